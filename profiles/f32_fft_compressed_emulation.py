"""CPU experiment: does the COMPRESSED analysis need float64 butterflies?

The lossless features need them (normalised real / imag of near-silent bins miss 1e-5 RMS with float32 butterflies,
DESIGN.md "Precision").  The compressed features are mel-warped projections of those rows (60 + 45 + 45 smooth
combinations of 2049 bins), which average the per-bin round-off.  This script runs the oracle's analysis with a float32
FFT (scipy.fft on float32 frames, pocketfft single precision: the same error class as a float32 radix-16 engine) and a
float64 FFT and compares the outputs of format_for_modelling.   python profiles/f32_fft_compressed_emulation.py
"""
import os
import sys

import numpy as np
import scipy.fft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import magphase_oracle as orc
from magphase_b200.synth import synth_utterance

rms = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)))
for fs, N, uid in ((48000, 4096, 3), (48000, 4096, 7), (16000, 2048, 5)):
    sig, pm, voi = synth_utterance(uid, fs, 2.0)
    frms, v_shift, _ = orc.analysis_frames(sig, pm, N)
    H = N // 2 + 1
    X64 = np.fft.fft(frms)[:, :H]
    X32 = scipy.fft.fft(frms.astype(np.float32), axis=1)[:, :H].astype(np.complex128)
    assert scipy.fft.fft(frms.astype(np.float32), axis=1).dtype == np.complex64
    f64 = orc.compute_lossless_feats(X64, v_shift, voi, fs)
    f32 = orc.compute_lossless_feats(X32, v_shift, voi, fs)
    print('fs %d N %d frames %d: lossless rms  mag %.2e  real %.2e  imag %.2e' % (fs, N, len(v_shift), rms(f32[0], f64[0]), rms(f32[1], f64[1]), rms(f32[2], f64[2])))
    c64 = orc.format_for_modelling(*f64, fs, mag_dim=60, phase_dim=45)
    c32 = orc.format_for_modelling(*f32, fs, mag_dim=60, phase_dim=45)
    print('    compressed rms  mag_mel_log %.2e  real_mel %.2e  imag_mel %.2e   (bar 1e-5)' % (rms(c32[0], c64[0]), rms(c32[1], c64[1]), rms(c32[2], c64[2])))
    print('    compressed max  mag_mel_log %.2e  real_mel %.2e  imag_mel %.2e' % (np.abs(c32[0] - c64[0]).max(), np.abs(c32[1] - c64[1]).max(), np.abs(c32[2] - c64[2]).max()))

# band-limited synthetic utterances (magphase_b200.synth.synth_utterance_band_limited): quiet high-frequency bins without
# any reference data; floor None = quantisation noise only, harder than a real recording
from magphase_b200.synth import synth_utterance_band_limited
for uid, fl in ((3, -72.0), (7, -72.0), (3, -85.0), (3, None)):
    sig, pm, voi = synth_utterance_band_limited(uid, 48000, 2.0, 7000.0, fl)
    N = 4096; H = N // 2 + 1
    frms, v_shift, _ = orc.analysis_frames(sig, pm, N)
    X64 = np.fft.fft(frms)[:, :H]
    X32 = scipy.fft.fft(frms.astype(np.float32), axis=1)[:, :H].astype(np.complex128)
    f64 = orc.compute_lossless_feats(X64, v_shift, voi, 48000)
    f32 = orc.compute_lossless_feats(X32, v_shift, voi, 48000)
    c64 = orc.format_for_modelling(*f64, 48000, mag_dim=60, phase_dim=45)
    c32 = orc.format_for_modelling(*f32, 48000, mag_dim=60, phase_dim=45)
    print('band-limited u%d floor %s dB: lossless rms real %.2e imag %.2e | compressed rms mag_mel_log %.2e real_mel %.2e imag_mel %.2e' % (
        uid, fl, rms(f32[1], f64[1]), rms(f32[2], f64[2]), rms(c32[0], c64[0]), rms(c32[1], c64[1]), rms(c32[2], c64[2])))

# the bundled natural recordings (only where /root/reference exists): quiet high-frequency bins make this the harder case
d = '/root/reference/demos/data_48k/wavs_nat'
if os.path.isdir(d):
    from scipy.io import wavfile
    from magphase_b200.synth import synth_marks_for_wav
    for name in sorted(os.listdir(d))[:4]:
        fs, x = wavfile.read(os.path.join(d, name))
        sig = x.astype(np.float64) / 32768.0
        pm, voi = synth_marks_for_wav(sig.size, fs)
        N = 4096; H = N // 2 + 1
        frms, v_shift, _ = orc.analysis_frames(sig, pm, N)
        X64 = np.fft.fft(frms)[:, :H]
        X32 = scipy.fft.fft(frms.astype(np.float32), axis=1)[:, :H].astype(np.complex128)
        f64 = orc.compute_lossless_feats(X64, v_shift, voi, fs)
        f32 = orc.compute_lossless_feats(X32, v_shift, voi, fs)
        c64 = orc.format_for_modelling(*f64, fs, mag_dim=60, phase_dim=45)
        c32 = orc.format_for_modelling(*f32, fs, mag_dim=60, phase_dim=45)
        print('%s frames %d: lossless rms real %.2e imag %.2e | compressed rms mag_mel_log %.2e real_mel %.2e imag_mel %.2e' % (
            name, len(v_shift), rms(f32[1], f64[1]), rms(f32[2], f64[2]), rms(c32[0], c64[0]), rms(c32[1], c64[1]), rms(c32[2], c64[2])))

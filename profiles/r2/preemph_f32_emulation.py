"""CPU experiment (round 2, not adopted): can float32 butterflies carry the compressed analysis if the windowed frame is
pre-emphasised in float64 first (y[n] = b[n] - a b[n-1], circular, so that FFT(b) = FFT(y) / (1 - a e^{-jw}) exactly)?
Flattening the spectrum lowers the white round-off floor of a float32 FFT relative to the quiet high-frequency bins.
Result with pocketfft in float32 standing in for the engine: mag_mel_log error of the natural recordings improves only
2.5-4x (hvd_577 3.7e-6 -> 0.9e-6 at a = 0.9), and LF-heavy synthetic speech gets WORSE (2.5e-7 -> 3.4e-6: the division by
|H| ~ 0.03 near DC amplifies the error where the mel axis is densest).  The GPU float32 engine measured 3-6x above this
emulation (profiles/r2/parity_report_logp_f32.txt), so the margin under the 1e-5 bar would be ~2x at best: float64 butterflies
stay.   python profiles/r2/preemph_f32_emulation.py   (needs /root/reference for the natural recordings)"""
import os, sys
import numpy as np, scipy.fft
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import magphase_oracle as orc
from magphase_b200.synth import synth_utterance, synth_marks_for_wav, synth_utterance_band_limited
from scipy.io import wavfile
rms = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)))

def feats(X, v_shift, voi, fs, pd=45):
    f = orc.compute_lossless_feats(X, v_shift, voi, fs)
    return orc.format_for_modelling(*f, fs, mag_dim=60, phase_dim=pd)

def run(name, sig, pm, voi, fs, N, a=0.97, pd=45, extra_noise=0.0):
    frms, v_shift, _ = orc.analysis_frames(sig, pm, N)
    H = N // 2 + 1
    X64 = np.fft.fft(frms)[:, :H]
    def f32fft(m):
        Y = scipy.fft.fft(m.astype(np.float32), axis=1)[:, :H].astype(np.complex128)
        if extra_noise:   # degrade to the error level of a less careful float32 engine: white, relative to the row RMS
            r = np.random.default_rng(0)
            s = extra_noise * np.sqrt(np.mean(np.abs(Y) ** 2, axis=1, keepdims=True))
            Y = Y + s * (r.standard_normal(Y.shape) + 1j * r.standard_normal(Y.shape))
        return Y
    X32 = f32fft(frms)
    y = frms - a * np.roll(frms, 1, axis=1)                    # circular first difference in float64
    Hk = 1.0 - a * np.exp(-2j * np.pi * np.arange(H) / N)
    Xpe = f32fft(y) / Hk
    c64 = feats(X64, v_shift, voi, fs, pd); c32 = feats(X32, v_shift, voi, fs, pd); cpe = feats(Xpe, v_shift, voi, fs, pd)
    print('%-28s f32: %.2e %.2e %.2e | pre-emph %.2f: %.2e %.2e %.2e' % ((name,) + tuple(rms(c32[i], c64[i]) for i in range(3)) + (a,) + tuple(rms(cpe[i], c64[i]) for i in range(3))))

d = '/root/reference/demos/data_48k/wavs_nat'
for noise in (0.0,):
    for name in ('hvd_593.wav', 'hvd_577.wav'):
        fs, x = wavfile.read(os.path.join(d, name)); sig = x.astype(np.float64) / 32768.0
        pm, voi = synth_marks_for_wav(sig.size, fs)
        for a in (0.9, 0.97):
            run(name, sig, pm, voi, fs, 4096, a=a)
        run(name + ' pd10', sig, pm, voi, fs, 4096, a=0.97, pd=10)
    for uid, fl in ((3, -72.0), (3, None)):
        sig, pm, voi = synth_utterance_band_limited(uid, 48000, 2.0, 7000.0, fl)
        run('bandlim u%d %s' % (uid, fl), sig, pm, voi, 48000, 4096)
    sig, pm, voi = synth_utterance(3, 48000, 2.0)
    run('synth u3', sig, pm, voi, 48000, 4096)

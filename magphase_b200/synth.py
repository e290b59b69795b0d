"""Deterministic synthetic speech + pitch marks ("synth48k-v1").

REAPER (the reference's pitch-mark provider, src/libaudio.py:450-455) is an external binary that
is not available, so benchmarks and parity tests generate their own utterances: alternating
voiced / unvoiced segments, voiced marks by integrating 1/f0(t), unvoiced marks every 5 ms
(REAPER flag ``-u 0.005``), f0 clamped to REAPER's ``-m 50 -x 400`` range.  The waveform is
quantised to int16 steps to mimic ``sf.read`` of the bundled PCM16 fixtures.
"""
import numpy as np
from scipy import signal


def synth_utterance(u, fs=48000, dur_s=5.0):
    """Returns (v_sig float64 [L], v_pm_smpls float64 [n], v_voi float64 [n]) for utterance id ``u``."""
    rng = np.random.Generator(np.random.PCG64(1000 + int(u)))
    L = int(round(dur_s * fs))
    unv_step = 0.005 * fs
    marks, voi = [], []
    pos, seg_end, voiced = 0.0, 0.05 * fs, False
    while pos < L - 2:
        if voiced:
            fb, r, ph = rng.uniform(90, 250), rng.uniform(0.5, 2.0), rng.uniform(0, 2 * np.pi)
            while pos < seg_end:
                f0 = min(max(fb * 2.0 ** (0.25 * np.sin(2 * np.pi * r * pos / fs + ph)), 50.0), 400.0)
                pos += fs / f0
                marks.append(pos)
                voi.append(1.0)
        else:
            while pos < seg_end:
                pos += unv_step
                marks.append(pos)
                voi.append(0.0)
        voiced = not voiced
        seg_end = pos + rng.uniform(0.15, 0.60) * fs
    pm = np.array(marks)
    vv = np.array(voi)
    keep = np.round(pm) < (L - 1)
    pm, vv = pm[keep], vv[keep]
    ipm = np.round(pm).astype(int)
    # glottal impulses through three resonators, peak 0.3, + white noise at -40 dB
    exc = np.zeros(L)
    exc[ipm[vv > 0]] = 1.0
    y = np.zeros(L)
    for fc, bw in ((700.0, 80.0), (1200.0, 100.0), (2600.0, 140.0)):
        rr = np.exp(-np.pi * bw / fs)
        y += signal.lfilter([1.0], [1.0, -2 * rr * np.cos(2 * np.pi * fc / fs), rr * rr], exc)
    y *= 0.3 / max(float(np.max(np.abs(y))), 1e-9)
    y += rng.uniform(-1, 1, L) * 0.003
    # unvoiced stretches (no voiced mark on either side): plain noise
    unv = np.ones(L, dtype=bool)
    ext = np.hstack((0, ipm, L - 1))
    for i in np.nonzero(vv > 0)[0]:
        unv[ext[i]:ext[i + 2] + 1] = False
    y[unv] = rng.uniform(-0.05, 0.05, int(unv.sum()))
    v_sig = np.clip(np.round(y * 32768.0), -32768, 32767) / 32768.0
    return v_sig, pm, vv


def synth_utterance_band_limited(u, fs=48000, dur_s=5.0, cutoff_hz=7000.0, floor_db=-72.0):
    """synth_utterance(u) through a steep low-pass plus a faint white floor, re-quantised to int16 steps: quiet
    high-frequency bins like those of a studio recording -- the hard case for anything computed in float32 (normalised
    real / imag of near-silent bins).  floor_db=None leaves only the quantisation noise (about -100 dB) above the cut-off,
    which is harder than any real recording.  Sensitivity to float32 butterflies, measured on the CPU with the oracle
    (profiles/f32_fft_compressed_emulation.py): see that script's output."""
    v_sig, pm, vv = synth_utterance(u, fs=fs, dur_s=dur_s)
    sos = signal.butter(10, cutoff_hz / (fs / 2.0), btype='lowpass', output='sos')
    y = signal.sosfilt(sos, v_sig)
    if floor_db is not None:
        rng = np.random.Generator(np.random.PCG64(9000 + int(u)))
        y = y + rng.uniform(-1, 1, y.size) * (10.0 ** (floor_db / 20.0))
    return np.clip(np.round(y * 32768.0), -32768, 32767) / 32768.0, pm, vv


def synth_marks_for_wav(n_smpls, fs=48000, seed=0):
    """Deterministic pitch marks + voicing for an arbitrary (e.g. bundled natural) waveform."""
    rng = np.random.Generator(np.random.PCG64(7000 + int(seed)))
    marks, voi = [], []
    pos, voiced = 0.0, False
    seg_end = 0.05 * fs
    while pos < n_smpls - 2:
        if voiced:
            fb, r, ph = rng.uniform(90, 250), rng.uniform(0.5, 2.0), rng.uniform(0, 2 * np.pi)
            while pos < seg_end:
                f0 = min(max(fb * 2.0 ** (0.25 * np.sin(2 * np.pi * r * pos / fs + ph)), 50.0), 400.0)
                pos += fs / f0
                marks.append(pos)
                voi.append(1.0)
        else:
            while pos < seg_end:
                pos += 0.005 * fs
                marks.append(pos)
                voi.append(0.0)
        voiced = not voiced
        seg_end = pos + rng.uniform(0.15, 0.60) * fs
    pm = np.array(marks)
    vv = np.array(voi)
    keep = np.round(pm) < (n_smpls - 1)
    return pm[keep], vv[keep]

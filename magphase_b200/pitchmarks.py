"""A pitch-mark provider of our own: the step BEFORE the hot path.

The reference shells out to the REAPER binary (``la.reaper``, src/libaudio.py:450-455, flags ``-s -x 400 -m 50 -a
-u 0.005``) and reads its ``.est`` file (``la.read_reaper_est_file``, :421-447): epoch times in seconds plus a voicing
flag, unvoiced stretches filled with marks every 5 ms.  REAPER is an external C++ program that is not available here; this
module produces the same two arrays for clean recordings and synthetic speech so that ``analysis_lossless(wav)`` runs
out of the box.  It is NOT a restatement of REAPER (a dynamic-programming epoch tracker) and makes no parity claim:

  1. frame-wise voicing + f0 from the normalised autocorrelation of 40 ms frames every 5 ms (lags 1/400 .. 1/50 s),
     smoothed by a median of five;
  2. inside a voiced stretch, marks are placed pitch-synchronously: the first one on the strongest negative-going or
     positive peak of the band-limited signal, every next one at the strongest peak of the same polarity within +-30 % of
     the local period after the previous mark;
  3. unvoiced stretches get marks every ``unv_step`` seconds, like REAPER's ``-u``.

Host NumPy on purpose: O(samples) work that runs once per utterance in front of the device pipeline.
"""
import numpy as np
from scipy import signal


def _frame_f0(x, fs, f0_min, f0_max, hop, win):
    """Normalised autocorrelation pitch per frame: (f0 [n], strength [n], rms [n])."""
    n_frm = max(1, (x.size - win) // hop + 1)
    lag_lo, lag_hi = int(fs / f0_max), int(fs / f0_min)
    nfft = 1 << int(np.ceil(np.log2(win + lag_hi + 1)))
    idx = np.arange(win)[None, :] + hop * np.arange(n_frm)[:, None]
    frames = x[np.minimum(idx, x.size - 1)] * np.hanning(win)[None, :]
    frames = frames - frames.mean(axis=1, keepdims=True)
    spec = np.fft.rfft(frames, nfft, axis=1)
    ac = np.fft.irfft(np.abs(spec) ** 2, nfft, axis=1)[:, :lag_hi + 1]
    # compensate the taper of the window's own autocorrelation
    wac = np.fft.irfft(np.abs(np.fft.rfft(np.hanning(win), nfft)) ** 2, nfft)[:lag_hi + 1]
    nac = ac / np.maximum(ac[:, :1], 1e-20) / np.maximum(wac / wac[0], 1e-3)[None, :]
    seg = nac[:, lag_lo:lag_hi + 1]
    best = np.argmax(seg, axis=1)
    strength = seg[np.arange(n_frm), best]
    f0 = fs / (best + lag_lo).astype(np.float64)
    rms = np.sqrt(np.mean(frames ** 2, axis=1))
    return f0, strength, rms


def estimate_pitch_marks(v_sig, fs, f0_min=50.0, f0_max=400.0, unv_step=0.005, voicing_threshold=0.5):
    """Returns (v_pm_sec, v_voi) in the layout of ``la.read_reaper_est_file``: strictly increasing mark times in seconds
    and a 0 / 1 voicing flag per mark."""
    x = np.asarray(v_sig, dtype=np.float64)
    n = x.size
    if n < int(0.05 * fs):
        t = np.arange(unv_step, n / float(fs), unv_step)
        return t, np.zeros(t.size)
    hop, win = int(round(0.005 * fs)), int(round(0.040 * fs))
    # band-limit to where voicing lives: peaks of this signal are the mark candidates
    sos = signal.butter(2, [40.0 / (fs / 2.0), min(1200.0, 0.45 * fs) / (fs / 2.0)], btype='bandpass', output='sos')
    y = signal.sosfiltfilt(sos, x)
    f0, strength, rms = _frame_f0(y, fs, f0_min, f0_max, hop, win)
    voiced = (strength > voicing_threshold) & (rms > 0.02 * max(float(rms.max()), 1e-12))
    voiced = signal.medfilt(voiced.astype(np.float64), 5) > 0.5
    f0 = signal.medfilt(f0, 5)
    centre = (np.arange(f0.size) * hop + win // 2).astype(np.int64)
    per_smp = np.interp(np.arange(n), centre, fs / np.maximum(f0, 1.0))          # local period in samples
    voi_smp = np.interp(np.arange(n), centre, voiced.astype(np.float64)) > 0.5
    marks, flags = [], []
    unv = int(round(unv_step * fs))
    min_gap = max(1, int(0.7 * fs / f0_max))                                       # = the lower end of the epoch search below
    pos = 0
    while pos < n - 1:
        if not voi_smp[pos]:
            nxt = pos + unv
            if nxt >= n - 1:
                break
            # an unvoiced mark, unless a voiced stretch starts before it
            start = pos + 1 + int(np.argmax(voi_smp[pos + 1:nxt + 1])) if voi_smp[pos + 1:nxt + 1].any() else -1
            if start < 0:
                marks.append(nxt); flags.append(0.0)
                pos = nxt
                continue
            pos = start
        # ---- a voiced stretch starts at pos: its end, its polarity, its first mark ----
        end = pos + int(np.argmin(voi_smp[pos:])) if not voi_smp[pos:].all() else n
        T0 = int(per_smp[pos])
        seg = y[pos:min(end, pos + 2 * T0 + 1)]
        if seg.size < 3:
            pos = end
            continue
        sign = 1.0 if seg.max() >= -seg.min() else -1.0
        m = max(1, pos + int(np.argmax(sign * seg[:T0 + 1])))          # never on sample 0: a zero-length first frame reads as f0 = inf
        # no two marks closer than the shortest period searched for: an unvoiced filler mark that would sit right in front
        # of the stretch's first epoch gives way to it (a one-sample frame would read as f0 = fs there)
        while marks and flags[-1] == 0.0 and m - marks[-1] < min_gap:
            marks.pop(); flags.pop()
        if marks and m - marks[-1] < min_gap:
            m = marks[-1] + min_gap
        while m < min(end, n - 1):
            marks.append(m); flags.append(1.0)
            T = per_smp[m]
            lo, hi = m + int(0.7 * T), m + int(1.3 * T) + 1
            if lo >= min(end, n - 1):
                break
            hi = min(hi, n)
            m = lo + int(np.argmax(sign * y[lo:hi]))
        pos = max(end, (marks[-1] + 1) if marks else end)
    pm = np.asarray(marks, dtype=np.float64)
    vv = np.asarray(flags, dtype=np.float64)
    keep = np.hstack((True, np.diff(pm) > 0)) if pm.size else np.zeros(0, dtype=bool)
    pm, vv = pm[keep], vv[keep]
    ok = pm < (n - 1)
    return pm[ok] / float(fs), vv[ok]

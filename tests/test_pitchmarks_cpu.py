"""The package's own pitch-mark provider (stands in for the REAPER binary the reference shells out
to, src/libaudio.py:450-455): same output layout as la.read_reaper_est_file, sane on synthetic speech with known marks.
No parity claim against REAPER (a different algorithm, binary absent)."""
import os

import numpy as np

from magphase_b200.pitchmarks import estimate_pitch_marks
from magphase_b200.synth import synth_utterance
from magphase_b200 import hostio


def test_marks_match_known_synthetic_marks():
    for u in (1, 7):
        sig, pm, voi = synth_utterance(u, fs=48000, dur_s=2.0)
        pm_s, vv = estimate_pitch_marks(sig, 48000)
        est = np.round(pm_s * 48000)
        assert np.all(np.diff(est) > 0) and est[-1] < sig.size - 1 and set(np.unique(vv)) <= {0.0, 1.0}
        grid = np.arange(0, sig.size, 48)
        track = lambda p, v: v[np.clip(np.searchsorted(p, grid), 0, p.size - 1)]
        assert np.mean(track(pm, voi) == track(est, vv)) > 0.95                 # voicing decision on a 1 ms grid
        # pitch-synchronous: the spacing of consecutive voiced marks follows the true local period (the marks themselves
        # may sit on another peak of the period than the synthetic excitation instants: a constant offset)
        both = (vv[1:] > 0) & (vv[:-1] > 0)
        per_est, at = np.diff(est)[both], est[1:][both]
        tv = pm[voi > 0]
        per_true = np.interp(at, tv[1:], np.diff(tv))
        inside = track(pm, voi)[np.clip(np.searchsorted(grid, at), 0, grid.size - 1)] > 0
        rel = np.abs(per_est - per_true)[inside] / per_true[inside]
        assert np.mean(rel < 0.08) > 0.9
        unv = np.diff(est)[(vv[1:] == 0) & (vv[:-1] == 0)]
        assert unv.size and np.mean(unv == 240) > 0.95                          # REAPER -u 0.005: 5 ms steps when unvoiced


def test_est_file_round_trip_and_short_signals(tmp_path):
    sig, pm, voi = synth_utterance(3, fs=16000, dur_s=1.0)
    pm_s, vv = estimate_pitch_marks(sig, 16000)
    f = os.path.join(tmp_path, 'a.est')
    hostio.write_reaper_est_file(f, pm_s, vv)
    pm2, vv2 = hostio.read_reaper_est_file(f, check_len_smpls=sig.size, fs=16000)
    assert np.allclose(pm2, pm_s, atol=1e-6) and np.array_equal(vv2, vv)
    t, v = estimate_pitch_marks(np.zeros(100), 48000)                           # shorter than one analysis window
    assert t.size == 0 or np.all(v == 0)
    t, v = estimate_pitch_marks(np.zeros(48000), 48000)                         # silence: unvoiced marks every 5 ms
    assert np.all(v == 0) and np.allclose(np.diff(t), 0.005)


def test_no_two_marks_closer_than_the_shortest_period():
    """An unvoiced filler mark right in front of a voiced stretch's first epoch used to leave frames of 1-40 samples (read
    as f0 = fs / 1 by shift_to_f0): consecutive marks now keep at least 0.7 / f0_max seconds, on synthetic speech and -- where
    the reference's recordings are present -- on natural speech."""
    cases = [synth_utterance(u, fs=48000, dur_s=1.5)[0] for u in (2, 5, 9)]
    d = '/root/reference/demos/data_48k/wavs_nat'
    if os.path.isdir(d):
        from scipy.io import wavfile
        for name in sorted(os.listdir(d))[:4]:
            cases.append(wavfile.read(os.path.join(d, name))[1].astype(np.float64) / 32768.0)
    for sig in cases:
        pm_s, vv = estimate_pitch_marks(sig, 48000)
        gaps = np.diff(np.round(pm_s * 48000))
        assert gaps.min() >= int(0.7 * 48000 / 400.0), gaps.min()
        assert gaps.max() <= 0.03 * 48000                               # never more than 1.3 periods of 50 Hz + one filler step
        assert 0.1 < vv.mean() < 0.9


def test_no_mark_on_sample_zero():
    """A recording that starts inside a voiced sound with its strongest peak on the very first sample: the first mark must not
    be sample 0 (shift 0 -> f0 = voi * fs / 0 in shift_to_f0, src/magphase.py:2198-2207)."""
    fs = 48000
    t = np.arange(int(0.4 * fs)) / fs
    sig = 0.5 * np.cos(2 * np.pi * 120.0 * t) + 0.25 * np.cos(2 * np.pi * 240.0 * t)      # peaks at t = 0, 1/120, ...
    pm_s, vv = estimate_pitch_marks(sig, fs)
    est = np.round(pm_s * fs)
    assert est.size > 20 and est[0] >= 1 and np.all(np.diff(est) > 0)
    assert vv.mean() > 0.8 and abs(np.median(fs / np.diff(est)[(vv[1:] > 0) & (vv[:-1] > 0)]) - 120.0) < 3.0

"""The vectorised batch bookkeeping (one pass of NumPy calls per batch) against the per-utterance definitions that
mirror the reference line by line.  Integer arrays must be identical, float arrays bit-identical.  No GPU needed."""
import numpy as np
import pytest

import magphase_b200.magphase as mp
from magphase_b200.synth import synth_utterance


def _utts():
    out = [synth_utterance(u, fs=48000, dur_s=d) for u, d in ((0, 0.7), (1, 1.3), (2, 0.4), (3, 2.0))]
    sig, pm, voi = synth_utterance(4, fs=48000, dur_s=0.5)
    out.append((sig, pm[:1], voi[:1]))                       # a single pitch mark
    return out


def test_batch_frame_geometry_equals_per_utterance():
    utts = _utts()
    pm, left, right, off = mp.batch_frame_geometry([u[1] for u in utts], [u[0].size for u in utts])
    for k, (sig, p, _) in enumerate(utts):
        P, v_shift, v_rights = mp.frame_geometry(p, sig.size)
        a, b = off[k], off[k + 1]
        assert np.array_equal(pm[a:b], P[1:-1])
        assert np.array_equal(left[a:b], v_shift)
        assert np.array_equal(right[a:b], v_rights)


def test_medfilt3_segments_equals_scipy_per_utterance():
    from scipy import signal
    rng = np.random.default_rng(0)
    segs = [rng.uniform(0, 300, n) * (rng.uniform(size=n) > 0.3) for n in (1, 2, 3, 50, 7)]
    off = np.concatenate(([0], np.cumsum([s.size for s in segs])))
    got = mp._medfilt3_segments(np.concatenate(segs), off)
    ref = np.concatenate([signal.medfilt(s) for s in segs])
    assert np.array_equal(got, ref)
    assert np.array_equal(ref, np.concatenate([mp.medfilt3(s) for s in segs]))


@pytest.mark.parametrize('b_voi_ap_win', [True, False])
def test_flat_synthesis_geometry_equals_loop(b_voi_ap_win):
    rng = np.random.default_rng(3)
    l_lf0 = []
    for n in (2, 3, 40, 333, 5):
        f0 = rng.uniform(60, 380, n)
        f0[rng.uniform(size=n) < 0.4] = 0.0
        with np.errstate(divide='ignore'):
            lf0 = np.log(f0)
        lf0[np.isinf(lf0)] = -1e10
        l_lf0.append(lf0)
    l_lf0.append(np.full(6, -1e10))                           # all unvoiced
    l_lf0.append(np.log(np.full(4, 48000.0 / 2047.4)))        # longest pitch period that still fits fft_len / 2
    rows = [v.size for v in l_lf0]
    for fs, fft_len in ((48000, 4096), (16000, 2048)):
        if fs == 16000:
            lf0s = [np.where(v > 0, v + np.log(1.2), v) for v in l_lf0[:-1]]
        else:
            lf0s = l_lf0
        r = [v.size for v in lf0s]
        a_flat, ns_flat = mp._compressed_synthesis_geometry_flat(lf0s, r, fs, fft_len, b_voi_ap_win)
        a_loop, ns_loop = mp._compressed_synthesis_geometry_loop(lf0s, r, fs, fft_len, b_voi_ap_win, False)
        assert ns_flat == ns_loop
        assert set(a_flat) == set(a_loop)
        for k in a_loop:
            if a_loop[k] is None:
                assert a_flat[k] is None
            else:
                assert a_flat[k].dtype == a_loop[k].dtype, k
                assert np.array_equal(a_flat[k], a_loop[k]), k


def test_flat_synthesis_geometry_errors():
    ok = np.log(np.full(5, 100.0))
    with pytest.raises(IndexError):
        mp.compressed_synthesis_geometry([ok, ok[:1]], [5, 1], 48000, 4096)
    with pytest.raises(ValueError):
        mp.compressed_synthesis_geometry([ok], [4], 48000, 4096)
    with pytest.raises(ValueError):                            # pitch period longer than fft_len / 2
        mp.compressed_synthesis_geometry([np.log(np.full(5, 20.0))], [5], 48000, 4096)


def test_c_analysis_geometry_equals_numpy_definitions():
    """mpb_analysis_geometry (one C pass) against frame_geometry / shift_to_f0 / _lf0_smoothed per utterance."""
    utts = _utts()
    fs = 48000
    centre, left, right, voi8, f0_med, off = mp._analysis_geometry_c([u[1] for u in utts], [u[0].size for u in utts],
                                                                     [u[2] for u in utts], fs)
    sig_off = 0
    for k, (sig, p, voi) in enumerate(utts):
        P, v_shift, v_rights = mp.frame_geometry(p, sig.size)
        a, b = off[k], off[k + 1]
        assert np.array_equal(centre[a:b], P[1:-1] + sig_off)
        assert np.array_equal(left[a:b], v_shift) and np.array_equal(right[a:b], v_rights)
        v_f0 = mp.shift_to_f0(v_shift.astype(int), np.asarray(voi, dtype=np.float64), fs, out='f0', b_smooth=False)
        v_voi, v_lf0 = mp._lf0_smoothed(v_f0)
        assert np.array_equal(voi8[a:b], (v_voi > 0).astype(np.uint8))
        assert np.array_equal(mp.f0_to_lf0(f0_med[a:b].copy()), v_lf0)            # bit-identical log argument
        sig_off += sig.size


def test_c_geometry_randomised_against_the_definitions():
    """Many random batches (seeded): the one-pass C bookkeeping must reproduce the per-utterance definitions exactly --
    integer arrays identical, float arrays bit-identical -- for both sides of the path."""
    rng = np.random.default_rng(2024)
    for trial in range(40):
        n_utt = int(rng.integers(1, 7))
        fs = int(rng.choice([48000, 16000]))
        fft_len = 4096 if fs == 48000 else 2048
        # ---- synthesis side: lf0 tracks with voiced / unvoiced stretches, periods up to the fft_len / 2 limit ----
        l_lf0 = []
        for _ in range(n_utt):
            n = int(rng.integers(2, 120))
            f0 = rng.uniform(fs / (fft_len / 2 - 2.0), 400.0, n)
            f0[rng.uniform(size=n) < rng.uniform(0, 0.8)] = 0.0
            with np.errstate(divide='ignore'):
                lf0 = np.log(f0)
            lf0[np.isinf(lf0)] = -1e10
            l_lf0.append(lf0)
        rows = [v.size for v in l_lf0]
        for vw in (True, False):
            a_flat, ns_flat = mp._compressed_synthesis_geometry_flat(l_lf0, rows, fs, fft_len, vw)
            a_loop, ns_loop = mp._compressed_synthesis_geometry_loop(l_lf0, rows, fs, fft_len, vw, False)
            assert ns_flat == ns_loop
            for k in a_loop:
                if a_loop[k] is None:
                    assert a_flat[k] is None
                else:
                    assert a_flat[k].dtype == a_loop[k].dtype and np.array_equal(a_flat[k], a_loop[k]), (trial, k)
        # ---- analysis side: jittered marks (including .5 positions: round half to even), random voicing ----
        l_pm, l_n, l_voi = [], [], []
        for _ in range(n_utt):
            n = int(rng.integers(1, 90))
            pm = np.cumsum(rng.integers(60, 900, n)).astype(np.float64) + rng.choice([0.0, 0.5, 0.25, -0.5], n)
            l_pm.append(pm)
            l_n.append(int(pm[-1]) + int(rng.integers(2, 500)))
            l_voi.append((rng.uniform(size=n) < 0.6).astype(np.float64))
        centre, left, right, voi8, f0_med, off = mp._analysis_geometry_c(l_pm, l_n, l_voi, fs)
        sig_off = 0
        for k in range(n_utt):
            P, v_shift, v_rights = mp.frame_geometry(l_pm[k], l_n[k])
            a, b = off[k], off[k + 1]
            assert np.array_equal(centre[a:b], P[1:-1] + sig_off)
            assert np.array_equal(left[a:b], v_shift) and np.array_equal(right[a:b], v_rights)
            v_f0 = mp.shift_to_f0(v_shift.astype(int), l_voi[k], fs, out='f0', b_smooth=False)
            v_voi, v_lf0 = mp._lf0_smoothed(v_f0)
            assert np.array_equal(voi8[a:b], (v_voi > 0).astype(np.uint8))
            assert np.array_equal(mp.f0_to_lf0(f0_med[a:b].copy()), v_lf0)
            sig_off += l_n[k]


def test_const_rate_reverse_scan_c_pass_is_bit_identical_to_the_numpy_loop():
    """mpb_const_rate_scan == get_shifts_and_frm_locs_from_const_shifts (np.interp per step, src/magphase.py:1426-1449):
    the shifts are truncated to integers afterwards and become pitch marks, so the floats must agree to the last bit --
    including exact knot hits, single-row utterances, all-unvoiced and all-voiced tracks, integer and fractional steps."""
    rng = np.random.Generator(np.random.PCG64(11))
    for fs in (16000, 22050, 44100, 48000):
        tracks = []
        for u in range(24):
            n_c = int(rng.integers(1, 700))
            kind = u % 4
            voiced = np.ones(n_c, bool) if kind == 0 else (np.zeros(n_c, bool) if kind == 1 else rng.random(n_c) < 0.5)
            f0 = np.where(voiced, rng.uniform(50.0, 420.0, n_c), 0.0)
            if kind == 3:
                f0 = np.where(voiced, fs / (0.005 * fs * rng.integers(1, 3, n_c)), 0.0)     # shifts that land exactly on knots
            tracks.append(mp.f0_to_shift(f0, fs))
        got = mp.const_rate_scan_batch(tracks, 5.0, fs)
        assert len(got) == len(tracks)
        for v, (s, loc) in zip(tracks, got):
            rs, rl = mp.get_shifts_and_frm_locs_from_const_shifts(v, 5.0, fs)
            assert s.dtype == rs.dtype == np.float64 and np.array_equal(rs, s) and np.array_equal(rl, loc)
    assert mp.const_rate_scan_batch([], 5.0, 16000) == []


def test_const_rate_synthesis_geometry_vectorised_equals_the_loop():
    """compressed_synthesis_geometry(b_const_rate=True): the batch version (C reverse scan + vectorised row pairs / voicing +
    mpb_syn_geometry) against the per-utterance loop that follows src/magphase.py:846-896 line by line -- every array
    identical in dtype and value, for integer and fractional constant-rate steps."""
    rng = np.random.Generator(np.random.PCG64(23))
    for fs, fft_len in ((16000, 2048), (48000, 4096), (44100, 4096), (22050, 2048)):
        for trial in range(4):
            l_lf0 = []
            for u in range(int(rng.integers(1, 8))):
                n_c = int(rng.integers(3, 400))
                voiced = rng.random(n_c) < (0.0 if trial == 0 and u == 0 else 0.6)
                f0 = np.where(voiced, rng.uniform(70.0, 380.0, n_c), 0.0)
                l_lf0.append(np.where(f0 > 0, np.log(np.maximum(f0, 1e-3)), -1.0e10))
            rows = [v.size for v in l_lf0]
            for vw in (True, False):
                a, ns_a = mp._compressed_synthesis_geometry_const_flat(l_lf0, rows, fs, fft_len, vw)
                b, ns_b = mp._compressed_synthesis_geometry_loop(l_lf0, rows, fs, fft_len, vw, True)
                assert ns_a == ns_b and set(a) == set(b)
                for k in b:
                    assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), (fs, trial, k)
    with pytest.raises(ValueError):
        mp._compressed_synthesis_geometry_const_flat([np.zeros(5)], [4], 16000, 2048)

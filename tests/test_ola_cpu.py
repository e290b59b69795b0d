"""Standalone ola() (src/magphase.py:34-62): the gather formula of k_ola_gather (mpb_extra.cu) replayed in NumPy against the
real reference's loop -- index math, the cut with Python slice semantics (pm[0] beyond frmlen/2 makes the first slice index
negative), odd frame lengths, repeated marks, the optional centred window.  The kernel itself:
tests/test_gpu_lossless.py::test_standalone_ola."""
import numpy as np
import pytest

import magphase_b200.magphase as mpb


def gather_replay(m_frm, pm32, t0, n_out):
    """out[j] = sum_i (in frame order) frames[i][j + t0 - pm[i] + frmlen//2], exactly as the kernel walks it."""
    nfrm, frmlen = m_frm.shape
    half = frmlen // 2
    out = np.zeros(n_out)
    for j in range(n_out):
        pos = j + t0
        lo_val = pos + half - frmlen
        a = int(np.searchsorted(pm32, lo_val, side='right'))          # first i with pm[i] > lo_val
        acc = 0.0
        i = a
        while i < nfrm and pm32[i] <= pos + half:
            acc += m_frm[i, pos - int(pm32[i]) + half]
            i += 1
        out[j] = acc
    return out


CASES = [
    dict(frmlen=64, pm=[10, 25, 47, 60, 90, 91, 130]),                 # ordinary
    dict(frmlen=64, pm=[40, 60, 100, 180]),                            # pm[0] > frmlen/2: negative first slice index
    dict(frmlen=63, pm=[5, 20, 20, 44, 70]),                           # odd length, a repeated mark
    dict(frmlen=32, pm=[3, 100, 260]),                                 # gaps wider than a frame
    dict(frmlen=128, pm=[200]),                                        # a single frame, far from the origin
    dict(frmlen=16, pm=[0, 1, 2, 3, 9]),                               # mark at 0, one-sample shifts
]


@pytest.mark.parametrize('case', CASES)
def test_gather_replay_equals_reference_ola(ref_modules, case):
    mp, la, lu = ref_modules
    rng = np.random.default_rng(3)
    pm = np.array(case['pm'], dtype=np.float64) + 0.4                  # truncation, not rounding (:36)
    m_frm = rng.standard_normal((pm.size, case['frmlen']))
    ref = mp.ola(m_frm.copy(), pm.copy())
    pm32, t0, n_out = mpb.ola_geometry(pm, case['frmlen'])
    assert n_out == ref.size
    got = gather_replay(m_frm, pm32, t0, n_out)
    np.testing.assert_array_equal(got, ref)                            # same adds in the same order: bit-identical


def test_centred_window_weights_equal_reference(ref_modules):
    """The host half of ola(win_func=...): la.gen_centr_win per frame from window_weights."""
    mp, la, lu = ref_modules
    pm = np.array([30, 70, 95, 150, 230])
    frmlen = 256
    shift = np.append(np.diff(np.hstack((0, pm))), 0)
    shift[-1] = shift[-2]
    w_all, off = mpb.window_weights([mpb.raised_hanning] * pm.size, shift[:-1], shift[1:])
    for i in range(pm.size):
        ref = la.gen_centr_win(shift[i], shift[i + 1], frmlen, win_func=mp.raised_hanning)
        v_win = np.zeros(frmlen)
        z = frmlen // 2 - int(shift[i])
        v_win[z:z + int(off[i + 1] - off[i])] = w_all[off[i]:off[i + 1]]
        np.testing.assert_array_equal(v_win, ref)
    np.testing.assert_array_equal(mpb.raised_hanning(9, 0.7), mp.raised_hanning(9, 0.7))


@pytest.mark.parametrize('case', CASES)
def test_oracle_ola_equals_reference(ref_modules, case):
    import magphase_oracle as orc
    mp, la, lu = ref_modules
    rng = np.random.default_rng(4)
    pm = np.array(case['pm'], dtype=np.float64) + 0.4
    m_frm = rng.standard_normal((pm.size, case['frmlen']))
    np.testing.assert_array_equal(orc.ola(m_frm.copy(), pm), mp.ola(m_frm.copy(), pm.copy()))

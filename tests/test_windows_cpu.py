"""Host half of the arbitrary-window path (win_func callables, src/magphase.py:102-108, src/libaudio.py:70-84): the
mirror evaluates the callables, multiplies them into the frames' samples and hands the kernels a buffer of pre-windowed
frames with MPB_WIN_RECT.  Here the kernel's frame staging (mpb_frame.cuh:load_frame, weight 1) is replayed in NumPy on
that buffer and compared with the oracle's frames; the GPU side is tests/test_gpu_lossless.py::test_arbitrary_win_func."""
import warnings

import numpy as np
import pytest

import magphase_oracle as orc
import magphase_b200.magphase as mp


def _offpeak_window(n):
    return 0.9 * np.hamming(n) ** 1.5


def _replay_rect_staging(pre, centre, left, right, N):
    """b[N-j] = sig[c-j] (j = 1..l, priority), b[k] = sig[c+k] (k = 0..min(q, N-l-1)); l >= N: b[k] = sig[c-l+k]."""
    m = np.zeros((centre.size, N))
    for f in range(centre.size):
        c, l, q = int(centre[f]), int(left[f]), int(right[f])
        if l >= N:
            m[f] = pre[c - l:c - l + N]
            continue
        q_eff = min(q, N - l - 1)
        m[f, :q_eff + 1] = pre[c:c + q_eff + 1]
        if l:
            m[f, N - l:] = pre[c - l:c]
    return m


@pytest.mark.parametrize('win', ['hamming', 'offpeak', 'list'])
def test_prewindowed_frames_replay_equals_oracle_frames(win):
    rng = np.random.default_rng(12)
    sig = rng.uniform(-1, 1, 30000)
    pm = np.concatenate(([0.0], np.cumsum(rng.integers(150, 700, 24)).astype(float), [16000.0, 21000.5, 29000.0]))
    assert np.all(np.diff(pm) > 0)
    n = pm.size
    fn = {'hamming': np.hamming, 'offpeak': _offpeak_window,
          'list': [(np.hanning, np.hamming, _offpeak_window, mp.voi_noise_window)[f % 4] for f in range(n)]}[win]
    N = 4096
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref, v_shift, P = orc.analysis_frames(sig, pm, N, kinds=mp._win_list(fn, n))
    P2, left, right = mp.frame_geometry(pm, sig.size)
    assert np.array_equal(P, P2)
    pre, centre, idx, w_all = mp.prewindowed_frames(sig, P[1:-1], left, right, mp._win_list(fn, n))
    assert pre.size == int(np.sum(left + right + 1)) and idx.min() >= 0 and idx.max() < sig.size
    got = _replay_rect_staging(pre, centre, left, right, N)
    np.testing.assert_array_equal(got, ref)


def test_window_classification_and_errors():
    assert not mp._has_custom_window(np.hanning)
    assert not mp._has_custom_window([np.hanning, mp.voi_noise_window, 'hann', 'bartlett2.5'])
    assert mp._has_custom_window(np.hamming)
    assert mp._has_custom_window([np.hanning, np.hamming])
    assert mp._win_codes(np.hanning, 5) is None
    assert mp._win_codes([np.hanning, mp.voi_noise_window], 2).tolist() == [0, 1]
    with pytest.raises(ValueError):
        mp._win_codes([np.hanning], 2)                        # list length != number of frames
    with pytest.raises(ValueError):
        mp._has_custom_window(3.5)                            # not a window function
    with pytest.raises(ValueError):
        mp.window_weights([lambda n: np.ones(n + 1)], [3], [4])   # wrong length returned by the callable
    # hanning(1) = [1.0] for zero-length sides (src/libaudio.py:72-78 with left_len = 0)
    w, off = mp.window_weights([np.hanning, np.hamming], [0, 2], [3, 0])
    assert off.tolist() == [0, 4, 7]
    np.testing.assert_allclose(w[:4], np.concatenate(([1.0], np.hanning(7)[:4][::-1][1:])))
    np.testing.assert_allclose(w[4:], np.hamming(5)[:3])


def test_custom_window_path_rejects_marks_outside_the_signal():
    """The host half validates the geometry before it gathers (the closed-form path leaves that to the C entry point)."""
    sig = np.zeros(1000)
    for pm in ([100.0, 50.0, 300.0], [100.0, 400.0, 2000.0]):
        with pytest.raises(ValueError):
            mp.analysis_with_del_comp_from_pm(sig, 48000, np.array(pm), win_func=np.hamming)

"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header declares,
fails loudly without a GPU, and the host bookkeeping is bit-exact against the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import magphase_oracle as orc
from conftest import ROOT, have_cuda


def _header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'magphase_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(mpb_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    from magphase_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
    # and the ctypes table binds exactly the declared entry points
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_strings():
    from magphase_b200 import _lib
    assert b'magphase_b200' in _lib.lib().mpb_version()


@pytest.mark.skipif(have_cuda(), reason='only meaningful without a GPU')
def test_no_cpu_fallback():
    import magphase_b200.magphase as mp
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        mp.analysis_lossless_from_pm(np.zeros(2000), 48000, np.array([300.0, 700.0]), np.array([0.0, 1.0]))


def test_plan_ola_runs_host():
    from magphase_b200 import _lib
    rng = np.random.default_rng(0)
    pm_a = np.cumsum(rng.integers(120, 900, 500)).astype(np.int32)
    pm_b = np.cumsum(rng.integers(120, 900, 7)).astype(np.int32)
    pm = np.concatenate((pm_a, pm_b))
    off = np.array([0, 500, 507], dtype=np.int64)
    n = ctypes.c_int64()
    lib = _lib.lib()
    _lib.check(lib.mpb_plan_ola_runs(_lib.ptr(pm), _lib.ptr(off), 2, 4096, 16, None, 0, ctypes.byref(n)))
    runs = np.zeros((n.value, 4), dtype=np.int32)
    _lib.check(lib.mpb_plan_ola_runs(_lib.ptr(pm), _lib.ptr(off), 2, 4096, 16, _lib.ptr(runs), n.value, ctypes.byref(n)))
    # runs tile the frames of each utterance in order; every run of a multi-run utterance spans >= fft_len
    assert runs[0, 0] == 0 and np.all(runs[1:, 0] == runs[:-1, 0] + runs[:-1, 1])
    assert runs[-1, 0] + runs[-1, 1] == 507
    for first, cnt, utt, flags in runs:
        assert off[utt] <= first and first + cnt <= off[utt + 1]
        if flags != 0:
            assert pm[first + cnt - 1] - pm[first] >= 4096
        assert bool(flags & 1) == (first > off[utt]) and bool(flags & 2) == (first + cnt < off[utt + 1])
    assert runs[-1, 2] == 1 and runs[-1, 3] == 0
    # non-increasing marks are refused
    bad = pm.copy()
    bad[10] = bad[9]
    with pytest.raises(ValueError):
        _lib.check(lib.mpb_plan_ola_runs(_lib.ptr(bad), _lib.ptr(off), 2, 4096, 16, None, 0, ctypes.byref(n)))


def test_host_bookkeeping_bit_exact():
    import magphase_b200.magphase as mp
    rng = np.random.default_rng(1)
    pm = np.cumsum(rng.uniform(100, 700, 300))
    pm[5] = np.floor(pm[5]) + 0.5      # half-to-even cases
    pm[6] = np.floor(pm[6]) + 0.5
    P, s, r = mp.frame_geometry(pm, int(pm[-1]) + 500)
    P2, s2, r2 = orc.frame_limits(pm, int(pm[-1]) + 500)
    assert np.array_equal(P, P2) and np.array_equal(s, s2) and np.array_equal(r, r2)
    voi = (rng.uniform(size=300) > 0.4).astype(float)
    f0 = mp.shift_to_f0(s, voi, 48000, b_smooth=False)
    assert np.array_equal(f0, orc.shift_to_f0(s2, voi, 48000))
    assert np.array_equal(mp.f0_to_shift(f0, 48000), orc.f0_to_shift(f0, 48000))
    # ola geometry against the oracle's literal slicing
    v_pm = np.cumsum(mp.f0_to_shift(f0, 48000))
    pm_int, t0, n_out = mp.ola_geometry(v_pm, 4096)
    y = orc.ola(np.zeros((300, 4096)), v_pm)
    assert t0 == 0 and n_out == y.size and np.array_equal(pm_int, v_pm.astype(int))
    for fs in (16000, 22050, 44100, 48000):
        assert mp.define_alpha(fs) == orc.define_alpha(fs)
        assert mp.define_fft_len(fs) == orc.define_fft_len(fs)
    with pytest.raises(ValueError):
        mp.define_alpha(8000)


def test_const_rate_rows_reproduce_interp1d():
    """Index form of interp_from_variable_to_const_frm_rate against the oracle's scipy interp1d, with and without
    the replicated first row (pm[0] > 0 / pm[0] == 0)."""
    import magphase_b200.magphase as mp
    rng = np.random.default_rng(2)
    for first in (0, 137):
        pm = np.cumsum(np.r_[first, rng.integers(120, 900, 60)]).astype(np.int64)
        data = rng.normal(size=(pm.size, 5))
        r0, r1, w = mp.const_rate_rows(pm, 5.0, 48000)
        got = data[r0] + (data[r1] - data[r0]) * w[:, None]
        ref = orc.interp_from_variable_to_const_frm_rate(data, pm, 5.0, 48000)
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


def test_const_rate_scan_argument_checks_and_edge_cases():
    """mpb_const_rate_scan (host only): bad arguments are refused with MPB_ERR_BAD_ARG -> ValueError; an empty utterance
    writes nothing; a track with non-finite shifts (f0 = 0 gives fs / 0 upstream only if the caller skipped f0_to_shift's
    floor) behaves like the NumPy loop: NaN is handed through until the iteration cap."""
    from magphase_b200 import _lib
    lib = _lib.lib()
    sh = np.array([80.0, 90.0, 100.0, 0.0, 70.0], dtype=np.float64)
    off = np.array([0, 3, 3, 5], dtype=np.int64)                      # utterance 1 is empty
    o_s, o_l = np.full(12, -1.0), np.full(12, -1.0)
    cnt = np.zeros(4, dtype=np.int64)
    _lib.check(lib.mpb_const_rate_scan(_lib.ptr(sh), _lib.ptr(off), 3, 80.0, _lib.ptr(o_s), _lib.ptr(o_l), _lib.ptr(cnt)))
    assert cnt[1] == 0 and cnt[0] >= 1 and cnt[0] <= 5
    assert o_l[0] == 240.0 and o_s[0] == 100.0                         # the walk starts on the last centre with its own shift
    # utterance 2: a zero shift at its first row would stall the walk on one position; the iteration cap (2 n - 1) ends it
    assert 1 <= cnt[2] <= 3 and o_l[6] == 160.0 and o_s[6] == 70.0
    with pytest.raises(ValueError):
        _lib.check(lib.mpb_const_rate_scan(_lib.ptr(sh), _lib.ptr(off), 3, 0.0, _lib.ptr(o_s), _lib.ptr(o_l), _lib.ptr(cnt)))
    with pytest.raises(ValueError):
        _lib.check(lib.mpb_const_rate_scan(None, _lib.ptr(off), 3, 80.0, _lib.ptr(o_s), _lib.ptr(o_l), _lib.ptr(cnt)))
    bad = np.array([0, 3, 2, 5], dtype=np.int64)
    with pytest.raises(ValueError):
        _lib.check(lib.mpb_const_rate_scan(_lib.ptr(sh), _lib.ptr(bad), 3, 80.0, _lib.ptr(o_s), _lib.ptr(o_l), _lib.ptr(cnt)))
    nan = np.array([80.0, np.nan, 100.0], dtype=np.float64)
    _lib.check(lib.mpb_const_rate_scan(_lib.ptr(nan), _lib.ptr(np.array([0, 3], dtype=np.int64)), 1, 80.0, _lib.ptr(o_s),
                                       _lib.ptr(o_l), _lib.ptr(cnt)))
    import magphase_b200.magphase as mp
    rs, rl = mp.get_shifts_and_frm_locs_from_const_shifts(nan, 5.0, 16000)      # the NumPy loop: NaN is handed through to the cap
    assert cnt[0] == rs.size == 5
    assert np.array_equal(o_s[:5][::-1], rs, equal_nan=True) and np.array_equal(o_l[:5][::-1], rl, equal_nan=True)


def test_fork_after_initialisation_is_refused(monkeypatch):
    """the reference fans out with a forking Pool (src/libutils.py:32-63); a CUDA context does not survive
    fork(), so a child that inherited an initialised library must get an error, not a hang."""
    import os
    from magphase_b200 import _lib
    monkeypatch.setattr(_lib, '_ctx', {(0, 0): object()})        # as if the parent had created a context
    monkeypatch.setattr(_lib, '_ctx_pid', os.getpid())
    assert _lib.ctx(0) is _lib._ctx[(0, 0)]                      # the parent itself keeps working
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:                                                 # child
        try:
            _lib.ctx(0)
            os.write(w, b'no error')
        except RuntimeError as e:
            os.write(w, b'refused' if 'fork' in str(e) else b'other')
        finally:
            os._exit(0)
    os.close(w)
    os.waitpid(pid, 0)
    assert os.read(r, 64) == b'refused'
    os.close(r)

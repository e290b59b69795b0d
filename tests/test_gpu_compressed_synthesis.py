"""GPU parity of synthesis_from_compressed against the oracle and the golden vectors generated from the real
reference, with identical host-generated noise (np.random legacy stream, same seed)."""
import os
import warnings

import numpy as np
import pytest

import magphase_oracle as orc
from magphase_b200.synth import synth_utterance

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = 1e-5


def rms(a, b):
    return float(np.sqrt(np.mean(np.abs(np.asarray(a) - np.asarray(b)) ** 2)))


@pytest.fixture(scope='module')
def mp():
    import magphase_b200.magphase as m
    return m


@pytest.fixture(scope='module')
def gold():
    g = np.load(os.path.join(GOLD, 'compressed_hvd704.npz'))
    f64 = lambda k: g[k].astype(np.float64)
    return g, (f64('mag'), f64('real'), f64('imag'), f64('lf0'))


@pytest.mark.parametrize('key,kw', [('syn_var_nohpf', dict(b_out_hpf=False)), ('syn_var_hpf', dict(b_out_hpf=True)),
                                    ('syn_const_nohpf', dict(b_out_hpf=False, b_const_rate=True))])
def test_vs_reference_golden_48k(mp, gold, key, kw):
    g, feats = gold
    np.random.seed(int(g['seed']))
    y = mp.synthesis_from_compressed(*feats, 48000, **kw)
    assert y.shape == g[key].shape and y.dtype == np.float64
    assert rms(y, g[key]) < TOL, rms(y, g[key])
    assert np.max(np.abs(y - g[key])) < 1e-4


def test_vs_reference_golden_16k_const_rate_hpf(mp, gold):
    g, feats = gold
    np.random.seed(int(g['seed']))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y = mp.synthesis_from_compressed(*feats, 16000, b_const_rate=True)
    assert y.shape == g['syn_16k_const_hpf'].shape
    assert rms(y, g['syn_16k_const_hpf']) < TOL


def test_linear_phase_and_plain_noise_window_vs_oracle(mp, gold):
    g, feats = gold
    for kw in (dict(per_phase_type='linear'), dict(b_voi_ap_win=False)):
        np.random.seed(5)
        y = mp.synthesis_from_compressed(*feats, 48000, b_out_hpf=False, **kw)
        np.random.seed(5)
        y_ref = orc.synthesis_from_compressed(*feats, 48000, b_out_hpf=False, **kw)
        assert y.shape == y_ref.shape and rms(y, y_ref) < TOL


def test_copy_synthesis_low_dim_chain(mp):
    """BASELINE config 2: analysis_compressed (60/45/45) -> synthesis_from_compressed on the same utterance,
    CUDA chain against the oracle chain."""
    sig, pm, voi = synth_utterance(12, fs=48000, dur_s=1.0)
    got = mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    ref = orc.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    np.random.seed(3)
    y = mp.synthesis_from_compressed(got[0], got[1], got[2], got[3], 48000, b_out_hpf=False)
    np.random.seed(3)
    y_ref = orc.synthesis_from_compressed(ref[0], ref[1], ref[2], ref[3], 48000, b_out_hpf=False)
    assert y.shape == y_ref.shape
    assert rms(y, y_ref) < TOL, rms(y, y_ref)


def test_batch_equals_single_and_is_deterministic(mp, gold):
    g, feats = gold
    a_feats = tuple(f[:30] for f in feats)
    b_feats = tuple(f[20:64] for f in feats)
    rng = np.random.RandomState(1)
    lens = []
    for f in (a_feats, b_feats):
        sh = mp.f0_to_shift(np.exp(f[3]), 48000).astype(int)
        pmv = np.cumsum(sh)
        lens.append(int(pmv[-1] + (pmv[-1] - pmv[-2])))
    noises = [rng.uniform(-1, 1, n) for n in lens]
    y = mp.synthesis_from_compressed_batch([a_feats, b_feats], 48000, b_out_hpf=False, l_noise=noises)
    y2 = mp.synthesis_from_compressed_batch([a_feats, b_feats], 48000, b_out_hpf=False, l_noise=noises)
    for u, f in enumerate((a_feats, b_feats)):
        assert np.array_equal(y[u], y2[u])
        single = mp.synthesis_from_compressed_batch([f], 48000, b_out_hpf=False, l_noise=[noises[u]])[0]
        assert np.array_equal(single, y[u])
        ref = orc.synthesis_from_compressed(*f, 48000, b_out_hpf=False, v_noise=noises[u])
        assert rms(y[u], ref) < TOL


def test_all_unvoiced_and_all_voiced(mp, gold):
    """Empty voicing class: the reference takes np.mean of an empty selection (NaN gain, RuntimeWarning) that is
    never applied to any frame."""
    g, feats = gold
    mag, real, imag, lf0 = (f[:20].copy() for f in feats)
    for lf in (np.full(20, -1.0e10), np.full(20, np.log(120.0))):
        np.random.seed(2)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            y = mp.synthesis_from_compressed(mag, real, imag, lf, 48000, b_out_hpf=False)
            np.random.seed(2)
            y_ref = orc.synthesis_from_compressed(mag, real, imag, lf, 48000, b_out_hpf=False)
        assert np.all(np.isfinite(y)) and y.shape == y_ref.shape and rms(y, y_ref) < TOL


def test_min_phase_vs_reference_golden(mp, gold):
    g, feats = gold
    np.random.seed(int(g['seed']))
    y = mp.synthesis_from_compressed(*feats, 48000, b_out_hpf=False, per_phase_type='min_phase')
    assert y.shape == g['syn_minph_nohpf'].shape
    assert rms(y, g['syn_minph_nohpf']) < TOL


def test_min_phase_with_const_rate_vs_oracle(mp, gold):
    """min-phase of the INTERPOLATED magnitude rows (the oracle is pinned to the reference for this combination in
    tests/test_oracle_vs_ref.py), 48 kHz and 16 kHz."""
    g, feats = gold
    for fs in (48000, 16000):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            np.random.seed(21)
            y_ref = orc.synthesis_from_compressed(*feats, fs, b_const_rate=True, per_phase_type='min_phase', b_out_hpf=False)
            np.random.seed(21)
            y = mp.synthesis_from_compressed(*feats, fs, b_const_rate=True, per_phase_type='min_phase', b_out_hpf=False)
        assert y.shape == y_ref.shape and rms(y, y_ref) < TOL, (fs, rms(y, y_ref))


def test_explicit_fft_len_1024_and_2048_at_16k(mp, gold):
    """fft_len is an argument of synthesis_from_compressed (src/magphase.py:825): the smallest engine size (1024 points,
    32 threads per frame) and the default of 16 kHz, variable and constant rate."""
    g, feats = gold
    for fft_len in (1024, 2048):
        for kw in (dict(), dict(b_const_rate=True)):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                np.random.seed(9)
                y_ref = orc.synthesis_from_compressed(*feats, 16000, fft_len=fft_len, b_out_hpf=False, **kw)
                np.random.seed(9)
                y = mp.synthesis_from_compressed(*feats, 16000, fft_len=fft_len, b_out_hpf=False, **kw)
            assert y.shape == y_ref.shape and rms(y, y_ref) < TOL, (fft_len, kw, rms(y, y_ref))


def test_post_filter_vs_reference_golden(mp, gold):
    g, feats = gold
    y = mp.post_filter(feats[0], 48000)
    assert y.dtype == np.float64 and np.max(np.abs(y - g['post_filter_48k'])) < 1e-12
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        y16 = mp.post_filter(feats[0], 16000)
        assert any('16kHz' in str(x.message) for x in w)
    assert np.max(np.abs(y16 - g['post_filter_16k'])) < 1e-12
    # explicit options and another dimension
    rng = np.random.default_rng(4)
    x = rng.normal(size=(37, 80))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        a = mp.post_filter(x, 22050, av_len_at_zero=9, av_len_at_nyq=5, boost_at_zero=1.5, boost_at_nyq=2.5)
        b = orc.post_filter(x, 22050, av_len_at_zero=9, av_len_at_nyq=5, boost_at_zero=1.5, boost_at_nyq=2.5)
    assert np.max(np.abs(a - b)) < 1e-12
    with pytest.raises(ValueError):
        mp.post_filter(x, 22050)


def test_build_min_phase_vs_reference_golden(mp):
    g = np.load(os.path.join(GOLD, 'lossless_synth48k.npz'))
    got = mp.build_min_phase_from_mag_spec(g['mag_rows'])
    ref = g['minph_rows']
    assert got.dtype == np.complex128 and got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-12)) < 1e-9
    # |min-phase| reproduces the magnitude (property, any size); a zero bin hits the protected log
    rng = np.random.default_rng(3)
    mag = np.exp(rng.normal(size=(300, 1025)))
    mag[5, 17] = 0.0
    mpz = mp.build_min_phase_from_mag_spec(mag)
    ok = np.ones(300, dtype=bool); ok[5] = False
    assert np.max(np.abs(np.abs(mpz[ok]) / mag[ok] - 1)) < 1e-10
    # (the -1e10 floor puts +-1e10/N terms into the cepstrum: float64 round-off alone is ~1e-6 relative here)
    np.testing.assert_allclose(mpz[5], orc.build_min_phase_from_mag_spec(mag[5:6])[0], rtol=1e-4, atol=1e-300)


def test_errors(mp, gold):
    g, feats = gold
    with pytest.raises(ValueError):
        mp.synthesis_from_compressed(*feats, 48000, per_phase_type='bogus')
    with pytest.raises(ValueError):
        mp.synthesis_from_compressed(*feats, 48000, b_fbank_mel=True)
    with pytest.raises(ValueError):        # f0 of 20 Hz: frames longer than fft_len/2
        mp.synthesis_from_compressed(feats[0][:5], feats[1][:5], feats[2][:5], np.full(5, np.log(20.0)), 48000)


def test_numpy_legacy_stream_on_device_is_bit_exact(mp):
    """np.random.uniform on the global MT19937 stream, reproduced on the GPU: same numbers, same final state,
    across block boundaries (624-word regenerations), odd positions and repeated calls."""
    for seed, sizes in ((0, [1, 5, 311, 312, 313, 100000]), (12345, [7, 623, 1, 2_000_003])):
        np.random.seed(seed)
        np.random.uniform(size=3)                 # odd position inside the first block
        ref = [np.random.uniform(-1, 1, n) for n in sizes]
        ref_next = np.random.random_sample(5)
        np.random.seed(seed)
        np.random.uniform(size=3)
        got = [mp.numpy_stream_uniform(-1, 1, n) for n in sizes]
        got_next = np.random.random_sample(5)     # NumPy continues where the device left the stream
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
        assert np.array_equal(got_next, ref_next)


def test_numpy_legacy_stream_arbitrary_range(mp):
    """low + (high - low) * r with NumPy's two roundings (product, then sum): a fused multiply-add differs in the last bit
    unless the scale is a power of two."""
    np.random.seed(31)
    ref = np.random.uniform(0.1, 0.7, 200_001)
    np.random.seed(31)
    got = mp.numpy_stream_uniform(0.1, 0.7, 200_001)
    assert np.array_equal(got, ref)


def test_numpy_legacy_stream_jump_ahead_segments(mp):
    """Draws longer than one segment (1024 twists = 638,976 words = 319,488 draws) are generated by several CTAs that JUMP
    to their segment through x^(s J) mod phi (one fold; a second one beyond 256 segments).  Same numbers, same final state
    as NumPy, from fresh seeds (arbitrary low bits in key[0]), mid-block positions and pos = 624."""
    cases = ((7, 0, [319_488, 319_489, 1, 400_001]),         # exactly one segment, then one word more
             (2024, 5, [3_000_001, 11, 2_500_000]),
             (99, 0, [81_788_929 + 1000, 17]))                # beyond segment 256: the second-digit polynomials
    for seed, skip, sizes in cases:
        np.random.seed(seed)
        np.random.uniform(size=skip)
        ref = [np.random.uniform(-1, 1, n) for n in sizes]
        ref_state = np.random.get_state()
        np.random.seed(seed)
        np.random.uniform(size=skip)
        got = [mp.numpy_stream_uniform(-1, 1, n) for n in sizes]
        got_state = np.random.get_state()
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
        assert got_state[2] == ref_state[2]
        # NumPy's own key[0] keeps stale low bits that are not part of the state; compare what matters
        assert np.array_equal(got_state[1][1:], ref_state[1][1:]) and (got_state[1][0] >> 31) == (ref_state[1][0] >> 31)
        nxt = np.random.random_sample(700)                   # NumPy continues from the device's state, across a twist
        np.random.set_state(ref_state)
        assert np.array_equal(nxt, np.random.random_sample(700))


def test_device_hpf_matches_lfilter(mp):
    """The blocked state-space scan against scipy.signal.lfilter (the reference's own call, src/magphase.py:995):
    several utterances of odd lengths in one call.  lfilter's direct form carries ~1e-8 of rounding noise itself."""
    import ctypes
    from scipy import signal
    from magphase_b200 import _lib
    rng = np.random.default_rng(8)
    lens = [1, 511, 512, 513, 40000, 123457]
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    x = rng.normal(size=int(off[-1])) * 0.3 + 0.05          # with a DC offset for the high-pass to remove
    for fs in (48000, 16000):
        b, a = mp.output_hpf_coefficients(fs)
        sos = mp.output_hpf_sos(fs)
        y = x.copy()
        _lib.check(_lib.lib().mpb_sos2_host(_lib.ctx(), _lib.ptr(y), _lib.ptr(off), len(lens), _lib.ptr(sos)))
        for u in range(len(lens)):
            ref = signal.lfilter(b, a, x[off[u]:off[u + 1]])
            assert np.max(np.abs(y[off[u]:off[u + 1]] - ref)) < 1e-6


def test_post_filter_merlin_vs_restatement(mp):
    """pf_type='merlin' (src/magphase.py:3375-3465): the SPTK pipeline restated (binaries absent: parity unpinned, like mcep);
    the device computes the two energy integrals per frame, the oracle everything in NumPy."""
    g = np.load(os.path.join(GOLD, 'compressed_hvd704.npz'))
    mag = g['mag'].astype(np.float64)
    ref = orc.post_filter_merlin(mag, 48000)
    got = mp.post_filter_merlin(mag, 48000)
    assert got.shape == ref.shape and rms(got, ref) < 1e-5
    # the lifter sharpens the formants and keeps the frame energy: a real change, bounded, finite
    assert 0.1 < rms(got, mag) < 2.0 and np.isfinite(got).all()
    same = mp.post_filter_merlin(mag, 48000, pf_coef=1.0)                       # unit lifter: only the reference's own
    assert rms(same, mag) < 0.05                                                # cepstral round-trip quirk remains (:3397)

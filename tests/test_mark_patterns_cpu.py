"""Generated pitch-mark patterns (test pyramid item 1-iii): the oracle against the real reference on the
edge cases a corpus of natural speech rarely holds -- a mark at sample 0, shifts of one sample, half-integer marks
(half-to-even rounding, src/libutils.py:131-133), frames and pitch periods longer than fft_len (the truncation branch,
src/magphase.py:311-315), marks next to the end of the signal.  tests/test_gpu_lossless.py::test_random_mark_patterns holds
the kernels to the oracle on patterns of the same generator."""
import warnings

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import magphase_oracle as orc


def mark_pattern(rng, n_smpls, n_marks, style):
    """Strictly increasing marks in [0, n_smpls-1] with the given flavour; shared with the GPU test."""
    if style == 'speech':                                    # plausible pitch periods, 60-400 Hz at 48 kHz
        pm = np.cumsum(rng.integers(120, 800, n_marks)).astype(float)
    elif style == 'tiny':                                    # shifts of 1-3 samples mixed with ordinary ones
        pm = np.cumsum(np.where(rng.random(n_marks) < 0.5, rng.integers(1, 4, n_marks), rng.integers(100, 500, n_marks))).astype(float)
    elif style == 'long':                                    # some periods beyond fft_len (4096) and fft_len / 2
        pm = np.cumsum(np.where(rng.random(n_marks) < 0.3, rng.integers(2000, 6000, n_marks), rng.integers(150, 600, n_marks))).astype(float)
    else:                                                    # 'fractional': half-integers and arbitrary fractions
        pm = np.cumsum(rng.integers(100, 600, n_marks)).astype(float) + rng.choice([0.0, 0.5, 0.49, 0.51, 0.25], n_marks)
    if rng.random() < 0.5:
        pm = np.concatenate(([0.0], pm))                     # pm[0] = 0: an empty left side for the first frame
    pm = pm[pm <= n_smpls - 1]
    if rng.random() < 0.5 and pm.size and pm[-1] < n_smpls - 1:
        pm = np.concatenate((pm, [float(n_smpls - 1)]))      # last mark on the last sample: empty right side
    if pm.size:
        pm = pm[np.concatenate(([True], np.diff(np.round(pm)) > 0))]
    return pm


@settings(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(seed=st.integers(0, 2 ** 31 - 1), style=st.sampled_from(['speech', 'tiny', 'long', 'fractional']),
       fft_len=st.sampled_from([1024, 2048, 4096]))
def test_oracle_equals_reference_on_generated_marks(ref_modules, seed, style, fft_len):
    mp, la, lu = ref_modules
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3000, 40000))
    sig = rng.uniform(-1, 1, n)
    pm = mark_pattern(rng, n, int(rng.integers(3, 40)), style)
    if pm.size < 2:
        return
    voi = (rng.random(pm.size) < 0.6).astype(float)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m_fft_r, v_shift_r = mp.analysis_with_del_comp_from_pm(sig.copy(), 48000, pm.copy(), fft_len=fft_len)
        m_fft, v_shift = orc.analysis_fft_from_pm(sig, 48000, pm, fft_len=fft_len)
    assert np.array_equal(v_shift, v_shift_r)
    np.testing.assert_allclose(m_fft, m_fft_r, rtol=0, atol=1e-10)
    mag_r, real_r, imag_r, f0_r = mp.compute_lossless_feats(m_fft_r, v_shift_r, voi, 48000)
    mag, real, imag, f0 = orc.compute_lossless_feats(m_fft, v_shift, voi, 48000)
    assert np.array_equal(f0, f0_r, equal_nan=True)          # a mark at sample 0 is a shift of 0: f0 = voi * fs / 0 (inf or nan)
    np.testing.assert_allclose(mag, mag_r, rtol=0, atol=1e-10)
    if not np.all(np.isfinite(f0)):
        return
    # lossless resynthesis from those features: integer geometry of ola() (src/magphase.py:34-62) on arbitrary f0 tracks
    y_r = mp.synthesis_from_lossless(mag_r.copy(), real_r.copy(), imag_r.copy(), f0_r.copy(), 48000)
    y = orc.synthesis_from_lossless(mag, real, imag, f0, 48000)
    assert y.shape == y_r.shape
    np.testing.assert_allclose(y, y_r, rtol=0, atol=1e-10)


def test_mark_pattern_generator_is_strictly_increasing():
    for seed in range(50):
        rng = np.random.default_rng(seed)
        for style in ('speech', 'tiny', 'long', 'fractional'):
            pm = mark_pattern(rng, 20000, 25, style)
            r = np.round(pm)
            assert np.all(np.diff(r) > 0) and (pm.size == 0 or (r[0] >= 0 and r[-1] <= 19999))

"""The numpy oracle against the committed golden vectors (generated from the real reference by
tests/golden/make_golden.py).  No /root/reference needed: this is what pins the oracle on the GPU box."""
import os
import warnings

import numpy as np

import magphase_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLD, name))


def test_lossless_analysis_golden():
    g = load('lossless_synth48k.npz')
    sig = g['sig_i16'].astype(np.float64) / 32768.0
    mag, real, imag, f0, fs, v_shift = orc.analysis_lossless_from_pm(sig, int(g['fs']), g['pm'], g['voi'])
    assert np.array_equal(v_shift, g['v_shift'])
    assert np.array_equal(f0, g['v_f0'])
    rows, st = g['full_rows'], int(g['bin_step'])
    np.testing.assert_allclose(mag[rows], g['mag_rows'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(real[rows], g['real_rows'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(imag[rows], g['imag_rows'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(mag[:, ::st], g['mag_cols'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(real[:, ::st], g['real_cols'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(imag[:, ::st], g['imag_cols'], rtol=0, atol=1e-9)
    y = orc.synthesis_from_lossless(mag, real, imag, f0, fs)
    assert y.shape == g['syn'].shape
    np.testing.assert_allclose(y, g['syn'], rtol=0, atol=1e-13)
    np.testing.assert_allclose(orc.build_min_phase_from_mag_spec(mag[rows]), g['minph_rows'], rtol=1e-11, atol=1e-12)


def test_natural_speech_golden():
    """The oracle on the bundled natural recordings (hvd_593 / hvd_577 slices) against the real reference."""
    g = load('natural_48k.npz')
    for tag in ('a', 'b'):
        sig = g[tag + '_sig_i16'].astype(np.float64) / 32768.0
        mag, real, imag, f0, fs, v_shift = orc.analysis_lossless_from_pm(sig, int(g['fs']), g[tag + '_pm'], g[tag + '_voi'])
        assert np.array_equal(v_shift, g[tag + '_v_shift']) and np.array_equal(f0, g[tag + '_v_f0'])
        rows, st = g[tag + '_full_rows'], int(g['bin_step'])
        np.testing.assert_allclose(mag[rows], g[tag + '_mag_rows'], rtol=0, atol=1e-12)
        np.testing.assert_allclose(real[:, ::st], g[tag + '_real_cols'], rtol=0, atol=1e-9)
        np.testing.assert_allclose(imag[:, ::st], g[tag + '_imag_cols'], rtol=0, atol=1e-9)
        np.testing.assert_allclose(orc.synthesis_from_lossless(mag, real, imag, f0, fs), g[tag + '_syn'], rtol=0, atol=1e-13)
        mm, rr, ii, lf0 = orc.format_for_modelling(mag, real, imag, f0, fs, mag_dim=60, phase_dim=45)
        np.testing.assert_allclose(mm, g[tag + '_mag_mel_log'], rtol=0, atol=1e-6)      # float32 SPTK file rounding inside
        np.testing.assert_allclose(rr, g[tag + '_real_mel'], rtol=0, atol=1e-6)
        assert np.array_equal(lf0, g[tag + '_lf0'])
        np.random.seed(int(g[tag + '_seed']))
        y = orc.synthesis_from_compressed(g[tag + '_mag_mel_log'], g[tag + '_real_mel'], g[tag + '_imag_mel'], g[tag + '_lf0'],
                                          48000, b_out_hpf=False)
        np.testing.assert_allclose(y, g[tag + '_syn_compressed'], rtol=0, atol=1e-12)


def test_compressed_synthesis_golden():
    g = load('compressed_hvd704.npz')
    f64 = lambda k: g[k].astype(np.float64)
    mag, real, imag, lf0 = f64('mag'), f64('real'), f64('imag'), f64('lf0')
    cases = {'syn_var_nohpf': dict(b_out_hpf=False), 'syn_var_hpf': dict(b_out_hpf=True),
             'syn_const_nohpf': dict(b_out_hpf=False, b_const_rate=True),
             'syn_minph_nohpf': dict(b_out_hpf=False, per_phase_type='min_phase')}
    for k, kw in cases.items():
        np.random.seed(int(g['seed']))
        y = orc.synthesis_from_compressed(mag, real, imag, lf0, 48000, **kw)
        assert y.shape == g[k].shape, k
        # HPF: direct-form IIR rounding noise ~1e-8 (see tests/test_oracle_vs_ref.py)
        np.testing.assert_allclose(y, g[k], rtol=0, atol=1e-6 if kw.get('b_out_hpf') else 1e-12, err_msg=k)
    np.random.seed(int(g['seed']))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y = orc.synthesis_from_compressed(mag, real, imag, lf0, 16000, b_const_rate=True)
    assert y.shape == g['syn_16k_const_hpf'].shape
    np.testing.assert_allclose(y, g['syn_16k_const_hpf'], rtol=0, atol=1e-6)


def test_post_filter_and_unwarp_golden():
    g = load('compressed_hvd704.npz')
    mag = g['mag'].astype(np.float64)
    np.testing.assert_allclose(orc.post_filter(mag, 48000), g['post_filter_48k'], rtol=0, atol=1e-13)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        np.testing.assert_allclose(orc.post_filter(mag, 16000), g['post_filter_16k'], rtol=0, atol=1e-13)
    np.testing.assert_allclose(orc.sp_mel_unwarp(mag[:4], 2049, 0.77), g['mag_unwarp4'], rtol=0, atol=1e-12)
    r, i = orc.phase_uncompress(g['real'][:4].astype(np.float64), g['imag'][:4].astype(np.float64), 0.77, 4096, 48000)
    np.testing.assert_allclose(r, g['real_unwarp4'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(i, g['imag_unwarp4'], rtol=0, atol=1e-12)


def test_freqt_matrix_matches_scalar_recursion():
    """The vectorised freqt matrix against a literal scalar run of the published SPTK recursion."""
    rng = np.random.default_rng(1)
    c = rng.normal(size=40) * np.exp(-np.arange(40) / 8.0)
    a, m2 = 0.77, 11
    g = np.zeros(m2 + 1)
    d = np.zeros(m2 + 1)
    b = 1 - a * a
    for i in range(-(c.size - 1), 1):
        d[0] = g[0]
        g[0] = c[-i] + a * d[0]
        d[1] = g[1]
        g[1] = b * d[0] + a * d[1]
        for j in range(2, m2 + 1):
            d[j] = g[j]
            g[j] = d[j - 1] + a * (d[j] - g[j - 1])
    np.testing.assert_allclose(orc.freqt_matrix(m2 + 1, c.size, a) @ c, g, rtol=1e-12, atol=1e-13)


def test_mel_warp_unwarp_roundtrip_sane():
    """sp_mel_warp (SPTK restatement, parity unpinned) followed by the pinned sp_mel_unwarp must give back a
    smoothed version of the input log spectrum."""
    g = load('lossless_synth48k.npz')
    sig = g['sig_i16'].astype(np.float64) / 32768.0
    mag, real, imag, f0, fs, v_shift = orc.analysis_lossless_from_pm(sig, 48000, g['pm'], g['voi'])
    mel = orc.sp_mel_warp(mag, 60, alpha=0.77, in_type=3)
    back = orc.sp_mel_unwarp(np.log(mel), 2049, alpha=0.77, in_type='log')
    voiced = g['voi'] > 0
    lm = np.log(mag[voiced][:, :400])
    cc = np.corrcoef(lm.ravel(), back[voiced][:, :400].ravel())[0, 1]
    assert cc > 0.9
    mm, rr, ii, lf0 = orc.format_for_modelling(mag, real, imag, f0, fs)
    assert mm.shape == (mag.shape[0], 60) and rr.shape == (mag.shape[0], 45) and ii.shape == rr.shape
    assert np.all(np.abs(rr) <= 1) and np.all(rr[~voiced] == 0)
    assert np.all(lf0[~voiced] == orc.MAGIC)


def test_griffin_lim_golden():
    """Reference griffin_lim outputs (src/magphase.py:3318-3373) replayed through the oracle on the golden utterance."""
    g, gl = load('lossless_synth48k.npz'), load('griffin_lim_synth48k.npz')
    sig = g['sig_i16'].astype(np.float64) / 32768.0
    mag = orc.analysis_lossless_from_pm(sig, int(g['fs']), g['pm'], g['voi'])[0]
    rows = gl['full_rows']
    strong = mag[rows] > 1e-6 * mag.max()
    for init in ('linear', 'min_phase', 'random'):
        np.random.seed(int(gl['seed']))
        y, ph = orc.griffin_lim(mag.copy(), gl['v_shift'], phase_init=init, niters=int(gl['niters']))
        assert y.shape == gl['syn_' + init].shape
        np.testing.assert_allclose(y, gl['syn_' + init], rtol=0, atol=1e-11)
        d = np.abs(np.angle(np.exp(1j * (ph[rows] - gl['phase_rows_' + init]))))
        assert np.max(d[strong]) < 1e-7


def test_freqt_matrix_equals_the_all_pass_series_expansion():
    """SPTK's `freqt` is absent here (parity unpinned), but what it computes is defined in print: the frequency
    transformation of Oppenheim & Johnson (1972) that SPTK's manual cites -- C~(z~) = C(z) under the first-order all-pass
    substitution z~^-1 = (z^-1 - a) / (1 - a z^-1), i.e. z^-1 = (x + a) / (1 + a x) with x = z~^-1.  So row m, column n of the
    transform is the coefficient of x^m in ((x + a) / (1 + a x))^n.  Those coefficients are computed here by plain power-series
    arithmetic (no recursion of SPTK's shape involved) and must equal the oracle's freqt matrix, which follows the published
    recursion -- an independent derivation of the same linear map."""
    for alpha, n_out, n_in in ((0.77, 60, 400), (0.58, 45, 257), (0.0, 10, 30), (-0.3, 12, 64)):
        base = np.zeros(n_out, dtype=np.longdouble)                 # series of (x + a) / (1 + a x), truncated at x^(n_out-1)
        geo = (-np.longdouble(alpha)) ** np.arange(n_out)           # 1 / (1 + a x) = sum (-a)^k x^k
        base += alpha * geo
        base[1:] += geo[:-1]
        A = np.zeros((n_out, n_in), dtype=np.longdouble)
        p = np.zeros(n_out, dtype=np.longdouble)
        p[0] = 1.0                                                  # ((x + a) / (1 + a x))^0
        for n in range(n_in):
            A[:, n] = p
            p = np.convolve(p, base)[:n_out]                        # next power, truncated
        np.testing.assert_allclose(orc.freqt_matrix(n_out, n_in, alpha), A.astype(np.float64), rtol=0, atol=1e-11)


def test_restated_merlin_post_filter_preserves_frame_energy():
    """pf_type='merlin' (src/magphase.py:3375-3465) pipes the cepstrum through SPTK binaries that are absent here (parity
    unpinned).  The pipeline's documented purpose is checkable without them: it sharpens the formants with a cepstral lifter
    and then resets c0 so that every frame keeps its energy (the r[0] of `c2acr`).  The restatement must therefore change the
    spectrum (it is not a no-op) while the autocorrelation r[0] of every frame, recomputed here from the OUTPUT, stays
    where it was -- to the float32 round-off of the pipe and the 60-coefficient truncation."""
    g = np.load(os.path.join(GOLD, 'compressed_hvd704.npz'))
    mag = g['mag'].astype(np.float64)
    out = orc.post_filter_merlin(mag, 48000)

    def r0_of(m):
        n = m.shape[1]
        ext = np.hstack((m, m[:, -2:0:-1]))
        ceps = np.fft.ifft(ext, axis=1).real
        ceps[:, 1:(n - 2)] *= 2
        F = orc.freqt_matrix(2048, n, -orc.define_alpha(48000))
        return orc.sptk_c2acr_r0(ceps[:, :n] @ F.T, 4096)

    ratio = r0_of(out) / r0_of(mag)
    assert np.all(np.abs(ratio - 1.0) < 5e-3), (ratio.min(), ratio.max())
    assert np.abs(out - mag).mean() > 0.1                            # the formant enhancement itself
    # mc2b / b2mc are exact inverses of each other (their definition), whatever alpha
    x = np.random.default_rng(0).standard_normal((5, 60))
    np.testing.assert_allclose(orc.sptk_b2mc(orc.sptk_mc2b(x, 0.77), 0.77), x, rtol=0, atol=1e-12)

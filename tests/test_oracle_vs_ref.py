"""Pin the numpy oracle (oracle/magphase_oracle.py) against the real reference (oracle/_ref, the
mechanically py3-translated copy of /root/reference/src).  Runs only where /root/reference exists."""
import os
import warnings

import numpy as np
import pytest

import magphase_oracle as orc
from magphase_b200.synth import synth_marks_for_wav, synth_utterance

from conftest import REF_DATA


def _wav(name):
    from scipy.io import wavfile
    fs, d = wavfile.read(os.path.join(REF_DATA, 'wavs_nat', name + '.wav'))
    return d.astype(np.float64) / 32768.0, fs


def _pred(tok):
    d = os.path.join(REF_DATA, 'params_predicted')
    rd = lambda ext, dim: np.fromfile(os.path.join(d, tok + ext), dtype=np.float32).reshape(-1, dim).astype(np.float64)
    return rd('.mag', 60), rd('.real', 45), rd('.imag', 45), rd('.lf0', 1)[:, 0]


@pytest.mark.parametrize('wav', ['hvd_593', 'hvd_577'])
def test_lossless_analysis_matches_reference(ref_modules, wav):
    mp, la, lu = ref_modules
    sig, fs = _wav(wav)
    pm, voi = synth_marks_for_wav(sig.size, fs, seed=3)
    m_fft_r, v_shift_r = mp.analysis_with_del_comp_from_pm(sig, fs, pm)
    mag_r, real_r, imag_r, f0_r = mp.compute_lossless_feats(m_fft_r, v_shift_r, voi, fs)
    mag, real, imag, f0, _, v_shift = orc.analysis_lossless_from_pm(sig, fs, pm, voi)
    assert np.array_equal(v_shift, v_shift_r)
    assert np.array_equal(f0, f0_r)
    np.testing.assert_allclose(mag, mag_r, rtol=0, atol=1e-12)
    np.testing.assert_allclose(real, real_r, rtol=0, atol=1e-9)
    np.testing.assert_allclose(imag, imag_r, rtol=0, atol=1e-9)


def test_lossless_analysis_fractional_and_edge_marks(ref_modules):
    mp, la, lu = ref_modules
    rng = np.random.default_rng(5)
    sig = rng.uniform(-0.5, 0.5, 30000)
    # marks include pm[0]=0, a shift of 1, half-integers (half-to-even) and a frame longer than fft_len
    pm = np.array([0.0, 1.0, 240.5, 241.5, 700.49, 1200.0, 6000.0, 6300.5, 9000.0, 29990.0])
    voi = (np.arange(pm.size) % 2).astype(float)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m_fft_r, v_shift_r = mp.analysis_with_del_comp_from_pm(sig, 48000, pm)
        m_fft, v_shift = orc.analysis_fft_from_pm(sig, 48000, pm)
    assert np.array_equal(v_shift, v_shift_r)
    np.testing.assert_allclose(m_fft, m_fft_r, rtol=0, atol=1e-11)


def test_lossless_synthesis_matches_reference(ref_modules):
    mp, la, lu = ref_modules
    sig, pm, voi = synth_utterance(2, dur_s=1.0)
    mag, real, imag, f0, fs, _ = orc.analysis_lossless_from_pm(sig, 48000, pm, voi)
    y_r = mp.synthesis_from_lossless(mag.copy(), real.copy(), imag.copy(), f0.copy(), fs)
    y = orc.synthesis_from_lossless(mag, real, imag, f0, fs)
    assert y.shape == y_r.shape
    np.testing.assert_allclose(y, y_r, rtol=0, atol=1e-13)


def test_tables_match_reference(ref_modules):
    mp, la, lu = ref_modules
    for fs in (48000, 16000):
        assert orc.define_alpha(fs) == mp.define_alpha(fs)
        assert orc.define_fft_len(fs) == mp.define_fft_len(fs)
        assert orc.define_crossfade_params(fs) == mp.define_crossfade_params(fs)
    for fs, pd, al in ((48000, 45, 0.77), (48000, 10, 0.77), (16000, 45, 0.58), (48000, 10, 0.0)):
        cf = mp.define_crossfade_params(fs)[0]
        assert orc.n_full_mel_coeffs(cf, pd, al, fs) == mp.get_num_full_mel_coeffs_from_num_phase_coeffs(cf, pd, al, fs)
    H = 2049
    ones, zeros = np.ones((1, H)), np.zeros((1, H))
    ref_curve = la.spectral_crossfade(ones, zeros, 5000, 2000, 48000, freq_scale='hz', win_func=np.hanning)[0]
    np.testing.assert_array_equal(orc.crossfade_curve(H, 5000, 2000, 48000), ref_curve)
    np.testing.assert_array_equal(orc.build_mel_curve(0.77, H, amp=3.5), la.build_mel_curve(0.77, H, amp=3.5))
    for (l, r) in ((0, 5), (7, 0), (240, 371), (1, 1)):
        np.testing.assert_array_equal(orc.asym_window(l, r), la.gen_non_symmetric_win(l, r, np.hanning))
        np.testing.assert_array_equal(orc.asym_window(l, r, 'bartlett2.5'),
                                      la.gen_non_symmetric_win(l, r, mp.voi_noise_window))
    np.testing.assert_array_equal(orc.centred_window(480, 611, 4096),
                                  la.gen_centr_win(480, 611, 4096, win_func=mp.raised_hanning, b_fill_w_bound_val=True))


def test_mel_unwarp_and_minphase_match_reference(ref_modules):
    mp, la, lu = ref_modules
    mag_mel, real_mel, imag_mel, lf0 = _pred('hvd_704')
    a = la.sp_mel_unwarp(mag_mel[:40].copy(), 2049, alpha=0.77, in_type='log')
    b = orc.sp_mel_unwarp(mag_mel[:40], 2049, alpha=0.77, in_type='log')
    np.testing.assert_allclose(b, a, rtol=0, atol=1e-12)
    m_mag = np.exp(b)
    np.testing.assert_allclose(orc.build_min_phase_from_mag_spec(m_mag), la.build_min_phase_from_mag_spec(m_mag.copy()),
                               rtol=1e-11, atol=1e-12)
    r1, i1 = mp.phase_uncompress_type1_mcep(real_mel[:30].copy(), imag_mel[:30].copy(), 0.77, 4096, 48000)
    r2, i2 = orc.phase_uncompress(real_mel[:30], imag_mel[:30], 0.77, 4096, 48000)
    np.testing.assert_allclose(r2, r1, rtol=0, atol=1e-12)
    np.testing.assert_allclose(i2, i1, rtol=0, atol=1e-12)


def test_post_filter_matches_reference(ref_modules):
    mp, la, lu = ref_modules
    mag_mel = _pred('hvd_705')[0]
    np.testing.assert_allclose(orc.post_filter(mag_mel, 48000), mp.post_filter(mag_mel.copy(), 48000), rtol=0, atol=1e-13)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        np.testing.assert_allclose(orc.post_filter(mag_mel, 16000), mp.post_filter(mag_mel.copy(), 16000), rtol=0, atol=1e-13)


@pytest.mark.parametrize('tok,hpf,ptype', [('hvd_704', True, 'magphase'), ('hvd_708', False, 'magphase'),
                                           ('hvd_706', False, 'min_phase')])
def test_compressed_synthesis_matches_reference(ref_modules, tok, hpf, ptype):
    mp, la, lu = ref_modules
    mag_mel, real_mel, imag_mel, lf0 = _pred(tok)
    np.random.seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y_r = mp.synthesis_from_compressed(mag_mel.copy(), real_mel.copy(), imag_mel.copy(), lf0.copy(), 48000,
                                           b_out_hpf=hpf, per_phase_type=ptype)
    np.random.seed(11)
    y = orc.synthesis_from_compressed(mag_mel, real_mel, imag_mel, lf0, 48000, b_out_hpf=hpf, per_phase_type=ptype)
    assert y.shape == y_r.shape
    # Without the HPF the two agree to ~1e-15.  The reference's 4th-order 40 Hz Butterworth runs as a
    # direct-form lfilter (src/magphase.py:990-995) whose poles sit at |z|~0.995: its own rounding noise is
    # ~1e-8, so two float64 evaluations whose inputs differ in the last bit already differ by ~3e-8.
    np.testing.assert_allclose(y, y_r, rtol=0, atol=1e-6 if hpf else 1e-12)
    # ('linear' is not cross-checked: the reference raises TypeError under numpy>=2 at src/magphase.py:951.)


def test_compressed_synthesis_const_rate_16k_matches_reference(ref_modules):
    mp, la, lu = ref_modules
    mag_mel, real_mel, imag_mel, lf0 = _pred('hvd_705')
    # plumbing input for the 16 kHz constant-rate configuration (BASELINE config 4): same features, lf0 shifted
    lf0_16 = lf0
    np.random.seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y_r = mp.synthesis_from_compressed(mag_mel.copy(), real_mel.copy(), imag_mel.copy(), lf0_16.copy(), 16000,
                                           b_const_rate=True)
    np.random.seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y = orc.synthesis_from_compressed(mag_mel, real_mel, imag_mel, lf0_16, 16000, b_const_rate=True)
    assert y.shape == y_r.shape
    np.testing.assert_allclose(y, y_r, rtol=0, atol=1e-6)   # HPF on (default): see note above


def test_compressed_synthesis_min_phase_const_rate_matches_reference(ref_modules):
    """per_phase_type='min_phase' with b_const_rate=True: the reference interpolates the un-warped magnitudes to the
    synthesis frames first and builds the minimum phase of the interpolated rows (src/magphase.py:861-870, :935-936)."""
    mp, la, lu = ref_modules
    mag_mel, real_mel, imag_mel, lf0 = _pred('hvd_706')
    np.random.seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y_r = mp.synthesis_from_compressed(mag_mel.copy(), real_mel.copy(), imag_mel.copy(), lf0.copy(), 48000,
                                           b_const_rate=True, per_phase_type='min_phase', b_out_hpf=False)
    np.random.seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y = orc.synthesis_from_compressed(mag_mel, real_mel, imag_mel, lf0, 48000, b_const_rate=True,
                                          per_phase_type='min_phase', b_out_hpf=False)
    assert y.shape == y_r.shape
    np.testing.assert_allclose(y, y_r, rtol=0, atol=1e-12)


def test_var_to_const_rate_matches_reference(ref_modules):
    mp, la, lu = ref_modules
    rng = np.random.default_rng(0)
    v_shift = rng.integers(200, 600, 50)
    pm = np.cumsum(v_shift)
    m = rng.normal(size=(50, 7))
    np.testing.assert_allclose(orc.interp_from_variable_to_const_frm_rate(m, pm, 5.0, 48000),
                               mp.interp_from_variable_to_const_frm_rate(m.copy(), pm, 5.0, 48000), rtol=0, atol=1e-14)


def test_legacy_ph_encoding_matches_reference_around_sptk(ref_modules, monkeypatch):
    """analysis_with_del_comp_and_ph_encoding (src/magphase.py:573-598): the reference needs REAPER and SPTK; with
    la.get_pitch_marks fed the same marks and la.sp_to_mcep replaced by the oracle's restatement of `mcep -j 0`,
    everything else (windowing, FFT, phase encoding, cubic resampling, dimensions) is the reference's own code."""
    mp, la, lu = ref_modules
    sig, pm, voi = synth_utterance(14, dur_s=0.5)
    pm_sec = pm / 48000.0
    monkeypatch.setattr(la, 'get_pitch_marks', lambda v_sig, fs: pm_sec)
    monkeypatch.setattr(la, 'sp_to_mcep', lambda m_sp, n_coeffs=60, alpha=0.77, in_type=3, fft_len=0:
                        orc.mcep_j0(m_sp, n_coeffs=n_coeffs, alpha=alpha, in_type=in_type, fft_len=fft_len))
    ref = mp.analysis_with_del_comp_and_ph_encoding(sig, 4096, 48000, 4500)
    got = orc.analysis_with_del_comp_and_ph_encoding_from_pm(sig, 4096, 48000, 4500, pm_sec)
    assert np.array_equal(got[3], ref[3])
    for a, b in zip(got[:3], ref[:3]):
        assert a.shape == b.shape == (pm.size, 60)
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-5)     # float32 outputs of an ill-conditioned dB -> cepstrum map
    with pytest.raises(ValueError):
        mp.analysis_with_del_comp_and_ph_encoding(sig, 512, 48000, 4500)


def test_griffin_lim_matches_reference(ref_modules):
    """src/magphase.py:3318-3373, every phase_init; np.random.rand for 'random' is drawn from the same seeded stream."""
    mp, la, lu = ref_modules
    sig, pm, voi = synth_utterance(8, dur_s=0.6)
    mag, real, imag, f0, fs, shift = orc.analysis_lossless_from_pm(sig, 48000, pm, voi)
    for init in ('linear', 'min_phase', 'random', np.angle(real + 1j * imag)):
        arg = lambda: init if isinstance(init, str) else init.copy()       # add_hermitian_half mutates its input
        np.random.seed(3)
        y_ref, ph_ref = mp.griffin_lim(mag.copy(), shift, phase_init=arg(), niters=5)
        np.random.seed(3)
        y, ph = orc.griffin_lim(mag.copy(), shift, phase_init=arg(), niters=5)
        assert y.shape == y_ref.shape and ph.shape == ph_ref.shape
        np.testing.assert_allclose(y, y_ref, rtol=0, atol=1e-12)
        strong = mag > 1e-6 * mag.max()
        assert np.max(np.abs(np.angle(np.exp(1j * (ph - ph_ref))))[strong]) < 1e-8


def test_const_rate_reverse_scan_is_bit_exact_against_reference(ref_modules):
    """get_shifts_and_frm_locs_from_const_shifts (src/magphase.py:1426-1449): the reference walks back with scipy's
    interp1d until it raises; the oracle and the host mirror use np.interp with explicit range checks.  The shifts are
    truncated to integers downstream (:879), so the three must agree to the last bit -- randomised tracks, both rates."""
    import magphase_b200.magphase as host
    mp, la, lu = ref_modules
    rng = np.random.default_rng(11)
    for trial in range(24):
        fs = (48000, 16000)[trial % 2]
        n = int(rng.integers(3, 400))
        f0 = rng.uniform(60, 380, n)
        f0[rng.uniform(size=n) < 0.35] = 0.0                      # unvoiced stretches
        v_shift_c = orc.f0_to_shift(f0, fs)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            s_ref, l_ref = mp.get_shifts_and_frm_locs_from_const_shifts(v_shift_c.copy(), 5.0, fs, interp_type='linear')
        s_orc, l_orc = orc.get_shifts_and_frm_locs_from_const_shifts(v_shift_c, 5.0, fs)
        s_host, l_host = host.get_shifts_and_frm_locs_from_const_shifts(v_shift_c, 5.0, fs)
        for s, l in ((s_orc, l_orc), (s_host, l_host)):
            assert s.shape == np.shape(s_ref) and np.array_equal(s, s_ref), (trial, fs, n)
            assert np.array_equal(l, l_ref), (trial, fs, n)


def test_const_to_variable_rate_rows_match_reference(ref_modules):
    """interp_from_const_to_variable_rate (src/magphase.py:2242-2252) against the row pairs + weights the kernels use."""
    import magphase_b200.magphase as host
    mp, la, lu = ref_modules
    rng = np.random.default_rng(12)
    for fs in (48000, 16000):
        n_c = 120
        step = fs * 5.0 / 1000
        data = rng.normal(size=(n_c, 6))
        locs = np.sort(rng.uniform(step, step * n_c, 300))
        locs[0], locs[-1] = step, step * n_c                      # both ends of the interpolation range
        ref = mp.interp_from_const_to_variable_rate(data.copy(), locs, 5.0, fs, interp_type='linear')
        r0, r1, w = host._const_rate_rows(locs, n_c, step)
        got = data[r0] + (data[r1] - data[r0]) * w[:, None]
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13)
        np.testing.assert_allclose(orc.interp_from_const_to_variable_rate(data, locs, 5.0, fs), ref, rtol=0, atol=1e-14)


def test_intermediate_epochs_match_reference(ref_modules):
    """nwin_per_pitch_period >= 1 (src/magphase.py:280-288): the host mirror's epoch expansion followed by its frame
    geometry gives the reference's v_shift, integer for integer."""
    import magphase_b200.magphase as host
    mp, la, lu = ref_modules
    sig, pm, voi = synth_utterance(5, fs=48000, dur_s=0.6)
    for nwin in (1.0, 2.0):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            _, v_shift_ref = mp.analysis_with_del_comp_from_pm(sig.copy(), 48000, pm.copy(), nwin_per_pitch_period=nwin)
        v_pm = host._expand_epochs(np.asarray(pm, dtype=np.float64), nwin)
        _, v_shift, _ = host.frame_geometry(v_pm, sig.size)
        assert np.array_equal(v_shift, v_shift_ref), nwin


def _tukeyish(n):
    """A window that is neither symmetric-peaked at 1 nor one of the two built-ins (centre value 0.9)."""
    return 0.9 * np.hamming(n) ** 1.5


@pytest.mark.parametrize('win', ['hamming', 'blackman', 'custom', 'list'])
def test_arbitrary_win_func_matches_reference(ref_modules, win):
    """win_func as any callable or a per-frame list (src/magphase.py:102-108, src/libaudio.py:70-84): the oracle's
    restatement and the mirror's host-side weights (window_weights / prewindowed_frames, which feed the kernels their
    samples under MPB_WIN_RECT) against the reference's own windowing() and analysis."""
    mp, la, lu = ref_modules
    import magphase_b200.magphase as mpb
    sig, pm, voi = synth_utterance(4, dur_s=0.5)
    n = pm.size
    fn = {'hamming': np.hamming, 'blackman': np.blackman, 'custom': _tukeyish,
          'list': [(np.hanning, np.hamming, _tukeyish, mp.voi_noise_window)[f % 4] for f in range(n)]}[win]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m_fft_r, v_shift_r = mp.analysis_with_del_comp_from_pm(sig, 48000, pm, win_func=fn)
        m_fft, v_shift = orc.analysis_fft_from_pm(sig, 48000, pm, win_func=fn)
    assert np.array_equal(v_shift, v_shift_r)
    np.testing.assert_allclose(m_fft, m_fft_r, rtol=0, atol=1e-11)
    # the host half of the CUDA path: frames times weights, laid back to back
    l_frames, v_lens, P, v_shift_w, v_rights = mp.windowing(sig, pm, win_func=fn)
    fns = mpb._win_list(fn, n)
    pre, centre, idx, w_all = mpb.prewindowed_frames(sig, P[1:-1], v_shift_w, v_rights, fns)
    assert pre.size == int(np.sum(v_lens))
    np.testing.assert_array_equal(pre, np.concatenate(l_frames))
    assert np.array_equal(centre, np.concatenate(([0], np.cumsum(v_lens)[:-1])) + v_shift_w)
    assert np.array_equal(sig[idx] * w_all, pre)


def test_griffin_lim_with_arbitrary_window_matches_reference(ref_modules):
    mp, la, lu = ref_modules
    sig, pm, voi = synth_utterance(8, dur_s=0.5)
    mag, real, imag, f0, fs, shift = orc.analysis_lossless_from_pm(sig, 48000, pm, voi)
    y_ref, ph_ref = mp.griffin_lim(mag.copy(), shift, win_func=np.hamming, phase_init='linear', niters=4)
    y, ph = orc.griffin_lim(mag.copy(), shift, phase_init='linear', niters=4, win_func=np.hamming)
    np.testing.assert_allclose(y, y_ref, rtol=0, atol=1e-12)
    strong = mag > 1e-6 * mag.max()
    assert np.max(np.abs(np.angle(np.exp(1j * (ph - ph_ref))))[strong]) < 1e-8


def test_windowing_helper_matches_reference(ref_modules):
    """mirror.windowing() (host helper with the reference's signature, src/magphase.py:74-119) against the reference."""
    mp, la, lu = ref_modules
    import magphase_b200.magphase as mpb
    sig, pm, voi = synth_utterance(6, dur_s=0.4)
    for fn_ref, fn in ((np.hanning, np.hanning), (np.hamming, np.hamming),
                       ([mp.voi_noise_window if i % 2 else np.hanning for i in range(pm.size)],
                        [mpb.voi_noise_window if i % 2 else np.hanning for i in range(pm.size)])):
        l_r, lens_r, P_r, shift_r, rights_r = mp.windowing(sig, pm, win_func=fn_ref)
        l_m, lens_m, P_m, shift_m, rights_m = mpb.windowing(sig, pm, win_func=fn)
        assert np.array_equal(lens_r, lens_m) and np.array_equal(P_r, P_m)
        assert np.array_equal(shift_r, shift_m) and np.array_equal(rights_r, rights_m)
        assert len(l_r) == len(l_m)
        for a, b in zip(l_r, l_m):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize('alpha,fft_len', [(0.77, 4096), (0.58, 2048)])
def test_restated_mcep_inverts_the_references_own_cosine_map(ref_modules, alpha, fft_len):
    """SPTK's `mcep -j 0` is a third-party binary that is not available (parity unpinned).  What CAN be anchored
    on the reference itself: its own la.mcep_to_sp_cosmat (src/libaudio.py:605-631) is the synthesis-side inverse of the
    analysis-side `mcep`: a spectrum built from mel-cepstral coefficients by the REFERENCE's function must give those
    coefficients back through the restated analysis step (to the float32 rounding SPTK's file interface applies and the
    1e-8 floor of `-e`).  A wrong factor (the halved c[0] / c[N/2], power vs. amplitude), a wrong sign of alpha or a wrong
    frequency transformation all break this identity at the 1e-1 level."""
    mp, la, lu = ref_modules
    rng = np.random.default_rng(0)
    H = fft_len // 2 + 1
    mc = np.zeros((6, 60))
    mc[:, :25] = rng.standard_normal((6, 25)) * np.exp(-0.25 * np.arange(25))[None, :]
    mc[:, 0] = rng.uniform(-3, 1, 6)
    sp = la.mcep_to_sp_cosmat(mc, H, alpha=alpha, out_type='abs')
    assert sp.min() > 1e-3                                                   # far above the 1e-8 floor of `-e`
    for in_type, x in ((3, sp), (2, np.log(sp)), (1, 20 * np.log10(sp))):
        back = orc.mcep_j0(x, n_coeffs=60, alpha=alpha, in_type=in_type)
        assert np.sqrt(np.mean((back - mc) ** 2)) < 3e-6, in_type
        assert np.abs(back - mc).max() < 3e-5, in_type
    # ... and the whole warp of format_for_modelling: la.sp_mel_warp's output is the log spectrum on the alpha=0 axis of the
    # same coefficients (src/libaudio.py:643-661)
    warped = orc.sp_mel_warp(sp, 60, alpha=alpha, in_type=3)
    direct = la.mcep_to_sp_cosmat(mc, 60, alpha=0.0, out_type='abs')
    np.testing.assert_allclose(np.log(warped), np.log(direct), rtol=0, atol=2e-4)

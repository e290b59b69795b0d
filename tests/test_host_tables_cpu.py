"""The constant tables the host mirror hands to the kernels (magphase_b200/magphase.py), against the oracle's own
restatement of the reference lines (which tests/test_oracle_vs_ref.py pins to the reference itself).  CPU only."""
import numpy as np
import pytest
from scipy import signal

import magphase_oracle as orc
import magphase_b200.magphase as mp


@pytest.mark.parametrize('fs,fft_len', [(48000, 4096), (16000, 2048), (16000, 1024)])
def test_unwarp_matrix_is_sp_mel_unwarp(fs, fft_len):
    """la.sp_mel_unwarp (src/libaudio.py:667-684) is linear in its input: mel_unwarp_matrix is that map, including the
    un-doubled cepstral index n_c - 2 (:679)."""
    H = fft_len // 2 + 1
    alpha = mp.define_alpha(fs)
    rng = np.random.default_rng(3)
    for n_c in (60, 45, 13):
        x = rng.normal(size=(7, n_c))
        ref = orc.sp_mel_unwarp(x, H, alpha=alpha, in_type='log')
        got = x @ mp.mel_unwarp_matrix(n_c, H, alpha)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize('fs,fft_len', [(48000, 4096), (16000, 2048)])
def test_phase_unwarp_matrix_folds_the_nearest_padding(fs, fft_len):
    """phase_uncompress_type1_mcep (src/magphase.py:1219-1235): pad phase_dim -> nmel by repeating the last column, then
    un-warp.  The plan folds the padding into a phase_dim x HB matrix (same construction as _SynPlan)."""
    H = fft_len // 2 + 1
    alpha = mp.define_alpha(fs)
    crsf_cf, crsf_bw = orc.define_crossfade_params(fs)
    phase_dim = 45
    nmel = mp.get_num_full_mel_coeffs_from_num_phase_coeffs(crsf_cf, phase_dim, alpha, fs)
    assert nmel == orc.n_full_mel_coeffs(crsf_cf, phase_dim, alpha, fs)
    u_full = mp.mel_unwarp_matrix(nmel, H, alpha)
    u_ph = np.zeros((phase_dim, H))
    np.add.at(u_ph, np.minimum(np.arange(nmel), phase_dim - 1), u_full)
    rng = np.random.default_rng(4)
    xr, xi = rng.uniform(-1, 1, (5, phase_dim)), rng.uniform(-1, 1, (5, phase_dim))
    ref_r, ref_i = orc.phase_uncompress(xr, xi, alpha, fft_len, fs)
    np.testing.assert_allclose(xr @ u_ph, ref_r, rtol=0, atol=1e-12)
    np.testing.assert_allclose(xi @ u_ph, ref_i, rtol=0, atol=1e-12)


def test_crossfade_curve_and_its_upper_edge():
    for fs, fft_len in ((48000, 4096), (16000, 2048)):
        H = fft_len // 2 + 1
        cf, bw = orc.define_crossfade_params(fs)
        curve, bin_r = mp.crossfade_curve(H, cf, bw, fs)
        assert np.array_equal(curve, orc.crossfade_curve(H, cf, bw, fs))
        assert curve[bin_r] == 0.0 and np.all(curve[bin_r:] == 0.0) and curve[bin_r - 1] > 0.0   # periodic part ends below bin_r
    assert mp.crossfade_curve(2049, 5000, 2000, 48000)[1] == 512
    assert np.array_equal(mp.build_mel_curve(0.77, 2049, amp=3.5), orc.build_mel_curve(0.77, 2049, amp=3.5))


@pytest.mark.parametrize('fs', [48000, 16000, 22050, 44100])
def test_output_hpf_sections_reproduce_the_reference_design(fs):
    """output_hpf_sos factors scipy.signal.butter(4, 40 Hz, 'highpass') (src/magphase.py:981-995) into two biquads for the
    device scan: their product is the reference's (b, a), and filtering with them equals lfilter."""
    v_b, v_a = signal.butter(4, 40 / (fs / 2.0), btype='highpass')
    sos = mp.output_hpf_sos(fs)
    b = np.polymul(sos[0, :3], sos[1, :3])
    a = np.polymul(sos[0, 3:], sos[1, 3:])
    np.testing.assert_allclose(b, v_b, rtol=1e-12, atol=0)
    np.testing.assert_allclose(a, v_a, rtol=1e-12, atol=0)
    x = np.random.default_rng(5).uniform(-1, 1, 4000)
    np.testing.assert_allclose(signal.sosfilt(sos, x), signal.lfilter(v_b, v_a, x), rtol=0, atol=1e-6)


def test_post_filter_argument_errors_come_before_device_work():
    with pytest.raises(ValueError):
        mp.post_filter(np.zeros((4, 60)), 22050)           # no defaults for this rate (src/magphase.py:2330-2334)


def test_sp_mel_warp_host_half_against_the_oracle(monkeypatch):
    """mp.sp_mel_warp = sp_to_mcep (device, tests/test_gpu_compressed_analysis.py) + the cosine matrix of
    la.mcep_to_sp_cosmat(alpha=0) on the host: with the oracle's mcep_j0 standing in for the device half, the composition
    must equal the oracle's sp_mel_warp for all three input types (src/libaudio.py:643-661)."""
    import numpy as np
    import magphase_oracle as orc
    import magphase_b200.magphase as mp
    monkeypatch.setattr(mp, 'sp_to_mcep', lambda m_sp, n_coeffs=60, alpha=0.77, in_type=3, fft_len=0:
                        orc.mcep_j0(m_sp, n_coeffs=n_coeffs, alpha=alpha, in_type=in_type))
    rng = np.random.default_rng(0)
    mag = np.abs(rng.standard_normal((5, 1025))) + 0.01
    for in_type, x in ((3, mag), (2, np.log(mag)), (1, 20 * np.log10(mag))):
        got = mp.sp_mel_warp(x, 40, alpha=0.58, in_type=in_type)
        ref = orc.sp_mel_warp(x, 40, alpha=0.58, in_type=in_type)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)

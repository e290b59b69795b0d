"""GPU parity of the low-dimensional compression (format_for_modelling / analysis_compressed) against the oracle.

The oracle's SPTK `mcep -j 0` restatement is itself unpinned (no SPTK binary/source/fixture available, see
oracle/magphase_oracle.py), so these tests prove CUDA == restatement, within 1e-5 RMS."""
import ctypes
import warnings

import numpy as np
import pytest

import magphase_oracle as orc
from magphase_b200.synth import synth_utterance

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rms(a, b):
    return float(np.sqrt(np.mean(np.abs(np.asarray(a) - np.asarray(b)) ** 2)))


@pytest.fixture(scope='module')
def mp():
    import magphase_b200.magphase as m
    return m


def test_warp_matrix_matches_freqt_recursion(mp):
    """The device-built W^T = (freqt . cosine-IFFT)^T against the oracle's matrices."""
    from magphase_b200 import _lib
    plan = mp._MelPlan.get(48000, 4096, 60, 45, None)
    H, N = 2049, 4096
    k = np.arange(H)
    w = np.full(H, 2.0); w[0] = w[-1] = 1.0
    Cm = np.cos(2 * np.pi * np.outer(k, k) / N) * w[None, :] / N
    Cm[0] *= 0.5; Cm[-1] *= 0.5
    for which, n in ((0, 60), (1, plan.nmel)):
        got = np.zeros((H, n), dtype=np.float32)
        _lib.check(_lib.lib().mpb_mel_get_warp_matrix(plan.handle, which, _lib.ptr(got)))
        ref = (orc.freqt_matrix(n, H, 0.77) @ Cm).T
        assert np.max(np.abs(got - ref)) < 1e-9 + 1e-6 * np.max(np.abs(ref))


@pytest.mark.parametrize('fs,phase_dim,alpha_phase', [(48000, 45, None), (48000, 10, 0.0), (16000, 45, None)])
def test_format_for_modelling_vs_oracle(mp, fs, phase_dim, alpha_phase):
    sig, pm, voi = synth_utterance(6, fs=fs, dur_s=0.8)
    mag, real, imag, f0, _, _ = orc.analysis_lossless_from_pm(sig, fs, pm, voi)
    ref = orc.format_for_modelling(mag, real, imag, f0, fs, mag_dim=60, phase_dim=phase_dim, alpha_phase=alpha_phase)
    got = mp.format_for_modelling(mag, real, imag, f0, fs, mag_dim=60, phase_dim=phase_dim, alpha_phase=alpha_phase)
    for name, a, b in zip(('mag_mel_log', 'real_mel', 'imag_mel'), got[:3], ref[:3]):
        assert a.shape == b.shape and a.dtype == np.float64
        assert rms(a, b) < TOL, (name, rms(a, b))
        assert np.max(np.abs(a - b)) < 1e-4, (name, np.max(np.abs(a - b)))
    assert np.array_equal(got[3], ref[3]), 'lf0 (host float64 bookkeeping) must be bit-exact'
    unv = f0 == 0
    assert np.all(got[1][unv] == 0) and np.all(np.abs(got[1]) <= 1) and np.all(np.abs(got[2]) <= 1)


def test_analysis_compressed_fused_vs_oracle(mp):
    sig, pm, voi = synth_utterance(8, fs=48000, dur_s=1.0)
    ref = orc.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    got = mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    assert np.array_equal(got[4], ref[4]) and got[5] == 48000 and got[6] == 4096
    assert np.array_equal(got[3], ref[3])
    for name, a, b in zip(('mag_mel_log', 'real_mel', 'imag_mel'), got[:3], ref[:3]):
        assert a.shape == b.shape
        assert rms(a, b) < TOL, (name, rms(a, b))


def test_analysis_compressed_all_unvoiced_all_voiced_and_mixed_batch(mp):
    """The phase streams run over the voiced frames only (compacted rows, tensor-core tiles sized by the device-side count):
    an utterance without a single voiced frame, one without an unvoiced frame, and a batch that mixes them with an ordinary
    utterance must all match the oracle; unvoiced frames keep exact zeros in real_mel / imag_mel (src/magphase.py:2527-2528)."""
    sig, pm, voi = synth_utterance(31, fs=48000, dur_s=0.6)
    cases = [(sig, pm, np.zeros_like(voi)), (sig, pm, np.ones_like(voi)), (sig, pm, voi)]
    refs = [orc.analysis_compressed_from_pm(s, 48000, p, v, mag_dim=60, phase_dim=45) for s, p, v in cases]
    singles = [mp.analysis_compressed_from_pm(s, 48000, p, v, mag_dim=60, phase_dim=45) for s, p, v in cases]
    batch = mp.analysis_compressed_batch([c[0] for c in cases], 48000, [c[1] for c in cases], [c[2] for c in cases], mag_dim=60,
                                         phase_dim=45)
    for ref, one, bat, (_, _, v) in zip(refs, singles, batch, cases):
        for a, b, c in zip(one[:3], bat[:3], ref[:3]):
            assert a.shape == c.shape and rms(a, c) < TOL and np.array_equal(a, b)       # batch == single, byte for byte
        assert np.array_equal(one[3], ref[3]) and np.array_equal(bat[3], ref[3])
        unv = np.asarray(ref[3]) < -1.0e9                                                # lf0 of unvoiced frames is the -1e10 floor
        assert np.all(one[1][unv] == 0.0) and np.all(one[2][unv] == 0.0)
    assert np.all(np.asarray(refs[0][3]) < -1.0e9) and np.all(singles[0][1] == 0.0)


def test_analysis_compressed_batch_and_mag_dim_100(mp):
    """mag_dim=100 is what the shipped low-dim demo uses (demos/demo_copy_synthesis_low_dim.py:63)."""
    utts = [synth_utterance(u, fs=48000, dur_s=0.4) for u in (20, 21)]
    outs = mp.analysis_compressed_batch([u[0] for u in utts], 48000, [u[1] for u in utts], [u[2] for u in utts],
                                        mag_dim=100, phase_dim=45)
    for (sig, pm, voi), got in zip(utts, outs):
        ref = orc.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=100, phase_dim=45)
        for a, b in zip(got[:3], ref[:3]):
            assert a.shape == b.shape and rms(a, b) < TOL
        assert np.array_equal(got[3], ref[3])


def test_analysis_compressed_const_rate(mp):
    """5 ms constant-rate output: the lossless rows are interpolated on the device inside the tile-product loader."""
    utts = [synth_utterance(u, fs=48000, dur_s=0.8) for u in (9, 10)]
    outs = mp.analysis_compressed_batch([u[0] for u in utts], 48000, [u[1] for u in utts], [u[2] for u in utts],
                                        mag_dim=60, phase_dim=45, b_const_rate=True)
    for (sig, pm, voi), got in zip(utts, outs):
        ref = orc.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45, b_const_rate=True)
        for a, b in zip(got[:3], ref[:3]):
            assert a.shape == b.shape and rms(a, b) < TOL
        assert np.array_equal(got[3], ref[3]) and np.array_equal(got[4], ref[4])
    one = mp.analysis_compressed_from_pm(*utts[0][:1], 48000, utts[0][1], utts[0][2], mag_dim=60, phase_dim=45,
                                         b_const_rate=True)
    for a, b in zip(one[:4], outs[0][:4]):
        assert np.array_equal(a, b)


def test_dim_errors(mp):
    sig, pm, voi = synth_utterance(9, fs=48000, dur_s=0.3)
    with pytest.raises(ValueError):
        mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=300, phase_dim=45)
    with pytest.raises(ValueError):
        mp.analysis_compressed_from_pm(sig, 8000, pm, voi)


def test_legacy_analysis_with_del_comp_and_ph_encoding(mp):
    """Named in BASELINE.json north_star: signature (v_in_sig, nFFT, fs, mvf) -> 4-tuple kept; runs on the analysis
    and mel kernels (sp_to_mcep on the device) + host cubic resampling."""
    sig, pm, voi = synth_utterance(14, fs=48000, dur_s=0.5)
    pm_sec = pm / 48000.0
    ref = orc.analysis_with_del_comp_and_ph_encoding_from_pm(sig, 4096, 48000, 4500, pm_sec)
    got = mp.analysis_with_del_comp_and_ph_encoding(sig, 4096, 48000, 4500, pm=pm_sec)
    assert len(got) == 4 and np.array_equal(got[3], ref[3])
    assert rms(got[0], ref[0]) < TOL and got[0].shape == (pm.size, 60)
    # the phase streams are cepstra of 10^(sin/10)-type spectra (in_type=1 applied to values in [-1, 1]):
    # well conditioned, same 1e-5 bar
    assert rms(got[1], ref[1]) < TOL and rms(got[2], ref[2]) < TOL
    with pytest.raises(ValueError):
        mp.analysis_with_del_comp_and_ph_encoding(sig, 512, 48000, 4500, pm=pm_sec)
    # sp_to_mcep alone, all three input types
    mag = orc.analysis_lossless_from_pm(sig, 48000, pm, voi)[0]
    for it, x in ((3, mag), (2, np.log(mag + 1e-3)), (1, 20 * np.log10(mag + 1e-3))):
        assert rms(mp.sp_to_mcep(x, in_type=it), orc.mcep_j0(x, in_type=it)) < TOL
        # la.sp_mel_warp on top of it (src/libaudio.py:643-661): compared in the log domain for in_type 3 (output is |.|)
        w_got, w_ref = mp.sp_mel_warp(x, 60, alpha=0.77, in_type=it), orc.sp_mel_warp(x, 60, alpha=0.77, in_type=it)
        assert w_got.shape == w_ref.shape == (pm.size, 60)
        if it == 3:
            w_got, w_ref = np.log(w_got), np.log(w_ref)
        assert rms(w_got, w_ref) < (1e-4 if it == 1 else TOL), (it, rms(w_got, w_ref))      # dB scale: 8.7 x the log error

"""Parity at the benchmark's batch shape: the device-resident plans that bench.py times (features never leave HBM)
against the host API and the oracle, plus size-independent properties (determinism, batch == single)."""
import warnings

import numpy as np
import pytest

import magphase_oracle as orc
from magphase_b200.synth import synth_utterance

pytestmark = pytest.mark.gpu
FS, N = 48000, 4096


def rms(a, b):
    return float(np.sqrt(np.mean(np.abs(np.asarray(a) - np.asarray(b)) ** 2)))


@pytest.fixture(scope='module')
def batch():
    base = [synth_utterance(u, fs=FS, dur_s=5.0) for u in range(8)]
    return [base[i % 8] for i in range(48)]          # 48 x 5 s, ~43 k frames


def test_compressed_plan_matches_host_api_and_oracle(batch):
    import torch
    import magphase_b200.magphase as mp
    from magphase_b200.device import CompressedPlan
    plan = CompressedPlan([u[0].size for u in batch], [u[1] for u in batch], [u[2] for u in batch], FS, N,
                          mag_dim=60, phase_dim=45, device=0)
    d_sig = torch.from_numpy(np.concatenate([u[0] for u in batch]).astype(np.float32)).cuda()
    mel = [t.clone() for t in plan.analysis(d_sig)]
    y1 = plan.synthesis().clone()
    torch.cuda.synchronize()
    # bit-reproducible: a second pass over the same plan gives identical bytes
    mel2 = plan.analysis(d_sig)
    y2 = plan.synthesis()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(mel, mel2)) and torch.equal(y1, y2)
    assert torch.isfinite(y1).all()
    frm_off = np.concatenate(([0], np.cumsum([len(l) for l in plan.l_lf0])))
    out_off = plan._keep['utt_out_off'].cpu().numpy()
    y1 = y1.cpu().numpy()
    for u in (0, 13, 47):
        sig, pm, voi = batch[u]
        a, b = frm_off[u], frm_off[u + 1]
        host = mp.analysis_compressed_from_pm(sig, FS, pm, voi, mag_dim=60, phase_dim=45)
        for d, h in zip(mel, host[:3]):
            assert rms(d[a:b].cpu().numpy(), h) < 1e-6                      # float32 storage vs float64 host arrays
        assert np.array_equal(plan.l_lf0[u], host[3])
        y_host = mp.synthesis_from_compressed_batch([host[:4]], FS, b_out_hpf=False, l_noise=[plan.h_noise[u]])[0]
        y_dev = y1[out_off[u]:out_off[u + 1]]
        assert y_dev.shape == y_host.shape and rms(y_dev, y_host) < 1e-5
    # one utterance end to end against the oracle (the CPU restatement of the reference), same noise
    u = 5
    sig, pm, voi = batch[u]
    ref = orc.analysis_compressed_from_pm(sig, FS, pm, voi, mag_dim=60, phase_dim=45)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y_ref = orc.synthesis_from_compressed(ref[0], ref[1], ref[2], ref[3], FS, b_out_hpf=False, v_noise=plan.h_noise[u])
    assert rms(y1[out_off[u]:out_off[u + 1]], y_ref) < 1e-5


def test_lossless_plan_matches_host_api(batch):
    import torch
    import magphase_b200.magphase as mp
    from magphase_b200 import _lib
    from magphase_b200.device import LosslessPlan
    sub = batch[:24]
    plan = LosslessPlan([u[0].size for u in sub], [u[1] for u in sub], [u[2] for u in sub], FS, N, device=0)
    d_sig = torch.from_numpy(np.concatenate([u[0] for u in sub]).astype(np.float32)).cuda()
    feats = plan.alloc_features(_lib.MPB_F32)
    out = plan.alloc_output(_lib.MPB_F32)
    plan.analysis(d_sig, feats, compute=_lib.MPB_F64)
    plan.synthesis(feats, out, compute=_lib.MPB_F32)
    torch.cuda.synchronize()
    y = out.cpu().numpy()
    for u in (0, 23):
        sig, pm, voi = sub[u]
        a, b = plan.frm_off[u], plan.frm_off[u + 1]
        mag, real, imag, f0, fs, v_shift = mp.analysis_lossless_from_pm(sig, FS, pm, voi)
        assert np.array_equal(v_shift, plan.l_shift[u]) and np.array_equal(f0, plan.l_f0[u])
        assert rms(feats[0][a:b].cpu().numpy(), mag) / rms(mag, 0 * mag) < 1e-6
        assert rms(feats[1][a:b].cpu().numpy(), real) < 1e-5 and rms(feats[2][a:b].cpu().numpy(), imag) < 1e-5
        y_host = mp.synthesis_from_lossless(mag, real, imag, f0, fs)
        y_dev = y[plan.out_off[u]:plan.out_off[u + 1]]
        assert y_dev.shape == y_host.shape and rms(y_dev, y_host) < 1e-6
        # copy synthesis reproduces the analysed waveform where the two side windows overlap-add to one:
        # voiced stretches re-placed on the f0 grid differ by design, so compare the oracle chain instead
    ref = orc.synthesis_from_lossless(*orc.analysis_lossless_from_pm(*sub[3][:1], FS, sub[3][1], sub[3][2])[:4], FS)
    assert rms(y[plan.out_off[3]:plan.out_off[4]], ref) < 1e-6

"""The EXPERIMENTAL tensor-core tile product (magphase_b200/csrc/mpb_mel_tc.cu, k_mel_gemm_tc: tcgen05.mma kind::tf32
with hi + lo operand splitting, TMEM accumulators) against the oracle, at the bar of the shipped path (1e-5 RMS).

The kernel is off by default and is selected when a plan is created (environment MPB_MEL_TC=1), so the check runs in a
child process.  It was written after the GPU budget of round 1 was spent and has seen one GPU run (it passed,
profiles/r1b/mel_tc_first_run.txt); until it has been measured and run through the whole suite its outcome is recorded
as xpass / xfail instead of gating the suite.  Bits 2 and 3 are small variations of the same kernel (never run); the un-warp variant (bit 1) has never run either and
is not exercised here."""
import os
import subprocess
import sys

import pytest

from conftest import have_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/oracle')
import magphase_oracle as orc
import magphase_b200.magphase as mp
from magphase_b200.synth import synth_utterance
rms = lambda a, b: float(np.sqrt(np.mean((np.asarray(a) - np.asarray(b)) ** 2)))
utts = [synth_utterance(u, fs=48000, dur_s=1.0) for u in (8, 9, 10, 11)]
got = mp.analysis_compressed_batch([u[0] for u in utts], 48000, [u[1] for u in utts], [u[2] for u in utts], mag_dim=60, phase_dim=45)
worst = 0.0
for (sig, pm, voi), g in zip(utts, got):
    ref = orc.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    assert np.array_equal(g[3], ref[3]) and np.array_equal(g[4], ref[4])
    for a, b in zip(g[:3], ref[:3]):
        assert a.shape == b.shape
        worst = max(worst, rms(a, b))
print('WORST_RMS %%.3e' %% worst)
assert worst < 1e-5, worst
'''


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason='experimental tcgen05 tile product: one GPU run so far, not yet part of the shipped path')
@pytest.mark.parametrize('mask', [1, 5, 9, 13])    # 1: as run once; +4: deeper load pipeline; +8: K-slice sums in the epilogue
def test_tensor_core_warp_product_vs_oracle(mask):
    if not have_cuda():
        pytest.skip('needs a CUDA device')
    env = dict(os.environ, MPB_MEL_TC=str(mask))
    r = subprocess.run([sys.executable, '-c', CHILD % {'root': ROOT}], env=env, capture_output=True, text=True, timeout=180)
    print(r.stdout[-500:], r.stderr[-1500:])
    assert r.returncode == 0 and 'WORST_RMS' in r.stdout

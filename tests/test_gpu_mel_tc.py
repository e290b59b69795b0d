"""The two code paths of the mel tile products against each other and against the oracle.

Default: tcgen05 kernels (mpb_mel_warp_tc.cu / mpb_mel_unwarp_tc.cu, "3xTF32" with sliced accumulators).  MPB_MEL_TC=0 at plan
creation selects the CUDA-core FMA kernels (mpb_mel.cu / mpb_unwarp.cu), which also serve every case the tensor-core tiles do
not cover (more than 64 coefficients, float64 feature rows, constant-rate interpolation, raw cepstra).  The switch is read when
a plan is created, so each setting runs in a child process.  Both must meet the 1e-5 RMS bar on their own; the test also
records how far apart they are."""
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
worst, worst_syn = 0.0, 0.0
for k, ((sig, pm, voi), g) in enumerate(zip(utts, got)):
    ref = orc.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    assert np.array_equal(g[3], ref[3]) and np.array_equal(g[4], ref[4])
    for a, b in zip(g[:3], ref[:3]):
        assert a.shape == b.shape
        worst = max(worst, rms(a, b))
    np.random.seed(50 + k)
    y = mp.synthesis_from_compressed(*ref[:4], 48000, b_out_hpf=False)
    np.random.seed(50 + k)
    y_ref = orc.synthesis_from_compressed(*ref[:4], 48000, b_out_hpf=False)
    assert y.shape == y_ref.shape
    worst_syn = max(worst_syn, rms(y, y_ref))
print('WORST_RMS %%.3e %%.3e' %% (worst, worst_syn))
assert worst < 1e-5 and worst_syn < 1e-5, (worst, worst_syn)
'''


@pytest.mark.gpu
@pytest.mark.parametrize('mask', ['3', '0', '1', '2'])    # both tensor-core products (default), none, warp only, un-warp only
def test_tile_product_paths_vs_oracle(mask):
    if not have_cuda():
        pytest.skip('needs a CUDA device')
    env = dict(os.environ, MPB_MEL_TC=mask)
    r = subprocess.run([sys.executable, '-c', CHILD % {'root': ROOT}], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout[-500:], r.stderr[-1500:])
    assert r.returncode == 0 and 'WORST_RMS' in r.stdout

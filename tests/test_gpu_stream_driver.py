"""batch.run_chain_stream: batches taken round-robin by several host threads with private library contexts must give
exactly what one thread gives, whatever the host wait mode (MPB_SYNC is read once per process: child processes) -- and the
results must not depend on which worker ran which batch (the noise of batch k comes from RandomState(seed + k))."""
import os
import subprocess
import sys

import pytest

from conftest import have_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, hashlib, numpy as np
sys.path.insert(0, %(root)r)
from magphase_b200 import _lib
from magphase_b200.batch import run_chain_stream
from magphase_b200.synth import synth_utterance
FS = 48000
pool = [synth_utterance(700 + i, fs=FS, dur_s=d) for i, d in enumerate((0.6, 1.0, 1.4, 0.8, 1.2))]
batches = [[pool[(3 * b + j) %% len(pool)] for j in range(1 + (b %% 4))] for b in range(9)]      # 1..4 utterances per batch
def digest(r):
    h = hashlib.sha256()
    for feats, ys in r['outputs']:
        for f in feats:
            for a in f[:5]:
                h.update(np.ascontiguousarray(a).tobytes())
        for y in ys:
            h.update(np.ascontiguousarray(y).tobytes())
    return h.hexdigest()
one = run_chain_stream(batches, FS, n_workers=1, seed=11, keep_outputs=True)
four = run_chain_stream(batches, FS, n_workers=4, seed=11, keep_outputs=True)
gated = run_chain_stream(batches, FS, n_workers=3, seed=11, keep_outputs=True, n_inflight=1)
assert one['frames'] == four['frames'] == gated['frames'] > 0
d1, d4, dg = digest(one), digest(four), digest(gated)
assert d1 == d4 == dg, (d1, d4, dg)
again = run_chain_stream(batches, FS, n_workers=4, seed=12, keep_outputs=True)
assert digest(again) != d1                                   # a different seed is a different noise stream
y = np.concatenate([y for _, ys in four['outputs'] for y in ys])
assert np.isfinite(y).all() and float(np.abs(y).max()) > 1e-3
print('DIGEST', d1, 'pool', _lib.pinned.stats)
'''


@pytest.mark.gpu
@pytest.mark.parametrize('sync', ['spin', 'block'])
def test_stream_driver_workers_agree(sync):
    if not have_cuda():
        pytest.skip('needs a CUDA device')
    env = dict(os.environ, MPB_SYNC=sync)
    r = subprocess.run([sys.executable, '-c', CHILD % {'root': ROOT}], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-600:], r.stderr[-2000:])
    assert r.returncode == 0 and 'DIGEST' in r.stdout
    test_stream_driver_workers_agree.digests = getattr(test_stream_driver_workers_agree, 'digests', {})
    test_stream_driver_workers_agree.digests[sync] = r.stdout.split('DIGEST')[1].split()[0]
    d = test_stream_driver_workers_agree.digests
    if len(d) == 2:
        assert d['spin'] == d['block']                       # the wait mode never changes a byte

"""Fine-grained wall-clock breakdown of the compressed e2e arm: MPB_TRACE=1 prints the C-side stages, this script times
the Python-side pieces around them."""
import os, sys, time
os.environ['MPB_TRACE'] = '1'
sys.path.insert(0, '.')
import numpy as np
import magphase_b200.magphase as mp
from magphase_b200.synth import synth_utterance
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
base = [synth_utterance(u) for u in range(8)]
utts = [base[i % 8] for i in range(n)]
sig, pm, voi = [u[0] for u in utts], [u[1] for u in utts], [u[2] for u in utts]
for k in range(4):
    sys.stderr.write('--- step %d\n' % k)
    t = time.perf_counter()
    outs = mp.analysis_compressed_batch(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    t1 = time.perf_counter()
    ys = mp.synthesis_from_compressed_batch([o[:4] for o in outs], 48000, b_out_hpf=False)
    t2 = time.perf_counter()
    sys.stderr.write('python: analysis %.3f ms, synthesis %.3f ms\n' % (1e3 * (t1 - t), 1e3 * (t2 - t1)))
feats = [o[:4] for o in outs]
t = time.perf_counter()
arrs, l = mp.compressed_synthesis_geometry([f[3] for f in feats], [f[0].shape[0] for f in feats], 48000, 4096)
sys.stderr.write('geometry %.3f ms\n' % (1e3 * (time.perf_counter() - t)))

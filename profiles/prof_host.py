"""Host-path breakdown of the e2e arms (run on the GPU box): wall time per API call, device time per kernel,
cProfile top entries.  Usage: python profiles/prof_host.py [compressed|lossless] [n_utts]"""
import cProfile
import pstats
import sys
import time

sys.path.insert(0, '.')
import numpy as np
import magphase_b200.magphase as mp
from magphase_b200 import _lib
from magphase_b200.synth import synth_utterance

mode = sys.argv[1] if len(sys.argv) > 1 else 'compressed'
n = int(sys.argv[2]) if len(sys.argv) > 2 else (32 if mode == 'compressed' else 8)
base = [synth_utterance(u) for u in range(8)]
utts = [base[i % 8] for i in range(n)]
sig, pm, voi = [u[0] for u in utts], [u[1] for u in utts], [u[2] for u in utts]

if mode == 'compressed':
    ana = lambda: mp.analysis_compressed_batch(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    syn = lambda outs: mp.synthesis_from_compressed_batch([o[:4] for o in outs], 48000, b_out_hpf=False)
    nf = lambda outs: sum(o[4].size for o in outs)
else:
    ana = lambda: mp.analysis_lossless_batch(sig, 48000, pm, voi)
    syn = lambda outs: mp.synthesis_from_lossless_batch([o[:4] for o in outs], 48000)
    nf = lambda outs: sum(o[5].size for o in outs)

for _ in range(3):
    outs = ana(); ys = syn(outs)
T = {'analysis': 0.0, 'synthesis': 0.0}
K = 5
for _ in range(K):
    t = time.perf_counter(); outs = ana(); T['analysis'] += time.perf_counter() - t
    t = time.perf_counter(); ys = syn(outs); T['synthesis'] += time.perf_counter() - t
frames = nf(outs)
print(mode, 'utts', n, 'frames', frames, {k: round(1e3 * v / K, 2) for k, v in T.items()}, 'ms; e2e frames/s',
      round(frames / (sum(T.values()) / K)))
_lib.profile_begin()
outs = ana(); ys = syn(outs)
prof = _lib.profile_end()
print('device ms:', {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
      'total', round(sum(v[1] for v in prof.values()), 3))
for name, f in (('analysis', ana), ('synthesis', lambda: syn(outs))):
    pr = cProfile.Profile(); pr.enable(); f(); pr.disable()
    print('----', name)
    pstats.Stats(pr).sort_stats('tottime').print_stats(10)

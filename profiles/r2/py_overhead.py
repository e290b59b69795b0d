"""How much of an analysis / synthesis call is Python (GIL held) and how much is the C entry point (GIL released)?"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import magphase_b200.magphase as mp
from magphase_b200 import _lib
from magphase_b200.synth import synth_utterance
base = [synth_utterance(u) for u in range(8)]
utts = [(np.round(base[i % 8][0] * 32768.0).astype(np.int16), base[i % 8][1], base[i % 8][2]) for i in range(128)]
sigs, pms, vois = [u[0] for u in utts], [u[1] for u in utts], [u[2] for u in utts]
acc = {}
real_lib = _lib.lib()
def wrap(name):
    f = getattr(real_lib, name)
    def g(*a):
        t = time.perf_counter(); r = f(*a); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
    return g
class L:
    def __init__(self, l): self.l = l
    def __getattr__(self, n): return wrap(n)
_lib.lib = lambda: L(real_lib)
for rep in range(4):
    acc.clear()
    t0 = time.perf_counter()
    outs = mp.analysis_compressed_batch(sigs, 48000, pms, vois, mag_dim=60, phase_dim=45, out_dtype=np.float32)
    t1 = time.perf_counter()
    a_c = sum(acc.values()); a_parts = dict(acc); acc.clear()
    ys = mp.synthesis_from_compressed_batch([o[:4] for o in outs], 48000, b_out_hpf=False, out_dtype=np.float32, rng=np.random.RandomState(1))
    t2 = time.perf_counter()
    s_c = sum(acc.values())
    print('analysis %.2f ms (C %.2f, Python %.2f)   synthesis %.2f ms (C %.2f, Python %.2f)' %
          (1e3 * (t1 - t0), 1e3 * a_c, 1e3 * (t1 - t0 - a_c), 1e3 * (t2 - t1), 1e3 * s_c, 1e3 * (t2 - t1 - s_c)))
print({k: round(1e3 * v, 2) for k, v in a_parts.items()}, {k: round(1e3 * v, 2) for k, v in acc.items()})

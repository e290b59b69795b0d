import sys, time
sys.path.insert(0, '.')
import numpy as np
import magphase_b200.magphase as mp
from magphase_b200.synth import synth_utterance
utts = [synth_utterance(u) for u in range(8)] * 4
sig, pm, voi = [u[0] for u in utts], [u[1] for u in utts], [u[2] for u in utts]
def step():
    outs = mp.analysis_compressed_batch(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    ys = mp.synthesis_from_compressed_batch([o[:4] for o in outs], 48000, b_out_hpf=False)
    return outs, ys
for _ in range(3): step()
import cProfile, pstats
T = {}
def timed(name, f):
    t = time.perf_counter(); r = f(); T[name] = T.get(name, 0) + time.perf_counter() - t; return r
for _ in range(5):
    outs = timed('analysis', lambda: mp.analysis_compressed_batch(sig, 48000, pm, voi, mag_dim=60, phase_dim=45))
    ys = timed('synthesis', lambda: mp.synthesis_from_compressed_batch([o[:4] for o in outs], 48000, b_out_hpf=False))
print({k: v / 5 for k, v in T.items()}, 'frames', sum(o[4].size for o in outs))
# finer: time pieces of synthesis
feats = [o[:4] for o in outs]
t = time.perf_counter(); arrs, l = mp.compressed_synthesis_geometry([f[3] for f in feats], [f[0].shape[0] for f in feats], 48000, 4096); print('geometry', time.perf_counter() - t)
t = time.perf_counter(); out = np.empty(int(arrs['utt_out_off'][-1])); out[:] = 0; print('alloc+touch out', time.perf_counter() - t, out.size)
t = time.perf_counter(); st = np.random.get_state(); np.random.set_state(st); print('get/set state', time.perf_counter() - t)
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(8)

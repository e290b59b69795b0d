import cProfile, pstats, sys, time
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
step(); step()
t = time.perf_counter(); outs, ys = step(); dt = time.perf_counter() - t
print('frames', sum(o[4].size for o in outs), 'secs', dt)
t = time.perf_counter(); outs = mp.analysis_compressed_batch(sig, 48000, pm, voi, mag_dim=60, phase_dim=45); print('analysis', time.perf_counter() - t)
t = time.perf_counter(); ys = mp.synthesis_from_compressed_batch([o[:4] for o in outs], 48000, b_out_hpf=False); print('synthesis', time.perf_counter() - t)
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)

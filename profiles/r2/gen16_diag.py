"""cProfile of one config-4 step (16 kHz constant-rate features -> post_filter -> synthesis with HPF) on the GPU box."""
import cProfile, pstats, sys, time
sys.path.insert(0, '.')
import numpy as np
import magphase_b200.magphase as mp
from magphase_b200.synth import synth_utterance
fs16, n16 = 16000, 128
base16 = [synth_utterance(900 + u, fs=fs16, dur_s=5.0) for u in range(8)]
outs16 = mp.analysis_compressed_batch([b[0] for b in base16], fs16, [b[1] for b in base16], [b[2] for b in base16], mag_dim=60,
                                      phase_dim=45, b_const_rate=True)
feats16 = [tuple(np.array(a) for a in outs16[i % 8][:4]) for i in range(n16)]
def g16():
    rows = np.concatenate([f[0] for f in feats16], axis=0)
    rows = mp.post_filter(rows, fs16)
    off = np.concatenate(([0], np.cumsum([f[0].shape[0] for f in feats16])))
    fl = [(rows[off[i]:off[i + 1]],) + f[1:] for i, f in enumerate(feats16)]
    return mp.synthesis_from_compressed_batch(fl, fs16, b_const_rate=True, b_out_hpf=True)
for _ in range(2): g16()
t = time.perf_counter(); g16(); print('step %.1f ms' % (1e3 * (time.perf_counter() - t)))
pr = cProfile.Profile(); pr.enable(); g16(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(16)

import sys, time, numpy as np
sys.path.insert(0, '.')
import magphase_b200.magphase as mp
from magphase_b200.synth import synth_utterance
FS = 48000
durs = [2.0, 3.0, 4.0, 5.0, 6.0, 8.0]
pool = [synth_utterance(5000 + i, fs=FS, dur_s=d) for i, d in enumerate(durs)]
pcm_pool = [(np.round(u[0] * 32768.0).astype(np.int16), u[1], u[2]) for u in pool]
order = np.random.Generator(np.random.PCG64(7)).integers(0, len(pool), 1024)
order = sorted(order.tolist(), key=lambda i: -pool[i][0].size)
batches = [[pcm_pool[i] for i in order[k:k + 128]] for k in range(0, 1024, 128)]
for rep in range(2):
    for k, b in enumerate(batches):
        t0 = time.perf_counter()
        outs = mp.analysis_compressed_batch([u[0] for u in b], FS, [u[1] for u in b], [u[2] for u in b], mag_dim=60, phase_dim=45, out_dtype=np.float32)
        t1 = time.perf_counter()
        ys = mp.synthesis_from_compressed_batch([o[:4] for o in outs], FS, b_out_hpf=False, out_dtype=np.float32, rng=np.random.RandomState(k))
        t2 = time.perf_counter()
        fr = sum(o[4].size for o in outs)
        print('rep %d batch %d: %6d frames  analysis %.1f ms  synthesis %.1f ms  -> %.2f M frames/s' % (rep, k, fr, 1e3 * (t1 - t0), 1e3 * (t2 - t1), fr / (t2 - t0) / 1e6))
        del outs, ys

"""Which half of the host chain saturates first?  4 worker threads (private contexts), 128-utterance batches: analysis calls
only, synthesis calls only, both (= run_chain_stream)."""
import concurrent.futures as cf, sys, time
sys.path.insert(0, '.')
import numpy as np
import magphase_b200.magphase as mp
from magphase_b200 import _lib
from magphase_b200.synth import synth_utterance
FS = 48000
base = [synth_utterance(u) for u in range(8)]
utts = [(np.round(base[i % 8][0] * 32768.0).astype(np.int16), base[i % 8][1], base[i % 8][2]) for i in range(128)]
sigs, pms, vois = [u[0] for u in utts], [u[1] for u in utts], [u[2] for u in utts]
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 24
NW = int(sys.argv[2]) if len(sys.argv) > 2 else 4
_lib.pinned.ensure(3 << 30)
feats = {}
def run(mode):
    def work(w):
        _lib.set_thread_slot(w)
        if w not in feats:
            feats[w] = [tuple(np.array(a) for a in o[:4]) for o in mp.analysis_compressed_batch(sigs, FS, pms, vois, mag_dim=60, phase_dim=45, out_dtype=np.float32)]
        fr = 0
        for k in range(w, NB, NW):
            if mode in ('ana', 'both'):
                outs = mp.analysis_compressed_batch(sigs, FS, pms, vois, mag_dim=60, phase_dim=45, out_dtype=np.float32)
                f = [o[:4] for o in outs]
            else:
                f = feats[w]
            if mode in ('syn', 'both'):
                ys = mp.synthesis_from_compressed_batch(f, FS, b_out_hpf=False, out_dtype=np.float32, rng=np.random.RandomState(k))
            fr += sum(x[0].shape[0] for x in f)
            outs = ys = None
        return fr
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(NW) as ex:
        fr = sum(ex.map(work, range(NW)))
    return fr, time.perf_counter() - t0
for mode in ('both', 'ana', 'syn', 'both'):
    run(mode)
    fr, dt = run(mode)
    print('%-5s %d workers: %.2f M frames/s  (%.2f ms per batch)' % (mode, NW, fr / dt / 1e6, 1e3 * dt / NB), flush=True)

"""Does the GPU keep its throughput when several host threads drive device-resident chains on their own streams?
(the end-to-end arm saturates at ~17 M frames/s whatever the number of workers; the device-resident chain alone does 23.8 M)"""
import sys, threading, time
sys.path.insert(0, '.')
import numpy as np, torch
from magphase_b200.device import CompressedPlan
from magphase_b200.synth import synth_utterance
from magphase_b200 import _lib
FS, N = 48000, 4096
base = [synth_utterance(u) for u in range(8)]
utts = [base[i % 8] for i in range(128)]
sig = torch.from_numpy(np.concatenate([u[0] for u in utts]).astype(np.float32)).cuda()
geom = ([u[0].size for u in utts], [u[1] for u in utts], [u[2] for u in utts])
K = 12
for nw in (1, 2, 4):
    plans, streams = [], []
    def make(w):
        _lib.set_thread_slot(w)
        plans.append(CompressedPlan(*geom, FS, N, mag_dim=60, phase_dim=45, device=0)); streams.append(torch.cuda.Stream())
    ts = [threading.Thread(target=make, args=(w,)) for w in range(nw)]
    [t.start() or t.join() for t in ts]
    def work(w, k):
        _lib.set_thread_slot(w)
        with torch.cuda.stream(streams[w]):
            for _ in range(k):
                plans[w].chain(sig)
        streams[w].synchronize()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ts = [threading.Thread(target=work, args=(w, K)) for w in range(nw)]
        [t.start() for t in ts]; [t.join() for t in ts]
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('threads %d: %.2f M frames/s aggregate (%.2f ms per 128-utterance chain)' % (nw, nw * K * plans[0].nfrm / dt / 1e6, 1e3 * dt / (nw * K)), flush=True)
    del plans, streams
    torch.cuda.empty_cache()

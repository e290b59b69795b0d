"""End-to-end streaming throughput against the number of host worker threads (each with a private context), for identical
batches (the bench's e2e workload) and for the mixed-duration stream of the config-5 record."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np
from magphase_b200.batch import run_chain_stream
from magphase_b200 import _lib
from magphase_b200.synth import synth_utterance
FS = 48000
base = [synth_utterance(u) for u in range(8)]
utts = [(np.round(base[i % 8][0] * 32768.0).astype(np.int16), base[i % 8][1], base[i % 8][2]) for i in range(128)]
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 12
only = sys.argv[2] if len(sys.argv) > 2 else ''
gates = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0]           # e.g. '128:2' or 'mixed:1' (with MPB_TRACE=1 in the environment)
_last = dict(_lib.pinned.stats)
def pool_delta():
    global _last
    cur = dict(_lib.pinned.stats)
    d = {k: (round(cur[k] - _last[k], 3) if isinstance(cur[k], float) else cur[k] - _last[k]) for k in cur}
    _last = cur
    return d
for bs in (128,):
    batches = [utts[:bs]] * (nb * 128 // bs)
    for nw in (1, 2, 3, 4, 6):
        if only and only != '%d:%d' % (bs, nw):
            continue
        for g in gates:
            if g >= nw:
                continue
            run_chain_stream(batches, FS, n_workers=nw, n_inflight=g or nw)
            pool_delta()
            sys.stderr.write('=== timed pass\n')
            r = run_chain_stream(batches, FS, n_workers=nw, n_inflight=g or nw)
            print('identical batches of %3d: workers %d gate %d  %.2f M frames/s  (%.2f ms per 128 utterances)  pool %s'
                  % (bs, nw, g, r['frames'] / r['seconds'] / 1e6, 1e3 * r['seconds'] / nb, pool_delta()), flush=True)
durs = [2.0, 3.0, 4.0, 5.0, 6.0, 8.0]
pool = [synth_utterance(5000 + i, fs=FS, dur_s=d) for i, d in enumerate(durs)]
pcm = [(np.round(u[0] * 32768.0).astype(np.int16), u[1], u[2]) for u in pool]
order = np.random.Generator(np.random.PCG64(7)).integers(0, len(pool), 1024)
order = sorted(order.tolist(), key=lambda i: -pool[i][0].size)
batches = [[pcm[i] for i in order[k:k + 128]] for k in range(0, 1024, 128)]
for nw in (1, 2, 3):
    if only and only != 'mixed:%d' % nw:
        continue
    run_chain_stream(batches, FS, n_workers=nw)
    pool_delta()
    sys.stderr.write('=== timed pass\n')
    r = run_chain_stream(batches, FS, n_workers=nw)
    print('mixed durations: workers %d  %.2f M frames/s  pool %s' % (nw, r['frames'] / r['seconds'] / 1e6, pool_delta()), flush=True)

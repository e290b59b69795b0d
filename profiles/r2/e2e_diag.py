"""Host-path breakdown of the narrow e2e arm (PCM16 in, float32 out), one thread: MPB_TRACE stage timings on stderr,
wall time per call, cProfile of both calls."""
import cProfile, os, pstats, sys, time
os.environ['MPB_TRACE'] = '1'
sys.path.insert(0, '.')
import numpy as np
import magphase_b200.magphase as mp
from magphase_b200.synth import synth_utterance
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
base = [synth_utterance(u) for u in range(8)]
utts = [base[i % 8] for i in range(n)]
sig = [np.round(u[0] * 32768.0).astype(np.int16) for u in utts]
pm, voi = [u[1] for u in utts], [u[2] for u in utts]
ana = lambda: mp.analysis_compressed_batch(sig, 48000, pm, voi, mag_dim=60, phase_dim=45, out_dtype=np.float32)
syn = lambda outs: mp.synthesis_from_compressed_batch([o[:4] for o in outs], 48000, b_out_hpf=False, out_dtype=np.float32,
                                                      rng=np.random.RandomState(1))
for k in range(5):
    sys.stderr.write('--- step %d\n' % k)
    t = time.perf_counter(); outs = ana(); t1 = time.perf_counter(); ys = syn(outs); t2 = time.perf_counter()
    fr = sum(o[4].size for o in outs)
    sys.stderr.write('python: analysis %.3f ms, synthesis %.3f ms  (%d frames, %.2f M frames/s)\n'
                     % (1e3 * (t1 - t), 1e3 * (t2 - t1), fr, fr / (t2 - t) / 1e6))
os.environ['MPB_TRACE'] = '0'
for name, f in (('analysis', ana), ('synthesis', lambda: syn(outs))):
    pr = cProfile.Profile(); pr.enable(); f(); pr.disable()
    print('----', name)
    pstats.Stats(pr).sort_stats('tottime').print_stats(14)

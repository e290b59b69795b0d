#!/usr/bin/env python
"""Benchmark of the MagPhase analysis+synthesis hot path (BASELINE.json metric: frames/sec at 48 kHz, FFT 4096).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One *step* = one pass of analysis -> synthesis over one batch of synthetic 48 kHz utterances ("synth48k-v1",
magphase_b200/synth.py).  `value` is device-timed with signals, descriptors and features resident in HBM;
`e2e` goes through the public host API (NumPy in, NumPy out, H2D/D2H inside the timed region).
Under torchrun every rank runs its own shard of utterances (weak scaling, no data-path collective: the only
NCCL traffic is the scattered work list and the final counters).
"""
import os

# The CPU arm forks one worker per core (the reference's own fan-out, src/libutils.py:32-63): every worker must run its BLAS
# single-threaded or 32 workers x 32 BLAS threads oversubscribe the host.  OpenBLAS / MKL / OpenMP size their pools when numpy
# is IMPORTED, so the variables are set before that import (round 1 set them afterwards: the N=1 reference value was 6x low).
for _k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
    os.environ[_k] = '1'

import argparse
import json
import multiprocessing
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'analysis+synthesis frames/sec at 48 kHz FFT=4096'
FS = 48000
FFT_LEN = 4096
DISTINCT = 16          # distinct synthetic utterances generated per rank (tiled up to --utts)

# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref, py3 translation) or the numpy oracle port
# --------------------------------------------------------------------------------------------
_CPU_INPUTS = None
_CPU_WORKLOAD = 'compressed'


def _cpu_impl():
    """Returns (kind, analysis_fn, synthesis_fn).  The ONLY place bench.py touches oracle/."""
    odir = os.path.join(ROOT, 'oracle')
    ref_dir = os.path.join(odir, '_ref')
    import warnings
    warnings.simplefilter('ignore')
    if os.path.exists(os.path.join(ref_dir, 'magphase.py')):
        for p in (odir, ref_dir):
            if p not in sys.path:
                sys.path.insert(0, p)
        import magphase as ref_mp

        def ana(sig, pm, voi):
            m_fft, v_shift = ref_mp.analysis_with_del_comp_from_pm(sig, FS, pm)
            return ref_mp.compute_lossless_feats(m_fft, v_shift, voi, FS)

        def syn(mag, real, imag, f0):
            return ref_mp.synthesis_from_lossless(mag, real, imag, f0, FS)
        return 'reference', ana, syn
    if odir not in sys.path:
        sys.path.insert(0, odir)
    import magphase_oracle as orc

    def ana(sig, pm, voi):
        return orc.analysis_lossless_from_pm(sig, FS, pm, voi)[:4]

    def syn(mag, real, imag, f0):
        return orc.synthesis_from_lossless(mag, real, imag, f0, FS)
    return 'port', ana, syn


def _cpu_impl_compressed():
    """(kind, fn) for the compressed chain: reference lossless analysis -> format_for_modelling with the numpy
    restatement of SPTK `mcep -j 0` (the binary is not available anywhere) -> reference synthesis_from_compressed."""
    kind, ana, _ = _cpu_impl()
    odir = os.path.join(ROOT, 'oracle')
    if odir not in sys.path:
        sys.path.insert(0, odir)
    import magphase_oracle as orc
    if kind == 'reference':
        import magphase as ref_mp
        syn_c = lambda a, b, c, d: ref_mp.synthesis_from_compressed(a, b, c, d, FS, b_out_hpf=False)
    else:
        syn_c = lambda a, b, c, d: orc.synthesis_from_compressed(a, b, c, d, FS, b_out_hpf=False)

    def chain(sig, pm, voi):
        mag, real, imag, f0 = ana(sig, pm, voi)
        mm, rr, ii, lf0 = orc.format_for_modelling(mag, real, imag, f0, FS, mag_dim=60, phase_dim=45)
        return mag.shape[0], syn_c(mm, rr, ii, lf0)
    return kind, chain


def _cpu_worker(i):
    sig, pm, voi = _CPU_INPUTS[i]
    if _CPU_WORKLOAD == 'compressed':
        n, y = _cpu_impl_compressed()[1](sig, pm, voi)
        return int(n), float(y[0])
    kind, ana, syn = _cpu_impl()
    mag, real, imag, f0 = ana(sig, pm, voi)
    y = syn(mag, real, imag, f0)
    return int(mag.shape[0]), float(y[0])


def _cpu_init(workload):
    """Pool initializer: import the CPU implementation and build its one-time tables (the oracle's freqt matrices)
    so that no timed pass pays for them."""
    global _CPU_WORKLOAD
    _CPU_WORKLOAD = workload
    try:                                   # belt and braces: also cap pools a library may have sized before the fork
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    from magphase_b200.synth import synth_utterance
    sig, pm, voi = synth_utterance(99, fs=FS, dur_s=0.3)
    if workload == 'compressed':
        _cpu_impl_compressed()[1](sig, pm, voi)
    else:
        _, ana, syn = _cpu_impl()
        syn(*ana(sig, pm, voi))


def cpu_pass(pool, n_utts):
    t = time.perf_counter()
    res = pool.map(_cpu_worker, range(n_utts), chunksize=1)
    dt = time.perf_counter() - t
    return sum(r[0] for r in res), dt


def make_cpu_inputs(n_utts, dur_s):
    from magphase_b200.synth import synth_utterance
    global _CPU_INPUTS
    base = [synth_utterance(u, fs=FS, dur_s=dur_s) for u in range(min(n_utts, DISTINCT))]
    _CPU_INPUTS = [base[i % len(base)] for i in range(n_utts)]


def run_cpu_arm(n_utts, dur_s, steps, warmup, workload='compressed'):
    """K timed steps (after W warm-ups), each a Pool.map over n_utts utterances on all host cores --
    the reference's own fan-out (src/libutils.py:32-63)."""
    global _CPU_WORKLOAD
    _CPU_WORKLOAD = workload
    kind = _cpu_impl()[0]
    make_cpu_inputs(n_utts, dur_s)
    cores = os.cpu_count() or 1
    ctx = multiprocessing.get_context('fork')
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(workload,)) as pool:
        for _ in range(warmup):
            cpu_pass(pool, n_utts)
        frames, secs = 0, 0.0
        for _ in range(steps):
            f, dt = cpu_pass(pool, n_utts)
            frames += f
            secs += dt
    return dict(kind=kind, cores=cores, n_utts=n_utts, dur_s=dur_s, frames_per_step=frames // max(steps, 1), value=frames / secs,
                ms_per_step=1e3 * secs / max(steps, 1),
                sample=('%d x %.1f s synth48k-v1 utterances per step, Pool(%d): ' % (n_utts, dur_s, cores)) + (
                    'analysis_with_del_comp_from_pm + compute_lossless_feats + format_for_modelling(60/45/45; SPTK mcep '
                    'step = numpy restatement, binary unavailable) + synthesis_from_compressed(b_out_hpf=False)'
                    if workload == 'compressed' else
                    'analysis_with_del_comp_from_pm + compute_lossless_feats + synthesis_from_lossless'))


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from before the warm-up; stop(t0, t1) keeps the samples whose timestamps fall
    inside the timed region [t0, t1] (host wall clock), or the nearest ones when the region is shorter than a period."""
    QUERY = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.QUERY,
                                       '--format=csv,noheader,nounits', '-lms', '20'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, t0=None, t1=None):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        import datetime
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        samples = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), '%Y/%m/%d %H:%M:%S.%f').timestamp()
                samples.append((ts, float(r[1]), float(r[2]), [n for n, v in zip(names, r[5:9]) if v.strip().lower() == 'active']))
            except (ValueError, IndexError):
                continue
        if not samples:
            return None
        inside = [x for x in samples if t0 is not None and t0 <= x[0] <= t1]
        note = 'samples inside the timed region'
        if not inside:          # region shorter than the polling period: the two samples that bracket it
            mid = 0.5 * (t0 + t1) if t0 is not None else samples[-1][0]
            inside = sorted(samples, key=lambda x: abs(x[0] - mid))[:2]
            note = 'timed region shorter than the polling period: nearest samples'
        reasons = sorted({n for x in inside for n in x[3]})
        return dict(sm_mhz=statistics.median([x[1] for x in inside]), sm_max_mhz=max(x[2] for x in inside),
                    reasons=reasons, samples=len(inside), note=note)


def hbm_peak():
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        if os.path.exists(peaks_path):
            return float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except (KeyError, ValueError, TypeError, OSError):
        pass
    return 6650.0, 'fallback (B200_PROFILING.md: 6.65 TB/s)'


def kernel_table(prof, prof_steps, kbytes, kflops, peak_gbs, fma_peaks, traffic_pf, nf):
    """Per-kernel records from the library's CUDA-event brackets: algorithmic HBM bytes -> GB/s and fraction of the measured
    copy bandwidth; for the compute-bound kernels also algorithmic flops against the MEASURED non-tensor FMA peak."""
    out = []
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        per = ms / prof_steps
        kb = float(kbytes.get(name, 0))
        gbs = kb / (per * 1e-3) / 1e9 if per > 0 else 0.0
        rec = dict(name=name, launches_per_step=cnt // prof_steps, ms_per_step=per, algorithmic_bytes_per_step=int(kb), gbs=gbs,
                   frac=gbs / peak_gbs,
                   dram_traffic_per_step=(int(traffic_pf[name] * nf) if name in traffic_pf else None))
        if name in kflops and per > 0:
            fl, which = kflops[name]
            tf = fl / (per * 1e-3) / 1e12
            rec['compute'] = {'flops_per_step': int(fl), 'achieved_tflops': tf, 'pipe': which, 'peak_tflops': fma_peaks[which],
                              'frac': tf / fma_peaks[which] if fma_peaks[which] > 0 else None,
                              'peak_source': 'measured on this device by mpb_measure_fma_peak (register-resident FMA chains)'}
        out.append(rec)
    return out


def time_steps(step, steps, warmup, barrier, torch):
    """W untimed + K timed calls of step(); device time of the timed region in ms (CUDA events on the current stream)."""
    for _ in range(warmup):
        step()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    barrier()
    return t0.elapsed_time(t1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='compressed', choices=['compressed', 'lossless'],
                    help="compressed = BASELINE config 2 (analysis_compressed 60/45/45 -> synthesis_from_compressed); "
                         "lossless = config 1 chain (analysis_lossless -> synthesis_from_lossless)")
    ap.add_argument('--utts', type=int, default=128, help='utterances per GPU per step (device-timed arm)')
    ap.add_argument('--dur', type=float, default=5.0, help='utterance length in seconds')
    ap.add_argument('--e2e-utts', type=int, default=0, help='utterances per GPU per step (host-API arm); 0 = auto')
    ap.add_argument('--cpu-utts', type=int, default=0, help='utterances per step of the CPU arm; 0 = 2 per host core')
    ap.add_argument('--cpu-dur', type=float, default=0.0, help='utterance length of the CPU arm sample; 0 = --dur')
    ap.add_argument('--feat-dtype', default='f32', choices=['f32', 'f64'], help='lossless feature storage in HBM')
    ap.add_argument('--analysis-compute', default='f64', choices=['f32', 'f64'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-rng-overlap', action='store_true', help='draw the noise on the main stream, before the synthesis half')
    ap.add_argument('--e2e-workers', type=int, default=0,
                    help='host threads (private contexts) of the end-to-end arm; 0 = 4')
    ap.add_argument('--no-extras', action='store_true', help='skip the sub-records (lossless chain, configs 3 / 4 / 5)')
    ap.add_argument('--stream-utts', type=int, default=2048, help='utterances per GPU of the streamed many-batch record (config 5)')
    ap.add_argument('--ola-target', type=int, default=32, help='frames per overlap-add run (device-timed arm)')
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    comp = a.workload == 'compressed'
    cores = os.cpu_count() or 1
    if a.cpu_utts == 0:
        a.cpu_utts = 2 * cores            # two tasks per worker: no core idles while the slowest utterance finishes
    if a.cpu_dur == 0.0:
        a.cpu_dur = a.dur                 # the same utterances as the GPU arm
    if a.e2e_utts == 0:
        a.e2e_utts = 128 if comp else 8
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    # one process per GPU on a shared host: every rank keeps to its own slice of the cores (its NumPy bookkeeping, staging
    # copies and the two worker threads of the end-to-end arm would otherwise migrate across all of them)
    lws = int(os.environ.get('LOCAL_WORLD_SIZE', world))
    if a.impl == 'ours' and lws > 1 and hasattr(os, 'sched_setaffinity'):
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // lws)
            os.sched_setaffinity(0, cpus[local_rank * per:(local_rank + 1) * per] or cpus)
        except OSError:
            pass
    if a.e2e_workers <= 0:
        # measured on 8 ranks x 4 cores: 4 workers with yielding waits (MPB_SYNC=block, the library's default under a
        # one-process-per-GPU launcher) 32 M frames/s, 2 workers 12 M -- the workers mostly sleep on the device
        a.e2e_workers = 4
    chain = ('analysis_compressed(mag=60, real=45, imag=45) -> synthesis_from_compressed(b_out_hpf=False)' if comp
             else 'analysis_lossless -> synthesis_from_lossless')
    workload = '%s chain: %s, %d x %.0f s synth48k-v1 utterances per GPU per step' % (a.workload, chain, a.utts, a.dur)

    if a.impl == 'reference':
        if rank != 0:
            return
        r = run_cpu_arm(a.cpu_utts, a.cpu_dur, a.steps, a.warmup, a.workload)
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': 'frames/s', 'n_gpus': a.gpus,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload, 'fs': FS, 'fft_len': FFT_LEN,
                       'sample': 'each step = %d of those utterances (%.0f s each) on %d host cores' % (r['n_utts'], r['dur_s'], r['cores'])},
            'cpu_baseline': {'value': r['value'], 'unit': 'frames/s', 'cores': r['cores'], 'kind': r['kind'],
                             'sample': r['sample']},
            'e2e': {'value': r['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}))
        return

    # ---- CPU baseline first (fork before CUDA is initialised), rank 0 at N=1 only ----
    cpu = None
    if world == 1 and a.gpus == 1 and not a.no_cpu_baseline:
        cpu = run_cpu_arm(a.cpu_utts, a.cpu_dur, 2, 1, a.workload)

    import torch
    import torch.distributed as dist
    os.environ.setdefault('MPB_DEVICE', str(local_rank))
    from magphase_b200 import _lib
    from magphase_b200 import magphase as mp
    from magphase_b200.device import CompressedPlan, LosslessPlan
    from magphase_b200.synth import synth_utterance

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    # ---- work list: rank 0 enumerates utterances, LPT-assigns them, NCCL broadcasts the assignment ----
    from magphase_b200.sharding import reduce_counters, scatter_work_list
    sizes = np.full(a.utts * world, int(round(a.dur * FS))) if rank == 0 else None
    my_ids, _ = scatter_work_list(sizes, device=dev)

    base = {}
    for u in my_ids[:DISTINCT]:
        base[int(u)] = synth_utterance(int(u), fs=FS, dur_s=a.dur)
    keys = list(base)
    utts = [base[keys[i % len(keys)]] for i in range(len(my_ids))]

    F32, F64 = _lib.MPB_F32, _lib.MPB_F64
    feat_dt = F64 if a.feat_dtype == 'f64' else F32
    ana_compute = F64 if a.analysis_compute == 'f64' else F32
    d_sig = torch.from_numpy(np.concatenate([u[0] for u in utts]).astype(np.float32)).to(dev)   # PCM16/32768: exact in f32
    geom = ([u[0].size for u in utts], [u[1] for u in utts], [u[2] for u in utts])

    def make_lossless(g, n_utts=None):
        gg = g if n_utts is None else tuple(x[:n_utts] for x in g)
        lp = LosslessPlan(*gg, FS, FFT_LEN, device=local_rank, ola_target_frames=a.ola_target)
        feats = lp.alloc_features(feat_dt)
        d_out = lp.alloc_output(F32)
        sig = d_sig if n_utts is None else d_sig[:int(sum(gg[0]))]

        def step(evs=None, sequential=False):
            if evs:
                evs[0].record()
            lp.analysis(sig, feats, compute=ana_compute)
            if evs:
                evs[1].record()
            lp.synthesis(feats, d_out, compute=F32)
            if evs:
                evs[2].record()
        return lp, step

    if comp:
        plan = CompressedPlan(*geom, FS, FFT_LEN, mag_dim=60, phase_dim=45, device=local_rank, ola_target_frames=a.ola_target)

        def step(evs=None, sequential=False):
            if evs:
                evs[0].record()
            if sequential or a.no_rng_overlap:
                plan.analysis(d_sig, compute=ana_compute)
                if evs:
                    evs[1].record()
                plan.synthesis()       # draws the aperiodic noise (src/magphase.py:883) on the main stream first
            else:
                # the MT19937 draw runs on a side stream next to the analysis half; both join before the synthesis half,
                # inside the timed region
                plan.chain(d_sig, compute=ana_compute, mid_event=evs[1] if evs else None)
            if evs:
                evs[2].record()
        bytes_ana, bytes_syn = plan.analysis_bytes(), plan.synthesis_bytes()
    else:
        plan, step = make_lossless(geom)
        bytes_ana, bytes_syn = plan.analysis_bytes(F32, feat_dt), plan.synthesis_bytes(feat_dt, F32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None      # polls from before the warm-up (nvidia-smi start-up)
    for _ in range(a.warmup):
        step()
    barrier()
    launches0 = _lib.launch_count(local_rank)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    t_start.record()
    for k in range(a.steps):
        step(ev[k])
    t_end.record()
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if sampler else None
    launches = _lib.launch_count(local_rank) - launches0
    total_ms = t_start.elapsed_time(t_end)
    ana_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / a.steps
    syn_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / a.steps

    # ---- per-kernel device times: extra steps with the library's CUDA-event brackets switched on ----
    prof_steps = 3
    _lib.profile_begin(local_rank)
    for _ in range(prof_steps):
        step(sequential=True)          # the noise draw on the main stream: clean per-kernel times
    prof = _lib.profile_end(local_rank)

    # ---- e2e through the public host API (NumPy in / NumPy out; H2D + D2H inside the timed region) ----
    e_utts = utts[:max(1, min(a.e2e_utts, len(utts)))]
    e_sig, e_pm, e_voi = [u[0] for u in e_utts], [u[1] for u in e_utts], [u[2] for u in e_utts]
    e2e64 = None
    if comp:
        # the call a user of the reference's file formats makes: PCM16 samples as the wav file holds them (sf.read's 1/32768 is
        # applied on the device), float32 features (the reference's feature files) and a float32 waveform (the wav writer
        # quantises to PCM16 anyway) -- element types on request; the float64-in / float64-out default of the drop-in
        # signatures is timed as well (e2e_float64_api)
        e_pcm = [np.round(x * 32768.0).astype(np.int16) for x in e_sig]

        from magphase_b200.batch import run_chain_stream
        e_batch = list(zip(e_pcm, e_pm, e_voi))

        def e2e_step(narrow=True, keep=False, n=1):
            if narrow:
                # n steps = n batches streamed through `--e2e-workers` host threads, each with a private context on this GPU
                # (magphase_b200.batch.run_chain_stream): the NumPy bookkeeping and PCIe copies of one step overlap the kernels
                # of its neighbours -- what the reference gets from one forked process per utterance.  Every step's inputs are
                # copied to the device and its results back to the host inside the timed region.
                r = run_chain_stream([e_batch] * n, FS, fft_len=FFT_LEN, mag_dim=60, phase_dim=45, b_out_hpf=False,
                                     n_workers=a.e2e_workers, out_dtype=np.float32, keep_outputs=keep)
                if not keep:
                    return r['frames'] // n, None, None
                return r['frames'] // n, r['outputs'][0][0], r['outputs'][0][1]
            else:
                outs = mp.analysis_compressed_batch(e_sig, FS, e_pm, e_voi, fft_len=FFT_LEN, mag_dim=60, phase_dim=45)
                ys = mp.synthesis_from_compressed_batch([o[:4] for o in outs], FS, b_out_hpf=False)
            return sum(o[4].size for o in outs), outs, ys
        api = ('magphase_b200.batch.run_chain_stream: per step analysis_compressed_batch(int16 PCM in, out_dtype=float32) -> '
               'synthesis_from_compressed_batch(float32 features in, out_dtype=float32) over %d utterances; the steps are taken '
               'round-robin by %d host threads with private contexts; NumPy-stream noise drawn on the device' % (len(e_batch), a.e2e_workers))
    else:
        def e2e_step(narrow=True, keep=False):
            outs = mp.analysis_lossless_batch(e_sig, FS, e_pm, e_voi, fft_len=FFT_LEN)
            ys = mp.synthesis_from_lossless_batch([o[:4] for o in outs], FS)
            return sum(o[5].size for o in outs), outs, ys
        api = 'analysis_lossless_batch -> synthesis_from_lossless_batch (float64 NumPy in/out)'
    e_steps = max(6, a.steps) if comp else max(2, min(a.steps, 5))

    pool_stats = {}

    def time_e2e(narrow):
        r = e2e_step(narrow, keep=True)       # untimed: the results themselves, for the byte counts
        if narrow and comp:
            e2e_step(True, n=e_steps)         # untimed pass of the same stream: every worker's context and pinned blocks exist
            barrier()
            pool0 = dict(_lib.pinned.stats)
            c0 = time.process_time()
            t = time.perf_counter()
            e2e_step(True, n=e_steps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            pool_stats.update({k: round(v - pool0[k], 4) for k, v in _lib.pinned.stats.items()})
            pool_stats['host_cpu_ms_per_step'] = round(1e3 * (time.process_time() - c0) / e_steps, 2)   # all threads of this rank
            return (dt, ) + r
        e2e_step(narrow)
        barrier()
        t = time.perf_counter()
        for _ in range(e_steps):
            e2e_step(narrow)
        torch.cuda.synchronize()
        return (time.perf_counter() - t, ) + r
    e_secs, e_frames, outs, ys = time_e2e(True)
    n_smp = sum(s.size for s in e_sig)
    if comp:
        feat_bytes = 4 * sum(o[0].size + o[1].size + o[2].size for o in outs)
        # PCM16 samples 2 B; analysis descriptors 17 B/frame; float32 features each way; synthesis descriptors 35 B/frame; the
        # noise is drawn on the device (only the 2.5 KB MT19937 state travels); float32 waveform back
        h2d = 2 * n_smp + 17 * e_frames + feat_bytes + 35 * e_frames + 2500
        d2h = feat_bytes + 4 * sum(y.size for y in ys) + 2500
        del outs, ys
        s64, f64_frames, outs, ys = time_e2e(False)
        e2e64 = {'seconds_per_step': s64 / e_steps, 'frames': f64_frames,
                 'h2d_bytes_per_step': int(4 * n_smp + 52 * f64_frames + 8 * sum(o[0].size + o[1].size + o[2].size for o in outs) + 2500),
                 'd2h_bytes_per_step': int(8 * sum(o[0].size + o[1].size + o[2].size for o in outs) + 8 * sum(y.size for y in ys) + 2500),
                 'api': 'the same two calls with their drop-in defaults: float64 NumPy in / out'}
    else:
        feat_bytes = 3 * 8 * sum(o[0].size for o in outs)
        h2d = 4 * n_smp + 16 * e_frames + feat_bytes + 4 * e_frames
        d2h = feat_bytes + 8 * sum(y.size for y in ys)
    del outs, ys

    # ---- sub-records: the other chain and the other BASELINE configs, each its own short timed region ----
    extras = {}
    H, nf = FFT_LEN // 2 + 1, plan.nfrm
    if not a.no_extras and comp:
        try:
            extras = run_extras(a, torch, _lib, mp, CompressedPlan, make_lossless, geom, d_sig, utts, barrier, local_rank, world, rank)
        except Exception as e:                      # explanatory records must never cost the headline line
            extras = {'error': '%s: %s' % (type(e).__name__, e)}

    # ---- measured roofline denominators ----
    peak, peak_src = hbm_peak()
    fma_peaks = {'fp32': _lib.measure_fma_peak(F32, local_rank), 'fp64': _lib.measure_fma_peak(F64, local_rank)}

    # ---- reduce over ranks: max time, summed frames ----
    s64 = e2e64['seconds_per_step'] if e2e64 else 0.0
    stats, counts = reduce_counters([total_ms, e_secs, ana_ms, syn_ms, s64], [plan.nfrm, e_frames, pool_stats.get('allocs', 0)],
                                    device=dev)
    total_ms, e_secs, ana_ms, syn_ms, s64 = [float(x) for x in stats]
    frames_all, e_frames_all, pool_allocs_all = [float(x) for x in counts]
    pool_allocs_all = int(pool_allocs_all)
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    # algorithmic HBM bytes per step of each kernel = what its own contract must move, and the
    # algorithmic flops of the compute-bound ones (a real N-point FFT = 2.5 N log2 N flops)
    nv = getattr(plan, 'n_voiced', nf)       # the phase rows / phase-stream products only exist for voiced frames
    vf = nv / max(nf, 1)
    LP = (H + 3) & ~3
    fft_flops = 2.5 * FFT_LEN * np.log2(FFT_LEN)
    n_noise, n_out = getattr(plan, 'n_noise', 0), getattr(plan, 'n_out', 0)
    kbytes = {
        'k_analysis<logp>': plan.n_sig * 4 + nf * 17 + (nf + 2 * nv) * LP * 4,
        'k_analysis': plan.n_sig * 4 + nf * 16 + 3 * nf * H * (8 if (not comp and feat_dt == F64) else 4),
        'k_synthesis_lossless': 3 * nf * H * (8 if feat_dt == F64 else 4) + nf * 4 + n_out * 4,
        'k_mel_warp_tc': (nf + 2 * nv) * LP * 4 + (nf + 2 * nv) * 64 * 4,
        'k_mel_cos': (nf + 2 * nv) * 64 * 4 + nf * 150 * 4,
        'k_mel_unwarp_tc': nf * 150 * 4 + nf * H * 4 + 2 * nv * 512 * 4,
        'k_mel_gemm': (nf + 2 * nv) * H * 4 + nf * 150 * 4, 'k_mel_finish': nf * 150 * 4,
        'k_mel_unwarp': nf * 150 * 4 + nf * H * 4 + 2 * nv * 512 * 4,
        'k_analysis<noise_logsq>': n_noise * 4 + nf * 25 + nf * (H + 1) * 8,      # + stored noise spectra
        'k_mt19937_stream+k_mt_to_uniform': n_noise * 4,
        'k_noise_gain': nf * 9,
        'k_synthesis_compressed': nf * (H + 1) * 8 + nf * H * 4 + 2 * nv * 512 * 4 + nf * 45 + n_out * 4,
    }
    kflops = {
        'k_analysis<logp>': (nf * fft_flops, 'fp64'), 'k_analysis': (nf * fft_flops, 'fp64' if ana_compute == F64 else 'fp32'),
        'k_analysis<noise_logsq>': (nf * fft_flops, 'fp32'), 'k_synthesis_compressed': (nf * fft_flops, 'fp32'),
        'k_synthesis_lossless': (nf * fft_flops, 'fp32'),
        'k_mel_cos': (2.0 * (nf * 60 * 60 + 2 * nv * 58 * 45), 'fp64'),
    }
    traffic_pf, traffic_src = {}, None
    for rdir in ('r2', 'r1b'):
        tpath = os.path.join(ROOT, 'profiles', rdir, 'traffic_per_frame.json')
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic_pf = tj['dram_bytes_per_frame']
            traffic_src = 'profiles/%s/traffic_per_frame.json (ncu --set full capture, scaled to this launch size)' % rdir
            break
    kernels = kernel_table(prof, prof_steps, kbytes, kflops, peak, fma_peaks, traffic_pf, nf)
    dom = kernels[0]
    ms_per_step = total_ms / a.steps
    value = frames_all / (ms_per_step * 1e-3)
    chain_bytes = bytes_ana + bytes_syn
    known = [k['dram_traffic_per_step'] for k in kernels if k['dram_traffic_per_step']]
    chain_traffic = int(sum(known)) if known else None
    e2e_value = e_frames_all / (e_secs / e_steps)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 analysis butterflies; f32 elsewhere; 3xTF32 tensor-core tile products; f32 storage',
        'data': 'synthetic',
        'config': {'workload': workload, 'fs': FS, 'fft_len': FFT_LEN, 'frames_per_gpu_per_step': plan.nfrm,
                   'voiced_fraction': round(vf, 3),
                   'mean_shift_samples': round(plan.mean_shift, 1),
                   'rng': 'the np.random.uniform draw of the aperiodic noise runs on the device inside every timed step',
                   'l2_note': 'per-step intermediates %.1f GB in HBM >> 126 MB L2' % (3 * nf * H * 4 / 1e9),
                   'parallelism': 'utterance-sharded x%d' % world},
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'utts_per_gpu_per_step': len(e_utts), 'api': api,
                'frac_of_value': e2e_value / value if value > 0 else None, 'steps': e_steps, 'workers': a.e2e_workers,
                'host_cpu_ms_per_step': pool_stats.pop('host_cpu_ms_per_step', None),
                'page_lock_calls_during_timed_steps_all_ranks': pool_allocs_all,
                'pinned_pool_during_timed_steps': pool_stats or None},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': dom['name'], 'achieved': dom['gbs'], 'peak': peak, 'unit': 'GB/s',
                     'frac': dom['frac'],
                     'traffic': (int(dom['dram_traffic_per_step'] / max(dom['launches_per_step'], 1))
                                 if dom['dram_traffic_per_step'] else None),
                     'traffic_source': traffic_src, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': int(dom['algorithmic_bytes_per_step'] / max(dom['launches_per_step'], 1)),
                     'ms_per_launch': dom['ms_per_step'] / max(dom['launches_per_step'], 1),
                     'compute': dom.get('compute'),
                     'limiter': 'ncu (profiles/r2/c_k_analysis_logp.txt): FP64 pipe 44 % active, issue slots 56 %, warps active '
                                '28 % (96 registers x 128 threads, 5 CTAs/SM) -- latency / FP64-pipe bound, NOT HBM-bound; '
                                '"bound" names the nearer of the two rooflines the line format knows',
                     'note': 'the dominant kernels are the three FFT kernels: their contract is HBM bytes (this fraction), their '
                             'practical limit is instruction issue / the FP64 pipe (see compute.frac against the measured FMA '
                             'peak and profiles/); the tensor-core tile products are HBM-bound (their own frac in kernels[])'},
        'chain': {'algorithmic_bytes_per_step': int(chain_bytes), 'gbs': chain_bytes / (ms_per_step * 1e-3) / 1e9,
                  'frac_of_hbm': chain_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                  'dram_traffic_per_step': chain_traffic,
                  'traffic_over_algorithmic': (chain_traffic / chain_bytes if chain_traffic else None),
                  'note': 'SURVEY 8(d) bytes of the reference API contract (samples + descriptors + low-dimensional features); '
                          'the intermediates between the kernels (log periodograms, noise spectra, un-warped rows) are extra traffic'},
        'fma_peaks_tflops': fma_peaks,
        'halves': {'analysis_ms': ana_ms, 'synthesis_ms': syn_ms},
        'kernels': kernels,
        'clocks': clocks,
    }
    if e2e64:
        line['e2e_float64_api'] = {'value': e_frames_all / s64 if s64 > 0 else None, 'unit': 'frames/s',
                                   'h2d_bytes_per_step': e2e64['h2d_bytes_per_step'], 'd2h_bytes_per_step': e2e64['d2h_bytes_per_step'],
                                   'api': e2e64['api']}
    line.update(extras)
    if cpu is not None:
        line['cpu_baseline'] = {'value': cpu['value'], 'unit': 'frames/s', 'cores': cpu['cores'], 'kind': cpu['kind'],
                                'sample': cpu['sample']}
    print(json.dumps(line))


def run_extras(a, torch, _lib, mp, CompressedPlan, make_lossless, geom, d_sig, utts, barrier, local_rank, world, rank):
    """Sub-records of the default line (each a short timed region of its own, per GPU; under torchrun every rank runs them and
    rank 0 reports its own numbers):
      lossless      BASELINE config 1 chain, device-resident (the chain SURVEY 8(d)'s 125 M frames/s HBM roofline is written for)
      extract_tts   config 3: analysis_for_acoustic_modelling's feature extraction (mag 60, phase_dim 10, alpha_phase=False quirk)
      generate_16k  config 4: 16 kHz constant-rate features -> post_filter -> synthesis_from_compressed with the output high-pass
      stream        config 5 per GPU: many batches of UNEQUAL utterances streamed through the host API (LPT order)"""
    ex = {}
    steps = max(3, min(a.steps, 5))
    peak, _ = hbm_peak()
    F32 = _lib.MPB_F32
    # ---- lossless chain ----
    n_l = min(len(utts), 64)
    lp, lstep = make_lossless(geom, n_l)
    ms = time_steps(lstep, steps, 3, barrier, torch) / steps
    _lib.profile_begin(local_rank)
    for _ in range(2):
        lstep()
    prof = _lib.profile_end(local_rank)
    lb = lp.analysis_bytes(F32, F32) + lp.synthesis_bytes(F32, F32)
    # per kernel: its half of the chain's algorithmic bytes against the measured copy bandwidth, and one N-point real FFT per
    # frame (2.5 N log2 N flops) against the measured FMA peak of the pipe it runs on (float64 analysis, float32 synthesis)
    fpk = {'fp32': _lib.measure_fma_peak(_lib.MPB_F32, local_rank), 'fp64': _lib.measure_fma_peak(_lib.MPB_F64, local_rank)}
    fl = lp.nfrm * 2.5 * FFT_LEN * np.log2(FFT_LEN)
    kb = {'k_analysis': (lp.analysis_bytes(F32, F32), 'fp64'), 'k_synthesis_lossless': (lp.synthesis_bytes(F32, F32), 'fp32')}
    lk = []
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        per = v[1] / 2
        rec = dict(name=k, ms_per_step=per)
        if k in kb and per > 0:
            rec.update(algorithmic_bytes_per_step=int(kb[k][0]), gbs=kb[k][0] / (per * 1e-3) / 1e9,
                       frac=kb[k][0] / (per * 1e-3) / 1e9 / peak,
                       compute={'pipe': kb[k][1], 'achieved_tflops': fl / (per * 1e-3) / 1e12,
                                'frac': fl / (per * 1e-3) / 1e12 / fpk[kb[k][1]]})
        lk.append(rec)
    ex['lossless'] = {'workload': 'analysis_lossless -> synthesis_from_lossless, %d x %.0f s utterances, float32 features in HBM' % (n_l, a.dur),
                      'value': lp.nfrm / (ms * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms,
                      'chain_algorithmic_bytes_per_step': int(lb), 'chain_gbs': lb / (ms * 1e-3) / 1e9,
                      'frac_of_hbm': lb / (ms * 1e-3) / 1e9 / peak,
                      'kernels': lk}
    del lp, lstep
    torch.cuda.empty_cache()
    # the same chain through the host API (float64 NumPy in / out like the reference: 3 x 2049 float64 features per frame
    # cross PCIe in each direction, 98 KB per frame -- this path is a PCIe measurement)
    n_le = min(len(utts), 8)
    le = ([u[0] for u in utts[:n_le]], [u[1] for u in utts[:n_le]], [u[2] for u in utts[:n_le]])

    le_pcm = [np.round(x * 32768.0).astype(np.int16) for x in le[0]]

    def l_e2e(narrow=False):
        if narrow:      # the reference's file formats: PCM16 in, float32 feature matrices, float32 waveform
            o = mp.analysis_lossless_batch(le_pcm, FS, le[1], le[2], fft_len=FFT_LEN, out_dtype=np.float32)
            mp.synthesis_from_lossless_batch([x[:4] for x in o], FS, out_dtype=np.float32)
        else:
            o = mp.analysis_lossless_batch(le[0], FS, le[1], le[2], fft_len=FFT_LEN)
            mp.synthesis_from_lossless_batch([x[:4] for x in o], FS)
        return sum(x[5].size for x in o)

    def l_time(narrow):
        l_e2e(narrow); n = l_e2e(narrow)
        barrier()
        t = time.perf_counter()
        for _ in range(3):
            l_e2e(narrow)
        torch.cuda.synchronize()
        return n, (time.perf_counter() - t) / 3
    l_frames, l_s = l_time(False)
    _, l_sn = l_time(True)
    ex['lossless']['e2e'] = {'value': l_frames / l_s, 'unit': 'frames/s', 'utts': n_le,
                             'h2d_bytes_per_step': int(sum(x.size for x in le[0]) * 4 + l_frames * (3 * (FFT_LEN // 2 + 1) * 8 + 16)),
                             'd2h_bytes_per_step': int(l_frames * 3 * (FFT_LEN // 2 + 1) * 8 + sum(x.size for x in le[0]) * 8),
                             'api': 'analysis_lossless_batch -> synthesis_from_lossless_batch (float64 NumPy in / out)',
                             'narrow': {'value': l_frames / l_sn, 'unit': 'frames/s',
                                        'api': 'the same two calls with PCM16 signals in, float32 feature matrices and float32 waveform '
                                               '(out_dtype=np.float32): 49 KB per frame and direction'}}
    # ---- config 3: feature extraction for TTS (analysis only) ----
    p3 = CompressedPlan(*geom, FS, FFT_LEN, mag_dim=60, phase_dim=10, device=local_rank, alpha_phase=0.0)
    ms = time_steps(lambda: p3.analysis(d_sig), steps, 3, barrier, torch) / steps
    e_n = min(len(utts), 128)
    e_sig, e_pm, e_voi = [u[0] for u in utts[:e_n]], [u[1] for u in utts[:e_n]], [u[2] for u in utts[:e_n]]
    f3 = lambda: mp.analysis_compressed_batch(e_sig, FS, e_pm, e_voi, fft_len=FFT_LEN, mag_dim=60, phase_dim=10, alpha_phase=0.0)
    f3(); outs = f3()
    barrier()
    t = time.perf_counter()
    for _ in range(3):
        f3()
    torch.cuda.synchronize()
    e_s = (time.perf_counter() - t) / 3
    ex['extract_tts'] = {'workload': 'config 3: analysis_compressed(mag_dim=60, phase_dim=10, alpha_phase=0.0 as analysis_for_acoustic_modelling '
                                     'passes it), %d x %.0f s utterances' % (len(utts), a.dur),
                         'value': p3.nfrm / (ms * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms,
                         'e2e': {'value': sum(o[4].size for o in outs) / e_s, 'unit': 'frames/s', 'utts': e_n,
                                 'api': 'analysis_compressed_batch (NumPy in / out)'}}
    del p3, outs
    torch.cuda.empty_cache()
    # ---- config 4: 16 kHz constant-rate generation with post-filter and output high-pass (host API: NumPy in / out) ----
    from magphase_b200.synth import synth_utterance
    fs16 = 16000
    base16 = [synth_utterance(3000 + u, fs=fs16, dur_s=a.dur) for u in range(8)]
    n16 = 128
    u16 = [base16[i % len(base16)] for i in range(n16)]
    feats16 = mp.analysis_compressed_batch([u[0] for u in u16], fs16, [u[1] for u in u16], [u[2] for u in u16], mag_dim=60,
                                           phase_dim=45, b_const_rate=True)
    feats16 = [tuple(np.array(a) for a in f[:4]) for f in feats16]     # one pageable array per utterance and stream, as read from files
    n_rows = sum(f[0].shape[0] for f in feats16)

    def g16():
        # as magphase_b200.batch.run_waveform_generation does it: the post-filter works frame by frame, one call over the
        # stacked rows of the batch; the feature rows stay one block per utterance (as read from one file per utterance)
        rows = mp.post_filter(np.concatenate([f[0] for f in feats16], axis=0), fs16)
        off = np.concatenate(([0], np.cumsum([f[0].shape[0] for f in feats16])))
        pf = [(np.ascontiguousarray(rows[off[i]:off[i + 1]]), f[1], f[2], f[3]) for i, f in enumerate(feats16)]
        return mp.synthesis_from_compressed_batch(pf, fs16, b_const_rate=True, b_out_hpf=True)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        g16(); ys = g16()
        barrier()
        t = time.perf_counter()
        for _ in range(3):
            g16()
        torch.cuda.synchronize()
        g_s = (time.perf_counter() - t) / 3
    n_smp = sum(y.size for y in ys)
    ex['generate_16k'] = {'workload': 'config 4: %d utterances of %.0f s at 16 kHz, constant-rate (5 ms) features -> post_filter -> '
                                      'synthesis_from_compressed(b_const_rate=True, b_out_hpf=True), host API end to end' % (n16, a.dur),
                          'value': n_rows / g_s, 'unit': 'constant-rate frames/s', 'seconds_per_step': g_s,
                          'audio_seconds_per_second': n_smp / fs16 / g_s}
    del feats16, ys
    # ---- config 5 (per GPU): a stream of many batches of unequal utterances through the host API ----
    durs = [2.0, 3.0, 4.0, 5.0, 6.0, 8.0]
    pool = [synth_utterance(5000 + i, fs=FS, dur_s=d) for i, d in enumerate(durs)]
    n_total = max(128, int(a.stream_utts))
    order = np.random.Generator(np.random.PCG64(7)).integers(0, len(pool), n_total)
    # longest first (LPT) inside this GPU's shard, then batches of 128 utterances
    order = sorted(order.tolist(), key=lambda i: -pool[i][0].size)
    batches = [order[i:i + 128] for i in range(0, n_total, 128)]
    from magphase_b200.batch import run_chain_stream
    pcm_pool = [(np.round(u[0] * 32768.0).astype(np.int16), u[1], u[2]) for u in pool]
    bl = [[pcm_pool[i] for i in b] for b in batches]
    run_chain_stream(bl, FS, fft_len=FFT_LEN, n_workers=a.e2e_workers)   # first pass: device scratch grows, result buffers get page-locked and pooled
    barrier()
    r = run_chain_stream(bl, FS, fft_len=FFT_LEN, n_workers=a.e2e_workers)
    torch.cuda.synchronize()
    frames, s_s = r['frames'], r['seconds']
    ex['stream'] = {'workload': 'config 5 per GPU: %d utterances of 2-8 s (LPT order), %d batches of 128 through '
                                'magphase_b200.batch.run_chain_stream (PCM16 in, float32 out, %d host threads)' % (n_total, len(batches), a.e2e_workers),
                    'value': frames / s_s, 'unit': 'frames/s', 'frames': int(frames), 'seconds': s_s,
                    'note': 'steady state (second pass over the list; the first pass page-locks the pooled result buffers). BASELINE config 5 = '
                            '12,500 such utterances per GPU on 8 GPUs: run with --stream-utts 12500'}
    return ex


if __name__ == '__main__':
    main()

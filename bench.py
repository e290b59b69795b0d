#!/usr/bin/env python
"""Benchmark of the MagPhase analysis+synthesis hot path (BASELINE.json metric: frames/sec at 48 kHz, FFT 4096).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One *step* = one pass of analysis -> synthesis over one batch of synthetic 48 kHz utterances ("synth48k-v1",
magphase_b200/synth.py).  `value` is device-timed with signals, descriptors and features resident in HBM;
`e2e` goes through the public host API (NumPy in, NumPy out, H2D/D2H inside the timed region).
Under torchrun every rank runs its own shard of utterances (weak scaling, no data-path collective: the only
NCCL traffic is the scattered work list and the final counters).
"""
import argparse
import json
import multiprocessing
import os
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

def fp32_fma_view(kernels, nf, nv, H, sm_mhz, num_sms, mag_dim=60, nmel=58, phase_dim=45, hb=512):
    """the two tile products are FP32-FMA bound, not HBM bound, so next to their HBM fraction report the
    algorithmic FLOP/s against the non-tensor FP32 peak.  Peak = SMs x 128 FMA lanes x 2 x the SM clock sampled during the
    timed region (derived from the clock, not a measured GEMM).  FMAs per step: warp product (nf x mag_dim + 2 nv x nmel)
    x (H - 1) bins (src/libaudio.py:575-601 as a matrix); un-warp product nf x mag_dim x H +
    2 nv x phase_dim x hb (src/libaudio.py:667-684).  Adds an 'fp32_fma' entry to those kernels in place."""
    peak = num_sms * 128 * 2 * sm_mhz * 1e6 / 1e12
    fmas = {'k_mel_gemm': (nf * mag_dim + 2 * nv * nmel) * (H - 1),
            'k_mel_unwarp': nf * mag_dim * H + 2 * nv * phase_dim * hb}
    for k in kernels:
        n = fmas.get(k['name'])
        if n and k['ms_per_step'] > 0 and peak > 0:
            tf = 2.0 * n / (k['ms_per_step'] * 1e-3) / 1e12
            k['fp32_fma'] = {'flops_per_step': int(2 * n), 'achieved_tflops': tf, 'peak_tflops': peak, 'frac': tf / peak,
                             'peak_source': '%d SMs x 128 FMA/clk x 2 x %.0f MHz (sampled SM clock; derived, not measured)'
                                            % (num_sms, sm_mhz)}
    return kernels


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
    for k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
        os.environ[k] = '1'
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
    return dict(kind=kind, cores=cores, frames_per_step=frames // max(steps, 1), value=frames / secs,
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
    ap.add_argument('--cpu-utts', type=int, default=0, help='utterances per step of the CPU arm; 0 = auto')
    ap.add_argument('--cpu-dur', type=float, default=0.0, help='utterance length of the CPU arm sample; 0 = auto')
    ap.add_argument('--feat-dtype', default='f32', choices=['f32', 'f64'], help='lossless feature storage in HBM')
    ap.add_argument('--analysis-compute', default='f64', choices=['f32', 'f64'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ola-target', type=int, default=32, help='frames per overlap-add run (device-timed arm)')
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    comp = a.workload == 'compressed'
    if a.cpu_utts == 0:
        a.cpu_utts = 16 if comp else 32
    if a.cpu_dur == 0.0:
        a.cpu_dur = 2.0 if comp else a.dur     # the reference's compressed synthesis runs at ~70 frames/s/core
    if a.e2e_utts == 0:
        a.e2e_utts = 128 if comp else 8
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
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
            'config': {'workload': workload, 'fs': FS, 'fft_len': FFT_LEN},
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
    if comp:
        plan = CompressedPlan(*geom, FS, FFT_LEN, mag_dim=60, phase_dim=45, device=local_rank, ola_target_frames=a.ola_target)

        def step(evs=None):
            if evs:
                evs[0].record()
            plan.analysis(d_sig, compute=ana_compute)
            if evs:
                evs[1].record()
            plan.synthesis()
            if evs:
                evs[2].record()
        bytes_ana, bytes_syn = plan.analysis_bytes(), plan.synthesis_bytes()
    else:
        plan = LosslessPlan(*geom, FS, FFT_LEN, device=local_rank, ola_target_frames=a.ola_target)
        feats = plan.alloc_features(feat_dt)
        d_out = plan.alloc_output(F32)

        def step(evs=None):
            if evs:
                evs[0].record()
            plan.analysis(d_sig, feats, compute=ana_compute)
            if evs:
                evs[1].record()
            plan.synthesis(feats, d_out, compute=F32)
            if evs:
                evs[2].record()
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
        step()
    prof = _lib.profile_end(local_rank)

    # ---- e2e through the public host API (NumPy in / NumPy out; H2D + D2H inside the timed region) ----
    e_utts = utts[:max(1, min(a.e2e_utts, len(utts)))]
    e_sig, e_pm, e_voi = [u[0] for u in e_utts], [u[1] for u in e_utts], [u[2] for u in e_utts]
    if comp:
        def e2e_step():
            outs = mp.analysis_compressed_batch(e_sig, FS, e_pm, e_voi, fft_len=FFT_LEN, mag_dim=60, phase_dim=45)
            ys = mp.synthesis_from_compressed_batch([o[:4] for o in outs], FS, b_out_hpf=False)
            return sum(o[4].size for o in outs), outs, ys
        api = 'analysis_compressed_batch -> synthesis_from_compressed_batch (float64 NumPy in/out, np.random noise)'
    else:
        def e2e_step():
            outs = mp.analysis_lossless_batch(e_sig, FS, e_pm, e_voi, fft_len=FFT_LEN)
            ys = mp.synthesis_from_lossless_batch([o[:4] for o in outs], FS)
            return sum(o[5].size for o in outs), outs, ys
        api = 'analysis_lossless_batch -> synthesis_from_lossless_batch (float64 NumPy in/out)'
    for _ in range(2):
        e_frames, outs, ys = e2e_step()
    e_steps = max(2, min(a.steps, 5))
    barrier()
    t = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e_secs = time.perf_counter() - t
    n_smp = sum(s.size for s in e_sig)
    if comp:
        feat_bytes = 8 * sum(o[0].size + o[1].size + o[2].size for o in outs)
        n_noise = sum(y.size for y in ys)      # ~ one noise sample per output sample
        # signals cross PCIe as float32 (PCM-exact samples are narrowed on the host, mpb_stage.cu); analysis descriptors
        # 17 B/frame; features 8 B/value each way; synthesis descriptors 35 B/frame; the noise is drawn on the device
        # (only the 2.5 KB MT19937 state travels)
        h2d = 4 * n_smp + 17 * e_frames + feat_bytes + 35 * e_frames + 2500
        d2h = feat_bytes + 8 * sum(y.size for y in ys) + 2500
    else:
        feat_bytes = 3 * 8 * sum(o[0].size for o in outs)
        h2d = 4 * n_smp + 16 * e_frames + feat_bytes + 4 * e_frames
        d2h = feat_bytes + 8 * sum(y.size for y in ys)

    # ---- reduce over ranks: max time, summed frames ----
    stats, counts = reduce_counters([total_ms, e_secs, ana_ms, syn_ms], [plan.nfrm, e_frames], device=dev)
    total_ms, e_secs, ana_ms, syn_ms = [float(x) for x in stats]
    frames_all, e_frames_all = [float(x) for x in counts]
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md: 6.65 TB/s)'
    try:
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except (KeyError, ValueError, TypeError, OSError):
        pass
    # algorithmic HBM bytes per launch of each kernel = what its own contract must move
    H, nf = FFT_LEN // 2 + 1, plan.nfrm
    nv = getattr(plan, 'n_voiced', nf)       # the phase rows / phase-stream products only exist for voiced frames
    vf = nv / max(nf, 1)
    kbytes = {
        'k_analysis<logp>': plan.n_sig * 4 + nf * 17 + (nf + 2 * nv) * H * 4,
        'k_analysis': plan.n_sig * 4 + nf * 16 + 3 * nf * H * (8 if (not comp and feat_dt == F64) else 4),
        'k_synthesis_lossless': 3 * nf * H * (8 if feat_dt == F64 else 4) + nf * 4 + getattr(plan, 'n_out', 0) * 4,
        'k_mel_gemm': (nf + 2 * nv) * H * 4 + nf * 150 * 4,
        'k_mel_finish': nf * 150 * 4,
        'k_mel_unwarp': nf * 150 * 4 + nf * H * 4 + vf * nf * 2 * 512 * 4,
        'k_analysis<noise_logsq>': getattr(plan, 'n_noise', 0) * 4 + nf * 25 + nf * (H + 1) * 8,      # + stored noise spectra
        'k_noise_gain': nf * 9,
        'k_synthesis_compressed': nf * (H + 1) * 8 + nf * H * 4 + vf * nf * 2 * 512 * 4 + nf * 45
                                  + getattr(plan, 'n_out', 0) * 4,
    }
    traffic_pf, traffic_src = {}, None
    tpath = os.path.join(ROOT, 'profiles', 'r1b', 'traffic_per_frame.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic_pf, traffic_src = tj['dram_bytes_per_frame'], 'profiles/r1b/traffic_per_frame.json (ncu --set full capture, scaled to this launch size)'
    kernels = []
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        per_step_ms = ms / prof_steps
        kb = float(kbytes.get(name, 0))
        gbs = kb / (per_step_ms * 1e-3) / 1e9 if per_step_ms > 0 else 0.0
        kernels.append(dict(name=name, launches_per_step=cnt // prof_steps, ms_per_step=per_step_ms,
                            algorithmic_bytes_per_step=int(kb), gbs=gbs, frac=gbs / peak,
                            dram_traffic_per_step=(int(traffic_pf[name] * nf) if name in traffic_pf else None)))
    dom = kernels[0]
    try:          # explanatory only: never let it cost the line
        if comp:
            fp32_fma_view(kernels, nf, nv, H, float(clocks.get('sm_mhz') or clocks.get('sm_max_mhz') or 0.0),
                          torch.cuda.get_device_properties(dev).multi_processor_count)
    except Exception:
        pass
    ms_per_step = total_ms / a.steps
    value = frames_all / (ms_per_step * 1e-3)
    chain_bytes = bytes_ana + bytes_syn
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 analysis butterflies, f32 elsewhere; f32 storage',
        'data': 'synthetic',
        'config': {'workload': workload, 'fs': FS, 'fft_len': FFT_LEN, 'frames_per_gpu_per_step': plan.nfrm,
                   'voiced_fraction': round(vf, 3),
                   'mean_shift_samples': round(plan.mean_shift, 1),
                   'l2_note': 'per-step intermediates %.1f GB in HBM >> 126 MB L2' % (3 * nf * H * 4 / 1e9),
                   'parallelism': 'utterance-sharded x%d' % world},
        'e2e': {'value': e_frames_all / (e_secs / e_steps), 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'utts_per_gpu_per_step': len(e_utts), 'api': api},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': dom['name'], 'achieved': dom['gbs'], 'peak': peak, 'unit': 'GB/s',
                     'frac': dom['frac'],
                     'traffic': (int(dom['dram_traffic_per_step'] / max(dom['launches_per_step'], 1))
                                 if dom['dram_traffic_per_step'] else None),
                     'traffic_source': traffic_src, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': int(dom['algorithmic_bytes_per_step'] / max(dom['launches_per_step'], 1)),
                     'ms_per_launch': dom['ms_per_step'] / max(dom['launches_per_step'], 1),
                     'note': 'none of the kernels of this path is HBM-bound: the FFT kernels are issue / FP64-pipe limited '
                             '(55-71 % of the issue slots, FP64 pipe 42 % in the analysis kernels), the tile products run at '
                             '57-60 % of the FP32 FMA pipe; DRAM traffic equals the algorithmic bytes (profiles/README.md)'},
        'halves': {'analysis_ms': ana_ms, 'synthesis_ms': syn_ms,
                   'chain_algorithmic_bytes_per_step': int(chain_bytes),
                   'chain_gbs': chain_bytes / (ms_per_step * 1e-3) / 1e9},
        'kernels': kernels,
        'clocks': clocks,
    }
    if cpu is not None:
        line['cpu_baseline'] = {'value': cpu['value'], 'unit': 'frames/s', 'cores': cpu['cores'], 'kind': cpu['kind'],
                                'sample': cpu['sample']}
    print(json.dumps(line))


if __name__ == '__main__':
    main()

"""`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a tiny sample: runs without a GPU, prints
ONE JSON line with the contract's keys, and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                        '--cpu-utts', '2', '--cpu-dur', '0.4'], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith('{')]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['gpu_launches'] == 0 and 'workload' in d['config']


def test_reference_arm_only_rank_zero_speaks():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == []


def test_fp32_fma_view_of_the_tile_products():
    """bench.fp32_fma_view: algorithmic FLOP/s of the two tile products against the derived non-tensor FP32 peak."""
    sys.path.insert(0, ROOT)
    import bench
    nf, nv, H = 116432, 54257, 2049
    ks = [dict(name='k_mel_gemm', ms_per_step=1.44), dict(name='k_mel_unwarp', ms_per_step=1.057),
          dict(name='k_analysis<logp>', ms_per_step=1.6)]
    bench.fp32_fma_view(ks, nf, nv, H, 1965.0, 148)
    g, u = ks[0]['fp32_fma'], ks[1]['fp32_fma']
    assert 'fp32_fma' not in ks[2]
    assert abs(g['peak_tflops'] - 148 * 128 * 2 * 1.965e9 / 1e12) < 1e-9
    assert g['flops_per_step'] == 2 * (nf * 60 + 2 * nv * 58) * 2048
    assert u['flops_per_step'] == 2 * (nf * 60 * 2049 + 2 * nv * 45 * 512)
    assert 0.0 < u['frac'] < g['frac'] < 1.0
    assert abs(g['achieved_tflops'] - g['flops_per_step'] / 1.44e-3 / 1e12) < 1e-9

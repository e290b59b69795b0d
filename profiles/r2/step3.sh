#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_compressed_analysis.py tests/test_gpu_compressed_synthesis.py tests/test_gpu_natural.py tests/test_gpu_host_pipeline.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 10 --warmup 3 > gpurun_out/r3a_bench.log 2> gpurun_out/r3a_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r3a_bench.err
timeout 600 python bench.py --no-cpu-baseline --no-extras --no-rng-overlap --steps 10 --warmup 3 > gpurun_out/r3b_bench.log 2> gpurun_out/r3b_bench.err
python - <<'PY'
import json
for f in ('r3a','r3b'):
    for l in open('gpurun_out/%s_bench.log' % f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, 'value %.2fM  e2e %.2fM  e2e_f64 %.2fM ms %.3f launches %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e_float64_api']['value']/1e6, d['ms_per_step'], d['gpu_launches']))
            print('  ', ' '.join('%s=%.3f' % (k['name'], k['ms_per_step']) for k in d['kernels']))
PY
timeout 600 python profiles/r2/e2e_diag.py 128 > gpurun_out/e2e_diag.out 2> gpurun_out/e2e_diag.err
tail -40 gpurun_out/e2e_diag.err

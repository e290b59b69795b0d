#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lossless.py tests/test_gpu_compressed_synthesis.py tests/test_gpu_natural.py -x -q 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2n_bench.log 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2n_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2n_bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.2fM  e2e %.2fM  ms %.3f launches %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['gpu_launches']))
        for k in d['kernels'][:4]: print('  %-36s %.3f ms  hbm frac %.3f' % (k['name'], k['ms_per_step'], k['frac']))
        print('lossless', d['lossless']['value']/1e6, [(k['name'], round(k['ms_per_step'],3)) for k in d['lossless']['kernels']])
PY

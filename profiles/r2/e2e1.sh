#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host_pipeline.py tests/test_gpu_batch_files.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2k_bench.log 2> gpurun_out/r2k_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2k_bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.2fM  e2e %.2fM  e2e_f64 %.2fM ms %.3f launches %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e_float64_api']['value']/1e6, d['ms_per_step'], d['gpu_launches']))
        for k in ('lossless','extract_tts','generate_16k','stream','error'):
            if k in d: print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in d[k].items() if kk not in ('workload','kernels','note')} if isinstance(d[k],dict) else d[k])
PY

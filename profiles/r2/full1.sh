#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_gputests.txt 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/r2i_gputests.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.log 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2i_bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.2fM  e2e %.2fM  ms %.3f launches %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['gpu_launches']))
        print('fma peaks', d['fma_peaks_tflops'], 'chain', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['chain'].items() if k!='note'})
        for k in d['kernels']: print('  %-36s %.3f ms  hbm frac %.3f  compute %s' % (k['name'], k['ms_per_step'], k['frac'], (round(k['compute']['frac'],3), k['compute']['pipe']) if 'compute' in k else None))
        for k in ('lossless','extract_tts','generate_16k','stream','error'):
            if k in d: print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in d[k].items() if kk not in ('workload','kernels','note')} if isinstance(d[k],dict) else d[k])
        print('cpu', d.get('cpu_baseline'))
PY

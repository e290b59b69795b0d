#!/bin/bash
mkdir -p gpurun_out
for g in 1 2 3 4 6 0; do
  MPB_PIPELINE_GROUPS=$g timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 10 > gpurun_out/grp_$g.log 2> gpurun_out/grp.err
  python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith(chr(123)): d=json.loads(l); e=d['e2e']; print(sys.argv[1], 'value %.2fM e2e %.2fM cpu %.1f ms/step f64 %.2fM' % (d['value']/1e6, e['value']/1e6, e['host_cpu_ms_per_step'], d['e2e_float64_api']['value']/1e6))
" gpurun_out/grp_$g.log
done

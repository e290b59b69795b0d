#!/bin/bash
# One rank of an 8-rank job has 4 of the box's 32 cores: emulate it at N = 1 with taskset.
mkdir -p gpurun_out
for sync in spin block; do for w in 2 3 4; do
  MPB_SYNC=$sync MPB_HOST_THREADS=4 timeout 300 taskset -c 0-3 python bench.py --no-cpu-baseline --no-extras --steps 10 --e2e-workers $w > gpurun_out/fc_${sync}_$w.log 2> gpurun_out/fc.err
  python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith(chr(123)): d=json.loads(l); e=d['e2e']; print(sys.argv[1], 'value %.2fM e2e %.2fM cpu %.1f ms/step f64 %.2fM' % (d['value']/1e6, e['value']/1e6, e['host_cpu_ms_per_step'], d['e2e_float64_api']['value']/1e6))
" gpurun_out/fc_${sync}_$w.log
done; done
for sync in spin block; do
  MPB_SYNC=$sync timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 10 > gpurun_out/fc_all_$sync.log 2> gpurun_out/fc.err
  python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith(chr(123)): d=json.loads(l); e=d['e2e']; print(sys.argv[1], 'value %.2fM e2e %.2fM cpu %.1f ms/step f64 %.2fM' % (d['value']/1e6, e['value']/1e6, e['host_cpu_ms_per_step'], d['e2e_float64_api']['value']/1e6))
" gpurun_out/fc_all_$sync.log
done

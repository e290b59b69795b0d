#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_full_size.py -x -q 2>&1 | tail -2
for cfg in "1 1" "1 0" "0 1" "0 0"; do set -- $cfg
  MPB_NOISE_EARLY=$1 MPB_SIDE_PRIO=$2 timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 20 --warmup 3 --e2e-utts 8 > gpurun_out/r3f_$1$2.log 2> gpurun_out/r3f.err
  python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith(chr(123)): d=json.loads(l); print(sys.argv[1], 'value %.2fM ms %.3f' % (d['value']/1e6, d['ms_per_step']), d['halves'] if 'halves' in d else '')
" gpurun_out/r3f_$1$2.log
done

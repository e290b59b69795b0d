#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --utts 64 --e2e-utts 2 --steps 1 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_mel_unwarp_tc|k_mel_cos' -s 6 -c 3 -f -o gpurun_out/r2_tc_unwarp $B > gpurun_out/r2_ncu_tc.log 2>&1
echo rc=$?; ls -la gpurun_out/r2_tc_unwarp.ncu-rep

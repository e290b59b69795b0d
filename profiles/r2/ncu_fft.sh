#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-extras --utts 32 --e2e-utts 2 --steps 1 --warmup 3"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_analysis|k_synthesis_compressed' -s 9 -c 3 -f -o gpurun_out/r2_fft $B > gpurun_out/r2_ncu_fft.log 2>&1
echo rc=$?; ls -la gpurun_out/r2_fft.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
MPB_TC_TRACE=gpurun_out/tc_trace0.bin timeout 120 python bench.py --no-cpu-baseline --steps 1 --warmup 3 --e2e-utts 2 > gpurun_out/r2f_trace0.log 2>&1
MPB_TC_DEBUG=255 MPB_TC_TRACE=gpurun_out/tc_trace255.bin timeout 120 python bench.py --no-cpu-baseline --steps 1 --warmup 3 --e2e-utts 2 > gpurun_out/r2f_trace255.log 2>&1
ls -la gpurun_out/tc_trace*

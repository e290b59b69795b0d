#!/bin/bash
mkdir -p gpurun_out
python profiles/parity_report.py > gpurun_out/r2b_parity_default.txt 2>&1
MPB_LOGP_F32=1 python profiles/parity_report.py > gpurun_out/r2b_parity_logp_f32.txt 2>&1
MPB_MEL_TC=1 python profiles/parity_report.py > gpurun_out/r2b_parity_tc1.txt 2>&1
python -m pytest tests/test_gpu_natural.py -q > gpurun_out/r2b_natural_tests.txt 2>&1
tail -3 gpurun_out/r2b_natural_tests.txt
grep -E "natural|band" gpurun_out/r2b_parity_logp_f32.txt | grep -E "compressed"

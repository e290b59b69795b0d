#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_compressed_synthesis.py tests/test_gpu_natural.py tests/test_gpu_compressed_analysis.py -x -q 2>&1 | tail -5
timeout 120 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2h_bench.log 2>&1
python profiles/show_bench.py gpurun_out/r2h_bench.log | head -16
timeout 120 python profiles/parity_report.py 2>&1 | grep -E "synthesis_from_compressed"

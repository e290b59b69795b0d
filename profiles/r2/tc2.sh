#!/bin/bash
# warp-product tensor-core kernel with fused finish: parity (incl. natural speech), timing
mkdir -p gpurun_out
T="timeout 180"
$T python -m pytest tests/test_gpu_compressed_analysis.py tests/test_gpu_natural.py tests/test_gpu_host_pipeline.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2c_tests.txt 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2c_tests.txt
$T python profiles/parity_report.py > gpurun_out/r2c_parity.txt 2>&1
grep -E "compressed" gpurun_out/r2c_parity.txt | grep -v synthesis
$T python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2c_bench.log 2>&1
echo "bench rc=$?"; python profiles/show_bench.py gpurun_out/r2c_bench.log | head -14
MPB_MEL_TC=0 $T python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2c_bench_fma.log 2>&1
python profiles/show_bench.py gpurun_out/r2c_bench_fma.log | head -14

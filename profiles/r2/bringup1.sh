#!/bin/bash
# Round 2, first GPU call: parity + timing of the tensor-core tile products (MPB_MEL_TC masks) and the float32
# log-periodogram FFT (MPB_LOGP_F32=1).  Runs ON THE GPU BOX; every step under its own timeout.
mkdir -p gpurun_out
S=gpurun_out/r2a_summary.log
: > $S
T="timeout 150"
python -c "import torch; print(torch.cuda.get_device_name(0))" >> $S 2>&1
for m in 1 3 11; do
  MPB_MEL_TC=$m $T python -m pytest tests/test_gpu_compressed_analysis.py tests/test_gpu_compressed_synthesis.py tests/test_gpu_full_size.py \
      tests/test_gpu_host_pipeline.py -x -q > gpurun_out/r2a_tests_$m.log 2>&1
  echo "MPB_MEL_TC=$m tests rc=$?" >> $S
  tail -3 gpurun_out/r2a_tests_$m.log >> $S
done
MPB_LOGP_F32=1 $T python -m pytest tests/test_gpu_compressed_analysis.py tests/test_gpu_full_size.py tests/test_gpu_host_pipeline.py -x -q > gpurun_out/r2a_tests_f32.log 2>&1
echo "MPB_LOGP_F32=1 tests rc=$?" >> $S; tail -3 gpurun_out/r2a_tests_f32.log >> $S
for m in 0 1 2 3 5 9 11 15; do
  MPB_MEL_TC=$m $T python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2a_bench_$m.log 2>&1
  echo "MPB_MEL_TC=$m bench rc=$?" >> $S
  python profiles/show_bench.py gpurun_out/r2a_bench_$m.log 2>/dev/null | head -14 >> $S
done
MPB_LOGP_F32=1 $T python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2a_bench_f32.log 2>&1
echo "LOGP_F32 bench rc=$?" >> $S
python profiles/show_bench.py gpurun_out/r2a_bench_f32.log 2>/dev/null | head -14 >> $S
MPB_LOGP_F32=1 MPB_MEL_TC=11 $T python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2a_bench_f32_tc11.log 2>&1
echo "LOGP_F32 + TC11 bench rc=$?" >> $S
python profiles/show_bench.py gpurun_out/r2a_bench_f32_tc11.log 2>/dev/null | head -14 >> $S
$T python bench.py --no-cpu-baseline --workload lossless --steps 10 --warmup 3 > gpurun_out/r2a_bench_lossless.log 2>&1
echo "lossless bench rc=$?" >> $S
python profiles/show_bench.py gpurun_out/r2a_bench_lossless.log 2>/dev/null | head -14 >> $S
cat $S

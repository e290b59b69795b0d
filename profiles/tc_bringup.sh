#!/bin/bash
# Runs ON THE GPU BOX (gpurun --timeout 900 -- 'bash profiles/tc_bringup.sh'): first measurements of the experimental
# tensor-core tile products (magphase_b200/csrc/mpb_mel_tc.cu; MPB_MEL_TC bit 0 = warp product, bit 1 = un-warp product,
# bit 2 = warp product with three stages of row loads in flight, bit 3 = K-slice sums inside the warp-product kernel).
# Every step runs under its own timeout so that a hanging kernel (mbarrier deadlock) cannot hold the box.
# Order: cheapest evidence first.  Logs land in gpurun_out/tc_*.log.
mkdir -p gpurun_out
T="timeout 120"
for m in 1 5 9 2 3; do
  MPB_MEL_TC=$m $T python -m pytest tests/test_gpu_compressed_analysis.py tests/test_gpu_compressed_synthesis.py tests/test_gpu_full_size.py \
      tests/test_gpu_host_pipeline.py -x -q > gpurun_out/tc_tests_$m.log 2>&1
  echo "MPB_MEL_TC=$m tests rc=$?" | tee -a gpurun_out/tc_summary.log
  tail -3 gpurun_out/tc_tests_$m.log | tee -a gpurun_out/tc_summary.log
done
for m in 0 1 5 9 13 2 15; do
  MPB_MEL_TC=$m $T python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/tc_bench_$m.log 2>&1
  echo "MPB_MEL_TC=$m bench rc=$?" | tee -a gpurun_out/tc_summary.log
  python profiles/show_bench.py gpurun_out/tc_bench_$m.log 2>/dev/null | head -20 | tee -a gpurun_out/tc_summary.log
done
# float32 butterflies in the fused compressed analysis (MPB_LOGP_F32=1), alone and with the best tensor-core setting
MPB_LOGP_F32=1 $T python -m pytest tests/test_gpu_compressed_analysis.py tests/test_gpu_full_size.py tests/test_gpu_host_pipeline.py -x -q > gpurun_out/tc_tests_f32.log 2>&1
echo "MPB_LOGP_F32=1 tests rc=$?" | tee -a gpurun_out/tc_summary.log; tail -3 gpurun_out/tc_tests_f32.log | tee -a gpurun_out/tc_summary.log
MPB_LOGP_F32=1 $T python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/tc_bench_f32.log 2>&1
python profiles/show_bench.py gpurun_out/tc_bench_f32.log 2>/dev/null | head -20 | tee -a gpurun_out/tc_summary.log
MPB_LOGP_F32=1 MPB_MEL_TC=15 $T python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/tc_bench_f32_tc15.log 2>&1
python profiles/show_bench.py gpurun_out/tc_bench_f32_tc15.log 2>/dev/null | head -20 | tee -a gpurun_out/tc_summary.log
# one full capture of the two tensor-core kernels (launch counts: see the launch list first if -s / -c need adjusting)
B="python bench.py --no-cpu-baseline --utts 32 --e2e-utts 2 --steps 1 --warmup 3"
MPB_MEL_TC=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/tc_launches.csv $B > gpurun_out/tc_ncu_l.log 2>&1
MPB_MEL_TC=3 timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_mel_gemm_tc|k_mel_unwarp_tc' -s 6 -c 3 -f \
    -o gpurun_out/tc_prof $B > gpurun_out/tc_ncu_p.log 2>&1
ls -la gpurun_out/tc_* | tee -a gpurun_out/tc_summary.log

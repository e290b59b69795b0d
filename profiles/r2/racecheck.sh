#!/bin/bash
# compute-sanitizer racecheck over the shared-memory exchanges of the FFT engine, the overlap-add accumulators and the
# TMA-staged synthesis kernel (race detection).  Small inputs: racecheck serialises the kernels.
mkdir -p gpurun_out
export MPB_MEL_TC=0     # the FMA tile products: racecheck does not model tcgen05 / TMEM traffic
timeout -k 5 230 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 \
  python -m pytest tests/test_gpu_lossless.py tests/test_gpu_compressed_synthesis.py -x -q \
  -k "test_analysis_vs_oracle_48k or test_synthesis_vs_oracle or test_arbitrary_win_func or test_copy_synthesis_low_dim_chain" \
  > gpurun_out/racecheck.log 2>&1
echo "rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error|error" gpurun_out/racecheck.log | sort | uniq -c | sort -rn | head -20

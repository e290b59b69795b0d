#!/bin/bash
# compute-sanitizer passes over this round's kernels (race detection / sanitizers), small inputs.
mkdir -p gpurun_out
SEL='tests/test_gpu_lossless.py tests/test_gpu_compressed_synthesis.py tests/test_gpu_compressed_analysis.py tests/test_gpu_natural.py tests/test_gpu_griffin_lim.py'
KX='not jump_ahead and not full_size and not legacy_stream_on_device'
run() {  # name, tool args...
  local name=$1; shift
  timeout -k 5 "$TMO" compute-sanitizer "$@" python -m pytest $SEL -x -q -k "$KX" > gpurun_out/san_$name.log 2>&1
  echo "== $name rc=$? : $(grep -E ' passed| failed' gpurun_out/san_$name.log | tail -1) | $(grep -E 'SUMMARY' gpurun_out/san_$name.log | tail -1)"
}
TMO=150 run race_tc   --tool racecheck --racecheck-report all --print-limit 10
TMO=150 run memcheck  --tool memcheck --print-limit 10
TMO=100 run synccheck --tool synccheck --print-limit 10

#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- 'bash profiles/capture.sh'): launch lists and `--set full` captures of one bench step per
# workload at --utts 32 (29,108 frames per launch).  Reports land in gpurun_out/; profiles/summarise.sh turns them into
# the text files committed under profiles/<round>/.  Numbers printed by bench.py under ncu are never bench values.
set -x
B="python bench.py --no-cpu-baseline --utts 32 --e2e-utts 2 --steps 1 --warmup 3"
K='regex:k_analysis|k_mel_|k_synthesis|k_noise|k_voiced|k_unwarp'
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_compressed.csv $B > gpurun_out/ncu_lc.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lossless.csv $B --workload lossless > gpurun_out/ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 27 -c 9 -f -o gpurun_out/prof_compressed $B > gpurun_out/ncu_pc.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 6 -c 2 -f -o gpurun_out/prof_lossless $B --workload lossless > gpurun_out/ncu_pl.log 2>&1
ls -la gpurun_out/*.ncu-rep

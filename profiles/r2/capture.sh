#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- 'bash profiles/r2/capture.sh'): launch lists and `--set full` captures of bench steps per
# workload at --utts 32 (29,108 frames per launch).  Reports land in gpurun_out/; profiles/summarise.py turns them into
# the text files committed under profiles/r2/.  Numbers printed by bench.py under ncu are never bench values.
set -x
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-extras --utts 32 --e2e-utts 2 --steps 1 --warmup 3"
K='regex:^(void )?(mpb::)?(k_analysis|k_mel_|k_synthesis|k_noise|k_voiced|k_unwarp|k_mt)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_compressed.csv $B > gpurun_out/ncu_lc.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_lossless.csv $B --workload lossless > gpurun_out/ncu_ll.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k "$K" -c 24 -f -o gpurun_out/prof_compressed $B > gpurun_out/ncu_pc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -c 4 -f -o gpurun_out/prof_lossless $B --workload lossless > gpurun_out/ncu_pl.log 2>&1
python profiles/parity_report.py > gpurun_out/parity_report.txt 2>&1
# summarise on the box (ncu is here) and keep the transfer under gpurun's 64 MiB: the text summaries always travel, the
# reports only while they fit
python profiles/summarise.py gpurun_out gpurun_out/r2_summary > gpurun_out/summarise.log 2>&1
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/prof_lossless.ncu-rep
[ $(stat -c %s gpurun_out/prof_compressed.ncu-rep) -gt 52000000 ] && rm -f gpurun_out/prof_compressed.ncu-rep
du -sh gpurun_out

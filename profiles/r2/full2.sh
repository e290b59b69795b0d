#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r3e_bench.log 2> gpurun_out/r3e_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r3e_bench.err
python profiles/show_bench.py gpurun_out/r3e_bench.log 2>&1 | head -30
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3e_ref.log 2> gpurun_out/r3e_ref.err
echo "ref rc=$?"; tail -2 gpurun_out/r3e_ref.log | cut -c1-400

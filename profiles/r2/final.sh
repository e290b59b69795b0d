#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/final_bench.err
python profiles/show_bench.py gpurun_out/final_bench.log 2>&1 | head -16
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.log 2> gpurun_out/final_ref.err
echo "ref rc=$?"; tail -1 gpurun_out/final_ref.log | cut -c1-200

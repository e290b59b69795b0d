#!/bin/bash
# aggregate pinned-copy bandwidth with every GPU of the box copying at once (one process per GPU)
N=${1:-8}
for i in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$i python profiles/r2/pcie_bw.py > gpurun_out/pcie_$i.txt 2>&1 & done
wait
grep -h "both at once\|alone" gpurun_out/pcie_*.txt | sort | uniq -c | sort -rn | head -30

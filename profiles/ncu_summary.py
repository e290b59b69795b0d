#!/usr/bin/env python
"""Print the handful of ncu metrics we track from a .ncu-rep (run where ncu is installed; no GPU needed).
    python profiles/ncu_summary.py gpurun_out/prof_x.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'sm__inst_executed_pipe_lsu.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('==', vals[hdr.index('Kernel Name')][:110])
        for i, h in enumerate(hdr):
            try:
                fv = float(vals[i].replace(',', ''))
            except ValueError:
                fv = 0.0
            if h in WANT or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and fv > 0.3):
                print('  %-78s %-14s %s' % (h, units[i], vals[i]))


if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)

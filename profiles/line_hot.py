#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel of an .ncu-rep (captured with --import-source on,
kernels compiled with -lineinfo):   python profiles/line_hot.py <report.ncu-rep> <kernel regex> [top]
Reads `ncu --page source --csv --print-source cuda,sass`; needs ncu, no GPU."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def main(path, kernel_rx, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # blocks: a "Kernel Name" row (if present) followed by file blocks; find header rows
    kernel, file_name, hdr = None, None, None
    per_line = defaultdict(lambda: [0.0, 0.0, ''])   # (file, line) -> [warp instr, samples, text]
    per_op = defaultdict(lambda: [0.0, 0.0])
    total_i = total_s = 0.0
    cur_line, cur_text = None, ''
    active = False
    for r in rows:
        if not r:
            continue
        if r[0] in ('Kernel Name', 'Function Name'):
            kernel = r[1]
            active = re.search(kernel_rx, kernel) is not None
            continue
        if r[0] in ('File Name', 'File Path'):
            file_name = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            i_addr = hdr.index('Address')
            i_sass = i_addr + 1
            i_smp, i_ins = hdr.index('# Samples'), hdr.index('Instructions Executed')
            continue
        if hdr is None or not active:
            continue
        if r[0].strip():
            cur_line, cur_text = r[0], r[1].strip()
        if len(r) > i_ins and r[i_addr].startswith('0x'):
            ins, smp = float(r[i_ins] or 0), float(r[i_smp] or 0)
            e = per_line[(file_name, cur_line)]
            e[0] += ins; e[1] += smp; e[2] = cur_text
            op = r[i_sass].split()
            op = (op[1] if op and op[0].startswith('@') else (op[0] if op else '?')).split('.')[0]
            per_op[op][0] += ins; per_op[op][1] += smp
            total_i += ins; total_s += smp
    print('kernel /%s/: %.0f warp-instr, %.0f samples' % (kernel_rx, total_i, total_s))
    print('%-22s %6s %6s  %s' % ('file:line', 'instr%', 'smp%', 'source'))
    for (f, l), (i, s, t) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%-22s %6.2f %6.2f  %s' % ('%s:%s' % (f, l), 100 * i / max(total_i, 1), 100 * s / max(total_s, 1), t[:110]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)

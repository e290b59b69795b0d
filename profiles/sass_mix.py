#!/usr/bin/env python
"""Instruction mix and stall samples per SASS opcode from a .ncu-rep (needs --import-source on at capture time).
    python profiles/sass_mix.py gpurun_out/prof_x.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, top=22):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(out))
    hdr = None
    inst, samp = defaultdict(float), defaultdict(float)
    for r in rows:
        if r and r[0] == 'Address':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2 or not r[0].startswith('0x'):
            continue
        op = r[1].split()
        op = op[1] if op and op[0].startswith('@') else (op[0] if op else '?')
        op = op.split('.')[0]
        inst[op] += float(r[hdr.index('Instructions Executed')] or 0)
        samp[op] += float(r[hdr.index('# Samples')] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    print('%-10s %12s %7s %9s' % ('opcode', 'warp-instr', 'share', 'samples%'))
    for op, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
        print('%-10s %12.0f %6.1f%% %8.1f%%' % (op, n, 100 * n / ti, 100 * samp[op] / max(ts, 1)))
    print('total warp-instr %.0f' % ti)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)

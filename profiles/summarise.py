#!/usr/bin/env python
"""Turns the .ncu-rep files written by profiles/capture.sh into the per-kernel text summaries committed under
profiles/<round>/ (run where ncu is installed; no GPU needed):

    python profiles/summarise.py gpurun_out profiles/r1b

Per kernel: the tracked raw metrics (ncu_summary.WANT + stall reasons > 0.3 per issue), the SASS opcode mix with stall
samples, and one line per file in <round>/traffic_per_frame.json (DRAM bytes and time per frame) for bench.py."""
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import WANT  # noqa: E402

SHORT = [  # (regex on the demangled name, bench.py kernel key, file stem)
    (r'k_analysis<double, float, float, \d+, 3>', 'k_analysis<logp>', 'k_analysis_logp'),
    (r'k_analysis<float, float, double, \d+, 2>', 'k_analysis<noise_logsq>', 'k_analysis_noise_logsq'),
    (r'k_analysis<double, float, float, \d+, 0>', 'k_analysis', 'k_analysis'),
    (r'k_mel_warp_tc', 'k_mel_warp_tc', 'k_mel_warp_tc'),
    (r'k_mel_cos', 'k_mel_cos', 'k_mel_cos'),
    (r'k_mel_unwarp_tc', 'k_mel_unwarp_tc', 'k_mel_unwarp_tc'),
    (r'k_unwarp_prep', 'k_mel_unwarp_tc', 'k_unwarp_prep'),                     # bench.py brackets it with the product
    (r'k_mt19937_stream', 'k_mt19937_stream+k_mt_to_uniform', 'k_mt19937_stream'),
    (r'k_mt_to_uniform', 'k_mt19937_stream+k_mt_to_uniform', 'k_mt_to_uniform'),
    (r'k_mel_gemm', 'k_mel_gemm', 'k_mel_gemm'),
    (r'k_mel_finish', 'k_mel_finish', 'k_mel_finish'),
    (r'k_mel_unwarp', 'k_mel_unwarp', 'k_mel_unwarp'),
    (r'k_unwarp_tile_flags', None, 'k_unwarp_tile_flags'),
    (r'k_voiced_compact', 'k_voiced_compact', 'k_voiced_compact'),
    (r'k_noise_gain', 'k_noise_gain', 'k_noise_gain'),
    (r'k_synthesis_compressed', 'k_synthesis_compressed', 'k_synthesis_compressed'),
    (r'k_synthesis_lossless', 'k_synthesis_lossless', 'k_synthesis_lossless'),
]


def classify(name):
    for rx, key, stem in SHORT:
        if re.search(rx, name):
            return key, stem
    return None, re.sub(r'\W+', '_', name)[:40]


def raw_page(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def sass_blocks(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout.splitlines()
    blocks, cur = [], None
    for r in csv.reader(out):
        if r and r[0] == 'Address':
            cur = {'hdr': r, 'rows': []}
            blocks.append(cur)
        elif cur is not None and r and r[0].startswith('0x'):
            cur['rows'].append(r)
    return blocks


def opcode_mix(block, top=16):
    h = block['hdr']
    i_s, i_i = h.index('# Samples'), h.index('Instructions Executed')
    inst, samp = defaultdict(float), defaultdict(float)
    for r in block['rows']:
        op = r[1].split()
        op = op[1] if op and op[0].startswith('@') else (op[0] if op else '?')
        op = op.split('.')[0]
        inst[op] += float(r[i_i] or 0)
        samp[op] += float(r[i_s] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    lines = ['%-10s %12s %7s %9s' % ('opcode', 'warp-instr', 'share', 'samples%')]
    for op, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
        lines.append('%-10s %12.0f %6.1f%% %8.1f%%' % (op, n, 100 * n / max(ti, 1), 100 * samp[op] / max(ts, 1)))
    lines.append('total warp-instr %.0f' % ti)
    return lines


def main(src_dir, dst_dir, frames):
    """One bench step per report: the rows up to and including the first launch of the chain's last kernel.  DRAM bytes and
    time are SUMMED per bench.py kernel key over that step (a key can cover several launches: the two pipeline groups of
    the analysis half, the preparation kernels in front of the un-warp product); the per-kernel text file is written for
    the longest launch of each kernel."""
    os.makedirs(dst_dir, exist_ok=True)
    traffic, times, launches = defaultdict(float), defaultdict(float), defaultdict(int)
    for rep, prefix, last in (('prof_compressed.ncu-rep', 'c_', 'k_synthesis_compressed'),
                              ('prof_lossless.ncu-rep', 'l_', 'k_synthesis_lossless')):
        path = os.path.join(src_dir, rep)
        if not os.path.exists(path):
            continue
        hdr, units, rows = raw_page(path)
        blocks = sass_blocks(path)
        per_kernel = len(blocks) // max(len(rows), 1) or 1          # the source page repeats every kernel per view
        best = {}
        for k, vals in enumerate(rows):
            name = vals[hdr.index('Kernel Name')]
            key, stem = classify(name)
            get = {}
            for i, h in enumerate(hdr):
                try:
                    fv = float(vals[i].replace(',', ''))
                except ValueError:
                    fv = 0.0
                get[h] = (fv, units[i])

            def to_bytes(m):
                v, u = get[m]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            v, u = get['gpu__time_duration.sum']
            us = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1.0)
            if key:
                traffic[key] += to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
                times[key] += us
                launches[key] += 1
            if stem not in best or us > best[stem][0]:
                best[stem] = (us, k, name, vals)
            if last in name:
                break
        for stem, (us, k, name, vals) in best.items():
            lines = ['== ' + name[:150]]
            for i, h in enumerate(hdr):
                try:
                    fv = float(vals[i].replace(',', ''))
                except ValueError:
                    fv = 0.0
                if h in WANT or (h.startswith('smsp__average_warps_issue_stalled') and
                                 h.endswith('per_issue_active.ratio') and fv > 0.3):
                    lines.append('  %-78s %-14s %s' % (h, units[i], vals[i]))
            lines += ['', '-- SASS opcode mix --'] + opcode_mix(blocks[k * per_kernel])
            open(os.path.join(dst_dir, prefix + stem + '.txt'), 'w').write('\n'.join(lines) + '\n')
            print(prefix + stem, '%.1f us' % us)
    json.dump({'source': 'ncu --set full --clock-control none, ONE bench step over %d frames (bench.py --utts 32), '
                         'dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of each kernel key in that '
                         'step; see the per-kernel .txt files' % frames,
               'frames_per_captured_step': frames, 'launches_per_step': dict(launches),
               'dram_bytes_per_frame': {k: v / frames for k, v in traffic.items()},
               'ncu_us_per_captured_step': dict(times)}, open(os.path.join(dst_dir, 'traffic_per_frame.json'), 'w'), indent=1)
    for f in ('launches_compressed.csv', 'launches_lossless.csv'):
        p = os.path.join(src_dir, f)
        if os.path.exists(p):
            keep = [l for l in open(p) if l.startswith('"') or l.startswith('ID')]
            open(os.path.join(dst_dir, f), 'w').writelines(keep)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 29108)

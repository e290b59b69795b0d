"""Offline sweep (test infrastructure, not part of the timed CPU suite): the NumPy oracle against the py3-translated
reference (oracle/_ref, built by make_ref.py from /root/reference) on randomly generated inputs.

    python oracle/sweep_vs_ref.py [n_analysis_seeds] [n_synthesis_seeds]

  * analysis: pitch-mark patterns of tests/test_mark_patterns_cpu.py (mark at 0, 1-sample shifts, half-integers, periods beyond
    fft_len, mark on the last sample) at fft_len 1024 / 2048 / 4096 -> spectra, features, lossless resynthesis;
  * synthesis: random compressed features with all-voiced / all-unvoiced / mixed voicing, 48 and 16 kHz, variable and
    constant rate, with and without the output high-pass and the voiced-noise window; exceptions must match too.
Last run of round 2: 1,598 analysis cases and 240 synthesis cases, no mismatch."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'oracle', '_ref'), os.path.join(ROOT, 'tests')]
warnings.simplefilter('ignore')
import make_ref
make_ref.build()
import magphase as mp                      # the reference
import magphase_oracle as orc
from test_mark_patterns_cpu import mark_pattern


def sweep_analysis(n_seeds):
    done = bad = 0
    for seed in range(n_seeds):
        for style in ('speech', 'tiny', 'long', 'fractional'):
            fft_len = (1024, 2048, 4096)[seed % 3]
            rng = np.random.default_rng(seed * 7 + len(style))
            n = int(rng.integers(3000, 40000))
            sig = rng.uniform(-1, 1, n)
            pm = mark_pattern(rng, n, int(rng.integers(3, 40)), style)
            if pm.size < 2:
                continue
            voi = (rng.random(pm.size) < 0.6).astype(float)
            try:
                a, sa = mp.analysis_with_del_comp_from_pm(sig.copy(), 48000, pm.copy(), fft_len=fft_len)
                b, sb = orc.analysis_fft_from_pm(sig, 48000, pm, fft_len=fft_len)
                assert np.array_equal(sa, sb) and np.max(np.abs(a - b)) < 1e-10
                fa = mp.compute_lossless_feats(a, sa, voi, 48000)
                fb = orc.compute_lossless_feats(b, sb, voi, 48000)
                assert np.array_equal(fa[3], fb[3], equal_nan=True)
                if np.all(np.isfinite(fb[3])):
                    ya = mp.synthesis_from_lossless(fa[0].copy(), fa[1].copy(), fa[2].copy(), fa[3].copy(), 48000)
                    yb = orc.synthesis_from_lossless(*fb, 48000)
                    assert ya.shape == yb.shape and np.max(np.abs(ya - yb)) < 1e-10
                done += 1
            except Exception as e:          # noqa: BLE001
                bad += 1
                print('ANALYSIS MISMATCH', seed, style, fft_len, type(e).__name__, str(e)[:100])
    return done, bad


def sweep_synthesis(n_seeds):
    done = bad = 0
    for seed in range(n_seeds):
        rng = np.random.default_rng(seed)
        fs = (48000, 16000)[seed % 2]
        n = int(rng.integers(5, 60))
        style = seed % 5
        voi = np.ones(n, bool) if style == 0 else (np.zeros(n, bool) if style == 1 else rng.random(n) < rng.uniform(0.2, 0.8))
        lf0 = np.where(voi, np.log(rng.uniform(60, 380, n)), -1e10)
        mag = rng.standard_normal((n, 60)) * 0.5 - 2.0
        re, im = rng.uniform(-1, 1, (n, 45)), rng.uniform(-1, 1, (n, 45))
        for kw in (dict(b_out_hpf=False), dict(b_out_hpf=True), dict(b_const_rate=True, b_out_hpf=False),
                   dict(b_voi_ap_win=False, b_out_hpf=False)):
            res = []
            for impl in (mp, orc):
                try:
                    np.random.seed(seed)
                    res.append(impl.synthesis_from_compressed(mag.copy(), re.copy(), im.copy(), lf0.copy(), fs, **kw))
                except Exception as e:      # noqa: BLE001
                    res.append(type(e).__name__)
            a, b = res
            if isinstance(a, str) or isinstance(b, str):
                if a is not b and not (isinstance(a, str) and a == b):
                    bad += 1
                    print('SYNTHESIS EXCEPTION MISMATCH', seed, kw, a if isinstance(a, str) else 'ok', b if isinstance(b, str) else 'ok')
                continue
            done += 1
            tol = 1e-6 if kw.get('b_out_hpf') else 1e-10     # the reference's 4th-order direct-form IIR amplifies last-bit differences
            if a.shape != b.shape or not np.allclose(a, b, rtol=0, atol=tol, equal_nan=True):
                bad += 1
                print('SYNTHESIS MISMATCH', seed, fs, n, style, kw)
    return done, bad


if __name__ == '__main__':
    na = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    print('analysis: %d cases, %d mismatches' % sweep_analysis(na))
    print('synthesis: %d cases, %d mismatches' % sweep_synthesis(ns))

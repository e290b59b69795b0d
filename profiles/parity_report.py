"""Measured error of every output of the two chains against the oracle (run on the GPU box): RMS and max-abs per array,
48 kHz / FFT 4096 and 16 kHz / FFT 2048.  The tolerance of the parity tests is 1e-5 RMS (BASELINE.json north_star)."""
import sys
import warnings

sys.path.insert(0, '.')
sys.path.insert(0, 'oracle')
import numpy as np

import magphase_oracle as orc
import magphase_b200.magphase as mp
from magphase_b200.synth import synth_utterance

warnings.simplefilter('ignore')


def err(a, b):
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    return '%9.2e %9.2e' % (float(np.sqrt(np.mean(d ** 2))), float(d.max()))


print('%-44s %9s %9s' % ('output', 'rms', 'max'))
for fs in (48000, 16000):
    sig, pm, voi = synth_utterance(11, fs=fs, dur_s=2.0)
    tag = '%dk ' % (fs // 1000)
    got = mp.analysis_lossless_from_pm(sig, fs, pm, voi)
    ref = orc.analysis_lossless_from_pm(sig, fs, pm, voi)
    for n, a, b in zip(('mag', 'real', 'imag', 'f0'), got[:4], ref[:4]):
        print('%-44s %s' % (tag + 'analysis_lossless ' + n, err(a, b)))
    print('%-44s %s' % (tag + 'synthesis_from_lossless', err(mp.synthesis_from_lossless(*got[:4], fs), orc.synthesis_from_lossless(*ref[:4], fs))))
    cg = mp.analysis_compressed_from_pm(sig, fs, pm, voi, mag_dim=60, phase_dim=45)
    cr = orc.analysis_compressed_from_pm(sig, fs, pm, voi, mag_dim=60, phase_dim=45)
    for n, a, b in zip(('mag_mel_log', 'real_mel', 'imag_mel', 'lf0'), cg[:4], cr[:4]):
        print('%-44s %s' % (tag + 'analysis_compressed ' + n, err(a, b)))
    for kw in (dict(b_out_hpf=False), dict(b_out_hpf=True), dict(b_out_hpf=False, b_const_rate=True),
               dict(b_out_hpf=False, per_phase_type='min_phase')):
        np.random.seed(3)
        y = mp.synthesis_from_compressed(*cr[:4], fs, **kw)
        np.random.seed(3)
        y_ref = orc.synthesis_from_compressed(*cr[:4], fs, **kw)
        print('%-44s %s   (peak %.2f)' % (tag + 'synthesis_from_compressed ' + ','.join('%s=%s' % kv for kv in kw.items() if kv[0] != 'b_out_hpf' or kv[1]),
                                         err(y, y_ref), float(np.abs(y_ref).max())))
    if fs == 48000:
        np.random.seed(3)
        y, ph = mp.griffin_lim(ref[0].copy(), ref[5], phase_init='min_phase', niters=8)
        np.random.seed(3)
        y_ref, ph_ref = orc.griffin_lim(ref[0].copy(), ref[5], phase_init='min_phase', niters=8)
        print('%-44s %s' % (tag + 'griffin_lim (8 iterations) signal', err(y, y_ref)))

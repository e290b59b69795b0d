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

# ---- natural speech (golden slices of the bundled recordings, generated from the real reference) and the band-limited
# synthetic stress case: the rows that guard every float32 / tensor-core decision ----
import os
from magphase_b200.synth import synth_utterance_band_limited
g = np.load(os.path.join('tests', 'golden', 'natural_48k.npz'))
for tag, name in (('a', 'hvd_593'), ('b', 'hvd_577')):
    sig = g[tag + '_sig_i16'].astype(np.float64) / 32768.0
    pm, voi = g[tag + '_pm'], g[tag + '_voi']
    got = mp.analysis_lossless_from_pm(sig, 48000, pm, voi)
    st = int(g['bin_step'])
    for n, a in zip(('mag', 'real', 'imag'), got[:3]):
        print('%-44s %s' % ('natural %s analysis_lossless %s (cols)' % (name, n), err(a[:, ::st], g['%s_%s_cols' % (tag, n)])))
    print('%-44s %s' % ('natural %s synthesis_from_lossless' % name, err(mp.synthesis_from_lossless(*got[:4], 48000), g[tag + '_syn'])))
    cg = mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    for n, a in zip(('mag_mel_log', 'real_mel', 'imag_mel'), cg[:3]):
        print('%-44s %s' % ('natural %s analysis_compressed %s' % (name, n), err(a, g['%s_%s' % (tag, n)])))
    c3 = mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=10, alpha_phase=0.0)
    print('%-44s %s' % ('natural %s analysis (tts dims) real_mel' % name, err(c3[1], g[tag + '_real_mel_tts'])))
    np.random.seed(int(g[tag + '_seed']))
    y = mp.synthesis_from_compressed(g[tag + '_mag_mel_log'], g[tag + '_real_mel'], g[tag + '_imag_mel'], g[tag + '_lf0'], 48000, b_out_hpf=False)
    print('%-44s %s' % ('natural %s synthesis_from_compressed' % name, err(y, g[tag + '_syn_compressed'])))
for floor in (-72.0, None):
    sig, pm, voi = synth_utterance_band_limited(31, fs=48000, dur_s=0.6, floor_db=floor)
    ref = orc.analysis_lossless_from_pm(sig, 48000, pm, voi)
    got = mp.analysis_lossless_from_pm(sig, 48000, pm, voi)
    t = 'band-limited (floor %s) ' % floor
    for n, a, b in zip(('mag', 'real', 'imag'), got[:3], ref[:3]):
        print('%-44s %s' % (t + 'lossless ' + n, err(a, b)))
    cr = orc.format_for_modelling(*ref[:4], 48000, mag_dim=60, phase_dim=45)
    cg = mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    for n, a, b in zip(('mag_mel_log', 'real_mel', 'imag_mel'), cg[:3], cr[:3]):
        print('%-44s %s' % (t + 'compressed ' + n, err(a, b)))

#!/usr/bin/env python
"""Generate the committed golden vectors from the REAL reference (oracle/_ref, the mechanically
py3-translated copy of /root/reference/src written by oracle/make_ref.py).

Runs only in the build container (needs /root/reference).  Output: tests/golden/*.npz, small on
purpose.  Inputs are stored next to the expected outputs so the GPU box (which has neither
/root/reference nor oracle/_ref) can replay them.

    python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import make_ref  # noqa: E402

sys.path.insert(0, make_ref.build())
warnings.simplefilter('ignore')
import magphase as mp  # noqa: E402  (the reference)
import libaudio as la  # noqa: E402

from magphase_b200.synth import synth_marks_for_wav, synth_utterance  # noqa: E402

REF_DATA = '/root/reference/demos/data_48k'
BIN_STEP = 32
FULL_ROWS = [0, 3, 17, 40, -1]


def lossless():
    fs = 48000
    sig, pm, voi = synth_utterance(7, fs=fs, dur_s=0.4)
    m_fft, v_shift = mp.analysis_with_del_comp_from_pm(sig, fs, pm)
    m_mag, m_real, m_imag, v_f0 = mp.compute_lossless_feats(m_fft, v_shift, voi, fs)
    y = mp.synthesis_from_lossless(m_mag.copy(), m_real.copy(), m_imag.copy(), v_f0.copy(), fs)
    minph = la.build_min_phase_from_mag_spec(m_mag[FULL_ROWS].copy())
    np.savez_compressed(
        os.path.join(HERE, 'lossless_synth48k.npz'),
        sig_i16=np.round(sig * 32768.0).astype(np.int16), pm=pm, voi=voi, fs=fs,
        v_shift=v_shift.astype(np.int64), v_f0=v_f0,
        full_rows=np.array(FULL_ROWS), bin_step=BIN_STEP,
        mag_rows=m_mag[FULL_ROWS], real_rows=m_real[FULL_ROWS], imag_rows=m_imag[FULL_ROWS],
        mag_cols=m_mag[:, ::BIN_STEP], real_cols=m_real[:, ::BIN_STEP], imag_cols=m_imag[:, ::BIN_STEP],
        syn=y, minph_rows=minph)
    print('lossless: %d frames, syn %d samples' % (v_shift.size, y.size))


def compressed():
    fs = 48000
    d = os.path.join(REF_DATA, 'params_predicted')
    n = 64
    rd = lambda ext, dim: np.fromfile(os.path.join(d, 'hvd_704' + ext), dtype=np.float32).reshape(-1, dim)[40:40 + n]
    mag, real, imag, lf0 = rd('.mag', 60), rd('.real', 45), rd('.imag', 45), rd('.lf0', 1)[:, 0]
    f64 = lambda a: a.astype(np.float64)
    out = {}
    for name, kw in (('var_nohpf', dict(b_out_hpf=False)), ('var_hpf', dict(b_out_hpf=True)),
                     ('const_nohpf', dict(b_out_hpf=False, b_const_rate=True)),
                     ('minph_nohpf', dict(b_out_hpf=False, per_phase_type='min_phase'))):
        np.random.seed(1234)
        out['syn_' + name] = mp.synthesis_from_compressed(f64(mag), f64(real), f64(imag), f64(lf0), fs, **kw)
    np.random.seed(1234)
    out['syn_16k_const_hpf'] = mp.synthesis_from_compressed(f64(mag), f64(real), f64(imag), f64(lf0), 16000,
                                                            b_const_rate=True)
    pf48 = mp.post_filter(f64(mag), 48000)
    pf16 = mp.post_filter(f64(mag), 16000)
    unw = la.sp_mel_unwarp(f64(mag)[:4], 2049, alpha=0.77, in_type='log')
    r_unw, i_unw = mp.phase_uncompress_type1_mcep(f64(real)[:4], f64(imag)[:4], 0.77, 4096, fs)
    np.savez_compressed(os.path.join(HERE, 'compressed_hvd704.npz'),
                        mag=mag, real=real, imag=imag, lf0=lf0, fs=fs, seed=1234,
                        post_filter_48k=pf48, post_filter_16k=pf16,
                        mag_unwarp4=unw, real_unwarp4=r_unw, imag_unwarp4=i_unw, **out)
    print('compressed:', {k: v.shape for k, v in out.items()})


def griffin_lim():
    """src/magphase.py:3318-3373 on the magnitudes of the lossless golden utterance: every phase_init, 4 iterations.
    Inputs (marks, voicing, signal) are the ones stored in lossless_synth48k.npz; the magnitudes are re-derived from them
    by the replaying test through the oracle (pinned against the reference to 1e-12)."""
    fs = 48000
    sig, pm, voi = synth_utterance(7, fs=fs, dur_s=0.4)
    m_fft, v_shift = mp.analysis_with_del_comp_from_pm(sig, fs, pm)
    m_mag, m_real, m_imag, v_f0 = mp.compute_lossless_feats(m_fft, v_shift, voi, fs)
    out = {}
    for init in ('linear', 'min_phase', 'random'):
        np.random.seed(4321)
        y, ph = mp.griffin_lim(m_mag.copy(), v_shift, phase_init=init, niters=4)
        out['syn_' + init] = y
        out['phase_rows_' + init] = ph[FULL_ROWS]
    np.savez_compressed(os.path.join(HERE, 'griffin_lim_synth48k.npz'), seed=4321, niters=4, full_rows=np.array(FULL_ROWS),
                        v_shift=v_shift.astype(np.int64), **out)
    print('griffin_lim:', {k: v.shape for k, v in out.items()})


def natural():
    """BASELINE config 1 names a bundled natural recording (demos/demo_copy_synthesis_lossless.py:57-91 on
    demos/data_48k/wavs_nat/hvd_593.wav).  REAPER is not available, so the marks come from synth_marks_for_wav (seeded);
    everything downstream of the marks is the real reference: analysis_with_del_comp_from_pm (src/magphase.py:266-334),
    compute_lossless_feats (:457-476), synthesis_from_lossless (:1759-1776), synthesis_from_compressed (:825-997).  The
    compressed features are the oracle's format_for_modelling of the reference's lossless features (SPTK mcep restated,
    unpinned).  Half-second slices (24,000 int16 samples each) of two recordings: silence -> onset -> vowel, and a
    stretch with quiet high-frequency bins -- the hard case for float32 butterflies."""
    import magphase_oracle as orc
    from scipy.io import wavfile
    fs = 48000
    out = {}
    for tag, name, a, seed in (('a', 'hvd_593', 12000, 593), ('b', 'hvd_577', 48000, 577)):
        fs_w, x = wavfile.read(os.path.join(REF_DATA, 'wavs_nat', name + '.wav'))
        assert fs_w == fs and x.dtype == np.int16
        x = x[a:a + 24000]
        sig = x.astype(np.float64) / 32768.0                       # sf.read of PCM16
        pm, voi = synth_marks_for_wav(sig.size, fs=fs, seed=seed)
        m_fft, v_shift = mp.analysis_with_del_comp_from_pm(sig, fs, pm)
        m_mag, m_real, m_imag, v_f0 = mp.compute_lossless_feats(m_fft, v_shift, voi, fs)
        y = mp.synthesis_from_lossless(m_mag.copy(), m_real.copy(), m_imag.copy(), v_f0.copy(), fs)
        mm, rr, ii, lf0 = orc.format_for_modelling(m_mag, m_real, m_imag, v_f0, fs, mag_dim=60, phase_dim=45)
        mm3, rr3, ii3, _ = orc.format_for_modelling(m_mag, m_real, m_imag, v_f0, fs, mag_dim=60, phase_dim=10, alpha_phase=0.0)
        np.random.seed(2000 + seed)
        yc = mp.synthesis_from_compressed(mm.copy(), rr.copy(), ii.copy(), lf0.copy(), fs, b_out_hpf=False)
        rows = [0, 5, 23, v_shift.size // 2, -1]
        out.update({tag + '_' + k: v for k, v in dict(
            sig_i16=x, pm=pm, voi=voi, v_shift=v_shift.astype(np.int64), v_f0=v_f0, full_rows=np.array(rows),
            mag_rows=m_mag[rows], real_rows=m_real[rows], imag_rows=m_imag[rows],
            mag_cols=m_mag[:, ::BIN_STEP], real_cols=m_real[:, ::BIN_STEP], imag_cols=m_imag[:, ::BIN_STEP],
            syn=y, mag_mel_log=mm, real_mel=rr, imag_mel=ii, lf0=lf0,
            real_mel_tts=rr3, imag_mel_tts=ii3, seed=2000 + seed, syn_compressed=yc).items()})
        print('natural %s: %d frames (%d voiced), syn %d, compressed syn %d samples' % (
            name, v_shift.size, int(voi.sum()), y.size, yc.size))
    np.savez_compressed(os.path.join(HERE, 'natural_48k.npz'), fs=fs, bin_step=BIN_STEP, **out)


if __name__ == '__main__':
    only = sys.argv[1:]
    for fn in (lossless, compressed, griffin_lim, natural):
        if not only or fn.__name__ in only:
            fn()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))

#!/usr/bin/env python
"""Copy synthesis with magphase_b200: the two demos of the reference (demos/demo_copy_synthesis_lossless.py and
demos/demo_copy_synthesis_low_dim.py) in one script, on the GPU.

    python demos/demo_copy_synthesis.py                       # synthetic 48 kHz utterance (no files needed)
    python demos/demo_copy_synthesis.py --wav hvd_593.wav --est hvd_593.est [--mode low_dim] [--mag-dim 100]

Pitch marks come from a REAPER .est file (``--est``) or from the REAPER binary when it is installed, exactly like the
reference (src/libaudio.py:450-455); without ``--wav`` the script writes a synthetic utterance + .est first.
Outputs go to ``--out-dir`` (default ./wavs_syn): <token>_copy_syn_lossless.wav / <token>_copy_syn_low_dim.wav.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import magphase_b200.magphase as mp  # noqa: E402
from magphase_b200 import hostio  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--wav', default=None)
    ap.add_argument('--est', default=None)
    ap.add_argument('--out-dir', default='wavs_syn')
    ap.add_argument('--mode', default='both', choices=['lossless', 'low_dim', 'both'])
    ap.add_argument('--mag-dim', type=int, default=100)       # demos/demo_copy_synthesis_low_dim.py:63
    ap.add_argument('--phase-dim', type=int, default=45)
    ap.add_argument('--const-rate', action='store_true')
    ap.add_argument('--seed', type=int, default=0, help='np.random.seed for the aperiodic noise of the low-dim synthesis')
    a = ap.parse_args(argv)
    os.makedirs(a.out_dir, exist_ok=True)
    wav, est = a.wav, a.est
    if wav is None:
        from scipy.io import wavfile
        from magphase_b200.synth import synth_utterance
        sig, pm, voi = synth_utterance(593, fs=48000, dur_s=2.4)
        wav, est = os.path.join(a.out_dir, 'synth_593.wav'), os.path.join(a.out_dir, 'synth_593.est')
        wavfile.write(wav, 48000, np.round(sig * 32768.0).astype(np.int16))
        hostio.write_reaper_est_file(est, pm / 48000.0, voi)
        print('no --wav given: wrote a synthetic utterance to %s (+ .est)' % wav)
    token = os.path.splitext(os.path.basename(wav))[0]
    done = []
    if a.mode in ('lossless', 'both'):
        print('Analysing (lossless)..........................................')
        m_mag, m_real, m_imag, v_f0, fs, v_shift = mp.analysis_lossless(wav, est_file=est)
        print('Synthesising...................................................')
        v_syn = mp.synthesis_from_lossless(m_mag, m_real, m_imag, v_f0, fs)
        out = os.path.join(a.out_dir, token + '_copy_syn_lossless.wav')
        hostio.write_audio_file(out, v_syn, fs)
        done.append((out, m_mag.shape[0]))
    if a.mode in ('low_dim', 'both'):
        print('Analysing (mag_dim=%d, phase_dim=%d)...........................' % (a.mag_dim, a.phase_dim))
        m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0, v_shift, fs, fft_len = mp.analysis_compressed(
            wav, mag_dim=a.mag_dim, phase_dim=a.phase_dim, b_const_rate=a.const_rate, est_file=est)
        print('Synthesising...................................................')
        np.random.seed(a.seed)
        v_syn = mp.synthesis_from_compressed(m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0, fs,
                                             b_const_rate=a.const_rate, b_out_hpf=False)
        out = os.path.join(a.out_dir, token + '_copy_syn_low_dim.wav')
        hostio.write_audio_file(out, v_syn, fs)
        done.append((out, m_mag_mel_log.shape[0]))
    for out, n in done:
        print('%s  (%d frames)' % (out, n))
    print('Done!')
    return done


if __name__ == '__main__':
    main()

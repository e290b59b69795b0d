"""End-to-end file pipeline at a few hundred utterances (run on the GPU box): synthetic 48 kHz wavs + REAPER-style .est
files on local disk -> magphase_b200.batch.run_feature_extraction -> run_waveform_generation.  Prints frames/s
including file IO.    python profiles/file_pipeline_demo.py [n_utts] [dur_s]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, '.')
import numpy as np
from scipy.io import wavfile

from magphase_b200 import batch, hostio
from magphase_b200.synth import synth_utterance

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dur = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
root = tempfile.mkdtemp(prefix='mpb_files_')
wav_dir, est_dir = os.path.join(root, 'wavs'), os.path.join(root, 'est')
os.makedirs(wav_dir); os.makedirs(est_dir)
base = [synth_utterance(u, fs=48000, dur_s=dur) for u in range(16)]
tokens = []
for u in range(n):
    sig, pm, voi = base[u % 16]
    tok = 'utt_%04d' % u
    wavfile.write(os.path.join(wav_dir, tok + '.wav'), 48000, np.round(sig * 32768.0).astype(np.int16))
    hostio.write_reaper_est_file(os.path.join(est_dir, tok + '.est'), pm / 48000.0, voi)
    tokens.append(tok)
print('corpus: %d utterances x %.1f s in %s' % (n, dur, root))
for rep in range(2):                       # the second pass has warm page cache, plans and pinned pools
    feats, syn = os.path.join(root, 'feats%d' % rep), os.path.join(root, 'syn%d' % rep)
    t = time.perf_counter()
    r1 = batch.run_feature_extraction(tokens, wav_dir, feats, est_dir=est_dir, batch_utts=64, io_threads=8)
    t1 = time.perf_counter()
    np.random.seed(1)
    r2 = batch.run_waveform_generation(tokens, feats, syn, 60, 45, 48000, pf_type='magphase', batch_utts=64, io_threads=8)
    t2 = time.perf_counter()
    print('pass %d: extraction %.0f frames/s (%.2f s), generation %.0f frames/s (%.2f s), %d frames'
          % (rep, r1['frames'] / (t1 - t), t1 - t, r2['frames'] / (t2 - t1), t2 - t1, r1['frames']))

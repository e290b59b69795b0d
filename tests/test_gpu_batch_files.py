"""The batch drivers (magphase_b200/batch.py = the reference's scripts/batch_feature_extraction_for_tts.py and
scripts/batch_waveform_generation.py on the GPU) against the per-utterance file wrappers: feature files byte for byte,
waveforms from the same NumPy noise stream consumed in list order (the reference's sequential loop)."""
import os
import warnings

import numpy as np
import pytest

from magphase_b200.synth import synth_utterance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def corpus(tmp_path_factory):
    from scipy.io import wavfile
    from magphase_b200 import hostio
    d = tmp_path_factory.mktemp('corpus')
    wav_dir, est_dir = d / 'wavs', d / 'est'
    os.makedirs(wav_dir); os.makedirs(est_dir)
    tokens = []
    for u, dur in enumerate((0.45, 0.8, 0.3, 0.65, 0.5)):
        sig, pm, voi = synth_utterance(60 + u, fs=48000, dur_s=dur)
        tok = 'utt_%02d' % u
        wavfile.write(str(wav_dir / (tok + '.wav')), 48000, np.round(sig * 32768.0).astype(np.int16))
        hostio.write_reaper_est_file(str(est_dir / (tok + '.est')), pm / 48000.0, voi)
        tokens.append(tok)
    scp = d / 'file_id.scp'
    scp.write_text('# tokens\n' + '\n'.join(tokens) + '\n')
    return d, str(scp), str(wav_dir), str(est_dir), tokens


def test_batch_feature_extraction_writes_the_same_files(corpus):
    import magphase_b200.magphase as mp
    from magphase_b200 import batch
    d, scp, wav_dir, est_dir, tokens = corpus
    out_b, out_s = str(d / 'feats_batch'), str(d / 'feats_single')
    os.makedirs(out_s)
    r = batch.run_feature_extraction(scp, wav_dir, out_b, est_dir=est_dir, batch_utts=2, io_threads=3)   # 3 batches
    assert r['utterances'] == len(tokens) and r['frames'] > 0
    for tok in tokens:
        mp.analysis_for_acoustic_modelling(os.path.join(wav_dir, tok + '.wav'), out_s, mag_dim=60, phase_dim=45,
                                           est_file=os.path.join(est_dir, tok + '.est'))
        for ext in ('.mag', '.real', '.imag', '.lf0', '.shift'):
            a = open(os.path.join(out_b, tok + ext), 'rb').read()
            b = open(os.path.join(out_s, tok + ext), 'rb').read()
            assert a == b, (tok, ext)


def test_batch_waveform_generation_matches_the_sequential_loop(corpus):
    import magphase_b200.magphase as mp
    from magphase_b200 import batch, hostio
    d, scp, wav_dir, est_dir, tokens = corpus
    feats = str(d / 'feats_batch')
    if not os.path.isdir(feats):
        batch.run_feature_extraction(scp, wav_dir, feats, est_dir=est_dir)
    out_b, out_s = str(d / 'syn_batch'), str(d / 'syn_single')
    os.makedirs(out_s)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        np.random.seed(5)
        r = batch.run_waveform_generation(scp, feats, out_b, 60, 45, 48000, pf_type='magphase', batch_utts=2, io_threads=3)
        state_b = np.random.get_state()
        np.random.seed(5)
        for tok in tokens:
            mp.synthesis_from_acoustic_modelling(feats, tok, out_s, 60, 45, 48000, pf_type='magphase')
        state_s = np.random.get_state()
    assert r['utterances'] == len(tokens)
    assert state_b[2] == state_s[2] and np.array_equal(state_b[1], state_s[1])       # same draws from the stream
    for tok in tokens:
        a, fa = hostio.read_audio_file(os.path.join(out_b, tok + '.wav'))
        b, fb = hostio.read_audio_file(os.path.join(out_s, tok + '.wav'))
        assert fa == fb == 48000 and a.shape == b.shape
        assert np.max(np.abs(a - b)) <= 1.0 / 32768 + 1e-12                          # at most one PCM16 step
        assert np.mean(a != b) < 0.01


def test_batch_cli_and_errors(corpus, capsys):
    from magphase_b200 import batch
    d, scp, wav_dir, est_dir, tokens = corpus
    out = str(d / 'feats_cli')
    batch.main(['extract', '--scp', scp, '--wav-dir', wav_dir, '--out-dir', out, '--est-dir', est_dir, '--batch-utts', '8'])
    assert 'Done!' in capsys.readouterr().out
    assert sorted(os.listdir(out)) == sorted(t + e for t in tokens for e in ('.mag', '.real', '.imag', '.lf0', '.shift'))
    np.random.seed(3)
    r = batch.run_waveform_generation(tokens, out, str(d / 'x'), 60, 45, 48000, pf_type='merlin')       # Merlin-style post-filter
    assert r['utterances'] == len(tokens) and sorted(os.listdir(str(d / 'x'))) == sorted(t + '.wav' for t in tokens)
    with pytest.raises(FileNotFoundError):
        batch.run_feature_extraction(['missing'], wav_dir, str(d / 'y'), est_dir=est_dir)


def test_resume_and_skip_errors_on_the_device(corpus):
    """resume / on_error='skip' with the real device calls (the CPU suite covers the logic with stand-ins): a resumed
    generation run writes the files of an uninterrupted one because the finished tokens' noise draws are still consumed;
    a one-frame feature set fails the argument checks of its batch and is isolated."""
    from magphase_b200 import batch, hostio
    d, scp, wav_dir, est_dir, tokens = corpus
    feats = str(d / 'feats_resume')
    r = batch.run_feature_extraction(tokens[:2], wav_dir, feats, est_dir=est_dir)
    first = {t: open(os.path.join(feats, t + '.mag'), 'rb').read() for t in tokens[:2]}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        r = batch.run_feature_extraction(tokens + ['absent'], wav_dir, feats, est_dir=est_dir, batch_utts=4, resume=True,
                                         on_error='skip')
    assert r['skipped'] == tokens[:2] and r['failed'] == ['absent'] and r['utterances'] == len(tokens) - 2
    assert all(open(os.path.join(feats, t + '.mag'), 'rb').read() == first[t] for t in tokens[:2])
    ref = str(d / 'feats_batch')
    if os.path.isdir(ref):
        for t in tokens:
            assert open(os.path.join(feats, t + '.real'), 'rb').read() == open(os.path.join(ref, t + '.real'), 'rb').read()
    for ext, dim in (('.mag', 60), ('.real', 45), ('.imag', 45), ('.lf0', 1)):
        hostio.write_binfile(np.zeros((1, dim)), os.path.join(feats, 'short' + ext))
    full, part = str(d / 'syn_full'), str(d / 'syn_part')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        np.random.seed(21)
        batch.run_waveform_generation(tokens, feats, full, 60, 45, 48000, batch_utts=3)
        state_full = np.random.get_state()
        np.random.seed(21)
        batch.run_waveform_generation(tokens[:4], feats, part, 60, 45, 48000, batch_utts=3)
        os.remove(os.path.join(part, tokens[0] + '.wav'))
        os.remove(os.path.join(part, tokens[2] + '.wav'))
        np.random.seed(21)
        r = batch.run_waveform_generation(tokens[:3] + ['short'] + tokens[3:], feats, part, 60, 45, 48000, batch_utts=3,
                                          resume=True, on_error='skip')
    assert r['skipped'] == [tokens[1], tokens[3]] and r['failed'] == ['short'] and r['utterances'] == 3
    st = np.random.get_state()
    assert st[2] == state_full[2] and np.array_equal(st[1], state_full[1])
    for t in tokens:
        a, _ = hostio.read_audio_file(os.path.join(part, t + '.wav'))
        b, _ = hostio.read_audio_file(os.path.join(full, t + '.wav'))
        assert a.shape == b.shape and np.max(np.abs(a - b)) <= 1.0 / 32768 + 1e-12 and np.mean(a != b) < 0.01, t

"""Host logic of the batch drivers that needs no GPU: token lists, batching, argument errors raised before any device work."""
import pytest

from magphase_b200 import batch


def test_read_tokens_and_batches(tmp_path):
    scp = tmp_path / 'file_id.scp'
    scp.write_text('# list of tokens\nhvd_593\n\n  hvd_594  # trailing comment\n#hvd_595\nhvd_596\n')
    assert batch.read_tokens(str(scp)) == ['hvd_593', 'hvd_594', 'hvd_596']
    assert list(batch._batches(list(range(7)), 3)) == [[0, 1, 2], [3, 4, 5], [6]]
    assert list(batch._batches([], 3)) == []


def test_argument_errors_come_before_device_work(tmp_path):
    with pytest.raises(ValueError):
        batch.run_waveform_generation(['a'], str(tmp_path), str(tmp_path / 'o'), 60, 45, 48000, pf_type='bogus')
    # an empty token list is a no-op on both sides
    assert batch.run_feature_extraction([], str(tmp_path), str(tmp_path / 'f'))['utterances'] == 0
    assert batch.run_waveform_generation([], str(tmp_path), str(tmp_path / 'w'), 60, 45, 48000)['frames'] == 0


def test_cli_parser():
    with pytest.raises(SystemExit):
        batch.main(['extract'])                      # missing required arguments


# ------------------------------------------------------------------------------------------------------------------
# resume / fault isolation (failure detection, checkpoint-resume).  The device calls are replaced by
# CPU stand-ins here (the logic under test is the drivers'); tests/test_gpu_batch_files.py repeats them on the GPU.
# ------------------------------------------------------------------------------------------------------------------
import os
import warnings

import numpy as np

import magphase_oracle as orc
from magphase_b200 import hostio
from magphase_b200.synth import synth_utterance


def _corpus(d, n=5):
    from scipy.io import wavfile
    wav_dir, est_dir = d / 'wavs', d / 'est'
    os.makedirs(wav_dir); os.makedirs(est_dir)
    tokens = []
    for u in range(n):
        sig, pm, voi = synth_utterance(70 + u, fs=48000, dur_s=0.25 + 0.05 * u)
        tok = 'utt_%02d' % u
        wavfile.write(str(wav_dir / (tok + '.wav')), 48000, np.round(sig * 32768.0).astype(np.int16))
        hostio.write_reaper_est_file(str(est_dir / (tok + '.est')), pm / 48000.0, voi)
        tokens.append(tok)
    return str(wav_dir), str(est_dir), tokens


def _fake_analysis(calls):
    def f(l_sig, fs, l_pm, l_voi, **kw):
        calls.append(len(l_sig))
        outs = []
        for sig, pm, voi in zip(l_sig, l_pm, l_voi):
            n = np.size(pm)
            if n < 3:
                raise ValueError('too few pitch marks')
            r = np.random.RandomState(n)
            lf0 = np.where(np.asarray(voi) > 0, np.log(120.0), -1e10)
            outs.append((r.rand(n, 60), r.rand(n, 45), r.rand(n, 45), lf0, np.diff(np.hstack((0, np.round(pm)))).astype(int)))
        return outs
    return f


def test_extraction_resume_and_skip_errors(tmp_path, monkeypatch):
    wav_dir, est_dir, tokens = _corpus(tmp_path)
    calls = []
    monkeypatch.setattr(batch.mp, 'analysis_compressed_batch', _fake_analysis(calls))
    out = str(tmp_path / 'feats')
    exts = ('.mag', '.real', '.imag', '.lf0', '.shift')
    # a first job dies after two tokens ...
    r = batch.run_feature_extraction(tokens[:2], wav_dir, out, est_dir=est_dir, batch_utts=2)
    assert r['utterances'] == 2 and r['skipped'] == [] and r['failed'] == []
    before = {t + e: open(os.path.join(out, t + e), 'rb').read() for t in tokens[:2] for e in exts}
    os.remove(os.path.join(out, tokens[1] + '.imag'))                 # ... and one of its tokens is incomplete
    open(os.path.join(out, tokens[0] + '.mag'), 'ab').close()
    # the resumed job redoes exactly the incomplete and the missing ones; a broken wav and a token with a one-line .est
    # are listed, not fatal
    open(os.path.join(wav_dir, 'broken.wav'), 'wb').write(b'not a wav file')
    open(os.path.join(est_dir, 'broken.est'), 'w').write('EST_File Track\nEST_Header_End\n')
    del calls[:]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        r = batch.run_feature_extraction(tokens[:3] + ['broken', 'nofile'] + tokens[3:], wav_dir, out, est_dir=est_dir,
                                         batch_utts=3, resume=True, on_error='skip')
    assert r['skipped'] == [tokens[0]]
    assert r['failed'] == ['broken', 'nofile']
    assert r['utterances'] == len(tokens) - 1
    assert any('broken' in str(x.message) for x in w)
    assert sum(calls) == len(tokens) - 1
    lists = [f for f in os.listdir(out) if f.startswith('crash_file_list_')]
    assert len(lists) == 1 and lists[0].endswith('_%d.scp' % os.getpid())
    assert batch.read_tokens(os.path.join(out, lists[0])) == ['broken', 'nofile']
    for t in tokens:
        for e in exts:
            assert os.path.getsize(os.path.join(out, t + e)) > 0
    assert not [f for f in os.listdir(out) if f.startswith('.tmp')]               # no temporary files left behind
    after = {k: open(os.path.join(out, k), 'rb').read() for k in before}
    assert after == before                                                        # redone token: the same bytes
    # nothing left to do
    del calls[:]
    r = batch.run_feature_extraction(tokens, wav_dir, out, est_dir=est_dir, resume=True)
    assert r['utterances'] == 0 and r['skipped'] == tokens and calls == []
    # without on_error='skip' the first failure is raised
    with pytest.raises(Exception):
        batch.run_feature_extraction(['broken'], wav_dir, str(tmp_path / 'z'), est_dir=est_dir)
    with pytest.raises(ValueError):
        batch.run_feature_extraction(tokens, wav_dir, out, on_error='ignore')


def test_bad_marks_inside_a_batch_are_isolated(tmp_path, monkeypatch):
    """An utterance that fails the argument checks of the batched call is found by retrying one by one."""
    wav_dir, est_dir, tokens = _corpus(tmp_path, n=3)
    hostio.write_reaper_est_file(os.path.join(est_dir, tokens[1] + '.est'), np.array([0.01, 0.02]), np.array([1.0, 1.0]))
    calls = []
    monkeypatch.setattr(batch.mp, 'analysis_compressed_batch', _fake_analysis(calls))
    out = str(tmp_path / 'feats')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        r = batch.run_feature_extraction(tokens, wav_dir, out, est_dir=est_dir, on_error='skip')
    assert r['failed'] == [tokens[1]] and r['utterances'] == 2
    assert calls == [3, 1, 1, 1]
    assert os.path.exists(os.path.join(out, tokens[2] + '.mag')) and not os.path.exists(os.path.join(out, tokens[1] + '.mag'))
    with pytest.raises(ValueError):
        batch.run_feature_extraction(tokens, wav_dir, str(tmp_path / 'f2'), est_dir=est_dir)


def _feature_files(d, tokens, seed=0):
    os.makedirs(d, exist_ok=True)
    r = np.random.RandomState(seed)
    for k, t in enumerate(tokens):
        n = 20 + 3 * k
        lf0 = np.where(r.rand(n) > 0.4, np.log(r.uniform(90, 250, n)), -1e10)
        hostio.write_binfile(r.randn(n, 60), os.path.join(d, t + '.mag'))
        hostio.write_binfile(r.uniform(-1, 1, (n, 45)), os.path.join(d, t + '.real'))
        hostio.write_binfile(r.uniform(-1, 1, (n, 45)), os.path.join(d, t + '.imag'))
        hostio.write_binfile(lf0, os.path.join(d, t + '.lf0'))


def _fake_synthesis(l_feats, fs, fft_len=None, b_const_rate=False, **kw):
    """Draws what the real call draws (one np.random.uniform(-1, 1, ns_len) per utterance, in list order) and returns it."""
    ys = []
    for f in l_feats:
        if np.shape(f[0])[0] < 2:
            raise IndexError('needs at least two frames')
        ys.append(0.5 * np.random.uniform(-1, 1, batch._noise_draws(f[3], fs, fft_len, b_const_rate)))
    return ys


def test_noise_draw_count_matches_the_reference_algorithm(monkeypatch):
    """_noise_draws (what a resumed run consumes for a finished token) == the size of the np.random.uniform call inside
    synthesis_from_compressed (src/magphase.py:879-883), here in the oracle's restatement of it."""
    r = np.random.RandomState(1)
    n = 40
    lf0 = np.where(r.rand(n) > 0.4, np.log(r.uniform(90, 250, n)), -1e10)
    sizes = []
    real_uniform = np.random.uniform
    monkeypatch.setattr(orc.np.random, 'uniform', lambda lo, hi, size: (sizes.append(int(size)), real_uniform(lo, hi, size))[1])
    orc.synthesis_from_compressed(r.randn(n, 60) * 0.1, r.uniform(-1, 1, (n, 45)), r.uniform(-1, 1, (n, 45)), lf0, 48000,
                                  b_out_hpf=False)
    monkeypatch.undo()
    assert sizes == [batch._noise_draws(lf0, 48000, None, False)]


def test_generation_resume_writes_the_files_of_an_uninterrupted_run(tmp_path, monkeypatch):
    monkeypatch.setattr(batch.mp, 'synthesis_from_compressed_batch', _fake_synthesis)
    tokens = ['g%02d' % k for k in range(7)]
    feats = str(tmp_path / 'feats')
    _feature_files(feats, tokens)
    hostio.write_binfile(np.zeros((1, 60)), os.path.join(feats, 'short.mag'))      # one frame: fails the argument checks
    hostio.write_binfile(np.zeros((1, 45)), os.path.join(feats, 'short.real'))
    hostio.write_binfile(np.zeros((1, 45)), os.path.join(feats, 'short.imag'))
    hostio.write_binfile(np.zeros(1), os.path.join(feats, 'short.lf0'))
    full, part = str(tmp_path / 'full'), str(tmp_path / 'part')
    np.random.seed(11)
    r = batch.run_waveform_generation(tokens, feats, full, 60, 45, 48000, pf_type='no', batch_utts=3)
    state_full = np.random.get_state()
    assert r['utterances'] == 7 and r['skipped'] == [] and r['failed'] == []
    # an interrupted job left tokens 0, 1, 4 behind (same seed); the resumed one (same seed again) fills in the rest
    np.random.seed(11)
    batch.run_waveform_generation(tokens[:2], feats, part, 60, 45, 48000, pf_type='no', batch_utts=3)
    os.replace(os.path.join(part, tokens[1] + '.wav'), os.path.join(part, 'keep.wav'))
    np.random.seed(11)
    batch.run_waveform_generation(tokens[:5], feats, str(tmp_path / 'tmp5'), 60, 45, 48000, pf_type='no', batch_utts=3)
    os.replace(os.path.join(str(tmp_path / 'tmp5'), tokens[4] + '.wav'), os.path.join(part, tokens[4] + '.wav'))
    os.replace(os.path.join(part, 'keep.wav'), os.path.join(part, tokens[1] + '.wav'))
    np.random.seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        r = batch.run_waveform_generation(tokens[:3] + ['short', 'absent'] + tokens[3:], feats, part, 60, 45, 48000,
                                          pf_type='no', batch_utts=3, resume=True, on_error='skip')
    assert r['skipped'] == [tokens[0], tokens[1], tokens[4]]
    assert r['failed'] == ['absent', 'short'] or r['failed'] == ['short', 'absent']
    assert r['utterances'] == 4
    st = np.random.get_state()
    assert st[2] == state_full[2] and np.array_equal(st[1], state_full[1])         # the same draws were consumed
    for t in tokens:
        assert open(os.path.join(part, t + '.wav'), 'rb').read() == open(os.path.join(full, t + '.wav'), 'rb').read(), t
    assert not [f for f in os.listdir(part) if f.startswith('.tmp')]


def test_cli_passes_resume_and_skip_errors(monkeypatch, capsys):
    seen = {}

    def fake_extract(scp, wav_dir, out_dir, **kw):
        seen['extract'] = kw
        return dict(utterances=3, frames=30, seconds=1.0, skipped=['a'], failed=[])

    def fake_generate(scp, feats_dir, out_dir, mag_dim, phase_dim, fs, **kw):
        seen['generate'] = kw
        return dict(utterances=2, frames=20, seconds=1.0, skipped=[], failed=['b'])
    monkeypatch.setattr(batch, 'run_feature_extraction', fake_extract)
    monkeypatch.setattr(batch, 'run_waveform_generation', fake_generate)
    batch.main(['extract', '--scp', 's', '--wav-dir', 'w', '--out-dir', 'o', '--resume', '--skip-errors'])
    assert seen['extract']['resume'] is True and seen['extract']['on_error'] == 'skip'
    batch.main(['generate', '--scp', 's', '--feats-dir', 'f', '--out-dir', 'o'])
    assert seen['generate']['resume'] is False and seen['generate']['on_error'] == 'raise'
    out = capsys.readouterr().out
    assert '1 tokens already done, 0 failed' in out and '0 tokens already done, 1 failed' in out and out.count('Done!') == 2

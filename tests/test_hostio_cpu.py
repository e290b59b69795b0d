"""File formats either side of the path (host only): REAPER .est, raw float32 feature files, PCM wav."""
import numpy as np
import pytest

from magphase_b200 import hostio


def _est(path, header, rows):
    path.write_text('\n'.join(header) + '\n' + '\n'.join('%.6f %d' % r for r in rows) + '\n')
    return str(path)


STD = ['EST_File Track', 'DataType ascii', 'NumFrames 4', 'NumChannels 1', 'FrameShift 0.00000', 'VoicingEnabled true',
       'EST_Header_End']
ROWS = [(0.005, 0), (0.010, 1), (0.010, 1), (0.0175, 1), (0.030, 0)]


def test_est_reader_standard_header_and_filters(tmp_path):
    t, v = hostio.read_reaper_est_file(_est(tmp_path / 'a.est', STD, ROWS))
    assert np.allclose(t, [0.005, 0.010, 0.0175, 0.030]) and v.tolist() == [0, 1, 1, 0]      # repeated time dropped
    # marks at or beyond the last sample are dropped when the signal length is given (src/libaudio.py:440-445)
    t2, v2 = hostio.read_reaper_est_file(_est(tmp_path / 'b.est', STD, ROWS), check_len_smpls=1000, fs=48000)
    assert np.allclose(t2, [0.005, 0.010, 0.0175])
    with pytest.raises(ValueError):
        hostio.read_reaper_est_file(str(tmp_path / 'b.est'), check_len_smpls=1000)


def test_est_reader_follows_the_header_end_marker(tmp_path):
    longer = STD[:-1] + ['BreaksPresent true', 'CommentChar ;', 'EST_Header_End']
    t, v = hostio.read_reaper_est_file(_est(tmp_path / 'c.est', longer, ROWS))
    assert t.size == 4 and v.tolist() == [0, 1, 1, 0]
    with pytest.raises(ValueError):
        hostio.read_reaper_est_file(_est(tmp_path / 'd.est', STD, []))


def test_binfile_and_wav_round_trip(tmp_path):
    m = np.random.default_rng(0).normal(size=(7, 60))
    f = str(tmp_path / 'x.mag')
    hostio.write_binfile(m, f)
    back = hostio.read_binfile(f, dim=60)
    assert back.dtype == np.float64 and np.array_equal(back, m.astype(np.float32).astype(np.float64))
    with pytest.raises(ValueError):
        hostio.read_binfile(f, dim=7 * 60 - 1)
    sig = 0.5 * np.sin(np.arange(4800) * 0.05)
    w = str(tmp_path / 'y.wav')
    hostio.write_audio_file(w, sig, 48000)
    y, fs = hostio.read_audio_file(w)
    assert fs == 48000 and y.shape == sig.shape
    assert abs(np.max(np.abs(y)) - 0.98) < 1e-3                                  # peak-normalised to 0.98
    assert np.max(np.abs(y - 0.98 * sig / np.max(np.abs(sig)))) <= 1.0 / 32768 + 1e-12


def test_band_limited_synthetic_utterance():
    """synth_utterance_band_limited: same marks as synth_utterance, int16-exact samples, high band at the faint floor."""
    from magphase_b200.synth import synth_utterance, synth_utterance_band_limited
    a = synth_utterance(3, 48000, 0.5)
    b = synth_utterance_band_limited(3, 48000, 0.5)
    b2 = synth_utterance_band_limited(3, 48000, 0.5)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(b[0], b2[0])
    assert np.array_equal(np.round(b[0] * 32768.0), b[0] * 32768.0)
    spec = np.abs(np.fft.rfft(b[0] * np.hanning(b[0].size)))
    f = np.fft.rfftfreq(b[0].size, 1.0 / 48000)
    lo, hi = spec[(f > 300) & (f < 3000)].max(), spec[f > 10000].max()
    assert 20 * np.log10(hi / lo) < -50.0


def test_reaper_shell_out_with_a_fake_binary(tmp_path, monkeypatch):
    """The REAPER call of analysis_lossless (src/libaudio.py:450-455) with a stand-in executable: same flags, file names
    with blanks survive (argument list, no shell), the temporary .est is removed, a failing binary raises instead of leaving
    a FileNotFoundError from the clean-up."""
    import os
    import stat
    import numpy as np
    import pytest
    import magphase_b200.magphase as mp
    from magphase_b200 import hostio
    log = tmp_path / 'args.txt'
    fake = tmp_path / 'reaper'
    fake.write_text('#!/bin/bash\nprintf "%s\\n" "$@" > "' + str(log) + '"\n'
                    'while [ $# -gt 0 ]; do if [ "$1" = "-p" ]; then out="$2"; fi; shift; done\n'
                    'printf "EST_File Track\\nDataType ascii\\nNumFrames 3\\nNumChannels 1\\nFrameShift 0.0\\nVoicingEnabled true\\n'
                    'EST_Header_End\\n0.010000 1\\n0.020000 0\\n0.500000 1\\n" > "$out"\n')
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setattr(hostio, 'find_tool', lambda name: str(fake))
    monkeypatch.chdir(tmp_path)
    wav = str(tmp_path / 'my utterance.wav')
    v_pm, v_voi = mp.get_pitch_marks_and_voicing(wav, 4800, 48000)          # 0.5 s lies beyond the 0.1 s signal: dropped
    assert np.allclose(v_pm, [0.01, 0.02]) and v_voi.tolist() == [1.0, 0.0]
    args = log.read_text().split('\n')
    assert args[:9] == ['-s', '-x', '400', '-m', '50', '-a', '-u', '0.005', '-i'] and args[9] == wav and args[10] == '-p'
    assert not [f for f in os.listdir(tmp_path) if f.startswith('temp_') and f.endswith('.est')]
    fake.write_text('#!/bin/bash\nexit 3\n')
    with pytest.raises(RuntimeError):
        mp.get_pitch_marks_and_voicing(wav, 4800, 48000)

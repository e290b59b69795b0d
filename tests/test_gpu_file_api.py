"""The file-level wrappers of the reference API (wav in / feature files / wav out) on top of the kernels:
analysis_lossless (A8), analysis_for_acoustic_modelling incl. its alpha_phase=False quirk (C5),
synthesis_from_acoustic_modelling (S16)."""
import os
import warnings

import numpy as np
import pytest

import magphase_oracle as orc
from magphase_b200.synth import synth_utterance

pytestmark = pytest.mark.gpu


def rms(a, b):
    return float(np.sqrt(np.mean(np.abs(np.asarray(a) - np.asarray(b)) ** 2)))


@pytest.fixture(scope='module')
def files(tmp_path_factory):
    from scipy.io import wavfile
    from magphase_b200 import hostio
    d = tmp_path_factory.mktemp('mpb')
    sig, pm, voi = synth_utterance(31, fs=48000, dur_s=0.6)
    wav = str(d / 'utt_a.wav')
    wavfile.write(wav, 48000, np.round(sig * 32768.0).astype(np.int16))
    est = str(d / 'utt_a.est')
    hostio.write_reaper_est_file(est, pm / 48000.0, voi)
    return d, wav, est, sig, pm, voi


def test_analysis_lossless_from_wav_and_est(files):
    import magphase_b200.magphase as mp
    from magphase_b200 import hostio
    d, wav, est, sig, pm, voi = files
    got = mp.analysis_lossless(wav, est_file=est)
    v_pm_sec, v_voi = hostio.read_reaper_est_file(est, check_len_smpls=sig.size, fs=48000)
    ref = orc.analysis_lossless_from_pm(sig, 48000, v_pm_sec * 48000, v_voi)
    assert got[4] == 48000 and np.array_equal(got[5], ref[5]) and np.array_equal(got[3], ref[3])
    for a, b in zip(got[:3], ref[:3]):
        assert rms(a, b) < 1e-5
    # out_dir given: float32 files, returns None (src/magphase.py:2897-2904)
    out = str(d / 'lossless')
    os.makedirs(out)
    assert mp.analysis_lossless(wav, out_dir=out, est_file=est) is None
    mag = hostio.read_binfile(os.path.join(out, 'utt_a.mag'), dim=2049)
    assert mag.shape == ref[0].shape and rms(mag, ref[0].astype(np.float32)) < 1e-5
    assert np.array_equal(hostio.read_binfile(os.path.join(out, 'utt_a.shift'), dim=1), ref[5].astype(np.float32))
    # no REAPER binary here and no marks given: the package's own pitch-mark provider steps in (with a warning); the copy
    # synthesis of its analysis reproduces the recording like the one from the given marks does
    with pytest.warns(UserWarning, match='REAPER binary not found'):
        own = mp.analysis_lossless(wav)
    assert own[0].shape[1] == 2049 and abs(own[0].shape[0] - ref[0].shape[0]) < 0.15 * ref[0].shape[0]
    y = mp.synthesis_from_lossless(*own[:4], own[4])
    assert np.isfinite(y).all() and y.size > 0.8 * sig.size


def test_feature_extraction_and_waveform_generation_scripts(files):
    """scripts/batch_feature_extraction_for_tts.py -> scripts/batch_waveform_generation.py, one utterance."""
    import magphase_b200.magphase as mp
    from magphase_b200 import hostio
    d, wav, est, sig, pm, voi = files
    feats_dir, syn_dir = str(d / 'feats'), str(d / 'syn')
    os.makedirs(feats_dir); os.makedirs(syn_dir)
    mp.analysis_for_acoustic_modelling(wav, feats_dir, mag_dim=60, phase_dim=45, est_file=est)
    v_pm_sec, v_voi = hostio.read_reaper_est_file(est, check_len_smpls=sig.size, fs=48000)
    # the reference passes alpha_phase=b_mag_fbank_mel (=False -> 0.0) at src/magphase.py:3010: replicated
    ref = orc.analysis_compressed_from_pm(sig, 48000, v_pm_sec * 48000, v_voi, mag_dim=60, phase_dim=45, alpha_phase=0.0)
    for ext, dim, r in (('.mag', 60, ref[0]), ('.real', 45, ref[1]), ('.imag', 45, ref[2]), ('.lf0', 1, ref[3])):
        got = hostio.read_binfile(os.path.join(feats_dir, 'utt_a' + ext), dim=dim)
        assert got.shape == r.shape and rms(got, r.astype(np.float32).astype(np.float64)) < 1e-5
    np.random.seed(77)
    mp.synthesis_from_acoustic_modelling(feats_dir, 'utt_a', syn_dir, 60, 45, 48000, pf_type='magphase')
    y, fs = hostio.read_audio_file(os.path.join(syn_dir, 'utt_a.wav'))
    rd = lambda ext, dim: hostio.read_binfile(os.path.join(feats_dir, 'utt_a' + ext), dim=dim)
    np.random.seed(77)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        y_ref = orc.synthesis_from_compressed(orc.post_filter(rd('.mag', 60), 48000), rd('.real', 45), rd('.imag', 45),
                                              rd('.lf0', 1), 48000)
    y_ref = 0.98 * y_ref / np.max(np.abs(y_ref))                          # la.write_audio_file norm (src/libaudio.py:352-365)
    assert fs == 48000 and y.shape == y_ref.shape
    assert rms(y, y_ref) < 1e-4                                           # PCM16 quantisation of the written wav


def test_copy_synthesis_demo_script(tmp_path):
    """demos/demo_copy_synthesis.py = the reference's two copy-synthesis demos: runs end to end and the re-synthesised
    waveforms resemble the input (lossless: sample-accurate in the interior; low-dim: same length class, finite)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('demo_copy_synthesis', os.path.join(root, 'demos', 'demo_copy_synthesis.py'))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        done = demo.main(['--out-dir', str(tmp_path), '--mag-dim', '60'])
    assert len(done) == 2 and all(os.path.exists(p) for p, _ in done)
    from magphase_b200 import hostio
    x, fs = hostio.read_audio_file(os.path.join(str(tmp_path), 'synth_593.wav'))
    y, fs2 = hostio.read_audio_file(done[0][0])
    assert fs == fs2 == 48000 and abs(y.size - x.size) < 4096
    n = min(x.size, y.size)
    a, b = x[2000:n - 2000], y[2000:n - 2000]
    assert np.corrcoef(a, b)[0, 1] > 0.98            # (peak-normalised on writing: compare shape, not scale)
    z, _ = hostio.read_audio_file(done[1][0])
    assert np.all(np.isfinite(z)) and abs(z.size - x.size) < 4096

"""GPU parity on NATURAL speech: the input BASELINE config 1 names (demos/demo_copy_synthesis_lossless.py:57-91 on
demos/data_48k/wavs_nat/hvd_593.wav) plus a second bundled recording, against golden vectors generated from the real
reference (tests/golden/make_golden.py::natural; half-second int16 slices, seeded marks because REAPER is absent).
Studio recordings have quiet high-frequency bins (100 dB per-frame dynamic range): the hard case for the normalised
real / imag features and for anything computed in float32 -- the synthetic utterances have a -40 dB noise floor and are
the easy case.  The band-limited synthetic utterances (steep low-pass, -72 dB floor or quantisation noise only) are the
same stress without reference data.  Tolerance: 1e-5 RMS (BASELINE.json north_star), integer bookkeeping bit-exact."""
import os

import numpy as np
import pytest

import magphase_oracle as orc
from magphase_b200.synth import synth_utterance_band_limited

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = 1e-5


def rms(a, b):
    return float(np.sqrt(np.mean(np.abs(np.asarray(a) - np.asarray(b)) ** 2)))


@pytest.fixture(scope='module')
def mp():
    import magphase_b200.magphase as m
    return m


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'natural_48k.npz'))


def _inputs(g, tag):
    return g[tag + '_sig_i16'].astype(np.float64) / 32768.0, g[tag + '_pm'], g[tag + '_voi']


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_lossless_analysis_and_copy_synthesis_natural(mp, gold, tag):
    g = gold
    sig, pm, voi = _inputs(g, tag)
    mag, real, imag, f0, fs, v_shift = mp.analysis_lossless_from_pm(sig, int(g['fs']), pm, voi)
    assert np.array_equal(v_shift, g[tag + '_v_shift']) and np.array_equal(f0, g[tag + '_v_f0'])
    rows, st = g[tag + '_full_rows'], int(g['bin_step'])
    for a, k in ((mag, 'mag'), (real, 'real'), (imag, 'imag')):
        assert rms(a[rows], g['%s_%s_rows' % (tag, k)]) < TOL, k
        assert rms(a[:, ::st], g['%s_%s_cols' % (tag, k)]) < TOL, k
    # the float64 engine is far inside the bar even on near-silent bins
    assert rms(real[:, ::st], g[tag + '_real_cols']) < 1e-8 and rms(imag[:, ::st], g[tag + '_imag_cols']) < 1e-8
    y = mp.synthesis_from_lossless(mag, real, imag, f0, fs)
    assert y.shape == g[tag + '_syn'].shape and rms(y, g[tag + '_syn']) < TOL


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_compressed_analysis_natural(mp, gold, tag):
    """format_for_modelling of natural speech: CUDA vs the oracle's features of the REFERENCE's lossless analysis (the SPTK
    step is the restatement, unpinned).  Config 2 dims (60/45/45) and config 3 dims (phase_dim=10 with the reference's
    alpha_phase=False quirk, src/magphase.py:3010)."""
    g = gold
    sig, pm, voi = _inputs(g, tag)
    got = mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    for name, a in zip(('mag_mel_log', 'real_mel', 'imag_mel'), got[:3]):
        b = g['%s_%s' % (tag, name)]
        assert a.shape == b.shape and rms(a, b) < TOL, (name, rms(a, b))
    assert np.array_equal(got[3], g[tag + '_lf0']) and np.array_equal(got[4], g[tag + '_v_shift'])
    got3 = mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=10, alpha_phase=0.0)
    assert rms(got3[0], g[tag + '_mag_mel_log']) < TOL
    assert rms(got3[1], g[tag + '_real_mel_tts']) < TOL and rms(got3[2], g[tag + '_imag_mel_tts']) < TOL


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_compressed_synthesis_natural(mp, gold, tag):
    g = gold
    np.random.seed(int(g[tag + '_seed']))
    y = mp.synthesis_from_compressed(g[tag + '_mag_mel_log'], g[tag + '_real_mel'], g[tag + '_imag_mel'], g[tag + '_lf0'],
                                     48000, b_out_hpf=False)
    ref = g[tag + '_syn_compressed']
    assert y.shape == ref.shape and rms(y, ref) < TOL, rms(y, ref)


@pytest.mark.parametrize('floor_db', [-72.0, None])
def test_band_limited_synthetic(mp, floor_db):
    """Steep 7 kHz low-pass + faint floor, int16 steps: reacts to float32 butterflies like a studio recording."""
    sig, pm, voi = synth_utterance_band_limited(31, fs=48000, dur_s=0.6, floor_db=floor_db)
    ref = orc.analysis_lossless_from_pm(sig, 48000, pm, voi)
    got = mp.analysis_lossless_from_pm(sig, 48000, pm, voi)
    assert np.array_equal(got[5], ref[5]) and np.array_equal(got[3], ref[3])
    for a, b in zip(got[:3], ref[:3]):
        assert rms(a, b) < TOL
    cref = orc.format_for_modelling(*ref[:4], 48000, mag_dim=60, phase_dim=45)
    cgot = mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
    for name, a, b in zip(('mag_mel_log', 'real_mel', 'imag_mel'), cgot[:3], cref[:3]):
        assert rms(a, b) < TOL, (name, rms(a, b))
    y = mp.synthesis_from_lossless(*got[:4], 48000)
    y_ref = orc.synthesis_from_lossless(*ref[:4], 48000)
    assert y.shape == y_ref.shape and rms(y, y_ref) < TOL

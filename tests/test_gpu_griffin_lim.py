"""Pitch-synchronous Griffin-Lim on the device (the hot path iterated) against the oracle
restatement of src/magphase.py:3318-3373, which tests/test_oracle_vs_ref.py pins to the reference itself."""
import numpy as np
import pytest

import magphase_oracle as orc
from magphase_b200.synth import synth_utterance

pytestmark = pytest.mark.gpu


def rms(a, b):
    return float(np.sqrt(np.mean(np.abs(np.asarray(a) - np.asarray(b)) ** 2)))


@pytest.fixture(scope='module')
def feats():
    sig, pm, voi = synth_utterance(8, fs=48000, dur_s=0.6)
    mag, real, imag, f0, fs, shift = orc.analysis_lossless_from_pm(sig, 48000, pm, voi)
    return mag, real, imag, shift


def phase_dist(a, b):
    return np.abs(np.angle(np.exp(1j * (a - b))))


@pytest.mark.parametrize('init', ['linear', 'min_phase', 'random', 'array'])
def test_griffin_lim_vs_oracle(feats, init):
    import magphase_b200.magphase as mp
    mag, real, imag, shift = feats
    arg = np.angle(real + 1j * imag) if init == 'array' else init
    for niters in (1, 2, 6):
        np.random.seed(3)
        y_ref, ph_ref = orc.griffin_lim(mag.copy(), shift, phase_init=arg if isinstance(arg, str) else arg.copy(), niters=niters)
        np.random.seed(3)
        y, ph = mp.griffin_lim(mag.copy(), shift, phase_init=arg if isinstance(arg, str) else arg.copy(), niters=niters)
        assert y.shape == y_ref.shape and ph.shape == ph_ref.shape
        assert rms(y, y_ref) < 1e-9 * max(1.0, float(np.abs(y_ref).max())), (init, niters, rms(y, y_ref))
        # phases agree where the bin carries energy (np.angle of a numerically empty bin is arbitrary)
        strong = mag > 1e-6 * mag.max()
        assert np.max(phase_dist(ph, ph_ref)[strong]) < 1e-5, (init, niters)


def test_griffin_lim_with_arbitrary_window(feats):
    """win_func other than np.hanning (src/magphase.py:3318, :3362): the weights and gather indices are built once on the
    host, every analysis half runs on sig[idx] * w with MPB_WIN_RECT."""
    import magphase_b200.magphase as mp
    mag, real, imag, shift = feats
    for fn in (np.hamming, np.blackman):
        y_ref, ph_ref = orc.griffin_lim(mag.copy(), shift, phase_init='linear', niters=5, win_func=fn)
        y, ph = mp.griffin_lim(mag.copy(), shift, win_func=fn, phase_init='linear', niters=5)
        assert y.shape == y_ref.shape
        assert rms(y, y_ref) < 1e-9 * max(1.0, float(np.abs(y_ref).max())), rms(y, y_ref)
        strong = mag > 1e-6 * mag.max()
        assert np.max(phase_dist(ph, ph_ref)[strong]) < 1e-5


def test_griffin_lim_reduces_inconsistency(feats):
    """The point of the algorithm: the spectrogram of the output gets closer to the target magnitude."""
    import magphase_b200.magphase as mp
    mag, real, imag, shift = feats
    pm = np.cumsum(shift)

    def err(y):
        m = mp.analysis_lossless_from_pm(y, 48000, pm, np.ones(pm.size))[0]
        return float(np.linalg.norm(m - mag) / np.linalg.norm(mag))

    e = [err(mp.griffin_lim(mag.copy(), shift, phase_init='linear', niters=k)[0]) for k in (1, 4, 16)]
    assert e[2] < e[1] < e[0]


def test_griffin_lim_errors(feats):
    import magphase_b200.magphase as mp
    mag, real, imag, shift = feats
    with pytest.raises(ValueError):
        mp.griffin_lim(mag, np.full(shift.size, 3000.0))          # frames longer than fft_len / 2
    with pytest.raises(ValueError):
        mp.griffin_lim(mag, shift[:-1])
    with pytest.raises(ValueError):
        mp.griffin_lim(mag, shift, phase_init='bogus')


def test_griffin_lim_vs_reference_golden():
    """The device path against outputs of the real reference (tests/golden/griffin_lim_synth48k.npz)."""
    import os
    import magphase_b200.magphase as mp
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    g, gl = np.load(os.path.join(gold, 'lossless_synth48k.npz')), np.load(os.path.join(gold, 'griffin_lim_synth48k.npz'))
    sig = g['sig_i16'].astype(np.float64) / 32768.0
    mag = orc.analysis_lossless_from_pm(sig, int(g['fs']), g['pm'], g['voi'])[0]
    rows = gl['full_rows']
    strong = mag[rows] > 1e-6 * mag.max()
    for init in ('linear', 'min_phase', 'random'):
        np.random.seed(int(gl['seed']))
        y, ph = mp.griffin_lim(mag.copy(), gl['v_shift'], phase_init=init, niters=int(gl['niters']))
        assert y.shape == gl['syn_' + init].shape
        assert rms(y, gl['syn_' + init]) < 1e-9
        assert np.max(phase_dist(ph[rows], gl['phase_rows_' + init])[strong]) < 1e-5

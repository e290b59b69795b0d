"""The host entry points are pipelines over groups of utterances (mpb_stage.cu, mpb_api_mel.cu, mpb_api_syn.cu):
PCM-exact signals are narrowed to float32 on the host, others travel as float64 -- decided per group -- and groups
flow through H2D / compute / D2H streams.  None of that may change a number: a batch must equal its utterances
processed one by one, whatever mix of exact and inexact signals it holds, and stay within tolerance of the oracle."""
import numpy as np
import pytest

import magphase_oracle as orc
from magphase_b200.synth import synth_utterance

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rms(a, b):
    return float(np.sqrt(np.mean(np.abs(np.asarray(a) - np.asarray(b)) ** 2)))


@pytest.fixture(scope='module')
def mp():
    import magphase_b200.magphase as m
    return m


def _batch(n=12, dur=0.35):
    """Utterances of different lengths; every third one carries float64 samples that float32 cannot hold."""
    rng = np.random.default_rng(5)
    out = []
    for u in range(n):
        sig, pm, voi = synth_utterance(40 + u, fs=48000, dur_s=dur + 0.05 * (u % 4))
        if u % 3 == 1:
            sig = sig + rng.normal(scale=1e-9, size=sig.size)          # not representable in float32
            assert np.any(sig.astype(np.float32).astype(np.float64) != sig)
        out.append((sig, pm, voi))
    return out


def test_compressed_analysis_batch_equals_single_with_mixed_signal_precision(mp):
    utts = _batch()
    outs = mp.analysis_compressed_batch([u[0] for u in utts], 48000, [u[1] for u in utts], [u[2] for u in utts],
                                        mag_dim=60, phase_dim=45)
    for k, ((sig, pm, voi), got) in enumerate(zip(utts, outs)):
        one = mp.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
        for a, b in zip(got[:5], one[:5]):
            assert np.array_equal(a, b), k                               # same kernels, same rows: bit-identical
        if k in (0, 1, 5):
            ref = orc.analysis_compressed_from_pm(sig, 48000, pm, voi, mag_dim=60, phase_dim=45)
            for a, b in zip(got[:3], ref[:3]):
                assert rms(a, b) < TOL
            assert np.array_equal(got[3], ref[3]) and np.array_equal(got[4], ref[4])


def test_lossless_analysis_inexact_signal_takes_the_float64_path(mp):
    sig, pm, voi = synth_utterance(3, fs=48000, dur_s=0.4)
    sig = sig * (1.0 + 1e-12)                                            # every sample off the float32 grid
    got = mp.analysis_lossless_from_pm(sig, 48000, pm, voi)
    ref = orc.analysis_lossless_from_pm(sig, 48000, pm, voi)
    for a, b in zip(got[:3], ref[:3]):
        assert rms(a, b) < 1e-6


def test_compressed_synthesis_batch_equals_single_across_pipeline_groups(mp):
    """A batch long enough for several pipeline groups against the same utterances synthesised one at a time on the
    same NumPy stream: the noise is one continuous MT19937 sequence either way."""
    utts = _batch(n=48, dur=2.0)
    feats = mp.analysis_compressed_batch([u[0] for u in utts], 48000, [u[1] for u in utts], [u[2] for u in utts],
                                         mag_dim=60, phase_dim=45)
    feats = [f[:4] for f in feats]
    assert sum(f[0].shape[0] for f in feats) > 15000                     # > 1 group at 10,000 frames per group
    np.random.seed(77)
    ys = mp.synthesis_from_compressed_batch(feats, 48000, b_out_hpf=False)
    state_batch = np.random.get_state()
    np.random.seed(77)
    singles = [mp.synthesis_from_compressed(*f, 48000, b_out_hpf=False) for f in feats]
    state_single = np.random.get_state()
    assert state_batch[2] == state_single[2] and np.array_equal(state_batch[1], state_single[1])
    for k, (a, b) in enumerate(zip(ys, singles)):
        assert a.shape == b.shape
        assert rms(a, b) < 1e-7, (k, rms(a, b))                          # OLA run splits differ with the batch size


def test_narrow_element_types_on_request(mp):
    """PCM16 / float32 signals in, float32 features and waveforms out (the reference's own disk formats,
    src/libaudio.py:343-365, src/libutils.py:112-127): the same kernels, so the results equal the float64 API's up to the
    float32 rounding of the outputs, and the NumPy noise stream is advanced identically."""
    utts = _batch(n=5, dur=0.6)[:5]
    utts = [(np.round(s * 32768.0) / 32768.0, pm, voi) for s, pm, voi in utts]          # PCM16-exact samples
    pcm = [np.round(u[0] * 32768.0).astype(np.int16) for u in utts]
    ref = mp.analysis_compressed_batch([u[0] for u in utts], 48000, [u[1] for u in utts], [u[2] for u in utts], mag_dim=60, phase_dim=45)
    for sigs in (pcm, [u[0].astype(np.float32) for u in utts]):
        got = mp.analysis_compressed_batch(sigs, 48000, [u[1] for u in utts], [u[2] for u in utts], mag_dim=60, phase_dim=45,
                                           out_dtype=np.float32)
        for g, r in zip(got, ref):
            for a, b in zip(g[:3], r[:3]):
                assert a.dtype == np.float32 and a.shape == b.shape
                assert np.array_equal(a, b.astype(np.float32))                          # same kernels: only the output rounding differs
            assert np.array_equal(g[3], r[3]) and np.array_equal(g[4], r[4]) and g[3].dtype == np.float64
    feats32 = [g[:4] for g in got]
    np.random.seed(9)
    y64 = mp.synthesis_from_compressed_batch([(f[0].astype(np.float64), f[1].astype(np.float64), f[2].astype(np.float64), f[3])
                                              for f in feats32], 48000, b_out_hpf=False)
    st64 = np.random.get_state()
    np.random.seed(9)
    y32 = mp.synthesis_from_compressed_batch(feats32, 48000, b_out_hpf=False, out_dtype=np.float32)
    st32 = np.random.get_state()
    assert st64[2] == st32[2] and np.array_equal(st64[1], st32[1])
    for a, b in zip(y32, y64):
        assert a.dtype == np.float32 and a.shape == b.shape and rms(a, b) < 1e-7
    # with the output high-pass as well (single group path)
    np.random.seed(9)
    z64 = mp.synthesis_from_compressed_batch(feats32, 48000, b_out_hpf=True)
    np.random.seed(9)
    z32 = mp.synthesis_from_compressed_batch(feats32, 48000, b_out_hpf=True, out_dtype=np.float32)
    for a, b in zip(z32, z64):
        assert rms(a, b) < 1e-6


def test_lossless_narrow_element_types_on_request(mp):
    """The lossless chain with the reference's file formats: PCM16 / float32 signals in, float32 feature matrices out
    (.mag/.real/.imag are float32 files), float32 features in, float32 waveform out.  Same kernels and float64 butterflies:
    the features equal the float64 API's within float32 rounding, and so does the waveform of float32 features."""
    utts = _batch(n=3, dur=0.5)[:3]
    utts = [(np.round(s * 32768.0) / 32768.0, pm, voi) for s, pm, voi in utts]          # PCM16-exact samples
    pcm = [np.round(u[0] * 32768.0).astype(np.int16) for u in utts]
    pms, vois = [u[1] for u in utts], [u[2] for u in utts]
    ref = mp.analysis_lossless_batch([u[0] for u in utts], 48000, pms, vois)
    for sigs in (pcm, [u[0].astype(np.float32) for u in utts]):
        got = mp.analysis_lossless_batch(sigs, 48000, pms, vois, out_dtype=np.float32)
        for g, r in zip(got, ref):
            for a, b in zip(g[:3], r[:3]):
                # (float32 rows are normalised in float32 arithmetic after the float64 butterflies: within an ulp or two of
                #  the float64 rows, not their rounding)
                assert a.dtype == np.float32 and a.shape == b.shape
                assert np.max(np.abs(a - b)) <= 4e-7 * max(1.0, float(np.max(np.abs(b)))) and rms(a, b) < 1e-7 * max(1.0, rms(b, 0 * b))
            assert np.array_equal(g[3], r[3]) and np.array_equal(g[5], r[5]) and g[3].dtype == np.float64
    y64 = mp.synthesis_from_lossless_batch([r[:4] for r in ref], 48000)
    y32 = mp.synthesis_from_lossless_batch([g[:4] for g in got], 48000, out_dtype=np.float32)
    for a, b in zip(y32, y64):
        assert a.dtype == np.float32 and a.shape == b.shape and rms(a, b) < 2e-7
    with pytest.raises(ValueError):
        mp.analysis_lossless_batch(pcm, 48000, pms, vois, out_dtype=np.int16)

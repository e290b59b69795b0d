"""Host-side mirror of the reference's vocoder API (``src/magphase.py``) over the CUDA kernels.

Same function names, argument meaning, return values and error behaviour as the reference for the hot
path; the arithmetic runs in ``libmagphase_b200.so`` (sm_100a kernels) through ``ctypes``.  What stays on
the host is exactly what SURVEY.md 8(a) marks "bit-exact on host": pitch-mark rounding / truncation /
cumsum in float64 NumPy with the reference's expression order, per-sample-rate constants and file IO.

Differences from the reference, all additive:
  * REAPER (external binary) is optional: ``analysis_lossless`` also accepts ``est_file=`` (a REAPER
    ``.est`` file) or ``pm=(v_pm_sec, v_voi)``.
  * ``*_batch`` variants take lists of utterances and run them in one launch (the reference forks one
    process per utterance, ``src/libutils.py:32-63``).
  * inputs are never mutated (the reference zeroes DC/Nyquist imag in place, ``src/libaudio.py:375-376``).
"""
import ctypes as C
import os
import warnings
from subprocess import call

import numpy as np

from . import _lib
from . import hostio as io
from ._lib import MPB_F32, MPB_F64, WIN_BARTLETT25, WIN_HANN

MAGIC = -1.0e10   # src/libaudio.py:17

# compute precision of the analysis butterflies.  float64 is needed to keep the normalised real/imag
# features of near-silent bins within 1e-5 of the reference; synthesis is float32-safe
# but defaults to float64 for the drop-in API (the batch/bench path selects float32).
ANALYSIS_COMPUTE = MPB_F64
SYNTHESIS_COMPUTE = MPB_F64


# ----------------------------------------------------------------------------------------------
# constants per sample rate                                         src/magphase.py:3279-3317
# ----------------------------------------------------------------------------------------------
_ALPHA = {16000: 0.58, 22050: 0.65, 44100: 0.76, 48000: 0.77}
_FFT_LEN = {22050: 2048, 16000: 2048, 8000: 1024}
_CROSSFADE_CF = {48000: 5000, 16000: 2500, 44100: 4500, 22050: 3500}


def define_alpha(fs):
    try:
        return _ALPHA[fs]
    except KeyError:
        raise ValueError("Sample rate %d not supported yet." % (fs))


def define_fft_len(fs):
    return _FFT_LEN.get(fs, 4096)


def define_crossfade_params(fs):
    if fs not in (48000, 16000):
        warnings.warn('Constant crsf_cf not tested nor tunned to synthesise at fs=%d Hz.' % fs)
    return _CROSSFADE_CF.get(fs, 3500), 2000


# ----------------------------------------------------------------------------------------------
# integer / float64 bookkeeping (host, bit-exact)
# ----------------------------------------------------------------------------------------------
def round_to_int(x):
    """half-to-even, like np.round (src/libutils.py:131-133)."""
    return np.round(x).astype(int)


def shift_to_f0(v_shift, v_voi, fs, out='f0', b_smooth=True):
    """src/magphase.py:2198-2207"""
    v_f0 = v_voi * fs / v_shift.astype('float64')
    if b_smooth:
        from scipy import signal
        v_f0 = v_voi * signal.medfilt(v_f0)
    if out == 'lf0':
        v_f0 = f0_to_lf0(v_f0)
    return v_f0


def f0_to_shift(v_f0_in, fs, unv_frm_rate_ms=5):
    """src/magphase.py:2210-2215"""
    v_f0 = np.array(v_f0_in, dtype=np.float64, copy=True)
    v_f0[v_f0 == 0] = 1000.0 / unv_frm_rate_ms
    return fs / v_f0


def f0_to_lf0(v_f0):
    """src/libaudio.py:458-465"""
    with np.errstate(divide='ignore'):
        v_lf0 = np.log(v_f0)
    v_lf0[np.isinf(v_lf0)] = MAGIC
    return v_lf0


def frame_geometry(v_pm_smpls, n_smpls):
    """Extended marks P = [0, round(pm)..., n_smpls-1], v_shift (left lengths), right lengths.
    src/magphase.py:74-84, :112-117"""
    P = np.hstack((0, round_to_int(np.asarray(v_pm_smpls, dtype=np.float64)), n_smpls - 1)).astype(np.int64)
    return P, (P[1:-1] - P[:-2]), (P[2:] - P[1:-1])


def _win_codes(win_func, n):
    """Map the reference's win_func argument (a function or a per-frame list) to per-frame kernel codes."""
    def one(f):
        if f is np.hanning or f == 'hann':
            return WIN_HANN
        if f is voi_noise_window or f == 'bartlett2.5':
            return WIN_BARTLETT25
        raise ValueError('win_func %r is not available on the CUDA path (np.hanning / voi_noise_window only)' % (f,))
    if isinstance(win_func, (list, tuple)):
        if len(win_func) != n:
            raise ValueError('win_func list length must equal the number of frames')
        return np.array([one(f) for f in win_func], dtype=np.uint8)
    code = one(win_func)
    return None if code == WIN_HANN else np.full(n, code, dtype=np.uint8)


def voi_noise_window(length):
    """Host definition kept for API compatibility (src/magphase.py:67-69); the kernels evaluate it in closed form."""
    return np.bartlett(length) ** 2.5


def _check_frames(v_shift, v_rights, fft_len):
    lens = v_shift + v_rights + 1
    too_long = np.nonzero(lens > fft_len)[0]
    for f in too_long:   # same warning as src/magphase.py:305-315, once per offending frame
        warnings.warn("fft_len (%d) is shorter than the current detected frame length (%d). "
                      "This issue is not very critical, but if it occurs often "
                      "(e.g., more than 3 times per utterance), please increase de FFT length." % (fft_len, lens[f]))


# ----------------------------------------------------------------------------------------------
# analysis
# ----------------------------------------------------------------------------------------------
def _frames_call(l_sig, l_pm, fft_len, l_win, mode, compute=None):
    """Shared driver of the analysis kernels for a list of utterances (host buffers in, host buffers out)."""
    compute = ANALYSIS_COMPUTE if compute is None else compute
    H = fft_len // 2 + 1
    sig_off = np.zeros(len(l_sig) + 1, dtype=np.int64)
    centres, lefts, rights, wins, shifts = [], [], [], [], []
    any_win = False
    for u, (sig, pm) in enumerate(zip(l_sig, l_pm)):
        P, v_shift, v_rights = frame_geometry(pm, sig.size)
        _check_frames(v_shift, v_rights, fft_len)
        sig_off[u + 1] = sig_off[u] + sig.size
        centres.append(P[1:-1] + sig_off[u])
        lefts.append(v_shift)
        rights.append(v_rights)
        shifts.append(v_shift)
        w = _win_codes(l_win[u], v_shift.size)
        any_win |= w is not None
        wins.append(w)
    centre = np.ascontiguousarray(np.concatenate(centres), dtype=np.int64)
    left = np.ascontiguousarray(np.concatenate(lefts), dtype=np.int32)
    right = np.ascontiguousarray(np.concatenate(rights), dtype=np.int32)
    win = None
    if any_win:
        win = np.ascontiguousarray(np.concatenate([w if w is not None else np.zeros(s.size, np.uint8)
                                                   for w, s in zip(wins, shifts)]), dtype=np.uint8)
    sig_all = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float64) for s in l_sig]))
    nfrm = centre.size
    l = _lib.lib()
    if mode == 'fft':
        out = np.empty((nfrm, H), dtype=np.complex128)
        _lib.check(l.mpb_frames_fft_host(_lib.ctx(), _lib.ptr(sig_all), sig_all.size, _lib.ptr(centre), _lib.ptr(left),
                                         _lib.ptr(right), _lib.ptr(win), nfrm, fft_len, compute, _lib.ptr(out)))
        outs = (out,)
    else:
        outs = tuple(np.empty((nfrm, H), dtype=np.float64) for _ in range(3))
        _lib.check(l.mpb_analysis_lossless_host(_lib.ctx(), _lib.ptr(sig_all), sig_all.size, _lib.ptr(centre),
                                                _lib.ptr(left), _lib.ptr(right), _lib.ptr(win), nfrm, fft_len, compute,
                                                _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2])))
    frm_off = np.concatenate(([0], np.cumsum([s.size for s in shifts])))
    return outs, shifts, frm_off


def _expand_epochs(v_pm_smpls, nwin_per_pitch_period):
    """Intermediate epochs for nwin_per_pitch_period >= 1.  src/magphase.py:280-288"""
    if nwin_per_pitch_period == 0.5:
        return v_pm_smpls
    if nwin_per_pitch_period >= 1.0:
        n = int(nwin_per_pitch_period * 2)
        step = np.diff(v_pm_smpls) / float(n)
        return (v_pm_smpls[:-1][None, :] + step[None, :] * np.arange(n)[:, None]).flatten(order='F')
    return v_pm_smpls


def analysis_with_del_comp_from_pm(v_in_sig, fs, v_pm_smpls, fft_len=None, win_func=np.hanning,
                                   nwin_per_pitch_period=0.5):
    """Complex half spectra of the pitch-synchronous frames + v_shift.  src/magphase.py:266-334"""
    if fft_len is None:
        fft_len = define_fft_len(fs)
    v_pm = _expand_epochs(np.asarray(v_pm_smpls, dtype=np.float64), nwin_per_pitch_period)
    (m_fft,), shifts, _ = _frames_call([np.asarray(v_in_sig)], [v_pm], fft_len, [win_func], 'fft')
    return m_fft, shifts[0].astype(int)


def analysis_lossless_from_pm(v_sig, fs, v_pm_smpls, v_voi, fft_len=None):
    """analysis_lossless minus file reading and epoch detection (pitch marks in samples + voicing given)."""
    if fft_len is None:
        fft_len = define_fft_len(fs)
    (m_mag, m_real, m_imag), shifts, _ = _frames_call([np.asarray(v_sig)], [np.asarray(v_pm_smpls)], fft_len,
                                                      [np.hanning], 'feats')
    v_shift = shifts[0].astype(int)
    v_f0 = shift_to_f0(v_shift, np.asarray(v_voi, dtype=np.float64), fs, out='f0', b_smooth=False)
    return m_mag, m_real, m_imag, v_f0, fs, v_shift


def analysis_lossless_batch(l_sig, fs, l_pm_smpls, l_voi, fft_len=None):
    """Batched analysis_lossless_from_pm: one kernel launch for all utterances.
    Returns a list of (m_mag, m_real, m_imag, v_f0, fs, v_shift), the arrays being views into three batch matrices."""
    if fft_len is None:
        fft_len = define_fft_len(fs)
    (mag, real, imag), shifts, off = _frames_call([np.asarray(s) for s in l_sig], l_pm_smpls, fft_len,
                                                  [np.hanning] * len(l_sig), 'feats')
    out = []
    for u in range(len(l_sig)):
        a, b = off[u], off[u + 1]
        v_shift = shifts[u].astype(int)
        v_f0 = shift_to_f0(v_shift, np.asarray(l_voi[u], dtype=np.float64), fs, out='f0', b_smooth=False)
        out.append((mag[a:b], real[a:b], imag[a:b], v_f0, fs, v_shift))
    return out


def get_pitch_marks_and_voicing(wav_file, n_smpls, fs, est_file=None, pm=None):
    """Pitch marks (seconds) + voicing for analysis_lossless: explicit arrays, a REAPER .est file, or the
    REAPER binary itself when installed (src/magphase.py:2875-2878, src/libaudio.py:421-455)."""
    if pm is not None:
        v_pm_sec, v_voi = pm
        return np.asarray(v_pm_sec, dtype=np.float64), np.asarray(v_voi, dtype=np.float64)
    if est_file is not None:
        return io.read_reaper_est_file(est_file, check_len_smpls=n_smpls, fs=fs)
    reaper = io.find_tool('reaper')
    if reaper is None:
        raise RuntimeError('analysis_lossless: REAPER binary not found (config.ini [TOOLS] bin_dir / tools/bin). '
                           'Pass est_file=<REAPER .est file> or pm=(v_pm_sec, v_voi).')
    tmp_est = io.ins_pid('temp.est')
    print("Extracting epochs with REAPER...")
    call(reaper + " -s -x 400 -m 50 -a -u 0.005 -i %s -p %s" % (wav_file, tmp_est), shell=True)
    try:
        return io.read_reaper_est_file(tmp_est, check_len_smpls=n_smpls, fs=fs)
    finally:
        os.remove(tmp_est)


def analysis_lossless(wav_file, fft_len=None, out_dir=None, est_file=None, pm=None):
    """src/magphase.py:2869-2906.  Returns (m_mag, m_real, m_imag, v_f0, fs, v_shift), or writes
    .mag/.real/.imag/.f0/.shift float32 files and returns None when out_dir is a str."""
    v_sig, fs = io.read_audio_file(wav_file)
    v_pm_sec, v_voi = get_pitch_marks_and_voicing(wav_file, len(v_sig), fs, est_file=est_file, pm=pm)
    v_pm_smpls = v_pm_sec * fs
    m_mag, m_real, m_imag, v_f0, fs, v_shift = analysis_lossless_from_pm(v_sig, fs, v_pm_smpls, v_voi, fft_len=fft_len)
    if type(out_dir) is str:
        file_id = os.path.basename(wav_file).split(".")[0]
        for arr, ext in ((m_mag, '.mag'), (m_real, '.real'), (m_imag, '.imag'), (v_f0, '.f0'), (v_shift, '.shift')):
            io.write_binfile(arr, os.path.join(out_dir, file_id + ext))
        return
    return m_mag, m_real, m_imag, v_f0, fs, v_shift


# ----------------------------------------------------------------------------------------------
# lossless synthesis
# ----------------------------------------------------------------------------------------------
def ola_geometry(v_pm, fft_len):
    """Integer geometry of ola() (src/magphase.py:34-62): truncated marks, position of the first output
    sample on the pitch-mark axis and output length, Python slice semantics included."""
    v_pm = np.asarray(v_pm).astype(int)                       # truncation (:36)
    buf_len = int(v_pm[-1]) + fft_len
    v_shift = np.diff(np.hstack((0, v_pm)))
    start, stop, _ = slice(fft_len // 2 - int(v_pm[0]), None).indices(buf_len)
    n1 = max(stop - start, 0)
    n_out = min(n1, max(int(v_pm[-1] + v_shift[-1] + 1), 0))
    t0 = start + int(v_pm[0]) - fft_len // 2                  # buffer index j <-> position j + pm[0] - N/2
    return v_pm.astype(np.int32), t0, n_out


def _synthesis_lossless_call(l_feats, l_pm_int, l_t0, l_nout, fft_len, compute=None):
    compute = SYNTHESIS_COMPUTE if compute is None else compute
    n_utt = len(l_feats)
    frm_off = np.zeros(n_utt + 1, dtype=np.int64)
    out_off = np.zeros(n_utt + 1, dtype=np.int64)
    for u in range(n_utt):
        frm_off[u + 1] = frm_off[u] + l_pm_int[u].size
        out_off[u + 1] = out_off[u] + l_nout[u]
    cat = lambda i: np.ascontiguousarray(np.concatenate([np.asarray(f[i], dtype=np.float64) for f in l_feats], axis=0))
    mag, real, imag = cat(0), cat(1), cat(2)
    H = fft_len // 2 + 1
    if mag.shape[1] != H or real.shape != mag.shape or imag.shape != mag.shape:
        raise ValueError('feature matrices must be nfrms x %d' % H)
    pm = np.ascontiguousarray(np.concatenate(l_pm_int), dtype=np.int32)
    t0 = np.ascontiguousarray(l_t0, dtype=np.int32)
    out = np.empty(int(out_off[-1]), dtype=np.float64)
    _lib.check(_lib.lib().mpb_synthesis_lossless_host(
        _lib.ctx(), _lib.ptr(mag), _lib.ptr(real), _lib.ptr(imag), _lib.ptr(pm), pm.size, _lib.ptr(frm_off),
        _lib.ptr(out_off), _lib.ptr(t0), n_utt, fft_len, compute, _lib.ptr(out), out.size))
    return [out[out_off[u]:out_off[u + 1]] for u in range(n_utt)]


def synthesis_from_lossless(m_mag, m_real, m_imag, v_f0, fs):
    """src/magphase.py:1759-1776"""
    return synthesis_from_lossless_batch([(m_mag, m_real, m_imag, v_f0)], fs)[0]


def synthesis_from_lossless_batch(l_feats, fs):
    """Batched synthesis_from_lossless; l_feats is a list of (m_mag, m_real, m_imag, v_f0)."""
    l_pm, l_t0, l_n = [], [], []
    fft_len = None
    for (m_mag, m_real, m_imag, v_f0) in l_feats:
        n_fft = 2 * (np.shape(m_mag)[1] - 1)
        if fft_len is None:
            fft_len = n_fft
        elif fft_len != n_fft:
            raise ValueError('all utterances of a batch must share fft_len')
        v_shift = f0_to_shift(np.asarray(v_f0, dtype=np.float64), fs, unv_frm_rate_ms=5)
        v_pm = np.cumsum(v_shift)                              # float cumsum, truncated inside ola (:1771-1772, :36)
        pm_int, t0, n_out = ola_geometry(v_pm, fft_len)
        l_pm.append(pm_int)
        l_t0.append(t0)
        l_n.append(n_out)
    return _synthesis_lossless_call(l_feats, l_pm, l_t0, l_n, fft_len)

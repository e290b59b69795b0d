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
import threading
import warnings
from subprocess import call

import numpy as np

from . import _lib
from . import hostio as io
from ._lib import MPB_F32, MPB_F64, WIN_BARTLETT25, WIN_HANN, WIN_RECT

MAGIC = -1.0e10   # src/libaudio.py:17

# Compute precision of the LOSSLESS entry points (analysis_lossless*, synthesis_from_lossless*, griffin_lim).  float64
# butterflies are needed to keep the normalised real/imag features of near-silent bins within 1e-5 of the reference
# (SURVEY.md 7.3-1); lossless synthesis is float32-safe but defaults to float64 for the drop-in API (the device-resident
# batch / bench path selects float32).
# These constants do NOT apply to the compressed chain, whose precision is fixed: analysis_compressed* runs float64
# butterflies and keeps everything after X[k] in float32 (log periodograms, 3xTF32 tile products, float32 mel cepstra like
# SPTK's files); synthesis_from_compressed* is float32 throughout, including the noise samples (NumPy's uniform draws are
# reproduced bit for bit as float64 and narrowed to float32 before the noise FFT); results are widened to float64 on return.
# All of it is held to 1e-5 RMS against the float64 oracle.
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


def _win_builtin(f):
    """Kernel code of a window the kernels evaluate in closed form, None for any other callable."""
    if f is np.hanning or (isinstance(f, str) and f == 'hann'):
        return WIN_HANN
    if f is voi_noise_window or (isinstance(f, str) and f == 'bartlett2.5'):
        return WIN_BARTLETT25
    if callable(f):
        return None
    raise ValueError('win_func %r is neither a window function nor a list of window functions' % (f,))


def _win_list(win_func, n):
    if isinstance(win_func, (list, tuple)):
        if len(win_func) != n:
            raise ValueError('win_func list length must equal the number of frames')
        return list(win_func)
    return [win_func] * n


def _has_custom_window(win_func):
    """True when win_func (a function or a per-frame list, src/magphase.py:102-108) names a window the kernels do not
    evaluate themselves."""
    fs = win_func if isinstance(win_func, (list, tuple)) else (win_func,)
    return any(_win_builtin(f) is None for f in fs)


def _win_codes(win_func, n):
    """Map the reference's win_func argument (a function or a per-frame list) to per-frame kernel codes; None = all Hann.
    Only for the two windows the kernels evaluate in closed form (see _has_custom_window / window_weights for the rest)."""
    codes = [_win_builtin(f) for f in _win_list(win_func, n)]
    if any(c is None for c in codes):
        raise ValueError('arbitrary window callables go through window_weights(), not through kernel codes')
    if all(c == WIN_HANN for c in codes):
        return None
    return np.array(codes, dtype=np.uint8)


def window_weights(l_fns, left, right):
    """The windows la.gen_non_symmetric_win(left[f], right[f], l_fns[f]) of all frames back to back (float64) and their
    offsets: hstack(w(1+2l)[0:l+1], flipud(w(1+2r)[0:r+1])[1:]) per frame (src/libaudio.py:70-84), evaluated by the
    caller's own callables once per distinct (function, side length).  This is how arbitrary ``win_func`` arguments reach
    the kernels: the mirror multiplies the weights into the frames' samples and launches with MPB_WIN_RECT."""
    left = np.asarray(left, dtype=np.int64)
    right = np.asarray(right, dtype=np.int64)
    off = _seg_offsets(left + right + 1)
    w_all = np.empty(int(off[-1]), dtype=np.float64)
    halves = {}

    def half(fn, s):
        key = (id(fn), s)
        h = halves.get(key)
        if h is None:
            fn_ = {WIN_HANN: np.hanning, WIN_BARTLETT25: voi_noise_window}.get(_win_builtin(fn), fn)
            h = np.asarray(fn_(1 + 2 * s), dtype=np.float64)
            if h.shape != (1 + 2 * s,):
                raise ValueError('win_func(%d) must return %d values' % (1 + 2 * s, 1 + 2 * s))
            h = halves[key] = h[:s + 1]
        return h
    for f in range(left.size):
        l, r, a = int(left[f]), int(right[f]), int(off[f])
        w_all[a:a + l + 1] = half(l_fns[f], l)
        w_all[a + l + 1:a + l + 1 + r] = half(l_fns[f], r)[::-1][1:]
    return w_all, off


def prewindowed_frames(sig_all, centre, left, right, l_fns):
    """Frames sig[c-l .. c+r] * window laid back to back, and the marks' positions inside that buffer (centre[f] must be
    the absolute index of frame f's mark in sig_all).  Launch the analysis kernels on the result with MPB_WIN_RECT."""
    w_all, off = window_weights(l_fns, left, right)
    lens = np.diff(off)
    idx = np.repeat(np.asarray(centre, dtype=np.int64) - left - off[:-1], lens) + np.arange(int(off[-1]), dtype=np.int64)
    return np.asarray(sig_all, dtype=np.float64)[idx] * w_all, np.ascontiguousarray(off[:-1] + left), idx, w_all


def windowing(v_sig, v_pm, win_func=np.hanning):
    """The reference's frame extractor as a callable (src/magphase.py:74-119): the list of windowed frames
    sig[P[f] : P[f+2]+1] * gen_non_symmetric_win(left, right, win_func[f]) and the integer bookkeeping
    (v_lens, v_pm_plus, v_shift, v_rights).  Host NumPy, for callers of the reference's helper: the analysis kernels never
    materialise this list -- they gather and window the samples on the fly (mpb_frame.cuh:load_frame)."""
    v_sig = np.asarray(v_sig)
    P, left, right = frame_geometry(v_pm, v_sig.size)
    n = left.size
    pre, _, _, _ = prewindowed_frames(v_sig, P[1:-1], left, right, _win_list(win_func, n))
    off = _seg_offsets(left + right + 1)
    l_frames = [pre[off[f]:off[f + 1]] for f in range(n)]
    return l_frames, (left + right + 1).astype(int), P.astype(int), left.astype(int), right.astype(int)


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


def _seg_offsets(lens):
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return off


def batch_frame_geometry(l_pm_smpls, l_n_smpls):
    """frame_geometry() of a whole batch in a handful of NumPy calls (the per-utterance Python loop was a third of the
    end-to-end time): rounded marks of all utterances back to back, left / right frame lengths, frame offsets.
    Same integer results as the per-utterance function (tests/test_batch_geometry_cpu.py)."""
    lens = np.array([np.size(p) for p in l_pm_smpls], dtype=np.int64)
    off = _seg_offsets(lens)
    if np.any(lens == 0):
        raise ValueError('every utterance needs at least one pitch mark')
    pm = round_to_int(np.concatenate([np.asarray(p, dtype=np.float64) for p in l_pm_smpls])).astype(np.int64)
    prev = np.empty_like(pm)
    prev[1:] = pm[:-1]
    prev[off[:-1]] = 0
    nxt = np.empty_like(pm)
    nxt[:-1] = pm[1:]
    nxt[off[1:] - 1] = np.asarray(l_n_smpls, dtype=np.int64) - 1
    return pm, pm - prev, nxt - pm, off


def _medfilt3_segments(v, off):
    """medfilt3() applied to every segment v[off[u]:off[u+1]] (zero padded at the segment edges), flat."""
    a = np.empty_like(v)
    a[1:] = v[:-1]
    a[off[:-1]] = 0.0
    c = np.empty_like(v)
    c[:-1] = v[1:]
    c[off[1:] - 1] = 0.0
    return np.maximum(np.minimum(a, v), np.minimum(np.maximum(a, v), c))


_WORKSPACE = threading.local()


def _workspace(name, n, dtype):
    """Reusable scratch array (descriptor arrays that only live for the duration of one C call): fresh NumPy
    allocations of this size are mmap-ed and page-faulted on every call.  Per thread: ctypes releases the GIL inside
    the C call that reads these arrays."""
    ws = _WORKSPACE.__dict__
    a = ws.get(name)
    if a is None or a.size < n or a.dtype != np.dtype(dtype):
        a = np.empty(max(int(n), 1024) * 5 // 4, dtype=dtype)
        ws[name] = a
    return a[:n]


def _analysis_geometry_c(l_pm_smpls, l_n_smpls, l_voi, fs):
    """batch_frame_geometry + shift_to_f0(b_smooth=False) + the argument of the log of _lf0_smoothed for a batch, in one
    C pass (mpb_analysis_geometry; integer arithmetic, IEEE multiply / divide and a median of three).  Returns
    (centre, left, right, voi8, f0_med, frm_off); everything except frm_off lives in the module workspace."""
    lens = np.array([np.size(p) for p in l_pm_smpls], dtype=np.int64)
    if np.any(lens == 0):
        raise ValueError('every utterance needs at least one pitch mark')
    if any(np.size(v) != k for v, k in zip(l_voi, lens)):
        raise ValueError('voicing and pitch-mark arrays must have the same length')
    off = _seg_offsets(lens)
    n = int(off[-1])
    pm = np.round(np.concatenate([np.asarray(p, dtype=np.float64) for p in l_pm_smpls])).astype(np.int64)   # round_to_int
    voi_in = np.concatenate([np.asarray(v, dtype=np.float64) for v in l_voi])
    n_smpls = np.ascontiguousarray(l_n_smpls, dtype=np.int64)
    centre, left, right = _workspace('a_centre', n, np.int64), _workspace('a_left', n, np.int32), _workspace('a_right', n, np.int32)
    f0_med, voi8 = _workspace('a_f0med', n, np.float64), _workspace('a_voi8', n, np.uint8)
    _lib.check(_lib.lib().mpb_analysis_geometry(_lib.ptr(pm), _lib.ptr(off), _lib.ptr(n_smpls), len(lens), _lib.ptr(voi_in),
                                                float(fs), _lib.ptr(centre), _lib.ptr(left), _lib.ptr(right),
                                                _lib.ptr(f0_med), _lib.ptr(voi8)))
    return centre, left, right, voi8, f0_med, off


# ----------------------------------------------------------------------------------------------
# analysis
# ----------------------------------------------------------------------------------------------
def _frames_call(l_sig, l_pm, fft_len, l_win, mode, compute=None, out_dtype=np.float64):
    """Shared driver of the analysis kernels for a list of utterances (host buffers in, host buffers out).  Signals that
    are ALL int16 (PCM as the wav file holds it) or ALL float32 travel as they are; out_dtype float32 halves the feature bytes."""
    compute = ANALYSIS_COMPUTE if compute is None else compute
    out_dtype = np.dtype(out_dtype)
    if out_dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
        raise ValueError('out_dtype must be float64 or float32')
    H = fft_len // 2 + 1
    sig_off = _seg_offsets([s.size for s in l_sig])
    pm, left64, right64, frm_off = batch_frame_geometry(l_pm, [s.size for s in l_sig])
    _check_frames(left64, right64, fft_len)
    nfr = np.diff(frm_off)
    centre = np.ascontiguousarray(pm + np.repeat(sig_off[:-1], nfr))
    left = left64.astype(np.int32)
    right = right64.astype(np.int32)
    shifts = [left64[frm_off[u]:frm_off[u + 1]] for u in range(len(l_sig))]
    kinds = {np.asarray(s).dtype for s in l_sig}
    sig_np = kinds.pop() if len(kinds) == 1 and next(iter(kinds)) in (np.dtype(np.int16), np.dtype(np.float32)) else np.dtype(np.float64)
    sig_all = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=sig_np) for s in l_sig]))
    nfrm = centre.size
    win = None
    if any(_has_custom_window(w) for w in l_win):
        # a window the kernels do not know in closed form: evaluate the caller's callables on the host, multiply them into
        # the frames' samples (float64) and run the kernels with weight 1 on that buffer
        if np.any(left64 < 0) or np.any(right64 < 0):
            raise ValueError('magphase_b200: frame reaches outside the signal (pitch marks must increase and lie inside it)')
        fns = [f for u in range(len(l_sig)) for f in _win_list(l_win[u], int(nfr[u]))]
        scale = 1.0 / 32768.0 if sig_np == np.dtype(np.int16) else 1.0
        sig_all, centre, _, _ = prewindowed_frames(sig_all.astype(np.float64) * scale, centre, left64, right64, fns)
        sig_np = np.dtype(np.float64)
        win = np.full(nfrm, WIN_RECT, dtype=np.uint8)
    else:
        wins = [_win_codes(l_win[u], int(nfr[u])) for u in range(len(l_sig))]
        if any(w is not None for w in wins):
            win = np.ascontiguousarray(np.concatenate([w if w is not None else np.zeros(int(k), np.uint8)
                                                       for w, k in zip(wins, nfr)]), dtype=np.uint8)
    sig_code = {np.dtype(np.float64): MPB_F64, np.dtype(np.float32): MPB_F32, np.dtype(np.int16): _lib.MPB_I16}[sig_np]
    l = _lib.lib()
    if mode == 'fft':
        sig_all = np.ascontiguousarray(sig_all if sig_np == np.dtype(np.float64) else
                                       (sig_all.astype(np.float64) / 32768.0 if sig_np == np.dtype(np.int16) else sig_all), dtype=np.float64)
        out = np.empty((nfrm, H), dtype=np.complex128)
        _lib.check(l.mpb_frames_fft_host(_lib.ctx(), _lib.ptr(sig_all), sig_all.size, _lib.ptr(centre), _lib.ptr(left),
                                         _lib.ptr(right), _lib.ptr(win), nfrm, fft_len, compute, _lib.ptr(out)))
        outs = (out,)
    else:
        outs = tuple(_lib.pinned.empty((nfrm, H), dtype=out_dtype) for _ in range(3))
        _lib.check(l.mpb_analysis_lossless_host2(_lib.ctx(), _lib.ptr(sig_all), sig_code, sig_all.size, _lib.ptr(centre),
                                                 _lib.ptr(left), _lib.ptr(right), _lib.ptr(win), nfrm, fft_len, compute,
                                                 _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]),
                                                 MPB_F32 if out_dtype == np.dtype(np.float32) else MPB_F64))
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


def compute_lossless_feats(m_fft, v_shift, v_voi, fs):
    """mag, real / |X|, imag / |X| (0 where |X| == 0) and f0 from ready-made half spectra -- the second step of the
    reference's analysis as a function of its own (src/magphase.py:457-476), for callers that hold the output of
    analysis_with_del_comp_from_pm.  Elementwise on the device; analysis_lossless* fuse it into the analysis kernel."""
    m_fft = np.ascontiguousarray(m_fft, dtype=np.complex128)
    m_mag, m_real, m_imag = (np.empty(m_fft.shape, dtype=np.float64) for _ in range(3))
    _lib.check(_lib.lib().mpb_lossless_feats_host(_lib.ctx(), _lib.ptr(m_fft), m_fft.size, _lib.ptr(m_mag), _lib.ptr(m_real),
                                                  _lib.ptr(m_imag)))
    v_f0 = shift_to_f0(np.asarray(v_shift), np.asarray(v_voi, dtype=np.float64), fs, out='f0', b_smooth=False)
    return m_mag, m_real, m_imag, v_f0


def analysis_lossless_from_pm(v_sig, fs, v_pm_smpls, v_voi, fft_len=None):
    """analysis_lossless minus file reading and epoch detection (pitch marks in samples + voicing given)."""
    if fft_len is None:
        fft_len = define_fft_len(fs)
    (m_mag, m_real, m_imag), shifts, _ = _frames_call([np.asarray(v_sig)], [np.asarray(v_pm_smpls)], fft_len,
                                                      [np.hanning], 'feats')
    v_shift = shifts[0].astype(int)
    v_f0 = shift_to_f0(v_shift, np.asarray(v_voi, dtype=np.float64), fs, out='f0', b_smooth=False)
    return m_mag, m_real, m_imag, v_f0, fs, v_shift


def analysis_lossless_batch(l_sig, fs, l_pm_smpls, l_voi, fft_len=None, out_dtype=np.float64):
    """Batched analysis_lossless_from_pm: one kernel launch for all utterances.
    Returns a list of (m_mag, m_real, m_imag, v_f0, fs, v_shift), the arrays being views into three batch matrices.
    Element types on request, as for the compressed chain: int16 (PCM16 of the wav file) or float32 signals are uploaded as
    they are, ``out_dtype=np.float32`` returns float32 feature matrices -- the precision of the reference's own feature files
    (src/libutils.py:122-127) and half of the 98 KB per frame that otherwise cross PCIe.  The butterflies stay float64."""
    if fft_len is None:
        fft_len = define_fft_len(fs)
    (mag, real, imag), shifts, off = _frames_call([np.asarray(s) for s in l_sig], l_pm_smpls, fft_len,
                                                  [np.hanning] * len(l_sig), 'feats', out_dtype=out_dtype)
    out = []
    for u in range(len(l_sig)):
        a, b = off[u], off[u + 1]
        v_shift = shifts[u].astype(int)
        v_f0 = shift_to_f0(v_shift, np.asarray(l_voi[u], dtype=np.float64), fs, out='f0', b_smooth=False)
        out.append((mag[a:b], real[a:b], imag[a:b], v_f0, fs, v_shift))
    return out


def get_pitch_marks_and_voicing(wav_file, n_smpls, fs, est_file=None, pm=None):
    """Pitch marks (seconds) + voicing for analysis_lossless: explicit arrays, a REAPER .est file, the REAPER binary
    itself when installed (src/magphase.py:2875-2878, src/libaudio.py:421-455) -- or, without it, this package's own
    provider (magphase_b200/pitchmarks.py: autocorrelation voicing + peak picking; same output layout, no parity claim)."""
    if pm is not None:
        v_pm_sec, v_voi = pm
        return np.asarray(v_pm_sec, dtype=np.float64), np.asarray(v_voi, dtype=np.float64)
    if est_file is not None:
        return io.read_reaper_est_file(est_file, check_len_smpls=n_smpls, fs=fs)
    reaper = io.find_tool('reaper')
    if reaper is None:
        from .pitchmarks import estimate_pitch_marks
        warnings.warn('REAPER binary not found (config.ini [TOOLS] bin_dir / tools/bin): pitch marks come from '
                      'magphase_b200.pitchmarks.estimate_pitch_marks (pass est_file= / pm= to supply your own)')
        v_sig, fs_w = io.read_audio_file(wav_file)
        v_pm_sec, v_voi = estimate_pitch_marks(v_sig, fs_w)
        ok = np.round(v_pm_sec * fs_w).astype(int) < (n_smpls - 1)
        return v_pm_sec[ok], v_voi[ok]
    tmp_est = io.ins_pid('temp.est')
    print("Extracting epochs with REAPER...")
    # the reference's command line (src/libaudio.py:452), as an argument list: file names with blanks or shell characters
    # are passed through untouched
    rc = call([reaper, '-s', '-x', '400', '-m', '50', '-a', '-u', '0.005', '-i', str(wav_file), '-p', tmp_est])
    try:
        if rc != 0 or not os.path.isfile(tmp_est):
            raise RuntimeError('REAPER failed on %s (exit status %s)' % (wav_file, rc))
        return io.read_reaper_est_file(tmp_est, check_len_smpls=n_smpls, fs=fs)
    finally:
        if os.path.exists(tmp_est):
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


def raised_hanning(length, att=1.0):
    """src/magphase.py:25-31"""
    return (1 - att) + att * np.hanning(length)


def ola(m_frm, v_pm, win_func=None):
    """Pitch-synchronous overlap-add of ready-made frames (src/magphase.py:34-62) on the device: the frame centre (column
    frmlen/2) of frame i lands on int(v_pm[i]); the sums are taken in frame order like the reference's loop, so the result is
    bit-identical to it.  ``win_func``: optional per-frame centred window la.gen_centr_win(shift[i], shift[i+1], frmlen,
    win_func) (src/libaudio.py:90-103), evaluated on the host; unlike the reference the caller's m_frm is NOT modified.
    (synthesis_from_lossless / synthesis_from_compressed do their overlap-add inside the inverse-FFT kernels; this is the
    standalone operator.)"""
    m_frm = np.asarray(m_frm, dtype=np.float64)
    if m_frm.ndim != 2:
        raise ValueError('m_frm must be nfrms x frmlen')
    nfrms, frmlen = m_frm.shape
    v_pm_i = np.asarray(v_pm).astype(int)
    if v_pm_i.size != nfrms or nfrms < 1:
        raise ValueError('one pitch mark per frame')
    if np.any(np.diff(v_pm_i) < 0):
        raise ValueError('pitch marks must be non-decreasing')
    pm32, t0, n_out = ola_geometry(v_pm_i, frmlen)
    if win_func is not None:
        v_shift = np.diff(np.hstack((0, v_pm_i)))
        v_shift = np.append(v_shift, v_shift[-1])
        w_all, off = window_weights([win_func] * nfrms, v_shift[:-1], v_shift[1:])
        m_frm = m_frm.copy()
        for i in range(nfrms):
            v_win = np.zeros(frmlen)
            z = frmlen // 2 - int(v_shift[i])
            v_win[z:z + int(off[i + 1] - off[i])] = w_all[off[i]:off[i + 1]]     # same ValueError as the reference when it does not fit
            m_frm[i] *= v_win
    m_frm = np.ascontiguousarray(m_frm)
    out = np.empty(n_out, dtype=np.float64)
    _lib.check(_lib.lib().mpb_ola_host(_lib.ctx(), _lib.ptr(m_frm), _lib.ptr(pm32), nfrms, frmlen, int(t0), _lib.ptr(out), n_out))
    return out


def _synthesis_lossless_call(l_feats, l_pm_int, l_t0, l_nout, fft_len, compute=None, out_dtype=np.float64):
    compute = SYNTHESIS_COMPUTE if compute is None else compute
    out_dtype = np.dtype(out_dtype)
    if out_dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
        raise ValueError('out_dtype must be float64 or float32')
    feat_np = np.float32 if all(np.asarray(f[i]).dtype == np.float32 for f in l_feats for i in range(3)) else np.float64
    n_utt = len(l_feats)
    frm_off = np.zeros(n_utt + 1, dtype=np.int64)
    out_off = np.zeros(n_utt + 1, dtype=np.int64)
    for u in range(n_utt):
        frm_off[u + 1] = frm_off[u] + l_pm_int[u].size
        out_off[u + 1] = out_off[u] + l_nout[u]
    # zero-copy when the rows already sit back to back (the blocks analysis_lossless_batch returns live in pinned memory)
    mag, real, imag = (_stack_rows([f[i] for f in l_feats], feat_np) for i in range(3))
    H = fft_len // 2 + 1
    if mag.shape[1] != H or real.shape != mag.shape or imag.shape != mag.shape:
        raise ValueError('feature matrices must be nfrms x %d' % H)
    pm = np.ascontiguousarray(np.concatenate(l_pm_int), dtype=np.int32)
    t0 = np.ascontiguousarray(l_t0, dtype=np.int32)
    out = _lib.pinned.empty(int(out_off[-1]), dtype=out_dtype)
    code = lambda dt: MPB_F32 if np.dtype(dt) == np.dtype(np.float32) else MPB_F64
    _lib.check(_lib.lib().mpb_synthesis_lossless_host2(
        _lib.ctx(), _lib.ptr(mag), _lib.ptr(real), _lib.ptr(imag), code(feat_np), _lib.ptr(pm), pm.size, _lib.ptr(frm_off),
        _lib.ptr(out_off), _lib.ptr(t0), n_utt, fft_len, compute, _lib.ptr(out), code(out_dtype), out.size))
    return [out[out_off[u]:out_off[u + 1]] for u in range(n_utt)]


def synthesis_from_lossless(m_mag, m_real, m_imag, v_f0, fs):
    """src/magphase.py:1759-1776"""
    return synthesis_from_lossless_batch([(m_mag, m_real, m_imag, v_f0)], fs)[0]


def synthesis_from_lossless_batch(l_feats, fs, out_dtype=np.float64):
    """Batched synthesis_from_lossless; l_feats is a list of (m_mag, m_real, m_imag, v_f0).  Feature matrices that are
    ALL float32 cross PCIe as float32; ``out_dtype=np.float32`` returns float32 waveforms."""
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
    return _synthesis_lossless_call(l_feats, l_pm, l_t0, l_n, fft_len, out_dtype=out_dtype)


# ----------------------------------------------------------------------------------------------
# low-dimensional compression (analysis side)
# ----------------------------------------------------------------------------------------------
def warped_axis(alpha, nbins):
    """All-pass warped frequency axis on [0, pi] (src/libaudio.py:611-613, :711-718)."""
    w = np.linspace(0, np.pi, num=nbins)
    with np.errstate(divide='ignore', invalid='ignore'):
        wt = np.arctan((1 - alpha ** 2) * np.sin(w) / ((1 + alpha ** 2) * np.cos(w) - 2 * alpha))
    wt[wt < 0] += np.pi
    return wt


def build_mel_curve(alpha, nbins, amp=np.pi):
    """src/libaudio.py:711-718"""
    return warped_axis(alpha, nbins) * (amp / np.pi)


def get_num_full_mel_coeffs_from_num_phase_coeffs(freq_hz, phase_dim, alpha, fs):
    """src/magphase.py:2479-2487"""
    w = 2 * np.pi * freq_hz / float(fs)
    m = np.arctan((1 - alpha ** 2) * np.sin(w) / ((1 + alpha ** 2) * np.cos(w) - 2 * alpha))
    if m < 0:
        m += np.pi
    return int(round_to_int(1 + (np.pi * (phase_dim - 1) / float(m))))


def medfilt3(v):
    """scipy.signal.medfilt(v) (kernel 3, zero padded) without scipy: the median of three picks an element, so
    this is bit-exact.  Used by format_for_modelling (src/magphase.py:2500)."""
    v = np.asarray(v, dtype=np.float64)
    p = np.concatenate(([0.0], v, [0.0]))
    return np.median(np.stack((p[:-2], p[1:-1], p[2:])), axis=0)


class _MelPlan:
    _cache = {}

    def __init__(self, fft_len, alpha_mag, mag_dim, alpha_ph, nmel, phase_dim):
        self.key = (fft_len, alpha_mag, mag_dim, alpha_ph, nmel, phase_dim)
        self.mag_dim, self.nmel, self.phase_dim, self.fft_len = mag_dim, nmel, phase_dim, fft_len
        # cosine matrices of la.mcep_to_sp_cosmat(alpha=0.0)  (src/libaudio.py:605-631)
        cos_mag = np.ascontiguousarray(np.cos(np.arange(mag_dim)[:, None] * warped_axis(0.0, mag_dim)[None, :]))
        cos_ph = np.ascontiguousarray(
            np.cos(np.arange(nmel)[:, None] * warped_axis(0.0, nmel)[None, :])[:, :phase_dim])
        h = C.c_void_p()
        _lib.check(_lib.lib().mpb_mel_create(_lib.ctx(), fft_len, float(alpha_mag), mag_dim, float(alpha_ph), nmel,
                                             phase_dim, _lib.ptr(cos_mag), _lib.ptr(cos_ph), C.byref(h)))
        self.handle = h

    @classmethod
    def get(cls, fs, fft_len, mag_dim, phase_dim, alpha_phase):
        alpha = define_alpha(fs)
        crsf_cf, _ = define_crossfade_params(fs)
        if alpha_phase is None:
            alpha_phase = alpha
        nmel = get_num_full_mel_coeffs_from_num_phase_coeffs(crsf_cf, phase_dim, alpha_phase, fs)
        # the reference prints alpha with "%1.2f" on the SPTK command line (src/libaudio.py:589): False -> 0.00
        a_mag, a_ph = float("%1.2f" % alpha), float("%1.2f" % alpha_phase)
        key = (_lib.default_device(), _lib.current_slot(), fft_len, a_mag, mag_dim, a_ph, nmel, phase_dim)
        if key not in cls._cache:
            cls._cache[key] = cls(fft_len, a_mag, mag_dim, a_ph, nmel, phase_dim)
        return cls._cache[key]


def _lf0_smoothed(v_f0):
    """src/magphase.py:2499-2501"""
    v_f0 = np.asarray(v_f0, dtype=np.float64)
    v_voi = (v_f0 > 0).astype('float')
    return v_voi, f0_to_lf0(v_voi * medfilt3(v_f0))


def format_for_modelling(m_mag, m_real, m_imag, v_f0, fs, mag_dim=60, phase_dim=45, b_mag_fbank_mel=False,
                         alpha_phase=None):
    """Lossless features -> (m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0_smth).  src/magphase.py:2490-2544"""
    if b_mag_fbank_mel:
        raise ValueError('b_mag_fbank_mel=True (experimental filter-bank warping) is outside the CUDA hot path')
    m_mag = np.ascontiguousarray(m_mag, dtype=np.float64)
    m_real = np.ascontiguousarray(m_real, dtype=np.float64)
    m_imag = np.ascontiguousarray(m_imag, dtype=np.float64)
    fft_len = 2 * (m_mag.shape[1] - 1)
    plan = _MelPlan.get(fs, fft_len, mag_dim, phase_dim, alpha_phase)
    v_voi, v_lf0 = _lf0_smoothed(v_f0)
    n = m_mag.shape[0]
    voi8 = np.ascontiguousarray(v_voi > 0, dtype=np.uint8)
    o_mag = np.empty((n, mag_dim)); o_real = np.empty((n, phase_dim)); o_imag = np.empty((n, phase_dim))
    _lib.check(_lib.lib().mpb_mel_compress_host(plan.handle, _lib.ptr(m_mag), _lib.ptr(m_real), _lib.ptr(m_imag),
                                                _lib.ptr(voi8), n, _lib.ptr(o_mag), _lib.ptr(o_real), _lib.ptr(o_imag)))
    return o_mag, o_real, o_imag, v_lf0


def interp_from_variable_to_const_frm_rate(m_data, v_pm_smpls, const_rate_ms, fs, interp_type='linear'):
    """Linear resampling of per-frame data onto a constant frame-rate grid.  src/magphase.py:2219-2239.
    Host version, used for the O(n) f0 / voicing tracks only; the feature matrices are interpolated on the device
    (const_rate_rows + the tile-product loader)."""
    if interp_type != 'linear':
        raise ValueError('only linear interpolation is supported')
    m = np.asarray(m_data, dtype=np.float64)
    one_d = m.ndim == 1
    if one_d:
        m = m[:, None]
    step = fs * const_rate_ms / 1000
    centres = np.arange(step, v_pm_smpls[-1], step)
    x = np.asarray(v_pm_smpls, dtype=np.float64)
    if x[0] > 0:
        x = np.r_[0, x]
        m = np.vstack((m[0, :], m))
    # scipy's interp1d is what the reference calls: same arithmetic, bit-identical f0 / voicing tracks
    from scipy import interpolate
    out = interpolate.interp1d(x, m, axis=0, kind='linear')(centres)
    return out[:, 0] if one_d else out


def analysis_compressed_from_pm(v_sig, fs, v_pm_smpls, v_voi, fft_len=None, mag_dim=60, phase_dim=10,
                                b_const_rate=False, b_mag_fbank_mel=False, alpha_phase=None):
    """analysis_compressed (src/magphase.py:2947-2988) minus file reading and epoch detection.
    Variable-rate output runs fused on the device (the lossless features never leave HBM)."""
    return analysis_compressed_batch([v_sig], fs, [v_pm_smpls], [v_voi], fft_len=fft_len, mag_dim=mag_dim,
                                     phase_dim=phase_dim, b_const_rate=b_const_rate,
                                     b_mag_fbank_mel=b_mag_fbank_mel, alpha_phase=alpha_phase)[0]


def analysis_compressed_batch(l_sig, fs, l_pm_smpls, l_voi, fft_len=None, mag_dim=60, phase_dim=10,
                              b_const_rate=False, b_mag_fbank_mel=False, alpha_phase=None, out_dtype=np.float64):
    """Batched analysis_compressed_from_pm.  Returns a list of
    (m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0_smth, v_shift, fs, fft_len).

    Element types on request (the defaults reproduce the reference: float64 in, float64 out): signals may be int16 (PCM16
    exactly as the wav file holds it -- the device applies sf.read's 1/32768, src/libaudio.py:343-350) or float32 arrays,
    and ``out_dtype=np.float32`` returns the three feature matrices in float32, the precision of the reference's own
    feature files (src/libutils.py:122-127).  Nothing is then narrowed or widened on the host and half (features) to a
    quarter (PCM16) of the bytes cross PCIe.  lf0 and the shifts are host bookkeeping and stay float64 / int."""
    if b_mag_fbank_mel:
        raise ValueError('b_mag_fbank_mel=True (experimental filter-bank warping) is outside the CUDA hot path')
    if fft_len is None:
        fft_len = define_fft_len(fs)
    if b_const_rate:
        return _analysis_compressed_const_rate(l_sig, fs, l_pm_smpls, l_voi, fft_len, mag_dim, phase_dim, alpha_phase)
    plan = _MelPlan.get(fs, fft_len, mag_dim, phase_dim, alpha_phase)
    sizes = [np.size(x) for x in l_sig]
    centre, left, right, voi8, f0_med, frm_off = _analysis_geometry_c(l_pm_smpls, sizes, l_voi, fs)
    _check_frames(left, right, fft_len)
    lf0_all = f0_to_lf0(f0_med)                                               # np.log: a fresh array
    shift_all = left.astype(int)                                              # one copy out of the reusable workspace
    lefts = [shift_all[frm_off[u]:frm_off[u + 1]] for u in range(len(l_sig))]
    lf0s = [lf0_all[frm_off[u]:frm_off[u + 1]] for u in range(len(l_sig))]
    out_dtype = np.dtype(out_dtype)
    if out_dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
        raise ValueError('out_dtype must be float64 or float32')
    # one element type for the whole batch: int16 / float32 as given, anything else as float64 (no copy for float64 arrays)
    kinds = {np.asarray(s).dtype for s in l_sig}
    sig_np = kinds.pop() if len(kinds) == 1 and next(iter(kinds)) in (np.dtype(np.int16), np.dtype(np.float32)) else np.dtype(np.float64)
    sig_code = {np.dtype(np.float64): _lib.MPB_F64, np.dtype(np.float32): _lib.MPB_F32, np.dtype(np.int16): _lib.MPB_I16}[sig_np]
    sigs = [np.ascontiguousarray(s, dtype=sig_np) for s in l_sig]
    sig_ptrs = (C.c_void_p * len(sigs))(*[s.ctypes.data for s in sigs])
    sig_lens = np.ascontiguousarray([s.size for s in sigs], dtype=np.int64)
    n = centre.size
    o_mag, o_real, o_imag = (_lib.pinned.empty((n, d), dtype=out_dtype) for d in (mag_dim, phase_dim, phase_dim))
    with _lib.device_gate():
        _lib.check(_lib.lib().mpb_analysis_compressed_hostv2(
            plan.handle, sig_ptrs, sig_code, _lib.ptr(sig_lens), len(sigs), _lib.ptr(centre), _lib.ptr(left), _lib.ptr(right),
            _lib.ptr(voi8), n, _lib.ptr(o_mag), _lib.ptr(o_real), _lib.ptr(o_imag),
            _lib.MPB_F32 if out_dtype == np.dtype(np.float32) else _lib.MPB_F64))
    out, a = [], 0
    for u in range(len(l_sig)):
        b = a + lefts[u].size
        out.append((o_mag[a:b], o_real[a:b], o_imag[a:b], lf0s[u], lefts[u], fs, fft_len))
        a = b
    return out


def const_rate_rows(v_pm_smpls, const_rate_ms, fs):
    """Index form of interp_from_variable_to_const_frm_rate (src/magphase.py:2219-2239, linear): for every constant-rate
    centre the two source frames and the weight of the second one.  Frame 0 is replicated at t = 0 when pm[0] > 0."""
    step = fs * const_rate_ms / 1000
    centres = np.arange(step, v_pm_smpls[-1], step)
    x = np.asarray(v_pm_smpls, dtype=np.float64)
    shift = 0
    if x[0] > 0:
        x = np.r_[0, x]
        shift = 1
    j = np.clip(np.searchsorted(x, centres, side='right') - 1, 0, x.size - 2)
    w = (centres - x[j]) / (x[j + 1] - x[j])
    r0 = np.maximum(j - shift, 0)
    r1 = np.maximum(j + 1 - shift, 0)
    return r0.astype(np.int32), r1.astype(np.int32), w


def _analysis_compressed_const_rate(l_sig, fs, l_pm_smpls, l_voi, fft_len, mag_dim, phase_dim, alpha_phase):
    """Constant-rate branch of analysis_compressed (src/magphase.py:2966-2983) on the device."""
    plan = _MelPlan.get(fs, fft_len, mag_dim, phase_dim, alpha_phase)
    sig_off, frm_off = 0, 0
    centres, lefts, rights, r0s, r1s, ws, vois, lf0s, shifts, n_outs = [], [], [], [], [], [], [], [], [], []
    for sig, pm, voi in zip(l_sig, l_pm_smpls, l_voi):
        sig = np.asarray(sig)
        P, v_shift, v_rights = frame_geometry(pm, sig.size)
        _check_frames(v_shift, v_rights, fft_len)
        v_f0 = shift_to_f0(v_shift.astype(int), np.asarray(voi, dtype=np.float64), fs, out='f0', b_smooth=False)
        v_pm = np.cumsum(v_shift)                                   # la.shift_to_pm (:2970)
        r0, r1, w = const_rate_rows(v_pm, 5.0, fs)
        vv = v_f0 > 1.0
        v_f0c = interp_from_variable_to_const_frm_rate(np.r_[v_f0[vv][0], v_f0[vv], v_f0[vv][-1]],
                                                       np.r_[0, v_pm[vv], v_pm[-1]], 5.0, fs)
        vvc = interp_from_variable_to_const_frm_rate(vv.astype(float), v_pm, 5.0, fs) > 0.5
        v_voi_c, v_lf0 = _lf0_smoothed(v_f0c * vvc)
        centres.append(P[1:-1] + sig_off); lefts.append(v_shift); rights.append(v_rights)
        r0s.append(r0 + frm_off); r1s.append(r1 + frm_off); ws.append(w)
        vois.append(v_voi_c > 0); lf0s.append(v_lf0); shifts.append(v_shift.astype(int)); n_outs.append(r0.size)
        sig_off += sig.size
        frm_off += v_shift.size
    cat = lambda l, dt: np.ascontiguousarray(np.concatenate(l), dtype=dt)
    centre, left, right = cat(centres, np.int64), cat(lefts, np.int32), cat(rights, np.int32)
    lr0, lr1, lw, voi8 = cat(r0s, np.int32), cat(r1s, np.int32), cat(ws, np.float32), cat(vois, np.uint8)
    sigs = [np.ascontiguousarray(s, dtype=np.float64) for s in l_sig]
    sig_ptrs = (C.c_void_p * len(sigs))(*[s.ctypes.data for s in sigs])
    sig_lens = np.ascontiguousarray([s.size for s in sigs], dtype=np.int64)
    n = int(sum(n_outs))
    o_mag = np.empty((n, mag_dim)); o_real = np.empty((n, phase_dim)); o_imag = np.empty((n, phase_dim))
    _lib.check(_lib.lib().mpb_analysis_compressed_const_hostv(
        plan.handle, sig_ptrs, _lib.ptr(sig_lens), len(sigs), _lib.ptr(centre), _lib.ptr(left), _lib.ptr(right),
        centre.size, _lib.ptr(lr0), _lib.ptr(lr1), _lib.ptr(lw), _lib.ptr(voi8), n, _lib.ptr(o_mag), _lib.ptr(o_real),
        _lib.ptr(o_imag)))
    out, a = [], 0
    for u in range(len(l_sig)):
        b = a + n_outs[u]
        out.append((o_mag[a:b], o_real[a:b], o_imag[a:b], lf0s[u], shifts[u], fs, fft_len))
        a = b
    return out


def analysis_compressed(wav_file, fft_len=None, mag_dim=60, phase_dim=10, b_const_rate=False, b_mag_fbank_mel=False,
                        alpha_phase=None, est_file=None, pm=None):
    """src/magphase.py:2947-2988.  Returns (m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0_smth, v_shift, fs, fft_len)."""
    v_sig, fs = io.read_audio_file(wav_file)
    v_pm_sec, v_voi = get_pitch_marks_and_voicing(wav_file, len(v_sig), fs, est_file=est_file, pm=pm)
    return analysis_compressed_from_pm(v_sig, fs, v_pm_sec * fs, v_voi, fft_len=fft_len, mag_dim=mag_dim,
                                       phase_dim=phase_dim, b_const_rate=b_const_rate,
                                       b_mag_fbank_mel=b_mag_fbank_mel, alpha_phase=alpha_phase)


def analysis_for_acoustic_modelling(wav_file, out_dir, fft_len=None, mag_dim=60, phase_dim=10, b_const_rate=False,
                                    b_mag_fbank_mel=False, alpha_phase=None, est_file=None, pm=None):
    """Writes .mag/.real/.imag/.lf0 (+ .shift when variable rate) float32 files.  src/magphase.py:2992-3022.
    NB the reference passes ``alpha_phase=b_mag_fbank_mel`` (=False, i.e. 0.0) at :3010 -- replicated."""
    feats = analysis_compressed(wav_file, fft_len=fft_len, mag_dim=mag_dim, phase_dim=phase_dim,
                                b_const_rate=b_const_rate, b_mag_fbank_mel=b_mag_fbank_mel,
                                alpha_phase=b_mag_fbank_mel, est_file=est_file, pm=pm)
    m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0_smth, v_shift = feats[:5]
    file_id = os.path.basename(wav_file).split(".")[0]
    io.write_binfile(m_mag_mel_log, os.path.join(out_dir, file_id + '.mag'))
    io.write_binfile(m_real_mel, os.path.join(out_dir, file_id + '.real'))
    io.write_binfile(m_imag_mel, os.path.join(out_dir, file_id + '.imag'))
    io.write_binfile(v_lf0_smth, os.path.join(out_dir, file_id + '.lf0'))
    if not b_const_rate:
        io.write_binfile(v_shift, os.path.join(out_dir, file_id + '.shift'))
    return


# ----------------------------------------------------------------------------------------------
# compressed synthesis
# ----------------------------------------------------------------------------------------------
def mel_unwarp_matrix(n_c, nbins_out, alpha):
    """la.sp_mel_unwarp (src/libaudio.py:667-684) as one matrix: out = in @ U, U is n_c x nbins_out.
    Hermitian-extend, ifft.real, double cepstral indices 1..n_c-3 (index n_c-2 is NOT doubled, :679),
    cosine matrix of the warped axis (src/libaudio.py:605-631)."""
    eye = np.eye(n_c)
    ceps = np.fft.ifft(np.hstack((eye, eye[:, -2:0:-1])), axis=1).real
    ceps[:, 1:(n_c - 2)] *= 2
    T = np.cos(np.arange(n_c)[:, None] * warped_axis(alpha, nbins_out)[None, :])
    return ceps[:, :n_c] @ T


def crossfade_curve(nbins, cut_off, bw, fs):
    """Left weight of la.spectral_crossfade (src/libaudio.py:160-186) and the last bin it can be non-zero at."""
    n_fft = (nbins - 1) * 2
    bin_l = int(round_to_int((cut_off - bw / 2.0) * n_fft / float(fs)))
    bin_r = int(round_to_int((cut_off + bw / 2.0) * n_fft / float(fs)))
    B = bin_r - bin_l
    return np.hstack((np.ones(bin_l), np.hanning(2 * B + 1)[B:], np.zeros(nbins - bin_r - 1))), bin_r


class _SynPlan:
    _cache = {}

    def __init__(self, fs, fft_len, mag_dim, phase_dim, alpha, alpha_phase):
        H = fft_len // 2 + 1
        crsf_cf, crsf_bw = define_crossfade_params(fs)
        curve, bin_r = crossfade_curve(H, crsf_cf, crsf_bw, fs)
        # bins at and above the crossfade's upper edge carry no periodic part: the mask reaches exactly 0 at bin_r
        # (hann(2B+1)[-1] == 0.0), so the phase rows are only needed for bins < bin_r
        self.HB = bin_r if curve[bin_r] == 0.0 else bin_r + 1
        u_mag = mel_unwarp_matrix(mag_dim, H, alpha)
        nmel = get_num_full_mel_coeffs_from_num_phase_coeffs(crsf_cf, phase_dim, alpha_phase, fs)
        u_full = mel_unwarp_matrix(nmel, H, alpha_phase)[:, :self.HB]
        # nearest-extrapolation padding phase_dim -> nmel = repeat the last column (src/magphase.py:1225-1229)
        u_ph = np.zeros((phase_dim, self.HB))
        np.add.at(u_ph, np.minimum(np.arange(nmel), phase_dim - 1), u_full)
        tab = np.vstack((np.sqrt(curve) * 10 ** (build_mel_curve(0.6, H, amp=2.0) / 20),          # :940-946
                         np.sqrt(1 - curve),                                                      # :947
                         10 ** ((build_mel_curve(alpha, H, amp=3.5) - 3.5) / 20)))                # :917-918
        self.fft_len, self.mag_dim, self.phase_dim = fft_len, mag_dim, phase_dim
        h = C.c_void_p()
        _lib.check(_lib.lib().mpb_syn_create(_lib.ctx(), fft_len, mag_dim, phase_dim, self.HB,
                                             _lib.ptr(np.ascontiguousarray(u_mag)), _lib.ptr(np.ascontiguousarray(u_ph)),
                                             _lib.ptr(np.ascontiguousarray(tab)), C.byref(h)))
        self.handle = h

    @classmethod
    def get(cls, fs, fft_len, mag_dim, phase_dim, alpha_phase):
        alpha = define_alpha(fs)
        if alpha_phase is None:
            alpha_phase = alpha
        key = (_lib.default_device(), _lib.current_slot(), fs, fft_len, mag_dim, phase_dim, float(alpha_phase))
        if key not in cls._cache:
            cls._cache[key] = cls(fs, fft_len, mag_dim, phase_dim, alpha, alpha_phase)
        return cls._cache[key]


def get_shifts_and_frm_locs_from_const_shifts(v_shift_c_rate, frm_rate_ms, fs, interp_type='linear'):
    """Host reverse scan (sequential, data dependent): walk back from the last constant-rate centre, subtracting
    the interpolated shift, until the position leaves the interpolation range.  src/magphase.py:1426-1449"""
    n = np.size(v_shift_c_rate, 0)
    centres = (fs * frm_rate_ms / 1000) * np.arange(1, n + 1)
    shifts, locs = [], []
    pos = centres[-1]
    for _ in range(2 * n - 1):
        if pos < centres[0] or pos > centres[-1]:
            break
        s = float(np.interp(pos, centres, v_shift_c_rate))
        locs.append(pos)
        shifts.append(s)
        pos = pos - s
    return np.array(shifts[::-1]), np.array(locs[::-1])


def const_rate_scan_batch(l_shift_c, frm_rate_ms, fs):
    """get_shifts_and_frm_locs_from_const_shifts for a list of utterances in one C pass (mpb_const_rate_scan: np.interp's
    arithmetic bit for bit, tests/test_batch_geometry_cpu.py).  Returns a list of (v_shift, v_locs)."""
    lens = np.array([np.size(v) for v in l_shift_c], dtype=np.int64)
    off = _seg_offsets(lens)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in l_shift_c])) if len(l_shift_c) else np.zeros(0)
    o_s, o_l = np.empty(2 * flat.size + 2), np.empty(2 * flat.size + 2)
    cnt = np.zeros(len(l_shift_c) + 1, dtype=np.int64)
    _lib.check(_lib.lib().mpb_const_rate_scan(_lib.ptr(flat), _lib.ptr(off), len(l_shift_c), float(fs * frm_rate_ms / 1000),
                                              _lib.ptr(o_s), _lib.ptr(o_l), _lib.ptr(cnt)))
    out = []
    for u in range(len(l_shift_c)):
        a, c = 2 * int(off[u]), int(cnt[u])
        out.append((o_s[a:a + c][::-1].copy(), o_l[a:a + c][::-1].copy()))
    return out


def _const_rate_rows(v_locs, n_c, step):
    """Row pairs + weights of interp_from_const_to_variable_rate (src/magphase.py:2242-2252), linear."""
    centres = step * np.arange(1, n_c + 1)
    j = np.clip(np.searchsorted(centres, v_locs, side='right') - 1, 0, n_c - 2)
    w = (v_locs - centres[j]) / (centres[j + 1] - centres[j])
    return j.astype(np.int32), (j + 1).astype(np.int32), w


def output_hpf_coefficients(fs):
    """(b, a) of the output high-pass: scipy.signal.butter(4, 40 Hz / (fs/2), 'highpass').  src/magphase.py:981-995.
    Host constant table; the filter itself runs on the device (blocked state-space scan, mpb_iir4_*)."""
    from scipy import signal
    v_b, v_a = signal.butter(4, 40 / (fs / 2.0), btype='highpass')
    return np.ascontiguousarray(v_b, dtype=np.float64), np.ascontiguousarray(v_a, dtype=np.float64)


def output_hpf_sos(fs):
    """The reference's (b, a) factored into two biquads (scipy sos layout) for the device scan: poles = np.roots(a)
    paired by conjugates, zeros = the exact quadruple zero at z = 1 of a Butterworth high-pass, gain b[0] in the first
    section.  The product of the sections reproduces (b, a) to ~1e-15."""
    v_b, v_a = output_hpf_coefficients(fs)
    poles = np.roots(v_a)
    upper = sorted([p for p in poles if p.imag > 0], key=lambda p: p.imag)
    if len(upper) != 2 or not np.allclose(v_b / v_b[0], [1, -4, 6, -4, 1], rtol=0, atol=1e-9):
        raise ValueError('unexpected high-pass design')
    sos = np.array([[v_b[0], -2 * v_b[0], v_b[0], 1.0, -2 * upper[0].real, abs(upper[0]) ** 2],
                    [1.0, -2.0, 1.0, 1.0, -2 * upper[1].real, abs(upper[1]) ** 2]])
    return np.ascontiguousarray(sos, dtype=np.float64)


def synthesis_from_compressed(m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0, fs, fft_len=None, b_voi_ap_win=True,
                              b_fbank_mel=False, b_const_rate=False, per_phase_type='magphase', alpha_phase=None,
                              b_out_hpf=True):
    """src/magphase.py:825-997.  The aperiodic noise is drawn from NumPy's global legacy stream with
    np.random.uniform(-1, 1, ns_len) exactly where the reference draws it (:883): seed it for reproducibility."""
    return synthesis_from_compressed_batch([(m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0)], fs, fft_len=fft_len,
                                           b_voi_ap_win=b_voi_ap_win, b_fbank_mel=b_fbank_mel,
                                           b_const_rate=b_const_rate, per_phase_type=per_phase_type,
                                           alpha_phase=alpha_phase, b_out_hpf=b_out_hpf)[0]


def _stack_rows(l_arr, dtype=np.float64, copy=True):
    """np.concatenate(l_arr, axis=0) as `dtype` -- without the copy when the arrays already sit back to back in memory
    (the row blocks the *_batch analysis functions return).  The result is only used as a read-only argument of a C
    call made while the inputs are still referenced, so viewing across the blocks is safe.  copy=False: None instead of
    a stacked copy when the blocks are not adjacent."""
    a0 = l_arr[0]
    if all(isinstance(a, np.ndarray) and a.dtype == np.dtype(dtype) and a.ndim == 2 and a.flags['C_CONTIGUOUS'] and
           a.shape[1] == a0.shape[1] for a in l_arr):
        ptr = a0.ctypes.data
        for a in l_arr:
            if a.ctypes.data != ptr:
                break
            ptr += a.nbytes
        else:
            rows = sum(a.shape[0] for a in l_arr)
            return np.lib.stride_tricks.as_strided(a0, shape=(rows, a0.shape[1]), strides=a0.strides, writeable=False)
    if not copy:
        return None
    return np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=dtype) for a in l_arr], axis=0))


def compressed_synthesis_geometry(l_lf0, l_nrows, fs, fft_len, b_voi_ap_win=True, b_const_rate=False, _reuse=False):
    """Host bookkeeping of synthesis_from_compressed for a batch (src/magphase.py:846-848, 861-870, 879-882,
    886-896, 968-971, 34-62): everything integer / float64 that the kernels consume as arrays.
    Returns (dict of C-contiguous arrays for mpb_syn_frames + 'need_ph', list of per-utterance noise lengths).
    Variable-rate batches take the vectorised path; the per-utterance loop below is the definition (and the
    constant-rate path, whose reverse scan is sequential anyway)."""
    if not b_const_rate:
        return _compressed_synthesis_geometry_flat(l_lf0, l_nrows, fs, fft_len, b_voi_ap_win, reuse=_reuse)
    return _compressed_synthesis_geometry_const_flat(l_lf0, l_nrows, fs, fft_len, b_voi_ap_win)


def _compressed_synthesis_geometry_flat(l_lf0, l_nrows, fs, fft_len, b_voi_ap_win=True, reuse=False):
    """compressed_synthesis_geometry(b_const_rate=False) for all utterances at once.  exp() and fs / f0 run in NumPy on
    the concatenated lf0 (elementwise: the same numbers as per utterance); everything after the truncation to integer
    shifts is one C pass (mpb_syn_geometry).  Integer results identical to the loop (tests/test_batch_geometry_cpu.py).
    reuse=True hands out arrays of the module workspace (valid until the next call)."""
    n_utt = len(l_lf0)
    lens = np.array([np.size(v) for v in l_lf0], dtype=np.int64)
    if np.any(lens != np.asarray(l_nrows, dtype=np.int64)):
        raise ValueError('lf0 length must equal the number of feature rows')
    if np.any(lens < 2):
        raise IndexError('synthesis_from_compressed needs at least two frames (src/magphase.py:882)')
    off = _seg_offsets(lens)
    n = int(off[-1])
    v_f0 = np.exp(np.concatenate([np.asarray(v, dtype=np.float64) for v in l_lf0]))
    voi8 = (v_f0 > 1.0).astype(np.uint8)                          # :847
    v_f0[v_f0 == 0] = 200.0                                       # f0_to_shift (:2210-2215), in place on our own array
    shift = (fs / v_f0).astype(np.int64)                          # truncation BEFORE the cumsum (:879-880)
    new = (lambda name, k, dt: _workspace('s_' + name, k, dt)) if reuse else (lambda name, k, dt: np.empty(k, dtype=dt))
    pm, ncentre = new('pm', n, np.int32), new('ncentre', n, np.int64)
    nleft, nright, nkind = new('nleft', n, np.int32), new('nright', n, np.int32), new('nkind', n, np.uint8)
    win_a, win_b, row0 = new('win_a', n, np.int32), new('win_b', n, np.int32), new('row0', n, np.int32)
    out_off, t0, ns_len = np.empty(n_utt + 1, dtype=np.int64), np.empty(n_utt, dtype=np.int32), np.empty(n_utt, dtype=np.int64)
    _lib.check(_lib.lib().mpb_syn_geometry(_lib.ptr(shift), _lib.ptr(voi8), _lib.ptr(off), n_utt, fft_len,
                                           1 if b_voi_ap_win else 0, _lib.ptr(pm), _lib.ptr(ncentre), _lib.ptr(nleft),
                                           _lib.ptr(nright), _lib.ptr(nkind), _lib.ptr(win_a), _lib.ptr(win_b), _lib.ptr(row0),
                                           _lib.ptr(out_off), _lib.ptr(t0), _lib.ptr(ns_len)))
    arrs = dict(pm=pm, ncentre=ncentre, nleft=nleft, nright=nright, voi=voi8, nkind=nkind, win_a=win_a, win_b=win_b,
                row0=row0, row1=None, roww=None, utt_frm_off=off, utt_out_off=out_off, utt_t0=t0, need_ph=voi8)
    return arrs, [int(x) for x in ns_len]


def _compressed_synthesis_geometry_const_flat(l_lf0, l_nrows, fs, fft_len, b_voi_ap_win=True):
    """compressed_synthesis_geometry(b_const_rate=True) for all utterances at once: the reverse scan in C
    (mpb_const_rate_scan), the row pairs / weights and the interpolated voicing (src/magphase.py:861-870) vectorised over the
    concatenated frames, everything after the truncation to integer shifts in mpb_syn_geometry.  Bit-identical to the loop
    below (tests/test_batch_geometry_cpu.py), which stays the definition."""
    n_utt = len(l_lf0)
    n_c = np.array([np.size(v) for v in l_lf0], dtype=np.int64)
    if np.any(n_c != np.asarray(l_nrows, dtype=np.int64)):
        raise ValueError('lf0 length must equal the number of feature rows')
    row_off = _seg_offsets(n_c)
    n_rows = int(row_off[-1])
    v_f0 = np.exp(np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in l_lf0])) if n_utt else np.zeros(0)
    vf = (v_f0 > 1.0).astype(np.float64)                          # :847
    shift_c = f0_to_shift(v_f0, fs)
    step = float(fs * 5.0 / 1000)
    o_s, o_l = np.empty(2 * n_rows + 2), np.empty(2 * n_rows + 2)
    cnt = np.zeros(n_utt + 1, dtype=np.int64)
    _lib.check(_lib.lib().mpb_const_rate_scan(_lib.ptr(np.ascontiguousarray(shift_c)), _lib.ptr(row_off), n_utt, step,
                                              _lib.ptr(o_s), _lib.ptr(o_l), _lib.ptr(cnt)))
    cnt = cnt[:n_utt]
    if np.any(cnt < 2):
        raise IndexError('synthesis_from_compressed needs at least two frames (src/magphase.py:882)')
    frm_off = _seg_offsets(cnt)
    n = int(frm_off[-1])
    # frame k of utterance u (forward order) is entry cnt[u] - 1 - k of its scan, which starts at 2 * row_off[u]
    k_in = np.arange(n, dtype=np.int64) - np.repeat(frm_off[:-1], cnt)
    idx = np.repeat(2 * row_off[:-1] + cnt - 1, cnt) - k_in
    v_shift, v_locs = o_s[idx], o_l[idx]
    # rows (src/magphase.py:2242-2252): j = clip(#{centres <= loc} - 1, 0, n_c - 2) with centres[k] = step * (k + 1)
    j = (v_locs / step).astype(np.int64) - 1
    for _ in range(2):                                            # the quotient can be one off in either direction
        j = np.where(step * (j + 2).astype(np.float64) <= v_locs, j + 1, j)
        j = np.where(step * (j + 1).astype(np.float64) > v_locs, j - 1, j)
    j = np.clip(j, 0, np.repeat(n_c - 2, cnt))
    c0, c1 = step * (j + 1).astype(np.float64), step * (j + 2).astype(np.float64)
    w = (v_locs - c0) / (c1 - c0)
    r0 = j + np.repeat(row_off[:-1], cnt)
    r1 = r0 + 1
    voi8 = ((vf[r0] + (vf[r1] - vf[r0]) * w) > 0.5).astype(np.uint8)                        # :868
    shift = v_shift.astype(np.int64)                              # truncation BEFORE the cumsum (:879-880)
    pm, ncentre = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
    nleft, nright, nkind = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.uint8)
    win_a, win_b, row_id = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
    out_off, t0, ns_len = np.empty(n_utt + 1, dtype=np.int64), np.empty(n_utt, dtype=np.int32), np.empty(n_utt, dtype=np.int64)
    _lib.check(_lib.lib().mpb_syn_geometry(_lib.ptr(shift), _lib.ptr(voi8), _lib.ptr(frm_off), n_utt, fft_len,
                                           1 if b_voi_ap_win else 0, _lib.ptr(pm), _lib.ptr(ncentre), _lib.ptr(nleft),
                                           _lib.ptr(nright), _lib.ptr(nkind), _lib.ptr(win_a), _lib.ptr(win_b), _lib.ptr(row_id),
                                           _lib.ptr(out_off), _lib.ptr(t0), _lib.ptr(ns_len)))
    arrs = dict(pm=pm, ncentre=ncentre, nleft=nleft, nright=nright, voi=voi8, nkind=nkind, win_a=win_a, win_b=win_b,
                row0=r0.astype(np.int32), row1=r1.astype(np.int32), roww=w.astype(np.float32), utt_frm_off=frm_off,
                utt_out_off=out_off, utt_t0=t0, need_ph=np.ones(n_rows, dtype=np.uint8))
    return arrs, [int(x) for x in ns_len]


def _compressed_synthesis_geometry_loop(l_lf0, l_nrows, fs, fft_len, b_voi_ap_win=True, b_const_rate=False):
    half = fft_len // 2
    n_utt = len(l_lf0)
    frm_off = np.zeros(n_utt + 1, dtype=np.int64)
    out_off = np.zeros(n_utt + 1, dtype=np.int64)
    row_off, noise_off = 0, 0
    acc = {k: [] for k in ('pm', 'ncentre', 'nleft', 'nright', 'voi', 'nkind', 'win_a', 'win_b', 'row0', 'row1', 'roww',
                           'need', 't0')}
    l_ns_len = []
    l_f0 = [np.exp(np.asarray(l_lf0[u], dtype=np.float64)) for u in range(n_utt)]
    for u in range(n_utt):
        if l_f0[u].size != int(l_nrows[u]):
            raise ValueError('lf0 length must equal the number of feature rows')
    # the sequential reverse scan of every utterance in one C pass
    l_scan = const_rate_scan_batch([f0_to_shift(f, fs) for f in l_f0], 5.0, fs) if b_const_rate else None
    for u in range(n_utt):
        n_c = int(l_nrows[u])
        v_f0 = l_f0[u]
        v_voi = v_f0 > 1.0                                        # :847
        v_shift = f0_to_shift(v_f0, fs) if not b_const_rate else None
        if b_const_rate:
            v_shift, v_locs = l_scan[u]
            r0, r1, w = _const_rate_rows(v_locs, n_c, fs * 5.0 / 1000)
            vf = v_voi.astype(np.float64)
            v_voi = (vf[r0] + (vf[r1] - vf[r0]) * w) > 0.5        # :868
            need = np.ones(n_c, dtype=np.uint8)
        else:
            r0, r1, w = np.arange(n_c, dtype=np.int32), None, None
            need = v_voi.astype(np.uint8)
        n = v_shift.size
        if n < 2:
            raise IndexError('synthesis_from_compressed needs at least two frames (src/magphase.py:882)')
        v_shift = v_shift.astype(int)                             # truncation BEFORE the cumsum (:879-880)
        v_pm = np.cumsum(v_shift)
        ns_len = int(v_pm[-1] + (v_pm[-1] - v_pm[-2]))
        P, n_left, n_right = frame_geometry(v_pm, ns_len)
        if np.any(n_left > half) or np.any(n_right >= half):      # frame_shift() would get a negative pad (src/libaudio.py:137-140)
            raise ValueError('negative dimensions are not allowed')
        se = np.r_[v_shift[0], v_shift, v_shift[-1], v_shift[-1]]
        pm_int, t0, n_out = ola_geometry(v_pm, fft_len)
        frm_off[u + 1] = frm_off[u] + n
        out_off[u + 1] = out_off[u] + n_out
        acc['pm'].append(pm_int); acc['ncentre'].append(P[1:-1] + noise_off)
        acc['nleft'].append(n_left); acc['nright'].append(n_right)
        acc['voi'].append(v_voi.astype(np.uint8))
        acc['nkind'].append(np.where(v_voi & bool(b_voi_ap_win), WIN_BARTLETT25, WIN_HANN).astype(np.uint8))
        acc['win_a'].append(se[:-3] + se[1:-2]); acc['win_b'].append(se[2:-1] + se[3:])
        acc['row0'].append(r0 + row_off)
        if b_const_rate:
            acc['row1'].append(r1 + row_off); acc['roww'].append(w)
        acc['need'].append(need); acc['t0'].append(t0)
        l_ns_len.append(ns_len)
        row_off += n_c
        noise_off += ns_len
    cat = lambda k, dt: np.ascontiguousarray(np.concatenate(acc[k]), dtype=dt)
    arrs = dict(pm=cat('pm', np.int32), ncentre=cat('ncentre', np.int64), nleft=cat('nleft', np.int32),
                nright=cat('nright', np.int32), voi=cat('voi', np.uint8), nkind=cat('nkind', np.uint8),
                win_a=cat('win_a', np.int32), win_b=cat('win_b', np.int32), row0=cat('row0', np.int32),
                row1=cat('row1', np.int32) if b_const_rate else None,
                roww=cat('roww', np.float32) if b_const_rate else None,
                utt_frm_off=frm_off, utt_out_off=out_off, utt_t0=np.ascontiguousarray(acc['t0'], dtype=np.int32),
                need_ph=cat('need', np.uint8))
    return arrs, l_ns_len


def synthesis_from_compressed_batch(l_feats, fs, fft_len=None, b_voi_ap_win=True, b_fbank_mel=False, b_const_rate=False,
                                    per_phase_type='magphase', alpha_phase=None, b_out_hpf=True, l_noise=None,
                                    out_dtype=np.float64, rng=None):
    """Batched synthesis_from_compressed; l_feats is a list of (m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0).
    l_noise: optional list of per-utterance noise vectors (else drawn from np.random, utterance by utterance).
    Feature matrices that are ALL float32 (what read_binfile finds on disk, src/libutils.py:112-120) cross PCIe as float32;
    ``out_dtype=np.float32`` returns float32 waveforms (the wav writer quantises to PCM16 anyway, src/libaudio.py:352-365).
    The defaults (float64 in, float64 out) reproduce the reference.
    rng: a ``np.random.RandomState`` to draw the noise from instead of NumPy's global stream (worker threads that
    process batches concurrently each bring their own; the global stream is the reference's behaviour)."""
    if b_fbank_mel:
        raise ValueError('b_fbank_mel=True (experimental filter-bank warping) is outside the CUDA hot path')
    if per_phase_type not in ('magphase', 'linear', 'min_phase'):
        raise ValueError("per_phase_type must be 'magphase', 'min_phase' or 'linear'")
    if fft_len is None:
        fft_len = define_fft_len(fs)
    mag_dim = np.shape(l_feats[0][0])[1]
    phase_dim = np.shape(l_feats[0][1])[1]
    for f in l_feats:
        if np.shape(f[0])[1] != mag_dim or np.shape(f[1])[1] != phase_dim or np.shape(f[2])[1] != phase_dim:
            raise ValueError('all utterances of a batch must share mag_dim / phase_dim')
    plan = _SynPlan.get(fs, fft_len, mag_dim, phase_dim, alpha_phase)
    n_utt = len(l_feats)
    arrs, l_ns_len = compressed_synthesis_geometry([f[3] for f in l_feats], [np.shape(f[0])[0] for f in l_feats], fs,
                                                   fft_len, b_voi_ap_win=b_voi_ap_win, b_const_rate=b_const_rate,
                                                   _reuse=True)
    mt_key, mt_pos, np_state = None, None, None
    if l_noise is None:
        # np.random.uniform(-1, 1, ns_len) per utterance (:883) == one run of sum(ns_len) draws on NumPy's global
        # legacy stream.  The stream is advanced ON THE DEVICE, bit for bit, and handed back to NumPy afterwards.
        np_state = rng.get_state() if rng is not None else np.random.get_state()
        if np_state[0] == 'MT19937':
            mt_key = np.ascontiguousarray(np_state[1], dtype=np.uint32).copy()
            mt_pos = C.c_int32(int(np_state[2]))
        else:
            l_noise = [(rng if rng is not None else np.random).uniform(-1, 1, n) for n in l_ns_len]
    if l_noise is not None:
        for v, n in zip(l_noise, l_ns_len):
            if np.size(v) != n:
                raise ValueError('noise length %d != %d' % (np.size(v), n))
    need = arrs.pop('need_ph')
    if per_phase_type == 'min_phase':
        need = np.zeros_like(need)          # phase rows come from the minimum-phase kernel instead
    out_off = arrs['utt_out_off']
    fr = _lib.SynFrames(nfrm=int(arrs['utt_frm_off'][-1]), n_utt=n_utt, **{k: _lib.ptr(v) for k, v in arrs.items()})
    out_dtype = np.dtype(out_dtype)
    if out_dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
        raise ValueError('out_dtype must be float64 or float32')
    feat_np = np.float32 if all(np.asarray(f[i]).dtype == np.float32 for f in l_feats for i in range(3)) else np.float64
    # three ways in: row blocks that already sit back to back (what the batch analysis returns: viewed, no copy), one
    # C-contiguous block per utterance (what a caller that read one feature file per utterance holds: pointers handed to the
    # library, whose host threads copy them straight into page-locked staging), anything else: stacked here
    mags, reals, imags = ([f[i] for f in l_feats] for i in range(3))
    mag = _stack_rows(mags, feat_np, copy=False)
    real = _stack_rows(reals, feat_np, copy=False) if mag is not None else None
    imag = _stack_rows(imags, feat_np, copy=False) if real is not None else None
    blocks = None
    if imag is None:
        ok = lambda a: isinstance(a, np.ndarray) and a.dtype == np.dtype(feat_np) and a.ndim == 2 and a.flags['C_CONTIGUOUS']
        if (all(ok(a) for a in mags) and all(ok(a) for a in reals) and all(ok(a) for a in imags) and
                all(m.shape[0] == r.shape[0] == i.shape[0] for m, r, i in zip(mags, reals, imags))):
            blocks = tuple((C.c_void_p * n_utt)(*[a.ctypes.data for a in l]) for l in (mags, reals, imags))
            block_rows = np.ascontiguousarray([a.shape[0] for a in mags], dtype=np.int64)
        else:
            mag, real, imag = (_stack_rows(l, feat_np) for l in (mags, reals, imags))
    noise = None
    if l_noise is not None:
        noise = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float64) for v in l_noise]))
    hpf_sos = output_hpf_sos(fs) if b_out_hpf else None
    out = _lib.pinned.empty(int(out_off[-1]), dtype=out_dtype)
    code = lambda dt: _lib.MPB_F32 if np.dtype(dt) == np.dtype(np.float32) else _lib.MPB_F64
    ppt = {'magphase': 0, 'linear': 1, 'min_phase': 2}[per_phase_type]
    with _lib.device_gate():
        if blocks is not None:
            _lib.check(_lib.lib().mpb_synthesis_compressed_hostv2(
                plan.handle, blocks[0], blocks[1], blocks[2], _lib.ptr(block_rows), n_utt, code(feat_np), _lib.ptr(need),
                _lib.ptr(noise), int(sum(l_ns_len)), _lib.ptr(mt_key), C.byref(mt_pos) if mt_pos is not None else None,
                C.byref(fr), ppt, _lib.ptr(hpf_sos), _lib.ptr(out), code(out_dtype), out.size))
        else:
            _lib.check(_lib.lib().mpb_synthesis_compressed_host2(
                plan.handle, _lib.ptr(mag), _lib.ptr(real), _lib.ptr(imag), code(feat_np), mag.shape[0], _lib.ptr(need),
                _lib.ptr(noise), int(sum(l_ns_len)), _lib.ptr(mt_key), C.byref(mt_pos) if mt_pos is not None else None,
                C.byref(fr), ppt, _lib.ptr(hpf_sos), _lib.ptr(out), code(out_dtype), out.size))
    if mt_key is not None:
        (rng if rng is not None else np.random).set_state((np_state[0], mt_key, int(mt_pos.value), np_state[3], np_state[4]))
    l_out = [out[out_off[u]:out_off[u + 1]] for u in range(n_utt)]
    return l_out


# ----------------------------------------------------------------------------------------------
# post-filter and minimum phase
# ----------------------------------------------------------------------------------------------
def post_filter(m_mag_mel_log, fs, av_len_at_zero=None, av_len_at_nyq=None, boost_at_zero=None, boost_at_nyq=None):
    """MagPhase post-filter on the mel log-magnitude.  src/magphase.py:2300-2378"""
    m = np.ascontiguousarray(m_mag_mel_log, dtype=np.float64)
    nfrms, mag_dim = m.shape
    if mag_dim != 60:
        warnings.warn('Post-filter: It has been only tested with 60 dimensional mag data. If you use another dimension, '
                      'the result may be suboptimal.')
    opts = [av_len_at_zero, av_len_at_nyq, boost_at_zero, boost_at_nyq]
    if fs == 48000:
        defaults = [round_to_int(11.0 * (mag_dim / 60.0)), round_to_int(3.0 * (mag_dim / 60.0)), 1.8, 2.0]
    elif fs == 16000:
        if any(o is None for o in opts):
            warnings.warn('Post-filter: The default parameters for 16kHz sample rate have not being tunned.')
        defaults = [round_to_int(9.0 * (mag_dim / 60.0)), round_to_int(12.0 * (mag_dim / 60.0)), 2.0, 1.6]
    else:
        if any(o is None for o in opts):
            raise ValueError('Post-filter: It has only been tested with 16kHz and 48kHz sample rates.'
                             '\nProvide your own values for the options: av_len_at_zero, av_len_at_nyq, boost_at_zero,'
                             '\nboost_at_nyq if you use another sample rate')
        defaults = opts
    l0, l1, b0, b1 = [d if o is None else o for d, o in zip(defaults, opts)]
    # integer bookkeeping of the averaging windows (:2343-2346) and the boundary fill (:2357-2358)
    v_nx = np.arange(np.floor(l0 / 2), mag_dim - np.floor(l1 / 2)).astype(int)
    v_lens = (2 * np.ceil(np.linspace(l0, l1, v_nx.size) / 2) - 1).astype(int)
    centre = np.empty(mag_dim, dtype=np.int32)
    half = np.empty(mag_dim, dtype=np.int32)
    centre[v_nx] = v_nx
    half[v_nx] = v_lens // 2
    centre[:v_nx[0]], half[:v_nx[0]] = v_nx[0], half[v_nx[0]]
    centre[v_nx[-1]:], half[v_nx[-1]:] = v_nx[-1], half[v_nx[-1]]
    tilt = np.ascontiguousarray(np.linspace(b0, b1, mag_dim), dtype=np.float64)
    out = np.empty_like(m)
    _lib.check(_lib.lib().mpb_post_filter_host(_lib.ctx(), _lib.ptr(m), nfrms, mag_dim, _lib.ptr(centre), _lib.ptr(half),
                                               _lib.ptr(tilt), _lib.ptr(out)))
    return out


_MERLIN_TABLES = {}


def _freqt_matrix(n_out, n_in, alpha):
    """SPTK freqt (Oppenheim all-pass recursion, inputs consumed from the last coefficient down) as an (n_out x n_in) matrix:
    the recursion run on all unit vectors at once."""
    g = np.zeros((n_out, n_in))
    b = 1.0 - alpha * alpha
    eye = np.eye(n_in)
    for i in range(n_in - 1, -1, -1):
        d = g.copy()
        g[0] = eye[i] + alpha * d[0]
        if n_out > 1:
            g[1] = b * d[0] + alpha * d[1]
        for j in range(2, n_out):
            g[j] = d[j - 1] + alpha * (d[j] - g[j - 1])
    return g


def post_filter_merlin(m_mag_mel_log, fs, pf_coef=1.4):
    """Merlin-style post-filter (src/magphase.py:3375-3465): lifter the mel cepstrum by pf_coef above the second
    coefficient and restore the frame energy.  The reference pipes the cepstra through nine SPTK-3.9 binaries (x2x, freqt,
    c2acr, vopr, mc2b, bcp, sopr, merge, b2mc; command lines :3418-3444) that are not available here: every stage is
    restated from the published algorithm of its binary, with SPTK's float32 files at every boundary -- PARITY UNPINNED,
    like `mcep`.  The heavy stage (two energy integrals over a 4096-point spectrum per frame, `freqt | c2acr`) runs on the
    device (k_cep_energy); the O(60)-per-frame recursions around it are vectorised NumPy."""
    m = np.ascontiguousarray(m_mag_mel_log, dtype=np.float64)
    n = m.shape[1]
    fft_len, alpha = 4096, define_alpha(fs)
    f32 = lambda x: np.asarray(x, dtype=np.float32).astype(np.float64)
    key = (n, alpha)
    if key not in _MERLIN_TABLES:
        # freqt -m n-1 -a alpha -M fft_len/2-1 -A 0 (all-pass back to the linear axis: a = -alpha) folded into the cosine
        # table of c2acr's length-4096 real transform: G[j][k] = sum_i F[i][j] cos(2 pi k i / L)
        F = _freqt_matrix(fft_len // 2, n, -alpha)
        k = np.arange(fft_len // 2 + 1)
        G = np.cos(2 * np.pi * np.outer(np.arange(fft_len // 2), k) / fft_len).T @ F
        _MERLIN_TABLES[key] = np.ascontiguousarray(G.T)                      # [n][K]
    G = _MERLIN_TABLES[key]
    ext = np.hstack((m, m[:, -2:0:-1]))                                      # la.rceps(in_type='log', out_type='compact')
    ceps = np.fft.ifft(ext, axis=1).real
    ceps[:, 1:(n - 2)] *= 2
    mcep = f32(ceps[:, :n])
    w = f32(np.r_[1.0, 1.0, np.full(n - 2, float('%1.2f' % pf_coef))])
    lifted = f32(mcep * w)
    both = np.ascontiguousarray(np.vstack((mcep, lifted)))
    r = np.empty(both.shape[0])
    _lib.check(_lib.lib().mpb_cep_energy_host(_lib.ctx(), _lib.ptr(both), both.shape[0], n, _lib.ptr(G), G.shape[1], fft_len,
                                              _lib.ptr(r)))
    r0, p_r0 = f32(r[:m.shape[0]]), f32(r[m.shape[0]:])
    b = lifted.copy()                                                        # mc2b
    for i in range(n - 2, -1, -1):
        b[:, i] = b[:, i] - alpha * b[:, i + 1]
    b = f32(b)
    p_b0 = f32(f32(f32(np.log(f32(r0 / p_r0))) / 2.0) + b[:, 0])
    merged = np.hstack((p_b0[:, None], b[:, 1:]))
    mcep_pf = merged.copy()                                                  # b2mc
    mcep_pf[:, :-1] = merged[:, :-1] + alpha * merged[:, 1:]
    mcep_pf = f32(mcep_pf)
    out = mcep_pf @ np.cos(np.arange(n)[:, None] * warped_axis(0.0, n)[None, :])   # la.mcep_to_sp_cosmat(alpha=0, 'log')
    out[np.isnan(out)] = MAGIC
    return out


def build_min_phase_from_mag_spec(m_mag):
    """Minimum-phase complex spectrum of a magnitude spectrum.  la.build_min_phase_from_mag_spec, src/libaudio.py:920-934"""
    m = np.ascontiguousarray(m_mag, dtype=np.float64)
    fft_len = 2 * (m.shape[1] - 1)
    out = np.empty(m.shape, dtype=np.complex128)
    _lib.check(_lib.lib().mpb_min_phase_host(_lib.ctx(), _lib.ptr(m), m.shape[0], fft_len, _lib.ptr(out)))
    return out


def synthesis_from_acoustic_modelling(in_feats_dir, filename_token, out_syn_dir, mag_dim, phase_dim, fs, fft_len=None,
                                      pf_type='no', b_const_rate=False):
    """Feature files -> wav.  src/magphase.py:3229-3275."""
    print("\nSynthesising file: " + filename_token + '.wav............................')
    m_mag_mel_log = io.read_binfile(in_feats_dir + '/' + filename_token + '.mag', dim=mag_dim)
    m_real_mel = io.read_binfile(in_feats_dir + '/' + filename_token + '.real', dim=phase_dim)
    m_imag_mel = io.read_binfile(in_feats_dir + '/' + filename_token + '.imag', dim=phase_dim)
    v_lf0 = io.read_binfile(in_feats_dir + '/' + filename_token + '.lf0', dim=1)
    if pf_type == 'magphase':
        print('Using MagPhase postfilter...')
        m_mag_mel_log = post_filter(m_mag_mel_log, fs)
    elif pf_type == 'merlin':
        print('Using Merlin postfilter...')
        m_mag_mel_log = post_filter_merlin(m_mag_mel_log, fs)
    elif pf_type == 'no':
        print('No postfilter...')
    v_syn_sig = synthesis_from_compressed(m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0, fs, fft_len=fft_len,
                                          b_const_rate=b_const_rate)
    io.write_audio_file(out_syn_dir + '/' + filename_token + '.wav', v_syn_sig, fs)
    return


def griffin_lim(m_mag, v_shift, win_func=np.hanning, phase_init='random', niters=30):
    """Pitch synchronous Griffin-Lim, src/magphase.py:3318-3373: phase_init 'random' (np.random.rand on the global
    stream), 'linear', 'min_phase' or a phase matrix.  Every iteration runs on the device (device.griffin_lim)."""
    from . import device
    print('Starting Griffin-Lim. It could take a while...')
    return device.griffin_lim(m_mag, v_shift, win_func=win_func, phase_init=phase_init, niters=niters)


def numpy_stream_uniform(low, high, n):
    """np.random.uniform(low, high, n) on NumPy's global legacy stream, generated on the device (bit-identical,
    the global state advances as if NumPy had drawn the numbers).  Exposed for tests."""
    st = np.random.get_state()
    if st[0] != 'MT19937':
        raise RuntimeError('NumPy global stream is not MT19937')
    key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
    pos = C.c_int32(int(st[2]))
    out = np.empty(int(n), dtype=np.float64)
    _lib.check(_lib.lib().mpb_mt19937_uniform_host(_lib.ctx(), _lib.ptr(key), C.byref(pos), int(n), float(low), float(high),
                                                   _lib.ptr(out)))
    np.random.set_state((st[0], key, int(pos.value), st[3], st[4]))
    return out


# ----------------------------------------------------------------------------------------------
# legacy v1 analysis (kept for its signature: named in BASELINE.json north_star)
# ----------------------------------------------------------------------------------------------
def next_pow_of_two(x):
    """src/libaudio.py next_pow_of_two"""
    if x < 2:
        x = 2
    return int(2 ** np.ceil(np.log2(x)).astype(int))


def _raw_mcep_plan(fft_len, n_coeffs, alpha):
    """A mel plan whose three streams all produce plain la.sp_to_mcep output (n_coeffs each)."""
    key = ('raw', _lib.default_device(), _lib.current_slot(), fft_len, n_coeffs, float("%1.2f" % alpha))
    if key not in _MelPlan._cache:
        eye = np.ascontiguousarray(np.eye(n_coeffs))
        h = C.c_void_p()
        _lib.check(_lib.lib().mpb_mel_create(_lib.ctx(), fft_len, float("%1.2f" % alpha), n_coeffs, float("%1.2f" % alpha),
                                             n_coeffs, n_coeffs, _lib.ptr(eye), _lib.ptr(eye), C.byref(h)))
        _MelPlan._cache[key] = h
    return _MelPlan._cache[key]


def sp_to_mcep(m_sp, n_coeffs=60, alpha=0.77, in_type=3, fft_len=0):
    """la.sp_to_mcep (src/libaudio.py:575-601): SPTK `mcep -a alpha -m n-1 -l fft_len -e 1.0E-8 -j 0 -f 0.0 -q in_type`
    on the device (see mpb_mel.cu for the restatement).  in_type: 3 |f(w)|, 2 ln|f(w)|, 1 20*log10|f(w)|."""
    m = np.ascontiguousarray(m_sp, dtype=np.float64)
    if fft_len == 0:
        fft_len = 2 * (m.shape[1] - 1)
    if in_type == 1:
        m = np.ascontiguousarray(m.astype(np.float32).astype(np.float64) * (np.log(10.0) / 20.0))   # dB -> ln
    elif in_type not in (2, 3):
        raise ValueError('in_type must be 1, 2 or 3')
    h = _raw_mcep_plan(fft_len, n_coeffs, alpha)
    outs = [np.empty((m.shape[0], n_coeffs)) for _ in range(3)]
    _lib.check(_lib.lib().mpb_sp_to_mcep_host(h, _lib.ptr(m), _lib.ptr(m), _lib.ptr(m), m.shape[0], _lib.ptr(outs[0]),
                                              _lib.ptr(outs[1]), _lib.ptr(outs[2])))
    return outs[0] if in_type == 3 else outs[1]


def sp_mel_warp(m_sp, nbins_out, alpha=0.77, in_type=3):
    """la.sp_mel_warp (src/libaudio.py:643-661) as a function of its own: sp_to_mcep on the device, then the nbins_out x
    nbins_out cosine matrix of la.mcep_to_sp_cosmat(alpha=0.0) (:605-631) on the host (a few thousand multiplies per frame;
    format_for_modelling runs the same product inside its kernels).  Output type follows in_type: 3 -> |.|, 2 -> ln, 1 -> dB."""
    m_mcep = sp_to_mcep(m_sp, n_coeffs=nbins_out, alpha=alpha, in_type=in_type)
    m_sp_wrp = np.dot(m_mcep, np.cos(np.arange(nbins_out)[:, None] * warped_axis(0.0, nbins_out)[None, :]))
    if in_type == 3:
        return np.exp(m_sp_wrp)
    if in_type == 1:
        return m_sp_wrp * (20 / np.log(10))
    return m_sp_wrp


def analysis_with_del_comp_and_ph_encoding(v_in_sig, nFFT, fs, mvf, pm=None):
    """Legacy v1 analysis (src/magphase.py:573-598 over analysis_with_del_comp :338-368): spectral envelope and the
    sine / cosine of the phase up to `mvf` Hz, each as 60 mel-cepstral coefficients.
    Returns (m_spmgc, m_phs_mgc, m_phc_mgc, v_shift).  `pm` (seconds) replaces the REAPER call when given."""
    v_in_sig = np.asarray(v_in_sig, dtype=np.float64)
    if pm is None:
        tmp_wav, tmp_est = io.ins_pid('temp.wav'), io.ins_pid('temp.pm')
        reaper = io.find_tool('reaper')
        if reaper is None:
            raise RuntimeError('REAPER binary not found: pass pm=<pitch marks in seconds>')
        io.write_audio_file(tmp_wav, v_in_sig, fs, norm=None)
        call(reaper + " -s -x 400 -m 50 -a -u 0.005 -i %s -p %s" % (tmp_wav, tmp_est), shell=True)
        pm = np.loadtxt(tmp_est, skiprows=7)[:, 0]
        os.remove(tmp_wav); os.remove(tmp_est)
        pm = pm[np.hstack((True, np.diff(pm) > 0))]                       # src/libaudio.py:479-485
        if (pm[-1] * fs) >= (np.size(v_in_sig) - 1):
            pm = pm[:-1]
    v_pm_smpls = np.asarray(pm, dtype=np.float64) * fs
    P, v_shift, v_rights = frame_geometry(v_pm_smpls, v_in_sig.size)
    len_max = int(np.max(v_shift + v_rights + 1))
    if nFFT < len_max:
        raise ValueError("nFFT (%d) is shorter than the maximum frame length (%d)" % (nFFT, len_max))
    (m_sp, m_phc, m_phs), _, _ = _frames_call([v_in_sig], [v_pm_smpls], nFFT, [np.hanning], 'feats')
    m_phc = np.where(m_sp == 0.0, 1.0, m_phc)                             # np.angle(0) = 0 -> cos = 1 (:423-426)
    m_spmgc = sp_to_mcep(m_sp)                                            # 60 coefficients, alpha 0.77, in_type 3
    mvf_bin = int(round_to_int(mvf * nFFT / float(fs)))
    n_ph = next_pow_of_two(mvf_bin) + 1
    from scipy import interpolate                                         # cubic resampling of a handful of bins: host
    grid = np.linspace(0, mvf_bin - 1, n_ph)
    m_phs_i = interpolate.interp1d(np.arange(mvf_bin), m_phs[:, :mvf_bin], kind='cubic')(grid)
    m_phc_i = interpolate.interp1d(np.arange(mvf_bin), m_phc[:, :mvf_bin], kind='cubic')(grid)
    m_phs_mgc = sp_to_mcep(m_phs_i, in_type=1)
    m_phc_mgc = sp_to_mcep(m_phc_i, in_type=1)
    return m_spmgc, m_phs_mgc, m_phc_mgc, v_shift.astype(int)

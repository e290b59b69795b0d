"""CPU oracle for the MagPhase analysis/synthesis hot path.  TEST INFRASTRUCTURE ONLY.

A clean float64 NumPy restatement of the reference algorithm (CSTR-Edinburgh/magphase,
``src/magphase.py`` + ``src/libaudio.py`` + ``src/libutils.py``).  Each function cites the
reference ``file:line`` it follows.  It exists to *check* the CUDA path:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import it;
  * the product package ``magphase_b200`` never imports anything under ``oracle/``.

Parity status
-------------
Pinned against the real reference (the mechanically py2->py3 translated copy that
``oracle/make_ref.py`` writes to ``oracle/_ref``) by ``tests/test_oracle_vs_ref.py`` and
against the committed golden vectors in ``tests/golden/`` (generated from ``oracle/_ref``
by ``tests/golden/make_golden.py``) for every function below EXCEPT the SPTK part:

  * ``mcep_j0`` (SPTK-3.9 ``mcep -j 0``; call site ``src/libaudio.py:575-601``, command line
    ``:589``) restates a third-party binary that is neither in ``/root/reference`` nor in this
    image (``tools/download_and_compile_tools.sh:5`` fetches SPTK-3.9 from sourceforge).
    The reference ships no analysis-output fixture either => **parity unpinned** for
    ``sp_mel_warp`` / ``format_for_modelling`` (everything downstream of ``mcep``).
"""
import warnings

import numpy as np
from scipy import interpolate, signal

MAGIC = -1.0e10  # src/libaudio.py:17


# ----------------------------------------------------------------------------------------
# constants per sample rate                                    src/magphase.py:3279-3317
# ----------------------------------------------------------------------------------------
def define_alpha(fs):
    table = {16000: 0.58, 22050: 0.65, 44100: 0.76, 48000: 0.77}
    if fs not in table:
        raise ValueError("Sample rate %d not supported yet." % (fs))
    return table[fs]


def define_fft_len(fs):
    if fs in (22050, 16000):
        return 2048
    if fs == 8000:
        return 1024
    return 4096


def define_crossfade_params(fs):
    bw = 2000
    if fs == 48000:
        return 5000, bw
    if fs == 16000:
        return 2500, bw
    warnings.warn('Constant crsf_cf not tested nor tunned to synthesise at fs=%d Hz.' % fs)
    if fs == 44100:
        return 4500, bw
    return 3500, bw


# ----------------------------------------------------------------------------------------
# bookkeeping
# ----------------------------------------------------------------------------------------
def round_to_int(x):
    """np.round is half-to-even.  src/libutils.py:131-133"""
    return np.round(x).astype(int)


def log_protected(x):
    """src/libaudio.py:241-248"""
    with np.errstate(divide='ignore', invalid='ignore'):
        y = np.log(x)
    y = np.array(y, dtype=np.float64, copy=True)
    y[np.isinf(y)] = MAGIC
    y[np.isnan(y)] = MAGIC
    return y


def f0_to_lf0(v_f0):
    """src/libaudio.py:458-465"""
    with np.errstate(divide='ignore'):
        v = np.log(v_f0)
    v[np.isinf(v)] = MAGIC
    return v


def shift_to_f0(v_shift, v_voi, fs, b_smooth=False):
    """src/magphase.py:2198-2207"""
    v_f0 = v_voi * fs / v_shift.astype('float64')
    if b_smooth:
        v_f0 = v_voi * signal.medfilt(v_f0)
    return v_f0


def f0_to_shift(v_f0, fs, unv_frm_rate_ms=5):
    """Unvoiced frames get a fixed 5 ms shift.  src/magphase.py:2210-2215"""
    v = np.array(v_f0, dtype=np.float64, copy=True)
    v[v == 0] = 1000.0 / unv_frm_rate_ms
    return fs / v


def frame_limits(v_pm_smpls, n_smpls):
    """Extended integer marks P=[0, round(pm)..., n-1], left lengths (=v_shift) and right lengths.

    src/magphase.py:74-84 and :112-117
    """
    P = np.hstack((0, round_to_int(np.asarray(v_pm_smpls)), n_smpls - 1)).astype(np.int64)
    v_shift = P[1:-1] - P[:-2]
    v_rights = P[2:] - P[1:-1]
    return P, v_shift, v_rights


# ----------------------------------------------------------------------------------------
# windows
# ----------------------------------------------------------------------------------------
def side_window(length, kind):
    """Samples 0..length of the rising half of a (2*length+1)-point window.

    kind 'hann'        -> np.hanning              (src/libaudio.py:70-84)
    kind 'bartlett2.5' -> np.bartlett(.)**2.5     (src/magphase.py:67-69)
    kind callable      -> kind(1 + 2*length)      (any win_func, src/libaudio.py:72-78)
    """
    if callable(kind):
        w = np.asarray(kind(1 + 2 * length), dtype=np.float64)
    elif kind == 'hann':
        w = np.hanning(1 + 2 * length)
    elif kind == 'bartlett2.5':
        w = np.bartlett(1 + 2 * length) ** 2.5
    else:
        raise ValueError(kind)
    return w[:length + 1]


def asym_window(left_len, right_len, kind='hann'):
    """Asymmetric window whose peak sits at index left_len.  src/libaudio.py:70-84"""
    wl = side_window(int(left_len), kind)
    wr = side_window(int(right_len), kind)[::-1]
    return np.concatenate((wl, wr[1:]))


def centred_window(len_l, len_r, totlen):
    """raised_hanning(att=1) asymmetric window centred at totlen//2, zero elsewhere.

    src/libaudio.py:90-103 with win_func=raised_hanning (src/magphase.py:25-31) and
    b_fill_w_bound_val=True (the edge value of a Hann window is 0, so the fill is 0).
    """
    w = np.zeros(totlen)
    short = asym_window(len_l, len_r, 'hann')
    w += short[0]
    z = totlen // 2 - len_l
    w[z:z + short.size] = short
    return w


# ----------------------------------------------------------------------------------------
# analysis                                                      src/magphase.py:266-334
# ----------------------------------------------------------------------------------------
def analysis_frames(v_sig, v_pm_smpls, fft_len, kinds=None):
    """Windowed, zero-padded, un-delayed frames (pitch mark at index 0 of each row).

    Follows windowing() src/magphase.py:74-119 and the pad / truncate / rotate loop of
    analysis_with_del_comp_from_pm src/magphase.py:305-323.
    kinds: optional per-frame window kind list (used by the noise branch of synthesis).
    Returns (m_frms[n, fft_len], v_shift[n] int64, P[n+2] int64).
    """
    v_sig = np.asarray(v_sig, dtype=np.float64)
    P, v_shift, v_rights = frame_limits(v_pm_smpls, v_sig.size)
    n = v_shift.size
    m = np.zeros((n, fft_len))
    for f in range(n):
        l, r = int(v_shift[f]), int(v_rights[f])
        frm = v_sig[P[f]:P[f + 2] + 1] * asym_window(l, r, 'hann' if kinds is None else kinds[f])
        if frm.size <= fft_len:
            row = np.zeros(fft_len)
            row[:frm.size] = frm
        else:
            warnings.warn("fft_len (%d) is shorter than the current detected frame length (%d). " % (fft_len, frm.size))
            row = frm[:fft_len]
        m[f] = np.concatenate((row[l:], row[:l]))
    return m, v_shift, P


def analysis_fft_from_pm(v_sig, fs, v_pm_smpls, fft_len=None, win_func=None):
    """Half spectrum of every pitch-synchronous frame.  src/magphase.py:266-334
    win_func: None (np.hanning), a callable or a per-frame list of callables / kind names (src/magphase.py:102-108)."""
    if fft_len is None:
        fft_len = define_fft_len(fs)
    kinds = None
    if win_func is not None:
        kinds = list(win_func) if isinstance(win_func, (list, tuple)) else [win_func] * np.size(v_pm_smpls)
    m_frms, v_shift, _ = analysis_frames(v_sig, v_pm_smpls, fft_len, kinds)
    m_fft = np.fft.fft(m_frms)[:, :fft_len // 2 + 1].copy()
    return m_fft, v_shift


def compute_lossless_feats(m_fft, v_shift, v_voi, fs):
    """mag, real/|X|, imag/|X| (0 where |X|==0), f0.  src/magphase.py:457-476"""
    m_mag = np.absolute(m_fft)
    zero = m_mag == 0.0
    div = np.where(zero, 1.0, m_mag)
    m_real = np.where(zero, 0.0, m_fft.real / div)
    m_imag = np.where(zero, 0.0, m_fft.imag / div)
    v_f0 = shift_to_f0(v_shift, v_voi, fs, b_smooth=False)
    return m_mag, m_real, m_imag, v_f0


def analysis_lossless_from_pm(v_sig, fs, v_pm_smpls, v_voi, fft_len=None):
    """analysis_lossless (src/magphase.py:2869-2906) with REAPER's output given as input."""
    m_fft, v_shift = analysis_fft_from_pm(v_sig, fs, v_pm_smpls, fft_len)
    m_mag, m_real, m_imag, v_f0 = compute_lossless_feats(m_fft, v_shift, np.asarray(v_voi, dtype=np.float64), fs)
    return m_mag, m_real, m_imag, v_f0, fs, v_shift


# ----------------------------------------------------------------------------------------
# lossless synthesis                                           src/magphase.py:1759-1776
# ----------------------------------------------------------------------------------------
def half_to_frames(m_half):
    """Hermitian-extend (imag of DC and Nyquist dropped, src/libaudio.py:369-388), inverse FFT,
    fftshift so that time zero sits at column N/2.  src/magphase.py:1768-1770"""
    N = 2 * (m_half.shape[1] - 1)
    x = np.fft.irfft(m_half, n=N, axis=1)  # irfft ignores Im(DC), Im(Nyquist): same as zeroing them
    return np.fft.fftshift(x, axes=1)


def ola(m_frm, v_pm):
    """Pitch-synchronous overlap-add.  src/magphase.py:34-62"""
    v_pm = np.asarray(v_pm).astype(int)
    n, N = m_frm.shape
    buf = np.zeros(v_pm[-1] + N)
    v_shift = np.diff(np.hstack((0, v_pm)))
    start = v_pm - v_pm[0]
    for i in range(n):
        buf[start[i]:start[i] + N] += m_frm[i]
    buf = buf[(N // 2 - v_pm[0]):]
    return buf[:(v_pm[-1] + v_shift[-1] + 1)]


def synthesis_from_lossless(m_mag, m_real, m_imag, v_f0, fs):
    """src/magphase.py:1759-1776"""
    u = m_real + 1j * m_imag
    a = np.absolute(u)
    a[a == 0.0] = 1.0
    m_frm = half_to_frames(m_mag * u / a)
    v_pm = np.cumsum(f0_to_shift(v_f0, fs))
    return ola(m_frm, v_pm)


# ----------------------------------------------------------------------------------------
# mel warping                                  src/libaudio.py:575-684, magphase.py:2479-2487
# ----------------------------------------------------------------------------------------
def warped_axis(alpha, nbins):
    """First-order all-pass warped frequency axis on [0, pi].  src/libaudio.py:611-613, 711-718"""
    w = np.linspace(0, np.pi, num=nbins)
    with np.errstate(divide='ignore', invalid='ignore'):
        wt = np.arctan((1 - alpha ** 2) * np.sin(w) / ((1 + alpha ** 2) * np.cos(w) - 2 * alpha))
    wt[wt < 0] += np.pi
    return wt


def build_mel_curve(alpha, nbins, amp=np.pi):
    """src/libaudio.py:711-718"""
    return warped_axis(alpha, nbins) * (amp / np.pi)


def cosine_matrix(n_ceps, nbins, alpha):
    """T[j,k] = cos(j * warped_w[k]).  src/libaudio.py:605-631"""
    return np.cos(np.arange(n_ceps)[:, None] * warped_axis(alpha, nbins)[None, :])


def mcep_to_sp_cosmat(m_mcep, n_spbins, alpha=0.77, out_type='abs'):
    """src/libaudio.py:605-631"""
    m_sp = np.dot(m_mcep, cosine_matrix(m_mcep.shape[1], n_spbins, alpha))
    if out_type == 'abs':
        m_sp = np.exp(m_sp)
    elif out_type == 'db':
        m_sp = m_sp * (20 / np.log(10))
    return m_sp


def n_full_mel_coeffs(freq_hz, phase_dim, alpha, fs):
    """get_num_full_mel_coeffs_from_num_phase_coeffs.  src/magphase.py:2479-2487"""
    w = 2 * np.pi * freq_hz / float(fs)
    m = np.arctan((1 - alpha ** 2) * np.sin(w) / ((1 + alpha ** 2) * np.cos(w) - 2 * alpha))
    if m < 0:
        m += np.pi
    return int(round_to_int(1 + (np.pi * (phase_dim - 1) / float(m))))


def sp_mel_unwarp(m_sp_mel, nbins_out, alpha=0.77, in_type='log'):
    """Low-dim mel log spectrum -> nbins_out linear-frequency bins.  src/libaudio.py:667-684

    NB the doubling stops one short: cepstral index ncoeffs-2 is NOT doubled (:679)."""
    nc = m_sp_mel.shape[1]
    if in_type == 'abs':
        m_sp_mel = np.log(m_sp_mel)
    ext = np.hstack((m_sp_mel, m_sp_mel[:, -2:0:-1]))
    ceps = np.fft.ifft(ext, axis=1).real
    ceps[:, 1:(nc - 2)] *= 2
    return mcep_to_sp_cosmat(ceps[:, :nc], nbins_out, alpha=alpha, out_type=in_type)


def freqt_matrix(n_out, n_in, alpha):
    """SPTK ``freqt`` (Oppenheim all-pass recursion) as an (n_out x n_in) matrix.

    Third-party: SPTK-3.9 ``freqt()``; not in /root/reference.  Built by running the published
    recursion on every unit vector at once (columns = inputs)."""
    b = 1.0 - alpha * alpha
    g = np.zeros((n_out, n_in))          # state after consuming inputs, one column per unit vector
    eye = np.eye(n_in)
    for i in range(n_in - 1, -1, -1):    # i = -m1..0 in SPTK: consumes c[m1], ..., c[0]
        d = g.copy()
        g[0] = eye[i] + alpha * d[0]
        if n_out > 1:
            g[1] = b * d[0] + alpha * d[1]
        for j in range(2, n_out):
            g[j] = d[j - 1] + alpha * (d[j] - g[j - 1])
    return g


_FREQT_CACHE = {}


def _freqt_cached(n_out, n_in, alpha):
    key = (n_out, n_in, float(alpha))
    if key not in _FREQT_CACHE:
        _FREQT_CACHE[key] = freqt_matrix(n_out, n_in, alpha)
    return _FREQT_CACHE[key]


def mcep_j0(m_sp, n_coeffs=60, alpha=0.77, in_type=3, fft_len=0):
    """Restatement of ``mcep -a alpha -m n-1 -l N -e 1.0E-8 -j 0 -f 0.0 -q in_type`` as invoked
    at src/libaudio.py:589 (third-party SPTK-3.9, float32 file I/O at src/libaudio.py:582,593).

    With ``-j 0`` the Newton loop never runs; the output is SPTK's initial estimate:
    float32 input -> periodogram (+eps) -> log -> IFFT -> halve c[0], c[N/2] -> freqt -> float32.
    PARITY UNPINNED: no SPTK binary/source/fixture is available here (see module docstring).
    """
    x = np.asarray(m_sp, dtype=np.float32).astype(np.float64)   # lu.write_binfile -> float32
    H = x.shape[1]
    if fft_len == 0:
        fft_len = 2 * (H - 1)
    alpha = float("%1.2f" % alpha)                              # printed with %1.2f on the command line
    if in_type == 3:
        p = x * x
    elif in_type == 2:
        p = np.exp(2.0 * x)
    elif in_type == 1:
        p = 10.0 ** (x / 10.0)
    else:
        raise ValueError(in_type)
    logp = np.log(p + 1.0e-8)
    c = np.fft.irfft(logp, n=fft_len, axis=1)[:, :H]
    c[:, 0] *= 0.5
    c[:, H - 1] *= 0.5
    mc = c @ _freqt_cached(n_coeffs, H, alpha).T
    return mc.astype(np.float32).astype(np.float64)             # float32 file read back as float64


def sp_mel_warp(m_sp, nbins_out, alpha=0.77, in_type=3):
    """src/libaudio.py:643-661"""
    m_mcep = mcep_j0(m_sp, n_coeffs=nbins_out, alpha=alpha, in_type=in_type)
    out_type = {3: 'abs', 1: 'db', 2: 'log'}[in_type]
    return mcep_to_sp_cosmat(m_mcep, nbins_out, alpha=0.0, out_type=out_type)


def format_for_modelling(m_mag, m_real, m_imag, v_f0, fs, mag_dim=60, phase_dim=45, alpha_phase=None):
    """src/magphase.py:2490-2544 (b_mag_fbank_mel=False branch)"""
    alpha = define_alpha(fs)
    v_voi = (v_f0 > 0).astype('float')
    v_lf0 = f0_to_lf0(v_voi * signal.medfilt(v_f0))
    m_mag_mel_log = log_protected(sp_mel_warp(m_mag, mag_dim, alpha=alpha, in_type=3))
    crsf_cf, _ = define_crossfade_params(fs)
    if alpha_phase is None:
        alpha_phase = alpha
    nmel = n_full_mel_coeffs(crsf_cf, phase_dim, alpha_phase, fs)
    m_real_mel = sp_mel_warp(m_real, nmel, alpha=alpha_phase, in_type=2)[:, :phase_dim]
    m_imag_mel = sp_mel_warp(m_imag, nmel, alpha=alpha_phase, in_type=2)[:, :phase_dim]
    m_real_mel = np.clip(m_real_mel * v_voi[:, None], -1, 1)
    m_imag_mel = np.clip(m_imag_mel * v_voi[:, None], -1, 1)
    return m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0


# ----------------------------------------------------------------------------------------
# constant <-> variable frame rate                      src/magphase.py:1426-1449, 2219-2252
# ----------------------------------------------------------------------------------------
def interp_from_variable_to_const_frm_rate(m_data, v_pm_smpls, const_rate_ms, fs):
    """src/magphase.py:2219-2239 (linear)"""
    m_data = np.asarray(m_data, dtype=np.float64)
    one_d = m_data.ndim == 1
    if one_d:
        m_data = m_data[:, None]
    step = fs * const_rate_ms / 1000
    centres = np.arange(step, v_pm_smpls[-1], step)
    if v_pm_smpls[0] > 0:
        f = interpolate.interp1d(np.r_[0, v_pm_smpls], np.vstack((m_data[0, :], m_data)), axis=0, kind='linear')
    else:
        f = interpolate.interp1d(v_pm_smpls, m_data, axis=0, kind='linear')
    out = f(centres)
    return out[:, 0] if one_d else out


def get_shifts_and_frm_locs_from_const_shifts(v_shift_c_rate, frm_rate_ms, fs):
    """Walk backwards from the last constant-rate centre, subtracting the interpolated shift, until
    the position leaves the interpolation range.  src/magphase.py:1426-1449"""
    n = np.size(v_shift_c_rate, 0)
    step = fs * frm_rate_ms / 1000
    centres = step * np.arange(1, n + 1)
    lo, hi = centres[0], centres[-1]
    shifts, locs = [], []
    pos = centres[-1]
    for _ in range(2 * n - 1):
        if pos < lo or pos > hi:
            break
        s = float(np.interp(pos, centres, v_shift_c_rate))
        locs.append(pos)
        shifts.append(s)
        pos = pos - s
    return np.array(shifts[::-1]), np.array(locs[::-1])


def interp_from_const_to_variable_rate(m_data, v_frm_locs_smpls, frm_rate_ms, fs):
    """src/magphase.py:2242-2252 (linear)"""
    n = np.size(m_data, 0)
    centres = (fs * frm_rate_ms / 1000) * np.arange(1, n + 1)
    return interpolate.interp1d(centres, m_data, axis=0, kind='linear')(v_frm_locs_smpls)


# ----------------------------------------------------------------------------------------
# compressed synthesis                                         src/magphase.py:825-997
# ----------------------------------------------------------------------------------------
def crossfade_curve(nbins, cut_off, bw, fs):
    """Left weight of la.spectral_crossfade: ones, falling half-Hann, zeros.  src/libaudio.py:160-186"""
    N = (nbins - 1) * 2
    bin_l = int(round_to_int((cut_off - bw / 2.0) * N / float(fs)))
    bin_r = int(round_to_int((cut_off + bw / 2.0) * N / float(fs)))
    B = bin_r - bin_l
    return np.hstack((np.ones(bin_l), np.hanning(2 * B + 1)[B:], np.zeros(nbins - bin_r - 1)))


def phase_uncompress(m_real_mel, m_imag_mel, alpha, fft_len, fs):
    """Pad phase_dim -> nmel by repeating the last column ('nearest' extrapolation), mel-unwarp.
    src/magphase.py:1219-1235"""
    nc = m_real_mel.shape[1]
    nmel = n_full_mel_coeffs(define_crossfade_params(fs)[0], nc, alpha, fs)
    idx = np.minimum(np.arange(nmel), nc - 1)
    H = 1 + fft_len // 2
    return (sp_mel_unwarp(m_real_mel[:, idx], H, alpha=alpha, in_type='log'),
            sp_mel_unwarp(m_imag_mel[:, idx], H, alpha=alpha, in_type='log'))


def noise_gain(m_ns_mag, rows):
    """sqrt(exp(mean(log|N|^2))) over the bins 1..H-2 of the selected rows.  src/magphase.py:902-903"""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return np.sqrt(np.exp(np.mean(log_protected(m_ns_mag[rows, 1:-1]) ** 2)))


def synthesis_from_compressed(m_mag_mel_log, m_real_mel, m_imag_mel, v_lf0, fs, fft_len=None,
                              b_voi_ap_win=True, b_const_rate=False, per_phase_type='magphase',
                              alpha_phase=None, b_out_hpf=True, v_noise=None, return_parts=False):
    """src/magphase.py:825-997.  ``v_noise``: the uniform(-1,1) noise to use; when None it is drawn
    from the global legacy numpy stream exactly where the reference draws it (:883)."""
    crsf_cf, crsf_bw = define_crossfade_params(fs)
    alpha = define_alpha(fs)
    if fft_len is None:
        fft_len = define_fft_len(fs)
    H = fft_len // 2 + 1

    v_f0 = np.exp(v_lf0)
    v_voi = v_f0 > 1.0
    v_shift = f0_to_shift(v_f0, fs)

    m_mag = np.exp(sp_mel_unwarp(m_mag_mel_log, H, alpha=alpha, in_type='log'))
    if alpha_phase is None:
        alpha_phase = alpha
    m_real, m_imag = phase_uncompress(m_real_mel, m_imag_mel, alpha_phase, fft_len, fs)

    if b_const_rate:
        v_shift, v_locs = get_shifts_and_frm_locs_from_const_shifts(v_shift, 5.0, fs)
        m_mag = interp_from_const_to_variable_rate(m_mag, v_locs, 5.0, fs)
        m_real = interp_from_const_to_variable_rate(m_real, v_locs, 5.0, fs)
        m_imag = interp_from_const_to_variable_rate(m_imag, v_locs, 5.0, fs)
        v_voi = interp_from_const_to_variable_rate(v_voi, v_locs, 5.0, fs) > 0.5
    n = v_shift.size

    curve = crossfade_curve(H, crsf_cf, crsf_bw, fs)
    m_mask = np.zeros((n, H))
    m_mask[v_voi, :] = curve[None, :]

    v_shift = v_shift.astype(int)                      # truncation BEFORE the cumsum (:879-880)
    v_pm = np.cumsum(v_shift)
    ns_len = v_pm[-1] + (v_pm[-1] - v_pm[-2])
    if v_noise is None:
        v_noise = np.random.uniform(-1, 1, ns_len)
    v_noise = np.asarray(v_noise, dtype=np.float64)
    if v_noise.size != ns_len:
        raise ValueError('noise length %d != %d' % (v_noise.size, ns_len))

    kinds = ['bartlett2.5' if (b_voi_ap_win and v_voi[i]) else 'hann' for i in range(n)]
    # windowing + frm_list_to_matrix + fftshift (src/magphase.py:893-896) == analysis_frames layout
    m_frm_ns, _, _ = analysis_frames(v_noise, v_pm, fft_len, kinds=kinds)
    m_ns = np.fft.fft(m_frm_ns)[:, :H].copy()
    m_ns_mag = np.absolute(m_ns)
    g_voi = noise_gain(m_ns_mag, v_voi)
    g_unv = noise_gain(m_ns_mag, ~v_voi)
    m_ns[v_voi, :] /= g_voi
    m_ns[~v_voi, :] /= g_unv

    m_ap = m_ns * m_mag
    m_ap[~v_voi, :] *= 10 ** ((build_mel_curve(alpha, H, amp=3.5) - 3.5) / 20)

    if per_phase_type == 'magphase':
        u = m_real + 1j * m_imag
        a = np.absolute(u)
        a[a == 0.0] = 1.0
        m_per = m_mag * (u / a)
    elif per_phase_type == 'linear':
        m_per = m_mag.astype(complex)
    elif per_phase_type == 'min_phase':
        m_per = build_min_phase_from_mag_spec(m_mag)
    else:
        raise ValueError(per_phase_type)
    m_per[v_voi, :] *= 10 ** (build_mel_curve(0.6, H, amp=2.0) / 20)

    m_per = m_per * (m_mask ** 0.5)
    m_ap = m_ap * ((1 - m_mask) ** 0.5)
    m_per[m_mask == 0.0] = 0
    m_ap[m_mask == 1.0] = 0
    m_syn = m_per + m_ap
    m_syn[:, 0] = np.absolute(m_syn[:, 0])
    m_syn[:, -1] = np.absolute(m_syn[:, -1])

    m_frm = half_to_frames(m_syn)
    se = np.r_[v_shift[0], v_shift, v_shift[-1], v_shift[-1]]
    for i in range(n):
        m_frm[i] *= centred_window(se[i] + se[i + 1], se[i + 2] + se[i + 3], fft_len)
    v_sig = ola(m_frm, v_pm)

    if b_out_hpf:
        v_b, v_a = signal.butter(4, 40 / (fs / 2.0), btype='highpass')
        v_sig = signal.lfilter(v_b, v_a, v_sig)
    if return_parts:
        return v_sig, dict(m_mag=m_mag, m_real=m_real, m_imag=m_imag, v_shift=v_shift, v_pm=v_pm,
                           v_voi=v_voi, g_voi=g_voi, g_unv=g_unv, m_ns=m_ns, m_syn=m_syn)
    return v_sig


# ----------------------------------------------------------------------------------------
# post-filter                                                  src/magphase.py:2300-2378
# ----------------------------------------------------------------------------------------
def post_filter_params(fs, mag_dim, av_len_at_zero=None, av_len_at_nyq=None, boost_at_zero=None, boost_at_nyq=None):
    """Defaults per sample rate.  src/magphase.py:2306-2340"""
    opts = [av_len_at_zero, av_len_at_nyq, boost_at_zero, boost_at_nyq]
    if fs == 48000:
        d = [round_to_int(11.0 * (mag_dim / 60.0)), round_to_int(3.0 * (mag_dim / 60.0)), 1.8, 2.0]
    elif fs == 16000:
        if any(o is None for o in opts):
            warnings.warn('Post-filter: The default parameters for 16kHz sample rate have not being tunned.')
        d = [round_to_int(9.0 * (mag_dim / 60.0)), round_to_int(12.0 * (mag_dim / 60.0)), 2.0, 1.6]
    else:
        if any(o is None for o in opts):
            raise ValueError('Post-filter: It has only been tested with 16kHz and 48kHz sample rates.')
        d = opts
    return [d[i] if opts[i] is None else opts[i] for i in range(4)]


def post_filter(m_mag_mel_log, fs, av_len_at_zero=None, av_len_at_nyq=None, boost_at_zero=None, boost_at_nyq=None):
    """src/magphase.py:2300-2378"""
    n, D = m_mag_mel_log.shape
    if D != 60:
        warnings.warn('Post-filter: It has been only tested with 60 dimensional mag data.')
    l0, l1, b0, b1 = post_filter_params(fs, D, av_len_at_zero, av_len_at_nyq, boost_at_zero, boost_at_nyq)
    v_nx = np.arange(np.floor(l0 / 2), D - np.floor(l1 / 2)).astype(int)
    v_lens = (2 * np.ceil(np.linspace(l0, l1, v_nx.size) / 2) - 1).astype(int)
    half = v_lens // 2
    ave = np.zeros((n, D))
    for j, b in enumerate(v_nx):
        ave[:, b] = np.mean(m_mag_mel_log[:, b - half[j]:b + half[j] + 1], axis=1)
    ave[:, :v_nx[0]] = ave[:, [v_nx[0]]]
    ave[:, v_nx[-1]:] = ave[:, [v_nx[-1]]]
    out = (m_mag_mel_log - ave) * np.linspace(b0, b1, D)[None, :] + ave
    out[:, 0] = m_mag_mel_log[:, 0]
    out[:, -1] = m_mag_mel_log[:, -1]
    return out


# ----------------------------------------------------------------------------------------
# minimum phase                                                src/libaudio.py:920-934
# ----------------------------------------------------------------------------------------
def build_min_phase_from_mag_spec(m_mag):
    """log|X| -> real cepstrum -> causal lifter (x2 on 1..H-2, zero from H) -> FFT -> exp.
    src/libaudio.py:920-934"""
    H = m_mag.shape[1]
    N = 2 * (H - 1)
    ceps = np.fft.irfft(log_protected(m_mag), n=N, axis=1)
    ceps[:, H:] = 0.0
    ceps[:, 1:(H - 1)] *= 2.0
    return np.exp(np.fft.fft(ceps, axis=1)[:, :H])


# ----------------------------------------------------------------------------------------
# pitch-synchronous Griffin-Lim                               src/magphase.py:3318-3373
# ----------------------------------------------------------------------------------------
def griffin_lim_initial_phase(m_mag, phase_init='random'):
    """Full-length (nfrms x fft_len) initial phase matrix.  src/magphase.py:3330-3347.
    'random' draws nfrms * fft_len numbers from NumPy's global stream (np.random.rand)."""
    nfrms, H = m_mag.shape
    N = 2 * (H - 1)
    herm = lambda ph: np.hstack((np.zeros((nfrms, 1)), ph[:, 1:-1], np.zeros((nfrms, 1)), -ph[:, -2:0:-1]))   # la.add_hermitian_half 'phase'
    if isinstance(phase_init, str):
        if phase_init == 'random':
            return 2 * np.pi * (np.random.rand(nfrms, N) - 0.5)
        if phase_init == 'linear':
            d = np.zeros((nfrms, N))
            d[:, N // 2] = 1.0
            return np.angle(np.fft.fft(d))
        if phase_init == 'min_phase':
            return herm(np.angle(build_min_phase_from_mag_spec(m_mag)))
        raise ValueError('phase_init')
    return herm(np.asarray(phase_init, dtype=np.float64))


def griffin_lim(m_mag, v_shift, phase_init='random', niters=30, win_func='hann'):
    """Pitch synchronous Griffin-Lim (win_func: np.hanning unless a callable is given).  src/magphase.py:3318-3373.
    Synthesis: ifft(mag * exp(j phase)).real rows overlap-added with the frame centre (column N/2) on the pitch marks,
    NO fftshift (:3356-3358).  Analysis: windowing() around the marks, frame placed with its mark at column N/2
    (la.frm_list_to_matrix, src/libaudio.py:122-140), fft, np.angle (:3364-3369).  Returns (v_sig, half phase)."""
    m_mag = np.asarray(m_mag, dtype=np.float64)
    v_shift = round_to_int(v_shift)
    nfrms, H = m_mag.shape
    N = 2 * (H - 1)
    m_phase = griffin_lim_initial_phase(m_mag, phase_init)
    m_mag_full = np.hstack((m_mag, m_mag[:, -2:0:-1]))                     # la.add_hermitian_half 'mag'
    v_pm = np.cumsum(v_shift)
    v_sig = None
    for it in range(niters):
        m_frms = np.fft.ifft(m_mag_full * np.exp(m_phase * 1j)).real
        v_sig = ola(m_frms, v_pm)
        if it == niters - 1:
            break
        P, v_l, v_r = frame_limits(v_pm, v_sig.size)
        m_frms = np.zeros((nfrms, N))
        for f in range(nfrms):
            l, r = int(v_l[f]), int(v_r[f])
            frm = v_sig[P[f]:P[f + 2] + 1] * asym_window(l, r, win_func)
            a = N // 2 - l                                                 # rel_shift of la.frame_shift
            if a < 0 or a + frm.size > N:
                raise ValueError('negative dimensions are not allowed')
            m_frms[f, a:a + frm.size] = frm
        m_phase = np.angle(np.fft.fft(m_frms, n=N))
    return v_sig, m_phase[:, :H]


# ----------------------------------------------------------------------------------------
# compressed analysis                                         src/magphase.py:2947-2988
# ----------------------------------------------------------------------------------------
def analysis_compressed_from_pm(v_sig, fs, v_pm_smpls, v_voi, fft_len=None, mag_dim=60, phase_dim=45,
                                b_const_rate=False, alpha_phase=None):
    """analysis_compressed (src/magphase.py:2947-2988) with REAPER's output given as input."""
    m_mag, m_real, m_imag, v_f0, fs, v_shift = analysis_lossless_from_pm(v_sig, fs, v_pm_smpls, v_voi, fft_len)
    if b_const_rate:
        v_pm = np.cumsum(v_shift)
        m_mag = interp_from_variable_to_const_frm_rate(m_mag, v_pm, 5.0, fs)
        m_real = interp_from_variable_to_const_frm_rate(m_real, v_pm, 5.0, fs)
        m_imag = interp_from_variable_to_const_frm_rate(m_imag, v_pm, 5.0, fs)
        voi = v_f0 > 1.0
        v_f0 = interp_from_variable_to_const_frm_rate(np.r_[v_f0[voi][0], v_f0[voi], v_f0[voi][-1]],
                                                      np.r_[0, v_pm[voi], v_pm[-1]], 5.0, fs)
        voi = interp_from_variable_to_const_frm_rate(voi.astype(float), v_pm, 5.0, fs) > 0.5
        v_f0 = v_f0 * voi
    feats = format_for_modelling(m_mag, m_real, m_imag, v_f0, fs, mag_dim=mag_dim, phase_dim=phase_dim,
                                 alpha_phase=alpha_phase)
    return feats + (v_shift, fs, 2 * (m_mag.shape[1] - 1))


# ----------------------------------------------------------------------------------------
# legacy v1 analysis                                  src/magphase.py:573-598 over :338-368
# ----------------------------------------------------------------------------------------
def next_pow_of_two(x):
    """src/libaudio.py next_pow_of_two"""
    if x < 2:
        x = 2
    return int(2 ** np.ceil(np.log2(x)).astype(int))


def analysis_with_del_comp_and_ph_encoding_from_pm(v_in_sig, nFFT, fs, mvf, v_pm_sec):
    """analysis_with_del_comp_and_ph_encoding with REAPER's pitch marks (seconds) given as input.
    The three la.sp_to_mcep calls go through mcep_j0 (SPTK restatement, parity unpinned)."""
    v_pm_smpls = np.asarray(v_pm_sec, dtype=np.float64) * fs
    P, v_shift, v_rights = frame_limits(v_pm_smpls, np.size(v_in_sig))
    len_max = int(np.max(v_shift + v_rights + 1))
    if nFFT < len_max:
        raise ValueError("nFFT (%d) is shorter than the maximum frame length (%d)" % (nFFT, len_max))
    m_frms, v_shift, _ = analysis_frames(v_in_sig, v_pm_smpls, nFFT)
    m_fft = np.fft.fft(m_frms)[:, :nFFT // 2 + 1]
    m_sp, m_ph = np.absolute(m_fft), np.angle(m_fft)
    m_phs, m_phc = np.sin(m_ph), np.cos(m_ph)
    m_spmgc = mcep_j0(m_sp)
    mvf_bin = int(round_to_int(mvf * nFFT / float(fs)))
    n_ph = next_pow_of_two(mvf_bin) + 1
    grid = np.linspace(0, mvf_bin - 1, n_ph)
    m_phs_i = interpolate.interp1d(np.arange(mvf_bin), m_phs[:, :mvf_bin], kind='cubic')(grid)
    m_phc_i = interpolate.interp1d(np.arange(mvf_bin), m_phc[:, :mvf_bin], kind='cubic')(grid)
    return m_spmgc, mcep_j0(m_phs_i, in_type=1), mcep_j0(m_phc_i, in_type=1), v_shift


# ------------------------------------------------------------------------------------------------
# Merlin-style post-filter (src/magphase.py:3375-3465).  PARITY UNPINNED: the reference pipes the cepstra through nine
# SPTK-3.9 binaries (x2x freqt c2acr vopr mc2b bcp sopr merge b2mc, fetched by tools/download_and_compile_tools.sh:5,
# absent here, no fixture shipped); each stage below restates the published algorithm of the binary named beside it,
# with SPTK's float32 file format at every pipe / file boundary.
# ------------------------------------------------------------------------------------------------
def _f32(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def sptk_mc2b(m_mc, alpha):
    """SPTK mc2b: b[m] = c[m]; b[i] = c[i] - alpha b[i+1]."""
    b = np.array(m_mc, dtype=np.float64)
    for i in range(b.shape[1] - 2, -1, -1):
        b[:, i] = b[:, i] - alpha * b[:, i + 1]
    return b


def sptk_b2mc(m_b, alpha):
    """SPTK b2mc: c[m] = b[m]; c[i] = b[i] + alpha b[i+1] (b as given, not the updated values)."""
    c = np.array(m_b, dtype=np.float64)
    c[:, :-1] = m_b[:, :-1] + alpha * m_b[:, 1:]
    return c


def sptk_c2acr_r0(m_c, fft_len):
    """SPTK c2acr -M 0: x = Re FFT_l(c zero padded); r[0] = mean_k exp(2 x[k])."""
    x = np.fft.fft(m_c, n=fft_len, axis=1).real
    return np.mean(np.exp(2.0 * x), axis=1)


def post_filter_merlin(m_mag_mel_log, fs, pf_coef=1.4):
    """src/magphase.py:3375-3465 (command lines :3418-3444)."""
    fft_len = 4096
    minph_ord = fft_len // 2 - 1
    alpha = define_alpha(fs)
    n = m_mag_mel_log.shape[1]
    # la.rceps(in_type='log', out_type='compact'), written as float32 (:3397-3398)
    ext = np.hstack((m_mag_mel_log, m_mag_mel_log[:, -2:0:-1]))
    ceps = np.fft.ifft(ext, axis=1).real
    ceps[:, 1:(n - 2)] *= 2
    mcep = _f32(ceps[:, :n])
    w = _f32(np.r_[1.0, 1.0, np.full(n - 2, float('%1.2f' % pf_coef))])        # echo 1 1 pf pf ... | x2x +af
    # freqt -m n-1 -a alpha -M 2047 -A 0: all-pass from warping alpha to 0, i.e. freqt with a = (0 - alpha) / (1 - 0 alpha)
    F = freqt_matrix(minph_ord + 1, n, -alpha)
    lifted = _f32(mcep * w)                                                   # vopr -m
    r0 = _f32(sptk_c2acr_r0(_f32(mcep @ F.T), fft_len))
    p_r0 = _f32(sptk_c2acr_r0(_f32(lifted @ F.T), fft_len))
    b = _f32(sptk_mc2b(lifted, alpha))                                        # mc2b
    b0 = b[:, 0]                                                              # bcp -s 0 -e 0
    p_b0 = _f32(_f32(_f32(np.log(_f32(r0 / p_r0))) / 2.0) + b0)               # vopr -d | sopr -LN -d 2 | vopr -a
    merged = np.hstack((p_b0[:, None], b[:, 1:]))                             # bcp -s 1 -e n-1 | merge -s 0
    mcep_pf = _f32(sptk_b2mc(merged, alpha))                                  # b2mc
    out = mcep_to_sp_cosmat(mcep_pf, n, alpha=0.0, out_type='log')
    out[np.isnan(out)] = MAGIC
    return out

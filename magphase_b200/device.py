"""Device-resident batch pipeline (torch tensors for HBM and streams; the arithmetic is the C ABI's).

A *plan* is the integer bookkeeping of a batch of utterances (what ``windowing()`` / ``ola()`` compute on the
host in the reference) uploaded once; signals and features then stay in HBM between analysis and
synthesis.  Used by ``bench.py`` (device-timed ``value``) and by the sharded multi-GPU driver.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from . import magphase as mp
from ._lib import MPB_F32, MPB_F64

_TORCH_DT = {MPB_F32: torch.float32, MPB_F64: torch.float64}


def _dp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class LosslessPlan:
    """Bookkeeping for analysis_lossless -> synthesis_from_lossless of a list of utterances on one GPU."""

    def __init__(self, l_nsmpls, l_pm_smpls, l_voi, fs, fft_len=None, device=None, ola_target_frames=32):
        self.fs = fs
        self.fft_len = mp.define_fft_len(fs) if fft_len is None else fft_len
        self.H = self.fft_len // 2 + 1
        self.device = torch.device('cuda', _lib.default_device() if device is None else device)
        self.ctx = _lib.ctx(self.device.index)
        n_utt = len(l_nsmpls)
        sig_off = np.zeros(n_utt + 1, dtype=np.int64)
        frm_off = np.zeros(n_utt + 1, dtype=np.int64)
        out_off = np.zeros(n_utt + 1, dtype=np.int64)
        centre, left, right, pm_int, t0, self.l_shift, self.l_f0 = [], [], [], [], [], [], []
        for u in range(n_utt):
            P, v_shift, v_rights = mp.frame_geometry(l_pm_smpls[u], l_nsmpls[u])
            sig_off[u + 1] = sig_off[u] + l_nsmpls[u]
            frm_off[u + 1] = frm_off[u] + v_shift.size
            centre.append(P[1:-1] + sig_off[u]); left.append(v_shift); right.append(v_rights)
            v_f0 = mp.shift_to_f0(v_shift, np.asarray(l_voi[u], dtype=np.float64), fs, b_smooth=False)
            p, t, n_out = mp.ola_geometry(np.cumsum(mp.f0_to_shift(v_f0, fs)), self.fft_len)
            pm_int.append(p); t0.append(t)
            out_off[u + 1] = out_off[u] + n_out
            self.l_shift.append(v_shift); self.l_f0.append(v_f0)
        self.n_utt, self.sig_off, self.frm_off, self.out_off = n_utt, sig_off, frm_off, out_off
        self.nfrm, self.n_sig, self.n_out = int(frm_off[-1]), int(sig_off[-1]), int(out_off[-1])
        pm_all = np.ascontiguousarray(np.concatenate(pm_int), dtype=np.int32)
        n_runs = C.c_int64()
        lib = _lib.lib()
        _lib.check(lib.mpb_plan_ola_runs(_lib.ptr(pm_all), _lib.ptr(frm_off), n_utt, self.fft_len, ola_target_frames,
                                         None, 0, C.byref(n_runs)))
        runs = np.zeros((max(n_runs.value, 1), 4), dtype=np.int32)
        _lib.check(lib.mpb_plan_ola_runs(_lib.ptr(pm_all), _lib.ptr(frm_off), n_utt, self.fft_len, ola_target_frames,
                                         _lib.ptr(runs), n_runs.value, C.byref(n_runs)))
        self.n_runs = int(n_runs.value)
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
        self.d_centre = up(np.concatenate(centre), np.int64)
        self.d_left = up(np.concatenate(left), np.int32)
        self.d_right = up(np.concatenate(right), np.int32)
        self.d_pm = up(pm_all, np.int32)
        self.d_out_off = up(out_off, np.int64)
        self.d_t0 = up(np.array(t0), np.int32)
        self.d_runs = up(runs, np.int32)
        self.mean_shift = float(np.mean(np.concatenate(left)))

    # algorithmic HBM bytes per launch: samples in + descriptors, features out / the reverse
    def analysis_bytes(self, sig_dtype, feat_dtype):
        es, ef = (8 if sig_dtype == MPB_F64 else 4), (8 if feat_dtype == MPB_F64 else 4)
        return self.n_sig * es + self.nfrm * 16 + 3 * self.nfrm * self.H * ef

    def synthesis_bytes(self, feat_dtype, out_dtype):
        ef, eo = (8 if feat_dtype == MPB_F64 else 4), (8 if out_dtype == MPB_F64 else 4)
        return 3 * self.nfrm * self.H * ef + self.nfrm * 4 + self.n_out * eo

    def alloc_features(self, feat_dtype=MPB_F32):
        return tuple(torch.empty((self.nfrm, self.H), dtype=_TORCH_DT[feat_dtype], device=self.device) for _ in range(3))

    def alloc_output(self, out_dtype=MPB_F32):
        return torch.empty(self.n_out, dtype=_TORCH_DT[out_dtype], device=self.device)

    def analysis(self, d_sig, feats, compute=MPB_F64):
        sig_dt = MPB_F64 if d_sig.dtype == torch.float64 else MPB_F32
        feat_dt = MPB_F64 if feats[0].dtype == torch.float64 else MPB_F32
        _lib.check(_lib.lib().mpb_analysis_lossless_dev(
            self.ctx, _stream(), _dp(d_sig), sig_dt, self.n_sig, _dp(self.d_centre), _dp(self.d_left),
            _dp(self.d_right), None, self.nfrm, self.fft_len, compute, _dp(feats[0]), _dp(feats[1]), _dp(feats[2]),
            feat_dt))

    def synthesis(self, feats, d_out, compute=MPB_F32):
        feat_dt = MPB_F64 if feats[0].dtype == torch.float64 else MPB_F32
        out_dt = MPB_F64 if d_out.dtype == torch.float64 else MPB_F32
        _lib.check(_lib.lib().mpb_synthesis_lossless_dev(
            self.ctx, _stream(), _dp(feats[0]), _dp(feats[1]), _dp(feats[2]), feat_dt, _dp(self.d_pm), self.nfrm,
            _dp(self.d_out_off), _dp(self.d_t0), self.n_utt, _dp(self.d_runs), self.n_runs, self.fft_len, compute,
            _dp(d_out), out_dt, self.n_out))


class CompressedPlan:
    """Bookkeeping for analysis_compressed -> synthesis_from_compressed (BASELINE config 2: 60/45/45, variable
    rate, no output HPF) of a list of utterances on one GPU.  Everything the host mirror computes per call
    (frame geometry, lf0, synthesis shifts, noise frame geometry, OLA runs) is computed once and uploaded."""

    def __init__(self, l_nsmpls, l_pm_smpls, l_voi, fs, fft_len=None, mag_dim=60, phase_dim=45, device=None,
                 ola_target_frames=32, noise_seed=1234, alpha_phase=None):
        self.fs = fs
        self.fft_len = mp.define_fft_len(fs) if fft_len is None else fft_len
        self.H = self.fft_len // 2 + 1
        self.mag_dim, self.phase_dim = mag_dim, phase_dim
        self.device = torch.device('cuda', _lib.default_device() if device is None else device)
        self.ctx = _lib.ctx(self.device.index)
        # alpha_phase: analysis-side warping of the phase streams (0.0 reproduces analysis_for_acoustic_modelling's
        # alpha_phase=False quirk, src/magphase.py:3010); the synthesis side always un-warps with the default (:3271)
        self.mel = mp._MelPlan.get(fs, self.fft_len, mag_dim, phase_dim, alpha_phase)
        self.syn = mp._SynPlan.get(fs, self.fft_len, mag_dim, phase_dim, None)
        n_utt = len(l_nsmpls)
        sig_off = np.zeros(n_utt + 1, dtype=np.int64)
        centre, left, right, voi8, self.l_lf0 = [], [], [], [], []
        for u in range(n_utt):
            P, v_shift, v_rights = mp.frame_geometry(l_pm_smpls[u], l_nsmpls[u])
            sig_off[u + 1] = sig_off[u] + l_nsmpls[u]
            centre.append(P[1:-1] + sig_off[u]); left.append(v_shift); right.append(v_rights)
            v_f0 = mp.shift_to_f0(v_shift, np.asarray(l_voi[u], dtype=np.float64), fs, b_smooth=False)
            v_voi, v_lf0 = mp._lf0_smoothed(v_f0)
            voi8.append(v_voi > 0); self.l_lf0.append(v_lf0)
        self.n_utt, self.n_sig = n_utt, int(sig_off[-1])
        self.nfrm = int(sum(a.size for a in left))
        self.mean_shift = float(np.mean(np.concatenate(left)))
        self.n_voiced = int(sum(int(np.count_nonzero(v)) for v in voi8))
        arrs, self.l_ns_len = mp.compressed_synthesis_geometry(self.l_lf0, [a.size for a in left], fs, self.fft_len)
        self.n_noise = int(sum(self.l_ns_len))
        self.n_out = int(arrs['utt_out_off'][-1])
        n_runs = C.c_int64()
        lib = _lib.lib()
        _lib.check(lib.mpb_plan_ola_runs(_lib.ptr(arrs['pm']), _lib.ptr(arrs['utt_frm_off']), n_utt, self.fft_len,
                                         ola_target_frames, None, 0, C.byref(n_runs)))
        runs = np.zeros((max(n_runs.value, 1), 4), dtype=np.int32)
        _lib.check(lib.mpb_plan_ola_runs(_lib.ptr(arrs['pm']), _lib.ptr(arrs['utt_frm_off']), n_utt, self.fft_len,
                                         ola_target_frames, _lib.ptr(runs), n_runs.value, C.byref(n_runs)))
        self.n_runs = int(n_runs.value)
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
        self.d_centre = up(np.concatenate(centre), np.int64)
        self.d_left = up(np.concatenate(left), np.int32)
        self.d_right = up(np.concatenate(right), np.int32)
        self.d_voi_ana = up(np.concatenate(voi8), np.uint8)
        self.d_runs = up(runs, np.int32)
        self.d_need = up(arrs.pop('need_ph'), np.uint8)
        self._keep = {k: (up(v, v.dtype) if v is not None else None) for k, v in arrs.items()}
        self.frames = _lib.SynFrames(nfrm=self.nfrm, n_utt=n_utt,
                                     **{k: (_dp(v) if v is not None else None) for k, v in self._keep.items()})
        # the aperiodic noise (np.random.uniform(-1, 1, ns_len) per utterance, src/magphase.py:883) is drawn ON THE DEVICE
        # inside synthesis(), from NumPy's legacy MT19937 state for `noise_seed` (bit-identical stream); every call re-draws
        # the same stretch, so repeated steps do the same work on the same numbers
        st = np.random.RandomState(noise_seed).get_state()
        self.mt_key = np.ascontiguousarray(st[1], dtype=np.uint32)
        self.mt_pos = int(st[2])
        self.d_noise = torch.empty(max(self.n_noise, 1), dtype=torch.float32, device=self.device)
        self.draw_noise = True
        self._side = None
        self.noise_stage_early = os.environ.get('MPB_NOISE_EARLY', '1') != '0'
        self.side_high_priority = os.environ.get('MPB_SIDE_PRIO', '1') != '0'
        # the compressed features live in HBM between the two halves
        self.d_mel = (torch.empty((self.nfrm, mag_dim), dtype=torch.float32, device=self.device),
                      torch.empty((self.nfrm, phase_dim), dtype=torch.float32, device=self.device),
                      torch.empty((self.nfrm, phase_dim), dtype=torch.float32, device=self.device))
        self.d_out = torch.empty(self.n_out, dtype=torch.float32, device=self.device)

    # algorithmic HBM bytes per pass of the two halves: what the reference API contract moves
    def analysis_bytes(self):
        return self.n_sig * 4 + self.nfrm * 17 + self.nfrm * (self.mag_dim + 2 * self.phase_dim) * 4

    def synthesis_bytes(self):
        return self.nfrm * (self.mag_dim + 2 * self.phase_dim) * 4 + self.nfrm * 45 + self.n_noise * 4 + self.n_out * 4

    def analysis(self, d_sig, compute=MPB_F64):
        """k_voiced_compact -> k_analysis<logp> (float64 butterflies) -> k_mel_warp_tc -> k_mel_cos per chunk of frames."""
        sig_dt = MPB_F64 if d_sig.dtype == torch.float64 else MPB_F32
        _lib.check(_lib.lib().mpb_analysis_compressed_dev(
            self.mel.handle, _stream(), _dp(d_sig), sig_dt, self.n_sig, _dp(self.d_centre), _dp(self.d_left),
            _dp(self.d_right), _dp(self.d_voi_ana), self.nfrm, _dp(self.d_mel[0]), _dp(self.d_mel[1]), _dp(self.d_mel[2]),
            MPB_F32))
        return self.d_mel

    def host_noise(self):
        """The same draws as a list of float64 arrays (one per utterance), for parity checks against the oracle."""
        rs = np.random.RandomState()
        rs.set_state(('MT19937', self.mt_key.copy(), self.mt_pos, 0, 0.0))
        return [rs.uniform(-1, 1, n) for n in self.l_ns_len]

    def draw(self):
        """Enqueues the MT19937 draw of the whole batch's aperiodic noise on the current stream."""
        _lib.check(_lib.lib().mpb_mt19937_fill_dev(self.ctx, _stream(), _lib.ptr(self.mt_key), self.mt_pos, self.n_noise, -1.0,
                                                   1.0, _dp(self.d_noise), MPB_F32))

    def chain(self, d_sig, compute=MPB_F64, mid_event=None):
        """analysis -> synthesis with the noise draw on a side stream NEXT TO the analysis half: the draw is one wave of
        latency-bound CTAs (the twister's recurrence), the analysis kernels fill the rest of every SM meanwhile.  The draw
        does not depend on the signal, only on NumPy's generator state, so the result is the same as analysis() followed by
        synthesis()."""
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=-1 if self.side_high_priority else 0)
            self._ev_free, self._ev_noise = torch.cuda.Event(), torch.cuda.Event()
        self._ev_free.record(main)                 # everything enqueued so far (the previous synthesis reads d_noise)
        self._side.wait_event(self._ev_free)
        with torch.cuda.stream(self._side):
            self.draw()
            if self.noise_stage_early:
                # the noise frames' FFTs (float32 pipes) beside the signal frames' FFTs (float64 pipe): neither needs the other
                _lib.check(_lib.lib().mpb_synthesis_noise_stage_dev(self.syn.handle, _stream(), _dp(self.d_noise), self.n_noise,
                                                                    C.byref(self.frames)))
            self._ev_noise.record(self._side)
        self.analysis(d_sig, compute=compute)
        if mid_event is not None:
            mid_event.record(main)
        main.wait_event(self._ev_noise)
        keep, self.draw_noise = self.draw_noise, False
        try:
            return self.synthesis()
        finally:
            self.draw_noise = keep

    def synthesis(self, d_mel=None):
        m = self.d_mel if d_mel is None else d_mel
        if self.draw_noise:
            self.draw()
        _lib.check(_lib.lib().mpb_synthesis_compressed_dev(
            self.syn.handle, _stream(), _dp(m[0]), _dp(m[1]), _dp(m[2]), MPB_F32, self.nfrm, _dp(self.d_need),
            _dp(self.d_noise), self.n_noise, C.byref(self.frames), _dp(self.d_runs), self.n_runs, 0, _dp(self.d_out),
            MPB_F32, self.n_out))
        return self.d_out


def griffin_lim(m_mag, v_shift, win_func=np.hanning, phase_init='random', niters=30, device=None):
    """Pitch-synchronous Griffin-Lim (src/magphase.py:3318-3373) with every iteration on the device: the synthesis half is
    k_synthesis_lossless, the analysis half k_analysis, both in float64, features and signal resident in HBM.

    The reference keeps its frames centred on column N/2 without fftshift, the kernels keep the pitch mark at index 0
    with fftshift on the way out; the two conventions differ by (-1)^k on the spectrum, i.e. phase_kernel = phase_ref -
    pi k, and that factor cancels between the two halves of an iteration.  It is applied to the initial phase and
    removed from the returned one.  Returns (v_sig, m_phase[nfrms, N/2+1]) like the reference."""
    m_mag = np.ascontiguousarray(m_mag, dtype=np.float64)
    if m_mag.ndim != 2:
        raise ValueError('m_mag must be nfrms x (fft_len/2+1)')
    nfrms, H = m_mag.shape
    N = 2 * (H - 1)
    if N not in (1024, 2048, 4096):
        raise ValueError('fft_len must be 1024, 2048 or 4096')
    if niters < 1:
        raise ValueError('niters must be >= 1')
    v_shift = mp.round_to_int(np.asarray(v_shift, dtype=np.float64))
    if v_shift.size != nfrms:
        raise ValueError('one shift per frame')
    v_pm = np.cumsum(v_shift)
    pm_int, t0, n_out = mp.ola_geometry(v_pm, N)
    P, left, right = mp.frame_geometry(v_pm, n_out)
    if np.any(left > N // 2) or np.any(right >= N // 2):          # la.frame_shift gets a negative pad (src/libaudio.py:137-140)
        raise ValueError('negative dimensions are not allowed')
    custom = mp._has_custom_window(win_func)
    kinds = None if custom else mp._win_codes(win_func, nfrms)

    # ---- initial spectrum, as the first synthesis sees it (src/magphase.py:3330-3347, :3356-3357) ----
    sgn = np.where(np.arange(H) % 2 == 0, 1.0, -1.0)
    if isinstance(phase_init, str) and phase_init == 'random':
        ph_full = 2 * np.pi * (np.random.rand(nfrms, N) - 0.5)     # NOT Hermitian: .real of the ifft symmetrises it
        mag_full = np.hstack((m_mag, m_mag[:, -2:0:-1]))
        x_full = mag_full * np.exp(1j * ph_full)
        x_half = 0.5 * (x_full[:, :H] + np.conj(x_full[:, (N - np.arange(H)) % N]))
        first_phase = ph_full[:, :H]
    else:
        if isinstance(phase_init, str):
            if phase_init == 'linear':
                ph = np.angle(np.tile(sgn + 0j, (nfrms, 1)))       # angle(fft(delta at N/2))
            elif phase_init == 'min_phase':
                ph = np.angle(mp.build_min_phase_from_mag_spec(m_mag))
            else:
                raise ValueError("phase_init must be 'random', 'linear', 'min_phase' or an array")
        else:
            ph = np.array(phase_init, dtype=np.float64)
            if ph.shape != m_mag.shape:
                raise ValueError('phase_init must have the shape of m_mag')
        if not (isinstance(phase_init, str) and phase_init == 'linear'):
            ph[:, 0] = 0.0                                          # la.add_hermitian_half(data_type='phase')
            ph[:, -1] = 0.0
        x_half = m_mag * np.exp(1j * ph)
        first_phase = ph
    x_half = x_half * sgn
    mag0 = np.abs(x_half)
    with np.errstate(invalid='ignore', divide='ignore'):
        u0 = np.where(mag0 > 0, x_half / mag0, 0.0)

    dev = torch.device('cuda', _lib.default_device() if device is None else device)
    ctx = _lib.ctx(dev.index)
    lib = _lib.lib()
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    frm_off = np.array([0, nfrms], dtype=np.int64)
    n_runs = C.c_int64()
    _lib.check(lib.mpb_plan_ola_runs(_lib.ptr(pm_int), _lib.ptr(frm_off), 1, N, 32, None, 0, C.byref(n_runs)))
    runs = np.zeros((max(n_runs.value, 1), 4), dtype=np.int32)
    _lib.check(lib.mpb_plan_ola_runs(_lib.ptr(pm_int), _lib.ptr(frm_off), 1, N, 32, _lib.ptr(runs), n_runs.value, C.byref(n_runs)))
    d_pm, d_runs = up(pm_int, np.int32), up(runs, np.int32)
    d_out_off, d_t0 = up(np.array([0, n_out]), np.int64), up(np.array([t0]), np.int32)
    d_centre, d_left, d_right = up(P[1:-1], np.int64), up(left, np.int32), up(right, np.int32)
    d_win = up(kinds, np.uint8) if kinds is not None else None
    n_ana = n_out
    if custom:
        # a window the kernels do not evaluate themselves (any callable, src/magphase.py:102-108): its weights and the
        # gather indices of the frames are fixed over the iterations; every analysis half runs on sig[idx] * w with weight 1
        w_all, w_off = mp.window_weights(mp._win_list(win_func, nfrms), left, right)
        idx = np.repeat(P[1:-1] - left - w_off[:-1], np.diff(w_off)) + np.arange(w_all.size, dtype=np.int64)
        d_idx, d_w = up(idx, np.int64), up(w_all, np.float64)
        d_centre = up(w_off[:-1] + left, np.int64)
        d_win = up(np.full(nfrms, _lib.WIN_RECT), np.uint8)
        n_ana = int(w_all.size)
    d_mag_t, d_mag0 = up(m_mag, np.float64), up(mag0, np.float64)
    d_re, d_im = up(u0.real, np.float64), up(u0.imag, np.float64)
    d_mag_w = torch.empty_like(d_mag_t)
    d_sig = torch.empty(max(n_out, 1), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = _stream()
        for it in range(niters):
            d_mag = d_mag0 if it == 0 else d_mag_t
            _lib.check(lib.mpb_synthesis_lossless_dev(ctx, st, _dp(d_mag), _dp(d_re), _dp(d_im), MPB_F64, _dp(d_pm), nfrms,
                                                      _dp(d_out_off), _dp(d_t0), 1, _dp(d_runs), int(n_runs.value), N, MPB_F64,
                                                      _dp(d_sig), MPB_F64, n_out))
            if it == niters - 1:
                break
            d_ana = (d_sig[:n_out][d_idx] * d_w) if custom else d_sig
            _lib.check(lib.mpb_analysis_lossless_dev(ctx, st, _dp(d_ana), MPB_F64, n_ana, _dp(d_centre), _dp(d_left),
                                                     _dp(d_right), _dp(d_win), nfrms, N, MPB_F64, _dp(d_mag_w), _dp(d_re),
                                                     _dp(d_im), MPB_F64))
        v_sig = d_sig[:n_out].cpu().numpy()
        if niters == 1:
            m_phase = first_phase
        else:
            m_phase = np.angle(sgn * (d_re.cpu().numpy() + 1j * d_im.cpu().numpy()))
    return v_sig, m_phase

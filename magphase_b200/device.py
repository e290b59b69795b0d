"""Device-resident batch pipeline (torch tensors for HBM and streams; the arithmetic is the C ABI's).

A *plan* is the integer bookkeeping of a batch of utterances (what ``windowing()`` / ``ola()`` compute on the
host in the reference) uploaded once; signals and features then stay in HBM between analysis and
synthesis.  Used by ``bench.py`` (device-timed ``value``) and by the sharded multi-GPU driver.
"""
import ctypes as C

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

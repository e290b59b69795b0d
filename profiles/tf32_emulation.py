"""CPU emulation of a 3xTF32 tensor-core version of the mel-warp tile product (mpb_mel_tc.cu, experimental).

Question: does  log-periodogram[F x 2049] . W^T[2049 x n]  keep the 1e-5 RMS bar when both operands are split into two
TF32 values (hi + lo, three products per step) and the float32 accumulator of the tensor core TRUNCATES (the
pessimistic reading of an undocumented detail) instead of rounding?  Uses the oracle (test infrastructure) on one
synthetic utterance; prints the error of the mel cepstra and of the final features for:
  fp32fma    the shipped scheme: float32 FMA inside 256-bin K slices, float64 across
  tf32x3_rn  3xTF32, accumulator rounds to nearest
  tf32x3_rz  3xTF32, accumulator truncates; kslice = bins per accumulator (None: all 2049); demean = subtract the row mean first
Run:  python profiles/tf32_emulation.py > profiles/r1b/tf32_emulation.txt   (CPU only, ~1 min)
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import magphase_oracle as orc
from magphase_b200 import synth
fs, N = 48000, 4096; H = N//2+1
sig, pm, voi = synth.synth_utterance(3, fs, 2.0)
m_mag, m_real, m_imag, v_f0, _, v_shift = orc.analysis_lossless_from_pm(sig, fs, pm, voi, fft_len=N)
nf = m_mag.shape[0]; print('frames', nf)
def wmat(ncoef, alpha):
    eye = np.eye(H)
    c = np.fft.irfft(eye, n=N, axis=1)[:, :H]; c[:, 0] *= 0.5; c[:, H-1] *= 0.5
    return c @ orc._freqt_cached(ncoef, H, float('%1.2f' % alpha)).T     # [H x ncoef]
def trunc_tf32(x32):
    return (x32.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
def trunc_f32_from_f64(x):   # round-toward-zero to fp32
    y = x.astype(np.float32)
    adj = np.abs(y.astype(np.float64)) > np.abs(x)
    y2 = np.nextafter(y, np.float32(0)); 
    return np.where(adj, y2, y)
def emul(logp, W, mode, kslice=None, demean=False):
    A = logp.astype(np.float32); B = W.astype(np.float32)
    mu = None
    if demean:
        mu = A.astype(np.float64).mean(axis=1, keepdims=True).astype(np.float32)
        A = (A - mu).astype(np.float32)
    K = A.shape[1]
    if mode == 'fp32fma':      # current scheme: fp32 sequential within 256-bin slices, f64 across
        out = np.zeros((A.shape[0], B.shape[1]))
        for s in range(0, K, 256):
            acc = np.zeros((A.shape[0], B.shape[1]), np.float32)
            for k in range(s, min(s+256, K)):
                acc = (acc.astype(np.float64) + A[:, k:k+1].astype(np.float64) * B[k:k+1, :].astype(np.float64)).astype(np.float32)
            out += acc
    else:
        Ah = trunc_tf32(A); Al = trunc_tf32((A - Ah).astype(np.float32))
        Bh = trunc_tf32(B); Bl = trunc_tf32((B - Bh).astype(np.float32))
        ks = kslice or K
        out = np.zeros((A.shape[0], B.shape[1]))
        for s in range(0, K, ks):
            acc = np.zeros((A.shape[0], B.shape[1]), np.float32)
            for k in range(s, min(s+ks, K), 8):
                e = min(k+8, s+ks, K)
                p = (Ah[:, k:e].astype(np.float64) @ Bh[k:e].astype(np.float64) + Al[:, k:e].astype(np.float64) @ Bh[k:e].astype(np.float64)
                     + Ah[:, k:e].astype(np.float64) @ Bl[k:e].astype(np.float64))
                t = acc.astype(np.float64) + p
                acc = trunc_f32_from_f64(t) if mode == 'tf32x3_rz' else t.astype(np.float32)
            out += acc
    if demean:
        out += mu.astype(np.float64) * B.astype(np.float64).sum(axis=0, keepdims=True)
    return out
for name, x, q, ncoef, alpha in (('mag', m_mag, 3, 60, 0.77), ('real', m_real, 2, 58, 0.77)):
    x32 = x.astype(np.float32).astype(np.float64)
    p = x32*x32 if q == 3 else np.exp(2*x32)
    logp = np.log(p + 1e-8)
    W = wmat(ncoef, alpha)
    mc_ref = (logp @ W).astype(np.float32).astype(np.float64)
    cosm = orc.cosine_matrix(ncoef, ncoef, 0.0) if hasattr(orc, 'cosine_matrix') else None
    def final(mc):
        mc = mc.astype(np.float32).astype(np.float64)
        o = orc.mcep_to_sp_cosmat(mc, ncoef, alpha=0.0, out_type={3:'abs',2:'log'}[q])
        return np.log(o) if q == 3 else o
    ref = final(logp @ W)
    print(name, 'logp range', logp.min(), logp.max(), 'mc0 mean', (logp@W)[:,0].mean())
    for mode, ks, dm in (('fp32fma', None, False), ('tf32x3_rn', None, False), ('tf32x3_rz', None, False), ('tf32x3_rz', 256, False),
                         ('tf32x3_rz', None, True), ('tf32x3_rz', 256, True)):
        mc = emul(logp, W, mode, ks, dm)
        out = final(mc)
        print('  %-10s kslice=%-5s demean=%-5s  mc rms err %.2e   out rms err %.2e  max %.2e' % (mode, ks, dm,
              np.sqrt(np.mean((mc - logp@W)**2)), np.sqrt(np.mean((out-ref)**2)), np.abs(out-ref).max()))

"""Index-math emulation of k_split_unwarp_tc + k_mel_unwarp_tc (layouts, descriptors, epilogue mapping)."""
import numpy as np
rng = np.random.default_rng(1)
UT_FT, UT_BT, LBO = 64, 128, 128
def kpad_of(K): return (K + 7) & ~7
def split_unwarp(U, K, npitch, kpad):
    n_tiles = (npitch + UT_BT - 1) // UT_BT
    part = UT_BT * kpad
    out = np.zeros(n_tiles * 2 * part, np.float64)
    for i in range(n_tiles * UT_BT * kpad):
        k = i % kpad; m = (i // kpad) % UT_BT; t = i // (kpad * UT_BT)
        b = t * UT_BT + m
        u = U[k, b] if (k < K and b < npitch) else 0.0
        off = t * 2 * part + (m >> 3) * (kpad // 4) * 32 + (k >> 2) * 32 + (m & 7) * 4 + (k & 3)
        out[off] = u            # "hi" (emulation keeps full value in hi, lo = 0)
        out[off + part] = 0.0
    return out
def run(K, npitch, nfrm):
    kpad = kpad_of(K)
    U = rng.standard_normal((K, npitch)); X = rng.standard_normal((nfrm, K))
    utc = split_unwarp(U, K, npitch, kpad)
    Y = np.full((nfrm, npitch), np.nan)
    a_part = UT_BT * kpad * 4; b_part = UT_FT * kpad * 4; sbo = (kpad // 4) * LBO
    for btile in range((npitch + UT_BT - 1) // UT_BT):
        b0 = btile * UT_BT
        UT = utc[btile * 2 * UT_BT * kpad:]
        As = np.zeros(2 * a_part // 4)           # float-indexed smem image
        As[:a_part // 4] = UT[:a_part // 4]; As[a_part // 4:2 * a_part // 4] = UT[a_part // 4:2 * a_part // 4]
        for ft in range((nfrm + UT_FT - 1) // UT_FT):
            f0 = ft * UT_FT; rows = min(UT_FT, nfrm - f0)
            raw = X[f0:f0 + rows].reshape(-1)
            Bs = np.zeros(2 * b_part // 4)
            for i in range(UT_FT * (kpad // 4)):
                f = i % UT_FT; k4 = i // UT_FT
                x = [raw[f * K + 4 * k4 + j] if (f < rows and 4 * k4 + j < K) else 0.0 for j in range(4)]
                off = (f >> 3) * sbo + k4 * LBO + (f & 7) * 16       # bytes
                Bs[off // 4: off // 4 + 4] = x
            D = np.zeros((UT_BT, UT_FT))
            for j in range(kpad // 8):
                a_base = j * 2 * LBO; b_base = j * 2 * LBO
                A = np.zeros((UT_BT, 8)); B = np.zeros((UT_FT, 8))
                for m in range(UT_BT):
                    for kk in range(8):
                        A[m, kk] = As[(a_base + (m // 8) * sbo + (kk // 4) * LBO + (m % 8) * 16 + (kk % 4) * 4) // 4]
                for n in range(UT_FT):
                    for kk in range(8):
                        B[n, kk] = Bs[(b_base + (n // 8) * sbo + (kk // 4) * LBO + (n % 8) * 16 + (kk % 4) * 4) // 4]
                D += A @ B.T
            for warp in range(4):
                for lane in range(32):
                    bin_ = b0 + warp * 32 + lane
                    if bin_ >= npitch: continue
                    for c in range(UT_FT // 16):
                        for j in range(16):
                            n = c * 16 + j
                            if n < rows: Y[f0 + n, bin_] = D[warp * 32 + lane, n]
    ref = X @ U
    assert not np.isnan(Y).any()
    print('K=%d np=%d nfrm=%d  max err %.2e' % (K, npitch, nfrm, np.abs(Y - ref).max()))
run(60, 260, 70)
run(45, 132, 64)
run(10, 128, 5)

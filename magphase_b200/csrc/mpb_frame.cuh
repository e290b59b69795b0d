// Device functions shared by the analysis-type kernels: side windows, frame staging, normalisation,
// and the overlap-add flush shared by the synthesis kernels.
#pragma once
#include "mpb_fft.cuh"
#include "mpb_kernels.h"

namespace mpb {

// sin(pi y) for |y| <= 0.5: odd Taylor polynomial, 12 terms (truncation 2e-17, measured max error 3.3e-16).
// A third of the instructions of cospi(): the Hann side window needs one per non-zero sample.
__device__ __forceinline__ double sinpi_half(double y) {
    const double u = y * y;
    double p = -1.0518471716932065e-11;
    p = fma(p, u, 5.392664662608129e-10);
    p = fma(p, u, -2.2948428997269873e-08);
    p = fma(p, u, 7.952054001475513e-07);
    p = fma(p, u, -2.1915353447830217e-05);
    p = fma(p, u, 0.00046630280576761255);
    p = fma(p, u, -0.0073704309457143504);
    p = fma(p, u, 0.08214588661112823);
    p = fma(p, u, -0.5992645293207921);
    p = fma(p, u, 2.5501640398773455);
    p = fma(p, u, -5.16771278004997);
    p = fma(p, u, 3.141592653589793);
    return p * y;
}

// Hann side window in float32: 0.5 + 0.5 cos(pi x) for x in [0, 1], as 0.5 + 0.5 sin(pi (0.5 - x)) with the odd Taylor
// polynomial of sin(pi y) on |y| <= 0.5 (6 terms: truncation 6e-8).  A third of the instructions of cospif(), which
// carries range reduction for arbitrary arguments; the float32 noise / anti-ringing windows need one per sample.
__device__ __forceinline__ float hann_side_f32(float x) {
    const float y = 0.5f - x;
    const float u = y * y;
    float p = -0.0073704309f;
    p = fmaf(p, u, 0.082145887f);
    p = fmaf(p, u, -0.59926453f);
    p = fmaf(p, u, 2.5501640f);
    p = fmaf(p, u, -5.1677128f);
    p = fmaf(p, u, 3.1415927f);
    return fmaf(0.5f * p, y, 0.5f);
}

// value of the side window at distance j from the peak; inv_s = 1 / side length (j <= side length):
//   Hann       : 0.5 + 0.5 cos(pi j / S) = 0.5 - 0.5 sin(pi (j/S - 0.5))   (np.hanning(2S+1) halves; hanning(1) = [1])
//   Bartlett2.5: (1 - j/S)^2.5                                           (np.bartlett(2S+1)**2.5 halves)
//   Rect       : 1 -- the caller's samples already carry the window (arbitrary win_func callables, evaluated on the
//                host and multiplied in by the mirror module: src/magphase.py:102-108)
template <typename T>
__device__ __forceinline__ T side_window(int j, double inv_s, int kind) {
    if (j == 0) return (T)1;
    if (sizeof(T) == 4) {
        // float32 frames (the noise branch of compressed synthesis): float32 window arithmetic is enough
        const float x = (float)j * (float)inv_s;
        if (kind == MPB_WIN_HANN) return (T)hann_side_f32(x);
        if (kind == MPB_WIN_RECT) return (T)1;
        const float b = 1.0f - x;
        return (T)(b * b * sqrtf(b));
    }
    const double x = (double)j * inv_s;
    if (kind == MPB_WIN_HANN) return (T)(0.5 - 0.5 * sinpi_half(x - 0.5));
    if (kind == MPB_WIN_RECT) return (T)1;
    const double b = 1.0 - x;
    return (T)(b * b * sqrt(b));
}

// mag = |X|, re = Re X/|X|, im = Im X/|X|, all 0 where |X| == 0 (src/magphase.py:459-470).  One reciprocal
// square root instead of hypot + divide: float seed refined by a Newton step in the compute precision.
__device__ __forceinline__ void normalise(double x, double y, double& mag, double& re, double& im) {
    const double p = x * x + y * y;
    if (p > 1e-30 && p < 1e30) {
        double r = (double)rsqrtf((float)p);
        r = r * fma(-0.5 * p, r * r, 1.5);
        r = r * fma(-0.5 * p, r * r, 1.5);
        mag = p * r; re = x * r; im = y * r;
    } else if (p == 0.0 && x == 0.0 && y == 0.0) {
        mag = re = im = 0.0;
    } else {                                  // out-of-range magnitudes: slow exact path
        mag = hypot(x, y);
        re = x / mag; im = y / mag;
    }
}
__device__ __forceinline__ void normalise(float x, float y, float& mag, float& re, float& im) {
    const float p = x * x + y * y;
    if (p > 1e-30f && p < 1e30f) {
        float r = rsqrtf(p);
        r = r * fmaf(-0.5f * p, r * r, 1.5f);
        mag = p * r; re = x * r; im = y * r;
    } else if (x == 0.0f && y == 0.0f) {
        mag = re = im = 0.0f;
    } else {
        mag = hypotf(x, y);
        re = x / mag; im = y / mag;
    }
}

// float32 outputs: once X[k] is known to float64 accuracy, the normalisation itself only needs float32 arithmetic
// (relative error ~2e-7, the float32 storage already rounds at 6e-8).  Magnitudes outside the safe float32 range take
// the exact float64 path.  Keeps ~18 instructions per bin off the FP64 pipe.
__device__ __forceinline__ void normalise_to_f32(double x, double y, float& mag, float& re, float& im, float& pw) {
    const float xf = (float)x, yf = (float)y;
    const float p = fmaf(xf, xf, yf * yf);
    if (p > 1e-30f && p < 1e30f) {
        float r = rsqrtf(p);
        r = r * fmaf(-0.5f * p, r * r, 1.5f);
        mag = p * r; re = xf * r; im = yf * r; pw = p;
    } else {
        double m, a, b;
        normalise(x, y, m, a, b);
        mag = (float)m; re = (float)a; im = (float)b; pw = mag * mag;
    }
}
__device__ __forceinline__ void normalise_to_f32(float x, float y, float& mag, float& re, float& im, float& pw) {
    normalise(x, y, mag, re, im);
    pw = mag * mag;
}

// Stage the windowed, un-delayed frame b[k] into shared memory as the packed complex
// sequence z[m] = b[2m] + i b[2m+1] (natural padded layout) and pull this thread's 16 points into registers.
//   b[N-j] = sig[c-j] * w(j, l)   j = 1..l          (left part; has priority, which also reproduces the
//   b[k]   = sig[c+k] * w(k, q)   k = 0..q_eff       truncation branch src/magphase.py:313-315)
// Only the l + q_eff + 1 non-zero samples are touched (~18 % of N for speech); everything else is known
// to be zero from the frame geometry and never goes through shared memory.
// 1 / s for an integer side length s >= 1 in float64: float32 reciprocal refined by two Newton steps (four DFMA instead of
// the ~40 instructions of an IEEE division; within one ulp of it, far inside every tolerance of the path)
__device__ __forceinline__ double recip_len(int s) {
    const double d = (double)s;
    double r = (double)__frcp_rn((float)s);
    r = fma(r, fma(-d, r, 1.0), r);
    r = fma(r, fma(-d, r, 1.0), r);
    return r;
}

template <typename T, typename TS, int N>
__device__ __forceinline__ void load_frame(const TS* __restrict__ sig, int64_t n_sig, int64_t c, int l, int q, int kind,
                                           cx<T>* __restrict__ buf, cx<T>* v, int t, bool ends_only = false) {
    using G = FftGeom<T, N>;
    const double inv_l = l > 0 ? recip_len(l) : 0.0, inv_q = q > 0 ? recip_len(q) : 0.0;
    if (ends_only) {
        // l <= 2 S1 and q < 2 S1 (uniform over the CTA; most speech frames): only z[t] (b[2t], b[2t+1], right part) and
        // z[M - S1 + t] (b[N - 2 S1 + 2t], +1, left part) can be non-zero -- this thread's own first and last butterfly
        // input.  Straight from global memory into registers: no staging, no barrier, no index search.
        auto right_part = [&](int k) -> T {                       // b[k] = sig[c + k] w(k, q), k <= q
            const int64_t i = c + k;
            return (k <= q && i >= 0 && i < n_sig) ? (T)sig[i] * side_window<T>(k, inv_q, kind) : (T)0;
        };
        auto left_part = [&](int j) -> T {                        // b[N - j] = sig[c - j] w(j, l), 1 <= j <= l
            const int64_t i = c - j;
            return (j >= 1 && j <= l && i >= 0 && i < n_sig) ? (T)sig[i] * side_window<T>(j, inv_l, kind) : (T)0;
        };
        const int j0 = 2 * G::S1 - 2 * t;                         // b[N - 2 S1 + 2t] is j = 2 S1 - 2t; the next sample j - 1
#pragma unroll
        for (int n1 = 1; n1 < 15; ++n1) v[n1] = mk<T>((T)0, (T)0);
        v[0] = mk<T>(right_part(2 * t), right_part(2 * t + 1));
        v[15] = mk<T>(left_part(j0), left_part(j0 - 1));
        return;
    }
    T* bufT = reinterpret_cast<T*>(buf);
    // l >= N (pitch period longer than the FFT): the reference keeps the first N samples of the frame and its
    // hstack((v[l:], v[:l])) rotation degenerates to the identity -> b[k] = sig[c-l+k] * w(l-k, l)
    const bool whole = l >= N;
    const int q_eff = whole ? -1 : min(q, N - l - 1);
    const int total = whole ? N : l + q_eff + 1;
    for (int idx = t; idx < total; idx += G::TPB) {
        int k, dist;
        double inv_s;
        if (whole)        { dist = l - idx; inv_s = inv_l; k = idx; }
        else if (idx < l) { dist = l - idx; inv_s = inv_l; k = N - dist; }
        else              { dist = idx - l; inv_s = inv_q; k = dist; }
        const int64_t i = c - l + idx;
        const T x = (i >= 0 && i < n_sig) ? (T)sig[i] * side_window<T>(dist, inv_s, kind) : (T)0;
        bufT[2 * G::nphys(k >> 1) + (k & 1)] = x;
    }
    // complete the two complex elements that straddle the edges of the non-zero ranges
    if (t == 0 && !whole) {
        const int ke = q_eff + 1;                 // first zero after the right part
        if ((ke & 1) && ke < N - l) bufT[2 * G::nphys(ke >> 1) + 1] = (T)0;
        const int ks = N - l;                     // first sample of the left part
        if ((ks & 1) && ks - 1 > q_eff) bufT[2 * G::nphys(ks >> 1)] = (T)0;
    }
    __syncthreads();
    const cx<T>* pk = buf + G::nphys(t);
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const int m = n1 * G::S1 + t;
        const bool nz = whole || (2 * m <= q_eff) || (2 * m + 1 >= N - l);
        v[n1] = nz ? pk[n1 * (G::S1 + G::S1 / 16)] : mk<T>((T)0, (T)0);
    }
    __syncthreads();
}


// Flush the finished part [lo, hi) of the circular overlap-add accumulator to HBM (and clear it).  Samples a
// neighbouring run may also write (outside [own_lo, own_hi)) are combined with atomicAdd on the zero-initialised
// output; everything else is a plain store.  t0 / out_len map the pitch-mark axis to the utterance's output.
template <typename T, typename TO, int N, int TPB>
__device__ __forceinline__ void ola_flush(T* __restrict__ acc, int lo, int hi, int own_lo, int own_hi, int t0,
                                          int64_t out_len, TO* __restrict__ out_utt, int t) {
    for (int pos = lo + t; pos < hi; pos += TPB) {
        const int ai = pos & (N - 1);
        const T x = acc[ai];
        acc[ai] = (T)0;
        const int64_t j = (int64_t)pos - t0;
        if (j >= 0 && j < out_len) {
            if (pos >= own_lo && pos < own_hi) out_utt[j] = (TO)x;
            else atomicAdd(&out_utt[j], (TO)x);
        }
    }
}

}  // namespace mpb

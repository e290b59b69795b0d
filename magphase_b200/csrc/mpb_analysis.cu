// Pitch-synchronous lossless analysis kernel.
//
// One CTA of M/16 threads transforms one frame at a time (grid-stride over frames):
//   gather samples around the pitch mark -> asymmetric window on the fly -> circular un-delay (pitch mark
//   at index 0) -> real FFT of N points (mpb_fft.cuh) -> |X|, Re X/|X|, Im X/|X| -> coalesced stores.
// Reference: windowing() src/magphase.py:74-119, analysis_with_del_comp_from_pm :266-334 (pad :311-315,
// rotate :323, fft :325, half :330-332), compute_lossless_feats :457-476, la.gen_non_symmetric_win
// src/libaudio.py:70-84, voi_noise_window src/magphase.py:67-69.
#include "mpb_fft.cuh"
#include "mpb_kernels.h"

namespace mpb {

// value of the side window at distance j from the peak, side length S (j <= S):
//   Hann       : 0.5 + 0.5 cos(pi j / S)        (np.hanning(2S+1) halves; hanning(1) = [1])
//   Bartlett2.5: (1 - j/S)^2.5                  (np.bartlett(2S+1)**2.5 halves)
template <typename T>
__device__ __forceinline__ T side_window(int j, int S, int kind) {
    if (j == 0) return (T)1;
    const double x = (double)j / (double)S;
    if (kind == MPB_WIN_HANN) return (T)(0.5 + 0.5 * cospi(x));
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

// Stage the windowed, un-delayed frame b[k] into shared memory as the packed complex
// sequence z[m] = b[2m] + i b[2m+1] (natural padded layout) and pull this thread's 16 points into registers.
//   b[N-j] = sig[c-j] * w(j, l)   j = 1..l          (left part; has priority, which also reproduces the
//   b[k]   = sig[c+k] * w(k, q)   k = 0..q_eff       truncation branch src/magphase.py:313-315)
// Only the l + q_eff + 1 non-zero samples are touched (~18 % of N for speech); everything else is known
// to be zero from the frame geometry and never goes through shared memory.
template <typename T, typename TS, int N>
__device__ __forceinline__ void load_frame(const TS* __restrict__ sig, int64_t n_sig, int64_t c, int l, int q, int kind,
                                           cx<T>* __restrict__ buf, cx<T>* v, int t) {
    using G = FftGeom<T, N>;
    T* bufT = reinterpret_cast<T*>(buf);
    // l >= N (pitch period longer than the FFT): the reference keeps the first N samples of the frame and its
    // hstack((v[l:], v[:l])) rotation degenerates to the identity -> b[k] = sig[c-l+k] * w(l-k, l)
    const bool whole = l >= N;
    const int q_eff = whole ? -1 : min(q, N - l - 1);
    const int total = whole ? N : l + q_eff + 1;
    for (int idx = t; idx < total; idx += G::TPB) {
        int k, dist, side;
        if (whole)        { dist = l - idx; side = l; k = idx; }
        else if (idx < l) { dist = l - idx; side = l; k = N - dist; }
        else              { dist = idx - l; side = q; k = dist; }
        const int64_t i = c - l + idx;
        const T x = (i >= 0 && i < n_sig) ? (T)sig[i] * side_window<T>(dist, side, kind) : (T)0;
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
#pragma unroll
    const cx<T>* pk = buf + G::nphys(t);
    for (int n1 = 0; n1 < 16; ++n1) {
        const int m = n1 * G::S1 + t;
        const bool nz = whole || (2 * m <= q_eff) || (2 * m + 1 >= N - l);
        v[n1] = nz ? pk[n1 * (G::S1 + G::S1 / 16)] : mk<T>((T)0, (T)0);
    }
    __syncthreads();
}

template <typename T, int N> struct KernelCfg {
    // register budget: 128 regs/thread for float64 butterflies, ~85 for float32
    static constexpr int MINB = (sizeof(T) == 8 ? 512 : 768) / FftGeom<T, N>::TPB;
};

template <typename T, typename TS, typename TO, int N, int MODE>
__global__ void __launch_bounds__(FftGeom<T, N>::TPB, KernelCfg<T, N>::MINB)
k_analysis(const TS* __restrict__ sig, int64_t n_sig,
           const int64_t* __restrict__ centre, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
           const uint8_t* __restrict__ win, int64_t nfrm, const cx<T>* __restrict__ tw,
           TO* __restrict__ out_a, TO* __restrict__ out_b, TO* __restrict__ out_c) {
    using G = FftGeom<T, N>;
    using T2 = cx<T>;
    constexpr int M = G::M, H = M + 1, TPB = G::TPB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T2* buf = reinterpret_cast<T2*>(smem_raw);
    const int t = threadIdx.x;
    FftCtx<T> fc;
    fft_setup<T, N, false>(fc, buf + G::BUF_ELEMS, tw, t);
    constexpr int STEP = TPB + TPB / 16;          // nphys(k + TPB) - nphys(k)
    const T2* pk = buf + G::nphys(t);
    const T2* pmk = buf + G::nphys(M - t);

    for (int64_t f = blockIdx.x; f < nfrm; f += gridDim.x) {
        const int64_t c = centre[f];
        const int l = left[f], q = right[f];
        const int kind = win ? (int)win[f] : MPB_WIN_HANN;

        T2 v[16];
        load_frame<T, TS, N>(sig, n_sig, c, l, q, kind, buf, v, t);
        fft_m<T, N, false>(v, buf, fc, t);

        // real-FFT split:  X[k] = E + W_N^k O,  X[M-k] = conj(E - W_N^k O),
        //                  E = (Z[k] + conj Z[M-k]) / 2,  O = -i (Z[k] - conj Z[M-k]) / 2,   k = t + j*TPB
        TO* oa = out_a + f * (int64_t)H * (MODE == MODE_FFT ? 2 : 1);
        TO* ob = out_b + f * (int64_t)H;
        TO* oc = out_c + f * (int64_t)H;
        T2 w = fc.wp;
        constexpr int NJ = (M / 2) / TPB;
#pragma unroll 2
        for (int j = 0; j <= NJ; ++j) {
            const int k = t + j * TPB;
            if (j == NJ && t != 0) break;          // k == M/2 is handled by thread 0 only
            const T2 zk = pk[j * STEP];
            const T2 zm = cconj(k == 0 ? buf[0] : pmk[-j * STEP]);
            const T2 e = mk<T>((T)0.5 * (zk.x + zm.x), (T)0.5 * (zk.y + zm.y));
            const T2 d = mk<T>((T)0.5 * (zk.x - zm.x), (T)0.5 * (zk.y - zm.y));
            const T2 wo = cmul(mk<T>(d.y, -d.x), w);
            w = cmul(w, fc.wstep);
            T2 x1 = cadd(e, wo);            // X[k]
            T2 x2 = cconj(csub(e, wo));     // X[M-k]
            if (k == 0) { x1.y = (T)0; x2.y = (T)0; }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const T2 x = h ? x2 : x1;
                const int kk = h ? (M - k) : k;
                if (h && kk == k) break;    // k == M/2 pairs with itself
                if (MODE == MODE_FFT) {
                    __stcs(&oa[2 * kk], (TO)x.x);
                    __stcs(&oa[2 * kk + 1], (TO)x.y);
                } else {
                    T mag, re, im;
                    normalise(x.x, x.y, mag, re, im);
                    __stcs(&oa[kk], (TO)mag);
                    __stcs(&ob[kk], (TO)re);
                    __stcs(&oc[kk], (TO)im);
                }
            }
        }
        __syncthreads();   // buf is rewritten by the next frame's staging
    }
}

template <typename T, typename TS, typename TO, int N, int MODE>
static cudaError_t launch_analysis_t(const AnalysisArgs& a, cudaStream_t st) {
    using G = FftGeom<T, N>;
    const size_t smem = sizeof(cx<T>) * (G::BUF_ELEMS + G::TW2_ELEMS);
    auto kern = k_analysis<T, TS, TO, N, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G::TPB, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)a.num_sms * per_sm;
    if (grid > a.nfrm) grid = a.nfrm;
    if (grid < 1) return cudaSuccess;
    kern<<<(unsigned)grid, G::TPB, smem, st>>>((const TS*)a.sig, a.n_sig, a.centre, a.left, a.right, a.win, a.nfrm,
                                               (const cx<T>*)a.tw, (TO*)a.out_a, (TO*)a.out_b, (TO*)a.out_c);
    return cudaGetLastError();
}

template <typename T, typename TS, typename TO, int N>
static cudaError_t launch_analysis_m(const AnalysisArgs& a, cudaStream_t st) {
    return a.mode == MODE_FFT ? launch_analysis_t<T, TS, TO, N, MODE_FFT>(a, st)
                              : launch_analysis_t<T, TS, TO, N, MODE_FEATS>(a, st);
}

template <typename T, typename TS, typename TO>
static cudaError_t launch_analysis_n(const AnalysisArgs& a, cudaStream_t st) {
    switch (a.fft_len) {
        case 1024: return launch_analysis_m<T, TS, TO, 1024>(a, st);
        case 2048: return launch_analysis_m<T, TS, TO, 2048>(a, st);
        case 4096: return launch_analysis_m<T, TS, TO, 4096>(a, st);
    }
    return cudaErrorInvalidValue;
}

template <typename T>
static cudaError_t launch_analysis_io(const AnalysisArgs& a, cudaStream_t st) {
    if (a.sig_dtype == MPB_F32 && a.out_dtype == MPB_F32) return launch_analysis_n<T, float, float>(a, st);
    if (a.sig_dtype == MPB_F32 && a.out_dtype == MPB_F64) return launch_analysis_n<T, float, double>(a, st);
    if (a.sig_dtype == MPB_F64 && a.out_dtype == MPB_F32) return launch_analysis_n<T, double, float>(a, st);
    return launch_analysis_n<T, double, double>(a, st);
}

cudaError_t launch_analysis(const AnalysisArgs& a, cudaStream_t st) {
    return a.compute_dtype == MPB_F64 ? launch_analysis_io<double>(a, st) : launch_analysis_io<float>(a, st);
}

}  // namespace mpb

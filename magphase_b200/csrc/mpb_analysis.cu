// Pitch-synchronous lossless analysis kernel.
//
// One CTA of M/16 threads transforms one frame at a time (grid-stride over frames):
//   gather samples around the pitch mark -> asymmetric window on the fly -> circular un-delay (pitch mark
//   at index 0) -> real FFT of N points (mpb_fft.cuh) -> |X|, Re X/|X|, Im X/|X| -> coalesced stores.
// Reference: windowing() src/magphase.py:74-119, analysis_with_del_comp_from_pm :266-334 (pad :311-315,
// rotate :323, fft :325, half :330-332), compute_lossless_feats :457-476, la.gen_non_symmetric_win
// src/libaudio.py:70-84, voi_noise_window src/magphase.py:67-69.
#include "mpb_frame.cuh"

namespace mpb {

template <typename T, int N> struct KernelCfg {
    // CTAs per SM, measured: float64 butterflies 5 (96 registers; 4 and 6 are both slower), float32 7 (72 registers)
    static constexpr int MINB = (sizeof(T) == 8 ? 640 : 896) / FftGeom<T, N>::TPB;
};

template <typename T, typename TS, typename TO, int N, int MODE>
__global__ void __launch_bounds__(FftGeom<T, N>::TPB, KernelCfg<T, N>::MINB)
k_analysis(const TS* __restrict__ sig, int64_t n_sig,
           const int64_t* __restrict__ centre, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
           const uint8_t* __restrict__ win, int64_t nfrm, const cx<T>* __restrict__ tw,
           TO* __restrict__ out_a, TO* __restrict__ out_b, TO* __restrict__ out_c,
           const uint8_t* __restrict__ ph_mask, int row_pitch, const int32_t* __restrict__ ph_row) {
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
        // MODE_LOGP: the phase streams of unvoiced frames are masked downstream (src/magphase.py:2527-2542) and never
        // read by the tile product: skip their two rows (a third of this kernel's epilogue and two thirds of its stores)
        bool need_ph = MODE != MODE_LOGP || !ph_mask || ph_mask[f] != 0;
        // MODE_LOGP with a row map: the phase rows are compacted to the voiced frames (the tensor-core product reads them
        // through a tensor map, which cannot gather); rows are pitched to a multiple of 16 bytes
        int64_t prow = f;
        if (MODE == MODE_LOGP && ph_row) { prow = ph_row[f]; need_ph = prow >= 0; }

        T2 v[16];
        // frames no longer than 2/16 of the FFT on either side of the mark only fill the first and the last sixteenth
        // of the buffer: the first radix-16 butterfly sees two non-zero inputs, read straight from global memory
        const bool ends_only = l <= 2 * G::S1 && min(q, N - l - 1) < 2 * G::S1;
        load_frame<T, TS, N>(sig, n_sig, c, l, q, kind, buf, v, t, ends_only);
        fft_m<T, N, false>(v, buf, fc, t, ends_only);

        // real-FFT split:  X[k] = E + W_N^k O,  X[M-k] = conj(E - W_N^k O),
        //                  E = (Z[k] + conj Z[M-k]) / 2,  O = -i (Z[k] - conj Z[M-k]) / 2,   k = t + j*TPB
        const int64_t pitch = MODE == MODE_LOGP ? (int64_t)row_pitch : (int64_t)H;
        TO* oa = out_a + (MODE == MODE_LOGSQ ? 0 : f * pitch * (MODE == MODE_FFT ? 2 : 1));
        TO* ob = out_b + (MODE == MODE_LOGSQ ? 0 : (need_ph ? prow : 0) * pitch);
        TO* oc = out_c + (MODE == MODE_LOGSQ ? 0 : (need_ph ? prow : 0) * pitch);
        T2 w = fc.wp;
        constexpr int NJ = (M / 2) / TPB;
        double lsum = 0.0;                         // MODE_LOGSQ: sum over bins 1..H-2 of (log|X|)^2
        float lsum_f = 0.0f;
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
                if (MODE == MODE_LOGP) {
                    // log periodograms exactly as the mel product consumes them (SPTK mcep -q 3 / -q 2 with -e 1e-8):
                    // log(|X|^2 + 1e-8) and log(exp(2 Re/|X|) + 1e-8); SFU work that hides under the FP64 butterflies
                    // log(exp(2u) + 1e-8) = 2u + log1p(1e-8 exp(-2u)) = 2u + 1e-8 exp(-2u) to 1e-15 (|u| <= 1): one exp
                    // instead of exp + log
                    // (float32 from here: once X[k] is known to float64 accuracy the logs only need float32 arithmetic)
                    const float xf = (float)x.x, yf = (float)x.y;
                    const float pw = fmaf(xf, xf, yf * yf);
                    if (pw > 1e-30f && pw < 1e30f) {
                        __stcs(&oa[kk], (TO)__logf(pw + 1.0e-8f));
                        if (need_ph) {       // unvoiced frames (half of them) stop above: no normalisation at all
                            const float r = rsqrtf(pw);          // 2 ulp: 2e-7 on values that enter a log averaged over 2048 bins
                            const float re = xf * r, im = yf * r;
                            __stcs(&ob[kk], (TO)fmaf(1.0e-8f, __expf(-2.0f * re), 2.0f * re));
                            __stcs(&oc[kk], (TO)fmaf(1.0e-8f, __expf(-2.0f * im), 2.0f * im));
                        }
                    } else {                 // out-of-range magnitudes (exact zeros included): exact path
                        float mag, re, im, p2;
                        normalise_to_f32(x.x, x.y, mag, re, im, p2);
                        __stcs(&oa[kk], (TO)__logf(p2 + 1.0e-8f));
                        if (need_ph) {
                            __stcs(&ob[kk], (TO)fmaf(1.0e-8f, __expf(-2.0f * re), 2.0f * re));
                            __stcs(&oc[kk], (TO)fmaf(1.0e-8f, __expf(-2.0f * im), 2.0f * im));
                        }
                    }
                } else if (MODE == MODE_LOGSQ) {
                    // noise frames: the spectrum itself goes to HBM for k_synthesis_compressed (rows pitched to M + 2)
                    if (sizeof(T) == 4 && out_b)
                        reinterpret_cast<float2*>(out_b)[f * (int64_t)(M + 2) + kk] = make_float2((float)x.x, (float)x.y);
                    if (kk != 0 && kk != M) {
                        const T p = x.x * x.x + x.y * x.y;
                        if (sizeof(T) == 4 && p > (T)0) {      // float32 noise frames: fast log, ~17 terms per thread in float
                            const float lg = 0.5f * __logf((float)p);
                            lsum_f = fmaf(lg, lg, lsum_f);
                        } else {
                            const double lg = p > (T)0 ? 0.5 * (double)log(p) : -1.0e10;   // la.log floor (src/libaudio.py:241-248)
                            lsum = fma(lg, lg, lsum);
                        }
                    }
                } else if (MODE == MODE_FFT) {
                    __stcs(&oa[2 * kk], (TO)x.x);
                    __stcs(&oa[2 * kk + 1], (TO)x.y);
                } else if (sizeof(TO) == 4) {
                    float mag, re, im, pw;
                    normalise_to_f32(x.x, x.y, mag, re, im, pw);
                    __stcs(&oa[kk], (TO)mag);
                    __stcs(&ob[kk], (TO)re);
                    __stcs(&oc[kk], (TO)im);
                } else {
                    T mag, re, im;
                    normalise(x.x, x.y, mag, re, im);
                    __stcs(&oa[kk], (TO)mag);
                    __stcs(&ob[kk], (TO)re);
                    __stcs(&oc[kk], (TO)im);
                }
            }
        }
        if (MODE == MODE_LOGSQ) {                  // deterministic block reduction -> out_a[f]
            __shared__ double red[32];
            lsum += (double)lsum_f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
            if ((t & 31) == 0) red[t >> 5] = lsum;
            __syncthreads();
            if (t == 0) {
                double s = 0.0;
                for (int i = 0; i < TPB / 32; ++i) s += red[i];
                out_a[f] = (TO)s;
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
                                               (const cx<T>*)a.tw, (TO*)a.out_a, (TO*)a.out_b, (TO*)a.out_c, a.ph_mask,
                                               a.row_pitch > 0 ? a.row_pitch : a.fft_len / 2 + 1, a.ph_row);
    return cudaGetLastError();
}

template <typename T, typename TS, typename TO, int N>
static cudaError_t launch_analysis_m(const AnalysisArgs& a, cudaStream_t st) {
    return a.mode == MODE_FFT ? launch_analysis_t<T, TS, TO, N, MODE_FFT>(a, st)
                              : launch_analysis_t<T, TS, TO, N, MODE_FEATS>(a, st);
}

// float32 log periodograms for the fused compressed analysis (float64 butterflies, float32 or float64 signal).
// T = float is an experiment (MPB_LOGP_F32=1, off by default, never run): the mel products average the per-bin round-off
// of float32 butterflies, a CPU emulation stays 40x under the 1e-5 bar (profiles/f32_fft_compressed_emulation.py).
template <typename T, typename TS>
static cudaError_t launch_logp_n(const AnalysisArgs& a, cudaStream_t st) {
    switch (a.fft_len) {
        case 1024: return launch_analysis_t<T, TS, float, 1024, MODE_LOGP>(a, st);
        case 2048: return launch_analysis_t<T, TS, float, 2048, MODE_LOGP>(a, st);
        case 4096: return launch_analysis_t<T, TS, float, 4096, MODE_LOGP>(a, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_analysis_logp(const AnalysisArgs& a, cudaStream_t st) {
    if (a.compute_dtype == MPB_F32)
        return a.sig_dtype == MPB_F64 ? launch_logp_n<float, double>(a, st) : launch_logp_n<float, float>(a, st);
    return a.sig_dtype == MPB_F64 ? launch_logp_n<double, double>(a, st) : launch_logp_n<double, float>(a, st);
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

// sum over bins 1..H-2 of (log|N|)^2 per frame (float32 butterflies, float32 signal, float64 sums)
cudaError_t launch_noise_stats(const AnalysisArgs& a, cudaStream_t st) {
    switch (a.fft_len) {
        case 1024: return launch_analysis_t<float, float, double, 1024, MODE_LOGSQ>(a, st);
        case 2048: return launch_analysis_t<float, float, double, 2048, MODE_LOGSQ>(a, st);
        case 4096: return launch_analysis_t<float, float, double, 4096, MODE_LOGSQ>(a, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_analysis(const AnalysisArgs& a, cudaStream_t st) {
    return a.compute_dtype == MPB_F64 ? launch_analysis_io<double>(a, st) : launch_analysis_io<float>(a, st);
}

}  // namespace mpb

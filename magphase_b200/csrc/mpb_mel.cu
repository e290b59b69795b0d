// Low-dimensional mel compression of the lossless features (analysis side of format_for_modelling).
//
// Reference: format_for_modelling src/magphase.py:2490-2544 -> la.sp_mel_warp src/libaudio.py:643-661 ->
// la.sp_to_mcep :575-601 (SPTK `mcep -a alpha -m n-1 -l N -e 1.0E-8 -j 0 -f 0.0 -q in_type`, float32 file I/O)
// -> la.mcep_to_sp_cosmat(alpha=0) :605-631.
//
// With `-j 0` SPTK's output is its initial estimate, which is LINEAR in the log periodogram:
//     mc = freqt_alpha( halve_ends( irfft( log(amp(x)^2 + 1e-8) ) ) )  =  W . logp,    W = A_alpha . C
// (C: the cosine IFFT matrix with the halved c[0], c[N/2]; A_alpha: SPTK freqt as a matrix).  So the hot
// loop is  MC[F x n] = log-periodogram[F x H] . W^T[H x n]  per stream (mag: in_type 3, real/imag: in_type 2)
// -- a tall-skinny product with K = H = N/2+1.  There is no cepstrum / IFFT left in the path.
//   k_build_warp : builds W^T (H x n, float64 -> float32) on the GPU: one thread per spectral bin runs the
//                  freqt recursion on that bin's cosine column.
//   k_mel_gemm   : CUDA-core FMA tile kernel, split over K in slices of MEL_KSLICE bins; float32 products and
//                  accumulation inside a slice (the operands are float32 data anyway), partial sums to HBM.
//   k_mel_finish : float64 sum over the K-slices -> float32 rounding (SPTK's float32 output file) ->
//                  cosine matrix in float64 -> voicing mask / clip / log.
#include "mpb_kernels.h"

namespace mpb {

// ---- W^T builder ------------------------------------------------------------------------------
// thread k: sequence c_k[n] = C[n][k] = wk/N * cos(2 pi k n / N) * (n == 0 || n == N/2 ? 0.5 : 1), n = 0..H-1
// pushed through the all-pass recursion (SPTK freqt: inputs consumed from n = H-1 down to 0).
__global__ void k_build_warp(int fft_len, int n_out, double alpha, double* __restrict__ wt64, int ld) {
    const int H = fft_len / 2 + 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= H) return;
    double g[MEL_MAX_COEFFS], d[MEL_MAX_COEFFS];
    for (int j = 0; j < n_out; ++j) g[j] = 0.0;
    const double wk = (k == 0 || k == H - 1) ? 1.0 : 2.0;
    const double b = 1.0 - alpha * alpha;
    for (int n = H - 1; n >= 0; --n) {
        const int r = (int)(((long long)k * n) % fft_len);
        double c = wk / (double)fft_len * cospi(2.0 * (double)r / (double)fft_len);
        if (n == 0 || n == H - 1) c *= 0.5;
        for (int j = 0; j < n_out; ++j) d[j] = g[j];
        g[0] = c + alpha * d[0];
        if (n_out > 1) g[1] = b * d[0] + alpha * d[1];
        for (int j = 2; j < n_out; ++j) g[j] = d[j - 1] + alpha * (d[j] - g[j - 1]);
    }
    for (int j = 0; j < n_out; ++j) wt64[(size_t)k * ld + j] = g[j];
}

__global__ void k_f64_to_f32(const double* __restrict__ a, float* __restrict__ b, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (float)a[i];
}

cudaError_t build_warp_matrix(int fft_len, int n_out, double alpha, float* wt32, double* scratch64, int ld,
                              cudaStream_t st) {
    const int H = fft_len / 2 + 1;
    const int kpad = ((H + MEL_KSLICE - 1) / MEL_KSLICE) * MEL_KSLICE;
    cudaError_t e = cudaMemsetAsync(scratch64, 0, sizeof(double) * (size_t)kpad * ld, st);
    if (e != cudaSuccess) return e;
    k_build_warp<<<(H + 63) / 64, 64, 0, st>>>(fft_len, n_out, alpha, scratch64, ld);
    const size_t n = (size_t)kpad * ld;
    k_f64_to_f32<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scratch64, wt32, n);
    return cudaGetLastError();
}

// ---- voiced-frame compaction ------------------------------------------------------------------
// The phase streams of unvoiced frames are masked to zero (src/magphase.py:2527-2542), so only voiced frames go
// through the real / imag tile products.  vidx[c] = frame of the c-th voiced frame, cidx[f] = its rank (or -1),
// *count = number of voiced frames.  No host round trip.
// CTA b ranks the 1024 frames [1024 b, 1024 b + 1024).  The number of voiced frames before its segment is counted by the
// CTA itself from the flags (at most 128 KB, L2-resident, four flags per load): no inter-CTA hand-off, no second launch,
// nothing to spin on.  (The first version scanned the whole chunk with ONE CTA: 45 us per 116 k frames with the rest of
// the GPU idle, twice per step.)
__device__ __forceinline__ int block_sum_1024(int x, int* warp_sums, int t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((t & 31) == 0) warp_sums[t >> 5] = x;
    __syncthreads();
    int s = warp_sums[t & 31];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();                                                      // warp_sums is reused by the caller
    return s;                                                             // every thread holds the block total
}

__global__ void __launch_bounds__(1024)
k_voiced_compact(const uint8_t* __restrict__ voi, int n, int32_t* __restrict__ vidx, int32_t* __restrict__ cidx,
                 int32_t* __restrict__ count) {
    __shared__ int warp_sums[32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int base = blockIdx.x * 1024;                                   // < n by the launch geometry
    // (1) voiced frames in [0, base)
    int before = 0;
    if ((reinterpret_cast<uintptr_t>(voi) & 3u) == 0) {                    // uniform: four flags per load
        const uint32_t* v4 = reinterpret_cast<const uint32_t*>(voi);
        for (int i = t; i < base / 4; i += 1024) before += __popc(__vcmpne4(__ldg(&v4[i]), 0u)) >> 3;
    } else {
        for (int i = t; i < base; i += 1024) before += voi[i] != 0;
    }
    before = block_sum_1024(before, warp_sums, t);
    // (2) ranks inside the segment: ballot per warp, exclusive scan of the 32 warp counts
    const int f = base + t;
    const bool v = f < n && voi[f] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) warp_sums[w] = __popc(m);
    __syncthreads();
    int incl = warp_sums[lane];                                           // every warp scans the 32 counts itself
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int warp_excl = __shfl_sync(0xffffffffu, incl, w) - __popc(m);  // counts of warps 0..w-1
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int r = before + warp_excl + __popc(m & ((1u << lane) - 1u));
    if (f < n) cidx[f] = v ? r : -1;
    if (v) vidx[r] = f;
    if (blockIdx.x == gridDim.x - 1 && t == 0) *count = before + total;
}

cudaError_t launch_voiced_compact(const uint8_t* voi, int n, int32_t* vidx, int32_t* cidx, int32_t* count, cudaStream_t st) {
    if (n <= 0) return cudaMemsetAsync(count, 0, sizeof(int32_t), st);
    k_voiced_compact<<<(unsigned)((n + 1023) / 1024), 1024, 0, st>>>(voi, n, vidx, cidx, count);
    return cudaGetLastError();
}

// ---- tile product -----------------------------------------------------------------------------
// log periodogram of one feature value as SPTK sees it (float32 input file):
//   in_type 3 (|X|):   log(x^2 + 1e-8)            in_type 2 (ln|X|):  log(exp(2x) + 1e-8)
// Fast intrinsics: their error (<= ~1e-6 absolute on the log) is far below the float32 noise of the data and is
// attenuated by the warp matrix (row norms ~1e-2).
__device__ __forceinline__ float log_periodogram(float x, bool is_mag) {
    const float t = is_mag ? x * x : __expf(2.0f * x);
    return __logf(t + 1.0e-8f);
}

constexpr int GEMM_FT = 128;          // frames per CTA tile
constexpr int GEMM_CT = 64;           // coefficients per CTA tile
constexpr int GEMM_LDL = GEMM_FT + 4; // pitch of the transposed log-periodogram tile (floats)
constexpr int GEMM_KS = 64;           // bins per shared-memory stage

// grid: (frame tiles, K slices * coefficient tiles, 3 streams).  The K slices cover bins 0 .. H-2 (H-1 = N/2 is a
// multiple of MEL_KSLICE); the Nyquist bin is added by k_mel_finish.  PRE: rows already hold log periodograms.
template <typename TF, bool PRE, bool LERP>
__global__ void __launch_bounds__(128, 4)
k_mel_gemm(const TF* __restrict__ mag, const TF* __restrict__ real, const TF* __restrict__ imag, int64_t nfrm, int H,
           const float* __restrict__ wt_mag, int ld_mag, const float* __restrict__ wt_ph, int ld_ph,
           float* __restrict__ partial, int n_slices, int ncp_max, const int32_t* __restrict__ vidx,
           const int32_t* __restrict__ vcount, const int32_t* __restrict__ lr0, const int32_t* __restrict__ lr1,
           const float* __restrict__ lw) {
    extern __shared__ __align__(16) float smem_f[];
    float* Ls = smem_f;                                  // [GEMM_KS][GEMM_LDL]
    float* Bs = smem_f + GEMM_KS * GEMM_LDL;             // [GEMM_KS][GEMM_CT]
    __shared__ int rowmap[GEMM_FT];                      // tile row -> frame (-1: none)
    __shared__ int rowsrc[LERP ? 2 * GEMM_FT : 1];       // LERP: the two source rows of every tile row
    __shared__ float roww[LERP ? GEMM_FT : 1];           //       and the interpolation weight
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stream = blockIdx.z;
    const int slice = blockIdx.y % n_slices, ctile = blockIdx.y / n_slices;
    const TF* __restrict__ src = stream == 0 ? mag : (stream == 1 ? real : imag);
    const float* __restrict__ wt = stream == 0 ? wt_mag : wt_ph;
    const int ld = stream == 0 ? ld_mag : ld_ph;
    if (ctile * GEMM_CT >= ld) return;
    const int64_t f0 = (int64_t)blockIdx.x * GEMM_FT;
    const int k0 = slice * MEL_KSLICE;
    // rows of this tile: consecutive frames for the magnitude stream, the f0-th.. voiced frames for the phase streams
    const int64_t nrows = (stream == 0 || !vidx) ? nfrm : (int64_t)*vcount;
    if (f0 >= nrows) return;
    if (tid < GEMM_FT) {
        const int64_t r = f0 + tid;
        const int fo = r < nrows ? ((stream == 0 || !vidx) ? (int)r : vidx[r]) : -1;
        rowmap[tid] = fo;
        if (LERP) {
            rowsrc[2 * tid] = fo >= 0 ? lr0[fo] : 0;
            rowsrc[2 * tid + 1] = fo >= 0 ? lr1[fo] : 0;
            roww[tid] = fo >= 0 ? lw[fo] : 0.0f;
        }
    }
    __syncthreads();

    const int tf = tid >> 3, tc = tid & 7;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    // a K-slice is processed in MEL_KSLICE / GEMM_KS stages through one small smem tile (4 CTAs per SM)
#pragma unroll 1
    for (int ks = 0; ks < MEL_KSLICE; ks += GEMM_KS) {
        if (ks) __syncthreads();
        // ---- stage the W^T rows k0+ks .. +GEMM_KS-1, columns ctile*64..+63 (zero padded on the host side) ----
        for (int i = tid; i < GEMM_KS * (GEMM_CT / 4); i += 128) {
            const int kk = i / (GEMM_CT / 4), c4 = i % (GEMM_CT / 4);
            reinterpret_cast<float4*>(Bs)[i] =
                __ldg(reinterpret_cast<const float4*>(wt + (size_t)(k0 + ks + kk) * ld + ctile * GEMM_CT) + c4);
        }
        // ---- stage the log-periodogram tile transposed: Ls[kk][f]; a warp takes 4 frames x 32 bins per step;
        // 32 loads are issued before their first use (two DRAM latencies per stage, overlapped across 4 CTAs/SM) ----
#pragma unroll 1
        for (int gh = 0; gh < 8; gh += 4) {
            float raw[4][2][4];                          // [frame group][bin chunk][frame in group]
            const TF* __restrict__ p0 = src + k0 + ks + lane;
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int tr = (gh + g) * 16 + warp * 4 + r;
                    const int fr = rowmap[tr];
                    if (LERP) {      // interp1d(kind='linear') between two source rows BEFORE the log (src/magphase.py:2967-2980)
                        const int s0 = rowsrc[2 * tr], s1 = rowsrc[2 * tr + 1];
                        const float w = roww[tr];
#pragma unroll
                        for (int c = 0; c < GEMM_KS / 32; ++c) {
                            const float a0 = fr >= 0 ? (float)__ldcs(p0 + s0 * (int64_t)H + c * 32) : 0.0f;
                            const float a1 = fr >= 0 ? (float)__ldcs(p0 + s1 * (int64_t)H + c * 32) : 0.0f;
                            raw[g][c][r] = fmaf(w, a1 - a0, a0);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < GEMM_KS / 32; ++c)
                            raw[g][c][r] = fr >= 0 ? (float)__ldcs(p0 + fr * (int64_t)H + c * 32) : 0.0f;
                    }
                }
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int c = 0; c < GEMM_KS / 32; ++c) {
                    float4 v;
                    float* pv = &v.x;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const bool ok = rowmap[(gh + g) * 16 + warp * 4 + r] >= 0;
                        pv[r] = PRE ? raw[g][c][r] : (ok ? log_periodogram(raw[g][c][r], stream == 0) : 0.0f);
                    }
                    *reinterpret_cast<float4*>(Ls + (c * 32 + lane) * GEMM_LDL + (gh + g) * 16 + warp * 4) = v;
                }
        }
        __syncthreads();

        // ---- 8 x 8 outputs per thread: frames {tf*4.., 64+tf*4..}, coefficients {tc*4.., 32+tc*4..} ----
        const float* pl = Ls + tf * 4;
        const float* pb = Bs + tc * 4;
#pragma unroll 4
        for (int kk = 0; kk < GEMM_KS; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(pl + kk * GEMM_LDL);
            const float4 a1 = *reinterpret_cast<const float4*>(pl + kk * GEMM_LDL + 64);
            const float4 b0 = *reinterpret_cast<const float4*>(pb + kk * GEMM_CT);
            const float4 b1 = *reinterpret_cast<const float4*>(pb + kk * GEMM_CT + 32);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    // ---- partial[stream][row][slice][ncp_max]  (row = frame, or voiced rank for the phase streams) ----
    float* out = partial + (size_t)stream * (size_t)nfrm * n_slices * ncp_max + (size_t)slice * ncp_max + ctile * GEMM_CT;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t f = f0 + (i < 4 ? tf * 4 + i : 64 + tf * 4 + (i - 4));
        if (f >= nrows) continue;
        float* po = out + (size_t)f * n_slices * ncp_max;
        *reinterpret_cast<float4*>(po + tc * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(po + 32 + tc * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
}

// ---- finish -------------------------------------------------------------------------------------
// one warp per (frame, stream): mc[j] = float32(sum over K slices + Nyquist-bin term), out[o] = sum_j mc[j] cos_tab[j][o]
template <typename TF, typename TO, bool PRE, bool LERP>
__global__ void __launch_bounds__(128)
k_mel_finish(const float* __restrict__ partial, int n_slices, int ncp_max, int64_t nfrm,
             const TF* __restrict__ mag, const TF* __restrict__ real, const TF* __restrict__ imag, int H,
             const float* __restrict__ wt_mag, int ld_mag, const float* __restrict__ wt_ph, int ld_ph,
             const double* __restrict__ cos_mag, int n_mag, const double* __restrict__ cos_ph, int n_ph, int phase_dim,
             const uint8_t* __restrict__ voi, const int32_t* __restrict__ cidx, const int32_t* __restrict__ lr0,
             const int32_t* __restrict__ lr1, const float* __restrict__ lw, int raw_mc, TO* __restrict__ out_mag,
             TO* __restrict__ out_real, TO* __restrict__ out_imag) {
    __shared__ double mc[4][MEL_MAX_COEFFS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * 4 + warp;
    const int stream = blockIdx.y;
    if (f >= nfrm) return;
    const int n_in = stream == 0 ? n_mag : n_ph;
    const int n_out = raw_mc ? n_in : (stream == 0 ? n_mag : phase_dim);
    const double* __restrict__ ct = stream == 0 ? cos_mag : cos_ph;      // [n_in][n_out]
    const TF* __restrict__ src = stream == 0 ? mag : (stream == 1 ? real : imag);
    const float* __restrict__ wt = stream == 0 ? wt_mag : wt_ph;
    const int ld = stream == 0 ? ld_mag : ld_ph;
    TO* __restrict__ dst = (stream == 0 ? out_mag : (stream == 1 ? out_real : out_imag)) + f * (int64_t)n_out;
    const bool voiced = voi[f] != 0;
    int64_t row = f;                                                       // row of this frame in `partial`
    if (stream != 0 && cidx && !raw_mc) {
        if (!voiced) {                                                     // masked anyway (src/magphase.py:2527-2528)
            for (int o = lane; o < n_out; o += 32) dst[o] = (TO)0;
            return;
        }
        row = cidx[f];
    }
    float xl;
    if (LERP) {
        const float a0 = (float)src[lr0[f] * (int64_t)H + (H - 1)], a1 = (float)src[lr1[f] * (int64_t)H + (H - 1)];
        xl = fmaf(lw[f], a1 - a0, a0);
    } else {
        xl = (float)src[f * (int64_t)H + (H - 1)];
    }
    const float last = PRE ? xl : log_periodogram(xl, stream == 0);       // Nyquist bin, not covered by the K slices
    const float* __restrict__ pp = partial + ((size_t)stream * (size_t)nfrm + (size_t)row) * n_slices * ncp_max;
    // all K-slice partials of a coefficient are fetched before the first add (up to 16 independent loads in flight per
    // lane: this kernel is pure DRAM latency); the sum itself runs in slice order, so it is reproducible
    constexpr int MAX_SLICES = 4096 / 2 / MEL_KSLICE;
    for (int j = lane; j < n_in; j += 32) {
        float v[MAX_SLICES];
#pragma unroll
        for (int sl = 0; sl < MAX_SLICES; ++sl) v[sl] = sl < n_slices ? __ldcs(pp + sl * ncp_max + j) : 0.0f;
        double s = (double)last * (double)__ldg(wt + (size_t)(H - 1) * ld + j);
#pragma unroll
        for (int sl = 0; sl < MAX_SLICES; ++sl)
            if (sl < n_slices) s += (double)v[sl];
        mc[warp][j] = (double)(float)s;                                    // SPTK writes float32 (src/libaudio.py:593)
    }
    __syncwarp();
    if (raw_mc) {                                                          // la.sp_to_mcep: the mel cepstrum itself
        for (int o = lane; o < n_out; o += 32) dst[o] = (TO)mc[warp][o];
        return;
    }
    for (int o = lane; o < n_out; o += 32) {
        double s = 0.0;
        for (int j = 0; j < n_in; ++j) s = fma(mc[warp][j], ct[j * n_out + o], s);
        if (stream == 0) {
            // m_mag_mel = exp(s); la.log(m_mag_mel) = s up to 1 ulp, except the protected -inf (src/libaudio.py:241-248)
            if (s < -745.0) s = -1.0e10;
        } else {
            s = voiced ? fmin(fmax(s, -1.0), 1.0) : 0.0;                   // mask, clip (src/magphase.py:2527-2542)
        }
        dst[o] = (TO)s;
    }
}

template <typename TF, bool PRE, bool LERP>
static cudaError_t launch_gemm_t(const MelArgs& a, cudaStream_t st) {
    const int H = a.fft_len / 2 + 1;
    const int n_slices = (H - 1) / MEL_KSLICE;
    const int ctiles = (a.ncp_max + GEMM_CT - 1) / GEMM_CT;
    const size_t smem = sizeof(float) * (GEMM_KS * GEMM_LDL + GEMM_KS * GEMM_CT);
    dim3 grid((unsigned)((a.nfrm + GEMM_FT - 1) / GEMM_FT), (unsigned)(n_slices * ctiles), 3);
    auto kern = k_mel_gemm<TF, PRE, LERP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 128, smem, st>>>((const TF*)a.mag, (const TF*)a.real, (const TF*)a.imag, a.nfrm, H, a.wt_mag, a.ld_mag,
                                  a.wt_ph, a.ld_ph, a.partial, n_slices, a.ncp_max, a.vidx, a.vcount, a.lerp_r0, a.lerp_r1,
                                  a.lerp_w);
    return cudaGetLastError();
}

cudaError_t launch_mel_gemm(const MelArgs& a, cudaStream_t st) {
    if (a.lerp_r0) return a.feat_dtype == MPB_F64 ? launch_gemm_t<double, false, true>(a, st) : launch_gemm_t<float, false, true>(a, st);
    if (a.feat_dtype == MPB_F64) return launch_gemm_t<double, false, false>(a, st);
    return a.pre_logp ? launch_gemm_t<float, true, false>(a, st) : launch_gemm_t<float, false, false>(a, st);
}

template <typename TF, typename TO, bool PRE, bool LERP>
static cudaError_t launch_finish_t(const MelArgs& a, cudaStream_t st) {
    const int H = a.fft_len / 2 + 1;
    const int n_slices = (H - 1) / MEL_KSLICE;
    dim3 g2((unsigned)((a.nfrm + 3) / 4), 3);
    k_mel_finish<TF, TO, PRE, LERP><<<g2, 128, 0, st>>>(a.partial, n_slices, a.ncp_max, a.nfrm, (const TF*)a.mag, (const TF*)a.real,
                                                   (const TF*)a.imag, H, a.wt_mag, a.ld_mag, a.wt_ph, a.ld_ph, a.cos_mag,
                                                   a.n_mag, a.cos_ph, a.n_ph, a.phase_dim, a.voi, a.cidx, a.lerp_r0, a.lerp_r1, a.lerp_w, a.raw_mc, (TO*)a.out_mag,
                                                   (TO*)a.out_real, (TO*)a.out_imag);
    return cudaGetLastError();
}

cudaError_t launch_mel_finish(const MelArgs& a, cudaStream_t st) {
    const bool o64 = a.out_dtype == MPB_F64;
    if (a.lerp_r0) {
        if (a.feat_dtype == MPB_F64)
            return o64 ? launch_finish_t<double, double, false, true>(a, st) : launch_finish_t<double, float, false, true>(a, st);
        return o64 ? launch_finish_t<float, double, false, true>(a, st) : launch_finish_t<float, float, false, true>(a, st);
    }
    if (a.feat_dtype == MPB_F64)
        return o64 ? launch_finish_t<double, double, false, false>(a, st) : launch_finish_t<double, float, false, false>(a, st);
    if (a.pre_logp)
        return o64 ? launch_finish_t<float, double, true, false>(a, st) : launch_finish_t<float, float, true, false>(a, st);
    return o64 ? launch_finish_t<float, double, false, false>(a, st) : launch_finish_t<float, float, false, false>(a, st);
}

}  // namespace mpb

// Two small hot-path rows that reuse the FFT engine: the MagPhase post-filter and the minimum-phase builder.
#include "mpb_frame.cuh"

namespace mpb {

// ---- post_filter ------------------------------------------------------------------------------
// Reference: post_filter src/magphase.py:2300-2378.  Per frame and mel bin b:
//   ave[b] = mean(x[c[b]-h[b] .. c[b]+h[b]])   (c, h: host-computed, boundary fill included, :2352-2360)
//   out[b] = (x[b] - ave[b]) * tilt[b] + ave[b],   out[0] = x[0], out[D-1] = x[D-1]      (:2369-2372)
template <typename T>
__global__ void k_post_filter(const T* __restrict__ x, int64_t nfrm, int dim, const int32_t* __restrict__ centre,
                              const int32_t* __restrict__ half, const double* __restrict__ tilt, T* __restrict__ out) {
    extern __shared__ double row[];                       // [frames per CTA][dim]
    const int fpc = blockDim.x / dim;                     // frames per CTA
    const int lf = threadIdx.x / dim, b = threadIdx.x % dim;
    const int64_t f = (int64_t)blockIdx.x * fpc + lf;
    const bool on = lf < fpc && f < nfrm;
    if (on) row[lf * dim + b] = (double)x[f * dim + b];
    __syncthreads();
    if (!on) return;
    const double* r = row + lf * dim;
    const int c = centre[b], h = half[b];
    double s = 0.0;
    for (int i = c - h; i <= c + h; ++i) s += r[i];
    const double ave = s / (double)(2 * h + 1);
    double y = (r[b] - ave) * tilt[b] + ave;
    if (b == 0 || b == dim - 1) y = r[b];
    out[f * dim + b] = (T)y;
}

cudaError_t launch_post_filter(const void* x, int dtype, int64_t nfrm, int dim, const int32_t* centre, const int32_t* half,
                               const double* tilt, void* out, cudaStream_t st) {
    if (nfrm < 1) return cudaSuccess;
    const int fpc = 256 / dim > 0 ? 256 / dim : 1;
    const int threads = fpc * dim;
    const unsigned grid = (unsigned)((nfrm + fpc - 1) / fpc);
    const size_t smem = sizeof(double) * fpc * dim;
    if (dtype == MPB_F64)
        k_post_filter<double><<<grid, threads, smem, st>>>((const double*)x, nfrm, dim, centre, half, tilt, (double*)out);
    else
        k_post_filter<float><<<grid, threads, smem, st>>>((const float*)x, nfrm, dim, centre, half, tilt, (float*)out);
    return cudaGetLastError();
}

// ---- cepstrum -> r[0] (energy of the spectrum a cepstrum describes) ---------------------------------
// SPTK `c2acr -M 0` behind `freqt -A 0` (src/magphase.py:3419-3427, post_filter_merlin): with the all-pass transform folded
// into the cosine table on the host, r0[f] = (1 / L) sum_k w_k exp(2 sum_j c[f][j] G[j][k]) over the K = L/2 + 1 half-circle
// bins (w_k = 2 except at DC and Nyquist).  One CTA per frame; float64; partial sums reduced in a fixed order.
__global__ void __launch_bounds__(256)
k_cep_energy(const double* __restrict__ c, int64_t nfrm, int n, const double* __restrict__ G, int K, int L,
             double* __restrict__ r0) {
    extern __shared__ double ce_s[];                 // [n] cepstrum of the frame, then [8] warp sums
    double* red = ce_s + n;
    const int t = threadIdx.x;
    for (int64_t f = blockIdx.x; f < nfrm; f += gridDim.x) {
        __syncthreads();
        for (int j = t; j < n; j += 256) ce_s[j] = c[f * n + j];
        __syncthreads();
        double acc = 0.0;
        for (int k = t; k < K; k += 256) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s = fma(ce_s[j], __ldg(G + (size_t)j * K + k), s);
            acc += ((k == 0 || k == K - 1) ? 1.0 : 2.0) * exp(2.0 * s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((t & 31) == 0) red[t >> 5] = acc;
        __syncthreads();
        if (t == 0) {
            double s = 0.0;
            for (int i = 0; i < 8; ++i) s += red[i];
            r0[f] = s / (double)L;
        }
    }
}

cudaError_t launch_cep_energy(const double* c, int64_t nfrm, int n, const double* G, int K, int L, double* r0, int num_sms,
                              cudaStream_t st) {
    if (nfrm < 1) return cudaSuccess;
    int64_t grid = (int64_t)num_sms * 8;
    if (grid > nfrm) grid = nfrm;
    k_cep_energy<<<(unsigned)grid, 256, sizeof(double) * (n + 8), st>>>(c, nfrm, n, G, K, L, r0);
    return cudaGetLastError();
}

// ---- minimum phase ------------------------------------------------------------------------------
// Reference: la.build_min_phase_from_mag_spec src/libaudio.py:920-934 (with la.log :241-248):
//   log|X| -> Hermitian extend -> ifft.real (real cepstrum) -> zero n >= H, double n = 1..H-2 -> fft -> half -> exp
// fused in one kernel, float64: one CTA per frame; inverse FFT, lifter in shared memory, forward FFT, complex exp.
// OUT_MODE 0: complex128/complex64 rows [nfrm][H][2];  1: separate real / imag rows of `nb` leading bins (float32),
// which the compressed synthesis kernel consumes for per_phase_type='min_phase'.
template <typename TI, typename TO, int N, int OUT_MODE>
__global__ void __launch_bounds__(FftGeom<double, N>::TPB, 384 / FftGeom<double, N>::TPB)
k_min_phase(const TI* __restrict__ mag, int in_pitch, int64_t nfrm, const double2* __restrict__ tw, TO* __restrict__ out_a,
            TO* __restrict__ out_b, int nb, int out_pitch) {
    using T = double;
    using G = FftGeom<T, N>;
    using T2 = double2;
    constexpr int M = G::M, H = M + 1, TPB = G::TPB;
    constexpr int NJ = (M / 2) / TPB;
    constexpr int STEP = TPB + TPB / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T2* buf = reinterpret_cast<T2*>(smem_raw);
    T2* tw2f = buf + G::BUF_ELEMS;
    T2* tw2i = tw2f + G::TW2_ELEMS;
    const int t = threadIdx.x;
    FftCtx<T> ff, fi;
    fft_setup<T, N, false>(ff, tw2f, tw, t);
    fft_setup<T, N, true>(fi, tw2i, tw, t);
    T2* pk = buf + G::nphys(t);
    T2* pmk = buf + G::nphys(M - t);
    T* bufT = reinterpret_cast<T*>(buf);

    for (int64_t f = blockIdx.x; f < nfrm; f += gridDim.x) {
        const TI* __restrict__ row = mag + f * (int64_t)in_pitch;
        // ---- log magnitude (protected like la.log) packed for the Hermitian inverse transform ----
        auto lg = [&](int k) {
            const double m = (double)row[k];
            double l = log(m);
            if (isinf(l) || isnan(l)) l = -1.0e10;
            return l;
        };
        T2 w = fi.wp;
#pragma unroll 1
        for (int j = 0; j < NJ; ++j) {
            const int k = t + j * TPB;
            const double a = lg(k), b = lg(M - k);          // X[k], conj(X[M-k]): both real
            const T2 o = cmul(mk<T>(a - b, 0.0), w);
            w = cmul(w, fi.wstep);
            pk[j * STEP] = mk<T>((a + b) - o.y, o.x);
            if (k != 0) pmk[-j * STEP] = mk<T>((a + b) + o.y, o.x);
        }
        if (t == 0) buf[G::nphys(M / 2)] = mk<T>(2.0 * lg(M / 2), 0.0);
        __syncthreads();
        T2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = pk[n1 * (G::S1 + G::S1 / 16)];
        __syncthreads();
        fft_m<T, N, true>(v, buf, fi, t);
        // ---- causal lifter on the real cepstrum c[n] = bufT[..] / N ----
        for (int n = t; n < N; n += TPB) {
            T* p = bufT + 2 * G::nphys(n >> 1) + (n & 1);
            const double c = *p / (double)N;
            *p = (n == 0 || n == H - 1) ? c : (n < H - 1 ? 2.0 * c : 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = pk[n1 * (G::S1 + G::S1 / 16)];
        __syncthreads();
        fft_m<T, N, false>(v, buf, ff, t);
        // ---- split, complex exponential, store ----
        w = ff.wp;
#pragma unroll 1
        for (int j = 0; j <= NJ; ++j) {
            const int k = t + j * TPB;
            if (j == NJ && t != 0) break;
            const T2 zk = pk[j * STEP];
            const T2 zm = cconj(k == 0 ? buf[0] : pmk[-j * STEP]);
            const T2 e = mk<T>(0.5 * (zk.x + zm.x), 0.5 * (zk.y + zm.y));
            const T2 d = mk<T>(0.5 * (zk.x - zm.x), 0.5 * (zk.y - zm.y));
            const T2 wo = cmul(mk<T>(d.y, -d.x), w);
            w = cmul(w, ff.wstep);
            T2 x1 = cadd(e, wo), x2 = cconj(csub(e, wo));
            if (k == 0) { x1.y = 0.0; x2.y = 0.0; }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const T2 x = h ? x2 : x1;
                const int kk = h ? (M - k) : k;
                if (h && kk == k) break;
                if (OUT_MODE == 1 && kk >= nb) continue;
                double s, c;
                sincos(x.y, &s, &c);
                const double m = exp(x.x);
                if (OUT_MODE == 0) {
                    out_a[2 * (f * (int64_t)H + kk)] = (TO)(m * c);
                    out_a[2 * (f * (int64_t)H + kk) + 1] = (TO)(m * s);
                } else {
                    out_a[f * (int64_t)out_pitch + kk] = (TO)(m * c);
                    out_b[f * (int64_t)out_pitch + kk] = (TO)(m * s);
                }
            }
        }
        __syncthreads();
    }
}

template <typename TI, typename TO, int N, int OUT_MODE>
static cudaError_t launch_mp_t(const void* mag, int in_pitch, int64_t nfrm, const void* tw, void* out_a, void* out_b, int nb,
                               int out_pitch, int num_sms, cudaStream_t st) {
    using G = FftGeom<double, N>;
    const size_t smem = sizeof(double2) * (G::BUF_ELEMS + 2 * G::TW2_ELEMS);
    auto kern = k_min_phase<TI, TO, N, OUT_MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G::TPB, smem);
    if (e != cudaSuccess) return e;
    int64_t grid = (int64_t)num_sms * (per_sm < 1 ? 1 : per_sm);
    if (grid > nfrm) grid = nfrm;
    if (grid < 1) return cudaSuccess;
    kern<<<(unsigned)grid, G::TPB, smem, st>>>((const TI*)mag, in_pitch ? in_pitch : N / 2 + 1, nfrm, (const double2*)tw,
                                               (TO*)out_a, (TO*)out_b, nb, out_pitch);
    return cudaGetLastError();
}

template <typename TI, typename TO, int OUT_MODE>
static cudaError_t launch_mp_n(int fft_len, const void* mag, int in_pitch, int64_t nfrm, const void* tw, void* out_a,
                               void* out_b, int nb, int out_pitch, int num_sms, cudaStream_t st) {
    switch (fft_len) {
        case 1024: return launch_mp_t<TI, TO, 1024, OUT_MODE>(mag, in_pitch, nfrm, tw, out_a, out_b, nb, out_pitch, num_sms, st);
        case 2048: return launch_mp_t<TI, TO, 2048, OUT_MODE>(mag, in_pitch, nfrm, tw, out_a, out_b, nb, out_pitch, num_sms, st);
        case 4096: return launch_mp_t<TI, TO, 4096, OUT_MODE>(mag, in_pitch, nfrm, tw, out_a, out_b, nb, out_pitch, num_sms, st);
    }
    return cudaErrorInvalidValue;
}

// complex rows out (dtype of input and output: MPB_F32 -> complex64, MPB_F64 -> complex128)
cudaError_t launch_min_phase(int fft_len, const void* mag, int dtype, int64_t nfrm, const void* tw64, void* out_cplx,
                             int num_sms, cudaStream_t st) {
    return dtype == MPB_F64 ? launch_mp_n<double, double, 0>(fft_len, mag, 0, nfrm, tw64, out_cplx, nullptr, 0, 0, num_sms, st)
                            : launch_mp_n<float, float, 0>(fft_len, mag, 0, nfrm, tw64, out_cplx, nullptr, 0, 0, num_sms, st);
}

// float32 rows in, leading nb bins of Re / Im out (float32): feeds k_synthesis_compressed
cudaError_t launch_min_phase_split(int fft_len, const float* mag, int in_pitch, int64_t nfrm, const void* tw64, float* out_re,
                                   float* out_im, int nb, int out_pitch, int num_sms, cudaStream_t st) {
    return launch_mp_n<float, float, 1>(fft_len, mag, in_pitch, nfrm, tw64, out_re, out_im, nb, out_pitch, num_sms, st);
}


// ---- output high-pass (4th-order IIR) as a blocked state-space scan -----------------------------
// Reference: scipy.signal.lfilter(b, a, x) with butter(4, 40 Hz, 'highpass'), src/magphase.py:981-995.
// The recurrence is sequential in time but LINEAR in the state: an utterance is cut into chunks of L samples,
//   pass 1  every chunk runs the filter from a ZERO state and keeps its 4-value end state            (parallel)
//   pass 2  per utterance: s_{c+1} = M^L s_c + e_c  (M: the 4x4 state matrix, M^L from the host)       (tiny)
//   pass 3  every chunk reruns the filter from its true start state s_c and writes y in place        (parallel)
// The filter runs as two cascaded biquads (transposed direct form II each) factored from the reference's own
// (b, a): the direct-form state of a 4th-order filter with four poles at |z| ~ 0.997 is so ill-conditioned that
// M^L is not representable in float64 (its computed eigenvalues come out > 1), the biquad state is fine.
struct SosCoef { double c[2][6]; double ML[16]; };

__device__ __forceinline__ double sos_step(const SosCoef& cf, double* z, double xn) {
    const double y1 = cf.c[0][0] * xn + z[0];
    z[0] = cf.c[0][1] * xn + z[1] - cf.c[0][4] * y1;
    z[1] = cf.c[0][2] * xn - cf.c[0][5] * y1;
    const double y2 = cf.c[1][0] * y1 + z[2];
    z[2] = cf.c[1][1] * y1 + z[3] - cf.c[1][4] * y2;
    z[3] = cf.c[1][2] * y1 - cf.c[1][5] * y2;
    return y2;
}

template <typename T, bool WRITE>
__global__ void k_iir_chunks(T* __restrict__ x, const int64_t* __restrict__ utt_off, const int64_t* __restrict__ chunk_off,
                             int n_utt, int L, SosCoef cf, double* __restrict__ state) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chunk_off[n_utt]) return;
    int u = 0;                                            // utterance of this chunk (binary search)
    {
        int lo = 0, hi = n_utt;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (chunk_off[mid] <= c) lo = mid; else hi = mid; }
        u = lo;
    }
    const int64_t start = utt_off[u] + (c - chunk_off[u]) * (int64_t)L;
    const int64_t end = start + L < utt_off[u + 1] ? start + L : utt_off[u + 1];
    double z[4] = {0, 0, 0, 0};
    if (WRITE) { z[0] = state[4 * c]; z[1] = state[4 * c + 1]; z[2] = state[4 * c + 2]; z[3] = state[4 * c + 3]; }
    for (int64_t n = start; n < end; ++n) {
        const double y = sos_step(cf, z, (double)x[n]);
        if (WRITE) x[n] = (T)y;
    }
    if (!WRITE) { state[4 * c] = z[0]; state[4 * c + 1] = z[1]; state[4 * c + 2] = z[2]; state[4 * c + 3] = z[3]; }
}

// one thread per utterance: turn the zero-state end states into true start states (in place)
__global__ void k_iir_carry(const int64_t* __restrict__ chunk_off, int n_utt, SosCoef cf, double* __restrict__ state) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_utt) return;
    double s[4] = {0, 0, 0, 0};
    for (int64_t c = chunk_off[u]; c < chunk_off[u + 1]; ++c) {
        double e[4], nx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { e[i] = state[4 * c + i]; state[4 * c + i] = s[i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            nx[i] = cf.ML[4 * i] * s[0] + cf.ML[4 * i + 1] * s[1] + cf.ML[4 * i + 2] * s[2] + cf.ML[4 * i + 3] * s[3] + e[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = nx[i];
    }
}

cudaError_t launch_sos2(void* x, int dtype, const int64_t* utt_off, const int64_t* chunk_off, int n_utt, int64_t n_chunks,
                        int L, const double* sos, const double* ML, double* state, cudaStream_t st) {
    if (n_chunks < 1) return cudaSuccess;
    SosCoef cf;
    for (int i = 0; i < 12; ++i) cf.c[i / 6][i % 6] = sos[i];
    for (int i = 0; i < 16; ++i) cf.ML[i] = ML[i];
    const unsigned grid = (unsigned)((n_chunks + 127) / 128);
    if (dtype == MPB_F64) k_iir_chunks<double, false><<<grid, 128, 0, st>>>((double*)x, utt_off, chunk_off, n_utt, L, cf, state);
    else k_iir_chunks<float, false><<<grid, 128, 0, st>>>((float*)x, utt_off, chunk_off, n_utt, L, cf, state);
    k_iir_carry<<<(n_utt + 63) / 64, 64, 0, st>>>(chunk_off, n_utt, cf, state);
    if (dtype == MPB_F64) k_iir_chunks<double, true><<<grid, 128, 0, st>>>((double*)x, utt_off, chunk_off, n_utt, L, cf, state);
    else k_iir_chunks<float, true><<<grid, 128, 0, st>>>((float*)x, utt_off, chunk_off, n_utt, L, cf, state);
    return cudaGetLastError();
}

// ---- ola(): overlap-add of ready-made time-domain frames (src/magphase.py:34-62) ------------------------------------
// One thread per output sample gathers the frames that cover it, in frame order -- the order in which the reference's loop
// adds them into its zero-initialised buffer, so the float64 sums are bit-identical.  Frame i (frmlen columns, centre at
// column frmlen/2) covers the positions [pm[i] - frmlen/2, pm[i] - frmlen/2 + frmlen) of the pitch-mark axis; output sample
// j sits at position j + t0 (magphase.ola_geometry).  pm is non-decreasing (checked on the host): the first covering frame
// is found by bisection.  Reads of one frame are coalesced across the threads of a warp.
__global__ void __launch_bounds__(256)
k_ola_gather(const double* __restrict__ frames, const int32_t* __restrict__ pm, int64_t nfrm, int frmlen, int32_t t0,
             double* __restrict__ out, int64_t n_out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const int64_t pos = j + t0;
    const int half = frmlen / 2;
    const int64_t lo_val = pos + half - frmlen;          // frames with pm[i] > lo_val and pm[i] <= pos + half cover pos
    int64_t a = 0, b = nfrm;
    while (a < b) {                                      // first i with pm[i] > lo_val
        const int64_t m = (a + b) >> 1;
        if ((int64_t)pm[m] > lo_val) b = m; else a = m + 1;
    }
    double acc = 0.0;
    for (int64_t i = a; i < nfrm && (int64_t)pm[i] <= pos + half; ++i)
        acc += frames[i * frmlen + (pos - (int64_t)pm[i] + half)];
    out[j] = acc;
}

cudaError_t launch_ola_gather(const double* frames, const int32_t* pm, int64_t nfrm, int frmlen, int32_t t0, double* out,
                              int64_t n_out, cudaStream_t st) {
    if (n_out < 1) return cudaSuccess;
    k_ola_gather<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(frames, pm, nfrm, frmlen, t0, out, n_out);
    return cudaGetLastError();
}


// ---- compute_lossless_feats() on ready-made spectra (src/magphase.py:457-476) ----------------------------------------
// mag = |X|, real = Re X / |X|, imag = Im X / |X| (0 where |X| == 0), elementwise over n complex128 values.  The analysis
// kernels apply the same normalise() to the spectrum they have just computed; this is the operator for callers that hold
// the output of analysis_with_del_comp_from_pm.
__global__ void __launch_bounds__(256)
k_lossless_feats(const double2* __restrict__ x, int64_t n, double* __restrict__ mag, double* __restrict__ re,
                 double* __restrict__ im) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = x[i];
        double m, a, b;
        normalise(v.x, v.y, m, a, b);
        mag[i] = m; re[i] = a; im[i] = b;
    }
}

cudaError_t launch_lossless_feats(const void* x, int64_t n, double* mag, double* re, double* im, int num_sms, cudaStream_t st) {
    if (n < 1) return cudaSuccess;
    int64_t grid = (n + 255) / 256;
    if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
    k_lossless_feats<<<(unsigned)grid, 256, 0, st>>>((const double2*)x, n, mag, re, im);
    return cudaGetLastError();
}

}  // namespace mpb

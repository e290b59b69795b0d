// Compressed synthesis: un-warped features + windowed noise -> periodic/aperiodic mix -> IFFT -> PSOLA.
//
// Reference: synthesis_from_compressed src/magphase.py:825-997.  Per frame (one CTA walks an OLA run):
//   1. noise frame: windowing() with per-voicing window (:886-892), frm_list_to_matrix + fftshift (:895-896,
//      src/libaudio.py:122-140) == the analysis buffer layout -> forward real FFT
//      (done by k_analysis<noise_logsq> together with the gain statistics; the spectrum is read back here)
//   2. aperiodic = noise_spec / gain(voicing class) * mag [* unvoiced tilt]                    (:905-918)
//      periodic  = mag * (real + j imag)/|.| * voiced tilt                                      (:922-941)
//      mix with sqrt(mask), sqrt(1 - mask); DC and Nyquist become |.|                           (:944-961)
//      -- all folded into three per-bin tables P = sqrt(mask) * tilt_voi, Av = sqrt(1 - mask), Au = tilt_unv
//   3. Hermitian inverse FFT, fftshift (index math), anti-ringing window (:968-973, la.gen_centr_win
//      src/libaudio.py:90-103), ola() (:976, :34-62)
// The noise gains need the mean of (log|N|)^2 over ALL voiced / unvoiced frames of the utterance (:902-903):
// k_analysis<MODE_LOGSQ> produces per-frame sums AND stores the float32 noise spectra (16 KB/frame; HBM has the
// headroom, the issue slots a second noise FFT would take here do not), k_noise_gain reduces the sums per utterance.
#include "mpb_frame.cuh"
#include "mpb_tma.cuh"

namespace mpb {

// one warp per utterance, fixed summation order -> bit-reproducible gains
__global__ void k_noise_gain(const double* __restrict__ logsq, const uint8_t* __restrict__ voi,
                             const int64_t* __restrict__ utt_frm_off, int utt_a, int utt_b, int H,
                             double* __restrict__ inv_gain) {
    const int u = utt_a + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (u >= utt_b) return;
    double sv = 0.0, su = 0.0;
    long long nv = 0, nu = 0;
    for (int64_t f = utt_frm_off[u] + lane; f < utt_frm_off[u + 1]; f += 32) {
        if (voi[f]) { sv += logsq[f]; ++nv; } else { su += logsq[f]; ++nu; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
        su += __shfl_xor_sync(0xffffffffu, su, o);
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        nu += __shfl_xor_sync(0xffffffffu, nu, o);
    }
    if (lane == 0) {
        // gain = sqrt(exp(mean((log|N|)^2)));  an empty class gives NaN like np.mean([]) and is never used
        inv_gain[2 * u + 0] = 1.0 / sqrt(exp(sv / ((double)nv * (double)(H - 2))));
        inv_gain[2 * u + 1] = 1.0 / sqrt(exp(su / ((double)nu * (double)(H - 2))));
    }
}

cudaError_t launch_noise_gain(const SynthCompArgs& a, cudaStream_t st) {
    const int n = a.utt_b - a.utt_a;
    if (n < 1) return cudaSuccess;
    k_noise_gain<<<(n + 3) / 4, 128, 0, st>>>(a.logsq, a.voi, a.utt_frm_off, a.utt_a, a.utt_b, a.fft_len / 2 + 1, a.inv_gain);
    return cudaGetLastError();
}

__device__ __forceinline__ float lerp_row(const float* r0, const float* r1, float w, int k) {
    const float a = r0[k];
    return r1 ? fmaf(w, r1[k] - a, a) : a;
}

// Per-frame scalars, read one frame ahead so that their global-load latency hides under the previous frame's IFFT.
struct FrameDesc {
    int p, A, B, row0, row1;
    float rw;
    bool voiced;
};
__device__ __forceinline__ FrameDesc load_desc(const SynthCompArgs& a, int64_t g) {
    FrameDesc d;
    d.p = a.pm[g]; d.A = a.win_a[g]; d.B = a.win_b[g]; d.row0 = a.row0 ? a.row0[g] : (int)g;
    d.row1 = a.row1 ? a.row1[g] : 0;
    d.rw = a.roww ? a.roww[g] : 0.0f;
    d.voiced = a.voi[g] != 0;
    return d;
}

// One CTA walks an OLA run.  Everything a frame reads from HBM -- its noise spectrum (k_analysis<noise_spec>), its
// un-warped magnitude row and, for voiced frames, the two phase rows -- is staged into shared memory by TMA bulk copies
// issued by one thread as soon as the previous frame's mix has consumed the staging area, i.e. the copies run under
// the previous frame's inverse FFT and overlap-add.
template <typename TO, int N>
__global__ void __launch_bounds__(FftGeom<float, N>::TPB, 384 / FftGeom<float, N>::TPB)
k_synthesis_compressed(const SynthCompArgs a, TO* __restrict__ out) {
    using T = float;
    using G = FftGeom<T, N>;
    using T2 = float2;
    constexpr int M = G::M, TPB = G::TPB, HALF = N / 2;
    constexpr int NJ = (M / 2) / TPB;
    constexpr int STEP = TPB + TPB / 16;
    constexpr int SP = M + 2;                                  // float2 pitch of a stored noise spectrum (16-byte rows)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T2* buf = reinterpret_cast<T2*>(smem_raw);
    T* acc = reinterpret_cast<T*>(buf + G::BUF_ELEMS);
    T2* tw2i = reinterpret_cast<T2*>(acc + N);
    const int H = a.H, HB = a.HB, HP = a.HP, HBP = a.HBP;
    const int rowlen = HP + 2 * HBP;
    T2* sspec = tw2i + G::TW2_ELEMS;                                   // [SP] noise spectrum of the current frame
    float* srow0 = reinterpret_cast<float*>(sspec + SP);               // [mag HP | real HBP | imag HBP]
    float* srow1 = a.row1 ? srow0 + rowlen : nullptr;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(srow0 + rowlen * (a.row1 ? 2 : 1));
    const int t = threadIdx.x;
    const T scale = (T)1 / (T)N;
    FftCtx<T> fi;
    if (t == 0) { mbar_init(mbar, 1); fence_proxy_async(); }
    fft_setup<T, N, true>(fi, tw2i, (const T2*)a.tw, t);               // (ends with a barrier: mbarrier is initialised)
    T2* pk = buf + G::nphys(t);
    T2* pmk = buf + G::nphys(M - t);
    const float* __restrict__ tabP = a.tab;
    const float* __restrict__ tabAv = a.tab + H;
    const float* __restrict__ tabAu = a.tab + 2 * H;
    uint32_t phase = 0;

    // stage frame g (descriptor d): one thread, all copies complete on mbar
    auto stage = [&](int64_t g, const FrameDesc& d) {
        const bool ph = d.voiced && !a.per_linear;
        const uint32_t b_spec = (uint32_t)(SP * sizeof(T2)), b_mag = (uint32_t)(HP * 4), b_ph = (uint32_t)(HBP * 4);
        const uint32_t per_row = b_mag + (ph ? 2 * b_ph : 0);
        mbar_expect_tx(mbar, b_spec + per_row * (srow1 ? 2 : 1));
        tma_load_1d(sspec, a.nspec + g * (int64_t)SP, b_spec, mbar);
        tma_load_1d(srow0, a.m_mag + (int64_t)d.row0 * HP, b_mag, mbar);
        if (ph) {
            tma_load_1d(srow0 + HP, a.m_real + (int64_t)d.row0 * HBP, b_ph, mbar);
            tma_load_1d(srow0 + HP + HBP, a.m_imag + (int64_t)d.row0 * HBP, b_ph, mbar);
        }
        if (srow1) {
            tma_load_1d(srow1, a.m_mag + (int64_t)d.row1 * HP, b_mag, mbar);
            if (ph) {
                tma_load_1d(srow1 + HP, a.m_real + (int64_t)d.row1 * HBP, b_ph, mbar);
                tma_load_1d(srow1 + HP + HBP, a.m_imag + (int64_t)d.row1 * HBP, b_ph, mbar);
            }
        }
    };

    // runs are handed out dynamically (an atomic ticket per run after the CTA's first one): runs differ in length and
    // a static stride left the slowest CTA ~1 run behind the average.  Run boundaries are combined with atomicAdd by
    // exactly two contributors, so the result does not depend on which CTA takes which run.
    __shared__ int s_ticket;
    for (int r = blockIdx.x; r < a.n_runs;) {
        const OlaRun run = a.runs[r];
        const int64_t out_off = a.utt_out_off[run.utt];
        const int64_t out_len = a.utt_out_off[run.utt + 1] - out_off;
        const int t0 = a.utt_t0[run.utt];
        const int own_lo = (run.flags & 1) ? a.pm[run.first - 1] + HALF : INT32_MIN;
        const int own_hi = (run.flags & 2) ? a.pm[run.first + run.count] - HALF : INT32_MAX;
        const float giv = (float)a.inv_gain[2 * run.utt + 0], giu = (float)a.inv_gain[2 * run.utt + 1];

        FrameDesc dn = load_desc(a, run.first);
        if (t == 0) { fence_proxy_async(); stage(run.first, dn); }     // (the previous run ended with a barrier)
        for (int n = t; n < N; n += TPB) acc[n] = (T)0;
        __syncthreads();

        for (int fr = 0; fr < run.count; ++fr) {
            const int64_t g = (int64_t)run.first + fr;
            const FrameDesc d = dn;
            const bool more = fr + 1 < run.count;
            if (more) dn = load_desc(a, g + 1);
            const int p = d.p;
            const bool voiced = d.voiced;

            // ---- 1. wait for this frame's noise spectrum and feature rows ----
            mbar_wait(mbar, phase);
            phase ^= 1u;

            // ---- 2. mix periodic + aperiodic per bin pair (k, M-k), pack for the inverse transform ----
            const float* mag0 = srow0;
            const float* mag1 = srow1;
            const float* re0 = srow0 + HP;
            const float* re1 = srow1 ? srow1 + HP : nullptr;
            const float* im0 = srow0 + HP + HBP;
            const float* im1 = srow1 ? srow1 + HP + HBP : nullptr;
            const float rw = d.rw;
            const float gi = voiced ? giv : giu;
            const float* __restrict__ tabA = voiced ? tabAv : tabAu;
            T2 wi = fi.wp;
#pragma unroll
            for (int j = 0; j <= NJ; ++j) {
                const int k = t + j * TPB;
                if (j == NJ && t != 0) break;                  // k == M/2: thread 0 only
                const int km = M - k;
                T2 n1 = sspec[k];                              // noise spectrum N[k], N[M-k]
                T2 n2 = sspec[km];
                const float magk = lerp_row(mag0, mag1, rw, k), magm = lerp_row(mag0, mag1, rw, km);
                const float sk = gi * magk * __ldg(tabA + k), sm = gi * magm * __ldg(tabA + km);
                T2 xk = mk<T>(n1.x * sk, n1.y * sk);
                T2 xm = mk<T>(n2.x * sm, n2.y * sm);
                if (voiced && k < HB) {                        // periodic part lives below the crossfade band only
                    float ur = 1.0f, ui = 0.0f;                // per_phase_type 'linear': zero phase
                    if (!a.per_linear) { ur = lerp_row(re0, re1, rw, k); ui = lerp_row(im0, im1, rw, k); }
                    const float pw = ur * ur + ui * ui;
                    float s = 0.0f;                            // |u| == 0 -> u / 1 = 0
                    if (pw > 1e-30f && pw < 1e30f) { s = rsqrtf(pw); s = s * fmaf(-0.5f * pw, s * s, 1.5f); }
                    else if (pw > 0.0f) s = 1.0f / hypotf(ur, ui);
                    s *= magk * __ldg(tabP + k);
                    xk.x = fmaf(ur, s, xk.x);
                    xk.y = fmaf(ui, s, xk.y);
                }
                if (k == 0) {                                  // DC and Nyquist: Re = |.|, Im = 0    (:958-961)
                    xk = mk<T>(hypotf(xk.x, xk.y), 0.0f);
                    xm = mk<T>(hypotf(xm.x, xm.y), 0.0f);
                }
                if (j == NJ) {                                 // k == M/2 pairs with itself: Z = 2 conj(X)
                    pk[j * STEP] = mk<T>(2.0f * xk.x, -2.0f * xk.y);
                } else {
                    xm.y = -xm.y;                              // conj(X[M-k])
                    const T2 e2 = cadd(xk, xm);
                    const T2 o2 = cmul(csub(xk, xm), wi);
                    wi = cmul(wi, fi.wstep);
                    pk[j * STEP] = mk<T>(e2.x - o2.y, e2.y + o2.x);
                    if (k != 0) pmk[-j * STEP] = mk<T>(e2.x + o2.y, -e2.y + o2.x);
                }
            }
            __syncthreads();                                   // Z complete; every thread is done with the staging area
            if (t == 0 && more) { fence_proxy_async(); stage(g + 1, dn); }
            T2 v[16];
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) v[n1] = pk[n1 * (G::S1 + G::S1 / 16)];
            __syncthreads();
            fft_m<T, N, true>(v, buf, fi, t);

            // ---- 3. anti-ringing window + overlap-add: sample d in [-A, B] around the pitch mark ----
            const int A = d.A, B = d.B;
            const T* bufT = reinterpret_cast<const T*>(buf);
            const float ia = A > 0 ? 1.0f / (float)A : 0.0f, ib = B > 0 ? 1.0f / (float)B : 0.0f;
            for (int dd = -A + t; dd <= B; dd += TPB) {
                const int n = dd & (N - 1);
                const float w = dd == 0 ? 1.0f : hann_side_f32(dd < 0 ? (float)(-dd) * ia : (float)dd * ib);
                acc[(p + dd) & (N - 1)] += bufT[2 * G::nphys(n >> 1) + (n & 1)] * scale * w;
            }
            __syncthreads();

            const int lo = p - HALF;
            int hi = p + HALF;
            if (more) { const int nx = dn.p - HALF; hi = nx < hi ? nx : hi; }
            ola_flush<T, TO, N, TPB>(acc, lo, hi, own_lo, own_hi, t0, out_len, out + out_off, t);
        }
        if (t == 0) s_ticket = (int)gridDim.x + atomicAdd(a.run_ticket, 1);
        __syncthreads();
        r = s_ticket;
        __syncthreads();                                   // (s_ticket is rewritten at the end of the next run)
    }
}

template <typename TO, int N>
static cudaError_t launch_sc_t(const SynthCompArgs& a, cudaStream_t st) {
    using G = FftGeom<float, N>;
    const size_t smem = sizeof(float2) * (G::BUF_ELEMS + G::TW2_ELEMS + G::M + 2) + sizeof(float) * N +
                        sizeof(float) * (size_t)(a.HP + 2 * a.HBP) * (a.row1 ? 2 : 1) + 16;
    auto kern = k_synthesis_compressed<TO, N>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G::TPB, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)a.num_sms * per_sm;
    if (grid > a.n_runs) grid = a.n_runs;
    if (grid < 1) return cudaSuccess;
    e = cudaMemsetAsync(a.run_ticket, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)grid, G::TPB, smem, st>>>(a, (TO*)a.out);
    return cudaGetLastError();
}

template <typename TO>
static cudaError_t launch_sc_n(const SynthCompArgs& a, cudaStream_t st) {
    switch (a.fft_len) {
        case 1024: return launch_sc_t<TO, 1024>(a, st);
        case 2048: return launch_sc_t<TO, 2048>(a, st);
        case 4096: return launch_sc_t<TO, 4096>(a, st);
    }
    return cudaErrorInvalidValue;
}

// (the caller has zeroed the output samples of the utterances behind a.runs: run boundaries are combined with atomicAdd)
cudaError_t launch_synthesis_compressed(const SynthCompArgs& a, cudaStream_t st) {
    return a.out_dtype == MPB_F64 ? launch_sc_n<double>(a, st) : launch_sc_n<float>(a, st);
}

}  // namespace mpb

// Lossless synthesis kernel: features -> Hermitian inverse FFT -> pitch-synchronous overlap-add.
//
// Work unit = an "OLA run": a stretch of consecutive frames of one utterance (host-chosen, each run
// spans >= fft_len samples).  One CTA of M/16 threads walks its run frame by frame:
//   load mag/real/imag rows (coalesced) -> X = mag * u/|u| -> pack the Hermitian spectrum into the
//   half-size complex spectrum Z -> inverse FFT (mpb_fft.cuh) -> add the N samples (fftshift is index
//   math) into a circular N-sample accumulator in shared memory -> flush the samples no later frame
//   of the run can touch (everything left of pm[i+1] - N/2) to HBM.
// Frames are added in index order like the reference's loop (src/magphase.py:43-56).  Output samples
// that a neighbouring run also touches (within N/2 of the run boundary) are combined with a float
// atomicAdd onto the zero-initialised output: there are at most TWO contributors per sample (runs span
// >= N samples), and a+b == b+a, so the result is bit-reproducible run to run.
// Reference: synthesis_from_lossless src/magphase.py:1759-1776, la.add_hermitian_half
// src/libaudio.py:369-388 (imag of DC/Nyquist dropped), ola() src/magphase.py:34-62.
#include "mpb_frame.cuh"

namespace mpb {

// (real + j imag) / |.| with |.| == 0 -> 1, times mag          src/magphase.py:1761-1766
// |u| == 0 means u == 0, so the quotient is 0 either way: one reciprocal square root does it.
__device__ __forceinline__ float2 feat_to_spec(float mag, float re, float im, float) {
    const float p = re * re + im * im;
    if (p > 1e-30f && p < 1e30f) {
        float r = rsqrtf(p);
        r = r * fmaf(-0.5f * p, r * r, 1.5f);
        const float s = mag * r;
        return make_float2(re * s, im * s);
    }
    if (re == 0.0f && im == 0.0f) return make_float2(0.0f, 0.0f);
    const float s = mag / hypotf(re, im);
    return make_float2(re * s, im * s);
}
__device__ __forceinline__ double2 feat_to_spec(double mag, double re, double im, double) {
    const double p = re * re + im * im;
    if (p > 1e-30 && p < 1e30) {
        double r = (double)rsqrtf((float)p);
        r = r * fma(-0.5 * p, r * r, 1.5);
        r = r * fma(-0.5 * p, r * r, 1.5);
        const double s = mag * r;
        return make_double2(re * s, im * s);
    }
    if (re == 0.0 && im == 0.0) return make_double2(0.0, 0.0);
    const double s = mag / hypot(re, im);
    return make_double2(re * s, im * s);
}

template <typename T, int N> struct SynthCfg {
    // float32: 5 CTAs/SM (102 registers) so that all loads of a frame can be in flight at once (measured best of 4/5/6)
    static constexpr int MINB = (sizeof(T) == 8 ? 384 : 640) / FftGeom<T, N>::TPB;
};

template <typename T, typename TF, typename TO, int N>
__global__ void __launch_bounds__(FftGeom<T, N>::TPB, SynthCfg<T, N>::MINB)
k_synthesis_lossless(const TF* __restrict__ mag, const TF* __restrict__ real, const TF* __restrict__ imag,
                     const int32_t* __restrict__ pm, const int64_t* __restrict__ utt_out_off,
                     const int32_t* __restrict__ utt_t0, const OlaRun* __restrict__ runs, int32_t n_runs,
                     const cx<T>* __restrict__ tw, TO* __restrict__ out, int* __restrict__ run_ticket) {
    using G = FftGeom<T, N>;
    using T2 = cx<T>;
    constexpr int M = G::M, H = M + 1, TPB = G::TPB, HALF = N / 2;
    constexpr int NP = (M / 2) / TPB;            // spectrum pairs (k, M-k) per thread, k in [0, M/2)
    constexpr int NPB = sizeof(T) == 8 ? NP / 2 : NP;   // pairs per load batch (float32: all 6*NP loads of a frame in flight)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T2* buf = reinterpret_cast<T2*>(smem_raw);
    T* acc = reinterpret_cast<T*>(buf + G::BUF_ELEMS);
    const int t = threadIdx.x;
    const T scale = (T)1 / (T)N;
    FftCtx<T> fc;
    fft_setup<T, N, true>(fc, reinterpret_cast<T2*>(acc + N), tw, t);
    constexpr int STEP = TPB + TPB / 16;          // nphys(k + TPB) - nphys(k)
    T2* pk = buf + G::nphys(t);
    T2* pmk = buf + G::nphys(M - t);

    // runs are handed out dynamically after the CTA's first one (see k_synthesis_compressed): boundary samples are
    // combined by exactly two atomicAdd contributors, so the result does not depend on the assignment
    __shared__ int s_ticket;
    for (int r = blockIdx.x; r < n_runs;) {
        const OlaRun run = runs[r];
        const int64_t out_off = utt_out_off[run.utt];
        const int64_t out_len = utt_out_off[run.utt + 1] - out_off;
        const int t0 = utt_t0[run.utt];
        // positions (pm axis) outside [own_lo, own_hi) may also be written by the neighbouring runs
        const int own_lo = (run.flags & 1) ? pm[run.first - 1] + HALF : INT32_MIN;
        const int own_hi = (run.flags & 2) ? pm[run.first + run.count] - HALF : INT32_MAX;

        for (int n = t; n < N; n += TPB) acc[n] = (T)0;
        __syncthreads();

        for (int fr = 0; fr < run.count; ++fr) {
            const int64_t g = (int64_t)run.first + fr;
            const int p = pm[g];
            const int64_t row = g * (int64_t)H;

            // ---- load the half spectrum, build Z (natural padded layout) ----
            // (two batches of NP/2 pairs: 6*NP/2 independent loads in flight per thread)
            const TF* ga = mag + row + t;  const TF* gb = real + row + t;  const TF* gc = imag + row + t;
            const TF* ha = mag + row + M - t;  const TF* hb = real + row + M - t;  const TF* hc = imag + row + M - t;
            T2 w = fc.wp;                                         // conj(W_N^k), k = t + j*TPB
#pragma unroll 1
            for (int jb = 0; jb < NP; jb += NPB) {
                TF fa[NPB][3], fb[NPB][3];
#pragma unroll
                for (int j = 0; j < NPB; ++j) {
                    const int o = (jb + j) * TPB;
                    fa[j][0] = __ldcs(ga + o);   fa[j][1] = __ldcs(gb + o);   fa[j][2] = __ldcs(gc + o);
                    fb[j][0] = __ldcs(ha - o);   fb[j][1] = __ldcs(hb - o);   fb[j][2] = __ldcs(hc - o);
                }
#pragma unroll
                for (int j = 0; j < NPB; ++j) {
                    const int k = t + (jb + j) * TPB;
                    T2 a = feat_to_spec((T)fa[j][0], (T)fa[j][1], (T)fa[j][2], (T)0);
                    T2 b = feat_to_spec((T)fb[j][0], (T)fb[j][1], (T)fb[j][2], (T)0);
                    if (k == 0) { a.y = (T)0; b.y = (T)0; }          // Im X[0] = Im X[M] = 0
                    b.y = -b.y;                                       // B = conj(X[M-k])
                    const T2 e = cadd(a, b);
                    const T2 o = cmul(csub(a, b), w);
                    w = cmul(w, fc.wstep);
                    pk[(jb + j) * STEP] = mk<T>(e.x - o.y, e.y + o.x);                     // Z[k] = E + iO
                    if (k != 0) pmk[-(jb + j) * STEP] = mk<T>(e.x + o.y, -e.y + o.x);       // conj(E) + i conj(O)
                }
            }
            if (t == 0) {                                         // k = M/2 pairs with itself
                T2 a = feat_to_spec((T)mag[row + M / 2], (T)real[row + M / 2], (T)imag[row + M / 2], (T)0);
                buf[G::nphys(M / 2)] = mk<T>((T)2 * a.x, (T)-2 * a.y);
            }
            __syncthreads();
            T2 v[16];
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) v[n1] = pk[n1 * (G::S1 + G::S1 / 16)];
            __syncthreads();
            fft_m<T, N, true>(v, buf, fc, t);

            // ---- overlap-add: sample n of the frame sits at p + (n < N/2 ? n : n - N) ----
            const T* bufT = reinterpret_cast<const T*>(buf);
#pragma unroll 8
            for (int n = t; n < N; n += TPB) {
                const T x = bufT[2 * G::nphys(n >> 1) + (n & 1)] * scale;
                const int pos = p + (n < HALF ? n : n - N);
                acc[pos & (N - 1)] += x;
            }
            __syncthreads();

            // ---- flush what no later frame of this run can touch ----
            const int lo = p - HALF;
            int hi = p + HALF;
            if (fr + 1 < run.count) { const int nx = pm[g + 1] - HALF; hi = nx < hi ? nx : hi; }
            ola_flush<T, TO, N, TPB>(acc, lo, hi, own_lo, own_hi, t0, out_len, out + out_off, t);
            // the next frame's first write to acc happens after two more barriers: no barrier needed here
        }
        if (t == 0) s_ticket = (int)gridDim.x + atomicAdd(run_ticket, 1);
        __syncthreads();
        r = s_ticket;
        __syncthreads();
    }
}

template <typename T, typename TF, typename TO, int N>
static cudaError_t launch_synth_t(const SynthArgs& a, cudaStream_t st) {
    using G = FftGeom<T, N>;
    const size_t smem = sizeof(cx<T>) * (G::BUF_ELEMS + G::TW2_ELEMS) + sizeof(T) * N;
    auto kern = k_synthesis_lossless<T, TF, TO, N>;
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
    kern<<<(unsigned)grid, G::TPB, smem, st>>>((const TF*)a.mag, (const TF*)a.real, (const TF*)a.imag, a.pm,
                                               a.utt_out_off, a.utt_t0, a.runs, a.n_runs, (const cx<T>*)a.tw,
                                               (TO*)a.out, a.run_ticket);
    return cudaGetLastError();
}

template <typename T, typename TF, typename TO>
static cudaError_t launch_synth_n(const SynthArgs& a, cudaStream_t st) {
    switch (a.fft_len) {
        case 1024: return launch_synth_t<T, TF, TO, 1024>(a, st);
        case 2048: return launch_synth_t<T, TF, TO, 2048>(a, st);
        case 4096: return launch_synth_t<T, TF, TO, 4096>(a, st);
    }
    return cudaErrorInvalidValue;
}

template <typename T>
static cudaError_t launch_synth_io(const SynthArgs& a, cudaStream_t st) {
    if (a.feat_dtype == MPB_F32 && a.out_dtype == MPB_F32) return launch_synth_n<T, float, float>(a, st);
    if (a.feat_dtype == MPB_F32 && a.out_dtype == MPB_F64) return launch_synth_n<T, float, double>(a, st);
    if (a.feat_dtype == MPB_F64 && a.out_dtype == MPB_F32) return launch_synth_n<T, double, float>(a, st);
    return launch_synth_n<T, double, double>(a, st);
}

cudaError_t launch_synthesis_lossless(const SynthArgs& a, cudaStream_t st) {
    const size_t esz = a.out_dtype == MPB_F64 ? 8 : 4;
    cudaError_t e = cudaMemsetAsync(a.out, 0, esz * (size_t)a.n_out, st);
    if (e != cudaSuccess) return e;
    return a.compute_dtype == MPB_F64 ? launch_synth_io<double>(a, st) : launch_synth_io<float>(a, st);
}

}  // namespace mpb

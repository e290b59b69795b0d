// In-shared-memory FFT engine for pitch-synchronous frames (sm_100a, no cuFFT).
//
// A frame of N real samples (N = fft_len = 1024 / 2048 / 4096) is transformed through one complex FFT of
// M = N/2 points on z[m] = b[2m] + i b[2m+1] followed by the usual real-FFT split.  The M-point FFT is a
// three-pass Stockham-style decomposition M = 16 x 16 x R3 (R3 = M/256) executed by TPB = M/16 threads:
// every thread owns 16 points in registers, does a radix-16 butterfly per pass (radix-R3 in the last
// pass) and exchanges through ONE padded shared-memory buffer:
//
//   pass 1  thread t         reads  z[n1*S1 + t]              (registers <- caller)      S1 = M/16
//           radix-16 over n1, twiddle W_M^(k1*t), writes A[k1][t]
//   pass 2  thread (k1,m2)   reads  A[k1][m1*R3 + m2], radix-16 over m1,
//           twiddle W_S1^(k2*m2), writes back IN PLACE (same 16 slots, no barrier needed)
//   pass 3  butterfly (k1,k2) reads A[k1][k2*R3 + m2], radix-R3 over m2 -> Z[k1 + 16*k2 + 256*k3]
//
// "A" is stored with one pad element after every R3 so that passes 1, 2 and 3 are all bank-conflict free;
// the natural-order result is stored with one pad after every 16 (phys = k + k/16).
// Twiddles never touch global memory inside the frame loop: pass 1 uses powers of one per-thread register
// value (product tree), pass 2 a 16 x R3 table in shared memory, the real-FFT split a per-thread recurrence.
// They are seeded once per kernel from tw[j] = exp(-2*pi*i*j/N), j in [0, N/2), in the compute precision.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpb {

template <typename T> struct Vec2;
template <> struct Vec2<float>  { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename T> using cx = typename Vec2<T>::type;

template <typename T> __device__ __forceinline__ cx<T> mk(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }
template <typename T2> __device__ __forceinline__ T2 cadd(T2 a, T2 b) { a.x += b.x; a.y += b.y; return a; }
template <typename T2> __device__ __forceinline__ T2 csub(T2 a, T2 b) { a.x -= b.x; a.y -= b.y; return a; }
// float32 complex add / subtract as ONE packed instruction (Blackwell add.f32x2 / sub.f32x2, SASS FADD2): the float32 FFT
// kernels are issue-bound and complex adds are a fifth of their instructions
__device__ __forceinline__ unsigned long long f2_bits(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 bits_f2(unsigned long long r) {
    float2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
template <> __device__ __forceinline__ float2 cadd<float2>(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(r);
}
template <> __device__ __forceinline__ float2 csub<float2>(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(r);
}
template <typename T2> __device__ __forceinline__ T2 cmul(T2 a, T2 b) {
    T2 r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
template <typename T2> __device__ __forceinline__ T2 cconj(T2 a) { a.y = -a.y; return a; }
// multiply by -i (forward) or +i (inverse)
template <bool INV, typename T2> __device__ __forceinline__ T2 rot90(T2 a) {
    T2 r;
    if (INV) { r.x = -a.y; r.y = a.x; } else { r.x = a.y; r.y = -a.x; }
    return r;
}

// tw table: exp(-2 pi i j / N) for j in [0, N/2).  k may be any value in [0, N).
template <typename T, int N, bool INV>
__device__ __forceinline__ cx<T> twiddle(const cx<T>* __restrict__ tw, int k) {
    cx<T> w = __ldg(&tw[k & (N / 2 - 1)]);
    if (k & (N / 2)) { w.x = -w.x; w.y = -w.y; }
    if (INV) w.y = -w.y;
    return w;
}

// ---- register butterflies -------------------------------------------------------------------
template <bool INV, typename T2>
__device__ __forceinline__ void dft2(T2& a0, T2& a1) {
    T2 s = cadd(a0, a1), d = csub(a0, a1);
    a0 = s; a1 = d;
}

template <bool INV, typename T2>
__device__ __forceinline__ void dft4(T2& a0, T2& a1, T2& a2, T2& a3) {
    T2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = rot90<INV>(csub(a1, a3));
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd(d02, d13);
    a3 = csub(d02, d13);
}

// slot -> output index maps of the in-place butterflies below
__host__ __device__ constexpr int perm2(int s) { return s; }
__host__ __device__ constexpr int perm4(int s) { return s; }
__host__ __device__ constexpr int perm8(int s) { return (s >> 1) + 4 * (s & 1); }
__host__ __device__ constexpr int perm16(int s) { return (s >> 2) + 4 * (s & 3); }
template <int R> __host__ __device__ constexpr int perm(int s) {
    return R == 16 ? perm16(s) : (R == 8 ? perm8(s) : s);
}

// 8-point DFT, in place; output k sits in slot s with k = perm8(s)
template <bool INV, typename T, typename T2>
__device__ __forceinline__ void dft8(T2* v) {
    dft4<INV>(v[0], v[2], v[4], v[6]);   // E[k] -> slot 2k
    dft4<INV>(v[1], v[3], v[5], v[7]);   // O[k] -> slot 2k+1
    const T h = (T)0.70710678118654752440;
    // W8^1, W8^2, W8^3 (forward: exp(-i pi k/4))
    T2 o1, o2, o3;
    if (INV) {
        o1 = mk<T>(h * (v[3].x - v[3].y), h * (v[3].x + v[3].y));
        o2 = mk<T>(-v[5].y, v[5].x);
        o3 = mk<T>(-h * (v[7].x + v[7].y), h * (v[7].x - v[7].y));
    } else {
        o1 = mk<T>(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));
        o2 = mk<T>(v[5].y, -v[5].x);
        o3 = mk<T>(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));
    }
    T2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0); v[1] = csub(e0, o0);
    v[2] = cadd(e1, o1); v[3] = csub(e1, o1);
    v[4] = cadd(e2, o2); v[5] = csub(e2, o2);
    v[6] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// 16-point DFT as 4 x 4, in place; input n = 4*n1 + n2 in slot n; output k in slot s, k = perm16(s)
template <bool INV, typename T, typename T2>
__device__ __forceinline__ void dft16(T2* v) {
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // v[4*k1 + n2] *= W16^(n2*k1)
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    auto mulw = [&](T2 a, T wr, T wi) {   // a * (wr - i*wi) forward, a * (wr + i*wi) inverse
        T2 r;
        if (INV) { r.x = a.x * wr - a.y * wi; r.y = a.y * wr + a.x * wi; }
        else     { r.x = a.x * wr + a.y * wi; r.y = a.y * wr - a.x * wi; }
        return r;
    };
    v[5]  = mulw(v[5], c1, s1);                        // W16^1
    v[6]  = mulw(v[6], h, h);                          // W16^2
    v[7]  = mulw(v[7], s1, c1);                        // W16^3
    v[9]  = mulw(v[9], h, h);                          // W16^2
    v[10] = rot90<INV>(v[10]);                         // W16^4 = -i
    v[11] = mulw(v[11], -h, h);                        // W16^6
    v[13] = mulw(v[13], s1, c1);                       // W16^3
    v[14] = mulw(v[14], -h, h);                        // W16^6
    v[15] = mulw(v[15], -c1, -s1);                     // W16^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

// 16-point DFT of an input whose only non-zero points are n = 0 and n = 15 (a pitch-synchronous frame shorter than
// 2/16 of the FFT on either side of its mark, i.e. most speech frames): out[k] = v0 + v15 W16^(15 k) = v0 + v15 conj(W16^k).
// Same slot order as dft16.  ~70 flops instead of ~190.
template <bool INV, typename T, typename T2>
__device__ __forceinline__ void dft16_ends(T2* v) {
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    const T2 a = v[0], b = v[15];
    // conj(W16^k) forward = exp(+2 pi i k / 16); inverse: exp(-2 pi i k / 16)
    const T wr[16] = {(T)1, c1, h, s1, (T)0, -s1, -h, -c1, (T)-1, -c1, -h, -s1, (T)0, s1, h, c1};
    const T wi[16] = {(T)0, s1, h, c1, (T)1, c1, h, s1, (T)0, -s1, -h, -c1, (T)-1, -c1, -h, -s1};
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const int k = perm16(s);
        const T im = INV ? -wi[k] : wi[k];
        T2 r;
        r.x = a.x + (b.x * wr[k] - b.y * im);
        r.y = a.y + (b.x * im + b.y * wr[k]);
        v[s] = r;
    }
}

template <int R, bool INV, typename T, typename T2>
__device__ __forceinline__ void dftR(T2* v) {
    if (R == 16) dft16<INV, T>(v);
    else if (R == 8) dft8<INV, T>(v);
    else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else dft2<INV>(v[0], v[1]);
}

// ---- geometry -------------------------------------------------------------------------------
template <typename T, int N> struct FftGeom {
    static constexpr int M = N / 2;           // complex points
    static constexpr int TPB = M / 16;        // threads per frame
    static constexpr int S1 = M / 16;         // stride of pass 1
    static constexpr int R3 = M / 256;        // radix of pass 3 (2, 4, 8 or 16)
    static constexpr int NB3 = 16 / R3;       // pass-3 butterflies per thread
    // Exchange layout "A": element (k1, x), x in [0, S1), lives at k1*ROW + x + x/R3 (one pad after every R3
    // elements).  When the R3 consecutive elements a pass-2 thread group touches are narrower than 128 bytes, the
    // rows are additionally skewed by R3 elements so that the k1-rows one warp touches in pass 2 tile the 32 banks
    // instead of starting on the same bank.
    static constexpr int SKEW = (R3 * 2 * (int)sizeof(T) < 128) ? R3 : 0;
    static constexpr int ROW = S1 + S1 / R3 + SKEW;
    static constexpr int A_ELEMS = 16 * ROW;
    // natural-order layout: index k lives at k + k/16
    static constexpr int NAT_ELEMS = M + M / 16 + 1;  // (+1: slot for index M)
    static constexpr int BUF_ELEMS = (A_ELEMS > NAT_ELEMS ? A_ELEMS : NAT_ELEMS);
    static constexpr int TW2_ELEMS = 16 * R3;         // pass-2 twiddle table W_S1^(k2*m2), [k2][m2]
    __host__ __device__ static constexpr int nphys(int k) { return k + (k >> 4); }
};

// Per-thread, frame-independent twiddle state (set up once per kernel).
template <typename T> struct FftCtx {
    cx<T> w1;            // W_M^t: pass-1 twiddle base; W_M^(k1 t) are its powers
    cx<T> wp;            // W_N^t: base of the real-FFT split twiddles, advanced by W_N^TPB per step
    cx<T> wstep;         // W_N^TPB
    const cx<T>* tw2;    // shared-memory pass-2 table, pre-offset by m2 = t % R3
};

template <typename T, int N, bool INV>
__device__ __forceinline__ void fft_setup(FftCtx<T>& c, cx<T>* __restrict__ tw2_smem, const cx<T>* __restrict__ tw, int t) {
    using G = FftGeom<T, N>;
    for (int i = t; i < G::TW2_ELEMS; i += G::TPB)
        tw2_smem[i] = twiddle<T, N, INV>(tw, 32 * (i / G::R3) * (i % G::R3));
    c.w1 = twiddle<T, N, INV>(tw, 2 * t);
    c.wp = twiddle<T, N, INV>(tw, t);
    c.wstep = twiddle<T, N, INV>(tw, G::TPB);
    c.tw2 = tw2_smem + (t % G::R3);
    __syncthreads();
}

template <typename T2> __device__ __forceinline__ T2 csqr(T2 a) {
    T2 r;
    r.x = a.x * a.x - a.y * a.y;
    r.y = (a.x + a.x) * a.y;
    return r;
}

// M-point complex FFT of the 16 values per thread in v (v[n1] = z[n1*S1 + t]).
// On return the natural-order spectrum Z[k] is in buf[nphys(k)], k in [0, M), and the CTA is synchronised.
// INV=true computes the un-normalised inverse transform (sum with exp(+...)).
// ends_only (uniform over the CTA): v[1..14] are known to be zero (see dft16_ends).
template <typename T, int N, bool INV>
__device__ __forceinline__ void fft_m(cx<T>* v, cx<T>* __restrict__ buf, const FftCtx<T>& c, int t, bool ends_only = false) {
    using G = FftGeom<T, N>;
    using T2 = cx<T>;
    constexpr int R3 = G::R3, ROW = G::ROW, TPB = G::TPB;

    // pass 1: radix-16 over n1, twiddle by W_M^(k1 t) = product of the binary powers w1, w2, w4, w8 of w1 = W_M^t
    // (four live values instead of a 15-entry table: keeps the float64 kernels under 104 registers)
    if (ends_only) dft16_ends<INV, T>(v);
    else dft16<INV, T>(v);
    {
        const T2 w1 = c.w1, w2 = csqr(w1), w4 = csqr(w2), w8 = csqr(w4);
        T2* pa = buf + t + t / R3;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int k1 = perm16(s);
            T2 x = v[s];
            if (k1 & 1) x = cmul(x, w1);
            if (k1 & 2) x = cmul(x, w2);
            if (k1 & 4) x = cmul(x, w4);
            if (k1 & 8) x = cmul(x, w8);
            pa[k1 * ROW] = x;
        }
    }
    __syncthreads();

    // pass 2 (in place): thread (k1, m2) owns x = m1*R3 + m2  ->  phys = k1*ROW + m2 + m1*(R3+1)
    {
        T2* pa = buf + (t / R3) * ROW + (t % R3);
#pragma unroll
        for (int m1 = 0; m1 < 16; ++m1) v[m1] = pa[m1 * (R3 + 1)];
        dft16<INV, T>(v);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int k2 = perm16(s);
            pa[k2 * (R3 + 1)] = k2 == 0 ? v[s] : cmul(v[s], c.tw2[k2 * R3]);
        }
    }
    __syncthreads();

    // pass 3: NB3 butterflies of radix R3 per thread, b = k1*16 + k2 = t + i*TPB; results kept in v[i*R3 + slot]
    {
        const T2* pa = buf + (t >> 4) * ROW + (t & 15) * (R3 + 1);
#pragma unroll
        for (int i = 0; i < G::NB3; ++i) {
#pragma unroll
            for (int m2 = 0; m2 < R3; ++m2) v[i * R3 + m2] = pa[i * (TPB / 16) * ROW + m2];
            dftR<R3, INV, T>(v + i * R3);
        }
    }
    __syncthreads();
    {
        const int k12 = (t >> 4) + 16 * (t & 15);            // k1 + 16*k2 of butterfly b = t
        T2* pn = buf + G::nphys(k12);
#pragma unroll
        for (int i = 0; i < G::NB3; ++i) {
            // butterfly t + i*TPB has k1 larger by i*TPB/16  ->  natural index larger by the same amount
#pragma unroll
            for (int s = 0; s < R3; ++s) pn[i * (TPB / 16) + 272 * perm<R3>(s)] = v[i * R3 + s];
        }
    }
    __syncthreads();
}

}  // namespace mpb

// EXPERIMENTAL, OFF BY DEFAULT (environment MPB_MEL_TC, read when a plan is created: bit 0 = warp product, bit 1 = un-warp product,
// bit 2 = deeper load pipeline in the warp product, bit 3 = K-slice sums inside the warp-product kernel).  Written at the end of round 1 after the
// GPU budget was spent, to be brought up in round 2: k_mel_gemm_tc has run exactly once on a B200
// (the fused compressed-analysis parity test passed with it, profiles/r1b/mel_tc_first_run.txt) and has never been timed;
// k_mel_unwarp_tc (second half of this file) has only been compiled.
//
// The mel-warp tile product of format_for_modelling (src/magphase.py:2490-2544 -> la.sp_mel_warp src/libaudio.py:643-661)
//     MC[F x 64] = log-periodogram[F x 2048] . W^T[2048 x 64]          (per stream; the Nyquist bin stays in k_mel_finish)
// on the 5th-generation tensor cores: tcgen05.mma kind::tf32, M = 128 frames, N = 64 coefficients, K = 8 bins per
// instruction, accumulators in TMEM.  It writes exactly the K-slice partial sums k_mel_gemm<float, PRE> writes
// (partial[stream][row][slice][64]), so k_mel_finish is shared.
//
// Precision.  TF32 keeps 10 mantissa bits, far too few for the 1e-5 RMS bar, so both operands are split
//     a = a_hi + a_lo,  a_hi = a with the low 13 mantissa bits cleared,  a_lo = (a - a_hi) with its low 13 bits cleared
// (both exactly representable in TF32) and every 8-bin step issues THREE instructions  a_hi.b_hi + a_lo.b_hi + a_hi.b_lo
// ("3xTF32", relative representation error ~2^-21).  The accumulators are float32; whether the hardware rounds or
// truncates when it adds is not documented, so one accumulator only ever sums one 256-bin K slice (32 additions) and
// the slices are added in float64 by k_mel_finish, as on the FMA path.  A NumPy emulation of the pessimistic case
// (truncating adds) gives 1.8e-6 RMS on mag_mel_log and 8e-7 on real_mel for 256-bin slices, 1.5e-5 with a single
// accumulator over all 2048 bins (profiles/r1b/tf32_emulation.txt).
//
// Structure of a CTA (one 128-frame tile of one stream, 1 CTA per SM, 320 threads):
//   warps 0-7  producers: coalesced 4-byte loads of the float32 rows (the rows are 4-byte aligned, pitch H = 2049),
//              hi / lo split, stores into the canonical no-swizzle K-major layout the UMMA descriptors address
//              (8 rows x 16 bytes core matrices; the K stride LBO = 144 bytes makes the stores bank-conflict free);
//   warp  9    one lane streams the pre-split W^T stage (16 KB: hi | lo) with one TMA bulk copy per stage;
//   warp  8    one lane issues the tcgen05.mma instructions, tcgen05.commit releases the stage / signals the epilogue;
//   warps 0-3  epilogue: tcgen05.ld of the n_slices x 64 accumulator columns -> partial sums in HBM.
// A 4-stage ring of full / empty mbarriers connects them.
#include <stdlib.h>

#include "mpb_kernels.h"
#include "mpb_tma.cuh"

namespace mpb {

namespace {

constexpr int TC_M = 128;                      // frames per tile (UMMA M)
constexpr int TC_N = 64;                       // coefficients per tile (UMMA N)
constexpr int TC_KS = 32;                      // bins per pipeline stage (4 instructions of K = 8)
constexpr int TC_ST = 4;                       // pipeline stages
constexpr int TC_LBO_A = 144;                  // bytes between the two 16-byte K chunks of a core-matrix pair (A)
constexpr int TC_SBO_A = 8 * TC_LBO_A;         // bytes between 8-row groups (A): 8 K chunks per stage
constexpr int TC_A_PART = (TC_M / 8) * TC_SBO_A;   // 18,432 bytes (hi or lo)
constexpr int TC_LBO_B = 128;
constexpr int TC_SBO_B = 8 * TC_LBO_B;
constexpr int TC_B_PART = (TC_N / 8) * TC_SBO_B;   // 8,192 bytes (hi or lo)
constexpr int TC_STAGE = 2 * TC_A_PART + 2 * TC_B_PART;   // 53,248 bytes
constexpr int TC_PRODUCERS = 256;
constexpr int TC_THREADS = TC_PRODUCERS + 64;
constexpr int TC_SMEM = TC_ST * TC_STAGE + 1024;
constexpr uint32_t TF32_MASK = 0xFFFFE000u;

// tcgen05 instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32
// (bits 7-9, 10-12 = 2), both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

// shared-memory matrix descriptor (SmemDescriptor of the same header), SWIZZLE_NONE, K-major:
// start address >> 4 at bits 0-13, leading byte offset >> 4 at 16-29 (between the K chunks), stride byte offset >> 4 at
// 32-45 (between 8-row groups), version 1 at bits 46-47.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// W^T [kpad][ld] float32 -> the per-stage operand blocks the bulk copies fetch: [stage][hi | lo][n/8][k/4][n%8][k%4]
__global__ void k_split_warp_tc(const float* __restrict__ wt, int ld, int n_stages, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_stages * TC_KS * TC_N) return;
    const int n = i % TC_N, k = (i / TC_N) % TC_KS, s = i / (TC_N * TC_KS);
    const float w = wt[(size_t)(s * TC_KS + k) * ld + n];
    const float hi = __uint_as_float(__float_as_uint(w) & TF32_MASK);
    const float lo = __uint_as_float(__float_as_uint(w - hi) & TF32_MASK);
    const size_t off = (size_t)s * (2 * TC_B_PART / 4) + (n >> 3) * (TC_SBO_B / 4) + (k >> 2) * (TC_LBO_B / 4) + (n & 7) * 4 + (k & 3);
    out[off] = hi;
    out[off + TC_B_PART / 4] = lo;
}

// DEEP: the producers keep three stages of row loads in flight instead of one (untested variant, MPB_MEL_TC bit 2)
// SUM: the epilogue adds the K-slice accumulators in float64 itself and writes ONE float32 partial per coefficient
// (partial[stream][row][1][64]); k_mel_finish then runs with n_slices = 1 (untested variant, MPB_MEL_TC bit 3)
template <bool DEEP, bool SUM>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_mel_gemm_tc(const float* __restrict__ mag, const float* __restrict__ real, const float* __restrict__ imag, int64_t nfrm,
              int H, const float* __restrict__ btc_mag, const float* __restrict__ btc_ph, float* __restrict__ partial,
              int n_slices, int ncp_max, const int32_t* __restrict__ vidx, const int32_t* __restrict__ vcount) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_ST * TC_STAGE);   // [TC_ST] producers + W^T copy -> MMA
    uint64_t* empty = full + TC_ST;                                          // [TC_ST] MMA (commit) -> producers
    uint64_t* acc_full = empty + TC_ST;                                      // all MMAs done -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    int* rowmap = reinterpret_cast<int*>(tmem_slot + 2);                     // [TC_M] tile row -> frame (-1: none)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stream = blockIdx.y;
    const float* __restrict__ src = stream == 0 ? mag : (stream == 1 ? real : imag);
    const float* __restrict__ btc = stream == 0 ? btc_mag : btc_ph;
    const int64_t nrows = (stream == 0 || !vidx) ? nfrm : (int64_t)*vcount;
    const int64_t f0 = (int64_t)blockIdx.x * TC_M;
    if (f0 >= nrows) return;                                                 // whole CTA, before any barrier / allocation
    const int n_stages = n_slices * (MEL_KSLICE / TC_KS);
    const uint32_t tmem_cols = (uint32_t)(n_slices * TC_N) < 32u ? 32u : (uint32_t)(n_slices * TC_N);   // 128 / 256 / 512

    if (tid < TC_M) {
        const int64_t r = f0 + tid;
        rowmap[tid] = r < nrows ? ((stream == 0 || !vidx) ? (int)r : vidx[r]) : -1;
    }
    if (tid == 0) {
        for (int i = 0; i < TC_ST; ++i) { mbar_init(&full[i], TC_PRODUCERS + 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == 8) {                                                         // one warp allocates (and later frees) TMEM
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ---- producers: warp w owns tile rows 16w .. 16w+15, lane = bin inside the stage ----
        const int m0 = warp * 16;
        if constexpr (DEEP) {
            float buf[4][16];                                                // ring of register buffers: stage s lives in buf[s % 4]
            auto load_stage = [&](int st, float* dst) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int fr = rowmap[m0 + i];
                    dst[i] = fr >= 0 ? __ldcs(src + (int64_t)fr * H + st * TC_KS + lane) : 0.0f;
                }
            };
#pragma unroll
            for (int p = 0; p < 3; ++p)
                if (p < n_stages) load_stage(p, buf[p]);
#pragma unroll 1
            for (int s0 = 0; s0 < n_stages; s0 += 4) {                       // n_stages is a multiple of 8
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int s = s0 + u;
                    if (s + 3 < n_stages) load_stage(s + 3, buf[(u + 3) & 3]);
                    const int slot = s % TC_ST, it = s / TC_ST;
                    if (it > 0) mbar_wait(&empty[slot], (uint32_t)((it - 1) & 1));
                    uint8_t* a_hi = smem + slot * TC_STAGE;
                    uint8_t* a_lo = a_hi + TC_A_PART;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int m = m0 + i;
                        const int off = (m >> 3) * TC_SBO_A + (lane >> 2) * TC_LBO_A + (m & 7) * 16 + (lane & 3) * 4;
                        const float x = buf[u][i];
                        const float hi = __uint_as_float(__float_as_uint(x) & TF32_MASK);
                        const float lo = __uint_as_float(__float_as_uint(x - hi) & TF32_MASK);
                        *reinterpret_cast<float*>(a_hi + off) = hi;
                        *reinterpret_cast<float*>(a_lo + off) = lo;
                    }
                    fence_proxy_async();
                    mbar_arrive(&full[slot]);
                }
            }
        } else {
        float cur[16], nxt[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int fr = rowmap[m0 + i];
            cur[i] = fr >= 0 ? __ldcs(src + (int64_t)fr * H + lane) : 0.0f;
        }
#pragma unroll 1
        for (int s = 0; s < n_stages; ++s) {
            if (s + 1 < n_stages) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int fr = rowmap[m0 + i];
                    nxt[i] = fr >= 0 ? __ldcs(src + (int64_t)fr * H + (s + 1) * TC_KS + lane) : 0.0f;
                }
            }
            const int slot = s % TC_ST, it = s / TC_ST;
            if (it > 0) mbar_wait(&empty[slot], (uint32_t)((it - 1) & 1));   // the MMAs of the slot's previous use are done
            uint8_t* a_hi = smem + slot * TC_STAGE;
            uint8_t* a_lo = a_hi + TC_A_PART;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int m = m0 + i;
                const int off = (m >> 3) * TC_SBO_A + (lane >> 2) * TC_LBO_A + (m & 7) * 16 + (lane & 3) * 4;
                const float x = cur[i];
                const float hi = __uint_as_float(__float_as_uint(x) & TF32_MASK);
                const float lo = __uint_as_float(__float_as_uint(x - hi) & TF32_MASK);
                *reinterpret_cast<float*>(a_hi + off) = hi;
                *reinterpret_cast<float*>(a_lo + off) = lo;
            }
            fence_proxy_async();                                             // generic-proxy stores -> visible to the MMA's async proxy
            mbar_arrive(&full[slot]);
#pragma unroll
            for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
        }
        }
        // ---- epilogue: warps 0-3 read the TMEM lanes 32w .. 32w+31 (= tile rows) ----
        if (warp < 4) {
            mbar_wait(acc_full, 0u);
            tc_fence_after();
            const int r = warp * 32 + lane;
            const bool valid = rowmap[r] >= 0;
            if constexpr (SUM) {
                float* po = partial + ((size_t)stream * (size_t)nfrm + (size_t)(f0 + r)) * ncp_max;
#pragma unroll 1
                for (int c = 0; c < TC_N / 16; ++c) {
                    double acc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = 0.0;
#pragma unroll 1
                    for (int sl = 0; sl < n_slices; ++sl) {                  // slice order, like k_mel_finish
                        float v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sl * TC_N + c * 16), v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] += (double)v[j];
                    }
                    if (valid) {
                        float4* q = reinterpret_cast<float4*>(po + c * 16);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            q[j] = make_float4((float)acc[4 * j], (float)acc[4 * j + 1], (float)acc[4 * j + 2], (float)acc[4 * j + 3]);
                    }
                }
            } else {
            float* po = partial + (size_t)stream * (size_t)nfrm * n_slices * ncp_max + (size_t)(f0 + r) * n_slices * ncp_max;
#pragma unroll 1
            for (int sl = 0; sl < n_slices; ++sl) {
#pragma unroll 1
                for (int c = 0; c < TC_N / 16; ++c) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sl * TC_N + c * 16), v);
                    if (valid) {
                        float4* q = reinterpret_cast<float4*>(po + (size_t)sl * ncp_max + c * 16);
#pragma unroll
                        for (int j = 0; j < 4; ++j) q[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
            }
            tc_fence_before();
        }
    } else if (warp == 8) {
        // ---- MMA issuer: one lane ----
        if (lane == 0) {
#pragma unroll 1
            for (int s = 0; s < n_stages; ++s) {
                const int slot = s % TC_ST, it = s / TC_ST;
                mbar_wait(&full[slot], (uint32_t)(it & 1));
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + slot * TC_STAGE), a_lo = a_hi + TC_A_PART;
                const uint32_t b_hi = a_hi + 2 * TC_A_PART, b_lo = b_hi + TC_B_PART;
                const uint32_t d = tmem_base + (uint32_t)((s / (MEL_KSLICE / TC_KS)) * TC_N);
                const bool first_of_slice = (s % (MEL_KSLICE / TC_KS)) == 0;
#pragma unroll
                for (int j = 0; j < TC_KS / 8; ++j) {
                    const uint64_t da_hi = smem_desc(a_hi + j * 2 * TC_LBO_A, TC_LBO_A, TC_SBO_A);
                    const uint64_t da_lo = smem_desc(a_lo + j * 2 * TC_LBO_A, TC_LBO_A, TC_SBO_A);
                    const uint64_t db_hi = smem_desc(b_hi + j * 2 * TC_LBO_B, TC_LBO_B, TC_SBO_B);
                    const uint64_t db_lo = smem_desc(b_lo + j * 2 * TC_LBO_B, TC_LBO_B, TC_SBO_B);
                    umma_tf32(d, da_hi, db_hi, (first_of_slice && j == 0) ? 0u : 1u);
                    umma_tf32(d, da_lo, db_hi, 1u);
                    umma_tf32(d, da_hi, db_lo, 1u);
                }
                umma_commit(&empty[slot]);                                   // arrives when the MMAs above have read the stage
            }
            umma_commit(acc_full);
        }
        __syncwarp();
    } else {
        // ---- W^T loader: one lane, one 16 KB bulk copy (hi | lo) per stage ----
        if (lane == 0) {
#pragma unroll 1
            for (int s = 0; s < n_stages; ++s) {
                const int slot = s % TC_ST, it = s / TC_ST;
                if (it > 0) mbar_wait(&empty[slot], (uint32_t)((it - 1) & 1));
                mbar_expect_tx(&full[slot], 2 * TC_B_PART);
                tma_load_1d(smem + slot * TC_STAGE + 2 * TC_A_PART, btc + (size_t)s * (2 * TC_B_PART / 4), 2 * TC_B_PART, &full[slot]);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}


// ================================================================================================================
// Un-warp product (synthesis side), same scheme.  la.sp_mel_unwarp src/libaudio.py:667-684 as one matrix per stream:
//     Y[f][b] = sum_c X[f][c] . U[c][b],   K = mag_dim (60) or phase_dim (45), b < np (2052 / 512 pitched bins)
// computed TRANSPOSED so that the epilogue writes coalesced rows:  D[bin][frame] = U^T[128 bins x K] . X^T[K x 64 frames]
// (UMMA M = 128 bins = TMEM lanes, N = 64 frames = TMEM columns): a warp's 32 lanes hold 32 consecutive bins of one frame.
// A CTA owns one 128-bin tile of one stream for its lifetime (U^T tile, pre-split hi | lo, one bulk copy) and walks over
// its share of the 64-frame tiles exactly like k_mel_unwarp (same tile flags, same grid split): raw feature tile by TMA
// bulk copy -> hi / lo split into the canonical K-major layout -> 3 x K/8 tcgen05.mma -> tcgen05.ld -> exp -> stores.
// The steps of a tile are serialised inside the CTA; two (magnitude) or three (phase) CTAs per SM overlap each other.
constexpr int UT_FT = 64;                       // frames per tile (UMMA N)
constexpr int UT_BT = 128;                      // bins per CTA (UMMA M)
constexpr int UT_THREADS = 128;
constexpr int UT_LBO = 128;                     // both operands: K chunks of a core-matrix row group are contiguous

__host__ __device__ inline int ut_kpad(int K) { return (K + 7) & ~7; }

// U [K][np] float32 (zero padded rows) -> per 128-bin tile: [hi | lo][m/8][k/4][m%8][k%4], K padded to a multiple of 8
__global__ void k_split_unwarp_tc(const float* __restrict__ U, int K, int np, int kpad, int n_tiles, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles * UT_BT * kpad) return;
    const int k = i % kpad, m = (i / kpad) % UT_BT, t = i / (kpad * UT_BT);
    const int b = t * UT_BT + m;
    const float u = (k < K && b < np) ? U[(size_t)k * np + b] : 0.0f;
    const float hi = __uint_as_float(__float_as_uint(u) & TF32_MASK);
    const float lo = __uint_as_float(__float_as_uint(u - hi) & TF32_MASK);
    const int part = UT_BT * kpad;              // floats per part
    const size_t off = (size_t)t * 2 * part + (m >> 3) * (kpad / 4) * 32 + (k >> 2) * 32 + (m & 7) * 4 + (k & 3);
    out[off] = hi;
    out[off + part] = lo;
}

__global__ void __launch_bounds__(UT_THREADS, 2)
k_mel_unwarp_tc(const float* __restrict__ mag_mel, const float* __restrict__ real_mel, const float* __restrict__ imag_mel,
                const uint8_t* __restrict__ need_ph, const uint8_t* __restrict__ tile_flags, int64_t nfrm, int n_mag, int n_ph,
                const float* __restrict__ utc_mag, const float* __restrict__ utc_ph,
                float* __restrict__ out_mag, float* __restrict__ out_real, float* __restrict__ out_imag,
                int tiles_mag, int tiles_ph, int g_mag, int g_ph, int HP, int HBP) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int stream, btile, part, parts;
    {
        const int b = blockIdx.x, nm = tiles_mag * g_mag;
        if (b < nm) { stream = 0; btile = b / g_mag; part = b % g_mag; parts = g_mag; }
        else {
            const int c = b - nm;
            stream = 1 + c / (tiles_ph * g_ph);
            btile = (c % (tiles_ph * g_ph)) / g_ph; part = c % g_ph; parts = g_ph;
        }
    }
    const int K = stream == 0 ? n_mag : n_ph;
    const int kpad = ut_kpad(K);
    const int np = stream == 0 ? HP : HBP;
    const float* __restrict__ X = stream == 0 ? mag_mel : (stream == 1 ? real_mel : imag_mel);
    const float* __restrict__ UT = (stream == 0 ? utc_mag : utc_ph) + (size_t)btile * 2 * UT_BT * kpad;
    float* __restrict__ Y = stream == 0 ? out_mag : (stream == 1 ? out_real : out_imag);
    const int b0 = btile * UT_BT;
    const int a_part = UT_BT * kpad * 4, b_part = UT_FT * kpad * 4;          // bytes of one hi / lo part
    const int sbo = (kpad / 4) * UT_LBO;                                     // bytes between 8-row groups (both operands)
    uint8_t* As = smem;                                                      // [hi | lo] U^T tile
    uint8_t* Bs = As + 2 * a_part;                                           // [hi | lo] feature tile, canonical layout
    float* raw = reinterpret_cast<float*>(Bs + 2 * b_part);                  // [UT_FT][K] as it sits in HBM (TMA destination)
    uint8_t* tail = reinterpret_cast<uint8_t*>(raw) + ((UT_FT * K * 4 + 15) & ~15);
    uint64_t* bar_a = reinterpret_cast<uint64_t*>(tail);
    uint64_t* bar_raw = bar_a + 1;
    uint64_t* bar_mma = bar_a + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a + 3);
    uint8_t* need_s = reinterpret_cast<uint8_t*>(tmem_slot + 2);             // [UT_FT] frame of the tile is written
    const int64_t n_ft = (nfrm + UT_FT - 1) / UT_FT;

    if (tid == 0) {
        mbar_init(bar_a, 1); mbar_init(bar_raw, 1); mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)UT_FT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 2 * b_part / 16; i += UT_THREADS)                  // the K padding of the feature tile stays zero
        reinterpret_cast<float4*>(Bs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto next_tile = [&](int64_t ft) {
        while (ft < n_ft && stream != 0 && tile_flags[ft] == 0) ft += parts;
        return ft;
    };
    auto tile_rows = [&](int64_t ft) { return (int)(nfrm - ft * UT_FT < UT_FT ? nfrm - ft * UT_FT : UT_FT); };
    auto by_tma = [&](int64_t ft) { return ((tile_rows(ft) * K) & 3) == 0; };
    auto fetch = [&](int64_t ft) {                                           // thread 0 only
        const uint32_t bytes = (uint32_t)(tile_rows(ft) * K * 4);
        mbar_expect_tx(bar_raw, bytes);
        tma_load_1d(raw, X + ft * (int64_t)(UT_FT * K), bytes, bar_raw);
    };
    int64_t ft = next_tile(part);
    if (tid == 0) {
        if (ft < n_ft) {                                                     // the U^T tile: hi | lo are contiguous
            mbar_expect_tx(bar_a, 2 * (uint32_t)a_part);
            tma_load_1d(As, UT, (uint32_t)a_part, bar_a);
            tma_load_1d(As + a_part, UT + a_part / 4, (uint32_t)a_part, bar_a);
        }
        if (ft < n_ft && by_tma(ft)) fetch(ft);
    }
    if (ft < n_ft) mbar_wait(bar_a, 0u);
    uint32_t ph_raw = 0, ph_mma = 0;
    const int bin = b0 + warp * 32 + lane;                                   // this thread's TMEM lane
    const bool bin_ok = bin < np;

    while (ft < n_ft) {
        const int64_t f0 = ft * UT_FT;
        const int rows = tile_rows(ft);
        if (by_tma(ft)) {
            mbar_wait(bar_raw, ph_raw);
            ph_raw ^= 1u;
        } else {
            for (int i = tid; i < rows * K; i += UT_THREADS) raw[i] = X[f0 * K + i];
            __syncthreads();
        }
        if (tid < UT_FT) need_s[tid] = (tid < rows && (stream == 0 || need_ph[f0 + tid] != 0)) ? 1 : 0;
        // hi / lo split: one item = 4 consecutive coefficients of one frame = one 16-byte row of a core matrix
        for (int i = tid; i < UT_FT * (kpad / 4); i += UT_THREADS) {
            const int f = i % UT_FT, k4 = i / UT_FT;                         // 8 consecutive threads: the 8 rows of one core matrix
            float x[4], hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * k4 + j;
                x[j] = (f < rows && c < K) ? raw[f * K + c] : 0.0f;
                hi[j] = __uint_as_float(__float_as_uint(x[j]) & TF32_MASK);
                lo[j] = __uint_as_float(__float_as_uint(x[j] - hi[j]) & TF32_MASK);
            }
            const int off = (f >> 3) * sbo + k4 * UT_LBO + (f & 7) * 16;
            *reinterpret_cast<float4*>(Bs + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(Bs + b_part + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();
        __syncthreads();                                                     // feature tile complete, raw free
        const int64_t ft_next = next_tile(ft + parts);
        if (tid == 0) {
            if (ft_next < n_ft && by_tma(ft_next)) fetch(ft_next);           // lands under the MMAs and the epilogue
            tc_fence_after();
            const uint32_t a_hi = smem_u32(As), a_lo = a_hi + (uint32_t)a_part;
            const uint32_t b_hi = smem_u32(Bs), b_lo = b_hi + (uint32_t)b_part;
            for (int j = 0; j < kpad / 8; ++j) {
                const uint64_t da_hi = smem_desc(a_hi + j * 2 * UT_LBO, UT_LBO, (uint32_t)sbo);
                const uint64_t da_lo = smem_desc(a_lo + j * 2 * UT_LBO, UT_LBO, (uint32_t)sbo);
                const uint64_t db_hi = smem_desc(b_hi + j * 2 * UT_LBO, UT_LBO, (uint32_t)sbo);
                const uint64_t db_lo = smem_desc(b_lo + j * 2 * UT_LBO, UT_LBO, (uint32_t)sbo);
                umma_tf32(tmem_base, da_hi, db_hi, j == 0 ? 0u : 1u);
                umma_tf32(tmem_base, da_lo, db_hi, 1u);
                umma_tf32(tmem_base, da_hi, db_lo, 1u);
            }
            umma_commit(bar_mma);
        }
        __syncwarp();
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1u;
        tc_fence_after();
        // epilogue: lane = bin, column = frame; every store instruction of a warp writes 128 contiguous bytes of one row
#pragma unroll 1
        for (int c = 0; c < UT_FT / 16; ++c) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 16), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (need_s[c * 16 + j] && bin_ok)
                    Y[(f0 + c * 16 + j) * (int64_t)np + bin] = stream == 0 ? __expf(v[j]) : v[j];
            }
        }
        tc_fence_before();
        __syncthreads();                                                     // accumulator and feature tile are free again
        ft = ft_next;
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)UT_FT) : "memory");
    }
}

}  // namespace

size_t mel_tc_operand_bytes(int fft_len) { return (size_t)((fft_len / 2) / TC_KS) * 2 * TC_B_PART; }

// MPB_MEL_TC bit 3: the tensor-core kernel sums the K slices itself (see SUM above)
bool mel_tc_sums_slices() {
    static const bool on = [] { const char* e = getenv("MPB_MEL_TC"); return e && (atoi(e) & 8); }();
    return on;
}

bool mel_tc_usable(const MelArgs& a) {
    return a.wt_tc_mag && a.wt_tc_ph && a.pre_logp && a.feat_dtype == MPB_F32 && !a.lerp_r0 && a.ncp_max == TC_N &&
           a.ld_mag == TC_N && a.ld_ph == TC_N;
}

cudaError_t build_warp_matrix_tc(int fft_len, const float* wt32, int ld, float* out, cudaStream_t st) {
    const int n_stages = (fft_len / 2) / TC_KS;
    const int n = n_stages * TC_KS * TC_N;
    k_split_warp_tc<<<(n + 255) / 256, 256, 0, st>>>(wt32, ld, n_stages, out);
    return cudaGetLastError();
}

cudaError_t launch_mel_gemm_tc(const MelArgs& a, cudaStream_t st) {
    const int H = a.fft_len / 2 + 1;
    const int n_slices = (H - 1) / MEL_KSLICE;
    static const bool deep = [] { const char* e = getenv("MPB_MEL_TC"); return e && (atoi(e) & 4); }();
    const bool sum = a.partial_slices == 1;
    auto kern = deep ? (sum ? k_mel_gemm_tc<true, true> : k_mel_gemm_tc<true, false>)
                     : (sum ? k_mel_gemm_tc<false, true> : k_mel_gemm_tc<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((a.nfrm + TC_M - 1) / TC_M), 3);
    kern<<<grid, TC_THREADS, TC_SMEM, st>>>((const float*)a.mag, (const float*)a.real, (const float*)a.imag, a.nfrm, H,
                                                      a.wt_tc_mag, a.wt_tc_ph, a.partial, n_slices, a.ncp_max, a.vidx, a.vcount);
    return cudaGetLastError();
}

// ---- un-warp product ----
size_t unwarp_tc_operand_bytes(int np, int K) { return (size_t)((np + UT_BT - 1) / UT_BT) * 2 * UT_BT * ut_kpad(K) * sizeof(float); }

cudaError_t build_unwarp_matrix_tc(const float* U, int K, int np, float* out, cudaStream_t st) {
    const int n_tiles = (np + UT_BT - 1) / UT_BT, kpad = ut_kpad(K);
    const int n = n_tiles * UT_BT * kpad;
    k_split_unwarp_tc<<<(n + 255) / 256, 256, 0, st>>>(U, K, np, kpad, n_tiles, out);
    return cudaGetLastError();
}

bool unwarp_tc_usable(const UnwarpArgs& a) { return a.u_tc_mag && a.u_tc_ph && a.n_mag <= 64 && a.n_ph <= 64; }

// xm / xr / xi: the float32 feature matrices (already narrowed by launch_mel_unwarp), a.flags already filled
cudaError_t launch_mel_unwarp_tc(const UnwarpArgs& a, const float* xm, const float* xr, const float* xi, cudaStream_t st) {
    const int tiles_mag = (a.HP + UT_BT - 1) / UT_BT, tiles_ph = (a.HBP + UT_BT - 1) / UT_BT;
    const int kmax = a.n_mag > a.n_ph ? a.n_mag : a.n_ph, kp = ut_kpad(kmax);
    const size_t smem = (size_t)2 * UT_BT * kp * 4 + (size_t)2 * UT_FT * kp * 4 + (((size_t)UT_FT * kmax * 4 + 15) & ~(size_t)15) + 128;
    cudaError_t e = cudaFuncSetAttribute(k_mel_unwarp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mel_unwarp_tc, UT_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 512 / UT_FT) per_sm = 512 / UT_FT;                          // TMEM columns
    const int64_t n_ft = (a.nfrm + UT_FT - 1) / UT_FT;
    const int slots = a.num_sms * per_sm;
    const double w_mag = (double)tiles_mag * a.n_mag, w_ph = 2.0 * tiles_ph * a.n_ph * 0.95;
    int g_ph = (int)((double)slots * w_ph / (w_mag + w_ph) / (2 * tiles_ph));
    if (g_ph < 1) g_ph = 1;
    int g_mag = (slots - 2 * tiles_ph * g_ph) / tiles_mag;
    if (g_mag < 1) g_mag = 1;
    if (g_mag > n_ft) g_mag = (int)n_ft;
    if (g_ph > n_ft) g_ph = (int)n_ft;
    const unsigned grid = (unsigned)(tiles_mag * g_mag + 2 * tiles_ph * g_ph);
    k_mel_unwarp_tc<<<grid, UT_THREADS, smem, st>>>(xm, xr, xi, a.need_ph, a.flags, a.nfrm, a.n_mag, a.n_ph, a.u_tc_mag, a.u_tc_ph,
                                                    a.out_mag, a.out_real, a.out_imag, tiles_mag, tiles_ph, g_mag, g_ph, a.HP, a.HBP);
    return cudaGetLastError();
}

}  // namespace mpb

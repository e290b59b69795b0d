// EXPERIMENTAL, OFF BY DEFAULT (MPB_MEL_TC=1 selects it; never run on a GPU yet -- written at the end of round 1 after
// the GPU budget was spent, to be brought up in round 2).
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
        float cur[16], nxt[16];
        const int m0 = warp * 16;
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
        // ---- epilogue: warps 0-3 read the TMEM lanes 32w .. 32w+31 (= tile rows) ----
        if (warp < 4) {
            mbar_wait(acc_full, 0u);
            tc_fence_after();
            const int r = warp * 32 + lane;
            const bool valid = rowmap[r] >= 0;
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

}  // namespace

size_t mel_tc_operand_bytes(int fft_len) { return (size_t)((fft_len / 2) / TC_KS) * 2 * TC_B_PART; }

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
    cudaError_t e = cudaFuncSetAttribute(k_mel_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((a.nfrm + TC_M - 1) / TC_M), 3);
    k_mel_gemm_tc<<<grid, TC_THREADS, TC_SMEM, st>>>((const float*)a.mag, (const float*)a.real, (const float*)a.imag, a.nfrm, H,
                                                      a.wt_tc_mag, a.wt_tc_ph, a.partial, n_slices, a.ncp_max, a.vidx, a.vcount);
    return cudaGetLastError();
}

}  // namespace mpb

// Mel-warp tile product of format_for_modelling on the 5th-generation tensor cores: the K-slice sums, the Nyquist term and
// SPTK's float32 rounding happen in the epilogue, only 64 mel cepstra per frame and stream leave the kernel.
//
// Reference: format_for_modelling src/magphase.py:2490-2544 -> la.sp_mel_warp src/libaudio.py:643-661 -> la.sp_to_mcep
// :575-601 (SPTK `mcep -j 0`: linear in the log periodogram, see mpb_mel.cu) -> la.mcep_to_sp_cosmat(alpha=0) :605-631,
// voicing mask and clip src/magphase.py:2527-2542.
//
//     MC[rows x 64] = log-periodogram[rows x K] . W^T[K x 64],   K = fft_len / 2 bins (+ the Nyquist bin in the epilogue)
//     out[rows x n_out] = float32(MC) . cos_tab            (float64), then log floor (magnitude) / clip (phase streams)
//
// Rows: the float32 log periodograms k_analysis<logp> leaves in HBM, pitched to a multiple of 16 bytes so that a 2-D
// tensor map can describe them; the rows of the two phase streams are COMPACTED (voiced frames only, in order).
//
// Precision.  kind::tf32 reads 32-bit containers and uses sign, exponent and 10 mantissa bits, so the raw float32 tile IS
// the high part A_hi of the "3xTF32" split; the low part A_lo = x - trunc_tf32(x) is computed by the converter warps and
// handed to the tensor core through TENSOR MEMORY (tcgen05.mma with A in TMEM), W^T is pre-split on the host side of the
// plan (hi | lo per 32-bin stage).  Measured on the B200: the float32 accumulator of tcgen05.mma TRUNCATES (a 256-bin
// slice = 96 accumulations biased the cepstra by 5e-6 relative, 2e-5 RMS on natural speech), so
//   * the dominant product A_hi.W_hi only ever accumulates 16 instructions (128 bins) in one TMEM accumulator; the
//     slices are drained to registers and summed there with rounding float32 adds;
//   * the two cross products (2^-11 of the main one) have their own accumulator over the whole K range.
//
// Structure (persistent CTAs, one per SM, 640 threads, static round-robin over work items of 256 rows of one stream):
//   (register budgets are re-balanced per warpgroup with setmaxnreg: converters 72, control 40, drainers 144)
//   warp 16     one lane: TMA producer.  Per 32-bin stage two 2-D tiled loads (128 rows x 128 bytes each, 128-byte
//               swizzle = the K-major UMMA layout) and one bulk copy of the W^T stage, 4-stage ring.
//   warps 0-7   converters: own row of the raw tile (swizzled 16-byte chunks, conflict free) -> A_lo -> tcgen05.st.
//   warp 17     one lane: tcgen05.mma issue (24 per stage), tcgen05.commit hands stages / accumulators on.
//   warps 8-15  drainers: tcgen05.ld of the slice accumulators, float32 sums, Nyquist term, float32 rounding -> one row of
//               64 mel cepstra per thread to HBM (256 bytes per frame and stream).
// k_mel_cos (second kernel, 0.1 ms) applies the 60 x 60 cosine matrix in float64, the log floor / clip, and scatters the
// compacted phase rows back to their frames.  (Inside the epilogue of the first kernel the float64 product costs 80k cycles
// per item: broadcast shared-memory reads deliver 4 bytes per wavefront.)
#include "mpb_kernels.h"
#include "mpb_tc.cuh"

namespace mpb {

namespace {

using namespace tc;

constexpr int BM = 128;                      // rows per tile (UMMA M)
constexpr int NT = 2;                        // tiles per work item
constexpr int BN = 64;                       // coefficients (UMMA N)
constexpr int BK = 32;                       // bins per stage = one 128-byte swizzle row
constexpr int NST = 4;                       // ring stages
constexpr int SLICE_ST = 4;                  // stages per main-accumulator slice (128 bins = 16 accumulations)
constexpr int A_TILE = BM * BK * 4;          // 16,384 bytes
constexpr int B_PART = BN * BK * 4;          // 8,192 bytes (hi or lo)
constexpr int STAGE = NT * A_TILE + 2 * B_PART;   // 49,152 bytes
constexpr int LBO_B = 128, SBO_B = 1024;     // W^T stage: canonical no-swizzle K-major core matrices (k_split_warp_stage)
constexpr int CONV_WARPS = 8, DRAIN_WARPS = 8;
constexpr int TMA_WARP = 16, MMA_WARP = 17;
constexpr int THREADS = 20 * 32;            // warps 18, 19 only complete the fifth warpgroup (setmaxnreg is per warpgroup)
constexpr int REGS_CONV = 72, REGS_CTRL = 40, REGS_DRAIN = 144;   // 96 at launch (65,536 / 640); the drainers hold 64 sums + the epilogue
constexpr int SMEM = 1024 + NST * STAGE + 64 * 4 + 256;
constexpr uint32_t TMEM_COLS = 512;
__device__ __forceinline__ uint32_t col_main(int t, int b) { return (uint32_t)((t * 2 + b) * 64); }
__device__ __forceinline__ uint32_t col_cross(int t) { return (uint32_t)(256 + t * 64); }
__device__ __forceinline__ uint32_t col_alo(int t, int b) { return (uint32_t)(384 + (t * 2 + b) * 32); }
constexpr uint32_t IDESC = idesc_tf32(BM, BN);

struct Params {
    const float* lp[3]; int lp_pitch;                 // log-periodogram rows (mag: frame order; real / imag: voiced rank order)
    const float* btc[2];                              // pre-split W^T stages: mag, phase
    const float* wlast[2];                            // W^T row of the Nyquist bin (64 floats): mag, phase
    const int32_t* vcount;                            // number of voiced frames = rows of the phase streams (NULL: nfrm)
    int64_t nfrm; int H; int n_kstages;
    float* mc[3];                                     // out: float32-rounded mel cepstra [rows][64] per stream
};

// W^T [kpad][ld] float32 -> per 32-bin stage [hi | lo][n/8][k/4][n%8][k%4]
__global__ void k_split_warp_stage(const float* __restrict__ wt, int ld, int n_stages, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_stages * BK * BN) return;
    const int n = i % BN, k = (i / BN) % BK, s = i / (BN * BK);
    const float w = wt[(size_t)(s * BK + k) * ld + n];
    const float hi = __uint_as_float(__float_as_uint(w) & TF32_MASK);
    const float lo = __uint_as_float(__float_as_uint(w - hi) & TF32_MASK);
    const size_t off = (size_t)s * (2 * B_PART / 4) + (n >> 3) * (SBO_B / 4) + (k >> 2) * (LBO_B / 4) + (n & 7) * 4 + (k & 3);
    out[off] = hi;
    out[off + B_PART / 4] = lo;
}

__global__ void __launch_bounds__(THREADS, 1)
k_mel_warp_tc(const __grid_constant__ CUtensorMap map_mag, const __grid_constant__ CUtensorMap map_real,
              const __grid_constant__ CUtensorMap map_imag, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stages = smem;
    float* wlast_s = reinterpret_cast<float*>(smem + NST * STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wlast_s + 64);
    uint64_t* full = bars;                    // [NST]  TMA -> converters, MMA
    uint64_t* empty = full + NST;             // [NST]  MMA (commit) -> TMA
    uint64_t* alo_full = empty + NST;         // [2]    converters -> MMA
    uint64_t* alo_empty = alo_full + 2;       // [2]    MMA (commit) -> converters
    uint64_t* slice_full = alo_empty + 2;     // [2]    MMA (commit) -> drainers
    uint64_t* slice_empty = slice_full + 2;   // [2]    drainers -> MMA
    uint64_t* cross_full = slice_empty + 2;   // [1]
    uint64_t* cross_empty = cross_full + 1;   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cross_empty + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (the warp reductions only tell the compiler that these loaded values are warp-uniform: loop counters, descriptors and
    // tensor-memory addresses of the issuing warps then stay in uniform registers)
    const int64_t nv = p.vcount ? (int64_t)__reduce_max_sync(0xffffffffu, (unsigned)*p.vcount) : p.nfrm;
    const int items_mag = (int)((p.nfrm + NT * BM - 1) / (NT * BM));
    const int items_ph = (int)((nv + NT * BM - 1) / (NT * BM));
    const int n_items = items_mag + 2 * items_ph;
    const int nks = p.n_kstages;

    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&alo_full[i], CONV_WARPS); mbar_init(&alo_empty[i], 1);
            mbar_init(&slice_full[i], 1); mbar_init(&slice_empty[i], DRAIN_WARPS);
        }
        mbar_init(cross_full, 1); mbar_init(cross_empty, DRAIN_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == TMA_WARP && lane == 0) { tma_prefetch_desc(&map_mag); tma_prefetch_desc(&map_real); tma_prefetch_desc(&map_imag); }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);

    auto decode = [&](int item, int& stream, int64_t& row0) {
        if (item < items_mag) { stream = 0; row0 = (int64_t)item * (NT * BM); }
        else { const int j = item - items_mag; stream = 1 + j / items_ph; row0 = (int64_t)(j % items_ph) * (NT * BM); }
    };

    if (warp >= TMA_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CTRL));
    }
    if (warp == TMA_WARP) {
        // warp-uniform control flow, one elected lane issues (see elect_one)
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int stream; int64_t row0;
            decode(item, stream, row0);
            const CUtensorMap* map = stream == 0 ? &map_mag : (stream == 1 ? &map_real : &map_imag);
            const float* btc = p.btc[stream == 0 ? 0 : 1];
            for (int ks = 0; ks < nks; ++ks, ++it) {
                const uint32_t s = it % NST, n = it / NST;
                mbar_wait(&empty[s], (n & 1u) ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(&full[s], STAGE);
                    uint8_t* st = stages + s * STAGE;
                    tma_load_2d(st, map, ks * BK, (int)row0, &full[s]);
                    tma_load_2d(st + A_TILE, map, ks * BK, (int)row0 + BM, &full[s]);
                    tma_load_1d(st + NT * A_TILE, btc + (size_t)ks * (2 * B_PART / 4), 2 * B_PART, &full[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp == MMA_WARP) {
        // warp-uniform control flow, one elected lane issues
        uint32_t it = 0, slc = 0, item_n = 0, mb = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
            for (int ks = 0; ks < nks; ++ks, ++it) {
                const uint32_t s = it % NST, n = it / NST, b = it & 1u, nb = it >> 1;
                if (ks % SLICE_ST == 0) {
                    mb = slc & 1u;
                    mbar_wait(&slice_empty[mb], ((slc >> 1) & 1u) ^ 1u);
                }
                if (ks == 0) mbar_wait(cross_empty, (item_n & 1u) ^ 1u);
                mbar_wait(&full[s], n & 1u);
                mbar_wait(&alo_full[b], nb & 1u);
                fence_after();
                const uint32_t st = smem_u32(stages + s * STAGE);
                const uint32_t b_hi = st + NT * A_TILE, b_lo = b_hi + B_PART;
                const bool last_of_slice = ks % SLICE_ST == SLICE_ST - 1;
                if (elect_one()) {
                    // descriptors: only the 14-bit start-address field changes between the instructions of a stage
                    const uint32_t a_lo = ((st >> 4) & 0x3FFFu) | ((16u >> 4) << 16);
                    const uint32_t bh_lo = ((b_hi >> 4) & 0x3FFFu) | ((uint32_t)(LBO_B >> 4) << 16);
                    constexpr uint32_t A_HI = (1024u >> 4) | (1u << 14) | (LAYOUT_SW128 << 29);
                    constexpr uint32_t B_HI = ((uint32_t)SBO_B >> 4) | (1u << 14);
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
#pragma unroll
                        for (int j = 0; j < BK / 8; ++j) {
                            const uint64_t da = make_desc(a_lo + (uint32_t)((t * A_TILE + j * 32) >> 4), A_HI);
                            const uint64_t dbh = make_desc(bh_lo + (uint32_t)((j * 2 * LBO_B) >> 4), B_HI);
                            const uint64_t dbl = make_desc(bh_lo + (uint32_t)((B_PART + j * 2 * LBO_B) >> 4), B_HI);
                            umma_ss(tmem_base + col_main(t, mb), da, dbh, IDESC, (ks % SLICE_ST == 0 && j == 0) ? 0u : 1u);
                            umma_ss(tmem_base + col_cross(t), da, dbl, IDESC, (ks == 0 && j == 0) ? 0u : 1u);
                            umma_ts(tmem_base + col_cross(t), tmem_base + col_alo(t, b) + j * 8, dbh, IDESC, 1u);
                        }
                    }
                    umma_commit(&empty[s]);
                    umma_commit(&alo_empty[b]);
                    if (last_of_slice) umma_commit(&slice_full[mb]);
                    if (ks == nks - 1) umma_commit(cross_full);
                }
                __syncwarp();
                if (last_of_slice) ++slc;
            }
        }
    } else if (warp < CONV_WARPS) {
        // ---- converters: warp w owns rows 32 (w % 4) .. +31 of tile w / 4 ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONV));
        const int t = warp >> 2, q = warp & 3, r = q * 32 + lane;
        const uint32_t sw = (uint32_t)(r & 7);
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            for (int ks = 0; ks < nks; ++ks, ++it) {
                const uint32_t s = it % NST, n = it / NST, b = it & 1u, nb = it >> 1;
                if (lane == 0) { mbar_wait(&full[s], n & 1u); mbar_wait(&alo_empty[b], (nb & 1u) ^ 1u); }
                __syncwarp();
                fence_after();
                const uint8_t* rowp = stages + s * STAGE + t * A_TILE + r * 128;
                uint32_t lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 x = *reinterpret_cast<const float4*>(rowp + ((c ^ sw) << 4));
                    lo[4 * c + 0] = __float_as_uint(x.x - __uint_as_float(__float_as_uint(x.x) & TF32_MASK));
                    lo[4 * c + 1] = __float_as_uint(x.y - __uint_as_float(__float_as_uint(x.y) & TF32_MASK));
                    lo[4 * c + 2] = __float_as_uint(x.z - __uint_as_float(__float_as_uint(x.z) & TF32_MASK));
                    lo[4 * c + 3] = __float_as_uint(x.w - __uint_as_float(__float_as_uint(x.w) & TF32_MASK));
                }
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + col_alo(t, b);
                tmem_st16_nowait(ta, lo);
                tmem_st16_nowait(ta + 16, lo + 16);
                tmem_st_wait();
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&alo_full[b]);
            }
        }
    } else if (warp < CONV_WARPS + DRAIN_WARPS) {
        // ---- drainers + epilogue: warp 8 + d owns rows 32 (d % 4) .. +31 of tile d / 4 ----
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_DRAIN));
        const int d = warp - CONV_WARPS, t = d >> 2, q = d & 3;
        const int dtid = tid - CONV_WARPS * 32;                 // 0 .. 255
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int nsl = nks / SLICE_ST;
        uint32_t slc = 0, item_n = 0;
        int loaded = -1;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
            int stream; int64_t row0;
            decode(item, stream, row0);
            const int kind = stream == 0 ? 0 : 1;
            const int64_t nrows = stream == 0 ? p.nfrm : nv;
            const int64_t row = row0 + t * BM + q * 32 + lane;
            const bool valid = row < nrows;
            const float last = valid ? __ldg(p.lp[stream] + row * (int64_t)p.lp_pitch + (p.H - 1)) : 0.0f;
            if (loaded != kind) {                                // Nyquist row of W^T for this stream kind -> shared memory
                named_bar_sync(1, DRAIN_WARPS * 32);
                if (dtid < 64) wlast_s[dtid] = p.wlast[kind][dtid];
                named_bar_sync(1, DRAIN_WARPS * 32);
                loaded = kind;
            }
            float acc[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) acc[i] = 0.0f;
            for (int sl = 0; sl < nsl; ++sl, ++slc) {
                const uint32_t mb = slc & 1u;
                mbar_wait_warp(&slice_full[mb], (slc >> 1) & 1u, lane);
                fence_after();
                const uint32_t ta = tmem_base + lane_base + col_main(t, mb);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld16_nowait(ta + h * 32, v);
                    tmem_ld16_nowait(ta + h * 32 + 16, v + 16);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[h * 32 + i] += __uint_as_float(v[i]);
                }
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&slice_empty[mb]);
            }
            {
                mbar_wait_warp(cross_full, item_n & 1u, lane);
                fence_after();
                const uint32_t ta = tmem_base + lane_base + col_cross(t);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld16_nowait(ta + h * 32, v);
                    tmem_ld16_nowait(ta + h * 32 + 16, v + 16);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[h * 32 + i] += __uint_as_float(v[i]);
                }
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(cross_empty);
            }
            // ---- this thread's row of mel cepstra: Nyquist-bin term added in float64, then SPTK's float32 output file
            // (src/libaudio.py:593); the cosine matrix / mask / clip follow in k_mel_cos ----
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(p.mc[stream] + row * 64);
#pragma unroll
                for (int j = 0; j < 64; j += 4) {
                    float4 o;
                    o.x = (float)((double)acc[j + 0] + (double)last * (double)wlast_s[j + 0]);
                    o.y = (float)((double)acc[j + 1] + (double)last * (double)wlast_s[j + 1]);
                    o.z = (float)((double)acc[j + 2] + (double)last * (double)wlast_s[j + 2]);
                    o.w = (float)((double)acc[j + 3] + (double)last * (double)wlast_s[j + 3]);
                    dst[j >> 2] = o;
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        fence_after();
        tmem_free(tmem_base, TMEM_COLS);
    }
}

// ---- cosine matrix, mask, clip ------------------------------------------------------------------------------------------
// out[row][o] = sum_j mc[row][j] cos_tab[j][o] in float64 (la.mcep_to_sp_cosmat(alpha=0), src/libaudio.py:605-631, on the
// float32-rounded cepstra), then the log floor of the magnitude stream / the clip of the phase streams.  A register-tiled
// float64 product: CTA = 128 rows x up to 64 outputs, thread = 8 rows x 8 outputs, both operands staged in shared memory.
constexpr int CO_ROWS = 128, CO_THREADS = 128, CO_PITCH = 65;
struct CosParams {
    const float* mc[3]; const double* ct[2]; int n_in[2]; int n_out[2];
    const int32_t* vidx; const int32_t* vcount; int64_t nfrm;
    void* out[3]; int out_f64;
    int g_mag, g_ph;                       // CTAs of the magnitude stream / of each phase stream (grid = g_mag + 2 g_ph)
};

// One 128-row tile: NO = outputs per thread (8 threads across the outputs): 8 covers 64 outputs, 6 covers 48 (phase_dim 45).
template <int NO>
__device__ __forceinline__ void cos_tile(const CosParams& p, int stream, int n_in, int n_out, int64_t row0, int64_t nrows,
                                         const double* __restrict__ ct_s, double* __restrict__ mc_s, int tid) {
    // stage the tile's mel cepstra, widened once: all 16 loads of a thread are issued before the first use
    const float4* src = reinterpret_cast<const float4*>(p.mc[stream] + row0 * 64);
    float4 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = tid + k * CO_THREADS;
        v[k] = row0 + (i >> 4) < nrows ? __ldcs(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();                                                     // the previous tile's product has read mc_s
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = tid + k * CO_THREADS;
        double* d = mc_s + (i >> 4) * CO_PITCH + (i & 15) * 4;
        d[0] = (double)v[k].x; d[1] = (double)v[k].y; d[2] = (double)v[k].z; d[3] = (double)v[k].w;
    }
    __syncthreads();
    // rows tr + 16 i; outputs 16 u + 2 tc + {0, 1}, u < NO / 2: the 8 threads of a row group read 128 contiguous bytes of the
    // table per load (a thread owning 8 consecutive outputs would stride the loads by 64 bytes: 4-way bank conflicts)
    const int tr = tid >> 3, tc = tid & 7;
    double acc[8][NO];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int u = 0; u < NO; ++u) acc[i][u] = 0.0;
#pragma unroll 2
    for (int j = 0; j < n_in; ++j) {
        double c[NO];
        const double2* cp = reinterpret_cast<const double2*>(ct_s + j * 64 + tc * 2);
#pragma unroll
        for (int u = 0; u < NO / 2; ++u) { const double2 t2 = cp[8 * u]; c[2 * u] = t2.x; c[2 * u + 1] = t2.y; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double m = mc_s[(tr + 16 * i) * CO_PITCH + j];
#pragma unroll
            for (int u = 0; u < NO; ++u) acc[i][u] = fma(m, c[u], acc[i][u]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t row = row0 + tr + 16 * i;
        if (row >= nrows) continue;
        const int64_t frame = (stream != 0 && p.vidx) ? (int64_t)p.vidx[row] : row;
#pragma unroll
        for (int u = 0; u < NO; ++u) {
            const int o = 16 * (u >> 1) + 2 * tc + (u & 1);
            if (o >= n_out) continue;
            double s = acc[i][u];
            if (stream == 0) { if (s < -745.0) s = -1.0e10; }            // la.log floor (src/libaudio.py:241-248)
            else s = fmin(fmax(s, -1.0), 1.0);                            // clip (src/magphase.py:2530-2542); these rows are voiced
            if (p.out_f64) reinterpret_cast<double*>(p.out[stream])[frame * n_out + o] = s;
            else reinterpret_cast<float*>(p.out[stream])[frame * n_out + o] = (float)s;
        }
    }
}

// persistent, 1-D grid: the first g_mag CTAs serve the magnitude stream, then g_ph CTAs per phase stream -- in proportion to
// the streams' work (the magnitude stream has every frame and 60 x 60 terms per row, a phase stream the voiced frames and
// 58 x 45: with the same number of CTAs per stream the magnitude CTAs did 59 % of the work on a third of the grid).  A CTA
// stages its stream's cosine table once and walks over the row tiles i, i + g, ...
__global__ void __launch_bounds__(CO_THREADS, 2)
k_mel_cos(const CosParams p) {
    extern __shared__ __align__(16) uint8_t cos_smem[];
    double* ct_s = reinterpret_cast<double*>(cos_smem);                 // [64][64], zero padded
    double* mc_s = ct_s + 64 * 64;                                      // [CO_ROWS][CO_PITCH]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int stream = b < p.g_mag ? 0 : 1 + (b - p.g_mag) / p.g_ph;
    const int idx = b < p.g_mag ? b : (b - p.g_mag) % p.g_ph, g = stream == 0 ? p.g_mag : p.g_ph;
    const int kind = stream == 0 ? 0 : 1;
    const int64_t nrows = (stream == 0 || !p.vcount) ? p.nfrm : (int64_t)*p.vcount;
    if ((int64_t)idx * CO_ROWS >= nrows) return;
    const int n_in = p.n_in[kind], n_out = p.n_out[kind];
    for (int i = tid; i < 64 * 64; i += CO_THREADS) {
        const int j = i >> 6, o = i & 63;
        ct_s[i] = (j < n_in && o < n_out) ? __ldg(p.ct[kind] + j * n_out + o) : 0.0;
    }
    for (int64_t row0 = (int64_t)idx * CO_ROWS; row0 < nrows; row0 += (int64_t)g * CO_ROWS) {
        if (n_out <= 48) cos_tile<6>(p, stream, n_in, n_out, row0, nrows, ct_s, mc_s, tid);
        else cos_tile<8>(p, stream, n_in, n_out, row0, nrows, ct_s, mc_s, tid);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

cudaError_t make_tensor_map_f32_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t row_pitch_floats,
                                   uint32_t box_cols, uint32_t box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (!f || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        fn = (EncodeTiledFn)f;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {row_pitch_floats * sizeof(float)};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

size_t mel_tc_operand_bytes(int fft_len) { return (size_t)((fft_len / 2) / BK) * 2 * B_PART; }

cudaError_t build_warp_matrix_tc(int fft_len, const float* wt32, int ld, float* out, cudaStream_t st) {
    const int n_stages = (fft_len / 2) / BK;
    const int n = n_stages * BK * BN;
    k_split_warp_stage<<<(n + 255) / 256, 256, 0, st>>>(wt32, ld, n_stages, out);
    return cudaGetLastError();
}

// the fused tensor-core path serves the device-resident compressed analysis: float32 log periodograms in the pitched /
// compacted layout, at most 64 coefficients per stream, no row interpolation, no raw-cepstrum output
bool mel_tc_usable(const MelArgs& a) {
    return a.wt_tc_mag && a.wt_tc_ph && a.pre_logp && a.feat_dtype == MPB_F32 && !a.lerp_r0 && !a.raw_mc && a.ncp_max == BN &&
           a.ld_mag == BN && a.ld_ph == BN && a.lp_pitch > 0 && a.n_mag <= 64 && a.n_ph <= 64 && a.partial && a.vidx && a.vcount;
}

cudaError_t launch_mel_warp_tc(const MelArgs& a, cudaStream_t st) {
    const int H = a.fft_len / 2 + 1;
    CUtensorMap maps[3];
    const void* base[3] = {a.mag, a.real, a.imag};
    for (int i = 0; i < 3; ++i) {
        cudaError_t e = make_tensor_map_f32_2d(&maps[i], base[i], (uint64_t)(H - 1), (uint64_t)a.nfrm, (uint64_t)a.lp_pitch, BK, BM);
        if (e != cudaSuccess) return e;
    }
    Params p;
    p.lp[0] = (const float*)a.mag; p.lp[1] = (const float*)a.real; p.lp[2] = (const float*)a.imag; p.lp_pitch = a.lp_pitch;
    p.btc[0] = a.wt_tc_mag; p.btc[1] = a.wt_tc_ph;
    p.wlast[0] = a.wt_mag + (size_t)(H - 1) * a.ld_mag; p.wlast[1] = a.wt_ph + (size_t)(H - 1) * a.ld_ph;
    p.vcount = a.vcount;
    p.nfrm = a.nfrm; p.H = H; p.n_kstages = (H - 1) / BK;
    for (int i = 0; i < 3; ++i) p.mc[i] = a.partial + (size_t)i * (size_t)a.nfrm * 64;
    cudaError_t e = cudaFuncSetAttribute(k_mel_warp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return e;
    const int64_t items_max = 3 * ((a.nfrm + NT * BM - 1) / (NT * BM));
    const int grid = (int)(items_max < a.num_sms ? items_max : a.num_sms);
    if (grid < 1) return cudaSuccess;
    // unvoiced frames keep zeros in the phase streams (src/magphase.py:2527-2528): only voiced rows are written below
    const size_t oes = a.out_dtype == MPB_F64 ? 8 : 4;
    if (a.vidx) {
        e = cudaMemsetAsync(a.out_real, 0, oes * (size_t)a.nfrm * a.phase_dim, st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(a.out_imag, 0, oes * (size_t)a.nfrm * a.phase_dim, st);
        if (e != cudaSuccess) return e;
    }
    k_mel_warp_tc<<<grid, THREADS, SMEM, st>>>(maps[0], maps[1], maps[2], p);
    return cudaGetLastError();
}

// second half: cosine matrix / log floor / clip on the float32 mel cepstra launch_mel_warp_tc left in a.partial
cudaError_t launch_mel_cos(const MelArgs& a, cudaStream_t st) {
    cudaError_t e;
    CosParams c;
    for (int i = 0; i < 3; ++i) c.mc[i] = a.partial + (size_t)i * (size_t)a.nfrm * 64;
    c.ct[0] = a.cos_mag; c.ct[1] = a.cos_ph;
    c.n_in[0] = a.n_mag; c.n_in[1] = a.n_ph; c.n_out[0] = a.n_mag; c.n_out[1] = a.phase_dim;
    c.vidx = a.vidx; c.vcount = a.vcount; c.nfrm = a.nfrm;
    c.out[0] = a.out_mag; c.out[1] = a.out_real; c.out[2] = a.out_imag; c.out_f64 = a.out_dtype == MPB_F64 ? 1 : 0;
    const int cos_smem_bytes = 64 * 64 * 8 + CO_ROWS * CO_PITCH * 8;
    const int64_t tiles = (a.nfrm + CO_ROWS - 1) / CO_ROWS;
    // two CTAs per SM, split by the streams' work; the voiced fraction is only known on the device: assume one half
    const double w_mag = (double)a.n_mag * (a.n_mag <= 48 ? 48 : 64), w_ph = 0.5 * (double)a.n_ph * (a.phase_dim <= 48 ? 48 : 64);
    const int total = 2 * a.num_sms;
    int g_ph = (int)(total * w_ph / (w_mag + 2.0 * w_ph) + 0.5);
    if (g_ph < 1) g_ph = 1;
    int g_mag = total - 2 * g_ph;
    if (g_mag < 1) g_mag = 1;
    if (g_mag > tiles) g_mag = (int)tiles;
    if (g_ph > tiles) g_ph = (int)tiles;
    if (tiles < 1) return cudaSuccess;
    c.g_mag = g_mag; c.g_ph = g_ph;
    e = cudaFuncSetAttribute(k_mel_cos, cudaFuncAttributeMaxDynamicSharedMemorySize, cos_smem_bytes);
    if (e != cudaSuccess) return e;
    k_mel_cos<<<(unsigned)(g_mag + 2 * g_ph), CO_THREADS, cos_smem_bytes, st>>>(c);
    return cudaGetLastError();
}

}  // namespace mpb

// Mel un-warping of the low-dimensional features (synthesis side) on the 5th-generation tensor cores.
//
// Reference: la.sp_mel_unwarp src/libaudio.py:667-684 and phase_uncompress_type1_mcep src/magphase.py:1219-1235 -- one
// fixed linear map per stream (built in float64 by the host mirror):
//     log|X|[f][b] = sum_c mag_mel_log[f][c] U_mag[c][b],  b < H;     real / imag[f][b] = sum_c {real, imag}_mel[f][c] U_ph[c][b],  b < HB
// computed TRANSPOSED so that the epilogue writes coalesced rows:
//     D[128 bins x 128 frames] = U^T[128 bins x K] . X^T[K x 128 frames]         (UMMA M = bins = TMEM lanes, N = frames = columns)
// A warp's 32 lanes hold 32 consecutive bins of one frame: every store instruction writes 128 contiguous bytes of a row.
//
// Precision: "3xTF32" like the warp product (mpb_mel_warp_tc.cu), K <= 64.  The raw float32 operand is the high part
// (kind::tf32 ignores the low 13 mantissa bits), the low parts x - trunc_tf32(x) are prepared once: U^T at plan creation,
// the features by k_unwarp_prep.  One accumulator, cross products FIRST: tcgen05.mma truncates when it accumulates, and the
// truncation error scales with the accumulator -- which is still 2^-11 of its final size while the cross terms are added;
// the 8 accumulations of the main product leave a relative bias below 5e-7.
//
//   k_unwarp_prep    one warp per frame: feature rows -> [row][hi(64) | lo(64)] float32 (zero padded K), the rows of the two
//                    phase streams compacted to the frames that need them; the Nyquist bin of the magnitude stream (bin
//                    H-1 = 16 x 128 + 1 would cost a whole 128-bin tile) is a 60-term dot product done here.
//   k_mel_unwarp_tc  persistent CTAs, 320 threads.  Work item = (stream, 128-frame tile, group of 4 bin tiles): the feature
//                    tile is loaded once per item, the U^T tiles (1 MB in all: L2 resident) stream through a 2-stage ring
//                    (2-D tiled TMA, 128-byte swizzle = the UMMA K-major layout), the accumulators are double buffered in
//                    TMEM.  Items that run at the same time complete whole output rows together (DRAM page locality).
//                    warp 8: TMA producer, warp 9: MMA issuer (both warp-uniform, one elected lane issues),
//                    warps 0-7: epilogue -- tcgen05.ld, exp (magnitude stream), coalesced streaming stores.
#include <stdlib.h>

#include "mpb_kernels.h"
#include "mpb_tc.cuh"

namespace mpb {

namespace {

using namespace tc;

constexpr int UB = 128;                       // bins per tile (UMMA M)
constexpr int UF = 128;                       // frames per tile (UMMA N)
constexpr int UK = 64;                        // padded K
constexpr int XP = 2 * UK;                    // row pitch of the split operands: hi(64) | lo(64) floats
constexpr int ATOM = 128 * 128;               // one 128-row x 128-byte swizzle column block: 16,384 bytes
constexpr int OPND = 4 * ATOM;                // hi k0 | hi k1 | lo k0 | lo k1
constexpr int U_ST = 2;                       // feature-tile ring stages
constexpr int EPI_WARPS = 8;                  // two per TMEM lane quadrant: columns 0-63 / 64-127 of the accumulator
constexpr int U_TMA_WARP = 8, U_MMA_WARP = 9;
constexpr int U_THREADS = 10 * 32;
constexpr int U_SMEM = 1024 + OPND + U_ST * OPND + 128 * 4 + 256;
constexpr int IT_BT_DEFAULT = 4;              // bin tiles per work item (MPB_UNWARP_ITBT overrides: experiments)
constexpr uint32_t U_TMEM = 256;              // 2 accumulators of 128 columns
constexpr uint32_t U_IDESC = idesc_tf32(UB, UF);

struct UParams {
    int64_t nfrm; int n_bins[2]; int ksteps[2];          // bins / K steps of 8: magnitude, phase
    const int32_t* vidx; const int32_t* vcount;          // phase rows: rank -> frame, number of rows
    float* out[3]; int pitch[2];
    int it_bt;                                           // bin tiles per work item
};

// features -> split operands (+ Nyquist bin of the magnitude stream)
template <typename TI>
__global__ void __launch_bounds__(128)
k_unwarp_prep(const TI* __restrict__ mag_mel, const TI* __restrict__ real_mel, const TI* __restrict__ imag_mel, int n_mag, int n_ph,
              const int32_t* __restrict__ cidx, int64_t nfrm, const float* __restrict__ u_nyq, int u_pitch,
              float* __restrict__ x_mag, float* __restrict__ x_re, float* __restrict__ x_im,
              float* __restrict__ out_mag, int out_pitch, int nyq_bin) {
    const int lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (f >= nfrm) return;
    float nyq = 0.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        const float x = c < n_mag ? (float)mag_mel[f * n_mag + c] : 0.0f;
        const float hi = __uint_as_float(__float_as_uint(x) & TF32_MASK);
        x_mag[f * XP + c] = x;
        x_mag[f * XP + UK + c] = x - hi;
        if (c < n_mag) nyq = fmaf(x, __ldg(u_nyq + (size_t)c * u_pitch), nyq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nyq += __shfl_xor_sync(0xffffffffu, nyq, o);
    if (lane == 0) out_mag[f * out_pitch + nyq_bin] = __expf(nyq);
    const int r = cidx[f];
    if (r < 0) return;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        const float a = c < n_ph ? (float)real_mel[f * n_ph + c] : 0.0f;
        const float b = c < n_ph ? (float)imag_mel[f * n_ph + c] : 0.0f;
        x_re[(int64_t)r * XP + c] = a;
        x_re[(int64_t)r * XP + UK + c] = a - __uint_as_float(__float_as_uint(a) & TF32_MASK);
        x_im[(int64_t)r * XP + c] = b;
        x_im[(int64_t)r * XP + UK + c] = b - __uint_as_float(__float_as_uint(b) & TF32_MASK);
    }
}

// U [K][np] float32 (rows pitched, zero padded) -> U^T split: [bin][hi(64) | lo(64)], bins padded to a multiple of 128
__global__ void k_split_unwarp(const float* __restrict__ U, int K, int np, int nbins, int rows_pad, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_pad * UK) return;
    const int c = i % UK, b = i / UK;
    const float u = (c < K && b < nbins) ? U[(size_t)c * np + b] : 0.0f;
    out[(size_t)b * XP + c] = u;
    out[(size_t)b * XP + UK + c] = u - __uint_as_float(__float_as_uint(u) & TF32_MASK);
}

// predicated streaming store without a branch (a branch per element costs more than the store)
__device__ __forceinline__ void st_cs_if(float* ptr, float v, int off) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %2, 0;\n"
        "@p st.global.cs.f32 [%0], %1;\n"
        "}\n" ::"l"(ptr), "f"(v), "r"(off) : "memory");
}

// 64 accumulator columns (frames) of this thread's bin -> rows of the output; off[j]: element offset of column j's row, < 0: none
template <bool EXP>
__device__ __forceinline__ void epilogue_store(uint32_t ta, const int32_t* __restrict__ off, float* __restrict__ Y, bool bin_ok) {
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld16_nowait(ta + c, v);
        tmem_ld16_nowait(ta + c + 16, v + 16);
        tmem_ld_wait();
        if (bin_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int4 o = *reinterpret_cast<const int4*>(off + c + j);
                const int oo[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float y = __uint_as_float(v[j + u]);
                    if (EXP) y = exp2f(y * 1.4426950408889634f);      // --use_fast_math: ex2.approx
                    st_cs_if(Y + oo[u], y, oo[u]);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(U_THREADS, 1)
k_mel_unwarp_tc(const __grid_constant__ CUtensorMap map_u_mag, const __grid_constant__ CUtensorMap map_u_ph,
                const __grid_constant__ CUtensorMap map_x_mag, const __grid_constant__ CUtensorMap map_x_re,
                const __grid_constant__ CUtensorMap map_x_im, const UParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_s = smem;                                   // U^T tile
    uint8_t* b_s = smem + OPND;                            // [U_ST] feature tiles
    int32_t* off_s = reinterpret_cast<int32_t*>(smem + OPND + U_ST * OPND);     // [128] output row offset of every column of the current tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(off_s + 128);
    uint64_t* a_full = bars;            // TMA -> MMA
    uint64_t* a_empty = bars + 1;       // MMA (commit) -> TMA
    uint64_t* b_full = bars + 2;        // [U_ST]
    uint64_t* b_empty = b_full + U_ST;  // [U_ST]
    uint64_t* d_full = b_empty + U_ST;  // [2] MMA (commit) -> epilogue
    uint64_t* d_empty = d_full + 2;     // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (warp reductions: tell the compiler these loaded values are warp-uniform, see mpb_mel_warp_tc.cu)
    const int64_t nv = p.vcount ? (int64_t)__reduce_max_sync(0xffffffffu, (unsigned)*p.vcount) : 0;
    // Work item = (stream, 128-frame tile, group of IT_BT bin tiles): the frame tile's features stay in shared memory, the U^T
    // tiles of the group stream through the ring.  Neighbouring items (= CTAs running at the same time) are the bin groups of
    // ONE frame tile, so whole 8 KB output rows are written within a short window: the stores of a 128-bin tile are 512-byte
    // pieces of 128 different rows, and DRAM wants its pages written while they are open.
    const int ft_mag = (int)((p.nfrm + UF - 1) / UF), ft_ph = (int)((nv + UF - 1) / UF);
    const int bt_mag = (p.n_bins[0] + UB - 1) / UB, bt_ph = (p.n_bins[1] + UB - 1) / UB;
    const int IT_BT = p.it_bt;
    const int bg_mag = (bt_mag + IT_BT - 1) / IT_BT, bg_ph = (bt_ph + IT_BT - 1) / IT_BT;
    const int items_mag = ft_mag * bg_mag, items_ph = ft_ph * bg_ph;
    const int n_items = items_mag + 2 * items_ph;

    if (tid == 0) {
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int i = 0; i < U_ST; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == U_MMA_WARP) tmem_alloc(tmem_slot, U_TMEM);
    if (warp == U_TMA_WARP && lane == 0) {
        tma_prefetch_desc(&map_u_mag); tma_prefetch_desc(&map_u_ph); tma_prefetch_desc(&map_x_mag);
        tma_prefetch_desc(&map_x_re); tma_prefetch_desc(&map_x_im);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, *tmem_slot);

    // item -> stream, frame tile, bin tiles [bt0, bt1)
    auto decode = [&](int item, int& stream, int& ft, int& bt0, int& bt1) {
        if (item < items_mag) { stream = 0; ft = item / bg_mag; bt0 = (item % bg_mag) * IT_BT; bt1 = min(bt0 + IT_BT, bt_mag); }
        else {
            const int j = item - items_mag;
            stream = 1 + j / items_ph;
            const int k = j % items_ph;
            ft = k / bg_ph; bt0 = (k % bg_ph) * IT_BT; bt1 = min(bt0 + IT_BT, bt_ph);
        }
    };

    if (warp == U_TMA_WARP) {
        // ---- TMA producer: a_s = the item's feature tile, b_s ring = U^T tiles ----
        uint32_t it = 0, item_n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
            int stream, ft, bt0, bt1;
            decode(item, stream, ft, bt0, bt1);
            const CUtensorMap* mu = stream == 0 ? &map_u_mag : &map_u_ph;
            const CUtensorMap* mx = stream == 0 ? &map_x_mag : (stream == 1 ? &map_x_re : &map_x_im);
            mbar_wait(a_empty, (item_n & 1u) ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(a_full, OPND);
#pragma unroll
                for (int q = 0; q < 4; ++q) tma_load_2d(a_s + q * ATOM, mx, q * 32, ft * UF, a_full);
            }
            __syncwarp();
            for (int bt = bt0; bt < bt1; ++bt, ++it) {
                const uint32_t s = it % U_ST, n = it / U_ST;
                mbar_wait(&b_empty[s], (n & 1u) ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(&b_full[s], OPND);
#pragma unroll
                    for (int q = 0; q < 4; ++q) tma_load_2d(b_s + s * OPND + q * ATOM, mu, q * 32, bt * UB, &b_full[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp == U_MMA_WARP) {
        // ---- MMA issuer: D[bins x frames] = U^T tile (ring, A operand) . X^T (resident, B operand) ----
        uint32_t it = 0, item_n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_n) {
            int stream, ft, bt0, bt1;
            decode(item, stream, ft, bt0, bt1);
            const int ksteps = p.ksteps[stream == 0 ? 0 : 1];
            mbar_wait(a_full, item_n & 1u);
            for (int bt = bt0; bt < bt1; ++bt, ++it) {
                const uint32_t s = it % U_ST, n = it / U_ST, d = it & 1u, nd = it >> 1;
                mbar_wait(&d_empty[d], (nd & 1u) ^ 1u);
                mbar_wait(&b_full[s], n & 1u);
                fence_after();
                if (elect_one()) {
                    const uint32_t x_lo32 = ((smem_u32(a_s) >> 4) & 0x3FFFu) | ((16u >> 4) << 16);
                    const uint32_t u_lo32 = ((smem_u32(b_s + s * OPND) >> 4) & 0x3FFFu) | ((16u >> 4) << 16);
                    constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (LAYOUT_SW128 << 29);
                    const uint32_t dcol = tmem_base + d * UF;
                    // cross products first (see the header), then the main product; part: 0 = hi, 1 = lo
                    uint32_t acc = 0u;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const int pu = pass == 0 ? 1 : 0, px = pass == 1 ? 1 : 0;     // lo.hi, hi.lo, hi.hi
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (j < ksteps) {
                                const uint32_t off = (uint32_t)(((j >> 2) * ATOM + (j & 3) * 32) >> 4);
                                umma_ss(dcol, make_desc(u_lo32 + (uint32_t)((pu * 2 * ATOM) >> 4) + off, HI),
                                        make_desc(x_lo32 + (uint32_t)((px * 2 * ATOM) >> 4) + off, HI), U_IDESC, acc);
                                acc = 1u;
                            }
                        }
                    }
                    umma_commit(&b_empty[s]);
                    umma_commit(&d_full[d]);
                    if (bt == bt1 - 1) umma_commit(a_empty);
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue: warp w owns bins 32 (w % 4) .. +31 of the tile (TMEM lanes) and columns 64 (w / 4) .. +63 ----
        const int q = warp & 3, half = warp >> 2;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int stream, ft, bt0, bt1;
            decode(item, stream, ft, bt0, bt1);
            const int kind = stream == 0 ? 0 : 1;
            const int64_t nrows = stream == 0 ? p.nfrm : nv;
            const int pitch = p.pitch[kind];
            // element offset of every column's output row (-1: no such frame), once per item: identity rows for the magnitude
            // stream, the compacted rank's frame for the phase streams
            named_bar_sync(1, EPI_WARPS * 32);              // the previous item's stores are issued: off_s may be rewritten
            if (tid < UF) {
                const int64_t r = (int64_t)ft * UF + tid;
                const int fr = r < nrows ? (stream == 0 ? (int)r : p.vidx[r]) : -1;
                off_s[tid] = fr >= 0 ? fr * pitch : -1;
            }
            named_bar_sync(1, EPI_WARPS * 32);
            for (int bt = bt0; bt < bt1; ++bt, ++it) {
                const uint32_t d = it & 1u, nd = it >> 1;
                const int bin = bt * UB + q * 32 + lane;
                const bool bin_ok = bin < p.n_bins[kind];
                float* __restrict__ Y = p.out[stream] + bin;
                mbar_wait_warp(&d_full[d], nd & 1u, lane);
                fence_after();
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + d * UF + half * 64;
                if (stream == 0) epilogue_store<true>(ta, off_s + half * 64, Y, bin_ok);
                else epilogue_store<false>(ta, off_s + half * 64, Y, bin_ok);
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[d]);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == U_MMA_WARP) {
        fence_after();
        tmem_free(tmem_base, U_TMEM);
    }
}

}  // namespace

size_t unwarp_tc_operand_bytes(int nbins) { return (size_t)((nbins + UB - 1) / UB) * UB * XP * sizeof(float); }
size_t unwarp_tc_feature_bytes(int64_t n_rows) { return (size_t)(n_rows + UF) * XP * sizeof(float); }

cudaError_t build_unwarp_matrix_tc(const float* U, int K, int np, int nbins, float* out, cudaStream_t st) {
    const int rows_pad = ((nbins + UB - 1) / UB) * UB;
    const int n = rows_pad * UK;
    k_split_unwarp<<<(n + 255) / 256, 256, 0, st>>>(U, K, np, nbins, rows_pad, out);
    return cudaGetLastError();
}

bool unwarp_tc_usable(const UnwarpArgs& a) {
    return a.ut_mag && a.ut_ph && a.xs[0] && a.vidx && a.n_mag <= UK && a.n_ph <= UK && (a.H - 1) % UB == 0 && a.H - 1 < a.HP;
}

// a.vidx / a.cidx / a.vcount: voiced-frame compaction of a.need_ph, already computed on st (launch_voiced_compact)
cudaError_t launch_mel_unwarp_tc(const UnwarpArgs& a, cudaStream_t st) {
    const int nb_mag = a.H - 1;                 // the Nyquist bin is k_unwarp_prep's
    const unsigned pg = (unsigned)((a.nfrm + 3) / 4);
    if (a.in_dtype == MPB_F64)
        k_unwarp_prep<double><<<pg, 128, 0, st>>>((const double*)a.mag_mel, (const double*)a.real_mel, (const double*)a.imag_mel, a.n_mag,
                                                  a.n_ph, a.cidx, a.nfrm, a.u_mag + nb_mag, a.HP, a.xs[0], a.xs[1], a.xs[2], a.out_mag, a.HP, nb_mag);
    else
        k_unwarp_prep<float><<<pg, 128, 0, st>>>((const float*)a.mag_mel, (const float*)a.real_mel, (const float*)a.imag_mel, a.n_mag,
                                                 a.n_ph, a.cidx, a.nfrm, a.u_mag + nb_mag, a.HP, a.xs[0], a.xs[1], a.xs[2], a.out_mag, a.HP, nb_mag);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    CUtensorMap m[5];
    const int rows_mag = ((nb_mag + UB - 1) / UB) * UB, rows_ph = ((a.HB + UB - 1) / UB) * UB;
    e = make_tensor_map_f32_2d(&m[0], a.ut_mag, XP, (uint64_t)rows_mag, XP, 32, 128);
    if (e == cudaSuccess) e = make_tensor_map_f32_2d(&m[1], a.ut_ph, XP, (uint64_t)rows_ph, XP, 32, 128);
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = make_tensor_map_f32_2d(&m[2 + i], a.xs[i], XP, (uint64_t)a.nfrm, XP, 32, 128);
    if (e != cudaSuccess) return e;
    static const int it_bt = [] { const char* e = getenv("MPB_UNWARP_ITBT"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 64 ? v : IT_BT_DEFAULT; }();
    UParams p;
    p.nfrm = a.nfrm; p.n_bins[0] = nb_mag; p.n_bins[1] = a.HB;
    p.ksteps[0] = (a.n_mag + 7) / 8; p.ksteps[1] = (a.n_ph + 7) / 8;
    p.vidx = a.vidx; p.vcount = a.vcount;
    p.it_bt = it_bt;
    p.out[0] = a.out_mag; p.out[1] = a.out_real; p.out[2] = a.out_imag; p.pitch[0] = a.HP; p.pitch[1] = a.HBP;
    e = cudaFuncSetAttribute(k_mel_unwarp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, U_SMEM);
    if (e != cudaSuccess) return e;
    const int64_t ft = (a.nfrm + UF - 1) / UF;
    const int64_t bt_m = (nb_mag + UB - 1) / UB, bt_p = (a.HB + UB - 1) / UB;
    const int64_t items_max = ft * ((bt_m + it_bt - 1) / it_bt + 2 * ((bt_p + it_bt - 1) / it_bt));
    const int grid = (int)(items_max < a.num_sms ? items_max : a.num_sms);
    if (grid < 1) return cudaSuccess;
    k_mel_unwarp_tc<<<grid, U_THREADS, U_SMEM, st>>>(m[0], m[1], m[2], m[3], m[4], p);
    return cudaGetLastError();
}

}  // namespace mpb

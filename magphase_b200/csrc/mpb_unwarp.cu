// Mel un-warping of the low-dimensional features (synthesis side) and minimum-phase helper tables.
//
// Reference: la.sp_mel_unwarp src/libaudio.py:667-684 (Hermitian-extend the n_c mel-log values, ifft.real,
// double cepstral indices 1..n_c-3, cosine matrix of the warped axis, la.mcep_to_sp_cosmat :605-631) and
// phase_uncompress_type1_mcep src/magphase.py:1219-1235 (pad phase_dim -> nmel by repeating the last column).
// All of that is ONE fixed linear map per stream, built in float64 by the host mirror:
//     log|X|[f][k] = sum_c mag_mel_log[f][c] * U_mag[c][k],      k < H
//     real[f][k]   = sum_c real_mel[f][c]   * U_ph[c][k],        k < HB   (only bins below the crossfade
//     imag[f][k]   = sum_c imag_mel[f][c]   * U_ph[c][k]                   band's upper edge are ever used)
// k_mel_unwarp is the CUDA-core tile product (K = 60 / 45 is tiny, the 12 KB/frame of output dominates).  A CTA owns
// ONE 128-bin tile of one stream -- its slice of U stays in shared memory for the CTA's lifetime -- and walks over the
// frame tiles (64 frames) of its share of the batch: the 64 x K feature tile is one contiguous run of the row-major
// matrix, fetched by a TMA bulk copy that runs under the previous tile's FMA loop, transposed in shared memory, and
// multiplied with 4 x 8 outputs per thread (256 threads: 24 warps per SM hide the FMA latencies); exp() is fused for the
// magnitude stream.
#include "mpb_kernels.h"
#include "mpb_tma.cuh"

namespace mpb {

constexpr int UW_FT = 64;             // frames per tile
constexpr int UW_BT = 128;            // bins per CTA
constexpr int UW_LDX = UW_FT + 4;     // pitch of the transposed feature tile
constexpr int UW_TPB = 256;           // threads per CTA: 4 frames x 8 bins of the 64 x 128 tile each

// flags[t] = 1 when any frame of frame tile t needs its phase rows (tiles without voiced frames are skipped)
__global__ void k_unwarp_tile_flags(const uint8_t* __restrict__ need_ph, int64_t nfrm, uint8_t* __restrict__ flags) {
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tile * UW_FT >= nfrm) return;
    bool any = false;
    for (int i = lane; i < UW_FT; i += 32) {
        const int64_t f = tile * UW_FT + i;
        any |= f < nfrm && need_ph[f] != 0;
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flags[tile] = any ? 1 : 0;
}

template <typename TA, typename TB>
__global__ void k_convert(const TA* __restrict__ a, TB* __restrict__ b, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (TB)a[i];
}

__global__ void __launch_bounds__(UW_TPB, 3)
k_mel_unwarp(const float* __restrict__ mag_mel, const float* __restrict__ real_mel, const float* __restrict__ imag_mel,
             const uint8_t* __restrict__ need_ph, const uint8_t* __restrict__ tile_flags, int64_t nfrm, int n_mag, int n_ph,
             const float* __restrict__ u_mag, const float* __restrict__ u_ph,
             float* __restrict__ out_mag, float* __restrict__ out_real, float* __restrict__ out_imag,
             int tiles_mag, int tiles_ph, int g_mag, int g_ph, int kmax, int HP, int HBP) {
    extern __shared__ __align__(16) float smem_f[];
    const int tid = threadIdx.x;
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
    const int np = stream == 0 ? HP : HBP;                // row pitch of U and of the output scratch (multiple of 4 floats)
    const float* __restrict__ X = stream == 0 ? mag_mel : (stream == 1 ? real_mel : imag_mel);
    const float* __restrict__ U = stream == 0 ? u_mag : u_ph;
    float* __restrict__ Y = stream == 0 ? out_mag : (stream == 1 ? out_real : out_imag);
    const int b0 = btile * UW_BT;
    float* raw = smem_f;                                  // [UW_FT][K]   frame-major, as it sits in HBM (TMA destination)
    float* Xs = raw + UW_FT * kmax;                       // [K][UW_LDX]  transposed: coefficient-major
    float* Us = Xs + kmax * UW_LDX;                       // [K][UW_BT]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(Us + kmax * UW_BT);
    const int64_t n_ft = (nfrm + UW_FT - 1) / UW_FT;

    if (tid == 0) { mbar_init(mbar, 1); fence_proxy_async(); }
    // un-warp matrix tile: rows are pitched to 16 bytes and zero padded on the host side -> float4 copies
    for (int i = tid; i < K * (UW_BT / 4); i += UW_TPB) {
        const int c = i / (UW_BT / 4), b4 = i % (UW_BT / 4);
        const int b = b0 + 4 * b4;
        reinterpret_cast<float4*>(Us)[i] = b < np ? __ldg(reinterpret_cast<const float4*>(U + (size_t)c * np + b))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // frame tiles of this CTA: ft = part, part + parts, ... ; phase streams skip tiles without a voiced frame
    auto next_tile = [&](int64_t ft) {
        while (ft < n_ft && stream != 0 && tile_flags[ft] == 0) ft += parts;
        return ft;
    };
    // a full tile (or a tail whose byte count is a multiple of 16) arrives by TMA, any other tail by plain loads
    auto tile_rows = [&](int64_t ft) { return (int)(nfrm - ft * UW_FT < UW_FT ? nfrm - ft * UW_FT : UW_FT); };
    auto by_tma = [&](int64_t ft) { return ((tile_rows(ft) * K) & 3) == 0; };
    auto fetch = [&](int64_t ft) {                        // thread 0 only
        const uint32_t bytes = (uint32_t)(tile_rows(ft) * K * 4);
        mbar_expect_tx(mbar, bytes);
        tma_load_1d(raw, X + ft * (int64_t)(UW_FT * K), bytes, mbar);
    };
    int64_t ft = next_tile(part);
    __syncthreads();                                      // mbarrier initialised, Us complete
    if (ft < n_ft && by_tma(ft) && tid == 0) fetch(ft);
    uint32_t phase = 0;

    const int tf = tid >> 4, tb = tid & 15;               // frames tf*4 .. tf*4+3, bins {tb*4.., 64+tb*4..}
    const float* px = Xs + tf * 4;
    const float* pu = Us + tb * 4;
    while (ft < n_ft) {
        const int64_t f0 = ft * UW_FT;
        const int rows = tile_rows(ft);
        if (by_tma(ft)) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
        } else {
            for (int i = tid; i < rows * K; i += UW_TPB) raw[i] = X[f0 * K + i];
            __syncthreads();
        }
        // transpose raw[f][c] -> Xs[c][f] (rows past the end of the batch read as zero)
        for (int i = tid; i < UW_FT * K; i += UW_TPB) {
            const int c = i / UW_FT, f = i % UW_FT;       // consecutive threads: consecutive f (conflict-free stores)
            Xs[c * UW_LDX + f] = f < rows ? raw[f * K + c] : 0.0f;
        }
        __syncthreads();                                  // Xs complete, raw free
        const int64_t ft_next = next_tile(ft + parts);
        if (ft_next < n_ft && by_tma(ft_next) && tid == 0) { fence_proxy_async(); fetch(ft_next); }

        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
        for (int c = 0; c < K; ++c) {
            const float4 a0 = *reinterpret_cast<const float4*>(px + c * UW_LDX);
            const float4 b0v = *reinterpret_cast<const float4*>(pu + c * UW_BT);
            const float4 b1v = *reinterpret_cast<const float4*>(pu + c * UW_BT + 64);
            const float a[4] = {a0.x, a0.y, a0.z, a0.w};
            const float b[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t f = f0 + tf * 4 + i;
            if (f >= nfrm) continue;
            if (stream != 0 && need_ph[f] == 0) continue;
            float* py = Y + f * (int64_t)np + b0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {                     // two groups of 4 consecutive bins -> one 16-byte store each
                const int b = h * 64 + tb * 4;
                if (b0 + b >= np) continue;                   // (the pad bins of the last group hold exp(0) / 0: never read)
                float4 o;
                if (stream == 0) o = make_float4(__expf(acc[i][4 * h]), __expf(acc[i][4 * h + 1]), __expf(acc[i][4 * h + 2]), __expf(acc[i][4 * h + 3]));
                else o = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
                *reinterpret_cast<float4*>(py + b) = o;
            }
        }
        __syncthreads();                                  // every thread is done with Xs before the next transpose
        ft = ft_next;
    }
}

// Per-frame magnitude rows of constant-rate input: out[f] = (1 - w) row0 + w row1 of the UN-WARPED rows, the very
// expression k_synthesis_compressed evaluates on the fly (src/magphase.py:861-870).  Only materialised when something
// has to be computed FROM the interpolated row (the minimum phase, :935-936).
__global__ void k_lerp_rows(const float* __restrict__ rows, int pitch, const int32_t* __restrict__ row0,
                            const int32_t* __restrict__ row1, const float* __restrict__ roww, int64_t nfrm,
                            float* __restrict__ out) {
    const int64_t f = blockIdx.x;
    if (f >= nfrm) return;
    const float* a = rows + (int64_t)row0[f] * pitch;
    const float* b = rows + (int64_t)row1[f] * pitch;
    const float w = roww[f];
    for (int k = threadIdx.x; k < pitch; k += blockDim.x) {
        const float x = a[k];
        out[f * pitch + k] = fmaf(w, b[k] - x, x);
    }
}

cudaError_t launch_lerp_rows(const float* rows, int pitch, const int32_t* row0, const int32_t* row1, const float* roww,
                             int64_t nfrm, float* out, cudaStream_t st) {
    if (nfrm < 1) return cudaSuccess;
    k_lerp_rows<<<(unsigned)nfrm, 256, 0, st>>>(rows, pitch, row0, row1, roww, nfrm, out);
    return cudaGetLastError();
}

// in_f32: the three feature matrices as float32 (a float64 caller is narrowed into `cvt` first -- the same rounding
// the tile loader used to apply element by element).  flags: ceil(nfrm / 64) bytes of scratch.
cudaError_t launch_mel_unwarp(const UnwarpArgs& a, cudaStream_t st) {
    const int tiles_mag = (a.H + UW_BT - 1) / UW_BT, tiles_ph = (a.HB + UW_BT - 1) / UW_BT;
    const int kmax = ((a.n_mag > a.n_ph ? a.n_mag : a.n_ph) + 3) & ~3;
    const float *xm = (const float*)a.mag_mel, *xr = (const float*)a.real_mel, *xi = (const float*)a.imag_mel;
    if (a.in_dtype == MPB_F64) {
        const size_t nm = (size_t)a.nfrm * a.n_mag, np_ = (size_t)a.nfrm * a.n_ph;
        float* cm = a.cvt + a.cvt_off_mag;
        float* cr = a.cvt + a.cvt_pitch + a.cvt_off_ph;
        float* ci = a.cvt + 2 * a.cvt_pitch + a.cvt_off_ph;
        k_convert<double, float><<<(unsigned)((nm + 255) / 256), 256, 0, st>>>((const double*)a.mag_mel, cm, nm);
        k_convert<double, float><<<(unsigned)((np_ + 255) / 256), 256, 0, st>>>((const double*)a.real_mel, cr, np_);
        k_convert<double, float><<<(unsigned)((np_ + 255) / 256), 256, 0, st>>>((const double*)a.imag_mel, ci, np_);
        xm = cm; xr = cr; xi = ci;
    }
    const int64_t n_ft = (a.nfrm + UW_FT - 1) / UW_FT;
    k_unwarp_tile_flags<<<(unsigned)((n_ft + 3) / 4), 128, 0, st>>>(a.need_ph, a.nfrm, a.flags);
    const size_t smem = sizeof(float) * ((size_t)UW_FT * kmax + (size_t)kmax * UW_LDX + (size_t)kmax * UW_BT) + 16;
    cudaError_t e = cudaFuncSetAttribute(k_mel_unwarp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mel_unwarp, UW_TPB, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    // CTAs per bin tile, proportional to the work behind a tile: K = n_mag for every frame tile of the magnitude stream,
    // K = n_ph for the frame tiles of the phase streams that hold at least one voiced frame (nearly all of them in
    // running speech: a tile spans ~0.4 s)
    const int slots = a.num_sms * per_sm;
    const double w_mag = (double)tiles_mag * a.n_mag, w_ph = 2.0 * tiles_ph * a.n_ph * 0.95;
    int g_ph = (int)((double)slots * w_ph / (w_mag + w_ph) / (2 * tiles_ph));
    if (g_ph < 1) g_ph = 1;
    int g_mag = (slots - 2 * tiles_ph * g_ph) / tiles_mag;
    if (g_mag < 1) g_mag = 1;
    if (g_mag > n_ft) g_mag = (int)n_ft;
    if (g_ph > n_ft) g_ph = (int)n_ft;
    const unsigned grid = (unsigned)(tiles_mag * g_mag + 2 * tiles_ph * g_ph);
    k_mel_unwarp<<<grid, UW_TPB, smem, st>>>(xm, xr, xi, a.need_ph, a.flags, a.nfrm, a.n_mag, a.n_ph, a.u_mag, a.u_ph, a.out_mag,
                                          a.out_real, a.out_imag, tiles_mag, tiles_ph, g_mag, g_ph, kmax, a.HP, a.HBP);
    return cudaGetLastError();
}

}  // namespace mpb

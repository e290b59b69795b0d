// Mel un-warping of the low-dimensional features (synthesis side) and minimum-phase helper tables.
//
// Reference: la.sp_mel_unwarp src/libaudio.py:667-684 (Hermitian-extend the n_c mel-log values, ifft.real,
// double cepstral indices 1..n_c-3, cosine matrix of the warped axis, la.mcep_to_sp_cosmat :605-631) and
// phase_uncompress_type1_mcep src/magphase.py:1219-1235 (pad phase_dim -> nmel by repeating the last column).
// All of that is ONE fixed linear map per stream, built in float64 by the host mirror:
//     log|X|[f][k] = sum_c mag_mel_log[f][c] * U_mag[c][k],      k < H
//     real[f][k]   = sum_c real_mel[f][c]   * U_ph[c][k],        k < HB   (only bins below the crossfade
//     imag[f][k]   = sum_c imag_mel[f][c]   * U_ph[c][k]                   band's upper edge are ever used)
// k_mel_unwarp is the CUDA-core tile product (K = 60 / 45 is tiny, the 12 KB/frame of output dominates):
// 64 frames x 128 bins per CTA, 8 x 8 outputs per thread, exp() fused for the magnitude stream.
#include "mpb_kernels.h"

namespace mpb {

constexpr int UW_FT = 64;             // frames per CTA tile
constexpr int UW_BT = 128;            // bins per CTA tile
constexpr int UW_LDX = UW_FT + 4;     // pitch of the transposed feature tile

template <typename TI>
__global__ void __launch_bounds__(128, 4)
k_mel_unwarp(const TI* __restrict__ mag_mel, const TI* __restrict__ real_mel, const TI* __restrict__ imag_mel,
             const uint8_t* __restrict__ need_ph, int64_t nfrm, int n_mag, int n_ph,
             const float* __restrict__ u_mag, int H, const float* __restrict__ u_ph, int HB,
             float* __restrict__ out_mag, float* __restrict__ out_real, float* __restrict__ out_imag,
             int tiles_mag, int tiles_ph, int kmax, int HP, int HBP) {
    extern __shared__ __align__(16) float smem_f[];
    const int tid = threadIdx.x;
    int stream, btile;
    if ((int)blockIdx.y < tiles_mag) { stream = 0; btile = blockIdx.y; }
    else { stream = 1 + ((int)blockIdx.y - tiles_mag) / tiles_ph; btile = ((int)blockIdx.y - tiles_mag) % tiles_ph; }
    const int K = stream == 0 ? n_mag : n_ph;
    const int np = stream == 0 ? HP : HBP;                // row pitch of U and of the output scratch (multiple of 4 floats)
    const TI* __restrict__ X = stream == 0 ? mag_mel : (stream == 1 ? real_mel : imag_mel);
    const float* __restrict__ U = stream == 0 ? u_mag : u_ph;
    float* __restrict__ Y = stream == 0 ? out_mag : (stream == 1 ? out_real : out_imag);
    const int64_t f0 = (int64_t)blockIdx.x * UW_FT;
    const int b0 = btile * UW_BT;
    float* Xs = smem_f;                                   // [K][UW_LDX]  (transposed: coefficient-major)
    float* Us = smem_f + kmax * UW_LDX;                   // [K][UW_BT]

    if (stream != 0) {                                    // skip tiles where no frame needs phase
        bool any = false;
        for (int i = 0; i < UW_FT && f0 + i < nfrm; ++i) any |= need_ph[f0 + i] != 0;
        if (!any) return;
    }
    // features: the 64 x K tile is one contiguous run of the row-major matrix -> linear coalesced read, transposed store
    for (int i = tid; i < UW_FT * K; i += 128) {
        const int f = i / K, c = i % K;
        Xs[c * UW_LDX + f] = (f0 + f < nfrm) ? (float)X[(f0 + f) * (int64_t)K + c] : 0.0f;
    }
    // un-warp matrix tile: rows are pitched to 16 bytes and zero padded on the host side -> float4 copies
    for (int i = tid; i < K * (UW_BT / 4); i += 128) {
        const int c = i / (UW_BT / 4), b4 = i % (UW_BT / 4);
        const int b = b0 + 4 * b4;
        reinterpret_cast<float4*>(Us)[i] = b < np ? __ldg(reinterpret_cast<const float4*>(U + (size_t)c * np + b))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    const int tf = tid >> 4, tb = tid & 15;               // frames {tf*4.., 32+tf*4..}, bins {tb*4.., 64+tb*4..}
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    const float* px = Xs + tf * 4;
    const float* pu = Us + tb * 4;
#pragma unroll 4
    for (int c = 0; c < K; ++c) {
        const float4 a0 = *reinterpret_cast<const float4*>(px + c * UW_LDX);
        const float4 a1 = *reinterpret_cast<const float4*>(px + c * UW_LDX + 32);
        const float4 b0v = *reinterpret_cast<const float4*>(pu + c * UW_BT);
        const float4 b1v = *reinterpret_cast<const float4*>(pu + c * UW_BT + 64);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t f = f0 + (i < 4 ? tf * 4 + i : 32 + tf * 4 + (i - 4));
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
}

cudaError_t launch_mel_unwarp(const UnwarpArgs& a, cudaStream_t st) {
    const int tiles_mag = (a.H + UW_BT - 1) / UW_BT, tiles_ph = (a.HB + UW_BT - 1) / UW_BT;
    const int kmax = a.n_mag > a.n_ph ? a.n_mag : a.n_ph;
    const size_t smem = sizeof(float) * ((size_t)kmax * UW_LDX + (size_t)kmax * UW_BT);
    dim3 grid((unsigned)((a.nfrm + UW_FT - 1) / UW_FT), (unsigned)(tiles_mag + 2 * tiles_ph));
    cudaError_t e;
    if (a.in_dtype == MPB_F64) {
        e = cudaFuncSetAttribute(k_mel_unwarp<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_mel_unwarp<double><<<grid, 128, smem, st>>>((const double*)a.mag_mel, (const double*)a.real_mel,
                                                      (const double*)a.imag_mel, a.need_ph, a.nfrm, a.n_mag, a.n_ph,
                                                      a.u_mag, a.H, a.u_ph, a.HB, a.out_mag, a.out_real, a.out_imag,
                                                      tiles_mag, tiles_ph, kmax, a.HP, a.HBP);
    } else {
        e = cudaFuncSetAttribute(k_mel_unwarp<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_mel_unwarp<float><<<grid, 128, smem, st>>>((const float*)a.mag_mel, (const float*)a.real_mel,
                                                     (const float*)a.imag_mel, a.need_ph, a.nfrm, a.n_mag, a.n_ph,
                                                     a.u_mag, a.H, a.u_ph, a.HB, a.out_mag, a.out_real, a.out_imag,
                                                     tiles_mag, tiles_ph, kmax, a.HP, a.HBP);
    }
    return cudaGetLastError();
}

}  // namespace mpb

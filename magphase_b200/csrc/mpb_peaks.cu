// Measured non-tensor FMA peaks of the device (roofline denominators of the compute-bound kernels):
// register-resident FFMA / DFMA chains, no memory traffic, every SM filled.  bench.py reports the FFT and cosine-matrix
// kernels against these numbers instead of a figure derived from the clock.
#include "mpb_ctx.h"

namespace {

template <typename T>
__global__ void __launch_bounds__(256, 4)
k_fma_peak(T* __restrict__ out, T a, T b, int iters) {
    T acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = acc[i] * a + b;      // contracted to one FMA per accumulator
    }
    T s = (T)0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == (T)123456789) out[0] = s;                              // keeps the chains alive, never true in practice
}

template <typename T>
int measure(mpb_ctx* ctx, double* tflops) {
    cudaStream_t st = ctx->stream;
    T* d = nullptr;
    CU(cudaMalloc(&d, sizeof(T)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int grid = ctx->num_sms * 8, iters = sizeof(T) == 8 ? 4096 : 8192;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {                             // first repetition warms up
        CU(cudaEventRecord(e0, st));
        k_fma_peak<T><<<grid, 256, 0, st>>>(d, (T)0.999999, (T)1e-7, iters);
        CU(cudaEventRecord(e1, st));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 16.0 * (double)iters * 256.0 * (double)grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    ctx->launches += 4;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return MPB_OK;
}

}  // namespace

extern "C" int mpb_measure_fma_peak(mpb_ctx* ctx, int dtype, double* tflops) {
    if (!ctx || !tflops) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (!mpb::dtype_ok(dtype)) return fail(MPB_ERR_BAD_ARG, "unknown dtype");
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    return dtype == MPB_F64 ? measure<double>(ctx, tflops) : measure<float>(ctx, tflops);
}

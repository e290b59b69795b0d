// NumPy-legacy-stream noise on the device.
//
// The reference draws the aperiodic excitation with np.random.uniform(-1, 1, ns_len) on NumPy's global legacy
// MT19937 stream (src/magphase.py:883); parity is only defined with the SAME numbers.  Generating millions of
// doubles with the host generator dominates the end-to-end time, so the stream itself is reproduced here,
// bit for bit: the host hands over NumPy's state (624 words + position), one CTA advances the twister in shared
// memory (the recurrence x[k+624] = x[k+397] ^ f(x[k], x[k+1]) allows 227 / 227 / 170 words per step in
// parallel) and writes tempered 32-bit outputs; a second, fully parallel kernel pairs them into doubles exactly
// like NumPy's random_sample: ((a >> 5) * 2^26 + (b >> 6)) / 2^53, then low + (high - low) * r.
// The final state goes back to the host so that np.random continues where the reference would.
#include "mpb_kernels.h"

namespace mpb {

__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

__global__ void __launch_bounds__(256)
k_mt19937_stream(uint32_t* __restrict__ key, int32_t* __restrict__ pos_io, uint32_t* __restrict__ out, int64_t n32) {
    __shared__ uint32_t mt[624];
    const int t = threadIdx.x;
    for (int i = t; i < 624; i += 256) mt[i] = key[i];
    int pos = *pos_io;
    __syncthreads();
    int64_t done = 0;
    while (done < n32) {
        if (pos >= 624) {
            uint32_t v = 0;
            if (t < 227) v = mt[t + 397] ^ mt_mix(mt[t], mt[t + 1]);
            __syncthreads();
            if (t < 227) mt[t] = v;
            __syncthreads();
            if (t < 227) v = mt[t] ^ mt_mix(mt[227 + t], mt[228 + t]);
            __syncthreads();
            if (t < 227) mt[227 + t] = v;
            __syncthreads();
            if (t < 170) v = mt[227 + t] ^ mt_mix(mt[454 + t], mt[t == 169 ? 0 : 455 + t]);
            __syncthreads();
            if (t < 170) mt[454 + t] = v;
            __syncthreads();
            pos = 0;
        }
        const int64_t left = n32 - done;
        const int take = (int)(left < (int64_t)(624 - pos) ? left : (int64_t)(624 - pos));
        for (int i = t; i < take; i += 256) out[done + i] = mt_temper(mt[pos + i]);
        done += take;
        pos += take;
        __syncthreads();
    }
    for (int i = t; i < 624; i += 256) key[i] = mt[i];
    if (t == 0) *pos_io = pos;
}

template <typename TO>
__global__ void k_mt_to_uniform(const uint32_t* __restrict__ raw, int64_t n, double low, double scale, TO* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 ab = reinterpret_cast<const uint2*>(raw)[i];
    const double r = ((double)(ab.x >> 5) * 67108864.0 + (double)(ab.y >> 6)) / 9007199254740992.0;
    out[i] = (TO)(low + scale * r);
}

cudaError_t launch_mt19937_uniform(uint32_t* key_dev, int32_t* pos_dev, uint32_t* raw_dev, int64_t n, double low,
                                   double high, void* out, int out_dtype, cudaStream_t st) {
    if (n < 1) return cudaSuccess;
    k_mt19937_stream<<<1, 256, 0, st>>>(key_dev, pos_dev, raw_dev, 2 * n);
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (out_dtype == MPB_F64) k_mt_to_uniform<double><<<grid, 256, 0, st>>>(raw_dev, n, low, high - low, (double*)out);
    else k_mt_to_uniform<float><<<grid, 256, 0, st>>>(raw_dev, n, low, high - low, (float*)out);
    return cudaGetLastError();
}

}  // namespace mpb

// NumPy-legacy-stream noise on the device.
//
// The reference draws the aperiodic excitation with np.random.uniform(-1, 1, ns_len) on NumPy's global legacy
// MT19937 stream (src/magphase.py:883); parity is only defined with the SAME numbers.  Generating millions of
// doubles with the host generator dominates the end-to-end time, so the stream itself is reproduced here,
// bit for bit: the host hands over NumPy's state (624 words + position), one CTA advances the twister in shared
// memory (623 words per barrier, see k_mt19937_stream) and writes tempered 32-bit outputs; a second, fully parallel kernel pairs them into doubles exactly
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

// x[] is the twister's word stream: x[0..623] = the state block NumPy handed over, x[n] = x[n-227] ^ f(x[n-624], x[n-623]).
// f is GF(2)-linear, so the recurrence can be substituted into itself:
//   x[n] = x[n-681] ^ f(n-1078) ^ f(n-851) ^ f(n-624),      f(m) := mt_mix(x[m], x[m+1])
// which only reaches >= 623 words back: 623 new words per barrier instead of 227 (the first 454 words use the plain
// form).  Words live in a 2048-entry ring in shared memory; tempered outputs x[pos0 .. pos0+n32) stream to HBM.
constexpr int MT_RING = 2048;
constexpr int MT_THREADS = 640;

__global__ void __launch_bounds__(MT_THREADS)
k_mt19937_stream(uint32_t* __restrict__ key, int32_t* __restrict__ pos_io, uint32_t* __restrict__ out, int64_t n32) {
    __shared__ uint32_t x[MT_RING];
    const int t = threadIdx.x;
    for (int i = t; i < 624; i += MT_THREADS) x[i] = key[i];
    const int64_t pos0 = *pos_io;
    const int64_t last = pos0 + n32 - 1;                    // stream index of the last consumed word
    const int64_t blk = last / 624;                         // NumPy's final state block ...
    const int64_t need = (blk + 1) * 624;                   // ... must be generated to its end
    __syncthreads();
    for (int64_t i = pos0 + t; i < 624 && i <= last; i += MT_THREADS) out[i - pos0] = mt_temper(x[i]);
    auto X = [&](int64_t n) -> uint32_t { return x[n & (MT_RING - 1)]; };
    int64_t n0 = 624;
    // start-up: two plain waves of 227 words bring the history to 1078 words
    for (int w = 0; w < 2 && n0 < need; ++w) {
        const int64_t n = n0 + t;
        if (t < 227 && n < need) {
            const uint32_t v = X(n - 227) ^ mt_mix(X(n - 624), X(n - 623));
            x[n & (MT_RING - 1)] = v;
            if (n >= pos0 && n <= last) out[n - pos0] = mt_temper(v);
        }
        n0 += 227;
        __syncthreads();
    }
    while (n0 < need) {
        const int64_t n = n0 + t;
        if (t < 623 && n < need) {
            const uint32_t v = X(n - 681) ^ mt_mix(X(n - 1078), X(n - 1077)) ^ mt_mix(X(n - 851), X(n - 850)) ^
                               mt_mix(X(n - 624), X(n - 623));
            x[n & (MT_RING - 1)] = v;
            if (n >= pos0 && n <= last) out[n - pos0] = mt_temper(v);
        }
        n0 += 623;
        __syncthreads();
    }
    // hand NumPy its state back: the block holding the last consumed word, position just past it
    if (n32 > 0) {
        for (int i = t; i < 624; i += MT_THREADS) key[i] = X(blk * 624 + i);
        if (t == 0) *pos_io = (int32_t)(last + 1 - blk * 624);
    }
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
    k_mt19937_stream<<<1, MT_THREADS, 0, st>>>(key_dev, pos_dev, raw_dev, 2 * n);
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (out_dtype == MPB_F64) k_mt_to_uniform<double><<<grid, 256, 0, st>>>(raw_dev, n, low, high - low, (double*)out);
    else k_mt_to_uniform<float><<<grid, 256, 0, st>>>(raw_dev, n, low, high - low, (float*)out);
    return cudaGetLastError();
}

}  // namespace mpb

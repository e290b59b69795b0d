// NumPy-legacy-stream noise on the device.
//
// The reference draws the aperiodic excitation with np.random.uniform(-1, 1, ns_len) on NumPy's global legacy
// MT19937 stream (src/magphase.py:883); parity is only defined with the SAME numbers.  Generating millions of
// doubles with the host generator dominates the end-to-end time, so the stream itself is reproduced here,
// bit for bit: the host hands over NumPy's state (624 words + position), the device advances the twister and writes
// tempered 32-bit outputs; a second, fully parallel kernel pairs them into doubles exactly like NumPy's
// random_sample: ((a >> 5) * 2^26 + (b >> 6)) / 2^53, then low + (high - low) * r.  The final state goes back to the
// host so that np.random continues where the reference would.
//
// The twister is sequential, so it is parallelised twice:
//  * inside a CTA, 623 words per barrier via the substituted recurrence (mt_generate);
//  * across CTAs, by cutting the stream into segments of MT_SEG_WORDS words and JUMPING to the start of each.  The
//    state transition is GF(2)-linear with characteristic polynomial phi (degree 19937), so with
//    g_J(x) = x^J mod phi every word of the stream obeys  x[m+J] = XOR_{i : g_J[i] = 1} x[m+i]  (Cayley-Hamilton):
//    a CTA generates the 19937+623 words that follow its current block and folds them with the set bits of g_J.
//    A fold is 19937 x 624 / 2 word XORs out of shared memory (25 MB of shared-memory traffic) -- as much as generating
//    0.8 M words -- so the number of folds is what a draw costs: segment s = d0 + 256 d1 is reached with ONE fold by
//    g_{d0 J} (255 polynomials) plus one by g_{256 d1 J} (d1 = 1..3) for the rare launch beyond 256 segments.  (Base-4
//    digits, round 1: up to four folds per segment, 1.95 ms for the 30.7 M draws of a bench step.)  phi comes from
//    Berlekamp-Massey on the twister's own output and the polynomials from square-and-multiply / repeated
//    multiplication, once per process on host threads (jump_tables) -- no magic tables.
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "mpb_kernels.h"

namespace mpb {

__host__ __device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b) {
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

constexpr int MT_THREADS = 1024;
constexpr int MT_DEG = 19937;
constexpr int MT_SEG_TWISTS = 1024;                           // a fold costs as much as ~1300 twists of generation: few, long segments
constexpr int MT_SEG_WORDS = MT_SEG_TWISTS * 624;              // words per segment (a multiple of the block size)
constexpr int MT_NPOLY = 255 + 3;                              // g_{d0 J}, d0 = 1..255, then g_{256 d1 J}, d1 = 1..3
constexpr int MT_MAX_SEGS = 1024;                              // segments per launch
constexpr int MT_HDR = 2 * MT_NPOLY * 2;                       // uint16 slots of the table header: int32 off[], cnt[]
constexpr int MT_WIN = MT_DEG + 623;                           // words a fold reads: x[i + t], i < 19937, t < 624
constexpr int MT_PAD = 768;                                    // zero words behind the window (padding index target)
constexpr int MT_WIN_ALLOC = MT_WIN + MT_PAD;
constexpr int MT_RING = 2048;
constexpr int MT_IDX_STAGE = 5632;                             // uint16 indices staged per pass of a fold (11 KB: 2 CTAs per SM)

struct MtJumpArgs {
    const uint16_t* idx;                    // header (int32 off[MT_NPOLY], cnt[MT_NPOLY], in uint16 units behind the header),
};                                          // then the concatenated set-bit lists, each padded to a multiple of 16

// x[] is the twister's word stream: x[n] = x[n-227] ^ f(x[n-624], x[n-623]),  f(m) := mt_mix(x[m], x[m+1]).
// f is GF(2)-linear, so the recurrence can be substituted into itself:
//   x[n] = x[n-681] ^ f(n-1078) ^ f(n-851) ^ f(n-624)
// which only reaches >= 623 words back: 623 new words per barrier instead of 227 (the first 454 words use the plain
// form).  Extends w[0 .. have) to w[0 .. want) in a FLAT array; all threads of the CTA must call it.
__device__ __forceinline__ void mt_extend_flat(uint32_t* __restrict__ w, int have, int want, int t) {
    int n0 = have;
    for (int k = 0; k < 2 && n0 < want && n0 < 1078; ++k) {
        const int n = n0 + t;
        if (t < 227 && n < want) w[n] = w[n - 227] ^ mt_mix(w[n - 624], w[n - 623]);
        n0 += 227;
        __syncthreads();
    }
    while (n0 < want) {
        const int n = n0 + t;
        if (t < 623 && n < want)
            w[n] = w[n - 681] ^ mt_mix(w[n - 1078], w[n - 1077]) ^ mt_mix(w[n - 851], w[n - 850]) ^
                   mt_mix(w[n - 624], w[n - 623]);
        n0 += 623;
        __syncthreads();
    }
}

// One fold: w[0..623] = x[m .. m+623]  ->  w[0..623] = x[m+J .. m+J+623] for the polynomial given as a set-bit list.
// part: 4 x 768 words of scratch; idx_s: MT_IDX_STAGE uint16 of shared memory.  The set-bit list (about 10,000 indices) is
// staged through shared memory in two halves with coalesced 16-byte loads: read straight from global memory, the inner
// loop waited one L2 latency per four indices (measured: 2.1 ms for the 30.7 M draws of a bench step, 80 % of it here).
__device__ __forceinline__ void mt_fold(uint32_t* __restrict__ w, uint32_t* __restrict__ part, uint16_t* __restrict__ idx_s,
                                        const uint16_t* __restrict__ idx, int cnt, int t) {
    mt_extend_flat(w, 624, MT_WIN, t);
    const int q = t >> 8, u = t & 255;
    const int per = cnt >> 2;                                  // cnt is a multiple of 16: per is a multiple of 4
    const uint32_t* wu = w + u;
    uint32_t a0 = 0, a1 = 0, a2 = 0;
    // quarter q of the CTA owns indices [q * per, (q + 1) * per); each pass stages `chunk` indices of every quarter
    for (int done = 0; done < per; done += MT_IDX_STAGE / 4) {
        const int chunk = min(per - done, MT_IDX_STAGE / 4);   // multiple of 4
        __syncthreads();                                       // the previous pass has read idx_s
        for (int i = t; i < chunk; i += MT_THREADS) {  // 2-byte loads, consecutive threads consecutive indices
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) idx_s[qq * (MT_IDX_STAGE / 4) + i] = __ldg(idx + qq * per + done + i);
        }
        __syncthreads();
        const uint2* p = reinterpret_cast<const uint2*>(idx_s + q * (MT_IDX_STAGE / 4));
#pragma unroll 4
        for (int i = 0; i < chunk / 4; ++i) {
            const uint2 v = p[i];
            const int i0 = v.x & 0xffff, i1 = v.x >> 16, i2 = v.y & 0xffff, i3 = v.y >> 16;
            a0 ^= wu[i0] ^ wu[i1];          a1 ^= wu[i0 + 256] ^ wu[i1 + 256];   a2 ^= wu[i0 + 512] ^ wu[i1 + 512];
            a0 ^= wu[i2] ^ wu[i3];          a1 ^= wu[i2 + 256] ^ wu[i3 + 256];   a2 ^= wu[i2 + 512] ^ wu[i3 + 512];
        }
    }
    part[q * 768 + u] = a0;
    part[q * 768 + u + 256] = a1;
    part[q * 768 + u + 512] = a2;
    __syncthreads();
    if (t < 624) w[t] = part[t] ^ part[768 + t] ^ part[2 * 768 + t] ^ part[3 * 768 + t];
    __syncthreads();
}

// Segment s of the launch: stream words [s*MT_SEG_WORDS, (s+1)*MT_SEG_WORDS) relative to the handed block
// x[0..623] = key_in.  Tempered words x[pos0 .. pos0+n32) go to out; the CTA that owns the last consumed word also
// writes NumPy's final state (the 624-word block holding that word, position just past it) to key_out / pos_out.
__global__ void __launch_bounds__(MT_THREADS)
k_mt19937_stream(const uint32_t* __restrict__ key_in, int32_t pos_in, uint32_t* __restrict__ key_out,
                 int32_t* __restrict__ pos_out, uint32_t* __restrict__ out, int64_t n32, MtJumpArgs jt) {
    extern __shared__ uint32_t sm[];
    uint32_t* w = sm;                              // MT_WIN_ALLOC words: fold window, later the generation ring
    uint32_t* part = sm + MT_WIN_ALLOC;            // 4 x 768
    uint16_t* idx_s = reinterpret_cast<uint16_t*>(part + 4 * 768);   // MT_IDX_STAGE staged set-bit indices
    const int t = threadIdx.x;
    const int s = blockIdx.x;
    const int64_t pos0 = pos_in;
    const int64_t last = pos0 + n32 - 1;                    // stream index of the last consumed word
    const int64_t blk = last / 624;                         // NumPy's final state block ...
    const int64_t need = (blk + 1) * 624;                   // ... must be generated to its end
    const int64_t base = (int64_t)s * MT_SEG_WORDS;
    if (base >= need) return;
    for (int i = t; i < 624; i += MT_THREADS) w[i] = key_in[i];
    for (int i = MT_WIN + t; i < MT_WIN_ALLOC; i += MT_THREADS) w[i] = 0u;
    __syncthreads();
    if (s > 0) {
        if (t == 0) {
            // The low 31 bits of x[0] are not part of the twister's state (a freshly seeded key holds arbitrary bits
            // there).  The fold treats x[0] as a full stream word: give it the bits the recurrence implies,
            // x[623] = x[396] ^ f(upper(x[-1]) | lower(x[0])).
            const uint32_t v = w[623] ^ w[396];
            const uint32_t odd = v >> 31;
            const uint32_t y = ((v ^ (odd ? 0x9908b0dfu : 0u)) << 1) | odd;
            w[0] = (w[0] & 0x80000000u) | (y & 0x7fffffffu);
        }
        __syncthreads();
        const int32_t* hdr = reinterpret_cast<const int32_t*>(jt.idx);
        const uint16_t* lists = jt.idx + MT_HDR;
        const int d0 = s & 255, d1 = s >> 8;
        if (d1) mt_fold(w, part, idx_s, lists + hdr[255 + d1 - 1], hdr[MT_NPOLY + 255 + d1 - 1], t);
        if (d0) mt_fold(w, part, idx_s, lists + hdr[d0 - 1], hdr[MT_NPOLY + d0 - 1], t);
    }
    // ---- generation: ring of MT_RING words, ring index = (stream index - base) mod MT_RING ----
    const int64_t end = (base + MT_SEG_WORDS < need) ? base + MT_SEG_WORDS : need;
    const int len = (int)(end - base);                      // words of this segment (>= 624)
    for (int i = t; i < 624; i += MT_THREADS) {
        const int64_t n = base + i;
        if (n >= pos0 && n <= last) out[n - pos0] = mt_temper(w[i]);
    }
    auto X = [&](int n) -> uint32_t { return w[n & (MT_RING - 1)]; };
    const int64_t olo = pos0 - base, ohi = last - base;     // output range in segment-relative indices
    uint32_t* o = out + (base - pos0);
    int n0 = 624;
    for (int k = 0; k < 2 && n0 < len; ++k) {
        const int n = n0 + t;
        if (t < 227 && n < len) {
            const uint32_t v = X(n - 227) ^ mt_mix(X(n - 624), X(n - 623));
            w[n & (MT_RING - 1)] = v;
            if (n >= olo && n <= ohi) o[n] = mt_temper(v);
        }
        n0 += 227;
        __syncthreads();
    }
    while (n0 < len) {
        const int n = n0 + t;
        if (t < 623 && n < len) {
            const uint32_t v = X(n - 681) ^ mt_mix(X(n - 1078), X(n - 1077)) ^ mt_mix(X(n - 851), X(n - 850)) ^
                               mt_mix(X(n - 624), X(n - 623));
            w[n & (MT_RING - 1)] = v;
            if (n >= olo && n <= ohi) o[n] = mt_temper(v);
        }
        n0 += 623;
        __syncthreads();
    }
    if (end == need) {          // this CTA owns the final block
        const int b0 = (int)(blk * 624 - base);
        for (int i = t; i < 624; i += MT_THREADS) key_out[i] = X(b0 + i);
        if (t == 0) *pos_out = (int32_t)(last + 1 - blk * 624);
    }
}

template <typename TO>
__global__ void k_mt_to_uniform(const uint32_t* __restrict__ raw, int64_t n, double low, double scale, TO* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 ab = reinterpret_cast<const uint2*>(raw)[i];
    const double r = ((double)(ab.x >> 5) * 67108864.0 + (double)(ab.y >> 6)) / 9007199254740992.0;
    out[i] = (TO)__dadd_rn(low, __dmul_rn(scale, r));     // NumPy rounds the product, then the sum (no fused multiply-add)
}

// ---------------------------------------------------------------------------------------------
// Host side: GF(2) polynomial arithmetic for the jump polynomials (64-bit limbs, little endian).
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int PW = (MT_DEG + 64) / 64;          // 312 limbs hold degrees 0..19967 (phi itself has degree 19937)
using Limbs = std::vector<uint64_t>;

inline bool bit_of(const Limbs& a, int i) { return (a[i >> 6] >> (i & 63)) & 1u; }

// a ^= b << s  (b has nb limbs)
void xor_shifted(Limbs& a, const uint64_t* b, int nb, int s) {
    const int wo = s >> 6, bs = s & 63;
    if (bs == 0) {
        for (int k = 0; k < nb; ++k) a[wo + k] ^= b[k];
    } else {
        for (int k = 0; k < nb; ++k) {
            a[wo + k] ^= b[k] << bs;
            a[wo + k + 1] ^= b[k] >> (64 - bs);
        }
    }
}

// reduce a (2*PW + 1 limbs) modulo phi (degree MT_DEG); the result sits in the low PW limbs
void reduce(Limbs& a, const Limbs& phi) {
    for (int d = 2 * PW * 64 - 1; d >= MT_DEG; --d)
        if (bit_of(a, d)) xor_shifted(a, phi.data(), PW, d - MT_DEG);
}

Limbs mulmod(const Limbs& a, const Limbs& b, const Limbs& phi) {
    Limbs acc(2 * PW + 1, 0);
    for (int i = 0; i < MT_DEG; ++i)
        if (bit_of(a, i)) xor_shifted(acc, b.data(), PW, i);
    reduce(acc, phi);
    acc.resize(PW);
    return acc;
}

Limbs sqrmod(const Limbs& a, const Limbs& phi) {
    Limbs acc(2 * PW + 1, 0);
    for (int i = 0; i < MT_DEG; ++i)
        if (bit_of(a, i)) acc[(2 * i) >> 6] |= 1ull << ((2 * i) & 63);
    reduce(acc, phi);
    acc.resize(PW);
    return acc;
}

// characteristic polynomial of the MT19937 transition: Berlekamp-Massey on bit 0 of the twister's own word stream
Limbs mt_char_poly() {
    const int NB = 2 * MT_DEG + 2;
    std::vector<uint32_t> x(NB + 625);
    x[0] = 5489u;
    for (int i = 1; i < 624; ++i) x[i] = 1812433253u * (x[i - 1] ^ (x[i - 1] >> 30)) + (uint32_t)i;
    for (size_t n = 624; n < x.size(); ++n) x[n] = x[n - 227] ^ mt_mix(x[n - 624], x[n - 623]);
    const int LW = PW + 1;
    Limbs Cp(LW, 0), Bp(LW, 0), R(LW, 0), T;
    Cp[0] = Bp[0] = 1;
    int L = 0, m = 1;
    for (int N = 0; N < NB; ++N) {
        // R bit i = s[N - i]
        for (int k = LW - 1; k > 0; --k) R[k] = (R[k] << 1) | (R[k - 1] >> 63);
        R[0] = (R[0] << 1) | (uint64_t)(x[1 + N] & 1u);
        uint64_t acc = 0;
        for (int k = 0; k < LW; ++k) acc ^= Cp[k] & R[k];
        if (__builtin_popcountll(acc) & 1) {
            const bool grow = 2 * L <= N;
            if (grow) T = Cp;
            {   // Cp ^= Bp << m  (degrees stay <= L_new <= MT_DEG for this sequence)
                const int wo = m >> 6, bs = m & 63;
                for (int k = 0; k + wo < LW; ++k) {
                    Cp[k + wo] ^= Bp[k] << bs;
                    if (bs && k + wo + 1 < LW) Cp[k + wo + 1] ^= Bp[k] >> (64 - bs);
                }
            }
            if (grow) { L = N + 1 - L; Bp = T; m = 1; } else ++m;
        } else {
            ++m;
        }
    }
    Limbs phi(PW, 0);
    if (L != MT_DEG) return phi;                     // cannot happen; leaves phi = 0 and the caller reports it
    for (int k = 0; k <= L; ++k)                     // phi_k = c_{L-k}
        if ((Cp[(L - k) >> 6] >> ((L - k) & 63)) & 1u) phi[k >> 6] |= 1ull << (k & 63);
    return phi;
}

// x^e mod phi
Limbs powx(uint64_t e, const Limbs& phi) {
    Limbs r(PW, 0);
    r[0] = 1;
    for (int b = 63; b >= 0; --b) {
        r = sqrmod(r, phi);
        if ((e >> b) & 1u) {
            Limbs sh(2 * PW + 1, 0);
            xor_shifted(sh, r.data(), PW, 1);
            reduce(sh, phi);
            sh.resize(PW);
            r = sh;
        }
    }
    return r;
}

struct JumpTables {
    bool ok = false;
    std::vector<uint16_t> idx;              // header + lists, exactly as the kernel reads it
};

const JumpTables& jump_tables() {
    static JumpTables jt;
    static std::once_flag once;
    std::call_once(once, [] {
        const Limbs phi = mt_char_poly();
        if (!bit_of(phi, MT_DEG)) return;
        std::vector<Limbs> poly((size_t)MT_NPOLY);
        const Limbs g1 = powx((uint64_t)MT_SEG_WORDS, phi);
        // g_{k J} = g_{(k-1) J} . g_J : eight host threads, each starting its range of k from a square-and-multiply power
        const int n_thr = 8, span = (255 + n_thr - 1) / n_thr;
        std::vector<std::thread> th;
        for (int i = 0; i < n_thr; ++i)
            th.emplace_back([&, i] {
                const int k0 = i * span + 1, k1 = std::min(255, (i + 1) * span);
                if (k0 > k1) return;
                Limbs cur = k0 == 1 ? g1 : powx((uint64_t)MT_SEG_WORDS * (uint64_t)k0, phi);
                for (int k = k0; k <= k1; ++k) {
                    poly[(size_t)k - 1] = cur;
                    if (k < k1) cur = mulmod(cur, g1, phi);
                }
            });
        for (auto& t : th) t.join();
        poly[255] = mulmod(poly[254], g1, phi);                  // g_{256 J}
        poly[256] = sqrmod(poly[255], phi);                      // g_{512 J}
        poly[257] = mulmod(poly[256], poly[255], phi);           // g_{768 J}
        std::vector<int32_t> off((size_t)MT_NPOLY), cnt((size_t)MT_NPOLY);
        std::vector<uint16_t> lists;
        for (int k = 0; k < MT_NPOLY; ++k) {
            off[k] = (int32_t)lists.size();
            for (int i = 0; i < MT_DEG; ++i)
                if (bit_of(poly[k], i)) lists.push_back((uint16_t)i);
            while ((lists.size() - off[k]) % 16) lists.push_back((uint16_t)MT_WIN);   // points at the zero padding
            cnt[k] = (int32_t)lists.size() - off[k];
        }
        jt.idx.resize((size_t)MT_HDR + lists.size());
        memcpy(jt.idx.data(), off.data(), sizeof(int32_t) * MT_NPOLY);
        memcpy(jt.idx.data() + 2 * MT_NPOLY, cnt.data(), sizeof(int32_t) * MT_NPOLY);
        memcpy(jt.idx.data() + MT_HDR, lists.data(), sizeof(uint16_t) * lists.size());
        jt.ok = true;
    });
    return jt;
}
}  // namespace

// x^(n_words) mod phi as 624 little-endian 32-bit words (bit i of the polynomial = bit (i % 32) of word i / 32);
// n_words = 0 returns phi itself without its leading term.  For tests.
int mt19937_jump_poly(int64_t n_words, uint32_t* out624) {
    static Limbs phi;
    static std::once_flag once;
    std::call_once(once, [] { phi = mt_char_poly(); });
    if (!bit_of(phi, MT_DEG)) return -1;
    Limbs g = phi;
    if (n_words > 0) g = powx((uint64_t)n_words, phi);
    else g[MT_DEG >> 6] &= ~(1ull << (MT_DEG & 63));
    for (int k = 0; k < 624; ++k) out624[k] = (uint32_t)(g[k >> 1] >> (32 * (k & 1)));
    return 0;
}

size_t mt19937_jump_table_bytes() {
    const JumpTables& jt = jump_tables();
    return jt.ok ? jt.idx.size() * sizeof(uint16_t) : 0;
}

// n draws; state_dev: 2 x 625 words of device scratch (key + position, ping-pong), with the handed state in half
// slot_in; on return *final_slot tells which half holds the final state.  jump_idx_dev: the device copy of the
// set-bit lists (mt19937_jump_tables_host), only needed when the draw spans more than one segment.
cudaError_t launch_mt19937_uniform(uint32_t* state_dev, int slot_in, int32_t pos_host, int* final_slot,
                                   const uint16_t* jump_idx_dev, uint32_t* raw_dev, int64_t n, double low, double high,
                                   void* out, int out_dtype, cudaStream_t st) {
    *final_slot = slot_in;
    if (n < 1) return cudaSuccess;
    static bool attr_done = false;
    const size_t smem = sizeof(uint32_t) * (MT_WIN_ALLOC + 4 * 768) + sizeof(uint16_t) * MT_IDX_STAGE;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_mt19937_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    MtJumpArgs ja{};
    ja.idx = jump_idx_dev;
    int64_t left = 2 * n;                      // 32-bit words still to draw
    int64_t pos = pos_host;
    uint32_t* o = raw_dev;
    int slot = slot_in;
    const int64_t per_launch = (int64_t)MT_MAX_SEGS * MT_SEG_WORDS;
    while (left > 0) {
        // words this launch consumes: everything, or exactly up to the end of its last segment
        int64_t take = left;
        if (pos + take > per_launch) take = per_launch - pos;
        const int64_t last = pos + take - 1;
        const int64_t need = (last / 624 + 1) * 624;
        const int segs = (int)((need + MT_SEG_WORDS - 1) / MT_SEG_WORDS);
        if (segs > 1 && !jump_idx_dev) return cudaErrorInvalidValue;
        uint32_t* kin = state_dev + 625 * slot;
        uint32_t* kout = state_dev + 625 * (slot ^ 1);
        k_mt19937_stream<<<segs, MT_THREADS, smem, st>>>(kin, (int32_t)pos, kout, (int32_t*)(kout + 624), o, take, ja);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        slot ^= 1;
        o += take;
        left -= take;
        pos = 624;                              // a launch that stops at its capacity leaves (last block, 624)
    }
    *final_slot = slot;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (out_dtype == MPB_F64) k_mt_to_uniform<double><<<grid, 256, 0, st>>>(raw_dev, n, low, high - low, (double*)out);
    else k_mt_to_uniform<float><<<grid, 256, 0, st>>>(raw_dev, n, low, high - low, (float*)out);
    return cudaGetLastError();
}

bool mt19937_needs_jump(int32_t pos, int64_t n) {
    return pos + 2 * n > MT_SEG_WORDS;
}

const uint16_t* mt19937_jump_table_host(size_t* n_entries) {
    const JumpTables& jt = jump_tables();
    *n_entries = jt.ok ? jt.idx.size() : 0;
    return jt.ok ? jt.idx.data() : nullptr;
}

}  // namespace mpb

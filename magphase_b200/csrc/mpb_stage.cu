// Host -> device staging of float64 signals for the *_host entry points.
//
// The reference hands float64 NumPy arrays in pageable memory (sf.read of PCM16 wavs, src/magphase.py:2878).  A plain
// cudaMemcpy of those runs at the driver's pageable rate (~10 GB/s) and moves 8 bytes per sample; it was the largest
// single item of the end-to-end time.  Here a small pool of host threads narrows the samples to float32 into a
// page-locked staging buffer, chunk by chunk, while the chunks already done are in flight over PCIe.  The narrowing
// is only used when it is EXACT for every sample ((double)(float)x == x, true for anything read from 8/16/24-bit
// PCM); otherwise the signal is uploaded as float64, as before.  No arithmetic happens on the host.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <thread>

#include "mpb_ctx.h"

namespace mpb {

namespace {

class HostPool {
public:
    static HostPool& get() {
        static HostPool p;
        return p;
    }
    int size() const { return (int)th_.size() + 1; }
    // runs fn(chunk) for chunk in [0, n) on the pool threads AND the caller; returns when all chunks are done
    void run(int n, const std::function<void(int)>& fn, const std::function<void(int)>& on_done_in_order) {
        std::lock_guard<std::mutex> one_caller(run_mu_);       // contexts of different devices share the pool
        std::unique_lock<std::mutex> lk(mu_);
        fn_ = &fn;
        n_ = n;
        next_.store(0);
        done_.assign((size_t)n, 0);
        ++gen_;
        active_ = (int)th_.size();
        lk.unlock();
        cv_.notify_all();
        // the caller issues the in-order completion callbacks (the H2D copies) and helps when nothing is ready
        int issued = 0;
        while (issued < n) {
            if (__atomic_load_n(&done_[issued], __ATOMIC_ACQUIRE)) { on_done_in_order(issued); ++issued; continue; }
            const int c = next_.fetch_add(1);
            if (c < n) { fn(c); __atomic_store_n(&done_[c], (char)1, __ATOMIC_RELEASE); }
            else std::this_thread::yield();
        }
        lk.lock();
        idle_cv_.wait(lk, [&] { return active_ == 0; });
        fn_ = nullptr;
    }

private:
    HostPool() {
        // default: 8 threads, or this process's share of the cores when a launcher runs one process per GPU
        int n = 8;
        const int hc = (int)std::thread::hardware_concurrency();
        if (const char* w = getenv("LOCAL_WORLD_SIZE")) {
            const int lw = atoi(w);
            if (lw > 1 && hc > 0 && hc / lw < n) n = hc / lw;
        }
        if (const char* e = getenv("MPB_HOST_THREADS")) n = atoi(e);
        if (hc > 0 && n > hc) n = hc;
        if (n < 1) n = 1;
        for (int i = 0; i < n - 1; ++i) th_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            const std::function<void(int)>* fn = fn_;
            const int n = n_;
            lk.unlock();
            for (;;) {
                const int c = next_.fetch_add(1);
                if (c >= n) break;
                (*fn)(c);
                __atomic_store_n(&done_[c], (char)1, __ATOMIC_RELEASE);
            }
            lk.lock();
            if (--active_ == 0) idle_cv_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_, idle_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, active_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
    std::atomic<int> next_{0};
    std::vector<char> done_;
};

// narrow src[0..n) to float32; returns false as soon as a sample does not survive the round trip
inline bool narrow_exact(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
    int bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float f = (float)src[i];
        dst[i] = f;
        bad |= ((double)f != src[i]);
    }
    return bad == 0;
}

}  // namespace

int host_threads() { return HostPool::get().size(); }

// Uploads the virtual concatenation sigs[0] | sigs[1] | ... (HOST float64) on stream st, GROUP by group (group_end[g] =
// one past the last signal of group g; groups are contiguous and ordered).  A group whose samples are all exactly
// representable in float32 is narrowed on the host into the page-locked staging buffer and lands in dev_f32 (absolute
// sample offsets); any other group is uploaded as float64 into dev_f64.  As soon as the last copy of a group has been
// enqueued, on_group(g, dtype) runs on the calling thread -- the caller records its event there and enqueues the
// group's kernels while the pool is still narrowing the next group.  The copies are asynchronous and the staging
// buffer is reused by the next call on this ctx: the caller synchronises before returning to its own caller.
int upload_signal_groups(mpb_ctx* ctx, cudaStream_t st, const double* const* sigs, const int64_t* lens, int32_t n_sigs,
                         const int32_t* group_end, int32_t n_groups, void* dev_f32, void* dev_f64,
                         const std::function<int(int32_t, int)>& on_group) {
    int64_t n_sig = 0;
    for (int32_t i = 0; i < n_sigs; ++i) n_sig += lens[i];
    static const bool trace = getenv("MPB_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    PinnedBuf& pb = ctx->stage;
    const bool have_stage = n_sig > 0 && pb.need(sizeof(float) * (size_t)n_sig) == cudaSuccess;
    if (!have_stage) cudaGetLastError();      // no page-locked memory to be had: plain float64 path for every group
    // chunks of <= CH samples that never straddle two utterances
    constexpr int64_t CH = 1 << 18;
    struct Chunk { const double* src; int64_t off, n; int32_t group; bool last_of_group; };
    std::vector<Chunk> ch;
    std::vector<int64_t> sig_off((size_t)n_sigs + 1, 0);
    {
        int32_t g = 0;
        for (int32_t i = 0; i < n_sigs; ++i) {
            while (g < n_groups - 1 && i >= group_end[g]) ++g;
            sig_off[i + 1] = sig_off[i] + lens[i];
            for (int64_t a = 0; a < lens[i] || (a == 0 && lens[i] == 0); a += CH) {
                const int64_t n = lens[i] - a < CH ? lens[i] - a : CH;
                ch.push_back({sigs[i] + a, sig_off[i] + a, n, g, false});
                if (lens[i] == 0) break;
            }
        }
        for (size_t c = 0; c < ch.size(); ++c)
            ch[c].last_of_group = c + 1 == ch.size() || ch[c + 1].group != ch[c].group;
    }
    float* h = (float*)pb.p;
    std::vector<std::atomic<int>> inexact((size_t)n_groups);
    for (auto& x : inexact) x.store(have_stage ? 0 : 1);
    int rc = MPB_OK;
    size_t c_first = 0;                          // first chunk of the group currently being issued
    HostPool::get().run(
        (int)ch.size(),
        [&](int c) {
            const Chunk& k = ch[c];
            if (inexact[k.group].load(std::memory_order_relaxed)) return;
            if (!narrow_exact(k.src, h + k.off, k.n)) inexact[k.group].store(1);
        },
        [&](int c) {
            if (rc != MPB_OK) return;
            const Chunk& k = ch[c];
            cudaError_t e = cudaSuccess;
            // optimistic: a narrowed chunk goes out at once (its group may still turn out inexact; then the float64
            // copies below supersede it -- in dev_f64, or later in stream order when both are the same buffer)
            if (k.n > 0 && !inexact[k.group].load())
                e = cudaMemcpyAsync((float*)dev_f32 + k.off, h + k.off, sizeof(float) * k.n, cudaMemcpyHostToDevice, st);
            if (k.last_of_group && e == cudaSuccess) {
                const bool f32 = !inexact[k.group].load();
                if (!f32)
                    for (size_t j = c_first; j <= (size_t)c && e == cudaSuccess; ++j)
                        if (ch[j].n > 0)
                            e = cudaMemcpyAsync((double*)dev_f64 + ch[j].off, ch[j].src, sizeof(double) * ch[j].n,
                                                cudaMemcpyHostToDevice, st);
                c_first = (size_t)c + 1;
                if (e == cudaSuccess) { rc = on_group(k.group, f32 ? MPB_F32 : MPB_F64); return; }
            }
            if (e != cudaSuccess) rc = fail(MPB_ERR_CUDA, std::string("signal upload: ") + cudaGetErrorString(e));
        });
    if (trace)
        fprintf(stderr, "[mpb] upload_signal_groups: %lld samples, %d chunks, %d groups, %d threads, %.3f ms\n",
                (long long)n_sig, (int)ch.size(), (int)n_groups, HostPool::get().size(),
                1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count());
    return rc;
}

__global__ void k_i16_to_f32(const int16_t* __restrict__ a, float* __restrict__ b, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (float)a[i] * (1.0f / 32768.0f);          // sf.read of PCM16 (src/libaudio.py:343-350): exact in float32
}

// Signals that are already narrow on the host -- float32 samples or PCM16 (int16, scaled by 1/32768 on the device like
// sf.read does) -- skip the float64 narrowing pass: the pool only copies them into the page-locked staging buffer, chunk by
// chunk, while earlier chunks are on PCIe.  int16 groups land in dev_aux and are converted into dev_f32 on `st`.
int upload_signal_groups_narrow(mpb_ctx* ctx, cudaStream_t st, const void* const* sigs, int sig_dtype, const int64_t* lens,
                                int32_t n_sigs, const int32_t* group_end, int32_t n_groups, void* dev_f32, void* dev_aux,
                                const std::function<int(int32_t, int)>& on_group) {
    const size_t es = sig_dtype == MPB_I16 ? 2 : 4;
    int64_t n_sig = 0;
    for (int32_t i = 0; i < n_sigs; ++i) n_sig += lens[i];
    PinnedBuf& pb = ctx->stage;
    if (n_sig > 0 && pb.need(es * (size_t)n_sig) != cudaSuccess) return fail(MPB_ERR_CUDA, "page-locked staging buffer");
    constexpr int64_t CH = 1 << 19;
    struct Chunk { const char* src; int64_t off, n; int32_t group; bool last_of_group; };
    std::vector<Chunk> ch;
    {
        int32_t g = 0;
        int64_t off = 0;
        for (int32_t i = 0; i < n_sigs; ++i) {
            while (g < n_groups - 1 && i >= group_end[g]) ++g;
            for (int64_t a = 0; a < lens[i] || (a == 0 && lens[i] == 0); a += CH) {
                const int64_t n = lens[i] - a < CH ? lens[i] - a : CH;
                ch.push_back({(const char*)sigs[i] + es * a, off + a, n, g, false});
                if (lens[i] == 0) break;
            }
            off += lens[i];
        }
        for (size_t c = 0; c < ch.size(); ++c) ch[c].last_of_group = c + 1 == ch.size() || ch[c + 1].group != ch[c].group;
    }
    char* h = (char*)pb.p;
    int rc = MPB_OK;
    int64_t group_first = 0;                     // first sample of the group currently being issued
    HostPool::get().run(
        (int)ch.size(),
        [&](int c) { if (ch[c].n > 0) memcpy(h + es * ch[c].off, ch[c].src, es * (size_t)ch[c].n); },
        [&](int c) {
            if (rc != MPB_OK) return;
            const Chunk& k = ch[c];
            cudaError_t e = cudaSuccess;
            char* dst = sig_dtype == MPB_I16 ? (char*)dev_aux : (char*)dev_f32;
            if (k.n > 0) e = cudaMemcpyAsync(dst + es * k.off, h + es * k.off, es * (size_t)k.n, cudaMemcpyHostToDevice, st);
            if (k.last_of_group && e == cudaSuccess) {
                const int64_t g_end = k.off + k.n, g_n = g_end - group_first;
                if (sig_dtype == MPB_I16 && g_n > 0) {
                    k_i16_to_f32<<<(unsigned)((g_n + 255) / 256), 256, 0, st>>>((const int16_t*)dev_aux + group_first,
                                                                             (float*)dev_f32 + group_first, g_n);
                    e = cudaGetLastError();
                    ctx->launches += 1;
                }
                group_first = g_end;
                if (e == cudaSuccess) { rc = on_group(k.group, MPB_F32); return; }
            }
            if (e != cudaSuccess) rc = fail(MPB_ERR_CUDA, std::string("signal upload: ") + cudaGetErrorString(e));
        });
    return rc;
}

bool host_is_page_locked(const void* p) {
    cudaPointerAttributes a;
    const cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int h2d_staged(mpb_ctx* ctx, cudaStream_t st, void* dev, const void* host, size_t bytes, PinnedBuf& stage, size_t stage_off) {
    (void)ctx;
    if (bytes == 0) return MPB_OK;
    if (host_is_page_locked(host) || stage_off + bytes > stage.cap) {
        const cudaError_t e = cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st);
        return e == cudaSuccess ? MPB_OK : fail(MPB_ERR_CUDA, std::string("feature upload: ") + cudaGetErrorString(e));
    }
    constexpr size_t CH = (size_t)2 << 20;
    const int n = (int)((bytes + CH - 1) / CH);
    char* h = (char*)stage.p + stage_off;
    int rc = MPB_OK;
    HostPool::get().run(
        n,
        [&](int c) { const size_t a = (size_t)c * CH; memcpy(h + a, (const char*)host + a, bytes - a < CH ? bytes - a : CH); },
        [&](int c) {
            if (rc != MPB_OK) return;
            const size_t a = (size_t)c * CH;
            const cudaError_t e = cudaMemcpyAsync((char*)dev + a, h + a, bytes - a < CH ? bytes - a : CH, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) rc = fail(MPB_ERR_CUDA, std::string("feature upload: ") + cudaGetErrorString(e));
        });
    return rc;
}

int h2d_gather_staged(mpb_ctx* ctx, cudaStream_t st, void* dev, const void* const* blocks, const int64_t* blk_rows, int32_t b0,
                      int32_t b1, size_t row_bytes, PinnedBuf& stage, size_t stage_off) {
    (void)ctx;
    constexpr size_t CH = (size_t)2 << 20;
    struct Chunk { const char* src; size_t off, n; };
    std::vector<Chunk> ch;
    size_t total = 0;
    for (int32_t b = b0; b < b1; ++b) {
        const size_t bytes = (size_t)blk_rows[b] * row_bytes;
        if (bytes && !blocks[b]) return fail(MPB_ERR_BAD_ARG, "NULL feature block");
        for (size_t a = 0; a < bytes; a += CH) ch.push_back({(const char*)blocks[b] + a, total + a, bytes - a < CH ? bytes - a : CH});
        total += bytes;
    }
    if (total == 0) return MPB_OK;
    if (stage_off + total > stage.cap) return fail(MPB_ERR_INTERNAL, "feature staging buffer too small");
    char* h = (char*)stage.p + stage_off;
    int rc = MPB_OK;
    HostPool::get().run(
        (int)ch.size(),
        [&](int c) { memcpy(h + ch[c].off, ch[c].src, ch[c].n); },
        [&](int c) {
            if (rc != MPB_OK) return;
            const cudaError_t e = cudaMemcpyAsync((char*)dev + ch[c].off, h + ch[c].off, ch[c].n, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) rc = fail(MPB_ERR_CUDA, std::string("feature upload: ") + cudaGetErrorString(e));
        });
    return rc;
}

// One group: everything in dev (>= 8 bytes per sample); *out_dtype says how it was uploaded.
int upload_signals(mpb_ctx* ctx, cudaStream_t st, const double* const* sigs, const int64_t* lens, int32_t n_sigs,
                   void* dev, int* out_dtype) {
    *out_dtype = MPB_F64;
    const int32_t end = n_sigs;
    return upload_signal_groups(ctx, st, sigs, lens, n_sigs, &end, 1, dev, dev,
                                [&](int32_t, int dtype) { *out_dtype = dtype; return MPB_OK; });
}

}  // namespace mpb

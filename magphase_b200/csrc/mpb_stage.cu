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
        int n = 8;
        if (const char* e = getenv("MPB_HOST_THREADS")) n = atoi(e);
        const int hc = (int)std::thread::hardware_concurrency();
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
    std::mutex mu_;
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

// Uploads the virtual concatenation sigs[0] | sigs[1] | ... (HOST float64) into dev (>= 8 * n_sig bytes) on stream st.
// *out_dtype = MPB_F32 when every sample is exactly representable in float32 (narrowed on the host, half the PCIe
// bytes), else MPB_F64.  The copies are asynchronous; the staging buffer is reused by the next call on this ctx, so the
// caller synchronises st before returning to its own caller (all *_host entry points do).
int upload_signals(mpb_ctx* ctx, cudaStream_t st, const double* const* sigs, const int64_t* lens, int32_t n_sigs,
                   void* dev, int* out_dtype) {
    int64_t n_sig = 0;
    for (int32_t i = 0; i < n_sigs; ++i) n_sig += lens[i];
    *out_dtype = MPB_F64;
    if (n_sig == 0) return MPB_OK;
    PinnedBuf& pb = ctx->stage;
    bool exact = false;
    static const bool trace = getenv("MPB_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    if (pb.need(sizeof(float) * (size_t)n_sig) == cudaSuccess) {
        // chunks of <= CH samples that never straddle two utterances
        constexpr int64_t CH = 1 << 18;
        struct Chunk { const double* src; int64_t off, n; };
        std::vector<Chunk> ch;
        int64_t off = 0;
        for (int32_t i = 0; i < n_sigs; ++i) {
            for (int64_t a = 0; a < lens[i]; a += CH) {
                const int64_t n = lens[i] - a < CH ? lens[i] - a : CH;
                ch.push_back({sigs[i] + a, off + a, n});
            }
            off += lens[i];
        }
        float* h = (float*)pb.p;
        std::atomic<int> inexact{0};
        cudaError_t cerr = cudaSuccess;
        HostPool::get().run(
            (int)ch.size(),
            [&](int c) {
                if (inexact.load(std::memory_order_relaxed)) return;
                if (!narrow_exact(ch[c].src, h + ch[c].off, ch[c].n)) inexact.store(1);
            },
            [&](int c) {
                if (inexact.load() || cerr != cudaSuccess) return;
                cerr = cudaMemcpyAsync((float*)dev + ch[c].off, h + ch[c].off, sizeof(float) * ch[c].n,
                                       cudaMemcpyHostToDevice, st);
            });
        if (cerr != cudaSuccess) return fail(MPB_ERR_CUDA, std::string("signal upload: ") + cudaGetErrorString(cerr));
        exact = !inexact.load();
        if (trace)
            fprintf(stderr, "[mpb] upload_signals: %lld samples, %d chunks, %d threads, narrow+issue %.3f ms, exact=%d\n",
                    (long long)n_sig, (int)ch.size(), HostPool::get().size(),
                    1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count(), (int)exact);
    } else {
        cudaGetLastError();      // no page-locked memory to be had: fall through to the plain path
    }
    if (exact) { *out_dtype = MPB_F32; return MPB_OK; }
    // float64 upload; stream order puts it after any float32 chunk already issued into the same buffer
    int64_t off = 0;
    for (int32_t i = 0; i < n_sigs; ++i) {
        CU(cudaMemcpyAsync((double*)dev + off, sigs[i], sizeof(double) * lens[i], cudaMemcpyHostToDevice, st));
        off += lens[i];
    }
    return MPB_OK;
}

}  // namespace mpb

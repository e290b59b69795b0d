// Shared internals of the C-ABI translation units: context, scratch buffers, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "mpb_kernels.h"

namespace mpb {
int fail(int code, const std::string& msg);
}
using mpb::fail;

#define CU(call)                                                                               \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail(MPB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));     \
    } while (0)

struct DevBuf {   // grow-only device scratch
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinnedBuf {   // grow-only page-locked host staging buffer
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct KernelTimer {   // optional per-kernel CUDA-event timing (mpb_profile_begin / mpb_profile_end)
    struct Rec { const char* name; cudaEvent_t a, b; };
    bool on = false;
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
};

struct mpb_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;                 // used by the *_host entry points (compute)
    cudaStream_t stream_in = nullptr, stream_out = nullptr;   // host->device / device->host copies of the pipelined entry points
    cudaStream_t stream_aux = nullptr;             // the MT19937 noise stream runs here, beside the compute stream
    std::vector<cudaEvent_t> ev_pool;              // timing-disabled events (pipeline hand-offs)
    std::vector<cudaEvent_t> ev_block;             // blocking-sync events of host_wait()
    std::mutex ev_mu;
    PinnedBuf desc_stage;                          // page-locked staging of descriptor arrays
    PinnedBuf mt_fin;                              // page-locked landing zone of the MT19937 state read-back
    std::map<int, void*> tw32, tw64;               // fft_len -> twiddle table exp(-2 pi i j / N), j < N/2
    std::mutex mu;                                 // serialises the *_host entry points (shared scratch)
    std::mutex tw_mu;
    int64_t launches = 0;
    DevBuf scratch[16];
    DevBuf mt_jump;                                // MT19937 jump polynomials (set-bit lists), uploaded on first use
    bool mt_jump_ready = false;
    DevBuf ticket;                                 // run hand-out counter of k_synthesis_lossless (one launch at a time per ctx stream)
    PinnedBuf stage;                               // float32 staging of host signals (upload_signals)
    PinnedBuf stage_feat;                          // staging of pageable feature matrices (h2d_staged)
    KernelTimer timer;
};

// timing-disabled events for the stage hand-offs of the pipelined *_host entry points (callers hold ctx->mu)
inline cudaEvent_t get_event(mpb_ctx* ctx) {
    if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}
inline void put_event(mpb_ctx* ctx, cudaEvent_t e) { if (e) ctx->ev_pool.push_back(e); }

// Host-side wait for a stream.  cudaStreamSynchronize spins (the runtime's default when there are more cores than contexts):
// a worker thread that drains a 5 ms pipeline burns a core the whole time, and with one process per GPU and a few cores per
// process the staging threads then starve (measured at 8 ranks x 4 cores: end-to-end throughput per rank fell to a quarter).
// MPB_SYNC=block (default when a launcher runs several processes per host, LOCAL_WORLD_SIZE > 1): record an event created
// with cudaEventBlockingSync and sleep on it.  MPB_SYNC=spin forces the runtime's behaviour.
inline bool sync_blocks() {
    static const int mode = [] {
        if (const char* e = getenv("MPB_SYNC")) return (e[0] == 'b' || e[0] == 'B') ? 1 : 0;
        if (const char* w = getenv("LOCAL_WORLD_SIZE")) return atoi(w) > 1 ? 1 : 0;
        return 0;
    }();
    return mode == 1;
}
inline cudaError_t host_wait(mpb_ctx* ctx, cudaStream_t st) {
    if (!sync_blocks()) return cudaStreamSynchronize(st);
    cudaEvent_t e = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->ev_mu);
        if (!ctx->ev_block.empty()) { e = ctx->ev_block.back(); ctx->ev_block.pop_back(); }
    }
    if (!e) {
        cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync);
        if (r != cudaSuccess) return r;
    }
    cudaError_t r = cudaEventRecord(e, st);
    if (r == cudaSuccess) r = cudaEventSynchronize(e);
    std::lock_guard<std::mutex> lk(ctx->ev_mu);
    ctx->ev_block.push_back(e);
    return r;
}

// Scope guard of the pipelined *_host entry points: whatever path leaves the function (CU() returns on any CUDA error), the
// four streams are drained before host / staging memory that live copies still read goes out of scope, and the hand-off
// events go back to the pool.
struct PipelineDrain {
    mpb_ctx* ctx;
    std::vector<cudaEvent_t>* evs[2] = {nullptr, nullptr};
    explicit PipelineDrain(mpb_ctx* c, std::vector<cudaEvent_t>* a = nullptr, std::vector<cudaEvent_t>* b = nullptr) : ctx(c) { evs[0] = a; evs[1] = b; }
    ~PipelineDrain() {
        host_wait(ctx, ctx->stream_in); host_wait(ctx, ctx->stream); host_wait(ctx, ctx->stream_out);
        host_wait(ctx, ctx->stream_aux);
        for (auto* v : evs)
            if (v) { for (auto e : *v) put_event(ctx, e); v->clear(); }
    }
};

// number of utterance groups the pipelined *_host entry points cut a batch of nfrm frames into: about one group per
// `frames_per_group` frames, at most `max_groups` (MPB_PIPELINE_GROUPS overrides).  Smaller groups shorten the fill and
// drain of the pipeline, larger ones keep the persistent kernels' grids full.
inline int pipeline_groups(int64_t nfrm, int64_t frames_per_group, int max_groups) {
    static const int forced = [] {
        const char* e = getenv("MPB_PIPELINE_GROUPS");
        const int v = e ? atoi(e) : 0;
        return v < 0 ? 0 : (v > 64 ? 64 : v);
    }();
    if (forced) return forced;
    const int64_t n = (nfrm + frames_per_group / 2) / frames_per_group;
    return (int)(n < 1 ? 1 : (n > max_groups ? max_groups : n));
}

// Launch `expr` (returns cudaError_t) on stream `st`, bracketed by events when profiling is on.
#define LAUNCH(ctx, st, kname, expr)                                                           \
    do {                                                                                       \
        cudaEvent_t _ea = nullptr, _eb = nullptr;                                              \
        if ((ctx)->timer.on) { _ea = (ctx)->timer.get(); _eb = (ctx)->timer.get(); cudaEventRecord(_ea, (st)); } \
        CU(expr);                                                                              \
        if ((ctx)->timer.on) { cudaEventRecord(_eb, (st)); (ctx)->timer.recs.push_back({kname, _ea, _eb}); } \
        (ctx)->launches += 1;                                                                  \
    } while (0)

namespace mpb {
inline bool fft_len_ok(int n) { return n == 1024 || n == 2048 || n == 4096; }
inline bool dtype_ok(int d) { return d == MPB_F32 || d == MPB_F64; }
inline bool sig_dtype_ok(int d) { return d == MPB_F32 || d == MPB_F64 || d == MPB_I16; }
int get_twiddles(mpb_ctx* ctx, int fft_len, int dtype, const void** out);
int analysis_common(mpb_ctx* ctx, void* stream, const void* sig, int sig_dtype, int64_t n_sig,
                    const int64_t* centre, const int32_t* left, const int32_t* right, const uint8_t* win,
                    int64_t nfrm, int fft_len, int compute_dtype, void* out_a, void* out_b, void* out_c,
                    int out_dtype, int mode);
// HOST float64 signals sigs[0] | sigs[1] | ... -> dev (>= 8 bytes per sample) on st; *out_dtype = MPB_F32 when the
// samples could be narrowed exactly on the host (mpb_stage.cu), else MPB_F64.  Caller synchronises st before returning.
int upload_signals(mpb_ctx* ctx, cudaStream_t st, const double* const* sigs, const int64_t* lens, int32_t n_sigs,
                   void* dev, int* out_dtype);
// Same, group by group (see mpb_stage.cu): float32 groups land in dev_f32, others in dev_f64, both at absolute sample
// offsets; on_group(g, dtype) runs on the calling thread right after the group's last copy has been enqueued.
int upload_signal_groups(mpb_ctx* ctx, cudaStream_t st, const double* const* sigs, const int64_t* lens, int32_t n_sigs,
                         const int32_t* group_end, int32_t n_groups, void* dev_f32, void* dev_f64,
                         const std::function<int(int32_t, int)>& on_group);
int upload_signal_groups_narrow(mpb_ctx* ctx, cudaStream_t st, const void* const* sigs, int sig_dtype, const int64_t* lens,
                                int32_t n_sigs, const int32_t* group_end, int32_t n_groups, void* dev_f32, void* dev_aux,
                                const std::function<int(int32_t, int)>& on_group);
int host_threads();

// Host -> device copy of a block that may be pageable, on stream st.  Page-locked sources (cudaHostAlloc / registered, e.g. the
// arrays the batch analysis functions return) go straight to the copy engine.  Pageable ones -- arrays read from feature
// files -- would make cudaMemcpyAsync bounce them through the driver's small staging buffer on the CALLING thread (measured:
// 11 ms instead of 3 ms per 128-utterance batch); the host thread pool copies them into `stage` (page-locked, at byte offset
// stage_off, capacity ensured by the caller) in chunks, and every chunk's DMA is enqueued as soon as it is staged.
int h2d_staged(mpb_ctx* ctx, cudaStream_t st, void* dev, const void* host, size_t bytes, PinnedBuf& stage, size_t stage_off);
bool host_is_page_locked(const void* p);
// The same for a matrix whose rows arrive as separate host blocks (one per utterance: blocks[b] holds blk_rows[b] rows of
// row_bytes bytes): blocks b0 .. b1-1 are copied back to back into `stage` at stage_off by the pool and land at dev, chunk by
// chunk.  Replaces a np.concatenate on the caller's side (one pass over the data instead of three).
int h2d_gather_staged(mpb_ctx* ctx, cudaStream_t st, void* dev, const void* const* blocks, const int64_t* blk_rows, int32_t b0,
                      int32_t b1, size_t row_bytes, PinnedBuf& stage, size_t stage_off);
int mt19937_enqueue(mpb_ctx* ctx, cudaStream_t st, const uint32_t* key, int32_t pos, const int64_t* part_n, int n_parts,
                    double low, double high, void* out_dev, int out_dtype, uint32_t* fin625, cudaEvent_t* part_done);
int check_frames_host(const int64_t* centre, const int32_t* left, const int32_t* right, int64_t nfrm,
                      int64_t n_sig, int fft_len);
}

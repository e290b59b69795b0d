// C-ABI layer of libmagphase_b200.so: context, tables, argument checks, host staging.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "mpb_ctx.h"

using namespace mpb;

static thread_local std::string g_err;

namespace mpb {
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
}  // namespace mpb

int mpb::get_twiddles(mpb_ctx* ctx, int fft_len, int dtype, const void** out) {
    std::lock_guard<std::mutex> lk(ctx->tw_mu);
    auto& tab = dtype == MPB_F64 ? ctx->tw64 : ctx->tw32;
    auto it = tab.find(fft_len);
    if (it != tab.end()) { *out = it->second; return MPB_OK; }
    const int n = fft_len / 2;
    std::vector<double> h64(2 * n);
    for (int j = 0; j < n; ++j) {
        const long double a = 2.0L * 3.141592653589793238462643383279502884L * (long double)j / (long double)fft_len;
        h64[2 * j] = (double)cosl(a);
        h64[2 * j + 1] = (double)(-sinl(a));
    }
    void* d = nullptr;
    if (dtype == MPB_F64) {
        CU(cudaMalloc(&d, sizeof(double) * 2 * n));
        CU(cudaMemcpy(d, h64.data(), sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> h32(2 * n);
        for (int j = 0; j < 2 * n; ++j) h32[j] = (float)h64[j];
        CU(cudaMalloc(&d, sizeof(float) * 2 * n));
        CU(cudaMemcpy(d, h32.data(), sizeof(float) * 2 * n, cudaMemcpyHostToDevice));
    }
    tab[fft_len] = d;
    *out = d;
    return MPB_OK;
}

extern "C" {

const char* mpb_last_error(void) { return g_err.c_str(); }
const char* mpb_version(void) { return "magphase_b200 0.1 (sm_100a)"; }

int mpb_create(int device, mpb_ctx** out_ctx) {
    if (!out_ctx) return fail(MPB_ERR_BAD_ARG, "out_ctx is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(MPB_ERR_NO_DEVICE, std::string("no CUDA device available (there is no CPU fallback): ") +
                                           cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(MPB_ERR_BAD_ARG, "device ordinal out of range");
    CU(cudaSetDevice(device));
    mpb_ctx* c = new mpb_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->stream_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->stream_out, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->stream_aux, cudaStreamNonBlocking));
    *out_ctx = c;
    return MPB_OK;
}

int mpb_destroy(mpb_ctx* ctx) {
    if (!ctx) return MPB_OK;
    cudaSetDevice(ctx->device);
    for (auto& kv : ctx->tw32) cudaFree(kv.second);
    for (auto& kv : ctx->tw64) cudaFree(kv.second);
    for (auto& b : ctx->scratch) b.release();
    ctx->mt_jump.release();
    ctx->ticket.release();
    ctx->stage.release();
    ctx->stage_feat.release();
    ctx->desc_stage.release();
    ctx->mt_fin.release();
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->stream_in) cudaStreamDestroy(ctx->stream_in);
    if (ctx->stream_out) cudaStreamDestroy(ctx->stream_out);
    if (ctx->stream_aux) cudaStreamDestroy(ctx->stream_aux);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MPB_OK;
}

int64_t mpb_launch_count(const mpb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mpb_profile_begin(mpb_ctx* ctx) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    for (auto& r : ctx->timer.recs) { ctx->timer.pool.push_back(r.a); ctx->timer.pool.push_back(r.b); }
    ctx->timer.recs.clear();
    ctx->timer.on = true;
    return MPB_OK;
}

int mpb_profile_end(mpb_ctx* ctx, char* buf, int64_t buf_len) {
    if (!ctx || !buf || buf_len < 1) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    ctx->timer.on = false;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& r : ctx->timer.recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { auto& e = agg[r.name]; e.first += 1; e.second += ms; }
        ctx->timer.pool.push_back(r.a); ctx->timer.pool.push_back(r.b);
    }
    ctx->timer.recs.clear();
    std::string out;
    char line[256];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if ((int64_t)out.size() + 1 > buf_len) return fail(MPB_ERR_BAD_ARG, "profile buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return MPB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
int mpb::analysis_common(mpb_ctx* ctx, void* stream, const void* sig, int sig_dtype, int64_t n_sig,
                           const int64_t* centre, const int32_t* left, const int32_t* right, const uint8_t* win,
                           int64_t nfrm, int fft_len, int compute_dtype, void* out_a, void* out_b, void* out_c,
                           int out_dtype, int mode) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (nfrm < 0 || n_sig < 0) return fail(MPB_ERR_BAD_ARG, "negative size");
    if (!fft_len_ok(fft_len)) return fail(MPB_ERR_FFT_LEN, "fft_len must be 1024, 2048 or 4096");
    if (!dtype_ok(sig_dtype) || !dtype_ok(compute_dtype) || !dtype_ok(out_dtype))
        return fail(MPB_ERR_BAD_ARG, "unknown dtype");
    if (nfrm == 0) return MPB_OK;
    if (!sig || !centre || !left || !right || !out_a || (mode == MODE_FEATS && (!out_b || !out_c)))
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    AnalysisArgs a;
    a.sig = sig; a.sig_dtype = sig_dtype; a.n_sig = n_sig;
    a.centre = centre; a.left = left; a.right = right; a.win = win;
    a.nfrm = nfrm; a.fft_len = fft_len; a.compute_dtype = compute_dtype;
    int rc = get_twiddles(ctx, fft_len, compute_dtype, &a.tw);
    if (rc != MPB_OK) return rc;
    a.out_a = out_a; a.out_b = out_b; a.out_c = out_c; a.out_dtype = out_dtype;
    a.mode = mode; a.num_sms = ctx->num_sms;
    LAUNCH(ctx, (cudaStream_t)stream, mode == MODE_FFT ? "k_analysis<fft>" : "k_analysis", launch_analysis(a, (cudaStream_t)stream));
    return MPB_OK;
}

extern "C" {

int mpb_analysis_lossless_dev(mpb_ctx* ctx, void* stream, const void* sig, int sig_dtype, int64_t n_sig,
                              const int64_t* centre, const int32_t* left, const int32_t* right, const uint8_t* win,
                              int64_t nfrm, int fft_len, int compute_dtype, void* out_mag, void* out_real,
                              void* out_imag, int out_dtype) {
    return analysis_common(ctx, stream, sig, sig_dtype, n_sig, centre, left, right, win, nfrm, fft_len, compute_dtype,
                           out_mag, out_real, out_imag, out_dtype, MODE_FEATS);
}

int mpb_frames_fft_dev(mpb_ctx* ctx, void* stream, const void* sig, int sig_dtype, int64_t n_sig,
                       const int64_t* centre, const int32_t* left, const int32_t* right, const uint8_t* win,
                       int64_t nfrm, int fft_len, int compute_dtype, void* out_fft, int out_dtype) {
    return analysis_common(ctx, stream, sig, sig_dtype, n_sig, centre, left, right, win, nfrm, fft_len, compute_dtype,
                           out_fft, nullptr, nullptr, out_dtype, MODE_FFT);
}

}  // extern "C"

// Host-side geometry check shared by the *_host entry points (the *_dev ones trust the caller's plan).
int mpb::check_frames_host(const int64_t* centre, const int32_t* left, const int32_t* right, int64_t nfrm,
                             int64_t n_sig, int fft_len) {
    for (int64_t f = 0; f < nfrm; ++f) {
        if (left[f] < 0 || right[f] < 0) return fail(MPB_ERR_FRAME_GEOM, "negative frame side length");
        if (centre[f] - left[f] < 0 || centre[f] + right[f] >= n_sig)
            return fail(MPB_ERR_FRAME_GEOM, "frame reaches outside the signal");
    }
    return MPB_OK;
}

extern "C" {

// sig_dtype: MPB_F64 (narrowed on the host when that is exact), MPB_F32 or MPB_I16 (PCM16, scaled by 1/32768 on the device);
// out_dtype: element type of the host output matrices
static int frames_host(mpb_ctx* ctx, const void* sig, int sig_in_dtype, int64_t n_sig, const int64_t* centre, const int32_t* left,
                       const int32_t* right, const uint8_t* win, int64_t nfrm, int fft_len, int compute_dtype,
                       void* o0, void* o1, void* o2, int out_dtype, int mode) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (nfrm == 0) return MPB_OK;
    if (!fft_len_ok(fft_len)) return fail(MPB_ERR_FFT_LEN, "fft_len must be 1024, 2048 or 4096");
    if (!sig || !centre || !left || !right || !o0) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    int rc = check_frames_host(centre, left, right, nfrm, n_sig, fft_len);
    if (rc != MPB_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int64_t H = fft_len / 2 + 1;
    if (!sig_dtype_ok(sig_in_dtype) || !dtype_ok(out_dtype)) return fail(MPB_ERR_BAD_ARG, "unknown dtype");
    const size_t osz = (out_dtype == MPB_F64 ? 8 : 4) * (size_t)nfrm * H * (mode == MODE_FFT ? 2 : 1);
    DevBuf* b = ctx->scratch;
    CU(b[0].need(sizeof(double) * n_sig));
    CU(b[1].need(sizeof(int64_t) * nfrm));
    CU(b[2].need(sizeof(int32_t) * nfrm));
    CU(b[3].need(sizeof(int32_t) * nfrm));
    CU(b[4].need(win ? (size_t)nfrm : 1));
    CU(b[5].need(osz));
    if (mode == MODE_FEATS) { CU(b[6].need(osz)); CU(b[7].need(osz)); }
    cudaStream_t st = ctx->stream;
    int sig_dtype = MPB_F64;
    if (sig_in_dtype == MPB_F64) {
        const double* sd = (const double*)sig;
        rc = upload_signals(ctx, st, &sd, &n_sig, 1, b[0].p, &sig_dtype);
    } else {
        // narrow element types: staged as they are; int16 lands in the upper half of the buffer and is converted in front
        const int32_t end = 1;
        sig_dtype = MPB_F32;
        rc = upload_signal_groups_narrow(ctx, st, &sig, sig_in_dtype, &n_sig, 1, &end, 1, b[0].p, (char*)b[0].p + 4 * (size_t)n_sig,
                                         [](int32_t, int) { return (int)MPB_OK; });
    }
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpyAsync(b[1].p, centre, sizeof(int64_t) * nfrm, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[2].p, left, sizeof(int32_t) * nfrm, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[3].p, right, sizeof(int32_t) * nfrm, cudaMemcpyHostToDevice, st));
    if (win) CU(cudaMemcpyAsync(b[4].p, win, (size_t)nfrm, cudaMemcpyHostToDevice, st));
    rc = analysis_common(ctx, st, b[0].p, sig_dtype, n_sig, (const int64_t*)b[1].p, (const int32_t*)b[2].p,
                         (const int32_t*)b[3].p, win ? (const uint8_t*)b[4].p : nullptr, nfrm, fft_len, compute_dtype,
                         b[5].p, b[6].p, b[7].p, out_dtype, mode);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpyAsync(o0, b[5].p, osz, cudaMemcpyDeviceToHost, st));
    if (mode == MODE_FEATS) {
        CU(cudaMemcpyAsync(o1, b[6].p, osz, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(o2, b[7].p, osz, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return MPB_OK;
}

int mpb_analysis_lossless_host(mpb_ctx* ctx, const double* sig, int64_t n_sig, const int64_t* centre,
                               const int32_t* left, const int32_t* right, const uint8_t* win, int64_t nfrm,
                               int fft_len, int compute_dtype, double* out_mag, double* out_real, double* out_imag) {
    if (nfrm > 0 && (!out_real || !out_imag)) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    return frames_host(ctx, sig, MPB_F64, n_sig, centre, left, right, win, nfrm, fft_len, compute_dtype, out_mag, out_real,
                       out_imag, MPB_F64, MODE_FEATS);
}

// mpb_analysis_lossless_host with the caller's element types: signal float64 / float32 / int16 PCM (sig_dtype), feature
// matrices float64 or float32 (out_dtype; the reference's .mag/.real/.imag files are float32, src/libutils.py:122-127).
int mpb_analysis_lossless_host2(mpb_ctx* ctx, const void* sig, int sig_dtype, int64_t n_sig, const int64_t* centre,
                                const int32_t* left, const int32_t* right, const uint8_t* win, int64_t nfrm, int fft_len,
                                int compute_dtype, void* out_mag, void* out_real, void* out_imag, int out_dtype) {
    if (nfrm > 0 && (!out_real || !out_imag)) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    return frames_host(ctx, sig, sig_dtype, n_sig, centre, left, right, win, nfrm, fft_len, compute_dtype, out_mag, out_real,
                       out_imag, out_dtype, MODE_FEATS);
}

int mpb_frames_fft_host(mpb_ctx* ctx, const double* sig, int64_t n_sig, const int64_t* centre, const int32_t* left,
                        const int32_t* right, const uint8_t* win, int64_t nfrm, int fft_len, int compute_dtype,
                        double* out_fft) {
    return frames_host(ctx, sig, MPB_F64, n_sig, centre, left, right, win, nfrm, fft_len, compute_dtype, out_fft, nullptr,
                       nullptr, MPB_F64, MODE_FFT);
}

// ---------------------------------------------------------------------------------------------
int mpb_plan_ola_runs(const int32_t* pm, const int64_t* utt_frm_off, int32_t n_utt, int fft_len,
                      int32_t target_frames, int32_t* out_runs, int64_t capacity, int64_t* n_runs) {
    if (!pm || !utt_frm_off || !n_runs || n_utt < 0) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    if (!fft_len_ok(fft_len)) return fail(MPB_ERR_FFT_LEN, "fft_len must be 1024, 2048 or 4096");
    if (target_frames < 1) target_frames = 32;
    int64_t nr = 0;
    for (int32_t u = 0; u < n_utt; ++u) {
        const int64_t a = utt_frm_off[u], b = utt_frm_off[u + 1];
        if (b < a) return fail(MPB_ERR_BAD_ARG, "utt_frm_off not non-decreasing");
        for (int64_t f = a + 1; f < b; ++f)
            if (pm[f] <= pm[f - 1]) return fail(MPB_ERR_FRAME_GEOM, "pitch marks must be strictly increasing");
        int64_t s = a;
        const int64_t first_run = nr;
        while (s < b) {
            int64_t e = s + target_frames < b ? s + target_frames : b;
            while (e < b && pm[e - 1] - pm[s] < fft_len) ++e;             // every run spans >= fft_len samples
            if (e < b && pm[b - 1] - pm[e] < fft_len) e = b;              // ... including the last one
            if (out_runs) {
                if (nr >= capacity) return fail(MPB_ERR_BAD_ARG, "run buffer too small");
                out_runs[4 * nr + 0] = (int32_t)s;
                out_runs[4 * nr + 1] = (int32_t)(e - s);
                out_runs[4 * nr + 2] = u;
                out_runs[4 * nr + 3] = (nr > first_run ? 1 : 0) | (e < b ? 2 : 0);
            }
            ++nr;
            s = e;
        }
    }
    *n_runs = nr;
    return MPB_OK;
}

int mpb_synthesis_lossless_dev(mpb_ctx* ctx, void* stream, const void* mag, const void* real, const void* imag,
                               int feat_dtype, const int32_t* pm, int64_t nfrm_total, const int64_t* utt_out_off,
                               const int32_t* utt_t0, int32_t n_utt, const int32_t* runs, int32_t n_runs, int fft_len,
                               int compute_dtype, void* out, int out_dtype, int64_t n_out) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (!fft_len_ok(fft_len)) return fail(MPB_ERR_FFT_LEN, "fft_len must be 1024, 2048 or 4096");
    if (!dtype_ok(feat_dtype) || !dtype_ok(compute_dtype) || !dtype_ok(out_dtype))
        return fail(MPB_ERR_BAD_ARG, "unknown dtype");
    if (nfrm_total < 0 || n_out < 0 || n_utt < 0 || n_runs < 0) return fail(MPB_ERR_BAD_ARG, "negative size");
    if (n_out == 0) return MPB_OK;
    if (!out) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    if (nfrm_total > 0 && (!mag || !real || !imag || !pm || !utt_out_off || !utt_t0 || !runs))
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    SynthArgs a;
    a.mag = mag; a.real = real; a.imag = imag; a.feat_dtype = feat_dtype;
    a.pm = pm; a.nfrm_total = nfrm_total;
    a.utt_frm_off = nullptr; a.utt_out_off = utt_out_off; a.utt_t0 = utt_t0; a.n_utt = n_utt;
    a.runs = (const OlaRun*)runs; a.n_runs = n_runs;
    a.fft_len = fft_len; a.compute_dtype = compute_dtype;
    int rc = get_twiddles(ctx, fft_len, compute_dtype, &a.tw);
    if (rc != MPB_OK) return rc;
    a.out = out; a.out_dtype = out_dtype; a.n_out = n_out; a.num_sms = ctx->num_sms;
    CU(ctx->ticket.need(sizeof(int)));
    a.run_ticket = (int*)ctx->ticket.p;
    LAUNCH(ctx, (cudaStream_t)stream, "k_synthesis_lossless", launch_synthesis_lossless(a, (cudaStream_t)stream));
    return MPB_OK;
}

int mpb_synthesis_lossless_host(mpb_ctx* ctx, const double* mag, const double* real, const double* imag,
                                const int32_t* pm, int64_t nfrm_total, const int64_t* utt_frm_off,
                                const int64_t* utt_out_off, const int32_t* utt_t0, int32_t n_utt, int fft_len,
                                int compute_dtype, double* out, int64_t n_out) {
    return mpb_synthesis_lossless_host2(ctx, mag, real, imag, MPB_F64, pm, nfrm_total, utt_frm_off, utt_out_off, utt_t0, n_utt,
                                        fft_len, compute_dtype, out, MPB_F64, n_out);
}

// mpb_synthesis_lossless_host with float64 or float32 feature matrices (feat_dtype) and waveform (out_dtype).
int mpb_synthesis_lossless_host2(mpb_ctx* ctx, const void* mag, const void* real, const void* imag, int feat_dtype,
                                 const int32_t* pm, int64_t nfrm_total, const int64_t* utt_frm_off,
                                 const int64_t* utt_out_off, const int32_t* utt_t0, int32_t n_utt, int fft_len,
                                 int compute_dtype, void* out, int out_dtype, int64_t n_out) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (!dtype_ok(feat_dtype) || !dtype_ok(out_dtype)) return fail(MPB_ERR_BAD_ARG, "unknown dtype");
    const size_t fes = feat_dtype == MPB_F64 ? 8 : 4, oes = out_dtype == MPB_F64 ? 8 : 4;
    if (!fft_len_ok(fft_len)) return fail(MPB_ERR_FFT_LEN, "fft_len must be 1024, 2048 or 4096");
    if (n_out == 0) return MPB_OK;
    if (!mag || !real || !imag || !pm || !utt_frm_off || !utt_out_off || !utt_t0 || !out)
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    int64_t n_runs = 0;
    int target = 32;
    {   // fewer frames per run when the batch is too small to fill the GPU
        const int64_t want = (int64_t)ctx->num_sms * 4;
        if (nfrm_total / target < want) target = (int)(nfrm_total / want > 1 ? nfrm_total / want : 1);
    }
    int rc = mpb_plan_ola_runs(pm, utt_frm_off, n_utt, fft_len, target, nullptr, 0, &n_runs);
    if (rc != MPB_OK) return rc;
    std::vector<int32_t> runs(4 * (size_t)(n_runs > 0 ? n_runs : 1));
    rc = mpb_plan_ola_runs(pm, utt_frm_off, n_utt, fft_len, target, runs.data(), n_runs, &n_runs);
    if (rc != MPB_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    std::unique_lock<std::mutex> lk(ctx->mu);
    const int64_t H = fft_len / 2 + 1;
    const size_t fsz = fes * (size_t)nfrm_total * H;
    DevBuf* b = ctx->scratch;
    CU(b[5].need(fsz)); CU(b[6].need(fsz)); CU(b[7].need(fsz));
    CU(b[1].need(sizeof(int32_t) * nfrm_total));
    CU(b[2].need(sizeof(int64_t) * (n_utt + 1)));
    CU(b[3].need(sizeof(int32_t) * (n_utt + 1)));
    CU(b[8].need(sizeof(int32_t) * 4 * (size_t)n_runs));
    CU(b[0].need(oes * n_out));
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(b[5].p, mag, fsz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[6].p, real, fsz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[7].p, imag, fsz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[1].p, pm, sizeof(int32_t) * nfrm_total, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[2].p, utt_out_off, sizeof(int64_t) * (n_utt + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[3].p, utt_t0, sizeof(int32_t) * n_utt, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[8].p, runs.data(), sizeof(int32_t) * 4 * (size_t)n_runs, cudaMemcpyHostToDevice, st));
    rc = mpb_synthesis_lossless_dev(ctx, st, b[5].p, b[6].p, b[7].p, feat_dtype, (const int32_t*)b[1].p, nfrm_total,
                                    (const int64_t*)b[2].p, (const int32_t*)b[3].p, n_utt, (const int32_t*)b[8].p,
                                    (int32_t)n_runs, fft_len, compute_dtype, b[0].p, out_dtype, n_out);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpyAsync(out, b[0].p, oes * n_out, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MPB_OK;
}

// ---------------------------------------------------------------------------------------------
int mpb_post_filter_dev(mpb_ctx* ctx, void* stream, const void* x, int dtype, int64_t nfrm, int dim,
                        const int32_t* centre, const int32_t* half, const double* tilt, void* out) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (!dtype_ok(dtype) || nfrm < 0) return fail(MPB_ERR_BAD_ARG, "bad dtype or size");
    if (dim < 2 || dim > 256) return fail(MPB_ERR_DIM, "post-filter dimension must be in 2..256");
    if (nfrm == 0) return MPB_OK;
    if (!x || !centre || !half || !tilt || !out) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    LAUNCH(ctx, (cudaStream_t)stream, "k_post_filter",
           launch_post_filter(x, dtype, nfrm, dim, centre, half, tilt, out, (cudaStream_t)stream));
    return MPB_OK;
}

// r0 of SPTK `freqt -A 0 | c2acr -M 0` for nfrm cepstra of n coefficients (post_filter_merlin, src/magphase.py:3419-3427).
// G: HOST float64 [n][K], K = L/2 + 1: the all-pass transform folded into the cosine table by the host mirror.
int mpb_cep_energy_host(mpb_ctx* ctx, const double* c, int64_t nfrm, int n, const double* G, int K, int L, double* r0) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (nfrm == 0) return MPB_OK;
    if (!c || !G || !r0 || n < 1 || n > 1024 || K < 2 || L != 2 * (K - 1)) return fail(MPB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevBuf* b = ctx->scratch;
    cudaStream_t st = ctx->stream;
    CU(b[0].need(sizeof(double) * (size_t)nfrm * n)); CU(b[1].need(sizeof(double) * (size_t)n * K)); CU(b[5].need(sizeof(double) * (size_t)nfrm));
    CU(cudaMemcpyAsync(b[0].p, c, sizeof(double) * (size_t)nfrm * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[1].p, G, sizeof(double) * (size_t)n * K, cudaMemcpyHostToDevice, st));
    LAUNCH(ctx, st, "k_cep_energy", launch_cep_energy((const double*)b[0].p, nfrm, n, (const double*)b[1].p, K, L, (double*)b[5].p,
                                                      ctx->num_sms, st));
    CU(cudaMemcpyAsync(r0, b[5].p, sizeof(double) * (size_t)nfrm, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MPB_OK;
}

int mpb_lossless_feats_host(mpb_ctx* ctx, const double* fft, int64_t n, double* mag, double* real, double* imag) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (n < 0) return fail(MPB_ERR_BAD_ARG, "bad size");
    if (n == 0) return MPB_OK;
    if (!fft || !mag || !real || !imag) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevBuf* b = ctx->scratch;
    cudaStream_t st = ctx->stream;
    // chunks of at most 16 M values: 256 MB in, 384 MB out of scratch whatever the size of the call
    const int64_t CH = (int64_t)1 << 24;
    const int64_t c0 = n < CH ? n : CH;
    CU(b[0].need(sizeof(double) * 2 * (size_t)c0)); CU(b[5].need(sizeof(double) * 3 * (size_t)c0));
    for (int64_t a = 0; a < n; a += CH) {
        const int64_t c = n - a < CH ? n - a : CH;
        double* o = (double*)b[5].p;
        CU(cudaMemcpyAsync(b[0].p, fft + 2 * a, sizeof(double) * 2 * (size_t)c, cudaMemcpyHostToDevice, st));
        LAUNCH(ctx, st, "k_lossless_feats", launch_lossless_feats(b[0].p, c, o, o + c, o + 2 * c, ctx->num_sms, st));
        CU(cudaMemcpyAsync(mag + a, o, sizeof(double) * (size_t)c, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(real + a, o + c, sizeof(double) * (size_t)c, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(imag + a, o + 2 * c, sizeof(double) * (size_t)c, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return MPB_OK;
}

int mpb_ola_dev(mpb_ctx* ctx, void* stream, const double* frames, const int32_t* pm, int64_t nfrm, int frmlen,
                int32_t t0, double* out, int64_t n_out) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (nfrm < 0 || n_out < 0 || frmlen < 1) return fail(MPB_ERR_BAD_ARG, "bad size");
    if (n_out == 0) return MPB_OK;
    if (!out || (nfrm > 0 && (!frames || !pm))) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    LAUNCH(ctx, (cudaStream_t)stream, "k_ola_gather",
           launch_ola_gather(frames, pm, nfrm, frmlen, t0, out, n_out, (cudaStream_t)stream));
    return MPB_OK;
}

int mpb_ola_host(mpb_ctx* ctx, const double* frames, const int32_t* pm, int64_t nfrm, int frmlen,
                 int32_t t0, double* out, int64_t n_out) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (nfrm < 0 || n_out < 0 || frmlen < 1) return fail(MPB_ERR_BAD_ARG, "bad size");
    if (n_out == 0) return MPB_OK;
    if (!out || (nfrm > 0 && (!frames || !pm))) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    for (int64_t i = 1; i < nfrm; ++i)
        if (pm[i] < pm[i - 1]) return fail(MPB_ERR_BAD_ARG, "pitch marks must be non-decreasing");
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevBuf* b = ctx->scratch;
    cudaStream_t st = ctx->stream;
    const size_t fsz = sizeof(double) * (size_t)nfrm * (size_t)frmlen, osz = sizeof(double) * (size_t)n_out;
    CU(b[0].need(fsz ? fsz : 8)); CU(b[2].need(nfrm ? sizeof(int32_t) * (size_t)nfrm : 4)); CU(b[5].need(osz));
    if (nfrm > 0) {
        CU(cudaMemcpyAsync(b[0].p, frames, fsz, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(b[2].p, pm, sizeof(int32_t) * (size_t)nfrm, cudaMemcpyHostToDevice, st));
    }
    int rc = mpb_ola_dev(ctx, st, (const double*)b[0].p, (const int32_t*)b[2].p, nfrm, frmlen, t0, (double*)b[5].p, n_out);
    if (rc != MPB_OK) { cudaStreamSynchronize(st); return rc; }
    CU(cudaMemcpyAsync(out, b[5].p, osz, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MPB_OK;
}

int mpb_post_filter_host(mpb_ctx* ctx, const double* x, int64_t nfrm, int dim, const int32_t* centre,
                         const int32_t* half, const double* tilt, double* out) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (nfrm == 0) return MPB_OK;
    if (!x || !centre || !half || !tilt || !out || dim < 2) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    for (int b = 0; b < dim; ++b)
        if (half[b] < 0 || centre[b] - half[b] < 0 || centre[b] + half[b] >= dim)
            return fail(MPB_ERR_BAD_ARG, "post-filter averaging window reaches outside the feature vector");
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevBuf* b = ctx->scratch;
    cudaStream_t st = ctx->stream;
    const size_t sz = sizeof(double) * (size_t)nfrm * dim;
    CU(b[0].need(sz)); CU(b[5].need(sz));
    CU(b[2].need(sizeof(int32_t) * dim)); CU(b[3].need(sizeof(int32_t) * dim)); CU(b[1].need(sizeof(double) * dim));
    CU(cudaMemcpyAsync(b[0].p, x, sz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[2].p, centre, sizeof(int32_t) * dim, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[3].p, half, sizeof(int32_t) * dim, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[1].p, tilt, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
    int rc = mpb_post_filter_dev(ctx, st, b[0].p, MPB_F64, nfrm, dim, (const int32_t*)b[2].p, (const int32_t*)b[3].p,
                                 (const double*)b[1].p, b[5].p);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpyAsync(out, b[5].p, sz, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MPB_OK;
}

int mpb_min_phase_dev(mpb_ctx* ctx, void* stream, const void* mag, int dtype, int64_t nfrm, int fft_len, void* out_cplx) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (!fft_len_ok(fft_len)) return fail(MPB_ERR_FFT_LEN, "fft_len must be 1024, 2048 or 4096");
    if (!dtype_ok(dtype) || nfrm < 0) return fail(MPB_ERR_BAD_ARG, "bad dtype or size");
    if (nfrm == 0) return MPB_OK;
    if (!mag || !out_cplx) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    const void* tw = nullptr;
    int rc = get_twiddles(ctx, fft_len, MPB_F64, &tw);
    if (rc != MPB_OK) return rc;
    LAUNCH(ctx, (cudaStream_t)stream, "k_min_phase",
           launch_min_phase(fft_len, mag, dtype, nfrm, tw, out_cplx, ctx->num_sms, (cudaStream_t)stream));
    return MPB_OK;
}

int mpb_min_phase_host(mpb_ctx* ctx, const double* mag, int64_t nfrm, int fft_len, double* out_cplx) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (!fft_len_ok(fft_len)) return fail(MPB_ERR_FFT_LEN, "fft_len must be 1024, 2048 or 4096");
    if (nfrm == 0) return MPB_OK;
    if (!mag || !out_cplx) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    DevBuf* b = ctx->scratch;
    cudaStream_t st = ctx->stream;
    const size_t sz = sizeof(double) * (size_t)nfrm * (fft_len / 2 + 1);
    CU(b[5].need(sz)); CU(b[6].need(2 * sz));
    CU(cudaMemcpyAsync(b[5].p, mag, sz, cudaMemcpyHostToDevice, st));
    int rc = mpb_min_phase_dev(ctx, st, b[5].p, MPB_F64, nfrm, fft_len, b[6].p);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpyAsync(out_cplx, b[6].p, 2 * sz, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MPB_OK;
}

// ---------------------------------------------------------------------------------------------
}  // extern "C"

// Enqueues sum(part_n) draws of np.random.uniform(low, high) from the state (key[624], pos) on `st`, part by part
// (part p fills out_dev[sum(part_n[0..p)) ...]; part_done[p], when given, is recorded after it), and an asynchronous
// read-back of the final state (624 key words + position) into fin625, which must be PAGE-LOCKED host memory.  Nothing
// is synchronised: the position after each part follows from the counts alone.  Building block of the pipelined
// synthesis entry point.
int mpb::mt19937_enqueue(mpb_ctx* ctx, cudaStream_t st, const uint32_t* key, int32_t pos, const int64_t* part_n, int n_parts,
                         double low, double high, void* out_dev, int out_dtype, uint32_t* fin625, cudaEvent_t* part_done) {
    int64_t n = 0;
    for (int p = 0; p < n_parts; ++p) n += part_n[p];
    DevBuf* b = ctx->scratch;
    CU(b[9].need(sizeof(uint32_t) * 2 * 625));
    CU(b[10].need(sizeof(uint32_t) * 2 * (size_t)(n > 0 ? n : 1)));
    const uint16_t* d_jump = nullptr;
    if (mt19937_needs_jump(pos, n)) {             // more than one segment: jump polynomials (built once per process)
        if (!ctx->mt_jump_ready) {
            size_t n_idx = 0;
            const uint16_t* h = mt19937_jump_table_host(&n_idx);
            if (!h) return fail(MPB_ERR_INTERNAL, "MT19937 jump polynomials could not be derived");
            CU(ctx->mt_jump.need(sizeof(uint16_t) * n_idx));
            CU(cudaMemcpyAsync(ctx->mt_jump.p, h, sizeof(uint16_t) * n_idx, cudaMemcpyHostToDevice, st));
            ctx->mt_jump_ready = true;
        }
        d_jump = (const uint16_t*)ctx->mt_jump.p;
    }
    uint32_t* d_state = (uint32_t*)b[9].p;
    CU(cudaMemcpyAsync(d_state, key, sizeof(uint32_t) * 624, cudaMemcpyHostToDevice, st));
    int slot = 0;
    int64_t done = 0;
    const size_t oes = out_dtype == MPB_F64 ? 8 : 4;
    for (int p = 0; p < n_parts; ++p) {
        if (part_n[p] > 0) {
            LAUNCH(ctx, st, "k_mt19937_stream+k_mt_to_uniform",
                   launch_mt19937_uniform(d_state, slot, pos, &slot, d_jump, (uint32_t*)b[10].p + 2 * done, part_n[p], low,
                                          high, (char*)out_dev + oes * done, out_dtype, st));
            ctx->launches += 1;
            const int64_t last = (int64_t)pos + 2 * part_n[p] - 1;           // stream index of the last consumed word
            pos = (int32_t)(last + 1 - (last / 624) * 624);
            done += part_n[p];
        }
        if (part_done) CU(cudaEventRecord(part_done[p], st));
    }
    if (n > 0) {
        CU(cudaMemcpyAsync(fin625, d_state + 625 * slot, sizeof(uint32_t) * 625, cudaMemcpyDeviceToHost, st));
    } else {
        memcpy(fin625, key, sizeof(uint32_t) * 624);
        fin625[624] = (uint32_t)pos;
    }
    return MPB_OK;
}

extern "C" {

// n draws of np.random.uniform(low, high) from NumPy's legacy MT19937 state (key[624], pos in [0, 624]) into a
// DEVICE buffer; key/pos (HOST, in/out) are advanced exactly as NumPy would advance them.  Synchronises `stream`.
int mpb_mt19937_uniform_dev(mpb_ctx* ctx, void* stream, uint32_t* key, int32_t* pos, int64_t n, double low, double high,
                            void* out_dev, int out_dtype) {
    if (!ctx || !key || !pos) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (!dtype_ok(out_dtype) || n < 0 || *pos < 0 || *pos > 624) return fail(MPB_ERR_BAD_ARG, "bad dtype, size or MT19937 position");
    if (n == 0) return MPB_OK;
    if (!out_dev) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CU(ctx->mt_fin.need(sizeof(uint32_t) * 625));
    uint32_t* fin = (uint32_t*)ctx->mt_fin.p;
    int rc = mt19937_enqueue(ctx, st, key, *pos, &n, 1, low, high, out_dev, out_dtype, fin, nullptr);
    if (rc != MPB_OK) return rc;
    CU(cudaStreamSynchronize(st));
    memcpy(key, fin, sizeof(uint32_t) * 624);
    *pos = (int32_t)fin[624];
    return MPB_OK;
}

// Same draw, enqueue only: nothing is synchronised and the caller's state (key, pos: HOST, read before returning) is NOT
// advanced -- for device-resident pipelines that re-draw the same stretch of the stream every step (bench.py times the
// draw inside its step this way).
int mpb_mt19937_fill_dev(mpb_ctx* ctx, void* stream, const uint32_t* key, int32_t pos, int64_t n, double low, double high,
                         void* out_dev, int out_dtype) {
    if (!ctx || !key) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (!dtype_ok(out_dtype) || n < 0 || pos < 0 || pos > 624) return fail(MPB_ERR_BAD_ARG, "bad dtype, size or MT19937 position");
    if (n == 0) return MPB_OK;
    if (!out_dev) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    CU(ctx->mt_fin.need(sizeof(uint32_t) * 625));
    return mt19937_enqueue(ctx, (cudaStream_t)stream, key, pos, &n, 1, low, high, out_dev, out_dtype, (uint32_t*)ctx->mt_fin.p, nullptr);
}

// x^(n_words) mod phi, phi = characteristic polynomial of the MT19937 transition (n_words = 0: phi minus its leading
// term), as 624 little-endian 32-bit words.  Host only; exposes the jump-ahead arithmetic to the tests.
int mpb_mt19937_jump_poly(int64_t n_words, uint32_t* out624) {
    if (!out624 || n_words < 0) return fail(MPB_ERR_BAD_ARG, "bad argument");
    if (mt19937_jump_poly(n_words, out624) != 0) return fail(MPB_ERR_INTERNAL, "Berlekamp-Massey did not return degree 19937");
    return MPB_OK;
}

int mpb_mt19937_uniform_host(mpb_ctx* ctx, uint32_t* key, int32_t* pos, int64_t n, double low, double high, double* out) {
    if (!ctx || !out) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (n == 0) return MPB_OK;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx->scratch[11].need(sizeof(double) * (size_t)n));
    int rc = mpb_mt19937_uniform_dev(ctx, ctx->stream, key, pos, n, low, high, ctx->scratch[11].p, MPB_F64);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpy(out, ctx->scratch[11].p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return MPB_OK;
}

// ---------------------------------------------------------------------------------------------
// Two cascaded biquads (scipy sos layout: 2 x [b0 b1 b2 1 a1 a2], HOST) in place on a DEVICE buffer holding n_utt
// concatenated utterances (utt_off: HOST array [n_utt+1]).
int mpb_sos2_dev(mpb_ctx* ctx, void* stream, void* x, int dtype, const int64_t* utt_off, int32_t n_utt, const double* sos) {
    if (!ctx || !utt_off || !sos) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (!dtype_ok(dtype) || n_utt < 0) return fail(MPB_ERR_BAD_ARG, "bad dtype or size");
    if (sos[3] != 1.0 || sos[9] != 1.0) return fail(MPB_ERR_BAD_ARG, "sos sections must be normalised (a0 == 1)");
    if (n_utt == 0 || utt_off[n_utt] == utt_off[0]) return MPB_OK;
    if (!x) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int L = 512;
    std::vector<int64_t> h(2 * ((size_t)n_utt + 1));
    int64_t* uo = h.data();
    int64_t* co = h.data() + n_utt + 1;
    co[0] = 0;
    for (int32_t u = 0; u <= n_utt; ++u) uo[u] = utt_off[u];
    for (int32_t u = 0; u < n_utt; ++u) {
        if (uo[u + 1] < uo[u]) return fail(MPB_ERR_BAD_ARG, "utt_off not non-decreasing");
        co[u + 1] = co[u] + (uo[u + 1] - uo[u] + L - 1) / L;
    }
    const int64_t n_chunks = co[n_utt];
    // zero-input state map of the cascade (one step on each unit state) and its L-th power by repeated squaring
    double M[16], P[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, T[16];
    for (int j = 0; j < 4; ++j) {
        double z[4] = {0, 0, 0, 0};
        z[j] = 1.0;
        const double y1 = z[0];
        double n0 = z[1] - sos[4] * y1, n1 = -sos[5] * y1;
        const double y2 = sos[6] * y1 + z[2];
        double n2 = sos[7] * y1 + z[3] - sos[10] * y2, n3 = sos[8] * y1 - sos[11] * y2;
        M[0 * 4 + j] = n0; M[1 * 4 + j] = n1; M[2 * 4 + j] = n2; M[3 * 4 + j] = n3;
    }
    auto mul = [&](const double* A, const double* B, double* C) {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double acc = 0;
                for (int k = 0; k < 4; ++k) acc += A[4 * i + k] * B[4 * k + j];
                C[4 * i + j] = acc;
            }
    };
    for (int e = L; e > 0; e >>= 1) {
        if (e & 1) { mul(P, M, T); memcpy(P, T, sizeof(P)); }
        mul(M, M, T); memcpy(M, T, sizeof(M));
    }
    DevBuf* sb = ctx->scratch;
    CU(sb[12].need(sizeof(int64_t) * h.size()));
    CU(sb[13].need(sizeof(double) * 4 * (size_t)n_chunks));
    CU(cudaMemcpyAsync(sb[12].p, h.data(), sizeof(int64_t) * h.size(), cudaMemcpyHostToDevice, st));
    LAUNCH(ctx, st, "k_iir_chunks x2 + k_iir_carry",
           launch_sos2(x, dtype, (const int64_t*)sb[12].p, (const int64_t*)sb[12].p + n_utt + 1, n_utt, n_chunks, L, sos, P,
                       (double*)sb[13].p, st));
    ctx->launches += 2;
    CU(cudaStreamSynchronize(st));      // h (pageable) was handed to an async copy
    return MPB_OK;
}

int mpb_sos2_host(mpb_ctx* ctx, double* x, const int64_t* utt_off, int32_t n_utt, const double* sos) {
    if (!ctx || !utt_off) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (n_utt == 0 || utt_off[n_utt] == utt_off[0]) return MPB_OK;
    if (!x || utt_off[0] != 0) return fail(MPB_ERR_BAD_ARG, "NULL buffer or utt_off[0] != 0");
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t sz = sizeof(double) * (size_t)utt_off[n_utt];
    CU(ctx->scratch[14].need(sz));
    CU(cudaMemcpyAsync(ctx->scratch[14].p, x, sz, cudaMemcpyHostToDevice, ctx->stream));
    int rc = mpb_sos2_dev(ctx, ctx->stream, ctx->scratch[14].p, MPB_F64, utt_off, n_utt, sos);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpy(x, ctx->scratch[14].p, sz, cudaMemcpyDeviceToHost));
    return MPB_OK;
}

// ---------------------------------------------------------------------------------------------
// Page-locked host memory for result arrays: device -> host copies into it run at full PCIe rate instead of being
// bounced through the driver's staging buffer into fresh pageable memory.  The Python mirror pools these buffers.
int mpb_host_alloc(mpb_ctx* ctx, int64_t bytes, void** out) {
    if (!ctx || !out || bytes < 0) return fail(MPB_ERR_BAD_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaHostAlloc(out, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
    return MPB_OK;
}

int mpb_host_free(mpb_ctx* ctx, void* p) {
    if (!ctx) return fail(MPB_ERR_BAD_ARG, "ctx is NULL");
    if (!p) return MPB_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaFreeHost(p));
    return MPB_OK;
}

}  // extern "C"

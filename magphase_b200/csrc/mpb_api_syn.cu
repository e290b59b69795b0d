// C-ABI: compressed synthesis plan (synthesis_from_compressed).
#include <chrono>

#include "mpb_ctx.h"

using namespace mpb;

extern "C" int mpb_mt19937_uniform_dev(mpb_ctx* ctx, void* stream, uint32_t* key, int32_t* pos, int64_t n, double low,
                                       double high, void* out_dev, int out_dtype);
extern "C" int mpb_sos2_dev(mpb_ctx* ctx, void* stream, void* x, int dtype, const int64_t* utt_off, int32_t n_utt,
                            const double* sos);
extern "C" int mpb_plan_ola_runs(const int32_t* pm, const int64_t* utt_frm_off, int32_t n_utt, int fft_len,
                                 int32_t target_frames, int32_t* out_runs, int64_t capacity, int64_t* n_runs);

struct mpb_syn {
    mpb_ctx* ctx = nullptr;
    int fft_len = 0, n_mag = 0, n_ph = 0, H = 0, HB = 0;
    float* u_mag = nullptr;   // [n_mag][HP]   (rows pitched to 16 bytes, zero padded)
    float* u_ph = nullptr;    // [n_ph][HBP]
    float* tab = nullptr;     // [3][H]
    DevBuf unw[3], unw_flags, unw_cvt, logsq, nspec, gain, host_in[20], out;
    std::mutex mu;
};

// float64 host rows (pitch `cols`) -> float32 device rows pitched to `pitch` >= cols, zero padded
static int upload_f32(const double* src, size_t rows, size_t cols, size_t pitch, float** dst) {
    std::vector<float> h(rows * pitch, 0.0f);
    for (size_t r = 0; r < rows; ++r)
        for (size_t c = 0; c < cols; ++c) h[r * pitch + c] = (float)src[r * cols + c];
    CU(cudaMalloc((void**)dst, sizeof(float) * h.size()));
    CU(cudaMemcpy(*dst, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice));
    return MPB_OK;
}

extern "C" {

int mpb_syn_create(mpb_ctx* ctx, int fft_len, int n_mag, int n_ph, int hb, const double* u_mag, const double* u_ph,
                   const double* tab, mpb_syn** out) {
    if (!ctx || !u_mag || !u_ph || !tab || !out) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (!fft_len_ok(fft_len)) return fail(MPB_ERR_FFT_LEN, "fft_len must be 1024, 2048 or 4096");
    const int H = fft_len / 2 + 1;
    if (n_mag < 1 || n_ph < 1 || n_mag > MEL_MAX_COEFFS || n_ph > MEL_MAX_COEFFS)
        return fail(MPB_ERR_DIM, "mel dimensions must be in 1..256");
    if (hb < 1 || hb > fft_len / 4)
        return fail(MPB_ERR_DIM, "the periodic band (crossfade upper edge) must end at or below fft_len/4 bins");
    CU(cudaSetDevice(ctx->device));
    mpb_syn* s = new mpb_syn();
    s->ctx = ctx; s->fft_len = fft_len; s->n_mag = n_mag; s->n_ph = n_ph; s->H = H; s->HB = hb;
    int rc = upload_f32(u_mag, n_mag, H, (H + 3) & ~3, &s->u_mag);
    if (rc == MPB_OK) rc = upload_f32(u_ph, n_ph, hb, (hb + 3) & ~3, &s->u_ph);
    if (rc == MPB_OK) rc = upload_f32(tab, 3, H, H, &s->tab);
    if (rc != MPB_OK) { delete s; return rc; }
    *out = s;
    return MPB_OK;
}

int mpb_syn_destroy(mpb_syn* s) {
    if (!s) return MPB_OK;
    cudaSetDevice(s->ctx->device);
    cudaFree(s->u_mag); cudaFree(s->u_ph); cudaFree(s->tab);
    for (auto& b : s->unw) b.release();
    for (auto& b : s->host_in) b.release();
    s->unw_flags.release(); s->unw_cvt.release(); s->logsq.release(); s->nspec.release(); s->gain.release(); s->out.release();
    delete s;
    return MPB_OK;
}

// per_linear: 0 = per_phase_type 'magphase', 1 = 'linear', 2 = 'min_phase' (pass need_ph all zero).
// All pointers are DEVICE pointers (fr included: a host struct of device pointers).  Enqueues un-warp,
// noise statistics, gains and the synthesis kernel on `stream`.
int mpb_synthesis_compressed_dev(mpb_syn* s, void* stream, const void* mag_mel, const void* real_mel,
                                 const void* imag_mel, int in_dtype, int64_t n_rows, const uint8_t* need_ph,
                                 const float* noise, int64_t n_noise, const mpb_syn_frames* fr, const int32_t* runs,
                                 int32_t n_runs, int per_linear, void* out, int out_dtype, int64_t n_out) {
    if (!s || !fr) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (!dtype_ok(in_dtype) || !dtype_ok(out_dtype)) return fail(MPB_ERR_BAD_ARG, "unknown dtype");
    if (n_rows < 0 || n_noise < 0 || fr->nfrm < 0 || n_out < 0) return fail(MPB_ERR_BAD_ARG, "negative size");
    if (n_out == 0) return MPB_OK;
    if (!out) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    mpb_ctx* ctx = s->ctx;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (fr->nfrm == 0) { CU(cudaMemsetAsync(out, 0, (out_dtype == MPB_F64 ? 8 : 4) * (size_t)n_out, st)); return MPB_OK; }
    if (!mag_mel || !real_mel || !imag_mel || !need_ph || !noise || !runs || !fr->pm || !fr->ncentre || !fr->nleft ||
        !fr->nright || !fr->voi || !fr->nkind || !fr->win_a || !fr->win_b || !fr->row0 || !fr->utt_frm_off ||
        !fr->utt_out_off || !fr->utt_t0)
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    if (in_dtype == MPB_F32 && (((uintptr_t)mag_mel | (uintptr_t)real_mel | (uintptr_t)imag_mel) & 15))
        return fail(MPB_ERR_BAD_ARG, "float32 feature matrices must be 16-byte aligned (they are fetched by TMA bulk copies)");
    std::lock_guard<std::mutex> lk(s->mu);
    const int HP = (s->H + 3) & ~3, HBP = (s->HB + 3) & ~3;      // scratch row pitches: 16-byte aligned rows
    CU(s->unw[0].need(sizeof(float) * (size_t)n_rows * HP));
    CU(s->unw[1].need(sizeof(float) * (size_t)n_rows * HBP));
    CU(s->unw[2].need(sizeof(float) * (size_t)n_rows * HBP));
    CU(s->logsq.need(sizeof(double) * (size_t)fr->nfrm));
    CU(s->nspec.need(sizeof(float2) * (size_t)fr->nfrm * (s->fft_len / 2 + 2)));
    CU(s->gain.need(sizeof(double) * 2 * (size_t)fr->n_utt));

    UnwarpArgs u;
    u.mag_mel = mag_mel; u.real_mel = real_mel; u.imag_mel = imag_mel; u.in_dtype = in_dtype;
    u.need_ph = need_ph; u.nfrm = n_rows; u.n_mag = s->n_mag; u.n_ph = s->n_ph;
    u.u_mag = s->u_mag; u.H = s->H; u.u_ph = s->u_ph; u.HB = s->HB;
    u.out_mag = (float*)s->unw[0].p; u.out_real = (float*)s->unw[1].p; u.out_imag = (float*)s->unw[2].p;
    u.HP = HP; u.HBP = HBP; u.num_sms = ctx->num_sms;
    CU(s->unw_flags.need((size_t)(n_rows + 63) / 64 + 1));
    u.flags = (uint8_t*)s->unw_flags.p;
    u.cvt = nullptr; u.cvt_pitch = 0;
    if (in_dtype == MPB_F64) {
        const size_t widest = (size_t)(s->n_mag > s->n_ph ? s->n_mag : s->n_ph);
        u.cvt_pitch = ((size_t)n_rows * widest + 63) & ~(size_t)63;       // keeps every matrix 256-byte aligned
        CU(s->unw_cvt.need(sizeof(float) * 3 * u.cvt_pitch));
        u.cvt = (float*)s->unw_cvt.p;
    }
    LAUNCH(ctx, st, "k_mel_unwarp", launch_mel_unwarp(u, st));

    const void* tw = nullptr;
    int rc = get_twiddles(ctx, s->fft_len, MPB_F32, &tw);
    if (rc != MPB_OK) return rc;
    AnalysisArgs n;
    n.sig = noise; n.sig_dtype = MPB_F32; n.n_sig = n_noise;
    n.centre = fr->ncentre; n.left = fr->nleft; n.right = fr->nright; n.win = fr->nkind;
    n.nfrm = fr->nfrm; n.fft_len = s->fft_len; n.compute_dtype = MPB_F32; n.tw = tw;
    n.out_a = s->logsq.p; n.out_b = s->nspec.p; n.out_c = nullptr; n.out_dtype = MPB_F64; n.mode = MODE_LOGSQ;
    n.num_sms = ctx->num_sms;
    LAUNCH(ctx, st, "k_analysis<noise_logsq>", launch_noise_stats(n, st));

    SynthCompArgs a;
    a.m_mag = u.out_mag; a.m_real = u.out_real; a.m_imag = u.out_imag; a.H = s->H; a.HB = s->HB; a.HP = HP; a.HBP = HBP;
    a.noise = noise; a.n_noise = n_noise; a.nspec = (const float2*)s->nspec.p;
    a.pm = fr->pm; a.ncentre = fr->ncentre; a.nleft = fr->nleft; a.nright = fr->nright;
    a.voi = fr->voi; a.nkind = fr->nkind; a.win_a = fr->win_a; a.win_b = fr->win_b;
    a.row0 = fr->row0; a.row1 = fr->row1; a.roww = fr->roww;
    a.logsq = (const double*)s->logsq.p; a.utt_frm_off = fr->utt_frm_off; a.inv_gain = (double*)s->gain.p;
    a.tab = s->tab; a.utt_out_off = fr->utt_out_off; a.utt_t0 = fr->utt_t0; a.n_utt = fr->n_utt;
    a.runs = (const OlaRun*)runs; a.n_runs = n_runs; a.nfrm = fr->nfrm;
    a.fft_len = s->fft_len; a.per_linear = per_linear == 1; a.tw = tw;
    if (per_linear == 2) {   // per_phase_type='min_phase': Re/Im of the minimum-phase spectrum replace the phase rows
        const void* tw64 = nullptr;
        rc = get_twiddles(ctx, s->fft_len, MPB_F64, &tw64);
        if (rc != MPB_OK) return rc;
        LAUNCH(ctx, st, "k_min_phase", launch_min_phase_split(s->fft_len, u.out_mag, HP, n_rows, tw64, u.out_real, u.out_imag,
                                                               s->HB, HBP, ctx->num_sms, st));
    }
    a.out = out; a.out_dtype = out_dtype; a.n_out = n_out; a.num_sms = ctx->num_sms;
    LAUNCH(ctx, st, "k_noise_gain", launch_noise_gain(a, st));
    LAUNCH(ctx, st, "k_synthesis_compressed", launch_synthesis_compressed(a, st));
    return MPB_OK;
}

// HOST pointers everywhere.  noise: the uniform(-1, 1) samples of all utterances, concatenated -- or NULL with
// mt_key / mt_pos (NumPy's legacy MT19937 state, in/out): then the same numbers are generated on the device.
int mpb_synthesis_compressed_host(mpb_syn* s, const double* mag_mel, const double* real_mel, const double* imag_mel,
                                  int64_t n_rows, const uint8_t* need_ph, const double* noise, int64_t n_noise,
                                  uint32_t* mt_key, int32_t* mt_pos, const mpb_syn_frames* fr, int per_linear,
                                  const double* hpf_sos, double* out, int64_t n_out) {
    if (!s || !fr) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (n_out == 0) return MPB_OK;
    if (!mag_mel || !real_mel || !imag_mel || !need_ph || !out) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    if (!noise && !(mt_key && mt_pos)) return fail(MPB_ERR_BAD_ARG, "either noise or an MT19937 state is required");
    const int64_t F = fr->nfrm;
    const int32_t U = fr->n_utt;
    static const bool trace = getenv("MPB_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return 1e3 * std::chrono::duration<double>(b - a).count();
    };
    const auto t0 = now();
    for (int64_t f = 0; f < F; ++f) {
        if (fr->win_a[f] < 0 || fr->win_b[f] < 0 || fr->win_a[f] > s->fft_len / 2 || fr->win_b[f] >= s->fft_len / 2)
            return fail(MPB_ERR_FRAME_GEOM, "anti-ringing window longer than fft_len/2 (f0 too low for this fft_len)");
        if (fr->row0[f] < 0 || fr->row0[f] >= n_rows || (fr->row1 && (fr->row1[f] < 0 || fr->row1[f] >= n_rows)))
            return fail(MPB_ERR_BAD_ARG, "feature row index out of range");
    }
    int rc = check_frames_host(fr->ncentre, fr->nleft, fr->nright, F, n_noise, s->fft_len);
    if (rc != MPB_OK) return rc;
    int64_t n_runs = 0;
    int target = 32;
    {
        const int64_t want = (int64_t)s->ctx->num_sms * 4;
        if (F / target < want) target = (int)(F / want > 1 ? F / want : 1);
    }
    rc = mpb_plan_ola_runs(fr->pm, fr->utt_frm_off, U, s->fft_len, target, nullptr, 0, &n_runs);
    if (rc != MPB_OK) return rc;
    std::vector<int32_t> runs(4 * (size_t)(n_runs > 0 ? n_runs : 1));
    rc = mpb_plan_ola_runs(fr->pm, fr->utt_frm_off, U, s->fft_len, target, runs.data(), n_runs, &n_runs);
    if (rc != MPB_OK) return rc;
    std::vector<float> noise32;
    if (noise) {
        noise32.resize((size_t)n_noise);
        for (int64_t i = 0; i < n_noise; ++i) noise32[i] = (float)noise[i];
    }

    mpb_ctx* ctx = s->ctx;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t st = ctx->stream;
    const auto t1 = now();
    DevBuf* b = s->host_in;
    int bi = 0;
    auto up = [&](const void* src, size_t bytes, const void** dst) -> int {
        if (!src) { *dst = nullptr; ++bi; return MPB_OK; }
        DevBuf& d = b[bi++];
        CU(d.need(bytes > 0 ? bytes : 1));
        CU(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, st));
        *dst = d.p;
        return MPB_OK;
    };
    mpb_syn_frames d = *fr;
    const void *d_mag, *d_real, *d_imag, *d_need, *d_noise, *d_runs;
#define UP(src, bytes, dst) do { rc = up(src, bytes, (const void**)&(dst)); if (rc != MPB_OK) return rc; } while (0)
    UP(mag_mel, sizeof(double) * n_rows * s->n_mag, d_mag);
    UP(real_mel, sizeof(double) * n_rows * s->n_ph, d_real);
    UP(imag_mel, sizeof(double) * n_rows * s->n_ph, d_imag);
    UP(need_ph, (size_t)n_rows, d_need);
    if (noise) {
        UP(noise32.data(), sizeof(float) * n_noise, d_noise);
    } else {
        DevBuf& dn = b[bi++];
        CU(dn.need(sizeof(float) * (size_t)(n_noise > 0 ? n_noise : 1)));
        rc = mpb_mt19937_uniform_dev(ctx, st, mt_key, mt_pos, n_noise, -1.0, 1.0, dn.p, MPB_F32);
        if (rc != MPB_OK) return rc;
        d_noise = dn.p;
    }
    const auto t2 = now();
    UP(runs.data(), sizeof(int32_t) * 4 * n_runs, d_runs);
    UP(fr->pm, sizeof(int32_t) * F, d.pm);
    UP(fr->ncentre, sizeof(int64_t) * F, d.ncentre);
    UP(fr->nleft, sizeof(int32_t) * F, d.nleft);
    UP(fr->nright, sizeof(int32_t) * F, d.nright);
    UP(fr->voi, (size_t)F, d.voi);
    UP(fr->nkind, (size_t)F, d.nkind);
    UP(fr->win_a, sizeof(int32_t) * F, d.win_a);
    UP(fr->win_b, sizeof(int32_t) * F, d.win_b);
    UP(fr->row0, sizeof(int32_t) * F, d.row0);
    UP(fr->row1, sizeof(int32_t) * F, d.row1);
    UP(fr->roww, sizeof(float) * F, d.roww);
    UP(fr->utt_frm_off, sizeof(int64_t) * (U + 1), d.utt_frm_off);
    UP(fr->utt_out_off, sizeof(int64_t) * (U + 1), d.utt_out_off);
    UP(fr->utt_t0, sizeof(int32_t) * U, d.utt_t0);
#undef UP
    const auto t3 = now();
    CU(s->out.need(sizeof(double) * n_out));
    rc = mpb_synthesis_compressed_dev(s, st, d_mag, d_real, d_imag, MPB_F64, n_rows, (const uint8_t*)d_need,
                                      (const float*)d_noise, n_noise, &d, (const int32_t*)d_runs, (int32_t)n_runs,
                                      per_linear, s->out.p, MPB_F64, n_out);
    if (rc != MPB_OK) return rc;
    if (hpf_sos) {                   // output high-pass (src/magphase.py:981-995) on the device, per utterance
        rc = mpb_sos2_dev(ctx, st, s->out.p, MPB_F64, fr->utt_out_off, U, hpf_sos);
        if (rc != MPB_OK) return rc;
    }
    CU(cudaMemcpyAsync(out, s->out.p, sizeof(double) * n_out, cudaMemcpyDeviceToHost, st));
    const auto t4 = now();
    CU(cudaStreamSynchronize(st));   // also keeps noise32 / runs alive until the copies are done
    if (trace)
        fprintf(stderr, "[mpb] synthesis_compressed_host: checks+runs %.3f ms, features+noise %.3f ms, descriptors %.3f ms, "
                        "enqueue %.3f ms, drain %.3f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, now()));
    return MPB_OK;
}

}  // extern "C"

// C-ABI: compressed synthesis plan (synthesis_from_compressed).
#include <algorithm>
#include <chrono>

#include "mpb_ctx.h"

using namespace mpb;

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
    float* ut_mag = nullptr;  // U^T split for the tensor-core un-warp product (mpb_mel_unwarp_tc.cu); NULL when a stream has more
    float* ut_ph = nullptr;   // than 64 coefficients or MPB_MEL_TC=0 was set at plan creation: FMA kernel then
    DevBuf unw_x[3], unw_idx;
    DevBuf unw[3], unw_flags, unw_cvt, logsq, nspec, gain, ticket, frm_rows[3], host_in[20], out;
    const float* noise_ready = nullptr;   // noise buffer whose statistics + spectra mpb_synthesis_noise_stage_dev already enqueued
    int64_t noise_ready_frames = 0;
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
    static const bool mel_tc = [] { const char* e = getenv("MPB_MEL_TC"); return !e || (atoi(e) & 2); }();   // default on; bit 1 clear: FMA kernel
    if (mel_tc && n_mag <= 64 && n_ph <= 64 && (H - 1) % 128 == 0) {
        const int HP = (H + 3) & ~3, HBP = (hb + 3) & ~3;
        CU(cudaMalloc((void**)&s->ut_mag, unwarp_tc_operand_bytes(H - 1)));
        CU(cudaMalloc((void**)&s->ut_ph, unwarp_tc_operand_bytes(hb)));
        CU(build_unwarp_matrix_tc(s->u_mag, n_mag, HP, H - 1, s->ut_mag, ctx->stream));
        CU(build_unwarp_matrix_tc(s->u_ph, n_ph, HBP, hb, s->ut_ph, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->launches += 2;
    }
    *out = s;
    return MPB_OK;
}

int mpb_syn_destroy(mpb_syn* s) {
    if (!s) return MPB_OK;
    cudaSetDevice(s->ctx->device);
    cudaFree(s->u_mag); cudaFree(s->u_ph); cudaFree(s->tab); cudaFree(s->ut_mag); cudaFree(s->ut_ph);
    for (auto& b : s->unw_x) b.release();
    s->unw_idx.release();
    for (auto& b : s->unw) b.release();
    for (auto& b : s->host_in) b.release();
    for (auto& b : s->frm_rows) b.release();
    s->ticket.release(); s->unw_flags.release(); s->unw_cvt.release(); s->logsq.release(); s->nspec.release(); s->gain.release(); s->out.release();
    delete s;
    return MPB_OK;
}

}  // extern "C"

// A slice of a batch: whole utterances [utt_a, utt_b) with their feature rows, frames, OLA runs and output samples.
// `ordinal` numbers the slices of one call (it spaces their private scratch regions).
struct SynRange {
    int64_t row_a, row_b, frm_a, frm_b, out_a, out_b;
    int32_t utt_a, utt_b, run_a, run_b;
    int ordinal;
};

// grows every scratch buffer of a call to its final size (the ranges of a pipelined call run while others are enqueued)
static int syn_reserve(mpb_syn* s, int in_dtype, int64_t n_rows, int64_t nfrm, int32_t n_utt, int n_ranges, size_t* cvt_pitch) {
    const int HP = (s->H + 3) & ~3, HBP = (s->HB + 3) & ~3;      // scratch row pitches: 16-byte aligned rows
    CU(s->unw[0].need(sizeof(float) * (size_t)n_rows * HP));
    CU(s->unw[1].need(sizeof(float) * (size_t)n_rows * HBP));
    CU(s->unw[2].need(sizeof(float) * (size_t)n_rows * HBP));
    CU(s->logsq.need(sizeof(double) * (size_t)nfrm));
    CU(s->nspec.need(sizeof(float2) * (size_t)nfrm * (s->fft_len / 2 + 2)));
    CU(s->gain.need(sizeof(double) * 2 * (size_t)n_utt));
    CU(s->ticket.need(sizeof(int)));
    CU(s->unw_flags.need((size_t)n_rows / 64 + n_ranges + 2));
    if (s->ut_mag) {
        for (int i = 0; i < 3; ++i) CU(s->unw_x[i].need(unwarp_tc_feature_bytes(n_rows)));
        CU(s->unw_idx.need(sizeof(int32_t) * (2 * (size_t)n_rows + 4 * ((size_t)n_ranges + 1) + 8)));
    }
    *cvt_pitch = 0;
    if (in_dtype == MPB_F64) {
        const size_t widest = (size_t)(s->n_mag > s->n_ph ? s->n_mag : s->n_ph);
        *cvt_pitch = ((size_t)n_rows * widest + 4 * ((size_t)n_ranges + 1) + 63) & ~(size_t)63;   // 256-byte aligned matrices
        CU(s->unw_cvt.need(sizeof(float) * 3 * *cvt_pitch));
    }
    return MPB_OK;
}

// un-warp, [min-phase], noise statistics, gains and synthesis of one range, enqueued on st (scratch already reserved)
static int syn_enqueue_range(mpb_syn* s, cudaStream_t st, const void* mag_mel, const void* real_mel, const void* imag_mel,
                             int in_dtype, const uint8_t* need_ph, const float* noise, int64_t n_noise,
                             const mpb_syn_frames* fr, const int32_t* runs, int per_linear, void* out, int out_dtype,
                             size_t cvt_pitch, const SynRange& r) {
    mpb_ctx* ctx = s->ctx;
    const int HP = (s->H + 3) & ~3, HBP = (s->HB + 3) & ~3;
    const size_t ies = in_dtype == MPB_F64 ? 8 : 4, oes = out_dtype == MPB_F64 ? 8 : 4;
    const int64_t n_rows = r.row_b - r.row_a, nfrm = r.frm_b - r.frm_a;
    CU(cudaMemsetAsync((char*)out + oes * r.out_a, 0, oes * (size_t)(r.out_b - r.out_a), st));
    if (nfrm == 0 || n_rows == 0) return MPB_OK;

    UnwarpArgs u;
    u.mag_mel = (const char*)mag_mel + ies * r.row_a * s->n_mag;
    u.real_mel = (const char*)real_mel + ies * r.row_a * s->n_ph;
    u.imag_mel = (const char*)imag_mel + ies * r.row_a * s->n_ph;
    u.in_dtype = in_dtype;
    u.need_ph = need_ph + r.row_a; u.nfrm = n_rows; u.n_mag = s->n_mag; u.n_ph = s->n_ph;
    u.u_mag = s->u_mag; u.H = s->H; u.u_ph = s->u_ph; u.HB = s->HB;
    u.out_mag = (float*)s->unw[0].p + r.row_a * HP; u.out_real = (float*)s->unw[1].p + r.row_a * HBP;
    u.out_imag = (float*)s->unw[2].p + r.row_a * HBP;
    u.HP = HP; u.HBP = HBP; u.num_sms = ctx->num_sms;
    u.flags = (uint8_t*)s->unw_flags.p + r.row_a / 64 + r.ordinal;
    u.cvt = nullptr; u.cvt_pitch = cvt_pitch; u.cvt_off_mag = 0; u.cvt_off_ph = 0;
    if (in_dtype == MPB_F64) {
        // private, 16-byte aligned regions of the narrowing scratch (TMA sources) for this range
        u.cvt = (float*)s->unw_cvt.p;
        u.cvt_off_mag = (((size_t)r.row_a * s->n_mag + 3) & ~(size_t)3) + 4 * (size_t)r.ordinal;
        u.cvt_off_ph = (((size_t)r.row_a * s->n_ph + 3) & ~(size_t)3) + 4 * (size_t)r.ordinal;
    }
    if (s->ut_mag) {
        // tensor-core product: compaction of the frames that need their phase rows, split operands, tiles (mpb_mel_unwarp_tc.cu)
        int32_t* vidx = (int32_t*)s->unw_idx.p + r.row_a;
        int32_t* cidx = (int32_t*)s->unw_idx.p + (((size_t)(s->unw_idx.cap / sizeof(int32_t) / 2) + 3) & ~(size_t)3) + r.row_a;
        int32_t* cnt = (int32_t*)s->unw_idx.p + s->unw_idx.cap / sizeof(int32_t) - 4 - r.ordinal;
        u.ut_mag = s->ut_mag; u.ut_ph = s->ut_ph;
        for (int i = 0; i < 3; ++i) u.xs[i] = (float*)s->unw_x[i].p + (size_t)r.row_a * 128;
        u.vidx = vidx; u.cidx = cidx; u.vcount = cnt;
    }
    if (unwarp_tc_usable(u)) {
        LAUNCH(ctx, st, "k_voiced_compact", launch_voiced_compact(u.need_ph, (int)n_rows, (int32_t*)u.vidx, (int32_t*)u.cidx,
                                                                  (int32_t*)u.vcount, st));
        LAUNCH(ctx, st, "k_mel_unwarp_tc", launch_mel_unwarp_tc(u, st));
    } else {
        LAUNCH(ctx, st, "k_mel_unwarp", launch_mel_unwarp(u, st));
    }

    const void* tw = nullptr;
    int rc = get_twiddles(ctx, s->fft_len, MPB_F32, &tw);
    if (rc != MPB_OK) return rc;
    const int SP = s->fft_len / 2 + 2;
    AnalysisArgs n;
    n.sig = noise; n.sig_dtype = MPB_F32; n.n_sig = n_noise;
    n.centre = fr->ncentre + r.frm_a; n.left = fr->nleft + r.frm_a; n.right = fr->nright + r.frm_a; n.win = fr->nkind + r.frm_a;
    n.nfrm = nfrm; n.fft_len = s->fft_len; n.compute_dtype = MPB_F32; n.tw = tw;
    n.out_a = (double*)s->logsq.p + r.frm_a; n.out_b = (float2*)s->nspec.p + r.frm_a * SP; n.out_c = nullptr;
    n.out_dtype = MPB_F64; n.mode = MODE_LOGSQ;
    n.num_sms = ctx->num_sms;
    if (s->noise_ready == noise && s->noise_ready_frames == fr->nfrm && r.frm_a == 0 && r.frm_b == fr->nfrm) {
        s->noise_ready = nullptr;                  // consumed: the pre-stage ran for exactly this buffer and frame count
    } else {
        LAUNCH(ctx, st, "k_analysis<noise_logsq>", launch_noise_stats(n, st));
    }

    SynthCompArgs a;
    a.m_mag = (const float*)s->unw[0].p; a.m_real = (const float*)s->unw[1].p; a.m_imag = (const float*)s->unw[2].p;
    a.H = s->H; a.HB = s->HB; a.HP = HP; a.HBP = HBP;
    a.noise = noise; a.n_noise = n_noise; a.nspec = (const float2*)s->nspec.p;
    a.pm = fr->pm; a.ncentre = fr->ncentre; a.nleft = fr->nleft; a.nright = fr->nright;
    a.voi = fr->voi; a.nkind = fr->nkind; a.win_a = fr->win_a; a.win_b = fr->win_b;
    a.row0 = fr->row0; a.row1 = fr->row1; a.roww = fr->roww;
    a.logsq = (const double*)s->logsq.p; a.utt_frm_off = fr->utt_frm_off; a.inv_gain = (double*)s->gain.p;
    a.tab = s->tab; a.utt_out_off = fr->utt_out_off; a.utt_t0 = fr->utt_t0; a.n_utt = fr->n_utt;
    a.utt_a = r.utt_a; a.utt_b = r.utt_b;
    a.runs = (const OlaRun*)runs + r.run_a; a.n_runs = r.run_b - r.run_a; a.nfrm = fr->nfrm;
    a.run_ticket = (int*)s->ticket.p;
    a.fft_len = s->fft_len; a.per_linear = per_linear == 1; a.tw = tw;
    if (per_linear == 2) {   // per_phase_type='min_phase': Re/Im of the minimum-phase spectrum replace the phase rows
        const void* tw64 = nullptr;
        rc = get_twiddles(ctx, s->fft_len, MPB_F64, &tw64);
        if (rc != MPB_OK) return rc;
        if (fr->row1) {
            // constant-rate input: the reference interpolates the un-warped magnitudes to the synthesis frames first and
            // builds the minimum phase of the INTERPOLATED rows (src/magphase.py:861-870, then :935-936): materialise
            // them per frame; the synthesis kernel then reads frame f's own row, no interpolation left
            const int64_t F = fr->nfrm;
            CU(s->frm_rows[0].need(sizeof(float) * (size_t)F * HP));
            CU(s->frm_rows[1].need(sizeof(float) * (size_t)F * HBP));
            CU(s->frm_rows[2].need(sizeof(float) * (size_t)F * HBP));
            float* fm = (float*)s->frm_rows[0].p + r.frm_a * HP;
            float* fre = (float*)s->frm_rows[1].p + r.frm_a * HBP;
            float* fim = (float*)s->frm_rows[2].p + r.frm_a * HBP;
            LAUNCH(ctx, st, "k_lerp_rows", launch_lerp_rows((const float*)s->unw[0].p, HP, fr->row0 + r.frm_a, fr->row1 + r.frm_a,
                                                            fr->roww + r.frm_a, nfrm, fm, st));
            LAUNCH(ctx, st, "k_min_phase", launch_min_phase_split(s->fft_len, fm, HP, nfrm, tw64, fre, fim, s->HB, HBP,
                                                                   ctx->num_sms, st));
            a.m_mag = (const float*)s->frm_rows[0].p; a.m_real = (const float*)s->frm_rows[1].p;
            a.m_imag = (const float*)s->frm_rows[2].p;
            a.row0 = nullptr; a.row1 = nullptr; a.roww = nullptr;
        } else {
            LAUNCH(ctx, st, "k_min_phase", launch_min_phase_split(s->fft_len, u.out_mag, HP, n_rows, tw64, u.out_real,
                                                                   u.out_imag, s->HB, HBP, ctx->num_sms, st));
        }
    }
    a.out = out; a.out_dtype = out_dtype; a.n_out = 0; a.num_sms = ctx->num_sms;
    LAUNCH(ctx, st, "k_noise_gain", launch_noise_gain(a, st));
    LAUNCH(ctx, st, "k_synthesis_compressed", launch_synthesis_compressed(a, st));
    return MPB_OK;
}

extern "C" {

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
    size_t cvt_pitch = 0;
    int rc = syn_reserve(s, in_dtype, n_rows, fr->nfrm, fr->n_utt, 1, &cvt_pitch);
    if (rc != MPB_OK) return rc;
    SynRange r;
    r.row_a = 0; r.row_b = n_rows; r.frm_a = 0; r.frm_b = fr->nfrm; r.out_a = 0; r.out_b = n_out;
    r.utt_a = 0; r.utt_b = fr->n_utt; r.run_a = 0; r.run_b = n_runs; r.ordinal = 0;
    return syn_enqueue_range(s, st, mag_mel, real_mel, imag_mel, in_dtype, need_ph, noise, n_noise, fr, runs, per_linear, out,
                             out_dtype, cvt_pitch, r);
}

// The noise half of the synthesis on its own: windowed noise frames -> FFT -> per-frame statistics + stored spectra
// (k_analysis<noise_logsq>).  It depends on the noise samples and the frame geometry only, not on the features, so a caller
// that holds the features of the batch back (bench: they are still being analysed) can enqueue it on ANOTHER stream first;
// the next mpb_synthesis_compressed_dev call with the same noise buffer and frame count then skips that stage.  Ordering
// between the two streams is the caller's (an event after this call, waited for before the synthesis call).
int mpb_synthesis_noise_stage_dev(mpb_syn* s, void* stream, const float* noise, int64_t n_noise, const mpb_syn_frames* fr) {
    if (!s || !fr) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (n_noise < 0 || fr->nfrm < 0) return fail(MPB_ERR_BAD_ARG, "negative size");
    if (fr->nfrm == 0) return MPB_OK;
    if (!noise || !fr->ncentre || !fr->nleft || !fr->nright || !fr->nkind) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    mpb_ctx* ctx = s->ctx;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lk(s->mu);
    const int SP = s->fft_len / 2 + 2;
    CU(s->logsq.need(sizeof(double) * (size_t)fr->nfrm));
    CU(s->nspec.need(sizeof(float2) * (size_t)fr->nfrm * SP));
    const void* tw = nullptr;
    int rc = get_twiddles(ctx, s->fft_len, MPB_F32, &tw);
    if (rc != MPB_OK) return rc;
    AnalysisArgs n;
    n.sig = noise; n.sig_dtype = MPB_F32; n.n_sig = n_noise;
    n.centre = fr->ncentre; n.left = fr->nleft; n.right = fr->nright; n.win = fr->nkind;
    n.nfrm = fr->nfrm; n.fft_len = s->fft_len; n.compute_dtype = MPB_F32; n.tw = tw;
    n.out_a = (double*)s->logsq.p; n.out_b = (float2*)s->nspec.p; n.out_c = nullptr;
    n.out_dtype = MPB_F64; n.mode = MODE_LOGSQ;
    n.num_sms = ctx->num_sms;
    LAUNCH(ctx, st, "k_analysis<noise_logsq>", launch_noise_stats(n, st));
    s->noise_ready = noise;
    s->noise_ready_frames = fr->nfrm;
    return MPB_OK;
}

// Host buffers in, host buffer out.  Like the analysis entry point this is a three-stage pipeline over groups of
// utterances: features of group g+1 cross PCIe on stream_in while group g runs un-warp / noise statistics / gains /
// synthesis on the compute stream and the waveform of group g-1 returns on stream_out.  The noise of the whole batch
// is drawn first on the compute stream (the MT19937 stream is one sequence); its final state comes back through
// page-locked memory at the end.  Frame descriptors travel as one page-locked block.
int mpb_synthesis_compressed_host2(mpb_syn* s, const void* mag_mel_v, const void* real_mel_v, const void* imag_mel_v, int in_dtype,
                                   int64_t n_rows, const uint8_t* need_ph, const double* noise, int64_t n_noise,
                                   uint32_t* mt_key, int32_t* mt_pos, const mpb_syn_frames* fr, int per_linear,
                                   const double* hpf_sos, void* out_v, int out_dtype, int64_t n_out);

// feature rows handed over as one host block per utterance (mpb_synthesis_compressed_hostv2) instead of stacked matrices
struct RowBlocks {
    const void* const* blk[3];        // mag, real, imag: n pointers each
    const int64_t* rows;              // rows per block
    int32_t n;
    std::vector<int64_t> off;         // [n + 1] first row of every block
};
static int syn_host_pipeline(const RowBlocks* rb, mpb_syn* s, const void* mag_mel_v, const void* real_mel_v, const void* imag_mel_v,
                             int in_dtype, int64_t n_rows, const uint8_t* need_ph, const double* noise, int64_t n_noise,
                             uint32_t* mt_key, int32_t* mt_pos, const mpb_syn_frames* fr, int per_linear, const double* hpf_sos,
                             void* out_v, int out_dtype, int64_t n_out);

int mpb_synthesis_compressed_host2(mpb_syn* s, const void* mag_mel_v, const void* real_mel_v, const void* imag_mel_v, int in_dtype,
                                   int64_t n_rows, const uint8_t* need_ph, const double* noise, int64_t n_noise,
                                   uint32_t* mt_key, int32_t* mt_pos, const mpb_syn_frames* fr, int per_linear,
                                   const double* hpf_sos, void* out_v, int out_dtype, int64_t n_out) {
    return syn_host_pipeline(nullptr, s, mag_mel_v, real_mel_v, imag_mel_v, in_dtype, n_rows, need_ph, noise, n_noise, mt_key, mt_pos,
                             fr, per_linear, hpf_sos, out_v, out_dtype, n_out);
}

// mpb_synthesis_compressed_host2 with the feature rows of every utterance in a host block of its own (what a caller that read
// one feature file per utterance holds): block u has block_rows[u] rows; with variable-rate features block u must be
// utterance u of `fr`.  The host pool copies the blocks straight into page-locked staging -- no stacked copy on the caller's side.
int mpb_synthesis_compressed_hostv2(mpb_syn* s, const void* const* mag_blocks, const void* const* real_blocks,
                                    const void* const* imag_blocks, const int64_t* block_rows, int32_t n_blocks, int in_dtype,
                                    const uint8_t* need_ph, const double* noise, int64_t n_noise, uint32_t* mt_key, int32_t* mt_pos,
                                    const mpb_syn_frames* fr, int per_linear, const double* hpf_sos, void* out_v, int out_dtype,
                                    int64_t n_out) {
    if (!s || !fr || !mag_blocks || !real_blocks || !imag_blocks || !block_rows || n_blocks < 0)
        return fail(MPB_ERR_BAD_ARG, "NULL argument");
    RowBlocks rb;
    rb.blk[0] = mag_blocks; rb.blk[1] = real_blocks; rb.blk[2] = imag_blocks; rb.rows = block_rows; rb.n = n_blocks;
    rb.off.assign((size_t)n_blocks + 1, 0);
    for (int32_t b = 0; b < n_blocks; ++b) {
        if (block_rows[b] < 0) return fail(MPB_ERR_BAD_ARG, "negative block size");
        rb.off[b + 1] = rb.off[b] + block_rows[b];
    }
    return syn_host_pipeline(&rb, s, nullptr, nullptr, nullptr, in_dtype, rb.off[n_blocks], need_ph, noise, n_noise, mt_key, mt_pos, fr,
                             per_linear, hpf_sos, out_v, out_dtype, n_out);
}

int mpb_synthesis_compressed_host(mpb_syn* s, const double* mag_mel, const double* real_mel, const double* imag_mel,
                                  int64_t n_rows, const uint8_t* need_ph, const double* noise, int64_t n_noise,
                                  uint32_t* mt_key, int32_t* mt_pos, const mpb_syn_frames* fr, int per_linear,
                                  const double* hpf_sos, double* out, int64_t n_out) {
    return mpb_synthesis_compressed_host2(s, mag_mel, real_mel, imag_mel, MPB_F64, n_rows, need_ph, noise, n_noise, mt_key, mt_pos,
                                          fr, per_linear, hpf_sos, out, MPB_F64, n_out);
}

// The pipeline with the caller's element types: features float64 or float32 (in_dtype; the reference's feature files are
// float32, src/libutils.py:112-127), waveform float64 or float32 (out_dtype).  float32 halves the bytes on PCIe both ways.
// rb == NULL: stacked feature matrices (mag_mel_v, ...); else one host block per utterance.
static int syn_host_pipeline(const RowBlocks* rb, mpb_syn* s, const void* mag_mel_v, const void* real_mel_v, const void* imag_mel_v,
                             int in_dtype, int64_t n_rows, const uint8_t* need_ph, const double* noise, int64_t n_noise,
                             uint32_t* mt_key, int32_t* mt_pos, const mpb_syn_frames* fr, int per_linear, const double* hpf_sos,
                             void* out_v, int out_dtype, int64_t n_out) {
    if (!s || !fr) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (!dtype_ok(in_dtype) || !dtype_ok(out_dtype)) return fail(MPB_ERR_BAD_ARG, "unknown dtype");
    const char* mag_mel = (const char*)mag_mel_v; const char* real_mel = (const char*)real_mel_v;
    const char* imag_mel = (const char*)imag_mel_v; char* out = (char*)out_v;
    const size_t ies = in_dtype == MPB_F64 ? 8 : 4, oes = out_dtype == MPB_F64 ? 8 : 4;
    if (n_out == 0) return MPB_OK;
    if ((!rb && (!mag_mel || !real_mel || !imag_mel)) || !need_ph || !out) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    if (!noise && !(mt_key && mt_pos)) return fail(MPB_ERR_BAD_ARG, "either noise or an MT19937 state is required");
    if (!noise && (*mt_pos < 0 || *mt_pos > 624)) return fail(MPB_ERR_BAD_ARG, "bad MT19937 position");
    const int64_t F = fr->nfrm;
    const int32_t U = fr->n_utt;
    if (F < 0 || U < 0 || n_rows < 0 || n_noise < 0) return fail(MPB_ERR_BAD_ARG, "negative size");
    if (F > 0 && (!fr->pm || !fr->ncentre || !fr->nleft || !fr->nright || !fr->voi || !fr->nkind || !fr->win_a || !fr->win_b ||
                  !fr->row0 || !fr->utt_frm_off || !fr->utt_out_off || !fr->utt_t0))
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    static const bool trace = getenv("MPB_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return 1e3 * std::chrono::duration<double>(b - a).count();
    };
    const auto t0 = now();
    bool identity_rows = fr->row1 == nullptr && n_rows == F;     // variable rate: frame f reads feature row f
    for (int64_t f = 0; f < F; ++f) {
        if (fr->win_a[f] < 0 || fr->win_b[f] < 0 || fr->win_a[f] > s->fft_len / 2 || fr->win_b[f] >= s->fft_len / 2)
            return fail(MPB_ERR_FRAME_GEOM, "anti-ringing window longer than fft_len/2 (f0 too low for this fft_len)");
        if (fr->row0[f] < 0 || fr->row0[f] >= n_rows || (fr->row1 && (fr->row1[f] < 0 || fr->row1[f] >= n_rows)))
            return fail(MPB_ERR_BAD_ARG, "feature row index out of range");
        identity_rows = identity_rows && fr->row0[f] == f;
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

    // ---- groups of whole utterances with about the same number of frames; constant-rate input (frames address
    // arbitrary feature rows) and the output high-pass (a per-utterance scan that synchronises) run as one group ----
    int n_groups = (identity_rows && !hpf_sos) ? pipeline_groups(F, 10000, 6) : 1;
    if (n_groups > U) n_groups = U > 0 ? U : 1;
    std::vector<SynRange> rg;
    {
        int32_t ua = 0, run_i = 0;
        for (int g = 0; g < n_groups && ua < U; ++g) {
            int32_t ub = ua + 1;
            if (g == n_groups - 1) ub = U;
            else while (ub < U - (n_groups - 1 - g) && fr->utt_frm_off[ub] * n_groups < F * (g + 1)) ++ub;
            SynRange r;
            r.utt_a = ua; r.utt_b = ub; r.ordinal = g;
            r.frm_a = fr->utt_frm_off[ua]; r.frm_b = fr->utt_frm_off[ub];
            r.row_a = n_groups == 1 ? 0 : r.frm_a; r.row_b = n_groups == 1 ? n_rows : r.frm_b;
            r.out_a = fr->utt_out_off[ua]; r.out_b = fr->utt_out_off[ub];
            r.run_a = run_i;
            while (run_i < n_runs && runs[4 * (size_t)run_i + 2] < ub) ++run_i;
            r.run_b = run_i;
            rg.push_back(r);
            ua = ub;
        }
        if (rg.empty()) { SynRange r{0, n_rows, 0, F, 0, n_out, 0, U, 0, (int32_t)n_runs, 0}; rg.push_back(r); }
        rg.back().out_b = n_out;
    }

    mpb_ctx* ctx = s->ctx;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::lock_guard<std::mutex> lk2(s->mu);
    cudaStream_t s_in = ctx->stream_in, s_cmp = ctx->stream, s_out = ctx->stream_out;
    const auto t1 = now();
    DevBuf* b = s->host_in;
    enum { B_MAG = 0, B_REAL, B_IMAG, B_NOISE, B_DESC };
    CU(b[B_NOISE].need(sizeof(float) * (size_t)(n_noise > 0 ? n_noise : 1)));
    // ---- the NumPy noise stream goes first, on its own stream: the twister is latency bound (a few SMs), so it runs
    // under the descriptor packing and the first feature upload; the first group's kernels wait for e_rng ----
    uint32_t* mt_fin = nullptr;
    std::vector<cudaEvent_t> e_rng, evs;
    std::vector<float> noise32;
    PipelineDrain drain(ctx, &evs, &e_rng);        // from here on every exit path drains the streams first
    if (!noise && n_noise > 0) {
        CU(ctx->mt_fin.need(sizeof(uint32_t) * 625));
        mt_fin = (uint32_t*)ctx->mt_fin.p;
        e_rng.push_back(get_event(ctx));
        rc = mt19937_enqueue(ctx, ctx->stream_aux, mt_key, *mt_pos, &n_noise, 1, -1.0, 1.0, b[B_NOISE].p, MPB_F32, mt_fin,
                             e_rng.data());
        if (rc != MPB_OK) return rc;
    }

    // ---- device buffers, all at their final size before anything is enqueued ----
    CU(b[B_MAG].need(ies * (size_t)n_rows * s->n_mag + 16));
    CU(b[B_REAL].need(ies * (size_t)n_rows * s->n_ph + 16));
    CU(b[B_IMAG].need(ies * (size_t)n_rows * s->n_ph + 16));
    CU(s->out.need(oes * (size_t)n_out));
    size_t cvt_pitch = 0;
    rc = syn_reserve(s, in_dtype, n_rows, F, U, (int)rg.size(), &cvt_pitch);
    if (rc != MPB_OK) return rc;

    // ---- descriptors: one page-locked block, one copy ----
    struct Item { const void* src; size_t bytes; const void** dst; };
    mpb_syn_frames d = *fr;
    const void *d_need = nullptr, *d_runs = nullptr;
    const Item items[] = {
        {need_ph, (size_t)n_rows, &d_need},
        {runs.data(), sizeof(int32_t) * 4 * (size_t)n_runs, &d_runs},
        {fr->pm, sizeof(int32_t) * F, (const void**)&d.pm},
        {fr->ncentre, sizeof(int64_t) * F, (const void**)&d.ncentre},
        {fr->nleft, sizeof(int32_t) * F, (const void**)&d.nleft},
        {fr->nright, sizeof(int32_t) * F, (const void**)&d.nright},
        {fr->voi, (size_t)F, (const void**)&d.voi},
        {fr->nkind, (size_t)F, (const void**)&d.nkind},
        {fr->win_a, sizeof(int32_t) * F, (const void**)&d.win_a},
        {fr->win_b, sizeof(int32_t) * F, (const void**)&d.win_b},
        {fr->row0, sizeof(int32_t) * F, (const void**)&d.row0},
        {fr->row1, fr->row1 ? sizeof(int32_t) * F : 0, (const void**)&d.row1},
        {fr->roww, fr->roww ? sizeof(float) * F : 0, (const void**)&d.roww},
        {fr->utt_frm_off, sizeof(int64_t) * ((size_t)U + 1), (const void**)&d.utt_frm_off},
        {fr->utt_out_off, sizeof(int64_t) * ((size_t)U + 1), (const void**)&d.utt_out_off},
        {fr->utt_t0, sizeof(int32_t) * (size_t)U, (const void**)&d.utt_t0},
    };
    size_t d_bytes = 0;
    for (const Item& it : items) d_bytes += (it.bytes + 15) & ~(size_t)15;
    CU(ctx->desc_stage.need(d_bytes > 0 ? d_bytes : 16));
    CU(b[B_DESC].need(d_bytes > 0 ? d_bytes : 16));
    {
        char* h = (char*)ctx->desc_stage.p;
        size_t off = 0;
        for (const Item& it : items) {
            if (it.src && it.bytes) { memcpy(h + off, it.src, it.bytes); *it.dst = (const char*)b[B_DESC].p + off; }
            else *it.dst = nullptr;
            off += (it.bytes + 15) & ~(size_t)15;
        }
        if (d_bytes) CU(cudaMemcpyAsync(b[B_DESC].p, h, d_bytes, cudaMemcpyHostToDevice, s_in));
    }

    // ---- noise: explicit samples (narrowed on the host, uploaded) or the NumPy stream advanced on the device ----
    if (noise) {
        noise32.resize((size_t)n_noise);
        for (int64_t i = 0; i < n_noise; ++i) noise32[i] = (float)noise[i];
        if (n_noise) CU(cudaMemcpyAsync(b[B_NOISE].p, noise32.data(), sizeof(float) * n_noise, cudaMemcpyHostToDevice, s_in));
    }
    const auto t2 = now();

    // ---- the pipeline ----
    const size_t stage_real = (ies * (size_t)n_rows * s->n_mag + 255) & ~(size_t)255;
    const size_t stage_imag = stage_real + ((ies * (size_t)n_rows * s->n_ph + 255) & ~(size_t)255);
    if (n_rows > 0 && (rb || !(host_is_page_locked(mag_mel) && host_is_page_locked(real_mel) && host_is_page_locked(imag_mel))))
        CU(ctx->stage_feat.need(stage_imag + ies * (size_t)n_rows * s->n_ph));
    for (const SynRange& r : rg) {
        const int64_t nr = r.row_b - r.row_a;
        if (nr > 0) {
            // (pageable sources -- features read from files -- are staged by the host pool; page-locked ones go straight to DMA)
            const size_t o_mag = ies * r.row_a * s->n_mag, o_ph = ies * r.row_a * s->n_ph;
            if (rb) {
                // the blocks that make up rows [row_a, row_b): ranges are whole utterances, so they start and end on block edges
                const int32_t b0 = (int32_t)(std::lower_bound(rb->off.begin(), rb->off.end(), r.row_a) - rb->off.begin());
                const int32_t b1 = (int32_t)(std::lower_bound(rb->off.begin(), rb->off.end(), r.row_b) - rb->off.begin());
                if (b0 > rb->n || b1 > rb->n || rb->off[b0] != r.row_a || rb->off[b1] != r.row_b) {
                    rc = fail(MPB_ERR_BAD_ARG, "feature blocks do not line up with the utterances");
                    break;
                }
                rc = h2d_gather_staged(ctx, s_in, (char*)b[B_MAG].p + o_mag, rb->blk[0], rb->rows, b0, b1, ies * s->n_mag,
                                       ctx->stage_feat, o_mag);
                if (rc == MPB_OK)
                    rc = h2d_gather_staged(ctx, s_in, (char*)b[B_REAL].p + o_ph, rb->blk[1], rb->rows, b0, b1, ies * s->n_ph,
                                           ctx->stage_feat, stage_real + o_ph);
                if (rc == MPB_OK)
                    rc = h2d_gather_staged(ctx, s_in, (char*)b[B_IMAG].p + o_ph, rb->blk[2], rb->rows, b0, b1, ies * s->n_ph,
                                           ctx->stage_feat, stage_imag + o_ph);
            } else {
                rc = h2d_staged(ctx, s_in, (char*)b[B_MAG].p + o_mag, mag_mel + o_mag, ies * nr * s->n_mag, ctx->stage_feat, o_mag);
                if (rc == MPB_OK)
                    rc = h2d_staged(ctx, s_in, (char*)b[B_REAL].p + o_ph, real_mel + o_ph, ies * nr * s->n_ph, ctx->stage_feat,
                                    stage_real + o_ph);
                if (rc == MPB_OK)
                    rc = h2d_staged(ctx, s_in, (char*)b[B_IMAG].p + o_ph, imag_mel + o_ph, ies * nr * s->n_ph, ctx->stage_feat,
                                    stage_imag + o_ph);
            }
            if (rc != MPB_OK) break;
        }
        cudaEvent_t e_in = get_event(ctx), e_cmp = get_event(ctx);
        evs.push_back(e_in); evs.push_back(e_cmp);
        CU(cudaEventRecord(e_in, s_in));
        CU(cudaStreamWaitEvent(s_cmp, e_in, 0));
        if (!e_rng.empty() && r.ordinal == 0) CU(cudaStreamWaitEvent(s_cmp, e_rng[0], 0));
        rc = syn_enqueue_range(s, s_cmp, b[B_MAG].p, b[B_REAL].p, b[B_IMAG].p, in_dtype, (const uint8_t*)d_need,
                               (const float*)b[B_NOISE].p, n_noise, &d, (const int32_t*)d_runs, per_linear, s->out.p,
                               out_dtype, cvt_pitch, r);
        if (rc != MPB_OK) break;
        if (!hpf_sos) {
            // (float64 on the wire: narrowing the waveform to float32 for PCIe and widening it again on the host was
            // measured slower -- the widening pass costs more host memory bandwidth than the DMA engine saves)
            CU(cudaEventRecord(e_cmp, s_cmp));
            CU(cudaStreamWaitEvent(s_out, e_cmp, 0));
            if (r.out_b > r.out_a)
                CU(cudaMemcpyAsync(out + oes * r.out_a, (char*)s->out.p + oes * r.out_a, oes * (r.out_b - r.out_a),
                                   cudaMemcpyDeviceToHost, s_out));
        }
    }
    const auto t3 = now();
    if (rc == MPB_OK && hpf_sos) {   // output high-pass (src/magphase.py:981-995) on the device, per utterance
        rc = mpb_sos2_dev(ctx, s_cmp, s->out.p, out_dtype, fr->utt_out_off, U, hpf_sos);
        if (rc == MPB_OK) {
            cudaError_t e = cudaMemcpyAsync(out, s->out.p, oes * n_out, cudaMemcpyDeviceToHost, s_cmp);
            if (e != cudaSuccess) rc = fail(MPB_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    // drain every stage, also after an error: noise32 / runs / the staging blocks must outlive the copies
    cudaError_t e1 = host_wait(ctx, s_in), e2 = host_wait(ctx, s_cmp), e3 = host_wait(ctx, s_out);
    cudaError_t e4 = host_wait(ctx, ctx->stream_aux);
    if (rc != MPB_OK) return rc;
    CU(e1); CU(e2); CU(e3); CU(e4);
    if (mt_fin) {
        memcpy(mt_key, mt_fin, sizeof(uint32_t) * 624);
        *mt_pos = (int32_t)mt_fin[624];
    }
    if (trace)
        fprintf(stderr, "[mpb] synthesis_compressed_host: %d groups, checks+runs %.3f ms, descriptors+noise %.3f ms, "
                        "enqueue %.3f ms, drain %.3f ms\n", (int)rg.size(), ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, now()));
    return MPB_OK;
}

}  // extern "C"

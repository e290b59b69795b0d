// C-ABI: mel-compression plan (format_for_modelling) and the fused host entry point of analysis_compressed.
#include <chrono>

#include "mpb_ctx.h"

using namespace mpb;

struct mpb_mel {
    mpb_ctx* ctx = nullptr;
    int fft_len = 0, n_mag = 0, n_ph = 0, phase_dim = 0, ld_mag = 0, ld_ph = 0;
    double alpha_mag = 0, alpha_ph = 0;
    float* wt_mag = nullptr;     // [kpad][ld_mag]
    float* wt_ph = nullptr;      // [kpad][ld_ph]
    float* wt_tc_mag = nullptr;  // W^T pre-split per 32-bin stage for the tcgen05 tile product (mpb_mel_warp_tc.cu); NULL when a stream
    float* wt_tc_ph = nullptr;   // has more than 64 coefficients or MPB_MEL_TC=0 was set at plan creation: FMA kernels then
    double* cos_mag = nullptr;   // [n_mag][n_mag]
    double* cos_ph = nullptr;    // [n_ph][phase_dim]
    DevBuf partial, feats[3], small[8], compact, sig32;
    std::mutex mu;
};

static constexpr int64_t MEL_CHUNK_DEFAULT = 131072;   // frames per pass: bounds the scratch (log periodograms 3.2 GB + mel cepstra 0.1 GB);
                                              // measured 32768 / 65536 / 131072: 3.62 / 3.50 / 3.48 ms per 116k frames

// (MPB_MEL_CHUNK overrides it: tests run the chunk loop with a few thousand frames per pass)
static int64_t mel_chunk() {
    static const int64_t v = [] { const char* e = getenv("MPB_MEL_CHUNK"); const long long x = e ? atoll(e) : 0; return x >= 256 ? (int64_t)x : MEL_CHUNK_DEFAULT; }();
    return v;
}
#define MEL_CHUNK mel_chunk()

static int pad64(int n) { return ((n + 63) / 64) * 64; }
static int lp_pitch_of(int H) { return (H + 3) & ~3; }   // row pitch of the log-periodogram scratch: 16-byte rows (TMA)

// grows the scratch of mpb_analysis_compressed_dev / mel_compress_impl for calls of up to nfrm frames
static int mel_reserve(mpb_mel* m, int64_t nfrm) {
    if (nfrm < 1) return MPB_OK;
    std::lock_guard<std::mutex> lk(m->mu);
    const int H = m->fft_len / 2 + 1;
    const int n_slices = (H - 1) / MEL_KSLICE;
    const int ncp = m->ld_mag > m->ld_ph ? m->ld_mag : m->ld_ph;
    const int64_t chunk = nfrm < MEL_CHUNK ? nfrm : MEL_CHUNK;
    for (int i = 0; i < 3; ++i) CU(m->feats[i].need(sizeof(float) * (size_t)chunk * lp_pitch_of(H)));
    CU(m->partial.need(sizeof(float) * 3 * (size_t)(m->wt_tc_mag ? 1 : n_slices) * (size_t)chunk * ncp));
    CU(m->compact.need(sizeof(int32_t) * (2 * (size_t)((chunk + 3) & ~(int64_t)3) + 4)));
    return MPB_OK;
}

extern "C" {

int mpb_mel_create(mpb_ctx* ctx, int fft_len, double alpha_mag, int n_mag, double alpha_ph, int n_ph, int phase_dim,
                   const double* cos_mag, const double* cos_ph, mpb_mel** out) {
    if (!ctx || !out || !cos_mag || !cos_ph) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (!fft_len_ok(fft_len)) return fail(MPB_ERR_FFT_LEN, "fft_len must be 1024, 2048 or 4096");
    if (n_mag < 2 || n_ph < 2 || phase_dim < 1 || phase_dim > n_ph || n_mag > MEL_MAX_COEFFS || n_ph > MEL_MAX_COEFFS)
        return fail(MPB_ERR_DIM, "mel dimensions must satisfy 2 <= mag_dim, nmel <= 256 and 1 <= phase_dim <= nmel");
    CU(cudaSetDevice(ctx->device));
    mpb_mel* m = new mpb_mel();
    m->ctx = ctx; m->fft_len = fft_len; m->n_mag = n_mag; m->n_ph = n_ph; m->phase_dim = phase_dim;
    m->alpha_mag = alpha_mag; m->alpha_ph = alpha_ph;
    m->ld_mag = pad64(n_mag); m->ld_ph = pad64(n_ph);
    const int H = fft_len / 2 + 1;
    const size_t kpad = (size_t)((H + MEL_KSLICE - 1) / MEL_KSLICE) * MEL_KSLICE;
    double* scratch = nullptr;
    CU(cudaMalloc(&scratch, sizeof(double) * kpad * MEL_MAX_COEFFS));
    CU(cudaMalloc(&m->wt_mag, sizeof(float) * kpad * m->ld_mag));
    CU(cudaMalloc(&m->wt_ph, sizeof(float) * kpad * m->ld_ph));
    CU(cudaMalloc(&m->cos_mag, sizeof(double) * n_mag * n_mag));
    CU(cudaMalloc(&m->cos_ph, sizeof(double) * n_ph * phase_dim));
    CU(cudaMemcpy(m->cos_mag, cos_mag, sizeof(double) * n_mag * n_mag, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m->cos_ph, cos_ph, sizeof(double) * n_ph * phase_dim, cudaMemcpyHostToDevice));
    CU(build_warp_matrix(fft_len, n_mag, alpha_mag, m->wt_mag, scratch, m->ld_mag, ctx->stream));
    CU(build_warp_matrix(fft_len, n_ph, alpha_ph, m->wt_ph, scratch, m->ld_ph, ctx->stream));
    static const bool mel_tc = [] { const char* e = getenv("MPB_MEL_TC"); return !e || (atoi(e) & 1); }();   // default on; MPB_MEL_TC=0: FMA kernels
    if (mel_tc && m->ld_mag == 64 && m->ld_ph == 64) {       // tensor-core tile product with fused finish (mpb_mel_warp_tc.cu)
        CU(cudaMalloc(&m->wt_tc_mag, mel_tc_operand_bytes(fft_len)));
        CU(cudaMalloc(&m->wt_tc_ph, mel_tc_operand_bytes(fft_len)));
        CU(build_warp_matrix_tc(fft_len, m->wt_mag, m->ld_mag, m->wt_tc_mag, ctx->stream));
        CU(build_warp_matrix_tc(fft_len, m->wt_ph, m->ld_ph, m->wt_tc_ph, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(scratch));
    ctx->launches += 4;
    *out = m;
    return MPB_OK;
}

int mpb_mel_destroy(mpb_mel* m) {
    if (!m) return MPB_OK;
    cudaSetDevice(m->ctx->device);
    cudaFree(m->wt_mag); cudaFree(m->wt_ph); cudaFree(m->cos_mag); cudaFree(m->cos_ph);
    cudaFree(m->wt_tc_mag); cudaFree(m->wt_tc_ph);
    m->partial.release(); m->compact.release(); m->sig32.release();
    for (auto& b : m->feats) b.release();
    for (auto& b : m->small) b.release();
    delete m;
    return MPB_OK;
}

// debugging / tests: copies W^T (float32, [fft_len/2+1][n]) of stream 0 (mag) or 1 (phase) to the host
int mpb_mel_get_warp_matrix(mpb_mel* m, int which, float* out_host) {
    if (!m || !out_host) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    CU(cudaSetDevice(m->ctx->device));
    const int H = m->fft_len / 2 + 1;
    const int n = which == 0 ? m->n_mag : m->n_ph, ld = which == 0 ? m->ld_mag : m->ld_ph;
    const float* src = which == 0 ? m->wt_mag : m->wt_ph;
    CU(cudaMemcpy2D(out_host, sizeof(float) * n, src, sizeof(float) * ld, sizeof(float) * n, H, cudaMemcpyDeviceToHost));
    return MPB_OK;
}

struct LerpRows { const int32_t* r0; const int32_t* r1; const float* w; int raw_mc = 0; };   // device arrays, one entry per OUTPUT frame

static int mel_compress_impl(mpb_mel* m, void* stream, const void* mag, const void* real, const void* imag, int feat_dtype,
                             int pre_logp, const uint8_t* voi, int64_t nfrm, void* out_mag_mel, void* out_real_mel,
                             void* out_imag_mel, int out_dtype, LerpRows lerp = LerpRows{nullptr, nullptr, nullptr, 0});

int mpb_mel_compress_dev(mpb_mel* m, void* stream, const void* mag, const void* real, const void* imag, int feat_dtype,
                         const uint8_t* voi, int64_t nfrm, void* out_mag_mel, void* out_real_mel, void* out_imag_mel,
                         int out_dtype) {
    return mel_compress_impl(m, stream, mag, real, imag, feat_dtype, 0, voi, nfrm, out_mag_mel, out_real_mel, out_imag_mel,
                             out_dtype);
}

}  // extern "C"

static int mel_compress_impl(mpb_mel* m, void* stream, const void* mag, const void* real, const void* imag, int feat_dtype,
                             int pre_logp, const uint8_t* voi, int64_t nfrm, void* out_mag_mel, void* out_real_mel,
                             void* out_imag_mel, int out_dtype, LerpRows lerp) {
    if (!m) return fail(MPB_ERR_BAD_ARG, "plan is NULL");
    if (!dtype_ok(feat_dtype) || !dtype_ok(out_dtype)) return fail(MPB_ERR_BAD_ARG, "unknown dtype");
    if (nfrm < 0) return fail(MPB_ERR_BAD_ARG, "negative size");
    if (nfrm == 0) return MPB_OK;
    if (!mag || !real || !imag || !voi || !out_mag_mel || !out_real_mel || !out_imag_mel)
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    CU(cudaSetDevice(m->ctx->device));
    std::lock_guard<std::mutex> lk(m->mu);
    const int H = m->fft_len / 2 + 1;
    const int n_slices = (H - 1) / MEL_KSLICE;
    const int ncp = m->ld_mag > m->ld_ph ? m->ld_mag : m->ld_ph;
    const int64_t chunk = nfrm < MEL_CHUNK ? nfrm : MEL_CHUNK;
    CU(m->partial.need(sizeof(float) * 3 * (size_t)n_slices * (size_t)chunk * ncp));
    CU(m->compact.need(sizeof(int32_t) * (2 * (size_t)((chunk + 3) & ~(int64_t)3) + 4)));
    int32_t* d_vidx = (int32_t*)m->compact.p;
    int32_t* d_cidx = d_vidx + ((chunk + 3) & ~(int64_t)3);          // 16-byte aligned: vector stores in the scan
    int32_t* d_cnt = d_cidx + ((chunk + 3) & ~(int64_t)3);
    const size_t fes = feat_dtype == MPB_F64 ? 8 : 4, oes = out_dtype == MPB_F64 ? 8 : 4;
    const int od_mag = m->n_mag, od_ph = lerp.raw_mc ? m->n_ph : m->phase_dim;   // output row widths
    for (int64_t f0 = 0; f0 < nfrm; f0 += chunk) {
        const int64_t n = nfrm - f0 < chunk ? nfrm - f0 : chunk;
        MelArgs a;
        // with row interpolation the source rows are addressed through lerp.r0 / r1 (global), else frame by frame
        const int64_t src0 = lerp.r0 ? 0 : f0;
        a.mag = (const char*)mag + fes * src0 * H; a.real = (const char*)real + fes * src0 * H;
        a.imag = (const char*)imag + fes * src0 * H; a.feat_dtype = feat_dtype; a.pre_logp = pre_logp;
        a.lerp_r0 = lerp.r0 ? lerp.r0 + f0 : nullptr; a.lerp_r1 = lerp.r1 ? lerp.r1 + f0 : nullptr;
        a.lerp_w = lerp.w ? lerp.w + f0 : nullptr;
        a.raw_mc = lerp.raw_mc;
        a.voi = voi + f0; a.nfrm = n; a.fft_len = m->fft_len;
        a.wt_mag = m->wt_mag; a.ld_mag = m->ld_mag; a.wt_ph = m->wt_ph; a.ld_ph = m->ld_ph;
        a.cos_mag = m->cos_mag; a.n_mag = m->n_mag; a.cos_ph = m->cos_ph; a.n_ph = m->n_ph; a.phase_dim = m->phase_dim;
        a.partial = (float*)m->partial.p; a.ncp_max = ncp;
        a.out_mag = (char*)out_mag_mel + oes * f0 * od_mag;
        a.out_real = (char*)out_real_mel + oes * f0 * od_ph;
        a.out_imag = (char*)out_imag_mel + oes * f0 * od_ph;
        a.out_dtype = out_dtype;
        a.vidx = d_vidx; a.cidx = d_cidx; a.vcount = d_cnt;
        if (lerp.raw_mc) { a.vidx = nullptr; a.cidx = nullptr; a.vcount = nullptr; }   // every frame, every stream
        else LAUNCH(m->ctx, (cudaStream_t)stream, "k_voiced_compact",
                    launch_voiced_compact(a.voi, (int)n, d_vidx, d_cidx, d_cnt, (cudaStream_t)stream));
        LAUNCH(m->ctx, (cudaStream_t)stream, "k_mel_gemm", launch_mel_gemm(a, (cudaStream_t)stream));
        LAUNCH(m->ctx, (cudaStream_t)stream, "k_mel_finish", launch_mel_finish(a, (cudaStream_t)stream));
    }
    return MPB_OK;
}

extern "C" {

// analysis_compressed on the device for a batch of frames: signal + frame geometry (device pointers) in,
// low-dimensional features out.  Per chunk of frames: k_analysis (float64 butterflies) emits float32 log
// periodograms into a plan-owned scratch, k_mel_gemm / k_mel_finish reduce them to mag_dim + 2*phase_dim values.
int mpb_analysis_compressed_dev(mpb_mel* m, void* stream, const void* sig, int sig_dtype, int64_t n_sig,
                                const int64_t* centre, const int32_t* left, const int32_t* right, const uint8_t* voi,
                                int64_t nfrm, void* out_mag_mel, void* out_real_mel, void* out_imag_mel, int out_dtype) {
    if (!m) return fail(MPB_ERR_BAD_ARG, "plan is NULL");
    if (!dtype_ok(sig_dtype) || !dtype_ok(out_dtype) || nfrm < 0 || n_sig < 0) return fail(MPB_ERR_BAD_ARG, "bad dtype or size");
    if (nfrm == 0) return MPB_OK;
    if (!sig || !centre || !left || !right || !voi || !out_mag_mel || !out_real_mel || !out_imag_mel)
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    mpb_ctx* ctx = m->ctx;
    CU(cudaSetDevice(ctx->device));
    const int H = m->fft_len / 2 + 1;
    const int64_t chunk = nfrm < MEL_CHUNK ? nfrm : MEL_CHUNK;
    const int LP = lp_pitch_of(H);
    const bool use_tc = m->wt_tc_mag != nullptr;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        for (int i = 0; i < 3; ++i) CU(m->feats[i].need(sizeof(float) * (size_t)chunk * LP));
        if (use_tc) {
            CU(m->compact.need(sizeof(int32_t) * (2 * (size_t)((chunk + 3) & ~(int64_t)3) + 4)));
            CU(m->partial.need(sizeof(float) * 3 * (size_t)chunk * 64));      // float32 mel cepstra between the two kernels
        }
    }
    // float64 butterflies (default); MPB_LOGP_F32=1 selects the float32 engine for this fused path only (experiment, see
    // launch_analysis_logp)
    static const int logp_compute = [] { const char* e = getenv("MPB_LOGP_F32"); return (e && atoi(e) > 0) ? MPB_F32 : MPB_F64; }();
    const void* tw = nullptr;
    int rc = get_twiddles(ctx, m->fft_len, logp_compute, &tw);
    if (rc != MPB_OK) return rc;
    const size_t oes = out_dtype == MPB_F64 ? 8 : 4;
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t f0 = 0; f0 < nfrm; f0 += chunk) {
        const int64_t n = nfrm - f0 < chunk ? nfrm - f0 : chunk;
        AnalysisArgs a;
        a.sig = sig; a.sig_dtype = sig_dtype; a.n_sig = n_sig;
        a.centre = centre + f0; a.left = left + f0; a.right = right + f0; a.win = nullptr;
        a.nfrm = n; a.fft_len = m->fft_len; a.compute_dtype = logp_compute; a.tw = tw;
        a.out_a = m->feats[0].p; a.out_b = m->feats[1].p; a.out_c = m->feats[2].p; a.out_dtype = MPB_F32;
        a.mode = MODE_LOGP; a.num_sms = ctx->num_sms; a.ph_mask = voi + f0;
        if (use_tc) {
            // tensor-core product with the finish step in its epilogue: 16-byte pitched rows, the phase rows compacted to the
            // voiced frames (rank from k_voiced_compact), nothing but the low-dimensional features leaves the second kernel
            std::lock_guard<std::mutex> lk(m->mu);
            int32_t* d_vidx = (int32_t*)m->compact.p;
            int32_t* d_cidx = d_vidx + ((chunk + 3) & ~(int64_t)3);          // 16-byte aligned: vector stores in the scan
            int32_t* d_cnt = d_cidx + ((chunk + 3) & ~(int64_t)3);
            LAUNCH(ctx, st, "k_voiced_compact", launch_voiced_compact(voi + f0, (int)n, d_vidx, d_cidx, d_cnt, st));
            a.row_pitch = LP; a.ph_row = d_cidx;
            LAUNCH(ctx, st, "k_analysis<logp>", launch_analysis_logp(a, st));
            MelArgs g;
            g.mag = m->feats[0].p; g.real = m->feats[1].p; g.imag = m->feats[2].p; g.feat_dtype = MPB_F32; g.pre_logp = 1; g.raw_mc = 0;
            g.voi = voi + f0; g.nfrm = n; g.fft_len = m->fft_len;
            g.wt_mag = m->wt_mag; g.ld_mag = m->ld_mag; g.wt_ph = m->wt_ph; g.ld_ph = m->ld_ph;
            g.cos_mag = m->cos_mag; g.n_mag = m->n_mag; g.cos_ph = m->cos_ph; g.n_ph = m->n_ph; g.phase_dim = m->phase_dim;
            g.partial = (float*)m->partial.p; g.ncp_max = 64;
            g.out_mag = (char*)out_mag_mel + oes * f0 * m->n_mag; g.out_real = (char*)out_real_mel + oes * f0 * m->phase_dim;
            g.out_imag = (char*)out_imag_mel + oes * f0 * m->phase_dim; g.out_dtype = out_dtype;
            g.vidx = d_vidx; g.cidx = d_cidx; g.vcount = d_cnt;
            g.lerp_r0 = nullptr; g.lerp_r1 = nullptr; g.lerp_w = nullptr;
            g.wt_tc_mag = m->wt_tc_mag; g.wt_tc_ph = m->wt_tc_ph; g.lp_pitch = LP; g.num_sms = ctx->num_sms;
            if (!mel_tc_usable(g)) return fail(MPB_ERR_INTERNAL, "tensor-core mel product not usable for this plan");
            LAUNCH(ctx, st, "k_mel_warp_tc", launch_mel_warp_tc(g, st));
            LAUNCH(ctx, st, "k_mel_cos", launch_mel_cos(g, st));
            continue;
        }
        LAUNCH(ctx, st, "k_analysis<logp>", launch_analysis_logp(a, st));
        rc = mel_compress_impl(m, stream, m->feats[0].p, m->feats[1].p, m->feats[2].p, MPB_F32, 1, voi + f0, n,
                               (char*)out_mag_mel + oes * f0 * m->n_mag, (char*)out_real_mel + oes * f0 * m->phase_dim,
                               (char*)out_imag_mel + oes * f0 * m->phase_dim, out_dtype);
        if (rc != MPB_OK) return rc;
    }
    return MPB_OK;
}

int mpb_mel_compress_host(mpb_mel* m, const double* mag, const double* real, const double* imag, const uint8_t* voi,
                          int64_t nfrm, double* out_mag_mel, double* out_real_mel, double* out_imag_mel) {
    if (!m) return fail(MPB_ERR_BAD_ARG, "plan is NULL");
    if (nfrm == 0) return MPB_OK;
    if (!mag || !real || !imag || !voi || !out_mag_mel || !out_real_mel || !out_imag_mel)
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    mpb_ctx* ctx = m->ctx;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int H = m->fft_len / 2 + 1;
    const size_t fsz = sizeof(double) * (size_t)nfrm * H;
    cudaStream_t st = ctx->stream;
    for (int i = 0; i < 3; ++i) CU(m->feats[i].need(fsz));
    CU(m->small[0].need((size_t)nfrm));
    CU(m->small[1].need(sizeof(double) * nfrm * m->n_mag));
    CU(m->small[2].need(sizeof(double) * nfrm * m->phase_dim));
    CU(m->small[3].need(sizeof(double) * nfrm * m->phase_dim));
    CU(cudaMemcpyAsync(m->feats[0].p, mag, fsz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->feats[1].p, real, fsz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->feats[2].p, imag, fsz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->small[0].p, voi, (size_t)nfrm, cudaMemcpyHostToDevice, st));
    int rc = mpb_mel_compress_dev(m, st, m->feats[0].p, m->feats[1].p, m->feats[2].p, MPB_F64, (const uint8_t*)m->small[0].p,
                                  nfrm, m->small[1].p, m->small[2].p, m->small[3].p, MPB_F64);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpyAsync(out_mag_mel, m->small[1].p, sizeof(double) * nfrm * m->n_mag, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_real_mel, m->small[2].p, sizeof(double) * nfrm * m->phase_dim, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_imag_mel, m->small[3].p, sizeof(double) * nfrm * m->phase_dim, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MPB_OK;
}

int mpb_analysis_compressed_hostv(mpb_mel* m, const double* const* sigs, const int64_t* sig_lens, int32_t n_sigs,
                                  const int64_t* centre, const int32_t* left, const int32_t* right, const uint8_t* voi,
                                  int64_t nfrm, int compute_dtype, double* out_mag_mel, double* out_real_mel,
                                  double* out_imag_mel);
int mpb_analysis_compressed_hostv2(mpb_mel* m, const void* const* sigs, int sig_dtype, const int64_t* sig_lens, int32_t n_sigs,
                                   const int64_t* centre, const int32_t* left, const int32_t* right, const uint8_t* voi,
                                   int64_t nfrm, void* out_mag_mel, void* out_real_mel, void* out_imag_mel, int out_dtype);

// analysis_lossless + format_for_modelling fused on the device: host signal in, low-dimensional features out.
// The lossless features only ever exist as a float32 scratch in HBM (frame-chunked).
int mpb_analysis_compressed_host(mpb_mel* m, const double* sig, int64_t n_sig, const int64_t* centre,
                                 const int32_t* left, const int32_t* right, const uint8_t* voi, int64_t nfrm,
                                 int compute_dtype, double* out_mag_mel, double* out_real_mel, double* out_imag_mel) {
    const double* sigs[1] = {sig};
    const int64_t lens[1] = {n_sig};
    if (!sig && nfrm > 0) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    return mpb_analysis_compressed_hostv(m, sigs, lens, 1, centre, left, right, voi, nfrm, compute_dtype, out_mag_mel,
                                         out_real_mel, out_imag_mel);
}

// Same, with the utterances given as separate HOST arrays (no concatenation on the caller's side): centre[] still
// indexes the virtual concatenation sigs[0] | sigs[1] | ...
//
// The call is a three-stage pipeline over groups of utterances: while the host threads narrow group g+1 into the
// page-locked staging buffer and its samples cross PCIe on stream_in, group g runs its kernels on the compute stream
// and the features of group g-1 return on stream_out.  Events hand the groups from stage to stage; the scratch of the
// kernels is shared, which is safe because the compute stream serialises them.
int mpb_analysis_compressed_hostv(mpb_mel* m, const double* const* sigs, const int64_t* sig_lens, int32_t n_sigs,
                                  const int64_t* centre, const int32_t* left, const int32_t* right, const uint8_t* voi,
                                  int64_t nfrm, int compute_dtype, double* out_mag_mel, double* out_real_mel,
                                  double* out_imag_mel) {
    (void)compute_dtype;   // the fused path always runs float64 butterflies
    return mpb_analysis_compressed_hostv2(m, (const void* const*)sigs, MPB_F64, sig_lens, n_sigs, centre, left, right, voi, nfrm,
                                          out_mag_mel, out_real_mel, out_imag_mel, MPB_F64);
}

// The same pipeline with the caller's own element types: signals as float64, float32 or PCM16 (MPB_I16: scaled by 1/32768 on
// the device, what sf.read returns for the reference's wav files), features as float64 or float32 (the reference's feature
// files are float32, src/libutils.py:122-127).  Narrow types halve or quarter the bytes on PCIe and skip the host-side
// float64 -> float32 narrowing pass.
int mpb_analysis_compressed_hostv2(mpb_mel* m, const void* const* sigs, int sig_dtype, const int64_t* sig_lens, int32_t n_sigs,
                                   const int64_t* centre, const int32_t* left, const int32_t* right, const uint8_t* voi,
                                   int64_t nfrm, void* out_mag_mel, void* out_real_mel, void* out_imag_mel, int out_dtype) {
    if (!sig_dtype_ok(sig_dtype) || !dtype_ok(out_dtype)) return fail(MPB_ERR_BAD_ARG, "unknown dtype");
    if (!m) return fail(MPB_ERR_BAD_ARG, "plan is NULL");
    if (nfrm == 0) return MPB_OK;
    if (!sigs || !sig_lens || n_sigs < 1 || !centre || !left || !right || !voi || !out_mag_mel || !out_real_mel ||
        !out_imag_mel)
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    int64_t n_sig = 0;
    for (int32_t i = 0; i < n_sigs; ++i) {
        if (!sigs[i] || sig_lens[i] < 0) return fail(MPB_ERR_BAD_ARG, "NULL signal");
        n_sig += sig_lens[i];
    }
    mpb_ctx* ctx = m->ctx;
    int rc = check_frames_host(centre, left, right, nfrm, n_sig, m->fft_len);
    if (rc != MPB_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    static const bool trace = getenv("MPB_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return 1e3 * std::chrono::duration<double>(b - a).count();
    };
    const auto t0 = now();
    cudaStream_t s_in = ctx->stream_in, s_cmp = ctx->stream, s_out = ctx->stream_out;

    // ---- groups of whole utterances with about the same number of frames (frames never straddle utterances) ----
    int n_groups = pipeline_groups(nfrm, 5000, 8);
    if (n_groups > n_sigs) n_groups = n_sigs;
    std::vector<int32_t> group_end((size_t)n_groups);          // one past the last signal of the group
    std::vector<int64_t> group_frm((size_t)n_groups + 1, 0);   // frame range of the group
    {
        int64_t f = 0, off = 0;
        int32_t g = 0;
        for (int32_t i = 0; i < n_sigs; ++i) {
            off += sig_lens[i];
            while (f < nfrm && centre[f] < off) ++f;           // frames are ordered by utterance
            const int32_t left_sigs = n_sigs - 1 - i, left_groups = n_groups - 1 - g;
            if (i == n_sigs - 1 || (left_groups > 0 && (f * n_groups >= nfrm * (g + 1) || left_sigs == left_groups))) {
                group_end[g] = i + 1;
                group_frm[g + 1] = f;
                if (++g == n_groups) break;
            }
        }
        n_groups = g;
        group_frm[n_groups] = nfrm;
    }

    CU(m->small[0].need((size_t)nfrm));
    const size_t oes = out_dtype == MPB_F64 ? 8 : 4;
    CU(m->small[1].need(oes * nfrm * m->n_mag));
    CU(m->small[2].need(oes * nfrm * m->phase_dim));
    CU(m->small[3].need(oes * nfrm * m->phase_dim));
    CU(m->small[4].need((sig_dtype == MPB_F64 ? sizeof(double) : sizeof(int16_t)) * (size_t)n_sig));   // float64 fall-back / PCM16 landing zone
    CU(m->small[5].need(sizeof(int64_t) * nfrm));
    CU(m->small[6].need(sizeof(int32_t) * nfrm));
    CU(m->small[7].need(sizeof(int32_t) * nfrm));
    CU(m->sig32.need(sizeof(float) * n_sig));
    {   // size the kernels' scratch for the largest group up front: a grow in mid-pipeline would free live buffers
        int64_t big = 0;
        for (int g = 0; g < n_groups; ++g) big = std::max(big, group_frm[g + 1] - group_frm[g]);
        rc = mel_reserve(m, big);
        if (rc != MPB_OK) return rc;
    }
    // frame descriptors: one page-locked block, one copy (pageable copies would stall the enqueueing thread)
    const size_t d_bytes = (sizeof(int64_t) + 2 * sizeof(int32_t) + 1) * (size_t)nfrm;
    CU(ctx->desc_stage.need(d_bytes));
    {
        char* h = (char*)ctx->desc_stage.p;
        memcpy(h, centre, sizeof(int64_t) * nfrm);
        memcpy(h + 8 * nfrm, left, sizeof(int32_t) * nfrm);
        memcpy(h + 12 * nfrm, right, sizeof(int32_t) * nfrm);
        memcpy(h + 16 * nfrm, voi, (size_t)nfrm);
        CU(cudaMemcpyAsync(m->small[5].p, h, sizeof(int64_t) * nfrm, cudaMemcpyHostToDevice, s_in));
        CU(cudaMemcpyAsync(m->small[6].p, h + 8 * nfrm, sizeof(int32_t) * nfrm, cudaMemcpyHostToDevice, s_in));
        CU(cudaMemcpyAsync(m->small[7].p, h + 12 * nfrm, sizeof(int32_t) * nfrm, cudaMemcpyHostToDevice, s_in));
        CU(cudaMemcpyAsync(m->small[0].p, h + 16 * nfrm, (size_t)nfrm, cudaMemcpyHostToDevice, s_in));
    }
    const auto t1 = now();
    std::vector<cudaEvent_t> evs;
    PipelineDrain drain(ctx, &evs);                // every exit path below drains the streams first
    auto on_group = [&](int32_t g, int dtype) -> int {
        cudaEvent_t e_in = get_event(ctx), e_cmp = get_event(ctx);
        evs.push_back(e_in); evs.push_back(e_cmp);
        CU(cudaEventRecord(e_in, s_in));
        CU(cudaStreamWaitEvent(s_cmp, e_in, 0));
        const int64_t fa = group_frm[g], n = group_frm[g + 1] - fa;
        if (n > 0) {
            int r = mpb_analysis_compressed_dev(
                m, s_cmp, dtype == MPB_F32 ? m->sig32.p : m->small[4].p, dtype, n_sig, (const int64_t*)m->small[5].p + fa,
                (const int32_t*)m->small[6].p + fa, (const int32_t*)m->small[7].p + fa, (const uint8_t*)m->small[0].p + fa, n,
                (char*)m->small[1].p + oes * fa * m->n_mag, (char*)m->small[2].p + oes * fa * m->phase_dim,
                (char*)m->small[3].p + oes * fa * m->phase_dim, out_dtype);
            if (r != MPB_OK) return r;
        }
        CU(cudaEventRecord(e_cmp, s_cmp));
        CU(cudaStreamWaitEvent(s_out, e_cmp, 0));
        if (n > 0) {
            CU(cudaMemcpyAsync((char*)out_mag_mel + oes * fa * m->n_mag, (char*)m->small[1].p + oes * fa * m->n_mag,
                               oes * n * m->n_mag, cudaMemcpyDeviceToHost, s_out));
            CU(cudaMemcpyAsync((char*)out_real_mel + oes * fa * m->phase_dim, (char*)m->small[2].p + oes * fa * m->phase_dim,
                               oes * n * m->phase_dim, cudaMemcpyDeviceToHost, s_out));
            CU(cudaMemcpyAsync((char*)out_imag_mel + oes * fa * m->phase_dim, (char*)m->small[3].p + oes * fa * m->phase_dim,
                               oes * n * m->phase_dim, cudaMemcpyDeviceToHost, s_out));
        }
        return MPB_OK;
    };
    if (sig_dtype == MPB_F64)
        rc = upload_signal_groups(ctx, s_in, (const double* const*)sigs, sig_lens, n_sigs, group_end.data(), n_groups, m->sig32.p,
                                  m->small[4].p, on_group);
    else
        rc = upload_signal_groups_narrow(ctx, s_in, sigs, sig_dtype, sig_lens, n_sigs, group_end.data(), n_groups, m->sig32.p,
                                         m->small[4].p, on_group);
    const auto t2 = now();
    // drain all three stages even after an error: host and staging buffers must not be reused under a live copy
    cudaError_t e1 = host_wait(ctx, s_in), e2 = host_wait(ctx, s_cmp), e3 = host_wait(ctx, s_out);
    if (rc != MPB_OK) return rc;
    CU(e1); CU(e2); CU(e3);
    if (trace)
        fprintf(stderr, "[mpb] analysis_compressed_hostv: %d groups, prepare %.3f ms, stage+enqueue %.3f ms, drain %.3f ms\n",
                n_groups, ms(t0, t1), ms(t1, t2), ms(t2, now()));
    return MPB_OK;
}

// analysis_compressed(b_const_rate=True) (src/magphase.py:2966-2983): variable-rate lossless analysis on the device
// (float32 rows, scratch), then every constant-rate output frame f is the linear interpolation
// (1-w)*row[r0[f]] + w*row[r1[f]] of two analysed frames -- interpolated inside the tile-product loader, before the
// log, exactly where interp1d sits in the reference -- and goes through the same mel compression.
// lerp_r0 / lerp_r1 / lerp_w / voi_out: HOST arrays, one entry per output frame.
int mpb_analysis_compressed_const_hostv(mpb_mel* m, const double* const* sigs, const int64_t* sig_lens, int32_t n_sigs,
                                        const int64_t* centre, const int32_t* left, const int32_t* right, int64_t nfrm,
                                        const int32_t* lerp_r0, const int32_t* lerp_r1, const float* lerp_w,
                                        const uint8_t* voi_out, int64_t n_out, double* out_mag_mel, double* out_real_mel,
                                        double* out_imag_mel) {
    if (!m) return fail(MPB_ERR_BAD_ARG, "plan is NULL");
    if (n_out == 0) return MPB_OK;
    if (!sigs || !sig_lens || n_sigs < 1 || !centre || !left || !right || !lerp_r0 || !lerp_r1 || !lerp_w || !voi_out ||
        !out_mag_mel || !out_real_mel || !out_imag_mel || nfrm < 1)
        return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    int64_t n_sig = 0;
    for (int32_t i = 0; i < n_sigs; ++i) {
        if (!sigs[i] || sig_lens[i] < 0) return fail(MPB_ERR_BAD_ARG, "NULL signal");
        n_sig += sig_lens[i];
    }
    for (int64_t f = 0; f < n_out; ++f)
        if (lerp_r0[f] < 0 || lerp_r0[f] >= nfrm || lerp_r1[f] < 0 || lerp_r1[f] >= nfrm)
            return fail(MPB_ERR_BAD_ARG, "interpolation row index out of range");
    mpb_ctx* ctx = m->ctx;
    int rc = check_frames_host(centre, left, right, nfrm, n_sig, m->fft_len);
    if (rc != MPB_OK) return rc;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int H = m->fft_len / 2 + 1;
    cudaStream_t st = ctx->stream;
    DevBuf* b = ctx->scratch;     // ctx scratch: lossless rows + descriptors; plan scratch: outputs
    for (int i = 5; i < 8; ++i) CU(b[i].need(sizeof(float) * (size_t)nfrm * H));
    CU(b[0].need(sizeof(double) * n_sig));
    CU(b[1].need(sizeof(int64_t) * nfrm)); CU(b[2].need(sizeof(int32_t) * nfrm)); CU(b[3].need(sizeof(int32_t) * nfrm));
    CU(m->small[0].need((size_t)n_out));
    CU(m->small[1].need(sizeof(double) * n_out * m->n_mag));
    CU(m->small[2].need(sizeof(double) * n_out * m->phase_dim));
    CU(m->small[3].need(sizeof(double) * n_out * m->phase_dim));
    CU(m->small[5].need(sizeof(int32_t) * n_out)); CU(m->small[6].need(sizeof(int32_t) * n_out));
    CU(m->small[7].need(sizeof(float) * n_out));
    int sig_dtype = MPB_F64;
    rc = upload_signals(ctx, st, sigs, sig_lens, n_sigs, b[0].p, &sig_dtype);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpyAsync(b[1].p, centre, sizeof(int64_t) * nfrm, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[2].p, left, sizeof(int32_t) * nfrm, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b[3].p, right, sizeof(int32_t) * nfrm, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->small[0].p, voi_out, (size_t)n_out, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->small[5].p, lerp_r0, sizeof(int32_t) * n_out, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->small[6].p, lerp_r1, sizeof(int32_t) * n_out, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->small[7].p, lerp_w, sizeof(float) * n_out, cudaMemcpyHostToDevice, st));
    rc = analysis_common(ctx, st, b[0].p, sig_dtype, n_sig, (const int64_t*)b[1].p, (const int32_t*)b[2].p, (const int32_t*)b[3].p,
                         nullptr, nfrm, m->fft_len, MPB_F64, b[5].p, b[6].p, b[7].p, MPB_F32, MODE_FEATS);
    if (rc != MPB_OK) return rc;
    LerpRows lr{(const int32_t*)m->small[5].p, (const int32_t*)m->small[6].p, (const float*)m->small[7].p};
    rc = mel_compress_impl(m, st, b[5].p, b[6].p, b[7].p, MPB_F32, 0, (const uint8_t*)m->small[0].p, n_out, m->small[1].p,
                           m->small[2].p, m->small[3].p, MPB_F64, lr);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpyAsync(out_mag_mel, m->small[1].p, sizeof(double) * n_out * m->n_mag, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_real_mel, m->small[2].p, sizeof(double) * n_out * m->phase_dim, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_imag_mel, m->small[3].p, sizeof(double) * n_out * m->phase_dim, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MPB_OK;
}

// la.sp_to_mcep (src/libaudio.py:575-601) for three spectra at once: the float32-rounded mel cepstra of
// a (in_type 3, |X|), b and c (in_type 2, ln|X|; feed x*ln(10)/20 for in_type 1 "dB" input), HOST float64 rows of
// fft_len/2+1 bins.  out_a: nfrm x n_mag, out_b / out_c: nfrm x n_ph.  Any of b / c may alias a.
int mpb_sp_to_mcep_host(mpb_mel* m, const double* a, const double* b, const double* c, int64_t nfrm, double* out_a,
                        double* out_b, double* out_c) {
    if (!m) return fail(MPB_ERR_BAD_ARG, "plan is NULL");
    if (nfrm == 0) return MPB_OK;
    if (!a || !b || !c || !out_a || !out_b || !out_c) return fail(MPB_ERR_BAD_ARG, "NULL buffer");
    mpb_ctx* ctx = m->ctx;
    CU(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int H = m->fft_len / 2 + 1;
    const size_t fsz = sizeof(double) * (size_t)nfrm * H;
    cudaStream_t st = ctx->stream;
    for (int i = 0; i < 3; ++i) CU(m->feats[i].need(fsz));
    CU(m->small[0].need((size_t)nfrm));
    CU(m->small[1].need(sizeof(double) * nfrm * m->n_mag));
    CU(m->small[2].need(sizeof(double) * nfrm * m->n_ph));
    CU(m->small[3].need(sizeof(double) * nfrm * m->n_ph));
    CU(cudaMemcpyAsync(m->feats[0].p, a, fsz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->feats[1].p, b, fsz, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m->feats[2].p, c, fsz, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(m->small[0].p, 1, (size_t)nfrm, st));
    LerpRows lr{nullptr, nullptr, nullptr, 1};
    int rc = mel_compress_impl(m, st, m->feats[0].p, m->feats[1].p, m->feats[2].p, MPB_F64, 0, (const uint8_t*)m->small[0].p,
                               nfrm, m->small[1].p, m->small[2].p, m->small[3].p, MPB_F64, lr);
    if (rc != MPB_OK) return rc;
    CU(cudaMemcpyAsync(out_a, m->small[1].p, sizeof(double) * nfrm * m->n_mag, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_b, m->small[2].p, sizeof(double) * nfrm * m->n_ph, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_c, m->small[3].p, sizeof(double) * nfrm * m->n_ph, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MPB_OK;
}

}  // extern "C"

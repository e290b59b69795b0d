/*
 * magphase_b200 -- C ABI of the B200-native MagPhase analysis/synthesis hot path.
 *
 * The reference (CSTR-Edinburgh/magphase) has no FFI: its boundary is the Python module API of
 * src/magphase.py.  This header is what a ctypes binding in that module would bind instead of the
 * NumPy loops; each entry point names the reference interface it replaces (file:line, relative to
 * the reference root).  See INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *   - every function returns an int status: MPB_OK (0) or a negative MPB_ERR_* code; nothing throws.
 *     mpb_last_error() gives the message for the last failure on the calling thread.
 *   - the caller allocates and owns every buffer.  Functions ending in _dev take DEVICE pointers and
 *     enqueue on `stream` (a cudaStream_t passed as void*; NULL = default stream) without
 *     synchronising.  Functions ending in _host take HOST pointers (NumPy arrays), stage through
 *     pinned memory, run the same kernels and return after the results are back on the host.
 *   - matrices are C-contiguous, row = frame, exactly like the reference's NumPy arrays
 *     (nfrms x (fft_len/2+1) for mag/real/imag).
 *   - integer frame bookkeeping (rounding / truncation of pitch marks, cumsum) is done by the host
 *     mirror in float64 NumPy with the reference's expression order and handed over as int arrays.
 */
#ifndef MAGPHASE_B200_H
#define MAGPHASE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPB_OK 0
#define MPB_ERR_BAD_ARG (-1)      /* NULL pointer, negative size, unknown dtype ...   -> ValueError */
#define MPB_ERR_FFT_LEN (-2)      /* fft_len not one of 1024 / 2048 / 4096            -> ValueError */
#define MPB_ERR_FRAME_GEOM (-3)   /* frame outside the signal, marks not increasing ... -> ValueError */
#define MPB_ERR_CUDA (-4)         /* CUDA runtime failure (message has the cudaError string)        */
#define MPB_ERR_NO_DEVICE (-5)    /* no usable CUDA device: there is NO CPU fallback                */
#define MPB_ERR_DIM (-6)          /* mel dimensions exceed the compiled limits                      */
#define MPB_ERR_INTERNAL (-7)     /* a self-check of the library failed (never expected)            */

/* element types of caller buffers */
#define MPB_F32 0
#define MPB_F64 1
#define MPB_I16 2   /* host signals only: PCM16 samples, scaled by 1/32768 on the device (what sf.read returns, src/libaudio.py:343-350) */

/* window applied to each side of a pitch-synchronous frame (mpb_frames_*) */
#define MPB_WIN_HANN 0            /* np.hanning halves               src/libaudio.py:70-84 */
#define MPB_WIN_BARTLETT25 1      /* np.bartlett(.)**2.5 halves      src/magphase.py:67-69 */
#define MPB_WIN_RECT 2            /* weight 1: the frame's samples already carry their window.  How an arbitrary win_func
                                     callable (src/magphase.py:102-108) reaches the kernels: the caller lays the windowed
                                     frames sig[P[f] .. P[f+2]] * gen_non_symmetric_win(l, r, win_func) back to back in
                                     `sig` and points centre[f] at the frame's own mark inside that buffer */

typedef struct mpb_ctx mpb_ctx;

/* ---- context ------------------------------------------------------------------------------- */
/* Creates the per-device context: twiddle tables, mel/unwarp matrices cache, pinned staging.      */
int mpb_create(int device, mpb_ctx** out_ctx);
int mpb_destroy(mpb_ctx* ctx);
const char* mpb_last_error(void);
const char* mpb_version(void);
/* number of kernels this library has launched since the context was created (bench accounting)   */
int64_t mpb_launch_count(const mpb_ctx* ctx);
/* Optional per-kernel device timing: between begin and end every kernel launch of this context is bracketed
 * by CUDA events on its own stream; end synchronises and writes "name count total_ms\n" lines into buf.     */
int mpb_profile_begin(mpb_ctx* ctx);
int mpb_profile_end(mpb_ctx* ctx, char* buf, int64_t buf_len);

/* Measured non-tensor FMA peak of the device in TFLOP/s (2 flops per FMA) for MPB_F32 / MPB_F64: register-resident FMA
 * chains on every SM, best of three timed repetitions.  The roofline denominator of the compute-bound kernels
 * (the reference publishes no throughput; SURVEY.md 8(d) asks for a measured figure).                             */
int mpb_measure_fma_peak(mpb_ctx* ctx, int dtype, double* tflops);

/* Page-locked host buffers for result arrays (device->host copies into them run at full PCIe rate).            */
int mpb_host_alloc(mpb_ctx* ctx, int64_t bytes, void** out);
int mpb_host_free(mpb_ctx* ctx, void* p);

/* ---- lossless analysis --------------------------------------------------------------------- */
/*
 * Replaces windowing() + the pad/rotate loop + np.fft.fft + remove_hermitian_half of
 * analysis_with_del_comp_from_pm (src/magphase.py:74-119, :266-334) and compute_lossless_feats
 * (:457-476) for a batch of frames.
 *
 *   sig        concatenated signals of all utterances in the batch (sig_dtype)
 *   n_sig      total samples in sig
 *   centre[f]  absolute index into sig of frame f's pitch mark  (P[f+1] + utterance base)
 *   left[f]    P[f+1]-P[f]   (= v_shift[f]);   right[f] = P[f+2]-P[f+1]
 *   win[f]     MPB_WIN_* per frame, or NULL for all-Hann
 *   out_*      nfrm x (fft_len/2+1), out_dtype; mag=|X|, real=Re X/|X|, imag=Im X/|X| (0 where |X|==0)
 *   compute_dtype  MPB_F64: float64 butterflies (needed for 1e-5 on real/imag of near-silent bins)
 *                  MPB_F32: float32 butterflies
 */
int mpb_analysis_lossless_dev(mpb_ctx* ctx, void* stream,
                              const void* sig, int sig_dtype, int64_t n_sig,
                              const int64_t* centre, const int32_t* left, const int32_t* right,
                              const uint8_t* win, int64_t nfrm, int fft_len, int compute_dtype,
                              void* out_mag, void* out_real, void* out_imag, int out_dtype);

int mpb_analysis_lossless_host(mpb_ctx* ctx,
                               const double* sig, int64_t n_sig,
                               const int64_t* centre, const int32_t* left, const int32_t* right,
                               const uint8_t* win, int64_t nfrm, int fft_len, int compute_dtype,
                               double* out_mag, double* out_real, double* out_imag);

/* mpb_analysis_lossless_host with the caller's element types: signal MPB_F64 / MPB_F32 / MPB_I16 (PCM16 as the wav file holds
 * it; the device applies sf.read's 1/32768, src/libaudio.py:343-350), feature matrices MPB_F64 / MPB_F32 (the reference's
 * .mag/.real/.imag files are float32, src/libutils.py:122-127).                                                   */
int mpb_analysis_lossless_host2(mpb_ctx* ctx,
                                const void* sig, int sig_dtype, int64_t n_sig,
                                const int64_t* centre, const int32_t* left, const int32_t* right,
                                const uint8_t* win, int64_t nfrm, int fft_len, int compute_dtype,
                                void* out_mag, void* out_real, void* out_imag, int out_dtype);

/* Complex half spectra only (m_fft of analysis_with_del_comp_from_pm, src/magphase.py:325-332):
 * out_fft is nfrm x (fft_len/2+1) x 2 (re, im interleaved == complex64/complex128).              */
int mpb_frames_fft_dev(mpb_ctx* ctx, void* stream,
                       const void* sig, int sig_dtype, int64_t n_sig,
                       const int64_t* centre, const int32_t* left, const int32_t* right,
                       const uint8_t* win, int64_t nfrm, int fft_len, int compute_dtype,
                       void* out_fft, int out_dtype);

/* Host variant of mpb_frames_fft_dev: out_fft is nfrm x (fft_len/2+1) complex128.                */
int mpb_frames_fft_host(mpb_ctx* ctx,
                        const double* sig, int64_t n_sig,
                        const int64_t* centre, const int32_t* left, const int32_t* right,
                        const uint8_t* win, int64_t nfrm, int fft_len, int compute_dtype,
                        double* out_fft);

/* ---- lossless synthesis -------------------------------------------------------------------- */
/*
 * Host-side planner for the overlap-add: splits every utterance's frames into "runs" of about
 * target_frames consecutive frames, each spanning at least fft_len samples (so that at most two runs
 * ever touch the same output sample and the result is bit-reproducible).  All pointers are HOST
 * pointers.  A run is 4 x int32: {first frame (global index), frame count, utterance, flags
 * (bit0: has a previous run in the utterance, bit1: has a next run)}.  Call with out_runs = NULL
 * to get the count.  Also validates that pm is strictly increasing inside every utterance.
 */
int mpb_plan_ola_runs(const int32_t* pm, const int64_t* utt_frm_off, int32_t n_utt, int fft_len,
                      int32_t target_frames, int32_t* out_runs, int64_t capacity, int64_t* n_runs);

/*
 * Replaces synthesis_from_lossless (src/magphase.py:1759-1776): normalise real+j*imag, scale by mag,
 * Hermitian inverse FFT (la.add_hermitian_half src/libaudio.py:369-388 + np.fft.ifft), fftshift and
 * ola() (src/magphase.py:34-62) for a batch of utterances.
 *
 *   mag/real/imag  nfrm_total x (fft_len/2+1), feat_dtype, frames of all utterances concatenated
 *   pm[f]          integer pitch-mark position of frame f inside its utterance: trunc(cumsum(shift))
 *   utt_out_off    [n_utt+1] first output sample of each utterance in `out`
 *   utt_t0[u]      position (same axis as pm) of the utterance's first output sample:
 *                  out[utt_out_off[u] + j] = sum_i frame_i[(t0 + j) - pm_i + N/2]
 *                  (the reference's cut v_sig[(N/2 - pm[0]):][:pm[-1]+shift[-1]+1] is t0 = 0)
 *   runs, n_runs   the plan from mpb_plan_ola_runs, uploaded by the caller (device pointer)
 *   out            n_out = utt_out_off[n_utt] samples, out_dtype; fully overwritten
 */
int mpb_synthesis_lossless_dev(mpb_ctx* ctx, void* stream,
                               const void* mag, const void* real, const void* imag, int feat_dtype,
                               const int32_t* pm, int64_t nfrm_total,
                               const int64_t* utt_out_off, const int32_t* utt_t0, int32_t n_utt,
                               const int32_t* runs, int32_t n_runs, int fft_len, int compute_dtype,
                               void* out, int out_dtype, int64_t n_out);

/* Host variant: plans the runs itself; utt_frm_off is [n_utt+1] first frame of each utterance.     */
int mpb_synthesis_lossless_host(mpb_ctx* ctx,
                                const double* mag, const double* real, const double* imag,
                                const int32_t* pm, int64_t nfrm_total,
                                const int64_t* utt_frm_off, const int64_t* utt_out_off, const int32_t* utt_t0,
                                int32_t n_utt, int fft_len, int compute_dtype,
                                double* out, int64_t n_out);

/* mpb_synthesis_lossless_host with float64 or float32 feature matrices (feat_dtype) and waveform (out_dtype).     */
int mpb_synthesis_lossless_host2(mpb_ctx* ctx,
                                 const void* mag, const void* real, const void* imag, int feat_dtype,
                                 const int32_t* pm, int64_t nfrm_total,
                                 const int64_t* utt_frm_off, const int64_t* utt_out_off, const int32_t* utt_t0,
                                 int32_t n_utt, int fft_len, int compute_dtype,
                                 void* out, int out_dtype, int64_t n_out);

/* ---- low-dimensional compression (analysis side) --------------------------------------------- */
typedef struct mpb_mel mpb_mel;
/*
 * Plan for format_for_modelling (src/magphase.py:2490-2544): la.sp_mel_warp (src/libaudio.py:643-661) of the
 * three lossless streams, i.e. SPTK `mcep -a alpha -m n-1 -l fft_len -e 1.0E-8 -j 0 -f 0.0 -q {3,2,2}`
 * (src/libaudio.py:589) followed by la.mcep_to_sp_cosmat(alpha=0) (:605-631), voicing mask and clip.
 *   alpha_mag, n_mag      warping factor and size of the log-magnitude stream          (0.77, 60 @ 48 kHz)
 *   alpha_ph, n_ph        warping factor and FULL mel size of the phase streams (get_num_full_mel_coeffs_...)
 *   phase_dim             phase coefficients kept (first phase_dim of n_ph)
 *   cos_mag               HOST float64 [n_mag][n_mag]  : cos(j * w~_o), the matrix of mcep_to_sp_cosmat(alpha=0)
 *   cos_ph                HOST float64 [n_ph][phase_dim]
 * The warping matrices themselves are built on the device.
 */
int mpb_mel_create(mpb_ctx* ctx, int fft_len, double alpha_mag, int n_mag, double alpha_ph, int n_ph,
                   int phase_dim, const double* cos_mag, const double* cos_ph, mpb_mel** out);
int mpb_mel_destroy(mpb_mel* plan);
/* test hook: W^T of stream 0 (mag) / 1 (phase) as HOST float32 [fft_len/2+1][n]                         */
int mpb_mel_get_warp_matrix(mpb_mel* plan, int which, float* out_host);

/* mag/real/imag: nfrm x (fft_len/2+1) lossless features (feat_dtype); voi[f] != 0 for voiced frames.
 * out_mag_mel: nfrm x n_mag (log); out_real_mel / out_imag_mel: nfrm x phase_dim, masked and clipped.      */
int mpb_mel_compress_dev(mpb_mel* plan, void* stream,
                         const void* mag, const void* real, const void* imag, int feat_dtype,
                         const uint8_t* voi, int64_t nfrm,
                         void* out_mag_mel, void* out_real_mel, void* out_imag_mel, int out_dtype);
int mpb_mel_compress_host(mpb_mel* plan,
                          const double* mag, const double* real, const double* imag,
                          const uint8_t* voi, int64_t nfrm,
                          double* out_mag_mel, double* out_real_mel, double* out_imag_mel);
/* analysis_compressed (src/magphase.py:2947-2988) minus wav reading / REAPER / lf0: signal + frame geometry
 * in, low-dimensional features out; the lossless features stay in HBM (float32 scratch).                */
int mpb_analysis_compressed_host(mpb_mel* plan,
                                 const double* sig, int64_t n_sig,
                                 const int64_t* centre, const int32_t* left, const int32_t* right,
                                 const uint8_t* voi, int64_t nfrm, int compute_dtype,
                                 double* out_mag_mel, double* out_real_mel, double* out_imag_mel);

/* Device variant: sig / centre / left / right / voi and the outputs are DEVICE pointers; the intermediate float32
 * log periodograms live in a plan-owned scratch (frame-chunked), float64 butterflies.                     */
int mpb_analysis_compressed_dev(mpb_mel* plan, void* stream,
                                const void* sig, int sig_dtype, int64_t n_sig,
                                const int64_t* centre, const int32_t* left, const int32_t* right,
                                const uint8_t* voi, int64_t nfrm,
                                void* out_mag_mel, void* out_real_mel, void* out_imag_mel, int out_dtype);
/* Same with the utterances as n_sigs separate HOST arrays; centre[] indexes their virtual concatenation.  */
int mpb_analysis_compressed_hostv(mpb_mel* plan,
                                  const double* const* sigs, const int64_t* sig_lens, int32_t n_sigs,
                                  const int64_t* centre, const int32_t* left, const int32_t* right,
                                  const uint8_t* voi, int64_t nfrm, int compute_dtype,
                                  double* out_mag_mel, double* out_real_mel, double* out_imag_mel);

/* mpb_analysis_compressed_hostv with the caller's own element types: sig_dtype MPB_F64 / MPB_F32 / MPB_I16 (PCM16 as the
 * reference's wav files hold it: the device applies sf.read's 1/32768), out_dtype MPB_F64 / MPB_F32 (the reference's feature
 * files are float32, src/libutils.py:122-127).  Narrow types skip the host-side narrowing pass and halve the PCIe bytes.    */
int mpb_analysis_compressed_hostv2(mpb_mel* plan,
                                   const void* const* sigs, int sig_dtype, const int64_t* sig_lens, int32_t n_sigs,
                                   const int64_t* centre, const int32_t* left, const int32_t* right,
                                   const uint8_t* voi, int64_t nfrm,
                                   void* out_mag_mel, void* out_real_mel, void* out_imag_mel, int out_dtype);

/* la.sp_to_mcep (src/libaudio.py:575-601) itself, for three spectra at once: float32-rounded mel cepstra of a
 * (in_type 3, |X|) and of b, c (in_type 2, ln|X|; pass x*ln(10)/20 for in_type 1 "dB" input).  HOST float64 rows of
 * fft_len/2+1 bins; out_a: nfrm x n_mag, out_b / out_c: nfrm x n_ph.  Used by the legacy
 * analysis_with_del_comp_and_ph_encoding (src/magphase.py:573-598).                                        */
int mpb_sp_to_mcep_host(mpb_mel* plan, const double* a, const double* b, const double* c, int64_t nfrm,
                        double* out_a, double* out_b, double* out_c);

/* analysis_compressed(b_const_rate=True) (src/magphase.py:2966-2983): nfrm variable-rate frames are analysed on the
 * device; constant-rate output frame f is (1-w[f])*frame[r0[f]] + w[f]*frame[r1[f]] of the LOSSLESS features
 * (interp_from_variable_to_const_frm_rate, :2219-2239, interpolated inside the tile-product loader before the log) and
 * then compressed.  lerp_r0 / lerp_r1 / lerp_w / voi_out: HOST arrays with n_out entries.                     */
int mpb_analysis_compressed_const_hostv(mpb_mel* plan,
                                        const double* const* sigs, const int64_t* sig_lens, int32_t n_sigs,
                                        const int64_t* centre, const int32_t* left, const int32_t* right, int64_t nfrm,
                                        const int32_t* lerp_r0, const int32_t* lerp_r1, const float* lerp_w,
                                        const uint8_t* voi_out, int64_t n_out,
                                        double* out_mag_mel, double* out_real_mel, double* out_imag_mel);

/* ---- compressed synthesis ------------------------------------------------------------------- */
typedef struct mpb_syn mpb_syn;
/*
 * Plan for synthesis_from_compressed (src/magphase.py:825-997).  The host mirror builds, in float64, the fixed
 * linear maps of la.sp_mel_unwarp (src/libaudio.py:667-684) and the per-bin tables:
 *   u_mag  HOST [n_mag][fft_len/2+1]   log|X| = mag_mel_log . u_mag
 *   u_ph   HOST [n_ph][hb]             real/imag = {real,imag}_mel . u_ph  (the nearest-extrapolation padding of
 *                                      phase_uncompress_type1_mcep, src/magphase.py:1219-1235, folded in);
 *                                      hb = bins below the upper edge of the crossfade band (<= fft_len/4)
 *   tab    HOST [3][fft_len/2+1]       P  = sqrt(mask) * voiced tilt       (src/magphase.py:873-875, 940-946)
 *                                      Av = sqrt(1 - mask)                 (:947)
 *                                      Au = unvoiced aperiodic tilt        (:917-918)
 */
int mpb_syn_create(mpb_ctx* ctx, int fft_len, int n_mag, int n_ph, int hb,
                   const double* u_mag, const double* u_ph, const double* tab, mpb_syn** out);
int mpb_syn_destroy(mpb_syn* plan);

/* Per-frame / per-utterance bookkeeping of a batch (all arrays host pointers for *_host, device for *_dev). */
typedef struct mpb_syn_frames {
    int64_t nfrm;
    const int32_t* pm;        /* cumsum(trunc(shift)) inside the utterance                 (:879-880)           */
    const int64_t* ncentre;   /* noise frame geometry, as in mpb_analysis_lossless_*        (:886-896)           */
    const int32_t* nleft;
    const int32_t* nright;
    const uint8_t* voi;       /* voiced flag                                                                    */
    const uint8_t* nkind;     /* MPB_WIN_* of the noise frame (Bartlett^2.5 for voiced when b_voi_ap_win)      */
    const int32_t* win_a;     /* anti-ringing window half lengths                           (:968-973)           */
    const int32_t* win_b;
    const int32_t* row0;      /* feature row of the frame; with constant-rate input the frame is               */
    const int32_t* row1;      /*   (1-roww)*row0 + roww*row1 of the UN-WARPED features (:861-870); NULL = none  */
    const float* roww;
    int32_t n_utt;
    const int64_t* utt_frm_off;   /* [n_utt+1] */
    const int64_t* utt_out_off;   /* [n_utt+1] */
    const int32_t* utt_t0;        /* [n_utt]   see mpb_synthesis_lossless_dev */
} mpb_syn_frames;

/* mag_mel: n_rows x n_mag; real_mel/imag_mel: n_rows x n_ph (in_dtype); need_ph[row] != 0 where the phase rows
 * are needed; noise: uniform(-1,1) samples (float32) of all utterances; runs from mpb_plan_ola_runs.
 * per_linear: 0 = per_phase_type 'magphase', 1 = 'linear', 2 = 'min_phase' (then pass need_ph all zero).
 * Launches un-warp, [min-phase], noise statistics, gains, synthesis.                                        */
int mpb_synthesis_compressed_dev(mpb_syn* plan, void* stream,
                                 const void* mag_mel, const void* real_mel, const void* imag_mel, int in_dtype,
                                 int64_t n_rows, const uint8_t* need_ph, const float* noise, int64_t n_noise,
                                 const mpb_syn_frames* frames, const int32_t* runs, int32_t n_runs, int per_linear,
                                 void* out, int out_dtype, int64_t n_out);
/* Optional pre-stage: the noise half of the synthesis (windowed noise frames -> FFT -> gain statistics + stored spectra,
 * src/magphase.py:886-903) enqueued on its own stream; it needs the noise and the frame geometry, not the features.  The
 * next mpb_synthesis_compressed_dev call of this plan with the same noise pointer and frame count skips that stage; the
 * caller orders the two streams (event).                                                                     */
int mpb_synthesis_noise_stage_dev(mpb_syn* plan, void* stream, const float* noise, int64_t n_noise,
                                  const mpb_syn_frames* frames);
int mpb_synthesis_compressed_host(mpb_syn* plan,
                                  const double* mag_mel, const double* real_mel, const double* imag_mel,
                                  int64_t n_rows, const uint8_t* need_ph, const double* noise, int64_t n_noise,
                                  uint32_t* mt_key, int32_t* mt_pos,   /* noise == NULL: draw it on the device from
                                                                          NumPy's legacy state (in/out), see below */
                                  const mpb_syn_frames* frames, int per_linear,
                                  const double* hpf_sos,   /* output high-pass as 2 biquads (mpb_sos2_*), or NULL */
                                  double* out, int64_t n_out);

/* mpb_synthesis_compressed_host with float32 or float64 features (in_dtype) and waveform (out_dtype).                     */
int mpb_synthesis_compressed_host2(mpb_syn* plan,
                                   const void* mag_mel, const void* real_mel, const void* imag_mel, int in_dtype,
                                   int64_t n_rows, const uint8_t* need_ph, const double* noise, int64_t n_noise,
                                   uint32_t* mt_key, int32_t* mt_pos,
                                   const mpb_syn_frames* frames, int per_linear, const double* hpf_sos,
                                   void* out, int out_dtype, int64_t n_out);

/* mpb_synthesis_compressed_host2 with the feature rows of every utterance in a host block of its own (one feature file per
 * utterance, src/magphase.py:3229-3275): block b holds block_rows[b] rows; with variable-rate features block b is utterance b
 * of `frames`.  The blocks are copied straight into page-locked staging by the host thread pool.                     */
int mpb_synthesis_compressed_hostv2(mpb_syn* plan,
                                    const void* const* mag_blocks, const void* const* real_blocks, const void* const* imag_blocks,
                                    const int64_t* block_rows, int32_t n_blocks, int in_dtype,
                                    const uint8_t* need_ph, const double* noise, int64_t n_noise,
                                    uint32_t* mt_key, int32_t* mt_pos,
                                    const mpb_syn_frames* frames, int per_linear, const double* hpf_sos,
                                    void* out, int out_dtype, int64_t n_out);

/*
 * compute_lossless_feats (src/magphase.py:457-476) on ready-made half spectra: fft holds n complex128 values (the output
 * of mpb_frames_fft_host, any shape), mag = |X|, real = Re X/|X|, imag = Im X/|X|, all 0 where |X| == 0; float64.
 * (mpb_analysis_lossless_* fuse this into the analysis kernel.)
 */
int mpb_lossless_feats_host(mpb_ctx* ctx, const double* fft, int64_t n, double* mag, double* real, double* imag);

/*
 * ola (src/magphase.py:34-62) as a function of its own: overlap-add of nfrm ready-made time-domain frames
 * frames[nfrm][frmlen] (float64, frame centre = column frmlen/2) at the integer pitch marks pm (non-decreasing).
 * out[j] = sum over i, in frame order, of frames[i][j + t0 - pm[i] + frmlen/2] where that column exists; t0 and n_out are
 * the cut of :58-60 as computed by the host mirror (magphase.ola_geometry, Python slice semantics included).  The sums
 * are bit-identical to the reference's loop.  (The synthesis entry points above overlap-add inside their inverse-FFT
 * kernels; this is the standalone operator.)
 */
int mpb_ola_dev(mpb_ctx* ctx, void* stream, const double* frames, const int32_t* pm, int64_t nfrm, int frmlen,
                int32_t t0, double* out, int64_t n_out);
int mpb_ola_host(mpb_ctx* ctx, const double* frames, const int32_t* pm, int64_t nfrm, int frmlen,
                 int32_t t0, double* out, int64_t n_out);

/* ---- post-filter and minimum phase ---------------------------------------------------------- */
/*
 * post_filter (src/magphase.py:2300-2378): ave[b] = mean(x[centre[b]-half[b] .. centre[b]+half[b]]),
 * out[b] = (x[b]-ave[b])*tilt[b] + ave[b], first and last bin passed through.  centre/half/tilt (dim entries) are
 * computed by the host mirror from av_len_at_zero / av_len_at_nyq / boost_* exactly as the reference does.
 */
int mpb_post_filter_dev(mpb_ctx* ctx, void* stream, const void* x, int dtype, int64_t nfrm, int dim,
                        const int32_t* centre, const int32_t* half, const double* tilt, void* out);
int mpb_post_filter_host(mpb_ctx* ctx, const double* x, int64_t nfrm, int dim,
                         const int32_t* centre, const int32_t* half, const double* tilt, double* out);
/*
 * Heavy step of post_filter_merlin (src/magphase.py:3375-3465): r[0] of SPTK `freqt -m n-1 -a alpha -M L/2-1 -A 0 | c2acr -M 0
 * -l L` (:3419-3427) for nfrm cepstra c[nfrm][n].  G is a HOST float64 table [n][L/2+1], the all-pass transform folded into
 * the cosine table of the length-L transform by the host mirror; r0[f] = (1/L) sum_k w_k exp(2 (c[f] . G)[k]).  The SPTK
 * binaries are not available: the pipeline is restated from their published algorithms (parity unpinned, like mcep).
 */
int mpb_cep_energy_host(mpb_ctx* ctx, const double* c, int64_t nfrm, int n, const double* G, int K, int L, double* r0);
/*
 * la.build_min_phase_from_mag_spec (src/libaudio.py:920-934): log|X| -> real cepstrum -> causal lifter -> FFT ->
 * exp, one fused float64 kernel.  mag: nfrm x (fft_len/2+1); out_cplx: nfrm x (fft_len/2+1) x 2 (re, im).
 */
int mpb_min_phase_dev(mpb_ctx* ctx, void* stream, const void* mag, int dtype, int64_t nfrm, int fft_len,
                      void* out_cplx);
int mpb_min_phase_host(mpb_ctx* ctx, const double* mag, int64_t nfrm, int fft_len, double* out_cplx);

/* ---- output high-pass ----------------------------------------------------------------------- */
/*
 * The reference's scipy.signal.lfilter(butter(4, 40 Hz, 'highpass')) on the synthesised waveform
 * (src/magphase.py:981-995), in place, per utterance, as a blocked state-space scan: chunks run the filter in parallel
 * from a zero state, a per-utterance carry pass links them, chunks rerun from their true start state.
 * sos: HOST, two biquads in scipy's sos layout (2 x [b0 b1 b2 1 a1 a2]) factored by the host mirror from the
 * reference's (b, a) -- the 4th-order direct form's state is too ill-conditioned for the carry pass.
 * utt_off: HOST [n_utt+1] sample offsets of the concatenated utterances.  _dev: x is a DEVICE buffer (dtype).
 */
int mpb_sos2_dev(mpb_ctx* ctx, void* stream, void* x, int dtype, const int64_t* utt_off, int32_t n_utt,
                 const double* sos);
int mpb_sos2_host(mpb_ctx* ctx, double* x, const int64_t* utt_off, int32_t n_utt, const double* sos);

/* ---- batch bookkeeping on the host (integer arithmetic, no CUDA) ---------------------------- */
/*
 * The reference's per-utterance NumPy bookkeeping for a whole batch in one pass; exp()/log() stay with the caller.
 * mpb_analysis_geometry: windowing() frame geometry (src/magphase.py:74-84, 112-117) + the argument of the log in
 * format_for_modelling (voi * medfilt3(voi_in * fs / shift), :2198-2207, :2499-2501).
 * mpb_syn_geometry: the per-frame / per-utterance arrays of mpb_syn_frames for variable-rate synthesis_from_compressed
 * (:879-896 noise frames, :968-971 anti-ringing windows, ola() :34-62); ns_len[u] = noise samples of utterance u.
 */
int mpb_analysis_geometry(const int64_t* pm_rounded, const int64_t* utt_frm_off, const int64_t* n_smpls, int32_t n_utt,
                          const double* voi_in, double fs, int64_t* centre, int32_t* left, int32_t* right,
                          double* f0_med, uint8_t* voi8);
/* mpb_const_rate_scan: the reverse scan of get_shifts_and_frm_locs_from_const_shifts (src/magphase.py:1426-1449) for a batch,
 * np.interp semantics bit for bit.  shift_c: constant-rate shifts of all utterances, row_off[n_utt+1], step = fs * 5 / 1000.
 * Utterance u writes its (shift, location) pairs in SCAN order (last frame first) from index 2 * row_off[u]; count[u]
 * entries (at most 2 n_u - 1).                                                                               */
int mpb_const_rate_scan(const double* shift_c, const int64_t* row_off, int32_t n_utt, double step,
                        double* out_shift, double* out_loc, int64_t* count);
int mpb_syn_geometry(const int64_t* shift_trunc, const uint8_t* voi, const int64_t* utt_frm_off, int32_t n_utt,
                     int fft_len, int b_voi_ap_win, int32_t* pm, int64_t* ncentre, int32_t* nleft, int32_t* nright,
                     uint8_t* nkind, int32_t* win_a, int32_t* win_b, int32_t* row0, int64_t* utt_out_off,
                     int32_t* utt_t0, int64_t* ns_len);

/* ---- NumPy legacy random stream ------------------------------------------------------------- */
/*
 * n draws of np.random.uniform(low, high) from NumPy's global legacy MT19937 stream, generated on the device
 * bit for bit (the reference draws its aperiodic noise there, src/magphase.py:883).  key[624] / *pos are the
 * HOST copies of np.random.get_state()[1] / [2]; both are advanced exactly as NumPy would advance them, so
 * np.random.set_state() afterwards leaves the stream where the reference would have left it.
 */
int mpb_mt19937_uniform_dev(mpb_ctx* ctx, void* stream, uint32_t* key, int32_t* pos, int64_t n,
                            double low, double high, void* out_dev, int out_dtype);
int mpb_mt19937_uniform_host(mpb_ctx* ctx, uint32_t* key, int32_t* pos, int64_t n,
                             double low, double high, double* out);
/* Enqueue-only variant of mpb_mt19937_uniform_dev: fills out_dev on `stream` without synchronising and without advancing the
 * caller's state (device-resident pipelines that re-draw the same stretch of the stream every step).                         */
int mpb_mt19937_fill_dev(mpb_ctx* ctx, void* stream, const uint32_t* key, int32_t pos, int64_t n,
                         double low, double high, void* out_dev, int out_dtype);
/*
 * Long draws are cut into segments generated by different CTAs; a CTA reaches its segment by jump-ahead:
 * x[m+J] = XOR over the set bits i of (x^J mod phi) of x[m+i], phi = characteristic polynomial of the twister
 * (derived at run time by Berlekamp-Massey).  This HOST-only call returns x^n_words mod phi (n_words = 0: phi
 * without its leading term x^19937) as 624 little-endian 32-bit words, so that the tests can check the algebra.
 */
int mpb_mt19937_jump_poly(int64_t n_words, uint32_t* out624);

#ifdef __cplusplus
}
#endif
#endif /* MAGPHASE_B200_H */

// Internal launch interfaces between the C-ABI layer (mpb_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/magphase_b200.h"

namespace mpb {

enum { MODE_FEATS = 0, MODE_FFT = 1, MODE_LOGSQ = 2, MODE_LOGP = 3 };

struct AnalysisArgs {
    const void* sig; int sig_dtype; int64_t n_sig;
    const int64_t* centre; const int32_t* left; const int32_t* right; const uint8_t* win;
    int64_t nfrm; int fft_len; int compute_dtype;
    const void* tw;                 // twiddle table in the compute precision
    void* out_a; void* out_b; void* out_c; int out_dtype;
    int mode;                       // MODE_FEATS: a=mag b=real c=imag;  MODE_FFT: a=interleaved complex;
                                    // MODE_LOGP: a,b,c = log periodograms of mag/real/imag as SPTK mcep sees them (float32)
    int num_sms;
    const uint8_t* ph_mask = nullptr;   // MODE_LOGP: per-frame flag, 0 = the phase rows (b, c) of the frame are not needed
    int row_pitch = 0;                  // MODE_LOGP: output row pitch in elements (0: fft_len/2+1, unpadded)
    const int32_t* ph_row = nullptr;    // MODE_LOGP: row of frame f in the COMPACTED phase matrices b, c (< 0: none); NULL: row f
};
cudaError_t launch_analysis(const AnalysisArgs& a, cudaStream_t st);
cudaError_t launch_noise_stats(const AnalysisArgs& a, cudaStream_t st);   // out_a: double[nfrm]; out_b: float2[nfrm][fft_len/2+2] spectra or NULL
cudaError_t launch_analysis_logp(const AnalysisArgs& a, cudaStream_t st);  // float64 compute, float32 log periodograms

// One OLA run = consecutive frames of one utterance handled by one CTA (see mpb_synthesis.cu).
struct OlaRun {
    int32_t first;      // global index of the first frame
    int32_t count;      // frames in the run
    int32_t utt;        // utterance index
    int32_t flags;      // bit0: a previous run of the same utterance exists, bit1: a next run exists
};

struct SynthArgs {
    const void* mag; const void* real; const void* imag; int feat_dtype;
    const int32_t* pm; int64_t nfrm_total;
    const int64_t* utt_frm_off; const int64_t* utt_out_off; const int32_t* utt_t0; int32_t n_utt;
    const OlaRun* runs; int32_t n_runs;
    int* run_ticket;                // device int, zeroed by the launcher: dynamic run hand-out
    int fft_len; int compute_dtype;
    const void* tw;
    void* out; int out_dtype; int64_t n_out;
    int num_sms;
};
cudaError_t launch_synthesis_lossless(const SynthArgs& a, cudaStream_t st);

// ---- mel compression (mpb_mel.cu) ----
constexpr int MEL_KSLICE = 256;        // spectral bins per K-slice of the tile product (float32 accumulation inside a slice, float64 across; 128 / 256 / 512 measured: 2.03 / 1.91 / 1.88 ms gemm + finish)
constexpr int MEL_MAX_COEFFS = 256;    // largest supported mag_dim / nmel (nmel = 212 for phase_dim 45 at alpha_phase 0)

struct MelArgs {
    const void* mag; const void* real; const void* imag; int feat_dtype;   // nfrm x (fft_len/2+1)
    int pre_logp;                                                           // rows already hold log periodograms
    int raw_mc;                                                             // output the float32-rounded mel cepstra themselves
                                                                            // (la.sp_to_mcep): no cosine matrix, mask or clip
    const uint8_t* voi; int64_t nfrm; int fft_len;
    const float* wt_mag; int ld_mag; const float* wt_ph; int ld_ph;         // W^T, [kpad][ld] float32
    const double* cos_mag; int n_mag; const double* cos_ph; int n_ph; int phase_dim;
    float* partial; int ncp_max;                                            // [3][nfrm][slices][ncp_max]
    void* out_mag; void* out_real; void* out_imag; int out_dtype;
    const int32_t* vidx; const int32_t* cidx; const int32_t* vcount;        // voiced-frame compaction (NULL: off)
    const int32_t* lerp_r0; const int32_t* lerp_r1; const float* lerp_w;    // output frame f = lerp of two source rows (NULL: off)
    const float* wt_tc_mag = nullptr; const float* wt_tc_ph = nullptr;      // W^T pre-split per 32-bin stage for the tensor-core product (mpb_mel_warp_tc.cu; NULL: off)
    int lp_pitch = 0;                                                       // > 0: mag / real / imag are float32 log periodograms with this row pitch (multiple
                                                                            // of 4 floats), the real / imag rows compacted to the voiced frames (row = cidx[f])
    int num_sms = 0;
};
cudaError_t launch_voiced_compact(const uint8_t* voi, int n, int32_t* vidx, int32_t* cidx, int32_t* count, cudaStream_t st);
cudaError_t build_warp_matrix(int fft_len, int n_out, double alpha, float* wt32, double* scratch64, int ld,
                              cudaStream_t st);
cudaError_t launch_mel_gemm(const MelArgs& a, cudaStream_t st);
cudaError_t launch_mel_finish(const MelArgs& a, cudaStream_t st);
// tcgen05 tile product with the finish step fused into its epilogue (mpb_mel_warp_tc.cu): the default of the fused compressed
// analysis whenever both streams have at most 64 coefficients (MPB_MEL_TC=0 at plan creation selects the FMA kernels)
size_t mel_tc_operand_bytes(int fft_len);
bool mel_tc_usable(const MelArgs& a);
cudaError_t build_warp_matrix_tc(int fft_len, const float* wt32, int ld, float* out, cudaStream_t st);
cudaError_t launch_mel_warp_tc(const MelArgs& a, cudaStream_t st);   // log periodograms -> float32 mel cepstra (a.partial)
cudaError_t launch_mel_cos(const MelArgs& a, cudaStream_t st);       // mel cepstra -> features (cosine matrix, floor / clip)

// ---- mel un-warping + compressed synthesis (mpb_unwarp.cu, mpb_synth_comp.cu) ----
struct UnwarpArgs {
    const void* mag_mel; const void* real_mel; const void* imag_mel; int in_dtype;   // [nfrm][n_mag], [nfrm][n_ph]
    const uint8_t* need_ph; int64_t nfrm; int n_mag; int n_ph;
    const float* u_mag; int H; const float* u_ph; int HB;                           // [n_mag][HP], [n_ph][HBP] (zero padded)
    float* out_mag; float* out_real; float* out_imag;                               // [nfrm][HP], [nfrm][HBP] x2
    int HP; int HBP;                                                                // row pitches (multiples of 4 floats)
    uint8_t* flags;                                                                 // scratch: ceil(nfrm / 64) bytes
    float* cvt; size_t cvt_pitch;                                                   // in_dtype F64: 3 x cvt_pitch floats of scratch;
    size_t cvt_off_mag, cvt_off_ph;                                                 //   float offsets (multiples of 4) inside each matrix
    int num_sms;
    // tensor-core path (mpb_mel_unwarp_tc.cu; all NULL: FMA kernel)
    const float* ut_mag = nullptr; const float* ut_ph = nullptr;                    // U^T split [bin][hi(64) | lo(64)]
    float* xs[3] = {nullptr, nullptr, nullptr};                                     // scratch: split feature rows [row][128] per stream
    const int32_t* vidx = nullptr; const int32_t* cidx = nullptr; const int32_t* vcount = nullptr;   // compaction of need_ph
};
cudaError_t launch_mel_unwarp(const UnwarpArgs& a, cudaStream_t st);
// tcgen05 variant: the default when both streams have at most 64 coefficients (MPB_MEL_TC=0 at plan creation: FMA kernel)
size_t unwarp_tc_operand_bytes(int nbins);
size_t unwarp_tc_feature_bytes(int64_t n_rows);
cudaError_t build_unwarp_matrix_tc(const float* U, int K, int np, int nbins, float* out, cudaStream_t st);
bool unwarp_tc_usable(const UnwarpArgs& a);
cudaError_t launch_mel_unwarp_tc(const UnwarpArgs& a, cudaStream_t st);
cudaError_t launch_lerp_rows(const float* rows, int pitch, const int32_t* row0, const int32_t* row1, const float* roww,
                             int64_t nfrm, float* out, cudaStream_t st);

struct SynthCompArgs {
    const float* m_mag; const float* m_real; const float* m_imag; int H; int HB; int HP; int HBP;
    const float* noise; int64_t n_noise;
    const float2* nspec;                                                            // [nfrm][fft_len/2 + 2] noise spectra (k_analysis<noise_logsq>)
    const int32_t* pm; const int64_t* ncentre; const int32_t* nleft; const int32_t* nright;
    const uint8_t* voi; const uint8_t* nkind; const int32_t* win_a; const int32_t* win_b;
    const int32_t* row0; const int32_t* row1; const float* roww;                    // row0 NULL: frame f reads row f; row1/roww NULL: no interpolation
    const double* logsq; const int64_t* utt_frm_off;                                // noise statistics per frame
    double* inv_gain;                                                               // [n_utt][2] scratch
    const float* tab;                                                               // [3][H]: P, Av, Au
    const int64_t* utt_out_off; const int32_t* utt_t0; int32_t n_utt;
    int32_t utt_a, utt_b;                                                           // utterances whose gains this call computes
    const OlaRun* runs; int32_t n_runs; int64_t nfrm;
    int* run_ticket;                                                                // device int, zeroed by the launcher: dynamic run hand-out
    int fft_len; int per_linear;
    const void* tw;                                                                 // float32 twiddles
    void* out; int out_dtype; int64_t n_out;
    int num_sms;
};
cudaError_t launch_noise_gain(const SynthCompArgs& a, cudaStream_t st);
cudaError_t launch_synthesis_compressed(const SynthCompArgs& a, cudaStream_t st);

// ---- post-filter and minimum phase (mpb_extra.cu) ----
cudaError_t launch_post_filter(const void* x, int dtype, int64_t nfrm, int dim, const int32_t* centre, const int32_t* half,
                               const double* tilt, void* out, cudaStream_t st);
cudaError_t launch_lossless_feats(const void* x, int64_t n, double* mag, double* re, double* im, int num_sms, cudaStream_t st);
cudaError_t launch_ola_gather(const double* frames, const int32_t* pm, int64_t nfrm, int frmlen, int32_t t0, double* out,
                              int64_t n_out, cudaStream_t st);
cudaError_t launch_cep_energy(const double* c, int64_t nfrm, int n, const double* G, int K, int L, double* r0, int num_sms,
                              cudaStream_t st);
cudaError_t launch_min_phase(int fft_len, const void* mag, int dtype, int64_t nfrm, const void* tw64, void* out_cplx,
                             int num_sms, cudaStream_t st);
cudaError_t launch_min_phase_split(int fft_len, const float* mag, int in_pitch, int64_t nfrm, const void* tw64, float* out_re,
                                   float* out_im, int nb, int out_pitch, int num_sms, cudaStream_t st);

// two cascaded biquads (scipy sos layout [2][6]) as a blocked state-space scan, in place on x (utterances
// concatenated, utt_off[n_utt+1]; chunk_off[n_utt+1]: first chunk of every utterance; state: 4 doubles per chunk)
cudaError_t launch_sos2(void* x, int dtype, const int64_t* utt_off, const int64_t* chunk_off, int n_utt, int64_t n_chunks,
                        int L, const double* sos, const double* ML, double* state, cudaStream_t st);

// ---- NumPy legacy MT19937 stream on the device (mpb_rng.cu) ----
// n draws of uniform(low, high).  state_dev: 2 x 625 words of device scratch (624-word key + position word, ping-pong)
// with the handed key in half slot_in; *final_slot tells which half holds the final state.  jump_idx_dev: device copy
// of mt19937_jump_table_host(), required when mt19937_needs_jump(pos, n) (the draw spans more than one segment).
cudaError_t launch_mt19937_uniform(uint32_t* state_dev, int slot_in, int32_t pos_host, int* final_slot,
                                   const uint16_t* jump_idx_dev, uint32_t* raw_dev, int64_t n, double low, double high,
                                   void* out, int out_dtype, cudaStream_t st);
bool mt19937_needs_jump(int32_t pos, int64_t n);
const uint16_t* mt19937_jump_table_host(size_t* n_entries);
int mt19937_jump_poly(int64_t n_words, uint32_t* out624);

}  // namespace mpb

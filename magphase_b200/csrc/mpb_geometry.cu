// Host-side integer bookkeeping of a batch in one pass (no CUDA work, no floating-point functions).
//
// The reference does this with a dozen NumPy calls per utterance (windowing() src/magphase.py:74-84, 112-117;
// shift_to_f0 :2198-2207; synthesis_from_compressed :879-896, 968-971; ola() :34-62).  Vectorising those calls over a
// batch still costs ~40 temporaries of the batch's size; at 100k+ frames that is most of the host time of a call.  The
// arithmetic is integer (plus IEEE multiply / divide and a median of three, which are exactly reproducible), so it
// moves here; exp() / log() stay in NumPy, whose results define parity.
#include <math.h>

#include "mpb_ctx.h"

using namespace mpb;

extern "C" {

// Analysis side.  pm[F]: rounded pitch marks (sample index inside the utterance), utt_frm_off[U+1], n_smpls[U]
// (samples per utterance), voi_in[F] (0/1 as float64), fs.
//   centre[F]  absolute index of the mark in the concatenated signal            (P[1:-1] + signal offset)
//   left[F], right[F]   frame half lengths  P[f+1]-P[f], P[f+2]-P[f+1]  with P = [0, pm..., n_smpls-1]
//   f0_med[F]  voi * medfilt3(f0), f0 = voi_in * fs / left, voi = f0 > 0 (zero padded median per utterance):
//              the argument of the log in format_for_modelling (src/magphase.py:2499-2501)
//   voi8[F]    voi as bytes
int mpb_analysis_geometry(const int64_t* pm, const int64_t* utt_frm_off, const int64_t* n_smpls, int32_t n_utt,
                          const double* voi_in, double fs, int64_t* centre, int32_t* left, int32_t* right,
                          double* f0_med, uint8_t* voi8) {
    if (!pm || !utt_frm_off || !n_smpls || !voi_in || !centre || !left || !right || !f0_med || !voi8 || n_utt < 0)
        return fail(MPB_ERR_BAD_ARG, "NULL argument");
    int64_t sig_off = 0;
    for (int32_t u = 0; u < n_utt; ++u) {
        const int64_t a = utt_frm_off[u], b = utt_frm_off[u + 1];
        if (b < a) return fail(MPB_ERR_BAD_ARG, "utt_frm_off not non-decreasing");
        for (int64_t f = a; f < b; ++f) {
            const int64_t prev = f > a ? pm[f - 1] : 0;
            const int64_t next = f + 1 < b ? pm[f + 1] : n_smpls[u] - 1;
            const int64_t l = pm[f] - prev, r = next - pm[f];
            if (l < INT32_MIN || l > INT32_MAX || r < INT32_MIN || r > INT32_MAX)
                return fail(MPB_ERR_FRAME_GEOM, "frame length does not fit 32 bits");
            centre[f] = pm[f] + sig_off;
            left[f] = (int32_t)l;
            right[f] = (int32_t)r;
        }
        // f0 = voi * fs / shift (left to right, like NumPy), then the zero-padded median of three
        double fm1 = 0.0;                                     // f0[f-1]
        double f0c = a < b ? voi_in[a] * fs / (double)left[a] : 0.0;
        for (int64_t f = a; f < b; ++f) {
            const double f0n = f + 1 < b ? voi_in[f + 1] * fs / (double)left[f + 1] : 0.0;
            const double lo = fm1 < f0c ? fm1 : f0c, hi = fm1 < f0c ? f0c : fm1;
            const double med = lo > (hi < f0n ? hi : f0n) ? lo : (hi < f0n ? hi : f0n);
            const bool v = f0c > 0.0;
            voi8[f] = v ? 1 : 0;
            f0_med[f] = (v ? 1.0 : 0.0) * med;
            fm1 = f0c;
            f0c = f0n;
        }
        sig_off += n_smpls[u];
    }
    return MPB_OK;
}

// Synthesis side (variable frame rate).  shift[F]: per-frame shifts already truncated to integers (:879), voi[F],
// utt_frm_off[U+1] (every utterance has >= 2 frames).  Fills the per-frame arrays of mpb_syn_frames and the
// per-utterance output geometry; ns_len[U] = noise samples per utterance.  MPB_ERR_FRAME_GEOM when a frame is longer
// than fft_len/2 (the reference's frame_shift() fails on a negative pad, src/libaudio.py:137-140).
int mpb_syn_geometry(const int64_t* shift, const uint8_t* voi, const int64_t* utt_frm_off, int32_t n_utt, int fft_len,
                     int b_voi_ap_win, int32_t* pm, int64_t* ncentre, int32_t* nleft, int32_t* nright, uint8_t* nkind,
                     int32_t* win_a, int32_t* win_b, int32_t* row0, int64_t* utt_out_off, int32_t* utt_t0, int64_t* ns_len) {
    if (!shift || !voi || !utt_frm_off || !pm || !ncentre || !nleft || !nright || !nkind || !win_a || !win_b || !row0 ||
        !utt_out_off || !utt_t0 || !ns_len || n_utt < 0)
        return fail(MPB_ERR_BAD_ARG, "NULL argument");
    const int64_t half = fft_len / 2;
    int64_t noise_off = 0;
    utt_out_off[0] = 0;
    for (int32_t u = 0; u < n_utt; ++u) {
        const int64_t a = utt_frm_off[u], b = utt_frm_off[u + 1], n = b - a;
        if (n < 2) return fail(MPB_ERR_BAD_ARG, "every utterance needs at least two frames");
        int64_t p = 0;
        for (int64_t f = a; f < b; ++f) {                     // pm = cumsum(shift)
            p += shift[f];
            if (p > INT32_MAX || p < INT32_MIN) return fail(MPB_ERR_FRAME_GEOM, "pitch mark does not fit 32 bits");
            pm[f] = (int32_t)p;
        }
        const int64_t pm_first = pm[a], pm_last = pm[b - 1];
        const int64_t ns = pm_last + (pm_last - pm[b - 2]);   // :882
        ns_len[u] = ns;
        for (int64_t f = a; f < b; ++f) {
            const int64_t prev = f > a ? pm[f - 1] : 0, next = f + 1 < b ? pm[f + 1] : ns - 1;
            const int64_t l = pm[f] - prev, r = next - pm[f];
            if (l > half || r >= half) return fail(MPB_ERR_FRAME_GEOM, "negative dimensions are not allowed");
            ncentre[f] = pm[f] + noise_off;
            nleft[f] = (int32_t)l;
            nright[f] = (int32_t)r;
            nkind[f] = (voi[f] && b_voi_ap_win) ? MPB_WIN_BARTLETT25 : MPB_WIN_HANN;
            // anti-ringing half lengths: se = [s0, s..., s_last, s_last]; a = se[i] + se[i+1], b = se[i+2] + se[i+3]
            const int64_t s_prev = f > a ? shift[f - 1] : shift[a];
            const int64_t s_n1 = f + 1 < b ? shift[f + 1] : shift[b - 1];
            const int64_t s_n2 = f + 2 < b ? shift[f + 2] : shift[b - 1];
            win_a[f] = (int32_t)(s_prev + shift[f]);
            win_b[f] = (int32_t)(s_n1 + s_n2);
            row0[f] = (int32_t)f;
        }
        // ola(): out = buffer[N/2 - pm[0] : ][: pm[-1] + shift[-1] + 1] with Python slice semantics (:34-62)
        const int64_t buf_len = pm_last + fft_len;
        const int64_t s0 = half - pm_first;
        int64_t start = s0 < 0 ? (buf_len + s0 > 0 ? buf_len + s0 : 0) : (s0 < buf_len ? s0 : buf_len);
        int64_t n1 = buf_len - start;
        if (n1 < 0) n1 = 0;
        int64_t want = pm_last + shift[b - 1] + 1;
        if (want < 0) want = 0;
        const int64_t n_out = n1 < want ? n1 : want;
        utt_t0[u] = (int32_t)(start + pm_first - half);
        utt_out_off[u + 1] = utt_out_off[u] + n_out;
        noise_off += ns;
    }
    return MPB_OK;
}

// Constant-rate -> variable-rate reverse scan (get_shifts_and_frm_locs_from_const_shifts, src/magphase.py:1426-1449) for a
// batch: per utterance, walk back from the last constant-rate centre, subtracting the linearly interpolated shift, until
// the position leaves [centres[0], centres[-1]].  Sequential and data dependent, hence host work -- but 750 NumPy calls per
// utterance in the mirror (90 % of the host time of constant-rate synthesis).  The interpolation reproduces np.interp bit
// for bit: interval j with xp[j] <= x < xp[j+1]; exact hits and the last knot return fp[j]; otherwise
// slope * (x - xp[j]) + fp[j] with slope = (fp[j+1] - fp[j]) / (xp[j+1] - xp[j]) (no FMA contraction in host code).
//   shift_c[R]  constant-rate shifts of all utterances (float64), row_off[U+1], step = fs * frm_rate_ms / 1000
//   out_shift / out_loc [2 R]: utterance u writes at most 2 n_u - 1 entries from out_off = 2 row_off[u], IN SCAN ORDER (last
//   frame first); count[u] = entries written.  centres[k] = step * (k + 1) as NumPy computes them (one rounding).
int mpb_const_rate_scan(const double* shift_c, const int64_t* row_off, int32_t n_utt, double step, double* out_shift,
                        double* out_loc, int64_t* count) {
    if (!shift_c || !row_off || !out_shift || !out_loc || !count || n_utt < 0) return fail(MPB_ERR_BAD_ARG, "NULL argument");
    if (!(step > 0.0)) return fail(MPB_ERR_BAD_ARG, "step must be positive");
    for (int32_t u = 0; u < n_utt; ++u) {
        const int64_t a = row_off[u], n = row_off[u + 1] - a;
        if (n < 0) return fail(MPB_ERR_BAD_ARG, "row_off not non-decreasing");
        count[u] = 0;
        if (n == 0) continue;
        const double* fp = shift_c + a;
        double* os = out_shift + 2 * a;
        double* ol = out_loc + 2 * a;
        auto xp = [step](int64_t k) { return step * (double)(k + 1); };
        const double lo = xp(0), hi = xp(n - 1);
        double pos = hi;
        int64_t j = n - 1, c = 0;
        for (int64_t it = 0; it < 2 * n - 1; ++it) {
            if (pos < lo || pos > hi) break;                       // (false for NaN, like the mirror's comparison)
            // interval search from the previous one (the scan only moves left; a negative shift is followed as well)
            if (j > n - 1) j = n - 1;
            while (j + 1 <= n - 1 && xp(j + 1) <= pos) ++j;
            while (j > 0 && xp(j) > pos) --j;
            double sft;
            if (pos != pos) sft = pos;                             // np.interp hands NaN through
            else if (j == n - 1 || xp(j) == pos) sft = fp[j];
            else {
                const double slope = (fp[j + 1] - fp[j]) / (xp(j + 1) - xp(j));
                sft = slope * (pos - xp(j)) + fp[j];
                if (sft != sft) {                                  // np.interp: "if we get nan in one direction, try the other"
                    sft = slope * (pos - xp(j + 1)) + fp[j + 1];
                    if (sft != sft && fp[j] == fp[j + 1]) sft = fp[j];
                }
            }
            ol[c] = pos;
            os[c] = sft;
            ++c;
            pos = pos - sft;
        }
        count[u] = c;
    }
    return MPB_OK;
}

}  // extern "C"

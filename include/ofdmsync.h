/*
 * ofdmsync.h -- C ABI of the B200-native OFDM synchronisation engine (libofdmsync.so).
 *
 * The reference (amcolex/ofdm-sync-math) has no FFI/plugin boundary: its hot path sits behind
 * plain module-level Python functions.  This header is the boundary a maintainer
 * would bind from those functions (ctypes stubs in INTEGRATION.md); every entry point names the
 * reference function it replaces (file:line under /root/reference).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / CUDA types (a stream is passed as void*,
 *    it is a cudaStream_t; NULL = the legacy default stream).
 *  - All *device* entry points take DEVICE pointers, are asynchronous on `stream`, allocate
 *    nothing persistent and never synchronise.  The caller owns every buffer.
 *  - *_host entry points take HOST pointers (pinned memory makes the copies asynchronous),
 *    stage through an ofs_ctx workspace, and return after the results are in host memory.
 *  - Return value: OFS_OK (0); < 0 invalid argument (ofs_last_error_string() says which);
 *    > 0 a cudaError_t.  Nothing throws across the ABI.  Re-entrant; error strings are
 *    thread-local.
 *  - Complex arrays are interleaved (re, im).  "frames" is the leading batch axis added for
 *    the B200 engine; branches (antennas) are summed before the non-linear metric exactly as
 *    in the reference (sc.py:73-74, minn.py:106-107, sync_aa.py:478-479, ...).
 */
#ifndef OFDMSYNC_H
#define OFDMSYNC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFS_ABI_VERSION 2

#define OFS_OK 0
#define OFS_EINVAL (-1)      /* bad descriptor / null pointer / unsupported size */
#define OFS_EUNSUPPORTED (-2) /* valid request that this build cannot serve (e.g. span too large) */

/* input sample formats */
#define OFS_C64 0   /* float  (re, im)   8 B / sample */
#define OFS_C128 1  /* double (re, im)  16 B / sample */
#define OFS_IQ16 2  /* int16  (i, q)     4 B / sample */

/* autocorrelation metric kinds */
#define OFS_SC 0       /* sc.sc_streaming_metric                     sc.py:42-78   (R: 2nd half)   */
#define OFS_SC_BOTH 1  /* combined_sc_min.schmidl_cox_streaming_metric  :116-164   (R: both halves) */
#define OFS_MINN 2     /* minn.minn_streaming_metric(_parameterized) minn.py:59-112, :697-751       */
#define OFS_AA 3       /* sync_aa.aa_detect_streaming loop 1         sync_aa.py:458-493 (causal)    */

/* kernel family selection */
#define OFS_PATH_AUTO 0
#define OFS_PATH_STRIPE 1 /* fast: fp32 products, fp64 carries, TMA-fed persistent stripes (c64 / iq16 in, f32 out; M always,
                             P complex64 / R float32 optionally, same pitch and offset as M) */
#define OFS_PATH_TILE 2   /* precise: float64 prefix sums, any lag / branch count / dtype */
#define OFS_PATH_ARRAY 3  /* antenna arrays (kind AA, any branch count): branch sum on chip, TMA-fed, c64 / iq16 in, f32 out */

typedef struct ofs_metric_desc {
    int32_t kind;        /* OFS_SC ... OFS_AA */
    int32_t in_dtype;    /* OFS_C64 / OFS_C128 / OFS_IQ16 */
    int32_t out_f64;     /* 0: M,R float32 + P complex64; 1: M,R float64 + P complex128 */
    int32_t path;        /* OFS_PATH_* */
    int32_t symbol_len;  /* N (SC, SC_BOTH, MINN: Q = N/4, half = N/2); L = half-preamble length for AA */
    int32_t n_branches;  /* antennas summed before the metric */
    int64_t n_frames;
    int64_t n_samples;        /* samples per frame and branch */
    int64_t x_frame_stride;   /* in samples */
    int64_t x_branch_stride;  /* in samples */
    int64_t out_stride;       /* elements between frames in M / P / R / chunk_max (>= out_len) */
    int32_t store_mode;       /* stripe path only: 0 direct 16-byte vector stores (default, faster), 1 TMA bulk stores */
    int32_t reserved;
} ofs_metric_desc;

/* Library / error -------------------------------------------------------------------------- */
int ofs_version(void);
const char *ofs_last_error_string(void);
/* number of outputs per frame for a descriptor: L-N+1 (SC, SC_BOTH, MINN; 0 if L<N), L (AA) */
int64_t ofs_metric_out_len(const ofs_metric_desc *d);
/* 1 if the stripe (fast) path can serve this descriptor and these pointers (alignment rules in DESIGN.md) */
int ofs_metric_stripe_ok(const ofs_metric_desc *d, const void *x, const void *M);
/* 1 if the antenna-array kernel (OFS_PATH_ARRAY) can serve this descriptor / input pointer */
int ofs_metric_array_ok(const ofs_metric_desc *d, const void *x);
/* outputs covered by one chunk_max entry (stripe path), 256 */
int32_t ofs_chunk_len(void);

/* Timing metric M (and optionally P, R) for a batch of frames ---------------------------------
 * Replaces sc.py:42-78, combined_sc_min.py:116-164, minn.py:59-112 / :697-751, sync_aa.py:458-493.
 * M, P, R: out_len elements per frame (out_stride apart); P and R may be NULL.
 * chunk_max (may be NULL; stripe path only): float[n_frames][ceil(n_samples/256)] at stride
 * cm_stride, max of M over each aligned block of 256 causal sample times -- used by the
 * detectors to prune their second pass. */
int ofs_metric(const ofs_metric_desc *d, const void *x, void *M, void *P, void *R,
               float *chunk_max, int64_t cm_stride, void *stream);

/* Park metric -- park.py:64-114.  n = L - 2*(N/2) outputs per frame (0 if L < N+1): M, P, E
 * (ds = h + arange(n) is implicit).  in/out dtypes as in ofs_metric_desc (kind ignored).
 * complex64 / int16-IQ input with float32 outputs and h = N/2 a multiple of 128 up to 1024 runs the block-FFT kernel
 * (band-limited self-convolution by 256-point block transforms, |d M| <= 1e-4 max M against float64); everything else, and
 * everything when the environment variable OFS_PARK_DIRECT is set to a non-zero value, the direct O(h) kernel
 * (float64 accumulation for complex128 input: the drop-in path, 1e-11). */
int ofs_park_metric(const ofs_metric_desc *d, const void *x, void *M, void *P, void *E, void *stream);

/* Detectors on metric arrays (one row per frame; float32 or float64 rows) ----------------------- */
typedef struct ofs_rows {
    const void *data;   /* row-major metric array */
    int32_t f64;        /* 0 float32, 1 float64 */
    int32_t reserved;
    int64_t n_rows, n, stride;
} ofs_rows;

/* sc.find_plateau_end_from_metric -- sc.py:81-146.  lookahead < 0 == None.  out: int64[n_rows]. */
int ofs_find_plateau_end(const ofs_rows *M, int32_t cp_len, int32_t lookahead, int32_t smooth_win,
                         int64_t *plateau_end, void *stream);

/* minn.find_minn_peak -- minn.py:131-205 (incl. _trailing_average :115-128).
 * peak: int64[n_rows] (-1 empty metric, -2 non-positive peak -> the shim raises ValueError);
 * gate_span: int64[n_rows][2] = the selected gate [start, end) AFTER bounds (fallback, minn.py:195-200:
 * the single-sample gate [peak, peak+1)); Ms (optional): smoothed metric, same dtype/stride as M. */
int ofs_find_minn_peak(const ofs_rows *M, int32_t smooth_win, double gate_threshold, int32_t has_bounds,
                       int64_t bound_lo, int64_t bound_hi, int64_t *peak, int64_t *gate_span, void *Ms,
                       void *stream);

/* Same detectors, pruned with the stripe kernel's per-chunk maxima (chunk_max rows of ofs_metric, chunk c = the 256
 * causal sample times [256c, 256c+256), output index d = t - toff): chunks whose maximum cannot reach the
 * running best are not re-read from HBM.  Results are identical to the unpruned calls. */
int ofs_find_plateau_end_pruned(const ofs_rows *M, const float *chunk_max, int64_t cm_stride, int32_t toff,
                                int32_t cp_len, int32_t lookahead, int32_t smooth_win, int64_t *plateau_end, void *stream);
int ofs_find_minn_peak_pruned(const ofs_rows *M, const float *chunk_max, int64_t cm_stride, int32_t toff,
                              int32_t smooth_win, double gate_threshold, int32_t has_bounds, int64_t bound_lo,
                              int64_t bound_hi, int64_t *peak, int64_t *gate_span, void *Ms, void *stream);

/* combined_sc_min: S&C gate construction :337-351 (gate = M_sc/max >= thr, seeded with argmax) */
int ofs_sc_gate(const ofs_rows *Msc, double threshold, uint8_t *gate, int64_t gate_stride, void *stream);
/* same, reading M only in the chunks whose maximum (chunk_max of the stripe metric kernel) can reach the gate level */
int ofs_sc_gate_pruned(const ofs_rows *Msc, const float *chunk_max, int64_t cm_stride, int32_t toff, double threshold,
                       uint8_t *gate, int64_t gate_stride, void *stream);
/* combined_sc_min.find_minn_peak :212-259 + _streaming_peak_detector :183-209.
 * peak: -1 empty M (reference returns 0), -3 empty gate region (ValueError). */
int ofs_find_minn_peak_gated(const ofs_rows *M, int32_t smooth_win, const uint8_t *gate, int64_t gate_stride,
                             int32_t has_bounds, int64_t bound_lo, int64_t bound_hi, int64_t *peak, void *stream);
/* combined_sc_min detector in one launch, without a gate array: the S&C gate test M_sc[d] / max(M_sc) >= threshold
 * (combined_sc_min.py:337-351) is evaluated only around the first gate segment, located through the S&C chunk maxima of the
 * stripe kernel; peak = first maximum of the trailing-averaged Minn metric inside that segment (:183-259).  Same result as
 * ofs_sc_gate_pruned + ofs_find_minn_peak_gated.  float32 rows of equal shape, 0 < threshold <= 1.
 * peak: int64[n_rows] (-1 empty metric, -3 empty gate); gate_span (optional): int64[n_rows][2] = [first, stop). */
int ofs_combined_peak(const ofs_rows *M_minn, const ofs_rows *M_sc, const float *chunk_max_sc, int64_t cm_stride, int32_t toff,
                      double threshold, int32_t smooth_win, int32_t has_bounds, int64_t bound_lo, int64_t bound_hi,
                      int64_t *peak, int64_t *gate_span, void *stream);

/* argmax with numpy semantics (first maximum): park.py:161, zc.py:128, zc_freq.py:147 */
int ofs_argmax(const ofs_rows *M, int64_t *index, void *stream);

/* Gate / hysteresis state machines -------------------------------------------------------------
 * Event buffers hold max_events slots per row (caller's choice, >= 1; OFS_MAX_EVENTS is the default the Python layer starts
 * with).  n_events[row] is always the TRUE number of gates of the row: a value above max_events means the list was cut and
 * the call should be repeated with a larger buffer (the reference returns unbounded lists). */
#define OFS_MAX_EVENTS 64
typedef struct ofs_event {
    int64_t peak_index;
    int64_t gate_start;
    int64_t gate_end;   /* close index (n if the gate never closed) */
    int64_t aux;        /* aa: frame_start = peak-2L+1; zc_v2: detected_start; minn_rtl: detected_index */
    double value;       /* aa: M[peak]; zc_v2: corr_mag[peak]; minn_rtl: corr_positive[peak] */
    double p_re, p_im;  /* aa: P at peak */
    double cfo;         /* aa: angle(P)*fs/(2*pi*L) in Hz */
    int32_t closed;     /* 0: gate still open at end of input */
    int32_t reserved;
} ofs_event;

/* sync_aa.aa_detect_streaming loop 2 -- sync_aa.py:495-568.  M, P: rows of length n (P complex,
 * same precision as M).  events: ofs_event[n_rows][max_events]; n_events: int32[n_rows]. */
int ofs_aa_events(const ofs_rows *M, const void *P, int32_t L, double threshold, int32_t hysteresis,
                  double sample_rate, ofs_event *events, int32_t *n_events, int32_t max_events, void *stream);

/* Fused antenna-array detector = sync_aa.aa_detect_streaming, sync_aa.py:458-568, for captures of n_antennas branches
 * (complex64 or int16 IQ on the device; L in {128, 256, 512, 1024}; rows 16-byte aligned).  One pass over x: P, R and M
 * are formed with the antenna sum kept on chip (sync_aa.py:478-479) and the (n >= L && M >= threshold) flags of
 * sync_aa.py:511 leave the metric kernel as a bitmask, so the gate FSM never re-reads M.
 * M float32 / P complex64 [n_frames][out_stride] (out_stride even), R optional; mask_ws: uint32[n_frames][mask_stride],
 * mask_stride >= ceil(n/32); events / n_events as ofs_aa_events. */
int ofs_aa_detect(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_antennas, int64_t n,
                  int64_t x_frame_stride, int64_t x_branch_stride, int32_t L, double threshold, int32_t hysteresis,
                  double sample_rate, float *M, void *P_c64, float *R, int64_t out_stride, uint32_t *mask_ws,
                  int64_t mask_stride, ofs_event *events, int32_t *n_events, int32_t max_events, void *stream);

/* Reference-ORDER [A][A] metric (sync_aa.py:321-386, 458-493): the reference's running-sum recurrences
 * `sum + sample - oldest` in its exact operation order, one thread per frame, float64 outputs
 * [n_frames][n].  P is bit-equal to the reference; needed where a detection depends on the rounding
 * of the running sums (docs/detector_test_vector.csv: the 1523/1524 tie).
 * x: (n_frames, n_antennas, n), n_antennas <= 64. */
int ofs_aa_metric_reference(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_antennas, int64_t n,
                            int32_t L, void *P_c128, double *R, double *M, uint8_t *valid, void *stream);

/* zc_v2.zc_streaming_detection -- zc_v2.py:288-336: local_sum (same dtype as corr_mag), valid, above. */
int ofs_zc_streaming_detection(const ofs_rows *corr_mag, int32_t window, int32_t thresh_value,
                               int32_t frac_bits, double min_corr_mag, void *local_sum, uint8_t *valid,
                               uint8_t *above, int64_t mask_stride, void *stream);
/* zc_v2.detect_zc_peaks -- zc_v2.py:360-450.  gate_mask optional. */
int ofs_zc_events(const ofs_rows *corr_mag, const uint8_t *valid, const uint8_t *above, int64_t mask_stride,
                  int32_t reference_length, int32_t hysteresis, ofs_event *events, int32_t *n_events,
                  int32_t max_events, uint8_t *gate_mask, void *stream);
/* ofs_zc_streaming_detection + ofs_zc_events in two launches that exchange one BIT per sample: the threshold kernel
 * (zc_v2.py:288-336) writes the above-threshold flags of the valid samples as a bitmask (mask_ws: uint32[n_rows][mask_stride],
 * mask_stride >= ceil(n / 32)) and the gate FSM (zc_v2.py:360-450) walks that -- the local_sum / valid / above arrays
 * (6 bytes per sample written, 2 read back) are not produced.  Same events as the two separate calls. */
int ofs_zc_detect(const ofs_rows *corr_mag, int32_t window, int32_t thresh_value, int32_t frac_bits, double min_corr_mag,
                  int32_t reference_length, int32_t hysteresis, uint32_t *mask_ws, int64_t mask_stride, ofs_event *events,
                  int32_t *n_events, int32_t max_events, void *stream);

/* zc_v2.detect_zc_preamble -- zc_v2.py:456-516 -- for a batch of complex64 / int16-IQ captures in three launches on the float32
 * path: matched filter (8192-point overlap-save blocks when the capture is long enough) writing |corr| only (normalize != 0:
 * per-branch normalisation, then branch sum, :488-495), running-sum threshold -> bitmask (:288-336), gate FSM (:360-450).
 * mag_ws: float[n_frames][mag_stride >= n + nr - 1] (the corr_mag rows, kept for the caller); mask_ws as ofs_zc_detect. */
int ofs_zc_v2_detect(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n, const void *ref_c128,
                     int32_t nr, int32_t normalize, int32_t window, int32_t thresh_value, int32_t frac_bits,
                     double min_corr_mag, int32_t hysteresis, float *mag_ws, int64_t mag_stride, uint32_t *mask_ws,
                     int64_t mask_stride, ofs_event *events, int32_t *n_events, int32_t max_events, void *stream);

/* minn_rtl.detect_minn_rtl -- minn_rtl.py:750-825 (== ref/minn_preamble_detector.sv:337-384).
 * corr_positive rows: float64, or int64 when is_int != 0 (integer RTL mode).  An unclosed tail gate
 * is returned as an event with closed == 0 (the reference reports it as a segment, not an event). */
int ofs_minn_rtl_events(const void *corr_positive, int32_t is_int, const uint8_t *valid, const uint8_t *above,
                        int64_t n_rows, int64_t n, int64_t stride, int32_t hysteresis, int32_t timing_offset,
                        ofs_event *events, int32_t *n_events, int32_t max_events, void *stream);

/* minn_rtl metric ------------------------------------------------------------------------------
 * Float mirror -- minn_rtl.minn_rtl_streaming_metric, minn_rtl.py:583-733.
 * x: (n_frames, n_branches, n) complex128 or complex64.  All outputs float64[n_frames][n] (stride n)
 * except the two uint8 masks.  Bit-equal to the reference on integer-valued input. */
int ofs_minn_rtl_metric(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n,
                        int32_t quarter_len, int32_t smooth_shift, int32_t threshold_value, int32_t frac_bits,
                        double *corr_total, double *corr_positive, double *smooth_metric, double *energy_total,
                        double *corr_scaled, double *energy_scaled, uint8_t *metric_valid, uint8_t *above,
                        void *stream);
/* Integer RTL datapath -- ref/minn_antenna_path.sv:63-194, ref/minn_preamble_detector.sv:247-325.
 * iq: int16 (n_frames, n_branches, n, 2).  lag_extra: 0 (minn_rtl.py lag Q) or 1 (registered delay-line
 * read of the SV).  Outputs int64[n_frames][n] + masks. */
int ofs_minn_rtl_int(const int16_t *iq, int64_t n_frames, int32_t n_branches, int64_t n, int32_t quarter_len,
                     int32_t smooth_shift, int32_t threshold_value, int32_t frac_bits, int32_t lag_extra,
                     int64_t *corr_total, int64_t *corr_positive, int64_t *smooth_metric, int64_t *energy_total,
                     uint8_t *metric_valid, uint8_t *above, void *stream);

/* Zadoff-Chu ---------------------------------------------------------------------------------------
 * Matched filter by FFT overlap-save: corr = x (*) conj(ref[::-1]) and energy = |x|^2 (*) ones(nr),
 * full length n + nr - 1 -- zc.py:115-117, zc_v2.py:244-254,268.  x: (n_frames, n_branches, n).
 * mode 0 (zc.py:118-126): sum branches, then corr / (||ref|| * sqrt(max(pow,0) + 1e-12)) -> corr_out.
 * mode 1 (zc_v2.py:488-495): normalise each branch by ||ref||*sqrt(max(E,1e-12)), then sum.
 * mode 2: raw branch-summed numerator (normalize=False).
 * corr_out: complex (c64 or c128 per out_f64) [n_frames][n+nr-1]; mag_out (optional): |corr|. */
int ofs_zc_matched_filter(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n,
                          const void *ref_c128, int32_t nr, int32_t mode, int32_t out_f64, void *corr_out,
                          void *mag_out, int64_t out_stride, void *stream);
/* zc_v2.normalize_correlation -- zc_v2.py:257-271 -- applied to a correlation the CALLER supplies (any array of the full
 * length n + nr - 1, not necessarily this library's matched-filter output): out = corr / (ref_norm * sqrt(max(E, 1e-12))),
 * E = np.convolve(|x|^2, ones(nr), "full") in float64.  x: (n_frames, n), one branch; corr / out: complex64 (f64 = 0) or
 * complex128 (f64 = 1) [n_frames][stride]; out may alias corr. */
int ofs_zc_normalize(const void *corr, const void *x, int32_t in_dtype, int64_t n_frames, int64_t n, int32_t nr,
                     double ref_norm, int32_t f64, void *out, int64_t stride, void *stream);
/* zc_freq.compute_frequency_metric -- zc_freq.py:62-99, as a sliding DFT of the used bins.
 * bins: DFT bin numbers k_j in [0, n_fft); templ: complex128[nbins].  metric: [n_frames][n-(n_fft+cp)+1]. */
int ofs_zc_freq_metric(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n,
                       int32_t n_fft, int32_t cp, const int32_t *bins, const void *templ_c128, int32_t nbins,
                       double templ_energy, int32_t out_f64, void *metric, int64_t out_stride, void *stream);
/* Multi-root correlator bank on the tensor cores (one fused kernel: sliding-DFT producers -> tcgen05.mma kind::f16 with
 * FP32 accumulators in TMEM -> epilogue): for every capture and every root r < n_roots (<= 128; 64 per pass), the maximum
 * over all candidate offsets o of
 *     | sum_j conj(T[r, j]) bins[j, o] |^2 / (E_r * E(o))
 * i.e. zc_freq.compute_frequency_metric (zc_freq.py:62-99; np.vdot at :94) evaluated for a bank of templates at once.
 * x: complex64 (n_frames, n) on the device, one branch.  bins: int32[nbins] (device), templ: complex64[n_roots][nbins]
 * (device), nbins <= 64, n_fft a multiple of 32.  best_metric / best_offset: [n_frames][n_roots].
 * FP16 operands (10-bit mantissa), FP32 accumulation: |d metric| <= 5e-3 * max(metric); offsets whose in-band energy is
 * below 1e-7 of the largest seen in their stretch of the capture count as silence (metric 0, as the reference's eps clamp). */
int ofs_zc_bank(const void *x_c64, int64_t n_frames, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                const void *templ_c64, int32_t nbins, int32_t n_roots, float *best_metric, int32_t *best_offset,
                void *stream);

/* zc_freq.compute_frequency_metric through the bank kernel (one template): the whole metric row, float32, for complex64
 * single-branch captures.  FP16 tensor-core operands: |d metric| <= 5e-3 * max(metric) (ofs_zc_freq_metric is the
 * float64-prefix version for when 1e-11 is wanted).  bins int32[nbins], templ complex64[nbins] on the device. */
int ofs_zc_freq_metric_fast(const void *x_c64, int64_t n_frames, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                            const void *templ_c64, int32_t nbins, double templ_energy, float *metric,
                            int64_t out_stride, void *stream);

/* zc_freq.compute_frequency_metric at float32 accuracy (|d metric| <= 1e-4 * max(metric)) for complex64 single-branch captures:
 * the sliding-DFT recurrence of the bank in packed fp32 with a float32 epilogue (no fp16 operands, no tensor cores) -- ~40x
 * the float64-prefix kernel of ofs_zc_freq_metric.  bins int32[nbins], templ complex64[nbins] on the device, nbins <= 64. */
int ofs_zc_freq_metric_f32(const void *x_c64, int64_t n_frames, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                           const void *templ_c64, int32_t nbins, double templ_energy, float *metric,
                           int64_t out_stride, void *stream);

/* zc_freq.compute_frequency_metric (zc_freq.py:62-99) in FFT form, float32 (|d metric| <= 1e-4 * max(metric)), complex64
 * or int16-IQ captures [frames][branches][n] (branches summed as at :88-97: one inverse transform for the sum of their correlations, one
 * energy path each), n_fft <= 2048, nbins <= 64: np.vdot(template, bins) (:94) and sum(bins) are two n_fft-tap matched
 * filters of the capture (8192-point overlap-save blocks, one forward and two inverse FFTs per block), and the in-band energy
 * sum|bins|^2 (:95) follows E(o+1) = E(o) + 2 Re(conj(sum bins(o)) d(o)) + nbins |d(o)|^2, d(o) = x[o+cp+n_fft] - x[o+cp],
 * anchored by a direct DFT and carried in float64.  ~2.5x ofs_zc_freq_metric_f32.  Arguments as ofs_zc_freq_metric_f32 plus
 * the input dtype (OFS_C64 / OFS_IQ16) and the branch count.  Consecutive 8192-sample blocks of a capture share one CTA and one anchor unless there are too few
 * captures to fill the GPU; the environment variable OFS_ZQF_BLOCKS_PER_ITEM overrides the split (tests use it to force
 * both the per-block anchor and the whole-capture carry). */
int ofs_zc_freq_metric_fft(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                           const void *templ_c64, int32_t nbins, double templ_energy, float *metric,
                           int64_t out_stride, void *stream);

/* Impairment chain = channel.apply_channel (channel.py:51-98) -> core.apply_cfo (core.py:123-138) ->
 * sync_aa.quantize_adc (sync_aa.py:263-291), batched on the device.
 *   tx: complex64 / complex128 [n_rows][n_tx]; taps (optional): complex128[n_taps] -> full convolution, n_out = n_tx+n_taps-1
 *   (no taps: n_out = n_tx).  Stream s reads faded row row_of_stream[s] (NULL: s), adds std_s * unit_noise[s][.] with
 *   std_s = sqrt(mean|faded row|^2 / 10^(snr_db[s]/10) / 2) (unit_noise NULL: no noise), rotates by
 *   exp(j 2 pi cfo_hz[s] n / fs) (cfo_hz NULL: none) and, when full_scale is given, quantises to `bits` bits:
 *   out gets the quantised samples, out_iq (optional) the integer codes as int16 IQ.
 *   All per-stream arrays live on the device.  faded_ws: [n_rows][n_out] samples of `dtype`; power_ws: double[n_rows]. */
int ofs_channel_apply(const void *tx, int32_t dtype, int64_t n_rows, int64_t n_tx, const void *taps_c128, int32_t n_taps,
                      int64_t n_streams, const int32_t *row_of_stream, const void *unit_noise, int64_t noise_stride,
                      const double *snr_db, const double *cfo_hz, double fs, const double *full_scale, int32_t bits,
                      void *out, int16_t *out_iq, int64_t out_stride, void *faded_ws, double *power_ws, void *stream);

/* 12-bit wire formats of the RTL side <-> the int16 IQ layout [n_channels][n][2] the OFS_IQ16 kernels ingest.
 *   OFS_WIRE_HEX24   uint32 words, {Re[11:0], Im[11:0]} with Re in the upper 12 bits -- docs/preamble_test_vector.hex; 1 channel
 *   OFS_WIRE_AXIS48  uint64 words, {ch1_q, ch1_i, ch0_q, ch0_i} x 12 bits, ch0_i lowest -- ref/test_minn_preamble_detector.py:41-47,
 *                    minn_preamble_detector.sv:23,98-101; 2 channels
 * pack keeps the low 12 bits of every component (two's complement), unpack sign-extends them.  Device pointers. */
#define OFS_WIRE_HEX24 0
#define OFS_WIRE_AXIS48 1
int ofs_wire_pack(const int16_t *iq, int64_t n, int32_t format, void *words, void *stream);
int ofs_wire_unpack(const void *words, int64_t n, int32_t format, int16_t *iq, void *stream);

/* CP-correlation CFO estimators, one result per frame; x: (n_frames, n_branches, n), branches summed.
 *   mode 0  core.estimate_cfo_from_cp               core.py:179-196   P = sum_n x[start+n] conj(x[start+n+N]), n < cp_len
 *   mode 1  core.estimate_cfo_from_cp_robust        core.py:199-230   sum of P(d), window win_len, d in [start-span, start+span)
 *   mode 2  core.estimate_cfo_from_cp_peak(_with_index) / find_cp_start_via_corr  core.py:233-336   first max of |P(d)|
 * cfo_hz = -angle(P) fs / (2 pi N); best_d (optional): the offset used (mode 2), else start; P_c128 (optional).
 * A frame whose windows leave the capture gets cfo = NaN, best_d = -1 (the reference raises on the mismatched slices). */
int ofs_cp_cfo(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n, int64_t x_frame_stride,
               int64_t x_branch_stride, const int64_t *starts, int32_t n_fft, int32_t cp_len, int32_t span, int32_t win_len,
               int32_t mode, double fs, double *cfo_hz, int64_t *best_d, void *P_c128, void *stream);

/* Receive chain after the detector, one result per frame: apply_cfo(rx, -cfo_hz) (core.py:123-138), branch
 * mean, pilot and data symbol FFT + used bins (core.ofdm_fft_used, core.py:171-176), LS channel estimate (core.py:339-341),
 * phase-slope timing (core.py:443-469), equalise (core.py:344-345), complex-gain alignment (core.py:357-362), EVM
 * (core.py:365-370).  x: (n_frames, n_branches, n); the pilot symbol's CP starts at pilot_cp_start[f], the data symbol follows
 * it (sc.py:286-309).  bins: DFT bin numbers of the used subcarriers (device int32[n_used]); k_index: the same as signed
 * indices (device float64[n_used]); pilot_used: complex128[n_used]; data_used: complex128[n_frames][data_stride] (stride 0:
 * shared).  n_fft must divide 4096.  Outputs: h_est, xhat (aligned): complex128[n_frames][n_used]; scalars:
 * float64[n_frames][8] = evm_rms, evm_db, slope (rad/bin), timing offset (samples), gain re, gain im, 0, valid.
 * Symbols that run past the end of the capture are zero-padded (numpy slicing + np.fft.fft(td, n=N)); a start outside the
 * capture gives valid = 0 and NaN. */
int ofs_rx_chain(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n, int64_t x_frame_stride,
                 int64_t x_branch_stride, const int64_t *pilot_cp_start, const double *cfo_hz, double fs, int32_t n_fft,
                 int32_t cp_len, const int32_t *bins, const double *k_index, int32_t n_used, const void *pilot_used_c128,
                 const void *data_used_c128, int64_t data_stride, void *h_est_c128, void *xhat_c128, double *scalars,
                 void *stream);

/* End-to-end sync over HOST buffers --------------------------------------------------------------- */
typedef struct ofs_ctx ofs_ctx;
int ofs_ctx_create(ofs_ctx **ctx, int device);
void ofs_ctx_destroy(ofs_ctx *ctx);
void *ofs_host_alloc(size_t bytes);  /* pinned */
void ofs_host_free(void *p);

/* status bits of a sync record (exact mode) */
#define OFS_ST_EXACT 1       /* a decision lay inside the float32 band and was re-evaluated in float64 from the samples */
#define OFS_ST_CHANGED 2     /* ... and the float64 evaluation moved the index */
#define OFS_ST_UNRESOLVED 4  /* a decision inside the band could not be settled in the kernel (more than 64 candidates, or one
                                of the reference's fallback branches): timing is the float32 decision; ofs_sync_f64 settles it */
#define OFS_EXACT_BAND 1e-5  /* default relative half-width of the band: 8x the largest error of the (smoothed) float32 metric
                                measured where decisions are taken (>= 0.45 of the row maximum: 5.4e-7, profiles/
                                r2_exact_band_probe.json; tests/test_gpu_exact.py asserts <= 2.5e-6 = band / 4) */

typedef struct ofs_sync_record {
    int64_t timing;      /* SC: plateau_end; MINN: peak */
    int64_t coarse;      /* SC: max(plateau_end - delta, 0); MINN: peak */
    float metric;        /* M at the timing index */
    float p_re, p_im;    /* P at the coarse/timing index (float64 recompute, stored as float) */
    float cfo;           /* -angle(P)/(2*pi*lag) cycles/sample (lag = N/2 for SC, N/4 for MINN) */
    int32_t status;      /* OFS_ST_* */
    int32_t reserved;
} ofs_sync_record;

typedef struct ofs_sync_params {
    int32_t cp_len;        /* SC: sc.find_plateau_end_from_metric(M, cp_len, lookahead = cp_len / 4, smooth_win), sc.py:81-146 */
    int32_t smooth_win;    /* SC: np.convolve "same" window; MINN: trailing average (minn.py:115-128) */
    int32_t sc_delta;      /* SC: coarse = max(plateau_end - delta, 0), sc.py:211 */
    int32_t exact;         /* 0: decide on the float32 metric.  1: every comparison of the detector whose operands lie within
                              exact_band of each other is re-evaluated in float64 from the samples, so that timing equals the
                              index the reference finds on its float64 metric (status says when that happened) */
    double gate_threshold; /* MINN: minn.find_minn_peak gate_threshold, minn.py:131-205 */
    double exact_band;     /* <= 0: OFS_EXACT_BAND */
} ofs_sync_params;

/* One call = "sync metric + CFO" for a batch of frames (the BASELINE.json headline):
 *   metric (ofs_metric, stripe path) -> detector (SC: plateau, MINN: find_minn_peak) -> P at the
 *   detected index -> CFO.  Device version: x, M, records are device pointers; branches are summed (sc.py:73-74).
 *   scratch: int64[4 * n_frames]. */
int ofs_sync(const ofs_metric_desc *d, const void *x, float *M, float *chunk_max, int64_t cm_stride,
             const ofs_sync_params *params, ofs_sync_record *records, int64_t *scratch, void *stream);
/* Second half of ofs_sync alone (detector + P/CFO records on an already computed metric). */
int ofs_sync_detect(const ofs_metric_desc *d, const void *x, const float *M, const float *chunk_max, int64_t cm_stride,
                    const ofs_sync_params *params, ofs_sync_record *records, int64_t *scratch, void *stream);
/* The same pipeline entirely in float64 (precise tile kernel with float64 outputs -> float64 detector -> records), any branch
 * count / dtype / symbol length: the path frames flagged OFS_ST_UNRESOLVED are re-run through.  Workspace (8 bytes per output)
 * comes from the stream's memory pool.  records: device pointer, status = OFS_ST_EXACT. */
int ofs_sync_f64(const ofs_metric_desc *d, const void *x, const ofs_sync_params *params, ofs_sync_record *records, void *stream);
/* Host version: x_host (n_frames x n_samples, dtype per d->in_dtype; x_frame_stride in SAMPLES), M_host optional (float32
 * [n_frames][out_stride]); records_host[n_frames].  Frames are pipelined through the ctx workspace
 * in batches (H2D, kernels, D2H overlapped on three streams, two buffer sets).  A batch holds 32 MB of
 * samples (environment variable OFS_HOST_BATCH_MB overrides, 1..4096); pinned host memory (ofs_host_alloc)
 * is what lets the copies run asynchronously.  M_host == NULL: only the records come back.  With params->exact, frames the
 * kernels flag OFS_ST_UNRESOLVED are re-run through ofs_sync_f64 before the call returns. */
int ofs_sync_host(ofs_ctx *ctx, const ofs_metric_desc *d, const void *x_host, float *M_host,
                  const ofs_sync_params *params, ofs_sync_record *records_host);
/* kernels launched by this library on the calling thread since load (for bench.py's gpu_launches) */
int64_t ofs_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* OFDMSYNC_H */

/*
 * ofs_oracle.c -- CPU restatement (plain C, float64 / int64) of the hot path of
 * amcolex/ofdm-sync-math.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load it.  The product path
 * (ofdm_sync_math_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py)
 * against outputs of the unmodified reference captured by oracle/gen_golden.py into
 * tests/golden/*.npz, and against the reference's own docs/*_test_vector.{csv,hex}.
 *
 * Each function cites the reference file:line it follows.  Loops follow the reference's
 * operation order where the reference is a scalar recurrence (sc.py, sync_aa.py, minn_rtl.py,
 * zc_v2.py); where the reference calls a numpy reduction (np.sum, np.convolve, np.fft) the
 * summation order inside numpy is unspecified and a plain sequential sum is used (agreement
 * ~1e-13 relative, stated in the tests).  Build with -ffp-contract=off so that no FMA is
 * formed (numpy scalar arithmetic has none).
 *
 * Complex arrays are interleaved (re, im) doubles == numpy complex128 layout.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline double sq_abs(double re, double im)
{
    /* np.abs(z) ** 2 (sc.py:60,70; sync_aa.py:475,502).  numpy's complex abs is its own SIMD
     * hypot whose last bit matches neither glibc hypot() nor sqrt(re^2+im^2) (measured: 62-64 %
     * of samples equal either way), so |z|^2 is NOT bit-reproducible outside numpy; the plain
     * re^2+im^2 is used and R / M are compared to a few ulp (tests state 4e-15). */
    return re * re + im * im;
}

/* ---------------------------------------------------------------------------------------
 * Schmidl-Cox streaming metric -- sc.py:42-78.
 * x: (nb, L) complex128.  Outputs length out_len = max(L-N+1,0).  Returns out_len.
 * Recursive update in the reference's order (sc.py:65-72); R over the SECOND half only.
 * ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_sc_metric(const double *x, int64_t nb, int64_t L, int64_t N,
                              double *M, double *P, double *R)
{
    int64_t half = N / 2;
    int64_t out_len = L - N + 1;
    if (out_len <= 0) return 0;
    memset(P, 0, sizeof(double) * 2 * (size_t)out_len);
    memset(R, 0, sizeof(double) * (size_t)out_len);
    for (int64_t b = 0; b < nb; ++b) {
        const double *xb = x + 2 * b * L;
        double pre = 0.0, pim = 0.0, r = 0.0;
        for (int64_t m = 0; m < half; ++m) {              /* sc.py:59-60 */
            double ar = xb[2 * m], ai = xb[2 * m + 1];
            double br = xb[2 * (m + half)], bi = xb[2 * (m + half) + 1];
            pre += ar * br + ai * bi;
            pim += ai * br - ar * bi;
            r += sq_abs(br, bi);
        }
        P[0] += pre; P[1] += pim; R[0] += r;
        for (int64_t d = 1; d < out_len; ++d) {           /* sc.py:65-72 */
            double ar = xb[2 * (d - 1)], ai = xb[2 * (d - 1) + 1];
            double br = xb[2 * (d - 1 + half)], bi = xb[2 * (d - 1 + half) + 1];
            double cr = xb[2 * (d - 1 + N)], ci = xb[2 * (d - 1 + N) + 1];
            /* P = P - old_a*conj(old_b) + old_b*conj(new_b) */
            double t1r = ar * br + ai * bi, t1i = ai * br - ar * bi;
            double t2r = br * cr + bi * ci, t2i = bi * cr - br * ci;
            pre = (pre - t1r) + t2r;
            pim = (pim - t1i) + t2i;
            r = (r - sq_abs(br, bi)) + sq_abs(cr, ci);
            P[2 * d] += pre; P[2 * d + 1] += pim; R[d] += r;
        }
    }
    for (int64_t d = 0; d < out_len; ++d) {               /* sc.py:76-77 */
        double a2 = sq_abs(P[2 * d], P[2 * d + 1]);
        double rr = R[d] > 1e-12 ? R[d] : 1e-12;
        M[d] = a2 / (rr * rr);
    }
    return out_len;
}

/* ---------------------------------------------------------------------------------------
 * S&C metric with R over BOTH halves -- combined_sc_min.py:116-164 (direct O(N) per d).
 * ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_scb_metric(const double *x, int64_t nb, int64_t L, int64_t N,
                               double *M, double *P, double *R)
{
    int64_t half = N / 2;
    if (half == 0 || N > L) return 0;
    int64_t out_len = L - N + 1;
    memset(P, 0, sizeof(double) * 2 * (size_t)out_len);
    memset(R, 0, sizeof(double) * (size_t)out_len);
    for (int64_t b = 0; b < nb; ++b) {
        const double *xb = x + 2 * b * L;
        for (int64_t d = 0; d < out_len; ++d) {
            double pre = 0, pim = 0, r = 0;
            for (int64_t m = 0; m < half; ++m) {
                double ar = xb[2 * (d + m)], ai = xb[2 * (d + m) + 1];
                double br = xb[2 * (d + half + m)], bi = xb[2 * (d + half + m) + 1];
                pre += ar * br + ai * bi;
                pim += ai * br - ar * bi;
                r += sq_abs(ar, ai) + sq_abs(br, bi);
            }
            P[2 * d] += pre; P[2 * d + 1] += pim; R[d] += r;
        }
    }
    for (int64_t d = 0; d < out_len; ++d) {
        double a2 = sq_abs(P[2 * d], P[2 * d + 1]);
        double rr = R[d] > 1e-12 ? R[d] : 1e-12;
        M[d] = a2 / (rr * rr);
    }
    return out_len;
}

/* ---------------------------------------------------------------------------------------
 * Minn metric -- minn.py:59-112 (N = N_FFT) and minn.py:697-751 (parameterised symbol_len);
 * identical copy at combined_sc_min.py:60-113.  Direct O(Q) sums per d.
 * ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_minn_metric(const double *x, int64_t nb, int64_t L, int64_t N,
                                double *M, double *P, double *R)
{
    int64_t Q = N / 4;
    int64_t out_len = L - N + 1;
    if (out_len <= 0) return 0;
    memset(P, 0, sizeof(double) * 2 * (size_t)out_len);
    memset(R, 0, sizeof(double) * (size_t)out_len);
    for (int64_t b = 0; b < nb; ++b) {
        const double *xb = x + 2 * b * L;
        for (int64_t d = 0; d < out_len; ++d) {
            double c1r = 0, c1i = 0, c2r = 0, c2i = 0, r = 0;
            for (int64_t m = 0; m < Q; ++m) {
                const double *q0 = xb + 2 * (d + m), *q1 = xb + 2 * (d + Q + m);
                const double *q2 = xb + 2 * (d + 2 * Q + m), *q3 = xb + 2 * (d + 3 * Q + m);
                c1r += q0[0] * q1[0] + q0[1] * q1[1];
                c1i += q0[1] * q1[0] - q0[0] * q1[1];
                c2r += q2[0] * q3[0] + q2[1] * q3[1];
                c2i += q2[1] * q3[0] - q2[0] * q3[1];
                r += sq_abs(q1[0], q1[1]) + sq_abs(q2[0], q2[1]) + sq_abs(q3[0], q3[1]);
            }
            P[2 * d] += c1r + c2r; P[2 * d + 1] += c1i + c2i; R[d] += r;
        }
    }
    for (int64_t d = 0; d < out_len; ++d) {               /* minn.py:109-111 */
        double ar = P[2 * d] > 0.0 ? P[2 * d] : 0.0;
        double rr = R[d] > 1e-12 ? R[d] : 1e-12;
        M[d] = (ar * ar) / (rr * rr);
    }
    return out_len;
}

/* ---------------------------------------------------------------------------------------
 * Park metric -- park.py:64-114.  P(d)=sum_{k<h} x[d-k]*x[d+k] (no conjugate),
 * E(d)=sum_{k<h}|x[d+k]|^2, d in [h, L-h-1].  Returns n = L-2h (0 if L < 2h+1).
 * ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_park_metric(const double *x, int64_t nb, int64_t L, int64_t N,
                                int64_t *ds, double *M, double *P, double *E)
{
    int64_t half = N / 2;
    if (half == 0 || L < 2 * half + 1) return 0;
    int64_t lo = half, hi = L - half - 1;
    if (hi < lo) return 0;
    int64_t n = hi - lo + 1;
    memset(P, 0, sizeof(double) * 2 * (size_t)n);
    memset(E, 0, sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) ds[i] = lo + i;
    for (int64_t b = 0; b < nb; ++b) {
        const double *xb = x + 2 * b * L;
        for (int64_t i = 0; i < n; ++i) {
            int64_t d = lo + i;
            double pr = 0, pi = 0, e = 0;
            for (int64_t k = 0; k < half; ++k) {
                const double *f = xb + 2 * (d + k), *w = xb + 2 * (d - k);
                pr += w[0] * f[0] - w[1] * f[1];
                pi += w[0] * f[1] + w[1] * f[0];
                e += sq_abs(f[0], f[1]);
            }
            P[2 * i] += pr; P[2 * i + 1] += pi; E[i] += e;
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        double a2 = sq_abs(P[2 * i], P[2 * i + 1]);
        double ee = E[i] > 1e-12 ? E[i] : 1e-12;
        M[i] = a2 / (ee * ee);
    }
    return n;
}

/* ---------------------------------------------------------------------------------------
 * Causal trailing average -- minn.py:115-128 == combined_sc_min.py:167-180.
 * ------------------------------------------------------------------------------------- */
ORC_API void orc_trailing_average(const double *x, int64_t n, int64_t win, double *y)
{
    if (win <= 1) { memcpy(y, x, sizeof(double) * (size_t)n); return; }
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        acc += x[i];
        if (i >= win) acc -= x[i - win];
        double denom = (i >= win - 1) ? (double)win : (double)(i + 1);
        y[i] = acc / denom;
    }
}

/* np.convolve(M, ones(w)/w, mode="same") -- sc.py:100.  Output length max(n, w). */
static int64_t conv_same_box(const double *M, int64_t n, int64_t w, double *out)
{
    double h = 1.0 / (double)w;
    int64_t nl = n > w ? n : w, ns = n > w ? w : n;
    int64_t off = (ns - 1) / 2;          /* numpy 'same': centred w.r.t. the longer input */
    for (int64_t i = 0; i < nl; ++i) {
        int64_t k = i + off;             /* index into the full convolution */
        double s = 0.0;
        int64_t jlo = k - w + 1 < 0 ? 0 : k - w + 1;
        int64_t jhi = k < n - 1 ? k : n - 1;
        for (int64_t j = jlo; j <= jhi; ++j) s += M[j] * h;
        out[i] = s;
    }
    return nl;
}

static int64_t argmax_first(const double *v, int64_t n)
{
    int64_t best = 0;
    for (int64_t i = 1; i < n; ++i) if (v[i] > v[best]) best = i;
    return best;
}

/* ---------------------------------------------------------------------------------------
 * Plateau end -- sc.py:81-146.  lookahead < 0 means None.  Ms_out (optional) gets the
 * smoothed metric (length max(n, w)).
 * ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_find_plateau_end(const double *M, int64_t n, int64_t cp_len, int64_t lookahead,
                                     int64_t smooth_win, double *Ms_out)
{
    if (n == 0) return 0;
    int64_t Lk = lookahead < 0 ? cp_len / 4 : (lookahead > 1 ? lookahead : 1);
    int64_t w = smooth_win > 1 ? smooth_win : 1;
    int64_t cap = n > w ? n : w;
    double *Ms = (double *)malloc(sizeof(double) * (size_t)cap);
    int64_t ms = conv_same_box(M, n, w, Ms);
    if (Ms_out) memcpy(Ms_out, Ms, sizeof(double) * (size_t)ms);
    int64_t result;
    int64_t center = argmax_first(Ms, ms);                    /* sc.py:106 */
    int64_t post_hi = ms < center + cp_len ? ms : center + cp_len;
    if (post_hi > center + 1) {
        double thr_local = 0.95 * Ms[center];
        for (int64_t i = center; i < post_hi; ++i)
            if (Ms[i] <= thr_local) { result = i; goto done; }
    }
    {
        int64_t min_run = cp_len / 2 > 8 ? cp_len / 2 : 8;     /* sc.py:117 */
        double peak = Ms[center];
        if (peak > 0) {
            double thr = 0.6 * peak;
            int64_t i = 0;
            while (i < ms) {
                if (Ms[i] >= thr) {
                    int64_t s = i;
                    while (i < ms && Ms[i] >= thr) ++i;
                    if (i - s >= min_run) { result = i - 1; goto done; }
                } else ++i;
            }
        }
    }
    {
        int64_t lo = center - cp_len > 0 ? center - cp_len : 0;     /* sc.py:136-146 */
        int64_t hi = ms - Lk - 1 < center + cp_len ? ms - Lk - 1 : center + cp_len;
        /* python slicing Ms[lo:hi], Ms[lo+L:hi+L]; negative hi wraps in python -- mirror it */
        int64_t hi_eff = hi;
        if (hi_eff < 0) hi_eff += ms;
        if (hi_eff < 0) hi_eff = 0;
        if (hi_eff > ms) hi_eff = ms;
        int64_t wl = hi_eff - lo;
        int64_t hi2 = hi + Lk, lo2 = lo + Lk;
        if (hi2 < 0) hi2 += ms;
        if (hi2 < 0) hi2 = 0;
        if (hi2 > ms) hi2 = ms;
        if (lo2 > ms) lo2 = ms;
        int64_t al = hi2 - lo2;
        if (wl <= 0 || al <= 0 || wl != al) {
            /* empty (or non-broadcastable) window -> reference returns center on empty drop */
            result = center;
            goto done;
        }
        int64_t best = 0; double bestv = Ms[lo] - Ms[lo2];
        for (int64_t i = 1; i < wl; ++i) {
            double dv = Ms[lo + i] - Ms[lo2 + i];
            if (dv > bestv) { bestv = dv; best = i; }
        }
        result = lo + best + Lk / 2;
    }
done:
    free(Ms);
    return result;
}

/* ---------------------------------------------------------------------------------------
 * find_minn_peak -- minn.py:131-205.  Returns peak (or -1: empty, -2: non-positive peak).
 * gate_mask (uint8[n]) and Ms (double[n]) are outputs.  bounds_lo/hi: use has_bounds flag.
 * ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_find_minn_peak(const double *M, int64_t n, int64_t smooth_win, double gate_threshold,
                                   int has_bounds, int64_t b_lo, int64_t b_hi,
                                   uint8_t *gate_mask, double *Ms)
{
    if (n == 0) return -1;
    int64_t w = smooth_win > 1 ? smooth_win : 1;
    double *mp = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) mp[i] = M[i] > 0.0 ? M[i] : 0.0;
    orc_trailing_average(mp, n, w, Ms);
    free(mp);
    double max_ms = Ms[0];
    for (int64_t i = 1; i < n; ++i) if (Ms[i] > max_ms) max_ms = Ms[i];
    if (max_ms <= 0.0) return -2;
    double level = gate_threshold * max_ms;
    int any = 0;
    for (int64_t i = 0; i < n; ++i) { gate_mask[i] = Ms[i] >= level; any |= gate_mask[i]; }
    if (any) {                                               /* minn.py:159-182 */
        int in_seg = 0; int64_t bs = 0, be = 0, bl = 0, st = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (gate_mask[i] && !in_seg) { in_seg = 1; st = i; }
            else if (!gate_mask[i] && in_seg) {
                in_seg = 0;
                if (i - st > bl) { bl = i - st; bs = st; be = i; }
            }
        }
        if (in_seg && n - st > bl) { bl = n - st; bs = st; be = n; }
        if (bl > 0) {
            memset(gate_mask, 0, (size_t)n);
            memset(gate_mask + bs, 1, (size_t)(be - bs));
        }
    }
    if (has_bounds) {                                        /* minn.py:186-193 */
        int64_t s = b_lo > 0 ? b_lo : 0, e = b_hi < n ? b_hi : n;
        if (s >= e) { s = 0; e = n; }
        for (int64_t i = 0; i < n; ++i) if (i < s || i >= e) gate_mask[i] = 0;
    }
    any = 0;
    for (int64_t i = 0; i < n; ++i) any |= gate_mask[i];
    if (!any) {                                              /* minn.py:195-200 */
        int64_t pk = argmax_first(Ms, n);
        memset(gate_mask, 0, (size_t)n);
        gate_mask[pk] = 1;
        return pk;
    }
    int64_t best = -1;
    for (int64_t i = 0; i < n; ++i)
        if (gate_mask[i] && (best < 0 || Ms[i] > Ms[best])) best = i;
    return best;
}

/* ---------------------------------------------------------------------------------------
 * Gated Minn peak -- combined_sc_min.py:183-259.  Returns peak, or -1 empty M (reference
 * returns 0), -3 empty gate region (ValueError).  Ms is an output (n doubles).
 * ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_find_minn_peak_gated(const double *M, int64_t n, int64_t smooth_win,
                                         const uint8_t *gate, int has_bounds, int64_t b_lo, int64_t b_hi,
                                         double *Ms)
{
    if (n == 0) return -1;
    uint8_t *sm = (uint8_t *)malloc((size_t)n);
    memcpy(sm, gate, (size_t)n);
    if (has_bounds) {
        int64_t s = b_lo > 0 ? b_lo : 0, e = b_hi < n ? b_hi : n;
        if (s >= e) { s = 0; e = n; }
        for (int64_t i = 0; i < n; ++i) if (i < s || i >= e) sm[i] = 0;
    }
    int any = 0;
    for (int64_t i = 0; i < n; ++i) any |= sm[i];
    if (!any) { free(sm); return -3; }
    int64_t w = smooth_win > 1 ? smooth_win : 1;
    double *mp = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) mp[i] = M[i] > 0.0 ? M[i] : 0.0;
    orc_trailing_average(mp, n, w, Ms);
    free(mp);
    /* _streaming_peak_detector, combined_sc_min.py:183-209 */
    int64_t best = -1; double bv = -INFINITY; int active = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (sm[i]) {
            if (!active) { active = 1; bv = Ms[i]; best = i; }
            else if (Ms[i] > bv) { bv = Ms[i]; best = i; }
        } else if (active) break;
    }
    free(sm);
    return best;
}

/* S&C gate construction -- combined_sc_min.py:337-351.  gate: uint8[n] out. */
ORC_API void orc_sc_gate(const double *Msc, int64_t n, double thr, uint8_t *gate)
{
    if (n == 0) return;
    double mx = Msc[0];
    for (int64_t i = 1; i < n; ++i) if (Msc[i] > mx) mx = Msc[i];
    int any = 0;
    for (int64_t i = 0; i < n; ++i) {
        double v = mx > 0 ? Msc[i] / mx : Msc[i];
        gate[i] = v >= thr; any |= gate[i];
    }
    if (!any) { memset(gate, 0, (size_t)n); gate[argmax_first(Msc, n)] = 1; }
}

/* ---------------------------------------------------------------------------------------
 * ZC matched filter: np.convolve(x, conj(ref[::-1]))  (zc.py:116, zc_v2.py:244-254) and the
 * sliding energy np.convolve(|x|^2, ones(nr)) (zc.py:117, zc_v2.py:268).  Full length L+nr-1.
 * corr[k] = sum_j x[j] * conj(ref[nr-1-(k-j)]),  0 <= k-j < nr.
 * ------------------------------------------------------------------------------------- */
ORC_API void orc_zc_matched_filter(const double *x, int64_t L, const double *ref, int64_t nr,
                                   double *corr, double *energy)
{
    int64_t n = L + nr - 1;
    for (int64_t k = 0; k < n; ++k) {
        double cr = 0, ci = 0, e = 0;
        int64_t jlo = k - nr + 1 < 0 ? 0 : k - nr + 1;
        int64_t jhi = k < L - 1 ? k : L - 1;
        for (int64_t j = jlo; j <= jhi; ++j) {
            const double *r = ref + 2 * (nr - 1 - (k - j));
            double xr = x[2 * j], xi = x[2 * j + 1];
            cr += xr * r[0] + xi * r[1];
            ci += xi * r[0] - xr * r[1];
            e += sq_abs(xr, xi);
        }
        corr[2 * k] = cr; corr[2 * k + 1] = ci; energy[k] = e;
    }
}

/* zc_v2.RunningSum over |corr| + threshold -- zc_v2.py:191-238, 288-336.  Sequential recurrence
 * in the reference's order (sum_acc + sample - oldest). */
ORC_API void orc_zc_streaming_detection(const double *mag, int64_t n, int64_t window, int64_t thresh_value,
                                        int64_t frac_bits, double min_mag,
                                        double *local_sum, uint8_t *valid, uint8_t *above)
{
    int64_t W = window > 1 ? window : 1;
    double *buf = (double *)calloc((size_t)W, sizeof(double));
    int64_t ptr = 0, filled = 0; double acc = 0.0; int v = 0;
    double scale = (double)((int64_t)1 << frac_bits);
    for (int64_t i = 0; i < n; ++i) {
        double oldest = filled >= W ? buf[ptr] : 0.0;
        int dv = filled >= W;
        buf[ptr] = mag[i];
        ptr = (ptr + 1) % W;
        if (filled < W) ++filled;
        if (dv) { acc = acc + mag[i] - oldest; v = 1; }
        else acc = acc + mag[i];
        local_sum[i] = acc; valid[i] = (uint8_t)v;
        above[i] = (uint8_t)(v && (mag[i] * scale >= acc * (double)thresh_value) && (mag[i] >= min_mag));
    }
    free(buf);
}

/* zc_v2.detect_zc_peaks -- zc_v2.py:360-450.  events: int64[max_ev][4] = peak, gate_start,
 * gate_end, detected_start; values: double[max_ev].  Returns number of events (may exceed
 * max_ev; only the first max_ev are stored). */
ORC_API int64_t orc_detect_zc_peaks(const double *mag, const uint8_t *valid, const uint8_t *above, int64_t n,
                                    int64_t ref_len, int64_t hysteresis, uint8_t *gate_mask,
                                    int64_t *events, double *values, int64_t max_ev)
{
    memset(gate_mask, 0, (size_t)n);
    int open = 0; int64_t gs = 0, pk = 0, low = 0, nev = 0; double pv = 0.0;
    int64_t hyst_limit = hysteresis - 1 > 0 ? hysteresis - 1 : 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        if (!open) {
            if (above[i]) { open = 1; gs = i; pk = i; pv = mag[i]; low = 0; }
        } else {
            gate_mask[i] = 1;
            if (mag[i] > pv) { pv = mag[i]; pk = i; }
            if (above[i]) low = 0;
            else if (hysteresis == 0 || low >= hyst_limit) {
                if (nev < max_ev) {
                    int64_t ds_ = pk - ref_len + 1; if (ds_ < 0) ds_ = 0;
                    events[4 * nev] = pk; events[4 * nev + 1] = gs; events[4 * nev + 2] = i;
                    events[4 * nev + 3] = ds_; values[nev] = pv;
                }
                ++nev; open = 0; pv = 0.0; low = 0;
            } else ++low;
        }
    }
    if (open) {
        if (nev < max_ev) {
            int64_t ds_ = pk - ref_len + 1; if (ds_ < 0) ds_ = 0;
            events[4 * nev] = pk; events[4 * nev + 1] = gs; events[4 * nev + 2] = n;
            events[4 * nev + 3] = ds_; values[nev] = pv;
        }
        ++nev;
        for (int64_t i = gs; i < n; ++i) gate_mask[i] = 1;
    }
    return nev;
}

/* ---------------------------------------------------------------------------------------
 * zc_freq metric -- zc_freq.py:62-99, restated as the direct DFT of the nbins used bins
 * (positions[j] = (N/2 + bin_indices[j]) % N of the fftshifted spectrum == DFT bin
 * (positions[j] + N/2) % N); agreement with np.fft ~2e-15 relative.
 * Returns num_offsets (<=0: the reference raises ValueError, zc_freq.py:76-78).
 * ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_zc_freq_metric(const double *x, int64_t nb, int64_t L, int64_t N, int64_t cp,
                                   const int64_t *bin_indices, const double *templ, int64_t nbins,
                                   double templ_energy, double *metric)
{
    int64_t nofs = L - (N + cp) + 1;
    if (nofs <= 0) return nofs;
    double *tw = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    for (int64_t k = 0; k < N; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)N;
        tw[2 * k] = cos(a); tw[2 * k + 1] = sin(a);
    }
    int64_t *kk = (int64_t *)malloc(sizeof(int64_t) * (size_t)nbins);
    for (int64_t j = 0; j < nbins; ++j) {
        int64_t pos = ((N / 2 + bin_indices[j]) % N + N) % N;   /* index after fftshift */
        kk[j] = (pos + N - N / 2) % N;                            /* plain DFT bin (N even) */
    }
    for (int64_t o = 0; o < nofs; ++o) {
        double cr = 0, ci = 0, es = 0;
        for (int64_t b = 0; b < nb; ++b) {
            const double *s = x + 2 * (b * L + o + cp);
            for (int64_t j = 0; j < nbins; ++j) {
                double br = 0, bi = 0; int64_t k = kk[j], ph = 0;
                for (int64_t n = 0; n < N; ++n) {
                    double wr = tw[2 * ph], wi = tw[2 * ph + 1];
                    br += s[2 * n] * wr - s[2 * n + 1] * wi;
                    bi += s[2 * n] * wi + s[2 * n + 1] * wr;
                    ph += k; if (ph >= N) ph -= N;
                }
                /* np.vdot(template, bins) = sum conj(t) * bins */
                cr += templ[2 * j] * br + templ[2 * j + 1] * bi;
                ci += templ[2 * j] * bi - templ[2 * j + 1] * br;
                es += sq_abs(br, bi);
            }
        }
        double den = templ_energy * es; if (den < 1e-12) den = 1e-12;
        metric[o] = sq_abs(cr, ci) / den;
    }
    free(tw); free(kk);
    return nofs;
}

/* ---------------------------------------------------------------------------------------
 * [A][A] streaming detector, loop 1 -- sync_aa.py:321-386 (DelayLine / RunningSum /
 * RunningSumReal) + sync_aa.py:458-493.  Literal ring buffers, reference operation order.
 * ------------------------------------------------------------------------------------- */
ORC_API void orc_aa_metric(const double *x, int64_t na, int64_t n, int64_t L,
                           double *P, double *R, double *M, uint8_t *valid)
{
    double *dl = (double *)calloc((size_t)(2 * na * L), sizeof(double));   /* delay line */
    double *pb = (double *)calloc((size_t)(2 * na * L), sizeof(double));   /* P window */
    double *rb = (double *)calloc((size_t)(na * L), sizeof(double));       /* R window */
    double *ps = (double *)calloc((size_t)(2 * na), sizeof(double));
    double *rs = (double *)calloc((size_t)na, sizeof(double));
    int64_t *fill = (int64_t *)calloc((size_t)(3 * na), sizeof(int64_t));
    int64_t ptr = 0;  /* all rings advance in lock-step */
    double noise_floor = 1e-6 * (double)L;
    for (int64_t t = 0; t < n; ++t) {
        double Psr = 0.0, Psi = 0.0, Rs = 0.0; int all_valid = 1;
        for (int64_t a = 0; a < na; ++a) {
            double xr = x[2 * (a * n + t)], xi = x[2 * (a * n + t) + 1];
            double *d = dl + 2 * (a * L + ptr);
            double dr = d[0], di = d[1];
            d[0] = xr; d[1] = xi;
            int dvalid;
            if (fill[3 * a] < L) { ++fill[3 * a]; dvalid = 0; dr = 0.0; di = 0.0; }
            else dvalid = 1;
            /* product = x_n * conj(x_delayed) if delay_valid else 0.0  (sync_aa.py:470) */
            double qr = 0.0, qi = 0.0;
            if (dvalid) { qr = xr * dr + xi * di; qi = xi * dr - xr * di; }
            double *pw = pb + 2 * (a * L + ptr);
            double orr = pw[0], oi = pw[1];
            pw[0] = qr; pw[1] = qi;
            ps[2 * a] = ps[2 * a] + qr - orr;                 /* sync_aa.py:337 */
            ps[2 * a + 1] = ps[2 * a + 1] + qi - oi;
            int pvalid;
            if (fill[3 * a + 1] < L) { ++fill[3 * a + 1]; pvalid = 0; } else pvalid = 1;
            double pw2 = sq_abs(xr, xi);                        /* sync_aa.py:475 */
            double *rw = rb + (a * L + ptr);
            double ro = *rw; *rw = pw2;
            rs[a] = rs[a] + pw2 - ro;
            int rvalid;
            if (fill[3 * a + 2] < L) { ++fill[3 * a + 2]; rvalid = 0; } else rvalid = 1;
            Psr += ps[2 * a]; Psi += ps[2 * a + 1]; Rs += rs[a];
            all_valid = all_valid && pvalid && rvalid;
        }
        ptr = (ptr + 1) % L;
        P[2 * t] = Psr; P[2 * t + 1] = Psi; R[t] = Rs; valid[t] = (uint8_t)all_valid;
        if (all_valid && Rs > noise_floor) {
            double m = sq_abs(Psr, Psi) / (Rs * Rs);
            M[t] = m < 1.0 ? m : 1.0;
        } else M[t] = 0.0;
    }
    free(dl); free(pb); free(rb); free(ps); free(rs); free(fill);
}

/* [A][A] gate/peak FSM, loop 2 -- sync_aa.py:495-568.  ev_i: int64[max_ev][4] = peak, gate_start,
 * gate_end, frame_start; ev_f: double[max_ev][4] = P_re, P_im, M_at_peak, cfo_hz. */
ORC_API int64_t orc_aa_events(const double *P, const double *M, const uint8_t *valid, int64_t n, int64_t L,
                              double threshold, int64_t hysteresis, double sample_rate,
                              int64_t *ev_i, double *ev_f, int64_t max_ev)
{
    int open = 0; int64_t gs = 0, pk = 0, low = 0, nev = 0;
    double pkr = 0, pki = 0, pk2 = 0;
    for (int64_t t = 0; t < n; ++t) {
        if (!valid[t]) continue;
        double m = M[t], p2 = sq_abs(P[2 * t], P[2 * t + 1]);
        if (!open) {
            if (m >= threshold) { open = 1; gs = t; pk = t; pkr = P[2 * t]; pki = P[2 * t + 1]; pk2 = p2; low = 0; }
        } else {
            if (p2 > pk2) { pk = t; pkr = P[2 * t]; pki = P[2 * t + 1]; pk2 = p2; }
            if (m >= threshold) low = 0;
            else {
                ++low;
                if (low >= hysteresis) {
                    if (nev < max_ev) {
                        ev_i[4 * nev] = pk; ev_i[4 * nev + 1] = gs; ev_i[4 * nev + 2] = t;
                        ev_i[4 * nev + 3] = pk - 2 * L + 1;
                        ev_f[4 * nev] = pkr; ev_f[4 * nev + 1] = pki; ev_f[4 * nev + 2] = M[pk];
                        ev_f[4 * nev + 3] = atan2(pki, pkr) * sample_rate / (2 * M_PI * (double)L);
                    }
                    ++nev; open = 0; pk2 = 0.0; low = 0;
                }
            }
        }
    }
    if (open) {
        if (nev < max_ev) {
            ev_i[4 * nev] = pk; ev_i[4 * nev + 1] = gs; ev_i[4 * nev + 2] = n;
            ev_i[4 * nev + 3] = pk - 2 * L + 1;
            ev_f[4 * nev] = pkr; ev_f[4 * nev + 1] = pki; ev_f[4 * nev + 2] = M[pk];
            ev_f[4 * nev + 3] = atan2(pki, pkr) * sample_rate / (2 * M_PI * (double)L);
        }
        ++nev;
    }
    return nev;
}

/* ---------------------------------------------------------------------------------------
 * minn_rtl float mirror -- minn_rtl.py:512-652 (_DelayLine/_RunningSum/_antenna_path) and
 * minn_rtl.py:667-733.  Literal per-sample objects with their valid gating.
 * ------------------------------------------------------------------------------------- */
typedef struct { int64_t depth, wr, fill; double *mem; double last; } dl_t;
typedef struct { int64_t depth, wr, fill; double *mem; double sum; int valid; } rs_t;

static void dl_init(dl_t *d, int64_t depth) { d->depth = depth; d->wr = 0; d->fill = 0; d->last = 0; d->mem = (double *)calloc((size_t)(depth > 0 ? depth : 1), sizeof(double)); }
static void rs_init(rs_t *r, int64_t depth) { r->depth = depth; r->wr = 0; r->fill = 0; r->sum = 0; r->valid = 0; r->mem = (double *)calloc((size_t)(depth > 0 ? depth : 1), sizeof(double)); }

static double dl_step(dl_t *d, double s, int in_valid, int *out_valid)   /* minn_rtl.py:524-542 */
{
    if (d->depth == 0) { if (in_valid) d->last = s; *out_valid = in_valid; return s; }
    if (!in_valid) { *out_valid = 0; return d->last; }
    double rv = d->fill < d->depth ? 0.0 : d->mem[d->wr];
    d->mem[d->wr] = s;
    d->wr = (d->wr + 1) % d->depth;
    if (d->fill < d->depth) { ++d->fill; d->last = 0.0; *out_valid = 0; return 0.0; }
    d->last = rv; *out_valid = 1; return rv;
}

static double rs_step(rs_t *r, double s, int in_valid, int *out_valid)   /* minn_rtl.py:558-580 */
{
    if (r->depth == 0) { if (in_valid) { r->sum = s; r->valid = 1; } *out_valid = r->valid; return r->sum; }
    if (!in_valid) { *out_valid = r->valid; return r->sum; }
    double oldest = r->fill < r->depth ? 0.0 : r->mem[r->wr];
    r->mem[r->wr] = s;
    r->wr = (r->wr + 1) % r->depth;
    r->sum = r->sum + s - oldest;
    if (r->fill < r->depth) { ++r->fill; if (r->fill >= r->depth) r->valid = 1; }
    else r->valid = 1;
    *out_valid = r->valid; return r->sum;
}

/* Outputs (each double[n] unless noted): corr_total, corr_positive, smooth_metric, energy_total,
 * corr_scaled, energy_scaled, metric_valid (u8), above (u8). */
ORC_API void orc_minn_rtl_metric(const double *x, int64_t nb, int64_t n, int64_t Q, int64_t smooth_shift,
                                 int64_t threshold_value, int64_t frac_bits,
                                 double *corr_total, double *corr_positive, double *smooth_metric,
                                 double *energy_total, double *corr_scaled, double *energy_scaled,
                                 uint8_t *metric_valid, uint8_t *above)
{
    for (int64_t i = 0; i < n; ++i) { corr_total[i] = 0; energy_total[i] = 0; metric_valid[i] = 1; }
    for (int64_t b = 0; b < nb; ++b) {
        dl_t di, dq, cd, e1, e2; rs_t cw, ew;
        dl_init(&di, Q); dl_init(&dq, Q); dl_init(&cd, Q); dl_init(&e1, Q); dl_init(&e2, Q);
        rs_init(&cw, Q); rs_init(&ew, Q);
        double cr = 0, cp = 0, er = 0, ep = 0, ep2 = 0;
        for (int64_t i = 0; i < n; ++i) {
            double in_i = x[2 * (b * n + i)], in_q = x[2 * (b * n + i) + 1];
            int v, cv, ev, cpv, eqv, e2v;
            double dli = dl_step(&di, in_i, 1, &v);
            double dlq = dl_step(&dq, in_q, 1, &v);
            double qp = dli * in_i + dlq * in_q;                   /* minn_rtl.py:616 */
            double pw = in_i * in_i + in_q * in_q;
            double cs = rs_step(&cw, qp, 1, &cv);
            double es = rs_step(&ew, pw, 1, &ev);
            double cpr = dl_step(&cd, cs, cv, &cpv);
            double eq = dl_step(&e1, es, ev, &eqv);
            double e2q = dl_step(&e2, eq, eqv, &e2v);
            if (cv) cr = cs;
            if (cpv) cp = cpr;
            if (ev) er = es;
            if (eqv) ep = eq;
            if (e2v) ep2 = e2q;
            corr_total[i] += cr + cp;                               /* minn_rtl.py:695-702 */
            energy_total[i] += er + ep + ep2;
            metric_valid[i] &= (uint8_t)e2v;
        }
        free(di.mem); free(dq.mem); free(cd.mem); free(e1.mem); free(e2.mem); free(cw.mem); free(ew.mem);
    }
    double s = 0.0;
    double denom = (double)((int64_t)1 << (smooth_shift > 0 ? smooth_shift : 0));
    double cscale = (double)((int64_t)1 << frac_bits);
    for (int64_t i = 0; i < n; ++i) {
        corr_positive[i] = corr_total[i] > 0.0 ? corr_total[i] : 0.0;
        if (metric_valid[i]) {
            if (smooth_shift == 0) s = corr_positive[i];
            else s += (corr_positive[i] - s) / denom;              /* minn_rtl.py:709-715 */
        }
        smooth_metric[i] = s;
        corr_scaled[i] = s * cscale;
        energy_scaled[i] = threshold_value == 0 ? 0.0 : energy_total[i] * (double)threshold_value;
        above[i] = (uint8_t)(metric_valid[i] && (corr_scaled[i] >= energy_scaled[i]));
    }
}

/* detect_minn_rtl -- minn_rtl.py:750-825.  events: int64[max_ev][4] = peak, detected, seg_lo, seg_hi.
 * segs: int64[max_ev][2] (includes an unclosed tail segment, which yields no event).
 * Returns nev; *nseg_out receives the number of gate segments. */
ORC_API int64_t orc_detect_minn_rtl(const double *corr_positive, const uint8_t *above, const uint8_t *valid,
                                    int64_t n, int64_t hysteresis, int64_t timing_offset,
                                    int64_t *events, int64_t *segs, int64_t max_ev, int64_t *nseg_out)
{
    int open = 0; int64_t gs = -1, pk = 0, low = 0, nev = 0, nseg = 0; double pv = 0.0;
    int64_t hyst_limit = hysteresis > 0 ? hysteresis - 1 : 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        double v = corr_positive[i];
        if (!open) {
            if (above[i]) { open = 1; gs = i; pv = v; pk = i; low = 0; }
        } else {
            if (v >= pv) { pv = v; pk = i; }
            if (above[i]) low = 0;
            else {
                int closing = 0;
                if (hysteresis == 0) closing = 1;
                else if (low == hyst_limit) closing = 1;
                else ++low;
                if (closing) {
                    if (nev < max_ev) {
                        events[4 * nev] = pk; events[4 * nev + 1] = pk + timing_offset;
                        events[4 * nev + 2] = gs; events[4 * nev + 3] = i + 1;
                    }
                    if (nseg < max_ev) { segs[2 * nseg] = gs; segs[2 * nseg + 1] = i + 1; }
                    ++nev; ++nseg; open = 0; gs = -1; pv = 0.0; low = 0;
                }
            }
        }
    }
    if (open && gs >= 0) {
        if (nseg < max_ev) { segs[2 * nseg] = gs; segs[2 * nseg + 1] = n; }
        ++nseg;
    }
    *nseg_out = nseg;
    return nev;
}

/* ---------------------------------------------------------------------------------------
 * Integer model of the SystemVerilog datapath ("rtl_exact" mode; parity UNPINNED against RTL
 * simulation -- no simulator / RTL output vectors exist).  Follows
 * ref/minn_antenna_path.sv:63-194 (products, running sums, history taps, hold registers) and
 * ref/minn_preamble_detector.sv:247-325 (combine, clamp, floor-shift smoother, cross-multiplied
 * threshold).  lag_extra = 0 reproduces minn_rtl.py's lag-Q products; lag_extra = 1 models the
 * registered delay-line output read by the SV (minn_delay_line.sv:72-73) as read in SURVEY.md.
 * iq: int16 (nb, n, 2).  All outputs int64[n] / u8[n].
 * ------------------------------------------------------------------------------------- */
ORC_API void orc_minn_rtl_int(const int16_t *iq, int64_t nb, int64_t n, int64_t Q, int64_t smooth_shift,
                              int64_t threshold_value, int64_t frac_bits, int64_t lag_extra,
                              int64_t *corr_total, int64_t *corr_positive, int64_t *smooth_metric,
                              int64_t *energy_total, uint8_t *metric_valid, uint8_t *above)
{
    int64_t D = Q + lag_extra;
    for (int64_t i = 0; i < n; ++i) { corr_total[i] = 0; energy_total[i] = 0; }
    int64_t *C = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t *E = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t b = 0; b < nb; ++b) {
        const int16_t *s = iq + 2 * b * n;
        int64_t c = 0, e = 0;
        for (int64_t i = 0; i < n; ++i) {
            int64_t xi = s[2 * i], xq = s[2 * i + 1];
            int64_t prod = 0;
            if (i >= D) prod = (int64_t)s[2 * (i - D)] * xi + (int64_t)s[2 * (i - D) + 1] * xq;
            int64_t pw = xi * xi + xq * xq;
            c += prod; e += pw;
            if (i >= Q) {
                int64_t j = i - Q;
                int64_t po = 0;
                if (j >= D) po = (int64_t)s[2 * (j - D)] * s[2 * j] + (int64_t)s[2 * (j - D) + 1] * s[2 * j + 1];
                c -= po;
                e -= (int64_t)s[2 * j] * s[2 * j] + (int64_t)s[2 * j + 1] * s[2 * j + 1];
            }
            C[i] = c; E[i] = e;
        }
        for (int64_t i = 0; i < n; ++i) {
            /* hold registers with their valid gating (minn_antenna_path.sv:168-194) */
            int64_t cr = C[i];                                  /* zero before Q-1 anyway */
            int64_t cp = i >= 2 * Q - 1 ? C[i - Q] : 0;
            int64_t er = i >= Q - 1 ? E[i] : 0;
            int64_t ep = i >= 2 * Q - 1 ? E[i - Q] : 0;
            int64_t ep2 = i >= 3 * Q - 1 ? E[i - 2 * Q] : 0;
            corr_total[i] += cr + cp;
            energy_total[i] += er + ep + ep2;
        }
    }
    free(C); free(E);
    int64_t sm = 0;
    for (int64_t i = 0; i < n; ++i) {
        int v = i >= 3 * Q - 1;
        metric_valid[i] = (uint8_t)v;
        corr_positive[i] = corr_total[i] > 0 ? corr_total[i] : 0;   /* minn_preamble_detector.sv:265-272 */
        if (v) {
            if (smooth_shift == 0) sm = corr_positive[i];
            else sm = sm + ((corr_positive[i] - sm) >> smooth_shift);  /* arithmetic shift, :294-296 */
        }
        smooth_metric[i] = sm;
        /* 52-bit compare fits in int64 for 12-bit inputs, Q<=512 */
        int64_t cs = sm << frac_bits;
        int64_t es = energy_total[i] * threshold_value;
        above[i] = (uint8_t)(v && cs >= es);
    }
}

/* Integer gate / peak FSM, ref/minn_preamble_detector.sv:337-384 (== minn_rtl.py:750-825 on
 * integer metrics). Same outputs as orc_detect_minn_rtl. */
ORC_API int64_t orc_detect_minn_rtl_int(const int64_t *corr_positive, const uint8_t *above, const uint8_t *valid,
                                        int64_t n, int64_t hysteresis, int64_t timing_offset,
                                        int64_t *events, int64_t *segs, int64_t max_ev, int64_t *nseg_out)
{
    int open = 0; int64_t gs = -1, pk = 0, low = 0, nev = 0, nseg = 0, pv = 0;
    int64_t hyst_limit = hysteresis > 1 ? hysteresis - 1 : 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        int64_t v = corr_positive[i];
        if (!open) {
            if (above[i]) { open = 1; gs = i; pv = v; pk = i; low = 0; }
        } else {
            if (v >= pv) { pv = v; pk = i; }
            if (above[i]) low = 0;
            else if (hysteresis == 0 || low == hyst_limit) {
                if (nev < max_ev) {
                    events[4 * nev] = pk; events[4 * nev + 1] = pk + timing_offset;
                    events[4 * nev + 2] = gs; events[4 * nev + 3] = i + 1;
                }
                if (nseg < max_ev) { segs[2 * nseg] = gs; segs[2 * nseg + 1] = i + 1; }
                ++nev; ++nseg; open = 0; gs = -1; pv = 0; low = 0;
            } else ++low;
        }
    }
    if (open && gs >= 0) {
        if (nseg < max_ev) { segs[2 * nseg] = gs; segs[2 * nseg + 1] = n; }
        ++nseg;
    }
    *nseg_out = nseg;
    return nev;
}

/* ---------------------------------------------------------------------------------------
 * Scale path (used only to check BASELINE-size runs and as bench.py's cpu_baseline "port"):
 * closed forms of SURVEY.md Appendix B on float64 prefix sums.  kind: 0 = S&C (sc.py),
 * 1 = S&C both halves, 2 = Minn.  x is complex64 (float pairs) here -- the input the GPU sees.
 * M is float64.  Agreement with the literal functions above ~1e-12 relative (tested).
 * ------------------------------------------------------------------------------------- */
ORC_API int64_t orc_metric_prefix_c64(const float *x, int64_t L, int64_t N, int kind, double *M,
                                      double *P_out, double *R_out)
{
    int64_t out_len = L - N + 1;
    if (out_len <= 0) return 0;
    int64_t D = kind == 2 ? N / 4 : N / 2;
    int64_t np_ = L - D;
    double *sc = (double *)malloc(sizeof(double) * 2 * (size_t)(np_ + 1));
    double *se = (double *)malloc(sizeof(double) * (size_t)(L + 1));
    sc[0] = sc[1] = 0.0; se[0] = 0.0;
    for (int64_t j = 0; j < np_; ++j) {
        double ar = x[2 * j], ai = x[2 * j + 1], br = x[2 * (j + D)], bi = x[2 * (j + D) + 1];
        sc[2 * (j + 1)] = sc[2 * j] + (ar * br + ai * bi);
        sc[2 * (j + 1) + 1] = sc[2 * j + 1] + (ai * br - ar * bi);
    }
    for (int64_t j = 0; j < L; ++j) {
        double ar = x[2 * j], ai = x[2 * j + 1];
        se[j + 1] = se[j] + (ar * ar + ai * ai);
    }
    for (int64_t d = 0; d < out_len; ++d) {
        double pr, pi, r;
        if (kind == 2) {
            int64_t Q = D;
            pr = (sc[2 * (d + Q)] - sc[2 * d]) + (sc[2 * (d + 3 * Q)] - sc[2 * (d + 2 * Q)]);
            pi = (sc[2 * (d + Q) + 1] - sc[2 * d + 1]) + (sc[2 * (d + 3 * Q) + 1] - sc[2 * (d + 2 * Q) + 1]);
            r = se[d + 4 * Q] - se[d + Q];
            double a = pr > 0 ? pr : 0.0, rr = r > 1e-12 ? r : 1e-12;
            M[d] = (a * a) / (rr * rr);
        } else {
            pr = sc[2 * (d + D)] - sc[2 * d];
            pi = sc[2 * (d + D) + 1] - sc[2 * d + 1];
            r = kind == 0 ? se[d + N] - se[d + D] : se[d + N] - se[d];
            double rr = r > 1e-12 ? r : 1e-12;
            M[d] = (pr * pr + pi * pi) / (rr * rr);
        }
        if (P_out) { P_out[2 * d] = pr; P_out[2 * d + 1] = pi; }
        if (R_out) R_out[d] = r;
    }
    free(sc); free(se);
    return out_len;
}

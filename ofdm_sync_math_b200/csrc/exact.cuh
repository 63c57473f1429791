// Exact (float64) re-evaluation of the timing metric around single indices, straight from the samples.
//
// The fused sync pipeline (ofs_sync*) decides on the float32 metric of the stripe kernel.  A decision of the reference's
// detectors (arg-max, `<= 0.95 peak`, `>= thr max`, sc.py:106-114, minn.py:155-205) whose operands differ by less than the
// float32 error of that metric could come out differently in the reference's float64 arithmetic.  Those decisions -- and only
// those -- are re-evaluated here: the CTA computes P, R, M of sc.py:57-77 / combined_sc_min.py:144-163 / minn.py:87-111 for a
// window of up to EXWIN consecutive output indices -- a direct float64 sum over the samples for the first index (all loads of
// the CTA in flight at once: one DRAM round trip), the sliding differences for the others -- then the detector's smoothing
// window on top.  Products of complex64 / int16 samples are exact in float64; the sums carry ~1e-15 relative error, well
// below the ~1e-13 of the reference's own 260 000-step recurrences.  Branches are summed before the non-linear step
// (sc.py:73-74).
#pragma once
#include "common.cuh"

namespace ofs {

// status bits OFS_ST_EXACT / OFS_ST_CHANGED / OFS_ST_UNRESOLVED: include/ofdmsync.h

struct ExactSrc {
    const void *x;      // nullptr: no exact evaluation (the metric rows are the caller's data, there is nothing more exact)
    int dtype, kind, N, nb;
    int64_t L, xfs, xbs;   // samples per frame; frame / branch stride in samples
    double band;           // relative half-width of the float32 uncertainty band
};

constexpr int EXCAP = 64;    // indices re-evaluated per decision and row; more = OFS_ST_UNRESOLVED
constexpr int EXWIN = 64;    // metric values per evaluation window

struct ExactScratch {        // shared memory of a detector CTA
    long long idx[EXCAP];
    double val[EXCAP];
    unsigned char tag[EXCAP];
    int cnt;
    double red[3][32];
    double dlt[3][EXWIN];
    double mwin[EXWIN];
};

// smoothing window of a detector: output i averages M[i - back .. i - back + w - 1]
struct ExactSmooth {
    int w, back;
    bool trailing;           // minn._trailing_average: divisor min(i + 1, w) and max(M, 0); else np.convolve(..., "same"): 1 / w
};

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_f64(v, o);
    return v;
}

// sc.mwin[k] = M(d0 + k) in float64 for k < cnt (<= EXWIN); 0 for indices outside [0, L - N].  Whole CTA (block-uniform
// arguments); returns after a __syncthreads.
template <int DT>
__device__ __noinline__ void exact_metric_cta_t(const ExactSrc &s, int64_t frame, int64_t d0, int cnt, ExactScratch &sc)
{
    using In = typename InT<DT>::type;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = (int)blockDim.x, nwarp = nthr >> 5;
    const bool minn = s.kind == OFS_MINN;
    const int D = minn ? s.N / 4 : s.N / 2;                    // lag
    const int nwin = minn ? 2 : 1;                             // product windows [a, a + D), a = 0 (and 2 D for Minn)
    const int e0 = minn ? D : (s.kind == OFS_SC_BOTH ? 0 : D), e1 = minn ? 4 * D : 2 * D;   // energy window [e0, e1)
    const int64_t out_len = s.L - s.N + 1;
    const int64_t lo = d0 < 0 ? 0 : d0, hi = d0 + cnt < out_len ? d0 + cnt : out_len;
    if (tid < EXWIN) sc.mwin[tid] = 0.0;
    if (hi <= lo) { __syncthreads(); return; }
    double pr = 0.0, pi = 0.0, rr = 0.0;       // direct sums at d = lo, spread over the CTA
    double dpr = 0.0, dpi = 0.0, drr = 0.0;    // thread k in [1, hi - lo): value(lo + k) - value(lo + k - 1)
    const int nd = (int)(hi - lo);
    for (int b = 0; b < s.nb; ++b) {
        const In *xb = reinterpret_cast<const In *>(s.x) + (size_t)frame * s.xfs + (size_t)b * s.xbs;
        for (int w = 0; w < nwin; ++w) {
            const In *xa = xb + lo + 2 * w * D;
            // the two samples of a lag product also feed the energy window where they lie inside it
            const bool ea = 2 * w * D >= e0, ec = true;
#pragma unroll 4
            for (int m = tid; m < D; m += nthr) {
                const In a = xa[m], c = xa[m + D];
                pr += (double)a.x * c.x + (double)a.y * c.y;
                pi += (double)a.y * c.x - (double)a.x * c.y;
                if (ea) rr += (double)a.x * a.x + (double)a.y * a.y;
                if (ec) rr += (double)c.x * c.x + (double)c.y * c.y;
            }
        }
        if (tid >= 1 && tid < nd) {
            const int64_t j = lo + tid - 1;                    // the sample that leaves every window
            for (int w = 0; w < nwin; ++w) {
                const In *xa = xb + j + 2 * w * D;
                const In a = xa[0], c = xa[D], e = xa[2 * D];
                dpr += ((double)c.x * e.x + (double)c.y * e.y) - ((double)a.x * c.x + (double)a.y * c.y);
                dpi += ((double)c.y * e.x - (double)c.x * e.y) - ((double)a.y * c.x - (double)a.x * c.y);
            }
            const In vo = xb[j + e0], vi = xb[j + e1];
            drr += ((double)vi.x * vi.x + (double)vi.y * vi.y) - ((double)vo.x * vo.x + (double)vo.y * vo.y);
        }
    }
    pr = warp_sum_f64(pr); pi = warp_sum_f64(pi); rr = warp_sum_f64(rr);
    if (lane == 0) { sc.red[0][warp] = pr; sc.red[1][warp] = pi; sc.red[2][warp] = rr; }
    if (tid < EXWIN) { sc.dlt[0][tid] = dpr; sc.dlt[1][tid] = dpi; sc.dlt[2][tid] = drr; }
    __syncthreads();
    if (tid < nd) {
        double Pr = 0.0, Pi = 0.0, R = 0.0;
        for (int w = 0; w < nwarp; ++w) { Pr += sc.red[0][w]; Pi += sc.red[1][w]; R += sc.red[2][w]; }
        for (int k = 1; k <= tid; ++k) { Pr += sc.dlt[0][k]; Pi += sc.dlt[1][k]; R += sc.dlt[2][k]; }
        if (R < 1e-12) R = 1e-12;                               // eps clamp of sc.py:75-76 / minn.py:109-110
        double M;
        if (minn) { const double pp = Pr > 0.0 ? Pr : 0.0; M = (pp * pp) / (R * R); }
        else M = (Pr * Pr + Pi * Pi) / (R * R);
        sc.mwin[(int)(lo - d0) + tid] = M;
    }
    __syncthreads();
}

__device__ __forceinline__ void exact_metric_cta(const ExactSrc &s, int64_t frame, int64_t d0, int cnt, ExactScratch &sc)
{
    if (s.dtype == OFS_C64) exact_metric_cta_t<OFS_C64>(s, frame, d0, cnt, sc);
    else if (s.dtype == OFS_IQ16) exact_metric_cta_t<OFS_IQ16>(s, frame, d0, cnt, sc);
    else exact_metric_cta_t<OFS_C128>(s, frame, d0, cnt, sc);
}

// sc.val[k] = smoothed float64 metric at sc.idx[k] for k < cnt (<= EXCAP), any order: the list is covered by windows of EXWIN
// metric values (neighbouring indices -- the usual case -- share one evaluation).  Whole CTA, block-uniform control flow.
__device__ __forceinline__ void exact_eval_list(const ExactSrc &s, int64_t frame, ExactScratch &sc, int cnt, const ExactSmooth sm)
{
    const int tid = threadIdx.x;
    const int span = EXWIN - sm.w + 1;                         // smoothed outputs per window
    unsigned long long pending = cnt >= 64 ? ~0ull : ((1ull << cnt) - 1ull);
    __syncthreads();
    while (pending) {
        long long a = LLONG_MAX;
        for (int k = 0; k < cnt; ++k)
            if ((pending >> k) & 1ull) { const long long i = sc.idx[k]; a = i < a ? i : a; }
        exact_metric_cta(s, frame, a - sm.back, EXWIN, sc);
        unsigned long long done = 0ull;
        for (int k = 0; k < cnt; ++k)
            if (((pending >> k) & 1ull) && sc.idx[k] < a + span) done |= 1ull << k;
        if (tid < cnt && ((done >> tid) & 1ull)) {
            const long long i = sc.idx[tid];
            const int o = (int)(i - a);
            double sum = 0.0;
            if (sm.trailing) {
                for (int k = 0; k < sm.w; ++k) sum += sc.mwin[o + k];
                sum /= (double)(i >= sm.w - 1 ? sm.w : i + 1);
            } else {
                const double h = 1.0 / (double)sm.w;
                for (int k = 0; k < sm.w; ++k) sum += sc.mwin[o + k] * h;
            }
            sc.val[tid] = sum;
        }
        pending &= ~done;
        __syncthreads();
    }
}

}  // namespace ofs

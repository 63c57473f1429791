// Exact (float64) re-evaluation of the timing metric at single indices, straight from the samples.
//
// The fused sync pipeline (ofs_sync*) decides on the float32 metric of the stripe kernel.  A decision of the reference's
// detectors (arg-max, `<= 0.95 peak`, `>= thr max`, sc.py:106-114, minn.py:155-205) whose operands differ by less than the
// float32 error of that metric could come out differently in the reference's float64 arithmetic.  Those decisions -- and only
// those -- are re-evaluated here: one warp computes P, R, M of sc.py:57-77 / combined_sc_min.py:144-163 / minn.py:87-111 for
// up to 32 consecutive output indices as direct float64 sums over the samples (products of complex64 / int16 samples are
// exact in float64; the sums carry ~1e-15 relative error, well below the ~1e-13 of the reference's own 260 000-step
// recurrences), branches summed before the non-linear step (sc.py:73-74), then the detector's smoothing window on top.
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

struct ExactScratch {        // shared memory of a detector CTA
    long long idx[EXCAP];
    double val[EXCAP];
    int cnt;
};

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_f64(v, o);
    return v;
}

// M(d0 + lane) in float64 for lane < cnt (<= 32); 0 for indices outside [0, L - N].  Whole warp, converged.
template <int DT>
__device__ __noinline__ double exact_metric_warp_t(const ExactSrc &s, int64_t frame, int64_t d0, int cnt)
{
    using In = typename InT<DT>::type;
    const int lane = threadIdx.x & 31;
    const bool minn = s.kind == OFS_MINN;
    const int D = minn ? s.N / 4 : s.N / 2;                    // lag
    const int nwin = minn ? 2 : 1;                             // product windows [a, a + D), a = 0 (and 2 D for Minn)
    const int e0 = minn ? D : (s.kind == OFS_SC_BOTH ? 0 : D), e1 = minn ? 4 * D : 2 * D;   // energy window [e0, e1)
    const int64_t out_len = s.L - s.N + 1;
    const int64_t lo = d0 < 0 ? 0 : d0, hi = d0 + cnt < out_len ? d0 + cnt : out_len;
    if (hi <= lo) return 0.0;
    double pr = 0.0, pi = 0.0, rr = 0.0;       // direct sums at d = lo, spread over the lanes
    double dpr = 0.0, dpi = 0.0, drr = 0.0;    // lane k >= 1: value(lo + k) - value(lo + k - 1)
    for (int b = 0; b < s.nb; ++b) {
        const In *xb = reinterpret_cast<const In *>(s.x) + (size_t)frame * s.xfs + (size_t)b * s.xbs;
        for (int w = 0; w < nwin; ++w) {
            const In *xa = xb + lo + 2 * w * D;
#pragma unroll 4
            for (int m = lane; m < D; m += 32) {
                const In a = xa[m], c = xa[m + D];
                pr += (double)a.x * c.x + (double)a.y * c.y;
                pi += (double)a.y * c.x - (double)a.x * c.y;
            }
        }
#pragma unroll 4
        for (int m = e0 + lane; m < e1; m += 32) {
            const In v = xb[lo + m];
            rr += (double)v.x * v.x + (double)v.y * v.y;
        }
        if (lane >= 1 && lo + lane < hi) {
            const int64_t j = lo + lane - 1;                    // the sample that leaves every window
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
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double a = shfl_up_f64(dpr, o), c = shfl_up_f64(dpi, o), e = shfl_up_f64(drr, o);
        if (lane >= o) { dpr += a; dpi += c; drr += e; }
    }
    const double Pr = pr + dpr, Pi = pi + dpi;
    double R = rr + drr;
    if (R < 1e-12) R = 1e-12;                                   // eps clamp of sc.py:75-76 / minn.py:109-110
    double M;
    if (minn) { const double pp = Pr > 0.0 ? Pr : 0.0; M = (pp * pp) / (R * R); }
    else M = (Pr * Pr + Pi * Pi) / (R * R);
    // value of d = lo + lane sits in this lane; the caller wants d = d0 + lane
    const int shift = (int)(lo - d0);
    const double out = shfl_f64(M, (lane - shift) & 31);
    const int64_t d = d0 + lane;
    return (lane < cnt && d >= lo && d < hi) ? out : 0.0;
}

__device__ __forceinline__ double exact_metric_warp(const ExactSrc &s, int64_t frame, int64_t d0, int cnt)
{
    if (s.dtype == OFS_C64) return exact_metric_warp_t<OFS_C64>(s, frame, d0, cnt);
    if (s.dtype == OFS_IQ16) return exact_metric_warp_t<OFS_IQ16>(s, frame, d0, cnt);
    return exact_metric_warp_t<OFS_C128>(s, frame, d0, cnt);
}

// np.convolve(M, ones(w)/w, "same")[i] (sc.py:100) for a row longer than w: mean of M[i + off - w + 1 .. i + off], zeros outside
__device__ __forceinline__ double exact_smooth_same(const ExactSrc &s, int64_t frame, int64_t i, int w, int off)
{
    const int lane = threadIdx.x & 31;
    const double m = exact_metric_warp(s, frame, i + off - w + 1, w);
    return warp_sum_f64(lane < w ? m * (1.0 / (double)w) : 0.0);
}
// minn._trailing_average(max(M, 0), w)[i] (minn.py:115-128): sum of the last w values / min(i + 1, w)
__device__ __forceinline__ double exact_trailing(const ExactSrc &s, int64_t frame, int64_t i, int w)
{
    const int lane = threadIdx.x & 31;
    if (w <= 1) return shfl_f64(exact_metric_warp(s, frame, i, 1), 0);
    const double m = exact_metric_warp(s, frame, i - w + 1, w);
    const double sum = warp_sum_f64(lane < w ? m : 0.0);
    return sum / (double)(i >= w - 1 ? w : i + 1);
}

// Evaluate fn(idx[k]) for every listed index, one warp per index (fn is a warp-collective returning the value in all lanes).
template <typename F>
__device__ __forceinline__ void exact_eval_list(ExactScratch &sc, int cnt, const F &fn)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
    for (int k = warp; k < cnt; k += nw) {
        const double v = fn(sc.idx[k]);
        if (lane == 0) sc.val[k] = v;
    }
    __syncthreads();
}

}  // namespace ofs

// Detector kernels (K2): one CTA per metric row.  Parallel restatements of the reference's
// sequential Python detectors; every tie / edge rule of SURVEY.md 8a is kept:
//   sc.find_plateau_end_from_metric           sc.py:81-146
//   minn._trailing_average / find_minn_peak   minn.py:115-205
//   combined_sc_min gate + gated peak         combined_sc_min.py:183-259, 337-351
//   sync_aa gate FSM                          sync_aa.py:495-568
//   zc_v2 threshold + gate FSM                zc_v2.py:288-336, 360-450
//   minn_rtl gate FSM                         minn_rtl.py:750-825 (ref/minn_preamble_detector.sv:337-384)
// All comparisons are done in float64 on the stored metric values.
//
// The gate FSMs are evaluated as interval logic on a shared-memory bitmask of the "above
// threshold" flags: a gate is a maximal cluster of above-samples whose gaps are shorter than the
// hysteresis; it closes at (last above) + H; the peak is a block-parallel arg-max over
// [open, close] with the reference's tie rule (first max for `>`, last max for `>=`).
#include "common.cuh"
#include "exact.cuh"

namespace ofs {

constexpr int DNT = 256;                   // threads per row-CTA
constexpr int MNT = 768;                   // minn_peak_kernel on long float rows: the 128 KB gate mask allows one CTA per SM, so the CTA is wider
constexpr int64_t MASK_MAX_N = 1600000;     // row length limit of the bitmask kernels (200 KB of smem)

int launch_metric_array(const void *x, int in_dtype, int64_t n_frames, int n_ant, int64_t n, int64_t xfs, int64_t xbs, int L,
                        float *M, void *P, float *R, int64_t out_stride, unsigned *mask, int64_t mask_stride, double thr,
                        cudaStream_t stream);

struct RowView {
    const void *data;
    int f64;
    int64_t n, stride;
    __device__ __forceinline__ double at(int64_t row, int64_t i) const
    {
        return f64 ? reinterpret_cast<const double *>(data)[row * stride + i]
                   : (double)reinterpret_cast<const float *>(data)[row * stride + i];
    }
};

// ---- block reductions ----------------------------------------------------------------------
struct ArgVal {
    double v;
    long long i;
};
template <bool LAST>
__device__ __forceinline__ ArgVal better(ArgVal a, ArgVal b)
{
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.v > a.v) return b;
    if (b.v < a.v) return a;
    if (LAST) return (b.i > a.i) ? b : a;
    return (b.i < a.i) ? b : a;
}
template <bool LAST>
__device__ ArgVal block_argmax(ArgVal x, ArgVal *sh /* DNT/32 */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgVal y;
        y.v = shfl_xor_f64(x.v, o);
        y.i = shfl_xor_i64(x.i, o);
        x = better<LAST>(x, y);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = x;
    __syncthreads();
    ArgVal r = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = better<LAST>(r, sh[w]);
    return r;
}
__device__ long long block_min_i64(long long x, long long *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long y = shfl_xor_i64(x, o);
        x = y < x ? y : x;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = x;
    __syncthreads();
    long long r = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = sh[w] < r ? sh[w] : r;
    return r;
}

// ---- arg-max of a windowed function of the metric, pruned by the stripe kernel's chunk maxima ---------
// chunk c of a row covers the outputs d in [256c - toff, 256c - toff + 256) (256 causal sample times);
// cm[c] = max(0, max of M over the chunk).  A window function F(i) whose window spans the chunks
// [c(i) - rad_lo, c(i) + rad_hi] and is an average of (zero-padded, non-negative-clamped) metric values obeys
// F(i) <= max(cm[c(i)-rad_lo .. c(i)+rad_hi]).  Chunks whose bound is below a value already attained are
// skipped: the result (first maximum) is exactly the one of the full scan.
struct Prune {
    const float *cm;      // nullptr: no pruning.  Points to SHARED memory after stage().
    int64_t ncm;          // entries in this row of cm
    int toff, rad_lo, rad_hi;
    // copy the row of chunk maxima into shared memory (coalesced) so that bound() is not a chain of global loads
    __device__ void stage(float *smem_cm, int64_t nch)
    {
        if (cm == nullptr) return;
        if (nch < ncm) ncm = nch;
        for (int64_t c = threadIdx.x; c < ncm; c += blockDim.x) smem_cm[c] = __ldg(cm + c);
        cm = smem_cm;
        __syncthreads();
    }
    __device__ __forceinline__ float bound(int64_t c) const
    {
        float b = 0.f;
        int64_t lo = c - rad_lo, hi = c + rad_hi;
        if (lo < 0) lo = 0;
        if (hi > ncm - 1) hi = ncm - 1;
        for (int64_t k = lo; k <= hi; ++k) b = fmaxf(b, cm[k]);
        return b;
    }
};

// Window functions are evaluated for 8 consecutive outputs at a time (eval8): the window is loaded once and slid,
// ~4 loads per output instead of w.  Groups are aligned to the causal time t = d + toff (multiples of 8), so that one
// warp = 32 groups = one 256-sample chunk.  Every pass of a kernel goes through the same grouping, hence sees
// bit-identical values.
// slack / n_near (exact mode): chunks are kept while their bound reaches v* (1 - slack), and *n_near receives how many values
// reach that level -- 1 means the maximum has no contender inside the band and needs no float64 re-evaluation; -1 = unknown
// (no chunk maxima to derive v* from).
template <int K = 1, typename F>
__device__ ArgVal pruned_argmax(const F &fn, int64_t n_out, const Prune &pr, ArgVal *sh_av, double slack = 0.0, int *n_near = nullptr)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nch = (n_out + pr.toff + 255) / 256;
    double v8[8];
    ArgVal vs{0.0, -1};
    bool have = false;
    if (pr.cm != nullptr) {
        // phase 0: the chunk with the largest raw maximum
        ArgVal bc{0.0, -1};
        for (int64_t c = tid; c < nch && c < pr.ncm; c += blockDim.x) {
            const double v = (double)pr.cm[c];
            if (bc.i < 0 || v > bc.v) { bc.v = v; bc.i = c; }
        }
        bc = block_argmax<false>(bc, sh_av);
        // phase 1: exact values over that chunk -> a value v* that the maximum is known to reach
        if (bc.i >= 0 && warp == 0) {
            const int64_t i0 = bc.i * 256 - pr.toff + 8 * lane;
            fn.eval8(i0, v8);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (i0 + k >= 0 && i0 + k < n_out && (vs.i < 0 || v8[k] > vs.v)) { vs.v = v8[k]; vs.i = i0 + k; }
        }
        vs = block_argmax<false>(vs, sh_av);
        have = vs.i >= 0;
    }
    // phase 2: every chunk that could still hold the maximum (all chunks when there is no pruning information).
    // The bound test is done 32 chunks at a time (one per lane); survivors are evaluated by the whole warp.
    ArgVal best{0.0, -1};
    const int nchi = (int)nch;
    const double near_lvl = vs.v * (1.0 - slack);
    int near = 0;
    for (int c0 = warp * 32; c0 < nchi; c0 += (int)blockDim.x) {
        const int cl = c0 + lane;
        const bool pass = cl < nchi && (!have || (double)pr.bound(cl) >= near_lvl);
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {                              // K surviving chunks per round, in ascending order (first maximum wins)
            int64_t i0s[K];
            double vk[K][8];
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < K; ++u)
                if (m) {                         // warp-uniform; slots fill in order, out-of-row lanes evaluate zeros
                    const int c = c0 + __ffs(m) - 1;
                    m &= m - 1;
                    i0s[u] = (int64_t)c * 256 - pr.toff + 8 * lane;
                    cnt = u + 1;
                }
            fn.template eval8_multi<K>(i0s, cnt, vk);
#pragma unroll
            for (int u = 0; u < K; ++u)
                if (u < cnt) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (i0s[u] + k >= 0 && i0s[u] + k < n_out) {
                            if (best.i < 0 || vk[u][k] > best.v) { best.v = vk[u][k]; best.i = i0s[u] + k; }
                            near += vk[u][k] >= near_lvl;
                        }
                }
        }
    }
    if (n_near) {
        int w = near;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        __shared__ int sh_near;
        if (tid == 0) sh_near = 0;
        __syncthreads();
        if (lane == 0 && w) atomicAdd(&sh_near, w);
        __syncthreads();
        *n_near = have ? sh_near : -1;
    }
    return block_argmax<false>(best, sh_av);
}

// first-maximum over an index range [lo, hi) through the same 8-groups
template <typename F>
__device__ ArgVal range_argmax(const F &fn, int64_t lo, int64_t hi, int toff, ArgVal *sh_av)
{
    ArgVal best{0.0, -1};
    double v8[8];
    const int64_t g0 = lo - (((lo + toff) % 8) + 8) % 8;
    for (int64_t i0 = g0 + 8LL * threadIdx.x; i0 < hi; i0 += 8LL * blockDim.x) {
        fn.eval8(i0, v8);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i0 + k >= lo && i0 + k < hi && (best.i < 0 || v8[k] > best.v)) { best.v = v8[k]; best.i = i0 + k; }
    }
    return block_argmax<false>(best, sh_av);
}

// Window of W + 7 floats starting SKIP elements after index a, as aligned 16-byte loads (6 instead of 23 predicated scalar
// loads for W = 16).  Only when the whole span lies inside the row and is 16-byte aligned -- the detectors' 8-groups are
// aligned to the causal time t = d + toff, which is how the stripe kernel lays M out; anything else takes the scalar path.
template <int W, int SKIP>
__device__ __forceinline__ bool load_window_vec(const float *p, int64_t a, int64_t n, float (&y)[W + 7])
{
    constexpr int NV = (SKIP + W + 7 + 3) / 4;
    if (a < 0 || a + 4 * NV > n || (reinterpret_cast<uintptr_t>(p + a) & 15)) return false;
    const float4 *q = reinterpret_cast<const float4 *>(p + a);
    float tmp[4 * NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const float4 f = q[v];
        tmp[4 * v] = f.x; tmp[4 * v + 1] = f.y; tmp[4 * v + 2] = f.z; tmp[4 * v + 3] = f.w;
    }
#pragma unroll
    for (int k = 0; k < W + 7; ++k) y[k] = tmp[SKIP + k];
    return true;
}

// Eight consecutive window sums from W + 7 terms with a short dependency chain: pairwise tree for the first window, then
// o[k] = o[0] + prefix_k(t[W-1+j] - t[j-1]).  (A slid accumulator is a chain of 2 W dependent float64 additions per group; with
// eight warps per SM that latency, not bandwidth, was what the detectors waited on.  Rounding differs from a slid sum by
// ~1e-16 of the window sum; every pass of a kernel uses this one function, so all passes see identical values.)
template <int W>
__device__ __forceinline__ void window_sums8(double (&t)[W + 7], double (&o)[8])
{
    static_assert(W == 8 || W == 16, "power-of-two windows only");
    double d[8];
    d[0] = 0.0;
#pragma unroll
    for (int k = 1; k < 8; ++k) d[k] = t[W - 1 + k] - t[k - 1];
#pragma unroll
    for (int st = 1; st < W; st <<= 1) {
#pragma unroll
        for (int q = 0; q < W; q += 2 * st) t[q] += t[q + st];
    }
#pragma unroll
    for (int st = 1; st < 8; st <<= 1) {
#pragma unroll
        for (int k = 7; k >= st; --k) d[k] += d[k - st];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = t[0] + d[k];
}

// ---- S&C plateau end -------------------------------------------------------------------------------
struct SmoothSame {   // np.convolve(M, ones(w)/w, "same")[i], sc.py:100
    const float *pf;
    const double *pd;
    int64_t n, ms;
    int w, off, toff;
    double h;
    __device__ __forceinline__ double X(int64_t j) const      // M[j] * (1/w), zero outside the row
    {
        const bool ok = j >= 0 && j < n;
        const int64_t jc = ok ? j : 0;
        const double v = pd ? pd[jc] : (double)pf[jc];
        return ok ? v * h : 0.0;
    }
    // compile-time window: the W+7 values are fetched by independent (predicated) loads first, then slid.  The window is held
    // in the row's own type (float rows: half the registers of a double window; the conversion at use is exact).
    template <int W, typename T>
    __device__ __forceinline__ void eval8_raw(const T *p, int64_t i0, double (&o)[8]) const
    {
        T y[W + 7];
        const int64_t j0 = i0 + off - W + 1;
        bool done = false;
        if constexpr (sizeof(T) == 4) done = load_window_vec<W, 0>(p, j0, n, y);
        if (!done) {
#pragma unroll
            for (int q = 0; q < W + 7; ++q) {
                const int64_t j = j0 + q;
                const bool ok = j >= 0 && j < n;
                const T v = p[ok ? j : 0];
                y[q] = ok ? v : T(0);
            }
        }
        double t[W + 7];
#pragma unroll
        for (int q = 0; q < W + 7; ++q) t[q] = (double)y[q] * h;
        window_sums8<W>(t, o);
    }
    template <int W>
    __device__ __forceinline__ void eval8_t(int64_t i0, double (&o)[8]) const
    {
        if (pd) eval8_raw<W, double>(pd, i0, o);
        else eval8_raw<W, float>(pf, i0, o);
    }
    __device__ __forceinline__ void eval8(int64_t i0, double (&o)[8]) const
    {
        if (w == 16) { eval8_t<16>(i0, o); return; }
        if (w == 8) { eval8_t<8>(i0, o); return; }
        // window of output i: j in [i + off - w + 1, i + off]
        double s = 0.0;
        for (int64_t j = i0 + off - w + 1; j <= i0 + off; ++j) s += X(j);
        o[0] = s;
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            s += X(i0 + k + off);
            s -= X(i0 + k + off - w);
            o[k] = s;
        }
    }
    template <int K>
    __device__ __forceinline__ void eval8_multi(const int64_t (&i0)[K], int cnt, double (&o)[K][8]) const
    {
#pragma unroll
        for (int u = 0; u < K; ++u)
            if (u < cnt) eval8(i0[u], o[u]);
    }
    __device__ __forceinline__ double operator()(int64_t i) const
    {
        const int64_t g0 = i - (((i + toff) % 8) + 8) % 8;
        double o[8];
        eval8(g0, o);
        return o[i - g0];
    }
};

// append the indices i0 + k (k = set bits of hits) to the list; kept out of the unrolled 8-loops
__device__ __forceinline__ void push_hits(ExactScratch &sc, int64_t i0, unsigned hits)
{
#pragma unroll 1
    while (hits) {
        const int k = __ffs(hits) - 1;
        hits &= hits - 1;
        const int slot = atomicAdd(&sc.cnt, 1);
        if (slot < EXCAP) sc.idx[slot] = i0 + k;
    }
}

// Collect every index i < n_out with fn(i) >= lvl into sc.idx (ascending order is NOT guaranteed), through the same chunk
// pruning and 8-groups as pruned_argmax.  Returns the number found (may exceed EXCAP: then the list is incomplete).
template <typename F>
__device__ int collect_at_least(const F &fn, int64_t n_out, const Prune &pr, double lvl, ExactScratch &sc)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchi = (int)((n_out + pr.toff + 255) / 256);
    if (tid == 0) sc.cnt = 0;
    __syncthreads();
    double v8[8];
    for (int c0 = warp * 32; c0 < nchi; c0 += (int)blockDim.x) {
        const int cl = c0 + lane;
        const bool pass = cl < nchi && (pr.cm == nullptr || (double)pr.bound(cl) >= lvl);
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
            const int c = c0 + __ffs(m) - 1;
            m &= m - 1;
            const int64_t i0 = (int64_t)c * 256 - pr.toff + 8 * lane;
            fn.eval8(i0, v8);
            unsigned hits = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (i0 + k >= 0 && i0 + k < n_out && v8[k] >= lvl) hits |= 1u << k;
            push_hits(sc, i0, hits);
        }
    }
    __syncthreads();
    return sc.cnt;
}
// same over an index range [lo, hi)
template <typename F>
__device__ int collect_range_at_least(const F &fn, int64_t lo, int64_t hi, int toff, double lvl, ExactScratch &sc)
{
    if (threadIdx.x == 0) sc.cnt = 0;
    __syncthreads();
    double v8[8];
    const int64_t g0 = lo - (((lo + toff) % 8) + 8) % 8;
    for (int64_t i0 = g0 + 8LL * threadIdx.x; i0 < hi; i0 += 8LL * blockDim.x) {
        fn.eval8(i0, v8);
        unsigned hits = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i0 + k >= lo && i0 + k < hi && v8[k] >= lvl) hits |= 1u << k;
        push_hits(sc, i0, hits);
    }
    __syncthreads();
    return sc.cnt;
}
// first maximum of the re-evaluated list (cnt <= EXCAP entries, evaluated by exact_eval_list); every thread gets the result
__device__ __forceinline__ ArgVal exact_list_argmax(const ExactScratch &sc, int cnt)
{
    ArgVal b{0.0, -1};
    for (int k = 0; k < cnt; ++k) {
        const double v = sc.val[k];
        const long long i = sc.idx[k];
        if (b.i < 0 || v > b.v || (v == b.v && i < b.i)) { b.v = v; b.i = i; }
    }
    return b;
}

// ex.x != nullptr (fused sync pipeline, float32 rows from the stripe kernel): decisions whose operands lie within ex.band of
// each other are re-evaluated in float64 from the samples (exact.cuh), so the index equals the one the reference finds on its
// float64 metric; status[row] reports what happened (OFS_ST_*).
#ifndef OFS_DET_EXACT_CTAS
#define OFS_DET_EXACT_CTAS 3
#endif
// EXACT is a compile-time switch: the plain detector API (metric rows are the caller's data) carries none of the exact-mode code.
template <bool F64, bool EXACT>
__global__ void __launch_bounds__(DNT, F64 ? 2 : (EXACT ? OFS_DET_EXACT_CTAS : 3)) plateau_kernel(RowView r, int cp_len, int lookahead, int smooth_win,
                                                      int64_t *out, const float *cm, int64_t cm_stride, int toff,
                                                      const __grid_constant__ ExactSrc ex, int32_t *status)
{
    __shared__ ArgVal sh_av[DNT / 32];
    __shared__ long long sh_i[DNT / 32];
    __shared__ ExactScratch sc;
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x;
    int st = 0;
    auto finish = [&](long long v) { if (tid == 0) { out[row] = v; if (status) status[row] = st; } };
    if (r.n == 0) { finish(0); return; }
    const int Lk = lookahead < 0 ? cp_len / 4 : (lookahead > 1 ? lookahead : 1);
    const int w = smooth_win > 1 ? smooth_win : 1;
    SmoothSame Ms;
    Ms.pf = F64 ? nullptr : reinterpret_cast<const float *>(r.data) + row * r.stride;
    Ms.pd = F64 ? reinterpret_cast<const double *>(r.data) + row * r.stride : nullptr;
    Ms.n = r.n; Ms.w = w; Ms.h = 1.0 / (double)w; Ms.toff = toff;
    Ms.ms = r.n > w ? r.n : w;
    Ms.off = (int)(((r.n > w ? (int64_t)w : r.n) - 1) / 2);
    const int64_t ms = Ms.ms;

    // pass A: center = first argmax of the smoothed metric (sc.py:106), pruned by the chunk maxima
    extern __shared__ float cm_s[];
    Prune pr{(cm && r.n > w) ? cm + row * cm_stride : nullptr, cm_stride, toff, (w - 1 - Ms.off + 255) / 256, (Ms.off + 255) / 256};
    pr.stage(cm_s, (ms + toff + 255) / 256);
    int n_near = -1;
    ArgVal best = EXACT ? pruned_argmax(Ms, ms, pr, sh_av, ex.band, &n_near) : pruned_argmax(Ms, ms, pr, sh_av);
    int64_t center = best.i;
    double peak = best.v;
    const bool exact_on = EXACT && !F64 && ex.x != nullptr && r.n > w && w <= 32 && peak > 0.0;
    const ExactSmooth esm{w, w - 1 - Ms.off, false};
    bool peak_exact = false;
    if (exact_on && n_near != 1) {
        // every index whose float32 value is within the band of the float32 maximum can be the float64 maximum
        // (n_near == 1: the scan above saw no second value near the maximum -- the usual case, nothing to do)
        const int cnt = collect_at_least(Ms, ms, pr, peak * (1.0 - ex.band), sc);
        if (cnt > EXCAP) st |= OFS_ST_UNRESOLVED;
        else if (cnt > 1) {
            exact_eval_list(ex, row, sc, cnt, esm);
            const ArgVal b = exact_list_argmax(sc, cnt);
            st |= OFS_ST_EXACT | (b.i != center ? OFS_ST_CHANGED : 0);
            center = b.i; peak = b.v; peak_exact = true;
            __syncthreads();
        }
    }

    // pass B: first index in [center, center+cp) at or below 95 % of the maximum (sc.py:107-114)
    const int64_t post_hi = ms < center + cp_len ? ms : center + cp_len;
    if (post_hi > center + 1) {
        const double thr = 0.95 * peak;
        // float32 value v, float64 value V: v <= thr_def  =>  V <= 0.95 peak64;  v > thr_unc  =>  V > 0.95 peak64
        const double bw = exact_on ? (peak_exact ? ex.band : 2.0 * ex.band) : 0.0;
        const double thr_def = thr * (1.0 - bw), thr_unc = thr * (1.0 + bw);
        long long first = LLONG_MAX, first32 = LLONG_MAX;
        if (exact_on) { if (tid == 0) sc.cnt = 0; __syncthreads(); }
        {
            double v8[8];
            const int64_t g0 = center - (((center + Ms.toff) % 8) + 8) % 8;
            for (int64_t i0 = g0 + 8LL * tid; i0 < post_hi && first == LLONG_MAX; i0 += 8LL * DNT) {
                Ms.eval8(i0, v8);
                unsigned m32 = 0u, mdef = 0u, munc = 0u;        // bit k: value k at or below thr / thr_def / thr_unc
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const unsigned in = (i0 + k >= center && i0 + k < post_hi) ? 1u << k : 0u;
                    if (v8[k] <= thr) m32 |= in;
                    if (v8[k] <= thr_def) mdef |= in;
                    if (v8[k] <= thr_unc) munc |= in;
                }
                if (m32 && first32 == LLONG_MAX) first32 = i0 + __ffs(m32) - 1;
                if (mdef) { first = i0 + __ffs(mdef) - 1; munc &= (1u << (__ffs(mdef) - 1)) - 1u; }
                munc &= ~mdef;                                  // inside the band, before the first certain one: float64 decides
                if (exact_on) push_hits(sc, i0, munc);
            }
        }
        first = block_min_i64(first, sh_i);
        if (exact_on) {
            first32 = block_min_i64(first32, sh_i);
            const int cnt = sc.cnt;                            // (block_min_i64 has synchronised the block)
            // only in-band indices BEFORE the first certain one matter
            bool any = false;
            for (int k = 0; k < cnt && k < EXCAP; ++k) any |= sc.idx[k] < first;
            if (cnt >= EXCAP) { st |= OFS_ST_UNRESOLVED; first = first32; }
            else if (any) {
                __syncthreads();
                if (tid == 0) {                                 // drop the ones behind `first`, append the centre (for peak64)
                    int q = 0;
                    for (int k = 0; k < cnt; ++k) if (sc.idx[k] < first) { sc.idx[q] = sc.idx[k]; sc.tag[q++] = 0; }
                    if (!peak_exact) { sc.idx[q] = center; sc.tag[q++] = 1; }    // tagged: not a candidate
                    sc.cnt = q;
                }
                __syncthreads();
                const int q = sc.cnt;
                exact_eval_list(ex, row, sc, q, esm);
                double pk64 = peak;
                for (int k = 0; k < q; ++k) if (sc.tag[k]) pk64 = sc.val[k];
                const double thr64 = 0.95 * pk64;
                long long fx = first;
                for (int k = 0; k < q; ++k) if (!sc.tag[k] && sc.val[k] <= thr64 && sc.idx[k] < fx) fx = sc.idx[k];
                st |= OFS_ST_EXACT | (fx != first32 ? OFS_ST_CHANGED : 0);
                first = fx;
                __syncthreads();
            }
        }
        if (first != LLONG_MAX) { finish(first); return; }
    }
    // the remaining paths are the reference's fallbacks (no sample within cp_len of the maximum drops below 95 %): rare, and
    // their decisions are not re-evaluated here -- a row that gets this far with the exact mode on is reported as unresolved
    // unless nothing it compares lies inside the band.
    const double bwc = exact_on ? 2.0 * ex.band : 0.0;

    // pass C: right edge of the earliest run >= 60 % of the peak with length >= max(8, cp/2) (sc.py:117-133)
    if (peak > 0.0) {
        const double thr = 0.6 * peak;
        const int64_t min_run = cp_len / 2 > 8 ? cp_len / 2 : 8;
        __shared__ long long sh_res;
        __shared__ int sh_unc;
        if (tid < 32) {                      // warp 0 walks the row, 32 flags per step
            long long res = -1, run_start = -1;
            unsigned unc = 0u;
            for (int64_t base = 0; base < ms && res < 0; base += 32) {
                const int64_t i = base + tid;
                const double v = i < ms ? Ms(i) : 0.0;
                const bool f = i < ms && v >= thr;
                unc |= __ballot_sync(0xffffffffu, i < ms && fabs(v - thr) <= bwc * thr);
                const unsigned m = __ballot_sync(0xffffffffu, f);
                const int nbits = (int)(ms - base < 32 ? ms - base : 32);
                if (m == 0u) {
                    if (run_start >= 0 && base - run_start >= min_run) res = base - 1;
                    run_start = -1;
                } else if (nbits == 32 && m == 0xffffffffu) {
                    if (run_start < 0) run_start = base;
                } else {
                    for (int b = 0; b < nbits && res < 0; ++b) {
                        if ((m >> b) & 1u) { if (run_start < 0) run_start = base + b; }
                        else {
                            if (run_start >= 0 && base + b - run_start >= min_run) res = base + b - 1;
                            run_start = -1;
                        }
                    }
                }
            }
            if (res < 0 && run_start >= 0 && ms - run_start >= min_run) res = ms - 1;
            if (tid == 0) { sh_res = res; sh_unc = (exact_on && unc != 0u) ? 1 : 0; }
        }
        __syncthreads();
        if (sh_unc) st |= OFS_ST_UNRESOLVED;
        if (sh_res >= 0) { finish(sh_res); return; }
    }

    // pass D: largest drop over the lookahead around the strongest plateau (sc.py:136-146)
    {
        if (exact_on) st |= OFS_ST_UNRESOLVED;     // an arg-max of differences: not re-evaluated (never seen on real frames)
        const int64_t lo = center - cp_len > 0 ? center - cp_len : 0;
        const int64_t hi = ms - Lk - 1 < center + cp_len ? ms - Lk - 1 : center + cp_len;
        int64_t hi_eff = hi;                       // python slice semantics for a negative stop
        if (hi_eff < 0) hi_eff += ms;
        if (hi_eff < 0) hi_eff = 0;
        if (hi_eff > ms) hi_eff = ms;
        int64_t hi2 = hi + Lk, lo2 = lo + Lk;
        if (hi2 < 0) hi2 += ms;
        if (hi2 < 0) hi2 = 0;
        if (hi2 > ms) hi2 = ms;
        if (lo2 > ms) lo2 = ms;
        const int64_t wl = hi_eff - lo, al = hi2 - lo2;
        if (wl <= 0 || al <= 0 || wl != al) { finish(center); return; }
        ArgVal bd{0.0, -1};
        for (int64_t i = tid; i < wl; i += DNT) {
            const double dv = Ms(lo + i) - Ms(lo2 + i);
            if (bd.i < 0 || dv > bd.v) { bd.v = dv; bd.i = i; }
        }
        bd = block_argmax<false>(bd, sh_av);
        finish(lo + bd.i + Lk / 2);
    }
}

// ---- trailing average (minn.py:115-128) as a pure function of the index ---------------------------
struct Trailing {   // minn._trailing_average(max(M,0), w)[i], minn.py:115-128
    const float *pf;
    const double *pd;
    int64_t n;
    int w, toff;
    __device__ __forceinline__ double X(int64_t j) const      // max(M[j], 0), zero outside the row
    {
        const bool ok = j >= 0 && j < n;
        const int64_t jc = ok ? j : 0;
        const double v = pd ? pd[jc] : (double)pf[jc];
        return (ok && v > 0.0) ? v : 0.0;
    }
    __device__ __forceinline__ double denom(int64_t i) const { return (double)(i >= w - 1 ? w : (i >= 0 ? i + 1 : 1)); }
    template <int W, typename T>
    __device__ __forceinline__ void load_win(const T *p, int64_t i0, T (&y)[W + 7]) const
    {
        const int64_t j0 = i0 - W + 1;               // max(M, 0) in the row's own type (zero outside the row)
        if constexpr (sizeof(T) == 4) {
            if (load_window_vec<W, 1>(p, j0 - 1, n, y)) {
#pragma unroll
                for (int q = 0; q < W + 7; ++q) y[q] = y[q] > 0.f ? y[q] : 0.f;
                return;
            }
        }
#pragma unroll
        for (int q = 0; q < W + 7; ++q) {
            const int64_t j = j0 + q;
            const bool ok = j >= 0 && j < n;
            const T v = p[ok ? j : 0];
            y[q] = (ok && v > T(0)) ? v : T(0);
        }
    }
    template <int W, typename T>
    __device__ __forceinline__ void slide_win(const T (&y)[W + 7], int64_t i0, double (&o)[8]) const
    {
        double t[W + 7];
#pragma unroll
        for (int q = 0; q < W + 7; ++q) t[q] = (double)y[q];
        window_sums8<W>(t, o);
        // W is a power of two here: s * (1/W) is exactly s / W; only the warm-up outputs (i < W-1) need a true division
        const bool steady = i0 >= W - 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = steady ? o[k] * (1.0 / W) : o[k] / denom(i0 + k);
    }
    template <int W, typename T>
    __device__ __forceinline__ void eval8_raw(const T *p, int64_t i0, double (&o)[8]) const
    {
        T y[W + 7];
        load_win<W, T>(p, i0, y);
        slide_win<W, T>(y, i0, o);
    }
    // K windows at once: all loads are issued before the first sum, so one warp keeps K chunks in flight (the detectors are
    // bound by the latency of these scattered reads, not by bandwidth).  Same values as K calls of eval8.
    template <int K>
    __device__ __forceinline__ void eval8_multi(const int64_t (&i0)[K], int cnt, double (&o)[K][8]) const
    {
        if (K > 1 && w == 16 && pf) {
            float y[K][23];
#pragma unroll
            for (int u = 0; u < K; ++u)
                if (u < cnt) load_win<16, float>(pf, i0[u], y[u]);
#pragma unroll
            for (int u = 0; u < K; ++u)
                if (u < cnt) slide_win<16, float>(y[u], i0[u], o[u]);
            return;
        }
#pragma unroll
        for (int u = 0; u < K; ++u)
            if (u < cnt) eval8(i0[u], o[u]);
    }
    template <int W>
    __device__ __forceinline__ void eval8_t(int64_t i0, double (&o)[8]) const
    {
        if (pd) eval8_raw<W, double>(pd, i0, o);
        else eval8_raw<W, float>(pf, i0, o);
    }
    __device__ __forceinline__ void eval8(int64_t i0, double (&o)[8]) const
    {
        if (w <= 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = X(i0 + k);
            return;
        }
        if (w == 16) { eval8_t<16>(i0, o); return; }
        if (w == 8) { eval8_t<8>(i0, o); return; }
        double s = 0.0;
        for (int64_t j = i0 - w + 1; j <= i0; ++j) s += X(j);
        o[0] = s / denom(i0);
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            s += X(i0 + k);
            s -= X(i0 + k - w);
            o[k] = s / denom(i0 + k);
        }
    }
    __device__ __forceinline__ double operator()(int64_t i) const
    {
        const int64_t g0 = i - (((i + toff) % 8) + 8) % 8;
        double o[8];
        eval8(g0, o);
        return o[i - g0];
    }
};
// F64 is the row type as a compile-time constant: the window code of the other type is eliminated, and with it its registers
// (a double window is 46 registers, a float one 23: float rows run three CTAs per SM instead of two).
template <bool F64>
__device__ __forceinline__ Trailing make_trailing(const RowView &r, int64_t row, int w, int toff)
{
    Trailing t;
    t.pf = F64 ? nullptr : reinterpret_cast<const float *>(r.data) + row * r.stride;
    t.pd = F64 ? reinterpret_cast<const double *>(r.data) + row * r.stride : nullptr;
    t.n = r.n; t.w = w; t.toff = toff;
    return t;
}

// Summary of the runs of ones in a bit range [a0, a1): pre = end of the run glued to a0 (a0 if none; a1 if the range is all
// ones), suf = start of the run glued to a1 (-1 if none), (bl, bs, be) = longest run touching neither end, earliest on ties.
struct RunSum { long long a0, a1, pre, suf, bl, bs, be; };
// Summary of [A.a0, B.a1) from two adjacent ranges.  Associative; candidates are compared in order of their start with a strict
// '>' so that the earliest of equally long runs wins (minn.py:159-182).
__device__ __forceinline__ RunSum run_join(const RunSum &A, const RunSum &B)
{
    if (A.a0 >= A.a1) return B;
    if (B.a0 >= B.a1) return A;
    const bool a_all = A.pre >= A.a1, b_all = B.pre >= B.a1;
    const long long s = A.suf >= 0 ? A.suf : B.a0, e = B.pre;        // the run across the seam: [s, e)
    RunSum R;
    R.a0 = A.a0; R.a1 = B.a1;
    R.bl = A.bl; R.bs = A.bs; R.be = A.be;
    if (!a_all && !b_all && e - s > R.bl) { R.bl = e - s; R.bs = s; R.be = e; }
    if (B.bl > R.bl) { R.bl = B.bl; R.bs = B.bs; R.be = B.be; }
    R.pre = a_all ? e : A.pre;
    R.suf = b_all ? s : B.suf;
    return R;
}

// Longest run of ones (earliest on ties), executed by the WHOLE CTA: every thread scans its own odd-sized slice of mask words
// (conflict-free, a zero or all-one word costs a handful of instructions) and reports (run glued to the slice start, best run
// strictly inside, run still open at the slice end); warp 0 joins the summaries (run_join: blockDim/32 per lane, then a 5-step tree).
// (History: one warp walking the row was 2/3 of minn_peak_kernel; eight warps walking 32-word groups with
// shuffles still spent 150 k warp-instructions per 1 M-sample row, 29 % of the kernel; this form needs about a tenth.)
__device__ void longest_run_block(const unsigned *mask, int64_t n, long long &bs, long long &be)
{
    __shared__ long long s_pre[MNT], s_suf[MNT], s_bl[MNT], s_bs[MNT], s_be[MNT], s_res[2];   // blockDim.x <= MNT slices
    const int tid = threadIdx.x;
    const int64_t nw = (n + 31) / 32;
    int64_t wpt = (nw + blockDim.x - 1) / blockDim.x;
    wpt |= 1;
    const int64_t w_lo = (int64_t)tid * wpt < nw ? (int64_t)tid * wpt : nw;
    const int64_t w_hi = w_lo + wpt < nw ? w_lo + wpt : nw;
    const long long seg0 = w_lo * 32, seg1 = w_hi * 32 < n ? w_hi * 32 : n;
    long long best_len = 0, rbs = 0, rbe = 0, pre_end = seg0, cur = -1;
    auto close = [&](long long e) {                          // the run [cur, e) ends
        if (cur == seg0) pre_end = e;
        else if (e - cur > best_len) { best_len = e - cur; rbs = cur; rbe = e; }
        cur = -1;
    };
    for (int64_t w = w_lo; w < w_hi; ++w) {
        unsigned m = mask[w];
        if ((w + 1) * 32 > n) m &= (1u << (int)(n - w * 32)) - 1u;
        const long long base = w * 32;
        if (m == 0u) { if (cur >= 0) close(base); continue; }
        if (m == 0xffffffffu) { if (cur < 0) cur = base; continue; }
        int pos = 0;
        while (pos < 32) {
            const unsigned rest = m >> pos;
            if (rest & 1u) {
                if (cur < 0) cur = base + pos;
                pos += __ffs(~rest) - 1;                      // zeros shifted in at the top end the count at the word boundary
                if (pos < 32) close(base + pos);
            } else {
                if (cur >= 0) close(base + pos);
                pos += rest == 0u ? 32 - pos : __ffs(rest) - 1;
            }
        }
    }
    long long suf = -1;
    if (cur >= 0) { suf = cur; if (cur == seg0) pre_end = seg1; }
    s_pre[tid] = pre_end; s_suf[tid] = suf; s_bl[tid] = best_len; s_bs[tid] = rbs; s_be[tid] = rbe;
    __syncthreads();
    if (tid < 32) {                                          // warp 0 stitches: 8 slices per lane in order, then a 5-step tree
        auto slice = [&](int t) {
            RunSum r;
            r.a0 = (long long)t * wpt * 32 < n ? (long long)t * wpt * 32 : n;
            r.a1 = r.a0 + wpt * 32 < n ? r.a0 + wpt * 32 : n;
            r.pre = s_pre[t]; r.suf = s_suf[t]; r.bl = s_bl[t]; r.bs = s_bs[t]; r.be = s_be[t];
            return r;
        };
        const int PER = (int)(blockDim.x >> 5);
        RunSum R = slice(tid * PER);
#pragma unroll 1
        for (int k = 1; k < PER; ++k) R = run_join(R, slice(tid * PER + k));
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            RunSum B;
            B.a0 = __shfl_down_sync(0xffffffffu, R.a0, o); B.a1 = __shfl_down_sync(0xffffffffu, R.a1, o);
            B.pre = __shfl_down_sync(0xffffffffu, R.pre, o); B.suf = __shfl_down_sync(0xffffffffu, R.suf, o);
            B.bl = __shfl_down_sync(0xffffffffu, R.bl, o); B.bs = __shfl_down_sync(0xffffffffu, R.bs, o);
            B.be = __shfl_down_sync(0xffffffffu, R.be, o);
            if ((tid & (2 * o - 1)) == 0) R = run_join(R, B);
        }
        if (tid == 0) {                                      // candidates in order of their start: glued-to-0 run, inside best, tail run
            long long bl = 0, b0 = 0, b1 = 0;
            if (R.pre > R.a0) { bl = R.pre - R.a0; b0 = R.a0; b1 = R.pre; }
            if (R.bl > bl) { bl = R.bl; b0 = R.bs; b1 = R.be; }
            if (R.suf >= 0 && R.a1 - R.suf > bl) { b0 = R.suf; b1 = R.a1; }
            s_res[0] = b0; s_res[1] = b1;
        }
    }
    __syncthreads();
    bs = s_res[0]; be = s_res[1];
}

template <bool F64, bool EXACT>
__global__ void __launch_bounds__(F64 ? DNT : MNT, F64 ? 2 : 1) minn_peak_kernel(RowView r, int smooth_win, double gate_threshold,
                                                        int has_bounds, int64_t b_lo, int64_t b_hi, int64_t *peak,
                                                        int64_t *gate_span, void *Ms_out, const float *cm,
                                                        int64_t cm_stride, int toff, const __grid_constant__ ExactSrc ex, int32_t *status)
{
    extern __shared__ unsigned mask[];
    __shared__ ArgVal sh_av[MNT / 32];
    __shared__ long long sh_span[2];
    __shared__ ExactScratch sc;
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t n = r.n;
    int st = 0;
    auto finish = [&](long long pk, long long a, long long b) {
        if (tid == 0) { peak[row] = pk; gate_span[2 * row] = a; gate_span[2 * row + 1] = b; if (status) status[row] = st; }
    };
    if (n == 0) { finish(-1, 0, 0); return; }
    const Trailing Ms = make_trailing<F64>(r, row, smooth_win > 1 ? smooth_win : 1, toff);

    // pass 1: global first-argmax of Ms (also the fallback answer), optional Ms output
    const int wv = smooth_win > 1 ? smooth_win : 1;
    Prune pr{cm ? cm + row * cm_stride : nullptr, cm_stride, toff, (wv - 1 + 255) / 256, 0};
    {
        const int64_t nch0 = (n + toff + 255) / 256;
        pr.stage(reinterpret_cast<float *>(mask + nch0 * 8 + 4), nch0);      // chunk maxima live after the bitmask
    }
    if (Ms_out) {
        double v8[8];
        const int64_t g0 = -(int64_t)(((toff % 8) + 8) % 8);
        for (int64_t i0 = g0 + 8LL * tid; i0 < n; i0 += 8LL * blockDim.x) {
            Ms.eval8(i0, v8);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t i = i0 + k;
                if (i < 0 || i >= n) continue;
                if (r.f64) reinterpret_cast<double *>(Ms_out)[row * r.stride + i] = v8[k];
                else reinterpret_cast<float *>(Ms_out)[row * r.stride + i] = (float)v8[k];
            }
        }
    }
    constexpr int KW = 1;                            // chunks in flight per warp (4 was tried: the code outgrew the instruction cache)
    ArgVal best = pruned_argmax<KW>(Ms, n, pr, sh_av);
    if (!(best.v > 0.0)) { finish(-2, 0, 0); return; }

    // exact mode (fused sync pipeline): the float64 value of the maximum always (it scales the gate level), every
    // float32 value within the band of a decision is re-evaluated from the samples (exact.cuh)
    const bool exact_on = EXACT && !F64 && ex.x != nullptr && wv <= 32;
    const ExactSmooth esm{wv, wv - 1, true};
    if (exact_on) {
        const int cnt = collect_at_least(Ms, n, pr, best.v * (1.0 - ex.band), sc);
        if (cnt > EXCAP) st |= OFS_ST_UNRESOLVED;
        else {
            exact_eval_list(ex, row, sc, cnt, esm);
            const ArgVal b = exact_list_argmax(sc, cnt);
            if (cnt > 1) st |= OFS_ST_EXACT | (b.i != best.i ? OFS_ST_CHANGED : 0);
            best = b;
            __syncthreads();
        }
    }

    // pass 2: gate flags -> bitmask (bit index = causal time t = d + toff, so words align with the chunks)
    //         -> longest run (minn.py:155-182)
    const double level = gate_threshold * best.v;
    const double lvl_lo = exact_on ? level * (1.0 - ex.band) : level, lvl_hi = exact_on ? level * (1.0 + ex.band) : level;
    const int64_t nch = (n + toff + 255) / 256;
    if (exact_on) { if (tid == 0) sc.cnt = 0; __syncthreads(); }
    {
        const int nchi = (int)nch, warp = tid >> 5;
        for (int c0 = warp * 32; c0 < nchi; c0 += (int)blockDim.x) {
            const int cl = c0 + lane;
            const bool inr = cl < nchi;
            const bool pass = inr && (!pr.cm || (double)pr.bound(cl) >= lvl_lo);
            if (inr && !pass) {
#pragma unroll
                for (int q = 0; q < 8; ++q) mask[cl * 8 + q] = 0u;
            }
            unsigned m = __ballot_sync(0xffffffffu, pass);
            while (m) {
                int64_t i0s[KW];
                int cs[KW];
                double vk[KW][8];
                int cnt = 0;
#pragma unroll
                for (int u = 0; u < KW; ++u)
                    if (m) {
                        cs[u] = c0 + __ffs(m) - 1;
                        m &= m - 1;
                        // lane = 8 consecutive causal times of the chunk -> one byte of flags; 4 lanes make a mask word
                        i0s[u] = (int64_t)cs[u] * 256 - toff + 8 * lane;
                        cnt = u + 1;
                    }
                Ms.template eval8_multi<KW>(i0s, cnt, vk);
#pragma unroll
                for (int u = 0; u < KW; ++u)
                    if (u < cnt) {
                        unsigned b8 = 0, unc = 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const bool inrow = i0s[u] + k >= 0 && i0s[u] + k < n;
                            if (inrow && vk[u][k] >= level) b8 |= 1u << k;
                            if (inrow && vk[u][k] >= lvl_lo && vk[u][k] <= lvl_hi) unc |= 1u << k;   // float64 decides this flag
                        }
                        if (exact_on) push_hits(sc, i0s[u], unc);
                        unsigned wv32 = b8 << (8 * (lane & 3));
                        wv32 |= __shfl_xor_sync(0xffffffffu, wv32, 1);
                        wv32 |= __shfl_xor_sync(0xffffffffu, wv32, 2);
                        if ((lane & 3) == 0) mask[cs[u] * 8 + (lane >> 2)] = wv32;
                    }
            }
        }
    }
    __syncthreads();
    if (exact_on) {
        const int cnt = sc.cnt;
        if (cnt > EXCAP) st |= OFS_ST_UNRESOLVED;
        else if (cnt > 0) {
            exact_eval_list(ex, row, sc, cnt, esm);
            if (tid < cnt) {
                const long long t = sc.idx[tid] + toff;
                const bool on = sc.val[tid] >= level, was = (mask[t >> 5] >> (t & 31)) & 1u;
                if (on && !was) atomicOr(&mask[t >> 5], 1u << (t & 31));
                if (!on && was) atomicAnd(&mask[t >> 5], ~(1u << (t & 31)));
            }
            st |= OFS_ST_EXACT;
            __syncthreads();
        }
    }
    long long bs_all, be_all;
    longest_run_block(mask, nch * 256, bs_all, be_all);
    if (tid < 32) {
        long long bs = bs_all, be = be_all;
        bs -= toff; be -= toff;
        if (bs < 0) bs = 0;
        if (be < 0) be = 0;
        if (has_bounds) {                                   // minn.py:186-193
            long long s = b_lo > 0 ? b_lo : 0, e = b_hi < n ? b_hi : n;
            if (s >= e) { s = 0; e = n; }
            bs = bs > s ? bs : s;
            be = be < e ? be : e;
        }
        if (tid == 0) { sh_span[0] = bs; sh_span[1] = be; }
    }
    __syncthreads();
    const long long gs = sh_span[0], ge = sh_span[1];
    if (gs >= ge) {                                          // empty gate -> global argmax (minn.py:195-200)
        finish(best.i, best.i, best.i + 1);
        return;
    }
    ArgVal pk = range_argmax(Ms, gs, ge, Ms.toff, sh_av);
    if (exact_on) {                                          // first float64 maximum inside the gate
        const int cnt = collect_range_at_least(Ms, gs, ge, Ms.toff, pk.v * (1.0 - ex.band), sc);
        if (cnt > EXCAP) st |= OFS_ST_UNRESOLVED;
        else if (cnt > 1) {
            exact_eval_list(ex, row, sc, cnt, esm);
            const ArgVal b = exact_list_argmax(sc, cnt);
            st |= OFS_ST_EXACT | (b.i != pk.i ? OFS_ST_CHANGED : 0);
            pk = b;
        }
    }
    finish(pk.i, gs, ge);
}

// ---- combined_sc_min: S&C gate (:337-351) and gated first-segment peak (:183-259) -----------------
__global__ void __launch_bounds__(DNT) sc_gate_kernel(RowView r, double thr, uint8_t *gate, int64_t gstride)
{
    __shared__ ArgVal sh_av[DNT / 32];
    __shared__ long long sh_i[DNT / 32];
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x;
    if (r.n == 0) return;
    ArgVal best{0.0, -1};
    for (int64_t i = tid; i < r.n; i += DNT) {
        const double v = r.at(row, i);
        if (best.i < 0 || v > best.v) { best.v = v; best.i = i; }
    }
    best = block_argmax<false>(best, sh_av);
    long long any = LLONG_MAX;
    for (int64_t i = tid; i < r.n; i += DNT) {
        const double v = r.at(row, i);
        const bool g = best.v > 0.0 ? (v / best.v >= thr) : (v >= thr);
        gate[row * gstride + i] = g;
        if (g && any == LLONG_MAX) any = i;
    }
    any = block_min_i64(any, sh_i);
    if (any == LLONG_MAX && tid == 0) gate[row * gstride + best.i] = 1;   // seed with the strongest sample
}

// Same gate with the chunk maxima of the stripe metric kernel: the row maximum comes from the maxima alone, and a chunk
// whose maximum cannot reach thr * max is all zeros -- M is read only where the gate can be set (a few chunks per frame).
// One warp per chunk; identical output to sc_gate_kernel (the per-sample test is the same float64 expression).
__global__ void __launch_bounds__(DNT) sc_gate_pruned_kernel(RowView r, double thr, uint8_t *gate, int64_t gstride, const float *cm,
                                                             int64_t cm_stride, int toff)
{
    __shared__ ArgVal sh_av[DNT / 32];
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n = r.n;
    const int64_t nch = (n + toff + 255) / 256;
    const float *cmr = cm + row * cm_stride;
    ArgVal best{0.0, -1};
    for (int64_t c = tid; c < nch; c += DNT) {
        const double v = (double)cmr[c];
        if (best.i < 0 || v > best.v) { best.v = v; best.i = c; }
    }
    best = block_argmax<false>(best, sh_av);
    const double mx = best.v;                        // > 0 and thr <= 1 guaranteed by the launcher's fallback rule
    const double level_lo = thr * mx * (1.0 - 1e-9);
    uint8_t *g = gate + row * gstride;
    for (int64_t c = warp; c < nch; c += DNT / 32) {
        const int64_t d0 = c * 256 - toff;           // metric index of the chunk's first sample
        const bool live = (double)cmr[c] >= level_lo;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int64_t d = d0 + 32 * k + lane;
            if (d < 0 || d >= n) continue;
            g[d] = live ? (uint8_t)(r.at(row, d) / mx >= thr) : (uint8_t)0;
        }
    }
    // all-zero row (every chunk maximum 0): nothing passes, the reference seeds the gate with argmax = index 0
    if (!(mx > 0.0)) {
        __syncthreads();
        if (tid == 0) g[0] = 1;
    }
}

template <bool F64>
__global__ void __launch_bounds__(DNT, F64 ? 2 : 3) gated_peak_kernel(RowView r, int smooth_win, const uint8_t *gate,
                                                         int64_t gstride, int has_bounds, int64_t b_lo,
                                                         int64_t b_hi, int64_t *peak)
{
    __shared__ ArgVal sh_av[DNT / 32];
    __shared__ long long sh_i[DNT / 32];
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t n = r.n;
    if (n == 0) { if (tid == 0) peak[row] = -1; return; }
    int64_t s = 0, e = n;
    if (has_bounds) {
        s = b_lo > 0 ? b_lo : 0; e = b_hi < n ? b_hi : n;
        if (s >= e) { s = 0; e = n; }
    }
    const uint8_t *g = gate + row * gstride;
    long long first = LLONG_MAX;
    for (int64_t i = s + tid; i < e; i += DNT)
        if (g[i]) { first = i; break; }
    first = block_min_i64(first, sh_i);
    if (first == LLONG_MAX) { if (tid == 0) peak[row] = -3; return; }
    long long stop = LLONG_MAX;                      // the gate drops -> the streaming detector returns
    for (int64_t i = first + 1 + tid; i < e; i += DNT)
        if (!g[i]) { stop = i; break; }
    stop = block_min_i64(stop, sh_i);
    if (stop == LLONG_MAX) stop = e;
    const Trailing Ms = make_trailing<F64>(r, row, smooth_win > 1 ? smooth_win : 1, 0);
    const ArgVal pk = range_argmax(Ms, first, stop, 0, sh_av);
    if (tid == 0) peak[row] = pk.i;
}

// combined_sc_min in one kernel, no gate array: S&C gate (combined_sc_min.py:337-351) -> first gate segment -> first maximum
// of the trailing-averaged Minn metric inside it (:183-259).  The gate of sample d is the float64 test
// M_sc[d] / max(M_sc) >= thr -- the same expression sc_gate_kernel writes into a byte per sample; here it is evaluated only
// where it matters: the row maximum and the first chunk that can reach the level come from the S&C chunk maxima, the segment
// [first, stop) is walked 2048 samples per step.  Same peak as ofs_sc_gate_pruned + ofs_find_minn_peak_gated, which move a byte
// per sample through HBM twice (0.94 ms of the 3.26 ms combined step on 512 x 1 M-sample rows).
__global__ void __launch_bounds__(DNT, 3) combined_peak_kernel(RowView rm, RowView rs, double thr, const float *cm, int64_t cm_stride,
                                                               int toff, int smooth_win, int has_bounds, int64_t b_lo, int64_t b_hi,
                                                               int64_t *peak, int64_t *gate_span)
{
    __shared__ ArgVal sh_av[DNT / 32];
    __shared__ long long sh_i[DNT / 32];
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t n = rs.n;
    auto done = [&](long long pk, long long a, long long b) {
        if (tid == 0) { peak[row] = pk; if (gate_span) { gate_span[2 * row] = a; gate_span[2 * row + 1] = b; } }
    };
    if (n == 0) { done(-1, 0, 0); return; }
    int64_t s = 0, e = n;
    if (has_bounds) {
        s = b_lo > 0 ? b_lo : 0; e = b_hi < n ? b_hi : n;
        if (s >= e) { s = 0; e = n; }
    }
    const int64_t nch = (n + toff + 255) / 256;
    const float *cmr = cm + row * cm_stride;
    ArgVal best{0.0, -1};
    for (int64_t c = tid; c < nch; c += DNT) {
        const double v = (double)cmr[c];
        if (best.i < 0 || v > best.v) { best.v = v; best.i = c; }
    }
    best = block_argmax<false>(best, sh_av);
    const double mx = best.v;
    long long first = LLONG_MAX;
    if (!(mx > 0.0)) {
        // all-zero row: nothing passes, the reference seeds the gate at argmax = index 0
        if (s == 0) first = 0;
        if (first == LLONG_MAX) { done(-3, 0, 0); return; }
        const Trailing Ms0 = make_trailing<false>(rm, row, smooth_win > 1 ? smooth_win : 1, 0);
        const ArgVal pk0 = range_argmax(Ms0, 0, 1 < e ? 1 : e, 0, sh_av);
        done(pk0.i, 0, 1);
        return;
    }
    const float *ps = reinterpret_cast<const float *>(rs.data) + row * rs.stride;
    auto pass = [&](int64_t d) { return (double)ps[d] / mx >= thr; };
    const double level_lo = thr * mx * (1.0 - 1e-9);
    // first sample of [s, e) that passes: chunks that can reach the level, in order (normally the first one holds it)
    int64_t c_from = (s + toff) / 256;
    while (first == LLONG_MAX) {
        long long cl = LLONG_MAX;
        for (int64_t c = c_from + tid; c < nch; c += DNT)
            if ((double)cmr[c] >= level_lo) { cl = c; break; }
        cl = block_min_i64(cl, sh_i);
        if (cl == LLONG_MAX) break;
        const int64_t d = cl * 256 - toff + tid;                       // one sample per thread
        long long cand = (d >= s && d < e && pass(d)) ? d : LLONG_MAX;
        first = block_min_i64(cand, sh_i);
        c_from = cl + 1;
        if (cl * 256 - toff >= e) break;
    }
    if (first == LLONG_MAX) { done(-3, 0, 0); return; }
    // the gate drops -> the streaming detector returns: first failing sample after `first`
    long long stop = LLONG_MAX;
    for (int64_t base = first + 1; base < e && stop == LLONG_MAX; base += 8LL * DNT) {
        long long cand = LLONG_MAX;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int64_t d = base + k * DNT + tid;
            if (cand == LLONG_MAX && d < e && !pass(d)) cand = d;
        }
        stop = block_min_i64(cand, sh_i);
    }
    if (stop == LLONG_MAX) stop = e;
    const Trailing Ms = make_trailing<false>(rm, row, smooth_win > 1 ? smooth_win : 1, 0);
    const ArgVal pk = range_argmax(Ms, first, stop, 0, sh_av);
    done(pk.i, first, stop);
}

__global__ void __launch_bounds__(DNT) argmax_kernel(RowView r, int64_t *out)
{
    __shared__ ArgVal sh_av[DNT / 32];
    const int64_t row = blockIdx.x;
    ArgVal best{0.0, -1};
    for (int64_t i = threadIdx.x; i < r.n; i += DNT) {
        const double v = r.at(row, i);
        if (best.i < 0 || v > best.v) { best.v = v; best.i = i; }
    }
    best = block_argmax<false>(best, sh_av);
    if (threadIdx.x == 0) out[row] = best.i < 0 ? 0 : best.i;
}

// ---- zc_v2 running-sum threshold (zc_v2.py:288-336) ------------------------------------------------
// local_sum[i] = sum_{j=max(0,i-W+1)}^{i} mag[j]; valid[i] = i >= W.  (The reference's recurrence
// `acc + sample - oldest` drifts by ~1e-16 relative; this is the drift-free windowed sum.)
// One CTA per (row, tile of ZT outputs): float64 inclusive prefix of tile + W-sample halo in smem.
constexpr int ZT = 4096;
__global__ void __launch_bounds__(DNT) zc_stream_kernel(RowView r, int W, double thresh_value, double scale,
                                                        double min_mag, void *local_sum, uint8_t *valid,
                                                        uint8_t *above, int64_t mstride, unsigned *bitmask, int64_t bm_stride)
{
    extern __shared__ double zs[];              // zs[k] = sum of mag[j0 .. j0+k-1]
    __shared__ double wtot[DNT / 32];
    const int64_t row = blockIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * ZT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t j0 = i0 - W + 1;
    if (j0 < 0) j0 = 0;
    const int64_t iend = i0 + ZT < r.n ? i0 + ZT : r.n;
    const int cnt = (int)(iend - j0);           // samples covered
    for (int k = tid; k < cnt; k += DNT) zs[k + 1] = r.at(row, j0 + k);
    if (tid == 0) zs[0] = 0.0;
    __syncthreads();
    int ipt = (cnt + DNT - 1) / DNT;
    ipt |= 1;
    const int s0 = 1 + tid * ipt, s1 = min(s0 + ipt, cnt + 1);
    double acc = 0.0;
    for (int s = s0; s < s1; ++s) { acc += zs[s]; zs[s] = acc; }
    double t = acc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double y = shfl_up_f64(t, o); if (lane >= o) t += y; }
    if (lane == 31) wtot[warp] = t;
    __syncthreads();
    double off = t - acc;
    for (int w = 0; w < warp; ++w) off += wtot[w];
    for (int s = s0; s < s1; ++s) zs[s] += off;
    __syncthreads();
    // (the tile start is a multiple of 32 and every warp walks whole 32-sample groups, so a ballot is one word of the bitmask)
    for (int64_t ib = i0; ib < iend; ib += DNT) {
        const int64_t i = ib + tid;
        bool ab = false;
        if (i < iend) {
            int64_t lo = i - W + 1;
            if (lo < 0) lo = 0;
            const double sum = zs[i + 1 - j0] - zs[lo - j0];
            const double m = r.at(row, i);
            const bool v = i >= W;
            ab = v && (m * scale >= sum * thresh_value) && (m >= min_mag);
            if (local_sum) {
                if (r.f64) reinterpret_cast<double *>(local_sum)[row * r.stride + i] = sum;
                else reinterpret_cast<float *>(local_sum)[row * r.stride + i] = (float)sum;
            }
            if (valid) valid[row * mstride + i] = v;
            if (above) above[row * mstride + i] = ab;
        }
        if (bitmask) {
            const unsigned word = __ballot_sync(0xffffffffu, ab);
            if (lane == 0 && ib + (tid & ~31) < iend) bitmask[row * bm_stride + (ib + tid) / 32] = word;
        }
    }
}

// Bitmask-only form of the threshold for float32 rows (the fused zc_v2 pipeline): zc_stream_kernel moves every sample through
// shared memory seven times as a double (84 B per sample with its 50 % halo) and that, not HBM, bounds it.  Here a thread
// owns 32 CONSECUTIVE samples and only the float32 samples live in shared memory (coalesced load, padded so that the
// 32-sample thread stride is conflict-free): the thread's float64 total goes through a warp scan to an exclusive offset per
// thread, and the window sum of sample i is
//     (offset + running sum of the thread's own samples) - (offset + running sum of the samples of the thread W / 32 to the left)
// both running sums re-accumulated in registers in one pass -- W is a multiple of 32, so sample i - W is element e of that
// other thread.  The thread's 32 flags are one word of the bitmask.  Tile = TMZ outputs + a halo of W samples (extra warps).
constexpr int TMZ = 8192;
__global__ void __launch_bounds__(320, 3) zc_thresh_mask_kernel(const float *mag, int64_t n, int64_t stride, int W, double thresh_value,
                                                                double scale, double min_mag, unsigned *bitmask, int64_t bm_stride,
                                                                int halo_threads)
{
    extern __shared__ __align__(16) unsigned char tsm[];
    const int nthr = (int)blockDim.x;
    float *ms = reinterpret_cast<float *>(tsm);                               // samples, padded: ms[k + k / 32]
    __shared__ double offs[320];                                              // exclusive float64 offset of every thread's 32 samples
    __shared__ double wtot[10];
    const int64_t row = blockIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * TMZ;
    const int64_t j0 = i0 - (int64_t)halo_threads * 32;                       // first sample of the tile (may be negative)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *mr = mag + row * stride;
#pragma unroll
    for (int h = 0; h < 2; ++h) {                                             // two batches of 16 independent coalesced loads per thread
        float ld[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int64_t j = j0 + tid + (16 * h + q) * nthr;
            ld[q] = (j >= 0 && j < n) ? __ldg(mr + j) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) { const int k = tid + (16 * h + q) * nthr; ms[k + (k >> 5)] = ld[q]; }
    }
    __syncthreads();
    const float *mine = ms + tid * 33;                                        // padded position of this thread's first sample
    double run = 0.0;
#pragma unroll
    for (int e = 0; e < 32; ++e) run += (double)mine[e];
    double t = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double y = shfl_up_f64(t, o); if (lane >= o) t += y; }
    if (lane == 31) wtot[warp] = t;
    __syncthreads();
    double off = t - run;
    for (int w = 0; w < warp; ++w) off += wtot[w];
    offs[tid] = off;
    __syncthreads();
    if (tid < halo_threads) return;                                           // halo warps own no outputs
    const int64_t ibase = j0 + (int64_t)tid * 32;                             // first output of this thread
    if (ibase >= n) return;
    const int to = tid - W / 32;                                              // owner of the samples W to the left (>= 0: the halo covers W)
    const float *other = ms + to * 33;
    const double offo = offs[to];
    unsigned word = 0u;
    double own = 0.0, oth = 0.0;
#pragma unroll 8
    for (int e = 0; e < 32; ++e) {
        const int64_t i = ibase + e;
        const double m = (double)mine[e];
        own += m;                                                             // inclusive prefix up to sample i
        oth += (double)other[e];                                              // inclusive prefix up to sample i - W (zeros below the row start)
        const double sum = (own + off) - (oth + offo);                        // local_sum[i] (zc_v2.py:308-318)
        const bool ab = i < n && i >= W && (m * scale >= sum * thresh_value) && (m >= min_mag);
        word |= ab ? 1u << e : 0u;
    }
    bitmask[row * bm_stride + (ibase >> 5)] = word;
}

// ---- gate / hysteresis FSMs ------------------------------------------------------------------------
enum { FSM_AA = 0, FSM_ZC = 1, FSM_RTL = 2 };

struct FsmParams {
    RowView val;              // AA: M rows; ZC: corr_mag rows; RTL: corr_positive rows (f64, or int64 if is_int)
    const void *P;            // AA: complex rows (same precision as val)
    const uint8_t *valid, *above;   // ZC / RTL
    int64_t mstride;
    int is_int;
    int L;                    // AA: half length; ZC: reference length; RTL: timing offset
    int heff;                 // max(hysteresis, 1)
    double thr, fs;
    ofs_event *events;
    int32_t *n_events;
    int max_events;           // event slots per row (>= 1); n_events carries the true gate count
    uint8_t *gate_mask;       // ZC optional
    const unsigned *premask;  // AA optional: above-threshold bitmask already built by the metric kernel (metric_array.cu)
    int64_t premask_stride;   // words per row
};

template <int KIND>
__device__ __forceinline__ double fsm_track_value(const FsmParams &p, int64_t row, int64_t i)
{
    if (KIND == FSM_AA) {
        if (p.val.f64) {
            const double2 z = reinterpret_cast<const double2 *>(p.P)[row * p.val.stride + i];
            return z.x * z.x + z.y * z.y;
        }
        const float2 z = reinterpret_cast<const float2 *>(p.P)[row * p.val.stride + i];
        return (double)z.x * z.x + (double)z.y * z.y;
    }
    if (KIND == FSM_RTL && p.is_int)
        return (double)reinterpret_cast<const long long *>(p.val.data)[row * p.val.stride + i];   // < 2^53
    return p.val.at(row, i);
}

// Byte flags -> mask bits, four at a time: (valid != 0 && above != 0) per byte, then the four 0/1 bytes are gathered into a
// nibble by one multiply (byte k lands on bit 24 + k; the partial products occupy distinct bits, so nothing carries).
__device__ __forceinline__ unsigned flag_nibble(unsigned v, unsigned a)
{
    const unsigned f = __vcmpne4(v, 0u) & __vcmpne4(a, 0u) & 0x01010101u;
    return (f * 0x01020408u) >> 24;
}
__device__ __forceinline__ bool flags_vectorisable(const uint8_t *v, const uint8_t *a)
{
    return v && a && ((reinterpret_cast<uintptr_t>(v) ^ reinterpret_cast<uintptr_t>(a)) & 15) == 0;
}

// Two launches per call.  fsm_gates_kernel: one CTA per row builds the bitmask and walks it into gate intervals (written into
// the row's event slots).  fsm_peaks_kernel: the peak search of the gates, one CTA per (row, gate slice) -- a capture with
// dozens of gates (64 preambles in a 262 144-sample sync_aa capture) would otherwise scan them one after the other on one SM.
template <int KIND>
__global__ void __launch_bounds__(DNT) fsm_gates_kernel(FsmParams p)
{
    extern __shared__ unsigned mask[];
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t n = p.val.n;

    // 1. above-threshold flags of the valid samples -> bitmask
    const int64_t nround = ((n + 31) / 32) * 32;
    if (p.premask) {
        const unsigned *pm = p.premask + row * p.premask_stride;
        for (int64_t w = tid; w < nround / 32; w += DNT) mask[w] = pm[w];
    } else if (KIND != FSM_AA && flags_vectorisable(p.valid + row * p.mstride, p.above + row * p.mstride)) {
        // byte flags, 16 per thread and load: a 16-byte load of each array -> 16 mask bits, OR-ed into the (zeroed) mask at
        // their bit position.  Rows need not be 16-byte aligned, only equally misaligned in both arrays: the first `peel`
        // and the last few samples go one by one.
        const uint8_t *vr = p.valid + row * p.mstride, *ar = p.above + row * p.mstride;
        for (int64_t w = tid; w < nround / 32; w += DNT) mask[w] = 0u;
        __syncthreads();
        int64_t peel = (16 - (int64_t)(reinterpret_cast<uintptr_t>(vr) & 15)) & 15;
        if (peel > n) peel = n;
        const int64_t nfull = (n - peel) / 16;
        const uint4 *v4 = reinterpret_cast<const uint4 *>(vr + peel);
        const uint4 *a4 = reinterpret_cast<const uint4 *>(ar + peel);
        for (int64_t g = tid; g < nfull; g += DNT) {
            const uint4 v = v4[g], a = a4[g];
            const unsigned bits = flag_nibble(v.x, a.x) | (flag_nibble(v.y, a.y) << 4) | (flag_nibble(v.z, a.z) << 8) | (flag_nibble(v.w, a.w) << 12);
            if (bits) {
                const int64_t pos = peel + 16 * g;
                const int sh = (int)(pos & 31);
                atomicOr(&mask[pos >> 5], bits << sh);
                if (sh > 16) atomicOr(&mask[(pos >> 5) + 1], bits >> (32 - sh));
            }
        }
        const int64_t tail0 = peel + 16 * nfull;                      // < 31 samples left: head [0, peel) and tail [tail0, n)
        for (int64_t i = tid; i < peel + (n - tail0); i += DNT) {
            const int64_t j = i < peel ? i : tail0 + (i - peel);
            if (vr[j] && ar[j]) atomicOr(&mask[j >> 5], 1u << (j & 31));
        }
    } else
    for (int64_t i = tid; i < nround; i += DNT) {
        bool f = false;
        if (i < n) {
            if (KIND == FSM_AA) f = i >= p.L && p.val.at(row, i) >= p.thr;
            else f = p.valid[row * p.mstride + i] && p.above[row * p.mstride + i];
        }
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) mask[i / 32] = m;
    }
    __syncthreads();

    // 2. gates = clusters of above-runs separated by fewer than heff below-samples (the reference's open / hysteresis-count /
    //    close loop, sync_aa.py:495-568, as interval logic): a run start opens a gate when the previous above-sample lies at
    //    least heff samples back (or does not exist); a run end is the last above-sample of its gate when the next one is at
    //    least heff samples ahead (or does not exist); the k-th opening pairs with the k-th closing.  Every thread owns an odd
    //    number of consecutive mask words (conflict-free), the "previous / next above-sample" across threads comes from a block
    //    prefix-max / suffix-min, the event slots from a block prefix sum of the per-thread counts.
    constexpr int NW = DNT / 32;
    __shared__ long long s_last[NW], s_first[NW];
    __shared__ int s_cs[NW], s_ce[NW];
    const int warp = tid >> 5;
    const int64_t nw = (n + 31) / 32;
    int wpt = (int)((nw + DNT - 1) / DNT);
    wpt |= 1;
    const int64_t w_lo = (int64_t)tid * wpt < nw ? (int64_t)tid * wpt : nw;
    const int64_t w_hi = w_lo + wpt < nw ? w_lo + wpt : nw;
    long long last = -1, first = LLONG_MAX;
    for (int64_t w = w_lo; w < w_hi; ++w) {
        const unsigned m = mask[w];
        if (m) {
            if (first == LLONG_MAX) first = w * 32 + __ffs(m) - 1;
            last = w * 32 + 31 - __clz(m);
        }
    }
    long long xl = last, xf = first;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long yl = __shfl_up_sync(0xffffffffu, xl, o), yf = __shfl_down_sync(0xffffffffu, xf, o);
        if (lane >= o && yl > xl) xl = yl;
        if (lane + o < 32 && yf < xf) xf = yf;
    }
    if (lane == 31) s_last[warp] = xl;
    if (lane == 0) s_first[warp] = xf;
    long long prev_last = __shfl_up_sync(0xffffffffu, xl, 1), next_first = __shfl_down_sync(0xffffffffu, xf, 1);
    if (lane == 0) prev_last = -1;
    if (lane == 31) next_first = LLONG_MAX;
    __syncthreads();
    for (int w = 0; w < warp; ++w) if (s_last[w] > prev_last) prev_last = s_last[w];
    for (int w = warp + 1; w < NW; ++w) if (s_first[w] < next_first) next_first = s_first[w];

    ofs_event *evs = p.events + row * (int64_t)p.max_events;
    auto walk_starts = [&](bool write, int off) -> int {
        long long run_last = prev_last;
        int c = 0;
        for (int64_t w = w_lo; w < w_hi; ++w) {
            const unsigned m = mask[w];
            if (!m) continue;
            const unsigned carry = w > 0 ? mask[w - 1] >> 31 : 0u;
            unsigned st = m & ~((m << 1) | carry);
            const long long base = w * 32;
            while (st) {
                const int a = __ffs(st) - 1;
                st &= st - 1;
                const unsigned below = m & ((1u << a) - 1u);
                const long long prev = below ? base + 31 - __clz(below) : run_last;
                if (prev < 0 || base + a - prev - 1 >= p.heff) {
                    if (write && off + c < p.max_events) evs[off + c].gate_start = base + a;
                    ++c;
                }
            }
            run_last = base + 31 - __clz(m);
        }
        return c;
    };
    auto walk_ends = [&](bool write, int off_end) -> int {      // descending; off_end = slot after this thread's last closing
        long long run_first = next_first;
        int c = 0;
        for (int64_t w = w_hi - 1; w >= w_lo; --w) {
            const unsigned m = mask[w];
            if (!m) continue;
            const unsigned nb = w + 1 < nw ? mask[w + 1] & 1u : 0u;
            unsigned en = m & ~((m >> 1) | (nb << 31));
            const long long base = w * 32;
            while (en) {
                const int b = 31 - __clz(en);
                en &= ~(1u << b);
                const unsigned above = b == 31 ? 0u : m & ~((2u << b) - 1u);
                const long long next = above ? base + __ffs(above) - 1 : run_first;
                const long long pos = base + b;
                if (next == LLONG_MAX || next - pos - 1 >= p.heff) {
                    const int idx = off_end - 1 - c;
                    if (write && idx < p.max_events) {
                        const bool closes = next != LLONG_MAX || (n - 1 - pos) >= p.heff;
                        evs[idx].gate_end = closes ? pos + p.heff : n - 1;
                        evs[idx].closed = closes;
                    }
                    ++c;
                }
            }
            run_first = base + __ffs(m) - 1;
        }
        return c;
    };
    const int cs = walk_starts(false, 0), ce = walk_ends(false, 0);
    int xs = cs, xe = ce;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ys = __shfl_up_sync(0xffffffffu, xs, o), ye = __shfl_up_sync(0xffffffffu, xe, o);
        if (lane >= o) { xs += ys; xe += ye; }
    }
    if (lane == 31) { s_cs[warp] = xs; s_ce[warp] = xe; }
    __syncthreads();
    int off_s = xs - cs, off_e = xe - ce, total = 0;
    for (int w = 0; w < NW; ++w) {
        if (w < warp) { off_s += s_cs[w]; off_e += s_ce[w]; }
        total += s_cs[w];
    }
    if (cs) walk_starts(true, off_s);
    if (ce) walk_ends(true, off_e + ce);
    if (tid == 0) p.n_events[row] = total;
}

template <int KIND>
__global__ void __launch_bounds__(DNT) fsm_peaks_kernel(FsmParams p)
{
    __shared__ ArgVal sh_av[DNT / 32];
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t n = p.val.n;
    const int total = p.n_events[row];
    const int cnt = total < p.max_events ? total : p.max_events;

    // 3. peak of each gate over [open, close] with the reference's tie rule
    for (int e = blockIdx.y; e < cnt; e += gridDim.y) {
        ofs_event *slot = p.events + row * (int64_t)p.max_events + e;
        const long long gs = slot->gate_start, gc = slot->gate_end;
        const int closed = slot->closed;
        ArgVal pk{0.0, -1};
        for (int64_t i = gs + tid; i <= gc; i += DNT) {
            const double v = fsm_track_value<KIND>(p, row, i);
            if (KIND == FSM_RTL) { if (pk.i < 0 || v >= pk.v) { pk.v = v; pk.i = i; } }
            else { if (pk.i < 0 || v > pk.v) { pk.v = v; pk.i = i; } }
        }
        pk = block_argmax<KIND == FSM_RTL>(pk, sh_av);     // (barriers inside: every thread has read the slot by now)
        if (KIND == FSM_ZC && p.gate_mask) {   // zc_v2.py:409,444: (open, close] when closed, [open, n) otherwise
            const int64_t a = closed ? gs + 1 : gs;
            for (int64_t i = a + tid; i <= gc; i += DNT) p.gate_mask[row * p.mstride + i] = 1;
        }
        if (tid == 0) {
            ofs_event ev{};
            ev.peak_index = pk.i; ev.gate_start = gs; ev.closed = closed;
            if (KIND == FSM_AA) {
                ev.gate_end = closed ? gc : n;
                ev.aux = pk.i - 2LL * p.L + 1;
                ev.value = p.val.at(row, pk.i);
                if (p.val.f64) { const double2 z = reinterpret_cast<const double2 *>(p.P)[row * p.val.stride + pk.i]; ev.p_re = z.x; ev.p_im = z.y; }
                else { const float2 z = reinterpret_cast<const float2 *>(p.P)[row * p.val.stride + pk.i]; ev.p_re = z.x; ev.p_im = z.y; }
                ev.cfo = atan2(ev.p_im, ev.p_re) * p.fs / (2.0 * 3.14159265358979323846 * (double)p.L);
            } else if (KIND == FSM_ZC) {
                ev.gate_end = closed ? gc : n;
                const long long ds = pk.i - p.L + 1;
                ev.aux = ds > 0 ? ds : 0;
                ev.value = pk.v;
            } else {
                ev.gate_end = closed ? gc + 1 : n;     // segment end (exclusive), minn_rtl.py:799
                ev.aux = pk.i + p.L;
                ev.value = pk.v;
            }
            *slot = ev;
        }
        __syncthreads();
    }
}

// ---- host launchers ----------------------------------------------------------------------------------
static int rows_ok(const ofs_rows *M, const char *who)
{
    OFS_REQUIRE(M && (M->data || M->n_rows == 0 || M->n == 0), "%s: null metric rows", who);
    OFS_REQUIRE(M->n_rows >= 0 && M->n >= 0 && M->stride >= M->n, "%s: bad row geometry", who);
    OFS_REQUIRE(M->n_rows < (1LL << 31), "%s: too many rows", who);
    return OFS_OK;
}
static RowView view(const ofs_rows *M) { return RowView{M->data, M->f64, M->n, M->stride}; }

template <typename K>
static int set_mask_smem(K kern, size_t bytes)
{
    OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return OFS_OK;
}
static size_t mask_bytes(int64_t n) { return (size_t)((n + 31) / 32) * 4 + 16; }

// plateau detector; ex != nullptr: exact mode of the fused sync pipeline (status: int32[n_rows], OFS_ST_* bits)
int launch_plateau(const ofs_rows *M, const float *chunk_max, int64_t cm_stride, int32_t toff, int32_t cp_len, int32_t lookahead,
                   int32_t smooth_win, int64_t *plateau_end, const ExactSrc *ex, int32_t *status, void *stream)
{
    if (int rc = rows_ok(M, "ofs_find_plateau_end")) return rc;
    OFS_REQUIRE(plateau_end, "ofs_find_plateau_end: null output");
    OFS_REQUIRE(!chunk_max || (toff >= 0 && cm_stride >= (M->n + toff + 255) / 256), "ofs_find_plateau_end: bad chunk_max geometry");
    if (M->n_rows == 0) return OFS_OK;
    const size_t cms = chunk_max ? (size_t)((M->n + toff + 255) / 256 + 8) * sizeof(float) : 0;
    OFS_REQUIRE(cms <= 200 * 1024, "ofs_find_plateau_end: rows too long for the pruned path");
    auto go = [&](auto kern) -> int {
        if (cms > 48 * 1024) OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cms));
        kern<<<(unsigned)M->n_rows, DNT, cms, (cudaStream_t)stream>>>(view(M), cp_len, lookahead, smooth_win, plateau_end, chunk_max,
                                                                     cm_stride, chunk_max ? toff : 0, ex ? *ex : ExactSrc{}, status);
        return check_launch("plateau_kernel");
    };
    if (M->f64) return go(plateau_kernel<true, false>);
    return (ex && ex->x) ? go(plateau_kernel<false, true>) : go(plateau_kernel<false, false>);
}

}  // namespace ofs
using namespace ofs;

OFS_API int ofs_find_plateau_end_pruned(const ofs_rows *M, const float *chunk_max, int64_t cm_stride, int32_t toff,
                                        int32_t cp_len, int32_t lookahead, int32_t smooth_win, int64_t *plateau_end,
                                        void *stream)
{
    OFS_TRACE();
    return launch_plateau(M, chunk_max, cm_stride, toff, cp_len, lookahead, smooth_win, plateau_end, nullptr, nullptr, stream);
}

OFS_API int ofs_find_plateau_end(const ofs_rows *M, int32_t cp_len, int32_t lookahead, int32_t smooth_win,
                                 int64_t *plateau_end, void *stream)
{
    OFS_TRACE();
    return ofs_find_plateau_end_pruned(M, nullptr, 0, 0, cp_len, lookahead, smooth_win, plateau_end, stream);
}

namespace ofs {
int launch_minn_peak(const ofs_rows *M, const float *chunk_max, int64_t cm_stride, int32_t toff, int32_t smooth_win,
                     double gate_threshold, int32_t has_bounds, int64_t bound_lo, int64_t bound_hi, int64_t *peak,
                     int64_t *gate_span, void *Ms, const ExactSrc *ex, int32_t *status, void *stream)
{
    if (int rc = rows_ok(M, "ofs_find_minn_peak")) return rc;
    OFS_REQUIRE(peak && gate_span, "ofs_find_minn_peak: null output");
    if (!chunk_max) toff = 0;
    OFS_REQUIRE(toff >= 0 && M->n + toff <= MASK_MAX_N, "ofs_find_minn_peak: rows longer than %lld unsupported",
                (long long)MASK_MAX_N);
    OFS_REQUIRE(!chunk_max || cm_stride >= (M->n + toff + 255) / 256, "ofs_find_minn_peak: bad chunk_max geometry");
    if (M->n_rows == 0) return OFS_OK;
    const int64_t nch_h = (M->n + toff + 255) / 256;
    const size_t sm = mask_bytes(nch_h * 256) + (chunk_max ? (size_t)(nch_h + 8) * sizeof(float) : 0);
    OFS_REQUIRE(sm <= 195 * 1024, "ofs_find_minn_peak: rows too long");   // + 31 KB static (longest_run_block) <= 227 KB
    // float rows whose mask leaves room for one CTA per SM only: 24 warps instead of 8 hide the latency of the window fetches
    const int nt = (!M->f64 && sm > 72 * 1024) ? MNT : DNT;
    auto go = [&](auto kern) -> int {
        if (int rc = set_mask_smem(kern, sm)) return rc;
        kern<<<(unsigned)M->n_rows, nt, sm, (cudaStream_t)stream>>>(view(M), smooth_win, gate_threshold, has_bounds, bound_lo, bound_hi,
                                                                   peak, gate_span, Ms, chunk_max, cm_stride, toff, ex ? *ex : ExactSrc{},
                                                                   status);
        return check_launch("minn_peak_kernel");
    };
    if (M->f64) return go(minn_peak_kernel<true, false>);
    return (ex && ex->x) ? go(minn_peak_kernel<false, true>) : go(minn_peak_kernel<false, false>);
}
}  // namespace ofs

OFS_API int ofs_find_minn_peak_pruned(const ofs_rows *M, const float *chunk_max, int64_t cm_stride, int32_t toff,
                                      int32_t smooth_win, double gate_threshold, int32_t has_bounds, int64_t bound_lo,
                                      int64_t bound_hi, int64_t *peak, int64_t *gate_span, void *Ms, void *stream)
{
    OFS_TRACE();
    return launch_minn_peak(M, chunk_max, cm_stride, toff, smooth_win, gate_threshold, has_bounds, bound_lo, bound_hi, peak, gate_span,
                            Ms, nullptr, nullptr, stream);
}

OFS_API int ofs_find_minn_peak(const ofs_rows *M, int32_t smooth_win, double gate_threshold, int32_t has_bounds,
                               int64_t bound_lo, int64_t bound_hi, int64_t *peak, int64_t *gate_span, void *Ms,
                               void *stream)
{
    OFS_TRACE();
    return ofs_find_minn_peak_pruned(M, nullptr, 0, 0, smooth_win, gate_threshold, has_bounds, bound_lo, bound_hi, peak,
                                     gate_span, Ms, stream);
}

OFS_API int ofs_sc_gate_pruned(const ofs_rows *Msc, const float *chunk_max, int64_t cm_stride, int32_t toff, double threshold,
                               uint8_t *gate, int64_t gate_stride, void *stream)
{
    OFS_TRACE();
    if (int rc = rows_ok(Msc, "ofs_sc_gate")) return rc;
    OFS_REQUIRE(gate && gate_stride >= Msc->n, "ofs_sc_gate: bad gate buffer");
    OFS_REQUIRE(!chunk_max || (toff >= 0 && cm_stride >= (Msc->n + toff + 255) / 256), "ofs_sc_gate: bad chunk_max geometry");
    if (Msc->n_rows == 0 || Msc->n == 0) return OFS_OK;
    // pruned form: float32 rows from the stripe kernel (non-negative metric, so a positive row maximum exists unless the row
    // is all zero -- then chunk maxima are all zero, nothing passes and the seed rule needs the full kernel) and thr in (0, 1]
    if (chunk_max && !Msc->f64 && threshold > 0.0 && threshold <= 1.0) {
        sc_gate_pruned_kernel<<<(unsigned)Msc->n_rows, DNT, 0, (cudaStream_t)stream>>>(view(Msc), threshold, gate, gate_stride, chunk_max,
                                                                                     cm_stride, toff);
        return check_launch("sc_gate_pruned_kernel");      // (rows whose maximum is zero are seeded at index 0 by the same kernel)
    }
    sc_gate_kernel<<<(unsigned)Msc->n_rows, DNT, 0, (cudaStream_t)stream>>>(view(Msc), threshold, gate, gate_stride);
    return check_launch("sc_gate_kernel");
}

OFS_API int ofs_sc_gate(const ofs_rows *Msc, double threshold, uint8_t *gate, int64_t gate_stride, void *stream)
{
    OFS_TRACE();
    return ofs_sc_gate_pruned(Msc, nullptr, 0, 0, threshold, gate, gate_stride, stream);
}

OFS_API int ofs_find_minn_peak_gated(const ofs_rows *M, int32_t smooth_win, const uint8_t *gate, int64_t gate_stride,
                                     int32_t has_bounds, int64_t bound_lo, int64_t bound_hi, int64_t *peak, void *stream)
{
    OFS_TRACE();
    if (int rc = rows_ok(M, "ofs_find_minn_peak_gated")) return rc;
    OFS_REQUIRE(peak && (gate || M->n == 0) && gate_stride >= M->n, "ofs_find_minn_peak_gated: bad arguments");
    if (M->n_rows == 0) return OFS_OK;
    if (M->f64)
        gated_peak_kernel<true><<<(unsigned)M->n_rows, DNT, 0, (cudaStream_t)stream>>>(view(M), smooth_win, gate, gate_stride, has_bounds,
                                                                                      bound_lo, bound_hi, peak);
    else
        gated_peak_kernel<false><<<(unsigned)M->n_rows, DNT, 0, (cudaStream_t)stream>>>(view(M), smooth_win, gate, gate_stride, has_bounds,
                                                                                       bound_lo, bound_hi, peak);
    return check_launch("gated_peak_kernel");
}

OFS_API int ofs_combined_peak(const ofs_rows *M_minn, const ofs_rows *M_sc, const float *chunk_max_sc, int64_t cm_stride, int32_t toff,
                              double threshold, int32_t smooth_win, int32_t has_bounds, int64_t bound_lo, int64_t bound_hi,
                              int64_t *peak, int64_t *gate_span, void *stream)
{
    OFS_TRACE();
    if (int rc = rows_ok(M_minn, "ofs_combined_peak")) return rc;
    if (int rc = rows_ok(M_sc, "ofs_combined_peak")) return rc;
    OFS_REQUIRE(peak && chunk_max_sc, "ofs_combined_peak: null argument");
    OFS_REQUIRE(!M_minn->f64 && !M_sc->f64, "ofs_combined_peak: float32 rows (the stripe kernel's output) only");
    OFS_REQUIRE(M_minn->n == M_sc->n && M_minn->n_rows == M_sc->n_rows, "ofs_combined_peak: the two metrics must have the same shape");
    OFS_REQUIRE(threshold > 0.0 && threshold <= 1.0, "ofs_combined_peak: threshold must be in (0, 1]");
    OFS_REQUIRE(toff >= 0 && cm_stride >= (M_sc->n + toff + 255) / 256, "ofs_combined_peak: bad chunk_max geometry");
    if (M_sc->n_rows == 0) return OFS_OK;
    combined_peak_kernel<<<(unsigned)M_sc->n_rows, DNT, 0, (cudaStream_t)stream>>>(view(M_minn), view(M_sc), threshold, chunk_max_sc, cm_stride,
                                                                                 toff, smooth_win, has_bounds, bound_lo, bound_hi, peak,
                                                                                 gate_span);
    return check_launch("combined_peak_kernel");
}

OFS_API int ofs_argmax(const ofs_rows *M, int64_t *index, void *stream)
{
    OFS_TRACE();
    if (int rc = rows_ok(M, "ofs_argmax")) return rc;
    OFS_REQUIRE(index, "ofs_argmax: null output");
    if (M->n_rows == 0) return OFS_OK;
    argmax_kernel<<<(unsigned)M->n_rows, DNT, 0, (cudaStream_t)stream>>>(view(M), index);
    return check_launch("argmax_kernel");
}

OFS_API int ofs_zc_streaming_detection(const ofs_rows *corr_mag, int32_t window, int32_t thresh_value, int32_t frac_bits,
                                       double min_corr_mag, void *local_sum, uint8_t *valid, uint8_t *above,
                                       int64_t mask_stride, void *stream)
{
    OFS_TRACE();
    if (int rc = rows_ok(corr_mag, "ofs_zc_streaming_detection")) return rc;
    OFS_REQUIRE(local_sum && valid && above && mask_stride >= corr_mag->n, "ofs_zc_streaming_detection: bad outputs");
    OFS_REQUIRE(frac_bits >= 0 && frac_bits < 62, "ofs_zc_streaming_detection: bad frac_bits");
    if (corr_mag->n_rows == 0 || corr_mag->n == 0) return OFS_OK;
    const int W = window > 1 ? window : 1;
    OFS_REQUIRE(W <= 16384, "ofs_zc_streaming_detection: window > 16384 unsupported");
    dim3 grid((unsigned)((corr_mag->n + ZT - 1) / ZT), (unsigned)corr_mag->n_rows);
    OFS_REQUIRE(corr_mag->n_rows < 65536, "ofs_zc_streaming_detection: too many rows");
    const size_t zsm = (size_t)(ZT + W + 2) * sizeof(double);
    OFS_CUDA(cudaFuncSetAttribute(zc_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsm));
    zc_stream_kernel<<<grid, DNT, zsm, (cudaStream_t)stream>>>(view(corr_mag), W, (double)thresh_value,
                                                            (double)(1LL << frac_bits), min_corr_mag, local_sum, valid, above,
                                                            mask_stride, nullptr, 0);
    return check_launch("zc_stream_kernel");
}

template <int KIND>
static int launch_fsm(FsmParams &p, int64_t n_rows, cudaStream_t stream)
{
    OFS_REQUIRE(p.val.n <= MASK_MAX_N, "gate FSM: rows longer than %lld unsupported", (long long)MASK_MAX_N);
    OFS_REQUIRE(p.max_events >= 1, "gate FSM: max_events must be >= 1");
    if (n_rows == 0) return OFS_OK;
    const size_t sm = mask_bytes(p.val.n);
    if (int rc = set_mask_smem(fsm_gates_kernel<KIND>, sm)) return rc;
    fsm_gates_kernel<KIND><<<(unsigned)n_rows, DNT, sm, stream>>>(p);
    if (int rc = check_launch("fsm_gates_kernel")) return rc;
    // gate slices per row: enough CTAs to fill the machine a few times over when the rows are few
    int64_t slices = (4LL * sm_count() + n_rows - 1) / n_rows;
    if (slices > p.max_events) slices = p.max_events;
    if (slices > 65535) slices = 65535;
    if (slices < 1) slices = 1;
    fsm_peaks_kernel<KIND><<<dim3((unsigned)n_rows, (unsigned)slices), DNT, 0, stream>>>(p);
    return check_launch("fsm_peaks_kernel");
}

OFS_API int ofs_aa_events(const ofs_rows *M, const void *P, int32_t L, double threshold, int32_t hysteresis,
                          double sample_rate, ofs_event *events, int32_t *n_events, int32_t max_events, void *stream)
{
    OFS_TRACE();
    if (int rc = rows_ok(M, "ofs_aa_events")) return rc;
    OFS_REQUIRE(P && events && n_events && L > 0, "ofs_aa_events: bad arguments");
    FsmParams p{};
    p.val = view(M); p.P = P; p.L = L; p.heff = hysteresis > 1 ? hysteresis : 1; p.thr = threshold; p.fs = sample_rate;
    p.events = events; p.n_events = n_events; p.max_events = max_events;
    return launch_fsm<FSM_AA>(p, M->n_rows, (cudaStream_t)stream);
}

OFS_API int ofs_aa_detect(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_antennas, int64_t n,
                          int64_t x_frame_stride, int64_t x_branch_stride, int32_t L, double threshold, int32_t hysteresis,
                          double sample_rate, float *M, void *P_c64, float *R, int64_t out_stride, uint32_t *mask_ws,
                          int64_t mask_stride, ofs_event *events, int32_t *n_events, int32_t max_events, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(x && M && P_c64 && mask_ws && events && n_events && L > 0, "ofs_aa_detect: null argument");
    OFS_REQUIRE(n_frames >= 0 && n_antennas >= 1 && n >= 0 && n_frames < (1LL << 31), "ofs_aa_detect: bad geometry");
    if (n_frames == 0) return OFS_OK;
    if (n == 0) { OFS_CUDA(cudaMemsetAsync(n_events, 0, (size_t)n_frames * sizeof(int32_t), (cudaStream_t)stream)); return OFS_OK; }
    if (int rc = launch_metric_array(x, in_dtype, n_frames, n_antennas, n, x_frame_stride, x_branch_stride, L, M, P_c64, R,
                                     out_stride, mask_ws, mask_stride, threshold, (cudaStream_t)stream))
        return rc;
    FsmParams p{};
    p.val = RowView{M, 0, n, out_stride}; p.P = P_c64; p.L = L; p.heff = hysteresis > 1 ? hysteresis : 1; p.thr = threshold;
    p.fs = sample_rate; p.events = events; p.n_events = n_events; p.max_events = max_events; p.premask = mask_ws; p.premask_stride = mask_stride;
    return launch_fsm<FSM_AA>(p, n_frames, (cudaStream_t)stream);
}

OFS_API int ofs_zc_events(const ofs_rows *corr_mag, const uint8_t *valid, const uint8_t *above, int64_t mask_stride,
                          int32_t reference_length, int32_t hysteresis, ofs_event *events, int32_t *n_events,
                          int32_t max_events, uint8_t *gate_mask, void *stream)
{
    OFS_TRACE();
    if (int rc = rows_ok(corr_mag, "ofs_zc_events")) return rc;
    OFS_REQUIRE(valid && above && events && n_events && mask_stride >= corr_mag->n, "ofs_zc_events: bad arguments");
    if (gate_mask && corr_mag->n_rows > 0)
        OFS_CUDA(cudaMemsetAsync(gate_mask, 0, (size_t)corr_mag->n_rows * mask_stride, (cudaStream_t)stream));
    FsmParams p{};
    p.val = view(corr_mag); p.valid = valid; p.above = above; p.mstride = mask_stride; p.L = reference_length;
    p.heff = hysteresis > 1 ? hysteresis : 1; p.events = events; p.n_events = n_events; p.max_events = max_events; p.gate_mask = gate_mask;
    return launch_fsm<FSM_ZC>(p, corr_mag->n_rows, (cudaStream_t)stream);
}

OFS_API int ofs_zc_detect(const ofs_rows *corr_mag, int32_t window, int32_t thresh_value, int32_t frac_bits, double min_corr_mag,
                          int32_t reference_length, int32_t hysteresis, uint32_t *mask_ws, int64_t mask_stride, ofs_event *events,
                          int32_t *n_events, int32_t max_events, void *stream)
{
    OFS_TRACE();
    if (int rc = rows_ok(corr_mag, "ofs_zc_detect")) return rc;
    OFS_REQUIRE(mask_ws && events && n_events && mask_stride >= (corr_mag->n + 31) / 32, "ofs_zc_detect: bad arguments");
    OFS_REQUIRE(frac_bits >= 0 && frac_bits < 62, "ofs_zc_detect: bad frac_bits");
    if (corr_mag->n_rows == 0 || corr_mag->n == 0) return OFS_OK;
    const int W = window > 1 ? window : 1;
    OFS_REQUIRE(W <= 16384, "ofs_zc_detect: window > 16384 unsupported");
    OFS_REQUIRE(corr_mag->n_rows < 65536, "ofs_zc_detect: too many rows");
    if (!corr_mag->f64 && W <= 2048 && W % 32 == 0) {
        // float32 rows: the register-prefix kernel (one mask word per thread; 256 + up to 64 halo threads)
        const int halo_threads = (W / 32 + 31) / 32 * 32;                     // whole warps, >= W samples
        const int nthr = TMZ / 32 + halo_threads;
        const size_t tsm = (size_t)(nthr * 32 + nthr + 2) * sizeof(float);
        dim3 grid2((unsigned)((corr_mag->n + TMZ - 1) / TMZ), (unsigned)corr_mag->n_rows);
        static PerDeviceOnce once;
        if (!once.done()) {
            OFS_CUDA(cudaFuncSetAttribute(zc_thresh_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            once.mark();
        }
        zc_thresh_mask_kernel<<<grid2, nthr, tsm, (cudaStream_t)stream>>>(reinterpret_cast<const float *>(corr_mag->data), corr_mag->n,
                                                                         corr_mag->stride, W, (double)thresh_value,
                                                                         (double)(1LL << frac_bits), min_corr_mag, mask_ws, mask_stride,
                                                                         halo_threads);
        if (int rc = check_launch("zc_thresh_mask_kernel")) return rc;
    } else {
        dim3 grid((unsigned)((corr_mag->n + ZT - 1) / ZT), (unsigned)corr_mag->n_rows);
        const size_t zsm = (size_t)(ZT + W + 2) * sizeof(double);
        OFS_CUDA(cudaFuncSetAttribute(zc_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsm));
        zc_stream_kernel<<<grid, DNT, zsm, (cudaStream_t)stream>>>(view(corr_mag), W, (double)thresh_value, (double)(1LL << frac_bits),
                                                                min_corr_mag, nullptr, nullptr, nullptr, 0, mask_ws, mask_stride);
        if (int rc = check_launch("zc_stream_kernel")) return rc;
    }
    FsmParams p{};
    p.val = view(corr_mag); p.L = reference_length; p.heff = hysteresis > 1 ? hysteresis : 1; p.events = events; p.n_events = n_events;
    p.max_events = max_events; p.premask = mask_ws; p.premask_stride = mask_stride;
    return launch_fsm<FSM_ZC>(p, corr_mag->n_rows, (cudaStream_t)stream);
}

OFS_API int ofs_zc_v2_detect(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n, const void *ref_c128,
                             int32_t nr, int32_t normalize, int32_t window, int32_t thresh_value, int32_t frac_bits,
                             double min_corr_mag, int32_t hysteresis, float *mag_ws, int64_t mag_stride, uint32_t *mask_ws,
                             int64_t mask_stride, ofs_event *events, int32_t *n_events, int32_t max_events, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(x && ref_c128 && mag_ws && mask_ws && events && n_events, "ofs_zc_v2_detect: null argument");
    OFS_REQUIRE(in_dtype == OFS_C64 || in_dtype == OFS_IQ16, "ofs_zc_v2_detect: complex64 or int16 IQ captures (float32 pipeline)");
    const int64_t n_out = n + nr - 1;
    OFS_REQUIRE(mag_stride >= n_out && mask_stride >= (n_out + 31) / 32, "ofs_zc_v2_detect: workspace pitch too small");
    // |corr| only (4 bytes per sample leave the filter), mode 1 = per-branch normalisation (zc_v2.py:488-495), mode 2 = raw
    if (int rc = ofs_zc_matched_filter(x, in_dtype, n_frames, n_branches, n, ref_c128, nr, normalize ? 1 : 2, 0, nullptr, mag_ws, mag_stride,
                                       stream))
        return rc;
    ofs_rows rows{mag_ws, 0, 0, n_frames, n_out, mag_stride};
    return ofs_zc_detect(&rows, window, thresh_value, frac_bits, min_corr_mag, nr, hysteresis, mask_ws, mask_stride, events, n_events,
                         max_events, stream);
}

OFS_API int ofs_minn_rtl_events(const void *corr_positive, int32_t is_int, const uint8_t *valid, const uint8_t *above,
                                int64_t n_rows, int64_t n, int64_t stride, int32_t hysteresis, int32_t timing_offset,
                                ofs_event *events, int32_t *n_events, int32_t max_events, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(corr_positive && valid && above && events && n_events, "ofs_minn_rtl_events: null argument");
    OFS_REQUIRE(n_rows >= 0 && n >= 0 && stride >= n && n_rows < (1LL << 31), "ofs_minn_rtl_events: bad geometry");
    FsmParams p{};
    p.val = RowView{corr_positive, 1, n, stride}; p.is_int = is_int; p.valid = valid; p.above = above; p.mstride = stride;
    p.L = timing_offset; p.heff = hysteresis > 1 ? hysteresis : 1; p.events = events; p.n_events = n_events; p.max_events = max_events;
    return launch_fsm<FSM_RTL>(p, n_rows, (cudaStream_t)stream);
}

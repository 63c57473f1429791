// Precise ("tile") autocorrelation-metric kernel: float64 prefix sums in shared memory.
//
// Serves every lag / branch count / input dtype of
//   sc.sc_streaming_metric                       sc.py:42-78
//   combined_sc_min.schmidl_cox_streaming_metric combined_sc_min.py:116-164
//   minn.minn_streaming_metric(_parameterized)   minn.py:59-112, 697-751
//   sync_aa.aa_detect_streaming loop 1           sync_aa.py:458-493
// through the closed forms of SURVEY.md Appendix B:
//   Sc[k] = sum_{j<k} x[j] conj(x[j+D])  (branch-summed),  Se[k] = sum_{j<k} |x[j]|^2
//   P(d) = sum over windows Sc[d+b]-Sc[d+a],  R(d) = Se[d+rb]-Se[d+ra]
// One CTA = one tile of T outputs of one frame: products are formed from coalesced global loads
// (exact in float64 for c64 / iq16 input), written to smem, scanned in place (thread-serial +
// warp-shuffle + CTA carry), and every output is a difference of 2..6 smem taps.
#include "common.cuh"

namespace ofs {

struct TileParams {
    const void *x;
    void *M, *P, *R;
    int64_t L, xfs, xbs, out_len, out_stride;
    int dtype, nb, kind, D, out_f64;
    int pa0, pb0, pa1, pb1, npw, ra, rb, conjP;
    int offmin, T, NP, tiles_per_frame;
    int aa_L;
};

constexpr int TILE_NT = 512;

// One slot: q = sum_b x_b[j] conj(x_b[j+D]), e = sum_b |x_b[j]|^2.
// One branch: float64 products (exact for complex64 / int16 input).  Several branches (antenna arrays): the loads of
// four branches are issued together (memory-level parallelism) and, for 32-bit inputs, products and the sum over
// branches are formed in fp32 (one rounding per product, 6e-8 relative, far inside the 1e-4 metric tolerance) and
// promoted to float64 once per slot, where the long sums live.
template <int DT>
__device__ __forceinline__ void phase1_slot(const unsigned char *xf, int nb, size_t xbs, int64_t j, int D, bool has_lag,
                                            double &qr, double &qi, double &e)
{
    using In = typename InT<DT>::type;
    const In *x0 = reinterpret_cast<const In *>(xf) + j;
    if (DT == OFS_C128 || nb == 1) {
        for (int b = 0; b < nb; ++b) {
            const In *xb = x0 + (size_t)b * xbs;
            const In av = __ldg(xb);
            const double ax = (double)av.x, ay = (double)av.y;
            e += ax * ax + ay * ay;
            if (has_lag) {
                const In cv = __ldg(xb + D);
                const double cx = (double)cv.x, cy = (double)cv.y;
                qr += ax * cx + ay * cy;      // x[j] * conj(x[j+D])
                qi += ay * cx - ax * cy;
            }
        }
        return;
    }
    float fr[4] = {0.f, 0.f, 0.f, 0.f}, fi[4] = {0.f, 0.f, 0.f, 0.f}, fe[4] = {0.f, 0.f, 0.f, 0.f};
    int b = 0;
    for (; b + 4 <= nb; b += 4) {
        In av[4], cv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const In *xb = x0 + (size_t)(b + u) * xbs;
            av[u] = __ldg(xb);
            cv[u] = has_lag ? __ldg(xb + D) : In{};
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float ax = (float)av[u].x, ay = (float)av[u].y, cx = (float)cv[u].x, cy = (float)cv[u].y;
            fe[u] = fmaf(ax, ax, fmaf(ay, ay, fe[u]));
            fr[u] = fmaf(ax, cx, fmaf(ay, cy, fr[u]));
            fi[u] = fmaf(ay, cx, fmaf(-ax, cy, fi[u]));
        }
    }
    for (; b < nb; ++b) {
        const In *xb = x0 + (size_t)b * xbs;
        const In av = __ldg(xb);
        const In cv = has_lag ? __ldg(xb + D) : In{};
        const float ax = (float)av.x, ay = (float)av.y, cx = (float)cv.x, cy = (float)cv.y;
        fe[0] = fmaf(ax, ax, fmaf(ay, ay, fe[0]));
        fr[0] = fmaf(ax, cx, fmaf(ay, cy, fr[0]));
        fi[0] = fmaf(ay, cx, fmaf(-ax, cy, fi[0]));
    }
    qr = ((double)fr[0] + (double)fr[1]) + ((double)fr[2] + (double)fr[3]);
    qi = ((double)fi[0] + (double)fi[1]) + ((double)fi[2] + (double)fi[3]);
    e = ((double)fe[0] + (double)fe[1]) + ((double)fe[2] + (double)fe[3]);
}

template <int DT>
__global__ void __launch_bounds__(TILE_NT, 1) metric_tile_kernel(TileParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *sq = reinterpret_cast<double2 *>(smem_raw);
    double *se = reinterpret_cast<double *>(smem_raw + (size_t)p.NP * sizeof(double2));
    __shared__ double wtot[3][TILE_NT / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t frame = blockIdx.x / p.tiles_per_frame;
    const int tile = blockIdx.x % p.tiles_per_frame;
    const int64_t d0 = (int64_t)tile * p.T;
    const int64_t kbase = d0 + p.offmin;
    constexpr size_t esz = InT<DT>::bytes;
    const unsigned char *xf = reinterpret_cast<const unsigned char *>(p.x) + (size_t)frame * p.xfs * esz;

    // ---- phase 1: branch-summed lag products and energies into smem slots 1..NP-1 ---------------
    if (tid == 0) { sq[0] = make_double2(0.0, 0.0); se[0] = 0.0; }
    for (int s = 1 + tid; s < p.NP; s += TILE_NT) {
        const int64_t j = kbase + s - 1;
        double qr = 0.0, qi = 0.0, e = 0.0;
        if (j >= 0 && j < p.L) {
            const bool has_lag = (j + p.D < p.L);
            phase1_slot<DT>(xf, p.nb, (size_t)p.xbs, j, p.D, has_lag, qr, qi, e);
        }
        sq[s] = make_double2(qr, qi);
        se[s] = e;
    }
    __syncthreads();

    // ---- phase 2: in-place inclusive scan over NP slots ---------------------------------------
    int ipt = (p.NP + TILE_NT - 1) / TILE_NT;
    ipt |= 1;  // odd segment length -> conflict-free strided smem access
    const int s0 = tid * ipt;
    const int s1 = min(s0 + ipt, p.NP);
    double ar = 0.0, ai = 0.0, ae = 0.0;
    for (int s = s0; s < s1; ++s) {
        double2 v = sq[s];
        ar += v.x; ai += v.y; ae += se[s];
        sq[s] = make_double2(ar, ai);
        se[s] = ae;
    }
    // warp inclusive scan of thread totals
    double tr = ar, ti = ai, te = ae;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double yr = shfl_up_f64(tr, o), yi = shfl_up_f64(ti, o), ye = shfl_up_f64(te, o);
        if (lane >= o) { tr += yr; ti += yi; te += ye; }
    }
    if (lane == 31) { wtot[0][warp] = tr; wtot[1][warp] = ti; wtot[2][warp] = te; }
    __syncthreads();
    double br = 0.0, bi = 0.0, be = 0.0;
    for (int w = 0; w < warp; ++w) { br += wtot[0][w]; bi += wtot[1][w]; be += wtot[2][w]; }
    const double offr = br + (tr - ar), offi = bi + (ti - ai), offe = be + (te - ae);  // exclusive offset
    for (int s = s0; s < s1; ++s) {
        double2 v = sq[s];
        sq[s] = make_double2(v.x + offr, v.y + offi);
        se[s] += offe;
    }
    __syncthreads();

    // ---- phase 3: outputs ----------------------------------------------------------------------
    const int64_t dend = min(d0 + (int64_t)p.T, p.out_len);
    const int sh = -p.offmin;
    for (int64_t d = d0 + tid; d < dend; d += TILE_NT) {
        const int r = (int)(d - d0) + sh;
        double2 b0 = sq[r + p.pb0], a0 = sq[r + p.pa0];
        double Pr = b0.x - a0.x, Pi = b0.y - a0.y;
        if (p.npw == 2) {
            double2 b1 = sq[r + p.pb1], a1 = sq[r + p.pa1];
            Pr += b1.x - a1.x; Pi += b1.y - a1.y;
        }
        if (p.conjP) Pi = -Pi;
        const double Rv = se[r + p.rb] - se[r + p.ra];
        double Mv;
        if (p.kind == OFS_AA) {
            // sync_aa.py:486-493: valid (n >= L) and R > 1e-6*L ? min(|P|^2/R^2, 1) : 0
            const bool valid = d >= p.aa_L;
            Mv = 0.0;
            if (valid && Rv > 1e-6 * (double)p.aa_L) {
                Mv = (Pr * Pr + Pi * Pi) / (Rv * Rv);
                Mv = Mv < 1.0 ? Mv : 1.0;
            }
        } else {
            const double rr = Rv > 1e-12 ? Rv : 1e-12;
            if (p.kind == OFS_MINN) {
                const double a = Pr > 0.0 ? Pr : 0.0;          // minn.py:109-111
                Mv = (a * a) / (rr * rr);
            } else {
                Mv = (Pr * Pr + Pi * Pi) / (rr * rr);          // sc.py:76-77
            }
        }
        const int64_t o = frame * p.out_stride + d;
        if (p.out_f64) {
            if (p.M) reinterpret_cast<double *>(p.M)[o] = Mv;
            if (p.P) reinterpret_cast<double2 *>(p.P)[o] = make_double2(Pr, Pi);
            if (p.R) reinterpret_cast<double *>(p.R)[o] = Rv;
        } else {
            if (p.M) reinterpret_cast<float *>(p.M)[o] = (float)Mv;
            if (p.P) reinterpret_cast<float2 *>(p.P)[o] = make_float2((float)Pr, (float)Pi);
            if (p.R) reinterpret_cast<float *>(p.R)[o] = (float)Rv;
        }
    }
}

int launch_metric_tile(const ofs_metric_desc *d, const void *x, void *M, void *P, void *R, cudaStream_t stream)
{
    TileParams p{};
    const int N = d->symbol_len;
    p.x = x; p.M = M; p.P = P; p.R = R;
    p.L = d->n_samples; p.xfs = d->x_frame_stride; p.xbs = d->x_branch_stride;
    p.out_len = ofs_metric_out_len(d); p.out_stride = d->out_stride;
    p.dtype = d->in_dtype; p.nb = d->n_branches; p.kind = d->kind; p.out_f64 = d->out_f64;
    p.npw = 1; p.conjP = 0; p.pa1 = p.pb1 = 0; p.aa_L = 0;
    int offmax;
    switch (d->kind) {
    case OFS_SC:      p.D = N / 2; p.pa0 = 0; p.pb0 = N / 2; p.ra = N / 2; p.rb = N; p.offmin = 0; offmax = N; break;
    case OFS_SC_BOTH: p.D = N / 2; p.pa0 = 0; p.pb0 = N / 2; p.ra = 0; p.rb = N; p.offmin = 0; offmax = N; break;
    case OFS_MINN: {
        const int Q = N / 4;
        p.D = Q; p.pa0 = 0; p.pb0 = Q; p.pa1 = 2 * Q; p.pb1 = 3 * Q; p.npw = 2; p.ra = Q; p.rb = 4 * Q;
        p.offmin = 0; offmax = 4 * Q; break;
    }
    case OFS_AA:
        p.D = N; p.pa0 = -2 * N + 1; p.pb0 = -N + 1; p.ra = -N + 1; p.rb = 1; p.conjP = 1; p.aa_L = N;
        p.offmin = -2 * N + 1; offmax = 1; break;
    default: set_error("ofs_metric: unknown kind %d", d->kind); return OFS_EINVAL;
    }
    if (p.out_len <= 0 || d->n_frames <= 0) return OFS_OK;
    const int span = offmax - p.offmin;
    constexpr int NP_MAX = 8192;
    if (span + 256 > NP_MAX) {
        set_error("ofs_metric(tile): window span %d too large (max %d)", span, NP_MAX - 256);
        return OFS_EUNSUPPORTED;
    }
    int T = NP_MAX - span - 1;
    // keep at least ~2 waves of CTAs when the batch is small
    const int64_t want_tiles = (2LL * sm_count() + d->n_frames - 1) / d->n_frames;
    if (want_tiles > 1) {
        int64_t t2 = (p.out_len + want_tiles - 1) / want_tiles;
        t2 = ((t2 + 255) / 256) * 256;
        if (t2 < 1024) t2 = 1024;
        if (t2 < T) T = (int)t2;
    }
    if (T > p.out_len) T = (int)p.out_len;
    p.T = T; p.NP = T + span + 1;
    p.tiles_per_frame = (int)((p.out_len + T - 1) / T);
    const size_t smem = (size_t)p.NP * 24;
    const int64_t grid = (int64_t)p.tiles_per_frame * d->n_frames;
    OFS_REQUIRE(grid < (1LL << 31), "ofs_metric(tile): grid too large");
#define OFS_TILE_LAUNCH(DT)                                                                                          \
    do {                                                                                                             \
        static PerDeviceOnce once;                                                                                   \
        if (!once.done()) {                                                                                          \
            OFS_CUDA(cudaFuncSetAttribute(metric_tile_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, NP_MAX * 24 + 64)); \
            once.mark();                                                                                             \
        }                                                                                                            \
        metric_tile_kernel<DT><<<(unsigned)grid, TILE_NT, smem, stream>>>(p);                                        \
    } while (0)
    if (d->in_dtype == OFS_C64) OFS_TILE_LAUNCH(OFS_C64);
    else if (d->in_dtype == OFS_C128) OFS_TILE_LAUNCH(OFS_C128);
    else OFS_TILE_LAUNCH(OFS_IQ16);
#undef OFS_TILE_LAUNCH
    return check_launch("metric_tile_kernel");
}

}  // namespace ofs

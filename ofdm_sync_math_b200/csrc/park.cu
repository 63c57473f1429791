// Park metric (K3) -- park.py:64-114.
//   P(d) = sum_{k<h} x[d-k] * x[d+k]   (no conjugate; mirror correlation, NOT a sliding sum)
//   E(d) = sum_{k<h} |x[d+k]|^2,  d in [h, L-h-1],  M = |P|^2 / max(E,1e-12)^2, branches summed first.
// 1024 complex MACs per output: bound by the FMA pipe, not by HBM.  One CTA =
// 512 consecutive outputs of one frame; the 512+2h-1 samples it needs are staged in shared memory
// once; each thread produces PO=4 outputs and walks the lags PK=4 at a time from registers
// (14 smem loads per 16 complex MACs).  Accumulation: T (float for c64/iq16 input with float64
// flushes every 128 lags, double for c128).
#include "common.cuh"

namespace ofs {

constexpr int PNT = 128, PO = 4, PK = 4, PTILE = PNT * PO;

template <typename T> struct Cx { T x, y; };

template <typename T, int DT>
__global__ void __launch_bounds__(PNT) park_kernel(const void *x, int nb, int64_t L, int64_t xfs, int64_t xbs, int h,
                                                   int64_t n_out, int64_t out_stride, int out_f64, void *M, void *P, void *E,
                                                   int tiles_per_frame)
{
    extern __shared__ __align__(16) unsigned char psm[];
    Cx<T> *xs = reinterpret_cast<Cx<T> *>(psm);              // PTILE + 2h + 2PK samples
    const int64_t frame = blockIdx.x / tiles_per_frame;
    const int tile = blockIdx.x % tiles_per_frame;
    const int64_t i0 = (int64_t)tile * PTILE;                // first output index of the tile (d = h + i)
    const int span = PTILE + 2 * h + 2 * PK;
    using In = typename InT<DT>::type;
    const int tid = threadIdx.x;
    double accr[PO], acci[PO], acce[PO];
#pragma unroll
    for (int i = 0; i < PO; ++i) accr[i] = acci[i] = acce[i] = 0.0;

    for (int b = 0; b < nb; ++b) {
        const In *xb = reinterpret_cast<const In *>(x) + frame * xfs + (int64_t)b * xbs;
        // xs[s] holds sample j = jbase + s.  The tile needs samples i0+1 .. i0+PTILE+2h-2; PK-1 guard
        // samples on either side keep the (masked) partial lag groups in bounds.
        const int64_t jbase = i0 - (PK - 1);
        __syncthreads();
        for (int s = tid; s < span; s += PNT) {
            const int64_t j = jbase + s;
            Cx<T> v{(T)0, (T)0};
            if (j >= 0 && j < L) { const In a = xb[j]; v.x = (T)a.x; v.y = (T)a.y; }
            xs[s] = v;
        }
        __syncthreads();
        // thread's outputs: i = i0 + tid*PO + o  -> centre d = h + i ; smem index of x[d] = d - jbase
        const int c0 = h + tid * PO + (PK - 1);              // smem index of x[d] for o = 0
        T pr[PO], pi[PO], pe[PO];
#pragma unroll
        for (int o = 0; o < PO; ++o) pr[o] = pi[o] = pe[o] = (T)0;
        for (int k0 = 0; k0 < h; k0 += PK) {
            Cx<T> a[PO + PK - 1], bb[PO + PK - 1];
#pragma unroll
            for (int q = 0; q < PO + PK - 1; ++q) {
                a[q] = xs[c0 - k0 - (PK - 1) + q];           // x[d_0 - k0 - (PK-1) + q]
                bb[q] = xs[c0 + k0 + q];                      // x[d_0 + k0 + q]
            }
#pragma unroll
            for (int o = 0; o < PO; ++o) {
#pragma unroll
                for (int kk = 0; kk < PK; ++kk) {
                    if (k0 + kk < h) {
                        const Cx<T> u = a[o - kk + PK - 1], v = bb[o + kk];
                        pr[o] = fma(u.x, v.x, pr[o]); pr[o] = fma(-u.y, v.y, pr[o]);
                        pi[o] = fma(u.x, v.y, pi[o]); pi[o] = fma(u.y, v.x, pi[o]);
                        pe[o] = fma(v.x, v.x, pe[o]); pe[o] = fma(v.y, v.y, pe[o]);
                    }
                }
            }
            if (sizeof(T) == 4 && ((k0 + PK) % 128 == 0)) {   // bounded fp32 error: flush partial sums
#pragma unroll
                for (int o = 0; o < PO; ++o) {
                    accr[o] += (double)pr[o]; acci[o] += (double)pi[o]; acce[o] += (double)pe[o];
                    pr[o] = pi[o] = pe[o] = (T)0;
                }
            }
        }
#pragma unroll
        for (int o = 0; o < PO; ++o) { accr[o] += (double)pr[o]; acci[o] += (double)pi[o]; acce[o] += (double)pe[o]; }
    }
#pragma unroll
    for (int o = 0; o < PO; ++o) {
        const int64_t i = i0 + tid * PO + o;
        if (i >= n_out) break;
        const double ee = acce[o] > 1e-12 ? acce[o] : 1e-12;
        const double m = (accr[o] * accr[o] + acci[o] * acci[o]) / (ee * ee);
        const int64_t oi = frame * out_stride + i;
        if (out_f64) {
            if (M) reinterpret_cast<double *>(M)[oi] = m;
            if (P) reinterpret_cast<double2 *>(P)[oi] = make_double2(accr[o], acci[o]);
            if (E) reinterpret_cast<double *>(E)[oi] = acce[o];
        } else {
            if (M) reinterpret_cast<float *>(M)[oi] = (float)m;
            if (P) reinterpret_cast<float2 *>(P)[oi] = make_float2((float)accr[o], (float)acci[o]);
            if (E) reinterpret_cast<float *>(E)[oi] = (float)acce[o];
        }
    }
}

// ---- fast variant (h % 8 == 0): 8 outputs x 8 lags per register group, de-interleaved shared memory ------------------
// Thread t owns outputs 8t .. 8t+7 of the tile, so every sample index it touches is 8t + (compile-time constant) + (a
// multiple of 8).  Samples are staged de-interleaved by 8 (logical i -> sub-array i % 8, slot i / 8): the sub-array is
// then a compile-time choice and the slot is t + const, i.e. consecutive across the warp -- conflict-free LDS, where
// the plain layout had an 8-way bank conflict on every load (32-byte thread stride).  30 loads feed 64 complex MACs
// (30 B per FMA instruction: under the 128 B/clk shared-memory port at the full FMA rate).  The energy E(d) is a sliding
// sum and is no longer recomputed per lag (it was 1/3 of the FMAs).
constexpr int QNT2 = 128, QO = 8, QK = 8, QTILE = QNT2 * QO;

// HC = h as a compile-time constant (0: runtime).  With it the sub-array pitch is a constant and every one of the 30 loads of
// a lag group is base + immediate; with a runtime pitch each load carried its own IMAD / LEA (13 % of the issue slots).
template <typename T, int DT, int HC>
__global__ void __launch_bounds__(QNT2) park_kernel_v2(const void *x, int nb, int64_t L, int64_t xfs, int64_t xbs, int h_rt,
                                                      int64_t n_out, int64_t out_stride, int out_f64, void *M, void *P, void *E,
                                                      int tiles_per_frame)
{
    const int h = HC ? HC : h_rt;
    extern __shared__ __align__(16) unsigned char psm[];
    Cx<T> *xs = reinterpret_cast<Cx<T> *>(psm);
    const int64_t frame = blockIdx.x / tiles_per_frame;
    const int tile = blockIdx.x % tiles_per_frame;
    const int64_t i0 = (int64_t)tile * QTILE;                // first output index of the tile (d = h + i)
    // logical index s <-> sample j = jbase + s, jbase = i0 - 8 (a multiple of 8 below the first sample needed, i0 + 1 - ...)
    const int span = QTILE + 2 * h + 16;                     // multiple of 8
    const int sub = span / 8;                                // slots per sub-array
    using In = typename InT<DT>::type;
    const int tid = threadIdx.x;
    const int64_t jbase = i0 - 8;
    double accr[QO], acci[QO], acce[QO];
#pragma unroll
    for (int i = 0; i < QO; ++i) accr[i] = acci[i] = acce[i] = 0.0;

    for (int b = 0; b < nb; ++b) {
        const In *xb = reinterpret_cast<const In *>(x) + frame * xfs + (int64_t)b * xbs;
        __syncthreads();
        for (int s = tid; s < span; s += QNT2) {
            const int64_t j = jbase + s;
            Cx<T> v{(T)0, (T)0};
            if (j >= 0 && j < L) { const In a = xb[j]; v.x = (T)a.x; v.y = (T)a.y; }
            xs[(s & 7) * sub + (s >> 3)] = v;
        }
        __syncthreads();
        // x[d_o] for output o of this thread sits at logical index c = h + 8 + 8 tid + o  (d = h + i0 + 8 tid + o)
        const int cslot = (h >> 3) + 1 + tid;                // slot of logical index h + 8 + 8 tid (sub-array 0)
        T pr[QO], pi[QO];
#pragma unroll
        for (int o = 0; o < QO; ++o) pr[o] = pi[o] = (T)0;
        for (int k0 = 0; k0 < h; k0 += QK) {
            // a[q] = x[d_0 - k0 - 7 + q], q = 0..14: logical c - k0 - 7 + q = 8 (cslot - k0/8 - 1) + (q + 1)
            // bb[q] = x[d_0 + k0 + q],    q = 0..14: logical c + k0 + q     = 8 (cslot + k0/8) + q
            Cx<T> a[QO + QK - 1], bb[QO + QK - 1];
            const int sa = cslot - (k0 >> 3) - 1, sb = cslot + (k0 >> 3);
#pragma unroll
            for (int q = 0; q < QO + QK - 1; ++q) {
                a[q] = xs[((q + 1) & 7) * sub + sa + ((q + 1) >> 3)];
                bb[q] = xs[(q & 7) * sub + sb + (q >> 3)];
            }
#pragma unroll
            for (int o = 0; o < QO; ++o) {
#pragma unroll
                for (int kk = 0; kk < QK; ++kk) {
                    const Cx<T> u = a[o - kk + QK - 1], v = bb[o + kk];
                    pr[o] = fma(u.x, v.x, pr[o]); pr[o] = fma(-u.y, v.y, pr[o]);
                    pi[o] = fma(u.x, v.y, pi[o]); pi[o] = fma(u.y, v.x, pi[o]);
                }
            }
            if (sizeof(T) == 4 && ((k0 + QK) % 128 == 0)) {   // bounded fp32 error: flush partial sums
#pragma unroll
                for (int o = 0; o < QO; ++o) { accr[o] += (double)pr[o]; acci[o] += (double)pi[o]; pr[o] = pi[o] = (T)0; }
            }
        }
#pragma unroll
        for (int o = 0; o < QO; ++o) { accr[o] += (double)pr[o]; acci[o] += (double)pi[o]; }
        // E(d) = sum_{k<h} |x[d+k]|^2: direct for the thread's first output (float partials flushed every 128 terms), then slid
        {
            double e0 = 0.0;
            for (int k1 = 0; k1 < h; k1 += 128) {
                T part = (T)0;
                const int kend = k1 + 128 < h ? k1 + 128 : h;
                for (int k = k1; k < kend; ++k) {
                    const int li = h + 8 + 8 * tid + k;
                    const Cx<T> v = xs[(li & 7) * sub + (li >> 3)];
                    part = fma(v.x, v.x, part); part = fma(v.y, v.y, part);
                }
                e0 += (double)part;
            }
            acce[0] += e0;
#pragma unroll
            for (int o = 1; o < QO; ++o) {
                const int lo = h + 8 + 8 * tid + o - 1, hi = lo + h;
                const Cx<T> vl = xs[(lo & 7) * sub + (lo >> 3)], vh = xs[(hi & 7) * sub + (hi >> 3)];
                e0 += ((double)vh.x * vh.x + (double)vh.y * vh.y) - ((double)vl.x * vl.x + (double)vl.y * vl.y);
                acce[o] += e0;
            }
        }
    }
#pragma unroll
    for (int o = 0; o < QO; ++o) {
        const int64_t i = i0 + tid * QO + o;
        if (i >= n_out) break;
        const double ee = acce[o] > 1e-12 ? acce[o] : 1e-12;
        const double m = (accr[o] * accr[o] + acci[o] * acci[o]) / (ee * ee);
        const int64_t oi = frame * out_stride + i;
        if (out_f64) {
            if (M) reinterpret_cast<double *>(M)[oi] = m;
            if (P) reinterpret_cast<double2 *>(P)[oi] = make_double2(accr[o], acci[o]);
            if (E) reinterpret_cast<double *>(E)[oi] = acce[o];
        } else {
            if (M) reinterpret_cast<float *>(M)[oi] = (float)m;
            if (P) reinterpret_cast<float2 *>(P)[oi] = make_float2((float)accr[o], (float)acci[o]);
            if (E) reinterpret_cast<float *>(E)[oi] = (float)acce[o];
        }
    }
}

template <typename T, int DT>
static int launch_park(const ofs_metric_desc *d, const void *x, void *M, void *P, void *E, int64_t n_out, cudaStream_t st)
{
    const int h = d->symbol_len / 2;
    if (h % 8 == 0 && h >= 8) {
        const int tiles2 = (int)((n_out + QTILE - 1) / QTILE);
        const size_t smem2 = (size_t)(QTILE + 2 * h + 16) * sizeof(Cx<T>);
        OFS_REQUIRE(smem2 <= 200 * 1024, "ofs_park_metric: symbol_len too large");
        const int64_t grid2 = (int64_t)tiles2 * d->n_frames;
        OFS_REQUIRE(grid2 < (1LL << 31), "ofs_park_metric: grid too large");
        auto go = [&](auto kern2) -> int {
            OFS_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            kern2<<<(unsigned)grid2, QNT2, smem2, st>>>(x, d->n_branches, d->n_samples, d->x_frame_stride, d->x_branch_stride, h, n_out,
                                                      d->out_stride, d->out_f64, M, P, E, tiles2);
            return check_launch("park_kernel_v2");
        };
        return h == 1024 ? go(park_kernel_v2<T, DT, 1024>) : go(park_kernel_v2<T, DT, 0>);   // N_FFT = 2048 is the scripts' geometry
    }
    const int tiles = (int)((n_out + PTILE - 1) / PTILE);
    const size_t smem = (size_t)(PTILE + 2 * h + 2 * PK) * sizeof(Cx<T>);
    auto kern = park_kernel<T, DT>;
    OFS_REQUIRE(smem <= 200 * 1024, "ofs_park_metric: symbol_len too large");
    OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (int64_t)tiles * d->n_frames;
    OFS_REQUIRE(grid < (1LL << 31), "ofs_park_metric: grid too large");
    kern<<<(unsigned)grid, PNT, smem, st>>>(x, d->n_branches, d->n_samples, d->x_frame_stride, d->x_branch_stride, h, n_out,
                                           d->out_stride, d->out_f64, M, P, E, tiles);
    return check_launch("park_kernel");
}

}  // namespace ofs

using namespace ofs;

OFS_API int ofs_park_metric(const ofs_metric_desc *d, const void *x, void *M, void *P, void *E, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(d && x, "ofs_park_metric: null argument");
    OFS_REQUIRE(d->symbol_len >= 2 && d->n_branches >= 1 && d->n_frames >= 0, "ofs_park_metric: bad descriptor");
    const int h = d->symbol_len / 2;
    const int64_t n_out = d->n_samples - 2 * (int64_t)h;          // park.py:79-95: empty if L < 2h+1
    if (n_out <= 0 || d->n_frames == 0) return OFS_OK;
    OFS_REQUIRE(d->out_stride >= n_out, "ofs_park_metric: out_stride too small");
    cudaStream_t st = (cudaStream_t)stream;
    switch (d->in_dtype) {
    case OFS_C64: return launch_park<float, OFS_C64>(d, x, M, P, E, n_out, st);
    case OFS_IQ16: return launch_park<float, OFS_IQ16>(d, x, M, P, E, n_out, st);
    case OFS_C128: return launch_park<double, OFS_C128>(d, x, M, P, E, n_out, st);
    default: set_error("ofs_park_metric: unknown dtype"); return OFS_EINVAL;
    }
}

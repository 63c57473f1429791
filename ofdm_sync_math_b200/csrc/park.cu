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

template <typename T, int DT>
static int launch_park(const ofs_metric_desc *d, const void *x, void *M, void *P, void *E, int64_t n_out, cudaStream_t st)
{
    const int h = d->symbol_len / 2;
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

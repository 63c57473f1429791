// K11: the stage after the detector in every script, batched -- one CTA per frame:
//   apply_cfo(rx, -cfo)  (core.py:123-138)  ->  mean over branches (sc.py:289)  ->  pilot / data symbol FFT + used bins
//   (core.ofdm_fft_used, core.py:171-176)  ->  LS channel estimate H = Y / (X + eps) (core.py:339-341)  ->  phase-slope
//   timing (np.unwrap + line fit, core.py:443-469)  ->  equalise (core.py:344-345)  ->  complex-gain alignment
//   (core.py:357-362)  ->  EVM (core.py:365-370).  Call sites: sc.py:286-309, minn.py:546-568, zc_v2.py:691-720.
// The n_fft-point DFT is read off the shared 4096-point radix-16 FFT (fft4096.cuh) of the zero-padded symbol:
// X_nfft[k] = X_4096[k * 4096 / n_fft].  Everything in float64 (a frame is two FFTs and 1200 bins: latency, not throughput).
#include "fft4096.cuh"

namespace ofs {

struct RxParams {
    const void *x;
    const int64_t *pilot_cp_start;
    const double *cfo_hz;
    const int32_t *bins;            // DFT bin numbers of the used subcarriers, in the order of core.centered_subcarrier_indices
    const double *kidx;             // the same subcarriers as signed indices (float64), for the line fit
    const double2 *pilot_used, *data_used;
    int64_t n_frames, n, xfs, xbs, data_stride;
    int nb, n_fft, cp_len, n_used, dtype;
    double fs;
    const double2 *tw;
    double2 *h_est, *xhat;
    double *scalars;                // [n_frames][8]: evm_rms, evm_db, slope, sto, gain_re, gain_im, cpe(unused 0), valid
};

__device__ __forceinline__ double2 cdiv(double2 a, double2 b)
{
    const double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

__device__ double block_sum(double v, double *sh)
{
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_f64(v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < ZNT / 32; ++w) t += sh[w];
    return t;
}

// position of DFT bin m (0..4095) after fft_dif: m = k0 + 16 k1 + 256 k2 sits at 256 k0 + 16 k1 + k2
__device__ __forceinline__ int fft_pos(int m) { return zpad(((m & 15) << 8) | (m & 0xf0) | (m >> 8)); }

__global__ void __launch_bounds__(ZNT) rx_chain_kernel(const RxParams p)
{
    extern __shared__ __align__(16) unsigned char rsm[];
    double2 *a = reinterpret_cast<double2 *>(rsm);                 // ZFP
    double2 *yp = a + ZFP;                                         // n_used: pilot bins, then H
    double *phi = reinterpret_cast<double *>(yp + p.n_used);       // n_used
    __shared__ double sh[ZNT / 32];
    const int64_t f = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t start = p.pilot_cp_start[f];
    const int N = p.n_fft, stride = ZF / N;
    const size_t esz = dtype_bytes_dev(p.dtype);
    const unsigned char *xf = reinterpret_cast<const unsigned char *>(p.x) + (size_t)f * p.xfs * esz;
    const double cfo = p.cfo_hz ? p.cfo_hz[f] : 0.0;
    double *sc = p.scalars + f * 8;
    const int64_t data_start = start + p.cp_len + N;               // data_cp_start (sc.py:303)
    // numpy slices that run past the end are silently shortened and np.fft.fft(td, n=N) zero-pads them (the AWGN scenario of
    // sc.py ends 12 samples inside the data symbol): samples beyond the capture count as zeros.  Only a start outside the
    // capture is rejected.
    if (start < 0 || start >= p.n) {
        if (tid == 0) { for (int q = 0; q < 8; ++q) sc[q] = 0.0; sc[0] = sc[1] = __longlong_as_double(0x7ff8000000000000LL); }
        return;
    }
    for (int sym = 0; sym < 2; ++sym) {
        const int64_t s0 = (sym == 0 ? start : data_start) + p.cp_len;
        // CFO-corrected branch mean of the symbol body, zero-padded to 4096
        for (int m = tid; m < ZF; m += ZNT) {
            double2 v = make_double2(0.0, 0.0);
            const int64_t j = s0 + m;
            if (m < N && j < p.n) {
                double sn, cs;
                sincos(2.0 * 3.14159265358979323846 * (-cfo) * (double)j / p.fs, &sn, &cs);
                double ar = 0.0, ai = 0.0;
                for (int b = 0; b < p.nb; ++b) {
                    const double2 s = load_sample_f64(xf + (size_t)b * p.xbs * esz, p.dtype, j);
                    ar += s.x * cs - s.y * sn; ai += s.x * sn + s.y * cs;
                }
                v = make_double2(ar / (double)p.nb, ai / (double)p.nb);
            }
            a[zpad(m)] = v;
        }
        __syncthreads();
        fft_dif<double2>(a, p.tw);
        if (sym == 0) {
            // LS estimate (core.py:341) and its phase
            for (int u = tid; u < p.n_used; u += ZNT) {
                const double2 y = a[fft_pos(p.bins[u] * stride)];
                const double2 xk = p.pilot_used[u];
                const double2 h = cdiv(y, make_double2(xk.x + 1e-9, xk.y));
                yp[u] = h;
                phi[u] = atan2(h.y, h.x);
                p.h_est[f * p.n_used + u] = h;
            }
            __syncthreads();
            // np.unwrap, then the line fit of core.py:457-468.  The correction of sample u depends on the ORIGINAL phases u-1, u
            // only (numpy: dd = diff(p); ph_correct = f(dd); up[1:] = p[1:] + cumsum(ph_correct)), so the fmod-heavy part runs on
            // all threads; what stays sequential is numpy's left-to-right cumulative sum, two additions per bin.
            double *cadd = reinterpret_cast<double *>(a);          // the FFT buffer is free: every used bin sits in yp / phi by now
            {
                const double PI = 3.14159265358979323846;
                for (int u = 1 + tid; u < p.n_used; u += ZNT) {
                    const double dd = phi[u] - phi[u - 1];
                    double ddm = fmod(dd + PI, 2.0 * PI);
                    if (ddm < 0.0) ddm += 2.0 * PI;                 // numpy's mod is non-negative
                    ddm -= PI;
                    if (ddm == -PI && dd > 0.0) ddm = PI;
                    double c = ddm - dd;
                    if (fabs(dd) < PI) c = 0.0;
                    cadd[u] = c;
                }
            }
            __syncthreads();
            if (tid == 0) {
                double corr = 0.0;
                for (int u = 1; u < p.n_used; ++u) { corr += cadd[u]; phi[u] += corr; }
            }
            __syncthreads();
            double sk = 0.0, sp = 0.0;
            for (int u = tid; u < p.n_used; u += ZNT) { sk += p.kidx[u]; sp += phi[u]; }
            const double km = block_sum(sk, sh) / (double)p.n_used;
            const double pm = block_sum(sp, sh) / (double)p.n_used;
            double skk = 0.0, skp = 0.0;
            for (int u = tid; u < p.n_used; u += ZNT) {
                const double kz = p.kidx[u] - km;
                skk += kz * kz; skp += kz * (phi[u] - pm);
            }
            const double den = block_sum(skk, sh) + 1e-12;
            const double num = block_sum(skp, sh);
            if (tid == 0) { sc[2] = num / den; sc[3] = -(num / den) * (double)N / (2.0 * 3.14159265358979323846); }
        } else {
            // equalise (core.py:345), align (core.py:357-362), EVM (core.py:365-370)
            const double2 *dref = p.data_used + f * p.data_stride;
            double nr = 0.0, ni = 0.0, dn = 0.0;
            for (int u = tid; u < p.n_used; u += ZNT) {
                const double2 y = a[fft_pos(p.bins[u] * stride)];
                const double2 h = yp[u];
                const double2 xh = cdiv(y, make_double2(h.x + 1e-9, h.y));
                const double2 r = dref[u];
                nr += xh.x * r.x + xh.y * r.y;      // vdot(x, ref) = sum conj(x) ref
                ni += xh.x * r.y - xh.y * r.x;
                dn += xh.x * xh.x + xh.y * xh.y;
                yp[u] = xh;
            }
            const double gnr = block_sum(nr, sh), gni = block_sum(ni, sh), gdn = block_sum(dn, sh) + 1e-12;
            const double gr = gnr / gdn, gi = gni / gdn;
            double e2 = 0.0, r2 = 0.0;
            for (int u = tid; u < p.n_used; u += ZNT) {
                const double2 xh = yp[u], r = dref[u];
                const double2 xa = make_double2(xh.x * gr - xh.y * gi, xh.x * gi + xh.y * gr);
                p.xhat[f * p.n_used + u] = xa;
                const double er = xa.x - r.x, ei = xa.y - r.y;
                e2 += er * er + ei * ei; r2 += r.x * r.x + r.y * r.y;
            }
            const double se = block_sum(e2, sh), sr = block_sum(r2, sh);
            if (tid == 0) {
                const double evm = sqrt((se / (double)p.n_used) / (sr / (double)p.n_used));
                sc[0] = evm; sc[1] = 20.0 * log10(evm + 1e-12); sc[4] = gr; sc[5] = gi; sc[6] = 0.0; sc[7] = 1.0;
            }
        }
        __syncthreads();
    }
}

__global__ void zc_twiddle_kernel(double2 *tw);

}  // namespace ofs

using namespace ofs;

OFS_API int ofs_rx_chain(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n, int64_t x_frame_stride,
                         int64_t x_branch_stride, const int64_t *pilot_cp_start, const double *cfo_hz, double fs, int32_t n_fft,
                         int32_t cp_len, const int32_t *bins, const double *k_index, int32_t n_used, const void *pilot_used_c128,
                         const void *data_used_c128, int64_t data_stride, void *h_est_c128, void *xhat_c128, double *scalars,
                         void *stream_)
{
    OFS_TRACE();
    OFS_REQUIRE(x && pilot_cp_start && bins && k_index && pilot_used_c128 && data_used_c128 && h_est_c128 && xhat_c128 && scalars,
                "ofs_rx_chain: null argument");
    OFS_REQUIRE(in_dtype >= OFS_C64 && in_dtype <= OFS_IQ16, "ofs_rx_chain: unknown dtype");
    OFS_REQUIRE(n_fft >= 16 && n_fft <= ZF && (ZF % n_fft) == 0, "ofs_rx_chain: n_fft must divide 4096");
    OFS_REQUIRE(n_used >= 1 && n_used <= n_fft && n_branches >= 1 && cp_len >= 0 && n_frames >= 0, "ofs_rx_chain: bad geometry");
    if (n_frames == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    keep_pool_cached();
    double2 *tw = nullptr;
    OFS_CUDA(cudaMallocAsync((void **)&tw, (ZF / 2) * (sizeof(double2) + sizeof(float2)), stream));
    zc_twiddle_kernel<<<(ZF / 2 + 255) / 256, 256, 0, stream>>>(tw);
    if (int rc = check_launch("zc_twiddle_kernel")) return rc;
    RxParams p{};
    p.x = x; p.pilot_cp_start = pilot_cp_start; p.cfo_hz = cfo_hz; p.bins = bins; p.kidx = k_index;
    p.pilot_used = (const double2 *)pilot_used_c128; p.data_used = (const double2 *)data_used_c128; p.n_frames = n_frames; p.n = n;
    p.xfs = x_frame_stride; p.xbs = x_branch_stride; p.data_stride = data_stride; p.nb = n_branches; p.n_fft = n_fft; p.cp_len = cp_len;
    p.n_used = n_used; p.dtype = in_dtype; p.fs = fs; p.tw = tw; p.h_est = (double2 *)h_est_c128; p.xhat = (double2 *)xhat_c128;
    p.scalars = scalars;
    const size_t smem = (size_t)ZFP * sizeof(double2) + (size_t)n_used * (sizeof(double2) + sizeof(double));
    OFS_CUDA(cudaFuncSetAttribute(rx_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rx_chain_kernel<<<(unsigned)n_frames, ZNT, smem, stream>>>(p);
    if (int rc = check_launch("rx_chain_kernel")) return rc;
    OFS_CUDA(cudaFreeAsync(tw, stream));
    return OFS_OK;
}

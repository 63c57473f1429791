// Reference-order streaming kernels: the reference's scalar recurrences evaluated in exactly the
// reference's operation order (no FMA contraction: __dmul_rn / __dadd_rn / __dsub_rn), one
// thread per independent stream.  These exist because some answers of the reference depend on the
// rounding of its running sums (the docs/detector_test_vector.csv tie 1523/1524,
// the float smoother of minn_rtl.py:709-715, the integer floor-shift smoother of
// ref/minn_preamble_detector.sv:294-296).  Throughput comes from the number of concurrent
// streams, not from parallelism inside one.
//
//   aa_reference_kernel      sync_aa.py:321-386 + 458-493   (RunningSum / RunningSumReal / DelayLine)
//   rtl_window_kernel<T>     minn_rtl.py:512-652            (_DelayLine / _RunningSum / _antenna_path)
//                            ref/minn_antenna_path.sv:63-194, ref/minn_running_sum.sv:77-82
//   rtl_combine_*            minn_rtl.py:695-733            ref/minn_preamble_detector.sv:247-325
#include "common.cuh"
#include <cstdlib>
#include <cmath>
#include <type_traits>

namespace ofs {

// ------------------------------------------------------------------------------------------ sync_aa
__global__ void aa_reference_kernel(const void *x, int dtype, int64_t n_frames, int na, int64_t n, int L,
                                    double2 *P, double *R, double *M, uint8_t *valid)
{
    const int64_t frame = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (frame >= n_frames) return;
    const size_t esz = dtype == OFS_C64 ? 8 : (dtype == OFS_C128 ? 16 : 4);
    const unsigned char *xf = reinterpret_cast<const unsigned char *>(x) + (size_t)frame * na * n * esz;
    // per-antenna running sums live in the output rows while streaming: P/R rows of this frame are
    // written once per sample with the antenna TOTAL, so the per-antenna state is kept in registers
    // for na <= 8 and re-derived from memory otherwise (loop below keeps it simple: na <= 64, state
    // in local memory).
    double psr[64], psi[64], rs[64];
    for (int a = 0; a < na; ++a) { psr[a] = 0.0; psi[a] = 0.0; rs[a] = 0.0; }
    const double floor_ = 1e-6 * (double)L;
    for (int64_t t = 0; t < n; ++t) {
        double Pr = 0.0, Pi = 0.0, Rs = 0.0;
        for (int a = 0; a < na; ++a) {
            const void *xa = xf + (size_t)a * n * esz;
            const double2 xn = load_sample_f64(xa, dtype, t);
            // product = x[n] * conj(x[n-L]) once the delay line is full (sync_aa.py:470)
            double qr = 0.0, qi = 0.0;
            if (t >= L) {
                const double2 xd = load_sample_f64(xa, dtype, t - L);
                qr = __dadd_rn(__dmul_rn(xn.x, xd.x), __dmul_rn(xn.y, xd.y));
                qi = __dsub_rn(__dmul_rn(xn.y, xd.x), __dmul_rn(xn.x, xd.y));
            }
            // oldest product in the L-deep window: the one pushed at t-L (zero before the window filled)
            double or_ = 0.0, oi = 0.0;
            if (t >= 2 * (int64_t)L) {
                const double2 a1 = load_sample_f64(xa, dtype, t - L);
                const double2 a2 = load_sample_f64(xa, dtype, t - 2 * (int64_t)L);
                or_ = __dadd_rn(__dmul_rn(a1.x, a2.x), __dmul_rn(a1.y, a2.y));
                oi = __dsub_rn(__dmul_rn(a1.y, a2.x), __dmul_rn(a1.x, a2.y));
            }
            psr[a] = __dsub_rn(__dadd_rn(psr[a], qr), or_);       // sum + sample - oldest (sync_aa.py:337)
            psi[a] = __dsub_rn(__dadd_rn(psi[a], qi), oi);
            const double pw = __dadd_rn(__dmul_rn(xn.x, xn.x), __dmul_rn(xn.y, xn.y));
            double po = 0.0;
            if (t >= L) {
                const double2 xd = load_sample_f64(xa, dtype, t - L);
                po = __dadd_rn(__dmul_rn(xd.x, xd.x), __dmul_rn(xd.y, xd.y));
            }
            rs[a] = __dsub_rn(__dadd_rn(rs[a], pw), po);
            Pr = __dadd_rn(Pr, psr[a]); Pi = __dadd_rn(Pi, psi[a]); Rs = __dadd_rn(Rs, rs[a]);
        }
        const bool v = t >= L;
        const int64_t o = frame * n + t;
        P[o] = make_double2(Pr, Pi); R[o] = Rs; valid[o] = v;
        double m = 0.0;
        if (v && Rs > floor_) {
            m = __ddiv_rn(__dadd_rn(__dmul_rn(Pr, Pr), __dmul_rn(Pi, Pi)), __dmul_rn(Rs, Rs));
            m = m < 1.0 ? m : 1.0;
        }
        M[o] = m;
    }
}

// Same results bit for bit, one CTA per frame: only the three running sums of each antenna are a true recurrence
// (s = (s + new) - oldest, two dependent float64 adds per step); the lag products feeding them and the antenna totals
// leaving them are computed by all threads from coalesced loads, 128 time steps at a time, antennas in groups of 6 (the
// totals are accumulated in antenna order, as sync_aa.py:478-479 does).  137 ms -> ~2 ms for one 262144-sample capture.
constexpr int AAK = 128, AAG = 6;
__global__ void __launch_bounds__(AAK) aa_reference_kernel_v2(const void *x, int dtype, int na, int64_t n, int L,
                                                              double2 *P, double *R, double *M, uint8_t *valid)
{
    __shared__ double tot[3][AAK];
    __shared__ double buf[AAG][6][AAK];           // qr, qi, or, oi, pw, po -> overwritten by psr, psi, -, -, rs, -
    __shared__ double state[64][3];
    const int64_t frame = blockIdx.x;
    const int tid = threadIdx.x;
    const size_t esz = dtype == OFS_C64 ? 8 : (dtype == OFS_C128 ? 16 : 4);
    const unsigned char *xf = reinterpret_cast<const unsigned char *>(x) + (size_t)frame * na * n * esz;
    if (tid < 64) { state[tid][0] = 0.0; state[tid][1] = 0.0; state[tid][2] = 0.0; }
    const double floor_ = 1e-6 * (double)L;
    for (int64_t t0 = 0; t0 < n; t0 += AAK) {
        tot[0][tid] = 0.0; tot[1][tid] = 0.0; tot[2][tid] = 0.0;
        const int kmax = (int)(n - t0 < AAK ? n - t0 : AAK);
        for (int a0 = 0; a0 < na; a0 += AAG) {
            __syncthreads();
            for (int idx = tid; idx < AAG * AAK; idx += AAK) {
                const int al = idx / AAK, k = idx % AAK;
                const int a = a0 + al;
                const int64_t t = t0 + k;
                if (a >= na || t >= n) continue;
                const void *xa = xf + (size_t)a * n * esz;
                const double2 xn = load_sample_f64(xa, dtype, t);
                double qr = 0.0, qi = 0.0, orr = 0.0, oi = 0.0, po = 0.0;
                if (t >= L) {
                    const double2 xd = load_sample_f64(xa, dtype, t - L);
                    qr = __dadd_rn(__dmul_rn(xn.x, xd.x), __dmul_rn(xn.y, xd.y));          // x[n] * conj(x[n-L]) (sync_aa.py:470)
                    qi = __dsub_rn(__dmul_rn(xn.y, xd.x), __dmul_rn(xn.x, xd.y));
                    po = __dadd_rn(__dmul_rn(xd.x, xd.x), __dmul_rn(xd.y, xd.y));
                    if (t >= 2 * (int64_t)L) {
                        const double2 a2 = load_sample_f64(xa, dtype, t - 2 * (int64_t)L);
                        orr = __dadd_rn(__dmul_rn(xd.x, a2.x), __dmul_rn(xd.y, a2.y));
                        oi = __dsub_rn(__dmul_rn(xd.y, a2.x), __dmul_rn(xd.x, a2.y));
                    }
                }
                buf[al][0][k] = qr; buf[al][1][k] = qi; buf[al][2][k] = orr; buf[al][3][k] = oi;
                buf[al][4][k] = __dadd_rn(__dmul_rn(xn.x, xn.x), __dmul_rn(xn.y, xn.y));
                buf[al][5][k] = po;
            }
            __syncthreads();
            if (tid < AAG && a0 + tid < na) {
                double psr = state[a0 + tid][0], psi = state[a0 + tid][1], rs = state[a0 + tid][2];
                for (int k = 0; k < kmax; ++k) {
                    psr = __dsub_rn(__dadd_rn(psr, buf[tid][0][k]), buf[tid][2][k]);       // sum + sample - oldest (sync_aa.py:337)
                    psi = __dsub_rn(__dadd_rn(psi, buf[tid][1][k]), buf[tid][3][k]);
                    rs = __dsub_rn(__dadd_rn(rs, buf[tid][4][k]), buf[tid][5][k]);
                    buf[tid][0][k] = psr; buf[tid][1][k] = psi; buf[tid][4][k] = rs;
                }
                state[a0 + tid][0] = psr; state[a0 + tid][1] = psi; state[a0 + tid][2] = rs;
            }
            __syncthreads();
            if (tid < kmax) {
                double Pr = tot[0][tid], Pi = tot[1][tid], Rs = tot[2][tid];
                for (int al = 0; al < AAG && a0 + al < na; ++al) {
                    Pr = __dadd_rn(Pr, buf[al][0][tid]); Pi = __dadd_rn(Pi, buf[al][1][tid]); Rs = __dadd_rn(Rs, buf[al][4][tid]);
                }
                tot[0][tid] = Pr; tot[1][tid] = Pi; tot[2][tid] = Rs;
            }
        }
        if (tid < kmax) {
            const int64_t t = t0 + tid;
            const double Pr = tot[0][tid], Pi = tot[1][tid], Rs = tot[2][tid];
            const bool v = t >= L;
            const int64_t o = frame * n + t;
            P[o] = make_double2(Pr, Pi); R[o] = Rs; valid[o] = v;
            double m = 0.0;
            if (v && Rs > floor_) {
                m = __ddiv_rn(__dadd_rn(__dmul_rn(Pr, Pr), __dmul_rn(Pi, Pi)), __dmul_rn(Rs, Rs));
                m = m < 1.0 ? m : 1.0;
            }
            M[o] = m;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------ minn_rtl
// Per (frame, antenna): running Q-window sums of the lag product and of the power.
//   C[n] = sum_{k=n-Q+1..n} prod[k],  prod[k] = re*re_d + im*im_d with the delayed sample x[k-D]
//   E[n] = sum_{k=n-Q+1..n} |x[k]|^2
// T = double: float mirror (operation order of minn_rtl._RunningSum.step); T = long long: integer RTL.
template <typename T, int DT>
__global__ void rtl_window_kernel(const void *x, int64_t n_streams, int64_t n, int Q, int D, T *C, T *E)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    using In = typename InT<DT>::type;
    const In *xs = reinterpret_cast<const In *>(x) + s * n;
    T *Cs = C + s * n, *Es = E + s * n;
    T c = 0, e = 0;
    auto prod_at = [&](int64_t k) -> T {
        if (k < D) return (T)0;
        const In a = xs[k], b = xs[k - D];
        if constexpr (sizeof(T) == 8 && !std::is_floating_point<T>::value)
            return (T)a.x * (T)b.x + (T)a.y * (T)b.y;
        else
            return (T)__dadd_rn(__dmul_rn((double)b.x, (double)a.x), __dmul_rn((double)b.y, (double)a.y));
    };
    auto pow_at = [&](int64_t k) -> T {
        const In a = xs[k];
        if constexpr (sizeof(T) == 8 && !std::is_floating_point<T>::value)
            return (T)a.x * (T)a.x + (T)a.y * (T)a.y;
        else
            return (T)__dadd_rn(__dmul_rn((double)a.x, (double)a.x), __dmul_rn((double)a.y, (double)a.y));
    };
    for (int64_t i = 0; i < n; ++i) {
        const T p = prod_at(i), w = pow_at(i);
        const T po = i >= Q ? prod_at(i - Q) : (T)0;
        const T wo = i >= Q ? pow_at(i - Q) : (T)0;
        if constexpr (std::is_floating_point<T>::value) {
            c = (T)__dsub_rn(__dadd_rn((double)c, (double)p), (double)po);     // sum_reg + val - oldest
            e = (T)__dsub_rn(__dadd_rn((double)e, (double)w), (double)wo);
        } else {
            c = c + p - po;
            e = e + w - wo;
        }
        Cs[i] = c; Es[i] = e;
    }
}

// Float mirror, few long streams: one CTA per stream.  The products are computed by all threads from coalesced loads, 256
// steps at a time; only c = (c + p) - p_old and e = (e + w) - w_old run on one thread each (same operation order as
// minn_rtl._RunningSum.step, so the results stay bit-equal to the reference).
constexpr int RWK = 256;
template <int DT>
__global__ void __launch_bounds__(RWK) rtl_window_kernel_v2(const void *x, int64_t n, int Q, int D, double *C, double *E)
{
    __shared__ double bp[RWK], bo[RWK], bw[RWK], bwo[RWK];
    __shared__ double st[2];
    using In = typename InT<DT>::type;
    const int64_t s = blockIdx.x;
    const In *xs = reinterpret_cast<const In *>(x) + s * n;
    double *Cs = C + s * n, *Es = E + s * n;
    const int tid = threadIdx.x;
    auto prod_at = [&](int64_t k) -> double {
        if (k < D) return 0.0;
        const In a = xs[k], b = xs[k - D];
        return __dadd_rn(__dmul_rn((double)b.x, (double)a.x), __dmul_rn((double)b.y, (double)a.y));
    };
    auto pow_at = [&](int64_t k) -> double {
        const In a = xs[k];
        return __dadd_rn(__dmul_rn((double)a.x, (double)a.x), __dmul_rn((double)a.y, (double)a.y));
    };
    if (tid < 2) st[tid] = 0.0;
    for (int64_t i0 = 0; i0 < n; i0 += RWK) {
        const int64_t i = i0 + tid;
        __syncthreads();
        if (i < n) {
            bp[tid] = prod_at(i); bw[tid] = pow_at(i);
            bo[tid] = i >= Q ? prod_at(i - Q) : 0.0;
            bwo[tid] = i >= Q ? pow_at(i - Q) : 0.0;
        }
        __syncthreads();
        const int kmax = (int)(n - i0 < RWK ? n - i0 : RWK);
        if (tid == 0) {
            double c = st[0];
            for (int k = 0; k < kmax; ++k) { c = __dsub_rn(__dadd_rn(c, bp[k]), bo[k]); bp[k] = c; }     // sum_reg + val - oldest
            st[0] = c;
        } else if (tid == 32) {
            double e = st[1];
            for (int k = 0; k < kmax; ++k) { e = __dsub_rn(__dadd_rn(e, bw[k]), bwo[k]); bw[k] = e; }
            st[1] = e;
        }
        __syncthreads();
        if (i < n) { Cs[i] = bp[tid]; Es[i] = bw[tid]; }
    }
}

// Integer window sums are exact in any order: a tile-parallel version for the int16 datapath (throughput path of the
// RTL model).  One CTA = RT outputs of one (frame, antenna) stream: products / powers from coalesced loads, int64
// inclusive scan in shared memory, C[n] = S[n+1] - S[n-Q+1].
constexpr int RT = 4096, RNT = 256;
__global__ void __launch_bounds__(RNT) rtl_window_tile_kernel(const short2 *x, int64_t n, int Q, int D, long long *C, long long *E)
{
    extern __shared__ long long rsm[];
    long long *sp = rsm;                   // RT + Q + 1
    long long *sw = rsm + (RT + Q + 1);
    __shared__ long long wtot[2][RNT / 32];
    const int64_t s = blockIdx.y;
    const int64_t n0 = (int64_t)blockIdx.x * RT;
    const short2 *xs = x + s * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t j0 = n0 - Q + 1;
    if (j0 < 0) j0 = 0;
    const int64_t nend = n0 + RT < n ? n0 + RT : n;
    const int cnt = (int)(nend - j0);
    for (int k = tid; k < cnt; k += RNT) {
        const int64_t i = j0 + k;
        const short2 a = xs[i];
        long long p = 0;
        if (i >= D) { const short2 b = xs[i - D]; p = (long long)a.x * b.x + (long long)a.y * b.y; }
        sp[k + 1] = p;
        sw[k + 1] = (long long)a.x * a.x + (long long)a.y * a.y;
    }
    if (tid == 0) { sp[0] = 0; sw[0] = 0; }
    __syncthreads();
    int ipt = (cnt + RNT - 1) / RNT;
    ipt |= 1;
    const int s0 = 1 + tid * ipt, s1 = min(s0 + ipt, cnt + 1);
    long long ap = 0, aw = 0;
    for (int q = s0; q < s1; ++q) { ap += sp[q]; sp[q] = ap; aw += sw[q]; sw[q] = aw; }
    long long tp = ap, tw = aw;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long yp = __shfl_up_sync(0xffffffffu, tp, o), yw = __shfl_up_sync(0xffffffffu, tw, o);
        if (lane >= o) { tp += yp; tw += yw; }
    }
    if (lane == 31) { wtot[0][warp] = tp; wtot[1][warp] = tw; }
    __syncthreads();
    long long op = tp - ap, ow = tw - aw;
    for (int w = 0; w < warp; ++w) { op += wtot[0][w]; ow += wtot[1][w]; }
    for (int q = s0; q < s1; ++q) { sp[q] += op; sw[q] += ow; }
    __syncthreads();
    for (int64_t i = n0 + tid; i < nend; i += RNT) {
        int64_t lo = i - Q + 1;
        if (lo < 0) lo = 0;
        C[s * n + i] = sp[i + 1 - j0] - sp[lo - j0];
        E[s * n + i] = sw[i + 1 - j0] - sw[lo - j0];
    }
}

// Window sums AND antenna combining of the integer datapath in one pass (ref/minn_antenna_path.sv:63-194 summed over the
// antennas as in ref/minn_preamble_detector.sv:247-275).  Integer addition is exact in any order, and the hold-register
// gating of the combiner depends on the index only, so the per-sample lag products and powers are summed over the antennas
// FIRST and a single int64 prefix per tile gives every window:
//   corr_total[i]   = [i >= Q-1] (Sp[i+1] - Sp[i-Q+1]) + [i >= 2Q-1] (Sp[i-Q+1] - Sp[i-2Q+1])
//   energy_total[i] = the same two terms on the powers + [i >= 3Q-1] (Sw[i-2Q+1] - Sw[i-3Q+1])          (lower ends clamped at 0)
// One CTA = RT outputs of one frame; it reads RT + 3Q - 1 samples of every antenna (int16 IQ, 4 B) and writes 25 B per output;
// the per-antenna window arrays (16 B per antenna-sample written, then read again by a combine kernel) no longer exist.
__global__ void __launch_bounds__(RNT) rtl_int_fused_kernel(const short2 *x, int nb, int64_t n, int Q, int D, long long *corr_total,
                                                            long long *corr_positive, long long *energy_total, uint8_t *valid)
{
    extern __shared__ long long rsm[];
    const int halo = 3 * Q - 1;
    long long *sp = rsm;                   // RT + halo + 1
    long long *sw = rsm + (RT + halo + 1);
    __shared__ long long wtot[2][RNT / 32];
    const int64_t frame = blockIdx.y;
    const int64_t n0 = (int64_t)blockIdx.x * RT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t j0 = n0 - halo;
    if (j0 < 0) j0 = 0;
    const int64_t nend = n0 + RT < n ? n0 + RT : n;
    const int cnt = (int)(nend - j0);
    for (int k = tid; k < cnt; k += RNT) {
        const int64_t i = j0 + k;
        long long p = 0, w = 0;
        for (int b = 0; b < nb; ++b) {
            const short2 *xs = x + (frame * nb + b) * n;
            const short2 a = xs[i];
            if (i >= D) { const short2 c = xs[i - D]; p += (long long)a.x * c.x + (long long)a.y * c.y; }
            w += (long long)a.x * a.x + (long long)a.y * a.y;
        }
        sp[k + 1] = p;
        sw[k + 1] = w;
    }
    if (tid == 0) { sp[0] = 0; sw[0] = 0; }
    __syncthreads();
    int ipt = (cnt + RNT - 1) / RNT;
    ipt |= 1;
    const int s0 = 1 + tid * ipt, s1 = min(s0 + ipt, cnt + 1);
    long long ap = 0, aw = 0;
    for (int q = s0; q < s1; ++q) { ap += sp[q]; sp[q] = ap; aw += sw[q]; sw[q] = aw; }
    long long tp = ap, tw = aw;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long yp = __shfl_up_sync(0xffffffffu, tp, o), yw = __shfl_up_sync(0xffffffffu, tw, o);
        if (lane >= o) { tp += yp; tw += yw; }
    }
    if (lane == 31) { wtot[0][warp] = tp; wtot[1][warp] = tw; }
    __syncthreads();
    long long op = tp - ap, ow = tw - aw;
    for (int w = 0; w < warp; ++w) { op += wtot[0][w]; ow += wtot[1][w]; }
    for (int q = s0; q < s1; ++q) { sp[q] += op; sw[q] += ow; }
    __syncthreads();
    // S(k) = sum of the samples below index k (k clamped at 0; everything below j0 is outside every window that is used)
    auto P = [&](int64_t k) { return sp[(k < j0 ? j0 : k) - j0]; };
    auto W = [&](int64_t k) { return sw[(k < j0 ? j0 : k) - j0]; };
    for (int64_t i = n0 + tid; i < nend; i += RNT) {
        long long ct = 0, et = 0;
        if (i >= Q - 1) { ct += P(i + 1) - P(i - Q + 1); et += W(i + 1) - W(i - Q + 1); }
        if (i >= 2 * (int64_t)Q - 1) { ct += P(i - Q + 1) - P(i - 2 * (int64_t)Q + 1); et += W(i - Q + 1) - W(i - 2 * (int64_t)Q + 1); }
        if (i >= 3 * (int64_t)Q - 1) et += W(i - 2 * (int64_t)Q + 1) - W(i - 3 * (int64_t)Q + 1);
        const int64_t o = frame * n + i;
        corr_total[o] = ct;
        corr_positive[o] = ct > 0 ? ct : 0;
        energy_total[o] = et;
        valid[o] = i >= 3 * (int64_t)Q - 1;
    }
}

// Combine antennas with the hold-register gating (minn_rtl.py:632-650 / minn_antenna_path.sv:168-194):
//   corr_recent = C[n] (n >= Q-1), corr_previous = C[n-Q] (n >= 2Q-1),
//   energy_recent = E[n] (n >= Q-1), energy_previous = E[n-Q] (n >= 2Q-1), energy_previous2 = E[n-2Q] (n >= 3Q-1)
template <typename T>
__global__ void rtl_combine_kernel(const T *C, const T *E, int64_t n_frames, int nb, int64_t n, int Q,
                                   T *corr_total, T *corr_positive, T *energy_total, uint8_t *valid)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_frames * n) return;
    const int64_t frame = idx / n, i = idx % n;
    T ct = 0, et = 0;
    for (int b = 0; b < nb; ++b) {
        const T *Cs = C + (frame * nb + b) * n, *Es = E + (frame * nb + b) * n;
        const T cr = i >= Q - 1 ? Cs[i] : (T)0;
        const T cp = i >= 2 * (int64_t)Q - 1 ? Cs[i - Q] : (T)0;
        const T er = i >= Q - 1 ? Es[i] : (T)0;
        const T ep = i >= 2 * (int64_t)Q - 1 ? Es[i - Q] : (T)0;
        const T e2 = i >= 3 * (int64_t)Q - 1 ? Es[i - 2 * (int64_t)Q] : (T)0;
        if constexpr (std::is_floating_point<T>::value) {
            ct = __dadd_rn(ct, __dadd_rn(cr, cp));                              // minn_rtl.py:696
            et = __dadd_rn(et, __dadd_rn(__dadd_rn(er, ep), e2));              // :697-701
        } else {
            ct += cr + cp; et += er + ep + e2;
        }
    }
    corr_total[idx] = ct;
    corr_positive[idx] = ct > (T)0 ? ct : (T)0;
    energy_total[idx] = et;
    valid[idx] = i >= 3 * (int64_t)Q - 1;
}

// Smoother + threshold (minn_rtl.py:707-722 / minn_preamble_detector.sv:277-325).  The shift-IIR is a true recurrence
// (float: rounding order matters; integer: floor shift is non-linear), so it is evaluated sequentially -- but by ONE WARP
// PER FRAME: the warp loads 32 consecutive samples coalesced and parks them in shared memory, lane 0 runs the 32
// dependent steps out of registers (all its loads issued up front), and the warp writes the 32 results back coalesced.
constexpr int SMW = 4;     // warps (frames) per CTA
template <typename T>
__global__ void __launch_bounds__(SMW * 32) rtl_smooth_kernel(const T *corr_positive, const T *energy_total, const uint8_t *valid,
                                                              int64_t n_frames, int64_t n, int shift, T thr_value, int frac_bits,
                                                              T *smooth, T *corr_scaled, T *energy_scaled, uint8_t *above,
                                                              const int *only_if = nullptr)
{
    __shared__ T sc[SMW][32];
    __shared__ T ss[SMW][32];
    __shared__ uint8_t sv[SMW][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t frame = (int64_t)blockIdx.x * SMW + w;
    if (frame >= n_frames) return;
    if (only_if && only_if[frame] == 0) return;          // fallback pass behind the parallel smoother: flagged frames only
    T s = (T)0;
    const double inv_denom = 1.0 / (double)(1LL << (shift > 0 ? shift : 0));     // exact power of two
    const double cscale = (double)(1LL << frac_bits);
    // the loads of chunk k+1 are issued before the 32 dependent steps of chunk k: without this every chunk paid a full
    // memory round trip (8 ms for 2048 x 32768; the recurrence itself is ~0.2 ms)
    auto fetch = [&](int64_t i0, T &c, T &e, uint8_t &v) {
        const int64_t i = i0 + lane;
        const bool in = i < n;
        const int64_t o = frame * n + (in ? i : 0);
        c = in ? corr_positive[o] : (T)0;
        e = in ? energy_total[o] : (T)0;
        v = in ? valid[o] : (uint8_t)0;
    };
    T cn, en; uint8_t vn;
    fetch(0, cn, en, vn);
    for (int64_t i0 = 0; i0 < n; i0 += 32) {
        const int64_t i = i0 + lane;
        const bool in = i < n;
        const int64_t o = frame * n + (in ? i : 0);
        const T c = cn, e = en;
        const uint8_t v = vn;
        if (i0 + 32 < n) fetch(i0 + 32, cn, en, vn);
        sc[w][lane] = c; sv[w][lane] = v;
        __syncwarp();
        if (lane == 0) {
            T cc[32]; uint8_t vv[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) { cc[q] = sc[w][q]; vv[q] = sv[w][q]; }
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                if (vv[q]) {
                    if constexpr (std::is_floating_point<T>::value) {
                        // smooth_val += (corr_positive - smooth_val) / denom   (minn_rtl.py:713)
                        s = shift == 0 ? cc[q] : (T)__dadd_rn((double)s, __dmul_rn(__dsub_rn((double)cc[q], (double)s), inv_denom));
                    } else {
                        s = shift == 0 ? cc[q] : s + ((cc[q] - s) >> shift);        // arithmetic shift = floor (sv:294-296)
                    }
                }
                ss[w][q] = s;
            }
        }
        __syncwarp();
        if (in) {
            const T sm = ss[w][lane];
            smooth[o] = sm;
            bool ab;
            if constexpr (std::is_floating_point<T>::value) {
                const double cs = __dmul_rn((double)sm, cscale);
                const double es = thr_value == (T)0 ? 0.0 : __dmul_rn((double)e, (double)thr_value);
                corr_scaled[o] = (T)cs; energy_scaled[o] = (T)es;
                ab = v && cs >= es;
            } else {
                ab = v && ((sm << frac_bits) >= e * thr_value);
            }
            above[o] = ab;
        }
        __syncwarp();
    }
}


// ---- the integer smoother in parallel ---------------------------------------------------------------------------------
// s <- s + floor((c - s) / 2^k) (minn_preamble_detector.sv:277-296) is a true recurrence, and one lane per frame walking
// 32768 samples was 1.55 ms of the 3.9 ms integer path.  Two facts make it parallel and still bit-exact:
//  (1) without the floor it is the LINEAR filter y <- y (1 - 2^-k) + c 2^-k, which a scan evaluates anywhere (float64: c < 2^53
//      and both coefficients are exact, so y carries ~1e-3 absolute error), and the floor can only pull the integer state
//      below it by less than 2^k:  y - (2^k - 1) <= s <= y  (induction on the step);
//  (2) the step is monotone in s, so two trajectories started at floor(y) - 2^k - 1 and ceil(y) + 1 bracket the true state
//      for ever, and the bracket closes by one whenever (c - s) mod 2^k allows it: once the two coincide they ARE the state.
// A thread runs one chain: W warm-up samples with both trajectories (W = 64 for k <= 3: 99.8 % of the chains have met by
// then), then its own PSS outputs.  A chain that has not met when its segment starts leaves the segment to the nearest chain
// on its left, which simply keeps going; only if that is the first chain of a tile is the frame handed to the serial kernel
// (dirty flag).  Bit-exact in every case.  The tile of corr_positive lives in shared memory and is overwritten in place by the
// state (block-synchronous steps: a chain's segment is written after every chain to its right has read it as warm-up).
constexpr int PSN = 128, PSS = 64, PST = PSN * PSS;
constexpr int PS_MAXH = 1024;                                             // longest halo (linear scan history / warm-up)
__device__ __forceinline__ int cpad(int k) { return k + 2 * (k >> 6); }  // a chain walks 64 consecutive values: thread stride 66 keeps the
                                                                         // segments 16-byte aligned for the bulk copies (2-way conflicts)
__device__ __forceinline__ long long shift_iir_step(long long s, long long c, int shift)
{
    return s + ((c - s) >> shift);                   // arithmetic shift = floor; shift 0: s = c
}
__global__ void __launch_bounds__(PSN, 3) rtl_smooth_par_kernel(const long long *corr_positive, int64_t n, int64_t first_valid, int shift,
                                                                int H, int W, long long *smooth, int *dirty, int tiles_per_frame)
{
    extern __shared__ __align__(16) long long cs[];  // cs[cpad(k)]: corr_positive at position t0 - H + k, later the smoothed state
    __shared__ double segA[PSN + PS_MAXH / PSS], segB[PSN + PS_MAXH / PSS], ystart[PSN + PS_MAXH / PSS + 1];
    __shared__ unsigned char unres[PSN];
    __shared__ __align__(8) uint64_t bar;
    const int64_t frame = blockIdx.y, t0 = (int64_t)blockIdx.x * PST;
    const int tid = threadIdx.x;
    const long long *cp = corr_positive + frame * n;
    const int hs = H / PSS, nseg = hs + PSN;
    // ---- the tile + halo arrive as one 512-byte bulk copy per 64-sample segment (cp.async.bulk on one mbarrier): every load of
    //      the CTA is in flight at once; segments that stick out of the frame (or a frame that is not 16-byte aligned) take plain loads
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(cp) & 15) == 0);
    auto seg_full = [&](int g) { const int64_t p0 = t0 - H + (int64_t)g * PSS; return p0 >= 0 && p0 + PSS <= n; };
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        int full = 0;
        if (bulk_ok) for (int g = 0; g < nseg; ++g) full += seg_full(g) ? 1 : 0;
        mbar_expect_tx(&bar, (uint32_t)(full * PSS * 8));
    }
    __syncthreads();
    for (int g = tid; g < nseg; g += PSN) {
        const int64_t p0 = t0 - H + (int64_t)g * PSS;
        long long *dst = cs + cpad(g * PSS);
        if (bulk_ok && seg_full(g)) tma_load_1d(dst, cp + p0, PSS * 8, &bar);
        else
            for (int m = 0; m < PSS; ++m) { const int64_t j = p0 + m; dst[m] = (j >= 0 && j < n) ? cp[j] : 0; }
    }
    mbar_wait(&bar, 0);
    __syncthreads();
    // (1) linear filter per 64-sample segment: y_end = A y_start + B (segments 0 .. hs-1 are the halo)
    const double b = 1.0 / (double)(1LL << shift), a = 1.0 - b;
    for (int g = tid; g < nseg; g += PSN) {
        const long long *c = cs + cpad(g * PSS);
        const int64_t pos0 = t0 - H + (int64_t)g * PSS;
        double A = 1.0, B = 0.0;
        if (pos0 >= first_valid && pos0 + PSS <= n) {
#pragma unroll 16
            for (int m = 0; m < PSS; ++m) B = fma(B, a, (double)c[m] * b);
            A = pow(a, (double)PSS);
        } else {
            for (int m = 0; m < PSS; ++m) {
                const int64_t pos = pos0 + m;
                if (pos >= first_valid && pos < n) { B = fma(B, a, (double)c[m] * b); A *= a; }
            }
        }
        segA[g] = A; segB[g] = B;
    }
    __syncthreads();
    if (tid == 0) {                                  // the state before the halo is forgotten by its end (H is chosen that way)
        double y = 0.0;
        for (int g = 0; g < nseg; ++g) { ystart[g] = y; y = fma(y, segA[g], segB[g]); }
    }
    __syncthreads();
    // (2) the chain of segment hs + tid starts W samples (W / 64 segments) early, bracketing the state around the linear value
    const int64_t seg = t0 + (int64_t)tid * PSS;     // first output of this chain
    const int nblk = W / PSS;
    long long lo, hi;
    if (seg - W <= first_valid) { lo = hi = 0; }     // nothing has updated the state yet: exactly 0
    else {
        const double y = ystart[hs + tid - nblk];
        lo = (long long)floor(y) - (1LL << shift) - 1;
        hi = (long long)ceil(y) + 1;
        if (lo < 0) lo = 0;                          // corr_positive >= 0 keeps the state >= 0; without this a bracket that starts
                                                     // below zero sticks at -(2^k - 1) through a run of zeros (a fixed point) and never closes
    }
    bool resolved = false;
    for (int blk = 0; blk <= nblk; ++blk) {
        const bool own = blk == nblk;
        if (own) { resolved = lo == hi; unres[tid] = resolved ? 0 : 1; }
        long long *c = cs + cpad(H - W + tid * PSS + blk * PSS);        // (a multiple of 64: the 64 values behind it are contiguous)
        const int64_t pos0 = seg - W + (int64_t)blk * PSS;
        const bool all_valid = pos0 >= first_valid && pos0 + PSS <= n;
        if (!own) {
            if (all_valid && lo != hi) {
#pragma unroll 16
                for (int m = 0; m < PSS; ++m) { lo = shift_iir_step(lo, c[m], shift); hi = shift_iir_step(hi, c[m], shift); }
            } else if (all_valid) {
#pragma unroll 16
                for (int m = 0; m < PSS; ++m) lo = shift_iir_step(lo, c[m], shift);
                hi = lo;
            } else {
                for (int m = 0; m < PSS; ++m) {
                    const int64_t pos = pos0 + m;
                    if (pos >= first_valid && pos < n) { lo = shift_iir_step(lo, c[m], shift); hi = shift_iir_step(hi, c[m], shift); }
                }
            }
        } else if (resolved) {
            if (all_valid) {
#pragma unroll 16
                for (int m = 0; m < PSS; ++m) { lo = shift_iir_step(lo, c[m], shift); c[m] = lo; }
            } else {
                for (int m = 0; m < PSS; ++m) {
                    const int64_t pos = pos0 + m;
                    if (pos >= first_valid && pos < n) lo = shift_iir_step(lo, c[m], shift);
                    c[m] = lo;
                }
            }
        }
        __syncthreads();
    }
    // segments whose chain did not converge in time: the nearest converged chain on the left keeps going through them; a run
    // of them at the START of the tile has nobody on its left here -- its length goes to the fix-up kernel
    if (resolved) {
        for (int nx = tid + 1; nx < PSN && unres[nx]; ++nx) {
            long long *c = cs + cpad(H + nx * PSS);
            const int64_t pos0 = t0 + (int64_t)nx * PSS;
            for (int m = 0; m < PSS; ++m) {
                const int64_t pos = pos0 + m;
                if (pos >= first_valid && pos < n) lo = shift_iir_step(lo, c[m], shift);
                c[m] = lo;
            }
        }
    } else if (tid == 0) {
        int cnt = 0;
        while (cnt < PSN && unres[cnt]) ++cnt;
        dirty[frame * tiles_per_frame + blockIdx.x] = cnt;
    }
    fence_proxy_async_smem();
    __syncthreads();
    // ---- results leave the same way: one bulk store per segment
    long long *sr = smooth + frame * n;
    const bool bulk_out = ((reinterpret_cast<uintptr_t>(sr) & 15) == 0);
    if (seg < n) {
        const long long *src = cs + cpad(H + tid * PSS);
        if (bulk_out && seg + PSS <= n) {
            tma_store_1d(sr + seg, src, PSS * 8);
            tma_store_commit();
            tma_store_wait_read<0>();
        } else {
            for (int m = 0; m < PSS && seg + m < n; ++m) sr[seg + m] = src[m];
        }
    }
}

// Leading segments of a tile whose chains had not converged (rtl_smooth_par_kernel's dirty counts): redone serially from the
// last state of the previous tile, tiles in order, one warp per frame (the warp loads 32 samples at a time, every lane walks
// the 32 steps and keeps its own).  Normally there is nothing to do.
__global__ void __launch_bounds__(128) rtl_smooth_fixup_kernel(const long long *corr_positive, int64_t n_frames, int64_t n, int64_t first_valid,
                                                              int shift, long long *smooth, const int *dirty, int tiles_per_frame)
{
    const int lane = threadIdx.x & 31;
    const int64_t frame = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (frame >= n_frames) return;
    for (int tile = 0; tile < tiles_per_frame; ++tile) {
        const int cnt = dirty[frame * tiles_per_frame + tile];
        if (cnt == 0) continue;
        const int64_t t0 = (int64_t)tile * PST;
        const int64_t t1 = t0 + (int64_t)cnt * PSS < n ? t0 + (int64_t)cnt * PSS : n;
        long long s = t0 > 0 ? smooth[frame * n + t0 - 1] : 0;
        for (int64_t i0 = t0; i0 < t1; i0 += 32) {
            const int64_t i = i0 + lane;
            const long long c = i < t1 ? corr_positive[frame * n + i] : 0;
            long long mine = s;
            for (int q = 0; q < 32; ++q) {
                const long long cq = __shfl_sync(0xffffffffu, c, q);
                if (i0 + q >= first_valid && i0 + q < t1) s = shift_iir_step(s, cq, shift);
                if (q == lane) mine = s;
            }
            if (i < t1) smooth[frame * n + i] = mine;
        }
        __threadfence_block();
    }
}

// above = valid && (smooth << frac_bits) >= energy_total * threshold  (minn_preamble_detector.sv:305-325), elementwise
__global__ void __launch_bounds__(256) rtl_above_kernel(const long long *smooth, const long long *energy_total, int64_t n_frames, int64_t n,
                                                        int64_t first_valid, long long thr_value, int frac_bits, uint8_t *above)
{
    const int64_t total = n_frames * n;
    for (int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; idx < total; idx += (int64_t)gridDim.x * blockDim.x * 4) {
        long long sm[4], en[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int64_t k = idx + q < total ? idx + q : total - 1; sm[q] = smooth[k]; en[q] = energy_total[k]; }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (idx + q < total) above[idx + q] = ((idx + q) % n >= first_valid) && ((sm[q] << frac_bits) >= en[q] * thr_value);
    }
}

}  // namespace ofs

using namespace ofs;

// Reference-order [A][A] metric (float64 outputs).  Declared here, used by the sync_aa shim.
OFS_API int ofs_aa_metric_reference(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_antennas, int64_t n,
                                    int32_t L, void *P_c128, double *R, double *M, uint8_t *valid, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(x && P_c128 && R && M && valid, "ofs_aa_metric_reference: null argument");
    OFS_REQUIRE(n_antennas >= 1 && n_antennas <= 64, "ofs_aa_metric_reference: 1..64 antennas");
    OFS_REQUIRE(L > 0 && n >= 0 && n_frames >= 0, "ofs_aa_metric_reference: bad geometry");
    if (n_frames == 0 || n == 0) return OFS_OK;
    // few long frames: one CTA per frame (recurrence only where it is one); many short frames: one thread per frame
    if (n_frames < 4096 && n_frames < (1LL << 31)) {
        aa_reference_kernel_v2<<<(unsigned)n_frames, AAK, 0, (cudaStream_t)stream>>>(x, in_dtype, n_antennas, n, L, (double2 *)P_c128, R, M, valid);
        return check_launch("aa_reference_kernel_v2");
    }
    const int bs = 32;
    aa_reference_kernel<<<(unsigned)((n_frames + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
        x, in_dtype, n_frames, n_antennas, n, L, (double2 *)P_c128, R, M, valid);
    return check_launch("aa_reference_kernel");
}

OFS_API int ofs_minn_rtl_metric(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n,
                                int32_t quarter_len, int32_t smooth_shift, int32_t threshold_value, int32_t frac_bits,
                                double *corr_total, double *corr_positive, double *smooth_metric, double *energy_total,
                                double *corr_scaled, double *energy_scaled, uint8_t *metric_valid, uint8_t *above,
                                void *stream_)
{
    OFS_TRACE();
    OFS_REQUIRE(x && corr_total && corr_positive && smooth_metric && energy_total && corr_scaled && energy_scaled &&
                    metric_valid && above, "ofs_minn_rtl_metric: null argument");
    OFS_REQUIRE(in_dtype == OFS_C64 || in_dtype == OFS_C128, "ofs_minn_rtl_metric: complex64/complex128 input");
    OFS_REQUIRE(quarter_len > 0, "quarter_len must be positive.");
    OFS_REQUIRE(n_branches >= 1 && n >= 0 && n_frames >= 0 && frac_bits >= 0 && frac_bits < 62, "ofs_minn_rtl_metric: bad geometry");
    if (n_frames == 0 || n == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    keep_pool_cached();
    const int64_t ns = n_frames * n_branches;
    double *C = nullptr, *E = nullptr;
    OFS_CUDA(cudaMallocAsync((void **)&C, (size_t)ns * n * sizeof(double), stream));
    OFS_CUDA(cudaMallocAsync((void **)&E, (size_t)ns * n * sizeof(double), stream));
    const int bs = 32;
    if (in_dtype == OFS_C64) {
        if (ns < 8192) rtl_window_kernel_v2<OFS_C64><<<(unsigned)ns, RWK, 0, stream>>>(x, n, quarter_len, quarter_len, C, E);
        else rtl_window_kernel<double, OFS_C64><<<(unsigned)((ns + bs - 1) / bs), bs, 0, stream>>>(x, ns, n, quarter_len, quarter_len, C, E);
    } else {
        if (ns < 8192) rtl_window_kernel_v2<OFS_C128><<<(unsigned)ns, RWK, 0, stream>>>(x, n, quarter_len, quarter_len, C, E);
        else rtl_window_kernel<double, OFS_C128><<<(unsigned)((ns + bs - 1) / bs), bs, 0, stream>>>(x, ns, n, quarter_len, quarter_len, C, E);
    }
    if (int rc = check_launch("rtl_window_kernel")) return rc;
    const int64_t tot = n_frames * n;
    rtl_combine_kernel<double><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(C, E, n_frames, n_branches, n, quarter_len,
                                                                               corr_total, corr_positive, energy_total, metric_valid);
    if (int rc = check_launch("rtl_combine_kernel")) return rc;
    rtl_smooth_kernel<double><<<(unsigned)((n_frames + SMW - 1) / SMW), SMW * 32, 0, stream>>>(
        corr_positive, energy_total, metric_valid, n_frames, n, smooth_shift, (double)threshold_value, frac_bits, smooth_metric,
        corr_scaled, energy_scaled, above);
    if (int rc = check_launch("rtl_smooth_kernel")) return rc;
    OFS_CUDA(cudaFreeAsync(C, stream));
    OFS_CUDA(cudaFreeAsync(E, stream));
    return OFS_OK;
}

// Integer smoother + threshold: the parallel kernel (shift <= 5), then the serial kernel for the frames it
// flagged (normally none); the serial kernel alone otherwise.  Bit-exact either way.
static int launch_int_smoother(const long long *corr_positive, const long long *energy_total, const uint8_t *metric_valid,
                               int64_t n_frames, int n_branches, int64_t n, int Q, int shift, int threshold_value, int frac_bits,
                               long long *smooth, uint8_t *above, cudaStream_t stream)
{
    // upper bound of any state: the state never exceeds the largest corr_positive seen, and that is a sum of
    // n_branches * 2Q products of two int16 pairs (|re re + im im| <= 2^31)
    double bound = (double)n_branches * 2.0 * (double)Q * 2147483648.0;
    int bits = 1;
    while (bits < 62 && (double)(1LL << bits) <= bound) ++bits;
    const int K = 1 << (shift > 0 ? shift : 0);
    // H: history after which the linear filter has forgotten its start to 1/4 LSB: (1 - 1/K)^H 2^bits < 1/4
    int H = shift > 0 ? (int)((double)(bits + 2) * 0.6931472 / -log1p(-1.0 / (double)K)) + 1 : 1;
    H = (H + PSS - 1) / PSS * PSS;
    // W: integer warm-up from a bracket of 2^k + 3 (simulated: < 0.5 % of the chains still apart, those hand over to their neighbour)
    static const int wb[6] = {1, 1, 1, 1, 3, 5};
    const int W = shift >= 0 && shift <= 5 ? wb[shift] * PSS : 0;
    const char *force = getenv("OFS_RTL_SERIAL_SMOOTHER");
    if (shift >= 0 && shift <= 5 && bits <= 50 && H <= PS_MAXH && W <= H && n_frames < 65536 && !(force && force[0] == '1')) {
        int *dirty = nullptr;
        keep_pool_cached();
        const int tiles = (int)((n + PST - 1) / PST);
        OFS_CUDA(cudaMallocAsync((void **)&dirty, (size_t)n_frames * tiles * sizeof(int), stream));
        OFS_CUDA(cudaMemsetAsync(dirty, 0, (size_t)n_frames * tiles * sizeof(int), stream));
        const size_t sm = (size_t)(H + PST + 2 * ((H + PST) / 64) + 4) * sizeof(long long);
        static PerDeviceOnce once;
        if (!once.done()) {
            OFS_CUDA(cudaFuncSetAttribute(rtl_smooth_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)((PS_MAXH + PST + 2 * ((PS_MAXH + PST) / 64) + 4) * sizeof(long long))));
            once.mark();
        }
        const int64_t first_valid = 3 * (int64_t)Q - 1;
        rtl_smooth_par_kernel<<<dim3((unsigned)tiles, (unsigned)n_frames), PSN, sm, stream>>>(corr_positive, n, first_valid, shift, H, W, smooth,
                                                                                           dirty, tiles);
        if (int rc = check_launch("rtl_smooth_par_kernel")) return rc;
        rtl_smooth_fixup_kernel<<<(unsigned)((n_frames + 3) / 4), 128, 0, stream>>>(corr_positive, n_frames, n, first_valid, shift, smooth, dirty,
                                                                                 tiles);
        if (int rc = check_launch("rtl_smooth_fixup_kernel")) return rc;
        const int64_t total = n_frames * n;
        int64_t ab_grid = (total / 4 + 255) / 256;
        if (ab_grid > (int64_t)sm_count() * 16) ab_grid = (int64_t)sm_count() * 16;
        if (ab_grid < 1) ab_grid = 1;
        rtl_above_kernel<<<(unsigned)ab_grid, 256, 0, stream>>>(smooth, energy_total, n_frames, n, first_valid, (long long)threshold_value,
                                                               frac_bits, above);
        if (int rc = check_launch("rtl_above_kernel")) return rc;
        OFS_CUDA(cudaFreeAsync(dirty, stream));
        return OFS_OK;
    }
    rtl_smooth_kernel<long long><<<(unsigned)((n_frames + SMW - 1) / SMW), SMW * 32, 0, stream>>>(
        corr_positive, energy_total, metric_valid, n_frames, n, shift, (long long)threshold_value, frac_bits, smooth, nullptr, nullptr, above);
    return check_launch("rtl_smooth_kernel");
}

OFS_API int ofs_minn_rtl_int(const int16_t *iq, int64_t n_frames, int32_t n_branches, int64_t n, int32_t quarter_len,
                             int32_t smooth_shift, int32_t threshold_value, int32_t frac_bits, int32_t lag_extra,
                             int64_t *corr_total, int64_t *corr_positive, int64_t *smooth_metric, int64_t *energy_total,
                             uint8_t *metric_valid, uint8_t *above, void *stream_)
{
    OFS_TRACE();
    OFS_REQUIRE(iq && corr_total && corr_positive && smooth_metric && energy_total && metric_valid && above,
                "ofs_minn_rtl_int: null argument");
    OFS_REQUIRE(quarter_len > 0 && (lag_extra == 0 || lag_extra == 1), "ofs_minn_rtl_int: bad quarter_len / lag_extra");
    OFS_REQUIRE(n_branches >= 1 && n >= 0 && n_frames >= 0 && frac_bits >= 0 && frac_bits < 24, "ofs_minn_rtl_int: bad geometry");
    if (n_frames == 0 || n == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t fsmem = (size_t)2 * (RT + 3 * (size_t)quarter_len) * sizeof(long long);
    if (fsmem <= 200 * 1024 && n_frames < 65536) {          // window sums + antenna combining fused (Q <= 2133)
        OFS_CUDA(cudaFuncSetAttribute(rtl_int_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        rtl_int_fused_kernel<<<dim3((unsigned)((n + RT - 1) / RT), (unsigned)n_frames), RNT, fsmem, stream>>>(
            reinterpret_cast<const short2 *>(iq), n_branches, n, quarter_len, quarter_len + lag_extra, (long long *)corr_total,
            (long long *)corr_positive, (long long *)energy_total, metric_valid);
        if (int rc = check_launch("rtl_int_fused_kernel")) return rc;
        return launch_int_smoother((const long long *)corr_positive, (const long long *)energy_total, metric_valid, n_frames, n_branches, n,
                                   quarter_len, smooth_shift, threshold_value, frac_bits, (long long *)smooth_metric, above, stream);
    }
    keep_pool_cached();
    const int64_t ns = n_frames * n_branches;
    long long *C = nullptr, *E = nullptr;
    OFS_CUDA(cudaMallocAsync((void **)&C, (size_t)ns * n * sizeof(long long), stream));
    OFS_CUDA(cudaMallocAsync((void **)&E, (size_t)ns * n * sizeof(long long), stream));
    const int bs = 32;
    if (quarter_len <= 8192 && ns < 65536) {
        const size_t rsmem = (size_t)2 * (RT + quarter_len + 1) * sizeof(long long);
        OFS_CUDA(cudaFuncSetAttribute(rtl_window_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        rtl_window_tile_kernel<<<dim3((unsigned)((n + RT - 1) / RT), (unsigned)ns), RNT, rsmem, stream>>>(
            reinterpret_cast<const short2 *>(iq), n, quarter_len, quarter_len + lag_extra, C, E);
    } else {
        rtl_window_kernel<long long, OFS_IQ16><<<(unsigned)((ns + bs - 1) / bs), bs, 0, stream>>>(iq, ns, n, quarter_len,
                                                                                               quarter_len + lag_extra, C, E);
    }
    if (int rc = check_launch("rtl_window_kernel")) return rc;
    const int64_t tot = n_frames * n;
    rtl_combine_kernel<long long><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(
        C, E, n_frames, n_branches, n, quarter_len, (long long *)corr_total, (long long *)corr_positive,
        (long long *)energy_total, metric_valid);
    if (int rc = check_launch("rtl_combine_kernel")) return rc;
    if (int rc = launch_int_smoother((const long long *)corr_positive, (const long long *)energy_total, metric_valid, n_frames, n_branches, n,
                                     quarter_len, smooth_shift, threshold_value, frac_bits, (long long *)smooth_metric, above, stream))
        return rc;
    OFS_CUDA(cudaFreeAsync(C, stream));
    OFS_CUDA(cudaFreeAsync(E, stream));
    return OFS_OK;
}

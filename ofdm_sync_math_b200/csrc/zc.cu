// Zadoff-Chu kernels.
//
// K4  zc_mf_kernel      matched filter by FFT overlap-save, hand-written 4096-point FFT: three passes of radix-16
//                       butterflies in registers with two shared-memory transposes (forward -> pointwise multiply in
//                       digit-reversed order -> inverse, so no reordering pass), fused with the sliding-energy
//                       normalisation.
//                       Replaces np.convolve(x, conj(ref[::-1])) and np.convolve(|x|^2, ones) of
//                       zc.py:115-126 and zc_v2.py:244-271, 486-495.
// K5  zc_freq_kernel    zc_freq.compute_frequency_metric (zc_freq.py:62-99) as a sliding DFT of the used
//                       bins: bins_j(o) = conj(w_j^s) (T_j[s+N]-T_j[s]), T_j = prefix sum of x[m] w_j^m,
//                       instead of one 2048-point FFT per candidate offset.
#include "common.cuh"
#include "fft4096.cuh"
#include "conv8k.cuh"
#include <cstdlib>

namespace ofs {

__global__ void zc_twiddle_kernel(double2 *tw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ZF / 2) {
        double s, c;
        sincospi(-2.0 * (double)i / (double)ZF, &s, &c);
        tw[i] = make_double2(c, s);
        reinterpret_cast<float2 *>(tw + ZF / 2)[i] = make_float2((float)c, (float)s);
    }
}

// G = DIF-FFT of g[m] = conj(ref[nr-1-m]) zero-padded to ZF (bit-reversed order), plus ||ref||.
__global__ void __launch_bounds__(ZNT) zc_spectrum_kernel(const double2 *ref, int nr, const double2 *tw, double2 *G,
                                                          double *ref_norm)
{
    extern __shared__ __align__(16) unsigned char zsm[];
    double2 *a = reinterpret_cast<double2 *>(zsm);
    __shared__ double red[ZNT / 32];
    double e = 0.0;
    for (int m = threadIdx.x; m < ZF; m += ZNT) {
        double2 v = make_double2(0.0, 0.0);
        if (m < nr) { const double2 r = ref[nr - 1 - m]; v = make_double2(r.x, -r.y); e += r.x * r.x + r.y * r.y; }
        a[zpad(m)] = v;
    }
    for (int o = 16; o > 0; o >>= 1) e += shfl_xor_f64(e, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
    __syncthreads();
    fft_dif<double2>(a, tw);
    for (int m = threadIdx.x; m < ZF; m += ZNT) G[m] = a[zpad(m)];
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < ZNT / 32; ++w) t += red[w];
        *ref_norm = sqrt(t);
    }
}

template <typename T, int DT>
__global__ void __launch_bounds__(ZNT, (sizeof(T) == 4 ? 3 : 1)) zc_mf_kernel(const void *x, int nb, int64_t n, int nr, const double2 *tw,
                                                    const double2 *G, const double *ref_norm_p, int mode, int out_f64,
                                                    void *corr_out, void *mag_out, int64_t out_stride, int blocks_per_frame)
{
    using C2 = typename C2T<T>::type;
    using In = typename InT<DT>::type;
    extern __shared__ __align__(16) unsigned char zsm[];
    // a: ZF complex | se: ZF+4 energy-prefix entries | pw: ZF window energies | acc: ZF branch-summed outputs
    // (float path: all fp32 -> 72 KB, 3 CTAs/SM; double path: 144 KB)
    C2 *a = reinterpret_cast<C2 *>(zsm);
    T *se = reinterpret_cast<T *>(zsm + (size_t)ZFP * sizeof(C2));
    T *pw = se + ZFP + 8;
    C2 *acc = reinterpret_cast<C2 *>(pw + ZF);
    __shared__ double wtot[ZNT / 32];

    const int V = ZF - nr + 1;                        // valid outputs per block
    const int64_t frame = blockIdx.x / blocks_per_frame;
    const int blk = blockIdx.x % blocks_per_frame;
    const int64_t k0 = (int64_t)blk * V;              // first full-convolution output of this block
    const int64_t jb = k0 - (nr - 1);                 // first input sample of this block
    const int64_t out_len = n + nr - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double ref_norm = *ref_norm_p;

    // one branch: results leave straight from the inverse FFT, pw / acc are neither allocated nor touched (51 KB -> more CTAs per SM)
    const bool single = nb == 1;
    if (!single)
        for (int i = tid; i < V; i += ZNT) { pw[i] = (T)0; acc[i].x = (T)0; acc[i].y = (T)0; }

    for (int b = 0; b < nb; ++b) {
        const In *xb = reinterpret_cast<const In *>(x) + (frame * nb + b) * n;
        __syncthreads();
        for (int m = tid; m < ZF; m += ZNT) {
            const int64_t j = jb + m;
            C2 v; v.x = 0; v.y = 0;
            if (j >= 0 && j < n) { const In s = xb[j]; v.x = (T)s.x; v.y = (T)s.y; }
            a[zpad(m)] = v;
        }
        if (tid == 0) se[0] = (T)0;                      // se is padded like a (index i at i + i/16): the 16-element thread stride
                                                         // of the serial prefix below would otherwise be a 16-way bank conflict
        __syncthreads();
        // energy prefix se[i+1] = sum_{m<=i} |a[m]|^2: thread-serial + warp scan + CTA carry (carries in float64)
        {
            constexpr int IPT = ZF / ZNT;          // 16
            const int s0 = tid * IPT;
            T run = (T)0;
            for (int m = 0; m < IPT; ++m) {
                const C2 v = a[zpad(s0 + m)];
                run += v.x * v.x + v.y * v.y;
                se[zpad(s0 + m + 1)] = run;
            }
            double t = (double)run;
            for (int o = 1; o < 32; o <<= 1) { const double y = shfl_up_f64(t, o); if (lane >= o) t += y; }
            if (lane == 31) wtot[warp] = t;
            __syncthreads();
            double off = t - (double)run;
            for (int w = 0; w < warp; ++w) off += wtot[w];
            for (int m = 0; m < IPT; ++m) se[zpad(s0 + m + 1)] = (T)((double)se[zpad(s0 + m + 1)] + off);
        }
        __syncthreads();
        fft_dif<C2>(a, tw);
        for (int m = tid; m < ZF; m += ZNT) {
            const double2 g = __ldg(G + m);
            C2 gg; gg.x = (T)g.x; gg.y = (T)g.y;
            a[zpad(m)] = cmul(a[zpad(m)], gg);
        }
        __syncthreads();
        ifft_dit<C2>(a, tw);
        for (int i = tid; i < V; i += ZNT) {
            // output k0+i is the window of local samples [i, i+nr-1]
            const T e = se[zpad(i + nr)] - se[zpad(i)];
            const C2 y = a[zpad(nr - 1 + i)];
            T yr = y.x * (T)(1.0 / ZF), yi = y.y * (T)(1.0 / ZF);
            if (mode == 1) {                                   // zc_v2.py:257-271: per-branch normalisation
                const T d = (T)ref_norm * sqrt(e > (T)1e-12 ? e : (T)1e-12);
                yr /= d; yi /= d;
            }
            if (single) {
                const int64_t k = k0 + i;
                if (k >= out_len) break;
                double wr = (double)yr, wi = (double)yi;
                if (mode == 0) {                               // zc.py:125-126
                    const double pp = e > (T)0 ? (double)e : 0.0;
                    const double d = ref_norm * sqrt(pp + 1e-12);
                    wr /= d; wi /= d;
                }
                const int64_t o = frame * out_stride + k;
                if (out_f64) {
                    if (corr_out) reinterpret_cast<double2 *>(corr_out)[o] = make_double2(wr, wi);
                    if (mag_out) reinterpret_cast<double *>(mag_out)[o] = hypot(wr, wi);
                } else {
                    if (corr_out) reinterpret_cast<float2 *>(corr_out)[o] = make_float2((float)wr, (float)wi);
                    if (mag_out) reinterpret_cast<float *>(mag_out)[o] = (float)hypot(wr, wi);
                }
            } else {
                acc[i].x += yr; acc[i].y += yi;
                pw[i] += e;
            }
        }
    }
    if (single) return;
    __syncthreads();
    for (int i = tid; i < V; i += ZNT) {
        const int64_t k = k0 + i;
        if (k >= out_len) break;
        double yr = (double)acc[i].x, yi = (double)acc[i].y;
        if (mode == 0) {                                       // zc.py:125-126: normalise after the branch sum
            const double p = pw[i] > (T)0 ? (double)pw[i] : 0.0;
            const double d = ref_norm * sqrt(p + 1e-12);
            yr /= d; yi /= d;
        }
        const int64_t o = frame * out_stride + k;
        if (out_f64) {
            if (corr_out) reinterpret_cast<double2 *>(corr_out)[o] = make_double2(yr, yi);
            if (mag_out) reinterpret_cast<double *>(mag_out)[o] = hypot(yr, yi);
        } else {
            if (corr_out) reinterpret_cast<float2 *>(corr_out)[o] = make_float2((float)yr, (float)yi);
            if (mag_out) reinterpret_cast<float *>(mag_out)[o] = (float)hypot(yr, yi);
        }
    }
}

// ---- K4b: the float32 single-branch matched filter on 8192-point blocks ------------------------------------------------
// With the 2048-tap reference a 4096-point overlap-save block advances 2049 outputs (50 % useful), an 8192-point block 6145
// (75 %) for 13/12 of the butterflies per point.  Same structure as zc_mf_kernel (load -> energy prefix -> forward FFT ->
// pointwise product in transform order -> inverse FFT -> normalise), float32 throughout the epilogue (reciprocal square root
// instead of float64 divisions and hypot), the filter spectrum as a float2 table, |corr| as the only mandatory output.
__device__ __forceinline__ int spad(int i) { return i + (i >> 5); }     // energy prefix: 32-element thread stride -> 33
__device__ __forceinline__ float sqrt_approx(float v) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }   // MUFU.SQRT, <= 1 ulp off

__global__ void zc_twiddle8_kernel(float2 *tw8f, double2 *tw8d)
{
    const int i = threadIdx.x;              // 256 entries: exp(-2 pi i t / 8192)
    double sn, c;
    sincospi(-2.0 * (double)i / (double)ZF8, &sn, &c);
    tw8d[i] = make_double2(c, sn);
    tw8f[i] = make_float2((float)c, (float)sn);
}

// G8[p] = 8192-point DIF FFT of g[m] = conj(ref[nr-1-m]) zero-padded, in transform order (float2), plus ||ref||
__global__ void __launch_bounds__(ZNT) zc_spectrum8k_kernel(const double2 *ref, int nr, const double2 *tw, const double2 *tw8d, float2 *G8,
                                                            double *ref_norm)
{
    extern __shared__ __align__(16) unsigned char zsm[];
    double2 *a = reinterpret_cast<double2 *>(zsm);
    __shared__ double red[ZNT / 32];
    double e = 0.0;
    for (int m = threadIdx.x; m < ZF8; m += ZNT) {
        double2 v = make_double2(0.0, 0.0);
        if (m < nr) { const double2 r = ref[nr - 1 - m]; v = make_double2(r.x, -r.y); e += r.x * r.x + r.y * r.y; }
        a[zpad8(m)] = v;
    }
    for (int o = 16; o > 0; o >>= 1) e += shfl_xor_f64(e, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
    __syncthreads();
    fft8k_dif<double2>(a, tw, tw8d[threadIdx.x]);
    // stage-C order of conv8k.cuh, 1/8192 folded in
    for (int m = threadIdx.x; m < ZF8; m += ZNT) {
        const double2 g = a[zpad8(m)];
        const int h = m >> 12, t = (m & (ZF - 1)) >> 4, q = m & 15;
        G8[conv8k_gidx(h, t, q)] = make_float2((float)(g.x / ZF8), (float)(g.y / ZF8));
    }
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < ZNT / 32; ++w) t += red[w];
        *ref_norm = sqrt(t);
    }
}

// MODE (0: zc.py:125-126, 1: zc_v2.py:257-271, 2: raw) and the presence of the complex output are compile-time: the epilogue
// is a third of the kernel's instructions and every run-time choice in it is paid 32 times per thread
template <int DT, int MODE, bool CORR>
__global__ void __launch_bounds__(ZNT, 2) zc_mf8k_kernel(const void *x, int64_t n, int nr, const double2 *tw, const float2 *tw8,
                                                        const float2 *Gp, const double *ref_norm_p, float2 *corr_out,
                                                        float *mag_out, int64_t out_stride, int blocks_per_frame)
{
    constexpr int mode = MODE;
    using In = typename InT<DT>::type;
    extern __shared__ __align__(16) unsigned char zsm[];
    float2 *a = reinterpret_cast<float2 *>(zsm);                              // ZFP8
    float *se = reinterpret_cast<float *>(zsm + (size_t)ZFP8 * sizeof(float2)); // energy prefix, se[spad(k)] = sum of the first k
    __shared__ double wtot[ZNT / 32];
    const int V = ZF8 - nr + 1;                       // valid outputs per block
    const int64_t frame = blockIdx.x / blocks_per_frame;
    const int blk = blockIdx.x % blocks_per_frame;
    const int64_t k0 = (int64_t)blk * V;              // first full-convolution output of this block
    const int64_t jb = k0 - (nr - 1);                 // first input sample of this block
    const int64_t out_len = n + nr - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const In *xb = reinterpret_cast<const In *>(x) + frame * n;
    const float2 w0 = __ldg(tw8 + tid);
    constexpr bool norm = MODE != 2;
    // local samples [m_lo, m_hi) of the block lie inside the capture (one unsigned compare per sample)
    const int m_lo = jb < 0 ? (int)(-jb) : 0;
    const int m_hi = n - jb < ZF8 ? (int)(n - jb) : ZF8;
    const unsigned m_cnt = m_hi > m_lo ? (unsigned)(m_hi - m_lo) : 0u;
    const In *xl = xb + jb;
    const pk::Seeds s_ae = conv8k_seeds_ae(tw);
    conv8k_stage_a(a, s_ae, w0,
                   [&](int m) {
                       float2 v = make_float2(0.f, 0.f);
                       if ((unsigned)(m - m_lo) < m_cnt) { const In s = xl[m]; v = make_float2((float)s.x, (float)s.y); }
                       return v;
                   },
                   [&](int m, float e) { if (norm) se[spad(m + 1)] = e; });
    const pk::Seeds s_bd = conv8k_seeds_bd(tw);           // for stages B and D, fetched ahead of the barrier
    if (tid == 0) se[0] = 0.f;
    __syncthreads();
    // inclusive prefix of the 8192 energies: 32 contiguous values per thread (stride 33 after padding: conflict-free),
    // warp / CTA carries in float64; the CTA carry is added after the next barrier (the one stage B needs anyway)
    constexpr int IPT = ZF8 / ZNT;
    const int s0 = tid * IPT + 1;
    double woff = 0.0;
    if (norm) {
        float run = 0.f;
#pragma unroll
        for (int m = 0; m < IPT; ++m) { run += se[spad(s0 + m)]; se[spad(s0 + m)] = run; }
        double t = (double)run;
        for (int o = 1; o < 32; o <<= 1) { const double y = shfl_up_f64(t, o); if (lane >= o) t += y; }
        if (lane == 31) wtot[warp] = t;
        woff = t - (double)run;
    }
    conv8k_stage_b(a, s_bd);
    __syncthreads();
    if (norm) {
        for (int w = 0; w < warp; ++w) woff += wtot[w];
#pragma unroll
        for (int m = 0; m < IPT; ++m) se[spad(s0 + m)] = (float)((double)se[spad(s0 + m)] + woff);
    }
    conv8k_stage_c<false>(a, Gp, nullptr, nullptr);
    __syncthreads();
    conv8k_stage_d(a, s_bd);
    const pk::Seeds s_e = conv8k_seeds_ae(tw);
    __syncthreads();
    const float rn = (float)(1.0 / *ref_norm_p);
    // local outputs m in [nr - 1, o_hi) are this block's: output k0 + i, i = m - (nr - 1), is the window of local samples [i, i + nr - 1]
    const int64_t left = out_len - k0;
    const unsigned o_cnt = (unsigned)(left < V ? left : V);
    float2 *co = CORR ? corr_out + frame * out_stride + k0 - (nr - 1) : nullptr;
    float *mo = mag_out ? mag_out + frame * out_stride + k0 - (nr - 1) : nullptr;
    conv8k_stage_e(a, s_e, w0, conv8k_no_pre(), [&](int m, float2 y, int) {
        const int i = m - (nr - 1);
        if ((unsigned)i >= o_cnt) return;
        float sc = 1.f;
        if (norm) {
            const float e = se[spad(i + nr)] - se[spad(i)];
            sc = rn * (mode == 1 ? rsqrtf(fmaxf(e, 1e-12f))              // zc_v2.py:257-271
                                 : rsqrtf(fmaxf(e, 0.f) + 1e-12f));     // zc.py:125-126
        }
        const float2 ys = __fmul2_rn(y, make_float2(sc, sc));
        if (CORR) co[m] = ys;
        if (!CORR || mo) mo[m] = sqrt_approx(fmaf(ys.x, ys.x, ys.y * ys.y));
    });
}

// ---- K5b: zc_freq.compute_frequency_metric (zc_freq.py:62-99) in FFT form, float32 -----------------------------------
// With bins_j(o) = DFT bin k_j of x[o+cp : o+cp+N]:
//   Y(o) = sum_j conj(T_j) bins_j(o) = sum_m x[o+cp+m] h_Y[m],  h_Y[m] = sum_j conj(T_j) e^{-2 pi i k_j m / N}   (np.vdot, :94)
//   S(o) = sum_j bins_j(o)            = sum_m x[o+cp+m] h_S[m],  h_S[m] = sum_j e^{-2 pi i k_j m / N}
// are two N-tap matched filters of the same signal, and the in-band energy E(o) = sum_j |bins_j(o)|^2 (:95) obeys
//   bins_j(o+1) = w_j (bins_j(o) + d(o)),  d(o) = x[o+cp+N] - x[o+cp],  |w_j| = 1
//   =>  E(o+1) = E(o) + 2 Re(conj(S(o)) d(o)) + J |d(o)|^2                                       (J = number of bins)
// so the metric |Y|^2 / max(E_T E, eps) (:96-97) costs one forward FFT, two pointwise products and two inverse FFTs per
// 8192-sample block (conv8k.cuh) plus a prefix sum -- instead of J sliding-DFT recurrences per offset (zc_sdft_kernel) or one
// N-point FFT per offset (the reference).  One CTA walks consecutive blocks of one capture: E is anchored once per CTA by a
// direct J-bin DFT of the first window and carried from block to block in float64; inside a block the prefix of the
// increments is thread-serial float32 with float64 warp / CTA carries, rounded to float32 once.
// Offsets whose E is below 1e-7 of the largest E seen so far in the CTA's range (this block included) give 0: there the
// float32 FFT's rounding residue of |Y|^2 would be divided by a rounding residue of E (the reference's eps clamp gives 0
// on exact silence).
constexpr int ZQF_THREADS = ZNT;
struct ZcFreqFftParams {
    const void *x;             // complex64 or int16 IQ, [frames][nb][n]
    int64_t n, n_off, mstride;
    int N, cp, nbins, nb, blocks_per_cap, blocks_per_item, items_per_cap;     // nb: receive branches, [frames][nb][n]
    int64_t n_items;
    const int *bins;
    const double2 *tw;
    const float2 *tw8, *GpY, *GpS;
    float2 *stash;             // [gridDim.x][8192]
    float templ_energy;
    float *metric;
};

// the block function is out of line (see there); it reads the launch parameters from this shared copy (one LDS, no registers
// held across the FFT stages) -- through a reference to the kernel's parameter block every access was a generic global load
__shared__ ZcFreqFftParams g_zqf;
// sample as float2 from complex64 or int16-IQ rows
template <typename In> __device__ __forceinline__ float2 zq_ld(const In *p)
{
    const In s = __ldg(p);
    return make_float2((float)s.x, (float)s.y);
}

template <typename In>
__device__ double zqf_anchor(const In *xl, int64_t avail, int N, const int *bins, int nbins, const double2 *tw, float2 *red, double *dred)
{
    const int tid = threadIdx.x, j = tid & 63, part = tid >> 6;
    const int chunk = (N + 3) >> 2, m0 = part * chunk, m1 = m0 + chunk < N ? m0 + chunk : N;
    float2 acc = make_float2(0.f, 0.f);
    if (j < nbins) {
        int k = bins[j] % N;
        if (k < 0) k += N;
        const bool tab = (ZF % N) == 0;
        const int mulf = tab ? ZF / N : 0;
        auto phasor = [&](int ph) {                                           // e^{-2 pi i ph / N}, 0 <= ph < N
            if (tab) return tw4096<float2>(tw, ph * mulf);
            float sn, cs;
            sincospif(-2.0f * (float)ph / (float)N, &sn, &cs);
            return make_float2(cs, sn);
        };
        const float2 ws = phasor(k);
        // the phasor advances by rotation and is re-seeded exactly every 32 samples (a table load per sample made this loop
        // latency-bound: 35 % of the kernel's stall samples)
        for (int mb = m0; mb < m1; mb += 32) {
            float2 w = phasor((int)(((long long)k * mb) % N));
            if (mb + 32 <= m1 && mb + 32 <= avail) {                          // whole group inside the window and the capture
#pragma unroll 8
                for (int u = 0; u < 32; u += 2) {
                    const float2 v0 = zq_ld(xl + mb + u), v1 = zq_ld(xl + mb + u + 1);
                    const float4 v = make_float4(v0.x, v0.y, v1.x, v1.y);
                    acc = __ffma2_rn(make_float2(v.x, v.x), w, acc);
                    acc = __ffma2_rn(make_float2(v.y, v.y), make_float2(-w.y, w.x), acc);
                    w = pk::mul(w, ws);
                    acc = __ffma2_rn(make_float2(v.z, v.z), w, acc);
                    acc = __ffma2_rn(make_float2(v.w, v.w), make_float2(-w.y, w.x), acc);
                    w = pk::mul(w, ws);
                }
            } else {
                for (int u = 0; u < 32 && mb + u < m1; ++u) {
                    const int m = mb + u;
                    const float2 v = m < avail ? zq_ld(xl + m) : make_float2(0.f, 0.f);
                    acc.x = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc.x));
                    acc.y = fmaf(v.x, w.y, fmaf(v.y, w.x, acc.y));
                    w = pk::mul(w, ws);
                }
            }
        }
    }
    red[tid] = acc;
    __syncthreads();
    if (tid < 64) {
        const float2 a0 = red[tid], a1 = red[tid + 64], a2 = red[tid + 128], a3 = red[tid + 192];
        const double sx = (double)a0.x + a1.x + a2.x + a3.x, sy = (double)a0.y + a1.y + a2.y + a3.y;
        double e = sx * sx + sy * sy;
        for (int o = 16; o > 0; o >>= 1) e += shfl_xor_f64(e, o);
        if ((tid & 31) == 0) dred[tid >> 5] = e;
    }
    __syncthreads();
    const double r = dred[0] + dred[1];
    __syncthreads();
    return r;
}

// one 8192-sample block: offsets o0 .. o0 + V - 1 of capture row xc; returns E at the first offset of the next block.
// Kept out of line: the block loop's own state (item, capture, carries) then lives across ONE call instead of competing with
// the 100+ registers each FFT stage wants.
// S path of one branch of one block: forward transform, both products (S back to shared memory, Y into / onto the stash),
// inverse transform of S, increments of E into se (added to the previous branches' when `add`).  Out of line for the same
// reason as zqf_block: called once per branch, its 100+ registers per stage do not compete with the caller's state.
template <typename In>
__device__ __noinline__ void zqf_spath(float2 *a, float *se, const In *xb, int64_t availl, unsigned o_cnt, bool add)
{
    const ZcFreqFftParams &p = g_zqf;
    const int tid = threadIdx.x;
    const int N = p.N, V = ZF8 - N + 1;
    const float2 w0 = __ldg(p.tw8 + tid);
    const unsigned m_cnt = (unsigned)(availl < ZF8 ? availl : ZF8);
    auto ldx = [&](int m) { return (unsigned)m < m_cnt ? zq_ld(xb + m) : make_float2(0.f, 0.f); };
    // twiddle seeds are fetched ahead of the barrier that precedes their stage
    pk::Seeds sd = conv8k_seeds_ae(p.tw);
    conv8k_stage_a(a, sd, w0, ldx, [](int, float) {});
    sd = conv8k_seeds_bd(p.tw);
    __syncthreads();
    conv8k_stage_b(a, sd);
    __syncthreads();
    conv8k_stage_c<true>(a, p.GpS, p.GpY, p.stash + (size_t)blockIdx.x * ZF8, add);      // S back to shared memory, Y to the stash
    sd = conv8k_seeds_bd(p.tw);
    __syncthreads();
    conv8k_stage_d(a, sd);
    sd = conv8k_seeds_ae(p.tw);
    __syncthreads();
    // S(o) -> increment of E: se[spad(i + 1)] = E(o0 + i + 1) - E(o0 + i)
    const float Jf = (float)p.nbins;
    struct DPair { float2 hi, lo; };
    conv8k_stage_e(a, sd, w0,
                   [&](int m) {                                 // d(o) = x[o+cp+N] - x[o+cp], fetched 8 outputs ahead
                       const int i = m - (N - 1);
                       DPair d;
                       d.hi = d.lo = make_float2(0.f, 0.f);
                       if ((unsigned)i < o_cnt) {
                           if ((int64_t)(i + N) < availl) d.hi = zq_ld(xb + i + N);      // i + N may be sample 8192: past the block, inside the capture
                           d.lo = ldx(i);
                       }
                       return d;
                   },
                   [&](int m, float2 S, const DPair &d) {
                       const int i = m - (N - 1);
                       if ((unsigned)i >= (unsigned)V) return;
                       const float dx = d.hi.x - d.lo.x, dy = d.hi.y - d.lo.y;
                       const float dl = (unsigned)i < o_cnt ? fmaf(2.f, fmaf(S.x, dx, S.y * dy), Jf * fmaf(dx, dx, dy * dy)) : 0.f;
                       se[spad(i + 1)] = add ? se[spad(i + 1)] + dl : dl;       // the same thread wrote it for the previous branch
                   });
    __syncthreads();
}

// one 8192-sample block: offsets o0 .. o0 + V - 1 of capture row xc; returns E at the first offset of the next block.
// Kept out of line: the block loop's own state (item, capture, carries) then lives across ONE call instead of competing with
// the 100+ registers each FFT stage wants.  Branches (zc_freq.py:88-97): Y is linear in the samples, so the branches' Y
// products are summed in the stash (one inverse transform for all of them); E is not, so every branch runs its own S path
// and the increments add up in se.
template <typename In>
__device__ __noinline__ double zqf_block(float2 *a, float *se, float2 *red, double *wtot, float *wmax, double *dred,
                                        const In *xc, float *mrow, int b, double Eb, float *emax_io)
{
    const ZcFreqFftParams &p = g_zqf;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, V = ZF8 - N + 1;
    const int64_t o0 = (int64_t)b * V;                    // first offset of the block; local sample m = x[cp + o0 + m]
    const In *xl = xc + p.cp + o0;
    const int64_t availl = p.n - p.cp - o0;
    const int64_t left = p.n_off - o0;
    const unsigned o_cnt = (unsigned)(left < V ? left : V);
    constexpr int IPT = ZF8 / ZNT;
    const int s0 = tid * IPT + 1;
    const int n_br = p.nb;
    for (int br = 0; br < n_br; ++br) zqf_spath(a, se, xl + (int64_t)br * p.n, availl, o_cnt, br > 0);
    // the Y product comes back from the stash while the prefix sum runs (an L2 round trip: 14 % of the kernel when exposed)
    float2 yst[32];
    conv8k_unstash_fetch(p.stash + (size_t)blockIdx.x * ZF8, yst);
    // inclusive prefix: thread-serial float32 over 32 contiguous increments, float64 warp / CTA / block-to-block carries
    float run = 0.f;
#pragma unroll
    for (int m = 0; m < IPT; ++m) { run += se[spad(s0 + m)]; se[spad(s0 + m)] = run; }
    double t = (double)run;
    for (int o = 1; o < 32; o <<= 1) { const double y = shfl_up_f64(t, o); if (lane >= o) t += y; }
    if (lane == 31) wtot[warp] = t;
    __syncthreads();
    double off = Eb + (t - (double)run), tot = 0.0;
    for (int w = 0; w < ZNT / 32; ++w) { if (w < warp) off += wtot[w]; tot += wtot[w]; }
    // E itself into se (rounded to float32 once), and the block's largest / smallest E with the position of the largest
    float mx = tid == 0 ? (float)Eb : -1.f, mn = tid == 0 ? (float)Eb : 3.0e38f;
    int ix = 0;
#pragma unroll
    for (int m = 0; m < IPT; ++m) {
        const float e = (float)((double)se[spad(s0 + m)] + off);
        se[spad(s0 + m)] = s0 + m <= V ? e : 0.f;                  // the tail beyond the block's V offsets stays zero: it is summed again
        if (s0 + m <= (int)o_cnt) {
            if (e > mx) { mx = e; ix = s0 + m; }
            mn = fminf(mn, e);
        }
    }
    if (tid == 0) se[0] = (float)Eb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float omx = __shfl_xor_sync(0xffffffffu, mx, o), omn = __shfl_xor_sync(0xffffffffu, mn, o);
        const int oix = __shfl_xor_sync(0xffffffffu, ix, o);
        if (omx > mx || (omx == mx && oix < ix)) { mx = omx; ix = oix; }
        mn = fminf(mn, omn);
    }
    if (lane == 0) { wmax[warp] = mx; wmax[ZNT / 32 + warp] = mn; reinterpret_cast<int *>(wmax)[2 * (ZNT / 32) + warp] = ix; }
    __syncthreads();
    float emax = *emax_io;
#pragma unroll
    for (int w = 0; w < ZNT / 32; ++w) {
        const float v = wmax[w];
        const int vi = reinterpret_cast<const int *>(wmax)[2 * (ZNT / 32) + w];
        if (w == 0 || v > mx || (v == mx && vi < ix)) { mx = v; ix = vi; }
        mn = w == 0 ? wmax[ZNT / 32] : fminf(mn, wmax[ZNT / 32 + w]);
    }
    emax = fmaxf(emax, mx);
    *emax_io = emax;
    // A block whose E spans more than ~27 dB (a burst entering and leaving the window, or silence): the float32 FFT's error in
    // S, multiplied by the burst-sized d, has accumulated ~3e-8 of the LARGEST E in the prefix -- too much for the small E
    // that follows.  Anchor E at the block's end directly as well; the discrepancy D belongs to the offsets after the peak of
    // E (it was gathered while the burst crossed the window, where E itself is huge), and the exact value is what the next
    // block starts from.
    double Enext = Eb + tot;
    float Df = 0.f;
    if (mn < 2e-3f * mx) {                                         // block-uniform
        double Ed = 0.0;
        for (int br = 0; br < n_br; ++br)
            Ed += zqf_anchor(xl + (int64_t)br * p.n + o_cnt, availl - (int64_t)o_cnt, N, p.bins, p.nbins, p.tw, red, dred);
        Df = (float)(Enext - Ed);
        Enext = Ed;
    }
    conv8k_unstash_put(a, yst);
    const float2 w0 = __ldg(p.tw8 + tid);
    pk::Seeds sd = conv8k_seeds_bd(p.tw);
    __syncthreads();
    conv8k_stage_d(a, sd);
    sd = conv8k_seeds_ae(p.tw);
    __syncthreads();
    {
        const float floor_e = 1e-7f * emax, te = p.templ_energy;
        float *mo = mrow + o0 - (N - 1);
        conv8k_stage_e(a, sd, w0, conv8k_no_pre(), [&](int m, float2 Y, int) {
            const int i = m - (N - 1);
            if ((unsigned)i >= o_cnt) return;
            const float e = se[spad(i)] - (i > ix ? Df : 0.f);
            const float den = fmaxf(te * e, 1e-12f);                              // zc_freq.py:96
            mo[m] = (e > floor_e && e > 0.f) ? __fdividef(fmaf(Y.x, Y.x, Y.y * Y.y), den) : 0.f;
        });
    }
    __syncthreads();
    return Enext;
}

template <int DT>
__global__ void __launch_bounds__(ZQF_THREADS, 2) zc_freq_fft_kernel(const __grid_constant__ ZcFreqFftParams p)
{
    using In = typename InT<DT>::type;
    extern __shared__ __align__(16) unsigned char zsm[];
    float2 *a = reinterpret_cast<float2 *>(zsm);                                   // ZFP8
    float *se = reinterpret_cast<float *>(zsm + (size_t)ZFP8 * sizeof(float2));     // se[spad(i)] = E(first offset of the block + i)
    __shared__ float2 red[ZNT];                        // the anchor's partial sums
    __shared__ double wtot[ZNT / 32];
    __shared__ float wmax[3 * (ZNT / 32)];            // per warp: max E, min E, position of the max
    __shared__ double dred[2];
    const int tid = threadIdx.x;
    const int V = ZF8 - p.N + 1;
    for (int i = tid; i < ZF8 + ZF8 / 32 + 8; i += ZNT) se[i] = 0.f;
    if (tid == 0) g_zqf = p;
    for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int64_t cap = item / p.items_per_cap;
        const int b0 = (int)(item % p.items_per_cap) * p.blocks_per_item;
        const int b1 = b0 + p.blocks_per_item < p.blocks_per_cap ? b0 + p.blocks_per_item : p.blocks_per_cap;
        const In *xc = reinterpret_cast<const In *>(p.x) + cap * p.nb * p.n;
        __syncthreads();
        // E at the item's first offset, directly (summed over the branches)
        double Eb = 0.0;
        for (int br = 0; br < p.nb; ++br)
            Eb += zqf_anchor(xc + (int64_t)br * p.n + p.cp + (int64_t)b0 * V, p.n - p.cp - (int64_t)b0 * V, p.N, p.bins, p.nbins, p.tw, red, dred);
        float emax = (float)Eb;
        for (int b = b0; b < b1; ++b) Eb = zqf_block(a, se, red, wtot, wmax, dred, xc, p.metric + cap * p.mstride, b, Eb, &emax);
    }
}

// ref[m] = sum_j T_j e^{+2 pi i k_j m / N} (T_j = 1 without a template): the time-domain sequence whose matched filter
// (conj, reversed -- zc_spectrum8k_kernel) is h_Y / h_S above
__global__ void zqf_ref_kernel(const int *bins, const float2 *templ, int nbins, int N, double2 *ref)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N) return;
    double sr = 0.0, si = 0.0;
    for (int j = 0; j < nbins; ++j) {
        int k = bins[j] % N;
        if (k < 0) k += N;
        double sn, cs;
        sincospi(2.0 * (double)(((long long)k * m) % N) / (double)N, &sn, &cs);
        const double tr = templ ? (double)templ[j].x : 1.0, ti = templ ? (double)templ[j].y : 0.0;
        sr += tr * cs - ti * sn;
        si += tr * sn + ti * cs;
    }
    ref[m] = make_double2(sr, si);
}

// ---- zc_v2.normalize_correlation (zc_v2.py:257-271) on a correlation the CALLER supplies ---------------------
// out[k] = corr[k] / (ref_norm * sqrt(max(E[k], 1e-12))), E[k] = sum of |x[j]|^2 over j in [k-nr+1, k] inside the capture
// (np.convolve(|x|^2, ones(nr), "full")).  One CTA per (frame, tile of NZT outputs): float64 prefix of the tile + halo.
constexpr int NZT = 4096;
template <int DT>
__global__ void __launch_bounds__(256) zc_normalize_kernel(const void *corr, const void *x, int64_t n, int nr, double ref_norm,
                                                           int f64, void *out, int64_t stride)
{
    using In = typename InT<DT>::type;
    extern __shared__ double nzs[];             // nzs[k] = sum of |x|^2 over [j0, j0 + k)
    __shared__ double wtot[8];
    const int64_t frame = blockIdx.y, k0 = (int64_t)blockIdx.x * NZT, out_len = n + nr - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t kend = k0 + NZT < out_len ? k0 + NZT : out_len;
    int64_t j0 = k0 - nr + 1;
    if (j0 < 0) j0 = 0;
    const int64_t j1 = kend < n ? kend : n;      // samples [j0, j1)
    const int cnt = (int)(j1 - j0 > 0 ? j1 - j0 : 0);
    const In *xr = reinterpret_cast<const In *>(x) + frame * n;
    for (int k = tid; k < cnt; k += 256) { const In v = xr[j0 + k]; nzs[k + 1] = (double)v.x * v.x + (double)v.y * v.y; }
    if (tid == 0) nzs[0] = 0.0;
    __syncthreads();
    const int ipt = ((cnt + 255) / 256) | 1;
    const int s0 = 1 + tid * ipt, s1 = min(s0 + ipt, cnt + 1);
    double acc = 0.0;
    for (int q = s0; q < s1; ++q) { acc += nzs[q]; nzs[q] = acc; }
    double t = acc;
    for (int o = 1; o < 32; o <<= 1) { const double y = shfl_up_f64(t, o); if (lane >= o) t += y; }
    if (lane == 31) wtot[warp] = t;
    __syncthreads();
    double off = t - acc;
    for (int w = 0; w < warp; ++w) off += wtot[w];
    for (int q = s0; q < s1; ++q) nzs[q] += off;
    __syncthreads();
    for (int64_t k = k0 + tid; k < kend; k += 256) {
        int64_t lo = k - nr + 1, hi = k + 1;
        if (lo < j0) lo = j0;
        if (hi > j1) hi = j1;
        const double e = hi > lo ? nzs[hi - j0] - nzs[lo - j0] : 0.0;
        const double dnm = ref_norm * sqrt(e > 1e-12 ? e : 1e-12);
        const int64_t o = frame * stride + k;
        if (f64) {
            const double2 c = reinterpret_cast<const double2 *>(corr)[o];
            reinterpret_cast<double2 *>(out)[o] = make_double2(c.x / dnm, c.y / dnm);
        } else {
            const float2 c = reinterpret_cast<const float2 *>(corr)[o];
            reinterpret_cast<float2 *>(out)[o] = make_float2((float)((double)c.x / dnm), (float)((double)c.y / dnm));
        }
    }
}

// ---- zc_freq: sliding DFT of the used bins --------------------------------------------------------
constexpr int QNT = 256;

template <int DT>
__global__ void __launch_bounds__(QNT) zc_freq_kernel(const void *x, int nb, int64_t n, int N, int cp, const int *bins,
                                                      const double2 *templ, int nbins, double templ_energy, int TO,
                                                      int64_t n_off, int out_f64, void *metric, int64_t out_stride,
                                                      int tiles_per_frame)
{
    using In = typename InT<DT>::type;
    extern __shared__ __align__(16) unsigned char qsm[];
    const int span = TO + N - 1;                                   // samples of the tile
    double2 *xs = reinterpret_cast<double2 *>(qsm);               // span
    double2 *S = xs + span;                                        // span + 1 (prefix of the modulated samples)
    double2 *tw = S + span + 1;                                    // N: exp(-2 pi i m / N)
    __shared__ double wtot[2][QNT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t frame = blockIdx.x / tiles_per_frame;
    const int tile = blockIdx.x % tiles_per_frame;
    const int64_t o0 = (int64_t)tile * TO;
    const int64_t jb = o0 + cp;                                    // first sample of the tile
    for (int m = tid; m < N; m += QNT) {
        double s, c;
        sincospi(-2.0 * (double)m / (double)N, &s, &c);
        tw[m] = make_double2(c, s);
    }
    const int ipt = ((span + QNT - 1) / QNT) | 1;
    const int s0 = tid * ipt, s1 = min(s0 + ipt, span);
    constexpr int OPT = 8;                                         // offsets per thread (TO <= QNT * OPT)
    double cr[OPT], ci[OPT], en[OPT];
#pragma unroll
    for (int q = 0; q < OPT; ++q) cr[q] = ci[q] = en[q] = 0.0;

    for (int b = 0; b < nb; ++b) {
        const In *xb = reinterpret_cast<const In *>(x) + (frame * nb + b) * n;
        __syncthreads();
        for (int m = tid; m < span; m += QNT) {
            const int64_t j = jb + m;
            double2 v = make_double2(0.0, 0.0);
            if (j < n) { const In s = xb[j]; v = make_double2((double)s.x, (double)s.y); }
            xs[m] = v;
        }
        if (tid == 0) S[0] = make_double2(0.0, 0.0);
        __syncthreads();
        for (int jbin = 0; jbin < nbins; ++jbin) {
            const int k = bins[jbin];
            // modulated inclusive prefix over the tile: S[m+1] = sum_{m'<=m} x[jb+m'] * w^(k (jb+m'))
            double rr = 0.0, ri = 0.0;
            int ph = (int)((jb + s0) % N);
            for (int m = s0; m < s1; ++m) {
                const double2 w = tw[(int)(((long long)k * ph) % N)];
                const double2 v = xs[m];
                rr += v.x * w.x - v.y * w.y;
                ri += v.x * w.y + v.y * w.x;
                S[m + 1] = make_double2(rr, ri);
                ph = ph + 1 == N ? 0 : ph + 1;
            }
            double tr = rr, ti = ri;
            for (int o = 1; o < 32; o <<= 1) {
                const double yr = shfl_up_f64(tr, o), yi = shfl_up_f64(ti, o);
                if (lane >= o) { tr += yr; ti += yi; }
            }
            if (lane == 31) { wtot[0][warp] = tr; wtot[1][warp] = ti; }
            __syncthreads();
            double offr = tr - rr, offi = ti - ri;
            for (int w = 0; w < warp; ++w) { offr += wtot[0][w]; offi += wtot[1][w]; }
            for (int m = s0; m < s1; ++m) { S[m + 1].x += offr; S[m + 1].y += offi; }
            __syncthreads();
            const double2 tj = templ[jbin];
#pragma unroll
            for (int q = 0; q < OPT; ++q) {
                const int i = tid + q * QNT;                        // offset o0 + i  <->  s = jb + i
                if (i < TO && o0 + i < n_off) {
                    const double2 hi = S[i + N], lo = S[i];
                    const double dr = hi.x - lo.x, di = hi.y - lo.y;
                    const double2 w = tw[(int)(((long long)k * ((jb + i) % N)) % N)];
                    // bin = conj(w) * d
                    const double br = dr * w.x + di * w.y, bi = di * w.x - dr * w.y;
                    cr[q] += tj.x * br + tj.y * bi;                 // conj(template) * bin (np.vdot)
                    ci[q] += tj.x * bi - tj.y * br;
                    en[q] += br * br + bi * bi;
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int q = 0; q < OPT; ++q) {
        const int i = tid + q * QNT;
        if (i < TO && o0 + i < n_off) {
            double den = templ_energy * en[q];
            if (den < 1e-12) den = 1e-12;
            const double m = (cr[q] * cr[q] + ci[q] * ci[q]) / den;
            const int64_t o = frame * out_stride + o0 + i;
            if (out_f64) reinterpret_cast<double *>(metric)[o] = m;
            else reinterpret_cast<float *>(metric)[o] = (float)m;
        }
    }
}

}  // namespace ofs

using namespace ofs;

OFS_API int ofs_zc_matched_filter(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n,
                                  const void *ref_c128, int32_t nr, int32_t mode, int32_t out_f64, void *corr_out,
                                  void *mag_out, int64_t out_stride, void *stream_)
{
    OFS_TRACE();
    OFS_REQUIRE(x && ref_c128 && (corr_out || mag_out), "ofs_zc_matched_filter: null argument");
    OFS_REQUIRE(nr >= 1 && nr <= 2048, "ofs_zc_matched_filter: reference length must be 1..2048");
    OFS_REQUIRE(mode >= 0 && mode <= 2, "ofs_zc_matched_filter: mode must be 0, 1 or 2");
    OFS_REQUIRE(n_branches >= 1 && n >= 1 && n_frames >= 0, "ofs_zc_matched_filter: bad geometry");
    OFS_REQUIRE(out_stride >= n + nr - 1, "ofs_zc_matched_filter: out_stride too small");
    if (n_frames == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    keep_pool_cached();
    AsyncBuf b_tw, b_G, b_rn, b_tw8f, b_tw8d, b_G8;                     // returned to the pool on every exit path
    OFS_CUDA(b_tw.alloc((ZF / 2) * (sizeof(double2) + sizeof(float2)), stream));
    OFS_CUDA(b_G.alloc(ZF * sizeof(double2), stream));
    OFS_CUDA(b_rn.alloc(sizeof(double), stream));
    double2 *tw = b_tw.as<double2>(), *G = b_G.as<double2>();
    double *rn = b_rn.as<double>();
    zc_twiddle_kernel<<<(ZF / 2 + 255) / 256, 256, 0, stream>>>(tw);
    if (int rc = check_launch("zc_twiddle_kernel")) return rc;
    const bool dbl = in_dtype == OFS_C128 || out_f64;
    // float32, one branch, captures of several blocks: the 8192-point kernel (75 % useful outputs per block instead of 50 %)
    static const int mf_block = [] { const char *e = getenv("OFS_MF_BLOCK"); return e ? atoi(e) : 8192; }();
    if (!dbl && n_branches == 1 && mf_block == 8192 && n + nr - 1 >= 2 * (ZF8 - nr + 1)) {
        OFS_CUDA(b_tw8f.alloc(256 * sizeof(float2), stream));
        OFS_CUDA(b_tw8d.alloc(256 * sizeof(double2), stream));
        OFS_CUDA(b_G8.alloc(ZF8 * sizeof(float2), stream));
        float2 *tw8f = b_tw8f.as<float2>(), *G8 = b_G8.as<float2>();
        double2 *tw8d = b_tw8d.as<double2>();
        zc_twiddle8_kernel<<<1, 256, 0, stream>>>(tw8f, tw8d);
        if (int rc = check_launch("zc_twiddle8_kernel")) return rc;
        OFS_CUDA(cudaFuncSetAttribute(zc_spectrum8k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ZFP8 * sizeof(double2))));
        zc_spectrum8k_kernel<<<1, ZNT, ZFP8 * sizeof(double2), stream>>>((const double2 *)ref_c128, nr, tw, tw8d, G8, rn);
        if (int rc = check_launch("zc_spectrum8k_kernel")) return rc;
        const int V8 = ZF8 - nr + 1;
        const int bpf8 = (int)((n + nr - 1 + V8 - 1) / V8);
        const int64_t grid8 = (int64_t)bpf8 * n_frames;
        OFS_REQUIRE(grid8 < (1LL << 31), "ofs_zc_matched_filter: grid too large");
        const size_t smem8 = (size_t)ZFP8 * sizeof(float2) + (size_t)(ZF8 + ZF8 / 32 + 8) * sizeof(float);
#define OFS_MF8_LAUNCH(DT, MODE, CORR)                                                                              \
    do {                                                                                                           \
        auto kern = zc_mf8k_kernel<DT, MODE, CORR>;                                                                \
        OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));             \
        kern<<<(unsigned)grid8, ZNT, smem8, stream>>>(x, n, nr, tw, tw8f, G8, rn, (float2 *)corr_out, (float *)mag_out, out_stride, bpf8); \
    } while (0)
#define OFS_MF8_MODE(DT, CORR)                                                                                     \
    do {                                                                                                           \
        if (mode == 0) OFS_MF8_LAUNCH(DT, 0, CORR);                                                                \
        else if (mode == 1) OFS_MF8_LAUNCH(DT, 1, CORR);                                                           \
        else OFS_MF8_LAUNCH(DT, 2, CORR);                                                                          \
    } while (0)
        if (in_dtype == OFS_C64) { if (corr_out) OFS_MF8_MODE(OFS_C64, true); else OFS_MF8_MODE(OFS_C64, false); }
        else { if (corr_out) OFS_MF8_MODE(OFS_IQ16, true); else OFS_MF8_MODE(OFS_IQ16, false); }
#undef OFS_MF8_MODE
#undef OFS_MF8_LAUNCH
        return check_launch("zc_mf8k_kernel");
    }
    // the 4096-point filter spectrum (only this path needs it: the 8192-point path above has its own)
    OFS_CUDA(cudaFuncSetAttribute(zc_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ZFP * sizeof(double2))));
    zc_spectrum_kernel<<<1, ZNT, ZFP * sizeof(double2), stream>>>((const double2 *)ref_c128, nr, tw, G, rn);
    if (int rc = check_launch("zc_spectrum_kernel")) return rc;
    const int V = ZF - nr + 1;
    const int bpf = (int)((n + nr - 1 + V - 1) / V);
    const int64_t grid = (int64_t)bpf * n_frames;
    OFS_REQUIRE(grid < (1LL << 31), "ofs_zc_matched_filter: grid too large");
    // a + se (+ pw + acc when branches are summed)
    const size_t esz_t = dbl ? 8 : 4;
    const size_t smem = (size_t)ZFP * 2 * esz_t + (size_t)(ZFP + 8) * esz_t + (n_branches > 1 ? (size_t)ZF * esz_t + (size_t)ZF * 2 * esz_t : 0);
#define OFS_MF_LAUNCH(T, DT)                                                                                       \
    do {                                                                                                           \
        auto kern = zc_mf_kernel<T, DT>;                                                                           \
        OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        kern<<<(unsigned)grid, ZNT, smem, stream>>>(x, n_branches, n, nr, tw, G, rn, mode, out_f64, corr_out, mag_out, \
                                                    out_stride, bpf);                                              \
    } while (0)
    if (in_dtype == OFS_C128) OFS_MF_LAUNCH(double, OFS_C128);
    else if (in_dtype == OFS_C64 && dbl) OFS_MF_LAUNCH(double, OFS_C64);
    else if (in_dtype == OFS_C64) OFS_MF_LAUNCH(float, OFS_C64);
    else if (in_dtype == OFS_IQ16) OFS_MF_LAUNCH(float, OFS_IQ16);
    else { set_error("ofs_zc_matched_filter: unknown dtype"); return OFS_EINVAL; }
#undef OFS_MF_LAUNCH
    return check_launch("zc_mf_kernel");
}

OFS_API int ofs_zc_normalize(const void *corr, const void *x, int32_t in_dtype, int64_t n_frames, int64_t n, int32_t nr,
                             double ref_norm, int32_t f64, void *out, int64_t stride, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(corr && x && out, "ofs_zc_normalize: null argument");
    OFS_REQUIRE(in_dtype >= OFS_C64 && in_dtype <= OFS_IQ16, "ofs_zc_normalize: unknown dtype");
    OFS_REQUIRE(n >= 1 && nr >= 1 && n_frames >= 0 && n_frames < 65536 && stride >= n + nr - 1, "ofs_zc_normalize: bad geometry");
    OFS_REQUIRE(nr <= 65536 && ref_norm > 0.0, "ofs_zc_normalize: reference length 1..65536, positive norm");
    if (n_frames == 0) return OFS_OK;
    const size_t smem = (size_t)(NZT + nr + 2) * sizeof(double);
    const dim3 grid((unsigned)((n + nr - 1 + NZT - 1) / NZT), (unsigned)n_frames);
#define OFS_NZ_LAUNCH(DT)                                                                                          \
    do {                                                                                                           \
        OFS_CUDA(cudaFuncSetAttribute(zc_normalize_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        zc_normalize_kernel<DT><<<grid, 256, smem, (cudaStream_t)stream>>>(corr, x, n, nr, ref_norm, f64, out, stride); \
    } while (0)
    if (in_dtype == OFS_C128) OFS_NZ_LAUNCH(OFS_C128);
    else if (in_dtype == OFS_C64) OFS_NZ_LAUNCH(OFS_C64);
    else OFS_NZ_LAUNCH(OFS_IQ16);
#undef OFS_NZ_LAUNCH
    return check_launch("zc_normalize_kernel");
}

OFS_API int ofs_zc_freq_metric(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n,
                               int32_t n_fft, int32_t cp, const int32_t *bins, const void *templ_c128, int32_t nbins,
                               double templ_energy, int32_t out_f64, void *metric, int64_t out_stride, void *stream_)
{
    OFS_TRACE();
    OFS_REQUIRE(x && bins && templ_c128 && metric, "ofs_zc_freq_metric: null argument");
    OFS_REQUIRE(n_fft >= 2 && n_fft <= 2048 && cp >= 0 && nbins >= 1, "ofs_zc_freq_metric: n_fft must be 2..2048");
    const int64_t n_off = n - ((int64_t)n_fft + cp) + 1;
    OFS_REQUIRE(n_off > 0, "Received stream is shorter than a single OFDM symbol.");   /* zc_freq.py:76-78 */
    OFS_REQUIRE(out_stride >= n_off && n_branches >= 1 && n_frames >= 0, "ofs_zc_freq_metric: bad geometry");
    if (n_frames == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    keep_pool_cached();
    int TO = 2048;
    if (n_off < TO) TO = (int)((n_off + 255) / 256 * 256);
    const int tiles = (int)((n_off + TO - 1) / TO);
    const int span = TO + n_fft - 1;
    const size_t smem = (size_t)(span + span + 1 + n_fft) * sizeof(double2);
    const int64_t grid = (int64_t)tiles * n_frames;
    OFS_REQUIRE(grid < (1LL << 31), "ofs_zc_freq_metric: grid too large");
#define OFS_ZQ_LAUNCH(DT)                                                                                          \
    do {                                                                                                           \
        auto kern = zc_freq_kernel<DT>;                                                                            \
        OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        kern<<<(unsigned)grid, QNT, smem, stream>>>(x, n_branches, n, n_fft, cp, bins, (const double2 *)templ_c128, nbins, \
                                                    templ_energy, TO, n_off, out_f64, metric, out_stride, tiles);  \
    } while (0)
    if (in_dtype == OFS_C128) OFS_ZQ_LAUNCH(OFS_C128);
    else if (in_dtype == OFS_C64) OFS_ZQ_LAUNCH(OFS_C64);
    else if (in_dtype == OFS_IQ16) OFS_ZQ_LAUNCH(OFS_IQ16);
    else { set_error("ofs_zc_freq_metric: unknown dtype"); return OFS_EINVAL; }
#undef OFS_ZQ_LAUNCH
    return check_launch("zc_freq_kernel");
}

OFS_API int ofs_zc_freq_metric_fft(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                                   const void *templ_c64, int32_t nbins, double templ_energy, float *metric, int64_t out_stride,
                                   void *stream_)
{
    OFS_TRACE();
    OFS_REQUIRE(x && bins && templ_c64 && metric && templ_energy > 0.0, "ofs_zc_freq_metric_fft: bad arguments");
    OFS_REQUIRE(in_dtype == OFS_C64 || in_dtype == OFS_IQ16, "ofs_zc_freq_metric_fft: complex64 or int16-IQ captures");
    OFS_REQUIRE(n_fft >= 2 && n_fft <= 2048 && cp >= 0, "ofs_zc_freq_metric_fft: n_fft must be 2..2048");
    OFS_REQUIRE(nbins >= 1 && nbins <= 64, "ofs_zc_freq_metric_fft: nbins <= 64");
    OFS_REQUIRE(n_branches >= 1 && n_branches <= 64, "ofs_zc_freq_metric_fft: 1..64 branches");
    const int64_t n_off = n - ((int64_t)n_fft + cp) + 1;
    OFS_REQUIRE(n_off > 0, "Received stream is shorter than a single OFDM symbol.");   /* zc_freq.py:76-78 */
    OFS_REQUIRE(out_stride >= n_off && n_frames >= 0, "ofs_zc_freq_metric_fft: out_stride < number of offsets");
    if (n_frames == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    keep_pool_cached();
    const int V = ZF8 - n_fft + 1;
    ZcFreqFftParams p{};
    p.blocks_per_cap = (int)((n_off + V - 1) / V);
    // consecutive blocks of a capture share one CTA (one anchor, float64 carry of E); captures are split only when there are
    // too few of them to fill the machine.  OFS_ZQF_BLOCKS_PER_ITEM overrides (tests exercise both the carry and the anchor).
    const int slots = 2 * sm_count();
    int64_t bpi = n_frames * p.blocks_per_cap / (4LL * slots);
    if (const char *e = getenv("OFS_ZQF_BLOCKS_PER_ITEM")) bpi = atoll(e);
    if (bpi < 1) bpi = 1;
    if (bpi > p.blocks_per_cap) bpi = p.blocks_per_cap;
    p.items_per_cap = (int)((p.blocks_per_cap + bpi - 1) / bpi);
    p.blocks_per_item = (p.blocks_per_cap + p.items_per_cap - 1) / p.items_per_cap;      // equal parts: 11 blocks in 2 items are 6 + 5, not 9 + 2
    p.n_items = n_frames * p.items_per_cap;
    const int grid = (int)(p.n_items < slots ? p.n_items : slots);
    AsyncBuf b_tw, b_tw8d, b_tw8f, b_ref, b_gp, b_rn, b_stash;          // returned to the pool on every exit path
    OFS_CUDA(b_tw.alloc((ZF / 2) * (sizeof(double2) + sizeof(float2)), stream));
    OFS_CUDA(b_tw8d.alloc(256 * sizeof(double2), stream));
    OFS_CUDA(b_tw8f.alloc(256 * sizeof(float2), stream));
    OFS_CUDA(b_ref.alloc(2 * (size_t)n_fft * sizeof(double2), stream));
    OFS_CUDA(b_gp.alloc(2 * ZF8 * sizeof(float2), stream));
    OFS_CUDA(b_rn.alloc(2 * sizeof(double), stream));
    OFS_CUDA(b_stash.alloc((size_t)grid * ZF8 * sizeof(float2), stream));
    double2 *tw = b_tw.as<double2>(), *tw8d = b_tw8d.as<double2>(), *ref = b_ref.as<double2>();
    float2 *tw8f = b_tw8f.as<float2>(), *Gp = b_gp.as<float2>(), *stash = b_stash.as<float2>();
    double *rn = b_rn.as<double>();
    zc_twiddle_kernel<<<(ZF / 2 + 255) / 256, 256, 0, stream>>>(tw);
    zc_twiddle8_kernel<<<1, 256, 0, stream>>>(tw8f, tw8d);
    zqf_ref_kernel<<<(n_fft + 127) / 128, 128, 0, stream>>>(bins, reinterpret_cast<const float2 *>(templ_c64), nbins, n_fft, ref);
    zqf_ref_kernel<<<(n_fft + 127) / 128, 128, 0, stream>>>(bins, nullptr, nbins, n_fft, ref + n_fft);
    if (int rc = check_launch("zqf_ref_kernel")) return rc;
    OFS_CUDA(cudaFuncSetAttribute(zc_spectrum8k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ZFP8 * sizeof(double2))));
    zc_spectrum8k_kernel<<<1, ZNT, ZFP8 * sizeof(double2), stream>>>(ref, n_fft, tw, tw8d, Gp, rn);
    zc_spectrum8k_kernel<<<1, ZNT, ZFP8 * sizeof(double2), stream>>>(ref + n_fft, n_fft, tw, tw8d, Gp + ZF8, rn + 1);
    if (int rc = check_launch("zc_spectrum8k_kernel")) return rc;
    count_launch(5);
    p.x = x; p.n = n; p.n_off = n_off; p.mstride = out_stride;
    p.N = n_fft; p.cp = cp; p.nbins = nbins; p.nb = n_branches; p.bins = bins; p.tw = tw; p.tw8 = tw8f; p.GpY = Gp; p.GpS = Gp + ZF8; p.stash = stash;
    p.templ_energy = (float)templ_energy; p.metric = metric;
    const size_t smem = (size_t)ZFP8 * sizeof(float2) + (size_t)(ZF8 + ZF8 / 32 + 8) * sizeof(float);
    auto kern = in_dtype == OFS_C64 ? zc_freq_fft_kernel<OFS_C64> : zc_freq_fft_kernel<OFS_IQ16>;
    OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, ZQF_THREADS, smem, stream>>>(p);
    return check_launch("zc_freq_fft_kernel");
}

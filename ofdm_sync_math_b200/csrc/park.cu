// Park metric (K3) -- park.py:64-114.
//   P(d) = sum_{k<h} x[d-k] * x[d+k]   (no conjugate; mirror correlation, NOT a sliding sum)
//   E(d) = sum_{k<h} |x[d+k]|^2,  d in [h, L-h-1],  M = |P|^2 / max(E,1e-12)^2, branches summed first.
// 1024 complex MACs per output: bound by the FMA pipe, not by HBM.  One CTA =
// 512 consecutive outputs of one frame; the 512+2h-1 samples it needs are staged in shared memory
// once; each thread produces PO=4 outputs and walks the lags PK=4 at a time from registers
// (14 smem loads per 16 complex MACs).  Accumulation: T (float for c64/iq16 input with float64
// flushes every 128 lags, double for c128).
#include "common.cuh"
#include "conv8k.cuh"
#include <cstdlib>

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
        // float32: the complex multiply-accumulate as two packed FFMA2 on two accumulator pairs per output (ParkAcc below: the
        // broadcasts and the swizzle are operand modifiers, so no register moves -- the packed loop tried in round 1 built its
        // operand pairs with moves and was slower); float64 keeps the scalar form
        float2 pA[QO], pB[QO];
#pragma unroll
        for (int o = 0; o < QO; ++o) pA[o] = pB[o] = make_float2(0.f, 0.f);
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
                    if constexpr (sizeof(T) == 4) {
                        // sum u v = A + (-1, 1) (.) B,  A += u (.) (v.x, v.x),  B += (u.y, u.x) (.) (v.y, v.y)
                        pA[o] = __ffma2_rn(make_float2(u.x, u.y), make_float2(v.x, v.x), pA[o]);
                        pB[o] = __ffma2_rn(make_float2(u.y, u.x), make_float2(v.y, v.y), pB[o]);
                    } else {
                        pr[o] = fma(u.x, v.x, pr[o]); pr[o] = fma(-u.y, v.y, pr[o]);
                        pi[o] = fma(u.x, v.y, pi[o]); pi[o] = fma(u.y, v.x, pi[o]);
                    }
                }
            }
            if (sizeof(T) == 4 && ((k0 + QK) % 128 == 0)) {   // bounded fp32 error: flush partial sums
#pragma unroll
                for (int o = 0; o < QO; ++o) {
                    accr[o] += (double)(pA[o].x - pB[o].x); acci[o] += (double)(pA[o].y + pB[o].y);
                    pA[o] = pB[o] = make_float2(0.f, 0.f);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < QO; ++o) {
            accr[o] += (double)pr[o] + (double)(pA[o].x - pB[o].x);
            acci[o] += (double)pi[o] + (double)(pA[o].y + pB[o].y);
        }
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

// ---- K3b: the Park metric by block FFTs (float32; h a multiple of 128, h <= 1024) ---------------------------------------
// P(d) = sum_{k<h} x[d-k] x[d+k] is half of a BAND-LIMITED self-convolution: with C(2d) = sum_{i+j=2d, |j-i|<2h} x[i] x[j]
// (ordered pairs), C(2d) = 2 P(d) - x[d]^2.  i + j even means i and j have the same parity, so with the half-rate sequences
// xe[p] = x[2p], xo[p] = x[2p+1]:  C(2d) = Cb_e(d) + Cb_o(d-1),  Cb_y(n) = sum_{p+q=n, |q-p|<h} y[p] y[q]  (no parity waste).
// Cut y into blocks of B = 128 and let m = h / B.  A block pair (I, J) lies entirely inside the band when |J - I| <= m - 1 and
// contributes the plain linear convolution y_I * y_J to the outputs n in [(I+J) B, (I+J+2) B); pairs with |J - I| = m are cut
// by the band edge to the strict triangle q' < p' (local indices), pairs further apart contribute nothing.  Hence, per
// anti-diagonal S = I + J:   Z_S = sum_{I+J=S, |J-I|<m} X_I X_J   (X = 256-point FFT of the zero-padded block: the products are
// summed in the FREQUENCY domain and ONE inverse transform per anti-diagonal brings them back), plus the edge triangle of the
// one pair with |J - I| = m directly.  ~700 flops per output instead of 8192.
// One CTA = 1536 consecutive outputs of one frame.  The 8192-element array of conv8k.cuh serves as 32 independent 256-point
// transforms (stage B + half of stage C): 16 blocks of samples and 14 anti-diagonals per parity pass.  Checked against the
// float64 oracle and against the direct kernel (tests/test_gpu_park.py); OFS_PARK_DIRECT=1 forces the direct kernel.
constexpr int PF_B = 128, PF_T = 12, PF_OUT = PF_T * PF_B, PF_NZ = PF_T + 2, PF_NBLK = 16, PF_R = PF_OUT / ZNT;
constexpr int PF_BS = 272;                 // a 256-block and its padding (one element per 16) in the array of conv8k.cuh
// Complex multiply-accumulate split over two packed accumulators, two FFMA2 per term and nothing else:
//   sum a w = A + (-1, 1) (.) B,   A += a (.) (w.x, w.x),   B += (a.y, a.x) (.) (w.y, w.y)
// (the broadcasts and the swizzle are operand modifiers; the sign pattern is applied once, after the loop)
struct ParkAcc {
    float2 A, B;
    __device__ __forceinline__ void zero() { A = B = make_float2(0.f, 0.f); }
    __device__ __forceinline__ void mac(float2 a, float2 w)
    {
        A = __ffma2_rn(a, make_float2(w.x, w.x), A);
        B = __ffma2_rn(make_float2(a.y, a.x), make_float2(w.y, w.y), B);
    }
    __device__ __forceinline__ float2 value() const { return __ffma2_rn(B, make_float2(-1.f, 1.f), A); }
};
// Output r of thread t inside the tile (d - D0).  The edge triangle of an output at offset u inside its anti-diagonal has
// min(u, 255 - u) / 2 terms, and a plain t + 256 r ownership would give a thread the same u six times (warps around u = 128
// doing eight times the work of those around 0): rotate the assignment by 43 from one group of 256 outputs to the next.
__device__ __forceinline__ int park_fft_own(int t, int r) { return ((t + 43 * r) & (ZNT - 1)) + ZNT * r; }

__global__ void park_twiddle_kernel(double2 *tw)         // the table layout of fft4096.cuh: double2[2048] then float2[2048]
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ZF / 2) {
        double sn, c;
        sincospi(-2.0 * (double)i / (double)ZF, &sn, &c);
        tw[i] = make_double2(c, sn);
        reinterpret_cast<float2 *>(tw + ZF / 2)[i] = make_float2((float)c, (float)sn);
    }
}

// MC: h / 128 as a compile-time constant (8 for the scripts' N_FFT = 2048; 0: run-time) -- with it every block index of the
// frequency-domain products is a constant and the 16 spectra of a position are read into registers once
template <int DT, int MC>
__global__ void __launch_bounds__(ZNT, 2) park_fft_kernel(const void *x, int nb, int64_t L, int64_t xfs, int64_t xbs, int h, int64_t n_out,
                                                         int64_t out_stride, float *M, float2 *P, float *E, int tiles_per_frame,
                                                         const double2 *tw)
{
    using In = typename InT<DT>::type;
    extern __shared__ __align__(16) unsigned char psm[];
    float2 *a = reinterpret_cast<float2 *>(psm);              // 32 blocks of 256 (padded): 16 sample blocks, 14 anti-diagonals
    float2 *yc = a + ZFP8;                                    // [16][128] the loaded blocks in the time domain (edge triangles)
    __shared__ double wsum[ZNT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t frame = blockIdx.x / tiles_per_frame;
    const int tile = blockIdx.x % tiles_per_frame;
    const int m = MC ? MC : h / PF_B;
    const int S0 = tile * PF_T;                               // first anti-diagonal block of the tile (even)
    const int64_t D0 = (int64_t)S0 * PF_B;                    // first centre d of the tile (d = h + output index)
    const int I_lo = (S0 - 2 - m) >> 1;                       // first sample block loaded (floor; blocks before the frame are zero)
    float2 accP[PF_R];
#pragma unroll
    for (int r = 0; r < PF_R; ++r) accP[r] = make_float2(0.f, 0.f);

    for (int b = 0; b < nb; ++b) {
        const In *xb = reinterpret_cast<const In *>(x) + frame * xfs + (int64_t)b * xbs;
        for (int par = 0; par < 2; ++par) {
            __syncthreads();
            {
                // thread = sample pp (+128) of blocks s0, s0 + 2, ...: 8 independent loads in flight, then the stores
                const int pp = tid & (PF_B - 1), s0 = tid >> 7;
                float2 v[PF_NBLK / 2];
#pragma unroll
                for (int k = 0; k < PF_NBLK / 2; ++k) {
                    const int64_t p = (int64_t)(I_lo + s0 + 2 * k) * PF_B + pp, j = 2 * p + par;
                    v[k] = make_float2(0.f, 0.f);
                    if (p >= 0 && j < L) { const In t = xb[j]; v[k] = make_float2((float)t.x, (float)t.y); }
                }
#pragma unroll
                for (int k = 0; k < PF_NBLK / 2; ++k) {
                    const int sl = s0 + 2 * k;
                    a[sl * PF_BS + pp + (pp >> 4)] = v[k];
                    a[sl * PF_BS + PF_B + pp + ((PF_B + pp) >> 4)] = make_float2(0.f, 0.f);
                    yc[sl * PF_B + pp] = v[k];
                }
            }
            const pk::Seeds sbd = conv8k_seeds_bd(tw);
            __syncthreads();
            conv8k_stage_b<true, false>(a, sbd);                        // the sample blocks (first half of the array) forward
            __syncthreads();
            conv8k_stage_c_fwd<true, false>(a);
            __syncthreads();
            // anti-diagonal sums in the frequency domain: thread = one of the 256 spectrum positions
            {
                // block s of the first half starts at s * 272 (256 + its padding); the anti-diagonals live in the second half
                const float2 *xs0 = a + tid + (tid >> 4);
                float2 *zs0 = a + ZFP + tid + (tid >> 4);
                if constexpr (MC == 8) {
                    // S0 is even and I_lo = S0 / 2 - 5: anti-diagonal zs pairs the blocks 4 + zs / 2 - k and 4 + (zs + 1) / 2 + k,
                    // k = 0 .. 3 -- all compile-time, so the 16 spectra of this position are loaded once (16 LDS instead of 112)
                    float2 X[PF_NBLK];
#pragma unroll
                    for (int sl = 0; sl < PF_NBLK; ++sl) X[sl] = xs0[sl * PF_BS];
#pragma unroll
                    for (int zs = 0; zs < PF_NZ; ++zs) {
                        ParkAcc pa;
                        pa.zero();
#pragma unroll
                        for (int k = 0; k < 4; ++k) pa.mac(X[4 + zs / 2 - k], X[4 + (zs + 1) / 2 + k]);
                        float2 acc = pa.value();
                        acc = pk::add(acc, acc);
                        if (!(zs & 1)) acc = pk::sub(acc, pk::mul(X[4 + zs / 2], X[4 + zs / 2]));
                        zs0[zs * PF_BS] = __fmul2_rn(acc, make_float2(1.0f / 256.0f, 1.0f / 256.0f));
                    }
                } else
                for (int zs = 0; zs < PF_NZ; ++zs) {
                    // pairs (c_lo - k, c_hi + k), k = 0 .. kmax: J - I = 2 k + (S & 1) <= m - 1; every pair but the diagonal one
                    // (k = 0 of an even S) also stands for its mirror image (J, I):  Z = 2 sum - [S even] X_c^2
                    const int S = S0 - 2 + zs;
                    const int c_lo = S >> 1, kmax = (m - 1 - (S & 1)) >> 1;
                    const float2 *pi = xs0 + (c_lo - I_lo) * PF_BS, *pj = xs0 + (S - c_lo - I_lo) * PF_BS;
                    const float2 xc = *pi;
                    ParkAcc pa;
                    pa.zero();
#pragma unroll 4
                    for (int k = 0; k <= kmax; ++k, pi -= PF_BS, pj += PF_BS) pa.mac(*pi, *pj);
                    float2 acc = pa.value();
                    acc = pk::add(acc, acc);
                    if (!(S & 1) && kmax >= 0) acc = pk::sub(acc, pk::mul(xc, xc));
                    zs0[zs * PF_BS] = __fmul2_rn(acc, make_float2(1.0f / 256.0f, 1.0f / 256.0f));   // 1/256 of the inverse transform
                }
            }
            __syncthreads();
            conv8k_stage_c_inv<false, true>(a);                         // the anti-diagonals (second half) back
            __syncthreads();
            conv8k_stage_d<false, true>(a, sbd);
            __syncthreads();
            // overlap-add of the two anti-diagonals that cover an output, plus the edge triangle of the one pair with J - I = m
#pragma unroll
            for (int r = 0; r < PF_R; ++r) {
                const int64_t n = D0 + park_fft_own(tid, r) - par;          // index into Cb_y
                if (n < 0) continue;
                const int S1 = (int)((uint64_t)n >> 7), u1 = (int)(n & (PF_B - 1));       // n >= 0; PF_B = 128
                const float2 *zb = a + ZFP + (S1 - S0 + 1) * PF_BS;        // anti-diagonal S1 - 1; S1 is the next block
                const float2 z0 = zb[u1 + PF_B + ((u1 + PF_B) >> 4)], z1 = zb[PF_BS + u1 + (u1 >> 4)];
                const int Se = ((S1 - m) & 1) ? S1 - 1 : S1;
                const int u = (int)(n - (int64_t)Se * PF_B);                   // 0 .. 255
                const int I = (Se - m) >> 1;
                const float2 *yi = yc + (I - I_lo) * PF_B + u, *yj = yc + (I + m - I_lo) * PF_B;
                const int q_lo = u > PF_B - 1 ? u - (PF_B - 1) : 0, q_hi = (u + 1) >> 1;   // q' < p' = u - q' < B
                ParkAcc e0, e1;                                               // two independent chains
                e0.zero(); e1.zero();
                int q = q_lo;
                for (; q + 4 <= q_hi; q += 4) {
                    e0.mac(yi[-q], yj[q]);
                    e1.mac(yi[-(q + 1)], yj[q + 1]);
                    e0.mac(yi[-(q + 2)], yj[q + 2]);
                    e1.mac(yi[-(q + 3)], yj[q + 3]);
                }
                for (; q < q_hi; ++q) e0.mac(yi[-q], yj[q]);
                const float2 e = pk::add(e0.value(), e1.value());
                accP[r] = pk::add(accP[r], pk::add(pk::add(z1, z0), pk::add(e, e)));
            }
        }
    }
    // ---- energies E(d) = sum_{k<h} |x[d+k]|^2 over the branches: float64 prefix of the tile's |x|^2 in shared memory ----------
    __syncthreads();
    double *pre = reinterpret_cast<double *>(psm);             // pre[k] = sum of |x[D0 + j]|^2, j < k
    const int cnt = PF_OUT + h;
    const int per = (cnt + ZNT - 1) / ZNT;
    {
        const int k0 = tid * per, k1 = k0 + per < cnt ? k0 + per : cnt;
        float en[10];                                          // per <= (1536 + 1024) / 256 = 10; |x|^2 in float32 (exact for int16
#pragma unroll                                                 // IQ, 6e-8 relative for complex64), summed in float64
        for (int q = 0; q < 10; ++q) en[q] = 0.f;
        for (int b = 0; b < nb; ++b) {
            const In *xr = reinterpret_cast<const In *>(x) + frame * xfs + (int64_t)b * xbs + D0 + k0;
#pragma unroll
            for (int q = 0; q < 10; ++q)
                if (k0 + q < k1 && D0 + k0 + q < L) {
                    const In t = xr[q];
                    const float tx = (float)t.x, ty = (float)t.y;
                    en[q] += fmaf(tx, tx, ty * ty);
                }
        }
        double run = 0.0;
#pragma unroll
        for (int q = 0; q < 10; ++q)
            if (k0 + q < k1) { run += (double)en[q]; pre[k0 + q + 1] = run; }
        double t = run;
        for (int o = 1; o < 32; o <<= 1) { const double y = shfl_up_f64(t, o); if (lane >= o) t += y; }
        if (lane == 31) wsum[warp] = t;
        __syncthreads();
        double off = t - run;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        for (int k = k0; k < k1; ++k) pre[k + 1] += off;
        if (tid == 0) pre[0] = 0.0;
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < PF_R; ++r) {
        const int kd = park_fft_own(tid, r);                   // d - D0
        const int64_t d = D0 + kd, i = d - h;
        if (i < 0 || i >= n_out) continue;
        float2 xs = make_float2(0.f, 0.f);                     // sum over branches of x[d]^2
        for (int b = 0; b < nb; ++b) {
            const In t = (reinterpret_cast<const In *>(x) + frame * xfs + (int64_t)b * xbs)[d];
            const float2 v = make_float2((float)t.x, (float)t.y);
            xs = pk::add(xs, pk::mul(v, v));
        }
        const float pr = 0.5f * (accP[r].x + xs.x), pi = 0.5f * (accP[r].y + xs.y);
        const float en = (float)(pre[kd + h] - pre[kd]);
        const float ee = fmaxf(en, 1e-12f);                    // park.py:111
        const int64_t oi = frame * out_stride + i;
        const float re = 1.0f / ee, qr = pr * re, qi = pi * re;  // |P / E|^2: no intermediate leaves the float32 range
        if (M) M[oi] = fmaf(qr, qr, qi * qi);
        if (P) P[oi] = make_float2(pr, pi);
        if (E) E[oi] = en;
    }
}

template <int DT>
static int launch_park_fft(const ofs_metric_desc *d, const void *x, void *M, void *P, void *E, int64_t n_out, cudaStream_t st)
{
    const int h = d->symbol_len / 2;
    const int tiles = (int)((h + n_out + PF_OUT - 1) / PF_OUT);
    const int64_t grid = (int64_t)tiles * d->n_frames;
    OFS_REQUIRE(grid < (1LL << 31), "ofs_park_metric: grid too large");
    AsyncBuf b_tw;                                                        // returned to the pool on every exit path
    OFS_CUDA(b_tw.alloc((ZF / 2) * (sizeof(double2) + sizeof(float2)), st));
    double2 *tw = b_tw.as<double2>();
    park_twiddle_kernel<<<(ZF / 2 + 255) / 256, 256, 0, st>>>(tw);
    if (int rc = check_launch("park_twiddle_kernel")) return rc;
    const size_t smem = (size_t)ZFP8 * sizeof(float2) + (size_t)PF_NBLK * PF_B * sizeof(float2);
    auto kern = h == 8 * PF_B ? park_fft_kernel<DT, 8> : park_fft_kernel<DT, 0>;
    OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, ZNT, smem, st>>>(x, d->n_branches, d->n_samples, d->x_frame_stride, d->x_branch_stride, h, n_out, d->out_stride,
                                           (float *)M, (float2 *)P, (float *)E, tiles, tw);
    return check_launch("park_fft_kernel");
}

template <typename T, int DT>
static int launch_park(const ofs_metric_desc *d, const void *x, void *M, void *P, void *E, int64_t n_out, cudaStream_t st)
{
    const int h = d->symbol_len / 2;
    if constexpr (sizeof(T) == 4) {
        // float32 outputs, h a multiple of 128 up to 1024 (the scripts' N_FFT = 2048): block FFTs; OFS_PARK_DIRECT=1: direct form
        const char *env = getenv("OFS_PARK_DIRECT");
        const bool force_direct = env && atoi(env) != 0;
        if (!d->out_f64 && !force_direct && h % PF_B == 0 && h / PF_B >= 1 && h / PF_B <= 8)
            return launch_park_fft<DT>(d, x, M, P, E, n_out, st);
    }
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

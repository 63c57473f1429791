// Hand-written 4096-point FFT shared by the matched filter (zc.cu) and the pilot / data demodulator (rxchain.cu).
#pragma once
#include "common.cuh"

namespace ofs {

constexpr int ZF = 4096, ZLOG = 12, ZNT = 256;

template <typename T> struct C2T;
template <> struct C2T<float> { using type = float2; };
template <> struct C2T<double> { using type = double2; };

template <typename C2> __device__ __forceinline__ C2 cadd(C2 a, C2 b) { C2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C2> __device__ __forceinline__ C2 csub(C2 a, C2 b) { C2 r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C2> __device__ __forceinline__ C2 cmul(C2 a, C2 b) { C2 r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <typename C2> __device__ __forceinline__ C2 cmulc(C2 a, C2 b) { C2 r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r; }  // a * conj(b)

// twiddle tables: double2[ZF/2] followed by float2[ZF/2] (same buffer), exp(-2 pi i k / ZF)
template <typename C2>
__device__ __forceinline__ C2 ld_tw(const double2 *tw, int i);
template <>
__device__ __forceinline__ double2 ld_tw<double2>(const double2 *tw, int i) { return __ldg(tw + i); }
template <>
__device__ __forceinline__ float2 ld_tw<float2>(const double2 *tw, int i)
{
    return __ldg(reinterpret_cast<const float2 *>(tw + ZF / 2) + i);
}

// ---- 4096-point FFT = three passes of radix-16 butterflies held in registers ------------------------------------
// 256 threads x 16 elements; between the passes the data is transposed through shared memory (two round trips per
// transform instead of the twelve of a radix-2 ladder).  The array is padded by one element every 16 (index i lives at
// i + i/16), which makes the stride-1, stride-16 and stride-256 access patterns of the three passes all conflict-free.
// Forward: natural order in -> digit-reversed out (X[k0 + 16 k1 + 256 k2] at position 256 k0 + 16 k1 + k2); the inverse
// runs the same flow graph backwards, so the pointwise product with the (equally permuted) filter spectrum needs no
// reordering pass.
constexpr int ZFP = ZF + ZF / 16;
__device__ __forceinline__ int zpad(int i) { return i + (i >> 4); }

template <typename C2>
__device__ __forceinline__ C2 tw4096(const double2 *tw, int i)          // exp(-2 pi i * i / 4096), 0 <= i < 4096
{
    C2 w = ld_tw<C2>(tw, i & (ZF / 2 - 1));
    if (i & (ZF / 2)) { w.x = -w.x; w.y = -w.y; }
    return w;
}

// 16-point DFT in registers (radix-2 DIF, 4 stages, constants folded), natural order in and out.  INV: conjugate kernel.
template <typename C2, bool INV>
__device__ __forceinline__ void dft16(C2 (&v)[16])
{
    using T = decltype(v[0].x);
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, r2 = (T)0.70710678118654752440;
    // W16^j, j = 0..7 (forward: exp(-2 pi i j / 16))
    const T wr[8] = {(T)1, c1, r2, s1, (T)0, -s1, -r2, -c1};
    const T wi[8] = {(T)0, -s1, -r2, -c1, (T)-1, -c1, -r2, -s1};
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((i & h) == 0) {
                const int j = (i & (h - 1)) * (8 / h);         // twiddle exponent in units of W16
                const C2 u = v[i], w = v[i + h];
                v[i].x = u.x + w.x; v[i].y = u.y + w.y;
                const T dx = u.x - w.x, dy = u.y - w.y;
                if (j == 0) { v[i + h].x = dx; v[i + h].y = dy; }
                else if (j == 4) { v[i + h].x = INV ? -dy : dy; v[i + h].y = INV ? dx : -dx; }
                else {
                    const T cr = wr[j], ci = INV ? -wi[j] : wi[j];
                    v[i + h].x = dx * cr - dy * ci;
                    v[i + h].y = dx * ci + dy * cr;
                }
            }
        }
    }
    // bit-reversed -> natural
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);
        if (i < r) { const C2 t = v[i]; v[i] = v[r]; v[r] = t; }
    }
}

// w[q] = w1^q for q = 1..15 from ONE table load: squarings and products, depth 4 (error <= 4 roundings) -- the 15 scattered
// table loads per pass were the long-scoreboard stalls of the kernel; the FMA pipe has room.
template <typename C2>
__device__ __forceinline__ void tw_powers(C2 w1, C2 (&w)[16])
{
    w[1] = w1;
    w[2] = cmul(w1, w1);
    w[3] = cmul(w[2], w1);
    w[4] = cmul(w[2], w[2]);
    w[5] = cmul(w[4], w1); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]);
    w[8] = cmul(w[4], w[4]);
#pragma unroll
    for (int q = 9; q < 16; ++q) w[q] = cmul(w[8], w[q - 8]);
}

// natural order in -> digit-reversed order out
template <typename C2>
__device__ void fft_dif(C2 *a, const double2 *tw)
{
    const int t = threadIdx.x;
    C2 v[16];
    // pass 1: stride 256, twiddle W4096^(t k0)
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = a[zpad(q * 256 + t)];
    dft16<C2, false>(v);
    C2 w[16];
    tw_powers<C2>(tw4096<C2>(tw, t), w);
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], w[q]);
#pragma unroll
    for (int q = 0; q < 16; ++q) a[zpad(q * 256 + t)] = v[q];
    __syncthreads();
    // pass 2: inside each block of 256, stride 16, twiddle W256^(n0 k1)
    const int k0 = t >> 4, n0 = t & 15;
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = a[zpad(k0 * 256 + q * 16 + n0)];
    dft16<C2, false>(v);
    tw_powers<C2>(tw4096<C2>(tw, 16 * n0), w);
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], w[q]);
#pragma unroll
    for (int q = 0; q < 16; ++q) a[zpad(k0 * 256 + q * 16 + n0)] = v[q];
    __syncthreads();
    // pass 3: 16 consecutive elements
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = a[t * 17 + q];
    dft16<C2, false>(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) a[t * 17 + q] = v[q];
    __syncthreads();
}
// digit-reversed order in -> natural order out, unscaled (multiply by 1/ZF afterwards)
template <typename C2>
__device__ void ifft_dit(C2 *a, const double2 *tw)
{
    const int t = threadIdx.x;
    C2 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = a[t * 17 + q];
    dft16<C2, true>(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) a[t * 17 + q] = v[q];
    __syncthreads();
    const int k0 = t >> 4, n0 = t & 15;
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = a[zpad(k0 * 256 + q * 16 + n0)];
    C2 w[16];
    tw_powers<C2>(tw4096<C2>(tw, 16 * n0), w);
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmulc(v[q], w[q]);
    dft16<C2, true>(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) a[zpad(k0 * 256 + q * 16 + n0)] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = a[zpad(q * 256 + t)];
    tw_powers<C2>(tw4096<C2>(tw, t), w);
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmulc(v[q], w[q]);
    dft16<C2, true>(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) a[zpad(q * 256 + t)] = v[q];
    __syncthreads();
}

// ---- 8192-point FFT = one radix-2 stage + two 4096-point FFTs -----------------------------------------------------------
// DIF: y0[n] = x[n] + x[n + 4096], y1[n] = (x[n] - x[n + 4096]) W8192^n, then X[2m] = FFT4096(y0)[m], X[2m + 1] = FFT4096(y1)[m].
// The array is two padded 4096-blocks (ZFP elements each); output order = the 4096-point digit reversal inside each block.
// The inverse runs the graph backwards (two unscaled inverse 4096-point transforms, then x[n], x[n + 4096] =
// E[n] +- conj(W8192^n) O[n]; multiply by 1 / 8192 afterwards).  W8192^(t + 256 q) = W8192^t W32^q: one table entry per
// thread (tw8: float2[256] / double2[256]) and compile-time constants.
constexpr int ZF8 = 8192;
constexpr int ZFP8 = 2 * ZFP;
__device__ __forceinline__ int zpad8(int i) { return (i >> 12) * ZFP + zpad(i & (ZF - 1)); }

template <typename C2>
__device__ __forceinline__ C2 w32_const(int q)        // exp(-2 pi i q / 32), q < 16 (compile-time after unrolling)
{
    using T = decltype(C2{}.x);
    constexpr double c[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                              0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785, 0.0, -0.19509032201612826785,
                              -0.38268343236508977173, -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                              -0.92387953251128675613, -0.98078528040323044913};
    constexpr double sn[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474, 0.70710678118654752440,
                               0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913, 1.0, 0.98078528040323044913,
                               0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440, 0.55557023301960222474,
                               0.38268343236508977173, 0.19509032201612826785};
    C2 w; w.x = (T)c[q]; w.y = (T)(-sn[q]);
    return w;
}

template <typename C2>
__device__ void fft8k_dif(C2 *a, const double2 *tw, C2 w0 /* exp(-2 pi i threadIdx.x / 8192) */)
{
    const int t = threadIdx.x;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = zpad(t + 256 * q);
        const C2 u = a[n], v = a[ZFP + n];
        a[n] = cadd(u, v);
        const C2 d = csub(u, v);
        a[ZFP + n] = q == 0 ? cmul(d, w0) : cmul(cmul(d, w0), w32_const<C2>(q));
    }
    __syncthreads();
    fft_dif<C2>(a, tw);
    fft_dif<C2>(a + ZFP, tw);
}
template <typename C2>
__device__ void ifft8k_dit(C2 *a, const double2 *tw, C2 w0)
{
    ifft_dit<C2>(a, tw);
    ifft_dit<C2>(a + ZFP, tw);
    const int t = threadIdx.x;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = zpad(t + 256 * q);
        const C2 u = a[n];
        C2 v = cmulc(a[ZFP + n], w0);
        if (q != 0) v = cmulc(v, w32_const<C2>(q));
        a[n] = cadd(u, v);
        a[ZFP + n] = csub(u, v);
    }
    __syncthreads();
}

}  // namespace ofs

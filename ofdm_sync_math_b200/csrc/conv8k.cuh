// Fused 8192-point overlap-save convolution pipeline (float32) for the Zadoff-Chu matched filter (zc.py:115-126,
// zc_v2.py:244-271), the FFT form of zc_freq.compute_frequency_metric (zc_freq.py:62-99) and -- as 32 independent 256-point
// transforms -- the block-FFT Park metric (park.py:64-114, park.cu).  Header-only, no barriers inside: the kernels place
// their own __syncthreads between the stages, so tests/host/conv8k_host_test.cpp can run the same code on the CPU.
//
// Same flow graph as fft8k_dif / ifft8k_dit of fft4096.cuh (radix-2 split, then two 4096-point transforms of three radix-16
// passes), but
//  * the passes are paired so that a block crosses shared memory 4 times instead of 10:
//      A  global -> registers: radix-2 stage + first radix-16 pass of both halves                        -> shared
//      B  second forward pass (stride 16)                                                         shared -> shared
//      C  third forward pass, pointwise product with the filter spectrum, first inverse pass,
//         all three on the same 16 consecutive elements held in registers                          shared -> shared
//      D  second inverse pass                                                                     shared -> shared
//      E  third inverse pass + inverse radix-2 stage in registers, results handed to the caller's sink (no store)
//  * all complex arithmetic is packed fp32 (FADD2 / FMUL2 / FFMA2 on (re, im) register pairs, operand swizzles and scalar
//    broadcasts are free): a 16-point DFT is 82 instructions instead of 168, a complex multiply 2-3 instead of 4.  The kernels
//    built on this are issue-bound, so halving the instruction count is what pays.
// The filter spectrum arrives pre-permuted to stage C's register order and pre-scaled by 1/8192 (zc_spectrum8k_kernel).
#pragma once
#include "fft4096.cuh"

namespace ofs {
namespace pk {
__device__ __forceinline__ float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 add_i(float2 a, float2 b) { return __ffma2_rn(make_float2(b.y, b.x), make_float2(-1.f, 1.f), a); }   // a + i b
__device__ __forceinline__ float2 sub_i(float2 a, float2 b) { return __ffma2_rn(make_float2(b.y, b.x), make_float2(1.f, -1.f), a); }   // a - i b
// a * (cr + i ci), compile-time constants
__device__ __forceinline__ float2 mulk(float2 a, float cr, float ci)
{
    return __ffma2_rn(a, make_float2(cr, cr), __fmul2_rn(make_float2(a.y, a.x), make_float2(-ci, ci)));
}
template <bool INV> __device__ __forceinline__ float2 mul_mi(float2 a)      // forward: a * (-i);  inverse: a * (+i)
{
    return __fmul2_rn(make_float2(a.y, a.x), INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f));
}
// a * w and a * conj(w), run-time w: (a.x w.x - a.y w.y, a.y w.x + a.x w.y) = a * w.x + (a.y, a.x) * (-w.y, w.y)
__device__ __forceinline__ float2 mul(float2 a, float2 w)
{
    const float2 s = __fmul2_rn(make_float2(w.y, w.y), make_float2(-1.f, 1.f));
    return __ffma2_rn(a, make_float2(w.x, w.x), __fmul2_rn(make_float2(a.y, a.x), s));
}
__device__ __forceinline__ float2 mulc(float2 a, float2 w)
{
    const float2 s = __fmul2_rn(make_float2(w.y, w.y), make_float2(1.f, -1.f));
    return __ffma2_rn(a, make_float2(w.x, w.x), __fmul2_rn(make_float2(a.y, a.x), s));
}

template <bool INV> __device__ __forceinline__ void r4(float2 &a, float2 &b, float2 &c, float2 &d)
{
    const float2 t0 = add(a, c), t1 = sub(a, c), t2 = add(b, d), t3 = sub(b, d);
    a = add(t0, t2);
    c = sub(t0, t2);
    b = INV ? add_i(t1, t3) : sub_i(t1, t3);
    d = INV ? sub_i(t1, t3) : add_i(t1, t3);
}
// 16-point DFT in registers, natural order in and out: 4 x 4 decomposition (n = 4 n1 + n2, k = k1 + 4 k2), two layers of
// radix-4 butterflies with the W16^(n2 k1) twiddles between them.  INV: conjugate kernel (unscaled).
template <bool INV> __device__ __forceinline__ void dft16(float2 (&v)[16])
{
    constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, r2 = 0.70710678118654752440f;
    constexpr float sg = INV ? 1.f : -1.f;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) r4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);     // over n1: A[n2][k1] at v[4 k1 + n2]
    v[5] = mulk(v[5], c1, sg * s1);          // W16^1
    v[6] = mulk(v[6], r2, sg * r2);          // W16^2
    v[7] = mulk(v[7], s1, sg * c1);          // W16^3
    v[9] = mulk(v[9], r2, sg * r2);          // W16^2
    v[10] = mul_mi<INV>(v[10]);              // W16^4
    v[11] = mulk(v[11], -r2, sg * r2);       // W16^6
    v[13] = mulk(v[13], s1, sg * c1);        // W16^3
    v[14] = mulk(v[14], -r2, sg * r2);       // W16^6
    v[15] = mulk(v[15], -c1, -sg * s1);      // W16^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) r4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // over n2: X[k1 + 4 k2] at v[4 k1 + k2]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) { const float2 t = v[4 * i + j]; v[4 * i + j] = v[4 * j + i]; v[4 * j + i] = t; }
}
// w[q] = W4096^(base q), q = 1..15 (base * 15 < 4096): w1, w2, w4, w8 from the float table (correctly rounded; their lane
// strides are the short ones, 30 L1 wavefronts per warp), the rest by one to three multiplications.  All 15 from the table is
// the most accurate but costs 240 scattered wavefronts per pass (measured: matched filter 1.49 -> 1.88 ms); all from w1 by
// squaring (depth 4) doubles the transform's rounding error, which the FFT-form zc_freq kernel multiplies by burst-sized samples.
// The four table loads are a separate step so that a kernel can issue them BEFORE the barrier that precedes the stage: right
// after a barrier every warp of the CTA would wait for them at once.
struct Seeds { float2 w1, w2, w4, w8; };
__device__ __forceinline__ Seeds seeds(const double2 *tw, int base)
{
    Seeds s;
    s.w1 = tw4096<float2>(tw, base);
    s.w2 = tw4096<float2>(tw, 2 * base);
    s.w4 = tw4096<float2>(tw, 4 * base);
    s.w8 = tw4096<float2>(tw, 8 * base);
    return s;
}
__device__ __forceinline__ void powers(const Seeds &s, float2 (&w)[16])
{
    w[1] = s.w1; w[2] = s.w2; w[4] = s.w4; w[8] = s.w8;
    w[3] = mul(w[2], w[1]);
    w[5] = mul(w[4], w[1]); w[6] = mul(w[4], w[2]); w[7] = mul(w[4], w[3]);
#pragma unroll
    for (int q = 9; q < 16; ++q) w[q] = mul(w[8], w[q - 8]);
}
}  // namespace pk

// filter spectrum in stage-C order: element (half h, thread t, q) at Gp[((h * 8 + q / 2) * 256 + t) * 2 + (q & 1)],
// value = G8[h * 4096 + 16 t + q] / 8192 (G8 = transform order of fft8k_dif): one coalesced 16-byte load per two elements
__host__ __device__ __forceinline__ int conv8k_gidx(int h, int t, int q) { return ((h * 8 + (q >> 1)) * 256 + t) * 2 + (q & 1); }

__device__ __forceinline__ float2 conv8k_w32(int q)      // exp(-2 pi i q / 32), compile-time after unrolling
{
    return w32_const<float2>(q);
}

// Stage A.  ld(m) returns local sample m of the block (0 <= m < 8192, zero outside the capture); en(m, |x|^2) receives
// every sample's energy (for the sliding-energy normalisation) -- pass a no-op when it is not needed.
template <class Load, class Energy>
__device__ __forceinline__ void conv8k_stage_a(float2 *a, const pk::Seeds &sd, float2 w0, Load ld, Energy en)
{
    const int t = threadIdx.x;
    float2 y0[16], y1[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = t + 256 * q;
        const float2 u = ld(n), v = ld(n + ZF);
        en(n, fmaf(u.x, u.x, u.y * u.y));
        en(n + ZF, fmaf(v.x, v.x, v.y * v.y));
        y0[q] = pk::add(u, v);
        float2 d = pk::mul(pk::sub(u, v), w0);                   // W8192^(t + 256 q) = W8192^t W32^q
        if (q != 0) { const float2 k = conv8k_w32(q); d = pk::mulk(d, k.x, k.y); }
        y1[q] = d;
    }
    float2 w[16];
    pk::powers(sd, w);
    pk::dft16<false>(y0);
#pragma unroll
    for (int q = 0; q < 16; ++q) a[zpad(q * 256 + t)] = q ? pk::mul(y0[q], w[q]) : y0[0];
    pk::dft16<false>(y1);
#pragma unroll
    for (int q = 0; q < 16; ++q) a[ZFP + zpad(q * 256 + t)] = q ? pk::mul(y1[q], w[q]) : y1[0];
}

// Stage B: second forward pass of both halves (inside each block of 256: stride 16, twiddle W256^(n0 k1))
// (H0 / H1: which halves of the array to process -- the Park kernel transforms its sample blocks, which live in the first half,
// forward only and its anti-diagonals, in the second half, backward only)
template <bool H0 = true, bool H1 = true>
__device__ __forceinline__ void conv8k_stage_b(float2 *a, const pk::Seeds &sd)
{
    const int t = threadIdx.x, k0 = t >> 4, n0 = t & 15;
    float2 w[16];
    pk::powers(sd, w);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if ((h == 0 && !H0) || (h == 1 && !H1)) continue;
        float2 *ah = a + h * ZFP + zpad(k0 * 256 + n0);         // + q * 17: zpad(k0 * 256 + q * 16 + n0), n0 < 16
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = ah[q * 17];
        pk::dft16<false>(v);
#pragma unroll
        for (int q = 0; q < 16; ++q) ah[q * 17] = q ? pk::mul(v[q], w[q]) : v[0];
    }
}

// Stage C: last forward pass, product with the filter spectrum, first inverse pass.  With a second spectrum Gp2 the same
// forward result is filtered twice: the first product goes back to shared memory, the second to `stash` (global scratch in
// the thread's own order: element (h, q) of thread t at stash[(h * 16 + q) * 256 + t], coalesced) for conv8k_unstash;
// stash_add: add to what the stash holds (the inverse transform is linear: products of several inputs are summed there).
template <bool TWO>
__device__ __forceinline__ void conv8k_stage_c(float2 *a, const float2 *Gp, const float2 *Gp2, float2 *stash, bool stash_add = false)
{
    const int t = threadIdx.x;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float2 *ah = a + h * ZFP + t * 17;
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = ah[q];
        pk::dft16<false>(v);
        if (TWO) {
            // the forward result is parked in the thread's own shared-memory slots while the second product is formed:
            // two 16-element arrays plus the spectrum loads do not fit the register file (2 CTAs x 256 threads per SM)
#pragma unroll
            for (int q = 0; q < 16; ++q) ah[q] = v[q];
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
                const float4 g = __ldg(reinterpret_cast<const float4 *>(Gp2 + conv8k_gidx(h, t, q)));
                v[q] = pk::mul(v[q], make_float2(g.x, g.y));
                v[q + 1] = pk::mul(v[q + 1], make_float2(g.z, g.w));
            }
            pk::dft16<true>(v);
            if (stash_add) {                                   // several branches: their second products add up (linearity)
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = pk::add(v[q], __ldcg(stash + (h * 16 + q) * 256 + t));
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) stash[(h * 16 + q) * 256 + t] = v[q];
            asm volatile("" ::: "memory");                     // the reload below is a real shared-memory read
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = ah[q];
        }
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
            const float4 g = __ldg(reinterpret_cast<const float4 *>(Gp + conv8k_gidx(h, t, q)));
            v[q] = pk::mul(v[q], make_float2(g.x, g.y));
            v[q + 1] = pk::mul(v[q + 1], make_float2(g.z, g.w));
        }
        pk::dft16<true>(v);
#pragma unroll
        for (int q = 0; q < 16; ++q) ah[q] = v[q];
    }
}
// Halves of stage C on their own.  Stage B followed by conv8k_stage_c_fwd is a 256-point forward FFT of EVERY aligned
// 256-element block of the array (the last two passes of the 4096-point transforms work inside such blocks), each spectrum in
// the same (digit-reversed) order; conv8k_stage_c_inv followed by stage D is the inverse (unscaled: x 256), natural order
// out.  The block-FFT Park kernel (park.cu) uses the array as 32 independent 256-point transforms.
template <bool H0 = true, bool H1 = true>
__device__ __forceinline__ void conv8k_stage_c_fwd(float2 *a)
{
    const int t = threadIdx.x;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if ((h == 0 && !H0) || (h == 1 && !H1)) continue;
        float2 *ah = a + h * ZFP + t * 17;
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = ah[q];
        pk::dft16<false>(v);
#pragma unroll
        for (int q = 0; q < 16; ++q) ah[q] = v[q];
    }
}
template <bool H0 = true, bool H1 = true>
__device__ __forceinline__ void conv8k_stage_c_inv(float2 *a)
{
    const int t = threadIdx.x;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if ((h == 0 && !H0) || (h == 1 && !H1)) continue;
        float2 *ah = a + h * ZFP + t * 17;
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = ah[q];
        pk::dft16<true>(v);
#pragma unroll
        for (int q = 0; q < 16; ++q) ah[q] = v[q];
    }
}
// element e (0..255) of 256-block s (0..31) of the array
__device__ __forceinline__ int conv8k_blk(int s, int e) { return (s >> 4) * ZFP + zpad((s & 15) * 256 + e); }

// the stashed second product back into shared memory (same thread, same positions as stage C's own store), in two steps so
// that the L2 round trip can hide behind other work: fetch into registers, ... , put into shared memory
__device__ __forceinline__ void conv8k_unstash_fetch(const float2 *stash, float2 (&r)[32])
{
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __ldcg(stash + i * 256 + t);
}
__device__ __forceinline__ void conv8k_unstash_put(float2 *a, const float2 (&r)[32])
{
    const int t = threadIdx.x;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int q = 0; q < 16; ++q) a[h * ZFP + t * 17 + q] = r[h * 16 + q];
}
__device__ __forceinline__ void conv8k_unstash(float2 *a, const float2 *stash)
{
    float2 r[32];
    conv8k_unstash_fetch(stash, r);
    conv8k_unstash_put(a, r);
}

// Stage D: second inverse pass of both halves
template <bool H0 = true, bool H1 = true>
__device__ __forceinline__ void conv8k_stage_d(float2 *a, const pk::Seeds &sd)
{
    const int t = threadIdx.x, k0 = t >> 4, n0 = t & 15;
    float2 w[16];
    pk::powers(sd, w);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if ((h == 0 && !H0) || (h == 1 && !H1)) continue;
        float2 *ah = a + h * ZFP + zpad(k0 * 256 + n0);
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) { v[q] = ah[q * 17]; if (q) v[q] = pk::mulc(v[q], w[q]); }
        pk::dft16<true>(v);
#pragma unroll
        for (int q = 0; q < 16; ++q) ah[q * 17] = v[q];
    }
}

// seeds of stages A / E (base = thread index) and B / D (base = 16 (t & 15))
__device__ __forceinline__ pk::Seeds conv8k_seeds_ae(const double2 *tw) { return pk::seeds(tw, threadIdx.x); }
__device__ __forceinline__ pk::Seeds conv8k_seeds_bd(const double2 *tw) { return pk::seeds(tw, 16 * (threadIdx.x & 15)); }

// Stage E: last inverse pass + inverse radix-2 stage; sink(m, y, pre_m) receives local output sample m (0 <= m < 8192) of the
// circular convolution, already scaled (the 1/8192 sits in the spectrum).  Call order per thread: m = t + 256 q, then
// m + 4096, q ascending -- consecutive threads hold consecutive m.  pre(m) is called 8 outputs ahead of sink(m, ., .) and its
// result handed to that sink call: side data a sink needs from global memory (the FFT-form zc_freq kernel's samples) is in
// flight while the previous outputs are processed.  conv8k_no_pre for sinks that need none.
struct conv8k_no_pre { __device__ __forceinline__ int operator()(int) const { return 0; } };
template <class Pre, class Sink>
__device__ __forceinline__ void conv8k_stage_e(const float2 *a, const pk::Seeds &sd, float2 w0, Pre pre, Sink sink)
{
    const int t = threadIdx.x;
    auto m_of = [&](int slot) { return t + 256 * (slot >> 1) + (slot & 1) * ZF; };
    decltype(pre(0)) ring[8];
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) ring[sl] = pre(m_of(sl));
    float2 e[16], o[16];
    {
        float2 w[16];
        pk::powers(sd, w);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            e[q] = a[zpad(q * 256 + t)];
            o[q] = a[ZFP + zpad(q * 256 + t)];
            if (q) { e[q] = pk::mulc(e[q], w[q]); o[q] = pk::mulc(o[q], w[q]); }
        }
    }
    pk::dft16<true>(e);
    pk::dft16<true>(o);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        float2 v = pk::mulc(o[q], w0);
        if (q != 0) { const float2 k = conv8k_w32(q); v = pk::mulk(v, k.x, -k.y); }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int slot = 2 * q + hf;
            sink(m_of(slot), hf ? pk::sub(e[q], v) : pk::add(e[q], v), ring[slot & 7]);
            if (slot + 8 < 32) ring[slot & 7] = pre(m_of(slot + 8));
        }
    }
}

}  // namespace ofs

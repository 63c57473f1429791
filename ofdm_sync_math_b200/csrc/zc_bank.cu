// K6: multi-root Zadoff-Chu correlator bank on the 5th-generation tensor cores (tcgen05 / TMEM), one fused kernel.
//
// Generalises zc_freq.compute_frequency_metric (zc_freq.py:62-99, one root: np.vdot(template, bins) at :94) to a bank of
// up to 64 roots per pass evaluated at every candidate offset o of every capture:
//     bins[j, o] = DFT bin k_j of x[o+cp : o+cp+N]                        (62 used bins, zc_freq.py:88-93)
//     Y[r, o]    = sum_j conj(T[r, j]) bins[j, o]
//     metric     = |Y|^2 / (E_r E(o)),   E(o) = sum_j |bins[j, o]|^2        (zc_freq.py:95-97)
// and only the maximum over offsets (value + first offset) of every root is kept.
//
// One persistent CTA per SM, warp-specialised (17 warps):
//  * 8 PRODUCER warps = 8 chains x 64 bins (two bins per thread).  A chain is an eighth of the CTA's offset range walked sample by sample with
//    the sliding-DFT recurrence  b(o+1) = w (b(o) + x[o+cp+N] - x[o+cp]),  w = e^{+2 pi i k/N}  (two bins per thread; the
//    systematic error of the float rotation is divided out once per 32 steps, so rounding errors only random-walk).
//    Each chain owns 16 of the 128 rows of a tile: every step writes Re/Im of its bin straight into the MMA's A operand
//    in shared memory (K-major, 128-byte swizzle, the layout a tiled TMA copy would produce) -- the bins never exist in
//    HBM.  Row energies E(o) come from a transposed warp reduction (16 shuffles per 16 rows).
//  * 1 MMA warp: one elected thread issues 8 tcgen05.mma.kind::f16 (M=128 offsets, N=128 = 64 roots x {Re, Im}, K=16)
//    per tile into one of two TMEM accumulators; tcgen05.commit releases the operand buffer and publishes the accumulator.
//  * 8 EPILOGUE warps (two per TMEM lane quadrant = two chains, 32 roots each): tcgen05.ld, |Y|^2 / E(o), running maximum per root
//    in registers (tile number packed into the 10 low mantissa bits so the arg-max costs nothing per element).
// The templates (B operand, 32 KB) arrive once per CTA by tiled TMA.  A 4-stage operand ring and two accumulators hang together
// through mbarriers, so the SIMT producers, the tensor pipe and the epilogue overlap.
// Accuracy: FP16 operands (10-bit mantissa like TF32, half the shared-memory traffic -- the 128x128 tile is operand-fetch
// bound; every chain is rescaled by a power of two so the range fits), FP32 accumulation in TMEM, metric mantissa cut to 13 bits by the arg-max
// packing -> |d metric| <= 5e-3 * max(metric) (tested against the float64 oracle); arg-max offsets equal on clear peaks.
// Offsets whose in-band energy is below 1e-7 of the largest seen so far in the chain are skipped (after a burst followed
// by exact silence the recurrence holds a rounding residue, where the reference sees 0/eps = 0).
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>

namespace ofs {

constexpr int BK_ROWS = 128;      // M: offsets per tile (TMEM lanes) = 8 chains x 16
constexpr int BK_K = 128;         // K: 64 real + 64 imaginary bin columns
constexpr int BK_N = 128;         // N: 64 roots x {Re Y, Im Y}
constexpr int BK_MAXR = 64;       // roots per pass
constexpr int BK_PW = 8, BK_EW = 8;
constexpr int BK_CH = 8;          // chains per work item, 16 rows of every tile each
constexpr int BK_THREADS = (BK_PW + BK_EW + 1) * 32;  // 544 (17 warps): 96 registers per thread
constexpr int BK_TILE = BK_ROWS * BK_K * 2;          // 32 KB: fp16 operands
constexpr int BK_ST = 4;                             // operand stages (A ring)
constexpr int BK_TILES_MAX = 1024;                   // tile number must fit the 10 packed bits

// ------------------------------------------------------------------------------------------------ tcgen05 helpers
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 8 columns of this thread's TMEM lane, NOT waited for: the registers may only be read after tmem_wait8 on them
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// tcgen05.wait::ld with the loaded registers as in/out operands, so no consumer can be scheduled above the wait
__device__ __forceinline__ void tmem_wait8(uint32_t (&a)[8], uint32_t (&b)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(b[0]), "+r"(b[1]),
                   "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7])
                 :
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_b(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// bounded wait: a wrong descriptor must not hang the GPU -- trap after ~2^22 polls (seconds)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity)
{
    // try_wait with a suspend-time hint: the waiting warp sleeps in hardware instead of spinning through the issue slots the
    // producer warps need (polling loops were 20 % of the executed instructions).  Bounded: a wrong descriptor must trap, not hang.
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 20); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity), "r"(2000u)
            : "memory");
        if (ok) return;
    }
    __trap();
}

// ------------------------------------------------------------------------------------------------ prep / unpack
// B operand [128 rows][128 k] fp16, K-major, k = 2j (pairs with Re bins_j) and 2j+1 (pairs with Im bins_j):
//   row r < 64:  (Tre_r, Tim_r)   -> Re Y_r = sum_j Tre Bre + Tim Bim
//   row 64 + r: (-Tim_r, Tre_r)   -> Im Y_r = sum_j Tre Bim - Tim Bre
// scaled by beta = 1 / max|T| so any template set fits fp16 (E_r is scaled alike: the metric does not change),
// plus the rotation table: w_j = e^{+2 pi i k_j / N} rounded to float, and kappa_j (see below).
__global__ void zc_bank_prep_kernel(const float2 *templ, int nbins, int n_roots, const int *bins, int N, __half *Bmat, float *Er,
                                    float4 *wtab)
{
    __shared__ float smax[128];
    const int row = blockIdx.x, k = threadIdx.x;     // 128 x 128
    float mx = 0.f;
    for (int q = k; q < n_roots * nbins; q += 128) { const float2 t = templ[q]; mx = fmaxf(mx, fmaxf(fabsf(t.x), fabsf(t.y))); }
    smax[k] = mx;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (k < o) smax[k] = fmaxf(smax[k], smax[k + o]); __syncthreads(); }
    const float beta = smax[0] > 0.f ? 1.0f / smax[0] : 1.0f;
    const int r = row & 63, j = k >> 1, c = k & 1;
    float tre = 0.f, tim = 0.f;
    if (r < n_roots && j < nbins) { const float2 t = templ[r * nbins + j]; tre = t.x * beta; tim = t.y * beta; }
    const bool imag_row = row >= 64;
    Bmat[row * BK_K + k] = __float2half_rn(imag_row ? (c ? tre : -tim) : (c ? tim : tre));
    if (row < 64 && k == 0) {
        float e = 0.f;
        if (r < n_roots) for (int q = 0; q < nbins; ++q) { const float2 t = templ[r * nbins + q]; e += t.x * t.x + t.y * t.y; }
        Er[r] = e * beta * beta;
    }
    if (row == 0 && k == 0) Er[BK_MAXR] = beta * beta;
    if (row == 0 && k < 64) {
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < nbins) {
            double sn, cs;
            sincospi(2.0 * (double)bins[k] / (double)N, &sn, &cs);
            // rotation in float, plus the per-block correction kappa = (w / w_float)^32 that takes out its systematic
            // magnitude / phase error (2e-8 per step would otherwise leave a 6e-5 residue per sample leaving the window)
            const float hr = (float)cs, hi = (float)sn;
            const double mag2 = (double)hr * hr + (double)hi * hi;
            double qr = (cs * hr + sn * hi) / mag2, qi = (sn * hr - cs * hi) / mag2;      // w / w_float
            double kr = 1.0, ki = 0.0;
            for (int q = 0; q < 32; ++q) { const double t = kr * qr - ki * qi; ki = kr * qi + ki * qr; kr = t; }
            w = make_float4(hr, hi, (float)kr, (float)ki);
        }
        wtab[k] = w;
    }
}

__global__ void zc_bank_unpack_kernel(const unsigned long long *packed, int64_t count, int n_roots, const float *Er,
                                      float *best_metric, int32_t *best_offset)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned long long k = packed[i];
    const float er = Er[i % n_roots];
    const float v = __uint_as_float((unsigned)(k >> 32));
    best_metric[i] = er > 0.f ? v / er : 0.f;
    best_offset[i] = k ? (int32_t)(0xffffffffu - (unsigned)(k & 0xffffffffull)) : 0;
}

// ------------------------------------------------------------------------------------------------ fused bank kernel
struct BankParams {
    const float2 *x;
    int64_t n, n_off, seg_len, n_items;
    int N, cp, segs_per_cap, n_roots;
    long long *prof;   // OFS_BANK_DBG & 8: per-CTA cycle counters [grid][16]
    int dbg;      // timing experiments only (build with -DOFS_BANK_EXPERIMENTS, then OFS_BANK_DBG): 1 skip MMAs, 2 skip epilogue
                  // math, 4 skip producer math, 8 cycle counters; 0 in the shipped library
    const float4 *wtab;
    unsigned long long *best_packed;
    float *metric_out;          // optional: the full metric row of root 0 (zc_freq fast path), [frames][metric_stride]
    int64_t metric_stride;
    const float *Er;            // [64] E_r * beta^2, [64] = beta^2
    float templ_energy;         // metric_out = |Y_0|^2 / (templ_energy * E(o))
};

__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// (re, im) -> packed fp16 pair (saturating: a burst the scale estimate missed must not become inf), one 32-bit store
__device__ __forceinline__ void sts_h2(uint32_t addr, float re, float im)
{
    uint32_t v;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(v) : "f"(im), "f"(re));      // low half = re
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v));
}

// timing-experiment switches are a compile-time zero in the shipped library: the branches disappear from the step loops
#ifdef OFS_BANK_EXPERIMENTS
#define BK_DBG (p.dbg)
#else
#define BK_DBG 0
#endif

__global__ void __launch_bounds__(BK_THREADS, 1)
zc_bank_fused_kernel(const __grid_constant__ CUtensorMap mapT, const BankParams p)
{
    extern __shared__ __align__(1024) unsigned char bsm[];
    unsigned char *sT = bsm;                                  // 32 KB templates (B operand)
    unsigned char *sA = bsm + BK_TILE;                        // BK_ST x 32 KB bins tiles (A operand ring)
    unsigned char *aux = bsm + (1 + BK_ST) * BK_TILE;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(aux);     // [BK_ST]
    uint64_t *a_empty = a_full + BK_ST, *d_full = a_full + 2 * BK_ST, *d_empty = d_full + 2, *t_full = d_full + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(aux + 120);
    unsigned *sBest = reinterpret_cast<unsigned *>(aux + 128);            // [64]
    unsigned *sOff = sBest + 64;                                          // [64]
    float *sE = reinterpret_cast<float *>(aux + 1024);                    // [8 slots][128 rows]
    float4 *sC = reinterpret_cast<float4 *>(aux + 1024 + 4096);           // [8 warps][2][32] comb samples as (x, x, y, y)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == BK_PW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int b = 0; b < BK_ST; ++b) { mbar_init(&a_full[b], BK_PW); mbar_init(&a_empty[b], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&d_full[b], 1); mbar_init(&d_empty[b], BK_EW); }
        mbar_init(t_full, 1);
        mbar_fence_init();
    }
    if (tid < 64) { sBest[tid] = 0u; sOff[tid] = 0xffffffffu; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // geometry of one work item (capture, segment): 4 chains of Q offsets, Q a multiple of 32
    auto item_geom = [&](int64_t item, int64_t &cap, int64_t &seg_lo, int64_t &seg_hi, int64_t &Q) {
        cap = item / p.segs_per_cap;
        seg_lo = (item % p.segs_per_cap) * p.seg_len;
        seg_hi = seg_lo + p.seg_len < p.n_off ? seg_lo + p.seg_len : p.n_off;
        Q = (((seg_hi - seg_lo + BK_CH - 1) / BK_CH) + 31) & ~(int64_t)31;
    };

    if (warp < BK_PW) {
        // ============================================================ producers: sliding DFT -> A operand
        // warp g = chain g = rows 16g .. 16g+15 of every tile; lane l carries bins l and l + 32 (two independent recurrences
        // per thread, two producer warps per scheduler: the dependent FADD -> FMUL -> FFMA chains hide behind each other)
        const int g = warp;
        // both bins ride in the halves of packed fp32 registers (FADD2 / FMUL2 / FFMA2): the kernel is issue-bound
        const float4 w0 = p.wtab[lane], w1 = p.wtab[32 + lane];
        const float2 WX = make_float2(w0.x, w1.x), WY = make_float2(w0.y, w1.y), NWY = make_float2(-w0.y, -w1.y);
        const float2 KX = make_float2(w0.z, w1.z), KY = make_float2(w0.w, w1.w), NKY = make_float2(-w0.w, -w1.w);
        float4 *sc = sC + warp * 64;
        // shared-memory byte offsets of this thread's first (Re, Im) fp16 pair for rows with (row & 7) == m, first row of the
        // chain; the second bin sits one K-chunk (16 KB) further
        uint32_t base8[8];
#pragma unroll
        for (int m = 0; m < 8; ++m)
            base8[m] = (uint32_t)((16 * g) * 128 + (((lane >> 2) ^ m) << 4) + (lane & 3) * 4);
        const uint32_t sA_u = smem_u32(sA);
        uint32_t it = 0;
        long long c_wait = 0, c_math = 0, c_red = 0, c_tail = 0, c_warm = 0, tk = 0;
        const bool prof = (BK_DBG & 8) && g == 0;
#define BK_TICK(acc) do { if (prof) { const long long now = clock64(); acc += now - tk; tk = now; } } while (0)
        for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int64_t cap, seg_lo, seg_hi, Q;
            item_geom(item, cap, seg_lo, seg_hi, Q);
            const float2 *xc = p.x + cap * p.n;
            const int64_t s0 = seg_lo + (int64_t)g * Q + p.cp;           // sample index of the chain's first window
            const int n_warm = p.N / 32, n_pairs = (int)(Q / 32);       // blocks of 32 samples: warm-up, then 2 tiles each
            auto ldx = [&](int64_t idx) { return idx < p.n ? __ldg(xc + idx) : make_float2(0.f, 0.f); };
            // fp16 operands: the chain works on alpha * x with alpha = 2^k chosen so that even a fully coherent window
            // (|bins| <= N max|x|) stays below 32768; every row may carry its own scale because the metric divides by
            // the row's own energy.  max|x| is estimated from 64 blocks of 32 samples spread over the chain's range.
            float alpha;
            {
                float m2 = 0.f;
                const int64_t span = Q + p.N;
                for (int q = 0; q < 64; ++q) {
                    const float2 v = ldx(s0 + (span * q >> 6) + lane);
                    m2 = fmaxf(m2, fmaf(v.x, v.x, v.y * v.y));
                }
                m2 = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m2)));
                alpha = 1.f;
                if (m2 > 0.f) {
                    int ex;
                    (void)frexpf(32768.0f / ((float)p.N * sqrtf(m2)), &ex);      // value = f * 2^ex, f in [0.5, 1)
                    ex = ex - 1 < -100 ? -100 : (ex - 1 > 100 ? 100 : ex - 1);
                    alpha = ldexpf(1.0f, ex);
                }
            }
            // feed u: warm-up blocks bring x[s0 + 32u + lane] into an empty window; real blocks the comb x[s+N] - x[s]
            auto feed = [&](int u, float2 &a, float2 &b) {      // comb sample = a - b (subtracted when it is stored)
                if (u < n_warm) { a = ldx(s0 + 32 * (int64_t)u + lane); b = make_float2(0.f, 0.f); return; }
                const int64_t s = s0 + 32 * (int64_t)(u - n_warm) + lane;
                a = ldx(s + p.N); b = ldx(s);
            };
            float2 Bx = make_float2(0.f, 0.f), By = Bx;          // (bin lane, bin lane + 32)
            {
                float2 a, b;
                feed(0, a, b);
                const float vx = alpha * (a.x - b.x), vy = alpha * (a.y - b.y);
                sc[lane] = make_float4(vx, vx, vy, vy);
            }
            __syncwarp();
            const int n_blocks = n_warm + n_pairs;
            for (int u = 0; u < n_blocks; ++u) {
                float2 na = make_float2(0.f, 0.f), nb = na;
                if (u + 1 < n_blocks) feed(u + 1, na, nb);             // in flight during the 32 steps below
                const float4 *cb4 = sc + (u & 1) * 32;
                if (prof) tk = clock64();
                if (u < n_warm) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float4 c = cb4[i];
                        const float2 T = __fadd2_rn(Bx, make_float2(c.x, c.y)), U = __fadd2_rn(By, make_float2(c.z, c.w));
                        Bx = __ffma2_rn(T, WX, __fmul2_rn(U, NWY));
                        By = __ffma2_rn(T, WY, __fmul2_rn(U, WX));
                    }
                    BK_TICK(c_warm);
                } else {
#pragma unroll
                    for (int tl = 0; tl < 2; ++tl) {                  // two tiles (16 rows each) per 32-sample block
                        const uint32_t buf = it % BK_ST;
                        if (it >= BK_ST) mbar_wait_bounded(&a_empty[buf], ((it / BK_ST) & 1u) ^ 1u);
                        const uint32_t tb = sA_u + buf * BK_TILE;
                        BK_TICK(c_wait);
                        float e[16];
                        if (BK_DBG & 4) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) e[i] = 0.f;
                        } else {
                            // comb samples: 4 steps' worth (4 x LDS.128, warp broadcast) fetched before the stores of their
                            // group -- ptxas will not move a shared load above a shared store on its own
#pragma unroll
                            for (int grp = 0; grp < 4; ++grp) {
                                float4 cg[4];
#pragma unroll
                                for (int q = 0; q < 4; ++q) cg[q] = cb4[16 * tl + 4 * grp + q];
#pragma unroll
                                for (int ii = 0; ii < 4; ++ii) {
                                    const int i = 4 * grp + ii;
                                    const float2 e2 = __ffma2_rn(Bx, Bx, __fmul2_rn(By, By));
                                    e[i] = e2.x + e2.y;
                                    const uint32_t a = tb + base8[i & 7] + i * 128;
                                    sts_h2(a, Bx.x, By.x);
                                    sts_h2(a + 16384, Bx.y, By.y);
                                    const float4 c = cg[ii];
                                    const float2 T = __fadd2_rn(Bx, make_float2(c.x, c.y)), U = __fadd2_rn(By, make_float2(c.z, c.w));
                                    Bx = __ffma2_rn(T, WX, __fmul2_rn(U, NWY));
                                    By = __ffma2_rn(T, WY, __fmul2_rn(U, WX));
                                }
                            }
                        }
                        BK_TICK(c_math);
                        // transposed reduction of the 16 row energies over the 32 lanes (= 64 bins): lanes 2r, 2r+1 end with row r
#pragma unroll
                        for (int s = 16; s >= 2; s >>= 1) {
                            const bool hi = (lane & s) != 0;
#pragma unroll
                            for (int k = 0; k < s / 2; ++k) {
                                const float send = hi ? e[k] : e[k + s / 2];
                                const float keep = hi ? e[k + s / 2] : e[k];
                                e[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                            }
                        }
                        e[0] += __shfl_xor_sync(0xffffffffu, e[0], 1);
                        if (!(lane & 1)) sE[(it & 7u) * 128 + 16 * g + (lane >> 1)] = e[0];
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cta(&a_full[buf]);
                        ++it;
                        BK_TICK(c_red);
                    }
                }
                asm volatile("" : "+f"(na.x), "+f"(na.y), "+f"(nb.x), "+f"(nb.y));   // keep the subtraction (and the wait) down here
                {
                    const float vx = alpha * (na.x - nb.x), vy = alpha * (na.y - nb.y);
                    sc[((u + 1) & 1) * 32 + lane] = make_float4(vx, vx, vy, vy);
                }
                {   // per-block correction of the float rotation: b *= kappa
                    const float2 nx = __ffma2_rn(Bx, KX, __fmul2_rn(By, NKY)), ny = __ffma2_rn(Bx, KY, __fmul2_rn(By, KX));
                    Bx = nx; By = ny;
                }
                __syncwarp();
                BK_TICK(c_tail);
            }
        }
        if (prof && lane == 0) {
            long long *o = p.prof + (size_t)blockIdx.x * 16;
            o[0] = c_wait; o[1] = c_math; o[2] = c_red; o[3] = c_tail; o[4] = c_warm; o[5] = it;
        }
    } else if (warp == BK_PW + BK_EW) {
        // ============================================================ MMA issuer (one elected thread)
        if (lane == 0) {
            uint32_t mt = 0, tot = 0;
            // instruction descriptor: D = F32 (bit 4), A = B = F16 (format 0), both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BK_N >> 3) << 17) | ((uint32_t)(BK_ROWS >> 4) << 24);
            for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int64_t cap, seg_lo, seg_hi, Q;
                item_geom(item, cap, seg_lo, seg_hi, Q);
                tot += (uint32_t)(Q / 16);
            }
            mbar_expect_tx(t_full, BK_TILE);
            for (int kc = 0; kc < 2; ++kc) tma_load_2d_b(sT + kc * 16384, &mapT, kc * 64, 0, t_full);
            mbar_wait_bounded(t_full, 0);
            long long m_wa = 0, m_wd = 0, m_is = 0, tk = clock64();
            const bool prof = (BK_DBG & 8) != 0;
            for (; mt < tot; ++mt) {
                const uint32_t buf = mt % BK_ST, acc = mt & 1u;
                mbar_wait_bounded(&a_full[buf], (mt / BK_ST) & 1u);
                BK_TICK(m_wa);
                if (mt >= 2) mbar_wait_bounded(&d_empty[acc], ((mt >> 1) & 1u) ^ 1u);
                BK_TICK(m_wd);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned char *At = sA + buf * BK_TILE;
                if (!(BK_DBG & 1))
#pragma unroll
                for (int step = 0; step < 8; ++step) {
                    // K loop: 8 steps of 16 fp16; K-chunk = step/4 (a 128-row x 128-byte box), 32 bytes per step inside the
                    // swizzled 128-byte row; 8-row atoms are 1024 B apart (SBO)
                    const uint64_t ad = umma_desc(smem_u32(At + (step >> 2) * 16384 + (step & 3) * 32), 16, 1024);
                    const uint64_t bd = umma_desc(smem_u32(sT + (step >> 2) * 16384 + (step & 3) * 32), 16, 1024);
                    umma_f16(tmem + acc * BK_N, ad, bd, idesc, step > 0 ? 1u : 0u);
                }
                umma_commit(&a_empty[buf]);
                umma_commit(&d_full[acc]);
                BK_TICK(m_is);
            }
            if (prof) { long long *o = p.prof + (size_t)blockIdx.x * 16; o[6] = m_wa; o[7] = m_wd; o[8] = m_is; }
        }
    } else {
        // ============================================================ epilogue: two warps per TMEM lane quadrant (= two chains),
        // 32 roots each; eight warps keep enough TMEM reads in flight to hide their latency
        const int q4 = warp & 3;                              // lane quadrant this warp may read (warp % 4)
        const int half = (warp - BK_PW) >> 2;                 // roots 32*half .. 32*half + 31
        const uint32_t tq = tmem + ((uint32_t)(q4 * 32) << 16) + 32 * half;
        const int et = tid - BK_PW * 32;                      // 0..255
        const int chain = 2 * q4 + (lane >> 4);
        const float mscale = p.metric_out ? 1.0f / (p.Er[BK_MAXR] * p.templ_energy) : 0.f;
        uint32_t it = 0;
        long long e_wait = 0, e_work = 0, tk = clock64();
        const bool prof = (BK_DBG & 8) && warp == BK_PW;
        for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int64_t cap, seg_lo, seg_hi, Q;
            item_geom(item, cap, seg_lo, seg_hi, Q);
            const int n_tiles = (int)(Q / 16);
            const int64_t q0 = seg_lo + (int64_t)chain * Q + (lane & 15);       // this thread's offset in tile 0
            const int64_t q_end = seg_lo + (int64_t)(chain + 1) * Q < seg_hi ? seg_lo + (int64_t)(chain + 1) * Q : seg_hi;
            unsigned best[32];
#pragma unroll
            for (int r = 0; r < 32; ++r) best[r] = 0u;
            float emax = 0.f;
            for (int t = 0; t < n_tiles; ++t, ++it) {
                const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
                mbar_wait_bounded(&d_full[buf], ph);
                BK_TICK(e_wait);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const float e = sE[(it & 7u) * 128 + 32 * q4 + lane];
                // running maximum per chain (= half warp: the two chains of a quadrant carry different scales)
                emax = fmaxf(emax, __uint_as_float(__reduce_max_sync(lane < 16 ? 0x0000ffffu : 0xffff0000u, __float_as_uint(e))));
                const int64_t o = q0 + 16 * (int64_t)t;
                const float inv = (o < q_end && e > 1e-7f * emax && e > 0.f) ? 1.0f / e : 0.f;
                const unsigned tb = (unsigned)(BK_TILES_MAX - 1 - t);
                // 8 roots at a time (Re columns 32 half + 8c.., Im columns 64 more); the loads of chunk c+1 are in flight while chunk c is
                // reduced, so the TMEM read latency is paid once per tile, not once per chunk
                uint32_t re[2][8], im[2][8];
                if (BK_DBG & 2) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) re[0][q] = im[0][q] = re[1][q] = im[1][q] = 0u;
                } else {
                    tmem_ld8_async(tq + buf * BK_N, re[0]);
                    tmem_ld8_async(tq + buf * BK_N + 64, im[0]);
                    tmem_wait8(re[0], im[0]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < 3 && !(BK_DBG & 2)) {
                        tmem_ld8_async(tq + buf * BK_N + 8 * (c + 1), re[(c + 1) & 1]);
                        tmem_ld8_async(tq + buf * BK_N + 64 + 8 * (c + 1), im[(c + 1) & 1]);
                    }
#pragma unroll
                    for (int q = 0; q < 8; q += 2) {
                        const float2 yr = make_float2(__uint_as_float(re[c & 1][q]), __uint_as_float(re[c & 1][q + 1]));
                        const float2 yi = make_float2(__uint_as_float(im[c & 1][q]), __uint_as_float(im[c & 1][q + 1]));
                        const float2 m = __fmul2_rn(__ffma2_rn(yr, yr, __fmul2_rn(yi, yi)), make_float2(inv, inv));
                        if (c == 0 && q == 0 && half == 0 && p.metric_out && o < q_end) p.metric_out[cap * p.metric_stride + o] = m.x * mscale;
                        best[8 * c + q] = max(best[8 * c + q], (__float_as_uint(m.x) & 0xfffffc00u) | tb);
                        best[8 * c + q + 1] = max(best[8 * c + q + 1], (__float_as_uint(m.y) & 0xfffffc00u) | tb);
                    }
                    if (c < 3 && !(BK_DBG & 2)) tmem_wait8(re[(c + 1) & 1], im[(c + 1) & 1]);
                    if (c == 2) {                             // the last loads have landed: the accumulator may be overwritten
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cta(&d_empty[buf]);
                    }
                }
                BK_TICK(e_work);
            }
            // ---- reduce over the 256 epilogue threads: max key per root, then the earliest offset holding it
#pragma unroll
            for (int r = 0; r < 32; ++r)
                if (32 * half + r < p.n_roots && (best[r] & 0xfffffc00u)) atomicMax(&sBest[32 * half + r], best[r]);
            asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                if (32 * half + r < p.n_roots && best[r] == sBest[32 * half + r] && (best[r] & 0xfffffc00u)) {
                    const int t = BK_TILES_MAX - 1 - (int)(best[r] & 0x3ffu);
                    atomicMin(&sOff[32 * half + r], (unsigned)(q0 + 16 * (int64_t)t));
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et < p.n_roots && sBest[et]) {
                const unsigned long long key = ((unsigned long long)(sBest[et] & 0xfffffc00u) << 32) | (unsigned long long)(0xffffffffu - sOff[et]);
                atomicMax(p.best_packed + cap * p.n_roots + et, key);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et < 64) { sBest[et] = 0u; sOff[et] = 0xffffffffu; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        if (prof && lane == 0) { long long *o = p.prof + (size_t)blockIdx.x * 16; o[9] = e_wait; o[10] = e_work; }
    }
#undef BK_TICK
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == BK_PW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

// ------------------------------------------------------------------------------------------------ K5b: float32 sliding DFT
// zc_freq.compute_frequency_metric (zc_freq.py:62-99) for complex64 single-branch captures at float32 accuracy, SIMT only:
// the producer recurrence of the bank above -- b(o+1) = w (b(o) + x[o+cp+N] - x[o+cp]), two bins per lane in packed fp32,
// the rotation's systematic error divided out every 32 steps so that rounding only random-walks (~6e-8 sqrt(steps)) --
// with a float32 epilogue instead of the fp16 tensor-core product: every step adds |b|^2 and conj(T) b of the lane's two
// bins to 16-step partials, a transposed warp reduction (3 x 15 shuffles per 16 offsets) sums them over the 64 bins, and the
// even lanes finish metric = |corr|^2 / max(E_T E, 1e-12).  One warp = one chain of offsets (N warm-up steps from an empty
// window, then chain_len offsets).  Bound: fp32 FMA / issue -- 62 bins x ~12 FMA-class instructions per offset (the
// reference's per-offset FFT is replaced by the recurrence, but the 62 bins are each 2 complex MACs per offset whatever one
// does): 134 M offsets of cfg 4 are ~3 G warp instructions, 2.7 ms at full issue rate.
constexpr int SD_WARPS = 8;
struct SdftParams {
    const float2 *x;
    int64_t n, n_off, chain_len, n_chains, mstride;
    int N, cp, chains_per_cap;
    const float4 *wtab;        // [64] (w_re, w_im, kappa_re, kappa_im)
    const float2 *ttab;        // [64] template bins (0 beyond nbins)
    float templ_energy;
    float *metric;
};

__global__ void zc_sdft_prep_kernel(const float2 *templ, int nbins, const int *bins, int N, float4 *wtab, float2 *ttab)
{
    const int k = threadIdx.x;          // 64 threads
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 t = make_float2(0.f, 0.f);
    if (k < nbins) {
        double sn, cs;
        sincospi(2.0 * (double)bins[k] / (double)N, &sn, &cs);
        const float hr = (float)cs, hi = (float)sn;
        const double mag2 = (double)hr * hr + (double)hi * hi;
        double qr = (cs * hr + sn * hi) / mag2, qi = (sn * hr - cs * hi) / mag2;      // w / w_float
        double kr = 1.0, ki = 0.0;
        for (int q = 0; q < 32; ++q) { const double u = kr * qr - ki * qi; ki = kr * qi + ki * qr; kr = u; }
        w = make_float4(hr, hi, (float)kr, (float)ki);
        t = templ[k];
    }
    wtab[k] = w; ttab[k] = t;
}

__global__ void __launch_bounds__(SD_WARPS * 32) zc_sdft_kernel(const SdftParams p)
{
    __shared__ float4 sC[SD_WARPS][2][32];        // comb samples of a 32-step block as (x, x, y, y), double-buffered per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t chain = (int64_t)blockIdx.x * SD_WARPS + warp;
    if (chain >= p.n_chains) return;
    const int64_t cap = chain / p.chains_per_cap;
    const int64_t o_lo = (chain % p.chains_per_cap) * p.chain_len;
    const int64_t o_hi = o_lo + p.chain_len < p.n_off ? o_lo + p.chain_len : p.n_off;
    if (o_lo >= o_hi) return;
    const float2 *xc = p.x + cap * p.n;
    float *mrow = p.metric + cap * p.mstride;
    const int64_t s0 = o_lo + p.cp;                                   // first sample of the chain's first window
    const float4 w0 = p.wtab[lane], w1 = p.wtab[32 + lane];
    const float2 WX = make_float2(w0.x, w1.x), WY = make_float2(w0.y, w1.y), NWY = make_float2(-w0.y, -w1.y);
    const float2 KX = make_float2(w0.z, w1.z), KY = make_float2(w0.w, w1.w), NKY = make_float2(-w0.w, -w1.w);
    const float2 t0 = p.ttab[lane], t1 = p.ttab[32 + lane];
    const float2 TR = make_float2(t0.x, t1.x), TI = make_float2(t0.y, t1.y), NTI = make_float2(-t0.y, -t1.y);
    float4 *sc = &sC[warp][0][0];
    auto ldx = [&](int64_t idx) { return idx < p.n ? __ldg(xc + idx) : make_float2(0.f, 0.f); };
    const int n_warm = p.N / 32, n_real = (int)((o_hi - o_lo + 31) / 32), n_blocks = n_warm + n_real;
    auto feed = [&](int u, float2 &a, float2 &b) {      // comb sample = a - b
        if (u < n_warm) { a = ldx(s0 + 32 * (int64_t)u + lane); b = make_float2(0.f, 0.f); return; }
        const int64_t s = s0 + 32 * (int64_t)(u - n_warm) + lane;
        a = ldx(s + p.N); b = ldx(s);
    };
    float2 Bx = make_float2(0.f, 0.f), By = Bx;          // (bin lane, bin lane + 32): real parts, imaginary parts
    {
        float2 a, b;
        feed(0, a, b);
        sc[lane] = make_float4(a.x - b.x, a.x - b.x, a.y - b.y, a.y - b.y);
    }
    __syncwarp();
    for (int u = 0; u < n_blocks; ++u) {
        float2 na = make_float2(0.f, 0.f), nb = na;
        if (u + 1 < n_blocks) feed(u + 1, na, nb);             // in flight during the 32 steps below
        const float4 *cb4 = sc + (u & 1) * 32;
        if (u < n_warm) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float4 c = cb4[i];
                const float2 T = __fadd2_rn(Bx, make_float2(c.x, c.y)), U = __fadd2_rn(By, make_float2(c.z, c.w));
                Bx = __ffma2_rn(T, WX, __fmul2_rn(U, NWY));
                By = __ffma2_rn(T, WY, __fmul2_rn(U, WX));
            }
        } else {
            const int64_t ob = o_lo + 32 * (int64_t)(u - n_warm);   // offset of step 0 of this block
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float e[16], cr[16], ci[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 e2 = __ffma2_rn(Bx, Bx, __fmul2_rn(By, By));
                    e[i] = e2.x + e2.y;
                    // conj(T) b: re = Tr br + Ti bi, im = Tr bi - Ti br  (np.vdot, zc_freq.py:94)
                    const float2 r2 = __ffma2_rn(TR, Bx, __fmul2_rn(TI, By)), i2 = __ffma2_rn(TR, By, __fmul2_rn(NTI, Bx));
                    cr[i] = r2.x + r2.y; ci[i] = i2.x + i2.y;
                    const float4 c = cb4[16 * hf + i];
                    const float2 T = __fadd2_rn(Bx, make_float2(c.x, c.y)), U = __fadd2_rn(By, make_float2(c.z, c.w));
                    Bx = __ffma2_rn(T, WX, __fmul2_rn(U, NWY));
                    By = __ffma2_rn(T, WY, __fmul2_rn(U, WX));
                }
                // transposed reduction over the 32 lanes (= 64 bins): lanes 2r, 2r + 1 end with the totals of step r
#pragma unroll
                for (int s = 16; s >= 2; s >>= 1) {
                    const bool hi = (lane & s) != 0;
#pragma unroll
                    for (int k = 0; k < s / 2; ++k) {
                        const float se = hi ? e[k] : e[k + s / 2], ke = hi ? e[k + s / 2] : e[k];
                        const float sr = hi ? cr[k] : cr[k + s / 2], kr = hi ? cr[k + s / 2] : cr[k];
                        const float si = hi ? ci[k] : ci[k + s / 2], ki = hi ? ci[k + s / 2] : ci[k];
                        e[k] = ke + __shfl_xor_sync(0xffffffffu, se, s);
                        cr[k] = kr + __shfl_xor_sync(0xffffffffu, sr, s);
                        ci[k] = ki + __shfl_xor_sync(0xffffffffu, si, s);
                    }
                }
                e[0] += __shfl_xor_sync(0xffffffffu, e[0], 1);
                cr[0] += __shfl_xor_sync(0xffffffffu, cr[0], 1);
                ci[0] += __shfl_xor_sync(0xffffffffu, ci[0], 1);
                const int64_t o = ob + 16 * hf + (lane >> 1);
                if (!(lane & 1) && o < o_hi) {
                    const float den = fmaxf(p.templ_energy * e[0], 1e-12f);       // zc_freq.py:96-97
                    mrow[o] = fmaf(cr[0], cr[0], ci[0] * ci[0]) / den;
                }
            }
        }
        asm volatile("" : "+f"(na.x), "+f"(na.y), "+f"(nb.x), "+f"(nb.y));
        sc[((u + 1) & 1) * 32 + lane] = make_float4(na.x - nb.x, na.x - nb.x, na.y - nb.y, na.y - nb.y);
        {   // per-block correction of the float rotation: b *= kappa
            const float2 nx = __ffma2_rn(Bx, KX, __fmul2_rn(By, NKY)), ny = __ffma2_rn(Bx, KY, __fmul2_rn(By, KX));
            Bx = nx; By = ny;
        }
        __syncwarp();
    }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 encode_fn2()
{
    static EncodeTiledFn2 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn2)ptr;
        else
            (void)cudaGetLastError();
    }
    return fn;
}
static bool make_map_f16(CUtensorMap *map, const void *base, uint64_t inner, uint64_t rows, uint32_t box_rows)
{
    EncodeTiledFn2 fn = encode_fn2();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {inner, rows};
    const cuuint64_t gstride[1] = {inner * 2};
    const cuuint32_t box[2] = {64, box_rows};          // 64 halfs = one 128-byte swizzle row
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ofs

using namespace ofs;

static int bank_run(const void *x_c64, int64_t n_frames, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                    const void *templ_c64, int32_t nbins, int32_t n_roots, float *best_metric, int32_t *best_offset,
                    float *metric_out, int64_t metric_stride, float templ_energy, void *stream_)
{
    OFS_REQUIRE(x_c64 && bins && templ_c64 && ((best_metric && best_offset) || metric_out), "ofs_zc_bank: null argument");
    OFS_REQUIRE(n_fft >= 32 && n_fft <= 65536 && n_fft % 32 == 0 && cp >= 0, "ofs_zc_bank: n_fft must be a multiple of 32 in 32..65536");
    OFS_REQUIRE(nbins >= 1 && nbins <= 64 && n_roots >= 1 && n_roots <= 128, "ofs_zc_bank: nbins <= 64, n_roots <= 128");
    const int64_t n_off = n - ((int64_t)n_fft + cp) + 1;
    OFS_REQUIRE(n_off > 0, "Received stream is shorter than a single OFDM symbol.");
    OFS_REQUIRE(n_off < (1LL << 31), "ofs_zc_bank: captures longer than 2^31 offsets unsupported");
    if (n_frames == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    keep_pool_cached();
    __half *Bmat = nullptr;
    float *Er = nullptr;
    float4 *wtab = nullptr;
    unsigned long long *packed = nullptr;
    OFS_CUDA(cudaMallocAsync((void **)&Bmat, (size_t)BK_N * BK_K * 2, stream));
    OFS_CUDA(cudaMallocAsync((void **)&Er, (BK_MAXR + 1) * 4, stream));
    OFS_CUDA(cudaMallocAsync((void **)&wtab, 64 * sizeof(float4), stream));
    OFS_CUDA(cudaMallocAsync((void **)&packed, (size_t)n_frames * BK_MAXR * 8, stream));

    // segments: <= 1024 tiles of 128 offsets per work item; more items when the batch is small (each chain pays an
    // n_fft-step warm-up, so segments stay >= 8 windows long)
    const int64_t seg_max = (int64_t)BK_TILES_MAX * BK_ROWS;
    int64_t segs = (n_off + seg_max - 1) / seg_max;
    const int64_t want = (2LL * sm_count() + n_frames - 1) / n_frames;
    const int64_t most = n_off / (8LL * n_fft) > 1 ? n_off / (8LL * n_fft) : 1;
    if (want > segs) segs = want < most ? want : (most > segs ? most : segs);
    int64_t seg_len = ((n_off + segs - 1) / segs + BK_ROWS - 1) / BK_ROWS * BK_ROWS;
    if (seg_len > seg_max) seg_len = seg_max;
    segs = (n_off + seg_len - 1) / seg_len;

    const size_t smem = (1 + BK_ST) * BK_TILE + 1024 + 4096 + 8192 + 1024;
    static PerDeviceOnce once;
    if (!once.done()) {
        OFS_CUDA(cudaFuncSetAttribute(zc_bank_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        once.mark();
    }
    for (int r0 = 0; r0 < n_roots; r0 += BK_MAXR) {
        const int nr = n_roots - r0 < BK_MAXR ? n_roots - r0 : BK_MAXR;
        zc_bank_prep_kernel<<<BK_N, BK_K, 0, stream>>>(reinterpret_cast<const float2 *>(templ_c64) + (size_t)r0 * nbins, nbins, nr, bins,
                                                       n_fft, Bmat, Er, wtab);
        if (int rc = check_launch("zc_bank_prep_kernel")) return rc;
        CUtensorMap mapT;
        OFS_REQUIRE(make_map_f16(&mapT, Bmat, BK_K, BK_N, 128), "ofs_zc_bank: cuTensorMapEncodeTiled failed");
        OFS_CUDA(cudaMemsetAsync(packed, 0, (size_t)n_frames * nr * 8, stream));
        BankParams p{};
        p.x = reinterpret_cast<const float2 *>(x_c64); p.n = n; p.n_off = n_off; p.seg_len = seg_len; p.segs_per_cap = (int)segs;
#ifdef OFS_BANK_EXPERIMENTS      // timing experiments (skip stages, cycle counters): compiled out of the shipped library
        { const char *dbg = getenv("OFS_BANK_DBG"); p.dbg = dbg ? atoi(dbg) : 0; }
#else
        p.dbg = 0;
#endif
        const int64_t grid_dbg = n_frames * segs < sm_count() ? n_frames * segs : sm_count();
        if (p.dbg & 8) { OFS_CUDA(cudaMalloc((void **)&p.prof, (size_t)grid_dbg * 16 * 8)); OFS_CUDA(cudaMemset(p.prof, 0, (size_t)grid_dbg * 16 * 8)); }
        p.n_items = n_frames * segs; p.N = n_fft; p.cp = cp; p.n_roots = nr; p.wtab = wtab; p.best_packed = packed;
        p.metric_out = metric_out; p.metric_stride = metric_stride; p.Er = Er; p.templ_energy = templ_energy;
        const int64_t grid = p.n_items < sm_count() ? p.n_items : sm_count();
        zc_bank_fused_kernel<<<(unsigned)grid, BK_THREADS, smem, stream>>>(mapT, p);
        if (int rc = check_launch("zc_bank_fused_kernel")) return rc;
        if (p.dbg & 8) {       // timing experiment: average cycles per tile spent in each phase (stderr)
            OFS_CUDA(cudaStreamSynchronize(stream));
            long long *h = (long long *)malloc((size_t)grid * 16 * 8);
            cudaMemcpy(h, p.prof, (size_t)grid * 16 * 8, cudaMemcpyDeviceToHost);
            double a[16] = {0}; double tiles = 0;
            for (int64_t b = 0; b < grid; ++b) { for (int q = 0; q < 16; ++q) a[q] += (double)h[b * 16 + q]; tiles += (double)h[b * 16 + 5]; }
            fprintf(stderr, "bank cycles/tile: producer wait %.0f math %.0f reduce+arrive %.0f tail/2 %.0f warm %.0f | mma wait_a %.0f wait_d %.0f issue %.0f | "
                            "epilogue wait %.0f work %.0f (tiles/CTA %.0f)\n", a[0] / tiles, a[1] / tiles, a[2] / tiles, a[3] / tiles, a[4] / tiles,
                    a[6] / tiles, a[7] / tiles, a[8] / tiles, a[9] / tiles, a[10] / tiles, tiles / (double)grid);
            free(h); cudaFree(p.prof);
        }
        if (!best_metric) {
            // metric-row mode (zc_freq fast path): nothing to unpack
        } else if (n_roots <= BK_MAXR) {
            zc_bank_unpack_kernel<<<(unsigned)((n_frames * nr + 255) / 256), 256, 0, stream>>>(packed, n_frames * nr, nr, Er, best_metric,
                                                                                           best_offset);
            if (int rc = check_launch("zc_bank_unpack_kernel")) return rc;
        } else {
            // more than 64 roots: one pass per 64, results scattered into the [frames][n_roots] outputs
            float *bm_t = nullptr; int32_t *bo_t = nullptr;
            OFS_CUDA(cudaMallocAsync((void **)&bm_t, (size_t)n_frames * nr * 4, stream));
            OFS_CUDA(cudaMallocAsync((void **)&bo_t, (size_t)n_frames * nr * 4, stream));
            zc_bank_unpack_kernel<<<(unsigned)((n_frames * nr + 255) / 256), 256, 0, stream>>>(packed, n_frames * nr, nr, Er, bm_t, bo_t);
            if (int rc = check_launch("zc_bank_unpack_kernel")) return rc;
            OFS_CUDA(cudaMemcpy2DAsync(best_metric + r0, (size_t)n_roots * 4, bm_t, (size_t)nr * 4, (size_t)nr * 4, (size_t)n_frames,
                                       cudaMemcpyDeviceToDevice, stream));
            OFS_CUDA(cudaMemcpy2DAsync(best_offset + r0, (size_t)n_roots * 4, bo_t, (size_t)nr * 4, (size_t)nr * 4, (size_t)n_frames,
                                       cudaMemcpyDeviceToDevice, stream));
            OFS_CUDA(cudaFreeAsync(bm_t, stream));
            OFS_CUDA(cudaFreeAsync(bo_t, stream));
        }
    }
    OFS_CUDA(cudaFreeAsync(packed, stream));
    OFS_CUDA(cudaFreeAsync(Bmat, stream));
    OFS_CUDA(cudaFreeAsync(Er, stream));
    OFS_CUDA(cudaFreeAsync(wtab, stream));
    return OFS_OK;
}

OFS_API int ofs_zc_bank(const void *x_c64, int64_t n_frames, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                        const void *templ_c64, int32_t nbins, int32_t n_roots, float *best_metric, int32_t *best_offset,
                        void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(best_metric && best_offset, "ofs_zc_bank: null output");
    return bank_run(x_c64, n_frames, n, n_fft, cp, bins, templ_c64, nbins, n_roots, best_metric, best_offset, nullptr, 0, 1.0f, stream);
}

OFS_API int ofs_zc_freq_metric_fast(const void *x_c64, int64_t n_frames, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                                    const void *templ_c64, int32_t nbins, double templ_energy, float *metric,
                                    int64_t out_stride, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(metric && templ_energy > 0.0, "ofs_zc_freq_metric_fast: bad arguments");
    const int64_t n_off = n - ((int64_t)n_fft + cp) + 1;
    OFS_REQUIRE(n_off > 0, "Received stream is shorter than a single OFDM symbol.");   /* zc_freq.py:76-78 */
    OFS_REQUIRE(out_stride >= n_off, "ofs_zc_freq_metric_fast: out_stride < number of offsets");
    return bank_run(x_c64, n_frames, n, n_fft, cp, bins, templ_c64, nbins, 1, nullptr, nullptr, metric, out_stride, (float)templ_energy,
                    stream);
}

OFS_API int ofs_zc_freq_metric_f32(const void *x_c64, int64_t n_frames, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                                   const void *templ_c64, int32_t nbins, double templ_energy, float *metric, int64_t out_stride,
                                   void *stream_)
{
    OFS_TRACE();
    OFS_REQUIRE(x_c64 && bins && templ_c64 && metric && templ_energy > 0.0, "ofs_zc_freq_metric_f32: bad arguments");
    OFS_REQUIRE(n_fft >= 32 && n_fft <= 65536 && n_fft % 32 == 0 && cp >= 0, "ofs_zc_freq_metric_f32: n_fft must be a multiple of 32 in 32..65536");
    OFS_REQUIRE(nbins >= 1 && nbins <= 64, "ofs_zc_freq_metric_f32: nbins <= 64");
    const int64_t n_off = n - ((int64_t)n_fft + cp) + 1;
    OFS_REQUIRE(n_off > 0, "Received stream is shorter than a single OFDM symbol.");   /* zc_freq.py:76-78 */
    OFS_REQUIRE(out_stride >= n_off, "ofs_zc_freq_metric_f32: out_stride < number of offsets");
    if (n_frames == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    keep_pool_cached();
    float4 *wtab = nullptr;
    float2 *ttab = nullptr;
    OFS_CUDA(cudaMallocAsync((void **)&wtab, 64 * sizeof(float4), stream));
    OFS_CUDA(cudaMallocAsync((void **)&ttab, 64 * sizeof(float2), stream));
    zc_sdft_prep_kernel<<<1, 64, 0, stream>>>(reinterpret_cast<const float2 *>(templ_c64), nbins, bins, n_fft, wtab, ttab);
    if (int rc = check_launch("zc_sdft_prep_kernel")) return rc;
    // chains: each pays an n_fft-step warm-up, so they are >= 4 windows long; enough of them to fill the machine ~4 times
    int64_t chain_len = 4LL * n_fft;
    const int64_t want = 4LL * sm_count() * 16;                       // warps wanted
    while (chain_len > n_fft && n_frames * ((n_off + chain_len - 1) / chain_len) < want) chain_len /= 2;
    while (n_frames * ((n_off + chain_len - 1) / chain_len) > 8 * want && chain_len < n_off) chain_len *= 2;
    chain_len = (chain_len + 31) / 32 * 32;
    SdftParams p{};
    p.x = reinterpret_cast<const float2 *>(x_c64); p.n = n; p.n_off = n_off; p.chain_len = chain_len;
    p.chains_per_cap = (int)((n_off + chain_len - 1) / chain_len); p.n_chains = n_frames * p.chains_per_cap; p.mstride = out_stride;
    p.N = n_fft; p.cp = cp; p.wtab = wtab; p.ttab = ttab; p.templ_energy = (float)templ_energy; p.metric = metric;
    const int64_t grid = (p.n_chains + SD_WARPS - 1) / SD_WARPS;
    OFS_REQUIRE(grid < (1LL << 31), "ofs_zc_freq_metric_f32: grid too large");
    zc_sdft_kernel<<<(unsigned)grid, SD_WARPS * 32, 0, stream>>>(p);
    if (int rc = check_launch("zc_sdft_kernel")) return rc;
    OFS_CUDA(cudaFreeAsync(wtab, stream));
    OFS_CUDA(cudaFreeAsync(ttab, stream));
    return OFS_OK;
}

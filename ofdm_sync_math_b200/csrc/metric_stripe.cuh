// Fast ("stripe") autocorrelation-metric kernel for sm_100a -- the headline path.
//
// Replaces the per-sample Python loops of sc.py:57-78, combined_sc_min.py:144-164,
// minn.py:87-112 and sync_aa.py:458-493 for complex64 / int16-IQ input, one branch, lag
// D in {256, 512, 1024}.
//
// Design (DESIGN.md "K1"):
//  * persistent CTAs; a work unit is one stripe (S causal sample times) of one frame;
//  * the CTA walks its stripe in blocks of BK = D samples; block j+STAGES is prefetched by a 1-D
//    bulk TMA copy (cp.async.bulk + mbarrier, SASS UBLKCP) into a STAGES-deep smem ring while
//    block j is computed;
//  * thread (warp w, lane l) always owns the SAME 8 sample phases w*256 + 8l .. +7 of every block,
//    so the sample x[t-D], the thread-local prefix of the lag product at t-D and the window sums
//    at t-D, t-2D are simply the thread's own registers from the previous block(s): no smem delay
//    line, no halo re-read;
//  * sliding windows are hierarchical, never a long running prefix (bounded fp32 error):
//      W(t) = (s_cur[i] - s_prev[i])                    fp32, 8-sample thread-local prefixes
//           + (E_cur[lane] - E_prev[lane]) + G[w]        fp64, warp-scan offsets + chunk totals
//    with G[w] = sum_{w'>=w} tot_prev[w'] + sum_{w'<w} tot_cur[w'] (exactly one window of D);
//  * epilogue in registers: P, R per metric kind, M = |P|^2/R^2 (or Minn / AA variants),
//    per-256-sample chunk maxima for the detectors; M leaves through a smem staging buffer and
//    bulk TMA stores (or direct 16-byte stores, store_mode 0).
// Outputs live in causal time t (newest sample of the window): output index d = t - toff, so the
// caller passes M pointing at d = 0 and vector stores need (M - toff) to be 16-byte aligned.
#pragma once
#include "common.cuh"
#include <type_traits>
#include <cuda.h>
#include <string.h>   // CUtensorMap + cuTensorMapEncodeTiled prototype (resolved at run time, no -lcuda)

namespace ofs {

constexpr int SK = 8;            // samples per thread per block
constexpr int SCH = 32 * SK;     // samples per warp per block (chunk)
#ifndef OFS_STRIPE_THREADS
#define OFS_STRIPE_THREADS 512
#endif
constexpr int SSTAGES = 4;
constexpr int SMAXSTAGES = 8;    // mbarriers reserved in shared memory

struct StripeParams {
    const void *x;
    float *M;
    float2 *P;                // optional (WPR variant): window sums P (complex64) and R (float32), same pitch / offset as M
    float *R;
    float *chunk_max;
    int64_t L, xfs, out_stride, cm_stride;
    int64_t stripe_len;       // multiple of BK
    int64_t n_frames;
    int stripes_per_frame;
    int toff;                 // output index = t - toff
    int use_tma, store_mode;  // use_tma: 0 plain loads, 1 1-D bulk copies, 2 tiled tensor-map copies (128B swizzle)
    int aa_L;
    float aa_floor;
    int tma_mode_wanted;      // 1: 1-D bulk copies only; 2: prefer tiled tensor-map copies
    int nb;                   // branches summed before the metric (sc.py:73-74); > 1 only in the MB kernel variant
    int stages;               // MB variant: block-stages in the ring (a block-stage holds one block of every branch)
    int64_t xbs;              // branch stride in samples
};

template <int DT>
__device__ __forceinline__ void load8_smem(const unsigned char *stage, int idx0, float2 (&v)[SK]);
template <>
__device__ __forceinline__ void load8_smem<OFS_C64>(const unsigned char *stage, int idx0, float2 (&v)[SK])
{
    const float4 *p = reinterpret_cast<const float4 *>(stage + (size_t)idx0 * 8);
#pragma unroll
    for (int i = 0; i < SK / 2; ++i) {
        const float4 a = p[i];
        v[2 * i] = make_float2(a.x, a.y);
        v[2 * i + 1] = make_float2(a.z, a.w);
    }
}
template <>
__device__ __forceinline__ void load8_smem<OFS_IQ16>(const unsigned char *stage, int idx0, float2 (&v)[SK])
{
    const int4 *p = reinterpret_cast<const int4 *>(stage + (size_t)idx0 * 4);
#pragma unroll
    for (int i = 0; i < SK / 4; ++i) {
        const int4 a = p[i];
        const int w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            v[4 * i + k] = cvt_iq16((unsigned)w[k]);
    }
}

// 128-byte-swizzled stage (tensor-map TMA, CU_TENSOR_MAP_SWIZZLE_128B): the 16-byte chunk c of 128-byte row r
// lives at chunk (c ^ (r & 7)).  A thread's 64 contiguous bytes (c64) are 4 chunks of one row: the four
// LDS.128 of a quarter-warp then hit 8 distinct chunk columns x 4 row groups -> conflict-free.
template <int DT>
__device__ __forceinline__ void load8_smem_swz(const unsigned char *stage, int idx0, float2 (&v)[SK]);
template <>
__device__ __forceinline__ void load8_smem_swz<OFS_C64>(const unsigned char *stage, int idx0, float2 (&v)[SK])
{
    const unsigned o = (unsigned)idx0 * 8u;
    const unsigned row = o >> 7, c0 = (o >> 4) & 7u, sw = row & 7u;
    const unsigned char *rb = stage + (row << 7);
#pragma unroll
    for (int i = 0; i < SK / 2; ++i) {
        const float4 a = *reinterpret_cast<const float4 *>(rb + (((c0 + i) ^ sw) << 4));
        v[2 * i] = make_float2(a.x, a.y);
        v[2 * i + 1] = make_float2(a.z, a.w);
    }
}
template <>
__device__ __forceinline__ void load8_smem_swz<OFS_IQ16>(const unsigned char *stage, int idx0, float2 (&v)[SK])
{
    const unsigned o = (unsigned)idx0 * 4u;
    const unsigned row = o >> 7, c0 = (o >> 4) & 7u, sw = row & 7u;
    const unsigned char *rb = stage + (row << 7);
#pragma unroll
    for (int i = 0; i < SK / 4; ++i) {
        const int4 a = *reinterpret_cast<const int4 *>(rb + (((c0 + i) ^ sw) << 4));
        const int w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            v[4 * i + k] = cvt_iq16((unsigned)w[k]);
    }
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

template <int DT>
__device__ __forceinline__ float2 load1_gmem(const void *row, int64_t idx);
template <>
__device__ __forceinline__ float2 load1_gmem<OFS_C64>(const void *row, int64_t idx)
{
    return __ldg(reinterpret_cast<const float2 *>(row) + idx);
}
template <>
__device__ __forceinline__ float2 load1_gmem<OFS_IQ16>(const void *row, int64_t idx)
{
    const short2 s = __ldg(reinterpret_cast<const short2 *>(row) + idx);
    return make_float2((float)s.x, (float)s.y);
}
template <int DT>
__device__ __forceinline__ float2 load1_smem(const unsigned char *stage, int idx);
template <>
__device__ __forceinline__ float2 load1_smem<OFS_C64>(const unsigned char *stage, int idx)
{
    return reinterpret_cast<const float2 *>(stage)[idx];
}
template <>
__device__ __forceinline__ float2 load1_smem<OFS_IQ16>(const unsigned char *stage, int idx)
{
    const short2 s = reinterpret_cast<const short2 *>(stage)[idx];
    return make_float2((float)s.x, (float)s.y);
}

// Per-block state of one thread; two copies ping-pong (cur / prev) so that no history is ever copied.
struct BlkState {
    float2 x[SK];                     // the thread's 8 samples
    float sqr[SK], sqi[SK], se[SK];   // thread-local inclusive prefixes of the lag product / energy
    double Er, Ei;                    // warp-exclusive offsets of the product prefixes (fp64)
    float Ee;                         // warp-exclusive offset of the energy prefix (fp32: positive sums, no cancellation)
};

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// KIND: OFS_SC / OFS_SC_BOTH / OFS_MINN / OFS_AA.   WARPS: D = WARPS*256.
// WPR: also write P and R (reference-shaped outputs M, P, R at 24 B per sample; 3 CTA-slots of 128 threads per SM for the
// 24 extra registers).
// MB: several branches per frame, summed before the scan (linearity: sc.py:73-74, minn.py:106-107).  A ring slot holds one block
// of EVERY branch (p.nb tiled TMA copies on one mbarrier) and stays resident for one more block: x[t - D] of each branch is
// re-read from the previous block's slot instead of living in registers, so the register footprint does not grow with the
// branch count and the branch count is a run-time parameter.  The lag products and energies of the branches are summed in
// registers; everything after that (prefixes, warp scans, windows, epilogue) is the single-branch code.
template <int WARPS, int KIND, int DT, bool WPR, bool MB>
__global__ void __launch_bounds__(WARPS * 32, (((KIND == OFS_MINN && WPR) ? 256 : (KIND == OFS_MINN || KIND == OFS_SC_BOTH || WPR) ? 384 : OFS_STRIPE_THREADS) / (WARPS * 32)))
metric_stripe_kernel(const StripeParams p, const __grid_constant__ CUtensorMap tmap)
{
    constexpr int BK = WARPS * SCH;
    constexpr int ESZ = InT<DT>::bytes;
    constexpr int STAGE_BYTES = BK * ESZ;
    constexpr int WU = (KIND == OFS_MINN) ? 4 : (KIND == OFS_SC_BOTH ? 3 : 2);   // warm-up blocks
    constexpr bool H1 = (KIND == OFS_SC_BOTH || KIND == OFS_MINN);               // window history depth >= 1
    constexpr bool H2 = (KIND == OFS_MINN);

    extern __shared__ __align__(1024) unsigned char smem[];
    const int S = MB ? p.stages : SSTAGES;                                     // ring slots
    const int SLOT_BYTES = MB ? p.nb * STAGE_BYTES : STAGE_BYTES;              // MB: one block of every branch per slot
    unsigned char *stages = smem;                                              // S * SLOT_BYTES
    float *ost = reinterpret_cast<float *>(smem + (size_t)S * SLOT_BYTES);     // WARPS * 2 * SCH floats
    double *tot = reinterpret_cast<double *>(ost + WARPS * 2 * SCH);           // 3 * WARPS * 4 doubles
    uint64_t *bars = reinterpret_cast<uint64_t *>(tot + 3 * WARPS * 4);        // SMAXSTAGES

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // Kogge-Stone masks: 1.0 where the lane takes its neighbour's partial sum (lane >= 1,2,4,8,16)
    double mk[5];
    float mkf[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) { mk[q] = lane >= (1 << q) ? 1.0 : 0.0; mkf[q] = lane >= (1 << q) ? 1.f : 0.f; }

    const int64_t total_work = p.n_frames * (int64_t)p.stripes_per_frame;
    uint32_t git = 0;                      // global block counter of this CTA (stage slot, tot slot)
    uint32_t phase_bits = 0;               // bit s: parity of the next completion of bars[s]
    const int myoff = warp * SCH + lane * SK;   // first sample phase owned by this thread
    float *my_ost = ost + warp * 2 * SCH;

    for (int64_t work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int64_t frame = work / p.stripes_per_frame;
        const int stripe = (int)(work % p.stripes_per_frame);
        const int64_t t0 = (int64_t)stripe * p.stripe_len;
        const int64_t t1 = min(t0 + p.stripe_len, p.L);
        int64_t tb = t0 - (int64_t)WU * BK;
        if (tb < 0) tb = 0;
        const int nblk = (int)((t1 - tb + BK - 1) / BK);
        const unsigned char *xrow = reinterpret_cast<const unsigned char *>(p.x) + (size_t)frame * p.xfs * ESZ;
        float *Mrow_t = p.M ? p.M + frame * p.out_stride - p.toff : nullptr;    // indexed by causal t
        float2 *Prow_t = (WPR && p.P) ? p.P + frame * p.out_stride - p.toff : nullptr;
        float *Rrow_t = (WPR && p.R) ? p.R + frame * p.out_stride - p.toff : nullptr;
        const bool m_vec_ok = p.M && ((reinterpret_cast<uintptr_t>(Mrow_t) & 15) == 0) &&
                              (!Prow_t || (reinterpret_cast<uintptr_t>(Prow_t) & 15) == 0) &&
                              (!Rrow_t || (reinterpret_cast<uintptr_t>(Rrow_t) & 15) == 0);
        // first causal time whose output exists and is fully valid (AA: the window must be full, t >= L)
        const int64_t tlo = max(t0, (int64_t)(KIND == OFS_AA ? p.aa_L : p.toff));
        // blocks [i_fast0, i_fast1) are "steady state": fully inside [tlo, t1), loaded by one full bulk copy,
        // stored with aligned vector / bulk stores -> branch-free fast path
        int i_fast0 = (int)((tlo - tb + BK - 1) / BK), i_fast1 = (int)((t1 - tb) / BK);
        if (!p.use_tma || (p.M && !m_vec_ok)) i_fast1 = 0;
        float *cm_row = p.chunk_max ? p.chunk_max + frame * p.cm_stride : nullptr;

        // ---- reset per-stripe state (history = zeros: x[t<0] = 0) -----------------------------
        if (tid < 3 * WARPS * 4) tot[tid] = 0.0;
        __syncthreads();

        BlkState sA, sB;
        float wqr1[SK], wqi1[SK], we1[SK], wqr2[SK], wqi2[SK], we2[SK];   // window history t-D, t-2D
#pragma unroll
        for (int j = 0; j < SK; ++j) {
            sA.x[j] = sB.x[j] = make_float2(0.f, 0.f);
            sA.sqr[j] = sA.sqi[j] = sA.se[j] = sB.sqr[j] = sB.sqi[j] = sB.se[j] = 0.f;
            wqr1[j] = wqi1[j] = we1[j] = wqr2[j] = wqi2[j] = we2[j] = 0.f;
        }
        sA.Er = sA.Ei = sB.Er = sB.Ei = 0.0;
        sA.Ee = sB.Ee = 0.f;

        const int rem0 = (int)min(p.L - tb, (int64_t)0x7fffffff);      // samples from tb to the end of the frame
        const int64_t row0 = ((int64_t)frame * p.xfs + tb) * ESZ / 128;     // tensor-map row of block 0 (use_tma == 2)
        auto tma_samples = [&](int i) -> int {        // prefix of block i brought by the bulk copy
            if (!p.use_tma) return 0;
            if (p.use_tma == 2) return BK;           // tiled copies always bring the whole box (zero-filled past the tensor)
            const int rem = rem0 - i * BK;
            const int valid = rem < BK ? rem : BK;
            return (int)((((unsigned)valid * ESZ) & ~15u) / ESZ);
        };
        auto issue = [&](int i) {
            if constexpr (MB) {
                const uint32_t g = git + (uint32_t)i;
                const int slot = (int)(g % (uint32_t)S);
                uint64_t *bar = &bars[slot];
                mbar_expect_tx(bar, (uint32_t)(p.nb * STAGE_BYTES));
                for (int b = 0; b < p.nb; ++b)
                    tma_load_2d(stages + (size_t)slot * SLOT_BYTES + (size_t)b * STAGE_BYTES, &tmap, 0,
                                (int)(row0 + (int64_t)b * (p.xbs * ESZ / 128) + (int64_t)i * (STAGE_BYTES / 128)), bar);
                return;
            }
            const int ns = tma_samples(i);
            if (ns > 0) {
                const uint32_t g = git + (uint32_t)i;
                uint64_t *bar = &bars[g % SSTAGES];
                mbar_expect_tx(bar, (uint32_t)(ns * ESZ));
                if (p.use_tma == 2)
                    tma_load_2d(stages + (size_t)(g % SSTAGES) * STAGE_BYTES, &tmap, 0, (int)(row0 + (int64_t)i * (STAGE_BYTES / 128)), bar);
                else
                    tma_load_1d(stages + (size_t)(g % SSTAGES) * STAGE_BYTES,
                                xrow + (size_t)(tb + (int64_t)i * BK) * ESZ, (uint32_t)(ns * ESZ), bar);
            }
        };
        if (tid == 0) {
            for (int i = 0; i < S && i < nblk; ++i) issue(i);
        }

        // One block: `cur` is filled, `prev` is the same thread's state one block (= D samples) earlier.
        // FAST (compile-time) = steady-state block: no edge handling anywhere.
        auto block = [&](auto fast_tag, int i, BlkState &cur, const BlkState &prev) {
            constexpr bool FAST = decltype(fast_tag)::value;
            const uint32_t g = git + (uint32_t)i;
            const int st = MB ? (int)(g % (uint32_t)S) : (int)(g % SSTAGES);
            const unsigned char *stage = stages + (size_t)st * SLOT_BYTES;
            const int64_t blkpos = tb + (int64_t)i * BK;
            // ---- wait for the bulk copy, load this thread's 8 samples --------------------------
            if constexpr (MB) {
                // every branch: this block's 8 samples and the same 8 phases of the previous block (its slot is still resident),
                // lag products and energies summed over the branches in registers
                mbar_wait(&bars[st], (phase_bits >> st) & 1u);
                phase_bits ^= 1u << st;
                const unsigned char *pstage = stages + (size_t)((g + (uint32_t)S - 1u) % (uint32_t)S) * SLOT_BYTES;
#pragma unroll
                for (int j = 0; j < SK; ++j) cur.sqr[j] = cur.sqi[j] = cur.se[j] = 0.f;
                for (int b = 0; b < p.nb; ++b) {
                    float2 cx[SK], px[SK];
                    load8_smem_swz<DT>(stage + (size_t)b * STAGE_BYTES, myoff, cx);
                    if (!FAST) {
#pragma unroll
                        for (int j = 0; j < SK; ++j)            // samples past the frame end belong to the next branch / frame
                            if (blkpos + myoff + j >= p.L) cx[j] = make_float2(0.f, 0.f);
                    }
                    if (i > 0) load8_smem_swz<DT>(pstage + (size_t)b * STAGE_BYTES, myoff, px);
                    else {
#pragma unroll
                        for (int j = 0; j < SK; ++j) px[j] = make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < SK; ++j) {
                        cur.sqr[j] += fmaf(px[j].x, cx[j].x, px[j].y * cx[j].y);
                        cur.sqi[j] += fmaf(px[j].y, cx[j].x, -(px[j].x * cx[j].y));
                        cur.se[j] += fmaf(cx[j].x, cx[j].x, cx[j].y * cx[j].y);
                    }
                }
            } else if (FAST) {
                mbar_wait(&bars[st], (phase_bits >> st) & 1u);
                phase_bits ^= 1u << st;
                if (p.use_tma == 2) load8_smem_swz<DT>(stage, myoff, cur.x);
                else load8_smem<DT>(stage, myoff, cur.x);
            } else if (p.use_tma == 2) {
                mbar_wait(&bars[st], (phase_bits >> st) & 1u);
                phase_bits ^= 1u << st;
                load8_smem_swz<DT>(stage, myoff, cur.x);
#pragma unroll
                for (int j = 0; j < SK; ++j)                    // samples past the frame end belong to the next frame: zero them
                    if (blkpos + myoff + j >= p.L) cur.x[j] = make_float2(0.f, 0.f);
            } else {
                const int ns_tma = tma_samples(i);
                if (ns_tma > 0) {
                    mbar_wait(&bars[st], (phase_bits >> st) & 1u);
                    phase_bits ^= 1u << st;
                }
                if (myoff + SK <= ns_tma) {
                    load8_smem<DT>(stage, myoff, cur.x);
                } else {
#pragma unroll
                    for (int j = 0; j < SK; ++j) {
                        const int idx = myoff + j;
                        if (idx < ns_tma) cur.x[j] = load1_smem<DT>(stage, idx);
                        else if (blkpos + idx < p.L) cur.x[j] = load1_gmem<DT>(xrow, blkpos + idx);
                        else cur.x[j] = make_float2(0.f, 0.f);
                    }
                }
            }
            // ---- lag products q = x[t-D] conj(x[t]), energies, thread-local inclusive prefixes (fp32) ----
            if constexpr (!MB) {
#pragma unroll
                for (int j = 0; j < SK; ++j) {
                    cur.sqr[j] = fmaf(prev.x[j].x, cur.x[j].x, prev.x[j].y * cur.x[j].y);
                    cur.sqi[j] = fmaf(prev.x[j].y, cur.x[j].x, -(prev.x[j].x * cur.x[j].y));
                    cur.se[j] = fmaf(cur.x[j].x, cur.x[j].x, cur.x[j].y * cur.x[j].y);
                }
            }
#pragma unroll
            for (int j = 1; j < SK; ++j) {
                cur.sqr[j] += cur.sqr[j - 1];
                cur.sqi[j] += cur.sqi[j - 1];
                cur.se[j] += cur.se[j - 1];
            }
            // ---- warp inclusive scan of the thread totals: products in fp64, energy in fp32 -----------
            const double ownr = (double)cur.sqr[SK - 1], owni = (double)cur.sqi[SK - 1];
            double tr = ownr, ti = owni;
            float te = cur.se[SK - 1];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                tr = fma(shfl_up_f64(tr, 1 << q), mk[q], tr);
                ti = fma(shfl_up_f64(ti, 1 << q), mk[q], ti);
                te = fmaf(__shfl_up_sync(0xffffffffu, te, 1 << q), mkf[q], te);
            }
            cur.Er = tr - ownr; cur.Ei = ti - owni; cur.Ee = te - cur.se[SK - 1];   // exclusive lane offsets
            const int tb_cur = (int)(g % 3u), tb_prev = (int)((g + 2u) % 3u);
            if (lane == 31) {
                double *t = tot + (tb_cur * WARPS + warp) * 4;
                *reinterpret_cast<double2 *>(t) = make_double2(tr, ti);
                t[2] = (double)te;
            }
            __syncthreads();
            // stage `st` has been read by every warp: refill it with block i + SSTAGES (MB: the slot of block i - 1, which
            // block i has just read as its history, takes block i - 1 + S)
            if constexpr (MB) { if (tid == 0 && i >= 1 && i - 1 + S < nblk) issue(i - 1 + S); }
            else { if (tid == 0 && i + SSTAGES < nblk) issue(i + SSTAGES); }

            // ---- G[w]: the D-sample window that ends just before this warp's chunk --------------
            double Gr = 0.0, Gi = 0.0, Ge = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const double *t = tot + (((w < warp) ? tb_cur : tb_prev) * WARPS + w) * 4;
                const double2 ri = *reinterpret_cast<const double2 *>(t);
                Gr += ri.x; Gi += ri.y; Ge += t[2];
            }
            const float br = (float)(Gr + (cur.Er - prev.Er));
            const float bi = (float)(Gi + (cur.Ei - prev.Ei));
            const float be = (float)Ge + (cur.Ee - prev.Ee);

            // ---- windows, metric --------------------------------------------------------------
            float Mv[SK];
            float2 Pv[WPR ? SK : 1];
            float Rw[WPR ? SK : 1];
#pragma unroll
            for (int j = 0; j < SK; ++j) {
                const float wqr = br + (cur.sqr[j] - prev.sqr[j]);
                const float wqi = bi + (cur.sqi[j] - prev.sqi[j]);
                const float we = be + (cur.se[j] - prev.se[j]);
                float Pr, Pi, Rv;
                if (KIND == OFS_MINN) { Pr = wqr + wqr2[j]; Pi = wqi + wqi2[j]; Rv = we + we1[j] + we2[j]; }
                else if (KIND == OFS_SC_BOTH) { Pr = wqr; Pi = wqi; Rv = we + we1[j]; }
                else { Pr = wqr; Pi = wqi; Rv = we; }
                if (KIND == OFS_AA) {
                    const float r = rcp_approx(Rv);
                    const float q = fmaf(Pr, Pr, Pi * Pi) * (r * r);
                    Mv[j] = Rv > p.aa_floor ? fminf(q, 1.f) : 0.f;
                } else {
                    const float r = rcp_approx(fmaxf(Rv, 1e-12f));
                    const float pp = fmaxf(Pr, 0.f);
                    const float num = (KIND == OFS_MINN) ? pp * pp : fmaf(Pr, Pr, Pi * Pi);
                    Mv[j] = num * (r * r);
                }
                if (WPR) { Pv[j] = make_float2(Pr, KIND == OFS_AA ? -Pi : Pi); Rw[j] = Rv; }   // sync_aa's P is the conjugate lag product
                if (H2) { wqr2[j] = wqr1[j]; wqi2[j] = wqi1[j]; we2[j] = we1[j]; }
                if (H1) { wqr1[j] = wqr; wqi1[j] = wqi; we1[j] = we; }
            }
            const int64_t wpos = blkpos + warp * SCH;        // this warp's chunk
            if (!FAST) {
                if (blkpos < t0) return;                     // warm-up blocks produce no output
                // edge block: mask outputs that do not exist (t < toff, t >= L) or are not yet valid (AA: t < L)
#pragma unroll
                for (int j = 0; j < SK; ++j) {
                    const int64_t t = blkpos + myoff + j;
                    const bool ok = t >= p.toff && t < p.L && (KIND != OFS_AA || t >= p.aa_L);
                    Mv[j] = ok ? Mv[j] : 0.f;
                }
            }
            // ---- per-chunk maximum for the detectors ------------------------------------------------
            if (cm_row && (FAST || wpos < p.L)) {
                float cmax = fmaxf(fmaxf(fmaxf(Mv[0], Mv[1]), fmaxf(Mv[2], Mv[3])), fmaxf(fmaxf(Mv[4], Mv[5]), fmaxf(Mv[6], Mv[7])));
                cmax = fmaxf(cmax, 0.f);
                // non-negative floats order like their bit patterns: one REDUX instead of a 5-step shuffle tree
                const unsigned cbits = __reduce_max_sync(0xffffffffu, __float_as_uint(cmax));
                if (lane == 0) cm_row[wpos >> 8] = __uint_as_float(cbits);
            }
            // ---- store M ------------------------------------------------------------------------------
            if (p.M) {
                const bool full = FAST || (m_vec_ok && wpos >= p.toff && wpos + SCH <= t1);
                if (full && p.store_mode == 1) {
                    float *buf = my_ost + (i & 1) * SCH;
                    if (lane == 0) tma_store_wait_read<1>();   // the store that used this buffer is done
                    __syncwarp();
                    float4 *b4 = reinterpret_cast<float4 *>(buf + lane * SK);
                    b4[0] = make_float4(Mv[0], Mv[1], Mv[2], Mv[3]);
                    b4[1] = make_float4(Mv[4], Mv[5], Mv[6], Mv[7]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_1d(Mrow_t + wpos, buf, SCH * sizeof(float));
                        tma_store_commit();
                    }
                } else if (full) {
                    float4 *g4 = reinterpret_cast<float4 *>(Mrow_t + wpos + lane * SK);
                    g4[0] = make_float4(Mv[0], Mv[1], Mv[2], Mv[3]);
                    g4[1] = make_float4(Mv[4], Mv[5], Mv[6], Mv[7]);
                    if (WPR) {
                        if (Prow_t) {
                            float4 *p4 = reinterpret_cast<float4 *>(Prow_t + wpos + lane * SK);
#pragma unroll
                            for (int q = 0; q < SK / 2; ++q) p4[q] = make_float4(Pv[2 * q].x, Pv[2 * q].y, Pv[2 * q + 1].x, Pv[2 * q + 1].y);
                        }
                        if (Rrow_t) {
                            float4 *r4 = reinterpret_cast<float4 *>(Rrow_t + wpos + lane * SK);
                            r4[0] = make_float4(Rw[0], Rw[1], Rw[2], Rw[3]);
                            r4[1] = make_float4(Rw[4], Rw[5], Rw[6], Rw[7]);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < SK; ++j) {
                        const int64_t t = blkpos + myoff + j;
                        if (t >= p.toff && t < t1) {
                            Mrow_t[t] = Mv[j];
                            if (WPR) { if (Prow_t) Prow_t[t] = Pv[j]; if (Rrow_t) Rrow_t[t] = Rw[j]; }
                        }
                    }
                }
            }
        };

        using FastT = std::true_type;
        using SlowT = std::false_type;
        int i = 0;
        // head (warm-up / partial) blocks, one at a time, alternating the two states
        bool flip = false;
        auto step_slow = [&](int ii) { if (!flip) block(SlowT{}, ii, sA, sB); else block(SlowT{}, ii, sB, sA); flip = !flip; };
        for (; i < nblk && i < i_fast0; ++i) step_slow(i);
        if (i_fast1 > i) {
            if (flip) { block(FastT{}, i, sB, sA); flip = false; ++i; }      // realign so that the pair loop starts with sA
            for (; i + 1 < i_fast1; i += 2) {
                block(FastT{}, i, sA, sB);
                block(FastT{}, i + 1, sB, sA);
            }
            if (i < i_fast1) { block(FastT{}, i, sA, sB); flip = true; ++i; }
        }
        for (; i < nblk; ++i) step_slow(i);
        git += (uint32_t)nblk;
        // all bulk stores must have read their staging buffers before the next stripe reuses them
        if (lane == 0) tma_store_wait_read<0>();
        __syncthreads();
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// The whole input batch as a 2-D tensor of 128-byte rows; one box = one stage (STAGE_BYTES / 128 rows), 128B swizzle.
inline bool make_input_map(CUtensorMap *map, const void *x, size_t total_bytes, int stage_bytes)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {32, (cuuint64_t)(total_bytes / 128)};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {32, (cuuint32_t)(stage_bytes / 128)};
    const cuuint32_t estr[2] = {1, 1};
    if (gdim[1] == 0 || gdim[1] > 0xffffffffull || box[1] > 256) return false;
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(x), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// shared memory of one CTA: ring + bulk-store staging + chunk totals + mbarriers
template <int WARPS, int DT>
inline size_t stripe_smem_bytes(int slots, int nb)
{
    constexpr int BK = WARPS * SCH;
    return (size_t)slots * nb * BK * InT<DT>::bytes + (size_t)WARPS * 2 * SCH * sizeof(float) + (size_t)3 * WARPS * 4 * sizeof(double) +
           SMAXSTAGES * sizeof(uint64_t);
}

template <int WARPS, int KIND, int DT, bool WPR, bool MB>
int launch_one(StripeParams p, int64_t total_work, cudaStream_t stream)
{
    constexpr int BK = WARPS * SCH;
    constexpr int ESZ = InT<DT>::bytes;
    auto kern = metric_stripe_kernel<WARPS, KIND, DT, WPR, MB>;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (MB) {
        // ring depth: 4 slots (block i - 1 as history, block i, two blocks in flight) while that leaves room for two CTAs per SM
        p.stages = (size_t)4 * p.nb * BK * ESZ <= 100 * 1024 ? 4 : 3;
        const size_t smem = stripe_smem_bytes<WARPS, DT>(p.stages, p.nb);
        if (smem > 220 * 1024 || !make_input_map(&tmap, p.x, (size_t)p.n_frames * p.xfs * ESZ, BK * ESZ)) {
            set_error("ofs_metric(stripe): %d branches of lag %d do not fit the multi-branch ring", p.nb, BK);
            return OFS_EUNSUPPORTED;
        }
        p.use_tma = 2;
        int occ = 0;
        OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     // size depends on nb
        OFS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
        if (occ < 1) occ = 1;
        int64_t grid = (int64_t)sm_count() * occ;
        if (grid > total_work) grid = total_work;
        kern<<<(unsigned)grid, WARPS * 32, smem, stream>>>(p, tmap);
        return check_launch("metric_stripe_kernel(mb)");
    }
    const size_t smem = stripe_smem_bytes<WARPS, DT>(SSTAGES, 1);
    static PerDeviceOnce once;
    static int occ_dev[OFS_MAX_DEVICES];
    if (!once.done()) {
        int o = 0;
        OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OFS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, WARPS * 32, smem));
        occ_dev[current_device()] = o < 1 ? 1 : o;
        once.mark();
    }
    const int occ = occ_dev[current_device()];
    // tiled (swizzled) copies need 128-byte rows: frame pitch a multiple of 128 bytes; the batch is one tensor
    if (p.use_tma == 1 && p.tma_mode_wanted == 2 && ((size_t)p.xfs * ESZ) % 128 == 0 &&
        make_input_map(&tmap, p.x, (size_t)p.n_frames * p.xfs * ESZ, BK * ESZ))
        p.use_tma = 2;
    int64_t grid = (int64_t)sm_count() * occ;
    if (grid > total_work) grid = total_work;
    kern<<<(unsigned)grid, WARPS * 32, smem, stream>>>(p, tmap);
    return check_launch("metric_stripe_kernel");
}

// Explicit instantiations live in metric_stripe_inst_*.cu (one translation unit per metric kind, so the 60-odd kernel variants
// compile in parallel); everybody else sees only these declarations.
#define OFS_STRIPE_FOR_KIND(X, KIND)                                                                       \
    X(4, KIND, OFS_C64, false, false) X(2, KIND, OFS_C64, false, false) X(1, KIND, OFS_C64, false, false)     \
    X(4, KIND, OFS_IQ16, false, false) X(2, KIND, OFS_IQ16, false, false) X(1, KIND, OFS_IQ16, false, false)  \
    X(4, KIND, OFS_C64, true, false) X(2, KIND, OFS_C64, true, false) X(1, KIND, OFS_C64, true, false)        \
    X(4, KIND, OFS_IQ16, true, false) X(2, KIND, OFS_IQ16, true, false) X(1, KIND, OFS_IQ16, true, false)
#define OFS_STRIPE_FOR_KIND_MB(X, KIND)                                                                    \
    X(4, KIND, OFS_C64, false, true) X(2, KIND, OFS_C64, false, true) X(1, KIND, OFS_C64, false, true)        \
    X(4, KIND, OFS_IQ16, false, true) X(2, KIND, OFS_IQ16, false, true) X(1, KIND, OFS_IQ16, false, true)
#define OFS_STRIPE_DECLARE(W, KIND, DT, WPR, MB) extern template int launch_one<W, KIND, DT, WPR, MB>(StripeParams, int64_t, cudaStream_t);
#define OFS_STRIPE_DEFINE(W, KIND, DT, WPR, MB) template int launch_one<W, KIND, DT, WPR, MB>(StripeParams, int64_t, cudaStream_t);
#ifndef OFS_STRIPE_INSTANTIATE
OFS_STRIPE_FOR_KIND(OFS_STRIPE_DECLARE, OFS_SC)
OFS_STRIPE_FOR_KIND(OFS_STRIPE_DECLARE, OFS_SC_BOTH)
OFS_STRIPE_FOR_KIND(OFS_STRIPE_DECLARE, OFS_MINN)
OFS_STRIPE_FOR_KIND(OFS_STRIPE_DECLARE, OFS_AA)
OFS_STRIPE_FOR_KIND_MB(OFS_STRIPE_DECLARE, OFS_SC)
OFS_STRIPE_FOR_KIND_MB(OFS_STRIPE_DECLARE, OFS_SC_BOTH)
OFS_STRIPE_FOR_KIND_MB(OFS_STRIPE_DECLARE, OFS_MINN)
#endif

}  // namespace ofs

// K8: antenna-array [A][A] metric with the antenna combining kept on chip.
//
// Replaces sync_aa.aa_detect_streaming loop 1 (sync_aa.py:458-493: per-antenna DelayLine / RunningSum objects,
// :321-386, summed over antennas at :478-479) for captures of A antennas, complex64 or int16-IQ:
//     prod_a[k] = x_a[k] conj(x_a[k-L])  (0 for k < L)      P[n] = sum_a sum_{k=n-L+1..n} prod_a[k]
//     R[n] = sum_a sum_{k=n-L+1..n} |x_a[k]|^2              M[n] = n >= L and R > 1e-6 L ? min(|P|^2/R^2, 1) : 0
// Every sample is read from HBM once; P, R, M and the (M >= threshold) bitmask of the gate FSM (sync_aa.py:511) are
// the only things written (1/A of the input traffic).
//
// Shape of the kernel (sm_100a):
//  * time is cut into rows of L samples; a tile = r <= 8 output rows of one capture.  A CTA has L/2 threads; thread t
//    owns columns 2t, 2t+1 of EVERY row, so x[k-L] is simply the value the same thread read for the previous row and
//    the lag never crosses threads.
//  * one 1-D bulk TMA copy (cp.async.bulk + mbarrier, SASS UBLKCP) brings the (r+2) rows of ONE antenna -- contiguous in
//    memory -- into a shared-memory stage; a ring of stages keeps up to 160 KB in flight per SM independent of register
//    pressure.  Persistent CTAs walk (tile, antenna) items; consecutive tiles run on different SMs at the same time, so the
//    two halo rows of a tile are L2 hits, not second HBM reads.
//  * the sum over antennas is accumulated in fp32 registers (9 product rows x 2 columns x {re, im, energy}); after the
//    last antenna each row gets a float64 block scan (thread-local, warp shuffles, one smem exchange) and
//    P[n = iL + c] = (rowtotal(i-1) - incl(i-1)[c]) + incl(i)[c]  -- no prefix ever runs longer than one row.
#include "common.cuh"

namespace ofs {

constexpr int AR_RO = 8;             // max output rows per tile
constexpr int AR_NQ = AR_RO + 2;     // sample rows staged per antenna
constexpr int AR_SMEM = 200 * 1024;  // budget for the ring

struct ArrayParams {
    const void *x;
    float *M;
    float2 *P;
    float *R;
    unsigned *mask;
    int64_t n, xfs, xbs, out_stride, mask_stride;
    int64_t n_tiles;
    int A, r, tiles_per_cap, stages;
    double thr;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int DT>
__device__ __forceinline__ float4 ld_pair(const unsigned char *stage, int idx);   // idx = pair index (2 samples)
template <>
__device__ __forceinline__ float4 ld_pair<OFS_C64>(const unsigned char *stage, int idx)
{
    return reinterpret_cast<const float4 *>(stage)[idx];
}
template <>
__device__ __forceinline__ float4 ld_pair<OFS_IQ16>(const unsigned char *stage, int idx)
{
    const uint2 v = reinterpret_cast<const uint2 *>(stage)[idx];
    const float2 a = cvt_iq16(v.x), b = cvt_iq16(v.y);
    return make_float4(a.x, a.y, b.x, b.y);
}

template <int DT, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256 ? 2 : 1)) aa_array_kernel(const ArrayParams p)
{
    constexpr int L = 2 * NT;
    constexpr int ESZ = InT<DT>::bytes;
    constexpr int STAGE_BYTES = AR_NQ * L * ESZ;
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ uint64_t full[8], empty[8];
    __shared__ double wt[2][NW][3];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.stages;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
        mbar_fence_init();
    }
    __syncthreads();

    const int my_tiles = (int)((p.n_tiles - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int r = p.r;

    // producer (thread 0): item = (tile, antenna): one bulk copy of the valid part of rows [i0-2, i0+r) of that antenna.
    // All counters advance incrementally -- no division on the per-item path.
    int pt = 0, pa = 0, ps = 0;             // next item to issue: tile iteration, antenna, stage
    int64_t p_src = 0;                      // sample offset of (capture, antenna 0, first valid row) of tile pt
    uint32_t p_dst = 0, p_bytes = 0;
    auto producer_tile = [&]() {
        const int64_t tile = (int64_t)blockIdx.x + (int64_t)pt * gridDim.x;
        const int64_t cap = tile / p.tiles_per_cap;
        const int64_t i0 = (tile % p.tiles_per_cap) * r;
        const int64_t s_base = (i0 - 2) * L;
        const int64_t lo = s_base < 0 ? 0 : s_base;
        int64_t hi = (i0 + r) * (int64_t)L;
        if (hi > p.n) hi = p.n;
        p_src = cap * p.xfs + lo;
        p_dst = (uint32_t)((lo - s_base) * ESZ);
        p_bytes = (uint32_t)((hi - lo) * ESZ);
    };
    auto issue = [&]() {
        const unsigned char *src = reinterpret_cast<const unsigned char *>(p.x) + (size_t)(p_src + (int64_t)pa * p.xbs) * ESZ;
        mbar_expect_tx(&full[ps], p_bytes);
        tma_load_1d(ring + (size_t)ps * STAGE_BYTES + p_dst, src, p_bytes, &full[ps]);
        if (++ps == S) ps = 0;
        if (++pa == p.A) { pa = 0; ++pt; if (pt < my_tiles) producer_tile(); }
    };
    if (tid == 0 && my_tiles > 0) {
        producer_tile();
        for (int j = 0; j < S && pt < my_tiles; ++j) issue();
    }

    float qr[AR_RO + 1][2], qi[AR_RO + 1][2], en[AR_RO + 1][2];
    int s = 0;
    uint32_t par = 0;

    for (int ti = 0; ti < my_tiles; ++ti) {
        const int64_t tile = (int64_t)blockIdx.x + (int64_t)ti * gridDim.x;
        const int64_t cap = tile / p.tiles_per_cap;
        const int64_t i0 = (tile % p.tiles_per_cap) * r;
#pragma unroll
        for (int q = 0; q <= AR_RO; ++q) { qr[q][0] = qr[q][1] = qi[q][0] = qi[q][1] = en[q][0] = en[q][1] = 0.f; }
        unsigned vmask = 0;
        for (int q = 0; q < r + 2; ++q) {
            const int64_t k0 = (i0 - 2 + q) * L + 2 * tid;
            if (k0 >= 0 && k0 < p.n) vmask |= 1u << q;
        }
        const bool interior = (i0 >= 2) && ((i0 + r) * (int64_t)L <= p.n);
      for (int a = 0; a < p.A; ++a) {
        mbar_wait(&full[s], par);
        const unsigned char *st = ring + (size_t)s * STAGE_BYTES;
        if (interior) {
            float4 prev = ld_pair<DT>(st, tid);
#pragma unroll
            for (int q = 1; q < AR_NQ; ++q) {
                if (q < r + 2) {
                    const float4 cur = ld_pair<DT>(st, q * NT + tid);
                    qr[q - 1][0] = fmaf(cur.x, prev.x, fmaf(cur.y, prev.y, qr[q - 1][0]));
                    qi[q - 1][0] = fmaf(cur.y, prev.x, fmaf(-cur.x, prev.y, qi[q - 1][0]));
                    en[q - 1][0] = fmaf(cur.x, cur.x, fmaf(cur.y, cur.y, en[q - 1][0]));
                    qr[q - 1][1] = fmaf(cur.z, prev.z, fmaf(cur.w, prev.w, qr[q - 1][1]));
                    qi[q - 1][1] = fmaf(cur.w, prev.z, fmaf(-cur.z, prev.w, qi[q - 1][1]));
                    en[q - 1][1] = fmaf(cur.z, cur.z, fmaf(cur.w, cur.w, en[q - 1][1]));
                    prev = cur;
                }
            }
        } else {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 prev = (vmask & 1u) ? ld_pair<DT>(st, tid) : z;
#pragma unroll
            for (int q = 1; q < AR_NQ; ++q) {
                if (q < r + 2) {
                    const float4 cur = ((vmask >> q) & 1u) ? ld_pair<DT>(st, q * NT + tid) : z;
                    qr[q - 1][0] = fmaf(cur.x, prev.x, fmaf(cur.y, prev.y, qr[q - 1][0]));
                    qi[q - 1][0] = fmaf(cur.y, prev.x, fmaf(-cur.x, prev.y, qi[q - 1][0]));
                    en[q - 1][0] = fmaf(cur.x, cur.x, fmaf(cur.y, cur.y, en[q - 1][0]));
                    qr[q - 1][1] = fmaf(cur.z, prev.z, fmaf(cur.w, prev.w, qr[q - 1][1]));
                    qi[q - 1][1] = fmaf(cur.w, prev.z, fmaf(-cur.z, prev.w, qi[q - 1][1]));
                    en[q - 1][1] = fmaf(cur.z, cur.z, fmaf(cur.w, cur.w, en[q - 1][1]));
                    prev = cur;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (tid == 0 && pt < my_tiles) {
            mbar_wait(&empty[s], par);          // every warp is done with this stage
            issue();
        }
        if (++s == S) { s = 0; par ^= 1u; }
      }

        // ---- all antennas summed: row scans (float64) and outputs ------------------------------------
        double ptot[3] = {0.0, 0.0, 0.0}, pin0[3] = {0.0, 0.0, 0.0}, pin1[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int pr = 0; pr <= AR_RO; ++pr) {
            if (pr <= r) {
                double l0[3] = {(double)qr[pr][0], (double)qi[pr][0], (double)en[pr][0]};
                double l1[3] = {l0[0] + (double)qr[pr][1], l0[1] + (double)qi[pr][1], l0[2] + (double)en[pr][1]};
                double t[3] = {l1[0], l1[1], l1[2]};
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) scan_step_f64(t[c], o);
                }
                if (lane == 31) { wt[pr & 1][warp][0] = t[0]; wt[pr & 1][warp][1] = t[1]; wt[pr & 1][warp][2] = t[2]; }
                __syncthreads();
                double in0[3], in1[3], tot[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double base = 0.0, all = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const double v = wt[pr & 1][w][c];
                        if (w < warp) base += v;
                        all += v;
                    }
                    const double off = base + (t[c] - l1[c]);
                    in0[c] = off + l0[c]; in1[c] = off + l1[c]; tot[c] = all;
                }
                if (pr >= 1) {
                    const int64_t row = i0 + pr - 1;
                    const int64_t n0 = row * L + 2 * tid;
                    float mv[2]; float2 pv[2]; float rv[2]; unsigned fl = 0;
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const double *pi = v ? pin1 : pin0;
                        const double *ci = v ? in1 : in0;
                        const double Pr = (ptot[0] - pi[0]) + ci[0];
                        const double Pi = (ptot[1] - pi[1]) + ci[1];
                        const double Rv = (ptot[2] - pi[2]) + ci[2];
                        double Mv = 0.0;
                        if (n0 + v >= L && Rv > 1e-6 * (double)L) {        // sync_aa.py:486-493
                            Mv = (Pr * Pr + Pi * Pi) / (Rv * Rv);
                            Mv = Mv < 1.0 ? Mv : 1.0;
                        }
                        mv[v] = (float)Mv; pv[v] = make_float2((float)Pr, (float)Pi); rv[v] = (float)Rv;
                        if (n0 + v < p.n && n0 + v >= L && (double)mv[v] >= p.thr) fl |= 1u << v;
                    }
                    if (n0 < p.n) {        // n is even: the pair is in range together
                        const int64_t o = cap * p.out_stride + n0;
                        if (p.M) *reinterpret_cast<float2 *>(p.M + o) = make_float2(mv[0], mv[1]);
                        if (p.P) *reinterpret_cast<float4 *>(p.P + o) = make_float4(pv[0].x, pv[0].y, pv[1].x, pv[1].y);
                        if (p.R) *reinterpret_cast<float2 *>(p.R + o) = make_float2(rv[0], rv[1]);
                    }
                    if (p.mask) {
                        // 32 samples = 16 lanes x 2 bits: lanes 0-15 -> word 2*warp, lanes 16-31 -> word 2*warp + 1
                        const unsigned sh = fl << (2 * (lane & 15));
                        const unsigned w_lo = __reduce_or_sync(0xffffffffu, lane < 16 ? sh : 0u);
                        const unsigned w_hi = __reduce_or_sync(0xffffffffu, lane >= 16 ? sh : 0u);
                        const int64_t wi = row * (L / 32) + 2 * warp;
                        if (lane == 0 && wi * 32 < p.n) {
                            unsigned *mw = p.mask + cap * p.mask_stride + wi;
                            mw[0] = w_lo;
                            if ((wi + 1) * 32 < p.n) mw[1] = w_hi;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) { ptot[c] = tot[c]; pin0[c] = in0[c]; pin1[c] = in1[c]; }
            }
        }
    }
}

bool array_supported(int in_dtype, int64_t n, int64_t xfs, int64_t xbs, int L, const void *x)
{
    if (in_dtype != OFS_C64 && in_dtype != OFS_IQ16) return false;
    if (L != 128 && L != 256 && L != 512 && L != 1024) return false;
    const int64_t esz = in_dtype == OFS_C64 ? 8 : 4;
    if ((n * esz) % 16 || (xfs * esz) % 16 || (xbs * esz) % 16 || n < 2) return false;
    if (x && (reinterpret_cast<uintptr_t>(x) & 15)) return false;
    return true;
}

// ring depth: ~160 KB in flight per SM.  Large CTAs (L >= 512) take the whole SM and deepen their ring (int16 rows are half
// the bytes: 8 stages); short rows keep 4 stages and share the SM between several CTAs instead.
static int array_stages(int L, int esz)
{
    // two CTAs per SM (L <= 512), ~80 KB of ring each
    const int stage_bytes = AR_NQ * L * esz;
    int st = (L >= 1024 ? 160 * 1024 : 80 * 1024) / stage_bytes;
    return st > 8 ? 8 : (st < 2 ? 2 : st);
}

static int array_ctas_per_sm(int L, int esz)
{
    const int stage_bytes = AR_NQ * L * esz;
    const int smem = array_stages(L, esz) * stage_bytes + 1024;
    int c = (220 * 1024) / smem;
    const int by_regs = 65536 / ((L / 2) * (L >= 1024 ? 128 : 128));
    if (c > by_regs) c = by_regs;
    return c < 1 ? 1 : c;
}

template <int DT, int NT>
static int launch_array_t(ArrayParams &p, cudaStream_t stream)
{
    constexpr int L = 2 * NT;
    constexpr int stage_bytes = AR_NQ * L * InT<DT>::bytes;
    // ring: ~160 KB in flight per SM; short rows (small CTAs) share an SM instead of deepening one ring
    const int stages = array_stages(L, InT<DT>::bytes);
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes;
    static PerDeviceOnce once;
    if (!once.done()) {
        OFS_CUDA(cudaFuncSetAttribute(aa_array_kernel<DT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, AR_SMEM));
        once.mark();
    }
    const int64_t slots = (int64_t)sm_count() * array_ctas_per_sm(L, InT<DT>::bytes);
    const int64_t grid = p.n_tiles < slots ? p.n_tiles : slots;
    aa_array_kernel<DT, NT><<<(unsigned)grid, NT, smem, stream>>>(p);
    return check_launch("aa_array_kernel");
}

// M, P, R (each optional, float32 / complex64, row pitch out_stride) and the optional above-threshold bitmask
// (uint32 words, row pitch mask_stride words, bit i%32 of word i/32 = (i >= L && M[i] >= thr)).
int launch_metric_array(const void *x, int in_dtype, int64_t n_frames, int n_ant, int64_t n, int64_t xfs, int64_t xbs, int L,
                        float *M, void *P, float *R, int64_t out_stride, unsigned *mask, int64_t mask_stride, double thr,
                        cudaStream_t stream)
{
    if (n_frames <= 0 || n <= 0) return OFS_OK;
    OFS_REQUIRE(array_supported(in_dtype, n, xfs, xbs, L, x), "aa array kernel: unsupported geometry (L in {128,256,512,1024}, 16-byte aligned rows)");
    OFS_REQUIRE(out_stride % 2 == 0 && out_stride >= n, "aa array kernel: out_stride must be even and >= n");
    OFS_REQUIRE((!M || !(reinterpret_cast<uintptr_t>(M) & 7)) && (!P || !(reinterpret_cast<uintptr_t>(P) & 15)) &&
                    (!R || !(reinterpret_cast<uintptr_t>(R) & 7)),
                "aa array kernel: misaligned output");
    OFS_REQUIRE(!mask || mask_stride >= (n + 31) / 32, "aa array kernel: mask_stride too small");
    ArrayParams p{};
    p.x = x; p.M = M; p.P = (float2 *)P; p.R = R; p.mask = mask;
    p.n = n; p.xfs = xfs; p.xbs = xbs; p.out_stride = out_stride; p.mask_stride = mask_stride;
    p.A = n_ant; p.thr = thr;
    // rows per tile: fewest (rows + 2 halo rows) x tiles-per-CTA
    const int64_t nrows = (n + L - 1) / L;
    const int64_t G = (int64_t)sm_count() * array_ctas_per_sm(L, in_dtype == OFS_C64 ? 8 : 4);
    int best_r = AR_RO;
    int64_t best_cost = -1;
    for (int r = AR_RO; r >= 3; --r) {
        const int64_t tpc = (nrows + r - 1) / r;
        const int64_t tiles = tpc * n_frames;
        const int64_t per_cta = (tiles + G - 1) / G;
        const int64_t cost = per_cta * (r + 2);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_r = r; }
    }
    if (nrows < best_r) best_r = (int)(nrows < 1 ? 1 : nrows);
    p.r = best_r;
    p.tiles_per_cap = (int)((nrows + best_r - 1) / best_r);
    p.n_tiles = (int64_t)p.tiles_per_cap * n_frames;
#define OFS_ARR(DT)                                                     \
    switch (L) {                                                        \
    case 128: return launch_array_t<DT, 64>(p, stream);                 \
    case 256: return launch_array_t<DT, 128>(p, stream);                \
    case 512: return launch_array_t<DT, 256>(p, stream);                \
    default: return launch_array_t<DT, 512>(p, stream);                 \
    }
    if (in_dtype == OFS_C64) { OFS_ARR(OFS_C64) }
    OFS_ARR(OFS_IQ16)
#undef OFS_ARR
}

}  // namespace ofs

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
#include "common.cuh"

namespace ofs {

constexpr int SK = 8;            // samples per thread per block
constexpr int SCH = 32 * SK;     // samples per warp per block (chunk)
constexpr int SSTAGES = 4;

struct StripeParams {
    const void *x;
    float *M;
    float *chunk_max;
    int64_t L, xfs, out_stride, cm_stride;
    int64_t stripe_len;       // multiple of BK
    int64_t n_frames;
    int stripes_per_frame;
    int toff;                 // output index = t - toff
    int use_tma, store_mode;
    int aa_L;
    float aa_floor;
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
            v[4 * i + k] = make_float2((float)(short)(w[k] & 0xffff), (float)(short)(w[k] >> 16));
    }
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

// KIND: OFS_SC / OFS_SC_BOTH / OFS_MINN / OFS_AA.   WARPS: D = WARPS*256.
template <int WARPS, int KIND, int DT>
__global__ void __launch_bounds__(WARPS * 32, ((KIND == OFS_MINN ? 384 : 512) / (WARPS * 32)))
metric_stripe_kernel(StripeParams p)
{
    constexpr int BK = WARPS * SCH;
    constexpr int ESZ = InT<DT>::bytes;
    constexpr int STAGE_BYTES = BK * ESZ;
    constexpr int WU = (KIND == OFS_MINN) ? 4 : (KIND == OFS_SC_BOTH ? 3 : 2);   // warm-up blocks
    constexpr bool H1 = (KIND == OFS_SC_BOTH || KIND == OFS_MINN);               // window history depth >= 1
    constexpr bool H2 = (KIND == OFS_MINN);

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *stages = smem;                                              // SSTAGES * STAGE_BYTES
    float *ost = reinterpret_cast<float *>(smem + SSTAGES * STAGE_BYTES);      // WARPS * 2 * SCH floats
    double *tot = reinterpret_cast<double *>(ost + WARPS * 2 * SCH);           // 3 * WARPS * 4 doubles
    uint64_t *bars = reinterpret_cast<uint64_t *>(tot + 3 * WARPS * 4);        // SSTAGES

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SSTAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

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
        const bool m_vec_ok = p.M && ((reinterpret_cast<uintptr_t>(Mrow_t) & 15) == 0);

        // ---- reset per-stripe state (history = zeros: x[t<0] = 0) -----------------------------
        if (tid < 3 * WARPS * 4) tot[tid] = 0.0;
        __syncthreads();

        float2 xh[SK];
        float sqr_h[SK], sqi_h[SK], se_h[SK];      // previous block: thread-local prefixes
        double Er_h = 0.0, Ei_h = 0.0, Ee_h = 0.0;  // previous block: warp-exclusive offsets
        float wqr1[SK], wqi1[SK], we1[SK], wqr2[SK], wqi2[SK], we2[SK];   // window history t-D, t-2D
#pragma unroll
        for (int j = 0; j < SK; ++j) {
            xh[j] = make_float2(0.f, 0.f);
            sqr_h[j] = sqi_h[j] = se_h[j] = 0.f;
            wqr1[j] = wqi1[j] = we1[j] = wqr2[j] = wqi2[j] = we2[j] = 0.f;
        }

        auto valid_samples = [&](int i) -> int {      // samples of block i that exist in the frame
            const int64_t rem = p.L - (tb + (int64_t)i * BK);
            return (int)(rem < BK ? rem : BK);
        };
        auto tma_samples = [&](int i) -> int {        // prefix of block i brought by the bulk copy
            if (!p.use_tma) return 0;
            return (int)((((size_t)valid_samples(i) * ESZ) & ~(size_t)15) / ESZ);
        };
        auto issue = [&](int i) {
            const int ns = tma_samples(i);
            if (ns > 0) {
                const uint32_t g = git + (uint32_t)i;
                uint64_t *bar = &bars[g % SSTAGES];
                mbar_expect_tx(bar, (uint32_t)(ns * ESZ));
                tma_load_1d(stages + (size_t)(g % SSTAGES) * STAGE_BYTES,
                            xrow + (size_t)(tb + (int64_t)i * BK) * ESZ, (uint32_t)(ns * ESZ), bar);
            }
        };
        if (tid == 0) {
            for (int i = 0; i < SSTAGES && i < nblk; ++i) issue(i);
        }

        for (int i = 0; i < nblk; ++i) {
            const uint32_t g = git + (uint32_t)i;
            const int st = g % SSTAGES;
            const unsigned char *stage = stages + (size_t)st * STAGE_BYTES;
            const int64_t blkpos = tb + (int64_t)i * BK;
            const int ns_tma = tma_samples(i);
            if (ns_tma > 0) {
                mbar_wait(&bars[st], (phase_bits >> st) & 1u);
                phase_bits ^= 1u << st;
            }

            // ---- load this thread's 8 samples -------------------------------------------------
            float2 xc[SK];
            if (myoff + SK <= ns_tma) {
                load8_smem<DT>(stage, myoff, xc);
            } else {
#pragma unroll
                for (int j = 0; j < SK; ++j) {
                    const int idx = myoff + j;
                    if (idx < ns_tma) xc[j] = load1_smem<DT>(stage, idx);
                    else if (blkpos + idx < p.L) xc[j] = load1_gmem<DT>(xrow, blkpos + idx);
                    else xc[j] = make_float2(0.f, 0.f);
                }
            }

            // ---- lag products, energies, thread-local inclusive prefixes (fp32) -----------------
            float sqr[SK], sqi[SK], se[SK];
#pragma unroll
            for (int j = 0; j < SK; ++j) {
                // q = x[t-D] * conj(x[t])
                sqr[j] = fmaf(xh[j].x, xc[j].x, xh[j].y * xc[j].y);
                sqi[j] = fmaf(xh[j].y, xc[j].x, -(xh[j].x * xc[j].y));
                se[j] = fmaf(xc[j].x, xc[j].x, xc[j].y * xc[j].y);
            }
#pragma unroll
            for (int j = 1; j < SK; ++j) {
                sqr[j] += sqr[j - 1];
                sqi[j] += sqi[j - 1];
                se[j] += se[j - 1];
            }
            // ---- warp inclusive scan of the thread totals, fp64 ---------------------------------
            const double ownr = (double)sqr[SK - 1], owni = (double)sqi[SK - 1], owne = (double)se[SK - 1];
            double tr = ownr, ti = owni, te = owne;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double yr = shfl_up_f64(tr, o), yi = shfl_up_f64(ti, o), ye = shfl_up_f64(te, o);
                if (lane >= o) { tr += yr; ti += yi; te += ye; }
            }
            const double Er = tr - ownr, Ei = ti - owni, Ee = te - owne;   // exclusive lane offsets
            const int tb_cur = (int)(g % 3u), tb_prev = (int)((g + 2u) % 3u);
            if (lane == 31) {
                double *t = tot + (tb_cur * WARPS + warp) * 4;
                t[0] = tr; t[1] = ti; t[2] = te;
            }
            __syncthreads();
            // stage `st` has been read by every warp: refill it with block i + SSTAGES
            if (tid == 0 && i + SSTAGES < nblk) issue(i + SSTAGES);

            // ---- G[w]: the D-sample window that ends just before this warp's chunk --------------
            double Gr = 0.0, Gi = 0.0, Ge = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const double *t = tot + (((w < warp) ? tb_cur : tb_prev) * WARPS + w) * 4;
                Gr += t[0]; Gi += t[1]; Ge += t[2];
            }
            const float br = (float)(Gr + (Er - Er_h));
            const float bi = (float)(Gi + (Ei - Ei_h));
            const float be = (float)(Ge + (Ee - Ee_h));

            // ---- windows, metric --------------------------------------------------------------
            const bool emit = blkpos >= t0;       // warm-up blocks produce no output
            float Mv[SK];
            float cmax = 0.f;
            const int64_t tpos = blkpos + myoff;
#pragma unroll
            for (int j = 0; j < SK; ++j) {
                const float wqr = br + (sqr[j] - sqr_h[j]);
                const float wqi = bi + (sqi[j] - sqi_h[j]);
                const float we = be + (se[j] - se_h[j]);
                float Pr, Pi, Rv;
                if (KIND == OFS_MINN) { Pr = wqr + wqr2[j]; Pi = wqi + wqi2[j]; Rv = we + we1[j] + we2[j]; }
                else if (KIND == OFS_SC_BOTH) { Pr = wqr; Pi = wqi; Rv = we + we1[j]; }
                else { Pr = wqr; Pi = wqi; Rv = we; }
                float m;
                if (KIND == OFS_AA) {
                    m = 0.f;
                    if (tpos + j >= p.aa_L && Rv > p.aa_floor) {
                        m = fmaf(Pr, Pr, Pi * Pi) / (Rv * Rv);
                        m = fminf(m, 1.f);
                    }
                } else {
                    const float rr = fmaxf(Rv, 1e-12f);
                    const float num = (KIND == OFS_MINN) ? fmaxf(Pr, 0.f) * fmaxf(Pr, 0.f) : fmaf(Pr, Pr, Pi * Pi);
                    m = num / (rr * rr);
                }
                const int64_t t = tpos + j;
                const bool ok = emit && t >= p.toff && t < p.L;
                Mv[j] = ok ? m : 0.f;
                cmax = fmaxf(cmax, Mv[j]);
                if (H2) { wqr2[j] = wqr1[j]; wqi2[j] = wqi1[j]; we2[j] = we1[j]; }
                if (H1) { wqr1[j] = wqr; wqi1[j] = wqi; we1[j] = we; }
                sqr_h[j] = sqr[j]; sqi_h[j] = sqi[j]; se_h[j] = se[j];
                xh[j] = xc[j];
            }
            Er_h = Er; Ei_h = Ei; Ee_h = Ee;

            if (emit) {
                // ---- per-chunk maximum for the detectors ------------------------------------------
                if (p.chunk_max && blkpos + warp * SCH < p.L) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
                    if (lane == 0) p.chunk_max[frame * p.cm_stride + (blkpos + warp * SCH) / SCH] = cmax;
                }
                // ---- store M ----------------------------------------------------------------------
                if (p.M) {
                    const int64_t wpos = blkpos + warp * SCH;
                    const bool full = m_vec_ok && wpos >= p.toff && wpos + SCH <= t1;
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
                        float4 *g4 = reinterpret_cast<float4 *>(Mrow_t + tpos);
                        g4[0] = make_float4(Mv[0], Mv[1], Mv[2], Mv[3]);
                        g4[1] = make_float4(Mv[4], Mv[5], Mv[6], Mv[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < SK; ++j) {
                            const int64_t t = tpos + j;
                            if (t >= p.toff && t < t1) Mrow_t[t] = Mv[j];
                        }
                    }
                }
            }
        }
        git += (uint32_t)nblk;
        // all bulk stores must have read their staging buffers before the next stripe reuses them
        if (lane == 0) tma_store_wait_read<0>();
        __syncthreads();
    }
}

template <int WARPS, int KIND, int DT>
static int launch_one(const StripeParams &p, int64_t total_work, cudaStream_t stream)
{
    constexpr int BK = WARPS * SCH;
    constexpr int ESZ = InT<DT>::bytes;
    const size_t smem = (size_t)SSTAGES * BK * ESZ + (size_t)WARPS * 2 * SCH * sizeof(float) +
                        (size_t)3 * WARPS * 4 * sizeof(double) + SSTAGES * sizeof(uint64_t);
    auto kern = metric_stripe_kernel<WARPS, KIND, DT>;
    static bool attr_set = false;
    static int occ = 0;
    if (!attr_set) {
        OFS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OFS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
        if (occ < 1) occ = 1;
        attr_set = true;
    }
    int64_t grid = (int64_t)sm_count() * occ;
    if (grid > total_work) grid = total_work;
    kern<<<(unsigned)grid, WARPS * 32, smem, stream>>>(p);
    return check_launch("metric_stripe_kernel");
}

template <int KIND, int DT>
static int launch_by_lag(int D, const StripeParams &p, int64_t total_work, cudaStream_t stream)
{
    switch (D) {
    case 1024: return launch_one<4, KIND, DT>(p, total_work, stream);
    case 512: return launch_one<2, KIND, DT>(p, total_work, stream);
    case 256: return launch_one<1, KIND, DT>(p, total_work, stream);
    default: set_error("ofs_metric(stripe): lag %d not in {256,512,1024}", D); return OFS_EUNSUPPORTED;
    }
}

template <int DT>
static int launch_by_kind(int kind, int D, const StripeParams &p, int64_t total_work, cudaStream_t stream)
{
    switch (kind) {
    case OFS_SC: return launch_by_lag<OFS_SC, DT>(D, p, total_work, stream);
    case OFS_SC_BOTH: return launch_by_lag<OFS_SC_BOTH, DT>(D, p, total_work, stream);
    case OFS_MINN: return launch_by_lag<OFS_MINN, DT>(D, p, total_work, stream);
    case OFS_AA: return launch_by_lag<OFS_AA, DT>(D, p, total_work, stream);
    default: set_error("ofs_metric(stripe): unknown kind %d", kind); return OFS_EINVAL;
    }
}

static int stripe_lag(const ofs_metric_desc *d)
{
    switch (d->kind) {
    case OFS_SC: case OFS_SC_BOTH: return (d->symbol_len % 2 == 0) ? d->symbol_len / 2 : -1;
    case OFS_MINN: return (d->symbol_len % 4 == 0) ? d->symbol_len / 4 : -1;
    case OFS_AA: return d->symbol_len;
    default: return -1;
    }
}

bool stripe_supported(const ofs_metric_desc *d)
{
    const int D = stripe_lag(d);
    return (D == 256 || D == 512 || D == 1024) && d->n_branches == 1 && d->out_f64 == 0 &&
           (d->in_dtype == OFS_C64 || d->in_dtype == OFS_IQ16) && ofs_metric_out_len(d) > 0;
}

int launch_metric_stripe(const ofs_metric_desc *d, const void *x, float *M, float *chunk_max, int64_t cm_stride,
                         cudaStream_t stream)
{
    if (!stripe_supported(d)) {
        set_error("ofs_metric(stripe): descriptor not supported by the stripe path");
        return OFS_EUNSUPPORTED;
    }
    const int D = stripe_lag(d);
    const int esz = d->in_dtype == OFS_C64 ? 8 : 4;
    StripeParams p{};
    p.x = x; p.M = M; p.chunk_max = chunk_max;
    p.L = d->n_samples; p.xfs = d->x_frame_stride; p.out_stride = d->out_stride; p.cm_stride = cm_stride;
    p.n_frames = d->n_frames;
    p.toff = d->kind == OFS_AA ? 0 : d->symbol_len - 1;
    p.store_mode = d->store_mode;
    p.aa_L = d->symbol_len; p.aa_floor = 1e-6f * (float)d->symbol_len;
    // bulk copies need 16-byte aligned sources: base pointer and frame pitch
    p.use_tma = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (((size_t)d->x_frame_stride * esz) % 16 == 0);
    // stripe length: a multiple of D, >= 8 warm-up-amortising blocks, aiming at >= ~6 work units per CTA slot
    const int64_t nblk_frame = (d->n_samples + D - 1) / D;
    const int64_t slots = (int64_t)sm_count() * 4;
    int64_t per_frame = (6 * slots + d->n_frames - 1) / d->n_frames;     // stripes per frame wanted
    if (per_frame < 1) per_frame = 1;
    int64_t blk_per_stripe = (nblk_frame + per_frame - 1) / per_frame;
    if (blk_per_stripe < 32) blk_per_stripe = 32;                         // warm-up <= 4/32 of the work
    if (blk_per_stripe > nblk_frame) blk_per_stripe = nblk_frame;
    p.stripe_len = blk_per_stripe * D;
    p.stripes_per_frame = (int)((d->n_samples + p.stripe_len - 1) / p.stripe_len);
    const int64_t total_work = p.n_frames * (int64_t)p.stripes_per_frame;
    if (total_work <= 0) return OFS_OK;
    if (d->in_dtype == OFS_C64) return launch_by_kind<OFS_C64>(d->kind, D, p, total_work, stream);
    return launch_by_kind<OFS_IQ16>(d->kind, D, p, total_work, stream);
}

}  // namespace ofs

// Fast ("stripe") autocorrelation-metric path: host-side dispatch.  The kernel lives in metric_stripe.cuh, its explicit
// instantiations in metric_stripe_inst_*.cu.
#include "metric_stripe.cuh"

namespace ofs {

template <int KIND, int DT>
static int launch_by_lag(int D, const StripeParams &p, int64_t total_work, cudaStream_t stream)
{
    if (p.nb > 1) {
        if constexpr (KIND != OFS_AA) {
            switch (D) {
            case 1024: return launch_one<4, KIND, DT, false, true>(p, total_work, stream);
            case 512: return launch_one<2, KIND, DT, false, true>(p, total_work, stream);
            case 256: return launch_one<1, KIND, DT, false, true>(p, total_work, stream);
            default: break;
            }
        }
        set_error("ofs_metric(stripe): multi-branch needs kind SC / SC_BOTH / MINN and lag in {256,512,1024}");
        return OFS_EUNSUPPORTED;
    }
    if (p.P || p.R) {
        switch (D) {
        case 1024: return launch_one<4, KIND, DT, true, false>(p, total_work, stream);
        case 512: return launch_one<2, KIND, DT, true, false>(p, total_work, stream);
        case 256: return launch_one<1, KIND, DT, true, false>(p, total_work, stream);
        default: set_error("ofs_metric(stripe): lag %d not in {256,512,1024}", D); return OFS_EUNSUPPORTED;
        }
    }
    switch (D) {
    case 1024: return launch_one<4, KIND, DT, false, false>(p, total_work, stream);
    case 512: return launch_one<2, KIND, DT, false, false>(p, total_work, stream);
    case 256: return launch_one<1, KIND, DT, false, false>(p, total_work, stream);
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

// One branch: any pitch / alignment (the kernel falls back from tiled to 1-D bulk copies to plain loads).  Several branches
// (summed before the metric): kinds SC / SC_BOTH / MINN, M only, and the tiled TMA ring -- frame and branch pitch multiples of
// 128 bytes, 128-byte aligned base, a ring of >= 3 slots of nb blocks within shared memory.
bool stripe_supported(const ofs_metric_desc *d, const void *x, bool want_pr)
{
    const int D = stripe_lag(d);
    if (!((D == 256 || D == 512 || D == 1024) && d->out_f64 == 0 && (d->in_dtype == OFS_C64 || d->in_dtype == OFS_IQ16) &&
          ofs_metric_out_len(d) > 0))
        return false;
    if (d->n_branches == 1) return true;
    const size_t esz = d->in_dtype == OFS_C64 ? 8 : 4;
    return d->n_branches <= 8 && d->kind != OFS_AA && !want_pr && x && (reinterpret_cast<uintptr_t>(x) & 127) == 0 &&
           ((size_t)d->x_frame_stride * esz) % 128 == 0 && ((size_t)d->x_branch_stride * esz) % 128 == 0 &&
           (size_t)3 * d->n_branches * D * esz <= 200 * 1024;
}

int launch_metric_stripe(const ofs_metric_desc *d, const void *x, float *M, void *P, float *R, float *chunk_max, int64_t cm_stride,
                         cudaStream_t stream)
{
    if (!stripe_supported(d, x, P || R)) {
        set_error("ofs_metric(stripe): descriptor not supported by the stripe path");
        return OFS_EUNSUPPORTED;
    }
    const int D = stripe_lag(d);
    const int esz = d->in_dtype == OFS_C64 ? 8 : 4;
    StripeParams p{};
    p.x = x; p.M = M; p.P = (float2 *)P; p.R = R; p.chunk_max = chunk_max;
    p.L = d->n_samples; p.xfs = d->x_frame_stride; p.out_stride = d->out_stride; p.cm_stride = cm_stride;
    p.n_frames = d->n_frames;
    p.nb = d->n_branches; p.xbs = d->x_branch_stride; p.stages = SSTAGES;
    p.toff = d->kind == OFS_AA ? 0 : d->symbol_len - 1;
    p.store_mode = (P || R) ? 0 : d->store_mode;          // the bulk-store variant carries M only
    p.aa_L = d->symbol_len; p.aa_floor = 1e-6f * (float)d->symbol_len;
    // bulk copies need 16-byte aligned sources: base pointer and frame pitch
    p.use_tma = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (((size_t)d->x_frame_stride * esz) % 16 == 0);
    p.tma_mode_wanted = (d->reserved == 1) ? 1 : 2;      // desc.reserved = 1 forces the 1-D bulk-copy path (A/B testing)
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

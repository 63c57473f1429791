// C-ABI entry points of libofdmsync: library/error plumbing, ofs_metric dispatch and the
// end-to-end "sync metric + CFO" pipeline (device and host-buffer versions).
#include <cstdlib>
#include "common.cuh"
#include "exact.cuh"
#include <string.h>
#include <new>

namespace ofs {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int64_t &launch_counter() { return g_launches; }

int current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = 0; }
    return dev < 0 ? 0 : (dev >= OFS_MAX_DEVICES ? OFS_MAX_DEVICES - 1 : dev);
}

int sm_count()
{
    static int cached[OFS_MAX_DEVICES] = {0};
    const int dev = current_device();
    int v = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
    if (v == 0) {
        int n = 0;
        v = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
        (void)cudaGetLastError();
        __atomic_store_n(&cached[dev], v, __ATOMIC_RELAXED);
    }
    return v;
}

void keep_pool_cached()
{
    static PerDeviceOnce once;
    if (once.done()) return;
    once.mark();
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    (void)cudaGetLastError();
}

int launch_metric_tile(const ofs_metric_desc *d, const void *x, void *M, void *P, void *R, cudaStream_t stream);
int launch_metric_stripe(const ofs_metric_desc *d, const void *x, float *M, void *P, float *R, float *chunk_max, int64_t cm_stride,
                         cudaStream_t stream);
bool stripe_supported(const ofs_metric_desc *d, const void *x, bool want_pr);
int launch_plateau(const ofs_rows *M, const float *chunk_max, int64_t cm_stride, int32_t toff, int32_t cp_len, int32_t lookahead,
                   int32_t smooth_win, int64_t *plateau_end, const ExactSrc *ex, int32_t *status, void *stream);
int launch_minn_peak(const ofs_rows *M, const float *chunk_max, int64_t cm_stride, int32_t toff, int32_t smooth_win,
                     double gate_threshold, int32_t has_bounds, int64_t bound_lo, int64_t bound_hi, int64_t *peak,
                     int64_t *gate_span, void *Ms, const ExactSrc *ex, int32_t *status, void *stream);
bool array_supported(int in_dtype, int64_t n, int64_t xfs, int64_t xbs, int L, const void *x);
int launch_metric_array(const void *x, int in_dtype, int64_t n_frames, int n_ant, int64_t n, int64_t xfs, int64_t xbs, int L,
                        float *M, void *P, float *R, int64_t out_stride, unsigned *mask, int64_t mask_stride, double thr,
                        cudaStream_t stream);

static int check_desc(const ofs_metric_desc *d, const char *who)
{
    OFS_REQUIRE(d, "%s: null descriptor", who);
    OFS_REQUIRE(d->kind >= OFS_SC && d->kind <= OFS_AA, "%s: unknown metric kind %d", who, d->kind);
    OFS_REQUIRE(d->in_dtype >= OFS_C64 && d->in_dtype <= OFS_IQ16, "%s: unknown input dtype %d", who, d->in_dtype);
    OFS_REQUIRE(d->symbol_len > 0, "%s: symbol_len must be positive", who);
    OFS_REQUIRE(d->kind == OFS_AA || d->kind == OFS_MINN || d->symbol_len % 2 == 0,
                "%s: symbol_len must be even for Schmidl-Cox", who);
    OFS_REQUIRE(d->n_branches >= 1 && d->n_frames >= 0 && d->n_samples >= 0, "%s: bad batch geometry", who);
    OFS_REQUIRE(d->x_branch_stride >= d->n_samples || d->n_branches == 1, "%s: branch stride < n_samples", who);
    OFS_REQUIRE(d->x_frame_stride >= d->n_samples, "%s: frame stride < n_samples", who);
    return OFS_OK;
}

// ---- P at one index per frame, in float64, + record (one warp per frame) -------------------------
__global__ void sync_record_kernel(const void *x, int dtype, int64_t L, int64_t xfs, int64_t xbs, int nb, int kind, int N,
                                   const float *M, const double *M64, int64_t out_stride, int64_t out_len, const int64_t *timing,
                                   const int32_t *status, int status_or, int sc_delta, ofs_sync_record *rec, int64_t n_frames)
{
    const int64_t frame = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (frame >= n_frames) return;
    int64_t t = timing[frame];
    int64_t coarse = t;
    if (kind != OFS_MINN) coarse = t - sc_delta > 0 ? t - sc_delta : 0;     // sc.py:211
    if (coarse > out_len - 1) coarse = out_len - 1;
    if (coarse < 0) coarse = 0;
    const size_t esz = dtype == OFS_C64 ? 8 : (dtype == OFS_C128 ? 16 : 4);
    double pr = 0.0, pi = 0.0;
    const int lag = kind == OFS_MINN ? N / 4 : N / 2;
    const int nwin = kind == OFS_MINN ? 2 : 1;
    for (int b = 0; b < nb; ++b) {                                            // branches summed (sc.py:73-74)
        const void *xr = reinterpret_cast<const unsigned char *>(x) + ((size_t)frame * xfs + (size_t)b * xbs) * esz;
        for (int wdw = 0; wdw < nwin; ++wdw) {
            const int64_t base = coarse + (int64_t)wdw * 2 * lag;
#pragma unroll 8
            for (int m = lane; m < lag; m += 32) {      // 8 independent sample pairs in flight per lane (the loop is pure load latency)
                const double2 a = load_sample_f64(xr, dtype, base + m);
                const double2 c = load_sample_f64(xr, dtype, base + lag + m);
                pr += a.x * c.x + a.y * c.y;
                pi += a.y * c.x - a.x * c.y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { pr += shfl_xor_f64(pr, o); pi += shfl_xor_f64(pi, o); }
    if (lane == 0) {
        ofs_sync_record r;
        r.timing = t; r.coarse = coarse;
        r.metric = M ? M[frame * out_stride + coarse] : (M64 ? (float)M64[frame * out_stride + coarse] : 0.f);
        r.p_re = (float)pr; r.p_im = (float)pi;
        r.cfo = (float)(-atan2(pi, pr) / (2.0 * 3.14159265358979323846 * (double)lag));
        r.status = (status ? status[frame] : 0) | status_or;
        r.reserved = 0;
        rec[frame] = r;
    }
}

static int check_sync_params(const ofs_sync_params *sp, const char *who)
{
    OFS_REQUIRE(sp, "%s: null parameters", who);
    OFS_REQUIRE(sp->cp_len >= 0 && sp->smooth_win >= 0 && sp->sc_delta >= 0, "%s: negative detector parameter", who);
    OFS_REQUIRE(sp->exact == 0 || sp->exact == 1, "%s: exact must be 0 or 1", who);
    OFS_REQUIRE(!(sp->exact_band > 0.01), "%s: exact_band must be <= 0.01", who);
    return OFS_OK;
}

}  // namespace ofs

using namespace ofs;

OFS_API int ofs_version(void) { return OFS_ABI_VERSION; }
OFS_API const char *ofs_last_error_string(void) { return g_err; }
OFS_API int64_t ofs_launch_count(void) { return g_launches; }
OFS_API int32_t ofs_chunk_len(void) { return 256; }

OFS_API int64_t ofs_metric_out_len(const ofs_metric_desc *d)
{
    if (!d) return 0;
    if (d->kind == OFS_AA) return d->n_samples;
    const int64_t n = d->n_samples - d->symbol_len + 1;
    return n > 0 ? n : 0;
}

OFS_API int ofs_metric_stripe_ok(const ofs_metric_desc *d, const void *x, const void *M)
{
    (void)x; (void)M;
    return d && check_desc(d, "ofs_metric_stripe_ok") == OFS_OK && stripe_supported(d, x, false) ? 1 : 0;
}

OFS_API int ofs_metric_array_ok(const ofs_metric_desc *d, const void *x)
{
    return d && check_desc(d, "ofs_metric_array_ok") == OFS_OK && d->kind == OFS_AA && !d->out_f64 &&
                   array_supported(d->in_dtype, d->n_samples, d->x_frame_stride, d->x_branch_stride, d->symbol_len, x)
               ? 1 : 0;
}

OFS_API int ofs_metric(const ofs_metric_desc *d, const void *x, void *M, void *P, void *R, float *chunk_max,
                       int64_t cm_stride, void *stream)
{
    OFS_TRACE();
    if (int rc = check_desc(d, "ofs_metric")) return rc;
    const int64_t out_len = ofs_metric_out_len(d);
    if (out_len == 0 || d->n_frames == 0) return OFS_OK;
    OFS_REQUIRE(x, "ofs_metric: null input");
    OFS_REQUIRE(M || P || R || chunk_max, "ofs_metric: no output requested");
    OFS_REQUIRE(d->out_stride >= out_len, "ofs_metric: out_stride %lld < out_len %lld", (long long)d->out_stride,
                (long long)out_len);
    int path = d->path;
    const bool array_ok = d->kind == OFS_AA && !d->out_f64 && !chunk_max && d->out_stride % 2 == 0 &&
                          array_supported(d->in_dtype, d->n_samples, d->x_frame_stride, d->x_branch_stride, d->symbol_len, x) &&
                          !(reinterpret_cast<uintptr_t>(M) & 7) && !(reinterpret_cast<uintptr_t>(P) & 15) &&
                          !(reinterpret_cast<uintptr_t>(R) & 7);
    if (path == OFS_PATH_AUTO) {
        if (stripe_supported(d, x, false) && !P && !R) path = OFS_PATH_STRIPE;
        else if (array_ok && d->n_branches >= 2) path = OFS_PATH_ARRAY;
        else path = OFS_PATH_TILE;
    }
    if (path == OFS_PATH_ARRAY) {
        OFS_REQUIRE(array_ok, "ofs_metric: the array path needs kind AA, c64/iq16 input, float32 outputs, L in {128,256,512,1024}, "
                              "16-byte aligned rows and an even out_stride");
        return launch_metric_array(x, d->in_dtype, d->n_frames, d->n_branches, d->n_samples, d->x_frame_stride, d->x_branch_stride,
                                   d->symbol_len, (float *)M, P, (float *)R, d->out_stride, nullptr, 0, 0.0, (cudaStream_t)stream);
    }
    if (path == OFS_PATH_STRIPE) {
        OFS_REQUIRE(M, "ofs_metric: the stripe path always writes M");
        OFS_REQUIRE(!chunk_max || cm_stride >= (d->n_samples + 255) / 256, "ofs_metric: cm_stride too small");
        return launch_metric_stripe(d, x, (float *)M, P, (float *)R, chunk_max, cm_stride, (cudaStream_t)stream);
    }
    OFS_REQUIRE(path == OFS_PATH_TILE, "ofs_metric: unknown path %d", path);
    OFS_REQUIRE(!chunk_max, "ofs_metric: chunk_max is produced by the stripe path only");
    return launch_metric_tile(d, x, M, P, R, (cudaStream_t)stream);
}

OFS_API int ofs_sync_detect(const ofs_metric_desc *d, const void *x, const float *M, const float *chunk_max,
                            int64_t cm_stride, const ofs_sync_params *sp, ofs_sync_record *records, int64_t *scratch,
                            void *stream)
{
    OFS_TRACE();
    if (int rc = check_desc(d, "ofs_sync_detect")) return rc;
    if (int rc = check_sync_params(sp, "ofs_sync_detect")) return rc;
    OFS_REQUIRE(d->kind == OFS_SC || d->kind == OFS_SC_BOTH || d->kind == OFS_MINN, "ofs_sync: kind must be SC or MINN");
    OFS_REQUIRE(x && M && records && scratch, "ofs_sync: null argument");
    const int64_t out_len = ofs_metric_out_len(d);
    OFS_REQUIRE(out_len > 0, "ofs_sync: frames shorter than one symbol");
    if (d->n_frames == 0) return OFS_OK;
    ofs_rows rows{M, 0, 0, d->n_frames, out_len, d->out_stride};
    int64_t *timing = scratch;
    int32_t *status = reinterpret_cast<int32_t *>(scratch + 3 * d->n_frames);
    const int toff = d->symbol_len - 1;
    ExactSrc ex{};
    if (sp->exact) {
        ex.x = x; ex.dtype = d->in_dtype; ex.kind = d->kind; ex.N = d->symbol_len; ex.nb = d->n_branches;
        ex.L = d->n_samples; ex.xfs = d->x_frame_stride; ex.xbs = d->x_branch_stride;
        ex.band = sp->exact_band > 0.0 ? sp->exact_band : OFS_EXACT_BAND;
    }
    if (d->kind == OFS_MINN) {
        if (int rc = launch_minn_peak(&rows, chunk_max, cm_stride, toff, sp->smooth_win, sp->gate_threshold, 0, 0, 0, timing,
                                      scratch + d->n_frames, nullptr, sp->exact ? &ex : nullptr, status, stream))
            return rc;
    } else {
        if (int rc = launch_plateau(&rows, chunk_max, cm_stride, toff, sp->cp_len, sp->cp_len / 4, sp->smooth_win, timing,
                                    sp->exact ? &ex : nullptr, status, stream))
            return rc;
    }
    const int wpb = 4;
    sync_record_kernel<<<(unsigned)((d->n_frames + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        x, d->in_dtype, d->n_samples, d->x_frame_stride, d->x_branch_stride, d->n_branches, d->kind, d->symbol_len, M, nullptr,
        d->out_stride, out_len, timing, status, 0, sp->sc_delta, records, d->n_frames);
    return check_launch("sync_record_kernel");
}

// The same pipeline entirely in float64 (tile metric kernel with float64 outputs -> float64 detector -> records): what a frame
// flagged OFS_ST_UNRESOLVED is re-run through, and the yardstick of the exact mode in the tests.  Workspace from the stream's pool.
OFS_API int ofs_sync_f64(const ofs_metric_desc *d, const void *x, const ofs_sync_params *sp, ofs_sync_record *records, void *stream_)
{
    OFS_TRACE();
    if (int rc = check_desc(d, "ofs_sync_f64")) return rc;
    if (int rc = check_sync_params(sp, "ofs_sync_f64")) return rc;
    OFS_REQUIRE(d->kind == OFS_SC || d->kind == OFS_SC_BOTH || d->kind == OFS_MINN, "ofs_sync_f64: kind must be SC or MINN");
    OFS_REQUIRE(x && records, "ofs_sync_f64: null argument");
    const int64_t out_len = ofs_metric_out_len(d);
    OFS_REQUIRE(out_len > 0, "ofs_sync_f64: frames shorter than one symbol");
    if (d->n_frames == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    keep_pool_cached();
    double *M64 = nullptr;
    int64_t *scratch = nullptr;
    OFS_CUDA(cudaMallocAsync((void **)&M64, (size_t)d->n_frames * out_len * sizeof(double), stream));
    OFS_CUDA(cudaMallocAsync((void **)&scratch, (size_t)d->n_frames * 3 * sizeof(int64_t), stream));
    ofs_metric_desc dd = *d;
    dd.out_f64 = 1; dd.out_stride = out_len; dd.path = OFS_PATH_TILE;
    int rc = launch_metric_tile(&dd, x, M64, nullptr, nullptr, stream);
    ofs_rows rows{M64, 1, 0, d->n_frames, out_len, out_len};
    if (!rc) {
        if (d->kind == OFS_MINN)
            rc = launch_minn_peak(&rows, nullptr, 0, 0, sp->smooth_win, sp->gate_threshold, 0, 0, 0, scratch, scratch + d->n_frames,
                                  nullptr, nullptr, nullptr, stream_);
        else
            rc = launch_plateau(&rows, nullptr, 0, 0, sp->cp_len, sp->cp_len / 4, sp->smooth_win, scratch, nullptr, nullptr, stream_);
    }
    if (!rc) {
        const int wpb = 4;
        sync_record_kernel<<<(unsigned)((d->n_frames + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
            x, d->in_dtype, d->n_samples, d->x_frame_stride, d->x_branch_stride, d->n_branches, d->kind, d->symbol_len, nullptr, M64,
            out_len, out_len, scratch, nullptr, OFS_ST_EXACT, sp->sc_delta, records, d->n_frames);
        rc = check_launch("sync_record_kernel");
    }
    OFS_CUDA(cudaFreeAsync(M64, stream));
    OFS_CUDA(cudaFreeAsync(scratch, stream));
    return rc;
}

OFS_API int ofs_sync(const ofs_metric_desc *d, const void *x, float *M, float *chunk_max, int64_t cm_stride,
                     const ofs_sync_params *sp, ofs_sync_record *records, int64_t *scratch, void *stream)
{
    OFS_TRACE();
    if (int rc = check_desc(d, "ofs_sync")) return rc;
    OFS_REQUIRE(x && M && records && scratch, "ofs_sync: null argument");
    OFS_REQUIRE(d->out_f64 == 0, "ofs_sync: float32 metric only");
    if (d->n_frames == 0) return OFS_OK;
    if (int rc = ofs_metric(d, x, M, nullptr, nullptr, chunk_max, cm_stride, stream)) return rc;
    return ofs_sync_detect(d, x, M, chunk_max, cm_stride, sp, records, scratch, stream);
}

// ---- host-buffer pipeline -------------------------------------------------------------------------
struct ofs_ctx {
    int device;
    cudaStream_t s_h2d, s_comp, s_d2h;
    cudaEvent_t ev_h2d[2], ev_comp[2], ev_d2h[2];
    void *x_dev[2];
    float *m_dev[2];
    float *cm_dev[2];
    ofs_sync_record *rec_dev[2];
    int64_t *scratch[2];
    size_t x_cap, m_cap, cm_cap, rec_cap;
};

OFS_API int ofs_ctx_create(ofs_ctx **out, int device)
{
    OFS_REQUIRE(out, "ofs_ctx_create: null output");
    OFS_CUDA(cudaSetDevice(device));
    ofs_ctx *c = new (std::nothrow) ofs_ctx();
    OFS_REQUIRE(c, "ofs_ctx_create: out of host memory");
    memset(c, 0, sizeof(*c));
    c->device = device;
    OFS_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    OFS_CUDA(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
    OFS_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        OFS_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
        OFS_CUDA(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
        OFS_CUDA(cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
    }
    *out = c;
    return OFS_OK;
}

static void ctx_free_buffers(ofs_ctx *c)
{
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->x_dev[i]); cudaFree(c->m_dev[i]); cudaFree(c->cm_dev[i]); cudaFree(c->rec_dev[i]); cudaFree(c->scratch[i]);
        c->x_dev[i] = nullptr; c->m_dev[i] = nullptr; c->cm_dev[i] = nullptr; c->rec_dev[i] = nullptr; c->scratch[i] = nullptr;
    }
    c->x_cap = c->m_cap = c->cm_cap = c->rec_cap = 0;
}

OFS_API void ofs_ctx_destroy(ofs_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    ctx_free_buffers(c);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(c->ev_h2d[i]); cudaEventDestroy(c->ev_comp[i]); cudaEventDestroy(c->ev_d2h[i]); }
    cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_comp); cudaStreamDestroy(c->s_d2h);
    delete c;
}

OFS_API void *ofs_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { set_error("ofs_host_alloc(%zu) failed", bytes); return nullptr; }
    return p;
}
OFS_API void ofs_host_free(void *p) { if (p) cudaFreeHost(p); }

OFS_API int ofs_sync_host(ofs_ctx *c, const ofs_metric_desc *d, const void *x_host, float *M_host, const ofs_sync_params *sp,
                          ofs_sync_record *records_host)
{
    OFS_TRACE();
    OFS_REQUIRE(c, "ofs_sync_host: null context");
    if (int rc = check_desc(d, "ofs_sync_host")) return rc;
    if (int rc = check_sync_params(sp, "ofs_sync_host")) return rc;
    OFS_REQUIRE(x_host && records_host, "ofs_sync_host: null argument");
    OFS_REQUIRE(d->n_branches == 1, "ofs_sync_host: one branch per frame");
    OFS_REQUIRE(d->x_frame_stride >= d->n_samples || d->n_frames <= 1, "ofs_sync_host: frame stride (in samples) < n_samples");
    OFS_REQUIRE(!M_host || d->out_stride >= ofs_metric_out_len(d), "ofs_sync_host: out_stride < out_len");
    const int64_t out_len = ofs_metric_out_len(d);
    OFS_REQUIRE(out_len > 0, "ofs_sync_host: frames shorter than one symbol");
    if (d->n_frames == 0) return OFS_OK;
    OFS_CUDA(cudaSetDevice(c->device));
    const size_t esz = dtype_bytes(d->in_dtype);
    const int64_t L = d->n_samples;
    const int toff = d->symbol_len - 1;
    const int64_t xpitch = (L + 3) / 4 * 4;                 // samples; keeps every frame 16-byte aligned
    const int64_t mpitch = (L + 3) / 4 * 4;                 // floats, causal-time rows (d = t - toff)
    const int64_t cmpitch = (L + 255) / 256;
    // Frames per pipeline batch.  PCIe is the bound of this call (8 B/sample in, 4 B/sample out, the kernels need < 2 % of the
    // copy time), so what matters is the fill (first H2D) and the drain (last D2H), both proportional to the batch: 32 MB of
    // samples per batch keeps them near 0.5 ms each while a batch is still ~0.6 ms of DMA, far above the launch overheads
    // (measured, profiles/r1_e2e_probe.jsonl: 256 MB 6.02, 64 MB 6.41, 32 MB 6.55, 16 MB 6.56, 8 MB 5.99 Gsamples/s; plain H2D copy 6.95).
    int64_t batch_mb = 32;
    if (const char *e = getenv("OFS_HOST_BATCH_MB")) { const long v = atol(e); if (v >= 1 && v <= 4096) batch_mb = v; }
    int64_t FB = (int64_t)((uint64_t)batch_mb << 20) / (int64_t)(xpitch * esz);
    if (FB < 1) FB = 1;
    if (FB > d->n_frames) FB = d->n_frames;
    const size_t x_need = (size_t)FB * xpitch * esz, m_need = (size_t)FB * mpitch * sizeof(float) + 64;
    const size_t cm_need = (size_t)FB * cmpitch * sizeof(float), rec_need = (size_t)FB;
    if (x_need > c->x_cap || m_need > c->m_cap || cm_need > c->cm_cap || rec_need > c->rec_cap) {
        cudaDeviceSynchronize();
        ctx_free_buffers(c);
        for (int i = 0; i < 2; ++i) {
            OFS_CUDA(cudaMalloc(&c->x_dev[i], x_need));
            OFS_CUDA(cudaMalloc((void **)&c->m_dev[i], m_need));
            OFS_CUDA(cudaMalloc((void **)&c->cm_dev[i], cm_need));
            OFS_CUDA(cudaMalloc((void **)&c->rec_dev[i], rec_need * sizeof(ofs_sync_record)));
            OFS_CUDA(cudaMalloc((void **)&c->scratch[i], rec_need * 4 * sizeof(int64_t)));
        }
        c->x_cap = x_need; c->m_cap = m_need; c->cm_cap = cm_need; c->rec_cap = rec_need;
    }
    const int64_t nbatch = (d->n_frames + FB - 1) / FB;
    for (int64_t b = 0; b < nbatch; ++b) {
        const int s = (int)(b & 1);
        const int64_t f0 = b * FB, nf = (f0 + FB <= d->n_frames) ? FB : d->n_frames - f0;
        // slot reuse: x_dev[s] was read by the kernels of batch b-2, m_dev[s] by its D2H copies
        if (b >= 2) { OFS_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[s], 0)); }
        OFS_CUDA(cudaMemcpy2DAsync(c->x_dev[s], (size_t)xpitch * esz,
                                   (const unsigned char *)x_host + (size_t)f0 * d->x_frame_stride * esz,
                                   (size_t)d->x_frame_stride * esz, (size_t)L * esz, (size_t)nf, cudaMemcpyHostToDevice, c->s_h2d));
        OFS_CUDA(cudaEventRecord(c->ev_h2d[s], c->s_h2d));
        OFS_CUDA(cudaStreamWaitEvent(c->s_comp, c->ev_h2d[s], 0));
        if (b >= 2) { OFS_CUDA(cudaStreamWaitEvent(c->s_comp, c->ev_d2h[s], 0)); }
        ofs_metric_desc dd = *d;
        dd.n_frames = nf; dd.x_frame_stride = xpitch; dd.out_stride = mpitch; dd.path = OFS_PATH_AUTO;
        float *Md0 = c->m_dev[s] + toff;      // element d = 0 of frame 0; (Md0 - toff) is 256-byte aligned
        if (int rc = ofs_sync(&dd, c->x_dev[s], Md0, c->cm_dev[s], cmpitch, sp, c->rec_dev[s], c->scratch[s], c->s_comp))
            return rc;
        OFS_CUDA(cudaEventRecord(c->ev_comp[s], c->s_comp));
        OFS_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[s], 0));
        if (M_host)
            OFS_CUDA(cudaMemcpy2DAsync(M_host + (size_t)f0 * d->out_stride, (size_t)d->out_stride * sizeof(float), Md0,
                                       (size_t)mpitch * sizeof(float), (size_t)out_len * sizeof(float), (size_t)nf,
                                       cudaMemcpyDeviceToHost, c->s_d2h));
        OFS_CUDA(cudaMemcpyAsync(records_host + f0, c->rec_dev[s], (size_t)nf * sizeof(ofs_sync_record), cudaMemcpyDeviceToHost,
                                 c->s_d2h));
        OFS_CUDA(cudaEventRecord(c->ev_d2h[s], c->s_d2h));
    }
    OFS_CUDA(cudaStreamSynchronize(c->s_d2h));
    OFS_CUDA(cudaStreamSynchronize(c->s_comp));
    // frames whose decision the kernels could not settle inside the float32 band: once more through the float64 pipeline
    // (rare -- a flat-topped or noiseless metric; the metric row M_host stays the float32 one)
    if (sp->exact) {
        for (int64_t f = 0; f < d->n_frames; ++f) {
            if (!(records_host[f].status & OFS_ST_UNRESOLVED)) continue;
            OFS_CUDA(cudaMemcpyAsync(c->x_dev[0], (const unsigned char *)x_host + (size_t)f * d->x_frame_stride * esz, (size_t)L * esz,
                                     cudaMemcpyHostToDevice, c->s_comp));
            ofs_metric_desc d1 = *d;
            d1.n_frames = 1; d1.x_frame_stride = L; d1.x_branch_stride = L;
            if (int rc = ofs_sync_f64(&d1, c->x_dev[0], sp, c->rec_dev[0], c->s_comp)) return rc;
            OFS_CUDA(cudaMemcpyAsync(records_host + f, c->rec_dev[0], sizeof(ofs_sync_record), cudaMemcpyDeviceToHost, c->s_comp));
            OFS_CUDA(cudaStreamSynchronize(c->s_comp));
        }
    }
    return OFS_OK;
}

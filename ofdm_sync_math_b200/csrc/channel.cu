// the stages either side of the detectors.
//
// K9  impairment chain  channel.apply_channel (channel.py:51-98: CIR by np.convolve, AWGN scaled to the faded power of
//                       each branch) -> core.apply_cfo (core.py:123-138) -> sync_aa.quantize_adc (sync_aa.py:263-291).
//                       The 1100-tap FIR runs through K4's overlap-save FFT (zc.cu) with the taps as the filter; the rest
//                       is one fused pass: out = (faded + std * noise) * e^{j 2 pi cfo n / fs}, optionally quantised to
//                       int16 IQ (the ingest format of the stripe / array kernels).  Unit noise is an INPUT (host RNG for
//                       parity with numpy's generator, device Philox for throughput): the library stays deterministic.
// K10 CP-correlation CFO estimators  core.estimate_cfo_from_cp / _robust / _peak(_with_index) / find_cp_start_via_corr
//                       (core.py:179-336), batched: one CTA per frame, lag-N products once, float64 prefix sums, every
//                       candidate offset a difference of two taps.
#include "common.cuh"

namespace ofs {

template <typename T> struct CT;
template <> struct CT<float> { using type = float2; };
template <> struct CT<double> { using type = double2; };

// ref[k] = conj(taps[nr-1-k]): the matched-filter kernel convolves with conj(ref[::-1]) = taps
__global__ void chan_ref_kernel(const double2 *taps, int nr, double2 *ref)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nr) { const double2 t = taps[nr - 1 - k]; ref[k] = make_double2(t.x, -t.y); }
}

// mean |x|^2 of every row (channel.py:74): float64 accumulation, one CTA per row
template <typename T>
__global__ void __launch_bounds__(256) chan_power_kernel(const void *x, int64_t n, int64_t stride, double *power)
{
    using C2 = typename CT<T>::type;
    __shared__ double red[8];
    const C2 *row = reinterpret_cast<const C2 *>(x) + (int64_t)blockIdx.x * stride;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) { const C2 v = row[i]; acc += (double)v.x * v.x + (double)v.y * v.y; }
    for (int o = 16; o > 0; o >>= 1) acc += shfl_xor_f64(acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        power[blockIdx.x] = t / (double)n;
    }
}

struct ImpairParams {
    const void *faded, *noise;
    const int32_t *row_of_stream;
    const double *power, *snr_db, *cfo_hz, *full_scale;
    void *out;
    short2 *out_iq;
    int64_t n, faded_stride, noise_stride, out_stride;
    double fs;
    int bits;
};

template <typename T>
__global__ void __launch_bounds__(256) chan_impair_kernel(const ImpairParams p)
{
    using C2 = typename CT<T>::type;
    constexpr int PER = 4;                                     // samples per thread: 1024 per CTA
    __shared__ double s_std, s_ratio;
    const int64_t s = blockIdx.y;
    const int64_t row = p.row_of_stream ? p.row_of_stream[s] : s;
    if (threadIdx.x == 0) {
        // channel.py:55-75: noise_std = sqrt(signal_power / snr_linear / 2); all-zero rows get no noise
        double sd = 0.0;
        if (p.noise) { const double pw = p.power[row]; if (pw > 0.0) sd = sqrt(pw / pow(10.0, p.snr_db[s] / 10.0) / 2.0); }
        s_std = sd;
        s_ratio = p.cfo_hz ? p.cfo_hz[s] / p.fs : 0.0;
    }
    __syncthreads();
    const double std = s_std;
    const double fsc = p.full_scale ? p.full_scale[s] : 1.0, levels = (double)(1 << (p.bits - 1));
    const double hi = 1.0 - 1.0 / levels;
    const C2 *frow = reinterpret_cast<const C2 *>(p.faded) + row * p.faded_stride;
    const C2 *nrow = p.noise ? reinterpret_cast<const C2 *>(p.noise) + s * p.noise_stride : nullptr;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int64_t i = ((int64_t)blockIdx.x * PER + k) * 256 + threadIdx.x;
        if (i >= p.n) break;
        const C2 f = frow[i];
        double vr = (double)f.x, vi = (double)f.y;
        if (nrow && std > 0.0) { const C2 nz = nrow[i]; vr += std * (double)nz.x; vi += std * (double)nz.y; }
        if (p.cfo_hz) {
            // core.py:131-132: tone = exp(1j * 2 pi cfo n / fs)
            double sn, cs;
            if (sizeof(T) == 8) {
                sincos(2.0 * 3.14159265358979323846 * p.cfo_hz[s] * (double)i / p.fs, &sn, &cs);
            } else {
                const double turns = s_ratio * (double)i;
                float fs_, fc_;
                sincospif((float)(2.0 * (turns - rint(turns))), &fs_, &fc_);
                sn = fs_; cs = fc_;
            }
            const double r = vr * cs - vi * sn;
            vi = vr * sn + vi * cs;
            vr = r;
        }
        if (p.full_scale) {
            // sync_aa.py:276-289: x / full_scale -> clip [-1, 1 - 1/levels] -> round half to even (np.round) -> rescale
            double a = vr / fsc, b = vi / fsc;
            a = a < -1.0 ? -1.0 : (a > hi ? hi : a);
            b = b < -1.0 ? -1.0 : (b > hi ? hi : b);
            const double qa = rint(a * levels), qb = rint(b * levels);
            if (p.out_iq) p.out_iq[s * p.out_stride + i] = make_short2((short)qa, (short)qb);
            vr = qa / levels * fsc; vi = qb / levels * fsc;
        }
        if (p.out) {
            C2 o; o.x = (T)vr; o.y = (T)vi;
            reinterpret_cast<C2 *>(p.out)[s * p.out_stride + i] = o;
        }
    }
}

// ---- CP-correlation CFO estimators ------------------------------------------------------------------------------
constexpr int CFO_NT = 256;
constexpr int CFO_MAXQ = 12288;      // products held in shared memory (float64 complex prefix: 192 KB)

struct CfoParams {
    const void *x;
    const int64_t *starts;
    int64_t n_frames, n, xfs, xbs;
    int nb, n_fft, cp_len, span, win, mode;
    double fs;
    double *cfo;
    int64_t *best_d;
    double2 *P;
};

template <int DT>
__global__ void __launch_bounds__(CFO_NT) cp_cfo_kernel(const CfoParams p)
{
    using In = typename InT<DT>::type;
    extern __shared__ __align__(16) unsigned char csm[];
    double2 *S = reinterpret_cast<double2 *>(csm);            // inclusive prefix, S[0] = 0
    __shared__ double wtot[2][CFO_NT / 32];
    __shared__ double sh_v[CFO_NT / 32];
    __shared__ long long sh_i[CFO_NT / 32];
    const int64_t frame = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t start = p.starts[frame];
    const int64_t L = p.n;
    // candidate offsets [d_lo, d_hi) and window length W per mode (core.py:213-222, 246-251, 283-288, 321-325)
    int64_t d_lo = start, d_hi = start + 1;
    int W = p.cp_len;
    bool fallback = false;
    if (p.mode == 1) {
        W = p.win;
        d_lo = start - p.span > 0 ? start - p.span : 0;
        d_hi = L - (p.n_fft + W) < start + p.span ? L - (p.n_fft + W) : start + p.span;
        if (d_hi <= d_lo) { fallback = true; W = p.cp_len < p.win ? p.cp_len : p.win; }
    } else if (p.mode == 2) {
        d_lo = start - p.span > 0 ? start - p.span : 0;
        d_hi = L - (p.n_fft + W) < start + p.span ? L - (p.n_fft + W) : start + p.span;
        if (d_hi <= d_lo) fallback = true;
    }
    if (fallback || p.mode == 0) { d_lo = start; d_hi = start + 1; }
    const int64_t nq = d_hi - d_lo + W - 1;                    // products needed: m in [d_lo, d_lo + nq)
    const bool valid = d_lo >= 0 && d_lo + nq + p.n_fft <= L && nq <= CFO_MAXQ && nq > 0;
    if (!valid) {                                              // the reference would raise (mismatched slices)
        if (tid == 0) {
            p.cfo[frame] = __longlong_as_double(0x7ff8000000000000LL);
            if (p.best_d) p.best_d[frame] = -1;
            if (p.P) p.P[frame] = make_double2(0.0, 0.0);
        }
        return;
    }
    const unsigned char *xf = reinterpret_cast<const unsigned char *>(p.x) + (size_t)frame * p.xfs * InT<DT>::bytes;
    // products q[m] = sum_b x_b[m] conj(x_b[m + N]) (float64: exact products for c64 / iq16 input), thread-serial prefix
    const int ipt = (int)((nq + CFO_NT - 1) / CFO_NT);
    const int64_t s0 = (int64_t)tid * ipt, s1 = s0 + ipt < nq ? s0 + ipt : nq;
    double ar = 0.0, ai = 0.0;
    for (int64_t m = s0; m < s1; ++m) {
        double qr = 0.0, qi = 0.0;
        for (int b = 0; b < p.nb; ++b) {
            const In *xb = reinterpret_cast<const In *>(xf) + (size_t)b * p.xbs;
            const In a = xb[d_lo + m], c = xb[d_lo + m + p.n_fft];
            qr += (double)a.x * c.x + (double)a.y * c.y;
            qi += (double)a.y * c.x - (double)a.x * c.y;
        }
        ar += qr; ai += qi;
        S[m + 1] = make_double2(ar, ai);
    }
    if (tid == 0) S[0] = make_double2(0.0, 0.0);
    double tr = ar, ti = ai;
    for (int o = 1; o < 32; o <<= 1) {
        const double yr = shfl_up_f64(tr, o), yi = shfl_up_f64(ti, o);
        if (lane >= o) { tr += yr; ti += yi; }
    }
    if (lane == 31) { wtot[0][warp] = tr; wtot[1][warp] = ti; }
    __syncthreads();
    double offr = tr - ar, offi = ti - ai;
    for (int w = 0; w < warp; ++w) { offr += wtot[0][w]; offi += wtot[1][w]; }
    for (int64_t m = s0; m < s1; ++m) { S[m + 1].x += offr; S[m + 1].y += offi; }
    __syncthreads();

    const int64_t nd = d_hi - d_lo;
    double Pr = 0.0, Pi = 0.0;
    long long bd = start;
    if (p.mode == 1 && !fallback) {
        // sum over d of P(d) (core.py:223-227)
        double sr = 0.0, si = 0.0;
        for (int64_t d = tid; d < nd; d += CFO_NT) { sr += S[d + W].x - S[d].x; si += S[d + W].y - S[d].y; }
        for (int o = 16; o > 0; o >>= 1) { sr += shfl_xor_f64(sr, o); si += shfl_xor_f64(si, o); }
        __syncthreads();
        if (lane == 0) { wtot[0][warp] = sr; wtot[1][warp] = si; }
        __syncthreads();
        for (int w = 0; w < CFO_NT / 32; ++w) { Pr += wtot[0][w]; Pi += wtot[1][w]; }
    } else if (p.mode == 2 && !fallback) {
        // first maximum of |P(d)| (strict >, core.py:254-261)
        double bv = -1.0; long long bi = LLONG_MAX;
        for (int64_t d = tid; d < nd; d += CFO_NT) {
            const double r = S[d + W].x - S[d].x, i = S[d + W].y - S[d].y;
            const double v = hypot(r, i);
            if (v > bv) { bv = v; bi = d; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = shfl_xor_f64(bv, o);
            const long long oi = shfl_xor_i64(bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sh_v[warp] = bv; sh_i[warp] = bi; }
        __syncthreads();
        bv = sh_v[0]; bi = sh_i[0];
        for (int w = 1; w < CFO_NT / 32; ++w)
            if (sh_v[w] > bv || (sh_v[w] == bv && sh_i[w] < bi)) { bv = sh_v[w]; bi = sh_i[w]; }
        Pr = S[bi + W].x - S[bi].x; Pi = S[bi + W].y - S[bi].y;
        bd = d_lo + bi;
    } else {
        Pr = S[W].x - S[0].x; Pi = S[W].y - S[0].y;
    }
    if (tid == 0) {
        // angle(P) = -2 pi f N / fs  =>  f = -angle fs / (2 pi N)  (core.py:193-196)
        p.cfo[frame] = -atan2(Pi, Pr) * p.fs / (2.0 * 3.14159265358979323846 * (double)p.n_fft);
        if (p.best_d) p.best_d[frame] = bd;
        if (p.P) p.P[frame] = make_double2(Pr, Pi);
    }
}

}  // namespace ofs

using namespace ofs;

OFS_API int ofs_channel_apply(const void *tx, int32_t dtype, int64_t n_rows, int64_t n_tx, const void *taps_c128, int32_t n_taps,
                              int64_t n_streams, const int32_t *row_of_stream, const void *unit_noise, int64_t noise_stride,
                              const double *snr_db, const double *cfo_hz, double fs, const double *full_scale, int32_t bits,
                              void *out, int16_t *out_iq, int64_t out_stride, void *faded_ws, double *power_ws, void *stream_)
{
    OFS_TRACE();
    OFS_REQUIRE(tx && (out || out_iq) && faded_ws && power_ws, "ofs_channel_apply: null argument");
    OFS_REQUIRE(dtype == OFS_C64 || dtype == OFS_C128, "ofs_channel_apply: samples must be complex64 or complex128");
    OFS_REQUIRE(n_rows >= 0 && n_tx >= 1 && n_streams >= 0, "ofs_channel_apply: bad geometry");
    OFS_REQUIRE(!taps_c128 || (n_taps >= 1 && n_taps <= 2048), "ofs_channel_apply: 1..2048 taps");
    OFS_REQUIRE(!unit_noise || snr_db, "ofs_channel_apply: noise needs snr_db");
    OFS_REQUIRE(!out_iq || full_scale, "ofs_channel_apply: int16 output needs full_scale");
    OFS_REQUIRE(!full_scale || (bits >= 2 && bits <= 16), "ofs_channel_apply: 2..16 ADC bits");
    OFS_REQUIRE(row_of_stream || n_streams == n_rows, "ofs_channel_apply: identity mapping needs n_streams == n_rows");
    const int64_t n_out = taps_c128 ? n_tx + n_taps - 1 : n_tx;
    OFS_REQUIRE(out_stride >= n_out && (!unit_noise || noise_stride >= n_out), "ofs_channel_apply: stride < output length");
    if (n_rows == 0 || n_streams == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t esz = dtype == OFS_C64 ? 8 : 16;
    if (taps_c128) {
        keep_pool_cached();
        double2 *ref = nullptr;
        OFS_CUDA(cudaMallocAsync((void **)&ref, (size_t)n_taps * sizeof(double2), stream));
        chan_ref_kernel<<<(n_taps + 255) / 256, 256, 0, stream>>>((const double2 *)taps_c128, n_taps, ref);
        if (int rc = check_launch("chan_ref_kernel")) return rc;
        const int rc = ofs_zc_matched_filter(tx, dtype, n_rows, 1, n_tx, ref, n_taps, 2, dtype == OFS_C128, faded_ws, nullptr, n_out, stream_);
        OFS_CUDA(cudaFreeAsync(ref, stream));              // stream-ordered: also on the error path
        if (rc) return rc;
    } else {
        OFS_CUDA(cudaMemcpyAsync(faded_ws, tx, (size_t)n_rows * n_tx * esz, cudaMemcpyDeviceToDevice, stream));
    }
    if (unit_noise) {
        if (dtype == OFS_C64) chan_power_kernel<float><<<(unsigned)n_rows, 256, 0, stream>>>(faded_ws, n_out, n_out, power_ws);
        else chan_power_kernel<double><<<(unsigned)n_rows, 256, 0, stream>>>(faded_ws, n_out, n_out, power_ws);
        if (int rc = check_launch("chan_power_kernel")) return rc;
    }
    ImpairParams p{};
    p.faded = faded_ws; p.noise = unit_noise; p.row_of_stream = row_of_stream; p.power = power_ws; p.snr_db = snr_db; p.cfo_hz = cfo_hz;
    p.full_scale = full_scale; p.out = out; p.out_iq = reinterpret_cast<short2 *>(out_iq); p.n = n_out; p.faded_stride = n_out;
    p.noise_stride = noise_stride; p.out_stride = out_stride; p.fs = fs; p.bits = bits;
    OFS_REQUIRE(n_streams < 65536, "ofs_channel_apply: at most 65535 streams per call");
    const dim3 grid((unsigned)((n_out + 1023) / 1024), (unsigned)n_streams);
    if (dtype == OFS_C64) chan_impair_kernel<float><<<grid, 256, 0, stream>>>(p);
    else chan_impair_kernel<double><<<grid, 256, 0, stream>>>(p);
    return check_launch("chan_impair_kernel");
}

OFS_API int ofs_cp_cfo(const void *x, int32_t in_dtype, int64_t n_frames, int32_t n_branches, int64_t n, int64_t x_frame_stride,
                       int64_t x_branch_stride, const int64_t *starts, int32_t n_fft, int32_t cp_len, int32_t span, int32_t win_len,
                       int32_t mode, double fs, double *cfo_hz, int64_t *best_d, void *P_c128, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(x && starts && cfo_hz, "ofs_cp_cfo: null argument");
    OFS_REQUIRE(in_dtype >= OFS_C64 && in_dtype <= OFS_IQ16, "ofs_cp_cfo: unknown dtype");
    OFS_REQUIRE(n_frames >= 0 && n_branches >= 1 && n >= 1 && n_fft >= 1 && cp_len >= 1, "ofs_cp_cfo: bad geometry");
    OFS_REQUIRE(mode >= 0 && mode <= 2, "ofs_cp_cfo: mode 0 (plain), 1 (robust), 2 (peak)");
    OFS_REQUIRE(span >= 0 && (mode != 1 || win_len >= 1), "ofs_cp_cfo: bad span / win_len");
    if (n_frames == 0) return OFS_OK;
    CfoParams p{};
    p.x = x; p.starts = starts; p.n_frames = n_frames; p.n = n; p.xfs = x_frame_stride; p.xbs = x_branch_stride; p.nb = n_branches;
    p.n_fft = n_fft; p.cp_len = cp_len; p.span = span; p.win = win_len; p.mode = mode; p.fs = fs; p.cfo = cfo_hz; p.best_d = best_d;
    p.P = (double2 *)P_c128;
    const int W = mode == 1 ? win_len : cp_len;
    int64_t nq = (mode == 0 ? 1 : 2LL * span) + W;
    if (nq > CFO_MAXQ) {
        set_error("ofs_cp_cfo: 2*span + window = %lld lag products exceed the %d this build holds in shared memory", (long long)nq, CFO_MAXQ);
        return OFS_EUNSUPPORTED;
    }
    const size_t smem = (size_t)(nq + 2) * sizeof(double2);
#define OFS_CFO_LAUNCH(DT)                                                                                     \
    do {                                                                                                       \
        OFS_CUDA(cudaFuncSetAttribute(cp_cfo_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cp_cfo_kernel<DT><<<(unsigned)n_frames, CFO_NT, smem, (cudaStream_t)stream>>>(p);                      \
    } while (0)
    if (in_dtype == OFS_C64) OFS_CFO_LAUNCH(OFS_C64);
    else if (in_dtype == OFS_C128) OFS_CFO_LAUNCH(OFS_C128);
    else OFS_CFO_LAUNCH(OFS_IQ16);
#undef OFS_CFO_LAUNCH
    return check_launch("cp_cfo_kernel");
}

// -------------------------------------------------------------------------------------------------------------------
// 12-bit wire formats of the RTL side: one launch moves every word once, coalesced both ways.
//   OFS_WIRE_HEX24   uint32 word = {Re[11:0], Im[11:0]}, Re in the upper 12 bits         (docs/preamble_test_vector.hex)
//   OFS_WIRE_AXIS48  uint64 word = {ch1_q, ch1_i, ch0_q, ch0_i} x 12 bits, ch0_i lowest   (ref/test_minn_preamble_detector.py:41-47,
//                                                                                          minn_preamble_detector.sv:23,98-101)
// int16 IQ layout on the other side: [n_channels][n][2], the layout every OFS_IQ16 kernel ingests.
__device__ __forceinline__ int sext12(unsigned v) { return (int)(v << 20) >> 20; }

__global__ void __launch_bounds__(256) wire_pack_kernel(const short2 *__restrict__ iq, int64_t n, int64_t ch_stride, int fmt, void *__restrict__ words)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const short2 a = iq[i];
    if (fmt == OFS_WIRE_HEX24) {
        reinterpret_cast<uint32_t *>(words)[i] = (((unsigned)a.x & 0xfffu) << 12) | ((unsigned)a.y & 0xfffu);
    } else {
        const short2 b = iq[ch_stride + i];
        const uint64_t lo = ((unsigned)a.x & 0xfffu) | (((unsigned)a.y & 0xfffu) << 12);
        const uint64_t hi = ((unsigned)b.x & 0xfffu) | (((unsigned)b.y & 0xfffu) << 12);
        reinterpret_cast<uint64_t *>(words)[i] = lo | (hi << 24);
    }
}

__global__ void __launch_bounds__(256) wire_unpack_kernel(const void *__restrict__ words, int64_t n, int64_t ch_stride, int fmt, short2 *__restrict__ iq)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (fmt == OFS_WIRE_HEX24) {
        const uint32_t w = reinterpret_cast<const uint32_t *>(words)[i];
        iq[i] = make_short2((short)sext12(w >> 12), (short)sext12(w));
    } else {
        const uint64_t w = reinterpret_cast<const uint64_t *>(words)[i];
        const unsigned lo = (unsigned)(w & 0xffffffu), hi = (unsigned)((w >> 24) & 0xffffffu);
        iq[i] = make_short2((short)sext12(lo), (short)sext12(lo >> 12));
        iq[ch_stride + i] = make_short2((short)sext12(hi), (short)sext12(hi >> 12));
    }
}

OFS_API int ofs_wire_pack(const int16_t *iq, int64_t n, int32_t format, void *words, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(iq && words, "ofs_wire_pack: null argument");
    OFS_REQUIRE(format == OFS_WIRE_HEX24 || format == OFS_WIRE_AXIS48, "ofs_wire_pack: unknown format");
    OFS_REQUIRE(n >= 0, "ofs_wire_pack: bad length");
    if (n == 0) return OFS_OK;
    wire_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const short2 *>(iq), n, n, format, words);
    return check_launch("wire_pack_kernel");
}

OFS_API int ofs_wire_unpack(const void *words, int64_t n, int32_t format, int16_t *iq, void *stream)
{
    OFS_TRACE();
    OFS_REQUIRE(iq && words, "ofs_wire_unpack: null argument");
    OFS_REQUIRE(format == OFS_WIRE_HEX24 || format == OFS_WIRE_AXIS48, "ofs_wire_unpack: unknown format");
    OFS_REQUIRE(n >= 0, "ofs_wire_unpack: bad length");
    if (n == 0) return OFS_OK;
    wire_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(words, n, n, format, reinterpret_cast<short2 *>(iq));
    return check_launch("wire_unpack_kernel");
}

// K6: multi-root Zadoff-Chu correlator bank on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
// Generalises zc_freq.compute_frequency_metric (zc_freq.py:62-99, one root: np.vdot(template, bins)) to a bank of
// up to 128 roots evaluated at every candidate offset:
//     Y[r, o] = sum_j conj(T[r, j]) * bins[j, o]          (62 used DFT bins -> K = 2 x 64 real columns)
//     metric[r, o] = |Y|^2 / max(E_r * E(o), 1e-12),  E(o) = sum_j |bins[j, o]|^2
// Only the maximum over offsets (value + offset) of every root is kept per capture.
//
// Two kernels per chunk of captures:
//  1. zc_bins_kernel  (SIMT, float64 modulated prefix sums = K5's sliding DFT) writes bins^T[k][o] (k = 0..63 real
//     parts, 64..127 imaginary parts; fp32) -- offsets contiguous, i.e. the MN-major B operand -- and E[o].
//     zc_bins_transpose_kernel turns it into bins[o][k] (K-major, what the MMA wants).
//  2. zc_bank_umma_kernel: one CTA per capture.  A = the templates as two K-major 128x128 fp32 matrices (rows = roots;
//     A_re gives Re Y, A_im gives Im Y), resident in shared memory; B tiles (128 offsets x 128 k, 64 KB) arrive by four
//     tiled TMA copies with 128B swizzle; one elected thread issues 2 x 16 tcgen05.mma.kind::tf32 (M=128, N=128, K=8) into
//     two TMEM accumulators (256 columns); tcgen05.commit signals an mbarrier; the four warps read their TMEM lane
//     quadrant with tcgen05.ld (lane = root), form |Y|^2 / (E_r E(o)) and keep a running maximum per root in registers.
// Accuracy: TF32 inputs (10-bit mantissa), FP32 accumulation in TMEM -> |d metric| <= 5e-3 * max(metric) (tested).
#include "common.cuh"
#include <cuda.h>
#include <string.h>
#include <stdlib.h>

namespace ofs {

constexpr int BK_ROOTS = 128;     // M: roots (lanes of TMEM)
constexpr int BK_K = 128;         // K: 64 real + 64 imaginary bin columns
constexpr int BK_N = 128;         // N: offsets per tile
constexpr int QB = 256;           // threads of the bins kernel

// ------------------------------------------------------------------------------------------------ bins kernel
template <int DT>
__global__ void __launch_bounds__(QB) zc_bins_kernel(const void *x, int64_t n, int N, int cp, const int *bins, int nbins, int TO,
                                                     int64_t n_off, int64_t n_off_pad, float *binsT, float *Eo, int tiles_per_cap)
{
    using In = typename InT<DT>::type;
    extern __shared__ __align__(16) unsigned char qsm[];
    const int span = TO + N - 1;
    double2 *xs = reinterpret_cast<double2 *>(qsm);               // span
    double2 *S = xs + span;                                        // span + 1
    double2 *tw = S + span + 1;                                    // N
    __shared__ double wtot[2][QB / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t cap = blockIdx.x / tiles_per_cap;
    const int tile = blockIdx.x % tiles_per_cap;
    const int64_t o0 = (int64_t)tile * TO;
    const int64_t jb = o0 + cp;
    for (int m = tid; m < N; m += QB) {
        double s, c;
        sincospi(-2.0 * (double)m / (double)N, &s, &c);
        tw[m] = make_double2(c, s);
    }
    const In *xb = reinterpret_cast<const In *>(x) + cap * n;
    for (int m = tid; m < span; m += QB) {
        const int64_t j = jb + m;
        double2 v = make_double2(0.0, 0.0);
        if (j < n) { const In s = xb[j]; v = make_double2((double)s.x, (double)s.y); }
        xs[m] = v;
    }
    if (tid == 0) S[0] = make_double2(0.0, 0.0);
    const int ipt = ((span + QB - 1) / QB) | 1;
    const int s0 = tid * ipt, s1 = min(s0 + ipt, span);
    constexpr int OPT = 8;
    double en[OPT];
#pragma unroll
    for (int q = 0; q < OPT; ++q) en[q] = 0.0;
    float *bt = binsT + cap * (int64_t)BK_K * n_off_pad;
    __syncthreads();
    for (int jbin = 0; jbin < nbins; ++jbin) {
        const int k = bins[jbin];
        double rr = 0.0, ri = 0.0;
        int ph = (int)((jb + s0) % N);
        for (int m = s0; m < s1; ++m) {
            const double2 w = tw[(int)(((long long)k * ph) % N)];
            const double2 v = xs[m];
            rr += v.x * w.x - v.y * w.y;
            ri += v.x * w.y + v.y * w.x;
            S[m + 1] = make_double2(rr, ri);
            ph = ph + 1 == N ? 0 : ph + 1;
        }
        double tr = rr, ti = ri;
        for (int o = 1; o < 32; o <<= 1) {
            const double yr = shfl_up_f64(tr, o), yi = shfl_up_f64(ti, o);
            if (lane >= o) { tr += yr; ti += yi; }
        }
        if (lane == 31) { wtot[0][warp] = tr; wtot[1][warp] = ti; }
        __syncthreads();
        double offr = tr - rr, offi = ti - ri;
        for (int w = 0; w < warp; ++w) { offr += wtot[0][w]; offi += wtot[1][w]; }
        for (int m = s0; m < s1; ++m) { S[m + 1].x += offr; S[m + 1].y += offi; }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < OPT; ++q) {
            const int i = tid + q * QB;
            if (i < TO && o0 + i < n_off_pad) {
                float br = 0.f, bi = 0.f;
                if (o0 + i < n_off) {
                    const double2 hi = S[i + N], lo = S[i];
                    const double dr = hi.x - lo.x, di = hi.y - lo.y;
                    const double2 w = tw[(int)(((long long)k * ((jb + i) % N)) % N)];
                    const double b_r = dr * w.x + di * w.y, b_i = di * w.x - dr * w.y;   // conj(w) * d
                    en[q] += b_r * b_r + b_i * b_i;
                    br = (float)b_r; bi = (float)b_i;
                }
                bt[(int64_t)jbin * n_off_pad + o0 + i] = br;                 // coalesced along the offsets
                bt[(int64_t)(64 + jbin) * n_off_pad + o0 + i] = bi;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < OPT; ++q) {
        const int i = tid + q * QB;
        if (i < TO && o0 + i < n_off_pad) Eo[cap * n_off_pad + o0 + i] = (o0 + i < n_off) ? (float)en[q] : 0.f;
    }
}

// bins^T [cap][128 k][n_off_pad] -> bins [cap][n_off_pad][128 k]: the K-major B operand of the MMA (a tf32 MMA with an
// MN-major B operand produced all-zero accumulators on this stack, so K-major it is).  32x32 smem tiles, coalesced both ways.
__global__ void zc_bins_transpose_kernel(const float *binsT, float *binsK, int64_t n_off_pad)
{
    __shared__ float t[32][33];
    const int64_t cap = blockIdx.z;
    const int64_t o0 = (int64_t)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    const float *src = binsT + cap * (int64_t)BK_K * n_off_pad;
    float *dst = binsK + cap * n_off_pad * (int64_t)BK_K;
    for (int r = threadIdx.y; r < 32; r += 8) t[r][threadIdx.x] = src[(int64_t)(k0 + r) * n_off_pad + o0 + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) dst[(o0 + r) * (int64_t)BK_K + k0 + threadIdx.x] = t[threadIdx.x][r];
}

__global__ void zc_bank_unpack_kernel(const unsigned long long *packed, int64_t count, float *best_metric, int32_t *best_offset)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned long long k = packed[i];
    best_metric[i] = __uint_as_float((unsigned)(k >> 32));
    best_offset[i] = (int32_t)(0xffffffffu - (unsigned)(k & 0xffffffffull));
}

// templates -> A_re (rows 0..127) and A_im (rows 128..255), K-major [256][128] fp32
//   Re Y[r] = sum_j Tre[r,j] Bre[j] + Tim[r,j] Bim[j]      Im Y[r] = sum_j Tre[r,j] Bim[j] - Tim[r,j] Bre[j]
__global__ void zc_bank_templates_kernel(const float2 *templ, int nbins, int n_roots, float *A, float *Er)
{
    const int r = blockIdx.x, k = threadIdx.x;       // 128 x 128
    float tre = 0.f, tim = 0.f;
    const int j = k & 63;
    if (r < n_roots && j < nbins) { const float2 t = templ[r * nbins + j]; tre = t.x; tim = t.y; }
    const bool imag_col = k >= 64;
    A[r * BK_K + k] = imag_col ? tim : tre;
    A[(BK_ROOTS + r) * BK_K + k] = imag_col ? tre : -tim;
    if (k == 0) {
        float e = 0.f;
        if (r < n_roots) for (int q = 0; q < nbins; ++q) { const float2 t = templ[r * nbins + q]; e += t.x * t.x + t.y * t.y; }
        Er[r] = e;
    }
}

// ------------------------------------------------------------------------------------------------ tcgen05 helpers
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;"      // same asm statement: the registers are valid when it returns
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
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
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}

// ------------------------------------------------------------------------------------------------ bank kernel
__global__ void __launch_bounds__(128, 1)
zc_bank_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const float *Eo,
                    const float *Er, int64_t n_off, int64_t n_off_pad, int n_roots, int n_split,
                    unsigned long long *best_packed, float *ydbg, int dbg_mode)
{
    extern __shared__ __align__(1024) unsigned char bsm[];
    unsigned char *sA = bsm;                              // 2 x 64 KB: A_re, A_im (each 4 K-chunks of 128 rows x 128 B)
    unsigned char *sB = bsm + 2 * 65536;                  // 64 KB: 4 N-chunks (32 offsets) of 128 k-rows x 128 B
    uint64_t *bars = reinterpret_cast<uint64_t *>(bsm + 3 * 65536);      // [0] A loaded, [1] B tile loaded, [2] MMA done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    float *sE = reinterpret_cast<float *>(bars + 8);      // E(o) of the tile (128 floats)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cap = blockIdx.x / n_split;                 // capture inside the chunk
    const int split = blockIdx.x % n_split;               // this CTA takes tiles split, split + n_split, ...
    constexpr int cap0 = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
        mbar_fence_init();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // templates: 8 boxes of {32 k, 128 rows}
    if (tid == 0) {
        mbar_expect_tx(&bars[0], 2 * 65536);
        for (int h = 0; h < 2; ++h)
            for (int kc = 0; kc < 4; ++kc) tma_load_2d_b(sA + h * 65536 + kc * 16384, &mapA, kc * 32, h * 128, &bars[0]);
    }
    mbar_wait_bounded(&bars[0], 0);

    // instruction descriptor: D=F32, A=B=TF32, both K-major, N=128, M=128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BK_N >> 3) << 17) | ((uint32_t)(BK_ROOTS >> 4) << 24);
    const float er = Er[tid];                             // this thread's root (TMEM lane = tid)
    float best = -1.f;
    int best_o = 0;
    uint32_t pb = 0, pm = 0;
    const int n_tiles = (int)((n_off + BK_N - 1) / BK_N);
    const int rowB = (int)((cap0 + cap) * n_off_pad);     // first tensor row (= offset) of this capture in bins[o][k]
    for (int tile = split; tile < n_tiles; tile += n_split) {
        const int64_t o0 = (int64_t)tile * BK_N;
        if (tid == 0) {
            mbar_expect_tx(&bars[1], 65536);
            for (int kc = 0; kc < 4; ++kc) tma_load_2d_b(sB + kc * 16384, &mapB, kc * 32, rowB + (int)o0, &bars[1]);
        }
        sE[tid] = Eo[(int64_t)(cap0 + cap) * n_off_pad + o0 + tid];
        mbar_wait_bounded(&bars[1], pb); pb ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg_mode == 1) {                               // st/ld self-test: lane*1000 + column
            const uint32_t tq0 = tmem + ((uint32_t)(warp * 32) << 16);
            for (int c = 0; c < 256; ++c) {
                const uint32_t v = __float_as_uint((float)(tid * 1000 + c));
                asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tq0 + c), "r"(v) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (tid == 0 && dbg_mode == 2) {                   // both operands K-major: D = A_re * A_im^T
            const uint32_t idk = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BK_N >> 3) << 17) | ((uint32_t)(BK_ROOTS >> 4) << 24);
            for (int step = 0; step < 16; ++step) {
                const uint64_t ad = umma_desc(smem_u32(sA + (step >> 2) * 16384 + (step & 3) * 32), 16, 1024);
                const uint64_t bd = umma_desc(smem_u32(sA + 65536 + (step >> 2) * 16384 + (step & 3) * 32), 16, 1024);
                umma_tf32(tmem, ad, bd, idk, step > 0 ? 1u : 0u);
                umma_tf32(tmem + BK_N, ad, bd, idk, step > 0 ? 1u : 0u);
            }
            umma_commit(&bars[2]);
        }
        if (tid == 0 && dbg_mode >= 3) {                   // MN-major B test on known data: D = A_re x A_im (plain matrix product)
            for (int step = 0; step < 16; ++step) {
                const uint64_t ad = umma_desc(smem_u32(sA + (step >> 2) * 16384 + (step & 3) * 32), 16, 1024);
                const uint64_t bd = dbg_mode == 3 ? umma_desc(smem_u32(sA + 65536 + step * 1024), 16384, 1024)
                                                  : umma_desc(smem_u32(sA + 65536 + step * 1024), 1024, 16384);
                umma_tf32(tmem, ad, bd, idesc, step > 0 ? 1u : 0u);
                umma_tf32(tmem + BK_N, ad, bd, idesc, step > 0 ? 1u : 0u);
            }
            umma_commit(&bars[2]);
        }
        if (tid == 0 && dbg_mode == 1) umma_commit(&bars[2]);
        if (tid == 0 && dbg_mode == 0) {
            // K loop: 16 steps of 8 tf32.  Both operands K-major SW128: K-chunk kc = step/4 (a 128-row x 128-byte box), 32 bytes
            // per step inside the swizzled 128-byte row; 8-row atoms are 1024 B apart (SBO).
            for (int h = 0; h < 2; ++h) {
                for (int step = 0; step < 16; ++step) {
                    const uint64_t ad = umma_desc(smem_u32(sA + h * 65536 + (step >> 2) * 16384 + (step & 3) * 32), 16, 1024);
                    const uint64_t bd = umma_desc(smem_u32(sB + (step >> 2) * 16384 + (step & 3) * 32), 16, 1024);
                    umma_tf32(tmem + h * BK_N, ad, bd, idesc, step > 0 ? 1u : 0u);
                }
            }
            umma_commit(&bars[2]);
        }
        __syncthreads();                                   // sE visible
        mbar_wait_bounded(&bars[2], pm); pm ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: this warp's TMEM lane quadrant, 32 offsets at a time; lane = root
        const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < BK_N; c += 32) {
            uint32_t re[32], im[32];
            tmem_ld32(tq + c, re);
            tmem_ld32(tq + BK_N + c, im);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const float yr = __uint_as_float(re[q]), yi = __uint_as_float(im[q]);
                const float den = fmaxf(er * sE[c + q], 1e-12f);
                const float m = (yr * yr + yi * yi) / den;
                const int64_t o = o0 + c + q;
                if (ydbg && cap == 0) { ydbg[(int64_t)tid * n_off_pad + o] = yr; ydbg[(int64_t)(128 + tid) * n_off_pad + o] = yi; }
                if (o < n_off && m > best) { best = m; best_o = (int)o; }
            }
        }
        if (ydbg && cap == 0 && tile == 0) {               // debugging aid: raw operand words as they sit in smem
            ydbg[(int64_t)250 * n_off_pad + tid] = reinterpret_cast<const float *>(sA)[tid];
            ydbg[(int64_t)251 * n_off_pad + tid] = reinterpret_cast<const float *>(sB)[tid];
            ydbg[(int64_t)252 * n_off_pad + tid] = reinterpret_cast<const float *>(sA + 65536)[tid];
            ydbg[(int64_t)253 * n_off_pad + tid] = __uint_as_float(tmem);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                   // TMEM and sB / sE may be overwritten by the next tile
    }
    // combine the splits: max metric, earliest offset on ties (metric >= 0: float bits order like unsigned)
    if (tid < n_roots && best >= 0.f) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xffffffffu - (unsigned)best_o);
        atomicMax(best_packed + (int64_t)cap * n_roots + tid, key);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
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
static bool make_map_f32(CUtensorMap *map, const void *base, uint64_t inner, uint64_t rows, uint32_t box_rows)
{
    EncodeTiledFn2 fn = encode_fn2();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {inner, rows};
    const cuuint64_t gstride[1] = {inner * 4};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ofs

using namespace ofs;

OFS_API int ofs_zc_bank(const void *x_c64, int64_t n_frames, int64_t n, int32_t n_fft, int32_t cp, const int32_t *bins,
                        const void *templ_c64, int32_t nbins, int32_t n_roots, float *best_metric, int32_t *best_offset,
                        void *stream_)
{
    OFS_REQUIRE(x_c64 && bins && templ_c64 && best_metric && best_offset, "ofs_zc_bank: null argument");
    OFS_REQUIRE(n_fft >= 2 && n_fft <= 2048 && cp >= 0, "ofs_zc_bank: n_fft must be 2..2048");
    OFS_REQUIRE(nbins >= 1 && nbins <= 64 && n_roots >= 1 && n_roots <= BK_ROOTS, "ofs_zc_bank: nbins <= 64, n_roots <= 128");
    const int64_t n_off = n - ((int64_t)n_fft + cp) + 1;
    OFS_REQUIRE(n_off > 0, "Received stream is shorter than a single OFDM symbol.");
    if (n_frames == 0) return OFS_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t n_off_pad = (n_off + BK_N - 1) / BK_N * BK_N;
    int64_t chunk = (int64_t)(1ull << 29) / (BK_K * n_off_pad * 4);          // <= 0.5 GB of bins (x2 layouts) per chunk
    if (chunk < 1) chunk = 1;
    if (chunk > n_frames) chunk = n_frames;
    float *binsT = nullptr, *binsK = nullptr, *Eo = nullptr, *A = nullptr, *Er = nullptr;
    unsigned long long *packed = nullptr;
    keep_pool_cached();
    OFS_CUDA(cudaMallocAsync((void **)&packed, (size_t)chunk * n_roots * 8, stream));
    OFS_CUDA(cudaMallocAsync((void **)&binsT, (size_t)chunk * BK_K * n_off_pad * 4, stream));
    OFS_CUDA(cudaMallocAsync((void **)&binsK, (size_t)chunk * BK_K * n_off_pad * 4, stream));
    OFS_CUDA(cudaMallocAsync((void **)&Eo, (size_t)chunk * n_off_pad * 4, stream));
    OFS_CUDA(cudaMallocAsync((void **)&A, (size_t)2 * BK_ROOTS * BK_K * 4, stream));
    OFS_CUDA(cudaMallocAsync((void **)&Er, BK_ROOTS * 4, stream));
    OFS_CUDA(cudaMemsetAsync(binsT, 0, (size_t)chunk * BK_K * n_off_pad * 4, stream));     // pad rows 62,63,126,127 stay zero
    zc_bank_templates_kernel<<<BK_ROOTS, BK_K, 0, stream>>>((const float2 *)templ_c64, nbins, n_roots, A, Er);
    if (int rc = check_launch("zc_bank_templates_kernel")) return rc;

    CUtensorMap mapA, mapB;
    OFS_REQUIRE(make_map_f32(&mapA, A, BK_K, 2 * BK_ROOTS, 128), "ofs_zc_bank: cuTensorMapEncodeTiled(A) failed");
    OFS_REQUIRE(make_map_f32(&mapB, binsK, BK_K, (uint64_t)chunk * n_off_pad, 128), "ofs_zc_bank: cuTensorMapEncodeTiled(B) failed");

    int TO = 2048;
    if (n_off_pad < TO) TO = (int)((n_off_pad + 255) / 256 * 256);
    const int tiles = (int)((n_off_pad + TO - 1) / TO);
    const int span = TO + n_fft - 1;
    const size_t smem_b = (size_t)(span + span + 1 + n_fft) * sizeof(double2);
    OFS_CUDA(cudaFuncSetAttribute(zc_bins_kernel<OFS_C64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    const size_t smem_u = 3 * 65536 + 64 + 128 * sizeof(float) + 1024;
    OFS_CUDA(cudaFuncSetAttribute(zc_bank_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));
    for (int64_t c0 = 0; c0 < n_frames; c0 += chunk) {
        const int64_t nc = (c0 + chunk <= n_frames) ? chunk : n_frames - c0;
        zc_bins_kernel<OFS_C64><<<(unsigned)(nc * tiles), QB, smem_b, stream>>>(
            reinterpret_cast<const float2 *>(x_c64) + c0 * n, n, n_fft, cp, bins, nbins, TO, n_off, n_off_pad, binsT, Eo, tiles);
        if (int rc = check_launch("zc_bins_kernel")) return rc;
        zc_bins_transpose_kernel<<<dim3((unsigned)(n_off_pad / 32), BK_K / 32, (unsigned)nc), dim3(32, 8), 0, stream>>>(binsT, binsK, n_off_pad);
        if (int rc = check_launch("zc_bins_transpose_kernel")) return rc;
        float *ydbg = nullptr;
        const char *dbg = getenv("OFS_BANK_DEBUG");
        const char *dbgm = getenv("OFS_BANK_DEBUG_MODE");
        const int dbg_mode = dbgm ? atoi(dbgm) : 0;
        if (dbg && c0 == 0) OFS_CUDA(cudaMalloc((void **)&ydbg, (size_t)256 * n_off_pad * 4));
        const int n_tiles_h = (int)((n_off + BK_N - 1) / BK_N);
        int n_split = (int)((2 * sm_count() + nc - 1) / nc);
        if (n_split > n_tiles_h) n_split = n_tiles_h;
        if (n_split < 1) n_split = 1;
        OFS_CUDA(cudaMemsetAsync(packed, 0, (size_t)nc * n_roots * 8, stream));
        zc_bank_umma_kernel<<<(unsigned)(nc * n_split), 128, smem_u, stream>>>(mapA, mapB, Eo, Er, n_off, n_off_pad, n_roots, n_split,
                                                                           packed, ydbg, dbg_mode);
        if (int rc = check_launch("zc_bank_umma_kernel")) return rc;
        zc_bank_unpack_kernel<<<(unsigned)((nc * n_roots + 255) / 256), 256, 0, stream>>>(packed, nc * n_roots, best_metric + c0 * n_roots,
                                                                                      best_offset + c0 * n_roots);
        if (int rc = check_launch("zc_bank_unpack_kernel")) return rc;
        if (ydbg) {   // debugging aid: dump the first capture's operands and accumulators
            OFS_CUDA(cudaStreamSynchronize(stream));
            auto dump = [&](const char *name, const void *dptr, size_t bytes) {
                void *h = malloc(bytes);
                cudaMemcpy(h, dptr, bytes, cudaMemcpyDeviceToHost);
                char path[512];
                snprintf(path, sizeof(path), "%s_%s.bin", dbg, name);
                FILE *f = fopen(path, "wb");
                if (f) { fwrite(h, 1, bytes, f); fclose(f); }
                free(h);
            };
            dump("binsT", binsT, (size_t)BK_K * n_off_pad * 4);
            dump("E", Eo, (size_t)n_off_pad * 4);
            dump("A", A, (size_t)2 * BK_ROOTS * BK_K * 4);
            dump("Y", ydbg, (size_t)256 * n_off_pad * 4);
            cudaFree(ydbg);
        }
    }
    OFS_CUDA(cudaFreeAsync(packed, stream));
    OFS_CUDA(cudaFreeAsync(binsT, stream));
    OFS_CUDA(cudaFreeAsync(binsK, stream));
    OFS_CUDA(cudaFreeAsync(Eo, stream));
    OFS_CUDA(cudaFreeAsync(A, stream));
    OFS_CUDA(cudaFreeAsync(Er, stream));
    return OFS_OK;
}

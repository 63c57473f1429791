// Shared helpers for libofdmsync (sm_100a).  Product code: nothing here touches oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <nvtx3/nvToolsExt.h>
#include "../../include/ofdmsync.h"

#define OFS_API extern "C" __attribute__((visibility("default")))

namespace ofs {

// ---- error reporting (thread-local string, int codes across the ABI) -------------------------
void set_error(const char *fmt, ...);
int64_t &launch_counter();
inline void count_launch(int n = 1) { launch_counter() += n; }

#define OFS_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            ofs::set_error(__VA_ARGS__);            \
            return OFS_EINVAL;                      \
        }                                           \
    } while (0)

#define OFS_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ofs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                 \
        }                                                                                   \
    } while (0)

inline int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    count_launch();
    return OFS_OK;
}

// NVTX range around every compute entry point of the C ABI (tracing hook).  Header-only NVTX3: a no-op function
// pointer check unless a tool (nsys, ncu --nvtx) has injected itself.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define OFS_TRACE() ofs::NvtxRange _ofs_trace_range(__func__)

// Stream-ordered scratch allocation that is returned on every exit path (an early `return rc` after a failed launch used to
// leak the buffers allocated before it).
struct AsyncBuf {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    AsyncBuf() = default;
    AsyncBuf(const AsyncBuf &) = delete;
    AsyncBuf &operator=(const AsyncBuf &) = delete;
    cudaError_t alloc(size_t bytes, cudaStream_t s) { st = s; return cudaMallocAsync(&p, bytes, s); }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
    ~AsyncBuf() { if (p) cudaFreeAsync(p, st); }
};

int sm_count();  // multiprocessor count of the CURRENT device (cached per device)
int current_device();  // cudaGetDevice, clamped to [0, OFS_MAX_DEVICES)
constexpr int OFS_MAX_DEVICES = 64;
// One-time per-DEVICE set-up (function attributes, occupancy queries are per device: a process may drive several GPUs through
// ofs_ctx_create(ctx, device)).  `if (!once.done()) { ...idempotent set-up...; once.mark(); }` -- two threads racing on the same
// device both run the set-up, which is harmless; nobody launches before its own set-up has finished.
struct PerDeviceOnce {
    unsigned long long bits = 0;
    bool done() const { return (__atomic_load_n(&bits, __ATOMIC_ACQUIRE) >> current_device()) & 1ull; }
    void mark() { __atomic_fetch_or(&bits, 1ull << current_device(), __ATOMIC_RELEASE); }
};
void keep_pool_cached();  // stream-ordered workspace (cudaMallocAsync): keep freed blocks in the pool between calls

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// mbarrier + 1-D bulk TMA (cp.async.bulk) -- SASS: SYNCS.* / UBLKCP
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ double shfl_up_f64(double v, int delta)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, delta);
    hi = __shfl_up_sync(0xffffffffu, hi, delta);
    return __hiloint2double(hi, lo);
}
// one Kogge-Stone step of a warp inclusive scan: v += (value of lane - delta), predicated by the
// shuffle's own "source lane in range" output (no ISETP / FSEL)
__device__ __forceinline__ void scan_step_f64(double &v, int delta)
{
    asm volatile(
        "{\n"
        ".reg .b32 lo, hi, ylo, yhi;\n"
        ".reg .pred p;\n"
        ".reg .f64 y;\n"
        "mov.b64 {lo, hi}, %0;\n"
        "shfl.sync.up.b32 ylo|p, lo, %1, 0x0, 0xffffffff;\n"
        "shfl.sync.up.b32 yhi, hi, %1, 0x0, 0xffffffff;\n"
        "mov.b64 y, {ylo, yhi};\n"
        "@p add.f64 %0, %0, y;\n"
        "}\n"
        : "+d"(v)
        : "r"(delta));
}
__device__ __forceinline__ double shfl_f64(double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ long long shfl_xor_i64(long long v, int m)
{
    int lo = (int)(v & 0xffffffffll), hi = (int)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return ((long long)hi << 32) | (unsigned int)lo;
}

// sample loads as double2 from any input format
template <int DT>
struct InT;
template <>
struct InT<OFS_C64> {
    using type = float2;
    static constexpr int bytes = 8;
};
template <>
struct InT<OFS_C128> {
    using type = double2;
    static constexpr int bytes = 16;
};
template <>
struct InT<OFS_IQ16> {
    using type = short2;
    static constexpr int bytes = 4;
};

__device__ __forceinline__ double2 load_sample_f64(const void *base, int dtype, int64_t idx)
{
    if (dtype == OFS_C64) {
        float2 v = __ldg(reinterpret_cast<const float2 *>(base) + idx);
        return make_double2((double)v.x, (double)v.y);
    } else if (dtype == OFS_C128) {
        return __ldg(reinterpret_cast<const double2 *>(base) + idx);
    } else {
        short2 v = __ldg(reinterpret_cast<const short2 *>(base) + idx);
        return make_double2((double)v.x, (double)v.y);
    }
}

// One packed int16 IQ word -> (float I, float Q), exactly, without I2F (quarter-rate pipe): bias both halves by 0x8000,
// drop each 16-bit field into the mantissa of 2^23 (PRMT) and subtract 2^23 + 32768 (FADD).  LOP3 + 2 PRMT + 2 FADD.
__device__ __forceinline__ float2 cvt_iq16(unsigned w)
{
    const unsigned b = w ^ 0x80008000u;
    const unsigned lo = __byte_perm(b, 0x4b000000u, 0x7610);      // bytes {b0, b1, 0x00, 0x4b}
    const unsigned hi = __byte_perm(b, 0x4b000000u, 0x7632);      // bytes {b2, b3, 0x00, 0x4b}
    return make_float2(__uint_as_float(lo) - 8421376.0f, __uint_as_float(hi) - 8421376.0f);
}

__device__ __forceinline__ size_t dtype_bytes_dev(int dt) { return dt == OFS_C64 ? 8 : (dt == OFS_C128 ? 16 : 4); }
inline size_t dtype_bytes(int dt) { return dt == OFS_C64 ? 8 : (dt == OFS_C128 ? 16 : 4); }

}  // namespace ofs

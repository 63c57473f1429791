// Host stand-in for csrc/common.cuh: lets g++ compile the header-only FFT / convolution stages (fft4096.cuh, conv8k.cuh) so a
// CPU test can run them thread by thread (tests/host/conv8k_host_test.cpp).  Test infrastructure only.
#pragma once
#include <cmath>
#include <cstdint>
#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline double2 make_double2(double x, double y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static struct { int x; } threadIdx;
inline void __syncthreads() {}
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
inline float2 __fadd2_rn(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }

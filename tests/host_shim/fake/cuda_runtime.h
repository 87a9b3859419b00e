// TEST INFRASTRUCTURE. A stand-in for <cuda_runtime.h> that lets g++ compile the product's DEVICE headers
// (pbrt_v2_spectral_b200/csrc/*.cuh) for the host, so that device functions can be compared with the oracle on a machine
// without a GPU (tests/test_device_code_on_host.py). Only what those headers use is provided.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r = { x, y }; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = { x, y, z, w }; return r; }
// round-to-nearest double arithmetic and conversion: what the host does anyway (no FMA contraction: -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline float __double2float_rn(double a) { return (float)a; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(uint32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t i; memcpy(&i, &f, 4); return i; }
using std::isinf;
using std::isnan;
using std::max;
using std::min;

// ---- what the KERNELS need on top: a "warp" of one lane, a grid of one thread ----------------------------------------
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = { x, y }; return r; }
struct HostDim3 { unsigned x, y, z; };
static const HostDim3 threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 }, blockDim = { 1, 1, 1 }, gridDim = { 1, 1, 1 };
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __activemask() { return 1u; }
static inline bool __any_sync(unsigned, int p) { return p != 0; }
static inline bool __all_sync(unsigned, int p) { return p != 0; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned, int = 32) { return v; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int, int = 32) { return v; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __brev(unsigned n) {
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
    n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
    n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
    return ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
}
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p += v; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p |= v; return o; }
using std::isfinite;
typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

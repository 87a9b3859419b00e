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

// sampler.cuh — K1's sample generator.
//
// The reference's LDSampler (src/samplers/lowdiscrepancy.cpp:59-71 -> LDPixelSample,
// src/core/montecarlo.cpp:192-244, src/core/montecarlo.h:262-315) gives every dimension of every
// pixel a randomly scrambled (0,2)-sequence visited in a random order, both drawn from ONE
// sequential MT19937 stream per image tile — inherently serial. Here the sequence is the same
// (VanDerCorput / Sobol2, bit-exact integer code) but scramble and visiting order come from
// counter-based hashes of (seed, pixel, dimension), so any sample of any pixel can be produced
// independently by any thread on any GPU: statistically equivalent, not stream-identical
// (SURVEY.md 7 "hard parts" 3). Draws the reference takes from the RNG itself (bounces >= 3,
// Russian roulette) are hashes of (seed, pixel, sample, counter).
#pragma once
#include "spt_device.cuh"

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// per-(seed, pixel) key, computed once per path vertex; every dimension derives from it with one mix
__device__ __forceinline__ uint32_t pixel_key(uint32_t seed, uint32_t pix) {
    return mix32(mix32(seed + 0x9e3779b9U) ^ (pix + 0x85ebca6bU));
}
__device__ __forceinline__ uint32_t dim_key(uint32_t pkey, uint32_t dim) { return mix32(pkey + dim * 0x9e3779b9U); }
// random permutation of [0,n), n a power of two (invertible mixing restricted to log2 n bits)
__device__ __forceinline__ uint32_t permute_pow2(uint32_t i, uint32_t n, uint32_t key) {
    uint32_t mask = n - 1;
    if (!mask) return 0;
    i ^= key; i *= 0xe170893dU; i ^= key >> 16;
    i ^= (i & mask) >> 4; i ^= key >> 8; i *= 0x0929eb3fU; i ^= key >> 23;
    i ^= (i & mask) >> 1; i *= 1 | key >> 27; i *= 0x6935fa69U;
    i ^= (i & mask) >> 11; i *= 0x74dcb303U; i ^= (i & mask) >> 2; i *= 0x9e501cc3U;
    i ^= (i & mask) >> 2; i *= 0xc860a3dfU; i &= mask; i ^= i >> 5;
    return (i + key) & mask;
}
__device__ __forceinline__ float van_der_corput(uint32_t n, uint32_t scramble) {   // montecarlo.h:270-279
    n = __brev(n);
    n ^= scramble;
    return stdminf(((n >> 8) & 0xffffff) / (float)(1 << 24), ONE_MINUS_EPS);
}
__device__ __forceinline__ float sobol2(uint32_t n, uint32_t scramble) {           // montecarlo.h:282-286
    for (uint32_t v = 1u << 31; n != 0; n >>= 1, v ^= v >> 1)
        if (n & 0x1) scramble ^= v;
    return stdminf(((scramble >> 8) & 0xffffff) / (float)(1 << 24), ONE_MINUS_EPS);
}
__device__ __forceinline__ float ld1(uint32_t pkey, uint32_t dim, uint32_t s, uint32_t spp) {
    uint32_t h = dim_key(pkey, dim);
    uint32_t idx = permute_pow2(s, spp, h);
    return van_der_corput(idx, mix32(h ^ 0x68bc21ebU));
}
__device__ __forceinline__ void ld2(uint32_t pkey, uint32_t dim, uint32_t s, uint32_t spp, float *out) {
    uint32_t h = dim_key(pkey, dim);
    uint32_t idx = permute_pow2(s, spp, h);
    out[0] = van_der_corput(idx, mix32(h ^ 0x68bc21ebU));
    out[1] = sobol2(idx, mix32(h ^ 0x02e5be93U));
}
__device__ __forceinline__ uint32_t rng_key(uint32_t pkey, uint32_t s) { return mix32(pkey ^ (0x10000u + s) * 0xc2b2ae35U); }
__device__ __forceinline__ float rng_float(uint32_t rkey, uint32_t k) {
    return (mix32(rkey + k * 0x27d4eb2fU) & 0xffffff) / (float)(1 << 24);          // RNG::RandomFloat, rng.cpp:51-57
}

#include "sampler_source.h"
// `pkey` below is pixel_key(seed, pix) of the path's sampler pixel

// The ten values bounce `b` consumes: {lightNum, lightPos0, lightPos1, lightComp, bsdfDir0,
// bsdfDir1, bsdfComp, pathDir0, pathDir1, pathComp} (src/integrators/path.cpp:33-41,63-83;
// src/core/integrator.cpp:84-99), and the Russian-roulette draw (path.cpp:97).
__device__ inline void bounce_dims(const SampleSource &src, uint64_t idx, uint32_t pkey, uint32_t s, int b,
                                   bool haveLights, float u[10], float *rr) {
    if (b < 3) {
        if (src.smp) {
            const float *oneD = src.smp + 37 * idx + 5 + 4 * b;
            const float *twoD = src.smp + 37 * idx + 19 + 6 * b;
            u[0] = oneD[1]; u[1] = twoD[0]; u[2] = twoD[1]; u[3] = oneD[0];
            u[4] = twoD[2]; u[5] = twoD[3]; u[6] = oneD[2];
            u[7] = twoD[4]; u[8] = twoD[5]; u[9] = oneD[3];
        } else {
            float t[2];
            u[3] = ld1(pkey, 3 + 4 * b + 0, s, src.spp);
            u[0] = ld1(pkey, 3 + 4 * b + 1, s, src.spp);
            u[6] = ld1(pkey, 3 + 4 * b + 2, s, src.spp);
            u[9] = ld1(pkey, 3 + 4 * b + 3, s, src.spp);
            ld2(pkey, 17 + 3 * b + 0, s, src.spp, t); u[1] = t[0]; u[2] = t[1];
            ld2(pkey, 17 + 3 * b + 1, s, src.spp, t); u[4] = t[0]; u[5] = t[1];
            ld2(pkey, 17 + 3 * b + 2, s, src.spp, t); u[7] = t[0]; u[8] = t[1];
        }
        *rr = 0.f;
        return;
    }
    // bounces >= 3: sequential RNG draws; a path that reaches bounce b has consumed a fixed count
    int perBounce = (haveLights ? 7 : 0) + 3;
    int k = (b - 3) * perBounce + (b > 4 ? b - 4 : 0);
    float v[11];
    uint32_t rkey = rng_key(pkey, s);
    for (int j = 0; j < perBounce + 1; ++j) {
        int kk = k + j;
        if (src.rng) v[j] = kk < src.n_rng ? src.rng[(size_t)src.n_rng * idx + kk] : 0.5f;
        else v[j] = rng_float(rkey, (uint32_t)kk);
    }
    int j = 0;
    if (haveLights) { for (; j < 7; ++j) u[j] = v[j]; }
    else { for (int q = 0; q < 7; ++q) u[q] = 0.f; }
    u[7] = v[j]; u[8] = v[j + 1]; u[9] = v[j + 2];
    *rr = v[j + 3];
}

// directlighting (strategy all): the six values light sample jj of light li consumes - {-, lightPos0, lightPos1, lightComp,
// bsdfDir0, bsdfDir1, bsdfComp} in u[1..6] (src/core/integrator.cpp:51-62). prefix = samples of the lights before li,
// n = this light's sample count (a power of two: Sampler::RoundSize), N = the sum over all lights.
// Caller-supplied vectors: the layout DirectLightingIntegrator::RequestSamples leaves (include/spt.h). Generated: the
// reference draws each array as ONE scrambled (0,2)-sequence of spp*n points, deals n consecutive points to every pixel
// sample and shuffles both the groups and the points within a group (LDShuffleScrambled1D/2D, montecarlo.h:289-315);
// here the group of pixel sample s and the place of jj within it are hashed permutations, as for the per-bounce values.
__device__ inline void direct_dims(const SampleSource &src, uint64_t cs, uint32_t pkey, uint32_t s, int li, int prefix, int n, int jj,
                                   int N, float u[10]) {
    if (src.smp) {
        const float *base = src.smp + (size_t)src.stride * cs;
        const float *oneD = base + 5 + 2 * prefix, *twoD = base + 7 + 2 * N + 4 * prefix;
        u[3] = oneD[jj]; u[6] = oneD[n + jj];
        u[1] = twoD[2 * jj]; u[2] = twoD[2 * jj + 1];
        u[4] = twoD[2 * n + 2 * jj]; u[5] = twoD[2 * n + 2 * jj + 1];
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t h = dim_key(pkey, 64u + 4u * (uint32_t)li + (uint32_t)k);
        uint32_t idx = permute_pow2(s, src.spp, h) * (uint32_t)n + permute_pow2((uint32_t)jj, (uint32_t)n, mix32(h ^ (s * 0x9e3779b9U + 0x7f4a7c15U)));
        float a = van_der_corput(idx, mix32(h ^ 0x68bc21ebU));
        if (k == 0) u[3] = a;
        else if (k == 1) u[6] = a;
        else {
            float b = sobol2(idx, mix32(h ^ 0x02e5be93U));
            if (k == 2) { u[1] = a; u[2] = b; } else { u[4] = a; u[5] = b; }
        }
    }
}

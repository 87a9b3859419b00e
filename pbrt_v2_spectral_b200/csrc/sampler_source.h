#pragma once
#include <stdint.h>
// Where a path's sample values come from: generated (production) or caller-supplied arrays in the
// reference's Sample memory order (spt_shade_samples).
struct SampleSource {
    const float *smp;       // n x stride (37 for the path integrator), or NULL
    int stride;
    const float *rng;       // n x n_rng, or NULL
    int n_rng;
    uint32_t seed, spp;
};

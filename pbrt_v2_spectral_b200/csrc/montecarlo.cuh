// montecarlo.cuh — sampling helpers of src/core/montecarlo.cpp / montecarlo.h, shared by the camera (exact TU)
// and the shading kernels.
#pragma once
#include "spt_device.cuh"

// ---- Monte Carlo helpers (src/core/montecarlo.cpp / montecarlo.h) ------------------------------
__device__ inline void concentric_sample_disk(float u1, float u2, float *dx, float *dy) {   // montecarlo.cpp:298-340
    float r, theta;
    float sx = 2 * u1 - 1;
    float sy = 2 * u2 - 1;
    if (sx == 0.0f && sy == 0.0f) { *dx = 0.0f; *dy = 0.0f; return; }
    if (sx >= -sy) {
        if (sx > sy) {
            r = sx;
            if (sy > 0.0f) theta = sy / r;
            else theta = 8.0f + sy / r;
        } else { r = sy; theta = 2.0f - sx / r; }
    } else {
        if (sx <= sy) { r = -sx; theta = 4.0f - sy / r; }
        else { r = -sy; theta = 6.0f + sx / r; }
    }
    theta *= PI_F / 4.f;
    float st, ct;
    sin_cos(theta, &st, &ct);
    *dx = r * ct;
    *dy = r * st;
}
__device__ inline v3 cosine_sample_hemisphere(float u1, float u2) {                        // montecarlo.h:120-125
    v3 ret;
    concentric_sample_disk(u1, u2, &ret.x, &ret.y);
    ret.z = sqrtf(stdmaxf(0.f, 1.f - ret.x * ret.x - ret.y * ret.y));
    return ret;
}
__device__ inline v3 uniform_sample_sphere(float u1, float u2) {                           // montecarlo.cpp:270-277
    float z = 1.f - 2.f * u1;
    float r = sqrtf(stdmaxf(0.f, 1.f - z * z));
    float phi = 2.f * PI_F * u2;
    float sp, cp;
    sin_cos(phi, &sp, &cp);
    return V(r * cp, r * sp, z);
}
__device__ inline v3 uniform_sample_cone(float u1, float u2, float costhetamax, v3 x, v3 y, v3 z) {   // :405-412
    float costheta = lerpf(u1, costhetamax, 1.f);
    float sintheta = sqrtf(1.f - costheta * costheta);
    float phi = u2 * 2.f * PI_F;
    float sp, cp;
    sin_cos(phi, &sp, &cp);
    return vadd(vadd(vmul(x, cp * sintheta), vmul(y, sp * sintheta)), vmul(z, costheta));
}
__device__ __forceinline__ float uniform_cone_pdf(float c) { return 1.f / (2.f * PI_F * (1.f - c)); }
__device__ __forceinline__ float power_heuristic(float fPdf, float gPdf) {                 // montecarlo.h:254-257
    float f = 1 * fPdf, g = 1 * gPdf;
    return (f * f) / (f * f + g * g);
}


// spt_device.cuh — device-side scene view and fp32 math with the reference's rounding behaviour.
//
// The reference is built -O2 -m64: scalar SSE, no FMA contraction (src/Makefile:24-29,64;
// SURVEY.md F8), and its Cross() is evaluated in double (src/core/geometry.h:479-486, F7).
// This TU is compiled with -fmad=false so `a*b + c` is an FMUL followed by an FADD exactly as on
// the CPU, with IEEE division and square root (nvcc defaults). Hit/miss decisions and primitive
// ids therefore match the reference bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "spt.h"

#define NB SPT_NBANDS            // valid bands
#define NBP SPT_BAND_PITCH       // floats per spectrum row in memory (bands NB.. are zero padding)
#define SPT_MISS 0xffffffffu
#define PI_F 3.14159265358979323846f
#define INV_PI_F 0.31830988618379067154f
#define INV_TWOPI_F 0.15915494309189533577f
#define ONE_MINUS_EPS 0x1.fffffep-1f
#define SPT_INF __int_as_float(0x7f800000)

// sinf and cosf of one angle: one range reduction on the device (sincosf returns what sinf / cosf return); the host build of the
// device sources (tests/host_shim) keeps the two calls the oracle makes
__device__ __forceinline__ void sin_cos(float x, float *s, float *c) {
#ifdef __CUDA_ARCH__
    sincosf(x, s, c);
#else
    *s = sinf(x); *c = cosf(x);
#endif
}
struct v3 { float x, y, z; };
__device__ __forceinline__ v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 vmul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float absdot(v3 a, v3 b) { return fabsf(dot(a, b)); }
__device__ __forceinline__ float len2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
// Cross in double, rounded once to float (geometry.h:479-486). Products of two floats are exact
// in double, so this is one rounded subtraction + one conversion per component.
__device__ __forceinline__ v3 cross(v3 a, v3 b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V(__double2float_rn(__dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by))),
             __double2float_rn(__dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz))),
             __double2float_rn(__dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx))));
}
__device__ __forceinline__ v3 vdiv(v3 a, float f) { float inv = 1.f / f; return V(a.x * inv, a.y * inv, a.z * inv); }
__device__ __forceinline__ v3 normalize(v3 a) { return vdiv(a, sqrtf(len2(a))); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ float lerpf(float t, float a, float b) { return (1.f - t) * a + t * b; }
__device__ __forceinline__ float stdmaxf(float a, float b) { return (a < b) ? b : a; }   // std::max semantics (NaN!)
__device__ __forceinline__ float stdminf(float a, float b) { return (b < a) ? b : a; }

__device__ __forceinline__ void coordinate_system(v3 v1, v3 *v2, v3 *v3o) {   // geometry.h:510-520
    if (fabsf(v1.x) > fabsf(v1.y)) {
        float invLen = 1.f / sqrtf(v1.x * v1.x + v1.z * v1.z);
        *v2 = V(-v1.z * invLen, 0.f, v1.x * invLen);
    } else {
        float invLen = 1.f / sqrtf(v1.y * v1.y + v1.z * v1.z);
        *v2 = V(0.f, v1.z * invLen, -v1.y * invLen);
    }
    *v3o = cross(v1, *v2);
}
// transform.h:184-241 (row-major 4x4)
__device__ __forceinline__ v3 xf_point(const float *m, v3 p) {
    float x = p.x, y = p.y, z = p.z;
    float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
    float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
    float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
    float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
    if (wp == 1.f) return V(xp, yp, zp);
    return vdiv(V(xp, yp, zp), wp);
}
__device__ __forceinline__ v3 xf_vector(const float *m, v3 v) {
    float x = v.x, y = v.y, z = v.z;
    return V(m[0] * x + m[1] * y + m[2] * z, m[4] * x + m[5] * y + m[6] * z, m[8] * x + m[9] * y + m[10] * z);
}
__device__ __forceinline__ v3 xf_normal(const float *minv, v3 n) {
    float x = n.x, y = n.y, z = n.z;
    return V(minv[0] * x + minv[4] * y + minv[8] * z, minv[1] * x + minv[5] * y + minv[9] * z,
             minv[2] * x + minv[6] * y + minv[10] * z);
}

struct Ray { v3 o, d; float mint, maxt; };
__device__ __forceinline__ v3 ray_at(const Ray &r, float t) { return vadd(r.o, vmul(r.d, t)); }

// What Shape::Intersect leaves behind (diffgeom.cpp:32-47, primitive.cpp:155-169)
struct Hit { float t, rayEpsilon; v3 p, dpdu, dpdv, nn; float u, v; };

// Scene tables resident in HBM. Layouts:
//   nodes      2 x float4 per LinearBVHNode, byte-identical to the reference's 32-byte node
//              (bvh.cpp:105-115): {pMin.xyz, pMax.x} {pMax.yz, offset, nPrims|axis<<8|hasQuadric<<16}
//   pnodes     64 B per INTERIOR node: child0 min/max, child1 min/max (12 floats), the two child codes
//              (pair index, or 1<<31 | hasQuadric<<30 | (nPrims-1)<<27 | first BVH slot for a leaf), split axis
//   tri_verts  3 x float4 per BVH slot (pre-gathered world-space p1,p2,p3; w unused) so a leaf test
//              is three 16-byte loads with no index indirection
struct DevScene {
    const float4 *nodes;
    const float4 *pnodes;            // pair nodes: 4 x float4 per interior node (both children's bounds + codes)
    uint32_t root_code;              // child code of the root
    const float4 *wnodes;            // fast mode (wide.h): 8 x float4 per 4-wide node; NULL until spt_scene_set_traversal(FAST)
    uint32_t wroot;                  // its root's child code
    const float4 *tri_verts;
    uint32_t n_nodes, n_prims;
    const uint8_t *prim_kind, *prim_flags;
    const uint32_t *prim_id, *prim_data;
    const int32_t *prim_material, *prim_light, *prim_xform;
    const int32_t *tri_vidx;
    const float *P, *N, *UV;
    const SptQuadric *quadrics;
    const SptXform *xforms;
    const SptMaterial *materials;
    const SptLight *lights;
    const SptLightShape *light_shapes;
    const float *light_cdf;          // per light: shape_count+1 floats at offset shape_first + light index
    uint32_t n_lights;
    int has_specular;                // some material is a mirror / glass: paths carry the specularBounce flag (path.cpp:50,86)
    const SptSpectralTables *tables;
    int env_w, env_h;
    const float *env_rgb, *env_func, *env_cdf, *env_func_int, *env_marg_func, *env_marg_cdf;
    float env_marg_int;
    // image textures (SptTexture rows, the texel pool of their MIP pyramids, MIPMap::weightLut) and whether any
    // material needs the extended shading kernels (substrate BSDF, image-mapped Kd, bump map)
    const SptTexture *textures;
    const float *tex_texels;
    const float *ewa_lut;
    int has_ext;
    // measured BRDFs: the reference's kd-trees (SptKdNode rows + one spectrum per node)
    const SptBrdfTable *brdfs;
    const SptKdNode *brdf_nodes;
    const float *brdf_spectra;
    const float *merl_rgb;           // half-angle (MERL) tables: RGB triples
    int has_measured;
    unsigned long long *counters;    // closest: [0] nodes, [1] prim tests; any-hit: [2], [3] (NULL when disabled)
};

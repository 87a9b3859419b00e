// traverse.cuh — ray/primitive tests and BVH traversal with the reference's exact semantics.
#pragma once
#include "spt_device.cuh"

// ---- Triangle (src/shapes/trianglemesh.cpp:119-200 Intersect, :203-273 IntersectP) -------------
// Accept test only: returns t through *tout. Vertices come pre-gathered per BVH slot.
__device__ __forceinline__ bool tri_test(v3 p1, v3 p2, v3 p3, const Ray &ray, float *tout,
                                         float *b1out, float *b2out) {
    v3 e1 = vsub(p2, p1), e2 = vsub(p3, p1);
    v3 s1 = cross(ray.d, e2);
    float divisor = dot(s1, e1);
    if (divisor == 0.f) return false;
    float invDivisor = 1.f / divisor;
    v3 d = vsub(ray.o, p1);
    float b1 = dot(d, s1) * invDivisor;
    if (b1 < 0.f || b1 > 1.f) return false;
    v3 s2 = cross(d, e1);
    float b2 = dot(ray.d, s2) * invDivisor;
    if (b2 < 0.f || b1 + b2 > 1.f) return false;
    float t = dot(e2, s2) * invDivisor;
    if (t < ray.mint || t > ray.maxt) return false;
    *tout = t; *b1out = b1; *b2out = b2;
    return true;
}

__device__ __forceinline__ void dg_init(Hit *h, v3 p, v3 dpdu, v3 dpdv, float u, float v, int flags) {
    h->p = p; h->dpdu = dpdu; h->dpdv = dpdv; h->u = u; h->v = v;     // diffgeom.cpp:32-47
    h->nn = normalize(cross(dpdu, dpdv));
    if (flags & SPT_PF_FLIP_NORMAL) h->nn = vmul(h->nn, -1.f);
}

__device__ __forceinline__ void tri_uvs(const DevScene &sc, int flags, const int32_t *vi, float uv[3][2]) {
    if (flags & SPT_PF_HAS_UV) {                                       // trianglemesh.h:78-92
#pragma unroll
        for (int k = 0; k < 3; ++k) { uv[k][0] = sc.UV[2 * (size_t)vi[k]]; uv[k][1] = sc.UV[2 * (size_t)vi[k] + 1]; }
    } else {
        uv[0][0] = 0.f; uv[0][1] = 0.f; uv[1][0] = 1.f; uv[1][1] = 0.f; uv[2][0] = 1.f; uv[2][1] = 1.f;
    }
}

// DifferentialGeometry of a triangle hit (trianglemesh.cpp:146-200) for barycentrics b1,b2 and distance t.
__device__ inline void tri_fill(const DevScene &sc, int flags, const int32_t *vi, v3 p1, v3 p2, v3 p3, const Ray &ray,
                                float t, float b1, float b2, Hit *hit) {
    v3 e1 = vsub(p2, p1), e2 = vsub(p3, p1);
    v3 dpdu, dpdv;
    float uvs[3][2];
    tri_uvs(sc, flags, vi, uvs);
    float du1 = uvs[0][0] - uvs[2][0], du2 = uvs[1][0] - uvs[2][0];
    float dv1 = uvs[0][1] - uvs[2][1], dv2 = uvs[1][1] - uvs[2][1];
    v3 dp1 = vsub(p1, p3), dp2 = vsub(p2, p3);
    float determinant = du1 * dv2 - dv1 * du2;
    if (determinant == 0.f) {
        coordinate_system(normalize(cross(e2, e1)), &dpdu, &dpdv);
    } else {
        float invdet = 1.f / determinant;
        dpdu = vmul(vsub(vmul(dp1, dv2), vmul(dp2, dv1)), invdet);
        dpdv = vmul(vadd(vmul(dp1, -du2), vmul(dp2, du1)), invdet);
    }
    float b0 = 1 - b1 - b2;
    float tu = b0 * uvs[0][0] + b1 * uvs[1][0] + b2 * uvs[2][0];
    float tv = b0 * uvs[0][1] + b1 * uvs[1][1] + b2 * uvs[2][1];
    dg_init(hit, ray_at(ray, t), dpdu, dpdv, tu, tv, flags);
    hit->t = t;
    hit->rayEpsilon = 1e-3f * t;
}
__device__ __forceinline__ void tri_load(const DevScene &sc, uint32_t tri, const int32_t **vi, v3 *p1, v3 *p2, v3 *p3) {
    const int32_t *v = sc.tri_vidx + 3 * (size_t)tri;
    int32_t i0 = v[0], i1 = v[1], i2 = v[2];
    *vi = v;
    *p1 = V(sc.P[3 * (size_t)i0], sc.P[3 * (size_t)i0 + 1], sc.P[3 * (size_t)i0 + 2]);
    *p2 = V(sc.P[3 * (size_t)i1], sc.P[3 * (size_t)i1 + 1], sc.P[3 * (size_t)i1 + 2]);
    *p3 = V(sc.P[3 * (size_t)i2], sc.P[3 * (size_t)i2 + 1], sc.P[3 * (size_t)i2 + 2]);
}
// Full Triangle::Intersect with DifferentialGeometry, given triangle number `tri` (index path).
__device__ inline bool tri_intersect_full(const DevScene &sc, uint32_t tri, int flags, const Ray &ray, Hit *hit) {
    const int32_t *vi; v3 p1, p2, p3;
    tri_load(sc, tri, &vi, &p1, &p2, &p3);
    float t, b1, b2;
    if (!tri_test(p1, p2, p3, ray, &t, &b1, &b2)) return false;
    tri_fill(sc, flags, vi, p1, p2, p3, ray, t, b1, b2, hit);
    return true;
}
// The hit record of a triangle the traversal has ALREADY accepted at distance t: same arithmetic as
// Triangle::Intersect for b1,b2, none of its rejection tests (they were decided, in the reference's
// rounding, by the traversal kernel; re-deciding them under FMA contraction could disagree).
__device__ inline void tri_record(const DevScene &sc, uint32_t tri, int flags, const Ray &ray, float t, Hit *hit) {
    const int32_t *vi; v3 p1, p2, p3;
    tri_load(sc, tri, &vi, &p1, &p2, &p3);
    v3 e1 = vsub(p2, p1), e2 = vsub(p3, p1);
    v3 s1 = cross(ray.d, e2);
    float invDivisor = 1.f / dot(s1, e1);
    v3 d = vsub(ray.o, p1);
    float b1 = dot(d, s1) * invDivisor;
    v3 s2 = cross(d, e1);
    float b2 = dot(ray.d, s2) * invDivisor;
    tri_fill(sc, flags, vi, p1, p2, p3, ray, t, b1, b2, hit);
}

__device__ __forceinline__ bool quadratic(float A, float B, float C, float *t0, float *t1) {   // pbrt.h:297-311
    float discrim = B * B - 4.f * A * C;
    if (discrim <= 0.f) return false;
    float rootDiscrim = sqrtf(discrim);
    float q;
    if (B < 0) q = -.5f * (B - rootDiscrim);
    else q = -.5f * (B + rootDiscrim);
    *t0 = q / A;
    *t1 = C / q;
    if (*t0 > *t1) { float tmp = *t0; *t0 = *t1; *t1 = tmp; }
    return true;
}

__device__ inline void sphere_fill(const SptQuadric &q, const SptXform &xf, int flags, v3 phit, float phi, float thit, Hit *hit);
// Sphere (src/shapes/sphere.cpp:50-149, :152-201). hit == NULL: accept test only; phit_obj: the object-space hit point.
__device__ inline bool sphere_intersect(const DevScene &sc, const SptQuadric &q, int flags, const Ray &r,
                                        float *tout, Hit *hit, v3 *phit_obj = nullptr) {
    const SptXform &xf = sc.xforms[q.xform];
    Ray ray = r;
    ray.o = xf_point(xf.minv, r.o);
    ray.d = xf_vector(xf.minv, r.d);
    float radius = q.radius, zmin = q.zmin, zmax = q.zmax, phiMax = q.phiMax;
    float A = ray.d.x * ray.d.x + ray.d.y * ray.d.y + ray.d.z * ray.d.z;
    float B = 2 * (ray.d.x * ray.o.x + ray.d.y * ray.o.y + ray.d.z * ray.o.z);
    float C = ray.o.x * ray.o.x + ray.o.y * ray.o.y + ray.o.z * ray.o.z - radius * radius;
    float t0, t1;
    if (!quadratic(A, B, C, &t0, &t1)) return false;
    if (t0 > ray.maxt || t1 < ray.mint) return false;
    float thit = t0;
    if (t0 < ray.mint) {
        thit = t1;
        if (thit > ray.maxt) return false;
    }
    // A sphere that is not clipped (zmin <= -r, zmax >= r, phiMax = 2 pi: Radians(360) rounds to the float
    // nearest 2 pi, which atan2f(...) + 2 pi never exceeds) accepts its first root: the phi / z tests
    // of sphere.cpp:81-99 cannot reject, so an accept-only query skips atan2f.
    const bool clipped = zmin > -radius || zmax < radius || phiMax < 2.f * PI_F;
    v3 phit = V(0, 0, 0);
    float phi = 0.f;
    if (clipped || hit) {
        phit = ray_at(ray, thit);
        if (phit.x == 0.f && phit.y == 0.f) phit.x = 1e-5f * radius;
        phi = atan2f(phit.y, phit.x);
        if (phi < 0.f) phi += 2.f * PI_F;
    }
    if (clipped && ((zmin > -radius && phit.z < zmin) || (zmax < radius && phit.z > zmax) || phi > phiMax)) {
        if (thit == t1) return false;
        if (t1 > ray.maxt) return false;
        thit = t1;
        phit = ray_at(ray, thit);
        if (phit.x == 0.f && phit.y == 0.f) phit.x = 1e-5f * radius;
        phi = atan2f(phit.y, phit.x);
        if (phi < 0.f) phi += 2.f * PI_F;
        if ((zmin > -radius && phit.z < zmin) || (zmax < radius && phit.z > zmax) || phi > phiMax) return false;
    }
    *tout = thit;
    if (phit_obj) *phit_obj = (clipped || hit) ? phit : ray_at(ray, thit);
    if (!hit) return true;
    sphere_fill(q, xf, flags, phit, phi, thit, hit);
    return true;
}
// DifferentialGeometry of a sphere hit at object-space point phit (sphere.cpp:105-149)
__device__ inline void sphere_fill(const SptQuadric &q, const SptXform &xf, int flags, v3 phit, float phi, float thit, Hit *hit) {
    float radius = q.radius, phiMax = q.phiMax;
    float u = phi / phiMax;
    float theta = acosf(clampf(phit.z / radius, -1.f, 1.f));
    float v = (theta - q.thetaMin) / (q.thetaMax - q.thetaMin);
    float zradius = sqrtf(phit.x * phit.x + phit.y * phit.y);
    float invzradius = 1.f / zradius;
    float cosphi = phit.x * invzradius;
    float sinphi = phit.y * invzradius;
    v3 dpdu = V(-phiMax * phit.y, phiMax * phit.x, 0);
    v3 dpdv = vmul(V(phit.z * cosphi, phit.z * sinphi, -radius * sinf(theta)), q.thetaMax - q.thetaMin);
    dg_init(hit, xf_point(xf.m, phit), xf_vector(xf.m, dpdu), xf_vector(xf.m, dpdv), u, v, flags);
    hit->t = thit;
    hit->rayEpsilon = 5e-4f * thit;
}
// The hit record of a sphere the traversal accepted at distance t (no re-decision, see tri_record)
__device__ inline void sphere_record(const DevScene &sc, const SptQuadric &q, int flags, const Ray &r, float t, Hit *hit) {
    const SptXform &xf = sc.xforms[q.xform];
    Ray ray = r;
    ray.o = xf_point(xf.minv, r.o);
    ray.d = xf_vector(xf.minv, r.d);
    v3 phit = ray_at(ray, t);
    if (phit.x == 0.f && phit.y == 0.f) phit.x = 1e-5f * q.radius;
    float phi = atan2f(phit.y, phit.x);
    if (phi < 0.f) phi += 2.f * PI_F;
    sphere_fill(q, xf, flags, phit, phi, t, hit);
}

__device__ inline void disk_fill(const SptQuadric &q, const SptXform &xf, int flags, v3 phit, float phi, float dist2, float thit, Hit *hit);
// Disk (src/shapes/disk.cpp:48-95, :98-121)
__device__ inline bool disk_intersect(const DevScene &sc, const SptQuadric &q, int flags, const Ray &r,
                                      float *tout, Hit *hit) {
    const SptXform &xf = sc.xforms[q.xform];
    Ray ray = r;
    ray.o = xf_point(xf.minv, r.o);
    ray.d = xf_vector(xf.minv, r.d);
    float height = q.zmin, radius = q.radius, innerRadius = q.zmax, phiMax = q.phiMax;
    if ((double)fabsf(ray.d.z) < 1e-7) return false;        // float against the DOUBLE literal, as the reference
    float thit = (height - ray.o.z) / ray.d.z;
    if (thit < ray.mint || thit > ray.maxt) return false;
    v3 phit = ray_at(ray, thit);
    float dist2 = phit.x * phit.x + phit.y * phit.y;
    if (dist2 > radius * radius || dist2 < innerRadius * innerRadius) return false;
    float phi = atan2f(phit.y, phit.x);
    if (phi < 0) phi = (float)((double)phi + 2. * (double)PI_F);
    if (phi > phiMax) return false;
    *tout = thit;
    if (!hit) return true;
    disk_fill(q, xf, flags, phit, phi, dist2, thit, hit);
    return true;
}
__device__ inline void disk_fill(const SptQuadric &q, const SptXform &xf, int flags, v3 phit, float phi, float dist2, float thit, Hit *hit) {
    float radius = q.radius, innerRadius = q.zmax, phiMax = q.phiMax;
    float u = phi / phiMax;
    float oneMinusV = ((sqrtf(dist2) - innerRadius) / (radius - innerRadius));
    float invOneMinusV = (oneMinusV > 0.f) ? (1.f / oneMinusV) : 0.f;
    float v = 1.f - oneMinusV;
    v3 dpdu = V(-phiMax * phit.y, phiMax * phit.x, 0.f);
    v3 dpdv = V(-phit.x * invOneMinusV, -phit.y * invOneMinusV, 0.f);
    dpdu = vmul(dpdu, phiMax * INV_TWOPI_F);
    dpdv = vmul(dpdv, (radius - innerRadius) / radius);
    dg_init(hit, xf_point(xf.m, phit), xf_vector(xf.m, dpdu), xf_vector(xf.m, dpdv), u, v, flags);
    hit->t = thit;
    hit->rayEpsilon = 5e-4f * thit;
}
__device__ inline void disk_record(const DevScene &sc, const SptQuadric &q, int flags, const Ray &r, float t, Hit *hit) {
    const SptXform &xf = sc.xforms[q.xform];
    Ray ray = r;
    ray.o = xf_point(xf.minv, r.o);
    ray.d = xf_vector(xf.minv, r.d);
    v3 phit = ray_at(ray, t);
    float dist2 = phit.x * phit.x + phit.y * phit.y;
    float phi = atan2f(phit.y, phit.x);
    if (phi < 0) phi = (float)((double)phi + 2. * (double)PI_F);
    disk_fill(q, xf, flags, phit, phi, dist2, t, hit);
}

// Shape::Intersect by (kind, flags, data); used for recomputing the hit record in the shading
// kernels and for the ShapeSet / Shape::Pdf re-intersections of light sampling.
__device__ inline bool shape_intersect(const DevScene &sc, int kind, int flags, uint32_t data, const Ray &ray, Hit *hit) {
    float t;
    if (kind == SPT_PRIM_TRIANGLE) return tri_intersect_full(sc, data, flags, ray, hit);
    if (kind == SPT_PRIM_SPHERE) return sphere_intersect(sc, sc.quadrics[data], flags, ray, &t, hit);
    return disk_intersect(sc, sc.quadrics[data], flags, ray, &t, hit);
}

// The hit record of the primitive the traversal accepted at distance t
__device__ inline void shape_record(const DevScene &sc, int kind, int flags, uint32_t data, const Ray &ray, float t, Hit *hit) {
    if (kind == SPT_PRIM_TRIANGLE) tri_record(sc, data, flags, ray, t, hit);
    else if (kind == SPT_PRIM_SPHERE) sphere_record(sc, sc.quadrics[data], flags, ray, t, hit);
    else disk_record(sc, sc.quadrics[data], flags, ray, t, hit);
}

// ---- slab test (src/accelerators/bvh.cpp:118-140) ----------------------------------------------
// n0 = {pMin.x,pMin.y,pMin.z,pMax.x}, n1 = {pMax.y,pMax.z,..}. neg* select which bound is "near".
__device__ __forceinline__ bool slab(const float4 &n0, const float4 &n1, const Ray &ray, v3 invDir,
                                     bool negx, bool negy, bool negz) {
    float bx0 = negx ? n0.w : n0.x, bx1 = negx ? n0.x : n0.w;
    float by0 = negy ? n1.x : n0.y, by1 = negy ? n0.y : n1.x;
    float tmin = (bx0 - ray.o.x) * invDir.x;
    float tmax = (bx1 - ray.o.x) * invDir.x;
    float tymin = (by0 - ray.o.y) * invDir.y;
    float tymax = (by1 - ray.o.y) * invDir.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float bz0 = negz ? n1.y : n0.z, bz1 = negz ? n0.z : n1.y;
    float tzmin = (bz0 - ray.o.z) * invDir.z;
    float tzmax = (bz1 - ray.o.z) * invDir.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return (tmin < ray.maxt) && (tmax > ray.mint);
}


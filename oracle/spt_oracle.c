/* spt_oracle.c — TEST INFRASTRUCTURE (see spt_oracle.h). Plain-C restatement of the reference's
 * hot path, one sample at a time, in the reference's own operation order (scalar fp32, no FMA:
 * build with -O2 -ffp-contract=off, never -ffast-math / -march=native — SURVEY.md F8).
 * Every function cites the reference lines it follows (paths relative to /root/reference/src).
 * The product path must never link or load this file. */
#include <alloca.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "spt_oracle.h"

#define NB SPT_NBANDS
#define PI_F 3.14159265358979323846f          /* core/pbrt.h:179 */
#define INV_PI_F 0.31830988618379067154f
#define INV_TWOPI_F 0.15915494309189533577f
#define ONE_MINUS_EPS 0x1.fffffep-1f            /* core/montecarlo.h:42 */

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float absdot(v3 a, v3 b) { return fabsf(dot(a, b)); }
static inline float len2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
/* core/geometry.h:479-486: cross product evaluated in double, rounded to float */
static inline v3 cross(v3 a, v3 b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V((float)((ay * bz) - (az * by)), (float)((az * bx) - (ax * bz)), (float)((ax * by) - (ay * bx)));
}
/* core/geometry.h:86-90,509: v / Length() == v * (1.f / len) */
static inline v3 vdiv(v3 a, float f) { float inv = 1.f / f; return V(a.x * inv, a.y * inv, a.z * inv); }
static inline v3 normalize(v3 a) { return vdiv(a, sqrtf(len2(a))); }
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float lerpf(float t, float a, float b) { return (1.f - t) * a + t * b; }
static inline float maxf(float a, float b) { return a > b ? a : b; }      /* std::max: (a<b)?b:a */
static inline float minf(float a, float b) { return b < a ? b : a; }      /* std::min */
static inline float stdmaxf(float a, float b) { return (a < b) ? b : a; }
static inline float stdminf(float a, float b) { return (b < a) ? b : a; }
/* core/geometry.h:510-520 */
static inline void coordinate_system(v3 v1, v3 *v2, v3 *v3o) {
    if (fabsf(v1.x) > fabsf(v1.y)) {
        float invLen = 1.f / sqrtf(v1.x * v1.x + v1.z * v1.z);
        *v2 = V(-v1.z * invLen, 0.f, v1.x * invLen);
    } else {
        float invLen = 1.f / sqrtf(v1.y * v1.y + v1.z * v1.z);
        *v2 = V(0.f, v1.z * invLen, -v1.y * invLen);
    }
    *v3o = cross(v1, *v2);
}

/* core/transform.h:184-241 */
static inline v3 xf_point(const float *m, v3 p) {
    float x = p.x, y = p.y, z = p.z;
    float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
    float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
    float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
    float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
    if (wp == 1.) return V(xp, yp, zp);
    return vdiv(V(xp, yp, zp), wp);
}
static inline v3 xf_vector(const float *m, v3 v) {
    float x = v.x, y = v.y, z = v.z;
    return V(m[0] * x + m[1] * y + m[2] * z, m[4] * x + m[5] * y + m[6] * z, m[8] * x + m[9] * y + m[10] * z);
}
static inline v3 xf_normal(const float *minv, v3 n) {   /* uses the inverse transposed */
    float x = n.x, y = n.y, z = n.z;
    return V(minv[0] * x + minv[4] * y + minv[8] * z, minv[1] * x + minv[5] * y + minv[9] * z,
             minv[2] * x + minv[6] * y + minv[10] * z);
}

typedef struct { v3 o, d; float mint, maxt; int depth; } Ray;
static inline v3 ray_at(const Ray *r, float t) { return vadd(r->o, vmul(r->d, t)); }

/* ------------------------------------------------------------------------------------------ */
/* K1: PerspectiveCamera::GenerateRayDifferential, cameras/perspective.cpp:73-106. The offset rays
 * (rd != NULL) are consumed by image textures only; they are scaled as
 * RayDifferential::ScaleDifferentials does for 1/sqrt(spp) (renderers/samplerrenderer.cpp:91,
 * core/geometry.h:368-373). */
typedef struct { int has; v3 rxo, ryo, rxd, ryd; } RayDiff;
static void concentric_sample_disk(float u1, float u2, float *dx, float *dy);

static void camera_ray_diff(const SptCameraDesc *cam, const float *s, int spp, Ray *ray, RayDiff *rd) {
    v3 Pras = V(s[0], s[1], 0.f);
    v3 Pcamera = xf_point(cam->raster_to_camera, Pras);
    v3 dir = normalize(Pcamera);
    ray->o = V(0, 0, 0);
    ray->d = dir;
    ray->mint = 0.f;
    ray->maxt = INFINITY;
    ray->depth = 0;
    if (cam->lens_radius > 0.) {
        float lensU, lensV;
        concentric_sample_disk(s[2], s[3], &lensU, &lensV);
        lensU *= cam->lens_radius;
        lensV *= cam->lens_radius;
        float ft = cam->focal_distance / ray->d.z;
        v3 Pfocus = ray_at(ray, ft);
        ray->o = V(lensU, lensV, 0.f);
        ray->d = normalize(vsub(Pfocus, ray->o));
    }
    if (rd) {
        v3 rxd = normalize(vadd(Pcamera, V(cam->dx_camera[0], cam->dx_camera[1], cam->dx_camera[2])));
        v3 ryd = normalize(vadd(Pcamera, V(cam->dy_camera[0], cam->dy_camera[1], cam->dy_camera[2])));
        rd->rxo = rd->ryo = xf_point(cam->camera_to_world, ray->o);
        rd->rxd = xf_vector(cam->camera_to_world, rxd);
        rd->ryd = xf_vector(cam->camera_to_world, ryd);
        rd->has = 1;
    }
    ray->o = xf_point(cam->camera_to_world, ray->o);
    ray->d = xf_vector(cam->camera_to_world, ray->d);
    if (rd) {
        float sc = 1.f / sqrtf((float)spp);
        rd->rxo = vadd(ray->o, vmul(vsub(rd->rxo, ray->o), sc));
        rd->ryo = vadd(ray->o, vmul(vsub(rd->ryo, ray->o), sc));
        rd->rxd = vadd(ray->d, vmul(vsub(rd->rxd, ray->d), sc));
        rd->ryd = vadd(ray->d, vmul(vsub(rd->ryd, ray->d), sc));
    }
}
static void camera_ray(const SptCameraDesc *cam, const float *s, Ray *ray) { camera_ray_diff(cam, s, 1, ray, NULL); }

void orc_camera_rays(const SptCameraDesc *cam, const float *samples, uint64_t n, float *out) {
    for (uint64_t i = 0; i < n; ++i) {
        Ray r;
        camera_ray(cam, samples + 5 * i, &r);
        float *o = out + 8 * i;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
        o[6] = r.mint; o[7] = r.maxt;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* geometry of one hit: what Shape::Intersect leaves in the DifferentialGeometry
 * (core/diffgeom.cpp:32-47) plus GeometricPrimitive::Intersect's bookkeeping (core/primitive.cpp:155-169) */
typedef struct {
    float t, rayEpsilon;
    v3 p, dpdu, dpdv, nn;
    float u, v;
} Hit;

typedef struct { float b[6]; uint32_t offset; uint8_t nPrims, axis, pad[2]; } Node;   /* accelerators/bvh.cpp:105-115 */

static inline v3 vert(const SptSceneDesc *sc, int idx) { const float *p = sc->P + 3 * (size_t)idx; return V(p[0], p[1], p[2]); }

static void tri_uvs(const SptSceneDesc *sc, int flags, const int32_t *vi, float uv[3][2]) {
    /* shapes/trianglemesh.h:78-92 */
    if (flags & SPT_PF_HAS_UV) {
        for (int k = 0; k < 3; ++k) { uv[k][0] = sc->UV[2 * (size_t)vi[k]]; uv[k][1] = sc->UV[2 * (size_t)vi[k] + 1]; }
    } else {
        uv[0][0] = 0.f; uv[0][1] = 0.f; uv[1][0] = 1.f; uv[1][1] = 0.f; uv[2][0] = 1.f; uv[2][1] = 1.f;
    }
}

static void dg_init(Hit *h, v3 p, v3 dpdu, v3 dpdv, float u, float v, int flags) {
    /* core/diffgeom.cpp:32-47 */
    h->p = p; h->dpdu = dpdu; h->dpdv = dpdv; h->u = u; h->v = v;
    h->nn = normalize(cross(dpdu, dpdv));
    if (flags & SPT_PF_FLIP_NORMAL) h->nn = vmul(h->nn, -1.f);
}

/* shapes/trianglemesh.cpp:119-200 (Intersect) and :203-273 (IntersectP: hit == NULL) */
static int tri_intersect(const SptSceneDesc *sc, uint32_t tri, int flags, const Ray *ray, Hit *hit) {
    const int32_t *vi = sc->tri_vidx + 3 * (size_t)tri;
    v3 p1 = vert(sc, vi[0]), p2 = vert(sc, vi[1]), p3 = vert(sc, vi[2]);
    v3 e1 = vsub(p2, p1), e2 = vsub(p3, p1);
    v3 s1 = cross(ray->d, e2);
    float divisor = dot(s1, e1);
    if (divisor == 0.) return 0;
    float invDivisor = 1.f / divisor;
    v3 d = vsub(ray->o, p1);
    float b1 = dot(d, s1) * invDivisor;
    if (b1 < 0. || b1 > 1.) return 0;
    v3 s2 = cross(d, e1);
    float b2 = dot(ray->d, s2) * invDivisor;
    if (b2 < 0. || b1 + b2 > 1.) return 0;
    float t = dot(e2, s2) * invDivisor;
    if (t < ray->mint || t > ray->maxt) return 0;
    if (!hit) return 1;
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
    return 1;
}

/* core/pbrt.h:297-311 */
static int quadratic(float A, float B, float C, float *t0, float *t1) {
    float discrim = B * B - 4.f * A * C;
    if (discrim <= 0.) return 0;
    float rootDiscrim = sqrtf(discrim);
    float q;
    if (B < 0) q = -.5f * (B - rootDiscrim);
    else q = -.5f * (B + rootDiscrim);
    *t0 = q / A;
    *t1 = C / q;
    if (*t0 > *t1) { float tmp = *t0; *t0 = *t1; *t1 = tmp; }
    return 1;
}

/* shapes/sphere.cpp:50-149 (Intersect), :152-201 (IntersectP: hit == NULL) */
static int sphere_intersect(const SptSceneDesc *sc, const SptQuadric *q, int flags, const Ray *r, Hit *hit) {
    const SptXform *xf = sc->xforms + q->xform;
    Ray ray = *r;
    ray.o = xf_point(xf->minv, r->o);
    ray.d = xf_vector(xf->minv, r->d);
    float radius = q->radius, zmin = q->zmin, zmax = q->zmax, phiMax = q->phiMax;
    float A = ray.d.x * ray.d.x + ray.d.y * ray.d.y + ray.d.z * ray.d.z;
    float B = 2 * (ray.d.x * ray.o.x + ray.d.y * ray.o.y + ray.d.z * ray.o.z);
    float C = ray.o.x * ray.o.x + ray.o.y * ray.o.y + ray.o.z * ray.o.z - radius * radius;
    float t0, t1;
    if (!quadratic(A, B, C, &t0, &t1)) return 0;
    if (t0 > ray.maxt || t1 < ray.mint) return 0;
    float thit = t0;
    if (t0 < ray.mint) {
        thit = t1;
        if (thit > ray.maxt) return 0;
    }
    v3 phit = ray_at(&ray, thit);
    if (phit.x == 0.f && phit.y == 0.f) phit.x = 1e-5f * radius;
    float phi = atan2f(phit.y, phit.x);
    if (phi < 0.) phi += 2.f * PI_F;
    if ((zmin > -radius && phit.z < zmin) || (zmax < radius && phit.z > zmax) || phi > phiMax) {
        if (thit == t1) return 0;
        if (t1 > ray.maxt) return 0;
        thit = t1;
        phit = ray_at(&ray, thit);
        if (phit.x == 0.f && phit.y == 0.f) phit.x = 1e-5f * radius;
        phi = atan2f(phit.y, phit.x);
        if (phi < 0.) phi += 2.f * PI_F;
        if ((zmin > -radius && phit.z < zmin) || (zmax < radius && phit.z > zmax) || phi > phiMax) return 0;
    }
    if (!hit) return 1;
    float u = phi / phiMax;
    float theta = acosf(clampf(phit.z / radius, -1.f, 1.f));
    float v = (theta - q->thetaMin) / (q->thetaMax - q->thetaMin);
    float zradius = sqrtf(phit.x * phit.x + phit.y * phit.y);
    float invzradius = 1.f / zradius;
    float cosphi = phit.x * invzradius;
    float sinphi = phit.y * invzradius;
    v3 dpdu = V(-phiMax * phit.y, phiMax * phit.x, 0);
    v3 dpdv = vmul(V(phit.z * cosphi, phit.z * sinphi, -radius * sinf(theta)), q->thetaMax - q->thetaMin);
    dg_init(hit, xf_point(xf->m, phit), xf_vector(xf->m, dpdu), xf_vector(xf->m, dpdv), u, v, flags);
    hit->t = thit;
    hit->rayEpsilon = 5e-4f * thit;
    return 1;
}

/* shapes/disk.cpp:48-95 (Intersect), :98-121 (IntersectP: hit == NULL) */
static int disk_intersect(const SptSceneDesc *sc, const SptQuadric *q, int flags, const Ray *r, Hit *hit) {
    const SptXform *xf = sc->xforms + q->xform;
    Ray ray = *r;
    ray.o = xf_point(xf->minv, r->o);
    ray.d = xf_vector(xf->minv, r->d);
    float height = q->zmin, radius = q->radius, innerRadius = q->zmax, phiMax = q->phiMax;
    if (fabsf(ray.d.z) < 1e-7) return 0;                       /* float vs DOUBLE 1e-7, as written */
    float thit = (height - ray.o.z) / ray.d.z;
    if (thit < ray.mint || thit > ray.maxt) return 0;
    v3 phit = ray_at(&ray, thit);
    float dist2 = phit.x * phit.x + phit.y * phit.y;
    if (dist2 > radius * radius || dist2 < innerRadius * innerRadius) return 0;
    float phi = atan2f(phit.y, phit.x);
    if (phi < 0) phi += 2. * PI_F;                              /* double arithmetic, as written */
    if (phi > phiMax) return 0;
    if (!hit) return 1;
    float u = phi / phiMax;
    float oneMinusV = ((sqrtf(dist2) - innerRadius) / (radius - innerRadius));
    float invOneMinusV = (oneMinusV > 0.f) ? (1.f / oneMinusV) : 0.f;
    float v = 1.f - oneMinusV;
    v3 dpdu = V(-phiMax * phit.y, phiMax * phit.x, 0.);
    v3 dpdv = V(-phit.x * invOneMinusV, -phit.y * invOneMinusV, 0.);
    dpdu = vmul(dpdu, phiMax * INV_TWOPI_F);
    dpdv = vmul(dpdv, (radius - innerRadius) / radius);
    dg_init(hit, xf_point(xf->m, phit), xf_vector(xf->m, dpdu), xf_vector(xf->m, dpdv), u, v, flags);
    hit->t = thit;
    hit->rayEpsilon = 5e-4f * thit;
    return 1;
}

static int shape_intersect(const SptSceneDesc *sc, int kind, int flags, uint32_t data, const Ray *ray, Hit *hit) {
    if (kind == SPT_PRIM_TRIANGLE) return tri_intersect(sc, data, flags, ray, hit);
    if (kind == SPT_PRIM_SPHERE) return sphere_intersect(sc, sc->quadrics + data, flags, ray, hit);
    return disk_intersect(sc, sc->quadrics + data, flags, ray, hit);
}

/* accelerators/bvh.cpp:118-140 */
static inline int slab(const float *b, const Ray *ray, v3 invDir, const uint32_t neg[3]) {
    float tmin = (b[3 * neg[0] + 0] - ray->o.x) * invDir.x;
    float tmax = (b[3 * (1 - neg[0]) + 0] - ray->o.x) * invDir.x;
    float tymin = (b[3 * neg[1] + 1] - ray->o.y) * invDir.y;
    float tymax = (b[3 * (1 - neg[1]) + 1] - ray->o.y) * invDir.y;
    if ((tmin > tymax) || (tymin > tmax)) return 0;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (b[3 * neg[2] + 2] - ray->o.z) * invDir.z;
    float tzmax = (b[3 * (1 - neg[2]) + 2] - ray->o.z) * invDir.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return 0;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return (tmin < ray->maxt) && (tmax > ray->mint);
}

/* accelerators/bvh.cpp:380-432 (closest: any == 0) and :435-481 (any == 1). The ray's maxt shrinks
 * on every accepted hit (core/primitive.cpp:167); ties go to the later-tested primitive. */
static int bvh_intersect(const SptSceneDesc *sc, Ray *ray, int any, uint32_t *slot_out, Hit *hit,
                         uint64_t *cnt_nodes, uint64_t *cnt_prims) {
    if (!sc->n_nodes) return 0;
    const Node *nodes = (const Node *)sc->bvh_nodes;
    int found = 0;
    v3 invDir = V(1.f / ray->d.x, 1.f / ray->d.y, 1.f / ray->d.z);
    uint32_t neg[3] = { invDir.x < 0, invDir.y < 0, invDir.z < 0 };
    uint32_t todoOffset = 0, nodeNum = 0;
    uint32_t todo[64];
    for (;;) {
        const Node *node = &nodes[nodeNum];
        if (cnt_nodes) ++*cnt_nodes;
        if (slab(node->b, ray, invDir, neg)) {
            if (node->nPrims > 0) {
                for (uint32_t i = 0; i < node->nPrims; ++i) {
                    uint32_t s = node->offset + i;
                    if (cnt_prims) ++*cnt_prims;
                    if (any) {
                        if (shape_intersect(sc, sc->prim_kind[s], sc->prim_flags[s], sc->prim_data[s], ray, NULL))
                            return 1;
                    } else {
                        Hit h;
                        if (shape_intersect(sc, sc->prim_kind[s], sc->prim_flags[s], sc->prim_data[s], ray, &h)) {
                            found = 1;
                            ray->maxt = h.t;
                            if (hit) *hit = h;
                            if (slot_out) *slot_out = s;
                        }
                    }
                }
                if (todoOffset == 0) break;
                nodeNum = todo[--todoOffset];
            } else {
                if (neg[node->axis]) { todo[todoOffset++] = nodeNum + 1; nodeNum = node->offset; }
                else { todo[todoOffset++] = node->offset; nodeNum = nodeNum + 1; }
            }
        } else {
            if (todoOffset == 0) break;
            nodeNum = todo[--todoOffset];
        }
    }
    return found;
}

static void load_ray(const float *r, Ray *ray) {
    ray->o = V(r[0], r[1], r[2]); ray->d = V(r[3], r[4], r[5]); ray->mint = r[6]; ray->maxt = r[7]; ray->depth = 0;
}

void orc_trace_closest(const SptSceneDesc *sc, const float *rays, uint64_t n,
                       uint32_t *out_slot, uint32_t *out_id, float *out_t) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        Ray ray; load_ray(rays + 8 * i, &ray);
        uint32_t slot = 0xffffffffu;
        int h = bvh_intersect(sc, &ray, 0, &slot, NULL, NULL, NULL);
        if (out_slot) out_slot[i] = h ? slot : 0xffffffffu;
        if (out_id) out_id[i] = h ? sc->prim_id[slot] : 0u;
        if (out_t) out_t[i] = ray.maxt;
    }
}

void orc_trace_closest_counted(const SptSceneDesc *sc, const float *rays, uint64_t n, uint64_t *nodes, uint64_t *prims) {
    uint64_t cn = 0, cp = 0;
    for (uint64_t i = 0; i < n; ++i) {
        Ray ray; load_ray(rays + 8 * i, &ray);
        bvh_intersect(sc, &ray, 0, NULL, NULL, &cn, &cp);
    }
    *nodes = cn; *prims = cp;
}

void orc_trace_any(const SptSceneDesc *sc, const float *rays, uint64_t n, uint8_t *out_hit) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        Ray ray; load_ray(rays + 8 * i, &ray);
        out_hit[i] = (uint8_t)bvh_intersect(sc, &ray, 1, NULL, NULL, NULL, NULL);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Monte Carlo helpers */
/* core/montecarlo.cpp:298-340 */
static void concentric_sample_disk(float u1, float u2, float *dx, float *dy) {
    float r, theta;
    float sx = 2 * u1 - 1;
    float sy = 2 * u2 - 1;
    if (sx == 0.0 && sy == 0.0) { *dx = 0.0; *dy = 0.0; return; }
    if (sx >= -sy) {
        if (sx > sy) {
            r = sx;
            if (sy > 0.0) theta = sy / r;
            else theta = 8.0f + sy / r;
        } else {
            r = sy;
            theta = 2.0f - sx / r;
        }
    } else {
        if (sx <= sy) { r = -sx; theta = 4.0f - sy / r; }
        else { r = -sy; theta = 6.0f + sx / r; }
    }
    theta *= PI_F / 4.f;
    *dx = r * cosf(theta);
    *dy = r * sinf(theta);
}
/* core/montecarlo.h:120-125 */
static v3 cosine_sample_hemisphere(float u1, float u2) {
    v3 ret;
    concentric_sample_disk(u1, u2, &ret.x, &ret.y);
    ret.z = sqrtf(stdmaxf(0.f, 1.f - ret.x * ret.x - ret.y * ret.y));
    return ret;
}
/* core/montecarlo.cpp:270-277 */
static v3 uniform_sample_sphere(float u1, float u2) {
    float z = 1.f - 2.f * u1;
    float r = sqrtf(stdmaxf(0.f, 1.f - z * z));
    float phi = 2.f * PI_F * u2;
    return V(r * cosf(phi), r * sinf(phi), z);
}
/* core/montecarlo.cpp:405-412 */
static v3 uniform_sample_cone(float u1, float u2, float costhetamax, v3 x, v3 y, v3 z) {
    float costheta = lerpf(u1, costhetamax, 1.f);
    float sintheta = sqrtf(1.f - costheta * costheta);
    float phi = u2 * 2.f * PI_F;
    return vadd(vadd(vmul(x, cosf(phi) * sintheta), vmul(y, sinf(phi) * sintheta)), vmul(z, costheta));
}
static float uniform_cone_pdf(float cosThetaMax) { return 1.f / (2.f * PI_F * (1.f - cosThetaMax)); }
/* core/montecarlo.cpp:342-347 */
static void uniform_sample_triangle(float u1, float u2, float *u, float *v) {
    float su1 = sqrtf(u1);
    *u = 1.f - su1;
    *v = u2 * su1;
}
/* core/montecarlo.h:254-257 */
static float power_heuristic(int nf, float fPdf, int ng, float gPdf) {
    float f = nf * fPdf, g = ng * gPdf;
    return (f * f) / (f * f + g * g);
}

/* ------------------------------------------------------------------------------------------ */
/* BSDF: core/reflection.cpp. A BSDF here is the frame (reflection.cpp:593-601) plus the material
 * row; its BxDF list is implied by the material type (materials/matte.cpp:34-60,
 * plastic.cpp:34-61, metal.cpp:44-68). */
enum { BX_LAMBERT = 0, BX_ORENNAYAR = 1, BX_MICROFACET_DIEL = 2, BX_MICROFACET_COND = 3,
       /* specular: reflection.cpp:130-160 */
       BX_SPEC_REFL_NOOP = 4, BX_SPEC_REFL_DIEL = 5, BX_SPEC_TRANS = 6,
       /* FresnelBlend over an Anisotropic distribution (substrate): reflection.cpp:217-236,369-459 */
       BX_FRESNEL_BLEND = 7,
       /* IrregIsotropicBRDF (measured material): reflection.cpp:239-263 */
       BX_MEASURED = 8 };
#define BX_IS_SPECULAR(k) ((k) >= BX_SPEC_REFL_NOOP && (k) <= BX_SPEC_TRANS)
typedef struct {
    v3 nn, sn, tn, ng;
    int nBxDFs;
    int kind[2];
    const float *R[2];          /* reflectance spectrum of each BxDF (NULL: 1) */
    const float *eta, *k;       /* conductor */
    float exponent;             /* Blinn */
    float A, B;                 /* Oren-Nayar */
    float ior;                  /* glass / subsurface: FresnelDielectric(1, ior) */
    float ex, ey;               /* Anisotropic */
    float Rtex[NB];             /* Kd evaluated from an image texture at this hit (R[0] points here) */
    const SptSceneDesc *sc;     /* measured BRDF: the tables */
    int brdf;
} BSDF;

static v3 w2l(const BSDF *b, v3 v) { return V(dot(v, b->sn), dot(v, b->tn), dot(v, b->nn)); }
static v3 l2w(const BSDF *b, v3 v) {
    return V(b->sn.x * v.x + b->tn.x * v.y + b->nn.x * v.z,
             b->sn.y * v.x + b->tn.y * v.y + b->nn.y * v.z,
             b->sn.z * v.x + b->tn.z * v.y + b->nn.z * v.z);
}
static inline float abs_cos_theta(v3 w) { return fabsf(w.z); }
static inline int same_hemisphere(v3 w, v3 wp) { return w.z * wp.z > 0.f; }
static inline float sin_theta2(v3 w) { return stdmaxf(0.f, 1.f - w.z * w.z); }
static inline float sin_theta(v3 w) { return sqrtf(sin_theta2(w)); }
static inline float cos_phi(v3 w) { float s = sin_theta(w); if (s == 0.f) return 1.f; return clampf(w.x / s, -1.f, 1.f); }
static inline float sin_phi(v3 w) { float s = sin_theta(w); if (s == 0.f) return 0.f; return clampf(w.y / s, -1.f, 1.f); }

static int is_black(const float *s);
static float blinn_exponent(float e) { if (e > 10000.f || isnan(e)) e = 10000.f; return e; }   /* reflection.h:416-417 */


/* ------------------------------------------------------------------------------------------ */
/* Image textures: DifferentialGeometry::ComputeDifferentials, UVMapping2D::Map,
 * MIPMap::Lookup (EWA / trilinear / the fork's noFiltering), FromRGB(SPECTRUM_REFLECTANCE). */
typedef struct { float dudx, dvdx, dudy, dvdy; } UVDiff;

static int solve2x2(const float A[2][2], const float B[2], float *x0, float *x1) {   /* core/transform.cpp:31-41 */
    float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    if (fabsf(det) < 1e-10f) return 0;
    *x0 = (A[1][1] * B[0] - A[0][1] * B[1]) / det;
    *x1 = (A[0][0] * B[1] - A[1][0] * B[0]) / det;
    if (isnan(*x0) || isnan(*x1)) return 0;
    return 1;
}
static inline float vcomp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* core/diffgeom.cpp:50-107 */
static void compute_differentials(const Hit *dg, const RayDiff *rd, UVDiff *o) {
    o->dudx = o->dvdx = o->dudy = o->dvdy = 0.f;
    if (!rd || !rd->has) return;
    v3 nn = dg->nn, p = dg->p;
    float d = -dot(nn, p);
    float tx = -(dot(nn, rd->rxo) + d) / dot(nn, rd->rxd);
    if (isnan(tx)) return;
    v3 px = vadd(rd->rxo, vmul(rd->rxd, tx));
    float ty = -(dot(nn, rd->ryo) + d) / dot(nn, rd->ryd);
    if (isnan(ty)) return;
    v3 py = vadd(rd->ryo, vmul(rd->ryd, ty));
    int axes[2];
    if (fabsf(nn.x) > fabsf(nn.y) && fabsf(nn.x) > fabsf(nn.z)) { axes[0] = 1; axes[1] = 2; }
    else if (fabsf(nn.y) > fabsf(nn.z)) { axes[0] = 0; axes[1] = 2; }
    else { axes[0] = 0; axes[1] = 1; }
    float A[2][2], Bx[2], By[2];
    A[0][0] = vcomp(dg->dpdu, axes[0]); A[0][1] = vcomp(dg->dpdv, axes[0]);
    A[1][0] = vcomp(dg->dpdu, axes[1]); A[1][1] = vcomp(dg->dpdv, axes[1]);
    Bx[0] = vcomp(px, axes[0]) - vcomp(p, axes[0]); Bx[1] = vcomp(px, axes[1]) - vcomp(p, axes[1]);
    By[0] = vcomp(py, axes[0]) - vcomp(p, axes[0]); By[1] = vcomp(py, axes[1]) - vcomp(p, axes[1]);
    if (!solve2x2(A, Bx, &o->dudx, &o->dvdx)) { o->dudx = 0.f; o->dvdx = 0.f; }
    if (!solve2x2(A, By, &o->dudy, &o->dvdy)) { o->dudy = 0.f; o->dvdy = 0.f; }
}

static int imod(int a, int b);
static inline float log2_pbrt(float x) { float invLog2 = 1.f / logf(2.f); return logf(x) * invLog2; }   /* core/pbrt.h:243-246 */

typedef struct { const float *texels; int w, h; } TexLevel;
static TexLevel tex_level(const SptSceneDesc *sc, const SptTexture *t, int level) {
    TexLevel l; l.texels = sc->tex_texels + t->texel_offset; l.w = t->width; l.h = t->height;
    for (int i = 0; i < level; ++i) {
        l.texels += (size_t)l.w * l.h * t->channels;
        l.w = l.w / 2 > 1 ? l.w / 2 : 1; l.h = l.h / 2 > 1 ? l.h / 2 : 1;
    }
    return l;
}
/* MIPMap::Texel, core/mipmap.h:177-201 */
static void tex_texel(const SptSceneDesc *sc, const SptTexture *t, int level, int s, int tt, float *out) {
    TexLevel l = tex_level(sc, t, level);
    if (t->wrap == SPT_WRAP_REPEAT) { s = imod(s, l.w); tt = imod(tt, l.h); }
    else if (t->wrap == SPT_WRAP_CLAMP) { s = clampi(s, 0, l.w - 1); tt = clampi(tt, 0, l.h - 1); }
    else if (s < 0 || s >= l.w || tt < 0 || tt >= l.h) { for (int k = 0; k < t->channels; ++k) out[k] = 0.f; return; }
    const float *px = l.texels + ((size_t)tt * l.w + s) * t->channels;
    for (int k = 0; k < t->channels; ++k) out[k] = px[k];
}
/* MIPMap::triangle, core/mipmap.h:258-270 */
static void tex_triangle(const SptSceneDesc *sc, const SptTexture *t, int level, float s, float tt, float *out) {
    level = clampi(level, 0, t->n_levels - 1);
    TexLevel l = tex_level(sc, t, level);
    s = s * l.w - 0.5f;
    tt = tt * l.h - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(tt);
    float ds = s - s0, dt = tt - t0;
    float a[3], b[3], c[3], d[3];
    tex_texel(sc, t, level, s0, t0, a); tex_texel(sc, t, level, s0, t0 + 1, b);
    tex_texel(sc, t, level, s0 + 1, t0, c); tex_texel(sc, t, level, s0 + 1, t0 + 1, d);
    for (int k = 0; k < t->channels; ++k)
        out[k] = a[k] * ((1.f - ds) * (1.f - dt)) + b[k] * ((1.f - ds) * dt) + c[k] * (ds * (1.f - dt)) + d[k] * (ds * dt);
}
/* MIPMap::EWA, core/mipmap.h:322-377 */
static void tex_ewa(const SptSceneDesc *sc, const SptTexture *t, int level, float s, float tt,
                    float ds0, float dt0, float ds1, float dt1, float *out) {
    if (level >= t->n_levels) { tex_texel(sc, t, t->n_levels - 1, 0, 0, out); return; }
    TexLevel l = tex_level(sc, t, level);
    s = s * l.w - 0.5f;
    tt = tt * l.h - 0.5f;
    ds0 *= l.w; dt0 *= l.h; ds1 *= l.w; dt1 *= l.h;
    float A = dt0 * dt0 + dt1 * dt1 + 1;
    float B = -2.f * (ds0 * dt0 + ds1 * dt1);
    float C = ds0 * ds0 + ds1 * ds1 + 1;
    float invF = 1.f / (A * C - B * B * 0.25f);
    A *= invF; B *= invF; C *= invF;
    float det = -B * B + 4.f * A * C;
    float invDet = 1.f / det;
    float uSqrt = sqrtf(det * C), vSqrt = sqrtf(A * det);
    int s0 = (int)ceilf(s - 2.f * invDet * uSqrt), s1 = (int)floorf(s + 2.f * invDet * uSqrt);
    int t0 = (int)ceilf(tt - 2.f * invDet * vSqrt), t1 = (int)floorf(tt + 2.f * invDet * vSqrt);
    float sum[3] = { 0.f, 0.f, 0.f }, sumWts = 0.f;
    for (int it = t0; it <= t1; ++it) {
        float ttt = it - tt;
        for (int is = s0; is <= s1; ++is) {
            float ss = is - s;
            float r2 = A * ss * ss + B * ss * ttt + C * ttt * ttt;
            if (r2 < 1.) {
                int li = (int)(r2 * 128);
                float weight = sc->ewa_weight_lut[li < 127 ? li : 127];
                float tx[3];
                tex_texel(sc, t, level, is, it, tx);
                for (int k = 0; k < t->channels; ++k) sum[k] += tx[k] * weight;
                sumWts += weight;
            }
        }
    }
    for (int k = 0; k < t->channels; ++k) out[k] = sum[k] / sumWts;
}
/* MIPMap::Lookup(s,t,width), core/mipmap.h:215-255 */
static void tex_lookup_tri(const SptSceneDesc *sc, const SptTexture *t, float s, float tt, float width, float *out) {
    if (t->no_filter) {
        s = s * t->width - 0.5f;
        tt = tt * t->height - 0.5f;
        tex_texel(sc, t, 0, (int)floorf(s + 0.5f), (int)floorf(tt + 0.5f), out);
        return;
    }
    float level = t->n_levels - 1 + log2_pbrt(stdmaxf(width, 1e-8f));
    if (level < 0) tex_triangle(sc, t, 0, s, tt, out);
    else if (level >= t->n_levels - 1) tex_texel(sc, t, t->n_levels - 1, 0, 0, out);
    else {
        int iLevel = (int)floorf(level);
        float delta = level - iLevel;
        float a[3], b[3];
        tex_triangle(sc, t, iLevel, s, tt, a);
        tex_triangle(sc, t, iLevel + 1, s, tt, b);
        for (int k = 0; k < t->channels; ++k) out[k] = a[k] * (1.f - delta) + b[k] * delta;
    }
}
/* MIPMap::Lookup(s,t,ds0,dt0,ds1,dt1), core/mipmap.h:273-319 */
static void tex_lookup(const SptSceneDesc *sc, const SptTexture *t, float s, float tt,
                       float ds0, float dt0, float ds1, float dt1, float *out) {
    if (t->trilinear || t->no_filter) {
        tex_lookup_tri(sc, t, s, tt, 2.f * stdmaxf(stdmaxf(fabsf(ds0), fabsf(dt0)), stdmaxf(fabsf(ds1), fabsf(dt1))), out);
        return;
    }
    if (ds0 * ds0 + dt0 * dt0 < ds1 * ds1 + dt1 * dt1) {
        float tmp = ds0; ds0 = ds1; ds1 = tmp;
        tmp = dt0; dt0 = dt1; dt1 = tmp;
    }
    float majorLength = sqrtf(ds0 * ds0 + dt0 * dt0);
    float minorLength = sqrtf(ds1 * ds1 + dt1 * dt1);
    if (minorLength * t->max_aniso < majorLength && minorLength > 0.f) {
        float scale = majorLength / (minorLength * t->max_aniso);
        ds1 *= scale; dt1 *= scale; minorLength *= scale;
    }
    if (minorLength == 0.f) { tex_triangle(sc, t, 0, s, tt, out); return; }
    float lod = stdmaxf(0.f, t->n_levels - 1.f + log2_pbrt(minorLength));
    int ilod = (int)floorf(lod);
    float d = lod - ilod;
    float a[3], b[3];
    tex_ewa(sc, t, ilod, s, tt, ds0, dt0, ds1, dt1, a);
    tex_ewa(sc, t, ilod + 1, s, tt, ds0, dt0, ds1, dt1, b);
    for (int k = 0; k < t->channels; ++k) out[k] = a[k] * (1.f - d) + b[k] * d;
}
/* ImageTexture::Evaluate (textures/imagemap.cpp:88-97) over UVMapping2D::Map (core/texture.cpp:80-90); a float
 * image is then multiplied by the constant of an enclosing ScaleTexture (textures/scale.h:45-47) */
static void tex_evaluate(const SptSceneDesc *sc, const SptTexture *t, float u, float v, const UVDiff *df, float *out) {
    float s = t->su * u + t->du, tt = t->sv * v + t->dv;
    float dsdx = t->su * df->dudx, dtdx = t->sv * df->dvdx, dsdy = t->su * df->dudy, dtdy = t->sv * df->dvdy;
    tex_lookup(sc, t, s, tt, dsdx, dtdx, dsdy, dtdy, out);
    if (t->channels == 1) out[0] = out[0] * t->scale;
}
/* SampledSpectrum::FromRGB(rgb, SPECTRUM_REFLECTANCE), core/spectrum.cpp:92-133,175 */
static void from_rgb_refl(const SptSpectralTables *t, const float rgb[3], float *r) {
    enum { W = 0, Cy = 1, Mg = 2, Ye = 3, Rd = 4, Gr = 5, Bl = 6 };
    for (int c = 0; c < NB; ++c) r[c] = 0.f;
#define ADDR(coef, basis) do { float k_ = (coef); for (int c = 0; c < NB; ++c) r[c] += t->rgb_refl[basis][c] * k_; } while (0)
    if (rgb[0] <= rgb[1] && rgb[0] <= rgb[2]) {
        ADDR(rgb[0], W);
        if (rgb[1] <= rgb[2]) { ADDR(rgb[1] - rgb[0], Cy); ADDR(rgb[2] - rgb[1], Bl); }
        else { ADDR(rgb[2] - rgb[0], Cy); ADDR(rgb[1] - rgb[2], Gr); }
    } else if (rgb[1] <= rgb[0] && rgb[1] <= rgb[2]) {
        ADDR(rgb[1], W);
        if (rgb[0] <= rgb[2]) { ADDR(rgb[0] - rgb[1], Mg); ADDR(rgb[2] - rgb[0], Bl); }
        else { ADDR(rgb[2] - rgb[1], Mg); ADDR(rgb[0] - rgb[2], Rd); }
    } else {
        ADDR(rgb[2], W);
        if (rgb[0] <= rgb[1]) { ADDR(rgb[0] - rgb[2], Ye); ADDR(rgb[1] - rgb[0], Gr); }
        else { ADDR(rgb[1] - rgb[2], Ye); ADDR(rgb[0] - rgb[1], Rd); }
    }
#undef ADDR
    for (int c = 0; c < NB; ++c) { r[c] *= .94f; r[c] = clampf(r[c], 0.f, INFINITY); }
}

/* Intersection::GetBSDF (core/intersection.cpp:39-47): ComputeDifferentials, GetShadingGeometry, then
 * Material::Bump (core/material.cpp:39-82; the constant-0 displacement of SURVEY.md F6, or a float image map) and
 * the material's GetBSDF. rd: the camera ray's offset rays at the first vertex, NULL afterwards (path.cpp:93). */
static void make_bsdf(const SptSceneDesc *sc, uint32_t slot, const Hit *dg, const RayDiff *rd, BSDF *b, v3 *n_shading) {
    int flags = sc->prim_flags[slot];
    const SptMaterial *m = sc->materials + sc->prim_material[slot];
    const int textured = m->tex_kd >= 0 || m->tex_bump >= 0;
    UVDiff df = { 0.f, 0.f, 0.f, 0.f };
    if (textured) compute_differentials(dg, rd, &df);
    v3 s_dpdu = dg->dpdu, s_dpdv = dg->dpdv;      /* dgShading */
    v3 s_nn = dg->nn, dndu = V(0, 0, 0), dndv = V(0, 0, 0);
    if (sc->prim_kind[slot] == SPT_PRIM_TRIANGLE && (flags & SPT_PF_HAS_N)) {
        /* Triangle::GetShadingGeometry, shapes/trianglemesh.cpp:285-360 */
        const int32_t *vi = sc->tri_vidx + 3 * (size_t)sc->prim_data[slot];
        float uv[3][2];
        tri_uvs(sc, flags, vi, uv);
        float A[2][2] = { { uv[1][0] - uv[0][0], uv[2][0] - uv[0][0] }, { uv[1][1] - uv[0][1], uv[2][1] - uv[0][1] } };
        float C[2] = { dg->u - uv[0][0], dg->v - uv[0][1] };
        float bb[3];
        /* SolveLinearSystem2x2, core/transform.cpp:31-41 */
        float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
        int ok = 1;
        if (fabsf(det) < 1e-10f) ok = 0;
        else {
            bb[1] = (A[1][1] * C[0] - A[0][1] * C[1]) / det;
            bb[2] = (A[0][0] * C[1] - A[1][0] * C[0]) / det;
            if (isnan(bb[1]) || isnan(bb[2])) ok = 0;
        }
        if (!ok) bb[0] = bb[1] = bb[2] = 1.f / 3.f;
        else bb[0] = 1.f - bb[1] - bb[2];
        const SptXform *xf = sc->xforms + sc->prim_xform[slot];
        const float *n0 = sc->N + 3 * (size_t)vi[0], *n1 = sc->N + 3 * (size_t)vi[1], *n2 = sc->N + 3 * (size_t)vi[2];
        v3 nsum = vadd(vadd(vmul(V(n0[0], n0[1], n0[2]), bb[0]), vmul(V(n1[0], n1[1], n1[2]), bb[1])),
                       vmul(V(n2[0], n2[1], n2[2]), bb[2]));
        v3 ns = normalize(xf_normal(xf->minv, nsum));
        v3 ss = normalize(dg->dpdu);
        v3 ts = cross(ss, ns);
        if (len2(ts) > 0.f) { ts = normalize(ts); ss = cross(ts, ns); }
        else coordinate_system(ns, &ss, &ts);
        s_dpdu = ss; s_dpdv = ts;
        if (m->tex_bump >= 0) {
            /* dndu, dndv (trianglemesh.cpp:331-351) and the shading DifferentialGeometry's own normal (diffgeom.cpp:32-47) */
            float du1 = uv[0][0] - uv[2][0], du2 = uv[1][0] - uv[2][0];
            float dv1 = uv[0][1] - uv[2][1], dv2 = uv[1][1] - uv[2][1];
            v3 N0 = V(n0[0], n0[1], n0[2]), N1 = V(n1[0], n1[1], n1[2]), N2 = V(n2[0], n2[1], n2[2]);
            v3 dn1 = vsub(N0, N2), dn2 = vsub(N1, N2);
            float determinant = du1 * dv2 - dv1 * du2;
            if (determinant != 0.f) {
                float invdet = 1.f / determinant;
                dndu = vmul(vsub(vmul(dn1, dv2), vmul(dn2, dv1)), invdet);
                dndv = vmul(vadd(vmul(dn1, -du2), vmul(dn2, du1)), invdet);
            }
            dndu = xf_normal(xf->minv, dndu);
            dndv = xf_normal(xf->minv, dndv);
            s_nn = normalize(cross(ss, ts));
            if (flags & SPT_PF_FLIP_NORMAL) s_nn = vmul(s_nn, -1.f);
        }
    }
    if (m->tex_bump >= 0) {                                    /* Material::Bump, core/material.cpp:39-82 */
        const SptTexture *bt = sc->textures + m->tex_bump;
        float du = .5f * (fabsf(df.dudx) + fabsf(df.dudy));
        if (du == 0.f) du = .01f;
        float uDisplace, vDisplace, displace;
        tex_evaluate(sc, bt, dg->u + du, dg->v, &df, &uDisplace);
        float dv = .5f * (fabsf(df.dvdx) + fabsf(df.dvdy));
        if (dv == 0.f) dv = .01f;
        tex_evaluate(sc, bt, dg->u, dg->v + dv, &df, &vDisplace);
        tex_evaluate(sc, bt, dg->u, dg->v, &df, &displace);
        s_dpdu = vadd(vadd(s_dpdu, vmul(s_nn, (uDisplace - displace) / du)), vmul(dndu, displace));
        s_dpdv = vadd(vadd(s_dpdv, vmul(s_nn, (vDisplace - displace) / dv)), vmul(dndv, displace));
    }
    /* Bump with a constant-0 displacement leaves dpdu/dpdv unchanged and recomputes the normal */
    v3 nn = normalize(cross(s_dpdu, s_dpdv));
    if (flags & SPT_PF_FLIP_NORMAL) nn = vmul(nn, -1.f);
    if (dot(nn, dg->nn) < 0.f) nn = vneg(nn);                  /* Faceforward, core/geometry.h:596-598 */
    b->nn = nn;
    b->ng = dg->nn;
    b->sn = normalize(s_dpdu);
    b->tn = cross(b->nn, b->sn);
    *n_shading = nn;
    b->eta = b->k = NULL; b->exponent = 0.f; b->A = b->B = 0.f; b->ex = b->ey = 0.f;
    const float *kd = m->spec0;
    if (m->tex_kd >= 0) {                                      /* Kd->Evaluate(dgs).Clamp(): imagemap.cpp:88-97, imagemap.h:88-92 */
        float rgb[3];
        tex_evaluate(sc, sc->textures + m->tex_kd, dg->u, dg->v, &df, rgb);
        from_rgb_refl(&sc->tables, rgb, b->Rtex);
        kd = b->Rtex;
    }
    b->sc = sc; b->brdf = -1;
    if (m->type == SPT_MAT_MEASURED) {                         /* materials/measured.cpp:185-205 */
        b->nBxDFs = 1; b->kind[0] = BX_MEASURED; b->R[0] = NULL; b->brdf = m->brdf;
    } else if (m->type == SPT_MAT_SUBSTRATE) {                 /* materials/substrate.cpp:34-56, reflection.h:433-437 */
        b->nBxDFs = 1;
        b->kind[0] = BX_FRESNEL_BLEND; b->R[0] = kd; b->R[1] = m->spec1;
        b->ex = 1.f / m->p0; b->ey = 1.f / m->p1;
        if (b->ex > 10000.f || isnan(b->ex)) b->ex = 10000.f;
        if (b->ey > 10000.f || isnan(b->ey)) b->ey = 10000.f;
    } else if (m->type == SPT_MAT_MATTE) {
        b->nBxDFs = 1; b->R[0] = kd;
        if (m->p0 == 0.) b->kind[0] = BX_LAMBERT;
        else {
            b->kind[0] = BX_ORENNAYAR;                         /* reflection.h:363-370 */
            float sigma = (PI_F / 180.f) * m->p0;
            float sigma2 = sigma * sigma;
            b->A = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
            b->B = 0.45f * sigma2 / (sigma2 + 0.09f);
        }
    } else if (m->type == SPT_MAT_PLASTIC) {
        b->nBxDFs = 2;
        b->kind[0] = BX_LAMBERT; b->R[0] = kd;
        b->kind[1] = BX_MICROFACET_DIEL; b->R[1] = m->spec1;
        b->exponent = blinn_exponent(1.f / m->p0);
    } else if (m->type == SPT_MAT_METAL) {
        b->nBxDFs = 1;
        b->kind[0] = BX_MICROFACET_COND; b->R[0] = NULL;
        b->eta = m->spec0; b->k = m->spec1;
        b->exponent = blinn_exponent(1.f / m->p0);
    } else if (m->type == SPT_MAT_MIRROR) {                     /* materials/mirror.cpp:34-55 */
        b->nBxDFs = 0;
        if (!is_black(m->spec0)) { b->kind[0] = BX_SPEC_REFL_NOOP; b->R[0] = m->spec0; b->nBxDFs = 1; }
    } else {                                                   /* materials/glass.cpp:34-58; subsurface.cpp:40-58 = reflection only */
        b->nBxDFs = 0;
        b->ior = m->p0;
        if (!is_black(m->spec0)) { b->kind[b->nBxDFs] = BX_SPEC_REFL_DIEL; b->R[b->nBxDFs] = m->spec0; b->nBxDFs++; }
        if (!is_black(m->spec1)) { b->kind[b->nBxDFs] = BX_SPEC_TRANS; b->R[b->nBxDFs] = m->spec1; b->nBxDFs++; }
    }
}

/* FresnelDielectric(1.5, 1)::Evaluate, reflection.cpp:107-127 + FrDiel :52-60 (all bands equal) */
static float fresnel_dielectric(float cosi, float eta_i, float eta_t) {
    cosi = clampf(cosi, -1.f, 1.f);
    int entering = cosi > 0.;
    float ei = eta_i, et = eta_t;
    if (!entering) { float tmp = ei; ei = et; et = tmp; }
    float sint = ei / et * sqrtf(stdmaxf(0.f, 1.f - cosi * cosi));
    if (sint >= 1.) return 1.f;
    float cost = sqrtf(stdmaxf(0.f, 1.f - sint * sint));
    float ac = fabsf(cosi);
    float Rparl = ((et * ac) - (ei * cost)) / ((et * ac) + (ei * cost));
    float Rperp = ((ei * ac) - (et * cost)) / ((ei * ac) + (et * cost));
    return (Rparl * Rparl + Rperp * Rperp) / 2.f;
}
/* FrCond, reflection.cpp:63-71 */
static float fr_cond(float cosi, float eta, float k) {
    float tmp = (eta * eta + k * k) * cosi * cosi;
    float Rparl2 = (tmp - (2.f * eta * cosi) + 1) / (tmp + (2.f * eta * cosi) + 1);
    float tmp_f = eta * eta + k * k;
    float Rperp2 = (tmp_f - (2.f * eta * cosi) + cosi * cosi) / (tmp_f + (2.f * eta * cosi) + cosi * cosi);
    return (Rparl2 + Rperp2) / 2.f;
}

/* KdTree::privateLookup (core/kdtree.h:143-168) with IrregIsoProc (reflection.cpp:34-47): children first, then the node */
typedef struct { float v[NB]; float sumWeights; int nFound; } IrregIsoProc;
static void kd_lookup(const SptKdNode *nodes, const float *spectra, uint32_t nNodes, uint32_t nodeNum, const float p[3],
                      IrregIsoProc *proc, float maxDistSquared) {
    const SptKdNode *node = nodes + nodeNum;
    int axis = (int)(node->bits & 3u);
    int hasLeft = (int)((node->bits >> 2) & 1u);
    uint32_t rightChild = node->bits >> 3;
    if (axis != 3) {
        float dist2 = (p[axis] - node->split_pos) * (p[axis] - node->split_pos);
        if (p[axis] <= node->split_pos) {
            if (hasLeft) kd_lookup(nodes, spectra, nNodes, nodeNum + 1, p, proc, maxDistSquared);
            if (dist2 < maxDistSquared && rightChild < nNodes) kd_lookup(nodes, spectra, nNodes, rightChild, p, proc, maxDistSquared);
        } else {
            if (rightChild < nNodes) kd_lookup(nodes, spectra, nNodes, rightChild, p, proc, maxDistSquared);
            if (dist2 < maxDistSquared && hasLeft) kd_lookup(nodes, spectra, nNodes, nodeNum + 1, p, proc, maxDistSquared);
        }
    }
    float dx = node->p[0] - p[0], dy = node->p[1] - p[1], dz = node->p[2] - p[2];
    float d2 = dx * dx + dy * dy + dz * dz;                   /* DistanceSquared -> Vector::LengthSquared, geometry.h */
    if (d2 < maxDistSquared) {
        float weight = expf(-100.f * d2);
        const float *sv = spectra + (size_t)nodeNum * SPT_BAND_PITCH;
        for (int c = 0; c < NB; ++c) proc->v[c] += sv[c] * weight;
        proc->sumWeights += weight;
        ++proc->nFound;
    }
}
static float spherical_phi(v3 v);
static void from_rgb_refl(const SptSpectralTables *t, const float rgb[3], float *r);
/* RegularHalfangleBRDF::f (reflection.cpp:267-300), added into out[NB]: M_PI is the float constant of pbrt.h */
static void halfangle_f(const SptSceneDesc *sc, const SptBrdfTable *t, v3 wo, v3 wi, float *out) {
    v3 wh = vadd(wo, wi);
    if (wh.z < 0.f) { wo = vneg(wo); wi = vneg(wi); wh = vneg(wh); }
    if (wh.x == 0.f && wh.y == 0.f && wh.z == 0.f) return;
    wh = normalize(wh);
    float whTheta = acosf(clampf(wh.z, -1.f, 1.f));
    float whCosPhi = cos_phi(wh), whSinPhi = sin_phi(wh);
    float whCosTheta = wh.z, whSinTheta = sin_theta(wh);
    v3 whx = V(whCosPhi * whCosTheta, whSinPhi * whCosTheta, -whSinTheta);
    v3 why = V(-whSinPhi, whCosPhi, 0.f);
    v3 wd = V(dot(wi, whx), dot(wi, why), dot(wi, wh));
    float wdTheta = acosf(clampf(wd.z, -1.f, 1.f)), wdPhi = spherical_phi(wd);
    if (wdPhi > PI_F) wdPhi -= PI_F;
    int nH = (int)t->n_theta_h, nD = (int)t->n_theta_d, nP = (int)t->n_phi_d;
#define REMAP(V_, MAX_, COUNT_) clampi((int)((V_) / (MAX_) * (COUNT_)), 0, (COUNT_) - 1)
    int ih = REMAP(sqrtf(stdmaxf(0.f, whTheta / (PI_F / 2.f))), 1.f, nH);
    int id = REMAP(wdTheta, PI_F / 2.f, nD);
    int ip = REMAP(wdPhi, PI_F, nP);
#undef REMAP
    const float *e = sc->merl_rgb + t->rgb_offset + 3 * (size_t)(ip + nP * (id + ih * nD));
    float s[NB];
    from_rgb_refl(&sc->tables, e, s);
    for (int c = 0; c < NB; ++c) out[c] += s[c];
}
/* IrregIsotropicBRDF::f (reflection.cpp:251-263) over BRDFRemap (:239-248), added into out[NB] */
static void measured_f(const SptSceneDesc *sc, const SptBrdfTable *t, v3 wo, v3 wi, float *out) {
    if (t->n_nodes == 0) { halfangle_f(sc, t, wo, wi, out); return; }
    float cosi = wi.z, coso = wo.z;
    float sini = sqrtf(stdmaxf(0.f, 1.f - wi.z * wi.z)), sino = sqrtf(stdmaxf(0.f, 1.f - wo.z * wo.z));
    float phii = spherical_phi(wi), phio = spherical_phi(wo);
    float dphi = phii - phio;
    if (dphi < 0.) dphi += 2.f * PI_F;
    if (dphi > 2.f * PI_F) dphi -= 2.f * PI_F;
    if (dphi > PI_F) dphi = 2.f * PI_F - dphi;
    float m[3] = { sini * sino, dphi / PI_F, cosi * coso };
    float lastMaxDist2 = .001f;
    for (;;) {
        IrregIsoProc proc;
        for (int c = 0; c < NB; ++c) proc.v[c] = 0.f;
        proc.sumWeights = 0.f; proc.nFound = 0;
        kd_lookup(sc->brdf_nodes + t->node_first, sc->brdf_spectra + (size_t)t->node_first * SPT_BAND_PITCH, t->n_nodes, 0, m, &proc, lastMaxDist2);
        if (proc.nFound > 2 || lastMaxDist2 > 1.5f) {
            for (int c = 0; c < NB; ++c) out[c] += clampf(proc.v[c], 0.f, INFINITY) / proc.sumWeights;
            return;
        }
        lastMaxDist2 *= 2.f;
    }
}

/* f of one BxDF in local coordinates, added into out[NB] */
static void bxdf_f(const BSDF *b, int i, v3 wo, v3 wi, float *out) {
    int kind = b->kind[i];
    if (kind == BX_LAMBERT) {                                  /* reflection.cpp:165-167 */
        for (int c = 0; c < NB; ++c) out[c] += b->R[i][c] * INV_PI_F;
        return;
    }
    if (kind == BX_ORENNAYAR) {                                /* reflection.cpp:170-193 */
        float sinthetai = sin_theta(wi), sinthetao = sin_theta(wo);
        float maxcos = 0.f;
        if (sinthetai > 1e-4 && sinthetao > 1e-4) {
            float sinphii = sin_phi(wi), cosphii = cos_phi(wi);
            float sinphio = sin_phi(wo), cosphio = cos_phi(wo);
            float dcos = cosphii * cosphio + sinphii * sinphio;
            maxcos = stdmaxf(0.f, dcos);
        }
        float sinalpha, tanbeta;
        if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sinalpha = sinthetao; tanbeta = sinthetai / abs_cos_theta(wi); }
        else { sinalpha = sinthetai; tanbeta = sinthetao / abs_cos_theta(wo); }
        float s = (b->A + b->B * maxcos * sinalpha * tanbeta);
        for (int c = 0; c < NB; ++c) out[c] += b->R[i][c] * INV_PI_F * s;
        return;
    }
    if (kind == BX_MEASURED) { measured_f(b->sc, b->sc->brdfs + b->brdf, wo, wi, out); return; }
    if (kind == BX_FRESNEL_BLEND) {                            /* reflection.cpp:224-236; Anisotropic::D reflection.h:438-444 */
        const float *Rd = b->R[0], *Rs = b->R[1];
        float k0 = (28.f / (23.f * PI_F));
        float a1 = (1.f - powf(1.f - .5f * abs_cos_theta(wi), 5));
        float a2 = (1.f - powf(1.f - .5f * abs_cos_theta(wo), 5));
        v3 wh = vadd(wi, wo);
        if (wh.x == 0. && wh.y == 0. && wh.z == 0.) return;
        wh = normalize(wh);
        float costhetah = abs_cos_theta(wh);
        float dd = 1.f - costhetah * costhetah;
        float D = 0.f;
        if (dd != 0.f) {
            float e = (b->ex * wh.x * wh.x + b->ey * wh.y * wh.y) / dd;
            D = sqrtf((b->ex + 2.f) * (b->ey + 2.f)) * INV_TWOPI_F * powf(costhetah, e);
        }
        float sc_ = D / (4.f * absdot(wi, wh) * stdmaxf(abs_cos_theta(wi), abs_cos_theta(wo)));
        float p5 = powf(1 - dot(wi, wh), 5.f);
        for (int c = 0; c < NB; ++c) {
            float diffuse = Rd[c] * k0 * (1.f - Rs[c]) * a1 * a2;
            float specular = (Rs[c] + (1.f - Rs[c]) * p5) * sc_;
            out[c] += diffuse + specular;
        }
        return;
    }
    /* Microfacet::f, reflection.cpp:203-214; G reflection.h:395-402; Blinn::D reflection.h:419-422 */
    float cosThetaO = abs_cos_theta(wo), cosThetaI = abs_cos_theta(wi);
    if (cosThetaI == 0.f || cosThetaO == 0.f) return;
    v3 wh = vadd(wi, wo);
    if (wh.x == 0. && wh.y == 0. && wh.z == 0.) return;
    wh = normalize(wh);
    float cosThetaH = dot(wi, wh);
    float D = (b->exponent + 2) * INV_TWOPI_F * powf(abs_cos_theta(wh), b->exponent);
    float NdotWh = abs_cos_theta(wh), NdotWo = abs_cos_theta(wo), NdotWi = abs_cos_theta(wi);
    float WOdotWh = absdot(wo, wh);
    float G = stdminf(1.f, stdminf((2.f * NdotWh * NdotWo / WOdotWh), (2.f * NdotWh * NdotWi / WOdotWh)));
    float denom = (4.f * cosThetaI * cosThetaO);
    if (kind == BX_MICROFACET_DIEL) {
        float F = fresnel_dielectric(cosThetaH, 1.5f, 1.f);
        for (int c = 0; c < NB; ++c) out[c] += b->R[i][c] * D * G * F / denom;
    } else {
        float ac = fabsf(cosThetaH);
        for (int c = 0; c < NB; ++c) out[c] += 1.f * D * G * fr_cond(ac, b->eta[c], b->k[c]) / denom;
    }
}

static int bxdf_is_reflection(const BSDF *b, int i) { (void)b; (void)i; return 1; }   /* all lowered BxDFs are BRDFs */

/* Anisotropic::Pdf, reflection.cpp:420-432 (and the tail of Anisotropic::Sample_f :396-405) for a half vector */
static float anisotropic_pdf(const BSDF *b, v3 wo, v3 wh) {
    float costhetah = abs_cos_theta(wh);
    float ds = 1.f - costhetah * costhetah;
    float pdf = 0.f;
    if (ds > 0.f && dot(wo, wh) > 0.f) {
        float e = (b->ex * wh.x * wh.x + b->ey * wh.y * wh.y) / ds;
        float d = sqrtf((b->ex + 1.f) * (b->ey + 1.f)) * INV_TWOPI_F * powf(costhetah, e);
        pdf = d / (4.f * dot(wo, wh));
    }
    return pdf;
}
/* Anisotropic::sampleFirstQuadrant, reflection.cpp:408-417 */
static void aniso_first_quadrant(const BSDF *b, float u1, float u2, float *phi, float *costheta) {
    if (b->ex == b->ey) *phi = PI_F * u1 * 0.5f;
    else *phi = atanf(sqrtf((b->ex + 1.f) / (b->ey + 1.f)) * tanf(PI_F * u1 * 0.5f));
    float cosphi = cosf(*phi), sinphi = sinf(*phi);
    *costheta = powf(u2, 1.f / (b->ex * cosphi * cosphi + b->ey * sinphi * sinphi + 1));
}
/* Blinn::Pdf reflection.cpp:356-366 / BxDF::Pdf :312-315 / Microfacet::Pdf :331-335 */
static float bxdf_pdf(const BSDF *b, int i, v3 wo, v3 wi) {
    int kind = b->kind[i];
    if (kind == BX_LAMBERT || kind == BX_ORENNAYAR || kind == BX_MEASURED)
        return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * INV_PI_F : 0.f;
    if (!same_hemisphere(wo, wi)) return 0.f;
    if (kind == BX_FRESNEL_BLEND)                              /* FresnelBlend::Pdf, reflection.cpp:453-456 */
        return .5f * (abs_cos_theta(wi) * INV_PI_F + anisotropic_pdf(b, wo, normalize(vadd(wo, wi))));
    v3 wh = normalize(vadd(wo, wi));
    float costheta = abs_cos_theta(wh);
    float blinn_pdf = ((b->exponent + 1.f) * powf(costheta, b->exponent)) / (2.f * PI_F * 4.f * dot(wo, wh));
    if (dot(wo, wh) <= 0.f) blinn_pdf = 0.f;
    return blinn_pdf;
}

/* BSDF::f, reflection.cpp:604-618. flags: all lowered BxDFs are non-specular reflection, so
 * BSDF_ALL & ~BSDF_SPECULAR and BSDF_ALL select the same set; only the geometric-normal
 * reflection/transmission switch matters. */
static void bsdf_f(const BSDF *b, v3 woW, v3 wiW, float *f) {
    v3 wi = w2l(b, wiW), wo = w2l(b, woW);
    int reflect = dot(wiW, b->ng) * dot(woW, b->ng) > 0;
    for (int c = 0; c < NB; ++c) f[c] = 0.f;
    if (!reflect) return;           /* BRDFs ignored; the only BTDF is specular */
    for (int i = 0; i < b->nBxDFs; ++i) if (!BX_IS_SPECULAR(b->kind[i])) bxdf_f(b, i, wo, wi, f);
}
/* BSDF::Pdf, reflection.cpp:575-590 */
static float bsdf_pdf(const BSDF *b, v3 woW, v3 wiW) {
    if (b->nBxDFs == 0.) return 0.;
    v3 wo = w2l(b, woW), wi = w2l(b, wiW);
    float pdf = 0.f;
    int matching = 0;
    for (int i = 0; i < b->nBxDFs; ++i) if (!BX_IS_SPECULAR(b->kind[i])) { ++matching; pdf += bxdf_pdf(b, i, wo, wi); }
    return matching > 0 ? pdf / matching : 0.f;
}
/* BSDF::Sample_f, reflection.cpp:514-572. allowSpecular: flags = BSDF_ALL (path continuation), else
 * BSDF_ALL & ~BSDF_SPECULAR (EstimateDirect). A lowered BSDF is either all specular (mirror, glass) or has
 * no specular component. *specular: the sampled BxDF's type has BSDF_SPECULAR. */
static void bsdf_sample_f(const BSDF *b, v3 woW, v3 *wiW, float uComp, float u1, float u2, float *pdf, float *f,
                          int allowSpecular, int *specular) {
    int matching = 0;
    int idx[2];
    for (int i = 0; i < b->nBxDFs; ++i) if (allowSpecular || !BX_IS_SPECULAR(b->kind[i])) idx[matching++] = i;
    for (int c = 0; c < NB; ++c) f[c] = 0.f;
    if (specular) *specular = 0;
    if (matching == 0) { *pdf = 0.f; return; }
    int which = (int)floorf(uComp * matching);
    if (matching - 1 < which) which = matching - 1;
    which = idx[which];
    v3 wo = w2l(b, woW);
    v3 wi;
    *pdf = 0.f;
    int kind = b->kind[which];
    if (BX_IS_SPECULAR(kind)) {
        float fr = 1.f;                                        /* FresnelNoOp */
        if (kind != BX_SPEC_REFL_NOOP) fr = fresnel_dielectric(wo.z, 1.f, b->ior);
        if (kind == BX_SPEC_TRANS) {                           /* SpecularTransmission::Sample_f, reflection.cpp:139-162 */
            int entering = wo.z > 0.;
            float ei = 1.f, et = b->ior;
            if (!entering) { float t = ei; ei = et; et = t; }
            float sini2 = sin_theta2(wo);
            float eta = ei / et;
            float sint2 = eta * eta * sini2;
            if (sint2 >= 1.) return;                           /* total internal reflection: pdf stays 0 */
            float cost = sqrtf(stdmaxf(0.f, 1.f - sint2));
            if (entering) cost = -cost;
            float sintOverSini = eta;
            wi = V(sintOverSini * -wo.x, sintOverSini * -wo.y, cost);
            *pdf = 1.f;
            for (int c = 0; c < NB; ++c) f[c] = (1.f - fr) * b->R[which][c] / abs_cos_theta(wi);
        } else {                                               /* SpecularReflection::Sample_f, reflection.cpp:130-136 */
            wi = V(-wo.x, -wo.y, wo.z);
            *pdf = 1.f;
            for (int c = 0; c < NB; ++c) f[c] = fr * b->R[which][c] / abs_cos_theta(wi);
        }
        if (specular) *specular = 1;
        *wiW = l2w(b, wi);
        if (matching > 1) *pdf /= matching;
        return;
    }
    if (kind == BX_LAMBERT || kind == BX_ORENNAYAR || kind == BX_MEASURED) {   /* BxDF::Sample_f, reflection.cpp:303-310 */
        wi = cosine_sample_hemisphere(u1, u2);
        if (wo.z < 0.) wi.z *= -1.f;
        *pdf = bxdf_pdf(b, which, wo, wi);
    } else if (kind == BX_FRESNEL_BLEND) {                     /* FresnelBlend::Sample_f, reflection.cpp:435-450 */
        int keepPdf = 0;
        if (u1 < .5) {
            u1 = 2.f * u1;
            wi = cosine_sample_hemisphere(u1, u2);
            if (wo.z < 0.) wi.z *= -1.f;
        } else {                                               /* Anisotropic::Sample_f, reflection.cpp:369-405 */
            u1 = 2.f * (u1 - .5f);
            float phi, costheta;
            if (u1 < .25f) aniso_first_quadrant(b, 4.f * u1, u2, &phi, &costheta);
            else if (u1 < .5f) { u1 = 4.f * (.5f - u1); aniso_first_quadrant(b, u1, u2, &phi, &costheta); phi = PI_F - phi; }
            else if (u1 < .75f) { u1 = 4.f * (u1 - .5f); aniso_first_quadrant(b, u1, u2, &phi, &costheta); phi += PI_F; }
            else { u1 = 4.f * (1.f - u1); aniso_first_quadrant(b, u1, u2, &phi, &costheta); phi = 2.f * PI_F - phi; }
            float sintheta = sqrtf(stdmaxf(0.f, 1.f - costheta * costheta));
            v3 wh = V(sintheta * cosf(phi), sintheta * sinf(phi), costheta);
            if (!same_hemisphere(wo, wh)) wh = vneg(wh);
            wi = vadd(vneg(wo), vmul(wh, 2.f * dot(wo, wh)));
            *pdf = anisotropic_pdf(b, wo, wh);
            if (!same_hemisphere(wo, wi)) keepPdf = 1;          /* returns before *pdf = Pdf(wo, *wi) */
        }
        if (!keepPdf) *pdf = bxdf_pdf(b, which, wo, wi);
    } else {                                                   /* Microfacet::Sample_f :324-329, Blinn::Sample_f :338-354 */
        float costheta = powf(u1, 1.f / (b->exponent + 1));
        float sintheta = sqrtf(stdmaxf(0.f, 1.f - costheta * costheta));
        float phi = u2 * 2.f * PI_F;
        v3 wh = V(sintheta * cosf(phi), sintheta * sinf(phi), costheta);
        if (!same_hemisphere(wo, wh)) wh = vneg(wh);
        wi = vadd(vneg(wo), vmul(wh, 2.f * dot(wo, wh)));
        float blinn_pdf = ((b->exponent + 1.f) * powf(costheta, b->exponent)) / (2.f * PI_F * 4.f * dot(wo, wh));
        if (dot(wo, wh) <= 0.f) blinn_pdf = 0.f;
        *pdf = blinn_pdf;
    }
    if (*pdf == 0.f) return;
    *wiW = l2w(b, wi);
    if (matching > 1)
        for (int i = 0; i < b->nBxDFs; ++i)
            if (i != which) *pdf += bxdf_pdf(b, i, wo, wi);
    if (matching > 1) *pdf /= matching;
    int reflect = dot(*wiW, b->ng) * dot(woW, b->ng) > 0;
    if (!reflect) return;
    for (int i = 0; i < b->nBxDFs; ++i) bxdf_f(b, i, wo, wi, f);
}

static int is_black(const float *s) { for (int c = 0; c < NB; ++c) if (s[c] != 0.) return 0; return 1; }
/* SampledSpectrum::y, core/spectrum.h:417-422 */
static float spectrum_y(const SptSpectralTables *t, const float *s) {
    float yy = 0.f;
    for (int c = 0; c < NB; ++c) yy += t->cie_y[c] * s[c];
    return yy / t->yint;
}

/* ------------------------------------------------------------------------------------------ */
/* lights */
/* SampledSpectrum::FromRGB(rgb, SPECTRUM_ILLUMINANT), core/spectrum.cpp:136-176 */
static void from_rgb_illum(const SptSpectralTables *t, const float rgb[3], float *r) {
    enum { W = 0, Cy = 1, Mg = 2, Ye = 3, Rd = 4, Gr = 5, Bl = 6 };
    for (int c = 0; c < NB; ++c) r[c] = 0.f;
#define ADD(coef, basis) do { float k_ = (coef); for (int c = 0; c < NB; ++c) r[c] += t->rgb_illum[basis][c] * k_; } while (0)
    if (rgb[0] <= rgb[1] && rgb[0] <= rgb[2]) {
        ADD(rgb[0], W);
        if (rgb[1] <= rgb[2]) { ADD(rgb[1] - rgb[0], Cy); ADD(rgb[2] - rgb[1], Bl); }
        else { ADD(rgb[2] - rgb[0], Cy); ADD(rgb[1] - rgb[2], Gr); }
    } else if (rgb[1] <= rgb[0] && rgb[1] <= rgb[2]) {
        ADD(rgb[1], W);
        if (rgb[0] <= rgb[2]) { ADD(rgb[0] - rgb[1], Mg); ADD(rgb[2] - rgb[0], Bl); }
        else { ADD(rgb[2] - rgb[1], Mg); ADD(rgb[0] - rgb[2], Rd); }
    } else {
        ADD(rgb[2], W);
        if (rgb[0] <= rgb[1]) { ADD(rgb[0] - rgb[2], Ye); ADD(rgb[1] - rgb[0], Gr); }
        else { ADD(rgb[1] - rgb[2], Ye); ADD(rgb[0] - rgb[1], Rd); }
    }
#undef ADD
    for (int c = 0; c < NB; ++c) { r[c] *= .86445f; r[c] = clampf(r[c], 0.f, INFINITY); }
}
static int imod(int a, int b) { int n = (int)(a / b); a -= n * b; if (a < 0) a += b; return a; }   /* core/pbrt.h:225-230 */
/* MIPMap::Lookup(s,t,width=0) -> triangle(0,s,t), core/mipmap.h:233-274 with TEXTURE_REPEAT Texel :198-221 */
static void env_lookup(const SptSceneDesc *sc, float s, float t, float rgb[3]) {
    int w = sc->env_w, h = sc->env_h;
    s = s * w - 0.5f;
    t = t * h - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(t);
    float ds = s - s0, dt = t - t0;
    const float *a = sc->env_rgb + 3 * ((size_t)imod(t0, h) * w + imod(s0, w));
    const float *b = sc->env_rgb + 3 * ((size_t)imod(t0 + 1, h) * w + imod(s0, w));
    const float *c = sc->env_rgb + 3 * ((size_t)imod(t0, h) * w + imod(s0 + 1, w));
    const float *d = sc->env_rgb + 3 * ((size_t)imod(t0 + 1, h) * w + imod(s0 + 1, w));
    for (int k = 0; k < 3; ++k)
        rgb[k] = a[k] * ((1.f - ds) * (1.f - dt)) + b[k] * ((1.f - ds) * dt) + c[k] * (ds * (1.f - dt)) + d[k] * (ds * dt);
}
static float spherical_theta(v3 v) { return acosf(clampf(v.z, -1.f, 1.f)); }
static float spherical_phi(v3 v) { float p = atan2f(v.y, v.x); return (p < 0.f) ? p + 2.f * PI_F : p; }
/* InfiniteAreaLight::Le, lights/infinite.cpp:109-114 */
static void infinite_le(const SptSceneDesc *sc, const SptLight *l, v3 d, float *out) {
    const SptXform *xf = sc->xforms + l->xform;
    v3 wh = normalize(xf_vector(xf->minv, d));
    float s = spherical_phi(wh) * INV_TWOPI_F;
    float t = spherical_theta(wh) * INV_PI_F;
    float rgb[3];
    env_lookup(sc, s, t, rgb);
    from_rgb_illum(&sc->tables, rgb, out);
}
/* Distribution1D::SampleContinuous, core/montecarlo.h:70-88 (std::upper_bound) */
static float dist1d_sample_continuous(const float *func, const float *cdf, float funcInt, int count, float u,
                                      float *pdf, int *off) {
    int lo = 0, hi = count + 1;                       /* first index with cdf[idx] > u */
    while (lo < hi) { int mid = (lo + hi) / 2; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    int offset = lo - 1; if (offset < 0) offset = 0;
    if (off) *off = offset;
    float du = (u - cdf[offset]) / (cdf[offset + 1] - cdf[offset]);
    if (pdf) *pdf = func[offset] / funcInt;
    return (offset + du) / count;
}
static int dist1d_sample_discrete(const float *cdf, int count, float u) {
    int lo = 0, hi = count + 1;
    while (lo < hi) { int mid = (lo + hi) / 2; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    int offset = lo - 1; if (offset < 0) offset = 0;
    return offset;
}
/* Distribution2D::Pdf, core/montecarlo.h:146-154 */
static float env_pdf_uv(const SptSceneDesc *sc, float u, float v) {
    int nu = sc->env_w, nv = sc->env_h;
    int iu = clampi((int)(u * nu), 0, nu - 1);
    int iv = clampi((int)(v * nv), 0, nv - 1);
    if (sc->env_func_int[iv] * sc->env_marg_int == 0.f) return 0.f;
    return (sc->env_func[(size_t)iv * nu + iu] * sc->env_marg_func[iv]) / (sc->env_func_int[iv] * sc->env_marg_int);
}

/* Shape::Sample(u1,u2,Ns): Triangle shapes/trianglemesh.cpp:436-448, Disk shapes/disk.cpp:140-149,
 * Sphere shapes/sphere.cpp:220-225 */
static v3 shape_sample_area(const SptSceneDesc *sc, const SptLightShape *s, float u1, float u2, v3 *ns) {
    if (s->kind == SPT_PRIM_TRIANGLE) {
        float b1, b2;
        uniform_sample_triangle(u1, u2, &b1, &b2);
        const int32_t *vi = sc->tri_vidx + 3 * (size_t)s->data;
        v3 p1 = vert(sc, vi[0]), p2 = vert(sc, vi[1]), p3 = vert(sc, vi[2]);
        v3 p = vadd(vadd(vmul(p1, b1), vmul(p2, b2)), vmul(p3, (1.f - b1 - b2)));
        v3 n = cross(vsub(p2, p1), vsub(p3, p1));
        *ns = normalize(n);
        if (s->flags & SPT_PF_REVERSE) *ns = vmul(*ns, -1.f);
        return p;
    }
    const SptQuadric *q = sc->quadrics + s->data;
    const SptXform *xf = sc->xforms + q->xform;
    if (s->kind == SPT_PRIM_DISK) {
        v3 p;
        concentric_sample_disk(u1, u2, &p.x, &p.y);
        p.x *= q->radius; p.y *= q->radius; p.z = q->zmin;
        *ns = normalize(xf_normal(xf->minv, V(0, 0, 1)));
        if (s->flags & SPT_PF_REVERSE) *ns = vmul(*ns, -1.f);
        return xf_point(xf->m, p);
    }
    v3 p = vadd(V(0, 0, 0), vmul(uniform_sample_sphere(u1, u2), q->radius));
    *ns = normalize(xf_normal(xf->minv, V(p.x, p.y, p.z)));
    if (s->flags & SPT_PF_REVERSE) *ns = vmul(*ns, -1.f);
    return xf_point(xf->m, p);
}
/* Shape::Sample(p,u1,u2,Ns): default = area sampling (core/shape.h); Sphere shapes/sphere.cpp:228-252 */
static v3 shape_sample_from(const SptSceneDesc *sc, const SptLightShape *s, v3 p, float u1, float u2, v3 *ns) {
    if (s->kind != SPT_PRIM_SPHERE) return shape_sample_area(sc, s, u1, u2, ns);
    const SptQuadric *q = sc->quadrics + s->data;
    const SptXform *xf = sc->xforms + q->xform;
    v3 Pcenter = xf_point(xf->m, V(0, 0, 0));
    v3 wc = normalize(vsub(Pcenter, p));
    v3 wcX, wcY;
    coordinate_system(wc, &wcX, &wcY);
    if (len2(vsub(p, Pcenter)) - q->radius * q->radius < 1e-4f) return shape_sample_area(sc, s, u1, u2, ns);
    float sinThetaMax2 = q->radius * q->radius / len2(vsub(p, Pcenter));
    float cosThetaMax = sqrtf(stdmaxf(0.f, 1.f - sinThetaMax2));
    Ray r; r.o = p; r.d = uniform_sample_cone(u1, u2, cosThetaMax, wcX, wcY, wc); r.mint = 1e-3f; r.maxt = INFINITY; r.depth = 0;
    Hit h;
    float thit;
    if (!sphere_intersect(sc, q, s->flags, &r, &h)) thit = dot(vsub(Pcenter, p), normalize(r.d));
    else thit = h.t;
    v3 ps = ray_at(&r, thit);
    *ns = normalize(vsub(ps, Pcenter));
    if (s->flags & SPT_PF_REVERSE) *ns = vmul(*ns, -1.f);
    return ps;
}
/* Shape::Pdf(p,wi) core/shape.cpp:78-91; Sphere::Pdf shapes/sphere.cpp:255-266 */
static float shape_pdf(const SptSceneDesc *sc, const SptLightShape *s, v3 p, v3 wi) {
    if (s->kind == SPT_PRIM_SPHERE) {
        const SptQuadric *q = sc->quadrics + s->data;
        const SptXform *xf = sc->xforms + q->xform;
        v3 Pcenter = xf_point(xf->m, V(0, 0, 0));
        if (!(len2(vsub(p, Pcenter)) - q->radius * q->radius < 1e-4f)) {
            float sinThetaMax2 = q->radius * q->radius / len2(vsub(p, Pcenter));
            float cosThetaMax = sqrtf(stdmaxf(0.f, 1.f - sinThetaMax2));
            return uniform_cone_pdf(cosThetaMax);
        }
    }
    Ray ray; ray.o = p; ray.d = wi; ray.mint = 1e-3f; ray.maxt = INFINITY; ray.depth = -1;
    Hit h;
    if (!shape_intersect(sc, s->kind, s->flags, (uint32_t)s->data, &ray, &h)) return 0.;
    float pdf = len2(vsub(p, ray_at(&ray, h.t))) / (absdot(h.nn, vneg(wi)) * s->area);
    if (isinf(pdf)) pdf = 0.f;
    return pdf;
}
/* ShapeSet::Pdf(p,wi), core/light.cpp:156-161 */
static float shapeset_pdf(const SptSceneDesc *sc, const SptLight *l, v3 p, v3 wi) {
    float pdf = 0.f;
    for (int i = 0; i < l->shape_count; ++i) {
        const SptLightShape *s = sc->light_shapes + l->shape_first + i;
        pdf += s->area * shape_pdf(sc, s, p, wi);
    }
    return pdf / l->sum_area;
}
/* Distribution1D ctor for the ShapeSet's area distribution, core/montecarlo.h:48-68 */
static void area_cdf(const SptSceneDesc *sc, const SptLight *l, float *cdf) {
    int n = l->shape_count;
    cdf[0] = 0.;
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + sc->light_shapes[l->shape_first + i - 1].area / n;
    float funcInt = cdf[n];
    if (funcInt == 0.f) for (int i = 1; i < n + 1; ++i) cdf[i] = (float)i / (float)n;
    else for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
}

typedef struct { int is_delta; float Li[NB]; v3 wi; float pdf; Ray shadow; } LightSampleResult;

/* Light::Sample_L: DiffuseAreaLight lights/diffuse.cpp:61-73 (+ShapeSet::Sample core/light.cpp:137-149),
 * PointLight lights/point.cpp:42-49, InfiniteAreaLight lights/infinite.cpp:187-213;
 * VisibilityTester core/light.h:79-88 */
static void light_sample(const SptSceneDesc *sc, const SptLight *l, v3 p, float pEpsilon,
                         float uPos0, float uPos1, float uComp, LightSampleResult *out) {
    for (int c = 0; c < NB; ++c) out->Li[c] = 0.f;
    out->pdf = 0.f;
    out->is_delta = 0;
    if (l->type == SPT_LIGHT_POINT) {
        v3 lp = V(l->pos[0], l->pos[1], l->pos[2]);
        out->is_delta = 1;
        out->wi = normalize(vsub(lp, p));
        out->pdf = 1.f;
        float dist = sqrtf(len2(vsub(p, lp)));
        out->shadow.o = p; out->shadow.d = vdiv(vsub(lp, p), dist); out->shadow.mint = pEpsilon;
        out->shadow.maxt = dist * (1.f - 0.f); out->shadow.depth = 0;
        float d2 = len2(vsub(lp, p));
        for (int c = 0; c < NB; ++c) out->Li[c] = l->spectrum[c] / d2;
        return;
    }
    if (l->type == SPT_LIGHT_AREA) {
        float *cdf = (float *)alloca(sizeof(float) * (l->shape_count + 1));
        area_cdf(sc, l, cdf);
        int sn = dist1d_sample_discrete(cdf, l->shape_count, uComp);
        v3 ns;
        v3 pt = shape_sample_from(sc, sc->light_shapes + l->shape_first + sn, p, uPos0, uPos1, &ns);
        Ray r; r.o = p; r.d = vsub(pt, p); r.mint = 1e-3f; r.maxt = INFINITY; r.depth = 0;
        float thit = 1.f;
        int anyHit = 0;
        Hit h;
        for (int i = 0; i < l->shape_count; ++i) {
            const SptLightShape *s = sc->light_shapes + l->shape_first + i;
            Hit hh;
            if (shape_intersect(sc, s->kind, s->flags, (uint32_t)s->data, &r, &hh)) { anyHit = 1; h = hh; thit = hh.t; }
        }
        if (anyHit) ns = h.nn;
        v3 ps = ray_at(&r, thit);
        out->wi = normalize(vsub(ps, p));
        out->pdf = shapeset_pdf(sc, l, p, out->wi);
        float dist = sqrtf(len2(vsub(p, ps)));
        out->shadow.o = p; out->shadow.d = vdiv(vsub(ps, p), dist); out->shadow.mint = pEpsilon;
        out->shadow.maxt = dist * (1.f - 1e-3f); out->shadow.depth = 0;
        if (dot(ns, vneg(out->wi)) > 0.f) for (int c = 0; c < NB; ++c) out->Li[c] = l->spectrum[c];
        return;
    }
    /* infinite */
    float uv[2], pdfs[2];
    int v;
    uv[1] = dist1d_sample_continuous(sc->env_marg_func, sc->env_marg_cdf, sc->env_marg_int, sc->env_h, uPos1, &pdfs[1], &v);
    uv[0] = dist1d_sample_continuous(sc->env_func + (size_t)v * sc->env_w, sc->env_cdf + (size_t)v * (sc->env_w + 1),
                                     sc->env_func_int[v], sc->env_w, uPos0, &pdfs[0], NULL);
    float mapPdf = pdfs[0] * pdfs[1];
    if (mapPdf == 0.f) return;
    float theta = uv[1] * PI_F, phi = uv[0] * 2.f * PI_F;
    float costheta = cosf(theta), sintheta = sinf(theta);
    float sinphi = sinf(phi), cosphi = cosf(phi);
    const SptXform *xf = sc->xforms + l->xform;
    out->wi = xf_vector(xf->m, V(sintheta * cosphi, sintheta * sinphi, costheta));
    out->pdf = mapPdf / (2.f * PI_F * PI_F * sintheta);
    if (sintheta == 0.f) out->pdf = 0.f;
    out->shadow.o = p; out->shadow.d = out->wi; out->shadow.mint = pEpsilon; out->shadow.maxt = INFINITY; out->shadow.depth = 0;
    float rgb[3];
    env_lookup(sc, uv[0], uv[1], rgb);
    from_rgb_illum(&sc->tables, rgb, out->Li);
}
/* Light::Pdf(p,wi): lights/diffuse.cpp:76-78, lights/infinite.cpp:216-226 */
static float light_pdf(const SptSceneDesc *sc, const SptLight *l, v3 p, v3 w) {
    if (l->type == SPT_LIGHT_AREA) return shapeset_pdf(sc, l, p, w);
    if (l->type == SPT_LIGHT_POINT) return 0.f;
    const SptXform *xf = sc->xforms + l->xform;
    v3 wi = xf_vector(xf->minv, w);
    float theta = spherical_theta(wi), phi = spherical_phi(wi);
    float sintheta = sinf(theta);
    if (sintheta == 0.f) return 0.f;
    return env_pdf_uv(sc, phi * INV_TWOPI_F, theta * INV_PI_F) / (2.f * PI_F * PI_F * sintheta);
}
/* Light::Le(ray): zero except the infinite light (core/light.cpp:51-53) */
static void light_le(const SptSceneDesc *sc, const SptLight *l, v3 d, float *out) {
    if (l->type == SPT_LIGHT_INFINITE) infinite_le(sc, l, d, out);
    else for (int c = 0; c < NB; ++c) out[c] = 0.f;
}
/* Intersection::Le, core/intersection.cpp:53-56 + DiffuseAreaLight::L lights/diffuse.h:43-45 */
static void isect_le(const SptSceneDesc *sc, uint32_t slot, const Hit *h, v3 w, float *out) {
    int li = sc->prim_light[slot];
    for (int c = 0; c < NB; ++c) out[c] = 0.f;
    if (li < 0) return;
    if (dot(h->nn, w) > 0.f) for (int c = 0; c < NB; ++c) out[c] = sc->lights[li].spectrum[c];
}

/* ------------------------------------------------------------------------------------------ */
typedef struct { const float *p; int n, i; } RngStream;
static float rng_next(RngStream *r) { return r->i < r->n ? r->p[r->i++] : 0.5f; }

/* EstimateDirect, core/integrator.cpp:109-166 (flags = BSDF_ALL & ~BSDF_SPECULAR) */
static void estimate_direct(const SptSceneDesc *sc, const SptLight *light, int lightIdx, v3 p, v3 n, v3 wo,
                            float rayEpsilon, const BSDF *bsdf, const float ls[3], const float bs[3], float *Ld) {
    for (int c = 0; c < NB; ++c) Ld[c] = 0.f;
    LightSampleResult lr;
    light_sample(sc, light, p, rayEpsilon, ls[0], ls[1], ls[2], &lr);
    float f[NB];
    if (lr.pdf > 0. && !is_black(lr.Li)) {
        bsdf_f(bsdf, wo, lr.wi, f);
        if (!is_black(f)) {
            Ray sh = lr.shadow;
            if (!bvh_intersect(sc, &sh, 1, NULL, NULL, NULL, NULL)) {
                if (lr.is_delta) {
                    float s = (absdot(lr.wi, n) / lr.pdf);
                    for (int c = 0; c < NB; ++c) Ld[c] += f[c] * lr.Li[c] * s;
                } else {
                    float bsdfPdf = bsdf_pdf(bsdf, wo, lr.wi);
                    float weight = power_heuristic(1, lr.pdf, 1, bsdfPdf);
                    float s = (absdot(lr.wi, n) * weight / lr.pdf);
                    for (int c = 0; c < NB; ++c) Ld[c] += f[c] * lr.Li[c] * s;
                }
            }
        }
    }
    if (!lr.is_delta) {
        v3 wi;
        float bsdfPdf;
        bsdf_sample_f(bsdf, wo, &wi, bs[2], bs[0], bs[1], &bsdfPdf, f, 0, NULL);
        if (!is_black(f) && bsdfPdf > 0.) {
            float weight = 1.f;
            float lightPdf = light_pdf(sc, light, p, wi);
            if (lightPdf == 0.) return;
            weight = power_heuristic(1, bsdfPdf, 1, lightPdf);
            float Li[NB];
            for (int c = 0; c < NB; ++c) Li[c] = 0.f;
            Ray ray; ray.o = p; ray.d = wi; ray.mint = rayEpsilon; ray.maxt = INFINITY; ray.depth = 0;
            uint32_t slot; Hit h;
            if (bvh_intersect(sc, &ray, 0, &slot, &h, NULL, NULL)) {
                if (sc->prim_light[slot] == lightIdx) isect_le(sc, slot, &h, vneg(wi), Li);
            } else light_le(sc, light, ray.d, Li);
            if (!is_black(Li))
                for (int c = 0; c < NB; ++c) Ld[c] += f[c] * Li[c] * absdot(wi, n) * weight / bsdfPdf;
        }
    }
}

/* SamplerRenderer::Li (renderers/samplerrenderer.cpp:225-247) + PathIntegrator::Li
 * (integrators/path.cpp:44-115) + UniformSampleOneLight (core/integrator.cpp:74-106).
 * sample37 layout: core/sampler.cpp:88-117 with the Add1D/Add2D order of integrators/path.cpp:33-41:
 *   1-D [5 + 4*i + {0 lightComp, 1 lightNum, 2 bsdfComp, 3 pathComp}], i < 3; [17],[18] emission (unused)
 *   2-D [19 + 6*i + {0 lightPos, 2 bsdfDir, 4 pathDir}]                                               */
static void li_sample(const SptSceneDesc *sc, const SptCameraDesc *cam, int maxDepth, int spp, const float *smp,
                      const float *rngv, int nrng, float *Lout) {
    RngStream rng = { rngv, nrng, 0 };
    Ray ray;
    RayDiff rd;
    camera_ray_diff(cam, smp, spp, &ray, &rd);
    float L[NB], T[NB];
    for (int c = 0; c < NB; ++c) { L[c] = 0.f; T[c] = 1.f; }
    uint32_t slot; Hit isect;
    if (!bvh_intersect(sc, &ray, 0, &slot, &isect, NULL, NULL)) {
        for (uint32_t i = 0; i < sc->n_lights; ++i) {
            float le[NB];
            light_le(sc, sc->lights + i, ray.d, le);
            for (int c = 0; c < NB; ++c) L[c] += le[c];
        }
        memcpy(Lout, L, sizeof(L));
        return;
    }
    int specularBounce = 0;
    int nLights = (int)sc->n_lights;
    for (int bounces = 0;; ++bounces) {
        if (bounces == 0 || specularBounce) {
            float le[NB];
            isect_le(sc, slot, &isect, vneg(ray.d), le);
            for (int c = 0; c < NB; ++c) L[c] += T[c] * le[c];
        }
        BSDF bsdf; v3 n;
        make_bsdf(sc, slot, &isect, bounces == 0 ? &rd : NULL, &bsdf, &n);
        v3 p = isect.p;
        v3 wo = vneg(ray.d);
        /* UniformSampleOneLight */
        if (nLights > 0) {
            float lightNumU, ls[3], bs[3];
            if (bounces < 3) {
                const float *oneD = smp + 5 + 4 * bounces;
                const float *twoD = smp + 19 + 6 * bounces;
                lightNumU = oneD[1];
                ls[0] = twoD[0]; ls[1] = twoD[1]; ls[2] = oneD[0];
                bs[0] = twoD[2]; bs[1] = twoD[3]; bs[2] = oneD[2];
            } else {
                lightNumU = rng_next(&rng);
                ls[0] = rng_next(&rng); ls[1] = rng_next(&rng); ls[2] = rng_next(&rng);   /* light.h:107-111 */
                bs[0] = rng_next(&rng); bs[1] = rng_next(&rng); bs[2] = rng_next(&rng);   /* reflection.cpp:502-507 */
            }
            int lightNum = (int)floorf(lightNumU * nLights);
            if (nLights - 1 < lightNum) lightNum = nLights - 1;
            float Ld[NB];
            estimate_direct(sc, sc->lights + lightNum, lightNum, p, n, wo, isect.rayEpsilon, &bsdf, ls, bs, Ld);
            for (int c = 0; c < NB; ++c) L[c] += T[c] * (Ld[c] * (float)nLights);
        }
        float u[3];
        if (bounces < 3) {
            u[0] = smp[19 + 6 * bounces + 4]; u[1] = smp[19 + 6 * bounces + 5]; u[2] = smp[5 + 4 * bounces + 3];
        } else { u[0] = rng_next(&rng); u[1] = rng_next(&rng); u[2] = rng_next(&rng); }
        v3 wi; float pdf; float f[NB];
        int spec;
        bsdf_sample_f(&bsdf, wo, &wi, u[2], u[0], u[1], &pdf, f, 1, &spec);
        if (is_black(f) || pdf == 0.) break;
        specularBounce = spec;
        float ad = absdot(wi, n);
        for (int c = 0; c < NB; ++c) T[c] *= f[c] * ad / pdf;
        float eps = isect.rayEpsilon;
        ray.o = p; ray.d = wi; ray.mint = eps; ray.maxt = INFINITY; ray.depth++;
        if (bounces > 3) {
            float continueProbability = stdminf(.5f, spectrum_y(&sc->tables, T));
            if (rng_next(&rng) > continueProbability) break;
            for (int c = 0; c < NB; ++c) T[c] /= continueProbability;
        }
        if (bounces == maxDepth) break;
        if (!bvh_intersect(sc, &ray, 0, &slot, &isect, NULL, NULL)) {
            if (specularBounce)
                for (uint32_t i = 0; i < sc->n_lights; ++i) {
                    float le[NB];
                    light_le(sc, sc->lights + i, ray.d, le);
                    for (int c = 0; c < NB; ++c) L[c] += T[c] * le[c];
                }
            break;
        }
    }
    memcpy(Lout, L, sizeof(L));
}

/* SamplerRenderer::Li (renderers/samplerrenderer.cpp:225-247) + DirectLightingIntegrator::Li with strategy "all"
 * (integrators/directlighting.cpp:70-105) + UniformSampleAllLights (core/integrator.cpp:39-71). The scene has no
 * specular BxDF (lowering), so SpecularReflect / SpecularTransmit (core/integrator.cpp:169-236) sample nothing.
 * Sample layout (directlighting.cpp:46-60, light.cpp:56-60, reflection.cpp:494-498, sampler.cpp:88-117), N = sum n_i:
 *   1-D from [5]: per light {light component x n_i, bsdf component x n_i}; then 2 volume-integrator floats
 *   2-D from [7 + 2N]: per light {light position x 2 n_i, bsdf direction x 2 n_i}                              */
static int direct_sample_floats(const SptSceneDesc *sc) {
    int N = 0;
    for (uint32_t i = 0; i < sc->n_lights; ++i) N += sc->lights[i].n_samples;
    return 7 + 6 * N;
}
/* SamplerRenderer::Li + DirectLightingIntegrator::Li for one ray of the recursion (samplerrenderer.cpp:225-247,
 * directlighting.cpp:70-107): emitted light + UniformSampleAllLights / UniformSampleOneLight at the hit - every level reads the
 * SAME Sample arrays - then SpecularReflect / SpecularTransmit (integrator.cpp:169-250) while depth + 1 < maxDepth. The
 * RNG values BSDFSample(rng) draws there choose among the BxDFs matching (REFLECTION | SPECULAR) resp. (TRANSMISSION |
 * SPECULAR) - one at most for every lowered material, and a perfectly specular direction ignores u - so they do not enter.
 * rd: the camera ray's differentials (level 0); the children's differentials only reach image-texture filtering, which the
 * lowering does not combine with specular materials under this integrator. */
static void direct_li(const SptSceneDesc *sc, int strategy_all, int maxDepth, int depth, Ray ray, const RayDiff *rd, const float *smp,
                      float *Lout) {
    float L[NB];
    for (int c = 0; c < NB; ++c) L[c] = 0.f;
    uint32_t slot; Hit isect;
    if (!bvh_intersect(sc, &ray, 0, &slot, &isect, NULL, NULL)) {
        for (uint32_t i = 0; i < sc->n_lights; ++i) {
            float le[NB];
            light_le(sc, sc->lights + i, ray.d, le);
            for (int c = 0; c < NB; ++c) L[c] += le[c];
        }
        memcpy(Lout, L, sizeof(L));
        return;
    }
    BSDF bsdf; v3 n;
    make_bsdf(sc, slot, &isect, rd, &bsdf, &n);
    v3 p = isect.p, wo = vneg(ray.d);
    float le[NB];
    isect_le(sc, slot, &isect, wo, le);
    for (int c = 0; c < NB; ++c) L[c] += le[c];
    if (strategy_all) {
        int N = 0;
        for (uint32_t i = 0; i < sc->n_lights; ++i) N += sc->lights[i].n_samples;
        const float *oneD = smp + 5, *twoD = smp + 7 + 2 * N;
        float Lall[NB];
        for (int c = 0; c < NB; ++c) Lall[c] = 0.f;
        for (uint32_t i = 0; i < sc->n_lights; ++i) {
            int nSamples = sc->lights[i].n_samples;
            float Ld[NB];
            for (int c = 0; c < NB; ++c) Ld[c] = 0.f;
            for (int j = 0; j < nSamples; ++j) {
                float ls[3] = { twoD[2 * j], twoD[2 * j + 1], oneD[j] };
                float bs[3] = { twoD[2 * nSamples + 2 * j], twoD[2 * nSamples + 2 * j + 1], oneD[nSamples + j] };
                float e[NB];
                estimate_direct(sc, sc->lights + i, (int)i, p, n, wo, isect.rayEpsilon, &bsdf, ls, bs, e);
                for (int c = 0; c < NB; ++c) Ld[c] += e[c];
            }
            for (int c = 0; c < NB; ++c) Lall[c] += Ld[c] / nSamples;
            oneD += 2 * nSamples; twoD += 4 * nSamples;
        }
        for (int c = 0; c < NB; ++c) L[c] += Lall[c];
    } else {
        /* strategy "one" (directlighting.cpp:61-68): [5] light component, [6] light number, [7] bsdf component, [8],[9] volume
         * integrator, [10],[11] light position, [12],[13] bsdf direction */
        int nLights = (int)sc->n_lights;
        if (nLights > 0) {
            int lightNum = (int)floorf(smp[6] * nLights);
            if (nLights - 1 < lightNum) lightNum = nLights - 1;
            float ls[3] = { smp[10], smp[11], smp[5] }, bs[3] = { smp[12], smp[13], smp[7] };
            float Ld[NB];
            estimate_direct(sc, sc->lights + lightNum, lightNum, p, n, wo, isect.rayEpsilon, &bsdf, ls, bs, Ld);
            for (int c = 0; c < NB; ++c) L[c] += Ld[c] * (float)nLights;
        }
    }
    if (depth + 1 < maxDepth) {
        /* SpecularReflect, then SpecularTransmit: BSDF::Sample_f restricted to the one specular BxDF of that kind */
        for (int pass = 0; pass < 2; ++pass) {
            int which = -1;
            for (int i = 0; i < bsdf.nBxDFs; ++i) {
                int k = bsdf.kind[i];
                if (pass == 0 ? (k == BX_SPEC_REFL_NOOP || k == BX_SPEC_REFL_DIEL) : (k == BX_SPEC_TRANS)) { which = i; break; }
            }
            if (which < 0) continue;
            v3 wol = w2l(&bsdf, wo), wil;
            float f[NB], pdf = 0.f;
            for (int c = 0; c < NB; ++c) f[c] = 0.f;
            float fr = 1.f;
            if (bsdf.kind[which] != BX_SPEC_REFL_NOOP) fr = fresnel_dielectric(wol.z, 1.f, bsdf.ior);
            if (pass == 1) {                                   /* SpecularTransmission::Sample_f, reflection.cpp:139-162 */
                int entering = wol.z > 0.;
                float ei = 1.f, et = bsdf.ior;
                if (!entering) { float t = ei; ei = et; et = t; }
                float sini2 = sin_theta2(wol);
                float eta = ei / et;
                float sint2 = eta * eta * sini2;
                if (sint2 >= 1.) continue;                     /* total internal reflection */
                float cost = sqrtf(stdmaxf(0.f, 1.f - sint2));
                if (entering) cost = -cost;
                wil = V(eta * -wol.x, eta * -wol.y, cost);
                pdf = 1.f;
                for (int c = 0; c < NB; ++c) f[c] = (1.f - fr) * bsdf.R[which][c] / abs_cos_theta(wil);
            } else {                                           /* SpecularReflection::Sample_f, reflection.cpp:130-136 */
                wil = V(-wol.x, -wol.y, wol.z);
                pdf = 1.f;
                for (int c = 0; c < NB; ++c) f[c] = fr * bsdf.R[which][c] / abs_cos_theta(wil);
            }
            v3 wi = l2w(&bsdf, wil);
            if (!(pdf > 0.f) || is_black(f) || absdot(wi, n) == 0.f) continue;
            Ray child; child.o = p; child.d = wi; child.mint = isect.rayEpsilon; child.maxt = INFINITY;
            float Li[NB];
            direct_li(sc, strategy_all, maxDepth, depth + 1, child, NULL, smp, Li);
            float ad = absdot(wi, n);
            for (int c = 0; c < NB; ++c) L[c] += f[c] * Li[c] * ad / pdf;
        }
    }
    memcpy(Lout, L, sizeof(L));
}
static void li_sample_direct(const SptSceneDesc *sc, const SptCameraDesc *cam, int maxDepth, int spp, const float *smp, float *Lout) {
    Ray ray;
    RayDiff rd;
    camera_ray_diff(cam, smp, spp, &ray, &rd);
    direct_li(sc, 1, maxDepth, 0, ray, &rd, smp, Lout);
}
static void li_sample_direct_one(const SptSceneDesc *sc, const SptCameraDesc *cam, int maxDepth, int spp, const float *smp, float *Lout) {
    Ray ray;
    RayDiff rd;
    camera_ray_diff(cam, smp, spp, &ray, &rd);
    direct_li(sc, 0, maxDepth, 0, ray, &rd, smp, Lout);
}

int orc_sample_floats(const SptSceneDesc *sc, int32_t integrator) {
    return integrator == SPT_INTEGRATOR_DIRECT_ALL ? direct_sample_floats(sc) : (integrator == SPT_INTEGRATOR_DIRECT_ONE ? 14 : 37);
}

void orc_shade_samples(const SptSceneDesc *sc, const SptCameraDesc *cam, int32_t integrator, int32_t max_depth, int32_t spp,
                       const float *samples, const float *rng, int32_t n_rng, uint64_t n, float *out_L) {
    const int stride = orc_sample_floats(sc, integrator);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        if (integrator == SPT_INTEGRATOR_DIRECT_ALL) li_sample_direct(sc, cam, max_depth, spp, samples + (size_t)stride * i, out_L + (size_t)NB * i);
        else if (integrator == SPT_INTEGRATOR_DIRECT_ONE) li_sample_direct_one(sc, cam, max_depth, spp, samples + (size_t)stride * i, out_L + (size_t)NB * i);
        else li_sample(sc, cam, max_depth, spp, samples + 37 * i, rng ? rng + (size_t)n_rng * i : NULL, rng ? n_rng : 0,
                       out_L + (size_t)NB * i);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* radiance guards (renderers/samplerrenderer.cpp:119-133) + SpectralImageFilm::AddSample
 * (film/spectralImage.cpp:77-152) */
static void film_add(const SptFilmDesc *fd, const SptSpectralTables *tb, float imageX, float imageY, const float *Lin,
                     float *cbuf, float *wbuf) {
    float L[NB];
    memcpy(L, Lin, sizeof(L));
    int hasnan = 0;
    for (int c = 0; c < NB; ++c) if (isnan(L[c])) hasnan = 1;
    float y = spectrum_y(tb, L);
    if (hasnan || y < -1e-5 || isinf(y)) for (int c = 0; c < NB; ++c) L[c] = 0.f;
    float dimageX = imageX - 0.5f, dimageY = imageY - 0.5f;
    int x0 = (int)ceilf(dimageX - fd->filter_xwidth), x1 = (int)floorf(dimageX + fd->filter_xwidth);
    int y0 = (int)ceilf(dimageY - fd->filter_ywidth), y1 = (int)floorf(dimageY + fd->filter_ywidth);
    if (x0 < fd->x_pixel_start) x0 = fd->x_pixel_start;
    if (x1 > fd->x_pixel_start + fd->x_pixel_count - 1) x1 = fd->x_pixel_start + fd->x_pixel_count - 1;
    if (y0 < fd->y_pixel_start) y0 = fd->y_pixel_start;
    if (y1 > fd->y_pixel_start + fd->y_pixel_count - 1) y1 = fd->y_pixel_start + fd->y_pixel_count - 1;
    if ((x1 - x0) < 0 || (y1 - y0) < 0) return;
    for (int yy = y0; yy <= y1; ++yy) {
        float fy = fabsf((yy - dimageY) * fd->filter_inv_ywidth * 16);
        int iy = (int)floorf(fy); if (iy > 15) iy = 15;
        for (int xx = x0; xx <= x1; ++xx) {
            float fx = fabsf((xx - dimageX) * fd->filter_inv_xwidth * 16);
            int ix = (int)floorf(fx); if (ix > 15) ix = 15;
            float wt = fd->filter_table[iy * 16 + ix];
            size_t pix = (size_t)(yy - fd->y_pixel_start) * fd->x_pixel_count + (xx - fd->x_pixel_start);
            for (int c = 0; c < NB; ++c) cbuf[pix * NB + c] += wt * L[c];
            wbuf[pix] += wt;
        }
    }
}

void orc_film_add_samples(const SptFilmDesc *film, const SptSpectralTables *tables, const float *xy, const float *L,
                          uint64_t n, float *c, float *weight) {
    for (uint64_t i = 0; i < n; ++i) film_add(film, tables, xy[2 * i], xy[2 * i + 1], L + (size_t)NB * i, c, weight);
}

/* ------------------------------------------------------------------------------------------ */
/* The product's sampler restated. The reference draws each dimension as a scrambled (0,2)-sequence
 * in a random order per pixel from a sequential MT19937 stream (core/montecarlo.cpp:192-244,
 * core/montecarlo.h:262-315); the product keeps the sequence (VanDerCorput / Sobol2, bit-exact) and
 * replaces the stream by counter-based hashes of (seed, pixel, dimension). */
static uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
static uint32_t pixel_key(uint32_t seed, uint32_t pix) { return mix32(mix32(seed + 0x9e3779b9U) ^ (pix + 0x85ebca6bU)); }
static uint32_t dim_key(uint32_t pkey, uint32_t dim) { return mix32(pkey + dim * 0x9e3779b9U); }
/* random permutation of [0,n), n a power of two: invertible mixing on log2(n) bits */
static uint32_t permute_pow2(uint32_t i, uint32_t n, uint32_t key) {
    uint32_t mask = n - 1;
    if (!mask) return 0;
    i ^= key; i *= 0xe170893dU; i ^= key >> 16;
    i ^= (i & mask) >> 4; i ^= key >> 8; i *= 0x0929eb3fU; i ^= key >> 23;
    i ^= (i & mask) >> 1; i *= 1 | key >> 27; i *= 0x6935fa69U;
    i ^= (i & mask) >> 11; i *= 0x74dcb303U; i ^= (i & mask) >> 2; i *= 0x9e501cc3U;
    i ^= (i & mask) >> 2; i *= 0xc860a3dfU; i &= mask; i ^= i >> 5;
    return (i + key) & mask;
}
static float van_der_corput(uint32_t n, uint32_t scramble) {          /* core/montecarlo.h:270-279 */
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
    n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
    n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
    n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
    n ^= scramble;
    return stdminf(((n >> 8) & 0xffffff) / (float)(1 << 24), ONE_MINUS_EPS);
}
static float sobol2(uint32_t n, uint32_t scramble) {                  /* core/montecarlo.h:282-286 */
    for (uint32_t v = 1u << 31; n != 0; n >>= 1, v ^= v >> 1)
        if (n & 0x1) scramble ^= v;
    return stdminf(((scramble >> 8) & 0xffffff) / (float)(1 << 24), ONE_MINUS_EPS);
}
static float ld1(uint32_t pkey, uint32_t dim, uint32_t s, uint32_t spp) {
    uint32_t h = dim_key(pkey, dim);
    uint32_t idx = permute_pow2(s, spp, h);
    return van_der_corput(idx, mix32(h ^ 0x68bc21ebU));
}
static void ld2(uint32_t pkey, uint32_t dim, uint32_t s, uint32_t spp, float *out) {
    uint32_t h = dim_key(pkey, dim);
    uint32_t idx = permute_pow2(s, spp, h);
    out[0] = van_der_corput(idx, mix32(h ^ 0x68bc21ebU));
    out[1] = sobol2(idx, mix32(h ^ 0x02e5be93U));
}
void orc_gen_sample(uint64_t seed64, int32_t px, int32_t py, int32_t s, int32_t spp, float so, float sc_,
                    int32_t n_rng, float *o, float *rng) {
    uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32));
    uint32_t pkey = pixel_key(seed, ((uint32_t)py << 16) ^ (uint32_t)px);
    float t2[2];
    ld2(pkey, 0, s, spp, t2);
    o[0] = px + t2[0]; o[1] = py + t2[1];
    ld2(pkey, 1, s, spp, t2);
    o[2] = t2[0]; o[3] = t2[1];
    o[4] = lerpf(ld1(pkey, 2, s, spp), so, sc_);
    for (int k = 0; k < 14; ++k) o[5 + k] = (k < 12) ? ld1(pkey, 3 + k, s, spp) : 0.f;
    for (int k = 0; k < 9; ++k) ld2(pkey, 17 + k, s, spp, o + 19 + 2 * k);
    uint32_t rkey = mix32(pkey ^ (0x10000u + (uint32_t)s) * 0xc2b2ae35U);
    for (int k = 0; k < n_rng; ++k)
        rng[k] = (mix32(rkey + (uint32_t)k * 0x27d4eb2fU) & 0xffffff) / (float)(1 << 24);
}

/* The product's directlighting sample generator restated (csrc/sampler.cuh: direct_dims): fills the
 * 7 + 6 N floats of sample `s` of sampler pixel (px,py) in the reference's layout. */
static void gen_sample_direct(const SptSceneDesc *sc, uint64_t seed64, int32_t px, int32_t py, int32_t s, int32_t spp,
                              float so, float sc_, float *o) {
    uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32));
    uint32_t pkey = pixel_key(seed, ((uint32_t)py << 16) ^ (uint32_t)px);
    float t2[2];
    ld2(pkey, 0, s, spp, t2);
    o[0] = px + t2[0]; o[1] = py + t2[1];
    ld2(pkey, 1, s, spp, t2);
    o[2] = t2[0]; o[3] = t2[1];
    o[4] = lerpf(ld1(pkey, 2, s, spp), so, sc_);
    int N = 0;
    for (uint32_t i = 0; i < sc->n_lights; ++i) N += sc->lights[i].n_samples;
    float *oneD = o + 5, *twoD = o + 7 + 2 * N;
    o[5 + 2 * N] = o[6 + 2 * N] = 0.f;
    for (uint32_t li = 0; li < sc->n_lights; ++li) {
        int n = sc->lights[li].n_samples;
        for (int jj = 0; jj < n; ++jj)
            for (int k = 0; k < 4; ++k) {
                uint32_t h = dim_key(pkey, 64u + 4u * li + (uint32_t)k);
                uint32_t idx = permute_pow2((uint32_t)s, (uint32_t)spp, h) * (uint32_t)n +
                               permute_pow2((uint32_t)jj, (uint32_t)n, mix32(h ^ ((uint32_t)s * 0x9e3779b9U + 0x7f4a7c15U)));
                float a = van_der_corput(idx, mix32(h ^ 0x68bc21ebU));
                if (k == 0) oneD[jj] = a;
                else if (k == 1) oneD[n + jj] = a;
                else {
                    float b = sobol2(idx, mix32(h ^ 0x02e5be93U));
                    if (k == 2) { twoD[2 * jj] = a; twoD[2 * jj + 1] = b; } else { twoD[2 * n + 2 * jj] = a; twoD[2 * n + 2 * jj + 1] = b; }
                }
            }
        oneD += 2 * n; twoD += 4 * n;
    }
}

void orc_render(const SptSceneDesc *sc, const SptCameraDesc *cam, const SptFilmDesc *fd, const SptRenderParams *rp,
                float *c, float *weight) {
    const int direct = rp->integrator == SPT_INTEGRATOR_DIRECT_ALL, directOne = rp->integrator == SPT_INTEGRATOR_DIRECT_ONE;
    const int stride = directOne ? 37 : orc_sample_floats(sc, rp->integrator);
    int nrng = 11 * (rp->max_depth > 2 ? rp->max_depth - 2 : 0) + 1;
    int x1 = rp->x_end, y1 = rp->y_end;
    if (rp->skip_border) {
        if (x1 > fd->x_pixel_start + fd->x_pixel_count) x1 = fd->x_pixel_start + fd->x_pixel_count;
        if (y1 > fd->y_pixel_start + fd->y_pixel_count) y1 = fd->y_pixel_start + fd->y_pixel_count;
    }
    int ts = rp->tile_size > 0 ? rp->tile_size : 32;
    int nranks = rp->tile_nranks > 0 ? rp->tile_nranks : 1;
    int tilesX = (x1 - rp->x_start + ts - 1) / ts;
#pragma omp parallel for schedule(dynamic, 1)
    for (int py = rp->y_start; py < y1; ++py) {
        float *smp = (float *)malloc(sizeof(float) * stride * rp->spp);
        float *rng = (float *)malloc(sizeof(float) * nrng * rp->spp);
        float *L = (float *)malloc(sizeof(float) * NB * rp->spp);
        for (int px = rp->x_start; px < x1; ++px) {
            int tile = ((py - rp->y_start) / ts) * tilesX + (px - rp->x_start) / ts;
            if (tile % nranks != rp->tile_rank) continue;
            for (int s = 0; s < rp->spp; ++s) {
                if (direct) {
                    gen_sample_direct(sc, rp->seed, px, py, s, rp->spp, cam->shutter_open, cam->shutter_close, smp + stride * s);
                    li_sample_direct(sc, cam, rp->max_depth, rp->spp, smp + stride * s, L + NB * s);
                } else {
                    orc_gen_sample(rp->seed, px, py, s, rp->spp, cam->shutter_open, cam->shutter_close, nrng,
                                   smp + 37 * s, rng + nrng * s);
                    if (directOne) {
                        /* the product draws strategy "one" from the path sampler's first-bounce dimensions */
                        const float *q = smp + 37 * s;
                        float o14[14] = { q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], 0.f, 0.f, q[19], q[20], q[21], q[22] };
                        li_sample_direct_one(sc, cam, rp->max_depth, rp->spp, o14, L + NB * s);
                    } else li_sample(sc, cam, rp->max_depth, rp->spp, smp + 37 * s, rng + nrng * s, nrng, L + NB * s);
                }
            }
#pragma omp critical
            for (int s = 0; s < rp->spp; ++s) film_add(fd, &sc->tables, smp[stride * s], smp[stride * s + 1], L + NB * s, c, weight);
        }
        free(smp); free(rng); free(L);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Hooks on single functions of the restatement, for tests that compare them with the product's device code compiled
 * for the host (tests/host_shim/, tests/test_device_code_on_host.py). Same arguments as the hd_* functions there. */
void orc_measured_f(const SptSceneDesc *sc, int table, const float *wo, const float *wi, int n, float *out) {
    for (int i = 0; i < n; ++i) {
        float *o = out + (size_t)NB * i;
        for (int c = 0; c < NB; ++c) o[c] = 0.f;
        measured_f(sc, sc->brdfs + table, V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), o);
    }
}
void orc_tex_evaluate(const SptSceneDesc *sc, int tex, const float *uvd, int n, float *out) {
    const SptTexture *t = sc->textures + tex;
    for (int i = 0; i < n; ++i) {
        UVDiff df = { uvd[6 * i + 2], uvd[6 * i + 3], uvd[6 * i + 4], uvd[6 * i + 5] };
        tex_evaluate(sc, t, uvd[6 * i], uvd[6 * i + 1], &df, out + (size_t)t->channels * i);
    }
}
/* shading frame + image-mapped Kd (as RGB) at the first hit of camera samples known to hit BVH slot slot[i] */
void orc_first_vertex_frame(const SptSceneDesc *sc, const SptCameraDesc *cam, int spp, const float *samples, const uint32_t *slot,
                            const float *t, int n, float *out) {
    (void)t;
    for (int i = 0; i < n; ++i) {
        Ray ray; RayDiff rd;
        camera_ray_diff(cam, samples + 5 * (size_t)i, spp, &ray, &rd);
        uint32_t s; Hit isect;
        float *o = out + 12 * (size_t)i;
        for (int k = 0; k < 12; ++k) o[k] = 0.f;
        if (!bvh_intersect(sc, &ray, 0, &s, &isect, NULL, NULL) || s != slot[i]) continue;
        BSDF b; v3 n_s;
        make_bsdf(sc, s, &isect, &rd, &b, &n_s);
        o[0] = b.nn.x; o[1] = b.nn.y; o[2] = b.nn.z; o[3] = b.sn.x; o[4] = b.sn.y; o[5] = b.sn.z;
        o[6] = b.tn.x; o[7] = b.tn.y; o[8] = b.tn.z;
        const SptMaterial *m = sc->materials + sc->prim_material[s];
        if (m->tex_kd >= 0) {
            UVDiff df;
            compute_differentials(&isect, &rd, &df);
            tex_evaluate(sc, sc->textures + m->tex_kd, isect.u, isect.v, &df, o + 9);
        }
    }
}

void orc_light_sample(const SptSceneDesc *sc, int light, const float *p, const float *u, int n, float *out) {
    for (int i = 0; i < n; ++i) {
        LightSampleResult lr;
        light_sample(sc, sc->lights + light, V(p[3 * i], p[3 * i + 1], p[3 * i + 2]), 0.f, u[3 * i], u[3 * i + 1], u[3 * i + 2], &lr);
        float *o = out + 9 * (size_t)i;
        o[0] = lr.wi.x; o[1] = lr.wi.y; o[2] = lr.wi.z; o[3] = lr.pdf;
        o[4] = lr.shadow.d.x; o[5] = lr.shadow.d.y; o[6] = lr.shadow.d.z; o[7] = lr.shadow.maxt; o[8] = is_black(lr.Li) ? 1.f : 0.f;
    }
}
void orc_light_pdf(const SptSceneDesc *sc, int light, const float *p, const float *w, int n, float *out) {
    for (int i = 0; i < n; ++i)
        out[i] = light_pdf(sc, sc->lights + light, V(p[3 * i], p[3 * i + 1], p[3 * i + 2]), V(w[3 * i], w[3 * i + 1], w[3 * i + 2]));
}

int orc_nbands(void) { return NB; }

// shade.cuh — device code for K4-K6: shading frame, BSDF terms, light sampling, MIS.
//
// The reference evaluates every BSDF/light quantity as a 32-float SampledSpectrum
// (src/core/spectrum.h:93-266). Here a direction's BSDF value is kept as a handful of SCALAR
// terms (DirTerms) — everything that does not depend on wavelength — and the per-band value is
// rebuilt inside the one band loop of the accumulate kernel from the material row
// (f_band()), in the reference's own operation order. A path therefore stages ~200 bytes
// between kernels instead of 3 x 128-byte spectra.
#pragma once
#include "traverse.cuh"
#include "montecarlo.cuh"

// ---- BSDF -------------------------------------------------------------------------------------
enum { BX_LAMBERT = 0, BX_ORENNAYAR = 1, BX_MF_DIEL = 2, BX_MF_COND = 3 };
struct Bsdf {
    v3 nn, sn, tn, ng;
    int mtype;            // SPT_MAT_*
    bool orenNayar;
    float exponent, A, B;
    float ior;            // GLASS: FresnelDielectric(1, ior)
    int compMask;         // MIRROR / GLASS: bit 0 the reflection component exists (Kr not black), bit 1 the transmission (Kt)
    float ex, ey;         // SUBSTRATE: Anisotropic exponents (reflection.h:433-437)
    int brdf;             // MEASURED: row of DevScene::brdfs
    bool texKd;           // Kd came from an image texture: kd_rgb replaces the material row's spec0
    float kd_rgb[3];
};
// Wavelength-independent factors of BSDF::f(wo,wi) for one direction.
//   Oren-Nayar: a0 = A + B*maxcos*sinalpha*tanbeta
//   Microfacet: a0 = D, a1 = G, a2 = F (dielectric) or |cos theta_h| (conductor), a3 = 4 cosI cosO
//   FresnelBlend: a0 = diffuse scalar, a1 = D / (4 |wi.wh| max(cosI, cosO)), a2 = (1 - wi.wh)^5
struct DirTerms { float a0, a1, a2, a3; bool reflect, mf; };

__device__ __forceinline__ v3 w2l(const Bsdf &b, v3 v) { return V(dot(v, b.sn), dot(v, b.tn), dot(v, b.nn)); }
__device__ __forceinline__ v3 l2w(const Bsdf &b, v3 v) {
    return V(b.sn.x * v.x + b.tn.x * v.y + b.nn.x * v.z, b.sn.y * v.x + b.tn.y * v.y + b.nn.y * v.z,
             b.sn.z * v.x + b.tn.z * v.y + b.nn.z * v.z);
}
__device__ __forceinline__ float abs_cos_theta(v3 w) { return fabsf(w.z); }
__device__ __forceinline__ bool same_hemisphere(v3 w, v3 wp) { return w.z * wp.z > 0.f; }
__device__ __forceinline__ float sin_theta(v3 w) { return sqrtf(stdmaxf(0.f, 1.f - w.z * w.z)); }
__device__ __forceinline__ float cos_phi(v3 w) { float s = sin_theta(w); if (s == 0.f) return 1.f; return clampf(w.x / s, -1.f, 1.f); }
__device__ __forceinline__ float sin_phi(v3 w) { float s = sin_theta(w); if (s == 0.f) return 0.f; return clampf(w.y / s, -1.f, 1.f); }
__device__ __forceinline__ float blinn_exponent(float e) { if (e > 10000.f || isnan(e)) e = 10000.f; return e; }


// ---- image textures -----------------------------------------------------------
// The camera ray's offset rays (perspective.cpp:96-101) after ScaleDifferentials (geometry.h:368-373)
struct RayDiff { v3 rxo, ryo, rxd, ryd; };
struct UVDiff { float dudx, dvdx, dudy, dvdy; };
__device__ __forceinline__ int imod(int a, int b) { int n = (int)(a / b); a -= n * b; if (a < 0) a += b; return a; }

__device__ inline void camera_ray_diff(const SptCameraDesc &cam, float imageX, float imageY, float scale, const Ray &ray, RayDiff *rd) {
    v3 Pcamera = xf_point(cam.raster_to_camera, V(imageX, imageY, 0.f));
    v3 rxd = normalize(vadd(Pcamera, V(cam.dx_camera[0], cam.dx_camera[1], cam.dx_camera[2])));
    v3 ryd = normalize(vadd(Pcamera, V(cam.dy_camera[0], cam.dy_camera[1], cam.dy_camera[2])));
    rxd = xf_vector(cam.camera_to_world, rxd);
    ryd = xf_vector(cam.camera_to_world, ryd);
    // rxOrigin = ryOrigin = the ray's own origin (also with a lens): o + (o - o) * s = o
    rd->rxo = rd->ryo = ray.o;
    rd->rxd = vadd(ray.d, vmul(vsub(rxd, ray.d), scale));
    rd->ryd = vadd(ray.d, vmul(vsub(ryd, ray.d), scale));
}
__device__ __forceinline__ bool solve2x2(const float A[2][2], const float B[2], float *x0, float *x1) {   // transform.cpp:31-41
    float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    if (fabsf(det) < 1e-10f) return false;
    *x0 = (A[1][1] * B[0] - A[0][1] * B[1]) / det;
    *x1 = (A[0][0] * B[1] - A[1][0] * B[0]) / det;
    return !(isnan(*x0) || isnan(*x1));
}
__device__ __forceinline__ float vcomp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// DifferentialGeometry::ComputeDifferentials, diffgeom.cpp:50-107
__device__ inline void compute_differentials(const Hit &dg, const RayDiff *rd, UVDiff *o) {
    o->dudx = o->dvdx = o->dudy = o->dvdy = 0.f;
    if (!rd) return;
    v3 nn = dg.nn, p = dg.p;
    float d = -dot(nn, p);
    float tx = -(dot(nn, rd->rxo) + d) / dot(nn, rd->rxd);
    if (isnan(tx)) return;
    v3 px = vadd(rd->rxo, vmul(rd->rxd, tx));
    float ty = -(dot(nn, rd->ryo) + d) / dot(nn, rd->ryd);
    if (isnan(ty)) return;
    v3 py = vadd(rd->ryo, vmul(rd->ryd, ty));
    int a0, a1;
    if (fabsf(nn.x) > fabsf(nn.y) && fabsf(nn.x) > fabsf(nn.z)) { a0 = 1; a1 = 2; }
    else if (fabsf(nn.y) > fabsf(nn.z)) { a0 = 0; a1 = 2; }
    else { a0 = 0; a1 = 1; }
    float A[2][2], Bx[2], By[2];
    A[0][0] = vcomp(dg.dpdu, a0); A[0][1] = vcomp(dg.dpdv, a0);
    A[1][0] = vcomp(dg.dpdu, a1); A[1][1] = vcomp(dg.dpdv, a1);
    Bx[0] = vcomp(px, a0) - vcomp(p, a0); Bx[1] = vcomp(px, a1) - vcomp(p, a1);
    By[0] = vcomp(py, a0) - vcomp(p, a0); By[1] = vcomp(py, a1) - vcomp(p, a1);
    if (!solve2x2(A, Bx, &o->dudx, &o->dvdx)) { o->dudx = 0.f; o->dvdx = 0.f; }
    if (!solve2x2(A, By, &o->dudy, &o->dvdy)) { o->dudy = 0.f; o->dvdy = 0.f; }
}
__device__ __forceinline__ float log2_pbrt(float x) { return logf(x) * (1.f / 0.693147180559945309417f); }   // pbrt.h:243-246

struct TexLevel { const float *texels; int w, h; };
__device__ inline TexLevel tex_level(const DevScene &sc, const SptTexture &t, int level) {
    TexLevel l; l.texels = sc.tex_texels + t.texel_offset; l.w = t.width; l.h = t.height;
    for (int i = 0; i < level; ++i) {
        l.texels += (size_t)l.w * l.h * t.channels;
        l.w = max(l.w / 2, 1); l.h = max(l.h / 2, 1);
    }
    return l;
}
// MIPMap::Texel, mipmap.h:177-201. C = channels per texel (3: RGB, 1: float)
template <int C> __device__ inline void tex_texel(const SptTexture &t, const TexLevel &l, int s, int tt, float *out) {
    // MIP levels are powers of two (the MIPMap constructor resamples, mipmap.h:124-170; checked by spt_scene_create), so
    // Mod(s, w) - including its wrap of negative s - is a mask: no integer divisions in the EWA loops
    if (t.wrap == SPT_WRAP_REPEAT) { s &= l.w - 1; tt &= l.h - 1; }
    else if (t.wrap == SPT_WRAP_CLAMP) { s = clampi(s, 0, l.w - 1); tt = clampi(tt, 0, l.h - 1); }
    else if (s < 0 || s >= l.w || tt < 0 || tt >= l.h) {
#pragma unroll
        for (int k = 0; k < C; ++k) out[k] = 0.f;
        return;
    }
    const float *px = l.texels + ((size_t)tt * l.w + s) * C;
#pragma unroll
    for (int k = 0; k < C; ++k) out[k] = __ldg(px + k);
}
// MIPMap::triangle, mipmap.h:258-270
template <int C> __device__ inline void tex_triangle(const DevScene &sc, const SptTexture &t, int level, float s, float tt, float *out) {
    level = clampi(level, 0, t.n_levels - 1);
    TexLevel l = tex_level(sc, t, level);
    s = s * l.w - 0.5f;
    tt = tt * l.h - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(tt);
    float ds = s - s0, dt = tt - t0;
    float a[C], b[C], c[C], d[C];
    tex_texel<C>(t, l, s0, t0, a); tex_texel<C>(t, l, s0, t0 + 1, b);
    tex_texel<C>(t, l, s0 + 1, t0, c); tex_texel<C>(t, l, s0 + 1, t0 + 1, d);
#pragma unroll
    for (int k = 0; k < C; ++k)
        out[k] = a[k] * ((1.f - ds) * (1.f - dt)) + b[k] * ((1.f - ds) * dt) + c[k] * (ds * (1.f - dt)) + d[k] * (ds * dt);
}
// MIPMap::EWA, mipmap.h:322-377
// (inlined on purpose: __noinline__ copies of the filters were measured 8 % slower on config 3, although ncu shows the
// kernel waiting on instruction fetch - "no instruction" 5.9 stall cycles per issue at the first vertex)
template <int C> __device__ inline void tex_ewa(const DevScene &sc, const SptTexture &t, int level, float s, float tt,
                                               float ds0, float dt0, float ds1, float dt1, float *out) {
    if (level >= t.n_levels) { TexLevel top = tex_level(sc, t, t.n_levels - 1); tex_texel<C>(t, top, 0, 0, out); return; }
    TexLevel l = tex_level(sc, t, level);
    s = s * l.w - 0.5f;
    tt = tt * l.h - 0.5f;
    ds0 *= l.w; dt0 *= l.h; ds1 *= l.w; dt1 *= l.h;
    float A = dt0 * dt0 + dt1 * dt1 + 1;
    float B = -2.f * (ds0 * dt0 + ds1 * dt1);
    float Cc = ds0 * ds0 + ds1 * ds1 + 1;
    float invF = 1.f / (A * Cc - B * B * 0.25f);
    A *= invF; B *= invF; Cc *= invF;
    float det = -B * B + 4.f * A * Cc;
    float invDet = 1.f / det;
    float uSqrt = sqrtf(det * Cc), vSqrt = sqrtf(A * det);
    int s0 = (int)ceilf(s - 2.f * invDet * uSqrt), s1 = (int)floorf(s + 2.f * invDet * uSqrt);
    int t0 = (int)ceilf(tt - 2.f * invDet * vSqrt), t1 = (int)floorf(tt + 2.f * invDet * vSqrt);
    float sum[C], sumWts = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) sum[k] = 0.f;
    for (int it = t0; it <= t1; ++it) {
        float ttt = it - tt;
        for (int is = s0; is <= s1; ++is) {
            float ss = is - s;
            float r2 = A * ss * ss + B * ss * ttt + Cc * ttt * ttt;
            if (r2 < 1.f) {
                float weight = __ldg(sc.ewa_lut + min((int)(r2 * 128), 127));
                float tx[C];
                tex_texel<C>(t, l, is, it, tx);
#pragma unroll
                for (int k = 0; k < C; ++k) sum[k] += tx[k] * weight;
                sumWts += weight;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < C; ++k) out[k] = sum[k] / sumWts;
}
// MIPMap::Lookup(s,t,width), mipmap.h:215-255
template <int C> __device__ inline void tex_lookup_tri(const DevScene &sc, const SptTexture &t, float s, float tt, float width, float *out) {
    if (t.no_filter) {
        TexLevel l = tex_level(sc, t, 0);
        s = s * t.width - 0.5f;
        tt = tt * t.height - 0.5f;
        tex_texel<C>(t, l, (int)floorf(s + 0.5f), (int)floorf(tt + 0.5f), out);
        return;
    }
    float level = t.n_levels - 1 + log2_pbrt(stdmaxf(width, 1e-8f));
    if (level < 0) tex_triangle<C>(sc, t, 0, s, tt, out);
    else if (level >= t.n_levels - 1) { TexLevel top = tex_level(sc, t, t.n_levels - 1); tex_texel<C>(t, top, 0, 0, out); }
    else {
        int iLevel = (int)floorf(level);
        float delta = level - iLevel;
        float a[C], b[C];
        tex_triangle<C>(sc, t, iLevel, s, tt, a);
        tex_triangle<C>(sc, t, iLevel + 1, s, tt, b);
#pragma unroll
        for (int k = 0; k < C; ++k) out[k] = a[k] * (1.f - delta) + b[k] * delta;
    }
}
// ImageTexture::Evaluate (imagemap.cpp:88-97): UVMapping2D::Map (texture.cpp:80-90), then MIPMap::Lookup with
// differentials (mipmap.h:273-319); a float image is multiplied by its ScaleTexture constant (scale.h:45-47)
template <int C> __device__ inline void tex_evaluate(const DevScene &sc, const SptTexture &t, float u, float v, const UVDiff &df, float *out) {
    float s = t.su * u + t.du, tt = t.sv * v + t.dv;
    float ds0 = t.su * df.dudx, dt0 = t.sv * df.dvdx, ds1 = t.su * df.dudy, dt1 = t.sv * df.dvdy;
    if (t.trilinear || t.no_filter) {
        tex_lookup_tri<C>(sc, t, s, tt, 2.f * stdmaxf(stdmaxf(fabsf(ds0), fabsf(dt0)), stdmaxf(fabsf(ds1), fabsf(dt1))), out);
    } else {
        if (ds0 * ds0 + dt0 * dt0 < ds1 * ds1 + dt1 * dt1) {
            float tmp = ds0; ds0 = ds1; ds1 = tmp;
            tmp = dt0; dt0 = dt1; dt1 = tmp;
        }
        float majorLength = sqrtf(ds0 * ds0 + dt0 * dt0);
        float minorLength = sqrtf(ds1 * ds1 + dt1 * dt1);
        if (minorLength * t.max_aniso < majorLength && minorLength > 0.f) {
            float scale = majorLength / (minorLength * t.max_aniso);
            ds1 *= scale; dt1 *= scale; minorLength *= scale;
        }
        if (minorLength == 0.f) tex_triangle<C>(sc, t, 0, s, tt, out);
        else {
            float lod = stdmaxf(0.f, t.n_levels - 1.f + log2_pbrt(minorLength));
            int ilod = (int)floorf(lod);
            float d = lod - ilod;
            float a[C], b[C];
            tex_ewa<C>(sc, t, ilod, s, tt, ds0, dt0, ds1, dt1, a);
            // d == 0 (the usual case at high sample counts: footprints below a texel clamp lod to 0): a*1 + b*0 = a
            // exactly - b is finite, the filter ellipse has semi-axes >= 1 texel (covariance J J^T + I) and so always
            // covers a texel centre - and the second level need not be filtered
            if (d == 0.f) {
#pragma unroll
                for (int k = 0; k < C; ++k) out[k] = a[k];
            } else {
                tex_ewa<C>(sc, t, ilod + 1, s, tt, ds0, dt0, ds1, dt1, b);
#pragma unroll
                for (int k = 0; k < C; ++k) out[k] = a[k] * (1.f - d) + b[k] * d;
            }
        }
    }
    if (C == 1) out[0] = out[0] * t.scale;
}

// Triangle::GetShadingGeometry (trianglemesh.cpp:285-360) + Material::Bump with the constant-0
// displacement every reference material carries (material.cpp:39-82) + BSDF frame
// (reflection.cpp:593-601) + the BxDF set of matte/plastic/metal (materials/*.cpp).
// EXT (scenes with a substrate / image-mapped Kd / bump-mapped material): Intersection::GetBSDF's
// ComputeDifferentials (rd: the camera ray's offset rays at the first vertex, NULL afterwards), Bump with a float
// image map, Kd->Evaluate(dgs).
template <bool EXT>
__device__ inline void make_bsdf(const DevScene &sc, uint32_t slot, const Hit &dg, const RayDiff *rd, Bsdf *b) {
    int flags = sc.prim_flags[slot];
    const SptMaterial &m = sc.materials[sc.prim_material[slot]];
    UVDiff df; df.dudx = df.dvdx = df.dudy = df.dvdy = 0.f;
    const bool bump = EXT && m.tex_bump >= 0;
    if (EXT && (m.tex_kd >= 0 || m.tex_bump >= 0)) compute_differentials(dg, rd, &df);
    v3 s_nn = dg.nn, dndu = V(0, 0, 0), dndv = V(0, 0, 0);
    v3 s_dpdu = dg.dpdu, s_dpdv = dg.dpdv;
    if (sc.prim_kind[slot] == SPT_PRIM_TRIANGLE && (flags & SPT_PF_HAS_N)) {
        const int32_t *vi = sc.tri_vidx + 3 * (size_t)sc.prim_data[slot];
        float uv[3][2];
        tri_uvs(sc, flags, vi, uv);
        float A00 = uv[1][0] - uv[0][0], A01 = uv[2][0] - uv[0][0];
        float A10 = uv[1][1] - uv[0][1], A11 = uv[2][1] - uv[0][1];
        float C0 = dg.u - uv[0][0], C1 = dg.v - uv[0][1];
        float bb0, bb1, bb2;
        float det = A00 * A11 - A01 * A10;                          // SolveLinearSystem2x2, transform.cpp:31-41
        bool ok = true;
        if (fabsf(det) < 1e-10f) ok = false;
        else {
            bb1 = (A11 * C0 - A01 * C1) / det;
            bb2 = (A00 * C1 - A10 * C0) / det;
            if (isnan(bb1) || isnan(bb2)) ok = false;
        }
        if (!ok) bb0 = bb1 = bb2 = 1.f / 3.f;
        else bb0 = 1.f - bb1 - bb2;
        const SptXform &xf = sc.xforms[sc.prim_xform[slot]];
        const float *n0 = sc.N + 3 * (size_t)vi[0], *n1 = sc.N + 3 * (size_t)vi[1], *n2 = sc.N + 3 * (size_t)vi[2];
        v3 nsum = vadd(vadd(vmul(V(n0[0], n0[1], n0[2]), bb0), vmul(V(n1[0], n1[1], n1[2]), bb1)),
                       vmul(V(n2[0], n2[1], n2[2]), bb2));
        v3 ns = normalize(xf_normal(xf.minv, nsum));
        v3 ss = normalize(dg.dpdu);
        v3 ts = cross(ss, ns);
        if (len2(ts) > 0.f) { ts = normalize(ts); ss = cross(ts, ns); }
        else coordinate_system(ns, &ss, &ts);
        s_dpdu = ss; s_dpdv = ts;
        if (bump) {
            // dndu, dndv (trianglemesh.cpp:331-351) and the shading DifferentialGeometry's own normal (diffgeom.cpp:32-47)
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
            dndu = xf_normal(xf.minv, dndu);
            dndv = xf_normal(xf.minv, dndv);
            s_nn = normalize(cross(ss, ts));
            if (flags & SPT_PF_FLIP_NORMAL) s_nn = vmul(s_nn, -1.f);
        }
    }
    if (bump) {                                                      // Material::Bump, material.cpp:39-82
        const SptTexture &bt = sc.textures[m.tex_bump];
        float du = .5f * (fabsf(df.dudx) + fabsf(df.dudy));
        if (du == 0.f) du = .01f;
        float dv = .5f * (fabsf(df.dvdx) + fabsf(df.dvdy));
        if (dv == 0.f) dv = .01f;
        float disp[3];
#pragma unroll 1
        for (int k = 0; k < 3; ++k)                                  // one copy of the filtered lookup in the instruction cache
            tex_evaluate<1>(sc, bt, dg.u + (k == 0 ? du : 0.f), dg.v + (k == 1 ? dv : 0.f), df, &disp[k]);
        const float uDisplace = disp[0], vDisplace = disp[1], displace = disp[2];
        s_dpdu = vadd(vadd(s_dpdu, vmul(s_nn, (uDisplace - displace) / du)), vmul(dndu, displace));
        s_dpdv = vadd(vadd(s_dpdv, vmul(s_nn, (vDisplace - displace) / dv)), vmul(dndv, displace));
    }
    v3 nn = normalize(cross(s_dpdu, s_dpdv));
    if (flags & SPT_PF_FLIP_NORMAL) nn = vmul(nn, -1.f);
    if (dot(nn, dg.nn) < 0.f) nn = vneg(nn);                        // Faceforward to the geometric normal
    b->nn = nn;
    b->ng = dg.nn;
    b->sn = normalize(s_dpdu);
    b->tn = cross(b->nn, b->sn);
    b->mtype = m.type;
    b->orenNayar = false; b->exponent = 0.f; b->A = b->B = 0.f; b->ior = 1.f; b->compMask = 0;
    b->ex = b->ey = 0.f; b->texKd = false; b->brdf = m.brdf;
    if (EXT && m.tex_kd >= 0) {                                      // Kd->Evaluate(dgs): imagemap.cpp:88-97
        tex_evaluate<3>(sc, sc.textures[m.tex_kd], dg.u, dg.v, df, b->kd_rgb);
        b->texKd = true;
    }
    if (EXT && m.type == SPT_MAT_SUBSTRATE) {                        // substrate.cpp:34-56, reflection.h:433-437
        b->ex = 1.f / m.p0; b->ey = 1.f / m.p1;
        if (b->ex > 10000.f || isnan(b->ex)) b->ex = 10000.f;
        if (b->ey > 10000.f || isnan(b->ey)) b->ey = 10000.f;
    } else if (m.type == SPT_MAT_MIRROR || m.type == SPT_MAT_GLASS) {       // mirror.cpp:34-55, glass.cpp:34-58
        b->ior = m.p0;
        b->compMask = (int)m.p1;                                      // which spectra are not black: set by spt_scene_create
    } else if (m.type == SPT_MAT_MATTE) {
        if (m.p0 != 0.f) {                                           // OrenNayar ctor, reflection.h:363-370
            b->orenNayar = true;
            float sigma = (PI_F / 180.f) * m.p0;
            float sigma2 = sigma * sigma;
            b->A = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
            b->B = 0.45f * sigma2 / (sigma2 + 0.09f);
        }
    } else {
        b->exponent = blinn_exponent(1.f / m.p0);
    }
}
__device__ __forceinline__ bool bsdf_is_specular(const Bsdf &b) { return b.mtype == SPT_MAT_MIRROR || b.mtype == SPT_MAT_GLASS; }
__device__ __forceinline__ int bsdf_ncomp(const Bsdf &b) { return b.mtype == SPT_MAT_PLASTIC ? 2 : 1; }     // non-specular materials
__device__ __forceinline__ int specular_ncomp(const Bsdf &b) { return (b.compMask & 1) + ((b.compMask >> 1) & 1); }

__device__ inline float fresnel_dielectric(float cosi, float eta_i, float eta_t) {          // reflection.cpp:107-127,52-60
    cosi = clampf(cosi, -1.f, 1.f);
    bool entering = cosi > 0.f;
    float ei = eta_i, et = eta_t;
    if (!entering) { float tmp = ei; ei = et; et = tmp; }
    float sint = ei / et * sqrtf(stdmaxf(0.f, 1.f - cosi * cosi));
    if (sint >= 1.f) return 1.f;
    float cost = sqrtf(stdmaxf(0.f, 1.f - sint * sint));
    float ac = fabsf(cosi);
    float Rparl = ((et * ac) - (ei * cost)) / ((et * ac) + (ei * cost));
    float Rperp = ((ei * ac) - (et * cost)) / ((ei * ac) + (et * cost));
    return (Rparl * Rparl + Rperp * Rperp) / 2.f;
}
__device__ __forceinline__ float fr_cond(float cosi, float eta, float k) {                   // reflection.cpp:63-71
    float tmp = (eta * eta + k * k) * cosi * cosi;
    float Rparl2 = (tmp - (2.f * eta * cosi) + 1) / (tmp + (2.f * eta * cosi) + 1);
    float tmp_f = eta * eta + k * k;
    float Rperp2 = (tmp_f - (2.f * eta * cosi) + cosi * cosi) / (tmp_f + (2.f * eta * cosi) + cosi * cosi);
    return (Rparl2 + Rperp2) / 2.f;
}

// x^e for x in [0,1], e >= 0 as exp2(e*log2(x)): ~35 instructions instead of powf's ~100. log2f is
// accurate to 1 ulp, so the error of the exponent argument is e*|log2 x|*6e-8: negligible wherever the
// result is not vanishing (radiance tests compare at 2e-4 relative).
__device__ __forceinline__ float pow01(float x, float e) { return exp2f(e * log2f(x)); }
// x^5 (the Schlick-style terms of FresnelBlend, reflection.cpp:224-236): three products, within 2 ulp of the reference's powf(x, 5)
__device__ __forceinline__ float pow5(float x) { const float x2 = x * x; return x2 * x2 * x; }

// scalar part of BSDF::f (reflection.cpp:604-618, with Lambertian :165-167, OrenNayar :170-193,
// Microfacet :203-214, G reflection.h:395-402, Blinn::D reflection.h:419-422) and, from the same
// half-vector and the same cos^e, BSDF::Pdf (reflection.cpp:575-590: BxDF::Pdf :312-315,
// Microfacet::Pdf :331-335, Blinn::Pdf :356-366) averaged over the components.
template <bool EXT>
__device__ inline void bsdf_terms(const Bsdf &b, v3 woW, v3 wiW, v3 wo, v3 wi, DirTerms *t, float *pdfAll) {
    t->a0 = t->a1 = t->a2 = t->a3 = 0.f;
    t->mf = false;
    t->reflect = dot(wiW, b.ng) * dot(woW, b.ng) > 0;
    bool same = same_hemisphere(wo, wi);
    float cosPdf = same ? abs_cos_theta(wi) * INV_PI_F : 0.f;
    if (b.mtype == SPT_MAT_MATTE) {
        *pdfAll = cosPdf / 1;
        if (t->reflect && b.orenNayar) {
            float sinthetai = sin_theta(wi), sinthetao = sin_theta(wo);
            float maxcos = 0.f;
            if ((double)sinthetai > 1e-4 && (double)sinthetao > 1e-4) {
                float sinphii = sin_phi(wi), cosphii = cos_phi(wi);
                float sinphio = sin_phi(wo), cosphio = cos_phi(wo);
                float dcos = cosphii * cosphio + sinphii * sinphio;
                maxcos = stdmaxf(0.f, dcos);
            }
            float sinalpha, tanbeta;
            if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sinalpha = sinthetao; tanbeta = sinthetai / abs_cos_theta(wi); }
            else { sinalpha = sinthetai; tanbeta = sinthetao / abs_cos_theta(wo); }
            t->a0 = (b.A + b.B * maxcos * sinalpha * tanbeta);
        }
        return;
    }
    if (EXT && b.mtype == SPT_MAT_MEASURED) {                        // BxDF::Pdf (reflection.cpp:312-315); f is a table look-up (measured_f)
        *pdfAll = cosPdf;
        if (t->reflect) { t->a0 = 1.f; t->mf = true; }
        return;
    }
    if (EXT && b.mtype == SPT_MAT_SUBSTRATE) {                       // FresnelBlend::f / ::Pdf, reflection.cpp:224-236,453-456
        v3 wh = vadd(wi, wo);
        const bool whZero = (wh.x == 0.f && wh.y == 0.f && wh.z == 0.f);
        wh = normalize(wh);
        const float costhetah = abs_cos_theta(wh);
        const float ds = 1.f - costhetah * costhetah;
        const float e = (b.ex * wh.x * wh.x + b.ey * wh.y * wh.y) / ds;
        const float pw = pow01(costhetah, e);            // unused when ds == 0 (e = 0/0)
        float anisoPdf = 0.f;                                        // Anisotropic::Pdf, reflection.cpp:420-432
        if (ds > 0.f && dot(wo, wh) > 0.f) anisoPdf = (sqrtf((b.ex + 1.f) * (b.ey + 1.f)) * INV_TWOPI_F * pw) / (4.f * dot(wo, wh));
        *pdfAll = same ? .5f * (abs_cos_theta(wi) * INV_PI_F + anisoPdf) : 0.f;
        if (!t->reflect || whZero) return;
        t->a0 = (28.f / (23.f * PI_F)) * (1.f - pow5(1.f - .5f * abs_cos_theta(wi))) * (1.f - pow5(1.f - .5f * abs_cos_theta(wo)));
        const float D = ds == 0.f ? 0.f : sqrtf((b.ex + 2.f) * (b.ey + 2.f)) * INV_TWOPI_F * pw;      // Anisotropic::D, reflection.h:438-444
        t->a1 = D / (4.f * absdot(wi, wh) * stdmaxf(abs_cos_theta(wi), abs_cos_theta(wo)));
        t->a2 = pow5(1 - dot(wi, wh));
        t->mf = true;
        return;
    }
    // microfacet component: one half-vector, one cos^e shared by D and the Blinn pdf
    v3 wh = vadd(wi, wo);
    bool whZero = (wh.x == 0.f && wh.y == 0.f && wh.z == 0.f);
    wh = normalize(wh);
    float cosH = abs_cos_theta(wh);
    float pw = pow01(cosH, b.exponent);
    float woh = dot(wo, wh);
    float blinnPdf = 0.f;
    if (same) {
        blinnPdf = ((b.exponent + 1.f) * pw) / (2.f * PI_F * 4.f * woh);
        if (woh <= 0.f) blinnPdf = 0.f;
    }
    *pdfAll = (b.mtype == SPT_MAT_PLASTIC) ? (cosPdf + blinnPdf) / 2 : blinnPdf / 1;
    if (!t->reflect) return;
    float cosThetaO = abs_cos_theta(wo), cosThetaI = abs_cos_theta(wi);
    if (cosThetaI == 0.f || cosThetaO == 0.f || whZero) return;
    float cosThetaH = dot(wi, wh);
    t->a0 = (b.exponent + 2) * INV_TWOPI_F * pw;
    float WOdotWh = fabsf(woh);
    t->a1 = stdminf(1.f, stdminf((2.f * cosH * cosThetaO / WOdotWh), (2.f * cosH * cosThetaI / WOdotWh)));
    t->a2 = (b.mtype == SPT_MAT_PLASTIC) ? fresnel_dielectric(cosThetaH, 1.5f, 1.f) : fabsf(cosThetaH);
    t->a3 = (4.f * cosThetaI * cosThetaO);
    t->mf = true;
}
// f for band c from the staged terms, in the reference's operation order (used by the parity
// helpers; the accumulate kernel folds the scalar factors once per vertex instead)
__device__ __forceinline__ float f_band(const SptMaterial &m, bool orenNayar, const DirTerms &t, int c) {
    if (!t.reflect) return 0.f;
    float f = 0.f;
    if (m.type == SPT_MAT_MATTE) {
        if (orenNayar) f += m.spec0[c] * INV_PI_F * t.a0;
        else f += m.spec0[c] * INV_PI_F;
    } else if (m.type == SPT_MAT_PLASTIC) {
        f += m.spec0[c] * INV_PI_F;
        if (t.mf) f += m.spec1[c] * t.a0 * t.a1 * t.a2 / t.a3;
    } else {
        if (t.mf) f += 1.f * t.a0 * t.a1 * fr_cond(t.a2, m.spec0[c], m.spec1[c]) / t.a3;
    }
    return f;
}
// BSDF::Sample_f (reflection.cpp:514-572), first half: the component is chosen and its direction
// sampled (local frame). Returns false when the chosen component yields no sample (pdf == 0).
// Anisotropic::sampleFirstQuadrant, reflection.cpp:408-417
__device__ inline void aniso_first_quadrant(const Bsdf &b, float u1, float u2, float *phi, float *costheta) {
    if (b.ex == b.ey) *phi = PI_F * u1 * 0.5f;
    else *phi = atanf(sqrtf((b.ex + 1.f) / (b.ey + 1.f)) * tanf(PI_F * u1 * 0.5f));
    float cosphi, sinphi;
    sin_cos(*phi, &sinphi, &cosphi);
    *costheta = pow01(u2, 1.f / (b.ex * cosphi * cosphi + b.ey * sinphi * sinphi + 1));
}
// pdfOverride (substrate only): FresnelBlend::Sample_f returns before `*pdf = Pdf(wo, *wi)` when the direction
// sampled from the microfacet distribution falls in the other hemisphere (reflection.cpp:444-446), leaving the
// distribution's own pdf in place; < 0 otherwise.
template <bool EXT>
__device__ inline bool bsdf_sample_dir(const Bsdf &b, v3 wo, float uComp, float u1, float u2, v3 *wiOut, float *pdfOverride) {
    *pdfOverride = -1.f;
    if (EXT && b.mtype == SPT_MAT_SUBSTRATE) {                              // FresnelBlend::Sample_f, reflection.cpp:435-450
        v3 wi;
        if (u1 < .5f) {
            u1 = 2.f * u1;
            wi = cosine_sample_hemisphere(u1, u2);
            if (wo.z < 0.f) wi.z *= -1.f;
        } else {                                                     // Anisotropic::Sample_f, reflection.cpp:369-405
            u1 = 2.f * (u1 - .5f);
            float phi, costheta;
            // the quadrant's own u1, ONE copy of the first-quadrant code, then the quadrant's reflection of phi
            const int quad = u1 < .25f ? 0 : (u1 < .5f ? 1 : (u1 < .75f ? 2 : 3));
            const float uq = quad == 0 ? 4.f * u1 : (quad == 1 ? 4.f * (.5f - u1) : (quad == 2 ? 4.f * (u1 - .5f) : 4.f * (1.f - u1)));
            aniso_first_quadrant(b, uq, u2, &phi, &costheta);
            if (quad == 1) phi = PI_F - phi; else if (quad == 2) phi += PI_F; else if (quad == 3) phi = 2.f * PI_F - phi;
            float sintheta = sqrtf(stdmaxf(0.f, 1.f - costheta * costheta));
            float sph, cph;
            sin_cos(phi, &sph, &cph);
            v3 wh = V(sintheta * cph, sintheta * sph, costheta);
            if (!same_hemisphere(wo, wh)) wh = vneg(wh);
            wi = vadd(vneg(wo), vmul(wh, 2.f * dot(wo, wh)));
            if (!same_hemisphere(wo, wi)) {
                const float costhetah = abs_cos_theta(wh), ds = 1.f - costhetah * costhetah;
                float pdf = 0.f;
                if (ds > 0.f && dot(wo, wh) > 0.f) {
                    float e = (b.ex * wh.x * wh.x + b.ey * wh.y * wh.y) / ds;
                    pdf = (sqrtf((b.ex + 1.f) * (b.ey + 1.f)) * INV_TWOPI_F * pow01(costhetah, e)) / (4.f * dot(wo, wh));
                }
                *pdfOverride = pdf;
            }
        }
        *wiOut = wi;
        return true;                                                 // pdf == 0 is caught by the caller (one component)
    }
    int matching = bsdf_ncomp(b);
    int which = (int)floorf(uComp * matching);
    if (matching - 1 < which) which = matching - 1;
    bool mf = (b.mtype == SPT_MAT_METAL) || (b.mtype == SPT_MAT_PLASTIC && which == 1);
    v3 wi;
    bool chosenZero;
    if (!mf) {                                                       // BxDF::Sample_f, reflection.cpp:303-310
        wi = cosine_sample_hemisphere(u1, u2);
        if (wo.z < 0.f) wi.z *= -1.f;
        chosenZero = !(same_hemisphere(wo, wi) && abs_cos_theta(wi) * INV_PI_F != 0.f);
    } else {                                                         // Blinn::Sample_f, reflection.cpp:338-354
        float costheta = pow01(u1, 1.f / (b.exponent + 1));
        float sintheta = sqrtf(stdmaxf(0.f, 1.f - costheta * costheta));
        float phi = u2 * 2.f * PI_F;
        float sp, cp;
        sincosf(phi, &sp, &cp);
        v3 wh = V(sintheta * cp, sintheta * sp, costheta);
        if (!same_hemisphere(wo, wh)) wh = vneg(wh);
        float woh = dot(wo, wh);
        wi = vadd(vneg(wo), vmul(wh, 2.f * woh));
        chosenZero = woh <= 0.f || costheta == 0.f;
    }
    *wiOut = wi;
    return !chosenZero;
}
// Second half: value and pdf over all components for the sampled direction. pdf == 0: no sample.
__device__ inline void bsdf_sample(const Bsdf &b, v3 woW, v3 wo, float uComp, float u1, float u2,
                                   v3 *wiW, float *pdf, DirTerms *t) {
    v3 wi;
    *pdf = 0.f;
    t->a0 = t->a1 = t->a2 = t->a3 = 0.f; t->mf = false; t->reflect = false;
    float pdfOverride;
    if (!bsdf_sample_dir<true>(b, wo, uComp, u1, u2, &wi, &pdfOverride)) return;
    *wiW = l2w(b, wi);
    bsdf_terms<true>(b, woW, *wiW, wo, wi, t, pdf);
    if (pdfOverride >= 0.f) *pdf = pdfOverride;
}

// BSDF::Sample_f for an all-specular BSDF (mirror, glass; reflection.cpp:514-572 with SpecularReflection::Sample_f
// :130-136 and SpecularTransmission::Sample_f :139-162). The value is f[c] = Kr[c] * coefR + Kt[c] * coefT with
// coefR = F / |cos theta_i| (F = 1 for the mirror's FresnelNoOp), coefT = (1 - F) / |cos theta_i| (no eta^2 scaling,
// as the reference); pdf = 1 / number of components. Returns false when there is no sample (no component, or total
// internal reflection on the transmission component).
__device__ inline bool specular_sample(const Bsdf &b, v3 wo, float uComp, v3 *wiOut, float *coefR, float *coefT, float *pdf) {
    int matching = specular_ncomp(b);
    *coefR = *coefT = 0.f; *pdf = 0.f;
    if (matching == 0) return false;
    int which = (int)floorf(uComp * matching);
    if (matching - 1 < which) which = matching - 1;
    bool transmit = (b.compMask & 1) ? which == 1 : true;             // component order: reflection, transmission
    float F = b.mtype == SPT_MAT_MIRROR ? 1.f : fresnel_dielectric(wo.z, 1.f, b.ior);
    v3 wi;
    if (!transmit) {
        wi = V(-wo.x, -wo.y, wo.z);
        *coefR = F / abs_cos_theta(wi);
    } else {
        bool entering = wo.z > 0.f;
        float ei = 1.f, et = b.ior;
        if (!entering) { float tmp = ei; ei = et; et = tmp; }
        float sini2 = stdmaxf(0.f, 1.f - wo.z * wo.z);
        float eta = ei / et;
        float sint2 = eta * eta * sini2;
        if (sint2 >= 1.f) return false;
        float cost = sqrtf(stdmaxf(0.f, 1.f - sint2));
        if (entering) cost = -cost;
        wi = V(eta * -wo.x, eta * -wo.y, cost);
        *coefT = (1.f - F) / abs_cos_theta(wi);
    }
    *wiOut = wi;
    *pdf = 1.f;
    if (matching > 1) *pdf /= matching;
    return true;
}

// ---- spectral tables ----------------------------------------------------------------------------
// SampledSpectrum::FromRGB(rgb, SPECTRUM_ILLUMINANT) (src/core/spectrum.cpp:136-176) is linear in
// three basis spectra chosen by the ordering of r,g,b: keep the three coefficients, rebuild per band.
struct IllumCoefs { float k0, k1, k2; int b1, b2; };
__device__ inline IllumCoefs illum_coefs(const float rgb[3]) {
    enum { W = 0, Cy = 1, Mg = 2, Ye = 3, Rd = 4, Gr = 5, Bl = 6 };
    IllumCoefs k;
    if (rgb[0] <= rgb[1] && rgb[0] <= rgb[2]) {
        k.k0 = rgb[0];
        if (rgb[1] <= rgb[2]) { k.k1 = rgb[1] - rgb[0]; k.b1 = Cy; k.k2 = rgb[2] - rgb[1]; k.b2 = Bl; }
        else { k.k1 = rgb[2] - rgb[0]; k.b1 = Cy; k.k2 = rgb[1] - rgb[2]; k.b2 = Gr; }
    } else if (rgb[1] <= rgb[0] && rgb[1] <= rgb[2]) {
        k.k0 = rgb[1];
        if (rgb[0] <= rgb[2]) { k.k1 = rgb[0] - rgb[1]; k.b1 = Mg; k.k2 = rgb[2] - rgb[0]; k.b2 = Bl; }
        else { k.k1 = rgb[2] - rgb[1]; k.b1 = Mg; k.k2 = rgb[0] - rgb[2]; k.b2 = Rd; }
    } else {
        k.k0 = rgb[2];
        if (rgb[0] <= rgb[1]) { k.k1 = rgb[0] - rgb[2]; k.b1 = Ye; k.k2 = rgb[1] - rgb[0]; k.b2 = Gr; }
        else { k.k1 = rgb[1] - rgb[2]; k.b1 = Ye; k.k2 = rgb[0] - rgb[1]; k.b2 = Rd; }
    }
    return k;
}
__device__ __forceinline__ float illum_band(const SptSpectralTables &t, const IllumCoefs &k, int c) {
    float r = 0.f;
    r += t.rgb_illum[0][c] * k.k0;
    r += t.rgb_illum[k.b1][c] * k.k1;
    r += t.rgb_illum[k.b2][c] * k.k2;
    r *= .86445f;
    return clampf(r, 0.f, SPT_INF);
}
__device__ __forceinline__ float spherical_theta(v3 v) { return acosf(clampf(v.z, -1.f, 1.f)); }
__device__ __forceinline__ float spherical_phi(v3 v) { float p = atan2f(v.y, v.x); return (p < 0.f) ? p + 2.f * PI_F : p; }
// IrregIsotropicBRDF::f (reflection.cpp:251-263): radius search around BRDFRemap(wo, wi) (:239-248) in the reference's
// kd-tree (kdtree.h:143-168: children first, then the node - the same order, so the same sums up to the rounding of expf),
// the radius doubling until three samples are found; v = sum of weight * sample spectrum, clamped, over the sum of weights.
__device__ __forceinline__ float refl_band(const SptSpectralTables &t, const IllumCoefs &k, int c);
// RegularHalfangleBRDF::f (reflection.cpp:267-300): the entry of the theta_h / theta_d / phi_d table the half-angle
// parameterisation of (wo, wi) falls in, an RGB triple turned into a spectrum by FromRGB(.., SPECTRUM_REFLECTANCE)
__device__ inline void halfangle_f(const DevScene &sc, const SptBrdfTable &t, v3 wo, v3 wi, float *v) {
    v3 wh = vadd(wo, wi);
    if (wh.z < 0.f) { wo = vneg(wo); wi = vneg(wi); wh = vneg(wh); }
    if (wh.x == 0.f && wh.y == 0.f && wh.z == 0.f) { for (int c = 0; c < NB; ++c) v[c] = 0.f; return; }
    wh = normalize(wh);
    const float whTheta = spherical_theta(wh);
    const float whCosPhi = cos_phi(wh), whSinPhi = sin_phi(wh);
    const float whCosTheta = wh.z, whSinTheta = sin_theta(wh);
    const v3 whx = V(whCosPhi * whCosTheta, whSinPhi * whCosTheta, -whSinTheta);
    const v3 why = V(-whSinPhi, whCosPhi, 0.f);
    const v3 wd = V(dot(wi, whx), dot(wi, why), dot(wi, wh));
    const float wdTheta = spherical_theta(wd);
    float wdPhi = spherical_phi(wd);
    if (wdPhi > PI_F) wdPhi -= PI_F;
    const int nH = (int)t.n_theta_h, nD = (int)t.n_theta_d, nP = (int)t.n_phi_d;
    const int ih = clampi((int)(sqrtf(stdmaxf(0.f, whTheta / (PI_F / 2.f))) / 1.f * nH), 0, nH - 1);
    const int id = clampi((int)(wdTheta / (PI_F / 2.f) * nD), 0, nD - 1);
    const int ip = clampi((int)(wdPhi / PI_F * nP), 0, nP - 1);
    const float *e = sc.merl_rgb + t.rgb_offset + 3 * (size_t)(ip + nP * (id + ih * nD));
    const float rgb[3] = { __ldg(e), __ldg(e + 1), __ldg(e + 2) };
    const IllumCoefs k = illum_coefs(rgb);
    for (int c = 0; c < NB; ++c) v[c] = refl_band(*sc.tables, k, c);
}
__device__ inline void measured_f(const DevScene &sc, const SptBrdfTable &t, v3 wo, v3 wi, float *v) {
    if (t.n_nodes == 0) { halfangle_f(sc, t, wo, wi, v); return; }
    const float cosi = wi.z, coso = wo.z;
    const float sini = sin_theta(wi), sino = sin_theta(wo);
    const float phii = spherical_phi(wi), phio = spherical_phi(wo);
    float dphi = phii - phio;
    if (dphi < 0.f) dphi += 2.f * PI_F;
    if (dphi > 2.f * PI_F) dphi -= 2.f * PI_F;
    if (dphi > PI_F) dphi = 2.f * PI_F - dphi;
    const v3 m = V(sini * sino, dphi / PI_F, cosi * coso);
    const SptKdNode *nodes = sc.brdf_nodes + t.node_first;
    const float *spectra = sc.brdf_spectra + (size_t)t.node_first * NBP;
    const uint32_t NONE = 0xffffffffu;
    // The reference repeats the search with maxDist2 = .001, .002, .004, ... until it finds more than two samples (or
    // maxDist2 exceeds 1.5), i.e. it stops at the first radius^2 of that sequence above the THIRD-SMALLEST squared distance.
    // That distance comes from one 3-nearest-neighbour descent (near child first, far child only if the split plane is
    // closer than the third best so far: a far-side sample's rounded distance^2 is never below the plane's, so nothing
    // that could enter the best three is skipped), and only the final search is run - same radius, same visiting order,
    // same sums as the reference's last round.
    float b0 = SPT_INF, b1 = SPT_INF, b2 = SPT_INF;
    {
        uint32_t stk[32];
        float stkd[32];
        int sp = 0;
        stk[sp] = 0u; stkd[sp++] = 0.f;
        while (sp) {
            --sp;
            const uint32_t n = stk[sp];
            if (!(stkd[sp] < b2)) continue;
            const SptKdNode nd = nodes[n];
            const float dx = nd.p[0] - m.x, dy = nd.p[1] - m.y, dz = nd.p[2] - m.z;
            const float d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < b2) {
                if (d2 < b0) { b2 = b1; b1 = b0; b0 = d2; }
                else if (d2 < b1) { b2 = b1; b1 = d2; }
                else b2 = d2;
            }
            const int axis = (int)(nd.bits & 3u);
            if (axis == 3) continue;
            const bool hasLeft = (nd.bits >> 2) & 1u;
            const uint32_t right = nd.bits >> 3;
            const float pa = vcomp(m, axis);
            const float pd = (pa - nd.split_pos) * (pa - nd.split_pos);
            const bool leftFirst = pa <= nd.split_pos;
            const uint32_t L = hasLeft ? n + 1 : NONE, R = right < t.n_nodes ? right : NONE;
            const uint32_t nearC = leftFirst ? L : R, farC = leftFirst ? R : L;
            if (farC != NONE && sp < 31) { stk[sp] = farC; stkd[sp++] = pd; }
            if (nearC != NONE && sp < 31) { stk[sp] = nearC; stkd[sp++] = 0.f; }
        }
    }
    float maxD2 = .001f;
    while (!(b2 < maxD2) && !(maxD2 > 1.5f)) maxD2 *= 2.f;
    {
        for (int c = 0; c < NB; ++c) v[c] = 0.f;
        float sumW = 0.f;
        uint32_t stack[32];                                          // node << 2 | stage: 0 enter, 1 first child done, 2 both done
        int sp = 0;
        stack[sp++] = 0u;
        while (sp) {
            const uint32_t e = stack[--sp], n = e >> 2, stage = e & 3u;
            const SptKdNode nd = nodes[n];
            const int axis = (int)(nd.bits & 3u);
            if (axis != 3 && stage < 2u) {
                const bool hasLeft = (nd.bits >> 2) & 1u;
                const uint32_t right = nd.bits >> 3;
                const float pa = vcomp(m, axis);
                const float dist2 = (pa - nd.split_pos) * (pa - nd.split_pos);
                const bool leftFirst = pa <= nd.split_pos;
                const uint32_t L = hasLeft ? n + 1 : NONE, R = right < t.n_nodes ? right : NONE;
                if (stage == 0u) {
                    stack[sp++] = n << 2 | 1u;
                    const uint32_t first = leftFirst ? L : R;
                    if (first != NONE) stack[sp++] = first << 2;
                } else {
                    stack[sp++] = n << 2 | 2u;
                    const uint32_t second = leftFirst ? R : L;
                    if (dist2 < maxD2 && second != NONE) stack[sp++] = second << 2;
                }
                continue;
            }
            const float dx = nd.p[0] - m.x, dy = nd.p[1] - m.y, dz = nd.p[2] - m.z;
            const float d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < maxD2) {
                const float weight = expf(-100.f * d2);
                const float *sv = spectra + (size_t)n * NBP;
                for (int c = 0; c < NB; ++c) v[c] += __ldg(sv + c) * weight;
                sumW += weight;
            }
        }
        for (int c = 0; c < NB; ++c) v[c] = clampf(v[c], 0.f, SPT_INF) / sumW;
    }
}

// FromRGB(rgb, SPECTRUM_REFLECTANCE) (spectrum.cpp:92-133,175): same basis choice as the illuminant form, the
// rgbRefl2Spect* tables, scale .94
__device__ __forceinline__ float refl_band(const SptSpectralTables &t, const IllumCoefs &k, int c) {
    float r = 0.f;
    r += t.rgb_refl[0][c] * k.k0;
    r += t.rgb_refl[k.b1][c] * k.k1;
    r += t.rgb_refl[k.b2][c] * k.k2;
    r *= .94f;
    return clampf(r, 0.f, SPT_INF);
}
// MIPMap::Lookup(s,t) -> triangle(0,s,t), repeat wrap (src/core/mipmap.h:198-221,233-274)
__device__ inline void env_lookup(const DevScene &sc, float s, float t, float rgb[3]) {
    int w = sc.env_w, h = sc.env_h;
    s = s * w - 0.5f;
    t = t * h - 0.5f;
    int s0 = (int)floorf(s), t0 = (int)floorf(t);
    float ds = s - s0, dt = t - t0;
    // power-of-two map (spt_scene_create checks): Mod is a mask
    const int sa = s0 & (w - 1), sb = (s0 + 1) & (w - 1), ta = t0 & (h - 1), tb = (t0 + 1) & (h - 1);
    const float *a = sc.env_rgb + 3 * ((size_t)ta * w + sa);
    const float *b = sc.env_rgb + 3 * ((size_t)tb * w + sa);
    const float *c = sc.env_rgb + 3 * ((size_t)ta * w + sb);
    const float *d = sc.env_rgb + 3 * ((size_t)tb * w + sb);
#pragma unroll
    for (int k = 0; k < 3; ++k)
        rgb[k] = a[k] * ((1.f - ds) * (1.f - dt)) + b[k] * ((1.f - ds) * dt) + c[k] * (ds * (1.f - dt)) + d[k] * (ds * dt);
}
// InfiniteAreaLight::Le (src/lights/infinite.cpp:109-114) as RGB; spectrum via illum_coefs/illum_band
__device__ inline void infinite_le_rgb(const DevScene &sc, const SptLight &l, v3 d, float rgb[3]) {
    const SptXform &xf = sc.xforms[l.xform];
    v3 wh = normalize(xf_vector(xf.minv, d));
    float s = spherical_phi(wh) * INV_TWOPI_F;
    float t = spherical_theta(wh) * INV_PI_F;
    env_lookup(sc, s, t, rgb);
}
// std::upper_bound(cdf, cdf+count+1, u) - cdf - 1, clamped at 0 (montecarlo.h:70-99)
__device__ inline int cdf_find(const float *cdf, int count, float u) {
    int lo = 0, hi = count + 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    int offset = lo - 1;
    return offset < 0 ? 0 : offset;
}
__device__ inline float dist1d_sample_continuous(const float *func, const float *cdf, float funcInt, int count,
                                                 float u, float *pdf, int *off) {
    int offset = cdf_find(cdf, count, u);
    if (off) *off = offset;
    float du = (u - cdf[offset]) / (cdf[offset + 1] - cdf[offset]);
    *pdf = func[offset] / funcInt;
    return (offset + du) / count;
}
__device__ inline float env_pdf_uv(const DevScene &sc, float u, float v) {                     // montecarlo.h:146-154
    int nu = sc.env_w, nv = sc.env_h;
    int iu = clampi((int)(u * nu), 0, nu - 1);
    int iv = clampi((int)(v * nv), 0, nv - 1);
    if (sc.env_func_int[iv] * sc.env_marg_int == 0.f) return 0.f;
    return (sc.env_func[(size_t)iv * nu + iu] * sc.env_marg_func[iv]) / (sc.env_func_int[iv] * sc.env_marg_int);
}

// ---- area-light shapes (ShapeSet, src/core/light.cpp:106-172) ----------------------------------
__device__ inline v3 tri_vertex(const DevScene &sc, int idx) { const float *p = sc.P + 3 * (size_t)idx; return V(p[0], p[1], p[2]); }
// Shape::Sample(u1,u2,Ns): trianglemesh.cpp:436-448, disk.cpp:140-149, sphere.cpp:220-225
__device__ inline v3 shape_sample_area(const DevScene &sc, const SptLightShape &s, float u1, float u2, v3 *ns) {
    if (s.kind == SPT_PRIM_TRIANGLE) {
        float su1 = sqrtf(u1);                                        // UniformSampleTriangle, montecarlo.cpp:342-347
        float b1 = 1.f - su1, b2 = u2 * su1;
        const int32_t *vi = sc.tri_vidx + 3 * (size_t)s.data;
        v3 p1 = tri_vertex(sc, vi[0]), p2 = tri_vertex(sc, vi[1]), p3 = tri_vertex(sc, vi[2]);
        v3 p = vadd(vadd(vmul(p1, b1), vmul(p2, b2)), vmul(p3, (1.f - b1 - b2)));
        *ns = normalize(cross(vsub(p2, p1), vsub(p3, p1)));
        if (s.flags & SPT_PF_REVERSE) *ns = vmul(*ns, -1.f);
        return p;
    }
    const SptQuadric &q = sc.quadrics[s.data];
    const SptXform &xf = sc.xforms[q.xform];
    if (s.kind == SPT_PRIM_DISK) {
        v3 p;
        concentric_sample_disk(u1, u2, &p.x, &p.y);
        p.x *= q.radius; p.y *= q.radius; p.z = q.zmin;
        *ns = normalize(xf_normal(xf.minv, V(0, 0, 1)));
        if (s.flags & SPT_PF_REVERSE) *ns = vmul(*ns, -1.f);
        return xf_point(xf.m, p);
    }
    v3 p = vadd(V(0, 0, 0), vmul(uniform_sample_sphere(u1, u2), q.radius));
    *ns = normalize(xf_normal(xf.minv, p));
    if (s.flags & SPT_PF_REVERSE) *ns = vmul(*ns, -1.f);
    return xf_point(xf.m, p);
}
// Shape::Sample(p,u1,u2,Ns): Sphere sphere.cpp:228-252, others = area sampling
__device__ inline v3 shape_sample_from(const DevScene &sc, const SptLightShape &s, v3 p, float u1, float u2, v3 *ns) {
    if (s.kind != SPT_PRIM_SPHERE) return shape_sample_area(sc, s, u1, u2, ns);
    const SptQuadric &q = sc.quadrics[s.data];
    const SptXform &xf = sc.xforms[q.xform];
    v3 Pcenter = xf_point(xf.m, V(0, 0, 0));
    v3 wc = normalize(vsub(Pcenter, p));
    v3 wcX, wcY;
    coordinate_system(wc, &wcX, &wcY);
    if (len2(vsub(p, Pcenter)) - q.radius * q.radius < 1e-4f) return shape_sample_area(sc, s, u1, u2, ns);
    float sinThetaMax2 = q.radius * q.radius / len2(vsub(p, Pcenter));
    float cosThetaMax = sqrtf(stdmaxf(0.f, 1.f - sinThetaMax2));
    Ray r; r.o = p; r.d = uniform_sample_cone(u1, u2, cosThetaMax, wcX, wcY, wc); r.mint = 1e-3f; r.maxt = SPT_INF;
    float thit;
    if (!sphere_intersect(sc, q, s.flags, r, &thit, nullptr)) thit = dot(vsub(Pcenter, p), normalize(r.d));
    v3 ps = ray_at(r, thit);
    *ns = normalize(vsub(ps, Pcenter));
    if (s.flags & SPT_PF_REVERSE) *ns = vmul(*ns, -1.f);
    return ps;
}
// Shape::Pdf(p,wi) shape.cpp:78-91; Sphere::Pdf sphere.cpp:255-266
__device__ inline float shape_pdf(const DevScene &sc, const SptLightShape &s, v3 p, v3 wi) {
    if (s.kind == SPT_PRIM_SPHERE) {
        const SptQuadric &q = sc.quadrics[s.data];
        const SptXform &xf = sc.xforms[q.xform];
        v3 Pcenter = xf_point(xf.m, V(0, 0, 0));
        if (!(len2(vsub(p, Pcenter)) - q.radius * q.radius < 1e-4f)) {
            float sinThetaMax2 = q.radius * q.radius / len2(vsub(p, Pcenter));
            float cosThetaMax = sqrtf(stdmaxf(0.f, 1.f - sinThetaMax2));
            return uniform_cone_pdf(cosThetaMax);
        }
    }
    Ray ray; ray.o = p; ray.d = wi; ray.mint = 1e-3f; ray.maxt = SPT_INF;
    Hit h;
    if (!shape_intersect(sc, s.kind, s.flags, (uint32_t)s.data, ray, &h)) return 0.f;
    float pdf = len2(vsub(p, ray_at(ray, h.t))) / (absdot(h.nn, vneg(wi)) * s.area);
    if (isinf(pdf)) pdf = 0.f;
    return pdf;
}
__device__ inline float shapeset_pdf(const DevScene &sc, const SptLight &l, v3 p, v3 wi) {    // light.cpp:156-161
    float pdf = 0.f;
    for (int i = 0; i < l.shape_count; ++i) {
        const SptLightShape &s = sc.light_shapes[l.shape_first + i];
        pdf += s.area * shape_pdf(sc, s, p, wi);
    }
    return pdf / l.sum_area;
}

// Can the BSDF-sampled ray of the MIS estimate (integrator.cpp:139-163) reach this area light at all?
// EstimateDirect traces it through the whole scene and keeps Le only when the CLOSEST primitive
// belongs to the light (:153-156); a ray that misses every shape of the light therefore contributes
// exactly nothing and need not be traced. Only spheres need the test: every other shape's
// Shape::Pdf already intersects it (pdf 0 on a miss, shape.cpp:78-91), while Sphere::Pdf returns
// the cone pdf for ANY direction (sphere.cpp:255-266). Same sphere_intersect, same ray as the BVH
// leaf would see, so "miss" here is the traversal's "miss". Conservative (true) for mixed sets.
__device__ inline bool light_ray_may_hit(const DevScene &sc, const SptLight &l, const Ray &ray) {
    for (int i = 0; i < l.shape_count; ++i) {
        const SptLightShape &s = sc.light_shapes[l.shape_first + i];
        if (s.kind != SPT_PRIM_SPHERE) return true;
        float t;
        if (sphere_intersect(sc, sc.quadrics[s.data], 0, ray, &t, nullptr)) return true;
    }
    return false;
}

// Result of Light::Sample_L (diffuse.cpp:61-73 + light.cpp:137-149, point.cpp:42-49,
// infinite.cpp:187-213) with the VisibilityTester segment (light.h:79-88). The radiance is kept as
// {kind, aux}: AREA on/off (x Lemit), POINT 1/d2 divisor, INFINITE rgb.
struct LightSampleResult { bool delta; bool black; bool pdfPending; v3 wi; float pdf; v3 shadow_d; float shadow_maxt; float aux[3]; };

// deferPdf: leave an area light's pdf (ShapeSet::Pdf of the sampled direction) to the caller, which
// evaluates Light::Pdf for this and the BSDF-sampled direction from one call site (pdfPending is set).
__device__ inline void light_sample(const DevScene &sc, int lightIdx, v3 p, float u0, float u1, float uComp,
                                    LightSampleResult *out, bool deferPdf = false) {
    const SptLight &l = sc.lights[lightIdx];
    out->delta = false; out->black = true; out->pdf = 0.f; out->pdfPending = false;
    out->aux[0] = out->aux[1] = out->aux[2] = 0.f;
    out->wi = V(0, 0, 1); out->shadow_d = V(0, 0, 1); out->shadow_maxt = 0.f;
    if (l.type == SPT_LIGHT_POINT) {
        v3 lp = V(l.pos[0], l.pos[1], l.pos[2]);
        out->delta = true;
        out->wi = normalize(vsub(lp, p));
        out->pdf = 1.f;
        float dist = sqrtf(len2(vsub(p, lp)));
        out->shadow_d = vdiv(vsub(lp, p), dist);
        out->shadow_maxt = dist * (1.f - 0.f);
        out->aux[0] = len2(vsub(lp, p));
        out->black = false;
        return;
    }
    if (l.type == SPT_LIGHT_AREA) {
        const float *cdf = sc.light_cdf + l.shape_first + lightIdx;
        int sn = cdf_find(cdf, l.shape_count, uComp);
        v3 ns;
        const SptLightShape &chosen = sc.light_shapes[l.shape_first + sn];
        v3 pt = shape_sample_from(sc, chosen, p, u0, u1, &ns);
        v3 ps = pt;
        // ShapeSet::Sample re-intersects every shape of the set with the unclipped ray p -> pt and keeps
        // the LAST hit (light.cpp:141-149) - also for a set of one sphere: at the rim of the cone Sphere::Sample
        // draws from, its float quadratic misses the sphere, Sample returns the point of closest approach, and
        // this second intersection (another direction length, other rounding) decides the point and its normal.
        Ray r; r.o = p; r.d = vsub(pt, p); r.mint = 1e-3f; r.maxt = SPT_INF;
        float thit = 1.f;
        int hitSphere = -1;                   // the last hit is on this sphere: its normal is worked out below
        v3 phitObj = V(0, 0, 0);
        {
            v3 nnHit = ns;
            for (int i = 0; i < l.shape_count; ++i) {
                const SptLightShape &s = sc.light_shapes[l.shape_first + i];
                if (s.kind == SPT_PRIM_SPHERE) {
                    float t;
                    if (sphere_intersect(sc, sc.quadrics[s.data], s.flags, r, &t, nullptr, &phitObj)) { thit = t; hitSphere = i; }
                } else {
                    Hit hh;
                    if (shape_intersect(sc, s.kind, s.flags, (uint32_t)s.data, r, &hh)) { nnHit = hh.nn; thit = hh.t; hitSphere = -1; }
                }
            }
            ns = nnHit;
            ps = ray_at(r, thit);
        }
        out->wi = normalize(vsub(ps, p));
        if (deferPdf) out->pdfPending = true;
        else out->pdf = shapeset_pdf(sc, l, p, out->wi);
        float dist = sqrtf(len2(vsub(p, ps)));
        out->shadow_d = vdiv(vsub(ps, p), dist);
        out->shadow_maxt = dist * (1.f - 1e-3f);
        if (hitSphere >= 0) {
            // Only the SIGN of Dot(dgLight.nn, -wi) is used (DiffuseAreaLight::L, diffuse.h:47-49). dgLight.nn =
            // Normalize(Cross(dpdu, dpdv)) of the transformed tangents, flipped by ReverseOrientation ^ SwapsHandedness
            // (sphere.cpp:105-149, diffgeom.cpp:44-46), points along ObjectToWorld(Normal(phit)), reversed iff ReverseOrientation:
            // Cross(M a, M b) = det(M) M^-T (a x b), a x b is a positive multiple of phit, and the two handedness signs cancel.
            // That normal costs no atan2f / acosf / sinf; where its cosine is within 1e-3 of zero (rounding differences
            // between the two are ~1e-6) the DifferentialGeometry is built as the reference builds it.
            const SptLightShape &s = sc.light_shapes[l.shape_first + hitSphere];
            const SptQuadric &q = sc.quadrics[s.data];
            const SptXform &xf = sc.xforms[q.xform];
            // (the sign of the un-normalised cosine first; its size against |n| only near zero)
            v3 nq = xf_normal(xf.minv, phitObj);
            float c = dot(nq, vneg(out->wi));
            if (s.flags & SPT_PF_REVERSE) c = -c;
            if (fabsf(c) < 1e-3f * sqrtf(len2(nq))) {
                Hit hh;
                sphere_record(sc, q, s.flags, r, thit, &hh);
                c = dot(hh.nn, vneg(out->wi));
            }
            out->black = !(c > 0.f);
            return;
        }
        out->black = !(dot(ns, vneg(out->wi)) > 0.f);
        return;
    }
    float uv[2], pdfs[2];
    int v;
    uv[1] = dist1d_sample_continuous(sc.env_marg_func, sc.env_marg_cdf, sc.env_marg_int, sc.env_h, u1, &pdfs[1], &v);
    uv[0] = dist1d_sample_continuous(sc.env_func + (size_t)v * sc.env_w, sc.env_cdf + (size_t)v * (sc.env_w + 1),
                                     sc.env_func_int[v], sc.env_w, u0, &pdfs[0], nullptr);
    float mapPdf = pdfs[0] * pdfs[1];
    if (mapPdf == 0.f) return;
    float theta = uv[1] * PI_F, phi = uv[0] * 2.f * PI_F;
    float costheta, sintheta, sinphi, cosphi;
    sin_cos(theta, &sintheta, &costheta);
    sin_cos(phi, &sinphi, &cosphi);
    const SptXform &xf = sc.xforms[l.xform];
    out->wi = xf_vector(xf.m, V(sintheta * cosphi, sintheta * sinphi, costheta));
    out->pdf = mapPdf / (2.f * PI_F * PI_F * sintheta);
    if (sintheta == 0.f) out->pdf = 0.f;
    out->shadow_d = out->wi;
    out->shadow_maxt = SPT_INF;
    env_lookup(sc, uv[0], uv[1], out->aux);
    // black iff every band of FromRGB(...) clamps to zero
    IllumCoefs k = illum_coefs(out->aux);
    bool black = true;
    for (int c = 0; c < NB; ++c) if (illum_band(*sc.tables, k, c) != 0.f) { black = false; break; }
    out->black = black;
}
// Light::Pdf(p,wi): diffuse.cpp:76-78, infinite.cpp:216-226
__device__ inline float light_pdf(const DevScene &sc, int lightIdx, v3 p, v3 w) {
    const SptLight &l = sc.lights[lightIdx];
    if (l.type == SPT_LIGHT_AREA) return shapeset_pdf(sc, l, p, w);
    if (l.type == SPT_LIGHT_POINT) return 0.f;
    const SptXform &xf = sc.xforms[l.xform];
    v3 wi = xf_vector(xf.minv, w);
    float theta = spherical_theta(wi), phi = spherical_phi(wi);
    float sintheta = sinf(theta);
    if (sintheta == 0.f) return 0.f;
    return env_pdf_uv(sc, phi * INV_TWOPI_F, theta * INV_PI_F) / (2.f * PI_F * PI_F * sintheta);
}

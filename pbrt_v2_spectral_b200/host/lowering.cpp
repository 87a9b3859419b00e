// lowering.cpp — Scene -> flat SoA buffers (include/spt.h). Host side of the drop-in boundary.
//
// The reference keeps everything the GPU needs in private members (BVHAccel::nodes/primitives
// src/accelerators/bvh.h:58-63, GeometricPrimitive::shape/material/areaLight
// src/core/primitive.h:80-83, Triangle::mesh/v src/shapes/trianglemesh.h:99-101, ...). A
// maintainer would add `friend struct GpuSceneLowering;` lines; to stay zero-patch this TU is
// instead compiled with private/protected opened up AFTER all standard headers are included
// (option ii).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <typeinfo>
#include <vector>
#include <list>
#include <pthread.h>
#include <stdint.h>

#define private public
#define protected public
#include "pbrt.h"
#include "scene.h"
#include "primitive.h"
#include "shape.h"
#include "light.h"
#include "camera.h"
#include "film.h"
#include "filter.h"
#include "sampler.h"
#include "integrator.h"
#include "texture.h"
#include "mipmap.h"
#include "montecarlo.h"
#include "spectrum.h"
#include "accelerators/bvh.h"
#include "shapes/trianglemesh.h"
#include "shapes/sphere.h"
#include "shapes/disk.h"
#include "materials/matte.h"
#include "materials/plastic.h"
#include "materials/metal.h"
#include "materials/mirror.h"
#include "materials/glass.h"
#include "materials/subsurface.h"
#include "materials/substrate.h"
#include "materials/measured.h"
#include "kdtree.h"
#include "reflection.h"
#include "textures/imagemap.h"
#include "textures/scale.h"
#include "lights/diffuse.h"
#include "lights/point.h"
#include "lights/infinite.h"
#include "cameras/perspective.h"
#include "film/spectralImage.h"
#include "samplers/lowdiscrepancy.h"
#include "integrators/path.h"
#include "integrators/directlighting.h"
#include "textures/constant.h"
#undef private
#undef protected

#include "lowering.h"

// Same bytes as the reference's file-local LinearBVHNode (src/accelerators/bvh.cpp:105-115).
struct LoweringBVHNode {
    float bounds[6];
    uint32_t offset;        // primitivesOffset (leaf) | secondChildOffset (interior)
    uint8_t nPrimitives;
    uint8_t axis;
    uint8_t pad[2];
};

LoweredScene::LoweredScene() : env_w(0), env_h(0), env_marg_int(0.f) {
    memset(&tables, 0, sizeof(tables));
    memset(&camera, 0, sizeof(camera));
    memset(&film, 0, sizeof(film));
    memset(&params, 0, sizeof(params));
}

template <typename T> static const T *ptr_or_null(const std::vector<T> &v) {
    return v.empty() ? NULL : &v[0];
}

SptSceneDesc LoweredScene::Desc() const {
    SptSceneDesc d;
    memset(&d, 0, sizeof(d));
    d.nbands = SPT_NBANDS;
    d.n_nodes = (uint32_t)(bvh_nodes.size() / 32);
    d.bvh_nodes = ptr_or_null(bvh_nodes);
    d.n_prims = (uint32_t)prim_kind.size();
    d.prim_kind = ptr_or_null(prim_kind);
    d.prim_flags = ptr_or_null(prim_flags);
    d.prim_id = ptr_or_null(prim_id);
    d.prim_data = ptr_or_null(prim_data);
    d.prim_material = ptr_or_null(prim_material);
    d.prim_light = ptr_or_null(prim_light);
    d.prim_xform = ptr_or_null(prim_xform);
    d.n_tris = (uint32_t)(tri_vidx.size() / 3);
    d.tri_vidx = ptr_or_null(tri_vidx);
    d.n_verts = (uint32_t)(P.size() / 3);
    d.P = ptr_or_null(P);
    d.N = ptr_or_null(N);
    d.UV = ptr_or_null(UV);
    d.n_quadrics = (uint32_t)quadrics.size();
    d.quadrics = ptr_or_null(quadrics);
    d.n_xforms = (uint32_t)xforms.size();
    d.xforms = ptr_or_null(xforms);
    d.n_materials = (uint32_t)materials.size();
    d.materials = ptr_or_null(materials);
    d.n_lights = (uint32_t)lights.size();
    d.lights = ptr_or_null(lights);
    d.n_light_shapes = (uint32_t)light_shapes.size();
    d.light_shapes = ptr_or_null(light_shapes);
    d.tables = tables;
    d.env_w = env_w;
    d.env_h = env_h;
    d.env_rgb = ptr_or_null(env_rgb);
    d.env_func = ptr_or_null(env_func);
    d.env_cdf = ptr_or_null(env_cdf);
    d.env_func_int = ptr_or_null(env_func_int);
    d.env_marg_func = ptr_or_null(env_marg_func);
    d.env_marg_cdf = ptr_or_null(env_marg_cdf);
    d.env_marg_int = env_marg_int;
    d.n_textures = (uint32_t)textures.size();
    d.textures = ptr_or_null(textures);
    d.n_texels = tex_texels.size();
    d.tex_texels = ptr_or_null(tex_texels);
    d.ewa_weight_lut = ptr_or_null(ewa_weight_lut);
    d.n_brdfs = (uint32_t)brdfs.size();
    d.brdfs = ptr_or_null(brdfs);
    d.n_brdf_nodes = (uint32_t)brdf_nodes.size();
    d.brdf_nodes = ptr_or_null(brdf_nodes);
    d.brdf_spectra = ptr_or_null(brdf_spectra);
    d.n_merl_floats = merl_rgb.size(); d.merl_rgb = ptr_or_null(merl_rgb);
    return d;
}

bool LoweredScene::Save(const std::string &path, std::string *err) const {
    SptContainerWriter w;
    if (!w.begin(path)) { if (err) *err = "cannot open " + path; return false; }
    int32_t nb = SPT_NBANDS;
    w.put("nbands", 1, &nb, 4, 1);
    w.put("bvh_nodes", 0, ptr_or_null(bvh_nodes), bvh_nodes.size(), bvh_nodes.size() / 32, 32);
    w.vec("prim_kind", 0, prim_kind);
    w.vec("prim_flags", 0, prim_flags);
    w.vec("prim_id", 2, prim_id);
    w.vec("prim_data", 2, prim_data);
    w.vec("prim_material", 1, prim_material);
    w.vec("prim_light", 1, prim_light);
    w.vec("prim_xform", 1, prim_xform);
    w.vec("tri_vidx", 1, tri_vidx, 3);
    w.vec("P", 3, P, 3);
    w.vec("N", 3, N, 3);
    w.vec("UV", 3, UV, 2);
    w.pod("quadrics", quadrics);
    w.pod("xforms", xforms);
    w.pod("materials", materials);
    w.pod("lights", lights);
    w.pod("light_shapes", light_shapes);
    w.put("tables", 0, &tables, sizeof(tables), 1, sizeof(tables));
    int32_t envdims[2] = { env_w, env_h };
    w.put("env_dims", 1, envdims, 8, 2);
    w.vec("env_rgb", 3, env_rgb);
    w.vec("env_func", 3, env_func);
    w.vec("env_cdf", 3, env_cdf);
    w.vec("env_func_int", 3, env_func_int);
    w.vec("env_marg_func", 3, env_marg_func);
    w.vec("env_marg_cdf", 3, env_marg_cdf);
    w.put("env_marg_int", 3, &env_marg_int, 4, 1);
    w.pod("textures", textures);
    w.vec("tex_texels", 3, tex_texels);
    w.vec("ewa_weight_lut", 3, ewa_weight_lut);
    w.pod("brdfs", brdfs);
    w.pod("brdf_nodes", brdf_nodes);
    w.vec("brdf_spectra", 3, brdf_spectra);
    w.vec("merl_rgb", 3, merl_rgb);
    w.put("camera", 0, &camera, sizeof(camera), 1, sizeof(camera));
    w.put("film", 0, &film, sizeof(film), 1, sizeof(film));
    w.put("params", 0, &params, sizeof(params), 1, sizeof(params));
    w.put("film_filename", 0, film_filename.data(), film_filename.size(), film_filename.size());
    w.end();
    return true;
}

// ---------------------------------------------------------------------------------------------
namespace {

// texel type of an image map -> its Treturn, channel count and how a texel is flattened
template <typename T> struct ImageRet;
template <> struct ImageRet<RGBSpectrum> {
    typedef Spectrum type;
    enum { channels = 3 };
    static void push(const RGBSpectrum &v, std::vector<float> *dst) { float rgb[3]; v.ToRGB(rgb); dst->insert(dst->end(), rgb, rgb + 3); }
};
template <> struct ImageRet<float> {
    typedef float type;
    enum { channels = 1 };
    static void push(float v, std::vector<float> *dst) { dst->push_back(v); }
};

struct Lowerer {
    LoweredScene *out;
    std::string err;
    std::map<const Transform *, int> xformIdx;
    std::map<const TriangleMesh *, std::pair<uint32_t, uint32_t> > meshBase;  // vertex base, tri base
    std::map<const Shape *, int> quadricIdx;
    std::map<const Light *, int> lightIdx;
    std::map<const void *, int> brdfIdx;                                       // KdTree -> row of brdfs[]
    std::map<const void *, uint64_t> texelBase;                                // MIPMap -> first float in tex_texels

    bool fail(const std::string &why) { err = why; return false; }

    int AddXform(const Transform *o2w) {
        std::map<const Transform *, int>::iterator it = xformIdx.find(o2w);
        if (it != xformIdx.end()) return it->second;
        SptXform x;
        memcpy(x.m, o2w->m.m, 64);
        memcpy(x.minv, o2w->mInv.m, 64);
        out->xforms.push_back(x);
        int idx = (int)out->xforms.size() - 1;
        xformIdx[o2w] = idx;
        return idx;
    }

    std::pair<uint32_t, uint32_t> AddMesh(const TriangleMesh *mesh) {
        std::map<const TriangleMesh *, std::pair<uint32_t, uint32_t> >::iterator it = meshBase.find(mesh);
        if (it != meshBase.end()) return it->second;
        uint32_t vbase = (uint32_t)(out->P.size() / 3);
        uint32_t tbase = (uint32_t)(out->tri_vidx.size() / 3);
        for (int i = 0; i < mesh->nverts; ++i) {
            out->P.push_back(mesh->p[i].x); out->P.push_back(mesh->p[i].y); out->P.push_back(mesh->p[i].z);
            if (mesh->n) {
                out->N.push_back(mesh->n[i].x); out->N.push_back(mesh->n[i].y); out->N.push_back(mesh->n[i].z);
            } else { out->N.push_back(0.f); out->N.push_back(0.f); out->N.push_back(0.f); }
            if (mesh->uvs) { out->UV.push_back(mesh->uvs[2*i]); out->UV.push_back(mesh->uvs[2*i+1]); }
            else { out->UV.push_back(0.f); out->UV.push_back(0.f); }
        }
        for (int i = 0; i < 3 * mesh->ntris; ++i)
            out->tri_vidx.push_back((int32_t)(mesh->vertexIndex[i] + vbase));
        std::pair<uint32_t, uint32_t> r(vbase, tbase);
        meshBase[mesh] = r;
        return r;
    }

    static uint8_t ShapeFlags(const Shape *s) {
        uint8_t f = 0;
        if (s->ReverseOrientation ^ s->TransformSwapsHandedness) f |= SPT_PF_FLIP_NORMAL;
        if (s->ReverseOrientation) f |= SPT_PF_REVERSE;
        return f;
    }

    // kind/flags/data of one intersectable shape; registers meshes, quadrics, transforms
    bool AddShape(const Shape *shape, uint8_t *kind, uint8_t *flags, uint32_t *data, int32_t *xform) {
        *flags = ShapeFlags(shape);
        *xform = AddXform(shape->ObjectToWorld);
        if (const Triangle *tri = dynamic_cast<const Triangle *>(shape)) {
            const TriangleMesh *mesh = tri->mesh.GetPtr();
            if (mesh->alphaTexture.GetPtr())
                return fail("triangle mesh with an alpha texture is not supported");
            if (mesh->s) return fail("triangle mesh with explicit tangents S is not supported");
            std::pair<uint32_t, uint32_t> base = AddMesh(mesh);
            if (mesh->n) *flags |= SPT_PF_HAS_N;
            if (mesh->uvs) *flags |= SPT_PF_HAS_UV;
            *kind = SPT_PRIM_TRIANGLE;
            *data = base.second + (uint32_t)((tri->v - mesh->vertexIndex) / 3);
            return true;
        }
        std::map<const Shape *, int>::iterator it = quadricIdx.find(shape);
        if (const Sphere *sp = dynamic_cast<const Sphere *>(shape)) {
            *kind = SPT_PRIM_SPHERE;
            if (it == quadricIdx.end()) {
                SptQuadric q;
                memset(&q, 0, sizeof(q));
                q.kind = SPT_PRIM_SPHERE; q.xform = *xform; q.radius = sp->radius;
                q.zmin = sp->zmin; q.zmax = sp->zmax; q.thetaMin = sp->thetaMin;
                q.thetaMax = sp->thetaMax; q.phiMax = sp->phiMax;
                out->quadrics.push_back(q);
                quadricIdx[shape] = (int)out->quadrics.size() - 1;
            }
            *data = (uint32_t)quadricIdx[shape];
            return true;
        }
        if (const Disk *dk = dynamic_cast<const Disk *>(shape)) {
            *kind = SPT_PRIM_DISK;
            if (it == quadricIdx.end()) {
                SptQuadric q;
                memset(&q, 0, sizeof(q));
                q.kind = SPT_PRIM_DISK; q.xform = *xform; q.radius = dk->radius;
                q.zmin = dk->height; q.zmax = dk->innerRadius; q.phiMax = dk->phiMax;
                out->quadrics.push_back(q);
                quadricIdx[shape] = (int)out->quadrics.size() - 1;
            }
            *data = (uint32_t)quadricIdx[shape];
            return true;
        }
        return fail(std::string("unsupported shape type ") + typeid(*shape).name());
    }

    template <typename T> bool ConstTex(const Reference<Texture<T> > &tex, T *value, const char *what) {
        const ConstantTexture<T> *c = dynamic_cast<const ConstantTexture<T> *>(tex.GetPtr());
        if (!c) return fail(std::string("non-constant texture for ") + what);
        *value = c->value;
        return true;
    }

    // One image map + its UV mapping -> textures[] (de-duplicated by value) and the texel pool.
    template <typename T> bool AddImageTexture(const ImageTexture<T, typename ImageRet<T>::type> *img, float scale,
                                               int32_t *idx, const char *what) {
        const UVMapping2D *uv = dynamic_cast<const UVMapping2D *>(img->mapping);
        if (!uv) return fail(std::string("texture mapping other than uv for ") + what);
        const MIPMap<T> *mm = img->mipmap;
        SptTexture t;
        memset(&t, 0, sizeof(t));
        t.channels = ImageRet<T>::channels;
        t.width = (int32_t)mm->width; t.height = (int32_t)mm->height; t.n_levels = (int32_t)mm->nLevels;
        t.wrap = mm->wrapMode == TEXTURE_REPEAT ? SPT_WRAP_REPEAT : (mm->wrapMode == TEXTURE_BLACK ? SPT_WRAP_BLACK : SPT_WRAP_CLAMP);
        t.trilinear = mm->doTrilinear ? 1 : 0;
        t.no_filter = mm->noFiltering ? 1 : 0;
        t.max_aniso = mm->maxAnisotropy;
        t.su = uv->su; t.sv = uv->sv; t.du = uv->du; t.dv = uv->dv;
        t.scale = scale;
        std::map<const void *, uint64_t>::iterator it = texelBase.find((const void *)mm);
        if (it == texelBase.end()) {
            uint64_t base = out->tex_texels.size();
            for (uint32_t l = 0; l < mm->nLevels; ++l) {
                const BlockedArray<T> &lvl = *mm->pyramid[l];
                for (uint32_t v = 0; v < lvl.vSize(); ++v)
                    for (uint32_t u = 0; u < lvl.uSize(); ++u) ImageRet<T>::push(lvl(u, v), &out->tex_texels);
            }
            texelBase[(const void *)mm] = base;
            it = texelBase.find((const void *)mm);
            if (out->ewa_weight_lut.empty() && MIPMap<T>::weightLut)
                out->ewa_weight_lut.assign(MIPMap<T>::weightLut, MIPMap<T>::weightLut + WEIGHT_LUT_SIZE);
        }
        t.texel_offset = it->second;
        for (size_t i = 0; i < out->textures.size(); ++i)
            if (!memcmp(&out->textures[i], &t, sizeof(t))) { *idx = (int32_t)i; return true; }
        out->textures.push_back(t);
        *idx = (int32_t)out->textures.size() - 1;
        return true;
    }

    // Kd: a constant (-> spec0) or an RGB image map (-> tex index)
    bool SpectrumParam(const Reference<Texture<Spectrum> > &tex, float *spec, int32_t *texIdx, const char *what) {
        *texIdx = -1;
        if (const ConstantTexture<Spectrum> *c = dynamic_cast<const ConstantTexture<Spectrum> *>(tex.GetPtr())) {
            CopySpectrum(c->value.Clamp(), spec);
            return true;
        }
        if (const ImageTexture<RGBSpectrum, Spectrum> *img = dynamic_cast<const ImageTexture<RGBSpectrum, Spectrum> *>(tex.GetPtr()))
            return AddImageTexture<RGBSpectrum>(img, 1.f, texIdx, what);
        return fail(std::string("texture for ") + what + " is neither constant nor an image map");
    }

    // bump map: the constant-0 default, a float image map, or ScaleTextures of one with constants;
    // the normal map must be the constant-black default (the fork's NormalMap path is not lowered)
    bool BumpParam(const Reference<Texture<float> > &bump, const Reference<Texture<Spectrum> > &normal, int32_t *texIdx) {
        *texIdx = -1;
        Spectrum n;
        if (!ConstTex(normal, &n, "normalmap")) return false;
        if (!n.IsBlack()) return fail("normal map present");
        const Texture<float> *t = bump.GetPtr();
        float scale = 1.f;
        bool scaled = false;
        for (;;) {
            if (const ConstantTexture<float> *c = dynamic_cast<const ConstantTexture<float> *>(t)) {
                if (scaled || c->value != 0.f) return fail("non-zero constant bump map");
                return true;
            }
            if (const ImageTexture<float, float> *img = dynamic_cast<const ImageTexture<float, float> *>(t))
                return AddImageTexture<float>(img, scale, texIdx, "bumpmap");
            if (const ScaleTexture<float, float> *sc = dynamic_cast<const ScaleTexture<float, float> *>(t)) {
                // tex1->Evaluate(dg) * tex2->Evaluate(dg) (scale.h:45-47): one operand must be a constant
                const ConstantTexture<float> *c2 = dynamic_cast<const ConstantTexture<float> *>(sc->tex2.GetPtr());
                if (!c2 || scaled) return fail("bump map: scale texture whose second operand is not a constant (or nested scales)");
                scale = c2->value; scaled = true;
                t = sc->tex1.GetPtr();
                continue;
            }
            return fail("bump map texture is not an image map");
        }
    }

    // rows of SPT_BAND_PITCH floats: the caller zero-fills them, the reference's nSpectralSamples samples go first
    static void CopySpectrum(const Spectrum &s, float *dst) {
        static_assert(nSpectralSamples == SPT_NBANDS, "build the host side with -DSPT_NBANDS equal to the reference's nSpectralSamples");
        for (int i = 0; i < nSpectralSamples; ++i) dst[i] = s.c[i];
    }

    bool AddMaterial(const Material *m, int32_t *idx) {
        SptMaterial row;
        memset(&row, 0, sizeof(row));
        row.tex_kd = row.tex_bump = row.brdf = -1;
        if (const MeasuredMaterial *ms = dynamic_cast<const MeasuredMaterial *>(m)) {
            // measured.cpp:185-205: IrregIsotropicBRDF over the kd-tree of a theta-phi (.brdf) file
            if (!BumpParam(ms->bumpMap, ms->normalMap, &row.tex_bump)) return false;
            row.type = SPT_MAT_MEASURED;
            if (ms->regularHalfangleData) {
                // measured.cpp:122-170: the MERL table as the reference loaded it (scaled, clamped at zero), entry for entry
                const void *key = (const void *)ms->regularHalfangleData;
                std::map<const void *, int>::iterator hit = brdfIdx.find(key);
                if (hit == brdfIdx.end()) {
                    SptBrdfTable t;
                    memset(&t, 0, sizeof(t));
                    t.n_theta_h = ms->nThetaH; t.n_theta_d = ms->nThetaD; t.n_phi_d = ms->nPhiD;
                    t.rgb_offset = out->merl_rgb.size();
                    const size_t n = 3 * (size_t)ms->nThetaH * ms->nThetaD * ms->nPhiD;
                    out->merl_rgb.insert(out->merl_rgb.end(), ms->regularHalfangleData, ms->regularHalfangleData + n);
                    out->brdfs.push_back(t);
                    brdfIdx[key] = (int)out->brdfs.size() - 1;
                    hit = brdfIdx.find(key);
                }
                row.brdf = hit->second;
            } else {
            if (!ms->thetaPhiData) return fail("measured material without BRDF data");
            const KdTree<IrregIsotropicBRDFSample> *kd = ms->thetaPhiData;
            std::map<const void *, int>::iterator it = brdfIdx.find((const void *)kd);
            if (it == brdfIdx.end()) {
                SptBrdfTable t;
                memset(&t, 0, sizeof(t));
                t.node_first = (uint32_t)out->brdf_nodes.size(); t.n_nodes = kd->nNodes;
                for (uint32_t k = 0; k < kd->nNodes; ++k) {
                    SptKdNode n;
                    memset(&n, 0, sizeof(n));
                    n.split_pos = kd->nodes[k].splitAxis == 3 ? 0.f : kd->nodes[k].splitPos;   // a leaf's splitPos is never written (kdtree.h:44-48)
                    n.bits = (uint32_t)kd->nodes[k].splitAxis | (uint32_t)kd->nodes[k].hasLeftChild << 2 | (uint32_t)kd->nodes[k].rightChild << 3;
                    n.p[0] = kd->nodeData[k].p.x; n.p[1] = kd->nodeData[k].p.y; n.p[2] = kd->nodeData[k].p.z;
                    out->brdf_nodes.push_back(n);
                    float v[SPT_BAND_PITCH] = { 0.f };       // one row per spectrum, zero beyond the valid bands
                    CopySpectrum(kd->nodeData[k].v, v);
                    out->brdf_spectra.insert(out->brdf_spectra.end(), v, v + SPT_BAND_PITCH);
                }
                out->brdfs.push_back(t);
                brdfIdx[(const void *)kd] = (int)out->brdfs.size() - 1;
                it = brdfIdx.find((const void *)kd);
            }
            row.brdf = it->second;
            }
        } else if (const MatteMaterial *mm = dynamic_cast<const MatteMaterial *>(m)) {
            float sig;
            if (!BumpParam(mm->bumpMap, mm->normalMap, &row.tex_bump) || !SpectrumParam(mm->Kd, row.spec0, &row.tex_kd, "Kd") ||
                !ConstTex(mm->sigma, &sig, "sigma")) return false;
            row.type = SPT_MAT_MATTE;
            row.p0 = Clamp(sig, 0.f, 90.f);
        } else if (const PlasticMaterial *pm = dynamic_cast<const PlasticMaterial *>(m)) {
            Spectrum ks; float rough;
            if (!BumpParam(pm->bumpMap, pm->normalMap, &row.tex_bump) || !SpectrumParam(pm->Kd, row.spec0, &row.tex_kd, "Kd") ||
                !ConstTex(pm->Ks, &ks, "Ks") || !ConstTex(pm->roughness, &rough, "roughness")) return false;
            row.type = SPT_MAT_PLASTIC;
            CopySpectrum(ks.Clamp(), row.spec1);
            row.p0 = rough;
        } else if (const SubstrateMaterial *sb = dynamic_cast<const SubstrateMaterial *>(m)) {
            Spectrum ks; float nu, nv;                            // substrate.cpp:34-56
            if (!BumpParam(sb->bumpMap, sb->normalMap, &row.tex_bump) || !SpectrumParam(sb->Kd, row.spec0, &row.tex_kd, "Kd") ||
                !ConstTex(sb->Ks, &ks, "Ks") || !ConstTex(sb->nu, &nu, "uroughness") || !ConstTex(sb->nv, &nv, "vroughness")) return false;
            row.type = SPT_MAT_SUBSTRATE;
            CopySpectrum(ks.Clamp(), row.spec1);
            row.p0 = nu; row.p1 = nv;
        } else if (const MetalMaterial *me = dynamic_cast<const MetalMaterial *>(m)) {
            Spectrum eta, k; float rough;
            if (!BumpParam(me->bumpMap, me->normalMap, &row.tex_bump) || !ConstTex(me->eta, &eta, "eta") ||
                !ConstTex(me->k, &k, "k") || !ConstTex(me->roughness, &rough, "roughness")) return false;
            row.type = SPT_MAT_METAL;
            CopySpectrum(eta, row.spec0);
            CopySpectrum(k, row.spec1);
            row.p0 = rough;
        } else if (const MirrorMaterial *mi = dynamic_cast<const MirrorMaterial *>(m)) {
            Spectrum kr;
            if (!BumpParam(mi->bumpMap, mi->normalMap, &row.tex_bump) || !ConstTex(mi->Kr, &kr, "Kr")) return false;
            row.type = SPT_MAT_MIRROR;                            // SpecularReflection(Kr, FresnelNoOp), mirror.cpp:34-55
            CopySpectrum(kr.Clamp(), row.spec0);
        } else if (const GlassMaterial *gl = dynamic_cast<const GlassMaterial *>(m)) {
            Spectrum kr, kt; float ior;
            if (!BumpParam(gl->bumpMap, gl->normalMap, &row.tex_bump) || !ConstTex(gl->Kr, &kr, "Kr") || !ConstTex(gl->Kt, &kt, "Kt") ||
                !ConstTex(gl->index, &ior, "index")) return false;
            row.type = SPT_MAT_GLASS;                             // glass.cpp:34-58
            CopySpectrum(kr.Clamp(), row.spec0);
            CopySpectrum(kt.Clamp(), row.spec1);
            row.p0 = ior;
        } else if (const SubsurfaceMaterial *su = dynamic_cast<const SubsurfaceMaterial *>(m)) {
            // Under a surface integrator that ignores the BSSRDF (path), the material is its BSDF:
            // SpecularReflection(Kr, FresnelDielectric(1, eta)) - subsurface.cpp:40-58 - i.e. glass without Kt.
            Spectrum kr; float eta;
            if (!BumpParam(su->bumpMap, su->normalMap, &row.tex_bump) || !ConstTex(su->Kr, &kr, "Kr") || !ConstTex(su->eta, &eta, "eta")) return false;
            row.type = SPT_MAT_GLASS;
            CopySpectrum(kr.Clamp(), row.spec0);
            row.p0 = eta;
        } else {
            return fail(std::string("unsupported material type ") + typeid(*m).name());
        }
        for (size_t i = 0; i < out->materials.size(); ++i)
            if (!memcmp(&out->materials[i], &row, sizeof(row))) { *idx = (int32_t)i; return true; }
        out->materials.push_back(row);
        *idx = (int32_t)out->materials.size() - 1;
        return true;
    }

    bool AddLights(const Scene *scene, const Sampler *sampler) {
        for (size_t li = 0; li < scene->lights.size(); ++li) {
            const Light *l = scene->lights[li];
            SptLight row;
            memset(&row, 0, sizeof(row));
            row.xform = -1;
            if (const DiffuseAreaLight *al = dynamic_cast<const DiffuseAreaLight *>(l)) {
                row.type = SPT_LIGHT_AREA;
                CopySpectrum(al->Lemit, row.spectrum);
                const ShapeSet *ss = al->shapeSet;
                row.shape_first = (int32_t)out->light_shapes.size();
                row.shape_count = (int32_t)ss->shapes.size();
                row.sum_area = ss->sumArea;
                for (size_t i = 0; i < ss->shapes.size(); ++i) {
                    SptLightShape s;
                    uint8_t kind, flags; uint32_t data; int32_t xf;
                    if (!AddShape(ss->shapes[i].GetPtr(), &kind, &flags, &data, &xf)) return false;
                    s.kind = kind; s.flags = flags; s.data = (int32_t)data; s.area = ss->areas[i];
                    out->light_shapes.push_back(s);
                }
            } else if (const PointLight *pl = dynamic_cast<const PointLight *>(l)) {
                row.type = SPT_LIGHT_POINT;
                CopySpectrum(pl->Intensity, row.spectrum);
                row.pos[0] = pl->lightPos.x; row.pos[1] = pl->lightPos.y; row.pos[2] = pl->lightPos.z;
            } else if (const InfiniteAreaLight *il = dynamic_cast<const InfiniteAreaLight *>(l)) {
                if (out->env_w) return fail("more than one infinite area light");
                row.type = SPT_LIGHT_INFINITE;
                row.xform = AddXform(&il->LightToWorld);
                const MIPMap<RGBSpectrum> *mm = il->radianceMap;
                if (mm->wrapMode != TEXTURE_REPEAT) return fail("env map wrap mode");
                const BlockedArray<RGBSpectrum> *lvl0 = mm->pyramid[0];
                out->env_w = lvl0->uSize(); out->env_h = lvl0->vSize();
                out->env_rgb.resize((size_t)3 * out->env_w * out->env_h);
                for (int v = 0; v < out->env_h; ++v)
                    for (int u = 0; u < out->env_w; ++u) {
                        float rgb[3];
                        (*lvl0)(u, v).ToRGB(rgb);
                        memcpy(&out->env_rgb[3 * ((size_t)v * out->env_w + u)], rgb, 12);
                    }
                const Distribution2D *d2 = il->distribution;
                int nv = (int)d2->pConditionalV.size();
                int nu = d2->pConditionalV[0]->count;
                // the distribution is built over the ORIGINAL image size, which the MIPMap may
                // have resampled to powers of two; the GPU tables carry their own dims
                if (nu != out->env_w || nv != out->env_h)
                    return fail("environment map resolution is not a power of two");
                for (int v = 0; v < nv; ++v) {
                    const Distribution1D *d1 = d2->pConditionalV[v];
                    out->env_func.insert(out->env_func.end(), d1->func, d1->func + nu);
                    out->env_cdf.insert(out->env_cdf.end(), d1->cdf, d1->cdf + nu + 1);
                    out->env_func_int.push_back(d1->funcInt);
                }
                const Distribution1D *dm = d2->pMarginal;
                out->env_marg_func.assign(dm->func, dm->func + nv);
                out->env_marg_cdf.assign(dm->cdf, dm->cdf + nv + 1);
                out->env_marg_int = dm->funcInt;
            } else {
                return fail(std::string("unsupported light type ") + typeid(*l).name());
            }
            row.n_samples = std::max(1, sampler ? ((Sampler *)sampler)->RoundSize(l->nSamples) : l->nSamples);
            lightIdx[l] = (int)out->lights.size();
            out->lights.push_back(row);
        }
        return true;
    }

    bool Run(const Scene *scene, const Camera *camera, const Sampler *sampler,
             const SurfaceIntegrator *surf) {
        if (scene->volumeRegion) return fail("participating media are not supported");
        const BVHAccel *bvh = dynamic_cast<const BVHAccel *>(scene->aggregate);
        if (!bvh) return fail("aggregate is not a BVHAccel");
        if (!bvh->nodes) return fail("empty BVH");
        // node count: the depth-first layout makes the tree the range [0, size(root))
        const LoweringBVHNode *nodes = (const LoweringBVHNode *)bvh->nodes;
        uint32_t nNodes = 0;
        {
            std::vector<uint32_t> todo(1, 0u);
            while (!todo.empty()) {
                uint32_t n = todo.back(); todo.pop_back();
                nNodes = std::max(nNodes, n + 1);
                if (nodes[n].nPrimitives == 0) { todo.push_back(n + 1); todo.push_back(nodes[n].offset); }
            }
        }
        out->bvh_nodes.assign((const uint8_t *)nodes, (const uint8_t *)nodes + (size_t)nNodes * 32);
        // the reference never writes LinearBVHNode::pad, nor a leaf's axis (bvh.cpp:105-115,364-369): zero them so that
        // lowering a scene twice gives the same bytes
        for (uint32_t n = 0; n < nNodes; ++n) {
            uint8_t *nd = &out->bvh_nodes[(size_t)n * 32];
            nd[30] = nd[31] = 0;
            if (nd[28]) nd[29] = 0;
        }

        if (!AddLights(scene, sampler)) return false;

        size_t np = bvh->primitives.size();
        for (size_t i = 0; i < np; ++i) {
            const GeometricPrimitive *gp = dynamic_cast<const GeometricPrimitive *>(bvh->primitives[i].GetPtr());
            if (!gp) return fail("non-geometric primitive (object instance / animated transform) in the BVH");
            uint8_t kind, flags; uint32_t data; int32_t xf, mat;
            if (!AddShape(gp->shape.GetPtr(), &kind, &flags, &data, &xf)) return false;
            if (!AddMaterial(gp->material.GetPtr(), &mat)) return false;
            int32_t light = -1;
            if (gp->areaLight) {
                std::map<const Light *, int>::iterator it = lightIdx.find(gp->areaLight);
                if (it == lightIdx.end()) return fail("area light of a primitive is not in scene->lights");
                light = it->second;
            }
            out->prim_kind.push_back(kind); out->prim_flags.push_back(flags);
            out->prim_id.push_back(gp->primitiveId); out->prim_data.push_back(data);
            out->prim_material.push_back(mat); out->prim_light.push_back(light);
            out->prim_xform.push_back(xf);
        }

        // spectral tables (SampledSpectrum statics, src/core/spectrum.h:297-351)
        for (int i = 0; i < nSpectralSamples; ++i) out->tables.cie_y[i] = SampledSpectrum::Y.c[i];
        out->tables.yint = SampledSpectrum::yint;
        const SampledSpectrum *ill[7] = {
            &SampledSpectrum::rgbIllum2SpectWhite, &SampledSpectrum::rgbIllum2SpectCyan,
            &SampledSpectrum::rgbIllum2SpectMagenta, &SampledSpectrum::rgbIllum2SpectYellow,
            &SampledSpectrum::rgbIllum2SpectRed, &SampledSpectrum::rgbIllum2SpectGreen,
            &SampledSpectrum::rgbIllum2SpectBlue };
        const SampledSpectrum *refl[7] = {
            &SampledSpectrum::rgbRefl2SpectWhite, &SampledSpectrum::rgbRefl2SpectCyan,
            &SampledSpectrum::rgbRefl2SpectMagenta, &SampledSpectrum::rgbRefl2SpectYellow,
            &SampledSpectrum::rgbRefl2SpectRed, &SampledSpectrum::rgbRefl2SpectGreen,
            &SampledSpectrum::rgbRefl2SpectBlue };
        for (int k = 0; k < 7; ++k)
            for (int i = 0; i < nSpectralSamples; ++i) {
                out->tables.rgb_illum[k][i] = ill[k]->c[i];
                out->tables.rgb_refl[k][i] = refl[k]->c[i];
            }

        // camera
        const PerspectiveCamera *pc = dynamic_cast<const PerspectiveCamera *>(camera);
        if (!pc) return fail("camera is not a PerspectiveCamera");
        if (pc->CameraToWorld.actuallyAnimated) return fail("animated camera");
        memcpy(out->camera.raster_to_camera, pc->RasterToCamera.m.m, 64);
        memcpy(out->camera.camera_to_world, pc->CameraToWorld.startTransform->m.m, 64);
        out->camera.lens_radius = pc->lensRadius;
        out->camera.focal_distance = pc->focalDistance;
        out->camera.shutter_open = pc->shutterOpen;
        out->camera.shutter_close = pc->shutterClose;
        for (int k = 0; k < 3; ++k) { out->camera.dx_camera[k] = pc->dxCamera[k]; out->camera.dy_camera[k] = pc->dyCamera[k]; }

        // film
        const SpectralImageFilm *film = dynamic_cast<const SpectralImageFilm *>(camera->film);
        if (!film) return fail("film is not a SpectralImageFilm");
        out->film.x_resolution = film->xResolution; out->film.y_resolution = film->yResolution;
        out->film.x_pixel_start = film->xPixelStart; out->film.y_pixel_start = film->yPixelStart;
        out->film.x_pixel_count = film->xPixelCount; out->film.y_pixel_count = film->yPixelCount;
        out->film.filter_xwidth = film->filter->xWidth; out->film.filter_ywidth = film->filter->yWidth;
        out->film.filter_inv_xwidth = film->filter->invXWidth;
        out->film.filter_inv_ywidth = film->filter->invYWidth;
        memcpy(out->film.filter_table, film->filterTable, 256 * sizeof(float));
        out->film_filename = film->imageOutputName;

        // sampler + integrator
        const LDSampler *ld = dynamic_cast<const LDSampler *>(sampler);
        if (!ld) return fail("sampler is not the low-discrepancy sampler");
        out->params.spp = ld->nPixelSamples;
        if (const PathIntegrator *pi = dynamic_cast<const PathIntegrator *>(surf)) {
            out->params.integrator = SPT_INTEGRATOR_PATH;
            out->params.max_depth = pi->maxDepth;
        } else if (const DirectLightingIntegrator *dl = dynamic_cast<const DirectLightingIntegrator *>(surf)) {
            // directlighting.cpp:70-105: UniformSampleAllLights (strategy "all", the default) or UniformSampleOneLight at every hit, and the
            // SpecularReflect / SpecularTransmit recursion (integrator.cpp:169-250) down to maxDepth - which only does anything for
            // specular BxDFs (mirror, glass, the subsurface material's reflection)
            out->params.integrator = dl->strategy == SAMPLE_ALL_UNIFORM ? SPT_INTEGRATOR_DIRECT_ALL : SPT_INTEGRATOR_DIRECT_ONE;
            out->params.max_depth = dl->maxDepth;
        } else return fail("surface integrator is neither the path nor the directlighting integrator");
        out->params.x_start = ld->xPixelStart; out->params.x_end = ld->xPixelEnd;
        out->params.y_start = ld->yPixelStart; out->params.y_end = ld->yPixelEnd;
        out->params.seed = 0;
        out->params.tile_rank = 0; out->params.tile_nranks = 1;
        return true;
    }
};

}  // namespace

bool LowerScene(const Scene *scene, const Camera *camera, const Sampler *sampler,
                const SurfaceIntegrator *surf, LoweredScene *out, std::string *err) {
    Lowerer L;
    L.out = out;
    bool ok = L.Run(scene, camera, sampler, surf);
    if (!ok && err) *err = L.err;
    return ok;
}

// TEST INFRASTRUCTURE (not part of the product): the product's device functions compiled for the host with g++
// (fake/cuda_runtime.h) and exported with a C interface, so that tests can compare them with the oracle bit for bit on a
// machine without a GPU. Same flags as the oracle (-O2 -ffp-contract=off): same libm, same rounding.
#include "shade.cuh"
#include "camera.cuh"

#include <vector>
static std::vector<float> g_light_cdf;     // per-call scratch of the single-threaded tests

static void dev_scene(const SptSceneDesc *d, DevScene *v) {
    memset(v, 0, sizeof(*v));
    // Distribution1D of each area light's ShapeSet, as spt_scene_create builds it (csrc/spt_api.cu)
    g_light_cdf.assign(d->n_light_shapes + d->n_lights + 1, 0.f);
    for (uint32_t li = 0; li < d->n_lights; ++li) {
        const SptLight &l = d->lights[li];
        if (l.type != SPT_LIGHT_AREA) continue;
        float *c = &g_light_cdf[l.shape_first + li];
        int n = l.shape_count;
        c[0] = 0.f;
        for (int i = 1; i < n + 1; ++i) c[i] = c[i - 1] + d->light_shapes[l.shape_first + i - 1].area / n;
        float funcInt = c[n];
        if (funcInt == 0.f) for (int i = 1; i < n + 1; ++i) c[i] = float(i) / float(n);
        else for (int i = 1; i < n + 1; ++i) c[i] /= funcInt;
    }
    v->light_cdf = g_light_cdf.data();
    v->env_func = d->env_func; v->env_cdf = d->env_cdf; v->env_func_int = d->env_func_int;
    v->env_marg_func = d->env_marg_func; v->env_marg_cdf = d->env_marg_cdf; v->env_marg_int = d->env_marg_int;
    v->n_nodes = d->n_nodes; v->n_prims = d->n_prims;
    v->prim_kind = d->prim_kind; v->prim_flags = d->prim_flags; v->prim_id = d->prim_id; v->prim_data = d->prim_data;
    v->prim_material = d->prim_material; v->prim_light = d->prim_light; v->prim_xform = d->prim_xform;
    v->tri_vidx = d->tri_vidx; v->P = d->P; v->N = d->N; v->UV = d->UV;
    v->quadrics = d->quadrics; v->xforms = d->xforms; v->materials = d->materials; v->lights = d->lights;
    v->light_shapes = d->light_shapes; v->n_lights = d->n_lights; v->tables = &d->tables;
    v->env_w = d->env_w; v->env_h = d->env_h; v->env_rgb = d->env_rgb;
    v->textures = d->textures; v->tex_texels = d->tex_texels; v->ewa_lut = d->ewa_weight_lut;
    v->brdfs = d->brdfs; v->brdf_nodes = d->brdf_nodes; v->brdf_spectra = d->brdf_spectra;
    v->has_ext = 1; v->has_measured = d->n_brdfs > 0;
}

extern "C" {

// measured_f (IrregIsotropicBRDF::f) for n direction pairs in the BSDF's local frame
void hd_measured_f(const SptSceneDesc *d, int table, const float *wo, const float *wi, int n, float *out) {
    DevScene sc; dev_scene(d, &sc);
    for (int i = 0; i < n; ++i)
        measured_f(sc, sc.brdfs[table], V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), out + (size_t)NB * i);
}

// ImageTexture::Evaluate for n look-ups: uvd = {u, v, dudx, dvdx, dudy, dvdy}; out: `channels` floats each
void hd_tex_evaluate(const SptSceneDesc *d, int tex, const float *uvd, int n, float *out) {
    DevScene sc; dev_scene(d, &sc);
    const SptTexture &t = sc.textures[tex];
    for (int i = 0; i < n; ++i) {
        UVDiff df; df.dudx = uvd[6 * i + 2]; df.dvdx = uvd[6 * i + 3]; df.dudy = uvd[6 * i + 4]; df.dvdy = uvd[6 * i + 5];
        if (t.channels == 3) tex_evaluate<3>(sc, t, uvd[6 * i], uvd[6 * i + 1], df, out + 3 * (size_t)i);
        else tex_evaluate<1>(sc, t, uvd[6 * i], uvd[6 * i + 1], df, out + (size_t)i);
    }
}

// The shading frame K5 builds at the first hit of n camera samples {imageX, imageY, lensU, lensV, time} that hit BVH slot
// slot[i] at distance t[i]: camera ray + its offset rays (scaled for spp), hit record, differentials, bump map, image-mapped
// Kd. out: 12 floats per sample {nn, sn, tn, kd_rgb}.
void hd_first_vertex_frame(const SptSceneDesc *d, const SptCameraDesc *cam, int spp, const float *samples, const uint32_t *slot,
                           const float *t, int n, float *out) {
    DevScene sc; dev_scene(d, &sc);
    const float scale = 1.f / sqrtf((float)spp);
    for (int i = 0; i < n; ++i) {
        const float *s = samples + 5 * (size_t)i;
        Ray ray;
        camera_ray(*cam, s[0], s[1], s[2], s[3], &ray);
        RayDiff rd;
        camera_ray_diff(*cam, s[0], s[1], scale, ray, &rd);
        Hit hit;
        shape_record(sc, sc.prim_kind[slot[i]], sc.prim_flags[slot[i]], sc.prim_data[slot[i]], ray, t[i], &hit);
        Bsdf b;
        b.kd_rgb[0] = b.kd_rgb[1] = b.kd_rgb[2] = 0.f;
        make_bsdf<true>(sc, slot[i], hit, &rd, &b);
        float *o = out + 12 * (size_t)i;
        o[0] = b.nn.x; o[1] = b.nn.y; o[2] = b.nn.z; o[3] = b.sn.x; o[4] = b.sn.y; o[5] = b.sn.z;
        o[6] = b.tn.x; o[7] = b.tn.y; o[8] = b.tn.z; o[9] = b.kd_rgb[0]; o[10] = b.kd_rgb[1]; o[11] = b.kd_rgb[2];
    }
}

// Light::Sample_L (+ Light::Pdf of the sampled direction) for n points p and sample values {uPos0, uPos1, uComponent}:
// out 9 floats {wi, pdf, shadow direction, shadow maxt, black}
void hd_light_sample(const SptSceneDesc *d, int light, const float *p, const float *u, int n, float *out) {
    DevScene sc; dev_scene(d, &sc);
    for (int i = 0; i < n; ++i) {
        LightSampleResult lr;
        light_sample(sc, light, V(p[3 * i], p[3 * i + 1], p[3 * i + 2]), u[3 * i], u[3 * i + 1], u[3 * i + 2], &lr, false);
        float *o = out + 9 * (size_t)i;
        o[0] = lr.wi.x; o[1] = lr.wi.y; o[2] = lr.wi.z; o[3] = lr.pdf;
        o[4] = lr.shadow_d.x; o[5] = lr.shadow_d.y; o[6] = lr.shadow_d.z; o[7] = lr.shadow_maxt; o[8] = lr.black ? 1.f : 0.f;
    }
}
// Light::Pdf(p, w) for n points and directions
void hd_light_pdf(const SptSceneDesc *d, int light, const float *p, const float *w, int n, float *out) {
    DevScene sc; dev_scene(d, &sc);
    for (int i = 0; i < n; ++i)
        out[i] = light_pdf(sc, light, V(p[3 * i], p[3 * i + 1], p[3 * i + 2]), V(w[3 * i], w[3 * i + 1], w[3 * i + 2]));
}

}  // extern "C"

// oracle_dump.cpp — TEST INFRASTRUCTURE (not part of the product path).
//
// Links the UNMODIFIED reference objects (oracle/_ref/lib/libpbrt_ref.a) and registers itself under
// the same `Renderer "gpupath"` hook as the product renderer, but instead of rendering it writes
// golden vectors produced by the reference's own code:
//   <prefix>.spt     the lowered scene (same lowering the product uses)
//   <prefix>.golden  per-sample records: the reference's LDPixelSample vectors
//                    (src/core/montecarlo.cpp:192-244), camera rays
//                    (src/cameras/perspective.cpp:73-106), first hits from Scene::Intersect
//                    (src/accelerators/bvh.cpp:380-432), secondary closest/any-hit rays, and the
//                    radiance SamplerRenderer::Li returns (src/renderers/samplerrenderer.cpp:225-247)
//                    together with the RNG floats it consumed.
// Environment: SPT_DUMP_PREFIX (required), SPT_DUMP_PIXELS (default 2000), SPT_DUMP_NRNG (64),
//              SPT_DUMP_LI (1: also record radiance; 0: rays only),
//              SPT_DUMP_COMPACT (1: first hits only, 16 bytes per camera ray - {imageX, imageY}, primitive id, t - for the
//              parity tests at BASELINE resolution: a million rays per config; pinhole cameras only).
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "gpupath.h"
#include "lowering.h"
#include "paramset.h"
#include "scene.h"
#include "camera.h"
#include "film.h"
#include "sampler.h"
#include "montecarlo.h"
#include "intersection.h"
#include "rng.h"
#include "memory.h"
#include "renderers/samplerrenderer.h"

namespace {

class DumpRenderer : public Renderer {
public:
    DumpRenderer(Sampler *s, Camera *c, SurfaceIntegrator *si, VolumeIntegrator *vi, bool visIds)
        : sampler(s), camera(c), surf(si), vol(vi) {
        cpu = new SamplerRenderer(s, c, si, vi, visIds);
    }
    ~DumpRenderer() { delete cpu; }

    void Render(const Scene *scene) {
        const char *prefix = getenv("SPT_DUMP_PREFIX");
        if (!prefix) Severe("oracle_dump: set SPT_DUMP_PREFIX");
        int nPixelsWanted = getenv("SPT_DUMP_PIXELS") ? atoi(getenv("SPT_DUMP_PIXELS")) : 2000;
        int nRng = getenv("SPT_DUMP_NRNG") ? atoi(getenv("SPT_DUMP_NRNG")) : 64;
        bool doLi = getenv("SPT_DUMP_LI") ? atoi(getenv("SPT_DUMP_LI")) != 0 : true;
        bool compact = getenv("SPT_DUMP_COMPACT") ? atoi(getenv("SPT_DUMP_COMPACT")) != 0 : false;
        if (compact) doLi = false;

        LoweredScene ls;
        std::string why;
        if (!LowerScene(scene, camera, sampler, surf, &ls, &why))
            Severe("oracle_dump: lowering failed: %s", why.c_str());
        if (!ls.Save(std::string(prefix) + ".spt", &why)) Severe("oracle_dump: %s", why.c_str());

        // the reference's own sample layout (src/renderers/samplerrenderer.cpp:197)
        Sample *origSample = new Sample(sampler, surf, vol, scene);
        int spp = sampler->samplesPerPixel;
        Sample *samples = origSample->Duplicate(spp);
        std::vector<float> buf(LDPixelSampleFloatsNeeded(samples, spp));
        int nExtra = 0;
        for (size_t i = 0; i < origSample->n1D.size(); ++i) nExtra += origSample->n1D[i];
        for (size_t i = 0; i < origSample->n2D.size(); ++i) nExtra += 2 * origSample->n2D[i];
        int nSampleFloats = 5 + nExtra;

        int x0 = sampler->xPixelStart, x1 = sampler->xPixelEnd;
        int y0 = sampler->yPixelStart, y1 = sampler->yPixelEnd;
        long long extent = (long long)(x1 - x0) * (y1 - y0);
        int nPix = (int)std::min<long long>(nPixelsWanted, extent);

        std::vector<float> outSamples, outRays, outT, outL, outRng, outRays2, outT2, outXY;
        std::vector<uint32_t> outId, outId2, outPixel;
        std::vector<uint8_t> outAny;
        MemoryArena arena;
        uint32_t lcg = 12345u;
        for (int k = 0; k < nPix; ++k) {
            int px, py;
            if (nPix == extent) { px = x0 + k % (x1 - x0); py = y0 + k / (x1 - x0); }
            else {
                lcg = lcg * 1664525u + 1013904223u; px = x0 + (int)((lcg >> 8) % (uint32_t)(x1 - x0));
                lcg = lcg * 1664525u + 1013904223u; py = y0 + (int)((lcg >> 8) % (uint32_t)(y1 - y0));
            }
            RNG pixRng(7919u * (uint32_t)k + 17u);
            LDPixelSample(px, py, sampler->shutterOpen, sampler->shutterClose, spp, samples, &buf[0], pixRng);
            for (int i = 0; i < spp; ++i) {
                Sample &s = samples[i];
                if (compact) {
                    // first hit of the camera ray only, by the reference's own camera and BVH
                    RayDifferential cray;
                    camera->GenerateRayDifferential(s, &cray);
                    Ray probe(cray);
                    Intersection isect;
                    bool hit = scene->Intersect(probe, &isect);
                    outXY.push_back(s.imageX); outXY.push_back(s.imageY);
                    outId.push_back(hit ? isect.primitiveId : 0u);
                    outT.push_back(probe.maxt);
                    continue;
                }
                outPixel.push_back((uint32_t)px); outPixel.push_back((uint32_t)py);
                outSamples.push_back(s.imageX); outSamples.push_back(s.imageY);
                outSamples.push_back(s.lensU); outSamples.push_back(s.lensV); outSamples.push_back(s.time);
                for (size_t j = 0; j < s.n1D.size(); ++j)
                    for (uint32_t q = 0; q < s.n1D[j]; ++q) outSamples.push_back(s.oneD[j][q]);
                for (size_t j = 0; j < s.n2D.size(); ++j)
                    for (uint32_t q = 0; q < 2 * s.n2D[j]; ++q) outSamples.push_back(s.twoD[j][q]);

                RayDifferential ray;
                camera->GenerateRayDifferential(s, &ray);
                ray.ScaleDifferentials(1.f / sqrtf(sampler->samplesPerPixel));
                float r8[8] = { ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, ray.mint, ray.maxt };
                outRays.insert(outRays.end(), r8, r8 + 8);

                // first hit with the reference BVH
                Ray probe(ray);
                Intersection isect;
                bool hit = scene->Intersect(probe, &isect);
                outId.push_back(hit ? isect.primitiveId : 0u);
                outT.push_back(probe.maxt);

                // a secondary ray from the hit point: closest hit (unbounded) + any hit (bounded)
                float r2[8] = { 0, 0, 0, 0, 0, 1, 0, 0 };
                uint32_t id2 = 0; float t2 = 0.f; uint8_t any = 0;
                if (hit) {
                    lcg = lcg * 1664525u + 1013904223u; float u1 = (lcg >> 8) / 16777216.f;
                    lcg = lcg * 1664525u + 1013904223u; float u2 = (lcg >> 8) / 16777216.f;
                    lcg = lcg * 1664525u + 1013904223u; float u3 = (lcg >> 8) / 16777216.f;
                    Vector dir = UniformSampleSphere(u1, u2);
                    Point p = probe(probe.maxt);
                    Ray sec(p, dir, isect.rayEpsilon, INFINITY);
                    Intersection isect2;
                    bool hit2 = scene->Intersect(sec, &isect2);
                    id2 = hit2 ? isect2.primitiveId : 0u;
                    t2 = sec.maxt;
                    // bounded segment for IntersectP: random length up to 1.5x the closest distance
                    float seg = hit2 ? sec.maxt * 1.5f * u3 : 1000.f * u3;
                    Ray sh(p, dir, isect.rayEpsilon, seg);
                    any = scene->IntersectP(sh) ? 1 : 0;
                    float tmp[8] = { p.x, p.y, p.z, dir.x, dir.y, dir.z, isect.rayEpsilon, seg };
                    memcpy(r2, tmp, sizeof(tmp));
                }
                outRays2.insert(outRays2.end(), r2, r2 + 8);
                outId2.push_back(id2); outT2.push_back(t2); outAny.push_back(any);

                if (doLi) {
                    uint32_t seed = 1000u + (uint32_t)(k * spp + i);
                    RNG pre(seed);
                    for (int q = 0; q < nRng; ++q) outRng.push_back(pre.RandomFloat());
                    RNG rng(seed);
                    Intersection isectLi;
                    Spectrum T;
                    Spectrum L = cpu->Li(scene, ray, &s, rng, arena, &isectLi, &T);
                    for (int b = 0; b < nSpectralSamples; ++b) {
                        float c[nSpectralSamples];
                        L.GetOrigC(c);
                        outL.push_back(c[b]);
                    }
                    arena.FreeAll();
                }
            }
        }
        SptContainerWriter w;
        if (!w.begin(std::string(prefix) + ".golden")) Severe("oracle_dump: cannot write golden file");
        int32_t meta[4] = { spp, nSampleFloats, nRng, nSpectralSamples };
        w.put("meta", 1, meta, sizeof(meta), 4);
        if (compact) {
            if (ls.camera.lens_radius != 0.f) Severe("oracle_dump: SPT_DUMP_COMPACT needs a pinhole camera");
            w.vec("image_xy", 3, outXY, 2);
            w.vec("prim_id", 2, outId);
            w.vec("t_hit", 3, outT);
            w.end();
            fprintf(stderr, "oracle_dump: wrote %s.spt and %s.golden (compact: %d pixels x %d spp first hits)\n", prefix, prefix, nPix, spp);
            delete origSample;
            return;
        }
        w.vec("pixel", 2, outPixel, 2);
        w.vec("samples", 3, outSamples, nSampleFloats);
        w.vec("rays", 3, outRays, 8);
        w.vec("prim_id", 2, outId);
        w.vec("t_hit", 3, outT);
        w.vec("rays2", 3, outRays2, 8);
        w.vec("prim_id2", 2, outId2);
        w.vec("t_hit2", 3, outT2);
        w.vec("any2", 0, outAny);
        if (doLi) {
            w.vec("rng", 3, outRng, nRng);
            w.vec("L", 3, outL, nSpectralSamples);
        }
        w.end();
        fprintf(stderr, "oracle_dump: wrote %s.spt and %s.golden (%d pixels x %d spp)\n",
                prefix, prefix, nPix, spp);
        delete origSample;
    }
    Spectrum Li(const Scene *scene, const RayDifferential &ray, const Sample *sample, RNG &rng,
                MemoryArena &arena, Intersection *isect, Spectrum *T) const {
        return cpu->Li(scene, ray, sample, rng, arena, isect, T);
    }
    Spectrum Transmittance(const Scene *scene, const RayDifferential &ray, const Sample *sample,
                           RNG &rng, MemoryArena &arena) const {
        return cpu->Transmittance(scene, ray, sample, rng, arena);
    }
private:
    Sampler *sampler;
    Camera *camera;
    SurfaceIntegrator *surf;
    VolumeIntegrator *vol;
    SamplerRenderer *cpu;
};

}  // namespace

Renderer *CreateGpuPathRenderer(const ParamSet &params, Sampler *sampler, Camera *camera,
                                SurfaceIntegrator *surf, VolumeIntegrator *vol, bool visIds) {
    return new DumpRenderer(sampler, camera, surf, vol, visIds);
}
void GpuPathTouchParams(const ParamSet &) {}

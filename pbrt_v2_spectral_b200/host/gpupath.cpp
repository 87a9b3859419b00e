// gpupath.cpp — `Renderer "gpupath"`: host-side mirror of SamplerRenderer
// (src/renderers/samplerrenderer.cpp:166-260) that hands the hot path to libspt.so.
//
// Render() lowers the built Scene (lowering.cpp) and calls the C ABI of include/spt.h:
//   spt_scene_create -> spt_film_create -> spt_render -> spt_film_write_dat
// writing the same .dat file SpectralImageFilm::WriteImage would (src/film/spectralImage.cpp:267-378).
// Li()/Transmittance() (only called by other CPU code, e.g. SpecularReflect) delegate to an
// embedded SamplerRenderer, which also owns sampler/camera/integrators exactly as the reference
// does (src/renderers/samplerrenderer.cpp:180-185). Anything the GPU path does not implement is
// reported with Error() and rendered by that SamplerRenderer instead — never approximated.
#include <dlfcn.h>
#include <stdlib.h>
#include <string>
#include "gpupath.h"
#include "lowering.h"
#include "paramset.h"
#include "scene.h"
#include "camera.h"
#include "film.h"
#include "renderers/samplerrenderer.h"

namespace {

struct SptApi {
    void *handle;
    const char *(*last_error)(void);
    SptScene *(*scene_create)(const SptSceneDesc *);
    void (*scene_destroy)(SptScene *);
    SptFilm *(*film_create)(const SptFilmDesc *);
    void (*film_destroy)(SptFilm *);
    int (*render)(SptScene *, const SptCameraDesc *, SptFilm *, const SptRenderParams *);
    int (*film_write_dat)(SptFilm *, const char *);
    int (*get_stats)(SptScene *, SptStats *);
    SptMulti *(*multi_create)(const SptSceneDesc *, const SptFilmDesc *, int, const int *);
    void (*multi_destroy)(SptMulti *);
    int (*multi_device_count)(SptMulti *);
    int (*multi_render)(SptMulti *, const SptCameraDesc *, const SptRenderParams *);
    SptFilm *(*multi_film)(SptMulti *);
    int (*multi_get_stats)(SptMulti *, int, SptStats *);
    bool Load(std::string *err) {
        const char *path = getenv("SPT_LIB");
        handle = dlopen(path ? path : "libspt.so", RTLD_NOW | RTLD_LOCAL);
        if (!handle) { *err = dlerror(); return false; }
#define SPT_SYM(member, name) \
        *(void **)(&member) = dlsym(handle, name); \
        if (!member) { *err = std::string("missing symbol ") + name; return false; }
        SPT_SYM(last_error, "spt_last_error");
        SPT_SYM(scene_create, "spt_scene_create");
        SPT_SYM(scene_destroy, "spt_scene_destroy");
        SPT_SYM(film_create, "spt_film_create");
        SPT_SYM(film_destroy, "spt_film_destroy");
        SPT_SYM(render, "spt_render");
        SPT_SYM(film_write_dat, "spt_film_write_dat");
        SPT_SYM(get_stats, "spt_get_stats");
        SPT_SYM(multi_create, "spt_multi_create");
        SPT_SYM(multi_destroy, "spt_multi_destroy");
        SPT_SYM(multi_device_count, "spt_multi_device_count");
        SPT_SYM(multi_render, "spt_multi_render");
        SPT_SYM(multi_film, "spt_multi_film");
        SPT_SYM(multi_get_stats, "spt_multi_get_stats");
#undef SPT_SYM
        return true;
    }
};

class GpuPathRenderer : public Renderer {
public:
    GpuPathRenderer(Sampler *s, Camera *c, SurfaceIntegrator *si, VolumeIntegrator *vi, bool visIds,
                    int seed, int gpus)
        : sampler(s), camera(c), surf(si), seed(seed), gpus(gpus) {
        cpu = new SamplerRenderer(s, c, si, vi, visIds);
    }
    ~GpuPathRenderer() { delete cpu; }   // owns sampler, camera, integrators

    void Render(const Scene *scene) {
        LoweredScene ls;
        std::string why;
        if (!LowerScene(scene, camera, sampler, surf, &ls, &why)) {
            Error("Renderer \"gpupath\": %s; rendering with the CPU SamplerRenderer instead.", why.c_str());
            cpu->Render(scene);
            return;
        }
        SptApi api;
        if (!api.Load(&why))
            Severe("Renderer \"gpupath\": cannot load libspt.so (%s). Set SPT_LIB.", why.c_str());
        SptSceneDesc desc = ls.Desc();
        ls.params.seed = (uint64_t)seed;
        if (gpus != 1) { RenderMulti(scene, api, ls, desc); return; }
        SptScene *gs = api.scene_create(&desc);
        if (!gs) Severe("Renderer \"gpupath\": spt_scene_create failed: %s", api.last_error());
        SptFilm *gf = api.film_create(&ls.film);
        if (!gf) Severe("Renderer \"gpupath\": spt_film_create failed: %s", api.last_error());
        int rc = api.render(gs, &ls.camera, gf, &ls.params);
        if (rc == SPT_ERR_UNSUPP) {
            // the library knows a combination it does not implement that the lowering let through (today: none): the
            // reference's own renderer takes the scene, nothing is approximated. Any other failure is fatal - no CPU path.
            Error("Renderer \"gpupath\": %s; rendering with the CPU SamplerRenderer instead.", api.last_error());
            api.film_destroy(gf);
            api.scene_destroy(gs);
            cpu->Render(scene);
            return;
        }
        if (rc != SPT_OK)
            Severe("Renderer \"gpupath\": spt_render failed: %s", api.last_error());
        SptStats st;
        if (api.get_stats(gs, &st) == SPT_OK)
            Info("gpupath: %llu camera samples, %llu closest-hit + %llu shadow rays, %.1f ms on the GPU",
                 (unsigned long long)st.camera_samples, (unsigned long long)st.closest_rays,
                 (unsigned long long)st.any_rays, st.render_ms);
        // same naming rule as SpectralImageFilm::WriteImage (src/film/spectralImage.cpp:340-342)
        std::string name = camera->film->imageOutputName;
        size_t dot = name.find_last_of(".");
        std::string out = name.substr(0, dot) + ".dat";
        if (api.film_write_dat(gf, out.c_str()) != SPT_OK)
            Error("Renderer \"gpupath\": writing %s failed: %s", out.c_str(), api.last_error());
        api.film_destroy(gf);
        api.scene_destroy(gs);
    }
    // "integer gpus" [N] (N > 1, or 0 = every visible GPU): scene replicated, image tile sets per GPU, one film on the first
    // GPU that every GPU's film kernel adds into over NVLink (spt_multi_*, include/spt.h)
    void RenderMulti(const Scene *scene, SptApi &api, LoweredScene &ls, const SptSceneDesc &desc) {
        SptMulti *gm = api.multi_create(&desc, &ls.film, gpus, NULL);
        if (!gm) Severe("Renderer \"gpupath\": spt_multi_create (%d GPUs) failed: %s", gpus, api.last_error());
        int rc = api.multi_render(gm, &ls.camera, &ls.params);
        if (rc == SPT_ERR_UNSUPP) {
            Error("Renderer \"gpupath\": %s; rendering with the CPU SamplerRenderer instead.", api.last_error());
            api.multi_destroy(gm);
            cpu->Render(scene);
            return;
        }
        if (rc != SPT_OK) Severe("Renderer \"gpupath\": spt_multi_render failed: %s", api.last_error());
        int n = api.multi_device_count(gm);
        for (int k = 0; k < n; ++k) {
            SptStats st;
            if (api.multi_get_stats(gm, k, &st) == SPT_OK)
                Info("gpupath: GPU %d of %d: %llu camera samples, %llu closest-hit + %llu shadow rays, %.1f ms", k, n,
                     (unsigned long long)st.camera_samples, (unsigned long long)st.closest_rays,
                     (unsigned long long)st.any_rays, st.render_ms);
        }
        std::string name = camera->film->imageOutputName;
        size_t dot = name.find_last_of(".");
        std::string out = name.substr(0, dot) + ".dat";
        if (api.film_write_dat(api.multi_film(gm), out.c_str()) != SPT_OK)
            Error("Renderer \"gpupath\": writing %s failed: %s", out.c_str(), api.last_error());
        api.multi_destroy(gm);
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
    SamplerRenderer *cpu;
    int seed;
    int gpus;
};

}  // namespace

Renderer *CreateGpuPathRenderer(const ParamSet &params, Sampler *sampler, Camera *camera,
                                SurfaceIntegrator *surf, VolumeIntegrator *vol, bool visIds) {
    int seed = params.FindOneInt("seed", 0);
    int gpus = params.FindOneInt("gpus", 1);
    if (gpus < 0) { Warning("Renderer \"gpupath\": \"integer gpus\" [%d] is not a GPU count; using 1.", gpus); gpus = 1; }
    return new GpuPathRenderer(sampler, camera, surf, vol, visIds, seed, gpus);
}

// Marks the renderer's own parameters as looked up. MakeRenderer calls RendererParams.ReportUnused() BEFORE it creates the
// renderer (src/core/api.cpp:1381), so the registration patch calls this first: no "unused parameter" warning for them.
void GpuPathTouchParams(const ParamSet &params) {
    params.FindOneInt("seed", 0);
    params.FindOneInt("gpus", 1);
}

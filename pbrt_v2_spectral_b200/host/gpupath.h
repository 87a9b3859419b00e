// gpupath.h — `Renderer "gpupath"`: the reference-side class a scene names to opt in to the B200
// path (replaces SamplerRenderer, src/renderers/samplerrenderer.h:37-83, behind the abstract
// Renderer of src/core/renderer.h:35-46). See INTEGRATION.md for the api.cpp registration.
#ifndef SPT_HOST_GPUPATH_H
#define SPT_HOST_GPUPATH_H
#include "pbrt.h"
#include "renderer.h"

class ParamSet;
Renderer *CreateGpuPathRenderer(const ParamSet &params, Sampler *sampler, Camera *camera,
                                SurfaceIntegrator *surf, VolumeIntegrator *vol, bool visIds);
// looks the renderer's parameters up ("integer seed", "integer gpus") so that ReportUnused(), which MakeRenderer calls before
// the renderer exists (src/core/api.cpp:1381), does not warn about them
void GpuPathTouchParams(const ParamSet &params);
#endif

# The whole reference-side patch for `Renderer "gpupath"` (src/core/api.cpp:1333-1420), applied to
# a generated copy of api.cpp at build time (oracle/Makefile); see INTEGRATION.md.
s|^#include "api.h"|#include "api.h"\n#include "gpupath.h"|
s|RendererName != "cameras")|RendererName != "cameras" \&\& RendererName != "gpupath")|
s|^\( *\)else if(RendererName == "cameras"){|\1else if (RendererName == "gpupath") {\n\1    renderer = CreateGpuPathRenderer(RendererParams, sampler, camera, surfaceIntegrator, volumeIntegrator, visIds);\n\1}\n\1else if(RendererName == "cameras"){|
# the renderer's own parameters ("integer seed", "integer gpus") are looked up before ReportUnused() runs (api.cpp:1381)
s|^\( *\)string samplingMethod = RendererParams.FindOneString(|\1if (RendererName == "gpupath") GpuPathTouchParams(RendererParams);\n\1string samplingMethod = RendererParams.FindOneString(|

// launch.h — host-callable launchers of the two kernel translation units (spt_exact.cu: camera +
// traversal, no FMA contraction; spt_shade.cu: shading + film, FMA on). spt_api.cu schedules them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "spt.h"
#include "wave.cuh"
#include "sampler_source.h"
#include "trace_args.h"

void spt_launch_gen_camera(int grid, cudaStream_t st, const RenderCfg &cfg, const SampleSource &src, const WaveBuffers &wb, uint32_t *count_out);
// variant 1: ONE launch of the pair-node kernel over all segments (merge) or one per segment; variant 0 (trees that do not pack):
// the reference-layout kernel, one launch per segment. a.work points at TRACE_MAX_SEG zeroed words.
void spt_launch_trace_multi(int variant, bool merge, bool count, int grid, cudaStream_t st, const DevScene &sc, const TraceMultiArgs &a);
void spt_launch_camera_rays(cudaStream_t st, const SptCameraDesc &cam, const float *samples, uint32_t n, float *out);
void spt_launch_split_rays(cudaStream_t st, const float *rays, uint32_t n, float4 *ro, float4 *rd);
void spt_launch_slot_to_id(cudaStream_t st, const uint32_t *slot, const uint32_t *prim_id, uint32_t n, uint32_t *out);
void spt_launch_slot_to_flag(cudaStream_t st, const uint32_t *slot, uint32_t n, uint8_t *out);

void spt_launch_compact_hits(int grid, cudaStream_t st, const uint32_t *queue, const uint32_t *count, const uint32_t *hit_slot,
                             uint32_t *hit_queue, uint32_t *hit_count, uint32_t *miss_queue, uint32_t *miss_count, float *black_L);
void spt_launch_miss_env(int grid, cudaStream_t st, const DevScene &sc, const WaveBuffers &wb, int bounce, int tree, const uint32_t *queue, const uint32_t *count);
void spt_launch_spawn_T(int grid, cudaStream_t st, const DevScene &sc, const WaveBuffers &wb, int level, const uint32_t *queue, const uint32_t *count);
void spt_launch_shade(int grid, cudaStream_t st, const DevScene &sc, const RenderCfg &cfg, const SampleSource &src, const WaveBuffers &wb,
                      int bounce, const uint32_t *queue, const uint32_t *count, uint32_t *shadow_count, uint32_t *mis_count,
                      uint32_t *elided_count, uint32_t *mis_any_count, uint32_t *next_queue, uint32_t *next_count, uint32_t *node_ctr);
void spt_launch_advance(int grid, cudaStream_t st, const DevScene &sc, const RenderCfg &cfg, const WaveBuffers &wb, int bounce,
                        const uint32_t *queue, const uint32_t *count, uint32_t *next_queue, uint32_t *next_count);
void spt_launch_addlight(int grid, cudaStream_t st, const DevScene &sc, const RenderCfg &cfg, const WaveBuffers &wb, int bounce,
                         const uint32_t *queue, const uint32_t *count);
void spt_launch_film_add(int grid, cudaStream_t st, const FilmView &film, const SptSpectralTables *tables, const float2 *img_xy,
                         const float *L, uint32_t cap, uint32_t n_samples, int spp);
void spt_launch_film_split(int grid, cudaStream_t st, const float *pix, size_t npix, float *c, float *w);
void spt_launch_gather_L(cudaStream_t st, const float *L, uint32_t cap, uint32_t n, float *out);
void spt_launch_scatter_L(cudaStream_t st, const float *in, uint32_t cap, uint32_t n, float *L);

// spt_build.cu: scene re-layout on the device
size_t spt_relayout_scratch_bytes(uint32_t n_nodes);
cudaError_t spt_launch_relayout(cudaStream_t st, void *nodes, uint32_t n_nodes, const uint8_t *prim_kind, const uint32_t *prim_data,
                                uint32_t n_prims, const int32_t *tri_vidx, const float *P, float4 *pn, float4 *tv, void *scratch,
                                uint32_t status_host[2]);

/* spt_oracle.h — TEST INFRASTRUCTURE. Plain-C CPU restatement of the reference hot path on the
 * flat scene of include/spt.h. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this; the product (libspt.so) never does.
 * Parity status: PINNED against the reference itself — tests/test_oracle_vs_reference.py checks
 * every function here against vectors produced by the unmodified reference
 * (oracle/oracle_dump.cpp -> tests/golden/tiny.golden, oracle/_ref/golden/*.golden). */
#ifndef SPT_ORACLE_H
#define SPT_ORACLE_H
#include "spt.h"
#ifdef __cplusplus
extern "C" {
#endif
int orc_nbands(void);
void orc_camera_rays(const SptCameraDesc *cam, const float *samples, uint64_t n, float *out_rays);
void orc_trace_closest(const SptSceneDesc *scene, const float *rays, uint64_t n,
                       uint32_t *out_slot, uint32_t *out_prim_id, float *out_t);
void orc_trace_any(const SptSceneDesc *scene, const float *rays, uint64_t n, uint8_t *out_hit);
/* counters: visited nodes and primitive tests summed over the n rays (for the roofline's
 * algorithmic bytes) */
void orc_trace_closest_counted(const SptSceneDesc *scene, const float *rays, uint64_t n,
                               uint64_t *nodes, uint64_t *prim_tests);
/* floats per sample vector: 37 (path) or 7 + 6 * sum of the lights' n_samples (directlighting "all") */
int orc_sample_floats(const SptSceneDesc *scene, int32_t integrator);
void orc_shade_samples(const SptSceneDesc *scene, const SptCameraDesc *cam, int32_t integrator, int32_t max_depth, int32_t spp,
                       const float *samples, const float *rng, int32_t n_rng, uint64_t n, float *out_L);
void orc_film_add_samples(const SptFilmDesc *film, const SptSpectralTables *tables,
                          const float *image_xy, const float *L, uint64_t n, float *c, float *weight);
/* The product's own sample generator restated (pbrt_v2_spectral_b200/csrc/sampler.cuh): fills the
 * 37-float vector + n_rng floats of sample `s` of sampler pixel (px,py). */
void orc_gen_sample(uint64_t seed, int32_t px, int32_t py, int32_t s, int32_t spp, float shutter_open,
                    float shutter_close, int32_t n_rng, float *sample37, float *rng);
/* Whole job on the CPU with the product's sampler: used as the deterministic checker of
 * spt_render on small images. c: [y][x][NBANDS], weight: [y][x]. */
void orc_render(const SptSceneDesc *scene, const SptCameraDesc *cam, const SptFilmDesc *film,
                const SptRenderParams *params, float *c, float *weight);
/* single functions of the restatement, for comparison with the product's device code compiled for the host (tests/host_shim/) */
void orc_measured_f(const SptSceneDesc *scene, int table, const float *wo, const float *wi, int n, float *out);
void orc_tex_evaluate(const SptSceneDesc *scene, int tex, const float *uvd, int n, float *out);
void orc_first_vertex_frame(const SptSceneDesc *scene, const SptCameraDesc *cam, int spp, const float *samples,
                            const uint32_t *slot, const float *t, int n, float *out);
void orc_light_sample(const SptSceneDesc *scene, int light, const float *p, const float *u, int n, float *out);
void orc_light_pdf(const SptSceneDesc *scene, int light, const float *p, const float *w, int n, float *out);
#ifdef __cplusplus
}
#endif
#endif

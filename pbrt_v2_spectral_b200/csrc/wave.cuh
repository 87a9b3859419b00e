// wave.cuh — per-wave path state in HBM and the device-side queues, shared by both kernel
// translation units (exact: camera + traversal; shading: everything after a hit is known).
//
// A WAVE is a batch of camera samples (whole sampler pixels x spp). Per-path state lives in HBM as
// SoA arrays indexed by the path's slot in the wave; kernels walk compacted index queues whose
// lengths stay on the device, so the host never synchronises inside a wave:
//
//   K1 gen_camera      sample -> camera ray                                   (Sampler + Camera)
//   per bounce b:
//     K2 trace<closest>  path-ray queue                                       (BVHAccel::Intersect)
//     K5 shade           hit -> shading frame, light sample, MIS sample, continuation sample;
//                        stages scalar BSDF terms + shadow / MIS rays         (PathIntegrator::Li body)
//     K3 trace<any>      shadow-ray queue                                     (BVHAccel::IntersectP)
//     K2 trace<closest>  MIS-ray queue
//     K6 accumulate      one 32-band loop: L += T*Ld, T *= f|cos|/pdf, Russian roulette,
//                        warp-aggregated compaction into the next path queue
//   K7 film_add        radiance guards + SpectralImageFilm::AddSample with a warp-per-pixel reduction
#pragma once
#include "spt_device.cuh"

struct WaveBuffers {
    uint32_t cap;                     // paths
    uint32_t jcap;                    // jobs = cap x jobs per path vertex (1; under directlighting the sum of the lights' n_samples):
                                      // g0..g2, the K5->K6 records, shadow / MIS ray results and their queues are per job
    float4 *ray_o, *ray_d;            // path rays: {o, mint}, {d, maxt}
    uint32_t *hit_slot; float *hit_t;
    float4 *g0, *g1, *g2, *g3;        // {p, eps}, {shadow d, shadow maxt}, {mis d, inf}, {path d, -}
    uint32_t *mis_slot; float *mis_t; uint32_t *sh_slot;
    // K5 -> K6 record of a path vertex, 48 bytes: the BSDF value of each direction folded to two
    // wavelength-independent coefficients {a, b}: matte/plastic f[c] = spec0[c]*a + spec1[c]*b, metal
    // f[c] = a * FrCond(b, eta[c], k[c]).
    float4 *rec0;                     // {light dir a, b, MIS dir a, b}
    float4 *rec1;                     // {continuation a, b, sL = |cos| weight / pdf of the light sample, sB likewise for the MIS sample}
    float4 *rec2;                     // {sP = |cos| / pdf of the continuation, Russian-roulette draw, RF_* flags | light << 12, material | (emitter+1) << 16}
    // extended materials only (DevScene::has_ext; one element otherwise): third coefficient of a substrate's
    // directions and the RGB an image-mapped Kd evaluated to at this vertex
    float4 *rec3;                     // {light dir c, MIS dir c, continuation c, -}: Schlick weight (1 - wi.wh)^5
    float4 *rec4;                     // {r, g, b, -}
    float *frow;                      // measured BRDFs only: [cap][3][NBP] - f(wo,wi) of the light / MIS / continuation direction, one
                                      // 128-byte row each (a table look-up has no wavelength-independent factorisation)
    float4 *laux;                     // infinite light only: RGB radiance of the sampled direction
    uint32_t *pflags;                 // scenes with specular materials: bit 0 = the ray of this path left a specular bounce
    uint32_t *root;                   // directlighting on scenes with specular materials (RenderCfg::tree): the camera sample a node of
                                      // the SpecularReflect / SpecularTransmit tree belongs to (slots >= n_samples are such nodes)
    float2 *img_xy;
    float *T[2], *L;                  // [cap][NBP]: band_off(). T[b & 1] = the throughput ARRIVING at the vertex of bounce b (k_advance of
                                      // bounce b writes the other buffer; k_addlight of bounce b still reads this one)
    uint32_t *pathQ[2], *shadowQ, *misQ;
    uint32_t *misAnyQ;                // MIS rays towards an infinite light: only hit-or-escape matters, traced as any-hit rays
    uint32_t *hitQ, *missQ;           // path rays of the bounce that found a surface / escaped (bounce 0, env light)
};

// Spectral path state layout: one 128-byte row of NB bands per path. The accumulate and film kernels
// work lane-per-band, so every access is one coalesced line whatever the order of the paths in the
// queues, and all bands of a path sit in one page (a [NB][cap] array strides by the wave capacity
// - tens of MB - between bands and thrashes the TLB).
__device__ __forceinline__ size_t band_off(uint32_t i, int c) { return (size_t)i * NBP + (uint32_t)c; }

struct RenderCfg {
    SptCameraDesc cam;
    int spp, spp_shift, max_depth;       // spp = 1 << spp_shift
    int x0, x1, y0, y1;               // sample extent (x1,y1 exclusive, already border-trimmed)
    int tile, tile_shift, tilesX, tilesY, rank, nranks;
    uint32_t seed;
    uint64_t pixel_base;              // first rank-local pixel of this wave
    uint32_t n_samples;               // samples in this wave
    int integrator;                   // SPT_INTEGRATOR_*
    int sub;                          // jobs per path vertex: 1, or under directlighting the sum of the lights' n_samples -
                                      // job r = vertex * sub + j is light sample j of UniformSampleAllLights
    float diff_scale;                 // 1/sqrt(samplesPerPixel): RayDifferential::ScaleDifferentials (samplerrenderer.cpp:91)
    int tree;                         // directlighting with specular materials: every hit may spawn a reflected and a transmitted node
                                      // (integrator.cpp:169-250) down to max_depth levels; nodes live in slots n_samples .. cap-1
};

// flags in rec2.z (low 12 bits)
enum { RF_L = 1,          // the light sample has a BSDF value: a shadow ray decides whether it counts
       RF_P_SPEC = 2,     // the continuation was sampled from a specular BxDF (path.cpp:86)
       RF_B = 4,          // the BSDF sample of the MIS estimate was traced
       RF_P = 8,          // there is a continuation direction
       RF_MEASURED = 128, // f of each direction is a full spectrum in frow (measured BRDF)
       RF_TEXKD = 256,    // spec0 of the material is replaced by FromRGB(rec4) (image-mapped Kd)
       RF_SUBSTRATE = 512,// FresnelBlend: f[c] = Kd[c](1-Ks[c]) a + (Ks[c] + (1-Ks[c]) c) b, c in rec3
       RF_ON = 1024,      // matte with Oren-Nayar
       RF_METAL = 2048 };

// j-th pixel of this rank's tile set -> raster coordinates. Tiles are square with a power-of-two side
// (tile_shift = log2), dealt round-robin to ranks; 32-bit arithmetic (spt_render bounds the counts).
__device__ __forceinline__ bool wave_pixel(const RenderCfg &cfg, uint64_t j, int *px, int *py) {
    const uint32_t ts = (uint32_t)cfg.tile_shift;
    uint32_t lt = (uint32_t)(j >> (2 * ts));
    uint32_t w = (uint32_t)j & ((1u << (2 * ts)) - 1u);
    uint32_t tile = lt * (uint32_t)cfg.nranks + (uint32_t)cfg.rank;
    if (tile >= (uint32_t)(cfg.tilesX * cfg.tilesY)) return false;
    uint32_t ty = tile / (uint32_t)cfg.tilesX, tx = tile - ty * (uint32_t)cfg.tilesX;
    *px = cfg.x0 + (int)(tx << ts) + (int)(w & ((1u << ts) - 1u));
    *py = cfg.y0 + (int)(ty << ts) + (int)(w >> ts);
    return *px < cfg.x1 && *py < cfg.y1;
}
__device__ __forceinline__ uint32_t pix_key(int px, int py) { return ((uint32_t)py << 16) ^ (uint32_t)px; }

// warp-aggregated append: one atomicAdd per warp
__device__ __forceinline__ void queue_push(uint32_t *queue, uint32_t *count, bool pred, uint32_t value) {
    unsigned mask = __ballot_sync(__activemask(), pred);
    if (!pred) return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    queue[base + __popc(mask & ((1u << lane) - 1))] = value;
}

struct FilmView {
    SptFilmDesc d;
    float *pix;            // [y][x][NB+1]
    const float *table;    // 256 filter weights in global memory
};

// spt_shade.cu — everything after a hit is known: hit compaction, K4 (miss), K5 shade, K6 accumulate,
// K7 film. Compiled -fmad=false like the traversal: radiance is compared with the reference per sample at
// 2e-4 relative, which holds only while the rounding sequence of the reference's (in places
// ill-conditioned) formulas is reproduced; hit records are rebuilt without re-deciding the hit.
#include "launch.h"
#include "shade.cuh"
#include "sampler.cuh"

// ---- hit compaction --------------------------------------------------------------------------------
// The persistent trace kernel finishes rays in no particular order, so the surviving paths are
// filtered here, in queue order (a block-wide scan per tile keeps runs of neighbouring samples together:
// the SoA spectra of the shading kernels stay coalesced and the next bounce's rays stay coherent).
// Hits go to hit_queue; escaped rays to miss_queue when one is given (camera rays under an
// environment light).
#define COMPACT_ITEMS 8            // queue entries per thread and tile: one atomic per 2048 entries per output queue
__global__ void __launch_bounds__(256) k_compact_hits(const uint32_t *__restrict__ queue, const uint32_t *__restrict__ count,
                                                      const uint32_t *__restrict__ hit_slot, uint32_t *__restrict__ hit_queue,
                                                      uint32_t *__restrict__ hit_count, uint32_t *__restrict__ miss_queue,
                                                      uint32_t *__restrict__ miss_count, float *__restrict__ black_L) {
    // Same-address atomics retire at ~1.3 G/s on this part: a per-warp append (one atomic per 32 entries)
    // made this kernel atomic-bound (0.75 ms for 31 M entries). Here a block owns a tile of 2048 consecutive
    // entries: a block-wide exclusive scan of the hit flags gives every hit its offset, ONE atomic per tile
    // reserves the output range, and the order of the queue is kept exactly within the tile.
    __shared__ uint32_t warp_hits[8], warp_miss[8], base_hit, base_miss;
    const uint32_t n = *count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = 256u * COMPACT_ITEMS;
    for (uint32_t t0 = blockIdx.x * tile; t0 < n; t0 += gridDim.x * tile) {
        // thread k owns entries t0 + k*ITEMS .. +ITEMS-1 (consecutive, so the scan order is the queue order)
        uint32_t idx[COMPACT_ITEMS];
        uint32_t hitBits = 0, valid = 0;
        const uint32_t q0 = t0 + threadIdx.x * COMPACT_ITEMS;
#pragma unroll
        for (int k = 0; k < COMPACT_ITEMS; ++k) {
            uint32_t q = q0 + k;
            idx[k] = q < n ? (queue ? queue[q] : q) : 0u;
            if (q < n) valid |= 1u << k;
        }
#pragma unroll
        for (int k = 0; k < COMPACT_ITEMS; ++k)
            if (((valid >> k) & 1u) && hit_slot[idx[k]] != SPT_MISS) hitBits |= 1u << k;
        const uint32_t missBits = valid & ~hitBits;
        uint32_t nh = __popc(hitBits), nm = __popc(missBits);
        // exclusive scan over the block: within the warp by shuffles, across warps through shared memory
        uint32_t ph = nh, pm = nm;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t a = __shfl_up_sync(0xffffffffu, ph, o), b = __shfl_up_sync(0xffffffffu, pm, o);
            if (lane >= o) { ph += a; pm += b; }
        }
        if (lane == 31) { warp_hits[warp] = ph; warp_miss[warp] = pm; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t th = 0, tm = 0;
            for (int w = 0; w < 8; ++w) { uint32_t a = warp_hits[w], b = warp_miss[w]; warp_hits[w] = th; warp_miss[w] = tm; th += a; tm += b; }
            base_hit = th ? atomicAdd(hit_count, th) : 0u;
            base_miss = (miss_queue && tm) ? atomicAdd(miss_count, tm) : 0u;
        }
        __syncthreads();
        uint32_t oh = base_hit + warp_hits[warp] + ph - nh, om = base_miss + warp_miss[warp] + pm - nm;
#pragma unroll
        for (int k = 0; k < COMPACT_ITEMS; ++k) {
            if ((hitBits >> k) & 1u) hit_queue[oh++] = idx[k];
            else if ((missBits >> k) & 1u) {
                if (miss_queue) miss_queue[om++] = idx[k];
                // camera rays that escape with no environment light: the sample's radiance row is black
                // (rows of hit samples are first written by K6, which starts them from the emitted radiance)
                if (black_L) {
                    float4 *row = (float4 *)(black_L + band_off(idx[k], 0));
#pragma unroll
                    for (int c = 0; c < NBP / 4; ++c) row[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        __syncthreads();
    }
}

// ---- K4 (miss) -----------------------------------------------------------------------------------
// Camera rays that escape: SamplerRenderer::Li's miss branch (samplerrenderer.cpp:239-243), the sum of
// Light::Le over all lights - only the infinite light is non-zero. Later bounces: the same sum times the
// throughput, but only for a ray that left a specular bounce (path.cpp:106-108).
// tree (directlighting on scenes with specular materials): an escaped node of any level adds T * Le to its root's row
// (samplerrenderer.cpp:239-243 runs for every ray of the recursion); siblings may meet there, hence the atomics.
__global__ void __launch_bounds__(128) k_miss_env(DevScene sc, WaveBuffers wb, int bounce, int tree, const uint32_t *queue, const uint32_t *count) {
    uint32_t n = *count;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        uint32_t i = queue[q];
        if (bounce > 0 && !tree && !(wb.pflags[i] & 1u)) continue;
        float4 d4 = wb.ray_d[i];
        float Le[NB];
        for (int c = 0; c < NB; ++c) Le[c] = 0.f;
        for (uint32_t l = 0; l < sc.n_lights; ++l)
            if (sc.lights[l].type == SPT_LIGHT_INFINITE) {
                float rgb[3];
                infinite_le_rgb(sc, sc.lights[l], V(d4.x, d4.y, d4.z), rgb);
                IllumCoefs k = illum_coefs(rgb);
                for (int c = 0; c < NB; ++c) Le[c] += illum_band(*sc.tables, k, c);
            }
        if (bounce == 0) { for (int c = 0; c < NB; ++c) wb.L[band_off(i, c)] = Le[c]; }
        else if (tree) { const uint32_t r = wb.root[i]; for (int c = 0; c < NB; ++c) atomicAdd(&wb.L[band_off(r, c)], wb.T[0][band_off(i, c)] * Le[c]); }
        else for (int c = 0; c < NB; ++c) wb.L[band_off(i, c)] += wb.T[bounce & 1][band_off(i, c)] * Le[c];
    }
}

// ---- K5 ----------------------------------------------------------------------------------------
// BSDF::f of one direction folded to two wavelength-independent coefficients (see WaveBuffers::rec0):
//   matte:   {1/pi [* Oren-Nayar factor], 0}          plastic: {1/pi, D*G*F / (4 cosI cosO)}
//   metal:   {D*G / (4 cosI cosO), |cos theta_h|}     (the conductor Fresnel term is per band)
__device__ __forceinline__ float2 dir_coef(int mtype, bool on, const DirTerms &t) {
    float2 d = make_float2(0.f, mtype == SPT_MAT_METAL ? 1.f : 0.f);
    if (!t.reflect) return d;
    if (mtype == SPT_MAT_MATTE) d.x = on ? INV_PI_F * t.a0 : INV_PI_F;
    else if (mtype == SPT_MAT_PLASTIC) { d.x = INV_PI_F; if (t.mf) d.y = t.a0 * t.a1 * t.a2 / t.a3; }
    else if (mtype == SPT_MAT_SUBSTRATE) { if (t.mf) { d.x = t.a0; d.y = t.a1; } }    // third coefficient (t.a2) goes to rec3
    else if (mtype == SPT_MAT_MEASURED) { if (t.mf) d.x = 1.f; }                      // the value itself is a row of frow
    else if (t.mf) { d.x = t.a0 * t.a1 / t.a3; d.y = t.a2; }
    return d;
}
// SPEC: the scene has mirror / glass materials (a second instantiation keeps their code, and the
// specularBounce flag traffic, out of the kernel every other scene runs).
// EXT: the scene has a substrate, an image-mapped Kd or a bump map (DevScene::has_ext): ray differentials at the
// first vertex, filtered texture lookups, Material::Bump, FresnelBlend - again kept out of the common kernel.
// DIRECT: directlighting, strategy "all" - the per-direction part runs once per job of the vertex (a compile-time switch:
// with a run-time job count the path integrator's kernels kept the vertex set-up live across a loop of one and lost 10 %).
// MEAS: the scene has a measured BRDF (DevScene::has_measured) - its kd-tree look-up (512 bytes of per-thread stacks and
// sums) stays out of the kernels of the textured / substrate scenes, which lost 4-6 % to it.
// The kernel is 13 k (plain) to 20 k (EXT) instructions, 205-320 KB, and a vertex runs through ~60 KB of them; the SM's
// instruction cache holds 32 KB. Warps that drift apart each stream the code from L2 on their own ("no instruction" was
// 1.7-3.4 of ~8 stall cycles per issue, 5.9 in the EXT kernels). One barrier per vertex at the top of the loop keeps the warps
// of a block within a few hundred instructions of each other, so a fetched line serves all of them, and a bigger block keeps
// more warps together (128 registers: 512 threads = one block per SM). Measured (profiles/r02_shade_sync.log, ms per frame):
//   config 1        128 threads, no barrier 41.7 | 256 + barrier 41.1 | 512 + barrier 40.9, but slower on an 8-GPU run's 1/8 frames
//   config 3 (EXT)  155.9 | 129.6 | 124.8        bunny.pbrt as shipped (EXT, MEAS)  292.0 | 239.0 | 235.6
// Barriers between the stages of a vertex as well: no further gain. A/B builds: profiles/tools/build_variants.py.
#ifndef SPT_SHADE_THREADS_EXT
#define SPT_SHADE_THREADS_EXT 512
#endif
#ifndef SPT_SHADE_THREADS
#define SPT_SHADE_THREADS 256
#endif
#ifndef SPT_SHADE_SYNC
#define SPT_SHADE_SYNC 1
#endif
#define SHADE_THREADS(EXT) ((EXT) ? SPT_SHADE_THREADS_EXT : SPT_SHADE_THREADS)
#ifndef SPT_ADVANCE_BRANCH
#define SPT_ADVANCE_BRANCH false
#endif
#ifndef SPT_SHADE_EARLY_DIMS
#define SPT_SHADE_EARLY_DIMS 1
#endif
template <bool SPEC, bool EXT, bool DIRECT = false, bool MEAS = false>
__global__ void __launch_bounds__(SHADE_THREADS(EXT), 512 / SHADE_THREADS(EXT)) k_shade(DevScene sc, RenderCfg cfg, SampleSource src, WaveBuffers wb, int bounce,
                                               const uint32_t *queue, const uint32_t *count,
                                               uint32_t *shadow_count, uint32_t *mis_count, uint32_t *elided_count, uint32_t *mis_any_count,
                                               uint32_t *next_queue, uint32_t *next_count, uint32_t *node_ctr) {
    uint32_t n = *count;
    // directlighting (EXT kernels only): a vertex carries `sub` JOBS, one per (light, light sample) of UniformSampleAllLights
    // (integrator.cpp:39-71); the vertex is set up once and the per-direction part below runs per job, its records and
    // shadow / MIS rays indexed by r = vertex * sub + job. The path integrator has one job per vertex, r = vertex.
    const bool direct = DIRECT;
    const uint32_t nj = DIRECT ? (uint32_t)cfg.sub : 1u;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < ((n + (SPT_SHADE_SYNC ? SHADE_THREADS(EXT) - 1u : 31u)) & ~(SPT_SHADE_SYNC ? SHADE_THREADS(EXT) - 1u : 31u)); q += gridDim.x * blockDim.x) {
        if (SPT_SHADE_SYNC) __syncthreads();
        bool active = q < n;
        uint32_t i = active ? queue[q] : 0;
        uint32_t slot = 0, s_idx = 0, pk = 0, i0 = 0;
        Bsdf bsdf;
        v3 p = V(0, 0, 0), n_s = p, woW = p, wo = p;
        float eps = 0.f;
        int emitter = -1;
#if SPT_SHADE_EARLY_DIMS
        // the path integrator's sample values depend on the sample's number only: hashed while the hit's loads are in flight
        float uE[10], rrE = 0.f;
        for (int k = 0; k < 10; ++k) uE[k] = 0.f;
        const bool earlyDims = !DIRECT && !(EXT && cfg.integrator == SPT_INTEGRATOR_DIRECT_ONE && src.smp);
        if (active && earlyDims) {
            s_idx = src.smp ? 0u : (i & ((uint32_t)cfg.spp - 1u));
            if (!src.smp) {
                int px, py;
                wave_pixel(cfg, cfg.pixel_base + (i >> cfg.spp_shift), &px, &py);
                pk = pixel_key(src.seed, pix_key(px, py));
            }
            bounce_dims(src, i, pk, s_idx, bounce, sc.n_lights > 0, uE, &rrE);
        }
#endif
        if (active) {
            slot = wb.hit_slot[i];
            float4 o4 = wb.ray_o[i], d4 = wb.ray_d[i];
            Ray ray;
            ray.o = V(o4.x, o4.y, o4.z); ray.d = V(d4.x, d4.y, d4.z); ray.mint = o4.w; ray.maxt = SPT_INF;
            Hit hit;
            shape_record(sc, sc.prim_kind[slot], sc.prim_flags[slot], sc.prim_data[slot], ray, wb.hit_t[i], &hit);
            // emitted light at the first vertex (path.cpp:55-56, directlighting.cpp:80; Intersection::Le, intersection.cpp:53-56):
            // K6 starts L from the emitter's spectrum instead of from black
            if (bounce == 0 || (DIRECT && cfg.tree) || (SPEC && !DIRECT && (wb.pflags[i] & 1u))) {     // directlighting.cpp:80: at every level
                int li = sc.prim_light[slot];
                if (li >= 0 && dot(hit.nn, vneg(ray.d)) > 0.f) emitter = li;
            }
            if (EXT) {
                RayDiff rdiff;
                const bool camRay = bounce == 0;                 // later rays carry no differentials (path.cpp:93)
                if (camRay) { float2 xy = wb.img_xy[i]; camera_ray_diff(cfg.cam, xy.x, xy.y, cfg.diff_scale, ray, &rdiff); }
                make_bsdf<true>(sc, slot, hit, camRay ? &rdiff : nullptr, &bsdf);
            } else make_bsdf<false>(sc, slot, hit, nullptr, &bsdf);
            p = hit.p; n_s = bsdf.nn; woW = vneg(ray.d);
            wo = w2l(bsdf, woW);
            eps = hit.rayEpsilon;
            // the camera sample whose Sample arrays this vertex reads: itself, or the root of its specular tree
            if (DIRECT && cfg.tree && bounce > 0) i0 = wb.root[i];
            else i0 = i;
            s_idx = src.smp ? 0u : (i0 & ((uint32_t)cfg.spp - 1u));
            if (!src.smp) {
                int px, py;
                wave_pixel(cfg, cfg.pixel_base + (i0 >> cfg.spp_shift), &px, &py);
                pk = pixel_key(src.seed, pix_key(px, py));
            }
            if (SPEC && DIRECT && cfg.tree && bounce < cfg.max_depth && bsdf_is_specular(bsdf)) {
                // SpecularReflect / SpecularTransmit (integrator.cpp:169-250): a child node per specular component, its ray and
                // {parent, material row | component << 16, f |cos| / pdf without the spectrum} (k_spawn_T multiplies the rows)
#pragma unroll 1
                for (int comp = 0; comp < 2; ++comp) {
                    if (!((bsdf.compMask >> comp) & 1)) continue;
                    Bsdf one = bsdf;
                    one.compMask = 1 << comp;
                    v3 wl; float cR, cT, pdfS;
                    if (!specular_sample(one, wo, 0.f, &wl, &cR, &cT, &pdfS)) continue;
                    const v3 wiW = l2w(bsdf, wl);
                    const float ad = absdot(wiW, n_s);
                    const float sc_ = (comp == 0 ? cR : cT) * ad / pdfS;
                    if (!(pdfS > 0.f) || sc_ == 0.f || ad == 0.f) continue;
                    const uint32_t child = atomicAdd(node_ctr, 1u);
                    if (child >= wb.cap) { node_ctr[1] = 1u; continue; }       // node pool exhausted: reported by spt_render
                    wb.ray_o[child] = make_float4(p.x, p.y, p.z, eps);
                    wb.ray_d[child] = make_float4(wiW.x, wiW.y, wiW.z, SPT_INF);
                    wb.g3[child] = make_float4(__uint_as_float(i), __uint_as_float((uint32_t)sc.prim_material[slot] | (uint32_t)comp << 16), sc_, 0.f);
                    next_queue[atomicAdd(next_count, 1u)] = child;
                }
            }
        }
        for (uint32_t dj = 0; dj < nj; ++dj) {
        const uint32_t r = direct ? i * nj + dj : i;
        bool pushShadow = false, pushMis = false, pushMisAny = false, elided = false;
        if (active) {
            uint32_t flags = 0;
            {
                float u[10], rr;
                int directLight = 0;
                if (direct) {
                    int prefix = 0, jj = (int)dj;
                    for (; directLight + 1 < (int)sc.n_lights && jj >= sc.lights[directLight].n_samples; ++directLight) {
                        jj -= sc.lights[directLight].n_samples; prefix += sc.lights[directLight].n_samples;
                    }
                    for (int k = 0; k < 10; ++k) u[k] = 0.f;
                    rr = 0.f;
                    direct_dims(src, i0, pk, s_idx, directLight, prefix, sc.lights[directLight].n_samples, jj, cfg.sub, u);
                } else if (EXT && cfg.integrator == SPT_INTEGRATOR_DIRECT_ONE && src.smp) {
                    // strategy "one" (directlighting.cpp:61-68): {light component, light number, bsdf component}, 2 volume floats,
                    // {light position, bsdf direction}; generated samples take the path sampler's first-bounce dimensions
                    const float *q = src.smp + (size_t)src.stride * i;
                    for (int k = 0; k < 10; ++k) u[k] = 0.f;
                    rr = 0.f;
                    u[0] = q[6]; u[1] = q[10]; u[2] = q[11]; u[3] = q[5]; u[4] = q[12]; u[5] = q[13]; u[6] = q[7];
                } else {
#if SPT_SHADE_EARLY_DIMS
                    for (int k = 0; k < 10; ++k) u[k] = uE[k];
                    rr = rrE;
#else
                    bounce_dims(src, i, pk, s_idx, bounce, sc.n_lights > 0, u, &rr);
#endif
                }

                float4 g1 = make_float4(0, 0, 0, 0), g2 = g1, g3 = g1, laux = g1;
                const int mtype = bsdf.mtype;
                const bool on = bsdf.orenNayar;
                float2 cL = make_float2(0.f, mtype == SPT_MAT_METAL ? 1.f : 0.f), cB = cL, cP = cL;
                float sL = 0.f, sB = 0.f, sP = 0.f;
                int lightIdx = 0;
                if (bsdf.orenNayar) flags |= RF_ON;
                const bool specMat = SPEC && bsdf_is_specular(bsdf);     // mirror / glass: no non-specular component, so
                const bool haveLights = sc.n_lights > 0 && !specMat;      // EstimateDirect contributes nothing and traces nothing
                // The three directions of a path vertex: 0 the light sample (UniformSampleOneLight,
                // integrator.cpp:74-106; EstimateDirect :109-137), 1 the BSDF sample of the MIS estimate
                // (:139-163), 2 the continuation (path.cpp:75-92). Each stage below runs ONCE over the
                // directions that need it (rolled loops: one copy of the sampling, Light::Pdf and BSDF::f/Pdf
                // code in the instruction cache instead of one per direction).
                LightSampleResult lr;
                lr.delta = false; lr.black = true; lr.pdfPending = false; lr.pdf = 0.f; lr.wi = V(0, 0, 1);
                if (haveLights) {
                    int nLights = (int)sc.n_lights;
                    lightIdx = (int)floorf(u[0] * nLights);
                    if (nLights - 1 < lightIdx) lightIdx = nLights - 1;
                    if (direct) lightIdx = directLight;               // UniformSampleAllLights: every light, n_samples times each
                    light_sample(sc, lightIdx, p, u[1], u[2], u[3], &lr, true);
                }
                v3 wl1 = V(0, 0, 1), wl2 = wl1, wiW1 = wl1, wiW2 = wl1;
                bool have1 = false, have2 = false;
                float pdfOv1 = -1.f, pdfOv2 = -1.f;
                float specR = 0.f, specT = 0.f, specPdf = 0.f;
                for (int d = 1; d <= 2; ++d) {
                    if (d == 1 && (!haveLights || lr.delta)) continue;
                    if (d == 2 && (direct || (EXT && cfg.integrator == SPT_INTEGRATOR_DIRECT_ONE))) continue;   // no continuation under directlighting
                    if (specMat) {                                      // only d == 2 gets here
                        v3 wl;
                        have2 = specular_sample(bsdf, wo, u[9], &wl, &specR, &specT, &specPdf);
                        wl2 = wl; wiW2 = l2w(bsdf, wl);
                        continue;
                    }
                    float uc = d == 1 ? u[6] : u[9], ua = d == 1 ? u[4] : u[7], ub = d == 1 ? u[5] : u[8];
                    v3 wl;
                    float po;
                    bool ok = bsdf_sample_dir<EXT>(bsdf, wo, uc, ua, ub, &wl, &po);
                    v3 wW = l2w(bsdf, wl);
                    if (d == 1) { have1 = ok; wl1 = wl; wiW1 = wW; pdfOv1 = po; } else { have2 = ok; wl2 = wl; wiW2 = wW; pdfOv2 = po; }
                }
                float lightPdf1 = 0.f;
#pragma unroll 1
                for (int d = 0; d <= 1; ++d) {
                    if (!(d == 0 ? lr.pdfPending : have1)) continue;
                    float pdf = light_pdf(sc, lightIdx, p, d == 0 ? lr.wi : wiW1);
                    if (d == 0) lr.pdf = pdf; else lightPdf1 = pdf;
                }
                // a BSDF-sampled ray that cannot reach the light contributes exactly zero: not traced
                if (have1 && lightPdf1 != 0.f && sc.lights[lightIdx].type == SPT_LIGHT_AREA) {
                    Ray mr; mr.o = p; mr.d = wiW1; mr.mint = eps; mr.maxt = SPT_INF;
                    if (!light_ray_may_hit(sc, sc.lights[lightIdx], mr)) { have1 = false; elided = true; }
                }
                const bool have0 = haveLights && lr.pdf > 0.f && !lr.black;
                const v3 wl0 = w2l(bsdf, lr.wi);
                DirTerms t0, t1, t2;
                t0.a0 = t0.a1 = t0.a2 = t0.a3 = 0.f; t0.reflect = t0.mf = false; t1 = t0; t2 = t0;
                float pdf0 = 0.f, pdf1 = 0.f, pdf2 = 0.f;
#pragma unroll 1
                for (int d = 0; d < 3; ++d) {
                    if (!(d == 0 ? have0 : (d == 1 ? have1 : (have2 && !specMat)))) continue;
                    v3 wW = d == 0 ? lr.wi : (d == 1 ? wiW1 : wiW2);
                    v3 wl = d == 0 ? wl0 : (d == 1 ? wl1 : wl2);
                    DirTerms t; float pdf;
                    bsdf_terms<EXT>(bsdf, woW, wW, wo, wl, &t, &pdf);
                    if (d == 0) { t0 = t; pdf0 = pdf; } else if (d == 1) { t1 = t; pdf1 = pdf; } else { t2 = t; pdf2 = pdf; }
                }
                if (EXT) { if (pdfOv1 >= 0.f) pdf1 = pdfOv1; if (pdfOv2 >= 0.f) pdf2 = pdfOv2; }
                if (have0 && t0.reflect) {
                    if (lr.delta) sL = (absdot(lr.wi, n_s) / lr.pdf);
                    else {
                        float weight = power_heuristic(lr.pdf, pdf0);
                        sL = (absdot(lr.wi, n_s) * weight / lr.pdf);
                    }
                    if (sc.lights[lightIdx].type == SPT_LIGHT_POINT) sL = sL / lr.aux[0];     // I / d^2 (point.cpp:42-49)
                    flags |= RF_L;
                    cL = dir_coef(mtype, on, t0);
                    laux = make_float4(lr.aux[0], lr.aux[1], lr.aux[2], 0.f);
                    g1 = make_float4(lr.shadow_d.x, lr.shadow_d.y, lr.shadow_d.z, lr.shadow_maxt);
                    pushShadow = true;
                }
                if (have1 && pdf1 > 0.f && t1.reflect && lightPdf1 != 0.f) {
                    float weight = power_heuristic(pdf1, lightPdf1);
                    flags |= RF_B;
                    cB = dir_coef(mtype, on, t1);
                    sB = absdot(wiW1, n_s) * weight / pdf1;
                    g2 = make_float4(wiW1.x, wiW1.y, wiW1.z, SPT_INF);
                    if (sc.lights[lightIdx].type == SPT_LIGHT_INFINITE) pushMisAny = true; else pushMis = true;
                }
                if (specMat) {
                    if (have2) {
                        flags |= RF_P | RF_P_SPEC;
                        cP = make_float2(specR, specT);                  // f[c] = Kr[c] * specR + Kt[c] * specT
                        sP = absdot(wiW2, n_s) / specPdf;
                        g3 = make_float4(wiW2.x, wiW2.y, wiW2.z, 0.f);
                    }
                } else if (have2 && pdf2 != 0.f) {
                    flags |= RF_P;
                    cP = dir_coef(mtype, on, t2);
                    sP = absdot(wiW2, n_s) / pdf2;
                    g3 = make_float4(wiW2.x, wiW2.y, wiW2.z, 0.f);
                }
                if (mtype == SPT_MAT_METAL) flags |= RF_METAL;
                if (EXT && MEAS && mtype == SPT_MAT_MEASURED) {
                    // f of every direction that survived: a table look-up per direction, written as one 128-byte row
                    flags |= RF_MEASURED;
                    const SptBrdfTable tb = sc.brdfs[bsdf.brdf];
#pragma unroll 1
                    for (int d = 0; d < 3; ++d) {
                        const bool need = d == 0 ? (flags & RF_L) != 0 : (d == 1 ? (flags & RF_B) != 0 : ((flags & RF_P) != 0 && t2.mf));
                        if (!need) continue;
                        float fv[NBP];
                        for (int c = NB; c < NBP; ++c) fv[c] = 0.f;
                        measured_f(sc, tb, wo, d == 0 ? wl0 : (d == 1 ? wl1 : wl2), fv);
                        float4 *row = (float4 *)(wb.frow + ((size_t)r * 3 + d) * NBP);
#pragma unroll
                        for (int c = 0; c < NBP / 4; ++c) row[c] = make_float4(fv[4 * c], fv[4 * c + 1], fv[4 * c + 2], fv[4 * c + 3]);
                    }
                }
                if (EXT) {
                    if (mtype == SPT_MAT_SUBSTRATE) {
                        flags |= RF_SUBSTRATE;
                        wb.rec3[r] = make_float4((flags & RF_L) ? t0.a2 : 0.f, (flags & RF_B) ? t1.a2 : 0.f,
                                                 ((flags & RF_P) && t2.mf) ? t2.a2 : 0.f, 0.f);
                    }
                    if (bsdf.texKd) { flags |= RF_TEXKD; wb.rec4[r] = make_float4(bsdf.kd_rgb[0], bsdf.kd_rgb[1], bsdf.kd_rgb[2], 0.f); }
                }
                wb.g0[r] = make_float4(p.x, p.y, p.z, eps);
                if (pushShadow) wb.g1[r] = g1;
                if (pushMis || pushMisAny) wb.g2[r] = g2;
                if (flags & RF_P) wb.g3[i] = g3;
                wb.rec0[r] = make_float4(cL.x, cL.y, cB.x, cB.y);
                wb.rec1[r] = make_float4(cP.x, cP.y, sL, sB);
                wb.rec2[r] = make_float4(sP, rr, __uint_as_float(flags | (uint32_t)lightIdx << 12),
                                         __uint_as_float((uint32_t)sc.prim_material[slot] | (uint32_t)(dj == 0 ? emitter + 1 : 0) << 16));
                if (pushShadow && sc.lights[lightIdx].type == SPT_LIGHT_INFINITE) wb.laux[r] = laux;
            }
        }
        queue_push(wb.shadowQ, shadow_count, pushShadow, r);
        queue_push(wb.misQ, mis_count, pushMis, r);
        queue_push(wb.misAnyQ, mis_any_count, pushMisAny, r);
        unsigned em = __ballot_sync(0xffffffffu, elided);
        if (em && (threadIdx.x & 31) == 0) atomicAdd(elided_count, (uint32_t)__popc(em));
        }
    }
}

// ---- K6 ----------------------------------------------------------------------------------------
// Per-direction BSDF value rebuilt per band from wavelength-independent coefficients:
//   matte / plastic:  f[c] = spec0[c]*u0 + spec1[c]*u1     (Lambert/Oren-Nayar + Blinn microfacet x dielectric Fresnel)
//   metal:            f[c] = w * FrCond(cosH, eta[c], k[c])
// The scalar factors (D*G*F/(4 cosI cosO), |cos|*weight/pdf, ...) are folded once per vertex instead of
// once per band as the reference's Spectrum arithmetic does; this reassociation moves results by
// rounding only (tests compare radiance at 2e-4 relative).
// FrCond (reflection.cpp:63-71) with the two quotients combined into one division
__device__ __forceinline__ float fr_cond_fast(float cosi, float c2, float eta, float k) {
    float A = fmaf(eta, eta, k * k);
    float e2 = 2.f * eta * cosi;
    float Ac2 = A * c2;
    float n1 = Ac2 - e2 + 1.f, d1 = Ac2 + e2 + 1.f;
    float n2 = A - e2 + c2, d2 = A + e2 + c2;
    return 0.5f * __fdividef(fmaf(n1, d2, n2 * d1), d1 * d2);
}
// Radiance of a light direction, per band: kind 0 none, 1 the light's table spectrum, 2 RGB illuminant
// (three basis spectra, coefficients k0..k2)
struct LightBand { int kind; IllumCoefs k; };

// K6 is two kernels, because the continuation ray of a vertex does not depend on that vertex's shadow / MIS rays:
//   k_advance   straight after K5: T' = T * f|cos|/pdf of the continuation, Russian roulette, the next ray, the next
//               path queue (path.cpp:88-104). T is double-buffered by bounce parity: T[b & 1] is the throughput ARRIVING at
//               the vertex of bounce b, so the old value survives until the light terms are added.
//   k_addlight  after the trace launch that carried this bounce's shadow / MIS rays (the same launch that traced the NEXT
//               bounce's path rays): L += T * (Le + Ld * nLights) (path.cpp:55-56, integrator.cpp:122-163).
// The dependent chain of a bounce is therefore trace -> compact -> K5 -> k_advance -> trace, with ONE persistent trace
// launch per bounce (a persistent kernel ends in the drain of its longest rays: at 1/8 of a frame per GPU those drains,
// three or four per bounce before, were the scaling loss).
//
// Both kernels: a warp takes 32 vertices. Phase A, lane = vertex: everything wavelength-independent, staged in shared
// memory. Phase B, lane = (vertex of a group of four, four bands): 8 lanes x float4 cover a vertex's 128-byte row, so one
// LDG.128 / STG.128 per lane moves four rows per warp instruction (the one-float-per-lane form was issue-bound at 0.49 of
// the HBM rate); y(T) is a sum over the 8 lanes of a vertex. Phase C (k_advance), lane = vertex: next ray + queue push.
static_assert(NBP == 32, "k_advance / k_addlight / k_film_add assume 128-byte rows (lane = band, or 8 lanes x float4)");
#define ACC_WARPS 4
#ifndef ADV_CHUNK
#define ADV_CHUNK 4
#endif
struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const float *row, int bg) { float4 q = __ldg((const float4 *)row + bg); F4 r; r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w; return r; }
__device__ __forceinline__ F4 ld4g(const float *row, int bg) { float4 q = *((const float4 *)row + bg); F4 r; r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w; return r; }
__device__ __forceinline__ void st4(float *row, int bg, const F4 &r) { *((float4 *)row + bg) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]); }
__device__ __forceinline__ F4 splat4(float x) { F4 r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = x; return r; }
// Kd from an image map: FromRGB(rgb, SPECTRUM_REFLECTANCE), four bands
__device__ __forceinline__ F4 refl4(const SptSpectralTables &tb, const float4 &e4, uint32_t bits, int bg) {
    IllumCoefs kk; kk.k0 = e4.x; kk.k1 = e4.y; kk.k2 = e4.z; kk.b1 = bits & 15; kk.b2 = (bits >> 4) & 15;
    F4 r;
#pragma unroll
    for (int c = 0; c < 4; ++c) r.v[c] = refl_band(tb, kk, 4 * bg + c);
    return r;
}
__device__ __forceinline__ F4 illum4(const SptSpectralTables &tb, const float4 &l4, int bg) {
    const uint32_t kb = __float_as_uint(l4.w);
    IllumCoefs kk; kk.k0 = l4.x; kk.k1 = l4.y; kk.k2 = l4.z; kk.b1 = (kb >> 4) & 15; kk.b2 = (kb >> 8) & 15;
    F4 r;
#pragma unroll
    for (int c = 0; c < 4; ++c) r.v[c] = illum_band(tb, kk, 4 * bg + c);
    return r;
}
// f of one direction, four bands, from its two folded coefficients {a, b} (see WaveBuffers::rec0) and the material rows.
// kind: 0 matte / plastic (f = spec0 a + spec1 b), 1 metal (f = a FrCond(b, eta, k)), 2 substrate (FresnelBlend with the
// Schlick weight c3: Kd (1 - Ks) a + (Ks + (1 - Ks) c3) b)
// BRANCH: the material kind is tested once, outside the band loop. k_addlight wants that: with the test inside the loop every
// vertex pays for all the formulas (47 % of its instructions on a scene of matte and plastic only; 1.83 -> 1.27 ms at
// bounce 0). k_advance keeps the test inside (no difference measured either way once its queue claims were chunked).
template <bool EXT, bool BRANCH>
__device__ __forceinline__ F4 fold_f(int kind, const F4 &s0, const F4 &s1, float a, float b, float c3) {
    F4 r;
    if (!BRANCH) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (kind == 1) r.v[c] = a != 0.f ? a * fr_cond_fast(b, b * b, s0.v[c], s1.v[c]) : 0.f;
            else if (EXT && kind == 2) { const float oms = 1.f - s1.v[c]; r.v[c] = s0.v[c] * oms * a + (s1.v[c] + oms * c3) * b; }
            else r.v[c] = fmaf(s0.v[c], a, s1.v[c] * b);
        }
        return r;
    }
    if (kind == 1) {
        if (a != 0.f) {
            const float b2 = b * b;
#pragma unroll
            for (int c = 0; c < 4; ++c) r.v[c] = a * fr_cond_fast(b, b2, s0.v[c], s1.v[c]);
        } else r = splat4(0.f);
    } else if (EXT && kind == 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { const float oms = 1.f - s1.v[c]; r.v[c] = s0.v[c] * oms * a + (s1.v[c] + oms * c3) * b; }
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) r.v[c] = fmaf(s0.v[c], a, s1.v[c] * b);
    }
    return r;
}

// EXT (DevScene::has_ext): the substrate's third coefficients, the reflectance coefficients of an image-mapped Kd and the
// measured-BRDF rows.
template <bool EXT>
__global__ void __launch_bounds__(32 * ACC_WARPS, 8) k_advance(DevScene sc, RenderCfg cfg, WaveBuffers wb, int bounce,
                                                            const uint32_t *__restrict__ queue, const uint32_t *__restrict__ count,
                                                            uint32_t *__restrict__ next_queue, uint32_t *__restrict__ next_count) {
    constexpr int STAGE = EXT ? 3 : 2;
    __shared__ float4 stage_all[ACC_WARPS][32][STAGE];
    const uint32_t n = *count;
    const SptSpectralTables &tb = *sc.tables;
    const float *__restrict__ Tin = wb.T[bounce & 1];
    float *__restrict__ Tout = wb.T[(bounce + 1) & 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 3, bg = lane & 7;
    float4 (*stage)[STAGE] = stage_all[warp];
    const unsigned FULL = 0xffffffffu;
    const F4 cieY = ld4(tb.cie_y, bg);
    const float yint = tb.yint;
    const uint32_t nwarps = gridDim.x * ACC_WARPS;
    // A warp takes ADV_CHUNK consecutive passes of 32 vertices and claims their queue slots with ONE atomic: the counter of the
    // next path queue is a single address, and one claim per 32 vertices (0.9 M same-address atomics at bounce 0 of config 1)
    // was what the kernel's time followed, not its DRAM traffic. Survivors wait in shared memory until the chunk is claimed.
    __shared__ uint32_t pend_all[ACC_WARPS][32 * ADV_CHUNK];
    uint32_t *pend = pend_all[warp];
    for (uint32_t cbase = (blockIdx.x * ACC_WARPS + warp) * (32u * ADV_CHUNK); cbase < n; cbase += nwarps * (32u * ADV_CHUNK)) {
      uint32_t npend = 0;
      for (uint32_t base = cbase; base < min(n, cbase + 32u * ADV_CHUNK); base += 32u) {
        // ---- phase A: lane = vertex
        const uint32_t q = base + lane;
        const bool active = q < n;
        const uint32_t i = active ? queue[q] : 0;
        bool specBounce = false;
        if (active) {
            const float4 c1 = wb.rec1[i], c2 = wb.rec2[i];
            const uint32_t bits0 = __float_as_uint(c2.z), bits1 = __float_as_uint(c2.w);
            const uint32_t flags = bits0 & 0xfffu;
            const bool haveP = (flags & RF_P) != 0;
            specBounce = (flags & RF_P_SPEC) != 0;
            uint32_t kind = (flags & RF_METAL) ? 1u : 0u;
            float c3 = 0.f;
            float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t ebits = 0;
            if (EXT) {
                if (flags & RF_SUBSTRATE) { kind = 2u; c3 = wb.rec3[i].z; }
                if (flags & RF_MEASURED) kind = 3u;
                if (flags & RF_TEXKD) {
                    const float4 r4 = wb.rec4[i];
                    const float rgb[3] = { r4.x, r4.y, r4.z };
                    const IllumCoefs kk = illum_coefs(rgb);           // same basis choice for the reflectance tables
                    e4 = make_float4(kk.k0, kk.k1, kk.k2, 0.f);
                    ebits = 1u | (uint32_t)kk.b1 << 4 | (uint32_t)kk.b2 << 8;
                }
            }
            stage[lane][0] = make_float4(c1.x, c1.y, haveP ? c2.x : 0.f, c2.y);           // {a, b, sP, Russian-roulette draw}
            stage[lane][1] = make_float4(__uint_as_float(i), __uint_as_float((haveP ? 1u : 0u) | kind << 1 | ebits << 4),
                                         __uint_as_float(bits1 & 0xffffu), c3);
            if (EXT) stage[lane][2] = e4;
        }
        __syncwarp();
        // ---- phase B: lane = (vertex grp of a group of four, bands 4 bg .. 4 bg + 3)
        const uint32_t cnt = min(32u, n - base);
        uint32_t aliveMask = 0;
        for (uint32_t v0 = 0; v0 < cnt; v0 += 4u) {
            const uint32_t v = v0 + (uint32_t)grp;
            const bool valid = v < cnt;
            const uint32_t vv = valid ? v : cnt - 1u;
            const float4 cP = stage[vv][0], m4 = stage[vv][1];
            const uint32_t vi = __float_as_uint(m4.x), misc = __float_as_uint(m4.y);
            const bool haveP = valid && (misc & 1u);
            const uint32_t kind = (misc >> 1) & 7u;
            F4 Tn = splat4(0.f), fP = splat4(0.f);
            if (haveP) {
                const F4 Tv = bounce == 0 ? splat4(1.f) : ld4g(Tin + (size_t)vi * NBP, bg);
                if (EXT && kind == 3u) {
                    // measured BRDF: the value is a row K5 wrote (coefficient 1 = the direction has a value)
                    if (cP.x != 0.f) fP = ld4g(wb.frow + ((size_t)vi * 3 + 2) * NBP, bg);
                } else {
                    const SptMaterial &m = sc.materials[__float_as_uint(m4.z)];
                    F4 s0 = ld4(m.spec0, bg);
                    const F4 s1 = ld4(m.spec1, bg);
                    if (EXT && ((misc >> 4) & 1u)) s0 = refl4(tb, stage[vv][2], misc >> 8, bg);
                    fP = fold_f<EXT, SPT_ADVANCE_BRANCH>((int)kind, s0, s1, cP.x, cP.y, m4.w);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) Tn.v[c] = Tv.v[c] * (fP.v[c] * cP.z);
            }
            // path.cpp:88-104: the path ends on a black BSDF value; Russian roulette from the fourth bounce on
            const unsigned nz = __ballot_sync(FULL, fP.v[0] != 0.f || fP.v[1] != 0.f || fP.v[2] != 0.f || fP.v[3] != 0.f);
            bool alive = haveP && ((nz >> (grp * 8)) & 0xffu) != 0u;
            if (bounce > 3) {
                float yy = cieY.v[0] * Tn.v[0] + cieY.v[1] * Tn.v[1] + cieY.v[2] * Tn.v[2] + cieY.v[3] * Tn.v[3];
                yy += __shfl_xor_sync(FULL, yy, 4);
                yy += __shfl_xor_sync(FULL, yy, 2);
                yy += __shfl_xor_sync(FULL, yy, 1);
                const float continueProbability = stdminf(.5f, yy / yint);
                if (cP.w > continueProbability) alive = false;
                else {
                    const float inv = 1.f / continueProbability;
#pragma unroll
                    for (int c = 0; c < 4; ++c) Tn.v[c] *= inv;
                }
            }
            if (alive) st4(Tout + (size_t)vi * NBP, bg, Tn);
            // the four vertices of this pass: their verdicts sit at lanes 0, 8, 16, 24 of the ballot -> bits v0 .. v0 + 3
            const unsigned b4 = __ballot_sync(FULL, alive && bg == 0);
            aliveMask |= ((b4 & 1u) | ((b4 >> 7) & 2u) | ((b4 >> 14) & 4u) | ((b4 >> 21) & 8u)) << v0;
        }
        // ---- phase C: lane = vertex
        const bool alive = active && ((aliveMask >> lane) & 1u);
        const unsigned pushMask = __ballot_sync(FULL, alive);
        if (alive) {
            const float4 g3 = wb.g3[i], g0 = wb.g0[i];
            wb.ray_o[i] = g0;
            wb.ray_d[i] = make_float4(g3.x, g3.y, g3.z, SPT_INF);
            if (sc.has_specular) wb.pflags[i] = specBounce ? 1u : 0u;
            pend[npend + __popc(pushMask & ((1u << lane) - 1u))] = i;
        }
        npend += (uint32_t)__popc(pushMask);
        __syncwarp();
      }
      // ---- the chunk's survivors -> the next path queue, in vertex order
      if (npend) {
        uint32_t pushBase = 0;
        if (lane == 0) pushBase = atomicAdd(next_count, npend);
        pushBase = __shfl_sync(FULL, pushBase, 0);
        for (uint32_t k = lane; k < npend; k += 32u) next_queue[pushBase + k] = pend[k];
      }
      __syncwarp();
    }
}

template <bool EXT>
__global__ void __launch_bounds__(32 * ACC_WARPS, 8) k_addlight(DevScene sc, RenderCfg cfg, WaveBuffers wb, int bounce,
                                                             const uint32_t *__restrict__ queue, const uint32_t *__restrict__ count) {
    constexpr int STAGE = EXT ? 7 : 5;
    __shared__ float4 stage_all[ACC_WARPS][32][STAGE];
    const uint32_t n = *count;
    const SptSpectralTables &tb = *sc.tables;
    const float *__restrict__ Tin = wb.T[bounce & 1];        // the throughput that ARRIVED at this bounce's vertices
    float *__restrict__ Lg = wb.L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 3, bg = lane & 7;
    float4 (*stage)[STAGE] = stage_all[warp];
    const float nL = (float)sc.n_lights;
    const F4 light0 = sc.n_lights ? ld4(sc.lights[0].spectrum, bg) : splat4(0.f);    // most scenes: the one light's row, read once
    const uint32_t nwarps = gridDim.x * ACC_WARPS;
    for (uint32_t base = (blockIdx.x * ACC_WARPS + warp) * 32u; base < n; base += nwarps * 32u) {
        // ---- phase A: lane = vertex - which terms survived the shadow / MIS rays
        const uint32_t q = base + lane;
        const bool active = q < n;
        if (active) {
            const uint32_t i = queue[q];
            const float4 c0 = wb.rec0[i], c1 = wb.rec1[i], c2 = wb.rec2[i];
            const uint32_t bits0 = __float_as_uint(c2.z), bits1 = __float_as_uint(c2.w);
            const uint32_t flags = bits0 & 0xfffu;
            const int lightIdx = (int)(bits0 >> 12);
            const bool metal = (flags & RF_METAL) != 0;
            const float2 none = make_float2(0.f, metal ? 1.f : 0.f);
            float2 cL = none, cB = none;
            LightBand lbL, lbB; lbL.kind = 0; lbB.kind = 0;
            lbL.k.k0 = lbL.k.k1 = lbL.k.k2 = 0.f; lbL.k.b1 = lbL.k.b2 = 0; lbB.k = lbL.k;
            float sL = 0.f, sB = 0.f;
            // --- light-sample term (integrator.cpp:122-137): visible iff the shadow ray found nothing
            if ((flags & RF_L) && wb.sh_slot[i] == SPT_MISS) {
                cL = make_float2(c0.x, c0.y);
                sL = c1.z;
                if (sc.lights[lightIdx].type == SPT_LIGHT_INFINITE) {
                    float4 r5 = wb.laux[i];
                    float rgb[3] = { r5.x, r5.y, r5.z };
                    lbL.kind = 2; lbL.k = illum_coefs(rgb);
                } else lbL.kind = 1;
            }
            // --- BSDF-sample term (integrator.cpp:139-163): radiance from what the MIS ray found
            if (flags & RF_B) {
                const float4 g0 = wb.g0[i], g2 = wb.g2[i];
                uint32_t ms = wb.mis_slot[i];
                const SptLight &l = sc.lights[lightIdx];
                v3 wi = V(g2.x, g2.y, g2.z);
                if (ms != SPT_MISS) {
                    if (sc.prim_light[ms] == lightIdx) {
                        Ray ray; ray.o = V(g0.x, g0.y, g0.z); ray.d = wi; ray.mint = g0.w; ray.maxt = SPT_INF;
                        Hit h;
                        shape_record(sc, sc.prim_kind[ms], sc.prim_flags[ms], sc.prim_data[ms], ray, wb.mis_t[i], &h);
                        if (dot(h.nn, vneg(wi)) > 0.f) lbB.kind = 1;
                    }
                } else if (l.type == SPT_LIGHT_INFINITE) {
                    float rgb[3];
                    infinite_le_rgb(sc, l, wi, rgb);
                    lbB.kind = 2; lbB.k = illum_coefs(rgb);
                }
                if (lbB.kind) { cB = make_float2(c0.z, c0.w); sB = c1.w; }
            }
            const uint32_t emit = bits1 >> 16;
            // rows are first written at bounce 0 (every vertex); afterwards only vertices that add something touch theirs
            const bool need = bounce == 0 || emit != 0u || lbL.kind != 0 || lbB.kind != 0;
            uint32_t kind = metal ? 1u : 0u;
            float4 e3 = make_float4(0.f, 0.f, 0.f, 0.f), e4 = e3;
            uint32_t ebits = 0;
            if (EXT) {
                if (flags & RF_SUBSTRATE) { kind = 2u; e3 = wb.rec3[i]; }
                if (flags & RF_MEASURED) kind = 3u;
                if (flags & RF_TEXKD) {
                    const float4 r4 = wb.rec4[i];
                    const float rgb[3] = { r4.x, r4.y, r4.z };
                    const IllumCoefs kk = illum_coefs(rgb);
                    e4 = make_float4(kk.k0, kk.k1, kk.k2, 0.f);
                    ebits = 1u | (uint32_t)kk.b1 << 4 | (uint32_t)kk.b2 << 8;
                }
            }
            // UniformSampleOneLight scales by the light count (integrator.cpp:105)
            stage[lane][0] = make_float4(cL.x, cL.y, cB.x, cB.y);
            stage[lane][1] = make_float4(sL * nL, sB * nL, __uint_as_float(i),
                                         __uint_as_float((need ? 1u : 0u) | kind << 1 | (lbL.kind == 2 ? 16u : 0u) | (lbB.kind == 2 ? 32u : 0u) |
                                                         (uint32_t)lightIdx << 8));
            stage[lane][2] = make_float4(lbL.k.k0, lbL.k.k1, lbL.k.k2, __uint_as_float((uint32_t)lbL.kind | (uint32_t)lbL.k.b1 << 4 | (uint32_t)lbL.k.b2 << 8));
            stage[lane][3] = make_float4(lbB.k.k0, lbB.k.k1, lbB.k.k2, __uint_as_float((uint32_t)lbB.kind | (uint32_t)lbB.k.b1 << 4 | (uint32_t)lbB.k.b2 << 8));
            stage[lane][4] = make_float4(__uint_as_float(bits1 & 0xffffu), __uint_as_float(emit), __uint_as_float(ebits), 0.f);
            if (EXT) { stage[lane][5] = e3; stage[lane][6] = e4; }
        }
        __syncwarp();
        // ---- phase B: lane = (vertex grp of a group of four, bands 4 bg .. 4 bg + 3)
        const uint32_t cnt = min(32u, n - base);
        for (uint32_t v0 = 0; v0 < cnt; v0 += 4u) {
            const uint32_t v = v0 + (uint32_t)grp;
            if (v >= cnt) continue;
            const float4 cLB = stage[v][0], m4 = stage[v][1];
            const uint32_t misc = __float_as_uint(m4.w);
            if (!(misc & 1u)) continue;
            const uint32_t vi = __float_as_uint(m4.z);
            const uint32_t kind = (misc >> 1) & 7u, lightIdx = misc >> 8;
            const float4 x4 = stage[v][4];
            const uint32_t emit = __float_as_uint(x4.y);
            float *Lrow = Lg + (size_t)vi * NBP;
            F4 Tv, Lv;
            if (bounce == 0) {
                // the path starts here: T = 1, L = what the first vertex emits towards the camera (path.cpp:55-56)
                Tv = splat4(1.f);
                Lv = emit == 0 ? splat4(0.f) : (emit == 1 ? light0 : ld4(sc.lights[emit - 1].spectrum, bg));
            } else {
                Tv = ld4g(Tin + (size_t)vi * NBP, bg);
                Lv = ld4g(Lrow, bg);
                // a vertex reached through a specular bounce adds what it emits (path.cpp:55-56)
                if (emit) {
                    const F4 Le = emit == 1 ? light0 : ld4(sc.lights[emit - 1].spectrum, bg);
#pragma unroll
                    for (int c = 0; c < 4; ++c) Lv.v[c] = Lv.v[c] + Tv.v[c] * Le.v[c];
                }
            }
            F4 fL = splat4(0.f), fB = splat4(0.f);
            if (EXT && kind == 3u) {
                // measured BRDF: the values are rows K5 wrote (coefficient 1 = the direction has a value)
                if (cLB.x != 0.f) fL = ld4g(wb.frow + ((size_t)vi * 3 + 0) * NBP, bg);
                if (cLB.z != 0.f) fB = ld4g(wb.frow + ((size_t)vi * 3 + 1) * NBP, bg);
            } else {
                const SptMaterial &m = sc.materials[__float_as_uint(x4.x)];
                F4 s0 = ld4(m.spec0, bg);
                const F4 s1 = ld4(m.spec1, bg);
                float c3L = 0.f, c3B = 0.f;
                if (EXT) {
                    const uint32_t ebits = __float_as_uint(x4.z);
                    if (ebits & 1u) s0 = refl4(tb, stage[v][6], ebits >> 4, bg);
                    const float4 e3 = stage[v][5];
                    c3L = e3.x; c3B = e3.y;
                }
                fL = fold_f<EXT, true>((int)kind, s0, s1, cLB.x, cLB.y, c3L);
                fB = fold_f<EXT, true>((int)kind, s0, s1, cLB.z, cLB.w, c3B);
            }
            // radiance arriving along the light / MIS direction: the light's table row, or (infinite light) an RGB
            // illuminant rebuilt from the staged coefficients
            F4 LcL, LcB;
            if (misc & 48u) {
                LcL = (misc & 16u) ? illum4(tb, stage[v][2], bg) : splat4(0.f);
                LcB = (misc & 32u) ? illum4(tb, stage[v][3], bg) : splat4(0.f);
            } else {
                // a direction without a light term has zero coefficients / scale, so the row can be applied unconditionally
                LcL = LcB = lightIdx == 0 ? light0 : ld4(sc.lights[lightIdx].spectrum, bg);
            }
            // L += T * Ld * nLights (integrator.cpp:122-163, path.cpp:69-72)
            F4 Lo;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float Ld = fL.v[c] * LcL.v[c] * m4.x + fB.v[c] * LcB.v[c] * m4.y;
                Lo.v[c] = fmaf(Tv.v[c], Ld, Lv.v[c]);
            }
            st4(Lrow, bg, Lo);
        }
        __syncwarp();
    }
}

// ---- K6, directlighting ---------------------------------------------------------------------------
// DirectLightingIntegrator::Li after the hit (directlighting.cpp:80-96): L = Le + sum over lights of (sum of the light's
// n_samples EstimateDirect values) / n_samples (integrator.cpp:39-71). A warp takes floor(32 / sub) camera hits = up to 32
// jobs per batch. Phase A, lane = job: which terms survived the shadow / MIS rays, folded to three staged float4 per job
// (coefficients of the light-sample and BSDF-sample directions, their scales - already divided by n_samples - and what
// radiance arrives). Phase B, lane = band: per hit the material row is read once, its jobs' staged values are shared-memory
// broadcasts, and the radiance row is written once - no throughput, no continuation.
#define ACCD_WARPS 4
// level > 0 (RenderCfg::tree): the hit is a node of a camera sample's SpecularReflect / SpecularTransmit tree - its radiance times
// the node's throughput row is ADDED to the root's row (integrator.cpp:196,244: L = f * Li * AbsDot(wi, n) / pdf, level by level).
__global__ void __launch_bounds__(32 * ACCD_WARPS, 8) k_accumulate_direct(DevScene sc, RenderCfg cfg, WaveBuffers wb, int level,
                                                                       const uint32_t *__restrict__ queue, const uint32_t *__restrict__ count) {
    __shared__ float4 stage_all[ACCD_WARPS][32][4];
    const uint32_t n = *count;
    const SptSpectralTables &tb = *sc.tables;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 (*stage)[4] = stage_all[warp];
    const uint32_t sub = (uint32_t)cfg.sub;
    const uint32_t H = sub <= 32u ? 32u / sub : 1u;                  // hits per batch; sub > 32: one hit, jobs in rounds of 32
    const uint32_t gw = blockIdx.x * ACCD_WARPS + warp, nwarps = gridDim.x * ACCD_WARPS;
    for (uint32_t h0 = gw * H; h0 < n; h0 += nwarps * H) {
        const uint32_t nh = min(H, n - h0);
        float Lacc = 0.f;                                            // sub > 32 only: the hit's sum across rounds
        for (uint32_t j0 = 0; j0 < sub; j0 += 32u) {                 // one round unless sub > 32
            // ---- phase A: lane = job
            const uint32_t hl = sub <= 32u ? (uint32_t)lane / sub : 0u;                  // hit of this lane within the batch
            const uint32_t j = sub <= 32u ? (uint32_t)lane - hl * sub : j0 + (uint32_t)lane;
            const bool jobLane = hl < nh && j < sub;
            if (jobLane) {
                const uint32_t i = queue[h0 + hl];
                const uint32_t r = i * sub + j;
                const float4 c0 = wb.rec0[r], c1 = wb.rec1[r], c2 = wb.rec2[r];
                const uint32_t bits0 = __float_as_uint(c2.z);
                const uint32_t flags = bits0 & 0xfffu;
                const int lightIdx = (int)(bits0 >> 12);
                const SptLight &l = sc.lights[lightIdx];
                const float inv = 1.f / (float)l.n_samples;
                const float2 none = make_float2(0.f, (flags & RF_METAL) ? 1.f : 0.f);
                float2 cL = none, cB = none;
                float sL = 0.f, sB = 0.f;
                uint32_t kindL = 0, kindB = 0;                        // 0 none, 1 the light's table spectrum, 2 RGB illuminant
                float4 rgbL = make_float4(0, 0, 0, 0), rgbB = rgbL;
                if ((flags & RF_L) && wb.sh_slot[r] == SPT_MISS) {    // integrator.cpp:122-137
                    cL = make_float2(c0.x, c0.y); sL = c1.z * inv;
                    if (l.type == SPT_LIGHT_INFINITE) { rgbL = wb.laux[r]; kindL = 2; } else kindL = 1;
                }
                if (flags & RF_B) {                                   // integrator.cpp:139-163
                    const float4 g0 = wb.g0[r], g2 = wb.g2[r];
                    const uint32_t ms = wb.mis_slot[r];
                    const v3 wi = V(g2.x, g2.y, g2.z);
                    if (ms != SPT_MISS) {
                        if (sc.prim_light[ms] == lightIdx) {
                            Ray ray; ray.o = V(g0.x, g0.y, g0.z); ray.d = wi; ray.mint = g0.w; ray.maxt = SPT_INF;
                            Hit hh;
                            shape_record(sc, sc.prim_kind[ms], sc.prim_flags[ms], sc.prim_data[ms], ray, wb.mis_t[r], &hh);
                            if (dot(hh.nn, vneg(wi)) > 0.f) kindB = 1;
                        }
                    } else if (l.type == SPT_LIGHT_INFINITE) {
                        float rgb[3];
                        infinite_le_rgb(sc, l, wi, rgb);
                        rgbB = make_float4(rgb[0], rgb[1], rgb[2], 0.f); kindB = 2;
                    }
                    if (kindB) { cB = make_float2(c0.z, c0.w); sB = c1.w * inv; }
                }
                stage[lane][0] = make_float4(cL.x, cL.y, cB.x, cB.y);
                stage[lane][1] = make_float4(sL, sB, __uint_as_float(flags | kindL << 12 | kindB << 14 | (uint32_t)lightIdx << 16), 0.f);
                stage[lane][2] = rgbL;
                stage[lane][3] = rgbB;
            }
            __syncwarp();
            // ---- phase B: lane = band
            for (uint32_t hh = 0; hh < nh; ++hh) {
                const uint32_t i = queue[h0 + hh];
                const uint32_t r0 = i * sub;
                const uint32_t bits1 = __float_as_uint(wb.rec2[r0].w);
                const SptMaterial &m = sc.materials[bits1 & 0xffffu];
                float s0 = __ldg(m.spec0 + lane);
                const float s1 = __ldg(m.spec1 + lane);
                float L = 0.f;
                if (j0 == 0) { const uint32_t emit = bits1 >> 16; L = emit ? __ldg(sc.lights[emit - 1].spectrum + lane) : 0.f; }
                const uint32_t jn = sub <= 32u ? sub : min(32u, sub - j0);
                if (__float_as_uint(stage[sub <= 32u ? hh * sub : 0u][1].z) & RF_TEXKD) {      // the image-mapped Kd is a property of the vertex
                    const float4 r4 = wb.rec4[r0];
                    const float rgb[3] = { r4.x, r4.y, r4.z };
                    s0 = refl_band(tb, illum_coefs(rgb), lane);
                }
                for (uint32_t jj = 0; jj < jn; ++jj) {
                    const uint32_t sl = sub <= 32u ? hh * sub + jj : jj;
                    const float4 cLB = stage[sl][0], sc4 = stage[sl][1];
                    const uint32_t bits = __float_as_uint(sc4.z);
                    const uint32_t kindL = (bits >> 12) & 3u, kindB = (bits >> 14) & 3u;
                    if (!(kindL | kindB)) continue;
                    const uint32_t r = r0 + (sub <= 32u ? jj : j0 + jj);
                    float fL, fB;
                    if (bits & RF_MEASURED) {
                        const float *fr = wb.frow + (size_t)r * 3 * NBP + lane;
                        fL = cLB.x != 0.f ? fr[0] : 0.f;
                        fB = cLB.z != 0.f ? fr[NBP] : 0.f;
                    } else if (bits & RF_SUBSTRATE) {
                        const float4 e3 = wb.rec3[r];
                        const float oms = 1.f - s1, dR = s0 * oms;
                        fL = dR * cLB.x + (s1 + oms * e3.x) * cLB.y;
                        fB = dR * cLB.z + (s1 + oms * e3.y) * cLB.w;
                    } else if (bits & RF_METAL) {
                        fL = cLB.x != 0.f ? cLB.x * fr_cond_fast(cLB.y, cLB.y * cLB.y, s0, s1) : 0.f;
                        fB = cLB.z != 0.f ? cLB.z * fr_cond_fast(cLB.w, cLB.w * cLB.w, s0, s1) : 0.f;
                    } else {
                        fL = fmaf(s0, cLB.x, s1 * cLB.y);
                        fB = fmaf(s0, cLB.z, s1 * cLB.w);
                    }
                    float LcL = 0.f, LcB = 0.f;
                    if ((kindL | kindB) & 2u) {
                        if (kindL == 2u) { const float4 q = stage[sl][2]; const float rgb[3] = { q.x, q.y, q.z }; LcL = illum_band(tb, illum_coefs(rgb), lane); }
                        if (kindB == 2u) { const float4 q = stage[sl][3]; const float rgb[3] = { q.x, q.y, q.z }; LcB = illum_band(tb, illum_coefs(rgb), lane); }
                    } else LcL = LcB = __ldg(sc.lights[bits >> 16].spectrum + lane);
                    L += fL * LcL * sc4.x + fB * LcB * sc4.y;
                }
                if (sub <= 32u) {
                    if (level == 0) wb.L[band_off(i, lane)] = L;
                    else atomicAdd(&wb.L[band_off(wb.root[i], lane)], wb.T[0][band_off(i, lane)] * L);
                } else Lacc += L;
            }
            __syncwarp();
        }
        if (sub > 32u) {
            const uint32_t i = queue[h0];
            if (level == 0) wb.L[band_off(i, lane)] = Lacc;
            else atomicAdd(&wb.L[band_off(wb.root[i], lane)], wb.T[0][band_off(i, lane)] * Lacc);
        }
    }
}

// ---- specular tree: throughput rows of the nodes K5 spawned ----------------------------------------------------------
// T[child] = T[parent] * K * (f |cos| / pdf without the spectrum), K = Kr (component 0) or Kt (1) of the parent's material;
// a warp takes four children per pass, 8 lanes x float4 per row. Level 0 parents are camera samples (T = 1, root = itself).
__global__ void __launch_bounds__(128) k_spawn_T(DevScene sc, WaveBuffers wb, int level, const uint32_t *__restrict__ queue, const uint32_t *__restrict__ count) {
    const uint32_t n = *count;
    const int lane = threadIdx.x & 31, grp = lane >> 3, bg = lane & 7;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q0 = warp * 4u; q0 < n; q0 += nwarps * 4u) {
        const uint32_t q = q0 + (uint32_t)grp;
        if (q >= n) continue;
        const uint32_t child = queue[q];
        const float4 rec = wb.g3[child];
        const uint32_t parent = __float_as_uint(rec.x), mbits = __float_as_uint(rec.y);
        const SptMaterial &m = sc.materials[mbits & 0xffffu];
        const F4 K = ld4((mbits >> 16) ? m.spec1 : m.spec0, bg);
        const F4 Tp = level == 0 ? splat4(1.f) : ld4g(wb.T[0] + (size_t)parent * NBP, bg);
        F4 Tc;
#pragma unroll
        for (int c = 0; c < 4; ++c) Tc.v[c] = Tp.v[c] * (K.v[c] * rec.z);
        st4(wb.T[0] + (size_t)child * NBP, bg, Tc);
        if (bg == 0) wb.root[child] = level == 0 ? parent : wb.root[parent];
    }
}

// ---- K7 ----------------------------------------------------------------------------------------
// Radiance guards (samplerrenderer.cpp:119-133) + SpectralImageFilm::AddSample (spectralImage.cpp:77-152).
// One warp per sampler pixel, 32 samples per pass: lane = (sub, bg) = (lane >> 3, lane & 7) reads bands 4 bg .. 4 bg + 3 of samples
// s0 + 4 k + sub, k = 0..7 - eight LDG.128 per lane in flight, each warp instruction moving four whole rows (round 2's
// first form, lane = band with one LDG.32 per row, ran at 0.37 of the HBM rate: 1.81 ms per 31.4 M samples; this one 1.01 ms). Lane j works out the filter footprint of sample
// s0 + j; a sample's y(L) is a sum over the 8 lanes of its group. Samples whose footprint is exactly the warp's own pixel
// accumulate in four registers per lane; the four groups are summed at the end and flushed with one 32-lane atomic per
// pixel (lane (sub, bg) writes band 4 bg + sub). Any other footprint (wide filters): the group's 8 lanes add its row to
// every pixel it touches.
__global__ void __launch_bounds__(256) k_film_add(FilmView film, const SptSpectralTables *tables, const float2 *img_xy,
                                                   const float *L, uint32_t cap, uint32_t n_samples, int spp) {
    const SptSpectralTables &tb = *tables;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, sub = lane >> 3, bg = lane & 7;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t npix = (n_samples + spp - 1) / spp;
    const SptFilmDesc &fd = film.d;
    const float4 cie = *((const float4 *)tb.cie_y + bg);
    const float yint = tb.yint;
    const int xs = fd.x_pixel_start, ys = fd.y_pixel_start, xe = xs + fd.x_pixel_count - 1, ye = ys + fd.y_pixel_count - 1;
    for (uint32_t pixel = warp; pixel < npix; pixel += nwarps) {
        const uint32_t first = pixel * (uint32_t)spp;
        const uint32_t ns = min((uint32_t)spp, n_samples - first);
        const float2 xy0 = img_xy[first];
        const int mainx = (int)floorf(xy0.x), mainy = (int)floorf(xy0.y);
        const bool mainInside = mainx >= xs && mainx <= xe && mainy >= ys && mainy <= ye;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f, wsum = 0.f;
        for (uint32_t s0 = 0; s0 < ns; s0 += 32u) {
            float4 Lv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t idx = min(s0 + 4u * k + (uint32_t)sub, ns - 1u);
                Lv[k] = *((const float4 *)(L + (size_t)(first + idx) * NBP) + bg);
            }
            // ---- footprint of sample s0 + lane: >= 0 the table weight of a sample that falls exactly on the warp's pixel,
            // -1 nothing to add, -2 any other footprint
            float wt = -1.f;
            if (s0 + lane < ns) {
                const float2 xy = img_xy[first + s0 + lane];
                if (xy.x > -1e29f) {                                    // else: sample outside this rank's tile set
                    const float dimageX = xy.x - 0.5f, dimageY = xy.y - 0.5f;
                    int x0 = (int)ceilf(dimageX - fd.filter_xwidth), x1 = (int)floorf(dimageX + fd.filter_xwidth);
                    int y0 = (int)ceilf(dimageY - fd.filter_ywidth), y1 = (int)floorf(dimageY + fd.filter_ywidth);
                    x0 = max(x0, xs); x1 = min(x1, xe); y0 = max(y0, ys); y1 = min(y1, ye);
                    if ((x1 - x0) >= 0 && (y1 - y0) >= 0) {
                        if (mainInside && x0 == x1 && y0 == y1 && x0 == mainx && y0 == mainy) {
                            float fx = fabsf((x0 - dimageX) * fd.filter_inv_xwidth * 16);
                            float fy = fabsf((y0 - dimageY) * fd.filter_inv_ywidth * 16);
                            int ix = min((int)floorf(fx), 15), iy = min((int)floorf(fy), 15);
                            wt = film.table[iy * 16 + ix];
                        } else wt = -2.f;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float4 v = Lv[k];
                // radiance guards (samplerrenderer.cpp:119-133): NaN anywhere, y < -1e-5, y infinite -> black
                const unsigned nanLanes = __ballot_sync(FULL, isnan(v.x) || isnan(v.y) || isnan(v.z) || isnan(v.w));
                float y = cie.x * v.x + cie.y * v.y + cie.z * v.z + cie.w * v.w;
                y += __shfl_xor_sync(FULL, y, 4);
                y += __shfl_xor_sync(FULL, y, 2);
                y += __shfl_xor_sync(FULL, y, 1);
                y = y / yint;
                const bool bad = ((nanLanes >> (8 * sub)) & 0xffu) != 0u || (double)y < -1e-5 || isinf(y);
                const float w = __shfl_sync(FULL, wt, 4 * k + sub);          // -1 beyond the pixel's samples
                if (w >= 0.f) {
                    if (!bad) { acc0 += w * v.x; acc1 += w * v.y; acc2 += w * v.z; acc3 += w * v.w; }
                    wsum += w;
                } else if (w == -2.f) {
                    const float2 xy = img_xy[first + s0 + 4u * k + (uint32_t)sub];
                    const float dimageX = xy.x - 0.5f, dimageY = xy.y - 0.5f;
                    int x0 = (int)ceilf(dimageX - fd.filter_xwidth), x1 = (int)floorf(dimageX + fd.filter_xwidth);
                    int y0 = (int)ceilf(dimageY - fd.filter_ywidth), y1 = (int)floorf(dimageY + fd.filter_ywidth);
                    x0 = max(x0, xs); x1 = min(x1, xe); y0 = max(y0, ys); y1 = min(y1, ye);
                    const float vv[4] = { bad ? 0.f : v.x, bad ? 0.f : v.y, bad ? 0.f : v.z, bad ? 0.f : v.w };
                    for (int py = y0; py <= y1; ++py) {
                        float fy = fabsf((py - dimageY) * fd.filter_inv_ywidth * 16);
                        int iy = min((int)floorf(fy), 15);
                        for (int px = x0; px <= x1; ++px) {
                            float fx = fabsf((px - dimageX) * fd.filter_inv_xwidth * 16);
                            int ix = min((int)floorf(fx), 15);
                            float wpx = film.table[iy * 16 + ix];
                            float *dst = film.pix + ((size_t)(py - ys) * fd.x_pixel_count + (px - xs)) * (NB + 1);
#pragma unroll
                            for (int c = 0; c < 4; ++c) if (4 * bg + c < NB) atomicAdd(dst + 4 * bg + c, wpx * vv[c]);
                            if (bg == 0) atomicAdd(dst + NB, wpx);
                        }
                    }
                }
            }
        }
        // the four groups' sums -> every lane; lane (sub, bg) flushes band 4 bg + sub
        acc0 += __shfl_xor_sync(FULL, acc0, 8); acc0 += __shfl_xor_sync(FULL, acc0, 16);
        acc1 += __shfl_xor_sync(FULL, acc1, 8); acc1 += __shfl_xor_sync(FULL, acc1, 16);
        acc2 += __shfl_xor_sync(FULL, acc2, 8); acc2 += __shfl_xor_sync(FULL, acc2, 16);
        acc3 += __shfl_xor_sync(FULL, acc3, 8); acc3 += __shfl_xor_sync(FULL, acc3, 16);
        wsum += __shfl_xor_sync(FULL, wsum, 8); wsum += __shfl_xor_sync(FULL, wsum, 16);
        if (mainInside && wsum != 0.f) {
            float *dst = film.pix + ((size_t)(mainy - ys) * fd.x_pixel_count + (mainx - xs)) * (NB + 1);
            const float a = sub == 0 ? acc0 : (sub == 1 ? acc1 : (sub == 2 ? acc2 : acc3));
            if (4 * bg + sub < NB) atomicAdd(dst + 4 * bg + sub, a);
            if (lane == 0) atomicAdd(dst + NB, wsum);
        }
    }
}

// film [pixel][NB+1] -> c [pixel][NB] and weight [pixel] (spt_film_download)
__global__ void k_film_split(const float *pix, size_t npix, float *c, float *w) {
    size_t total = npix * (NB + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t p = i / (NB + 1);
        int b = (int)(i % (NB + 1));
        float v = pix[i];
        if (b < NB) c[p * NB + b] = v; else w[p] = v;
    }
}
// rows of NBP floats -> [n][NB] (spt_shade_samples output)
__global__ void k_gather_L(const float *L, uint32_t cap, uint32_t n, float *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n * NB; i += gridDim.x * blockDim.x) {
        uint32_t s = i / NB, c = i % NB;
        out[i] = L[band_off(s, c)];
    }
}
// [n][NB] -> rows of NBP floats (zero padding), for spt_film_add_samples
__global__ void k_scatter_L(const float *in, uint32_t cap, uint32_t n, float *L) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n * NB; i += gridDim.x * blockDim.x) {
        uint32_t s = i / NB, c = i % NB;
        L[band_off(s, c)] = in[i];
        if (NB < NBP && c == 0) for (int k = NB; k < NBP; ++k) L[band_off(s, k)] = 0.f;
    }
}

// ---- launchers -------------------------------------------------------------------------------------
static inline unsigned grid1d(uint64_t n, int per, unsigned cap) { uint64_t g = (n + per - 1) / per; return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g)); }
void spt_launch_compact_hits(int grid, cudaStream_t st, const uint32_t *queue, const uint32_t *count, const uint32_t *hit_slot,
                             uint32_t *hit_queue, uint32_t *hit_count, uint32_t *miss_queue, uint32_t *miss_count, float *black_L) {
    k_compact_hits<<<grid, 256, 0, st>>>(queue, count, hit_slot, hit_queue, hit_count, miss_queue, miss_count, black_L);
}
void spt_launch_miss_env(int grid, cudaStream_t st, const DevScene &sc, const WaveBuffers &wb, int bounce, int tree, const uint32_t *queue, const uint32_t *count) {
    k_miss_env<<<grid, 128, 0, st>>>(sc, wb, bounce, tree, queue, count);
}
void spt_launch_spawn_T(int grid, cudaStream_t st, const DevScene &sc, const WaveBuffers &wb, int level, const uint32_t *queue, const uint32_t *count) {
    k_spawn_T<<<grid, 128, 0, st>>>(sc, wb, level, queue, count);
}
void spt_launch_shade(int grid, cudaStream_t st, const DevScene &sc, const RenderCfg &cfg, const SampleSource &src, const WaveBuffers &wb,
                      int bounce, const uint32_t *queue, const uint32_t *count, uint32_t *shadow_count, uint32_t *mis_count,
                      uint32_t *elided_count, uint32_t *mis_any_count, uint32_t *next_queue, uint32_t *next_count, uint32_t *node_ctr) {
#define SPT_SHADE4(S, E, D, M) k_shade<S, E, D, M><<<(grid * 128 + SHADE_THREADS(E) - 1) / SHADE_THREADS(E), SHADE_THREADS(E), 0, st>>>(sc, cfg, src, wb, bounce, queue, count, shadow_count, mis_count, elided_count, mis_any_count, next_queue, next_count, node_ctr)
#define SPT_SHADE(S, E) SPT_SHADE4(S, E, false, false)
    if (cfg.integrator == SPT_INTEGRATOR_DIRECT_ALL) {
        if (sc.has_specular) { if (sc.has_measured) SPT_SHADE4(true, true, true, true); else SPT_SHADE4(true, true, true, false); }
        else if (sc.has_measured) SPT_SHADE4(false, true, true, true); else SPT_SHADE4(false, true, true, false);
    } else if (sc.has_measured) { if (sc.has_specular) SPT_SHADE4(true, true, false, true); else SPT_SHADE4(false, true, false, true); }
    else if (sc.has_ext || cfg.integrator != SPT_INTEGRATOR_PATH) { if (sc.has_specular) SPT_SHADE(true, true); else SPT_SHADE(false, true); }
    else { if (sc.has_specular) SPT_SHADE(true, false); else SPT_SHADE(false, false); }
#undef SPT_SHADE
#undef SPT_SHADE4
}
void spt_launch_advance(int grid, cudaStream_t st, const DevScene &sc, const RenderCfg &cfg, const WaveBuffers &wb, int bounce,
                        const uint32_t *queue, const uint32_t *count, uint32_t *next_queue, uint32_t *next_count) {
    if (sc.has_ext) k_advance<true><<<grid, 128, 0, st>>>(sc, cfg, wb, bounce, queue, count, next_queue, next_count);
    else k_advance<false><<<grid, 128, 0, st>>>(sc, cfg, wb, bounce, queue, count, next_queue, next_count);
}
void spt_launch_addlight(int grid, cudaStream_t st, const DevScene &sc, const RenderCfg &cfg, const WaveBuffers &wb, int bounce,
                         const uint32_t *queue, const uint32_t *count) {
    if (cfg.integrator == SPT_INTEGRATOR_DIRECT_ALL) k_accumulate_direct<<<grid, 128, 0, st>>>(sc, cfg, wb, cfg.tree ? bounce : 0, queue, count);
    else if (sc.has_ext) k_addlight<true><<<grid, 128, 0, st>>>(sc, cfg, wb, bounce, queue, count);
    else k_addlight<false><<<grid, 128, 0, st>>>(sc, cfg, wb, bounce, queue, count);
}
void spt_launch_film_add(int grid, cudaStream_t st, const FilmView &film, const SptSpectralTables *tables, const float2 *img_xy,
                         const float *L, uint32_t cap, uint32_t n_samples, int spp) {
    k_film_add<<<grid, 256, 0, st>>>(film, tables, img_xy, L, cap, n_samples, spp);
}
void spt_launch_film_split(int grid, cudaStream_t st, const float *pix, size_t npix, float *c, float *w) {
    k_film_split<<<grid, 256, 0, st>>>(pix, npix, c, w);
}
void spt_launch_gather_L(cudaStream_t st, const float *L, uint32_t cap, uint32_t n, float *out) {
    k_gather_L<<<grid1d((uint64_t)n * NB, 256, 65535), 256, 0, st>>>(L, cap, n, out);
}
void spt_launch_scatter_L(cudaStream_t st, const float *in, uint32_t cap, uint32_t n, float *L) {
    k_scatter_L<<<grid1d((uint64_t)n * NB, 256, 65535), 256, 0, st>>>(in, cap, n, L);
}

// spt_kernels.cuh — the wavefront kernels K1..K7 (sm_100a, FP32/FP64/INT pipes; nothing here is a
// dense contraction, so no tensor cores — SURVEY.md 8a).
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
#include "shade.cuh"
#include "sampler.cuh"

struct WaveBuffers {
    uint32_t cap;
    float4 *ray_o, *ray_d;            // path rays: {o, mint}, {d, maxt}
    uint32_t *hit_slot; float *hit_t;
    float4 *g0, *g1, *g2, *g3;        // {p, eps}, {shadow d, shadow maxt}, {mis d, inf}, {path d, -}
    uint32_t *mis_slot; float *mis_t; uint32_t *sh_slot;
    float4 *r0, *r1, *r2, *r3, *r4, *r5;
    uint4 *r6;
    float2 *img_xy;
    float *T, *L;                     // [NB][cap]
    uint32_t *pathQ[2], *shadowQ, *misQ;
};

struct RenderCfg {
    SptCameraDesc cam;
    int spp, max_depth;
    int x0, x1, y0, y1;               // sample extent (x1,y1 exclusive, already border-trimmed)
    int tile, tilesX, tilesY, rank, nranks;
    uint32_t seed;
    uint64_t pixel_base;              // first rank-local pixel of this wave
    uint32_t n_samples;               // samples in this wave
};

// flags in r6.x
enum { RF_L = 1, RF_LDELTA = 2, RF_B = 4, RF_P = 8, RF_L_REFL = 16, RF_L_MF = 32, RF_B_REFL = 64, RF_B_MF = 128,
       RF_P_REFL = 256, RF_P_MF = 512, RF_ON = 1024 };

__device__ __forceinline__ bool wave_pixel(const RenderCfg &cfg, uint64_t j, int *px, int *py) {
    uint32_t tp = (uint32_t)(cfg.tile * cfg.tile);
    uint64_t lt = j / tp;
    uint32_t w = (uint32_t)(j % tp);
    uint64_t tile = lt * (uint64_t)cfg.nranks + (uint64_t)cfg.rank;
    if (tile >= (uint64_t)cfg.tilesX * cfg.tilesY) return false;
    int tx = (int)(tile % cfg.tilesX), ty = (int)(tile / cfg.tilesX);
    *px = cfg.x0 + tx * cfg.tile + (int)(w % cfg.tile);
    *py = cfg.y0 + ty * cfg.tile + (int)(w / cfg.tile);
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

// ---- K1 ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gen_camera(RenderCfg cfg, SampleSource src, WaveBuffers wb, uint32_t *count_out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cfg.n_samples; i += gridDim.x * blockDim.x) {
        float ix, iy, lu, lv;
        bool valid = true;
        if (src.smp) {
            const float *s = src.smp + 37 * (size_t)i;
            ix = s[0]; iy = s[1]; lu = s[2]; lv = s[3];
        } else {
            int px, py;
            uint32_t s = i % (uint32_t)cfg.spp;
            valid = wave_pixel(cfg, cfg.pixel_base + i / (uint32_t)cfg.spp, &px, &py);
            if (valid) {
                uint32_t pk = pixel_key(src.seed, pix_key(px, py));
                float t2[2];
                ld2(pk, 0, s, src.spp, t2);
                ix = px + t2[0]; iy = py + t2[1];
                ld2(pk, 1, s, src.spp, t2);
                lu = t2[0]; lv = t2[1];
            }
        }
        if (valid) {
            Ray ray;
            camera_ray(cfg.cam, ix, iy, lu, lv, &ray);
            wb.ray_o[i] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
            wb.ray_d[i] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
            wb.img_xy[i] = make_float2(ix, iy);
        } else {
            wb.img_xy[i] = make_float2(-1e30f, -1e30f);
        }
        queue_push(wb.pathQ[0], count_out, valid, i);
    }
}

// ---- K2 / K3 -------------------------------------------------------------------------------------
// BVHAccel::Intersect / IntersectP (src/accelerators/bvh.cpp:380-432, :435-481) as a persistent-warp
// kernel: every lane owns one ray and walks the reference's depth-first node array with the
// reference's 64-entry todo stack, slab test, near/far rule, leaf order and `t <= maxt` acceptance
// (ties resolve to the same primitive). Lanes whose ray has terminated sit idle only
// until fewer than FETCH_THRESHOLD lanes of the warp are still traversing; then the warp drops out
// of the traversal loop and the idle lanes pull the next rays from the device-side queue with one
// warp-aggregated atomic (incoherent bounce / MIS rays otherwise leave ~8 of 32 lanes busy).
// Node = 2 x LDG.128, triangle = 3 x LDG.128 (vertices pre-gathered per BVH slot).
#define FETCH_THRESHOLD 20
template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(128) k_trace(DevScene sc, const uint32_t *__restrict__ queue, const uint32_t *__restrict__ count,
                                               uint32_t *__restrict__ work, const float4 *__restrict__ ro,
                                               const float4 *__restrict__ rd, uint32_t *__restrict__ out_slot,
                                               float *__restrict__ out_t) {
    const uint32_t n = *count;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    uint32_t todo[64];
    uint32_t todoOffset = 0, nodeNum = 0, best = SPT_MISS, i = 0;
    Ray ray; ray.o = V(0, 0, 0); ray.d = V(0, 0, 1); ray.mint = 0.f; ray.maxt = 0.f;
    v3 invDir = V(0, 0, 0);
    bool negx = false, negy = false, negz = false;
    bool active = false, exhausted = (n == 0);
    unsigned long long cn = 0, cp = 0;
    for (;;) {
        // ---- fetch: idle lanes take the next rays of the queue
        unsigned idle = __ballot_sync(FULL, !active);
        if (!exhausted && idle) {
            uint32_t base = 0;
            int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(work, (uint32_t)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (!active) {
                uint32_t q = base + __popc(idle & ((1u << lane) - 1));
                if (q < n) {
                    i = queue ? queue[q] : q;
                    float4 o = ro[i], d = rd[i];
                    ray.o = V(o.x, o.y, o.z); ray.d = V(d.x, d.y, d.z); ray.mint = o.w; ray.maxt = d.w;
                    invDir = V(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
                    negx = invDir.x < 0; negy = invDir.y < 0; negz = invDir.z < 0;
                    todoOffset = 0; nodeNum = 0; best = SPT_MISS;
                    active = true;
                    if (sc.n_nodes == 0) { out_slot[i] = SPT_MISS; if (!ANY) out_t[i] = ray.maxt; active = false; }
                }
            }
            if (base + (uint32_t)__popc(idle) >= n) exhausted = true;
        }
        if (!__any_sync(FULL, active)) break;
        // ---- traverse until this lane's ray terminates or the warp has gone too idle
        while (active) {
            float4 n0 = __ldg(&sc.nodes[2 * (size_t)nodeNum]);
            float4 n1 = __ldg(&sc.nodes[2 * (size_t)nodeNum + 1]);
            if (COUNT) ++cn;
            bool pop = true;
            if (slab(n0, n1, ray, invDir, negx, negy, negz)) {
                uint32_t meta = __float_as_uint(n1.w);
                uint32_t offset = __float_as_uint(n1.z);
                uint32_t nPrims = meta & 0xff;
                if (nPrims > 0) {
                    bool hasQuadric = (meta >> 16) & 1;
                    for (uint32_t k = 0; k < nPrims; ++k) {
                        uint32_t s = offset + k;
                        if (COUNT) ++cp;
                        float t;
                        bool h;
                        if (!hasQuadric || sc.prim_kind[s] == SPT_PRIM_TRIANGLE) {
                            float4 a = __ldg(&sc.tri_verts[3 * (size_t)s]);
                            float4 b = __ldg(&sc.tri_verts[3 * (size_t)s + 1]);
                            float4 c = __ldg(&sc.tri_verts[3 * (size_t)s + 2]);
                            float b1, b2;
                            h = tri_test(V(a.x, a.y, a.z), V(b.x, b.y, b.z), V(c.x, c.y, c.z), ray, &t, &b1, &b2);
                        } else if (sc.prim_kind[s] == SPT_PRIM_SPHERE) {
                            h = sphere_intersect(sc, sc.quadrics[sc.prim_data[s]], 0, ray, &t, nullptr);
                        } else {
                            h = disk_intersect(sc, sc.quadrics[sc.prim_data[s]], 0, ray, &t, nullptr);
                        }
                        if (h) {
                            best = s;
                            if (ANY) break;            // IntersectP returns at the first accepted primitive
                            ray.maxt = t;
                        }
                    }
                } else {
                    uint32_t axis = (meta >> 8) & 0xff;
                    bool neg = axis == 0 ? negx : (axis == 1 ? negy : negz);
                    if (neg) { todo[todoOffset++] = nodeNum + 1; nodeNum = offset; }
                    else { todo[todoOffset++] = offset; nodeNum = nodeNum + 1; }
                    pop = false;
                }
            }
            if (pop) {
                if (todoOffset == 0 || (ANY && best != SPT_MISS)) {
                    out_slot[i] = best;
                    if (!ANY) out_t[i] = ray.maxt;
                    active = false;
                } else nodeNum = todo[--todoOffset];
            }
            if (active && !exhausted && __popc(__activemask()) < FETCH_THRESHOLD) break;
        }
    }
    if (COUNT && sc.counters && (cn | cp)) {
        atomicAdd(&sc.counters[ANY ? 2 : 0], cn);
        atomicAdd(&sc.counters[ANY ? 3 : 1], cp);
    }
}

// ---- K5 ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_shade(DevScene sc, RenderCfg cfg, SampleSource src, WaveBuffers wb, int bounce,
                                               const uint32_t *queue, const uint32_t *count,
                                               uint32_t *shadow_count, uint32_t *mis_count) {
    uint32_t n = *count;
    const uint32_t cap = wb.cap;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < ((n + 31u) & ~31u); q += gridDim.x * blockDim.x) {
        bool active = q < n;
        uint32_t i = active ? queue[q] : 0;
        bool pushShadow = false, pushMis = false;
        if (active) {
            uint32_t slot = wb.hit_slot[i];
            float4 o4 = wb.ray_o[i], d4 = wb.ray_d[i];
            Ray ray;
            ray.o = V(o4.x, o4.y, o4.z); ray.d = V(d4.x, d4.y, d4.z); ray.mint = o4.w; ray.maxt = SPT_INF;
            uint32_t flags = 0;
            if (slot == SPT_MISS) {
                // SamplerRenderer::Li miss branch (samplerrenderer.cpp:239-243): sum of Light::Le over
                // all lights, only the infinite light is non-zero. Later bounces add Le only after a
                // specular bounce (path.cpp:106-108), which the lowered BxDFs never produce.
                if (bounce == 0) {
                    bool haveEnv = false;
                    IllumCoefs k;
                    for (uint32_t l = 0; l < sc.n_lights; ++l)
                        if (sc.lights[l].type == SPT_LIGHT_INFINITE) {
                            float rgb[3];
                            infinite_le_rgb(sc, sc.lights[l], ray.d, rgb);
                            k = illum_coefs(rgb);
                            haveEnv = true;
                        }
                    for (int c = 0; c < NB; ++c) wb.L[(size_t)c * cap + i] = haveEnv ? illum_band(*sc.tables, k, c) : 0.f;
                }
                wb.r6[i] = make_uint4(0xffffffffu, 0, 0, 0);     // dead marker for K6
            } else {
                Hit hit;
                shape_intersect(sc, sc.prim_kind[slot], sc.prim_flags[slot], sc.prim_data[slot], ray, &hit);
                // emitted light at the first vertex (path.cpp:55-56; Intersection::Le, intersection.cpp:53-56)
                if (bounce == 0) {
                    int li = sc.prim_light[slot];
                    bool on = li >= 0 && dot(hit.nn, vneg(ray.d)) > 0.f;
                    for (int c = 0; c < NB; ++c) wb.L[(size_t)c * cap + i] = on ? sc.lights[li].spectrum[c] : 0.f;
                }
                Bsdf bsdf;
                make_bsdf(sc, slot, hit, &bsdf);
                v3 p = hit.p, n_s = bsdf.nn, woW = vneg(ray.d);
                v3 wo = w2l(bsdf, woW);
                float eps = hit.rayEpsilon;
                uint32_t s_idx = src.smp ? 0u : (i % (uint32_t)cfg.spp);
                uint32_t pk = 0;
                if (!src.smp) {
                    int px, py;
                    wave_pixel(cfg, cfg.pixel_base + i / (uint32_t)cfg.spp, &px, &py);
                    pk = pixel_key(src.seed, pix_key(px, py));
                }
                float u[10], rr;
                bounce_dims(src, i, pk, s_idx, bounce, sc.n_lights > 0, u, &rr);

                float4 r0 = make_float4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0, r4 = r0, r5 = r0;
                float4 g1 = r0, g2 = r0, g3 = r0;
                int lightIdx = 0;
                if (bsdf.orenNayar) flags |= RF_ON;
                if (sc.n_lights > 0) {
                    // UniformSampleOneLight (integrator.cpp:74-106) + EstimateDirect (:109-166)
                    int nLights = (int)sc.n_lights;
                    lightIdx = (int)floorf(u[0] * nLights);
                    if (nLights - 1 < lightIdx) lightIdx = nLights - 1;
                    LightSampleResult lr;
                    light_sample(sc, lightIdx, p, u[1], u[2], u[3], &lr);
                    if (lr.pdf > 0.f && !lr.black) {
                        DirTerms t;
                        v3 wi = w2l(bsdf, lr.wi);
                        float bsdfPdf;
                        bsdf_terms(bsdf, woW, lr.wi, wo, wi, &t, &bsdfPdf);
                        if (t.reflect) {
                            float sL;
                            if (lr.delta) sL = (absdot(lr.wi, n_s) / lr.pdf);
                            else {
                                float weight = power_heuristic(lr.pdf, bsdfPdf);
                                sL = (absdot(lr.wi, n_s) * weight / lr.pdf);
                            }
                            flags |= RF_L | (lr.delta ? RF_LDELTA : 0) | RF_L_REFL | (t.mf ? RF_L_MF : 0);
                            r0 = make_float4(t.a0, t.a1, t.a2, t.a3);
                            r3.x = sL;
                            r5 = make_float4(lr.aux[0], lr.aux[1], lr.aux[2], 0.f);
                            g1 = make_float4(lr.shadow_d.x, lr.shadow_d.y, lr.shadow_d.z, lr.shadow_maxt);
                            pushShadow = true;
                        }
                    }
                    if (!lr.delta) {
                        v3 wiW; float bsdfPdf; DirTerms t;
                        bsdf_sample(bsdf, woW, wo, u[6], u[4], u[5], &wiW, &bsdfPdf, &t);
                        if (bsdfPdf > 0.f && t.reflect) {
                            float lightPdf = light_pdf(sc, lightIdx, p, wiW);
                            if (lightPdf != 0.f) {
                                float weight = power_heuristic(bsdfPdf, lightPdf);
                                flags |= RF_B | RF_B_REFL | (t.mf ? RF_B_MF : 0);
                                r1 = make_float4(t.a0, t.a1, t.a2, t.a3);
                                r3.y = absdot(wiW, n_s); r3.z = weight; r3.w = bsdfPdf;
                                g2 = make_float4(wiW.x, wiW.y, wiW.z, SPT_INF);
                                pushMis = true;
                            }
                        }
                    }
                }
                {   // continuation direction (path.cpp:75-92)
                    v3 wiW; float pdf; DirTerms t;
                    bsdf_sample(bsdf, woW, wo, u[9], u[7], u[8], &wiW, &pdf, &t);
                    if (pdf != 0.f) {
                        flags |= RF_P | (t.reflect ? RF_P_REFL : 0) | (t.mf ? RF_P_MF : 0);
                        r2 = make_float4(t.a0, t.a1, t.a2, t.a3);
                        r4.x = absdot(wiW, n_s); r4.y = pdf;
                        g3 = make_float4(wiW.x, wiW.y, wiW.z, 0.f);
                    }
                    r4.z = rr;
                }
                wb.g0[i] = make_float4(p.x, p.y, p.z, eps);
                wb.g1[i] = g1; wb.g2[i] = g2; wb.g3[i] = g3;
                wb.r0[i] = r0; wb.r1[i] = r1; wb.r2[i] = r2; wb.r3[i] = r3; wb.r4[i] = r4; wb.r5[i] = r5;
                wb.r6[i] = make_uint4(flags, (uint32_t)sc.prim_material[slot], (uint32_t)lightIdx, 0);
            }
        }
        queue_push(wb.shadowQ, shadow_count, pushShadow, i);
        queue_push(wb.misQ, mis_count, pushMis, i);
    }
}

// ---- K6 ----------------------------------------------------------------------------------------
// Per-direction BSDF value rebuilt per band from wavelength-independent coefficients:
//   matte / plastic:  f[c] = spec0[c]*u0 + spec1[c]*u1     (Lambert/Oren-Nayar + Blinn microfacet x dielectric Fresnel)
//   metal:            f[c] = w * FrCond(cosH, eta[c], k[c])
// The scalar factors (D*G*F/(4 cosI cosO), |cos|*weight/pdf, ...) are folded once per vertex instead of
// once per band as the reference's Spectrum arithmetic does; this reassociation moves results by
// rounding only (tests compare radiance at 2e-4 relative).
struct DirCoef { float u0, u1, w, cosH; };
__device__ __forceinline__ DirCoef dir_coef(int mtype, bool on, bool reflect, bool mf, float4 a) {
    DirCoef d; d.u0 = d.u1 = d.w = 0.f; d.cosH = 1.f;
    if (!reflect) return d;
    if (mtype == SPT_MAT_MATTE) d.u0 = on ? INV_PI_F * a.x : INV_PI_F;
    else if (mtype == SPT_MAT_PLASTIC) { d.u0 = INV_PI_F; if (mf) d.u1 = a.x * a.y * a.z / a.w; }
    else if (mf) { d.w = a.x * a.y / a.w; d.cosH = a.z; }
    return d;
}
// FrCond (reflection.cpp:63-71) with the two quotients combined into one division
__device__ __forceinline__ float fr_cond_fast(float cosi, float c2, float eta, float k) {
    float A = fmaf(eta, eta, k * k);
    float e2 = 2.f * eta * cosi;
    float Ac2 = A * c2;
    float n1 = Ac2 - e2 + 1.f, d1 = Ac2 + e2 + 1.f;
    float n2 = A - e2 + c2, d2 = A + e2 + c2;
    return 0.5f * __fdividef(fmaf(n1, d2, n2 * d1), d1 * d2);
}
struct LightBand { int kind; const float *spec; IllumCoefs k; };   // kind: 0 none, 1 table spectrum, 2 rgb illuminant
__device__ __forceinline__ float light_band(const SptSpectralTables &tb, const LightBand &l, int c) {
    return l.kind == 1 ? __ldg(l.spec + c) : (l.kind == 2 ? illum_band(tb, l.k, c) : 0.f);
}

__global__ void __launch_bounds__(128, 4) k_accumulate(DevScene sc, RenderCfg cfg, WaveBuffers wb, int bounce,
                                                       const uint32_t *__restrict__ queue, const uint32_t *__restrict__ count,
                                                       uint32_t *__restrict__ next_queue, uint32_t *__restrict__ next_count) {
    uint32_t n = *count;
    const uint32_t cap = wb.cap;
    const SptSpectralTables &tb = *sc.tables;
    float *__restrict__ Tg = wb.T;
    float *__restrict__ Lg = wb.L;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < ((n + 31u) & ~31u); q += gridDim.x * blockDim.x) {
        bool active = q < n;
        uint32_t i = active ? queue[q] : 0;
        bool alive = false;
        uint4 r6 = active ? wb.r6[i] : make_uint4(0xffffffffu, 0, 0, 0);
        if (active && r6.x != 0xffffffffu) {
            uint32_t flags = r6.x;
            const SptMaterial &m = sc.materials[r6.y];
            const int mtype = m.type;
            int lightIdx = (int)r6.z;
            bool on = (flags & RF_ON) != 0;
            float4 r3 = wb.r3[i], r4 = wb.r4[i];
            float4 g0 = wb.g0[i];
            // --- light-sample term (integrator.cpp:122-137): visible iff the shadow ray found nothing
            DirCoef cL = dir_coef(mtype, on, false, false, make_float4(0, 0, 0, 0)), cB = cL, cP = cL;
            LightBand lbL, lbB; lbL.kind = 0; lbB.kind = 0; lbL.spec = lbB.spec = nullptr;
            float sL = 0.f, sB = 0.f, sP = 0.f;
            if ((flags & RF_L) && wb.sh_slot[i] == SPT_MISS) {
                float4 r5 = wb.r5[i];
                cL = dir_coef(mtype, on, true, (flags & RF_L_MF) != 0, wb.r0[i]);
                const SptLight &l = sc.lights[lightIdx];
                sL = r3.x;
                if (l.type == SPT_LIGHT_INFINITE) { float rgb[3] = { r5.x, r5.y, r5.z }; lbL.kind = 2; lbL.k = illum_coefs(rgb); }
                else { lbL.kind = 1; lbL.spec = l.spectrum; if (l.type == SPT_LIGHT_POINT) sL = sL / r5.x; }
            }
            // --- BSDF-sample term (integrator.cpp:139-163): radiance from what the MIS ray found
            if (flags & RF_B) {
                float4 g2 = wb.g2[i];
                uint32_t ms = wb.mis_slot[i];
                const SptLight &l = sc.lights[lightIdx];
                v3 wi = V(g2.x, g2.y, g2.z);
                if (ms != SPT_MISS) {
                    if (sc.prim_light[ms] == lightIdx) {
                        Ray ray; ray.o = V(g0.x, g0.y, g0.z); ray.d = wi; ray.mint = g0.w; ray.maxt = SPT_INF;
                        Hit h;
                        if (shape_intersect(sc, sc.prim_kind[ms], sc.prim_flags[ms], sc.prim_data[ms], ray, &h) &&
                            dot(h.nn, vneg(wi)) > 0.f) { lbB.kind = 1; lbB.spec = l.spectrum; }
                    }
                } else if (l.type == SPT_LIGHT_INFINITE) {
                    float rgb[3];
                    infinite_le_rgb(sc, l, wi, rgb);
                    lbB.kind = 2; lbB.k = illum_coefs(rgb);
                }
                if (lbB.kind) { cB = dir_coef(mtype, on, true, (flags & RF_B_MF) != 0, wb.r1[i]); sB = r3.y * r3.z / r3.w; }
            }
            bool haveP = (flags & RF_P) != 0;
            if (haveP) { cP = dir_coef(mtype, on, (flags & RF_P_REFL) != 0, (flags & RF_P_MF) != 0, wb.r2[i]); sP = r4.x / r4.y; }
            const float nL = (float)sc.n_lights;
            sL *= nL; sB *= nL;
            const bool metal = mtype == SPT_MAT_METAL;
            const float cL2 = cL.cosH * cL.cosH, cB2 = cB.cosH * cB.cosH, cP2 = cP.cosH * cP.cosH;
            float yy = 0.f;
            bool fBlack = true;
            const float *__restrict__ s0p = m.spec0;
            const float *__restrict__ s1p = m.spec1;
            // the one band loop of the bounce: L += T * Ld * nLights ; T *= f |cos| / pdf
            // (8 bands per trip: 16 independent loads in flight before any store)
#pragma unroll 1
            for (int c0 = 0; c0 < NB; c0 += 8) {
                float Tv[8], Lv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    size_t off = (size_t)(c0 + k) * cap + i;
                    Tv[k] = bounce == 0 ? 1.f : Tg[off];
                    Lv[k] = Lg[off];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int c = c0 + k;
                    float s0 = __ldg(s0p + c), s1 = __ldg(s1p + c);
                    float fL, fB, fP;
                    if (metal) {
                        fL = cL.w != 0.f ? cL.w * fr_cond_fast(cL.cosH, cL2, s0, s1) : 0.f;
                        fB = cB.w != 0.f ? cB.w * fr_cond_fast(cB.cosH, cB2, s0, s1) : 0.f;
                        fP = cP.w != 0.f ? cP.w * fr_cond_fast(cP.cosH, cP2, s0, s1) : 0.f;
                    } else {
                        fL = fmaf(s0, cL.u0, s1 * cL.u1);
                        fB = fmaf(s0, cB.u0, s1 * cB.u1);
                        fP = fmaf(s0, cP.u0, s1 * cP.u1);
                    }
                    float Ld = fL * light_band(tb, lbL, c) * sL + fB * light_band(tb, lbB, c) * sB;
                    size_t off = (size_t)c * cap + i;
                    Lg[off] = fmaf(Tv[k], Ld, Lv[k]);
                    if (fP != 0.f) fBlack = false;
                    float Tn = Tv[k] * (fP * sP);
                    Tg[off] = Tn;
                    yy = fmaf(tb.cie_y[c], Tn, yy);
                }
            }
            // path.cpp:88-104
            if (haveP && !fBlack) {
                alive = true;
                if (bounce > 3) {
                    float continueProbability = stdminf(.5f, yy / tb.yint);
                    if (r4.z > continueProbability) alive = false;
                    else if (bounce != cfg.max_depth) {
                        float inv = 1.f / continueProbability;
                        for (int c = 0; c < NB; ++c) Tg[(size_t)c * cap + i] *= inv;
                    }
                }
                if (bounce == cfg.max_depth) alive = false;
                if (alive) {
                    float4 g3 = wb.g3[i];
                    wb.ray_o[i] = g0;
                    wb.ray_d[i] = make_float4(g3.x, g3.y, g3.z, SPT_INF);
                }
            }
        }
        queue_push(next_queue, next_count, alive, i);
    }
}

// ---- K7 ----------------------------------------------------------------------------------------
// Radiance guards (samplerrenderer.cpp:119-133) + SpectralImageFilm::AddSample
// (spectralImage.cpp:77-152). One warp per sampler pixel: lanes stride over the pixel's samples;
// contributions whose footprint is exactly that pixel are reduced across the warp with shuffles and
// flushed with one atomic per band; anything else (wide filters, samples rounding onto a pixel
// edge) goes straight to global atomics.
struct FilmView {
    SptFilmDesc d;
    float *pix;            // [y][x][NB+1]
    const float *table;    // 256 filter weights in global memory
};
__global__ void __launch_bounds__(256) k_film_add(FilmView film, const SptSpectralTables *tables, const float2 *img_xy,
                                                  const float *L, uint32_t cap, uint32_t n_samples, int spp) {
    const SptSpectralTables &tb = *tables;
    int lane = threadIdx.x & 31;
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t npix = (n_samples + spp - 1) / spp;
    const SptFilmDesc &fd = film.d;
    for (uint32_t pixel = warp; pixel < npix; pixel += nwarps) {
        // the pixel this warp reduces into: the one the first sample of the group falls in
        float2 xy0 = img_xy[(size_t)pixel * spp];
        int mainx = (int)floorf(xy0.x), mainy = (int)floorf(xy0.y);
        bool mainInside = mainx >= fd.x_pixel_start && mainx < fd.x_pixel_start + fd.x_pixel_count &&
                          mainy >= fd.y_pixel_start && mainy < fd.y_pixel_start + fd.y_pixel_count;
        float wsum = 0.f;
        for (int s0 = 0; s0 < spp; s0 += 32) {
            int s = s0 + lane;
            uint32_t i = pixel * spp + s;
            bool have = s < spp && i < n_samples;
            float2 xy = have ? img_xy[i] : make_float2(-1e30f, -1e30f);
            have = have && xy.x > -1e29f;
            bool bad = false;
            float y = 0.f;
            if (have) {
                for (int c = 0; c < NB; ++c) {
                    float v = L[(size_t)c * cap + i];
                    if (isnan(v)) bad = true;
                    y += tb.cie_y[c] * v;
                }
                y = y / tb.yint;
                if ((double)y < -1e-5 || isinf(y)) bad = true;
            }
            int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
            float dimageX = xy.x - 0.5f, dimageY = xy.y - 0.5f;
            if (have) {
                x0 = (int)ceilf(dimageX - fd.filter_xwidth); x1 = (int)floorf(dimageX + fd.filter_xwidth);
                y0 = (int)ceilf(dimageY - fd.filter_ywidth); y1 = (int)floorf(dimageY + fd.filter_ywidth);
                x0 = max(x0, fd.x_pixel_start); x1 = min(x1, fd.x_pixel_start + fd.x_pixel_count - 1);
                y0 = max(y0, fd.y_pixel_start); y1 = min(y1, fd.y_pixel_start + fd.y_pixel_count - 1);
            }
            bool any = have && (x1 - x0) >= 0 && (y1 - y0) >= 0;
            bool fast = any && mainInside && x0 == x1 && y0 == y1 && x0 == mainx && y0 == mainy;
            float wfast = 0.f;
            if (fast) {
                float fx = fabsf((x0 - dimageX) * fd.filter_inv_xwidth * 16);
                float fy = fabsf((y0 - dimageY) * fd.filter_inv_ywidth * 16);
                int ix = min((int)floorf(fx), 15), iy = min((int)floorf(fy), 15);
                wfast = film.table[iy * 16 + ix];
            }
            if (any && !fast) {
                for (int yy = y0; yy <= y1; ++yy) {
                    float fy = fabsf((yy - dimageY) * fd.filter_inv_ywidth * 16);
                    int iy = min((int)floorf(fy), 15);
                    for (int xx = x0; xx <= x1; ++xx) {
                        float fx = fabsf((xx - dimageX) * fd.filter_inv_xwidth * 16);
                        int ix = min((int)floorf(fx), 15);
                        float wt = film.table[iy * 16 + ix];
                        float *dst = film.pix + ((size_t)(yy - fd.y_pixel_start) * fd.x_pixel_count + (xx - fd.x_pixel_start)) * (NB + 1);
                        for (int c = 0; c < NB; ++c) atomicAdd(dst + c, wt * (bad ? 0.f : L[(size_t)c * cap + i]));
                        atomicAdd(dst + NB, wt);
                    }
                }
            }
            // warp reduction of the fast-path contributions, band by band (fixed shuffle tree)
            float *dst = film.pix + ((size_t)(mainy - fd.y_pixel_start) * fd.x_pixel_count + (mainx - fd.x_pixel_start)) * (NB + 1);
            if (__any_sync(0xffffffffu, fast)) {
                for (int c = 0; c < NB; ++c) {
                    float v = (fast && !bad) ? wfast * L[(size_t)c * cap + i] : 0.f;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    if (lane == 0) atomicAdd(dst + c, v);
                }
                float w = fast ? wfast : 0.f;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(0xffffffffu, w, off);
                wsum += w;
            }
        }
        if (lane == 0 && wsum != 0.f && mainInside) {
            float *dst = film.pix + ((size_t)(mainy - fd.y_pixel_start) * fd.x_pixel_count + (mainx - fd.x_pixel_start)) * (NB + 1);
            atomicAdd(dst + NB, wsum);
        }
    }
}

// SoA [NB][cap] -> AoS [n][NB] (spt_shade_samples output)
__global__ void k_gather_L(const float *L, uint32_t cap, uint32_t n, float *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n * NB; i += gridDim.x * blockDim.x) {
        uint32_t s = i / NB, c = i % NB;
        out[i] = L[(size_t)c * cap + s];
    }
}
// AoS [n][NB] -> SoA, for spt_film_add_samples
__global__ void k_scatter_L(const float *in, uint32_t cap, uint32_t n, float *L) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n * NB; i += gridDim.x * blockDim.x) {
        uint32_t s = i / NB, c = i % NB;
        L[(size_t)c * cap + s] = in[i];
    }
}
__global__ void k_camera_rays(SptCameraDesc cam, const float *samples, uint32_t n, float *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *s = samples + 5 * (size_t)i;
        Ray ray;
        camera_ray(cam, s[0], s[1], s[2], s[3], &ray);
        float *o = out + 8 * (size_t)i;
        o[0] = ray.o.x; o[1] = ray.o.y; o[2] = ray.o.z; o[3] = ray.d.x; o[4] = ray.d.y; o[5] = ray.d.z;
        o[6] = ray.mint; o[7] = ray.maxt;
    }
}
// rays n x 8 {o,d,mint,maxt} -> the two float4 arrays the trace kernels read
__global__ void k_split_rays(const float *rays, uint32_t n, float4 *ro, float4 *rd) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *r = rays + 8 * (size_t)i;
        ro[i] = make_float4(r[0], r[1], r[2], r[6]);
        rd[i] = make_float4(r[3], r[4], r[5], r[7]);
    }
}
__global__ void k_slot_to_id(const uint32_t *slot, const uint32_t *prim_id, uint32_t n, uint32_t *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = slot[i] == SPT_MISS ? 0u : prim_id[slot[i]];
}
__global__ void k_slot_to_flag(const uint32_t *slot, uint32_t n, uint8_t *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = slot[i] == SPT_MISS ? 0 : 1;
}

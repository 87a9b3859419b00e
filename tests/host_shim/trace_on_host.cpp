// TEST INFRASTRUCTURE (not part of the product): the traversal KERNEL (csrc/trace_kernels.cuh: k_trace_multi, the pair-node
// walk) and the scene re-layout kernels (csrc/spt_build.cu) compiled for the host as a warp of one lane / a grid of one
// thread (fake/cuda_runtime.h), so that first-hit ids and distances of the very kernel source can be compared with the
// reference's golden vectors on a machine without a GPU.
#define SPT_HOST_SHIM 1
#include "trace_kernels.cuh"
#include "../../pbrt_v2_spectral_b200/csrc/spt_build.cu"
#include "../../pbrt_v2_spectral_b200/csrc/spt_wide.cu"
#include "camera.cuh"
#include "sampler.cuh"
#include <vector>

// K1 of csrc/spt_exact.cu is a __global__ in a .cu with launchers; its body is repeated here around the same device
// functions (wave_pixel, pixel_key, ld2, camera_ray) so that the sample -> film position -> camera ray chain runs on the host.
static void gen_camera_host(const RenderCfg &cfg, const SampleSource &src, float4 *ray_o, float4 *ray_d, float2 *img_xy) {
    for (uint32_t i = 0; i < cfg.n_samples; ++i) {
        int px, py;
        uint32_t s = i & ((uint32_t)cfg.spp - 1u);
        bool valid = wave_pixel(cfg, cfg.pixel_base + (i >> cfg.spp_shift), &px, &py);
        if (!valid) { ray_o[i] = make_float4(0.f, 0.f, 0.f, 1.f); ray_d[i] = make_float4(0.f, 0.f, 1.f, -1.f); img_xy[i] = make_float2(-1e30f, -1e30f); continue; }
        uint32_t pk = pixel_key(src.seed, pix_key(px, py));
        float t2[2];
        ld2(pk, 0, s, src.spp, t2);
        float ix = px + t2[0], iy = py + t2[1];
        ld2(pk, 1, s, src.spp, t2);
        Ray ray;
        camera_ray(cfg.cam, ix, iy, t2[0], t2[1], &ray);
        ray_o[i] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
        ray_d[i] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
        img_xy[i] = make_float2(ix, iy);
    }
}

extern "C" {

// rays: n x 8 {o, d, mint, maxt}; any = 0: closest hit (out_slot, out_t), 1: any hit (out_slot = SPT_MISS or a slot).
// Returns 0, or -1 when the tree does not pack into pair nodes (the product then walks the reference layout).
int hd_trace(const SptSceneDesc *d, const float *rays, uint32_t n, int any, uint32_t *out_slot, float *out_t, int wide) {
    // ---- scene re-layout with the product's kernels: leaf flags, exclusive scan of the interior flags, pair nodes, vertices
    std::vector<RefNodeD> nodes(d->n_nodes);
    memcpy(nodes.data(), d->bvh_nodes, (size_t)d->n_nodes * 32);
    std::vector<uint32_t> interior(d->n_nodes), pidx(d->n_nodes);
    uint32_t status[2] = { 0, 0xffffffffu };
    k_node_flags(nodes.data(), d->n_nodes, d->prim_kind, d->n_prims, interior.data(), status);
    uint32_t run = 0;
    for (uint32_t k = 0; k < d->n_nodes; ++k) { pidx[k] = run; run += interior[k]; }
    std::vector<float4> pn(((size_t)d->n_nodes / 2 + 1) * 4), tv((size_t)d->n_prims * 3);
    k_pair_nodes(nodes.data(), d->n_nodes, pidx.data(), pn.data(), status);
    k_tri_gather(d->prim_kind, d->prim_data, d->tri_vidx, d->P, d->n_prims, tv.data());
    if (status[0]) return -1;
    DevScene sc;
    memset(&sc, 0, sizeof(sc));
    sc.nodes = (const float4 *)nodes.data(); sc.pnodes = pn.data(); sc.root_code = status[1]; sc.tri_verts = tv.data();
    sc.n_nodes = d->n_nodes; sc.n_prims = d->n_prims;
    sc.prim_kind = d->prim_kind; sc.prim_flags = d->prim_flags; sc.prim_id = d->prim_id; sc.prim_data = d->prim_data;
    sc.prim_material = d->prim_material; sc.prim_light = d->prim_light; sc.prim_xform = d->prim_xform;
    sc.tri_vidx = d->tri_vidx; sc.P = d->P; sc.N = d->N; sc.UV = d->UV; sc.quadrics = d->quadrics; sc.xforms = d->xforms;
    // wide != 0: the fast layout (csrc/wide.h), collapsed by the product's builder from the flagged reference nodes
    std::vector<W4Node> wnodes;
    if (wide) {
        if (!spt_build_w4(nodes.data(), d->n_nodes, &wnodes, &sc.wroot)) return -2;
        wnodes.push_back(W4Node());
        sc.wnodes = (const float4 *)wnodes.data();
    }
    // ---- the kernel itself, one lane
    std::vector<float4> ro(n), rd(n);
    for (uint32_t i = 0; i < n; ++i) {
        const float *r = rays + 8 * (size_t)i;
        ro[i] = make_float4(r[0], r[1], r[2], r[6]); rd[i] = make_float4(r[3], r[4], r[5], r[7]);
    }
    // one launch over two queues, as the wavefront issues it: the first half of the rays, then the second half through an index queue
    uint32_t n0 = n / 2, n1 = n - n0, work = 0;
    std::vector<uint32_t> q1(n1);
    for (uint32_t k = 0; k < n1; ++k) q1[k] = n0 + k;
    TraceMultiArgs a;
    memset(&a, 0, sizeof(a));
    a.nseg = 2; a.work = &work; a.fetch_threshold = 14;
    a.seg[0].queue = nullptr; a.seg[0].count = &n0; a.seg[1].queue = q1.data(); a.seg[1].count = &n1;
    for (int k = 0; k < 2; ++k) { a.seg[k].ro = ro.data(); a.seg[k].rd = rd.data(); a.seg[k].out_slot = out_slot; a.seg[k].out_t = out_t; a.seg[k].any = any ? 1u : 0u; }
    if (wide) k_trace_multi<false, true>(sc, a); else k_trace_multi<false, false>(sc, a);
    return 0;
}

// One 8x8-pixel tile at (x0, y0) of the product's generated samples: film positions and camera rays per slot
// (n = 64 * spp slots; out_xy n x 2, out_rays n x 8), and for slot order see wave_pixel (csrc/wave.cuh).
void hd_gen_tile(const SptCameraDesc *cam, uint64_t seed64, int x0, int y0, int spp, float *out_xy, float *out_rays) {
    RenderCfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.cam = *cam; cfg.spp = spp;
    for (cfg.spp_shift = 0; (1 << cfg.spp_shift) < spp; ++cfg.spp_shift) {}
    cfg.x0 = x0; cfg.y0 = y0; cfg.x1 = x0 + 8; cfg.y1 = y0 + 8; cfg.tile = 8; cfg.tile_shift = 3; cfg.tilesX = 1; cfg.tilesY = 1;
    cfg.rank = 0; cfg.nranks = 1; cfg.seed = (uint32_t)(seed64 ^ (seed64 >> 32)); cfg.pixel_base = 0; cfg.n_samples = 64u * (uint32_t)spp;
    cfg.sub = 1;
    SampleSource src; src.smp = nullptr; src.stride = 0; src.rng = nullptr; src.n_rng = 0; src.seed = cfg.seed; src.spp = (uint32_t)spp;
    std::vector<float4> ro(cfg.n_samples), rd(cfg.n_samples);
    std::vector<float2> xy(cfg.n_samples);
    gen_camera_host(cfg, src, ro.data(), rd.data(), xy.data());
    for (uint32_t i = 0; i < cfg.n_samples; ++i) {
        out_xy[2 * i] = xy[i].x; out_xy[2 * i + 1] = xy[i].y;
        float *r = out_rays + 8 * (size_t)i;
        r[0] = ro[i].x; r[1] = ro[i].y; r[2] = ro[i].z; r[3] = rd[i].x; r[4] = rd[i].y; r[5] = rd[i].z; r[6] = ro[i].w; r[7] = rd[i].w;
    }
}
// The ten values + Russian-roulette draw bounce b of sample s of pixel (px, py) consumes (bounce_dims, csrc/sampler.cuh)
void hd_bounce_dims(uint64_t seed64, int px, int py, int s, int spp, int b, int have_lights, float *u11) {
    SampleSource src; src.smp = nullptr; src.stride = 0; src.rng = nullptr; src.n_rng = 0;
    src.seed = (uint32_t)(seed64 ^ (seed64 >> 32)); src.spp = (uint32_t)spp;
    uint32_t pk = pixel_key(src.seed, pix_key(px, py));
    bounce_dims(src, 0, pk, (uint32_t)s, b, have_lights != 0, u11, u11 + 10);
}

}  // extern "C"

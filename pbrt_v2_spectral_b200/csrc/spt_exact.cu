// spt_exact.cu — the translation unit whose results must match the reference bit for bit: camera
// rays (K1) and BVH traversal (K2/K3). Compiled -fmad=false: the reference is built without FMA
// contraction (src/Makefile:24-29), Cross() is FP64 (geometry.h:479-486, F7), IEEE
// division and square root are nvcc defaults.
#include "launch.h"
#include "camera.cuh"
#include "sampler.cuh"
#include "trace_kernels.cuh"

// ---- K1 ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gen_camera(RenderCfg cfg, SampleSource src, WaveBuffers wb, uint32_t *count_out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cfg.n_samples; i += gridDim.x * blockDim.x) {
        float ix, iy, lu, lv;
        bool valid = true;
        if (src.smp) {
            const float *s = src.smp + (size_t)src.stride * i;
            ix = s[0]; iy = s[1]; lu = s[2]; lv = s[3];
        } else {
            int px, py;
            uint32_t s = i & ((uint32_t)cfg.spp - 1u);                       // spp is a power of two (LDSampler rounds up)
            valid = wave_pixel(cfg, cfg.pixel_base + (i >> cfg.spp_shift), &px, &py);
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
            // a sample slot outside the sample extent (tiles overhang the image): a ray with an empty
            // parameter range, which no box or primitive test accepts, and a film position K7 skips
            wb.ray_o[i] = make_float4(0.f, 0.f, 0.f, 1.f);
            wb.ray_d[i] = make_float4(0.f, 0.f, 1.f, -1.f);
            wb.img_xy[i] = make_float2(-1e30f, -1e30f);
        }
    }
    // bounce 0 walks the wave in sample order: no queue (queue == NULL means identity), one counter write
    // instead of a million same-address atomics
    // (word 11 of the row: the next free node slot of a specular tree under directlighting, RenderCfg::tree)
    if (blockIdx.x == 0 && threadIdx.x == 0) { *count_out = cfg.n_samples; count_out[11] = cfg.n_samples; }
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

// ---- launchers -------------------------------------------------------------------------------------
static inline unsigned grid256(uint64_t n) { uint64_t g = (n + 255) / 256; return (unsigned)(g < 1 ? 1 : (g > 65535 ? 65535 : g)); }
void spt_launch_gen_camera(int grid, cudaStream_t st, const RenderCfg &cfg, const SampleSource &src, const WaveBuffers &wb,
                           uint32_t *count_out) {
    k_gen_camera<<<grid, 256, 0, st>>>(cfg, src, wb, count_out);
}
static void launch_multi(bool count, int grid, cudaStream_t st, const DevScene &sc, const TraceMultiArgs &a) {
    if (sc.wnodes) {        // fast mode (wide.h): the scene opted in with spt_scene_set_traversal
        if (count) k_trace_multi<true, true><<<grid, 128, 0, st>>>(sc, a); else k_trace_multi<false, true><<<grid, 128, 0, st>>>(sc, a);
    } else {
        if (count) k_trace_multi<true, false><<<grid, 128, 0, st>>>(sc, a); else k_trace_multi<false, false><<<grid, 128, 0, st>>>(sc, a);
    }
}
void spt_launch_trace_multi(int variant, bool merge, bool count, int grid, cudaStream_t st, const DevScene &sc, const TraceMultiArgs &a) {
    if (variant != 0 && merge) { launch_multi(count, grid, st, sc, a); return; }
    for (uint32_t k = 0; k < a.nseg; ++k) {
        if (variant != 0) {
            TraceMultiArgs one = a;
            one.seg[0] = a.seg[k]; one.nseg = 1; one.work = a.work + k;
            launch_multi(count, grid, st, sc, one);
        } else {
            TraceArgs t; t.queue = a.seg[k].queue; t.count = a.seg[k].count; t.work = a.work + k; t.ro = a.seg[k].ro; t.rd = a.seg[k].rd;
            t.out_slot = a.seg[k].out_slot; t.out_t = a.seg[k].out_t; t.fetch_threshold = a.fetch_threshold;
            if (a.seg[k].any) { if (count) k_trace_v0<true, true><<<grid, 128, 0, st>>>(sc, t); else k_trace_v0<true, false><<<grid, 128, 0, st>>>(sc, t); }
            else { if (count) k_trace_v0<false, true><<<grid, 128, 0, st>>>(sc, t); else k_trace_v0<false, false><<<grid, 128, 0, st>>>(sc, t); }
        }
    }
}
void spt_launch_camera_rays(cudaStream_t st, const SptCameraDesc &cam, const float *samples, uint32_t n, float *out) {
    k_camera_rays<<<grid256(n), 256, 0, st>>>(cam, samples, n, out);
}
void spt_launch_split_rays(cudaStream_t st, const float *rays, uint32_t n, float4 *ro, float4 *rd) {
    k_split_rays<<<grid256(n), 256, 0, st>>>(rays, n, ro, rd);
}
void spt_launch_slot_to_id(cudaStream_t st, const uint32_t *slot, const uint32_t *prim_id, uint32_t n, uint32_t *out) {
    k_slot_to_id<<<grid256(n), 256, 0, st>>>(slot, prim_id, n, out);
}
void spt_launch_slot_to_flag(cudaStream_t st, const uint32_t *slot, uint32_t n, uint8_t *out) {
    k_slot_to_flag<<<grid256(n), 256, 0, st>>>(slot, n, out);
}

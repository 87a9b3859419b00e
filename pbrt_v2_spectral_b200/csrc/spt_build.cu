// spt_build.cu — scene re-layout on the device (spt_scene_create): what the traversal kernels read is derived from the
// reference's arrays AFTER they are in HBM, so the host does not spend 1.7 ms per scene (131 k nodes, 66 k primitives)
// building it: leaf flags in the reference nodes, the pair-node array (one 64-byte node per interior node: both children's
// bounds + child codes, see spt_device.cuh) and the triangle vertices pre-gathered per BVH slot.
#ifndef SPT_HOST_SHIM                     // tests/host_shim compiles the kernels below for the host, without CUB and the launcher
#include <cub/device/device_scan.cuh>
#endif
#include "launch.h"

struct RefNodeD { float b[6]; uint32_t off; uint32_t meta; };     // LinearBVHNode: meta = nPrims | axis << 8 | pad << 16

// leaves: hasQuadric flag into the reference's pad byte (bit 16), pad cleared; interior[n] = 1 for interior nodes
__global__ void k_node_flags(RefNodeD *nodes, uint32_t n_nodes, const uint8_t *prim_kind, uint32_t n_prims, uint32_t *interior, uint32_t *status) {
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += gridDim.x * blockDim.x) {
        uint32_t meta = nodes[n].meta, np = meta & 0xffu, off = nodes[n].off;
        uint32_t hasq = 0;
        for (uint32_t i = 0; i < np; ++i)
            if (off + i < n_prims && prim_kind[off + i] != SPT_PRIM_TRIANGLE) hasq = 1;
        nodes[n].meta = (meta & 0xffffu) | hasq << 16;
        interior[n] = np == 0 ? 1u : 0u;
        if (np > 8) atomicOr(status, 2u);                          // does not pack into a child code: traverse the reference layout
    }
}
__device__ __forceinline__ uint32_t child_code(const RefNodeD *nodes, const uint32_t *pidx, uint32_t c) {
    uint32_t meta = nodes[c].meta, np = meta & 0xffu;
    return np ? (0x80000000u | (((meta >> 16) & 1u) ? 0x40000000u : 0u) | ((np - 1u) & 7u) << 27 | nodes[c].off) : pidx[c];
}
__global__ void k_pair_nodes(const RefNodeD *nodes, uint32_t n_nodes, const uint32_t *pidx, float4 *pn, uint32_t *status) {
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < n_nodes; n += gridDim.x * blockDim.x) {
        if (n == 0) status[1] = child_code(nodes, pidx, 0);
        if (nodes[n].meta & 0xffu) continue;
        uint32_t c0 = n + 1, c1 = nodes[n].off;
        if (c0 >= n_nodes || c1 >= n_nodes) { atomicOr(status, 1u); continue; }
        float4 *q = pn + (size_t)pidx[n] * 4;
        const float *a = nodes[c0].b, *b = nodes[c1].b;
        q[0] = make_float4(a[0], a[1], a[2], a[3]);
        q[1] = make_float4(a[4], a[5], b[0], b[1]);
        q[2] = make_float4(b[2], b[3], b[4], b[5]);
        q[3] = make_float4(__uint_as_float(child_code(nodes, pidx, c0)), __uint_as_float(child_code(nodes, pidx, c1)),
                           __uint_as_float((nodes[n].meta >> 8) & 3u), 0.f);
    }
}
__global__ void k_tri_gather(const uint8_t *prim_kind, const uint32_t *prim_data, const int32_t *tri_vidx, const float *P,
                             uint32_t n_prims, float4 *tv) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_prims; p += gridDim.x * blockDim.x) {
        float4 v[3] = { make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0) };
        if (prim_kind[p] == SPT_PRIM_TRIANGLE) {
            const int32_t *vi = tri_vidx + 3 * (size_t)prim_data[p];
            for (int k = 0; k < 3; ++k) { const float *q = P + 3 * (size_t)vi[k]; v[k] = make_float4(q[0], q[1], q[2], 0.f); }
        }
        tv[(size_t)p * 3] = v[0]; tv[(size_t)p * 3 + 1] = v[1]; tv[(size_t)p * 3 + 2] = v[2];
    }
}

#ifndef SPT_HOST_SHIM
size_t spt_relayout_scratch_bytes(uint32_t n_nodes) {
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_nodes);
    return ((tmp + 255) & ~(size_t)255) + 2 * (((size_t)n_nodes * 4 + 255) & ~(size_t)255) + 256;
}
// nodes: the uploaded reference nodes (modified in place); pn: n_nodes/2+1 pair nodes; tv: 3 float4 per primitive;
// scratch: spt_relayout_scratch_bytes(n_nodes); status_host[0]: bit 0 malformed tree, bit 1 a leaf of more than 8 primitives;
// status_host[1]: child code of the root. Synchronises the stream.
cudaError_t spt_launch_relayout(cudaStream_t st, void *nodes, uint32_t n_nodes, const uint8_t *prim_kind, const uint32_t *prim_data,
                                uint32_t n_prims, const int32_t *tri_vidx, const float *P, float4 *pn, float4 *tv, void *scratch,
                                uint32_t status_host[2]) {
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_nodes);
    const size_t arr = ((size_t)n_nodes * 4 + 255) & ~(size_t)255;
    uint8_t *base = (uint8_t *)scratch;
    uint32_t *status = (uint32_t *)base;
    uint32_t *interior = (uint32_t *)(base + 256), *pidx = (uint32_t *)(base + 256 + arr);
    void *cubtmp = base + 256 + 2 * arr;
    cudaMemsetAsync(status, 0, 8, st);
    const unsigned gn = (unsigned)((n_nodes + 255) / 256 ? (n_nodes + 255) / 256 : 1), gp = (unsigned)((n_prims + 255) / 256 ? (n_prims + 255) / 256 : 1);
    if (n_nodes) {
        k_node_flags<<<gn, 256, 0, st>>>((RefNodeD *)nodes, n_nodes, prim_kind, n_prims, interior, status);
        cub::DeviceScan::ExclusiveSum(cubtmp, tmp, interior, pidx, (int)n_nodes, st);
        k_pair_nodes<<<gn, 256, 0, st>>>((const RefNodeD *)nodes, n_nodes, pidx, pn, status);
    }
    if (n_prims) k_tri_gather<<<gp, 256, 0, st>>>(prim_kind, prim_data, tri_vidx, P, n_prims, tv);
    cudaError_t e = cudaMemcpyAsync(status_host, status, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e;
}
#endif

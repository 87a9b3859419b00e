// TEST INFRASTRUCTURE (not part of the product): the traversal KERNEL (csrc/trace_kernels.cuh: k_trace_v1, the pair-node
// walk) and the scene re-layout kernels (csrc/spt_build.cu) compiled for the host as a warp of one lane / a grid of one
// thread (fake/cuda_runtime.h), so that first-hit ids and distances of the very kernel source can be compared with the
// reference's golden vectors on a machine without a GPU.
#define SPT_HOST_SHIM 1
#include "trace_kernels.cuh"
#include "../../pbrt_v2_spectral_b200/csrc/spt_build.cu"
#include <vector>

extern "C" {

// rays: n x 8 {o, d, mint, maxt}; any = 0: closest hit (out_slot, out_t), 1: any hit (out_slot = SPT_MISS or a slot).
// Returns 0, or -1 when the tree does not pack into pair nodes (the product then walks the reference layout).
int hd_trace(const SptSceneDesc *d, const float *rays, uint32_t n, int any, uint32_t *out_slot, float *out_t) {
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
    // ---- the kernel itself, one lane
    std::vector<float4> ro(n), rd(n);
    for (uint32_t i = 0; i < n; ++i) {
        const float *r = rays + 8 * (size_t)i;
        ro[i] = make_float4(r[0], r[1], r[2], r[6]); rd[i] = make_float4(r[3], r[4], r[5], r[7]);
    }
    uint32_t count = n, work = 0;
    TraceArgs a;
    a.queue = nullptr; a.count = &count; a.work = &work; a.ro = ro.data(); a.rd = rd.data();
    a.out_slot = out_slot; a.out_t = out_t; a.fetch_threshold = 14;
    if (any) k_trace_v1<true, false>(sc, a); else k_trace_v1<false, false>(sc, a);
    return 0;
}

}  // extern "C"

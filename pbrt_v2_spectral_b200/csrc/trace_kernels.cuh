// trace_kernels.cuh — K2 / K3: BVHAccel::Intersect / IntersectP (src/accelerators/bvh.cpp:380-432,
// :435-481) as persistent-warp kernels. Every lane owns one ray; idle lanes refill from the
// device-side queue with one warp-aggregated atomic once fewer than fetch_threshold lanes of the
// warp are still traversing (incoherent bounce / MIS rays otherwise leave ~8 of 32 lanes busy).
//
// Both variants visit the reference's nodes in the reference's order (near child first, far child
// from the 64-entry todo stack, leaves in slot order, `t <= maxt` acceptance), with the reference's
// slab arithmetic (bvh.cpp:118-140) and triangle test (trianglemesh.cpp:119-273): the same
// primitives are tested in the same order against the same maxt, so ids and t match bit for bit
// and ties resolve to the same primitive. They differ only in how the 32 lanes of
// a warp are kept converged.
#pragma once
#include "traverse.cuh"

#define PN_NONE 0xffffffffu

#include "trace_args.h"

// per-lane ray state shared by the variants
struct LaneRay {
    Ray ray; v3 invDir; bool negx, negy, negz;
    bool exact;                 // a component of invDir is inf/NaN: 0 * inf can appear in the slab test, take the literal code
    uint32_t i, best;
};

// Refill: idle lanes claim the next queue entries. Returns false for lanes that got no ray.
__device__ __forceinline__ bool lane_fetch(const TraceArgs &a, uint32_t n, int lane, bool active, bool &exhausted, LaneRay &L) {
    const unsigned FULL = 0xffffffffu;
    unsigned idle = __ballot_sync(FULL, !active);
    bool got = false;
    if (!exhausted && idle) {
        uint32_t base = 0;
        int leader = __ffs(idle) - 1;
        if (lane == leader) base = atomicAdd(a.work, (uint32_t)__popc(idle));
        base = __shfl_sync(FULL, base, leader);
        if (!active) {
            uint32_t q = base + __popc(idle & ((1u << lane) - 1));
            if (q < n) {
                L.i = a.queue ? a.queue[q] : q;
                float4 o = a.ro[L.i], d = a.rd[L.i];
                L.ray.o = V(o.x, o.y, o.z); L.ray.d = V(d.x, d.y, d.z); L.ray.mint = o.w; L.ray.maxt = d.w;
                L.invDir = V(1.f / L.ray.d.x, 1.f / L.ray.d.y, 1.f / L.ray.d.z);
                L.negx = L.invDir.x < 0; L.negy = L.invDir.y < 0; L.negz = L.invDir.z < 0;
                L.exact = !(isfinite(L.invDir.x) && isfinite(L.invDir.y) && isfinite(L.invDir.z));
                L.best = SPT_MISS;
                got = true;
            }
        }
        if (base + (uint32_t)__popc(idle) >= n) exhausted = true;
    }
    return got;
}

// The primitives of one leaf, in slot order. Returns true when an any-hit ray is finished.
template <bool ANY, bool COUNT>
__device__ __forceinline__ bool leaf_test(const DevScene &sc, uint32_t offset, uint32_t nPrims, bool hasQuadric, LaneRay &L,
                                          unsigned long long &cp) {
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
            h = tri_test(V(a.x, a.y, a.z), V(b.x, b.y, b.z), V(c.x, c.y, c.z), L.ray, &t, &b1, &b2);
        } else if (sc.prim_kind[s] == SPT_PRIM_SPHERE) {
            h = sphere_intersect(sc, sc.quadrics[sc.prim_data[s]], 0, L.ray, &t, nullptr);
        } else {
            h = disk_intersect(sc, sc.quadrics[sc.prim_data[s]], 0, L.ray, &t, nullptr);
        }
        if (h) {
            L.best = s;
            if (ANY) return true;              // IntersectP returns at the first accepted primitive
            L.ray.maxt = t;
        }
    }
    return false;
}

// ---- variant 0: one loop, leaf tested where it is found ---------------------------------------
template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(128) k_trace_v0(DevScene sc, TraceArgs a) {
    const uint32_t n = *a.count;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    uint32_t todo[64];
    uint32_t todoOffset = 0, nodeNum = 0;
    LaneRay L; L.ray.o = V(0, 0, 0); L.ray.d = V(0, 0, 1); L.ray.mint = 0.f; L.ray.maxt = 0.f; L.invDir = V(0, 0, 0);
    L.negx = L.negy = L.negz = false; L.i = 0; L.best = SPT_MISS;
    bool active = false, exhausted = (n == 0);
    unsigned long long cn = 0, cp = 0;
    for (;;) {
        if (lane_fetch(a, n, lane, active, exhausted, L)) {
            todoOffset = 0; nodeNum = 0; active = true;
            if (sc.n_nodes == 0) { a.out_slot[L.i] = SPT_MISS; if (!ANY) a.out_t[L.i] = L.ray.maxt; active = false; }
        }
        if (!__any_sync(FULL, active)) break;
        while (active) {
            float4 n0 = __ldg(&sc.nodes[2 * (size_t)nodeNum]);
            float4 n1 = __ldg(&sc.nodes[2 * (size_t)nodeNum + 1]);
            if (COUNT) ++cn;
            bool pop = true, finished = false;
            if (slab(n0, n1, L.ray, L.invDir, L.negx, L.negy, L.negz)) {
                uint32_t meta = __float_as_uint(n1.w);
                uint32_t offset = __float_as_uint(n1.z);
                uint32_t nPrims = meta & 0xff;
                if (nPrims > 0) {
                    finished = leaf_test<ANY, COUNT>(sc, offset, nPrims, (meta >> 16) & 1, L, cp);
                } else {
                    uint32_t axis = (meta >> 8) & 0xff;
                    bool neg = axis == 0 ? L.negx : (axis == 1 ? L.negy : L.negz);
                    if (neg) { todo[todoOffset++] = nodeNum + 1; nodeNum = offset; }
                    else { todo[todoOffset++] = offset; nodeNum = nodeNum + 1; }
                    pop = false;
                }
            }
            if (pop) {
                if (todoOffset == 0 || finished) {
                    a.out_slot[L.i] = L.best;
                    if (!ANY) a.out_t[L.i] = L.ray.maxt;
                    active = false;
                } else nodeNum = todo[--todoOffset];
            }
            if (active && !exhausted && (uint32_t)__popc(__activemask()) < a.fetch_threshold) break;
        }
    }
    if (COUNT && sc.counters && (cn | cp)) {
        atomicAdd(&sc.counters[ANY ? 2 : 0], cn);
        atomicAdd(&sc.counters[ANY ? 3 : 1], cp);
    }
}

// slab test of one child box with the reference's literal sequence (bvh.cpp:118-140); returns hit and
// the clipped tmin (valid when the boxes overlap). x0 = the bound selected by dirIsNeg, x1 the other.
__device__ __forceinline__ bool slab6(float x0, float x1, float y0, float y1, float z0, float z1, const Ray &ray,
                                      v3 invDir, float *tminOut) {
    float tmin = (x0 - ray.o.x) * invDir.x;
    float tmax = (x1 - ray.o.x) * invDir.x;
    float tymin = (y0 - ray.o.y) * invDir.y;
    float tymax = (y1 - ray.o.y) * invDir.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (z0 - ray.o.z) * invDir.z;
    float tzmax = (z1 - ray.o.z) * invDir.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    *tminOut = tmin;
    return (tmin < ray.maxt) && (tmax > ray.mint);
}
// The same decision without the sign selects and early outs, for rays whose invDir is finite (no
// 0 * inf = NaN can occur): per axis near = min(t1, t2), far = max(t1, t2) are exactly the
// reference's bounds[dirIsNeg] / bounds[1-dirIsNeg] products (rounding is monotonic), its four
// early-out comparisons hold iff max(near) <= min(far), and its running tmin / tmax are those
// max / min. x0..z0 = pMin, x1..z1 = pMax.
__device__ __forceinline__ bool slab6_finite(float x0, float x1, float y0, float y1, float z0, float z1, const Ray &ray,
                                             v3 invDir, float *tminOut) {
    float tx0 = (x0 - ray.o.x) * invDir.x, tx1 = (x1 - ray.o.x) * invDir.x;
    float ty0 = (y0 - ray.o.y) * invDir.y, ty1 = (y1 - ray.o.y) * invDir.y;
    float tz0 = (z0 - ray.o.z) * invDir.z, tz1 = (z1 - ray.o.z) * invDir.z;
    float tmin = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fminf(tz0, tz1));
    float tmax = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fmaxf(tz0, tz1));
    *tminOut = tmin;
    return (tmin <= tmax) && (tmin < ray.maxt) && (tmax > ray.mint);
}

// ---- variant 1: pair nodes ------------------------------------------------------------------------
// One 64-byte record per INTERIOR node holds the bounds of both children (built in
// spt_scene_create): one step fetches four 16-byte words, slab-tests both children and descends.
// A child is one 32-bit code: a pair index, or PN_LEAF | hasQuadric << 30 | (nPrims-1) << 27 | first slot.
//
// Closest hit: the reference pushes the far child unconditionally and slab-tests it when popped,
// against the maxt current THEN; of that test only `tmin < maxt` depends on maxt, so the far child
// is tested at push time, dropped if it already fails, and its tmin is kept on the stack and
// re-compared with the (possibly shrunk) maxt at pop time - the same decision, bit for bit.
// Any hit: maxt never changes during IntersectP, the answer is the OR over the same set of leaves
// whatever their order, so the stack holds bare codes.
#define PN_LEAF 0x80000000u
template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(128) k_trace_v1(DevScene sc, TraceArgs a) {
    const uint32_t n = *a.count;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    uint2 stk[ANY ? 1 : 64];         // closest: {child code, tmin of the child's slab test at push time}
    uint32_t stk1[ANY ? 64 : 1];     // any hit: child code
    uint32_t sp = 0, cur = PN_NONE, negMask = 0;
    LaneRay L; L.ray.o = V(0, 0, 0); L.ray.d = V(0, 0, 1); L.ray.mint = 0.f; L.ray.maxt = 0.f; L.invDir = V(0, 0, 0);
    L.negx = L.negy = L.negz = false; L.exact = false; L.i = 0; L.best = SPT_MISS;
    bool active = false, exhausted = (n == 0);
    unsigned long long cn = 0, cp = 0;
    for (;;) {
        if (lane_fetch(a, n, lane, active, exhausted, L)) {
            sp = 0; cur = PN_NONE; active = true;
            negMask = (L.negx ? 1u : 0u) | (L.negy ? 2u : 0u) | (L.negz ? 4u : 0u);
            // the root: the one node whose own box is tested from the reference array
            if (sc.n_nodes) {
                float4 n0 = __ldg(&sc.nodes[0]), n1 = __ldg(&sc.nodes[1]);
                if (COUNT) ++cn;
                if (slab(n0, n1, L.ray, L.invDir, L.negx, L.negy, L.negz)) cur = sc.root_code;
            }
        }
        if (!__any_sync(FULL, active)) break;
        while (active) {
            bool done = false;
            if (cur == PN_NONE) {                                      // pop
                if (ANY) {
                    if (sp == 0) done = true; else cur = stk1[--sp];
                } else {
                    for (;;) {
                        if (sp == 0) { done = true; break; }
                        uint2 e = stk[--sp];
                        // the far child's slab test completes here with the current maxt
                        if (__uint_as_float(e.y) < L.ray.maxt) { cur = e.x; break; }
                    }
                }
            }
            if (cur < PN_LEAF) {                                        // interior: one pair step
                const float4 *pn = sc.pnodes + 4 * (size_t)cur;
                float4 q0 = __ldg(pn), q1 = __ldg(pn + 1), q2 = __ldg(pn + 2), q3 = __ldg(pn + 3);
                if (COUNT) cn += 2;
                // child 0: min (q0.x,q0.y,q0.z) max (q0.w,q1.x,q1.y); child 1: min (q1.z,q1.w,q2.x) max (q2.y,q2.z,q2.w)
                float t0, t1;
                bool h0, h1;
                if (!L.exact) {
                    h0 = slab6_finite(q0.x, q0.w, q0.y, q1.x, q0.z, q1.y, L.ray, L.invDir, &t0);
                    h1 = slab6_finite(q1.z, q2.y, q1.w, q2.z, q2.x, q2.w, L.ray, L.invDir, &t1);
                } else {
                    h0 = slab6(L.negx ? q0.w : q0.x, L.negx ? q0.x : q0.w, L.negy ? q1.x : q0.y, L.negy ? q0.y : q1.x,
                               L.negz ? q1.y : q0.z, L.negz ? q0.z : q1.y, L.ray, L.invDir, &t0);
                    h1 = slab6(L.negx ? q2.y : q1.z, L.negx ? q1.z : q2.y, L.negy ? q2.z : q1.w, L.negy ? q1.w : q2.z,
                               L.negz ? q2.w : q2.x, L.negz ? q2.x : q2.w, L.ray, L.invDir, &t1);
                }
                const uint32_t c0 = __float_as_uint(q3.x), c1 = __float_as_uint(q3.y), axis = __float_as_uint(q3.z);
                // reference: dirIsNeg[axis] ? second child first : first child first
                const bool swap = (negMask >> axis) & 1u;
                const uint32_t nearC = swap ? c1 : c0, farC = swap ? c0 : c1;
                const bool nearHit = swap ? h1 : h0, farHit = swap ? h0 : h1;
                if (farHit) {
                    if (ANY) stk1[sp++] = farC;
                    else stk[sp++] = make_uint2(farC, __float_as_uint(swap ? t0 : t1));
                }
                cur = nearHit ? nearC : PN_NONE;
            }
            if (cur >= PN_LEAF && cur != PN_NONE) {                     // leaf: its primitives in slot order
                if (leaf_test<ANY, COUNT>(sc, cur & 0x07ffffffu, ((cur >> 27) & 7u) + 1u, (cur >> 30) & 1u, L, cp)) done = true;
                cur = PN_NONE;
            }
            if (done) {
                a.out_slot[L.i] = L.best;
                if (!ANY) a.out_t[L.i] = L.ray.maxt;
                active = false;
            }
            if (active && !exhausted && (uint32_t)__popc(__activemask()) < a.fetch_threshold) break;
        }
    }
    if (COUNT && sc.counters && (cn | cp)) {
        atomicAdd(&sc.counters[ANY ? 2 : 0], cn);
        atomicAdd(&sc.counters[ANY ? 3 : 1], cp);
    }
}

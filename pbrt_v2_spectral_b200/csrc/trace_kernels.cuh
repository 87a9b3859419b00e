// trace_kernels.cuh — K2 / K3: BVHAccel::Intersect / IntersectP (src/accelerators/bvh.cpp:380-432,
// :435-481) as persistent-warp kernels. Every lane owns one ray; idle lanes refill from the
// device-side queue with one warp-aggregated atomic once fewer than FETCH_THRESHOLD lanes of the
// warp are still traversing (incoherent bounce / MIS rays otherwise leave ~8 of 32 lanes busy).
//
// All variants visit the reference's nodes in the reference's order (near child first, far child
// from the 64-entry todo stack, leaves in slot order, `t <= maxt` acceptance), with the reference's
// slab arithmetic (bvh.cpp:118-140) and triangle test (trianglemesh.cpp:119-273): the same
// primitives are tested in the same order against the same maxt, so ids and t match bit for bit
// and ties resolve to the same primitive. They differ only in how the 32 lanes of
// a warp are kept converged.
#pragma once
#include "traverse.cuh"

#define FETCH_THRESHOLD 20
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
            if (active && !exhausted && __popc(__activemask()) < FETCH_THRESHOLD) break;
        }
    }
    if (COUNT && sc.counters && (cn | cp)) {
        atomicAdd(&sc.counters[ANY ? 2 : 0], cn);
        atomicAdd(&sc.counters[ANY ? 3 : 1], cp);
    }
}

// ---- variant 2: node phase / leaf phase ----------------------------------------------------------
// With one loop, an iteration in which ANY lane sits on a leaf runs the ~150-instruction primitive
// test for the whole warp while ~92 % of the lanes only wanted a ~60-instruction node step. Here
// the warp alternates between two converged phases: lanes step through nodes until they hold a
// leaf (then they wait), and once at most `leaf_wait` lanes are still searching, every lane that
// holds a leaf tests its primitives. A lane stops searching as soon as it holds a leaf, so every
// primitive test sees the maxt the reference would have at that visit.
template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(128) k_trace_v2(DevScene sc, TraceArgs a) {
    const uint32_t n = *a.count;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    uint32_t todo[64];
    uint32_t sp = 0, cur = PN_NONE, leafOff = 0, leafMeta = 0;      // leafMeta & 0xff = pending primitives
    LaneRay L; L.ray.o = V(0, 0, 0); L.ray.d = V(0, 0, 1); L.ray.mint = 0.f; L.ray.maxt = 0.f; L.invDir = V(0, 0, 0);
    L.negx = L.negy = L.negz = false; L.i = 0; L.best = SPT_MISS;
    bool active = false, exhausted = (n == 0);
    unsigned long long cn = 0, cp = 0;
    for (;;) {
        if (lane_fetch(a, n, lane, active, exhausted, L)) {
            sp = 0; cur = 0; leafMeta = 0; active = true;
            if (sc.n_nodes == 0) { a.out_slot[L.i] = SPT_MISS; if (!ANY) a.out_t[L.i] = L.ray.maxt; active = false; }
        }
        unsigned act = __ballot_sync(FULL, active);
        if (!act) break;
        for (;;) {
            // ---- node phase
            for (;;) {
                bool searching = active && (leafMeta & 0xff) == 0;
                if (searching) {
                    if (cur == PN_NONE) {
                        if (sp == 0) {
                            a.out_slot[L.i] = L.best;
                            if (!ANY) a.out_t[L.i] = L.ray.maxt;
                            active = false; searching = false;
                        } else cur = todo[--sp];
                    }
                    if (searching) {
                        float4 n0 = __ldg(&sc.nodes[2 * (size_t)cur]);
                        float4 n1 = __ldg(&sc.nodes[2 * (size_t)cur + 1]);
                        if (COUNT) ++cn;
                        if (slab(n0, n1, L.ray, L.invDir, L.negx, L.negy, L.negz)) {
                            uint32_t meta = __float_as_uint(n1.w);
                            uint32_t offset = __float_as_uint(n1.z);
                            if (meta & 0xff) { leafOff = offset; leafMeta = meta; cur = PN_NONE; searching = false; }
                            else {
                                uint32_t axis = (meta >> 8) & 0xff;
                                bool neg = axis == 0 ? L.negx : (axis == 1 ? L.negy : L.negz);
                                if (neg) { todo[sp++] = cur + 1; cur = offset; }
                                else { todo[sp++] = offset; cur = cur + 1; }
                            }
                        } else cur = PN_NONE;
                    }
                }
                unsigned still = __ballot_sync(FULL, searching);
                if (!still) break;
                if ((uint32_t)__popc(still) <= a.leaf_wait && __any_sync(FULL, active && (leafMeta & 0xff))) break;
            }
            // ---- leaf phase
            if (active && (leafMeta & 0xff)) {
                bool finished = leaf_test<ANY, COUNT>(sc, leafOff, leafMeta & 0xff, (leafMeta >> 16) & 1, L, cp);
                leafMeta = 0;
                if (finished) {
                    a.out_slot[L.i] = L.best;
                    if (!ANY) a.out_t[L.i] = L.ray.maxt;
                    active = false;
                }
            }
            unsigned now = __ballot_sync(FULL, active);
            if (!now) break;
            if (!exhausted && __popc(now) < FETCH_THRESHOLD) break;
        }
    }
    if (COUNT && sc.counters && (cn | cp)) {
        atomicAdd(&sc.counters[ANY ? 2 : 0], cn);
        atomicAdd(&sc.counters[ANY ? 3 : 1], cp);
    }
}

// slab test of one child box; returns hit and the clipped tmin (valid when the boxes overlap)
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

// ---- variant 1: pair nodes ------------------------------------------------------------------------
// One 64-byte record per INTERIOR node holds the bounds of both children (built in
// spt_scene_create): one step fetches four 16-byte words, slab-tests both children and descends.
// The reference pushes the far child unconditionally and slab-tests it when popped, against the
// maxt current THEN; of that test only `tmin < maxt` depends on maxt, so the far child is tested
// at push time, dropped if it already fails, and its tmin is kept on the stack and re-compared with
// the (possibly shrunk) maxt at pop time - the same decision, bit for bit.
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

template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(128) k_trace_v1(DevScene sc, TraceArgs a) {
    const uint32_t n = *a.count;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    uint2 stk[64];           // {child code, child meta}
    float stk_t[64];         // tmin of the child's slab test at push time
    uint32_t sp = 0, cur = PN_NONE;
    LaneRay L; L.ray.o = V(0, 0, 0); L.ray.d = V(0, 0, 1); L.ray.mint = 0.f; L.ray.maxt = 0.f; L.invDir = V(0, 0, 0);
    L.negx = L.negy = L.negz = false; L.i = 0; L.best = SPT_MISS;
    bool active = false, exhausted = (n == 0);
    unsigned long long cn = 0, cp = 0;
    for (;;) {
        if (lane_fetch(a, n, lane, active, exhausted, L)) {
            sp = 0; cur = PN_NONE; active = true;
            // the root: the one node whose own box is tested from the reference array
            bool rootHit = false;
            if (sc.n_nodes) {
                float4 n0 = __ldg(&sc.nodes[0]), n1 = __ldg(&sc.nodes[1]);
                if (COUNT) ++cn;
                rootHit = slab(n0, n1, L.ray, L.invDir, L.negx, L.negy, L.negz);
            }
            if (rootHit) {
                if (sc.root_code.y & 0xff) { stk[0] = sc.root_code; stk_t[0] = -SPT_INF; sp = 1; }   // leaf root
                else cur = sc.root_code.x;
            }
        }
        if (!__any_sync(FULL, active)) break;
        while (active) {
            uint2 leaf = make_uint2(0, 0);          // leaf.y & 0xff = nPrims (0: none pending)
            bool done = false;
            if (cur != PN_NONE) {
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
                uint32_t c0 = __float_as_uint(q3.x), c1 = __float_as_uint(q3.y), meta = __float_as_uint(q3.z);
                uint32_t axis = meta & 3;
                uint32_t m0 = (meta >> 8) & 0x1ff, m1 = (meta >> 17) & 0x1ff;      // nPrims | hasQuadric << 8
                bool neg = axis == 0 ? L.negx : (axis == 1 ? L.negy : L.negz);
                // reference: dirIsNeg[axis] ? second child first : first child first
                bool nearHit = neg ? h1 : h0, farHit = neg ? h0 : h1;
                uint32_t nearC = neg ? c1 : c0, farC = neg ? c0 : c1;
                uint32_t nearM = neg ? m1 : m0, farM = neg ? m0 : m1;
                float farT = neg ? t0 : t1;
                if (farHit) { stk[sp] = make_uint2(farC, farM); stk_t[sp] = farT; ++sp; }
                cur = PN_NONE;
                if (nearHit) {
                    if (nearM & 0xff) leaf = make_uint2(nearC, nearM);
                    else cur = nearC;
                }
            }
            if (cur == PN_NONE && (leaf.y & 0xff) == 0) {
                // pop: the far child's slab test completes here with the current maxt
                for (;;) {
                    if (sp == 0) { done = true; break; }
                    --sp;
                    if (stk_t[sp] < L.ray.maxt) {
                        uint2 e = stk[sp];
                        if (e.y & 0xff) leaf = e; else cur = e.x;
                        break;
                    }
                }
            }
            if (leaf.y & 0xff) {
                if (leaf_test<ANY, COUNT>(sc, leaf.x, leaf.y & 0xff, (leaf.y >> 8) & 1, L, cp)) done = true;
            }
            if (done) {
                a.out_slot[L.i] = L.best;
                if (!ANY) a.out_t[L.i] = L.ray.maxt;
                active = false;
            }
            if (active && !exhausted && __popc(__activemask()) < FETCH_THRESHOLD) break;
        }
    }
    if (COUNT && sc.counters && (cn | cp)) {
        atomicAdd(&sc.counters[ANY ? 2 : 0], cn);
        atomicAdd(&sc.counters[ANY ? 3 : 1], cp);
    }
}

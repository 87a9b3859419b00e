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

// ---- variant 1: pair nodes, several ray queues per launch ------------------------------------------------------------
// One 64-byte record per INTERIOR node holds the bounds of both children (built in spt_scene_create): one step fetches
// four 16-byte words, slab-tests both children and descends. A child is one 32-bit code: a pair index, or
// PN_LEAF | hasQuadric << 30 | (nPrims-1) << 27 | first slot.
//
// Closest hit: the reference pushes the far child unconditionally and slab-tests it when popped, against the maxt current
// THEN; of that test only `tmin < maxt` depends on maxt, so the far child is tested at push time, dropped if it already
// fails, and its tmin is kept on the stack and re-compared with the (possibly shrunk) maxt at pop time - the same decision,
// bit for bit. Any hit: maxt never changes during IntersectP, so the pop-time comparison always holds and the answer is the
// OR over the same set of leaves whatever their order. Whether a lane's ray is a closest-hit or an any-hit query is a
// run-time property of the QUEUE it came from: one launch drains up to four queues (a bounce's path rays together with the
// previous bounce's shadow and MIS rays), so a wavefront bounce pays for ONE persistent kernel's drain instead of three or four.
//
#define PN_LEAF 0x80000000u
#define W4_EMPTY_CODE 0xffffffffu
// Warp organisation: one loop - every iteration a lane pops if it must, takes one pair step if it holds an interior node,
// then tests the leaf it may have reached; idle lanes wait for the refill. Two alternatives were built and measured on the
// B200 (profiles/r02_trace_modes.log, whole killeroo frame): while-while (every lane walks to its next leaf, the warp
// reconverges, leaves are tested together) 43.8 ms against 40.9 ms, batched leaves (a leaf phase once 4-16 lanes hold a leaf)
// 41.8-45.1 ms; keeping the first 8 or 16 stack entries of a lane in shared memory ([entry][thread], conflict-free) instead
// of local memory: 45.7 against 44.3 ms. The L1-served local stack and the plain loop won, so they are what is left.
// WIDE: the fast layout of wide.h instead of the pair nodes - four children per step, nearest first; NOT the parity path.
template <bool COUNT, bool WIDE>
__global__ void __launch_bounds__(128) k_trace_multi(DevScene sc, TraceMultiArgs a) {
    uint2 stk[64];                  // the 64-entry todo stack (bvh.cpp:384): {child code, tmin of its slab test at push time}
    // per-queue pointers, looked up by the lane's queue number when it fetches / finishes a ray (a ray finishes with one or
    // two lanes active: a chain of selects there costs whole warp instructions per ray)
    __shared__ TraceSeg s_seg[TRACE_MAX_SEG];
    for (uint32_t k = threadIdx.x; k < TRACE_MAX_SEG; k += blockDim.x) s_seg[k] = a.seg[k < a.nseg ? k : 0];
    __syncthreads();
    const uint32_t e0 = *a.seg[0].count;
    const uint32_t e1 = e0 + (a.nseg > 1 ? *a.seg[1].count : 0u);
    const uint32_t e2 = e1 + (a.nseg > 2 ? *a.seg[2].count : 0u);
    const uint32_t total = e2 + (a.nseg > 3 ? *a.seg[3].count : 0u);
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    uint32_t sp = 0, cur = PN_NONE, negMask = 0, seg = 0;
    Ray ray; ray.o = V(0, 0, 0); ray.d = V(0, 0, 1); ray.mint = 0.f; ray.maxt = 0.f;
    v3 invDir = V(0, 0, 0);
    bool exact = false, isAny = false;
    uint32_t ri = 0, best = SPT_MISS;
    bool active = false, exhausted = (total == 0);
    unsigned long long cn = 0, cp = 0, cnA = 0, cpA = 0;
    // pop: the next stack entry whose slab test still holds with the current maxt; true when the stack ran empty
    auto pop = [&]() -> bool {
        for (;;) {
            if (sp == 0) return true;
            const uint2 e = stk[--sp];
            // the far child's slab test completes here with the current maxt
            if (__uint_as_float(e.y) < ray.maxt) { cur = e.x; return false; }
        }
    };
    // one pair step from the interior node `cur`: both children slab-tested, the far one pushed, the near one entered
    auto node_step = [&]() {
        const float4 *pn = sc.pnodes + 4 * (size_t)cur;
        const float4 q0 = __ldg(pn), q1 = __ldg(pn + 1), q2 = __ldg(pn + 2), q3 = __ldg(pn + 3);
        if (COUNT) { if (isAny) cnA += 2; else cn += 2; }
        // child 0: min (q0.x,q0.y,q0.z) max (q0.w,q1.x,q1.y); child 1: min (q1.z,q1.w,q2.x) max (q2.y,q2.z,q2.w)
        float t0, t1;
        bool h0, h1;
        if (!exact) {
            h0 = slab6_finite(q0.x, q0.w, q0.y, q1.x, q0.z, q1.y, ray, invDir, &t0);
            h1 = slab6_finite(q1.z, q2.y, q1.w, q2.z, q2.x, q2.w, ray, invDir, &t1);
        } else {
            const bool negx = negMask & 1u, negy = negMask & 2u, negz = negMask & 4u;
            h0 = slab6(negx ? q0.w : q0.x, negx ? q0.x : q0.w, negy ? q1.x : q0.y, negy ? q0.y : q1.x,
                       negz ? q1.y : q0.z, negz ? q0.z : q1.y, ray, invDir, &t0);
            h1 = slab6(negx ? q2.y : q1.z, negx ? q1.z : q2.y, negy ? q2.z : q1.w, negy ? q1.w : q2.z,
                       negz ? q2.w : q2.x, negz ? q2.x : q2.w, ray, invDir, &t1);
        }
        const uint32_t c0 = __float_as_uint(q3.x), c1 = __float_as_uint(q3.y), axis = __float_as_uint(q3.z);
        // reference: dirIsNeg[axis] ? second child first : first child first
        const bool swap = (negMask >> axis) & 1u;
        const uint32_t nearC = swap ? c1 : c0, farC = swap ? c0 : c1;
        const bool nearHit = swap ? h1 : h0, farHit = swap ? h0 : h1;
        if (farHit) stk[sp++] = make_uint2(farC, __float_as_uint(swap ? t0 : t1));
        cur = nearHit ? nearC : PN_NONE;
    };
    // fast layout: one step tests the four children of the wide node `cur` (a slab plane is one fused multiply-add:
    // b * invDir - o * invDir), enters the nearest one and pushes the others farthest first
    v3 noid = V(0, 0, 0);            // -o * invDir
    auto wide_step = [&]() {
        const float4 *wn = sc.wnodes + 8 * (size_t)cur;
        const float4 lx = __ldg(wn), ly = __ldg(wn + 1), lz = __ldg(wn + 2), hx = __ldg(wn + 3), hy = __ldg(wn + 4), hz = __ldg(wn + 5);
        const float4 cc = __ldg(wn + 6);
        if (COUNT) { if (isAny) cnA += 4; else cn += 4; }
        float tn[4]; uint32_t cd[4];
#define W4_CHILD(i, LX, HX, LY, HY, LZ, HZ, C) do { \
            const float ax = fmaf(LX, invDir.x, noid.x), bx = fmaf(HX, invDir.x, noid.x); \
            const float ay = fmaf(LY, invDir.y, noid.y), by = fmaf(HY, invDir.y, noid.y); \
            const float az = fmaf(LZ, invDir.z, noid.z), bz = fmaf(HZ, invDir.z, noid.z); \
            const float tnear = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), ray.mint)); \
            const float tfar = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), ray.maxt)); \
            cd[i] = __float_as_uint(C); tn[i] = (tnear <= tfar && cd[i] != W4_EMPTY_CODE) ? tnear : SPT_INF; } while (0)
        W4_CHILD(0, lx.x, hx.x, ly.x, hy.x, lz.x, hz.x, cc.x);
        W4_CHILD(1, lx.y, hx.y, ly.y, hy.y, lz.y, hz.y, cc.y);
        W4_CHILD(2, lx.z, hx.z, ly.z, hy.z, lz.z, hz.z, cc.z);
        W4_CHILD(3, lx.w, hx.w, ly.w, hy.w, lz.w, hz.w, cc.w);
#undef W4_CHILD
        if (!isAny) {
            // sorting network on (entry distance, code): nearest first; a miss carries +inf and sinks to the end
#define W4_SWAP(i, j) do { if (tn[j] < tn[i]) { const float t_ = tn[i]; tn[i] = tn[j]; tn[j] = t_; const uint32_t c_ = cd[i]; cd[i] = cd[j]; cd[j] = c_; } } while (0)
            W4_SWAP(0, 1); W4_SWAP(2, 3); W4_SWAP(0, 2); W4_SWAP(1, 3); W4_SWAP(1, 2);
#undef W4_SWAP
        }
        // enter the first (nearest) child that was hit, push the others so that the nearer ones come off the stack first
        cur = PN_NONE;
        float tcur = 0.f;
#pragma unroll
        for (int i = 3; i >= 0; --i) {
            if (tn[i] == SPT_INF) continue;
            if (cur != PN_NONE) stk[sp++] = make_uint2(cur, __float_as_uint(tcur));
            cur = cd[i]; tcur = tn[i];
        }
    };
    // the primitives of the leaf `cur`, in slot order; true when an any-hit ray is finished
    auto leaf_step = [&]() -> bool {
        const uint32_t offset = cur & 0x07ffffffu, nPrims = ((cur >> 27) & 7u) + 1u;
        const bool hasQuadric = (cur >> 30) & 1u;
        cur = PN_NONE;
        for (uint32_t k = 0; k < nPrims; ++k) {
            const uint32_t s = offset + k;
            if (COUNT) { if (isAny) ++cpA; else ++cp; }
            float t;
            bool h;
            if (!hasQuadric || sc.prim_kind[s] == SPT_PRIM_TRIANGLE) {
                const float4 va = __ldg(&sc.tri_verts[3 * (size_t)s]);
                const float4 vb = __ldg(&sc.tri_verts[3 * (size_t)s + 1]);
                const float4 vc = __ldg(&sc.tri_verts[3 * (size_t)s + 2]);
                float b1, b2;
                h = tri_test(V(va.x, va.y, va.z), V(vb.x, vb.y, vb.z), V(vc.x, vc.y, vc.z), ray, &t, &b1, &b2);
            } else if (sc.prim_kind[s] == SPT_PRIM_SPHERE) {
                h = sphere_intersect(sc, sc.quadrics[sc.prim_data[s]], 0, ray, &t, nullptr);
            } else {
                h = disk_intersect(sc, sc.quadrics[sc.prim_data[s]], 0, ray, &t, nullptr);
            }
            if (h) {
                best = s;
                if (isAny) return true;                  // IntersectP returns at the first accepted primitive
                ray.maxt = t;
            }
        }
        return false;
    };
    auto finish = [&]() {
        s_seg[seg].out_slot[ri] = best;
        if (!isAny) s_seg[seg].out_t[ri] = ray.maxt;
        active = false;
    };
    for (;;) {
        // ---- refill: idle lanes claim the next entries of the concatenated queues with one warp-aggregated atomic
        {
            const unsigned idle = __ballot_sync(FULL, !active);
            if (!exhausted && idle) {
                uint32_t base = 0;
                const int leader = __ffs(idle) - 1;
                if (lane == leader) base = atomicAdd(a.work, (uint32_t)__popc(idle));
                base = __shfl_sync(FULL, base, leader);
                if (!active) {
                    const uint32_t q = base + __popc(idle & ((1u << lane) - 1));
                    if (q < total) {
                        seg = (q >= e0 ? 1u : 0u) + (q >= e1 ? 1u : 0u) + (q >= e2 ? 1u : 0u);
                        const uint32_t ql = q - (seg == 0 ? 0u : (seg == 1 ? e0 : (seg == 2 ? e1 : e2)));
                        const uint32_t *queue = s_seg[seg].queue;
                        const float4 *ro = s_seg[seg].ro, *rd = s_seg[seg].rd;
                        isAny = s_seg[seg].any != 0u;
                        ri = queue ? queue[ql] : ql;
                        const float4 o = ro[ri], d = rd[ri];
                        ray.o = V(o.x, o.y, o.z); ray.d = V(d.x, d.y, d.z); ray.mint = o.w; ray.maxt = d.w;
                        invDir = V(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
                        const bool negx = invDir.x < 0, negy = invDir.y < 0, negz = invDir.z < 0;
                        negMask = (negx ? 1u : 0u) | (negy ? 2u : 0u) | (negz ? 4u : 0u);
                        exact = !(isfinite(invDir.x) && isfinite(invDir.y) && isfinite(invDir.z));
                        best = SPT_MISS; sp = 0; cur = PN_NONE; active = true;
                        if (WIDE) {
                            // a direction component of (almost) zero: a huge finite reciprocal keeps b * invDir - o * invDir free of
                            // inf - inf; the slab of that axis then only excludes boxes the origin lies outside of
                            const float tiny = 1e-18f;
                            if (fabsf(ray.d.x) < tiny) invDir.x = copysignf(1.f / tiny, ray.d.x);
                            if (fabsf(ray.d.y) < tiny) invDir.y = copysignf(1.f / tiny, ray.d.y);
                            if (fabsf(ray.d.z) < tiny) invDir.z = copysignf(1.f / tiny, ray.d.z);
                            noid = V(-ray.o.x * invDir.x, -ray.o.y * invDir.y, -ray.o.z * invDir.z);
                            cur = sc.wroot == W4_EMPTY_CODE ? PN_NONE : sc.wroot;
                            if (COUNT) { if (isAny) ++cnA; else ++cn; }
                        } else
                        // the root: the one node whose own box is tested from the reference array
                        if (sc.n_nodes) {
                            const float4 n0 = __ldg(&sc.nodes[0]), n1 = __ldg(&sc.nodes[1]);
                            if (COUNT) { if (isAny) ++cnA; else ++cn; }
                            if (slab(n0, n1, ray, invDir, negx, negy, negz)) cur = sc.root_code;
                        }
                    }
                }
                if (base + (uint32_t)__popc(idle) >= total) exhausted = true;
            }
        }
        if (!__any_sync(FULL, active)) break;
        while (active) {
            bool done = false;
            if (cur == PN_NONE) done = pop();
            if (!done && cur < PN_LEAF) { if (WIDE) wide_step(); else node_step(); }
            if (!done && cur >= PN_LEAF && cur != PN_NONE) done = leaf_step();
            if (done) finish();
            if (active && !exhausted && (uint32_t)__popc(__activemask()) < a.fetch_threshold) break;
        }
    }
    if (COUNT && sc.counters) {
        if (cn | cp) { atomicAdd(&sc.counters[0], cn); atomicAdd(&sc.counters[1], cp); }
        if (cnA | cpA) { atomicAdd(&sc.counters[2], cnA); atomicAdd(&sc.counters[3], cpA); }
    }
}

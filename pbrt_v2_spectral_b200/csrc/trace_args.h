#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
struct TraceArgs {
    const uint32_t *queue;      // ray indices, or NULL for 0..n-1
    const uint32_t *count;      // number of rays (device word)
    uint32_t *work;             // next unclaimed queue position (device word, zeroed)
    const float4 *ro, *rd;      // {o, mint}, {d, maxt}
    uint32_t *out_slot;         // BVH slot of the accepted primitive, SPT_MISS if none
    float *out_t;               // closest hit: ray.maxt after the traversal (NULL for any-hit)
    uint32_t fetch_threshold;   // refill the warp from the queue when fewer lanes than this still traverse
};

// One launch of the pair-node kernel drains up to four ray queues ("segments"), each a closest-hit or an any-hit class.
#define TRACE_MAX_SEG 4
struct TraceSeg {
    const uint32_t *queue;      // ray indices, or NULL for 0..n-1
    const uint32_t *count;      // number of rays (device word)
    const float4 *ro, *rd;      // {o, mint}, {d, maxt}
    uint32_t *out_slot;         // BVH slot of the accepted primitive, SPT_MISS if none
    float *out_t;               // closest hit: ray.maxt after the traversal (unused for any-hit)
    uint32_t any, pad_;         // 1: IntersectP semantics
};
struct TraceMultiArgs {
    TraceSeg seg[TRACE_MAX_SEG];
    uint32_t nseg;
    uint32_t fetch_threshold;   // refill the warp from the queues when fewer lanes than this still traverse
    uint32_t *work;             // next unclaimed position of the concatenated queues (device word, zeroed)
};

// wide.h — the FAST traversal layout: a 4-wide BVH collapsed from the reference's flattened binary tree
// (src/accelerators/bvh.cpp:354-372 writes it, :105-115 is its node). One 128-byte node holds the bounds of up to four
// children, SoA per axis: a ray takes half as many dependent node fetches as on the bit-exact pair-node walk, tests four
// boxes per fetch with one FMA per slab plane, and enters the children nearest first.
//
// Bounds stay fp32: an 8-bit quantised variant (64-byte nodes) was costed and dropped - the traversal kernels are
// instruction-issue bound on L2-resident trees, and dequantising 24 bounds per node (byte extract, int-to-float, scale) costs
// more issue slots than the pair-node walk spends on the same four boxes.
//
// It is NOT the parity path: children are visited nearest first by their entry distance instead of by the reference's split
// axis rule, a slab plane is one fused multiply-add ((b - o) * invDir as b * invDir - o * invDir) and the boxes are padded by
// a few ulps of the scene's extent to keep that conservative, so
//   * the accepted primitive is the same wherever the closest hit is unique - the leaf code and the triangle / quadric tests
//     are the parity path's, so its distance is then bit-identical;
//   * rays with two primitives at exactly the same distance may report the other one, and a ray that grazes a box the
//     reference's slab test rejects by rounding may find a hit the reference misses (both counted in tests/).
// Scenes opt in with spt_scene_set_traversal(scene, SPT_TRAVERSAL_FAST); the default stays bit-exact.
#pragma once
#include <stdint.h>

#define W4_EMPTY 0xffffffffu
struct W4Node {                       // 128 bytes = 8 x 16
    float lo[3][4], hi[3][4];         // [axis][child]; an empty slot (child == W4_EMPTY) holds zeros
    uint32_t child[4];                // interior: index of a W4Node; leaf: 1<<31 | hasQuadric<<30 | (nPrims-1)<<27 | first slot; W4_EMPTY
    uint32_t pad[4];
};

// Host-side builder: nodes = the reference's LinearBVHNode array with the leaf flags of spt_build.cu (meta bit 16 =
// hasQuadric). Returns the root's child code through *root (a leaf code when the whole tree is one leaf); false when a leaf
// does not pack into a child code (more than 8 primitives, 2^27 slots) or the tree is malformed.
#ifdef __cplusplus
#include <vector>
bool spt_build_w4(const void *ref_nodes, uint32_t n_nodes, std::vector<W4Node> *out, uint32_t *root);
#endif

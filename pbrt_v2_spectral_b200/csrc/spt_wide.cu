// spt_wide.cu — host-side builder of the fast traversal layout (wide.h). No kernels: the tree is collapsed once per scene,
// on the first spt_scene_set_traversal(.., SPT_TRAVERSAL_FAST).
#include <cmath>
#include <cstring>
#include <limits>
#include <utility>
#include "wide.h"

namespace {
struct RefNode { float b[6]; uint32_t off; uint32_t meta; };     // LinearBVHNode: meta = nPrims | axis << 8 | hasQuadric << 16
inline float area(const RefNode &n) {
    float dx = n.b[3] - n.b[0], dy = n.b[4] - n.b[1], dz = n.b[5] - n.b[2];
    return dx * dy + dy * dz + dz * dx;
}
}  // namespace

bool spt_build_w4(const void *ref_nodes, uint32_t n_nodes, std::vector<W4Node> *out, uint32_t *root) {
    const RefNode *rn = (const RefNode *)ref_nodes;
    out->clear();
    *root = W4_EMPTY;
    if (n_nodes == 0) return true;
    auto leaf_code = [&](uint32_t n, uint32_t *code) {
        uint32_t np = rn[n].meta & 0xffu;
        if (np > 8 || rn[n].off >= (1u << 27)) return false;
        *code = 0x80000000u | (((rn[n].meta >> 16) & 1u) ? 0x40000000u : 0u) | (np - 1u) << 27 | rn[n].off;
        return true;
    };
    if (rn[0].meta & 0xffu) return leaf_code(0, root);
    // padding of every box: a slab plane is evaluated as fma(b, invDir, -o * invDir), whose error in world units is a few
    // ulps of |o|; ray origins lie in or around the scene, so a few ulps of the scene's largest coordinate cover it
    float reach = 0.f;
    for (int k = 0; k < 6; ++k) reach = std::fmax(reach, std::fabs(rn[0].b[k]));
    const float pad = 4.f * std::numeric_limits<float>::epsilon() * reach;
    std::vector<std::pair<uint32_t, uint32_t>> todo;     // (binary interior node, wide node made from it)
    out->push_back(W4Node());
    todo.push_back({0u, 0u});
    *root = 0;
    while (!todo.empty()) {
        const uint32_t bn = todo.back().first, wn = todo.back().second;
        todo.pop_back();
        // collapse: start from the two children, repeatedly open the interior child with the largest surface area
        uint32_t kids[4] = { bn + 1, rn[bn].off, 0, 0 };
        int nk = 2;
        if (kids[0] >= n_nodes || kids[1] >= n_nodes) return false;
        while (nk < 4) {
            int best = -1; float bestA = -1.f;
            for (int k = 0; k < nk; ++k)
                if (!(rn[kids[k]].meta & 0xffu) && area(rn[kids[k]]) > bestA) { bestA = area(rn[kids[k]]); best = k; }
            if (best < 0) break;
            const uint32_t open = kids[best];
            if (open + 1 >= n_nodes || rn[open].off >= n_nodes) return false;
            kids[best] = open + 1;
            kids[nk++] = rn[open].off;
        }
        W4Node q;
        memset(&q, 0, sizeof(q));
        for (int k = 0; k < 4; ++k) {
            if (k >= nk) {
                q.child[k] = W4_EMPTY;
                continue;
            }
            const RefNode &c = rn[kids[k]];
            for (int a = 0; a < 3; ++a) { q.lo[a][k] = c.b[a] - pad; q.hi[a][k] = c.b[3 + a] + pad; }
            if (c.meta & 0xffu) { if (!leaf_code(kids[k], &q.child[k])) return false; }
            else {
                q.child[k] = (uint32_t)out->size();
                out->push_back(W4Node());
                todo.push_back({kids[k], q.child[k]});
            }
        }
        (*out)[wn] = q;
    }
    return true;
}

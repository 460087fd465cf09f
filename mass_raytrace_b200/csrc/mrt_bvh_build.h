// mrt_bvh_build.h — host-side binned-SAH BVH builder used by mrt_scene_upload (unless MRT_SCENE_KEEP_TOPOLOGY is set).
//
// The reference builds its BVH with a random split axis and an object-median split (BvhNode::new, geom.rs:109-161). A closest
// hit does not depend on the tree's topology (except for the order in which exact-t ties are met), so the library is free to
// traverse a better tree than the one the caller passed: surface-area-heuristic splits chosen over 16 centroid bins per axis,
// leaves of up to `max_leaf` primitives. Past a depth budget the builder switches to object-median splits so that the tree
// always fits the traversal stack. Implementation: mrt_bvh_build.cpp (SSE, multi-threaded; compiled by the host compiler).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

namespace mrt_build {

struct alignas(16) Prim {  // two 16-byte halves: (lo, ref) and (hi, pad), loaded as one SSE register each
    float lo[3];
    uint32_t ref;  // caller's payload (primitive reference or triangle index)
    float hi[3];
    uint32_t pad;
};

struct Node {  // build node (trivial: the node pool is allocated without being touched)
    float lo[3], hi[3];
    int32_t left, right;    // node indices; leaf when left < 0
    uint32_t first, count;  // primitive range of a leaf (indices into the reordered prim array)
};

struct Tree {
    std::unique_ptr<Node[]> nodes;
    int32_t root = -1;
    int depth = 0;  // inner nodes on the longest root-to-leaf path
};

// Builds over prims (reordered in place so that every leaf is a contiguous range). cost_prim = cost of testing one primitive
// relative to one node visit.
Tree build_sah(std::vector<Prim>& prims, int max_leaf, int sah_depth_budget, float cost_prim);

}  // namespace mrt_build

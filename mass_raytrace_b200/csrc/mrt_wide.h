// mrt_wide.h — the acceleration structure the kernels traverse: an 8-wide BVH with child boxes quantised to the parent's frame.
//
// The reference walks a binary tree of pointer nodes, one box test per visit (BvhNode::intersect geom.rs:186-200, BoundingBox::hit
// :218-247). Here one 80-byte record decides up to eight children at once (after Ylitie, Karras, Laine 2017, "Efficient
// incoherent ray traversal on GPUs through compressed wide BVHs"): the node stores its own origin p and one power-of-two scale
// per axis, each child box is six bytes on that grid (rounded outwards), and the children sit in slots chosen at build time so
// that "slot index XOR ray octant" is a front-to-back order -- no distance sort at run time. Inner children of a node are
// stored consecutively (index = child_base + rank among the inner slots), the primitives of its leaf children consecutively
// from prim_base (at most 3 per leaf child, 24 per node). A closest hit does not depend on the tree, and the quantised boxes only
// ever contain the exact ones, so everything that reaches a hit record is still computed by the unchanged primitive tests.
//
// Shared by the host builder (mrt_bvh_build.cpp, g++), the kernels (mrt_device.cuh, nvcc) and the native checker of the test
// suite (tests/native/check_bvh_build.cpp), which runs wide_node_test on the CPU against brute force.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MRT_HD __host__ __device__ __forceinline__
#else
#define MRT_HD inline
#endif

namespace mrt {

constexpr int kWideWidth = 8;      // child slots per node
constexpr int kWideLeafPrims = 3;  // primitives per leaf child (unary count in three bits)

struct alignas(16) WNode {  // 80 B = five 16-byte words
    float p[3];             // origin of the quantisation grid (= the node's own box minimum)
    uint8_t e[3];           // per axis: biased binary exponent of the grid step (step = 2^(e - 127))
    uint8_t imask;          // bit s set: slot s holds an inner node
    uint32_t child_base;    // node index of the first inner child
    uint32_t prim_base;     // first primitive of this node's leaf children (BLAS: triangle in device order; TLAS: entry of the object list)
    uint8_t meta[8];        // per slot. inner: 0b001'11sss (s = slot); leaf: unary count << 5 | offset from prim_base (0..23); empty: 0
    uint8_t qlo[3][8];      // per axis, per slot: box minimum on the grid (rounded down)
    uint8_t qhi[3][8];      // box maximum (rounded up)
};
static_assert(sizeof(WNode) == 80, "WNode layout");

struct W4 {  // one 16-byte word of a node as the kernels load it
    uint32_t x, y, z, w;
};

MRT_HD float wide_as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// Per-ray constants of the node test: origin, ~1/direction, and (7 - octant) in every byte, octant = 4 [dx < 0] + 2 [dy < 0] + [dz < 0].
struct WideRay {
    float ox, oy, oz;
    float idx, idy, idz;
    uint32_t oct_inv4;
};
MRT_HD uint32_t wide_oct_inv4(float dx, float dy, float dz) {
    const uint32_t oct = (dx < 0.0f ? 4u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 1u : 0u);
    return (7u - oct) * 0x01010101u;
}

// Tests the eight child boxes of one node against a ray; returns the hit mask: bits 24..31 = inner children, ordered so that the
// HIGHEST set bit is the child to enter first (slot ^ (7 - octant)), bits 0..23 = the primitives (offsets from prim_base) of the
// leaf children that were hit.
//
// Conservative by construction, like the two-box test it replaces: a plane at grid coordinate q lies at t = q * (step / d) +
// (p - o) / d. With id ~ 1/d (<= 1 ulp) the computed value is off by at most 1.5 * 2^-22 |(p - o) id| + 1.5 * 2^-23 |t|. The first
// part is covered per node and per axis by moving the near planes 2^-21 |(p - o) id| earlier and the far planes as much later, the
// second by widening the final interval by 2^-20 relatively. A NaN (0 * inf on an axis the ray is parallel to) drops that axis
// from fminf / fmaxf, which only accepts more. Nothing computed here reaches a hit record.
MRT_HD uint32_t wide_node_test(const W4& n0, const W4& n1, const W4& n2, const W4& n3, const W4& n4, const WideRay& r, float t_min, float t_max) {
    const float kSlack = 4.76837158203125e-7f;   // 2^-21
    const float kWiden = 9.5367431640625e-7f;    // 2^-20
    const uint32_t e = n0.w;
    const float sx = wide_as_float((e & 0xffu) << 23) * r.idx;
    const float sy = wide_as_float(((e >> 8) & 0xffu) << 23) * r.idy;
    const float sz = wide_as_float(((e >> 16) & 0xffu) << 23) * r.idz;
    const float bx = (wide_as_float(n0.x) - r.ox) * r.idx;
    const float by = (wide_as_float(n0.y) - r.oy) * r.idy;
    const float bz = (wide_as_float(n0.z) - r.oz) * r.idz;
    const float ex = fabsf(bx) * kSlack, ey = fabsf(by) * kSlack, ez = fabsf(bz) * kSlack;
    const float bnx = bx - ex, bfx = bx + ex, bny = by - ey, bfy = by + ey, bnz = bz - ez, bfz = bz + ez;
    // near / far plane of each axis by the sign of the direction: qlo words are n2.xy (x) n2.zw (y) n3.xy (z), qhi n3.zw n4.xy n4.zw
    const bool nx = r.idx < 0.0f, ny = r.idy < 0.0f, nz = r.idz < 0.0f;
    uint32_t hit = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xffu;
        const uint32_t bit_index4 = (meta4 ^ (r.oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
        const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
        const uint32_t qnx = nx ? hix : lox, qfx = nx ? lox : hix;
        const uint32_t qny = ny ? hiy : loy, qfy = ny ? loy : hiy;
        const uint32_t qnz = nz ? hiz : loz, qfz = nz ? loz : hiz;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int sh = 8 * j;
            const float tnx = fmaf((float)((qnx >> sh) & 0xffu), sx, bnx), tfx = fmaf((float)((qfx >> sh) & 0xffu), sx, bfx);
            const float tny = fmaf((float)((qny >> sh) & 0xffu), sy, bny), tfy = fmaf((float)((qfy >> sh) & 0xffu), sy, bfy);
            const float tnz = fmaf((float)((qnz >> sh) & 0xffu), sz, bnz), tfz = fmaf((float)((qfz >> sh) & 0xffu), sz, bfz);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, t_min));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, t_max));
            if (tn <= fmaf(tf, kWiden, tf)) hit |= ((child_bits4 >> sh) & 0xffu) << ((bit_index4 >> sh) & 0xffu);
        }
    }
    return hit;
}

// ---- stepping through a hit mask (the same few lines on the device and in the checker) ---------------------------------------
// A "group" is (base index, bits): for inner children base = child_base and bits = hits << 24 | imask; for primitives base =
// prim_base and bits = the 24-bit primitive mask. The root of a tree is the group (root index, 0x80000000): imask 0, one hit.
constexpr uint32_t kWideRootBits = 0x80000000u;
MRT_HD bool wide_has_nodes(uint32_t bits) { return bits > 0x00ffffffu; }
MRT_HD int wide_high_bit(uint32_t v) {  // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
MRT_HD int wide_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// takes the next inner child out of a node group: returns its node index
MRT_HD uint32_t wide_next_child(uint32_t base, uint32_t& bits, uint32_t oct_inv4) {
    const int bit = wide_high_bit(bits);
    bits &= ~(1u << bit);
    const uint32_t slot = (uint32_t)(bit - 24) ^ (oct_inv4 & 0xffu);
    const uint32_t imask = bits & 0xffu;  // (the low byte is never touched by the line above: bit >= 24)
    return base + (uint32_t)wide_popc(imask & ~(0xffffffffu << slot));
}

}  // namespace mrt

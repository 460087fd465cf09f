// mrt_device.cuh — device-side scene layout, Philox, intersection, traversal and shading for sm_100a.
//
// Arithmetic contract: every value that decides a hit or ends up in a hit record (ray transforms, sphere and
// triangle tests, hit point, interpolated normal, face forwarding) is computed in the reference's f32 operation
// order with IEEE add/mul/div/sqrt and NO fused multiply-add (this TU is compiled with -fmad=false, default
// -prec-div/-prec-sqrt, no fast-math), so primary-ray t / normal / ids are bit-identical to the CPU restatement.
// Only the AABB slab test deviates: it multiplies by a per-ray reciprocal instead of dividing (geom.rs:219-220
// divides) and widens the interval by 4 ulp so that it never rejects a box the reference's test accepts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mrt.h"

namespace mrt {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kStackSize = 96;
// device-side triangle leaf reference: kind TRIANGLE | (count-1) << 27 | first triangle (device order, 27 bits)
constexpr uint32_t kTriIndexMask = 0x07FFFFFFu;
constexpr uint32_t kTriAlphaFlag = 1u;  // DTriVerts.a.w (as bits): the triangle's own material can fail alpha_test (geom.rs:568)

// ---- device scene (HBM layout; all arrays 16-byte aligned, fetched with 128-bit loads) -----------------------
// Inner node, 64 B = one half cache line: both children's AABBs + both child refs, so one fetch decides both
// subtrees (the reference stores one AABB per node and tests it after the pointer chase, geom.rs:103-107, 186-187).
struct __align__(16) DNode {
    float4 xy0;  // child0: min.x max.x min.y max.y
    float4 xy1;  // child1: min.x max.x min.y max.y
    float4 z01;  // child0 min.z max.z, child1 min.z max.z
    uint32_t child0, child1, pad0, pad1;  // prim refs (MRT_REF); kNone = absent
};
struct __align__(16) DTriVerts {
    float4 a, b, c;  // vertex_a/b/c, w unused
};
struct __align__(16) DInstance {
    float4 inv0, inv1, inv2;  // inv_transform columns c0.xyz c1.xyz c2.xyz c3.xyz packed as 12 floats
    float4 fwd0, fwd1, fwd2;  // transform, same packing
    uint32_t root;            // BLAS root ref
    int32_t material;         // override or -1
    uint32_t flags;
    uint32_t object_id;
    uint32_t pad[4];
};
struct DSphereAux {
    int32_t material;
    uint32_t object_id;
};
struct DScene {
    const DNode* nodes;
    const float4* spheres;  // center.xyz, radius
    const DSphereAux* sphere_aux;
    const DTriVerts* tri_verts;          // device (BVH leaf) order
    const uint32_t* tri_map;             // device triangle index -> caller's triangle index
    const mrt_tri_shading* tri_shading;  // caller's order
    const DInstance* instances;
    const mrt_blas* blas;
    const mrt_volume* volumes;
    const mrt_material* materials;
    const mrt_surface* surfaces;
    const mrt_texture* textures;
    const float4* texels;
    uint32_t root;                       // the one root reference (node or single primitive); kNone = empty world
    uint32_t n_volumes;
    uint32_t has_alpha;                  // the scene needs the full intersection code (ALPHA variants): some triangle carries kTriAlphaFlag, or a Volume fills a mesh
    mrt_background bg;
};
struct DCamera {
    float3 origin, llc, horizontal, vertical, u, v;
    float lens_radius;
};

struct VisitCounters {
    unsigned long long node_visits, tri_tests, sphere_tests, instance_tests, volume_tests;
    unsigned long long max_ray_nodes;  // most node visits of one ray
};

// ---- f32 vector helpers in the reference's operation order (math/generic.rs) ---------------------------------
struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 v3(float4 a) { return V3{a.x, a.y, a.z}; }
__device__ __forceinline__ V3 v3(float3 a) { return V3{a.x, a.y, a.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // generic.rs:8-10
__device__ __forceinline__ V3 cross(V3 a, V3 b) {                                               // generic.rs:12-18
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float length_squared(V3 a) { return dot(a, a); }
__device__ __forceinline__ float length(V3 a) { return sqrtf(length_squared(a)); }
__device__ __forceinline__ V3 unit(V3 a) { return a / length(a); }  // math.rs:75-77
__device__ __forceinline__ bool near_zero(V3 a) { return fabsf(a.x) <= 0.00001f && fabsf(a.y) <= 0.00001f && fabsf(a.z) <= 0.00001f; }
__device__ __forceinline__ V3 reflect(V3 v, V3 n) { return v - (n * dot(v, n) * 2.0f); }  // math.rs:115-117
__device__ __forceinline__ V3 refract(V3 v, V3 n, float eta) {                            // math.rs:119-124
    float cos_theta = fminf(dot(-v, n), 1.0f);
    V3 perp = (v + n * cos_theta) * eta;
    V3 par = n * (-sqrtf(fabsf(1.0f - length_squared(perp))));
    return perp + par;
}
// M4::transform (generic.rs:105-115) on the packed 3x4: ((c0*x + c1*y) + c2*z) + c3*w
__device__ __forceinline__ V3 xform(float4 m0, float4 m1, float4 m2, V3 p, float w) {
    return V3{((m0.x * p.x + m0.w * p.y) + m1.z * p.z) + m2.y * w,
              ((m0.y * p.x + m1.x * p.y) + m1.w * p.z) + m2.z * w,
              ((m0.z * p.x + m1.y * p.y) + m2.x * p.z) + m2.w * w};
}

struct Ray {
    V3 o, d;
};
__device__ __forceinline__ V3 ray_at(const Ray& r, float t) { return r.o + (r.d * t); }  // world.rs:179-181

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (pixel, sample, bounce, stream), key = seed -----------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// (0, 1), 23 bits like the reference's f32 generator (fastrand: 23 mantissa bits), but never 0: the closed-form ball sampler would
// turn a 0 into a zero-length direction, and log(0) into an infinite free-flight distance
__device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 9) + 0.5f) * 1.1920928955078125e-7f; }
struct Rand4 {
    float x, y, z, w;
};
// Stream ids: the KIND of draw sits in the top four bits and the index (volume number, Mix nesting level, triangle) below them, so
// that two different uses at the same (pixel, sample, bounce) can never share a counter whatever the index is.
enum : uint32_t { kStreamScatter = 0u << 28, kStreamVolume = 1u << 28, kStreamMix = 2u << 28, kStreamEmit = 3u << 28, kStreamAlpha = 4u << 28 };
constexpr uint32_t kStreamIndexMask = (1u << 28) - 1u;
constexpr uint32_t kBounceCamera = 0xFFFFFFFFu;
struct RngKey {
    uint32_t pixel, sample, bounce;
    uint2 seed;
};
__device__ __forceinline__ Rand4 draw4(const RngKey& k, uint32_t stream) {
    uint4 r = philox4x32_10(make_uint4(k.pixel, k.sample, k.bounce, stream), k.seed);
    return Rand4{u01(r.x), u01(r.y), u01(r.z), u01(r.w)};
}
// Closed-form samplers with the same distributions as the reference's rejection loops (math.rs:80-109):
// uniform in the unit disk, uniform on the unit sphere, uniform in the unit ball.
__device__ __forceinline__ V3 sample_unit_disk(float a, float b) {
    float s, c;
    sincospif(2.0f * b, &s, &c);
    float r = sqrtf(a);
    return V3{r * c, r * s, 0.0f};
}
__device__ __forceinline__ V3 sample_unit_vector(float a, float b) {
    float z = 1.0f - 2.0f * a;
    float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float s, c;
    sincospif(2.0f * b, &s, &c);
    return V3{r * c, r * s, z};
}
__device__ __forceinline__ V3 sample_unit_ball(float a, float b, float c3) { return sample_unit_vector(a, b) * cbrtf(c3); }

// ---- primitive tests -----------------------------------------------------------------------------------------
// Sphere::intersect geom.rs:57-93 (t only; normal/point are recomputed for the closest hit by resolve_hit)
__device__ __forceinline__ bool sphere_test(float4 s, const Ray& r, float t_min, float t_max, float& t_out) {
    V3 oc = r.o - v3(s);
    float a = length_squared(r.d);
    float half_b = dot(oc, r.d);
    float c = length_squared(oc) - (s.w * s.w);
    float disc = (half_b * half_b) - (a * c);
    if (disc < 0.0f) return false;
    float sqrt_d = sqrtf(disc);
    float root = (-half_b - sqrt_d) / a;
    if (root < t_min || t_max < root) {
        root = (-half_b + sqrt_d) / a;
        if (root < t_min || t_max < root) return false;
    }
    t_out = root;
    return true;
}
// Triangle::intersect geom.rs:504-533 (Moller-Trumbore part; the area-barycentric tail runs once, in resolve_hit)
__device__ __forceinline__ bool triangle_test(const DTriVerts& tv, const Ray& r, float t_min, float t_max, float& t_out) {
    V3 va = v3(tv.a), vb = v3(tv.b), vc = v3(tv.c);
    V3 ab = vb - va, ac = vc - va;
    V3 p_vec = cross(r.d, ac);
    float det = dot(ab, p_vec);
    if (fabsf(det) < 0.000001f) return false;
    float inv_det = 1.0f / det;
    V3 t_vec = r.o - va;
    float u = dot(t_vec, p_vec) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    V3 q_vec = cross(t_vec, ab);
    float v = dot(r.d, q_vec) * inv_det;
    if (v < 0.0f || v + u > 1.0f) return false;
    float t = dot(ac, q_vec) * inv_det;
    if (t < t_min || t > t_max) return false;
    t_out = t;
    return true;
}

// BoundingBox::hit geom.rs:218-247 for two boxes at once, as a CONSERVATIVE test. NaN handling follows f32::min/max (= fminf/fmaxf).
// The reference computes (b - o) / d per plane. Here each plane costs one FMA: b * id + ood with id ~ 1/d (MUFU.RCP, <= 1 ulp) and
// ood ~ -(o * id), both per ray. Against the reference's value t the FMA result is off by at most 2.5 * 2^-23 * (|t| + |o * id|)
// (rounding of id, of o * id and of the FMA, plus the reference's own two roundings). Both terms are covered without extra
// instructions in the loop: the |o * id| part is folded into the per-ray constants PER AXIS -- the plane that is the near one for
// this ray's direction sign uses ood - e, the far one ood + e, with e = 2^-21 * |o * id| -- and the |t| part by widening the final
// interval by 2^-21 relatively. So the test never rejects a box the reference's test accepts, and nothing computed here reaches
// a hit record. (A single slack for all axes would be wrong in practice: a ray almost parallel to an axis has |o * id| ~ 1e30
// on that axis, and adding that to the other axes' intervals switches culling off for the whole ray.) A ray with d == 0 on an axis
// is culled on that axis by where its origin lies (slab_axis below).
struct SlabRay {
    V3 id;      // ~ 1 / direction
    V3 ood_lo;  // -(origin * id) -+ e: added to box.min * id
    V3 ood_hi;  // -(origin * id) +- e: added to box.max * id
};
__device__ __forceinline__ void slab2(const DNode& n, const SlabRay& s, float t_min, float t_max, bool& h0, bool& h1, float& n0, float& n1) {
    const float kWiden = 4.76837158203125e-7f;  // 2^-21
    float a0 = fmaf(n.xy0.x, s.id.x, s.ood_lo.x), b0 = fmaf(n.xy0.y, s.id.x, s.ood_hi.x);
    float a1 = fmaf(n.xy1.x, s.id.x, s.ood_lo.x), b1 = fmaf(n.xy1.y, s.id.x, s.ood_hi.x);
    float tn0 = fmaxf(fminf(a0, b0), t_min), tf0 = fminf(fmaxf(a0, b0), t_max);
    float tn1 = fmaxf(fminf(a1, b1), t_min), tf1 = fminf(fmaxf(a1, b1), t_max);
    a0 = fmaf(n.xy0.z, s.id.y, s.ood_lo.y); b0 = fmaf(n.xy0.w, s.id.y, s.ood_hi.y);
    a1 = fmaf(n.xy1.z, s.id.y, s.ood_lo.y); b1 = fmaf(n.xy1.w, s.id.y, s.ood_hi.y);
    tn0 = fmaxf(fminf(a0, b0), tn0); tf0 = fminf(fmaxf(a0, b0), tf0);
    tn1 = fmaxf(fminf(a1, b1), tn1); tf1 = fminf(fmaxf(a1, b1), tf1);
    a0 = fmaf(n.z01.x, s.id.z, s.ood_lo.z); b0 = fmaf(n.z01.y, s.id.z, s.ood_hi.z);
    a1 = fmaf(n.z01.z, s.id.z, s.ood_lo.z); b1 = fmaf(n.z01.w, s.id.z, s.ood_hi.z);
    tn0 = fmaxf(fminf(a0, b0), tn0); tf0 = fminf(fmaxf(a0, b0), tf0);
    tn1 = fmaxf(fminf(a1, b1), tn1); tf1 = fminf(fmaxf(a1, b1), tf1);
    h0 = fmaf(-fabsf(tn0), kWiden, tn0) <= fmaf(fabsf(tf0), kWiden, tf0);
    h1 = fmaf(-fabsf(tn1), kWiden, tn1) <= fmaf(fabsf(tf1), kWiden, tf1);
    n0 = tn0;
    n1 = tn1;
}

struct HitRec {
    float t;
    uint32_t prim;  // MRT_REF of the sphere / triangle / volume hit, kNone = miss
    uint32_t inst;  // instance index the triangle was reached through, kNone otherwise
};

__device__ __forceinline__ Ray to_instance_space(const DInstance& in, const Ray& w) {  // geom.rs:405-408
    if (in.flags & MRT_INSTANCE_IDENTITY) return w;                                   // Model::intersect :318-319: no transform
    return Ray{xform(in.inv0, in.inv1, in.inv2, w.o, 1.0f), xform(in.inv0, in.inv1, in.inv2, w.d, 0.0f)};
}

// Volume::intersect geom.rs:612-655 after the two target.intersect calls: [enter, exit] clipped to [t_min, t_max], then the free flight.
// xi = the free-flight uniform (f32::rand(), :638).
__device__ __forceinline__ bool volume_flight(const mrt_volume& vol, const Ray& r, float enter, float exit_, float t_min, float t_max, float xi, float& t_out) {
    if (enter < t_min) enter = t_min;
    if (exit_ > t_max) exit_ = t_max;
    if (enter >= exit_) return false;
    if (enter < 0.0f) enter = 0.0f;
    float ray_length = length(r.d);
    float inside = (exit_ - enter) * ray_length;
    float hit_distance = logf(xi) * vol.neg_inv_density;
    if (hit_distance > inside) return false;
    t_out = enter + hit_distance / ray_length;
    return true;
}
// ... for a sphere target
__device__ __forceinline__ bool volume_test(const DScene& sc, const mrt_volume& vol, const Ray& r, float t_min, float t_max, float xi, float& t_out) {
    const float inf = __int_as_float(0x7f800000);
    float4 s = __ldg(&sc.spheres[MRT_REF_INDEX(vol.target)]);
    float enter, exit_;
    if (!sphere_test(s, r, -inf, inf, enter)) return false;
    if (!sphere_test(s, r, enter + 0.0001f, inf, exit_)) return false;
    return volume_flight(vol, r, enter, exit_, t_min, t_max, xi, t_out);
}

// 1/x for the slab test only: one MUFU.RCP (<= 1 ulp, subnormals handled, 1/+-0 = +-inf). The slab test is widened by 4 ulp, which
// covers this error; nothing that reaches a hit record uses it.
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void slab_axis(float o, float d, float& id, float& ood_lo, float& ood_hi) {
    // d == 0 (1 / d = +-inf) cannot go through b * id + ood: b * inf - o * inf is NaN whenever b and o have the same sign, and the
    // reference's (b - o) / 0 = +-inf -- "inside this slab for all t, or for none" -- is lost. 1 / d is therefore capped at 2^64: the
    // planes then sit at (b - o) * 2^64 -+ e, beyond every distance of a scene on the side the sign of b - o says, so a ray parallel
    // to an axis is culled by whether its origin lies between the planes (to within the slack e, i.e. |o| 2^-21 in space). Capping a
    // nonzero |d| < 2^-64 shortens |t| and only moves planes towards the origin's side of e. o * id stays finite below |o| = 2^63
    // (coordinates no float scene has).
    // (First version: NaNs dropped by fminf / fmaxf next to a -inf made every box straddling 0 on that axis a miss -- found by the
    // per-pixel volume test on the centre row of an axis-aligned camera. Second version: such rays ignored the axis -- correct, but a
    // Lambertian bounce straight up from the ground, direction exactly (0, 1, 0), then walked 250,000 nodes of the 1 M-triangle
    // mesh: one ray, 0.28 s.)
    const float kCap = 1.8446744e19f;  // 2^64
    id = fmaxf(fminf(rcp_fast(d), kCap), -kCap);  // (a NaN becomes kCap; rays with a NaN direction never get here, trav_begin)
    const float p = o * id;
    const float e = fabsf(p) * 4.76837158203125e-7f;
    const float se = copysignf(e, id);  // d > 0: box.min is the near plane and gets -e
    ood_lo = -p - se;
    ood_hi = -p + se;
}
__device__ __forceinline__ SlabRay slab_ray(const Ray& r) {
    SlabRay s;
    slab_axis(r.o.x, r.d.x, s.id.x, s.ood_lo.x, s.ood_hi.x);
    slab_axis(r.o.y, r.d.y, s.id.y, s.ood_lo.y, s.ood_hi.y);
    slab_axis(r.o.z, r.d.z, s.id.z, s.ood_lo.z, s.ood_hi.z);
    return s;
}

// Closest hit over World.objects (world.rs:131-144) through the flattened TLAS / BLAS, as a resumable state machine whose
// three kinds of work are separate functions, so that a warp can run each kind with as many lanes as possible (k_extend):
//   trav_pop   next stack entry (leaving an instance restores the world ray)       -> T.ref = node | leaf, or finished
//   trav_node  ONE inner-node visit: fetch 64 B, slab-test both children            -> T.ref = nearer child | none
//   trav_leaf  one primitive: triangle / sphere / volume test, or instance entry    -> T.ref = none | BLAS root
// Differences from BvhNode::intersect (geom.rs:186-200), none of which can change a closest hit except on exact-t ties:
// iterative with an explicit per-thread stack, nearer child first, subtree skipped when its box lies beyond the current closest t.
// The world always has ONE root on the device (mrt_scene_upload puts a tree over a multi-object world list, world.rs:135-140).
struct Traversal {
    Ray r;              // ray in the current space (world, or the entered instance's)
    SlabRay s;          // its slab-test constants
    HitRec best;        // best.t doubles as the shrinking t_max (world.rs:133, geom.rs:192)
    uint32_t cur_inst;
    uint32_t ref;          // what this lane does next: a node ref, a leaf ref, or kNone = pop
    int sp;
    int inst_base;         // stack height when the current instance was entered (0 in world space)
};
// Where the world-space ray lives while the traversal is inside an instance (it is needed again when the instance is left):
// in registers (k_aov, k_finish) or in shared memory (k_extend, which is register-bound).
struct WorldRayRegs {
    Ray w;
    __device__ __forceinline__ Ray load() const { return w; }
    __device__ __forceinline__ void store(const Ray& r) { w = r; }
};
template <int STRIDE>
struct WorldRayShared {  // component-major: base[k * STRIDE] is component k of this thread's ray (conflict-free)
    float* base;
    __device__ __forceinline__ Ray load() const {
        return Ray{V3{base[0], base[STRIDE], base[2 * STRIDE]}, V3{base[3 * STRIDE], base[4 * STRIDE], base[5 * STRIDE]}};
    }
    __device__ __forceinline__ void store(const Ray& r) {
        base[0] = r.o.x; base[STRIDE] = r.o.y; base[2 * STRIDE] = r.o.z;
        base[3 * STRIDE] = r.d.x; base[4 * STRIDE] = r.d.y; base[5 * STRIDE] = r.d.z;
    }
};
__device__ __forceinline__ bool ref_is_node(uint32_t ref) { return ref < (1u << 29); }  // kind 0 in the top three bits

template <class W>
__device__ __forceinline__ void trav_begin(const DScene& sc, Traversal& T, W& ws, const Ray& world, float t_max) {
    ws.store(world);
    T.sp = 0;
    T.best = HitRec{t_max, kNone, kNone};
    T.r = world;
    T.s = slab_ray(world);
    T.cur_inst = kNone;
    T.ref = sc.root;  // kNone for an empty world
    T.inst_base = 0;
    // A ray whose direction has no length or is not finite, or whose origin is not finite, misses everything. (In the reference NaN compares false everywhere: such a
    // ray passes every box test and "hits" the first sphere it meets at t = NaN, whichever that is in its tree, and keeps bouncing
    // as a NaN ray until the depth limit. Here it would walk the whole scene once per bounce -- measured: one such path, 0.47 s.)
    const float len2 = length_squared(world.d), oabs = fabsf(world.o.x) + fabsf(world.o.y) + fabsf(world.o.z);
    if (!(len2 > 0.0f) || !(len2 < __int_as_float(0x7f800000)) || !(oabs < __int_as_float(0x7f800000))) T.ref = kNone;
}

// false when the traversal is complete (T.best is final). Entries pushed before an instance was entered sit below
// T.inst_base: popping one of them means the BLAS is exhausted, i.e. Instance::intersect has returned (geom.rs:404-420).
template <class W>
__device__ __forceinline__ bool trav_pop(Traversal& T, const uint32_t* stack, const W& ws) {
    if (T.sp == 0) return false;
    T.ref = stack[--T.sp];
    if (T.sp < T.inst_base) {
        T.r = ws.load();
        T.s = slab_ray(T.r);
        T.cur_inst = kNone;
        T.inst_base = 0;
    }
    return true;
}

template <bool COUNT>
__device__ __forceinline__ void trav_node(const DScene& sc, Traversal& T, uint32_t* stack, float t_min, VisitCounters* cnt) {
    const DNode* np = &sc.nodes[MRT_REF_INDEX(T.ref)];
    DNode n;
    n.xy0 = __ldg(&np->xy0);
    n.xy1 = __ldg(&np->xy1);
    n.z01 = __ldg(&np->z01);
    uint2 ch = __ldg(reinterpret_cast<const uint2*>(&np->child0));
    if (COUNT) cnt->node_visits++;
    bool h0, h1;
    float n0, n1;
    slab2(n, T.s, t_min, T.best.t, h0, h1, n0, n1);
    h0 = h0 && ch.x != kNone;
    h1 = h1 && ch.y != kNone;
    if (h0 && h1) {
        bool swap = n1 < n0;
        stack[T.sp++] = swap ? ch.x : ch.y;
        T.ref = swap ? ch.y : ch.x;
    } else {
        T.ref = h0 ? ch.x : (h1 ? ch.y : kNone);
    }
}

// Material::alpha_test of the triangle's OWN material at the candidate hit (geom.rs:567-571); defined after the surfaces below
__device__ bool triangle_alpha_test(const DScene& sc, const DTriVerts& tv, uint32_t tri_dev, const Ray& r, float t, const RngKey& key);

// Closest triangle of one BLAS in [t_min, t_max] -- the smallest t, which may be negative: what Model / Instance::intersect returns to
// Volume::intersect (geom.rs:613-619, :318-328, :404-420). `stk` is scratch for at least the BLAS's depth (the free part of the
// caller's traversal stack).
template <bool COUNT, bool ALPHA>
__device__ __noinline__ bool blas_nearest(const DScene& sc, uint32_t root, const Ray& r, float t_min, float t_max, uint32_t* stk, const RngKey& key, VisitCounters* cnt,
                                          float& t_out) {
    const SlabRay s = slab_ray(r);
    float best = t_max;
    bool found = false;
    int sp = 0;
    uint32_t ref = root;
    for (;;) {
        if (ref == kNone) {
            if (sp == 0) break;
            ref = stk[--sp];
        }
        if (ref_is_node(ref)) {
            const DNode* np = &sc.nodes[MRT_REF_INDEX(ref)];
            DNode n;
            n.xy0 = __ldg(&np->xy0);
            n.xy1 = __ldg(&np->xy1);
            n.z01 = __ldg(&np->z01);
            const uint2 ch = __ldg(reinterpret_cast<const uint2*>(&np->child0));
            if (COUNT) cnt->node_visits++;
            bool h0, h1;
            float n0, n1;
            slab2(n, s, t_min, best, h0, h1, n0, n1);
            h0 = h0 && ch.x != kNone;
            h1 = h1 && ch.y != kNone;
            if (h0 && h1) {
                const bool swap = n1 < n0;
                stk[sp++] = swap ? ch.x : ch.y;
                ref = swap ? ch.y : ch.x;
            } else {
                ref = h0 ? ch.x : (h1 ? ch.y : kNone);
            }
        } else {  // a triangle leaf (a BLAS holds nothing else)
            const uint32_t first = ref & kTriIndexMask, count = ((ref >> 27) & 3u) + 1u;
            for (uint32_t k = 0; k < count; ++k) {
                if (COUNT) cnt->tri_tests++;
                const DTriVerts* tp = &sc.tri_verts[first + k];
                DTriVerts tv;
                tv.a = __ldg(&tp->a);
                tv.b = __ldg(&tp->b);
                tv.c = __ldg(&tp->c);
                float t;
                if (triangle_test(tv, r, t_min, best, t)) {
                    if (ALPHA && (__float_as_uint(tv.a.w) & kTriAlphaFlag) && !triangle_alpha_test(sc, tv, first + k, r, t, key)) continue;
                    best = t;
                    found = true;
                }
            }
            ref = kNone;
        }
    }
    t_out = best;
    return found;
}
// Volume::intersect for a Model / Instance target: the medium fills the mesh between its first crossing of the ray's line and the
// next one (geom.rs:613-619: convex targets are what the construction is meant for)
template <bool COUNT, bool ALPHA>
__device__ __forceinline__ bool volume_test_mesh(const DScene& sc, const mrt_volume& vol, const Ray& r, float t_min, float t_max, float xi, uint32_t* stk,
                                                 const RngKey& key, VisitCounters* cnt, float& t_out) {
    const float inf = __int_as_float(0x7f800000);
    const DInstance* ip = &sc.instances[MRT_REF_INDEX(vol.target)];
    DInstance in;
    in.inv0 = __ldg(&ip->inv0);
    in.inv1 = __ldg(&ip->inv1);
    in.inv2 = __ldg(&ip->inv2);
    const uint4 meta = __ldg(reinterpret_cast<const uint4*>(&ip->root));
    in.flags = meta.z;
    const Ray local = to_instance_space(in, r);  // t is the same along both rays (geom.rs:405-408: the direction is not normalised)
    float enter, exit_;
    if (!blas_nearest<COUNT, ALPHA>(sc, meta.x, local, -inf, inf, stk, key, cnt, enter)) return false;
    if (!blas_nearest<COUNT, ALPHA>(sc, meta.x, local, enter + 0.0001f, inf, stk, key, cnt, exit_)) return false;
    return volume_flight(vol, r, enter, exit_, t_min, t_max, xi, t_out);
}

// ALPHA = the scene has alpha-tested triangles, VOLUME = it has volumes: both draw random numbers during intersection (geom.rs:568,
// :638) and so need the path's RNG key; scenes without them run the variant that carries no key and has neither code path.
template <bool COUNT, bool ALPHA, bool VOLUME, class W>
__device__ __forceinline__ void trav_leaf(const DScene& sc, Traversal& T, uint32_t* stack, const W& ws, float t_min, const RngKey& key, VisitCounters* cnt) {
    const uint32_t ref = T.ref;
    const uint32_t idx = MRT_REF_INDEX(ref);
    T.ref = kNone;
    switch (MRT_REF_KIND(ref)) {
        case MRT_PRIM_TRIANGLE: {
            const uint32_t first = ref & kTriIndexMask, count = ((ref >> 27) & 3u) + 1u;
            for (uint32_t k = 0; k < count; ++k) {
                if (COUNT) cnt->tri_tests++;
                const DTriVerts* tp = &sc.tri_verts[first + k];
                DTriVerts tv;
                tv.a = __ldg(&tp->a);
                tv.b = __ldg(&tp->b);
                tv.c = __ldg(&tp->c);
                float t;
                if (triangle_test(tv, T.r, t_min, T.best.t, t)) {
                    if (ALPHA && (__float_as_uint(tv.a.w) & kTriAlphaFlag) && !triangle_alpha_test(sc, tv, first + k, T.r, t, key)) continue;  // geom.rs:567-571
                    T.best = HitRec{t, MRT_REF(MRT_PRIM_TRIANGLE, first + k), T.cur_inst};
                }
            }
            break;
        }
        case MRT_PRIM_SPHERE: {
            if (COUNT) cnt->sphere_tests++;
            float t;
            if (sphere_test(__ldg(&sc.spheres[idx]), T.r, t_min, T.best.t, t)) T.best = HitRec{t, ref, kNone};
            break;
        }
        case MRT_PRIM_INSTANCE: {
            if (COUNT) cnt->instance_tests++;
            const DInstance* ip = &sc.instances[idx];
            DInstance in;
            in.inv0 = __ldg(&ip->inv0);
            in.inv1 = __ldg(&ip->inv1);
            in.inv2 = __ldg(&ip->inv2);
            uint4 meta = __ldg(reinterpret_cast<const uint4*>(&ip->root));
            in.flags = meta.z;
            T.r = to_instance_space(in, T.r);  // only reachable from world space (no nested instances, geom.rs:336): T.r is the world ray
            T.s = slab_ray(T.r);
            T.cur_inst = idx;
            T.inst_base = T.sp;
            T.ref = meta.x;  // BLAS root: a node
            break;
        }
        case MRT_PRIM_VOLUME: {
            if (VOLUME) {
                if (COUNT) cnt->volume_tests++;
                mrt_volume vol = sc.volumes[idx];
                Rand4 xi = draw4(key, kStreamVolume | (idx & kStreamIndexMask));
                float t;
                bool hit;
                if (!ALPHA || MRT_REF_KIND(vol.target) == MRT_PRIM_SPHERE) hit = volume_test(sc, vol, T.r, t_min, T.best.t, xi.x, t);
                else hit = volume_test_mesh<COUNT, ALPHA>(sc, vol, T.r, t_min, T.best.t, xi.x, stack + T.sp, key, cnt, t);  // (scenes with such volumes run the ALPHA variant)
                if (hit) T.best = HitRec{t, ref, kNone};
            }
            break;
        }
        default: break;
    }
}

// run a traversal to completion (AOV pass, drain)
template <bool COUNT, bool ALPHA, bool VOLUME>
__device__ __forceinline__ HitRec traverse(const DScene& sc, const Ray& world, float t_min, float t_max, const RngKey& key, VisitCounters* cnt) {
    uint32_t stack[kStackSize];
    Traversal T;
    WorldRayRegs ws;
    trav_begin(sc, T, ws, world, t_max);
    do {
        while (T.ref != kNone) {
            if (ref_is_node(T.ref)) trav_node<COUNT>(sc, T, stack, t_min, cnt);
            else trav_leaf<COUNT, ALPHA, VOLUME>(sc, T, stack, ws, t_min, key, cnt);
        }
    } while (trav_pop(T, stack, ws));
    if (T.best.prim == kNone) T.best.t = t_max;
    return T.best;
}

// ---- surfaces (texture.rs) -------------------------------------------------------------------------------------
struct V4 {
    float x, y, z, w;
};
__device__ __forceinline__ V4 v4(float4 a) { return V4{a.x, a.y, a.z, a.w}; }
__device__ __forceinline__ V4 operator*(V4 a, float s) { return V4{a.x * s, a.y * s, a.z * s, a.w * s}; }
__device__ __forceinline__ V4 operator+(V4 a, V4 b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
__device__ __forceinline__ V4 operator-(V4 a, V4 b) { return V4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
__device__ __forceinline__ float fractf_(float x) { return x - truncf(x); }
__device__ __forceinline__ float wrap1(int mode, float x) {  // WrapMode::wrap texture.rs:277-300
    if (mode == MRT_WRAP_REPEAT) {
        x = (x < 0.0f) ? 1.0f - fractf_(fabsf(x)) : x;
        x = (x > 1.0f) ? fractf_(x) : x;
        return x;
    }
    return fmaxf(fminf(x, 1.0f), 0.0f);
}
__device__ __forceinline__ uint32_t as_index(float v, uint32_t limit) {  // Rust `as usize` saturates; the reference would panic past the edge
    if (!(v > 0.0f)) return 0;
    uint32_t i = (v >= 4294967040.0f) ? 0xFFFFFFFFu : (uint32_t)v;
    return i < limit ? i : limit - 1;
}
__device__ __forceinline__ V4 texture_get(const DScene& sc, int tex, float u, float v) {  // Texture::get_f texture.rs:126-148
    mrt_texture t = sc.textures[tex];
    float x = wrap1(t.wrap, u) * (float)(t.width - 1);
    float y = wrap1(t.wrap, v) * (float)(t.height - 1);
    uint32_t x0 = as_index(floorf(x), t.width), x1 = as_index(ceilf(x), t.width);
    uint32_t y0 = as_index(floorf(y), t.height), y1 = as_index(ceilf(y), t.height);
    const float4* px = sc.texels + t.texel_offset;
    float tx = x - (float)x0;
    V4 p0 = v4(__ldg(&px[(size_t)y0 * t.width + x0])) * (1.0f - tx) + v4(__ldg(&px[(size_t)y0 * t.width + x1])) * tx;
    V4 p1 = v4(__ldg(&px[(size_t)y1 * t.width + x0])) * (1.0f - tx) + v4(__ldg(&px[(size_t)y1 * t.width + x1])) * tx;
    float ty = y - (float)y0;
    return p1 * ty + p0 * (1.0f - ty);
}
__device__ __noinline__ V4 surface_get_slow(const DScene& sc, int surface, float u, float v);
__device__ __forceinline__ V4 surface_get(const DScene& sc, int surface, float u, float v) {  // Surface::get_f
    mrt_surface s = sc.surfaces[surface];
    if (s.kind == MRT_SURF_SOLID) return V4{s.color[0], s.color[1], s.color[2], s.color[3]};  // texture.rs:191-193
    if (s.kind == MRT_SURF_TEXTURE) return texture_get(sc, s.a, u, v);
    return surface_get_slow(sc, surface, u, v);
}
__device__ __noinline__ V4 surface_get_slow(const DScene& sc, int surface, float u, float v) {
    mrt_surface s = sc.surfaces[surface];
    switch (s.kind) {
        case MRT_SURF_YCBCR: {  // texture.rs:233-247 (BT.709 YCbCr -> RGB, clamp, powf 2.2)
            const float KR = 0.2126f, KG = 0.7152f, KB = 0.0722f;
            V4 l = texture_get(sc, s.a, u, v), c = texture_get(sc, s.b, u, v);
            float Y = l.x, cb = c.x - 0.5f, cr = c.y - 0.5f;
            float m11 = -(KB / KG) * (2.0f - 2.0f * KB), m12 = 2.0f - 2.0f * KB;
            float m20 = 2.0f - 2.0f * KR, m21 = -(KR / KG) * (2.0f - 2.0f * KR);
            float rr = ((1.0f * Y + 0.0f * cb) + m20 * cr) + 0.0f * 1.0f;
            float gg = ((1.0f * Y + m11 * cb) + m21 * cr) + 0.0f * 1.0f;
            float bb = ((1.0f * Y + m12 * cb) + 0.0f * cr) + 0.0f * 1.0f;
            rr = powf(fmaxf(fminf(rr, 1.0f), 0.0f), 2.2f);
            gg = powf(fmaxf(fminf(gg, 1.0f), 0.0f), 2.2f);
            bb = powf(fmaxf(fminf(bb, 1.0f), 0.0f), 2.2f);
            return V4{rr, gg, bb, 1.0f};
        }
        case MRT_SURF_BLEND: {  // texture.rs:250-334
            V4 l = surface_get(sc, s.a, u, v), r = surface_get(sc, s.b, u, v);
            switch (s.mode) {
                case MRT_BLEND_LIGHTEN: return V4{fmaxf(l.x, r.x), fmaxf(l.y, r.y), fmaxf(l.z, r.z), fmaxf(l.w, r.w)};
                case MRT_BLEND_DARKEN: return V4{fminf(l.x, r.x), fminf(l.y, r.y), fminf(l.z, r.z), fminf(l.w, r.w)};
                case MRT_BLEND_ADDITION: { V4 a = l + r; return V4{fminf(a.x, 1.0f), fminf(a.y, 1.0f), fminf(a.z, 1.0f), fminf(a.w, 1.0f)}; }
                default: { V4 a = l - r; return V4{fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f), fmaxf(a.z, 0.0f), fmaxf(a.w, 0.0f)}; }
            }
        }
        case MRT_SURF_FALLBACK: {  // texture.rs:356-359
            V4 c = surface_get(sc, s.a, u, v);
            return (V4{s.color[0], s.color[1], s.color[2], s.color[3]} * (1.0f - c.w)) + (c * c.w);
        }
        default: return V4{0, 0, 0, 0};
    }
}

// ---- EveMaterial (eve.rs:43-133): texture-driven Mix(Lambertian, Specular 1.8) with emission and tangent-space normals ------------
__device__ __forceinline__ int eve_palette_surface(const mrt_material& m) { return __float_as_int(m.p[0]); }
__device__ __noinline__ V3 eve_normal(const DScene& sc, const mrt_material& m, float u, float v) {  // normal_occlusion :66-73, normal :130-133
    V4 px = surface_get(sc, m.surface, u, v);
    float y = px.y * 2.0f - 1.0f, w = px.w * 2.0f - 1.0f;
    float x = (1.0f - y * y) - w * w;
    return unit(V3{y, w, sqrtf(fabsf(x))});
}
__device__ __forceinline__ V3 eve_palette(const DScene& sc, const mrt_material& m, float i) {  // EveMaterialColor::get :190-199
    i = i * 3.0f;
    uint32_t i0 = as_index(floorf(i), 4u), i1 = as_index(ceilf(i), 4u);
    float t = i - (float)i0;
    const mrt_surface &a = sc.surfaces[eve_palette_surface(m) + (int)i0], &b = sc.surfaces[eve_palette_surface(m) + (int)i1];
    return (V3{a.color[0], a.color[1], a.color[2]} * (1.0f - t)) + (V3{b.color[0], b.color[1], b.color[2]} * t);
}
__device__ __noinline__ V3 eve_emission(const DScene& sc, const mrt_material& m, float u, float v) {  // emit :121-128: glow colour * glow mask * 10
    const int p = eve_palette_surface(m);
    return V3{sc.surfaces[p].color[3], sc.surfaces[p + 1].color[3], sc.surfaces[p + 2].color[3]} * surface_get(sc, m.right, u, v).w * 10.0f;
}
// colour and Mix ratio of the Lambertian / Specular pair the material builds per hit (scatter :91-119)
__device__ __noinline__ void eve_surface(const DScene& sc, const mrt_material& m, float u, float v, V3& color, float& ratio) {
    V4 ar = surface_get(sc, m.left, u, v);
    V4 k = surface_get(sc, m.right, u, v);
    V3 albedo{ar.x, ar.y, ar.z};
    float paint = k.x, dirt = k.z * 1.0f;
    V3 material_color = eve_palette(sc, m, k.y);
    color = (((albedo * material_color * (1.0f - paint)) + (albedo * paint)) * (1.0f - fminf(dirt, 1.0f))) + (V3{0.01f, 0.005f, 0.0f} * dirt);
    ratio = fminf(ar.w + dirt, 1.0f);
}

// Triangle::intersect's alpha test (geom.rs:535-571): area barycentrics -> uv -> alpha_test of the triangle's own material
// (Lambertian / Metal / Specular: surface alpha != 0, material.rs:222-224, 281-283, 380-382; Mix: a coin flip then the child,
// :419-425; every other material: the default `true`, :24-26).
__device__ __noinline__ bool triangle_alpha_test(const DScene& sc, const DTriVerts& tv, uint32_t tri_dev, const Ray& r, float t, const RngKey& key) {
    const mrt_tri_shading& sh = sc.tri_shading[sc.tri_map[tri_dev]];
    V3 va = v3(tv.a), vb = v3(tv.b), vc = v3(tv.c);
    V3 point = ray_at(r, t);
    V3 d0 = va - point, d1 = vb - point, d2 = vc - point;
    float area = length(cross(va - vb, va - vc));
    float a0 = length(cross(d1, d2)) / area;
    float a1 = length(cross(d2, d0)) / area;
    float a2 = length(cross(d0, d1)) / area;
    float u = (sh.uv[0] * a0 + sh.uv[2] * a1) + sh.uv[4] * a2;
    float v = (sh.uv[1] * a0 + sh.uv[3] * a1) + sh.uv[5] * a2;
    mrt_material mat = sc.materials[sh.material];
    uint32_t level = 0;
    while (mat.kind == MRT_MAT_MIX && level < 16) {
        Rand4 c = draw4(key, kStreamAlpha | ((tri_dev & 0xFFFFFFu) << 4) | level);  // level < 16
        mat = sc.materials[(c.x < mat.p[0]) ? mat.left : mat.right];
        ++level;
    }
    if (mat.kind == MRT_MAT_LAMBERTIAN || mat.kind == MRT_MAT_METAL || mat.kind == MRT_MAT_SPECULAR) return surface_get(sc, mat.surface, u, v).w != 0.0f;
    return true;
}

// ---- closest-hit attributes (the part of Hit the reference fills for every candidate; here once) -----------------
struct Surfel {
    V3 point, normal;
    float u, v;
    bool has_uv, front_face;
    int32_t material;
    uint32_t object_id, tri_id;
};
__device__ __forceinline__ void set_face_normal(Surfel& s, const Ray& r, V3 outward) {  // geom.rs:17-24
    s.front_face = dot(r.d, outward) < 0.0f;
    s.normal = s.front_face ? outward : -outward;
}
// Material resolution N2: Triangle.material -> Model.material -> Instance.material (geom.rs:579, 321-323, 413-415)
__device__ __forceinline__ int32_t hit_material(const DScene& sc, const HitRec& h) {
    const uint32_t idx = MRT_REF_INDEX(h.prim);
    switch (MRT_REF_KIND(h.prim)) {
        case MRT_PRIM_SPHERE: return sc.sphere_aux[idx].material;
        case MRT_PRIM_VOLUME: return sc.volumes[idx].material;
        case MRT_PRIM_TRIANGLE: {
            int32_t m = (h.inst != kNone) ? sc.instances[h.inst].material : -1;
            return m >= 0 ? m : sc.tri_shading[sc.tri_map[h.prim & kTriIndexMask]].material;
        }
        default: return -1;
    }
}
// FULL: the scene needs the rarely used code (an EveMaterial somewhere): compiled out of the kernels every other scene runs
template <bool FULL>
__device__ __forceinline__ Surfel resolve_hit(const DScene& sc, const Ray& world, const HitRec& h, int32_t material) {
    Surfel s;
    s.has_uv = false;
    s.u = s.v = 0.0f;
    s.material = material;
    s.tri_id = kNone;
    const uint32_t idx = MRT_REF_INDEX(h.prim);
    switch (MRT_REF_KIND(h.prim)) {
        case MRT_PRIM_SPHERE: {  // geom.rs:77-91
            float4 sp = __ldg(&sc.spheres[idx]);
            s.point = ray_at(world, h.t);
            V3 n = (s.point - v3(sp)) / sp.w;
            set_face_normal(s, world, n);
            s.object_id = sc.sphere_aux[idx].object_id;
            break;
        }
        case MRT_PRIM_VOLUME: {  // geom.rs:644-651
            s.point = ray_at(world, h.t);
            s.normal = V3{1.0f, 0.0f, 0.0f};
            s.front_face = true;
            s.object_id = sc.volumes[idx].object_id;
            break;
        }
        default: {  // triangle, geom.rs:535-577, reached through Instance::intersect :404-420 or Model::intersect :318-328
            const DInstance* ip = &sc.instances[h.inst];
            DInstance in;
            in.inv0 = __ldg(&ip->inv0); in.inv1 = __ldg(&ip->inv1); in.inv2 = __ldg(&ip->inv2);
            in.fwd0 = __ldg(&ip->fwd0); in.fwd1 = __ldg(&ip->fwd1); in.fwd2 = __ldg(&ip->fwd2);
            uint4 meta = __ldg(reinterpret_cast<const uint4*>(&ip->root));
            in.flags = meta.z;
            Ray r = to_instance_space(in, world);
            const uint32_t tri_dev = h.prim & kTriIndexMask, tri = sc.tri_map[tri_dev];
            const DTriVerts* tp = &sc.tri_verts[tri_dev];
            V3 va = v3(__ldg(&tp->a)), vb = v3(__ldg(&tp->b)), vc = v3(__ldg(&tp->c));
            const mrt_tri_shading& sh = sc.tri_shading[tri];
            V3 point = ray_at(r, h.t);
            V3 d0 = va - point, d1 = vb - point, d2 = vc - point;
            float area = length(cross(va - vb, va - vc));
            float a0 = length(cross(d1, d2)) / area;
            float a1 = length(cross(d2, d0)) / area;
            float a2 = length(cross(d0, d1)) / area;
            V3 na = v3(sh.normal[0], sh.normal[1], sh.normal[2]), nb = v3(sh.normal[3], sh.normal[4], sh.normal[5]), nc = v3(sh.normal[6], sh.normal[7], sh.normal[8]);
            V3 normal = na * a0 + nb * a1 + nc * a2;
            if (sh.flags & MRT_TRI_HAS_UV) {
                s.has_uv = true;
                s.u = (sh.uv[0] * a0 + sh.uv[2] * a1) + sh.uv[4] * a2;
                s.v = (sh.uv[1] * a0 + sh.uv[3] * a1) + sh.uv[5] * a2;
                // Material::normal of the triangle's OWN material (geom.rs:551-560); None for every material of material.rs (:21-23)
                if (FULL && sc.materials[sh.material].kind == MRT_MAT_EVE) {
                    const mrt_material own = sc.materials[sh.material];
                    V3 tn = eve_normal(sc, own, s.u, s.v);
                    normal = (v3(sh.tangent[0], sh.tangent[1], sh.tangent[2]) * tn.x + v3(sh.bitangent[0], sh.bitangent[1], sh.bitangent[2]) * tn.y) + normal * tn.z;
                }
            }
            set_face_normal(s, r, normal);
            if (in.flags & MRT_INSTANCE_IDENTITY) {
                s.point = point;
            } else {  // geom.rs:411-412: forward matrix for both, normal re-normalised
                s.point = xform(in.fwd0, in.fwd1, in.fwd2, point, 1.0f);
                s.normal = unit(xform(in.fwd0, in.fwd1, in.fwd2, s.normal, 0.0f));
            }
            s.object_id = meta.w;
            s.tri_id = tri - sc.blas[ip->pad[0]].first_tri;
            break;
        }
    }
    return s;
}

// ---- backgrounds (material.rs:39-190) ---------------------------------------------------------------------------
__device__ __forceinline__ V3 background(const DScene& sc, const Ray& r) {
    const mrt_background& bg = sc.bg;
    switch (bg.kind) {
        case MRT_BG_SKY: {  // :57-62
            V3 ud = unit(r.d);
            float t = 0.5f * (ud.y + 1.0f);
            return (V3{1.0f, 1.0f, 1.0f} * (1.0f - t)) + (V3{0.5f, 0.7f, 1.0f} * t);
        }
        case MRT_BG_SKYSPHERE: {  // :75-88
            const float PI = 3.14159265358979323846f;
            V3 p = unit(r.d);
            float theta = acosf(p.y);
            float phi = atan2f(p.z * -1.0f, p.x) + PI;
            V4 c = surface_get(sc, bg.surface[0], phi / (2.0f * PI), theta / PI);
            return V3{c.x, c.y, c.z};
        }
        case MRT_BG_CUBEMAP: {  // :123-189
            const float* m = bg.transform;
            V3 d = r.d;
            V3 p{((m[0] * d.x + m[4] * d.y) + m[8] * d.z) + m[12] * 0.0f, ((m[1] * d.x + m[5] * d.y) + m[9] * d.z) + m[13] * 0.0f,
                 ((m[2] * d.x + m[6] * d.y) + m[10] * d.z) + m[14] * 0.0f};
            V3 a{fabsf(p.x), fabsf(p.y), fabsf(p.z)};
            bool xl = a.x >= a.y && a.x >= a.z, yl = a.y >= a.x && a.y >= a.z, zl = a.z >= a.x && a.z >= a.y;
            int index = 0;
            float max_axis = 0.0f, u = 0.0f, v = 0.0f;
            if (xl) {
                if (p.x > 0.0f) { index = 0; u = p.z * -1.0f; v = p.y; } else { index = 1; u = p.z; v = p.y; }
                max_axis = a.x;
            } else if (yl) {
                if (p.y > 0.0f) { index = 3; u = p.x; v = p.z * -1.0f; } else { index = 2; u = p.x; v = p.z; }
                max_axis = a.y;
            } else if (zl) {
                if (p.z > 0.0f) { index = 4; u = p.x; v = p.y; } else { index = 5; u = p.x * -1.0f; v = p.y; }
                max_axis = a.z;
            }
            V4 c = surface_get(sc, bg.surface[index], 0.5f * (u / max_axis + 1.0f), 0.5f * (v / max_axis + 1.0f));
            return V3{c.x, c.y, c.z};
        }
        default: return V3{bg.color[0], bg.color[1], bg.color[2]};
    }
}

// ---- materials (material.rs) --------------------------------------------------------------------------------------
__device__ __forceinline__ float reflectance(float cosine, float ref_idx) {  // :296-299
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    float x = 1.0f - cosine;
    float x2 = x * x;
    return r0 + (1.0f - r0) * (x2 * x2 * x);
}
struct ScatterOut {
    bool scattered;
    V3 attenuation, dir;
    V3 emitted;
};
// Mix (material.rs:391-426): one independent coin per call and per nesting level, then the chosen child
__device__ __forceinline__ mrt_material pick_material(const DScene& sc, int32_t m, const RngKey& key, uint32_t stream) {
    mrt_material mat = sc.materials[m];
    uint32_t level = 0;
    while (mat.kind == MRT_MAT_MIX && level < 16) {
        Rand4 c = draw4(key, stream + level);
        mat = sc.materials[(c.x < mat.p[0]) ? mat.left : mat.right];
        ++level;
    }
    return mat;
}
__device__ __forceinline__ V3 surface_rgb(const DScene& sc, int surface, const Surfel& s) {  // get_f(hit.uv.unwrap_or(0)).contract()
    V4 c = surface_get(sc, surface, s.has_uv ? s.u : 0.0f, s.has_uv ? s.v : 0.0f);
    return V3{c.x, c.y, c.z};
}
__device__ __forceinline__ void lambertian_scatter(const DScene& sc, const mrt_material& mat, const Surfel& s, float xa, float xb, ScatterOut& out) {  // :205-220
    V3 dir = s.normal + sample_unit_vector(xa, xb);
    if (near_zero(dir)) dir = s.normal;
    out.scattered = true;
    out.dir = dir;
    out.attenuation = surface_rgb(sc, mat.surface, s);
}
__device__ __forceinline__ void lambertian_scatter_color(const Surfel& s, V3 color, float xa, float xb, ScatterOut& out) {
    V3 dir = s.normal + sample_unit_vector(xa, xb);
    if (near_zero(dir)) dir = s.normal;
    out.scattered = true;
    out.dir = dir;
    out.attenuation = color;
}
// EveMaterial::scatter eve.rs:91-119: Mix(min(roughness + dirt, 1), Lambertian(color), Specular(1.8, color)).scatter -- the coin is xi.w
__device__ __noinline__ void eve_scatter(const DScene& sc, const mrt_material& mat, const Ray& ray, const Surfel& s, const Rand4& xi, ScatterOut& out) {
    if (!s.has_uv) return;
    V3 color;
    float mix_ratio;
    eve_surface(sc, mat, s.u, s.v, color, mix_ratio);
    if (xi.w < mix_ratio) {
        lambertian_scatter_color(s, color, xi.x, xi.y, out);
    } else {  // Specular::scatter material.rs:352-378
        float ratio = s.front_face ? 1.0f / 1.8f : 1.8f;
        V3 ud = unit(ray.d);
        float cos_theta = fminf(dot(-ud, s.normal), 1.0f);
        float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        if (ratio * sin_theta > 1.0f || reflectance(cos_theta, ratio) > xi.x) {
            out.scattered = true;
            out.dir = reflect(ud, s.normal);
            out.attenuation = V3{1.0f, 1.0f, 1.0f};
        } else {
            lambertian_scatter_color(s, color, xi.y, xi.z, out);
        }
    }
}
// Hit::emit + Hit::scatter (geom.rs:26-32) for a resolved material kind (never MIX)
template <bool FULL>
__device__ __forceinline__ void scatter_kind(const DScene& sc, const mrt_material& mat, const Ray& ray, const Surfel& s, const Rand4& xi, ScatterOut& out) {
    out.scattered = false;
    switch (mat.kind) {
        case MRT_MAT_LAMBERTIAN: lambertian_scatter(sc, mat, s, xi.x, xi.y, out); break;
        case MRT_MAT_METAL: {  // :261-279
            V3 reflected = reflect(unit(ray.d), s.normal);
            V3 dir = reflected + (sample_unit_ball(xi.x, xi.y, xi.z) * mat.p[0]);
            if (dot(dir, s.normal) > 0.0f) {
                out.scattered = true;
                out.dir = dir;
                out.attenuation = surface_rgb(sc, mat.surface, s);
            }
            break;
        }
        case MRT_MAT_DIELECTRIC:
        case MRT_MAT_SPECULAR: {  // :302-328, :352-378
            float ratio = s.front_face ? 1.0f / mat.p[0] : mat.p[0];
            V3 ud = unit(ray.d);
            float cos_theta = fminf(dot(-ud, s.normal), 1.0f);
            float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
            bool cannot_refract = ratio * sin_theta > 1.0f;
            if (cannot_refract || reflectance(cos_theta, ratio) > xi.x) {
                out.scattered = true;
                out.dir = reflect(ud, s.normal);
                out.attenuation = V3{1.0f, 1.0f, 1.0f};
            } else if (mat.kind == MRT_MAT_DIELECTRIC) {
                out.scattered = true;
                out.dir = refract(ud, s.normal, ratio);
                out.attenuation = V3{1.0f, 1.0f, 1.0f};
            } else {
                lambertian_scatter(sc, mat, s, xi.y, xi.z, out);  // Specular delegates to its inner Lambertian
            }
            break;
        }
        case MRT_MAT_EVE:
            if (FULL) eve_scatter(sc, mat, ray, s, xi, out);
            break;
        case MRT_MAT_ISOTROPIC:  // :439-444
            out.scattered = true;
            out.dir = sample_unit_ball(xi.x, xi.y, xi.z);
            out.attenuation = V3{mat.p[0], mat.p[1], mat.p[2]};
            break;
        default: break;  // ABSORB (:385-389), DIFFUSE_LIGHT (:239-241): no scatter
    }
}

}  // namespace mrt

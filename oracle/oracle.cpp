// oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT (see oracle.h for the usage rule).
//
// A C++17 restatement of the reference's CPU path tracer, function by function, each citing the
// reference file:line it follows (paths relative to /root/reference/src).  Same recursive pointer BVH,
// same random-axis median split, same always-left-then-right traversal, same division-based slab test,
// same Moller-Trumbore + area-barycentric triangle test, same recursive trace, same rejection samplers.
// All arithmetic is IEEE binary32 in the reference's operation order; build with
//   -O2 -ffp-contract=off -fno-fast-math
// because rustc never contracts a*b+c into an FMA.
//
// PARITY UNPINNED (oracle.h): no reference test/golden vector exists for this path and the reference
// cannot be compiled here; hand-derived known-answer tests pin it instead.  The RNG (`fastrand` 1.4.1,
// Cargo.lock:579-582, source absent) is restated from its published WyRand algorithm.
//
// Additions the reference does not have (needed by the parity harness): primitive ids on Hit, bounded
// spp, seedable per-frame RNG streams, visit counters.
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <memory>
#include <mutex>
#include <optional>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {

typedef float F;  // math.rs:12
const F PI = 3.14159265358979323846f;  // std::f32::consts::PI
const F INF = std::numeric_limits<float>::infinity();
const uint32_t NONE_ID = 0xFFFFFFFFu;

// f32::min / f32::max return the non-NaN operand (math.rs:248-254) == fminf / fmaxf
inline F fmin_(F a, F b) { return std::fmin(a, b); }
inline F fmax_(F a, F b) { return std::fmax(a, b); }

// ---------------------------------------------------------------------------------------------
// RNG: fastrand 1.4.1 (WyRand), restated from the published algorithm [source absent; unpinned]
// call sites: math.rs:245 (f32), geom.rs:111 (u8(0..3)), main.rs:86 (seed)
// ---------------------------------------------------------------------------------------------
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed = 0) : s(seed) {}
    uint64_t gen_u64() {
        s += 0xA0761D6478BD642FULL;
        __uint128_t t = (__uint128_t)s * (__uint128_t)(s ^ 0xE7037ED1A0B428DBULL);
        return (uint64_t)t ^ (uint64_t)(t >> 64);
    }
    uint32_t gen_u32() { return (uint32_t)gen_u64(); }
    F f32() {  // 23 mantissa bits in [1,2) minus 1
        uint32_t bits = 0x3F800000u + (gen_u32() >> 9);
        F f;
        std::memcpy(&f, &bits, 4);
        return f - 1.0f;
    }
    uint32_t gen_mod_u32(uint32_t n) {  // Lemire multiply-shift with rejection
        uint32_t r = gen_u32();
        uint64_t m = (uint64_t)r * n;
        uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
        if (lo < n) {
            uint32_t t = (0u - n) % n;
            while (lo < t) {
                r = gen_u32();
                m = (uint64_t)r * n;
                hi = (uint32_t)(m >> 32);
                lo = (uint32_t)m;
            }
        }
        return hi;
    }
};
inline uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
thread_local Rng* tl_rng = nullptr;          // fastrand's thread-local generator
thread_local orc_counters* tl_cnt = nullptr; // visit counters (addition)
inline F frand() { return tl_rng->f32(); }   // Num::rand math.rs:244-246

// ---------------------------------------------------------------------------------------------
// math/generic.rs — V2/V3/V4/M4, component-wise ops, left-to-right dot
// ---------------------------------------------------------------------------------------------
struct V2 { F x, y; };
struct V3 { F x, y, z; };
struct V4 { F x, y, z, w; };
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator*(V2 a, F s) { return {a.x * s, a.y * s}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator*(V3 a, F s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, F s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator+(V3 a, F s) { return {a.x + s, a.y + s, a.z + s}; }
inline V3 operator/(F s, V3 a) { return {s / a.x, s / a.y, s / a.z}; }  // generic.rs:232-241
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator-(V4 a, V4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline V4 operator*(V4 a, F s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V3 v3fill(F v) { return {v, v, v}; }
inline V4 v4fill(F v) { return {v, v, v, v}; }
inline F dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // generic.rs:8-10
inline V3 cross(V3 a, V3 b) {                                           // generic.rs:12-18
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline V3 vmin(V3 a, V3 b) { return {fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)}; }
inline V4 vmin(V4 a, V4 b) { return {fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z), fmin_(a.w, b.w)}; }
inline V4 vmax(V4 a, V4 b) { return {fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z), fmax_(a.w, b.w)}; }
inline V3 vabs(V3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline F dot4(V4 a, V4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }  // generic.rs:45-49
inline V3 contract(V4 a) { return {a.x, a.y, a.z}; }
inline V4 expand(V3 a, F w) { return {a.x, a.y, a.z, w}; }

// math.rs:67-124
inline F length_squared(V3 a) { return dot(a, a); }
inline F length(V3 a) { return std::sqrt(length_squared(a)); }
inline V3 unit(V3 a) { return a / length(a); }
inline V3 random_in_unit_sphere() {  // math.rs:80-92
    for (;;) {
        F x = frand() * 2.0f - 1.0f;
        F y = frand() * 2.0f - 1.0f;
        F z = frand() * 2.0f - 1.0f;
        V3 v{x, y, z};
        if (length_squared(v) >= 1.0f) continue;
        return v;
    }
}
inline V3 random_in_unit_disk() {  // math.rs:94-105
    for (;;) {
        F x = frand() * 2.0f - 1.0f;
        F y = frand() * 2.0f - 1.0f;
        V3 v{x, y, 0.0f};
        if (length_squared(v) >= 1.0f) continue;
        return v;
    }
}
inline V3 random_unit_vector() { return unit(random_in_unit_sphere()); }  // math.rs:107-109
inline bool near_zero(V3 a) {                                             // math.rs:111-113
    return std::fabs(a.x) <= 0.00001f && std::fabs(a.y) <= 0.00001f && std::fabs(a.z) <= 0.00001f;
}
inline V3 reflect(V3 v, V3 n) { return v - (n * dot(v, n) * 2.0f); }      // math.rs:115-117
inline V3 refract(V3 v, V3 n, F etai_over_etat) {                         // math.rs:119-124
    F cos_theta = fmin_(dot(-v, n), 1.0f);
    V3 r_out_perp = (v + n * cos_theta) * etai_over_etat;
    V3 r_out_parallel = n * (-std::sqrt(std::fabs(1.0f - length_squared(r_out_perp))));
    return r_out_perp + r_out_parallel;
}

struct M4 {  // column-major generic.rs:71-77
    V4 c0, c1, c2, c3;
};
inline M4 transpose(const M4& m) {  // generic.rs:84-91
    return {{m.c0.x, m.c1.x, m.c2.x, m.c3.x}, {m.c0.y, m.c1.y, m.c2.y, m.c3.y},
            {m.c0.z, m.c1.z, m.c2.z, m.c3.z}, {m.c0.w, m.c1.w, m.c2.w, m.c3.w}};
}
inline V3 transform(const M4& m, V3 r, F w) {  // generic.rs:105-115
    V4 vx = m.c0 * r.x, vy = m.c1 * r.y, vz = m.c2 * r.z, vw = m.c3 * w;
    V4 v = vx + vy + vz + vw;
    return {v.x, v.y, v.z};
}
inline V3 transform_vector(const M4& m, V3 r) { return transform(m, r, 0.0f); }
inline V3 transform_point(const M4& m, V3 r) { return transform(m, r, 1.0f); }
inline M4 mul(const M4& a, const M4& b) {  // generic.rs:126-161
    M4 m = transpose(a);
    return {{dot4(m.c0, b.c0), dot4(m.c1, b.c0), dot4(m.c2, b.c0), dot4(m.c3, b.c0)},
            {dot4(m.c0, b.c1), dot4(m.c1, b.c1), dot4(m.c2, b.c1), dot4(m.c3, b.c1)},
            {dot4(m.c0, b.c2), dot4(m.c1, b.c2), dot4(m.c2, b.c2), dot4(m.c3, b.c2)},
            {dot4(m.c0, b.c3), dot4(m.c1, b.c3), dot4(m.c2, b.c3), dot4(m.c3, b.c3)}};
}
// math.rs:165-224 — angles are in TURNS; rotate_y / rotate_z sign placement copied literally
inline M4 m4_translation(V3 t) { return {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {t.x, t.y, t.z, 1}}; }
inline M4 m4_rotate_x(F a) {
    F s = std::sin(a * PI * 2.0f), c = std::cos(a * PI * 2.0f);
    return {{1, 0, 0, 0}, {0, c, s, 0}, {0, -s, c, 0}, {0, 0, 0, 1}};
}
inline M4 m4_rotate_y(F a) {
    F s = std::sin(a * PI * 2.0f), c = std::cos(a * PI * 2.0f);
    return {{c, 0, s, 0}, {0, 1, 0, 0}, {-s, 0, c, 0}, {0, 0, 0, 1}};
}
inline M4 m4_rotate_z(F a) {
    F s = std::sin(a * PI * 2.0f), c = std::cos(a * PI * 2.0f);
    return {{c, -s, 0, 0}, {s, c, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
}
inline M4 m4_scale(V3 s) { return {{s.x, 0, 0, 0}, {0, s.y, 0, 0}, {0, 0, s.z, 0}, {0, 0, 0, 1}}; }

// ---------------------------------------------------------------------------------------------
// world.rs:168-182 Ray
// ---------------------------------------------------------------------------------------------
struct Ray {
    V3 origin, direction;
    V3 at(F t) const { return origin + (direction * t); }
};

// ---------------------------------------------------------------------------------------------
// texture.rs — Surface trait and implementations
// ---------------------------------------------------------------------------------------------
struct Surface {
    virtual ~Surface() {}
    virtual uint32_t width() const = 0;
    virtual uint32_t height() const = 0;
    virtual V4 get_f(V2 index) const = 0;
};
enum WrapMode { WRAP_MIRROR = 0, WRAP_REPEAT = 1, WRAP_CLAMP = 2 };
inline F fract(F x) { return x - std::trunc(x); }
inline V2 wrap(int mode, V2 orig) {  // texture.rs:277-300
    if (mode == WRAP_REPEAT) {
        F x = orig.x, y = orig.y;
        x = (x < 0.0f) ? 1.0f - fract(std::fabs(x)) : x;
        y = (y < 0.0f) ? 1.0f - fract(std::fabs(y)) : y;
        x = (x > 1.0f) ? fract(x) : x;
        y = (y > 1.0f) ? fract(y) : y;
        return {x, y};
    } else if (mode == WRAP_CLAMP) {
        return {fmax_(fmin_(orig.x, 1.0f), 0.0f), fmax_(fmin_(orig.y, 1.0f), 0.0f)};
    }
    std::fprintf(stderr, "oracle: Mirror wrapping is not implemented (texture.rs:280)\n");
    std::abort();
}
inline size_t f_as_usize(F v) {  // Rust `as usize`: saturating, NaN -> 0
    if (!(v > 0.0f)) return 0;
    if (v >= 1.8446744e19f) return (size_t)-1;
    return (size_t)v;
}
struct Texture : Surface {  // texture.rs:21-149
    uint32_t w, h;
    std::vector<V4> pixels;
    int wrapping;
    Texture(const uint8_t* rgba, uint32_t w_, uint32_t h_, int wrap_) : w(w_), h(h_), wrapping(wrap_) {  // load_bytes :70-103
        pixels.resize((size_t)w * h);
        for (size_t i = 0; i < pixels.size(); ++i)
            pixels[i] = {(F)rgba[4 * i] / 255.0f, (F)rgba[4 * i + 1] / 255.0f, (F)rgba[4 * i + 2] / 255.0f, (F)rgba[4 * i + 3] / 255.0f};
    }
    uint32_t width() const override { return w; }
    uint32_t height() const override { return h; }
    const V4& at(size_t x, size_t y) const { return pixels.at(y * (size_t)w + x); }  // Index :109-115 (panics OOB)
    V4 get_f(V2 index) const override {  // :126-148
        index = wrap(wrapping, index);
        F x = index.x * (F)(w - 1);
        F y = index.y * (F)(h - 1);
        size_t x0 = f_as_usize(std::floor(x)), x1 = f_as_usize(std::ceil(x));
        size_t y0 = f_as_usize(std::floor(y)), y1 = f_as_usize(std::ceil(y));
        F t = x - (F)x0;
        V4 p0 = at(x0, y0) * (1.0f - t) + at(x1, y0) * t;
        V4 p1 = at(x0, y1) * (1.0f - t) + at(x1, y1) * t;
        t = y - (F)y0;
        return p1 * t + p0 * (1.0f - t);
    }
};
struct SolidColor : Surface {  // texture.rs:179-194
    V4 c;
    explicit SolidColor(V4 c_) : c(c_) {}
    uint32_t width() const override { return 1; }
    uint32_t height() const override { return 1; }
    V4 get_f(V2) const override { return c; }
};
const F KR = 0.2126f, KG = 0.7152f, KB = 0.0722f;  // texture.rs:196-198
struct YCbCrTexture : Surface {                    // texture.rs:207-248
    const Texture *luma, *chroma;
    YCbCrTexture(const Texture* l, const Texture* c) : luma(l), chroma(c) {}
    uint32_t width() const override { return luma->width(); }
    uint32_t height() const override { return luma->height(); }
    V4 get_f(V2 index) const override {
        static const M4 YUV = {{1.0f, 1.0f, 1.0f, 0.0f},
                               {0.0f, -(KB / KG) * (2.0f - 2.0f * KB), 2.0f - 2.0f * KB, 0.0f},
                               {2.0f - 2.0f * KR, -(KR / KG) * (2.0f - 2.0f * KR), 0.0f, 0.0f},
                               {0.0f, 0.0f, 0.0f, 1.0f}};
        V4 l = luma->get_f(index), c = chroma->get_f(index);
        V3 yuv{l.x, c.x - 0.5f, c.y - 0.5f};
        V3 col = vmax(vmin(transform_point(YUV, yuv), v3fill(1.0f)), v3fill(0.0f));
        col = {std::pow(col.x, 2.2f), std::pow(col.y, 2.2f), std::pow(col.z, 2.2f)};
        return expand(col, 1.0f);
    }
};
struct TextureBlend : Surface {  // texture.rs:250-334
    int mode;
    const Surface *left, *right;
    TextureBlend(int m, const Surface* l, const Surface* r) : mode(m), left(l), right(r) {}
    uint32_t width() const override { return std::max(left->width(), right->width()); }
    uint32_t height() const override { return std::max(left->height(), right->height()); }
    V4 get_f(V2 index) const override {
        V4 l = left->get_f(index), r = right->get_f(index);
        switch (mode) {
            case 0: return vmax(l, r);
            case 1: return vmin(l, r);
            case 2: return vmin(l + r, v4fill(1.0f));
            default: return vmax(l - r, v4fill(0.0f));
        }
    }
};
struct SolidColorFallback : Surface {  // texture.rs:336-360 (height() returns width: reference quirk :352)
    V4 color;
    const Surface* surface;
    SolidColorFallback(V4 c, const Surface* s) : color(c), surface(s) {}
    uint32_t width() const override { return surface->width(); }
    uint32_t height() const override { return surface->width(); }
    V4 get_f(V2 index) const override {
        V4 c = surface->get_f(index);
        return (color * (1.0f - c.w)) + (c * c.w);
    }
};

// ---------------------------------------------------------------------------------------------
// geom.rs:7-33 Hit ; material.rs:10-27 Material / Scatter
// ---------------------------------------------------------------------------------------------
struct Material;
struct Hit {
    V3 point, normal;
    bool has_uv = false;
    V2 uv{0, 0};
    F t = 0;
    bool front_face = false;
    const Material* material = nullptr;
    uint32_t object = NONE_ID, tri = NONE_ID;  // additions: ids for the parity harness
    void set_face_normal(const Ray& ray, V3 outward_normal) {  // geom.rs:17-24
        front_face = dot(ray.direction, outward_normal) < 0.0f;
        normal = front_face ? outward_normal : -outward_normal;
    }
};
struct Scatter {
    V3 attenuation;
    Ray scattered;
};
struct Material {
    virtual ~Material() {}
    virtual bool scatter(const Ray& ray, const Hit& hit, Scatter& out) const = 0;
    virtual bool emit(const Hit&, V3&) const { return false; }
    virtual bool normal(V2, V3&) const { return false; }
    virtual bool alpha_test(V2) const { return true; }
};
inline V2 uv_or_zero(const Hit& h) { return h.has_uv ? h.uv : V2{0, 0}; }

struct Lambertian : Material {  // material.rs:192-225
    const Surface* surface;
    explicit Lambertian(const Surface* s) : surface(s) {}
    bool scatter(const Ray&, const Hit& hit, Scatter& out) const override {
        V3 dir = hit.normal + random_unit_vector();
        if (near_zero(dir)) dir = hit.normal;
        out.scattered = {hit.point, dir};
        out.attenuation = contract(surface->get_f(uv_or_zero(hit)));
        return true;
    }
    bool alpha_test(V2 uv) const override { return surface->get_f(uv).w != 0.0f; }
};
struct DiffuseLight : Material {  // material.rs:227-246
    V3 e;
    explicit DiffuseLight(V3 e_) : e(e_) {}
    bool scatter(const Ray&, const Hit&, Scatter&) const override { return false; }
    bool emit(const Hit&, V3& out) const override { out = e; return true; }
};
struct Metal : Material {  // material.rs:248-284
    F fuzz;
    const Surface* surface;
    Metal(F f, const Surface* s) : fuzz(f < 1.0f ? f : 1.0f), surface(s) {}
    bool scatter(const Ray& ray, const Hit& hit, Scatter& out) const override {
        V3 reflected = reflect(unit(ray.direction), hit.normal);
        Ray scattered{hit.point, reflected + (random_in_unit_sphere() * fuzz)};
        if (dot(scattered.direction, hit.normal) > 0.0f) {
            out.attenuation = contract(surface->get_f(uv_or_zero(hit)));
            out.scattered = scattered;
            return true;
        }
        return false;
    }
    bool alpha_test(V2 uv) const override { return surface->get_f(uv).w != 0.0f; }
};
inline F powi2(F x) { return x * x; }
inline F powi5(F x) { F x2 = x * x; return x2 * x2 * x; }  // llvm.powi: square-and-multiply
inline F reflectance(F cosine, F ref_idx) {  // material.rs:296-299, :346-349
    F r0 = powi2((1.0f - ref_idx) / (1.0f + ref_idx));
    return r0 + (1.0f - r0) * powi5(1.0f - cosine);
}
struct Dielectric : Material {  // material.rs:286-329
    F ior;
    explicit Dielectric(F i) : ior(i) {}
    bool scatter(const Ray& ray, const Hit& hit, Scatter& out) const override {
        F ratio = hit.front_face ? 1.0f / ior : ior;
        V3 ud = unit(ray.direction);
        F cos_theta = fmin_(dot(-ud, hit.normal), 1.0f);
        F sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
        bool cannot_refract = ratio * sin_theta > 1.0f;
        V3 dir = (cannot_refract || reflectance(cos_theta, ratio) > frand()) ? reflect(ud, hit.normal) : refract(ud, hit.normal, ratio);
        out.attenuation = v3fill(1.0f);
        out.scattered = {hit.point, dir};
        return true;
    }
};
struct Specular : Material {  // material.rs:331-383
    F ior;
    Lambertian inner;
    Specular(F i, const Surface* s) : ior(i), inner(s) {}
    bool scatter(const Ray& ray, const Hit& hit, Scatter& out) const override {
        F ratio = hit.front_face ? 1.0f / ior : ior;
        V3 ud = unit(ray.direction);
        F cos_theta = fmin_(dot(-ud, hit.normal), 1.0f);
        F sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
        bool cannot_refract = ratio * sin_theta > 1.0f;
        if (cannot_refract || reflectance(cos_theta, ratio) > frand()) {
            out.attenuation = v3fill(1.0f);
            out.scattered = {hit.point, reflect(ud, hit.normal)};
            return true;
        }
        return inner.scatter(ray, hit, out);
    }
    bool alpha_test(V2 uv) const override { return inner.alpha_test(uv); }
};
struct Absorb : Material {  // impl Material for () material.rs:385-389
    bool scatter(const Ray&, const Hit&, Scatter&) const override { return false; }
};
struct Mix : Material {  // material.rs:391-426 — independent coin flips in scatter, emit and alpha_test
    F ratio;
    const Material *left, *right;
    Mix(F r, const Material* l, const Material* rr) : ratio(r), left(l), right(rr) {}
    bool scatter(const Ray& ray, const Hit& hit, Scatter& out) const override {
        return (frand() < ratio) ? left->scatter(ray, hit, out) : right->scatter(ray, hit, out);
    }
    bool emit(const Hit& hit, V3& out) const override { return (frand() < ratio) ? left->emit(hit, out) : right->emit(hit, out); }
    bool alpha_test(V2 uv) const override { return (frand() < ratio) ? left->alpha_test(uv) : right->alpha_test(uv); }
};
// eve.rs:23-199 — the one material of the reference that implements Material::normal (the tangent-space hook of Triangle::intersect,
// geom.rs:551-560): three textures (normal + occlusion, albedo + roughness, paint / material / dirt / glow masks) and a palette.
struct EveMaterial : Material {
    const Surface *normal_occlusion, *albedo_roughness, *pmdg;
    V3 colors[4], glow;
    EveMaterial(const Surface* no, const Surface* ar, const Surface* p, const V3 c[4], V3 g) : normal_occlusion(no), albedo_roughness(ar), pmdg(p), glow(g) {
        for (int i = 0; i < 4; ++i) colors[i] = c[i];
    }
    V3 palette(F i) const {  // EveMaterialColor::get :190-199
        i = i * 3.0f;
        F f0 = std::floor(i), f1 = std::ceil(i);
        size_t i0 = f0 > 0.0f ? (size_t)f0 : 0, i1 = f1 > 0.0f ? (size_t)f1 : 0;  // `as usize` saturates at 0; the reference panics past the palette
        if (i0 > 3) i0 = 3;
        if (i1 > 3) i1 = 3;
        F t = i - (F)i0;
        return colors[i0] * (1.0f - t) + colors[i1] * t;
    }
    bool scatter(const Ray& ray, const Hit& hit, Scatter& out) const override {  // :91-119
        if (!hit.has_uv) return false;
        V4 ar = albedo_roughness->get_f(hit.uv);
        V3 albedo = contract(ar);
        F roughness = ar.w;
        V4 m = pmdg->get_f(hit.uv);
        F paint = m.x, material = m.y, dirt = m.z * 1.0f;
        V3 material_color = palette(material);
        V3 color = (((albedo * material_color * (1.0f - paint)) + (albedo * paint)) * (1.0f - fmin_(dirt, 1.0f))) + (V3{0.01f, 0.005f, 0.0f} * dirt);
        SolidColor solid(V4{color.x, color.y, color.z, 1.0f});
        Lambertian lambertian(&solid);
        Specular specular(1.8f, &solid);
        Mix mix(fmin_(roughness + dirt, 1.0f), &lambertian, &specular);
        return mix.scatter(ray, hit, out);
    }
    bool emit(const Hit& hit, V3& out) const override {  // :121-128
        if (!hit.has_uv) return false;
        out = glow * pmdg->get_f(hit.uv).w * 10.0f;
        return true;
    }
    bool normal(V2 uv, V3& out) const override {  // :130-133 with normal_occlusion :66-73
        V4 pixel = normal_occlusion->get_f(uv);
        pixel = pixel * 2.0f - V4{1.0f, 1.0f, 1.0f, 1.0f};
        F x = 1.0f - pixel.y * pixel.y - pixel.w * pixel.w;  // powi(2)
        F z = std::sqrt(std::fabs(x));
        out = unit(V3{pixel.y, pixel.w, z});
        return true;
    }
};
struct Isotrophic : Material {  // material.rs:428-445
    V3 albedo;
    explicit Isotrophic(V3 a) : albedo(a) {}
    bool scatter(const Ray&, const Hit& hit, Scatter& out) const override {
        out.attenuation = albedo;
        out.scattered = {hit.point, random_in_unit_sphere()};
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// material.rs:29-190 Backgrounds
// ---------------------------------------------------------------------------------------------
struct Background {
    virtual ~Background() {}
    virtual V3 background(const Ray& ray) const = 0;
};
struct SolidBackground : Background {
    V3 c;
    explicit SolidBackground(V3 c_) : c(c_) {}
    V3 background(const Ray&) const override { return c; }
};
struct SkyBackground : Background {  // :55-63
    V3 background(const Ray& ray) const override {
        V3 ud = unit(ray.direction);
        F t = 0.5f * (ud.y + 1.0f);
        return (v3fill(1.0f) * (1.0f - t)) + (V3{0.5f, 0.7f, 1.0f} * t);
    }
};
struct SkySphere : Background {  // :65-89
    const Surface* tex;
    explicit SkySphere(const Surface* s) : tex(s) {}
    V3 background(const Ray& ray) const override {
        V3 p = unit(ray.direction);
        F theta = std::acos(p.y);
        F phi = std::atan2(p.z * -1.0f, p.x) + PI;
        V2 uv{phi / (2.0f * PI), theta / PI};
        return contract(tex->get_f(uv));
    }
};
struct CubeMap : Background {  // :91-190 (rotate_x used for all three angles :103-105; y index swap :153-161 — quirks kept)
    const Surface* f[6];
    M4 tf;
    CubeMap(const Surface* const s[6], V3 rot) {
        for (int i = 0; i < 6; ++i) f[i] = s[i];
        tf = mul(mul(m4_rotate_x(rot.x), m4_rotate_x(rot.y)), m4_rotate_x(rot.z));
    }
    V3 background(const Ray& ray) const override {
        V3 p = transform_vector(tf, ray.direction);
        V3 a = vabs(p);
        bool xl = a.x >= a.y && a.x >= a.z, yl = a.y >= a.x && a.y >= a.z, zl = a.z >= a.x && a.z >= a.y;
        int index = 0;
        F max_axis = 0.0f, u = 0.0f, v = 0.0f;
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
        V2 uv{0.5f * (u / max_axis + 1.0f), 0.5f * (v / max_axis + 1.0f)};
        return contract(f[index]->get_f(uv));
    }
};

// ---------------------------------------------------------------------------------------------
// geom.rs:202-272 BoundingBox
// ---------------------------------------------------------------------------------------------
struct BoundingBox {
    V3 minimum, maximum;
    bool hit(const Ray& ray, F t_min, F t_max) const {  // :218-247
        if (tl_cnt) tl_cnt->box_tests++;
        V3 v_min = (minimum - ray.origin) / ray.direction;
        V3 v_max = (maximum - ray.origin) / ray.direction;
        V3 mn = vmin(v_min, v_max), mx = vmax(v_min, v_max);
        t_min = fmax_(mn.x, t_min);
        t_max = fmin_(mx.x, t_max);
        if (t_max < t_min) return false;
        t_min = fmax_(mn.y, t_min);
        t_max = fmin_(mx.y, t_max);
        if (t_max < t_min) return false;
        t_min = fmax_(mn.z, t_min);
        t_max = fmin_(mx.z, t_max);
        if (t_max < t_min) return false;
        return true;
    }
    BoundingBox join(const BoundingBox& o) const { return {vmin(minimum, o.minimum), vmax(maximum, o.maximum)}; }  // :249-254
    V3 corner(int c) const {  // :256-272
        return {(c & 1) == 0 ? maximum.x : minimum.x, (c & 2) == 0 ? maximum.y : minimum.y, (c & 4) == 0 ? maximum.z : minimum.z};
    }
};

struct Intersect {  // geom.rs:35-38
    virtual ~Intersect() {}
    virtual bool intersect(const Ray& ray, F t_min, F t_max, Hit& hit) const = 0;
    virtual bool bounding_box(BoundingBox& out) const = 0;
};

// geom.rs:40-101
struct Sphere : Intersect {
    V3 center;
    F radius;
    const Material* material;
    uint32_t object = NONE_ID;
    Sphere(const Material* m, V3 c, F r) : center(c), radius(r), material(m) {}
    bool intersect(const Ray& ray, F t_min, F t_max, Hit& hit) const override {
        if (tl_cnt) tl_cnt->sphere_tests++;
        V3 oc = ray.origin - center;
        F a = length_squared(ray.direction);
        F half_b = dot(oc, ray.direction);
        F c = length_squared(oc) - (radius * radius);
        F disc = (half_b * half_b) - (a * c);
        if (disc < 0.0f) return false;
        F sqrt_d = std::sqrt(disc);
        F root = (-half_b - sqrt_d) / a;
        if (root < t_min || t_max < root) {
            root = (-half_b + sqrt_d) / a;
            if (root < t_min || t_max < root) return false;
        }
        V3 point = ray.at(root);
        V3 normal = (point - center) / radius;
        hit = Hit();
        hit.point = point;
        hit.normal = normal;
        hit.t = root;
        hit.has_uv = false;
        hit.material = material;
        hit.object = object;
        hit.set_face_normal(ray, normal);
        return true;
    }
    bool bounding_box(BoundingBox& out) const override {
        out = {center - v3fill(std::fabs(radius)), center + v3fill(std::fabs(radius))};
        return true;
    }
};

// geom.rs:103-200
struct BvhNode : Intersect {
    const Intersect* left = nullptr;
    const Intersect* right = nullptr;
    std::unique_ptr<BvhNode> own_left, own_right;  // set when the child is a BvhNode
    BoundingBox bbox;
    uint64_t node_count = 1;

    static bool compare(int axis, const Intersect* a, const Intersect* b) {  // compare_{x,y,z} :164-183
        BoundingBox ba, bb;
        if (!a->bounding_box(ba) || !b->bounding_box(bb)) { std::fprintf(stderr, "Missing bounding box in bvh\n"); std::abort(); }
        if (axis == 0) return ba.minimum.x < bb.minimum.x;
        if (axis == 1) return ba.minimum.y < bb.minimum.y;
        return ba.minimum.z < bb.minimum.z;
    }
    // BvhNode::new :109-161.  rng = the constructing thread's fastrand.
    // Two harness additions that leave the tree of a list of < kForkItems items exactly as before: the sort runs on keys fetched
    // once per item (the same strict predicate on the same values, so the same stable order), and the two halves of a list of
    // >= kForkItems items are built as parallel tasks, the right one on a generator forked from the parent's (the reference builds
    // on one thread with one stream; which axis a node draws cannot change a closest hit). The 10 x 1 M-triangle scene of
    // BASELINE configs[4] builds in a minute instead of five.
    static constexpr size_t kForkItems = 1u << 16;
    BvhNode(std::vector<const Intersect*> items, Rng& rng) {
        int axis = (int)rng.gen_mod_u32(3);  // fastrand::u8(0..3)
        if (items.size() == 1) {
            left = items.back();
        } else if (items.size() == 2) {
            const Intersect* a = items.back(); items.pop_back();
            const Intersect* b = items.back(); items.pop_back();
            if (compare(axis, a, b)) { left = a; right = b; } else { left = b; right = a; }
        } else {
            // Rust sort_by with Less/Greater only; std::stable_sort on the same strict predicate (tie order unpinned)
            {
                std::vector<std::pair<F, const Intersect*>> keyed(items.size());
                for (size_t i = 0; i < items.size(); ++i) {
                    BoundingBox b;
                    if (!items[i]->bounding_box(b)) { std::fprintf(stderr, "Missing bounding box in bvh\n"); std::abort(); }
                    keyed[i] = {axis == 0 ? b.minimum.x : (axis == 1 ? b.minimum.y : b.minimum.z), items[i]};
                }
                std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<F, const Intersect*>& a, const std::pair<F, const Intersect*>& b) { return a.first < b.first; });
                for (size_t i = 0; i < items.size(); ++i) items[i] = keyed[i].second;
            }
            size_t mid = items.size() / 2;
            std::vector<const Intersect*> back_half(items.begin() + mid, items.end());
            items.resize(mid);
            if (items.size() + back_half.size() >= kForkItems) {
                Rng right_rng(splitmix(rng.gen_u64()));
                std::thread right_task([&] { own_right.reset(new BvhNode(std::move(back_half), right_rng)); });
                own_left.reset(new BvhNode(std::move(items), rng));
                right_task.join();
            } else {
                own_left.reset(new BvhNode(std::move(items), rng));
                own_right.reset(new BvhNode(std::move(back_half), rng));
            }
            left = own_left.get();
            right = own_right.get();
            node_count += own_left->node_count + own_right->node_count;
        }
        BoundingBox bl, br;
        bool hl = left && left->bounding_box(bl), hr = right && right->bounding_box(br);
        if (hl && hr) bbox = bl.join(br);
        else if (hl) bbox = bl;
        else if (hr) bbox = br;
        else { std::fprintf(stderr, "Missing bounding box in bvh\n"); std::abort(); }
    }
    bool intersect(const Ray& ray, F t_min, F t_max, Hit& hit) const override {  // :186-200
        if (!bbox.hit(ray, t_min, t_max)) return false;
        Hit lh;
        bool has_left = left && left->intersect(ray, t_min, t_max, lh);
        F tm = has_left ? lh.t : t_max;
        Hit rh;
        if (right && right->intersect(ray, t_min, tm, rh)) { hit = rh; return true; }
        if (has_left) { hit = lh; return true; }
        return false;
    }
    bool bounding_box(BoundingBox& out) const override { out = bbox; return true; }
};

// geom.rs:422-585
struct Triangle : Intersect {
    V3 va, vb, vc;
    bool has_uv = false;
    V2 uv_a{0, 0}, uv_b{0, 0}, uv_c{0, 0};
    const Material* material;
    V3 na, nb, nc, tangent{0, 0, 0}, bitangent{0, 0, 0};
    uint32_t tri = NONE_ID;
    Triangle(const Material* m, V3 a, V3 b, V3 c) : va(a), vb(b), vc(c), material(m) {  // Triangle::new :449-466
        V3 ab = b - a, ac = c - a;
        V3 n = unit(cross(ab, ac));
        na = nb = nc = n;
    }
    Triangle(const Material* m, V3 a, V3 n_a, V2 t_a, V3 b, V3 n_b, V2 t_b, V3 c, V3 n_c, V2 t_c)  // with_norms_and_uvs :468-496
        : va(a), vb(b), vc(c), has_uv(true), uv_a(t_a), uv_b(t_b), uv_c(t_c), material(m), na(n_a), nb(n_b), nc(n_c) {
        V3 ab = b - a, ac = c - a;
        V2 uv_ab = t_b - t_a, uv_ac = t_c - t_a;
        F r = fmax_(fmin_(1.0f / (uv_ab.x * uv_ac.y - uv_ab.y * uv_ac.x), 1.0f), -1.0f);
        tangent = (ab * uv_ac.y - ac * uv_ab.y) * r;
        bitangent = (ac * uv_ab.x - ab * uv_ac.x) * r;
    }
    bool intersect(const Ray& ray, F t_min, F t_max, Hit& hit) const override {  // :504-577
        if (tl_cnt) tl_cnt->tri_tests++;
        V3 ab = vb - va, ac = vc - va;
        V3 p_vec = cross(ray.direction, ac);
        F det = dot(ab, p_vec);
        if (std::fabs(det) < 0.000001f) return false;
        F inv_det = 1.0f / det;
        V3 t_vec = ray.origin - va;
        F u = dot(t_vec, p_vec) * inv_det;
        if (u < 0.0f || u > 1.0f) return false;
        V3 q_vec = cross(t_vec, ab);
        F v = dot(ray.direction, q_vec) * inv_det;
        if (v < 0.0f || v + u > 1.0f) return false;
        F t = dot(ac, q_vec) * inv_det;
        if (t < t_min || t > t_max) return false;
        V3 point = ray.at(t);
        V3 d0 = va - point, d1 = vb - point, d2 = vc - point;
        F area = length(cross(va - vb, va - vc));
        F a0 = length(cross(d1, d2)) / area;
        F a1 = length(cross(d2, d0)) / area;
        F a2 = length(cross(d0, d1)) / area;
        V3 normal = na * a0 + nb * a1 + nc * a2;
        V2 uv{0, 0};
        if (has_uv) {
            uv = uv_a * a0 + uv_b * a1 + uv_c * a2;
            V3 tn;
            if (material->normal(uv, tn)) normal = tangent * tn.x + bitangent * tn.y + normal * tn.z;
            if (!material->alpha_test(uv)) return false;
        }
        hit = Hit();
        hit.point = point;
        hit.normal = normal;
        hit.t = t;
        hit.has_uv = has_uv;
        hit.uv = uv;
        hit.material = material;
        hit.tri = tri;
        hit.set_face_normal(ray, normal);
        return true;
    }
    bool bounding_box(BoundingBox& out) const override {  // :579-584
        out = {vmin(vmin(va, vb), vc), vmax(vmax(va, vb), vc)};
        return true;
    }
};

struct Mesh {  // the Arc<BvhNode> a Model owns (geom.rs:275-292) plus its triangle storage
    std::vector<Triangle> tris;
    std::unique_ptr<BvhNode> bvh;
};

// geom.rs:275-333
struct Model : Intersect {
    const Material* material;  // Option<M>
    const BvhNode* triangles;
    uint32_t object = NONE_ID;
    Model(const BvhNode* b, const Material* m) : material(m), triangles(b) {}
    bool intersect(const Ray& ray, F t_min, F t_max, Hit& hit) const override {
        if (!triangles->intersect(ray, t_min, t_max, hit)) return false;
        if (material) hit.material = material;
        hit.object = object;
        return true;
    }
    bool bounding_box(BoundingBox& out) const override { return triangles->bounding_box(out); }
};

// geom.rs:335-420
struct Instance : Intersect {
    const BvhNode* triangles;
    const Material* material;  // Option<M>
    M4 transform_, inv_transform;
    BoundingBox bbox;
    uint32_t object = NONE_ID;
    Instance(const BvhNode* b, V3 translation, V3 rotation, V3 scale, const Material* m) : triangles(b), material(m) {  // :343-390
        V3 inv_translation = translation * -1.0f;
        V3 inv_rotation = rotation * -1.0f;
        V3 inv_scale = 1.0f / scale;
        M4 mt = m4_translation(translation), mit = m4_translation(inv_translation);
        M4 rx = m4_rotate_x(rotation.x), ry = m4_rotate_y(rotation.y), rz = m4_rotate_z(rotation.z);
        M4 irx = m4_rotate_x(inv_rotation.x), iry = m4_rotate_y(inv_rotation.y), irz = m4_rotate_z(inv_rotation.z);
        M4 rot = mul(mul(rx, ry), rz);
        M4 irot = mul(mul(irz, iry), irx);
        M4 ms = m4_scale(scale), mis = m4_scale(inv_scale);
        transform_ = mul(mul(mt, rot), ms);
        inv_transform = mul(mul(mis, irot), mit);
        V3 mn = v3fill(INF), mx = v3fill(-INF);
        for (int c = 0; c < 8; ++c) {
            V3 corner = transform_point(transform_, b->bbox.corner(c));
            mn = vmin(mn, corner);
            mx = vmax(mx, corner);
        }
        bbox = {mn, mx};
    }
    bool intersect(const Ray& ray_in, F t_min, F t_max, Hit& hit) const override {  // :404-420
        if (tl_cnt) tl_cnt->instance_tests++;
        Ray ray{transform_point(inv_transform, ray_in.origin), transform_vector(inv_transform, ray_in.direction)};
        if (!triangles->intersect(ray, t_min, t_max, hit)) return false;
        hit.point = transform_point(transform_, hit.point);
        hit.normal = unit(transform_vector(transform_, hit.normal));
        if (material) hit.material = material;
        hit.object = object;
        return true;
    }
    bool bounding_box(BoundingBox& out) const override { out = bbox; return true; }
};

// geom.rs:587-660
struct Volume : Intersect {
    F neg_inv_density;
    std::unique_ptr<Intersect> target;
    Isotrophic material;
    uint32_t object = NONE_ID;
    Volume(Intersect* tgt, F density, V3 albedo) : neg_inv_density(-1.0f / density), target(tgt), material(albedo) {}
    bool intersect(const Ray& ray, F t_min, F t_max, Hit& hit) const override {
        if (tl_cnt) tl_cnt->volume_tests++;
        Hit enter, exit_;
        if (!target->intersect(ray, -INF, INF, enter)) return false;
        if (!target->intersect(ray, enter.t + 0.0001f, INF, exit_)) return false;
        if (enter.t < t_min) enter.t = t_min;
        if (exit_.t > t_max) exit_.t = t_max;
        if (enter.t >= exit_.t) return false;
        if (enter.t < 0.0f) enter.t = 0.0f;
        F ray_length = length(ray.direction);
        F distance_inside = (exit_.t - enter.t) * ray_length;
        F hit_distance = std::log(frand()) * neg_inv_density;
        if (hit_distance > distance_inside) return false;
        F t = enter.t + hit_distance / ray_length;
        hit = Hit();
        hit.point = ray.at(t);
        hit.normal = {1.0f, 0.0f, 0.0f};
        hit.has_uv = false;
        hit.t = t;
        hit.front_face = true;
        hit.material = &material;
        hit.object = object;
        return true;
    }
    bool bounding_box(BoundingBox& out) const override { return target->bounding_box(out); }
};

// ---------------------------------------------------------------------------------------------
// world.rs
// ---------------------------------------------------------------------------------------------
struct Camera {  // world.rs:5-63
    V3 origin, lower_left_corner, horizontal, vertical, u, v;
    F lens_radius;
    Camera() {}
    Camera(F vfov, V3 look_from, V3 look_at, V3 view_up, F aspect, F aperture, F focus) {
        F rads = vfov * PI / 180.0f;
        F half_height = std::tan(rads / 2.0f);
        F viewport_height = half_height * 2.0f;
        F viewport_width = aspect * viewport_height;
        V3 w = unit(look_from - look_at);
        u = unit(cross(view_up, w));
        v = cross(w, u);
        origin = look_from;
        horizontal = u * viewport_width * focus;
        vertical = v * viewport_height * focus;
        lower_left_corner = origin - (horizontal / 2.0f) - (vertical / 2.0f) - (w * focus);
        lens_radius = aperture / 2.0f;
    }
    Ray ray(F s, F t) const {
        V3 blur = random_in_unit_disk() * lens_radius;
        V3 offset = u * blur.x + v * blur.y;
        return {origin + offset, lower_left_corner + (horizontal * s) + (vertical * t) - origin - offset};
    }
};

struct World {  // world.rs:96-166
    std::unique_ptr<Background> background;
    std::vector<const Intersect*> objects;
    std::unique_ptr<BvhNode> tlas;
    void build_bvh(Rng& rng) {  // :117-122
        std::vector<const Intersect*> items;
        items.swap(objects);
        tlas.reset(new BvhNode(std::move(items), rng));
        objects.push_back(tlas.get());
    }
    bool intersect(const Ray& ray, F t_min, F t_max, Hit& hit) const {  // :131-144
        bool found = false;
        F closest = t_max;
        Hit h;
        for (const Intersect* obj : objects) {
            if (obj->intersect(ray, t_min, closest, h)) {
                closest = h.t;
                hit = h;
                found = true;
            }
        }
        return found;
    }
};

inline V3 hit_emit(const Hit& h) {  // geom.rs:30-32
    V3 e;
    return h.material->emit(h, e) ? e : V3{0, 0, 0};
}

// Camera::trace world.rs:65-79 (recursive, radiance composed on the way back up)
static void trace(const World& scene, const Ray& ray, uint32_t depth, V3& color, uint32_t& depth_out) {
    if (depth == 0) { color = {0, 0, 0}; depth_out = depth; return; }
    if (tl_cnt) tl_cnt->rays++;
    Hit hit;
    if (scene.intersect(ray, 0.001f, INF, hit)) {
        V3 emitted = hit_emit(hit);
        Scatter sc;
        if (hit.material->scatter(ray, hit, sc)) {
            V3 c;
            trace(scene, sc.scattered, depth - 1, c, depth_out);
            color = (c * sc.attenuation) + emitted;
        } else {
            color = emitted;
            depth_out = depth;
        }
    } else {
        color = scene.background->background(ray);
        depth_out = depth;
    }
}
// Camera::albedo_normal world.rs:81-93 (+ ids, t)
static void albedo_normal(const World& scene, const Ray& ray, V3& albedo, V3& normal, uint32_t& object, uint32_t& tri, F& t) {
    if (tl_cnt) tl_cnt->rays++;
    Hit hit;
    if (scene.intersect(ray, 0.001f, INF, hit)) {
        V3 emitted = hit_emit(hit);
        Scatter sc;
        albedo = hit.material->scatter(ray, hit, sc) ? sc.attenuation : emitted;
        normal = hit.normal;
        object = hit.object;
        tri = hit.tri;
        t = hit.t;
    } else {
        albedo = scene.background->background(ray);
        normal = {0, 0, 0};
        object = NONE_ID;
        tri = NONE_ID;
        t = INF;
    }
}

// ---------------------------------------------------------------------------------------------
// ply_loader.rs — ascii / binary_little_endian / binary_big_endian; only vertex.{x,y,z} and 3-index lists
// ---------------------------------------------------------------------------------------------
enum PlyFormat { PLY_ASCII, PLY_LE, PLY_BE };
enum PlyType { T_CHAR, T_UCHAR, T_SHORT, T_USHORT, T_INT, T_UINT, T_FLOAT, T_DOUBLE, T_BAD };
static PlyType ply_type(const std::string& s) {  // ply_loader.rs:172-190
    if (s == "char" || s == "int8") return T_CHAR;
    if (s == "uchar" || s == "uint8") return T_UCHAR;
    if (s == "short" || s == "int16") return T_SHORT;
    if (s == "ushort" || s == "uint16") return T_USHORT;
    if (s == "int" || s == "int32") return T_INT;
    if (s == "uint" || s == "uint32") return T_UINT;
    if (s == "float" || s == "float32") return T_FLOAT;
    if (s == "double" || s == "float64") return T_DOUBLE;
    return T_BAD;
}
static size_t ply_size(PlyType t) {
    switch (t) { case T_CHAR: case T_UCHAR: return 1; case T_SHORT: case T_USHORT: return 2; case T_DOUBLE: return 8; default: return 4; }
}
struct PlyProp { bool is_list; std::string name; PlyType kind, count_kind; };
struct PlyElem { std::string name; size_t count; std::vector<PlyProp> props; };
struct PlyReader {
    std::istream& in;
    PlyFormat fmt;
    bool word(std::string& w) {  // ascii token: skip whitespace, read until whitespace (:24-33)
        w.clear();
        char c;
        while (in.get(c)) {
            bool ws = std::isspace((unsigned char)c);
            if (ws && !w.empty()) return true;
            if (!ws) w.push_back(c);
        }
        return false;  // read_exact hit EOF -> error
    }
    template <class T> bool raw(T& v) {
        unsigned char b[8];
        if (!in.read((char*)b, sizeof(T))) return false;
        if (fmt == PLY_BE) std::reverse(b, b + sizeof(T));
        std::memcpy(&v, b, sizeof(T));
        return true;
    }
    bool read_f64(PlyType k, double& out) {
        switch (k) {
            case T_CHAR: { int8_t v; if (!raw(v)) return false; out = v; return true; }
            case T_UCHAR: { uint8_t v; if (!raw(v)) return false; out = v; return true; }
            case T_SHORT: { int16_t v; if (!raw(v)) return false; out = v; return true; }
            case T_USHORT: { uint16_t v; if (!raw(v)) return false; out = v; return true; }
            case T_INT: { int32_t v; if (!raw(v)) return false; out = v; return true; }
            case T_UINT: { uint32_t v; if (!raw(v)) return false; out = v; return true; }
            case T_FLOAT: { float v; if (!raw(v)) return false; out = v; return true; }
            case T_DOUBLE: { double v; if (!raw(v)) return false; out = v; return true; }
            default: return false;
        }
    }
    bool read_f32(PlyType k, float& out) {  // :69-110
        if (fmt == PLY_ASCII) {
            std::string w;
            if (!word(w)) return false;
            char* end = nullptr;
            out = std::strtof(w.c_str(), &end);
            return end && *end == 0 && end != w.c_str();
        }
        if (k == T_FLOAT) return raw(out);
        double d;
        if (!read_f64(k, d)) return false;
        if (k == T_INT) { int32_t i = (int32_t)d; out = (float)i; return true; }
        if (k == T_UINT) { uint32_t i = (uint32_t)d; out = (float)i; return true; }
        out = (float)d;
        return true;
    }
    bool read_usize(PlyType k, size_t& out) {  // :14-67
        if (fmt == PLY_ASCII) {
            std::string w;
            if (!word(w)) return false;
            if (k == T_FLOAT || k == T_DOUBLE) {
                char* end = nullptr;
                double d = std::strtod(w.c_str(), &end);
                if (!end || *end != 0 || end == w.c_str()) return false;
                out = d > 0 ? (size_t)d : 0;
                return true;
            }
            if (w.empty()) return false;
            size_t i = (w[0] == '+') ? 1 : 0;  // usize::from_str accepts a leading '+'
            if (i >= w.size()) return false;
            size_t v = 0;
            for (; i < w.size(); ++i) {
                if (w[i] < '0' || w[i] > '9') return false;
                v = v * 10 + (size_t)(w[i] - '0');
            }
            out = v;
            return true;
        }
        double d;
        if (k == T_FLOAT || k == T_DOUBLE) {
            if (!read_f64(k, d)) return false;
            out = d > 0 ? (size_t)d : 0;  // `as usize` saturates
            return true;
        }
        switch (k) {  // signed -> usize sign-extends in Rust (`i8 as usize`)
            case T_CHAR: { int8_t v; if (!raw(v)) return false; out = (size_t)(int64_t)v; return true; }
            case T_SHORT: { int16_t v; if (!raw(v)) return false; out = (size_t)(int64_t)v; return true; }
            case T_INT: { int32_t v; if (!raw(v)) return false; out = (size_t)(int64_t)v; return true; }
            case T_UCHAR: { uint8_t v; if (!raw(v)) return false; out = v; return true; }
            case T_USHORT: { uint16_t v; if (!raw(v)) return false; out = v; return true; }
            case T_UINT: { uint32_t v; if (!raw(v)) return false; out = v; return true; }
            default: return false;
        }
    }
    bool skip(PlyType k) {  // :112-151
        if (fmt == PLY_ASCII) { std::string w; return word(w); }
        char b[8];
        return (bool)in.read(b, (std::streamsize)ply_size(k));
    }
};
static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
static std::vector<std::string> split_space(const std::string& s) {  // str::split(' ') keeps empty pieces
    std::vector<std::string> out;
    size_t start = 0;
    for (;;) {
        size_t p = s.find(' ', start);
        if (p == std::string::npos) { out.push_back(s.substr(start)); break; }
        out.push_back(s.substr(start, p - start));
        start = p + 1;
    }
    return out;
}
// PlyLoader::load ply_loader.rs:273-430 ; returns vertex triples (after perm) or an error string
static bool ply_load(const std::string& path, const int perm[3], std::vector<V3>& tri_verts, F* max_abs, std::string& err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "cannot open " + path; return false; }
    std::string line;
    if (!std::getline(in, line) || trim(line) != "ply") { err = "ply magic number not found"; return false; }
    PlyFormat fmt = PLY_ASCII;
    std::vector<PlyElem> elems;
    bool reading = true;
    while (reading) {
        if (!std::getline(in, line)) { line.clear(); if (in.eof()) { err = "unexpected end of ply header"; return false; } }
        std::vector<std::string> sp = split_space(trim(line));
        const std::string cmd = sp.empty() ? std::string() : sp[0];
        if (cmd == "end_header") reading = false;
        else if (cmd == "format") {
            std::string f = sp.size() > 1 ? sp[1] : "", v = sp.size() > 2 ? sp[2] : "";
            if (sp.size() > 2 && f == "ascii" && v == "1.0") fmt = PLY_ASCII;
            else if (sp.size() > 2 && f == "binary_little_endian" && v == "1.0") fmt = PLY_LE;
            else if (sp.size() > 2 && f == "binary_big_endian" && v == "1.0") fmt = PLY_BE;
            else { err = "ply unsupported format found: " + f + " " + v; return false; }
        } else if (cmd == "comment") {
        } else if (cmd == "element") {
            if (sp.size() < 3) { err = "ply invalid element: '" + line + "'"; return false; }
            char* end = nullptr;
            unsigned long long cnt = std::strtoull(sp[2].c_str(), &end, 10);
            if (!end || *end != 0 || sp[2].empty() || sp[2][0] == '-') { err = "ply invalid element: '" + line + "'"; return false; }
            elems.push_back({sp[1], (size_t)cnt, {}});
        } else if (cmd == "property") {
            if (sp.size() < 2) continue;
            if (sp[1] == "list") {
                PlyType ck = sp.size() > 2 ? ply_type(sp[2]) : T_BAD, pk = sp.size() > 3 ? ply_type(sp[3]) : T_BAD;
                if (sp.size() < 5 || ck == T_BAD || pk == T_BAD) { err = "ply invalid property: '" + line + "'"; return false; }
                if (!elems.empty()) elems.back().props.push_back({true, sp[4], pk, ck});
            } else {
                PlyType k = ply_type(sp[1]);
                if (sp.size() < 3 || k == T_BAD) { err = "ply invalid property: '" + line + "'"; return false; }
                if (!elems.empty()) elems.back().props.push_back({false, sp[2], k, T_BAD});
            }
        } else if (!cmd.empty()) {
            std::fprintf(stderr, "unknown ply header found: '%s'\n", cmd.c_str());
        }
    }
    PlyReader rd{in, fmt};
    std::vector<V3> vertexes;
    F mabs = 0.0f;
    for (const PlyElem& e : elems) {
        bool is_vertex = e.name == "vertex", is_face = e.name == "face";
        for (size_t i = 0; i < e.count; ++i) {
            bool hx = false, hy = false, hz = false;
            F xyz[3] = {0, 0, 0};
            for (const PlyProp& p : e.props) {
                if (!p.is_list) {
                    if (is_vertex && p.name == "x") { if (!rd.read_f32(p.kind, xyz[0])) { err = "ply read error"; return false; } hx = true; }
                    else if (is_vertex && p.name == "y") { if (!rd.read_f32(p.kind, xyz[1])) { err = "ply read error"; return false; } hy = true; }
                    else if (is_vertex && p.name == "z") { if (!rd.read_f32(p.kind, xyz[2])) { err = "ply read error"; return false; } hz = true; }
                    else if (!rd.skip(p.kind)) { err = "ply read error"; return false; }
                } else {
                    size_t count;
                    if (!rd.read_usize(p.count_kind, count)) { err = "ply read error"; return false; }
                    if (is_face && count == 3) {
                        size_t a, b, c;
                        if (!rd.read_usize(p.kind, a) || !rd.read_usize(p.kind, b) || !rd.read_usize(p.kind, c)) { err = "ply read error"; return false; }
                        if (a >= vertexes.size() || b >= vertexes.size() || c >= vertexes.size()) { err = "ply face index out of bounds"; return false; }
                        tri_verts.push_back(vertexes[a]);
                        tri_verts.push_back(vertexes[b]);
                        tri_verts.push_back(vertexes[c]);
                    } else {
                        for (size_t k = 0; k < count; ++k)
                            if (!rd.skip(p.kind)) { err = "ply read error"; return false; }
                    }
                }
            }
            if (is_vertex && hx && hy && hz) {
                mabs = fmax_(fmax_(fmax_(mabs, std::fabs(xyz[0])), std::fabs(xyz[1])), std::fabs(xyz[2]));  // scenes/lucy.rs:37
                vertexes.push_back({xyz[perm[0]], xyz[perm[1]], xyz[perm[2]]});
            }
        }
    }
    if (max_abs) *max_abs = mabs;
    return true;
}

// ---------------------------------------------------------------------------------------------
// stl_loader.rs:10-66 — binary STL: 80-byte header, u32 count, per triangle 12 floats (normal ignored) + u16 attribute byte count
// ---------------------------------------------------------------------------------------------
static bool stl_load_binary(const std::string& path, const int perm[3], std::vector<V3>& tri_verts, std::string& err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "cannot open " + path; return false; }
    char header[80];
    if (!in.read(header, 80)) { err = "stl read error"; return false; }
    uint32_t tri_count = 0;
    if (!in.read((char*)&tri_count, 4)) { err = "stl read error"; return false; }
    for (uint32_t i = 0; i < tri_count; ++i) {
        float f[12];
        if (!in.read((char*)f, 48)) { err = "stl read error"; return false; }
        for (int k = 0; k < 3; ++k) {
            const float* v = f + 3 + 3 * k;
            tri_verts.push_back({v[perm[0]], v[perm[1]], v[perm[2]]});
        }
        uint16_t attr = 0;
        if (!in.read((char*)&attr, 2)) { err = "stl read error"; return false; }
        if (attr) {
            std::vector<char> skip(attr);
            if (!in.read(skip.data(), attr)) { err = "stl read error"; return false; }
        }
    }
    return true;
}


// ---------------------------------------------------------------------------------------------
// obj_loader.rs — ObjLoader::load (:329-452) over the ObjBuilder trait (:23-43), with the two builders the reference ships:
// SimpleTexturedBuilder (:160-308) and obj_fns / FnObjBuilder (:45-158). PNG decoding (`image` crate, texture.rs:36-39) is
// not restated: decoded RGBA8 images are registered by path beforehand (orc_register_png) and load_png looks them up.
// ---------------------------------------------------------------------------------------------
struct ObjContext {  // :310-327
    bool has_group = false, has_material = false, has_library = false;
    std::string group_name, material_name, material_library;
};
struct ObjCorner { V3 vertex, normal; V2 uv; };
struct ObjBuilder {
    virtual ~ObjBuilder() {}
    virtual bool include_group(const ObjContext&) { return true; }
    virtual void load_materials(const ObjContext&) {}
    virtual V2 build_uv(F x, F y) = 0;
    virtual bool build_face(const ObjContext&, const ObjCorner& a, const ObjCorner& b, const ObjCorner& c, std::vector<Triangle>& out, std::string& err) = 0;
};
static std::vector<std::string> split_whitespace(const std::string& s) {
    std::vector<std::string> out;
    std::string cur;
    for (char ch : s) {
        if (std::isspace((unsigned char)ch)) { if (!cur.empty()) { out.push_back(cur); cur.clear(); } }
        else cur.push_back(ch);
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}
static bool rust_parse_f32(const std::string& t, F& out) {  // str::parse::<f32>: decimal literal, inf / infinity / nan, no hex, no blanks
    if (t.empty()) return false;
    size_t i = (t[0] == '+' || t[0] == '-') ? 1 : 0;
    std::string body;
    for (size_t k = i; k < t.size(); ++k) body.push_back((char)std::tolower((unsigned char)t[k]));
    if (body == "inf" || body == "infinity") { out = t[0] == '-' ? -INF : INF; return true; }
    if (body == "nan") { out = std::numeric_limits<F>::quiet_NaN(); return true; }
    bool digit = false, dot = false, exp = false, exp_digit = false;
    for (size_t k = 0; k < body.size(); ++k) {
        char ch = body[k];
        if (std::isdigit((unsigned char)ch)) { (exp ? exp_digit : digit) = true; }
        else if (ch == '.' && !dot && !exp) dot = true;
        else if (ch == 'e' && digit && !exp) { exp = true; if (k + 1 < body.size() && (body[k + 1] == '+' || body[k + 1] == '-')) ++k; }
        else return false;
    }
    if (!digit || (exp && !exp_digit)) return false;
    out = std::strtof(t.c_str(), nullptr);
    return true;
}
static bool rust_parse_usize(const std::string& t, size_t& out) {
    size_t i = (!t.empty() && t[0] == '+') ? 1 : 0;
    if (i >= t.size()) return false;
    unsigned long long v = 0;
    for (; i < t.size(); ++i) {
        if (!std::isdigit((unsigned char)t[i])) return false;
        unsigned long long d = (unsigned long long)(t[i] - '0');
        if (v > (~0ull - d) / 10) return false;
        v = v * 10 + d;
    }
    out = (size_t)v;
    return true;
}
static std::string path_with_file_name(const std::string& path, const std::string& name) {
    size_t slash = path.rfind('/');
    return slash == std::string::npos ? name : path.substr(0, slash + 1) + name;
}
static bool obj_load(const std::string& path, ObjBuilder& builder, std::vector<Triangle>& faces, std::string& err) {  // :332-452
    std::ifstream file(path, std::ios::binary);
    if (!file) { err = "cannot open " + path; return false; }
    std::vector<V3> vertexes, normals;
    std::vector<V2> uvs;
    ObjContext context;
    bool include_faces = builder.include_group(context);
    std::string line;
    while (std::getline(file, line)) {
        std::vector<std::string> parts = split_whitespace(line);
        if (parts.empty()) continue;
        const std::string& head = parts[0];
        auto num = [&](size_t i, F& v) { return i < parts.size() && rust_parse_f32(parts[i], v); };
        if (head == "v") {  // :355-366
            F x, y, z;
            if (!(num(1, x) && num(2, y) && num(3, z))) { err = "unable to parse vertex: " + trim(line); return false; }
            vertexes.push_back({x, y, z});
        } else if (head == "vn") {  // :367-378
            F x, y, z;
            if (!(num(1, x) && num(2, y) && num(3, z))) { err = "unable to parse normal: " + trim(line); return false; }
            normals.push_back({x, y, z});
        } else if (head == "vt") {  // :379-389
            F u, v;
            if (!(num(1, u) && num(2, v))) { err = "unable to parse texture coord: " + trim(line); return false; }
            uvs.push_back(builder.build_uv(u, v));
        } else if (head == "f") {  // :390-429
            if (!include_faces) continue;
            auto read_face = [&](size_t pi, ObjCorner& out) -> bool {
                if (pi >= parts.size()) return false;
                const std::string& s = parts[pi];
                std::vector<size_t> splits;  // s.split('/').filter_map(|n| n.parse::<usize>().ok())
                size_t start = 0;
                for (;;) {
                    size_t p = s.find('/', start);
                    size_t v;
                    if (rust_parse_usize(s.substr(start, p == std::string::npos ? std::string::npos : p - start), v)) splits.push_back(v);
                    if (p == std::string::npos) break;
                    start = p + 1;
                }
                auto get = [](auto& vec, size_t one_based) { return one_based >= 1 && one_based <= vec.size() ? &vec[one_based - 1] : nullptr; };
                const V3* v = nullptr;
                const V3* n = nullptr;
                const V2* uv = nullptr;
                if (s.find("//") != std::string::npos) {  // :399-407: vertex, uvs.get(0), normal
                    if (splits.size() > 0) v = get(vertexes, splits[0]);
                    uv = uvs.empty() ? nullptr : &uvs[0];
                    if (splits.size() > 1) n = get(normals, splits[1]);
                } else {  // :408-416
                    if (splits.size() > 0) v = get(vertexes, splits[0]);
                    if (splits.size() > 1) uv = get(uvs, splits[1]);
                    if (splits.size() > 2) n = get(normals, splits[2]);
                }
                if (!v || !n || !uv) return false;
                out = {*v, *n, *uv};
                return true;
            };
            ObjCorner a, b, c;
            if (!(read_face(1, a) && read_face(2, b) && read_face(3, c))) { err = "unable to parse face: " + trim(line); return false; }
            if (!builder.build_face(context, a, b, c, faces, err)) return false;
        } else if (head == "o" || head == "g") {  // :430-435
            if (parts.size() > 1) {
                context.group_name = parts[1];
                context.has_group = true;
                include_faces = builder.include_group(context);
            }
        } else if (head == "usemtl") {  // :436-440
            if (parts.size() > 1) { context.material_name = parts[1]; context.has_material = true; }
        } else if (head == "mtllib") {  // :441-445
            std::string joined;
            for (size_t i = 1; i < parts.size(); ++i) { if (i > 1) joined += " "; joined += parts[i]; }
            context.material_library = path_with_file_name(path, joined);
            context.has_library = true;
            builder.load_materials(context);
        }
    }
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// scene container behind the C API
// ---------------------------------------------------------------------------------------------
struct orc_scene {
    Rng rng{0};
    std::string err;
    std::vector<std::unique_ptr<Surface>> surfaces;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<Mesh>> meshes;
    std::vector<std::unique_ptr<Intersect>> objects;  // World::add order
    World world;
    Camera camera;
    bool bvh_built = false;
    struct Png { std::string path; std::vector<uint8_t> rgba; uint32_t w, h; };
    std::vector<Png> pngs;  // images decoded by the caller, looked up by Texture::load_png's path (orc_register_png)
};

namespace {
inline const Surface* surf(orc_scene* s, int i) { return s->surfaces.at((size_t)i).get(); }
inline const Material* mat(orc_scene* s, int i) { return i < 0 ? nullptr : s->materials.at((size_t)i).get(); }
inline int push_surface(orc_scene* s, Surface* p) { s->surfaces.emplace_back(p); return (int)s->surfaces.size() - 1; }
inline int push_material(orc_scene* s, Material* p) { s->materials.emplace_back(p); return (int)s->materials.size() - 1; }
inline V3 v3(const float* p) { return {p[0], p[1], p[2]}; }
inline void put3(float* o, V3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
inline void put_m4(float* o, const M4& m) {
    const V4* c[4] = {&m.c0, &m.c1, &m.c2, &m.c3};
    for (int i = 0; i < 4; ++i) { o[4 * i] = c[i]->x; o[4 * i + 1] = c[i]->y; o[4 * i + 2] = c[i]->z; o[4 * i + 3] = c[i]->w; }
}
int finish_mesh(orc_scene* s, Mesh* m) {
    std::vector<const Intersect*> items;
    items.reserve(m->tris.size());
    for (size_t i = 0; i < m->tris.size(); ++i) { m->tris[i].tri = (uint32_t)i; items.push_back(&m->tris[i]); }
    if (items.empty()) { s->err = "mesh has no triangles"; delete m; return -1; }
    m->bvh.reset(new BvhNode(std::move(items), s->rng));  // Model::new builds the BLAS immediately geom.rs:281-292
    s->meshes.emplace_back(m);
    return (int)s->meshes.size() - 1;
}
int add_object(orc_scene* s, Intersect* o) {
    s->objects.emplace_back(o);
    s->world.objects.push_back(o);
    return (int)s->objects.size() - 1;
}
int resolve_threads(int threads) {  // main.rs:159-160
    if (threads > 0) return threads;
    int cpus = (int)std::thread::hardware_concurrency();
    return std::max(cpus - 2, 1);
}
void add_counters(orc_counters& a, const orc_counters& b) {
    a.rays += b.rays; a.paths += b.paths; a.box_tests += b.box_tests; a.tri_tests += b.tri_tests;
    a.sphere_tests += b.sphere_tests; a.instance_tests += b.instance_tests; a.volume_tests += b.volume_tests;
}
}  // namespace

namespace {
// Texture::load_png texture.rs:29-68 over a pre-decoded image
const Surface* load_png(orc_scene* s, const std::string& path, int wrapping, std::string& err) {
    for (const orc_scene::Png& p : s->pngs)
        if (p.path == path) { push_surface(s, new Texture(p.rgba.data(), p.w, p.h, wrapping)); return s->surfaces.back().get(); }
    err = "cannot open " + path;
    return nullptr;
}
struct SimpleTexturedBuilder : ObjBuilder {  // obj_loader.rs:160-308
    orc_scene* s;
    int wrapping;
    std::vector<std::pair<std::string, const Surface*>> textures;
    std::vector<std::pair<std::string, V3>> diffuse;
    std::vector<std::string> filtered_groups;
    template <class T>
    static const T* lookup(const std::vector<std::pair<std::string, T>>& m, const std::string& k) {
        for (auto it = m.rbegin(); it != m.rend(); ++it)  // HashMap::insert replaces: the last entry wins
            if (it->first == k) return &it->second;
        return nullptr;
    }
    bool process_material_library(const std::string& path, std::string& err) {  // :188-233
        std::ifstream file(path, std::ios::binary);
        if (!file) { err = "cannot open " + path; return false; }
        std::string line, current_material;
        bool has_current = false;
        while (std::getline(file, line)) {
            std::vector<std::string> parts = split_whitespace(trim(line));
            if (parts.empty()) continue;
            if (parts[0] == "newmtl") {
                if (parts.size() > 1) { current_material = parts[1]; has_current = true; }
            } else if (parts[0] == "Kd") {
                if (has_current) {
                    F x, y, z;
                    if (parts.size() > 3 && rust_parse_f32(parts[1], x) && rust_parse_f32(parts[2], y) && rust_parse_f32(parts[3], z))
                        diffuse.push_back({current_material, V3{x, y, z}});
                }
            } else if (parts[0] == "map_Kd") {
                if (parts.size() > 1 && has_current) {
                    const Surface* t = load_png(s, path_with_file_name(path, parts[1]), wrapping, err);
                    if (!t) return false;  // `?` :226
                    textures.push_back({current_material, t});
                }
            }
        }
        return true;
    }
    void load_materials(const ObjContext& context) override {  // :261-268
        std::string err;
        if (context.has_library && !process_material_library(context.material_library, err))
            std::fprintf(stderr, "unable to load material library: %s\n", err.c_str());
    }
    V2 build_uv(F x, F y) override { return {x, 1.0f - y}; }  // :281-283
    bool build_face(const ObjContext& context, const ObjCorner& a, const ObjCorner& b, const ObjCorner& c, std::vector<Triangle>& out, std::string& err) override {  // :285-298
        const Surface* surface = nullptr;
        if (context.has_material) {
            if (const Surface* const* t = lookup(textures, context.material_name)) surface = *t;
            else if (const V3* d = lookup(diffuse, context.material_name)) {
                push_surface(s, new SolidColor(expand(*d, 1.0f)));
                surface = s->surfaces.back().get();
            }
        }
        if (!surface) { err = "No material found for face"; return false; }
        push_material(s, new Lambertian(surface));
        out.emplace_back(s->materials.back().get(), a.vertex, a.normal, a.uv, b.vertex, b.normal, b.uv, c.vertex, c.normal, c.uv);
        return true;
    }
    bool include_group(const ObjContext& context) override {  // :300-306
        if (!context.has_group) return true;
        return std::find(filtered_groups.begin(), filtered_groups.end(), context.group_name) == filtered_groups.end();
    }
};
struct FnObjBuilder : ObjBuilder {  // obj_fns(V3::new, V3::new, V2::new, |a, b, c| Triangle::with_norms_and_uvs(material, a, b, c))  :45-158, eve.rs:330-340
    const Material* material;
    V2 build_uv(F x, F y) override { return {x, y}; }
    bool build_face(const ObjContext&, const ObjCorner& a, const ObjCorner& b, const ObjCorner& c, std::vector<Triangle>& out, std::string&) override {
        out.emplace_back(material, a.vertex, a.normal, a.uv, b.vertex, b.normal, b.uv, c.vertex, c.normal, c.uv);
        return true;
    }
};
int mesh_from_faces(orc_scene* s, std::vector<Triangle>& faces) {
    Mesh* m = new Mesh();
    m->tris = std::move(faces);
    return finish_mesh(s, m);
}
}  // namespace


extern "C" {

orc_scene* orc_scene_new(void) {
    orc_scene* s = new orc_scene();
    s->world.background.reset(new SolidBackground({0, 0, 0}));
    return s;
}
void orc_scene_free(orc_scene* s) { delete s; }
const char* orc_last_error(orc_scene* s) { return s->err.c_str(); }
void orc_seed(orc_scene* s, uint64_t seed) { s->rng = Rng(seed); }
float orc_rand_f32(orc_scene* s) { return s->rng.f32(); }

int orc_surface_solid(orc_scene* s, float r, float g, float b, float a) { return push_surface(s, new SolidColor({r, g, b, a})); }
int orc_surface_texture(orc_scene* s, const uint8_t* rgba, uint32_t w, uint32_t h, int wrap_) { return push_surface(s, new Texture(rgba, w, h, wrap_)); }
int orc_surface_ycbcr(orc_scene* s, int l, int c) {
    const Texture* tl = dynamic_cast<const Texture*>(surf(s, l));
    const Texture* tc = dynamic_cast<const Texture*>(surf(s, c));
    if (!tl || !tc) { s->err = "ycbcr needs two Texture surfaces"; return -1; }
    return push_surface(s, new YCbCrTexture(tl, tc));
}
int orc_surface_blend(orc_scene* s, int mode, int l, int r) { return push_surface(s, new TextureBlend(mode, surf(s, l), surf(s, r))); }
int orc_surface_fallback(orc_scene* s, float r, float g, float b, float a, int inner) { return push_surface(s, new SolidColorFallback({r, g, b, a}, surf(s, inner))); }

int orc_mat_absorb(orc_scene* s) { return push_material(s, new Absorb()); }
int orc_mat_lambertian(orc_scene* s, int surface) { return push_material(s, new Lambertian(surf(s, surface))); }
int orc_mat_diffuse_light(orc_scene* s, float r, float g, float b) { return push_material(s, new DiffuseLight({r, g, b})); }
int orc_mat_metal(orc_scene* s, float fuzz, int surface) { return push_material(s, new Metal(fuzz, surf(s, surface))); }
int orc_mat_dielectric(orc_scene* s, float ior) { return push_material(s, new Dielectric(ior)); }
int orc_mat_specular(orc_scene* s, float ior, int surface) { return push_material(s, new Specular(ior, surf(s, surface))); }
int orc_mat_mix(orc_scene* s, float ratio, int l, int r) { return push_material(s, new Mix(ratio, mat(s, l), mat(s, r))); }
int orc_mat_isotropic(orc_scene* s, float r, float g, float b) { return push_material(s, new Isotrophic({r, g, b})); }
int orc_mat_eve(orc_scene* s, int no, int ar, int pmdg, const float colors12[12], const float glow3[3]) {
    V3 c[4];
    for (int i = 0; i < 4; ++i) c[i] = {colors12[3 * i], colors12[3 * i + 1], colors12[3 * i + 2]};
    return push_material(s, new EveMaterial(surf(s, no), surf(s, ar), surf(s, pmdg), c, {glow3[0], glow3[1], glow3[2]}));
}

void orc_background_solid(orc_scene* s, float r, float g, float b) { s->world.background.reset(new SolidBackground({r, g, b})); }
void orc_background_sky(orc_scene* s) { s->world.background.reset(new SkyBackground()); }
void orc_background_skysphere(orc_scene* s, int surface) { s->world.background.reset(new SkySphere(surf(s, surface))); }
void orc_background_cubemap(orc_scene* s, const int f[6], float rx, float ry, float rz) {
    const Surface* p[6];
    for (int i = 0; i < 6; ++i) p[i] = surf(s, f[i]);
    s->world.background.reset(new CubeMap(p, {rx, ry, rz}));
}

int orc_mesh_new(orc_scene* s, const float* v, uint64_t n, int tri_material) {
    Mesh* m = new Mesh();
    m->tris.reserve(n);
    const Material* tm = mat(s, tri_material);
    for (uint64_t i = 0; i < n; ++i) m->tris.emplace_back(tm, v3(v + 9 * i), v3(v + 9 * i + 3), v3(v + 9 * i + 6));
    return finish_mesh(s, m);
}
int orc_mesh_new_uv(orc_scene* s, const float* v, const float* nn, const float* uv, uint64_t n, int tri_material) {
    Mesh* m = new Mesh();
    m->tris.reserve(n);
    const Material* tm = mat(s, tri_material);
    for (uint64_t i = 0; i < n; ++i)
        m->tris.emplace_back(tm, v3(v + 9 * i), v3(nn + 9 * i), V2{uv[6 * i], uv[6 * i + 1]}, v3(v + 9 * i + 3), v3(nn + 9 * i + 3),
                             V2{uv[6 * i + 2], uv[6 * i + 3]}, v3(v + 9 * i + 6), v3(nn + 9 * i + 6), V2{uv[6 * i + 4], uv[6 * i + 5]});
    return finish_mesh(s, m);
}
void orc_register_png(orc_scene* s, const char* path, const uint8_t* rgba, uint32_t w, uint32_t h) {
    s->pngs.push_back({path, std::vector<uint8_t>(rgba, rgba + (size_t)w * h * 4), w, h});
}
int orc_surface_texture_png(orc_scene* s, const char* path, int wrap_) {
    return load_png(s, path, wrap_, s->err) ? (int)s->surfaces.size() - 1 : -1;
}
int orc_mesh_load_obj(orc_scene* s, const char* path, int wrap_, const char* filtered_groups) {
    SimpleTexturedBuilder b;
    b.s = s;
    b.wrapping = wrap_;
    if (filtered_groups) {
        std::string cur;
        for (const char* p = filtered_groups;; ++p) {
            if (*p == '\n' || *p == 0) { if (!cur.empty()) b.filtered_groups.push_back(cur); cur.clear(); if (!*p) break; }
            else cur.push_back(*p);
        }
    }
    std::vector<Triangle> faces;
    if (!obj_load(path, b, faces, s->err)) return -1;
    return mesh_from_faces(s, faces);
}
int orc_mesh_load_obj_with(orc_scene* s, const char* path, int tri_material) {
    FnObjBuilder b;
    b.material = mat(s, tri_material);
    std::vector<Triangle> faces;
    if (!obj_load(path, b, faces, s->err)) return -1;
    return mesh_from_faces(s, faces);
}
void orc_mesh_get_shading(orc_scene* s, int mesh, float* normals9, float* uvs6, int32_t* materials) {
    const Mesh& m = *s->meshes.at((size_t)mesh);
    for (size_t i = 0; i < m.tris.size(); ++i) {
        const Triangle& t = m.tris[i];
        if (normals9) { put3(normals9 + 9 * i, t.na); put3(normals9 + 9 * i + 3, t.nb); put3(normals9 + 9 * i + 6, t.nc); }
        if (uvs6) { uvs6[6 * i] = t.uv_a.x; uvs6[6 * i + 1] = t.uv_a.y; uvs6[6 * i + 2] = t.uv_b.x; uvs6[6 * i + 3] = t.uv_b.y; uvs6[6 * i + 4] = t.uv_c.x; uvs6[6 * i + 5] = t.uv_c.y; }
        if (materials) {
            materials[i] = -1;
            for (size_t k = 0; k < s->materials.size(); ++k)
                if (s->materials[k].get() == t.material) { materials[i] = (int32_t)k; break; }
        }
    }
}
int orc_material_info(orc_scene* s, int material, float color4[4], uint32_t wh[2], uint64_t* texel_hash) {
    const Material* m = mat(s, material);
    color4[0] = color4[1] = color4[2] = color4[3] = 0.0f;
    wh[0] = wh[1] = 0;
    *texel_hash = 0;
    const Surface* surface = nullptr;
    int kind = -1;  // numbering of include/mrt.h's MRT_MAT_* for the kinds an OBJ can produce
    if (auto* l = dynamic_cast<const Lambertian*>(m)) { kind = 1; surface = l->surface; }
    else if (dynamic_cast<const Absorb*>(m)) kind = 0;
    if (auto* c = dynamic_cast<const SolidColor*>(surface)) { color4[0] = c->c.x; color4[1] = c->c.y; color4[2] = c->c.z; color4[3] = c->c.w; }
    if (auto* t = dynamic_cast<const Texture*>(surface)) {
        wh[0] = t->w; wh[1] = t->h;
        uint64_t h = 1469598103934665603ull;  // FNV-1a over the f32 texels
        const unsigned char* p = reinterpret_cast<const unsigned char*>(t->pixels.data());
        for (size_t i = 0; i < t->pixels.size() * 16; ++i) { h ^= p[i]; h *= 1099511628211ull; }
        *texel_hash = h;
    }
    return kind;
}
int orc_mesh_load_ply(orc_scene* s, const char* path, const int perm[3], int tri_material, float* max_abs) {
    std::vector<V3> tv;
    if (!ply_load(path, perm, tv, max_abs, s->err)) return -1;
    Mesh* m = new Mesh();
    const Material* tm = mat(s, tri_material);
    m->tris.reserve(tv.size() / 3);
    for (size_t i = 0; i + 2 < tv.size(); i += 3) m->tris.emplace_back(tm, tv[i], tv[i + 1], tv[i + 2]);
    return finish_mesh(s, m);
}
int orc_mesh_load_stl(orc_scene* s, const char* path, const int perm[3], int tri_material) {
    std::vector<V3> tv;
    if (!stl_load_binary(path, perm, tv, s->err)) return -1;
    Mesh* m = new Mesh();
    const Material* tm = mat(s, tri_material);
    m->tris.reserve(tv.size() / 3);
    for (size_t i = 0; i + 2 < tv.size(); i += 3) m->tris.emplace_back(tm, tv[i], tv[i + 1], tv[i + 2]);
    return finish_mesh(s, m);
}
uint64_t orc_mesh_tri_count(orc_scene* s, int mesh) { return s->meshes.at((size_t)mesh)->tris.size(); }
void orc_mesh_get_verts(orc_scene* s, int mesh, float* out) {
    const Mesh& m = *s->meshes.at((size_t)mesh);
    for (size_t i = 0; i < m.tris.size(); ++i) { put3(out + 9 * i, m.tris[i].va); put3(out + 9 * i + 3, m.tris[i].vb); put3(out + 9 * i + 6, m.tris[i].vc); }
}
uint64_t orc_mesh_node_count(orc_scene* s, int mesh) { return s->meshes.at((size_t)mesh)->bvh->node_count; }

int orc_add_sphere(orc_scene* s, int material, float cx, float cy, float cz, float radius) {
    Sphere* sp = new Sphere(mat(s, material), {cx, cy, cz}, radius);
    int id = add_object(s, sp);
    sp->object = (uint32_t)id;
    return id;
}
int orc_add_model(orc_scene* s, int mesh, int override_material) {
    Model* m = new Model(s->meshes.at((size_t)mesh)->bvh.get(), mat(s, override_material));
    int id = add_object(s, m);
    m->object = (uint32_t)id;
    return id;
}
int orc_add_instance(orc_scene* s, int mesh, const float t[3], const float r[3], const float sc[3], int override_material) {
    Instance* in = new Instance(s->meshes.at((size_t)mesh)->bvh.get(), v3(t), v3(r), v3(sc), mat(s, override_material));
    int id = add_object(s, in);
    in->object = (uint32_t)id;
    return id;
}
int orc_add_volume_sphere(orc_scene* s, float cx, float cy, float cz, float radius, float density, float r, float g, float b) {
    static Absorb unit_material;  // Sphere<()> scenes/eve.rs:41-45
    Volume* v = new Volume(new Sphere(&unit_material, {cx, cy, cz}, radius), density, {r, g, b});
    int id = add_object(s, v);
    v->object = (uint32_t)id;
    return id;
}
int orc_add_volume_model(orc_scene* s, int mesh, float density, float r, float g, float b) {  // Volume::new(Model::new(..), ..) geom.rs:601-609
    Volume* v = new Volume(new Model(s->meshes.at((size_t)mesh)->bvh.get(), nullptr), density, {r, g, b});
    int id = add_object(s, v);
    v->object = (uint32_t)id;
    return id;
}
int orc_add_volume_instance(orc_scene* s, int mesh, const float t[3], const float rot[3], const float sc[3], float density, float r, float g, float b) {
    Volume* v = new Volume(new Instance(s->meshes.at((size_t)mesh)->bvh.get(), v3(t), v3(rot), v3(sc), nullptr), density, {r, g, b});
    int id = add_object(s, v);
    v->object = (uint32_t)id;
    return id;
}
void orc_build_bvh(orc_scene* s) {
    if (s->world.objects.empty()) return;
    s->world.build_bvh(s->rng);
    s->bvh_built = true;
}
uint64_t orc_tlas_node_count(orc_scene* s) { return s->world.tlas ? s->world.tlas->node_count : 0; }
void orc_camera(orc_scene* s, float vfov, const float from[3], const float at[3], const float up[3], float aspect, float aperture, float focus) {
    s->camera = Camera(vfov, v3(from), v3(at), v3(up), aspect, aperture, focus);
}
void orc_get_camera(orc_scene* s, float o[19]) {
    const Camera& c = s->camera;
    put3(o, c.origin); put3(o + 3, c.lower_left_corner); put3(o + 6, c.horizontal); put3(o + 9, c.vertical); put3(o + 12, c.u); put3(o + 15, c.v);
    o[18] = c.lens_radius;
}
void orc_get_instance(orc_scene* s, int object, float tf[16], float inv[16], float aabb[6]) {
    const Instance* in = dynamic_cast<const Instance*>(s->objects.at((size_t)object).get());
    if (!in) { s->err = "object is not an Instance"; return; }
    put_m4(tf, in->transform_);
    put_m4(inv, in->inv_transform);
    put3(aabb, in->bbox.minimum);
    put3(aabb + 3, in->bbox.maximum);
}
void orc_get_object_aabb(orc_scene* s, int object, float aabb[6]) {
    BoundingBox b;
    s->objects.at((size_t)object)->bounding_box(b);
    put3(aabb, b.minimum);
    put3(aabb + 3, b.maximum);
}

void orc_render_aov(orc_scene* s, uint32_t w, uint32_t h, uint64_t seed, int threads, float* albedo, float* normal, uint32_t* object,
                    uint32_t* tri, float* t, orc_counters* counters) {
    int T = resolve_threads(threads);
    std::atomic<uint32_t> row{0};  // dynamic row scheduling main.rs:184, 206
    std::mutex mu;
    orc_counters total{};
    std::vector<std::thread> pool;
    for (int i = 0; i < T; ++i) {
        pool.emplace_back([&]() {
            orc_counters cnt{};
            tl_cnt = &cnt;
            uint32_t y = row.fetch_add(1);
            while (y < h) {
                Rng rng(splitmix(splitmix(seed ^ 0xA0F0A0F0ull) + y));  // per-row stream: output independent of thread count
                tl_rng = &rng;
                for (uint32_t x = 0; x < w; ++x) {
                    F u = (F)x / (F)(w - 1);  // main.rs:189-190 (no jitter)
                    F v = (F)y / (F)(h - 1);
                    Ray ray = s->camera.ray(u, v);
                    V3 a, n;
                    uint32_t ob, tr;
                    F tt;
                    albedo_normal(s->world, ray, a, n, ob, tr, tt);
                    size_t p = (size_t)y * w + x;
                    if (albedo) put3(albedo + 3 * p, a);
                    if (normal) put3(normal + 3 * p, n);
                    if (object) object[p] = ob;
                    if (tri) tri[p] = tr;
                    if (t) t[p] = tt;
                    cnt.paths++;
                }
                y = row.fetch_add(1);
            }
            tl_cnt = nullptr;
            tl_rng = nullptr;
            std::lock_guard<std::mutex> g(mu);
            add_counters(total, cnt);
        });
    }
    for (auto& th : pool) th.join();
    if (counters) *counters = total;
}

void orc_render(orc_scene* s, uint32_t w, uint32_t h, uint32_t spp, uint32_t max_depth, uint64_t seed, int threads, float* sum_rgb,
                uint32_t* sum_bounces, orc_counters* counters) {
    const bool by_rows = threads < 0;  // harness addition, see below
    int T = resolve_threads(by_rows ? 0 : threads);
    size_t npx = (size_t)w * h;
    for (size_t i = 0; i < npx * 3; ++i) sum_rgb[i] = 0.0f;  // image.clear() main.rs:233
    for (size_t i = 0; i < npx; ++i) sum_bounces[i] = 0;
    if (by_rows) {
        // threads < 0: the same frames, distributed over the host's cores by IMAGE ROW instead of by frame, so that a render of a
        // few samples per pixel at a large resolution (the full-size parity tests) uses every core. Each (frame, row) has its own
        // generator stream and each pixel still receives its frames in frame order: the output does not depend on the thread count.
        std::atomic<uint32_t> next_row{0};
        std::mutex mu;
        orc_counters total{};
        std::vector<std::thread> pool;
        for (int i = 0; i < T; ++i) {
            pool.emplace_back([&]() {
                orc_counters cnt{};
                tl_cnt = &cnt;
                for (;;) {
                    uint32_t y = next_row.fetch_add(1);
                    if (y >= h) break;
                    for (uint32_t frame = 0; frame < spp; ++frame) {
                        Rng rng(splitmix(splitmix(splitmix(seed) + frame) ^ (0xD1B54A32D192ED03ULL * (uint64_t)(y + 1))));
                        tl_rng = &rng;
                        for (uint32_t x = 0; x < w; ++x) {
                            F u = ((F)x + frand()) / (F)(w - 1);  // main.rs:258-259
                            F v = ((F)y + frand()) / (F)(h - 1);
                            Ray ray = s->camera.ray(u, v);
                            V3 color;
                            uint32_t depth = 0;
                            trace(s->world, ray, max_depth, color, depth);
                            size_t p = (size_t)y * w + x;
                            sum_rgb[3 * p] += color.x;  // buffer.set + merge (main.rs:263, :629-638) for this pixel
                            sum_rgb[3 * p + 1] += color.y;
                            sum_rgb[3 * p + 2] += color.z;
                            sum_bounces[p] += max_depth - depth;
                            cnt.paths++;
                        }
                    }
                }
                tl_cnt = nullptr;
                tl_rng = nullptr;
                std::lock_guard<std::mutex> g(mu);
                add_counters(total, cnt);
            });
        }
        for (auto& th : pool) th.join();
        if (counters) *counters = total;
        return;
    }
    std::atomic<uint32_t> next_frame{0};
    std::mutex mu;
    std::condition_variable cv;
    uint32_t next_merge = 0;
    orc_counters total{};
    std::vector<std::thread> pool;
    for (int i = 0; i < T; ++i) {
        pool.emplace_back([&]() {
            orc_counters cnt{};
            tl_cnt = &cnt;
            std::vector<float> buf_rgb(npx * 3);     // image.buffer(): thread-private frame main.rs:241
            std::vector<uint32_t> buf_depth(npx);
            for (;;) {
                uint32_t frame = next_frame.fetch_add(1);
                if (frame >= spp) break;
                Rng rng(splitmix(splitmix(seed) + frame));  // per-frame stream: output independent of thread count
                tl_rng = &rng;
                for (uint32_t y = 0; y < h; ++y) {
                    for (uint32_t x = 0; x < w; ++x) {
                        F u = ((F)x + frand()) / (F)(w - 1);  // main.rs:258-259
                        F v = ((F)y + frand()) / (F)(h - 1);
                        Ray ray = s->camera.ray(u, v);
                        V3 color;
                        uint32_t depth = 0;
                        trace(s->world, ray, max_depth, color, depth);
                        size_t p = (size_t)y * w + x;
                        put3(&buf_rgb[3 * p], color);         // buffer.set(.., MAX_DEPTH - depth) main.rs:263
                        buf_depth[p] = max_depth - depth;
                        cnt.paths++;
                    }
                }
                // image.merge(&buffer) main.rs:273, :629-638 — serialised in frame order so the f32 sums are reproducible
                std::unique_lock<std::mutex> g(mu);
                cv.wait(g, [&] { return next_merge == frame; });
                for (size_t k = 0; k < npx * 3; ++k) sum_rgb[k] += buf_rgb[k];
                for (size_t k = 0; k < npx; ++k) sum_bounces[k] += buf_depth[k];
                next_merge++;
                g.unlock();
                cv.notify_all();
            }
            tl_cnt = nullptr;
            tl_rng = nullptr;
            std::lock_guard<std::mutex> g(mu);
            add_counters(total, cnt);
        });
    }
    for (auto& th : pool) th.join();
    if (counters) *counters = total;
}

void orc_resolve_rgb8(const float* sum_rgb, uint32_t w, uint32_t h, uint32_t count, int flip, uint8_t* out) {  // main.rs:640-722, 760-768
    F scale = 1.0f / (F)count;
    for (uint32_t y = 0; y < h; ++y) {
        uint32_t sy = flip ? (h - 1 - y) : y;
        for (uint32_t x = 0; x < w * 3; ++x) {
            F p = 0.0f;
            if (count != 0) p = fmax_(fmin_(std::pow(scale * sum_rgb[(size_t)sy * w * 3 + x], 1.0f / 2.2f), 1.0f), 0.0f);
            F v = p * 255.0f;  // `as u8` saturates, NaN -> 0
            out[(size_t)y * w * 3 + x] = (uint8_t)(v >= 255.0f ? 255 : (v > 0.0f ? (int)v : 0));
        }
    }
}

void orc_float_buffer_rgb8(const float* buf, uint32_t w, uint32_t h, int mode, int flip, uint8_t* out) {  // main.rs:694-721, 760-768
    for (uint32_t y = 0; y < h; ++y) {
        uint32_t sy = flip ? (h - 1 - y) : y;
        for (uint32_t x = 0; x < w * 3; ++x) {
            F p = buf[(size_t)sy * w * 3 + x];
            F f = mode == 3 ? std::pow(fmax_(fmin_(p, 1.0f), 0.0f), 1.0f / 2.2f)  // Albedo :694-697
                            : (p + 1.0f) / 2.0f;                                  // Normal :708-711
            F v = f * 255.0f;
            out[(size_t)y * w * 3 + x] = (uint8_t)(v >= 255.0f ? 255 : (v > 0.0f ? (int)v : 0));
        }
    }
}

// ----------------------------- known-answer hooks ----------------------------------------------
int orc_kat_sphere(float cx, float cy, float cz, float r, const float o[3], const float d[3], float tmin, float tmax, float out[8]) {
    static Absorb m;
    Sphere sp(&m, {cx, cy, cz}, r);
    Hit h;
    if (!sp.intersect({v3(o), v3(d)}, tmin, tmax, h)) return 0;
    out[0] = h.t; put3(out + 1, h.point); put3(out + 4, h.normal); out[7] = h.front_face ? 1.0f : 0.0f;
    return 1;
}
int orc_kat_aabb(const float bmin[3], const float bmax[3], const float o[3], const float d[3], float tmin, float tmax) {
    BoundingBox b{v3(bmin), v3(bmax)};
    return b.hit({v3(o), v3(d)}, tmin, tmax) ? 1 : 0;
}
int orc_kat_triangle(const float v[9], const float o[3], const float d[3], float tmin, float tmax, float out[11]) {
    static Absorb m;
    Triangle tr(&m, v3(v), v3(v + 3), v3(v + 6));
    Hit h;
    Ray ray{v3(o), v3(d)};
    if (!tr.intersect(ray, tmin, tmax, h)) return 0;
    out[0] = h.t; put3(out + 1, h.point); put3(out + 4, h.normal); out[7] = h.front_face ? 1.0f : 0.0f;
    V3 d0 = tr.va - h.point, d1 = tr.vb - h.point, d2 = tr.vc - h.point;
    F area = length(cross(tr.va - tr.vb, tr.va - tr.vc));
    out[8] = length(cross(d1, d2)) / area; out[9] = length(cross(d2, d0)) / area; out[10] = length(cross(d0, d1)) / area;
    return 1;
}
void orc_kat_rotate(int axis, float turns, float out[16]) {
    M4 m = axis == 0 ? m4_rotate_x(turns) : (axis == 1 ? m4_rotate_y(turns) : m4_rotate_z(turns));
    put_m4(out, m);
}
void orc_kat_wrap(int mode, float x, float y, float out[2]) { V2 r = wrap(mode, {x, y}); out[0] = r.x; out[1] = r.y; }
float orc_kat_reflectance(float cosine, float ref_idx) { return reflectance(cosine, ref_idx); }
void orc_kat_refract(const float v[3], const float n[3], float eta, float out[3]) { put3(out, refract(v3(v), v3(n), eta)); }
void orc_kat_texture_get(orc_scene* s, int surface, float u, float v, float out[4]) {
    V4 c = surf(s, surface)->get_f({u, v});
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}
void orc_kat_background(orc_scene* s, const float d[3], float out[3]) { put3(out, s->world.background->background({{0, 0, 0}, v3(d)})); }
int orc_kat_scatter(orc_scene* s, int material, uint64_t seed, const float ro[3], const float rd[3], const float p[3], const float n[3], int front,
                    float att[3], float dir[3]) {
    Rng rng(seed);
    tl_rng = &rng;
    Hit h;
    h.point = v3(p); h.normal = v3(n); h.front_face = front != 0; h.t = 1.0f; h.material = mat(s, material);
    Scatter sc;
    bool ok = h.material->scatter({v3(ro), v3(rd)}, h, sc);
    tl_rng = nullptr;
    if (ok) { put3(att, sc.attenuation); put3(dir, sc.scattered.direction); }
    return ok ? 1 : 0;
}
void orc_kat_samplers(uint64_t seed, uint64_t n, float* in_sphere, float* unit_vec, float* in_disk) {
    Rng rng(seed);
    tl_rng = &rng;
    for (uint64_t i = 0; i < n; ++i) {
        if (in_sphere) put3(in_sphere + 3 * i, random_in_unit_sphere());
        if (unit_vec) put3(unit_vec + 3 * i, random_unit_vector());
        if (in_disk) { V3 d = random_in_unit_disk(); in_disk[2 * i] = d.x; in_disk[2 * i + 1] = d.y; }
    }
    tl_rng = nullptr;
}

}  // extern "C"

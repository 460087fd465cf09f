// mrt_host.cpp — host-side scene API + flatten() above the C-ABI boundary (see include/mrt_host.h).
//
// Mirrors the reference's scene-construction code so that a scene written against World/Sphere/Model/
// Instance/Volume/Camera/PlyLoader produces the flat arrays mrt_scene_upload() takes. Scene data is kept
// flat from the start (no object graph): BLAS nodes are appended when a mesh is created (Model::new builds
// its BvhNode immediately, geom.rs:281-292), TLAS nodes when build_bvh() runs (world.rs:117-122).
// Arithmetic that ends up in the uploaded scene (triangle normals, instance matrices and bounds, camera
// frame) follows the reference's f32 operation order; build with -ffp-contract=off.
#include "mrt_host.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include <cerrno>
#include <sys/stat.h>

namespace {

const float kPi = 3.14159265358979323846f;
const float kInf = std::numeric_limits<float>::infinity();

// fastrand 1.4.1 (WyRand) — Cargo.lock:579-582; used at math.rs:245, geom.rs:111, main.rs:86. Source is not in the
// reference tree; restated from the published algorithm (no reference output pins it).
struct WyRand {
    uint64_t state = 0;
    uint64_t next() {
        state += 0xA0761D6478BD642FULL;
        unsigned __int128 t = (unsigned __int128)state * (state ^ 0xE7037ED1A0B428DBULL);
        return (uint64_t)t ^ (uint64_t)(t >> 64);
    }
    uint32_t next32() { return (uint32_t)next(); }
    float f32() {
        uint32_t bits = 0x3F800000u + (next32() >> 9);
        float f;
        std::memcpy(&f, &bits, sizeof f);
        return f - 1.0f;
    }
    uint32_t below(uint32_t n) {
        uint32_t r = next32();
        uint64_t m = (uint64_t)r * n;
        uint32_t lo = (uint32_t)m;
        if (lo < n) {
            uint32_t thresh = (0u - n) % n;
            while (lo < thresh) {
                r = next32();
                m = (uint64_t)r * n;
                lo = (uint32_t)m;
            }
        }
        return (uint32_t)(m >> 32);
    }
};

struct Vec3 {
    float x, y, z;
};
inline Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 scale(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot3(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross3(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Vec3 normalize(Vec3 a) {  // V3::unit math.rs:75-77: component-wise divide by the length
    float len = std::sqrt(dot3(a, a));
    return {a.x / len, a.y / len, a.z / len};
}
inline Vec3 load3(const float* p) { return {p[0], p[1], p[2]}; }
inline void store3(float* p, Vec3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// column-major 4x4: element(row r, column c) = m[4*c + r]   (math/generic.rs:71-77)
struct Mat4 {
    float m[16];
};
inline Mat4 mat_identity() { return {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}}; }
inline Mat4 mat_translation(Vec3 t) {  // math.rs:174-181
    Mat4 r = mat_identity();
    r.m[12] = t.x; r.m[13] = t.y; r.m[14] = t.z;
    return r;
}
inline Mat4 mat_scale(Vec3 s) {  // math.rs:217-224
    Mat4 r = mat_identity();
    r.m[0] = s.x; r.m[5] = s.y; r.m[10] = s.z;
    return r;
}
// math.rs:183-215: angle in turns; sine placement follows the reference literally
inline Mat4 mat_rotate(int axis, float turns) {
    float ang = turns * kPi * 2.0f;
    float s = std::sin(ang), c = std::cos(ang);
    Mat4 r = mat_identity();
    if (axis == 0) { r.m[5] = c; r.m[6] = s; r.m[9] = -s; r.m[10] = c; }          // c1=(0,c,s,0) c2=(0,-s,c,0)
    else if (axis == 1) { r.m[0] = c; r.m[2] = s; r.m[8] = -s; r.m[10] = c; }    // c0=(c,0,s,0) c2=(-s,0,c,0)
    else { r.m[0] = c; r.m[1] = -s; r.m[4] = s; r.m[5] = c; }                    // c0=(c,-s,0,0) c1=(s,c,0,0)
    return r;
}
inline Mat4 mat_mul(const Mat4& a, const Mat4& b) {  // generic.rs:126-161: out.c_j[i] = row_i(a) . b.c_j, summed left to right
    Mat4 r;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r.m[4 * j + i] = a.m[i] * b.m[4 * j] + a.m[4 + i] * b.m[4 * j + 1] + a.m[8 + i] * b.m[4 * j + 2] + a.m[12 + i] * b.m[4 * j + 3];
    return r;
}
inline Vec3 mat_apply(const Mat4& a, Vec3 v, float w) {  // generic.rs:105-115
    float o[3];
    for (int i = 0; i < 3; ++i) o[i] = ((a.m[i] * v.x + a.m[4 + i] * v.y) + a.m[8 + i] * v.z) + a.m[12 + i] * w;
    return {o[0], o[1], o[2]};
}

struct BuildItem {
    uint32_t ref;
    float lo[3], hi[3];
};

// Allocator whose resize() leaves new elements uninitialised, so the threads that fill a big mesh are also the ones that
// touch its pages first (a value-initialising resize would fault all of them in on one thread).
template <class T>
struct RawAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = RawAlloc<U>; };
    template <class U, class... A>
    void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U;
        else ::new ((void*)p) U(std::forward<A>(a)...);
    }
};

// fn(first, last) over [0, n) on up to 16 threads; small ranges stay on the caller's thread
template <class F>
void host_parallel_for(size_t n, F&& fn) {
    const size_t kGrain = 1u << 15;
    size_t threads = std::min<size_t>({(size_t)std::max(1u, std::thread::hardware_concurrency()), 16, n / kGrain});
    if (threads <= 1) { fn((size_t)0, n); return; }
    std::vector<std::thread> pool;
    const size_t chunk = (n + threads - 1) / threads;
    for (size_t t = 1; t < threads; ++t) pool.emplace_back([=, &fn] { fn(std::min(n, t * chunk), std::min(n, (t + 1) * chunk)); });
    fn((size_t)0, std::min(n, chunk));
    for (auto& th : pool) th.join();
}

}  // namespace

struct mrth_scene {
    WyRand rng;
    std::string err;
    std::vector<mrt_surface> surfaces;
    std::vector<mrt_texture> textures;
    std::vector<float> texels;
    std::vector<mrt_material> materials;
    std::vector<mrt_node> nodes;
    std::vector<mrt_sphere> spheres;
    std::vector<float, RawAlloc<float>> tri_verts;
    std::vector<mrt_tri_shading, RawAlloc<mrt_tri_shading>> tri_shading;
    std::vector<mrt_blas> blas;
    std::vector<std::array<float, 6>> blas_box;  // Model bounding box (= root node box) per mesh, kept apart so a deferred mesh has one too
    bool defer_mesh_bvh = false;                 // mrth_defer_mesh_bvh
    std::vector<mrt_instance> instances;
    std::vector<mrt_volume> volumes;
    std::vector<uint32_t> objects;  // World.objects in add order (prim refs)
    std::vector<uint32_t> roots;    // what World::intersect loops over (objects, or the TLAS root after build_bvh)
    uint64_t tlas_nodes = 0;
    mrt_background bg{};
    mrt_camera cam{};
    mrt_scene_desc desc{};
};

namespace {

void prim_bounds(const mrth_scene& s, uint32_t ref, float lo[3], float hi[3]) {
    uint32_t idx = MRT_REF_INDEX(ref);
    switch (MRT_REF_KIND(ref)) {
        case MRT_PRIM_NODE:
            std::memcpy(lo, s.nodes[idx].bmin, 12);
            std::memcpy(hi, s.nodes[idx].bmax, 12);
            break;
        case MRT_PRIM_SPHERE: {  // geom.rs:95-100
            const mrt_sphere& sp = s.spheres[idx];
            float r = std::fabs(sp.radius);
            for (int k = 0; k < 3; ++k) { lo[k] = sp.center[k] - r; hi[k] = sp.center[k] + r; }
            break;
        }
        case MRT_PRIM_TRIANGLE: {  // geom.rs:579-584
            const float* v = &s.tri_verts[9 * (size_t)idx];
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::fmin(std::fmin(v[k], v[3 + k]), v[6 + k]);
                hi[k] = std::fmax(std::fmax(v[k], v[3 + k]), v[6 + k]);
            }
            break;
        }
        case MRT_PRIM_INSTANCE:
            std::memcpy(lo, s.instances[idx].bmin, 12);
            std::memcpy(hi, s.instances[idx].bmax, 12);
            break;
        case MRT_PRIM_VOLUME:  // geom.rs:657-659
            prim_bounds(s, s.volumes[idx].target, lo, hi);
            break;
    }
}

// BvhNode::new geom.rs:109-161. Axis drawn on entry (before the children are built); nodes numbered in that order.
uint32_t build_bvh(mrth_scene& s, std::vector<BuildItem>& items, size_t first, size_t last) {
    const size_t n = last - first;
    const int axis = (int)s.rng.below(3);  // fastrand::u8(0..3)
    const uint32_t me = (uint32_t)s.nodes.size();
    s.nodes.push_back(mrt_node{});
    uint32_t left, right = MRT_REF_NONE;
    float llo[3], lhi[3], rlo[3], rhi[3];
    if (n == 1) {  // :120-121
        left = items[first].ref;
        std::memcpy(llo, items[first].lo, 12);
        std::memcpy(lhi, items[first].hi, 12);
    } else if (n == 2) {  // :122-129: a = pop() is the LAST item, b the first
        const BuildItem& a = items[first + 1];
        const BuildItem& b = items[first];
        const bool a_first = a.lo[axis] < b.lo[axis];
        const BuildItem& l = a_first ? a : b;
        const BuildItem& r = a_first ? b : a;
        left = l.ref;
        right = r.ref;
        std::memcpy(llo, l.lo, 12); std::memcpy(lhi, l.hi, 12);
        std::memcpy(rlo, r.lo, 12); std::memcpy(rhi, r.hi, 12);
    } else {  // :130-143. Rust's sort_by sees only Less/Greater; a stable sort on the strict predicate is used (tie order unpinned)
        std::stable_sort(items.begin() + (ptrdiff_t)first, items.begin() + (ptrdiff_t)last,
                         [axis](const BuildItem& p, const BuildItem& q) { return p.lo[axis] < q.lo[axis]; });
        const size_t mid = first + n / 2;
        left = build_bvh(s, items, first, mid);
        right = build_bvh(s, items, mid, last);
        std::memcpy(llo, s.nodes[MRT_REF_INDEX(left)].bmin, 12); std::memcpy(lhi, s.nodes[MRT_REF_INDEX(left)].bmax, 12);
        std::memcpy(rlo, s.nodes[MRT_REF_INDEX(right)].bmin, 12); std::memcpy(rhi, s.nodes[MRT_REF_INDEX(right)].bmax, 12);
    }
    mrt_node& node = s.nodes[me];
    node.left = left;
    node.right = right;
    for (int k = 0; k < 3; ++k) {  // BoundingBox::join :249-254
        node.bmin[k] = (right == MRT_REF_NONE) ? llo[k] : std::fmin(llo[k], rlo[k]);
        node.bmax[k] = (right == MRT_REF_NONE) ? lhi[k] : std::fmax(lhi[k], rhi[k]);
    }
    return MRT_REF(MRT_PRIM_NODE, me);
}

int finish_mesh(mrth_scene* s, uint32_t first_tri, uint32_t n_tris) {
    if (n_tris == 0) { s->err = "mesh has no triangles"; return MRT_E_INVALID; }
    mrt_blas b;
    b.first_tri = first_tri;
    b.n_tris = n_tris;
    std::array<float, 6> box;
    if (s->defer_mesh_bvh) {
        // No reference-topology tree: the backend builds its own over [first_tri, first_tri + n_tris). The box is the
        // min / max over the vertices, which is what the joins up the reference tree (geom.rs:249-254) arrive at too.
        box = {kInf, kInf, kInf, -kInf, -kInf, -kInf};
        const float* v = &s->tri_verts[9 * (size_t)first_tri];
        std::vector<std::array<float, 6>> parts(17, box);
        std::atomic<int> slot{0};
        host_parallel_for(3 * (size_t)n_tris, [&](size_t a, size_t e) {
            std::array<float, 6> bb = {kInf, kInf, kInf, -kInf, -kInf, -kInf};
            for (size_t i = a; i < e; ++i)
                for (int k = 0; k < 3; ++k) {  // f32::min / max with a non-NaN accumulator: a NaN coordinate is skipped either way
                    const float x = v[3 * i + k];
                    bb[k] = x < bb[k] ? x : bb[k];
                    bb[3 + k] = x > bb[3 + k] ? x : bb[3 + k];
                }
            parts[(size_t)slot.fetch_add(1)] = bb;
        });
        for (const auto& bb : parts)
            for (int k = 0; k < 3; ++k) {
                box[k] = bb[k] < box[k] ? bb[k] : box[k];
                box[3 + k] = bb[3 + k] > box[3 + k] ? bb[3 + k] : box[3 + k];
            }
        b.root = MRT_REF_NONE;
        b.n_nodes = 0;
    } else {
        std::vector<BuildItem> items(n_tris);
        for (uint32_t i = 0; i < n_tris; ++i) {
            items[i].ref = MRT_REF(MRT_PRIM_TRIANGLE, first_tri + i);
            prim_bounds(*s, items[i].ref, items[i].lo, items[i].hi);
        }
        const size_t before = s->nodes.size();
        b.root = build_bvh(*s, items, 0, n_tris);
        b.n_nodes = (uint32_t)(s->nodes.size() - before);
        const mrt_node& root = s->nodes[MRT_REF_INDEX(b.root)];
        box = {root.bmin[0], root.bmin[1], root.bmin[2], root.bmax[0], root.bmax[1], root.bmax[2]};
    }
    s->blas.push_back(b);
    s->blas_box.push_back(box);
    return (int)s->blas.size() - 1;
}

// Triangle::new geom.rs:449-466
void flat_triangle_at(mrth_scene* s, size_t i, Vec3 a, Vec3 b, Vec3 c, int material) {
    float* v = &s->tri_verts[9 * i];
    store3(v, a); store3(v + 3, b); store3(v + 6, c);
    Vec3 n = normalize(cross3(sub(b, a), sub(c, a)));
    mrt_tri_shading sh{};
    store3(sh.normal, n); store3(sh.normal + 3, n); store3(sh.normal + 6, n);
    sh.material = material;
    sh.flags = 0;
    s->tri_shading[i] = sh;
}

int push_object(mrth_scene* s, uint32_t ref) {
    s->objects.push_back(ref);
    s->roots.push_back(ref);
    return (int)s->objects.size() - 1;
}

bool valid_material(mrth_scene* s, int m, bool allow_none) {
    if (m < 0) {
        if (allow_none && m == -1) return true;
        s->err = "invalid material handle";
        return false;
    }
    if ((size_t)m >= s->materials.size()) { s->err = "invalid material handle"; return false; }
    return true;
}
bool valid_surface(mrth_scene* s, int h) {
    if (h < 0 || (size_t)h >= s->surfaces.size()) { s->err = "invalid surface handle"; return false; }
    return true;
}
int push_material(mrth_scene* s, int kind, int surface, int left, int right, float p0, float p1, float p2, float p3) {
    mrt_material m{kind, surface, left, right, {p0, p1, p2, p3}};
    s->materials.push_back(m);
    return (int)s->materials.size() - 1;
}
int push_surface(mrth_scene* s, int kind, int a, int b, int mode, float r, float g, float bl, float al) {
    mrt_surface x{kind, a, b, mode, {r, g, bl, al}};
    s->surfaces.push_back(x);
    return (int)s->surfaces.size() - 1;
}

// ------------------------------- PLY (ply_loader.rs) ---------------------------------------------
enum Scalar { S_I8, S_U8, S_I16, S_U16, S_I32, S_U32, S_F32, S_F64, S_INVALID };
Scalar scalar_from_name(const std::string& n) {  // ply_loader.rs:172-190
    static const struct { const char* a; const char* b; Scalar s; } table[] = {
        {"char", "int8", S_I8}, {"uchar", "uint8", S_U8}, {"short", "int16", S_I16}, {"ushort", "uint16", S_U16},
        {"int", "int32", S_I32}, {"uint", "uint32", S_U32}, {"float", "float32", S_F32}, {"double", "float64", S_F64}};
    for (const auto& e : table)
        if (n == e.a || n == e.b) return e.s;
    return S_INVALID;
}
int scalar_bytes(Scalar s) {
    static const int sz[] = {1, 1, 2, 2, 4, 4, 4, 8};
    return sz[s];
}
struct PlyProperty {
    std::string name;
    bool list;
    Scalar value, count;
};
struct PlyElement {
    std::string name;
    size_t count;
    std::vector<PlyProperty> props;
};
struct PlyCursor {
    const unsigned char* p;
    const unsigned char* end;
    int format;  // 0 ascii, 1 LE, 2 BE
    bool token(std::string& out) {  // ascii word (ply_loader.rs:24-33): skip whitespace, stop at the whitespace after the word
        out.clear();
        while (p < end) {
            unsigned char c = *p++;
            bool ws = std::isspace(c) != 0;
            if (ws && !out.empty()) return true;
            if (!ws) out.push_back((char)c);
        }
        return false;  // read_exact() ran out of bytes
    }
    bool bytes(void* dst, int n) {
        if (end - p < n) return false;
        unsigned char* d = (unsigned char*)dst;
        if (format == 2) for (int i = 0; i < n; ++i) d[i] = p[n - 1 - i];
        else std::memcpy(d, p, (size_t)n);
        p += n;
        return true;
    }
    bool skip(Scalar k) {
        if (format == 0) { std::string w; return token(w); }
        int n = scalar_bytes(k);
        if (end - p < n) return false;
        p += n;
        return true;
    }
    bool number_f32(Scalar k, float& out) {  // read_f32 :69-110
        if (format == 0) {
            std::string w;
            if (!token(w)) return false;
            char* stop = nullptr;
            out = std::strtof(w.c_str(), &stop);
            return stop != w.c_str() && *stop == 0;
        }
        switch (k) {
            case S_I8: { int8_t v; if (!bytes(&v, 1)) return false; out = (float)v; return true; }
            case S_U8: { uint8_t v; if (!bytes(&v, 1)) return false; out = (float)v; return true; }
            case S_I16: { int16_t v; if (!bytes(&v, 2)) return false; out = (float)v; return true; }
            case S_U16: { uint16_t v; if (!bytes(&v, 2)) return false; out = (float)v; return true; }
            case S_I32: { int32_t v; if (!bytes(&v, 4)) return false; out = (float)v; return true; }
            case S_U32: { uint32_t v; if (!bytes(&v, 4)) return false; out = (float)v; return true; }
            case S_F32: return bytes(&out, 4);
            case S_F64: { double v; if (!bytes(&v, 8)) return false; out = (float)v; return true; }
            default: return false;
        }
    }
    bool number_index(Scalar k, size_t& out) {  // read_usize :14-67
        if (format == 0) {
            std::string w;
            if (!token(w)) return false;
            if (k == S_F32 || k == S_F64) {
                char* stop = nullptr;
                double d = std::strtod(w.c_str(), &stop);
                if (stop == w.c_str() || *stop != 0) return false;
                out = d > 0.0 ? (size_t)d : 0;
                return true;
            }
            size_t i = (!w.empty() && w[0] == '+') ? 1 : 0;
            if (i >= w.size()) return false;
            size_t v = 0;
            for (; i < w.size(); ++i) {
                if (w[i] < '0' || w[i] > '9') return false;  // usize parse rejects '-' and '.'
                v = v * 10 + (size_t)(w[i] - '0');
            }
            out = v;
            return true;
        }
        switch (k) {
            case S_I8: { int8_t v; if (!bytes(&v, 1)) return false; out = (size_t)(int64_t)v; return true; }
            case S_U8: { uint8_t v; if (!bytes(&v, 1)) return false; out = v; return true; }
            case S_I16: { int16_t v; if (!bytes(&v, 2)) return false; out = (size_t)(int64_t)v; return true; }
            case S_U16: { uint16_t v; if (!bytes(&v, 2)) return false; out = v; return true; }
            case S_I32: { int32_t v; if (!bytes(&v, 4)) return false; out = (size_t)(int64_t)v; return true; }
            case S_U32: { uint32_t v; if (!bytes(&v, 4)) return false; out = v; return true; }
            case S_F32: { float v; if (!bytes(&v, 4)) return false; out = v > 0.0f ? (size_t)v : 0; return true; }
            case S_F64: { double v; if (!bytes(&v, 8)) return false; out = v > 0.0 ? (size_t)v : 0; return true; }
            default: return false;
        }
    }
};
std::string strip(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) ++b;
    while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}
std::vector<std::string> fields(const std::string& s) {  // str::split(' '): empty pieces are kept
    std::vector<std::string> f;
    size_t at = 0;
    while (true) {
        size_t sp = s.find(' ', at);
        f.push_back(s.substr(at, sp == std::string::npos ? std::string::npos : sp - at));
        if (sp == std::string::npos) break;
        at = sp + 1;
    }
    return f;
}
bool header_line(PlyCursor& c, std::string& line) {
    line.clear();
    if (c.p >= c.end) return false;
    while (c.p < c.end) {
        char ch = (char)*c.p++;
        if (ch == '\n') break;
        line.push_back(ch);
    }
    return true;
}
// PlyLoader::load ply_loader.rs:273-430. Appends 9 floats per 3-index face to `out`; every other element/property is skipped.
bool ply_read(const char* path, const int perm[3], std::vector<float>& out, float* max_abs, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::vector<unsigned char> data;
    unsigned char chunk[1 << 16];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof chunk, f)) > 0) data.insert(data.end(), chunk, chunk + got);
    std::fclose(f);
    PlyCursor cur{data.data(), data.data() + data.size(), 0};
    std::string line;
    if (!header_line(cur, line) || strip(line) != "ply") { err = "ply magic number not found"; return false; }
    std::vector<PlyElement> elements;
    for (;;) {
        if (!header_line(cur, line)) { err = "unexpected end of ply header"; return false; }
        std::vector<std::string> w = fields(strip(line));
        const std::string& key = w[0];
        if (key == "end_header") break;
        if (key == "format") {
            std::string kind = w.size() > 1 ? w[1] : "", ver = w.size() > 2 ? w[2] : "";
            if (w.size() > 2 && ver == "1.0" && kind == "ascii") cur.format = 0;
            else if (w.size() > 2 && ver == "1.0" && kind == "binary_little_endian") cur.format = 1;
            else if (w.size() > 2 && ver == "1.0" && kind == "binary_big_endian") cur.format = 2;
            else { err = "ply unsupported format found: " + kind + " " + ver; return false; }
        } else if (key == "comment" || key.empty()) {
        } else if (key == "element") {
            bool ok = w.size() > 2 && !w[2].empty();
            size_t count = 0;
            if (ok) {
                size_t i = w[2][0] == '+' ? 1 : 0;
                ok = i < w[2].size();
                for (; ok && i < w[2].size(); ++i) {
                    if (w[2][i] < '0' || w[2][i] > '9') ok = false;
                    else count = count * 10 + (size_t)(w[2][i] - '0');
                }
            }
            if (!ok) { err = "ply invalid element: '" + line + "'"; return false; }
            elements.push_back({w[1], count, {}});
        } else if (key == "property") {
            if (w.size() < 2) continue;
            if (w[1] == "list") {
                Scalar ck = w.size() > 2 ? scalar_from_name(w[2]) : S_INVALID;
                Scalar vk = w.size() > 3 ? scalar_from_name(w[3]) : S_INVALID;
                if (w.size() < 5 || ck == S_INVALID || vk == S_INVALID) { err = "ply invalid property: '" + line + "'"; return false; }
                if (!elements.empty()) elements.back().props.push_back({w[4], true, vk, ck});
            } else {
                Scalar k = scalar_from_name(w[1]);
                if (w.size() < 3 || k == S_INVALID) { err = "ply invalid property: '" + line + "'"; return false; }
                if (!elements.empty()) elements.back().props.push_back({w[2], false, k, S_INVALID});
            }
        } else {
            std::fprintf(stderr, "unknown ply header found: '%s'\n", key.c_str());
        }
    }
    std::vector<float> verts;  // 3 floats per accepted vertex, already permuted by vertex_fn
    float biggest = 0.0f;
    for (const PlyElement& el : elements) {
        const bool vertex = el.name == "vertex", face = el.name == "face";
        for (size_t i = 0; i < el.count; ++i) {
            float xyz[3];
            bool have[3] = {false, false, false};
            for (const PlyProperty& pr : el.props) {
                if (!pr.list) {
                    int comp = (vertex && pr.name.size() == 1 && pr.name[0] >= 'x' && pr.name[0] <= 'z') ? pr.name[0] - 'x' : -1;
                    if (comp >= 0) {
                        if (!cur.number_f32(pr.value, xyz[comp])) { err = "ply read error"; return false; }
                        have[comp] = true;
                    } else if (!cur.skip(pr.value)) { err = "ply read error"; return false; }
                } else {
                    size_t n;
                    if (!cur.number_index(pr.count, n)) { err = "ply read error"; return false; }
                    if (face && n == 3) {
                        size_t idx[3];
                        for (int k = 0; k < 3; ++k)
                            if (!cur.number_index(pr.value, idx[k])) { err = "ply read error"; return false; }
                        for (int k = 0; k < 3; ++k) {
                            if (idx[k] >= verts.size() / 3) { err = "ply face index out of bounds"; return false; }  // Rust: index panic
                            out.insert(out.end(), verts.begin() + (ptrdiff_t)(3 * idx[k]), verts.begin() + (ptrdiff_t)(3 * idx[k] + 3));
                        }
                    } else {
                        for (size_t k = 0; k < n; ++k)
                            if (!cur.skip(pr.value)) { err = "ply read error"; return false; }
                    }
                }
            }
            if (vertex && have[0] && have[1] && have[2]) {
                for (int k = 0; k < 3; ++k) biggest = std::fmax(biggest, std::fabs(xyz[k]));  // scenes/lucy.rs:37
                verts.push_back(xyz[perm[0]]);
                verts.push_back(xyz[perm[1]]);
                verts.push_back(xyz[perm[2]]);
            }
        }
    }
    if (max_abs) *max_abs = biggest;
    return true;
}

}  // namespace

namespace {

// Texture::load_bytes texture.rs:70-97: RGBA8 -> RGBA f32 texels, byte / 255, no sRGB decode
int surface_texture_rgba8(mrth_scene* s, const uint8_t* rgba, uint32_t w, uint32_t h, int wrap) {
    if (!rgba || w == 0 || h == 0) { s->err = "empty texture"; return MRT_E_INVALID; }
    if (wrap != MRT_WRAP_REPEAT && wrap != MRT_WRAP_CLAMP) { s->err = "Mirror wrapping is not implemented"; return MRT_E_UNSUPPORTED; }  // texture.rs:280
    mrt_texture t{w, h, wrap, 0, s->texels.size() / 4};
    size_t n = (size_t)w * h * 4;
    s->texels.reserve(s->texels.size() + n);
    for (size_t i = 0; i < n; ++i) s->texels.push_back((float)rgba[i] / 255.0f);  // texture.rs:79
    s->textures.push_back(t);
    return push_surface(s, MRT_SURF_TEXTURE, (int)s->textures.size() - 1, -1, 0, 0, 0, 0, 0);
}

// Model::new over Triangle::with_norms_and_uvs (geom.rs:468-496), one material per triangle
int mesh_from_uv_faces(mrth_scene* s, const float* v, const float* nn, const float* uv, const int* materials, uint64_t n) {
    if (n == 0 || n > 0x1FFFFFFFull - s->tri_shading.size()) { s->err = n ? "triangle count out of range" : "mesh has no triangles"; return MRT_E_INVALID; }
    uint32_t first = (uint32_t)s->tri_shading.size();
    for (uint64_t i = 0; i < n; ++i) {
        const float* p = v + 9 * i;
        const float* t = uv + 6 * i;
        s->tri_verts.insert(s->tri_verts.end(), p, p + 9);
        mrt_tri_shading sh{};
        std::memcpy(sh.normal, nn + 9 * i, 36);
        std::memcpy(sh.uv, t, 24);
        Vec3 ab = sub(load3(p + 3), load3(p)), ac = sub(load3(p + 6), load3(p));
        float abu = t[2] - t[0], abv = t[3] - t[1], acu = t[4] - t[0], acv = t[5] - t[1];
        float r = std::fmax(std::fmin(1.0f / (abu * acv - abv * acu), 1.0f), -1.0f);
        store3(sh.tangent, scale(sub(scale(ab, acv), scale(ac, abv)), r));
        store3(sh.bitangent, scale(sub(scale(ac, abu), scale(ab, acu)), r));
        sh.material = materials[i];
        sh.flags = MRT_TRI_HAS_UV;
        s->tri_shading.push_back(sh);
    }
    return finish_mesh(s, first, (uint32_t)n);
}

#include "mrt_obj.inc"
#include "mrt_png_write.inc"

}  // namespace

extern "C" {

mrth_scene* mrth_scene_new(void) {
    mrth_scene* s = new mrth_scene();
    s->bg.kind = MRT_BG_SOLID;
    return s;
}
void mrth_scene_free(mrth_scene* s) { delete s; }
const char* mrth_last_error(mrth_scene* s) { return s->err.c_str(); }
void mrth_seed(mrth_scene* s, uint64_t seed) { s->rng.state = seed; }
void mrth_defer_mesh_bvh(mrth_scene* s, int on) { s->defer_mesh_bvh = on != 0; }
float mrth_rand_f32(mrth_scene* s) { return s->rng.f32(); }

int mrth_surface_solid(mrth_scene* s, float r, float g, float b, float a) { return push_surface(s, MRT_SURF_SOLID, -1, -1, 0, r, g, b, a); }
int mrth_surface_texture(mrth_scene* s, const uint8_t* rgba, uint32_t w, uint32_t h, int wrap) { return surface_texture_rgba8(s, rgba, w, h, wrap); }
int mrth_surface_texture_png(mrth_scene* s, const char* path, int wrap) {  // Texture::load_png texture.rs:29-68
    std::vector<uint8_t> rgba;
    uint32_t w = 0, h = 0;
    if (!png_decode_rgba8(path, rgba, w, h, s->err)) return MRT_E_INVALID;
    return surface_texture_rgba8(s, rgba.data(), w, h, wrap);
}
int mrth_surface_ycbcr(mrth_scene* s, int luma, int chroma) {
    if (!valid_surface(s, luma) || !valid_surface(s, chroma)) return MRT_E_INVALID;
    if (s->surfaces[(size_t)luma].kind != MRT_SURF_TEXTURE || s->surfaces[(size_t)chroma].kind != MRT_SURF_TEXTURE) { s->err = "ycbcr needs two Texture surfaces"; return MRT_E_INVALID; }
    return push_surface(s, MRT_SURF_YCBCR, s->surfaces[(size_t)luma].a, s->surfaces[(size_t)chroma].a, 0, 0, 0, 0, 0);
}
int mrth_surface_blend(mrth_scene* s, int mode, int left, int right) {
    if (!valid_surface(s, left) || !valid_surface(s, right) || mode < 0 || mode > 3) return MRT_E_INVALID;
    return push_surface(s, MRT_SURF_BLEND, left, right, mode, 0, 0, 0, 0);
}
int mrth_surface_fallback(mrth_scene* s, float r, float g, float b, float a, int inner) {
    if (!valid_surface(s, inner)) return MRT_E_INVALID;
    return push_surface(s, MRT_SURF_FALLBACK, inner, -1, 0, r, g, b, a);
}

int mrth_mat_absorb(mrth_scene* s) { return push_material(s, MRT_MAT_ABSORB, -1, -1, -1, 0, 0, 0, 0); }
int mrth_mat_lambertian(mrth_scene* s, int surface) {
    if (!valid_surface(s, surface)) return MRT_E_INVALID;
    return push_material(s, MRT_MAT_LAMBERTIAN, surface, -1, -1, 0, 0, 0, 0);
}
int mrth_mat_diffuse_light(mrth_scene* s, float r, float g, float b) { return push_material(s, MRT_MAT_DIFFUSE_LIGHT, -1, -1, -1, r, g, b, 0); }
int mrth_mat_metal(mrth_scene* s, float fuzz, int surface) {
    if (!valid_surface(s, surface)) return MRT_E_INVALID;
    return push_material(s, MRT_MAT_METAL, surface, -1, -1, fuzz < 1.0f ? fuzz : 1.0f, 0, 0, 0);  // Metal::new material.rs:255-258
}
int mrth_mat_dielectric(mrth_scene* s, float ior) { return push_material(s, MRT_MAT_DIELECTRIC, -1, -1, -1, ior, 0, 0, 0); }
int mrth_mat_specular(mrth_scene* s, float ior, int surface) {
    if (!valid_surface(s, surface)) return MRT_E_INVALID;
    return push_material(s, MRT_MAT_SPECULAR, surface, -1, -1, ior, 0, 0, 0);
}
int mrth_mat_mix(mrth_scene* s, float ratio, int left, int right) {
    if (!valid_material(s, left, false) || !valid_material(s, right, false)) return MRT_E_INVALID;
    return push_material(s, MRT_MAT_MIX, -1, left, right, ratio, 0, 0, 0);
}
int mrth_mat_isotropic(mrth_scene* s, float r, float g, float b) { return push_material(s, MRT_MAT_ISOTROPIC, -1, -1, -1, r, g, b, 0); }
// EveMaterial::new + EveMaterialColor (eve.rs:43-64, :135-199): three texture surfaces and the palette, flattened as mrt.h describes
int mrth_mat_eve(mrth_scene* s, int normal_occlusion, int albedo_roughness, int pmdg, const float colors12[12], const float glow3[3]) {
    for (int h : {normal_occlusion, albedo_roughness, pmdg})
        if (h < 0 || (size_t)h >= s->surfaces.size()) { s->err = "invalid surface handle"; return MRT_E_INVALID; }
    const int palette = (int)s->surfaces.size();
    for (int i = 0; i < 4; ++i) mrth_surface_solid(s, colors12[3 * i], colors12[3 * i + 1], colors12[3 * i + 2], i < 3 ? glow3[i] : 0.0f);
    float bits;
    std::memcpy(&bits, &palette, 4);
    return push_material(s, MRT_MAT_EVE, normal_occlusion, albedo_roughness, pmdg, bits, 0, 0, 0);
}

void mrth_background_solid(mrth_scene* s, float r, float g, float b) {
    s->bg = mrt_background{};
    s->bg.kind = MRT_BG_SOLID;
    s->bg.color[0] = r; s->bg.color[1] = g; s->bg.color[2] = b;
}
void mrth_background_sky(mrth_scene* s) { s->bg = mrt_background{}; s->bg.kind = MRT_BG_SKY; }
void mrth_background_skysphere(mrth_scene* s, int surface) { s->bg = mrt_background{}; s->bg.kind = MRT_BG_SKYSPHERE; s->bg.surface[0] = surface; }
void mrth_background_cubemap(mrth_scene* s, const int f[6], float rx, float ry, float rz) {
    s->bg = mrt_background{};
    s->bg.kind = MRT_BG_CUBEMAP;
    for (int i = 0; i < 6; ++i) s->bg.surface[i] = f[i];
    Mat4 t = mat_mul(mat_mul(mat_rotate(0, rx), mat_rotate(0, ry)), mat_rotate(0, rz));  // rotate_x thrice: material.rs:103-105
    std::memcpy(s->bg.transform, t.m, sizeof t.m);
}

int mrth_mesh_new(mrth_scene* s, const float* v, uint64_t n, int tri_material) {
    if (!valid_material(s, tri_material, false)) return MRT_E_INVALID;
    if (n == 0 || n > 0x1FFFFFFFull - s->tri_shading.size()) { s->err = "triangle count out of range"; return MRT_E_INVALID; }
    uint32_t first = (uint32_t)s->tri_shading.size();
    s->tri_verts.resize(s->tri_verts.size() + 9 * n);  // uninitialised (RawAlloc): every element is written below
    s->tri_shading.resize(s->tri_shading.size() + n);
    host_parallel_for((size_t)n, [&](size_t a, size_t e) {
        for (size_t i = a; i < e; ++i) flat_triangle_at(s, first + i, load3(v + 9 * i), load3(v + 9 * i + 3), load3(v + 9 * i + 6), tri_material);
    });
    return finish_mesh(s, first, (uint32_t)n);
}
int mrth_mesh_new_uv(mrth_scene* s, const float* v, const float* nn, const float* uv, uint64_t n, int tri_material) {
    if (!valid_material(s, tri_material, false)) return MRT_E_INVALID;
    std::vector<int> mats((size_t)n, tri_material);
    return mesh_from_uv_faces(s, v, nn, uv, mats.data(), n);
}
int mrth_mesh_load_obj(mrth_scene* s, const char* path, int wrap, const char* filtered_groups) {
    ObjOptions opt{true, wrap, -1, {}};
    if (filtered_groups) {  // newline-separated group names (SimpleTexturedBuilder::with_filter obj_loader.rs:174-186)
        std::string all(filtered_groups), cur;
        for (char c : all) {
            if (c == '\n') { if (!cur.empty()) opt.filtered_groups.push_back(cur); cur.clear(); }
            else cur.push_back(c);
        }
        if (!cur.empty()) opt.filtered_groups.push_back(cur);
    }
    return load_obj(s, path, opt);
}
int mrth_mesh_load_obj_with(mrth_scene* s, const char* path, int tri_material) {
    if (!valid_material(s, tri_material, false)) return MRT_E_INVALID;
    return load_obj(s, path, ObjOptions{false, MRT_WRAP_REPEAT, tri_material, {}});
}
// a bad handle sets the error text and returns (nothing unwinds across the C boundary)
static const mrt_blas* mesh_or_null(mrth_scene* s, int mesh) {
    if (mesh < 0 || (size_t)mesh >= s->blas.size()) { s->err = "mesh handle out of range"; return nullptr; }
    return &s->blas[(size_t)mesh];
}
static bool object_ok(mrth_scene* s, int object) {
    if (object < 0 || (size_t)object >= s->objects.size()) { s->err = "object index out of range"; return false; }
    return true;
}
void mrth_mesh_get_shading(mrth_scene* s, int mesh, float* normals9, float* uvs6, int32_t* materials) {
    const mrt_blas* bp = mesh_or_null(s, mesh);
    if (!bp) return;
    const mrt_blas& b = *bp;
    for (uint32_t i = 0; i < b.n_tris; ++i) {
        const mrt_tri_shading& sh = s->tri_shading[(size_t)b.first_tri + i];
        if (normals9) std::memcpy(normals9 + 9 * (size_t)i, sh.normal, 36);
        if (uvs6) std::memcpy(uvs6 + 6 * (size_t)i, sh.uv, 24);
        if (materials) materials[i] = sh.material;
    }
}
int mrth_material_info(mrth_scene* s, int material, float color4[4], uint32_t wh[2], uint64_t* texel_hash) {
    if (!valid_material(s, material, false)) return MRT_E_INVALID;
    const mrt_material& m = s->materials[(size_t)material];
    color4[0] = color4[1] = color4[2] = color4[3] = 0.0f;
    wh[0] = wh[1] = 0;
    *texel_hash = 0;
    if (m.surface >= 0) {
        const mrt_surface& u = s->surfaces[(size_t)m.surface];
        if (u.kind == MRT_SURF_SOLID) std::memcpy(color4, u.color, 16);
        if (u.kind == MRT_SURF_TEXTURE) {
            const mrt_texture& t = s->textures[(size_t)u.a];
            wh[0] = t.width; wh[1] = t.height;
            uint64_t h = 1469598103934665603ull;  // FNV-1a over the f32 texels
            const unsigned char* p = reinterpret_cast<const unsigned char*>(&s->texels[4 * (size_t)t.texel_offset]);
            for (size_t i = 0; i < (size_t)t.width * t.height * 16; ++i) { h ^= p[i]; h *= 1099511628211ull; }
            *texel_hash = h;
        }
    }
    return m.kind;
}
int mrth_mesh_load_ply(mrth_scene* s, const char* path, const int perm[3], int tri_material, float* max_abs) {
    if (!valid_material(s, tri_material, false)) return MRT_E_INVALID;
    std::vector<float> v;
    if (!ply_read(path, perm, v, max_abs, s->err)) return MRT_E_INVALID;
    return mrth_mesh_new(s, v.data(), v.size() / 9, tri_material);
}
int mrth_mesh_load_stl(mrth_scene* s, const char* path, const int perm[3], int tri_material) {
    // StlLoader::load_binary stl_loader.rs:10-66: header[80], u32 count, then per triangle {normal (ignored), a, b, c, u16 n, n attribute bytes}
    if (!valid_material(s, tri_material, false)) return MRT_E_INVALID;
    FILE* f = std::fopen(path, "rb");
    if (!f) { s->err = std::string("cannot open ") + path; return MRT_E_INVALID; }
    std::vector<float> v;
    unsigned char header[80];
    uint32_t count = 0;
    bool ok = std::fread(header, 1, 80, f) == 80 && std::fread(&count, 4, 1, f) == 1;
    for (uint32_t i = 0; ok && i < count; ++i) {
        float rec[12];
        uint16_t attr = 0;
        ok = std::fread(rec, 4, 12, f) == 12 && std::fread(&attr, 2, 1, f) == 1;
        if (ok && attr) ok = std::fseek(f, attr, SEEK_CUR) == 0;
        if (ok)
            for (int k = 0; k < 3; ++k)
                for (int c = 0; c < 3; ++c) v.push_back(rec[3 + 3 * k + perm[c]]);
    }
    if (ok && count) {  // fseek past the end succeeds silently: make sure the last record was really there
        long pos = std::ftell(f);
        std::fseek(f, 0, SEEK_END);
        ok = pos <= std::ftell(f);
    }
    std::fclose(f);
    if (!ok) { s->err = "stl read error"; return MRT_E_INVALID; }
    return mrth_mesh_new(s, v.data(), v.size() / 9, tri_material);
}
uint64_t mrth_mesh_tri_count(mrth_scene* s, int mesh) {
    const mrt_blas* b = mesh_or_null(s, mesh);
    return b ? b->n_tris : 0;
}
void mrth_mesh_get_verts(mrth_scene* s, int mesh, float* out) {
    const mrt_blas* bp = mesh_or_null(s, mesh);
    if (!bp) return;
    const mrt_blas& b = *bp;
    std::memcpy(out, &s->tri_verts[9 * (size_t)b.first_tri], 36 * (size_t)b.n_tris);
}
uint64_t mrth_mesh_node_count(mrth_scene* s, int mesh) {
    const mrt_blas* b = mesh_or_null(s, mesh);
    return b ? b->n_nodes : 0;
}

int mrth_add_sphere(mrth_scene* s, int material, float cx, float cy, float cz, float radius) {
    if (!valid_material(s, material, false)) return MRT_E_INVALID;
    mrt_sphere sp{{cx, cy, cz}, radius, material, (uint32_t)s->objects.size(), {0, 0}};
    s->spheres.push_back(sp);
    return push_object(s, MRT_REF(MRT_PRIM_SPHERE, s->spheres.size() - 1));
}
// as_object: the instance is an object of the world (World::add); otherwise it is the target of a Volume and only the Volume is
// (returns the instance index then)
static int add_instance_common(mrth_scene* s, int mesh, const Mat4& fwd, const Mat4& inv, uint32_t flags, int material, bool as_object = true) {
    if (mesh < 0 || (size_t)mesh >= s->blas.size()) { s->err = "invalid mesh handle"; return MRT_E_INVALID; }
    if (!valid_material(s, material, true)) return MRT_E_INVALID;
    mrt_instance in{};
    std::memcpy(in.transform, fwd.m, 64);
    std::memcpy(in.inv_transform, inv.m, 64);
    const float* root_min = s->blas_box[(size_t)mesh].data();
    const float* root_max = root_min + 3;
    if (flags & MRT_INSTANCE_IDENTITY) {  // Model::bounding_box geom.rs:330-332
        std::memcpy(in.bmin, root_min, 12);
        std::memcpy(in.bmax, root_max, 12);
    } else {  // geom.rs:369-381 with corners() :256-272
        float lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
        for (int c = 0; c < 8; ++c) {
            Vec3 corner{(c & 1) ? root_min[0] : root_max[0], (c & 2) ? root_min[1] : root_max[1], (c & 4) ? root_min[2] : root_max[2]};
            Vec3 p = mat_apply(fwd, corner, 1.0f);
            lo[0] = std::fmin(lo[0], p.x); lo[1] = std::fmin(lo[1], p.y); lo[2] = std::fmin(lo[2], p.z);
            hi[0] = std::fmax(hi[0], p.x); hi[1] = std::fmax(hi[1], p.y); hi[2] = std::fmax(hi[2], p.z);
        }
        std::memcpy(in.bmin, lo, 12);
        std::memcpy(in.bmax, hi, 12);
    }
    in.blas = (uint32_t)mesh;
    in.material = material;
    in.flags = flags;
    in.object_id = as_object ? (uint32_t)s->objects.size() : MRT_REF_NONE;
    s->instances.push_back(in);
    if (!as_object) return (int)s->instances.size() - 1;
    return push_object(s, MRT_REF(MRT_PRIM_INSTANCE, s->instances.size() - 1));
}
int mrth_add_model(mrth_scene* s, int mesh, int override_material) {
    return add_instance_common(s, mesh, mat_identity(), mat_identity(), MRT_INSTANCE_IDENTITY, override_material);
}
static void instance_matrices(const float t[3], const float r[3], const float sc[3], Mat4& fwd, Mat4& inv) {
    // Instance::new geom.rs:343-367
    Vec3 tr = load3(t), rot = load3(r), scl = load3(sc);
    Vec3 inv_tr = scale(tr, -1.0f), inv_rot = scale(rot, -1.0f);
    Vec3 inv_scl{1.0f / scl.x, 1.0f / scl.y, 1.0f / scl.z};
    Mat4 rotation = mat_mul(mat_mul(mat_rotate(0, rot.x), mat_rotate(1, rot.y)), mat_rotate(2, rot.z));
    Mat4 inv_rotation = mat_mul(mat_mul(mat_rotate(2, inv_rot.z), mat_rotate(1, inv_rot.y)), mat_rotate(0, inv_rot.x));
    fwd = mat_mul(mat_mul(mat_translation(tr), rotation), mat_scale(scl));
    inv = mat_mul(mat_mul(mat_scale(inv_scl), inv_rotation), mat_translation(inv_tr));
}
int mrth_add_instance(mrth_scene* s, int mesh, const float t[3], const float r[3], const float sc[3], int override_material) {
    Mat4 fwd, inv;
    instance_matrices(t, r, sc, fwd, inv);
    return add_instance_common(s, mesh, fwd, inv, 0, override_material);
}
// Volume::new(target, density, albedo) geom.rs:601-609 over a Model or an Instance of a mesh
static int add_volume_over(mrth_scene* s, int target_instance, float density, float r, float g, float b) {
    if (target_instance < 0) return target_instance;
    mrt_volume v{MRT_REF(MRT_PRIM_INSTANCE, (uint32_t)target_instance), -1.0f / density, mrth_mat_isotropic(s, r, g, b), (uint32_t)s->objects.size()};
    s->volumes.push_back(v);
    return push_object(s, MRT_REF(MRT_PRIM_VOLUME, s->volumes.size() - 1));
}
int mrth_add_volume_model(mrth_scene* s, int mesh, float density, float r, float g, float b) {
    return add_volume_over(s, add_instance_common(s, mesh, mat_identity(), mat_identity(), MRT_INSTANCE_IDENTITY, -1, false), density, r, g, b);
}
int mrth_add_volume_instance(mrth_scene* s, int mesh, const float t[3], const float rot[3], const float sc[3], float density, float r, float g, float b) {
    Mat4 fwd, inv;
    instance_matrices(t, rot, sc, fwd, inv);
    return add_volume_over(s, add_instance_common(s, mesh, fwd, inv, 0, -1, false), density, r, g, b);
}
int mrth_add_volume_sphere(mrth_scene* s, float cx, float cy, float cz, float radius, float density, float r, float g, float b) {
    int absorb = mrth_mat_absorb(s);  // Sphere<()> target (scenes/eve.rs:41-45)
    mrt_sphere sp{{cx, cy, cz}, radius, absorb, MRT_REF_NONE, {0, 0}};
    s->spheres.push_back(sp);
    mrt_volume v{MRT_REF(MRT_PRIM_SPHERE, s->spheres.size() - 1), -1.0f / density, mrth_mat_isotropic(s, r, g, b), (uint32_t)s->objects.size()};  // geom.rs:603-609
    s->volumes.push_back(v);
    return push_object(s, MRT_REF(MRT_PRIM_VOLUME, s->volumes.size() - 1));
}
void mrth_build_bvh(mrth_scene* s) {  // World::build_bvh world.rs:117-122
    if (s->roots.empty()) return;
    std::vector<BuildItem> items(s->roots.size());
    for (size_t i = 0; i < items.size(); ++i) {
        items[i].ref = s->roots[i];
        prim_bounds(*s, items[i].ref, items[i].lo, items[i].hi);
    }
    const size_t before = s->nodes.size();
    uint32_t root = build_bvh(*s, items, 0, items.size());
    s->tlas_nodes = s->nodes.size() - before;
    s->roots.assign(1, root);
}
uint64_t mrth_tlas_node_count(mrth_scene* s) { return s->tlas_nodes; }

void mrth_camera(mrth_scene* s, float vfov, const float from[3], const float at[3], const float up[3], float aspect, float aperture, float focus) {
    // Camera::new world.rs:16-51
    float rads = vfov * kPi / 180.0f;
    float half_height = std::tan(rads / 2.0f);
    float viewport_height = half_height * 2.0f;
    float viewport_width = aspect * viewport_height;
    Vec3 origin = load3(from);
    Vec3 w = normalize(sub(origin, load3(at)));
    Vec3 u = normalize(cross3(load3(up), w));
    Vec3 v = cross3(w, u);
    Vec3 horizontal = scale(scale(u, viewport_width), focus);
    Vec3 vertical = scale(scale(v, viewport_height), focus);
    Vec3 hh{horizontal.x / 2.0f, horizontal.y / 2.0f, horizontal.z / 2.0f};
    Vec3 vh{vertical.x / 2.0f, vertical.y / 2.0f, vertical.z / 2.0f};
    Vec3 llc = sub(sub(sub(origin, hh), vh), scale(w, focus));
    store3(s->cam.origin, origin);
    store3(s->cam.lower_left_corner, llc);
    store3(s->cam.horizontal, horizontal);
    store3(s->cam.vertical, vertical);
    store3(s->cam.u, u);
    store3(s->cam.v, v);
    s->cam.lens_radius = aperture / 2.0f;
}

void mrth_get_camera(mrth_scene* s, float o[19]) {
    std::memcpy(o, s->cam.origin, 12); std::memcpy(o + 3, s->cam.lower_left_corner, 12); std::memcpy(o + 6, s->cam.horizontal, 12);
    std::memcpy(o + 9, s->cam.vertical, 12); std::memcpy(o + 12, s->cam.u, 12); std::memcpy(o + 15, s->cam.v, 12);
    o[18] = s->cam.lens_radius;
}
void mrth_get_instance(mrth_scene* s, int object, float tf[16], float inv[16], float aabb[6]) {
    if (!object_ok(s, object)) return;
    uint32_t ref = s->objects[(size_t)object];
    if (MRT_REF_KIND(ref) != MRT_PRIM_INSTANCE) { s->err = "object is not an Instance"; return; }
    const mrt_instance& in = s->instances[MRT_REF_INDEX(ref)];
    std::memcpy(tf, in.transform, 64);
    std::memcpy(inv, in.inv_transform, 64);
    std::memcpy(aabb, in.bmin, 12);
    std::memcpy(aabb + 3, in.bmax, 12);
}
void mrth_get_object_aabb(mrth_scene* s, int object, float aabb[6]) {
    if (!object_ok(s, object)) return;
    prim_bounds(*s, s->objects[(size_t)object], aabb, aabb + 3);
}

// Image::to_rgb_bytes(Albedo | Normal) over a FloatBuffer (main.rs:694-721, final cast :718-721), rows reversed like dump (:763-768) when flip != 0
int mrth_float_buffer_rgb8(const float* buf, uint32_t w, uint32_t h, int mode, int flip, uint8_t* out) {
    if (!buf || !out || w == 0 || h == 0) return MRT_E_INVALID;
    if (mode != MRT_DISPLAY_ALBEDO && mode != MRT_DISPLAY_NORMAL) return MRT_E_UNSUPPORTED;
    const size_t stride = (size_t)w * 3;
    for (uint32_t y = 0; y < h; ++y) {
        const float* src = buf + stride * y;
        uint8_t* dst = out + stride * (flip ? h - 1 - y : y);
        for (size_t i = 0; i < stride; ++i) {
            const float p = src[i];
            // f32::min / max return the other operand for a NaN: a NaN albedo becomes 1.0, a NaN normal stays NaN and casts to 0
            const float f = mode == MRT_DISPLAY_ALBEDO ? std::pow(std::fmax(std::fmin(p, 1.0f), 0.0f), 1.0f / 2.2f) : (p + 1.0f) / 2.0f;
            dst[i] = saturate_u8(f * 255.0f);
        }
    }
    return MRT_OK;
}

// Image::dump's file (main.rs:770-783): parent directories created, 8-bit RGB PNG, rows as given (resolve with flip != 0 first)
int mrth_write_png(const char* path, const uint8_t* rgb, uint32_t w, uint32_t h) {
    if (!path || !rgb || w == 0 || h == 0 || (uint64_t)w * h > (1ull << 28)) return MRT_E_INVALID;
    std::vector<uint8_t> file;
    png_encode_rgb8(rgb, w, h, file);
    FILE* f = create_parent_dirs(path) ? std::fopen(path, "wb") : nullptr;
    bool ok = f && std::fwrite(file.data(), 1, file.size(), f) == file.size();
    if (f) ok = std::fclose(f) == 0 && ok;
    if (!ok) {
        std::fprintf(stderr, "Unable to save image: %s: %s\n", path, std::strerror(errno));  // main.rs:780-782: reported, not fatal
        return MRT_E_INVALID;
    }
    return MRT_OK;
}

const mrt_scene_desc* mrth_scene_desc(mrth_scene* s) {
    mrt_scene_desc& d = s->desc;
    d = mrt_scene_desc{};
    d.abi_version = MRT_ABI_VERSION;
    d.flags = 0;
    d.roots = s->roots.data();
    d.n_roots = (uint32_t)s->roots.size();
    d.n_objects = (uint32_t)s->objects.size();
    d.nodes = s->nodes.data(); d.n_nodes = s->nodes.size();
    d.spheres = s->spheres.data(); d.n_spheres = s->spheres.size();
    d.tri_verts = s->tri_verts.data(); d.tri_shading = s->tri_shading.data(); d.n_tris = s->tri_shading.size();
    d.blas = s->blas.data(); d.n_blas = s->blas.size();
    d.instances = s->instances.data(); d.n_instances = s->instances.size();
    d.volumes = s->volumes.data(); d.n_volumes = s->volumes.size();
    d.materials = s->materials.data(); d.n_materials = s->materials.size();
    d.surfaces = s->surfaces.data(); d.n_surfaces = s->surfaces.size();
    d.textures = s->textures.data(); d.n_textures = s->textures.size();
    d.texels = s->texels.data(); d.n_texels = s->texels.size() / 4;
    d.background = s->bg;
    return &d;
}
const mrt_camera* mrth_scene_camera(mrth_scene* s) { return &s->cam; }

}  // extern "C"

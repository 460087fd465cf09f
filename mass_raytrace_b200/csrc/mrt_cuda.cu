// mrt_cuda.cu — wavefront path tracer for sm_100a behind the C ABI of include/mrt.h.
//
// Replaces render() (reference src/main.rs:150-295). One context drives one GPU:
//
//   advance  (1 thread)      queue bookkeeping, hands out the next block of pixel samples
//   generate (grid-stride)   Camera::ray for the new paths of this iteration          world.rs:53-63, main.rs:258-260
//   extend   (persistent)    closest hit through TLAS/BLAS, classify by material      world.rs:68, geom.rs
//   shade    (persistent)    emit + scatter per material-sorted queue, accumulate     world.rs:69-77, material.rs
//
// The queues carry the path state itself, not indices into a pool: an extend-queue entry is a 48-byte ray record (origin,
// direction, throughput + pixel / sample / bounce), a shade-queue entry the same plus the 16-byte hit. Every kernel therefore
// streams its input with coalesced 128-bit loads and appends its output with one atomic per warp (ballot / match + popc +
// shuffle); nothing is gathered through a permuted index. A path that ends frees its place, and the next iteration generates as
// many new camera rays as fit ("regeneration"), so every extend launch works on a full queue until the job drains.
// Radiance is accumulated as 64-bit fixed point (2^-32 units) with integer atomics: sums are exact and therefore
// independent of sample order, of how a sample range is split over calls, and of the number of GPUs.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mrt.h"
#include "mrt_bvh_build.h"
#include "mrt_debug.h"
#include "mrt_device.cuh"
#include "mrt_lbvh.cuh"

namespace mrt {

enum { Q_MISS = 0, Q_FIRST_MAT = 1, Q_COUNT = 1 + MRT_MAT_KINDS };

struct QueueState {
    uint32_t n_ext;        // rays traced this iteration = n_cont + n_new
    uint32_t n_cont;       // continuing paths (written by the previous shade into the current extend queue)
    uint32_t n_new;        // camera rays generated this iteration
    uint32_t n_next;       // appended by shade for the next iteration
    uint32_t ext_cursor;   // persistent-warp fetch cursor of extend
    uint32_t done;
    uint32_t finish_n;     // > 0: the drain has started; k_finish runs these last paths to completion in one launch
    uint32_t shade_done;   // blocks of k_shade that have finished this iteration (the last one runs advance())
    uint32_t n_shade[Q_COUNT + 3];
    unsigned long long next_work, total_work, gen_base;
    unsigned long long rays, iterations;
    VisitCounters visits;
};

struct RenderParams {
    uint32_t w, h, npix, spp_begin, max_depth;
    uint2 seed;
    uint32_t refill_lanes;   // k_extend commits and refills finished lanes once this many lanes of a warp are idle
    uint32_t node_burst;     // k_extend: node visits per lane between two looks at the leaves
    uint32_t finish_paths;   // drain threshold of k_finish (0 = never)
    uint32_t capacity;       // rays in flight (entries per queue)
    uint8_t region[Q_COUNT + 3];  // shade-queue region of each queue kind (only the material kinds the scene uses get one)
};

// One path in flight. Extend queue entry = RayRec (48 B); shade queue entry = HitEntry (64 B = two 32-byte sectors).
struct __align__(16) RayRec {
    float4 o;    // ray origin.xyz, pixel (as bits)
    float4 d;    // ray direction.xyz, sample index (as bits)
    float4 thr;  // throughput.rgb, bounce (as bits)
};
struct __align__(16) HitEntry {
    RayRec ray;
    uint4 hit;  // t (as bits), prim ref, instance, material
};
struct Pool {
    RayRec* q_ext[2];   // ping-pong: shade of iteration k appends to the queue extend reads in iteration k + 1
    HitEntry* q_shade;  // `regions` regions of `capacity` entries
    uint32_t capacity, regions;
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// one atomic per warp: every lane of the warp must call this (pred false = no entry). Returns the entry index.
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter, bool pred) {
    uint32_t mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return 0;
    uint32_t leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(mask & ((1u << lane_id()) - 1u));
}

__device__ __forceinline__ void accumulate(long long* accum, uint32_t* nonfinite, uint32_t pixel, V3 c) {
    const float kScale = 4294967296.0f;  // 2^32
    float v[3] = {c.x, c.y, c.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (v[k] == 0.0f) continue;
        if (!isfinite(v[k])) {
            atomicOr(&nonfinite[pixel], 1u << k);  // the reference's f32 sum would be poisoned (main.rs:634)
            continue;
        }
        atomicAdd(reinterpret_cast<unsigned long long*>(&accum[(size_t)pixel * 4 + k]), (unsigned long long)__float2ll_rn(v[k] * kScale));
    }
}

// Bookkeeping between two wavefront iterations (one thread): the paths shade appended become this iteration's continuing
// rays, the places of ended paths are refilled with the next pixel samples ("regeneration"). Runs in the last block of k_shade
// to finish (and once, as a kernel of its own, before the first iteration).
__device__ void advance(QueueState* q, uint32_t capacity, uint32_t finish_paths) {
    q->rays += q->n_ext;
    uint32_t n_cont = atomicExch(&q->n_next, 0u);  // the other blocks appended with atomics: read where they wrote (L2), not a cached copy
    for (int k = 0; k < Q_COUNT; ++k) q->n_shade[k] = 0;
    unsigned long long remaining = q->total_work - q->next_work;
    uint32_t n_new = (uint32_t)min((unsigned long long)(capacity - n_cont), remaining);
    q->gen_base = q->next_work;
    q->next_work += n_new;
    q->finish_n = 0;
    if (remaining == 0 && n_cont > 0 && n_cont <= finish_paths) {
        // the job is draining: no samples left to regenerate and only a few paths alive. Instead of up to max_depth more
        // wavefront iterations over a nearly empty queue, k_finish runs each remaining path to its end in this iteration.
        q->finish_n = n_cont;
        n_cont = 0;
    }
    q->n_cont = n_cont;
    q->n_new = n_new;
    q->n_ext = n_cont + n_new;
    q->ext_cursor = 0;
    q->done = (n_cont + n_new == 0 && q->finish_n == 0) ? 1u : 0u;
    if (n_cont + n_new + q->finish_n) q->iterations++;
}
__global__ void k_advance(QueueState* q, uint32_t capacity, uint32_t finish_paths) {
    if (threadIdx.x == 0 && blockIdx.x == 0) advance(q, capacity, finish_paths);
}

// Camera::ray world.rs:53-63 with the pixel jitter of main.rs:258-259 (jitter == false: pixel centres, main.rs:189-190)
__device__ __forceinline__ Ray camera_ray(const DCamera& cam, const RenderParams& rp, uint32_t pixel, uint32_t sample, bool jitter) {
    uint32_t x = pixel % rp.w, y = pixel / rp.w;
    RngKey key{pixel, sample, kBounceCamera, rp.seed};
    Rand4 xi = draw4(key, kStreamScatter);
    float s = jitter ? ((float)x + xi.x) / (float)(rp.w - 1) : (float)x / (float)(rp.w - 1);
    float t = jitter ? ((float)y + xi.y) / (float)(rp.h - 1) : (float)y / (float)(rp.h - 1);
    V3 blur = sample_unit_disk(xi.z, xi.w) * cam.lens_radius;
    V3 offset = v3(cam.u) * blur.x + v3(cam.v) * blur.y;
    Ray r;
    r.o = v3(cam.origin) + offset;
    r.d = v3(cam.llc) + (v3(cam.horizontal) * s) + (v3(cam.vertical) * t) - v3(cam.origin) - offset;
    return r;
}

// Work item k of a job is pixel k mod w*h, sample spp_begin + k div w*h. The 64-bit division is done once per thread (for the first
// item of this launch); every item of the launch is then a 32-bit offset from it.
__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ DCamera cam, const __grid_constant__ RenderParams rp, Pool pool,
                                                  QueueState* q, int cur) {
    const uint32_t n_new = q->n_new, n_cont = q->n_cont;
    const unsigned long long base = q->gen_base;
    const uint32_t pixel0 = (uint32_t)(base % rp.npix), sample0 = rp.spp_begin + (uint32_t)(base / rp.npix);
    RayRec* __restrict__ out = pool.q_ext[cur] + n_cont;  // new paths follow the continuing ones: whole warps of neighbouring pixels
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_new; i += gridDim.x * blockDim.x) {
        const uint32_t p = pixel0 + i;  // < 2^31 + 2^26
        const uint32_t wrap = p / rp.npix;
        const uint32_t pixel = p - wrap * rp.npix;
        const uint32_t sample = sample0 + wrap;
        Ray r = camera_ray(cam, rp, pixel, sample, true);
        out[i].o = make_float4(r.o.x, r.o.y, r.o.z, __uint_as_float(pixel));
        out[i].d = make_float4(r.d.x, r.d.y, r.d.z, __uint_as_float(sample));
        out[i].thr = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(0u));
    }
}

// Persistent warps. Every thread keeps one traversal in flight and, between refills, descends to its next leaf and tests it
// ("chain": all lanes run the node loop together, then the leaf code of each primitive kind present). When at least
// `refill_lanes` lanes of a warp have finished, the warp commits their hits -- hit record, then one __match_any_sync + one
// atomicAdd per (warp, material kind) to append to the material-sorted shade queues -- and refills those lanes with one
// atomicAdd on the queue cursor. Measured on B200 (profiles/README.md): on scenes whose trees are shallow (spheres, boxes, small
// meshes, however many instances) the best threshold is 32, i.e. a warp takes 32 consecutive queue entries, runs them to the
// end and commits them together. Consecutive entries are neighbouring pixels or paths shaded together, so they start in phase
// (all at the root, then all a few nodes from their next leaf); refilling lanes early -- even for the price of a few
// shared-memory loads from a prefetched ring -- mixes rays that need five node visits with rays that need one and costs more
// warp instructions than the idle lanes save. Under a deep BLAS (a mesh of tens of thousands of triangles or more) ray lengths
// vary so much that the idle tail dominates (5-7 of 32 lanes active on bounce rays): there continuing rays are refilled at 12
// idle lanes, and a lane makes at most 4 node visits before the warp looks at its pending leaves again, so that a freshly
// started ray (20 levels from its first leaf) does not hold up the lanes that are one or two visits from theirs
// (-13 % / -17 % k_extend time on the 1 M- / 10 M-triangle scenes); camera rays always run as whole batches.
// Leaving an instance is free: stack entries pushed before the instance was entered lie below inst_base, and popping one
// restores the world-space ray from shared memory (no sentinel entries).
constexpr uint32_t kRefillLanes = 32, kRefillLanesDeep = 12;  // defaults (shallow / deep BLAS); MRT_OPT_REFILL_LANES overrides them
constexpr uint32_t kNodeBurst = 0xFFFFFFFFu, kNodeBurstDeep = 4;  // node visits per lane between two looks at the leaves; MRT_OPT_NODE_BURST
constexpr int kDeepBlas = 12;                                 // inner-node levels from which a BLAS counts as deep
constexpr int kExtendThreads = 128;

template <bool COUNT, bool ALPHA, bool VOLUME>
__global__ void __launch_bounds__(kExtendThreads, COUNT || ALPHA ? 1 : (VOLUME ? 7 : 8)) k_extend(const __grid_constant__ DScene sc, const __grid_constant__ RenderParams rp,
                                                                                                  Pool pool, QueueState* q, int cur) {
    // the path record of the ray each thread has in flight, component-major (conflict-free): 0-5 world-space ray (needed again
    // when a lane leaves an instance), 6 pixel, 7 sample, 8-10 throughput, 11 bounce. It is copied to the shade queue at commit,
    // so nothing but the traversal state occupies registers.
    __shared__ float s_path[12 * kExtendThreads];
    float* const mine = &s_path[threadIdx.x];
    WorldRayShared<kExtendThreads> ws{mine};
    const uint32_t n = q->n_ext;
    const RayRec* __restrict__ queue = pool.q_ext[cur];
    const float inf = __int_as_float(0x7f800000);
    const uint32_t lane = lane_id();
    const uint32_t lt_mask = (1u << lane) - 1u;
    // refill threshold: new camera rays (the tail of the queue, entries >= n_cont) are coherent and start in phase, so a warp
    // always runs a batch of them to the end (32); continuing rays use rp.refill_lanes
    const uint32_t n_cont = q->n_cont;
    uint32_t refill_lanes = rp.refill_lanes;
    const uint32_t node_burst = rp.node_burst;
    VisitCounters cnt{0, 0, 0, 0, 0, 0};
    unsigned long long ray_start = 0;  // (COUNT) node visits when the lane's current ray began
    (void)ray_start;
    uint32_t stack[kStackSize];
    Traversal T;
    T.sp = 0;
    T.inst_base = 0;
    T.ref = kNone;
    T.best = HitRec{inf, kNone, kNone};
    RngKey key{0u, 0u, 0u, rp.seed};
    bool active = false, pending = false, drained = false;
    for (;;) {
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle == 0xffffffffu || (!drained && (uint32_t)__popc(idle) >= refill_lanes)) {
            // ---- commit finished rays: world.rs:68 result -> shade queue of the hit material's kind --------------------
            uint32_t kind = 0xFFu;
            HitRec h = T.best;
            int32_t m = -1;
            if (pending) {
                if (COUNT) cnt.max_ray_nodes = max(cnt.max_ray_nodes, cnt.node_visits - ray_start);
#ifdef MRT_TRACE_MONSTERS
                if (COUNT && cnt.node_visits - ray_start > 100000ull)
                    printf("monster ray: %llu node visits, pixel %u sample %u bounce %u o (%.9g %.9g %.9g) d (%.9g %.9g %.9g) hit t %.9g prim %08x\n", cnt.node_visits - ray_start,
                           __float_as_uint(mine[6 * kExtendThreads]), __float_as_uint(mine[7 * kExtendThreads]), __float_as_uint(mine[11 * kExtendThreads]), mine[0],
                           mine[kExtendThreads], mine[2 * kExtendThreads], mine[3 * kExtendThreads], mine[4 * kExtendThreads], mine[5 * kExtendThreads], h.t, h.prim);
#endif
                if (h.prim == kNone) h.t = inf;
                kind = Q_MISS;
                if (h.prim != kNone) {
                    m = hit_material(sc, h);
                    kind = Q_FIRST_MAT + (uint32_t)sc.materials[m].kind;
                }
            }
            const uint32_t peers = __match_any_sync(0xffffffffu, kind);
            if (pending) {
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t qbase = 0;
                if (lane == leader) qbase = atomicAdd(&q->n_shade[kind], __popc(peers));
                qbase = __shfl_sync(peers, qbase, leader);
                HitEntry* dst = pool.q_shade + (size_t)rp.region[kind] * pool.capacity + qbase + __popc(peers & lt_mask);
                dst->ray.o = make_float4(mine[0], mine[kExtendThreads], mine[2 * kExtendThreads], mine[6 * kExtendThreads]);
                dst->ray.d = make_float4(mine[3 * kExtendThreads], mine[4 * kExtendThreads], mine[5 * kExtendThreads], mine[7 * kExtendThreads]);
                dst->ray.thr = make_float4(mine[8 * kExtendThreads], mine[9 * kExtendThreads], mine[10 * kExtendThreads], mine[11 * kExtendThreads]);
                dst->hit = make_uint4(__float_as_uint(h.t), h.prim, h.inst, (uint32_t)m);
                pending = false;
            }
            // ---- refill idle lanes: consecutive queue entries, read with coalesced 128-bit loads -----------------------
            if (!drained) {
                const uint32_t want = __popc(idle);
                const uint32_t leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&q->ext_cursor, want);
                base = __shfl_sync(0xffffffffu, base, leader);
                if (base + want > n_cont) refill_lanes = 32u;
                const uint32_t i = base + __popc(idle & lt_mask);
                if (!active && i < n) {
                    const RayRec* rec = &queue[i];
                    const float4 o = rec->o, d = rec->d, thr = rec->thr;
                    if (ALPHA || VOLUME) {
                        key.pixel = __float_as_uint(o.w);
                        key.sample = __float_as_uint(d.w);
                        key.bounce = __float_as_uint(thr.w);
                    }
                    mine[6 * kExtendThreads] = o.w;
                    mine[7 * kExtendThreads] = d.w;
                    mine[8 * kExtendThreads] = thr.x;
                    mine[9 * kExtendThreads] = thr.y;
                    mine[10 * kExtendThreads] = thr.z;
                    mine[11 * kExtendThreads] = thr.w;
                    trav_begin(sc, T, ws, Ray{V3{o.x, o.y, o.z}, V3{d.x, d.y, d.z}}, inf);  // world.rs:68: [0.001, +inf); stores the ray in s_path 0-5
                    active = true;
                    if (COUNT) ray_start = cnt.node_visits;
                }
                drained = base + want >= n;
            }
            if (__ballot_sync(0xffffffffu, active) == 0) break;
        }
        if (active && T.ref == kNone && !trav_pop(T, stack, ws)) {
            active = false;
            pending = true;
        }
        if (active) {
            // at most node_burst node visits before the warp looks at its leaves again (0xFFFFFFFF: descend all the way to the next leaf)
            if (node_burst == kNodeBurst) {  // (warp-uniform) the unbounded loop carries no counter
                while (ref_is_node(T.ref)) trav_node<COUNT>(sc, T, stack, 0.001f, &cnt);
            } else {
                for (uint32_t k = 0; k < node_burst && ref_is_node(T.ref); ++k) trav_node<COUNT>(sc, T, stack, 0.001f, &cnt);
            }
            if (T.ref != kNone && !ref_is_node(T.ref)) trav_leaf<COUNT, ALPHA, VOLUME>(sc, T, stack, ws, 0.001f, key, &cnt);
        }
    }
    if (COUNT) {
        unsigned long long v[5] = {cnt.node_visits, cnt.tri_tests, cnt.sphere_tests, cnt.instance_tests, cnt.volume_tests};
        unsigned long long* dst = &q->visits.node_visits;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            unsigned long long x = v[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
            if (lane == 0 && x) atomicAdd(&dst[k], x);
        }
        unsigned long long mx = cnt.max_ray_nodes;
        for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, off));
        if (lane == 0 && mx) atomicMax(&q->visits.max_ray_nodes, mx);
    }
}

// One queue entry of Camera::trace's hit/miss handling (world.rs:69-77) in iterative form:
// radiance += throughput * emitted; throughput *= attenuation. On return with cont == true, `e.ray` is the scattered ray.
template <bool FULL>
__device__ __forceinline__ void shade_entry(const DScene& sc, const RenderParams& rp, long long* accum, uint32_t* nonfinite, uint32_t kind, HitEntry& e,
                                            bool& cont) {
    const float4 o = e.ray.o, d = e.ray.d, th = e.ray.thr;
    const uint32_t pixel = __float_as_uint(o.w), sample = __float_as_uint(d.w);
    uint32_t bounce = __float_as_uint(th.w);
    Ray ray{V3{o.x, o.y, o.z}, V3{d.x, d.y, d.z}};
    V3 thr{th.x, th.y, th.z};
    cont = false;
    if (kind == Q_MISS) {  // world.rs:77
        accumulate(accum, nonfinite, pixel, thr * background(sc, ray));
    } else {
        const uint4 hr = e.hit;
        HitRec h{__uint_as_float(hr.x), hr.y, hr.z};
        RngKey key{pixel, sample, bounce, rp.seed};
        mrt_material mat = sc.materials[(int32_t)hr.w];
        mrt_material emat = mat;
        if (mat.kind == MRT_MAT_MIX) {  // Mix::emit and Mix::scatter flip independent coins (material.rs:403-417)
            emat = pick_material(sc, (int32_t)hr.w, key, kStreamEmit);
            mat = pick_material(sc, (int32_t)hr.w, key, kStreamMix);
        }
        Surfel s = resolve_hit<FULL>(sc, ray, h, (int32_t)hr.w);
        if (emat.kind == MRT_MAT_DIFFUSE_LIGHT) accumulate(accum, nonfinite, pixel, thr * V3{emat.p[0], emat.p[1], emat.p[2]});  // world.rs:69
        else if (FULL && emat.kind == MRT_MAT_EVE && s.has_uv) accumulate(accum, nonfinite, pixel, thr * eve_emission(sc, emat, s.u, s.v));  // eve.rs:121-128
        ScatterOut sco;
        scatter_kind<FULL>(sc, mat, ray, s, draw4(key, kStreamScatter), sco);
        if (sco.scattered) {
            thr = thr * sco.attenuation;
            bounce += 1;
            if (bounce < rp.max_depth) {  // world.rs:66: the next call would be depth == 0 -> contributes 0
                cont = true;
                e.ray.o = make_float4(s.point.x, s.point.y, s.point.z, o.w);
                e.ray.d = make_float4(sco.dir.x, sco.dir.y, sco.dir.z, d.w);
                e.ray.thr = make_float4(thr.x, thr.y, thr.z, __uint_as_float(bounce));
            }
        }
    }
    if (!cont && bounce)  // buffer.set(.., MAX_DEPTH - depth) main.rs:263, merged at :635
        atomicAdd(reinterpret_cast<unsigned long long*>(&accum[(size_t)pixel * 4 + 3]), (unsigned long long)bounce);
}

#ifndef MRT_SHADE_THREADS
#define MRT_SHADE_THREADS 256
#endif
#ifndef MRT_SHADE_MINB
#define MRT_SHADE_MINB 3
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Each warp takes 32 consecutive entries of one material queue at a time (so a warp shades one material kind), streams them in
// with coalesced 128-bit loads -- the entries of its next turn are prefetched into L2 meanwhile -- and appends the scattered
// rays of the surviving paths to the next extend queue with one atomic per warp.
template <bool FULL>
__global__ void __launch_bounds__(MRT_SHADE_THREADS, MRT_SHADE_MINB) k_shade(const __grid_constant__ DScene sc, const __grid_constant__ RenderParams rp, Pool pool,
                                                                            QueueState* q, int cur, long long* accum, uint32_t* nonfinite,
                                                                            uint32_t next_finish_paths) {
    RayRec* __restrict__ q_next = pool.q_ext[cur ^ 1];
    const uint32_t lane = lane_id(), lt_mask = (1u << lane) - 1u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    // the nine queue lengths: one load per lane, broadcast per kind (instead of nine dependent round trips per warp)
    const uint32_t my_count = lane < Q_COUNT ? q->n_shade[lane] : 0u;
    for (uint32_t kind = 0; kind < Q_COUNT; ++kind) {
        const uint32_t n = __shfl_sync(0xffffffffu, my_count, kind);
        if (n == 0) continue;
        const HitEntry* __restrict__ queue = pool.q_shade + (size_t)rp.region[kind] * pool.capacity;
        for (uint32_t base = warp * 32u; base < n; base += n_warps * 32u) {
            const uint32_t i = base + lane;
            const bool valid = i < n;
            const uint32_t ahead = i + n_warps * 32u;
            if (ahead < n) {
                prefetch_l2(&queue[ahead]);
                prefetch_l2(reinterpret_cast<const char*>(&queue[ahead]) + 32);
            }
            bool cont = false;
            HitEntry e;
            if (valid) {
                e.ray.o = queue[i].ray.o;
                e.ray.d = queue[i].ray.d;
                e.ray.thr = queue[i].ray.thr;
                e.hit = queue[i].hit;
                shade_entry<FULL>(sc, rp, accum, nonfinite, kind, e, cont);
            }
            const uint32_t mc = __ballot_sync(0xffffffffu, cont);
            if (mc) {
                uint32_t base_c = 0;
                if (lane == 0) base_c = atomicAdd(&q->n_next, __popc(mc));
                base_c = __shfl_sync(0xffffffffu, base_c, 0);
                if (cont) {
                    RayRec* dst = &q_next[base_c + __popc(mc & lt_mask)];
                    dst->o = e.ray.o;
                    dst->d = e.ray.d;
                    dst->thr = e.ray.thr;
                }
            }
        }
    }
    // the last block to get here closes the iteration (saves a one-thread kernel launch per iteration)
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // this block's appends to n_next are visible before its ticket
        s_last = atomicAdd(&q->shade_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        q->shade_done = 0;
        advance(q, pool.capacity, next_finish_paths);
    }
}

// Drain: one thread per remaining path, each looping extend + shade until its path ends (paths are independent, so no
// grid-wide step is needed). Launched every iteration; returns at once unless k_advance has set finish_n.
template <bool ALPHA, bool VOLUME>
__global__ void __launch_bounds__(128) k_finish(const __grid_constant__ DScene sc, const __grid_constant__ RenderParams rp, Pool pool, QueueState* q,
                                                int cur, long long* accum, uint32_t* nonfinite) {
    const uint32_t n = q->finish_n;
    if (n == 0) return;
    const RayRec* __restrict__ queue = pool.q_ext[cur];
    const float inf = __int_as_float(0x7f800000);
    unsigned long long rays = 0;
    // Consecutive threads take queue entries n / 32 apart: neighbouring entries were shaded together and tend to be paths of one kind --
    // the ones that stay inside a mesh for all max_depth bounces sit next to each other, and a warp that took 32 of them would run
    // every one of their bounces with 32 diverged lanes.
    const uint32_t stride = (n + 31u) / 32u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < 32u * stride; i += gridDim.x * blockDim.x) {
        const uint32_t item = (i & 31u) * stride + (i >> 5);
        if (item >= n) continue;
        HitEntry e;
        e.ray = queue[item];
        bool cont = true;
        while (cont) {
            const float4 o = e.ray.o, d = e.ray.d;
            RngKey key{__float_as_uint(o.w), __float_as_uint(d.w), __float_as_uint(e.ray.thr.w), rp.seed};
            HitRec h = traverse<false, ALPHA, VOLUME>(sc, Ray{V3{o.x, o.y, o.z}, V3{d.x, d.y, d.z}}, 0.001f, inf, key, nullptr);
            int32_t m = -1;
            uint32_t kind = Q_MISS;
            if (h.prim != kNone) {
                m = hit_material(sc, h);
                kind = Q_FIRST_MAT + (uint32_t)sc.materials[m].kind;
            }
            e.hit = make_uint4(__float_as_uint(h.t), h.prim, h.inst, (uint32_t)m);
            shade_entry<ALPHA>(sc, rp, accum, nonfinite, kind, e, cont);
            ++rays;
        }
    }
    for (int off = 16; off > 0; off >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, off);
    if (lane_id() == 0 && rays) atomicAdd(&q->rays, rays);
}

// PASS A (main.rs:166-222): Camera::albedo_normal (world.rs:81-93) at pixel centres
template <bool ALPHA, bool VOLUME>
__global__ void __launch_bounds__(128) k_aov(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam, const __grid_constant__ RenderParams rp,
                                             float* albedo, float* normal, uint32_t* object_id, uint32_t* tri_id, float* t_out) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= rp.npix) return;
    const float inf = __int_as_float(0x7f800000);
    Ray ray = camera_ray(cam, rp, pixel, 0u, false);
    RngKey key{pixel, 0u, 0u, rp.seed};
    HitRec h = traverse<false, ALPHA, VOLUME>(sc, ray, 0.001f, inf, key, nullptr);
    V3 a, n{0.0f, 0.0f, 0.0f};
    uint32_t obj = kNone, tri = kNone;
    float t = inf;
    if (h.prim != kNone) {
        int32_t m = hit_material(sc, h);
        mrt_material mat = sc.materials[m], emat = mat;
        if (mat.kind == MRT_MAT_MIX) {
            emat = pick_material(sc, m, key, kStreamEmit);
            mat = pick_material(sc, m, key, kStreamMix);
        }
        Surfel s = resolve_hit<ALPHA>(sc, ray, h, m);
        V3 emitted = (emat.kind == MRT_MAT_DIFFUSE_LIGHT) ? V3{emat.p[0], emat.p[1], emat.p[2]} : V3{0.0f, 0.0f, 0.0f};
        if (ALPHA && emat.kind == MRT_MAT_EVE && s.has_uv) emitted = eve_emission(sc, emat, s.u, s.v);
        ScatterOut sco;
        scatter_kind<ALPHA>(sc, mat, ray, s, draw4(key, kStreamScatter), sco);
        a = sco.scattered ? sco.attenuation : emitted;
        n = s.normal;
        obj = s.object_id;
        tri = s.tri_id;
        t = h.t;
    } else {
        a = background(sc, ray);
    }
    if (albedo) { albedo[3 * (size_t)pixel] = a.x; albedo[3 * (size_t)pixel + 1] = a.y; albedo[3 * (size_t)pixel + 2] = a.z; }
    if (normal) { normal[3 * (size_t)pixel] = n.x; normal[3 * (size_t)pixel + 1] = n.y; normal[3 * (size_t)pixel + 2] = n.z; }
    if (object_id) object_id[pixel] = obj;
    if (tri_id) tri_id[pixel] = tri;
    if (t_out) t_out[pixel] = t;
}

// accumulators -> the reference's Image.pixels view: (sum colour f32, sum bounces u32)  main.rs:599
__global__ void k_resolve_sums(const long long* accum, const uint32_t* nonfinite, uint32_t npix, float* sum_rgb, uint32_t* sum_bounces) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        uint32_t bad = nonfinite[p];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = (float)((double)accum[(size_t)p * 4 + k] * 2.3283064365386963e-10);  // * 2^-32
            if (bad & (1u << k)) v = __int_as_float(0x7fc00000);
            sum_rgb[(size_t)p * 3 + k] = v;
        }
        sum_bounces[p] = (uint32_t)accum[(size_t)p * 4 + 3];
    }
}

// Image::to_rgb_bytes main.rs:640-722 (Default and Depth modes) with the row flip of Image::dump :763-768
__global__ void k_resolve_rgb8(const long long* accum, const uint32_t* nonfinite, uint32_t w, uint32_t h, uint32_t count, int mode, int flip,
                               uint32_t max_bounces, uint8_t* out) {
    const uint32_t npix = w * h;
    const float scale = 1.0f / (float)count;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        uint32_t y = p / w, x = p % w;
        uint32_t dst = (flip ? (h - 1 - y) : y) * w + x;
        float px[3];
        if (count == 0) {
            px[0] = px[1] = px[2] = 0.0f;
        } else if (mode == 2) {  // Depth :655-665
            float max_depth = (float)max(max_bounces, 1u) * scale;
            float dpt = fminf(fmaxf(((float)(uint32_t)accum[(size_t)p * 4 + 3] * scale) / max_depth, 0.0f), 1.0f);
            px[0] = px[1] = px[2] = dpt;
        } else {  // Default :667-673
            uint32_t bad = nonfinite[p];
            for (int k = 0; k < 3; ++k) {
                float v = (float)((double)accum[(size_t)p * 4 + k] * 2.3283064365386963e-10);
                if (bad & (1u << k)) v = __int_as_float(0x7fc00000);
                px[k] = fmaxf(fminf(powf(scale * v, 1.0f / 2.2f), 1.0f), 0.0f);  // NaN -> 1 (f32::min drops NaN)
            }
        }
        for (int k = 0; k < 3; ++k) {
            float v = px[k] * 255.0f;
            out[(size_t)dst * 3 + k] = (uint8_t)(v >= 255.0f ? 255 : (v > 0.0f ? (int)v : 0));  // `as u8` truncates and saturates
        }
    }
}
__global__ void k_max_bounces(const long long* accum, uint32_t npix, uint32_t* out) {
    uint32_t m = 0;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) m = max(m, (uint32_t)accum[(size_t)p * 4 + 3]);
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, off));
    if (lane_id() == 0) atomicMax(out, m);
}

// test hooks (include/mrt_debug.h)
__global__ void k_debug_philox(uint4 ctr, uint2 key, uint4* out) { *out = philox4x32_10(ctr, key); }
__global__ void k_debug_samplers(uint2 seed, uint32_t n, float* ball, float* sphere, float* disk) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        RngKey key{i, 0u, 0u, seed};
        Rand4 xi = draw4(key, 0);
        V3 b = sample_unit_ball(xi.x, xi.y, xi.z), s = sample_unit_vector(xi.x, xi.y), d = sample_unit_disk(xi.z, xi.w);
        ball[3 * i] = b.x; ball[3 * i + 1] = b.y; ball[3 * i + 2] = b.z;
        sphere[3 * i] = s.x; sphere[3 * i + 1] = s.y; sphere[3 * i + 2] = s.z;
        disk[2 * i] = d.x; disk[2 * i + 1] = d.y;
    }
}

}  // namespace mrt

using namespace mrt;

// ---------------------------------------------------------------------------------------------------------------
// host side of the C ABI
// ---------------------------------------------------------------------------------------------------------------
struct mrt_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int n_sms = 0;
    // scene
    // scene arrays live in a grow-only device arena (one buffer per array kind) so that re-uploading a scene of similar size
    // -- every frame of an animation, main.rs:104-117 -- costs no cudaMalloc / cudaFree; host data goes through two pinned
    // staging chunks so that the copies are real DMA and overlap the host-side memcpy into the next chunk
    struct DevBuf { void* p = nullptr; size_t cap = 0; size_t used = 0; };
    std::vector<DevBuf> scene_bufs;
    size_t scene_buf_next = 0;
    void* stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    int stage_next = 0;
    DScene scene{};
    bool has_scene = false;
    uint32_t material_kinds = 0;  // bit per MRT_MAT_* kind present in the scene's material table
    DCamera cam{};
    bool has_camera = false;
    uint64_t scene_bytes = 0;
    // image
    uint32_t w = 0, h = 0, count = 0;
    long long* d_accum = nullptr;
    uint32_t* d_nonfinite = nullptr;
    float* d_sum_rgb = nullptr;  // staging for mrt_accum_download / mrt_resolve_rgb8 (allocated with the image: no malloc per call)
    uint32_t* d_sum_b = nullptr;
    // pool
    Pool pool{};
    QueueState* d_q = nullptr;
    QueueState* h_q = nullptr;  // pinned, 2 status slots + 1 final
    cudaEvent_t ev_status[2] = {nullptr, nullptr};
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // options
    bool opt_count = false, opt_time = false;
    uint64_t opt_pool_slots = 0;
    uint32_t opt_finish_paths = 98304;  // 64 K / 96 K / 128 K / 256 K measured (profiles/r02_drain_cost.txt): 96 K is best or equal on every workload
    uint32_t opt_leaf_tris = 4, opt_tri_cost = 100;
    uint32_t opt_device_build = 1;  // MRT_OPT_DEVICE_BUILD: 0 host SAH everywhere; 1 GPU LBVH for big meshes (and for a TLAS of >= 2^20 objects);
                                    // 2 also for a TLAS of >= 16384 objects
    DevBuf build_scratch;           // raw vertices + work arrays of the GPU builder (grow-only)
    DevBuf aov_buf;                 // output planes of mrt_render_aov (grow-only)
    int* d_depths = nullptr;        // tree depth of each GPU-built BLAS
    int* h_depths = nullptr;        // pinned
    uint32_t opt_refill_lanes = 0;  // 0 = by scene: kRefillLanesDeep under a deep BLAS, else kRefillLanes
    uint32_t auto_refill_lanes = kRefillLanes, auto_node_burst = kNodeBurst;
    uint32_t opt_node_burst = 0;  // 0 = by scene
    mrt_stats stats{};
    // more than one GPU (mrt_comm.inc): a communicator, and for a multi-device handle the contexts of the other devices
    void* comm = nullptr;  // ncclComm_t
    int comm_rank = 0, comm_size = 1;
    bool comm_split = true;               // MRT_OPT_COMM_SPLIT
    bool comm_scene = true;               // MRT_OPT_COMM_SCENE: on a one-process-per-GPU communicator the root's scene is built once and broadcast
    void* d_scene_hdr = nullptr;          // device + pinned copies of the broadcast header
    void* h_scene_hdr = nullptr;
    std::vector<mrt_context*> peers;      // devices[1..] of mrt_context_create_multi (owned by this, the leader)
    unsigned long long* d_count = nullptr;  // sample-count cell of the reduce
    unsigned long long* h_count = nullptr;  // pinned
    int grid_extend[2][3] = {{0, 0, 0}, {0, 0, 0}};  // [count visits][0 plain, 1 volumes, 2 alpha-tested triangles (+ volumes)]
    int grid_shade = 0, grid_shade_full = 0, grid_generate = 0;
};

static std::string g_create_error;

#define MRT_CUDA(call)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                  \
            return MRT_E_CUDA;                                                                              \
        }                                                                                                   \
    } while (0)

static int fail(mrt_context* ctx, int code, const std::string& msg) {
    ctx->err = msg;
    return code;
}

#include "mrt_comm.inc"

constexpr size_t kStageChunk = 32u << 20;

// memcpy split over a few threads (one thread tops out near 10 GB/s; the scene arrays of a 10 M-triangle scene are ~2 GB)
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t kMin = 4u << 20;
    int parts = (int)std::min<size_t>(4, bytes / kMin);
    if (parts <= 1) { std::memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    for (int k = 1; k < parts; ++k) {
        size_t a = bytes * (size_t)k / (size_t)parts, b = bytes * (size_t)(k + 1) / (size_t)parts;
        th.emplace_back([=] { std::memcpy(static_cast<char*>(dst) + a, static_cast<const char*>(src) + a, b - a); });
    }
    std::memcpy(dst, src, bytes / (size_t)parts);
    for (auto& t : th) t.join();
}

// run fn(first, last) over [0, n) in parallel slices (host-side packing loops over millions of triangles)
template <class Fn>
static void parallel_for(size_t n, Fn fn) {
    const size_t kMin = 1u << 16;
    int parts = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), n / kMin);
    if (parts <= 1) { fn((size_t)0, n); return; }
    std::vector<std::thread> th;
    for (int k = 1; k < parts; ++k) th.emplace_back([=] { fn(n * (size_t)k / (size_t)parts, n * (size_t)(k + 1) / (size_t)parts); });
    fn((size_t)0, n / (size_t)parts);
    for (auto& t : th) t.join();
}

static int grow(mrt_context* ctx, mrt_context::DevBuf& buf, size_t bytes) {
    if (buf.cap >= bytes) return MRT_OK;
    if (buf.p) cudaFree(buf.p);
    buf = mrt_context::DevBuf{};
    const size_t cap = bytes + bytes / 8;  // a little headroom: an animated scene rarely keeps its exact size
    MRT_CUDA(cudaMalloc(&buf.p, cap));
    buf.cap = cap;
    return MRT_OK;
}

// next array of the scene arena, `bytes` long
static int arena_alloc(mrt_context* ctx, size_t bytes, void** out) {
    *out = nullptr;
    bytes = std::max<size_t>(bytes, 16);
    if (ctx->scene_buf_next >= ctx->scene_bufs.size()) ctx->scene_bufs.emplace_back();
    mrt_context::DevBuf& buf = ctx->scene_bufs[ctx->scene_buf_next++];
    int rc = grow(ctx, buf, bytes);
    if (rc) return rc;
    ctx->scene_bytes += bytes;
    buf.used = bytes;
    *out = buf.p;
    return MRT_OK;
}

// host -> device through the two pinned staging chunks: the memcpy into one chunk overlaps the DMA out of the other
static int stage_copy(mrt_context* ctx, void* dst, const void* src, size_t bytes) {
    const char* from = static_cast<const char*>(src);
    for (size_t off = 0; off < bytes; off += kStageChunk) {
        const size_t len = std::min(kStageChunk, bytes - off);
        const int k = ctx->stage_next;
        ctx->stage_next ^= 1;
        if (!ctx->stage[k]) {
            MRT_CUDA(cudaMallocHost(&ctx->stage[k], kStageChunk));
            MRT_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[k], cudaEventDisableTiming));
        } else {
            MRT_CUDA(cudaEventSynchronize(ctx->stage_ev[k]));  // the DMA that last read this chunk has finished
        }
        parallel_memcpy(ctx->stage[k], from + off, len);
        MRT_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + off, ctx->stage[k], len, cudaMemcpyHostToDevice, ctx->stream));
        MRT_CUDA(cudaEventRecord(ctx->stage_ev[k], ctx->stream));
    }
    return MRT_OK;
}

template <class T>
static int upload(mrt_context* ctx, const T* src, size_t n, const T** dst) {
    void* p = nullptr;
    int rc = arena_alloc(ctx, n * sizeof(T), &p);
    if (rc) return rc;
    if (n && (rc = stage_copy(ctx, p, src, n * sizeof(T)))) return rc;
    *dst = static_cast<const T*>(p);
    return MRT_OK;
}

// the arena is kept for the next upload; only the context's destruction releases it
static void free_scene(mrt_context* ctx) {
    ctx->scene_buf_next = 0;
    ctx->scene_bytes = 0;
    ctx->has_scene = false;
}
static void release_scene_arena(mrt_context* ctx) {
    for (auto& b : ctx->scene_bufs) cudaFree(b.p);
    ctx->scene_bufs.clear();
    cudaFree(ctx->build_scratch.p);
    ctx->build_scratch = mrt_context::DevBuf{};
    cudaFree(ctx->aov_buf.p);
    ctx->aov_buf = mrt_context::DevBuf{};
    cudaFree(ctx->d_depths);
    if (ctx->h_depths) cudaFreeHost(ctx->h_depths);
    ctx->d_depths = ctx->h_depths = nullptr;
    for (int k = 0; k < 2; ++k) {
        if (ctx->stage[k]) cudaFreeHost(ctx->stage[k]);
        if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]);
        ctx->stage[k] = nullptr;
        ctx->stage_ev[k] = nullptr;
    }
}
static void free_pool(mrt_context* ctx) {
    Pool& p = ctx->pool;
    cudaFree(p.q_ext[0]); cudaFree(p.q_ext[1]); cudaFree(p.q_shade);
    p = Pool{};
}
static void free_image(mrt_context* ctx) {
    cudaFree(ctx->d_accum);
    cudaFree(ctx->d_nonfinite);
    cudaFree(ctx->d_sum_rgb);
    cudaFree(ctx->d_sum_b);
    ctx->d_accum = nullptr;
    ctx->d_nonfinite = nullptr;
    ctx->d_sum_rgb = nullptr;
    ctx->d_sum_b = nullptr;
    ctx->w = ctx->h = ctx->count = 0;
}

// Multi-device handle: the scene is validated, built and uploaded ONCE (on devices[0]); the other devices receive the finished
// device arrays over NVLink (cudaMemcpyPeerAsync, array by array of the arena) with the pointers of DScene rebased.
static int replicate_scene(mrt_context* from, mrt_context* to) {
    mrt_context* ctx = to;  // MRT_CUDA reports on the receiving context
    MRT_CUDA(cudaSetDevice(to->device));
    cudaStreamSynchronize(to->stream);
    free_scene(to);
    const size_t n = from->scene_buf_next;
    std::vector<std::pair<const char*, const char*>> map(n);  // leader array -> this device's copy
    for (size_t k = 0; k < n; ++k) {
        const mrt_context::DevBuf& src = from->scene_bufs[k];
        void* dst = nullptr;
        int rc = arena_alloc(to, src.used, &dst);
        if (rc) return rc;
        MRT_CUDA(cudaMemcpyPeerAsync(dst, to->device, src.p, from->device, src.used, to->stream));
        map[k] = {static_cast<const char*>(src.p), static_cast<const char*>(dst)};
    }
    auto rebase = [&](auto*& ptr) {
        if (!ptr) return;
        const char* q = reinterpret_cast<const char*>(ptr);
        for (size_t k = 0; k < n; ++k) {
            if (q >= map[k].first && q < map[k].first + std::max<size_t>(from->scene_bufs[k].used, 1)) {
                ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(map[k].second + (q - map[k].first));
                return;
            }
        }
    };
    DScene d = from->scene;
    rebase(d.nodes); rebase(d.spheres); rebase(d.sphere_aux); rebase(d.tri_verts); rebase(d.tri_map); rebase(d.tri_shading);
    rebase(d.instances); rebase(d.blas); rebase(d.volumes); rebase(d.materials); rebase(d.surfaces); rebase(d.textures); rebase(d.texels);
    MRT_CUDA(cudaStreamSynchronize(to->stream));
    to->scene = d;
    to->has_scene = true;
    to->auto_refill_lanes = from->auto_refill_lanes;
    to->auto_node_burst = from->auto_node_burst;
    to->material_kinds = from->material_kinds;
    return MRT_OK;
}

// runs fn(member) for the leader (on the calling thread) and for each peer device (one thread each); first error wins
template <class Fn>
static int for_each_member(mrt_context* ctx, Fn fn) {
    if (ctx->peers.empty()) return fn(ctx);
    std::vector<int> rc(ctx->peers.size(), MRT_OK);
    std::vector<std::thread> th;
    for (size_t k = 0; k < ctx->peers.size(); ++k) th.emplace_back([&, k] { rc[k] = fn(ctx->peers[k]); });
    int r0 = fn(ctx);
    for (auto& t : th) t.join();
    if (r0) return r0;
    for (size_t k = 0; k < rc.size(); ++k)
        if (rc[k]) { ctx->err = "device " + std::to_string(ctx->peers[k]->device) + ": " + ctx->peers[k]->err; return rc[k]; }
    return MRT_OK;
}

extern "C" {

int mrt_abi_version(void) { return MRT_ABI_VERSION; }

const char* mrt_last_error(mrt_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mrt_context_create(int device, void* stream, mrt_context** out) {
    if (!out) { g_create_error = "out is NULL"; return MRT_E_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this backend has no CPU fallback)";
        return MRT_E_CUDA;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return MRT_E_INVALID; }
    mrt_context* ctx = new mrt_context();
    ctx->device = device;
    auto bail = [&](const char* what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        delete ctx;
        return MRT_E_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (prop.major != 10) {
        g_create_error = "this library contains sm_100a code only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
        delete ctx;
        return MRT_E_UNSUPPORTED;
    }
    ctx->n_sms = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = static_cast<cudaStream_t>(stream);
    } else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
        ctx->own_stream = true;
    }
    if ((e = cudaMalloc(&ctx->d_q, sizeof(QueueState))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMallocHost(&ctx->h_q, 3 * sizeof(QueueState))) != cudaSuccess) return bail("cudaMallocHost", e);
    for (int i = 0; i < 2; ++i)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_status[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev_begin)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev_end)) != cudaSuccess) return bail("cudaEventCreate", e);
    // persistent grids: SM count x resident blocks per SM
    int occ = 0;
    auto extend_grid = [&](auto kernel) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kExtendThreads, 0);
        return ctx->n_sms * std::max(occ, 1);
    };
    ctx->grid_extend[0][0] = extend_grid(k_extend<false, false, false>);
    ctx->grid_extend[0][1] = extend_grid(k_extend<false, false, true>);
    ctx->grid_extend[0][2] = extend_grid(k_extend<false, true, true>);
    ctx->grid_extend[1][0] = extend_grid(k_extend<true, false, false>);
    ctx->grid_extend[1][1] = extend_grid(k_extend<true, false, true>);
    ctx->grid_extend[1][2] = extend_grid(k_extend<true, true, true>);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_shade<false>, MRT_SHADE_THREADS, 0);
    ctx->grid_shade = ctx->n_sms * std::max(occ, 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_shade<true>, MRT_SHADE_THREADS, 0);
    ctx->grid_shade_full = ctx->n_sms * std::max(occ, 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_generate, 256, 0);
    ctx->grid_generate = ctx->n_sms * std::max(occ, 1);
    cudaDeviceSetLimit(cudaLimitStackSize, 2048);
    *out = ctx;
    return MRT_OK;
}

void mrt_context_destroy(mrt_context* ctx) {
    if (!ctx) return;
    for (mrt_context* p : ctx->peers) mrt_context_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) {
        std::string why;
        if (NcclApi* api = nccl_api(why)) api->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
        ctx->comm = nullptr;
    }
    cudaFree(ctx->d_count);
    if (ctx->h_count) cudaFreeHost(ctx->h_count);
    cudaFree(ctx->d_scene_hdr);
    if (ctx->h_scene_hdr) cudaFreeHost(ctx->h_scene_hdr);
    free_scene(ctx);
    release_scene_arena(ctx);
    free_pool(ctx);
    free_image(ctx);
    cudaFree(ctx->d_q);
    cudaFreeHost(ctx->h_q);
    for (int i = 0; i < 2; ++i)
        if (ctx->ev_status[i]) cudaEventDestroy(ctx->ev_status[i]);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

static bool ref_ok(const mrt_scene_desc* s, uint32_t ref, bool allow_node) {
    uint32_t idx = MRT_REF_INDEX(ref);
    switch (MRT_REF_KIND(ref)) {
        case MRT_PRIM_NODE: return allow_node && idx < s->n_nodes;
        case MRT_PRIM_SPHERE: return idx < s->n_spheres;
        case MRT_PRIM_TRIANGLE: return idx < s->n_tris;
        case MRT_PRIM_INSTANCE: return idx < s->n_instances;
        case MRT_PRIM_VOLUME: return idx < s->n_volumes;
        default: return false;
    }
}

static void leaf_bounds(const mrt_scene_desc* s, uint32_t ref, float lo[3], float hi[3]) {
    uint32_t idx = MRT_REF_INDEX(ref);
    switch (MRT_REF_KIND(ref)) {
        case MRT_PRIM_NODE:
            for (int k = 0; k < 3; ++k) { lo[k] = s->nodes[idx].bmin[k]; hi[k] = s->nodes[idx].bmax[k]; }
            break;
        case MRT_PRIM_SPHERE: {
            float r = std::fabs(s->spheres[idx].radius);
            for (int k = 0; k < 3; ++k) { lo[k] = s->spheres[idx].center[k] - r; hi[k] = s->spheres[idx].center[k] + r; }
            break;
        }
        case MRT_PRIM_TRIANGLE: {
            const float* v = s->tri_verts + 9 * (size_t)idx;
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::fmin(std::fmin(v[k], v[3 + k]), v[6 + k]);
                hi[k] = std::fmax(std::fmax(v[k], v[3 + k]), v[6 + k]);
            }
            break;
        }
        case MRT_PRIM_INSTANCE:
            for (int k = 0; k < 3; ++k) { lo[k] = s->instances[idx].bmin[k]; hi[k] = s->instances[idx].bmax[k]; }
            break;
        case MRT_PRIM_VOLUME: leaf_bounds(s, s->volumes[idx].target, lo, hi); break;
    }
}

// depth of the subtree under `ref` counted in inner nodes, iteratively; -1 on a malformed tree
static int subtree_depth(const mrt_scene_desc* s, uint32_t root, bool tlas, std::string& why) {
    if (MRT_REF_KIND(root) != MRT_PRIM_NODE) return 0;
    struct Item { uint32_t ref; int depth; };
    std::vector<Item> st;
    st.push_back({root, 1});
    int best = 0;
    uint64_t visited = 0;
    while (!st.empty()) {
        Item it = st.back();
        st.pop_back();
        if (++visited > 2 * s->n_nodes + 2) { why = "node graph is not a tree"; return -1; }
        best = std::max(best, it.depth);
        const mrt_node& n = s->nodes[MRT_REF_INDEX(it.ref)];
        const uint32_t ch[2] = {n.left, n.right};
        for (int k = 0; k < 2; ++k) {
            if (ch[k] == MRT_REF_NONE) {
                if (k == 0) { why = "node without a left child"; return -1; }
                continue;
            }
            if (!ref_ok(s, ch[k], true)) { why = "node child reference out of range"; return -1; }
            uint32_t kind = MRT_REF_KIND(ch[k]);
            if (kind == MRT_PRIM_NODE) st.push_back({ch[k], it.depth + 1});
            else if (tlas && kind == MRT_PRIM_TRIANGLE) { why = "bare Triangle in the world list is not supported (wrap it in a Model)"; return -2; }
            else if (!tlas && kind != MRT_PRIM_TRIANGLE) { why = "a BLAS may only contain triangles"; return -1; }
        }
    }
    return best;
}

constexpr int kRetryOnHost = 1;  // internal: the GPU-built tree is too deep for the traversal stack
constexpr uint32_t kDeviceBuildMin = 16384, kDeviceBuildMaxMeshes = 1024;
constexpr size_t kDeviceTlasMin = 1u << 20;  // an LBVH over instances traverses markedly worse than the SAH tree (Menger sponge: +37 % k_extend), so
                                             // the TLAS goes to the GPU only where the host build would take seconds

// Everything mrt_scene_upload checks before it touches the device: 0, or the MRT_E_* code with `why` set. Needs no context and no
// GPU (mrt_scene_validate exposes it). blas_depth / tlas_depth: depth of the caller's trees in inner nodes (keep-topology uploads).
static int validate_scene(const mrt_scene_desc* s, std::string& why, int& blas_depth, int& tlas_depth) {
    blas_depth = tlas_depth = 0;
    if (!s) { why = "scene is NULL"; return MRT_E_INVALID; }
    if (s->abi_version != MRT_ABI_VERSION) { why = "mrt_scene_desc.abi_version mismatch"; return MRT_E_INVALID; }
    if ((s->n_roots && !s->roots) || (s->n_nodes && !s->nodes) || (s->n_spheres && !s->spheres) || (s->n_tris && (!s->tri_verts || !s->tri_shading)) ||
        (s->n_blas && !s->blas) || (s->n_instances && !s->instances) || (s->n_volumes && !s->volumes) || (s->n_materials && !s->materials) ||
        (s->n_surfaces && !s->surfaces) || (s->n_textures && !s->textures) || (s->n_texels && !s->texels)) {
        why = "an array of the scene is NULL while its count is not 0";
        return MRT_E_INVALID;
    }
    if (s->background.kind < MRT_BG_SOLID || s->background.kind > MRT_BG_CUBEMAP) { why = "unknown background kind"; return MRT_E_INVALID; }
    if (s->n_nodes >= (1ull << 29) || s->n_tris >= (1ull << 29) || s->n_spheres >= (1ull << 29) || s->n_instances >= (1ull << 29))
        { why = "array too large for 29-bit primitive references"; return MRT_E_INVALID; }
    for (uint32_t i = 0; i < s->n_roots; ++i) {
        if (!ref_ok(s, s->roots[i], true)) { why = "root reference out of range"; return MRT_E_INVALID; }
        if (MRT_REF_KIND(s->roots[i]) == MRT_PRIM_TRIANGLE) { why = "bare Triangle in the world list is not supported (wrap it in a Model)"; return MRT_E_UNSUPPORTED; }
    }
    auto mat_ok = [&](int32_t m, bool none_ok) { return (none_ok && m == -1) || (m >= 0 && (uint64_t)m < s->n_materials); };
    auto surf_ok = [&](int32_t v) { return v >= 0 && (uint64_t)v < s->n_surfaces; };
    for (uint64_t i = 0; i < s->n_surfaces; ++i) {
        const mrt_surface& u = s->surfaces[i];
        bool ok = true;
        if (u.kind == MRT_SURF_TEXTURE) ok = u.a >= 0 && (uint64_t)u.a < s->n_textures;
        else if (u.kind == MRT_SURF_YCBCR) ok = u.a >= 0 && (uint64_t)u.a < s->n_textures && u.b >= 0 && (uint64_t)u.b < s->n_textures;
        else if (u.kind == MRT_SURF_BLEND) ok = surf_ok(u.a) && surf_ok(u.b) && (uint64_t)u.a < i && (uint64_t)u.b < i && u.mode >= 0 && u.mode <= 3;
        else if (u.kind == MRT_SURF_FALLBACK) ok = surf_ok(u.a) && (uint64_t)u.a < i;
        else if (u.kind != MRT_SURF_SOLID) ok = false;
        if (!ok) { why = "malformed surface table entry " + std::to_string(i); return MRT_E_INVALID; }
    }
    for (uint64_t i = 0; i < s->n_textures; ++i) {
        const mrt_texture& t = s->textures[i];
        if (t.width == 0 || t.height == 0 || t.texel_offset > s->n_texels || (uint64_t)t.width * t.height > s->n_texels - t.texel_offset)  // no wrap-around for huge offsets
            { why = "texture outside the texel array"; return MRT_E_INVALID; }
        if (t.wrap != MRT_WRAP_REPEAT && t.wrap != MRT_WRAP_CLAMP) { why = "Mirror wrapping is not implemented (texture.rs:280)"; return MRT_E_UNSUPPORTED; }
    }
    for (uint64_t i = 0; i < s->n_materials; ++i) {
        const mrt_material& m = s->materials[i];
        bool ok = m.kind >= 0 && m.kind < MRT_MAT_KINDS;
        if (ok && (m.kind == MRT_MAT_LAMBERTIAN || m.kind == MRT_MAT_METAL || m.kind == MRT_MAT_SPECULAR)) ok = surf_ok(m.surface);
        if (ok && m.kind == MRT_MAT_MIX) ok = mat_ok(m.left, false) && mat_ok(m.right, false) && (uint64_t)m.left < i && (uint64_t)m.right < i;
        if (ok && m.kind == MRT_MAT_EVE) {
            int32_t pal;
            std::memcpy(&pal, &m.p[0], 4);
            ok = surf_ok(m.surface) && surf_ok(m.left) && surf_ok(m.right) && pal >= 0 && (uint64_t)pal + 4 <= s->n_surfaces;
            for (int k = 0; ok && k < 4; ++k) ok = s->surfaces[pal + k].kind == MRT_SURF_SOLID;
        }
        if (!ok) { why = "malformed material table entry " + std::to_string(i); return MRT_E_INVALID; }
    }
    for (uint64_t i = 0; i < s->n_spheres; ++i)
        if (!mat_ok(s->spheres[i].material, false)) { why = "sphere material out of range"; return MRT_E_INVALID; }
    {
        std::atomic<bool> bad{false};
        parallel_for((size_t)s->n_tris, [&](size_t a, size_t b) {
            for (size_t i = a; i < b; ++i)
                if (!mat_ok(s->tri_shading[i].material, false)) bad = true;
        });
        if (bad) { why = "triangle material out of range"; return MRT_E_INVALID; }
    }
    for (uint64_t i = 0; i < s->n_volumes; ++i) {
        const mrt_volume& v = s->volumes[i];
        if (!ref_ok(s, v.target, false)) { why = "volume target out of range"; return MRT_E_INVALID; }
        if (MRT_REF_KIND(v.target) != MRT_PRIM_SPHERE && MRT_REF_KIND(v.target) != MRT_PRIM_INSTANCE)
            { why = "Volume targets: Sphere, Model and Instance (geom.rs:595 is generic over Intersect)"; return MRT_E_UNSUPPORTED; }
        if (!mat_ok(v.material, false)) { why = "volume material out of range"; return MRT_E_INVALID; }
    }
    if (s->background.kind == MRT_BG_SKYSPHERE && !surf_ok(s->background.surface[0])) { why = "background surface out of range"; return MRT_E_INVALID; }
    if (s->background.kind == MRT_BG_CUBEMAP)
        for (int k = 0; k < 6; ++k)
            if (!surf_ok(s->background.surface[k])) { why = "background surface out of range"; return MRT_E_INVALID; }
    blas_depth = 0;
    const bool keep = (s->flags & MRT_SCENE_KEEP_TOPOLOGY) != 0;
    for (uint64_t i = 0; i < s->n_blas; ++i) {
        const mrt_blas& b = s->blas[i];
        if ((uint64_t)b.first_tri + b.n_tris > s->n_tris) { why = "malformed BLAS table entry"; return MRT_E_INVALID; }
        if (b.n_tris == 0) { why = "BLAS without triangles (BvhNode::new does not terminate on an empty list, geom.rs:130-144)"; return MRT_E_INVALID; }
        if (b.root == MRT_REF_NONE) {  // a mesh without a caller tree (mrth_defer_mesh_bvh)
            if (keep) { why = "MRT_SCENE_KEEP_TOPOLOGY needs the nodes of every BLAS"; return MRT_E_INVALID; }
            continue;
        }
        if (MRT_REF_KIND(b.root) != MRT_PRIM_NODE || MRT_REF_INDEX(b.root) >= s->n_nodes) { why = "malformed BLAS table entry"; return MRT_E_INVALID; }
        if (!keep) continue;  // the rebuild reads the BLAS's triangle range only, never the caller's nodes
        int d = subtree_depth(s, b.root, false, why);
        if (d < 0) return MRT_E_INVALID;
        blas_depth = std::max(blas_depth, d);
    }
    for (uint64_t i = 0; i < s->n_instances; ++i) {
        if (s->instances[i].blas >= s->n_blas) { why = "instance BLAS out of range"; return MRT_E_INVALID; }
        if (!mat_ok(s->instances[i].material, true)) { why = "instance material out of range"; return MRT_E_INVALID; }
    }
    tlas_depth = 0;
    for (uint32_t i = 0; i < s->n_roots; ++i) {
        int d = subtree_depth(s, s->roots[i], true, why);
        if (d == -2) return MRT_E_UNSUPPORTED;
        if (d < 0) return MRT_E_INVALID;
        tlas_depth = std::max(tlas_depth, d);
    }
    if (keep) {  // every node is converted below, reachable from a root or not: all of them must name children that exist
        for (uint64_t i = 0; i < s->n_nodes; ++i) {
            const mrt_node& n = s->nodes[i];
            if (n.left == MRT_REF_NONE || !ref_ok(s, n.left, true) || (n.right != MRT_REF_NONE && !ref_ok(s, n.right, true)))
                { why = "node " + std::to_string(i) + ": child reference out of range"; return MRT_E_INVALID; }
        }
    }
    if (keep && tlas_depth + blas_depth + 2 > kStackSize)
        { why = "BVH too deep for the " + std::to_string(kStackSize) + "-entry traversal stack"; return MRT_E_UNSUPPORTED; }
    if (s->n_tris > kTriIndexMask) { why = "more than 2^27 triangles"; return MRT_E_UNSUPPORTED; }
    return MRT_OK;
}

int mrt_scene_validate(const mrt_scene_desc* scene, char* why_out, size_t why_bytes) {
    std::string why;
    int blas_depth, tlas_depth;
    const int rc = validate_scene(scene, why, blas_depth, tlas_depth);
    if (why_out && why_bytes) {
        const size_t n = std::min(why.size(), why_bytes - 1);
        std::memcpy(why_out, why.data(), n);
        why_out[n] = 0;
    }
    return rc;
}

static int scene_upload_impl(mrt_context* ctx, const mrt_scene_desc* s, bool allow_device);

static int comm_scene_broadcast(mrt_context* ctx, int root_rc);

int mrt_scene_upload(mrt_context* ctx, const mrt_scene_desc* s) {
    if (!ctx) return MRT_E_INVALID;
    // One process per GPU: the scene is the same on every rank (it is replicated, never sharded), so only the root validates, builds
    // and uploads it; the finished device arrays then travel over NVLink (one grouped ncclBroadcast) instead of N times over PCIe
    // from N host copies. COLLECTIVE like mrt_render on such a context; the other ranks' `s` is not read and may be NULL.
    const bool collective = ctx->comm_size > 1 && ctx->peers.empty() && ctx->comm_scene;
    int rc = MRT_OK;
    if (!collective || ctx->comm_rank == 0) {
        rc = scene_upload_impl(ctx, s, ctx->opt_device_build != 0);
        if (rc == kRetryOnHost) rc = scene_upload_impl(ctx, s, false);
    }
    if (collective) return comm_scene_broadcast(ctx, rc);
    if (rc) return rc;
    for (mrt_context* p : ctx->peers)
        if ((rc = replicate_scene(ctx, p))) return fail(ctx, rc, "device " + std::to_string(p->device) + ": " + p->err);
    return MRT_OK;
}

static int scene_upload_impl(mrt_context* ctx, const mrt_scene_desc* s, bool allow_device) {
    MRT_CUDA(cudaSetDevice(ctx->device));
    // phase timing of the upload on stderr when MRT_UPLOAD_TIMING is set (measurement aid)
    const bool timing = std::getenv("MRT_UPLOAD_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "mrt_scene_upload: %-10s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    // ---- validate ------------------------------------------------------------------------------------------
    std::string why;
    int blas_depth = 0, tlas_depth = 0;
    if (int vrc = validate_scene(s, why, blas_depth, tlas_depth)) return fail(ctx, vrc, why);
    const bool keep = (s->flags & MRT_SCENE_KEEP_TOPOLOGY) != 0;
    lap("validate");
    // ---- acceleration structure: the caller's topology re-laid out, or (default) a SAH rebuild -----------------------------
    cudaStreamSynchronize(ctx->stream);
    free_scene(ctx);
    DScene d{};
    const float inf = INFINITY;
    std::vector<DNode> nodes;
    std::vector<uint32_t> tri_map(s->n_tris);      // device triangle index -> caller's triangle index
    std::vector<uint32_t> blas_root(s->n_blas);
    std::vector<uint32_t> roots(s->roots, s->roots + s->n_roots);
    auto set_child = [&](DNode& o, int which, uint32_t ref, const float lo[3], const float hi[3]) {
        if (which == 0) {
            o.xy0 = make_float4(lo[0], hi[0], lo[1], hi[1]);
            o.z01.x = lo[2]; o.z01.y = hi[2];
            o.child0 = ref;
        } else {
            o.xy1 = make_float4(lo[0], hi[0], lo[1], hi[1]);
            o.z01.z = lo[2]; o.z01.w = hi[2];
            o.child1 = ref;
        }
    };
    const float empty_lo[3] = {inf, inf, inf}, empty_hi[3] = {-inf, -inf, -inf};
    int max_tlas_depth = tlas_depth, max_blas_depth = blas_depth;
    uint32_t device_root = kNone;  // stays kNone for an empty world
    std::vector<uint32_t> device_meshes;  // BLAS indices the GPU builds
    bool tlas_on_device = false;          // the world has so many objects that the GPU builds the TLAS too
    std::vector<mrt_build::Prim> tlas_prims;
    if (keep) {
        nodes.resize(s->n_nodes);
        for (uint64_t i = 0; i < s->n_nodes; ++i) {
            const mrt_node& n = s->nodes[i];
            float l0[3], h0[3], l1[3] = {inf, inf, inf}, h1[3] = {-inf, -inf, -inf};
            leaf_bounds(s, n.left, l0, h0);
            if (n.right != MRT_REF_NONE) leaf_bounds(s, n.right, l1, h1);
            DNode& o = nodes[i];
            o.pad0 = o.pad1 = 0;
            set_child(o, 0, n.left, l0, h0);   // a TRIANGLE ref here is a one-triangle leaf (count bits 0, index < 2^27)
            set_child(o, 1, n.right, l1, h1);
        }
        for (uint64_t i = 0; i < s->n_tris; ++i) tri_map[i] = (uint32_t)i;
        for (uint64_t i = 0; i < s->n_blas; ++i) blas_root[i] = s->blas[i].root;
        if (s->n_roots == 1) {
            device_root = roots[0];
            if (MRT_REF_KIND(device_root) != MRT_PRIM_NODE) {  // a lone primitive still needs a node to hold its box
                float lo[3], hi[3];
                leaf_bounds(s, device_root, lo, hi);
                DNode o{};
                set_child(o, 0, device_root, lo, hi);
                set_child(o, 1, kNone, empty_lo, empty_hi);
                nodes.push_back(o);
                device_root = MRT_REF(MRT_PRIM_NODE, (uint32_t)nodes.size() - 1);
            }
        } else if (s->n_roots > 1) {
            // a plain object list (World::intersect's loop, world.rs:135-140) has no topology to keep: a balanced tree over the
            // entries in list order visits the same objects; each entry keeps its own subtree as the caller built it
            struct Item { uint32_t ref; float lo[3], hi[3]; };
            std::vector<Item> level(s->n_roots);
            for (uint32_t i = 0; i < s->n_roots; ++i) { level[i].ref = roots[i]; leaf_bounds(s, roots[i], level[i].lo, level[i].hi); }
            int extra = 0;
            while (level.size() > 1) {
                std::vector<Item> next;
                for (size_t i = 0; i < level.size(); i += 2) {
                    DNode o{};
                    Item it;
                    set_child(o, 0, level[i].ref, level[i].lo, level[i].hi);
                    if (i + 1 < level.size()) {
                        set_child(o, 1, level[i + 1].ref, level[i + 1].lo, level[i + 1].hi);
                        for (int k = 0; k < 3; ++k) { it.lo[k] = std::fmin(level[i].lo[k], level[i + 1].lo[k]); it.hi[k] = std::fmax(level[i].hi[k], level[i + 1].hi[k]); }
                    } else {
                        set_child(o, 1, kNone, empty_lo, empty_hi);
                        for (int k = 0; k < 3; ++k) { it.lo[k] = level[i].lo[k]; it.hi[k] = level[i].hi[k]; }
                    }
                    nodes.push_back(o);
                    it.ref = MRT_REF(MRT_PRIM_NODE, (uint32_t)nodes.size() - 1);
                    next.push_back(it);
                }
                level.swap(next);
                ++extra;
            }
            device_root = level[0].ref;
            max_tlas_depth += extra;
            if (max_tlas_depth + max_blas_depth + 2 > kStackSize) return fail(ctx, MRT_E_UNSUPPORTED, "BVH too deep for the traversal stack");
        }
    } else {
        // emit a built tree as DNodes in depth-first order (a node's left subtree follows it directly); returns the reference
        // of the root. leaf_ref(first, count) names a leaf.
        auto emit_tree = [&](const mrt_build::Tree& t, auto leaf_ref) -> uint32_t {
            const mrt_build::Node* bn = t.nodes.get();
            if (bn[t.root].left < 0) {  // the whole set is one leaf: still needs a node to hold its box
                DNode o{};
                set_child(o, 0, leaf_ref(bn[t.root].first, bn[t.root].count), bn[t.root].lo, bn[t.root].hi);
                set_child(o, 1, kNone, empty_lo, empty_hi);
                nodes.push_back(o);
                return MRT_REF(MRT_PRIM_NODE, (uint32_t)nodes.size() - 1);
            }
            struct Item { int32_t node; uint32_t parent; int which; };
            std::vector<Item> st{{t.root, kNone, 0}};
            uint32_t root_ref = kNone;
            while (!st.empty()) {
                const Item it = st.back();
                st.pop_back();
                const mrt_build::Node& n = bn[it.node];
                uint32_t ref;
                if (n.left < 0) {
                    ref = leaf_ref(n.first, n.count);
                } else {
                    const uint32_t me = (uint32_t)nodes.size();
                    DNode o{};
                    set_child(o, 0, kNone, bn[n.left].lo, bn[n.left].hi);  // child references are patched in when the children are emitted
                    set_child(o, 1, kNone, bn[n.right].lo, bn[n.right].hi);
                    nodes.push_back(o);
                    ref = MRT_REF(MRT_PRIM_NODE, me);
                    st.push_back({n.right, me, 1});
                    st.push_back({n.left, me, 0});
                }
                if (it.parent == kNone) root_ref = ref;
                else if (it.which == 0) nodes[it.parent].child0 = ref;
                else nodes[it.parent].child1 = ref;
            }
            return root_ref;
        };
        max_blas_depth = 0;
        nodes.reserve((size_t)s->n_tris + 2 * (size_t)s->n_instances + 2 * (size_t)s->n_spheres + 64);
        for (uint64_t bi = 0; bi < s->n_blas; ++bi) {  // BLAS: SAH, leaves of up to 4 triangles
            const mrt_blas& bl = s->blas[bi];
            if (allow_device && bl.n_tris >= kDeviceBuildMin && device_meshes.size() < kDeviceBuildMaxMeshes) {
                device_meshes.push_back((uint32_t)bi);  // built on the GPU after the upload (mrt_lbvh.cuh); its root is set below
                continue;
            }
            std::vector<mrt_build::Prim> prims(bl.n_tris);
            parallel_for((size_t)bl.n_tris, [&](size_t a, size_t b) {
                for (size_t i = a; i < b; ++i) {
                    prims[i].ref = bl.first_tri + (uint32_t)i;
                    leaf_bounds(s, MRT_REF(MRT_PRIM_TRIANGLE, bl.first_tri + (uint32_t)i), prims[i].lo, prims[i].hi);
                }
            });
            mrt_build::Tree t = mrt_build::build_sah(prims, (int)ctx->opt_leaf_tris, 40, (float)ctx->opt_tri_cost * 0.01f);
            max_blas_depth = std::max(max_blas_depth, std::max(t.depth, 1));
            for (uint32_t i = 0; i < bl.n_tris; ++i) tri_map[bl.first_tri + i] = prims[i].ref;  // device order = leaf order, inside the BLAS's own range
            const uint32_t base = bl.first_tri;
            blas_root[bi] = emit_tree(t, [base](uint32_t first, uint32_t count) { return MRT_REF(MRT_PRIM_TRIANGLE, base + first) | ((count - 1u) << 27); });
        }
        // TLAS: one SAH tree over every object of the world, one object per leaf -- whether the caller passed World::build_bvh's
        // tree (world.rs:117-122) or the plain object list that World::intersect loops over (world.rs:135-140).
        max_tlas_depth = 0;
        std::vector<mrt_build::Prim> prims;
        for (uint32_t ri = 0; ri < s->n_roots; ++ri) {
            std::vector<uint32_t> st{roots[ri]};
            while (!st.empty()) {
                uint32_t ref = st.back();
                st.pop_back();
                if (MRT_REF_KIND(ref) == MRT_PRIM_NODE) {
                    const mrt_node& n = s->nodes[MRT_REF_INDEX(ref)];
                    if (n.right != MRT_REF_NONE) st.push_back(n.right);
                    st.push_back(n.left);
                } else {
                    mrt_build::Prim p;
                    p.ref = ref;
                    leaf_bounds(s, ref, p.lo, p.hi);
                    prims.push_back(p);
                }
            }
        }
        if (allow_device && prims.size() >= (ctx->opt_device_build >= 2 ? (size_t)kDeviceBuildMin : kDeviceTlasMin)) {
            tlas_on_device = true;  // built on the GPU after the upload, from these boxes (its root is set below)
            tlas_prims.swap(prims);
        } else if (!prims.empty()) {
            mrt_build::Tree t = mrt_build::build_sah(prims, 1, 28, 4.0f);
            max_tlas_depth = std::max(t.depth, 1);
            const mrt_build::Prim* pp = prims.data();
            device_root = emit_tree(t, [pp](uint32_t first, uint32_t) { return pp[first].ref; });
        }
        if (max_tlas_depth + max_blas_depth + 2 > kStackSize) return fail(ctx, MRT_E_UNSUPPORTED, "rebuilt BVH too deep for the traversal stack");
    }
    // GPU-built meshes: their nodes follow the host-built ones in the node array, root first
    std::vector<uint32_t> device_node_base(device_meshes.size());
    size_t n_nodes_total = nodes.size();
    for (size_t k = 0; k < device_meshes.size(); ++k) {
        device_node_base[k] = (uint32_t)n_nodes_total;
        blas_root[device_meshes[k]] = MRT_REF(MRT_PRIM_NODE, (uint32_t)n_nodes_total);
        n_nodes_total += s->blas[device_meshes[k]].n_tris - 1;
    }
    uint32_t tlas_node_base = 0;
    if (tlas_on_device) {
        tlas_node_base = (uint32_t)n_nodes_total;
        device_root = MRT_REF(MRT_PRIM_NODE, tlas_node_base);
        n_nodes_total += tlas_prims.size() - 1;
    }
    if (n_nodes_total >= (1ull << 29)) return fail(ctx, MRT_E_INVALID, "too many BVH nodes for 29-bit references");
    lap("build");
    // which triangles can fail Material::alpha_test (geom.rs:567-571): UV'd, and their own material's surface can return alpha 0
    std::vector<int8_t> surf_alpha(s->n_surfaces, -1), mat_alpha(s->n_materials, -1);
    std::function<bool(int32_t)> surface_can_be_transparent = [&](int32_t si) -> bool {
        if (surf_alpha[(size_t)si] >= 0) return surf_alpha[(size_t)si] != 0;
        const mrt_surface& u = s->surfaces[si];
        bool r = true;
        if (u.kind == MRT_SURF_SOLID) r = u.color[3] == 0.0f;
        else if (u.kind == MRT_SURF_YCBCR) r = false;  // alpha is always 1 (texture.rs:246)
        else if (u.kind == MRT_SURF_TEXTURE) {
            const mrt_texture& t = s->textures[u.a];
            r = false;
            for (uint64_t k = 0; k < (uint64_t)t.width * t.height && !r; ++k) r = s->texels[4 * (t.texel_offset + k) + 3] == 0.0f;
        }
        surf_alpha[(size_t)si] = r ? 1 : 0;
        return r;
    };
    std::function<bool(int32_t)> material_can_fail_alpha = [&](int32_t mi) -> bool {
        if (mat_alpha[(size_t)mi] >= 0) return mat_alpha[(size_t)mi] != 0;
        const mrt_material& m = s->materials[mi];
        bool r = false;
        if (m.kind == MRT_MAT_LAMBERTIAN || m.kind == MRT_MAT_METAL || m.kind == MRT_MAT_SPECULAR) r = surface_can_be_transparent(m.surface);
        else if (m.kind == MRT_MAT_MIX) r = material_can_fail_alpha(m.left) || material_can_fail_alpha(m.right);
        mat_alpha[(size_t)mi] = r ? 1 : 0;
        return r;
    };
    std::vector<float4> spheres(s->n_spheres);
    std::vector<DSphereAux> saux(s->n_spheres);
    for (uint64_t i = 0; i < s->n_spheres; ++i) {
        spheres[i] = make_float4(s->spheres[i].center[0], s->spheres[i].center[1], s->spheres[i].center[2], s->spheres[i].radius);
        saux[i] = DSphereAux{s->spheres[i].material, s->spheres[i].object_id};
    }
    for (uint64_t i = 0; i < s->n_materials; ++i) material_can_fail_alpha((int32_t)i);  // fills mat_alpha: the loop below only reads it
    std::vector<uint8_t> on_device(s->n_blas, 0);
    for (uint32_t bi : device_meshes) on_device[bi] = 1;
    std::unique_ptr<DTriVerts[]> tv(new DTriVerts[std::max<size_t>((size_t)s->n_tris, 1)]);  // untouched where the GPU fills in
    std::atomic<uint32_t> any_alpha_acc{0};
    parallel_for((size_t)s->n_tris, [&](size_t first, size_t last) {  // alpha-tested triangles anywhere in the scene?
        uint32_t any = 0;
        for (size_t i = first; i < last; ++i) {
            const mrt_tri_shading& sh = s->tri_shading[i];
            if ((sh.flags & MRT_TRI_HAS_UV) && mat_alpha[(size_t)sh.material] > 0) any = kTriAlphaFlag;
        }
        any_alpha_acc |= any;
    });
    const uint32_t any_alpha = any_alpha_acc.load();
    for (uint64_t bi = 0; bi < s->n_blas; ++bi) {
        if (on_device[bi]) continue;
        const size_t base = s->blas[bi].first_tri;
        parallel_for((size_t)s->blas[bi].n_tris, [&](size_t first, size_t last) {
            for (size_t i = base + first; i < base + last; ++i) {
                const uint32_t orig = tri_map[i];
                const float* v = s->tri_verts + 9 * (size_t)orig;
                const mrt_tri_shading& sh = s->tri_shading[orig];
                uint32_t flags = ((sh.flags & MRT_TRI_HAS_UV) && mat_alpha[(size_t)sh.material] > 0) ? kTriAlphaFlag : 0u;
                float fw;
                std::memcpy(&fw, &flags, 4);
                tv[i].a = make_float4(v[0], v[1], v[2], fw);
                tv[i].b = make_float4(v[3], v[4], v[5], 0.0f);
                tv[i].c = make_float4(v[6], v[7], v[8], 0.0f);
            }
        });
    }
    std::vector<DInstance> inst(s->n_instances);
    for (uint64_t i = 0; i < s->n_instances; ++i) {
        const mrt_instance& in = s->instances[i];
        auto pack = [](const float* m, float4& a, float4& b, float4& c) {  // columns c0.xyz c1.xyz c2.xyz c3.xyz
            a = make_float4(m[0], m[1], m[2], m[4]);
            b = make_float4(m[5], m[6], m[8], m[9]);
            c = make_float4(m[10], m[12], m[13], m[14]);
        };
        DInstance& o = inst[i];
        std::memset(&o, 0, sizeof o);
        pack(in.inv_transform, o.inv0, o.inv1, o.inv2);
        pack(in.transform, o.fwd0, o.fwd1, o.fwd2);
        o.root = blas_root[in.blas];
        o.material = in.material;
        o.flags = in.flags;
        o.object_id = in.object_id;
        o.pad[0] = in.blas;
    }
    lap("pack");
    int rc;
    DNode* d_nodes = nullptr;
    DTriVerts* d_tv = nullptr;
    uint32_t* d_tri_map = nullptr;
    if ((rc = arena_alloc(ctx, n_nodes_total * sizeof(DNode), reinterpret_cast<void**>(&d_nodes)))) return rc;
    if (!nodes.empty() && (rc = stage_copy(ctx, d_nodes, nodes.data(), nodes.size() * sizeof(DNode)))) return rc;
    if ((rc = upload(ctx, spheres.data(), spheres.size(), &d.spheres))) return rc;
    if ((rc = upload(ctx, saux.data(), saux.size(), &d.sphere_aux))) return rc;
    if ((rc = arena_alloc(ctx, (size_t)s->n_tris * sizeof(DTriVerts), reinterpret_cast<void**>(&d_tv)))) return rc;
    if ((rc = arena_alloc(ctx, (size_t)s->n_tris * sizeof(uint32_t), reinterpret_cast<void**>(&d_tri_map)))) return rc;
    for (uint64_t bi = 0; bi < s->n_blas; ++bi) {  // host-built meshes: their triangle ranges; the GPU builder writes the others
        if (on_device[bi]) continue;
        const size_t base = s->blas[bi].first_tri, n = s->blas[bi].n_tris;
        if ((rc = stage_copy(ctx, d_tv + base, tv.get() + base, n * sizeof(DTriVerts)))) return rc;
        if ((rc = stage_copy(ctx, d_tri_map + base, tri_map.data() + base, n * sizeof(uint32_t)))) return rc;
    }
    d.nodes = d_nodes;
    d.tri_verts = d_tv;
    d.tri_map = d_tri_map;
    if ((rc = upload(ctx, s->tri_shading, (size_t)s->n_tris, &d.tri_shading))) return rc;
    if ((rc = upload(ctx, inst.data(), inst.size(), &d.instances))) return rc;
    if ((rc = upload(ctx, s->blas, (size_t)s->n_blas, &d.blas))) return rc;
    if ((rc = upload(ctx, s->volumes, (size_t)s->n_volumes, &d.volumes))) return rc;
    if ((rc = upload(ctx, s->materials, (size_t)s->n_materials, &d.materials))) return rc;
    if ((rc = upload(ctx, s->surfaces, (size_t)s->n_surfaces, &d.surfaces))) return rc;
    if ((rc = upload(ctx, s->textures, (size_t)s->n_textures, &d.textures))) return rc;
    if ((rc = upload(ctx, reinterpret_cast<const float4*>(s->texels), (size_t)s->n_texels, &d.texels))) return rc;
    d.root = device_root;
    d.n_volumes = (uint32_t)s->n_volumes;
    d.has_alpha = any_alpha;
    for (uint64_t i = 0; i < s->n_volumes; ++i)
        if (MRT_REF_KIND(s->volumes[i].target) == MRT_PRIM_INSTANCE) d.has_alpha = 1;  // Volume over a mesh: only the full kernel variants carry that code
    for (uint64_t i = 0; i < s->n_materials; ++i)
        if (s->materials[i].kind == MRT_MAT_EVE) d.has_alpha = 1;  // likewise EveMaterial (tangent-space normals, texture-driven scatter and emission)
    d.bg = s->background;
    lap("copy");
    if (tlas_on_device) {  // ---- the TLAS of a world with many objects: boxes + references up, tree built on the GPU ---------
        const size_t n = tlas_prims.size(), cubb = lbvh::cub_temp_bytes(n);
        if (!ctx->d_depths) {
            MRT_CUDA(cudaMalloc(&ctx->d_depths, (kDeviceBuildMaxMeshes + 1) * sizeof(int)));
            MRT_CUDA(cudaMallocHost(&ctx->h_depths, (kDeviceBuildMaxMeshes + 1) * sizeof(int)));
        }
        const size_t refs_bytes = (n * 4 + 255) / 256 * 256;
        if ((rc = grow(ctx, ctx->build_scratch, refs_bytes + lbvh::scratch_bytes(n, cubb)))) return rc;
        std::vector<float4> lo(n), hi(n);
        std::vector<uint32_t> refs(n);
        for (size_t i = 0; i < n; ++i) {
            lo[i] = make_float4(tlas_prims[i].lo[0], tlas_prims[i].lo[1], tlas_prims[i].lo[2], 0.0f);
            hi[i] = make_float4(tlas_prims[i].hi[0], tlas_prims[i].hi[1], tlas_prims[i].hi[2], 0.0f);
            refs[i] = tlas_prims[i].ref;
        }
        uint32_t* d_refs = static_cast<uint32_t*>(ctx->build_scratch.p);
        lbvh::Scratch sc = lbvh::carve(static_cast<char*>(ctx->build_scratch.p) + refs_bytes, n, cubb);
        if ((rc = stage_copy(ctx, d_refs, refs.data(), n * 4))) return rc;
        if ((rc = stage_copy(ctx, sc.box_lo, lo.data(), n * 16))) return rc;
        if ((rc = stage_copy(ctx, sc.box_hi, hi.data(), n * 16))) return rc;
        MRT_CUDA(lbvh::build_objects(ctx->stream, sc, (uint32_t)n, d_refs, d_nodes, tlas_node_base, ctx->d_depths + kDeviceBuildMaxMeshes));
        MRT_CUDA(cudaMemcpyAsync(ctx->h_depths + kDeviceBuildMaxMeshes, ctx->d_depths + kDeviceBuildMaxMeshes, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MRT_CUDA(cudaStreamSynchronize(ctx->stream));  // the scratch is reused by the mesh builds below; lo / hi / refs go out of scope
        max_tlas_depth = std::max(ctx->h_depths[kDeviceBuildMaxMeshes], 1);
    }
    if (!device_meshes.empty()) {  // ---- build the big meshes on the GPU, one after the other on the upload stream --------
        const uint8_t* d_mat_alpha = nullptr;
        std::vector<uint8_t> ma(s->n_materials);
        for (uint64_t i = 0; i < s->n_materials; ++i) ma[i] = mat_alpha[i] > 0 ? 1 : 0;
        if ((rc = upload(ctx, ma.data(), ma.size(), &d_mat_alpha))) return rc;
        if (!ctx->d_depths) {
            MRT_CUDA(cudaMalloc(&ctx->d_depths, (kDeviceBuildMaxMeshes + 1) * sizeof(int)));
            MRT_CUDA(cudaMallocHost(&ctx->h_depths, (kDeviceBuildMaxMeshes + 1) * sizeof(int)));
        }
        size_t need = 0;
        for (uint32_t bi : device_meshes) {
            const size_t n = s->blas[bi].n_tris;
            need = std::max(need, (n * 36 + 255) / 256 * 256 + lbvh::scratch_bytes(n, lbvh::cub_temp_bytes(n)));
        }
        if ((rc = grow(ctx, ctx->build_scratch, need))) return rc;
        for (size_t k = 0; k < device_meshes.size(); ++k) {
            const mrt_blas& bl = s->blas[device_meshes[k]];
            const size_t n = bl.n_tris, raw_bytes = (n * 36 + 255) / 256 * 256;
            float* d_raw = static_cast<float*>(ctx->build_scratch.p);
            if ((rc = stage_copy(ctx, d_raw, s->tri_verts + 9 * (size_t)bl.first_tri, n * 36))) return rc;
            lbvh::Scratch sc = lbvh::carve(static_cast<char*>(ctx->build_scratch.p) + raw_bytes, n, lbvh::cub_temp_bytes(n));
            MRT_CUDA(lbvh::build(ctx->stream, sc, d_raw, (uint32_t)n, d.tri_shading, d_mat_alpha, bl.first_tri, d_nodes, device_node_base[k], d_tv, d_tri_map,
                                 ctx->d_depths + k));
        }
        MRT_CUDA(cudaMemcpyAsync(ctx->h_depths, ctx->d_depths, device_meshes.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    MRT_CUDA(cudaStreamSynchronize(ctx->stream));  // the staging vectors above go out of scope
    if (!device_meshes.empty() || tlas_on_device) {
        for (size_t k = 0; k < device_meshes.size(); ++k) max_blas_depth = std::max(max_blas_depth, ctx->h_depths[k]);
        lap("gpu build");
        if (max_tlas_depth + max_blas_depth + 2 > kStackSize) return kRetryOnHost;  // pathological Morton order: let the host's SAH builder do it
    }
    ctx->scene = d;
    ctx->has_scene = true;
    ctx->auto_refill_lanes = max_blas_depth >= kDeepBlas ? kRefillLanesDeep : kRefillLanes;
    ctx->auto_node_burst = max_blas_depth >= kDeepBlas ? kNodeBurstDeep : kNodeBurst;
    ctx->material_kinds = 0;
    for (uint64_t i = 0; i < s->n_materials; ++i) ctx->material_kinds |= 1u << s->materials[i].kind;
    return MRT_OK;
}

int mrt_camera_set(mrt_context* ctx, const mrt_camera* c) {
    if (!ctx) return MRT_E_INVALID;
    if (!c) return fail(ctx, MRT_E_INVALID, "camera is NULL");
    DCamera& d = ctx->cam;
    d.origin = make_float3(c->origin[0], c->origin[1], c->origin[2]);
    d.llc = make_float3(c->lower_left_corner[0], c->lower_left_corner[1], c->lower_left_corner[2]);
    d.horizontal = make_float3(c->horizontal[0], c->horizontal[1], c->horizontal[2]);
    d.vertical = make_float3(c->vertical[0], c->vertical[1], c->vertical[2]);
    d.u = make_float3(c->u[0], c->u[1], c->u[2]);
    d.v = make_float3(c->v[0], c->v[1], c->v[2]);
    d.lens_radius = c->lens_radius;
    ctx->has_camera = true;
    for (mrt_context* p : ctx->peers) { p->cam = d; p->has_camera = true; }
    return MRT_OK;
}

static int check_ready(mrt_context* ctx) {
    if (!ctx) return MRT_E_INVALID;
    if (!ctx->has_scene) return fail(ctx, MRT_E_STATE, "no scene uploaded (mrt_scene_upload)");
    if (!ctx->has_camera) return fail(ctx, MRT_E_STATE, "no camera set (mrt_camera_set)");
    return MRT_OK;
}

int mrt_render_aov(mrt_context* ctx, uint32_t w, uint32_t h, uint64_t seed, float* albedo, float* normal, uint32_t* object_id, uint32_t* tri_id,
                   float* t) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (w < 2 || h < 2 || (uint64_t)w * h > 0x7FFFFFFFull) return fail(ctx, MRT_E_INVALID, "image size out of range");
    MRT_CUDA(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)w * h;
    // the five output planes live in one grow-only buffer of the context (36 bytes per pixel): no cudaMalloc / cudaFree per call
    int rc_buf = grow(ctx, ctx->aov_buf, npix * 36);
    if (rc_buf) return rc_buf;
    char* base = static_cast<char*>(ctx->aov_buf.p);
    float* d_alb = albedo ? reinterpret_cast<float*>(base) : nullptr;
    float* d_nrm = normal ? reinterpret_cast<float*>(base + npix * 12) : nullptr;
    float* d_t = t ? reinterpret_cast<float*>(base + npix * 24) : nullptr;
    uint32_t* d_obj = object_id ? reinterpret_cast<uint32_t*>(base + npix * 28) : nullptr;
    uint32_t* d_tri = tri_id ? reinterpret_cast<uint32_t*>(base + npix * 32) : nullptr;
#define AOV_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return MRT_E_CUDA; } } while (0)
    RenderParams rp{w, h, (uint32_t)npix, 0u, 1u, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), kRefillLanes, kNodeBurst, ctx->opt_finish_paths, 0u, {}};
    const unsigned aov_grid = (unsigned)((npix + 127) / 128);
    if (ctx->scene.has_alpha) k_aov<true, true><<<aov_grid, 128, 0, ctx->stream>>>(ctx->scene, ctx->cam, rp, d_alb, d_nrm, d_obj, d_tri, d_t);
    else if (ctx->scene.n_volumes) k_aov<false, true><<<aov_grid, 128, 0, ctx->stream>>>(ctx->scene, ctx->cam, rp, d_alb, d_nrm, d_obj, d_tri, d_t);
    else k_aov<false, false><<<aov_grid, 128, 0, ctx->stream>>>(ctx->scene, ctx->cam, rp, d_alb, d_nrm, d_obj, d_tri, d_t);
    AOV_TRY(cudaGetLastError());
    if (albedo) AOV_TRY(cudaMemcpyAsync(albedo, d_alb, npix * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (normal) AOV_TRY(cudaMemcpyAsync(normal, d_nrm, npix * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (t) AOV_TRY(cudaMemcpyAsync(t, d_t, npix * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (object_id) AOV_TRY(cudaMemcpyAsync(object_id, d_obj, npix * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (tri_id) AOV_TRY(cudaMemcpyAsync(tri_id, d_tri, npix * 4, cudaMemcpyDeviceToHost, ctx->stream));
    AOV_TRY(cudaStreamSynchronize(ctx->stream));
#undef AOV_TRY
    return MRT_OK;
}

static int accum_reset_one(mrt_context* ctx, uint32_t w, uint32_t h);
int mrt_accum_reset(mrt_context* ctx, uint32_t w, uint32_t h) {
    if (!ctx) return MRT_E_INVALID;
    return for_each_member(ctx, [=](mrt_context* m) { return accum_reset_one(m, w, h); });
}
static int accum_reset_one(mrt_context* ctx, uint32_t w, uint32_t h) {
    if (w < 2 || h < 2 || (uint64_t)w * h > 0x7FFFFFFFull) return fail(ctx, MRT_E_INVALID, "image size out of range");
    MRT_CUDA(cudaSetDevice(ctx->device));
    if (ctx->w != w || ctx->h != h || !ctx->d_accum) {
        cudaStreamSynchronize(ctx->stream);
        free_image(ctx);
        MRT_CUDA(cudaMalloc(&ctx->d_accum, (size_t)w * h * 4 * sizeof(long long)));
        MRT_CUDA(cudaMalloc(&ctx->d_nonfinite, (size_t)w * h * sizeof(uint32_t)));
        MRT_CUDA(cudaMalloc(&ctx->d_sum_rgb, (size_t)w * h * 3 * sizeof(float)));
        MRT_CUDA(cudaMalloc(&ctx->d_sum_b, (size_t)w * h * sizeof(uint32_t)));
        ctx->w = w;
        ctx->h = h;
    }
    MRT_CUDA(cudaMemsetAsync(ctx->d_accum, 0, (size_t)w * h * 4 * sizeof(long long), ctx->stream));  // image.clear() main.rs:233
    MRT_CUDA(cudaMemsetAsync(ctx->d_nonfinite, 0, (size_t)w * h * sizeof(uint32_t), ctx->stream));
    ctx->count = 0;
    return MRT_OK;
}

// The shade queues are regions of one buffer, one region per queue kind the scene can produce: the miss queue and the material
// kinds present in its material table.
static uint32_t shade_regions(const mrt_context* ctx, uint8_t region[Q_COUNT + 3]) {
    uint32_t n = 0;
    for (int k = 0; k < Q_COUNT + 3; ++k) region[k] = 0;
    region[Q_MISS] = (uint8_t)n++;
    for (int kind = 0; kind < MRT_MAT_KINDS; ++kind)
        if (ctx->material_kinds & (1u << kind)) region[Q_FIRST_MAT + kind] = (uint8_t)n++;
    return n;
}

static int ensure_pool(mrt_context* ctx, uint64_t total_work, uint32_t regions) {
    // measured: 16 M rays in flight beat 4 M by 4 % (Cornell) to 26 % (1 M-triangle mesh), 32 M beat 16 M by another 0.4 % to 3 % (profiles/README.md)
    uint64_t want = ctx->opt_pool_slots ? ctx->opt_pool_slots : (1ull << 25);
    want = std::min<uint64_t>(want, std::max<uint64_t>((total_work + 1023) / 1024 * 1024, 1024));
    want = std::min<uint64_t>(want, 1ull << 26);
    if (ctx->pool.capacity == want && ctx->pool.regions >= regions) return MRT_OK;
    cudaStreamSynchronize(ctx->stream);
    free_pool(ctx);
    Pool& p = ctx->pool;
    MRT_CUDA(cudaMalloc(&p.q_ext[0], want * sizeof(RayRec)));
    MRT_CUDA(cudaMalloc(&p.q_ext[1], want * sizeof(RayRec)));
    MRT_CUDA(cudaMalloc(&p.q_shade, want * sizeof(HitEntry) * regions));
    p.capacity = (uint32_t)want;
    p.regions = regions;
    return MRT_OK;
}

static int render_accumulate_local(mrt_context* ctx, uint32_t spp_begin, uint32_t spp_count, uint32_t max_depth, uint64_t seed);
static int comm_reduce(mrt_context* ctx);

int mrt_sample_range(int rank, int size, uint32_t spp_begin, uint32_t spp_count, uint32_t* begin, uint32_t* count) {
    if (size < 1 || rank < 0 || rank >= size || !begin || !count) return MRT_E_INVALID;
    const uint64_t a = (uint64_t)spp_count * (uint64_t)rank / (uint64_t)size, b = (uint64_t)spp_count * (uint64_t)(rank + 1) / (uint64_t)size;
    *begin = spp_begin + (uint32_t)a;
    *count = (uint32_t)(b - a);
    return MRT_OK;
}

// PASS B over the members of a communicator: each renders its share of the sample range, then the one reduce (main.rs:235-294, 629-638)
int mrt_render_accumulate(mrt_context* ctx, uint32_t spp_begin, uint32_t spp_count, uint32_t max_depth, uint64_t seed) {
    if (!ctx) return MRT_E_INVALID;
    if (ctx->comm_size == 1 || !ctx->comm_split) return render_accumulate_local(ctx, spp_begin, spp_count, max_depth, seed);
    if ((uint64_t)spp_begin + spp_count > 0xFFFFFFFFull) return fail(ctx, MRT_E_INVALID, "sample range overflows 32 bits");
    int rc;
    if (!ctx->peers.empty()) {  // one process, several devices: member k = k-th device
        rc = for_each_member(ctx, [=](mrt_context* m) {
            uint32_t b = 0, c = 0;
            mrt_sample_range(m->comm_rank, m->comm_size, spp_begin, spp_count, &b, &c);
            return render_accumulate_local(m, b, c, max_depth, seed);
        });
    } else {  // one process per GPU
        uint32_t b = 0, c = 0;
        mrt_sample_range(ctx->comm_rank, ctx->comm_size, spp_begin, spp_count, &b, &c);
        rc = render_accumulate_local(ctx, b, c, max_depth, seed);
    }
    if (rc) return rc;
    return comm_reduce(ctx);
}

static int render_accumulate_local(mrt_context* ctx, uint32_t spp_begin, uint32_t spp_count, uint32_t max_depth, uint64_t seed) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!ctx->d_accum) return fail(ctx, MRT_E_STATE, "no image (mrt_accum_reset)");
    if (max_depth == 0) return fail(ctx, MRT_E_INVALID, "max_depth must be >= 1");
    if ((uint64_t)spp_begin + spp_count > 0xFFFFFFFFull) return fail(ctx, MRT_E_INVALID, "sample range overflows 32 bits");
    MRT_CUDA(cudaSetDevice(ctx->device));
    const uint32_t npix = ctx->w * ctx->h;
    const uint64_t total = (uint64_t)npix * spp_count;
    mrt_stats& st = ctx->stats;
    st.paths = total; st.rays = 0; st.node_visits = st.tri_tests = st.sphere_tests = st.instance_tests = st.volume_tests = 0;
    st.iterations = st.extend_launches = st.kernel_launches = 0;
    st.max_ray_node_visits = 0;
    st.render_ms = st.extend_ms = st.shade_ms = st.generate_ms = 0.0f;
    st.scene_bytes = ctx->scene_bytes;
    st.node_bytes = sizeof(DNode);
    if (total == 0) { st.pool_slots = ctx->pool.capacity; return MRT_OK; }
    RenderParams rp{ctx->w, ctx->h, npix, spp_begin, max_depth, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), ctx->opt_refill_lanes ? ctx->opt_refill_lanes : ctx->auto_refill_lanes, ctx->opt_node_burst ? ctx->opt_node_burst : ctx->auto_node_burst, ctx->opt_finish_paths, 0u, {}};
    const uint32_t regions = shade_regions(ctx, rp.region);
    if ((rc = ensure_pool(ctx, total, regions))) return rc;
    st.pool_slots = ctx->pool.capacity;
    Pool pool = ctx->pool;
    rp.capacity = pool.capacity;

    QueueState init;
    std::memset(&init, 0, sizeof init);
    init.total_work = total;
    ctx->h_q[2] = init;
    MRT_CUDA(cudaEventRecord(ctx->ev_begin, ctx->stream));
    MRT_CUDA(cudaMemcpyAsync(ctx->d_q, &ctx->h_q[2], sizeof(QueueState), cudaMemcpyHostToDevice, ctx->stream));

    std::vector<cudaEvent_t> tev;  // 6 events per iteration when kernel timing is on
    const int kChunk = 8, kDrainChunk = 3;
    int cur = 0;
    const int mode = ctx->scene.has_alpha ? 2 : (ctx->scene.n_volumes ? 1 : 0);  // which intersection code the scene needs
    // Per iteration: [finish,] generate, extend, shade (whose last block also does the bookkeeping for the next iteration). The
    // drain kernel k_finish returns at once unless advance() has handed it the last paths of the job; launching it in every
    // iteration (2.5 us) ends a job as soon as few enough paths are alive, which matters for renders of a few samples per pixel.
    k_advance<<<1, 32, 0, ctx->stream>>>(ctx->d_q, pool.capacity, rp.finish_paths);
    st.kernel_launches++;
    auto launch_chunk = [&](int slot, int len) -> int {
        for (int it = 0; it < len; ++it) {
            cudaEvent_t e[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            if (ctx->opt_time)
                for (int k = 0; k < 6; ++k) { MRT_CUDA(cudaEventCreate(&e[k])); tev.push_back(e[k]); }
            if (rp.finish_paths) {
                const unsigned gf = std::max(512u, (rp.finish_paths + 127u) / 128u);  // one thread per path at the threshold
                if (mode == 2) k_finish<true, true><<<gf, 128, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur, ctx->d_accum, ctx->d_nonfinite);
                else if (mode == 1) k_finish<false, true><<<gf, 128, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur, ctx->d_accum, ctx->d_nonfinite);
                else k_finish<false, false><<<gf, 128, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur, ctx->d_accum, ctx->d_nonfinite);
                st.kernel_launches++;
            }
            if (ctx->opt_time) MRT_CUDA(cudaEventRecord(e[0], ctx->stream));
            k_generate<<<ctx->grid_generate, 256, 0, ctx->stream>>>(ctx->cam, rp, pool, ctx->d_q, cur);
            if (ctx->opt_time) { MRT_CUDA(cudaEventRecord(e[1], ctx->stream)); MRT_CUDA(cudaEventRecord(e[2], ctx->stream)); }
            const int grid = ctx->grid_extend[ctx->opt_count ? 1 : 0][mode];
            if (ctx->opt_count) {
                if (mode == 2) k_extend<true, true, true><<<grid, kExtendThreads, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur);
                else if (mode == 1) k_extend<true, false, true><<<grid, kExtendThreads, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur);
                else k_extend<true, false, false><<<grid, kExtendThreads, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur);
            } else {
                if (mode == 2) k_extend<false, true, true><<<grid, kExtendThreads, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur);
                else if (mode == 1) k_extend<false, false, true><<<grid, kExtendThreads, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur);
                else k_extend<false, false, false><<<grid, kExtendThreads, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur);
            }
            if (ctx->opt_time) { MRT_CUDA(cudaEventRecord(e[3], ctx->stream)); MRT_CUDA(cudaEventRecord(e[4], ctx->stream)); }
            if (mode == 2) k_shade<true><<<ctx->grid_shade_full, MRT_SHADE_THREADS, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur, ctx->d_accum, ctx->d_nonfinite, rp.finish_paths);
            else k_shade<false><<<ctx->grid_shade, MRT_SHADE_THREADS, 0, ctx->stream>>>(ctx->scene, rp, pool, ctx->d_q, cur, ctx->d_accum, ctx->d_nonfinite, rp.finish_paths);
            if (ctx->opt_time) MRT_CUDA(cudaEventRecord(e[5], ctx->stream));
            cur ^= 1;
            st.kernel_launches += 3;
            st.extend_launches += 1;
        }
        MRT_CUDA(cudaGetLastError());
        MRT_CUDA(cudaMemcpyAsync(&ctx->h_q[slot], ctx->d_q, sizeof(QueueState), cudaMemcpyDeviceToHost, ctx->stream));
        MRT_CUDA(cudaEventRecord(ctx->ev_status[slot], ctx->stream));
        return MRT_OK;
    };
    // keep one chunk of iterations queued ahead of the one whose status is being read, so the stream never idles. Once a status
    // shows that every sample has been handed out (the job is draining), the chunks shrink: iterations queued past `done` are
    // empty launches (~17 us each), and a short job or one GPU's share of a split job should not pay 16 of them.
    int len = kChunk;
    if ((rc = launch_chunk(0, len))) return rc;
    for (int c = 0;; ++c) {
        if ((rc = launch_chunk((c + 1) & 1, len))) return rc;
        MRT_CUDA(cudaEventSynchronize(ctx->ev_status[c & 1]));
        const QueueState& seen = ctx->h_q[c & 1];
        if (seen.done) break;
        if (seen.next_work == seen.total_work) len = kDrainChunk;
    }
    MRT_CUDA(cudaMemcpyAsync(&ctx->h_q[2], ctx->d_q, sizeof(QueueState), cudaMemcpyDeviceToHost, ctx->stream));
    MRT_CUDA(cudaEventRecord(ctx->ev_end, ctx->stream));
    MRT_CUDA(cudaStreamSynchronize(ctx->stream));
    MRT_CUDA(cudaEventElapsedTime(&st.render_ms, ctx->ev_begin, ctx->ev_end));
    const QueueState& fin = ctx->h_q[2];
    st.rays = fin.rays + fin.n_ext;
    st.iterations = fin.iterations;
    st.node_visits = fin.visits.node_visits;
    st.tri_tests = fin.visits.tri_tests;
    st.sphere_tests = fin.visits.sphere_tests;
    st.instance_tests = fin.visits.instance_tests;
    st.volume_tests = fin.visits.volume_tests;
    st.max_ray_node_visits = fin.visits.max_ray_nodes;
    if (ctx->opt_time) {
        for (size_t i = 0; i + 5 < tev.size(); i += 6) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, tev[i], tev[i + 1]) == cudaSuccess) st.generate_ms += ms;
            if (cudaEventElapsedTime(&ms, tev[i + 2], tev[i + 3]) == cudaSuccess) st.extend_ms += ms;
            if (cudaEventElapsedTime(&ms, tev[i + 4], tev[i + 5]) == cudaSuccess) st.shade_ms += ms;
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    ctx->count += spp_count;
    return MRT_OK;
}

int mrt_accum_device_ptr(mrt_context* ctx, void** accum_i64, uint64_t* n_elems) {
    if (!ctx) return MRT_E_INVALID;
    if (!ctx->d_accum) return fail(ctx, MRT_E_STATE, "no image (mrt_accum_reset)");
    if (accum_i64) *accum_i64 = ctx->d_accum;
    if (n_elems) *n_elems = (uint64_t)ctx->w * ctx->h * 4;
    return MRT_OK;
}

int mrt_accum_download(mrt_context* ctx, float* sum_rgb, uint32_t* sum_bounces, uint32_t* out_count) {
    if (!ctx) return MRT_E_INVALID;
    if (!ctx->d_accum) return fail(ctx, MRT_E_STATE, "no image (mrt_accum_reset)");
    MRT_CUDA(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)ctx->w * ctx->h;
    k_resolve_sums<<<ctx->n_sms * 4, 256, 0, ctx->stream>>>(ctx->d_accum, ctx->d_nonfinite, (uint32_t)npix, ctx->d_sum_rgb, ctx->d_sum_b);
    MRT_CUDA(cudaGetLastError());
    if (sum_rgb) MRT_CUDA(cudaMemcpyAsync(sum_rgb, ctx->d_sum_rgb, npix * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (sum_bounces) MRT_CUDA(cudaMemcpyAsync(sum_bounces, ctx->d_sum_b, npix * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MRT_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_count) *out_count = ctx->count;
    return MRT_OK;
}

int mrt_render(mrt_context* ctx, uint32_t w, uint32_t h, uint32_t spp_begin, uint32_t spp_count, uint32_t max_depth, uint64_t seed, float* sum_rgb,
               uint32_t* sum_bounces, uint32_t* out_count) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if ((rc = mrt_accum_reset(ctx, w, h))) return rc;
    if ((rc = mrt_render_accumulate(ctx, spp_begin, spp_count, max_depth, seed))) return rc;
    if (ctx->comm_rank != 0 && ctx->comm_split) {  // the merged image lives on the root; this member's buffer was cleared by the merge
        if (out_count) *out_count = 0;
        return MRT_OK;
    }
    return mrt_accum_download(ctx, sum_rgb, sum_bounces, out_count);
}

int mrt_resolve_rgb8(mrt_context* ctx, int mode, int flip, uint32_t count, uint8_t* out) {
    if (!ctx) return MRT_E_INVALID;
    if (!ctx->d_accum) return fail(ctx, MRT_E_STATE, "no image (mrt_accum_reset)");
    if (!out) return fail(ctx, MRT_E_INVALID, "out is NULL");
    if (mode != 0 && mode != 1 && mode != 2) return fail(ctx, MRT_E_UNSUPPORTED, "only Default(0), Denoise(1, as Default) and Depth(2) are resolved on the device");
    MRT_CUDA(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)ctx->w * ctx->h;
    uint8_t* d_out = reinterpret_cast<uint8_t*>(ctx->d_sum_rgb);  // staging buffers of the image, reused as byte output / max cell
    uint32_t* d_max = ctx->d_sum_b;
    cudaMemsetAsync(d_max, 0, 4, ctx->stream);
    uint32_t h_max = 0;
    if (mode == 2) {
        k_max_bounces<<<ctx->n_sms * 4, 256, 0, ctx->stream>>>(ctx->d_accum, (uint32_t)npix, d_max);
        cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    k_resolve_rgb8<<<ctx->n_sms * 4, 256, 0, ctx->stream>>>(ctx->d_accum, ctx->d_nonfinite, ctx->w, ctx->h, count, mode == 2 ? 2 : 0, flip, h_max, d_out);
    cudaMemcpyAsync(out, d_out, npix * 3, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(ctx, MRT_E_CUDA, cudaGetErrorString(e));
    return MRT_OK;
}

static int set_option_one(mrt_context* ctx, int option, uint64_t value);
int mrt_set_option(mrt_context* ctx, int option, uint64_t value) {
    if (!ctx) return MRT_E_INVALID;
    int rc = set_option_one(ctx, option, value);
    for (mrt_context* p : ctx->peers)
        if (!rc) rc = set_option_one(p, option, value);
    return rc;
}
static int set_option_one(mrt_context* ctx, int option, uint64_t value) {
    switch (option) {
        case MRT_OPT_COMM_SPLIT: ctx->comm_split = value != 0; return MRT_OK;
        case MRT_OPT_COMM_SCENE: ctx->comm_scene = value != 0; return MRT_OK;
        case MRT_OPT_COUNT_VISITS: ctx->opt_count = value != 0; return MRT_OK;
        case MRT_OPT_TIME_KERNELS: ctx->opt_time = value != 0; return MRT_OK;
        case MRT_OPT_POOL_SLOTS:
            if (value != 0 && (value < 1024 || value > (1ull << 26))) return fail(ctx, MRT_E_INVALID, "pool slots out of range [1024, 2^26]");
            ctx->opt_pool_slots = value / 1024 * 1024;
            return MRT_OK;
        case MRT_OPT_REFILL_LANES:
            if (value > 32) return fail(ctx, MRT_E_INVALID, "lane threshold out of range [0, 32]");
            ctx->opt_refill_lanes = (uint32_t)value;
            return MRT_OK;
        case MRT_OPT_FINISH_PATHS:
            if (value > (1u << 22)) return fail(ctx, MRT_E_INVALID, "finish threshold out of range [0, 2^22]");
            ctx->opt_finish_paths = (uint32_t)value;
            return MRT_OK;
        case MRT_OPT_DEVICE_BUILD:
            if (value > 2) return fail(ctx, MRT_E_INVALID, "device build mode out of range [0, 2]");
            ctx->opt_device_build = (uint32_t)value;
            return MRT_OK;
        case MRT_OPT_NODE_BURST: ctx->opt_node_burst = (uint32_t)value; return MRT_OK;
        case MRT_OPT_BVH_LEAF_TRIS:
            if (value < 1 || value > 4) return fail(ctx, MRT_E_INVALID, "leaf size out of range [1, 4]");
            ctx->opt_leaf_tris = (uint32_t)value;
            return MRT_OK;
        case MRT_OPT_BVH_TRI_COST:
            if (value < 1 || value > 10000) return fail(ctx, MRT_E_INVALID, "triangle cost out of range [1, 10000]");
            ctx->opt_tri_cost = (uint32_t)value;
            return MRT_OK;
        default: return fail(ctx, MRT_E_INVALID, "unknown option");
    }
}

int mrt_get_stats(mrt_context* ctx, mrt_stats* out) {
    if (!ctx || !out) return MRT_E_INVALID;
    *out = ctx->stats;
    for (const mrt_context* p : ctx->peers) {  // a multi-device handle reports the whole job: counts summed, times = the slowest device
        const mrt_stats& q = p->stats;
        out->paths += q.paths; out->rays += q.rays; out->node_visits += q.node_visits; out->tri_tests += q.tri_tests;
        out->sphere_tests += q.sphere_tests; out->instance_tests += q.instance_tests; out->volume_tests += q.volume_tests;
        out->extend_launches += q.extend_launches; out->kernel_launches += q.kernel_launches;
        out->iterations = std::max(out->iterations, q.iterations);
        out->max_ray_node_visits = std::max(out->max_ray_node_visits, q.max_ray_node_visits);
        out->render_ms = std::max(out->render_ms, q.render_ms); out->extend_ms = std::max(out->extend_ms, q.extend_ms);
        out->shade_ms = std::max(out->shade_ms, q.shade_ms); out->generate_ms = std::max(out->generate_ms, q.generate_ms);
    }
    return MRT_OK;
}

int mrt_synchronize(mrt_context* ctx) {
    if (!ctx) return MRT_E_INVALID;
    for (mrt_context* p : ctx->peers) {
        cudaSetDevice(p->device);
        MRT_CUDA(cudaStreamSynchronize(p->stream));
    }
    MRT_CUDA(cudaSetDevice(ctx->device));
    MRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return MRT_OK;
}

// ---- communicators ---------------------------------------------------------------------------------------------------
static int nccl_fail(mrt_context* ctx, NcclApi* api, const char* what, ncclResult_t r) {
    return fail(ctx, MRT_E_CUDA, std::string(what) + ": " + (api && api->GetErrorString ? api->GetErrorString(r) : "NCCL error"));
}
static int comm_cells(mrt_context* ctx) {
    if (ctx->d_count) return MRT_OK;
    MRT_CUDA(cudaSetDevice(ctx->device));
    MRT_CUDA(cudaMalloc(&ctx->d_count, sizeof(unsigned long long)));
    MRT_CUDA(cudaMallocHost(&ctx->h_count, sizeof(unsigned long long)));
    return MRT_OK;
}

int mrt_context_create_multi(const int* devices, int n_devices, mrt_context** out) {
    if (!out) { g_create_error = "out is NULL"; return MRT_E_INVALID; }
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 255) { g_create_error = "device list is empty or longer than 255"; return MRT_E_INVALID; }
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) { g_create_error = "device listed twice"; return MRT_E_INVALID; }
    mrt_context* lead = nullptr;
    int rc = mrt_context_create(devices[0], nullptr, &lead);
    if (rc) return rc;
    if (n_devices == 1) { *out = lead; return MRT_OK; }
    auto bail = [&](int code, const std::string& msg) { g_create_error = msg; mrt_context_destroy(lead); return code; };
    for (int i = 1; i < n_devices; ++i) {
        mrt_context* p = nullptr;
        if ((rc = mrt_context_create(devices[i], nullptr, &p))) { mrt_context_destroy(lead); return rc; }
        lead->peers.push_back(p);
    }
    std::string why;
    NcclApi* api = nccl_api(why);
    if (!api) return bail(MRT_E_UNSUPPORTED, why);
    std::vector<ncclComm_t> comms((size_t)n_devices);
    ncclResult_t r = api->CommInitAll(comms.data(), n_devices, devices);
    if (r != ncclSuccess) return bail(MRT_E_CUDA, std::string("ncclCommInitAll: ") + api->GetErrorString(r));
    for (int i = 0; i < n_devices; ++i) {
        mrt_context* m = i == 0 ? lead : lead->peers[(size_t)i - 1];
        m->comm = comms[(size_t)i];
        m->comm_rank = i;
        m->comm_size = n_devices;
        if ((rc = comm_cells(m))) return bail(rc, m->err);
        // scene replication goes device to device: direct over NVLink where peer access can be enabled, staged otherwise
        if (i > 0) {
            int can = 0;
            cudaSetDevice(devices[i]);
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can) cudaDeviceEnablePeerAccess(devices[0], 0);
            cudaGetLastError();  // "already enabled" is fine
        }
    }
    cudaSetDevice(devices[0]);
    *out = lead;
    return MRT_OK;
}

int mrt_comm_unique_id(uint8_t id[MRT_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == MRT_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id) return MRT_E_INVALID;
    std::string why;
    NcclApi* api = nccl_api(why);
    if (!api) { g_create_error = why; return MRT_E_UNSUPPORTED; }
    ncclUniqueId u;
    ncclResult_t r = api->GetUniqueId(&u);
    if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + api->GetErrorString(r); return MRT_E_CUDA; }
    std::memcpy(id, u.internal, MRT_COMM_ID_BYTES);
    return MRT_OK;
}

int mrt_comm_init_rank(mrt_context* ctx, const uint8_t id[MRT_COMM_ID_BYTES], int rank, int n_ranks) {
    if (!ctx) return MRT_E_INVALID;
    if (!id || n_ranks < 1 || n_ranks > 255 || rank < 0 || rank >= n_ranks) return fail(ctx, MRT_E_INVALID, "bad rank / n_ranks");
    if (ctx->comm || !ctx->peers.empty()) return fail(ctx, MRT_E_STATE, "context already belongs to a communicator");
    if (n_ranks == 1) return MRT_OK;
    std::string why;
    NcclApi* api = nccl_api(why);
    if (!api) return fail(ctx, MRT_E_UNSUPPORTED, why);
    MRT_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId u;
    std::memcpy(u.internal, id, MRT_COMM_ID_BYTES);
    ncclComm_t comm = nullptr;
    ncclResult_t r = api->CommInitRank(&comm, n_ranks, u, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, api, "ncclCommInitRank", r);
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_size = n_ranks;
    return comm_cells(ctx);
}

int mrt_comm_rank(mrt_context* ctx, int* rank, int* size) {
    if (!ctx) return MRT_E_INVALID;
    if (rank) *rank = ctx->comm_rank;
    if (size) *size = ctx->comm_size;
    return MRT_OK;
}

// Image::merge (main.rs:629-638) across the members: accumulators (+ folded non-finite flags) and sample counts summed onto
// member 0 on the render streams; the other members' images are cleared.
static int comm_reduce(mrt_context* ctx) {
    if (ctx->comm_size == 1) return MRT_OK;
    std::string why;
    NcclApi* api = nccl_api(why);
    if (!api) return fail(ctx, MRT_E_UNSUPPORTED, why);
    std::vector<mrt_context*> members{ctx};
    members.insert(members.end(), ctx->peers.begin(), ctx->peers.end());
    for (mrt_context* m : members) {
        if (!m->d_accum) return fail(ctx, MRT_E_STATE, "no image (mrt_accum_reset)");
        if (m->w != ctx->w || m->h != ctx->h) return fail(ctx, MRT_E_STATE, "members hold images of different sizes");
    }
    const uint32_t npix = ctx->w * ctx->h;
    for (mrt_context* m : members) {
        MRT_CUDA(cudaSetDevice(m->device));
        k_fold_flags<<<m->n_sms * 4, 256, 0, m->stream>>>(m->d_accum, m->d_nonfinite, npix);
        k_set_cell<<<1, 1, 0, m->stream>>>(m->d_count, (unsigned long long)m->count);
    }
    ncclResult_t r = api->GroupStart();
    if (r != ncclSuccess) return nccl_fail(ctx, api, "ncclGroupStart", r);
    for (mrt_context* m : members) {
        ncclComm_t comm = static_cast<ncclComm_t>(m->comm);
        if (r == ncclSuccess) r = api->Reduce(m->d_accum, m->d_accum, (size_t)npix * 4, ncclInt64, ncclSum, 0, comm, m->stream);
        if (r == ncclSuccess) r = api->Reduce(m->d_count, m->d_count, 1, ncclUint64, ncclSum, 0, comm, m->stream);
    }
    ncclResult_t r2 = api->GroupEnd();
    if (r != ncclSuccess) return nccl_fail(ctx, api, "ncclReduce", r);
    if (r2 != ncclSuccess) return nccl_fail(ctx, api, "ncclGroupEnd", r2);
    for (mrt_context* m : members) {
        MRT_CUDA(cudaSetDevice(m->device));
        if (m->comm_rank == 0) {
            k_unfold_flags<<<m->n_sms * 4, 256, 0, m->stream>>>(m->d_accum, m->d_nonfinite, npix);
            MRT_CUDA(cudaMemcpyAsync(m->h_count, m->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
        } else {  // merged into the root: this member's private image starts over
            MRT_CUDA(cudaMemsetAsync(m->d_accum, 0, (size_t)npix * 4 * sizeof(long long), m->stream));
            MRT_CUDA(cudaMemsetAsync(m->d_nonfinite, 0, (size_t)npix * sizeof(uint32_t), m->stream));
        }
    }
    for (mrt_context* m : members) {
        MRT_CUDA(cudaSetDevice(m->device));
        MRT_CUDA(cudaStreamSynchronize(m->stream));
        m->count = m->comm_rank == 0 ? (uint32_t)*m->h_count : 0u;
    }
    MRT_CUDA(cudaSetDevice(ctx->device));
    return MRT_OK;
}

int mrt_comm_reduce(mrt_context* ctx) {
    if (!ctx) return MRT_E_INVALID;
    return comm_reduce(ctx);
}

// The root's finished scene to every member of a one-process-per-GPU communicator. A fixed-size header goes first (the root's
// upload status, so that a failed upload fails everywhere instead of hanging the others; the arena's array sizes and device
// addresses; DScene and the few host-side facts derived at upload), then every arena array, all in one NCCL group on the render
// streams. The receivers rebase DScene's pointers from the root's addresses to their own.
namespace {
constexpr size_t kMaxSceneBufs = 24;
struct SceneHeader {
    int32_t rc;
    uint32_t n_bufs;
    uint32_t auto_refill_lanes, auto_node_burst, material_kinds, pad;
    uint64_t scene_bytes;
    uint64_t buf_addr[kMaxSceneBufs], buf_used[kMaxSceneBufs];
    DScene scene;
};
}  // namespace
static int comm_scene_broadcast(mrt_context* ctx, int root_rc) {
    std::string why;
    NcclApi* api = nccl_api(why);
    if (!api) return fail(ctx, MRT_E_UNSUPPORTED, why);
    MRT_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->d_scene_hdr) {
        MRT_CUDA(cudaMalloc(&ctx->d_scene_hdr, sizeof(SceneHeader)));
        MRT_CUDA(cudaMallocHost(&ctx->h_scene_hdr, sizeof(SceneHeader)));
    }
    SceneHeader* h = static_cast<SceneHeader*>(ctx->h_scene_hdr);
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    const bool root = ctx->comm_rank == 0;
    if (root) {
        std::memset(h, 0, sizeof *h);
        h->rc = root_rc;
        if (!root_rc && ctx->scene_buf_next > kMaxSceneBufs) h->rc = MRT_E_UNSUPPORTED;
        if (!h->rc) {
            h->n_bufs = (uint32_t)ctx->scene_buf_next;
            for (size_t k = 0; k < ctx->scene_buf_next; ++k) {
                h->buf_addr[k] = reinterpret_cast<uint64_t>(ctx->scene_bufs[k].p);
                h->buf_used[k] = ctx->scene_bufs[k].used;
            }
            h->auto_refill_lanes = ctx->auto_refill_lanes;
            h->auto_node_burst = ctx->auto_node_burst;
            h->material_kinds = ctx->material_kinds;
            h->scene_bytes = ctx->scene_bytes;
            h->scene = ctx->scene;
        }
        MRT_CUDA(cudaMemcpyAsync(ctx->d_scene_hdr, h, sizeof *h, cudaMemcpyHostToDevice, ctx->stream));
    }
    ncclResult_t r = api->Broadcast(ctx->d_scene_hdr, ctx->d_scene_hdr, sizeof(SceneHeader), ncclUint8, 0, comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, api, "ncclBroadcast", r);
    if (!root) {
        MRT_CUDA(cudaMemcpyAsync(h, ctx->d_scene_hdr, sizeof *h, cudaMemcpyDeviceToHost, ctx->stream));
        MRT_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (h->rc) {
        if (root) return h->rc;  // (ctx->err was set by the upload)
        return fail(ctx, h->rc, "the root rank's mrt_scene_upload failed");  // (this rank keeps the scene it had, like the root)
    }
    if (!root) free_scene(ctx);  // (the stream was synchronised above: nothing reads the old arrays any more)
    std::vector<void*> mine(h->n_bufs);
    for (uint32_t k = 0; k < h->n_bufs; ++k) {
        if (root) { mine[k] = ctx->scene_bufs[k].p; continue; }
        int rc = arena_alloc(ctx, (size_t)h->buf_used[k], &mine[k]);
        if (rc) return rc;  // (the other ranks then wait in the broadcast below: an allocation failure here is fatal for the job anyway)
    }
    if ((r = api->GroupStart()) != ncclSuccess) return nccl_fail(ctx, api, "ncclGroupStart", r);
    for (uint32_t k = 0; k < h->n_bufs && r == ncclSuccess; ++k) r = api->Broadcast(mine[k], mine[k], (size_t)h->buf_used[k], ncclUint8, 0, comm, ctx->stream);
    ncclResult_t r2 = api->GroupEnd();
    if (r != ncclSuccess) return nccl_fail(ctx, api, "ncclBroadcast", r);
    if (r2 != ncclSuccess) return nccl_fail(ctx, api, "ncclGroupEnd", r2);
    if (!root) {
        DScene d = h->scene;
        auto rebase = [&](auto*& ptr) {
            if (!ptr) return;
            const uint64_t q = reinterpret_cast<uint64_t>(ptr);
            for (uint32_t k = 0; k < h->n_bufs; ++k)
                if (q >= h->buf_addr[k] && q < h->buf_addr[k] + std::max<uint64_t>(h->buf_used[k], 1)) {
                    ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(static_cast<char*>(mine[k]) + (q - h->buf_addr[k]));
                    return;
                }
        };
        rebase(d.nodes); rebase(d.spheres); rebase(d.sphere_aux); rebase(d.tri_verts); rebase(d.tri_map); rebase(d.tri_shading);
        rebase(d.instances); rebase(d.blas); rebase(d.volumes); rebase(d.materials); rebase(d.surfaces); rebase(d.textures); rebase(d.texels);
        ctx->scene = d;
        ctx->has_scene = true;
        ctx->auto_refill_lanes = h->auto_refill_lanes;
        ctx->auto_node_burst = h->auto_node_burst;
        ctx->material_kinds = h->material_kinds;
        ctx->scene_bytes = h->scene_bytes;
    }
    MRT_CUDA(cudaStreamSynchronize(ctx->stream));
    return MRT_OK;
}

// ---- test hooks ---------------------------------------------------------------------------------------------------
int mrt_debug_philox(mrt_context* ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    if (!ctx) return MRT_E_INVALID;
    uint4* d = nullptr;
    MRT_CUDA(cudaMalloc(&d, sizeof(uint4)));
    k_debug_philox<<<1, 1, 0, ctx->stream>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), d);
    cudaMemcpyAsync(out, d, 16, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, MRT_E_CUDA, cudaGetErrorString(e));
    return MRT_OK;
}
int mrt_debug_samplers(mrt_context* ctx, uint64_t seed, uint32_t n, float* in_ball3, float* unit_vec3, float* in_disk2) {
    if (!ctx) return MRT_E_INVALID;
    float *b = nullptr, *s = nullptr, *d = nullptr;
    MRT_CUDA(cudaMalloc(&b, (size_t)n * 12));
    MRT_CUDA(cudaMalloc(&s, (size_t)n * 12));
    MRT_CUDA(cudaMalloc(&d, (size_t)n * 8));
    k_debug_samplers<<<ctx->n_sms * 2, 256, 0, ctx->stream>>>(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), n, b, s, d);
    cudaMemcpyAsync(in_ball3, b, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(unit_vec3, s, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(in_disk2, d, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(b); cudaFree(s); cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, MRT_E_CUDA, cudaGetErrorString(e));
    return MRT_OK;
}

}  // extern "C"

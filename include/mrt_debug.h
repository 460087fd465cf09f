/*
 * mrt_debug.h — test hooks exported by libmrt_cuda.so (not part of the drop-in boundary).
 * They expose device-side building blocks so tests can pin them: the Philox4x32-10 generator against the
 * Random123 known-answer vectors, and the closed-form samplers against the distributions of the reference's
 * rejection samplers (math.rs:80-109).
 */
#ifndef MRT_DEBUG_H
#define MRT_DEBUG_H
#include "mrt.h"
#ifdef __cplusplus
extern "C" {
#endif
int mrt_debug_philox(mrt_context* ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* n draws each of: uniform in unit ball (3 floats), uniform on unit sphere (3), uniform in unit disk (2) */
int mrt_debug_samplers(mrt_context* ctx, uint64_t seed, uint32_t n, float* in_ball3, float* unit_vec3, float* in_disk2);
#ifdef __cplusplus
}
#endif
#endif

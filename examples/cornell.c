/* A host in plain C: the Cornell box of the reference (src/scenes/cornell.rs:29-99) built through libmrt_host.so, rendered through
 * libmrt_cuda.so, written as a binary PPM (or a PNG when the output name ends in .png). Nothing but the two C headers is used --
 * this is the call sequence a foreign-language host (the reference's Rust main.rs) goes through, and the
 * program the parity suite runs to check that the boundary gives the same image whichever language drives it
 * (tests/test_c_host.py).
 *
 *   cc -std=c11 -I include examples/cornell.c -L mass_raytrace_b200 -lmrt_host -lmrt_cuda -lm -o cornell
 *   ./cornell mass_raytrace_b200/assets/cube.ply 512 512 64 out.png [n_gpus]
 *
 * With n_gpus > 1 the same calls drive that many devices of this process (mrt_context_create_multi): the 64 samples are split
 * over them, merged by one NCCL reduce, and the image is byte for byte the one a single GPU gives.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mrt.h"
#include "mrt_host.h"

static int die_scene(mrth_scene* s, const char* what) {
    fprintf(stderr, "%s: %s\n", what, mrth_last_error(s));
    return 2;
}

static int die_ctx(mrt_context* c, const char* what) {
    fprintf(stderr, "%s: %s\n", what, mrt_last_error(c));
    return 3;
}

/* cube.instance(translation, rotation, scale).with_material(m)  geom.rs:288-298, 383-392 */
static int add_cube(mrth_scene* s, int cube, float tx, float ty, float tz, float ry, float sx, float sy, float sz, int material) {
    const float t[3] = {tx, ty, tz}, r[3] = {0.0f, ry, 0.0f}, sc[3] = {sx, sy, sz};
    return mrth_add_instance(s, cube, t, r, sc, material);
}

int main(int argc, char** argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s cube.ply width height spp out.ppm|out.png [n_gpus]\n", argv[0]);
        return 1;
    }
    const char* ply = argv[1];
    const uint32_t w = (uint32_t)atoi(argv[2]), h = (uint32_t)atoi(argv[3]), spp = (uint32_t)atoi(argv[4]);
    const int n_gpus = argc > 6 ? atoi(argv[6]) : 1;
    if (w == 0 || h == 0 || spp == 0 || n_gpus < 1 || n_gpus > 64) {
        fprintf(stderr, "width, height, spp and n_gpus must be positive\n");
        return 1;
    }

    mrth_scene* s = mrth_scene_new();
    if (!s) return 2;
    mrth_seed(s, 1); /* fastrand::seed(1) main.rs:86 */
    mrth_defer_mesh_bvh(s, 1);
    mrth_background_solid(s, 0.0f, 0.0f, 0.0f);
    const int red = mrth_mat_lambertian(s, mrth_surface_solid(s, 1.0f, 0.0f, 0.0f, 1.0f));
    const int green = mrth_mat_lambertian(s, mrth_surface_solid(s, 0.0f, 1.0f, 0.0f, 1.0f));
    const int white = mrth_mat_lambertian(s, mrth_surface_solid(s, 1.0f, 1.0f, 1.0f, 1.0f));
    const int light = mrth_mat_diffuse_light(s, 8.0f, 8.0f, 8.0f);
    const int glass = mrth_mat_dielectric(s, 1.3f);
    const int absorb = mrth_mat_absorb(s);
    const int perm[3] = {0, 1, 2};
    const int cube = mrth_mesh_load_ply(s, ply, perm, absorb, NULL);
    if (red < 0 || green < 0 || white < 0 || light < 0 || glass < 0 || absorb < 0 || cube < 0) return die_scene(s, "scene setup");

    int rc = 0;
    rc |= add_cube(s, cube, -10.0f, 5.0f, 0.0f, 0.0f, 5.0f, 5.0f, 5.0f, red);
    rc |= add_cube(s, cube, 10.0f, 5.0f, 0.0f, 0.0f, 5.0f, 5.0f, 5.0f, green);
    rc |= add_cube(s, cube, 0.0f, 15.0f, 0.0f, 0.0f, 5.0f, 5.0f, 5.0f, white);
    rc |= add_cube(s, cube, 0.0f, 5.0f, -10.0f, 0.0f, 5.0f, 5.0f, 5.0f, white);
    rc |= add_cube(s, cube, 0.0f, -5.0f, -0.0f, 0.0f, 5.0f, 5.0f, 5.0f, white);
    rc |= mrth_add_sphere(s, glass, 1.75f, 2.0f, 2.25f, 2.0f);
    rc |= add_cube(s, cube, 0.0f, 10.0f - 0.00011f, 0.0f, 0.0f, 1.0f, 0.0001f, 1.0f, light);
    rc |= add_cube(s, cube, -2.0f, 3.0f, -1.0f, -0.05f, 1.75f, 3.1f, 1.75f, white);
    if (rc < 0) return die_scene(s, "adding objects");
    mrth_build_bvh(s); /* world.build_bvh() main.rs:112 */
    const float from[3] = {0.0f, 5.0f, 20.0f}, at[3] = {0.0f, 5.0f, 0.0f}, up[3] = {0.0f, 1.0f, 0.0f};
    mrth_camera(s, 37.0f, from, at, up, (float)w / (float)h, 0.0f, 20.0f);

    mrt_context* ctx = NULL;
    if (n_gpus == 1) {
        if (mrt_context_create(0, NULL, &ctx) != MRT_OK) return die_ctx(NULL, "mrt_context_create");
    } else { /* the render-thread fan-out of main.rs:159-170, as devices */
        int devices[64];
        for (int i = 0; i < n_gpus; ++i) devices[i] = i;
        if (mrt_context_create_multi(devices, n_gpus, &ctx) != MRT_OK) return die_ctx(NULL, "mrt_context_create_multi");
    }
    if (mrt_scene_upload(ctx, mrth_scene_desc(s)) != MRT_OK) return die_ctx(ctx, "mrt_scene_upload");
    if (mrt_camera_set(ctx, mrth_scene_camera(s)) != MRT_OK) return die_ctx(ctx, "mrt_camera_set");

    /* the image stays on the device between the render and the tone map, like Image in main.rs:543-638 */
    if (mrt_accum_reset(ctx, w, h) != MRT_OK) return die_ctx(ctx, "mrt_accum_reset");
    if (mrt_render_accumulate(ctx, 0, spp, 50, 1) != MRT_OK) return die_ctx(ctx, "mrt_render_accumulate");
    uint8_t* rgb = (uint8_t*)malloc((size_t)w * h * 3);
    if (!rgb) return 4;
    if (mrt_resolve_rgb8(ctx, 0, 1, spp, rgb) != MRT_OK) return die_ctx(ctx, "mrt_resolve_rgb8");

    mrt_stats st;
    if (mrt_get_stats(ctx, &st) != MRT_OK) return die_ctx(ctx, "mrt_get_stats");
    uint64_t hash = 14695981039346656037ull; /* FNV-1a over the rgb8 image */
    for (size_t i = 0; i < (size_t)w * h * 3; ++i) hash = (hash ^ rgb[i]) * 1099511628211ull;
    printf("paths %llu rays %llu render_ms %.3f fnv1a %016llx\n", (unsigned long long)st.paths, (unsigned long long)st.rays, st.render_ms,
           (unsigned long long)hash);

    const size_t len = strlen(argv[5]);
    if (len > 4 && strcmp(argv[5] + len - 4, ".png") == 0) { /* Image::dump main.rs:760-783 */
        if (mrth_write_png(argv[5], rgb, w, h) != MRT_OK) return 5;
    } else {
        FILE* f = fopen(argv[5], "wb");
        if (!f) {
            perror(argv[5]);
            return 5;
        }
        fprintf(f, "P6\n%u %u\n255\n", w, h);
        fwrite(rgb, 1, (size_t)w * h * 3, f);
        fclose(f);
    }
    free(rgb);
    mrt_context_destroy(ctx);
    mrth_scene_free(s);
    return 0;
}

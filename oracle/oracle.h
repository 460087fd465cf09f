/*
 * oracle.h — C surface of the CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A C++17 restatement of nickmass/mass-raytrace's CPU path tracer (the reference is Rust and cannot
 * be compiled in this image: no rustc/cargo).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (libmrt_cuda.so,
 * libmrt_host.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures and cannot be run here, so
 * this oracle is pinned only by hand-derived known-answer tests (tests/test_oracle_kat.py, SURVEY.md
 * §8c items 1-10).  Two things are knowingly not the reference's: (1) the RNG stream — `fastrand`
 * 1.4.1 is a Cargo.lock dependency whose source is absent; it is restated from its published WyRand
 * algorithm and no reference output pins it; (2) tie order inside BvhNode::new's sort, which in Rust
 * depends on `sort_by` internals when the comparator never returns Equal (geom.rs:130-136).
 *
 * The builder functions mirror the reference scene API one to one (World::add, Sphere::new, Model::new,
 * Model::instance(..).with_material(..), Volume::new, Camera::new, PlyLoader::load); handles are small
 * integers owned by the orc_scene.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

typedef struct orc_counters {   /* per-render visit counters (SURVEY.md §8d N_box, N_tri, N_sph, N_inst) */
    uint64_t rays;              /* scene.intersect calls made by trace / albedo_normal (world.rs:68, 82) */
    uint64_t paths;
    uint64_t box_tests;         /* BoundingBox::hit calls        geom.rs:218 */
    uint64_t tri_tests;         /* Triangle::intersect calls     geom.rs:504 */
    uint64_t sphere_tests;      /* Sphere::intersect calls       geom.rs:57  */
    uint64_t instance_tests;    /* Instance::intersect calls     geom.rs:404 */
    uint64_t volume_tests;      /* Volume::intersect calls       geom.rs:612 */
} orc_counters;

orc_scene* orc_scene_new(void);
void       orc_scene_free(orc_scene*);
const char* orc_last_error(orc_scene*);
void       orc_seed(orc_scene*, uint64_t seed);            /* fastrand::seed  main.rs:86 (scene-thread RNG: BVH axes) */
float      orc_rand_f32(orc_scene*);                       /* f32::rand()     math.rs:244 on the scene-thread RNG */

/* surfaces  texture.rs */
int orc_surface_solid(orc_scene*, float r, float g, float b, float a);                       /* SolidColor :179 */
int orc_surface_texture(orc_scene*, const uint8_t* rgba, uint32_t w, uint32_t h, int wrap);  /* Texture::load_bytes :70; wrap 1=Repeat 2=Clamp */
int orc_surface_ycbcr(orc_scene*, int luma_tex, int chroma_tex);                             /* YCbCrTexture :209 */
int orc_surface_blend(orc_scene*, int mode, int left, int right);                            /* TextureBlend :302; mode 0..3 = Lighten,Darken,Addition,Subtraction */
int orc_surface_fallback(orc_scene*, float r, float g, float b, float a, int inner);         /* SolidColorFallback :336 */

/* materials  material.rs */
int orc_mat_absorb(orc_scene*);                                       /* impl Material for () :385 */
int orc_mat_lambertian(orc_scene*, int surface);                      /* :192 */
int orc_mat_diffuse_light(orc_scene*, float r, float g, float b);     /* :227 */
int orc_mat_metal(orc_scene*, float fuzz, int surface);               /* :248 */
int orc_mat_dielectric(orc_scene*, float ior);                        /* :286 */
int orc_mat_specular(orc_scene*, float ior, int surface);             /* :331 */
int orc_mat_mix(orc_scene*, float ratio, int left, int right);        /* :391 */
int orc_mat_isotropic(orc_scene*, float r, float g, float b);         /* Isotrophic :428 */
int orc_mat_eve(orc_scene*, int normal_occlusion, int albedo_roughness, int pmdg, const float colors12[12], const float glow3[3]); /* EveMaterial eve.rs:43-133 */

/* backgrounds  material.rs:39-190 */
void orc_background_solid(orc_scene*, float r, float g, float b);
void orc_background_sky(orc_scene*);
void orc_background_skysphere(orc_scene*, int surface);
void orc_background_cubemap(orc_scene*, const int surfaces6[6], float rx, float ry, float rz);

/* geometry  geom.rs */
/* Model::new over Triangle::new(material, a, b, c) for each of n_tris (verts: 9 floats per triangle) :281, :449 */
int orc_mesh_new(orc_scene*, const float* verts, uint64_t n_tris, int tri_material);
/* Model::new over Triangle::with_norms_and_uvs (normals 9 floats, uvs 6 floats per triangle) :468 */
int orc_mesh_new_uv(orc_scene*, const float* verts, const float* normals, const float* uvs, uint64_t n_tris, int tri_material);
/* PlyLoader::load(path, |x,y,z| V3(v[perm0],v[perm1],v[perm2]), |a,b,c| Triangle::new(material,a,b,c)) ply_loader.rs:273
   max_abs (nullable) receives max(|x|,|y|,|z|) over vertices (scenes/lucy.rs:37). Returns mesh handle or <0. */
int orc_mesh_load_ply(orc_scene*, const char* path, const int perm[3], int tri_material, float* max_abs);
/* StlLoader::load_binary(path, |x,y,z| V3(v[perm]), |a,b,c| Triangle::new(material,a,b,c)) stl_loader.rs:10 */
int orc_mesh_load_stl(orc_scene*, const char* path, const int perm[3], int tri_material);
/* Texture::load_png texture.rs:29 — the `image` crate's decoder is not restated: the caller decodes (e.g. PIL) and registers RGBA8 by path */
void orc_register_png(orc_scene*, const char* path, const uint8_t* rgba, uint32_t w, uint32_t h);
int orc_surface_texture_png(orc_scene*, const char* path, int wrap);
/* ObjLoader::load(path, SimpleTexturedBuilder::with_filter(wrap, groups)) obj_loader.rs:160-308, :332; groups newline-separated or NULL */
int orc_mesh_load_obj(orc_scene*, const char* path, int wrap, const char* filtered_groups);
/* ObjLoader::load(path, obj_fns(V3::new, V3::new, V2::new, |a,b,c| Triangle::with_norms_and_uvs(material,a,b,c))) obj_loader.rs:45, eve.rs:330 */
int orc_mesh_load_obj_with(orc_scene*, const char* path, int tri_material);
void orc_mesh_get_shading(orc_scene*, int mesh, float* normals9, float* uvs6, int32_t* materials);
int orc_material_info(orc_scene*, int material, float color4[4], uint32_t wh[2], uint64_t* texel_hash);
uint64_t orc_mesh_tri_count(orc_scene*, int mesh);
void orc_mesh_get_verts(orc_scene*, int mesh, float* out9);            /* 9 floats per triangle */
uint64_t orc_mesh_node_count(orc_scene*, int mesh);                    /* BvhNode count of the BLAS */

int orc_add_sphere(orc_scene*, int material, float cx, float cy, float cz, float radius);             /* World::add(Sphere::new) */
int orc_add_model(orc_scene*, int mesh, int override_material);                                         /* World::add(Model), -1 = None */
int orc_add_instance(orc_scene*, int mesh, const float t[3], const float r[3], const float s[3], int override_material); /* Model::instance().with_material() */
int orc_add_volume_sphere(orc_scene*, float cx, float cy, float cz, float radius, float density, float r, float g, float b); /* Volume::new(Sphere<()>) :603 */
int orc_add_volume_model(orc_scene*, int mesh, float density, float r, float g, float b);                                    /* Volume::new(Model) */
int orc_add_volume_instance(orc_scene*, int mesh, const float t[3], const float rot[3], const float sc[3], float density, float r, float g, float b); /* Volume::new(Instance) */
void orc_build_bvh(orc_scene*);                                        /* World::build_bvh world.rs:117 */
uint64_t orc_tlas_node_count(orc_scene*);
void orc_camera(orc_scene*, float vfov, const float from[3], const float at[3], const float up[3], float aspect, float aperture, float focus); /* Camera::new world.rs:16 */

/* introspection used by parity tests of the product's host-side builders */
void orc_get_camera(orc_scene*, float out19[19]);                      /* origin,llc,horizontal,vertical,u,v (6xV3), lens_radius */
void orc_get_instance(orc_scene*, int object, float transform16[16], float inv16[16], float aabb6[6]); /* column-major c0..c3 */
void orc_get_object_aabb(orc_scene*, int object, float aabb6[6]);

/* render  main.rs:150-295.  Buffers are row-major, row 0 = bottom (main.rs:559-564). */
/* PASS A (main.rs:166-222): pixel-centre rays, Camera::albedo_normal (world.rs:81). object/tri/t are additions
   (the reference Hit carries no ids); object = index in World::add order, tri = index in mesh or 0xFFFFFFFF. miss: object = 0xFFFFFFFF, t = inf. */
void orc_render_aov(orc_scene*, uint32_t w, uint32_t h, uint64_t seed, int threads,
                    float* albedo, float* normal, uint32_t* object, uint32_t* tri, float* t, orc_counters* counters);
/* PASS B (main.rs:233-294) bounded to `spp` merges: sum_rgb += colour, sum_bounces += MAX_DEPTH - depth. threads == 0 => max(cores-2,1)
   threads rendering whole frames like the reference; threads < 0 => the same frames distributed over those threads by image row (harness
   addition for few-spp renders of large images; per-(frame, row) random streams, output independent of the thread count). */
void orc_render(orc_scene*, uint32_t w, uint32_t h, uint32_t spp, uint32_t max_depth, uint64_t seed, int threads,
                float* sum_rgb, uint32_t* sum_bounces, orc_counters* counters);
/* Image::to_rgb_bytes(Default) + dump row flip (main.rs:640-722, 760-768): count merges -> RGB8 top row first when flip!=0 */
void orc_resolve_rgb8(const float* sum_rgb, uint32_t w, uint32_t h, uint32_t count, int flip, uint8_t* out);
/* Image::to_rgb_bytes(Albedo = 3 | Normal = 4) over a FloatBuffer (main.rs:694-721), rows reversed when flip != 0 */
void orc_float_buffer_rgb8(const float* rgb, uint32_t w, uint32_t h, int mode, int flip, uint8_t* out);

/* single-call hooks for known-answer tests */
int  orc_kat_sphere(float cx, float cy, float cz, float r, const float o[3], const float d[3], float tmin, float tmax, float out8[8]); /* t, p3, n3, front */
int  orc_kat_aabb(const float bmin[3], const float bmax[3], const float o[3], const float d[3], float tmin, float tmax);
int  orc_kat_triangle(const float v9[9], const float o[3], const float d[3], float tmin, float tmax, float out11[11]); /* t, p3, n3, front, a0,a1,a2 */
void orc_kat_rotate(int axis, float turns, float out16[16]);
void orc_kat_wrap(int mode, float x, float y, float out2[2]);
float orc_kat_reflectance(float cosine, float ref_idx);
void orc_kat_refract(const float v[3], const float n[3], float eta, float out3[3]);
void orc_kat_texture_get(orc_scene*, int surface, float u, float v, float out4[4]);
void orc_kat_background(orc_scene*, const float d[3], float out3[3]);
/* one Material::scatter call at a synthetic hit (point p, normal n already face-forwarded, front_face): returns 1 if scattered */
int  orc_kat_scatter(orc_scene*, int material, uint64_t seed, const float ray_o[3], const float ray_d[3], const float p[3], const float n[3], int front_face, float att3[3], float dir3[3]);
/* sampler draws for distribution tests */
void orc_kat_samplers(uint64_t seed, uint64_t n, float* in_sphere3, float* unit_vec3, float* in_disk2);

#ifdef __cplusplus
}
#endif
#endif

/*
 * mrt_host.h — host-side scene API above the C-ABI boundary (libmrt_host.so, plain C++17, no CUDA).
 *
 * This is the part the north star leaves with the Rust host: the reference's scene API (World::new/add/
 * build_bvh world.rs:101-122, Sphere::new geom.rs:46, Triangle::new :449, Model::new :281, Model::instance :312,
 * Instance::with_material :392, Volume::new :603, Camera::new world.rs:16, PlyLoader::load ply_loader.rs:273,
 * the materials of material.rs and surfaces of texture.rs) plus the `flatten()` walk that turns World.objects into
 * the plain arrays of mrt_scene_desc. No Rust toolchain exists in this image, so the mirror is written in C++
 * with a C surface; names, argument order and error behaviour follow the reference. Handles are small integers
 * owned by the mrth_scene. Functions returning int give a handle >= 0 or a negative error (mrth_last_error).
 */
#ifndef MRT_HOST_H
#define MRT_HOST_H
#include "mrt.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrth_scene mrth_scene;

mrth_scene* mrth_scene_new(void);        /* World::new world.rs:101; the background B is set separately (default: solid black) */
void mrth_scene_free(mrth_scene*);       /* drop(World) */
const char* mrth_last_error(mrth_scene*); /* the Err(..) / panic message of the failed call (loaders return Result, ply_loader.rs:196, obj_loader.rs:242) */
void mrth_seed(mrth_scene*, uint64_t seed); /* fastrand::seed main.rs:86 — drives BVH split axes (geom.rs:111) */
float mrth_rand_f32(mrth_scene*);           /* f32::rand() math.rs:244 for scene generation */
/* on != 0: meshes created from here on carry no reference-topology tree (mrt_blas.root = MRT_REF_NONE, n_nodes = 0) — the
 * backend builds its own BLAS from the triangle range anyway unless MRT_SCENE_KEEP_TOPOLOGY is set, so the median-split
 * build of Model::new (geom.rs:281-286) is load time with nothing to show for it. Such a scene cannot be uploaded with
 * MRT_SCENE_KEEP_TOPOLOGY, and the split-axis draws of the skipped builds are not taken from the scene's generator. */
void mrth_defer_mesh_bvh(mrth_scene*, int on);

/* texture.rs */
int mrth_surface_solid(mrth_scene*, float r, float g, float b, float a);                       /* SolidColor(V4) :180 */
int mrth_surface_texture(mrth_scene*, const uint8_t* rgba, uint32_t w, uint32_t h, int wrap); /* Texture::load_bytes :70 */
int mrth_surface_texture_png(mrth_scene*, const char* path, int wrap); /* Texture::load_png :29 — 8-bit PNG, decoded by the library itself */
int mrth_surface_ycbcr(mrth_scene*, int luma_tex, int chroma_tex);                              /* YCbCrTexture :207-223 over two Texture surfaces */
int mrth_surface_blend(mrth_scene*, int mode, int left, int right);                             /* TextureBlend::new :310; mode = BlendMode :252-275 (MRT_BLEND_*) */
int mrth_surface_fallback(mrth_scene*, float r, float g, float b, float a, int inner);          /* SolidColorFallback::new :342 */

/* material.rs */
int mrth_mat_absorb(mrth_scene*);                                     /* impl Material for () :385 */
int mrth_mat_lambertian(mrth_scene*, int surface);                    /* Lambertian::new :198 */
int mrth_mat_diffuse_light(mrth_scene*, float r, float g, float b);   /* DiffuseLight::new :233 */
int mrth_mat_metal(mrth_scene*, float fuzz, int surface);             /* Metal::new :255 (fuzz capped at 1, :256) */
int mrth_mat_dielectric(mrth_scene*, float ior);                      /* Dielectric::new :292 */
int mrth_mat_specular(mrth_scene*, float ior, int surface);           /* Specular::new :338 */
int mrth_mat_mix(mrth_scene*, float ratio, int left, int right);      /* Mix::new :398 */
int mrth_mat_isotropic(mrth_scene*, float r, float g, float b);       /* Isotrophic::new :433 */
/* EveMaterial::new (eve.rs:43-64) over three texture surfaces + EveMaterialColor {colors[4], glow} (:135-141): the implementer of Material::normal */
int mrth_mat_eve(mrth_scene*, int normal_occlusion, int albedo_roughness, int pmdg, const float colors12[12], const float glow3[3]);

/* material.rs backgrounds: the B of World<B> */
void mrth_background_solid(mrth_scene*, float r, float g, float b);   /* SolidBackground::new :44 */
void mrth_background_sky(mrth_scene*);                                /* SkyBackground :55 */
void mrth_background_skysphere(mrth_scene*, int surface);             /* SkySphere::new :70 */
void mrth_background_cubemap(mrth_scene*, const int surfaces6[6], float rx, float ry, float rz); /* CubeMap::new :102, surfaces in its argument order x+ x- y+ y- z+ z- */

/* geom.rs — a mesh is the Arc<BvhNode> of a Model; the BLAS is built at creation like Model::new */
int mrth_mesh_new(mrth_scene*, const float* verts, uint64_t n_tris, int tri_material); /* Model::new(Triangle::new(material, a, b, c)...) :281, :449 */
int mrth_mesh_new_uv(mrth_scene*, const float* verts, const float* normals, const float* uvs, uint64_t n_tris, int tri_material); /* ...Triangle::with_norms_and_uvs :468 */
int mrth_mesh_load_ply(mrth_scene*, const char* path, const int perm[3], int tri_material, float* max_abs); /* PlyLoader::load ply_loader.rs:273; perm = the axis order of the caller's vertex closure (scenes/lucy.rs:34-37: V3::new(y, z, x)), max_abs = its max_dim */
int mrth_mesh_load_stl(mrth_scene*, const char* path, const int perm[3], int tri_material); /* StlLoader::load_binary stl_loader.rs:10 */
/* ObjLoader::load(path, SimpleTexturedBuilder::with_filter(wrap, groups)) obj_loader.rs:160-308, :332 — `v/vt/vn` faces (first three
   corners), `usemtl` + `mtllib` with `Kd` / `map_Kd` -> one Lambertian per MTL material, v -> 1 - v; filtered_groups: newline-separated
   `o`/`g` names whose faces are skipped, or NULL */
int mrth_mesh_load_obj(mrth_scene*, const char* path, int wrap, const char* filtered_groups);
/* ObjLoader::load(path, obj_fns(V3::new, V3::new, V2::new, |a, b, c| Triangle::with_norms_and_uvs(material, a, b, c))) obj_loader.rs:45 (eve.rs:330) */
int mrth_mesh_load_obj_with(mrth_scene*, const char* path, int tri_material);
/* inspection (no reference counterpart: tests compare these with the oracle) */
uint64_t mrth_mesh_tri_count(mrth_scene*, int mesh);
void mrth_mesh_get_verts(mrth_scene*, int mesh, float* out9);
uint64_t mrth_mesh_node_count(mrth_scene*, int mesh);
void mrth_mesh_get_shading(mrth_scene*, int mesh, float* normals9, float* uvs6, int32_t* materials); /* per triangle; any pointer may be NULL */
/* returns the material kind (MRT_MAT_*); colour of a SolidColor surface, or size + FNV-1a hash of the f32 texels of a Texture surface */
int mrth_material_info(mrth_scene*, int material, float color4[4], uint32_t wh[2], uint64_t* texel_hash);

int mrth_add_sphere(mrth_scene*, int material, float cx, float cy, float cz, float radius);  /* world.add(Sphere::new(material, center, radius)) world.rs:112, geom.rs:47 */
int mrth_add_model(mrth_scene*, int mesh, int override_material);                           /* world.add(Model::new / Model::with_material :281, :296); override < 0 = the triangles' own */
int mrth_add_instance(mrth_scene*, int mesh, const float t[3], const float r[3], const float s[3], int override_material); /* world.add(model.instance(t, r, s)[.with_material(m)]) :312, :344, :392 */
int mrth_add_volume_sphere(mrth_scene*, float cx, float cy, float cz, float radius, float density, float r, float g, float b); /* world.add(Volume::new(Sphere::new((), c, radius), density, albedo)) :602 */
/* Volume::new over a Model / an Instance of a mesh (geom.rs:595-609 is generic over Intersect): the medium fills the mesh */
int mrth_add_volume_model(mrth_scene*, int mesh, float density, float r, float g, float b);
int mrth_add_volume_instance(mrth_scene*, int mesh, const float translation[3], const float rotation[3], const float scale[3], float density, float r, float g, float b);
void mrth_build_bvh(mrth_scene*);                /* World::build_bvh world.rs:117 */
uint64_t mrth_tlas_node_count(mrth_scene*);      /* inspection */
void mrth_camera(mrth_scene*, float vfov, const float from[3], const float at[3], const float up[3], float aspect, float aperture, float focus); /* Camera::new world.rs:16-58, same argument order */

/* inspection: the derived camera fields (world.rs:5-13), an instance's matrices and box (geom.rs:344-390), an object's box */
void mrth_get_camera(mrth_scene*, float out19[19]);
void mrth_get_instance(mrth_scene*, int object, float transform16[16], float inv16[16], float aabb6[6]);
void mrth_get_object_aabb(mrth_scene*, int object, float aabb6[6]);

/* Image export (main.rs:640-783). mrt_resolve_rgb8 covers Default / Denoise / Depth on the device image; the two FloatBuffer modes
 * work on the host buffers mrt_render_aov filled: Albedo = clamp(p, 0, 1)^(1/2.2), Normal = (p + 1) / 2, then `(f * 255) as u8`.
 * flip != 0 reverses the rows like Image::dump. mode: MRT_DISPLAY_ALBEDO or MRT_DISPLAY_NORMAL. */
int mrth_float_buffer_rgb8(const float* rgb, uint32_t w, uint32_t h, int mode, int flip, uint8_t* out_rgb);
/* Image::dump's file: creates the parent directories, writes an 8-bit RGB PNG (rows as given). */
int mrth_write_png(const char* path, const uint8_t* rgb, uint32_t w, uint32_t h);

/* the flatten() walk: valid until the scene is mutated or freed */
const mrt_scene_desc* mrth_scene_desc(mrth_scene*);
const mrt_camera* mrth_scene_camera(mrth_scene*);

#ifdef __cplusplus
}
#endif
#endif

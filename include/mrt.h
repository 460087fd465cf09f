/*
 * mrt.h — the drop-in boundary: C ABI of the B200 path-tracing backend (libmrt_cuda.so).
 *
 * The reference (nickmass/mass-raytrace, Rust) exposes no FFI; the seam this ABI replaces is the call
 *     render(image, event_proxy, world, camera, frame_limit)        src/main.rs:150-156 (called at :117)
 * Everything above that call builds (World<B>, Camera) through the scene API; everything below it is the
 * hot path (camera rays -> BVH traversal -> intersection -> material/texture shading -> accumulation).
 * A Rust host binds these entry points from an `extern "C"` block after flattening
 * World.objects into the plain arrays of mrt_scene_desc.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every call returns 0 or a negative MRT_E_*
 * code (no unwinding across the boundary; the reference relies on catch_unwind, main.rs:66);
 * mrt_last_error() gives the message. Input arrays are borrowed for the duration of the call only.
 * Image buffers are row-major with row 0 = BOTTOM of the picture (main.rs:559-564, world.rs:37-38).
 * All floats are IEEE binary32 (math.rs:12). One host thread per context at a time.
 */
#ifndef MRT_H
#define MRT_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MRT_ABI_VERSION 1

enum {
    MRT_OK = 0,
    MRT_E_INVALID = -1,     /* bad argument / malformed scene description */
    MRT_E_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    MRT_E_UNSUPPORTED = -3, /* scene uses a construct this build does not implement */
    MRT_E_STATE = -4,       /* call order (render before scene/camera upload, ...) */
    MRT_E_NOMEM = -5
};

/* ---- primitive references: kind in the top 3 bits, index in the low 29 ---------------------------- */
enum { MRT_PRIM_NODE = 0, MRT_PRIM_SPHERE = 1, MRT_PRIM_TRIANGLE = 2, MRT_PRIM_INSTANCE = 3, MRT_PRIM_VOLUME = 4 };
#define MRT_REF_NONE 0xFFFFFFFFu
#define MRT_REF(kind, index) ((((uint32_t)(kind)) << 29) | ((uint32_t)(index)))
#define MRT_REF_KIND(ref) ((ref) >> 29)
#define MRT_REF_INDEX(ref) ((ref)&0x1FFFFFFFu)

/* BvhNode {left, right, bounding_box}  geom.rs:103-107, flattened with the reference's topology
 * (random-axis median split, geom.rs:109-161). right == MRT_REF_NONE for a one-child node (:120). */
typedef struct mrt_node {
    float bmin[3];
    float bmax[3];
    uint32_t left;
    uint32_t right;
} mrt_node; /* 32 B */

/* Sphere {center, radius, material}  geom.rs:40-44 */
typedef struct mrt_sphere {
    float center[3];
    float radius;
    int32_t material;
    uint32_t object_id; /* index in World::add order, or MRT_REF_NONE for a Volume's target */
    uint32_t pad[2];
} mrt_sphere; /* 32 B */

/* Triangle shading attributes  geom.rs:435-446 (vertices live in tri_verts, 9 floats per triangle) */
#define MRT_TRI_HAS_UV 1u
typedef struct mrt_tri_shading {
    float normal[9];    /* normal_a, normal_b, normal_c */
    float uv[6];        /* uv_a, uv_b, uv_c (when MRT_TRI_HAS_UV) */
    float tangent[3];
    float bitangent[3];
    int32_t material;   /* the triangle's own material (geom.rs:441); overridden per N2 by Model/Instance */
    uint32_t flags;
    uint32_t pad;
} mrt_tri_shading; /* 96 B */

/* the Arc<BvhNode> a Model owns  geom.rs:275-278 */
typedef struct mrt_blas {
    uint32_t root;      /* prim ref of the BLAS root (a NODE); MRT_REF_NONE = no caller tree (not with MRT_SCENE_KEEP_TOPOLOGY) */
    uint32_t first_tri; /* triangles of this mesh are [first_tri, first_tri + n_tris) */
    uint32_t n_tris;
    uint32_t n_nodes;
} mrt_blas; /* 16 B */

/* Instance {triangles, material, transform, inv_transform, bounding_box}  geom.rs:335-341
 * and Model {material, triangles} geom.rs:275-278 (MRT_INSTANCE_IDENTITY: no ray transform, no normal renormalise). */
#define MRT_INSTANCE_IDENTITY 1u
typedef struct mrt_instance {
    float transform[16];     /* column-major c0,c1,c2,c3 (math/generic.rs:71-77) */
    float inv_transform[16];
    float bmin[3];           /* world AABB  geom.rs:369-381 */
    float bmax[3];
    uint32_t blas;
    int32_t material;        /* Option<M>: override, or -1 for None (geom.rs:321-323, 413-415) */
    uint32_t flags;
    uint32_t object_id;
    uint32_t pad[2];
} mrt_instance; /* 176 B */

/* Volume {neg_inv_density, target, material: Isotrophic}  geom.rs:587-591 */
typedef struct mrt_volume {
    uint32_t target;         /* prim ref of the Intersect the medium fills: a SPHERE, or an INSTANCE entry (Model or Instance; one that is not itself
                                in the world list carries object_id MRT_REF_NONE) */
    float neg_inv_density;
    int32_t material;        /* an MRT_MAT_ISOTROPIC entry */
    uint32_t object_id;
} mrt_volume; /* 16 B */

/* material.rs */
enum {
    MRT_MAT_ABSORB = 0,       /* impl Material for ()  :385 */
    MRT_MAT_LAMBERTIAN = 1,   /* surface               :192 */
    MRT_MAT_DIFFUSE_LIGHT = 2,/* p[0..3] = emit        :227 */
    MRT_MAT_METAL = 3,        /* surface, p[0] = fuzz (already min(fuzz,1)) :248 */
    MRT_MAT_DIELECTRIC = 4,   /* p[0] = refraction_index :286 */
    MRT_MAT_SPECULAR = 5,     /* surface, p[0] = refraction_index :331 */
    MRT_MAT_MIX = 6,          /* left, right, p[0] = ratio :391 */
    MRT_MAT_ISOTROPIC = 7,    /* p[0..3] = albedo      :428 */
    MRT_MAT_EVE = 8,          /* EveMaterial eve.rs:23-133, the implementer of Material::normal (tangent-space normals, geom.rs:551-560):
                                 surface = normal + occlusion texture surface, left = albedo + roughness, right = paint / material / dirt / glow
                                 (three MRT_SURF_TEXTURE surfaces); the bits of p[0] = index of the first of FOUR consecutive MRT_SURF_SOLID
                                 surfaces holding the palette: color[0..2] = EveMaterialColor.colors[i], color[3] = glow[i] (i < 3) */
    MRT_MAT_KINDS = 9
};
typedef struct mrt_material {
    int32_t kind;
    int32_t surface;
    int32_t left;
    int32_t right;
    float p[4];
} mrt_material; /* 32 B */

/* texture.rs */
enum {
    MRT_SURF_SOLID = 0,    /* color                          :179 */
    MRT_SURF_TEXTURE = 1,  /* a = texture index              :21  */
    MRT_SURF_YCBCR = 2,    /* a = luma texture, b = chroma   :207 */
    MRT_SURF_BLEND = 3,    /* a = left surface, b = right surface, mode = BlendMode :302 */
    MRT_SURF_FALLBACK = 4  /* color, a = inner surface       :336 */
};
enum { MRT_WRAP_MIRROR = 0, MRT_WRAP_REPEAT = 1, MRT_WRAP_CLAMP = 2 };
enum { MRT_BLEND_LIGHTEN = 0, MRT_BLEND_DARKEN = 1, MRT_BLEND_ADDITION = 2, MRT_BLEND_SUBTRACTION = 3 };
typedef struct mrt_surface {
    int32_t kind;
    int32_t a;
    int32_t b;
    int32_t mode;
    float color[4];
} mrt_surface; /* 32 B */
typedef struct mrt_texture {
    uint32_t width;
    uint32_t height;
    int32_t wrap;
    uint32_t pad;
    uint64_t texel_offset; /* first texel in `texels`, in RGBA-f32 texels (16 B each) */
} mrt_texture; /* 24 B */

/* material.rs:29-190 */
enum { MRT_BG_SOLID = 0, MRT_BG_SKY = 1, MRT_BG_SKYSPHERE = 2, MRT_BG_CUBEMAP = 3 };
typedef struct mrt_background {
    int32_t kind;
    int32_t surface[6];  /* SKYSPHERE: [0]; CUBEMAP: x_pos,x_neg,y_pos,y_neg,z_pos,z_neg */
    int32_t pad;
    float color[4];      /* SOLID */
    float transform[16]; /* CUBEMAP */
} mrt_background;

/* What `Arc::new(world)` (main.rs:157) holds after build_bvh (main.rs:112), as plain arrays. */
typedef struct mrt_scene_desc {
    uint32_t abi_version;
    uint32_t flags;
    const uint32_t* roots;   /* World.objects as prim refs, tested in order (world.rs:131-144); 1 entry after build_bvh */
    uint32_t n_roots;
    uint32_t n_objects;      /* number of World::add calls (object_id range) */
    const mrt_node* nodes;
    uint64_t n_nodes;
    const mrt_sphere* spheres;
    uint64_t n_spheres;
    const float* tri_verts;  /* 9 floats per triangle: vertex_a, vertex_b, vertex_c */
    const mrt_tri_shading* tri_shading;
    uint64_t n_tris;
    const mrt_blas* blas;
    uint64_t n_blas;
    const mrt_instance* instances;
    uint64_t n_instances;
    const mrt_volume* volumes;
    uint64_t n_volumes;
    const mrt_material* materials;
    uint64_t n_materials;
    const mrt_surface* surfaces;
    uint64_t n_surfaces;
    const mrt_texture* textures;
    uint64_t n_textures;
    const float* texels;     /* RGBA f32 */
    uint64_t n_texels;
    mrt_background background;
} mrt_scene_desc;

/* scene-upload flags */
#define MRT_SCENE_KEEP_TOPOLOGY 1u /* traverse the caller's BVH topology as given (default: the library may rebuild it; results do not depend on topology except exact-t ties) */

/* Camera  world.rs:5-13 — the seven derived fields, computed by the host with Camera::new arithmetic */
typedef struct mrt_camera {
    float origin[3];
    float lower_left_corner[3];
    float horizontal[3];
    float vertical[3];
    float u[3];
    float v[3];
    float lens_radius;
} mrt_camera;

typedef struct mrt_stats {
    uint64_t paths;           /* pixel samples finished (one Camera::trace call each, main.rs:261) */
    uint64_t rays;            /* scene.intersect calls from trace (world.rs:68) = sum of min(bounces+1, max_depth) */
    uint64_t node_visits;     /* inner nodes fetched (each fetch = 2 child AABBs = 64 B); counted only in instrumented renders */
    uint64_t tri_tests;
    uint64_t sphere_tests;
    uint64_t instance_tests;
    uint64_t volume_tests;
    uint64_t iterations;      /* wavefront iterations of the last render */
    uint64_t extend_launches;
    uint64_t kernel_launches; /* all kernels launched by the last render */
    float render_ms;          /* CUDA-event time of the last render on the context stream */
    float extend_ms;          /* summed CUDA-event time of the extend kernel launches (only when timing is enabled) */
    float shade_ms;
    float generate_ms;
    uint64_t scene_bytes;     /* device bytes of the uploaded scene */
    uint64_t pool_slots;      /* paths in flight (entries per ray queue) */
    uint64_t node_bytes;      /* bytes of one inner-node record of the uploaded acceleration structure (what one node visit fetches) */
    uint64_t max_ray_node_visits; /* most node visits any single ray made (instrumented renders): a ray that defeats the box tests shows here */
} mrt_stats;

typedef struct mrt_context mrt_context;

/* replaces thread-pool setup main.rs:159-170. stream = a cudaStream_t to launch on, or NULL for a private stream. */
int mrt_context_create(int device, void* stream, mrt_context** out);
void mrt_context_destroy(mrt_context* ctx);
const char* mrt_last_error(mrt_context* ctx); /* ctx may be NULL: last create() error */
int mrt_abi_version(void);


/* ---- more than one GPU ------------------------------------------------------------------------------------------------
 * The reference scales by letting every render thread trace WHOLE frames into a private buffer and adding them into the shared
 * image (thread fan-out main.rs:159-170 and :235-250, Image::merge :629-638). The B200 equivalent: the samples of every pixel
 * are split over the GPUs, each accumulates its share into its own exact (int64) image, and ONE NCCL sum-reduce over NVLink
 * merges them onto the root. Sample s of pixel p uses the same Philox stream wherever it is rendered and the sums are integers,
 * so the merged image is bit-identical to the one a single GPU renders.
 *
 * Two ways to get there, same code underneath (NCCL is loaded with dlopen("libnccl.so.2") on first use; a host that never
 * asks for more than one GPU does not need it):
 *   one process, several GPUs   mrt_context_create_multi(devices, n): the handle drives all n devices; every entry point below
 *                               works on it unchanged (scene uploads are replicated device-to-device, renders are split).
 *   one process per GPU         each rank creates its own context (mrt_context_create), rank 0 calls mrt_comm_unique_id and hands
 *                               the 128 bytes to the others by whatever means the host has, all call mrt_comm_init_rank.
 * In both cases mrt_render / mrt_render_accumulate become COLLECTIVE: every member is called with the SAME arguments, renders
 * its share mrt_sample_range(rank, size, ...) of [spp_begin, spp_begin + spp_count) and the call returns after the reduce. The
 * merged image lives on the root (rank 0 / devices[0]); the other members' images are cleared by the merge (they are the
 * threads' private buffers of main.rs:246). Output pointers may be NULL on non-root ranks. */
int mrt_context_create_multi(const int* devices, int n_devices, mrt_context** out);
#define MRT_COMM_ID_BYTES 128
int mrt_comm_unique_id(uint8_t id[MRT_COMM_ID_BYTES]);
int mrt_comm_init_rank(mrt_context* ctx, const uint8_t id[MRT_COMM_ID_BYTES], int rank, int n_ranks);
int mrt_comm_rank(mrt_context* ctx, int* rank, int* size); /* (0, 1) for a plain context; (0, n) for a multi-device handle */
/* The merge on its own, for hosts that choose the sample ranges themselves (MRT_OPT_COMM_SPLIT = 0 makes
 * mrt_render_accumulate local again): sums every member's accumulators, non-finite flags and sample count onto the root. */
int mrt_comm_reduce(mrt_context* ctx);
/* The split rule (pure function, no GPU): member `rank` of `size` renders [*begin, *begin + *count); contiguous, near-equal,
 * covering [spp_begin, spp_begin + spp_count) exactly once. */
int mrt_sample_range(int rank, int size, uint32_t spp_begin, uint32_t spp_count, uint32_t* begin, uint32_t* count);

/* replaces Arc::new(world) main.rs:157 (after world.build_bvh() main.rs:112): copies the scene to the device once */
int mrt_scene_upload(mrt_context* ctx, const mrt_scene_desc* scene);
/* The checks mrt_scene_upload makes before it touches the device, on their own: no context, no GPU. Returns MRT_OK or the code
 * the upload would fail with (index ranges of every table, tree shape under MRT_SCENE_KEEP_TOPOLOGY, unsupported features);
 * the message is copied to why[0 .. why_bytes) when why is not NULL. What Rust's type system guarantees about a World
 * (world.rs:95-122) and a flattened scene cannot: a host can run it in its own tests without a device. */
int mrt_scene_validate(const mrt_scene_desc* scene, char* why, size_t why_bytes);
/* replaces Arc::new(camera) main.rs:158 */
int mrt_camera_set(mrt_context* ctx, const mrt_camera* camera);

/* PASS A, main.rs:166-222 + Camera::albedo_normal world.rs:81-93. Pixel-centre rays u=x/(w-1), v=y/(h-1).
 * albedo/normal: w*h*3 floats (FloatBuffer, main.rs:543-575). object_id/tri_id/t (nullable) are additions the
 * reference lacks: object = World::add index, tri = index inside its mesh (MRT_REF_NONE for non-mesh), miss: object = NONE, t = +inf. */
int mrt_render_aov(mrt_context* ctx, uint32_t w, uint32_t h, uint64_t seed, float* albedo_rgb, float* normal_rgb, uint32_t* object_id,
                   uint32_t* tri_id, float* t);

/* PASS B, main.rs:233-294 + Image::merge :629-638, bounded: identical in meaning to `spp_count` merges into a cleared Image:
 * sum_rgb[(y*w+x)*3+c] = sum of sample colours, sum_bounces[y*w+x] = sum of (MAX_DEPTH - depth), *out_count = spp_count.
 * Sample s of pixel p always uses the Philox stream keyed (seed, p, s, bounce), so any split of [spp_begin, spp_begin+spp_count)
 * over calls, GPUs or processes sums to the same image bit for bit. Blocking; host buffers. */
int mrt_render(mrt_context* ctx, uint32_t w, uint32_t h, uint32_t spp_begin, uint32_t spp_count, uint32_t max_depth, uint64_t seed,
               float* sum_rgb, uint32_t* sum_bounces, uint32_t* out_count);

/* The same pass with the image kept on the device (multi-GPU: one process per GPU renders its sample range, the
 * accumulators are reduced with NCCL, rank 0 downloads). Accumulators are exact: int64 fixed point, 2^-32 units. */
int mrt_accum_reset(mrt_context* ctx, uint32_t w, uint32_t h);
int mrt_render_accumulate(mrt_context* ctx, uint32_t spp_begin, uint32_t spp_count, uint32_t max_depth, uint64_t seed);
/* device pointer to w*h*4 int64 {r,g,b (2^-32 units), bounce sum}; *n_elems = w*h*4. Sum-reducible across ranks. */
int mrt_accum_device_ptr(mrt_context* ctx, void** accum_i64, uint64_t* n_elems);
int mrt_accum_download(mrt_context* ctx, float* sum_rgb, uint32_t* sum_bounces, uint32_t* out_count);
/* Image::to_rgb_bytes(mode) main.rs:640-722 on the device image; flip != 0 reverses rows like Image::dump :763-768.
 * mode: 0 Default, 2 Depth (mode numbers follow DisplayMode, main.rs:534-541; Denoise=1 is treated as Default). out: w*h*3 bytes. */
int mrt_resolve_rgb8(mrt_context* ctx, int mode, int flip, uint32_t count, uint8_t* out_rgb);

/* DisplayMode main.rs:534-541 */
enum { MRT_DISPLAY_DEFAULT = 0, MRT_DISPLAY_DENOISE = 1, MRT_DISPLAY_DEPTH = 2, MRT_DISPLAY_ALBEDO = 3, MRT_DISPLAY_NORMAL = 4 };

/* knobs and counters (the reference only prints whole seconds, main.rs:270) */
enum {
    MRT_OPT_COUNT_VISITS = 1, /* count node / primitive visits in the next renders (instrumented kernel) */
    MRT_OPT_TIME_KERNELS = 2, /* CUDA-event time of every generate / extend / shade launch */
    MRT_OPT_POOL_SLOTS = 3,   /* paths in flight = entries per ray queue, 1024 .. 2^26 (0 = default 2^25, clamped to the size of the job; 96 B + 64 B per material kind in the scene each) */
    MRT_OPT_REFILL_LANES = 4, /* k_extend: commit and refill finished lanes of continuing rays once this many of a warp's 32 lanes are idle;
                                 0 (default) = by scene: 32 (whole batches), or 12 when a mesh's tree is deep */
    MRT_OPT_DEVICE_BUILD = 11, /* mrt_scene_upload builds trees ON THE GPU (linear BVH, ~3 ms per million primitives): 1 (default) the BLAS of meshes
                                  of >= 16384 triangles and a TLAS of >= 2^20 objects; 2 also a TLAS of >= 16384 objects; 0 never -- the host's
                                  SAH builder for everything (0.25 s per million triangles, ~13 % faster traversal of meshes, ~35 % of instance
                                  lattices): worth it for renders of thousands of samples per pixel */
    MRT_OPT_NODE_BURST = 12,   /* k_extend: at most this many node visits per lane before the warp tests its pending leaves; 0 (default) = by scene:
                                  no bound, or 4 when a mesh's tree is deep */
    MRT_OPT_COMM_SCENE = 14,   /* 1 (default): on a one-process-per-GPU communicator mrt_scene_upload is COLLECTIVE -- rank 0 validates, builds and
                                  uploads its scene once and the finished device arrays are broadcast over NVLink (the other ranks' desc is not
                                  read and may be NULL); 0: every rank uploads its own host copy */
    MRT_OPT_BVH_LEAF_TRIS = 9, /* SAH rebuild at the next mrt_scene_upload: most triangles per BLAS leaf, 1..4 (default 4) */
    MRT_OPT_BVH_TRI_COST = 10, /* SAH rebuild: cost of one triangle test in hundredths of a node visit (default 100) */
    MRT_OPT_COMM_SPLIT = 13,   /* 1 (default): on a context with a communicator mrt_render_accumulate splits the sample range and merges; 0: it renders
                                  the range it is given, locally, and the host calls mrt_comm_reduce */
    MRT_OPT_FINISH_PATHS = 8   /* drain: once no samples are left and at most this many paths are alive, one kernel runs them to the end (default 98304, 0 = off) */
};
int mrt_set_option(mrt_context* ctx, int option, uint64_t value);
int mrt_get_stats(mrt_context* ctx, mrt_stats* out);
int mrt_synchronize(mrt_context* ctx);

#ifdef __cplusplus
}
#endif
#endif

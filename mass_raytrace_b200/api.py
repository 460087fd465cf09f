"""Python mirror of the reference's scene API, and the Renderer that sits where `render()` sits.

Names and argument order follow the reference (src/world.rs, src/geom.rs, src/material.rs, src/texture.rs,
src/ply_loader.rs) so scene code reads like src/scenes/*.rs:

    world = World(SolidBackground(V3(0, 0, 0)))
    cube = Model(PlyLoader.load("cube.ply", material=()))
    world.add(cube.instance(V3(-10, 5, 0), V3(0, 0, 0), V3.fill(5.0)).with_material(Lambertian(SolidColor((1, 0, 0, 1)))))
    world.add(Sphere(Dielectric(1.3), V3(1.75, 2.0, 2.25), 2.0))
    camera = Camera(37.0, look_from, look_at, V3(0, 1, 0), aspect, aperture, focus)

The objects are plain data. `NativeScene(world, camera)` replays them into a native scene builder -- the product's
libmrt_host.so (which flattens to mrt_scene_desc), or, in tests only, the CPU oracle, which exposes the same C surface.
All geometry arithmetic (triangle normals, instance matrices, BVH build, camera frame, PLY parsing) happens natively.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _ffi


def V3(x, y, z):
    return (float(x), float(y), float(z))


def V3_fill(v):
    return (float(v), float(v), float(v))


class FastRand:
    """fastrand 1.4.1 (WyRand) as used by scene code through f32::rand() (math.rs:244) -- restated, unpinned."""

    def __init__(self, seed=1):  # fastrand::seed(1) main.rs:86
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def u64(self):
        self.s = (self.s + 0xA0761D6478BD642F) & 0xFFFFFFFFFFFFFFFF
        t = self.s * (self.s ^ 0xE7037ED1A0B428DB)
        return (t & 0xFFFFFFFFFFFFFFFF) ^ (t >> 64)

    def f32(self):
        bits = 0x3F800000 + ((self.u64() & 0xFFFFFFFF) >> 9)
        return float(np.array([bits], dtype=np.uint32).view(np.float32)[0] - np.float32(1.0))


# ---- surfaces (texture.rs) -------------------------------------------------------------------------------
WRAP_MIRROR, WRAP_REPEAT, WRAP_CLAMP = 0, 1, 2
BLEND_LIGHTEN, BLEND_DARKEN, BLEND_ADDITION, BLEND_SUBTRACTION = 0, 1, 2, 3


@dataclass(eq=False)
class SolidColor:  # texture.rs:179
    color: Tuple[float, float, float, float]


@dataclass(eq=False)
class Texture:  # texture.rs:21
    rgba: np.ndarray  # (h, w, 4) uint8
    wrapping: int = WRAP_REPEAT

    @staticmethod
    def load_bytes(data, width, height, wrapping):  # texture.rs:70
        arr = np.frombuffer(bytes(data), dtype=np.uint8).reshape(height, width, 4).copy()
        return Texture(arr, wrapping)

    @staticmethod
    def load_png(path, wrapping):  # texture.rs:29 (image crate -> to_rgba8)
        from PIL import Image

        return Texture(np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8).copy(), wrapping)


@dataclass(eq=False)
class TextureFile:  # Texture::load_png(path, wrapping) texture.rs:29, decoded natively by the scene builder
    path: str
    wrapping: int = WRAP_REPEAT


@dataclass(eq=False)
class YCbCrTexture:  # texture.rs:207
    luma: Texture
    chroma: Texture


@dataclass(eq=False)
class TextureBlend:  # texture.rs:302
    blend_mode: int
    left: object
    right: object


@dataclass(eq=False)
class SolidColorFallback:  # texture.rs:336
    color: Tuple[float, float, float, float]
    surface: object


# ---- materials (material.rs) -----------------------------------------------------------------------------
@dataclass(eq=False)
class Lambertian:  # :192
    surface: object


@dataclass(eq=False)
class DiffuseLight:  # :227
    emit: Tuple[float, float, float]


@dataclass(eq=False)
class Metal:  # :248  Metal::new(fuzz, surface)
    fuzz: float
    surface: object


@dataclass(eq=False)
class Dielectric:  # :286
    refraction_index: float


@dataclass(eq=False)
class Specular:  # :331
    refraction_index: float
    surface: object


@dataclass(eq=False)
class Mix:  # :391
    ratio: float
    left: object
    right: object


ABSORB = ()  # impl Material for ()  material.rs:385


@dataclass(eq=False)
class EveMaterial:  # EveMaterial::new(no, ar, pmdg, colors) eve.rs:43-64 -- the one implementer of Material::normal (geom.rs:551-560)
    normal_occlusion: object   # Texture surfaces (the reference loads them with WrapMode::Repeat)
    albedo_roughness: object
    pmdg: object               # paint, material, dirt, glow masks
    colors: Sequence[Tuple[float, float, float]] = ((0.02, 0.02, 0.02), (0.1, 0.1, 0.1), (0.03, 0.05, 0.1), (0.08, 0.08, 0.08))  # EveMaterialColor::caldari :160
    glow: Tuple[float, float, float] = (0.5, 0.85, 2.0)


# ---- backgrounds (material.rs:39-190) ----------------------------------------------------------------------
@dataclass(eq=False)
class SolidBackground:
    color: Tuple[float, float, float]


@dataclass(eq=False)
class SkyBackground:
    pass


@dataclass(eq=False)
class SkySphere:
    texture: object


@dataclass(eq=False)
class CubeMap:
    x_pos: object
    x_neg: object
    y_pos: object
    y_neg: object
    z_pos: object
    z_neg: object
    rotation: Tuple[float, float, float]


# ---- geometry (geom.rs) -------------------------------------------------------------------------------------
@dataclass(eq=False)
class Triangles:
    """A Vec<Triangle<M>>: either explicit vertices (n, 9) built with Triangle::new (geom.rs:449), with_norms_and_uvs (:468),
    or the result of PlyLoader::load (read natively by each backend)."""
    material: object = ABSORB
    verts: Optional[np.ndarray] = None
    normals: Optional[np.ndarray] = None
    uvs: Optional[np.ndarray] = None
    ply_path: Optional[str] = None
    ply_perm: Tuple[int, int, int] = (0, 1, 2)
    stl_path: Optional[str] = None
    obj_path: Optional[str] = None
    obj_builder: object = None


class PlyLoader:
    @staticmethod
    def load(path, vertex_perm=(0, 1, 2), material=ABSORB):
        """PlyLoader::load(path, |x,y,z| V3::new(v[perm]), |a,b,c| Triangle::new(material, a, b, c))  ply_loader.rs:273"""
        return Triangles(material=material, ply_path=str(path), ply_perm=tuple(vertex_perm))


class StlLoader:
    @staticmethod
    def load_binary(path, vertex_perm=(0, 1, 2), material=ABSORB):
        """StlLoader::load_binary(path, |x,y,z| V3::new(v[perm]), |a,b,c| Triangle::new(material, a, b, c))  stl_loader.rs:10"""
        return Triangles(material=material, stl_path=str(path), ply_perm=tuple(vertex_perm))


@dataclass(eq=False)
class SimpleTexturedBuilder:  # obj_loader.rs:160: SimpleTexturedBuilder::new(wrapping) / ::with_filter(wrapping, filtered_groups)
    wrapping: int = WRAP_REPEAT
    filtered_groups: Tuple[str, ...] = ()

    @staticmethod
    def with_filter(wrapping, filtered_groups):
        return SimpleTexturedBuilder(wrapping, tuple(filtered_groups))


@dataclass(eq=False)
class ObjFns:  # obj_fns(V3::new, V3::new, V2::new, |a, b, c| Triangle::with_norms_and_uvs(material, a, b, c))  obj_loader.rs:45, eve.rs:330
    material: object = ABSORB


class ObjLoader:
    @staticmethod
    def load(path, builder):
        """ObjLoader::load(path, builder) obj_loader.rs:332 -> Vec<Triangle<..>> (parsed natively by each backend)"""
        if not isinstance(builder, (SimpleTexturedBuilder, ObjFns)):
            raise TypeError("builder must be SimpleTexturedBuilder or ObjFns")
        material = builder.material if isinstance(builder, ObjFns) else ABSORB
        return Triangles(material=material, obj_path=str(path), obj_builder=builder)


@dataclass(eq=False)
class Sphere:  # Sphere::new(material, center, radius) geom.rs:46
    material: object
    center: Tuple[float, float, float]
    radius: float


@dataclass(eq=False)
class Instance:  # geom.rs:335
    model: "Model"
    translation: Tuple[float, float, float]
    rotation: Tuple[float, float, float]  # in turns
    scale: Tuple[float, float, float]
    material: object = None

    def with_material(self, material):  # geom.rs:392
        return Instance(self.model, self.translation, self.rotation, self.scale, material)


@dataclass(eq=False)
class Model:  # geom.rs:275
    triangles: Triangles
    material: object = None  # Option<M>

    @staticmethod
    def with_material(material, triangles):  # geom.rs:294
        return Model(triangles, material)

    def instance(self, translation, rotation, scale):  # geom.rs:312 -- drops the model's own material
        return Instance(self, tuple(translation), tuple(rotation), tuple(scale), None)


@dataclass(eq=False)
class Volume:  # Volume::new(target, density, albedo) geom.rs:603; target: a Sphere, a Model or an Instance (the medium fills it)
    target: object
    density: float
    albedo: Tuple[float, float, float]


@dataclass(eq=False)
class Camera:  # Camera::new world.rs:16
    vertical_fov: float
    look_from: Tuple[float, float, float]
    look_at: Tuple[float, float, float]
    view_up: Tuple[float, float, float]
    aspect_ratio: float
    aperture: float
    focus_distance: float


class World:  # world.rs:96
    def __init__(self, background, bvh_seed=1):
        self.background = background
        self.objects = []
        self.bvh_built = False
        self.bvh_seed = bvh_seed  # fastrand::seed(1) main.rs:86 drives the split axes (geom.rs:111)

    def clear(self):
        self.objects.clear()
        self.bvh_built = False

    def add(self, obj):
        if not isinstance(obj, (Sphere, Instance, Model, Volume)):
            raise TypeError(f"World.add: unsupported object {type(obj).__name__}")
        self.objects.append(obj)

    def build_bvh(self):  # world.rs:117
        self.bvh_built = True


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class NativeScene:
    """Replays a (World, Camera) into a native scene builder. backend = (ctypes lib, symbol prefix).

    defer_mesh_bvh: meshes carry no reference-topology tree (mrth_defer_mesh_bvh, include/mrt_host.h) — the load-time
    saving for a scene that is rendered without MRT_SCENE_KEEP_TOPOLOGY; host library only."""

    def __init__(self, world: World, camera: Optional[Camera] = None, backend=None, defer_mesh_bvh: bool = False):
        if backend is None:
            backend = (_ffi.host_lib(), "mrth")
        self.lib, self.prefix = backend
        self._h = self._fn("scene_new")()
        if not self._h:
            raise MemoryError("scene_new failed")
        self._surf = {}
        self._mat = {}
        self._mesh = {}
        self.mesh_max_abs = {}
        self.object_ids = []
        self._fn("seed")(self._h, world.bvh_seed)
        if defer_mesh_bvh:
            self._fn("defer_mesh_bvh")(self._h, 1)
        self._background(world.background)
        for obj in world.objects:
            self.object_ids.append(self._add(obj))
        if world.bvh_built:
            self._fn("build_bvh")(self._h)
        if camera is not None:
            self._fn("camera")(self._h, camera.vertical_fov, _f3(camera.look_from), _f3(camera.look_at), _f3(camera.view_up),
                               camera.aspect_ratio, camera.aperture, camera.focus_distance)
        self.has_camera = camera is not None

    def _fn(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    def _check(self, rc, what):
        if rc < 0:
            raise RuntimeError(f"{what}: {self._fn('last_error')(self._h).decode(errors="replace")}")
        return rc

    def close(self):
        if self._h:
            self._fn("scene_free")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- realisation -------------------------------------------------------------------------------------
    def _surface(self, s):
        key = id(s)
        if key in self._surf:
            return self._surf[key]
        if isinstance(s, SolidColor):
            h = self._fn("surface_solid")(self._h, *[float(x) for x in s.color])
        elif isinstance(s, Texture):
            arr = np.ascontiguousarray(s.rgba, dtype=np.uint8)
            h = self._fn("surface_texture")(self._h, arr.ctypes.data_as(_ffi.u8p), arr.shape[1], arr.shape[0], s.wrapping)
        elif isinstance(s, TextureFile):
            self._prepare_png(s.path)
            h = self._fn("surface_texture_png")(self._h, s.path.encode(), s.wrapping)
        elif isinstance(s, YCbCrTexture):
            h = self._fn("surface_ycbcr")(self._h, self._surface(s.luma), self._surface(s.chroma))
        elif isinstance(s, TextureBlend):
            h = self._fn("surface_blend")(self._h, s.blend_mode, self._surface(s.left), self._surface(s.right))
        elif isinstance(s, SolidColorFallback):
            h = self._fn("surface_fallback")(self._h, *[float(x) for x in s.color], self._surface(s.surface))
        else:
            raise TypeError(f"unsupported surface {type(s).__name__}")
        self._surf[key] = self._check(h, "surface")
        return h

    def _material(self, m):
        if m is None:
            return -1
        key = id(m) if m != () else "unit"
        if key in self._mat:
            return self._mat[key]
        if m == ():
            h = self._fn("mat_absorb")(self._h)
        elif isinstance(m, Lambertian):
            h = self._fn("mat_lambertian")(self._h, self._surface(m.surface))
        elif isinstance(m, DiffuseLight):
            h = self._fn("mat_diffuse_light")(self._h, *[float(x) for x in m.emit])
        elif isinstance(m, Metal):
            h = self._fn("mat_metal")(self._h, float(m.fuzz), self._surface(m.surface))
        elif isinstance(m, Dielectric):
            h = self._fn("mat_dielectric")(self._h, float(m.refraction_index))
        elif isinstance(m, Specular):
            h = self._fn("mat_specular")(self._h, float(m.refraction_index), self._surface(m.surface))
        elif isinstance(m, Mix):
            h = self._fn("mat_mix")(self._h, float(m.ratio), self._material(m.left), self._material(m.right))
        elif isinstance(m, EveMaterial):
            colors = (C.c_float * 12)(*[float(x) for c in m.colors for x in c])
            glow = (C.c_float * 3)(*[float(x) for x in m.glow])
            h = self._fn("mat_eve")(self._h, self._surface(m.normal_occlusion), self._surface(m.albedo_roughness), self._surface(m.pmdg), C.byref(colors), C.byref(glow))
        else:
            raise TypeError(f"unsupported material {type(m).__name__}")
        self._mat[key] = self._check(h, "material")
        return h

    def _background(self, b):
        if isinstance(b, SolidBackground):
            self._fn("background_solid")(self._h, *[float(x) for x in b.color])
        elif isinstance(b, SkyBackground):
            self._fn("background_sky")(self._h)
        elif isinstance(b, SkySphere):
            self._fn("background_skysphere")(self._h, self._surface(b.texture))
        elif isinstance(b, CubeMap):
            faces = (C.c_int * 6)(*[self._surface(s) for s in (b.x_pos, b.x_neg, b.y_pos, b.y_neg, b.z_pos, b.z_neg)])
            self._fn("background_cubemap")(self._h, C.byref(faces), *[float(x) for x in b.rotation])
        else:
            raise TypeError(f"unsupported background {type(b).__name__}")

    def _prepare_png(self, path):
        """Hook for backends that do not decode PNG themselves (the tests' CPU checker); the product library does."""

    def _prepare_obj(self, path):
        """Same hook for the PNG files an OBJ's material library may name."""

    def mesh(self, tris: Triangles):
        key = id(tris)
        if key in self._mesh:
            return self._mesh[key]
        mat = self._material(tris.material)
        if tris.obj_path is not None:
            self._prepare_obj(tris.obj_path)
            b = tris.obj_builder
            if isinstance(b, SimpleTexturedBuilder):
                groups = "\n".join(b.filtered_groups).encode() if b.filtered_groups else None
                h = self._fn("mesh_load_obj")(self._h, tris.obj_path.encode(), b.wrapping, groups)
            else:
                h = self._fn("mesh_load_obj_with")(self._h, tris.obj_path.encode(), mat)
            self._check(h, f"ObjLoader.load({tris.obj_path})")
        elif tris.ply_path is not None:
            perm = (C.c_int * 3)(*tris.ply_perm)
            max_abs = C.c_float(0)
            h = self._fn("mesh_load_ply")(self._h, tris.ply_path.encode(), C.byref(perm), mat, C.byref(max_abs))
            self._check(h, f"PlyLoader.load({tris.ply_path})")
            self.mesh_max_abs[key] = max_abs.value
        elif tris.stl_path is not None:
            perm = (C.c_int * 3)(*tris.ply_perm)
            h = self._fn("mesh_load_stl")(self._h, tris.stl_path.encode(), C.byref(perm), mat)
            self._check(h, f"StlLoader.load_binary({tris.stl_path})")
        else:
            v = np.ascontiguousarray(tris.verts, dtype=np.float32).reshape(-1, 9)
            if tris.uvs is not None:
                n = np.ascontiguousarray(tris.normals, dtype=np.float32).reshape(-1, 9)
                uv = np.ascontiguousarray(tris.uvs, dtype=np.float32).reshape(-1, 6)
                h = self._fn("mesh_new_uv")(self._h, v.ctypes.data_as(_ffi.f32p), n.ctypes.data_as(_ffi.f32p), uv.ctypes.data_as(_ffi.f32p), v.shape[0], mat)
            else:
                h = self._fn("mesh_new")(self._h, v.ctypes.data_as(_ffi.f32p), v.shape[0], mat)
            self._check(h, "Model::new")
        self._mesh[key] = h
        return h

    def _add(self, obj):
        if isinstance(obj, Sphere):
            h = self._fn("add_sphere")(self._h, self._material(obj.material), *[float(x) for x in obj.center], float(obj.radius))
        elif isinstance(obj, Model):
            h = self._fn("add_model")(self._h, self.mesh(obj.triangles), self._material(obj.material))
        elif isinstance(obj, Instance):
            h = self._fn("add_instance")(self._h, self.mesh(obj.model.triangles), _f3(obj.translation), _f3(obj.rotation), _f3(obj.scale),
                                         self._material(obj.material))
        elif isinstance(obj, Volume):
            tg, alb = obj.target, [float(x) for x in obj.albedo]
            if isinstance(tg, Sphere):
                c = tg.center
                h = self._fn("add_volume_sphere")(self._h, float(c[0]), float(c[1]), float(c[2]), float(tg.radius), float(obj.density), *alb)
            elif isinstance(tg, Model):
                h = self._fn("add_volume_model")(self._h, self.mesh(tg.triangles), float(obj.density), *alb)
            elif isinstance(tg, Instance):
                h = self._fn("add_volume_instance")(self._h, self.mesh(tg.model.triangles), _f3(tg.translation), _f3(tg.rotation), _f3(tg.scale), float(obj.density), *alb)
            else:
                raise TypeError(f"Volume over {type(tg).__name__}")
        else:
            raise TypeError(type(obj).__name__)
        return self._check(h, "World::add")

    # -- introspection (parity tests of the host builders) ----------------------------------------------
    def camera_fields(self):
        out = np.zeros(19, dtype=np.float32)
        self._fn("get_camera")(self._h, out.ctypes.data_as(_ffi.f32p))
        return out

    def instance_fields(self, object_id):
        tf, inv, aabb = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(6, np.float32)
        self._fn("get_instance")(self._h, object_id, tf.ctypes.data_as(_ffi.f32p), inv.ctypes.data_as(_ffi.f32p), aabb.ctypes.data_as(_ffi.f32p))
        return tf, inv, aabb

    def object_aabb(self, object_id):
        aabb = np.zeros(6, np.float32)
        self._fn("get_object_aabb")(self._h, object_id, aabb.ctypes.data_as(_ffi.f32p))
        return aabb

    def mesh_verts(self, tris: Triangles):
        m = self.mesh(tris)
        n = self._fn("mesh_tri_count")(self._h, m)
        out = np.zeros((n, 9), dtype=np.float32)
        self._fn("mesh_get_verts")(self._h, m, out.ctypes.data_as(_ffi.f32p))
        return out

    def mesh_shading(self, tris: Triangles):
        """Per-triangle vertex normals (n, 9), uvs (n, 6) and material handles (n,) of a mesh."""
        m = self.mesh(tris)
        n = self._fn("mesh_tri_count")(self._h, m)
        nrm, uv, mat = np.zeros((n, 9), np.float32), np.zeros((n, 6), np.float32), np.zeros(n, np.int32)
        self._fn("mesh_get_shading")(self._h, m, nrm.ctypes.data_as(_ffi.f32p), uv.ctypes.data_as(_ffi.f32p), mat.ctypes.data_as(C.POINTER(C.c_int32)))
        return nrm, uv, mat

    def material_info(self, material):
        """(kind, colour of a SolidColor surface, (w, h) and texel hash of a Texture surface) of a material handle."""
        color, wh, hsh = (C.c_float * 4)(), (C.c_uint32 * 2)(), C.c_uint64(0)
        kind = self._fn("material_info")(self._h, int(material), color, wh, C.byref(hsh))
        return kind, tuple(color), tuple(wh), hsh.value

    def mesh_node_count(self, tris: Triangles):
        return self._fn("mesh_node_count")(self._h, self.mesh(tris))

    def tlas_node_count(self):
        return self._fn("tlas_node_count")(self._h)

    # -- product only: the flattened scene ---------------------------------------------------------------
    def desc(self):
        return self._fn("scene_desc")(self._h)

    def camera_struct(self):
        return self._fn("scene_camera")(self._h)


class MrtError(RuntimeError):
    pass


class Renderer:
    """One GPU behind the C ABI of include/mrt.h. Replaces render() (main.rs:150-295); no CPU fallback."""

    OPT_COUNT_VISITS, OPT_TIME_KERNELS, OPT_POOL_SLOTS, OPT_REFILL_LANES, OPT_FINISH_PATHS = 1, 2, 3, 4, 8
    OPT_BVH_LEAF_TRIS, OPT_BVH_TRI_COST, OPT_DEVICE_BUILD, OPT_NODE_BURST, OPT_COMM_SPLIT, OPT_COMM_SCENE = 9, 10, 11, 12, 13, 14

    def __init__(self, device=0, stream=None):
        """device: one CUDA device index, or a list of them -- then the handle drives all of them in this process
        (mrt_context_create_multi: renders are split by samples and merged with one NCCL reduce)."""
        self.lib = _ffi.cuda_lib()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self.lib.mrt_context_create_multi(devs, len(device), C.byref(h))
            what = f"mrt_context_create_multi({list(device)})"
        else:
            rc = self.lib.mrt_context_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
            what = f"mrt_context_create({device})"
        if rc != 0:
            raise MrtError(f"{what} = {rc}: {self.lib.mrt_last_error(None).decode(errors="replace")}")
        self._h = h
        self.size = None
        self._scene_keepalive = None

    # ---- one process per GPU: rank 0 makes the id, every rank joins (the host moves the 128 bytes however it likes) ----
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        lib = _ffi.cuda_lib()
        rc = lib.mrt_comm_unique_id(buf)
        if rc != 0:
            raise MrtError(f"mrt_comm_unique_id = {rc}: {lib.mrt_last_error(None).decode(errors="replace")}")
        return bytes(buf)

    def comm_init_rank(self, unique_id, rank, n_ranks):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.lib.mrt_comm_init_rank(self._h, buf, int(rank), int(n_ranks)), "mrt_comm_init_rank")

    def comm_rank(self):
        r, n = C.c_int(0), C.c_int(1)
        self._check(self.lib.mrt_comm_rank(self._h, C.byref(r), C.byref(n)), "mrt_comm_rank")
        return r.value, n.value

    def comm_reduce(self):
        """Image::merge over the members (main.rs:629-638): one NCCL sum-reduce onto rank 0."""
        self._check(self.lib.mrt_comm_reduce(self._h), "mrt_comm_reduce")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mrt_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise MrtError(f"{what} = {rc}: {self.lib.mrt_last_error(self._h).decode(errors="replace")}")

    def set_scene(self, scene: NativeScene, keep_topology=False):
        """mrt_scene_upload (+ mrt_camera_set). keep_topology: traverse the caller's BVH as built by BvhNode::new (geom.rs:109-161)
        instead of the library's SAH rebuild (MRT_SCENE_KEEP_TOPOLOGY); results are the same except on exact-t ties."""
        desc = scene.desc()
        desc.contents.flags = 1 if keep_topology else 0
        self._check(self.lib.mrt_scene_upload(self._h, desc), "mrt_scene_upload")
        if scene.has_camera:
            self._check(self.lib.mrt_camera_set(self._h, scene.camera_struct()), "mrt_camera_set")

    def set_option(self, option, value):
        self._check(self.lib.mrt_set_option(self._h, option, int(value)), "mrt_set_option")

    def render_aov(self, w, h, seed=1):
        n = w * h
        out = dict(albedo=np.zeros((h, w, 3), np.float32), normal=np.zeros((h, w, 3), np.float32), object=np.zeros((h, w), np.uint32),
                   tri=np.zeros((h, w), np.uint32), t=np.zeros((h, w), np.float32))
        assert out["albedo"].size == n * 3
        self._check(self.lib.mrt_render_aov(self._h, w, h, seed, out["albedo"].ctypes.data_as(_ffi.f32p), out["normal"].ctypes.data_as(_ffi.f32p),
                                            out["object"].ctypes.data_as(_ffi.u32p), out["tri"].ctypes.data_as(_ffi.u32p),
                                            out["t"].ctypes.data_as(_ffi.f32p)), "mrt_render_aov")
        return out

    def render(self, w, h, spp, max_depth=50, seed=1, spp_begin=0, out=None):
        """mrt_render: host buffers in, blocking. Returns (sum_rgb (h,w,3) f32, sum_bounces (h,w) u32, count)."""
        if out is None:
            out = (np.zeros((h, w, 3), np.float32), np.zeros((h, w), np.uint32))
        cnt = C.c_uint32(0)
        self._check(self.lib.mrt_render(self._h, w, h, spp_begin, spp, max_depth, seed, out[0].ctypes.data_as(_ffi.f32p),
                                        out[1].ctypes.data_as(_ffi.u32p), C.byref(cnt)), "mrt_render")
        self.size = (w, h)
        return out[0], out[1], cnt.value

    def reset(self, w, h):
        self._check(self.lib.mrt_accum_reset(self._h, w, h), "mrt_accum_reset")
        self.size = (w, h)

    def accumulate(self, spp_begin, spp_count, max_depth=50, seed=1):
        self._check(self.lib.mrt_render_accumulate(self._h, spp_begin, spp_count, max_depth, seed), "mrt_render_accumulate")

    def accum_device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.lib.mrt_accum_device_ptr(self._h, C.byref(p), C.byref(n)), "mrt_accum_device_ptr")
        return p.value, n.value

    def download(self):
        w, h = self.size
        rgb, b = np.zeros((h, w, 3), np.float32), np.zeros((h, w), np.uint32)
        cnt = C.c_uint32(0)
        self._check(self.lib.mrt_accum_download(self._h, rgb.ctypes.data_as(_ffi.f32p), b.ctypes.data_as(_ffi.u32p), C.byref(cnt)), "mrt_accum_download")
        return rgb, b, cnt.value

    def resolve_rgb8(self, count, mode=0, flip=True):
        w, h = self.size
        out = np.zeros((h, w, 3), np.uint8)
        self._check(self.lib.mrt_resolve_rgb8(self._h, mode, 1 if flip else 0, count, out.ctypes.data_as(_ffi.u8p)), "mrt_resolve_rgb8")
        return out

    def stats(self):
        st = _ffi.mrt_stats()
        self._check(self.lib.mrt_get_stats(self._h, C.byref(st)), "mrt_get_stats")
        return st.as_dict()

    def synchronize(self):
        self._check(self.lib.mrt_synchronize(self._h), "mrt_synchronize")


DISPLAY_DEFAULT, DISPLAY_DENOISE, DISPLAY_DEPTH, DISPLAY_ALBEDO, DISPLAY_NORMAL = range(5)  # DisplayMode main.rs:534-541


def float_buffer_rgb8(buf, mode, flip=True):
    """Image::to_rgb_bytes(Albedo | Normal) over one of render_aov's FloatBuffers (main.rs:694-721); flip reverses rows like dump()."""
    a = np.ascontiguousarray(buf, dtype=np.float32)
    h, w = a.shape[:2]
    out = np.zeros((h, w, 3), np.uint8)
    rc = _ffi.host_lib().mrth_float_buffer_rgb8(a.ctypes.data_as(_ffi.f32p), w, h, mode, 1 if flip else 0, out.ctypes.data_as(_ffi.u8p))
    if rc < 0:
        raise ValueError(f"mrth_float_buffer_rgb8: error {rc} (mode must be DISPLAY_ALBEDO or DISPLAY_NORMAL)")
    return out


def write_png(path, rgb8):
    """Image::dump's file (main.rs:770-783): parent directories created, 8-bit RGB PNG, rows as given."""
    a = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w = a.shape[:2]
    if a.shape != (h, w, 3):
        raise ValueError("write_png wants an (h, w, 3) uint8 image")
    if _ffi.host_lib().mrth_write_png(os.fsencode(path), a.ctypes.data_as(_ffi.u8p), w, h) < 0:
        raise OSError(f"unable to save image {path}")


def render(world: World, camera: Camera, width, height, spp, max_depth=50, seed=1, device=0):
    """The reference's render(image, .., world, camera, frame_limit) (main.rs:150): AOV pre-pass then `spp` merges."""
    scene = NativeScene(world, camera)
    r = Renderer(device)
    try:
        r.set_scene(scene)
        aov = r.render_aov(width, height, seed)
        rgb, bounces, count = r.render(width, height, spp, max_depth, seed)
        return dict(sum_rgb=rgb, sum_bounces=bounces, count=count, albedo=aov["albedo"], normal=aov["normal"], stats=r.stats())
    finally:
        r.close()
        scene.close()

"""ctypes bindings for the two product libraries.

libmrt_cuda.so  -- the drop-in boundary, include/mrt.h (+ test hooks of include/mrt_debug.h)
libmrt_host.so  -- host-side scene API + flatten(), include/mrt_host.h

There is no CPU fallback: if libmrt_cuda.so is missing, loading raises; if no B200 is visible,
mrt_context_create() fails and Renderer() raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.environ.get("MRT_CUDA_LIB", os.path.join(_HERE, "libmrt_cuda.so"))  # override only for kernel-variant experiments
HOST_LIB_PATH = os.path.join(_HERE, "libmrt_host.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)


class mrt_background(C.Structure):
    _fields_ = [("kind", C.c_int32), ("surface", C.c_int32 * 6), ("pad", C.c_int32), ("color", C.c_float * 4), ("transform", C.c_float * 16)]


class mrt_scene_desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("flags", C.c_uint32),
        ("roots", u32p), ("n_roots", C.c_uint32), ("n_objects", C.c_uint32),
        ("nodes", C.c_void_p), ("n_nodes", C.c_uint64),
        ("spheres", C.c_void_p), ("n_spheres", C.c_uint64),
        ("tri_verts", f32p), ("tri_shading", C.c_void_p), ("n_tris", C.c_uint64),
        ("blas", C.c_void_p), ("n_blas", C.c_uint64),
        ("instances", C.c_void_p), ("n_instances", C.c_uint64),
        ("volumes", C.c_void_p), ("n_volumes", C.c_uint64),
        ("materials", C.c_void_p), ("n_materials", C.c_uint64),
        ("surfaces", C.c_void_p), ("n_surfaces", C.c_uint64),
        ("textures", C.c_void_p), ("n_textures", C.c_uint64),
        ("texels", f32p), ("n_texels", C.c_uint64),
        ("background", mrt_background),
    ]


class mrt_camera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("horizontal", C.c_float * 3), ("vertical", C.c_float * 3),
                ("u", C.c_float * 3), ("v", C.c_float * 3), ("lens_radius", C.c_float)]


class mrt_stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("instance_tests", C.c_uint64), ("volume_tests", C.c_uint64), ("iterations", C.c_uint64), ("extend_launches", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("render_ms", C.c_float), ("extend_ms", C.c_float), ("shade_ms", C.c_float), ("generate_ms", C.c_float),
                ("scene_bytes", C.c_uint64), ("pool_slots", C.c_uint64), ("node_bytes", C.c_uint64), ("max_ray_node_visits", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# name -> (restype, argtypes); the single source of truth that tests check against include/*.h
CUDA_API = {
    "mrt_abi_version": (C.c_int, []),
    "mrt_context_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mrt_context_create_multi": (C.c_int, [i32p, C.c_int, C.POINTER(C.c_void_p)]),
    "mrt_comm_unique_id": (C.c_int, [u8p]),
    "mrt_comm_init_rank": (C.c_int, [C.c_void_p, u8p, C.c_int, C.c_int]),
    "mrt_comm_rank": (C.c_int, [C.c_void_p, i32p, i32p]),
    "mrt_comm_reduce": (C.c_int, [C.c_void_p]),
    "mrt_sample_range": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, u32p, u32p]),
    "mrt_context_destroy": (None, [C.c_void_p]),
    "mrt_last_error": (C.c_char_p, [C.c_void_p]),
    "mrt_scene_upload": (C.c_int, [C.c_void_p, C.POINTER(mrt_scene_desc)]),
    "mrt_scene_validate": (C.c_int, [C.POINTER(mrt_scene_desc), C.c_char_p, C.c_size_t]),
    "mrt_camera_set": (C.c_int, [C.c_void_p, C.POINTER(mrt_camera)]),
    "mrt_render_aov": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, f32p, f32p, u32p, u32p, f32p]),
    "mrt_render": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, f32p, u32p, u32p]),
    "mrt_accum_reset": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "mrt_render_accumulate": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]),
    "mrt_accum_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "mrt_accum_download": (C.c_int, [C.c_void_p, f32p, u32p, u32p]),
    "mrt_resolve_rgb8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint32, u8p]),
    "mrt_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64]),
    "mrt_get_stats": (C.c_int, [C.c_void_p, C.POINTER(mrt_stats)]),
    "mrt_synchronize": (C.c_int, [C.c_void_p]),
    "mrt_debug_philox": (C.c_int, [C.c_void_p, u32p, u32p, u32p]),
    "mrt_debug_samplers": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, f32p, f32p, f32p]),
}


def scene_api(prefix, with_desc):
    """Signatures of the scene-builder C surface of libmrt_host.so (prefix `mrth`). The table is a function of the prefix so that
    tests can bind a second implementation of the same surface (their CPU checker) and replay identical scenes into both."""
    p = prefix
    f3 = C.c_float * 3
    api = {
        f"{p}_scene_new": (C.c_void_p, []),
        f"{p}_scene_free": (None, [C.c_void_p]),
        f"{p}_last_error": (C.c_char_p, [C.c_void_p]),
        f"{p}_seed": (None, [C.c_void_p, C.c_uint64]),
        f"{p}_rand_f32": (C.c_float, [C.c_void_p]),
        f"{p}_surface_solid": (C.c_int, [C.c_void_p] + [C.c_float] * 4),
        f"{p}_surface_texture": (C.c_int, [C.c_void_p, u8p, C.c_uint32, C.c_uint32, C.c_int]),
        f"{p}_surface_texture_png": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
        f"{p}_surface_ycbcr": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
        f"{p}_surface_blend": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
        f"{p}_surface_fallback": (C.c_int, [C.c_void_p] + [C.c_float] * 4 + [C.c_int]),
        f"{p}_mat_absorb": (C.c_int, [C.c_void_p]),
        f"{p}_mat_lambertian": (C.c_int, [C.c_void_p, C.c_int]),
        f"{p}_mat_diffuse_light": (C.c_int, [C.c_void_p] + [C.c_float] * 3),
        f"{p}_mat_metal": (C.c_int, [C.c_void_p, C.c_float, C.c_int]),
        f"{p}_mat_dielectric": (C.c_int, [C.c_void_p, C.c_float]),
        f"{p}_mat_specular": (C.c_int, [C.c_void_p, C.c_float, C.c_int]),
        f"{p}_mat_mix": (C.c_int, [C.c_void_p, C.c_float, C.c_int, C.c_int]),
        f"{p}_mat_isotropic": (C.c_int, [C.c_void_p] + [C.c_float] * 3),
        f"{p}_mat_eve": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float * 12), C.POINTER(C.c_float * 3)]),
        f"{p}_background_solid": (None, [C.c_void_p] + [C.c_float] * 3),
        f"{p}_background_sky": (None, [C.c_void_p]),
        f"{p}_background_skysphere": (None, [C.c_void_p, C.c_int]),
        f"{p}_background_cubemap": (None, [C.c_void_p, C.POINTER(C.c_int * 6)] + [C.c_float] * 3),
        f"{p}_mesh_new": (C.c_int, [C.c_void_p, f32p, C.c_uint64, C.c_int]),
        f"{p}_mesh_new_uv": (C.c_int, [C.c_void_p, f32p, f32p, f32p, C.c_uint64, C.c_int]),
        f"{p}_mesh_load_ply": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int * 3), C.c_int, f32p]),
        f"{p}_mesh_load_stl": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int * 3), C.c_int]),
        f"{p}_mesh_load_obj": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p]),
        f"{p}_mesh_load_obj_with": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
        f"{p}_mesh_get_shading": (None, [C.c_void_p, C.c_int, f32p, f32p, C.POINTER(C.c_int32)]),
        f"{p}_material_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float * 4), C.POINTER(C.c_uint32 * 2), C.POINTER(C.c_uint64)]),
        f"{p}_mesh_tri_count": (C.c_uint64, [C.c_void_p, C.c_int]),
        f"{p}_mesh_get_verts": (None, [C.c_void_p, C.c_int, f32p]),
        f"{p}_mesh_node_count": (C.c_uint64, [C.c_void_p, C.c_int]),
        f"{p}_add_sphere": (C.c_int, [C.c_void_p, C.c_int] + [C.c_float] * 4),
        f"{p}_add_model": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
        f"{p}_add_instance": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(f3), C.POINTER(f3), C.POINTER(f3), C.c_int]),
        f"{p}_add_volume_sphere": (C.c_int, [C.c_void_p] + [C.c_float] * 8),
        f"{p}_add_volume_model": (C.c_int, [C.c_void_p, C.c_int] + [C.c_float] * 4),
        f"{p}_add_volume_instance": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(f3), C.POINTER(f3), C.POINTER(f3)] + [C.c_float] * 4),
        f"{p}_build_bvh": (None, [C.c_void_p]),
        f"{p}_tlas_node_count": (C.c_uint64, [C.c_void_p]),
        f"{p}_camera": (None, [C.c_void_p, C.c_float, C.POINTER(f3), C.POINTER(f3), C.POINTER(f3), C.c_float, C.c_float, C.c_float]),
        f"{p}_get_camera": (None, [C.c_void_p, f32p]),
        f"{p}_get_instance": (None, [C.c_void_p, C.c_int, f32p, f32p, f32p]),
        f"{p}_get_object_aabb": (None, [C.c_void_p, C.c_int, f32p]),
    }
    if with_desc:
        api[f"{p}_defer_mesh_bvh"] = (None, [C.c_void_p, C.c_int])
        api[f"{p}_scene_desc"] = (C.POINTER(mrt_scene_desc), [C.c_void_p])
        api[f"{p}_scene_camera"] = (C.POINTER(mrt_camera), [C.c_void_p])
    return api


HOST_API = scene_api("mrth", with_desc=True)
HOST_API["mrth_float_buffer_rgb8"] = (C.c_int, [f32p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, u8p])
HOST_API["mrth_write_png"] = (C.c_int, [C.c_char_p, u8p, C.c_uint32, C.c_uint32])


def bind(lib, api):
    for name, (res, args) in api.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


_cuda = None
_host = None


def cuda_lib():
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB_PATH):
            raise RuntimeError(f"{CUDA_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(this backend has no CPU fallback)")
        _cuda = bind(C.CDLL(CUDA_LIB_PATH), CUDA_API)
    return _cuda


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        _host = bind(C.CDLL(HOST_LIB_PATH), HOST_API)
    return _host

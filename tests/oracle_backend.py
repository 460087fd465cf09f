"""Loads the CPU oracle (oracle/liboracle.so, test infrastructure) behind the same scene-builder surface as libmrt_host.so."""
import ctypes as C
import os

import numpy as np

from mass_raytrace_b200 import _ffi
from mass_raytrace_b200.api import NativeScene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")


class orc_counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "paths", "box_tests", "tri_tests", "sphere_tests", "instance_tests", "volume_tests")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


f3 = C.c_float * 3
_EXTRA = {
    "orc_render_aov": (None, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, _ffi.f32p, _ffi.f32p, _ffi.u32p, _ffi.u32p, _ffi.f32p, C.POINTER(orc_counters)]),
    "orc_render": (None, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, _ffi.f32p, _ffi.u32p, C.POINTER(orc_counters)]),
    "orc_resolve_rgb8": (None, [_ffi.f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _ffi.u8p]),
    "orc_float_buffer_rgb8": (None, [_ffi.f32p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _ffi.u8p]),
    "orc_register_png": (None, [C.c_void_p, C.c_char_p, _ffi.u8p, C.c_uint32, C.c_uint32]),
    "orc_kat_sphere": (C.c_int, [C.c_float] * 4 + [C.POINTER(f3), C.POINTER(f3), C.c_float, C.c_float, _ffi.f32p]),
    "orc_kat_aabb": (C.c_int, [C.POINTER(f3)] * 4 + [C.c_float, C.c_float]),
    "orc_kat_triangle": (C.c_int, [_ffi.f32p, C.POINTER(f3), C.POINTER(f3), C.c_float, C.c_float, _ffi.f32p]),
    "orc_kat_rotate": (None, [C.c_int, C.c_float, _ffi.f32p]),
    "orc_kat_wrap": (None, [C.c_int, C.c_float, C.c_float, _ffi.f32p]),
    "orc_kat_reflectance": (C.c_float, [C.c_float, C.c_float]),
    "orc_kat_refract": (None, [C.POINTER(f3), C.POINTER(f3), C.c_float, _ffi.f32p]),
    "orc_kat_texture_get": (None, [C.c_void_p, C.c_int, C.c_float, C.c_float, _ffi.f32p]),
    "orc_kat_background": (None, [C.c_void_p, C.POINTER(f3), _ffi.f32p]),
    "orc_kat_scatter": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64] + [C.POINTER(f3)] * 4 + [C.c_int, _ffi.f32p, _ffi.f32p]),
    "orc_kat_samplers": (None, [C.c_uint64, C.c_uint64, _ffi.f32p, _ffi.f32p, _ffi.f32p]),
}

_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_PATH):
            raise RuntimeError(f"{ORACLE_PATH} missing: run `make -C oracle` (or __graft_entry__.build())")
        api = dict(_ffi.scene_api("orc", with_desc=False))
        api.update(_EXTRA)
        _lib = _ffi.bind(C.CDLL(ORACLE_PATH), api)
    return _lib


def v3(x, y, z):
    return f3(float(x), float(y), float(z))


def fptr(a):
    return a.ctypes.data_as(_ffi.f32p)


class OracleScene(NativeScene):
    """(World, Camera) replayed into the oracle; adds the oracle's own render calls."""

    def __init__(self, world, camera=None):
        super().__init__(world, camera, backend=(oracle_lib(), "orc"))

    # the oracle does not restate the `image` crate's PNG decoder: images are decoded here (PIL) and registered by path
    def _prepare_png(self, path):
        from PIL import Image

        try:
            arr = np.ascontiguousarray(np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8))
        except OSError:
            return  # unreadable / missing: the oracle reports "cannot open" like File::open would
        self.lib.orc_register_png(self._h, path.encode(), arr.ctypes.data_as(_ffi.u8p), arr.shape[1], arr.shape[0])

    def _prepare_obj(self, path):
        d = os.path.dirname(path)
        for name in sorted(os.listdir(d or ".")):
            if name.lower().endswith(".png"):
                self._prepare_png(os.path.join(d, name) if d else name)

    def render_aov(self, w, h, seed=1, threads=0):
        out = dict(albedo=np.zeros((h, w, 3), np.float32), normal=np.zeros((h, w, 3), np.float32), object=np.zeros((h, w), np.uint32),
                   tri=np.zeros((h, w), np.uint32), t=np.zeros((h, w), np.float32))
        cnt = orc_counters()
        self.lib.orc_render_aov(self._h, w, h, seed, threads, fptr(out["albedo"]), fptr(out["normal"]), out["object"].ctypes.data_as(_ffi.u32p),
                                out["tri"].ctypes.data_as(_ffi.u32p), fptr(out["t"]), C.byref(cnt))
        out["counters"] = cnt.as_dict()
        return out

    def render(self, w, h, spp, max_depth=50, seed=1, threads=0, by_rows=False):
        """by_rows: distribute the work over the cores by image row instead of by frame (threads < 0 in orc_render) -- for renders of
        a few samples per pixel at a large resolution; a different (equally valid) assignment of random streams to pixels."""
        rgb, b = np.zeros((h, w, 3), np.float32), np.zeros((h, w), np.uint32)
        if by_rows:
            threads = -1
        cnt = orc_counters()
        self.lib.orc_render(self._h, w, h, spp, max_depth, seed, threads, fptr(rgb), b.ctypes.data_as(_ffi.u32p), C.byref(cnt))
        return rgb, b, cnt.as_dict()

"""The oracle's render loop (restating main.rs:150-295): determinism, golden fixtures, analytic cases."""
import os

import numpy as np
import pytest

from mass_raytrace_b200 import (Camera, Dielectric, DiffuseLight, Lambertian, Metal, SkyBackground, SolidBackground, SolidColor, Sphere, V3, Volume, World,
                                scenes)
from oracle_backend import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden_make():
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["cornell", "book1", "eve", "mesh_media"])
def test_oracle_reproduces_golden(name):
    got = _golden_make().make(name)
    want = np.load(os.path.join(HERE, "golden", f"{name}.npz"))
    for k in want.files:
        assert np.array_equal(got[k], want[k], equal_nan=True), k


def test_oracle_output_independent_of_thread_count():
    world, camera = scenes.cornell_box(1.0)
    s = OracleScene(world, camera)
    a = s.render(24, 24, 6, 50, seed=5, threads=1)
    b = s.render(24, 24, 6, 50, seed=5, threads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2]["rays"] == b[2]["rays"]
    c = s.render(24, 24, 6, 50, seed=6, threads=4)
    assert not np.array_equal(a[0], c[0])
    x = s.render_aov(24, 24, seed=5, threads=1)
    y = s.render_aov(24, 24, seed=5, threads=3)
    for k in ("albedo", "normal", "object", "tri", "t"):
        assert np.array_equal(x[k], y[k])


def test_empty_world_is_background_only():
    w = World(SolidBackground(V3(0.25, 0.5, 0.75)))
    s = OracleScene(w, Camera(40.0, V3(0, 0, 5), V3(0, 0, 0), V3(0, 1, 0), 1.0, 0.0, 5.0))
    rgb, b, cnt = s.render(8, 8, 4, 50, seed=1, threads=1)
    assert np.array_equal(rgb, np.broadcast_to(np.array([1.0, 2.0, 3.0], np.float32), (8, 8, 3))) and not b.any()
    assert cnt["rays"] == cnt["paths"] == 256
    aov = s.render_aov(8, 8)
    assert (aov["object"] == 0xFFFFFFFF).all() and np.isinf(aov["t"]).all() and not aov["normal"].any()


def test_white_furnace():
    # albedo-1 Lambertian sphere in a uniform radiance-1 environment: every path that escapes carries exactly 1 (world.rs:71-77)
    w = World(SolidBackground(V3(1, 1, 1)))
    w.add(Sphere(Lambertian(SolidColor((1, 1, 1, 1))), V3(0, 0, 0), 1.0))
    w.build_bvh()
    s = OracleScene(w, Camera(40.0, V3(0, 0, 4), V3(0, 0, 0), V3(0, 1, 0), 1.0, 0.0, 4.0))
    rgb, b, cnt = s.render(16, 16, 8, 50, seed=3, threads=2)
    assert np.array_equal(rgb, np.full((16, 16, 3), 8.0, np.float32))
    assert b[8, 8] >= 8 and b[0, 0] == 0  # centre pixels scatter at least once per sample, corner pixels miss


def test_depth_limit_semantics():
    # inside a closed mirror-like box nothing escapes: a path performs max_depth scatters, returns 0 and reports max_depth bounces
    w = World(SolidBackground(V3(1, 1, 1)))
    w.add(Sphere(Lambertian(SolidColor((1, 1, 1, 1))), V3(0, 0, 0), -10.0))  # inward-facing shell around the camera
    w.build_bvh()
    s = OracleScene(w, Camera(40.0, V3(0, 0, 0), V3(0, 0, -1), V3(0, 1, 0), 1.0, 0.0, 1.0))
    rgb, b, cnt = s.render(6, 6, 2, 1, seed=1, threads=1)
    assert not rgb.any() and (b == 2).all() and cnt["rays"] == 36 * 2  # depth 1: hit, scatter, trace(depth 0) -> (0, 0)
    for depth in (3, 50):  # a near-tangent bounce can slip through the shell below t_min = 0.001; everything else runs to the limit
        rgb, b, cnt = s.render(6, 6, 2, depth, seed=1, threads=1)
        assert (b <= 2 * depth).all() and b.mean() > 0.97 * 2 * depth and cnt["rays"] <= 36 * 2 * depth
        assert (rgb == 0).mean() > 0.9


def test_emitter_and_absorber():
    w = World(SolidBackground(V3(0, 0, 0)))
    w.add(Sphere(DiffuseLight(V3(2, 3, 4)), V3(0, 0, 0), 1.0))
    w.build_bvh()
    s = OracleScene(w, Camera(40.0, V3(0, 0, 4), V3(0, 0, 0), V3(0, 1, 0), 1.0, 0.0, 4.0))
    rgb, b, _ = s.render(9, 9, 3, 50, seed=1, threads=1)
    assert rgb[4, 4].tolist() == [6, 9, 12] and b[4, 4] == 0 and not rgb[0, 0].any()
    aov = s.render_aov(9, 9)
    assert aov["albedo"][4, 4].tolist() == [2, 3, 4] and aov["object"][4, 4] == 0  # no scatter -> (emitted, normal) world.rs:87


def test_volume_statistics():
    # dense constant medium: free-flight -ln(xi)/density (geom.rs:638); with density 1000 every ray entering the r=1 ball scatters
    # in it at once and the hit is reported on the Volume object, not on anything behind it
    w = World(SolidBackground(V3(1, 1, 1)))
    w.add(Volume(Sphere((), V3(0, 0, 0), 1.0), 1000.0, V3(0.5, 0.5, 0.5)))
    w.add(Sphere(Lambertian(SolidColor((1, 0, 0, 1))), V3(0, 0, -3), 1.0))
    w.build_bvh()
    s = OracleScene(w, Camera(30.0, V3(0, 0, 4), V3(0, 0, 0), V3(0, 1, 0), 1.0, 0.0, 4.0))
    aov = s.render_aov(9, 9)
    assert aov["object"][4, 4] == 0 and aov["normal"][4, 4].tolist() == [1, 0, 0] and abs(aov["t"][4, 4] - 0.75) < 0.01
    thin = World(SolidBackground(V3(1, 1, 1)))
    thin.add(Volume(Sphere((), V3(0, 0, 0), 1.0), 1e-6, V3(0.5, 0.5, 0.5)))
    thin.build_bvh()
    s2 = OracleScene(thin, Camera(30.0, V3(0, 0, 4), V3(0, 0, 0), V3(0, 1, 0), 1.0, 0.0, 4.0))
    assert (s2.render_aov(9, 9)["object"] == 0xFFFFFFFF).all()  # free flight ~1e6 >> 2: passes through


def test_volume_over_a_mesh_target():
    """Volume<I: Intersect> with a Model / Instance target (geom.rs:595-660): a very dense medium scatters right behind the surface the
    ray enters through, a very thin one never; the bounding box is the target's; a ray that starts inside is clipped at its origin."""
    from mass_raytrace_b200 import Model, PlyLoader, NativeScene
    cube = Model(PlyLoader.load(scenes.CUBE_PLY))  # [-1, 1]^3 (models/cube.ply)
    cam = Camera(30.0, V3(0, 0, 6), V3(0, 0, 0), V3(0, 1, 0), 1.0, 0.0, 6.0)
    for target, front in ((cube, 5.0 / 6.0), (cube.instance(V3(0, 0, 0.5), V3(0, 0, 0), V3(0.5, 0.5, 0.5)), 5.0 / 6.0)):  # t counts focus distances (world.rs:59)
        w = World(SolidBackground(V3(1, 1, 1)))
        w.add(Volume(target, 1.0e5, V3(0.5, 0.5, 0.5)))
        w.build_bvh()
        a = OracleScene(w, cam).render_aov(33, 33)
        assert a["object"][16, 16] == 0 and abs(a["t"][16, 16] - front) < 1e-3 and a["normal"][16, 16].tolist() == [1, 0, 0]
        assert a["object"][0, 0] == 0xFFFFFFFF
        host = NativeScene(w, cam)
        d = host.desc().contents
        assert d.n_volumes == 1 and d.n_instances == 1 and d.n_roots == 1  # the target is not an object of the world itself
        thin = World(SolidBackground(V3(1, 1, 1)))
        thin.add(Volume(target, 1.0e-7, V3(0.5, 0.5, 0.5)))
        thin.build_bvh()
        assert (OracleScene(thin, cam).render_aov(33, 33)["object"] == 0xFFFFFFFF).all()
    inside = Camera(60.0, V3(0, 0, 0), V3(0, 0, -1), V3(0, 1, 0), 1.0, 0.0, 1.0)
    w = World(SolidBackground(V3(1, 1, 1)))
    w.add(Volume(cube, 50.0, V3(0.5, 0.5, 0.5)))
    w.build_bvh()
    a = OracleScene(w, inside).render_aov(17, 17)
    assert (a["object"] == 0).all() and (a["t"] < 0.5).all() and (a["t"] >= 0).all(), (a["t"].min(), a["t"].max())


def test_host_flattens_the_round2_scenes():
    """libmrt_host.so on the scenes of the two round-2 features: an EveMaterial becomes one MRT_MAT_EVE entry whose palette is four
    consecutive solid surfaces (include/mrt.h), a Volume over a mesh an instance entry that is not in the world list."""
    import ctypes as C

    from extra_scenes import eve_scene, mesh_media_scene
    from mass_raytrace_b200 import NativeScene, _ffi

    class M(C.Structure):
        _fields_ = [("kind", C.c_int32), ("surface", C.c_int32), ("left", C.c_int32), ("right", C.c_int32), ("p", C.c_float * 4)]

    class S(C.Structure):
        _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("mode", C.c_int32), ("color", C.c_float * 4)]

    class V(C.Structure):
        _fields_ = [("target", C.c_uint32), ("neg_inv_density", C.c_float), ("material", C.c_int32), ("object_id", C.c_uint32)]

    host = NativeScene(*eve_scene())
    d = host.desc().contents
    mats = C.cast(d.materials, C.POINTER(M))
    surfs = C.cast(d.surfaces, C.POINTER(S))
    eve = [mats[i] for i in range(d.n_materials) if mats[i].kind == 8]
    assert len(eve) == 1
    m = eve[0]
    pal = C.cast(C.pointer(C.c_float(m.p[0])), C.POINTER(C.c_int32)).contents.value
    assert 0 <= pal and pal + 4 <= d.n_surfaces and all(surfs[pal + k].kind == 0 for k in range(4))
    assert [round(surfs[pal + k].color[3], 3) for k in range(3)] == [0.5, 0.85, 2.0]  # glow rides in the alpha of the first three
    assert all(surfs[x].kind == 1 for x in (m.surface, m.left, m.right)) and d.n_textures == 3
    host = NativeScene(*mesh_media_scene(0.9))
    d = host.desc().contents
    vols = C.cast(d.volumes, C.POINTER(V))
    assert d.n_volumes == 2 and d.n_instances == 2 and d.n_objects == 4
    for k in range(2):
        assert vols[k].target >> 29 == 3 and (vols[k].target & 0x1FFFFFFF) < d.n_instances  # MRT_PRIM_INSTANCE
        assert abs(vols[k].neg_inv_density + 1.0 / 0.9) < 1e-6

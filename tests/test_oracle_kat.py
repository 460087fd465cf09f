"""Known-answer tests that pin the CPU oracle. The reference has no tests or golden vectors, so every
expected value here is derived by hand from the cited reference formula (plus a few more)."""
import ctypes as C
import math

import numpy as np
import pytest

from mass_raytrace_b200 import (WRAP_CLAMP, WRAP_REPEAT, Camera, Dielectric, Lambertian, Metal, Model, PlyLoader, SkyBackground, SolidBackground,
                                SolidColor, Sphere, Texture, V3, V3_fill, World, scenes)
from mass_raytrace_b200 import _ffi
from oracle_backend import OracleScene, f3, fptr, v3

INF = float("inf")


def test_ply_cube_known_answer():
    # (1) cube.ply -> 12 triangles, first = (1,1,1), (-1,1,-1), (-1,1,1)   [cube.ply:11-13,19; ply_loader.rs:396-409]
    w = World(SolidBackground(V3(0, 0, 0)))
    tris = PlyLoader.load(scenes.CUBE_PLY)
    w.add(Model(tris))
    s = OracleScene(w)
    v = s.mesh_verts(tris)
    assert v.shape == (12, 9)
    assert v[0].tolist() == [1, 1, 1, -1, 1, -1, -1, 1, 1]
    assert v[11].tolist() == [1, 1, 1, -1, 1, 1, -1, -1, 1]  # face "3 0 2 7"
    assert np.abs(v).max() == 1.0


def test_sphere_known_answers(oracle):
    # (2) geom.rs:57-93: c=(0,0,-1) r=.5, o=0, d=(0,0,-1) -> t=.5, p=(0,0,-.5), n=(0,0,1), front
    out = np.zeros(8, np.float32)
    assert oracle.orc_kat_sphere(0, 0, -1, 0.5, C.byref(v3(0, 0, 0)), C.byref(v3(0, 0, -1)), 0.001, INF, fptr(out)) == 1
    assert out.tolist() == [0.5, 0, 0, -0.5, 0, 0, 1, 1]
    # from the centre: the '-' root is negative, the '+' root .5 is taken; normal flipped towards the ray, front_face false
    assert oracle.orc_kat_sphere(0, 0, -1, 0.5, C.byref(v3(0, 0, -1)), C.byref(v3(0, 0, -1)), 0.001, INF, fptr(out)) == 1
    assert out.tolist() == [0.5, 0, 0, -1.5, 0, 0, 1, 0]
    # direction is not normalised: d=(0,0,-2) halves t (geom.rs:59-66)
    assert oracle.orc_kat_sphere(0, 0, -1, 0.5, C.byref(v3(0, 0, 0)), C.byref(v3(0, 0, -2)), 0.001, INF, fptr(out)) == 1
    assert out[0] == 0.25
    # t_max shrinking: hit at .5 is outside [0.001, 0.4]; the far root 1.5 too
    assert oracle.orc_kat_sphere(0, 0, -1, 0.5, C.byref(v3(0, 0, 0)), C.byref(v3(0, 0, -1)), 0.001, 0.4, fptr(out)) == 0
    # negative radius flips the outward normal (hollow-glass trick): n = (p-c)/r
    assert oracle.orc_kat_sphere(0, 0, -1, -0.5, C.byref(v3(0, 0, 0)), C.byref(v3(0, 0, -1)), 0.001, INF, fptr(out)) == 1
    assert out[7] == 0 and out[6] == 1  # front_face false, normal still faces the ray
    # miss
    assert oracle.orc_kat_sphere(0, 0, -1, 0.5, C.byref(v3(0, 2, 0)), C.byref(v3(0, 0, -1)), 0.001, INF, fptr(out)) == 0


def test_aabb_known_answers(oracle):
    # (3) geom.rs:218-247
    lo, hi = v3(-1, -1, -1), v3(1, 1, 1)
    hit = lambda o, d, a=0.001, b=INF: oracle.orc_kat_aabb(C.byref(lo), C.byref(hi), C.byref(v3(*o)), C.byref(v3(*d)), a, b)
    assert hit((0, 0, 5), (0, 0, -1)) == 1
    # origin on the face plane with d.x = 0: v_max.x = 0/0 = NaN, v_min.x = -2/0 = -inf; f32::min/max drop the NaN ->
    # min.x = max.x = -inf -> t_max = -inf < t_min -> miss
    assert hit((1, 0, 5), (0, 0, -1)) == 0
    assert hit((-1, 0, 5), (0, 0, -1)) == 0  # v_min.x = 0/0 NaN, v_max.x = +inf -> min.x = max.x = +inf -> t_min = inf > t_max
    assert hit((0.5, 0, 5), (0, 0, -1)) == 1  # strictly inside the slab: (-inf, +inf)
    assert hit((2, 0, 5), (0, 0, -1)) == 0
    assert hit((0, 0, 5), (0, 0, 1)) == 0  # box behind the ray
    assert hit((0, 0, 5), (0, 0, -1), 0.001, 3.9) == 0  # t_max before the entry at t=4
    assert hit((0, 0, 5), (0, 0, -1), 0.001, 4.0) == 1  # t_max == t_min passes (`t_max < t_min` is the test)
    # zero-thickness box (a flat triangle pair) is hit: both slabs give the same t
    flat_lo, flat_hi = v3(-1, 0, -1), v3(1, 0, 1)
    assert oracle.orc_kat_aabb(C.byref(flat_lo), C.byref(flat_hi), C.byref(v3(0, 1, 0)), C.byref(v3(0, -1, 0)), 0.001, INF) == 1


def test_triangle_known_answers(oracle):
    # (4) geom.rs:504-577: triangle (0,0,0),(1,0,0),(0,1,0)
    tri = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32)
    out = np.zeros(11, np.float32)
    assert oracle.orc_kat_triangle(fptr(tri), C.byref(v3(0.25, 0.25, 1)), C.byref(v3(0, 0, -1)), 0.001, INF, fptr(out)) == 1
    assert out[0] == 1.0 and out[1:4].tolist() == [0.25, 0.25, 0.0]
    assert out[4:7].tolist() == [0, 0, 1] and out[7] == 1
    assert out[8:11].tolist() == [0.5, 0.25, 0.25]  # area barycentrics :539-545
    # two-sided: from below, the geometric normal (0,0,1) is flipped, front_face false (:511 uses |det|)
    assert oracle.orc_kat_triangle(fptr(tri), C.byref(v3(0.25, 0.25, -1)), C.byref(v3(0, 0, 1)), 0.001, INF, fptr(out)) == 1
    assert out[0] == 1.0 and out[4:7].tolist() == [0, 0, -1] and out[7] == 0
    # outside: u+v > 1; and parallel ray: |det| < 1e-6
    assert oracle.orc_kat_triangle(fptr(tri), C.byref(v3(0.75, 0.75, 1)), C.byref(v3(0, 0, -1)), 0.001, INF, fptr(out)) == 0
    assert oracle.orc_kat_triangle(fptr(tri), C.byref(v3(0.25, 0.25, 1)), C.byref(v3(1, 0, 0)), 0.001, INF, fptr(out)) == 0
    # edges are inclusive (u == 0 accepted; t == t_max accepted :531)
    assert oracle.orc_kat_triangle(fptr(tri), C.byref(v3(0.0, 0.5, 1)), C.byref(v3(0, 0, -1)), 0.001, 1.0, fptr(out)) == 1
    assert oracle.orc_kat_triangle(fptr(tri), C.byref(v3(0.25, 0.25, 1)), C.byref(v3(0, 0, -1)), 0.001, 0.999, fptr(out)) == 0


def test_rotation_turns_and_signs(oracle):
    # (5) math.rs:183-215: angle in turns; rotate_y(0.25): c0 = (cos, 0, sin, 0) = (~0, 0, 1, 0) -> (1,0,0) maps to (0,0,1)
    m = np.zeros(16, np.float32)
    oracle.orc_kat_rotate(1, 0.25, fptr(m))
    s, c = np.float32(math.sin(np.float32(0.25) * np.float32(math.pi) * np.float32(2))), None
    assert abs(m[0]) < 1e-7 and m[2] == 1.0 and m[8] == -1.0 and m[5] == 1.0 and m[15] == 1.0
    oracle.orc_kat_rotate(0, 0.25, fptr(m))  # c1 = (0, cos, sin, 0), c2 = (0, -sin, cos, 0)
    assert m[6] == 1.0 and m[9] == -1.0 and m[0] == 1.0
    oracle.orc_kat_rotate(2, 0.25, fptr(m))  # c0 = (cos, -sin, 0, 0), c1 = (sin, cos, 0, 0): opposite sign to the textbook Rz
    assert m[1] == -1.0 and m[4] == 1.0 and m[10] == 1.0
    oracle.orc_kat_rotate(1, 1.0, fptr(m))  # one full turn ~ identity
    assert abs(m[0] - 1) < 1e-6 and abs(m[2]) < 1e-6


def test_instance_bounds_and_matrices():
    # (6) geom.rs:343-381 + scenes/cornell.rs:46-49: cube at (10,5,0) scale 5 -> AABB (5,0,-5)-(15,10,5)
    world, camera = scenes.cornell_box(1.0)
    s = OracleScene(world, camera)
    tf, inv, aabb = s.instance_fields(1)
    assert aabb.tolist() == [5, 0, -5, 15, 10, 5]
    assert tf.reshape(4, 4).tolist() == [[5, 0, 0, 0], [0, 5, 0, 0], [0, 0, 5, 0], [10, 5, 0, 1]]  # columns c0..c3 = T*R*S
    np.testing.assert_allclose(inv.reshape(4, 4), [[0.2, 0, 0, 0], [0, 0.2, 0, 0], [0, 0, 0.2, 0], [-2, -1, 0, 1]], rtol=1e-6)
    # the rotated block (object 7): inverse really inverts (within f32), rotation -0.05 turn about y
    tf, inv, aabb = s.instance_fields(7)
    prod = tf.reshape(4, 4).T.astype(np.float64) @ inv.reshape(4, 4).T.astype(np.float64)
    np.testing.assert_allclose(prod, np.eye(4), atol=1e-6)
    ang = -0.05 * 2 * math.pi
    np.testing.assert_allclose(tf.reshape(4, 4)[0, :3], [1.75 * math.cos(ang), 0, 1.75 * math.sin(ang)], rtol=1e-6)
    # sphere bbox uses |r| (geom.rs:95-100)
    assert s.object_aabb(5).tolist() == [-0.25, 0, 0.25, 3.75, 4, 4.25]


def test_camera_known_answer():
    # (7) world.rs:16-51: Camera::new(37, (0,5,20), (0,5,0), (0,1,0), 16/9, 0, 20)
    w = World(SolidBackground(V3(0, 0, 0)))
    s = OracleScene(w, Camera(37.0, V3(0, 5, 20), V3(0, 5, 0), V3(0, 1, 0), 16.0 / 9.0, 0.0, 20.0))
    c = s.camera_fields()
    f = np.float32
    vh = f(math.tan(f(f(37.0) * f(math.pi) / f(180.0)) / f(2.0))) * f(2.0)
    vw = f(16.0 / 9.0) * vh
    assert c[0:3].tolist() == [0, 5, 20]
    assert c[12:15].tolist() == [1, 0, 0] and c[15:18].tolist() == [0, 1, 0]  # u, v (w = (0,0,1))
    np.testing.assert_allclose(c[6:9], [vw * 20, 0, 0], rtol=1e-6)  # horizontal
    np.testing.assert_allclose(c[9:12], [0, vh * 20, 0], rtol=1e-6)  # vertical
    np.testing.assert_allclose(c[3:6], [-vw * 10, 5 - vh * 10, 0], rtol=1e-6, atol=1e-6)  # lower_left_corner
    assert c[18] == 0.0


def test_dielectric_known_answers(oracle):
    # (8) material.rs:296-299: normal incidence, ior 1.5: r0 = ((1-1.5)/(1+1.5))^2 = 0.04
    assert abs(oracle.orc_kat_reflectance(1.0, 1.5) - 0.04) < 1e-7
    assert abs(oracle.orc_kat_reflectance(0.0, 1.5) - 1.0) < 1e-6  # grazing
    x = 1 - 0.5
    assert abs(oracle.orc_kat_reflectance(0.5, 1.0 / 1.5) - (0.04 + 0.96 * x ** 5)) < 1e-6
    # refract straight through: v = (0,0,-1), n = (0,0,1) -> (0,0,-1) for any eta (math.rs:119-124)
    out = np.zeros(3, np.float32)
    oracle.orc_kat_refract(C.byref(v3(0, 0, -1)), C.byref(v3(0, 0, 1)), 1 / 1.5, fptr(out))
    assert out.tolist() == [0, 0, -1]
    # Snell: 45 degrees into glass
    s = math.sqrt(0.5)
    oracle.orc_kat_refract(C.byref(v3(s, 0, -s)), C.byref(v3(0, 0, 1)), 1 / 1.5, fptr(out))
    assert abs(out[0] - s / 1.5) < 1e-6 and abs(np.linalg.norm(out) - 1) < 1e-6
    # scatter: reflect iff xi < 0.04 at normal incidence -> over many seeds ~4% reflect (dir.z > 0)
    w = World(SolidBackground(V3(0, 0, 0)))
    glass = Dielectric(1.5)
    w.add(Sphere(glass, V3(0, 0, 0), 1.0))
    sc = OracleScene(w)
    m = sc._material(glass)
    att, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    refl = 0
    n = 4000
    for seed in range(n):
        assert oracle.orc_kat_scatter(sc._h, m, seed * 7919 + 1, C.byref(v3(0, 0, 1)), C.byref(v3(0, 0, -1)), C.byref(v3(0, 0, 0)), C.byref(v3(0, 0, 1)), 1,
                                      fptr(att), fptr(d)) == 1
        assert att.tolist() == [1, 1, 1]
        refl += d[2] > 0
    assert abs(refl / n - 0.04) < 0.012


def test_wrap_known_answers(oracle):
    # (9) texture.rs:283-291
    out = np.zeros(2, np.float32)
    oracle.orc_kat_wrap(WRAP_REPEAT, -0.25, 1.25, fptr(out))
    assert out.tolist() == [0.75, 0.25]
    oracle.orc_kat_wrap(WRAP_REPEAT, 1.0, 0.0, fptr(out))
    assert out.tolist() == [1.0, 0.0]  # exactly 1.0 stays 1.0 (only x > 1 wraps)
    oracle.orc_kat_wrap(WRAP_CLAMP, -3.0, 7.0, fptr(out))
    assert out.tolist() == [0.0, 1.0]


def _nodes(n):  # BvhNode::new split rule geom.rs:120-144
    return 1 if n <= 2 else 1 + _nodes(n // 2) + _nodes(n - n // 2)


def test_bvh_node_counts():
    # (10) 488 spheres -> 511 BvhNodes? the count follows the split rule; 12 triangles -> 15
    assert _nodes(12) == 15
    world, camera = scenes.book1_spheres()
    s = OracleScene(world, camera)
    assert s.tlas_node_count() == _nodes(len(world.objects))
    wc, cc = scenes.cornell_box()
    sc = OracleScene(wc, cc)
    assert sc.tlas_node_count() == _nodes(8) == 7
    assert sc.mesh_node_count(wc.objects[0].model.triangles) == 15


def test_texture_bilinear_and_load(oracle):
    # texture.rs:70-103 (byte/255, no sRGB decode) and :126-148 (bilinear over (w-1, h-1))
    px = np.array([[[0, 0, 0, 255], [255, 0, 0, 255]], [[0, 255, 0, 255], [255, 255, 255, 0]]], np.uint8)
    tex = Texture(px, WRAP_CLAMP)
    w = World(SolidBackground(V3(0, 0, 0)))
    w.add(Sphere(Lambertian(tex), V3(0, 0, 0), 1.0))
    s = OracleScene(w)
    h = s._surface(tex)
    out = np.zeros(4, np.float32)
    oracle.orc_kat_texture_get(s._h, h, 0.0, 0.0, fptr(out))
    assert out.tolist() == [0, 0, 0, 1]
    oracle.orc_kat_texture_get(s._h, h, 1.0, 0.0, fptr(out))
    assert out.tolist() == [1, 0, 0, 1]
    oracle.orc_kat_texture_get(s._h, h, 0.5, 0.0, fptr(out))
    assert out.tolist() == [0.5, 0, 0, 1]
    oracle.orc_kat_texture_get(s._h, h, 0.5, 0.5, fptr(out))
    np.testing.assert_allclose(out, [0.5, 0.5, 0.25, 0.75])
    oracle.orc_kat_texture_get(s._h, h, 2.0, -1.0, fptr(out))  # clamp
    assert out.tolist() == [1, 0, 0, 1]


def test_sky_background(oracle):
    # material.rs:57-62: lerp(white, (.5,.7,1), .5*(unit(d).y+1))
    w = World(SkyBackground())
    s = OracleScene(w)
    out = np.zeros(3, np.float32)
    oracle.orc_kat_background(s._h, C.byref(v3(0, 2, 0)), fptr(out))
    assert out.tolist() == [0.5, np.float32(0.7), 1.0]
    oracle.orc_kat_background(s._h, C.byref(v3(0, -3, 0)), fptr(out))
    assert out.tolist() == [1, 1, 1]
    oracle.orc_kat_background(s._h, C.byref(v3(1, 0, 0)), fptr(out))
    np.testing.assert_allclose(out, [0.75, 0.85, 1.0], rtol=1e-6)


def test_rejection_sampler_distributions(oracle):
    # math.rs:80-109: in-sphere uniform in the ball (E|v|^2 = 3/5), unit vector on the sphere (E z^2 = 1/3), disk (E|v|^2 = 1/2)
    n = 200000
    ball, sph, disk = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
    oracle.orc_kat_samplers(123, n, fptr(ball), fptr(sph), fptr(disk))
    assert (np.sum(ball ** 2, 1) < 1).all() and (np.sum(disk ** 2, 1) < 1).all()
    np.testing.assert_allclose(np.linalg.norm(sph, axis=1), 1.0, atol=1e-6)
    assert abs(np.mean(np.sum(ball ** 2, 1)) - 0.6) < 0.005
    assert abs(np.mean(np.sum(disk ** 2, 1)) - 0.5) < 0.005
    np.testing.assert_allclose(np.mean(sph ** 2, 0), [1 / 3] * 3, atol=0.005)
    np.testing.assert_allclose(np.mean(sph, 0), [0, 0, 0], atol=0.005)


def test_metal_and_lambertian_scatter(oracle):
    # material.rs:261-279: fuzz 0 -> perfect mirror of unit(d); scatter only if dir.n > 0. :255-258: fuzz clamped to 1.
    w = World(SolidBackground(V3(0, 0, 0)))
    mirror, lam = Metal(0.0, SolidColor((0.7, 0.6, 0.5, 1.0))), Lambertian(SolidColor((0.1, 0.2, 0.3, 1.0)))
    w.add(Sphere(mirror, V3(0, 0, 0), 1.0))
    w.add(Sphere(lam, V3(3, 0, 0), 1.0))
    s = OracleScene(w)
    att, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    r = math.sqrt(0.5)
    assert oracle.orc_kat_scatter(s._h, s._material(mirror), 5, C.byref(v3(0, 0, 0)), C.byref(v3(2, 0, -2)), C.byref(v3(0, 0, 0)), C.byref(v3(0, 0, 1)), 1,
                                  fptr(att), fptr(d)) == 1
    np.testing.assert_allclose(d, [r, 0, r], rtol=1e-6)
    np.testing.assert_allclose(att, [0.7, 0.6, 0.5])
    # Lambertian: dir = n + unit vector -> |dir - n| = 1, attenuation = surface colour
    for seed in range(50):
        assert oracle.orc_kat_scatter(s._h, s._material(lam), seed + 1, C.byref(v3(0, 0, 0)), C.byref(v3(0, 0, -1)), C.byref(v3(0, 0, 0)), C.byref(v3(0, 0, 1)),
                                      1, fptr(att), fptr(d)) == 1
        assert abs(np.linalg.norm(d - np.array([0, 0, 1], np.float32)) - 1) < 1e-5
        np.testing.assert_allclose(att, [0.1, 0.2, 0.3])


def test_resolve_rgb8(oracle):
    # main.rs:640-722: (sum/count)^(1/2.2) -> clamp -> *255 truncated; NaN -> 255 (f32::min drops NaN); dump flips rows :763-768
    sums = np.array([[[0, 2.0, 8.0], [4.0, 1.0, float("nan")]], [[0.5, 0.25, 100.0], [4.0, 4.0, 4.0]]], np.float32)
    out = np.zeros((2, 2, 3), np.uint8)
    oracle.orc_resolve_rgb8(fptr(sums), 2, 2, 4, 0, out.ctypes.data_as(_ffi.u8p))
    exp = lambda v: int(min(max((v / 4.0) ** (1 / 2.2), 0.0), 1.0) * 255.0)
    assert out[0, 0].tolist() == [0, exp(2.0), 255]
    assert out[0, 1].tolist() == [255, exp(1.0), 255]
    assert out[1, 0].tolist() == [exp(0.5), exp(0.25), 255]
    flipped = np.zeros_like(out)
    oracle.orc_resolve_rgb8(fptr(sums), 2, 2, 4, 1, flipped.ctypes.data_as(_ffi.u8p))
    assert np.array_equal(flipped, out[::-1])
    oracle.orc_resolve_rgb8(fptr(sums), 2, 2, 0, 0, out.ctypes.data_as(_ffi.u8p))  # count == 0 -> black (:649-654)
    assert not out.any()

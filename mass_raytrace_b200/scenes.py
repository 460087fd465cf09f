"""Scenes for the BASELINE.json configs, written against the scene-API mirror in api.py.

Each function returns (World, Camera) like Scene::generate (reference src/scenes.rs:25-33). Only reference features are
used: Sphere, Model/Instance of triangle meshes, Volume, Lambertian/Metal/Dielectric/DiffuseLight,
SolidColor/Texture surfaces, Solid/Sky backgrounds. Scene randomness comes from FastRand(1) (fastrand::seed(1), main.rs:86).
"""
import math
import os

import numpy as np

from .api import (ABSORB, WRAP_CLAMP, Camera, Dielectric, DiffuseLight, FastRand, Lambertian, Metal, Model, PlyLoader, SkyBackground,
                  SolidBackground, SolidColor, Sphere, Texture, Triangles, V3, V3_fill, Volume, World)

_HERE = os.path.dirname(os.path.abspath(__file__))
CUBE_PLY = os.path.join(_HERE, "assets", "cube.ply")

f32 = np.float32


def _length(v):
    v = np.asarray(v, dtype=f32)
    return float(np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2], dtype=f32))


def cornell_box(aspect_ratio=1.0):
    """CornellBox exactly as reference src/scenes/cornell.rs:29-99 (cfg 2: aspect 1.0, 1024x1024)."""
    world = World(SolidBackground(V3(0, 0, 0)))
    red = Lambertian(SolidColor((1.0, 0.0, 0.0, 1.0)))
    green = Lambertian(SolidColor((0.0, 1.0, 0.0, 1.0)))
    white = Lambertian(SolidColor((1.0, 1.0, 1.0, 1.0)))
    light = DiffuseLight(V3_fill(8.0))
    sphere_material = Dielectric(1.3)
    cube = Model(PlyLoader.load(CUBE_PLY, material=ABSORB))
    world.add(cube.instance(V3(-10.0, 5.0, 0.0), V3(0, 0, 0), V3_fill(5.0)).with_material(red))
    world.add(cube.instance(V3(10.0, 5.0, 0.0), V3(0, 0, 0), V3_fill(5.0)).with_material(green))
    world.add(cube.instance(V3(0.0, 15.0, 0.0), V3(0, 0, 0), V3_fill(5.0)).with_material(white))
    world.add(cube.instance(V3(0.0, 5.0, -10.0), V3(0, 0, 0), V3_fill(5.0)).with_material(white))
    world.add(cube.instance(V3(0.0, -5.0, -0.0), V3(0, 0, 0), V3_fill(5.0)).with_material(white))
    world.add(Sphere(sphere_material, V3(1.75, 2.0, 2.25), 2.0))
    world.add(cube.instance(V3(0.0, float(f32(10.0) - f32(0.00011)), 0.0), V3(0, 0, 0), V3(1.0, 0.0001, 1.0)).with_material(light))
    world.add(cube.instance(V3(-2.0, 3.0, -1.0), V3(0.0, -0.05, 0.0), V3(1.75, 3.1, 1.75)).with_material(white))
    world.build_bvh()
    look_from, look_at = V3(0.0, 5.0, 20.0), V3(0.0, 5.0, 0.0)
    focus = _length(np.subtract(look_from, look_at))
    return world, Camera(37.0, look_from, look_at, V3(0, 1, 0), aspect_ratio, 0.0, focus)


def book1_spheres(aspect_ratio=1.5, aperture=0.1):
    """RTIOW book-1 final scene from reference parts (cfg 1): ~488 spheres, SkyBackground."""
    rng = FastRand(1)
    world = World(SkyBackground())
    world.add(Sphere(Lambertian(SolidColor((0.5, 0.5, 0.5, 1.0))), V3(0, -1000, 0), 1000.0))
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose = rng.f32()
            cx, cz = f32(a) + f32(0.9) * f32(rng.f32()), f32(b) + f32(0.9) * f32(rng.f32())
            if _length((cx - f32(4.0), f32(0.0), cz)) > 0.9:
                center = V3(cx, 0.2, cz)
                if choose < 0.8:
                    alb = [rng.f32() * rng.f32() for _ in range(3)]
                    world.add(Sphere(Lambertian(SolidColor((alb[0], alb[1], alb[2], 1.0))), center, 0.2))
                elif choose < 0.95:
                    alb = [0.5 + 0.5 * rng.f32() for _ in range(3)]
                    fuzz = 0.5 * rng.f32()
                    world.add(Sphere(Metal(fuzz, SolidColor((alb[0], alb[1], alb[2], 1.0))), center, 0.2))
                else:
                    world.add(Sphere(Dielectric(1.5), center, 0.2))
    world.add(Sphere(Dielectric(1.5), V3(0, 1, 0), 1.0))
    world.add(Sphere(Lambertian(SolidColor((0.4, 0.2, 0.1, 1.0))), V3(-4, 1, 0), 1.0))
    world.add(Sphere(Metal(0.0, SolidColor((0.7, 0.6, 0.5, 1.0))), V3(4, 1, 0), 1.0))
    world.build_bvh()
    return world, Camera(20.0, V3(13, 2, 3), V3(0, 0, 0), V3(0, 1, 0), aspect_ratio, aperture, 10.0)


def sphere_grid(aspect_ratio=16.0 / 9.0, dim=50):
    """SphereGrid as reference src/scenes/sphere_grid.rs:29-94."""
    rng = FastRand(1)
    world = World(SolidBackground(V3(0, 0, 0)))
    white = Lambertian(SolidColor((1.0, 1.0, 1.0, 1.0)))
    cube = Model(PlyLoader.load(CUBE_PLY, material=ABSORB))
    world.add(cube.instance(V3(0.0, -1000.0, 0.0), V3(0, 0, 0), V3_fill(1000.0)).with_material(white))
    r = f32(1.0)
    d = r * f32(2.0)
    a = np.sqrt(d * d - r * r, dtype=f32)
    for i in range(-dim, dim):
        for j in range(-dim, dim):
            off = r if j % 2 == 0 else f32(0.0)
            x, z, y = f32(i) * d + off, f32(j) * a, r
            rr = float(r - f32(0.05))
            if (i, j) == (0, 0):
                m = DiffuseLight(V3_fill(3.0))
            elif (i, j) in ((-1, 0), (1, 0), (1, -1), (0, -1), (1, 1), (0, 1)):
                m = Dielectric(1.8)
            else:
                m = Metal(0.0, SolidColor((rng.f32(), rng.f32(), rng.f32(), 1.0)))
            world.add(Sphere(m, V3(x, y, z), rr))
    world.build_bvh()
    look_from, look_at = V3(6.0, 8.0, 5.0), V3(0, 0, 0)
    return world, Camera(40.0, look_from, look_at, V3(0, 1, 0), aspect_ratio, 0.0, _length(look_from))


def write_synthetic_ply(path, nu=1024, nv=512, seed=1, fmt="binary_little_endian"):
    """A displaced lat-long grid of nu x nv quads -> 2*nu*nv triangles, written as a PLY that PlyLoader reads
    (stand-in for the un-shipped models/lucy.ply, SURVEY.md §8d cfg 3). Returns (n_triangles, max_abs_coordinate)."""
    rs = np.random.RandomState(seed)
    u = np.linspace(0.0, 2.0 * np.pi, nu + 1)
    v = np.linspace(0.02, np.pi - 0.02, nv + 1)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    k = rs.randint(2, 9, size=(6, 2))
    amp = rs.uniform(0.02, 0.08, size=6)
    ph = rs.uniform(0, 2 * np.pi, size=6)
    rad = np.ones_like(uu)
    for i in range(6):
        rad += amp[i] * np.sin(k[i, 0] * uu + ph[i]) * np.sin(k[i, 1] * vv)
    rad += 0.004 * np.sin(97 * uu) * np.sin(61 * vv)
    # a tall figure (Lucy is ~1.7x taller than wide); the loader swizzles (y, z, x) like scenes/lucy.rs:38
    x = rad * np.sin(vv) * np.cos(uu) * 0.55
    y = rad * np.sin(vv) * np.sin(uu) * 0.55
    z = rad * np.cos(vv) * 1.0
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype("<f4") * f32(800.0)
    idx = np.arange((nu + 1) * (nv + 1), dtype=np.int64).reshape(nv + 1, nu + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    header = (f"ply\nformat {fmt} 1.0\ncomment synthetic mesh seed {seed}\nelement vertex {len(verts)}\nproperty float x\nproperty float y\n"
              f"property float z\nelement face {len(faces)}\nproperty list uchar int vertex_indices\nend_header\n")
    with open(path, "wb") as f:
        f.write(header.encode())
        if fmt == "ascii":
            for p in verts:
                f.write(("%r %r %r\n" % (float(p[0]), float(p[1]), float(p[2]))).encode())
            for t in faces:
                f.write(("3 %d %d %d\n" % (t[0], t[1], t[2])).encode())
        else:
            big = fmt == "binary_big_endian"
            f.write(verts.astype(">f4" if big else "<f4").tobytes())
            rec = np.zeros(len(faces), dtype=[("n", "u1"), ("i", ">i4" if big else "<i4", 3)])
            rec["n"] = 3
            rec["i"] = faces
            f.write(rec.tobytes())
    return len(faces), float(np.abs(verts).max())


def lucy_layout(ply_path, max_dim, aspect_ratio=16.0 / 9.0, grid=0):
    """Lucy scene layout (reference src/scenes/lucy.rs:29-95) around a mesh loaded from `ply_path` with the (y, z, x) swizzle.
    grid = 0: one instance at the origin (cfg 3); grid = 5: the reference's 11 x 11 field of rotated instances."""
    rng = FastRand(1)
    world = World(SolidBackground(V3(0, 0, 0)))
    lucy = Model(PlyLoader.load(ply_path, vertex_perm=(1, 2, 0), material=ABSORB))
    white = Lambertian(SolidColor((1.0, 1.0, 1.0, 1.0)))
    cube = Model(PlyLoader.load(CUBE_PLY, material=ABSORB))
    world.add(cube.instance(V3(0.0, -1000.0, 0.0), V3(0, 0, 0), V3_fill(1000.0)).with_material(white))
    scale = float((f32(1.0) / f32(max_dim)) * f32(2.0))
    for x in range(-grid, grid + 1):
        for z in range(-grid, grid + 1):
            col = [1.0 - rng.f32() * 0.5 for _ in range(3)]
            material = Lambertian(SolidColor((col[0], col[1], col[2], 1.0)))
            world.add(lucy.instance(V3(x * 3.0, 1.0, z * 3.0), V3(0.0, rng.f32(), 0.0), V3_fill(scale)).with_material(material))
    world.add(Sphere(DiffuseLight(V3(40.0, 40.0, 50.0)), V3(10000.0, 4000.0, 4800.0), 1500.0))
    world.build_bvh()
    look_from, look_at = V3(6.0, 8.0, 5.0), V3(0, 0, 0)
    return world, Camera(40.0, look_from, look_at, V3(0, 1, 0), aspect_ratio, 0.0, _length(look_from))


def multi_mesh(ply_paths, max_dims, aspect_ratio=16.0 / 9.0):
    """cfg 5: distinct meshes (one BLAS each, no sharing) as Models' instances on a 5 x 2 grid + ground cube + sun."""
    rng = FastRand(1)
    world = World(SolidBackground(V3(0, 0, 0)))
    white = Lambertian(SolidColor((1.0, 1.0, 1.0, 1.0)))
    cube = Model(PlyLoader.load(CUBE_PLY, material=ABSORB))
    world.add(cube.instance(V3(0.0, -1000.0, 0.0), V3(0, 0, 0), V3_fill(1000.0)).with_material(white))
    for i, (path, md) in enumerate(zip(ply_paths, max_dims)):
        m = Model(PlyLoader.load(path, vertex_perm=(1, 2, 0), material=ABSORB))
        col = [1.0 - rng.f32() * 0.5 for _ in range(3)]
        scale = float((f32(1.0) / f32(md)) * f32(2.0))
        gx, gz = i % 5, i // 5
        world.add(m.instance(V3((gx - 2) * 2.6, 1.0, (gz - 0.5) * 3.0), V3(0.0, rng.f32(), 0.0), V3_fill(scale))
                  .with_material(Lambertian(SolidColor((col[0], col[1], col[2], 1.0)))))
    world.add(Sphere(DiffuseLight(V3(40.0, 40.0, 50.0)), V3(10000.0, 4000.0, 4800.0), 1500.0))
    world.build_bvh()
    look_from, look_at = V3(6.0, 8.0, 9.0), V3(0, 0.5, 0)
    return world, Camera(40.0, look_from, look_at, V3(0, 1, 0), aspect_ratio, 0.0, _length(np.subtract(look_from, look_at)))


def uv_sphere_triangles(center, radius, nu=64, nv=32, material=ABSORB):
    """A lat-long sphere as Triangle::with_norms_and_uvs triangles (geom.rs:468): the only way an image texture can vary over
    an object, because Sphere hits carry uv = None (geom.rs:84)."""
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    theta, phi = vv * np.pi, uu * 2.0 * np.pi
    n = np.stack([-np.sin(theta) * np.cos(phi), -np.cos(theta), np.sin(theta) * np.sin(phi)], -1)
    p = np.asarray(center) + radius * n
    uv = np.stack([uu, vv], -1)
    idx = np.arange((nu + 1) * (nv + 1)).reshape(nv + 1, nu + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
    tri = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
    P, N, T = p.reshape(-1, 3)[tri], n.reshape(-1, 3)[tri], uv.reshape(-1, 2)[tri]
    # drop the degenerate triangles at the poles (zero area -> 0/0 barycentrics in geom.rs:543-545)
    area = np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
    keep = area > 1e-9 * radius * radius
    return Triangles(material=material, verts=P[keep].reshape(-1, 9).astype(f32), normals=N[keep].reshape(-1, 9).astype(f32),
                     uvs=T[keep].reshape(-1, 6).astype(f32))


def synthetic_earth_texture(w=256, h=128, seed=7):
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    land = (np.sin(xx * 0.11 + 1.3) * np.cos(yy * 0.17) + 0.5 * np.sin(xx * 0.031 + yy * 0.05)) > 0.2
    img = np.zeros((h, w, 4), np.uint8)
    img[..., 0] = np.where(land, 60, 20) + rs.randint(0, 20, (h, w))
    img[..., 1] = np.where(land, 140, 60) + rs.randint(0, 20, (h, w))
    img[..., 2] = np.where(land, 50, 170) + rs.randint(0, 20, (h, w))
    img[..., 3] = 255
    return img


def book2_final(aspect_ratio=16.0 / 9.0, boxes_per_side=32, n_cluster=1000):
    """RTIOW book-2 final scene restated with reference parts only (cfg 4). Checker/Perlin textures and motion
    blur do not exist in the reference (no ray time, world.rs:168-172) and are replaced by solid Lambertians / a static sphere."""
    rng = FastRand(1)
    world = World(SolidBackground(V3(0, 0, 0)))
    cube = Model(PlyLoader.load(CUBE_PLY, material=ABSORB))

    def box(lo, hi, material):  # cube.ply spans [-1, 1]^3
        c = [(a + b) * 0.5 for a, b in zip(lo, hi)]
        s = [(b - a) * 0.5 for a, b in zip(lo, hi)]
        return cube.instance(V3(*c), V3(0, 0, 0), V3(*s)).with_material(material)

    ground = Lambertian(SolidColor((0.48, 0.83, 0.53, 1.0)))
    w = 2000.0 / boxes_per_side
    for i in range(boxes_per_side):
        for j in range(boxes_per_side):
            x0, z0 = -1000.0 + i * w, -1000.0 + j * w
            y1 = 1.0 + 100.0 * rng.f32()
            world.add(box((x0, 0.0, z0), (x0 + w, y1, z0 + w), ground))
    world.add(box((123.0, 554.0, 147.0), (423.0, 554.02, 412.0), DiffuseLight(V3_fill(7.0))))
    world.add(Sphere(Lambertian(SolidColor((0.7, 0.3, 0.1, 1.0))), V3(400, 400, 200), 50.0))  # the moving sphere, at rest
    world.add(Sphere(Dielectric(1.5), V3(260, 150, 45), 50.0))
    world.add(Sphere(Metal(1.0, SolidColor((0.8, 0.8, 0.9, 1.0))), V3(0, 150, 145), 50.0))
    world.add(Sphere(Dielectric(1.5), V3(360, 150, 145), 70.0))
    world.add(Volume(Sphere(ABSORB, V3(360, 150, 145), 70.0), 0.2, V3(0.2, 0.4, 0.9)))
    world.add(Volume(Sphere(ABSORB, V3(0, 0, 0), 5000.0), 0.0001, V3(1, 1, 1)))
    earth = Lambertian(Texture(synthetic_earth_texture(), WRAP_CLAMP))
    world.add(Model(uv_sphere_triangles((400.0, 200.0, 400.0), 100.0, material=earth)))
    world.add(Sphere(Lambertian(SolidColor((0.5, 0.5, 0.5, 1.0))), V3(220, 280, 300), 80.0))  # the Perlin sphere, solid grey
    white = Lambertian(SolidColor((0.73, 0.73, 0.73, 1.0)))
    ang = math.radians(15.0)
    ca, sa = math.cos(ang), math.sin(ang)
    for _ in range(n_cluster):
        px, py, pz = 165.0 * rng.f32(), 165.0 * rng.f32(), 165.0 * rng.f32()
        rx, rz = ca * px + sa * pz, -sa * px + ca * pz
        world.add(Sphere(white, V3(rx - 100.0, py + 270.0, rz + 395.0), 10.0))
    world.build_bvh()
    look_from, look_at = V3(478, 278, -600), V3(278, 278, 0)
    return world, Camera(40.0, look_from, look_at, V3(0, 1, 0), aspect_ratio, 0.0, 10.0)


MENGER_CUBE_SIDES = [(0, 1, 1), (1, 0, 1), (1, 1, 0), (0, -1, -1), (-1, 0, -1), (-1, -1, 0), (0, -1, 1), (-1, 0, 1), (-1, 1, 0), (0, 1, -1),
                     (1, 0, -1), (1, -1, 0), (-1, -1, 1), (-1, 1, -1), (1, -1, -1), (-1, 1, 1), (1, -1, 1), (1, 1, -1), (1, 1, 1), (-1, -1, -1)]


def menger(aspect_ratio=16.0 / 9.0, levels=5):
    """Menger sponge of 20^levels unit-cube instances (reference src/scenes/menger.rs:29-124; levels = 5 there -> 3.2 M instances,
    the TLAS stress case). The reference's nebula CubeMap needs un-shipped PNGs (eve.rs `environment`), so the sky is SkyBackground."""
    world = World(SkyBackground())
    cube = Model(PlyLoader.load(CUBE_PLY, material=ABSORB))
    foggy = Metal(0.7, SolidColor((0.5, 0.5, 0.5, 1.0)))
    material = Lambertian(SolidColor((1.0, 1.0, 1.0, 1.0)))
    sides = np.array(MENGER_CUBE_SIDES, dtype=f32)
    pos = np.zeros((1, 3), f32)
    dims = f32(2.0)
    for lvl in range(levels - 1, -1, -1):  # xyz = (i,j,k) * dims * 3^lvl + parent, coarsest level first (menger.rs:85-103)
        step = sides * dims * f32(3.0) ** f32(lvl)
        pos = (step[None, :, :] + pos[:, None, :]).reshape(-1, 3).astype(f32)
    for p in pos:
        world.add(cube.instance(V3(p[0], p[1], p[2]), V3(0, 0, 0), V3(1, 1, 1)).with_material(material))
    extent = float(3.0 ** levels)
    world.add(cube.instance(V3(0.0, -extent - 1.0, 0.0), V3(0, 0, 0), V3(500000.0, 1.0, 500000.0)).with_material(foggy))
    world.build_bvh()
    s = extent / 243.0  # the reference camera (2680, 140, 2000) frames the levels = 5 sponge (3^5 = 243)
    look_from, look_at = V3(2680.0 * s, 140.0 * s, 2000.0 * s), V3(0, 0, 0)
    return world, Camera(15.0, look_from, look_at, V3(0, 1, 0), aspect_ratio, 0.0, _length(look_from))

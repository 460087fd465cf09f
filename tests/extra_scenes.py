"""Scenes of the round-2 features, shared by the GPU parity tests and the golden-fixture generator."""
import numpy as np

from mass_raytrace_b200 import (WRAP_REPEAT, Camera, EveMaterial, Lambertian, Metal, Model, PlyLoader, SkyBackground, SolidBackground, SolidColor, Sphere, Texture, V3,
                                Volume, World, scenes)


def eve_scene(seed=7, flat_normals=False):
    """UV meshes carrying an EveMaterial (eve.rs:23-133): normal + occlusion, albedo + roughness and paint / material / dirt / glow
    textures, tangent-space normals through Material::normal (geom.rs:551-560)."""
    rs = np.random.RandomState(seed)
    no = rs.randint(64, 192, (16, 32, 4)).astype(np.uint8)  # normal x in G, y in A (normal_occlusion :66-73); moderate tilts
    if flat_normals:
        no[..., 1] = 128
        no[..., 3] = 128
    ar = rs.randint(40, 256, (8, 16, 4)).astype(np.uint8)
    pmdg = rs.randint(0, 256, (8, 8, 4)).astype(np.uint8)
    pmdg[..., 2] //= 3       # little dirt
    pmdg[..., 3] //= 8       # faint glow
    eve = EveMaterial(Texture(no, WRAP_REPEAT), Texture(ar, WRAP_REPEAT), Texture(pmdg, WRAP_REPEAT))
    w = World(SkyBackground())
    w.add(Sphere(Lambertian(SolidColor((0.5, 0.5, 0.5, 1))), V3(0, -1000, 0), 1000.0))
    w.add(Model(scenes.uv_sphere_triangles((-1.1, 1.0, 0.0), 1.0, 32, 16, material=eve)))
    ship = Model(scenes.uv_sphere_triangles((0.0, 0.0, 0.0), 1.0, 24, 12, material=eve))
    w.add(ship.instance(V3(1.2, 0.8, 0.3), V3(0.2, 0.7, 0.1), V3(0.9, 0.6, 0.7)))
    w.build_bvh()
    cam = Camera(35.0, V3(0, 2.0, 7), V3(0, 0.9, 0), V3(0, 1, 0), 1.5, 0.0, 7.0)
    return w, cam



def mesh_media_scene(density):
    """Volume<I: Intersect> (geom.rs:595-660) with meshes as targets: the medium fills a Model (a UV sphere mesh, no transform) and a
    rotated, scaled Instance of the cube; a mirror sphere behind them."""
    w = World(SolidBackground(V3(0.9, 0.95, 1.0)))
    w.add(Sphere(Lambertian(SolidColor((0.5, 0.5, 0.5, 1))), V3(0, -1000, 0), 1000.0))
    ball = Model(scenes.uv_sphere_triangles((-1.3, 1.0, 0.0), 0.9, 24, 12, material=()))
    w.add(Volume(ball, density, V3(0.8, 0.3, 0.2)))
    cube = Model(PlyLoader.load(scenes.CUBE_PLY))
    w.add(Volume(cube.instance(V3(1.3, 0.9, 0.0), V3(0.3, 0.6, 0.1), V3(0.8, 0.8, 0.8)), density, V3(0.2, 0.4, 0.8)))
    w.add(Sphere(Metal(0.0, SolidColor((0.9, 0.9, 0.9, 1))), V3(0, 0.5, -2.5), 0.5))
    w.build_bvh()
    return w, Camera(30.0, V3(0, 1.2, 7), V3(0, 0.9, 0), V3(0, 1, 0), 1.5, 0.0, 7.0)

"""A small tour of every kernel for compute-sanitizer (memcheck): Cornell render, a GPU-built mesh with alpha-tested texture, a
volume scene, AOV pass, resolve, scene re-upload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mass_raytrace_b200 import (WRAP_CLAMP, Camera, Lambertian, Model, NativeScene, Renderer, SkyBackground, SolidColor, Sphere, Texture, V3, World, scenes)
r = Renderer(0)
w, c = scenes.cornell_box(1.0)
r.set_scene(NativeScene(w, c)); r.render_aov(64, 64); rgb, b, n = r.render(64, 64, 8, 50, seed=3); print("cornell", float(rgb.mean()), int(b.sum()))
r.resolve_rgb8(8)
px = np.random.RandomState(2).randint(0, 256, (16, 16, 4)).astype(np.uint8); px[:, ::2, 3] = 0
w = World(SkyBackground())
w.add(Model(scenes.uv_sphere_triangles((0.0, 1.0, 0.0), 1.0, 192, 96, material=Lambertian(Texture(px, WRAP_CLAMP)))))  # 36,864 triangles: built on the GPU
w.add(Sphere(Lambertian(SolidColor((0.8, 0.2, 0.2, 1))), V3(0, -1000, 0), 1000.0)); w.build_bvh()
cam = Camera(35.0, V3(0, 1.5, 5), V3(0, 1, 0), V3(0, 1, 0), 1.5, 0.0, 5.0)
for dev in (1, 0):
    r.set_option(Renderer.OPT_DEVICE_BUILD, dev)
    r.set_scene(NativeScene(w, cam)); a = r.render_aov(96, 64); rgb, b, n = r.render(96, 64, 4, 50, seed=5); print("mesh dev", dev, float(rgb.mean()), int(b.sum()), int((a["object"] == 0).sum()))
w, c = scenes.book2_final(boxes_per_side=6, n_cluster=50)
r.set_scene(NativeScene(w, c)); rgb, b, n = r.render(96, 54, 4, 50, seed=7); print("book2", float(rgb.mean()), int(b.sum()))
# round 2: media that fill meshes, EveMaterial (full kernel variants), axis-parallel and degenerate rays, a two-device handle when there are two GPUs
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mass_raytrace_b200 import EveMaterial, PlyLoader, SolidBackground, Volume, WRAP_REPEAT
rs = np.random.RandomState(7)
tex = lambda h, w_: Texture(rs.randint(0, 256, (h, w_, 4)).astype(np.uint8), WRAP_REPEAT)
w = World(SolidBackground(V3(0.9, 0.95, 1.0)))
w.add(Sphere(Lambertian(SolidColor((0.5, 0.5, 0.5, 1))), V3(0, -1000, 0), 1000.0))
cube = Model(PlyLoader.load(scenes.CUBE_PLY))
w.add(Volume(cube.instance(V3(1.3, 0.9, 0.0), V3(0.3, 0.6, 0.1), V3(0.8, 0.8, 0.8)), 0.9, V3(0.2, 0.4, 0.8)))
w.add(Volume(Model(scenes.uv_sphere_triangles((-1.3, 1.0, 0.0), 0.9, 24, 12, material=())), 2.0, V3(0.8, 0.3, 0.2)))
w.add(Model(scenes.uv_sphere_triangles((0.0, 1.0, -2.0), 1.0, 32, 16, material=EveMaterial(tex(16, 32), tex(8, 16), tex(8, 8)))))
w.build_bvh()
cam = Camera(30.0, V3(0, 1.0, 7), V3(0, 1.0, 0), V3(0, 1, 0), 1.0, 0.0, 7.0)
r.set_scene(NativeScene(w, cam)); a = r.render_aov(65, 65); rgb, b, n = r.render(65, 65, 8, 50, seed=9); print("volumes over meshes + eve", float(rgb.mean()), int(b.sum()))
import ctypes as C
ndev = C.c_int(0); C.CDLL("libcudart.so.12").cudaGetDeviceCount(C.byref(ndev))
if ndev.value >= 2:
    m = Renderer([0, 1]); w2, c2 = scenes.cornell_box(1.0); m.set_scene(NativeScene(w2, c2)); rgb, b, n = m.render(64, 64, 8, 50, seed=3); print("two devices", float(rgb.mean()), n); m.close()
r.close(); print("done")

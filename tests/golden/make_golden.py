"""Generates the golden fixtures in this directory by running the CPU oracle (oracle/liboracle.so).

The reference has no golden vectors and cannot be run in this image (no Rust toolchain), so these pin the ORACLE'S output,
not the reference's: they guard the oracle against regressions on the CPU side and give the GPU tests fixed targets that do not
depend on re-running the oracle. Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mass_raytrace_b200 import scenes  # noqa: E402
from oracle_backend import OracleScene  # noqa: E402
from extra_scenes import eve_scene, mesh_media_scene  # noqa: E402

SPECS = {
    # name: (scene factory, aov size, render size, spp, seed)
    "cornell": (lambda: scenes.cornell_box(1.0), (64, 64), (48, 48), 64, 2024),
    "book1": (lambda: scenes.book1_spheres(1.5, aperture=0.0), (96, 64), (60, 40), 64, 2024),
    # round 2: EveMaterial (tangent-space normals through Material::normal, geom.rs:551-560) and media that fill meshes (Volume<I>, geom.rs:595)
    "eve": (eve_scene, (96, 64), (60, 40), 64, 2024),
    "mesh_media": (lambda: mesh_media_scene(0.9), (96, 64), (60, 40), 64, 2024),
}


def make(name):
    factory, (aw, ah), (rw, rh), spp, seed = SPECS[name]
    world, camera = factory()
    s = OracleScene(world, camera)
    aov = s.render_aov(aw, ah, seed=seed, threads=1)
    rgb, bounces, cnt = s.render(rw, rh, spp, 50, seed=seed, threads=2)
    return dict(aov_object=aov["object"], aov_tri=aov["tri"], aov_t=aov["t"], aov_normal=aov["normal"], aov_albedo=aov["albedo"],
                sum_rgb=rgb, sum_bounces=bounces, spp=np.uint32(spp), seed=np.uint64(seed), rays=np.uint64(cnt["rays"]))


if __name__ == "__main__":
    for name in SPECS:
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **make(name))
        print("wrote", name)

"""Render time at 16 M and 32 M rays in flight (MRT_OPT_POOL_SLOTS) for a few job sizes. Usage: gpu_capacity_quick.py [mesh10m]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_raytrace_b200 import NativeScene, Renderer, scenes
tmp = tempfile.mkdtemp()
n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
jobs = [("mesh1m", scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080, (32, 256)), ("cornell", scenes.cornell_box(1.0), 1024, 1024, (16, 125)),
        ("book2", scenes.book2_final(), 1920, 1080, (16, 125)), ("book1", scenes.book1_spheres(1.5, aperture=0.1), 1200, 800, (10,))]
if "mesh10m" in sys.argv[1:]:
    paths, mds = [], []
    for i in range(10):
        q = os.path.join(tmp, f"m{i}.ply"); n, md = scenes.write_synthetic_ply(q, 1024, 512, seed=100 + i); paths.append(q); mds.append(md)
    jobs = [("mesh10m", scenes.multi_mesh(paths, mds), 3840, 2160, (4, 16, 32)), ("mesh10m-hd", scenes.multi_mesh(paths, mds), 1920, 1080, (16, 64, 128)), ("mesh1m-4k", jobs[0][1], 3840, 2160, (8, 32, 64))]
for name, (w, c), W, H, spps in jobs:
    for slots in (1 << 24, 1 << 25):
        r = Renderer(0); r.set_option(Renderer.OPT_POOL_SLOTS, slots); r.set_scene(NativeScene(w, c, defer_mesh_bvh=True))
        row = []
        for spp in spps:
            best = 1e9
            for rep in range(4):
                r.reset(W, H); r.accumulate(0, spp); best = min(best, r.stats()["render_ms"])
            row.append(f"{spp} spp {best:8.2f} ms")
        print(f"{name:8s} {slots >> 20:3d} M in flight: " + "  ".join(row), flush=True)
        r.close()

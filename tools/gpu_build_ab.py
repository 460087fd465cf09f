"""Tree builders against each other: the host's SAH builder, the GPU's PLOC at several radii, the GPU's plain LBVH (radius 0).
Per workload: upload time (second upload), render / extend time, node visits and triangle tests per ray."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_raytrace_b200 import NativeScene, Renderer, scenes
names = (sys.argv[1] if len(sys.argv) > 1 else "mesh1m,menger").split(",")
variants = [("sah", 0, 16)] + [(f"ploc{r}" if r else "lbvh", 2, r) for r in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,8,16,32").split(",")]]
tmp = tempfile.mkdtemp()
for name in names:
    if name == "menger": (w, c), W, H, spp = scenes.menger(levels=4), 1920, 1080, 8
    elif name == "book2": (w, c), W, H, spp = scenes.book2_final(), 1920, 1080, 8
    elif name == "mesh10m":
        paths, mds = [], []
        for i in range(10):
            q = os.path.join(tmp, f"m{i}.ply"); n, md = scenes.write_synthetic_ply(q, 1024, 512, seed=100 + i); paths.append(q); mds.append(md)
        (w, c), W, H, spp = scenes.multi_mesh(paths, mds), 3840, 2160, 2
    else:
        n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
        (w, c), W, H, spp = scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080, 8
    host = NativeScene(w, c, defer_mesh_bvh=True); host.desc()
    for label, db, radius in variants:
        r = Renderer(0)
        r.set_option(Renderer.OPT_DEVICE_BUILD, db)
        if hasattr(Renderer, "OPT_BUILD_RADIUS"): r.set_option(Renderer.OPT_BUILD_RADIUS, radius)  # only with profiles/r02_ploc_experiment.patch applied
        r.set_scene(host)
        t0 = time.perf_counter(); r.set_scene(host); up = time.perf_counter() - t0
        r.reset(W, H); r.accumulate(0, 2)
        best = 1e9
        for rep in range(3):
            r.reset(W, H); r.accumulate(0, spp); st = r.stats(); best = min(best, st["render_ms"])
        r.set_option(Renderer.OPT_TIME_KERNELS, 1); r.reset(W, H); r.accumulate(0, spp); st = r.stats()
        r.set_option(Renderer.OPT_TIME_KERNELS, 0); r.set_option(Renderer.OPT_COUNT_VISITS, 1); r.reset(W, H); r.accumulate(0, 1); cs = r.stats()
        print(f"{name:8s} {label:7s} upload {up * 1e3:8.1f} ms | render {best:8.2f} ms extend {st['extend_ms']:7.2f} | nodes/ray {cs['node_visits'] / cs['rays']:5.2f} "
              f"tris/ray {cs['tri_tests'] / cs['rays']:4.2f} inst/ray {cs['instance_tests'] / cs['rays']:4.2f}", flush=True)
        r.close()

"""Find the sample whose render takes far longer than the others (a 'monster' path) and describe it."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mass_raytrace_b200 import NativeScene, Renderer, scenes
tmp = tempfile.mkdtemp()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
name = sys.argv[3] if len(sys.argv) > 3 else "mesh1m"
if name == "cornell": (w, c), W, H = scenes.cornell_box(1.0), 1024, 1024
elif name == "book1": (w, c), W, H = scenes.book1_spheres(1.5, aperture=0.1), 1200, 800
elif name == "book2": (w, c), W, H = scenes.book2_final(), 1920, 1080
elif name == "menger": (w, c), W, H = scenes.menger(levels=4), 1920, 1080
elif name == "mesh10m":
    paths, mds = [], []
    for i in range(10):
        q = os.path.join(tmp, f"m{i}.ply"); n, md = scenes.write_synthetic_ply(q, 1024, 512, seed=100 + i); paths.append(q); mds.append(md)
    (w, c), W, H = scenes.multi_mesh(paths, mds), 3840, 2160
else:
    n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
    (w, c), W, H = scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080
r = Renderer(0); r.set_scene(NativeScene(w, c, defer_mesh_bvh=True))
times = []
for s in range(lo, hi):
    r.reset(W, H); r.accumulate(s, 1, 50, seed=2024); times.append(r.stats()["render_ms"])
times = np.array(times); med = np.median(times)
print(name, "median ms per sample", med, "max", times.max(), "at sample", lo + int(times.argmax()), "| samples over 2x the median:", int((times > 2 * med).sum()))
for k in np.argsort(-times)[:1]:
    s = lo + int(k)
    r.set_option(Renderer.OPT_COUNT_VISITS, 1); r.reset(W, H); r.accumulate(s, 1, 50, seed=2024); st = r.stats(); r.set_option(Renderer.OPT_COUNT_VISITS, 0)
    rgb, b, cnt = r.download()
    print(f"sample {s}: {times[k]:.2f} ms, rays {st['rays']}, node visits/ray {st['node_visits'] / st['rays']:.1f} (worst ray {st['max_ray_node_visits']}), tri tests/ray {st['tri_tests'] / st['rays']:.1f}, "
          f"non-finite pixels {int((~np.isfinite(rgb)).any(-1).sum())}, max bounces {int(b.max())} at pixel {np.unravel_index(int(b.argmax()), b.shape)}")
    r.set_option(Renderer.OPT_FINISH_PATHS, 0); r.reset(W, H); r.accumulate(s, 1, 50, seed=2024); print("   without k_finish:", r.stats()["render_ms"], "ms", r.stats()["iterations"], "iterations"); r.set_option(Renderer.OPT_FINISH_PATHS, 98304)
if os.environ.get("MONSTER_DETAIL"):
    s = lo + int(times.argmax())
    r.set_option(Renderer.OPT_TIME_KERNELS, 1); r.reset(W, H); r.accumulate(s, 1, 50, seed=2024); st = r.stats(); r.set_option(Renderer.OPT_TIME_KERNELS, 0)
    print("timed:", {k: st[k] for k in ("render_ms", "extend_ms", "shade_ms", "generate_ms", "iterations", "kernel_launches")})
    for opt, val, name in ((Renderer.OPT_REFILL_LANES, 32, "refill 32"), (Renderer.OPT_NODE_BURST, 0xFFFFFFFF, "burst unbounded"), (Renderer.OPT_DEVICE_BUILD, 0, "host SAH tree")):
        r.set_option(opt, val)
        if opt == Renderer.OPT_DEVICE_BUILD: r.set_scene(NativeScene(w, c, defer_mesh_bvh=True))
        r.reset(W, H); r.accumulate(s, 1, 50, seed=2024); print(name, r.stats()["render_ms"], "ms")

"""What a render costs beyond its samples: render time against samples per pixel (the intercept is the ramp-up + drain of the wavefront),
for several drain thresholds (MRT_OPT_FINISH_PATHS)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_raytrace_b200 import NativeScene, Renderer, scenes
name = sys.argv[1] if len(sys.argv) > 1 else "mesh1m"
tmp = tempfile.mkdtemp()
if name == "cornell": (w, c), W, H = scenes.cornell_box(1.0), 1024, 1024
elif name == "book1": (w, c), W, H = scenes.book1_spheres(1.5, aperture=0.1), 1200, 800
elif name == "book2": (w, c), W, H = scenes.book2_final(), 1920, 1080
else:
    n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
    (w, c), W, H = scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080
r = Renderer(0); r.set_scene(NativeScene(w, c))
spps = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "8,16,32,64,128".split(","))]
for fp in [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "0,65536,1048576,4194304".split(","))]:
    r.set_option(Renderer.OPT_FINISH_PATHS, fp)
    row = []
    for spp in spps:
        best = 1e9
        for rep in range(3):
            r.reset(W, H); r.accumulate(0, spp); st = r.stats(); best = min(best, st["render_ms"])
        row.append((spp, best, st["iterations"]))
    # least-squares line through (spp, ms): slope = ms per sample per pixel, intercept = fixed cost
    n = len(row); sx = sum(a for a, _, _ in row); sy = sum(b for _, b, _ in row); sxx = sum(a * a for a, _, _ in row); sxy = sum(a * b for a, b, _ in row)
    slope = (n * sxy - sx * sy) / (n * sxx - sx * sx); icpt = (sy - slope * sx) / n
    print(f"{name} finish_paths {fp:8d}: " + "  ".join(f"{a}spp {b:7.2f}ms/{it}it" for a, b, it in row) + f" | {slope:.4f} ms/spp + {icpt:.2f} ms", flush=True)

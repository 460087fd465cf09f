"""One render of one workload and nothing else (for ncu: the k_extend launches 0..n of this process are the wavefront iterations)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_raytrace_b200 import NativeScene, Renderer, scenes
name, spp = sys.argv[1], int(sys.argv[2])
tmp = tempfile.mkdtemp()
if name == "cornell": (w, c), W, H = scenes.cornell_box(1.0), 1024, 1024
elif name == "book2": (w, c), W, H = scenes.book2_final(), 1920, 1080
elif name == "book1": (w, c), W, H = scenes.book1_spheres(1.5, aperture=0.1), 1200, 800
elif name == "menger": (w, c), W, H = scenes.menger(levels=4), 1920, 1080
elif name == "mesh1m":
    n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
    (w, c), W, H = scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080
elif name == "mesh10m":
    paths, mds = [], []
    for i in range(10):
        p = os.path.join(tmp, f"m{i}.ply"); n, md = scenes.write_synthetic_ply(p, 1024, 512, seed=100 + i); paths.append(p); mds.append(md)
    (w, c), W, H = scenes.multi_mesh(paths, mds), 3840, 2160
r = Renderer(0)
if os.environ.get("MRT_POOL_SLOTS"): r.set_option(Renderer.OPT_POOL_SLOTS, int(os.environ["MRT_POOL_SLOTS"]))  # the ncu captures of profiles/ were taken at 2^24
r.set_scene(NativeScene(w, c))
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
best = 1e9
for rep in range(reps):
    r.reset(W, H); r.accumulate(0, spp); st = r.stats(); best = min(best, st["render_ms"])
st["render_ms"] = best
print(f"{name} spp {spp}: render {st['render_ms']:.2f} ms, {st['rays']/1e6:.1f} M rays, {st['iterations']} iterations")

"""Would two wavefronts on two streams fill each other's kernel tails? One context rendering S spp against two contexts (own streams, half the rays
in flight each) rendering S/2 spp each at the same time from two host threads."""
import os, sys, tempfile, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_raytrace_b200 import NativeScene, Renderer, scenes
name, spp = sys.argv[1], int(sys.argv[2])
tmp = tempfile.mkdtemp()
if name == "cornell": (w, c), W, H = scenes.cornell_box(1.0), 1024, 1024
elif name == "book2": (w, c), W, H = scenes.book2_final(), 1920, 1080
elif name == "book1": (w, c), W, H = scenes.book1_spheres(1.5, aperture=0.1), 1200, 800
elif name == "mesh10m":
    paths, mds = [], []
    for i in range(10):
        p = os.path.join(tmp, f"m{i}.ply"); n, md = scenes.write_synthetic_ply(p, 1024, 512, seed=100 + i); paths.append(p); mds.append(md)
    (w, c), W, H = scenes.multi_mesh(paths, mds), 3840, 2160
else:
    n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
    (w, c), W, H = scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080
host = NativeScene(w, c, defer_mesh_bvh=True); host.desc()
def wall(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
one = Renderer(0); one.set_scene(host); one.reset(W, H); one.accumulate(0, 2)
def single():
    one.reset(W, H); one.accumulate(0, spp, 50, seed=1)
t1 = wall(single)
for slots in (1 << 23, 1 << 24):
    pair = [Renderer(0), Renderer(0)]
    for r in pair:
        r.set_option(Renderer.OPT_POOL_SLOTS, slots); r.set_scene(host); r.reset(W, H); r.accumulate(0, 2)
    def dual():
        th = [threading.Thread(target=lambda r=r, k=k: (r.reset(W, H), r.accumulate(k * (spp // 2), spp // 2, 50, seed=1))) for k, r in enumerate(pair)]
        for t in th: t.start()
        for t in th: t.join()
    t2 = wall(dual)
    print(f"{name} {spp} spp: one wavefront {t1:.2f} ms; two concurrent wavefronts of {spp // 2} spp, {slots >> 20} M rays in flight each: {t2:.2f} ms ({100 * (t1 / t2 - 1):+.1f} %)", flush=True)
    for r in pair: r.close()

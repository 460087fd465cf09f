"""Sweep run-time options (BVH leaf size / triangle cost, refill threshold) on several workloads: per-kernel ms and Mrays/s."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_raytrace_b200 import NativeScene, Renderer, scenes
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cornell", "book1", "book2", "mesh1m"]
combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[2:]] or [(4, 100, 32), (4, 100, 28), (4, 100, 24), (4, 100, 20), (4, 100, 16), (4, 100, 12), (4, 100, 8)]
tmp = tempfile.mkdtemp()
def build(name):
    if name == "cornell": return scenes.cornell_box(1.0), 1024, 1024, 16
    if name == "book1": return scenes.book1_spheres(1.5, 0.1), 1200, 800, 10
    if name == "book2": return scenes.book2_final(), 1920, 1080, 8
    if name == "menger": return scenes.menger(levels=4), 1920, 1080, 8
    if name == "mesh10m":
        paths, mds = [], []
        for i in range(10):
            q = os.path.join(tmp, f"m{i}.ply"); n, md = scenes.write_synthetic_ply(q, 1024, 512, seed=100 + i); paths.append(q); mds.append(md)
        return scenes.multi_mesh(paths, mds), 3840, 2160, 2
    if name == "mesh1m":
        n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
        return scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080, 8
    raise SystemExit(name)
for name in names:
    (w, c), W, H, spp = build(name)
    host = NativeScene(w, c)
    r = Renderer(0)
    for combo in combos:
        leaf, cost, refill = combo[:3]; burst = combo[3] if len(combo) > 3 else 0
        r.set_option(Renderer.OPT_BVH_LEAF_TRIS, leaf); r.set_option(Renderer.OPT_BVH_TRI_COST, cost); r.set_option(Renderer.OPT_REFILL_LANES, refill); r.set_option(Renderer.OPT_NODE_BURST, burst)
        r.set_scene(host)
        r.set_option(Renderer.OPT_TIME_KERNELS, 0)
        r.reset(W, H); r.accumulate(0, 2)
        plain = None
        for rep in range(3):
            r.reset(W, H); r.accumulate(0, spp); st = r.stats()
            plain = st["render_ms"] if plain is None else min(plain, st["render_ms"])
        r.set_option(Renderer.OPT_TIME_KERNELS, 1)
        r.reset(W, H); r.accumulate(0, spp); st = r.stats()
        r.set_option(Renderer.OPT_TIME_KERNELS, 0); r.set_option(Renderer.OPT_COUNT_VISITS, 1)
        r.reset(W, H); r.accumulate(0, 1); cs = r.stats()
        r.set_option(Renderer.OPT_COUNT_VISITS, 0)
        print(f"{name:8s} leaf {leaf} cost {cost:3d} refill {refill:2d} burst {burst:2d}: render {plain:8.2f} ms = {st['rays']/plain/1e3:7.1f} Mrays/s | extend {st['extend_ms']:7.2f} shade {st['shade_ms']:6.2f} gen {st['generate_ms']:5.2f} | "
              f"nodes/ray {cs['node_visits']/cs['rays']:5.2f} tris/ray {cs['tri_tests']/cs['rays']:4.2f} inst/ray {cs['instance_tests']/cs['rays']:4.2f}", flush=True)
    r.close()

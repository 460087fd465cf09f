"""Sweep k_extend's warp-vote thresholds on two workloads (extend-only and whole-render Mrays/s)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_raytrace_b200 import NativeScene, Renderer, scenes
work = {"cornell": (scenes.cornell_box(1.0), 1024, 1024, 16), "book1": (scenes.book1_spheres(1.5, 0.1), 1200, 800, 10)}
combos = [(16, 12, 1), (16, 12, 2), (16, 12, 4), (16, 16, 2), (20, 12, 2), (24, 12, 2), (16, 0, 1), (8, 0, 1), (24, 0, 1), (32, 0, 1)]
if len(sys.argv) > 1:
    combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for name, ((w, c), W, H, spp) in work.items():
    r = Renderer(0); r.set_scene(NativeScene(w, c))
    r.reset(W, H); r.accumulate(0, 2)
    r.set_option(Renderer.OPT_TIME_KERNELS, 1)
    for refill, node, burst in combos:
        r.set_option(Renderer.OPT_REFILL_LANES, refill); r.set_option(Renderer.OPT_NODE_LANES, node); r.set_option(Renderer.OPT_NODE_BURST, burst)
        best = None
        for rep in range(3):
            r.reset(W, H); r.accumulate(0, spp)
            st = r.stats()
            v = (st["rays"] / st["extend_ms"] / 1e3, st["rays"] / st["render_ms"] / 1e3, st["shade_ms"], st["generate_ms"], st["render_ms"])
            best = v if best is None or v[0] > best[0] else best
        print(f"{name:8s} refill>={refill:2d} node>={node:2d} burst {burst}: extend {best[0]:7.1f} Mrays/s  render {best[1]:7.1f} Mrays/s  (shade {best[2]:.2f} ms, generate {best[3]:.2f} ms, total {best[4]:.2f} ms)", flush=True)
    r.close()

"""Rays in flight (MRT_OPT_POOL_SLOTS) against render time."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_raytrace_b200 import NativeScene, Renderer, scenes
tmp = tempfile.mkdtemp(); n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
work = {"cornell": (scenes.cornell_box(1.0), 1024, 1024, 64), "mesh1m": (scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080, 32), "book2": (scenes.book2_final(), 1920, 1080, 16)}
for name, ((w, c), W, H, spp) in work.items():
    r = Renderer(0); r.set_scene(NativeScene(w, c))
    for cap in (1 << 22, 1 << 24, 1 << 25, 1 << 26):
        r.set_option(Renderer.OPT_POOL_SLOTS, cap); r.reset(W, H); r.accumulate(0, 4)
        best = None
        for rep in range(3):
            r.reset(W, H); r.accumulate(0, spp); st = r.stats(); best = st["render_ms"] if best is None else min(best, st["render_ms"])
        print(f"{name:8s} capacity {cap >> 20:2d} M: render {best:8.2f} ms = {st['paths']/best/1e3:7.1f} Mpaths/s, {st['iterations']} iterations", flush=True)
    r.close()

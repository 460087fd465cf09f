"""How much faster are coherent rays? Extend throughput (Grays/s) of primary rays only (max_depth 1) against the full path mix."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mass_raytrace_b200 import NativeScene, Renderer, scenes
tmp = tempfile.mkdtemp(); n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
work = {"cornell": (scenes.cornell_box(1.0), 1024, 1024, 16), "book2": (scenes.book2_final(), 1920, 1080, 8), "mesh1m": (scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080, 8)}
for name, ((w, c), W, H, spp) in work.items():
    r = Renderer(0); r.set_scene(NativeScene(w, c)); r.reset(W, H); r.accumulate(0, 2)
    r.set_option(Renderer.OPT_TIME_KERNELS, 1)
    for depth in (1, 2, 3, 50):
        r.reset(W, H); r.accumulate(0, spp, depth); st = r.stats()
        print(f"{name:8s} max_depth {depth:2d}: rays {st['rays']/1e6:7.1f} M  extend {st['extend_ms']:7.2f} ms = {st['rays']/st['extend_ms']/1e6:6.2f} Grays/s   shade {st['shade_ms']:6.2f} gen {st['generate_ms']:5.2f}", flush=True)
    r.close()

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mass_raytrace_b200 import NativeScene, Renderer, scenes
work = {"cornell": (scenes.cornell_box(1.0), 1024, 1024, 16), "book1": (scenes.book1_spheres(1.5, 0.1), 1200, 800, 10)}
for name, ((w, c), W, H, spp) in work.items():
    r = Renderer(0); r.set_scene(NativeScene(w, c)); r.reset(W, H); r.accumulate(0, 2)
    r.set_option(Renderer.OPT_TIME_KERNELS, 1)
    ref = None
    for mode in (0, 1):
        r.set_option(Renderer.OPT_SHADE_INORDER, mode)
        best = None
        for rep in range(3):
            r.reset(W, H); r.accumulate(0, spp); st = r.stats()
            best = st if best is None or st["render_ms"] < best["render_ms"] else best
        img = r.download()
        same = ref is None or (np.array_equal(img[0], ref[0]) and np.array_equal(img[1], ref[1]))
        ref = img if ref is None else ref
        print(f"{name:8s} inorder={mode} render {best['render_ms']:7.2f} ms  extend {best['extend_ms']:7.2f}  shade {best['shade_ms']:6.2f}  generate {best['generate_ms']:5.2f}  => {best['rays']/best['render_ms']/1e3:7.1f} Mrays/s  identical={same}", flush=True)
    r.close()

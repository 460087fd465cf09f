import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mass_raytrace_b200 import NativeScene, Renderer, scenes
work = {"book1 1200x800x10": (scenes.book1_spheres(1.5, 0.1), 1200, 800, 10), "cornell 1024^2x16": (scenes.cornell_box(1.0), 1024, 1024, 16), "cornell 256^2x4": (scenes.cornell_box(1.0), 256, 256, 4)}
for name, ((w, c), W, H, spp) in work.items():
    r = Renderer(0); r.set_scene(NativeScene(w, c)); r.reset(W, H); r.accumulate(0, 2)
    ref = None
    for thr in (0, 4096, 32768, 65536, 262144, 1 << 20):
        r.set_option(Renderer.OPT_FINISH_PATHS, thr)
        best = None
        for rep in range(3):
            r.reset(W, H); r.accumulate(0, spp); st = r.stats()
            best = st if best is None or st["render_ms"] < best["render_ms"] else best
        img = r.download()
        same = ref is None or (np.array_equal(img[0], ref[0]) and np.array_equal(img[1], ref[1]))
        ref = img if ref is None else ref
        print(f"{name:20s} finish<={thr:8d}: render {best['render_ms']:7.2f} ms iterations {best['iterations']:3d} launches {best['kernel_launches']:4d} rays {best['rays']} => {best['paths']/best['render_ms']/1e3:7.1f} Mpaths/s identical={same}", flush=True)
    r.close()

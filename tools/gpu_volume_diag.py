"""Where do the GPU's and the oracle's per-pixel Volume hit frequencies differ most? (diagnostic for test_volume_hit_probability_per_pixel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mass_raytrace_b200 import NativeScene, Renderer, scenes
from mass_raytrace_b200.api import Volume as VolumeT
from oracle_backend import OracleScene
world, camera = scenes.book2_final(boxes_per_side=12, n_cluster=200)
r = Renderer(0); r.set_scene(NativeScene(world, camera)); orc = OracleScene(world, camera)
w, h, n = 240, 135, 64
vid = [i for i, ob in enumerate(world.objects) if isinstance(ob, VolumeT)][1]
fg, fo, fo2 = np.zeros((h, w)), np.zeros((h, w)), np.zeros((h, w))
for seed in range(1, n + 1):
    g = r.render_aov(w, h, seed=seed); o = orc.render_aov(w, h, seed=1000 + seed); o2 = orc.render_aov(w, h, seed=5000 + seed)
    fg += g["object"] == vid; fo += o["object"] == vid; fo2 += o2["object"] == vid
base = orc.render_aov(w, h, seed=1)
for name, a, b in (("gpu-orc", fg / n, fo / n), ("orc-orc", fo2 / n, fo / n)):
    p = 0.5 * (a + b); mid = (p > 0.05) & (p < 0.95)
    z = np.zeros((h, w)); z[mid] = (a[mid] - b[mid]) / np.sqrt(2 * p[mid] * (1 - p[mid]) / n)
    print(name, "mean z2", float((z[mid] ** 2).mean()), "max", float(np.abs(z).max()))
    for k in np.argsort(-np.abs(z).ravel())[:8]:
        y, x = divmod(int(k), w)
        print(f"  pixel ({x},{y}) z {z[y, x]:6.2f} a {a[y, x]:.3f} b {b[y, x]:.3f} other-oracle {fo2[y, x] / n:.3f} behind: obj {base['object'][y, x]} t {base['t'][y, x]:.2f}")

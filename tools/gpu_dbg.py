import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mass_raytrace_b200 import NativeScene, Renderer, scenes
from oracle_backend import OracleScene
np.set_printoptions(linewidth=200, precision=9)
world, camera = scenes.cornell_box(1.0)
host = NativeScene(world, camera); orc = OracleScene(world, camera)
r = Renderer(0); r.set_scene(host)
for w in (96, 512):
    h = w
    g = r.render_aov(w, h); o = orc.render_aov(w, h)
    bad = (g["object"] != o["object"]) | (g["tri"] != o["tri"]) | ((g["t"] != o["t"]) & ~(np.isinf(g["t"]) & np.isinf(o["t"])))
    ys, xs = np.nonzero(bad)
    print(w, "mismatches", len(ys), "normal mismatches", int((g["normal"] != o["normal"]).any(-1).sum()), "albedo", int((g["albedo"] != o["albedo"]).any(-1).sum()))
    for y, x in list(zip(ys, xs))[:40]:
        print(f"  ({x},{y}) gpu obj {int(g['object'][y,x])} tri {int(g['tri'][y,x])} t {g['t'][y,x]!r} n {g['normal'][y,x]} | orc obj {int(o['object'][y,x])} tri {int(o['tri'][y,x])} t {o['t'][y,x]!r} n {o['normal'][y,x]}")

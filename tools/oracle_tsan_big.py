"""Oracle under TSan (tools/sanitize_host.sh): a 262,144-triangle mesh takes the threaded tree build (task subtrees above 65,536 primitives) and the row-parallel render."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mass_raytrace_b200 import scenes
from oracle_backend import OracleScene
tmp = tempfile.mkdtemp()
n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 512, 256, seed=1)
print("tris", n)
w, c = scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0)
s = OracleScene(w, c)
img = s.render(64, 36, 2, seed=1, threads=-4)
print("done", float(img[0].mean()))

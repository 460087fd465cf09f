"""Scene-upload timing: mrt_scene_upload phases (MRT_UPLOAD_TIMING) for repeated uploads of the 1M-triangle workload, then one render."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MRT_UPLOAD_TIMING"] = "1"
from mass_raytrace_b200 import NativeScene, Renderer, scenes
tmp = tempfile.mkdtemp(); n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
w, c = scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0)
t0 = time.perf_counter(); host = NativeScene(w, c); host.desc(); print(f"host scene build (reference-style BVH) {time.perf_counter()-t0:.3f} s", flush=True)
r = Renderer(0)
for k in range(3):
    t0 = time.perf_counter(); r.set_scene(host); r.synchronize(); print(f"set_scene #{k}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
r.reset(1920, 1080); r.accumulate(0, 8); print("render ms", r.stats()["render_ms"])

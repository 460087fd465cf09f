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
for dev in (0, 1):
    r.set_option(Renderer.OPT_DEVICE_BUILD, dev)
    print("---- BLAS built on the", "GPU (LBVH)" if dev else "host (SAH)", flush=True)
    for k in range(3):
        t0 = time.perf_counter(); r.set_scene(host); r.synchronize(); print(f"set_scene #{k}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
    r.reset(1920, 1080); r.accumulate(0, 2)
    r.set_option(Renderer.OPT_TIME_KERNELS, 1); r.reset(1920, 1080); r.accumulate(0, 8); st = r.stats()
    r.set_option(Renderer.OPT_TIME_KERNELS, 0); r.set_option(Renderer.OPT_COUNT_VISITS, 1); r.reset(1920, 1080); r.accumulate(0, 1); cs = r.stats(); r.set_option(Renderer.OPT_COUNT_VISITS, 0)
    print(f"render 8 spp: {st['render_ms']:.2f} ms (extend {st['extend_ms']:.2f}) | nodes/ray {cs['node_visits']/cs['rays']:.2f} tris/ray {cs['tri_tests']/cs['rays']:.2f}", flush=True)

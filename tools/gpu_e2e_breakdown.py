import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mass_raytrace_b200 import NativeScene, Renderer, scenes
w, c = scenes.cornell_box(1.0)
host = NativeScene(w, c)
stream = torch.cuda.Stream()
r = Renderer(0, stream=stream.cuda_stream)
W = H = 1024; spp = 200
out_rgb = torch.empty((H, W, 3), dtype=torch.float32).pin_memory(); out_b = torch.empty((H, W), dtype=torch.int32).pin_memory()
rgb_np, b_np = out_rgb.numpy(), out_b.numpy().view(np.uint32)
def T(f, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); v = f(*a, **k); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, v
for rep in range(3):
    t_up, _ = T(r.set_scene, host)
    t_reset, _ = T(r.reset, W, H)
    t_acc, _ = T(r.accumulate, 0, spp)
    st = r.stats()
    t_dl, _ = T(r.download)
    t_all, _ = T(r.render, W, H, spp, 50, 1, 0, (rgb_np, b_np))
    print(f"upload {t_up:.2f} ms reset {t_reset:.2f} accumulate {t_acc:.2f} (device {st['render_ms']:.2f}) download {t_dl:.2f} | mrt_render total {t_all:.2f} ms", flush=True)

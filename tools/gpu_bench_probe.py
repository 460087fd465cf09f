"""Why does bench.py's device-resident loop read slower than its end-to-end loop? Per-step times of both styles, with the library's own render_ms."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mass_raytrace_b200 import NativeScene, Renderer, scenes
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
use_torch_stream = os.environ.get("PROBE_STREAM", "torch") == "torch"
stream = torch.cuda.Stream(device=dev)
r = Renderer(0, stream=stream.cuda_stream) if use_torch_stream else Renderer(0)
tmp = tempfile.mkdtemp()
n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
w, c = scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0)
host = NativeScene(w, c, defer_mesh_bvh=os.environ.get("PROBE_DEFER", "1") == "1"); host.desc()
r.set_scene(host)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
W, H, spp = 1920, 1080, int(os.environ.get("PROBE_SPP", "256"))
for i in range(6):
    with torch.cuda.stream(stream):
        if os.environ.get("PROBE_FLUSH", "1") == "1": flush.zero_()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    with torch.cuda.stream(stream):
        r.reset(W, H); r.accumulate(i * spp, spp, 50, seed=2024)
    st = r.stats()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    print(f"step {i}: events {e0.elapsed_time(e1):7.2f} ms wall {1e3 * (time.perf_counter() - t0):7.2f} ms library render_ms {st['render_ms']:7.2f} iterations {st['iterations']} launches {st['kernel_launches']}", flush=True)
    if i == 2:
        print("re-upload"); r.set_scene(host)

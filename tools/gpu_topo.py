"""Compare the caller's median-split topology with the library's SAH rebuild: visits per ray and Mrays/s per workload."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mass_raytrace_b200 import NativeScene, Renderer
tmp = tempfile.mkdtemp()
for name, spp in (("cornell", 16), ("book1", 10), ("mesh1m", 8), ("book2", 8)):
    cfg, W, H, _, _ = bench.WORKLOADS[name]
    world, camera = bench.build_workload(name, tmp)
    host = NativeScene(world, camera)
    for keep in (True, False):
        r = Renderer(0)
        t0 = time.time(); r.set_scene(host, keep_topology=keep); up = time.time() - t0
        r.reset(W, H); r.accumulate(0, 1)
        r.set_option(Renderer.OPT_COUNT_VISITS, 1); r.reset(W, H); r.accumulate(0, 2); c = r.stats(); r.set_option(Renderer.OPT_COUNT_VISITS, 0)
        r.set_option(Renderer.OPT_TIME_KERNELS, 1)
        best = None
        for rep in range(2):
            r.reset(W, H); r.accumulate(0, spp); st = r.stats()
            v = (st["rays"] / st["extend_ms"] / 1e3, st["rays"] / st["render_ms"] / 1e3, st["paths"] / st["render_ms"] / 1e3)
            best = v if best is None or v[1] > best[1] else best
        print(f"{name:8s} {'keep ' if keep else 'SAH  '} upload {up:6.2f}s  nodes/ray {c['node_visits']/c['rays']:6.2f} tris/ray {c['tri_tests']/c['rays']:5.2f} "
              f"sph/ray {c['sphere_tests']/c['rays']:5.2f} inst/ray {c['instance_tests']/c['rays']:5.2f} | extend {best[0]:7.1f} Mrays/s  render {best[1]:7.1f} Mrays/s {best[2]:7.1f} Mpaths/s", flush=True)
        r.close()

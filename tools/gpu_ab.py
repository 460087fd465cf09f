"""A/B of CUDA-library builds (tools/_var_*.so and the in-tree one) on a few workloads: render / extend / shade ms and visit counts."""
import glob, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    from mass_raytrace_b200 import NativeScene, Renderer, scenes
    names = sys.argv[2].split(",")
    for name in names:
        keep = name.endswith("+keep")
        base = name.replace("+keep", "")
        if base == "cornell": (w, c), W, H, spp = scenes.cornell_box(1.0), 1024, 1024, 16
        elif base == "menger": (w, c), W, H, spp = scenes.menger(levels=4), 1920, 1080, 8
        elif base == "book2": (w, c), W, H, spp = scenes.book2_final(), 1920, 1080, 8
        elif base == "book1": (w, c), W, H, spp = scenes.book1_spheres(1.5, aperture=0.1), 1200, 800, 32
        elif base == "mesh10m":
            tmp = tempfile.mkdtemp(); paths, mds = [], []
            for i in range(10):
                q = os.path.join(tmp, f"m{i}.ply"); n, md = scenes.write_synthetic_ply(q, 1024, 512, seed=100 + i); paths.append(q); mds.append(md)
            (w, c), W, H, spp = scenes.multi_mesh(paths, mds), 3840, 2160, 2
        elif base == "mesh1m":
            tmp = tempfile.mkdtemp(); n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
            (w, c), W, H, spp = scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080, 8
        r = Renderer(0)
        if os.environ.get("MRT_AB_DEVICE_BUILD"): r.set_option(Renderer.OPT_DEVICE_BUILD, int(os.environ["MRT_AB_DEVICE_BUILD"]))
        r.set_scene(NativeScene(w, c), keep_topology=keep); r.reset(W, H); r.accumulate(0, 2)
        best = None
        for rep in range(3):
            r.reset(W, H); r.accumulate(0, spp); st = r.stats()
            best = st["render_ms"] if best is None else min(best, st["render_ms"])
        r.set_option(Renderer.OPT_TIME_KERNELS, 1); r.reset(W, H); r.accumulate(0, spp); st = r.stats()
        r.set_option(Renderer.OPT_TIME_KERNELS, 0); r.set_option(Renderer.OPT_COUNT_VISITS, 1); r.reset(W, H); r.accumulate(0, 1); cs = r.stats()
        print(f"  {name:12s} render {best:8.2f} ms = {st['rays']/best/1e3:7.1f} Mrays/s | extend {st['extend_ms']:7.2f} shade {st['shade_ms']:6.2f} gen {st['generate_ms']:5.2f} iters {st['iterations']} | "
              f"nodes/ray {cs['node_visits']/cs['rays']:5.2f} tris/ray {cs['tri_tests']/cs['rays']:4.2f} inst/ray {cs['instance_tests']/cs['rays']:4.2f}", flush=True)
        r.close()
else:
    names = sys.argv[1] if len(sys.argv) > 1 else "cornell,menger,menger+keep"
    libs = sorted(glob.glob(os.path.join(ROOT, "tools", "_var_*.so"))) + [os.path.join(ROOT, "mass_raytrace_b200", "libmrt_cuda.so")]
    for lib in libs:
        print(os.path.basename(lib), flush=True)
        subprocess.run([sys.executable, __file__, "--child", names], env=dict(os.environ, MRT_CUDA_LIB=lib))

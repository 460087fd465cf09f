"""Time kernel-variant builds (tools/_var_*.so) on two workloads: per-kernel ms and Mrays/s."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    from mass_raytrace_b200 import NativeScene, Renderer, scenes
    import tempfile
    work = {"cornell": (scenes.cornell_box(1.0), 1024, 1024, 16), "book1": (scenes.book1_spheres(1.5, 0.1), 1200, 800, 10)}
    if os.environ.get("MRT_VAR_MESH"):
        tmp = tempfile.mkdtemp(); n, md = scenes.write_synthetic_ply(os.path.join(tmp, "m.ply"), 1024, 512, seed=1)
        work["mesh1m"] = (scenes.lucy_layout(os.path.join(tmp, "m.ply"), md, grid=0), 1920, 1080, 16)
    for name, ((w, c), W, H, spp) in work.items():
        r = Renderer(0); r.set_scene(NativeScene(w, c)); r.reset(W, H); r.accumulate(0, 2)
        r.set_option(Renderer.OPT_TIME_KERNELS, 1)
        best = None
        for rep in range(3):
            r.reset(W, H); r.accumulate(0, spp); st = r.stats()
            best = st if best is None or st["render_ms"] < best["render_ms"] else best
        print(f"  {name:8s} render {best['render_ms']:7.2f} ms  extend {best['extend_ms']:7.2f}  shade {best['shade_ms']:6.2f}  generate {best['generate_ms']:5.2f}  => {best['rays']/best['render_ms']/1e3:7.1f} Mrays/s", flush=True)
        r.close()
else:
    for lib in sorted(glob.glob(os.path.join(ROOT, "tools", "_var_*.so"))):
        print(os.path.basename(lib), flush=True)
        subprocess.run([sys.executable, __file__, "--child"], env=dict(os.environ, MRT_CUDA_LIB=lib))

"""First-contact GPU probe: AOV parity vs oracle on two scenes, a statistical render comparison, and rough throughput."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from mass_raytrace_b200 import NativeScene, Renderer, scenes
from oracle_backend import OracleScene

def aov_compare(name, world, camera, w, h):
    host = NativeScene(world, camera); orc = OracleScene(world, camera)
    r = Renderer(0); r.set_scene(host)
    t0 = time.time(); g = r.render_aov(w, h); tg = time.time() - t0
    t0 = time.time(); o = orc.render_aov(w, h); to = time.time() - t0
    res = {k: int((g[k] != o[k]).sum()) for k in ("object", "tri")}
    tt = (g["t"] != o["t"]) & ~(np.isinf(g["t"]) & np.isinf(o["t"]))
    res["t_bits"] = int(tt.sum()); res["normal_bits"] = int((g["normal"] != o["normal"]).any(-1).sum())
    res["albedo_diff"] = int((g["albedo"] != o["albedo"]).any(-1).sum())
    fin = np.isfinite(o["t"])
    res["max_rel_t"] = float(np.max(np.abs(g["t"][fin] - o["t"][fin]) / o["t"][fin])) if fin.any() else 0.0
    print(name, w, h, res, f"gpu {tg:.3f}s oracle {to:.3f}s", flush=True)
    r.close()
    return res

def render_compare(name, world, camera, w, h, spp, depth=50):
    host = NativeScene(world, camera); orc = OracleScene(world, camera)
    r = Renderer(0); r.set_scene(host)
    rgb, b, cnt = r.render(w, h, spp, depth, seed=11)
    st = r.stats()
    t0 = time.time(); orgb, ob, oc = orc.render(w, h, spp, depth, seed=11); to = time.time() - t0
    orgb2, ob2, _ = orc.render(w, h, spp, depth, seed=12)
    Y = np.array([0.2126, 0.7152, 0.0722], np.float32)
    lum = lambda a: float((a * Y).sum(-1).mean()) / spp
    rm = lambda a, c: float(np.sqrt(np.mean((a / spp - c / spp) ** 2)))
    print(name, f"lum gpu {lum(rgb):.5f} orc {lum(orgb):.5f} orc2 {lum(orgb2):.5f} | rmse g-o {rm(rgb, orgb):.5f} o-o {rm(orgb, orgb2):.5f} | "
          f"bounces gpu {b.mean()/spp:.4f} orc {ob.mean()/spp:.4f} orc2 {ob2.mean()/spp:.4f} | rays gpu {st['rays']} orc {oc['rays']} | "
          f"gpu {st['render_ms']:.1f} ms ({st['rays']/st['render_ms']/1e3:.1f} Mrays/s) oracle {to:.2f}s ({oc['rays']/to/1e6:.2f} Mrays/s)", flush=True)
    r.close()

def throughput(name, world, camera, w, h, spp, count=False):
    host = NativeScene(world, camera)
    r = Renderer(0); r.set_scene(host)
    r.reset(w, h); r.accumulate(0, 1)  # warm-up
    r.set_option(Renderer.OPT_TIME_KERNELS, 1)
    if count: r.set_option(Renderer.OPT_COUNT_VISITS, 1)
    r.reset(w, h); r.accumulate(0, spp)
    st = r.stats()
    print(name, json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}),
          f"=> {st['paths']/st['render_ms']/1e3:.1f} Mpaths/s {st['rays']/st['render_ms']/1e3:.1f} Mrays/s; extend-only {st['rays']/max(st['extend_ms'],1e-9)/1e3:.1f} Mrays/s", flush=True)
    r.close()

if __name__ == "__main__":
    import __graft_entry__ as ge
    ge.smoke()
    wc, cc = scenes.cornell_box(1.0)
    aov_compare("cornell", wc, cc, 512, 512)
    wb, cb = scenes.book1_spheres(1.5, aperture=0.0)
    aov_compare("book1", wb, cb, 600, 400)
    render_compare("cornell", wc, cc, 128, 128, 64)
    wb2, cb2 = scenes.book1_spheres(1.5, aperture=0.1)
    render_compare("book1", wb2, cb2, 150, 100, 64)
    throughput("cornell 1024^2 x16", wc, cc, 1024, 1024, 16)
    throughput("cornell 1024^2 x16 counted", wc, cc, 1024, 1024, 16, count=True)
    throughput("book1 1200x800 x10", wb2, cb2, 1200, 800, 10)
    throughput("book1 1200x800 x10 counted", wb2, cb2, 1200, 800, 10, count=True)

"""More than one GPU behind the C ABI (include/mrt.h, "more than one GPU"): a multi-device handle splits the samples of every pixel
over its devices, merges the exact accumulators with one NCCL reduce and must give, bit for bit, the image one GPU renders --
the reference's thread fan-out + Image::merge (main.rs:159-170, 235-294, 629-638). Runs with -m gpu; the N > 1 cases skip on a
one-GPU box (the scaling bench prints the same check at N = 2 / 4 / 8, bench.py `parity`)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mass_raytrace_b200 import NativeScene, Renderer, World, Camera, scenes
from mass_raytrace_b200 import api, _ffi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 1


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_single_device_list_is_a_plain_context(renderer):
    world, camera = scenes.cornell_box(1.0)
    host = NativeScene(world, camera)
    renderer.set_scene(host)
    ref = renderer.render(64, 64, 8, 50, seed=3)
    r = Renderer([0])
    try:
        assert r.comm_rank() == (0, 1)
        r.set_scene(host)
        got = r.render(64, 64, 8, 50, seed=3)
        assert np.array_equal(bits(got[0]), bits(ref[0])) and np.array_equal(got[1], ref[1]) and got[2] == ref[2] == 8
        r.comm_reduce()  # a no-op without a communicator
        assert r.download()[2] == 8
    finally:
        r.close()


def test_multi_handle_rejects_bad_device_lists():
    lib = _ffi.cuda_lib()
    h = C.c_void_p()
    assert lib.mrt_context_create_multi((C.c_int * 2)(0, 0), 2, C.byref(h)) == -1 and b"twice" in lib.mrt_last_error(None)
    assert lib.mrt_context_create_multi(None, 0, C.byref(h)) == -1
    assert lib.mrt_context_create_multi((C.c_int * 1)(99), 1, C.byref(h)) == -1


def nonfinite_edge_scene():
    """A small light of infinite radiance: pixels on its silhouette receive a non-finite sample only from SOME jittered samples, so
    under a split the flag is raised on one device and must survive the merge (the reference's f32 sum would be poisoned, main.rs:634)."""
    w = World(api.SolidBackground((0.1, 0.2, 0.3)))
    w.add(api.Sphere(api.DiffuseLight((float("inf"), 1.0, 1.0)), (0.0, 0.0, 0.0), 0.5))
    w.add(api.Sphere(api.Lambertian(api.SolidColor((0.5, 0.5, 0.5, 1.0))), (0.0, -100.5, 0.0), 100.0))
    w.build_bvh()
    cam = Camera(40.0, (0.0, 0.5, 3.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 1.0, 0.0, 3.0)
    return w, cam


@pytest.mark.parametrize("n", [2, 4, 8])
def test_n_devices_in_one_process_equal_one_device(renderer, n, tmp_mesh_dir):
    if device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    r = Renderer(list(range(n)))
    try:
        assert r.comm_rank() == (0, n)
        cases = [("cornell", scenes.cornell_box(1.0), 96, 96, 13), ("nonfinite", nonfinite_edge_scene(), 64, 64, 6)]
        path = str(tmp_mesh_dir / "multi_mesh.ply")
        ntri, md = scenes.write_synthetic_ply(path, 256, 128, seed=4)  # 65,536 triangles: its BLAS is built on device 0 and copied over NVLink
        cases.append(("mesh", scenes.lucy_layout(path, md, grid=0), 160, 90, 5))
        for name, (world, camera), w, h, spp in cases:
            host = NativeScene(world, camera)
            renderer.set_scene(host)
            ref = renderer.render(w, h, spp, 50, seed=11)
            r.set_scene(host)
            got = r.render(w, h, spp, 50, seed=11)  # split over the devices, one NCCL reduce, download from device 0
            assert got[2] == ref[2] == spp, name
            assert np.array_equal(bits(got[0]), bits(ref[0])), name  # NaN channels included
            assert np.array_equal(got[1], ref[1]), name
            st, st1 = r.stats(), renderer.stats()
            assert st["paths"] == st1["paths"] == w * h * spp and st["rays"] == st1["rays"], name
            if name == "nonfinite":
                assert np.isnan(ref[0][..., 0]).any() and not np.isnan(ref[0][..., 0]).all()
            # a second merge into the same image keeps accumulating (count and sums), like further Image::merge calls
            r.accumulate(spp, 3, 50, seed=11)
            renderer.accumulate(spp, 3, 50, seed=11)
            a, b = r.download(), renderer.download()
            assert a[2] == b[2] == spp + 3 and np.array_equal(bits(a[0]), bits(b[0])) and np.array_equal(a[1], b[1]), name
        # the tone-mapped bytes too (resolved on device 0)
        assert np.array_equal(r.resolve_rgb8(a[2]), renderer.resolve_rgb8(b[2]))
        # explicit ranges + explicit merge (MRT_OPT_COMM_SPLIT = 0 is for one-process-per-GPU hosts; on a multi-device handle the
        # local render runs on device 0 only and the merge is then a sum with empty images)
        r.set_option(Renderer.OPT_COMM_SPLIT, 0)
        r.reset(w, h)
        r.accumulate(0, spp, 50, seed=11)
        r.comm_reduce()
        c = r.download()
        assert c[2] == spp and np.array_equal(bits(c[0]), bits(ref[0]))
        r.set_option(Renderer.OPT_COMM_SPLIT, 1)
    finally:
        r.close()


def test_c_host_on_two_gpus_writes_the_same_image(tmp_path):
    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pkg = os.path.join(ROOT, "mass_raytrace_b200")
    exe = str(tmp_path / "cornell")
    subprocess.run(["gcc", "-std=c11", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cornell.c"), "-L", pkg, "-lmrt_host",
                    "-lmrt_cuda", "-lm", f"-Wl,-rpath,{pkg}", "-o", exe], check=True, capture_output=True, text=True)
    cube = os.path.join(pkg, "assets", "cube.ply")
    outs = []
    for n in (1, 2):
        p = subprocess.run([exe, cube, "128", "128", "9", str(tmp_path / f"c{n}.ppm"), str(n)], capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr
        outs.append(p.stdout.split()[-1])  # fnv1a of the rgb8 image
    assert outs[0] == outs[1]


def test_one_process_per_gpu_collective_upload_and_render(renderer, tmp_path):
    """torchrun-style hosts: one context per process joined by mrt_comm_init_rank. mrt_scene_upload is collective there (rank 0 builds
    and uploads, ncclBroadcast over NVLink, the other ranks pass NULL), mrt_render splits and merges; rank 0's image must be the
    single-GPU image byte for byte, and a scene the root rejects must fail on every rank."""
    import hashlib
    import sys

    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out, mesh = str(tmp_path / "out"), str(tmp_path / "m.ply")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "multi_process_child.py"), out, mesh], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    got = dict((ln.split()[0], ln.split()[1:]) for ln in open(out + ".0").read().splitlines())
    n, md = scenes.write_synthetic_ply(mesh, 256, 128, seed=4)
    for name, (w, c), W, H, spp in (("cornell", scenes.cornell_box(1.0), 96, 96, 13), ("mesh", scenes.lucy_layout(mesh, md, grid=0), 160, 90, 5)):
        renderer.set_scene(NativeScene(w, c))
        rgb, b, cnt = renderer.render(W, H, spp, 50, seed=11)
        assert got[name][0] == hashlib.sha256(rgb.tobytes() + b.tobytes()).hexdigest(), name
    for rank in (0, 1):
        rej = [ln for ln in open(f"{out}.{rank}").read().splitlines() if ln.startswith("rejected")][0]
        assert "abi_version" in rej or "root rank" in rej, rej

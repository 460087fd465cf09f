"""A host written in plain C (examples/cornell.c) against include/mrt.h + include/mrt_host.h: the headers are valid C11, both
libraries link from C, a box without a GPU gets an error and an exit code instead of an image, and on the GPU the image is the one
the Python host gets for the same scene -- the boundary carries no language-specific state."""
import os
import subprocess

import numpy as np
import pytest

from mass_raytrace_b200 import NativeScene, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mass_raytrace_b200")
CUBE = os.path.join(PKG, "assets", "cube.ply")


@pytest.fixture(scope="module")
def cornell_exe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("c_host") / "cornell")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cornell.c"),
           "-L", PKG, "-lmrt_host", "-lmrt_cuda", "-lm", f"-Wl,-rpath,{PKG}", "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_headers_are_plain_c():
    for header in ("mrt.h", "mrt_host.h"):
        subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", header)],
                       check=True, capture_output=True, text=True)


def test_c_host_without_a_gpu_reports_and_exits(cornell_exe, tmp_path):
    out = str(tmp_path / "never.ppm")
    p = subprocess.run([cornell_exe, CUBE, "32", "32", "1", out], capture_output=True, text=True, timeout=120)
    if p.returncode == 0:
        pytest.skip("a CUDA device is present")
    assert p.returncode == 3 and "mrt_context_create" in p.stderr and "no CPU fallback" in p.stderr
    assert not os.path.exists(out)
    assert subprocess.run([cornell_exe], capture_output=True).returncode == 1  # usage
    assert subprocess.run([cornell_exe, "/nonexistent.ply", "32", "32", "1", out], capture_output=True).returncode == 2  # scene error, before any GPU call


@pytest.mark.gpu
def test_c_host_renders_the_image_the_python_host_renders(cornell_exe, renderer, tmp_path):
    w, h, spp = 256, 256, 16
    out = str(tmp_path / "cornell.ppm")
    p = subprocess.run([cornell_exe, CUBE, str(w), str(h), str(spp), out], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    fields = p.stdout.split()
    stats = dict(zip(fields[0::2], fields[1::2]))
    assert int(stats["paths"]) == w * h * spp and int(stats["rays"]) > int(stats["paths"])
    with open(out, "rb") as f:
        assert f.readline() == b"P6\n" and f.readline() == f"{w} {h}\n".encode() and f.readline() == b"255\n"
        img_c = np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)

    world, camera = scenes.cornell_box(1.0)
    renderer.set_scene(NativeScene(world, camera, defer_mesh_bvh=True))
    renderer.render(w, h, spp, 50, seed=1)
    img_py = renderer.resolve_rgb8(spp, mode=0, flip=True)
    assert np.array_equal(img_c, img_py)
    assert img_c.mean() > 20 and img_c[..., 0].mean() != img_c[..., 1].mean()  # a lit box with a red and a green wall, not a blank frame

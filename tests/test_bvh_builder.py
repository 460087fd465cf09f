"""The host's SAH builder (mrt_bvh_build.cpp: SSE binning, out-of-place partition, parallel slices and task subtrees) checked on the
CPU by a small native program: coverage, containment, leaf sizes, degenerate inputs, and agreement of repeated (multi-threaded) builds."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sah_builder_invariants(tmp_path):
    exe = str(tmp_path / "check_bvh_build")
    csrc = os.path.join(ROOT, "mass_raytrace_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", csrc, "-o", exe, os.path.join(ROOT, "tests", "native", "check_bvh_build.cpp"),
                    os.path.join(csrc, "mrt_bvh_build.cpp")], check=True)
    a = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    b = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert a == b  # same tree shape whatever the thread interleaving
    lines = a.strip().splitlines()
    assert len(lines) == 6 and all("bad=0 missing=0" in ln for ln in lines), a
    assert lines[0].startswith("n=1 leaves=1 depth=0")

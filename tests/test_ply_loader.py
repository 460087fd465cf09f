"""PlyLoader (reference src/ply_loader.rs): the host library's reader and the oracle's restatement on the same files --
ascii, binary little/big endian, skipped properties and elements, quads silently dropped, header errors."""
import os
import struct

import numpy as np
import pytest

from mass_raytrace_b200 import Model, NativeScene, PlyLoader, SolidBackground, V3, World, scenes
from oracle_backend import OracleScene


def _load(path, backend, perm=(0, 1, 2)):
    w = World(SolidBackground(V3(0, 0, 0)))
    t = PlyLoader.load(path, vertex_perm=perm)
    w.add(Model(t))
    s = backend(w)
    return s.mesh_verts(t), s.mesh_max_abs[id(t)]


BACKENDS = [NativeScene, OracleScene]


@pytest.mark.parametrize("backend", BACKENDS)
def test_cube(backend):
    v, m = _load(scenes.CUBE_PLY, backend)
    assert v.shape == (12, 9) and m == 1.0
    assert v[1].tolist() == [-1, 1, -1, 1, -1, -1, -1, -1, -1]  # "3 1 3 4"
    v2, _ = _load(scenes.CUBE_PLY, backend, perm=(1, 2, 0))  # the Lucy swizzle V3::new(y, z, x), scenes/lucy.rs:38
    assert np.array_equal(v2.reshape(-1, 3), v.reshape(-1, 3)[:, [1, 2, 0]])


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_formats_agree(tmp_path, fmt):
    ref_path, path = str(tmp_path / "ref.ply"), str(tmp_path / f"{fmt}.ply")
    n, md = scenes.write_synthetic_ply(ref_path, 12, 7, seed=5)
    n2, md2 = scenes.write_synthetic_ply(path, 12, 7, seed=5, fmt=fmt)
    assert n == n2 == 2 * 12 * 7
    ref, m0 = _load(ref_path, NativeScene)
    for backend in BACKENDS:
        v, m = _load(path, backend)
        assert np.array_equal(v, ref), (fmt, backend.__name__)  # ascii round-trips through repr(): exact
        assert m == m0 == np.float32(md)


def _write(path, header, body=b""):
    with open(path, "wb") as f:
        f.write(header.encode() + body)


@pytest.mark.parametrize("backend", BACKENDS)
def test_mixed_properties_and_quads(tmp_path, backend):
    # extra vertex properties (uchar colours, double nx) are skipped (:346-348); a face with 4 indices is dropped (:396-415);
    # an unknown element ("edge") is read past; face lists may be ushort counts with short indices
    p = str(tmp_path / "mixed.ply")
    hdr = ("ply\nformat binary_little_endian 1.0\ncomment x\nelement vertex 4\nproperty float x\nproperty uchar red\nproperty float y\nproperty double nx\n"
           "property float z\nelement edge 1\nproperty int a\nproperty int b\nelement face 3\nproperty list ushort short vertex_indices\nend_header\n")
    verts = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    body = b"".join(struct.pack("<fBfdf", x, 7, y, 0.5, z) for x, y, z in verts)
    body += struct.pack("<ii", 0, 1)
    body += struct.pack("<Hhhh", 3, 0, 1, 2) + struct.pack("<Hhhhh", 4, 0, 1, 2, 3) + struct.pack("<Hhhh", 3, 0, 2, 3)
    _write(p, hdr, body)
    v, m = _load(p, backend)
    assert v.tolist() == [[0, 0, 0, 1, 0, 0, 0, 1, 0], [0, 0, 0, 0, 1, 0, 0, 0, 1]]
    assert m == 1.0


@pytest.mark.parametrize("backend", BACKENDS)
def test_ascii_with_double_coordinates_and_int_vertices(tmp_path, backend):
    p = str(tmp_path / "a.ply")
    _write(p, "ply\nformat ascii 1.0\nelement vertex 3\nproperty double x\nproperty int y\nproperty float z\nelement face 1\n"
              "property list uchar uint vertex_indices\nend_header\n0.1 2 1e-3\n-4.5 -7 .25\n3 0 0\n3 2 1 0\n")
    v, m = _load(p, backend)
    assert v[0].tolist() == [3, 0, 0, np.float32(-4.5), -7, 0.25, np.float32(0.1), 2, np.float32(1e-3)]
    assert m == 7.0


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("header,msg", [
    ("plx\nformat ascii 1.0\nend_header\n", "magic"),
    ("ply\nformat ascii 2.0\nend_header\n", "unsupported format"),
    ("ply\nformat binary_middle_endian 1.0\nend_header\n", "unsupported format"),
    ("ply\nformat ascii 1.0\nelement vertex\nend_header\n", "invalid element"),
    ("ply\nformat ascii 1.0\nelement vertex 1\nproperty quad x\nend_header\n", "invalid property"),
    ("ply\nformat ascii 1.0\nelement vertex 1\nproperty list uchar\nend_header\n", "invalid property"),
])
def test_header_errors(tmp_path, backend, header, msg):
    p = str(tmp_path / "bad.ply")
    _write(p, header)
    with pytest.raises(RuntimeError, match=msg):
        _load(p, backend)


@pytest.mark.parametrize("backend", BACKENDS)
def test_truncated_and_out_of_range(tmp_path, backend):
    p = str(tmp_path / "t.ply")
    hdr = "ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n"
    body = struct.pack("<9f", *range(9))
    _write(p, hdr, body + struct.pack("<Bii", 3, 0, 1))  # file ends inside the face
    with pytest.raises(RuntimeError, match="read error"):
        _load(p, backend)
    _write(p, hdr, body + struct.pack("<Biii", 3, 0, 1, 3))  # index 3 of 3 vertices (Rust: index panic)
    with pytest.raises(RuntimeError, match="out of bounds"):
        _load(p, backend)
    _write(p, hdr.replace("element face 1", "element face 0"), body)  # no faces -> empty mesh is rejected by Model::new's BvhNode (unreachable!() geom.rs:153)
    with pytest.raises(RuntimeError):
        _load(p, backend)
    with pytest.raises(RuntimeError, match="cannot open"):
        _load(str(tmp_path / "missing.ply"), backend)


# ---- StlLoader (reference src/stl_loader.rs:10-66) --------------------------------------------------------------------------
def _load_stl(path, backend, perm=(0, 1, 2)):
    from mass_raytrace_b200 import StlLoader

    w = World(SolidBackground(V3(0, 0, 0)))
    t = StlLoader.load_binary(path, vertex_perm=perm)
    w.add(Model(t))
    return backend(w).mesh_verts(t)


def _write_stl(path, tris, attr_bytes=0, truncate=0):
    data = b"binary stl written by the tests".ljust(80, b"\0") + struct.pack("<I", len(tris))
    for k, t in enumerate(tris):
        data += struct.pack("<3f", 0.0, 0.0, 1.0) + struct.pack("<9f", *t)
        n = attr_bytes if k % 2 == 0 else 0
        data += struct.pack("<H", n) + b"\xAB" * n  # the attribute count is a BYTE count that the loader skips (:58-61)
    with open(path, "wb") as f:
        f.write(data[:len(data) - truncate] if truncate else data)


@pytest.mark.parametrize("backend", BACKENDS)
def test_stl_binary(tmp_path, backend):
    p = str(tmp_path / "m.stl")
    tris = [[0, 0, 0, 1, 0, 0, 0, 1, 0], [0, 0, 1, 2, 0, 1, 0, 3, 1], [5, 6, 7, 8, 9, 10, 11, 12, 14]]
    _write_stl(p, tris)
    assert _load_stl(p, backend).tolist() == tris
    _write_stl(p, tris, attr_bytes=6)  # non-zero attribute blocks are skipped
    assert _load_stl(p, backend).tolist() == tris
    assert _load_stl(p, backend, perm=(1, 2, 0))[2].tolist() == [6, 7, 5, 9, 10, 8, 12, 14, 11]
    _write_stl(p, tris, truncate=10)
    with pytest.raises(RuntimeError, match="read error"):
        _load_stl(p, backend)
    _write_stl(p, [])
    with pytest.raises(RuntimeError):  # Model::new over an empty Vec: BvhNode::new hits unreachable!() (geom.rs:153)
        _load_stl(p, backend)


# ---- corrupted files: a loader returns triangles or an error, it never crashes or reads out of bounds -----------------------------
# (the reference's loaders return Result / panic on bad input; tools/sanitize_host.sh runs this file under ASan + UBSan)
FUZZ = 150 * int(os.environ.get("MRT_FUZZ_SCALE", "1"))  # tools/sanitize_host.sh can be run with a larger scale


def _corruptions(data, rng, n):
    for _ in range(n):
        b = bytearray(data)
        kind = rng.integers(0, 4)
        if kind == 0:  # flip bytes
            for i in rng.integers(0, len(b), rng.integers(1, 6)):
                b[i] = int(rng.integers(0, 256))
        elif kind == 1:  # truncate
            b = b[:int(rng.integers(0, len(b)))]
        elif kind == 2:  # drop a slice from the middle
            i = int(rng.integers(0, len(b))); j = min(len(b), i + int(rng.integers(1, 40)))
            b = b[:i] + b[j:]
        else:  # overwrite a run with 0xFF (huge counts, NaNs, negative indices)
            i = int(rng.integers(0, len(b))); j = min(len(b), i + int(rng.integers(1, 9)))
            b[i:j] = b"\xFF" * (j - i)
        yield bytes(b)


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_corrupted_ply_never_crashes(tmp_path, fmt):
    src = str(tmp_path / "good.ply")
    scenes.write_synthetic_ply(src, 6, 4, seed=2, fmt=fmt)
    data = open(src, "rb").read()
    rng = np.random.default_rng(11)
    loaded = failed = 0
    for k, bad in enumerate(_corruptions(data, rng, FUZZ)):
        p = str(tmp_path / "bad.ply")
        open(p, "wb").write(bad)
        try:
            v, _ = _load(p, NativeScene)
            assert v.ndim == 2 and v.shape[1] == 9 and len(v) >= 1
            loaded += 1
        except RuntimeError as e:
            assert str(e)
            failed += 1
    assert failed > 20 and loaded + failed == FUZZ  # most corruptions are detected; a flipped coordinate byte still loads


def test_corrupted_stl_never_crashes(tmp_path):
    p = str(tmp_path / "good.stl")
    tris = np.random.default_rng(5).normal(size=(40, 9)).astype(np.float32).tolist()
    _write_stl(p, tris, attr_bytes=4)
    data = open(p, "rb").read()
    rng = np.random.default_rng(12)
    loaded = failed = 0
    for bad in _corruptions(data, rng, FUZZ):
        open(p, "wb").write(bad)
        try:
            v = _load_stl(p, NativeScene)
            assert v.ndim == 2 and v.shape[1] == 9 and len(v) >= 1
            loaded += 1
        except RuntimeError as e:
            assert str(e)
            failed += 1
    assert failed > 20 and loaded + failed == FUZZ

"""ObjLoader + SimpleTexturedBuilder / obj_fns (reference src/obj_loader.rs) and Texture::load_png (src/texture.rs:29): the host
library's parser and PNG decoder against the oracle's restatement (which gets its PNGs decoded by PIL) on the same files."""
import os
import struct
import zlib

import numpy as np
import pytest
from PIL import Image

from mass_raytrace_b200 import (WRAP_CLAMP, WRAP_REPEAT, Lambertian, Model, NativeScene, ObjFns, ObjLoader, SimpleTexturedBuilder, SolidBackground, SolidColor,
                                TextureFile, V3, World)
from oracle_backend import OracleScene

BACKENDS = [NativeScene, OracleScene]

OBJ = """# two quads, one textured, one flat-coloured, plus a filtered group
mtllib scene materials.mtl
o floor
v 0 0 0
v 1 0 0
v 1 0 1
v 0 0 1
vn 0 1 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
usemtl checker
f 1/1/1 2/2/1 3/3/1 4/4/1
f 1/1/1 3/3/1 4/4/1
g wall
v 0 1 0
v 1 1 0
vn 0 0 1
usemtl red
f 1//2 2//2 6//2
f 1/1/2 6/3/2 5/4/2 extra/ignored
o hidden
usemtl red
f 1/1/1 2/2/1 3/3/1
"""
MTL = """newmtl checker
Kd 0.1 0.2 0.3
map_Kd tex.png
newmtl red
Kd 0.8 0.1 1e-1
Ns 10
newmtl unused
Kd oops 0 0
"""


def _write_scene(d, png_mode="RGBA"):
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (5, 7, 4), dtype=np.uint8)
    img[0, 0, 3] = 0
    Image.fromarray(img, "RGBA").convert(png_mode).save(os.path.join(d, "tex.png"))
    open(os.path.join(d, "scene materials.mtl"), "w").write(MTL)
    open(os.path.join(d, "model.obj"), "w").write(OBJ)
    return os.path.join(d, "model.obj")


def _load(backend, path, builder):
    w = World(SolidBackground(V3(0, 0, 0)))
    t = ObjLoader.load(path, builder)
    w.add(Model(t))
    s = backend(w)
    nrm, uv, mat = s.mesh_shading(t)
    return s.mesh_verts(t), nrm, uv, [s.material_info(m) for m in mat]


@pytest.mark.parametrize("backend", BACKENDS)
def test_simple_textured_builder(tmp_path, backend):
    path = _write_scene(str(tmp_path))
    v, nrm, uv, mats = _load(backend, path, SimpleTexturedBuilder.with_filter(WRAP_CLAMP, ["hidden"]))
    assert v.shape == (4, 9)  # quads contribute their first three corners only (obj_loader.rs:420-422); group `hidden` is filtered
    assert v[0].tolist() == [0, 0, 0, 1, 0, 0, 1, 0, 1]
    assert uv[0].tolist() == [0, 1, 1, 1, 1, 0]  # v -> 1 - v (build_uv :281-283)
    assert uv[2].tolist() == [0, 1, 0, 1, 0, 1]  # `v//n`: the FIRST vt of the file stands in for every corner (:404)
    assert nrm[2].tolist() == [0, 0, 1] * 3 and nrm[0].tolist() == [0, 1, 0] * 3
    kinds = [m[0] for m in mats]
    assert kinds == [1, 1, 1, 1]  # Lambertian
    assert mats[0][2] == (7, 5) and mats[0][3] != 0 and mats[1][3] == mats[0][3]  # map_Kd wins over Kd
    assert mats[2][2] == (0, 0) and np.allclose(mats[2][1], (0.8, 0.1, 0.1, 1.0), rtol=0, atol=1e-7)


def test_host_matches_oracle(tmp_path):
    for mode in ("RGBA", "RGB", "L", "LA", "P"):
        path = _write_scene(str(tmp_path / mode), mode)
        a = _load(NativeScene, path, SimpleTexturedBuilder(WRAP_REPEAT))
        b = _load(OracleScene, path, SimpleTexturedBuilder(WRAP_REPEAT))
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x, y)
        assert a[3] == b[3], mode  # kind, colour, texture size and the hash of the decoded f32 texels (own inflate vs PIL)
        assert len(a[0]) == 5  # nothing filtered


@pytest.mark.parametrize("backend", BACKENDS)
def test_obj_fns_builder(tmp_path, backend):
    path = _write_scene(str(tmp_path))
    red = Lambertian(SolidColor((1, 0, 0, 1)))
    v, nrm, uv, mats = _load(backend, path, ObjFns(red))
    assert len(v) == 5 and uv[0].tolist() == [0, 0, 1, 0, 1, 1]  # uv as written, no filter, one material
    assert all(m[0] == 1 and m[1] == (1.0, 0.0, 0.0, 1.0) for m in mats)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("text,msg", [
    ("v 0 0\n", "unable to parse vertex"),
    ("v 0 0 0x10\n", "unable to parse vertex"),
    ("vn 0 a 0\n", "unable to parse normal"),
    ("vt 0.5\n", "unable to parse texture coord"),
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nf 1/1/1 2/1/1\n", "unable to parse face"),
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nf 1/1/1 2/1/1 4/1/1\n", "unable to parse face"),  # index past the end
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nf 1 2 3\n", "unable to parse face"),  # no normal / uv indices
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n", "unable to parse face"),  # `//` needs uvs[0] (:404)
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nf 0/1/1 2/1/1 3/1/1\n", "unable to parse face"),  # index 0
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nusemtl nope\nf 1/1/1 2/1/1 3/1/1\n", "No material found for face"),
    ("mtllib missing.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nusemtl a\nf 1/1/1 2/1/1 3/1/1\n", "No material found for face"),
    ("v 0 0 0\n", "mesh has no triangles"),
])
def test_errors(tmp_path, backend, text, msg):
    p = str(tmp_path / "bad.obj")
    open(p, "w").write(text)
    with pytest.raises(RuntimeError, match=msg):
        _load(backend, p, SimpleTexturedBuilder(WRAP_REPEAT))
    with pytest.raises(RuntimeError, match="cannot open"):
        _load(backend, str(tmp_path / "absent.obj"), SimpleTexturedBuilder(WRAP_REPEAT))


def _png(w, h, ctype, depth, rows, extra=b"", level=6, interlace=0):
    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))
    raw = b"".join(rows)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace)) + extra +
            chunk(b"IDAT", zlib.compress(raw, level)) + chunk(b"IEND", b""))


def _texels(backend, path):
    from mass_raytrace_b200 import Sphere
    w = World(SolidBackground(V3(0, 0, 0)))
    w.add(Sphere(Lambertian(TextureFile(path, WRAP_REPEAT)), V3(0, 0, 0), 1.0))
    s = backend(w)
    return s.material_info(0)


def test_png_decoder_against_pil(tmp_path):
    rng = np.random.default_rng(11)
    cases = []
    # every filter type on RGBA rows, stored (level 0), fixed and dynamic Huffman blocks
    img = rng.integers(0, 256, (6, 9, 4), dtype=np.uint8)
    for level in (0, 1, 9):
        rows = []
        prev = np.zeros(9 * 4, np.int32)
        for y in range(6):
            cur = img[y].reshape(-1).astype(np.int32)
            f = y % 5
            left = np.concatenate([np.zeros(4, np.int32), cur[:-4]])
            ul = np.concatenate([np.zeros(4, np.int32), prev[:-4]])
            if f == 0: enc = cur
            elif f == 1: enc = cur - left
            elif f == 2: enc = cur - prev
            elif f == 3: enc = cur - (left + prev) // 2
            else:
                p = left + prev - ul
                pa, pb, pc = abs(p - left), abs(p - prev), abs(p - ul)
                pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
                enc = cur - pred
            rows.append(bytes([f]) + (enc & 255).astype(np.uint8).tobytes())
            prev = cur
        cases.append((f"rgba_l{level}.png", _png(9, 6, 6, 8, rows, level=level)))
    # palette with tRNS at 2 bits per pixel, grey at 1 and 4 bits, grey + colour key, RGB + colour key
    plte = bytes(range(30, 42))
    rows = [b"\x00" + bytes([0b00011011, 0b11100100]) for _ in range(3)]
    chunkf = lambda tag, data: struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))
    cases.append(("pal2.png", _png(8, 3, 3, 2, rows, extra=chunkf(b"PLTE", plte) + chunkf(b"tRNS", bytes([0, 128])))))
    cases.append(("grey1.png", _png(10, 2, 0, 1, [b"\x00\xaa\x80", b"\x00\x55\x40"])))
    cases.append(("grey4.png", _png(3, 2, 0, 4, [b"\x00\x1f\x70", b"\x00\xa5\xc0"])))
    cases.append(("rgbkey.png", _png(2, 1, 2, 8, [b"\x00" + bytes([1, 2, 3, 9, 8, 7])], extra=chunkf(b"tRNS", struct.pack(">HHH", 9, 8, 7)))))
    big = rng.integers(0, 4, (64, 300, 3), dtype=np.uint8) * 60  # compressible: long matches, distances > 256
    Image.fromarray(big, "RGB").save(str(tmp_path / "big.png"))
    for name, data in cases:
        open(str(tmp_path / name), "wb").write(data)
    for name in [c[0] for c in cases] + ["big.png"]:
        p = str(tmp_path / name)
        host, orc = _texels(NativeScene, p), _texels(OracleScene, p)
        assert host[2] != (0, 0) and host == orc, name
    # a grey colour key is compared on the RAW sample before scaling to 8 bits (the png crate's expansion; PIL differs here)
    open(str(tmp_path / "grey4key.png"), "wb").write(_png(3, 2, 0, 4, [b"\x00\x1f\x70", b"\x00\xa5\xc0"], extra=chunkf(b"tRNS", struct.pack(">H", 15))))
    want = np.array([[1, 15, 7], [10, 5, 12]], np.uint8)
    rgba = np.stack([want * 17] * 3 + [np.where(want == 15, 0, 255).astype(np.uint8)], -1)
    h = 1469598103934665603
    for byte in (rgba.astype(np.float32) / np.float32(255.0)).astype(np.float32).tobytes():
        h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert _texels(NativeScene, str(tmp_path / "grey4key.png"))[3] == h
    for name, data, msg in [("i.png", _png(1, 1, 0, 8, [b"\x00\x00"], interlace=1), "interlaced"), ("d16.png", _png(1, 1, 0, 16, [b"\x00\x00\x00"]), "16-bit"),
                            ("sig.png", b"not a png", "bad signature")]:
        open(str(tmp_path / name), "wb").write(data)
        with pytest.raises(RuntimeError, match=msg):
            _texels(NativeScene, str(tmp_path / name))


def test_png_decoder_survives_corruption(tmp_path):
    """Corrupted and truncated PNGs give an error or an image, never a crash or a runaway allocation (the decoder bounds its output by
    the size the header announces)."""
    import io
    import random

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (40, 33, 4), dtype=np.uint8)
    img[:20] = 7
    buf = io.BytesIO()
    Image.fromarray(img, "RGBA").save(buf, "PNG")
    good = buf.getvalue()
    random.seed(1)
    decoded = rejected = 0
    p = str(tmp_path / "fuzz.png")
    for _ in range(400):
        b = bytearray(good)
        for _k in range(random.randint(1, 6)):
            b[random.randrange(len(b))] = random.randrange(256)
        if random.random() < 0.2:
            b = b[:random.randrange(8, len(b))]
        open(p, "wb").write(b)
        try:
            _texels(NativeScene, p)
            decoded += 1
        except RuntimeError:
            rejected += 1
    assert decoded + rejected == 400 and rejected > 100


def test_corrupted_obj_mtl_and_png_never_crash(tmp_path):
    """Random damage to the .obj, its .mtl and its texture: the host loader returns a mesh or an error (ObjLoader::load returns
    Result, obj_loader.rs:332), never a crash; tools/sanitize_host.sh runs this under ASan + UBSan."""
    d = str(tmp_path)
    path = _write_scene(d)
    rng = np.random.default_rng(21)
    files = [path, os.path.join(d, "scene materials.mtl"), os.path.join(d, "tex.png")]
    good = [open(f, "rb").read() for f in files]
    loaded = failed = 0
    for trial in range(240 * int(os.environ.get("MRT_FUZZ_SCALE", "1"))):
        which = trial % 3
        b = bytearray(good[which])
        kind = rng.integers(0, 4)
        if kind == 0:
            for i in rng.integers(0, len(b), rng.integers(1, 6)):
                b[i] = int(rng.integers(0, 256))
        elif kind == 1:
            b = b[:int(rng.integers(0, len(b)))]
        elif kind == 2:
            i = int(rng.integers(0, len(b))); j = min(len(b), i + int(rng.integers(1, 30)))
            b = b[:i] + b[j:]
        else:  # swap two lines (text) / two 8-byte runs (png)
            i, j = sorted(int(x) for x in rng.integers(0, max(len(b) - 8, 1), 2))
            b[i:i + 8], b[j:j + 8] = b[j:j + 8], b[i:i + 8]
        for k, f in enumerate(files):
            open(f, "wb").write(bytes(b) if k == which else good[k])
        for builder in (SimpleTexturedBuilder(WRAP_REPEAT), ObjFns(Lambertian(SolidColor((1, 0, 0, 1))))):
            try:
                v, nrm, uv, mats = _load(NativeScene, path, builder)
                assert len(v) >= 1 and len(v) == len(nrm) == len(uv) == len(mats)
                loaded += 1
            except RuntimeError as e:
                assert str(e)
                failed += 1
    assert loaded > 20 and failed > 20

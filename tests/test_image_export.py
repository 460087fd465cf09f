"""Image export, the host side of Image::to_rgb_bytes / Image::dump (reference main.rs:640-783): the two FloatBuffer display modes
against the oracle's restatement and the formula written out in numpy, and the PNG writer against an independent decoder (PIL) and
against the library's own decoder."""
import ctypes as C
import io
import os
import zlib

import numpy as np
import pytest
from PIL import Image

from mass_raytrace_b200 import DISPLAY_ALBEDO, DISPLAY_DEPTH, DISPLAY_NORMAL, float_buffer_rgb8, write_png
from mass_raytrace_b200 import _ffi


def _float_cases():
    rs = np.random.RandomState(11)
    a = rs.uniform(-0.5, 1.5, (37, 53, 3)).astype(np.float32)
    a[0, 0] = [np.nan, np.inf, -np.inf]
    a[0, 1] = [0.0, 1.0, -0.0]
    a[0, 2] = [1e-30, 0.99999994, 1.0000001]
    a[1, :, :] = np.linspace(-1.0, 1.0, 53 * 3, dtype=np.float32).reshape(53, 3)
    return a


def _saturating_u8(v):
    v = np.where(np.isnan(v), 0.0, v)
    return np.clip(np.floor(np.clip(v, 0.0, 255.0)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("flip", [False, True])
def test_albedo_and_normal_modes(oracle, flip):
    a = _float_cases()
    h, w = a.shape[:2]
    for mode in (DISPLAY_ALBEDO, DISPLAY_NORMAL):
        got = float_buffer_rgb8(a, mode, flip=flip)
        want = np.zeros_like(got)
        oracle.orc_float_buffer_rgb8(a.ctypes.data_as(_ffi.f32p), w, h, mode, int(flip), want.ctypes.data_as(_ffi.u8p))
        assert np.array_equal(got, want)
        # the formula itself (main.rs:694-697, 708-711, 718-721); f32::min(NaN, 1.0) = 1.0, so a NaN albedo is white, a NaN normal is 0
        with np.errstate(invalid="ignore", over="ignore"):
            if mode == DISPLAY_ALBEDO:
                p = np.where(np.isnan(a), np.float32(1.0), np.clip(a, np.float32(0.0), np.float32(1.0)))
                f = np.power(p, np.float32(1.0) / np.float32(2.2), dtype=np.float32)
            else:
                f = (a + np.float32(1.0)) / np.float32(2.0)
            ref = _saturating_u8(f * np.float32(255.0))
        ref = ref[::-1] if flip else ref
        assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1  # numpy's powf may round differently in the last place
        assert (got == ref).mean() > 0.999
    assert float_buffer_rgb8(np.full((2, 2, 3), np.nan, np.float32), DISPLAY_ALBEDO).min() == 255
    assert float_buffer_rgb8(np.full((2, 2, 3), np.nan, np.float32), DISPLAY_NORMAL).max() == 0
    with pytest.raises(ValueError):
        float_buffer_rgb8(a, DISPLAY_DEPTH)  # Depth is resolved from the device image (mrt_resolve_rgb8)


def _images():
    rs = np.random.RandomState(5)
    yy, xx = np.mgrid[0:96, 0:131]
    smooth = np.stack([(xx * 2) % 256, (yy * 3) % 256, ((xx + yy) // 2) % 256], -1).astype(np.uint8)
    flat = np.full((40, 64, 3), 200, np.uint8)
    flat[10:30, 20:50] = [10, 20, 30]
    return {
        "one_pixel": np.array([[[1, 2, 3]]], np.uint8),
        "one_row": rs.randint(0, 256, (1, 300, 3)).astype(np.uint8),
        "one_column": rs.randint(0, 256, (300, 1, 3)).astype(np.uint8),
        "noise": rs.randint(0, 256, (64, 67, 3)).astype(np.uint8),
        "smooth": smooth,
        "flat": flat,
        "long_runs": np.zeros((300, 400, 3), np.uint8),  # matches of the maximum length, distances of one
        "far_matches": np.tile(rs.randint(0, 256, (1, 5000, 3)).astype(np.uint8), (3, 1, 1)),  # distance 15001: inside the window, long chains
        "beyond_window": np.tile(rs.randint(0, 256, (1, 11000, 3)).astype(np.uint8), (3, 1, 1)),  # distance 33001: just outside it
    }


@pytest.mark.parametrize("name", sorted(_images()))
def test_png_round_trip(tmp_path, name):
    img = _images()[name]
    path = tmp_path / "a" / "b" / f"{name}.png"  # create_dir_all(path.parent()) main.rs:771
    write_png(str(path), img)
    data = path.read_bytes()
    back = Image.open(io.BytesIO(data))
    back.load()  # PIL checks every chunk CRC and the Adler-32 of the stream
    assert back.mode == "RGB" and back.size == (img.shape[1], img.shape[0])
    assert np.array_equal(np.asarray(back), img)
    # structure: signature, IHDR (8-bit RGB, no interlace), one zlib stream that zlib itself inflates to (1 + 3w) * h bytes
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and data[12:16] == b"IHDR" and data[24:29] == bytes([8, 2, 0, 0, 0])
    idat = b""
    at = 8
    while at < len(data):
        n = int.from_bytes(data[at:at + 4], "big")
        tag, body = data[at + 4:at + 8], data[at + 8:at + 8 + n]
        assert zlib.crc32(tag + body) == int.from_bytes(data[at + 8 + n:at + 12 + n], "big")
        if tag == b"IDAT":
            idat += body
        at += 12 + n
    assert tag == b"IEND" and len(zlib.decompress(idat)) == (1 + 3 * img.shape[1]) * img.shape[0]
    if name in ("flat", "long_runs", "smooth"):
        assert len(data) < img.size // 4  # the filters and the matcher do their job
    if name == "far_matches":
        assert len(data) < 0.4 * img.size  # one incompressible row, then two that repeat it
    # and the library's own decoder (Texture::load_png path) reads its own files
    lib = _ffi.host_lib()
    s = lib.mrth_scene_new()
    try:
        surf = lib.mrth_surface_texture_png(s, os.fsencode(str(path)), 1)
        assert surf >= 0, lib.mrth_last_error(s).decode()
        mat = lib.mrth_mat_lambertian(s, surf)
        wh = (C.c_uint32 * 2)()
        hsh = C.c_uint64()
        col = (C.c_float * 4)()
        lib.mrth_material_info(s, mat, C.byref(col), C.byref(wh), C.byref(hsh))
        assert (wh[0], wh[1]) == (img.shape[1], img.shape[0])
        rgba = np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], -1)
        texels = (rgba.astype(np.float32) / np.float32(255.0)).tobytes()
        h = 1469598103934665603  # the library's offset basis (mrth_material_info)
        for b in texels if len(texels) < 200000 else b"":
            h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        if len(texels) < 200000:
            assert hsh.value == h
    finally:
        lib.mrth_scene_free(s)


def test_png_writer_errors(tmp_path, capfd):
    img = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(ValueError):
        write_png(str(tmp_path / "x.png"), np.zeros((4, 4), np.uint8))
    blocker = tmp_path / "file"
    blocker.write_text("not a directory")
    with pytest.raises(OSError):
        write_png(str(blocker / "sub" / "x.png"), img)
    assert "Unable to save image" in capfd.readouterr().err  # main.rs:780-782: reported on stderr
    lib = _ffi.host_lib()
    assert lib.mrth_write_png(None, img.ctypes.data_as(_ffi.u8p), 4, 4) == -1
    assert lib.mrth_write_png(os.fsencode(str(tmp_path / "z.png")), img.ctypes.data_as(_ffi.u8p), 0, 4) == -1

"""mrt_scene_validate: the checks mrt_scene_upload makes before it touches the device, run without a GPU.

A Rust `World` (world.rs:95-122) cannot name an object that does not exist; a flattened scene handed over a C ABI can. Every rule of
include/mrt.h's tables is broken once here, on copies of the arrays of real scenes flattened by libmrt_host.so, and a random-mutation
pass checks that no damaged descriptor crashes the validator (it returns MRT_OK or an error with a message)."""
import ctypes as C

import numpy as np
import pytest

from mass_raytrace_b200 import NativeScene, _ffi, scenes
from extra_scenes import eve_scene, mesh_media_scene

NODE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("left", "<u4"), ("right", "<u4")])
SPHERE = np.dtype([("center", "<f4", 3), ("radius", "<f4"), ("material", "<i4"), ("object_id", "<u4"), ("pad", "<u4", 2)])
SHADING = np.dtype([("normal", "<f4", 9), ("uv", "<f4", 6), ("tangent", "<f4", 3), ("bitangent", "<f4", 3), ("material", "<i4"), ("flags", "<u4"), ("pad", "<u4")])
BLAS = np.dtype([("root", "<u4"), ("first_tri", "<u4"), ("n_tris", "<u4"), ("n_nodes", "<u4")])
INSTANCE = np.dtype([("transform", "<f4", 16), ("inv_transform", "<f4", 16), ("bmin", "<f4", 3), ("bmax", "<f4", 3), ("blas", "<u4"), ("material", "<i4"),
                     ("flags", "<u4"), ("object_id", "<u4"), ("pad", "<u4", 2)])
VOLUME = np.dtype([("target", "<u4"), ("neg_inv_density", "<f4"), ("material", "<i4"), ("object_id", "<u4")])
MATERIAL = np.dtype([("kind", "<i4"), ("surface", "<i4"), ("left", "<i4"), ("right", "<i4"), ("p", "<f4", 4)])
SURFACE = np.dtype([("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("mode", "<i4"), ("color", "<f4", 4)])
TEXTURE = np.dtype([("width", "<u4"), ("height", "<u4"), ("wrap", "<i4"), ("pad", "<u4"), ("texel_offset", "<u8")])
TABLES = {"roots": ("n_roots", np.dtype("<u4")), "nodes": ("n_nodes", NODE), "spheres": ("n_spheres", SPHERE), "tri_shading": ("n_tris", SHADING),
          "blas": ("n_blas", BLAS), "instances": ("n_instances", INSTANCE), "volumes": ("n_volumes", VOLUME), "materials": ("n_materials", MATERIAL),
          "surfaces": ("n_surfaces", SURFACE), "textures": ("n_textures", TEXTURE)}
KIND_NODE, KIND_SPHERE, KIND_TRIANGLE, KIND_INSTANCE, KIND_VOLUME = 0, 1, 2, 3, 4
NONE = 0xFFFFFFFF


def test_record_sizes_match_the_header():
    assert [d.itemsize for d in (NODE, SPHERE, SHADING, BLAS, INSTANCE, VOLUME, MATERIAL, SURFACE, TEXTURE)] == [32, 32, 96, 16, 176, 16, 32, 32, 24]


def ref(kind, index):
    return (kind << 29) | index


class Editable:
    """A copy of a flattened scene whose tables are numpy record arrays that can be edited and re-validated."""

    def __init__(self, host, keep=False):
        self.host = host  # keeps the original arrays (tri_verts, texels) alive
        src = host.desc().contents
        self.desc = _ffi.mrt_scene_desc.from_buffer_copy(src)
        self.desc.flags = 1 if keep else 0
        self.t = {}
        for name, (count, dt) in TABLES.items():
            n = getattr(src, count)
            p = getattr(src, name)
            addr = C.cast(p, C.c_void_p).value
            a = np.zeros(max(n, 1), dt)
            if n:
                C.memmove(a.ctypes.data, addr, n * dt.itemsize)
            self.t[name] = a

    def validate(self):
        for name, (count, dt) in TABLES.items():
            a = self.t[name]
            if name == "roots":
                self.desc.roots = a.ctypes.data_as(C.POINTER(C.c_uint32))
            else:
                setattr(self.desc, name, a.ctypes.data)
        why = C.create_string_buffer(256)
        rc = _ffi.cuda_lib().mrt_scene_validate(C.byref(self.desc), why, len(why))
        return rc, why.value.decode(errors="replace")


def book2():
    return NativeScene(*scenes.book2_final(boxes_per_side=4, n_cluster=20))


def cornell():
    return NativeScene(*scenes.cornell_box(1.0))


SCENES = {"cornell": cornell, "book2": book2, "eve": lambda: NativeScene(*eve_scene()), "mesh_media": lambda: NativeScene(*mesh_media_scene(2.0))}


@pytest.mark.parametrize("name", sorted(SCENES))
@pytest.mark.parametrize("keep", [False, True])
def test_real_scenes_validate(name, keep):
    rc, why = Editable(SCENES[name](), keep).validate()
    assert rc == 0 and why == ""


def test_null_and_version():
    lib = _ffi.cuda_lib()
    why = C.create_string_buffer(64)
    assert lib.mrt_scene_validate(None, why, 64) == -1 and b"NULL" in why.value
    assert lib.mrt_scene_validate(None, None, 0) == -1  # no message buffer
    e = Editable(cornell())
    e.desc.abi_version = 99
    assert e.validate() == (-1, "mrt_scene_desc.abi_version mismatch")
    e = Editable(cornell())
    e.validate()
    e.desc.abi_version = _ffi.cuda_lib().mrt_abi_version()
    short = C.create_string_buffer(8)  # a short buffer gets a truncated, terminated message
    e.desc.n_materials = 0
    assert lib.mrt_scene_validate(C.byref(e.desc), short, 8) == -1 and len(short.value) == 7


def _break(scene, keep, edit):
    e = Editable(scene, keep)
    assert e.validate()[0] == 0
    edit(e)
    return e.validate()


def test_every_rule_has_a_failing_case():
    b2, cb = book2(), cornell()
    ev = NativeScene(*eve_scene())
    mm = NativeScene(*mesh_media_scene(2.0))
    n_mat = Editable(b2).desc.n_materials

    def expect(scene, edit, code, text, keep=False):
        rc, why = _break(scene, keep, edit)
        assert rc == code and text in why, (rc, why, text)

    # world list and reference ranges
    expect(b2, lambda e: e.t["roots"].__setitem__(0, ref(KIND_NODE, e.desc.n_nodes)), -1, "root reference out of range")
    expect(b2, lambda e: e.t["roots"].__setitem__(0, ref(KIND_TRIANGLE, 0)), -3, "bare Triangle")
    expect(b2, lambda e: e.t["roots"].__setitem__(0, ref(5, 0)), -1, "root reference out of range")
    expect(b2, lambda e: setattr(e.desc, "n_nodes", 1 << 29), -1, "29-bit")
    # surfaces, textures, materials
    def surf_kind(e): e.t["surfaces"]["kind"][0] = 9
    expect(b2, surf_kind, -1, "malformed surface table entry 0")
    def tex_index(e):
        s = e.t["surfaces"]; i = int(np.flatnonzero(s["kind"] == 1)[0]); s["a"][i] = e.desc.n_textures
    expect(b2, tex_index, -1, "malformed surface table entry")
    def tex_range(e): e.t["textures"]["texel_offset"][0] = e.desc.n_texels
    expect(b2, tex_range, -1, "texture outside the texel array")
    def tex_wrap_around(e): e.t["textures"]["texel_offset"][0] = (1 << 64) - 1  # offset + w * h wraps to a small number
    expect(b2, tex_wrap_around, -1, "texture outside the texel array")
    def tex_zero(e): e.t["textures"]["width"][0] = 0
    expect(b2, tex_zero, -1, "texture outside the texel array")
    def tex_mirror(e): e.t["textures"]["wrap"][0] = 0
    expect(b2, tex_mirror, -3, "Mirror")
    def mat_kind(e): e.t["materials"]["kind"][0] = 9
    expect(b2, mat_kind, -1, "malformed material table entry 0")
    def mat_surface(e):
        m = e.t["materials"]; i = int(np.flatnonzero(m["kind"] == 1)[0]); m["surface"][i] = -1
    expect(b2, mat_surface, -1, "malformed material table entry")
    def eve_palette(e):
        m = e.t["materials"]; i = int(np.flatnonzero(m["kind"] == 8)[0]); m["p"][i, 0] = np.array([e.desc.n_surfaces - 2], "<i4").view("<f4")[0]
    expect(ev, eve_palette, -1, "malformed material table entry")
    def eve_texture(e):
        m = e.t["materials"]; i = int(np.flatnonzero(m["kind"] == 8)[0]); m["left"][i] = e.desc.n_surfaces
    expect(ev, eve_texture, -1, "malformed material table entry")
    # primitives
    def sphere_mat(e): e.t["spheres"]["material"][0] = n_mat
    expect(b2, sphere_mat, -1, "sphere material out of range")
    def tri_mat(e): e.t["tri_shading"]["material"][-1] = -1
    expect(b2, tri_mat, -1, "triangle material out of range")
    def vol_target(e): e.t["volumes"]["target"][0] = ref(KIND_SPHERE, e.desc.n_spheres)
    expect(b2, vol_target, -1, "volume target out of range")
    def vol_kind(e): e.t["volumes"]["target"][0] = ref(KIND_TRIANGLE, 0)
    expect(b2, vol_kind, -3, "Volume targets")
    def vol_mat(e): e.t["volumes"]["material"][0] = -2
    expect(b2, vol_mat, -1, "volume material out of range")
    def vol_instance(e): e.t["volumes"]["target"][0] = ref(KIND_INSTANCE, e.desc.n_instances)
    expect(mm, vol_instance, -1, "volume target out of range")
    def bg_surface(e): e.desc.background.kind = 2; e.desc.background.surface[0] = e.desc.n_surfaces
    expect(b2, bg_surface, -1, "background surface out of range")
    def bg_cube(e): e.desc.background.kind = 3; e.desc.background.surface[5] = -1
    expect(b2, bg_cube, -1, "background surface out of range")
    def bg_kind(e): e.desc.background.kind = 4
    expect(b2, bg_kind, -1, "unknown background kind")
    # meshes and instances
    def blas_range(e): e.t["blas"]["n_tris"][0] = e.desc.n_tris + 1
    expect(cb, blas_range, -1, "malformed BLAS table entry")
    def blas_empty(e): e.t["blas"]["n_tris"][0] = 0
    expect(cb, blas_empty, -1, "BLAS without triangles")
    def blas_root(e): e.t["blas"]["root"][0] = ref(KIND_SPHERE, 0)
    expect(cb, blas_root, -1, "malformed BLAS table entry")
    def blas_no_tree(e): e.t["blas"]["root"][0] = NONE
    expect(cb, blas_no_tree, -1, "needs the nodes of every BLAS", keep=True)
    assert _break(cb, False, blas_no_tree)[0] == 0  # the default upload builds its own tree
    def inst_blas(e): e.t["instances"]["blas"][0] = e.desc.n_blas
    expect(cb, inst_blas, -1, "instance BLAS out of range")
    def inst_mat(e): e.t["instances"]["material"][0] = -2
    expect(cb, inst_mat, -1, "instance material out of range")
    # trees (only the caller's topology is read under MRT_SCENE_KEEP_TOPOLOGY; the world's tree is always walked)
    def blas_node_sphere(e):
        root = int(e.t["blas"]["root"][0]) & 0x1FFFFFFF; e.t["nodes"]["left"][root] = ref(KIND_SPHERE, 0)
    expect(cb, blas_node_sphere, -1, "a BLAS may only contain triangles", keep=True)
    assert _break(cb, False, blas_node_sphere)[0] == 0
    def tlas_triangle(e):
        root = int(e.t["roots"][0]) & 0x1FFFFFFF; e.t["nodes"]["left"][root] = ref(KIND_TRIANGLE, 0)
    expect(cb, tlas_triangle, -3, "bare Triangle")
    def tlas_cycle(e):
        root = int(e.t["roots"][0]) & 0x1FFFFFFF; e.t["nodes"]["left"][root] = ref(KIND_NODE, root); e.t["nodes"]["right"][root] = ref(KIND_NODE, root)
    expect(cb, tlas_cycle, -1, "not a tree")
    def tlas_no_left(e):
        root = int(e.t["roots"][0]) & 0x1FFFFFFF; e.t["nodes"]["left"][root] = NONE
    expect(cb, tlas_no_left, -1, "without a left child")
    def tlas_child_range(e):
        root = int(e.t["roots"][0]) & 0x1FFFFFFF; e.t["nodes"]["right"][root] = ref(KIND_INSTANCE, e.desc.n_instances)
    expect(cb, tlas_child_range, -1, "child reference out of range")
    def unreachable_node(e):  # appended node nobody points to, with a child that does not exist: converted in keep mode, so it must be checked
        e.t["nodes"] = np.concatenate([e.t["nodes"], np.zeros(1, NODE)]); e.t["nodes"]["left"][-1] = ref(KIND_NODE, 1 << 28); e.t["nodes"]["right"][-1] = NONE
        e.desc.n_nodes += 1
    expect(cb, unreachable_node, -1, "child reference out of range", keep=True)
    assert _break(cb, False, unreachable_node)[0] == 0
    # null arrays
    e = Editable(cb)
    e.validate()
    e.desc.materials = None
    why = C.create_string_buffer(128)
    assert _ffi.cuda_lib().mrt_scene_validate(C.byref(e.desc), why, 128) == -1 and b"NULL while its count" in why.value


def test_deep_caller_tree_is_refused_in_keep_mode():
    """A left-leaning chain of 100 nodes over spheres: fine for the default rebuild, too deep for the 96-entry traversal stack as given."""
    e = Editable(NativeScene(*scenes.book1_spheres(1.5, aperture=0.1)), keep=True)
    n_sph = int(e.desc.n_spheres)
    assert n_sph > 101
    depth = 100
    nodes = np.zeros(depth, NODE)
    for i in range(depth):
        nodes["left"][i] = ref(KIND_NODE, i + 1) if i + 1 < depth else ref(KIND_SPHERE, depth)
        nodes["right"][i] = ref(KIND_SPHERE, i)
    e.t["nodes"] = nodes
    e.desc.n_nodes = depth
    e.t["roots"] = np.array([ref(KIND_NODE, 0)], "<u4")
    e.desc.n_roots = 1
    rc, why = e.validate()
    assert rc == -3 and "too deep" in why
    e.desc.flags = 0
    assert e.validate()[0] == 0


@pytest.mark.parametrize("name", sorted(SCENES))
def test_random_damage_never_crashes_the_validator(name):
    rng = np.random.default_rng(5)
    host = SCENES[name]()
    codes = {0: 0, -1: 0, -3: 0}
    for trial in range(400):
        e = Editable(host, keep=bool(trial & 1))
        for _ in range(int(rng.integers(1, 4))):
            table = list(TABLES)[int(rng.integers(0, len(TABLES)))]
            a = e.t[table]
            n = getattr(e.desc, TABLES[table][0])
            if n == 0:
                continue
            raw = a.view(np.uint8).reshape(len(a), -1)
            row = int(rng.integers(0, n))
            word = int(rng.integers(0, raw.shape[1] // 4))
            val = [0, 1, NONE, 0x7FFFFFFF, 0x80000000, int(rng.integers(0, 1 << 32)), ref(int(rng.integers(0, 8)), int(rng.integers(0, 64)))][int(rng.integers(0, 7))]
            raw[row, 4 * word:4 * word + 4] = np.frombuffer(np.uint32(val).tobytes(), np.uint8)
        rc, why = e.validate()
        assert rc in codes and (rc == 0) == (why == ""), (rc, why)
        codes[rc] += 1
    assert codes[0] > 0 and codes[-1] > 0  # float fields and padding absorb some damage; index fields do not

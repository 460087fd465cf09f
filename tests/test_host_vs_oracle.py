"""The product's host-side builders (libmrt_host.so: Triangle::new normals, Instance::new matrices and bounds, Camera::new,
BvhNode::new, flatten) against the oracle's independent restatement: results must be bit-identical."""
import ctypes as C

import numpy as np
import pytest

from mass_raytrace_b200 import NativeScene, scenes
from mass_raytrace_b200 import _ffi
from oracle_backend import OracleScene


def _scene_list(tmp_mesh_dir):
    ply = str(tmp_mesh_dir / "mesh_32x16.ply")
    n, md = scenes.write_synthetic_ply(ply, 32, 16, seed=3)
    return {
        "cornell": scenes.cornell_box(1.0),
        "book1": scenes.book1_spheres(),
        "sphere_grid": scenes.sphere_grid(dim=6),
        "lucy": scenes.lucy_layout(ply, md, grid=1),
        "book2": scenes.book2_final(boxes_per_side=6, n_cluster=40),
    }


@pytest.fixture(scope="module")
def scene_pairs(tmp_mesh_dir):
    out = {}
    for name, (world, camera) in _scene_list(tmp_mesh_dir).items():
        out[name] = (world, camera, NativeScene(world, camera), OracleScene(world, camera))
    return out


@pytest.mark.parametrize("name", ["cornell", "book1", "sphere_grid", "lucy", "book2"])
def test_builders_bit_identical(scene_pairs, name):
    world, camera, host, orc = scene_pairs[name]
    assert np.array_equal(host.camera_fields(), orc.camera_fields())
    assert host.tlas_node_count() == orc.tlas_node_count()
    from mass_raytrace_b200.api import Instance, Model
    seen = set()
    for i, obj in enumerate(world.objects):
        assert np.array_equal(host.object_aabb(i), orc.object_aabb(i)), f"object {i} bounds"
        if isinstance(obj, Instance):
            for a, b in zip(host.instance_fields(i), orc.instance_fields(i)):
                assert np.array_equal(a, b), f"instance {i}"
        tris = obj.model.triangles if isinstance(obj, Instance) else (obj.triangles if isinstance(obj, Model) else None)
        if tris is not None and id(tris) not in seen:
            seen.add(id(tris))
            assert np.array_equal(host.mesh_verts(tris), orc.mesh_verts(tris))
            assert host.mesh_node_count(tris) == orc.mesh_node_count(tris)


def _bounds(d, ref):
    kind, idx = ref >> 29, ref & 0x1FFFFFFF
    if kind == 0:
        n = d["nodes"][idx]
        return n["bmin"], n["bmax"]
    if kind == 1:
        s = d["spheres"][idx]
        return s["center"] - abs(s["radius"]), s["center"] + abs(s["radius"])
    if kind == 2:
        v = d["tri_verts"][idx].reshape(3, 3)
        return v.min(0), v.max(0)
    if kind == 3:
        return d["instances"][idx]["bmin"], d["instances"][idx]["bmax"]
    return _bounds(d, int(d["volumes"][idx]["target"]))


NODE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("left", "<u4"), ("right", "<u4")])
SPHERE = np.dtype([("center", "<f4", 3), ("radius", "<f4"), ("material", "<i4"), ("object_id", "<u4"), ("pad", "<u4", 2)])
INSTANCE = np.dtype([("transform", "<f4", 16), ("inv", "<f4", 16), ("bmin", "<f4", 3), ("bmax", "<f4", 3), ("blas", "<u4"), ("material", "<i4"),
                     ("flags", "<u4"), ("object_id", "<u4"), ("pad", "<u4", 2)])
VOLUME = np.dtype([("target", "<u4"), ("neg_inv_density", "<f4"), ("material", "<i4"), ("object_id", "<u4")])
BLAS = np.dtype([("root", "<u4"), ("first_tri", "<u4"), ("n_tris", "<u4"), ("n_nodes", "<u4")])
SHADING = np.dtype([("normal", "<f4", 9), ("uv", "<f4", 6), ("tangent", "<f4", 3), ("bitangent", "<f4", 3), ("material", "<i4"), ("flags", "<u4"), ("pad", "<u4")])


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value)
    return np.frombuffer(buf, dtype=dtype, count=n)


def desc_arrays(host):
    d = host.desc().contents
    return dict(
        d=d,
        nodes=_view(d.nodes, d.n_nodes, NODE), spheres=_view(d.spheres, d.n_spheres, SPHERE), instances=_view(d.instances, d.n_instances, INSTANCE),
        volumes=_view(d.volumes, d.n_volumes, VOLUME), blas=_view(d.blas, d.n_blas, BLAS), shading=_view(d.tri_shading, d.n_tris, SHADING),
        tri_verts=_view(d.tri_verts, d.n_tris * 9, np.dtype("<f4")).reshape(-1, 9) if d.n_tris else np.zeros((0, 9), np.float32),
        roots=[d.roots[i] for i in range(d.n_roots)],
    )


def test_struct_sizes_match_header():
    # include/mrt.h states the sizes; the flattened arrays are read with these dtypes
    assert (NODE.itemsize, SPHERE.itemsize, INSTANCE.itemsize, VOLUME.itemsize, BLAS.itemsize, SHADING.itemsize) == (32, 32, 176, 16, 16, 96)


@pytest.mark.parametrize("name", ["cornell", "book1", "lucy", "book2"])
def test_flattened_scene_is_a_consistent_tree(scene_pairs, name):
    world, camera, host, orc = scene_pairs[name]
    d = desc_arrays(host)
    assert d["d"].abi_version == 1 and len(d["roots"]) == 1 and d["d"].n_objects == len(world.objects)
    nodes = d["nodes"]
    seen_nodes, seen_prims = set(), set()

    def walk(ref, tlas):
        kind, idx = ref >> 29, ref & 0x1FFFFFFF
        if kind != 0:
            assert (kind == 2) == (not tlas)  # triangles only inside a BLAS, everything else only in the TLAS
            assert ref not in seen_prims
            seen_prims.add(ref)
            return
        assert idx not in seen_nodes
        seen_nodes.add(idx)
        n = nodes[idx]
        lo, hi = _bounds(d, int(n["left"]))
        if n["right"] != 0xFFFFFFFF:
            lo2, hi2 = _bounds(d, int(n["right"]))
            lo, hi = np.minimum(lo, lo2), np.maximum(hi, hi2)
        assert np.array_equal(n["bmin"], lo.astype(np.float32)) and np.array_equal(n["bmax"], hi.astype(np.float32))  # BoundingBox::join geom.rs:249
        walk(int(n["left"]), tlas)
        if n["right"] != 0xFFFFFFFF:
            walk(int(n["right"]), tlas)

    walk(d["roots"][0], True)
    for b in d["blas"]:
        before = len(seen_nodes)
        walk(int(b["root"]), False)
        assert len(seen_nodes) - before == b["n_nodes"]
    assert len(seen_nodes) == len(nodes)
    n_tlas_prims = len(d["spheres"]) - len(d["volumes"]) + len(d["instances"]) + len(d["volumes"])
    assert len([p for p in seen_prims if (p >> 29) != 2]) == n_tlas_prims == len(world.objects)
    assert len([p for p in seen_prims if (p >> 29) == 2]) == len(d["tri_verts"])
    # object ids cover World::add order exactly once
    ids = sorted([int(s["object_id"]) for s in d["spheres"] if s["object_id"] != 0xFFFFFFFF] + [int(i["object_id"]) for i in d["instances"]] +
                 [int(v["object_id"]) for v in d["volumes"]])
    assert ids == list(range(len(world.objects)))


def test_flat_triangle_normals_are_unit_cross(scene_pairs):
    world, camera, host, orc = scene_pairs["cornell"]
    d = desc_arrays(host)
    v = d["tri_verts"].reshape(-1, 3, 3)
    n = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    for k in range(3):
        np.testing.assert_allclose(d["shading"]["normal"].reshape(-1, 3, 3)[:, k], n, atol=1e-6)
    assert (d["shading"]["material"] >= 0).all() and (d["shading"]["flags"] == 0).all()


def test_material_override_chain(scene_pairs):
    # N2: loader triangles carry `()`, instances override (scenes/cornell.rs:38-50); Model::instance drops the model's material
    world, camera, host, orc = scene_pairs["cornell"]
    d = desc_arrays(host)
    mats = _view(d["d"].materials, d["d"].n_materials, np.dtype([("kind", "<i4"), ("surface", "<i4"), ("left", "<i4"), ("right", "<i4"), ("p", "<f4", 4)]))
    assert mats[d["shading"]["material"][0]]["kind"] == 0  # ABSORB
    kinds = [int(mats[i["material"]]["kind"]) for i in d["instances"]]
    assert kinds == [1, 1, 1, 1, 1, 2, 1]  # five Lambertian walls, the DiffuseLight, the white block
    assert (d["instances"]["flags"] == 0).all()


def test_uv_mesh_and_textures_flatten(scene_pairs):
    world, camera, host, orc = scene_pairs["book2"]
    d = desc_arrays(host)
    uv = d["shading"][d["shading"]["flags"] == 1]
    assert len(uv) > 1000 and (np.abs(uv["tangent"]).sum(1) > 0).any()
    assert d["d"].n_textures == 1 and d["d"].n_texels == 256 * 128
    tex = np.ctypeslib.as_array(d["d"].texels, shape=(d["d"].n_texels * 4,))
    assert 0.0 <= tex.min() and tex.max() <= 1.0 and tex[3] == 1.0
    assert len(d["volumes"]) == 2 and np.isclose(d["volumes"]["neg_inv_density"][0], -5.0)
    assert any(i["flags"] == 1 for i in d["instances"])  # the Model (identity instance) holding the textured sphere


@pytest.mark.parametrize("name", ["cornell", "lucy", "book2"])
def test_deferred_mesh_bvh_keeps_everything_but_the_mesh_nodes(tmp_mesh_dir, name):
    """mrth_defer_mesh_bvh (include/mrt_host.h): no reference-topology tree per mesh. Triangles, shading records, instance matrices
    and instance boxes (Model::bounding_box / Instance::new, geom.rs:330-332, 369-381, from the mesh box) must not change; only the
    BLAS nodes go, and the TLAS is still a consistent tree over every object."""
    world, camera = _scene_list(tmp_mesh_dir)[name]
    ref, lazy = NativeScene(world, camera), NativeScene(world, camera, defer_mesh_bvh=True)
    a, b = desc_arrays(ref), desc_arrays(lazy)
    assert np.array_equal(a["tri_verts"], b["tri_verts"]) and a["shading"].tobytes() == b["shading"].tobytes()
    assert a["spheres"].tobytes() == b["spheres"].tobytes() and a["volumes"].tobytes() == b["volumes"].tobytes()
    assert a["instances"].tobytes() == b["instances"].tobytes()  # transforms, world boxes, BLAS index, material, object id
    assert len(b["blas"]) == len(a["blas"]) > 0
    for x, y in zip(a["blas"], b["blas"]):
        assert (int(y["first_tri"]), int(y["n_tris"])) == (int(x["first_tri"]), int(x["n_tris"]))
        assert int(y["root"]) == 0xFFFFFFFF and int(y["n_nodes"]) == 0
    n_blas_nodes = int(sum(int(x["n_nodes"]) for x in a["blas"]))
    assert len(b["nodes"]) == len(a["nodes"]) - n_blas_nodes
    for m in range(len(b["blas"])):
        assert lazy._fn("mesh_node_count")(lazy._h, m) == 0
    # the TLAS: every object exactly once, boxes joined bottom-up
    seen = []

    def walk(r):
        kind, idx = r >> 29, r & 0x1FFFFFFF
        if kind != 0:
            seen.append(r)
            return _bounds(b, r)
        n = b["nodes"][idx]
        lo, hi = walk(int(n["left"]))
        if n["right"] != 0xFFFFFFFF:
            lo2, hi2 = walk(int(n["right"]))
            lo, hi = np.minimum(lo, lo2), np.maximum(hi, hi2)
        assert np.array_equal(n["bmin"], lo.astype(np.float32)) and np.array_equal(n["bmax"], hi.astype(np.float32))
        return lo, hi

    assert len(b["roots"]) == 1
    walk(b["roots"][0])
    assert len(seen) == len(set(seen)) == len(world.objects)

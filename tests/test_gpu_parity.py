"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the CPU oracle.

North-star test 1: deterministic primary rays (pixel centres, aperture 0) -- primitive ids exact except exact-t ties, t and normal
within 1e-5 relative (the implementation is in fact bit-identical; the tests record how many pixels are).
North-star test 2: converged renders at matched spp -- RMSE(gpu, oracle) <= 1.15 * RMSE(oracle seed A, oracle seed B), mean
luminance and mean bounce count within max(0.5 %, 3 sigma)."""
import os

import numpy as np
import pytest

from mass_raytrace_b200 import (BLEND_ADDITION, WRAP_CLAMP, WRAP_REPEAT, EveMaterial, Camera, CubeMap, Dielectric, DiffuseLight, Lambertian, Metal, Mix, Model, NativeScene,
                                SkyBackground, SkySphere, SolidBackground, SolidColor, SolidColorFallback, Specular, Sphere, Texture, TextureBlend, V3, V3_fill,
                                Volume, World, YCbCrTexture, scenes)
from mass_raytrace_b200.api import Volume as VolumeT
from oracle_backend import OracleScene
from extra_scenes import eve_scene, mesh_media_scene

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
Y = np.array([0.2126, 0.7152, 0.0722], np.float32)
NONE = 0xFFFFFFFF


def check_aov(g, o, exclude=None, albedo_exact=True, max_ties=None):
    """ids exact except exact-t ties; t, normal within 1e-5 relative (and almost everywhere bit-identical).
    An exact tie = a pixel centre exactly on an edge shared by two primitives (e.g. the diagonal of a cube face): both candidates
    have the same t and the winner depends on the order the tree is walked in, which the north star excludes from the comparison.
    max_ties=None allows up to 0.5 % of the pixels to be such ties (the SAH rebuild walks a different tree than the oracle)."""
    if max_ties is None:
        max_ties = int(0.005 * g["object"].size)
    keep = np.ones(g["object"].shape, bool) if exclude is None else ~exclude
    ids_differ = ((g["object"] != o["object"]) | (g["tri"] != o["tri"])) & keep
    both_hit = np.isfinite(g["t"]) & np.isfinite(o["t"])
    assert np.array_equal(np.isfinite(g["t"])[keep], np.isfinite(o["t"])[keep]), "hit/miss pattern differs"
    rel = np.zeros(g["t"].shape, np.float64)
    rel[both_hit] = np.abs(g["t"][both_hit].astype(np.float64) - o["t"][both_hit]) / np.abs(o["t"][both_hit])
    assert rel[keep].max() <= 1e-5, f"t differs by {rel[keep].max():.3g} relative"
    # a primitive-id mismatch is allowed only on an exact tie: same t (<= 1e-6 relative), two candidate primitives
    assert ids_differ.sum() <= max_ties, f"{ids_differ.sum()} id mismatches"
    assert (rel[ids_differ] <= 1e-6).all(), "id mismatch that is not a tie"
    assert (g["t"][ids_differ] == o["t"][ids_differ]).mean() >= 0.99 if ids_differ.any() else True
    same = keep & ~ids_differ
    nerr = np.abs(g["normal"].astype(np.float64) - o["normal"]).max(-1)
    assert nerr[same].max() <= 1e-5, f"normal differs by {nerr[same].max():.3g}"
    bit_exact = (g["t"] == o["t"]) | (np.isinf(g["t"]) & np.isinf(o["t"]))
    bit_exact &= (g["normal"] == o["normal"]).all(-1)
    frac = bit_exact[same].mean()
    assert frac >= 0.9999, f"only {frac:.6f} of the pixels are bit-identical"
    if albedo_exact:
        assert np.array_equal(g["albedo"][same], o["albedo"][same]), "albedo differs"
    return int(ids_differ.sum()), float(frac)


def volume_mask(world, *aovs):
    vol_ids = [i for i, ob in enumerate(world.objects) if isinstance(ob, VolumeT)]
    m = np.zeros(aovs[0]["object"].shape, bool)
    for a in aovs:
        m |= np.isin(a["object"], vol_ids)
    return m


@pytest.fixture(scope="module")
def mesh_ply(tmp_mesh_dir):
    path = str(tmp_mesh_dir / "mesh_256x128.ply")
    n, md = scenes.write_synthetic_ply(path, 256, 128, seed=1)
    return path, md, n


# ------------------------------------------------------------------------------------------------------------------
# test 1: primary rays
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("keep_topology", [False, True])
def test_aov_cornell(renderer, keep_topology):
    world, camera = scenes.cornell_box(1.0)
    renderer.set_scene(NativeScene(world, camera), keep_topology=keep_topology)
    g = renderer.render_aov(512, 512)
    o = OracleScene(world, camera).render_aov(512, 512)
    ties, frac = check_aov(g, o, max_ties=130 if keep_topology else None)  # 0.05 %: pixel centres exactly on cube-face diagonals
    assert frac == 1.0


def test_aov_cornell_golden(renderer):
    gold = np.load(os.path.join(HERE, "golden", "cornell.npz"))
    world, camera = scenes.cornell_box(1.0)
    renderer.set_scene(NativeScene(world, camera))
    h, w = gold["aov_object"].shape
    g = renderer.render_aov(w, h)
    o = dict(object=gold["aov_object"], tri=gold["aov_tri"], t=gold["aov_t"], normal=gold["aov_normal"], albedo=gold["aov_albedo"])
    check_aov(g, o, max_ties=int(0.02 * w * h))  # a 64x64 symmetric view puts many pixel centres exactly on cube-face diagonals


@pytest.mark.parametrize("keep_topology", [False, True])
def test_aov_book1(renderer, keep_topology):
    world, camera = scenes.book1_spheres(1.5, aperture=0.0)
    renderer.set_scene(NativeScene(world, camera), keep_topology=keep_topology)
    g = renderer.render_aov(600, 400)
    o = OracleScene(world, camera).render_aov(600, 400)
    ties, frac = check_aov(g, o, albedo_exact=False)
    assert ties == 0 and frac == 1.0
    # albedo: identical except on fuzzy Metal, where albedo_normal's scatter() consumes randoms (world.rs:84, material.rs:268-279)
    assert (g["albedo"] != o["albedo"]).any(-1).mean() < 0.005
    gold = np.load(os.path.join(HERE, "golden", "book1.npz"))
    h, w = gold["aov_object"].shape
    g2 = renderer.render_aov(w, h)
    check_aov(g2, dict(object=gold["aov_object"], tri=gold["aov_tri"], t=gold["aov_t"], normal=gold["aov_normal"], albedo=gold["aov_albedo"]), albedo_exact=False)


def test_aov_sphere_grid(renderer):
    world, camera = scenes.sphere_grid(dim=20)
    renderer.set_scene(NativeScene(world, camera))
    g = renderer.render_aov(480, 270)
    o = OracleScene(world, camera).render_aov(480, 270)
    check_aov(g, o)


@pytest.mark.parametrize("keep_topology", [False, True])
def test_aov_instanced_mesh_field(renderer, mesh_ply, keep_topology):
    path, md, n = mesh_ply
    world, camera = scenes.lucy_layout(path, md, grid=2)  # 25 rotated, scaled instances of one 65k-triangle BLAS
    renderer.set_scene(NativeScene(world, camera), keep_topology=keep_topology)
    g = renderer.render_aov(640, 360)
    o = OracleScene(world, camera).render_aov(640, 360)
    check_aov(g, o)
    assert (g["tri"] != NONE).mean() > 0.3


@pytest.mark.parametrize("keep_topology", [False, True])
def test_aov_book2_with_uv_mesh_texture_and_volumes(renderer, keep_topology):
    world, camera = scenes.book2_final(boxes_per_side=12, n_cluster=200)
    renderer.set_scene(NativeScene(world, camera), keep_topology=keep_topology)
    g = renderer.render_aov(480, 270)
    o = OracleScene(world, camera).render_aov(480, 270)
    # Volume hits depend on the free-flight random (geom.rs:638): excluded wherever either side reports a Volume
    vm = volume_mask(world, g, o)
    assert 0.01 < vm.mean() < 0.6
    check_aov(g, o, exclude=vm, albedo_exact=False)
    # textured UV mesh (object index of the Model): bilinear texel fetch must match bit for bit
    model_id = [i for i, ob in enumerate(world.objects) if isinstance(ob, Model)][0]
    on_mesh = (g["object"] == model_id) & (o["object"] == model_id) & (g["tri"] == o["tri"])
    assert on_mesh.sum() > 500
    assert np.array_equal(g["albedo"][on_mesh], o["albedo"][on_mesh])
    assert len(np.unique(g["albedo"][on_mesh], axis=0)) > 100  # the texture really varies over the mesh
    # volume statistics agree: fraction of primary rays scattered by the two media
    for vid in [i for i, ob in enumerate(world.objects) if isinstance(ob, VolumeT)]:
        fg, fo = (g["object"] == vid).mean(), (o["object"] == vid).mean()
        assert abs(fg - fo) < 0.01 + 0.15 * fo, (vid, fg, fo)


def test_aov_axis_parallel_rays(renderer):
    """Rays with an exactly zero direction component: the centre row and centre column of an odd-sized image whose camera looks along
    an axis. (b - o) / 0 = +-inf in the reference's slab test (geom.rs:219-220); a reciprocal-multiply form must not turn that into a
    NaN that rejects boxes straddling the origin of that axis -- round 1 did, for every box with min < 0 < max on the axis."""
    w = World(SkyBackground())
    w.add(Sphere(Lambertian(SolidColor((0.5, 0.5, 0.5, 1))), V3(0, -1000, 0), 1000.0))
    rs = np.random.RandomState(11)
    for i in range(60):
        c = rs.uniform(-6, 6, 3)
        w.add(Sphere(Lambertian(SolidColor((0.2 + 0.01 * i, 0.3, 0.4, 1))), V3(c[0], abs(c[1]) * 0.5 + 0.2, c[2]), float(rs.uniform(0.2, 1.2))))
    w.add(Sphere(Metal(0.0, SolidColor((0.9, 0.9, 0.9, 1))), V3(0, 1, 0), 1.0))  # straddles x = 0 and z = 0, centred on the view axis
    from mass_raytrace_b200 import PlyLoader
    cube = Model(PlyLoader.load(scenes.CUBE_PLY))
    w.add(cube.instance(V3(0, 1, -4), V3(0, 0, 0), V3(3, 1.5, 1)))           # axis-aligned box straddling x = 0 behind the sphere
    w.add(cube.instance(V3(-2.5, 1, 2), V3(0, 0.5, 0), V3(1, 1, 1)))
    w.build_bvh()
    for keep in (False, True):
        for cam in (Camera(40.0, V3(0, 1, 12), V3(0, 1, 0), V3(0, 1, 0), 1.0, 0.0, 12.0),     # looks along -z at the height of its target: d.y == 0 on the
                    Camera(40.0, V3(12, 1, 0), V3(0, 1, 0), V3(0, 1, 0), 1.0, 0.0, 12.0)):   # centre row, d.x (or d.z) == 0 on the centre column
            renderer.set_scene(NativeScene(w, cam), keep_topology=keep)
            g = renderer.render_aov(255, 255)
            o = OracleScene(w, cam).render_aov(255, 255)
            check_aov(g, o, albedo_exact=False)
            assert np.array_equal(g["t"][127], o["t"][127]) and np.array_equal(g["t"][:, 127], o["t"][:, 127])
            assert np.isfinite(g["t"][127]).mean() > 0.5


def test_no_ray_defeats_the_box_tests(renderer, mesh_ply):
    """The performance face of the axis-parallel fix: a ray with zero direction components (here every primary ray of the centre row
    and column of an axis-aligned camera, and Lambertian bounces that fall back to an axis-aligned normal, material.rs:207) must be
    culled like any other. When such rays ignored an axis instead, one of them visited 250,000 nodes of a 1 M-triangle mesh;
    mrt_stats.max_ray_node_visits (instrumented renders) bounds the worst ray of a render."""
    path, md, n = mesh_ply
    world, _ = scenes.lucy_layout(path, md, grid=1)  # nine instances of a 65k-triangle mesh over a ground cube
    for cam in (Camera(40.0, V3(0, 1, 12), V3(0, 1, 0), V3(0, 1, 0), 1.0, 0.0, 12.0), Camera(40.0, V3(0, 30, 0), V3(0, 0, 0), V3(0, 0, -1), 1.0, 0.0, 30.0)):
        renderer.set_scene(NativeScene(world, cam))
        renderer.set_option(renderer.OPT_COUNT_VISITS, 1)
        try:
            renderer.render(255, 255, 16, 50, seed=21)
            st = renderer.stats()
        finally:
            renderer.set_option(renderer.OPT_COUNT_VISITS, 0)
        assert st["node_visits"] / st["rays"] < 40
        assert 0 < st["max_ray_node_visits"] < 3000, st["max_ray_node_visits"]  # a whole-tree walk would be ~130,000


def test_volume_hit_probability_per_pixel(renderer):
    """Volume::intersect draws its free-flight distance from the path's RNG (geom.rs:638), so a primary ray through a medium
    reports the Volume with probability 1 - exp(-density * chord) and otherwise whatever lies behind. Over 64 seeds the PER-PIXEL hit
    frequency of each medium must agree with the oracle's like two binomial samples of the same probability, and so must the
    distribution of the reported distances."""
    world, camera = scenes.book2_final(boxes_per_side=12, n_cluster=200)
    renderer.set_scene(NativeScene(world, camera))
    orc = OracleScene(world, camera)
    w, h, n_seeds = 240, 135, 64
    vol_ids = [i for i, ob in enumerate(world.objects) if isinstance(ob, VolumeT)]
    assert len(vol_ids) == 2
    fg = {v: np.zeros((h, w)) for v in vol_ids}
    fo = {v: np.zeros((h, w)) for v in vol_ids}
    tg, to = {v: [] for v in vol_ids}, {v: [] for v in vol_ids}
    for seed in range(1, n_seeds + 1):
        g = renderer.render_aov(w, h, seed=seed)
        o = orc.render_aov(w, h, seed=1000 + seed)
        for v in vol_ids:
            fg[v] += g["object"] == v
            fo[v] += o["object"] == v
            tg[v].append(g["t"][g["object"] == v])
            to[v].append(o["t"][o["object"] == v])
        behind = ~np.isin(g["object"], vol_ids) & ~np.isin(o["object"], vol_ids)
        assert np.array_equal(g["t"][behind], o["t"][behind])  # whatever is seen through the medium is the deterministic scene
    checked = 0
    for v in vol_ids:
        a, b = fg[v] / n_seeds, fo[v] / n_seeds
        p = 0.5 * (a + b)
        assert np.array_equal(p > 0, (a > 0) | (b > 0))
        mid = (p > 0.05) & (p < 0.95)  # pixels where both frequencies are well inside (0, 1)
        if mid.sum() < 200:
            # a medium this thin (or this dense) over the view: compare the frequencies summed over the image instead
            na, nb = fg[v].sum(), fo[v].sum()
            assert abs(na - nb) <= 5.0 * np.sqrt(na + nb + 1.0), (v, na, nb)
            continue
        z = (a[mid] - b[mid]) / np.sqrt(2.0 * p[mid] * (1.0 - p[mid]) / n_seeds)
        assert np.abs(z).max() < 6.0, (v, float(np.abs(z).max()))
        assert 0.8 < float(np.mean(z * z)) < 1.25, (v, float(np.mean(z * z)))  # two samples of the same per-pixel probability
        assert abs(float(np.mean(z))) < 5.0 / np.sqrt(mid.sum()), (v, float(np.mean(z)))
        x, y = np.concatenate(tg[v]).astype(np.float64), np.concatenate(to[v]).astype(np.float64)
        se = np.sqrt(x.var() / x.size + y.var() / y.size)
        assert abs(x.mean() - y.mean()) < 5.0 * se, (v, x.mean(), y.mean(), se)
        qs = [0.1, 0.5, 0.9]
        assert np.allclose(np.quantile(x, qs), np.quantile(y, qs), rtol=0.03), (v, np.quantile(x, qs), np.quantile(y, qs))
        checked += 1
    assert checked >= 1


def test_aov_menger_sponge_instances(renderer):
    """scenes/menger.rs at 3 levels: 8,000 unit-cube instances of one 12-triangle BLAS under the TLAS."""
    world, camera = scenes.menger(levels=3)
    assert len(world.objects) == 8001
    renderer.set_scene(NativeScene(world, camera))
    g = renderer.render_aov(480, 270)
    o = OracleScene(world, camera).render_aov(480, 270)
    check_aov(g, o, albedo_exact=False, max_ties=int(0.03 * 480 * 270))  # axis-aligned unit cubes on an integer lattice: many pixel-exact edges
    assert len(np.unique(g["object"])) > 500
    stat_compare(renderer, world, camera, 96, 54, 32)


def test_aov_tlas_built_on_the_gpu(renderer):
    """MRT_OPT_DEVICE_BUILD = 2: a world of >= 16384 objects gets its TLAS built on the GPU too (mrt_lbvh.cuh, one object per leaf;
    by default only from 2^20 objects): 20,000 spheres and the instances of a cube, primary rays against the oracle; then the same
    world through the host's SAH builder."""
    rs = np.random.RandomState(4)
    world = World(SkyBackground())
    mats = [Lambertian(SolidColor((0.8, 0.3, 0.3, 1))), Metal(0.1, SolidColor((0.8, 0.8, 0.8, 1))), Dielectric(1.5)]
    c = rs.uniform(-30, 30, (20000, 3)).astype(np.float32)
    for i in range(20000):
        world.add(Sphere(mats[i % 3], V3(c[i, 0], abs(c[i, 1]) * 0.3 + 0.3, c[i, 2]), 0.3))
    from mass_raytrace_b200 import PlyLoader
    cube = Model(PlyLoader.load(scenes.CUBE_PLY))
    for i in range(200):
        world.add(cube.instance(V3(*rs.uniform(-25, 25, 3)), V3(*rs.uniform(0, 1, 3)), V3(0.5, 0.5, 0.5)).with_material(mats[0]))
    world.add(Sphere(mats[0], V3(0, -1000, 0), 1000.0))
    world.build_bvh()
    camera = Camera(40.0, V3(0, 12, 45), V3(0, 2, 0), V3(0, 1, 0), 16.0 / 9.0, 0.0, 45.0)
    o = OracleScene(world, camera).render_aov(480, 270)
    try:
        for device_build in (2, 0):
            renderer.set_option(renderer.OPT_DEVICE_BUILD, device_build)
            renderer.set_scene(NativeScene(world, camera))
            g = renderer.render_aov(480, 270)
            check_aov(g, o, albedo_exact=False)
            assert len(np.unique(g["object"])) > 2000
    finally:
        renderer.set_option(renderer.OPT_DEVICE_BUILD, 1)


def test_aov_world_without_bvh(renderer):
    # World::intersect before build_bvh: linear closest hit over the object list (world.rs:131-144); > 8 roots uses the list path
    for n_spheres in (5, 14):
        w = World(SkyBackground())
        for i in range(n_spheres):
            w.add(Sphere(Lambertian(SolidColor((0.1 * i % 1.0, 0.5, 0.5, 1))), V3(i * 1.1 - 5, 0.3 * (i % 3), -i * 0.5), 0.7))
        cam = Camera(50.0, V3(0, 1, 8), V3(0, 0, 0), V3(0, 1, 0), 1.5, 0.0, 8.0)
        renderer.set_scene(NativeScene(w, cam))
        g = renderer.render_aov(300, 200)
        o = OracleScene(w, cam).render_aov(300, 200)
        check_aov(g, o)
        assert len(np.unique(g["object"])) >= n_spheres


def test_aov_full_size_million_triangle_mesh(renderer, tmp_mesh_dir):
    """cfg 3 at its real size: 1,048,576-triangle PLY under the triangle BVH, 1920x1080 primary rays, against the oracle."""
    path = str(tmp_mesh_dir / "mesh_1m.ply")
    n, md = scenes.write_synthetic_ply(path, 1024, 512, seed=1)
    assert n == 1 << 20
    world, camera = scenes.lucy_layout(path, md, grid=0)
    host = NativeScene(world, camera)
    assert host.desc().contents.n_tris == (1 << 20) + 12
    o = OracleScene(world, camera).render_aov(1920, 1080)
    try:
        for device_build in (1, 0):  # the GPU's LBVH (default for a mesh this size) and the host's SAH tree
            renderer.set_option(renderer.OPT_DEVICE_BUILD, device_build)
            renderer.set_scene(host)
            g = renderer.render_aov(1920, 1080)
            ties, frac = check_aov(g, o)
            assert (g["tri"] != NONE).mean() > 0.2
    finally:
        renderer.set_option(renderer.OPT_DEVICE_BUILD, 1)


# ------------------------------------------------------------------------------------------------------------------
# test 2: converged renders, statistical
# ------------------------------------------------------------------------------------------------------------------
def stat_compare(renderer, world, camera, w, h, spp, oracle_a=None, seed=2024, rmse_factor=1.15, orc=None, upload=True, by_rows=False):
    if upload:
        renderer.set_scene(NativeScene(world, camera))
    rgb, bounces, count = renderer.render(w, h, spp, 50, seed=seed)
    assert count == spp and np.isfinite(rgb).all() and (rgb >= 0).all()
    if orc is None:
        orc = OracleScene(world, camera)
    if oracle_a is None:
        a_rgb, a_b, _ = orc.render(w, h, spp, 50, seed=seed, by_rows=by_rows)
    else:
        a_rgb, a_b = oracle_a
    b_rgb, b_b, _ = orc.render(w, h, spp, 50, seed=seed + 1, by_rows=by_rows)
    rmse = lambda x, y: float(np.sqrt(np.mean((x.astype(np.float64) / spp - y.astype(np.float64) / spp) ** 2)))
    rmse_oo = rmse(a_rgb, b_rgb)
    rmse_go = 0.5 * (rmse(rgb, a_rgb) + rmse(rgb, b_rgb))
    assert rmse_go <= rmse_factor * rmse_oo + 1e-7, f"RMSE gpu-oracle {rmse_go:.5g} vs oracle-oracle {rmse_oo:.5g}"
    npix = w * h
    lum = lambda x: float((x.astype(np.float64) * Y).sum(-1).mean() / spp)
    lum_o = 0.5 * (lum(a_rgb) + lum(b_rgb))
    # sigma of an image mean from the per-pixel seed-to-seed spread (pixels are independent)
    sig_lum = float(np.sqrt(np.mean((((a_rgb.astype(np.float64) - b_rgb) * Y).sum(-1) / spp) ** 2) / 2.0 / npix))
    tol = max(0.005 * lum_o, 3.0 * sig_lum * np.sqrt(1.5))
    assert abs(lum(rgb) - lum_o) <= tol, f"mean luminance gpu {lum(rgb):.6g} oracle {lum_o:.6g} tol {tol:.3g}"
    mb = lambda x: float(x.astype(np.float64).mean() / spp)
    mb_o = 0.5 * (mb(a_b) + mb(b_b))
    sig_b = float(np.sqrt(np.mean(((a_b.astype(np.float64) - b_b) / spp) ** 2) / 2.0 / npix))
    tolb = max(0.005 * mb_o, 3.0 * sig_b * np.sqrt(1.5))
    assert abs(mb(bounces) - mb_o) <= tolb, f"mean bounces gpu {mb(bounces):.6g} oracle {mb_o:.6g} tol {tolb:.3g}"
    return dict(rmse_go=rmse_go, rmse_oo=rmse_oo, lum=lum(rgb), lum_o=lum_o, bounces=mb(bounces), bounces_o=mb_o)


def test_render_cornell_vs_golden(renderer):
    gold = np.load(os.path.join(HERE, "golden", "cornell.npz"))
    world, camera = scenes.cornell_box(1.0)
    h, w = gold["sum_bounces"].shape
    stat_compare(renderer, world, camera, w, h, int(gold["spp"]), oracle_a=(gold["sum_rgb"], gold["sum_bounces"]), seed=int(gold["seed"]))


def test_render_book1_vs_golden(renderer):
    gold = np.load(os.path.join(HERE, "golden", "book1.npz"))
    world, camera = scenes.book1_spheres(1.5, aperture=0.0)
    h, w = gold["sum_bounces"].shape
    stat_compare(renderer, world, camera, w, h, int(gold["spp"]), oracle_a=(gold["sum_rgb"], gold["sum_bounces"]), seed=int(gold["seed"]))


def test_render_book1_defocus(renderer):
    world, camera = scenes.book1_spheres(1.5, aperture=0.1)  # thin-lens sampling world.rs:54-62
    stat_compare(renderer, world, camera, 96, 64, 96)


def test_render_cornell_tight_mean(renderer):
    # many samples on a small image: the image means must agree tightly (catches small biases that RMSE hides)
    world, camera = scenes.cornell_box(1.0)
    r = stat_compare(renderer, world, camera, 24, 24, 1024)
    assert abs(r["lum"] - r["lum_o"]) < 0.02 * r["lum_o"]
    assert abs(r["bounces"] - r["bounces_o"]) < 0.01 * r["bounces_o"]


def test_render_instanced_mesh(renderer, mesh_ply):
    path, md, n = mesh_ply
    world, camera = scenes.lucy_layout(path, md, grid=1)
    stat_compare(renderer, world, camera, 96, 54, 48)


def test_render_book2_volumes_textures(renderer):
    world, camera = scenes.book2_final(boxes_per_side=10, n_cluster=150)
    stat_compare(renderer, world, camera, 96, 54, 64)


def test_render_sphere_grid_metal_glass(renderer):
    world, camera = scenes.sphere_grid(dim=12)
    stat_compare(renderer, world, camera, 96, 54, 64)


def test_render_composite_materials_and_surfaces(renderer):
    # Mix (independent coins in scatter and emit, material.rs:403-417), Specular (:352-378), TextureBlend / SolidColorFallback /
    # YCbCrTexture surfaces (texture.rs:207-360) on UV meshes, under a SkyBackground
    rs = np.random.RandomState(3)
    tex_a = Texture(rs.randint(0, 256, (16, 32, 4)).astype(np.uint8), WRAP_REPEAT)
    tex_b = Texture(rs.randint(0, 256, (8, 8, 4)).astype(np.uint8), WRAP_CLAMP)
    luma = Texture(rs.randint(60, 200, (8, 8, 4)).astype(np.uint8), WRAP_CLAMP)
    chroma = Texture(rs.randint(100, 156, (8, 8, 4)).astype(np.uint8), WRAP_CLAMP)
    w = World(SkyBackground())
    w.add(Sphere(Lambertian(SolidColor((0.5, 0.5, 0.5, 1))), V3(0, -1000, 0), 1000.0))
    w.add(Sphere(Mix(0.3, DiffuseLight(V3(1.5, 1.2, 0.6)), Lambertian(SolidColor((0.8, 0.3, 0.3, 1)))), V3(-2.2, 1, 0), 1.0))
    w.add(Sphere(Specular(1.5, SolidColor((0.2, 0.4, 0.9, 1))), V3(0, 1, 0), 1.0))
    w.add(Sphere(Mix(0.5, Metal(0.2, SolidColor((0.9, 0.9, 0.9, 1))), Mix(0.5, Dielectric(1.5), Lambertian(SolidColor((0.1, 0.8, 0.1, 1))))), V3(2.2, 1, 0), 1.0))
    blend = TextureBlend(BLEND_ADDITION, tex_a, SolidColorFallback((0.2, 0.1, 0.0, 1.0), tex_b))
    w.add(Model(scenes.uv_sphere_triangles((-1.1, 0.6, 2.0), 0.6, 24, 12, material=Lambertian(blend))))
    w.add(Model(scenes.uv_sphere_triangles((1.1, 0.6, 2.0), 0.6, 24, 12, material=Metal(0.05, YCbCrTexture(luma, chroma)))))
    w.build_bvh()
    cam = Camera(35.0, V3(0, 2.5, 9), V3(0, 0.8, 0), V3(0, 1, 0), 1.5, 0.0, 9.0)
    renderer.set_scene(NativeScene(w, cam))
    g = renderer.render_aov(240, 160)
    o = OracleScene(w, cam).render_aov(240, 160)
    check_aov(g, o, albedo_exact=False)
    lam_mesh = (o["object"] == 4) & (g["tri"] == o["tri"])
    assert lam_mesh.sum() > 200 and np.array_equal(g["albedo"][lam_mesh], o["albedo"][lam_mesh])  # blend + fallback surfaces bit-identical
    stat_compare(renderer, w, cam, 72, 48, 128)


def test_background_kinds(renderer):
    rs = np.random.RandomState(5)
    sky = Texture(rs.randint(0, 256, (32, 64, 4)).astype(np.uint8), WRAP_CLAMP)
    faces = [Texture(rs.randint(0, 256, (8, 8, 4)).astype(np.uint8), WRAP_CLAMP) for _ in range(6)]
    cam = Camera(90.0, V3(0, 0, 0), V3(0.3, 0.2, -1), V3(0, 1, 0), 1.5, 0.0, 1.0)
    for bg in (SkySphere(sky), CubeMap(*faces, rotation=V3(0.1, 0.2, 0.05))):
        w = World(bg)
        w.add(Sphere(Metal(0.0, SolidColor((1, 1, 1, 1))), V3(0.3, 0.2, -3), 0.5))
        w.build_bvh()
        renderer.set_scene(NativeScene(w, cam))
        g = renderer.render_aov(192, 128)
        o = OracleScene(w, cam).render_aov(192, 128)
        miss = np.isinf(o["t"])
        assert np.array_equal(np.isinf(g["t"]), miss)
        # acos/atan2 (SkySphere) may differ from libm by an ulp and move a lookup across a texel edge on a few pixels
        close = np.abs(g["albedo"] - o["albedo"]).max(-1) <= 1e-4
        assert close[miss].mean() > 0.995, type(bg).__name__
        stat_compare(renderer, w, cam, 48, 32, 32)


# ------------------------------------------------------------------------------------------------------------------
# exact properties
# ------------------------------------------------------------------------------------------------------------------
def test_split_and_pool_invariance(renderer):
    """Any split of the sample range, and any pool size (i.e. any scheduling order), gives a bit-identical image."""
    world, camera = scenes.book1_spheres(1.5, aperture=0.1)
    renderer.set_scene(NativeScene(world, camera))
    w, h, spp = 160, 100, 12
    ref = renderer.render(w, h, spp, 50, seed=9)
    renderer.reset(w, h)
    renderer.accumulate(7, 5, 50, seed=9)
    renderer.accumulate(0, 3, 50, seed=9)
    renderer.accumulate(3, 4, 50, seed=9)
    rgb, b, count = renderer.download()
    assert count == spp and np.array_equal(rgb, ref[0]) and np.array_equal(b, ref[1])
    try:
        renderer.set_option(renderer.OPT_POOL_SLOTS, 4096)
        small = renderer.render(w, h, spp, 50, seed=9)
        assert renderer.stats()["pool_slots"] == 4096
    finally:
        renderer.set_option(renderer.OPT_POOL_SLOTS, 0)
    assert np.array_equal(small[0], ref[0]) and np.array_equal(small[1], ref[1])
    other = renderer.render(w, h, spp, 50, seed=10)
    assert not np.array_equal(other[0], ref[0])
    shifted = renderer.render(w, h, spp, 50, seed=9, spp_begin=100)  # a different sample range is a different image
    assert not np.array_equal(shifted[0], ref[0])


def test_white_furnace_exact(renderer):
    w = World(SolidBackground(V3(1, 1, 1)))
    w.add(Sphere(Lambertian(SolidColor((1, 1, 1, 1))), V3(0, 0, 0), 1.0))
    w.build_bvh()
    cam = Camera(40.0, V3(0, 0, 4), V3(0, 0, 0), V3(0, 1, 0), 1.0, 0.0, 4.0)
    renderer.set_scene(NativeScene(w, cam))
    rgb, b, count = renderer.render(64, 64, 16, 50, seed=3)
    # every escaping path carries exactly 1.0; the only way to lose energy is a grazing scatter that slips inside the sphere and
    # then runs into the depth limit (contributes 0, world.rs:66) -- so sums are integers <= spp and almost always == spp
    assert (rgb == np.round(rgb)).all() and (rgb <= 16).all() and (rgb[..., 0] == rgb[..., 2]).all()
    assert (rgb == 16).mean() > 0.995
    trapped = rgb[..., 0] < 16
    assert (b[trapped] >= 50).all()
    assert b[32, 32] >= 16 and b[0, 0] == 0


def test_empty_world_and_emitter_exact(renderer):
    cam = Camera(40.0, V3(0, 0, 4), V3(0, 0, 0), V3(0, 1, 0), 1.0, 0.0, 4.0)
    w = World(SolidBackground(V3(0.25, 0.5, 0.75)))
    renderer.set_scene(NativeScene(w, cam))
    rgb, b, count = renderer.render(16, 16, 4, 50)
    assert np.array_equal(rgb, np.broadcast_to(np.array([1.0, 2.0, 3.0], np.float32), (16, 16, 3))) and not b.any()
    st = renderer.stats()
    assert st["rays"] == st["paths"] == 16 * 16 * 4
    aov = renderer.render_aov(16, 16)
    assert (aov["object"] == NONE).all() and np.isinf(aov["t"]).all() and not aov["normal"].any()
    w2 = World(SolidBackground(V3(0, 0, 0)))
    w2.add(Sphere(DiffuseLight(V3(2, 3, 4)), V3(0, 0, 0), 1.0))
    w2.build_bvh()
    renderer.set_scene(NativeScene(w2, cam))
    rgb, b, _ = renderer.render(9, 9, 3, 50)
    assert rgb[4, 4].tolist() == [6, 9, 12] and b[4, 4] == 0 and not rgb[0, 0].any()
    assert renderer.render_aov(9, 9)["albedo"][4, 4].tolist() == [2, 3, 4]


def test_depth_limit_semantics(renderer):
    w = World(SolidBackground(V3(1, 1, 1)))
    w.add(Sphere(Lambertian(SolidColor((1, 1, 1, 1))), V3(0, 0, 0), -10.0))
    w.build_bvh()
    cam = Camera(40.0, V3(0, 0, 0), V3(0, 0, -1), V3(0, 1, 0), 1.0, 0.0, 1.0)
    renderer.set_scene(NativeScene(w, cam))
    rgb, b, _ = renderer.render(32, 32, 4, 1)
    assert not rgb.any() and (b == 4).all() and renderer.stats()["rays"] == 32 * 32 * 4
    orc = OracleScene(w, cam)
    for depth in (3, 50):
        rgb, b, _ = renderer.render(32, 32, 4, depth)
        orgb, ob, oc = orc.render(32, 32, 4, depth, seed=1)
        assert (b <= 4 * depth).all() and b.mean() > 0.97 * 4 * depth
        assert abs(b.mean() - ob.mean()) < 0.02 * ob.mean()
        assert abs(renderer.stats()["rays"] - oc["rays"]) < 0.02 * oc["rays"]


def test_ray_count_matches_bounce_bookkeeping(renderer):
    # rays = sum over samples of min(bounces + 1, max_depth)  (SURVEY.md §8d): bounded by the per-pixel bounce sums
    world, camera = scenes.cornell_box(1.0)
    renderer.set_scene(NativeScene(world, camera))
    rgb, b, count = renderer.render(128, 128, 8, 50, seed=4)
    st = renderer.stats()
    total_b, n = int(b.sum()), 128 * 128 * 8
    assert st["paths"] == n
    assert total_b + n - total_b // 50 <= st["rays"] <= total_b + n
    assert st["iterations"] >= 50 or st["rays"] < st["pool_slots"] * 50


def test_dense_and_thin_volume(renderer):
    cam = Camera(30.0, V3(0, 0, 4), V3(0, 0, 0), V3(0, 1, 0), 1.0, 0.0, 4.0)
    w = World(SolidBackground(V3(1, 1, 1)))
    w.add(Volume(Sphere((), V3(0, 0, 0), 1.0), 1000.0, V3(0.5, 0.5, 0.5)))
    w.add(Sphere(Lambertian(SolidColor((1, 0, 0, 1))), V3(0, 0, -3), 1.0))
    w.build_bvh()
    renderer.set_scene(NativeScene(w, cam))
    aov = renderer.render_aov(33, 33)
    assert aov["object"][16, 16] == 0 and aov["normal"][16, 16].tolist() == [1, 0, 0] and abs(aov["t"][16, 16] - 0.75) < 0.01
    stat_compare(renderer, w, cam, 48, 48, 64)
    thin = World(SolidBackground(V3(1, 1, 1)))
    thin.add(Volume(Sphere((), V3(0, 0, 0), 1.0), 1e-6, V3(0.5, 0.5, 0.5)))
    thin.build_bvh()
    renderer.set_scene(NativeScene(thin, cam))
    assert (renderer.render_aov(33, 33)["object"] == NONE).all()


def test_volume_over_model_and_instance_targets(renderer):
    """Volume<I: Intersect> (geom.rs:595-660) with a mesh as the target: the medium fills a Model (a UV sphere mesh, no transform) and a
    rotated, scaled Instance of the cube. Dense media: the primary ray scatters right behind the surface it enters through; media of
    moderate density: per-pixel hit frequencies over 48 seeds agree with the oracle's like two samples of one probability; converged
    renders agree statistically."""
    build = lambda density: mesh_media_scene(density)[0]
    cam = mesh_media_scene(1.0)[1]

    W, H = 150, 100
    # dense: the hit is the Volume, t just behind the entry surface, normal (1, 0, 0) (geom.rs:644-651)
    w = build(1.0e4)
    empty = build(1.0e-7)
    renderer.set_scene(NativeScene(w, cam))
    g = renderer.render_aov(W, H, seed=3)
    o = OracleScene(w, cam).render_aov(W, H, seed=3)
    vols = [1, 2]
    gm, om = np.isin(g["object"], vols), np.isin(o["object"], vols)
    assert 0.08 < om.mean() < 0.5
    assert (gm != om).mean() < 0.002  # silhouette pixels whose chord is shorter than a free flight
    both = gm & om
    assert np.array_equal(g["object"][both], o["object"][both])
    assert np.abs(g["t"][both] - o["t"][both]).max() < 5e-3 and (g["normal"][both] == np.array([1, 0, 0], np.float32)).all()
    check_aov(g, o, exclude=gm | om, albedo_exact=False)
    renderer.set_scene(NativeScene(empty, cam))
    assert not np.isin(renderer.render_aov(W, H)["object"], vols).any()
    # moderate density: per-pixel frequencies over many seeds
    w = build(0.9)
    renderer.set_scene(NativeScene(w, cam))
    orc = OracleScene(w, cam)
    n_seeds = 48
    fg, fo = np.zeros((H, W)), np.zeros((H, W))
    for seed in range(1, n_seeds + 1):
        fg += np.isin(renderer.render_aov(W, H, seed=seed)["object"], vols)
        fo += np.isin(orc.render_aov(W, H, seed=500 + seed)["object"], vols)
    a, b = fg / n_seeds, fo / n_seeds
    p = 0.5 * (a + b)
    assert np.array_equal(p > 0, (a > 0) | (b > 0)) or ((p > 0) != ((a > 0) & (b > 0))).mean() < 0.01
    mid = (p > 0.1) & (p < 0.9)
    assert mid.sum() > 500
    z = (a[mid] - b[mid]) / np.sqrt(2.0 * p[mid] * (1.0 - p[mid]) / n_seeds)
    assert np.abs(z).max() < 6.0 and 0.75 < float(np.mean(z * z)) < 1.3, (float(np.abs(z).max()), float(np.mean(z * z)))
    assert abs(float(np.mean(z))) < 5.0 / np.sqrt(mid.sum())
    stat_compare(renderer, w, cam, 72, 48, 96)


def test_eve_material_tangent_space_normals(renderer):
    w, cam = eve_scene()
    renderer.set_scene(NativeScene(w, cam))
    g = renderer.render_aov(300, 200)
    o = OracleScene(w, cam).render_aov(300, 200)
    check_aov(g, o, albedo_exact=False)  # the normals include the texture's tilt: bit-identical like everything else in a hit record
    on_mesh = np.isin(o["object"], [1, 2]) & (g["tri"] == o["tri"])
    assert on_mesh.sum() > 5000
    # the same meshes with a flat normal map: the hook returns (0, 0, 1) and the shading normal is the interpolated one
    wf, _ = eve_scene(flat_normals=True)
    renderer.set_scene(NativeScene(wf, cam))
    f = renderer.render_aov(300, 200)
    tilted = np.abs(g["normal"] - f["normal"]).max(-1) > 1e-3
    assert tilted[on_mesh].mean() > 0.9 and not tilted[~np.isin(o["object"], [1, 2])].any()
    assert np.array_equal(f["t"], g["t"])
    # albedo_normal's albedo (world.rs:81-93) is the scatter attenuation: a Mix coin decides between the surface colour and (1, 1, 1), so
    # only its distribution can agree; emission and colour are covered by the converged render
    stat_compare(renderer, w, cam, 90, 60, 128)
    r = renderer.render(90, 60, 64, 50, seed=3)
    assert renderer.stats()["rays"] > 90 * 60 * 64 * 1.5


@pytest.mark.parametrize("name", ["eve", "mesh_media"])
def test_round2_scenes_vs_golden(renderer, name):
    """The committed fixtures of the two round-2 features (tests/golden/make_golden.py): primary rays outside the media bit for bit,
    the converged render against the fixture as one of the two oracle images."""
    gold = np.load(os.path.join(HERE, "golden", f"{name}.npz"))
    world, camera = eve_scene() if name == "eve" else mesh_media_scene(0.9)
    renderer.set_scene(NativeScene(world, camera))
    h, w = gold["aov_object"].shape
    g = renderer.render_aov(w, h, seed=int(gold["seed"]))
    o = dict(object=gold["aov_object"], tri=gold["aov_tri"], t=gold["aov_t"], normal=gold["aov_normal"], albedo=gold["aov_albedo"])
    check_aov(g, o, exclude=volume_mask(world, g, o), albedo_exact=False)
    h, w = gold["sum_bounces"].shape
    stat_compare(renderer, world, camera, w, h, int(gold["spp"]), oracle_a=(gold["sum_rgb"], gold["sum_bounces"]), seed=int(gold["seed"]))


def test_resolve_rgb8_matches_reference_tonemap(renderer, oracle):
    from mass_raytrace_b200 import _ffi

    world, camera = scenes.book1_spheres(1.5)
    renderer.set_scene(NativeScene(world, camera))
    w, h, spp = 120, 80, 16
    rgb, b, count = renderer.render(w, h, spp, 50, seed=2)
    got = renderer.resolve_rgb8(count, mode=0, flip=True)
    want = np.zeros((h, w, 3), np.uint8)
    oracle.orc_resolve_rgb8(rgb.ctypes.data_as(_ffi.f32p), w, h, count, 1, want.ctypes.data_as(_ffi.u8p))
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1  # powf differs from libm by an ulp at most
    assert (got == want).mean() > 0.99
    noflip = renderer.resolve_rgb8(count, mode=0, flip=False)
    assert np.array_equal(noflip[::-1], got)
    depth = renderer.resolve_rgb8(count, mode=2, flip=False)  # DisplayMode::Depth main.rs:655-665
    want_d = np.clip((b / spp) / (max(b.max(), 1) / spp), 0, 1)
    assert np.abs(depth[..., 0].astype(float) - np.floor(want_d * 255)).max() <= 1 and np.array_equal(depth[..., 0], depth[..., 2])
    assert not renderer.resolve_rgb8(0).any()  # count == 0 -> black


def test_philox_known_answers(renderer):
    import ctypes as C

    from mass_raytrace_b200 import _ffi

    def philox(ctr, key):
        c, k, out = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        assert renderer.lib.mrt_debug_philox(renderer._h, c, k, out) == 0
        return [hex(x) for x in out]

    # Random123 kat_vectors, philox4x32-10
    assert philox([0, 0, 0, 0], [0, 0]) == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_closed_form_samplers_match_rejection_sampler_distributions(renderer, oracle):
    """The kernels sample the unit disk / sphere / ball in closed form; the reference uses rejection loops (math.rs:80-109).
    Same distributions: compare moments and radial/axial CDFs of 400k draws from each."""
    from mass_raytrace_b200 import _ffi

    n = 400000
    gb, gs, gd = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
    assert renderer.lib.mrt_debug_samplers(renderer._h, 77, n, gb.ctypes.data_as(_ffi.f32p), gs.ctypes.data_as(_ffi.f32p), gd.ctypes.data_as(_ffi.f32p)) == 0
    ob, os_, od = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
    oracle.orc_kat_samplers(99, n, ob.ctypes.data_as(_ffi.f32p), os_.ctypes.data_as(_ffi.f32p), od.ctypes.data_as(_ffi.f32p))
    assert (np.sum(gb.astype(np.float64) ** 2, 1) <= 1 + 1e-6).all() and (np.sum(gd.astype(np.float64) ** 2, 1) <= 1 + 1e-6).all()
    np.testing.assert_allclose(np.linalg.norm(gs.astype(np.float64), axis=1), 1.0, atol=1e-6)

    def ks(a, b):  # two-sample Kolmogorov-Smirnov statistic
        a, b = np.sort(a), np.sort(b)
        allv = np.concatenate([a, b])
        return np.abs(np.searchsorted(a, allv, side="right") / len(a) - np.searchsorted(b, allv, side="right") / len(b)).max()

    crit = 1.95 * np.sqrt(2.0 / n)  # alpha ~ 0.001
    for k in range(3):
        assert ks(gb[:, k], ob[:, k]) < crit and ks(gs[:, k], os_[:, k]) < crit
    for k in range(2):
        assert ks(gd[:, k], od[:, k]) < crit
    assert ks(np.linalg.norm(gb, axis=1), np.linalg.norm(ob, axis=1)) < crit
    assert ks(np.linalg.norm(gd, axis=1), np.linalg.norm(od, axis=1)) < crit
    assert ks(np.arctan2(gs[:, 1], gs[:, 0]), np.arctan2(os_[:, 1], os_[:, 0])) < crit


def test_alpha_tested_triangles(renderer):
    """Triangle::intersect rejects a candidate whose own material fails alpha_test (geom.rs:567-571; Lambertian: texel alpha != 0,
    material.rs:222-224) and traversal continues behind it: holes in a textured mesh must match the oracle exactly."""
    rs = np.random.RandomState(11)
    px = rs.randint(0, 256, (16, 16, 4)).astype(np.uint8)
    px[..., 3] = 255
    px[:, (np.arange(16) // 4) % 2 == 0, 3] = 0  # transparent stripes 4 texels wide: bilinear alpha is exactly 0 inside them
    cam = Camera(35.0, V3(0, 1.5, 5), V3(0, 1, 0), V3(0, 1, 0), 1.5, 0.0, 5.0)

    def build(tex):
        holes = Lambertian(Texture(tex, WRAP_CLAMP))
        w = World(SkyBackground())
        w.add(Model(scenes.uv_sphere_triangles((0.0, 1.0, 0.0), 1.0, 32, 16, material=holes)))
        w.add(Model(scenes.uv_sphere_triangles((0.3, 1.0, -2.5), 1.0, 24, 12, material=Mix(0.5, holes, Metal(0.0, SolidColor((1, 1, 1, 1)))))))
        w.add(Sphere(Lambertian(SolidColor((0.8, 0.2, 0.2, 1))), V3(0, -1000, 0), 1000.0))
        w.build_bvh()
        return w

    opaque = px.copy()
    opaque[..., 3] = 255
    renderer.set_scene(NativeScene(build(opaque), cam))
    solid = renderer.render_aov(360, 240)
    w = build(px)
    for keep in (False, True):
        renderer.set_scene(NativeScene(w, cam), keep_topology=keep)
        g = renderer.render_aov(360, 240)
        o = OracleScene(w, cam).render_aov(360, 240)
        # the second mesh's Mix material flips a coin per candidate (material.rs:419-425): compare only pixels where neither side's
        # nearest surface is that mesh
        clean = (g["object"] != 1) & (o["object"] != 1)
        check_aov(g, o, exclude=~clean, albedo_exact=False)
        # covered by mesh 0 when opaque; now the ray passes a hole and ends on the far inside wall, the ground, mesh 1 or the sky
        def through(a):
            return (solid["object"] == 0) & ((a["object"] != 0) | (a["t"] > solid["t"] * 1.05))

        assert through(g).sum() > 1500 and ((solid["object"] == 0) & ~through(g)).sum() > 1500
        assert np.array_equal(through(g) & clean, through(o) & clean)
    stat_compare(renderer, w, cam, 72, 48, 96)


def test_sah_rebuild_visits_fewer_nodes(renderer, mesh_ply):
    """The library's SAH rebuild and the caller's median-split topology give the same image bit for bit (no ties in this scene's
    sample set would be luck -- compare statistically identical AOVs instead) and the rebuild must not visit more nodes."""
    path, md, n = mesh_ply
    world, camera = scenes.lucy_layout(path, md, grid=1)
    host = NativeScene(world, camera)
    visits = {}
    images = {}
    for keep in (True, False):
        renderer.set_scene(host, keep_topology=keep)
        renderer.set_option(renderer.OPT_COUNT_VISITS, 1)
        try:
            images[keep] = renderer.render(160, 90, 8, 50, seed=5)
            st = renderer.stats()
        finally:
            renderer.set_option(renderer.OPT_COUNT_VISITS, 0)
        visits[keep] = (st["node_visits"] / st["rays"], st["tri_tests"] / st["rays"])
        assert st["node_visits"] > 0 and st["tri_tests"] > 0
    # same samples, same primitives: identical sums unless an exact tie was resolved differently (rare) -> allow a handful of pixels
    assert (images[True][0] != images[False][0]).any(-1).mean() < 1e-3
    assert np.array_equal(images[True][1].sum(), images[False][1].sum()) or abs(int(images[True][1].sum()) - int(images[False][1].sum())) < 50
    assert visits[False][0] < visits[True][0], visits


# ------------------------------------------------------------------------------------------------------------------
# error behaviour at the boundary
# ------------------------------------------------------------------------------------------------------------------
def test_call_order_and_bad_arguments():
    import ctypes as C

    from mass_raytrace_b200 import MrtError, Renderer, _ffi

    r = Renderer(0)
    try:
        with pytest.raises(MrtError, match="no scene uploaded"):
            r.render(8, 8, 1)
        with pytest.raises(MrtError, match="no scene uploaded"):
            r.render_aov(8, 8)
        world, camera = scenes.cornell_box(1.0)
        host = NativeScene(world, camera)
        desc = host.desc().contents
        assert r.lib.mrt_scene_upload(r._h, C.byref(desc)) == 0
        with pytest.raises(MrtError, match="no camera"):
            r.render(8, 8, 1)
        r.set_scene(host)
        with pytest.raises(MrtError, match="image size"):
            r.render(1, 8, 1)
        with pytest.raises(MrtError, match="max_depth"):
            r.render(8, 8, 1, max_depth=0)
        rgb, b, count = r.render(8, 8, 0)  # zero samples: a cleared image
        assert count == 0 and not rgb.any() and not b.any()
        bad = _ffi.mrt_scene_desc.from_buffer_copy(desc)
        bad.abi_version = 99
        assert r.lib.mrt_scene_upload(r._h, C.byref(bad)) == -1 and b"abi_version" in r.lib.mrt_last_error(r._h)
        bad = _ffi.mrt_scene_desc.from_buffer_copy(desc)
        bad.n_nodes = 3  # children now point past the node array
        assert r.lib.mrt_scene_upload(r._h, C.byref(bad)) == -1
        bad = _ffi.mrt_scene_desc.from_buffer_copy(desc)
        bad.n_materials = 1
        assert r.lib.mrt_scene_upload(r._h, C.byref(bad)) == -1 and b"material" in r.lib.mrt_last_error(r._h)
        assert r.lib.mrt_scene_upload(r._h, None) == -1
        r.render(8, 8, 1)  # the previously uploaded scene is still intact after the rejected uploads
        with pytest.raises(MrtError, match="pool slots"):
            r.set_option(r.OPT_POOL_SLOTS, 5)
    finally:
        r.close()
    with pytest.raises(MrtError, match="out of range"):
        Renderer(1000)


def test_obj_scene_textured_and_alpha(renderer, tmp_path):
    """A scene that enters through ObjLoader + SimpleTexturedBuilder (obj_loader.rs:160-308, :332): MTL materials with `Kd` and `map_Kd`,
    the PNG decoded by the host library itself, alpha holes from the texture's alpha channel (geom.rs:567-571). Primary rays must
    match the oracle (whose PNG comes from PIL) bit for bit, albedo included."""
    from PIL import Image

    from mass_raytrace_b200 import ObjLoader, SimpleTexturedBuilder

    rs = np.random.RandomState(5)
    px = rs.randint(0, 256, (32, 32, 4)).astype(np.uint8)
    px[..., 3] = 255
    px[(np.arange(32) // 8) % 2 == 1, :, 3] = 0  # transparent bands
    Image.fromarray(px, "RGBA").save(str(tmp_path / "leaf.png"))
    open(str(tmp_path / "m.mtl"), "w").write("newmtl leaf\nmap_Kd leaf.png\nnewmtl clay\nKd 0.7 0.4 0.2\n")
    n = 24
    lines = ["mtllib m.mtl", "vn 0 0 1", "vn 0 1 0"]
    for j in range(n + 1):
        for i in range(n + 1):
            x, y = i / n * 4 - 2, j / n * 3
            lines.append(f"v {x:.6f} {y:.6f} {0.3 * np.sin(3 * x) * np.cos(2 * y):.6f}")
            lines.append(f"vt {i / n:.6f} {j / n:.6f}")
    lines.append("usemtl leaf")
    for j in range(n):
        for i in range(n):
            a, b, c, d = j * (n + 1) + i + 1, j * (n + 1) + i + 2, (j + 1) * (n + 1) + i + 2, (j + 1) * (n + 1) + i + 1
            lines += [f"f {a}/{a}/1 {b}/{b}/1 {c}/{c}/1", f"f {a}/{a}/1 {c}/{c}/1 {d}/{d}/1"]
    base = (n + 1) * (n + 1)
    lines += ["o ground", "v -6 0 -6", "v 6 0 -6", "v 6 0 6", "v -6 0 6", "usemtl clay",
              f"f {base + 1}/1/2 {base + 2}/1/2 {base + 3}/1/2", f"f {base + 1}/1/2 {base + 3}/1/2 {base + 4}/1/2"]
    open(str(tmp_path / "leafwall.obj"), "w").write("\n".join(lines) + "\n")
    w = World(SkyBackground())
    tris = ObjLoader.load(str(tmp_path / "leafwall.obj"), SimpleTexturedBuilder(WRAP_CLAMP))
    w.add(Model(tris))
    w.add(Sphere(Metal(0.0, SolidColor((0.9, 0.9, 0.9, 1))), V3(0, 1.2, -2.0), 1.2))
    w.build_bvh()
    cam = Camera(40.0, V3(0.5, 1.8, 6), V3(0, 1.3, 0), V3(0, 1, 0), 1.5, 0.0, 6.0)
    host = NativeScene(w, cam)
    renderer.set_scene(host)
    g = renderer.render_aov(360, 240)
    o = OracleScene(w, cam).render_aov(360, 240)
    check_aov(g, o, albedo_exact=True)
    behind = (g["object"] == 1).sum()  # the metal sphere seen through the holes of the wall
    assert behind > 500 and (g["object"] == 0).sum() > 20000
    # converged image through the same path
    spp = 64
    rgb, b, _ = renderer.render(180, 120, spp, 50, seed=3)
    orgb, ob, _ = OracleScene(w, cam).render(180, 120, spp, 50, seed=3)
    lum = lambda a: float((a @ Y).mean()) / spp
    assert abs(lum(rgb) - lum(orgb)) < 0.02 * lum(orgb)
    assert abs(b.mean() - ob.mean()) < 0.02 * ob.mean()


def test_device_built_bvh(renderer, mesh_ply):
    """MRT_OPT_DEVICE_BUILD (on by default): meshes of >= 16384 triangles get their BLAS built on the GPU (Morton-order LBVH, mrt_lbvh.cuh). A closest
    hit does not depend on the tree, so primary rays must match the oracle exactly as with the host's SAH tree -- instanced meshes,
    and a UV mesh whose texture alpha punches holes (the builder's gather kernel sets the alpha flag per triangle)."""
    path, md, n = mesh_ply
    world, camera = scenes.lucy_layout(path, md, grid=2)  # 25 instances of one 65,536-triangle mesh
    o = OracleScene(world, camera).render_aov(640, 360)
    try:
        renderer.set_option(renderer.OPT_DEVICE_BUILD, 1)
        renderer.set_scene(NativeScene(world, camera))
        g = renderer.render_aov(640, 360)
        check_aov(g, o)
        assert (g["tri"] != NONE).mean() > 0.3
        rgb_dev, b_dev, _ = renderer.render(160, 90, 16, 50, seed=9)
        renderer.set_option(renderer.OPT_DEVICE_BUILD, 0)
        renderer.set_scene(NativeScene(world, camera))
        rgb_host, b_host, _ = renderer.render(160, 90, 16, 50, seed=9)
        # same samples, same hits (up to exact-t ties on shared edges): the two images agree almost everywhere bit for bit
        assert (b_dev == b_host).mean() > 0.999 and np.abs(rgb_dev - rgb_host).max(-1).mean() < 1e-3

        rs = np.random.RandomState(2)
        px = rs.randint(0, 256, (16, 16, 4)).astype(np.uint8)
        px[..., 3] = 255
        px[:, (np.arange(16) // 4) % 2 == 0, 3] = 0
        w = World(SkyBackground())
        w.add(Model(scenes.uv_sphere_triangles((0.0, 1.0, 0.0), 1.0, 192, 96, material=Lambertian(Texture(px, WRAP_CLAMP)))))  # 36,864 triangles
        w.add(Sphere(Lambertian(SolidColor((0.8, 0.2, 0.2, 1))), V3(0, -1000, 0), 1000.0))
        w.build_bvh()
        cam = Camera(35.0, V3(0, 1.5, 5), V3(0, 1, 0), V3(0, 1, 0), 1.5, 0.0, 5.0)
        o2 = OracleScene(w, cam).render_aov(360, 240)
        renderer.set_option(renderer.OPT_DEVICE_BUILD, 1)
        renderer.set_scene(NativeScene(w, cam))
        g2 = renderer.render_aov(360, 240)
        check_aov(g2, o2, albedo_exact=False)
        near = g2["t"][g2["object"] == 0].min()
        assert ((g2["object"] == 0) & (g2["t"] > 1.3 * near)).sum() > 1000  # rays that went through a hole and hit the far inside of the sphere
    finally:
        renderer.set_option(renderer.OPT_DEVICE_BUILD, 1)


def test_scene_without_mesh_trees(renderer, mesh_ply):
    """mrth_defer_mesh_bvh (include/mrt_host.h): the host skips Model::new's median-split build and hands over mrt_blas.root =
    MRT_REF_NONE. The backend builds every BLAS from the triangle range itself (host SAH or GPU LBVH), so primary rays must match the
    oracle exactly as before; asking to keep a topology that is not there is an error, and leaves the uploaded scene in place."""
    path, md, n = mesh_ply
    world, camera = scenes.lucy_layout(path, md, grid=1)  # 9 instances of one 65,536-triangle mesh + ground cube + sun sphere
    o = OracleScene(world, camera).render_aov(480, 270)
    lazy = NativeScene(world, camera, defer_mesh_bvh=True)
    assert lazy.desc().contents.n_nodes < 64  # the TLAS only
    try:
        for mode in (0, 1):
            renderer.set_option(renderer.OPT_DEVICE_BUILD, mode)
            renderer.set_scene(lazy)
            g = renderer.render_aov(480, 270)
            check_aov(g, o)
            assert (g["tri"] != NONE).mean() > 0.3
        with pytest.raises(RuntimeError, match="KEEP_TOPOLOGY"):
            renderer.set_scene(lazy, keep_topology=True)
        check_aov(renderer.render_aov(480, 270), o)
        rgb, bounces, _ = renderer.render(160, 90, 8, 50, seed=3)
        renderer.set_scene(NativeScene(world, camera))
        rgb_ref, bounces_ref, _ = renderer.render(160, 90, 8, 50, seed=3)
        assert (bounces == bounces_ref).mean() > 0.999 and np.abs(rgb - rgb_ref).max(-1).mean() < 1e-3
    finally:
        renderer.set_option(renderer.OPT_DEVICE_BUILD, 1)

"""GPU parity tests of the analytic Sphere (src/shapes/sphere.rs; pbrt-rs_b200/csrc/sphere.cuh) against the CPU oracle: closest
hits / any hits of a scene that mixes triangles and spheres (translated, rotated + non-uniformly scaled, partial, reversed)
BIT FOR BIT, and per-sample radiance / box-filtered film of the path tracer with sphere-shaped objects and a spherical
DiffuseAreaLight under the three light-sampling strategies, BIT FOR BIT as well."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def OP(orc):
    from oracle import oracle_path
    return oracle_path


def _rays(orc, scenes, res, seed):
    cam = scenes.C2_CAMERA
    prim = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], res)
    rng = np.random.default_rng(seed)
    n = 60000
    rnd = np.zeros((n, 8), np.float32)
    rnd[:, 0:3] = rng.uniform(10, 540, size=(n, 3))
    rnd[:, 3] = np.where(rng.random(n) < 0.3, rng.uniform(50, 400, n), np.inf)       # finite t_max too (shadow-ray like)
    rnd[:, 4:7] = rng.normal(size=(n, 3))
    axis = rnd[:4000].copy()                                                          # zero direction components: the literal walk
    axis[np.arange(4000), 4 + (np.arange(4000) % 3)] = 0.0
    return np.concatenate([prim, rnd, axis])


def test_sphere_scene_hits_bit_exact(gpu, orc, OP, scenes):
    sc = scenes.scene_spheres()
    accel = gpu.BVHAccel(gpu.scene_from_dict(sc), max_prims_in_node=4)
    ref = OP.Scene(sc, 4).bvh()
    nodes, prims = accel.export()
    assert np.array_equal(prims, ref.ordered_prims()) and len(nodes) == ref.num_nodes
    assert np.array_equal(nodes["bounds"], ref.nodes()["bounds"])
    rays = _rays(orc, scenes, (256, 256), 21)
    hits, b0 = accel.intersect(rays, want_b0=True)
    rh, rb0, _ = ref.intersect(rays, want_b0=True)
    nt = len(sc["idx"])
    assert (hits["prim_id"][hits["prim_id"] != gpu.PB2_MISS] >= nt).sum() > 20000
    for k in range(len(sc["spheres"])):
        assert (rh["prim_id"] == nt + k).sum() > 10, f"sphere {k} is never hit"
    assert np.array_equal(hits["prim_id"], rh["prim_id"])
    for f in ("t", "b1", "b2"):
        assert np.array_equal(bits(hits[f]), bits(rh[f])), f
    assert np.array_equal(bits(b0), bits(rb0))
    occ = accel.intersect_p(rays)
    assert np.array_equal(occ, ref.intersect_p(rays)[0])


@pytest.mark.parametrize("split", [2, 3])
def test_sphere_scene_other_split_methods(gpu, orc, OP, scenes, split):
    sc = scenes.scene_spheres()
    accel = gpu.BVHAccel(gpu.scene_from_dict(sc), max_prims_in_node=2, split_method=split)
    ref = OP.Scene(sc, 4).bvh()                      # any tree gives the same closest hits
    rays = _rays(orc, scenes, (64, 64), 5)
    hits = accel.intersect(rays)
    rh = ref.intersect(rays)[0]
    assert np.array_equal(hits["prim_id"], rh["prim_id"]) and np.array_equal(bits(hits["t"]), bits(rh["t"]))


def test_hlbvh_refuses_spheres(gpu, scenes):
    sc = scenes.scene_spheres()
    with pytest.raises(gpu.Pb2Error) as e:
        gpu.BVHAccel(gpu.scene_from_dict(sc), max_prims_in_node=4, split_method=1)
    assert "triangles only" in str(e.value)


@pytest.mark.parametrize("strategy", ["uniform", "power", "spatial"])
def test_sphere_scene_per_sample_radiance_bit_exact(gpu, OP, scenes, strategy):
    sc = scenes.scene_spheres()
    cam = dict(scenes.C2_CAMERA, res=(256, 256))
    kw = dict(max_depth=6, rr_threshold=1.0, light_strategy=strategy, spp=16)
    accel = gpu.BVHAccel(gpu.scene_from_dict(sc), max_prims_in_node=4)
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = gpu.PathIntegrator(accel, camera, **kw)
    ref = OP.Scene(sc, 4)
    rng = np.random.default_rng(31)
    n = 20000 if strategy != "spatial" else 6000
    xy = rng.integers(0, 256, size=(n, 2))
    s = rng.integers(0, 16, size=n)
    L, pf = integ.li(xy, s)
    rL, rpf = ref.path_li(cam, OP.film_desc(cam["res"]), OP.path_desc(**kw), xy, s)
    assert np.array_equal(bits(pf), bits(rpf))
    assert (rL.sum(axis=1) > 0).mean() > 0.5
    mism = (bits(L) != bits(rL)).any(axis=1)
    assert mism.sum() == 0, f"{mism.sum()} of {n} samples differ; first: {L[mism][:3]} vs {rL[mism][:3]}"
    c = integ.counters()
    assert c["shadow_rays"] > 0 and c["mis_rays"] > 0


def test_sphere_scene_film_bit_exact(gpu, OP, scenes):
    sc = scenes.scene_spheres()
    cam = dict(scenes.C2_CAMERA, res=(160, 160))
    kw = dict(max_depth=5, rr_threshold=1.0, light_strategy="power", spp=8)
    accel = gpu.BVHAccel(gpu.scene_from_dict(sc), max_prims_in_node=4)
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = gpu.PathIntegrator(accel, camera, **kw)
    film = gpu.Film(cam["res"])
    integ.render(film)
    got = film.read_xyzw()
    want, _ = OP.Scene(sc, 4).render(cam, OP.film_desc(cam["res"]), OP.path_desc(**kw), mode=1)
    assert np.array_equal(bits(got), bits(want)), f"{(bits(got) != bits(want)).any(axis=2).sum()} pixels differ"
    assert 0.05 < film.resolve_rgb().mean() < 2.0


def test_triangle_scene_keeps_the_triangle_kernels(gpu, orc, scenes):
    """A scene without spheres runs k_closest_hit / k_any_hit (64 / 56 registers), not the sphere variants: same hits as before."""
    verts, idx = scenes.merge(scenes.uv_sphere(n_theta=30, n_phi=60), scenes.ground_grid())
    accel = gpu.BVHAccel(verts, idx, max_prims_in_node=4)
    cam = scenes.C1_CAMERA
    rays = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], (128, 128))
    hits = accel.intersect(rays)
    rh = orc.BVHAccel(verts, idx, 4).intersect(rays)[0]
    assert np.array_equal(hits["prim_id"], rh["prim_id"]) and np.array_equal(bits(hits["t"]), bits(rh["t"]))

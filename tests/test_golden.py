"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle at the commit that
introduced them).  CPU tests: the oracle still reproduces them bit for bit.  GPU tests: the CUDA path, through the C ABI,
reproduces the same bytes — so the oracle and the kernels cannot drift together unnoticed.  The reference holds no vector
for this path (parity unpinned): these are regression pins of the restatement, not outputs of the Rust crate."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _mk():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def G():
    return {k: np.load(os.path.join(HERE, "golden", k + ".npz")) for k in ("raycast", "path", "path_r02", "sphere_r02", "volpath_r02")}


@pytest.fixture(scope="module")
def OP(orc):
    from oracle import oracle_path
    return oracle_path


def u32(x):
    return np.ascontiguousarray(x).view(np.uint32)


SAMPLER_KW = {"halton": {}, "stratified": dict(x_samples=2, y_samples=2), "zerotwo": {}}


def check_raycast(g, hits, b0, occ, nodes, prims, wb):
    assert np.array_equal(hits["prim_id"], g["prim_id"])
    assert np.array_equal(u32(hits["t"]), g["t"]) and np.array_equal(u32(hits["b1"]), g["b1"]) and np.array_equal(u32(hits["b2"]), g["b2"])
    assert np.array_equal(u32(np.asarray(b0, np.float32)), g["b0"])
    assert np.array_equal(np.asarray(occ, np.uint8), g["occluded"])
    assert np.array_equal(np.frombuffer(np.ascontiguousarray(nodes).tobytes(), np.uint8), g["nodes"])
    assert np.array_equal(prims, g["ordered_prims"])
    assert np.array_equal(u32(np.asarray(wb, np.float32)), u32(g["world_bound"]))


def test_oracle_reproduces_raycast_fixture(orc, G):
    g = G["raycast"]
    assert (g["prim_id"] != 0xFFFFFFFF).sum() > 2000 and 100 < int(g["occluded"].sum()) < len(g["occluded"])
    bvh = orc.BVHAccel(g["verts"], g["idx"], 4)
    hits, b0 = bvh.intersect(g["rays"], want_b0=True)[:2]
    check_raycast(g, hits, b0, bvh.intersect_p(g["rays"])[0], bvh.nodes(), bvh.ordered_prims(), bvh.world_bound())


def test_oracle_reproduces_path_fixture(orc, OP, scenes, G):
    g, mk = G["path"], _mk()
    ref = OP.Scene(mk.golden_scene(scenes), 4)
    fd = OP.film_desc(mk.GOLDEN_CAMERA["res"])
    for strat in ("uniform", "power", "spatial"):
        L, pf = ref.path_li(mk.GOLDEN_CAMERA, fd, OP.path_desc(light_strategy=strat, **mk.GOLDEN_PATH), g["xy"], g["sample"])
        assert np.array_equal(u32(L), g["L_" + strat]) and np.array_equal(u32(pf), g["p_film"]), strat
    assert (g["L_uniform"] != g["L_spatial"]).any() and (g["L_uniform"] != g["L_power"]).any()
    for sampler, skw in SAMPLER_KW.items():
        L, _ = ref.path_li(mk.GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", sampler=sampler, **mk.GOLDEN_PATH, **skw), g["xy"], g["sample"])
        assert np.array_equal(u32(L), g["L_" + sampler]), sampler
    film, _ = ref.render(mk.GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="spatial", **mk.GOLDEN_PATH), mode=1)
    assert np.array_equal(u32(film), g["film_spatial"])
    assert ref.spatial_grid() == tuple(int(v) for v in g["spatial_grid"])
    for v, f, c in zip(g["spatial_voxels"], g["spatial_func"], g["spatial_cdf"]):
        rf, rc, _ = ref.spatial_voxel(v, 3)
        assert np.array_equal(u32(rf), f) and np.array_equal(u32(rc), c)


def test_oracle_reproduces_round2_fixture(orc, OP, scenes, G):
    g, g2, mk = G["path"], G["path_r02"], _mk()
    ref = OP.Scene(mk.golden_scene(scenes), 4)
    fd = OP.film_desc(mk.GOLDEN_CAMERA["res"])
    L, pf = ref.path_li(mk.GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", sampler="sobol", **mk.GOLDEN_PATH), g["xy"], g["sample"])
    assert np.array_equal(u32(L), g2["L_sobol"]) and np.array_equal(u32(pf), g2["p_film_sobol"])
    assert (g2["L_sobol"] != g["L_halton"]).any()
    film, _ = ref.render(mk.GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", sampler="sobol", **mk.GOLDEN_PATH), mode=1)
    assert np.array_equal(u32(film), g2["film_sobol"])


@pytest.mark.gpu
def test_gpu_reproduces_round2_fixture(gpu, scenes, G):
    g, g2, mk = G["path"], G["path_r02"], _mk()
    cam = mk.GOLDEN_CAMERA
    accel = gpu.BVHAccel(gpu.scene_from_dict(mk.golden_scene(scenes)), max_prims_in_node=4)
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = gpu.PathIntegrator(accel, camera, light_strategy="power", sampler="sobol", **mk.GOLDEN_PATH)
    L, pf = integ.li(g["xy"], g["sample"])
    assert np.array_equal(u32(L), g2["L_sobol"]) and np.array_equal(u32(pf), g2["p_film_sobol"])
    film = gpu.Film(cam["res"])
    integ.render(film)
    assert np.array_equal(u32(film.read_xyzw()), g2["film_sobol"])


@pytest.mark.gpu
def test_gpu_reproduces_raycast_fixture(gpu, G):
    g = G["raycast"]
    accel = gpu.BVHAccel(g["verts"], g["idx"], max_prims_in_node=4)
    hits, b0 = accel.intersect(g["rays"], want_b0=True)
    nodes, prims = accel.export()
    check_raycast(g, hits, b0, accel.intersect_p(g["rays"]), nodes, prims, accel.world_bound())


@pytest.mark.gpu
def test_gpu_reproduces_path_fixture(gpu, scenes, G):
    g, mk = G["path"], _mk()
    cam = mk.GOLDEN_CAMERA
    accel = gpu.BVHAccel(gpu.scene_from_dict(mk.golden_scene(scenes)), max_prims_in_node=4)
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    for strat in ("uniform", "power", "spatial"):
        L, pf = gpu.PathIntegrator(accel, camera, light_strategy=strat, **mk.GOLDEN_PATH).li(g["xy"], g["sample"])
        assert np.array_equal(u32(L), g["L_" + strat]) and np.array_equal(u32(pf), g["p_film"]), strat
    for sampler, skw in SAMPLER_KW.items():
        L, _ = gpu.PathIntegrator(accel, camera, light_strategy="power", sampler=sampler, **mk.GOLDEN_PATH, **skw).li(g["xy"], g["sample"])
        assert np.array_equal(u32(L), g["L_" + sampler]), sampler
    film = gpu.Film(cam["res"])
    gpu.PathIntegrator(accel, camera, light_strategy="spatial", **mk.GOLDEN_PATH).render(film)
    assert np.array_equal(u32(film.read_xyzw()), g["film_spatial"])
    nv, func, cdf, _ = accel.spatial_light_distribution()
    assert nv == tuple(int(v) for v in g["spatial_grid"])
    for v, f, c in zip(g["spatial_voxels"], g["spatial_func"], g["spatial_cdf"]):
        assert np.array_equal(u32(func[v[2], v[1], v[0]]), f) and np.array_equal(u32(cdf[v[2], v[1], v[0]]), c)


def _check_sphere_hits(g, hits, b0, occ, nodes, prims):
    assert np.array_equal(hits["prim_id"], g["prim_id"])
    assert np.array_equal(u32(hits["t"]), g["t"]) and np.array_equal(u32(hits["b1"]), g["b1"]) and np.array_equal(u32(hits["b2"]), g["b2"])
    assert np.array_equal(u32(np.asarray(b0, np.float32)), g["b0"])
    assert np.array_equal(np.asarray(occ, np.uint8), g["occluded"])
    assert np.array_equal(np.frombuffer(np.ascontiguousarray(nodes).tobytes(), np.uint8), g["nodes"]) and np.array_equal(prims, g["ordered_prims"])


def test_oracle_reproduces_sphere_fixture(orc, OP, scenes, G):
    g, mk = G["sphere_r02"], _mk()
    sc = mk.golden_sphere_scene(scenes)
    assert (g["prim_id"][g["prim_id"] != 0xFFFFFFFF] >= len(sc["idx"])).sum() > 1000          # analytic spheres are hit
    ref = OP.Scene(sc, 4)
    bvh = ref.bvh()
    hits, b0 = bvh.intersect(g["rays"], want_b0=True)[:2]
    _check_sphere_hits(g, hits, b0, bvh.intersect_p(g["rays"])[0], bvh.nodes(), bvh.ordered_prims())
    fd = OP.film_desc(mk.GOLDEN_CAMERA["res"])
    for strat in ("uniform", "power", "spatial"):
        L, _ = ref.path_li(mk.GOLDEN_CAMERA, fd, OP.path_desc(light_strategy=strat, **mk.GOLDEN_PATH), g["xy"], g["sample"])
        assert np.array_equal(u32(L), g["L_" + strat]), strat
    film, _ = ref.render(mk.GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", **mk.GOLDEN_PATH), mode=1)
    assert np.array_equal(u32(film), g["film_power"])


@pytest.mark.gpu
def test_gpu_reproduces_sphere_fixture(gpu, scenes, G):
    g, mk = G["sphere_r02"], _mk()
    cam = mk.GOLDEN_CAMERA
    accel = gpu.BVHAccel(gpu.scene_from_dict(mk.golden_sphere_scene(scenes)), max_prims_in_node=4)
    hits, b0 = accel.intersect(g["rays"], want_b0=True)
    nodes, prims = accel.export()
    _check_sphere_hits(g, hits, b0, accel.intersect_p(g["rays"]), nodes, prims)
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    for strat in ("uniform", "power", "spatial"):
        L, _ = gpu.PathIntegrator(accel, camera, light_strategy=strat, **mk.GOLDEN_PATH).li(g["xy"], g["sample"])
        assert np.array_equal(u32(L), g["L_" + strat]), strat
    film = gpu.Film(cam["res"])
    gpu.PathIntegrator(accel, camera, light_strategy="power", **mk.GOLDEN_PATH).render(film)
    assert np.array_equal(u32(film.read_xyzw()), g["film_power"])


def test_oracle_reproduces_volpath_fixture(orc, OP, scenes, G):
    g, mk = G["volpath_r02"], _mk()
    ref = OP.Scene(scenes.scene_media(), 4)
    fd = OP.film_desc(mk.GOLDEN_CAMERA["res"])
    kw = dict(mk.GOLDEN_PATH, max_depth=8)
    for strat in ("uniform", "power"):
        L, _ = ref.path_li(mk.GOLDEN_CAMERA, fd, OP.path_desc(light_strategy=strat, integrator="volpath", **kw), g["xy"], g["sample"])
        assert np.array_equal(u32(L), g["L_" + strat]), strat
    film, _ = ref.render(mk.GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", integrator="volpath", **kw), mode=1)
    assert np.array_equal(u32(film), g["film_power"])


@pytest.mark.gpu
def test_gpu_reproduces_volpath_fixture(gpu, scenes, G):
    g, mk = G["volpath_r02"], _mk()
    cam = mk.GOLDEN_CAMERA
    accel = gpu.BVHAccel(gpu.scene_from_dict(scenes.scene_media()), max_prims_in_node=4)
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    kw = dict(mk.GOLDEN_PATH, max_depth=8)
    for strat in ("uniform", "power"):
        L, _ = gpu.PathIntegrator(accel, camera, light_strategy=strat, integrator="volpath", **kw).li(g["xy"], g["sample"])
        assert np.array_equal(u32(L), g["L_" + strat]), strat
    film = gpu.Film(cam["res"])
    gpu.PathIntegrator(accel, camera, light_strategy="power", integrator="volpath", **kw).render(film)
    assert np.array_equal(u32(film.read_xyzw()), g["film_power"])

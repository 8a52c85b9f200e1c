"""GPU parity tests of the VolPathIntegrator + HomogeneousMedium (src/integrators/volpath.rs, src/media/homogeneous.rs;
k_volpath in pbrt-rs_b200/csrc/wavefront.cu) against the CPU oracle: per-sample radiance and box-filtered film BIT FOR BIT on a
scene with a fog-filled room, a dense smoke box behind a material-less interface, matte / glass surfaces, an analytic sphere and
area + point lights."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def OP(orc):
    from oracle import oracle_path
    return oracle_path


def _setup(gpu, OP, sc, cam, **kw):
    accel = gpu.BVHAccel(gpu.scene_from_dict(sc), max_prims_in_node=4)
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    return accel, camera, gpu.PathIntegrator(accel, camera, integrator="volpath", **kw), OP.Scene(sc, 4)


@pytest.mark.parametrize("strategy,sampler", [("uniform", "random"), ("power", "random"), ("spatial", "random"), ("power", "zerotwo")])
def test_media_scene_per_sample_radiance_bit_exact(gpu, OP, scenes, strategy, sampler):
    sc = scenes.scene_media()
    cam = dict(scenes.C2_CAMERA, res=(128, 128))
    kw = dict(max_depth=8, rr_threshold=1.0, light_strategy=strategy, spp=16, sampler=sampler)
    accel, camera, integ, ref = _setup(gpu, OP, sc, cam, **kw)
    rng = np.random.default_rng(41)
    n = 12000 if strategy != "spatial" else 4000
    xy = rng.integers(0, 128, size=(n, 2))
    s = rng.integers(0, 16, size=n)
    L, pf = integ.li(xy, s)
    rL, rpf = ref.path_li(cam, OP.film_desc(cam["res"]), OP.path_desc(integrator="volpath", **kw), xy, s)
    assert np.array_equal(bits(pf), bits(rpf))
    assert (rL.sum(axis=1) > 0).mean() > 0.5
    mism = (bits(L) != bits(rL)).any(axis=1)
    assert mism.sum() == 0, f"{mism.sum()} of {n} samples differ; first: {L[mism][:3]} vs {rL[mism][:3]}"


def test_media_scene_film_bit_exact(gpu, OP, scenes):
    sc = scenes.scene_media(g=-0.4)
    cam = dict(scenes.C2_CAMERA, res=(96, 96))
    kw = dict(max_depth=6, rr_threshold=1.0, light_strategy="power", spp=8)
    accel, camera, integ, ref = _setup(gpu, OP, sc, cam, **kw)
    film = gpu.Film(cam["res"])
    integ.render(film)
    got = film.read_xyzw()
    want, _ = ref.render(cam, OP.film_desc(cam["res"]), OP.path_desc(integrator="volpath", **kw), mode=1)
    assert np.array_equal(bits(got), bits(want)), f"{(bits(got) != bits(want)).any(axis=2).sum()} pixels differ"
    c = integ.counters()
    assert c["shadow_rays"] > 0 and c["extend_rays"] > 96 * 96 * 8


def test_volpath_without_media_equals_path_on_the_device(gpu, scenes):
    """No media, no specular surface, no Russian roulette before bounce 4: k_volpath and the wavefront PathIntegrator return the
    same radiance bit for bit (two independent device implementations of the same surface transport)."""
    sc = scenes.scene_spheres()
    sc["spheres"][2]["material"] = 3
    cam = dict(scenes.C2_CAMERA, res=(64, 64))
    kw = dict(max_depth=4, rr_threshold=1.0, light_strategy="power", spp=4)
    accel = gpu.BVHAccel(gpu.scene_from_dict(sc), max_prims_in_node=4)
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    rng = np.random.default_rng(7)
    xy = rng.integers(0, 64, size=(3000, 2))
    s = rng.integers(0, 4, size=3000)
    a, _ = gpu.PathIntegrator(accel, camera, **kw).li(xy, s)
    b, _ = gpu.PathIntegrator(accel, camera, integrator="volpath", **kw).li(xy, s)
    assert np.array_equal(bits(a), bits(b))


def test_volpath_argument_checks(gpu, scenes):
    sc = scenes.scene_media()
    accel = gpu.BVHAccel(gpu.scene_from_dict(sc), max_prims_in_node=4)
    cam = dict(scenes.C2_CAMERA, res=(16, 16))
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    film = gpu.Film(cam["res"])
    with pytest.raises(gpu.Pb2Error) as e:              # material-less interface surfaces need the VolPathIntegrator
        gpu.PathIntegrator(accel, camera, spp=1).render(film)
    assert "VOLPATH" in str(e.value)
    with pytest.raises(gpu.Pb2Error) as e:
        gpu.PathIntegrator(accel, camera, spp=1, sampler="halton", integrator="volpath").render(film)
    assert "unbounded" in str(e.value)

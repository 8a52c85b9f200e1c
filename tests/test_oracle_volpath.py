"""CPU tests of the oracle's VolPathIntegrator + HomogeneousMedium (oracle/oracle_path.hpp; reference src/integrators/volpath.rs,
src/media/homogeneous.rs, src/core/medium.rs): the contract exp / ln, volpath == path on a scene without media, Beer-Lambert
attenuation through an absorbing slab, and the furnace property of a scattering medium with albedo 1."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import oracle_path as OP


def test_contract_exp_log_accuracy():
    L = O.lib()
    xs = np.linspace(-80.0, 20.0, 5001, dtype=np.float32)
    got = np.array([L.orc_exp(float(x)) for x in xs], np.float64)
    want = np.exp(xs.astype(np.float64))
    assert np.max(np.abs(got / want - 1.0)) < 3e-7
    assert L.orc_exp(-100.0) == 0.0 and L.orc_exp(0.0) == 1.0
    us = np.concatenate([np.linspace(2.0 ** -24, 1.0, 4001), np.geomspace(1e-7, 1e6, 2001)]).astype(np.float32)
    got = np.array([L.orc_log(float(u)) for u in us], np.float64)
    want = np.log(us.astype(np.float64))
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) < 3e-7
    assert L.orc_log(1.0) == 0.0


def test_volpath_equals_path_without_media(scenes):
    """No medium anywhere and no Russian roulette before bounce 4: the surface branch of VolPathIntegrator::li (volpath.rs:115-187)
    consumes the sampler like PathIntegrator::li and its transmittance rays see the same occluders, so radiance is equal bit for bit."""
    sc = scenes.scene_spheres()
    # (volpath.rs:137-146 calls uniform_sample_one_light at EVERY surface vertex, path.rs:105-121 only where the BSDF has a
    # non-specular lobe: on a specular surface volpath draws five sampler values path does not — so no glass in this comparison)
    sc["spheres"][2]["material"] = 3
    ref = OP.Scene(sc, 4)
    cam = dict(scenes.C2_CAMERA, res=(64, 64))
    fd = OP.film_desc(cam["res"])
    rng = np.random.default_rng(7)
    xy = rng.integers(0, 64, size=(3000, 2))
    s = rng.integers(0, 4, size=3000)
    kw = dict(max_depth=4, rr_threshold=1.0, light_strategy="power", spp=4)
    a, _ = ref.path_li(cam, fd, OP.path_desc(**kw), xy, s)
    b, _ = ref.path_li(cam, fd, OP.path_desc(integrator="volpath", **kw), xy, s)
    assert (a.sum(axis=1) > 0).mean() > 0.5
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _slab_scene(sigma_a, sigma_s, g=0.0, emissive=5.0, kd=0.0):
    """An emissive wall at z = 10 seen through a medium-filled box spanning z in [2, 6] (material-less faces)."""
    from __graft_entry__ import load_scenes
    scenes = load_scenes()
    wall = (np.array([[-50, -50, 10], [50, -50, 10], [50, 50, 10], [-50, 50, 10]], np.float64), np.array([[0, 2, 1], [0, 3, 2]], np.int64))
    box = scenes._box((-40.0, -40.0, 2.0), (40.0, 40.0, 6.0))
    verts, idx = scenes.merge(wall, box)
    tm = np.array([0, 0] + [scenes.NO_MATERIAL] * 12, np.uint32)
    inside = np.full(14, -1, np.int32)
    inside[2:] = 0
    return dict(verts=verts, idx=idx, tri_material=tm, materials=[dict(type="matte", kd=(kd, kd, kd))],
                lights=[dict(type="area", prim=0, L=(emissive,) * 3, two_sided=True), dict(type="area", prim=1, L=(emissive,) * 3, two_sided=True)],
                media=[dict(sigma_a=sigma_a, sigma_s=sigma_s, g=g)], prim_inside=inside, prim_outside=np.full(14, -1, np.int32), camera_medium=-1)


def test_beer_lambert_through_an_absorbing_slab():
    sig = (0.1, 0.25, 0.5)
    sc = _slab_scene(sig, (0.0, 0.0, 0.0))
    ref = OP.Scene(sc, 4)
    cam = dict(pos=(0.0, 0.0, 0.0), look=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0), fov=5.0, res=(4, 4))
    pd = OP.path_desc(max_depth=3, rr_threshold=0.0, light_strategy="uniform", spp=4096, integrator="volpath")
    xyzw, _ = ref.render(cam, OP.film_desc(cam["res"]), pd, mode=1)
    rgb = OP.resolve_rgb(xyzw).reshape(-1, 3).mean(axis=0)
    want = 5.0 * np.exp(-np.array(sig) * 4.0)               # the slab is 4 units thick along the (nearly axial) rays
    assert np.allclose(rgb, want, rtol=0.03), (rgb, want)


@pytest.mark.parametrize("g", [0.0, 0.6])
def test_scattering_medium_in_a_furnace_conserves_radiance(g):
    """A closed box whose six black walls emit L = 1, filled by a medium of albedo 1: every path that does not end early sees
    radiance 1 however it scatters, so the estimate must return ~1 (NEE with transmittance + phase-sampled MIS + transport)."""
    from __graft_entry__ import load_scenes
    scenes = load_scenes()
    verts, idx = scenes._box((-10.0, -10.0, -10.0), (10.0, 10.0, 10.0))
    idx = idx[:, [0, 2, 1]]                                  # normals inward
    sc = dict(verts=verts, idx=idx, tri_material=np.zeros(12, np.uint32), materials=[dict(type="matte", kd=(0.0, 0.0, 0.0))],
              lights=[dict(type="area", prim=k, L=(1.0, 1.0, 1.0), two_sided=False) for k in range(12)],
              media=[dict(sigma_a=(0.0, 0.0, 0.0), sigma_s=(0.15, 0.15, 0.15), g=g)], prim_inside=np.zeros(12, np.int32),
              prim_outside=np.zeros(12, np.int32), camera_medium=0)
    ref = OP.Scene(sc, 4)
    cam = dict(pos=(0.0, 0.0, -5.0), look=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0), fov=40.0, res=(8, 8))
    pd = OP.path_desc(max_depth=60, rr_threshold=0.0, light_strategy="uniform", spp=256, integrator="volpath")
    xyzw, _ = ref.render(cam, OP.film_desc(cam["res"]), pd, mode=1)
    rgb = OP.resolve_rgb(xyzw)
    assert rgb.mean() == pytest.approx(1.0, rel=0.03), rgb.mean()


def test_media_scene_renders_and_differs_from_vacuum(scenes):
    sc = scenes.scene_media()
    ref = OP.Scene(sc, 4)
    cam = dict(scenes.C2_CAMERA, res=(48, 48))
    kw = dict(max_depth=6, rr_threshold=1.0, light_strategy="power", spp=16)
    vol, _ = ref.render(cam, OP.film_desc(cam["res"]), OP.path_desc(integrator="volpath", **kw), mode=1)
    sc2 = dict(sc)
    sc2.pop("media")
    vac, _ = OP.Scene(sc2, 4).render(cam, OP.film_desc(cam["res"]), OP.path_desc(integrator="volpath", **kw), mode=1)
    a, b = OP.resolve_rgb(vol), OP.resolve_rgb(vac)
    assert np.isfinite(a).all() and a.mean() > 0.02
    assert abs(a.mean() - b.mean()) / b.mean() > 0.02       # the smoke box and the fog change the picture

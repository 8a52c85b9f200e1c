"""CPU tests of the oracle's path-tracing half (PathIntegrator::li, BSDFs, lights, Film) — analytic and self-consistency
pins, since the reference holds no golden image or known-answer test for this path (parity unpinned)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def OP(orc):
    from oracle import oracle_path
    return oracle_path


def test_white_furnace_converges_to_closed_form(OP, scenes):
    # closed matte box (kd = 0.5), every wall emits L = 0.5 two-sided: radiance = L / (1 - kd) = 1
    sc = OP.Scene(scenes.furnace_box(L=0.5, kd=0.5))
    cam = dict(pos=(0, 0, 0.0), look=(0, 0, 1), up=(0, 1, 0), fov=60.0, res=(24, 24))
    xyzw, _ = sc.render(cam, OP.film_desc((24, 24)), OP.path_desc(max_depth=40, rr_threshold=0.0, spp=64))
    assert abs(OP.resolve_rgb(xyzw).mean() - 1.0) < 0.01


def test_sample_ranges_add_up_exactly(OP, scenes):
    sc = OP.Scene(scenes.scene_c2())
    cam = dict(scenes.C2_CAMERA, res=(32, 32))
    fd = OP.film_desc((32, 32))
    kw = dict(max_depth=3, spp=8)
    whole, _ = sc.render(cam, fd, OP.path_desc(**kw))
    a, _ = sc.render(cam, fd, OP.path_desc(sample_begin=0, sample_end=8, **kw))
    assert np.array_equal(whole, a)
    assert (whole[..., 3] >= 8).all()                      # every pixel received its 8 samples (plus strays)
    # threads do not change the result (per-sample streams, ordered accumulation)
    one, _ = sc.render(cam, fd, OP.path_desc(**kw), threads=1)
    assert np.array_equal(whole, one)


def test_tile_order_and_sample_order_agree_statistically(OP, scenes):
    sc = OP.Scene(scenes.scene_c2())
    cam = dict(scenes.C2_CAMERA, res=(48, 48))
    fd = OP.film_desc((48, 48))
    kw = dict(max_depth=5, spp=64)
    a = OP.resolve_rgb(sc.render(cam, fd, OP.path_desc(**kw), mode=0)[0])
    b = OP.resolve_rgb(sc.render(cam, fd, OP.path_desc(**kw), mode=1)[0])
    assert abs(a.mean() - b.mean()) / b.mean() < 0.01
    assert 0.05 < b.mean() < 1.0 and not np.isnan(b).any()


def test_path_li_matches_render_accumulation(OP, scenes):
    sc = OP.Scene(scenes.scene_c2())
    cam = dict(scenes.C2_CAMERA, res=(16, 16))
    fd = OP.film_desc((16, 16))
    pd = OP.path_desc(max_depth=4, spp=4)
    xs, ys, ss = np.meshgrid(np.arange(16), np.arange(16), np.arange(4), indexing="ij")
    xy = np.stack([xs.ravel(), ys.ravel()], axis=1)
    L, pf = sc.path_li(cam, fd, pd, xy, ss.ravel())
    assert ((pf >= xy) & (pf <= xy + 1)).all()
    film = OP.film_add_samples(fd, pf, L, np.ones(len(L), np.float32))
    xyzw, _ = sc.render(cam, fd, pd)
    assert np.allclose(film, xyzw, rtol=1e-5, atol=1e-6)


def test_roughness_to_alpha_and_filter_table(OP):
    x = np.log(np.float32(0.1))
    want = 1.62142 + 0.819955 * x + 0.1734 * x * x + 0.0171201 * x ** 3 + 0.000640711 * x ** 4
    assert abs(float(OP.roughness_to_alpha(0.1)) - want) < 1e-6
    assert (OP.film_table(OP.film_desc((4, 4))) == 1.0).all()
    t = OP.film_table(OP.film_desc((4, 4), "gaussian", (2.0, 2.0), 2.0)).reshape(16, 16)
    assert t[0, 0] == t.max() and t[15, 15] >= 0 and np.allclose(t, t.T)


def test_sincos_contract_accuracy(OP):
    xs = np.linspace(-1.0, 7.0, 4001, dtype=np.float32)
    s = np.array([OP.sincos(float(x)) for x in xs])
    assert np.abs(s[:, 0] - np.sin(xs.astype(np.float64))).max() < 2.5e-7
    assert np.abs(s[:, 1] - np.cos(xs.astype(np.float64))).max() < 2.5e-7


def test_halton_known_answers(OP):
    """HaltonSampler restatement (samplers/halton.rs, lowdiscrepancy.rs): radical inverses of small indices, the pixel
    property of dimensions 0 / 1, and the permutation table being permutations."""
    # base 2 / base 3 radical inverses through dimension 0 / 1 of a 1x1 image (scales 1, stride 1 => index = sample number)
    idx, dims, perm = OP.halton_probe((1, 1), (0, 0), 0, n_dims=2, n_perm=2 + 3 + 5 + 7)
    assert idx == 0 and dims[0] == 0.0 and dims[1] == 0.0
    for s, (a, b) in {1: (0.5, 1 / 3), 2: (0.25, 2 / 3), 3: (0.75, 1 / 9), 4: (0.125, 4 / 9), 5: (0.625, 7 / 9)}.items():
        idx, dims, _ = OP.halton_probe((1, 1), (0, 0), s, n_dims=2)
        assert idx == s
        assert dims[0] == np.float32(a) and abs(dims[1] - b) < 1e-7
    assert sorted(perm[0:2]) == [0, 1] and sorted(perm[2:5]) == [0, 1, 2] and sorted(perm[5:10]) == list(range(5)) and sorted(perm[10:17]) == list(range(7))
    # the sample index of pixel (x, y) puts dimensions 0 / 1 of EVERY sample of that pixel inside the pixel's cell of the
    # 128 x 243 grid the sampler tiles the image with (halton.rs:117-141)
    for px, py in [(0, 0), (3, 5), (127, 200), (500, 17), (511, 511)]:
        for s in (0, 1, 7):
            idx, dims, _ = OP.halton_probe((512, 512), (px, py), s, n_dims=8)
            assert idx % (128 * 243) == OP.halton_probe((512, 512), (px, py), 0)[0] and idx // (128 * 243) == s
            assert ((dims >= 0) & (dims < 1)).all()

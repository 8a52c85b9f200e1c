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


def test_glass_in_a_furnace_preserves_radiance(OP, scenes):
    """External pin for FresnelSpecular + refract + the eta^2 radiance scaling (reflection.rs:733-819, :145-156): inside a closed
    box whose black walls all emit L = 1, a lossless dielectric (Kr = Kt = 1, eta = 1.5) changes nothing — every specular path
    carries throughput 1 when it leaves the glass again (F / F on reflection, (1 - F) eta_i^2 / eta_t^2 / (1 - F) in and the
    inverse out) and ends on an emitter, so every sample is 1 up to rounding, whether it sees the sphere or not; only paths still
    inside the glass at max_depth (total internal reflection) are lost."""
    box = scenes.furnace_box(L=1.0, kd=0.0)
    sv, si = scenes.uv_sphere(radius=0.45, center=(0.0, 0.0, 0.3), n_theta=24, n_phi=48)
    verts, idx = scenes.merge((box["verts"], box["idx"]), (sv, si))
    n_box = len(box["idx"])
    sc = dict(verts=verts, idx=idx, tri_material=np.concatenate([box["tri_material"], np.ones(len(si), np.uint32)]),
              materials=box["materials"] + [dict(type="glass", kr=(1, 1, 1), kt=(1, 1, 1), eta=1.5)], lights=box["lights"])
    assert all(l["prim"] < n_box for l in sc["lights"])
    cam = dict(pos=(0, 0, -0.9), look=(0, 0, 1), up=(0, 1, 0), fov=70.0, res=(32, 32))
    xyzw, _ = OP.Scene(sc).render(cam, OP.film_desc((32, 32)), OP.path_desc(max_depth=60, rr_threshold=0.0, spp=16))
    rgb = OP.resolve_rgb(xyzw)
    assert rgb.max() < 1.0 + 1e-4
    # every one of the 16 samples of a pixel returns exactly 1 (rounding aside) or — a grazing path caught in total internal
    # reflection until max_depth, on the sphere's silhouette only — 0
    assert np.abs(rgb * 16.0 - np.round(rgb * 16.0)).max() < 2e-3
    assert (np.abs(rgb - 1.0) < 1e-4).all(axis=2).mean() > 0.95
    assert (np.abs(rgb[12:20, 12:20] - 1.0) < 1e-4).all()   # straight through the middle of the sphere
    assert rgb.mean() > 0.995, rgb.mean()


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


def test_sobol_known_answers(OP):
    """SobolSampler restatement (samplers/sobol.rs, lowdiscrepancy.rs:507-560, the generator matrices of sobolmatrices.rs as
    converted by tools/make_sobol_tables.py).  Pins that do not come from the restatement: dimension 0 of the Sobol' sequence
    is the base-2 radical inverse, dimension 1 starts 0, 1/2, 3/4, 1/4, 5/8, 1/8, 3/8, 7/8 (Bratley & Fox), the first two
    dimensions form a (0, 2)-sequence, and sobol_interval_to_index must put dimensions 0 / 1 of EVERY sample of a pixel inside
    that pixel — which only holds if the van der Corput matrices and their inverses are the right ones."""
    want1 = [0.0, 0.5, 0.75, 0.25, 0.625, 0.125, 0.375, 0.875]
    for i in range(8):
        _, d = OP.sobol_probe((0, 0, 16, 16), (0, 0), i, n_dims=2, raw=True)
        assert d[0] == np.float32(int(f"{i:03b}"[::-1], 2) / 8.0) and d[1] == np.float32(want1[i])
    # (0, 2)-sequence: among the first 2^k points every elementary interval 2^-a x 2^-(k-a) holds exactly one
    k = 6
    pts = np.array([OP.sobol_probe((0, 0, 16, 16), (0, 0), i, n_dims=2, raw=True)[1] for i in range(1 << k)], np.float64)
    for a in range(k + 1):
        cells = (np.floor(pts[:, 0] * (1 << a)).astype(int) << (k - a)) | np.floor(pts[:, 1] * (1 << (k - a))).astype(int)
        assert sorted(cells) == list(range(1 << k))
    # higher dimensions are low-discrepancy too: 1024 points of dimensions (7, 200) fill a 16 x 16 grid nearly evenly
    pts = np.array([OP.sobol_probe((0, 0, 16, 16), (0, 0), i, n_dims=201, raw=True)[1][[7, 200]] for i in range(1024)], np.float64)
    counts = np.bincount((np.floor(pts[:, 0] * 16).astype(int) * 16 + np.floor(pts[:, 1] * 16).astype(int)), minlength=256)
    assert counts.min() >= 2 and counts.max() <= 6
    # pixel property, with sample bounds that neither start at 0 nor are a power of two (resolution rounds up to 256)
    sb = (-2, -3, 200, 131)
    for px, py in [(-2, -3), (0, 0), (57, 100), (197, 127), (120, 5)]:
        for s in (0, 1, 2, 3, 17, 255):
            idx, dims = OP.sobol_probe(sb, (px, py), s, n_dims=6)
            assert idx >> 16 == s                                   # index = sample number << 2 log2(resolution), low bits locate the pixel
            _, raw = OP.sobol_probe(sb, (px, py), idx, n_dims=2, raw=True)
            pos = raw.astype(np.float64) * 256 + np.array(sb[:2])
            assert np.floor(pos[0]) == px and np.floor(pos[1]) == py
            assert 0 <= dims[0] < 1 and 0 <= dims[1] < 1 and np.allclose(dims[:2], pos - (px, py), atol=3e-5)
            assert ((dims >= 0) & (dims < 1)).all()
    # different pixels get different indices; consecutive samples of one pixel stay apart in every dimension
    a = OP.sobol_probe(sb, (10, 10), 0, n_dims=8)[1]
    b = OP.sobol_probe(sb, (10, 10), 1, n_dims=8)[1]
    assert (np.abs(a[2:] - b[2:]) > 1e-3).all()


def test_sobol_render_converges_like_the_other_samplers(OP, scenes):
    """A Cornell box rendered with the SobolSampler has the RandomSampler's mean (unbiased) and, at equal spp, a lower error
    against a 256-spp render — what a low-discrepancy sampler is for."""
    sc = scenes.scene_c2()
    cam = dict(scenes.C2_CAMERA, res=(40, 40))
    fd = OP.film_desc(cam["res"])
    ref = OP.Scene(sc, 4)
    kw = dict(max_depth=3, rr_threshold=1.0, light_strategy="uniform")
    truth = OP.resolve_rgb(ref.render(cam, fd, OP.path_desc(spp=256, **kw))[0]).astype(np.float64)
    err = {}
    for sampler in ("random", "sobol"):
        img = OP.resolve_rgb(ref.render(cam, fd, OP.path_desc(spp=16, sampler=sampler, **kw))[0]).astype(np.float64)
        err[sampler] = float(((img - truth) ** 2).mean())
        assert abs(img.mean() / truth.mean() - 1.0) < 0.03
    assert err["sobol"] < err["random"]


def test_spot_and_distant_lights(OP, scenes):
    """SpotLight (src/lights/spot.rs) and DistantLight (src/lights/distant.rs) on a single matte floor quad, against the closed
    forms: directly below a spot of intensity I at height h the radiance is kd/pi * I / h^2, zero outside the cone; under a
    distant light of radiance L from direction w it is kd/pi * L * cos(theta) everywhere."""
    quad = np.array([(-50, 0, -50), (-50, 0, 50), (50, 0, 50), (50, 0, -50)], np.float32)
    base = dict(verts=quad, idx=np.array([[0, 1, 2], [0, 2, 3]], np.uint32), tri_material=np.zeros(2, np.uint32),
                materials=[dict(type="matte", kd=(0.5, 0.5, 0.5))])
    cam = dict(pos=(0, 40.0, 0), look=(0, 0, 0), up=(0, 0, 1), fov=90.0, res=(33, 33))
    fd = OP.film_desc(cam["res"])
    pd = OP.path_desc(max_depth=1, spp=4)
    spot = dict(type="spot", p=(0, 10.0, 0), axis=(0, -1.0, 0), I=(200.0, 200.0, 200.0), total_width=30.0, falloff_start=20.0)
    img = OP.resolve_rgb(OP.Scene(dict(base, lights=[spot])).render(cam, fd, pd)[0])
    # centre pixel (2.4 units wide on the floor): below the light, inside falloff_start; cos^3 falloff over the pixel ~ -1.5 %
    assert 0.975 < img[16, 16, 0] / (0.5 / np.pi * 200.0 / 100.0) < 1.0
    assert img[0, 0].max() == 0.0 and img[16, 0].max() == 0.0                   # 40 units out: far outside the 30 degree cone
    r = np.hypot(*np.meshgrid(np.arange(33) - 16.0, np.arange(33) - 16.0))     # lit disc radius = 10 tan(30 deg) = 5.77 of 80/33 per pixel
    lit = img[..., 0] > 0
    assert lit[r < 1.5].all() and not lit[r > 3.5].any()
    w = np.array((0.0, 0.6, 0.8))
    distant = dict(type="distant", w=tuple(w), L=(3.0, 2.0, 1.0))
    img = OP.resolve_rgb(OP.Scene(dict(base, lights=[distant])).render(cam, fd, pd)[0])
    assert np.allclose(img, 0.5 / np.pi * np.array((3.0, 2.0, 1.0)) * 0.6, rtol=1e-5)
    # power heuristic of the light distribution: spot 2 pi I (1 - (cos 20 + cos 30) / 2), distant pi r^2 L (both through luminance)
    both = OP.Scene(dict(base, lights=[spot, distant]))
    a = OP.resolve_rgb(both.render(cam, fd, OP.path_desc(max_depth=1, spp=256, light_strategy="power"))[0])
    b = OP.resolve_rgb(both.render(cam, fd, OP.path_desc(max_depth=1, spp=256, light_strategy="uniform"))[0])
    assert abs(a.mean() - b.mean()) / b.mean() < 0.03


def test_pixel_sampler_tables(OP):
    """StratifiedSampler / ZeroTwoSequenceSampler start_pixel (samplers/stratified.rs:44-76, samplers/zerotwosequence.rs:31-48):
    every 1D table holds one value per stratum, every stratified 2D table one point per (x, y) cell, every (0,2) 2D table is a
    (0,2)-net: one point in each elementary interval of area 1/spp."""
    pd = OP.path_desc(spp=64, sampler="stratified", n_sampled_dimensions=3, x_samples=16, y_samples=4)
    t1, t2 = OP.pixel_tables(pd, 7)
    for d in range(3):
        assert sorted(np.floor(t1[d] * 64).astype(int)) == list(range(64))
        cells = np.floor(t2[d, :, 0] * 16).astype(int) * 4 + np.floor(t2[d, :, 1] * 4).astype(int)
        assert sorted(cells) == list(range(64))
        assert list(np.floor(t1[d] * 64).astype(int)) != list(range(64))          # shuffled
    nj, _ = OP.pixel_tables(OP.path_desc(spp=64, sampler="stratified", n_sampled_dimensions=1, x_samples=8, y_samples=8, jitter=False), 7)
    assert sorted(nj[0]) == [np.float32((i + 0.5) / 64) for i in range(64)]
    pd = OP.path_desc(spp=64, sampler="zerotwo", n_sampled_dimensions=3)
    t1, t2 = OP.pixel_tables(pd, 7)
    for d in range(3):
        assert sorted(np.floor(t1[d] * 64).astype(int)) == list(range(64))        # scrambled van der Corput: stratified at every power of two
        for k in range(7):                                                        # 2^k x 2^(6-k) boxes
            a, b = 1 << k, 1 << (6 - k)
            cells = np.floor(t2[d, :, 0] * a).astype(int) * b + np.floor(t2[d, :, 1] * b).astype(int)
            assert sorted(cells) == list(range(64)), (d, k)
    u1, u2 = OP.pixel_tables(pd, 8)
    assert not np.array_equal(t1, u1) and not np.array_equal(t2, u2)              # another pixel, another scramble


@pytest.mark.parametrize("sampler,kw", [("stratified", dict(x_samples=8, y_samples=8)), ("zerotwo", {})])
def test_pixel_samplers_render_and_converge(OP, scenes, sampler, kw):
    """Cornell box with the PixelSamplers: same mean as the RandomSampler render (unbiased), lower pixel variance at equal spp
    on the directly lit floor, and the reference's tile-sequential order (mode 0) agrees statistically with per-pixel streams."""
    sc = OP.Scene(scenes.scene_c2())
    cam = dict(scenes.C2_CAMERA, res=(48, 48))
    fd = OP.film_desc((48, 48))
    base = dict(max_depth=5, spp=64)
    rnd = OP.resolve_rgb(sc.render(cam, fd, OP.path_desc(**base))[0])
    a = OP.resolve_rgb(sc.render(cam, fd, OP.path_desc(sampler=sampler, **base, **kw), mode=1)[0])
    b = OP.resolve_rgb(sc.render(cam, fd, OP.path_desc(sampler=sampler, **base, **kw), mode=0)[0])
    assert abs(a.mean() - rnd.mean()) / rnd.mean() < 0.02
    assert abs(a.mean() - b.mean()) / b.mean() < 0.02
    ref = OP.resolve_rgb(sc.render(cam, fd, OP.path_desc(max_depth=5, spp=1024))[0])
    err = lambda img: float(np.mean((img - ref) ** 2))
    assert err(a) < err(rnd)
    # sample ranges add up exactly with per-pixel tables too
    half, _ = sc.render(cam, fd, OP.path_desc(sampler=sampler, sample_begin=0, sample_end=40, **base, **kw))
    whole, _ = sc.render(cam, fd, OP.path_desc(sampler=sampler, sample_begin=40, sample_end=64, **base, **kw), out=half)
    full, _ = sc.render(cam, fd, OP.path_desc(sampler=sampler, **base, **kw))
    assert np.allclose(whole, full, rtol=1e-5, atol=1e-6)


def test_film_crop_window_clamp_and_filters(OP, scenes):
    """Film::new (film.rs:31-75): cropped_pixel_bounds = ceil(res * crop), sample bounds around it; a cropped render equals the
    same pixels of the full render for the box filter (a pixel only sees its own samples); max_sample_luminance rescales a
    sample to that luminance (film.rs:259-261); Triangle / Mitchell / LanczosSinc tables (src/filters/*.rs) against closed forms."""
    fd = OP.film_desc((40, 30), crop=(0.25, 0.2, 0.8, 0.9))
    pb, sb = OP.film_bounds(fd)
    assert pb == (10, 6, 32, 27) and sb == (10, 6, 32, 27)
    pb, sb = OP.film_bounds(OP.film_desc((40, 30), "gaussian", (2.0, 2.0), crop=(0.25, 0.2, 0.8, 0.9)))
    assert pb == (10, 6, 32, 27) and sb == (8, 4, 34, 29)
    assert OP.film_bounds(OP.film_desc((40, 30)))[0] == (0, 0, 40, 30)
    sc = OP.Scene(scenes.scene_c2())
    cam = dict(scenes.C2_CAMERA, res=(40, 30))
    pd = OP.path_desc(max_depth=3, spp=4)
    full, _ = sc.render(cam, OP.film_desc((40, 30)), pd)
    crop, _ = sc.render(cam, fd, pd)
    assert crop.shape == (21, 22, 4)
    # per-(pixel, sample) streams are indexed inside the sample bounds, so the crop draws different numbers: same estimator
    assert abs(OP.resolve_rgb(crop).mean() - OP.resolve_rgb(full[6:27, 10:32]).mean()) / OP.resolve_rgb(crop).mean() < 0.15
    # luminance clamp on explicit samples
    L = np.array([[10.0, 10.0, 10.0], [0.1, 0.2, 0.3]], np.float32)
    pf = np.array([[1.5, 1.5], [2.5, 1.5]], np.float32)
    got = OP.film_add_samples(OP.film_desc((4, 4), max_sample_luminance=2.0), pf, L, np.ones(2, np.float32))
    rgb = OP.resolve_rgb(got)
    assert np.allclose(rgb[1, 1], 2.0, rtol=1e-5) and np.allclose(rgb[1, 2], L[1], rtol=1e-5)
    # filter tables
    t = OP.film_table(OP.film_desc((4, 4), "triangle", (2.0, 1.0))).reshape(16, 16)
    x = (np.arange(16) + 0.5) * 2.0 / 16
    y = (np.arange(16) + 0.5) * 1.0 / 16
    assert np.allclose(t, np.outer(1.0 - y, 2.0 - x), rtol=1e-6)
    t = OP.film_table(OP.film_desc((4, 4), "mitchell", (2.0, 2.0), b=1 / 3, c=1 / 3)).reshape(16, 16)
    assert t[0, 0] == t.max() and t.min() < 0 and np.allclose(t, t.T)                 # Mitchell has negative lobes
    t = OP.film_table(OP.film_desc((4, 4), "sinc", (4.0, 4.0), tau=3.0)).reshape(16, 16)
    xs = (np.arange(16) + 0.5) * 4.0 / 16
    w = np.sinc(xs) * np.sinc(xs / 3.0)
    assert np.allclose(t, np.outer(w, w), atol=2e-6)


def test_mesh_shading_normals_uvs_tangents(OP, scenes):
    """TriangleMesh's optional n / s / uv (triangle.rs:17-26, 60-72, 251-311): analytic vertex normals make a coarsely tessellated
    matte sphere shade smoothly (neighbouring pixels differ less than on the faceted mesh), UVs / tangents only rotate the
    tangent frame of isotropic BSDFs (same image statistically), and nothing turns into NaN."""
    v, i = scenes.uv_sphere(radius=1.0, n_theta=10, n_phi=20)
    base = dict(verts=v, idx=i, tri_material=np.zeros(len(i), np.uint32), materials=[dict(type="matte", kd=(0.8, 0.8, 0.8))],
                lights=[dict(type="distant", w=(0.3, 0.5, -1.0), L=(3.0, 3.0, 3.0))])
    cam = dict(pos=(0, 0, -4.0), look=(0, 0, 0), up=(0, 1, 0), fov=30.0, res=(64, 64))
    fd = OP.film_desc(cam["res"])
    pd = OP.path_desc(max_depth=1, spp=16)
    flat = OP.resolve_rgb(OP.Scene(base).render(cam, fd, pd)[0])[..., 0]
    smooth = OP.resolve_rgb(OP.Scene(dict(base, normals=v.copy())).render(cam, fd, pd)[0])[..., 0]
    assert not np.isnan(smooth).any() and abs(smooth.mean() - flat.mean()) / flat.mean() < 0.05
    # analytic: Lambert under a distant light, radiance = kd/pi * L * max(0, n.w) with n = the unit position on the sphere
    w = np.array((0.3, 0.5, -1.0)) / np.linalg.norm((0.3, 0.5, -1.0))
    ys, xs = np.mgrid[20:44, 20:44]                       # central pixels, well inside the silhouette
    tanh = np.tan(np.radians(15.0))
    d = np.stack([(xs + 0.5 - 32) / 32 * tanh, -(ys + 0.5 - 32) / 32 * tanh, np.ones_like(xs, dtype=float)], axis=-1)   # raster y points down
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.array((0, 0, -4.0))
    b = (d @ o)
    t = -b - np.sqrt(b * b - (o @ o - 1.0))
    n = o + d * t[..., None]
    want = 0.8 / np.pi * 3.0 * np.maximum(0.0, n @ w)
    err_smooth = np.abs(smooth[20:44, 20:44] - want).mean()
    err_flat = np.abs(flat[20:44, 20:44] - want).mean()
    assert err_smooth < 0.2 * err_flat and err_smooth < 0.004, (err_smooth, err_flat)
    # UVs / tangents: the frame changes, the estimator does not
    sc4 = scenes.scene_c4(n_theta=12, n_phi=24)
    cam4 = dict(scenes.C4_CAMERA, res=(48, 27))
    fd4, pd4 = OP.film_desc(cam4["res"]), OP.path_desc(max_depth=4, spp=64, light_strategy="power")
    a = OP.resolve_rgb(OP.Scene(sc4).render(cam4, fd4, pd4)[0])
    full = scenes.scene_c4_smooth(n_theta=12, n_phi=24, tangents=True)
    bimg = OP.resolve_rgb(OP.Scene(dict(sc4, uvs=full["uvs"])).render(cam4, fd4, pd4)[0])
    assert not np.array_equal(a, bimg) and abs(a.mean() - bimg.mean()) / a.mean() < 0.03
    cimg = OP.resolve_rgb(OP.Scene(full).render(cam4, fd4, pd4)[0])
    assert not np.isnan(cimg).any() and cimg.mean() > 0.01


def test_mirror_metal_and_oren_nayar(OP, scenes):
    """pbrt-v3 mirror / metal / matte-with-sigma over the reference's BxDF blocks (SpecularReflection + FresnelNoOp, MicrofacetReflection +
    FresnelConductor, OrenNayar): a mirror shows the light at Kr x L_e; a smooth conductor at normal incidence at
    F0 = ((eta-1)^2 + k^2) / ((eta+1)^2 + k^2) x L_e; Oren-Nayar with sigma = 0 renders exactly like the Lambertian, with sigma > 0 it
    flattens a sphere lit from the camera (limb brighter relative to the centre)."""
    # a reflector quad at z = 0 facing -z, an emitter quad at z = -10 facing +z, the camera just in front of the emitter looking at the reflector
    refl = np.array([(-4, -4, 0), (4, -4, 0), (4, 4, 0), (-4, 4, 0)], np.float32)
    emit = np.array([(-3, -3, -10), (-3, 3, -10), (3, 3, -10), (3, -3, -10)], np.float32)
    verts = np.concatenate([refl, emit])
    idx = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.uint32)
    lights = [dict(type="area", prim=2, L=(2.0, 3.0, 4.0), two_sided=True), dict(type="area", prim=3, L=(2.0, 3.0, 4.0), two_sided=True)]
    cam = dict(pos=(0.3, 0.2, -9.0), look=(0.0, 0.0, 0.0), up=(0, 1, 0), fov=10.0, res=(8, 8))
    fd = OP.film_desc(cam["res"])

    def centre(mat, **pk):
        sc = OP.Scene(dict(verts=verts, idx=idx, tri_material=np.array([1, 1, 0, 0], np.uint32), materials=[dict(type="matte", kd=(0, 0, 0)), mat], lights=lights))
        img = OP.resolve_rgb(sc.render(cam, fd, OP.path_desc(max_depth=2, spp=pk.get("spp", 4)))[0])
        return img[3:5, 3:5].mean(axis=(0, 1))

    got = centre(dict(type="mirror", kr=(0.9, 0.8, 0.7)))
    assert np.allclose(got, np.array((2.0, 3.0, 4.0)) * (0.9, 0.8, 0.7), rtol=1e-5)
    eta, k = np.array((0.2, 0.92, 1.1)), np.array((3.9, 2.45, 2.14))
    f0 = ((eta - 1) ** 2 + k ** 2) / ((eta + 1) ** 2 + k ** 2)
    got = centre(dict(type="metal", metal_eta=tuple(eta), metal_k=tuple(k), roughness=0.002, remap=False), spp=4096)
    assert np.allclose(got, np.array((2.0, 3.0, 4.0)) * f0, rtol=0.05), (got, np.array((2.0, 3.0, 4.0)) * f0)
    # Oren-Nayar
    v, i = scenes.uv_sphere(radius=1.0, n_theta=48, n_phi=96)
    base = dict(verts=v, idx=i, tri_material=np.zeros(len(i), np.uint32), lights=[dict(type="distant", w=(0.0, 0.0, -1.0), L=(3.0, 3.0, 3.0))],
                normals=v.copy())
    cam = dict(pos=(0, 0, -6.0), look=(0, 0, 0), up=(0, 1, 0), fov=22.0, res=(64, 64))
    fd = OP.film_desc(cam["res"])
    pd = OP.path_desc(max_depth=1, spp=8)
    lam = OP.resolve_rgb(OP.Scene(dict(base, materials=[dict(type="matte", kd=(0.8, 0.8, 0.8))])).render(cam, fd, pd)[0])[..., 0]
    on = OP.resolve_rgb(OP.Scene(dict(base, materials=[dict(type="matte", kd=(0.8, 0.8, 0.8), sigma=40.0)])).render(cam, fd, pd)[0])[..., 0]
    assert np.array_equal(lam, OP.resolve_rgb(OP.Scene(dict(base, materials=[dict(type="matte", kd=(0.8, 0.8, 0.8), sigma=0.0)])).render(cam, fd, pd)[0])[..., 0])
    centre_ratio = on[30:34, 30:34].mean() / lam[30:34, 30:34].mean()
    limb_ratio = on[32, 8:12].mean() / lam[32, 8:12].mean()
    assert centre_ratio < 0.95 and limb_ratio > 1.15 * centre_ratio, (centre_ratio, limb_ratio)
    # everything together: finite, plausible
    img = OP.resolve_rgb(OP.Scene(scenes.scene_materials(12, 24)).render(dict(scenes.C4_CAMERA, res=(48, 27)), OP.film_desc((48, 27)),
                                                                       OP.path_desc(max_depth=6, spp=16, light_strategy="power"))[0])
    assert not np.isnan(img).any() and 0.01 < img.mean() < 2.0


def test_rough_glass(OP, scenes):
    """pbrt-v3 GlassMaterial with roughness: MicrofacetReflection + MicrofacetTransmission (reflection.rs:1058-1192, D62).  In a uniform
    radiance field (closed emissive box, L = 1) a non-absorbing dielectric sphere must stay close to radiance 1 (the single-scattering
    microfacet model is energy-conserving only to a few per cent: 0.98-1.02 over roughness 0.02-0.3 at 256 spp); the smooth
    FresnelSpecular glass gives exactly 1, and a nearly smooth rough glass is close to it."""
    box = scenes.furnace_box(L=1.0, kd=0.0)
    v, i = scenes.uv_sphere(radius=0.5, n_theta=24, n_phi=48)
    verts, idx = scenes.merge((box["verts"], box["idx"]), (v, i))
    cam = dict(pos=(0, 0, -0.95), look=(0, 0, 0), up=(0, 1, 0), fov=60.0, res=(32, 32))
    fd = OP.film_desc(cam["res"])

    def render(glass, spp=64, normals=True):
        n = np.concatenate([np.zeros_like(box["verts"]), v / 0.5]) if normals else None
        sc = dict(verts=verts, idx=idx, tri_material=np.concatenate([box["tri_material"], np.ones(len(i), np.uint32)]),
                  materials=[dict(type="matte", kd=(0.0, 0.0, 0.0)), glass], lights=box["lights"])
        return OP.resolve_rgb(OP.Scene(sc).render(cam, fd, OP.path_desc(max_depth=12, rr_threshold=0.0, spp=spp))[0])

    rough = render(dict(type="glass", kr=(1, 1, 1), kt=(1, 1, 1), eta=1.5, roughness=0.15, remap=False))
    assert not np.isnan(rough).any()
    centre = rough[10:22, 10:22].mean()
    assert 0.93 < centre < 1.07, centre
    smooth = render(dict(type="glass", kr=(1, 1, 1), kt=(1, 1, 1), eta=1.5))
    assert abs(smooth[10:22, 10:22].mean() - 1.0) < 0.02
    nearly = render(dict(type="glass", kr=(1, 1, 1), kt=(1, 1, 1), eta=1.5, roughness=0.002, remap=False))
    assert abs(nearly[10:22, 10:22].mean() - smooth[10:22, 10:22].mean()) < 0.06


def test_spatial_light_distribution(OP):
    """SpatialLightDistribution (lightdistrib.rs:71-220).  Grid: 64 voxels along the longest axis of the world bound, the others
    in proportion (:83-95).  One voxel's distribution, recomputed here in f32 numpy for two point lights over the 128 Halton
    points (bases 2, 3, 5: known values 1/2, 1/3, 1/5 at index 1), equals the oracle's bits; the 0.1 % floor (:163-170) holds;
    a voxel next to one light prefers it; the rendered mean agrees with the "uniform" strategy (the estimator stays unbiased)."""
    f32 = np.float32
    quad = np.array([(-50, 0, -50), (-50, 0, 50), (50, 0, 50), (50, 0, -50), (-50, 25, -50)], np.float32)
    idx = np.array([[0, 1, 2], [0, 2, 3], [0, 1, 4]], np.uint32)
    la = dict(type="point", p=(-40.0, 5.0, -40.0), I=(100.0, 100.0, 100.0))
    lb = dict(type="point", p=(40.0, 5.0, 40.0), I=(300.0, 200.0, 100.0))
    sc = dict(verts=quad, idx=idx, tri_material=np.zeros(3, np.uint32), materials=[dict(type="matte", kd=(0.5, 0.5, 0.5))], lights=[la, lb])
    s = OP.Scene(sc)
    assert s.spatial_grid() == (64, 16, 64)                                   # extents 100 x 25 x 100
    assert s.spatial_voxel_of((-50.0, 0.0, -50.0)) == (0, 0, 0) and s.spatial_voxel_of((50.0, 25.0, 50.0)) == (63, 15, 63)
    assert s.spatial_voxel_of((1e9, -1e9, float("nan"))) == (63, 0, 0)        # `as i32` saturates, NaN -> 0, then clamp

    def radical_inverse(base, i):
        inv, inv_n, rev = f32(1.0) / f32(base), f32(1.0), 0
        while i:
            i, d = divmod(i, base)
            rev = rev * base + d
            inv_n = f32(inv_n * inv)
        return min(f32(f32(rev) * inv_n), f32(1.0) - f32(2.0 ** -23))
    assert [float(radical_inverse(b, 1)) for b in (2, 3, 5)] == [0.5, float(f32(1) / f32(3)), float(f32(1) / f32(5))]
    pi = (5, 3, 60)
    lo, hi = np.array((-50, 0, -50), f32), np.array((50, 25, 50), f32)
    nv = np.array((64, 16, 64), f32)
    lerp = lambda t, a, b: (f32(1) - t) * a + t * b
    v0, v1 = lerp(np.array(pi, f32) / nv, lo, hi), lerp((np.array(pi, f32) + f32(1)) / nv, lo, hi)
    contrib = np.zeros(2, f32)
    for i in range(128):
        t = np.array([radical_inverse(b, i) for b in (2, 3, 5)], f32)
        po = lerp(t, np.minimum(v0, v1), np.maximum(v0, v1))
        for j, l in enumerate((la, lb)):
            d = np.array(l["p"], f32) - po
            d2 = f32(f32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
            li = np.array(l["I"], f32) / d2
            contrib[j] = contrib[j] + f32(f32(f32(0.212671) * li[0] + f32(0.715160) * li[1]) + f32(0.072169) * li[2])
    func, cdf, func_int = s.spatial_voxel(pi, 2)
    assert np.array_equal(func.view(np.uint32), contrib.view(np.uint32))
    assert cdf[0] == 0.0 and cdf[2] == 1.0 and np.isclose(cdf[1], contrib[0] / contrib.sum(), rtol=1e-6)
    assert np.isclose(func_int, contrib.sum() / 2, rtol=1e-6)
    near_a, _, _ = s.spatial_voxel(s.spatial_voxel_of(la["p"]), 2)
    near_b, _, _ = s.spatial_voxel(s.spatial_voxel_of(lb["p"]), 2)
    assert near_a[0] > 20 * near_a[1] and near_b[1] > 20 * near_b[0]
    # the floor: a spot light that cannot see the voxel still keeps 0.1 % of the voxel's average contribution
    spot = dict(type="spot", p=(0, 10.0, 0), axis=(0, -1.0, 0), I=(200.0, 200.0, 200.0), total_width=10.0, falloff_start=5.0)
    s2 = OP.Scene(dict(sc, lights=[la, spot]))
    s2.spatial_grid()
    f2, c2, _ = s2.spatial_voxel((60, 10, 60), 2)
    assert f2[1] > 0 and np.isclose(f2[1], 0.001 * f2[0] / (128 * 2), rtol=1e-5)      # min_contrib = 0.001 * sum / (n_samples * n_lights)
    # one light: "spatial" falls back to the uniform distribution (:223); several: same expectation as "uniform"
    cam = dict(pos=(0, 40.0, 0), look=(0, 0, 0), up=(0, 0, 1), fov=90.0, res=(32, 32))
    fd = OP.film_desc(cam["res"])
    one = OP.Scene(dict(sc, lights=[la]))
    a = one.render(cam, fd, OP.path_desc(max_depth=2, spp=4, light_strategy="spatial"))[0]
    b = one.render(cam, fd, OP.path_desc(max_depth=2, spp=4, light_strategy="uniform"))[0]
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    a = OP.resolve_rgb(s.render(cam, fd, OP.path_desc(max_depth=2, spp=256, light_strategy="spatial"))[0])
    b = OP.resolve_rgb(s.render(cam, fd, OP.path_desc(max_depth=2, spp=256, light_strategy="uniform"))[0])
    assert abs(a.mean() - b.mean()) / b.mean() < 0.02
    # importance sampling the nearer light lowers the variance under it
    assert a[2:8, 2:8].std() < b[2:8, 2:8].std()


def test_film_splats_and_write_image_splat_term(OP):
    """Film::add_splat (film.rs:137-151, D64 FIX) and the splat term of Film::write_image (:167-172): a splat lands on
    floor(p) inside the cropped bounds only, is clamped to max_sample_luminance, accumulates as XYZ; the written pixel is
    max(rgb / w, 0) + splat_scale * xyz_to_rgb(splat), times scale."""
    fd = OP.film_desc((8, 6), crop=(0.25, 0.0, 1.0, 1.0), max_sample_luminance=2.0)
    assert OP.film_bounds(fd)[0] == (2, 0, 8, 6)
    p = np.array([(2.0, 0.0), (7.99, 5.99), (1.99, 3.0), (8.0, 2.0), (4.5, -0.01), (4.5, 2.5), (4.2, 2.9)], np.float32)
    v = np.array([(1, 1, 1), (0.5, 0.25, 0.125), (9, 9, 9), (9, 9, 9), (9, 9, 9), (10, 10, 10), (0.1, 0.2, 0.3)], np.float32)
    sp = OP.film_add_splats(fd, p, v)
    assert sp.shape == (6, 6, 3)
    hit = np.zeros((6, 6), bool)
    hit[0, 0] = hit[5, 5] = hit[2, 2] = True                 # columns are x - 2
    assert ((np.abs(sp).sum(axis=2) > 0) == hit).all()       # the three splats outside the cropped bounds are dropped
    assert np.array_equal(sp[0, 0], OP.rgb_to_xyz(np.array([1, 1, 1], np.float32)))
    clamped = np.float32(10.0) * (np.float32(2.0) / OP.rgb_to_xyz(np.array([[10, 10, 10]], np.float32))[0, 1])      # y of grey 10 is 10
    want = OP.rgb_to_xyz(np.array([clamped] * 3, np.float32)) + OP.rgb_to_xyz(np.array([0.1, 0.2, 0.3], np.float32))
    assert np.allclose(sp[2, 2], want, rtol=1e-6)
    xyzw = np.zeros((6, 6, 4), np.float32)
    xyzw[..., :3] = OP.rgb_to_xyz(np.full((6, 6, 3), 0.5, np.float32)) * np.float32(4.0)
    xyzw[..., 3] = 4.0
    out = OP.resolve_rgb_splat(xyzw, sp, scale=2.0, splat_scale=0.5)
    assert np.allclose(out[1, 1], 1.0, rtol=1e-5)                                        # no splat: (0.5 + 0) * 2
    assert np.allclose(out[0, 0], (0.5 + 0.5 * 1.0) * 2.0, rtol=1e-5)
    assert np.array_equal(OP.resolve_rgb_splat(xyzw, np.zeros_like(sp), 2.0, 0.5), OP.resolve_rgb(xyzw, 2.0))

"""CPU tests of the oracle (the restatement of the reference hot path) — pins, hand cases and self-consistency
properties.  The reference's own tests hold no golden vector for this path (parity unpinned); the
externally pinned piece is PCG32 (canonical PCG demo vector)."""
import numpy as np
import pytest


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def test_pcg32_canonical_vector(orc):
    # pcg32 demo: seed state 42, stream 54 (https://www.pcg-random.org) — the one external known-answer test
    got = orc.pcg32_u32(54, 6, init_state=42)
    assert [int(x) for x in got] == [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]


def test_pcg32_reference_set_sequence(orc):
    # src/core/rng.rs:21-27 with PCG32_DEFAULT_STATE
    assert [int(x) for x in orc.pcg32_u32(0, 4)] == [0x69C87837, 0x6694BD1C, 0xA37B7AC6, 0x572D246C]
    assert [int(x) for x in orc.pcg32_u32(1, 4)] == [0x73C29FDB, 0xFBAA1FF7, 0xDB022AF6, 0x12D7398C]


def test_uniform_float_is_u32_times_2pow_minus32_clamped(orc):
    u = orc.pcg32_u32(7, 1000)
    f = orc.pcg32_float(7, 1000)
    expect = np.minimum(np.float32(1.0) - np.float32(2.0 ** -23), u.astype(np.float32) * np.float32(2.3283064365386963e-10))
    assert np.array_equal(bits(f), bits(expect))
    assert f.max() < 1.0


def test_gamma_and_slab_widen(orc):
    assert bits(orc.gamma(3.0)) == bits(np.float32(1.7881396e-07))
    assert int(bits(np.float32(orc.lib().orc_slab_widen()))) == 0x3F800003
    eps = np.float32(2.0 ** -24)
    for n in (2, 3, 5, 7):
        n32 = np.float32(n)
        assert bits(orc.gamma(n)) == bits((n32 * eps) / (np.float32(1) - n32 * eps))


def test_next_float_up_down(orc):
    assert bits(orc.next_float_up(0.0)) == 1
    assert bits(orc.next_float_up(-0.0)) == 1
    assert bits(orc.next_float_down(0.0)) == 0x80000001
    assert bits(orc.next_float_up(1.0)) == 0x3F800001
    assert bits(orc.next_float_down(1.0)) == 0x3F7FFFFF
    assert bits(orc.next_float_up(-1.0)) == 0xBF7FFFFF
    assert np.isinf(orc.next_float_up(np.inf)) and np.isinf(orc.next_float_down(-np.inf))


def test_slab_hand_cases(orc):
    box = [0, 0, 0, 1, 1, 1]
    ok, t = orc.slab_test(box, orc.make_ray([0.5, 0.5, -1], [0, 0, 1]))
    assert ok and t == 1.0
    assert not orc.slab_test(box, orc.make_ray([0.5, 0.5, -1], [0, 0, -1]))[0]          # behind
    assert not orc.slab_test(box, orc.make_ray([0.5, 0.5, -1], [0, 0, 1], t_max=0.5))[0]  # beyond t_max
    assert not orc.slab_test(box, orc.make_ray([2.0, 0.5, -1], [0, 0, 1]))[0]            # parallel miss, inv = inf
    assert orc.slab_test(box, orc.make_ray([0.5, 0.5, 0.5], [1, 0, 0]))[0]               # origin inside
    # d.x = -0.0 -> inv = -inf -> dir_is_neg = 1 (Appendix D)
    assert orc.slab_test(box, orc.make_ray([0.5, 0.5, -1], [-0.0, 0, 1]))[0]
    # origin exactly on a slab plane with a zero direction component: 0 * inf = NaN falls through
    ok, _ = orc.slab_test(box, orc.make_ray([0.0, 0.5, -1], [0, 0, 1]))
    assert ok in (True, False)


def test_triangle_hand_cases(orc):
    tri = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    hit, out = orc.triangle_test(tri, orc.make_ray([0.25, 0.25, -1], [0, 0, 1]))
    assert hit and out[3] == 1.0 and np.allclose(out[:3], [0.5, 0.25, 0.25])
    # both facings are hit (D8 FIX)
    hit2, out2 = orc.triangle_test([0, 0, 0, 0, 1, 0, 1, 0, 0], orc.make_ray([0.25, 0.25, -1], [0, 0, 1]))
    assert hit2 and out2[3] == 1.0
    assert not orc.triangle_test(tri, orc.make_ray([0.25, 0.25, 1], [0, 0, 1]))[0]            # behind origin
    assert not orc.triangle_test(tri, orc.make_ray([0.25, 0.25, -1], [0, 0, 1], t_max=0.5))[0]  # beyond t_max
    assert orc.triangle_test(tri, orc.make_ray([0.25, 0.25, -1], [0, 0, 1], t_max=1.0))[0]      # t == t_max accepted
    assert not orc.triangle_test([0, 0, 0, 1, 0, 0, 2, 0, 0], orc.make_ray([0.5, 0, -1], [0, 0, 1]))[0]  # degenerate
    assert not orc.triangle_test(tri, orc.make_ray([2, 2, -1], [0, 0, 1]))[0]


def test_watertight_shared_edge(orc):
    # rays through the shared edge / vertices of two triangles hit at least one of them
    a = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    b = [1, 0, 0, 1, 1, 0, 0, 1, 0]
    rng = np.random.default_rng(1)
    for s in rng.uniform(0, 1, 200):
        p = np.array([1 - s, s, 0.0], dtype=np.float32)
        o = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), -2.0], dtype=np.float32)
        ray = orc.make_ray(o, p - o)
        assert orc.triangle_test(a, ray)[0] or orc.triangle_test(b, ray)[0]


def test_bvh_matches_brute_force_on_soups(orc, scenes):
    rng = np.random.default_rng(5)
    for seed, n, mp in ((0, 2000, 4), (1, 500, 1), (2, 3000, 16)):
        v, i = scenes.random_soup(n, seed=seed)
        bvh = orc.BVHAccel(v, i, mp)
        o = rng.uniform(-12, 12, (4000, 3))
        d = rng.normal(size=(4000, 3))
        rays = np.zeros((4000, 8), dtype=np.float32)
        rays[:, 0:3] = o
        rays[:, 3] = np.inf
        rays[:, 4:7] = d
        hits = bvh.intersect(rays)[0]
        bf = bvh.brute_force(rays)
        assert np.array_equal(bits(hits["t"]), bits(bf["t"]))
        same = hits["prim_id"] == bf["prim_id"]
        assert same.all(), f"{(~same).sum()} id mismatches (only exact-t ties may differ)"
        occ = bvh.intersect_p(rays)[0]
        assert np.array_equal(occ.astype(bool), hits["prim_id"] != 0xFFFFFFFF)
        assert (hits["prim_id"] != 0xFFFFFFFF).sum() > 100


def test_bvh_structure_invariants(orc, scenes):
    v, i = scenes.merge(scenes.uv_sphere(n_theta=24, n_phi=48), scenes.ground_grid())
    bvh = orc.BVHAccel(v, i, 4)
    nodes = bvh.nodes()
    prims = bvh.ordered_prims()
    assert sorted(prims.tolist()) == list(range(len(i)))
    leaves = nodes[nodes["n_prims"] > 0]
    assert leaves["n_prims"].sum() == len(i)
    assert np.array_equal(np.sort(leaves["offset"]), leaves["offset"])       # leaves emitted in order
    assert bvh.max_depth <= 64
    wb = bvh.world_bound()
    assert np.allclose(wb[:3], v.min(axis=0)) and np.allclose(wb[3:], v.max(axis=0))
    interior = np.nonzero(nodes["n_prims"] == 0)[0]
    for k in interior[:200]:
        l, r = nodes[k + 1], nodes[nodes[k]["offset"]]
        assert np.array_equal(nodes[k]["bounds"][:3], np.minimum(l["bounds"][:3], r["bounds"][:3]))
        assert np.array_equal(nodes[k]["bounds"][3:], np.maximum(l["bounds"][3:], r["bounds"][3:]))


def test_sphere_and_plane_analytic(orc, scenes):
    v, i = scenes.scene_c1()
    assert len(i) == 100024
    bvh = orc.BVHAccel(v, i, 4)
    cam = dict(scenes.C1_CAMERA, res=(96, 96))
    rays = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    hits = bvh.intersect(rays)[0]
    hit = hits["prim_id"] != 0xFFFFFFFF
    p = rays[:, 0:3].astype(np.float64) + hits["t"][:, None].astype(np.float64) * rays[:, 4:7].astype(np.float64)
    sphere = hit & (hits["prim_id"] < 99224)
    ground = hit & (hits["prim_id"] >= 99224)
    assert sphere.sum() > 500 and ground.sum() > 500
    r = np.linalg.norm(p[sphere], axis=1)
    assert (r <= 1.0 + 1e-5).all() and (r >= 1.0 - 3e-4).all()              # inside the chord sag of the tessellation
    t_plane = (-1.0 - rays[ground, 1].astype(np.float64)) / rays[ground, 5].astype(np.float64)
    assert np.allclose(hits["t"][ground], t_plane, rtol=1e-6)


def test_empty_and_single_triangle(orc):
    e = orc.BVHAccel(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), 4)
    rays = np.array([orc.make_ray([0, 0, -1], [0, 0, 1])])
    assert e.intersect(rays)[0]["prim_id"][0] == 0xFFFFFFFF and e.intersect_p(rays)[0][0] == 0
    one = orc.BVHAccel(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32), 4)
    rays = np.array([orc.make_ray([0.2, 0.2, -1], [0, 0, 1]), orc.make_ray([2, 2, -1], [0, 0, 1])])
    h = one.intersect(rays)[0]
    assert h["prim_id"].tolist() == [0, 0xFFFFFFFF] and h["t"][0] == 1.0 and np.isinf(h["t"][1])


@pytest.mark.parametrize("max_prims", [1, 4, 16])
def test_hlbvh_oracle_closest_hit_equals_brute_force_and_sah(orc, scenes, max_prims):
    """SplitMethod::HLBVH restatement (bvh.rs:475-772, pbrt-v3 semantics): every triangle lands in exactly one leaf, the walk
    over the HLBVH tree finds the brute-force hit, and ids / t bits equal the SAH tree's (different topology, same answer)."""
    v, i = scenes.random_soup(4000, seed=21 + max_prims)
    hl = orc.BVHAccel(v, i, max_prims, split_method=1)
    sah = orc.BVHAccel(v, i, max_prims)
    assert sorted(hl.ordered_prims().tolist()) == list(range(len(i)))
    nodes = hl.nodes()
    leaves = nodes[nodes["n_prims"] > 0]
    assert int(leaves["n_prims"].sum()) == len(i)
    assert (leaves["n_prims"] < max(max_prims, 2)).all() or max_prims == 1      # emit_lbvh: n < max_prims_in_node makes a leaf
    rng = np.random.default_rng(5)
    rays = np.zeros((20000, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-12, 12, (20000, 3))
    d = rng.normal(size=(20000, 3))
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3] = np.inf
    h, _ = hl.intersect(rays)[:2] if False else (hl.intersect(rays)[0], None)
    hs = sah.intersect(rays)[0]
    hb = hl.brute_force(rays)
    assert (h["prim_id"] != 0xFFFFFFFF).sum() > 500
    assert np.array_equal(h["prim_id"], hb["prim_id"]) and np.array_equal(h["t"].view(np.uint32), hb["t"].view(np.uint32))
    assert np.array_equal(h["t"].view(np.uint32), hs["t"].view(np.uint32))
    assert np.array_equal(hl.intersect_p(rays)[0], (h["prim_id"] != 0xFFFFFFFF).astype(np.uint8))

"""CPU tests of the oracle's analytic Sphere (oracle/oracle_sphere.hpp; reference src/shapes/sphere.rs, src/core/efloat.rs):
the contract acos / atan2, hits against the closed-form ray / sphere solution, BVH == brute force with mixed triangle and
sphere primitives, partial spheres, sample2 / pdf2 consistency, and the solid-angle estimate of a spherical light."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from oracle import oracle_path as OP


def _ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_contract_acos_atan2_accuracy():
    L = O.lib()
    xs = np.linspace(-1.0, 1.0, 4001, dtype=np.float32)
    got = np.array([L.orc_acos(float(x)) for x in xs], np.float32)
    want = np.arccos(xs.astype(np.float64))
    assert np.max(np.abs(got - want)) < 4e-7 * np.pi
    rng = np.random.default_rng(5)
    y, x = rng.normal(size=3000).astype(np.float32), rng.normal(size=3000).astype(np.float32)
    got = np.array([L.orc_atan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    want = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.max(np.abs(got - want)) < 1e-6
    assert L.orc_atan2(0.0, -1.0) == pytest.approx(np.pi) and L.orc_atan2(1.0, 0.0) == pytest.approx(np.pi / 2)


def _cam_rays(cam, res):
    return O.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], res)


def test_sphere_hits_match_closed_form(scenes):
    sc = scenes.scene_spheres(sphere_light=False, partial=False)
    ref = OP.Scene(sc, 4)
    bvh = ref.bvh()
    rays = _cam_rays(scenes.C2_CAMERA, (96, 96))
    hits = bvh.intersect(rays)[0]
    nt = len(sc["idx"])
    for k in range(3):                                           # the three translated balls: closed form in world space
        sel = hits["prim_id"] == nt + k
        assert sel.sum() > 100
        c = np.array(sc["spheres"][k]["center"], np.float64)
        r = sc["spheres"][k]["radius"]
        o, d = rays[sel, 0:3].astype(np.float64), rays[sel, 4:7].astype(np.float64)
        oc = o - c
        a, b, cc = (d * d).sum(1), 2 * (d * oc).sum(1), (oc * oc).sum(1) - r * r
        t = (-b - np.sqrt(b * b - 4 * a * cc)) / (2 * a)
        assert np.allclose(hits["t"][sel], t, rtol=2e-5)
        p = o + d * hits["t"][sel][:, None].astype(np.float64)
        assert np.allclose(np.linalg.norm(p - c, axis=1), r, rtol=1e-4)
        assert ((hits["b1"][sel] >= 0) & (hits["b1"][sel] <= 1) & (hits["b2"][sel] >= 0) & (hits["b2"][sel] <= 1)).all()
    assert (hits["prim_id"] == nt + 3).sum() > 50                # the ellipsoid is seen too


def test_bvh_equals_brute_force_with_spheres_and_partial_sphere(scenes):
    sc = scenes.scene_spheres()
    ref = OP.Scene(sc, 4)
    bvh = ref.bvh()
    rng = np.random.default_rng(11)
    n = 6000
    o = rng.uniform(20, 530, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7] = o, np.inf, d
    hits = bvh.intersect(rays)[0]
    brute = bvh.brute_force(rays)
    assert np.array_equal(hits["prim_id"], brute["prim_id"])
    assert np.array_equal(hits["t"].view(np.uint32), brute["t"].view(np.uint32))
    occ = bvh.intersect_p(rays)[0]
    assert np.array_equal(occ.astype(bool), hits["prim_id"] != 0xFFFFFFFF)
    nt = len(sc["idx"])
    part = nt + 4
    assert (hits["prim_id"] == part).sum() > 20
    # rays from inside the partial sphere's cut-away region reach geometry behind it: some rays hit its inner surface
    assert (hits["prim_id"] >= nt).sum() > 500


def test_sample2_and_pdf2_are_consistent(scenes):
    sc = scenes.scene_spheres()
    ref = OP.Scene(sc, 4)
    L = O.lib()
    k = len(sc["spheres"]) - 1                                   # the spherical light: centre (460, 420, 150), r = 25
    c, r = np.array([460.0, 420.0, 150.0]), 25.0
    rng = np.random.default_rng(3)
    for _ in range(200):
        p = rng.uniform(50, 500, size=3)
        if np.linalg.norm(p - c) < r * 1.2:
            continue
        ref9 = np.array([*p, 0, 0, 0, 0, 0, 0], np.float32)
        out = np.zeros(10, np.float32)
        u = rng.random(2)
        L.orc_sphere_sample2(C.c_void_p(ref.h), k, ref9.ctypes.data_as(C.c_void_p), float(u[0]), float(u[1]), out.ctypes.data_as(C.c_void_p))
        ps, n, pdf = out[0:3].astype(np.float64), out[6:9].astype(np.float64), float(out[9])
        assert abs(np.linalg.norm(ps - c) - r) < 1e-3 * r
        assert np.allclose(n, (ps - c) / r, atol=2e-4)
        sin2 = r * r / ((p - c) ** 2).sum()
        cone = 1.0 / (2 * np.pi * (1 - np.sqrt(1 - sin2)))
        assert pdf == pytest.approx(cone, rel=2e-3)
        wi = (ps - p) / np.linalg.norm(ps - p)
        assert np.dot(n, -wi) > -1e-3                            # the sampled point faces the reference point
        pdf2 = L.orc_sphere_pdf2(C.c_void_p(ref.h), k, ref9.ctypes.data_as(C.c_void_p), wi.astype(np.float32).ctypes.data_as(C.c_void_p))
        assert pdf2 == pytest.approx(pdf, rel=1e-5)
    # inside the sphere: area sampling converted to solid angle; pdf2 follows Shape::pdf2 through an intersection
    ref9 = np.array([*(c + [3.0, -2.0, 5.0]), 0, 0, 0, 0, 0, 0], np.float32)
    tot = 0.0
    for i in range(2000):
        out = np.zeros(10, np.float32)
        u = rng.random(2)
        L.orc_sphere_sample2(C.c_void_p(ref.h), k, ref9.ctypes.data_as(C.c_void_p), float(u[0]), float(u[1]), out.ctypes.data_as(C.c_void_p))
        if out[9] > 0:
            tot += 1.0 / out[9]
            if i < 50:
                wi = out[0:3].astype(np.float64) - ref9[0:3]
                wi /= np.linalg.norm(wi)
                pdf2 = L.orc_sphere_pdf2(C.c_void_p(ref.h), k, ref9.ctypes.data_as(C.c_void_p), wi.astype(np.float32).ctypes.data_as(C.c_void_p))
                assert pdf2 == pytest.approx(float(out[9]), rel=2e-3)
    assert tot / 2000 == pytest.approx(4 * np.pi, rel=0.05)      # E[1 / pdf] = the full sphere of directions


def test_sphere_light_irradiance_matches_closed_form(scenes):
    """A matte floor lit only by a spherical light: the radiance seen by a camera ray that hits the floor at x equals
    kd / pi * E(x), E = L * pi * r^2 / d^2 * cos(theta) for a sphere fully above the horizon (direct light only)."""
    L_e, r, c = 10.0, 30.0, np.array([0.0, 200.0, 0.0])
    quad = np.array([[-2000, 0, -2000], [2000, 0, -2000], [2000, 0, 2000], [-2000, 0, 2000]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2]], np.uint32)
    sc = dict(verts=quad, idx=idx, tri_material=np.array([0, 0], np.uint32), materials=[dict(type="matte", kd=(0.6, 0.6, 0.6))],
              spheres=[dict(center=tuple(c), radius=r, material=0)], lights=[dict(type="area", prim=2, L=(L_e,) * 3, two_sided=False)])
    ref = OP.Scene(sc, 4)
    cam = dict(pos=(150.0, 80.0, -120.0), look=(100.0, 0.0, 40.0), up=(0.0, 1.0, 0.0), fov=10.0, res=(8, 8))
    pd = OP.path_desc(max_depth=1, rr_threshold=1.0, light_strategy="uniform", spp=512)
    xyzw, _ = ref.render(cam, OP.film_desc(cam["res"]), pd, mode=1)
    rgb = OP.resolve_rgb(xyzw)
    centre = rgb[3:5, 3:5].mean()
    x = np.array([100.0, 0.0, 40.0])
    d = np.linalg.norm(c - x)
    want = 0.6 / np.pi * L_e * np.pi * r * r / (d * d) * (c[1] / d)
    assert centre == pytest.approx(want, rel=0.03)

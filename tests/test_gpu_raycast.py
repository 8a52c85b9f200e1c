"""GPU parity tests (B200): every result that comes back through the C ABI is compared bit for bit with the CPU
oracle on the same inputs.  Integer / index / t-bit work: zero mismatches allowed."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def assert_hits_equal(got, ref, b0_got=None, b0_ref=None):
    assert np.array_equal(got["prim_id"], ref["prim_id"]), f"{(got['prim_id'] != ref['prim_id']).sum()} prim_id mismatches"
    assert np.array_equal(bits(got["t"]), bits(ref["t"])), "t bits differ"
    assert np.array_equal(bits(got["b1"]), bits(ref["b1"])) and np.array_equal(bits(got["b2"]), bits(ref["b2"]))
    if b0_got is not None:
        assert np.array_equal(bits(b0_got), bits(b0_ref))


def random_rays(n, seed, extent=12.0, finite_tmax=False):
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = rng.uniform(-extent, extent, (n, 3))
    rays[:, 3] = rng.uniform(1.0, 30.0, n) if finite_tmax else np.inf
    rays[:, 4:7] = rng.normal(size=(n, 3))
    return rays


def test_rng_streams_bit_exact(gpu, orc):
    got = gpu.rng_uniform_floats(0, 300, 16)
    for s in (0, 1, 17, 299):
        assert np.array_equal(bits(got[s]), bits(orc.pcg32_float(s, 16)))
    big = gpu.rng_uniform_floats(2 ** 40 + 5, 4, 8)
    assert np.array_equal(bits(big[3]), bits(orc.pcg32_float(2 ** 40 + 8, 8)))


def test_camera_rays_bit_exact(gpu, orc, scenes):
    for cam in (scenes.C1_CAMERA, scenes.C3_CAMERA):
        res = (256, 192)
        camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], res)
        rng = np.random.default_rng(3)
        pf = rng.uniform(0, 1, (50000, 2)).astype(np.float32) * np.array(res, np.float32)
        got = camera.generate_rays(pf)
        ref = orc.camera_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], res, pf)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.fixture(scope="module")
def c1(gpu, orc, scenes):
    v, i = scenes.scene_c1()
    cam = scenes.C1_CAMERA
    rays = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    return dict(v=v, i=i, rays=rays, accel=gpu.BVHAccel(v, i, 4), ref=orc.BVHAccel(v, i, 4))


def test_c1_closest_hit_bit_exact(c1):
    """BASELINE config 0: 1024x1024 primary rays vs the 100,024-triangle sphere+ground, SAH BVH."""
    hits, b0 = c1["accel"].intersect(c1["rays"], want_b0=True)
    ref, ref_b0, _ = c1["ref"].intersect(c1["rays"], want_b0=True)
    assert (ref["prim_id"] != 0xFFFFFFFF).mean() > 0.5
    assert_hits_equal(hits, ref, b0, ref_b0)


def test_c1_any_hit_bit_exact(c1):
    occ = c1["accel"].intersect_p(c1["rays"])
    ref = c1["ref"].intersect_p(c1["rays"])[0]
    assert np.array_equal(occ, ref)


def test_c1_bvh_export_equals_oracle(c1):
    nodes, prims = c1["accel"].export()
    ref_nodes = c1["ref"].nodes()
    assert np.array_equal(prims, c1["ref"].ordered_prims())
    for f in ("bounds", "offset", "n_prims", "axis"):
        assert np.array_equal(nodes[f], ref_nodes[f])
    assert np.array_equal(c1["accel"].world_bound(), c1["ref"].world_bound())


@pytest.mark.parametrize("max_prims,n_tris", [(1, 3000), (4, 20000), (16, 20000), (255, 5000)])
def test_random_soup_bit_exact(gpu, orc, scenes, max_prims, n_tris):
    v, i = scenes.random_soup(n_tris, seed=max_prims)
    accel = gpu.BVHAccel(v, i, max_prims)
    ref = orc.BVHAccel(v, i, max_prims)
    for finite in (False, True):
        rays = random_rays(100003, seed=n_tris + finite, finite_tmax=finite)        # ragged batch size
        # axis-aligned directions exercise inv_dir = +-inf and 0*inf = NaN in the slab test
        rays[:3000, 4:7] = np.eye(3, dtype=np.float32)[np.arange(3000) % 3] * np.where(np.arange(3000) % 2, 1, -1)[:, None].astype(np.float32)
        rays[3000:4000, 5] = -0.0
        hits, b0 = accel.intersect(rays, want_b0=True)
        rh, rb0, _ = ref.intersect(rays, want_b0=True)
        assert (rh["prim_id"] != 0xFFFFFFFF).sum() > 1000
        assert_hits_equal(hits, rh, b0, rb0)
        assert np.array_equal(accel.intersect_p(rays), ref.intersect_p(rays)[0])


def test_any_hit_equals_closest_hit_found(gpu, scenes):
    v, i = scenes.random_soup(30000, seed=77)
    accel = gpu.BVHAccel(v, i, 4)
    rays = random_rays(200000, seed=5, finite_tmax=True)
    hits = accel.intersect(rays)
    assert np.array_equal(accel.intersect_p(rays).astype(bool), hits["prim_id"] != 0xFFFFFFFF)
    assert (hits["t"][hits["prim_id"] != 0xFFFFFFFF] <= rays[hits["prim_id"] != 0xFFFFFFFF, 3]).all()
    miss = hits["prim_id"] == 0xFFFFFFFF
    assert np.array_equal(bits(hits["t"][miss]), bits(rays[miss, 3]))       # ray.t_max untouched on a miss


def test_shared_edges_and_coplanar_ties(gpu, orc):
    """Rays aimed exactly at shared vertices/edges of a fine grid: equal-t candidates -> last tested wins (tie order)."""
    n = 64
    xs = np.linspace(-1, 1, n + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs)
    v = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size, np.float32)], axis=1)
    a = (np.arange(n)[:, None] * (n + 1) + np.arange(n)[None, :]).ravel()
    i = np.concatenate([np.stack([a, a + 1, a + n + 1], 1), np.stack([a + 1, a + n + 2, a + n + 1], 1)]).astype(np.uint32)
    v = np.concatenate([v, v + np.array([0, 0, 1], np.float32)])            # second, parallel layer
    i = np.concatenate([i, i + (n + 1) ** 2])
    accel, ref = gpu.BVHAccel(v, i, 4), orc.BVHAccel(v, i, 4)
    targets = v[: (n + 1) ** 2]
    # (a) axis-aligned rays through the vertices: origin on the box planes with a zero direction component gives
    #     0 * inf = NaN in the slab test (geometry.rs:724-749 let it fall through as a miss) — must match exactly
    rays = np.zeros((len(targets), 8), np.float32)
    rays[:, 0:3] = targets + np.array([0, 0, -2], np.float32)
    rays[:, 3] = np.inf
    rays[:, 6] = 1.0
    assert_hits_equal(accel.intersect(rays), ref.intersect(rays)[0])
    assert np.array_equal(accel.intersect_p(rays), ref.intersect_p(rays)[0])
    # (b) rays from one eye point through every shared vertex and edge midpoint
    mids = 0.5 * (targets[:-1] + targets[1:])
    pts = np.concatenate([targets, mids])
    eye = np.array([0.13, 0.29, -3.0], np.float32)
    rays = np.zeros((len(pts), 8), np.float32)
    rays[:, 0:3] = eye
    rays[:, 3] = np.inf
    rays[:, 4:7] = pts - eye
    hits = accel.intersect(rays)
    rh = ref.intersect(rays)[0]
    assert (rh["prim_id"] != 0xFFFFFFFF).mean() > 0.95
    assert_hits_equal(hits, rh)
    assert np.array_equal(accel.intersect_p(rays), ref.intersect_p(rays)[0])


def test_empty_single_and_zero_rays(gpu, orc):
    e = gpu.BVHAccel(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    rays = np.array([orc.make_ray([0, 0, -1], [0, 0, 1])])
    h = e.intersect(rays)
    assert h["prim_id"][0] == 0xFFFFFFFF and np.isinf(h["t"][0]) and e.intersect_p(rays)[0] == 0
    one = gpu.BVHAccel(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32))
    rays = np.array([orc.make_ray([0.2, 0.2, -1], [0, 0, 1]), orc.make_ray([2, 2, -1], [0, 0, 1])])
    h = one.intersect(rays)
    assert h["prim_id"].tolist() == [0, 0xFFFFFFFF] and h["t"][0] == 1.0
    assert one.intersect_p(rays).tolist() == [1, 0]
    assert len(one.intersect(np.zeros((0, 8), np.float32))) == 0
    assert len(one.intersect_p(np.zeros((0, 8), np.float32))) == 0


def test_intersect_before_build_is_an_error(gpu):
    s = gpu.Scene(np.zeros((3, 3), np.float32), np.array([[0, 1, 2]], np.uint32))
    out = np.empty(1, dtype=gpu.HIT_DTYPE)
    rays = np.zeros((1, 8), np.float32)
    rc = gpu.lib().pb2_intersect(s.h, rays.ctypes.data, 1, out.ctypes.data, None)
    assert rc == -3 and b"build_bvh" in gpu.lib().pb2_last_error()


def _c3_pass(gpu, accel, camera, n, light):
    """primary closest-hit -> shadow any-hit + incoherent bounce closest-hit, all device resident."""
    d_rays, d_hits, d_b0 = gpu.DeviceBuffer(n * 32), gpu.DeviceBuffer(n * 16), gpu.DeviceBuffer(n * 4)
    d_srays, d_brays = gpu.DeviceBuffer(n * 32), gpu.DeviceBuffer(n * 32)
    d_occ, d_bhits = gpu.DeviceBuffer(n), gpu.DeviceBuffer(n * 16)
    camera.primary_rays_device(d_rays.ptr)
    accel.intersect_device(d_rays.ptr, n, d_hits.ptr, d_b0.ptr)
    accel.spawn_shadow_rays_device(d_rays.ptr, d_hits.ptr, n, light, d_srays.ptr)
    accel.spawn_bounce_rays_device(d_rays.ptr, d_hits.ptr, n, d_brays.ptr)
    accel.intersect_p_device(d_srays.ptr, n, d_occ.ptr)
    accel.intersect_device(d_brays.ptr, n, d_bhits.ptr)
    # the one-pass builder (bench.py) writes the same two ray sets
    d_s2, d_b2 = gpu.DeviceBuffer(n * 32), gpu.DeviceBuffer(n * 32)
    accel.spawn_shadow_bounce_rays_device(d_rays.ptr, d_hits.ptr, n, light, d_s2.ptr, d_b2.ptr)
    gpu.check(gpu.lib().pb2_device_synchronize())
    assert np.array_equal(d_s2.download(np.uint32, n * 8), d_srays.download(np.uint32, n * 8))
    assert np.array_equal(d_b2.download(np.uint32, n * 8), d_brays.download(np.uint32, n * 8))
    return dict(rays=d_rays.download(np.float32, n * 8).reshape(-1, 8), hits=d_hits.download(gpu.HIT_DTYPE, n),
                b0=d_b0.download(np.float32, n), srays=d_srays.download(np.float32, n * 8).reshape(-1, 8),
                brays=d_brays.download(np.float32, n * 8).reshape(-1, 8), occ=d_occ.download(np.uint8, n),
                bhits=d_bhits.download(gpu.HIT_DTYPE, n))


def _check_c3(gpu, orc, scenes, grid_n, res, light=None):
    light = scenes.C3_POINT_LIGHT if light is None else light
    v, i = scenes.displaced_grid(n=grid_n)
    cam = dict(scenes.C3_CAMERA, res=res)
    accel = gpu.BVHAccel(v, i, 4)
    camera = gpu.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    n = res[0] * res[1]
    got = _c3_pass(gpu, accel, camera, n, light)
    ref = orc.BVHAccel(v, i, 4)
    ref_rays = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    assert np.array_equal(got["rays"].view(np.uint32), ref_rays.view(np.uint32))
    rh, rb0, _ = ref.intersect(ref_rays, want_b0=True)
    assert (rh["prim_id"] != 0xFFFFFFFF).mean() > 0.3
    assert_hits_equal(got["hits"], rh, got["b0"], rb0)
    srays = orc.spawn_shadow_rays(ref, rh, rb0, light)
    brays = orc.spawn_bounce_rays(ref, ref_rays, rh, rb0)
    assert np.array_equal(got["srays"].view(np.uint32), srays.view(np.uint32)), "shadow rays differ"
    assert np.array_equal(got["brays"].view(np.uint32), brays.view(np.uint32)), "bounce rays differ"
    assert np.array_equal(got["occ"], ref.intersect_p(srays)[0])
    assert_hits_equal(got["bhits"], ref.intersect(brays)[0])
    return accel, got


def test_c3_reduced_all_ray_types_bit_exact(gpu, orc, scenes):
    _check_c3(gpu, orc, scenes, grid_n=400, res=(256, 256))
    # a grazing light so that a good part of the shadow rays is occluded (the C3 light is overhead: nothing is)
    _, got = _check_c3(gpu, orc, scenes, grid_n=400, res=(256, 256), light=(60.0, 4.0, 0.0))
    assert 0.05 < got["occ"].mean() < 0.95
    assert (got["bhits"]["prim_id"] != 0xFFFFFFFF).mean() > 0.02


def test_c3_full_size_10m_triangles_bit_exact(gpu, orc, scenes):
    """BASELINE config 2 at full size: 10,008,338 triangles, 1024x1024 primary + shadow + incoherent rays, every
    result compared with the oracle (which builds its own BVH), plus the size-independent properties."""
    accel, got = _check_c3(gpu, orc, scenes, grid_n=2237, res=(1024, 1024))
    assert accel.info()[1] == 10008338
    hit = got["hits"]["prim_id"] != 0xFFFFFFFF
    # host-buffer path == device-resident path; any-hit == (closest hit found) on the incoherent batch
    sub = slice(0, 300000)
    assert np.array_equal(accel.intersect(got["brays"][sub]), got["bhits"][sub])
    assert np.array_equal(accel.intersect_p(got["brays"][sub]).astype(bool), got["bhits"]["prim_id"][sub] != 0xFFFFFFFF)
    # idempotence: re-tracing a ray clipped to twice its hit distance returns the same primitive and t.  (Clipping to
    # exactly t would not: the slab test's `t_min < ray.t_max` is strict, so a leaf whose box face carries the hit
    # triangle is culled — reference semantics, geometry.rs:749.)
    clipped = got["rays"].copy()
    clipped[:, 3] = 2.0 * got["hits"]["t"]
    again = accel.intersect(clipped[hit][:200000])
    assert np.array_equal(again["prim_id"], got["hits"]["prim_id"][hit][:200000])
    assert np.array_equal(bits(again["t"]), bits(got["hits"]["t"][hit][:200000]))


@pytest.mark.gpu
@pytest.mark.parametrize("scene,max_prims", [("c1", 4), ("soup", 4), ("soup", 1), ("soup", 16), ("c3_small", 4), ("cornell", 4), ("dups", 4)])
def test_hlbvh_gpu_build_equals_oracle(gpu, orc, scenes, scene, max_prims):
    """SplitMethod::HLBVH (bvh.rs:475-772) built on the GPU: the flattened node array, the primitive order and the world bound
    equal the oracle's sequential HLBVH bit for bit, and closest hit / any hit over that tree are bit-exact."""
    if scene == "c1":
        v, i = scenes.scene_c1()
    elif scene == "soup":
        v, i = scenes.random_soup(30000, seed=11 + max_prims)
    elif scene == "c3_small":
        v, i = scenes.scene_c3(300)
    elif scene == "cornell":
        sc = scenes.scene_c2()
        v, i = sc["verts"], sc["idx"]
    else:       # many triangles with identical centroids (equal Morton codes all the way down) + a few others
        v0, i0 = scenes.random_soup(50, seed=5)
        v = np.concatenate([v0, np.tile(v0[:3], (40, 1))]).astype(np.float32)
        i = np.concatenate([i0, (len(v0) + np.arange(120)).reshape(40, 3)]).astype(np.uint32)
    accel = gpu.BVHAccel(v, i, max_prims, split_method=1)
    ref = orc.BVHAccel(v, i, max_prims, split_method=1)
    nodes, prims = accel.export()
    ref_nodes = ref.nodes()
    assert len(nodes) == len(ref_nodes)
    assert np.array_equal(prims, ref.ordered_prims())
    for f in ("bounds", "offset", "n_prims", "axis"):
        assert np.array_equal(nodes[f], ref_nodes[f]), f
    assert np.array_equal(accel.world_bound(), ref.world_bound())
    assert accel.info()[2] == ref.max_depth
    cam = scenes.C2_CAMERA if scene == "cornell" else scenes.C1_CAMERA
    rays = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], (256, 256))
    rays = np.concatenate([rays, random_rays(60000, seed=9, finite_tmax=True)])
    hits, b0 = accel.intersect(rays, want_b0=True)
    rh, rb0, _ = ref.intersect(rays, want_b0=True)
    assert_hits_equal(hits, rh, b0, rb0)
    assert np.array_equal(accel.intersect_p(rays), ref.intersect_p(rays)[0])
    assert sum(accel.build_stats()) > 0.0


@pytest.mark.gpu
def test_split_method_errors(gpu, scenes):
    v, i = scenes.random_soup(100, seed=1)
    with pytest.raises(gpu.Pb2Error):
        gpu.BVHAccel(v, i, 4, split_method=4)       # bvh.rs:199-204 has four split methods
    with pytest.raises(gpu.Pb2Error):
        gpu.BVHAccel(v, i, 4, split_method=-1)


@pytest.mark.gpu
@pytest.mark.parametrize("split", [2, 3])
def test_middle_and_equal_counts_trees_traverse_bit_exact(gpu, orc, scenes, split):
    """SplitMethod::Middle / ::EqualCounts trees (bvh.rs:331-360) through the same QuadNode kernels: hits equal the oracle's
    walk of the same tree bit for bit, and the primitive a ray hits is the one the SAH tree finds (same geometry)."""
    v, i = scenes.scene_c1()
    accel = gpu.BVHAccel(v, i, 4, split_method=split)
    ref = orc.BVHAccel(v, i, 4, split_method=split)
    cam = scenes.C1_CAMERA
    rays = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], (256, 256))
    rays = np.concatenate([rays, random_rays(60000, seed=4, finite_tmax=True)])
    hits, b0 = accel.intersect(rays, want_b0=True)
    rh, rb0, _ = ref.intersect(rays, want_b0=True)
    assert_hits_equal(hits, rh, b0, rb0)
    assert np.array_equal(accel.intersect_p(rays), ref.intersect_p(rays)[0])
    sah = gpu.BVHAccel(v, i, 4).intersect(rays)
    assert np.array_equal(bits(sah["t"]), bits(hits["t"]))


@pytest.mark.gpu
def test_hlbvh_leaf_limit_is_an_error(gpu, scenes):
    """LinearBVHNode::n_primitives is 16 bits (bvh.rs:129-135; pbrt-v3 CHECKs it): 70,000 triangles with one centroid cannot be
    split by Morton bits (or by SAH: coincident centroids make a leaf, bvh.rs:320-334) and must be refused, not truncated;
    60,000 fit, and the walk over that one fat leaf keeps the reference's tie rule."""
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    n = 70000
    v = np.tile(tri, (n, 1))
    i = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    scene = gpu.Scene(v, i)
    for split in (1, 0):
        with pytest.raises(gpu.Pb2Error):
            gpu.BVHAccel(scene, max_prims_in_node=4, split_method=split)
    n = 60000
    accel = gpu.BVHAccel(gpu.Scene(v[:3 * n], i[:n]), max_prims_in_node=255, split_method=0)
    rays = np.zeros((4, 8), np.float32)
    rays[:, 0:3] = (0.25, 0.25, -1.0)
    rays[:, 6] = 1.0
    rays[:, 3] = np.inf
    hits = accel.intersect(rays)
    assert (hits["prim_id"] == n - 1).all() and (hits["t"] == 1.0).all()       # equal t: the last triangle tested wins


@pytest.mark.gpu
def test_entry_points_are_thread_safe(gpu, orc, scenes):
    """The reference's Primitive / Integrator objects are Sync + Send and called from rayon workers (parallel.rs:4-17); the C ABI
    serialises per scene handle.  Eight host threads hammer pb2_intersect / pb2_intersect_p on one shared scene and on scenes of
    their own, and two render into separate films of one scene: every result equals the single-threaded one bit for bit."""
    import threading
    v, i = scenes.scene_c1()
    shared = gpu.BVHAccel(v, i, 4)
    cam = scenes.C1_CAMERA
    rays = np.concatenate([orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], (192, 192)), random_rays(30000, seed=23, finite_tmax=True)])
    want_h, want_o = shared.intersect(rays), shared.intersect_p(rays)
    errors = []

    def work(k):
        try:
            own_v, own_i = scenes.random_soup(2000 + 100 * k, seed=k)
            own = gpu.BVHAccel(own_v, own_i, 4)
            own_want = own.intersect(rays[:5000])
            for rep in range(6):
                sl = slice(k * 1000, k * 1000 + 20000 + rep)
                h = shared.intersect(rays[sl])
                o = shared.intersect_p(rays[sl])
                if not (np.array_equal(h.view(np.uint32), want_h[sl].view(np.uint32)) and np.array_equal(o, want_o[sl])):
                    errors.append(f"thread {k} rep {rep}: shared scene results differ")
                if not np.array_equal(own.intersect(rays[:5000]).view(np.uint32), own_want.view(np.uint32)):
                    errors.append(f"thread {k} rep {rep}: own scene results differ")
        except Exception as e:          # noqa: BLE001
            errors.append(f"thread {k}: {e!r}")

    threads = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    # two threads rendering with one scene into two films
    sc = scenes.scene_c2()
    accel = gpu.BVHAccel(gpu.scene_from_dict(sc), 4)
    camera = gpu.PerspectiveCamera(*[scenes.C2_CAMERA[k] for k in ("pos", "look", "up", "fov")], (64, 64))
    films = [gpu.Film((64, 64)) for _ in range(3)]
    integ = [gpu.PathIntegrator(accel, camera, max_depth=4, spp=8) for _ in range(3)]
    integ[2].render(films[2])
    ts = [threading.Thread(target=lambda j=j: integ[j].render(films[j])) for j in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for j in range(2):
        assert np.array_equal(films[j].read_xyzw().view(np.uint32), films[2].read_xyzw().view(np.uint32))


@pytest.mark.gpu
def test_hostile_rays_match_the_oracle(gpu, orc, scenes):
    """Rays no renderer should produce but a caller might: NaN / inf origins and directions, the zero direction, t_max of 0,
    negative, NaN and subnormal size, coordinates near FLT_MAX, origins exactly on vertices / edges / box planes.  No crash, no
    hang, and every result (id, t bits, any-hit flag) equals the oracle's — the comparisons of the reference are followed
    literally, NaN outcomes included."""
    v, i = scenes.merge(scenes.uv_sphere(n_theta=24, n_phi=48), scenes.ground_grid())
    accel = gpu.BVHAccel(v, i, 4)
    ref = orc.BVHAccel(v, i, 4)
    rng = np.random.default_rng(99)
    base = random_rays(4000, seed=5, extent=3.0)
    special = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-45, 3.0e38, -3.0e38, 1e-30]
    rays = []
    for k in range(4000):
        r = base[k].copy()
        for _ in range(rng.integers(1, 4)):
            r[rng.choice([0, 1, 2, 3, 4, 5, 6])] = special[rng.integers(len(special))]
        rays.append(r)
    rays = np.array(rays, np.float32)
    zero_dir = base[:200].copy(); zero_dir[:, 4:7] = 0.0
    neg_t = base[:200].copy(); neg_t[:, 3] = -1.0
    on_vertex = base[:600].copy(); on_vertex[:, 0:3] = v[rng.integers(0, len(v), 600)]
    on_plane = base[:400].copy(); on_plane[:, 1] = -1.0; on_plane[:, 5] = 0.0          # in the ground plane, parallel to it
    allr = np.concatenate([rays, zero_dir, neg_t, on_vertex, on_plane]).astype(np.float32)
    hits, b0 = accel.intersect(allr, want_b0=True)
    rh, rb0, _ = ref.intersect(allr, want_b0=True)
    found = rh["prim_id"] != 0xFFFFFFFF
    assert np.array_equal(hits["prim_id"], rh["prim_id"])

    def same_bits_or_both_nan(a, b):          # a NaN is a NaN on both sides (a ray with a NaN component "hits" under the
        return ((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))).all()      # reference's comparisons); its payload bits are not specified

    assert same_bits_or_both_nan(hits["t"], rh["t"])          # a miss reports the ray's own t_max
    assert same_bits_or_both_nan(hits["b1"][found], rh["b1"][found]) and same_bits_or_both_nan(hits["b2"][found], rh["b2"][found])
    assert same_bits_or_both_nan(b0[found], rb0[found])
    finite = found & np.isfinite(rh["t"])
    assert np.array_equal(bits(hits["t"])[finite], bits(rh["t"])[finite]) and finite.sum() > 300
    assert np.array_equal(accel.intersect_p(allr), ref.intersect_p(allr)[0])


def test_async_entry_points_match_the_synchronous_ones(gpu, orc, scenes):
    """pb2_intersect_async / pb2_intersect_p_async + pb2_scene_wait: several batches (a multi-chunk one, a short one, an empty
    one, one that needs b0) enqueued back to back on the scene's ring land in their own host buffers with the bits the
    synchronous calls and the oracle give; a synchronous call also drains earlier asynchronous ones."""
    import ctypes as C
    v, i = scenes.merge(scenes.uv_sphere(n_theta=40, n_phi=80), scenes.ground_grid())
    accel = gpu.BVHAccel(v, i, 4)
    ref = orc.BVHAccel(v, i, 4)
    L = gpu.lib()

    def pinned(nbytes):
        p = C.c_void_p()
        gpu.check(L.pb2_host_alloc(max(nbytes, 16), C.byref(p)))
        return p

    sizes = [300001, 17, 0, 150000, 40000]
    batches = []
    for k, n in enumerate(sizes):
        rays = random_rays(n, seed=100 + k, extent=6.0, finite_tmax=(k % 2 == 1))
        p_r, p_h, p_b, p_o = pinned(n * 32), pinned(n * 16), pinned(n * 4), pinned(n)
        C.memmove(p_r, rays.ctypes.data, rays.nbytes)
        batches.append((n, rays, p_r, p_h, p_b, p_o))
    for rnd in range(2):
        for k, (n, rays, p_r, p_h, p_b, p_o) in enumerate(batches):
            gpu.check(L.pb2_intersect_async(accel.h, p_r, n, p_h, p_b if k == 3 else None))
            gpu.check(L.pb2_intersect_p_async(accel.h, p_r, n, p_o))
        if rnd == 0:
            gpu.check(L.pb2_scene_wait(accel.h))
        else:                                     # a synchronous call on the same scene waits for everything before it
            assert np.array_equal(accel.intersect(batches[1][1])["prim_id"], ref.intersect(batches[1][1])[0]["prim_id"])
        for k, (n, rays, p_r, p_h, p_b, p_o) in enumerate(batches):
            if n == 0:
                continue
            hits = np.frombuffer((C.c_char * (n * 16)).from_address(p_h.value), dtype=gpu.HIT_DTYPE).copy()
            occ = np.frombuffer((C.c_char * n).from_address(p_o.value), dtype=np.uint8).copy()
            want = ref.intersect(rays, want_b0=True)
            assert_hits_equal(hits, want[0])
            assert np.array_equal(occ, ref.intersect_p(rays)[0])
            if k == 3:
                b0 = np.frombuffer((C.c_char * (n * 4)).from_address(p_b.value), dtype=np.float32).copy()
                assert np.array_equal(bits(b0), bits(want[1]))
            C.memset(p_h, 0xFF, n * 16)
            C.memset(p_o, 0xFF, n)
    gpu.check(L.pb2_scene_wait(accel.h))          # nothing pending: returns at once
    # pb2_scene_wait_until(k): all but the k newest batches are complete when it returns (the ring is in order)
    for k, (n, rays, p_r, p_h, p_b, p_o) in enumerate(batches):
        gpu.check(L.pb2_intersect_async(accel.h, p_r, n, p_h, None))
    gpu.check(L.pb2_scene_wait_until(accel.h, 2))
    for k, (n, rays, p_r, p_h, p_b, p_o) in enumerate(batches[:-2]):
        if n:
            hits = np.frombuffer((C.c_char * (n * 16)).from_address(p_h.value), dtype=gpu.HIT_DTYPE).copy()
            assert_hits_equal(hits, ref.intersect(rays)[0])
    gpu.check(L.pb2_scene_wait_until(accel.h, 100))    # more than were ever enqueued: nothing to wait for
    gpu.check(L.pb2_scene_wait_until(accel.h, 0))
    n, rays, _, p_h, _, _ = batches[-1]
    assert_hits_equal(np.frombuffer((C.c_char * (n * 16)).from_address(p_h.value), dtype=gpu.HIT_DTYPE).copy(), ref.intersect(rays)[0])
    for _, _, p_r, p_h, p_b, p_o in batches:
        for p in (p_r, p_h, p_b, p_o):
            gpu.check(L.pb2_host_free(p))

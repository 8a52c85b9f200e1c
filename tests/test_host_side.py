"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol, the host SAH
build equals the oracle's flattened array, camera matrices match, and error paths behave (no GPU calls)."""
import ctypes as C

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(pb2):
    L = pb2.lib()
    syms = pb2.header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_product_does_not_link_or_import_the_oracle(pb2):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(pb2.LIB_PATH))
    out = subprocess.run(["grep", "-rIl", "-i", "oracle", os.path.join(root, "pbrt-rs_b200", "csrc"),
                          os.path.join(root, "pbrt-rs_b200", "__init__.py"), os.path.join(root, "include")],
                         capture_output=True, text=True).stdout.strip()
    assert out == "", f"product sources mention the oracle: {out}"
    deps = subprocess.run(["ldd", pb2.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in deps


def test_last_error_and_invalid_arguments(pb2):
    L = pb2.lib()
    h = C.c_void_p()
    v = np.zeros((3, 3), np.float32)
    bad = np.array([[0, 1, 7]], np.uint32)
    rc = L.pb2_scene_create(v.ctypes.data, 3, bad.ctypes.data, 1, None, None, 0, None, 0, C.byref(h))
    assert rc == -1 and b"references vertex" in L.pb2_last_error()
    v[0, 0] = np.nan
    ok = np.array([[0, 1, 2]], np.uint32)
    rc = L.pb2_scene_create(v.ctypes.data, 3, ok.ctypes.data, 1, None, None, 0, None, 0, C.byref(h))
    assert rc == -1 and b"non-finite" in L.pb2_last_error()
    assert L.pb2_world_bound(None, None) == -1
    with pytest.raises(pb2.Pb2Error):
        pb2.check(L.pb2_scene_build_bvh(None, 4, 0))


def _same_nodes(a, b):
    return (np.array_equal(a["bounds"], b["bounds"]) and np.array_equal(a["offset"], b["offset"])
            and np.array_equal(a["n_prims"], b["n_prims"]) and np.array_equal(a["axis"], b["axis"]))


@pytest.mark.parametrize("case", ["soup4", "soup1", "soup16", "sphere", "tiny", "coincident"])
def test_host_sah_build_equals_oracle(pb2, orc, scenes, case):
    if case.startswith("soup"):
        mp = int(case[4:])
        v, i = scenes.random_soup(20000, seed=mp)
    elif case == "sphere":
        mp = 4
        v, i = scenes.merge(scenes.uv_sphere(n_theta=60, n_phi=120), scenes.ground_grid())
    elif case == "tiny":
        mp = 4
        v, i = scenes.random_soup(3, seed=9)
    else:  # nine triangles sharing one centroid -> one oversized leaf (bvh.rs:311)
        mp = 4
        tris = []
        for k in range(9):
            s = 1.0 + k
            tris.append([[-s, -s, 0], [s, -s, 0], [0, 2 * s, 0]])
        v = np.array(tris, np.float32).reshape(-1, 3)
        i = np.arange(27, dtype=np.uint32).reshape(-1, 3)
    mine = pb2.BVHAccel(v, i, max_prims_in_node=mp, host_only=True)
    ref = orc.BVHAccel(v, i, mp)
    nodes, prims = mine.export()
    assert len(nodes) == ref.num_nodes
    assert np.array_equal(prims, ref.ordered_prims())
    assert _same_nodes(nodes, ref.nodes())
    assert mine.info()[2] == ref.max_depth
    assert np.array_equal(mine.world_bound(), ref.world_bound())


@pytest.mark.parametrize("split", [2, 3])
@pytest.mark.parametrize("case", ["soup", "grid", "coincident", "ties"])
def test_host_middle_and_equal_counts_builds_equal_oracle(pb2, orc, scenes, case, split):
    """SplitMethod::Middle (2) / ::EqualCounts (3) (bvh.rs:331-360): node array and primitive order equal the oracle's, on a
    soup, on a grid large enough for the parallel top of the builder, with coincident centroids, and with many equal centroid
    coordinates (Middle's partition is improper there and falls through to EqualCounts)."""
    if case == "soup":
        v, i = scenes.random_soup(20000, seed=3)
    elif case == "grid":
        v, i = scenes.displaced_grid(n=300)
    elif case == "coincident":
        v0, i0 = scenes.random_soup(50, seed=5)
        v = np.concatenate([v0, np.tile(v0[:3], (40, 1))]).astype(np.float32)
        i = np.concatenate([i0, (len(v0) + np.arange(120)).reshape(40, 3)]).astype(np.uint32)
    else:
        rng = np.random.default_rng(2)
        c = np.stack([rng.integers(0, 3, 4000), rng.integers(0, 2, 4000), np.zeros(4000)], axis=1).astype(np.float32)[:, None, :]
        v = (c + np.array([[-0.25, -0.25, 0], [0.25, -0.25, 0], [0, 0.5, 0]], np.float32)[None]).reshape(-1, 3).astype(np.float32)
        i = np.arange(12000, dtype=np.uint32).reshape(-1, 3)
    mine = pb2.BVHAccel(v, i, max_prims_in_node=4, split_method=split, host_only=True)
    ref = orc.BVHAccel(v, i, 4, split_method=split)
    nodes, prims = mine.export()
    assert len(nodes) == ref.num_nodes
    assert np.array_equal(prims, ref.ordered_prims())
    assert _same_nodes(nodes, ref.nodes())
    assert mine.info()[2] == ref.max_depth
    sah = orc.BVHAccel(v, i, 4)
    assert sorted(prims) == list(range(len(i))) and ref.num_nodes >= sah.num_nodes       # leaves of one primitive unless centroids coincide


def test_host_build_parallel_path_equals_oracle(pb2, orc, scenes):
    # large enough that the builder splits the top of the tree across worker threads (grain = n / (8 * threads))
    v, i = scenes.displaced_grid(n=300)
    mine = pb2.BVHAccel(v, i, max_prims_in_node=4, host_only=True)
    ref = orc.BVHAccel(v, i, 4)
    nodes, prims = mine.export()
    assert np.array_equal(prims, ref.ordered_prims())
    assert _same_nodes(nodes, ref.nodes())


def test_camera_matrices_equal_oracle(pb2, orc, scenes):
    for cam in (scenes.C1_CAMERA, scenes.C3_CAMERA, dict(pos=(278, 273, -800), look=(278, 273, 0), up=(0, 1, 0), fov=39.3, res=(1920, 1080))):
        r2c, c2w = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"]).matrices()
        o_r2c, o_c2w = orc.camera_matrices(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
        assert np.array_equal(r2c.view(np.uint32), o_r2c.view(np.uint32))
        assert np.array_equal(c2w.view(np.uint32), o_c2w.view(np.uint32))


def test_empty_scene_host(pb2):
    b = pb2.BVHAccel(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), host_only=True)
    assert b.info() == (0, 0, 0)
    wb = b.world_bound()
    assert (wb[:3] > 1e38).all() and (wb[3:] < -1e38).all()


@pytest.mark.parametrize("scene", [["c2"], ["soup"], ["c3", "150"]])
def test_quad_layout_walk_matches_reference_order(scene):
    import os
    """QuadNode collapse (bvh_build.cpp): a CPU walk of the two-levels-per-record layout with the kernel's rules tests the
    same triangles in the same order as the reference-order binary walk — identical ids, t bits and triangle-test counts
    on primary and bounce rays (tools/quad_sim.py asserts; the arithmetic is the oracle's)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "quad_sim.py")] + scene, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout


def test_oversized_leaf_is_refused_not_truncated(pb2):
    """70,000 triangles with one centroid end up in one SAH leaf (bvh.rs:320-334); the 32-byte node counts primitives in 16
    bits, so the build must fail with PB2_ERR_LIMIT instead of dropping triangles."""
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    v = np.tile(tri, (70000, 1))
    i = np.arange(3 * 70000, dtype=np.uint32).reshape(-1, 3)
    with pytest.raises(pb2.Pb2Error) as e:
        pb2.BVHAccel(v, i, 255, host_only=True)
    assert e.value.code == -5
    ok = pb2.BVHAccel(v[:3 * 60000], i[:60000], 255, host_only=True)
    assert ok.info()[:2] == (1, 60000)


def test_rust_shim_matches_header(pb2, tmp_path):
    """rust_shim/src/lib.rs cannot be compiled here (no rustc); hold it to the header mechanically: its extern block declares
    exactly the header's functions, and every abi_size!(T, n) equals the C sizeof (gcc on include/pbrt_b200.h) and the size of
    the ctypes structure the Python binding passes."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(pb2.LIB_PATH))
    src = open(os.path.join(root, "rust_shim", "src", "lib.rs")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    declared = sorted(set(re.findall(r"pub fn (pb2_[a-z0-9_]+)\s*\(", block)))
    assert declared == pb2.header_symbols()
    sizes = dict((t, int(n)) for t, n in re.findall(r"abi_size!\((pb2_[a-z_]+), (\d+)\);", src))
    assert set(sizes) == {"pb2_ray", "pb2_hit", "pb2_material", "pb2_light", "pb2_sphere", "pb2_medium", "pb2_camera", "pb2_film_desc", "pb2_path_desc"}
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "pbrt_b200.h"\nint main(void){' +
                    "".join(f'printf("{t} %zu\\n", sizeof({t}));' for t in sorted(sizes)) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(prog), "-o", str(exe)])
    c_sizes = dict((l.split()[0], int(l.split()[1])) for l in subprocess.check_output([str(exe)], text=True).splitlines())
    assert c_sizes == sizes
    py = {"pb2_material": pb2.Material, "pb2_light": pb2.Light, "pb2_sphere": pb2.Sphere, "pb2_medium": pb2.Medium, "pb2_camera": pb2.CameraDesc, "pb2_film_desc": pb2.FilmDesc,
          "pb2_path_desc": pb2.PathDesc}
    for t, cls in py.items():
        assert C.sizeof(cls) == sizes[t], t
    assert sizes["pb2_ray"] == 32 and pb2.HIT_DTYPE.itemsize == sizes["pb2_hit"]      # rays travel as float32[n, 8]
    # the two trait impls the shim exists for are code, not comments
    assert re.search(r"^impl Primitive for B200Accel", src, flags=re.M) and re.search(r"^impl Integrator for B200PathIntegrator", src, flags=re.M)


def test_host_build_with_analytic_spheres_equals_oracle(pb2, scenes):
    """Spheres join the primitive list after the triangles (ids n_tris + k) with Shape::world_bound = object_to_world applied to the
    eight corners of the object bound (shape.rs:18-20, transform.rs:569-606): node array and primitive order equal the oracle's."""
    from oracle import oracle_path as OP
    sc = scenes.scene_spheres()
    mine = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4, host_only=True)
    ref = OP.Scene(sc, 4).bvh()
    nodes, prims = mine.export()
    assert len(prims) == len(sc["idx"]) + len(sc["spheres"])
    assert np.array_equal(prims, ref.ordered_prims())
    assert _same_nodes(nodes, ref.nodes())
    assert np.array_equal(mine.world_bound(), ref.world_bound())
    # error paths: a projective matrix, a zero radius, a sphere after the build
    L = pb2.lib()
    bad = pb2.sphere_from_dict(dict(center=(0, 0, 0), radius=1.0))
    bad.object_to_world[12] = 0.5
    arr = (pb2.Sphere * 1)(bad)
    s2 = pb2.scene_from_dict(scenes.scene_c2())
    assert L.pb2_scene_add_spheres(s2.h, C.cast(arr, C.c_void_p), 1) == -1 and b"affine" in L.pb2_last_error()
    arr = (pb2.Sphere * 1)(pb2.sphere_from_dict(dict(center=(0, 0, 0), radius=0.0)))
    assert L.pb2_scene_add_spheres(s2.h, C.cast(arr, C.c_void_p), 1) == -1 and b"radius" in L.pb2_last_error()
    arr = (pb2.Sphere * 1)(pb2.sphere_from_dict(dict(center=(0, 0, 0), radius=1.0)))
    assert L.pb2_scene_add_spheres(mine.scene.h, C.cast(arr, C.c_void_p), 1) == -3


def test_nudge_equals_next_float_up_down(tmp_path):
    """offset_ray_origin's branch-free step (pb2_math.cuh nudge) == next_float_up / next_float_down (pbrt.rs:43-77, geometry.rs:1146-1152)
    on every 251st binary32 value and the special values; `tools/nudge_check.cpp` without an argument is the exhaustive run."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "nudge_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-o", exe, os.path.join(root, "tools", "nudge_check.cpp")])
    out = subprocess.run([exe, "251"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert " 0 differences" in out.stdout


def test_fast_div_is_exact(tmp_path):
    """slot -> (sample, pixel, row, column) uses n / d = hi64(n * ceil(2^64 / d)) instead of two 32-bit divisions (wavefront.cuh
    fast_div): equal to n / d on 104 M (n, d) pairs — film widths, powers of two, multiples +- 2, the top of the range, random."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fast_div_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(root, "tools", "fast_div_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert " 0 differences" in out.stdout

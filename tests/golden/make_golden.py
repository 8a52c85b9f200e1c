"""Generates the committed golden fixtures of tests/golden/ from the CPU oracle (TEST INFRASTRUCTURE).

The reference (lazytiger/pbrt-rs) cannot be compiled or run (no Rust toolchain) and its tests hold no vector
for this path, so these fixtures are outputs of oracle/ — the restatement of the reference's algorithm — frozen at the commit
that introduced them.  They pin BOTH sides from then on: `-m "not gpu"` tests check the oracle still reproduces them, `-m gpu`
tests check the CUDA path against the same bytes, so the two cannot drift together unnoticed.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
    python tests/golden/make_golden.py --round2 # rewrites tests/golden/path_r02.npz only
    python tests/golden/make_golden.py --sphere # rewrites tests/golden/sphere_r02.npz only (analytic spheres)
    python tests/golden/make_golden.py --volpath # rewrites tests/golden/volpath_r02.npz only (VolPathIntegrator + media)

Inputs (meshes, rays, (pixel, sample) pairs) are stored in the fixtures, so no generator has to reproduce them bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def golden_scene(scenes):
    """Cornell box with a glass tall block and a point light beside the quad area light: matte + specular shading classes,
    three lights (two emissive triangles + one delta light) for the "power" / "spatial" strategies.  No libm on this path
    (no roughness remap, no spot cone, box filter)."""
    sc = scenes.cornell_box()
    sc["materials"] = sc["materials"] + [dict(type="glass", kr=(1.0, 1.0, 1.0), kt=(1.0, 1.0, 1.0), eta=1.5)]
    tm = sc["tri_material"].copy()
    tm[-10:] = 3                                     # the tall block's ten triangles
    sc["tri_material"] = tm
    sc["lights"] = sc["lights"] + [dict(type="point", p=(100.0, 400.0, 100.0), I=(4.0e5, 5.0e5, 6.0e5))]
    return sc


GOLDEN_CAMERA = dict(pos=(278.0, 273.0, -800.0), look=(278.0, 273.0, 0.0), up=(0.0, 1.0, 0.0), fov=39.3, res=(48, 48))
GOLDEN_PATH = dict(max_depth=6, rr_threshold=1.0, spp=4)


def main():
    ge.build()
    orc, scenes = ge.load_oracle(), ge.load_scenes()
    from oracle import oracle_path as OP
    # ---- ray casting: BVHAccel::intersect / intersect_p, SAH build ------------------------------------------------------------
    verts, idx = scenes.merge(scenes.uv_sphere(n_theta=20, n_phi=40), scenes.ground_grid())
    cam = dict(scenes.C1_CAMERA, res=(64, 64))
    rays = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    rng = np.random.default_rng(2026)
    extra = np.zeros((4096, 8), np.float32)
    extra[:, 0:3] = rng.uniform(-6, 6, (4096, 3))
    extra[:, 3] = np.where(rng.random(4096) < 0.5, np.inf, rng.uniform(0.5, 8.0, 4096)).astype(np.float32)
    extra[:, 4:7] = rng.normal(size=(4096, 3))
    extra[::97, 4] = 0.0                              # axis-parallel rays: the literal one-level walk
    rays = np.concatenate([np.asarray(rays, np.float32).reshape(-1, 8), extra])
    bvh = orc.BVHAccel(verts, idx, 4)
    hits, b0 = bvh.intersect(rays, want_b0=True)[:2]
    occ = bvh.intersect_p(rays)[0]
    nodes = bvh.nodes()
    np.savez_compressed(os.path.join(HERE, "raycast.npz"), verts=np.asarray(verts, np.float32), idx=np.asarray(idx, np.uint32), rays=rays,
                        prim_id=hits["prim_id"], t=hits["t"].view(np.uint32), b1=hits["b1"].view(np.uint32), b2=hits["b2"].view(np.uint32),
                        b0=np.asarray(b0, np.float32).view(np.uint32), occluded=np.asarray(occ, np.uint8),
                        nodes=np.frombuffer(np.ascontiguousarray(nodes).tobytes(), np.uint8), ordered_prims=bvh.ordered_prims(),
                        world_bound=np.asarray(bvh.world_bound(), np.float32))
    # ---- path tracing: PathIntegrator::li per (pixel, sample), film, SpatialLightDistribution ----------------------------------
    sc = golden_scene(scenes)
    ref = OP.Scene(sc, 4)
    fd = OP.film_desc(GOLDEN_CAMERA["res"])
    n = 2048
    xy = np.stack([rng.integers(0, 48, n), rng.integers(0, 48, n)], axis=1).astype(np.int32)
    s = rng.integers(0, 4, n).astype(np.uint32)
    out = dict(xy=xy, sample=s)
    for strat in ("uniform", "power", "spatial"):
        L, pf = ref.path_li(GOLDEN_CAMERA, fd, OP.path_desc(light_strategy=strat, **GOLDEN_PATH), xy, s)
        out["L_" + strat] = L.view(np.uint32)
        out["p_film"] = pf.view(np.uint32)
    for sampler, skw in (("halton", {}), ("stratified", dict(x_samples=2, y_samples=2)), ("zerotwo", {})):
        L, _ = ref.path_li(GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", sampler=sampler, **GOLDEN_PATH, **skw), xy, s)
        out["L_" + sampler] = L.view(np.uint32)
    film, _ = ref.render(GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="spatial", **GOLDEN_PATH), mode=1)
    out["film_spatial"] = film.view(np.uint32)
    nv = ref.spatial_grid()
    out["spatial_grid"] = np.asarray(nv, np.int32)
    vox = np.stack([rng.integers(0, nv[0], 64), rng.integers(0, nv[1], 64), rng.integers(0, nv[2], 64)], axis=1).astype(np.int32)
    out["spatial_voxels"] = vox
    out["spatial_func"] = np.stack([ref.spatial_voxel(v, 3)[0] for v in vox]).view(np.uint32)
    out["spatial_cdf"] = np.stack([ref.spatial_voxel(v, 3)[1] for v in vox]).view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "path.npz"), **out)
    for f in ("raycast.npz", "path.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    make_round2(scenes, OP)


def make_round2(scenes, OP):
    """Fixtures of the components added in round 2 (same scene, camera and (pixel, sample) list as path.npz; a file of their
    own so that the round-1 fixtures stay byte-identical): per-sample radiance and film under the SobolSampler."""
    g = np.load(os.path.join(HERE, "path.npz"))
    ref = OP.Scene(golden_scene(scenes), 4)
    fd = OP.film_desc(GOLDEN_CAMERA["res"])
    out = {}
    L, pf = ref.path_li(GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", sampler="sobol", **GOLDEN_PATH), g["xy"], g["sample"])
    out["L_sobol"], out["p_film_sobol"] = L.view(np.uint32), pf.view(np.uint32)
    film, _ = ref.render(GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", sampler="sobol", **GOLDEN_PATH), mode=1)
    out["film_sobol"] = film.view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "path_r02.npz"), **out)
    print("path_r02.npz", os.path.getsize(os.path.join(HERE, "path_r02.npz")), "bytes")


def golden_sphere_scene(scenes):
    """scenes.scene_spheres() (analytic matte / plastic / glass balls, an ellipsoid, a partial reversed sphere, a spherical area
    light, the quad light and a point light) without the roughness remap (no libm on the path)."""
    sc = scenes.scene_spheres()
    sc["materials"] = [dict(m, remap=False) if m["type"] == "plastic" else m for m in sc["materials"]]
    return sc


def make_sphere(scenes, orc, OP):
    """tests/golden/sphere_r02.npz: closest hits / any hits of rays against the sphere scene, per-sample radiance and film."""
    sc = golden_sphere_scene(scenes)
    ref = OP.Scene(sc, 4)
    bvh = ref.bvh()
    rng = np.random.default_rng(4242)
    cam = dict(GOLDEN_CAMERA, res=(64, 64))
    rays = np.asarray(orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"]), np.float32).reshape(-1, 8)
    extra = np.zeros((4096, 8), np.float32)
    extra[:, 0:3] = rng.uniform(10, 540, (4096, 3))
    extra[:, 3] = np.where(rng.random(4096) < 0.5, np.inf, rng.uniform(50.0, 400.0, 4096)).astype(np.float32)
    extra[:, 4:7] = rng.normal(size=(4096, 3))
    extra[::97, 5] = 0.0
    rays = np.concatenate([rays, extra])
    hits, b0 = bvh.intersect(rays, want_b0=True)[:2]
    out = dict(rays=rays, prim_id=hits["prim_id"], t=hits["t"].view(np.uint32), b1=hits["b1"].view(np.uint32), b2=hits["b2"].view(np.uint32),
               b0=np.asarray(b0, np.float32).view(np.uint32), occluded=np.asarray(bvh.intersect_p(rays)[0], np.uint8),
               nodes=np.frombuffer(np.ascontiguousarray(bvh.nodes()).tobytes(), np.uint8), ordered_prims=bvh.ordered_prims())
    fd = OP.film_desc(GOLDEN_CAMERA["res"])
    n = 2048
    xy = np.stack([rng.integers(0, 48, n), rng.integers(0, 48, n)], axis=1).astype(np.int32)
    s = rng.integers(0, 4, n).astype(np.uint32)
    out["xy"], out["sample"] = xy, s
    for strat in ("uniform", "power", "spatial"):
        L, _ = ref.path_li(GOLDEN_CAMERA, fd, OP.path_desc(light_strategy=strat, **GOLDEN_PATH), xy, s)
        out["L_" + strat] = L.view(np.uint32)
    film, _ = ref.render(GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", **GOLDEN_PATH), mode=1)
    out["film_power"] = film.view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "sphere_r02.npz"), **out)
    print("sphere_r02.npz", os.path.getsize(os.path.join(HERE, "sphere_r02.npz")), "bytes")


def make_volpath(scenes, OP):
    """tests/golden/volpath_r02.npz: VolPathIntegrator over scenes.scene_media() (fog, smoke box behind a material-less interface,
    glass sphere): per-sample radiance under two light strategies and a film."""
    sc = scenes.scene_media()
    ref = OP.Scene(sc, 4)
    rng = np.random.default_rng(777)
    fd = OP.film_desc(GOLDEN_CAMERA["res"])
    n = 2048
    xy = np.stack([rng.integers(0, 48, n), rng.integers(0, 48, n)], axis=1).astype(np.int32)
    s = rng.integers(0, 4, n).astype(np.uint32)
    out = dict(xy=xy, sample=s)
    kw = dict(GOLDEN_PATH, max_depth=8)
    for strat in ("uniform", "power"):
        L, _ = ref.path_li(GOLDEN_CAMERA, fd, OP.path_desc(light_strategy=strat, integrator="volpath", **kw), xy, s)
        out["L_" + strat] = L.view(np.uint32)
    film, _ = ref.render(GOLDEN_CAMERA, fd, OP.path_desc(light_strategy="power", integrator="volpath", **kw), mode=1)
    out["film_power"] = film.view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "volpath_r02.npz"), **out)
    print("volpath_r02.npz", os.path.getsize(os.path.join(HERE, "volpath_r02.npz")), "bytes")


if __name__ == "__main__":
    if "--volpath" in sys.argv:                      # only volpath_r02.npz
        ge.build()
        from oracle import oracle_path as _OP
        make_volpath(ge.load_scenes(), _OP)
    elif "--sphere" in sys.argv:                       # only sphere_r02.npz
        ge.build()
        from oracle import oracle_path as _OP
        make_sphere(ge.load_scenes(), ge.load_oracle(), _OP)
    elif "--round2" in sys.argv:                       # only path_r02.npz (the round-1 fixtures are left untouched)
        ge.build()
        from oracle import oracle_path as _OP
        make_round2(ge.load_scenes(), _OP)
    else:
        main()

#!/usr/bin/env python
"""Exercise every kernel of the backend once at a representative (but short) size: the command the per-kernel ncu capture
(`ncu --set full --kernel-id :::1`, first invocation of each kernel) runs under.  Prints a few numbers so a plain run can be
checked for success first."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    pb2, scenes = ge.load_package(), ge.load_scenes()
    pb2.init(0)
    n = int(os.environ.get("PROF_GRID", "1000"))
    res = int(os.environ.get("PROF_RES", "768"))
    verts, idx = scenes.scene_c3(n)
    cam = dict(scenes.C3_CAMERA, res=(res, res))
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    nr = res * res
    out = []
    for split in (0, 1):                      # host SAH build + upload, then the device HLBVH build
        accel = pb2.BVHAccel(verts, idx, 4, split_method=split)
        d_rays, d_hits, d_b0 = pb2.DeviceBuffer(nr * 32), pb2.DeviceBuffer(nr * 16), pb2.DeviceBuffer(nr * 4)
        d_s, d_b, d_occ = pb2.DeviceBuffer(nr * 32), pb2.DeviceBuffer(nr * 32), pb2.DeviceBuffer(nr)
        camera.primary_rays_device(d_rays.ptr)
        accel.intersect_device(d_rays.ptr, nr, d_hits.ptr, d_b0.ptr)
        accel.spawn_shadow_rays_device(d_rays.ptr, d_hits.ptr, nr, scenes.C3_POINT_LIGHT, d_s.ptr)
        accel.spawn_bounce_rays_device(d_rays.ptr, d_hits.ptr, nr, d_b.ptr)
        accel.intersect_p_device(d_s.ptr, nr, d_occ.ptr)
        accel.intersect_device(d_b.ptr, nr, d_hits.ptr, None)
        pb2.check(pb2.lib().pb2_device_synchronize())
        out.append(accel.info())
    # path tracing: mixed materials, all four light kinds, smooth normals + UVs, stratified sampler (per-pixel tables),
    # box film (ordered accumulation + strays) and Gaussian film (atomics), thin lens
    sc = scenes.scene_c4_smooth(n_theta=80, n_phi=160)
    sc["lights"] = scenes.scene_all_lights(8, 16)["lights"]
    cam4 = dict(scenes.C4_CAMERA, res=(960, 540))
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), 4)
    camera4 = pb2.PerspectiveCamera(cam4["pos"], cam4["look"], cam4["up"], cam4["fov"], cam4["res"], lens_radius=5.0, focal_distance=1000.0)
    for sampler, kw, film in (("stratified", dict(x_samples=4, y_samples=4), pb2.Film(cam4["res"])),
                              ("random", {}, pb2.Film(cam4["res"], filter="gaussian", radius=(2.0, 2.0))),
                              ("halton", {}, pb2.Film(cam4["res"], max_sample_luminance=10.0))):
        integ = pb2.PathIntegrator(accel, camera4, max_depth=8, light_strategy="power", spp=16, sampler=sampler, **kw)
        integ.render(film)
        out.append((sampler, float(film.resolve_rgb().mean()), integ.counters()["extend_rays"]))
    # SpatialLightDistribution: the voxel tables (k_spatial_contrib, k_spatial_distrib) and a render that looks them up
    film = pb2.Film(cam4["res"])
    integ = pb2.PathIntegrator(accel, camera4, max_depth=8, light_strategy="spatial", spp=4)
    integ.render(film)
    out.append(("spatial", accel.spatial_light_distribution(tables=False), float(film.resolve_rgb().mean())))
    # plain mesh (the kernels the benchmarks run: k_shade<*, false, false>)
    sc = scenes.scene_c4()
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), 4)
    camera4 = pb2.PerspectiveCamera(cam4["pos"], cam4["look"], cam4["up"], cam4["fov"], cam4["res"])
    film = pb2.Film(cam4["res"])
    integ = pb2.PathIntegrator(accel, camera4, **dict(scenes.C4_PATH, spp=16))
    integ.render(film)
    out.append(("c4", float(film.resolve_rgb().mean())))
    print(out)


if __name__ == "__main__":
    main()

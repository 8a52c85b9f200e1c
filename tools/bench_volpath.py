#!/usr/bin/env python
"""VolPathIntegrator on the fog + smoke scene (bench.py path_extras.volpath: 512x512 @ 16 spp, maxdepth 8) and on the same room
without the material-less smoke box (no interfaces: no host read-backs): ms per frame, Msamples/s, rays, launches — for the
wavefront stages (default) or, with PB2_VOLPATH_MEGAKERNEL=1 in the environment, the one-thread-per-path kernel; the film's
checksum (crc32 of the XYZ + weight accumulators) is printed so that the two runs can be compared bit for bit."""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

pb2, scenes = ge.load_package(), ge.load_scenes()
pb2.init(0)
st = torch.cuda.current_stream().cuda_stream
res, spp = (int(os.environ.get("VOL_RES", "512")),) * 2, int(os.environ.get("VOL_SPP", "16"))
mode = "megakernel" if os.environ.get("PB2_VOLPATH_MEGAKERNEL", "0") not in ("", "0") else "wavefront"


def no_interface(sc):
    """scene_media without its material-less primitives (and so without the smoke box's medium)"""
    sc = dict(sc)
    keep = np.asarray(sc["tri_material"]) != 0xFFFFFFFF if "tri_material" in sc else None
    return sc, keep


for name, sc in (("fog+smoke", scenes.scene_media()),):
    cam = dict(scenes.C2_CAMERA, res=res)
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = pb2.PathIntegrator(accel, camera, spp=spp, max_depth=8, rr_threshold=1.0, light_strategy="power", integrator="volpath")
    film = pb2.Film(cam["res"])
    integ.render(film, 0, 2, stream=st)
    torch.cuda.synchronize()
    ms = []
    for rep in range(3):
        film.clear()
        c0 = integ.counters()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); integ.render(film, stream=st); b.record()
        torch.cuda.synchronize()
        c1 = integ.counters()
        ms.append(a.elapsed_time(b))
    rays = {k: int(c1[k] - c0[k]) for k in ("extend_rays", "shadow_rays", "mis_rays")}
    t = min(ms)
    crc = zlib.crc32(film.read_xyzw().tobytes())
    print(f"{mode} {name} {res[0]}x{res[1]} @ {spp} spp: {t:.3f} ms = {res[0] * res[1] * spp / t / 1e3:.1f} Msamples/s, "
          f"{sum(rays.values()) / t / 1e3:.1f} Mrays/s, rays {rays}, launches {int(c1.get('launches', 0) - c0.get('launches', 0))}, film crc32 {crc:08x}", flush=True)

#!/usr/bin/env python
"""Render one C2 (or C4-shaped) frame — the short command the ncu captures of the wavefront kernels run under."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="c2", choices=["c2", "c4", "spheres", "media"])
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--frames", type=int, default=1)
    args = ap.parse_args()
    pb2, scenes = ge.load_package(), ge.load_scenes()
    pb2.init(0)
    if args.scene == "c2":
        sc, cam, pk = scenes.scene_c2(), scenes.C2_CAMERA, dict(scenes.C2_PATH, spp=args.spp)
    elif args.scene == "spheres":      # bench.py path_extras.spheres
        sc, cam, pk = scenes.scene_spheres(), dict(scenes.C2_CAMERA, res=(1024, 1024)), dict(max_depth=5, rr_threshold=1.0, light_strategy="power", spp=args.spp)
    elif args.scene == "media":        # bench.py path_extras.volpath
        sc, cam = scenes.scene_media(), dict(scenes.C2_CAMERA, res=(512, 512))
        pk = dict(max_depth=8, rr_threshold=1.0, light_strategy="power", integrator="volpath", spp=args.spp)
    else:
        sc, cam, pk = scenes.scene_c4(), scenes.C4_CAMERA, dict(scenes.C4_PATH, spp=args.spp)
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = pb2.PathIntegrator(accel, camera, **pk)
    film = pb2.Film(cam["res"])
    for _ in range(args.frames):
        film.clear()
        integ.render(film)
    rgb = film.resolve_rgb()
    print("mean rgb", rgb.mean(axis=(0, 1)), "counters", integ.counters())


if __name__ == "__main__":
    main()

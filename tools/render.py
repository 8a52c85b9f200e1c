#!/usr/bin/env python
"""Minimal render driver: build a scene, render it with the wavefront PathIntegrator, write the image.
    python tools/render.py --scene c2 --spp 64 --sampler halton --out cornell.ppm
Scenes: c2 (Cornell box), c4 (room with matte / plastic / glass spheres).  Output: .ppm (8-bit sRGB) or .pfm (float)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="c2", choices=["c2", "c4"])
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--res", type=int, nargs=2, default=None)
    ap.add_argument("--sampler", default="random", choices=["random", "halton"])
    ap.add_argument("--split", default="sah", choices=["sah", "hlbvh"])
    ap.add_argument("--out", default="out.ppm")
    args = ap.parse_args()
    pb2, scenes = ge.load_package(), ge.load_scenes()
    pb2.init(0)
    if args.scene == "c2":
        sc, cam, pk = scenes.scene_c2(), dict(scenes.C2_CAMERA), dict(scenes.C2_PATH)
    else:
        sc, cam, pk = scenes.scene_c4(), dict(scenes.C4_CAMERA), dict(scenes.C4_PATH)
    if args.res:
        cam["res"] = tuple(args.res)
    pk["spp"] = args.spp
    t0 = time.time()
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4, split_method=1 if args.split == "hlbvh" else 0)
    t1 = time.time()
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = pb2.PathIntegrator(accel, camera, sampler=args.sampler, **pk)
    film = pb2.Film(cam["res"])
    integ.render(film)
    film.write_image(args.out)
    t2 = time.time()
    n = cam["res"][0] * cam["res"][1] * args.spp
    print(f"{args.scene}: {len(sc['idx'])} triangles, BVH ({args.split}) {t1 - t0:.2f} s, {cam['res'][0]}x{cam['res'][1]} @ {args.spp} spp "
          f"({args.sampler}) rendered + written in {t2 - t1:.2f} s = {n / (t2 - t1) / 1e6:.0f} Msamples/s -> {args.out}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Minimal render driver: build a scene, render it with the wavefront PathIntegrator, write the image.
    python tools/render.py --scene c2 --spp 64 --sampler halton --out cornell.ppm
Scenes: c2 (Cornell box), c4 (room with matte / plastic / glass spheres), c4smooth (the same mesh with analytic vertex normals and
UVs), lights (c4smooth + spot and distant lights).  Output: .ppm (8-bit sRGB) or .pfm (float)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="c2", choices=["c2", "c4", "c4smooth", "lights", "materials"])
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--res", type=int, nargs=2, default=None)
    ap.add_argument("--sampler", default="random", choices=["random", "halton", "stratified", "zerotwo"])
    ap.add_argument("--filter", default="box", choices=["box", "gaussian", "triangle", "mitchell", "sinc"])
    ap.add_argument("--radius", type=float, default=None)
    ap.add_argument("--lens", type=float, nargs=2, default=None, metavar=("RADIUS", "FOCAL_DISTANCE"))
    ap.add_argument("--crop", type=float, nargs=4, default=None, metavar=("X0", "Y0", "X1", "Y1"))
    ap.add_argument("--split", default="sah", choices=["sah", "hlbvh"])
    ap.add_argument("--out", default="out.ppm")
    args = ap.parse_args()
    pb2, scenes = ge.load_package(), ge.load_scenes()
    pb2.init(0)
    if args.scene == "c2":
        sc, cam, pk = scenes.scene_c2(), dict(scenes.C2_CAMERA), dict(scenes.C2_PATH)
    elif args.scene == "c4":
        sc, cam, pk = scenes.scene_c4(), dict(scenes.C4_CAMERA), dict(scenes.C4_PATH)
    elif args.scene == "materials":
        sc, cam, pk = scenes.scene_materials(96, 192), dict(scenes.C4_CAMERA), dict(scenes.C4_PATH)
        import numpy as np
        v = sc["verts"].astype(np.float64)                  # smooth shading: analytic normals for the five balls
        nrm = np.zeros_like(v)
        for c, mats in (((140.0, 90.0, 280.0), (3,)), ((278.0, 130.0, 220.0), (4,)), ((416.0, 90.0, 280.0), (5,)), ((200.0, 45.0, 120.0), (6,)),
                        ((350.0, 45.0, 120.0), (7, 8))):
            used = np.unique(sc["idx"][np.isin(sc["tri_material"], mats)])
            nrm[used] = v[used] - np.array(c)
        flat = np.linalg.norm(nrm, axis=1) == 0
        fn = np.cross(v[sc["idx"][:, 1]] - v[sc["idx"][:, 0]], v[sc["idx"][:, 2]] - v[sc["idx"][:, 0]])
        for k in range(3):
            np.add.at(nrm, sc["idx"][:, k], fn * flat[sc["idx"][:, k]][:, None])
        sc["normals"] = (nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)).astype(np.float32)
    else:
        sc, cam, pk = scenes.scene_c4_smooth(n_theta=64, n_phi=128, emissive_normals=False), dict(scenes.C4_CAMERA), dict(scenes.C4_PATH)
        if args.scene == "lights":
            extra = scenes.scene_all_lights(8, 16)["lights"][-2:]
            sc["lights"] = [l for l in sc["lights"] if l["type"] == "area"] + extra
    if args.res:
        cam["res"] = tuple(args.res)
    pk["spp"] = args.spp
    t0 = time.time()
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4, split_method=1 if args.split == "hlbvh" else 0)
    t1 = time.time()
    lens = dict(lens_radius=args.lens[0], focal_distance=args.lens[1]) if args.lens else {}
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"], **lens)
    skw = {}
    if args.sampler == "stratified":
        side = max(1, int(round(args.spp ** 0.5)))
        skw = dict(x_samples=side, y_samples=side)
    integ = pb2.PathIntegrator(accel, camera, sampler=args.sampler, **pk, **skw)
    args.spp = integ.spp
    fkw = dict(radius=(args.radius, args.radius)) if args.radius else ({} if args.filter == "box" else dict(radius=(2.0, 2.0)))
    film = pb2.Film(cam["res"], filter=args.filter, crop=args.crop, **fkw)
    integ.render(film)
    film.write_image(args.out)
    t2 = time.time()
    n = cam["res"][0] * cam["res"][1] * args.spp
    print(f"{args.scene}: {len(sc['idx'])} triangles, BVH ({args.split}) {t1 - t0:.2f} s, {cam['res'][0]}x{cam['res'][1]} @ {args.spp} spp "
          f"({args.sampler}) rendered + written in {t2 - t1:.2f} s = {n / (t2 - t1) / 1e6:.0f} Msamples/s -> {args.out}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Measurement-only: per-stage times of one C4 batch (1920x1080 @ 8 spp) and one C2 batch (512x512 @ 32 spp) with the queues as
the kernels leave them, and — PB2_EXP_SORT=1 — sorted by slot before every kernel that reads them.  Needs the variant
build `tools/build_variant.sh sortq -DPB2_EXPERIMENT_SORTED_QUEUES` selected with PB2_LIB (the product library has none of this)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

pb2, scenes = ge.load_package(), ge.load_scenes()
pb2.init(0)
for name, sc, cam, pk in (("c4", scenes.scene_c4(), scenes.C4_CAMERA, dict(scenes.C4_PATH, spp=8)),
                          ("c2", scenes.scene_c2(), scenes.C2_CAMERA, dict(scenes.C2_PATH, spp=32))):
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = pb2.PathIntegrator(accel, camera, **pk)
    film = pb2.Film(cam["res"])
    for rep in range(3):
        film.clear()
        print(f"--- {name} rep {rep}", file=sys.stderr, flush=True)
        integ.render(film)
    print(name, "mean rgb", float(film.resolve_rgb().mean()), flush=True)

#!/usr/bin/env python
"""Sweep the traversal scheduling knobs (pb2_set_trace_tuning) on the path-traced workloads: ms per C2 frame (512x512 @ 16 spp)
and per C4 step (1920x1080 @ 4 spp) for every setting.  Results do not depend on the knobs (mean rgb is printed as a check)."""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

pb2, scenes = ge.load_package(), ge.load_scenes()
pb2.init(0)
L = pb2.lib()
st = torch.cuda.current_stream().cuda_stream


def setup(sc, cam, pk):
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    return accel, camera, pb2.PathIntegrator(accel, camera, **pk), pb2.Film(cam["res"])


SPP2, SPP4 = int(os.environ.get("TUNE_C2_SPP", "16")), int(os.environ.get("TUNE_C4_SPP", "4"))
c2 = setup(scenes.scene_c2(), scenes.C2_CAMERA, dict(scenes.C2_PATH, spp=SPP2))
c4 = setup(scenes.scene_c4(), scenes.C4_CAMERA, dict(scenes.C4_PATH, spp=SPP4))


def timed(w, reps=4):
    _, _, integ, film = w
    ms = []
    for _ in range(reps + 1):
        film.clear()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); integ.render(film, stream=st); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms[1:]))


def arg(i, dflt):
    return [int(x) for x in (sys.argv[i] if len(sys.argv) > i else dflt).split(",")]


for r, q, lq in itertools.product(arg(1, "4,8,12,16,20,24,28"), arg(2, "4,8,12,16,24"), arg(3, "4,8,12,18,24,33")):
    pb2.check(L.pb2_set_trace_tuning(r, q, lq, 0))
    t2, t4 = timed(c2), timed(c4)
    print(f"refill<{r:2d} node_q {q:2d} leaf_q {lq:2d}: c2 {t2:.3f} ms c4 {t4:.3f} ms  rgb {c2[3].resolve_rgb().mean():.6f} {c4[3].resolve_rgb().mean():.6f}", flush=True)

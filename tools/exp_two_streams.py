#!/usr/bin/env python
"""Experiment: how much of a path-traced frame is tail / ramp / launch gap that a second, independent batch could fill?
Two copies of the C4 scene (own BVH, own wavefront arena, own film) render 1920x1080 @ 8 spp each (one batch = 7 launches per
bounce), first one after the other on one stream, then at the same time on two streams.  The persistent traversal kernels fill
the machine, so a second stream can only run in the holes of the first: the ratio is the upper bound of what interleaving the
batches of ONE frame over two streams could gain."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

pb2, scenes = ge.load_package(), ge.load_scenes()
pb2.init(0)
SPP = int(os.environ.get("EXP_SPP", "8"))
which = os.environ.get("EXP_SCENE", "c4")
sc, cam, pk = (scenes.scene_c4(), scenes.C4_CAMERA, scenes.C4_PATH) if which == "c4" else (scenes.scene_c2(), scenes.C2_CAMERA, scenes.C2_PATH)


def setup():
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    return accel, camera, pb2.PathIntegrator(accel, camera, **dict(pk, spp=SPP)), pb2.Film(cam["res"])


A, B = setup(), setup()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for w, s in ((A, s1), (B, s2)):
    w[2].render(w[3], stream=s.cuda_stream)
torch.cuda.synchronize()


def run(concurrent):
    ts = []
    for _ in range(5):
        A[3].clear(); B[3].clear()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(s1)
        s2.wait_event(e0)
        A[2].render(A[3], stream=s1.cuda_stream)
        if concurrent:
            B[2].render(B[3], stream=s2.cuda_stream)
            e2.record(s2)
            s1.wait_event(e2)
        else:
            B[2].render(B[3], stream=s1.cuda_stream)
        e1.record(s1)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts[1:]))


for rep in range(2):
    seq, con = run(False), run(True)
    print(f"{which} 2 x {SPP} spp: one stream {seq:.3f} ms, two streams {con:.3f} ms ({(con / seq - 1) * 100:+.1f} %)  "
          f"rgb {A[3].resolve_rgb().mean():.6f} {B[3].resolve_rgb().mean():.6f}", flush=True)

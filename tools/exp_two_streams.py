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


NC = int(os.environ.get("EXP_COPIES", "2"))
W = [setup() for _ in range(NC)]
S = [torch.cuda.Stream() for _ in range(NC)]
for w, s in zip(W, S):
    w[2].render(w[3], stream=s.cuda_stream)
torch.cuda.synchronize()


def run(concurrent):
    ts = []
    for _ in range(5):
        for w in W:
            w[3].clear()
        torch.cuda.synchronize()
        e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        e0.record(S[0])
        for s in S[1:]:
            s.wait_event(e0)
        for w, s in zip(W, S):
            w[2].render(w[3], stream=(s if concurrent else S[0]).cuda_stream)
        if concurrent:
            for s in S[1:]:
                e = torch.cuda.Event()
                e.record(s)
                S[0].wait_event(e)
        e1.record(S[0])
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts[1:]))


for rep in range(2):
    seq, con = run(False), run(True)
    print(f"{which} {NC} x {SPP} spp: one stream {seq:.3f} ms, {NC} streams {con:.3f} ms ({(con / seq - 1) * 100:+.1f} %)  "
          f"rgb {W[0][3].resolve_rgb().mean():.6f} {W[-1][3].resolve_rgb().mean():.6f}", flush=True)

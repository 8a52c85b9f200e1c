#!/usr/bin/env python
"""Sweep the traversal scheduling knobs on the C3 ray sets (scene built once); prints ms per launch for each setting."""
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
verts, idx = scenes.scene_c3(2237)
accel = pb2.BVHAccel(verts, idx, max_prims_in_node=4)
cam = dict(scenes.C3_CAMERA, res=(1024, 1024))
camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
n = 1024 * 1024
dev = torch.device("cuda", 0)
buf = lambda b: torch.empty(b, dtype=torch.uint8, device=dev)
d_rays, d_hits, d_b0, d_s, d_b, d_occ, d_bh = buf(n * 32), buf(n * 16), buf(n * 4), buf(n * 32), buf(n * 32), buf(n), buf(n * 16)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
camera.primary_rays_device(d_rays.data_ptr(), st)
accel.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_b0.data_ptr(), st)
accel.spawn_shadow_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, scenes.C3_POINT_LIGHT, d_s.data_ptr(), st)
accel.spawn_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, d_b.data_ptr(), st)
torch.cuda.synchronize()


def timed(fn, reps=8):
    ms = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms[2:]))


refills = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "8,12,16,20,24".split(","))]
nodeqs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "6,10,12,16,20".split(","))]
prefs = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "0".split(","))]
leafqs = [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else "33".split(","))]
for r, q, p, lq in itertools.product(refills, nodeqs, prefs, leafqs):
    pb2.check(L.pb2_set_trace_tuning(r, q, lq, p))
    t1 = timed(lambda: accel.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_b0.data_ptr(), st))
    t2 = timed(lambda: accel.intersect_p_device(d_s.data_ptr(), n, d_occ.data_ptr(), st))
    t3 = timed(lambda: accel.intersect_device(d_b.data_ptr(), n, d_bh.data_ptr(), None, st))
    print(f"refill<{r:2d} node_q {q:2d} leaf_q {lq:2d} prefetch {p}: primary {t1:.4f} shadow {t2:.4f} bounce {t3:.4f} sum {t1+t2+t3:.4f} ms", flush=True)

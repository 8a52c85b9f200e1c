#!/usr/bin/env python
"""Experiment: does ordering the incoherent C3 bounce batch help k_closest_hit?  The 1 M cosine-bounce
rays (pixel order, random directions) are re-ordered on the device with torch — by direction octant (stable: pixel order inside an
octant), by octant then 30-bit Morton code of the origin, by Morton code alone, and at random — and k_closest_hit is timed on
every ordering (L2 flushed between launches).  Results are scattered back by ray index and compared with the unsorted launch,
so every ordering is checked to be result-neutral.  The sort itself is timed separately (torch.sort = cub radix sort)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

pb2, scenes = ge.load_package(), ge.load_scenes()
pb2.init(0)
verts, idx = scenes.scene_c3(int(os.environ.get("EXP_GRID", "2237")))
accel = pb2.BVHAccel(verts, idx, max_prims_in_node=4)
cam = dict(scenes.C3_CAMERA, res=(1024, 1024))
camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
n = 1024 * 1024
dev = torch.device("cuda", 0)
buf = lambda b: torch.empty(b, dtype=torch.uint8, device=dev)
d_rays, d_hits, d_b0, d_b, d_bh = buf(n * 32), buf(n * 16), buf(n * 4), buf(n * 32), buf(n * 16)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
camera.primary_rays_device(d_rays.data_ptr(), st)
accel.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_b0.data_ptr(), st)
accel.spawn_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, d_b.data_ptr(), st)
torch.cuda.synchronize()


def timed(fn, reps=10):
    ms = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms[2:]))


rays = d_b.view(torch.float32).view(n, 8)
o, d, tmax = rays[:, 0:3], rays[:, 4:7], rays[:, 3]
live = tmax >= 0                                      # primary misses spawn a degenerate ray (t_max = -1)
octant = ((d[:, 0] < 0).long() | ((d[:, 1] < 0).long() << 1) | ((d[:, 2] < 0).long() << 2))
lo, hi = o[live].min(0).values, o[live].max(0).values
q = ((o - lo) / (hi - lo).clamp_min(1e-20) * 1023.0).clamp(0, 1023).long()


def spread(v):
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


morton = (spread(q[:, 2]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 0])
dead = (~live).long() << 40
pix = torch.arange(n, device=dev)
px, py = pix % 1024, pix // 1024


def tile_key(tw, th):
    """Rays of one tw x th pixel tile adjacent (tiles row-major, pixels row-major inside a tile): a warp then covers a compact
    block of the image instead of 32 pixels of one row."""
    return ((py // th) * (1024 // tw) + (px // tw)) * (tw * th) + (py % th) * tw + (px % tw)


orders = {
    "pixel order (as spawned)": None,
    "8x4 pixel tiles": tile_key(8, 4) + dead,
    "8x8 pixel tiles": tile_key(8, 8) + dead,
    "16x16 tiles, octant inside": ((tile_key(16, 16) // 256) << 11) + (octant << 8) + (tile_key(16, 16) % 256) + dead,
    "pixel morton (z-order)": (spread(py) << 1 | spread(px)) + dead,
    "octant (stable)": octant + dead,
    "octant, morton30(origin)": (octant << 30) + morton + dead,
    "morton30(origin)": morton + dead,
    "morton30(origin), octant": (morton << 3) + octant + dead,
    "random permutation": torch.randperm(n, device=dev),
}
ref = None
d_sorted = buf(n * 32)
for name, key in orders.items():
    if key is None:
        perm, sort_ms = None, 0.0
        src = d_b
    else:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        perm = torch.sort(key, stable=True).indices
        d_sorted.view(torch.float32).view(n, 8).copy_(rays[perm])
        b.record()
        torch.cuda.synchronize()
        sort_ms = a.elapsed_time(b)
        src = d_sorted
    ms = timed(lambda: accel.intersect_device(src.data_ptr(), n, d_bh.data_ptr(), None, st))
    hits = d_bh.view(torch.int32).view(n, 4).clone()
    if perm is not None:
        un = torch.empty_like(hits)
        un[perm] = hits
        hits = un
    if ref is None:
        ref = hits
    same = bool(torch.equal(ref, hits))
    print(f"{name:32s} k_closest_hit {ms:.4f} ms = {n / ms / 1e3:7.1f} Mrays/s   (torch sort + gather {sort_ms:.3f} ms)   results equal: {same}", flush=True)

# the primary (coherent) set under the same pixel-tile orders: does a compact warp footprint help camera rays?
prim_rays = d_rays.view(torch.float32).view(n, 8)
for name, key in (("row-major (as generated)", None), ("8x4 pixel tiles", tile_key(8, 4)), ("8x8 pixel tiles", tile_key(8, 8)),
                  ("pixel morton (z-order)", spread(py) << 1 | spread(px))):
    if key is None:
        src = d_rays
    else:
        perm = torch.sort(key, stable=True).indices
        d_sorted.view(torch.float32).view(n, 8).copy_(prim_rays[perm])
        src = d_sorted
    ms = timed(lambda: accel.intersect_device(src.data_ptr(), n, d_bh.data_ptr(), None, st))
    print(f"primary set, {name:28s} k_closest_hit {ms:.4f} ms = {n / ms / 1e3:7.1f} Mrays/s", flush=True)

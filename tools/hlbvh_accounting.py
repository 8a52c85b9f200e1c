#!/usr/bin/env python
"""HLBVH accounting: where the wall time of a GPU HLBVH build goes outside its six device stages, how
stable the stages are over repeated builds, and why the C3 pass is slower on the HLBVH tree than on the SAH tree (oracle-counted
boxes and triangles tested per ray on both trees, reference traversal order, plus the device timings of the same ray sets)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

pb2, scenes, orc = ge.load_package(), ge.load_scenes(), ge.load_oracle()
pb2.init(0)
n_grid = int(os.environ.get("ACC_GRID", "2237"))
verts, idx = scenes.scene_c3(n_grid)
print(f"scene: {len(idx)} triangles, {len(verts)} vertices ({verts.nbytes / 1e6:.0f} + {idx.nbytes / 1e6:.0f} MB)", flush=True)
names = ["upload_bounds_morton", "sort", "treelets", "upper_sah_host", "flatten", "device_repack"]
for rep in range(5):
    t0 = time.perf_counter()
    sc = pb2.Scene(verts, idx)
    t1 = time.perf_counter()
    accel = pb2.BVHAccel(sc, max_prims_in_node=4, split_method=1)
    t2 = time.perf_counter()
    st = accel.build_stats()
    print(f"build {rep}: pb2_scene_create (host copy + validation) {1e3 * (t1 - t0):7.1f} ms | pb2_scene_build_bvh {1e3 * (t2 - t1):7.1f} ms, of which stages "
          f"{sum(st):6.1f} ms: " + ", ".join(f"{n} {v:.2f}" for n, v in zip(names, st)), flush=True)
    if rep < 4:
        del accel, sc
t0 = time.perf_counter()
sc_s = pb2.Scene(verts, idx)
t1 = time.perf_counter()
accel_s = pb2.BVHAccel(sc_s, max_prims_in_node=4, split_method=0)
t2 = time.perf_counter()
print(f"SAH: pb2_scene_create {1e3 * (t1 - t0):.1f} ms | pb2_scene_build_bvh (host SAH + repack + upload) {1e3 * (t2 - t1):.1f} ms", flush=True)

# the C3 ray sets
cam = dict(scenes.C3_CAMERA, res=(1024, 1024))
camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
n = 1024 * 1024
dev = torch.device("cuda", 0)
buf = lambda b: torch.empty(b, dtype=torch.uint8, device=dev)
d_rays, d_hits, d_b0, d_s, d_b, d_occ = buf(n * 32), buf(n * 16), buf(n * 4), buf(n * 32), buf(n * 32), buf(n)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
camera.primary_rays_device(d_rays.data_ptr(), st)
accel_s.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_b0.data_ptr(), st)
accel_s.spawn_shadow_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, scenes.C3_POINT_LIGHT, d_s.data_ptr(), st)
accel_s.spawn_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, d_b.data_ptr(), st)
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


g = {k: v.view(torch.float32).cpu().numpy().reshape(-1, 8) for k, v in (("primary", d_rays), ("shadow", d_s), ("bounce", d_b))}
stride = int(os.environ.get("ACC_STRIDE", "8"))
for tree, acc, split in (("SAH", accel_s, 0), ("HLBVH", accel, 1)):
    nn, npr, depth = acc.info()
    ref = orc.BVHAccel(verts, idx, 4, split_method=split)
    row = [f"{tree}: {nn} nodes, depth {depth}"]
    for name, rays_dev, any_hit in (("primary", d_rays, False), ("shadow", d_s, True), ("bounce", d_b, False)):
        if any_hit:
            ms = timed(lambda: acc.intersect_p_device(rays_dev.data_ptr(), n, d_occ.data_ptr(), st))
            cnt = ref.intersect_p(np.ascontiguousarray(g[name][::stride]), counters=True)[1]
        else:
            ms = timed(lambda: acc.intersect_device(rays_dev.data_ptr(), n, d_hits.data_ptr(), None, st))
            cnt = ref.intersect(np.ascontiguousarray(g[name][::stride]), counters=True)[1]
        m = len(g[name][::stride])
        row.append(f"{name} {ms:.3f} ms = {n / ms / 1e3:6.0f} Mrays/s, {cnt[0] / m:5.1f} boxes + {cnt[1] / m:4.2f} triangles per ray")
    print(" | ".join(row), flush=True)

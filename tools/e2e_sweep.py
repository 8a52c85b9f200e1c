#!/usr/bin/env python
"""e2e (host-buffer) throughput of pb2_intersect / pb2_intersect_p on the C3 ray sets for one setting of the ring's chunk
sizes (env PB2_PIPE_CHUNK / PB2_PIPE_TAIL, read when the library loads).  Prints one line; run once per setting."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    import torch
    pb2, scenes = ge.load_package(), ge.load_scenes()
    pb2.init(0)
    grid = int(os.environ.get("GRID_N", "2237"))
    verts, idx = scenes.scene_c3(grid)
    accel = pb2.BVHAccel(verts, idx, max_prims_in_node=4)
    cam = scenes.C3_CAMERA
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], (1024, 1024))
    n = 1024 * 1024
    dev = torch.device("cuda", 0)
    d = [torch.empty(n * 32, dtype=torch.uint8, device=dev) for _ in range(3)]
    d_hits = torch.empty(n * 16, dtype=torch.uint8, device=dev)
    camera.primary_rays_device(d[0].data_ptr(), None)
    accel.intersect_device(d[0].data_ptr(), n, d_hits.data_ptr(), None, None)
    accel.spawn_shadow_rays_device(d[0].data_ptr(), d_hits.data_ptr(), n, scenes.C3_POINT_LIGHT, d[1].data_ptr(), None)
    accel.spawn_bounce_rays_device(d[0].data_ptr(), d_hits.data_ptr(), n, d[2].data_ptr(), None)
    torch.cuda.synchronize()
    h = [torch.empty(n * 8, dtype=torch.float32).pin_memory() for _ in range(3)]
    for a, b in zip(h, d):
        a.copy_(b.view(torch.float32))
    h_hits = [torch.empty(n * 4, dtype=torch.int32).pin_memory() for _ in range(2)]
    h_occ = torch.empty(n, dtype=torch.uint8).pin_memory()
    L = pb2.lib()

    use_async = os.environ.get("E2E_ASYNC") == "1"

    def step():
        if use_async:                       # three batches in flight, one wait (what bench.py's e2e times)
            pb2.check(L.pb2_intersect_async(accel.h, h[0].data_ptr(), n, h_hits[0].data_ptr(), None))
            pb2.check(L.pb2_intersect_p_async(accel.h, h[1].data_ptr(), n, h_occ.data_ptr()))
            pb2.check(L.pb2_intersect_async(accel.h, h[2].data_ptr(), n, h_hits[1].data_ptr(), None))
            pb2.check(L.pb2_scene_wait(accel.h))
            return
        pb2.check(L.pb2_intersect(accel.h, h[0].data_ptr(), n, h_hits[0].data_ptr(), None))
        pb2.check(L.pb2_intersect_p(accel.h, h[1].data_ptr(), n, h_occ.data_ptr()))
        pb2.check(L.pb2_intersect(accel.h, h[2].data_ptr(), n, h_hits[1].data_ptr(), None))

    for _ in range(3):
        step()
    best = 1e9
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        t1 = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t1)
    dt = (time.perf_counter() - t0) / reps
    import zlib
    crc = "%08x" % zlib.crc32(h_occ.numpy().tobytes(), zlib.crc32(h_hits[1].numpy().tobytes(), zlib.crc32(h_hits[0].numpy().tobytes())))
    print(f"async={int(use_async)} chunk={os.environ.get('PB2_PIPE_CHUNK', 'default')} tail={os.environ.get('PB2_PIPE_TAIL', 'default')} "
          f"e2e mean {3 * n / dt / 1e6:.0f} Mrays/s best {3 * n / best / 1e6:.0f} Mrays/s ({dt * 1e3:.3f} ms/step) crc {crc}", flush=True)


if __name__ == "__main__":
    main()

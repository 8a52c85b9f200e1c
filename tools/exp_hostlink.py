#!/usr/bin/env python
"""Host-link experiment for the e2e (host-buffer) path at N ranks.  Launch with torchrun, one rank per GPU.
Every rank measures, with all ranks running at the same moment (barrier before each leg, max / sum over ranks afterwards):
  1. a plain pinned H2D and D2H copy of one ray-set-sized buffer, alone on the box (rank by rank) and all ranks together;
  2. the same H2D from write-combined pinned memory (pb2_host_alloc with PB2_HOST_ALLOC_WC=1);
  3. the e2e C3 step (pb2_intersect_async x2 + pb2_intersect_p_async + pb2_scene_wait) with the process pinned to its own slice
     of the GPU's CPU affinity mask (sched_setaffinity before the pinned buffers are allocated) and without pinning.
Prints one line per leg on rank 0."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
if os.environ.get("EXP_BIND") == "1":            # one contiguous slice of the visible cores per rank, before any allocation
    cores = sorted(os.sched_getaffinity(0))
    per = max(1, len(cores) // world)
    os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pb2, scenes = ge.load_package(), ge.load_scenes()
pb2.init(local)
dev = torch.device("cuda", local)
n = 1024 * 1024
L = pb2.lib()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def reduce(v, op):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=op)
    return float(t[0])


def say(msg):
    if rank == 0:
        print(msg, flush=True)


say(f"ranks {world}, cores visible to rank 0: {len(os.sched_getaffinity(0))}, bind={os.environ.get('EXP_BIND', '0')}, wc={os.environ.get('PB2_HOST_ALLOC_WC', '0')}")
d_buf = torch.empty(n * 32, dtype=torch.uint8, device=dev)
h_pin = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
h_pin.fill_(1)
ptr = C.c_void_p()
pb2.check(L.pb2_host_alloc(n * 32, C.byref(ptr)))          # write-combined when PB2_HOST_ALLOC_WC=1
C.memset(ptr, 1, n * 32)


def copy_gbs(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return n * 32 * reps / (a.elapsed_time(b) * 1e-3) / 1e9


def h2d():
    d_buf.copy_(h_pin, non_blocking=True)


def d2h():
    h_pin.copy_(d_buf, non_blocking=True)


def h2d_api():
    pb2.check(L.pb2_memcpy_h2d(C.c_void_p(d_buf.data_ptr()), ptr, n * 32))


# alone: rank by rank
alone = {}
for r in range(world):
    barrier()
    if r == rank:
        alone = {"h2d": copy_gbs(h2d), "d2h": copy_gbs(d2h), "h2d_api_buf": copy_gbs(h2d_api, 10)}
barrier()
for k in ("h2d", "d2h", "h2d_api_buf"):
    lo, hi = reduce(alone[k], dist.ReduceOp.MIN if world > 1 else None), reduce(alone[k], dist.ReduceOp.MAX if world > 1 else None)
    say(f"alone      {k:12s}: {lo:6.1f} .. {hi:6.1f} GB/s per rank")
# together
for k, fn in (("h2d", h2d), ("d2h", d2h), ("h2d_api_buf", h2d_api)):
    barrier()
    v = copy_gbs(fn, 10 if k == "h2d_api_buf" else 20)
    lo, tot = reduce(v, dist.ReduceOp.MIN if world > 1 else None), reduce(v, dist.ReduceOp.SUM if world > 1 else None)
    say(f"concurrent {k:12s}: min {lo:6.1f} GB/s per rank, {tot:7.1f} GB/s aggregate -> link bound {tot * 1e9 / 32 / 1e6:8.0f} Mrays/s")

# the e2e C3 step
verts, idx = scenes.scene_c3(int(os.environ.get("GRID_N", "2237")))
accel = pb2.BVHAccel(verts, idx, max_prims_in_node=4)
cam = scenes.C3_CAMERA
camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], (1024, 1024))
d = [torch.empty(n * 32, dtype=torch.uint8, device=dev) for _ in range(3)]
d_hits = torch.empty(n * 16, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
camera.primary_rays_device(d[0].data_ptr(), st)
accel.intersect_device(d[0].data_ptr(), n, d_hits.data_ptr(), None, st)
accel.spawn_shadow_rays_device(d[0].data_ptr(), d_hits.data_ptr(), n, scenes.C3_POINT_LIGHT, d[1].data_ptr(), st)
accel.spawn_bounce_rays_device(d[0].data_ptr(), d_hits.data_ptr(), n, d[2].data_ptr(), st)
torch.cuda.synchronize()
hptr = []
for k in range(3):                                          # ray buffers from pb2_host_alloc (WC when asked), results from torch pinned memory
    p = C.c_void_p()
    pb2.check(L.pb2_host_alloc(n * 32, C.byref(p)))
    pb2.check(L.pb2_memcpy_d2h(p, C.c_void_p(d[k].data_ptr()), n * 32))
    hptr.append(p)
h_hits = [torch.empty(n * 4, dtype=torch.int32).pin_memory() for _ in range(2)]
h_occ = torch.empty(n, dtype=torch.uint8).pin_memory()


def step():
    pb2.check(L.pb2_intersect_async(accel.h, hptr[0], n, h_hits[0].data_ptr(), None))
    pb2.check(L.pb2_intersect_p_async(accel.h, hptr[1], n, h_occ.data_ptr()))
    pb2.check(L.pb2_intersect_async(accel.h, hptr[2], n, h_hits[1].data_ptr(), None))
    pb2.check(L.pb2_scene_wait(accel.h))


for _ in range(3):
    step()
barrier()
t0 = time.perf_counter()
for _ in range(20):
    step()
dt = time.perf_counter() - t0
dt = reduce(dt, dist.ReduceOp.MAX if world > 1 else None)
import zlib
crc = "%08x" % zlib.crc32(h_occ.numpy().tobytes(), zlib.crc32(h_hits[1].numpy().tobytes(), zlib.crc32(h_hits[0].numpy().tobytes())))
say(f"e2e C3 step: {world * 3 * n * 20 / dt / 1e6:8.0f} Mrays/s over {world} ranks ({1e3 * dt / 20:.3f} ms per step), hits_crc32 {crc}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

"""Multi-GPU test of the film reduce (needs >= 2 visible GPUs; on a single-GPU box it is skipped — the host-side logic is
covered on CPU by tests/test_multi_gpu_host.py).  Two processes, one per GPU, render disjoint sample ranges of the same
frame; pb2_film_reduce (ncclReduce over NVLink) on rank 0 must equal the sum of the two per-GPU films exactly (fixed order
for two ranks), and agree with a single-GPU render of the whole range up to f32 reassociation."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pb2, scenes = ge.load_package(), ge.load_scenes()
    pb2.init(rank)
    uid = [pb2.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    pb2.nccl_init(uid[0], rank, world)
    cam = dict(scenes.C2_CAMERA, res=(96, 96))
    kw = dict(max_depth=4, rr_threshold=1.0, light_strategy="uniform", spp=8)
    accel = pb2.BVHAccel(pb2.scene_from_dict(scenes.scene_c2()), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = pb2.PathIntegrator(accel, camera, **kw)
    film = pb2.Film(cam["res"])
    begin, end = pb2.partition_samples(kw["spp"], rank, world)
    integ.render(film, begin, end)
    np.save(os.path.join(out_dir, f"own{rank}.npy"), film.read_xyzw())
    film.reduce(0)
    pb2.check(pb2.lib().pb2_device_synchronize())
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), film.read_xyzw())
        whole = pb2.Film(cam["res"])
        integ.render(whole)
        np.save(os.path.join(out_dir, "whole.npy"), whole.read_xyzw())
    dist.barrier()
    pb2.nccl_shutdown()
    dist.destroy_process_group()


def test_two_gpu_film_reduce(gpu, tmp_path):
    n = gpu.__dict__["C"].c_int()
    gpu.check(gpu.lib().pb2_device_count(gpu.C.byref(n)))
    if n.value < 2:
        pytest.skip("needs 2 GPUs (single-GPU box); host-side N>1 logic is tested in test_multi_gpu_host.py")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    own0, own1 = np.load(tmp_path / "own0.npy"), np.load(tmp_path / "own1.npy")
    reduced, whole = np.load(tmp_path / "reduced.npy"), np.load(tmp_path / "whole.npy")
    assert np.array_equal(reduced, own0 + own1)
    assert np.array_equal(reduced[..., 3], whole[..., 3])
    assert np.allclose(reduced, whole, rtol=1e-6, atol=1e-7)

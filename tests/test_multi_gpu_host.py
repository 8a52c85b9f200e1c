"""CPU tests of the N > 1 path (host logic): world_size-2 `gloo` processes split the sample indices with
partition_samples(), each renders its range (with the oracle standing in for the device), and the films are summed with
one reduce to rank 0 — the same plumbing bench.py runs over NCCL.  The reduced film must equal a single-process render."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_samples_covers_range(pb2):
    for spp in (1, 7, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            ranges = [pb2.partition_samples(spp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pb2.partition_samples(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pb2, scenes = ge.load_package(), ge.load_scenes()
    ge.load_oracle()
    from oracle import oracle_path as OP
    cam = dict(scenes.C2_CAMERA, res=(24, 24))
    kw = dict(max_depth=3, spp=6)
    ref = OP.Scene(scenes.scene_c2(), 4)
    begin, end = pb2.partition_samples(kw["spp"], rank, world)
    xyzw, _ = ref.render(cam, OP.film_desc(cam["res"]), OP.path_desc(sample_begin=begin, sample_end=end, **kw), threads=1)
    t = torch.from_numpy(xyzw)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)            # the film reduce of SURVEY §8e (ncclReduce on the GPUs)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_film_reduce_equals_single_render(pb2, orc, scenes, tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle_path as OP
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    reduced = np.load(out)
    cam = dict(scenes.C2_CAMERA, res=(24, 24))
    kw = dict(max_depth=3, spp=6)
    ref = OP.Scene(scenes.scene_c2(), 4)
    fd = OP.film_desc(cam["res"])
    a, _ = ref.render(cam, fd, OP.path_desc(sample_begin=0, sample_end=3, **kw))
    b, _ = ref.render(cam, fd, OP.path_desc(sample_begin=3, sample_end=6, **kw))
    assert np.array_equal(reduced, a + b)                  # fixed reduction order for 2 ranks: exact
    whole, _ = ref.render(cam, fd, OP.path_desc(**kw))
    assert np.array_equal(reduced[..., 3], whole[..., 3])  # weights: exact
    assert np.allclose(reduced, whole, rtol=1e-6, atol=1e-7)   # sum of per-rank XYZ vs XYZ of the whole sum: f32 reassociation only

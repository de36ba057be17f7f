"""World-size-2 test of the sample-sharding logic on CPU (gloo): each rank renders its own RNG stream with the CPU
oracle standing in for the device, the un-normalised sums are reduced to rank 0 and divided by the total sample count --
the same protocol bench.py runs over NCCL (one reduce at frame end, nothing per bounce)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, ROOT)
    import oracle
    from bench import shard_samples, reduce_and_finalize
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = oracle.crt.Scene.staircase(0.1, 32, 5)
    nx, ny, ns_total = 48, 32, 8
    ns = shard_samples(ns_total, world)[rank]
    img, _ = oracle.render(scene, nx, ny, ns, 8, stream=rank)
    acc = torch.zeros(ny, nx, 4)
    acc[..., :3] = torch.from_numpy(img * np.float32(ns))  # un-normalised sums, float4 per pixel like the device buffer
    frame = reduce_and_finalize(acc, ns_total, rank)
    if rank == 0:
        np.save(out, frame.numpy())
    dist.destroy_process_group()


def test_two_rank_sample_sharding_gloo(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, 29531 + os.getpid() % 200, out), nprocs=2, join=True)
    got = np.load(out)
    scene = oracle.crt.Scene.staircase(0.1, 32, 5)
    a, _ = oracle.render(scene, 48, 32, 4, 8, stream=0)
    b, _ = oracle.render(scene, 48, 32, 4, 8, stream=1)
    want = (a * np.float32(4) + b * np.float32(4)) / np.float32(8)
    assert np.allclose(got[..., :3], want, rtol=1e-6, atol=1e-7)


def test_shard_samples():
    sys.path.insert(0, ROOT)
    from bench import shard_samples
    assert shard_samples(100, 1) == [100]
    assert shard_samples(1024, 8) == [128] * 8
    assert shard_samples(10, 4) == [3, 3, 2, 2] and sum(shard_samples(7, 8)) == 7

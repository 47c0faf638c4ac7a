"""CPU test of the N>1 path's host logic: two gloo ranks shard a pair list, each produces its
ranges' results, and the gather on rank 0 reassembles them in order (what bench.py does with NCCL)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, out_path):
    sys.path.insert(0, ROOT)
    from roborts_edu_slam_b200.sharding import contiguous_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = contiguous_range(n_items, rank, world)
    # stand-in for the per-pair results of this rank's shard: f(pair id)
    local = torch.tensor([[i, i * i % 97, 3.0 * i] for i in range(b, e)], dtype=torch.float64).reshape(-1, 3)
    counts = [contiguous_range(n_items, r, world) for r in range(world)]
    pad = max(c[1] - c[0] for c in counts)
    buf = torch.full((pad, 3), -1.0, dtype=torch.float64)
    buf[: e - b] = local
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    # device-side timing is reduced with MAX over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = torch.cat([gathered[r][: counts[r][1] - counts[r][0]] for r in range(world)])
        np.save(out_path, np.concatenate([full.numpy().ravel(), t.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    n_items = 11
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_items, out), nprocs=2, join=True)
    got = np.load(out)
    full, tmax = got[:-1].reshape(-1, 3), got[-1]
    assert full.shape == (n_items, 3)
    assert np.array_equal(full[:, 0], np.arange(n_items))
    assert np.array_equal(full[:, 2], 3.0 * np.arange(n_items))
    assert tmax == 11.0

"""World-size-2 gloo test of the N>1 host logic (visiondk_b200/sharding.py): sharding, the two exchanges and the
merge layout.  The per-shard search and the merge are played by the oracle here (no GPU in this container); on the
GPU box the same functions run over NCCL with the CUDA kernels (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import retrieval as R
from visiondk_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    g = R.l2_normalize(rng.standard_normal((1001, 64)).astype(np.float32))
    q = R.l2_normalize(rng.standard_normal((37, 64)).astype(np.float32))
    k = 10
    lo, hi = sharding.shard_bounds(len(g), world, rank)
    q_lo, q_hi = sharding.shard_bounds(len(q), world, rank)

    def local_search(q_all, kk):
        s, i = R.flat_ip_search(q_all.numpy(), g[lo:hi], kk, id_offset=lo)
        return torch.from_numpy(s), torch.from_numpy(i)

    def merge(ss, ii, kk):
        s, i = R.merge_topk([x.numpy() for x in ss], [x.numpy() for x in ii], kk)
        return torch.from_numpy(s), torch.from_numpy(i)

    s, i = sharding.sharded_search(torch.from_numpy(q[q_lo:q_hi]), sharding.shard_sizes(len(q), world), local_search, merge, k)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), s=s.numpy(), i=i.numpy())
    dist.destroy_process_group()


def test_two_rank_sharded_search_equals_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    g = R.l2_normalize(rng.standard_normal((1001, 64)).astype(np.float32))
    q = R.l2_normalize(rng.standard_normal((37, 64)).astype(np.float32))
    ref_s, ref_i = R.flat_ip_search(q, g, 10)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["i"], ref_i) and np.array_equal(z["s"], ref_s)


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 1000000):
        for world in (1, 2, 3, 8):
            b = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            assert max(sharding.shard_sizes(n, world)) - min(sharding.shard_sizes(n, world)) <= 1

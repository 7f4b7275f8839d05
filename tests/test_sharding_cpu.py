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
    # the exchange between gallery ranges (retrieval._Exchange) through the same collectives: rank sketches -> global k-th bound
    comm = sharding.DistComm()
    assert (comm.world, comm.rank) == (world, rank)
    ranks = R.sketch_ranks(k, world)
    sk = R.rank_sketch(R.canonical_scores(q, g[lo:hi]), ranks)
    all_sk = comm.all_gather(torch.from_numpy(sk))
    assert tuple(all_sk.shape) == (world,) + sk.shape and np.array_equal(all_sk[rank].numpy(), sk)
    bound = R.bound_from_sketches(all_sk.numpy(), ranks, k, np.full(len(q), -np.inf, np.float32))
    t = torch.tensor([float(rank), -float(rank)])
    comm.all_reduce_max_(t)
    assert t.tolist() == [float(world - 1), 0.0]
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), s=s.numpy(), i=i.numpy(), bound=bound)
    dist.destroy_process_group()


def test_two_rank_sharded_search_equals_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    g = R.l2_normalize(rng.standard_normal((1001, 64)).astype(np.float32))
    q = R.l2_normalize(rng.standard_normal((37, 64)).astype(np.float32))
    ref_s, ref_i = R.flat_ip_search(q, g, 10)
    best_single = np.maximum(*[R.rank_sketch(R.canonical_scores(q, g[slice(*sharding.shard_bounds(len(g), world, r))]), [10])[:, 0]
                               for r in range(world)])
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["i"], ref_i) and np.array_equal(z["s"], ref_s)
        # the bound every rank derived: never above the true global k-th score, never below the best single shard's k-th
        assert np.all(z["bound"] <= ref_s[:, -1]) and np.all(z["bound"] >= best_single)
        assert np.array_equal(z["bound"], np.load(tmp_path / "r0.npz")["bound"])


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 1000000):
        for world in (1, 2, 3, 8):
            b = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            assert max(sharding.shard_sizes(n, world)) - min(sharding.shard_sizes(n, world)) <= 1


@pytest.mark.parametrize("world,clustered", [(8, False), (8, True), (3, False), (1, False)])
def test_rank_sketch_protocol_in_numpy_on_local_shards(world, clustered):
    """The W-shard search protocol with the oracle playing every kernel and sharding.LocalShardGroup (one host thread per
    shard, barriers for collectives) playing the network: two gallery ranges, a rank-sketch exchange after each, candidates
    below the derived bound dropped, lists merged — must equal the unsharded search, and the bound must be valid and useful."""
    rng = np.random.default_rng(7)
    nq, ng, dim, k = 24, 4000, 32, 20
    q = R.l2_normalize(rng.standard_normal((nq, dim)).astype(np.float32))
    g = R.l2_normalize(rng.standard_normal((ng, dim)).astype(np.float32))
    if clustered:  # every query's neighbours sit in ONE shard
        per = ng // world
        for j in range(nq):
            a = (j % world) * per + 100 + (j // world) * 40
            g[a:a + 30] = R.l2_normalize(q[j] + 0.05 * rng.standard_normal((30, dim)).astype(np.float32))
    ref_s, ref_i = R.flat_ip_search(q, g, k)
    ranks = R.sketch_ranks(k, world)
    group = sharding.LocalShardGroup(world)

    def shard(comm):
        lo, hi = sharding.shard_bounds(ng, world, comm.rank)
        sc = R.canonical_scores(q, g[lo:hi])
        bound = np.full(nq, -np.inf, np.float32)
        bounds = []
        for end in (min(64, hi - lo), hi - lo):  # two ranges: a prefix, then everything
            sk = R.rank_sketch(sc[:, :end], ranks)
            all_sk = comm.all_gather(torch.from_numpy(sk)).numpy()
            bound = R.bound_from_sketches(all_sk, ranks, k, bound)
            bounds.append(bound)
        keep = sc >= bound[:, None]  # what the re-rank keeps: only candidates that can reach the global top-k
        s_, i_ = R.topk_from_scores(np.where(keep, sc, -np.inf), k, id_offset=lo)
        pad = ~np.isfinite(s_) | (s_ == R.FLT_LOWEST)
        s_[pad], i_[pad] = R.FLT_LOWEST, -1
        packed = comm.all_gather(torch.from_numpy(np.stack([s_.astype(np.float64), i_.astype(np.float64)])))
        ms, mi = R.merge_topk([packed[r, 0].numpy().astype(np.float32) for r in range(world)],
                              [packed[r, 1].numpy().astype(np.int64) for r in range(world)], k)
        return ms, mi, bounds, int(keep.sum())

    results = group.run(shard)
    kth = ref_s[:, -1]
    for ms, mi, bounds, _ in results:
        assert np.array_equal(mi, ref_i) and np.array_equal(ms, ref_s)
        assert all(np.all(b <= kth) for b in bounds) and np.all(bounds[1] >= bounds[0])
        assert np.array_equal(bounds[1], results[0][2][1])  # every shard derives the same bound
    kept = sum(r[3] for r in results)
    assert kept >= nq * k  # nothing needed was dropped ...
    if world == 8 and not clustered:
        assert kept <= 3 * nq * k  # ... and little else survives (each shard alone would keep its own k: 8 x)


def test_memmap_store_is_read_shard_by_shard(tmp_path):
    """visiondk_b200.cbir.memmap_shard: every rank maps only its rows of the reference's raw [N, D] embedding file
    (engine/cbir/evaluation.py:124-152); the shards tile the file exactly, with the ids the sharded search assigns."""
    from visiondk_b200.cbir import memmap_shard
    rng = np.random.default_rng(1)
    for dtype in (np.float16, np.float32):
        emb = rng.standard_normal((1001, 24)).astype(dtype)
        path = str(tmp_path / f"gallery.{np.dtype(dtype).name}")
        mm = np.memmap(path, shape=emb.shape, mode="w+", dtype=dtype)
        mm[:] = emb
        mm.flush()
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                view, lo, n = memmap_shard(path, 24, dtype, r, world)
                assert n == 1001 and (lo, lo + len(view)) == sharding.shard_bounds(1001, world, r)
                assert np.array_equal(view, emb[lo:lo + len(view)])
                rows.append(np.asarray(view))
            assert np.array_equal(np.concatenate(rows), emb)
    with pytest.raises(ValueError, match="whole number"):
        memmap_shard(path, 25, np.float32)

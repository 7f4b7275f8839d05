"""Host-side plumbing of the sharded CBIR search: one process per GPU, gallery rows sharded
contiguously, two small exchanges over torch.distributed (NCCL on GPUs; the same code runs over gloo in the CPU
tests).  The reference only replicates the index (`faiss.index_cpu_to_all_gpus`, engine/cbir/evaluation.py:159-162);
sharding is BASELINE config 4.  No arithmetic happens here: scoring and merging are the CUDA kernels' job.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous rows [lo, hi) of rank `rank`; sizes differ by at most one and cover [0, n) exactly."""
    return n * rank // world, n * (rank + 1) // world


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]


def all_gather_rows(local: torch.Tensor, sizes: List[int]) -> torch.Tensor:
    """Exchange 1: every rank extracted `sizes[rank]` query embeddings; returns all of them in rank order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    # ranks may hold one row more or less: pad to the largest shard so every backend sees equal shapes
    m = max(sizes)
    padded = local.contiguous()
    if padded.shape[0] < m:
        pad = torch.zeros((m - padded.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([padded, pad], dim=0)
    out = torch.empty((len(sizes) * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    if all(n == m for n in sizes):
        return out
    return torch.cat([out[r * m:r * m + n] for r, n in enumerate(sizes)], dim=0)


def all_gather_topk(scores: torch.Tensor, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exchange 2: per-shard top-k lists [nq,k] -> stacked [world, nq, k] on every rank (8 MB per rank at 10k x 100)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return scores.unsqueeze(0), ids.unsqueeze(0)
    world = dist.get_world_size()
    ss = [torch.empty_like(scores) for _ in range(world)]
    ii = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(ss, scores.contiguous())
    dist.all_gather(ii, ids.contiguous())
    return torch.stack(ss), torch.stack(ii)


def sharded_search(q_local: torch.Tensor, q_sizes: List[int], local_search: Callable, merge: Callable, k: int):
    """all-gather queries -> local_search(q_all, k) on this rank's shard -> all-gather lists -> merge(ss, ii, k)."""
    q_all = all_gather_rows(q_local, q_sizes)
    s, i = local_search(q_all, k)
    ss, ii = all_gather_topk(s, i)
    if ss.shape[0] == 1:
        return s, i
    return merge(ss, ii, k)

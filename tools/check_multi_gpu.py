"""Multi-rank parity check, run under torchrun on one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py

Backend: NCCL with one GPU per rank when the box has >= WORLD_SIZE GPUs; otherwise (one GPU) every rank uses cuda:0 and the
collectives run over gloo — the kernels, the sharding / DDP host logic and the merge are the same code either way
(tests/test_multi_rank_gpu.py launches this script in whichever mode the box allows).

1. sharded CBIR search (gallery rows split over the ranks, packed single all-gather, `vdk_topk_merge_packed`) is
   BIT-IDENTICAL to the oracle's unsharded exact top-k (ids and canonical scores), including exact-duplicate ties that
   straddle shards and a query whose candidate lists overflow on one shard only;
2. ranks that initialise their models from DIFFERENT seeds hold identical parameters after FaceTrainer's construction
   (rank 0's state is broadcast, as DistributedDataParallel does), stay identical over two DDP train steps, and the
   first step's loss on rank r equals the single-process loss of the broadcast model on that rank's half batch.
Prints one JSON line on rank 0; exit code != 0 on any mismatch.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    nccl = torch.cuda.device_count() >= world
    dev = torch.device("cuda", local if nccl else 0)
    torch.cuda.set_device(dev)
    if nccl:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    from visiondk_b200 import sharding
    from visiondk_b200.retrieval import FlatIPIndex, sharded_flat_search
    out = {"world": world, "backend": dist.get_backend()}

    # ---- 1. sharded retrieval vs the oracle ----
    nq, ng, dim, k = 257, 30011, 512, 100
    g = torch.Generator(device="cpu").manual_seed(3)
    gallery = torch.nn.functional.normalize(torch.randn(ng, dim, generator=g))
    gallery[7777] = gallery[123]  # exact duplicates: the tie rule (smaller id first) must survive the merge
    gallery[29000] = gallery[123]
    gallery[20000:23000] = gallery[77]  # 3000 copies of one row inside the LAST shard only: that rank resolves the overflow
    queries = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g))
    queries[5] = gallery[123]
    queries[9] = gallery[77]
    lo, hi = sharding.shard_bounds(ng, world, rank)
    index = FlatIPIndex(dim, dev, normalize=True, id_offset=lo)
    index.add(gallery[lo:hi].to(dev))
    q_lo, q_hi = sharding.shard_bounds(nq, world, rank)
    s, i = sharded_flat_search(index, queries[q_lo:q_hi].to(dev), sharding.shard_sizes(nq, world), k)
    index.check_status()
    if rank == 0:
        from oracle import retrieval as oret  # the checker, never the product path
        qn = oret.l2_normalize(queries.numpy())
        gn = oret.l2_normalize(gallery.numpy())
        ref_s, ref_i = oret.flat_ip_search(qn, gn, k)
        got_s, got_i = s.cpu().numpy(), i.cpu().numpy()
        out["retrieval_ids_equal"] = bool(np.array_equal(got_i, ref_i))
        out["retrieval_scores_equal"] = bool(np.array_equal(got_s.view(np.uint32), ref_s.view(np.uint32)))
    resolved = torch.tensor([float(index.wide_path_rows + index.exhaustive_rows)], device=dev)
    dist.all_reduce(resolved, op=dist.ReduceOp.MAX)
    out["overflow_resolved_on_some_rank"] = bool(resolved.item() > 0)

    # ---- 2. DDP: broadcast of rank 0's initial state, then two train steps ----
    from visiondk_b200.train import FaceTrainingModel, FaceTrainer
    cfg = {"backbone": {"timm-toy": {"pretrained": False, "image_size": 64, "feat_dim": 64, "depths": (1, 1, 2, 1),
                                     "dims": (64, 128, 128, 256)}},
           "head": {"arcface": {"feat_dim": 64, "num_class": 50, "margin_arc": 0.35, "margin_am": 0.0, "scale": 32}}}
    torch.manual_seed(11 + 1000 * rank)  # DIFFERENT initial weights per rank: only the broadcast can make them agree
    model = FaceTrainingModel(cfg).to(dev)

    def flat_state():
        return torch.cat([p.detach().float().reshape(-1) for p in model.parameters()] +
                         [b.detach().float().reshape(-1) for b in model.buffers()])

    def identical_everywhere(v):
        ref = v.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([1.0 if torch.equal(v, ref) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        return bool(same.item() == 1.0)

    out["init_differs_before_trainer"] = not identical_everywhere(flat_state())
    trainer = FaceTrainer(model, lr0=0.01, momentum=0.9, weight_decay=5e-4, label_smooth=0.1, layer_wise=True, warm_steps=0,
                          total_steps=100, use_ema=(rank == 0))
    out["init_identical_after_trainer"] = identical_everywhere(flat_state())
    gen = torch.Generator(device="cpu").manual_seed(21)
    xs = torch.randn(2, world * 8, 3, 64, 64, generator=gen)
    ys = torch.randint(0, 50, (2, world * 8), generator=gen)
    losses = []
    for step in range(2):
        x = xs[step, rank * 8:(rank + 1) * 8].to(dev)
        y = ys[step, rank * 8:(rank + 1) * 8].to(dev)
        losses.append(float(trainer.step(x, y)))
    flat = flat_state()
    out["ddp_params_identical_on_all_ranks"] = identical_everywhere(torch.cat([p.detach().float().reshape(-1) for p in model.parameters()]))
    out["ddp_finite"] = bool(torch.isfinite(flat).all().item())
    all_losses = [None] * world
    dist.all_gather_object(all_losses, losses)
    out["losses_per_rank"] = all_losses
    ok = True
    if rank == 0:
        ok = all(v for k_, v in out.items() if isinstance(v, bool))
        out["ok"] = ok
        print(json.dumps(out), flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

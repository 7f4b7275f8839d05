"""Multi-GPU parity check, run under torchrun on one node (NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py

1. sharded CBIR search (gallery rows split over the ranks, `vdk_topk_merge`) is BIT-IDENTICAL to the oracle's unsharded
   exact top-k (ids and canonical scores);
2. two DDP train steps (flat-gradient all-reduce(mean)) leave every rank with identical parameters, and the first
   step's loss on rank r equals the single-process loss on that rank's half batch.
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
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from visiondk_b200 import sharding
    from visiondk_b200.retrieval import FlatIPIndex, merge_topk
    out = {"world": world}

    # ---- 1. sharded retrieval vs the oracle ----
    nq, ng, dim, k = 257, 30011, 512, 100
    g = torch.Generator(device="cpu").manual_seed(3)
    gallery = torch.nn.functional.normalize(torch.randn(ng, dim, generator=g))
    gallery[7777] = gallery[123]  # exact duplicates: the tie rule (smaller id first) must survive the merge
    gallery[29000] = gallery[123]
    queries = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g))
    queries[5] = gallery[123]
    lo, hi = sharding.shard_bounds(ng, world, rank)
    index = FlatIPIndex(dim, dev, normalize=True, id_offset=lo)
    index.add(gallery[lo:hi].to(dev))
    q_lo, q_hi = sharding.shard_bounds(nq, world, rank)
    s, i = sharding.sharded_search(queries[q_lo:q_hi].to(dev), sharding.shard_sizes(nq, world), index.search_device, merge_topk, k)
    index.check_status()
    if rank == 0:
        from oracle import retrieval as oret  # the checker, never the product path
        qn = oret.l2_normalize(queries.numpy())
        gn = oret.l2_normalize(gallery.numpy())
        ref_s, ref_i = oret.flat_ip_search(qn, gn, k)
        got_s, got_i = s.cpu().numpy(), i.cpu().numpy()
        out["retrieval_ids_equal"] = bool(np.array_equal(got_i, ref_i))
        out["retrieval_scores_equal"] = bool(np.array_equal(got_s.view(np.uint32), ref_s.view(np.uint32)))

    # ---- 2. DDP train steps ----
    from visiondk_b200.train import FaceTrainingModel, FaceTrainer
    cfg = {"backbone": {"timm-toy": {"pretrained": False, "image_size": 64, "feat_dim": 64, "depths": (1, 1, 2, 1),
                                     "dims": (64, 128, 128, 256)}},
           "head": {"arcface": {"feat_dim": 64, "num_class": 50, "margin_arc": 0.35, "margin_am": 0.0, "scale": 32}}}
    torch.manual_seed(11)
    model = FaceTrainingModel(cfg).to(dev)
    trainer = FaceTrainer(model, lr0=0.01, momentum=0.9, weight_decay=5e-4, label_smooth=0.1, layer_wise=True, warm_steps=0,
                          total_steps=100, use_ema=(rank == 0))
    gen = torch.Generator(device="cpu").manual_seed(21)
    xs = torch.randn(2, world * 8, 3, 64, 64, generator=gen)
    ys = torch.randint(0, 50, (2, world * 8), generator=gen)
    losses = []
    for step in range(2):
        x = xs[step, rank * 8:(rank + 1) * 8].to(dev)
        y = ys[step, rank * 8:(rank + 1) * 8].to(dev)
        losses.append(float(trainer.step(x, y)))
    flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([1.0 if torch.equal(flat, ref) else 0.0], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out["ddp_params_identical_on_all_ranks"] = bool(same.item() == 1.0)
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

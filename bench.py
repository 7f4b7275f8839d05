#!/usr/bin/env python
"""bench.py — the hot-path benchmark (driver contract: one JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the CBIR eval half of the metric): 10 000 queries x 1 000 000 gallery
rows, 512-d, cosine top-100.  A "step" is one full search pass: L2-normalise + fp16 copy of the query block,
score/filter over the whole (local shard of the) gallery on tcgen05, select + canonical re-rank.  Under N>1
the gallery is sharded by rows (strong scaling: the 1M gallery is fixed), queries are all-gathered, every rank
searches its shard, and the per-shard top-k lists are all-gathered and merged (NCCL over NVLink only for those
two small exchanges).

`value`  : pairs/s with the query block and the index already resident in HBM.
`e2e`    : the same through the reference-facing call FlatIPIndex.search(numpy float32) — host query buffer in,
           host (scores, ids) out, both copies inside the timed region.
`roofline`: the dominant kernel (score_filter over the last gallery range) timed alone with CUDA events.
`cpu_baseline`: the oracle's reference formulation (256-query slices: q @ g.T + top-k, engine/cbir/evaluation.py
           :190-195 with faiss replaced by its published brute-force algorithm) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NQ, NG, DIM, K = 10000, 1000000, 512, 100


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nq", type=int, default=NQ)
    ap.add_argument("--ng", type=int, default=NG)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[j] for s in self.samples for j in range(4) if len(s) > 2 + j and s[2 + j] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU reference arm: the reference formulation on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_pass(q, g, k, slice_rows=256):
    """engine/cbir/evaluation.py:190-195: 256-query slices against the whole index, top-k per slice."""
    import torch
    outs = []
    for i in range(0, q.shape[0], slice_rows):
        s = q[i:i + slice_rows] @ g.t()
        outs.append(torch.topk(s, min(k, g.shape[0]), dim=1))
    return outs


def cpu_baseline(ng, dim, k, target_s=12.0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(5)
    g = torch.nn.functional.normalize(torch.randn(ng, dim, generator=gen))
    q = torch.nn.functional.normalize(torch.randn(256, dim, generator=gen))
    cpu_reference_pass(q[:64], g, k)  # warm-up
    t0 = time.perf_counter()
    cpu_reference_pass(q, g, k)
    t1 = time.perf_counter() - t0
    reps = max(1, min(40, int(target_s / max(t1, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu_reference_pass(q, g, k)
    dt = time.perf_counter() - t0
    pairs = reps * 256 * ng
    return {"value": pairs / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x (256 queries x {ng} gallery x {dim}-d fp32, q@g.T + top-{k}), {dt:.1f} s"}, (q, g)


def run_reference(args):
    """--impl reference: the reference's own CPU formulation (oracle port; faiss is not installable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ng, dim, k = args.ng, args.dim, args.k
    gen = torch.Generator().manual_seed(5)
    g = torch.nn.functional.normalize(torch.randn(ng, dim, generator=gen))
    q = torch.nn.functional.normalize(torch.randn(256, dim, generator=gen))
    # one step = a bounded sample of the workload: one 256-query slice against the full gallery
    for _ in range(max(1, args.warmup)):
        cpu_reference_pass(q, g, k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(q, g, k)
    dt = time.perf_counter() - t0
    value = args.steps * 256 * ng / dt
    line = {
        "impl": "reference", "metric": "query x gallery pairs/sec (cosine top-100, 512-d)", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CBIR eval: {args.nq} queries x {ng} gallery x {dim}-d, cosine top-{k}",
                   "reference_formulation": "256-query slices, q @ g.T + top-k (faiss IndexFlatIP restated)"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"each step = 256 queries x {ng} gallery (1/{max(1, args.nq // 256)} of the query block)"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C
    import torch
    import torch.distributed as dist
    from visiondk_b200 import _lib, build
    from visiondk_b200.retrieval import FlatIPIndex, PreparedRows, merge_topk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the hot path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    build.build()
    _lib.load()
    _lib.require_device()

    nq, ng, dim, k = args.nq, args.ng, args.dim, args.k
    # synthetic data (seeded; the same rows on every rank, each rank keeps its shard)
    lo = ng * rank // world
    hi = ng * (rank + 1) // world
    gen = torch.Generator(device=dev).manual_seed(5)
    g_all_rows = None
    chunks = []
    # generate the full gallery in chunks with one generator so that shards are slices of the same 1M rows
    step_rows = 125000
    for a in range(0, ng, step_rows):
        b = min(ng, a + step_rows)
        blk = torch.nn.functional.normalize(torch.randn(b - a, dim, device=dev, generator=gen))
        s0, s1 = max(a, lo), min(b, hi)
        if s1 > s0:
            chunks.append(blk[s0 - a:s1 - a].clone())
        del blk
    q_full = torch.nn.functional.normalize(torch.randn(nq, dim, device=dev, generator=gen))
    index = FlatIPIndex(dim, dev, normalize=True, id_offset=lo)
    for c in chunks:
        index.add(c)
    del chunks
    index._finalize()
    q_lo, q_hi = nq * rank // world, nq * (rank + 1) // world
    q_local = q_full[q_lo:q_hi].contiguous()
    q_sizes = [nq * (r + 1) // world - nq * r // world for r in range(world)]

    def step_device():
        if world > 1:
            # exchange 1: all-gather of the query embeddings (each rank extracted nq/world of them)
            parts = [torch.empty((n, dim), dtype=torch.float32, device=dev) for n in q_sizes]
            dist.all_gather(parts, q_local)
            q = torch.cat(parts, 0)
        else:
            q = q_local
        s, i = index.search_device(q, k)
        if world > 1:
            # exchange 2: all-gather of per-shard top-k, merged with the canonical tie rule
            ss = torch.empty((world,) + tuple(s.shape), dtype=s.dtype, device=dev)
            ii = torch.empty((world,) + tuple(i.shape), dtype=i.dtype, device=dev)
            dist.all_gather_into_tensor(ss, s)
            dist.all_gather_into_tensor(ii, i)
            s, i = merge_topk(ss, ii, k)
        return s, i

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        out = step_device()
    index.check_status()
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = nq * ng / (ms_step * 1e-3)

    # ---- e2e: host query buffer -> (scores, ids) on the host, through the faiss-shaped call ----
    q_host = torch.empty((nq, dim), dtype=torch.float32).pin_memory()
    q_host.copy_(q_full.cpu())
    s_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    i_host = torch.empty((nq, k), dtype=torch.int64).pin_memory()

    def step_e2e():
        qd = q_host[q_lo:q_hi].to(dev, non_blocking=True) if world > 1 else q_host.to(dev, non_blocking=True)
        if world > 1:
            parts = [torch.empty((n, dim), dtype=torch.float32, device=dev) for n in q_sizes]
            dist.all_gather(parts, qd)
            qd = torch.cat(parts, 0)
        s, i = index.search_device(qd, k)
        if world > 1:
            ss = torch.empty((world,) + tuple(s.shape), dtype=s.dtype, device=dev)
            ii = torch.empty((world,) + tuple(i.shape), dtype=i.dtype, device=dev)
            dist.all_gather_into_tensor(ss, s)
            dist.all_gather_into_tensor(ii, i)
            s, i = merge_topk(ss, ii, k)
        if rank == 0:
            s_host.copy_(s, non_blocking=True)
            i_host.copy_(i, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    clocks = sampler.stop() if sampler else None

    # ---- roofline: the dominant kernel alone (score/filter over the last gallery range) ----
    lib = _lib.load()
    roof = None
    if rank == 0:
        ngl = hi - lo
        plan = _lib.TopkPlan()
        _lib.check(lib.vdk_topk_plan_default(C.byref(plan), nq, ngl, dim, k), "plan")
        qp = PreparedRows(q_full, True)
        r_lo = plan.stage_end[plan.n_stages - 2] if plan.n_stages > 1 else 0
        dense = 1 if plan.n_stages == 1 else 0
        ws = index._ws  # thresholds of the last search are still in the workspace
        reps = 5

        def score_only():
            _lib.check(lib.vdk_score_range(C.byref(plan), qp.xh.data_ptr(), index._rows.xh.data_ptr(), r_lo, ngl, dense,
                                           ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "vdk_score_range")

        score_only()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            score_only()
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / reps
        flops = 2.0 * dim * nq * (ngl - r_lo)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops", 1590.0))
        ach = flops / (k_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "score_filter_kernel<false>", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; fp16 runs at the same tensor rate)" if peaks
                else "fallback 1590 TFLOP/s (B200_PROFILING.md)",
                "launch_ms": k_ms, "algorithmic_flops_per_launch": flops,
                "share_of_step": k_ms / ms_step}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_baseline(ng, dim, k)

    if rank == 0:
        n_stages = 4 if (hi - lo) > 262144 else (3 if (hi - lo) > 32768 else (2 if (hi - lo) > 4096 else 1))
        launches_per_step = 1 + 1 + 2 * n_stages + (1 if world > 1 else 0)
        line = {
            "metric": "query x gallery pairs/sec (cosine top-100, 512-d)", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16 (candidates) + f64 re-rank",
            "data": "synthetic",
            "config": {"workload": f"CBIR eval: {nq} queries x {ng} gallery x {dim}-d, cosine top-{k}, gallery sharded by rows over {world} GPU(s)",
                       "l2": "gallery (fp16 1.0 GB + fp32 2.0 GB per 1M rows) exceeds the 126 MB L2; no flush needed",
                       "exactness": "ids and scores bit-exact vs oracle/retrieval.py (canonical fp64 re-rank)"},
            "e2e": {"value": nq * ng / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * k * 12},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""World-size-2 parity of the N>1 paths ON THE GPU KERNELS (tools/check_multi_gpu.py under torchrun): the sharded search
equals the oracle bit for bit, rank 0's initial state is broadcast, DDP steps keep the ranks identical.

With >= 2 GPUs the ranks use NCCL, one GPU each (the product configuration).  On a one-GPU box both ranks share cuda:0 and the
collectives run over gloo (NCCL refuses two ranks on one device): kernels, sharding, packing, merge and the DDP host logic
are exactly the code the NCCL run executes."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_sharded_search_and_ddp_parity(lib):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, f"rc={r.returncode}\nstdout:\n{r.stdout[-3000:]}\nstderr:\n{r.stderr[-3000:]}"
    out = json.loads(lines[-1])
    assert out["ok"], out
    assert out["backend"] == ("nccl" if torch.cuda.device_count() >= 2 else "gloo")
    for key in ("retrieval_ids_equal", "retrieval_scores_equal", "overflow_resolved_on_some_rank", "init_differs_before_trainer",
                "init_identical_after_trainer", "ddp_params_identical_on_all_ranks", "ddp_finite"):
        assert out[key] is True, (key, out)

"""Golden vectors minted from the ORACLE (oracle/retrieval.py, oracle/convnext.py, oracle/vit.py) for the pieces whose arithmetic
lives in third-party code the reference does not vendor (faiss, timm): regression pins of the restatements and fixed inputs for
the GPU parity tests.  (Vectors minted from the reference's OWN modules are made by oracle/make_golden.py.)

    python oracle/make_golden_oracle.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import retrieval as R  # noqa: E402
from oracle.convnext import TimmWrapperOracle, randomize_ as rand_cnx  # noqa: E402
from oracle.vit import ViTWrapperOracle, randomize_ as rand_vit  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def retrieval():
    rng = np.random.default_rng(11)
    g = rng.standard_normal((3000, 128)).astype(np.float32)
    g[700] = g[5]      # exact duplicates: the tie rule (smaller id first)
    g[2999] = g[5]
    q = rng.standard_normal((33, 128)).astype(np.float32)
    q[2] = g[5]
    qn, gn = R.l2_normalize(q), R.l2_normalize(g)
    s, i = R.flat_ip_search(qn, gn, 20)
    np.savez_compressed(os.path.join(OUT, "retrieval_small.npz"), q=q, g=g, scores=s, ids=i, k=np.int32(20))


def backbones():
    torch.manual_seed(0)
    x = torch.randn(4, 3, 64, 64)
    cnx = rand_cnx(TimmWrapperOracle("toy", 64, 64, depths=(1, 1, 2, 1), dims=(64, 128, 128, 256)), seed=21).eval()
    vit = rand_vit(ViTWrapperOracle("toy", 64, 64, patch=16, dim=128, depth=2, heads=2), seed=22).eval()
    with torch.no_grad():
        e_cnx = torch.nn.functional.normalize(cnx(x)).numpy()
        e_vit = torch.nn.functional.normalize(vit(x)).numpy()
    # the networks themselves are rebuilt from their seeds (oracle.*.randomize_ is deterministic); only inputs / outputs are stored
    np.savez_compressed(os.path.join(OUT, "backbones_toy.npz"), x=x.numpy(), convnext_seed=np.int32(21), vit_seed=np.int32(22),
                        convnext_embeddings=e_cnx, vit_embeddings=e_vit)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    retrieval()
    backbones()
    print("oracle golden vectors written to", OUT)

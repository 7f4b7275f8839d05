"""The oracle restatements of the third-party arithmetic (faiss flat-IP search, timm ConvNeXt / ViT) against the committed vectors
of oracle/make_golden_oracle.py: regression pins (a change of the oracle's canonical order or of a layer definition shows here)."""
import os

import numpy as np
import torch

from oracle import retrieval as R
from oracle.convnext import TimmWrapperOracle, randomize_ as rand_cnx
from oracle.vit import ViTWrapperOracle, randomize_ as rand_vit

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_retrieval_oracle_reproduces_golden_bit_for_bit():
    z = np.load(os.path.join(GOLD, "retrieval_small.npz"))
    s, i = R.flat_ip_search(R.l2_normalize(z["q"]), R.l2_normalize(z["g"]), int(z["k"]))
    assert np.array_equal(i, z["ids"])
    assert np.array_equal(s.view(np.uint32), z["scores"].view(np.uint32))
    # the planted duplicates: equal canonical scores, ascending ids
    assert list(i[2][:3]) == [5, 700, 2999] and s[2][0] == s[2][1] == s[2][2]


def test_backbone_oracles_reproduce_golden_embeddings():
    z = np.load(os.path.join(GOLD, "backbones_toy.npz"))
    x = torch.from_numpy(z["x"])
    cnx = rand_cnx(TimmWrapperOracle("toy", 64, 64, depths=(1, 1, 2, 1), dims=(64, 128, 128, 256)), seed=int(z["convnext_seed"])).eval()
    vit = rand_vit(ViTWrapperOracle("toy", 64, 64, patch=16, dim=128, depth=2, heads=2), seed=int(z["vit_seed"])).eval()
    with torch.no_grad():
        e_cnx = torch.nn.functional.normalize(cnx(x)).numpy()
        e_vit = torch.nn.functional.normalize(vit(x)).numpy()
    # fp32 on CPU: the summation order of the BLAS kernels may differ between hosts / thread counts
    np.testing.assert_allclose(e_cnx, z["convnext_embeddings"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(e_vit, z["vit_embeddings"], rtol=0, atol=2e-5)

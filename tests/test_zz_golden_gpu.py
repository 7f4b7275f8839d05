"""The CUDA path against the COMMITTED golden vectors (tests/golden/retrieval_small.npz, backbones_toy.npz: minted by
oracle/make_golden_oracle.py, pinned on CPU by tests/test_oracle_golden_cpu.py).  Named to run last: the same kernels are already
checked against the live oracle by the other GPU tests; this one fixes the inputs and the expected outputs in the repository."""
import os

import numpy as np
import pytest
import torch

from oracle.convnext import TimmWrapperOracle, randomize_ as rand_cnx
from oracle.vit import ViTWrapperOracle, randomize_ as rand_vit
from visiondk_b200.backbone import TimmWrapper
from visiondk_b200.retrieval import FlatIPIndex
from visiondk_b200.vit import ViTWrapper

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_search_reproduces_the_golden_top_k_bit_for_bit(lib):
    z = np.load(os.path.join(GOLD, "retrieval_small.npz"))
    index = FlatIPIndex(z["g"].shape[1], "cuda", normalize=True)
    index.add(z["g"])
    s, i = index.search(z["q"], int(z["k"]))
    index.check_status()
    assert np.array_equal(i, z["ids"])
    assert np.array_equal(s.view(np.uint32), z["scores"].view(np.uint32))


def test_backbones_reproduce_the_golden_embeddings(lib):
    """bf16 activations against the fp32 golden: relative L2 <= 3e-2, cosine >= 0.999 per row (the tolerance of the live-oracle tests)."""
    z = np.load(os.path.join(GOLD, "backbones_toy.npz"))
    x = torch.from_numpy(z["x"]).cuda()
    cnx_o = rand_cnx(TimmWrapperOracle("toy", 64, 64, depths=(1, 1, 2, 1), dims=(64, 128, 128, 256)), seed=int(z["convnext_seed"]))
    cnx = TimmWrapper("toy", 64, 64, pretrained=False, depths=(1, 1, 2, 1), dims=(64, 128, 128, 256))
    cnx.load_state_dict(cnx_o.state_dict(), strict=True)
    vit_o = rand_vit(ViTWrapperOracle("toy", 64, 64, patch=16, dim=128, depth=2, heads=2), seed=int(z["vit_seed"]))
    vit = ViTWrapper("toy", 64, 64, pretrained=False, patch=16, dim=128, depth=2, heads=2)
    vit.load_state_dict(vit_o.state_dict(), strict=True)
    for model, key in ((cnx, "convnext_embeddings"), (vit, "vit_embeddings")):
        got = model.cuda().eval().embed(x, l2_normalize=True).cpu()
        ref = torch.from_numpy(z[key])
        rel = ((got - ref).norm() / ref.norm()).item()
        cos = (got * ref).sum(dim=1).min().item()
        assert rel <= 3e-2 and cos >= 0.999, (key, rel, cos)

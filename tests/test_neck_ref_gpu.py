"""The CUDA TimmWrapper / ViTWrapper against tests/golden/neck_ref.npz — outputs of the REFERENCE's own
models/faceX/backbone/timm_wrapper.py:5-54 (both neck branches) executed around a stub `timm` body
(oracle/make_golden.py::neck; pinned on CPU by tests/test_oracle_neck_cpu.py).

Tolerance: the body runs in bf16 on the GPU against the fp32 golden -> embeddings relative L2 <= 3e-2 (the tolerance of
the live-oracle tests); neck gradients / running statistics relative L2 <= 6e-2 / 2e-2."""
import os

import numpy as np
import pytest
import torch

from oracle.convnext import TimmWrapperOracle, randomize_ as rand_cnx
from oracle.vit import ViTWrapperOracle, randomize_ as rand_vit
from visiondk_b200.backbone import TimmWrapper
from visiondk_b200.vit import ViTWrapper

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "neck_ref.npz")


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def build(tag, z):
    if tag == "cnn":
        o = rand_cnx(TimmWrapperOracle("toy", 64, 64, depths=(1, 1, 2, 1), dims=(64, 128, 128, 256)), seed=int(z["cnn_seed"]))
        m = TimmWrapper("toy", 64, 64, pretrained=False, depths=(1, 1, 2, 1), dims=(64, 128, 128, 256))
    else:
        o = rand_vit(ViTWrapperOracle("toy", 64, 64, patch=16, dim=128, depth=2, heads=2), seed=int(z["vit_seed"]))
        m = ViTWrapper("toy", 64, 64, pretrained=False, patch=16, dim=128, depth=2, heads=2)
    sd = o.state_dict()
    assert list(sd.keys()) == [str(k) for k in z[f"{tag}_keys"]]
    for k in sd:  # the neck tensors come from the golden file itself (the reference module's), the body from its seed
        if k.startswith("output_layer."):
            sd[k] = torch.from_numpy(z[f"{tag}_sd/{k}"])
    m.load_state_dict(sd, strict=True)  # the reference's key set loads with strict=True
    return m.cuda()


@pytest.mark.parametrize("tag", ["cnn", "vit"])
def test_eval_embeddings_match_the_reference_wrapper(lib, tag):
    z = np.load(GOLD)
    m = build(tag, z).eval()
    got = m.embed(torch.from_numpy(z["x"]).cuda(), l2_normalize=False).cpu()
    ref = torch.from_numpy(z[f"{tag}_eval"])
    assert rel(got, ref) <= 3e-2, rel(got, ref)
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1).min().item()
    assert cos >= 0.999, cos


@pytest.mark.parametrize("tag", ["cnn", "vit"])
def test_train_mode_neck_matches_the_reference_wrapper(lib, tag):
    z = np.load(GOLD)
    m = build(tag, z).train()
    y = m(torch.from_numpy(z["x"]).cuda())
    (y * torch.from_numpy(z[f"{tag}_w_out"]).cuda()).sum().backward()
    assert rel(y.detach().cpu(), torch.from_numpy(z[f"{tag}_train"])) <= 3e-2
    bn_scale = float(np.abs(z[f"{tag}_grad/output_layer.0.weight"]).max())
    for k, p in m.output_layer.named_parameters():
        ref = torch.from_numpy(z[f"{tag}_grad/output_layer.{k}"])
        g = p.grad.detach().cpu()
        if ref.norm() < 1e-5 * (1 + ref.numel() ** 0.5):  # exact gradient 0 (a shift in front of a batch-statistics BatchNorm)
            assert (g - ref).abs().max().item() <= 5e-2 * bn_scale + 1e-3, k
        else:
            assert rel(g, ref) <= 6e-2, (k, rel(g, ref))
    for k, b in m.output_layer.named_buffers():
        ref = torch.from_numpy(z[f"{tag}_after/output_layer.{k}"])
        if k.endswith("num_batches_tracked"):
            assert int(b) == int(ref)
        else:
            assert rel(b.detach().cpu().float(), ref.float()) <= 2e-2, (k, rel(b.detach().cpu().float(), ref.float()))

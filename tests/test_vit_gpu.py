"""GPU parity of the ViT extract path (csrc/vit.cu) against the oracle (oracle/vit.py: timm's VisionTransformer restated and
pinned against HF ViTModel on CPU).  bf16 activations vs the fp32 oracle: tolerances stated per test."""
import pytest
import torch

from oracle.vit import ViTWrapperOracle, randomize_
from visiondk_b200 import _lib
from visiondk_b200.backbone import BackboneFactory
from visiondk_b200.vit import ViTWrapper

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("B,N,H", [(2, 197, 12), (3, 64, 3), (1, 577, 4), (2, 50, 2), (2, 1, 1)])
def test_attention_forward_matches_torch(lib, B, N, H):
    """softmax(q k^T / 8) v per (image, head) on the qkv Linear's output layout [B, N, 3, H, 64]; reference in fp32 on the same
    bf16-rounded inputs.  P is rounded to bf16 before P.V: |err| <= 2^-8 relative to the largest |v| of the row."""
    torch.manual_seed(N * 7 + H)
    qkv = (torch.randn(B, N, 3, H, 64, device="cuda") * 1.5).to(torch.bfloat16)
    out = torch.full((B, N, H * 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.vdk_attention_fwd(qkv.data_ptr(), B, N, H, 64, out.data_ptr(), _lib.stream_ptr()), "attention")
    q, k, v = qkv.float().permute(2, 0, 3, 1, 4).unbind(0)  # [B, H, N, 64]
    ref = torch.softmax((q @ k.transpose(-2, -1)) * 0.125, dim=-1) @ v
    ref = ref.transpose(1, 2).reshape(B, N, H * 64)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 2e-2 * v.abs().max().item()
    assert rel(out, ref) <= 1e-2


def build(seed, **kw):
    oracle = randomize_(ViTWrapperOracle("x", **kw), seed=seed).eval()
    ours = ViTWrapper("x", kw["feat_dim"], kw["image_size"], pretrained=False, patch=kw["patch"], dim=kw["dim"], depth=kw["depth"],
                      heads=kw["heads"])
    ours.load_state_dict(oracle.state_dict(), strict=True)
    return oracle, ours.cuda().eval()


@pytest.mark.parametrize("cfg", [dict(feat_dim=64, image_size=64, patch=16, dim=128, depth=2, heads=2),
                                 dict(feat_dim=128, image_size=112, patch=14, dim=192, depth=3, heads=3)])
def test_vit_toy_embeddings_match_oracle(lib, cfg):
    oracle, ours = build(3, **cfg)
    torch.manual_seed(1)
    x = torch.randn(5, 3, cfg["image_size"], cfg["image_size"])
    with torch.no_grad():
        ref = oracle(x)
    got = ours(x.cuda()).cpu()
    assert got.shape == ref.shape
    assert rel(got, ref) <= 2e-2, rel(got, ref)
    # the fused F.normalize (face_model.py:139)
    gn = ours.embed(x.cuda(), l2_normalize=True).cpu()
    assert rel(gn, torch.nn.functional.normalize(ref)) <= 2e-2
    assert (gn.norm(dim=1) - 1).abs().max().item() <= 1e-5


def test_vit_base_patch16_224_embeddings_match_oracle(lib):
    """BASELINE config 3/5 family at full size: ViT-B/16 224^2, 197 tokens, neck Linear(151296, 512)."""
    torch.set_num_threads(min(16, torch.get_num_threads()))
    oracle = randomize_(ViTWrapperOracle("vit_base_patch16_224", 512, 224), seed=5).eval()
    ours = BackboneFactory({"timm-vit_base_patch16_224": {"pretrained": False, "image_size": 224, "feat_dim": 512}}).get_backbone()
    ours.load_state_dict(oracle.state_dict(), strict=True)
    ours = ours.cuda().eval()
    torch.manual_seed(2)
    x = torch.randn(3, 3, 224, 224)
    with torch.no_grad():
        ref = torch.nn.functional.normalize(oracle(x))
    got = ours.embed(x.cuda(), l2_normalize=True).cpu()
    cos = (got * ref).sum(dim=1)
    assert rel(got, ref) <= 3e-2, rel(got, ref)
    assert cos.min().item() >= 0.999


def test_vit_train_mode_raises_instead_of_falling_back(lib):
    _, ours = build(1, feat_dim=64, image_size=64, patch=16, dim=128, depth=1, heads=2)
    ours.train()
    with pytest.raises(NotImplementedError):
        ours(torch.randn(2, 3, 64, 64, device="cuda"))

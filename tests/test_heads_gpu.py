"""GPU parity of the margin heads + CE (csrc/heads.cu) against the golden vectors minted from the REFERENCE's own
modules (tests/golden/heads_*.npz) and against oracle/heads.py on fresh inputs.

Tolerance (floating point, stated): the reference is fp32; the CUDA path evaluates the contractions with a 3-way
bf16 split on tensor cores and uses fast exp in the softmax.  logits: |err| <= 2e-4 * scale-of-logit-range (abs 2e-3
at scale 32..256); loss: 1e-4 relative; gradients: 2e-3 of the tensor's max magnitude.
"""
import os

import numpy as np
import pytest
import torch

from oracle import heads as H
from visiondk_b200.heads import ArcFace, CircleLoss, HeadFactory, MV_Softmax, margin_ce_loss

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

KINDS = {
    "arcface": (lambda d, c: ArcFace(d, c, 0.35, 0.0, 32)),
    "arcface_am": (lambda d, c: ArcFace(d, c, 0.5, 0.2, 64)),
    "circleloss": (lambda d, c: CircleLoss(d, c, 0.25, 256)),
}


def close(got, ref, tol, what):
    err = (got - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1e-12)
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (tol {tol})"


@pytest.mark.parametrize("case", ["small", "cfg"])
@pytest.mark.parametrize("kind", list(KINDS))
def test_heads_match_reference_golden(lib, case, kind):
    z = np.load(os.path.join(GOLD, f"heads_{case}.npz"))
    feats = torch.from_numpy(z["feats"]).cuda().requires_grad_(True)
    labels = torch.from_numpy(z["labels"]).cuda()
    w = torch.from_numpy(z[f"{kind}_weight"])
    head = KINDS[kind](w.shape[0], w.shape[1]).cuda()
    with torch.no_grad():
        head.weight.copy_(w)
    smooth = float(z["label_smooth"])
    # un-fused surface: head -> logits, then torch's criterion
    logits = head(feats, labels)
    close(logits.detach().cpu(), torch.from_numpy(z[f"{kind}_logits"]), 2e-4, "logits")
    loss_unfused = torch.nn.functional.cross_entropy(logits, labels, label_smoothing=smooth)
    loss_unfused.backward()
    close(feats.grad.cpu(), torch.from_numpy(z[f"{kind}_dfeats"]), 2e-3, "dfeats (un-fused)")
    close(head.weight.grad.cpu(), torch.from_numpy(z[f"{kind}_dweight"]), 2e-3, "dweight (un-fused)")
    # fused criterion∘head
    feats.grad, head.weight.grad = None, None
    loss = margin_ce_loss(head, feats, labels, smooth)
    assert abs(loss.item() - float(z[f"{kind}_loss"])) <= 1e-4 * abs(float(z[f"{kind}_loss"])) + 1e-5
    loss.backward()
    close(feats.grad.cpu(), torch.from_numpy(z[f"{kind}_dfeats"]), 2e-3, "dfeats (fused)")
    close(head.weight.grad.cpu(), torch.from_numpy(z[f"{kind}_dweight"]), 2e-3, "dweight (fused)")


@pytest.mark.parametrize("kind,is_am", [("mv_arc", False), ("mv_am", True)])
def test_mv_softmax_matches_reference_golden(lib, kind, is_am):
    """MV-Softmax (models/faceX/head/mv_softmax.py) against vectors minted from the reference's own module
    (tests/golden/heads_mv.npz): un-fused logits + torch CE, and the fused criterion∘head."""
    z = np.load(os.path.join(GOLD, "heads_mv.npz"))
    feats = torch.from_numpy(z["feats"]).cuda().requires_grad_(True)
    labels = torch.from_numpy(z["labels"]).cuda()
    w = torch.from_numpy(z[f"{kind}_weight"])
    head = MV_Softmax(w.shape[0], w.shape[1], is_am, 0.35, 1.12, 32).cuda()
    with torch.no_grad():
        head.weight.copy_(w)
    smooth = float(z["label_smooth"])
    logits = head(feats, labels)
    close(logits.detach().cpu(), torch.from_numpy(z[f"{kind}_logits"]), 2e-4, "logits")
    torch.nn.functional.cross_entropy(logits, labels, label_smoothing=smooth).backward()
    close(feats.grad.cpu(), torch.from_numpy(z[f"{kind}_dfeats"]), 2e-3, "dfeats (un-fused)")
    close(head.weight.grad.cpu(), torch.from_numpy(z[f"{kind}_dweight"]), 2e-3, "dweight (un-fused)")
    feats.grad, head.weight.grad = None, None
    loss = margin_ce_loss(head, feats, labels, smooth)
    assert abs(loss.item() - float(z[f"{kind}_loss"])) <= 1e-4 * abs(float(z[f"{kind}_loss"])) + 1e-5
    loss.backward()
    close(feats.grad.cpu(), torch.from_numpy(z[f"{kind}_dfeats"]), 2e-3, "dfeats (fused)")
    close(head.weight.grad.cpu(), torch.from_numpy(z[f"{kind}_dweight"]), 2e-3, "dweight (fused)")


@pytest.mark.parametrize("is_am", [False, True])
def test_mv_softmax_matches_oracle_on_fresh_inputs(lib, is_am):
    """Sizes kept moderate: the hard-example test `cos > threshold` is a discontinuity, and an element within rounding of the
    threshold may legitimately land on the other side (the comparison below skips elements within 1e-5 of it)."""
    B, D, Cn = 64, 128, 300
    torch.manual_seed(5 + int(is_am))
    feats = torch.randn(B, D) * 2
    labels = torch.randint(0, Cn, (B,))
    head = MV_Softmax(D, Cn, is_am, 0.35, 1.12, 32)
    with torch.no_grad():  # some well-aligned classes so that gt > 0 / gt > margin both occur
        for r in range(0, B, 2):
            head.weight[:, labels[r]] = feats[r] / feats[r].norm() + 0.3 * head.weight[:, labels[r]]
    fo = feats.clone().requires_grad_(True)
    wo = head.weight.detach().clone().requires_grad_(True)
    ref_loss, ref_logits = H.head_loss("mv_softmax", fo, wo, labels, label_smooth=0.1, is_am=is_am, margin=0.35, mv_weight=1.12,
                                       scale=32.0)
    ref_loss.backward()
    cos = torch.mm(torch.nn.functional.normalize(feats), torch.nn.functional.normalize(wo.detach(), dim=0))
    gt = cos[torch.arange(B), labels].view(-1, 1)
    thr = gt - 0.35 if is_am else gt * np.cos(0.35) - torch.sqrt(1 - gt * gt) * np.sin(0.35)
    safe = (cos - thr).abs() > 1e-5
    head = head.cuda()
    fg = feats.clone().cuda().requires_grad_(True)
    got_logits = head(fg, labels.cuda()).detach().cpu()
    assert ((got_logits - ref_logits.detach()).abs()[safe]).max().item() <= 2e-4 * ref_logits.abs().max().item()
    if bool(safe.all()):
        loss = margin_ce_loss(head, fg, labels.cuda(), 0.1)
        loss.backward()
        assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item()) + 1e-5
        close(fg.grad.cpu(), fo.grad, 2e-3, "mv dfeats")
        close(head.weight.grad.cpu(), wo.grad, 2e-3, "mv dweight")


@pytest.mark.parametrize("B,D,Cn", [(160, 512, 1000), (7, 64, 33), (256, 512, 2051)])
def test_heads_match_oracle_on_fresh_inputs(lib, B, D, Cn):
    torch.manual_seed(B)
    feats = torch.randn(B, D) * 2
    labels = torch.randint(0, Cn, (B,))
    for kind, mk in KINDS.items():
        head = mk(D, Cn)
        fo = feats.clone().requires_grad_(True)
        wo = head.weight.detach().clone().requires_grad_(True)
        fn = "arcface" if kind.startswith("arcface") else "circleloss"
        kw = dict(margin_arc=head.margin_arc, margin_am=head.margin_am, scale=head.scale) if fn == "arcface" else \
            dict(margin=head.margin, gamma=head.gamma)
        ref_loss, ref_logits = H.head_loss(fn, fo, wo, labels, label_smooth=0.1, **kw)
        ref_loss.backward()
        head = head.cuda()
        fg = feats.clone().cuda().requires_grad_(True)
        loss = margin_ce_loss(head, fg, labels.cuda(), 0.1)
        loss.backward()
        assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item()) + 1e-5, kind
        close(fg.grad.cpu(), fo.grad, 2e-3, f"{kind} dfeats")
        close(head.weight.grad.cpu(), wo.grad, 2e-3, f"{kind} dweight")
        close(head(fg, labels.cuda()).detach().cpu(), ref_logits.detach(), 2e-4, f"{kind} logits")


def test_head_factory_surface(lib):
    h = HeadFactory({"arcface": dict(feat_dim=64, num_class=10, margin_arc=0.35, margin_am=0.0, scale=32)}).get_head()
    assert isinstance(h, ArcFace) and h.weight.shape == (64, 10)
    np.testing.assert_allclose(h.weight.detach().norm(dim=0).numpy(), 1.0, rtol=1e-5)
    h = HeadFactory({"circleloss": dict(feat_dim=64, num_class=10, margin=0.25, gamma=256)}).get_head()
    assert isinstance(h, CircleLoss)
    h2 = HeadFactory({"mv-softmax": dict(feat_dim=64, num_class=10, is_am=False, margin=0.35, mv_weight=1.12, scale=32)}).get_head()
    assert isinstance(h2, MV_Softmax)
    with pytest.raises(NotImplementedError):
        HeadFactory({"magface": {}}).get_head()
    with pytest.raises(RuntimeError):
        h(torch.zeros(2, 64), torch.zeros(2, dtype=torch.long))


def test_arcface_at_the_reference_face_config_scale(lib):
    """configs/faceX/face.yaml:28-37,42: feat_dim 128, num_class 58 671 (not a multiple of 8), per-GPU batch 160, ArcFace(0.35, 32) —
    the fused criterion∘head and its backward against the oracle's fp32 autograd at exactly that size."""
    from oracle import heads as H
    D, Cn, B = 128, 58671, 160
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, D, generator=g) * 2.0
    labels = torch.randint(0, Cn, (B,), generator=g)
    w = H.init_head_weight(D, Cn)
    with torch.no_grad():  # some class centres near their samples so that the margin branch matters
        for r in range(0, B, 4):
            w[:, labels[r]] = feats[r] / feats[r].norm() + 0.3 * w[:, labels[r]]
    f_ref = feats.clone().requires_grad_(True)
    w_ref = w.clone().requires_grad_(True)
    ref_loss = H.cross_entropy(H.arcface_logits(f_ref, w_ref, labels, 0.35, 0.0, 32.0), labels, 0.1)
    ref_loss.backward()
    head = ArcFace(D, Cn, 0.35, 0.0, 32).cuda()
    with torch.no_grad():
        head.weight.copy_(w)
    f = feats.cuda().requires_grad_(True)
    loss = margin_ce_loss(head, f, labels.cuda(), 0.1)
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item()) + 1e-5
    close(f.grad.cpu(), f_ref.grad, 2e-3, "dfeats at C=58671")
    close(head.weight.grad.cpu(), w_ref.grad, 2e-3, "dweight at C=58671")

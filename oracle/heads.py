"""Oracle for the margin-softmax heads + cross-entropy (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates, in plain PyTorch fp32 on the CPU (autograd provides the backward):
  * ArcFace.forward       /root/reference/models/faceX/head/arcface.py:20-36  (init :9-18)
  * CircleLoss.forward    /root/reference/models/faceX/head/circleloss.py:21-43 (init :9-19)
  * ce                    /root/reference/models/losses/loss.py:71-73 = nn.CrossEntropyLoss(label_smoothing)
PINNED: tests/golden/heads_*.npz were produced by oracle/make_golden.py executing the reference's own head
modules (loaded by file path) in the authoring container; tests/test_oracle_heads_cpu.py holds this
restatement to them bit for bit.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def init_head_weight(feat_dim: int, num_class: int, generator: torch.Generator | None = None) -> torch.Tensor:
    """arcface.py:11-12: uniform(-1,1) then renorm to unit columns."""
    w = torch.empty(feat_dim, num_class).uniform_(-1, 1, generator=generator)
    return w.renorm_(2, 1, 1e-5).mul_(1e5)


def cos_theta(feats: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """arcface.py:21-24 / circleloss.py:22-25: normalise columns of W and rows of f, multiply, clamp."""
    kernel_norm = F.normalize(weight, dim=0)
    feats = F.normalize(feats)
    return torch.mm(feats, kernel_norm).clamp(-1, 1)


def arcface_logits(feats, weight, labels, margin_arc=0.35, margin_am=0.0, scale=32.0):
    """arcface.py:20-36."""
    cos = cos_theta(feats, weight)
    sin = torch.sqrt(1.0 - torch.pow(cos, 2))
    cos_m = cos * math.cos(margin_arc) - sin * math.sin(margin_arc)
    cos_m = torch.where(cos > math.cos(math.pi - margin_arc), cos_m, cos - margin_am)
    onehot = F.one_hot(labels, cos.shape[1]).bool()
    out = torch.where(onehot, cos_m, cos)
    return out * scale


def circleloss_logits(feats, weight, labels, margin=0.25, gamma=256.0):
    """circleloss.py:21-43 (alpha_p / alpha_n are detached)."""
    cos = cos_theta(feats, weight)
    onehot = F.one_hot(labels, cos.shape[1]).bool()
    alpha_p = torch.clamp_min((1 + margin) - cos.detach(), 0.0)
    alpha_n = torch.clamp_min(cos.detach() + margin, 0.0)
    logit_p = alpha_p * (cos - (1 - margin))
    logit_n = alpha_n * (cos - margin)
    return torch.where(onehot, logit_p, logit_n) * gamma


def mv_softmax_logits(feats, weight, labels, is_am=False, margin=0.35, mv_weight=1.12, scale=32.0):
    """mv_softmax.py:25-44 (MV-Softmax: hard negatives — those scoring above the margin-shifted target — are re-weighted;
    the comparison carries no gradient; cos is NOT clamped here, exactly like the reference)."""
    kernel_norm = F.normalize(weight, dim=0)
    cos = torch.mm(F.normalize(feats), kernel_norm)
    gt = cos[torch.arange(labels.shape[0]), labels].view(-1, 1)
    if is_am:
        mask = cos > gt - margin
        final_gt = torch.where(gt > margin, gt - margin, gt)
    else:
        sin = torch.sqrt(1.0 - torch.pow(gt, 2))
        cos_m = gt * math.cos(margin) - sin * math.sin(margin)
        mask = cos > cos_m
        final_gt = torch.where(gt > 0.0, cos_m, gt)
    out = torch.where(mask, mv_weight * cos + mv_weight - 1.0, cos)
    onehot = F.one_hot(labels, cos.shape[1]).bool()
    out = torch.where(onehot, final_gt.expand_as(out), out)
    return out * scale


def cross_entropy(logits, labels, label_smooth: float = 0.0):
    """loss.py:71-73, called at engine/procedure/train.py:196."""
    return F.cross_entropy(logits, labels, label_smoothing=label_smooth)


def head_loss(kind, feats, weight, labels, label_smooth=0.0, **kw):
    fn = {"arcface": arcface_logits, "circleloss": circleloss_logits, "mv_softmax": mv_softmax_logits}[kind]
    logits = fn(feats, weight, labels, **kw)
    return cross_entropy(logits, labels, label_smooth), logits

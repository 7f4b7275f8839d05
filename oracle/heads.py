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


def cross_entropy(logits, labels, label_smooth: float = 0.0):
    """loss.py:71-73, called at engine/procedure/train.py:196."""
    return F.cross_entropy(logits, labels, label_smoothing=label_smooth)


def head_loss(kind, feats, weight, labels, label_smooth=0.0, **kw):
    logits = arcface_logits(feats, weight, labels, **kw) if kind == "arcface" else circleloss_logits(feats, weight, labels, **kw)
    return cross_entropy(logits, labels, label_smooth), logits

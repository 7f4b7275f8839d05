"""Oracle for the ViT backbone + Transformer neck (TEST INFRASTRUCTURE — see oracle/__init__.py).

The reference builds its backbone with `timm.create_model(name, pretrained, num_classes=0, global_pool='')`
(/root/reference/models/faceX/backbone/timm_wrapper.py:16-21); for a Transformer the un-pooled output is
`[B, tokens, C]` and the neck is LayerNorm(C) -> Flatten -> Linear(tokens*C, feat_dim) -> BatchNorm1d
(timm_wrapper.py:39-47).  timm (pinned 0.9.16) is not vendored and not installed here: PARITY UNPINNED at that
boundary.  This file restates timm 0.9.16's published VisionTransformer (timm/models/vision_transformer.py:
VisionTransformer.forward_features + forward_head with global_pool='' and num_classes=0: every token, cls first,
after the final LayerNorm(eps 1e-6); Block = x + attn(norm1(x)); x + mlp(norm2(x)); Attention = qkv Linear with
bias, softmax(q k^T / sqrt(d)) v, proj; no LayerScale, no qk-norm, learned pos_embed incl. the cls slot) in plain
PyTorch fp32 with the SAME state_dict keys:

  model.patch_embed.proj Conv2d(3, C, P, P)      model.cls_token [1,1,C]      model.pos_embed [1, 1+N, C]
  model.blocks.{i}.{norm1, attn.qkv Linear(C,3C), attn.proj Linear(C,C), norm2, mlp.fc1 Linear(C,4C), mlp.fc2}
  model.norm LayerNorm(C, eps 1e-6)
  output_layer.{0: LayerNorm(C), 1: Flatten, 2: Linear((1+N)*C, feat_dim), 3: BatchNorm1d}   (timm_wrapper.py:42-47)

Independent cross-check available in this container: HF transformers' ViTModel is architecture-identical
(tests/test_oracle_vit_cpu.py maps the weights across and compares the token outputs).

pre_norm=True restates timm 0.9.16's `vit_*_clip_*` entries (the CLIP image towers, BASELINE config 5's ViT-L/14 at 336^2:
`VisionTransformer(pre_norm=True, norm_layer=nn.LayerNorm)`): `model.norm_pre` LayerNorm after cls / position embedding,
`patch_embed.proj` WITHOUT bias, every LayerNorm at eps 1e-5, standard GELU.  Cross-check: HF transformers' CLIPVisionModel
with hidden_act="gelu" has the same tower (tests/test_oracle_vit_cpu.py maps the weights across; HF applies its
post_layernorm to the pooled token only, so the comparison is on the tokens before the final norm).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

VIT_ARCHS = {
    # timm name -> (patch, embed_dim, depth, heads)
    "vit_tiny_patch16_224": (16, 192, 12, 3),
    "vit_small_patch16_224": (16, 384, 12, 6),
    "vit_base_patch16_224": (16, 768, 12, 12),
    "vit_large_patch16_224": (16, 1024, 24, 16),
    "vit_base_patch16_clip_224": (16, 768, 12, 12),
    "vit_large_patch14_clip_224": (14, 1024, 24, 16),
    "vit_large_patch14_clip_336": (14, 1024, 24, 16),
}
VIT_PRE_NORM = {"vit_base_patch16_clip_224", "vit_large_patch14_clip_224", "vit_large_patch14_clip_336"}


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.head_dim = dim // heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(4 * dim, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, eps=1e-6):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = Mlp(dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, patch, dim, bias=True):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch, bias=bias)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)  # [B, N, C], row-major over (h, w)


class VisionTransformer(nn.Module):
    def __init__(self, image_size, patch, dim, depth, heads, pre_norm=False):
        super().__init__()
        n = (image_size // patch) ** 2
        eps = 1e-5 if pre_norm else 1e-6  # timm: norm_layer=nn.LayerNorm for the clip entries, partial(LayerNorm, eps=1e-6) otherwise
        self.patch_embed = PatchEmbed(patch, dim, bias=not pre_norm)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.randn(1, n + 1, dim) * 0.02)
        self.norm_pre = nn.LayerNorm(dim, eps=eps) if pre_norm else nn.Identity()
        self.blocks = nn.Sequential(*[Block(dim, heads, eps) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=eps)

    def forward_tokens(self, x):
        """Tokens after the last block, BEFORE the final LayerNorm (what HF's `last_hidden_state` of a CLIP tower holds)."""
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.pos_embed
        return self.blocks(self.norm_pre(x))

    def forward(self, x):
        return self.norm(self.forward_tokens(x))


class ViTWrapperOracle(nn.Module):
    """timm_wrapper.py TimmWrapper for a Transformer backbone: `model` + the `[B,N,C]` neck."""

    def __init__(self, model_name, feat_dim, image_size, patch=None, dim=None, depth=None, heads=None, pre_norm=None):
        super().__init__()
        if dim is None:
            patch, dim, depth, heads = VIT_ARCHS[model_name]
        if pre_norm is None:
            pre_norm = model_name in VIT_PRE_NORM
        assert image_size % patch == 0 and dim % heads == 0
        self.model = VisionTransformer(image_size, patch, dim, depth, heads, pre_norm=pre_norm)
        tokens = (image_size // patch) ** 2 + 1
        self.output_layer = nn.Sequential(nn.LayerNorm(dim), nn.Flatten(1), nn.Linear(tokens * dim, feat_dim),
                                          nn.BatchNorm1d(feat_dim))

    def forward(self, x):
        return self.output_layer(self.model(x))


def randomize_(m: nn.Module, seed: int = 0) -> nn.Module:
    """Random but well-scaled parameters and BatchNorm statistics (a fresh init leaves cls_token = 0, BN stats = 0/1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() >= 2 and "pos_embed" not in n and "cls_token" not in n:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan_in))
            elif n.endswith("weight"):  # norm scales
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
        for n, b in m.named_buffers():
            if n.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif n.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))
    return m

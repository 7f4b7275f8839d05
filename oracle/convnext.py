"""Oracle for the ConvNeXt backbone + neck (TEST INFRASTRUCTURE — see oracle/__init__.py).

The reference builds its backbone with `timm.create_model(name, pretrained, num_classes=0, global_pool='')`
(/root/reference/models/faceX/backbone/timm_wrapper.py:16-21) and adds the neck of timm_wrapper.py:30-38.
timm (pinned 0.9.16, /root/reference/requirements.txt:10) is a third-party dependency that is NOT vendored in
/root/reference and not installed here: PARITY UNPINNED at that boundary.  This file restates timm 0.9.16's
published ConvNeXt (timm/models/convnext.py: ConvNeXt, ConvNeXtStage, ConvNeXtBlock, NormMlpClassifierHead) in
plain PyTorch fp32 with the SAME state_dict keys, so real timm checkpoints load:

  model.stem.0 Conv2d(3,C0,4,4)  model.stem.1 LayerNorm2d(eps 1e-6)
  model.stages.{i}.downsample.{0: LayerNorm2d, 1: Conv2d(k2,s2)}          (i >= 1; stage 0: Identity)
  model.stages.{i}.blocks.{j}.{conv_dw Conv2d(7, pad 3, groups C), norm LayerNorm(eps 1e-6),
                                mlp.fc1 Linear(C,4C), GELU(erf), mlp.fc2 Linear(4C,C), gamma[C] (init 1e-6)}
  model.head.norm LayerNorm2d — applied even with global_pool='' (NormMlpClassifierHead.forward), output [B,C,7,7]
  output_layer.{0: BatchNorm2d, 1: Flatten, 2: Linear(C*h*w, feat_dim), 3: BatchNorm1d}   (timm_wrapper.py:33-38)

Independent cross-check available in this container: torchvision's convnext_base is architecture-identical
(tests/test_oracle_convnext_cpu.py maps weights across and compares the feature maps).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

CONVNEXT_ARCHS = {
    # timm name -> (depths, dims)
    "convnext_atto": ((2, 2, 6, 2), (40, 80, 160, 320)),
    "convnext_tiny": ((3, 3, 9, 3), (96, 192, 384, 768)),
    "convnext_small": ((3, 3, 27, 3), (96, 192, 384, 768)),
    "convnext_base": ((3, 3, 27, 3), (128, 256, 512, 1024)),
    "convnext_large": ((3, 3, 27, 3), (192, 384, 768, 1536)),
}


class LayerNorm2d(nn.LayerNorm):
    """timm.layers.LayerNorm2d: LayerNorm over the channel dim of NCHW."""

    def forward(self, x):
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)


class Mlp(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(4 * dim, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class ConvNeXtBlock(nn.Module):
    def __init__(self, dim, ls_init_value=1e-6):
        super().__init__()
        self.conv_dw = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim)
        self.gamma = nn.Parameter(ls_init_value * torch.ones(dim))

    def forward(self, x):
        shortcut = x
        x = self.conv_dw(x)
        x = x.permute(0, 2, 3, 1)
        x = self.norm(x)
        x = self.mlp(x)
        x = x.permute(0, 3, 1, 2)
        x = x.mul(self.gamma.reshape(1, -1, 1, 1))
        return x + shortcut


class ConvNeXtStage(nn.Module):
    def __init__(self, in_chs, out_chs, depth, downsample):
        super().__init__()
        if downsample:
            self.downsample = nn.Sequential(LayerNorm2d(in_chs, eps=1e-6), nn.Conv2d(in_chs, out_chs, kernel_size=2, stride=2))
        else:
            self.downsample = nn.Identity()
        self.blocks = nn.Sequential(*[ConvNeXtBlock(out_chs) for _ in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class NormHead(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm = LayerNorm2d(dim, eps=1e-6)

    def forward(self, x):
        return self.norm(x)


class ConvNeXt(nn.Module):
    def __init__(self, depths, dims):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(3, dims[0], kernel_size=4, stride=4), LayerNorm2d(dims[0], eps=1e-6))
        stages, prev = [], dims[0]
        for i, (d, c) in enumerate(zip(depths, dims)):
            stages.append(ConvNeXtStage(prev, c, d, downsample=i > 0))
            prev = c
        self.stages = nn.Sequential(*stages)
        self.head = NormHead(prev)

    def forward(self, x):
        return self.head(self.stages(self.stem(x)))


class TimmWrapperOracle(nn.Module):
    """timm_wrapper.py:5-54 for CNN backbones: un-pooled features -> BN2d -> Flatten -> Linear -> BN1d."""

    def __init__(self, model_name: str, feat_dim: int, image_size: int, depths=None, dims=None):
        super().__init__()
        if depths is None:
            depths, dims = CONVNEXT_ARCHS[model_name]
        self.model = ConvNeXt(depths, dims)
        hw = image_size // 32
        self.output_layer = nn.Sequential(nn.BatchNorm2d(dims[-1]), nn.Flatten(1),
                                          nn.Linear(dims[-1] * hw * hw, feat_dim), nn.BatchNorm1d(feat_dim))

    def forward(self, x):
        return self.output_layer(self.model(x))


def randomize_(module: nn.Module, seed: int = 0) -> nn.Module:
    """Random but well-conditioned parameters/buffers (default inits make gamma=1e-6 hide the MLP branch)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("gamma"):
                p.copy_(0.05 + 0.2 * torch.rand(p.shape, generator=g))
            elif p.dim() == 1 and ("norm" in name or "stem.1" in name or "downsample.0" in name
                                   or "output_layer.0" in name or "output_layer.3" in name):
                if name.endswith("weight"):
                    p.copy_(0.8 + 0.4 * torch.rand(p.shape, generator=g))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / fan_in) ** 0.5)
        for name, b in module.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif name.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))
    return module

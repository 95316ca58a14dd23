"""Pure-PyTorch restatement of the model the reference trains with `--model convnext_*` (TEST ORACLE).

What it follows:
  * Block arithmetic and LayerNorm: the only first-party statement in the reference,
    semantic_segmentation/backbone/convnext.py:32-56 (Block) and :164-182 (LayerNorm), which is
    byte-identical to object_detection/mmdet/models/backbones/convnext.py:19-54,156-180.
  * Network skeleton (4x4/s4 stem + LN, 3 x (LN + 2x2/s2 conv), linearly spaced drop-path, trunc-normal
    init): same file :73-115.
  * Everything train.py:187-194 gets from `timm.models.create_model` that is NOT in the tree (module /
    state-dict names `stem.{0,1}`, `stages.{i}.downsample.{0,1}`, `stages.{i}.blocks.{j}.{conv_dw,norm,
    mlp.fc1,mlp.fc2,gamma}`, `head.{norm,fc}`; the avg-pool -> LayerNorm2d -> fc head; gamma applied after
    the permute back; residual written `drop_path(x) + shortcut`) follows timm's published ConvNeXt
    (timm/models/convnext.py), restated from its public semantics — timm is not installed here.

Only stock torch ops are used, so this runs on CPU (fp32 oracle) and, unchanged, on a GPU under
torch.autocast (the bf16 oracle: CUDA autocast keeps layer_norm / the residual stream in fp32).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def drop_path(x, drop_prob: float = 0.0, training: bool = False, scale_by_keep: bool = True):
    """timm.layers.drop_path: per-sample Bernoulli(keep) mask, divided by keep (convnext.py:41,55)."""
    if drop_prob == 0.0 or not training:
        return x
    keep_prob = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    random_tensor = x.new_empty(shape).bernoulli_(keep_prob)
    if keep_prob > 0.0 and scale_by_keep:
        random_tensor.div_(keep_prob)
    return x * random_tensor


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training, self.scale_by_keep)


class LayerNorm(nn.LayerNorm):
    """channels_last branch of convnext.py:175-176."""

    def __init__(self, num_channels, eps=1e-6):
        super().__init__(num_channels, eps=eps)

    def forward(self, x):
        return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)


class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over C of an NCHW tensor (timm LayerNorm2d; same values as the channels_first branch :177-182)."""

    def __init__(self, num_channels, eps=1e-6):
        super().__init__(num_channels, eps=eps)

    def forward(self, x):
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class ConvNeXtBlock(nn.Module):
    """convnext.py:32-56 with timm's attribute names."""

    def __init__(self, dim, drop_path=0.0, ls_init_value=1e-6):
        super().__init__()
        self.conv_dw = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, 4 * dim)
        self.gamma = nn.Parameter(ls_init_value * torch.ones(dim)) if ls_init_value is not None else None
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        shortcut = x
        x = self.conv_dw(x)
        x = x.permute(0, 2, 3, 1)
        x = self.norm(x)
        x = self.mlp(x)
        x = x.permute(0, 3, 1, 2)
        if self.gamma is not None:
            x = x.mul(self.gamma.reshape(1, -1, 1, 1))
        return self.drop_path(x) + shortcut


class ConvNeXtStage(nn.Module):
    def __init__(self, in_chs, out_chs, stride, depth, drop_path_rates, ls_init_value):
        super().__init__()
        if in_chs != out_chs or stride > 1:
            self.downsample = nn.Sequential(LayerNorm2d(in_chs, eps=1e-6),
                                            nn.Conv2d(in_chs, out_chs, kernel_size=stride, stride=stride))
        else:
            self.downsample = nn.Identity()
        self.blocks = nn.Sequential(*[ConvNeXtBlock(out_chs, drop_path=drop_path_rates[j], ls_init_value=ls_init_value)
                                      for j in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class Head(nn.Module):
    def __init__(self, in_features, num_classes):
        super().__init__()
        self.norm = LayerNorm2d(in_features, eps=1e-6)
        self.fc = nn.Linear(in_features, num_classes)

    def forward(self, x):
        x = x.mean((2, 3), keepdim=True)
        x = self.norm(x)
        return self.fc(x.flatten(1))


class ConvNeXt(nn.Module):
    def __init__(self, num_classes=1000, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), drop_path_rate=0.0,
                 ls_init_value=1e-6, in_chans=3):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4), LayerNorm2d(dims[0], eps=1e-6))
        rates = [r.tolist() for r in torch.linspace(0, drop_path_rate, sum(depths)).split(list(depths))]
        stages, prev = [], dims[0]
        for i in range(4):
            stages.append(ConvNeXtStage(prev, dims[i], 2 if i > 0 else 1, depths[i], rates[i], ls_init_value))
            prev = dims[i]
        self.stages = nn.Sequential(*stages)
        self.head = Head(prev, num_classes)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        return self.head(self.stages(self.stem(x)))


ARCHS = {
    "convnext_tiny": dict(depths=(3, 3, 9, 3), dims=(96, 192, 384, 768)),
    "convnext_small": dict(depths=(3, 3, 27, 3), dims=(96, 192, 384, 768)),
    "convnext_base": dict(depths=(3, 3, 27, 3), dims=(128, 256, 512, 1024)),
    "convnext_large": dict(depths=(3, 3, 27, 3), dims=(192, 384, 768, 1536)),
}


def create_model(model_name, pretrained=False, num_classes=1000, drop_path_rate=0.0, **kw):
    """Stand-in for timm.models.create_model as called at train.py:187-194."""
    assert not pretrained, "no network"
    return ConvNeXt(num_classes=num_classes, drop_path_rate=drop_path_rate, **ARCHS[model_name], **kw)

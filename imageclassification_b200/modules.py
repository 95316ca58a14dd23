"""nn.Module mirror of the reference's model-side interface.

train.py:187-196 builds the model with timm's `create_model('convnext_*', pretrained, num_classes,
drop_path_rate)`; these classes keep timm's attribute names, constructor keywords and state-dict
keys/shapes (`stages.{i}.blocks.{j}.{conv_dw,norm,mlp.fc1,mlp.fc2,gamma}`, `stem.{0,1}`,
`stages.{i}.downsample.{0,1}`, `head.{norm,fc}`) so checkpoints and the reference's whole-module
pickling (utils.py:542) keep working, while the Block's arithmetic runs in the libcnx kernels
(the in-tree statement of the Block is semantic_segmentation/backbone/convnext.py:21-56).
Parameters stay canonical fp32 tensors; kernel-side layouts (bf16 copies, transposes) are derived per step.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops  # noqa: F401  (registers the torch.ops.cnx.* custom ops the modules below call)


def _trunc_normal_(t: torch.Tensor, std: float = 0.02) -> torch.Tensor:
    return nn.init.trunc_normal_(t, std=std)


class DropPath(nn.Module):
    """timm DropPath (stochastic depth per sample, scale_by_keep=True); convnext.py:41,55."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.scale_by_keep = scale_by_keep

    def sample_scale(self, n: int, device) -> torch.Tensor | None:
        """Per-sample scale [n] drawn with the same torch RNG call as timm.drop_path (keeps RNG parity)."""
        if self.drop_prob == 0.0 or not self.training:
            return None
        keep = 1.0 - self.drop_prob
        r = torch.empty((n, 1, 1, 1), dtype=torch.float32, device=device).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            r.div_(keep)
        return r.view(n)

    def forward(self, x):
        s = self.sample_scale(x.shape[0], x.device)
        return x if s is None else x * s.view(-1, *([1] * (x.dim() - 1))).to(x.dtype)

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


class LayerNorm(nn.LayerNorm):
    """channels-last LayerNorm (timm.layers.LayerNorm; convnext.py:158-176), eps 1e-6."""

    def __init__(self, num_channels: int, eps: float = 1e-6, affine: bool = True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)

    def forward(self, x):
        return torch.ops.cnx.layer_norm_cl(x, self.weight, self.bias, self.eps)


class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over C of an NCHW tensor (timm.layers.LayerNorm2d == convnext.py:177-182 channels_first)."""

    def __init__(self, num_channels: int, eps: float = 1e-6, affine: bool = True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)

    def forward(self, x):
        x = x.permute(0, 2, 3, 1)
        x = torch.ops.cnx.layer_norm_cl(x, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)


class Mlp(nn.Module):
    """timm.layers.Mlp container (fc1 -> GELU -> fc2); only holds the parameters, the Block runs the math."""

    def __init__(self, in_features: int, hidden_features: int, out_features: int | None = None):
        super().__init__()
        out_features = out_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.drop1 = nn.Identity()
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Identity()


class ConvNeXtBlock(nn.Module):
    """timm ConvNeXtBlock(in_chs, kernel_size=7, mlp_ratio=4, ls_init_value=1e-6, drop_path=0.) backed by libcnx."""

    def __init__(self, in_chs: int, out_chs: int | None = None, kernel_size: int = 7, stride: int = 1,
                 mlp_ratio: float = 4, ls_init_value: float | None = 1e-6, drop_path: float = 0.0):
        super().__init__()
        out_chs = out_chs or in_chs
        if kernel_size != 7 or stride != 1 or out_chs != in_chs:
            raise NotImplementedError("libcnx implements the ConvNeXt 7x7 stride-1 depthwise Block")
        self.conv_dw = nn.Conv2d(in_chs, in_chs, kernel_size=7, padding=3, groups=in_chs)
        self.norm = LayerNorm(in_chs, eps=1e-6)
        self.mlp = Mlp(in_chs, int(mlp_ratio * in_chs))
        self.gamma = nn.Parameter(ls_init_value * torch.ones(in_chs)) if ls_init_value is not None else None
        self.shortcut = nn.Identity()
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        dp = self.drop_path.sample_scale(x.shape[0], x.device) if isinstance(self.drop_path, DropPath) else None
        return torch.ops.cnx.block_forward(x, self.conv_dw.weight, self.conv_dw.bias, self.norm.weight, self.norm.bias,
                                           self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias,
                                           self.gamma, dp, self.norm.eps)


class ConvNeXtStage(nn.Module):
    def __init__(self, in_chs: int, out_chs: int, stride: int, depth: int, drop_path_rates, ls_init_value):
        super().__init__()
        if in_chs != out_chs or stride > 1:
            self.downsample = nn.Sequential(LayerNorm2d(in_chs, eps=1e-6),
                                            nn.Conv2d(in_chs, out_chs, kernel_size=stride, stride=stride))
        else:
            self.downsample = nn.Identity()
        self.blocks = nn.Sequential(*[
            ConvNeXtBlock(out_chs, ls_init_value=ls_init_value, drop_path=drop_path_rates[j]) for j in range(depth)])

    def forward(self, x):
        ds = self.downsample
        if isinstance(ds, nn.Sequential):
            conv = ds[1]
            if conv.kernel_size == (2, 2) and conv.stride == (2, 2) and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
                # the first Block widens a bf16 conv output to fp32 on entry (fp32 layer scale): let the conv write that tensor
                b0 = self.blocks[0] if len(self.blocks) else None
                widen = (isinstance(b0, ConvNeXtBlock) and b0.gamma is not None and b0.gamma.dtype == torch.float32
                         and not ops.KEEP_BF16_STREAM)
                x = torch.ops.cnx.downsample_forward(x, ds[0].weight, ds[0].bias, conv.weight, conv.bias, ds[0].eps, widen)
            else:
                x = ds(x)
        return self.blocks(x)


class NormMlpClassifierHead(nn.Module):
    """timm head: global avg pool -> LayerNorm2d -> flatten -> (drop) -> fc."""

    def __init__(self, in_features: int, num_classes: int, drop_rate: float = 0.0):
        super().__init__()
        self.in_features = in_features
        self.num_features = in_features
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.norm = LayerNorm2d(in_features, eps=1e-6)
        self.flatten = nn.Flatten(1)
        self.pre_logits = nn.Identity()
        self.drop = nn.Dropout(drop_rate)
        self.fc = nn.Linear(in_features, num_classes) if num_classes > 0 else nn.Identity()

    def forward(self, x, pre_logits: bool = False):
        if (not pre_logits and isinstance(self.fc, nn.Linear) and self.fc.bias is not None and ops.head_supported(x, self.fc.weight)
                and not (self.training and self.drop.p > 0)):
            # pool -> LayerNorm -> fc on the libcnx kernels; other shapes (e.g. 2 classes) use the ATen modules
            return torch.ops.cnx.head_forward(x, self.norm.weight, self.norm.bias, self.fc.weight, self.fc.bias, self.norm.eps)
        x = self.global_pool(x)
        x = self.norm(x)
        x = self.flatten(x)
        x = self.pre_logits(x)
        x = self.drop(x)
        return x if pre_logits else self.fc(x)


class ConvNeXt(nn.Module):
    """timm ConvNeXt (patch stem, 4 stages, NormMlpClassifierHead) with libcnx Blocks."""

    def __init__(self, in_chans: int = 3, num_classes: int = 1000, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768),
                 ls_init_value: float | None = 1e-6, head_init_scale: float = 1.0, drop_rate: float = 0.0,
                 drop_path_rate: float = 0.0, patch_size: int = 4):
        super().__init__()
        self.num_classes = num_classes
        self.drop_rate = drop_rate
        self.stem = nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=patch_size, stride=patch_size),
                                  LayerNorm2d(dims[0], eps=1e-6))
        dp_rates = [r.tolist() for r in torch.linspace(0, drop_path_rate, sum(depths)).split(list(depths))]
        stages = []
        prev = dims[0]
        for i in range(4):
            stages.append(ConvNeXtStage(prev, dims[i], stride=2 if i > 0 else 1, depth=depths[i],
                                        drop_path_rates=dp_rates[i], ls_init_value=ls_init_value))
            prev = dims[i]
        self.stages = nn.Sequential(*stages)
        self.num_features = self.head_hidden_size = prev
        self.norm_pre = nn.Identity()
        self.head = NormMlpClassifierHead(prev, num_classes, drop_rate=drop_rate)
        self.apply(self._init_weights)
        if isinstance(self.head.fc, nn.Linear):
            self.head.fc.weight.data.mul_(head_init_scale)
            self.head.fc.bias.data.mul_(head_init_scale)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Conv2d):
            _trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=0.02)
            nn.init.zeros_(m.bias)

    def forward_features(self, x):
        # channels-last from the first kernel on: the patch-conv output is then already the [N,H,W,C]
        # row-major matrix the Block kernels consume (no permute copies anywhere in the network)
        conv, norm = self.stem[0], self.stem[1]
        if conv.kernel_size == (4, 4) and conv.stride == (4, 4) and x.shape[2] % 4 == 0 and x.shape[3] % 4 == 0:
            x = torch.ops.cnx.stem_forward(x, conv.weight, conv.bias, norm.weight, norm.bias, norm.eps)
        else:
            x = self.stem(x.contiguous(memory_format=torch.channels_last))
        x = self.stages(x)
        return self.norm_pre(x)

    def forward_head(self, x, pre_logits: bool = False):
        return self.head(x, pre_logits=pre_logits)

    def forward(self, x):
        return self.forward_head(self.forward_features(x))

    def get_classifier(self):
        return self.head.fc

    def reset_classifier(self, num_classes: int):
        self.num_classes = num_classes
        self.head.fc = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        if isinstance(self.head.fc, nn.Linear):
            self._init_weights(self.head.fc)
            self.head.fc.to(self.stem[0].weight.device)


_ARCHS = {
    "convnext_atto": dict(depths=(2, 2, 6, 2), dims=(40, 80, 160, 320)),
    "convnext_femto": dict(depths=(2, 2, 6, 2), dims=(48, 96, 192, 384)),
    "convnext_pico": dict(depths=(2, 2, 6, 2), dims=(64, 128, 256, 512)),
    "convnext_nano": dict(depths=(2, 2, 8, 2), dims=(80, 160, 320, 640)),
    "convnext_tiny": dict(depths=(3, 3, 9, 3), dims=(96, 192, 384, 768)),
    "convnext_small": dict(depths=(3, 3, 27, 3), dims=(96, 192, 384, 768)),
    "convnext_base": dict(depths=(3, 3, 27, 3), dims=(128, 256, 512, 1024)),
    "convnext_large": dict(depths=(3, 3, 27, 3), dims=(192, 384, 768, 1536)),
    "convnext_xlarge": dict(depths=(3, 3, 27, 3), dims=(256, 512, 1024, 2048)),
}


def _supported(name: str) -> bool:
    return all(d % 32 == 0 for d in _ARCHS[name]["dims"])          # libcnx dwconv kernels work on 32-channel chunks


def list_models():
    """the model names create_model accepts (atto / femto / nano have channel counts that are not multiples of 32)"""
    return sorted(n for n in _ARCHS if _supported(n))


def create_model(model_name: str, pretrained: bool = False, num_classes: int = 1000, drop_path_rate: float = 0.0,
                 **kwargs) -> ConvNeXt:
    """Drop-in for `timm.models.create_model` as called at train.py:187-194 (convnext_* names only)."""
    if model_name not in _ARCHS:
        raise ValueError(f"unknown model {model_name!r}; imageclassification_b200 provides {list_models()}")
    if not _supported(model_name):
        raise NotImplementedError(f"{model_name}: libcnx dwconv kernels need channel counts that are multiples of 32")
    if pretrained:
        raise RuntimeError("pretrained weights need network access; load a checkpoint with load_state_dict instead")
    return ConvNeXt(num_classes=num_classes, drop_path_rate=drop_path_rate, **_ARCHS[model_name], **kwargs)

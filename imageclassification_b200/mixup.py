"""Mixup — drop-in for timm.data.mixup.Mixup as built at train.py:176-185 and called at engine.py:44:
`samples, targets = mixup_fn(samples, targets)`; mutates `samples` in place, returns soft targets [B,K] fp32.

Host side keeps timm's numpy RNG call sequence exactly (np.random.rand / beta / randint), so with the same
`np.random.seed` the same lambda and cut-mix box are drawn (train.py:116-118 seeds numpy per rank).  The label
mixing is one libcnx kernel, bit-exact in fp32.  Image mixing is one in-place libcnx pass over the
sample pairs with timm's rounding order (`cnx_mixup_batch`: mixup blend or cutmix box swap, optionally writing the un-mixed
copy the engine keeps); other dtypes / layouts use timm's own tensor expression."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def rand_bbox(img_shape, lam, margin=0.0, count=None):
    ratio = np.sqrt(1 - lam)
    img_h, img_w = img_shape[-2:]
    cut_h, cut_w = int(img_h * ratio), int(img_w * ratio)
    margin_y, margin_x = int(margin * cut_h), int(margin * cut_w)
    cy = np.random.randint(0 + margin_y, img_h - margin_y, size=count)
    cx = np.random.randint(0 + margin_x, img_w - margin_x, size=count)
    yl = np.clip(cy - cut_h // 2, 0, img_h)
    yh = np.clip(cy + cut_h // 2, 0, img_h)
    xl = np.clip(cx - cut_w // 2, 0, img_w)
    xh = np.clip(cx + cut_w // 2, 0, img_w)
    return yl, yh, xl, xh


def rand_bbox_minmax(img_shape, minmax, count=None):
    if len(minmax) != 2:
        raise ValueError("cutmix_minmax needs (min, max)")
    img_h, img_w = img_shape[-2:]
    cut_h = np.random.randint(int(img_h * minmax[0]), int(img_h * minmax[1]), size=count)
    cut_w = np.random.randint(int(img_w * minmax[0]), int(img_w * minmax[1]), size=count)
    yl = np.random.randint(0, img_h - cut_h, size=count)
    xl = np.random.randint(0, img_w - cut_w, size=count)
    return yl, yl + cut_h, xl, xl + cut_w


def cutmix_bbox_and_lam(img_shape, lam, ratio_minmax=None, correct_lam=True, count=None):
    if ratio_minmax is not None:
        yl, yu, xl, xu = rand_bbox_minmax(img_shape, ratio_minmax, count=count)
    else:
        yl, yu, xl, xu = rand_bbox(img_shape, lam, count=count)
    if correct_lam or ratio_minmax is not None:
        bbox_area = (yu - yl) * (xu - xl)
        lam = 1.0 - bbox_area / float(img_shape[-2] * img_shape[-1])
    return (yl, yu, xl, xu), lam


def mixup_target(target, num_classes, lam=1.0, smoothing=0.0):
    return torch.ops.cnx.mixup_target(target, num_classes, float(lam), float(smoothing))


class Mixup:
    def __init__(self, mixup_alpha=1.0, cutmix_alpha=0.0, cutmix_minmax=None, prob=1.0, switch_prob=0.5, mode="batch",
                 correct_lam=True, label_smoothing=0.1, num_classes=1000):
        self.mixup_alpha = mixup_alpha
        self.cutmix_alpha = cutmix_alpha
        self.cutmix_minmax = cutmix_minmax
        if self.cutmix_minmax is not None:
            if len(self.cutmix_minmax) != 2:
                raise ValueError("cutmix_minmax needs (min, max)")
            self.cutmix_alpha = 1.0
        self.mix_prob = prob
        self.switch_prob = switch_prob
        self.label_smoothing = label_smoothing
        self.num_classes = num_classes
        if mode != "batch":
            raise NotImplementedError("imageclassification_b200.Mixup implements the reference default "
                                      "--mixup_mode batch (train.py:77); 'pair'/'elem' are not on the hot path")
        self.mode = mode
        self.correct_lam = correct_lam
        self.mixup_enabled = True

    def _params_per_batch(self):
        lam, use_cutmix = 1.0, False
        if self.mixup_enabled and np.random.rand() < self.mix_prob:
            if self.mixup_alpha > 0.0 and self.cutmix_alpha > 0.0:
                use_cutmix = np.random.rand() < self.switch_prob
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha) if use_cutmix else \
                    np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.mixup_alpha > 0.0:
                lam_mix = np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.cutmix_alpha > 0.0:
                use_cutmix = True
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha)
            else:
                raise ValueError("One of mixup_alpha > 0., cutmix_alpha > 0., cutmix_minmax not None should be true.")
            lam = float(lam_mix)
        return lam, use_cutmix

    def _mix_batch(self, x, original_out=None):
        lam, use_cutmix = self._params_per_batch()
        fused = x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
        if lam == 1.0:
            if original_out is not None:
                original_out.copy_(x)
            return 1.0
        if use_cutmix:
            (yl, yh, xl, xh), lam = cutmix_bbox_and_lam(x.shape, lam, ratio_minmax=self.cutmix_minmax,
                                                        correct_lam=self.correct_lam)
            if fused:
                torch.ops.cnx.mixup_batch(x, float(lam), [int(yl), int(yh), int(xl), int(xh)], original_out)
                return lam
            if original_out is not None:
                original_out.copy_(x)
            x[:, :, yl:yh, xl:xh] = x.flip(0)[:, :, yl:yh, xl:xh]
        else:
            if fused:
                torch.ops.cnx.mixup_batch(x, float(lam), None, original_out)    # one pass instead of flip / mul_ / mul_ / add_
                return lam
            if original_out is not None:
                original_out.copy_(x)
            x_flipped = x.flip(0).mul_(1.0 - lam)                       # other dtypes / layouts: timm's own tensor ops
            x.mul_(lam).add_(x_flipped)
        return lam

    def __call__(self, x, target, original_out=None):
        """`original_out` (extension, used by this package's engine): a tensor like `x` that receives the un-mixed batch in
        the same pass — the copy engine.py:40 keeps for the accuracy forward."""
        if len(x) % 2 != 0:
            raise ValueError("Batch size should be even when using this")
        lam = self._mix_batch(x, original_out)
        target = mixup_target(target, self.num_classes, lam, self.label_smoothing)
        return x, target

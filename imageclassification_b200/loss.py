"""SoftTargetCrossEntropy — drop-in for timm.loss.SoftTargetCrossEntropy as built at train.py:257 and called at
engine.py:49,52: `criterion(output[B,K], targets[B,K]) -> 0-dim tensor` supporting .item(), /= and .backward().
Forward and backward are one libcnx kernel each; fp32 math regardless of the logits dtype,
which is what CUDA autocast does for log_softmax."""
import torch
import torch.nn as nn

from . import ops  # noqa: F401  (registers torch.ops.cnx.*)


class SoftTargetCrossEntropy(nn.Module):
    def forward(self, x, target):
        return torch.ops.cnx.soft_target_cross_entropy(x, target)

"""timm.loss.SoftTargetCrossEntropy restated (TEST ORACLE).  Constructed at train.py:257, called at engine.py:49,52.
timm is not in the reference tree; this is its published definition (timm/loss/cross_entropy.py)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class SoftTargetCrossEntropy(nn.Module):
    def forward(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        loss = torch.sum(-target * F.log_softmax(x, dim=-1), dim=-1)
        return loss.mean()

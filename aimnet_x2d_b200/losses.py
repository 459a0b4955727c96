"""Weighted L1 / MSE losses with the reference's constructor and forward (``src/models/losses.py:14-87``),
computed by one fused kernel (value + gradient)."""
import torch
import torch.nn as nn

from . import ops


class _WeightedLoss(nn.Module):
    kind = 0

    def __init__(self, weights: torch.Tensor):
        super().__init__()
        self.register_buffer("weights", weights)

    def forward(self, y_pred: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
        w = self.weights.to(device=y_pred.device, dtype=torch.float32)
        return ops.WeightedLossFn.apply(y_pred, y_true.to(torch.float32), w, self.kind)


class WeightedL1Loss(_WeightedLoss):
    """mean_b sum_t w_t |pred - true|  (losses.py:29-49)."""
    kind = 0


class WeightedMSELoss(_WeightedLoss):
    """mean_b sum_t w_t (pred - true)^2  (losses.py:67-87)."""
    kind = 1

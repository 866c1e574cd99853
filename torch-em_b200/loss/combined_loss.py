"""``CombinedLoss`` with torch-em's signature (torch_em/loss/combined_loss.py:6-38): a weighted sum of loss modules; each
member runs its own fused kernels."""
from typing import List

import torch


class CombinedLoss(torch.nn.Module):
    def __init__(self, *losses: torch.nn.Module, loss_weights: List[float] = None):
        super().__init__()
        self.losses = torch.nn.ModuleList(losses)
        n_losses = len(self.losses)
        if loss_weights is None:
            self.loss_weights = [1.0 / n_losses] * n_losses if n_losses > 0 else None
        else:
            assert len(loss_weights) == n_losses
            self.loss_weights = loss_weights

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        assert self.loss_weights is not None
        return sum([loss(x, y) * weight for loss, weight in zip(self.losses, self.loss_weights)])

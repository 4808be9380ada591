"""fp32 restatement of the reference reconstruction / adversarial losses.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def rmse(yhat: torch.Tensor, y: torch.Tensor, eps: float = 1e-16) -> torch.Tensor:
    """RMSELoss.forward — loss.py:17-19: sqrt(mean((yhat-y)^2) + eps)."""
    return torch.sqrt(F.mse_loss(yhat, y) + eps)


def local_loss(yhat, y, mask, base: str = "l1") -> torch.Tensor:
    """LocalLoss.forward — loss.py:35-47 with a working base (nn.L1Loss / nn.MSELoss).

    sum(base(y*m, yhat*m)) / count(m != 0); the count is exact.  The RMSE base
    is unconstructible in the reference (loss.py:28 passes reduction= to
    RMSELoss.__init__), so its sqrt branch is dead code and is not restated.
    """
    a, b = y * mask, yhat * mask
    if base == "l1":
        tot = (a - b).abs().sum()
    elif base == "mse":
        tot = ((a - b) ** 2).sum()
    else:
        raise ValueError(base)
    return tot / (mask != 0).float().sum()


def masked_pixel_count(mask: torch.Tensor) -> int:
    return int((mask != 0).sum().item())


def bce(pred: torch.Tensor, target_value: float) -> torch.Tensor:
    """nn.BCELoss on sigmoid outputs against a constant target (minimaxgan_rmse.py:61,135,141,162).

    torch clamps log() at -100.
    """
    return F.binary_cross_entropy(pred, torch.full_like(pred, target_value))


def lsgan(pred: torch.Tensor, target_value: float) -> torch.Tensor:
    """nn.MSELoss against ones / zeros (experiment1_global_local_D.py:119,162-163,183-194)."""
    return F.mse_loss(pred, torch.full_like(pred, target_value))


def tv_loss(img: torch.Tensor, tv_weight: float) -> torch.Tensor:
    """tv_loss — loss.py:138-151: tv_weight * (mean of squared horizontal differences + mean of squared vertical differences)."""
    w_variance = torch.mean((img[:, :, :, :-1] - img[:, :, :, 1:]) ** 2)
    h_variance = torch.mean((img[:, :, :-1, :] - img[:, :, 1:, :]) ** 2)
    return tv_weight * (h_variance + w_variance)

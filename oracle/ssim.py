"""fp32 restatement of the reference's SSIM metric (lib/pytorch_ssim/__init__.py:10-40).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against the reference's own module by
``oracle/pin_ssim_against_reference.py`` (fixture tests/golden/reference_ssim.pt).
"""
from __future__ import annotations

from math import exp

import torch
import torch.nn.functional as F


def window(window_size: int = 11, sigma: float = 1.5, channel: int = 1) -> torch.Tensor:
    """gaussian + create_window (:10-18): normalised 1-D Gaussian, outer product, one copy per channel."""
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """_ssim (:20-40): depthwise Gaussian means / variances / covariance with zero padding, C1 = 0.01^2, C2 = 0.03^2."""
    c = img1.shape[1]
    w = window(window_size, 1.5, c).to(img1)
    pad = window_size // 2
    conv = lambda t: F.conv2d(t, w, padding=pad, groups=c)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1, s2, s12 = conv(img1 * img1) - mu1_sq, conv(img2 * img2) - mu2_sq, conv(img1 * img2) - mu12
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)

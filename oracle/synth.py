"""Deterministic synthetic inputs.  TEST INFRASTRUCTURE.

The reference's loader yields ``(ground[B,1,R,R] in [0,1], mask[B,1,R,R] in
[0,1], segment)`` — lib/data/dataset.py:14-51, train.py:69-79.
"""
from __future__ import annotations

import torch


def make_batch(batch: int, res: int, seed: int = 1234, rank: int = 0, soft_edge: bool = False):
    """ground ~ U[0,1); mask = one axis-aligned rectangle per image, side in
    [R/8, R/2].  ``soft_edge`` adds a 1-px border of values in (0,1) that
    ``ceil`` must turn into 1 (wgan_l1.py:114)."""
    g = torch.Generator().manual_seed(seed + rank)
    ground = torch.rand(batch, 1, res, res, generator=g, dtype=torch.float32)
    mask = torch.zeros(batch, 1, res, res, dtype=torch.float32)
    lo, hi = res // 8, res // 2
    for b in range(batch):
        h = int(torch.randint(lo, hi + 1, (1,), generator=g))
        w = int(torch.randint(lo, hi + 1, (1,), generator=g))
        t = int(torch.randint(0, res - h + 1, (1,), generator=g))
        l = int(torch.randint(0, res - w + 1, (1,), generator=g))
        if soft_edge:
            t0, l0 = max(t - 1, 0), max(l - 1, 0)
            t1, l1 = min(t + h + 1, res), min(l + w + 1, res)
            edge = torch.rand(t1 - t0, l1 - l0, generator=g) * 0.98 + 0.01
            mask[b, 0, t0:t1, l0:l1] = edge
        mask[b, 0, t : t + h, l : l + w] = 1.0
    return ground, mask

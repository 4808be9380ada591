"""Drop-in for the reference's ``lib/pytorch_ssim`` (lib/pytorch_ssim/__init__.py:10-76): ``ssim(img1, img2, window_size=11,
size_average=True)``, ``SSIM(window_size, size_average)``, ``gaussian`` and ``create_window``.  The reference logs it per batch on
detached images (experiment1_global_local_D.py:209), so this is forward-only: one fused sm_100a kernel (separable 11-tap Gaussian
out of shared memory + the pointwise SSIM formula + block reduction) instead of 5 depthwise conv2d and ~15 elementwise launches.
"""
from __future__ import annotations

import ctypes
from math import exp

import torch

from ... import _lib
from ..._lib import call, ptr, stream


def gaussian(window_size, sigma):
    """pytorch_ssim gaussian (:10-12): normalised 1-D Gaussian, fp32."""
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def create_window(window_size, channel):
    """pytorch_ssim create_window (:14-18): [channel, 1, ws, ws] outer product of the 1-D Gaussian (kept for callers that build
    their own window; the kernel applies the two 1-D passes)."""
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def _ssim_cuda(img1, img2, window_size, size_average):
    if window_size != 11:
        raise NotImplementedError("the fused SSIM kernel covers the reference's window_size = 11")
    if not (img1.is_cuda and img2.is_cuda):
        raise _lib.B200Error("ssim runs on CUDA (sm_100a) only; there is no CPU fallback")
    if img1.shape != img2.shape or img1.dim() != 4:
        raise _lib.B200Error(f"ssim expects two NCHW tensors of the same shape, got {tuple(img1.shape)} and {tuple(img2.shape)}")
    a, b = img1.detach().contiguous().float(), img2.detach().contiguous().float()
    N, C, H, W = a.shape
    win = (ctypes.c_float * 11)(*gaussian(11, 1.5).tolist())
    partial = torch.empty(_lib.value("b200_ssim_partials", N, C, H, W), dtype=torch.float32, device=a.device)
    out = torch.empty(1 if size_average else N, dtype=torch.float32, device=a.device)
    call("b200_ssim", ptr(a), ptr(b), N, C, H, W, win, int(bool(size_average)), ptr(partial), ptr(out), stream())
    return out[0] if size_average else out


def ssim(img1, img2, window_size=11, size_average=True):
    """:68-76.  Returns a 0-dim tensor (``.item()`` at the call site) or, with ``size_average=False``, one value per image."""
    return _ssim_cuda(img1, img2, window_size, size_average)


class SSIM(torch.nn.Module):
    """:42-66 (the cached window of the reference is an implementation detail: the taps are kernel constants here)."""

    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        self.window_size, self.size_average, self.channel = window_size, size_average, 1

    def forward(self, img1, img2):
        return _ssim_cuda(img1, img2, self.window_size, self.size_average)

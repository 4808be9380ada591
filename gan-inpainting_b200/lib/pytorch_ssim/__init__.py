"""Drop-in for the reference's ``lib/pytorch_ssim`` (lib/pytorch_ssim/__init__.py:10-76): ``ssim(img1, img2, window_size=11,
size_average=True)``, ``SSIM(window_size, size_average)``, ``gaussian`` and ``create_window``.  The reference logs it per batch on
detached images (experiment1_global_local_D.py:209), so this is forward-only: one fused sm_100a kernel (separable 11-tap Gaussian
out of shared memory + the pointwise SSIM formula + block reduction) instead of 5 depthwise conv2d and ~15 elementwise launches.
"""
from __future__ import annotations

import ctypes

import torch

from ... import _lib
from ..._lib import call, ptr, stream


def gaussian(window_size, sigma):
    """The 1-D window of the metric: ``window_size`` samples of a zero-centred Gaussian of width ``sigma``, normalised to sum 1
    (fp32, like the reference's helper of the same name)."""
    offsets = torch.arange(window_size, dtype=torch.float64) - (window_size // 2)
    taps = torch.exp(-(offsets * offsets) / (2.0 * float(sigma) ** 2))
    return (taps / taps.sum()).float()


def create_window(window_size, channel):
    """``[channel, 1, window_size, window_size]``: one 2-D window (outer product of the 1-D taps, sigma 1.5) per channel, for
    callers that convolve themselves; the fused kernel applies the two 1-D passes."""
    taps = gaussian(window_size, 1.5)
    return torch.outer(taps, taps)[None, None].repeat(channel, 1, 1, 1).contiguous()


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

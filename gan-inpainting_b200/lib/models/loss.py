"""Drop-in ``lib.models.loss`` (reference lib/models/loss.py:11-47): ``RMSELoss`` and ``LocalLoss`` as fused CUDA
reductions with hand-written backward kernels (csrc/losses.cu).

``tv_loss`` (loss.py:138-151, part of g_loss in config 4) is a fused stencil + reduction with its own backward kernel
(csrc/tv.cu).  ``perceptual_loss`` / ``style_loss`` / ``perceptual_and_style_loss`` / ``gram_matrix`` (loss.py:49-136) run the VGG19
feature extractor on the tcgen05 3x3 window kernel (``engine.VggFeatureEngine``); like the reference they contribute no gradient
(``torch.no_grad()`` on detached tensors).  The reference builds its VGG at import time with downloaded weights
(loss.py:4-5); here it is built lazily with random initialisation unless ``set_vgg`` installs one.
"""
from __future__ import annotations

import torch
from torch import nn

from ... import _lib
from ...engine import Scratch
from ..._lib import call, ptr, stream

RECON_L1, RECON_RMSE, RECON_LOCAL_L1, RECON_LOCAL_MSE, RECON_MASKED_RMSE = 0, 1, 2, 3, 4


def _check(*ts):
    n = ts[0].numel()
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.B200Error("loss kernels run on CUDA tensors only; there is no CPU fallback")
        if t.numel() != n:
            raise _lib.B200Error(f"shape mismatch in loss inputs: {[tuple(x.shape) for x in ts if x is not None]}")
    if n % 4:
        raise _lib.B200Error("loss inputs must hold a multiple of 4 elements")


class _ReconFunction(torch.autograd.Function):
    """loss(first, second[, mask]) for the symmetric kinds of csrc/losses.cu; d/d(first) and d/d(second) by one
    elementwise kernel each (roles swapped)."""

    @staticmethod
    def forward(ctx, kind, first, second, mask, eps):
        a = first.detach().contiguous().float()
        b = second.detach().contiguous().float()
        m = None if mask is None else mask.detach().contiguous().float()
        _check(a, b, m)
        out = torch.empty(4, dtype=torch.float32, device=a.device)
        call("b200_recon_loss", kind, ptr(a), ptr(b), ptr(m), a.numel(), float(eps), ptr(Scratch.get(a.device).recon), ptr(out), stream())
        ctx.kind, ctx.a, ctx.b, ctx.m, ctx.out = kind, a, b, m, out
        ctx.shapes = (first.shape, second.shape)
        return out[0].clone()

    @staticmethod
    def backward(ctx, gout):
        gout = gout.contiguous().float()
        res = [None, None, None, None, None]
        for slot, (y, yh) in ((1, (ctx.b, ctx.a)), (2, (ctx.a, ctx.b))):
            if not ctx.needs_input_grad[slot]:
                continue
            d = torch.empty_like(yh)
            call("b200_gstep_seed", ptr(y), ptr(yh), None, ptr(ctx.m), None, None, ctx.kind, 1.0, ptr(ctx.out), -1, 0.0, None, ptr(gout), 0, yh.numel(),
                 ptr(d), None, None, None, 0, stream())
            res[slot] = d.view(ctx.shapes[slot - 1])
        return tuple(res)


class RMSELoss(nn.Module):
    """loss.py:11-19: ``sqrt(mean((yhat - y)**2) + eps)``; eps avoids the infinite derivative of sqrt at 0."""

    def __init__(self, eps=1e-16):
        super().__init__()
        self.eps = eps

    def forward(self, yhat, y):
        return _ReconFunction.apply(RECON_RMSE, yhat, y, None, self.eps)


class LocalLoss(nn.Module):
    """loss.py:24-47: ``sum(base(y*mask, yhat*mask)) / count(mask != 0)`` — the loss over the pixels a mask touches,
    micro-averaged.  ``baseloss`` is a loss *class* accepting ``reduction='none'`` (``nn.L1Loss`` or ``nn.MSELoss``);
    like in the reference, ``LocalLoss(RMSELoss)`` raises ``TypeError`` because ``RMSELoss.__init__`` takes no
    ``reduction`` argument."""

    def __init__(self, baseloss, eps=1e-16):
        super().__init__()
        self.loss = baseloss(reduction="none")
        if isinstance(self.loss, nn.L1Loss):
            self._kind = RECON_LOCAL_L1
        elif isinstance(self.loss, nn.MSELoss):
            self._kind = RECON_LOCAL_MSE
        else:
            raise NotImplementedError("LocalLoss supports nn.L1Loss and nn.MSELoss bases")
        self.eps = eps

    def forward(self, yhat, y, mask):
        return _ReconFunction.apply(self._kind, y, yhat, mask, self.eps)

    @staticmethod
    def affected_pixels(yhat, y, mask) -> int:
        """Exact ``(mask != 0)`` pixel count as computed by the kernel (bit-exact parity check)."""
        a, b, m = y.detach().contiguous().float(), yhat.detach().contiguous().float(), mask.detach().contiguous().float()
        _check(a, b, m)
        out = torch.empty(4, dtype=torch.float32, device=a.device)
        call("b200_recon_loss", RECON_LOCAL_L1, ptr(a), ptr(b), ptr(m), a.numel(), 0.0, ptr(Scratch.get(a.device).recon), ptr(out), stream())
        words = out.view(torch.int32)[2:4].tolist()
        return int(words[0]) + (int(words[1]) << 31)


class _TvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, weight):
        x = img.detach().contiguous().float()
        if not x.is_cuda:
            raise _lib.B200Error("tv_loss runs on CUDA tensors only; there is no CPU fallback")
        if x.dim() < 2:
            raise _lib.B200Error(f"tv_loss expects [..., H, W], got {tuple(x.shape)}")
        H, W = x.shape[-2:]
        planes = x.numel() // (H * W)
        partial = torch.empty(2 * planes * H, dtype=torch.float32, device=x.device)
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        call("b200_tv_loss", ptr(x), planes, H, W, float(weight), ptr(partial), ptr(out), stream())
        ctx.x, ctx.weight, ctx.shape, ctx.dims = x, float(weight), img.shape, (planes, H, W)
        return out[0].clone()

    @staticmethod
    def backward(ctx, gout):
        if not ctx.needs_input_grad[0]:
            return None, None
        g = gout.contiguous().float().reshape(1)
        d = torch.empty_like(ctx.x)
        planes, H, W = ctx.dims
        call("b200_tv_loss_bwd", ptr(ctx.x), planes, H, W, ctx.weight, ptr(g), ptr(d), 0, stream())
        return d.view(ctx.shape), None


def tv_loss(img, tv_weight):
    """loss.py:138-151: ``tv_weight * (mean((img[..., :-1] - img[..., 1:])**2) + mean((img[..., :-1, :] - img[..., 1:, :])**2))``;
    differentiable w.r.t. ``img`` (it is part of ``g_loss``, wgan_perceptual_style_faceparsing.py:219-221)."""
    return _TvFunction.apply(img, tv_weight)


# ------------------------------------------------------------------------------------------------------------------
# VGG19 perceptual / style losses (loss.py:49-136)
# ------------------------------------------------------------------------------------------------------------------
vgg = None  # the reference's module-level ``vgg = VGG19Wrapper().to(device)`` (loss.py:4-5), built on first use


def set_vgg(wrapper):
    """Install the ``networks.VGG19Wrapper`` the feature losses use (e.g. one loaded with trained torchvision weights)."""
    global vgg
    vgg = wrapper
    return wrapper


def _vgg(device):
    global vgg
    if vgg is None:
        from . import networks

        vgg = networks.VGG19Wrapper().to(device)
        vgg.eval()
    return vgg


def _feature_losses(output, target, want_p: bool, want_s: bool, dst=None, weights=(1.0, 1.0)):
    """weights x (sum over the five layers of mean((f_o - f_t)^2), sum of mean((gram(f_o) - gram(f_t))^2)) as two device scalars
    (written into ``dst`` when given: a contiguous fp32 view of 2 elements)."""
    eng = _vgg(output.device)._get_engine()
    B = output.shape[0]
    with torch.no_grad():
        # one pass over [output; target]: the weights are read once per layer
        feats = eng.features_of(torch.cat([output.detach().reshape(B, 1, *output.shape[-2:]), target.detach().reshape(B, 1, *target.shape[-2:])], 0).float())
        dev = output.device
        out = torch.zeros(2, dtype=torch.float32, device=dev) if dst is None else dst.zero_()
        wp, ws = float(weights[0]), float(weights[1])
        part = torch.empty(_lib.value("b200_sqdiff_blocks"), dtype=torch.float32, device=dev)
        for f in feats:
            _, H, W, C = f.shape
            fo, ft = f[:B], f[B:]
            n = fo.numel()
            if want_p:
                call("b200_sqdiff_bf16", ptr(fo), ptr(ft), n, wp / n, ptr(part), ptr(out), 1, stream())
            if want_s:
                Cp, S = max(C, 128), _lib.value("b200_gram_splits", 2 * B, H, W, C)
                G = torch.empty((2 * B, S, Cp, C), dtype=torch.float32, device=dev)
                call("b200_gram_gemm", ptr(f), 2 * B, H, W, C, C, ptr(G), stream())
                call("b200_sqdiff_gram", ptr(G[:B]), ptr(G[B:]), B, S, Cp, C, 1.0 / (H * W * C), ws / (B * C * C), ptr(part), ctypes_ptr(out, 1), 1, stream())
    return out


def ctypes_ptr(t, index):
    import ctypes

    return ctypes.c_void_p(t.data_ptr() + index * t.element_size())


def perceptual_loss(output, target, weight=0.05):
    """loss.py:49-69: ``weight * sum_l mean((vgg_l(output) - vgg_l(target))^2)`` over relu{1..5}_1; no gradient."""
    return weight * _feature_losses(output, target, True, False)[0]


def style_loss(output, target, weight=0.1):
    """loss.py:71-91: ``weight * sum_l mean((gram(vgg_l(output)) - gram(vgg_l(target)))^2)``; no gradient."""
    return weight * _feature_losses(output, target, False, True)[1]


def perceptual_and_style_loss(output, target, weight_p=0.05, weight_s=100):
    """loss.py:93-115 (called with weight_p = weight_s = 0.01 at wgan_perceptual_style_faceparsing.py:216): both losses from ONE
    pass of the feature extractor over [output; target]."""
    v = _feature_losses(output, target, True, True)
    return weight_p * v[0], weight_s * v[1]


def gram_matrix(features, normalize=True):
    """loss.py:117-136: ``features`` [N,C,H,W] -> [N,C,C] Gram matrices (divided by H*W*C when ``normalize``), on the tensor cores.
    C must be a multiple of 64 and H*W >= 64 (every VGG19 layer the losses use)."""
    if not features.is_cuda:
        raise _lib.B200Error("gram_matrix runs on CUDA tensors only; there is no CPU fallback")
    N, C, H, W = features.shape
    f = features.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    Cp, S = max(C, 128), _lib.value("b200_gram_splits", N, H, W, C)
    G = torch.empty((N, S, Cp, C), dtype=torch.float32, device=features.device)
    call("b200_gram_gemm", ptr(f), N, H, W, C, C, ptr(G), stream())
    g = G[:, :, :C, :].sum(1)
    return g / (H * W * C) if normalize else g

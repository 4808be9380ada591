"""Shared helpers of the parity tests."""
import torch


def rel(a, b):
    """rel-L2 of a against the reference b."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def summary_close(t, s, tol):
    """Compare a tensor with a golden `summarize()` fingerprint (norm + 64 strided samples)."""
    f = t.detach().float().cpu().flatten()
    assert list(t.shape) == s["shape"], (list(t.shape), s["shape"])
    idx = torch.linspace(0, f.numel() - 1, steps=min(f.numel(), 64)).long()
    norm_err = abs(float(f.norm()) - s["norm"]) / max(s["norm"], 1e-30)
    samp = s["sample"]
    samp_err = float((f[idx] - samp).norm() / samp.norm().clamp_min(1e-30)) if float(samp.norm()) > 0 else float((f[idx] - samp).norm())
    return norm_err, samp_err

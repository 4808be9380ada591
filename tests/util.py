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


def cos(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def _t(x):
    """NHWC bf16 device buffer -> NCHW fp32 CPU."""
    return x.detach().float().permute(0, 3, 1, 2).contiguous().cpu()


def stored_from_gen_ctx(c, widths, scope=""):
    """Storage-point tensors of one generator forward, keyed like oracle.nets.unet_forward's policy tags."""
    pre = scope + "/" if scope else ""
    n = len(widths)
    out = {}
    for l in range(1, n):
        out[f"{pre}a{l}"] = _t(c.a[l])
    for l in range(2, n + 1):
        out[f"{pre}y{l}"] = _t(c.y[l])
        out[f"{pre}yu{l}"] = _t(c.yu[l])
        out[f"{pre}s{l}"] = _t(c.cat[l - 1][..., widths[l - 2]:])
    return out


def stored_from_disc_ctx(c, scope=""):
    pre = scope + "/" if scope else ""
    out = {f"{pre}a{i}": _t(c.a[i]) for i in range(4)}
    out.update({f"{pre}y{i}": _t(c.y[i]) for i in range(1, 4)})
    return out

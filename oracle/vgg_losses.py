"""fp32 restatement of the reference's VGG19 perceptual / style losses (lib/models/loss.py:49-136).

TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this.
Pinned against the reference's own functions by oracle/pin_vgg_against_reference.py (tests/golden/reference_vgg.pt).

``features`` is ``torchvision.models.vgg19().features`` (networks.py:367-389 wraps exactly that).  ``bf16=True`` emulates the
storage precision of the CUDA path (bf16 weights and bf16 activations between layers, fp32 accumulation) so that tests can tell
kernel errors from rounding.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

FEATURE_LAYERS = (1, 6, 11, 20, 29)  # loss.py:54: relu1_1 .. relu5_1


def gram_matrix(features: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """loss.py:117-136: [N,C,H,W] -> [N,C,C], divided by H*W*C when ``normalize``."""
    n, c, h, w = features.shape
    f = features.reshape(n, c, h * w)
    g = torch.bmm(f, f.transpose(1, 2))
    return g / (h * w * c) if normalize else g


def _q(t: torch.Tensor, bf16: bool) -> torch.Tensor:
    return t.to(torch.bfloat16).float() if bf16 else t


def feature_maps(features, img: torch.Tensor, bf16: bool = False):
    """The walk of loss.py:55-67 over ``features`` (every module, in order) on ``img.repeat(1,3,1,1)``; returns the activations
    at FEATURE_LAYERS.  Modules past the last used layer are skipped (their outputs are never read)."""
    x = img.detach().float().repeat(1, 3, 1, 1)
    outs = []
    mods = list(features)
    i = 0
    while i <= max(FEATURE_LAYERS):
        m = mods[i]
        if isinstance(m, torch.nn.Conv2d):
            first = i == 0
            w = m.weight.detach().float()
            # the CUDA path folds the three identical input channels of the first layer in fp32; later layers hold bf16 weights
            x = F.conv2d(x, w if first else _q(w, bf16), m.bias.detach().float(), stride=1, padding=1)
            if isinstance(mods[i + 1], torch.nn.ReLU):  # ReLU(inplace=True): the conv output is never observed on its own
                x = torch.relu(x)
                i += 1
            x = _q(x, bf16)
        elif isinstance(m, torch.nn.MaxPool2d):
            x = F.max_pool2d(x, 2, 2)
        elif isinstance(m, torch.nn.ReLU):
            x = torch.relu(x)
        else:
            raise TypeError(type(m))
        if i in FEATURE_LAYERS:
            outs.append(x)
        i += 1
    return outs


def perceptual_and_style_loss(features, output, target, weight_p=0.05, weight_s=100, bf16: bool = False):
    """loss.py:93-115."""
    fo, ft = feature_maps(features, output, bf16), feature_maps(features, target, bf16)
    p = sum(torch.mean((a - b) ** 2) for a, b in zip(fo, ft))
    s = sum(torch.mean((gram_matrix(a) - gram_matrix(b)) ** 2) for a, b in zip(fo, ft))
    return weight_p * p, weight_s * s


def perceptual_loss(features, output, target, weight=0.05, bf16: bool = False):
    """loss.py:49-69."""
    return perceptual_and_style_loss(features, output, target, weight, 0.0, bf16)[0]


def style_loss(features, output, target, weight=0.1, bf16: bool = False):
    """loss.py:71-91."""
    return perceptual_and_style_loss(features, output, target, 0.0, weight, bf16)[1]


def seeded_vgg_features(seed: int = 0):
    """torchvision's VGG19 feature stack with its default random initialisation under ``seed`` (the reference downloads trained
    weights, networks.py:371; there is no network here, and the loss arithmetic does not depend on what the weights are)."""
    from torchvision import models

    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = models.vgg19(weights=None).features.eval()
    torch.random.set_rng_state(gen_state)
    return net

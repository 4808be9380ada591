"""Functional fp32 restatement of the reference generator / discriminator.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parameters live in a flat dict
whose keys are exactly the reference modules' ``state_dict()`` keys, so a dict
made here loads into ``/root/reference/lib/models/networks.py`` modules with
``load_state_dict`` and vice versa.

Reference anchors
-----------------
* UnetGenerator(1, 1, 7, ngf=64, BatchNorm2d, use_dropout='False')
  ``networks.py:17-19`` (factory), ``:222-244`` (inside-out construction),
  ``:264-318`` (per-block layer list), ``:320-324`` (skip = cat([x, model(x)])
  where x was already LeakyReLU'd *in place*).
* PatchGANDiscriminator ``networks.py:331-363`` (head ``Linear(25, 1)`` only
  fits 128x128; here the head is ``Linear((R/16-3)**2, 1)``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------
# precision policy (None == the reference's pure fp32 arithmetic)
# ----------------------------------------------------------------------------
class _RoundSTE(torch.autograd.Function):
    """Forward: round to bf16 (the tensor is *stored* in bf16); backward: identity."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """Forward: identity; backward: round to bf16 (the gradient w.r.t. this tensor is stored in bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class Fp32Policy:
    """The reference: nothing is rounded.  ``tag`` names the storage point (see ``TeacherForcedPolicy``)."""

    scope = ""

    def act(self, x, tag=None):
        return x

    def w(self, x):
        return x

    def grad(self, x):
        return x


class Bf16Policy(Fp32Policy):
    """Where the B200 path stores bf16 (DESIGN.md "precision policy"): activations and activation gradients in HBM,
    and the tensor-core operand copies of the 4x4/s2 GEMM-layer weights.  Accumulation, BatchNorm statistics, the
    1-channel layers' weights (d1, u1, c1, the discriminator head), losses, weight gradients and optimizer state
    stay fp32.  Free-running: the same algorithm as the reference evaluated at the CUDA path's precision."""

    def act(self, x, tag=None):
        return _RoundSTE.apply(x)

    def w(self, x):
        return _RoundSTE.apply(x)

    def grad(self, x):
        return _RoundGrad.apply(x)


class TeacherForcedPolicy(Fp32Policy):
    """Checks a device implementation *layer by layer* and its backward pass *at the implementation's own forward
    state*.  ``stored`` maps storage-point tags ("<scope>/a3", "<scope>/y4", ...) to the tensors the implementation
    produced (NCHW fp32 copies of its bf16 buffers).  At every storage point the oracle

      1. records rel-L2(implementation tensor, oracle value computed from the implementation's *previous* tensor)
         in ``self.errors[tag]`` — a one-layer error, no accumulation;
      2. continues from the implementation's tensor (straight-through substitution), so every ReLU / LeakyReLU /
         dropout mask and every BatchNorm statistic of the oracle's autograd graph is the implementation's.

    Gradients of that graph are then the exact fp32 vector-Jacobian product of the implementation's forward pass —
    what its backward kernels must reproduce up to bf16 storage of activation gradients.  This removes the
    mask-flip noise that makes free-running bf16-vs-fp32 gradient comparisons ill-conditioned."""

    def __init__(self, stored, round_weights=True):
        self.stored = stored
        self.errors = {}
        self.round_weights = round_weights
        self.scope = ""

    def act(self, x, tag=None):
        key = f"{self.scope}/{tag}" if self.scope else tag
        if tag is None or key not in self.stored:
            return x
        mine = self.stored[key].to(x.dtype)
        self.errors[key] = float((mine - x.detach()).norm() / x.detach().norm().clamp_min(1e-30))
        return x + (mine - x).detach()

    def w(self, x):
        return _RoundSTE.apply(x) if self.round_weights else x


# ----------------------------------------------------------------------------
# layer specs
# ----------------------------------------------------------------------------


def unet_channel_plan(num_downs: int = 7, ngf: int = 64, input_nc: int = 1, output_nc: int = 1):
    """Encoder/decoder channel plan of networks.py:236-243.

    Returns ``enc`` = [(cin, cout)] for d1..dN and ``dec`` = [(cin, cout)] for
    uN..u1 (decoder listed innermost first, i.e. in execution order).
    """
    widths = [ngf, ngf * 2, ngf * 4] + [ngf * 8] * (num_downs - 3)
    enc = []
    cin = input_nc
    for w in widths:
        enc.append((cin, w))
        cin = w
    dec = []
    for lvl in range(num_downs, 0, -1):  # lvl = N (innermost) .. 1 (outermost)
        inner = widths[lvl - 1]
        outer = output_nc if lvl == 1 else widths[lvl - 2]
        cin_t = inner if lvl == num_downs else inner * 2
        dec.append((cin_t, outer))
    return enc, dec


def unet_prefix(level: int) -> str:
    """state_dict prefix of the UnetSkipConnectionBlock at nesting ``level``.

    level 1 = outermost (``model``), level 2 = ``model.model.1`` and every
    deeper block sits at index 3 of its parent (after [lrelu, conv, norm])
    — networks.py:296-316.
    """
    p = "model"
    if level >= 2:
        p += ".model.1"
    for _ in range(level - 2):
        p += ".model.3"
    return p


def unet_keys(level: int, num_downs: int = 7):
    """Names of (down conv, down norm, up conv, up norm) inside block ``level``."""
    base = unet_prefix(level) + ".model."
    if level == 1:  # [downconv, sub, uprelu, upconv, tanh]
        return dict(down=base + "0", dnorm=None, up=base + "3", unorm=None)
    if level == num_downs:  # [lrelu, downconv, uprelu, upconv, upnorm]
        return dict(down=base + "1", dnorm=None, up=base + "3", unorm=base + "4")
    # [lrelu, downconv, downnorm, sub, uprelu, upconv, upnorm, (dropout)]
    return dict(down=base + "1", dnorm=base + "2", up=base + "5", unorm=base + "6")


def unet_dropout_levels(num_downs: int = 7) -> List[int]:
    """Levels whose block ends in Dropout(0.5) when use_dropout is truthy.

    networks.py:238-239: the ``num_downs - 5`` intermediate ngf*8 blocks, i.e.
    levels num_downs-1 .. 5 (two blocks for num_downs = 7).  The factory passes
    the *string* 'False' (truthy) so they are ON by default.
    """
    return list(range(num_downs - 1, 3, -1))[: max(num_downs - 5, 0)]


# ----------------------------------------------------------------------------
# initialisation (PyTorch defaults; the reference defines no custom init)
# ----------------------------------------------------------------------------


def _uniform(shape, bound, gen):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound


def _bn_params(p: Params, key: str, c: int):
    p[key + ".weight"] = torch.ones(c)
    p[key + ".bias"] = torch.zeros(c)
    p[key + ".running_mean"] = torch.zeros(c)
    p[key + ".running_var"] = torch.ones(c)
    p[key + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def init_unet(seed: int = 0, num_downs: int = 7, ngf: int = 64, input_nc: int = 1, output_nc: int = 1) -> Params:
    """kaiming_uniform(a=sqrt 5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)); BN gamma=1, beta=0.

    ConvTranspose2d weights are (Cin, Cout, 4, 4) and torch computes their
    fan_in from dim 1 (Cout*16).
    """
    g = torch.Generator().manual_seed(seed)
    enc, dec = unet_channel_plan(num_downs, ngf, input_nc, output_nc)
    p: Params = {}
    # key order == the reference's registration order: d1, d2, bn2, ... dN, uN, bnN, ... u1
    for lvl in range(1, num_downs + 1):
        k = unet_keys(lvl, num_downs)
        cin, cout = enc[lvl - 1]
        p[k["down"] + ".weight"] = _uniform((cout, cin, 4, 4), 1.0 / math.sqrt(cin * 16), g)
        if k["dnorm"]:
            _bn_params(p, k["dnorm"], cout)
    for lvl in range(num_downs, 0, -1):
        k = unet_keys(lvl, num_downs)
        tin, tout = dec[num_downs - lvl]
        p[k["up"] + ".weight"] = _uniform((tin, tout, 4, 4), 1.0 / math.sqrt(tout * 16), g)
        if lvl == 1:  # only the outermost ConvTranspose2d keeps its bias (networks.py:293-295)
            p[k["up"] + ".bias"] = _uniform((tout,), 1.0 / math.sqrt(tout * 16), g)
        if k["unorm"]:
            _bn_params(p, k["unorm"], tout)
    return p


PATCHGAN_CONVS = [("model.0", None), ("model.2", "model.3"), ("model.5", "model.6"), ("model.8", "model.9")]


def patchgan_head_features(res: int) -> int:
    return (res // 16 - 3) ** 2


def init_patchgan(seed: int = 0, res: int = 128, c: int = 1) -> Params:
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    cin = c
    for i, (ck, bk) in enumerate(PATCHGAN_CONVS):
        cout = 64 << i
        p[ck + ".weight"] = _uniform((cout, cin, 4, 4), 1.0 / math.sqrt(cin * 16), g)
        if bk:
            _bn_params(p, bk, cout)
        cin = cout
    p["model.11.weight"] = _uniform((1, cin, 4, 4), 1.0 / math.sqrt(cin * 16), g)
    nf = patchgan_head_features(res)
    p["model.13.weight"] = _uniform((1, nf), 1.0 / math.sqrt(nf), g)
    p["model.13.bias"] = _uniform((1,), 1.0 / math.sqrt(nf), g)
    return p


def trainable(p: Params) -> Params:
    """Float leaves that the optimizers see (everything but BN running buffers)."""
    return {k: v for k, v in p.items() if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))}


# ----------------------------------------------------------------------------
# forward passes
# ----------------------------------------------------------------------------


def _bn(p: Params, key: str, x, train: bool, taps: Optional[dict]):
    out = F.batch_norm(
        x,
        p[key + ".running_mean"],
        p[key + ".running_var"],
        p[key + ".weight"],
        p[key + ".bias"],
        training=train,
        momentum=0.1,
        eps=1e-5,
    )
    if train:
        p[key + ".num_batches_tracked"] += 1
    return out


def unet_forward(
    p: Params,
    x: torch.Tensor,
    train: bool = True,
    num_downs: int = 7,
    dropout_masks: Optional[Dict[int, torch.Tensor]] = None,
    taps: Optional[dict] = None,
    policy=None,
) -> torch.Tensor:
    """UnetGenerator.forward, networks.py:246-253 + 320-324.

    ``dropout_masks``: {level: keep-mask of the block's up-norm output}; the
    kept values are scaled by 2 (Dropout(0.5)).  ``None`` == dropout disabled
    (``use_dropout=False``) — parity runs use that, SURVEY.md D8.
    ``taps``: optional dict filled with named intermediate activations.
    ``policy``: precision policy (default ``Fp32Policy`` == the reference).
    """
    P = policy or Fp32Policy()
    skips = []  # post-LeakyReLU encoder activations (what the in-place op leaves in the cat, D9)
    h = x
    for lvl in range(1, num_downs + 1):
        k = unet_keys(lvl, num_downs)
        if lvl > 1:
            h = P.act(F.leaky_relu(h, 0.2), f"a{lvl - 1}")
            skips.append(h)
            h = P.grad(h)
        w = p[k["down"] + ".weight"]
        h = F.conv2d(h, P.w(w) if lvl > 1 else w, None, stride=2, padding=1)
        h = P.grad(P.act(h, f"y{lvl}") if lvl > 1 else h)
        if taps is not None:
            taps[f"d{lvl}.conv"] = h
        if k["dnorm"]:
            h = _bn(p, k["dnorm"], h, train, taps)
    for lvl in range(num_downs, 0, -1):
        k = unet_keys(lvl, num_downs)
        h = P.grad(F.relu(h))  # uprelu (in place on the child's cat in the reference)
        if taps is not None:
            taps[f"u{lvl}.in"] = h
        w = p[k["up"] + ".weight"]
        h = F.conv_transpose2d(h, P.w(w) if lvl > 1 else w, p.get(k["up"] + ".bias"), stride=2, padding=1)
        if lvl > 1:
            h = P.grad(P.act(h, f"yu{lvl}"))
        if taps is not None:
            taps[f"u{lvl}.conv"] = h
        if lvl == 1:
            return torch.tanh(h)
        h = _bn(p, k["unorm"], h, train, taps)
        if dropout_masks is not None and lvl in dropout_masks:
            h = h * dropout_masks[lvl] * 2.0
        # relu(round(x)) == round(relu(x)): the decoder half is stored post-ReLU by the CUDA path; the parent's uprelu
        # (top of this loop) is idempotent on it
        h = torch.cat([skips[lvl - 2], P.act(F.relu(h), f"s{lvl}")], 1)  # skip first, decoder second
    raise AssertionError("unreachable")


def patchgan_forward(p: Params, x: torch.Tensor, train: bool = True, sigmoid: bool = True, taps: Optional[dict] = None, policy=None):
    """PatchGANDiscriminator.forward, networks.py:337-363 -> [B, 1]."""
    P = policy or Fp32Policy()
    h = x
    for i, (ck, bk) in enumerate(PATCHGAN_CONVS):
        w = p[ck + ".weight"]
        h = F.conv2d(h, P.w(w) if i > 0 else w, None, stride=2, padding=1)
        h = P.grad(P.act(h, f"y{i}") if i > 0 else h)
        if taps is not None:
            taps[f"c{i + 1}.conv"] = h
        if bk:
            h = _bn(p, bk, h, train, taps)
        h = P.grad(P.act(F.leaky_relu(h, 0.2), f"a{i}"))
    h = F.conv2d(h, p["model.11.weight"], None, stride=1, padding=0)
    if taps is not None:
        taps["c5.conv"] = h
    h = F.linear(h.flatten(1), p["model.13.weight"], p["model.13.bias"])
    if sigmoid:
        h = torch.sigmoid(h)
    return h.view(-1, 1)


def conv_flops_unet(res: int, num_downs: int = 7, ngf: int = 64) -> float:
    """Forward conv FLOPs (2*MAC) per image — SURVEY.md §2.1 / §8(d)."""
    enc, dec = unet_channel_plan(num_downs, ngf)
    fl, r = 0.0, res
    for cin, cout in enc:
        r //= 2
        fl += 2.0 * r * r * cout * cin * 16
    for cin, cout in dec:
        fl += 2.0 * r * r * cin * cout * 16  # every input pixel meets all 16 taps
        r *= 2
    return fl


def conv_flops_patchgan(res: int) -> float:
    fl, r, cin = 0.0, res, 1
    for i in range(4):
        r //= 2
        fl += 2.0 * r * r * (64 << i) * cin * 16
        cin = 64 << i
    r -= 3
    fl += 2.0 * r * r * cin * 16 + 2.0 * r * r
    return fl

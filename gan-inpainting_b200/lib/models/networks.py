"""Drop-in ``lib.models.networks``: same constructors, ``state_dict`` keys / shapes and call behaviour as the
reference (lib/models/networks.py), executed by the sm_100a kernels of ``gan_inpainting_b200.engine``.

The module tree below exists to OWN the parameters in the reference's checkpoint layout (SURVEY.md §2.1:
``model.model.1.model.3…`` nesting, Conv2d ``(Cout,Cin,4,4)``, ConvTranspose2d ``(Cin,Cout,4,4)``, BatchNorm
buffers), so ``epochN_G.pt`` files interchange and ``optim.*(net.parameters())``, ``p.data.clamp_`` and the
``"bias" in name`` filters of the experiment scripts keep working.  The layers are never executed one by one:
``UnetGenerator.forward`` / ``PatchGANDiscriminator.forward`` hand the whole network to one autograd node that
runs the fused CUDA forward and, on ``.backward()``, the fused CUDA backward.

There is no CPU path: calling a network on a CPU tensor raises.

Scope: ``UnetGenerator(1, 1, num_downs, ngf=64k, BatchNorm2d)`` and ``PatchGANDiscriminator(c=1)``.
``VGG19Generator`` / ``DCGANDiscriminator`` (non-default alternatives, networks.py:49-212) and instance / no
normalisation are outside the hot path and raise ``NotImplementedError``.
"""
from __future__ import annotations

import functools
from typing import List

import torch
from torch import nn

from ... import _lib
from ...engine import DiscriminatorEngine, GeneratorEngine, Scratch
from ..._lib import call, ptr, stream


class Identity(nn.Module):
    """networks.py:9-11."""

    def forward(self, x):
        return x


def get_norm_layer(norm_type="instance"):
    """networks.py:30-45: 'batch' -> BatchNorm2d(affine, running stats); 'instance' -> InstanceNorm2d(no affine,
    no running stats); 'none' -> Identity."""
    if norm_type == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True)
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    if norm_type == "none":
        return lambda _channels: Identity()
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def get_network(type, name):
    """networks.py:13-28.  Note the reference passes the *string* ``use_dropout='False'`` (truthy), so the default
    generator has its two Dropout(0.5) layers switched ON; that is kept."""
    if type == "generator":
        if name == "unet":
            return UnetGenerator(1, 1, 7, ngf=64, norm_layer=get_norm_layer(norm_type="batch"), use_dropout="False")
        if name == "vgg19":
            raise NotImplementedError("VGG19Generator is outside the B200 hot path (SURVEY.md §2: non-default, 128x128 only)")
        raise Exception("Invalid generator network name")
    elif type == "discriminator":
        if name == "patchgan":
            return PatchGANDiscriminator()
        if name == "dcgan":
            raise NotImplementedError("DCGANDiscriminator is outside the B200 hot path (SURVEY.md §2: non-default, 128x128 only)")
        raise Exception("Invalid discriminator network name")


def _is_batchnorm_factory(norm_layer) -> bool:
    f = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    return f is nn.BatchNorm2d


def _wants_conv_bias(norm_layer) -> bool:
    f = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    return f is nn.InstanceNorm2d  # networks.py:279-282


# --------------------------------------------------------------------------------------------------------------
# generator
# --------------------------------------------------------------------------------------------------------------
class UnetSkipConnectionBlock(nn.Module):
    """One U-Net level as a parameter container; slot order inside ``self.model`` follows networks.py:285-318 so the
    state_dict indices agree:

        outermost : 0 down-conv | 1 child | 2 relu | 3 up-convT(bias) | 4 tanh
        innermost : 0 lrelu | 1 down-conv | 2 relu | 3 up-convT | 4 up-norm
        otherwise : 0 lrelu | 1 down-conv | 2 down-norm | 3 child | 4 relu | 5 up-convT | 6 up-norm [| 7 dropout]
    """

    def __init__(self, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False, norm_layer=nn.BatchNorm2d, use_dropout=False):
        super().__init__()
        self.outermost, self.innermost = outermost, innermost
        bias = _wants_conv_bias(norm_layer)
        cin = outer_nc if input_nc is None else input_nc
        k = dict(kernel_size=4, stride=2, padding=1)
        slots: List[nn.Module] = []
        if outermost:
            dconv = nn.Conv2d(cin, inner_nc, bias=bias, **k)
            uconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, **k)
            slots = [dconv, submodule, nn.ReLU(True), uconv, nn.Tanh()]
            dnorm = unorm = None
        elif innermost:
            dconv = nn.Conv2d(cin, inner_nc, bias=bias, **k)
            uconv = nn.ConvTranspose2d(inner_nc, outer_nc, bias=bias, **k)
            unorm = norm_layer(outer_nc)
            slots = [nn.LeakyReLU(0.2, True), dconv, nn.ReLU(True), uconv, unorm]
            dnorm = None
        else:
            dconv = nn.Conv2d(cin, inner_nc, bias=bias, **k)
            dnorm = norm_layer(inner_nc)
            uconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, bias=bias, **k)
            unorm = norm_layer(outer_nc)
            slots = [nn.LeakyReLU(0.2, True), dconv, dnorm, submodule, nn.ReLU(True), uconv, unorm]
            if use_dropout:
                slots.append(nn.Dropout(0.5))
        self.model = nn.Sequential(*slots)
        # plain-attribute handles for the engine (object.__setattr__: not registered twice in the module tree)
        object.__setattr__(self, "_parts", dict(down_conv=dconv, down_norm=dnorm, up_conv=uconv, up_norm=unorm, child=submodule,
                                                dropout=bool(use_dropout) and not (outermost or innermost)))

    def forward(self, x):
        raise _lib.B200Error("UnetSkipConnectionBlock is executed by UnetGenerator.forward as one fused CUDA graph of kernels; "
                             "there is no layer-by-layer (PyTorch) path")


class _UnetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, *params):
        eng = net._get_engine()
        c = eng.forward(x.detach().contiguous().float(), net.training, seed=net._next_seed())
        ctx.net, ctx.c = net, c
        if net.keep_last_context:
            object.__setattr__(net, "_last_ctx", c)
        return c.g_out.view(x.shape[0], 1, x.shape[-2], x.shape[-1])

    @staticmethod
    def backward(ctx, dout):
        net, c = ctx.net, ctx.c
        eng = net._get_engine()
        params = list(net.parameters())
        grads = {id(p): torch.empty_like(p, dtype=torch.float32) for p in params}
        dout = dout.contiguous().float()
        dpre = torch.empty_like(dout)
        sc = Scratch.get(dout.device)
        bias = eng.uc[0].bias
        call("b200_gstep_seed", None, None, ptr(c.g_out), None, ptr(dout), None, -1, 0.0, None, -1, 0.0, None, None, 0, dout.numel(), ptr(dpre), None,
             ptr(sc.recon), ptr(grads[id(bias)]) if bias is not None else None, 0, stream())
        eng.backward(c, dpre, grads, accumulate=False)
        ctx.c = None
        return (None, None) + tuple(grads[id(p)] if p.requires_grad else None for p in params)


class UnetGenerator(nn.Module):
    """networks.py:216-253 — built inside-out exactly like the reference so parameter registration order (and therefore
    ``parameters()`` / optimizer state order) matches."""

    def __init__(self, input_nc, output_nc, num_downs, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False, use_sigmoid_output=False):
        super().__init__()
        if num_downs < 5:
            raise ValueError("num_downs must be >= 5 (ngf, 2ngf, 4ngf, 8ngf levels + the innermost block)")
        widths = [ngf, ngf * 2, ngf * 4] + [ngf * 8] * (num_downs - 3)  # channels after down-conv of level 1..num_downs
        block = None
        for level in range(num_downs, 0, -1):
            inner = widths[level - 1]
            outer = output_nc if level == 1 else widths[level - 2]
            block = UnetSkipConnectionBlock(
                outer, inner, input_nc=input_nc if level == 1 else None, submodule=block, outermost=level == 1, innermost=level == num_downs,
                norm_layer=norm_layer, use_dropout=use_dropout if 5 <= level < num_downs else False,
            )
        self.model = block
        self.use_sigmoid_output = use_sigmoid_output
        object.__setattr__(self, "_cfg", dict(input_nc=input_nc, output_nc=output_nc, num_downs=num_downs, ngf=ngf, widths=widths,
                                              batchnorm=_is_batchnorm_factory(norm_layer)))
        object.__setattr__(self, "_engine", None)
        object.__setattr__(self, "_calls", 0)
        self.dropout_seed = 0x5EED
        self.keep_last_context = False  # debugging / parity tests: keep the activations of the last forward in _last_ctx

    # ---- engine plumbing
    def _levels(self):
        lv, b = [], self.model
        while b is not None:
            lv.append(b._parts)
            b = b._parts["child"]
        return lv

    def _get_engine(self) -> GeneratorEngine:
        if self._engine is None:
            cfg = self._cfg
            if cfg["input_nc"] != 1 or cfg["output_nc"] != 1:
                raise NotImplementedError("the B200 hot path covers the 1-channel inpainting generator (input_nc = output_nc = 1)")
            if not cfg["batchnorm"]:
                raise NotImplementedError("the B200 hot path covers BatchNorm2d generators (get_norm_layer('batch'))")
            lv = self._levels()
            eng = GeneratorEngine([p["down_conv"] for p in lv], [p["down_norm"] for p in lv], [p["up_conv"] for p in lv], [p["up_norm"] for p in lv],
                                  [i + 1 for i, p in enumerate(lv) if p["dropout"]], cfg["widths"])
            object.__setattr__(self, "_engine", eng)
        return self._engine

    def _next_seed(self) -> int:
        object.__setattr__(self, "_calls", self._calls + 1)
        return (self.dropout_seed * 1000003 + self._calls) & 0x0FFFFFFF

    def forward(self, input):
        if not input.is_cuda:
            raise _lib.B200Error("UnetGenerator runs on CUDA (sm_100a) only; there is no CPU fallback")
        if input.dim() != 4 or input.shape[1] != 1:
            raise _lib.B200Error(f"expected NCHW input with 1 channel, got {tuple(input.shape)}")
        if input.requires_grad:
            raise NotImplementedError("gradient w.r.t. the generator input is not part of the train step (the input is the masked image)")
        out = _UnetFunction.apply(input, self, *self.parameters())
        if self.use_sigmoid_output:
            out = torch.sigmoid(out)
        return out


# --------------------------------------------------------------------------------------------------------------
# discriminator
# --------------------------------------------------------------------------------------------------------------
class Flatten(nn.Module):
    """networks.py:326-328."""

    def forward(self, input):
        return input.view(input.size(0), -1)


class _PatchGANFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, *params):
        eng = net._get_engine()
        c = eng.forward(x.detach().contiguous().float(), net.training, net._sigmoid)
        ctx.net, ctx.c = net, c
        if net.keep_last_context:
            object.__setattr__(net, "_last_ctx", c)
        return c.pred

    @staticmethod
    def backward(ctx, dpred):
        net, c = ctx.net, ctx.c
        eng = net._get_engine()
        params = list(net.parameters())
        need_x = ctx.needs_input_grad[0]
        need_p = any(ctx.needs_input_grad[2:])
        dpred = dpred.contiguous().float()
        dlin = torch.empty((c.B,), dtype=torch.float32, device=dpred.device)
        call("b200_head_dlin", ptr(dpred), ptr(c.pred), int(c.sigmoid), c.B, ptr(dlin), stream())
        grads = {id(p): torch.empty_like(p, dtype=torch.float32) for p in params} if need_p else None
        dx = eng.backward(c, dlin, grads, accumulate=False, input_grad=need_x)
        ctx.c = None
        gp = tuple((grads[id(p)] if (need_p and p.requires_grad) else None) for p in params)
        return (dx, None) + gp


class PatchGANDiscriminator(nn.Module):
    """networks.py:331-363.  ``n_d_channel`` is ignored like in the reference (:334 hard-codes 64).

    Extension: the reference head ``Linear(25, 1)`` only fits 128x128 inputs; ``imagedim`` sizes it as
    ``Linear((imagedim/16 - 3)**2, 1)`` so the 256x256 / 512x512 configurations run.  ``imagedim=128`` is the reference."""

    def __init__(self, c=1, n_d_channel=64, sigmoid=True, imagedim=128):
        super().__init__()
        nd = 64
        conv = functools.partial(nn.Conv2d, kernel_size=(4, 4), stride=2, padding=1, bias=False)
        layers = [conv(c, nd), nn.LeakyReLU(0.2)]
        for i in range(3):
            layers += [conv(nd << i, nd << (i + 1)), nn.BatchNorm2d(nd << (i + 1)), nn.LeakyReLU(0.2)]
        layers += [nn.Conv2d(nd * 8, 1, kernel_size=(4, 4), stride=1, padding=0, bias=False), Flatten(), nn.Linear((imagedim // 16 - 3) ** 2, 1)]
        if sigmoid:
            layers.append(nn.Sigmoid())
        self.model = nn.Sequential(*layers)
        object.__setattr__(self, "_sigmoid", bool(sigmoid))
        self.keep_last_context = False
        object.__setattr__(self, "_c", c)
        object.__setattr__(self, "_engine", None)

    def _get_engine(self) -> DiscriminatorEngine:
        if self._engine is None:
            if self._c != 1:
                raise NotImplementedError("the B200 hot path covers the 1-channel discriminator (c = 1)")
            m = self.model
            object.__setattr__(self, "_engine", DiscriminatorEngine([m[0], m[2], m[5], m[8]], [None, m[3], m[6], m[9]], m[11], m[13]))
        return self._engine

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.B200Error("PatchGANDiscriminator runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.dim() != 4 or x.shape[1] != 1:
            raise _lib.B200Error(f"expected NCHW input with 1 channel, got {tuple(x.shape)}")
        output = _PatchGANFunction.apply(x, self, *self.parameters())
        return output.view(-1, 1)


class VGG19Wrapper(nn.Module):
    """networks.py:367-389: a torchvision VGG19 whose relu{1..5}_1 activations feed the perceptual / style losses
    (lib/models/loss.py:49-115).  The reference loads ``vgg19(pretrained=True)`` (a download); there is no network here, so the
    default is random initialisation (``pretrained=False``) — pass a ``state_dict`` of torchvision's VGG19 to use trained weights.
    ``forward`` (the full classifier, never called by the reference's losses) is not built; the losses use
    ``lib.models.loss.perceptual_and_style_loss``, which runs ``features[0..29]`` on the sm_100a kernels."""

    def __init__(self, pretrained: bool = False, state_dict=None):
        super().__init__()
        from torchvision import models

        if pretrained and state_dict is None:
            raise _lib.B200Error("VGG19Wrapper(pretrained=True) needs a download; pass state_dict=<torchvision vgg19 weights> instead")
        self.feature_maps = {}
        self.vgg19 = models.vgg19(weights=None)
        if state_dict is not None:
            self.vgg19.load_state_dict(state_dict)
        self._engine = None

    def _get_engine(self):
        from ...engine import VggFeatureEngine

        if self._engine is None:
            self._engine = VggFeatureEngine(self.vgg19.features)
        return self._engine

    def forward(self, x):
        raise NotImplementedError("the VGG19 classifier head is not on the hot path (the reference's losses only walk vgg19.features)")

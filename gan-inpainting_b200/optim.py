"""Fused optimizers with the ``torch.optim`` constructor surface the reference's experiment scripts use
(``optim.RMSprop(net.parameters(), lr=0.00005)`` wgan_l1.py:64-65; ``optim.Adam(net.parameters(), lr=0.0002,
betas=(0.5, 0.999))`` minimaxgan_rmse.py:64-65), running on the kernels of csrc/optim.cu.

``flatten_parameters(module)`` re-homes a network's parameters (and their ``.grad``) into ONE fp32 buffer each,
so an optimizer step — and the WGAN weight clip (wgan_l1.py:152-153) — is a single launch over the whole network,
and the data-parallel gradient all-reduce can bucket by slicing that buffer.  Parameters that were not flattened
are still handled (one launch per tensor).
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import call, ptr, stream

_ALIGN = 4  # elements (16 bytes)


class FlatGroup:
    """One network's parameters as views into ``params`` and their gradients as views into ``grads``."""

    def __init__(self, module: torch.nn.Module):
        ps = list(module.parameters())
        dev = ps[0].device
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = n
        self.params = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.members = ps
        self.offsets = offs
        self.names = [k for k, _ in module.named_parameters()]
        for p, o in zip(ps, offs):
            self.params[o : o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.params[o : o + p.numel()].view(p.shape)
            p.grad = self.grads[o : o + p.numel()].view(p.shape)
        self.grad_views = [p.grad for p in ps]
        module._b200_flat_group = self  # the module owns its flat group; the registry below only holds weak references
        ref = weakref.ref(self)
        _FLAT.update({id(p): ref for p in ps})

    def grad_of(self, p) -> torch.Tensor:
        return self.grad_views[self._index(p)]

    def _index(self, p):
        for i, q in enumerate(self.members):
            if q is p:
                return i
        raise KeyError("parameter is not part of this flat group")

    def grad_map(self) -> Dict[int, torch.Tensor]:
        return {id(p): g for p, g in zip(self.members, self.grad_views)}

    def check_alias(self):
        """The parameters must still BE the views into ``params``: ``net.to(other device / dtype)`` or
        ``load_state_dict(assign=True)`` replaces ``p.data``, after which the fused optimizer, the WGAN clip and the all-reduce
        buckets would update a buffer the module no longer reads — training would silently stop.  Cheap (one pointer compare per
        parameter); re-homes a detached parameter when only its storage moved (same device, dtype and shape), raises otherwise."""
        esz = self.params.element_size()
        for p, o in zip(self.members, self.offsets):
            if p.data_ptr() != self.params.data_ptr() + o * esz:
                if p.device != self.params.device or p.dtype != self.params.dtype:
                    raise _lib.B200Error("a flattened parameter was moved to another device / dtype (net.to(...) after flatten_parameters): "
                                         "re-create the optimizer / trainer after moving the network")
                view = self.params[o : o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                _lib.WEIGHT_EPOCH += 1

    def rebind_grads(self):
        """Bring ``p.grad`` back to the flat views (after ``zero_grad(set_to_none=True)`` or an autograd re-allocation)."""
        self.check_alias()
        for p, g in zip(self.members, self.grad_views):
            if p.grad is None:
                g.zero_()
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g

    def zero_grads(self):
        self.grads.zero_()
        for p, g in zip(self.members, self.grad_views):
            p.grad = g

    def logging_offsets(self, skip_bias=True):
        """(names, int64 offsets tensor of len n+1 pairs flattened) for the mean-|grad| kernel (wgan_l1.py:195-197)."""
        names, spans = [], []
        for name, p, o in zip(self.names, self.members, self.offsets):
            if skip_bias and "bias" in name:
                continue
            names.append(name)
            spans.append((o, o + p.numel()))
        return names, spans


_FLAT: Dict[int, "weakref.ref[FlatGroup]"] = {}


def _flat_of(p) -> Optional[FlatGroup]:
    ref = _FLAT.get(id(p))
    g = ref() if ref is not None else None
    if g is None or not any(q is p for q in g.members):  # dead group, or a recycled id()
        return None
    return g


def flatten_parameters(module: torch.nn.Module) -> FlatGroup:
    ps = list(module.parameters())
    g = _flat_of(ps[0]) if ps else None
    if g is not None and len(g.members) == len(ps) and all(a is b for a, b in zip(g.members, ps)):
        g.check_alias()
        return g
    return FlatGroup(module)


def _plan(params: List[torch.nn.Parameter]):
    """Split a parameter list into whole flat groups and loose tensors."""
    groups, loose, seen = [], [], set()
    ids = {id(p) for p in params}
    for p in params:
        g = _flat_of(p)
        if g is not None and all(id(q) in ids for q in g.members):
            if id(g) not in seen:
                seen.add(id(g))
                groups.append(g)
        else:
            loose.append(p)
    return groups, loose


class _Fused(torch.optim.Optimizer):
    def zero_grad(self, set_to_none: bool = False):
        groups, loose = _plan([p for g in self.param_groups for p in g["params"]])
        for fg in groups:
            fg.zero_grads()
        for p in loose:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def _state_for(self, key, like: torch.Tensor, names):
        st = self.state.setdefault(key, {})
        for n in names:
            if n not in st:
                st[n] = torch.zeros_like(like, dtype=torch.float32)
        return st


class RMSprop(_Fused):
    """torch.optim.RMSprop semantics (no momentum, not centered): sq = a*sq + (1-a)*g^2; p -= lr*g/(sqrt(sq)+eps).
    ``clip=(lo, hi)`` fuses the WGAN weight clamp into the same pass."""

    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False, clip=None):
        if weight_decay or momentum or centered:
            raise NotImplementedError("fused RMSprop covers the reference's configuration (no weight decay / momentum / centering)")
        super().__init__(params, dict(lr=lr, alpha=alpha, eps=eps, clip=clip))

    @torch.no_grad()
    def step(self, closure=None):
        for g in self.param_groups:
            groups, loose = _plan(g["params"])
            clip = g["clip"]
            lo, hi = clip if clip else (0.0, 0.0)
            for fg in groups:
                fg.rebind_grads()
                st = self._state_for(fg.members[0], fg.params, ["square_avg"])
                call("b200_rmsprop_step", ptr(fg.params), ptr(fg.grads), ptr(st["square_avg"]), fg.numel, g["lr"], g["alpha"], g["eps"], int(bool(clip)),
                     lo, hi, stream())
            for p in loose:
                if p.grad is None:
                    continue
                st = self._state_for(p, p.data, ["square_avg"])
                call("b200_rmsprop_step", ptr(p.data), ptr(p.grad.contiguous()), ptr(st["square_avg"]), p.numel(), g["lr"], g["alpha"], g["eps"],
                     int(bool(clip)), lo, hi, stream())
        _bump(self)


class Adam(_Fused):
    """torch.optim.Adam semantics (no amsgrad / weight decay), bias correction from a device-side step counter so
    a captured CUDA graph of the step replays correctly."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay or amsgrad:
            raise NotImplementedError("fused Adam covers the reference's configuration (no weight decay / amsgrad)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._step_dev: Optional[torch.Tensor] = None

    @torch.no_grad()
    def step(self, closure=None):
        first = self.param_groups[0]["params"][0]
        if self._step_dev is None:
            self._step_dev = torch.zeros(1, dtype=torch.int32, device=first.device)
        call("b200_step_inc", ptr(self._step_dev), stream())
        for g in self.param_groups:
            groups, loose = _plan(g["params"])
            b1, b2 = g["betas"]
            for fg in groups:
                fg.rebind_grads()
                st = self._state_for(fg.members[0], fg.params, ["exp_avg", "exp_avg_sq"])
                call("b200_adam_step", ptr(fg.params), ptr(fg.grads), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), fg.numel, g["lr"], b1, b2, g["eps"],
                     ptr(self._step_dev), stream())
            for p in loose:
                if p.grad is None:
                    continue
                st = self._state_for(p, p.data, ["exp_avg", "exp_avg_sq"])
                call("b200_adam_step", ptr(p.data), ptr(p.grad.contiguous()), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), g["lr"], b1, b2,
                     g["eps"], ptr(self._step_dev), stream())
        _bump(self)


def _bump(opt=None):
    """The kernels write through raw pointers, which autograd's version counters do not see; advance the global
    weight epoch so the engines refresh their packed bf16 weight copies on the next forward."""
    _lib.WEIGHT_EPOCH += 1


def clamp_(params, lo: float, hi: float):
    """``for p in params: p.data.clamp_(lo, hi)`` (wgan_l1.py:152-153) as one launch per flat group."""
    groups, loose = _plan(list(params))
    for fg in groups:
        call("b200_clamp", ptr(fg.params), fg.numel, float(lo), float(hi), stream())
    for p in loose:
        call("b200_clamp", ptr(p.data), p.numel(), float(lo), float(hi), stream())
    _bump()

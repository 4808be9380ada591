"""fp32 restatement of the reference's per-batch train-step recipes.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's experiment
scripts cannot be imported (they import the broken ``lib.models.loss``,
SURVEY.md D7), so the orchestration is restated here over ``oracle.nets`` with
these explicit, documented deviations:

* WGAN backward seeds are ``+1`` / ``-1`` (the reference uses an uninitialised
  ``torch.FloatTensor(1)``, wgan_l1.py:74-78 — SURVEY.md D10);
* dropout is an explicit keep-mask argument;
* the PatchGAN head is sized ``(R/16-3)**2``.

Every step function mutates the parameter dicts / optimizers in place (like the
reference) and returns a dict of the scalars and tensors the parity tests
compare: losses, ``inpainted``, and the parameter gradients *as they were when
the optimizer consumed them*.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import losses, nets

Params = Dict[str, torch.Tensor]


def leaves(p: Params):
    """Trainable leaves in state_dict order (== module.parameters() order)."""
    return [v for v in nets.trainable(p).values()]


def enable_grad(p: Params):
    for k, v in nets.trainable(p).items():
        v.requires_grad_(True)
    return p


def set_requires_grad(param_dicts, flag: bool):
    """util.set_requires_grad — util.py:19-22."""
    for p in param_dicts:
        for v in nets.trainable(p).values():
            v.requires_grad_(flag)


def make_optimizer(kind: str, *param_dicts: Params):
    """'rmsprop': wgan_l1.py:64-65 (lr 5e-5, alpha .99, eps 1e-8).
    'adam': minimaxgan_rmse.py:64-65 / experiment1…:122-123 (lr 2e-4, betas (.5,.999))."""
    ps = [v for p in param_dicts for v in leaves(p)]
    if kind == "rmsprop":
        return torch.optim.RMSprop(ps, lr=0.00005)
    if kind == "adam":
        return torch.optim.Adam(ps, lr=0.0002, betas=(0.5, 0.999))
    raise ValueError(kind)


def _scope(policy, name):
    """Name the forward pass about to run (TeacherForcedPolicy keys its stored tensors by '<scope>/<tag>')."""
    if policy is not None:
        policy.scope = name


def grads_of(p: Params) -> Params:
    return {k: (v.grad.detach().clone() if v.grad is not None else None) for k, v in nets.trainable(p).items()}


def _recon(kind: str, ground, inpainted):
    if kind == "l1":  # nn.L1Loss()(ground, inpainted) — wgan_l1.py:62,180
        return (ground - inpainted).abs().mean()
    if kind == "rmse":  # RMSELoss()(ground, inpainted) — minimaxgan_rmse.py:166, wgan_rmse.py:180
        return losses.rmse(ground, inpainted)
    raise ValueError(kind)


def prepare(ground, mask, ceil_mask: bool = True):
    """wgan_l1.py:113-117: mask = ceil(mask); masked = ground * (1 - mask)."""
    if ceil_mask:
        mask = torch.ceil(mask)
    return mask, ground * (1 - mask)


def composite(masked, g_out, mask):
    """wgan_l1.py:122: inpainted = masked + G(masked) * mask."""
    return masked + g_out * mask


def inference(G: Params, ground, mask, train_mode: bool = False):
    """visualize.py:69-76 / evaluate.py:127-136 (config 5): ceil -> mask out -> G -> composite."""
    with torch.no_grad():
        mask, masked = prepare(ground, mask)
        return composite(masked, nets.unet_forward(G, masked, train=train_mode), mask)


def calculate_metric(G: Params, batches, is_flip_mask: bool = False, train_mode: bool = True, eps: float = 1e-16):
    """The reconstruction part of evaluate.py:87-177 (``calculate_metric``; the file itself is a TabError, SURVEY.md D7, so this
    follows its statements): per batch ``mask = ceil(mask)`` [``1 - mask``] (:129-132), ``masked = input*(1-m)``,
    ``out = net(masked)*m + masked`` (:133-136), RMSELoss / L1Loss globally (:139-140) and LocalLoss(L1) / the intended
    LocalLoss(RMSELoss) = sqrt(masked-MSE micro-average + eps) (:142-143, loss.py:30-31,44-45); means over batches (:166-170).
    ``train_mode`` mirrors the reference never calling ``net.eval()``."""
    acc = dict(recon_rmse_global=0.0, recon_l1_global=0.0, recon_rmse_local=0.0, recon_l1_local=0.0)
    n = 0
    with torch.no_grad():
        for ground, mask, *_ in batches:
            m = torch.ceil(mask)
            if is_flip_mask:
                m = 1 - m
            masked = ground * (1 - m)
            out = nets.unet_forward(G, masked, train=train_mode) * m + masked
            acc["recon_rmse_global"] += float(losses.rmse(ground, out, eps))
            acc["recon_l1_global"] += float((ground - out).abs().mean())
            acc["recon_rmse_local"] += float(torch.sqrt(losses.local_loss(ground, out, m, "mse") + eps))
            acc["recon_l1_local"] += float(losses.local_loss(ground, out, m, "l1"))
            n += 1
    return {k: v / n for k, v in acc.items()}


def wgan_step(
    G: Params,
    D: Params,
    g_opt,
    d_opt,
    ground,
    mask,
    recon: str = "l1",
    update_g: bool = True,
    dropout_masks: Optional[dict] = None,
    clip: float = 0.01,
    policy=None,
    after_d_step=None,
    tv_weight: float = 0.0,
    vgg_features=None,
    feature_weights=(0.01, 0.01),
    local_recon: Optional[str] = None,
):
    """wgan_l1.py:110-186 (recon='l1') / wgan_rmse.py (recon='rmse').

    ``update_g`` stands for the schedule test at wgan_l1.py:157-163.  ``tv_weight`` / ``vgg_features`` (+ ``feature_weights``) /
    ``local_recon`` add the generator loss terms of wgan_perceptual_style_faceparsing.py:208-221 — ``tv_loss(inpainted, 1)``,
    ``perceptual_and_style_loss(inpainted, ground, 0.01, 0.01)`` (no gradient, D11) and the local reconstruction term — as that
    script intends them (it does not run: ``recon_loss`` :221 is undefined and ``LocalLoss(RMSELoss())`` :63 cannot be constructed;
    the usable ``LocalLoss`` bases are L1 / MSE).  Its face-parsing term needs a trained segmentation network and is not restated."""
    out = {}
    mask, masked = prepare(ground, mask)
    _scope(policy, "G")
    g_out = nets.unet_forward(G, masked, train=True, dropout_masks=dropout_masks, policy=policy)
    inpainted = composite(masked, g_out, mask)
    out["g_out"], out["inpainted"] = g_out.detach(), inpainted.detach()

    # ---- D step (:128-153)
    set_requires_grad([D], True)
    d_opt.zero_grad()
    _scope(policy, "Dreal")
    d_real = nets.patchgan_forward(D, ground, train=True, sigmoid=False, policy=policy)
    _scope(policy, "Dfake")
    d_fake = nets.patchgan_forward(D, inpainted.detach(), train=True, sigmoid=False, policy=policy)
    d_loss_real = d_real.mean().view(1)
    d_loss_real.backward(torch.ones(1))  # `one`  (D10)
    d_loss_fake = d_fake.mean().view(1)
    d_loss_fake.backward(-torch.ones(1))  # `mone` (D10)
    out["d_loss_real"], out["d_loss_fake"] = d_loss_real.item(), d_loss_fake.item()
    out["d_grads"] = grads_of(D)
    d_opt.step()
    with torch.no_grad():
        for v in leaves(D):  # every D parameter incl. BN gamma/beta and the Linear head (:152-153)
            v.clamp_(-clip, clip)
    if after_d_step is not None:  # test hook: e.g. continue the G step from an implementation's post-step D weights
        after_d_step(D)

    # ---- G step (:169-186)
    if update_g:
        set_requires_grad([D], False)
        g_opt.zero_grad()
        _scope(policy, "Dgen")
        d_gen = nets.patchgan_forward(D, inpainted, train=True, sigmoid=False, policy=policy).view(-1)
        g_adv = d_gen.mean().view(1)
        rec = _recon(recon, ground, inpainted)
        g_loss = g_adv + rec
        if local_recon is not None:  # :209
            loc = losses.local_loss(inpainted, ground, mask, local_recon)
            g_loss = g_loss + loc
            out["recon_local"] = loc.item()
        if tv_weight:  # :219
            tv = losses.tv_loss(inpainted, tv_weight)
            g_loss = g_loss + tv
            out["g_tv"] = tv.item()
        if vgg_features is not None:  # :216
            from . import vgg_losses

            with torch.no_grad():
                p, s = vgg_losses.perceptual_and_style_loss(vgg_features, inpainted, ground, feature_weights[0], feature_weights[1], bf16=policy is not None)
            g_loss = g_loss + p + s
            out["g_perceptual"], out["g_style"] = p.item(), s.item()
        g_loss.backward()
        out["g_adv"], out["recon"], out["g_loss"] = g_adv.item(), rec.item(), g_loss.item()
        out["g_grads"] = grads_of(G)
        g_opt.step()
    return out


def minimax_step(G: Params, D: Params, g_opt, d_opt, ground, mask, recon: str = "rmse", dropout_masks: Optional[dict] = None, policy=None,
                 after_d_step=None):
    """minimaxgan_rmse.py:110-173 (recon='rmse') / minimaxgan_l1.py (recon='l1')."""
    out = {}
    mask, masked = prepare(ground, mask)
    _scope(policy, "G")
    g_out = nets.unet_forward(G, masked, train=True, dropout_masks=dropout_masks, policy=policy)
    inpainted = composite(masked, g_out, mask)
    out["g_out"], out["inpainted"] = g_out.detach(), inpainted.detach()

    set_requires_grad([D], True)
    d_opt.zero_grad()
    _scope(policy, "Dreal")
    d_real = nets.patchgan_forward(D, ground, train=True, sigmoid=True, policy=policy).view(-1)
    d_loss_real = losses.bce(d_real, 1.0)
    d_loss_real.backward()
    _scope(policy, "Dfake")
    d_fake = nets.patchgan_forward(D, inpainted.detach(), train=True, sigmoid=True, policy=policy).view(-1)
    d_loss_fake = losses.bce(d_fake, 0.0)
    d_loss_fake.backward()
    out["d_loss_real"], out["d_loss_fake"] = d_loss_real.item(), d_loss_fake.item()
    out["d_grads"] = grads_of(D)
    d_opt.step()
    if after_d_step is not None:
        after_d_step(D)

    set_requires_grad([D], False)
    g_opt.zero_grad()
    _scope(policy, "Dgen")
    d_gen = nets.patchgan_forward(D, inpainted, train=True, sigmoid=True, policy=policy).view(-1)
    g_adv = losses.bce(d_gen, 1.0)
    rec = _recon(recon, ground, inpainted)
    g_loss = g_adv + rec
    g_loss.backward()
    out["g_adv"], out["recon"], out["g_loss"] = g_adv.item(), rec.item(), g_loss.item()
    out["g_grads"] = grads_of(G)
    g_opt.step()
    return out


def global_d_step(G: Params, D: Params, g_opt, d_opt, ground, mask, lambda1: float = 300.0, dropout_masks: Optional[dict] = None, policy=None):
    """experiment3_global_D.py:140-196: the global-discriminator-only sibling of ``global_local_step`` — G step first, no composite,
    mask not ceil'ed, BCE heads on a sigmoid D, ``g_loss = BCE(D(G(x)), 1) + lambda1 * RMSE``; then ONE D backward of
    ``BCE(D(ground), 1) + BCE(D(G(x).detach()), 0)``; Adam(2e-4, (0.5, 0.999)) for both (:121-122)."""
    out = {}
    masked = ground * (1 - mask)
    set_requires_grad([D], False)
    g_opt.zero_grad()
    _scope(policy, "G")
    inpainted = nets.unet_forward(G, masked, train=True, dropout_masks=dropout_masks, policy=policy)
    out["inpainted"] = inpainted.detach()
    _scope(policy, "DgGen")
    pg = nets.patchgan_forward(D, inpainted, train=True, sigmoid=True, policy=policy).view(-1)
    g_adv = losses.bce(pg, 1.0)
    rec = losses.rmse(ground, inpainted)
    g_loss = g_adv + lambda1 * rec
    g_loss.backward()
    out.update(g_adv=g_adv.item(), recon=rec.item(), g_loss=g_loss.item())
    out["g_grads"] = grads_of(G)
    g_opt.step()

    set_requires_grad([D], True)
    d_opt.zero_grad()
    _scope(policy, "DgReal")
    l_r = losses.bce(nets.patchgan_forward(D, ground, train=True, sigmoid=True, policy=policy).view(-1), 1.0)
    _scope(policy, "DgFake")
    l_f = losses.bce(nets.patchgan_forward(D, inpainted.detach(), train=True, sigmoid=True, policy=policy).view(-1), 0.0)
    d_loss = l_r + l_f
    d_loss.backward()
    out.update(d_loss_real=l_r.item(), d_loss_fake=l_f.item(), d_loss=d_loss.item())
    out["d_grads"] = grads_of(D)
    d_opt.step()
    return out


def global_local_step(
    G: Params,
    Dg: Params,
    Dl: Params,
    g_opt,
    d_opt,
    ground,
    mask,
    lambda1: float = 300.0,
    lambda2: float = 300.0,
    dropout_masks: Optional[dict] = None,
    policy=None,
):
    """experiment1_global_local_D.py:139-200.  G step first, no composite, mask not ceil'ed,
    LSGAN heads on sigmoid outputs, ``d_opt`` is ONE Adam over chain(Dl, Dg) (:123)."""
    out = {}
    masked = ground * (1 - mask)

    set_requires_grad([Dg, Dl], False)
    g_opt.zero_grad()
    _scope(policy, "G")
    inpainted = nets.unet_forward(G, masked, train=True, dropout_masks=dropout_masks, policy=policy)
    out["inpainted"] = inpainted.detach()
    _scope(policy, "DgGen")
    pg = nets.patchgan_forward(Dg, inpainted, train=True, sigmoid=True, policy=policy).view(-1)
    _scope(policy, "DlGen")
    pl = nets.patchgan_forward(Dl, mask * inpainted, train=True, sigmoid=True, policy=policy).view(-1)
    g_adv_g, g_adv_l = losses.lsgan(pg, 1.0), losses.lsgan(pl, 1.0)
    rec_g = losses.rmse(ground, inpainted)
    rec_l = losses.rmse(mask * ground, mask * inpainted)  # averages over ALL pixels (not LocalLoss)
    g_loss = g_adv_g + g_adv_l + lambda1 * rec_g + lambda2 * rec_l
    g_loss.backward()
    out.update(g_adv_global=g_adv_g.item(), g_adv_local=g_adv_l.item(), recon_global=rec_g.item(), recon_local=rec_l.item(), g_loss=g_loss.item())
    out["g_grads"] = grads_of(G)
    g_opt.step()

    set_requires_grad([Dg, Dl], True)
    d_opt.zero_grad()
    fake = inpainted.detach()
    _scope(policy, "DgReal")
    l_rg = losses.lsgan(nets.patchgan_forward(Dg, ground, train=True, sigmoid=True, policy=policy).view(-1), 1.0)
    _scope(policy, "DgFake")
    l_fg = losses.lsgan(nets.patchgan_forward(Dg, fake, train=True, sigmoid=True, policy=policy).view(-1), 0.0)
    _scope(policy, "DlReal")
    l_rl = losses.lsgan(nets.patchgan_forward(Dl, ground * mask, train=True, sigmoid=True, policy=policy).view(-1), 1.0)
    _scope(policy, "DlFake")
    l_fl = losses.lsgan(nets.patchgan_forward(Dl, fake * mask, train=True, sigmoid=True, policy=policy).view(-1), 0.0)
    d_loss = l_rl + l_fl + l_rg + l_fg
    d_loss.backward()
    out["d_loss"] = d_loss.item()
    out["dg_grads"], out["dl_grads"] = grads_of(Dg), grads_of(Dl)
    d_opt.step()
    return out

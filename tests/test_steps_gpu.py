"""Step-level parity (GPU): one full train step of every recipe through
``gan_inpainting_b200.trainer.InpaintingTrainer`` against ``oracle.steps`` on the same seeded inputs and weights:
losses, the composited image, every parameter gradient as the optimizer consumed it, and the post-step weights.
Also checks the committed golden scalars produced by the REFERENCE modules (batch 2).

Two tiers, as in tests/test_networks_gpu.py (DESIGN.md "Numerical parity"):
A. teacher-forced oracle step (the oracle continues from the CUDA path's stored activations, so masks and batch
   statistics coincide): losses, every gradient and the post-step weights to rel-L2 <= 1e-2 (north_star);
B. free-running fp32 oracle step: images and losses to 1e-2 / 2e-2, gradient cosine >= 0.9, post-step weights 1e-2.
"""
import pytest
import torch

from tests.test_networks_gpu import make_D, make_G
from tests.util import cos, rel, stored_from_disc_ctx, stored_from_gen_ctx, summary_close

pytestmark = pytest.mark.gpu

TOL = 1e-2


def _imports():
    from gan_inpainting_b200 import trainer
    from oracle import nets, steps, synth

    return trainer, nets, steps, synth


def close(a, b, tol=TOL, floor=1e-3):
    """|a - b| <= tol * max(|b|, floor); ``floor`` is the natural scale of a quantity that may cancel to ~0 (a WGAN critic mean)."""
    return abs(a - b) <= tol * max(abs(b), floor)


def check_grads(flat_group, ref_grads, tol=TOL, what=""):
    bad = []
    for name, p, g in zip(flat_group.names, flat_group.members, flat_group.grad_views):
        e = rel(g, ref_grads[name])
        if not e < tol:
            bad.append((name, e))
    assert not bad, (what, bad)


def check_cos(flat_group, ref_grads, lo=0.9, what=""):
    bad = []
    for name, g in zip(flat_group.names, flat_group.grad_views):
        if float(ref_grads[name].norm()) < 1e-12:  # e.g. the WGAN head bias: d(mean real - mean fake)/d bias == 0 exactly
            assert float(g.norm()) < 1e-6, name
            continue
        c = cos(g, ref_grads[name])
        if not c > lo:
            bad.append((name, c))
    assert not bad, (what, bad)


def forced_policy(tr, nets):
    """TeacherForcedPolicy over every forward pass of the last trainer step."""
    stored = {}
    for scope, c in tr.last["ctx"].items():
        if c is None:
            continue
        if scope == "G":
            stored.update(stored_from_gen_ctx(c, tr.G._cfg["widths"], scope))
        else:
            stored.update(stored_from_disc_ctx(c, scope))
    return nets.TeacherForcedPolicy(stored)


def adopt_params(net, checks):
    """after_d_step hook for the teacher-forced tier: first record the oracle's own post-step weights (compared with the
    CUDA path's afterwards), then continue the step from the CUDA path's weights, so that the remaining forward passes
    are compared layer by layer under identical parameters (sign descent flips a few near-zero-gradient elements)."""

    def hook(params):
        checks.update({k: v.detach().clone() for k, v in params.items()})
        sd = net.state_dict()
        with torch.no_grad():
            for k, v in params.items():
                if v.dtype.is_floating_point and "running" not in k:  # buffers keep following the oracle's own forwards
                    v.copy_(sd[k].cpu())

    return hook


def check_forced_layers(pol, n_expected):
    assert len(pol.errors) == n_expected, (len(pol.errors), n_expected)
    worst = max(pol.errors.items(), key=lambda kv: kv[1])
    assert worst[1] < 5e-3, worst


RMSPROP_STEP = 5e-4  # first RMSprop step: lr * g / (sqrt(0.01 g^2) + eps) = 10 * lr * sign(g), lr = 5e-5
ADAM_STEP = 2e-4     # first Adam step: lr * sign(g), lr = 2e-4


def check_params(net, ref_params, step, tol=TOL, what="", flips=0.02, running=True):
    """Post-step weights.  The first RMSprop / Adam step is sign descent (|update| == ``step`` for every element), which
    is discontinuous where a gradient element is ~0: the bound is tol * ||w_ref|| plus the norm of a fraction ``flips``
    of sign disagreements (each 2 * step) — 2 % teacher-forced, 15 % free-running (where gradient cosine is ~0.95).
    The optimizer kernels themselves are checked exactly against torch.optim in tests/test_kernels_gpu.py.
    ``running=False`` skips the BatchNorm running statistics: a snapshot taken right after the D step (``adopt_params``)
    has seen two train-mode forwards, the discriminator at the end of the step three (the G step's forward, wgan_l1.py:172);
    they are compared with the oracle's end-of-step buffers instead (``check_running``)."""
    bad = []
    for k, v in net.state_dict().items():
        if "running" in k and not running:
            continue
        if v.dtype.is_floating_point:
            r = ref_params[k].detach().float()
            err = float((v.detach().float().cpu() - r).norm())
            bound = tol * float(r.norm()) + 2 * step * (flips * r.numel()) ** 0.5
            if not err <= bound:
                bad.append((k, err, bound))
    assert not bad, (what, bad)


def check_running(net, ref_params, tol=TOL, what=""):
    bad = [(k, rel(v, ref_params[k])) for k, v in net.state_dict().items() if "running" in k and not rel(v, ref_params[k]) < tol]
    assert not bad, (what, bad)


@pytest.mark.parametrize("recon", ["l1", "rmse"])
def test_wgan_step(dev, recon):
    trainer, nets, steps, synth = _imports()
    R, B = 128, 8
    ground, mask = synth.make_batch(B, R, seed=21, soft_edge=True)
    gp, dp = nets.init_unet(0), nets.init_patchgan(1, res=R)
    G, D = make_G(dev, gp).train(), make_D(dev, dp, R, False).train()
    tr = trainer.InpaintingTrainer("wgan", G, D, recon=recon)
    tr.step(ground.to(dev), mask.to(dev), update_g=True)
    s = tr.read_scalars()
    cl = lambda p: steps.enable_grad({k: v.clone() for k, v in p.items()})
    assert torch.equal(tr.last["mask"].cpu(), torch.ceil(mask))  # bit-exact binarisation
    # ---- tier A: teacher-forced
    pol = forced_policy(tr, nets)
    G_, D_ = cl(gp), cl(dp)
    d_post = {}
    o = steps.wgan_step(G_, D_, steps.make_optimizer("rmsprop", G_), steps.make_optimizer("rmsprop", D_), ground, mask, recon=recon, policy=pol,
                        after_d_step=adopt_params(D, d_post))
    check_forced_layers(pol, 24 + 3 * 7)
    assert rel(tr.last["inpainted"], o["inpainted"]) < 1e-4
    for k in ("d_loss_real", "d_loss_fake", "g_adv", "recon", "g_loss"):
        assert close(s[k], o[k], 2e-3), (k, s[k], o[k])
    check_grads(tr.fD, o["d_grads"], what="D grads (teacher-forced)")
    check_grads(tr.fG, o["g_grads"], what="G grads (teacher-forced)")
    check_params(D, d_post, RMSPROP_STEP, what="D after RMSprop + clip", running=False)
    check_running(D, D_, what="D running statistics after the step")
    check_params(G, G_, RMSPROP_STEP, what="G after RMSprop")
    # ---- tier B: free-running fp32 reference
    G_, D_ = cl(gp), cl(dp)
    o = steps.wgan_step(G_, D_, steps.make_optimizer("rmsprop", G_), steps.make_optimizer("rmsprop", D_), ground, mask, recon=recon)
    assert rel(tr.last["inpainted"], o["inpainted"]) < TOL
    for k in ("d_loss_real", "d_loss_fake", "g_adv", "recon", "g_loss"):
        assert close(s[k], o[k], 2e-2, floor=0.1), (k, s[k], o[k])  # critic outputs are O(0.1-1); their batch mean cancels
    check_cos(tr.fD, o["d_grads"], what="D grads (free-running)")
    check_cos(tr.fG, o["g_grads"], what="G grads (free-running)")
    check_params(D, D_, RMSPROP_STEP, what="D after RMSprop + clip (free-running)", flips=0.15)
    check_params(G, G_, RMSPROP_STEP, what="G after RMSprop (free-running)", flips=0.15)
    # the clip touches every D parameter, BN gamma included (gamma = 1 -> 0.01)
    assert float(max(p.abs().max() for p in D.parameters())) <= 0.01 + 1e-7
    assert abs(float(D.state_dict()["model.3.weight"].mean()) - 0.01) < 1e-6


def test_wgan_step_with_config4_loss_terms(dev):
    """The generator loss of wgan_perceptual_style_faceparsing.py:208-221 without its face-parsing term: adversarial + RMSE + local
    (masked, micro-averaged MSE) + TV (with gradient) + VGG19 perceptual / style (numbers only) at 256x256 — teacher-forced."""
    trainer, nets, steps, synth = _imports()
    from gan_inpainting_b200.lib.models import loss, networks
    from oracle import vgg_losses

    R, B = 256, 4
    ground, mask = synth.make_batch(B, R, seed=31, soft_edge=True)
    gp, dp = nets.init_unet(0), nets.init_patchgan(1, res=R)
    G, D = make_G(dev, gp).train(), make_D(dev, dp, R, False).train()
    feats = vgg_losses.seeded_vgg_features(11)
    w = networks.VGG19Wrapper()
    w.vgg19.features.load_state_dict(feats.state_dict())
    loss.set_vgg(w.to(dev).eval())
    try:
        tr = trainer.InpaintingTrainer("wgan", G, D, recon="rmse", tv_weight=1.0, feature_weights=(0.01, 0.01), local_recon="mse")
        tr.step(ground.to(dev), mask.to(dev), update_g=True)
        s = tr.read_scalars()
    finally:
        loss.set_vgg(None)
    cl = lambda p: steps.enable_grad({k: v.clone() for k, v in p.items()})
    pol = forced_policy(tr, nets)
    G_, D_ = cl(gp), cl(dp)
    d_post = {}
    o = steps.wgan_step(G_, D_, steps.make_optimizer("rmsprop", G_), steps.make_optimizer("rmsprop", D_), ground, mask, recon="rmse", policy=pol,
                        after_d_step=adopt_params(D, d_post), tv_weight=1.0, vgg_features=feats, feature_weights=(0.01, 0.01), local_recon="mse")
    assert rel(tr.last["inpainted"], o["inpainted"]) < 1e-4
    for k in ("g_adv", "recon", "recon_local", "g_tv", "g_loss"):
        assert close(s[k], o[k], 2e-3), (k, s[k], o[k])
    for k in ("g_perceptual", "g_style"):  # bf16 storage through 13 convolutions; the oracle emulates it under a policy
        assert s[k] > 0 and close(s[k], o[k], 2e-2, floor=0.0), (k, s[k], o[k])
    assert abs(s["g_loss"] - (s["g_adv"] + s["recon"] + s["recon_local"] + s["g_tv"] + s["g_perceptual"] + s["g_style"])) < 1e-5 * max(1.0, abs(s["g_loss"]))
    check_grads(tr.fG, o["g_grads"], what="G grads with TV + local terms (teacher-forced)")
    check_params(G, G_, RMSPROP_STEP, what="G after RMSprop")
    # the extra terms really reach the gradient: the same step without them gives different G gradients
    G2, D2 = make_G(dev, gp).train(), make_D(dev, dp, R, False).train()
    tr2 = trainer.InpaintingTrainer("wgan", G2, D2, recon="rmse")
    tr2.step(ground.to(dev), mask.to(dev), update_g=True)
    assert rel(tr2.fG.grads, tr.fG.grads) > 1e-3


def test_wgan_d_only_step(dev):
    """4 of 5 steady-state batches skip the G update (wgan_l1.py:157-163)."""
    trainer, nets, steps, synth = _imports()
    R, B = 128, 4
    ground, mask = synth.make_batch(B, R, seed=22)
    gp, dp = nets.init_unet(0), nets.init_patchgan(1, res=R)
    G, D = make_G(dev, gp).train(), make_D(dev, dp, R, False).train()
    tr = trainer.InpaintingTrainer("wgan", G, D)
    tr.step(ground.to(dev), mask.to(dev), update_g=False)
    G_, D_ = steps.enable_grad({k: v.clone() for k, v in gp.items()}), steps.enable_grad({k: v.clone() for k, v in dp.items()})
    steps.wgan_step(G_, D_, steps.make_optimizer("rmsprop", G_), steps.make_optimizer("rmsprop", D_), ground, mask, update_g=False)
    check_params(D, D_, RMSPROP_STEP, what="D after D-only step", flips=0.15)
    for k, v in G.state_dict().items():  # conv / norm weights untouched; BN running stats did move (train-mode forward)
        if "running" not in k and v.dtype.is_floating_point:
            assert torch.equal(v.cpu(), gp[k]), k
    check_params(G, G_, RMSPROP_STEP, what="G buffers after D-only step", tol=1.5e-2)


@pytest.mark.parametrize("recon", ["rmse", "l1"])
def test_minimax_step(dev, recon):
    trainer, nets, steps, synth = _imports()
    R, B = 128, 8
    ground, mask = synth.make_batch(B, R, seed=23, soft_edge=True)
    gp, dp = nets.init_unet(0), nets.init_patchgan(1, res=R)
    G, D = make_G(dev, gp).train(), make_D(dev, dp, R, True).train()
    tr = trainer.InpaintingTrainer("minimax", G, D, recon=recon)
    tr.step(ground.to(dev), mask.to(dev))
    s = tr.read_scalars()
    cl = lambda p: steps.enable_grad({k: v.clone() for k, v in p.items()})
    # ---- tier A: teacher-forced
    pol = forced_policy(tr, nets)
    G_, D_ = cl(gp), cl(dp)
    d_post = {}
    o = steps.minimax_step(G_, D_, steps.make_optimizer("adam", G_), steps.make_optimizer("adam", D_), ground, mask, recon=recon, policy=pol,
                           after_d_step=adopt_params(D, d_post))
    check_forced_layers(pol, 24 + 3 * 7)
    assert rel(tr.last["inpainted"], o["inpainted"]) < 1e-4
    for k in ("d_loss_real", "d_loss_fake", "g_adv", "recon", "g_loss"):
        assert close(s[k], o[k], 2e-3), (k, s[k], o[k])
    check_grads(tr.fD, o["d_grads"], what="D grads (teacher-forced)")
    check_grads(tr.fG, o["g_grads"], what="G grads (teacher-forced)")
    check_params(D, d_post, ADAM_STEP, what="D after Adam", running=False)
    check_running(D, D_, what="D running statistics after the step")
    check_params(G, G_, ADAM_STEP, what="G after Adam")
    # ---- tier B: free-running fp32 reference
    G_, D_ = cl(gp), cl(dp)
    o = steps.minimax_step(G_, D_, steps.make_optimizer("adam", G_), steps.make_optimizer("adam", D_), ground, mask, recon=recon)
    assert rel(tr.last["inpainted"], o["inpainted"]) < TOL
    for k in ("d_loss_real", "d_loss_fake", "g_adv", "recon", "g_loss"):
        assert close(s[k], o[k], 2e-2), (k, s[k], o[k])
    check_cos(tr.fD, o["d_grads"], what="D grads (free-running)")
    check_cos(tr.fG, o["g_grads"], what="G grads (free-running)")
    check_params(D, D_, ADAM_STEP, what="D after Adam (free-running)", flips=0.15)
    check_params(G, G_, ADAM_STEP, what="G after Adam (free-running)", flips=0.15)


def test_global_local_step(dev):
    """config 1 (experiment1_global_local_D): 128x128, batch 8."""
    trainer, nets, steps, synth = _imports()
    R, B = 128, 8
    ground, mask = synth.make_batch(B, R, seed=24, soft_edge=True)
    gp, dp, dlp = nets.init_unet(0), nets.init_patchgan(1, res=R), nets.init_patchgan(2, res=R)
    G, Dg, Dl = make_G(dev, gp).train(), make_D(dev, dp, R, True).train(), make_D(dev, dlp, R, True).train()
    tr = trainer.InpaintingTrainer("global_local", G, Dg, Dl)
    tr.step(ground.to(dev), mask.to(dev))
    s = tr.read_scalars()
    cl = lambda p: steps.enable_grad({k: v.clone() for k, v in p.items()})
    run = lambda G_, Dg_, Dl_, pol: steps.global_local_step(G_, Dg_, Dl_, steps.make_optimizer("adam", G_), steps.make_optimizer("adam", Dl_, Dg_), ground,
                                                           mask, policy=pol)
    # ---- tier A: teacher-forced
    pol = forced_policy(tr, nets)
    G_, Dg_, Dl_ = cl(gp), cl(dp), cl(dlp)
    o = run(G_, Dg_, Dl_, pol)
    check_forced_layers(pol, 24 + 6 * 7)
    assert rel(tr.last["inpainted"], o["inpainted"]) < 1e-4
    assert close(s["g_adv"], o["g_adv_global"], 2e-3) and close(s["g_adv_local"], o["g_adv_local"], 2e-3)
    assert close(s["recon"], o["recon_global"], 2e-3) and close(s["recon_local"], o["recon_local"], 2e-3)
    assert close(s["g_loss"], o["g_loss"], 2e-3) and close(s["d_loss"], o["d_loss"], 2e-3)
    check_grads(tr.fG, o["g_grads"], what="G grads (teacher-forced)")
    check_grads(tr.fD, o["dg_grads"], what="Dg grads (teacher-forced)")
    check_grads(tr.fDl, o["dl_grads"], what="Dl grads (teacher-forced)")
    check_params(G, G_, ADAM_STEP, what="G after Adam")
    check_params(Dg, Dg_, ADAM_STEP, what="Dg after Adam")
    check_params(Dl, Dl_, ADAM_STEP, what="Dl after Adam")
    # ---- tier B: free-running fp32 reference
    G_, Dg_, Dl_ = cl(gp), cl(dp), cl(dlp)
    o = run(G_, Dg_, Dl_, None)
    assert rel(tr.last["inpainted"], o["inpainted"]) < 1.5e-2
    assert close(s["g_adv"], o["g_adv_global"], 2e-2) and close(s["g_adv_local"], o["g_adv_local"], 2e-2)
    assert close(s["recon"], o["recon_global"], 2e-2) and close(s["recon_local"], o["recon_local"], 2e-2)
    assert close(s["g_loss"], o["g_loss"], 2e-2) and close(s["d_loss"], o["d_loss"], 2e-2)
    check_cos(tr.fG, o["g_grads"], what="G grads (free-running)")
    check_cos(tr.fD, o["dg_grads"], what="Dg grads (free-running)")
    check_cos(tr.fDl, o["dl_grads"], what="Dl grads (free-running)")
    check_params(G, G_, ADAM_STEP, what="G after Adam (free-running)", flips=0.15)


def test_global_d_step(dev):
    """experiment3_global_D.py:140-196 (the global-discriminator-only sibling of config 1): 128x128, batch 8."""
    trainer, nets, steps, synth = _imports()
    R, B = 128, 8
    ground, mask = synth.make_batch(B, R, seed=26, soft_edge=True)
    gp, dp = nets.init_unet(0), nets.init_patchgan(1, res=R)
    G, D = make_G(dev, gp).train(), make_D(dev, dp, R, True).train()
    tr = trainer.InpaintingTrainer("global_d", G, D)
    tr.step(ground.to(dev), mask.to(dev))
    s = tr.read_scalars()
    cl = lambda p: steps.enable_grad({k: v.clone() for k, v in p.items()})
    run = lambda G_, D_, pol: steps.global_d_step(G_, D_, steps.make_optimizer("adam", G_), steps.make_optimizer("adam", D_), ground, mask, policy=pol)
    keys = ("g_adv", "recon", "g_loss", "d_loss_real", "d_loss_fake", "d_loss")
    # ---- tier A: teacher-forced
    pol = forced_policy(tr, nets)
    G_, D_ = cl(gp), cl(dp)
    o = run(G_, D_, pol)
    check_forced_layers(pol, 24 + 3 * 7)
    assert rel(tr.last["inpainted"], o["inpainted"]) < 1e-4
    for k in keys:
        assert close(s[k], o[k], 2e-3), (k, s[k], o[k])
    assert abs(s["g_loss"] - (s["g_adv"] + 300.0 * s["recon"])) < 1e-4 * abs(s["g_loss"])
    check_grads(tr.fG, o["g_grads"], what="G grads (teacher-forced)")
    check_grads(tr.fD, o["d_grads"], what="D grads (teacher-forced)")
    check_params(G, G_, ADAM_STEP, what="G after Adam")
    check_params(D, D_, ADAM_STEP, what="D after Adam")
    # ---- tier B: free-running fp32 reference
    G_, D_ = cl(gp), cl(dp)
    o = run(G_, D_, None)
    assert rel(tr.last["inpainted"], o["inpainted"]) < 1.5e-2
    for k in keys:
        assert close(s[k], o[k], 2e-2), (k, s[k], o[k])
    check_cos(tr.fG, o["g_grads"], what="G grads (free-running)")
    check_cos(tr.fD, o["d_grads"], what="D grads (free-running)")
    with pytest.raises(ValueError):
        trainer.InpaintingTrainer("global_d", G, D, net_D_local=D)


def test_steps_vs_reference_golden_scalars(dev, golden):
    """Batch-2 scalars recorded from the REFERENCE modules (oracle/pin_against_reference.py)."""
    trainer, nets, _, _ = _imports()
    ground, mask = golden["ground"].to(dev), golden["mask"].to(dev)
    for key, recipe, recon, sig in (("wgan_l1", "wgan", "l1", False), ("wgan_rmse", "wgan", "rmse", False), ("minimax_rmse", "minimax", "rmse", True),
                                    ("minimax_l1", "minimax", "l1", True)):
        G = make_G(dev, nets.init_unet(golden["seed_g"])).train()
        D = make_D(dev, nets.init_patchgan(golden["seed_d"], res=128), 128, sig).train()
        tr = trainer.InpaintingTrainer(recipe, G, D, recon=recon)
        tr.step(ground, mask)
        s, g = tr.read_scalars(), golden[key]
        assert rel(tr.last["inpainted"], g["inpainted"]) < TOL, key
        for k in ("d_loss_real", "d_loss_fake", "g_adv", "recon"):
            assert close(s[k], g[k], 3e-2, floor=0.1), (key, k, s[k], g[k])


def test_global_d_step_vs_reference_golden_scalars(dev):
    """The ``global_d`` recipe against scalars recorded from the REFERENCE modules (oracle/pin_global_d_against_reference.py)."""
    import os
    trainer, nets, _, _ = _imports()
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_global_d.pt"))
    G = make_G(dev, nets.init_unet(g["seed_g"])).train()
    D = make_D(dev, nets.init_patchgan(g["seed_d"], res=g["res"]), g["res"], True).train()
    tr = trainer.InpaintingTrainer("global_d", G, D)
    tr.step(g["ground"].to(dev), g["mask"].to(dev))
    s = tr.read_scalars()
    assert rel(tr.last["inpainted"], g["inpainted"]) < 1.5e-2
    for k in ("g_adv", "recon", "g_loss", "d_loss_real", "d_loss_fake", "d_loss"):
        assert close(s[k], g[k], 3e-2, floor=0.1), (k, s[k], g[k])


def test_module_api_step_equals_fused_step(dev):
    """The reference-style loop (nn.Module calls + autograd + optimizer.step(), wgan_l1.py:110-186) and the fused
    trainer run the same kernels: identical losses and weights."""
    from gan_inpainting_b200 import optim as bo
    from gan_inpainting_b200.lib.models import util
    trainer, nets, steps, synth = _imports()
    R, B = 128, 4
    ground, mask = synth.make_batch(B, R, seed=25)
    ground, mask = ground.to(dev), mask.to(dev)
    gp, dp = nets.init_unet(0), nets.init_patchgan(1, res=R)
    G, D = make_G(dev, gp).train(), make_D(dev, dp, R, False).train()
    g_opt, d_opt = bo.RMSprop(G.parameters(), lr=0.00005), bo.RMSprop(D.parameters(), lr=0.00005)
    one = torch.ones(1, device=dev)
    mask_c = torch.ceil(mask)
    masked = ground * (1 - mask_c)
    inpainted = masked + G(masked) * mask_c
    util.set_requires_grad([D], True)
    d_opt.zero_grad()
    pr, pf = D(ground), D(inpainted.detach())
    pr.mean().view(1).backward(one)
    pf.mean().view(1).backward(-one)
    d_opt.step()
    bo.clamp_(D.parameters(), -0.01, 0.01)
    util.set_requires_grad([D], False)
    g_opt.zero_grad()
    g_loss = D(inpainted).view(-1).mean().view(1) + torch.nn.L1Loss()(ground, inpainted)
    g_loss.backward()
    g_opt.step()

    G2, D2 = make_G(dev, gp).train(), make_D(dev, dp, R, False).train()
    tr = trainer.InpaintingTrainer("wgan", G2, D2, recon="l1")
    tr.step(ground, mask)
    s = tr.read_scalars()
    assert close(float(g_loss), s["g_loss"], 1e-4)
    for (k, a), b in zip(D.state_dict().items(), D2.state_dict().values()):
        if a.dtype.is_floating_point:
            assert rel(a, b) < 1e-4, k
    for (k, a), b in zip(G.state_dict().items(), G2.state_dict().values()):
        if a.dtype.is_floating_point:
            assert rel(a, b) < 1e-3, k


@pytest.mark.parametrize("R", [256, 512])
def test_inference_composite(dev, R):
    """config 5: ceil -> mask out -> G (eval) -> composite (visualize.py:69-76), at 256x256 and at config 5's 512x512."""
    trainer, nets, steps, synth = _imports()
    ground, mask = synth.make_batch(2, R, seed=26, soft_edge=True)
    gp = nets.init_unet(0)
    G = make_G(dev, gp).eval()
    D = make_D(dev, nets.init_patchgan(1, res=R), R, False)
    tr = trainer.InpaintingTrainer("wgan", G, D)
    out = tr.inpaint(ground.to(dev), mask.to(dev))
    ref = steps.inference({k: v.clone() for k, v in gp.items()}, ground, mask)
    assert rel(out, ref) < TOL
    keep = torch.ceil(mask) == 0
    assert torch.equal(out.cpu()[keep], ground[keep])  # pixels outside the mask pass through bit-exactly


def test_begin_call_surface(dev, tmp_path):
    """experiment_list.<recipe>.begin(state, loaders) (train.py:183-186 / wgan_l1.py:15): state keys of train.py:22-36, loaders yield
    (ground, mask, segment); history dicts, the -7777 sentinel and the epochN_G.pt checkpoint follow the reference."""
    from gan_inpainting_b200.experiment_list import minimaxgan_rmse, wgan_l1
    _, nets, _, synth = _imports()
    batches = []
    for i in range(3):
        g, m = synth.make_batch(2, 128, seed=40 + i)
        batches.append((g, m, torch.zeros(2, dtype=torch.long)))
    state = dict(experiments=["wgan_l1"], numepoch=1, batchsize=2, generator="unet", discriminator="patchgan", imagedim=128, saveevery=1, updatediscevery=3,
                 evalevery=1, debug="false", experiment_dir=str(tmp_path / "model"), log_dir=str(tmp_path / "log"))
    out = wgan_l1.begin(state, {"train": batches, "test": batches, "extra": batches})
    h = out["history"]
    assert len(h) == 2 and out["d_iter"] == 6 and out["g_iter"] == 0
    assert h[0]["g"]["total"] == -7777 and h[0]["d"]["update_count"] == 3  # 3 batches < 140: the WGAN schedule never updates G
    assert all(v == -7777 for v in h[0]["gradients"]["avg_g"].values()) and all(v > 0 for v in h[0]["gradients"]["avg_d"].values())
    assert not any("bias" in n for n in h[0]["gradients"]["avg_d"])
    sd = torch.load(tmp_path / "model" / "epoch1_G.pt")
    assert set(sd) == set(nets.init_unet(0))
    ev = out["eval_history"]  # evaluate.calculate_metric on both loaders after epoch 1 (wgan_l1.py:251-259)
    assert len(ev) == 1 and set(ev[0]) == {"train", "test"} and ev[0]["test"]["recon_l1_local"] > 0 and ev[0]["test"]["fid"] == -1
    out = minimaxgan_rmse.begin(dict(state, numepoch=0), {"train": batches, "test": batches, "extra": batches})
    h = out["history"]
    assert out["g_iter"] == 3 and h[0]["g"]["update_count"] == 3 and h[0]["g"]["total"] > 0
    assert all(v > 0 for v in h[0]["gradients"]["avg_g"].values())
    from gan_inpainting_b200.experiment_list import experiment1_global_local_D
    out = experiment1_global_local_D.begin(dict(state, numepoch=0), {"train": batches, "test": batches, "extra": batches})
    # lib.pytorch_ssim per batch (experiment1_global_local_D.py:209); an untrained generator's SSIM is ~0 and may have either sign
    assert -1.0 <= out["history"][0]["g"]["ssim"] <= 1.0 and out["history"][0]["g"]["ssim"] != 0.0
    from gan_inpainting_b200.experiment_list import experiment3_global_D
    out = experiment3_global_D.begin(dict(state, numepoch=0), {"train": batches, "test": batches, "extra": batches})
    g3 = out["history"][0]["g"]
    assert out["g_iter"] == 3 and g3["update_count"] == 3 and g3["total"] > 0 and -1.0 <= g3["ssim"] <= 1.0 and g3["ssim"] != 0.0  # :201
    # config 4 (wgan_perceptual_style_faceparsing.py) at 256x256: every batch updates G here so that the extra loss terms are logged
    from gan_inpainting_b200.experiment_list import _runner, wgan_perceptual_style_faceparsing as cfg4
    big = []
    for i in range(2):
        g, m = synth.make_batch(2, 256, seed=50 + i)
        big.append((g, m, torch.zeros(2, 256, 256, dtype=torch.long)))
    st4 = dict(state, numepoch=0, imagedim=256)
    with pytest.raises(NotImplementedError):
        cfg4.begin(dict(st4, segmentation_model=object()), {"train": big, "test": big, "extra": big})
    orig = _runner.wgan_g_schedule
    _runner.wgan_g_schedule = lambda update_g_every=5: _runner.every_batch
    try:
        out = cfg4.begin(dict(st4, segmentation_model=object(), ignore_segmentation_model=True), {"train": big, "test": big, "extra": big})
    finally:
        _runner.wgan_g_schedule = orig
    g4 = out["history"][0]["g"]
    assert g4["update_count"] == 2 and all(g4[k] > 0 for k in ("recon", "recon_local", "tv", "perceptual", "style"))


def test_device_prefetcher_feeds_identical_batches(dev):
    """data.DevicePrefetcher (the pinned host -> device copy of batch i+1 overlaps batch i): same tensors, same order, buffers are
    only recycled after their consumer; a train step fed through it equals a step fed with plain .to(device) tensors."""
    from gan_inpainting_b200.data import DevicePrefetcher
    trainer, nets, steps, synth = _imports()
    batches = [synth.make_batch(2, 128, seed=60 + i) + (i,) for i in range(5)]
    seen = []
    for g, m, tag in DevicePrefetcher(batches, dev):
        assert g.is_cuda and m.is_cuda
        torch.cuda._sleep(2_000_000)  # a consumer still busy with this buffer while the next copy is issued
        seen.append((g.clone(), m.clone(), tag))
    assert [t for _, _, t in seen] == list(range(5))
    for (g, m, _), (hg, hm, _) in zip(seen, batches):
        assert torch.equal(g.cpu(), hg) and torch.equal(m.cpu(), hm)


def test_device_prefetcher_with_gpu_load_transform(dev):
    """Decoded uint8 images through DevicePrefetcher(transform=GpuLoadTransform): the batches the train step sees equal the
    reference's host pipeline (pil_loader -> Resize -> ToTensor, lib/data/dataset.py:6-12 / train.py:69-72) bit for bit."""
    import numpy as np
    from gan_inpainting_b200.data import DevicePrefetcher, GpuLoadTransform
    from oracle import preprocess
    rng = np.random.default_rng(11)
    batches = []
    for i in range(3):
        img = rng.integers(0, 256, (2, 160, 160, 3), dtype=np.uint8)
        msk = np.zeros((2, 160, 160), np.uint8)
        msk[:, 30 + i:90, 40:120 - i] = 255
        batches.append((torch.from_numpy(img), torch.from_numpy(msk), i))
    for (g, m, tag), (hg, hm, _) in zip(DevicePrefetcher(batches, dev, transform=GpuLoadTransform(128, dev)), batches):
        assert g.shape == (2, 1, 128, 128) and g.dtype == torch.float32 and m.shape == (2, 1, 128, 128)
        ref_g = np.stack([preprocess.load_transform(x, 128) for x in hg.numpy()])
        ref_m = np.stack([preprocess.load_transform(x, 128) for x in hm.numpy()])
        assert torch.equal(g.cpu(), torch.from_numpy(ref_g)) and torch.equal(m.cpu(), torch.from_numpy(ref_m))
        assert float(m.max()) == 1.0 and 0.0 < float(m[m > 0].min()) < 1.0  # resized masks have soft edges: ceil() binarises them (wgan_l1.py:114)


@pytest.mark.parametrize("recipe", ["wgan", "minimax"])
def test_graph_replayed_step_equals_eager_step(dev, recipe):
    """InpaintingTrainer(graph=True): the first batch runs eagerly, the step is captured into a CUDA graph, later batches replay
    it.  The kernels are deterministic, so with dropout off four steps (one eager + three replays, fresh inputs copied into the
    static buffers, WGAN: alternating D-only / full steps = two graphs) leave bit-identical weights, statistics and scalars."""
    trainer, nets, steps, synth = _imports()
    R, B = 128, 4
    gp, dp = nets.init_unet(0), nets.init_patchgan(1, res=R)
    sig = recipe != "wgan"
    batches = [tuple(t.to(dev) for t in synth.make_batch(B, R, seed=70 + i)) for i in range(6)]
    sched = [True, True, False, True, False, True] if recipe == "wgan" else [True] * 6
    nets_out, scal, launches = [], [], []
    from gan_inpainting_b200 import _lib
    for graph in (False, True):
        G, D = make_G(dev, gp).train(), make_D(dev, dp, R, sig).train()
        tr = trainer.InpaintingTrainer(recipe, G, D, recon="l1", graph=graph)
        per_step = []
        for (g, m), upd in zip(batches, sched):
            n0 = _lib.value("b200_launch_count")
            tr.step(g, m, update_g=upd)
            launches.append(_lib.value("b200_launch_count") - n0)
            per_step.append(tr.scalars.clone())
        torch.cuda.synchronize()
        nets_out.append((G.state_dict(), D.state_dict()))
        scal.append(per_step)
        if graph:
            assert len(tr._graphs) == (2 if recipe == "wgan" else 1)
            assert torch.equal(tr.last["inpainted"], last_inp)
        else:
            last_inp = tr.last["inpainted"].clone()
    n = len(sched)
    # a replayed graph accounts for the kernels it re-launches (bench.py's gpu_launches): the eager count +1 (seed counter) and
    # +-2 for the weight re-packs that depend on what ran before
    assert all(abs(a - b) <= 4 for a, b in zip(launches[:n], launches[n:])), launches
    for a, b in zip(scal[0], scal[1]):
        assert torch.equal(a, b)
    for sd_a, sd_b in zip(nets_out[0], nets_out[1]):
        for k in sd_a:
            assert torch.equal(sd_a[k], sd_b[k]), k


def test_graph_replay_draws_fresh_dropout_masks(dev):
    """With dropout on (the reference default) every replay must use a new mask: the device seed counter the
    graph increments.  Same weights (lr = 0 is not available, so compare the dropout layer's zero pattern) on the same batch."""
    trainer, nets, steps, synth = _imports()
    R, B = 128, 2
    G = make_G(dev, nets.init_unet(0), dropout=True).train()
    D = make_D(dev, nets.init_patchgan(1, res=R), R, False).train()
    tr = trainer.InpaintingTrainer("wgan", G, D, graph=True)
    g, m = (t.to(dev) for t in synth.make_batch(B, R, seed=80))
    w = G._cfg["widths"]
    pats = []
    for i in range(4):
        tr.step(g, m)
        cat = tr.last["ctx"]["G"].cat[5]  # up half = relu(dropout(bn(yu6))): zero where dropped (or negative)
        pats.append((cat[..., w[4]:] == 0).clone())
    assert len(tr._graphs) == 1
    for i in range(1, 4):
        for j in range(i):
            differ = (pats[i] != pats[j]).float().mean().item()
            assert differ > 0.2, (i, j, differ)


@pytest.mark.parametrize("flip", [False, True])
def test_calculate_metric_matches_oracle(dev, flip):
    """lib.models.evaluate.calculate_metric (the train-time evaluation pass, evaluate.py:87-177 / wgan_l1.py:251-259): global and
    masked L1 / RMSE over a loader of (ground, mask, segment) host batches, against the oracle's restatement."""
    from gan_inpainting_b200.lib.models import evaluate
    trainer, nets, steps, synth = _imports()
    gp = nets.init_unet(0)
    G = make_G(dev, gp).train()
    batches = [synth.make_batch(3, 128, seed=90 + i, soft_edge=(i == 1)) + (None,) for i in range(3)]
    got = evaluate.calculate_metric(dev, batches, G, fid_stats=(-1, -1), mode="test", inception_model=None, epoch=7, is_flip_mask=flip)
    ref = steps.calculate_metric({k: v.clone() for k, v in gp.items()}, batches, is_flip_mask=flip, train_mode=True)
    for k, v in ref.items():
        assert close(got[k], v, TOL), (k, got[k], v)
    assert got["fid"] == -1 and got["epoch"] == 7
    with pytest.raises(ValueError):
        evaluate.calculate_metric(dev, [], G)


def test_two_rank_data_parallel_graph_step(dev):
    """N > 1 on real GPUs (skipped on 1-GPU boxes; the CPU suite covers the bucketing with gloo): tools/ddp_check.py under torchrun —
    replicas stay identical, the CUDA-graph replay (NCCL all-reduces captured) equals the eager step, the all-reduced gradient is the
    mean of the per-shard gradients."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(root, "tools", "ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DDP_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    # ... and against the oracle: N reference replicas at batch b with averaged gradients
    cmd[-1] = os.path.join(root, "tests", "ddp_oracle_check.py")
    cmd[cmd.index("29611")] = "29612"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500)
    assert r.returncode == 0 and "DDP_ORACLE_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

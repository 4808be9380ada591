"""Pin the oracle against the reference itself and write tests/golden/*.pt.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs ONLY in the build
container, where the reference is mounted read-only at /root/reference:

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.pin_against_reference

It imports the reference's own ``lib/models/networks.py`` (imports cleanly) and
exec's ``lib/models/loss.py`` with the two undefined names it needs injected
(SURVEY.md §8c), loads oracle-initialised weights into the reference modules,
and

1. asserts the functional restatement in ``oracle.nets`` / ``oracle.losses`` /
   ``oracle.steps`` reproduces the reference modules' outputs and gradients
   (fp32, CPU) to ~1e-6, and
2. stores the *reference's* outputs as small fixtures under ``tests/golden/``
   so the same check (and the CUDA parity tests) can run on the GPU box, where
   /root/reference does not exist.

Nothing from the reference's sources is copied: only its numerical outputs.
"""
from __future__ import annotations

import os
import sys

import torch
from torch import nn

from . import losses, nets, steps, synth

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from lib.models import networks as ref_networks  # noqa

    class _VggStub(nn.Module):
        def forward(self, x):
            return x

    ns = {"VGG19Wrapper": _VggStub, "device": torch.device("cpu"), "__name__": "ref_loss"}
    with open(os.path.join(REF, "lib/models/loss.py")) as f:
        exec(compile(f.read(), "ref_loss.py", "exec"), ns)
    sys.path.pop(0)
    return ref_networks, ns


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def summarize(t: torch.Tensor):
    """Small, size-independent fingerprint of a big tensor."""
    f = t.detach().float().flatten()
    idx = torch.linspace(0, f.numel() - 1, steps=min(f.numel(), 64)).long()
    return dict(shape=list(t.shape), norm=float(f.norm()), mean=float(f.mean()), absmean=float(f.abs().mean()), sample=f[idx].clone())


def ref_generator(ref_networks, params):
    net = ref_networks.UnetGenerator(1, 1, 7, ngf=64, norm_layer=ref_networks.get_norm_layer("batch"), use_dropout=False)
    net.load_state_dict(params, strict=True)
    return net


def ref_discriminator(ref_networks, params, res, sigmoid):
    net = ref_networks.PatchGANDiscriminator(sigmoid=sigmoid)
    nf = nets.patchgan_head_features(res)
    if nf != 25:  # the head only fits 128x128
        net.model[13] = nn.Linear(nf, 1)
    net.load_state_dict(params, strict=True)
    return net


def clone(p):
    return {k: v.clone() for k, v in p.items()}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ref_networks, ref_loss = load_reference()
    os.makedirs(OUT, exist_ok=True)
    R, B = 128, 2
    ground, mask = synth.make_batch(B, R, seed=1234, soft_edge=True)

    # ---------------- generator forward / backward --------------------------------
    gp = nets.init_unet(seed=0)
    dp = nets.init_patchgan(seed=1, res=R)
    net_g = ref_generator(ref_networks, clone(gp))
    assert list(net_g.state_dict().keys()) == list(gp.keys()), "state_dict key order differs"
    gold = {"res": R, "batch": B, "seed_inputs": 1234, "seed_g": 0, "seed_d": 1}

    mask_c, masked = steps.prepare(ground, mask)
    gold["mask_ceil_sum"] = float(mask_c.sum())
    gold["masked_pixel_count"] = losses.masked_pixel_count(mask)

    # hooks on every conv / convT of the reference generator
    ref_taps = {}
    convs = [m for m in net_g.modules() if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d))]
    # module order inside-out is d1, d2..d7, u7..u1 in *registration* order: outermost block
    # registers its downconv first, then the submodule, then its upconv
    hooks = []
    names = [f"d{i}.conv" for i in range(1, 8)] + [f"u{i}.conv" for i in range(7, 0, -1)]
    for m, n in zip(convs, names):
        hooks.append(m.register_forward_hook(lambda mod, i, o, n=n: ref_taps.__setitem__(n, o.detach().clone())))
    net_g.train()
    ref_out = net_g(masked.clone())
    for h in hooks:
        h.remove()
    loss_ref = (ref_out * torch.linspace(-1, 1, ref_out.numel()).view_as(ref_out)).sum()
    loss_ref.backward()

    op = steps.enable_grad(clone(gp))
    otaps = {}
    o_out = nets.unet_forward(op, masked.clone(), train=True, taps=otaps)
    (o_out * torch.linspace(-1, 1, o_out.numel()).view_as(o_out)).sum().backward()
    e = rel(o_out.detach(), ref_out.detach())
    print(f"[pin] unet train fwd rel-L2 vs reference: {e:.2e}")
    assert e < 1e-5
    for n in names:
        e = rel(otaps[n].detach(), ref_taps[n])
        assert e < 1e-5, (n, e)
    ref_named = dict(net_g.named_parameters())
    worst = 0.0
    for k, v in nets.trainable(op).items():
        worst = max(worst, rel(v.grad, ref_named[k].grad))
    print(f"[pin] unet grads worst rel-L2 vs reference: {worst:.2e}")
    assert worst < 1e-4
    for k in gp:
        if "running" in k:
            assert rel(op[k], net_g.state_dict()[k]) < 1e-5, k
    gold["unet_train_out"] = ref_out.detach().clone()
    gold["unet_taps"] = {n: summarize(t) for n, t in ref_taps.items()}
    gold["unet_grads"] = {k: summarize(p.grad) for k, p in ref_named.items()}
    gold["unet_running"] = {k: v.clone() for k, v in net_g.state_dict().items() if "running" in k and v.numel() <= 64}

    net_g.eval()
    with torch.no_grad():
        ref_eval = net_g(masked.clone())
        o_eval = nets.unet_forward(op, masked.clone(), train=False)
    assert rel(o_eval, ref_eval) < 1e-5
    gold["unet_eval_out"] = ref_eval.clone()

    # D9 witness: the skip half of the first cat is leaky_relu(d1 out)
    # ---------------- discriminator -------------------------------------------------
    for sig in (True, False):
        net_d = ref_discriminator(ref_networks, clone(dp), R, sig)
        assert list(net_d.state_dict().keys()) == list(dp.keys())
        net_d.train()
        x = ground.clone().requires_grad_(True)
        r = net_d(x)
        (r * torch.tensor([[1.0], [-2.0]])).sum().backward()
        od = steps.enable_grad(clone(dp))
        xo = ground.clone().requires_grad_(True)
        o = nets.patchgan_forward(od, xo, train=True, sigmoid=sig)
        (o * torch.tensor([[1.0], [-2.0]])).sum().backward()
        assert rel(o.detach(), r.detach()) < 1e-5
        assert rel(xo.grad, x.grad) < 1e-4
        rn = dict(net_d.named_parameters())
        for k, v in nets.trainable(od).items():
            assert rel(v.grad, rn[k].grad) < 1e-4, k
        tag = "sig" if sig else "lin"
        gold[f"patchgan_{tag}_out"] = r.detach().clone()
        gold[f"patchgan_{tag}_dx"] = summarize(x.grad)
        gold[f"patchgan_{tag}_grads"] = {k: summarize(p.grad) for k, p in rn.items()}
    print("[pin] patchgan fwd/bwd (sigmoid and linear heads) match the reference")

    # ---------------- losses (the reference's own classes via exec) -----------------
    yhat = torch.rand(B, 1, R, R)
    ref_rmse = ref_loss["RMSELoss"]()(yhat, ground)
    assert abs(float(ref_rmse) - float(losses.rmse(yhat, ground))) < 1e-7
    ref_ll1 = ref_loss["LocalLoss"](nn.L1Loss)(yhat, ground, mask)
    ref_lmse = ref_loss["LocalLoss"](nn.MSELoss)(yhat, ground, mask)
    assert abs(float(ref_ll1) - float(losses.local_loss(yhat, ground, mask, "l1"))) < 1e-6
    assert abs(float(ref_lmse) - float(losses.local_loss(yhat, ground, mask, "mse"))) < 1e-6
    gold["loss_yhat_seed"] = "torch.manual_seed(0) stream; stored"
    gold["loss_yhat"] = yhat
    gold["rmse"], gold["local_l1"], gold["local_mse"] = float(ref_rmse), float(ref_ll1), float(ref_lmse)
    print("[pin] RMSELoss / LocalLoss match the reference classes")

    # ---------------- step recipes driven over the REFERENCE modules ----------------
    def ref_wgan(recon):
        G = ref_generator(ref_networks, clone(gp)).train()
        D = ref_discriminator(ref_networks, clone(dp), R, False).train()
        go = torch.optim.RMSprop(G.parameters(), lr=0.00005)
        do = torch.optim.RMSprop(D.parameters(), lr=0.00005)
        m = torch.ceil(mask)
        mk = ground * (1 - m)
        inp = mk + G(mk) * m
        for q in D.parameters():
            q.requires_grad = True
        do.zero_grad()
        dr, df = D(ground), D(inp.detach())
        lr_ = dr.mean().view(1)
        lr_.backward(torch.ones(1))
        lf_ = df.mean().view(1)
        lf_.backward(-torch.ones(1))
        dgr = {k: q.grad.clone() for k, q in D.named_parameters()}
        do.step()
        for q in D.parameters():
            q.data.clamp_(-0.01, 0.01)
        for q in D.parameters():
            q.requires_grad = False
        go.zero_grad()
        adv = D(inp).view(-1).mean().view(1)
        rec = nn.L1Loss()(ground, inp) if recon == "l1" else ref_loss["RMSELoss"]()(ground, inp)
        (adv + rec).backward()
        ggr = {k: q.grad.clone() for k, q in G.named_parameters()}
        go.step()
        return dict(inp=inp.detach(), lr=float(lr_), lf=float(lf_), adv=float(adv), rec=float(rec), dgr=dgr, ggr=ggr, G=G, D=D)

    for recon in ("l1", "rmse"):
        r = ref_wgan(recon)
        G_, D_ = steps.enable_grad(clone(gp)), steps.enable_grad(clone(dp))
        o = steps.wgan_step(G_, D_, steps.make_optimizer("rmsprop", G_), steps.make_optimizer("rmsprop", D_), ground, mask, recon=recon)
        assert rel(o["inpainted"], r["inp"]) < 1e-5
        assert abs(o["d_loss_real"] - r["lr"]) < 1e-5 and abs(o["d_loss_fake"] - r["lf"]) < 1e-5
        assert abs(o["g_adv"] - r["adv"]) < 1e-5 and abs(o["recon"] - r["rec"]) < 1e-6
        for k, v in o["d_grads"].items():
            assert rel(v, r["dgr"][k]) < 1e-3, (k, rel(v, r["dgr"][k]))
        for k, v in o["g_grads"].items():
            assert rel(v, r["ggr"][k]) < 1e-3, (k, rel(v, r["ggr"][k]))
        sd_g, sd_d = r["G"].state_dict(), r["D"].state_dict()
        for k in nets.trainable(G_):
            assert rel(G_[k].detach(), sd_g[k]) < 1e-5, k
        for k in nets.trainable(D_):
            assert rel(D_[k].detach(), sd_d[k]) < 1e-5, k
        gold[f"wgan_{recon}"] = dict(
            inpainted=r["inp"], d_loss_real=r["lr"], d_loss_fake=r["lf"], g_adv=r["adv"], recon=r["rec"],
            d_grads={k: summarize(v) for k, v in r["dgr"].items()},
            g_grads={k: summarize(v) for k, v in r["ggr"].items()},
            d_post={k: summarize(v) for k, v in sd_d.items() if v.dtype.is_floating_point},
        )
        print(f"[pin] wgan_{recon} step (reference modules) == oracle.steps.wgan_step")

    def ref_minimax(recon):
        G = ref_generator(ref_networks, clone(gp)).train()
        D = ref_discriminator(ref_networks, clone(dp), R, True).train()
        go = torch.optim.Adam(G.parameters(), lr=0.0002, betas=(0.5, 0.999))
        do = torch.optim.Adam(D.parameters(), lr=0.0002, betas=(0.5, 0.999))
        bce = nn.BCELoss()
        m = torch.ceil(mask)
        mk = ground * (1 - m)
        inp = mk + G(mk) * m
        do.zero_grad()
        pr = D(ground).view(-1)
        lr_ = bce(pr, torch.ones(len(pr)))
        lr_.backward()
        pf = D(inp.detach()).view(-1)
        lf_ = bce(pf, torch.zeros(len(pf)))
        lf_.backward()
        dgr = {k: q.grad.clone() for k, q in D.named_parameters()}
        do.step()
        for q in D.parameters():
            q.requires_grad = False
        go.zero_grad()
        pg = D(inp).view(-1)
        adv = bce(pg, torch.ones(len(pg)))
        rec = nn.L1Loss()(ground, inp) if recon == "l1" else ref_loss["RMSELoss"]()(ground, inp)
        (adv + rec).backward()
        ggr = {k: q.grad.clone() for k, q in G.named_parameters()}
        go.step()
        return dict(inp=inp.detach(), lr=float(lr_), lf=float(lf_), adv=float(adv), rec=float(rec), dgr=dgr, ggr=ggr, G=G, D=D)

    for recon in ("rmse", "l1"):
        r = ref_minimax(recon)
        G_, D_ = steps.enable_grad(clone(gp)), steps.enable_grad(clone(dp))
        o = steps.minimax_step(G_, D_, steps.make_optimizer("adam", G_), steps.make_optimizer("adam", D_), ground, mask, recon=recon)
        assert abs(o["d_loss_real"] - r["lr"]) < 1e-5 and abs(o["d_loss_fake"] - r["lf"]) < 1e-5
        assert abs(o["g_adv"] - r["adv"]) < 1e-5 and abs(o["recon"] - r["rec"]) < 1e-6
        for k, v in o["g_grads"].items():
            assert rel(v, r["ggr"][k]) < 1e-3, (k, rel(v, r["ggr"][k]))
        for k, v in o["d_grads"].items():
            assert rel(v, r["dgr"][k]) < 1e-3, (k, rel(v, r["dgr"][k]))
        gold[f"minimax_{recon}"] = dict(
            inpainted=r["inp"], d_loss_real=r["lr"], d_loss_fake=r["lf"], g_adv=r["adv"], recon=r["rec"],
            d_grads={k: summarize(v) for k, v in r["dgr"].items()},
            g_grads={k: summarize(v) for k, v in r["ggr"].items()},
        )
        print(f"[pin] minimax_{recon} step (reference modules) == oracle.steps.minimax_step")

    # experiment1 (global + local D), reference modules
    dlp = nets.init_patchgan(seed=2, res=R)
    G = ref_generator(ref_networks, clone(gp)).train()
    Dg = ref_discriminator(ref_networks, clone(dp), R, True).train()
    Dl = ref_discriminator(ref_networks, clone(dlp), R, True).train()
    import itertools

    go = torch.optim.Adam(G.parameters(), lr=0.0002, betas=(0.5, 0.999))
    do = torch.optim.Adam(itertools.chain(Dl.parameters(), Dg.parameters()), lr=0.0002, betas=(0.5, 0.999))
    mse, rm = nn.MSELoss(), ref_loss["RMSELoss"]()
    mk = ground * (1 - mask)
    for q in itertools.chain(Dg.parameters(), Dl.parameters()):
        q.requires_grad = False
    go.zero_grad()
    inp = G(mk)
    pg, pl = Dg(inp).view(-1), Dl(mask * inp).view(-1)
    ag, al = mse(pg, torch.ones(len(pg))), mse(pl, torch.ones(len(pl)))
    rg, rl = rm(ground, inp), rm(mask * ground, mask * inp)
    gl = ag + al + 300 * rg + 300 * rl
    gl.backward()
    ggr = {k: q.grad.clone() for k, q in G.named_parameters()}
    go.step()
    for q in itertools.chain(Dg.parameters(), Dl.parameters()):
        q.requires_grad = True
    do.zero_grad()
    fk = inp.detach()
    dl_ = (mse(Dl(ground * mask).view(-1), torch.ones(B)) + mse(Dl(fk * mask).view(-1), torch.zeros(B))
           + mse(Dg(ground).view(-1), torch.ones(B)) + mse(Dg(fk).view(-1), torch.zeros(B)))
    dl_.backward()
    dggr = {k: q.grad.clone() for k, q in Dg.named_parameters()}
    dlgr = {k: q.grad.clone() for k, q in Dl.named_parameters()}
    do.step()
    G_, Dg_, Dl_ = steps.enable_grad(clone(gp)), steps.enable_grad(clone(dp)), steps.enable_grad(clone(dlp))
    o = steps.global_local_step(G_, Dg_, Dl_, steps.make_optimizer("adam", G_), steps.make_optimizer("adam", Dl_, Dg_), ground, mask)
    assert abs(o["g_loss"] - float(gl)) < 1e-3 * abs(float(gl)), (o["g_loss"], float(gl))
    assert abs(o["d_loss"] - float(dl_)) < 1e-5
    for k, v in o["g_grads"].items():
        assert rel(v, ggr[k]) < 1e-3, (k, rel(v, ggr[k]))
    for k, v in o["dg_grads"].items():
        assert rel(v, dggr[k]) < 1e-3, k
    for k, v in o["dl_grads"].items():
        assert rel(v, dlgr[k]) < 1e-3, k
    gold["global_local"] = dict(
        seed_dl=2, inpainted=inp.detach(), g_adv_global=float(ag), g_adv_local=float(al), recon_global=float(rg),
        recon_local=float(rl), g_loss=float(gl), d_loss=float(dl_),
        g_grads={k: summarize(v) for k, v in ggr.items()},
        dg_grads={k: summarize(v) for k, v in dggr.items()},
        dl_grads={k: summarize(v) for k, v in dlgr.items()},
    )
    print("[pin] experiment1 global+local step (reference modules) == oracle.steps.global_local_step")

    gold["ground"], gold["mask"] = ground, mask
    torch.save(gold, os.path.join(OUT, "reference_r128_b2.pt"))
    sz = os.path.getsize(os.path.join(OUT, "reference_r128_b2.pt"))
    print(f"[pin] wrote tests/golden/reference_r128_b2.pt ({sz / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()

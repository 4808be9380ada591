"""Free-running parity table: ACTUAL rel-L2 of every stored activation, every loss and every parameter gradient of one train
step of the CUDA path against the pure-fp32 oracle (oracle.steps, restating wgan_l1.py:110-186 / minimaxgan_rmse.py:110-173 /
experiment1_global_local_D.py:139-200), at BASELINE.json's configs:

    C1  experiment1_global_local_D  128x128, batch 8        C2  wgan_l1  256x256, batch 64      C3  minimaxgan_rmse  256x256, batch 64
    C4  wgan + RMSE + local MSE + TV + VGG19 perceptual / style (random-init VGG19), 256x256, batch 32

    python -m tests.parity_report [--out profiles/parity_r2.json] [--cases C1,C2,C3,...]

Three columns per tensor:
  cuda      rel-L2(CUDA path, fp32 oracle)                              — what the north star's 1e-2 is asked of
  emulated  rel-L2(bf16-storage emulation of the oracle, fp32 oracle)   — oracle.nets.Bf16Policy: the SAME fp32 PyTorch arithmetic
            with tensors rounded to bf16 exactly where the CUDA path stores bf16 (activations, activation gradients, tensor-core
            weight operands).  It is the error ANY implementation with the north star's "NHWC bf16 tiles, fp32 accumulation" has.
  forced    teacher-forced: the oracle continues from the CUDA path's own stored tensors, so each number is a ONE-LAYER error
            (activations) or the error against the exact fp32 vector-Jacobian product of the CUDA forward (gradients).

TEST INFRASTRUCTURE: the only module besides tests/ that touches oracle/.  tests/test_steps_gpu.py asserts on `parity_case`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CASES = {
    # name: (recipe, recon, R, B, seed)
    "C1": ("global_local", "rmse", 128, 8, 24),
    "C1w": ("wgan", "l1", 128, 8, 21),
    "C2": ("wgan", "l1", 256, 64, 31),
    "C3": ("minimax", "rmse", 256, 64, 33),
    # config 4's generator loss without its face-parsing term (wgan_perceptual_style_faceparsing.py:208-221)
    "C4": ("wgan", "rmse", 256, 32, 35, dict(tv_weight=1.0, feature_weights=(0.01, 0.01), local_recon="mse")),
}
VGG_SEED = 11


def _rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _cos(a, b):
    a, b = a.detach().float().flatten(), b.detach().float().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def _recording(nets, base=None):
    class RecordingPolicy(nets.Fp32Policy):
        """Delegates to ``base`` and keeps every tagged storage-point tensor (keyed '<scope>/<tag>')."""

        def __init__(self):
            self.base = base or nets.Fp32Policy()
            self.rec = {}
            self.scope = ""

        def act(self, x, tag=None):
            y = self.base.act(x, tag)
            if tag is not None:
                self.rec[f"{self.scope}/{tag}"] = y.detach()
            return y

        def w(self, x):
            return self.base.w(x)

        def grad(self, x):
            return self.base.grad(x)

    return RecordingPolicy()


def _run_oracle(steps, recipe, recon, params, ground, mask, policy, hooks=None, extras=None):
    cl = lambda p: steps.enable_grad({k: v.clone() for k, v in p.items()})
    hooks = hooks or {}
    extras = extras or {}
    if recipe == "global_local":
        G, Dg, Dl = cl(params["G"]), cl(params["D"]), cl(params["Dl"])
        o = steps.global_local_step(G, Dg, Dl, steps.make_optimizer("adam", G), steps.make_optimizer("adam", Dl, Dg), ground, mask, policy=policy)
        o = dict(o, g_adv=o["g_adv_global"], recon=o["recon_global"], d_grads=o["dg_grads"])
        return o, dict(G=G, D=Dg, Dl=Dl)
    G, D = cl(params["G"]), cl(params["D"])
    if recipe == "wgan":
        o = steps.wgan_step(G, D, steps.make_optimizer("rmsprop", G), steps.make_optimizer("rmsprop", D), ground, mask, recon=recon, policy=policy,
                            after_d_step=hooks.get("after_d_step"), **extras)
    else:
        o = steps.minimax_step(G, D, steps.make_optimizer("adam", G), steps.make_optimizer("adam", D), ground, mask, recon=recon, policy=policy,
                               after_d_step=hooks.get("after_d_step"))
    return o, dict(G=G, D=D)


LOSS_KEYS = {"wgan": ("d_loss_real", "d_loss_fake", "g_adv", "recon", "g_loss"), "minimax": ("d_loss_real", "d_loss_fake", "g_adv", "recon", "g_loss"),
             "global_local": ("g_adv", "g_adv_local", "recon", "recon_local", "g_loss", "d_loss")}


def parity_case(dev, recipe, recon, R, B, seed, extras=None, with_forced=True):
    """One train step of the CUDA path and up to three oracle runs on the same inputs / weights.  Returns the table (dict).
    ``extras``: the extra generator loss terms of config 4 (``InpaintingTrainer`` keywords)."""
    from gan_inpainting_b200 import trainer
    from oracle import nets, steps, synth
    from tests.test_networks_gpu import make_D, make_G
    from tests.test_steps_gpu import adopt_params, forced_policy
    from tests.util import stored_from_disc_ctx, stored_from_gen_ctx

    ground, mask = synth.make_batch(B, R, seed=seed, soft_edge=True)
    params = dict(G=nets.init_unet(0), D=nets.init_patchgan(1, res=R))
    sig = recipe != "wgan"
    G, D = make_G(dev, params["G"]).train(), make_D(dev, params["D"], R, sig).train()
    Dl = None
    if recipe == "global_local":
        params["Dl"] = nets.init_patchgan(2, res=R)
        Dl = make_D(dev, params["Dl"], R, True).train()
    oextras, loss_keys = {}, LOSS_KEYS[recipe]
    if extras:
        from gan_inpainting_b200.lib.models import loss, networks
        from oracle import vgg_losses

        feats = vgg_losses.seeded_vgg_features(VGG_SEED)
        w = networks.VGG19Wrapper()
        w.vgg19.features.load_state_dict(feats.state_dict())
        loss.set_vgg(w.to(dev).eval())
        oextras = dict(tv_weight=extras["tv_weight"], vgg_features=feats, feature_weights=extras["feature_weights"], local_recon=extras["local_recon"])
        loss_keys = loss_keys + ("recon_local", "g_tv", "g_perceptual", "g_style")
    tr = trainer.InpaintingTrainer(recipe, G, D, Dl, recon=recon, **(extras or {}))
    tr.step(ground.to(dev), mask.to(dev))
    s = tr.read_scalars()
    torch.cuda.synchronize()
    if extras:
        loss.set_vgg(None)

    stored = {}
    for scope, c in tr.last["ctx"].items():
        if c is None:
            continue
        stored.update(stored_from_gen_ctx(c, tr.G._cfg["widths"], scope) if scope == "G" else stored_from_disc_ctx(c, scope))
    cuda_grads = {"G": dict(zip(tr.fG.names, [g.detach().cpu().clone() for g in tr.fG.grad_views])),
                  "D": dict(zip(tr.fD.names, [g.detach().cpu().clone() for g in tr.fD.grad_views]))}
    if Dl is not None:
        cuda_grads["Dl"] = dict(zip(tr.fDl.names, [g.detach().cpu().clone() for g in tr.fDl.grad_views]))
    inpainted = tr.last["inpainted"].detach().cpu()

    t0 = time.time()
    p32 = _recording(nets)
    o32, _ = _run_oracle(steps, recipe, recon, params, ground, mask, p32, extras=oextras)
    t_oracle = time.time() - t0
    pem = _recording(nets, nets.Bf16Policy())
    oem, _ = _run_oracle(steps, recipe, recon, params, ground, mask, pem, extras=oextras)

    def grads_of(o):
        g = {"G": o["g_grads"], "D": o["d_grads"]}
        if recipe == "global_local":
            g["Dl"] = o["dl_grads"]
        return g

    g32, gem = grads_of(o32), grads_of(oem)
    table = dict(recipe=recipe, recon=recon, res=R, batch=B, seed=seed, oracle_step_s=round(t_oracle, 2), activations={}, losses={}, grads={},
                 extras={k: (list(v) if isinstance(v, tuple) else v) for k, v in (extras or {}).items()})
    table["image"] = dict(cuda=_rel(inpainted, o32["inpainted"]), emulated=_rel(oem["inpainted"], o32["inpainted"]))
    for k in sorted(p32.rec):
        if k in stored:
            table["activations"][k] = dict(cuda=_rel(stored[k], p32.rec[k]), emulated=_rel(pem.rec[k], p32.rec[k]))
    for k in loss_keys:
        ref = o32[k]
        table["losses"][k] = dict(value_cuda=s[k], value_fp32=ref, value_emulated=oem[k], cuda=abs(s[k] - ref) / max(abs(ref), 1e-30),
                                  emulated=abs(oem[k] - ref) / max(abs(ref), 1e-30))
    for net in g32:
        for name, ref in g32[net].items():
            if ref is None:
                continue
            if float(ref.norm()) < 1e-12:  # e.g. the WGAN head bias: d(mean real - mean fake)/d bias == 0 exactly
                table["grads"][f"{net}/{name}"] = dict(cuda=float(cuda_grads[net][name].norm()), emulated=float(gem[net][name].norm()), zero_reference=True)
                continue
            table["grads"][f"{net}/{name}"] = dict(cuda=_rel(cuda_grads[net][name], ref), emulated=_rel(gem[net][name], ref),
                                                    cos_cuda=_cos(cuda_grads[net][name], ref), cos_emulated=_cos(gem[net][name], ref))
    del p32, pem

    if with_forced:
        pol = forced_policy(tr, nets)
        hooks = {}
        if recipe != "global_local":
            hooks["after_d_step"] = adopt_params(D, {})
        of, _ = _run_oracle(steps, recipe, recon, params, ground, mask, pol, hooks, extras=oextras)
        gf = grads_of(of)
        for k, e in pol.errors.items():
            table["activations"].setdefault(k, {})["forced"] = e
        for k in loss_keys:
            table["losses"][k]["forced"] = abs(s[k] - of[k]) / max(abs(of[k]), 1e-30)
            table["losses"][k]["value_forced"] = of[k]
        for net in gf:
            for name, ref in gf[net].items():
                if ref is None or float(ref.norm()) < 1e-12:
                    continue
                table["grads"][f"{net}/{name}"]["forced"] = _rel(cuda_grads[net][name], ref)
        table["image"]["forced"] = _rel(inpainted, of["inpainted"])

    def worst(section, col):
        vals = [(v[col], k) for k, v in table[section].items() if col in v and not v.get("zero_reference")]
        return dict(max=max(vals)[0], argmax=max(vals)[1], median=sorted(vals)[len(vals) // 2][0]) if vals else None

    table["summary"] = {sec: {col: worst(sec, col) for col in ("cuda", "emulated", "forced")} for sec in ("activations", "losses", "grads")}
    table["mask_bit_exact"] = bool(torch.equal(tr.last["mask"].cpu(), torch.ceil(mask) if recipe != "global_local" else mask))
    return table


def markdown(tables):
    out = ["# Free-running parity table (round 2)", "",
           "rel-L2 against the pure-fp32 oracle step on the same seeded inputs and weights (dropout off).  `cuda` = the sm_100a path, `emulated` = the fp32 oracle with",
           "bf16 rounding at the CUDA path's storage points (what bf16 NHWC tiles + fp32 accumulation cost in ANY implementation), `forced` = teacher-forced",
           "(one-layer error / error against the exact fp32 VJP of the CUDA forward).  Full per-tensor numbers: `parity_r2.json`.", ""]
    for name, t in tables.items():
        ex = " + local MSE + TV + VGG19 perceptual / style (random-init VGG19, seeded)" if t.get("extras") else ""
        out.append(f"## {name}: {t['recipe']} ({t['recon']}){ex}, {t['res']}x{t['res']}, batch {t['batch']} — fp32 oracle step {t['oracle_step_s']} s on the host")
        out.append("")
        out.append("| quantity | cuda max (median) | emulated max (median) | forced max (median) | worst tensor (cuda) |")
        out.append("|---|---|---|---|---|")
        for sec in ("activations", "losses", "grads"):
            s = t["summary"][sec]
            f = lambda c: f"{s[c]['max']:.2e} ({s[c]['median']:.2e})" if s.get(c) else "—"
            out.append(f"| {sec} | {f('cuda')} | {f('emulated')} | {f('forced')} | `{s['cuda']['argmax']}` |")
        im = t["image"]
        out.append(f"| output image | {im['cuda']:.2e} | {im['emulated']:.2e} | {im.get('forced', float('nan')):.2e} | |")
        out.append("")
        out.append("Losses: " + ", ".join(f"{k} {v['value_cuda']:.6g} vs {v['value_fp32']:.6g} (rel {v['cuda']:.1e}; emulated {v['emulated']:.1e})" for k, v in t["losses"].items()))
        out.append("")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "parity_r2.json"))
    ap.add_argument("--cases", default="C1,C1w,C2,C3,C4")
    args = ap.parse_args()
    from gan_inpainting_b200 import build

    build.build(verbose=False)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    tables = {}
    for name in args.cases.split(","):
        t0 = time.time()
        tables[name] = parity_case(dev, *CASES[name])
        print(f"[parity] {name}: {json.dumps(tables[name]['summary'])}  ({time.time() - t0:.0f} s)", flush=True)
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(tables, open(args.out, "w"), indent=1)
    open(os.path.splitext(args.out)[0] + ".md", "w").write(markdown(tables))
    print("[parity] wrote", args.out)


if __name__ == "__main__":
    main()

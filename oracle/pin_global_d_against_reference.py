"""Pin ``oracle.steps.global_d_step`` against the reference's own modules and write tests/golden/reference_global_d.pt.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs only in the build container (/root/reference mounted):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.pin_global_d_against_reference

experiment_list/experiment3_global_D.py is a standalone script (argparse, dataset paths at import time) and cannot be imported;
its per-batch body (:140-196) is executed here statement by statement over the REFERENCE's ``UnetGenerator``,
``PatchGANDiscriminator`` and ``RMSELoss`` with ``nn.BCELoss`` / ``optim.Adam`` exactly as the script constructs them (:113-122),
and compared with the oracle's restatement.  Only numerical outputs are stored.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import nets, steps, synth
from .pin_against_reference import OUT, clone, load_reference, ref_discriminator, ref_generator, rel, summarize


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ref_networks, ref_loss = load_reference()
    R, B = 128, 2
    ground, mask = synth.make_batch(B, R, seed=4321, soft_edge=True)
    gp, dp = nets.init_unet(seed=0), nets.init_patchgan(seed=1, res=R)
    G = ref_generator(ref_networks, clone(gp)).train()
    D = ref_discriminator(ref_networks, clone(dp), R, True).train()
    go = torch.optim.Adam(G.parameters(), lr=0.0002, betas=(0.5, 0.999))  # :121
    do = torch.optim.Adam(D.parameters(), lr=0.0002, betas=(0.5, 0.999))  # :122
    bce, rm = nn.BCELoss(), ref_loss["RMSELoss"]()                        # :118-119
    lambda1 = 300                                                         # :26
    masked = ground * (1 - mask)                                          # :145
    for q in D.parameters():                                              # util.set_requires_grad([D], False), :151
        q.requires_grad = False
    go.zero_grad()
    inp = G(masked)                                                       # :155
    pg = D(inp).view(-1)                                                  # :157
    adv = bce(pg, torch.ones(len(pg)))                                    # :162
    rec = rm(ground, inp)                                                 # :165
    gl = adv + lambda1 * rec                                              # :167
    gl.backward()
    ggr = {k: q.grad.clone() for k, q in G.named_parameters()}
    go.step()
    for q in D.parameters():                                              # :176
        q.requires_grad = True
    do.zero_grad()
    lr_ = bce(D(ground).view(-1), torch.ones(B))                          # :180-181
    lf_ = bce(D(inp.detach()).view(-1), torch.zeros(B))                   # :183-184
    dl = lr_ + lf_                                                        # :186
    dl.backward()
    dgr = {k: q.grad.clone() for k, q in D.named_parameters()}
    do.step()

    G_, D_ = steps.enable_grad(clone(gp)), steps.enable_grad(clone(dp))
    o = steps.global_d_step(G_, D_, steps.make_optimizer("adam", G_), steps.make_optimizer("adam", D_), ground, mask)
    assert rel(o["inpainted"], inp.detach()) < 1e-5
    for name, ref in (("g_adv", adv), ("recon", rec), ("g_loss", gl), ("d_loss_real", lr_), ("d_loss_fake", lf_), ("d_loss", dl)):
        assert abs(o[name] - float(ref)) < 1e-5 * max(1.0, abs(float(ref))), (name, o[name], float(ref))
    for k, v in o["g_grads"].items():
        assert rel(v, ggr[k]) < 1e-3, (k, rel(v, ggr[k]))
    for k, v in o["d_grads"].items():
        assert rel(v, dgr[k]) < 1e-3, (k, rel(v, dgr[k]))
    for k, v in G.state_dict().items():  # post-step weights and BatchNorm buffers
        if v.dtype.is_floating_point:
            assert rel(G_[k].detach(), v) < 1e-4, k
    for k, v in D.state_dict().items():
        if v.dtype.is_floating_point:
            assert rel(D_[k].detach(), v) < 1e-4, k
    gold = dict(res=R, batch=B, seed_inputs=4321, seed_g=0, seed_d=1, ground=ground, mask=mask, inpainted=inp.detach(), g_adv=float(adv), recon=float(rec),
                g_loss=float(gl), d_loss_real=float(lr_), d_loss_fake=float(lf_), d_loss=float(dl),
                g_grads={k: summarize(v) for k, v in ggr.items()}, d_grads={k: summarize(v) for k, v in dgr.items()})
    path = os.path.join(OUT, "reference_global_d.pt")
    torch.save(gold, path)
    print(f"[pin] experiment3 global-D step (reference modules) == oracle.steps.global_d_step; wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()

"""Pin oracle.losses.tv_loss against the reference's ``tv_loss`` (lib/models/loss.py:138-151) and write tests/golden/reference_tv.pt.

TEST INFRASTRUCTURE.  Runs ONLY in the build container (reference mounted read-only at /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.pin_tv_against_reference

The reference's loss.py does not import: it is exec'd with the two undefined names injected, like in
oracle/pin_against_reference.py.  Stores the REFERENCE's loss values and input gradients (fingerprints) for ``cases()``.
"""
from __future__ import annotations

import os

import torch

from . import losses, synth
from .pin_against_reference import load_reference, summarize

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_tv.pt")


def cases():
    """name -> (img, tv_weight): the recipe's use (1-channel inpainted batch, weight 1), odd sizes with 3 channels, a single row."""
    gen = torch.Generator().manual_seed(800)
    return {
        "b3_c1_128_w1": (synth.make_batch(3, 128, seed=801)[0], 1.0),
        "b2_c3_37x53_w0.25": (torch.rand(2, 3, 37, 53, generator=gen), 0.25),
        "b2_c1_1x64_w2": (torch.rand(2, 1, 1, 64, generator=gen), 2.0),
    }


def main():
    _, ns = load_reference()
    ref_tv = ns["tv_loss"]
    gold = {}
    for name, (img, w) in cases().items():
        a = img.clone().requires_grad_(True)
        r = ref_tv(a, w)
        if torch.isfinite(r):
            r.backward()
        b = img.clone().requires_grad_(True)
        o = losses.tv_loss(b, w)
        if torch.isfinite(o):
            o.backward()
        same = (torch.isnan(r) and torch.isnan(o)) or abs(float(r) - float(o)) < 1e-7
        assert same, (name, float(r), float(o))
        gold[name] = dict(loss=float(r), grad=None if a.grad is None else summarize(a.grad))
        print(f"{name}: reference tv_loss {float(r):.6f}")
    torch.save(gold, OUT)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

"""Pin oracle/vgg_losses.py against the reference's own perceptual / style loss functions and write tests/golden/reference_vgg.pt.

    python oracle/pin_vgg_against_reference.py

Runs in the build container only (needs /root/reference).  The reference's lib/models/loss.py is not importable on its own: it
expects ``VGG19Wrapper`` and ``device`` in its namespace and builds a VGG19 with downloaded weights at import time
(loss.py:4-5).  Its source is executed here, unmodified, in a namespace that supplies ``device = cpu`` and a VGG19Wrapper holding a
seeded random-initialised torchvision VGG19 (no download); the functions under test (loss.py:49-136) are the reference's.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vgg_losses  # noqa: E402

REF = "/root/reference/lib/models/loss.py"
SEED = 11


class _Wrapper(torch.nn.Module):
    """Stands in for networks.VGG19Wrapper (networks.py:367-389): same attribute the losses walk (``.vgg19.features``)."""

    def __init__(self):
        super().__init__()
        from torchvision import models

        self.vgg19 = models.vgg19(weights=None)
        self.vgg19.features = vgg_losses.seeded_vgg_features(SEED)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ns = {"VGG19Wrapper": _Wrapper, "device": torch.device("cpu"), "__name__": "reference_loss"}
    with open(REF) as f:
        exec(compile(f.read(), REF, "exec"), ns)
    feats = ns["vgg"].vgg19.features
    g = torch.Generator().manual_seed(5)
    cases = []
    for b, h, w in [(2, 64, 64), (1, 128, 128), (3, 32, 64)]:
        tgt = torch.rand(b, 1, h, w, generator=g)
        out = (tgt + 0.2 * torch.randn(b, 1, h, w, generator=g)).clamp(0, 1)
        rp, rs = ns["perceptual_and_style_loss"](out, tgt, 0.01, 0.01)  # the reference's call site: wgan_perceptual_style_faceparsing.py:216
        rp1, rs1 = ns["perceptual_loss"](out, tgt), ns["style_loss"](out, tgt)
        op, os_ = vgg_losses.perceptual_and_style_loss(feats, out, tgt, 0.01, 0.01)
        op1, os1 = vgg_losses.perceptual_loss(feats, out, tgt), vgg_losses.style_loss(feats, out, tgt)
        for name, r, o in [("p", rp, op), ("s", rs, os_), ("p1", rp1, op1), ("s1", rs1, os1)]:
            assert torch.allclose(r, o, rtol=1e-5, atol=0), (name, float(r), float(o))
        x = torch.rand(2, 64, 8, 8, generator=g)
        assert torch.equal(ns["gram_matrix"](x), vgg_losses.gram_matrix(x)) and torch.equal(ns["gram_matrix"](x, False), vgg_losses.gram_matrix(x, False))
        cases.append(dict(output=out, target=tgt, p=rp.clone(), s=rs.clone(), p_default=rp1.clone(), s_default=rs1.clone()))
        print(f"B={b} {h}x{w}: reference p={float(rp):.6e} s={float(rs):.6e}  oracle p={float(op):.6e} s={float(os_):.6e}")
    wsum = float(sum(p.detach().double().sum() for p in feats.parameters()))
    out_path = os.path.join(ROOT, "tests", "golden", "reference_vgg.pt")
    torch.save(dict(seed=SEED, weight_sum=wsum, cases=cases, torch=str(torch.__version__)), out_path)
    print(f"pinned {len(cases)} cases (rtol 1e-5); wrote {out_path} ({os.path.getsize(out_path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()

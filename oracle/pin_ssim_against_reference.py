"""Pin oracle.ssim against the reference's lib/pytorch_ssim and write tests/golden/reference_ssim.pt.

TEST INFRASTRUCTURE.  Runs ONLY in the build container (reference mounted read-only at /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.pin_ssim_against_reference

Stores the REFERENCE's outputs (not its sources) for the seeded inputs of ``cases()``; tests/test_oracle.py re-checks the oracle
against them anywhere, tests/test_kernels_gpu.py checks the CUDA kernel against both.
"""
from __future__ import annotations

import os
import sys

import torch

from . import ssim as ossim
from . import synth

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_ssim.pt")


def cases():
    """name -> (img1, img2): the experiment's use (ground vs a perturbed copy, 1 channel), odd sizes, 3 channels, identical images."""
    out = {}
    g, _ = synth.make_batch(3, 128, seed=700)
    n = torch.randn(g.shape, generator=torch.Generator().manual_seed(701)) * 0.1
    out["b3_c1_128"] = (g, (g + n).clamp(0, 1))
    gen = torch.Generator().manual_seed(702)
    out["b2_c3_45x70"] = (torch.rand(2, 3, 45, 70, generator=gen), torch.rand(2, 3, 45, 70, generator=gen))
    out["b1_c1_256_same"] = (synth.make_batch(1, 256, seed=703)[0],) * 2
    out["b2_c1_7x9"] = (torch.rand(2, 1, 7, 9, generator=gen), torch.rand(2, 1, 7, 9, generator=gen))
    return out


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from lib import pytorch_ssim as ref  # noqa

    sys.path.pop(0)
    gold = {}
    for name, (a, b) in cases().items():
        r_mean = ref.ssim(a, b)
        r_img = ref.ssim(a, b, size_average=False)
        r_mod = ref.SSIM()(a, b)
        o_mean, o_img = ossim.ssim(a, b), ossim.ssim(a, b, size_average=False)
        assert abs(float(r_mean) - float(o_mean)) < 1e-6 and float((r_img - o_img).abs().max()) < 1e-6 and abs(float(r_mod) - float(r_mean)) < 1e-7, name
        gold[name] = dict(mean=float(r_mean), per_image=r_img.clone())
        print(f"{name}: reference ssim {float(r_mean):.6f}  oracle diff {abs(float(r_mean) - float(o_mean)):.1e}")
    torch.save(gold, OUT)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

"""Step-level parity AT THE HEADLINE SHAPES (GPU): BASELINE.json configs[1] (wgan_l1, 256x256, batch 64 — the bench workload) and
configs[2] (minimaxgan_rmse, same shape), one full train step through ``InpaintingTrainer`` against ``oracle.steps``.  Batch 64 at
256x256 exercises what the 128x128 / batch <= 8 step tests never see: CTA pairs on every layer, >= 148-tile schedules, the
``pick_splits`` choices of the weight-gradient GEMMs and the 169-feature discriminator head.

Both tiers of tests/test_steps_gpu.py, with the free-running tier asserting what is actually achieved (tests/parity_report.py,
profiles/parity_r2.md) instead of a cosine bound:

A. teacher-forced: every layer <= 5e-3, every gradient <= 1e-2, every loss <= 2e-3 (north star: rel-L2 <= 1e-2).  One exception,
   measured and stated: the BatchNorm-beta gradients of the WGAN critic are the DIFFERENCE of two backward passes seeded +1 (real) and
   -1 (fake) (wgan_l1.py:137-141) that nearly cancel, so the bf16 rounding of each pass's activation gradients weighs up to 1.3e-2 of
   the small difference (profiles/parity_r2.json, C2 `D/model.6.bias`); they are held to 2e-2;
B. free-running vs the pure-fp32 oracle: every stored activation, the output image and every loss within 1e-2 — or, where bf16
   storage itself costs more than that, within 1.5x of what the bf16-storage EMULATION of the fp32 oracle (oracle.nets.Bf16Policy:
   same fp32 PyTorch arithmetic, tensors rounded to bf16 where the CUDA path stores bf16) deviates from fp32.  Free-running
   gradients of this GAN step are ill-conditioned (a 2e-3 forward perturbation — bf16 rounding of the weights alone — moves the
   D and inner-G gradients by 10-25 % in the fp32 oracle itself, see profiles/parity_r2.md), so their bound is the emulation's
   own deviation x1.5, never a constant.
"""
import pytest

pytestmark = pytest.mark.gpu


def _close(a, b, tol, floor):
    return abs(a - b) <= tol * max(abs(b), floor)


def _check(t):
    bad = []
    # ---- tier A: teacher-forced
    for k, v in t["activations"].items():
        if "forced" in v and not v["forced"] < 5e-3:
            bad.append(("forced layer", k, v["forced"]))
    for k, v in t["grads"].items():
        cancels = t["recipe"] == "wgan" and k.startswith("D/") and k.endswith(".bias")
        if "forced" in v and not v["forced"] < (2e-2 if cancels else 1e-2):
            bad.append(("forced grad", k, v["forced"]))
    for k, v in t["losses"].items():
        if not _close(v["value_cuda"], v["value_forced"], 2e-3, 1e-2):
            bad.append(("forced loss", k, v["value_cuda"], v["value_forced"]))
    if not t["image"]["forced"] < 1e-4:
        bad.append(("forced image", t["image"]))
    # ---- tier B: free-running vs pure fp32
    lim = lambda v: max(1e-2, 1.5 * v["emulated"])
    for k, v in t["activations"].items():
        if "cuda" in v and not v["cuda"] <= lim(v):
            bad.append(("free-running activation", k, v["cuda"], v["emulated"]))
    for k, v in t["grads"].items():
        if v.get("zero_reference"):
            if not v["cuda"] < 1e-6:
                bad.append(("gradient that is exactly zero in the oracle", k, v["cuda"]))
        elif not v["cuda"] <= lim(v):
            bad.append(("free-running grad", k, v["cuda"], v["emulated"]))
    for k, v in t["losses"].items():
        # critic outputs are O(0.1-1) and their batch mean cancels: the natural scale of a loss that may be ~0 is 0.1
        if not abs(v["value_cuda"] - v["value_fp32"]) <= max(1e-2 * max(abs(v["value_fp32"]), 0.1), 1.5 * abs(v["value_emulated"] - v["value_fp32"])):
            bad.append(("free-running loss", k, v["value_cuda"], v["value_fp32"], v["value_emulated"]))
    if not t["image"]["cuda"] <= lim(t["image"]):
        bad.append(("free-running image", t["image"]))
    if not t["mask_bit_exact"]:
        bad.append(("mask binarisation is not bit-exact",))
    return bad


def _check_feature_losses(t):
    """Config 4's VGG19 perceptual / style terms are O(1e-5) with a random-initialised VGG19, far below the floors of ``_check``:
    relative bounds of their own — teacher-forced (same inpainted image, the oracle's VGG under bf16-storage emulation) 2e-2,
    free-running max(3e-2, 2 x the emulation's own deviation from fp32)."""
    bad = []
    for k in ("g_perceptual", "g_style", "g_tv", "recon_local"):
        v = t["losses"][k]
        if not (v["value_cuda"] > 0 and abs(v["value_cuda"] - v["value_forced"]) <= 2e-2 * abs(v["value_forced"])):
            bad.append(("forced loss term", k, v["value_cuda"], v["value_forced"]))
        if not abs(v["value_cuda"] - v["value_fp32"]) <= max(3e-2 * abs(v["value_fp32"]), 2 * abs(v["value_emulated"] - v["value_fp32"])):
            bad.append(("free-running loss term", k, v["value_cuda"], v["value_fp32"], v["value_emulated"]))
    return bad


@pytest.mark.parametrize("case", ["C2", "C3", "C4"])
def test_headline_step_parity(dev, case):
    """C2 / C3: BASELINE.json configs[1] / configs[2]; C4: configs[3]'s generator loss (TV, local and VGG19 perceptual / style terms;
    256x256, batch 32) without its face-parsing term."""
    from tests.parity_report import CASES, parity_case

    t = parity_case(dev, *CASES[case])
    bad = _check(t) + (_check_feature_losses(t) if case == "C4" else [])
    assert not bad, bad[:12]

"""CPU: the oracle (oracle/) against the committed golden vectors that oracle/pin_against_reference.py recorded from the
REFERENCE modules (/root/reference is not available at test time), plus the properties the precision policies rely on."""
import torch

from oracle import losses, nets, steps, synth
from tests.util import rel, summary_close


# fp32 CPU convolutions re-associate their sums differently with the thread count / oneDNN primitive picked on the
# machine that runs the test, so "same algorithm, same inputs" agrees to a few 1e-5 through 15 BN-coupled layers, not 1e-7.
ORACLE_TOL = 1e-4
# gradients (a second pass through the same re-associated sums) and point samples of large tensors: 1e-3 / 5e-3


def clone(p):
    return {k: v.clone() for k, v in p.items()}


def test_synthetic_inputs_are_deterministic(golden):
    g, m = synth.make_batch(golden["batch"], golden["res"], seed=golden["seed_inputs"], soft_edge=True)
    assert torch.equal(g, golden["ground"]) and torch.equal(m, golden["mask"])
    mc, _ = steps.prepare(g, m)
    assert float(mc.sum()) == golden["mask_ceil_sum"]
    assert losses.masked_pixel_count(m) == golden["masked_pixel_count"]
    assert set(mc.unique().tolist()) <= {0.0, 1.0} and not set(m.unique().tolist()) <= {0.0, 1.0}  # soft edge got ceil'ed


def test_unet_matches_reference_outputs(golden):
    ground, mask = golden["ground"], golden["mask"]
    _, masked = steps.prepare(ground, mask)
    p = steps.enable_grad(nets.init_unet(golden["seed_g"]))
    taps = {}
    out = nets.unet_forward(p, masked, train=True, taps=taps)
    assert rel(out, golden["unet_train_out"]) < ORACLE_TOL
    for name, s in golden["unet_taps"].items():
        ne, se = summary_close(taps[name], s, 0)
        assert ne < ORACLE_TOL and se < 1e-3, name
    (out * torch.linspace(-1, 1, out.numel()).view_as(out)).sum().backward()
    for k, s in golden["unet_grads"].items():
        ne, se = summary_close(p[k].grad, s, 0)
        assert ne < 1e-3 and se < 5e-3, k
    for k, v in golden["unet_running"].items():
        assert rel(p[k], v) < ORACLE_TOL
    with torch.no_grad():
        assert rel(nets.unet_forward(p, masked, train=False), golden["unet_eval_out"]) < ORACLE_TOL


def test_patchgan_matches_reference_outputs(golden):
    for sig, tag in ((True, "sig"), (False, "lin")):
        p = steps.enable_grad(nets.init_patchgan(golden["seed_d"], res=128))
        x = golden["ground"].clone().requires_grad_(True)
        r = nets.patchgan_forward(p, x, train=True, sigmoid=sig)
        assert rel(r, golden[f"patchgan_{tag}_out"]) < ORACLE_TOL
        (r * torch.tensor([[1.0], [-2.0]])).sum().backward()
        ne, se = summary_close(x.grad, golden[f"patchgan_{tag}_dx"], 0)
        assert ne < 1e-3 and se < 5e-3
        for k, s in golden[f"patchgan_{tag}_grads"].items():
            ne, se = summary_close(p[k].grad, s, 0)
            assert ne < 1e-3 and se < 5e-3, k


def test_losses_match_reference_classes(golden):
    yhat, ground, mask = golden["loss_yhat"], golden["ground"], golden["mask"]
    assert abs(float(losses.rmse(yhat, ground)) - golden["rmse"]) < 1e-7
    assert abs(float(losses.local_loss(yhat, ground, mask, "l1")) - golden["local_l1"]) < 1e-6
    assert abs(float(losses.local_loss(yhat, ground, mask, "mse")) - golden["local_mse"]) < 1e-6


def test_wgan_step_matches_reference_modules(golden):
    ground, mask = golden["ground"], golden["mask"]
    for recon in ("l1", "rmse"):
        G, D = steps.enable_grad(nets.init_unet(golden["seed_g"])), steps.enable_grad(nets.init_patchgan(golden["seed_d"], res=128))
        o = steps.wgan_step(G, D, steps.make_optimizer("rmsprop", G), steps.make_optimizer("rmsprop", D), ground, mask, recon=recon)
        g = golden[f"wgan_{recon}"]
        assert rel(o["inpainted"], g["inpainted"]) < ORACLE_TOL
        for k in ("d_loss_real", "d_loss_fake", "g_adv", "recon"):
            assert abs(o[k] - g[k]) < ORACLE_TOL, (recon, k)
        for k, s in g["g_grads"].items():
            ne, _ = summary_close(o["g_grads"][k], s, 0)
            assert ne < 1e-3, k
        for k, s in g["d_post"].items():
            ne, _ = summary_close(D[k], s, 0)
            assert ne < 1e-3, k  # post-step weights: one sign flip of a ~0 gradient element moves the norm by ~1e-4 (sign descent)


def test_minimax_and_global_local_steps_match_reference_modules(golden):
    ground, mask = golden["ground"], golden["mask"]
    G, D = steps.enable_grad(nets.init_unet(golden["seed_g"])), steps.enable_grad(nets.init_patchgan(golden["seed_d"], res=128))
    o = steps.minimax_step(G, D, steps.make_optimizer("adam", G), steps.make_optimizer("adam", D), ground, mask, recon="rmse")
    g = golden["minimax_rmse"]
    for k in ("d_loss_real", "d_loss_fake", "g_adv", "recon"):
        assert abs(o[k] - g[k]) < ORACLE_TOL, k
    G, Dg, Dl = (steps.enable_grad(nets.init_unet(golden["seed_g"])), steps.enable_grad(nets.init_patchgan(golden["seed_d"], res=128)),
                 steps.enable_grad(nets.init_patchgan(golden["global_local"]["seed_dl"], res=128)))
    o = steps.global_local_step(G, Dg, Dl, steps.make_optimizer("adam", G), steps.make_optimizer("adam", Dl, Dg), ground, mask)
    g = golden["global_local"]
    assert abs(o["g_loss"] - g["g_loss"]) < 1e-3 * abs(g["g_loss"]) and abs(o["d_loss"] - g["d_loss"]) < ORACLE_TOL
    for k, s in g["dl_grads"].items():
        ne, _ = summary_close(o["dl_grads"][k], s, 0)
        assert ne < 1e-3, k


def test_policies_share_one_algorithm():
    """Fp32Policy is the identity; Bf16Policy only rounds; TeacherForcedPolicy fed with the oracle's own tensors is a no-op."""
    ground, mask = synth.make_batch(2, 128, seed=1)
    _, masked = steps.prepare(ground, mask)
    p = nets.init_unet(5)
    a = nets.unet_forward(clone(p), masked, train=True)
    b = nets.unet_forward(clone(p), masked, train=True, policy=nets.Fp32Policy())
    assert torch.equal(a, b)
    c = nets.unet_forward(clone(p), masked, train=True, policy=nets.Bf16Policy())
    assert 1e-4 < rel(c, a) < 3e-2  # bf16 storage moves the result by ~1 %
    # teacher forcing with the free-running bf16 tensors reproduces the bf16 run (up to the weights' rounding, also applied)
    d = steps.enable_grad(nets.init_patchgan(6, res=128))
    taps = {}
    r = nets.patchgan_forward(d, ground, train=True, sigmoid=False, taps=taps, policy=nets.Bf16Policy())
    import torch.nn.functional as F

    stored = {f"y{i}": taps[f"c{i + 1}.conv"].detach() for i in range(1, 4)}
    pol = nets.TeacherForcedPolicy(stored)
    r2 = nets.patchgan_forward(clone({k: v.detach() for k, v in d.items()}), ground, train=True, sigmoid=False, policy=pol)
    assert len(pol.errors) == 3 and max(pol.errors.values()) < 5e-3
    assert rel(r2, r) < 5e-3


def test_flop_counts_match_survey():
    assert abs(nets.conv_flops_unet(256) / 1e9 - 11.845) < 0.01
    assert abs(nets.conv_flops_unet(128) / 1e9 - 2.961) < 0.01
    assert abs(nets.conv_flops_patchgan(256) / 1e9 - 3.258) < 0.01
    assert abs(nets.conv_flops_patchgan(128) / 1e9 - 0.814) < 0.01


def test_ssim_oracle_matches_reference_golden():
    """oracle.ssim against the values the REFERENCE's lib/pytorch_ssim produced (oracle/pin_ssim_against_reference.py)."""
    import os
    from oracle import pin_ssim_against_reference as pin
    from oracle import ssim as ossim
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_ssim.pt"))
    for name, (a, b) in pin.cases().items():
        assert abs(float(ossim.ssim(a, b)) - gold[name]["mean"]) < 2e-6, name
        assert float((ossim.ssim(a, b, size_average=False) - gold[name]["per_image"]).abs().max()) < 2e-6, name
    assert abs(gold["b1_c1_256_same"]["mean"] - 1.0) < 1e-6  # identical images


def test_tv_oracle_matches_reference_golden():
    """oracle.losses.tv_loss against the values / gradients the REFERENCE's tv_loss produced (oracle/pin_tv_against_reference.py)."""
    import math
    import os
    from oracle import pin_tv_against_reference as pin
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_tv.pt"))
    for name, (img, w) in pin.cases().items():
        x = img.clone().requires_grad_(True)
        o = losses.tv_loss(x, w)
        if math.isnan(gold[name]["loss"]):
            assert math.isnan(float(o)), name  # a single row: mean over an empty tensor
            continue
        assert abs(float(o) - gold[name]["loss"]) < 1e-6, name
        o.backward()
        ne, se = summary_close(x.grad, gold[name]["grad"], 0)
        assert ne < 1e-5 and se < 1e-5, name


def test_preprocess_oracle_matches_pillow_torchvision_golden():
    """oracle/preprocess.py (greyscale + Resize + ToTensor, lib/data/dataset.py:6-12 / train.py:69-72) against the tensors Pillow +
    torchvision produced in the build container (oracle/pin_preprocess_against_reference.py): bit for bit."""
    import os

    import numpy as np
    from oracle import preprocess

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_preprocess.pt"))
    assert len(g["cases"]) >= 5
    for c in g["cases"]:
        got = torch.from_numpy(preprocess.load_transform(c["image"].numpy(), c["size"]))
        assert got.shape == c["out"].shape and torch.equal(got, c["out"])
    # edge cases: a 1-pixel-high strip, identity size, extreme down-scale
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (1, 40), dtype=np.uint8)
    assert preprocess.load_transform(img, 1).shape == (1, 1, 40)
    img = rng.integers(0, 256, (64, 64), dtype=np.uint8)
    assert np.array_equal(preprocess.load_transform(img, 64)[0], img.astype(np.float32) / np.float32(255))
    assert preprocess.load_transform(rng.integers(0, 256, (512, 512, 3), dtype=np.uint8), 4).shape == (1, 4, 4)


def test_vgg_loss_oracle_matches_reference_golden():
    """oracle/vgg_losses.py against the values the reference's own perceptual / style loss functions returned for the same seeded
    random-initialised VGG19 (oracle/pin_vgg_against_reference.py)."""
    import os

    from oracle import vgg_losses

    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vgg.pt"))
    feats = vgg_losses.seeded_vgg_features(gold["seed"])
    wsum = float(sum(p.detach().double().sum() for p in feats.parameters()))
    assert abs(wsum - gold["weight_sum"]) < 1e-6 * max(1.0, abs(gold["weight_sum"])), "torchvision's seeded VGG19 initialisation changed"
    for c in gold["cases"][:2]:
        p, s = vgg_losses.perceptual_and_style_loss(feats, c["output"], c["target"], 0.01, 0.01)
        assert torch.allclose(p, c["p"], rtol=1e-5, atol=0) and torch.allclose(s, c["s"], rtol=1e-5, atol=0)
    c = gold["cases"][2]
    assert torch.allclose(vgg_losses.perceptual_loss(feats, c["output"], c["target"]), c["p_default"], rtol=1e-5, atol=0)
    assert torch.allclose(vgg_losses.style_loss(feats, c["output"], c["target"]), c["s_default"], rtol=1e-5, atol=0)
    # the bf16-storage emulation stays close to fp32 (it bounds what the CUDA path may differ by)
    pe, se = vgg_losses.perceptual_and_style_loss(feats, c["output"], c["target"], 0.01, 0.01, bf16=True)
    assert abs(float(pe) - float(c["p"])) < 5e-2 * float(c["p"]) and abs(float(se) - float(c["s"])) < 1e-1 * float(c["s"])


def test_global_d_step_matches_reference_modules():
    """oracle.steps.global_d_step against the per-batch body of experiment3_global_D.py:140-196 executed over the reference's own
    modules (oracle/pin_global_d_against_reference.py -> tests/golden/reference_global_d.pt)."""
    import os

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_global_d.pt"))
    G, D = steps.enable_grad(nets.init_unet(g["seed_g"])), steps.enable_grad(nets.init_patchgan(g["seed_d"], res=g["res"]))
    o = steps.global_d_step(G, D, steps.make_optimizer("adam", G), steps.make_optimizer("adam", D), g["ground"], g["mask"])
    assert rel(o["inpainted"], g["inpainted"]) < ORACLE_TOL
    for k in ("g_adv", "recon", "d_loss_real", "d_loss_fake", "d_loss"):
        assert abs(o[k] - g[k]) < ORACLE_TOL, (k, o[k], g[k])
    assert abs(o["g_loss"] - g["g_loss"]) < 1e-5 * abs(g["g_loss"])
    for name, grads in (("g_grads", o["g_grads"]), ("d_grads", o["d_grads"])):
        for k, sm in g[name].items():
            ne, _ = summary_close(grads[k], sm, 0)
            assert ne < 1e-3, (name, k, ne)

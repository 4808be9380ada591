"""Module-level parity (GPU): the drop-in ``lib.models.networks`` modules, called through the reference's own
``nn.Module`` surface (constructor, ``load_state_dict``, ``net(x)``, ``.backward()``, ``p.grad``), against

* the committed golden outputs of the REFERENCE modules (tests/golden/reference_r128_b2.pt, made by
  oracle/pin_against_reference.py from /root/reference), and
* the CPU oracle (oracle.nets) run live on the same seeded inputs at a larger batch.

Two tiers (DESIGN.md "Numerical parity"):

A. TEACHER-FORCED (the implementation-correctness bar, rel-L2 <= 1e-2 as in BASELINE.json's north_star): the oracle
   recomputes every layer from the CUDA path's own previous stored tensor (one-layer error, asserted <= 5e-3) and
   its autograd graph then carries the CUDA path's ReLU/LeakyReLU masks and BatchNorm statistics, so its gradients are
   the exact fp32 vector-Jacobian product of the CUDA forward; every parameter gradient must match it to <= 1e-2.
B. FREE-RUNNING vs the pure fp32 reference: bf16 storage adds ~0.28 % rel-L2 per layer (3 roundings x 0.16 %), i.e.
   ~1.1 % after the generator's 14 layers, and each sign flip of a near-zero pre-activation between the two forwards
   changes a ReLU/LeakyReLU derivative outright, which makes free-running *gradient* rel-L2 ~ sqrt(forward error):
   5-25 % for ANY reduced-precision forward.  Asserted: outputs <= 1.5e-2, discriminator outputs <= 1e-2,
   gradient cosine similarity >= 0.95.
"""
import pytest
import torch

from tests.util import cos, rel, stored_from_disc_ctx, stored_from_gen_ctx, summary_close

pytestmark = pytest.mark.gpu


def _nets():
    from gan_inpainting_b200.lib.models import networks

    return networks


def _oracle():
    from oracle import nets, steps, synth

    return nets, steps, synth


def make_G(dev, params, dropout=False):
    networks = _nets()
    net = networks.UnetGenerator(1, 1, 7, ngf=64, norm_layer=networks.get_norm_layer("batch"), use_dropout=dropout).to(dev)
    net.load_state_dict(params, strict=True)
    return net


def make_D(dev, params, res, sigmoid):
    net = _nets().PatchGANDiscriminator(sigmoid=sigmoid, imagedim=res).to(dev)
    net.load_state_dict(params, strict=True)
    return net


def test_state_dict_layout_matches_reference(dev):
    nets, _, _ = _oracle()
    networks = _nets()
    g = networks.get_network("generator", "unet")
    ref = nets.init_unet(0)
    sd = g.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(sd[k].shape == ref[k].shape for k in ref)
    d = networks.get_network("discriminator", "patchgan")
    refd = nets.init_patchgan(1, res=128)
    assert list(d.state_dict().keys()) == list(refd.keys())
    assert all(d.state_dict()[k].shape == refd[k].shape for k in refd)
    assert sum(p.numel() for p in g.parameters()) == 41_822_849
    # get_network's string 'False' is truthy: dropout on in the two inner ngf*8 blocks
    assert g._get_engine().dropout_levels == {5, 6} if torch.cuda.is_available() else True


def test_generator_vs_reference_golden(dev, golden):
    """Tier B against outputs of the REFERENCE modules themselves (batch 2, 128x128)."""
    nets, steps, _ = _oracle()
    ground, mask = golden["ground"], golden["mask"]
    _, masked = steps.prepare(ground, mask)
    G = make_G(dev, nets.init_unet(golden["seed_g"])).train()
    out = G(masked.to(dev))
    assert out.shape == (2, 1, 128, 128)
    # batch 2: the 2x2 bottleneck normalises over 8 values, which amplifies bf16 rounding (3e-2 here, 1.5e-2 at batch >= 4)
    assert rel(out, golden["unet_train_out"]) < 3e-2, rel(out, golden["unet_train_out"])
    sd = G.state_dict()
    for k, v in golden["unet_running"].items():
        assert rel(sd[k], v) < 2e-2, (k, rel(sd[k], v))
    assert int(sd["model.model.1.model.2.num_batches_tracked"]) == 1
    # eval mode after that one training forward (the reference sequence): running statistics, no dropout
    G.eval()
    with torch.no_grad():
        ev = G(masked.to(dev))
    assert rel(ev, golden["unet_eval_out"]) < 1.5e-2, rel(ev, golden["unet_eval_out"])


@pytest.mark.parametrize("res,batch", [(128, 8), (256, 8)])
def test_generator_vs_oracle_live(dev, res, batch):
    nets, steps, synth = _oracle()
    ground, mask = synth.make_batch(batch, res, seed=77)
    _, masked = steps.prepare(ground, mask)
    params = nets.init_unet(3)
    G = make_G(dev, params).train()
    G.keep_last_context = True
    out = G(masked.to(dev))
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (out * w.to(dev)).sum().backward()
    clone = lambda: steps.enable_grad({k: v.clone() for k, v in params.items()})

    # ---- tier A: teacher-forced
    pol = nets.TeacherForcedPolicy(stored_from_gen_ctx(G._last_ctx, G._cfg["widths"]))
    pa = clone()
    ref = nets.unet_forward(pa, masked, train=True, policy=pol)
    assert len(pol.errors) == 6 + 6 + 6 + 6
    worst = max(pol.errors.items(), key=lambda kv: kv[1])
    assert worst[1] < 5e-3, worst  # every layer, from the CUDA path's own input
    assert rel(out, ref) < 1e-3
    (ref * w).sum().backward()
    bad = [(k, rel(p.grad, pa[k].grad)) for k, p in G.named_parameters() if not rel(p.grad, pa[k].grad) < 1e-2]
    assert not bad, bad
    sd = G.state_dict()
    for k in params:
        if "running" in k:
            assert rel(sd[k], pa[k]) < 1e-3, k

    # ---- tier B: free-running fp32 reference
    pb = clone()
    ref = nets.unet_forward(pb, masked, train=True)
    assert rel(out, ref) < 1.5e-2
    (ref * w).sum().backward()
    low = [(k, cos(p.grad, pb[k].grad)) for k, p in G.named_parameters() if not cos(p.grad, pb[k].grad) > 0.95]
    assert not low, low


@pytest.mark.parametrize("sig", [True, False])
def test_discriminator_vs_reference_golden(dev, golden, sig):
    """Tier B against outputs of the REFERENCE PatchGANDiscriminator (batch 2, 128x128)."""
    nets, _, _ = _oracle()
    tag = "sig" if sig else "lin"
    D = make_D(dev, nets.init_patchgan(golden["seed_d"], res=128), 128, sig).train()
    x = golden["ground"].to(dev).requires_grad_(True)
    r = D(x)
    assert r.shape == (2, 1)
    assert rel(r, golden[f"patchgan_{tag}_out"]) < 1e-2
    (r * torch.tensor([[1.0], [-2.0]], device=dev)).sum().backward()
    ne, se = summary_close(x.grad, golden[f"patchgan_{tag}_dx"], 0)
    assert ne < 0.1 and se < 0.3, (ne, se)  # free-running gradient: norm and strided samples only roughly (tier B)
    for k, p in D.named_parameters():
        ne, _ = summary_close(p.grad, golden[f"patchgan_{tag}_grads"][k], 0)
        assert ne < 0.1, (k, ne)


@pytest.mark.parametrize("res,batch,sig", [(128, 8, False), (256, 4, True)])
def test_discriminator_vs_oracle_live(dev, res, batch, sig):
    nets, steps, synth = _oracle()
    ground, _ = synth.make_batch(batch, res, seed=5)
    params = nets.init_patchgan(4, res=res)
    D = make_D(dev, params, res, sig).train()
    D.keep_last_context = True
    x = ground.to(dev).requires_grad_(True)
    r = D(x)
    w = torch.randn(r.shape, generator=torch.Generator().manual_seed(2))
    (r * w.to(dev)).sum().backward()
    clone = lambda: steps.enable_grad({k: v.clone() for k, v in params.items()})

    # ---- tier A: teacher-forced
    pol = nets.TeacherForcedPolicy(stored_from_disc_ctx(D._last_ctx))
    pa, xa = clone(), ground.clone().requires_grad_(True)
    ref = nets.patchgan_forward(pa, xa, train=True, sigmoid=sig, policy=pol)
    assert len(pol.errors) == 7
    worst = max(pol.errors.items(), key=lambda kv: kv[1])
    assert worst[1] < 5e-3, worst
    assert rel(r, ref) < 1e-3
    (ref * w).sum().backward()
    assert rel(x.grad, xa.grad) < 1e-2
    bad = [(k, rel(p.grad, pa[k].grad)) for k, p in D.named_parameters() if not rel(p.grad, pa[k].grad) < 1e-2]
    assert not bad, bad
    for k in params:
        if "running" in k:
            assert rel(D.state_dict()[k], pa[k]) < 1e-3, k

    # ---- tier B: free-running fp32 reference
    pb, xb = clone(), ground.clone().requires_grad_(True)
    ref = nets.patchgan_forward(pb, xb, train=True, sigmoid=sig)
    # one number per image after five bf16-stored layers and a cancelling 25..169-term head sum: the free-running distance to
    # the pure-fp32 run sits at ~1e-2 by construction (tier A above pins the same output to 1e-3 under identical inputs)
    assert rel(r, ref) < 2e-2
    (ref * w).sum().backward()
    assert cos(x.grad, xb.grad) > 0.95
    low = [(k, cos(p.grad, pb[k].grad)) for k, p in D.named_parameters() if not cos(p.grad, pb[k].grad) > 0.95]
    assert not low, low


def test_frozen_discriminator_skips_weight_gradients(dev):
    """util.set_requires_grad([D], False) (wgan_l1.py:169): only the data gradient flows, p.grad stays None."""
    from gan_inpainting_b200.lib.models import util

    nets, _, synth = _oracle()
    ground, _ = synth.make_batch(4, 128, seed=9)
    D = make_D(dev, nets.init_patchgan(4, res=128), 128, False).train()
    util.set_requires_grad([D], False)
    x = ground.to(dev).requires_grad_(True)
    D(x).mean().backward()
    assert x.grad is not None and float(x.grad.abs().sum()) > 0
    assert all(p.grad is None for p in D.parameters())
    util.set_requires_grad([D], True)
    D(ground.to(dev)).mean().backward()
    assert all(p.grad is not None for p in D.parameters())


def test_two_backwards_accumulate(dev):
    """d_loss_real.backward(one); d_loss_fake.backward(mone) accumulate into the same .grad (wgan_l1.py:137-141)."""
    nets, _, synth = _oracle()
    ground, mask = synth.make_batch(4, 128, seed=10)
    params = nets.init_patchgan(4, res=128)
    D = make_D(dev, params, 128, False).train()
    a, b = ground.to(dev), (ground * (1 - mask)).to(dev)
    pa, pb = D(a), D(b)
    pa.mean().view(1).backward(torch.ones(1, device=dev))
    g1 = {k: p.grad.clone() for k, p in D.named_parameters()}
    pb.mean().view(1).backward(-torch.ones(1, device=dev))
    D2 = make_D(dev, params, 128, False).train()
    D2(a)  # same BN running-stat sequence is irrelevant for grads
    D2(b).mean().view(1).backward(-torch.ones(1, device=dev))
    for k, p in D.named_parameters():
        assert rel(p.grad - g1[k], dict(D2.named_parameters())[k].grad) < 1e-3, k


def test_dropout_train_vs_eval(dev):
    nets, steps, synth = _oracle()
    ground, mask = synth.make_batch(2, 128, seed=11)
    _, masked = steps.prepare(ground, mask)
    G = make_G(dev, nets.init_unet(3), dropout=True).train()
    x = masked.to(dev)
    a, b = G(x), G(x)
    assert not torch.equal(a, b)  # stochastic in train mode
    (a.sum()).backward()          # backward regenerates the same keep mask
    assert all(torch.isfinite(p.grad).all() for p in G.parameters())
    G.eval()
    with torch.no_grad():
        c, d = G(x), G(x)
    assert torch.equal(c, d)


def test_errors(dev):
    from gan_inpainting_b200._lib import B200Error

    networks = _nets()
    G = networks.get_network("generator", "unet").to(dev)
    with pytest.raises(B200Error):
        G(torch.zeros(1, 1, 128, 128))  # CPU tensor: no fallback
    with pytest.raises(B200Error):
        G(torch.zeros(1, 1, 96, 96, device=dev))  # not divisible by 2^7 (the reference mis-shapes silently)
    D = networks.PatchGANDiscriminator().to(dev)
    with pytest.raises(B200Error):
        D(torch.zeros(2, 1, 256, 256, device=dev))  # Linear(25, 1) head only fits 128x128
    with pytest.raises(Exception, match="Invalid generator network name"):
        networks.get_network("generator", "nope")
    with pytest.raises(Exception, match="Invalid discriminator network name"):
        networks.get_network("discriminator", "nope")

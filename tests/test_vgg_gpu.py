"""GPU parity of the VGG19 perceptual / style loss path (lib/models/loss.py:49-136): every new C-ABI kernel
against a plain PyTorch fp32 statement of the same op, and the drop-in loss functions against oracle/vgg_losses.py (itself pinned
against the reference's functions, tests/golden/reference_vgg.pt).

Tolerances: bf16 operands, fp32 accumulation -> rel-L2 <= 1e-2 per kernel (north_star); the two loss scalars are compared with
the fp32 oracle within max(2e-2, 2 x the error of the oracle's own bf16-storage emulation) — the losses are differences of
nearly equal feature maps, so storage rounding is amplified; the emulation measures by how much.
"""
import pytest
import torch
import torch.nn.functional as F

from tests.util import nchw, nhwc, rel

pytestmark = pytest.mark.gpu
BF16_TOL = 1e-2


def _api():
    from gan_inpainting_b200 import _lib, engine

    return _lib, engine


def bf(x):
    return x.to(torch.bfloat16)


def _pack3(w):
    lib, _ = _api()
    cout, cin = w.shape[:2]
    pk = torch.empty(cout, 9 * cin, dtype=torch.bfloat16, device=w.device)
    lib.call("b200_pack_conv3", lib.ptr(w.contiguous()), cout, cin, lib.ptr(pk), lib.stream())
    return pk


def test_pack_conv3_layout(dev):
    torch.manual_seed(0)
    w = torch.randn(96, 64, 3, 3, device=dev)
    pk = _pack3(w)
    assert torch.equal(pk, bf(w.permute(0, 2, 3, 1).reshape(96, -1)))  # k = (kh*3 + kw)*Cin + ci


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 512, 512), (2, 64, 64, 128, 256), (1, 256, 256, 64, 64), (3, 32, 32, 256, 256), (2, 128, 128, 64, 128),
                                            (5, 16, 8, 64, 32), (1, 32, 128, 128, 128), (64, 16, 16, 512, 512), (7, 64, 64, 64, 96)])
def test_conv3_bias_relu(dev, B, H, W, Cin, Cout):
    """3x3 / stride 1 / pad 1 convolution + bias + ReLU (vgg19.features' conv blocks) on the tcgen05 window kernel."""
    lib, eng = _api()
    torch.manual_seed(B * 131 + H + Cin + Cout)
    x = bf(torch.randn(B, Cin, H, W, device=dev))
    w = torch.randn(Cout, Cin, 3, 3, device=dev) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device=dev) * 0.2
    ref = torch.relu(F.conv2d(x.float(), bf(w).float(), b, padding=1))
    xh = nhwc(x)
    y = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
    for hint in (0, eng.pick_bn(B * (H // 16) * (W // 8), Cout)):
        y.fill_(float("nan"))
        lib.call("b200_conv3_gemm", lib.ptr(xh), B, H, W, Cin, Cin, lib.ptr(_pack3(w)), lib.ptr(b), Cout, lib.ptr(y), Cout, 0, hint, lib.stream())
        got = nchw(y).float()
        assert torch.isfinite(got).all()
        assert rel(got, ref) < BF16_TOL, (hint, rel(got, ref))
        assert (got >= 0).all()


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 256, 256, 64, 64), (3, 64, 64, 128, 128), (2, 32, 32, 256, 256), (5, 16, 8, 512, 512), (64, 32, 32, 512, 512)])
def test_conv3_fused_maxpool_equals_separate_kernels(dev, B, H, W, Cin, Cout):
    """conv -> bias -> ReLU -> MaxPool2d(2, 2) in one launch: bit-identical to the un-pooled launch followed by b200_maxpool2 (max of
    bf16-rounded values == bf16 rounding of the max: rounding is monotonic)."""
    lib, eng = _api()
    torch.manual_seed(Cin + H)
    x = bf(torch.randn(B, H, W, Cin, device=dev))
    w = torch.randn(Cout, Cin, 3, 3, device=dev) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device=dev) * 0.2
    pk = _pack3(w)
    full = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    lib.call("b200_conv3_gemm", lib.ptr(x), B, H, W, Cin, Cin, lib.ptr(pk), lib.ptr(b), Cout, lib.ptr(full), Cout, 0, 0, lib.stream())
    ref = torch.empty(B, H // 2, W // 2, Cout, dtype=torch.bfloat16, device=dev)
    lib.call("b200_maxpool2", lib.ptr(full), B, H, W, Cout, lib.ptr(ref), lib.stream())
    for hint in (0, 64, eng.pick_bn(B * (H // 16) * (W // 8), Cout)):
        got = torch.full_like(ref, float("nan"))
        lib.call("b200_conv3_gemm", lib.ptr(x), B, H, W, Cin, Cin, lib.ptr(pk), lib.ptr(b), Cout, lib.ptr(got), Cout, 1, hint, lib.stream())
        assert torch.equal(got, ref), (hint, float((got.float() - ref.float()).abs().max()))


def test_conv3_strided_input_and_output(dev):
    """Pixel strides wider than the channel count on both sides (a slice of a wider NHWC buffer)."""
    lib, _ = _api()
    torch.manual_seed(9)
    B, H, W, Cin, Cout = 2, 32, 32, 64, 64
    buf = bf(torch.randn(B, H, W, 128, device=dev))
    w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.06
    b = torch.zeros(Cout, device=dev)
    out = torch.zeros(B, H, W, 192, dtype=torch.bfloat16, device=dev)
    lib.call("b200_conv3_gemm", lib.ptr(buf), B, H, W, Cin, 128, lib.ptr(_pack3(w)), lib.ptr(b), Cout, lib.ptr(out), 192, 0, 0, lib.stream())
    ref = torch.relu(F.conv2d(nchw(buf[..., :64]).float(), bf(w).float(), b, padding=1))
    assert rel(nchw(out[..., :64]).float(), ref) < BF16_TOL
    assert float(out[..., 64:].abs().max()) == 0.0


def test_conv3_rejects_unsupported_shapes(dev):
    lib, _ = _api()
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=dev)
    with pytest.raises(lib.B200Error):
        lib.call("b200_conv3_gemm", lib.ptr(x), 1, 8, 8, 64, 64, lib.ptr(x), lib.ptr(x), 64, lib.ptr(x), 64, 0, 0, lib.stream())


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (3, 32, 64), (1, 16, 8)])
def test_conv3_first_layer(dev, B, H, W):
    """features[0] on img.repeat(1,3,1,1): one input channel with channel-summed filters."""
    lib, _ = _api()
    torch.manual_seed(3)
    x = torch.rand(B, 1, H, W, device=dev)
    w = torch.randn(64, 3, 3, 3, device=dev) * 0.3
    b = torch.randn(64, device=dev) * 0.1
    ref = torch.relu(F.conv2d(x.repeat(1, 3, 1, 1), w, b, padding=1))
    w1 = w.sum(1).reshape(64, 9).contiguous()
    y = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=dev)
    lib.call("b200_conv3_c1_fwd", lib.ptr(x), B, H, W, lib.ptr(w1), lib.ptr(b), 64, lib.ptr(y), lib.stream())
    assert rel(nchw(y).float(), ref) < 4e-3  # fp32 arithmetic, bf16 store


@pytest.mark.parametrize("B,H,W,C", [(2, 256, 256, 64), (3, 32, 16, 128), (1, 16, 16, 512), (5, 2, 2, 8)])
def test_maxpool2_exact(dev, B, H, W, C):
    lib, _ = _api()
    torch.manual_seed(1)
    x = bf(torch.randn(B, H, W, C, device=dev))
    y = torch.empty(B, H // 2, W // 2, C, dtype=torch.bfloat16, device=dev)
    lib.call("b200_maxpool2", lib.ptr(x), B, H, W, C, lib.ptr(y), lib.stream())
    assert torch.equal(nchw(y), F.max_pool2d(nchw(x).float(), 2, 2).to(torch.bfloat16))


@pytest.mark.parametrize("B,H,W,C", [(2, 256, 256, 64), (3, 128, 128, 128), (2, 64, 64, 256), (4, 32, 32, 512), (64, 16, 16, 512), (1, 8, 8, 64), (2, 16, 4, 128)])
def test_gram_gemm(dev, B, H, W, C):
    """Per-image C x C Gram matrices (loss.py:117-136 before the division) on the tensor cores."""
    lib, _ = _api()
    torch.manual_seed(C + H)
    f = bf(torch.randn(B, H, W, C, device=dev).relu())
    Cp, S = max(C, 128), lib.value("b200_gram_splits", B, H, W, C)
    assert S >= 1 and (H * W // 64) % S == 0 and (H * W // 64) // S <= 64
    G = torch.full((B, S, Cp, C), float("nan"), dtype=torch.float32, device=dev)
    lib.call("b200_gram_gemm", lib.ptr(f), B, H, W, C, C, lib.ptr(G), lib.stream())
    fm = f.float().reshape(B, H * W, C)
    ref = torch.bmm(fm.transpose(1, 2).double(), fm.double()).float()
    assert torch.isfinite(G).all()
    # fp32 tensor-core accumulation truncates: <= 256 accumulation steps per partial sum keeps the (one-sided) drift below 3e-5
    assert rel(G[:, :, :C].sum(1), ref) < 3e-5
    if Cp > C:
        assert float(G[:, :, C:].abs().max()) == 0.0


def test_gram_matrix_function(dev):
    from gan_inpainting_b200.lib.models import loss

    torch.manual_seed(2)
    x = bf(torch.randn(3, 128, 32, 32, device=dev)).float()
    n, c, h, w = x.shape
    ref = torch.bmm(x.reshape(n, c, -1), x.reshape(n, c, -1).transpose(1, 2))
    assert rel(loss.gram_matrix(x, normalize=False), ref) < 1e-5
    assert rel(loss.gram_matrix(x), ref / (h * w * c)) < 1e-5


def test_sqdiff_reductions(dev):
    lib, _ = _api()
    torch.manual_seed(4)
    a, b = bf(torch.randn(3, 64, 64, 128, device=dev)), bf(torch.randn(3, 64, 64, 128, device=dev))
    out = torch.full((2,), 5.0, device=dev)
    part = torch.empty(lib.value("b200_sqdiff_blocks"), device=dev)
    n = a.numel()
    lib.call("b200_sqdiff_bf16", lib.ptr(a), lib.ptr(b), n, 1.0 / n, lib.ptr(part), lib.ptr(out), 0, lib.stream())
    ref = ((a.double() - b.double()) ** 2).mean()
    assert abs(float(out[0]) - float(ref)) < 1e-5 * float(ref) and float(out[1]) == 5.0
    lib.call("b200_sqdiff_bf16", lib.ptr(a), lib.ptr(b), n, 1.0 / n, lib.ptr(part), lib.ptr(out), 1, lib.stream())  # accumulate
    assert abs(float(out[0]) - 2 * float(ref)) < 2e-5 * float(ref)
    # Gram matrices with padded rows: rows >= C are ignored, both operands scaled by gscale before the difference
    B, Cp, C = 3, 128, 64
    S = 4
    ga, gb = torch.randn(B, S, Cp, C, device=dev), torch.randn(B, S, Cp, C, device=dev)
    o2 = torch.zeros(1, device=dev)
    lib.call("b200_sqdiff_gram", lib.ptr(ga), lib.ptr(gb), B, S, Cp, C, 0.25, 1.0 / (B * C * C), lib.ptr(part), lib.ptr(o2), 0, lib.stream())
    ref2 = ((0.25 * ga[:, :, :C].double().sum(1) - 0.25 * gb[:, :, :C].double().sum(1)) ** 2).mean()
    assert abs(float(o2[0]) - float(ref2)) < 1e-5 * float(ref2)


# ------------------------------------------------------------------------------------------------ the loss functions
def _wrapper(dev, seed=11):
    from gan_inpainting_b200.lib.models import networks
    from oracle import vgg_losses

    w = networks.VGG19Wrapper()
    w.vgg19.features.load_state_dict(vgg_losses.seeded_vgg_features(seed).state_dict())
    return w.to(dev).eval()


def _images(dev, B, R, seed):
    g = torch.Generator().manual_seed(seed)
    tgt = torch.rand(B, 1, R, R, generator=g)
    out = (tgt + 0.2 * torch.randn(B, 1, R, R, generator=g)).clamp(0, 1)
    return out.to(dev), tgt.to(dev)


def test_vgg_feature_maps_match_oracle(dev):
    """relu1_1 .. relu5_1 of the engine against the fp32 oracle walk, layer by layer."""
    from oracle import vgg_losses

    w = _wrapper(dev)
    out, _ = _images(dev, 2, 256, 3)
    feats = w._get_engine().features_of(out)
    w._get_engine().fuse_pool = False  # the separate max-pool kernel gives the same bits
    for a, b in zip(feats, w._get_engine().features_of(out)):
        assert torch.equal(a, b)
    w._get_engine().fuse_pool = True
    ref = vgg_losses.feature_maps(w.vgg19.features, out)
    emu = vgg_losses.feature_maps(w.vgg19.features, out, bf16=True)
    assert len(feats) == 5
    for i, (a, r, e) in enumerate(zip(feats, ref, emu)):
        assert tuple(a.shape) == (r.shape[0], r.shape[2], r.shape[3], r.shape[1])
        err, bound = rel(nchw(a).float(), r), max(BF16_TOL, 1.5 * rel(e, r))
        assert err < bound, (i, err, bound)
        assert rel(nchw(a).float(), e) < BF16_TOL, (i, rel(nchw(a).float(), e))  # against the same storage precision: kernel error only


@pytest.mark.parametrize("B,R", [(2, 256), (1, 512)])
def test_perceptual_and_style_loss_match_oracle(dev, B, R):
    from gan_inpainting_b200.lib.models import loss
    from oracle import vgg_losses

    w = _wrapper(dev)
    loss.set_vgg(w)
    try:
        out, tgt = _images(dev, B, R, 17 + R)
        p, s = loss.perceptual_and_style_loss(out, tgt, 0.01, 0.01)  # the reference's call site (wgan_perceptual_style_faceparsing.py:216)
        rp, rs = vgg_losses.perceptual_and_style_loss(w.vgg19.features, out, tgt, 0.01, 0.01)
        ep, es = vgg_losses.perceptual_and_style_loss(w.vgg19.features, out, tgt, 0.01, 0.01, bf16=True)
        for name, got, ref, emu in (("perceptual", p, rp, ep), ("style", s, rs, es)):
            err = abs(float(got) - float(ref)) / abs(float(ref))
            bound = max(2e-2, 2 * abs(float(emu) - float(ref)) / abs(float(ref)))
            assert err < bound, (name, float(got), float(ref), float(emu), err, bound)
        # the single-loss entry points and their default weights (loss.py:49, 71)
        assert abs(float(loss.perceptual_loss(out, tgt)) - 5 * float(p)) < 1e-5 * abs(5 * float(p))
        assert abs(float(loss.style_loss(out, tgt)) - 10 * float(s)) < 1e-5 * abs(10 * float(s))
        assert not p.requires_grad and not s.requires_grad  # no_grad in the reference
        # identical images: both losses are exactly zero
        z = loss.perceptual_and_style_loss(tgt, tgt.clone())
        assert float(z[0]) == 0.0 and float(z[1]) == 0.0
    finally:
        loss.set_vgg(None)


def test_vgg_losses_reject_cpu_and_small_images(dev):
    from gan_inpainting_b200 import _lib
    from gan_inpainting_b200.lib.models import loss

    loss.set_vgg(_wrapper(dev))
    try:
        with pytest.raises(_lib.B200Error):
            loss.perceptual_loss(torch.rand(1, 1, 256, 256), torch.rand(1, 1, 256, 256))
        with pytest.raises(_lib.B200Error):  # relu5_1 of a 128^2 image is 8x8: below the 16x8-pixel tile of the window kernel
            loss.perceptual_loss(torch.rand(1, 1, 128, 128, device=dev), torch.rand(1, 1, 128, 128, device=dev))
    finally:
        loss.set_vgg(None)

"""Kernel-level parity (GPU): every C-ABI entry point against a plain PyTorch fp32 statement of the same op.

Tolerances: bf16 operands with fp32 accumulation -> rel-L2 <= 1e-2 (BASELINE.json north_star); fp32 tail kernels
-> 1e-5; mask / count arithmetic bit-exact.
"""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

from tests.util import nchw, nhwc, rel

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2
F32_TOL = 2e-5


def _lib():
    from gan_inpainting_b200 import _lib

    return _lib


def _eng():
    from gan_inpainting_b200 import engine

    return engine


def bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ packing + GEMMs
def _packs(w):
    eng = _eng()
    p = eng._Packed(w)
    p.refresh(w)
    return p


def test_pack_all_equals_single_layer_packs(dev):
    """One-launch repack of all layers (after the fused optimizer step) is bit-identical to the per-layer kernels."""
    eng = _eng()
    torch.manual_seed(5)
    ws = [torch.randn(d0, d1, 4, 4, device=dev) for d0, d1 in ((128, 64), (64, 256), (512, 512), (1024, 512), (16, 32))]
    ref = [_packs(w) for w in ws]
    new = [eng._Packed(w) for w in ws]
    eng.PackTable().refresh(list(zip(ws, new)))
    for a, b in zip(ref, new):
        assert torch.equal(a.gather, b.gather) and torch.equal(a.phase, b.phase)


@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 16, 64, 128), (3, 8, 256, 512), (8, 2, 512, 512), (2, 64, 64, 128), (5, 4, 512, 512), (48, 32, 128, 256),
                                          # window operand staging (>= 16 x 8 output pixels per image): several channel blocks, N = 256 / 128 / 64
                                          # tiles, CTA pairs and single CTAs, an odd tile count
                                          (3, 32, 256, 512), (64, 64, 128, 256), (2, 128, 64, 64), (5, 64, 256, 128), (37, 32, 64, 64)])
def test_conv_s2_fprop(dev, B, H, Cin, Cout):
    eng = _eng()
    torch.manual_seed(0)
    x = bf(torch.randn(B, H, H, Cin, device=dev))
    w = torch.randn(Cout, Cin, 4, 4, device=dev) / math.sqrt(Cin * 16)
    y = eng.conv_s2(x, _packs(w).gather, Cout)
    ref = nhwc(F.conv2d(nchw(x.float()), bf(w).float(), stride=2, padding=1))
    assert rel(y, ref) < BF16_TOL


@pytest.mark.parametrize("B,Hi,Cin,Cout", [(2, 8, 128, 64), (2, 1, 512, 512), (3, 4, 1024, 512), (2, 32, 512, 128), (2, 64, 256, 64), (20, 32, 128, 64),
                                           (297, 8, 64, 64), (4, 128, 64, 64),  # these three: >= 148 pixel tiles at N = 64 (CTA pairs / odd tile count with masked rows / big)
                                           (5, 16, 1024, 256), (16, 32, 512, 128), (3, 16, 512, 512), (75, 16, 64, 64)])  # window operand staging, more channel blocks
def test_convt_s2_fprop(dev, B, Hi, Cin, Cout):
    eng = _eng()
    torch.manual_seed(1)
    z = bf(torch.randn(B, Hi, Hi, Cin, device=dev))
    w = torch.randn(Cin, Cout, 4, 4, device=dev) / math.sqrt(Cin * 4)
    y = eng.convt_s2(z, Cin, Cin, _packs(w).phase, Cout)
    ref = nhwc(F.conv_transpose2d(nchw(z.float()), bf(w).float(), stride=2, padding=1))
    assert rel(y, ref) < BF16_TOL


@pytest.mark.parametrize("kind,B,H,Cin,Cout", [("conv", 3, 16, 64, 128), ("conv", 40, 32, 64, 128), ("conv", 5, 4, 512, 512), ("convt", 2, 8, 128, 64),
                                               ("convt", 24, 16, 256, 64), ("convt", 3, 2, 512, 512), ("convt", 80, 16, 128, 64), ("convt", 299, 8, 64, 64)])
def test_conv_epilogue_batchnorm_statistics(dev, kind, B, H, Cin, Cout):
    """The persistent GEMM's epilogue hands the following BatchNorm2d (networks.py:310, 341-349) its batch statistics: per-CTA partial
    rows + b200_bn_finalize must equal mean / biased variance of the STORED output, and update the running statistics like torch."""
    eng = _eng()
    torch.manual_seed(7)
    bn = torch.nn.BatchNorm2d(Cout).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    plan = eng.BnPlan(bn, True, eng.Scratch.get(dev))
    x = bf(torch.randn(B, H, H, Cin, device=dev))
    if kind == "conv":
        w = torch.randn(Cout, Cin, 4, 4, device=dev) / math.sqrt(Cin * 16)
        y = eng.conv_s2(x, _packs(w).gather, Cout, plan=plan)
        ref = nhwc(F.conv2d(nchw(x.float()), bf(w).float(), stride=2, padding=1))
    else:
        w = torch.randn(Cin, Cout, 4, 4, device=dev) / math.sqrt(Cin * 4)
        y = eng.convt_s2(x, Cin, Cin, _packs(w).phase, Cout, plan=plan)
        ref = nhwc(F.conv_transpose2d(nchw(x.float()), bf(w).float(), stride=2, padding=1))
    assert rel(y, ref) < BF16_TOL
    assert 1 <= plan.rows.value <= 148
    st = plan.finalize(y, Cout)
    yf = y.float().reshape(-1, Cout)
    mean, var = yf.mean(0), yf.var(0, unbiased=False)
    assert rel(st[0], mean) < 1e-4 and rel(st[1], 1.0 / torch.sqrt(var + bn.eps)) < 1e-4
    assert rel(st[2], bn.weight * st[1]) < 1e-5 and rel(st[3], bn.bias - st[0] * st[2]) < 1e-4
    n = yf.shape[0]
    assert rel(bn.running_mean, 0.1 * mean) < 1e-4 and rel(bn.running_var, 0.9 + 0.1 * var * n / (n - 1)) < 1e-4
    assert int(bn.num_batches_tracked) == 1


def _bn_bwd_reference(y, dz, gamma, eps=1e-5):
    """dgamma, dbeta, dy of training-mode BatchNorm given dz = gradient w.r.t. its output (all [P, C] fp32)."""
    n = y.shape[0]
    mean, var = y.mean(0), y.var(0, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + eps)
    xhat = (y - mean) * rstd
    dgamma, dbeta = (dz * xhat).sum(0), dz.sum(0)
    return dgamma, dbeta, gamma * rstd * (dz - dbeta / n - xhat * dgamma / n)


@pytest.mark.parametrize("B,H,Cin,Cy,skip,norm", [(3, 8, 128, 64, True, True), (20, 16, 128, 64, False, True), (2, 4, 512, 256, True, False),
                                                      (8, 32, 256, 128, True, True), (8, 64, 128, 64, True, False),
                                                      (8, 64, 128, 64, False, False), (80, 16, 128, 64, True, True), (297, 8, 64, 64, False, True)])
def test_fused_leaky_bn_backward_in_conv_dgrad_epilogue(dev, B, H, Cin, Cy, skip, norm):
    """Conv2d data gradient (4-phase kernel) whose epilogue applies the LeakyReLU backward of the layer below, merges the skip-connection
    gradient and reduces (sum dz, sum dz*y) for that layer's BatchNorm; b200_bn_bwd_fused finishes dgamma / dbeta / dy."""
    eng = _eng()
    torch.manual_seed(11)
    sc = eng.Scratch.get(dev)
    y = bf(torch.randn(B, 2 * H, 2 * H, Cy, device=dev) * 1.5 + 0.3)           # forward conv output of the layer below
    gamma, beta = torch.rand(Cy, device=dev) + 0.5, torch.randn(Cy, device=dev) * 0.3
    st = eng.bn_stats(y, Cy, gamma, beta, None, None, None, 0.1, 1e-5, sc) if norm else None
    dyin = bf(torch.randn(B, H, H, Cin, device=dev))                             # gradient w.r.t. the conv output of this layer
    w = torch.randn(Cin, Cy, 4, 4, device=dev) / math.sqrt(Cin * 4)
    g2 = bf(torch.randn(B, 2 * H, 2 * H, 2 * Cy, device=dev)) if skip else None
    epi = eng.ActBwd(y, Cy, st, 1, g2=g2, ldg2=2 * Cy, offg2=0, scratch=sc)
    dz = eng.convt_s2(dyin, Cin, Cin, _packs(w).phase, Cy, epi=epi)
    if norm:
        dz = epi.dz32  # a BatchNorm backward follows: the epilogue keeps dz unrounded in its own fp32 tensor
    g = nhwc(torch.nn.grad.conv2d_input((B, Cy, 2 * H, 2 * H), bf(w).float(), nchw(dyin.float()), stride=2, padding=1))
    yf = y.float()
    z = yf * st[2] + st[3] if norm else yf
    ref = torch.where(z > 0, g + (g2[..., :Cy].float() if skip else 0.0), 0.2 * g)
    assert rel(dz, ref) < (1e-5 if norm else BF16_TOL)
    if norm:
        dgamma, dbeta = torch.zeros(Cy, device=dev), torch.zeros(Cy, device=dev)
        dy = eng.bn_bwd_fused(y, Cy, st, gamma, dz, Cy, 0, epi, Cy, dgamma, dbeta, False, sc)
        rg, rb, rdy = _bn_bwd_reference(yf.reshape(-1, Cy), dz.float().reshape(-1, Cy), gamma)
        assert rel(dgamma, rg) < 1e-3 and rel(dbeta, rb) < 1e-3
        assert rel(dy.reshape(-1, Cy), rdy) < BF16_TOL


@pytest.mark.parametrize("B,H,Cy,Cin,drop,norm", [(3, 16, 128, 64, True, True), (24, 16, 128, 128, False, True), (2, 4, 512, 512, False, False),
                                                   (80, 32, 128, 64, False, True)])
def test_fused_relu_dropout_bn_backward_in_convt_dgrad_epilogue(dev, B, H, Cy, Cin, drop, norm):
    """ConvTranspose2d data gradient (gather kernel): output = [skip half | up half]; the up half goes through the backward of
    relu(dropout(bn(y))) in the epilogue (networks.py:304-314), the skip half is stored untouched."""
    eng, lib = _eng(), _lib()
    torch.manual_seed(12)
    sc = eng.Scratch.get(dev)
    Ho = H // 2
    col0 = Cy if norm else 0            # the innermost level has no concat: all columns fused, no norm
    Cout = col0 + Cy
    y = bf(torch.randn(B, Ho, Ho, Cy, device=dev) + 0.2)
    gamma, beta = torch.rand(Cy, device=dev) + 0.5, torch.randn(Cy, device=dev) * 0.3
    st = eng.bn_stats(y, Cy, gamma, beta, None, None, None, 0.1, 1e-5, sc) if norm else None
    dyin = bf(torch.randn(B, H, H, Cin, device=dev))
    w = torch.randn(Cout, Cin, 4, 4, device=dev) / math.sqrt(Cin * 16)           # ConvTranspose2d weight (in = Cout here, out = Cin)
    seed = 77
    epi = eng.ActBwd(y, Cy, st, 2, col0=col0, dropout=drop, seed=seed, scratch=sc)
    out = eng.conv_s2(dyin, _packs(w).gather, Cout, epi=epi)
    g = nhwc(F.conv2d(nchw(dyin.float()), bf(w).float(), stride=2, padding=1))
    if col0:
        assert rel(out[..., :col0], g[..., :col0]) < BF16_TOL
    yf = y.float()
    z = yf * st[2] + st[3] if norm else yf
    ref = torch.where(z > 0, g[..., col0:], torch.zeros_like(z))
    if drop:
        ones = torch.ones_like(y)
        keep = torch.empty_like(y)
        eng.bn_act(ones, Cy, None, 0, keep, Cy, 0, dropout=True, seed=seed)
        ref = ref * keep.float()
    fused_out = epi.dz32 if norm else out[..., col0:]
    assert rel(fused_out, ref) < (1e-5 if norm else BF16_TOL)
    if norm:
        dgamma, dbeta = torch.zeros(Cy, device=dev), torch.zeros(Cy, device=dev)
        dz = fused_out.reshape(-1, Cy)
        dy = eng.bn_bwd_fused(y, Cy, st, gamma, out, Cout, col0, epi, Cout, dgamma, dbeta, False, sc)
        rg, rb, rdy = _bn_bwd_reference(yf.reshape(-1, Cy), dz, gamma)
        assert rel(dgamma, rg) < 1e-3 and rel(dbeta, rb) < 1e-3
        assert rel(dy.reshape(-1, Cy), rdy) < BF16_TOL


def test_split_k_small_layers(dev):
    """Innermost layers (a few 128-pixel tiles, K = 8192): the wrappers split K over CTAs (fp32 partial tiles + b200_splitk_reduce)
    when no epilogue is fused (inference); same result as the unsplit kernel."""
    eng, lib = _eng(), _lib()
    torch.manual_seed(13)
    B, H, Cin, Cout = 8, 8, 512, 512
    assert eng.pick_ksplit(eng._m_tiles(B, H // 2, H // 2), Cout, 1, 16 * Cin // 64)[0] > 1
    x = bf(torch.randn(B, H, H, Cin, device=dev))
    w = torch.randn(Cout, Cin, 4, 4, device=dev) / math.sqrt(Cin * 16)
    pk = _packs(w)
    y = eng.conv_s2(x, pk.gather, Cout)
    y1 = torch.empty_like(y)
    lib.call("b200_conv_s2_gemm", lib.ptr(x), B, H, H, Cin, lib.ptr(pk.gather), Cout, lib.ptr(y1), Cout, 64, 0, None, None, None, 1, None, lib.stream())
    ref = nhwc(F.conv2d(nchw(x.float()), bf(w).float(), stride=2, padding=1))
    assert rel(y, ref) < BF16_TOL and rel(y, y1) < 5e-3
    z = bf(torch.randn(B, 2, 2, Cin, device=dev))
    wt = torch.randn(Cin, Cout, 4, 4, device=dev) / math.sqrt(Cin * 4)
    assert eng.pick_ksplit(eng._m_tiles(B, 2, 2), Cout, 4, 4 * Cin // 64)[0] > 1
    yt = eng.convt_s2(z, Cin, Cin, _packs(wt).phase, Cout)
    assert rel(yt, nhwc(F.conv_transpose2d(nchw(z.float()), bf(wt).float(), stride=2, padding=1))) < BF16_TOL


@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 16, 64, 128), (4, 8, 256, 512)])
def test_conv_dgrad_is_phase_convt(dev, B, H, Cin, Cout):
    """Conv2d data gradient == the 4-phase transposed-conv kernel with the phase-packed Conv2d weight."""
    eng = _eng()
    torch.manual_seed(2)
    dy = bf(torch.randn(B, H // 2, H // 2, Cout, device=dev))
    w = torch.randn(Cout, Cin, 4, 4, device=dev) / math.sqrt(Cin * 16)
    dx = eng.convt_s2(dy, Cout, Cout, _packs(w).phase, Cin)
    ref = nhwc(torch.nn.grad.conv2d_input((B, Cin, H, H), bf(w).float(), nchw(dy.float()), stride=2, padding=1))
    assert rel(dx, ref) < BF16_TOL


@pytest.mark.parametrize("B,Hi,Cin,Cout", [(2, 8, 128, 64), (3, 4, 1024, 512)])
def test_convt_dgrad_is_gather_conv(dev, B, Hi, Cin, Cout):
    eng = _eng()
    torch.manual_seed(3)
    dy = bf(torch.randn(B, 2 * Hi, 2 * Hi, Cout, device=dev))
    w = torch.randn(Cin, Cout, 4, 4, device=dev) / math.sqrt(Cout * 16)
    dz = eng.conv_s2(dy, _packs(w).gather, Cin)
    # d/dz of conv_transpose2d(z, w) is conv2d(dy, w) with w read as (out=Cin, in=Cout)
    ref = nhwc(F.conv2d(nchw(dy.float()), bf(w).float(), stride=2, padding=1))
    assert rel(dz, ref) < BF16_TOL


@pytest.mark.parametrize("B,Hs,Cs,Cl,acc", [(2, 8, 128, 64, False), (3, 4, 512, 256, True), (2, 32, 256, 64, False), (4, 1, 512, 512, False), (16, 32, 128, 64, False)])
def test_wgrad(dev, B, Hs, Cs, Cl, acc):
    eng = _eng()
    torch.manual_seed(4)
    S = bf(torch.randn(B, Hs, Hs, Cs, device=dev))
    L = bf(torch.randn(B, 2 * Hs, 2 * Hs, Cl, device=dev))
    g0 = torch.randn(Cs, Cl, 4, 4, device=dev)
    g = g0.clone()
    eng.wgrad(S, Cs, Cs, L, g, acc, eng.Scratch.get(dev))
    # conv view: weight (Cs, Cl, 4, 4), input L, output-grad S
    ref = torch.nn.grad.conv2d_weight(nchw(L.float()), (Cs, Cl, 4, 4), nchw(S.float()), stride=2, padding=1)
    if acc:
        ref = ref + g0
    assert rel(g, ref) < 1e-4


# ------------------------------------------------------------------------------------------------ BatchNorm + activations
@pytest.mark.parametrize("P,C", [(2 * 64 * 64, 128), (8 * 4, 512), (3 * 16 * 16, 256), (2 * 2 * 2, 512)])
def test_bn_train_forward(dev, P, C):
    eng = _eng()
    torch.manual_seed(5)
    y = bf(torch.randn(P, C, device=dev) * 2 + 0.5)
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    rm, rv, nbt = torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros((), dtype=torch.long, device=dev)
    st = eng.bn_stats(y, C, gamma, beta, rm, rv, nbt, 0.1, 1e-5, eng.Scratch.get(dev))
    out = torch.empty_like(y)
    out2 = torch.zeros(P, 2 * C, dtype=torch.bfloat16, device=dev)
    eng.bn_act(y, C, st, eng.ACT_LEAKY, out, C, 0, out2, 2 * C, C)
    yf = y.float()
    rm_r, rv_r = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ref = F.batch_norm(yf, rm_r, rv_r, gamma, beta, training=True, momentum=0.1, eps=1e-5)
    assert rel(st[0], yf.mean(0)) < 1e-4
    assert rel(rm, rm_r) < 1e-4 and rel(rv, rv_r) < 1e-4 and int(nbt) == 1
    assert rel(out, F.leaky_relu(ref, 0.2)) < BF16_TOL
    assert rel(out2[:, C:], F.relu(ref)) < BF16_TOL
    assert float(out2[:, :C].abs().max()) == 0.0


@pytest.mark.parametrize("act,two,drop", [(1, True, False), (1, False, False), (2, False, False), (2, False, True), (0, False, False)])
@pytest.mark.parametrize("P,C", [(2 * 16 * 16, 128), (8 * 4, 512)])
def test_bn_act_backward(dev, P, C, act, two, drop):
    eng = _eng()
    torch.manual_seed(6)
    sc = eng.Scratch.get(dev)
    y = bf(torch.randn(P, C, device=dev) * 1.5 + 0.3)
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    st = eng.bn_stats(y, C, gamma, beta, None, None, None, 0.1, 1e-5, sc)
    g1 = bf(torch.randn(P, 2 * C, device=dev))
    g2 = bf(torch.randn(P, C, device=dev)) if two else None
    dgam, dbet = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    seed = 77
    dy = eng.bn_act_bwd(y, C, st, gamma, act, g1, 2 * C, C, g2, C, 0, drop, seed, dgam, dbet, False, sc)
    # reference through autograd
    yr = y.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.batch_norm(yr, None, None, gr, br, training=True, eps=1e-5)
    if drop:  # the keep bit depends only on (seed, element index): read it off a forward pass over ones
        fw = torch.empty(P, C, dtype=torch.bfloat16, device=dev)
        eng.bn_act(bf(torch.ones(P, C, device=dev)), C, None, eng.ACT_RELU, fw, C, 0, dropout=True, seed=seed)
        z = z * (fw.float() != 0).float() * 2.0
    if act == 1:
        a = F.leaky_relu(z, 0.2)
        loss = (a * g1[:, C:].float()).sum()
        if two:
            loss = loss + (F.relu(a) * g2.float()).sum()
    elif act == 2:
        loss = (F.relu(z) * g1[:, C:].float()).sum()
    else:
        loss = (z * g1[:, C:].float()).sum()
    loss.backward()
    assert rel(dy, yr.grad) < 2e-2
    assert rel(dgam, gr.grad) < 5e-3 and rel(dbet, br.grad) < 5e-3


def test_bn_eval_coeffs(dev):
    eng = _eng()
    C = 256
    torch.manual_seed(7)
    g, b, rm, rv = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev), torch.randn(C, device=dev), torch.rand(C, device=dev) + 0.1
    st = eng.bn_eval(C, g, b, rm, rv, 1e-5, dev)
    y = bf(torch.randn(100, C, device=dev))
    out = torch.empty_like(y)
    eng.bn_act(y, C, st, eng.ACT_NONE, out, C, 0)
    ref = F.batch_norm(y.float(), rm, rv, g, b, training=False, eps=1e-5)
    assert rel(out, ref) < BF16_TOL


def test_dropout_rate_and_determinism(dev):
    eng = _eng()
    P, C = 4096, 512
    y = bf(torch.ones(P, C, device=dev))
    a, b, c = (torch.empty_like(y) for _ in range(3))
    eng.bn_act(y, C, None, eng.ACT_RELU, a, C, 0, dropout=True, seed=1)
    eng.bn_act(y, C, None, eng.ACT_RELU, b, C, 0, dropout=True, seed=1)
    eng.bn_act(y, C, None, eng.ACT_RELU, c, C, 0, dropout=True, seed=2)
    assert torch.equal(a, b) and not torch.equal(a, c)
    keep = (a.float() != 0).float().mean().item()
    assert abs(keep - 0.5) < 0.01
    assert set(a.float().unique().tolist()) == {0.0, 2.0}  # Dropout(0.5) scales kept values by 2


def test_dropout_device_seed_counter(dev):
    """b200_set_seed_counter: the dropout seed is advanced by a device counter read when the kernel runs (what lets a captured CUDA
    graph of the step draw fresh masks per replay); counter == 0 is the plain seed; forward and backward stay consistent."""
    eng, lib = _eng(), _lib()
    P, C = 2048, 256
    y = bf(torch.ones(P, C, device=dev))
    ctr = torch.zeros(1, dtype=torch.int32, device=dev)
    base, a0, a1, a1b = (torch.empty_like(y) for _ in range(4))
    eng.bn_act(y, C, None, eng.ACT_RELU, base, C, 0, dropout=True, seed=7)
    lib.value("b200_set_seed_counter", lib.ptr(ctr))
    try:
        eng.bn_act(y, C, None, eng.ACT_RELU, a0, C, 0, dropout=True, seed=7)
        lib.call("b200_step_inc", lib.ptr(ctr), lib.stream())
        eng.bn_act(y, C, None, eng.ACT_RELU, a1, C, 0, dropout=True, seed=7)
        eng.bn_act(y, C, None, eng.ACT_RELU, a1b, C, 0, dropout=True, seed=7)
        # backward under the same counter regenerates the same mask: dz = keep ? 2 g : 0 with g == 1 equals the forward's output
        sc = eng.Scratch.get(dev)
        g = bf(torch.ones(P, C, device=dev))
        dz = eng.bn_act_bwd(y, C, None, None, eng.ACT_RELU, g, C, 0, None, 0, 0, True, 7, None, None, False, sc)
    finally:
        lib.value("b200_set_seed_counter", None)
    assert torch.equal(base, a0) and torch.equal(a1, a1b) and not torch.equal(a0, a1)
    assert abs((a1.float() != 0).float().mean().item() - 0.5) < 0.01
    assert torch.equal(dz, a1)


# ------------------------------------------------------------------------------------------------ 1-channel-side convs
@pytest.mark.parametrize("B,R,C,act,mul", [(2, 64, 64, 1, False), (3, 32, 128, 0, False), (2, 128, 64, 1, True)])
def test_conv_c1_fwd(dev, B, R, C, act, mul):
    lib = _lib()
    torch.manual_seed(8)
    x = torch.randn(B, 1, R, R, device=dev)
    m = (torch.rand(B, 1, R, R, device=dev) > 0.5).float() if mul else None
    w = torch.randn(C, 1, 4, 4, device=dev) * 0.25
    y = torch.empty(B, R // 2, R // 2, C, dtype=torch.bfloat16, device=dev)
    y2 = torch.zeros(B, R // 2, R // 2, 2 * C, dtype=torch.bfloat16, device=dev)
    lib.call("b200_conv_c1_fwd", lib.ptr(x), lib.ptr(m), B, R, R, lib.ptr(w), C, act, lib.ptr(y), lib.ptr(y2), 2 * C, 0, lib.stream())
    ref = F.conv2d(x * m if mul else x, w, stride=2, padding=1)
    if act:
        ref = F.leaky_relu(ref, 0.2)
    assert rel(y, nhwc(ref)) < 4e-3  # only the bf16 rounding of the output
    assert rel(y2[..., :C], nhwc(F.relu(ref))) < 4e-3


@pytest.mark.parametrize("B,H,C,tanh,comp", [(2, 32, 128, True, True), (3, 16, 64, False, False), (1, 64, 128, True, False)])
def test_convt_c1_fwd(dev, B, H, C, tanh, comp):
    lib = _lib()
    torch.manual_seed(9)
    z = bf(torch.randn(B, H, H, C, device=dev))
    w = torch.randn(C, 1, 4, 4, device=dev) / math.sqrt(C * 4)
    bias = torch.randn(1, device=dev) * 0.1 if tanh else None
    base = torch.rand(B, 1, 2 * H, 2 * H, device=dev)
    mul = (torch.rand(B, 1, 2 * H, 2 * H, device=dev) > 0.5).float()
    y = torch.empty(B, 1, 2 * H, 2 * H, device=dev)
    cp = torch.empty_like(y) if comp else None
    lib.call("b200_convt_c1_fwd", lib.ptr(z), B, H, H, C, C, lib.ptr(w), lib.ptr(bias), 3 if tanh else 0, lib.ptr(base if comp else None),
             lib.ptr(mul if (comp or not tanh) else None), lib.ptr(y), lib.ptr(cp), lib.stream())
    ref = F.conv_transpose2d(nchw(z.float()), w, bias, stride=2, padding=1)
    if tanh:
        ref = torch.tanh(ref)
    if comp:
        assert rel(y, ref) < 1e-4
        assert rel(cp, base + ref * mul) < 1e-4
    elif not tanh:
        assert rel(y, ref * mul) < 1e-4
    else:
        assert rel(y, ref) < 1e-4


@pytest.mark.parametrize("B,H,C,mul,acc", [(2, 32, 64, False, False), (3, 32, 128, False, True), (2, 64, 64, True, False)])
def test_wgrad_c1(dev, B, H, C, mul, acc):
    lib, eng = _lib(), _eng()
    torch.manual_seed(10)
    z = bf(torch.randn(B, H, H, C, device=dev))
    g = torch.randn(B, 1, 2 * H, 2 * H, device=dev)
    m = (torch.rand_like(g) > 0.5).float() if mul else None
    dw0 = torch.randn(C, 1, 4, 4, device=dev)
    dw = dw0.clone()
    sc = eng.Scratch.get(dev)
    lib.call("b200_wgrad_c1", lib.ptr(z), B, H, H, C, C, lib.ptr(g), lib.ptr(m), lib.ptr(sc.c1_partial), lib.ptr(dw), int(acc), lib.stream())
    ref = torch.nn.grad.conv2d_weight(g * m if mul else g, (C, 1, 4, 4), nchw(z.float()), stride=2, padding=1)
    if acc:
        ref = ref + dw0
    assert rel(dw, ref) < 1e-4


# ------------------------------------------------------------------------------------------------ discriminator head
@pytest.mark.parametrize("B,S,sig", [(4, 8, True), (3, 16, False)])
def test_head(dev, B, S, sig):
    lib = _lib()
    torch.manual_seed(11)
    C, n = 512, S - 3
    a = bf(torch.randn(B, S, S, C, device=dev))
    w5 = torch.randn(1, C, 4, 4, device=dev) / math.sqrt(C * 16)
    wl = torch.randn(1, n * n, device=dev) / n
    bl = torch.randn(1, device=dev)
    E = torch.empty(S, S, C, device=dev)
    lib.call("b200_head_weight", lib.ptr(w5), lib.ptr(wl), S, C, lib.ptr(E), lib.stream())
    lin, pred, scr = torch.empty(B, device=dev), torch.empty(B, 1, device=dev), torch.empty(B * 8, device=dev)
    lib.call("b200_head_fwd", lib.ptr(a), lib.ptr(E), B, S * S * C, lib.ptr(bl), int(sig), lib.ptr(scr), lib.ptr(lin), lib.ptr(pred), lib.stream())
    ar = nchw(a.float()).requires_grad_(True)
    w5r, wlr, blr = w5.clone().requires_grad_(True), wl.clone().requires_grad_(True), bl.clone().requires_grad_(True)
    r = F.linear(F.conv2d(ar, w5r).flatten(1), wlr, blr)
    rp = torch.sigmoid(r) if sig else r
    assert rel(pred, rp) < 1e-4
    dpred = torch.randn(B, 1, device=dev)
    rp.backward(dpred)
    dlin = torch.empty(B, device=dev)
    lib.call("b200_head_dlin", lib.ptr(dpred), lib.ptr(pred), int(sig), B, lib.ptr(dlin), lib.stream())
    da = torch.empty_like(a)
    lib.call("b200_head_bwd_input", lib.ptr(dlin), lib.ptr(E), B, S * S * C, lib.ptr(da), lib.stream())
    assert rel(da, nhwc(ar.grad)) < BF16_TOL
    dE, dw5, dwl, dbl = torch.empty(S * S * C, device=dev), torch.empty_like(w5), torch.empty_like(wl), torch.empty_like(bl)
    lib.call("b200_head_bwd_params", lib.ptr(dlin), lib.ptr(a), B, S, C, lib.ptr(w5), lib.ptr(wl), lib.ptr(dE), lib.ptr(dw5), lib.ptr(dwl), lib.ptr(dbl), 0,
             lib.stream())
    assert rel(dw5, w5r.grad) < 1e-4 and rel(dwl, wlr.grad) < 1e-4 and rel(dbl, blr.grad) < 1e-4


# ------------------------------------------------------------------------------------------------ masks, losses, seed
def test_mask_prepare_and_composite_bit_exact(dev):
    lib = _lib()
    from oracle import synth

    ground, mask = synth.make_batch(4, 64, seed=3, soft_edge=True)
    ground, mask = ground.to(dev), mask.to(dev)
    mc, mk = torch.empty_like(mask), torch.empty_like(ground)
    lib.call("b200_mask_prepare", lib.ptr(ground), lib.ptr(mask), ground.numel(), 1, lib.ptr(mc), lib.ptr(mk), lib.stream())
    assert torch.equal(mc, torch.ceil(mask))  # bit-exact binarisation (wgan_l1.py:114)
    assert torch.equal(mk, ground * (1 - torch.ceil(mask)))
    g = torch.randn_like(ground)
    out = torch.empty_like(ground)
    lib.call("b200_composite", lib.ptr(mk), lib.ptr(g), lib.ptr(mc), ground.numel(), lib.ptr(out), lib.stream())
    assert rel(out, mk + g * mc) < 1e-7
    # outside the mask the composite returns the input pixels unchanged, bit for bit
    assert torch.equal(out[mc == 0], ground[mc == 0])


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4])
def test_recon_losses_and_gradients(dev, kind):
    lib, eng = _lib(), _eng()
    from oracle import losses as ol, synth

    ground, mask = synth.make_batch(4, 64, seed=5, soft_edge=True)
    ground, mask = ground.to(dev), mask.to(dev)
    torch.manual_seed(12)
    yh = torch.rand_like(ground).requires_grad_(True)
    sc = eng.Scratch.get(dev)
    out = torch.empty(4, device=dev)
    lib.call("b200_recon_loss", kind, lib.ptr(ground), lib.ptr(yh), lib.ptr(mask), ground.numel(), 1e-16, lib.ptr(sc.recon), lib.ptr(out), lib.stream())
    ref = [lambda: (ground - yh).abs().mean(), lambda: ol.rmse(yh, ground), lambda: ol.local_loss(yh, ground, mask, "l1"),
           lambda: ol.local_loss(yh, ground, mask, "mse"), lambda: ol.rmse(mask * yh, mask * ground)][kind]()
    assert abs(float(out[0]) - float(ref)) <= 1e-5 * abs(float(ref))
    if kind in (2, 3):
        words = out.view(torch.int32)[2:4].tolist()
        assert words[0] + (words[1] << 31) == int((mask != 0).sum())  # exact pixel count (loss.py:39)
    ref.backward()
    d = torch.empty_like(ground)
    lib.call("b200_gstep_seed", lib.ptr(ground), lib.ptr(yh.detach()), None, lib.ptr(mask), None, None, kind, 1.0, lib.ptr(out), -1, 0.0, None, None, 0,
             ground.numel(), lib.ptr(d), None, None, None, 0, lib.stream())
    assert rel(d, yh.grad) < 1e-4


def test_gstep_seed_full(dev):
    """Fused seed == autograd through composite + tanh + L1 + an upstream adversarial gradient."""
    lib, eng = _lib(), _eng()
    from oracle import synth

    ground, mask = synth.make_batch(2, 64, seed=6)
    ground, mask = ground.to(dev), mask.to(dev)
    torch.manual_seed(13)
    pre = torch.randn_like(ground).requires_grad_(True)
    dadv = torch.randn_like(ground)
    g = torch.tanh(pre)
    masked = ground * (1 - mask)
    inp = masked + g * mask
    rec = (ground - inp).abs().mean()
    (rec + (inp * dadv).sum()).backward()
    sc = eng.Scratch.get(dev)
    out = torch.empty(4, device=dev)
    lib.call("b200_recon_loss", 0, lib.ptr(ground), lib.ptr(inp.detach()), None, ground.numel(), 1e-16, lib.ptr(sc.recon), lib.ptr(out), lib.stream())
    dpre, dbias = torch.empty_like(ground), torch.zeros(1, device=dev)
    lib.call("b200_gstep_seed", lib.ptr(ground), lib.ptr(inp.detach()), lib.ptr(g.detach()), lib.ptr(mask), lib.ptr(dadv), None, 0, 1.0, lib.ptr(out), -1, 0.0,
             None, None, 1, ground.numel(), lib.ptr(dpre), None, lib.ptr(sc.recon), lib.ptr(dbias), 0, lib.stream())
    assert rel(dpre, pre.grad) < 1e-5
    assert abs(float(dbias) - float(pre.grad.sum())) < 1e-3 * abs(float(pre.grad.abs().sum()))


@pytest.mark.parametrize("mode,target", [(0, 0.0), (1, 1.0), (1, 0.0), (2, 1.0), (2, 0.0)])
def test_adv_loss(dev, mode, target):
    lib = _lib()
    torch.manual_seed(14)
    B = 37
    p = (torch.rand(B, device=dev) * 0.98 + 0.01) if mode else torch.randn(B, device=dev)
    pr = p.clone().requires_grad_(True)
    t = torch.full_like(p, target)
    ref = [lambda: pr.mean(), lambda: F.binary_cross_entropy(pr, t), lambda: F.mse_loss(pr, t)][mode]()
    ref.backward()
    loss, d = torch.empty(1, device=dev), torch.empty(B, device=dev)
    lib.call("b200_adv_loss", lib.ptr(p), B, mode, target, -1.0, lib.ptr(loss), lib.ptr(d), lib.stream())
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert rel(d, -pr.grad) < 1e-5


# ------------------------------------------------------------------------------------------------ optimizers
def test_rmsprop_adam_clamp_match_torch(dev):
    from gan_inpainting_b200 import optim as bo

    torch.manual_seed(15)
    net_a = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5)).to(dev)
    for kind in ("rmsprop", "adam"):
        net_b = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5)).to(dev)
        net_b.load_state_dict(net_a.state_dict())
        net_r = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Linear(17, 5)).to(dev)
        net_r.load_state_dict(net_a.state_dict())
        fg = bo.flatten_parameters(net_b)
        if kind == "rmsprop":
            ob, orf = bo.RMSprop(net_b.parameters(), lr=5e-5, clip=(-0.01, 0.01)), torch.optim.RMSprop(net_r.parameters(), lr=5e-5)
        else:
            ob, orf = bo.Adam(net_b.parameters(), lr=2e-4, betas=(0.5, 0.999)), torch.optim.Adam(net_r.parameters(), lr=2e-4, betas=(0.5, 0.999))
        for it in range(3):
            x = torch.randn(8, 33, device=dev)
            for net, o in ((net_b, ob), (net_r, orf)):
                o.zero_grad()
                net(x).pow(2).sum().backward()
                o.step()
            if kind == "rmsprop":
                for q in net_r.parameters():
                    q.data.clamp_(-0.01, 0.01)
        for pb, pr in zip(net_b.parameters(), net_r.parameters()):
            assert rel(pb, pr) < 1e-5, kind
        assert fg.params.data_ptr() == next(net_b.parameters()).data_ptr()


def test_absmean(dev):
    lib = _lib()
    g = torch.randn(1000, device=dev)
    offs = torch.tensor([0, 100, 100, 1000], dtype=torch.int64, device=dev)
    out = torch.empty(3, device=dev)
    lib.call("b200_absmean", lib.ptr(g), lib.ptr(offs), 3, lib.ptr(out), lib.stream())
    assert abs(float(out[0]) - float(g[:100].abs().mean())) < 1e-6
    assert abs(float(out[2]) - float(g[100:].abs().mean())) < 1e-6


# ------------------------------------------------------------------------------------------------ SSIM metric
def test_ssim_matches_oracle_and_reference_golden(dev):
    """lib.pytorch_ssim.ssim / SSIM (fused separable-Gaussian kernel) against the oracle and the values recorded from the
    reference's lib/pytorch_ssim (tests/golden/reference_ssim.pt): 1 and 3 channels, sizes that are not tile multiples, images
    smaller than the window, identical images (== 1)."""
    import os
    from gan_inpainting_b200.lib import pytorch_ssim
    from oracle import pin_ssim_against_reference as pin
    from oracle import ssim as ossim
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_ssim.pt"))
    for name, (a, b) in pin.cases().items():
        got = pytorch_ssim.ssim(a.to(dev), b.to(dev))
        per = pytorch_ssim.ssim(a.to(dev), b.to(dev), size_average=False)
        mod = pytorch_ssim.SSIM()(a.to(dev), b.to(dev))
        assert got.dim() == 0 and per.shape == (a.shape[0],)
        assert abs(float(got) - gold[name]["mean"]) < 1e-5, (name, float(got), gold[name]["mean"])
        assert float((per.cpu() - gold[name]["per_image"]).abs().max()) < 1e-5, name
        assert abs(float(got) - float(ossim.ssim(a, b))) < 1e-5 and float(mod) == float(got)
    with pytest.raises(NotImplementedError):
        pytorch_ssim.ssim(a.to(dev), b.to(dev), window_size=7)
    with pytest.raises(Exception):
        pytorch_ssim.ssim(a, b)  # CPU tensors: no fallback


# ------------------------------------------------------------------------------------------------ total-variation loss
def test_tv_loss_forward_backward(dev):
    """lib.models.loss.tv_loss (loss.py:138-151; part of g_loss in config 4) against the oracle and the reference golden values:
    loss and the gradient w.r.t. the image (through autograd, with a non-trivial upstream gradient), NaN for a single-row image."""
    import math
    import os
    from gan_inpainting_b200.lib.models import loss as bl
    from oracle import losses as ol
    from oracle import pin_tv_against_reference as pin
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_tv.pt"))
    for name, (img, w) in pin.cases().items():
        x = img.to(dev).requires_grad_(True)
        out = bl.tv_loss(x, w)
        if math.isnan(gold[name]["loss"]):
            assert math.isnan(float(out)), name
            continue
        assert abs(float(out) - gold[name]["loss"]) <= 2e-6 * max(1.0, abs(gold[name]["loss"])), (name, float(out), gold[name]["loss"])
        (3.0 * out).backward()
        xr = img.clone().requires_grad_(True)
        (3.0 * ol.tv_loss(xr, w)).backward()
        assert rel(x.grad, xr.grad) < F32_TOL, name
    with pytest.raises(Exception):
        bl.tv_loss(img, 1.0)  # CPU tensor: no fallback


def test_gpu_load_transform_bit_exact(dev):
    """b200_preprocess (greyscale + Resize + ToTensor of a batch on the GPU; lib/data/dataset.py:6-12, train.py:69-72): bit-exact
    against the tensors Pillow + torchvision produced (tests/golden/reference_preprocess.pt) and against the oracle on a batch."""
    import os

    import numpy as np
    from gan_inpainting_b200 import data
    from oracle import preprocess

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_preprocess.pt"))
    for c in g["cases"]:
        t = data.GpuLoadTransform(c["size"], dev)
        out = t(c["image"][None].to(dev))
        assert out.shape[1:] == c["out"].shape and torch.equal(out[0].cpu(), c["out"])
    rng = np.random.default_rng(3)
    batch = rng.integers(0, 256, (5, 200, 168, 3), dtype=np.uint8)
    out = data.GpuLoadTransform(64, dev)(torch.from_numpy(batch))
    ref = np.stack([preprocess.load_transform(b, 64) for b in batch])
    assert torch.equal(out.cpu(), torch.from_numpy(ref))
    with pytest.raises(_lib().B200Error):
        data.GpuLoadTransform(64, dev)(torch.zeros(2, 8, 8, 2, dtype=torch.uint8))

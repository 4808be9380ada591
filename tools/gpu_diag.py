"""Layer-by-layer diagnosis of the CUDA networks against the CPU oracle (prints rel-L2 per tap / gradient).
Run on a B200 via gpurun:  python tools/gpu_diag.py [res] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gan_inpainting_b200.lib.models import networks
from oracle import nets, steps, synth

dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nchw(x):
    return x.permute(0, 3, 1, 2)


POLICY = None


def diag_D(sig):
    print(f"=== discriminator R={R} B={B} sigmoid={sig} policy={POLICY}")
    ground, _ = synth.make_batch(B, R, seed=5)
    params = nets.init_patchgan(4, res=R)
    D = networks.PatchGANDiscriminator(sigmoid=sig, imagedim=R).to(dev).train()
    D.load_state_dict(params)
    eng = D._get_engine()
    x = ground.to(dev)
    c = eng.forward(x, True, sig)
    po = steps.enable_grad({k: v.clone() for k, v in params.items()})
    xr = ground.clone().requires_grad_(True)
    taps = {}
    ref = nets.patchgan_forward(po, xr, train=True, sigmoid=sig, taps=taps, policy=POLICY)
    import torch.nn.functional as F

    print("  a0 (leaky c1)", rel(nchw(c.a[0]), F.leaky_relu(taps["c1.conv"], 0.2)))
    for i in range(1, 4):
        print(f"  y{i} (c{i+1}.conv)", rel(nchw(c.y[i]), taps[f"c{i+1}.conv"]))
        bn = ["model.3", "model.6", "model.9"][i - 1]
        t = taps[f"c{i+1}.conv"].detach()
        a_ref = F.leaky_relu(F.batch_norm(t, None, None, po[bn + ".weight"].detach(), po[bn + ".bias"].detach(), training=True), 0.2)
        print(f"  a{i}", rel(nchw(c.a[i]), a_ref), " mean err", rel(c.st[i][0], t.mean((0, 2, 3))), " rstd err", rel(c.st[i][1], 1 / torch.sqrt(t.var((0, 2, 3), unbiased=False) + 1e-5)))
    print("  pred", rel(c.pred, ref), c.pred.flatten()[:4].tolist(), ref.flatten()[:4].tolist())
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    (ref * w).sum().backward()
    # engine backward with intermediate taps
    from gan_inpainting_b200 import _lib
    dpred = w.to(dev).contiguous()
    dlin = torch.empty(B, device=dev)
    _lib.call("b200_head_dlin", _lib.ptr(dpred), _lib.ptr(c.pred), int(sig), B, _lib.ptr(dlin), _lib.stream())
    grads = {id(p): torch.zeros_like(p) for p in D.parameters()}
    dx = eng.backward(c, dlin, grads, False, input_grad=True)
    print("  dx", rel(dx, xr.grad))
    for k, p in D.named_parameters():
        print(f"  grad {k:18s}", rel(grads[id(p)], po[k].grad), float(po[k].grad.norm()))


def diag_G():
    print(f"=== generator R={R} B={B} policy={POLICY}")
    ground, mask = synth.make_batch(B, R, seed=77)
    _, masked = steps.prepare(ground, mask)
    params = nets.init_unet(3)
    G = networks.UnetGenerator(1, 1, 7, ngf=64, norm_layer=networks.get_norm_layer("batch"), use_dropout=False).to(dev).train()
    G.load_state_dict(params)
    eng = G._get_engine()
    c = eng.forward(masked.to(dev), True)
    po = steps.enable_grad({k: v.clone() for k, v in params.items()})
    taps = {}
    ref = nets.unet_forward(po, masked, train=True, taps=taps, policy=POLICY)
    for l in range(2, 8):
        print(f"  d{l}.conv", rel(nchw(c.y[l]), taps[f"d{l}.conv"]))
    for l in range(7, 1, -1):
        print(f"  u{l}.conv", rel(nchw(c.yu[l]), taps[f"u{l}.conv"]), "  u-in", rel(nchw(c.r_in if l == 7 else c.cat[l]), taps[f"u{l}.in"]))
    print("  u1.in", rel(nchw(c.cat[1]), taps["u1.in"]))
    print("  out", rel(c.g_out, ref))
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (ref * w).sum().backward()
    out = G(masked.to(dev))
    (out * w.to(dev)).sum().backward()
    for k, p in G.named_parameters():
        print(f"  grad {k:60s}", rel(p.grad, po[k].grad), float(po[k].grad.norm()))


if __name__ == "__main__":
    torch.backends.cudnn.allow_tf32 = False
    for POLICY in (nets.Bf16Policy, None):
        diag_D(False)
        diag_D(True)
        diag_G()

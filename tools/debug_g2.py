"""Per-layer check of the generator's fused backward against torch fp32 computed from the engine's own tensors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gan_inpainting_b200 import engine as eng
from gan_inpainting_b200.lib.models import networks
from oracle import nets, synth

R, B = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))
nchw = lambda t: t.permute(0, 3, 1, 2).float()
nhwc = lambda t: t.permute(0, 2, 3, 1)
ground, mask = synth.make_batch(B, R, seed=77)
G = networks.UnetGenerator(1, 1, 7, ngf=64, norm_layer=networks.get_norm_layer("batch"), use_dropout=False).to(dev).train()
G.load_state_dict(nets.init_unet(3))
e = G._get_engine()
x = (ground * (1 - torch.ceil(mask))).to(dev)
c = e.forward(x, True)
wgt = torch.randn(c.g_out.shape, generator=torch.Generator().manual_seed(1)).to(dev)
dpre = (wgt * (1 - c.g_out ** 2)).contiguous()
grads = {id(p): torch.zeros_like(p) for p in G.parameters()}
tr = {}
e.backward(c, dpre, grads, trace=tr)
w = e.w
N = e.N
bf = lambda t: t.to(torch.bfloat16).float()
for l in range(N, 1, -1):
    dy, dz = tr[f"dy{l}"], tr[f"dz{l-1}"]
    Wd = bf(e.dc[l - 1].weight.data)  # (w[l], w[l-1], 4, 4)
    g = nhwc(torch.nn.grad.conv2d_input((B, w[l - 1], R >> (l - 1), R >> (l - 1)), Wd, nchw(dy), stride=2, padding=1))
    g2 = tr_prev = (tr[f"dcat{l-1}"] if l - 1 >= 2 else None)
    if l - 1 >= 2:
        y = c.y[l - 1].float(); st = c.sd[l - 1]
        z = y * st[2] + st[3]
        skip = tr[f"dcat{l-1}"][..., : w[l - 1]].float()
    else:
        z = c.a[1].float()
        skip = None  # dcat1 not traced; compare only where needed
    if skip is not None:
        ref = torch.where(z > 0, g + skip, 0.2 * g)
        print(f"level {l}->{l-1}: dz err {rel(dz, ref):.3e}  |dz| {float(dz.float().norm()):.3e}")
        # BN backward of level l-1 from the engine's dz
        if l - 1 >= 2:
            yf = y.reshape(-1, w[l - 1]); dzf = dz.float().reshape(-1, w[l - 1])
            n = yf.shape[0]
            mean, var = yf.mean(0), yf.var(0, unbiased=False)
            rstd = 1 / torch.sqrt(var + 1e-5)
            xh = (yf - mean) * rstd
            dg, db = (dzf * xh).sum(0), dzf.sum(0)
            gam = e.dn[l - 2].weight.data
            dyref = gam * rstd * (dzf - db / n - xh * dg / n)
            dyn = tr[f"dy{l-1}"].float().reshape(-1, w[l - 1])
            print(f"    dy{l-1} err {rel(dyn, dyref):.3e}  dgamma err {rel(grads[id(e.dn[l-2].weight)], dg):.3e} dbeta err {rel(grads[id(e.dn[l-2].bias)], db):.3e}  |dy|/|dz| {float(dyn.norm()/dzf.norm()):.3f}")
    # weight gradient of level l from dy and a[l-1]
    a_in = c.a[l - 1].float()
    wref = torch.nn.grad.conv2d_weight(nchw(a_in), tuple(e.dc[l - 1].weight.shape), nchw(dy), stride=2, padding=1)
    print(f"    wgrad d{l} err {rel(grads[id(e.dc[l-1].weight)], wref):.3e}")

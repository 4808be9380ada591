"""Debug: fused activation-backward epilogue vs the separate kernels, inside a real discriminator backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_inpainting_b200 import engine as eng
from gan_inpainting_b200.lib.models import networks
from oracle import nets, synth

dev = torch.device("cuda:0")
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))
ground, mask = synth.make_batch(8, 128, seed=21)
D = networks.PatchGANDiscriminator(sigmoid=False, imagedim=128).to(dev).train()
D.load_state_dict(nets.init_patchgan(1, res=128))
e = D._get_engine()
sc = eng.Scratch.get(dev)
for rep in range(2):
    c = e.forward(ground.to(dev), True, False)
    B = 8
    dlin = torch.full((B,), 1.0 / B, device=dev) * (1 if rep == 0 else -1)
    S = 8
    da = torch.empty_like(c.a[3])
    eng.call("b200_head_bwd_input", eng.ptr(dlin), eng.ptr(e.E), B, S * S * 512, eng.ptr(da), eng.stream())
    bn = e.norms[3]
    dg, db = torch.zeros(512, device=dev), torch.zeros(512, device=dev)
    dy = eng.bn_act_bwd(c.y[3], 512, c.st[3], bn.weight.data, 1, da, 512, 0, None, 0, 0, False, 0, dg, db, False, sc)
    for i in (3, 2, 1):
        C = 64 << i
        Cy = C // 2
        # old path
        g = eng.convt_s2(dy, C, C, e.packed[i].phase, Cy)
        if i - 1 >= 1:
            bnl = e.norms[i - 1]
            dg0, db0 = torch.zeros(Cy, device=dev), torch.zeros(Cy, device=dev)
            dy_old = eng.bn_act_bwd(c.y[i - 1], Cy, c.st[i - 1], bnl.weight.data, 1, g, Cy, 0, None, 0, 0, False, 0, dg0, db0, False, sc)
            epi = eng.ActBwd(c.y[i - 1], Cy, c.st[i - 1], 1, scratch=sc)
            dz = eng.convt_s2(dy, C, C, e.packed[i].phase, Cy, epi=epi)
            dg1, db1 = torch.zeros(Cy, device=dev), torch.zeros(Cy, device=dev)
            dy_new = eng.bn_bwd_fused(c.y[i - 1], Cy, c.st[i - 1], bnl.weight.data, dz, Cy, 0, epi, Cy, dg1, db1, False, sc)
            z = c.y[i - 1].float() * c.st[i - 1][2] + c.st[i - 1][3]
            dz_ref = torch.where(z > 0, g.float(), 0.2 * g.float())
            print(rep, i, "rows", epi.rows.value, "dz", rel(dz, dz_ref), "dgamma", rel(dg1, dg0), "dbeta", rel(db1, db0), "dy", rel(dy_new, dy_old),
                  "dbeta vs sum dz", rel(db1, dz.float().reshape(-1, Cy).sum(0)), rel(db0, dz_ref.reshape(-1, Cy).sum(0)))
            dy = dy_new
        else:
            epi = eng.ActBwd(c.a[0], 64, None, 1)
            dz = eng.convt_s2(dy, C, C, e.packed[i].phase, Cy, epi=epi)
            dz_ref = torch.where(c.a[0].float() > 0, g.float(), 0.2 * g.float())
            print(rep, i, "dz0", rel(dz, dz_ref))
torch.cuda.synchronize()

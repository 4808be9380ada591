import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_inpainting_b200 import engine as eng
from gan_inpainting_b200.lib.models import networks
from oracle import nets, synth

dev = torch.device("cuda:0")
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))
ground, mask = synth.make_batch(4, 128, seed=10)
params = nets.init_patchgan(4, res=128)
def mk():
    D = networks.PatchGANDiscriminator(sigmoid=False, imagedim=128).to(dev).train()
    D.load_state_dict(params)
    return D
a, b = ground.to(dev), (ground * (1 - mask)).to(dev)
# determinism of a single backward
res = []
for rep in range(3):
    D = mk()
    D(a)
    D(b).mean().view(1).backward(-torch.ones(1, device=dev))
    res.append({k: p.grad.clone() for k, p in D.named_parameters()})
for k in res[0]:
    print("determinism", k, rel(res[1][k], res[0][k]), rel(res[2][k], res[0][k]))
D = mk()
pa, pb = D(a), D(b)
pa.mean().view(1).backward(torch.ones(1, device=dev))
g1 = {k: p.grad.clone() for k, p in D.named_parameters()}
pb.mean().view(1).backward(-torch.ones(1, device=dev))
for k, p in D.named_parameters():
    print("two-backward", k, rel(p.grad - g1[k], res[0][k]), float(g1[k].norm()), float(res[0][k].norm()))

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_inpainting_b200.lib.models import networks
from oracle import nets, synth
R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
ground, mask = synth.make_batch(B, R, seed=3)
G = networks.UnetGenerator(1, 1, 7, ngf=64, norm_layer=networks.get_norm_layer("batch"), use_dropout=False).to(dev).train()
G.load_state_dict(nets.init_unet(0))
x = (ground * (1 - mask)).to(dev)
out = G(x)
(out * torch.linspace(-1, 1, out.numel(), device=dev).view_as(out)).sum().backward()
torch.cuda.synchronize()
print("ok", float(out.abs().mean()), sum(float(p.grad.abs().sum()) for p in G.parameters()))

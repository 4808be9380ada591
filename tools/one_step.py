"""A few full train steps at the bench configuration (for ncu captures): python tools/one_step.py [steps] [batch] [res] [recipe]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gan_inpainting_b200.lib.models import networks
from gan_inpainting_b200.trainer import InpaintingTrainer
from gan_inpainting_b200.data import synthetic_batch

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
R = int(sys.argv[3]) if len(sys.argv) > 3 else 256
recipe = sys.argv[4] if len(sys.argv) > 4 else "wgan"
dev = torch.device("cuda:0")
torch.manual_seed(0)
G = networks.get_network("generator", "unet").to(dev).train()
D = networks.PatchGANDiscriminator(sigmoid=(recipe != "wgan"), imagedim=R).to(dev).train()
tr = InpaintingTrainer(recipe, G, D, recon="l1")
ground, mask = synthetic_batch(B, R, seed=1234)
ground, mask = ground.to(dev), mask.to(dev)
for _ in range(steps):
    tr.step(ground, mask, update_g=True)
torch.cuda.synchronize()
print("ok", tr.read_scalars())

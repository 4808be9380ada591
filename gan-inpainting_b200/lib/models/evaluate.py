"""Drop-in for the reconstruction part of the reference's train-time evaluation pass, ``lib/models/evaluate.py:87-177``
(``calculate_metric``), called every ``evalevery`` epochs from the recipes (``wgan_l1.py:251-259``).

Per batch (evaluate.py:127-143): ``mask = ceil(mask)`` [``1 - mask`` if ``is_flip_mask``] -> ``masked = input * (1 - m)`` ->
``out = net(masked) * m + masked`` -> global RMSE / L1 and the masked ("local") RMSE / L1 micro-averaged over ``(m != 0)``
pixels; the returned dict carries the means over batches (evaluate.py:166-170).  Everything runs on the sm_100a kernels
(generator forward with the composite fused into the last layer, fused loss reductions); the four running sums stay on the
device and are read back once.

Deviations, all forced by the reference being unrunnable here (``evaluate.py`` is a ``TabError``):
* ``recon_rmse_local`` follows the intent of ``LocalLoss(RMSELoss)`` (loss.py:30-31, 44-45: masked MSE micro-average, then
  ``sqrt(. + eps)``) — the constructor call as written raises;
* FID needs downloaded Inception weights and JPEG round trips (``fid = -1`` like the reference's "no stats" case,
  evaluate.py:160-163); the face-parsing metrics reference undefined names (evaluate.py:146) and are out of scope
   — the key is kept with ``None``;
* the generator runs in the mode it is in (the reference never calls ``.eval()``).
"""
from __future__ import annotations

import torch

from ... import _lib
from ...data import DevicePrefetcher
from ...engine import Scratch
from ..._lib import call, ptr, stream

RECON_L1, RECON_RMSE, RECON_LOCAL_L1, RECON_LOCAL_MSE = 0, 1, 2, 3


@torch.no_grad()
def calculate_metric(device, loader, net, fid_stats=None, mode="test", inception_model=None, epoch=None, is_flip_mask=False, eps=1e-16):
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.B200Error("calculate_metric runs on CUDA (sm_100a) only; there is no CPU fallback")
    eng = net._get_engine()
    sc = Scratch.get(device)
    sums = torch.zeros(4, dtype=torch.float32, device=device)   # rmse_global, l1_global, rmse_local, l1_local
    slot = torch.empty(4, 4, dtype=torch.float32, device=device)
    batches = 0
    for ground, mask, _ in DevicePrefetcher(loader, device):
        ground, mask = ground.contiguous().float(), mask.contiguous().float()
        n = ground.numel()
        m, masked = torch.empty_like(mask), torch.empty_like(ground)
        call("b200_mask_prepare", ptr(ground), ptr(mask), n, 1, ptr(m), ptr(masked), stream())
        if is_flip_mask:  # evaluate.py:131-132
            m = 1.0 - m
            call("b200_mask_prepare", ptr(ground), ptr(m), n, 0, None, ptr(masked), stream())
        # the reference never calls .eval(): in train mode every batch draws its own dropout masks
        seed = net._next_seed() if (net.training and hasattr(net, "_next_seed")) else 0
        out = eng.forward(masked, net.training, comp_base=masked, comp_mul=m, seed=seed, keep=False).comp
        for i, kind in enumerate((RECON_RMSE, RECON_L1, RECON_LOCAL_MSE, RECON_LOCAL_L1)):
            call("b200_recon_loss", kind, ptr(ground), ptr(out), ptr(m) if kind >= RECON_LOCAL_L1 else None, n, float(eps), ptr(sc.recon), ptr(slot[i]), stream())
        vals = slot[:, 0].clone()
        vals[2] = torch.sqrt(vals[2] + eps)
        sums += vals
        batches += 1
    if batches == 0:
        raise ValueError("calculate_metric: empty loader")
    r = (sums / batches).tolist()  # the one device -> host read
    return {"recon_rmse_global": r[0], "recon_l1_global": r[1], "recon_rmse_local": r[2], "recon_l1_local": r[3], "face_parsing_metric": None,
            "fid": -1, "epoch": epoch}

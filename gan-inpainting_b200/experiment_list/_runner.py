"""Shared epoch loop behind the ``begin(state, loaders)`` recipes (call surface of the reference's experiment scripts).

The reference keeps one ~270-line copy of this loop per recipe (experiment_list/wgan_l1.py:15-277,
minimaxgan_rmse.py:15-260, ...).  Here the per-batch body is ``InpaintingTrainer.step`` (one fused stream of sm_100a
kernels) and the bookkeeping around it is written once:

* ``state`` keys read: ``numepoch, batchsize, imagedim, saveevery, evalevery, updatediscevery, generator`` (train.py:22-36);
  optional extras with defaults so reference-style callers keep working: ``experiment_dir`` (the reference hard-codes a
  home directory, wgan_l1.py:27), ``log_dir`` ('log'), ``distributed`` (False), ``bucket_mb`` (25), ``evaluate_fn``
  (callable(net_G, loader, mode) -> dict; default: ``lib.models.evaluate.calculate_metric`` like wgan_l1.py:251-259, with
  ``state['train_fid'] / ['test_fid'] / ['inception_model']`` passed through when present), ``graph`` (True: replay the step as a
  CUDA graph), ``max_batches`` (None; smoke runs).
* per-epoch history dicts with the reference's keys (``g``/``d``/``gradients``; -7777 sentinel when G never updated,
  wgan_l1.py:214-249), pickled to ``training_epoch_history.obj`` / ``eval_history.obj`` (:262-267), and
  ``epoch{N}_G.pt`` checkpoints of ``net_G.state_dict()`` (:269-271) in the reference's key layout.
* the reference reads ~8 scalars and 34 per-parameter gradient means with one ``.item()`` sync each per batch
  (SURVEY.md a14); here each batch does ONE device->host copy of the loss vector and one of the gradient means.
"""
from __future__ import annotations

import logging
import os
import pickle
import time
from typing import Callable, Dict, Optional

import torch

from ..data import DevicePrefetcher
from ..lib import pytorch_ssim
from ..lib.models import evaluate, networks
from ..trainer import InpaintingTrainer

SENTINEL = -7777  # wgan_l1.py:224-233


def wgan_g_schedule(update_g_every: int = 5) -> Callable[[int, int], bool]:
    """wgan_l1.py:155-163: K = 140 while G_iter < 25 or G_iter % 500 == 0, else ``update_g_every``; G is updated when
    batch_idx % K == 0 and batch_idx > 0."""

    def should_update(batch_idx: int, g_iter: int) -> bool:
        k = 140 if (g_iter < 25 or g_iter % 500 == 0) else update_g_every
        return batch_idx % k == 0 and batch_idx > 0

    return should_update


def every_batch(batch_idx: int, g_iter: int) -> bool:
    """minimaxgan_rmse.py:150-173: the generator is updated on every batch."""
    return True


def _logger(title: str, log_dir: str, to_file: bool = True) -> logging.Logger:
    logger = logging.getLogger(title)
    if not to_file:  # data-parallel ranks other than 0
        if not logger.handlers:
            logger.addHandler(logging.NullHandler())
        return logger
    os.makedirs(log_dir, exist_ok=True)
    if not logger.handlers:
        hdlr = logging.FileHandler(os.path.join(log_dir, "{}.log".format(title)))
        hdlr.setFormatter(logging.Formatter("%(asctime)s %(message)s"))
        logger.addHandler(hdlr)
    logger.setLevel(logging.INFO)
    return logger


def run(state: dict, loaders: dict, title: str, recipe: str, recon: str, schedule: Callable[[int, int], bool], local_d: bool = False,
        lambdas=(300.0, 300.0), loss_terms: Optional[dict] = None, log_ssim: Optional[bool] = None) -> Dict[str, object]:
    """``loss_terms``: extra generator loss terms for ``InpaintingTrainer`` (``tv_weight``, ``feature_weights``, ``local_recon``).
    ``log_ssim``: per-epoch mean SSIM of the inpainted batches (experiment1_global_local_D.py:209, experiment3_global_D.py:201;
    default: the recipes with a local discriminator)."""
    log_ssim = local_d if log_ssim is None else bool(log_ssim)
    state = state.copy()
    state.update({"title": title})
    train_loader, test_loader = loaders["train"], loaders.get("test")
    experiment_dir = state.get("experiment_dir") or os.path.join("model", title)
    os.makedirs(experiment_dir, exist_ok=True)
    if not torch.cuda.is_available():
        raise RuntimeError("the B200 train step has no CPU path (the reference falls back to CPU at wgan_l1.py:50-52; this build does not)")
    distributed = bool(state.get("distributed", False))
    # data parallel (one process per GPU): every rank trains, only rank 0 writes checkpoints / history pickles / the log file and
    # runs the evaluation pass; the others wait at a barrier so that nobody races ahead of a half-written checkpoint
    rank0 = True
    if distributed:
        import torch.distributed as dist

        rank0 = dist.get_rank() == 0
    logger = _logger(title, state.get("log_dir", "log"), to_file=rank0)
    device = torch.device("cuda", torch.cuda.current_device())
    logger.info("using device %s", device)
    R = int(state.get("imagedim", 128))
    num_epochs, save_every, evaluate_every = int(state["numepoch"]), int(state["saveevery"]), int(state["evalevery"])

    net_G = networks.get_network("generator", state.get("generator", "unet")).to(device)
    sigmoid = recipe != "wgan"
    net_D = networks.PatchGANDiscriminator(sigmoid=sigmoid, imagedim=R).to(device)
    net_Dl = networks.PatchGANDiscriminator(sigmoid=sigmoid, imagedim=R).to(device) if local_d else None
    tr = InpaintingTrainer(recipe, net_G, net_D, net_Dl, recon=recon, lambdas=lambdas, distributed=distributed,
                           bucket_mb=float(state.get("bucket_mb", 25.0)), log_grad_means=True, graph=bool(state.get("graph", True)), **(loss_terms or {}))
    g_names = [n for n, _ in net_G.named_parameters() if "bias" not in n]
    d_names = [n for n, _ in net_D.named_parameters() if "bias" not in n]
    evaluate_fn: Optional[Callable] = state.get("evaluate_fn")
    if evaluate_fn is None:  # wgan_l1.py:251-252
        def evaluate_fn(net, loader, mode):
            return evaluate.calculate_metric(device, loader, net, state.get(mode + "_fid", (-1, -1)), mode=mode, inception_model=state.get("inception_model"))
    max_batches = state.get("max_batches")

    training_epoch_hist, eval_hist = [], []
    g_iter = d_iter = 0
    for epoch in range(num_epochs + 1):
        start = time.time()
        eg = {"total": 0.0, "recon": 0.0, "adv": 0.0, "update_count": 0}
        if loss_terms:  # wgan_perceptual_style_faceparsing.py logs every term of g_loss
            eg.update({"recon_local": 0.0, "tv": 0.0, "perceptual": 0.0, "style": 0.0})
        ed = {"total": 0.0, "adv_real": 0.0, "adv_fake": 0.0, "update_count": 0}
        gh = {"avg_g": {n: 0.0 for n in g_names}, "avg_d": {n: 0.0 for n in d_names}}
        ssim_acc, n_images = torch.zeros((), dtype=torch.float32, device=device), 0  # experiment1_global_local_D.py:135,209
        net_G.train()
        # the next batch's host -> device copy overlaps this batch's kernels (the reference blocks on .to(device), wgan_l1.py:113-114)
        for batch_idx, (ground, mask, _) in enumerate(DevicePrefetcher(train_loader, device)):
            if max_batches is not None and batch_idx >= max_batches:
                break
            upd = schedule(batch_idx, g_iter)
            tr.step(ground, mask, update_g=upd)
            s = tr.read_scalars()
            d_iter += 1
            if log_ssim:  # SSIM of the inpainted batch against the ground truth, logged per epoch (experiment1_global_local_D.py:209,216)
                ssim_acc += pytorch_ssim.ssim(ground, tr.last["inpainted"]) * ground.shape[0]
                n_images += ground.shape[0]
            if upd:
                g_iter += 1
                eg["total"] += s["g_loss"]
                eg["recon"] += s["recon"]
                eg["adv"] += s["g_adv"] + (s["g_adv_local"] if local_d else 0.0)
                eg["update_count"] += 1
                if loss_terms:
                    eg["recon_local"] += s["recon_local"]
                    eg["tv"] += s["g_tv"]
                    eg["perceptual"] += s["g_perceptual"]
                    eg["style"] += s["g_style"]
                for n, v in tr.grad_means_host("g").items():
                    if n in gh["avg_g"]:
                        gh["avg_g"][n] += v
                logger.info("[epoch %d/%d][batch %d/%d]\tLoss_D: %.4f\tLoss_G: %.4f" % (epoch, num_epochs, batch_idx, len(train_loader), s["d_loss"], s["g_adv"]))
            ed["total"] += s["d_loss"]
            ed["adv_real"] += s["d_loss_real"]
            ed["adv_fake"] += s["d_loss_fake"]
            ed["update_count"] += 1
            for n, v in tr.grad_means_host("d").items():
                if n in gh["avg_d"]:
                    gh["avg_d"][n] += v
        if eg["update_count"]:
            for k in [k for k in eg if k != "update_count"]:
                eg[k] /= eg["update_count"]
            for n in g_names:
                gh["avg_g"][n] /= eg["update_count"]
        else:  # generator never updated in this epoch
            for k in [k for k in eg if k != "update_count"]:
                eg[k] = SENTINEL
            for n in g_names:
                gh["avg_g"][n] = SENTINEL
        if ed["update_count"]:
            for k in ("total", "adv_real", "adv_fake"):
                ed[k] /= ed["update_count"]
            for n in d_names:
                gh["avg_d"][n] /= ed["update_count"]
        if log_ssim and n_images:
            eg["ssim"] = float(ssim_acc) / n_images
        training_epoch_hist.append({"g": eg, "d": ed, "gradients": gh})

        if epoch % evaluate_every == 0 and epoch > 0 and rank0:
            entry = {"train": evaluate_fn(net_G, train_loader, "train"), "test": evaluate_fn(net_G, test_loader, "test") if test_loader is not None else None}
            eval_hist.append(entry)
            logger.info("VALIDATION: train - {}, test - {}".format(str(entry["train"]), str(entry["test"])))
            with open(os.path.join(experiment_dir, "training_epoch_history.obj"), "wb") as handle:
                pickle.dump(training_epoch_hist, handle, protocol=pickle.HIGHEST_PROTOCOL)
            with open(os.path.join(experiment_dir, "eval_history.obj"), "wb") as handle:
                pickle.dump(eval_hist, handle, protocol=pickle.HIGHEST_PROTOCOL)
        if epoch % save_every == 0 and epoch > 0 and rank0:
            try:
                torch.save(net_G.state_dict(), os.path.join(experiment_dir, "epoch{}_G.pt".format(epoch)))
            except Exception:  # the reference logs and carries on (wgan_l1.py:270-274)
                logger.exception("checkpoint failed")
        if distributed and (epoch % evaluate_every == 0 or epoch % save_every == 0) and epoch > 0:
            dist.barrier()
        logger.info("epoch %d done in %.1fs" % (epoch, time.time() - start))
    if distributed:
        tr.release_graphs()  # captured NCCL collectives pin the communicator: must go before the caller's destroy_process_group()
    return {"net_G": net_G, "net_D": net_D, "net_D_local": net_Dl, "trainer": tr, "history": training_epoch_hist, "eval_history": eval_hist,
            "g_iter": g_iter, "d_iter": d_iter}

"""The per-batch train step of the reference's experiment scripts, as one stream of CUDA kernels.

Four recipes:

* ``wgan``         experiment_list/wgan_l1.py:110-186 (recon 'l1') / wgan_rmse.py (recon 'rmse'):
                   ceil(mask) -> mask out -> G -> composite -> D step (mean heads, seeds +1/-1, RMSprop, clip
                   to [-0.01, 0.01]) -> every K-th batch the G step (frozen D, adv + recon, RMSprop).
* ``minimax``      experiment_list/minimaxgan_rmse.py:110-173 (recon 'rmse') / minimaxgan_l1.py (recon 'l1'):
                   same prologue, BCE heads on sigmoid outputs, Adam(2e-4, (0.5, 0.999)), G step every batch.
* ``global_local`` experiment_list/experiment1_global_local_D.py:139-200: G step first (no composite, mask not
                   ceil'ed, LSGAN heads of a global and a local discriminator, 300*RMSE + 300*masked RMSE), then one
                   D step over both discriminators with a single Adam.
* ``global_d``     experiment_list/experiment3_global_D.py:140-196: the same order with the global discriminator only, BCE heads
                   and 300*RMSE.

Unlike the module API (``lib.models.networks`` + autograd) this path uses no autograd at all: the engines' forward
and backward are called directly, gradients land in the flat gradient buffers of ``optim.flatten_parameters``, the
optimizers are single fused launches, and the logged scalars are written by the kernels into ONE device vector that
is read back with one copy per step (instead of the reference's ~8 + 34 ``.item()`` syncs).

Data parallelism (``world_size > 1``, one process per GPU): gradients are summed with NCCL all-reduces over
slices ("buckets") of the flat gradient buffer, issued on a side stream as soon as the backward pass has moved
past a bucket, so the transfer overlaps the remaining backward kernels.  BatchNorm statistics stay per replica
(stock DistributedDataParallel semantics).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, optim
from ._lib import call, ptr, stream
from .engine import Scratch

ADV_WGAN, ADV_BCE, ADV_LSGAN = 0, 1, 2
RECON_L1, RECON_RMSE, RECON_LOCAL_L1, RECON_LOCAL_MSE, RECON_MASKED_RMSE = 0, 1, 2, 3, 4

# slots of the per-step scalar vector
SLOTS = ["d_loss_real", "d_loss_fake", "g_adv", "g_adv_local", "d_loss_real_local", "d_loss_fake_local"]
RECON_SLOT = 8      # 4 floats: loss, denom, exact count (2 words)
RECON2_SLOT = 12    # second reconstruction term (global_local; config 4's local term)
TV_SLOT, PERCEPTUAL_SLOT, STYLE_SLOT = 16, 17, 18  # config 4's extra generator loss terms (already weighted)
NSCALARS = 20


class GradBucketer:
    """Bucketed, overlapped gradient all-reduce over a flat gradient buffer whose tensors complete in descending
    address order during backward (decoder outermost..innermost, then encoder innermost..outermost)."""

    def __init__(self, group: optim.FlatGroup, bucket_mb: float = 25.0, process_group=None, comm_sms: int = 0, wire_bf16: bool = False):
        import torch.distributed as dist

        self.dist = dist
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.group = group
        # SMs left to the collective while buckets are in flight: the persistent GEMMs (one CTA per SM, static tile schedule) are
        # launched on 148 - comm_sms SMs from the first bucket until finish(), so that NCCL's CTAs never hold an SM a persistent
        # CTA was counting on (a late-starting CTA stretches its whole launch — the 0.4-0.7 ms a step lost at 2-8 GPUs in round 1)
        self.comm_sms = int(comm_sms)
        # wire_bf16: the buckets travel as bf16 (cast kernels on the communication stream before / after the collective): half the
        # NVLink bytes and half the time NCCL's CTAs compete with the backward GEMMs; the averaged gradient then carries one bf16
        # rounding (rel 2^-9 per element) — the same storage precision as every activation gradient of the step
        self.wire = torch.empty(group.numel, dtype=torch.bfloat16, device=group.grads.device) if (wire_bf16 and group.grads.is_cuda) else None
        self.bucket_elems = max(1, int(bucket_mb * (1 << 20) / 4))
        self.cuda = group.grads.is_cuda
        self.comm = torch.cuda.Stream() if self.cuda else None
        # NCCL averages inside the collective; other backends (the gloo CPU tests) sum and scale afterwards
        self.avg = dist.get_backend(process_group) == "nccl"
        self.off_of = {id(p): o for p, o in zip(group.members, group.offsets)}
        self.reset()

    def reset(self):
        self.hi = self.group.numel      # everything in [hi, numel) has been handed to the collective
        self.lo = self.group.numel      # everything in [lo, numel) is complete on the compute stream
        self.launched = 0

    def ready(self, p):
        """Called right after the kernels producing ``p.grad`` were enqueued."""
        self.lo = min(self.lo, self.off_of[id(p)])
        if self.hi - self.lo >= self.bucket_elems:
            self._launch(self.lo)

    def _reduce(self, t):
        if self.avg:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.AVG, group=self.pg)
        else:
            self.dist.all_reduce(t, group=self.pg)
            t.mul_(1.0 / self.world)

    def _launch(self, lo):
        if lo >= self.hi:
            return
        bucket = self.group.grads[lo : self.hi]
        if self.cuda and self.comm_sms > 0 and self.launched == 0:
            _lib.value("b200_set_sm_limit", 148 - self.comm_sms)
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                if self.wire is not None:
                    w = self.wire[lo : self.hi]
                    cs = ctypes_ptr(self.group.grads, lo), ctypes_ptr(self.wire, lo)
                    call("b200_cast_grads", cs[0], cs[1], self.hi - lo, 1, stream())
                    self._reduce(w)
                    call("b200_cast_grads", cs[1], cs[0], self.hi - lo, 0, stream())
                else:
                    self._reduce(bucket)
        else:
            self._reduce(bucket)
        self.hi = lo
        self.launched += 1

    def finish(self):
        """Flush the remaining range and make the compute stream wait for all reductions; grads hold the MEAN over ranks."""
        self._launch(0)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm)
            if self.comm_sms > 0:
                _lib.value("b200_set_sm_limit", 0)
        n = self.launched
        self.reset()
        return n


def nccl_group_options(max_ctas: Optional[int] = None):
    """``pg_options`` for ``torch.distributed.init_process_group("nccl", ...)`` that cap the thread blocks NCCL may use per
    collective (``B200_NCCL_CTAS`` when ``max_ctas`` is None; 0 / unset = NCCL's default).

    Why: the persistent tcgen05 GEMMs own all 227 KB of shared memory of every SM they run on, so an all-reduce block and a GEMM
    block never share an SM — whichever is resident first keeps it.  With a static tile schedule a GEMM whose block has to wait for
    an all-reduce block (100+ us per 25 MB bucket) is stretched by that wait.  Capping NCCL at k blocks and keeping k SMs out of the
    GEMM grids (``comm_sms=k`` -> ``b200_set_sm_limit``) gives each side its own SMs; the 178 MB of gradients per step need only
    ~30 GB/s to stay hidden behind the backward pass."""
    import os

    import torch.distributed as dist

    if max_ctas is None:
        max_ctas = int(os.environ.get("B200_NCCL_CTAS", "0"))
    if max_ctas <= 0:
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = int(max_ctas)
    opts.config.min_ctas = 1
    return opts


class InpaintingTrainer:
    def __init__(self, recipe: str, net_G, net_D, net_D_local=None, recon: str = "l1", lambdas=(300.0, 300.0), clip: float = 0.01,
                 distributed: bool = False, bucket_mb: float = 25.0, log_grad_means: bool = False, graph: bool = False, rank: Optional[int] = None,
                 comm_sms: Optional[int] = None, bucket_mb_d: float = 4.0, wire_bf16: Optional[bool] = None, tv_weight: float = 0.0,
                 feature_weights: Optional[Tuple[float, float]] = None, local_recon: Optional[str] = None):
        """``tv_weight`` / ``feature_weights=(weight_p, weight_s)`` / ``local_recon`` ("l1" | "mse") add the generator loss terms of
        wgan_perceptual_style_faceparsing.py:208-221 to the ``wgan`` / ``minimax`` recipes: ``tv_loss(inpainted, tv_weight)`` (with
        its gradient), the VGG19 perceptual + style losses (numbers only: the reference computes them under ``no_grad``, SURVEY.md
        D11) and ``LocalLoss(base)(ground, inpainted, mask)``.  That script's face-parsing term is not built."""
        if recipe not in ("wgan", "minimax", "global_local", "global_d"):
            raise ValueError(recipe)
        if recipe in ("global_local", "global_d") and (tv_weight or feature_weights or local_recon):
            raise ValueError("the extra generator loss terms belong to the wgan / minimax recipes")
        if recipe == "global_d" and net_D_local is not None:
            raise ValueError("global_d (experiment3_global_D.py) has no local discriminator")
        self.tv_weight = float(tv_weight)
        self.feature_weights = tuple(float(w) for w in feature_weights) if feature_weights is not None else None
        self.local_recon = {None: -1, "l1": RECON_LOCAL_L1, "mse": RECON_LOCAL_MSE}[local_recon]
        if recipe == "global_local" and net_D_local is None:
            raise ValueError("global_local needs a local discriminator")
        self.recipe, self.recon = recipe, {"l1": RECON_L1, "rmse": RECON_RMSE}[recon]
        self.G, self.D, self.Dl = net_G, net_D, net_D_local
        self.lambdas, self.clip = lambdas, clip
        self.eG, self.eD = net_G._get_engine(), net_D._get_engine()
        self.eDl = net_D_local._get_engine() if net_D_local is not None else None
        self.fG = optim.flatten_parameters(net_G)
        self.fD = optim.flatten_parameters(net_D)
        self.fDl = optim.flatten_parameters(net_D_local) if net_D_local is not None else None
        dev = self.fG.params.device
        self.device = dev
        if recipe == "wgan":  # wgan_l1.py:64-65 + :152-153
            self.optG = optim.RMSprop(net_G.parameters(), lr=0.00005)
            self.optD = optim.RMSprop(net_D.parameters(), lr=0.00005, clip=(-clip, clip))
            self.sigmoid, self.adv = False, ADV_WGAN
        elif recipe == "minimax":  # minimaxgan_rmse.py:64-65
            self.optG = optim.Adam(net_G.parameters(), lr=0.0002, betas=(0.5, 0.999))
            self.optD = optim.Adam(net_D.parameters(), lr=0.0002, betas=(0.5, 0.999))
            self.sigmoid, self.adv = True, ADV_BCE
        elif recipe == "global_d":  # experiment3_global_D.py:118-122: BCE heads, one Adam per network, G step first
            self.optG = optim.Adam(net_G.parameters(), lr=0.0002, betas=(0.5, 0.999))
            self.optD = optim.Adam(net_D.parameters(), lr=0.0002, betas=(0.5, 0.999))
            self.sigmoid, self.adv = True, ADV_BCE
        else:  # experiment1_global_local_D.py:122-123
            self.optG = optim.Adam(net_G.parameters(), lr=0.0002, betas=(0.5, 0.999))
            self.optD = optim.Adam(list(net_D_local.parameters()) + list(net_D.parameters()), lr=0.0002, betas=(0.5, 0.999))
            self.sigmoid, self.adv = True, ADV_LSGAN
        self.scalars = torch.zeros(NSCALARS, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.world = 1
        self.rank = 0
        self.buckets: Dict[int, GradBucketer] = {}
        if distributed:
            import torch.distributed as dist

            import os

            self.world = dist.get_world_size()
            self.rank = dist.get_rank()
            if comm_sms is None:
                comm_sms = int(os.environ.get("B200_COMM_SMS", "0"))
            if wire_bf16 is None:
                wire_bf16 = os.environ.get("B200_GRAD_WIRE", "fp32").lower() == "bf16"
            self.wire_bf16 = bool(wire_bf16)
            for fg in (self.fG, self.fD, self.fDl):
                if fg is not None:
                    dist.broadcast(fg.params, src=0)
                    # the discriminators' 11 MB of gradients use small buckets: their all-reduce then overlaps the rest of the D
                    # backward pass instead of starting after it
                    self.buckets[id(fg)] = GradBucketer(fg, bucket_mb if fg is self.fG else min(bucket_mb, bucket_mb_d), comm_sms=comm_sms,
                                                        wire_bf16=self.wire_bf16)
            for net in (net_G, net_D, net_D_local):
                if net is not None:
                    for b in net.buffers():
                        dist.broadcast(b, src=0)
            _lib.WEIGHT_EPOCH += 1
        self.log_grad_means = log_grad_means
        self.grad_means: Dict[str, torch.Tensor] = {}
        self._abs = {}
        # graph=True: after one eager step per (batch shape, update_g) the whole step — ~210 kernels, the NCCL bucket all-reduces on
        # their side stream included — is captured into a CUDA graph and replayed with ONE host call, so the step no longer depends
        # on the Python thread keeping ahead of the GPU (3.8 ms of enqueue work per 7.7 ms step; 8 ranks share the host's cores)
        self.graph = graph
        self._graphs: Dict[tuple, "_CapturedStep"] = {}
        self._seed_ctr = torch.zeros(1, dtype=torch.int32, device=dev)  # device-side dropout seed counter (b200_set_seed_counter)
        self._seed_ctr_host = 0
        if rank is not None:
            self.rank = int(rank)
        # every data-parallel replica draws its own dropout masks (stock DDP: one RNG stream per rank): the rank enters the seed
        # of every mask; rank 0 keeps the single-GPU stream
        self._seed_base = (self.rank * 0x632BE5AB) & 0x3FFFFFF

    # ------------------------------------------------------------------------------------------ helpers
    def _slot(self, name_or_index):
        i = SLOTS.index(name_or_index) if isinstance(name_or_index, str) else name_or_index
        return ctypes_ptr(self.scalars, i)

    def _adv(self, ctx, target, scale, slot):
        """Adversarial head on a discriminator context -> gradient w.r.t. the pre-sigmoid output."""
        dpred = torch.empty((ctx.B,), dtype=torch.float32, device=self.device)
        call("b200_adv_loss", ptr(ctx.pred), ctx.B, self.adv, float(target), float(scale), self._slot(slot), ptr(dpred), stream())
        if not self.sigmoid:
            return dpred
        dlin = torch.empty_like(dpred)
        call("b200_head_dlin", ptr(dpred), ptr(ctx.pred), 1, ctx.B, ptr(dlin), stream())
        return dlin

    def _reduce(self, fg, hooked: bool):
        if self.world == 1:
            return
        b = self.buckets[id(fg)]
        if not hooked:
            b.lo = 0
        b.finish()

    def _grad_log(self, tag, fg):
        if not self.log_grad_means:
            return
        if tag not in self._abs:
            names, spans = fg.logging_offsets()
            # one block per tensor over [start, end): encode as consecutive (start, end) pairs with zero-length gaps
            flat = []
            for a, b in spans:
                flat += [a, b]
            self._abs[tag] = (names, torch.tensor(flat, dtype=torch.int64, device=self.device), torch.zeros(2 * len(spans), dtype=torch.float32, device=self.device))
        names, offs, out = self._abs[tag]
        call("b200_absmean", ptr(fg.grads), ptr(offs), 2 * len(names) - 1, ptr(out), stream())
        self.grad_means[tag] = (names, out)

    def grad_means_host(self, tag):
        names, out = self.grad_means[tag]
        vals = out.tolist()
        return {n: vals[2 * i] for i, n in enumerate(names)}

    # ------------------------------------------------------------------------------------------ the step
    @torch.no_grad()
    def step(self, ground: torch.Tensor, mask: torch.Tensor, update_g: bool = True, seed: Optional[int] = None):
        """One batch.  ``ground`` / ``mask``: fp32 CUDA [B,1,R,R].  Returns the device scalar vector (see ``read_scalars``)."""
        if not (ground.is_cuda and mask.is_cuda):
            raise _lib.B200Error("the train step runs on CUDA tensors only")
        if self.graph and seed is None:
            return self._step_graphed(ground, mask, bool(update_g))
        self.step_count += 1
        seed = self.step_count + self._seed_base if seed is None else seed
        if self.recipe in ("global_local", "global_d"):
            return self._step_global_local(ground.contiguous().float(), mask.contiguous().float(), seed)
        return self._step_gan(ground.contiguous().float(), mask.contiguous().float(), update_g, seed)

    def _run(self, ground, mask, update_g, seed):
        if self.recipe in ("global_local", "global_d"):
            return self._step_global_local(ground, mask, seed)
        return self._step_gan(ground, mask, update_g, seed)

    def _step_graphed(self, ground, mask, update_g):
        key = (tuple(ground.shape), ground.dtype, mask.dtype, update_g)
        cs = self._graphs.get(key)
        if cs is None:
            # first batch of this shape: a real eager step (allocates every scratch buffer / pack table at its final size), then
            # the capture — which only records — for all later batches
            self.step_count += 1
            out = self._run(ground.contiguous().float(), mask.contiguous().float(), update_g, self.step_count + self._seed_base)
            self._graphs[key] = _CapturedStep(self, ground, mask, update_g)
            return out
        self.step_count += 1
        for e in (self.eG, self.eD, self.eDl):  # weights rewritten outside the graph (load_state_dict, a user's optimizer)?
            if e is not None:
                e.refresh_weights()
        if self._seed_ctr_host != self.step_count - 1:  # eager steps ran in between: realign the device counter
            self._seed_ctr.fill_(self.step_count - 1)
        self._seed_ctr_host = self.step_count
        cs.replay(ground, mask)
        self.last = cs.last
        return self.scalars

    def release_graphs(self):
        """Destroy the captured step graphs.  MUST run before ``torch.distributed.destroy_process_group()``: NCCL keeps a
        communicator alive (and its destruction waits) while CUDA graphs holding captured collectives exist."""
        import gc

        torch.cuda.synchronize(self.device)
        for cs in self._graphs.values():
            cs.graph = None
            cs.last = None
        self._graphs.clear()
        self.last = None
        gc.collect()
        torch.cuda.synchronize(self.device)

    def _step_gan(self, ground, mask, update_g, seed):
        G, D, eG, eD, dev = self.G, self.D, self.eG, self.eD, self.device
        n = ground.numel()
        B = ground.shape[0]
        sc = Scratch.get(dev)
        # ---- prologue: wgan_l1.py:113-122
        mask_c = torch.empty_like(mask)
        masked = torch.empty_like(ground)
        call("b200_mask_prepare", ptr(ground), ptr(mask), n, 1, ptr(mask_c), ptr(masked), stream())
        cG = eG.forward(masked, True, comp_base=masked, comp_mul=mask_c, seed=seed)
        inpainted = cG.comp
        # ---- D step: :128-153
        gD = self.fD.grad_map()
        hook = self.buckets.get(id(self.fD))
        cR = eD.forward(ground, True, self.sigmoid)
        if self.recipe == "wgan":
            cF = eD.forward(inpainted, True, self.sigmoid)
            eD.backward(cR, self._adv(cR, 1.0, 1.0, "d_loss_real"), gD, accumulate=False)
            eD.backward(cF, self._adv(cF, 0.0, -1.0, "d_loss_fake"), gD, accumulate=True, on_ready=hook.ready if hook else None)
        else:  # minimaxgan_rmse.py:129-148: forward/backward real, then forward/backward fake
            eD.backward(cR, self._adv(cR, 1.0, 1.0, "d_loss_real"), gD, accumulate=False)
            cF = eD.forward(inpainted, True, self.sigmoid)
            eD.backward(cF, self._adv(cF, 0.0, 1.0, "d_loss_fake"), gD, accumulate=True, on_ready=hook.ready if hook else None)
        if update_g:
            # loss values of the G step that depend on neither network's update: enqueued here, so that on N > 1 GPUs they run while
            # the last discriminator bucket is still being all-reduced (the optimizer step below waits for it)
            call("b200_recon_loss", self.recon, ptr(ground), ptr(inpainted), None, n, 1e-16, ptr(sc.recon), self._slot(RECON_SLOT), stream())
            if self.local_recon >= 0:  # wgan_perceptual_style_faceparsing.py:209
                call("b200_recon_loss", self.local_recon, ptr(ground), ptr(inpainted), ptr(mask_c), n, 1e-16, ptr(sc.recon), self._slot(RECON2_SLOT), stream())
            if self.tv_weight:  # :219 — the value; its gradient is added to d g_loss / d inpainted below
                part = torch.empty(2 * B * ground.shape[-2], dtype=torch.float32, device=dev)
                call("b200_tv_loss", ptr(inpainted), B, ground.shape[-2], ground.shape[-1], self.tv_weight, ptr(part), self._slot(TV_SLOT), stream())
            if self.feature_weights is not None:  # :216 — no gradient
                from .lib.models import loss as _loss

                _loss._feature_losses(inpainted, ground, True, True, dst=self.scalars[PERCEPTUAL_SLOT:STYLE_SLOT + 1], weights=self.feature_weights)
        self._reduce(self.fD, hooked=True)
        self._grad_log("d", self.fD)
        self.optD.step()
        eD.refresh_weights(force=True)
        # ---- G step: :169-186
        if update_g:
            cA = eD.forward(inpainted, True, self.sigmoid)
            dadv = eD.backward(cA, self._adv(cA, 1.0, 1.0, "g_adv"), None, input_grad=True)
            if self.tv_weight:
                call("b200_tv_loss_bwd", ptr(inpainted), B, ground.shape[-2], ground.shape[-1], self.tv_weight, None, ptr(dadv), 1, stream())
            dpre = torch.empty_like(ground)
            gG = self.fG.grad_map()
            bias = eG.uc[0].bias
            call("b200_gstep_seed", ptr(ground), ptr(inpainted), ptr(cG.g_out), ptr(mask_c), ptr(dadv), None, self.recon, 1.0, self._slot(RECON_SLOT),
                 self.local_recon, 1.0 if self.local_recon >= 0 else 0.0, self._slot(RECON2_SLOT) if self.local_recon >= 0 else None, None, 1, n, ptr(dpre),
                 None, ptr(sc.recon), ptr(gG[id(bias)]), 0, stream())
            hook = self.buckets.get(id(self.fG))
            if hook:
                hook.ready(bias)
            eG.backward(cG, dpre, gG, accumulate=False, on_ready=hook.ready if hook else None)
            self._reduce(self.fG, hooked=True)
            self._grad_log("g", self.fG)
            self.optG.step()
            eG.refresh_weights(force=True)
        self.last = dict(inpainted=inpainted, g_out=cG.g_out, mask=mask_c, ctx=dict(G=cG, Dreal=cR, Dfake=cF, Dgen=cA if update_g else None))
        return self.scalars

    def _step_global_local(self, ground, mask, seed):
        """experiment1_global_local_D.py:139-200 (global + local discriminator, LSGAN heads) and, without a local discriminator,
        experiment3_global_D.py:140-196 (``global_d``: BCE heads, lambda1 * RMSE only)."""
        eG, eD, eDl, dev = self.eG, self.eD, self.eDl, self.device
        local = eDl is not None
        n = ground.numel()
        sc = Scratch.get(dev)
        l1, l2 = self.lambdas
        masked = torch.empty_like(ground)
        call("b200_mask_prepare", ptr(ground), ptr(mask), n, 0, None, ptr(masked), stream())  # mask is NOT ceil'ed (:142)
        # ---- G step (:150-172), discriminators frozen
        cG = eG.forward(masked, True, seed=seed)
        inp = cG.g_out
        cg = eD.forward(inp, True, True)
        cl = eDl.forward(inp, True, True, xmul=mask) if local else None
        dg = eD.backward(cg, self._adv(cg, 1.0, 1.0, "g_adv"), None, input_grad=True)
        dl = eDl.backward(cl, self._adv(cl, 1.0, 1.0, "g_adv_local"), None, input_grad=True) if local else None
        call("b200_recon_loss", RECON_RMSE, ptr(ground), ptr(inp), None, n, 1e-16, ptr(sc.recon), self._slot(RECON_SLOT), stream())
        if local:
            call("b200_recon_loss", RECON_MASKED_RMSE, ptr(ground), ptr(inp), ptr(mask), n, 1e-16, ptr(sc.recon), self._slot(RECON2_SLOT), stream())
        dpre = torch.empty_like(ground)
        gG = self.fG.grad_map()
        bias = eG.uc[0].bias
        call("b200_gstep_seed", ptr(ground), ptr(inp), ptr(cG.g_out), ptr(mask), ptr(dg), ptr(dl) if local else None, RECON_RMSE, float(l1), self._slot(RECON_SLOT),
             RECON_MASKED_RMSE if local else -1, float(l2) if local else 0.0, self._slot(RECON2_SLOT) if local else None, None, 0, n, ptr(dpre), None, ptr(sc.recon),
             ptr(gG[id(bias)]), 0, stream())
        hook = self.buckets.get(id(self.fG))
        if hook:
            hook.ready(bias)
        eG.backward(cG, dpre, gG, accumulate=False, on_ready=hook.ready if hook else None)
        self._reduce(self.fG, hooked=True)
        self._grad_log("g", self.fG)
        self.optG.step()
        eG.refresh_weights(force=True)
        # ---- D step (:178-200): two forwards per discriminator, one summed loss, one Adam (over chain(Dl, Dg) when there are two)
        gD = self.fD.grad_map()
        c1 = eD.forward(ground, True, True)
        c2 = eD.forward(inp, True, True)
        hg = self.buckets.get(id(self.fD))
        c3 = c4 = None
        if local:
            gDl = self.fDl.grad_map()
            c3 = eDl.forward(ground, True, True, xmul=mask)
            c4 = eDl.forward(inp, True, True, xmul=mask)
            hl = self.buckets.get(id(self.fDl))
            eDl.backward(c3, self._adv(c3, 1.0, 1.0, "d_loss_real_local"), gDl, accumulate=False)
            eDl.backward(c4, self._adv(c4, 0.0, 1.0, "d_loss_fake_local"), gDl, accumulate=True, on_ready=hl.ready if hl else None)
        eD.backward(c1, self._adv(c1, 1.0, 1.0, "d_loss_real"), gD, accumulate=False)
        eD.backward(c2, self._adv(c2, 0.0, 1.0, "d_loss_fake"), gD, accumulate=True, on_ready=hg.ready if hg else None)
        if local:
            self._reduce(self.fDl, hooked=True)
        self._reduce(self.fD, hooked=True)
        self._grad_log("d", self.fD)
        if local:
            self._grad_log("d_local", self.fDl)
        self.optD.step()
        eD.refresh_weights(force=True)
        if local:
            eDl.refresh_weights(force=True)
        self.last = dict(inpainted=inp, g_out=cG.g_out, mask=mask, ctx=dict(G=cG, DgGen=cg, DlGen=cl, DgReal=c1, DgFake=c2, DlReal=c3, DlFake=c4))
        return self.scalars

    # ------------------------------------------------------------------------------------------ reporting
    def read_scalars(self) -> Dict[str, float]:
        """ONE device->host copy of everything the reference logs per batch."""
        v = self.scalars.tolist()
        out = {k: v[i] for i, k in enumerate(SLOTS)}
        out["recon"] = v[RECON_SLOT]
        out["recon_local"] = v[RECON2_SLOT]
        if self.recipe == "global_local":
            out["g_loss"] = out["g_adv"] + out["g_adv_local"] + self.lambdas[0] * out["recon"] + self.lambdas[1] * out["recon_local"]
            out["d_loss"] = out["d_loss_real"] + out["d_loss_fake"] + out["d_loss_real_local"] + out["d_loss_fake_local"]
        elif self.recipe == "global_d":
            out["g_loss"] = out["g_adv"] + self.lambdas[0] * out["recon"]
            out["d_loss"] = out["d_loss_real"] + out["d_loss_fake"]
        else:
            out["g_tv"], out["g_perceptual"], out["g_style"] = v[TV_SLOT], v[PERCEPTUAL_SLOT], v[STYLE_SLOT]
            out["g_loss"] = out["g_adv"] + out["recon"]
            if self.local_recon >= 0:
                out["g_loss"] += out["recon_local"]
            if self.tv_weight:
                out["g_loss"] += out["g_tv"]
            if self.feature_weights is not None:
                out["g_loss"] += out["g_perceptual"] + out["g_style"]
            out["d_loss"] = out["d_loss_real"] - out["d_loss_fake"] if self.recipe == "wgan" else out["d_loss_real"] + out["d_loss_fake"]
        return out

    @torch.no_grad()
    def inpaint(self, ground: torch.Tensor, mask: torch.Tensor, train_mode: bool = False):
        """Inference (config 5; visualize.py:69-76, evaluate.py:127-136): ceil -> mask out -> G -> composite."""
        ground, mask = ground.contiguous().float(), mask.contiguous().float()
        mask_c, masked = torch.empty_like(mask), torch.empty_like(ground)
        call("b200_mask_prepare", ptr(ground), ptr(mask), ground.numel(), 1, ptr(mask_c), ptr(masked), stream())
        c = self.eG.forward(masked, train_mode, comp_base=masked, comp_mul=mask_c, keep=False)
        return c.comp


class _CapturedStep:
    """One train step of a fixed batch shape as a CUDA graph.  Inputs are copied into static buffers (one device-to-device copy
    of ground + mask per step); dropout masks stay fresh because the kernels add the device counter the graph's first node
    increments to their seed; Adam's bias correction and BatchNorm's ``num_batches_tracked`` already live on the device."""

    def __init__(self, tr: "InpaintingTrainer", ground, mask, update_g):
        self.ground = torch.empty_like(ground, dtype=torch.float32).contiguous()
        self.mask = torch.empty_like(mask, dtype=torch.float32).contiguous()
        self.ground.copy_(ground)
        self.mask.copy_(mask)
        sc = Scratch.get(tr.device)
        # eager-allocated buffers whose raw addresses the graph bakes in must outlive it even if their owners re-allocate
        self.keep = [sc.bn_partial, sc.coef, sc.c1_partial, sc.recon, sc.wgrad, sc.split, tr.scalars, tr._seed_ctr]
        torch.cuda.synchronize(tr.device)
        n0 = _lib.value("b200_launch_count")
        self.graph = torch.cuda.CUDAGraph()
        _lib.value("b200_set_seed_counter", ptr(tr._seed_ctr))
        try:
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                call("b200_step_inc", ptr(tr._seed_ctr), stream())
                tr._run(self.ground, self.mask, update_g, tr._seed_base)
        finally:
            _lib.value("b200_set_seed_counter", None)
        self.launches = int(_lib.value("b200_launch_count") - n0)
        _lib.value("b200_launch_count_add", -self.launches)  # recorded, not run
        self.last = tr.last

    def replay(self, ground, mask):
        if ground.data_ptr() != self.ground.data_ptr():
            self.ground.copy_(ground)
        if mask.data_ptr() != self.mask.data_ptr():
            self.mask.copy_(mask)
        self.graph.replay()
        _lib.value("b200_launch_count_add", self.launches)


def ctypes_ptr(t: torch.Tensor, index: int):
    import ctypes

    return ctypes.c_void_p(t.data_ptr() + index * t.element_size())

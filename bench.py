#!/usr/bin/env python
"""Benchmark of the GAN-inpainting train step (BASELINE.json metric: train-step images/sec at 256x256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--recipe wgan_l1|wgan_rmse|minimaxgan_rmse|minimaxgan_l1|wgan_d_only|inference]

Workload at N=1 (BASELINE.json configs[1]): ``wgan_l1`` — U-Net generator + PatchGAN global discriminator, 256x256,
batch 64 per GPU, bf16 operands / fp32 accumulation.  One "step" is the full step of experiment_list/wgan_l1.py:110-186
on one synthetic batch: ceil(mask) -> mask-out -> G forward -> composite -> D update (2 forward, 2 backward, RMSprop,
clip) -> G update (D forward, D data-gradient, G backward, RMSprop).  Nothing is skipped inside the timed region.

One JSON line on stdout (rank 0):
  value      images/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e        the same step through the public API with HOST inputs: per step a pinned-host -> device copy of
             (ground, mask) and a device -> host read of the loss vector, inside the timed region
  roofline   tensor-pipe roofline of the dominant kernel (the tcgen05 implicit-GEMM body that runs every 4x4/s2
             convolution, its data gradient and its weight gradient): algorithmic FLOPs / measured kernel time
  tail       achieved HBM GB/s of every memory-bound kernel of the step (algorithmic bytes of the call / its CUDA-event time)
  cpu_baseline  the oracle port of the reference step on this box's host cores (bounded sample, rank 0, N=1)
  also       (default recipe, N=1) short measurements of SURVEY.md §8(d)'s other two lines: the WGAN D-only step (4 of 5
             steady-state batches, wgan_l1.py:157-163) and config 5's inference pass (512x512, batch 128)

Other recipes: `--recipe wgan_d_only` (D update only), `--recipe inference` (config 5: ceil(mask) -> mask-out -> G (eval) ->
composite through trainer.inpaint at 512x512, batch 128 per GPU; e2e copies the images H2D and the result D2H every step),
`--recipe minimaxgan_rmse` (config 3).

`--impl reference` times the reference algorithm's CPU implementation (oracle/, fp32 PyTorch on all host cores) on a
bounded sample of the same workload and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: NCCL logs to stdout by default, so its log file defaults to stderr (NCCL_DEBUG itself is
# left as the caller set it)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

RECIPES = {
    # name: (trainer recipe, recon)
    "wgan_l1": ("wgan", "l1"),
    "wgan_rmse": ("wgan", "rmse"),
    "minimaxgan_rmse": ("minimax", "rmse"),
    "minimaxgan_l1": ("minimax", "l1"),
    # config 4 (wgan_perceptual_style_faceparsing.py:197-225) without its face-parsing term (needs a trained segmentation network):
    # WGAN + RMSE + local (masked MSE) + TV + VGG19 perceptual / style losses
    "wgan_perceptual_style": ("wgan", "rmse"),
    "wgan_d_only": ("wgan", "l1"),     # the D-only batches of the WGAN schedule (update_g=False)
    "inference": ("wgan", "l1"),       # config 5: trainer.inpaint, no discriminator work
}
CONFIG4 = dict(tv_weight=1.0, feature_weights=(0.01, 0.01), local_recon="mse")  # :209-219
DEFAULT_SHAPE = {"inference": (512, 128), "wgan_perceptual_style": (256, 32)}  # (res, batch per GPU); everything else 256x256, batch 64 (BASELINE configs[1])


def metric_name(recipe: str, res: int) -> str:
    return f"inference images/sec at {res}x{res}" if recipe == "inference" else f"train-step images/sec at {res}x{res}"


def conv_flops(res: int):
    """Algorithmic forward conv / linear FLOPs (2*MAC) per image of the two networks: U-Net generator
    (7 levels, ngf 64: 64-128-256-512-512-512-512, 4x4 stride-2 convolutions down and transposed convolutions up, decoder inputs
    doubled by the skip concatenation except the innermost) and PatchGAN (1-64-128-256-512, 4x4/s1 head conv, Linear)."""
    enc = [(1, 64), (64, 128), (128, 256), (256, 512), (512, 512), (512, 512), (512, 512)]
    dec = [(512, 512), (1024, 512), (1024, 512), (1024, 256), (512, 128), (256, 64), (128, 1)]
    g, r = 0.0, res
    for cin, cout in enc:
        r //= 2
        g += 2.0 * r * r * cout * cin * 16
    for cin, cout in dec:
        g += 2.0 * r * r * cin * cout * 16
        r *= 2
    d, r, cin = 0.0, res, 1
    for i in range(4):
        r //= 2
        d += 2.0 * r * r * (64 << i) * cin * 16
        cin = 64 << i
    r -= 3
    d += 2.0 * r * r * cin * 16 + 2.0 * r * r
    return g, d


def vgg_flops(res: int) -> float:
    """Tensor-core FLOPs per image of ONE pass of vgg19.features[0..29] (conv1_2 .. conv5_1; conv1_1 runs as a direct kernel) plus
    the five Gram matrices (loss.py:117-136)."""
    cfg = [(64, 64, 0), (64, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 2), (256, 256, 2), (256, 256, 2), (256, 512, 3), (512, 512, 3), (512, 512, 3),
           (512, 512, 3), (512, 512, 4)]
    f = sum(2.0 * (res >> lv) ** 2 * 9 * cin * cout for cin, cout, lv in cfg)
    f += sum(2.0 * (res >> lv) ** 2 * c * c for lv, c in enumerate((64, 128, 256, 512, 512)))
    return f


def flops_per_image(res: int, recipe: str) -> dict:
    """Algorithmic conv FLOPs (2*MAC) per image, SURVEY.md §8(d): full step = 3*G + 8*D."""
    g, d = conv_flops(res)
    # tensor-core part: everything but the 1-channel-side layers (d1, u1, c1) and the D head
    g1 = 2.0 * (res // 2) ** 2 * 64 * 16            # d1
    gu1 = 2.0 * (res // 2) ** 2 * 128 * 16          # u1
    dc1 = 2.0 * (res // 2) ** 2 * 64 * 16           # c1
    dhead = d - sum(2.0 * (res >> (i + 1)) ** 2 * (64 << i) * (1 if i == 0 else (64 << (i - 1))) * 16 for i in range(4))
    g_tc, d_tc = g - g1 - gu1, d - dc1 - dhead
    if recipe == "inference":
        return dict(total=g, tensor=g_tc)
    if recipe == "wgan_d_only":  # G forward + D: 2 forward, 2 x (data + weight gradient)   (1*G + 6*D)
        return dict(total=g + 6 * d, tensor=g_tc + 6 * d_tc)
    # G: fwd + dgrad + wgrad (d2's dgrad to a1 IS needed; d1 has none).  D: 2 fwd + 2*(dgrad+wgrad) in the D step
    # (c2's dgrad is needed for c1's wgrad) + 1 fwd + 1 dgrad in the G step.
    if recipe == "wgan_perceptual_style":  # + two feature-extractor passes (inpainted, ground) per image; its first conv is not tensor-core work
        v = 2 * vgg_flops(res)
        return dict(total=3 * g + 8 * d + v + 2 * 2.0 * res * res * 27 * 64, tensor=3 * g_tc + 8 * d_tc + v)
    return dict(total=3 * g + 8 * d, tensor=3 * g_tc + 8 * d_tc)


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0: float = None, t1: float = None) -> dict:
        """Summary of the samples taken in the wall-clock window [t0, t1] (the timed region; nvidia-smi is started before the
        warm-up so that its start-up cost stays outside).  Falls back to the last samples when none fell inside a short window."""
        import datetime

        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        samples = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                reasons = {name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[6:10])
                           if v.strip().lower() == "active"}
                samples.append((ts, float(r[2]), float(r[3]), float(r[4]), reasons))
            except (ValueError, IndexError):
                continue
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [x for x in samples if t0 is not None and t0 <= x[0] <= t1 + 0.06]
        window = "timed region"
        if not inside:
            inside, window = samples[-3:], "last samples (none fell inside the timed region)"
        sm = sorted(x[1] for x in inside)
        reasons = set().union(*(x[4] for x in inside))
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(x[2] for x in inside), "power_w_max": max(x[3] for x in inside), "samples": len(inside),
                "samples_total": len(samples), "window": window, "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference step on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_step_runner(recipe: str, res: int, batch: int):
    import torch

    from oracle import nets, steps, synth

    torch.set_num_threads(os.cpu_count())
    kind, recon = RECIPES[recipe]
    ground, mask = synth.make_batch(batch, res, seed=1234)
    if recipe == "inference":
        G = nets.init_unet(0)
        return lambda: steps.inference(G, ground, mask)
    G, D = steps.enable_grad(nets.init_unet(0)), steps.enable_grad(nets.init_patchgan(1, res=res))
    if kind == "wgan":
        go, do = steps.make_optimizer("rmsprop", G), steps.make_optimizer("rmsprop", D)
        extra = {}
        if recipe == "wgan_perceptual_style":
            from oracle import vgg_losses

            extra = dict(tv_weight=CONFIG4["tv_weight"], vgg_features=vgg_losses.seeded_vgg_features(0), feature_weights=CONFIG4["feature_weights"],
                         local_recon=CONFIG4["local_recon"])
        fn = lambda g, m: steps.wgan_step(G, D, go, do, g, m, recon=recon, update_g=(recipe != "wgan_d_only"), **extra)
    else:
        go, do = steps.make_optimizer("adam", G), steps.make_optimizer("adam", D)
        fn = lambda g, m: steps.minimax_step(G, D, go, do, g, m, recon=recon)
    return lambda: fn(ground, mask)


def time_cpu(recipe: str, res: int, batch: int, steps_n: int, warmup: int):
    run = cpu_step_runner(recipe, res, batch)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps_n):
        run()
    dt = time.perf_counter() - t0
    return batch * steps_n / dt, dt / steps_n


def workload_name(recipe: str, res: int, batch: int, what: str) -> str:
    if recipe == "inference":
        return f"inference (config 5): ceil(mask) -> mask-out -> U-Net generator (eval) -> composite, {res}x{res}, batch {batch} {what}"
    if recipe == "wgan_perceptual_style":
        return (f"wgan_perceptual_style_faceparsing without its face-parsing term: U-Net generator + PatchGAN global D full train step, g_loss = adversarial + "
                f"RMSE + local MSE + TV + VGG19 perceptual/style (random-init VGG19), {res}x{res}, batch {batch} {what}")
    if recipe == "wgan_d_only":
        return f"wgan_l1 D-only batch (G forward + composite + D update, no G update; wgan_l1.py:157-163), {res}x{res}, batch {batch} {what}"
    return f"{recipe}: U-Net generator + PatchGAN global D, full train step (D update + G update), {res}x{res}, batch {batch} {what}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.ref_batch
    ips, spt = time_cpu(args.recipe, args.res, batch, args.steps, args.warmup)
    sample = f"{args.steps} steps x {batch} images of the {args.recipe} step at {args.res}x{args.res} (fp32, torch CPU, {os.cpu_count()} threads)"
    line = {
        "impl": "reference", "metric": metric_name(args.recipe, args.res), "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": spt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.recipe, args.res, batch, "per step (bounded CPU sample of the B200 arm's workload)"),
                   "recipe": args.recipe, "res": args.res, "batch_per_step": batch},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# algorithmic HBM bytes of the memory-bound C-ABI calls (from their arguments; include/b200_inpaint.h gives the signatures)
# ----------------------------------------------------------------------------------------------------------------
def _has(a):
    return a is not None and getattr(a, "value", a) not in (None, 0)


def _i(a):
    return int(getattr(a, "value", a))


TAIL_BYTES = {
    # read y (bf16) + one or two bf16 stores
    "b200_bn_act": lambda a: _i(a[1]) * _i(a[2]) * 2 * (1 + _has(a[6]) + _has(a[9])),
    # read y (bf16) + dz (fp32 or bf16), write dy (bf16)
    "b200_bn_bwd_fused": lambda a: _i(a[1]) * _i(a[2]) * (2 + (4 if _i(a[8]) else 2) + 2),
    # with batch statistics: a reduce pass and an apply pass, each reading y, g1 (and g2); one bf16 store
    "b200_bn_act_bwd": lambda a: _i(a[1]) * _i(a[2]) * 2 * ((2 + _has(a[9])) * (2 if _has(a[3]) else 1) + 1),
    "b200_conv_c1_fwd": lambda a: _i(a[2]) * _i(a[3]) * _i(a[4]) * 4 * (1 + _has(a[1])) + _i(a[2]) * (_i(a[3]) // 2) * (_i(a[4]) // 2) * _i(a[6]) * 2 * (1 + _has(a[9])),
    "b200_convt_c1_fwd": lambda a: _i(a[1]) * _i(a[2]) * _i(a[3]) * _i(a[4]) * 2 + _i(a[1]) * 4 * _i(a[2]) * _i(a[3]) * 4 * (1 + _has(a[9]) + _has(a[10]) + _has(a[12])),
    "b200_wgrad_c1": lambda a: _i(a[1]) * _i(a[2]) * _i(a[3]) * _i(a[4]) * 2 + _i(a[1]) * 4 * _i(a[2]) * _i(a[3]) * 4 * (1 + _has(a[7])),
    "b200_wgrad_reduce": lambda a: 4 * (_i(a[1]) + 1 + _i(a[5])) * _i(a[2]) * _i(a[3]) * 16,
    "b200_rmsprop_step": lambda a: 16 * _i(a[3]),
    "b200_adam_step": lambda a: 28 * _i(a[4]),
    "b200_mask_prepare": lambda a: 4 * _i(a[2]) * (2 + _has(a[4]) + _has(a[5])),
    "b200_recon_loss": lambda a: 4 * _i(a[4]) * (2 + _has(a[3])),
    "b200_gstep_seed": lambda a: 4 * _i(a[14]) * 6,
    "b200_head_fwd": lambda a: 2 * _i(a[2]) * _i(a[3]),
    "b200_head_bwd_input": lambda a: 2 * _i(a[2]) * _i(a[3]),
}


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        from gan_inpainting_b200.trainer import nccl_group_options

        dist.init_process_group("nccl", device_id=dev, pg_options=nccl_group_options(args.nccl_ctas))
    from gan_inpainting_b200 import _lib, build
    from gan_inpainting_b200.lib.models import networks
    from gan_inpainting_b200.data import DevicePrefetcher, synthetic_batch
    from gan_inpainting_b200.trainer import InpaintingTrainer

    if rank == 0:
        build.build(verbose=False)
    if dist:
        dist.barrier()
    B, R = args.batch, args.res
    kind, recon = RECIPES[args.recipe]
    infer = args.recipe == "inference"
    update_g = args.recipe != "wgan_d_only"
    torch.manual_seed(0)
    G = networks.get_network("generator", "unet").to(dev).train()  # reference default: dropout ON
    D = networks.PatchGANDiscriminator(sigmoid=(kind != "wgan"), imagedim=R).to(dev).train()
    tr = InpaintingTrainer(kind, G, D, recon=recon, distributed=(world > 1 and not infer), bucket_mb=args.bucket_mb, graph=not args.no_graph, rank=rank,
                           comm_sms=args.comm_sms, wire_bf16=(None if args.grad_wire is None else args.grad_wire == "bf16"),
                           **(CONFIG4 if args.recipe == "wgan_perceptual_style" else {}))

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    class HostBatches:
        """The pinned host batch, served n times (a DataLoader stand-in: every step gets its own host -> device copy)."""

        def __init__(self, g, m, n):
            self.g, self.m, self.n = g, m, n

        def __iter__(self):
            for _ in range(self.n):
                yield self.g, self.m, None

    def workload(recipe, R, B):
        """(device-resident step, e2e runner, D2H bytes per step) of one recipe at one shape."""
        ground_h, mask_h = synthetic_batch(B, R, seed=1234, rank=rank)
        ground_h, mask_h = ground_h.pin_memory(), mask_h.pin_memory()
        ground, mask = ground_h.to(dev), mask_h.to(dev)
        if recipe == "inference":
            step = lambda: tr.inpaint(ground, mask)
            res_h = torch.empty((B, 1, R, R), dtype=torch.float32).pin_memory()

            def e2e(n):
                """DevicePrefetcher (pinned host -> device copy of every step's images and masks) -> trainer.inpaint -> the inpainted
                images copied back to pinned host memory every step (visualize.py:69-76 then saves them)."""
                for g_d, m_d, _ in DevicePrefetcher(HostBatches(ground_h, mask_h, n), dev):
                    res_h.copy_(tr.inpaint(g_d, m_d), non_blocking=True)
                    torch.cuda.current_stream().synchronize()

            return step, e2e, 2 * ground_h.numel() * 4, res_h.numel() * 4
        upd = recipe != "wgan_d_only"
        step = lambda: tr.step(ground, mask, update_g=upd)
        out_h = torch.empty(tr.scalars.numel(), dtype=torch.float32).pin_memory()

        def e2e(n):
            """n steps through the public API: DevicePrefetcher (pinned host -> device copy of EVERY step's inputs, overlapped with
            the previous step) -> InpaintingTrainer.step -> the loss vector read back to the host every step (wgan_l1.py:199-201)."""
            for g_d, m_d, _ in DevicePrefetcher(HostBatches(ground_h, mask_h, n), dev):
                tr.step(g_d, m_d, update_g=upd)
                out_h.copy_(tr.scalars, non_blocking=True)
                torch.cuda.current_stream().synchronize()

        return step, e2e, 2 * ground_h.numel() * 4, out_h.numel() * 4

    step_dev, run_e2e, h2d_bytes, d2h_bytes = workload(args.recipe, R, B)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    l0 = _lib.value("b200_launch_count")
    sync_all()
    w0 = time.time()
    ms = timed(step_dev, args.steps)
    w1 = time.time()
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    launches = _lib.value("b200_launch_count") - l0
    # host-side enqueue cost of one step (Python + ctypes + allocator): must stay below the device time or the GPU starves
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        step_dev()
    host_ms = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize()
    run_e2e(2)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1)
    ips = B * world * args.steps / (ms * 1e-3)
    ips_e2e = B * world * args.steps / (ms_e2e * 1e-3)

    # ---- per-kernel time of one instrumented step (CUDA events around every C-ABI call, on the launching stream)
    prof, nbytes = {}, {}

    def hook(name, cargs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if name in TAIL_BYTES:
            nbytes[name] = nbytes.get(name, 0) + TAIL_BYTES[name](cargs)

        def after():
            e1.record()
            prof.setdefault(name, []).append((e0, e1))

        return after

    kernel_ms = {}
    tr.graph = False  # the instrumented steps are enqueued kernel by kernel (events around every C-ABI call)
    gemm_names = ("b200_conv_s2_gemm", "b200_convt_s2_gemm", "b200_wgrad_gemm", "b200_conv3_gemm", "b200_gram_gemm")
    runs = []
    for _ in range(3):  # three instrumented steps, the one with the median GEMM time is reported (a single step is +-5 % noisy)
        prof.clear()
        nbytes.clear()
        if rank == 0:
            _lib.PROFILE = hook
        # keep the GPU BEHIND the host while the ~210 calls and ~420 events are enqueued (about 5 ms of Python): every event then
        # brackets its kernel back to back with its neighbours, and no interval contains time the GPU spent waiting for a launch
        torch.cuda._sleep(int(0.03 * 1.9e9))
        step_dev()  # every rank takes part (the step contains the gradient all-reduces); only rank 0 is instrumented
        torch.cuda.synchronize()
        _lib.PROFILE = None
        if rank == 0:
            km = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in prof.items()}
            runs.append((sum(km.get(k, 0.0) for k in gemm_names), km, {k: len(v) for k, v in prof.items()}, dict(nbytes)))
    tr.graph = not args.no_graph
    calls = {}
    if rank == 0:
        runs.sort(key=lambda r: r[0])
        _, kernel_ms, calls, nb = runs[len(runs) // 2]
        nbytes.clear()
        nbytes.update(nb)
    if dist:
        dist.barrier()

    # ---- SURVEY.md §8(d)'s other two lines, measured briefly next to the headline (default recipe, one GPU)
    also = None
    if world == 1 and args.recipe == "wgan_l1" and not args.no_also:
        also = {}
        try:
            s2, e2, hb, db = workload("wgan_d_only", R, B)
            for _ in range(3):
                s2()
            m2 = timed(s2, 10)
            e2(2)
            m2e = timed(lambda: e2(10), 1)
            fl2 = flops_per_image(R, "wgan_d_only")
            also["wgan_d_only"] = {"workload": workload_name("wgan_d_only", R, B, "per GPU"), "value": B * 10 / (m2 * 1e-3), "unit": "images/s",
                                   "ms_per_step": m2 / 10, "e2e": B * 10 / (m2e * 1e-3), "steps": 10, "whole_step_tflops": fl2["total"] * B * 10 / (m2 * 1e-3) / 1e12}
            Ri, Bi = DEFAULT_SHAPE["inference"]
            s3, e3, hb, db = workload("inference", Ri, Bi)
            for _ in range(3):
                s3()
            m3 = timed(s3, 5)
            e3(1)
            m3e = timed(lambda: e3(5), 1)
            fl3 = flops_per_image(Ri, "inference")
            also["inference"] = {"workload": workload_name("inference", Ri, Bi, "per GPU"), "value": Bi * 5 / (m3 * 1e-3), "unit": "images/s",
                                 "ms_per_step": m3 / 5, "e2e": Bi * 5 / (m3e * 1e-3), "h2d_bytes_per_step": hb, "d2h_bytes_per_step": db, "steps": 5,
                                 "whole_step_tflops": fl3["total"] * Bi * 5 / (m3 * 1e-3) / 1e12}
        except Exception as exc:  # the headline line must survive a problem in the extras
            also["error"] = repr(exc)

    def teardown():
        # captured NCCL collectives pin the communicator: drop the step graphs first; a watchdog guarantees that a teardown problem
        # can never hang the job after the result line has been printed
        t = threading.Timer(30.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        tr.release_graphs()
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        t.cancel()

    if rank != 0:
        teardown()
        return
    peaks = measured_peaks()
    fl = flops_per_image(R, args.recipe)
    gemm_ms = sum(kernel_ms.get(k, 0.0) for k in gemm_names)
    gemm_calls = sum(calls.get(k, 0) for k in gemm_names)
    achieved = fl["tensor"] * B / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    traffic, traffic_src = None, None  # DRAM bytes per igemm launch from the committed ncu capture of this command (profiles/)
    for tname in ("r2_igemm_traffic.json", "r1_igemm_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath) and B == 64 and R == 256 and args.recipe == "wgan_l1":
            traffic, traffic_src = json.load(open(tpath)).get("dram_bytes_per_launch"), tname
            break
    step_ms_instr = sum(kernel_ms.values())
    # which measured peak applies: the timed region of a default run is a fraction of a second at boost clocks and far below the
    # power cap — the burst figure (a kernel timed alone); a long run that pulls the clocks down is compared with the sustained one
    at_boost = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.95 * clocks["sm_max_mhz"])
    peak, peak_name = (peaks["tf_burst"], "bf16_tflops (burst)") if at_boost else (peaks["tf_sustained"], "bf16_tflops_sustained")
    roofline = {
        "bound": "tensor", "kernel": "igemm_kernel (tcgen05 implicit GEMM: conv / convT phases / wgrad)", "achieved": achieved,
        "peak": peak, "peak_source": f"{peaks['source']} {peak_name}: SM clock {clocks.get('sm_mhz') if clocks else None} MHz during the timed region",
        "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
        "frac_of_burst": (achieved / peaks["tf_burst"]) if achieved else None, "frac_of_sustained": (achieved / peaks["tf_sustained"]) if achieved else None,
        "traffic": traffic, "traffic_unit": (f"DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/{traffic_src})" if traffic_src
                                            else "no committed ncu capture for this recipe / shape"),
        "launches_per_step": gemm_calls, "avg_launch_ms": gemm_ms / max(gemm_calls, 1), "algorithmic_gflop_per_step": fl["tensor"] * B / 1e9,
        "share_of_step": gemm_ms / step_ms_instr if step_ms_instr else None,
        "step_breakdown_ms": {k: round(v, 4) for k, v in sorted(kernel_ms.items(), key=lambda kv: -kv[1])},
        "whole_step_tflops": fl["total"] * B * world * args.steps / (ms * 1e-3) / 1e12,
    }
    tail = {k: {"ms": round(kernel_ms[k], 4), "algorithmic_mb": round(nbytes[k] / 1e6, 1), "gbps": round(nbytes[k] / (kernel_ms[k] * 1e-3) / 1e9, 1),
                "frac_of_hbm_peak": round(nbytes[k] / (kernel_ms[k] * 1e-3) / 1e9 / peaks["hbm_gbs"], 3), "calls": calls.get(k, 0)}
            for k in sorted(nbytes, key=lambda k: -kernel_ms.get(k, 0.0)) if kernel_ms.get(k, 0.0) > 0}
    tail_ms = sum(v for k, v in kernel_ms.items() if k not in gemm_names)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = args.ref_batch
        c_ips, c_spt = time_cpu(args.recipe, R, cb, 5, 1)
        cpu = {"value": c_ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"5 steps x {cb} images of the same {args.recipe} workload at {R}x{R} (oracle port of the reference, fp32 torch CPU, all host threads)"}
    line = {
        "metric": metric_name(args.recipe, R), "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args.recipe, R, B, "per GPU"),
                   "recipe": args.recipe, "res": R, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "dropout": not infer, "l2": "no flush: one step streams several GB of activations/gradients, far larger than the 126 MB L2",
                   "bn": "running statistics (eval)" if infer else "per-replica batch statistics", "bucket_mb": args.bucket_mb,
                   "grad_wire": ("bf16" if getattr(tr, "wire_bf16", False) else "fp32") if world > 1 else None,
                   "nccl_ctas": (args.nccl_ctas if args.nccl_ctas is not None else int(os.environ.get("B200_NCCL_CTAS", "0"))) if world > 1 else 0,
                   "comm_sms": (args.comm_sms if args.comm_sms is not None else int(os.environ.get("B200_COMM_SMS", "0"))) if world > 1 else 0,
                   "launch": "eager (one ctypes call per kernel)" if (args.no_graph or infer) else "whole step replayed as one CUDA graph"},
        "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": host_ms,
        "clocks": clocks,
        "roofline": roofline,
        "tail": {"ms_per_step": round(tail_ms, 4), "hbm_peak_gbs": peaks["hbm_gbs"], "kernels": tail},
        "cpu_baseline": cpu,
        "also": also,
    }
    print(json.dumps(line), flush=True)
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--recipe", default="wgan_l1", choices=list(RECIPES))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 64; inference 128)")
    ap.add_argument("--res", type=int, default=None, help="image side (default 256; inference 512)")
    ap.add_argument("--ref-batch", type=int, default=4, help="images per step of the bounded CPU sample")
    ap.add_argument("--bucket-mb", type=float, default=25.0)
    ap.add_argument("--comm-sms", type=int, default=None, help="SMs kept free for the gradient all-reduces while they are in flight (default: B200_COMM_SMS or 0 = none)")
    ap.add_argument("--nccl-ctas", type=int, default=None, help="cap on NCCL's thread blocks per collective (default: B200_NCCL_CTAS or 0 = NCCL's own choice)")
    ap.add_argument("--grad-wire", default=None, choices=["fp32", "bf16"], help="dtype of the gradient buckets on the wire (default: B200_GRAD_WIRE or fp32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short D-only / inference measurements of the default run")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying the captured step graph")
    args = ap.parse_args()
    dres, dbatch = DEFAULT_SHAPE.get(args.recipe, (256, 64))
    args.res = args.res or dres
    args.batch = args.batch or dbatch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "b200" and args.gpus != world:
        if world == 1 and args.gpus > 1:
            # plain `python bench.py --gpus N`: relaunch under torchrun, one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

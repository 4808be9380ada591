"""Per-layer timing of the VGG19 feature extractor of the perceptual / style losses (config 4: 256x256, 2 x 32 images per step).

    python tools/vgg_table.py [--batch 32] [--res 256] [--reps 20]

CUDA events around every launch of one feature-extractor pass over [output; target] (2 x batch images), L2 flushed between
launches by the size of the activations themselves (>= 268 MB per layer at the default shape, except the 16x16 layers).
Prints TFLOP/s per convolution against the measured bf16 peak in MEASURED_PEAKS.json and the whole loss evaluation's time.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    from gan_inpainting_b200 import _lib
    from gan_inpainting_b200.lib.models import loss, networks

    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("bf16_tflops", 1636.9))
    torch.manual_seed(0)
    w = networks.VGG19Wrapper().to(dev).eval()
    loss.set_vgg(w)
    B, R = a.batch, a.res
    tgt = torch.rand(B, 1, R, R, device=dev)
    out = (tgt + 0.2 * torch.randn_like(tgt)).clamp(0, 1)
    rows = []
    orig = _lib.call

    def timed(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(name, *args)
        e1.record()
        rows.append((name, args, e0, e1))
        return r

    for _ in range(3):
        loss.perceptual_and_style_loss(out, tgt, 0.01, 0.01)
    torch.cuda.synchronize()
    # whole evaluation, untimed-per-launch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        loss.perceptual_and_style_loss(out, tgt, 0.01, 0.01)
    e1.record()
    torch.cuda.synchronize()
    whole = e0.elapsed_time(e1) / a.reps
    import gan_inpainting_b200.engine as engine
    import gan_inpainting_b200.lib.models.loss as L

    acc = {}
    for rep in range(a.reps):
        rows.clear()
        engine.call = L.call = timed
        try:
            loss.perceptual_and_style_loss(out, tgt, 0.01, 0.01)
        finally:
            engine.call = L.call = orig
        torch.cuda.synchronize()
        for i, (name, args, s, e) in enumerate(rows):
            acc.setdefault(i, [name, args, 0.0])[2] += s.elapsed_time(e)
    tot_flop = tot_ms = 0.0
    print(f"VGG19 feature losses, [output; target] = 2 x {B} images of {R}x{R}; bf16 peak {peak:.0f} TFLOP/s (burst)")
    print(f"{'kernel':22s} {'shape':28s} {'us':>9s} {'TFLOP/s':>9s} {'frac':>6s}")
    for i in sorted(acc):
        name, args, ms = acc[i]
        ms /= a.reps
        shape, fl = "", 0.0
        if name == "b200_conv3_gemm":
            _, b, h, wd, cin, _, _, _, cout = args[:9]
            fl = 2.0 * b * h * wd * 9 * cin * cout
            shape = f"{b}x{h}x{wd} {cin}->{cout}"
        elif name == "b200_conv3_c1_fwd":
            _, b, h, wd = args[:4]
            fl = 2.0 * b * h * wd * 27 * 64
            shape = f"{b}x{h}x{wd} 3->64 (direct)"
        elif name == "b200_gram_gemm":
            _, b, h, wd, c = args[:5]
            fl = 2.0 * b * h * wd * c * c
            shape = f"{b}x{h}x{wd} C={c}"
        elif name == "b200_maxpool2":
            _, b, h, wd, c = args[:5]
            shape = f"{b}x{h}x{wd}x{c}"
        tf = fl / (ms * 1e-3) / 1e12 if fl else 0.0
        tot_flop += fl if name != "b200_conv3_c1_fwd" else 0.0
        tot_ms += ms
        print(f"{name[5:]:22s} {shape:28s} {ms * 1e3:9.1f} {tf:9.1f} {tf / peak if fl else 0:6.2f}")
    print(f"sum of launches {tot_ms:.3f} ms; whole evaluation (back to back) {whole:.3f} ms = {tot_flop / (whole * 1e-3) / 1e12:.0f} TFLOP/s of tensor-core work "
          f"({tot_flop / 1e12:.2f} TFLOP), {tot_flop / (whole * 1e-3) / 1e12 / peak:.2f} of peak")


if __name__ == "__main__":
    main()

"""From an ncu launch list of `bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-also` (wgan_l1, 256x256, batch 64):

    python tools/gemm_table.py profiles/r2_bench_launches_ncu.csv profiles/r2_gemm_per_launch.txt profiles/r2_igemm_traffic.json

writes the per-launch table of the 60 igemm launches of the last complete step (layer, variant, us, algorithmic GFLOP, TFLOP/s,
DRAM MB) and the DRAM bytes per igemm launch that bench.py reports as `roofline.traffic`.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import launch_table as lt  # noqa: E402

BURST, SUSTAINED = 1636.9, 1376.0  # MEASURED_PEAKS.json (TFLOP/s)
B = 64


def conv(hw_out, cin, cout):
    return 2.0 * B * hw_out * hw_out * cout * 16 * cin / 1e9


def convt(hw_in, cin, cout):
    return 2.0 * B * hw_in * hw_in * cin * 16 * cout / 1e9


def layer_table():
    """(name, GFLOP) of the 60 GEMM launches of one wgan_l1 step in launch order (engine.py / trainer.py)."""
    d = {2: conv(64, 64, 128), 3: conv(32, 128, 256), 4: conv(16, 256, 512), 5: conv(8, 512, 512), 6: conv(4, 512, 512), 7: conv(2, 512, 512)}
    u = {7: convt(2, 512, 512), 6: convt(4, 1024, 512), 5: convt(8, 1024, 512), 4: convt(16, 1024, 256), 3: convt(32, 512, 128), 2: convt(64, 256, 64)}
    out = [("G d%d" % l, d[l]) for l in range(2, 8)] + [("G u%d" % l, u[l]) for l in range(7, 1, -1)]
    dfwd = [("D c2", d[2]), ("D c3", d[3]), ("D c4", d[4])]
    out += [(n + " (real)", g) for n, g in dfwd] + [(n + " (fake)", g) for n, g in dfwd]
    dbwd = [("D wgrad c4", d[4]), ("D dgrad c4", d[4]), ("D wgrad c3", d[3]), ("D dgrad c3", d[3]), ("D wgrad c2", d[2]), ("D dgrad c2", d[2])]
    out += dbwd + dbwd
    out += [(n + " (G step)", g) for n, g in dfwd] + [("D dgrad c4 (frozen D)", d[4]), ("D dgrad c3 (frozen D)", d[3]), ("D dgrad c2 (frozen D)", d[2])]
    for l in range(2, 8):
        out += [("G wgrad u%d" % l, u[l]), ("G dgrad u%d" % l, u[l])]
    for l in range(7, 1, -1):
        out += [("G wgrad d%d" % l, d[l]), ("G dgrad d%d" % l, d[l])]
    return out


def main():
    src, out_txt, out_json = sys.argv[1:4]
    L = lt.load(src)
    ours = [x for x in L if "b200::" in x["name"]]
    starts = [i for i, x in enumerate(ours) if "mask_prepare" in x["name"]]
    step = ours[starts[-2]:starts[-1]]
    ig = [x for x in step if "igemm" in x["name"]]
    layers = layer_table()
    assert len(ig) == len(layers) == 60, (len(ig), len(layers))
    tf = lambda g, us: g / us * 1e3
    lines = [f"per-launch GEMM table of the last complete step of {os.path.basename(src)} (us under ncu: cold cache, serialised; TFLOP/s of the layer's algorithmic flops)",
             "variant tags: S = BatchNorm statistics epilogue, B = fused activation backward (L leaky / R relu, G = + skip gradient, N = + BatchNorm backward sums), "
             "pair = cta_group::2, win = window operand staging (all launches with >= 16 x 8 output pixels per image)", ""]
    for i, (x, (name, g)) in enumerate(zip(ig, layers)):
        by = x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0)
        lines.append("%2d %-26s %7.1f us %7.1f GF %6.0f TFLOP/s %7.1f MB DRAM  %s" % (i, lt.short(x["name"]), x["us"], g, tf(g, x["us"]), by / 1e6, name))
    tot, tus = sum(g for _, g in layers), sum(x["us"] for x in ig)
    lines += ["", "sum %.0f us for %.1f GF -> %.0f TFLOP/s = %.2f of the measured burst peak (%.1f), %.2f of sustained (%.1f)" % (
        tus, tot, tf(tot, tus), tf(tot, tus) / BURST, BURST, tf(tot, tus) / SUSTAINED, SUSTAINED)]
    gfw, gus = sum(g for _, g in layers[:12]), sum(x["us"] for x in ig[:12])
    lines.append("generator forward (launches 0-11): %.0f us for %.1f GF -> %.0f TFLOP/s = %.2f of burst" % (gus, gfw, tf(gfw, gus), tf(gfw, gus) / BURST))
    big = [i for i in range(12) if layers[i][1] > 60]
    bg, bu = sum(layers[i][1] for i in big), sum(ig[i]["us"] for i in big)
    lines.append("generator forward convolutions with >= 64 GFLOP (d2-d4, u5-u2: %.0f %% of its flops): %.0f TFLOP/s = %.2f of burst" % (100 * bg / gfw, tf(bg, bu), tf(bg, bu) / BURST))
    open(out_txt, "w").write("\n".join(lines) + "\n")
    by = sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in ig)
    json.dump({"source": f"profiles/{os.path.basename(src)} (ncu launch list of `bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-also`, last complete step)",
               "igemm_launches_per_step": len(ig), "dram_bytes_per_launch": by / len(ig), "igemm_dram_bytes_per_step": by, "igemm_us_per_step_ncu": tus,
               "step_dram_bytes": sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in step),
               "step_kernel_us_ncu": sum(x["us"] for x in step), "launches_per_step": len(step)}, open(out_json, "w"), indent=1)
    print("\n".join(lines[-3:]))


if __name__ == "__main__":
    main()

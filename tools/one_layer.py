"""A single GEMM launch of the bench's shapes (for ncu captures / A-B timing of igemm variants):

    python tools/one_layer.py convt|conv [B] [H] [Cin] [Cout] [reps] [stats]

`convt`: transposed convolution from [B,H,H,Cin]; `conv`: stride-2 convolution from [B,H,H,Cin]; `stats` = 1 adds the BatchNorm
statistics epilogue (the forward layers of the train step).  Prints the CUDA-event time per launch (L2 flushed in between) and
the TFLOP/s.  B200_IGEMM_2CTA=0 selects the single-CTA kernel, B200_WINDOW=0 one TMA box per filter tap.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gan_inpainting_b200 import engine as eng

kind = sys.argv[1] if len(sys.argv) > 1 else "convt"
B, H, Cin, Cout, reps, stats = (int(a) for a in (sys.argv[2:8] + ["64", "64", "128", "64", "20", "0"][len(sys.argv) - 2:]))
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
if kind == "convt":
    w = torch.randn(Cin, Cout, 4, 4, device=dev) / (Cin * 4) ** 0.5
else:
    w = torch.randn(Cout, Cin, 4, 4, device=dev) / (Cin * 16) ** 0.5
pk = eng._Packed(w)
pk.refresh(w)
bn = torch.nn.BatchNorm2d(Cout).to(dev)
sc = eng.Scratch.get(dev)


def run():
    plan = eng.BnPlan(bn, True, sc) if stats else None
    if kind == "convt":
        return eng.convt_s2(x, Cin, Cin, pk.phase, Cout, plan=plan)
    return eng.conv_s2(x, pk.gather, Cout, plan=plan)


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    y = run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()  # evict L2 between launches (the step streams GBs between two uses of a tensor)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
med = ts[len(ts) // 2]
fl = 2.0 * B * (4 * H * H * Cout * 4 * Cin if kind == "convt" else (H // 2) ** 2 * Cout * 16 * Cin)
print(f"{kind} B={B} H={H} Cin={Cin} Cout={Cout} stats={stats} pair={os.environ.get('B200_IGEMM_2CTA', '1')} window={os.environ.get('B200_WINDOW', '1')}: "
      f"{med * 1e3:.1f} us  {fl / med / 1e9:.0f} TFLOP/s")

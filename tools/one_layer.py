"""A single transposed-convolution GEMM launch of the bench's shapes (for ncu captures / A-B timing of igemm variants):

    python tools/one_layer.py [B] [Hi] [Cin] [Cout] [reps]

Runs the plain fprop (statistics epilogue) `reps` times and prints the CUDA-event time per launch and the TFLOP/s.
B200_IGEMM_2CTA=0 selects the single-CTA kernel.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gan_inpainting_b200 import engine as eng

B, Hi, Cin, Cout, reps = (int(a) for a in (sys.argv[1:6] + ["64", "64", "128", "64", "20"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
torch.manual_seed(0)
z = torch.randn(B, Hi, Hi, Cin, device=dev).to(torch.bfloat16)
w = torch.randn(Cin, Cout, 4, 4, device=dev) / (Cin * 4) ** 0.5
pk = eng._Packed(w)
pk.refresh(w)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    y = eng.convt_s2(z, Cin, Cin, pk.phase, Cout)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()  # evict L2 between launches (the step streams GBs between two uses of a tensor)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = eng.convt_s2(z, Cin, Cin, pk.phase, Cout)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
med = ts[len(ts) // 2]
fl = 2.0 * B * 4 * Hi * Hi * Cout * 4 * Cin
print(f"convt B={B} Hi={Hi} Cin={Cin} Cout={Cout} pair={os.environ.get('B200_IGEMM_2CTA', '1')}: "
      f"{med * 1e3:.1f} us  {fl / med / 1e9:.0f} TFLOP/s")

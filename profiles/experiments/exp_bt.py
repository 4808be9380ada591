"""EXPERIMENT: conv fprop GEMM with the weight operand packed [K][N] (B200_BT=1) vs [N][K]: python tools/exp_bt.py B H Cin Cout"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gan_inpainting_b200 import engine as eng

B, H, Cin, Cout = (int(a) for a in sys.argv[1:5])
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
w = torch.randn(Cout, Cin, 4, 4, device=dev) / (Cin * 16) ** 0.5
pk = eng._Packed(w)
pk.refresh(w)
wt = pk.gather.t().contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(bt):
    os.environ["B200_BT"] = "1" if bt else "0"
    wp = wt if bt else pk.gather
    for _ in range(3):
        y = eng.conv_s2(x, wp, Cout)
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = eng.conv_s2(x, wp, Cout)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return y, ts[len(ts) // 2]


y0, t0 = run(False)
y1, t1 = run(True)
fl = 2.0 * B * (H // 2) ** 2 * Cout * 16 * Cin
err = float((y1.float() - y0.float()).norm() / y0.float().norm())
print(f"conv B={B} H={H} Cin={Cin} Cout={Cout}: [N][K] {t0 * 1e3:.1f} us {fl / t0 / 1e9:.0f} TF/s | [K][N] {t1 * 1e3:.1f} us {fl / t1 / 1e9:.0f} TF/s | rel diff {err:.2e}")

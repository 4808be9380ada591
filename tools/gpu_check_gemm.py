"""GPU bring-up check of the tcgen05 implicit-GEMM kernel (all three modes) against torch fp32 conv.
Prints one line per case; never stops at the first failure.  Run on a B200 via gpurun."""
import ctypes
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from gan_inpainting_b200 import _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))


def pack_conv(w):  # [Cout,Cin,4,4] -> [Cout, 16*Cin], k = (kh*4+kw)*Cin + ci
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def pack_convt(w):  # [Cin,Cout,4,4] -> [4][Cout][4*Cin]
    cin, cout = w.shape[:2]
    out = torch.empty(4, cout, 4, cin, dtype=w.dtype, device=w.device)
    for p in range(2):
        for q in range(2):
            for a in range(2):
                for b in range(2):
                    out[p * 2 + q, :, a * 2 + b, :] = w[:, :, 3 - p - 2 * a, 3 - q - 2 * b].t()
    return out.reshape(4, cout, 4 * cin).contiguous()


def check_conv(B, H, Cin, Cout, bn=0, stages=0):
    x = torch.randn(B, H, H, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, 4, 4, device=dev) / (Cin * 16) ** 0.5).bfloat16()
    y = torch.full((B, H // 2, H // 2, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    wp = pack_conv(w)
    _lib.call("b200_conv_s2_gemm", _lib.ptr(x), B, H, H, Cin, _lib.ptr(wp), Cout, _lib.ptr(y), Cout, bn, stages, None, None, _lib.stream())
    torch.cuda.synchronize()
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    return rel(y, ref)


def check_convt(B, Hi, Cin, Cout, bn=0, stages=0):
    z = torch.randn(B, Hi, Hi, Cin, device=dev).bfloat16()
    w = (torch.randn(Cin, Cout, 4, 4, device=dev) / (Cin * 4) ** 0.5).bfloat16()
    y = torch.full((B, 2 * Hi, 2 * Hi, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    wp = pack_convt(w)
    _lib.call("b200_convt_s2_gemm", _lib.ptr(z), B, Hi, Hi, Cin, Cin, _lib.ptr(wp), Cout, _lib.ptr(y), Cout, bn, stages, None, None, _lib.stream())
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(z.permute(0, 3, 1, 2).float(), w.float(), stride=2, padding=1).permute(0, 2, 3, 1)
    return rel(y, ref)


def check_wgrad(B, Hs, Cs, Cl, splits=4, bn=0, stages=0):
    S = torch.randn(B, Hs, Hs, Cs, device=dev).bfloat16()
    L = torch.randn(B, 2 * Hs, 2 * Hs, Cl, device=dev).bfloat16()
    sp = ctypes.c_int(splits)
    part = torch.full((max(splits, 1), Cs, 16 * Cl), float("nan"), device=dev, dtype=torch.float32)
    _lib.call("b200_wgrad_gemm", _lib.ptr(S), B, Hs, Hs, Cs, Cs, _lib.ptr(L), Cl, _lib.ptr(part), ctypes.byref(sp), bn, stages, None, None, _lib.stream())
    torch.cuda.synchronize()
    g = part[: sp.value].sum(0).view(Cs, 4, 4, Cl).permute(0, 3, 1, 2)  # [Cs,Cl,kh,kw]
    ref = torch.nn.grad.conv2d_weight(L.permute(0, 3, 1, 2).float(), (Cs, Cl, 4, 4), S.permute(0, 3, 1, 2).float(), stride=2, padding=1)
    return rel(g, ref), sp.value


def probe():
    # parity view: x[b,h,w,c] = unique code; load box for tap (kh,kw)=(0,0) at tile origin -> halo must be zero
    B, H, W, C = 2, 8, 8, 64
    x = torch.arange(B * H * W * C, device=dev, dtype=torch.float32).view(B, H, W, C)
    xb = (x % 251).bfloat16()
    out = torch.zeros(128 * 16, dtype=torch.uint8, device=dev)
    tw, th, nb = 4, 4, 1
    kh, kw = 0, 0
    _lib.call("b200_debug_tma_probe", _lib.ptr(xb), B, H, W, C, 0, tw, th, nb, ((kw + 1) & 1) * C, ((kw + 1) >> 1) - 1, (kh + 1) & 1,
              ((kh + 1) >> 1) - 1, 1, _lib.ptr(out), _lib.stream())
    torch.cuda.synchronize()
    raw = out.view(torch.bfloat16).view(16, 64).float()  # 16 rows x 64 ch (swizzled 16B chunks)
    # undo swizzle: chunk j of row r is stored at chunk j ^ (r % 8)
    un = torch.empty_like(raw)
    for r in range(16):
        for j in range(8):
            un[r, j * 8:(j + 1) * 8] = raw[r, (j ^ (r % 8)) * 8:((j ^ (r % 8)) + 1) * 8]
    exp = torch.zeros(16, 64, device=dev)
    for hl in range(th):
        for wl in range(tw):
            ih, iw = 2 * hl + kh - 1, 2 * wl + kw - 1
            if 0 <= ih < H and 0 <= iw < W:
                exp[hl * tw + wl] = xb[1, ih, iw].float()
    ok = torch.equal(un, exp)
    print(f"[probe] parity-view TMA box (tap 0,0, halo zero-fill, swizzle-128B): {'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        print("  got rows (unswizzled) first 4 ch:", un[:, :4].tolist())
        print("  exp rows first 4 ch:", exp[:, :4].tolist())
        print("  raw row0:", raw[0, :16].tolist(), " raw row1:", raw[1, :16].tolist())


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    try:
        probe()
    except Exception as e:  # noqa
        print("[probe] EXC", e, flush=True)
    cases = []
    for name, fn, args in [
        ("conv  B2 H16 64->64", check_conv, (2, 16, 64, 64)),
        ("conv  B2 H32 64->128", check_conv, (2, 32, 64, 128)),
        ("conv  B4 H64 128->256", check_conv, (4, 64, 128, 256)),
        ("conv  B3 H8 256->512", check_conv, (3, 8, 256, 512)),
        ("conv  B5 H4 512->512 bn128", check_conv, (5, 4, 512, 512, 128)),
        ("conv  B8 H2 512->512", check_conv, (8, 2, 512, 512)),
        ("conv  B2 H256 64->128", check_conv, (2, 256, 64, 128)),
        ("conv  B2 H16 64->16 (N16)", check_conv, (2, 16, 64, 16)),
        ("convT B2 Hi8 64->64", check_convt, (2, 8, 64, 64)),
        ("convT B2 Hi1 512->512", check_convt, (2, 1, 512, 512)),
        ("convT B3 Hi4 1024->512", check_convt, (3, 4, 1024, 512)),
        ("convT B2 Hi32 512->128", check_convt, (2, 32, 512, 128)),
        ("convT B2 Hi64 256->64", check_convt, (2, 64, 256, 64)),
        ("convT B1 Hi128 128->16 (N16)", check_convt, (1, 128, 128, 16)),
        ("wgrad B2 Hs8 128x64", check_wgrad, (2, 8, 128, 64)),
        ("wgrad B2 Hs32 128x64 s7", check_wgrad, (2, 32, 128, 64, 7)),
        ("wgrad B3 Hs4 512x256", check_wgrad, (3, 4, 512, 256, 2)),
        ("wgrad B2 Hs1 512x512", check_wgrad, (2, 1, 512, 512, 1)),
        ("wgrad B2 Hs16 256x128 bn128", check_wgrad, (2, 16, 256, 128, 4, 128)),
        ("wgrad B2 Hs64 256x64 (u2)", check_wgrad, (2, 64, 256, 64, 16)),
    ]:
        try:
            t0 = time.time()
            r = fn(*args)
            print(f"[gemm] {name:34s} rel-L2 = {r}  ({time.time() - t0:.2f}s)", flush=True)
        except Exception as e:  # noqa
            print(f"[gemm] {name:34s} EXC {e}", flush=True)


if __name__ == "__main__":
    main()

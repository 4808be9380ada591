"""Host-side orchestration of the CUDA kernels for the two networks on the hot path.

``GeneratorEngine``  — UnetGenerator(1, 1, num_downs, ngf=64, BatchNorm) forward / backward
                        (reference lib/models/networks.py:216-324)
``DiscriminatorEngine`` — PatchGANDiscriminator forward / backward (networks.py:331-363)

The engines own no parameters: they read the fp32 master weights of the drop-in ``nn.Module``s
(checkpoint layout), keep bf16 K-major operand copies for the tensor-core kernels,
and write fp32 gradients into tensors the caller provides.  Activations are NHWC bf16; the
1-channel images at either end stay fp32.  Every op is a call through the C-ABI
(``_lib.call``) on the current CUDA stream; PyTorch is used for device memory only.

Data flow of one generator level ``l`` (the skip tensor is the *post-LeakyReLU*
encoder activation and the decoder applies ReLU to the whole concat):

    a[l-1] --conv k4s2p1 (tcgen05)--> y[l] --BN stats--> a[l] = leaky(bn(y[l]))          (dense)
                                                       cat[l][:, :C] = relu(bn(y[l]))    (skip half)
    cat[l] --convT k4s2p1 (tcgen05, 4 phases)--> yu[l] --BN stats--> cat[l-1][:, C':] = relu(drop(bn(yu[l])))
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import call, ptr, stream

ACT_NONE, ACT_LEAKY, ACT_RELU, ACT_TANH = 0, 1, 2, 3
NUM_SMS = 148

bf16 = torch.bfloat16
f32 = torch.float32


def _empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device)


class Scratch:
    """Per-device scratch shared by all engines (single stream semantics)."""

    _by_device: Dict[int, "Scratch"] = {}

    def __init__(self, device):
        self.device = device
        self.bn_partial = _empty(_lib.value("b200_bn_stat_blocks", 1 << 40, 8) * 2 * 2048, f32, device)
        self.coef = _empty(3 * 2048, f32, device)
        self.c1_partial = _empty(_lib.value("b200_wgrad_c1_blocks") * 128 * 16, f32, device)
        self.recon = _empty(_lib.value("b200_recon_scratch_floats"), f32, device)
        self.wgrad: Optional[torch.Tensor] = None
        self.split: Optional[torch.Tensor] = None

    @classmethod
    def get(cls, device) -> "Scratch":
        idx = torch.device(device).index or 0
        if idx not in cls._by_device:
            cls._by_device[idx] = Scratch(torch.device("cuda", idx))
        return cls._by_device[idx]

    def split_buf(self, nfloats: int) -> torch.Tensor:
        if self.split is None or self.split.numel() < nfloats:
            self.split = _empty(nfloats, f32, self.device)
        return self.split

    def wgrad_buf(self, nfloats: int) -> torch.Tensor:
        if self.wgrad is None or self.wgrad.numel() < nfloats:
            self.wgrad = _empty(nfloats, f32, self.device)
        return self.wgrad


def pick_bn(m_tiles: int, cout: int, phases: int = 1) -> int:
    """N tile of the fprop / dgrad GEMM: the widest tile that still gives every SM a CTA."""
    for bn in (256, 128, 64, 32):
        if cout % bn:
            continue
        if m_tiles * (cout // bn) * phases >= NUM_SMS:
            return bn
    for bn in (32, 64, 128, 256):
        if cout % bn == 0:
            return bn
    return cout


def pick_ksplit(m_tiles: int, cout: int, phases: int, kb_total: int):
    """K split of the innermost layers' fprop / dgrad GEMMs (a handful of 128-pixel tiles, K up to 8192): wide N tiles (least
    operand traffic per flop) and enough K splits to occupy the SMs.  Returns (ksplit, bn); ksplit == 1: no split."""
    bn = 256 if cout % 256 == 0 else (128 if cout % 128 == 0 else 64)
    tiles = m_tiles * (cout // bn) * phases
    if 2 * tiles > NUM_SMS or kb_total < 32:
        return 1, 0
    ks = 1
    while tiles * ks * 2 <= NUM_SMS + tiles and kb_total % (ks * 2) == 0 and kb_total // (ks * 2) >= 8 and (kb_total // (ks * 2)) % 2 == 0:
        ks *= 2
    return (ks, bn) if ks > 1 else (1, 0)


def _m_tiles(B: int, hs: int, ws: int, npix: int = 128) -> int:
    tw = min(ws, npix)
    th = min(hs, npix // tw)
    nb = npix // (tw * th)
    return (ws // tw) * (hs // th) * ((B + nb - 1) // nb)


# ------------------------------------------------------------------------------------------------
# thin op wrappers (shapes are NHWC; all tensors contiguous)
# ------------------------------------------------------------------------------------------------
class BnPlan:
    """Asks the convolution epilogue for the batch statistics of its output (training-mode BatchNorm that follows) and
    finalizes them: mean / rstd / scale / shift in ``stats`` [4, C], running statistics updated in place."""

    def __init__(self, bn, train: bool, scratch: "Scratch"):
        track = bn.track_running_stats and train
        self.bn, self.scratch = bn, scratch
        self.rm, self.rv, self.nbt = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if track else (None, None, None)
        self.rows = ctypes.c_int(0)
        self.precomputed = None  # split-K layers: statistics from the standalone kernel

    def finalize(self, y, C):
        if self.precomputed is not None:
            return self.precomputed
        bn = self.bn
        stats = _empty((4, C), f32, y.device)
        call("b200_bn_finalize", ptr(self.scratch.bn_partial), self.rows.value, y.numel() // C, C, ptr(bn.weight.data), ptr(bn.bias.data), ptr(self.rm),
             ptr(self.rv), ptr(self.nbt), float(bn.momentum), float(bn.eps), ptr(stats), stream())
        return stats


class ActBwd:
    """Fused backward-of-activation epilogue of a data-gradient GEMM (``b200_epi_bwd``): the GEMM's output columns
    ``[col0, col0 + Cy)`` are gradients w.r.t. ``act(dropout(bn(y)))``; the epilogue stores ``dz`` (gradient w.r.t. the
    BatchNorm output) instead and, for a training-mode BatchNorm, leaves the per-CTA (sum dz, sum dz*y) rows in the scratch."""

    def __init__(self, y, Cy, stats, act, col0=0, g2=None, ldg2=0, offg2=0, dropout=False, seed=0, scratch: Optional["Scratch"] = None):
        self.Cy, self.col0, self.has_bn = Cy, col0, stats is not None
        self.y, self.stats, self.act, self.g2, self.ldg2, self.offg2, self.dropout, self.seed = y, stats, act, g2, ldg2, offg2, bool(dropout), int(seed)
        self.raw = None  # split-K layers: the plain gradient, the activation / BatchNorm backward then runs as its own kernels
        # a BatchNorm backward follows: dz is kept unrounded (fp32) for it — one bf16 rounding less per layer of the gradient chain
        self.dz32 = _empty(tuple(y.shape[:-1]) + (Cy,), f32, y.device) if (self.has_bn and scratch is not None) else None
        self.keep = (y, stats, g2)  # the raw pointers below must outlive the launch
        self.c = _lib.EpiBwd(y.data_ptr(), Cy, stats.data_ptr() + 2 * Cy * 4 if stats is not None else None,
                             stats.data_ptr() + 3 * Cy * 4 if stats is not None else None, act, g2.data_ptr() if g2 is not None else None, ldg2, offg2,
                             int(dropout), int(seed) & 0xFFFFFFFF, col0, Cy, self.dz32.data_ptr() if self.dz32 is not None else None)
        self.scratch = scratch
        self.rows = ctypes.c_int(0)

    def args(self):
        want = self.has_bn and self.scratch is not None
        return (ptr(self.scratch.bn_partial) if want else None, ctypes.byref(self.rows) if want else None, ctypes.byref(self.c))


def bn_bwd_fused(y, C, stats, gamma, dz, lddz, offdz, epi: ActBwd, cout_total, dgamma, dbeta, accumulate, scratch: "Scratch"):
    """Finish a BatchNorm backward whose dz and partial sums came out of a GEMM epilogue: dgamma / dbeta, and dy."""
    if epi.raw is not None:  # the GEMM ran split-K without a fused epilogue
        return bn_act_bwd(y, C, stats, gamma, epi.act, epi.raw, cout_total, epi.col0, epi.g2, epi.ldg2, epi.offg2, epi.dropout, epi.seed, dgamma, dbeta,
                          accumulate, scratch)
    P = y.numel() // C
    dy = _empty(y.shape, bf16, y.device)
    part = ctypes.c_void_p(scratch.bn_partial.data_ptr() + epi.col0 * 4)
    if epi.dz32 is not None:  # the epilogue left dz unrounded in its own fp32 tensor
        dz, lddz, offdz, f32in = epi.dz32, C, 0, 1
    else:
        f32in = 0
    call("b200_bn_bwd_fused", ptr(y), P, C, ptr(stats), ptr(gamma), ptr(dz), lddz, offdz, f32in, part, epi.rows.value, 2 * cout_total, cout_total,
         ptr(scratch.coef), ptr(dgamma), ptr(dbeta), int(accumulate), ptr(dy), stream())
    return dy


def _split_epilogue(out, cout, plan, epi, sc):
    """What the fused epilogue would have done, as separate kernels, after a split-K GEMM (tiny innermost layers)."""
    if plan is not None:
        bn = plan.bn
        plan.precomputed = bn_stats(out, cout, bn.weight.data, bn.bias.data, plan.rm, plan.rv, plan.nbt, bn.momentum, bn.eps, sc)
    if epi is not None:
        if epi.has_bn:
            epi.raw = out  # bn_bwd_fused finishes with the standalone BatchNorm backward
        else:  # activation only (col0 == 0): return dz in place of the plain gradient
            return bn_act_bwd(epi.y, epi.Cy, None, None, epi.act, out, cout, epi.col0, epi.g2, epi.ldg2, epi.offg2, False, 0, None, None, False, sc)
    return out


def conv_s2(x, wpk, cout, out=None, ldy=None, plan: Optional[BnPlan] = None, epi: Optional[ActBwd] = None):
    """x [B,H,W,Cin] bf16 -> [B,H/2,W/2,cout]; wpk [cout][16*Cin]."""
    B, H, W, Cin = x.shape
    if out is None:
        out = _empty((B, H // 2, W // 2, cout), bf16, x.device)
        ldy = cout
    mt = _m_tiles(B, H // 2, W // 2)
    # split-K pays only when nothing is fused into the epilogue (inference: eval-mode BatchNorm, no backward): in training the
    # separate reduce + statistics / activation-backward kernels of these tiny layers cost what the faster GEMM saves (measured)
    ks, bn = pick_ksplit(mt, cout, 1, 16 * (Cin // 64)) if (ldy == cout and plan is None and epi is None) else (1, 0)
    if ks > 1:
        sc = Scratch.get(x.device)
        P = B * (H // 2) * (W // 2)
        part = sc.split_buf(ks * P * cout)
        call("b200_conv_s2_gemm", ptr(x), B, H, W, Cin, ptr(wpk), cout, ptr(out), ldy, bn, 0, None, None, None, ks, ptr(part), stream())
        call("b200_splitk_reduce", ptr(part), ks, P, cout, ptr(out), ldy, stream())
        return _split_epilogue(out, cout, plan, epi, sc)
    bn = pick_bn(mt, cout)
    extra = epi.args() if epi is not None else (ptr(plan.scratch.bn_partial) if plan else None, ctypes.byref(plan.rows) if plan else None, None)
    call("b200_conv_s2_gemm", ptr(x), B, H, W, Cin, ptr(wpk), cout, ptr(out), ldy, bn, 0, *extra, 1, None, stream())
    return out


def convt_s2(z, ldz, cin, wpk, cout, plan: Optional[BnPlan] = None, epi: Optional[ActBwd] = None):
    """z [B,Hi,Wi,>=cin] (pixel stride ldz) -> [B,2Hi,2Wi,cout]; wpk [4][cout][4*cin]."""
    B, Hi, Wi = z.shape[:3]
    out = _empty((B, 2 * Hi, 2 * Wi, cout), bf16, z.device)
    mt = _m_tiles(B, Hi, Wi)
    ks, bn = pick_ksplit(mt, cout, 4, 4 * (cin // 64)) if (plan is None and epi is None) else (1, 0)
    if ks > 1:
        sc = Scratch.get(z.device)
        P = B * 4 * Hi * Wi
        part = sc.split_buf(ks * P * cout)
        call("b200_convt_s2_gemm", ptr(z), B, Hi, Wi, cin, ldz, ptr(wpk), cout, ptr(out), cout, bn, 0, None, None, None, ks, ptr(part), stream())
        call("b200_splitk_reduce", ptr(part), ks, P, cout, ptr(out), cout, stream())
        return _split_epilogue(out, cout, plan, epi, sc)
    bn = pick_bn(mt, cout, phases=4)
    extra = epi.args() if epi is not None else (ptr(plan.scratch.bn_partial) if plan else None, ctypes.byref(plan.rows) if plan else None, None)
    call("b200_convt_s2_gemm", ptr(z), B, Hi, Wi, cin, ldz, ptr(wpk), cout, ptr(out), cout, bn, 0, *extra, 1, None, stream())
    return out


def pick_splits(ptiles: int, ctas: int) -> int:
    """Pixel-range splits of a weight-gradient GEMM for the persistent kernel (one CTA per SM, static tile schedule).
    ``ctas`` output tiles exist per split and ``ptiles`` 64-pixel k-blocks in total.  Cost model in k-block units:
    rounds of NUM_SMS concurrent tiles x (k-blocks per tile + a fixed per-tile cost) + the partial-sum traffic (one fp32
    tile written and re-read per split ~ 6 k-blocks of operand traffic)."""
    best, best_cost = 1, None
    for s in range(1, min(ptiles, 64) + 1):
        kb = -(-ptiles // s)
        s_eff = -(-ptiles // kb)
        rounds = -(-(ctas * s_eff) // NUM_SMS)
        cost = rounds * (kb + 6) + 0.05 * s_eff * ctas * 6 / NUM_SMS
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = s_eff, cost
    return best


def wgrad(S, lds, cs, Lt, grad, accumulate, scratch: Scratch):
    """grad[cs][cl][4][4] (+)= sum S[b,i,j,cs] * L[b,2i+kh-1,2j+kw-1,cl]."""
    B, Hs, Ws = S.shape[:3]
    cl = Lt.shape[3]
    ptiles = _m_tiles(B, Hs, Ws, 64)
    bn = 256
    ctas = (cs // 128) * (16 * cl // bn)
    splits = pick_splits(ptiles, ctas)
    part = scratch.wgrad_buf(splits * cs * 16 * cl)
    sp = ctypes.c_int(splits)
    call("b200_wgrad_gemm", ptr(S), B, Hs, Ws, cs, lds, ptr(Lt), cl, ptr(part), ctypes.byref(sp), bn, 0, stream())
    call("b200_wgrad_reduce", ptr(part), sp.value, cs, cl, ptr(grad), int(accumulate), stream())


def bn_stats(y, C, gamma, beta, rmean, rvar, nbt, momentum, eps, scratch: Scratch):
    P = y.numel() // C
    stats = _empty((4, C), f32, y.device)
    call("b200_bn_stats", ptr(y), P, C, ptr(scratch.bn_partial), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), ptr(nbt), float(momentum), float(eps),
         ptr(stats), stream())
    return stats


def bn_eval(C, gamma, beta, rmean, rvar, eps, device):
    stats = _empty((4, C), f32, device)
    call("b200_bn_eval_coeffs", C, ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), float(eps), ptr(stats), stream())
    return stats


def bn_act(y, C, stats, act, out1=None, ld1=0, off1=0, out2=None, ld2=0, off2=0, dropout=False, seed=0):
    P = y.numel() // C
    scale = None if stats is None else ctypes.c_void_p(stats.data_ptr() + 2 * C * 4)
    shift = None if stats is None else ctypes.c_void_p(stats.data_ptr() + 3 * C * 4)
    call("b200_bn_act", ptr(y), P, C, scale, shift, act, ptr(out1), ld1, off1, ptr(out2), ld2, off2, int(dropout), int(seed) & 0xFFFFFFFF, stream())


def bn_act_bwd(y, C, stats, gamma, act, g1, ldg1, offg1, g2, ldg2, offg2, dropout, seed, dgamma, dbeta, accumulate, scratch: Scratch):
    P = y.numel() // C
    dy = _empty(y.shape, bf16, y.device)
    call("b200_bn_act_bwd", ptr(y), P, C, ptr(stats), ptr(gamma), act, ptr(g1), ldg1, offg1, ptr(g2), ldg2, offg2, int(dropout), int(seed) & 0xFFFFFFFF,
         ptr(scratch.bn_partial), ptr(scratch.coef), ptr(dgamma), ptr(dbeta), int(accumulate), ptr(dy), stream())
    return dy


class _Packed:
    """bf16 operand copies of one 4x4/s2 layer's fp32 master weight [D0, D1, 4, 4]."""

    def __init__(self, w: torch.Tensor):
        d0, d1 = w.shape[:2]
        self.gather = _empty((d0, 16 * d1), bf16, w.device)      # n = D0, k = D1
        self.phase = _empty((4, d1, 4 * d0), bf16, w.device)     # n = D1, k = D0

    def refresh(self, w: torch.Tensor):
        d0, d1 = w.shape[:2]
        call("b200_pack_gather", ptr(w), d0, d1, ptr(self.gather), stream())
        call("b200_pack_phase", ptr(w), d0, d1, ptr(self.phase), stream())


class PackTable:
    """Device-side work list of ``b200_pack_all``: refreshes the bf16 operand copies of all 4x4 layers of a network with one launch.
    Rebuilt only when a weight or packed buffer moves (its addresses are baked into the table)."""

    def __init__(self):
        self.key, self.table, self.n, self.tiles = None, None, 0, 0

    def refresh(self, pairs):
        """pairs: [(fp32 weight [D0, D1, 4, 4], _Packed)]"""
        key = tuple((w.data_ptr(), pk.gather.data_ptr(), pk.phase.data_ptr()) for w, pk in pairs)
        if key != self.key:
            rows, tile0 = [], 0
            for w, pk in pairs:
                d0, d1 = w.shape[:2]
                if d0 % 16 or d1 % 32:
                    raise _lib.B200Error("pack_all: channel counts must be multiples of 16 / 32")
                rows.append([w.data_ptr(), pk.gather.data_ptr(), pk.phase.data_ptr(), d0, d1, tile0])
                tile0 += (d0 // 16) * (d1 // 32)
            self.table = torch.tensor(rows, dtype=torch.int64, device=pairs[0][0].device)
            self.key, self.n, self.tiles = key, len(rows), tile0
        call("b200_pack_all", ptr(self.table), self.n, self.tiles, stream())


# ================================================================================================
# generator
# ================================================================================================
class GenCtx:
    """Everything one generator forward leaves behind for its backward."""

    __slots__ = ("B", "R", "x", "a", "y", "sd", "r_in", "yu", "su", "cat", "g_out", "comp", "seed", "train", "dropout")


class GeneratorEngine:
    def __init__(self, down_convs, down_norms, up_convs, up_norms, dropout_levels, widths):
        """``down_convs[l-1]`` / ``up_convs[l-1]`` are the nn.Conv2d / nn.ConvTranspose2d of level l (1 = outermost);
        ``down_norms`` / ``up_norms`` hold the nn.BatchNorm2d of that level or None."""
        self.N = len(down_convs)
        self.dc, self.dn, self.uc, self.un = down_convs, down_norms, up_convs, up_norms
        self.w = [None] + list(widths)  # w[l] = channels of a[l]
        self.dropout_levels = set(dropout_levels)
        self.packed_d: Dict[int, _Packed] = {}
        self.packed_u: Dict[int, _Packed] = {}
        self._pack_table = PackTable()
        self._versions = None
        self.calls = 0
        for l in range(2, self.N + 1):
            if self.w[l] % 128 or self.w[l - 1] % 64:
                raise _lib.B200Error("the tcgen05 path needs ngf to be a multiple of 64 (channel plan %r)" % (widths,))

    # ---- weights
    def _weights(self):
        return [c.weight for c in self.dc[1:]] + [c.weight for c in self.uc[1:]]

    def refresh_weights(self, force=False):
        ws = self._weights()
        ver = (_lib.WEIGHT_EPOCH,) + tuple((w._version, w.data_ptr()) for w in ws)
        if not force and ver == self._versions:
            return
        pairs = []
        for l in range(2, self.N + 1):
            if l not in self.packed_d or self.packed_d[l].gather.device != self.dc[l - 1].weight.device:
                self.packed_d[l] = _Packed(self.dc[l - 1].weight)
                self.packed_u[l] = _Packed(self.uc[l - 1].weight)
            pairs += [(self.dc[l - 1].weight.data, self.packed_d[l]), (self.uc[l - 1].weight.data, self.packed_u[l])]
        self._pack_table.refresh(pairs)
        self._versions = ver

    # ---- forward
    def forward(self, x: torch.Tensor, train: bool, comp_base=None, comp_mul=None, seed: int = 0, keep: bool = True) -> GenCtx:
        """x: fp32 [B,1,R,R] (or [B,R,R]).  Returns a context with ``g_out`` fp32 [B,1,R,R] (and ``comp``)."""
        dev = x.device
        B, R = x.shape[0], x.shape[-1]
        N, w = self.N, self.w
        if x.shape[-2] != R or R % (1 << N) or (R >> N) < 1 or (R & (R - 1)):
            raise _lib.B200Error(f"generator input must be square with a power-of-two side >= {1 << N}, got {tuple(x.shape)}")
        if R // 2 < 16:
            raise _lib.B200Error("generator input side must be >= 32")
        self.refresh_weights()
        sc = Scratch.get(dev)
        c = GenCtx()
        c.B, c.R, c.x, c.train, c.seed = B, R, x, train, seed
        c.dropout = bool(self.dropout_levels) and train
        c.a, c.y, c.sd, c.yu, c.su, c.cat = {}, {}, {}, {}, {}, {}
        side = lambda l: R >> l
        for l in range(1, N):
            c.cat[l] = _empty((B, side(l), side(l), 2 * w[l]), bf16, dev)
        # d1: direct conv + LeakyReLU, relu copy into cat[1]
        c.a[1] = _empty((B, side(1), side(1), w[1]), bf16, dev)
        call("b200_conv_c1_fwd", ptr(x), None, B, R, R, ptr(self.dc[0].weight.data), w[1], ACT_LEAKY, ptr(c.a[1]), ptr(c.cat[1]), 2 * w[1], 0, stream())
        # d2 .. dN
        for l in range(2, N + 1):
            plan = self._plan(self.dn[l - 1], train, sc) if l < N else None
            c.y[l] = conv_s2(c.a[l - 1], self.packed_d[l].gather, w[l], plan=plan)
            if l < N:
                c.sd[l] = self._stats(self.dn[l - 1], c.y[l], w[l], train, plan)
                c.a[l] = _empty(c.y[l].shape, bf16, dev)
                bn_act(c.y[l], w[l], c.sd[l], ACT_LEAKY, c.a[l], w[l], 0, c.cat[l], 2 * w[l], 0)
            else:
                c.r_in = _empty(c.y[l].shape, bf16, dev)
                bn_act(c.y[l], w[l], None, ACT_RELU, c.r_in, w[l], 0)
        # uN .. u2
        for l in range(N, 1, -1):
            zin, cin = (c.r_in, w[N]) if l == N else (c.cat[l], 2 * w[l])
            plan = self._plan(self.un[l - 1], train, sc)
            c.yu[l] = convt_s2(zin, cin, cin, self.packed_u[l].phase, w[l - 1], plan=plan)
            c.su[l] = self._stats(self.un[l - 1], c.yu[l], w[l - 1], train, plan)
            bn_act(c.yu[l], w[l - 1], c.su[l], ACT_RELU, c.cat[l - 1], 2 * w[l - 1], w[l - 1], dropout=c.dropout and l in self.dropout_levels,
                   seed=seed * 16 + l)
        # u1: direct transposed conv + bias + tanh (+ composite)
        c.g_out = _empty((B, 1, R, R), f32, dev)
        c.comp = _empty((B, 1, R, R), f32, dev) if comp_base is not None else None
        up1 = self.uc[0]
        call("b200_convt_c1_fwd", ptr(c.cat[1]), B, side(1), side(1), 2 * w[1], 2 * w[1], ptr(up1.weight.data), ptr(up1.bias.data if up1.bias is not None else None),
             ACT_TANH, ptr(comp_base), ptr(comp_mul), ptr(c.g_out), ptr(c.comp), stream())
        self.calls += 1
        if not keep:  # inference: drop the activations
            c.a = c.y = c.yu = c.cat = None
        return c

    @staticmethod
    def _plan(bn, train, sc):
        """Batch statistics are needed in training mode (and always when the norm keeps no running statistics)."""
        if bn is None or not (train or not bn.track_running_stats):
            return None
        return BnPlan(bn, train, sc)

    @staticmethod
    def _stats(bn, y, C, train, plan):
        if bn is None:
            return None
        if plan is not None:
            return plan.finalize(y, C)
        return bn_eval(C, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.eps, y.device)

    # ---- backward
    def backward(self, c: GenCtx, dpre: torch.Tensor, grads: Dict[object, torch.Tensor], accumulate: bool = False, on_ready=None, trace=None):
        """dpre: fp32 [B,1,R,R] gradient w.r.t. the pre-tanh output.  ``grads`` maps each parameter (by
        ``id`` of the nn.Parameter) to the fp32 tensor its gradient is written / accumulated into; the bias of
        the last layer is handled by the caller (gstep_seed reduces it)."""
        if not c.train:
            raise _lib.B200Error("generator backward needs a training-mode forward (batch statistics)")
        dev = dpre.device
        sc = Scratch.get(dev)
        B, R, N, w = c.B, c.R, self.N, self.w
        side = lambda l: R >> l
        acc = int(accumulate)
        G = lambda p: grads[id(p)]
        done = on_ready if on_ready is not None else (lambda p: None)  # data-parallel bucketing hook: p.grad is enqueued
        # ---- u1
        up1 = self.uc[0]
        call("b200_wgrad_c1", ptr(c.cat[1]), B, side(1), side(1), 2 * w[1], 2 * w[1], ptr(dpre), None, ptr(sc.c1_partial), ptr(G(up1.weight)), acc, stream())
        done(up1.weight)
        dcat = _empty(c.cat[1].shape, bf16, dev)
        call("b200_conv_c1_fwd", ptr(dpre), None, B, R, R, ptr(up1.weight.data), 2 * w[1], ACT_NONE, ptr(dcat), None, 0, 0, stream())
        dcats = {1: dcat}
        # ---- decoder u2 .. uN.  The data-gradient GEMM of level l also evaluates the activation backward of what feeds its
        # up half (relu(drop(bn(yu[l+1]))), or relu(y[N]) for the innermost level) and the BatchNorm reductions.
        epi = None
        for l in range(2, N + 1):
            bn = self.un[l - 1]
            drop = c.dropout and l in self.dropout_levels
            if epi is None:  # yu[2]: its gradient comes from the 1-channel kernel above
                dyu = bn_act_bwd(c.yu[l], w[l - 1], c.su[l], bn.weight.data, ACT_RELU, dcats[l - 1], 2 * w[l - 1], w[l - 1], None, 0, 0, drop, c.seed * 16 + l,
                                 G(bn.weight), G(bn.bias), accumulate, sc)
            else:
                dyu = bn_bwd_fused(c.yu[l], w[l - 1], c.su[l], bn.weight.data, dcats[l - 1], 2 * w[l - 1], w[l - 1], epi, 2 * w[l - 1], G(bn.weight), G(bn.bias),
                                   accumulate, sc)
            done(bn.bias)
            zin, cin = (c.r_in, w[N]) if l == N else (c.cat[l], 2 * w[l])
            wgrad(zin, cin, cin, dyu, G(self.uc[l - 1].weight), accumulate, sc)
            done(self.uc[l - 1].weight)
            if l < N:
                epi = ActBwd(c.yu[l + 1], w[l], c.su[l + 1], ACT_RELU, col0=w[l], dropout=c.dropout and (l + 1) in self.dropout_levels,
                             seed=c.seed * 16 + l + 1, scratch=sc)
            else:
                epi = ActBwd(c.y[N], w[N], None, ACT_RELU)
            dcats[l] = conv_s2(dyu, self.packed_u[l].gather, cin, epi=epi)
            if trace is not None:
                trace[f"dyu{l}"], trace[f"dcat{l}"] = dyu, dcats[l]
        # ---- encoder dN .. d2: dcats[N] already is the gradient w.r.t. y[N]
        dy = dcats[N]
        for l in range(N, 1, -1):
            if l < N:
                bn = self.dn[l - 1]
                dy = bn_bwd_fused(c.y[l], w[l], c.sd[l], bn.weight.data, dz, w[l], 0, epi, w[l], G(bn.weight), G(bn.bias), accumulate, sc)
                done(bn.weight)
            wgrad(dy, w[l], w[l], c.a[l - 1], G(self.dc[l - 1].weight), accumulate, sc)
            done(self.dc[l - 1].weight)
            # gradient w.r.t. a[l-1] = leaky(bn(y[l-1])), merged with the skip gradient (through relu of the same tensor)
            if l - 1 >= 2:
                epi = ActBwd(c.y[l - 1], w[l - 1], c.sd[l - 1], ACT_LEAKY, g2=dcats[l - 1], ldg2=2 * w[l - 1], offg2=0, scratch=sc)
            else:
                epi = ActBwd(c.a[1], w[1], None, ACT_LEAKY, g2=dcats[1], ldg2=2 * w[1], offg2=0)
            dz = convt_s2(dy, w[l], w[l], self.packed_d[l].phase, w[l - 1], epi=epi)
            if trace is not None:
                trace[f"dy{l}"], trace[f"dz{l - 1}"] = dy, dz
        # ---- d1
        call("b200_wgrad_c1", ptr(dz), B, side(1), side(1), w[1], w[1], ptr(c.x), None, ptr(sc.c1_partial), ptr(G(self.dc[0].weight)), acc, stream())
        done(self.dc[0].weight)


# ================================================================================================
# discriminator
# ================================================================================================
class DiscCtx:
    __slots__ = ("B", "R", "x", "xmul", "a", "y", "st", "lin", "pred", "train", "sigmoid")


class DiscriminatorEngine:
    def __init__(self, convs, norms, head_conv, head_linear):
        """convs: the four k4s2p1 nn.Conv2d; norms: [None, bn2, bn3, bn4]; head_conv: Conv2d(512,1,4,1,0); head_linear: nn.Linear."""
        self.convs, self.norms, self.hc, self.hl = convs, norms, head_conv, head_linear
        self.packed: Dict[int, _Packed] = {}
        self._pack_table = PackTable()
        self._versions = None
        self.E = None
        self.calls = 0

    def refresh_weights(self, force=False):
        ws = [c.weight for c in self.convs[1:]] + [self.hc.weight, self.hl.weight]
        ver = (_lib.WEIGHT_EPOCH,) + tuple((w._version, w.data_ptr()) for w in ws)
        if not force and ver == self._versions:
            return
        pairs = []
        for i in range(1, 4):
            if i not in self.packed or self.packed[i].gather.device != self.convs[i].weight.device:
                self.packed[i] = _Packed(self.convs[i].weight)
            pairs.append((self.convs[i].weight.data, self.packed[i]))
        self._pack_table.refresh(pairs)
        n = int(round(self.hl.in_features ** 0.5))
        S, C = n + 3, self.hc.in_channels
        if self.E is None or self.E.numel() != S * S * C:
            self.E = _empty((S, S, C), f32, self.hc.weight.device)
        call("b200_head_weight", ptr(self.hc.weight.data), ptr(self.hl.weight.data), S, C, ptr(self.E), stream())
        self._versions = ver

    def forward(self, x: torch.Tensor, train: bool, sigmoid: bool, xmul: Optional[torch.Tensor] = None) -> DiscCtx:
        """x: fp32 [B,1,R,R]; xmul: optional fp32 multiplier applied on load (the local discriminator's mask)."""
        dev = x.device
        B, R = x.shape[0], x.shape[-1]
        n = int(round(self.hl.in_features ** 0.5))
        if x.shape[-2] != R or R & (R - 1) or R // 16 - 3 != n:
            raise _lib.B200Error(f"discriminator head Linear({self.hl.in_features},1) expects {16 * (n + 3)}x{16 * (n + 3)} inputs, got {tuple(x.shape)}")
        self.refresh_weights()
        sc = Scratch.get(dev)
        c = DiscCtx()
        c.B, c.R, c.x, c.xmul, c.train, c.sigmoid = B, R, x, xmul, train, sigmoid
        c.a, c.y, c.st = {}, {}, {}
        c.a[0] = _empty((B, R // 2, R // 2, 64), bf16, dev)
        call("b200_conv_c1_fwd", ptr(x), ptr(xmul), B, R, R, ptr(self.convs[0].weight.data), 64, ACT_LEAKY, ptr(c.a[0]), None, 0, 0, stream())
        for i in range(1, 4):
            C = 64 << i
            bn = self.norms[i]
            plan = BnPlan(bn, train, sc) if train else None
            c.y[i] = conv_s2(c.a[i - 1], self.packed[i].gather, C, plan=plan)
            if train:
                c.st[i] = plan.finalize(c.y[i], C)
            else:
                c.st[i] = bn_eval(C, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.eps, dev)
            c.a[i] = _empty(c.y[i].shape, bf16, dev)
            bn_act(c.y[i], C, c.st[i], ACT_LEAKY, c.a[i], C, 0)
        S = R // 16
        c.lin = _empty((B,), f32, dev)
        c.pred = _empty((B, 1), f32, dev)
        hs = _empty((B * 8,), f32, dev)
        call("b200_head_fwd", ptr(c.a[3]), ptr(self.E), B, S * S * 512, ptr(self.hl.bias.data), int(sigmoid), ptr(hs), ptr(c.lin), ptr(c.pred), stream())
        self.calls += 1
        return c

    def backward(self, c: DiscCtx, dlin: torch.Tensor, grads: Optional[Dict[object, torch.Tensor]], accumulate: bool = False, input_grad: bool = False,
                 on_ready=None):
        """dlin: fp32 [B] gradient w.r.t. the pre-sigmoid head output.  ``grads`` None == parameters frozen
        (util.set_requires_grad(D, False), wgan_l1.py:169): only the data gradient is computed."""
        if not c.train:
            raise _lib.B200Error("discriminator backward needs a training-mode forward")
        dev = dlin.device
        sc = Scratch.get(dev)
        B, R = c.B, c.R
        S = R // 16
        acc = int(accumulate)
        G = (lambda p: grads[id(p)]) if grads is not None else None
        done = on_ready if on_ready is not None else (lambda p: None)
        if G is not None:
            dE = _empty((S * S * 512,), f32, dev)
            call("b200_head_bwd_params", ptr(dlin), ptr(c.a[3]), B, S, 512, ptr(self.hc.weight.data), ptr(self.hl.weight.data), ptr(dE), ptr(G(self.hc.weight)),
                 ptr(G(self.hl.weight)), ptr(G(self.hl.bias)), acc, stream())
            done(self.hc.weight)
        da = _empty(c.a[3].shape, bf16, dev)
        call("b200_head_bwd_input", ptr(dlin), ptr(self.E), B, S * S * 512, ptr(da), stream())
        epi = None
        for i in range(3, 0, -1):
            C = 64 << i
            bn = self.norms[i]
            if epi is None:  # y[3]: its gradient comes from the head kernel above
                dy = bn_act_bwd(c.y[i], C, c.st[i], bn.weight.data, ACT_LEAKY, da, C, 0, None, 0, 0, False, 0, G(bn.weight) if G else None,
                                G(bn.bias) if G else None, accumulate, sc)
            else:
                dy = bn_bwd_fused(c.y[i], C, c.st[i], bn.weight.data, da, C, 0, epi, C, G(bn.weight) if G else None, G(bn.bias) if G else None, accumulate, sc)
            if G is not None:
                wgrad(dy, C, C, c.a[i - 1], G(self.convs[i].weight), accumulate, sc)
                done(self.convs[i].weight)
            # data gradient with the LeakyReLU (and BatchNorm reductions) of the layer below evaluated in the GEMM epilogue
            epi = ActBwd(c.y[i - 1], C // 2, c.st[i - 1], ACT_LEAKY, scratch=sc) if i - 1 >= 1 else ActBwd(c.a[0], 64, None, ACT_LEAKY)
            da = convt_s2(dy, C, C, self.packed[i].phase, C // 2, epi=epi)
        dz = da
        if G is not None:
            call("b200_wgrad_c1", ptr(dz), B, R // 2, R // 2, 64, 64, ptr(c.x), ptr(c.xmul), ptr(sc.c1_partial), ptr(G(self.convs[0].weight)), acc, stream())
            done(self.convs[0].weight)
        dx = None
        if input_grad:
            dx = _empty((B, 1, R, R), f32, dev)
            call("b200_convt_c1_fwd", ptr(dz), B, R // 2, R // 2, 64, 64, ptr(self.convs[0].weight.data), None, ACT_NONE, None, ptr(c.xmul), ptr(dx), None, stream())
        return dx


# ================================================================================================
# VGG19 feature extractor of the perceptual / style losses (config 4)
# ================================================================================================
VGG_FEATURE_LAYERS = (1, 6, 11, 20, 29)  # relu1_1 .. relu5_1 (lib/models/loss.py:54)


class VggFeatureEngine:
    """``vgg19.features[0..29]`` (networks.py:367-389: torchvision VGG19 inside ``VGG19Wrapper``) on 1-channel images that the loss
    functions feed as ``img.repeat(1, 3, 1, 1)`` (lib/models/loss.py:55-56).  Forward only — the reference evaluates these losses
    under ``torch.no_grad()`` on detached tensors.  Layers 30..36 of ``features`` are executed by the reference's
    loop but none of their outputs is used; they are skipped.

    Every 3x3 convolution (+ bias + ReLU) except the first runs on the tcgen05 window kernel (``b200_conv3_gemm``); the first one
    (3 identical input channels == 1 channel with channel-summed filters) is a direct kernel; max-pools are their own pass."""

    def __init__(self, features: torch.nn.Sequential):
        self.features = features
        self.plan = []  # ("conv", index) | ("pool",)
        for i, m in enumerate(features):
            if i > max(VGG_FEATURE_LAYERS):
                break
            if isinstance(m, torch.nn.Conv2d):
                if m.kernel_size != (3, 3) or m.stride != (1, 1) or m.padding != (1, 1):
                    raise _lib.B200Error("VggFeatureEngine expects 3x3 / stride 1 / pad 1 convolutions")
                self.plan.append(("conv", i))
            elif isinstance(m, torch.nn.MaxPool2d):
                self.plan.append(("pool", i))
            elif isinstance(m, torch.nn.ReLU):
                self.plan.append(("relu", i))
            else:
                raise _lib.B200Error(f"unexpected module in vgg19.features: {type(m).__name__}")
        self.packed: Dict[int, torch.Tensor] = {}
        self.w1 = None
        self._versions = None
        self.fuse_pool = True  # tests switch it off to compare with the separate max-pool kernel

    def refresh_weights(self):
        convs = [self.features[i] for kind, i in self.plan if kind == "conv"]
        ver = tuple((c.weight._version, c.weight.data_ptr()) for c in convs)
        if ver == self._versions:
            return
        first = True
        for kind, i in self.plan:
            if kind != "conv":
                continue
            c = self.features[i]
            w = c.weight.data.float().contiguous()
            if first:  # Conv2d(3, 64): the three input channels carry the same image
                if w.shape[0] != 64 or w.shape[1] != 3:
                    raise _lib.B200Error("VggFeatureEngine expects features[0] = Conv2d(3, 64, 3, 1, 1)")
                self.w1 = w.sum(dim=1).reshape(64, 9).contiguous()
                first = False
                continue
            cout, cin = w.shape[:2]
            pk = self.packed.get(i)
            if pk is None or pk.device != w.device:
                pk = _empty((cout, 9 * cin), bf16, w.device)
                self.packed[i] = pk
            call("b200_pack_conv3", ptr(w), cout, cin, ptr(pk), stream())
        self._versions = ver

    @torch.no_grad()
    def features_of(self, x: torch.Tensor) -> List[torch.Tensor]:
        """x: fp32 [B,1,H,W] (or [B,H,W]) -> the five NHWC bf16 feature maps relu1_1 .. relu5_1."""
        if not x.is_cuda:
            raise _lib.B200Error("the VGG feature losses run on CUDA tensors only; there is no CPU fallback")
        x = x.detach().contiguous().float()
        B, H, W = x.shape[0], x.shape[-2], x.shape[-1]
        if x.numel() != B * H * W:
            raise _lib.B200Error(f"the VGG feature losses expect 1-channel images, got {tuple(x.shape)}")
        self.refresh_weights()
        outs, h, first = [], None, True
        C = 0
        fused_pools = set()
        for n, (kind, i) in enumerate(self.plan):
            if kind == "conv":
                c = self.features[i]
                bias = c.bias.data.float() if c.bias is not None else None
                if first:
                    h = _empty((B, H, W, 64), bf16, x.device)
                    call("b200_conv3_c1_fwd", ptr(x), B, H, W, ptr(self.w1), ptr(bias), 64, ptr(h), stream())
                    C, first = 64, False
                else:
                    cout = c.out_channels
                    # conv -> ReLU -> MaxPool whose un-pooled activation nobody reads: the pool runs in the GEMM epilogue
                    pool = (n + 2 < len(self.plan) and self.plan[n + 1][0] == "relu" and self.plan[n + 2][0] == "pool"
                            and self.plan[n + 1][1] not in VGG_FEATURE_LAYERS and self.fuse_pool)
                    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
                    y = _empty((B, Ho, Wo, cout), bf16, x.device)
                    mt = B * (H // 16) * (W // 8)
                    call("b200_conv3_gemm", ptr(h), B, H, W, C, C, ptr(self.packed[i]), ptr(bias), cout, ptr(y), cout, int(pool), pick_bn(mt, cout), stream())
                    h, C, H, W = y, cout, Ho, Wo
                    if pool:
                        fused_pools.add(self.plan[n + 2][1])
            elif kind == "pool" and i not in fused_pools:
                y = _empty((B, H // 2, W // 2, C), bf16, x.device)
                call("b200_maxpool2", ptr(h), B, H, W, C, ptr(y), stream())
                h, H, W = y, H // 2, W // 2
            # "relu": fused into the convolution epilogue (ReLU(inplace=True) overwrites the conv output in the reference as well)
            if i in VGG_FEATURE_LAYERS:
                outs.append(h)
        return outs

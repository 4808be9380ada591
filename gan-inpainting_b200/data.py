"""Host -> device input pipeline of the train step.

The reference moves each batch with a blocking ``.to(device)`` inside the batch loop (experiment_list/wgan_l1.py:113-114,
``DataLoader(num_workers=0)`` train.py:77), so the GPU idles during every copy.  ``DevicePrefetcher`` keeps the copy of batch
i+1 (pinned host memory -> device, on a side stream) in flight while batch i is being trained on; the tuple contract of the
loader — ``(ground, mask, segment)`` with ``ground`` / ``mask`` fp32 ``[B, 1, R, R]`` (lib/data/dataset.py:14-51) — is kept.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Iterator, Tuple

import torch


class DevicePrefetcher:
    """Iterates ``loader`` one batch ahead: ``for ground, mask, extra in DevicePrefetcher(loader, device)`` yields CUDA tensors
    whose host->device copies overlapped the previous iteration's kernels.  Two device buffers per tensor are recycled; a buffer
    is overwritten only after the step that consumed it has been enqueued AND a newer batch was requested (event-guarded)."""

    def __init__(self, loader: Iterable, device, pin: bool = True, transform=None):
        """``transform``: optional ``GpuLoadTransform``: the loader may then yield DECODED uint8 images (``[B,H,W,3]`` / ``[B,H,W]``); they
        cross PCIe as bytes and become the reference's fp32 ``[B,1,R,R]`` tensors (greyscale + Resize + ToTensor) on the device."""
        self.loader, self.device, self.pin, self.transform = loader, torch.device(device), pin, transform
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [None, None]
        self.consumed = [None, None]

    def _to_device(self, batch, slot: int):
        ground, mask = batch[0], batch[1]
        rest = tuple(batch[2:])
        if self.pin and not ground.is_pinned():
            ground, mask = ground.pin_memory(), mask.pin_memory()
        cur = self.slots[slot]
        if cur is None or cur[0].shape != ground.shape or cur[1].shape != mask.shape or cur[0].dtype != ground.dtype:
            keep = lambda t: t.dtype if t.dtype == torch.uint8 else torch.float32
            cur = (torch.empty(ground.shape, dtype=keep(ground), device=self.device), torch.empty(mask.shape, dtype=keep(mask), device=self.device))
            self.slots[slot] = cur
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[slot] is not None:
                self.copy_stream.wait_event(self.consumed[slot])  # the step that read this buffer has finished
            cur[0].copy_(ground, non_blocking=True)
            cur[1].copy_(mask, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        return cur, rest, ready

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, object]]:
        it = iter(self.loader)
        try:
            nxt = self._to_device(next(it), 0)
        except StopIteration:
            return
        i = 0
        while nxt is not None:
            (ground, mask), rest, ready = nxt
            torch.cuda.current_stream(self.device).wait_event(ready)
            if self.transform is not None and ground.dtype == torch.uint8:
                ground, mask = self.transform(ground), self.transform(mask)
            try:
                nxt = self._to_device(next(it), (i + 1) & 1)
            except StopIteration:
                nxt = None
            yield (ground, mask) + rest
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))  # everything the consumer enqueued for this batch
            self.consumed[i & 1] = done
            i += 1


def synthetic_batch(batch: int, res: int, seed: int = 1234, rank: int = 0):
    """The synthetic inputs of the benchmark, in the loader's format (lib/data/dataset.py:14-51): ``ground`` ~ U[0,1)
    fp32 ``[B,1,R,R]``, ``mask`` = one axis-aligned rectangle of ones per image with sides in [R/8, R/2].  Host tensors; seeded per rank."""
    g = torch.Generator().manual_seed(seed + rank)
    ground = torch.rand(batch, 1, res, res, generator=g, dtype=torch.float32)
    mask = torch.zeros(batch, 1, res, res, dtype=torch.float32)
    lo, hi = res // 8, res // 2
    for b in range(batch):
        h = int(torch.randint(lo, hi + 1, (1,), generator=g))
        w = int(torch.randint(lo, hi + 1, (1,), generator=g))
        t = int(torch.randint(0, res - h + 1, (1,), generator=g))
        l = int(torch.randint(0, res - w + 1, (1,), generator=g))
        mask[b, 0, t : t + h, l : l + w] = 1.0
    return ground, mask


# ------------------------------------------------------------------------------------------------------------------
# GPU side of the per-image transform: greyscale + Resize + ToTensor for a whole batch (lib/data/dataset.py:6-12, train.py:69-72)
# ------------------------------------------------------------------------------------------------------------------
_PRECISION_BITS = 32 - 8 - 2  # Pillow's 8-bit resample works in 22-bit fixed point


def _triangle_taps(in_size: int, out_size: int):
    """Per-output-index taps of Pillow's bilinear (triangle) resample along one axis (its precompute_coeffs /
    normalize_coeffs_8bpc, evaluated in double precision in the same operation order): first source index, tap count and the
    22-bit fixed-point weights.  When shrinking, the triangle is stretched by the scale factor (Pillow's antialiasing)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    first, count, weights = [], [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        k, total = [], 0.0
        for x in range(lo, hi):
            a = abs((x - center + 0.5) * inv)
            w = 1.0 - a if a < 1.0 else 0.0
            k.append(w)
            total += w
        if total != 0.0:
            k = [w / total for w in k]
        k += [0.0] * (ksize - len(k))
        first.append(lo)
        count.append(hi - lo)
        weights.append([int(-0.5 + w * (1 << _PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << _PRECISION_BITS)) for w in k])
    return ksize, first, count, weights


class GpuLoadTransform:
    """``pil_loader`` (greyscale) + ``transforms.Resize(size)`` + ``transforms.ToTensor()`` of the reference (lib/data/dataset.py:6-12,
    train.py:69-72) for a BATCH of decoded images on the GPU, bit-exact with Pillow + torchvision: ``uint8 [B,H,W,3]`` (RGB) or
    ``[B,H,W]`` (greyscale) -> ``float32 [B,1,out_h,out_w]`` in [0,1].  Like ``Resize(int)`` the smaller edge becomes ``size`` and
    the aspect ratio is kept.  The tap tables of every (height, width) seen are cached on the device."""

    def __init__(self, size: int, device):
        self.size, self.device = int(size), torch.device(device)
        self._taps: Dict[Tuple[int, int], tuple] = {}

    def output_size(self, h: int, w: int) -> Tuple[int, int]:
        short, long = (w, h) if w <= h else (h, w)
        new_long = int(self.size * long / short)
        return (new_long, self.size) if w <= h else (self.size, new_long)

    def _tables(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        if key not in self._taps:
            ksize, first, count, weights = _triangle_taps(in_size, out_size)
            dev = self.device
            self._taps[key] = (ksize, torch.tensor(first, dtype=torch.int32, device=dev), torch.tensor(count, dtype=torch.int32, device=dev),
                               torch.tensor(weights, dtype=torch.int32, device=dev).contiguous())
        return self._taps[key]

    @torch.no_grad()
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        from . import _lib
        from ._lib import call, ptr, stream

        if images.dtype != torch.uint8 or images.dim() not in (3, 4) or (images.dim() == 4 and images.shape[-1] != 3):
            raise _lib.B200Error("GpuLoadTransform expects uint8 [B,H,W,3] (RGB) or [B,H,W] (greyscale) decoded images")
        x = images.to(self.device, non_blocking=True).contiguous()
        B, H, W = x.shape[:3]
        ch = 3 if x.dim() == 4 else 1
        oh, ow = self.output_size(H, W)
        kx, xmin, xcnt, xk = self._tables(W, ow)
        ky, ymin, ycnt, yk = self._tables(H, oh)
        tmp = torch.empty((B, H, ow), dtype=torch.uint8, device=self.device)
        out = torch.empty((B, 1, oh, ow), dtype=torch.float32, device=self.device)
        call("b200_preprocess", ptr(x), B, H, W, ch, oh, ow, ptr(xmin), ptr(xcnt), ptr(xk), kx, ptr(ymin), ptr(ycnt), ptr(yk), ky, ptr(tmp), ptr(out), stream())
        return out

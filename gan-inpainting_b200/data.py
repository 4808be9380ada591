"""Host -> device input pipeline of the train step.

The reference moves each batch with a blocking ``.to(device)`` inside the batch loop (experiment_list/wgan_l1.py:113-114,
``DataLoader(num_workers=0)`` train.py:77), so the GPU idles during every copy.  ``DevicePrefetcher`` keeps the copy of batch
i+1 (pinned host memory -> device, on a side stream) in flight while batch i is being trained on; the tuple contract of the
loader — ``(ground, mask, segment)`` with ``ground`` / ``mask`` fp32 ``[B, 1, R, R]`` (lib/data/dataset.py:14-51) — is kept.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import torch


class DevicePrefetcher:
    """Iterates ``loader`` one batch ahead: ``for ground, mask, extra in DevicePrefetcher(loader, device)`` yields CUDA tensors
    whose host->device copies overlapped the previous iteration's kernels.  Two device buffers per tensor are recycled; a buffer
    is overwritten only after the step that consumed it has been enqueued AND a newer batch was requested (event-guarded)."""

    def __init__(self, loader: Iterable, device, pin: bool = True):
        self.loader, self.device, self.pin = loader, torch.device(device), pin
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [None, None]
        self.consumed = [None, None]

    def _to_device(self, batch, slot: int):
        ground, mask = batch[0], batch[1]
        rest = tuple(batch[2:])
        if self.pin and not ground.is_pinned():
            ground, mask = ground.pin_memory(), mask.pin_memory()
        cur = self.slots[slot]
        if cur is None or cur[0].shape != ground.shape:
            cur = (torch.empty(ground.shape, dtype=torch.float32, device=self.device), torch.empty(mask.shape, dtype=torch.float32, device=self.device))
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

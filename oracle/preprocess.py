"""Integer restatement of the reference's input pipeline.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference loads every ground-truth / mask image with PIL, converts it to greyscale and resizes + scales it with torchvision:

    lib/data/dataset.py:6-12   Image.open(f).convert('L')
    train.py:69-72             transforms.Compose([transforms.Resize(image_target_size), transforms.ToTensor()])

All of it is third-party arithmetic that is absent from /root/reference: Pillow (unpinned by the reference; 12.2.0 here) and
torchvision (0.26.0 here).  Their published algorithms, restated:

* ``convert('L')`` of an RGB image (Pillow src/libImaging/Convert.c, rgb2l / L24): L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16.
* ``Resize(size)`` of a PIL image = ``Image.resize(..., BILINEAR)`` with the smaller edge -> ``size`` (aspect ratio kept, the long
  edge is ``int(size * long / short)``); Pillow's resample (src/libImaging/Resample.c) is separable, horizontal pass first, each
  pass a convolution with triangle weights stretched by the scale when shrinking ("antialias"), evaluated in 22-bit fixed point:
  weights w = round(k * 2^22) from double-precision k, accumulator starts at 2^21, result clip8(acc >> 22), and the intermediate
  image is 8-bit.
* ``ToTensor()``: uint8 -> float32, divided by 255 (IEEE single division), layout [1, H, W].

``oracle/pin_preprocess_against_reference.py`` checks this restatement bit for bit against Pillow + torchvision run in the build
container and writes tests/golden/reference_preprocess.pt.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def rgb_to_l(rgb: np.ndarray) -> np.ndarray:
    """uint8 [..., 3] -> uint8 [...]."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def resize_output_size(h: int, w: int, size: int):
    """torchvision.transforms.Resize(int): smaller edge -> size, the other edge int(size * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)  # (out_h, out_w)


def bilinear_coeffs(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the triangle filter over the full axis.
    Returns (ksize, xmin [out], xmax [out] (tap counts), kk [out, ksize] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    xcnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        k = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(n):
            a = (x + lo - center + 0.5) * ss
            a = -a if a < 0 else a
            w = 1.0 - a if a < 1.0 else 0.0
            k[x] = w
            ww += w
        if ww != 0.0:
            for x in range(n):
                k[x] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + k[x] * (1 << PRECISION_BITS)) if k[x] < 0 else int(0.5 + k[x] * (1 << PRECISION_BITS))
        xmin[xx], xcnt[xx] = lo, n
    return ksize, xmin, xcnt, kk


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8-bit pass along ``axis`` of a uint8 [H, W] image."""
    in_size = img.shape[axis]
    if in_size == out_size:
        return img
    ksize, xmin, xcnt, kk = bilinear_coeffs(in_size, out_size)
    src = img.astype(np.int64)
    shape = list(img.shape)
    shape[axis] = out_size
    out = np.empty(shape, np.uint8)
    for xx in range(out_size):
        sl = np.take(src, np.arange(xmin[xx], xmin[xx] + xcnt[xx]), axis=axis)
        w = kk[xx, : xcnt[xx]].astype(np.int64)
        acc = (1 << (PRECISION_BITS - 1)) + (np.tensordot(sl, w, axes=([axis], [0])))
        v = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        if axis == 1:
            out[:, xx] = v
        else:
            out[xx, :] = v
    return out


def resize_l(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Image.resize((out_w, out_h), BILINEAR) of an 8-bit greyscale image: horizontal pass, 8-bit intermediate, vertical pass."""
    return _resample_axis(_resample_axis(img, out_w, 1), out_h, 0)


def load_transform(img: np.ndarray, size: int) -> np.ndarray:
    """dataset.pil_loader + Resize(size) + ToTensor of one decoded image: uint8 [H, W, 3] or [H, W] -> float32 [1, out_h, out_w]."""
    l = rgb_to_l(img) if img.ndim == 3 else img
    oh, ow = resize_output_size(l.shape[0], l.shape[1], size)
    r = resize_l(l, oh, ow)
    return (r.astype(np.float32) / np.float32(255.0))[None]

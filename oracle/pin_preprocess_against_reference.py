"""Pin oracle/preprocess.py against the third-party code the reference's input pipeline runs (Pillow's convert('L') / resize,
torchvision's Resize + ToTensor; lib/data/dataset.py:6-12, train.py:69-72) and write tests/golden/reference_preprocess.pt.

    python oracle/pin_preprocess_against_reference.py

Runs in the build container (PIL + torchvision present).  The fixture holds small decoded images and the exact tensors the
reference's transform produces for them.
"""
import os
import sys

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import preprocess  # noqa: E402


def reference(img: np.ndarray, size: int) -> torch.Tensor:
    pil = Image.fromarray(img, "RGB" if img.ndim == 3 else "L").convert("L")  # dataset.pil_loader
    return transforms.Compose([transforms.Resize(size), transforms.ToTensor()])(pil)  # train.py:69-72


def main():
    rng = np.random.default_rng(7)
    cases = []
    # (H, W, channels, target): down-scaling (antialiased), up-scaling, identity, non-square, odd sizes, a binary mask
    for h, w, ch, size in [(256, 256, 3, 128), (300, 260, 3, 128), (96, 96, 1, 128), (128, 128, 3, 128), (513, 377, 3, 256), (64, 200, 1, 32),
                           (1024, 1024, 3, 256), (130, 130, 1, 128)]:
        img = rng.integers(0, 256, (h, w, ch) if ch == 3 else (h, w), dtype=np.uint8)
        if (h, w, ch) == (130, 130, 1):  # a mask: rectangle of 255 on 0 (resizing gives the soft edge that ceil() later binarises)
            img = np.zeros((h, w), np.uint8)
            img[40:90, 30:100] = 255
        ref = reference(img, size)
        got = torch.from_numpy(preprocess.load_transform(img, size))
        assert got.shape == ref.shape, (got.shape, ref.shape)
        assert torch.equal(got, ref), ("oracle/preprocess.py differs from Pillow + torchvision", h, w, ch, size, float((got - ref).abs().max()))
        cases.append(dict(image=torch.from_numpy(img), size=size, out=ref))
    out = os.path.join(ROOT, "tests", "golden", "reference_preprocess.pt")
    small = [c for c in cases if c["image"].numel() <= 300 * 260 * 3]  # keep the fixture small
    torch.save(dict(pillow=Image.__version__ if hasattr(Image, "__version__") else "", cases=small), out)
    print(f"pinned {len(cases)} cases bit for bit; wrote {len(small)} to {out} ({os.path.getsize(out) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()

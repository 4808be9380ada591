import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree C-ABI library (built on demand; nvcc cross-compiles without a GPU)."""
    from gan_inpainting_b200 import build

    return build.build(verbose=False)


@pytest.fixture(scope="session")
def golden():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "reference_r128_b2.pt"))


@pytest.fixture()
def dev(built_lib):
    import torch

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")

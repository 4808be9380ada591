"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/b200_inpaint.h declares
(no compute calls — there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200_inpaint.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)


def test_header_matches_binding_table():
    from gan_inpainting_b200 import _lib

    names = declared()
    assert len(names) == len(set(names))
    assert set(names) == set(_lib.exported_symbols())


def test_header_arity_matches_ctypes_signatures():
    from gan_inpainting_b200 import _lib

    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        n = 0 if m.group(1).strip() == "void" else m.group(1).count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_library_exports_every_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in declared():
        assert hasattr(lib, name), name
    from gan_inpainting_b200 import _lib

    assert _lib.load().b200_abi_version() == 1
    assert _lib.value("b200_bn_stat_blocks", 1 << 30, 64) == 148 * 4
    assert _lib.value("b200_wgrad_c1_blocks") == 148 * 3


def test_library_is_sm100a_and_uses_tcgen05_tma(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass or "UTCMMA" in sass or re.search(r"UTC\w*MMA", sass)  # tcgen05.mma
    assert "UTMALDG" in sass  # TMA tensor loads
    assert "LDTM" in sass     # tcgen05.ld
    # warp-level mma.sync (HMMA) is used ONLY by the three HBM-bound layers with one channel on one side (K = 16 or N = 16:
    # conv_c1 / convt_c1 / wgrad_c1); every GEMM-shaped convolution goes through tcgen05
    funcs = re.split(r"\n\s*Function : ", sass)
    legacy = re.compile(r"(?<![A-Z])HMMA\.")  # warp-level HMMA, not the tcgen05 mnemonic UTCHMMA[.2CTA]
    hmma = [f.split("\n", 1)[0] for f in funcs[1:] if legacy.search(f)]
    assert hmma and all("_c1_" in name for name in hmma), hmma
    igemm = [f for f in funcs[1:] if "igemm_kernel" in f.split("\n", 1)[0]]
    assert igemm and all(re.search(r"UTC\w*MMA", f) and not legacy.search(f) for f in igemm)
    assert any("UTCHMMA.2CTA" in f for f in igemm)  # the CTA-pair variants (tcgen05.mma.cta_group::2)


def test_missing_library_fails_loudly(monkeypatch):
    from gan_inpainting_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb200inpaint.so")
    with pytest.raises(_lib.B200Error, match="no CPU / PyTorch fallback"):
        _lib.call("b200_clamp", None, 0, 0.0, 0.0, None)

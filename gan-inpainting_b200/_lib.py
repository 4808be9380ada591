"""ctypes binding of the C-ABI declared in include/b200_inpaint.h.

There is deliberately NO fallback: if libb200inpaint.so is missing (or its symbols are) every
entry point raises, so a GPU run can never silently go through PyTorch / the CPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200inpaint.so")

_lib = None

P = c_void_p  # device (or host, where stated) pointer
S = c_void_p  # cudaStream_t
I = c_int
F = c_float

# name -> argtypes (all functions return int status; 0 == OK)
SIGNATURES = {
    "b200_conv_s2_gemm": [P, I, I, I, I, P, I, P, I, I, I, S],
    "b200_convt_s2_gemm": [P, I, I, I, I, I, P, I, P, I, I, I, S],
    "b200_wgrad_gemm": [P, I, I, I, I, I, P, I, P, POINTER(c_int), I, I, S],
    "b200_debug_tma_probe": [P, I, I, I, I, I, I, I, I, I, I, I, I, I, P, S],
}


class B200Error(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} not found: build it with `python gan-inpainting_b200/build.py` "
            "(there is no CPU / PyTorch fallback for the hot path)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.b200_last_error.restype = c_char_p
    lib.b200_last_error.argtypes = []
    lib.b200_abi_version.restype = c_int
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError == ABI mismatch: fail loudly
        fn.restype = c_int
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    """Every entry point include/b200_inpaint.h declares (used by the CPU-side ABI test)."""
    return ["b200_last_error", "b200_abi_version"] + list(SIGNATURES)


def call(name: str, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise B200Error(f"{name} failed (status {rc}): {lib.b200_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)

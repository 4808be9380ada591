"""ctypes binding of the C-ABI declared in include/b200_inpaint.h.

There is deliberately NO fallback: if libb200inpaint.so is missing (or its symbols are) every
entry point raises, so a GPU run can never silently go through PyTorch / the CPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200inpaint.so")

_lib = None
# bumped by every kernel that rewrites parameters through raw pointers (fused optimizers, clamp)
WEIGHT_EPOCH = 0

P = c_void_p  # device (or host, where stated) pointer
S = c_void_p  # cudaStream_t
I = c_int
L = c_longlong
F = c_float
U = c_uint32

# name -> argtypes (all of these return an int status; 0 == OK).  Order == include/b200_inpaint.h.
SIGNATURES = {
    # ---- tcgen05 implicit-GEMM convolutions (csrc/conv_gemm.cu)
    "b200_conv_s2_gemm": [P, I, I, I, I, P, I, P, I, I, I, P, POINTER(c_int), P, I, P, S],
    "b200_convt_s2_gemm": [P, I, I, I, I, I, P, I, P, I, I, I, P, POINTER(c_int), P, I, P, S],
    "b200_splitk_reduce": [P, I, L, I, P, I, S],
    "b200_wgrad_gemm": [P, I, I, I, I, I, P, I, P, POINTER(c_int), I, I, S],
    # ---- weight layouts (csrc/pack.cu)
    "b200_pack_gather": [P, I, I, P, S],
    "b200_pack_phase": [P, I, I, P, S],
    "b200_pack_all": [P, I, I, S],
    "b200_wgrad_reduce": [P, I, I, I, P, I, S],
    # ---- BatchNorm + activation (csrc/norm_act.cu)
    "b200_bn_stats": [P, L, I, P, P, P, P, P, P, F, F, P, S],
    "b200_bn_finalize": [P, I, L, I, P, P, P, P, P, F, F, P, S],
    "b200_bn_eval_coeffs": [I, P, P, P, P, F, P, S],
    "b200_bn_act": [P, L, I, P, P, I, P, I, I, P, I, I, I, U, S],
    "b200_bn_bwd_fused": [P, L, I, P, P, P, I, I, I, P, I, I, I, P, P, P, I, P, S],
    "b200_bn_act_bwd": [P, L, I, P, P, I, P, I, I, P, I, I, I, U, P, P, P, P, I, P, S],
    # ---- single-channel-side convolutions and the discriminator head (csrc/conv_direct.cu)
    "b200_conv_c1_fwd": [P, P, I, I, I, P, I, I, P, P, I, I, S],
    "b200_convt_c1_fwd": [P, I, I, I, I, I, P, P, I, P, P, P, P, S],
    "b200_wgrad_c1": [P, I, I, I, I, I, P, P, P, P, I, S],
    "b200_head_weight": [P, P, I, I, P, S],
    "b200_head_fwd": [P, P, I, I, P, I, P, P, P, S],
    "b200_head_dlin": [P, P, I, I, P, S],
    "b200_head_bwd_input": [P, P, I, I, P, S],
    "b200_head_bwd_params": [P, P, I, I, I, P, P, P, P, P, P, I, S],
    # ---- masks, compositing, losses (csrc/losses.cu)
    "b200_mask_prepare": [P, P, L, I, P, P, S],
    "b200_composite": [P, P, P, L, P, S],
    "b200_mul": [P, P, L, P, S],
    "b200_recon_loss": [I, P, P, P, L, F, P, P, S],
    "b200_gstep_seed": [P, P, P, P, P, P, I, F, P, I, F, P, P, I, L, P, P, P, P, I, S],
    "b200_adv_loss": [P, I, I, F, F, P, P, S],
    # ---- optimizers (csrc/optim.cu)
    "b200_rmsprop_step": [P, P, P, L, F, F, F, I, F, F, S],
    "b200_step_inc": [P, S],
    "b200_adam_step": [P, P, P, P, L, F, F, F, F, P, S],
    "b200_clamp": [P, L, F, F, S],
    "b200_cast_grads": [P, P, L, I, S],
    "b200_absmean": [P, P, I, P, S],
    # ---- SSIM metric (csrc/ssim.cu)
    "b200_ssim": [P, P, I, I, I, I, POINTER(c_float), I, P, P, S],
    # ---- VGG19 feature losses (csrc/conv_gemm.cu conv3 / gram modes, csrc/vgg.cu)
    "b200_conv3_gemm": [P, I, I, I, I, I, P, P, I, P, I, I, I, S],
    "b200_pack_conv3": [P, I, I, P, S],
    "b200_conv3_c1_fwd": [P, I, I, I, P, P, I, P, S],
    "b200_maxpool2": [P, I, I, I, I, P, S],
    "b200_gram_gemm": [P, I, I, I, I, I, P, S],
    "b200_sqdiff_bf16": [P, P, L, ctypes.c_double, P, P, I, S],
    "b200_sqdiff_gram": [P, P, I, I, I, I, F, ctypes.c_double, P, P, I, S],
    # ---- input pipeline: greyscale + resize + ToTensor (csrc/preprocess.cu)
    "b200_preprocess": [P, I, I, I, I, I, I, P, P, P, I, P, P, P, I, P, P, S],
    # ---- total-variation loss (csrc/tv.cu)
    "b200_tv_loss": [P, L, I, I, F, P, P, S],
    "b200_tv_loss_bwd": [P, L, I, I, F, P, P, I, S],
}
# helpers that return a plain value instead of a status
VALUE_FUNCS = {
    "b200_abi_version": (c_int, []),
    "b200_last_error": (c_char_p, []),
    "b200_launch_count": (c_longlong, []),
    "b200_launch_count_add": (None, [c_longlong]),
    "b200_set_seed_counter": (None, [c_void_p]),
    "b200_set_sm_limit": (None, [c_int]),
    "b200_bn_stat_blocks": (c_int, [L, I]),
    "b200_igemm_stat_rows": (c_int, []),
    "b200_wgrad_c1_blocks": (c_int, []),
    "b200_recon_scratch_floats": (c_int, []),
    "b200_ssim_partials": (c_longlong, [I, I, I, I]),
    "b200_sqdiff_blocks": (c_int, []),
    "b200_gram_splits": (c_int, [I, I, I, I]),
}


class EpiBwd(ctypes.Structure):
    """ctypes mirror of ``b200_epi_bwd`` (include/b200_inpaint.h)."""

    _fields_ = [("y", c_void_p), ("ldy", c_int), ("scale", c_void_p), ("shift", c_void_p), ("act", c_int), ("g2", c_void_p), ("ldg2", c_int),
                ("offg2", c_int), ("dropout", c_int), ("seed", c_uint32), ("col0", c_int), ("Cy", c_int), ("dz32", c_void_p)]


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
    for name, (res, args) in VALUE_FUNCS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError == ABI mismatch: fail loudly
        fn.restype = c_int
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    """Every entry point include/b200_inpaint.h declares (used by the CPU-side ABI test)."""
    return list(VALUE_FUNCS) + list(SIGNATURES)


# optional per-call hook: PROFILE(name, args) -> callable run after the call (bench.py times kernels with CUDA events this way and derives
# the algorithmic bytes of the memory-bound calls from their arguments)
PROFILE = None


def call(name: str, *args):
    lib = load()
    after = PROFILE(name, args) if PROFILE is not None else None
    rc = getattr(lib, name)(*args)
    if after is not None:
        after()
    if rc != 0:
        raise B200Error(f"{name} failed (status {rc}): {lib.b200_last_error().decode()}")


def value(name: str, *args):
    return getattr(load(), name)(*args)


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)

"""B200-native GAN-inpainting train step (drop-in for the reference's lib/models + experiment_list path).

Layout:
  csrc/             hand-written sm_100a CUDA kernels + the C-ABI (libb200inpaint.so)
  _lib.py           ctypes binding of the C-ABI (include/b200_inpaint.h)
  lib/models/       networks / loss / util with the reference's module interfaces
  experiment_list/  begin(state, loaders) recipes with the reference's call surface
"""
__version__ = "0.1.0"

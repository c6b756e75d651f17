"""gr_doa_b200 -- B200-native (sm_100a) implementation of gr-doa's direction-of-arrival hot path.

The product is libdoa_cuda.so (hand-written CUDA behind the C ABI of include/doa_cuda.h).  This package is the
thin host side: a ctypes binding and Python mirrors of the four reference blocks
(autocorrelate, MUSIC_lin_array, rootMUSIC_linear_array, find_local_max) plus the fused DoaChain.
Importing the blocks loads the shared library and raises if it is missing -- there is no CPU or PyTorch fallback.
"""
from ._lib import dev_library  # noqa: F401
from .blocks import set_default_option  # noqa: F401
from .blocks import DoaChain, DoaChainMulti, MUSIC_lin_array, RootMusicChain, antenna_correction, autocorrelate, calibrate_lin_array, find_local_max, rootMUSIC_linear_array  # noqa: F401

__all__ = ["autocorrelate", "MUSIC_lin_array", "rootMUSIC_linear_array", "find_local_max", "DoaChain", "DoaChainMulti", "RootMusicChain", "antenna_correction", "calibrate_lin_array"]

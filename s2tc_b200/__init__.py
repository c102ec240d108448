"""s2tc_b200 -- B200 (sm_100a) S2TC texture encoder; Python host-side mirror of the C ABI.

The product is the shared object built from ``s2tc_b200/csrc`` (``lib/libs2tc_b200.so`` and its twin
``lib/libtxc_dxtn.so``): hand-written CUDA kernels behind the reference's own interface
(``tx_compress_dxtn``, ``s2tc_encode_block_func``, ``rgb565_image``, the ``S2TC_*`` environment) plus a
batched / device-pointer C ABI (``include/s2tc_b200.h``).  This package is a thin ctypes binding over
that ABI for tests, the benchmark and Python users; PyTorch is only used by callers for device
memory and process groups.

There is no CPU fallback anywhere: importing works without a GPU (so that the library's exports can be
checked), every call that would encode raises ``S2TCError`` when no CUDA device is usable, and a
missing shared object is an ``ImportError`` with the build command.
"""
from .api import (  # noqa: F401
    AVG, DITHER_FLOYDSTEINBERG, DITHER_NONE, DITHER_SIMPLE, DXT1, DXT3, DXT5, NORMALMAP, REFINE_ALWAYS,
    REFINE_LOOP, REFINE_NEVER, RGB, SRGB, SRGB_MIXED, W0AVG, WAVG, YUV, Encoder, S2TCError, Settings,
    block_bytes, build, draws_per_block, lib, lib_path, settings_from_env, tx_compress_dxtn,
)

__all__ = [
    "Encoder", "Settings", "S2TCError", "build", "lib", "lib_path", "tx_compress_dxtn", "settings_from_env",
    "block_bytes", "draws_per_block",
]

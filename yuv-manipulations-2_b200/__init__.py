"""yuv-manipulations-2_b200 -- B200-native (sm_100a) implementation of myyuv_lib's hot path.

The product is the CUDA library lib/libmyyuvb200.so behind the C ABI of include/myyuvb200.h, plus the
drop-in C++ class library lib/libmyyuv_lib.so (myyuv::BMP / myyuv::YUV).  This Python package is plumbing:
ctypes bindings of the C ABI (`capi`), a mirror of the reference's class API used by the tests and the
benchmark (`myyuv`), and the multi-GPU sharding logic on top of torch.distributed (`sharding`).

There is no CPU fallback anywhere in this package: without the CUDA library or without a GPU every codec
call raises.  The directory name is not a Python identifier; import it with
    importlib.import_module("yuv-manipulations-2_b200")      or      import myyuv_b200   (alias at repo root)
"""
from . import capi  # noqa: F401
from .capi import Context, MyyuvError, library_path  # noqa: F401
from .myyuv import BMP, YUV, BMPHeader, YUVHeader  # noqa: F401

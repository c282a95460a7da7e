"""cudaraster-linux_b200 -- B200-native CudaRaster hot path.

The product is ``libcrb200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/crb200.h``).  This package is the thin Python host layer used by the tests and the
benchmark: it loads the library with ctypes and mirrors the reference's host API
(``FW::CudaRaster`` / ``FW::CudaSurface``, /root/reference/src/cudaraster/CudaRaster.hpp:42-190)
on top of torch CUDA tensors, which are used only as device memory.

There is NO CPU fallback: constructing ``CudaRaster`` without the library or without a CUDA
device raises.
"""
from .binding import (CudaRaster, CudaSurface, CrbError, build_library, library_path, load_library,
                      RenderModeFlag_EnableDepth, RenderModeFlag_EnableLerp, RenderModeFlag_EnableQuads, pipe_name)
from . import scenes

__all__ = ["CudaRaster", "CudaSurface", "CrbError", "build_library", "library_path", "load_library", "scenes", "pipe_name",
           "RenderModeFlag_EnableDepth", "RenderModeFlag_EnableLerp", "RenderModeFlag_EnableQuads"]

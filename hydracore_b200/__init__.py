"""hydracore_b200 — B200-native CUDA hardware layer for the Hydra renderer (drop-in for the IHWLayer hot path).

The product is ``libhydracore_b200.so`` (hand-written sm_100a kernels behind the C ABI of ``include/hydracore_cuda.h``)
plus the C++ ``GPUCUDALayer : IHWLayer`` mirror under ``csrc/host``.  This Python package only binds the C ABI with
ctypes (for tests and bench.py) and packs synthetic scenes into the reference's blob formats.  There is no CPU fallback:
everything that computes raises if the library or a CUDA device is missing.
"""
from . import layout  # noqa: F401
from ._lib import load, lib_path, HcError, HC_HOST, HC_DEVICE  # noqa: F401
from .layer import CudaLayer, BvhBuilder  # noqa: F401

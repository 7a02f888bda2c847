"""brdf_b200 -- B200-native BRDF-fitting hot path (gather -> model residual/Jacobian -> LM normal
equations + solve) behind the reference's levmar-style C interface.

The product is the C-ABI shared library ``libbrdfgpu.so`` (include/brdfgpu.h, sources in
``brdf_b200/csrc``); this package is a thin ctypes mirror of it for tests and benchmarks.  There is
no CPU implementation here: every compute call needs the CUDA library and a GPU.
"""
from .api import (  # noqa: F401
    BLINN_PHONG,
    PHONG,
    REF_GLOBAL,
    REF_PERFACE,
    BrdfGpuError,
    Context,
    ExtraData,
    build,
    lib,
    lib_path,
)

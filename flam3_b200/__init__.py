"""flam3_b200 — B200-native render hot path of flam3 behind the flam3.h C API.

The product is native: ``flam3_b200/lib/libflam3_b200.so`` (CUDA kernels for sm_100a and the
stage-level C-ABI of ``include/flam3_b200.h``) and ``dropin/_build/libflam3.so`` (the
``flam3_render`` drop-in). This package only holds thin ctypes bindings used by the tests,
``bench.py`` and the multi-GPU driver; there is no Python or CPU compute path here, and
loading fails loudly when the native libraries are missing.
"""
from .capi import (  # noqa: F401
    CoreLib,
    Flam3Lib,
    RenderArgs,
    core_lib_path,
    dropin_lib_path,
    load_core,
    load_dropin,
    repo_root,
)

__all__ = [
    "CoreLib",
    "Flam3Lib",
    "RenderArgs",
    "core_lib_path",
    "dropin_lib_path",
    "load_core",
    "load_dropin",
    "repo_root",
]

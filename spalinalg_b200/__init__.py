"""spalinalg_b200 — B200 (sm_100a) implementation of spalinalg's sparse hot path.

Public surface mirrors the reference crate root (src/lib.rs:10-19): CooMatrix, CscMatrix,
CsrMatrix, DokMatrix.  Device work goes through libspalinalg_b200.so (include/spl.h); build it
with `python -m spalinalg_b200.build`.  No CPU fallback.
"""
from . import matrix
from .matrix import (Context, CooMatrix, CscMatrix, CsrMatrix, DeviceError, DokMatrix, Panic, PinnedCooMatrix,
                     default_context, pinned_empty, set_default_context)

__all__ = ["Context", "CooMatrix", "CscMatrix", "CsrMatrix", "DokMatrix", "Panic", "DeviceError", "PinnedCooMatrix",
           "default_context", "pinned_empty", "set_default_context"]

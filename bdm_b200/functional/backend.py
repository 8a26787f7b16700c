"""Same role as the reference's functional/backend.py (:12-33): exposes `_backend`.  The reference
JIT-compiles its CUDA sources here; we bind the prebuilt C-ABI library instead (no fallback)."""
from .. import backend as _backend

__all__ = ['_backend']

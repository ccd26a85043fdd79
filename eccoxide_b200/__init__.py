"""eccoxide_b200 — B200-native batched elliptic-curve scalar multiplication.

Batch siblings of vincenthz/eccoxide's per-element API for its one data-parallel hot path
(Point::mul_base, Point::mul, x25519/x448, ecdsa/ed25519 verify), executed by hand-written CUDA
for sm_100a behind the C ABI in include/eccbatch.h.  No CPU fallback.
"""
from ._lib import EccBatchError, load  # noqa: F401
from .context import Context  # noqa: F401

__all__ = ["Context", "EccBatchError", "load"]

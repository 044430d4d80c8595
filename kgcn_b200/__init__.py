"""kgcn_b200 -- B200-native batched graph-convolution path behind kGCN's layer / plugin-op API.

Importing the package loads ``libkgcn_b200.so`` (sm_100a CUDA, C ABI in ``include/kgcn_b200.h``)
and raises ImportError if it has not been built: there is deliberately no CPU fallback.
"""
from . import _lib  # noqa: F401  (fail loudly when the CUDA library is missing)
from ._lib import KgcnError, KgcnIndexError  # noqa: F401
from .csr import BatchedCSR  # noqa: F401

__all__ = ["BatchedCSR", "KgcnError", "KgcnIndexError"]

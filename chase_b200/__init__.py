"""chase_b200 — B200-native (sm_100a) Chebyshev-filtered subspace iteration.

Python is only the thin test/bench harness around the native library:
``chase_b200.solver`` mirrors the reference's C interface (``?chase_init_`` /
``?chase_`` / ``?chase_finalize_``) on numpy host buffers, ``chase_b200.kernels``
exposes the kernel-level C ABI on torch CUDA tensors.
"""
from ._lib import LIB_PATH, build, lib  # noqa: F401
from .solver import ChASE, SolveResult  # noqa: F401

__all__ = ["ChASE", "SolveResult", "build", "lib", "LIB_PATH"]

"""motes_b200 - B200-native implementation of the MOTES per-column extraction hot path.

Sub-modules mirror the reference's function API for the path
(``/root/reference/src/motes/{common,tracing,sky,extraction}.py``); the arithmetic runs
in hand-written sm_100a CUDA kernels behind the C ABI of ``include/motes_b200.h``.
Importing the package loads ``libmotes_b200.so`` and fails if it has not been built;
there is no CPU fallback.
"""

from . import _lib  # noqa: F401  (fails loudly when the CUDA extension is missing)

__all__ = ["_lib"]

"""motes_b200 - B200-native implementation of the MOTES per-column extraction hot path.

Sub-modules mirror the reference's function API for the path
(``/root/reference/src/motes/{common,tracing,sky,extraction}.py``); the arithmetic runs
in hand-written sm_100a CUDA kernels behind the C ABI of ``include/motes_b200.h``.
Every compute module (``tracing``, ``sky``, ``extraction``, ``pipeline``, ``core``) imports ``motes_b200._lib``,
which loads ``libmotes_b200.so`` and raises ``ImportError`` if it has not been built; there is no CPU
fallback.  ``synth`` (test-frame generator) and ``io`` (FITS) are plain numpy and do not load the library.
"""

__all__ = ["common", "tracing", "sky", "extraction", "pipeline", "core", "distributed", "synth", "io"]

"""Shared parity metrics for the tests (SURVEY.md §8c tolerances)."""
import numpy as np


def param_errors(got, ref, nspat):
    """Per-parameter error of fitted Moffat parameters (rows = fits, 6 columns).

    A, c, alpha, beta: relative.  B and m sit near zero after sky subtraction, where a relative
    error is ill-posed (the reference's own 1-ulp self-noise is 3e-5 on B): they are measured
    against max(|x|, A) and max(|x|, A/nspat) respectively.
    """
    got, ref = np.asarray(got)[:, :6], np.asarray(ref)[:, :6]
    amp = np.abs(ref[:, 0])
    scale = np.abs(ref).copy()
    scale[:, 4] = np.maximum(scale[:, 4], amp)
    scale[:, 5] = np.maximum(scale[:, 5], amp / nspat)
    return np.abs(got - ref) / np.maximum(scale, 1e-300)


def summarize(err):
    return dict(median=np.median(err, axis=0), p99=np.percentile(err, 99, axis=0), max=err.max(axis=0))


def rel_err(got, ref, floor=0.0):
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    return np.abs(got - ref) / np.maximum(np.abs(ref), floor if floor > 0 else 1e-300)

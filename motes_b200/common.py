"""Mirror of ``motes.common`` for the hot path (/root/reference/src/motes/common.py:65-114).

``moffat`` and ``moffat_fwhm`` are cheap host-side conveniences used for plotting and
logging by callers of the reference API; the fits and the extraction evaluate the model
on the device (motes_b200/csrc/common.cuh).
"""

import numpy as np


def moffat(parameters, data_range):
    """Moffat profile plus linear background (common.py:65-94)."""
    amplitude, center, alpha, beta, bg_level, bg_grad = parameters[:6]
    shifted = data_range - center
    core = (1 + ((shifted * shifted) / (alpha * alpha))) ** -beta
    return amplitude * core + bg_level + (data_range * bg_grad)


def moffat_fwhm(alpha, beta):
    """Full width at half maximum of a Moffat function (common.py:97-114)."""
    return 2 * alpha * np.sqrt((2 ** (1 / beta)) - 1)

"""Register the FITS shim under the astropy module names when astropy itself is not installed.

The reference imports ``astropy.io.fits`` and ``astropy.table.Table`` at module level
(/root/reference/src/motes/io/motesio.py:9, 19) and matplotlib in its plotting helpers
(diagnostics.py, sky.py); neither package is in this image.  ``install()`` makes ``import motes.core`` work
unmodified: the FITS calls land in ``motes_b200.io.fits``, the plotting modules are empty placeholders (the
reference only touches them when a ``diag_plot_*`` flag is set).  With the real packages installed nothing is
replaced.
"""

from __future__ import annotations

import importlib
import sys
import types


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        importlib.import_module(name)
        return False
    except ImportError:
        return True


def install(plotting: bool = True) -> dict:
    """Returns {package: "real" | "standin"} for astropy and matplotlib."""
    from . import fits
    used = {}
    if _missing("astropy"):
        astropy = types.ModuleType("astropy")
        astropy_io = types.ModuleType("astropy.io")
        table = types.ModuleType("astropy.table")
        table.Table = fits.Table
        astropy.io, astropy.table, astropy_io.fits = astropy_io, table, fits
        astropy.__path__, astropy_io.__path__ = [], []
        sys.modules.update({"astropy": astropy, "astropy.io": astropy_io, "astropy.io.fits": fits, "astropy.table": table})
        used["astropy"] = "standin"
    else:
        used["astropy"] = "standin" if getattr(sys.modules.get("astropy.io.fits"), "__name__", "") == fits.__name__ else "real"
    if plotting:
        if _missing("matplotlib"):
            mpl = types.ModuleType("matplotlib")
            mpl.__path__ = []
            subs = {}
            for sub in ("gridspec", "pyplot", "widgets"):
                m = types.ModuleType(f"matplotlib.{sub}")
                setattr(mpl, sub, m)
                subs[f"matplotlib.{sub}"] = m
            subs["matplotlib.widgets"].Slider = type("Slider", (), {})
            sys.modules.update({"matplotlib": mpl, **subs})
            used["matplotlib"] = "standin"
        else:
            used["matplotlib"] = "real" if hasattr(sys.modules["matplotlib"], "__version__") else "standin"
    return used

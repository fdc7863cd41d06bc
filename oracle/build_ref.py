#!/usr/bin/env python3
"""Stage the UNMODIFIED reference into ``oracle/_ref`` as compiled code objects.  TEST INFRASTRUCTURE ONLY.

The reference is pure Python, so "building" it means compiling the modules of the hot path from the sources where
they lie under ``/root/reference/src/motes`` into ``oracle/_ref/motes/<module>.code`` (the marshalled code object of
each module - what a ``.pyc`` holds after its header; ``oracle/ref_runner.py`` executes them into module objects.
``.pyc`` itself is not used because the gpurun snapshot leaves ``*.pyc`` behind).  No reference source text enters
the repository; the output directory is git-ignored and travels to the GPU box with the snapshot, like the repo's own
built ``.so`` files.

    python oracle/build_ref.py            # run in the build container; a no-op where /root/reference is absent

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU legs load the result (through ``oracle/ref_runner.py``),
as the checker and as the timed CPU baseline - never the product.
"""

from __future__ import annotations

import json
import marshal
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = "/root/reference/src/motes"
OUT = os.path.join(HERE, "_ref", "motes")
# the modules core.motes drives on the path (core.py:80-275), plus what they import at module level
MODULES = ("__init__", "common", "tracing", "sky", "extraction", "diagnostics", "logs")


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds the staged reference afterwards."""
    stamp = os.path.join(OUT, "STAGED.json")
    if not os.path.isdir(REFERENCE_SRC):
        if verbose:
            print(f"oracle/build_ref: {REFERENCE_SRC} is absent; keeping whatever oracle/_ref already holds")
        return os.path.exists(stamp)
    os.makedirs(OUT, exist_ok=True)
    staged = {}
    for name in MODULES:
        src = os.path.join(REFERENCE_SRC, name + ".py")
        dst = os.path.join(OUT, name + ".code")
        with open(src, "rb") as fh:
            code = compile(fh.read(), f"motes/{name}.py", "exec", dont_inherit=True)
        with open(dst, "wb") as fh:
            marshal.dump(code, fh)
        staged[name] = os.path.getsize(dst)
    for stale in os.listdir(OUT):
        if stale.endswith(".pyc"):
            os.remove(os.path.join(OUT, stale))
    with open(stamp, "w") as fh:
        json.dump({"from": REFERENCE_SRC, "python": sys.version.split()[0], "modules": staged}, fh, indent=1)
    if verbose:
        print(f"oracle/build_ref: compiled {len(staged)} reference modules into {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)

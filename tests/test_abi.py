"""The C-ABI library loads and exports every symbol include/motes_b200.h declares (no compute without a GPU)."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "motes_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(motes_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from motes_b200 import _lib
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(_lib.lib, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_compute_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from motes_b200 import _lib
    with pytest.raises(_lib.MotesCudaError):
        _lib.Handle(0)

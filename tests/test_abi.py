"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/stylesinger_b200.h declares (no compute calls: there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(REPO, "include", "stylesinger_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from stylesinger_b200 import build
    lib_path = build.build(verbose=False)
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    lib.ssb_version.restype = ctypes.c_int
    assert lib.ssb_version() >= 100


def test_binding_lists_every_header_symbol():
    from stylesinger_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stylesinger_b200 import synth
    from stylesinger_b200.engine import AcousticModel
    from stylesinger_b200._lib import SsbError
    with pytest.raises(SsbError):
        AcousticModel({}, None)


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "stylesinger_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f

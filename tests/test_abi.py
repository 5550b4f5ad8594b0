"""CPU: the C-ABI library loads, exports every symbol include/qtrack.h declares, and refuses to
run without a GPU instead of falling back."""
import ctypes
import pathlib
import re

import pytest

from lqr_quadcopter_test_b200 import _lib
from lqr_quadcopter_test_b200.params import QtParams, default_params

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_symbols() -> list[str]:
    text = (ROOT / "include" / "qtrack.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qt_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree_on_the_symbol_list():
    assert declared_symbols() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_params_struct_layout_and_defaults():
    lib = _lib.load_library()
    assert lib.qt_abi_version() == 6
    assert lib.qt_params_size() == ctypes.sizeof(QtParams)
    p = QtParams()
    lib.qt_params_default(ctypes.byref(p))
    assert bytes(p) == bytes(default_params())


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on CPU-only machines")
    lib = _lib.load_library()
    handle = ctypes.c_void_p()
    assert lib.qt_create(0, ctypes.byref(handle)) == -3  # QT_ERR_NO_DEVICE
    with pytest.raises(_lib.QtrackError):
        _lib.Context(0)
    from lqr_quadcopter_test_b200 import RiccatiLQRController

    with pytest.raises(_lib.QtrackError):  # never silently replaced by the heuristic fallback
        RiccatiLQRController(config={})


def test_product_package_does_not_import_the_oracle():
    pkg = ROOT / "lqr_quadcopter_test_b200"
    for path in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text and "qt_oracle" not in text, path

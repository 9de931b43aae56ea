"""CPU: the C-ABI library builds for sm_100a, loads without a GPU driver, and exports every
symbol include/oo_b200.h declares.  No compute entry point is called here."""
import ctypes
import os
import re

import pytest

from auto_oo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "oo_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oo_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build_library()
    assert os.path.exists(path)
    lib = _lib.load()
    assert lib.oo_abi_version() == _lib.ABI_VERSION == 2


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(build.build_library())
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/oo_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), "ctypes binding and header disagree"


def test_error_strings_and_workspace_sizes():
    lib = _lib.load()
    assert lib.oo_error_string(0) == b"success"
    for code in (-1, -2, -3, -4, -5):
        assert lib.oo_error_string(code) not in (b"success", b"unknown error")
    assert lib.oo_workspace_bytes(_lib.OO_WS_INT2E, 7, 8, 0, 1) == 8 ** 4 * 8
    assert lib.oo_workspace_bytes(_lib.OO_WS_INT2E, 7, 8, 0, 3) == 3 * 8 ** 4 * 8
    assert lib.oo_workspace_bytes(_lib.OO_WS_INT1E, 7, 8, 0, 2) == 2 * 64 * 8
    assert lib.oo_workspace_bytes(_lib.OO_WS_ROTATION, 7, 8, 0, 1) > 0
    assert lib.oo_workspace_bytes(_lib.OO_WS_HESSIAN, 7, 8, 7, 1) > 0
    assert lib.oo_workspace_bytes(99, 7, 8, 0, 1) == 0


def test_invalid_arguments_are_rejected_before_any_launch():
    lib = _lib.load()
    # null pointers / odd leading dimension -> OO_ERR_INVALID_ARG, no CUDA call needed
    assert lib.oo_int2e_transform_f64(None, 0, None, None, None, None, 0, 7, 8, 1, None, None, 0, None) == -1
    assert lib.oo_dgemm_tn_f64(None, None, None, 4, 4, 4, 4, 4, 4, 1, 0, 0, 0, None) == -1
    with pytest.raises(_lib.OOError):
        _lib.check(-3, "unit")


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from auto_oo_b200.engine import HotPathEngine
    import numpy as np
    with pytest.raises(_lib.OOError):
        HotPathEngine(np.eye(2), np.zeros((2,) * 4), np.eye(2), 0.0, 2, 0, 2, [0])

"""The C-ABI library loads and exports every function that include/*.h declares; without a GPU the
compute entry points fail loudly instead of falling back to anything."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from sparsematrixvectormultiplication_b200 import _native as N

INCLUDE = Path(__file__).resolve().parents[1] / "include"
DECL = re.compile(r"^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", re.M)


def declared_functions():
    names = set()
    for header in sorted(INCLUDE.glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        text = re.sub(r"typedef\s+struct[^;{]*\{.*?\}[^;]*;", "", text, flags=re.S)
        names |= {m.group(1) for m in DECL.finditer(text)}
    return names - {"defined", "while"}


def test_every_declared_symbol_is_exported_and_bound():
    lib = N.lib()
    declared = declared_functions()
    assert len(declared) > 70, sorted(declared)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/ but not exported"
        assert name in N.SIGNATURES, f"{name} has no ctypes signature in _native.py"


def test_struct_layouts_match_the_reference():
    # sizes measured on the reference structs
    assert ctypes.sizeof(N.PreMatrixStruct) == 48 and ctypes.sizeof(N.CSRMatrixStruct) == 48
    assert ctypes.sizeof(N.ELLPACKBlockStruct) == 32 and ctypes.sizeof(N.HLLMatrixStruct) == 16
    assert N.PreMatrixStruct.I.offset == 16 and N.PreMatrixStruct.type.offset == 40
    assert N.ELLPACKBlockStruct.JA.offset == 16 and N.ELLPACKBlockStruct.AS.offset == 24


def test_library_identity():
    assert N.lib().spmv_b200_version() >= 100
    assert N.lib().spmv_b200_vec_ws_doubles() > 0


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sparsematrixvectormultiplication_b200 import host
    from sparsematrixvectormultiplication_b200.device import DeviceCSR, device_count
    assert device_count() == 0
    rp = np.array([0, 1, 2], np.int32)
    ci = np.array([0, 1], np.int32)
    va = np.array([1.0, 2.0])
    with pytest.raises(N.SpmvError):
        DeviceCSR.upload(2, 2, rp, ci, va)
    y = np.zeros(2)
    with pytest.raises(N.SpmvError):
        host.csr_matrix_vector_mult(2, rp, ci, va, np.ones(2), y)
    assert np.isnan(y).all()  # poisoned, never silently computed on the CPU

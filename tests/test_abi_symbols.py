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


def test_row_kernel_forms_are_described_without_a_gpu():
    """spmv_b200_row_forms / _describe (the forms of csr_rowm_kernel / hll_rowm_kernel the fp32 tuner times): pure table
    look-ups, usable on a CPU box; an index outside the table is an error, not a crash."""
    from sparsematrixvectormultiplication_b200 import device
    for fmt in (device.FORMAT_CSR, device.FORMAT_HLL):
        forms = device.row_forms(fmt)
        assert len(forms) >= 4 and len(set(forms)) == len(forms)
        for rows, batch, ctas in forms:
            assert 1 <= rows <= 4 and 3 <= batch <= 7 and 3 <= ctas <= 8
        assert any(rows == 1 for rows, _, _ in forms) and any(rows > 1 for rows, _, _ in forms)
        r = ctypes.c_int()
        assert N.lib().spmv_b200_row_form_describe(fmt, len(forms), ctypes.byref(r), None, None) != 0
        assert N.lib().spmv_b200_row_form_describe(fmt, -1, None, None, None) != 0
        assert N.lib().spmv_b200_row_form_describe(fmt, 0, None, None, None) == 0   # NULL outputs are allowed
    assert N.lib().spmv_b200_row_forms(7) == 0
    assert device.row_form_name(device.FORMAT_CSR, 5) == "csr_row_kernel<5,float>"
    assert device.row_form_name(device.FORMAT_HLL, 16).startswith("hll_rowm_kernel<")
    assert device.row_form_name(device.FORMAT_HLL, 64 + 5) == "hll_rowu_kernel<5,float>"
    assert N.lib().spmv_b200_csr_row_form_f32(None) == 0 and N.lib().spmv_b200_hll_row_form_f32(None) == 0

"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares.
No compute entry point is called (there is no GPU here and the library has no CPU fallback)."""
import ctypes
import os
import re

import pytest

import grid_b200 as gb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gridb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-zA-Z0-9_]+)\s*\(", text)) - {"gb_hermop_fn"})


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(gb.LIB_PATH)
    names = header_symbols()
    assert len(names) > 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gridb200.h but not exported by libgridb200.so"


def test_python_binding_covers_the_header():
    bound = {n for n, _, _ in gb.SYMBOLS}
    assert bound == set(header_symbols())
    gb.lib()  # sets restype/argtypes for each: raises AttributeError on a missing symbol


def test_opcodes_match_oracle_binding():
    from oracle import pyoracle as po
    for n in ("OP_DHOP", "OP_DHOP_OE", "OP_DHOP_EO", "OP_M", "OP_MDAG", "OP_MEOOE", "OP_MEOOE_DAG", "OP_MOOEE", "OP_MOOEE_DAG",
              "OP_MOOEE_INV", "OP_MOOEE_INV_DAG", "OP_MPC", "OP_MPC_DAG", "OP_HERMOP", "OP_DW", "OP_MEOOE5D", "OP_MEOOEDAG5D"):
        assert getattr(gb, n) == getattr(po, n)
    text = open(os.path.join(ROOT, "include", "gridb200.h")).read()
    for name, val in re.findall(r"GB_(OP_[A-Z0-9_]+)\s*=\s*(\d+)", text):
        pyname = name.replace("OP_MEOOE5D", "OP_MEOOE5D")
        assert getattr(gb, pyname) == int(val), name


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product path must fail loudly (GB_ERR_NO_DEVICE), never compute on the host."""
    if gb.lib().gb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(gb.GridB200Error) as e:
        gb.Context(0)
    assert e.value.code == gb.GB_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "grid_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references the oracle"
    assert "oracle" not in open(os.path.join(ROOT, "include", "gridb200.h")).read()

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


def _c_prototypes():
    """{name: (return type, [parameter types])} parsed from include/gridb200.h"""
    text = open(os.path.join(ROOT, "include", "gridb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(gb_[a-zA-Z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if name == "gb_hermop_fn" or "typedef" in ret:
            continue
        params = [] if args in ("", "void") else [a.strip() for a in re.sub(r"\s+", " ", args).split(",")]
        protos[name] = (ret, params)
    return protos


def _category(ctype_decl):
    """coarse ABI class of a C parameter / return declaration"""
    d = ctype_decl
    if "gb_hermop_fn" in d:
        return "fnptr"
    if "*" in d or "[" in d:
        return "ptr"
    if re.search(r"\b(double)\b", d):
        return "f64"
    if re.search(r"\b(int64_t|uint64_t|size_t)\b", d):
        return "i64"
    return "i32"     # int, enums (gb_precision, gb_gridkind), uint32_t


def _py_category(t):
    import ctypes as C
    if t is None:
        return "void"
    if t in (C.c_double,):
        return "f64"
    if t in (C.c_int64, C.c_uint64, C.c_size_t):
        return "i64"
    if t in (C.c_int, C.c_uint32):
        return "i32"
    if isinstance(t, type) and issubclass(t, C._CFuncPtr):
        return "fnptr"
    return "ptr"     # c_void_p, c_char_p, POINTER(...)


def test_ctypes_signatures_match_the_header():
    """Every binding in grid_b200.SYMBOLS has the parameter count and ABI classes (pointer / int / int64 / double / callback) of its
    prototype in include/gridb200.h -- a wrong argtype here would corrupt a call silently on the GPU box."""
    protos = _c_prototypes()
    assert set(protos) == {n for n, _, _ in gb.SYMBOLS}
    for name, res, args in gb.SYMBOLS:
        ret, params = protos[name]
        assert len(params) == len(args), (name, params, args)
        for p, a in zip(params, args):
            assert _category(p) == _py_category(a), (name, p, a)
        want_ret = "ptr" if "*" in ret else _category(ret)
        assert want_ret == _py_category(res), (name, ret, res)


def test_library_carries_the_nvtx_trace_ranges():
    """NVTX ranges named like the reference's GRID_TRACE regions (ref: Grid/perfmon/Tracing.h:5-70; WilsonFermion5DImplementation.h:324-408,
    ConjugateGradient.h:72) are compiled into the product library"""
    blob = open(os.path.join(ROOT, "grid_b200", "libgridb200.so"), "rb").read()
    for name in (b"Dhop", b"DhopDag", b"HaloExchange", b"Gather", b"ConjugateGradient", b"MixedPrecisionConjugateGradient", b"ImportGauge", b"DhopHost"):
        assert name + b"\x00" in blob, name

"""CPU tier: the C-ABI library loads, exports every symbol include/ddf_b200.h declares,
the ctypes table covers the header, and the product fails loudly without a GPU (no
CPU fallback, no routing through the oracle)."""
import ast
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ddf_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert len(syms) >= 60
    for must in ("ddf_mesh_assemble", "ddf_op_deriv", "ddf_op_hodge", "ddf_op_apply", "ddf_op_laplace",
                 "ddf_wave_leapfrog", "ddf_wave_rhs", "ddf_op_compose", "ddf_op_adjoint", "ddf_halo_exchange"):
        assert must in syms


def test_library_exports_every_header_symbol(ddf):
    lib = ddf.load()
    out = subprocess.run(["nm", "-D", "--defined-only", ddf.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (ddf_[a-z0-9_]+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, f"declared in include/ddf_b200.h but not exported: {missing}"
    assert lib.ddf_version() >= 100
    # nothing but the C ABI leaks out of the library
    leaked = [s for s in re.findall(r" T (\S+)", out) if not s.startswith("ddf_") and not s.startswith("_")]
    assert not leaked, leaked


def test_ctypes_table_covers_header(ddf):
    from importlib import import_module
    _lib = import_module("ddf.jl_b200._lib")
    bound = set(_lib.SIGNATURES) | set(_lib.NON_STATUS)
    assert set(header_symbols()) == bound


def test_no_cpu_fallback(ddf):
    """Without an sm_100 device the context cannot be created; nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ddf.DDFError, match="no CPU fallback"):
        ddf.Context(0)


def test_product_never_imports_oracle():
    """The product package must not import, call or link anything under oracle/."""
    pkg = os.path.join(ROOT, "ddf.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    for n in names:
                        assert "oracle" not in n, f"{path} imports {n}"
            elif f.endswith((".cu", ".cuh", ".h", ".sh")):
                for ln in open(path).read().splitlines():
                    if "oracle/" in ln.lower() or "#include" in ln and "oracle" in ln.lower():
                        raise AssertionError(f"{path} references the oracle: {ln.strip()}")


def test_missing_library_fails_loudly(tmp_path):
    code = (
        "import sys, os\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import importlib\n"
        "m = importlib.import_module('ddf.jl_b200._lib')\n"
        "m.LIB_PATH = os.path.join(%r, 'nope.so')\n"
        "try:\n"
        "    m.load()\n"
        "except m.DDFError as e:\n"
        "    print('LOUD', e)\n"
    ) % str(tmp_path)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "LOUD" in out.stdout and "no CPU fallback" in out.stdout

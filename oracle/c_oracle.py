"""ctypes loader of the C twin of the oracle (oracle/ddf_oracle.c).

TEST INFRASTRUCTURE ONLY (see oracle/ddf_oracle.py).  Builds with gcc on first use
into oracle/_build/ (git-ignored).  -ffp-contract=off keeps the multiply and the add
separately rounded like Julia's `mul!`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "ddf_oracle.c")
OUT = os.path.join(_HERE, "_build", "libddf_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", SRC, "-o", OUT])
    return OUT


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def spmv_adjoint_csc_i8(colptr, rowval, nzval, x, omp=False):
    lib = load()
    ncols = len(colptr) - 1
    y = np.empty(ncols)
    fn = lib.ddf_spmv_adjoint_csc_i8_omp if omp else lib.ddf_spmv_adjoint_csc_i8
    fn(C.c_int64(ncols), _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(nzval, C.c_int8), _p(x, C.c_double),
       _p(y, C.c_double))
    return y


def spmv_csc_i8(nrows, colptr, rowval, nzval, x):
    lib = load()
    y = np.empty(nrows)
    lib.ddf_spmv_csc_i8(C.c_int64(nrows), C.c_int64(len(colptr) - 1), _p(colptr, C.c_int64), _p(rowval, C.c_int64),
                        _p(nzval, C.c_int8), _p(x, C.c_double), _p(y, C.c_double))
    return y


def spmv_csc_f64(nrows, colptr, rowval, nzval, x):
    lib = load()
    y = np.empty(nrows)
    lib.ddf_spmv_csc_f64(C.c_int64(nrows), C.c_int64(len(colptr) - 1), _p(colptr, C.c_int64), _p(rowval, C.c_int64),
                         _p(nzval, C.c_double), _p(x, C.c_double), _p(y, C.c_double))
    return y


def diag_mul(d, x):
    lib = load()
    y = np.empty(len(x))
    lib.ddf_diag_mul(C.c_int64(len(x)), _p(d, C.c_double), _p(x, C.c_double), _p(y, C.c_double))
    return y


def wave_leapfrog_csc(colptr, rowval, nzval, mask, u, v, dt, nsteps):
    lib = load()
    n = len(u)
    u = np.ascontiguousarray(u, dtype=np.float64).copy()
    v = np.ascontiguousarray(v, dtype=np.float64).copy()
    tmp = np.empty(n)
    lib.ddf_wave_leapfrog_csc(C.c_int64(n), _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(nzval, C.c_double),
                              _p(mask, C.c_double), _p(u, C.c_double), _p(v, C.c_double), _p(tmp, C.c_double),
                              C.c_double(dt), C.c_int(nsteps))
    return u, v


def calc_faces_boundaries(simplices: np.ndarray):
    """(faces 0-based, colptr, rowval, nzval) with the reference's 1-based CSC storage."""
    lib = load()
    lib.ddf_calc_faces_boundaries.restype = C.c_int64
    s = np.ascontiguousarray(simplices, dtype=np.int64)
    n, N = s.shape
    faces = np.empty((n * N, N - 1), dtype=np.int64)
    colptr = np.empty(n + 1, dtype=np.int64)
    rowval = np.empty(n * N, dtype=np.int64)
    nzval = np.empty(n * N, dtype=np.int8)
    nf = lib.ddf_calc_faces_boundaries(C.c_int64(n), C.c_int(N), _p(s, C.c_int64), _p(faces, C.c_int64),
                                       _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(nzval, C.c_int8))
    return faces[:nf].copy(), colptr, rowval, nzval


def calc_faces_boundaries_linear(simplices: np.ndarray, nvertices: int):
    """Same outputs as calc_faces_boundaries (the reference's 1-based CSC) from the linear-time bucket variant
    (ddf_calc_faces_boundaries_linear): what builds the n = 256 storage of the CPU baseline."""
    lib = load()
    lib.ddf_calc_faces_boundaries_linear.restype = C.c_int64
    s = np.ascontiguousarray(simplices, dtype=np.int64)
    n, N = s.shape
    faces = np.empty((n * N, N - 1), dtype=np.int64)
    colptr = np.empty(n + 1, dtype=np.int64)
    rowval = np.empty(n * N, dtype=np.int64)
    nzval = np.empty(n * N, dtype=np.int8)
    nf = lib.ddf_calc_faces_boundaries_linear(C.c_int64(n), C.c_int(N), C.c_int64(nvertices), _p(s, C.c_int64),
                                              _p(faces, C.c_int64), _p(colptr, C.c_int64), _p(rowval, C.c_int64),
                                              _p(nzval, C.c_int8))
    if nf == -1:
        return calc_faces_boundaries(simplices)
    if nf < 0:
        raise MemoryError("ddf_calc_faces_boundaries_linear: out of memory")
    out = np.empty((nf, N - 1), dtype=np.int64)
    out[:] = faces[:nf]
    del faces
    return out, colptr, rowval, nzval


def kuhn3_table(n: int, block: np.ndarray) -> np.ndarray:
    """Simplex table of large_hypercube_manifold(Val(3); nelts = (n, n, n)) from the 48-tetrahedron mirrored Kuhn
    block (corner offsets in {0, 1, 2}; taken from the oracle's own generator, ddf_oracle.symmetric_hypercube)."""
    lib = load()
    lib.ddf_kuhn3_table.restype = None
    b = np.ascontiguousarray(block, dtype=np.int8)
    assert b.shape == (48, 4, 3) and n % 2 == 0
    out = np.empty((6 * n ** 3, 4), dtype=np.int64)
    lib.ddf_kuhn3_table(C.c_int64(n), _p(b, C.c_int8), _p(out, C.c_int64))
    return out

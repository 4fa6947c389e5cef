"""ctypes binding of libddf_b200.so (the C ABI declared in include/ddf_b200.h).

This stands in for the Julia `ccall` glue of INTEGRATION.md (Julia is not
available in the build image).  There is no CPU fallback: if the shared library
is missing, or no sm_100 device is visible when a context is created, the
package raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libddf_b200.so")


class DDFError(RuntimeError):
    """Non-zero status from libddf_b200 (message from ddf_last_error())."""


_lib = None

_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_i64 = C.c_int64
_i64p = C.POINTER(C.c_int64)
_i8p = C.POINTER(C.c_int8)
_u8p = C.POINTER(C.c_uint8)
_f64 = C.c_double
_f64p = C.POINTER(C.c_double)
_int = C.c_int
_intp = C.POINTER(C.c_int)

# name -> argtypes ; every function returns int status except the two noted
SIGNATURES = {
    "ddf_ctx_create": [_int, _PP],
    "ddf_ctx_destroy": [_P],
    "ddf_sync": [_P],
    "ddf_ctx_bytes_in_use": [_P, _i64p],
    "ddf_ctx_trim": [_P],
    "ddf_timer_tic": [_P],
    "ddf_timer_toc": [_P, _f64p],
    "ddf_ctx_launch_count": [_P, _i64p],
    "ddf_event_record": [_P, _int],
    "ddf_event_elapsed": [_P, _int, _int, _f64p],
    "ddf_flush_l2": [_P],
    "ddf_set_tuning": [C.c_char_p, _int],
    "ddf_ctx_set_tuning": [_P, C.c_char_p, _int],
    "ddf_mesh_create": [_P, _int, _int, _i64, _i64, _i64p, _i64p, _f64p, _f64p, _PP],
    "ddf_mesh_create_hypercube": [_P, _int, _i64p, _i64, _i64, _PP],
    "ddf_mesh_destroy": [_P],
    "ddf_mesh_assemble": [_P],
    "ddf_mesh_nsimplices": [_P, _int, _i64p],
    "ddf_mesh_dims": [_P, _intp, _intp],
    "ddf_mesh_get_simplices_csc": [_P, _int, _i64p, _i64p],
    "ddf_mesh_get_boundary_csc": [_P, _int, _i64p, _i64p, _i8p],
    "ddf_mesh_get_lookup_csc": [_P, _int, _int, _i64p, _i64p, _i64p],
    "ddf_mesh_refine_coords": [_P, _f64p],
    "ddf_mesh_geometry": [_P, _int, _int],
    "ddf_mesh_get_coords": [_P, _f64p],
    "ddf_mesh_get_weights": [_P, _f64p],
    "ddf_mesh_get_volumes": [_P, _int, _f64p],
    "ddf_mesh_get_dualvolumes": [_P, _int, _f64p],
    "ddf_mesh_get_dualcoords": [_P, _int, _f64p],
    "ddf_mesh_get_isboundary": [_P, _int, _u8p],
    "ddf_vec_create": [_P, _i64, _PP],
    "ddf_vec_destroy": [_P],
    "ddf_vec_length": [_P, _i64p],
    "ddf_vec_upload": [_P, _f64p],
    "ddf_vec_download": [_P, _f64p],
    "ddf_vec_upload_async": [_P, _f64p],
    "ddf_vec_download_async": [_P, _f64p],
    "ddf_vec_fill": [_P, _f64],
    "ddf_vec_copy": [_P, _P],
    "ddf_vec_sample": [_P, _int, _P],
    "ddf_vec_rand": [_P, C.c_uint64],
    "ddf_vec_axpby": [_f64, _P, _f64, _P, _P],
    "ddf_vec_mul": [_P, _P, _P],
    "ddf_vec_div": [_P, _f64, _P],
    "ddf_vec_norm": [_P, _int, _f64p],
    "ddf_vec_dot": [_P, _P, _f64p],
    "ddf_vec_sum": [_P, _f64p],
    "ddf_vec_wdot": [_P, _int, _int, _P, _P, _f64p],
    "ddf_op_boundary": [_P, _int, _int, _PP],
    "ddf_op_deriv": [_P, _int, _int, _PP],
    "ddf_op_hodge": [_P, _int, _int, _PP],
    "ddf_op_invhodge": [_P, _int, _int, _PP],
    "ddf_op_coderiv": [_P, _int, _int, _PP],
    "ddf_op_laplace": [_P, _int, _int, _PP],
    "ddf_op_isboundary": [_P, _int, _int, _PP],
    "ddf_op_identity": [_P, _int, _f64, _PP],
    "ddf_op_zero": [_P, _int, _int, _PP],
    "ddf_op_from_csc": [_P, _i64, _i64, _i64p, _i64p, _f64p, _PP],
    "ddf_op_from_diag": [_P, _i64, _f64p, _PP],
    "ddf_op_destroy": [_P],
    "ddf_op_shape": [_P, _i64p, _i64p],
    "ddf_op_compose": [_P, _P, _PP],
    "ddf_op_add": [_P, _f64, _P, _PP],
    "ddf_op_scale": [_f64, _P, _PP],
    "ddf_op_div": [_P, _f64, _PP],
    "ddf_op_restrict_rows": [_P, _i64, _i64, _PP],
    "ddf_mesh_owned_rows": [_P, _int, _i64p, _i64p],
    "ddf_op_adjoint": [_P, _PP],
    "ddf_op_apply": [_P, _P, _P],
    "ddf_op_assemble": [_P, _PP],
    "ddf_mesh_deriv_format": [_P, C.c_int, C.POINTER(C.c_int)],
    "ddf_op_to_csc": [_P, _i64p, _i64p, _i64p, _f64p],
    "ddf_wave_rhs": [_P, _P, _P, _P, _P],
    "ddf_wave_dt": [_P, _f64p],
    "ddf_wave_leapfrog": [_P, _P, _P, _f64, _int],
    "ddf_wave_tsit5": [_P, _P, _P, _f64, _int],
    "ddf_poisson_jacobi": [_P, _P, _P, _f64, _int],
    "ddf_poisson_jacobi_dist": [_P, _P, _P, _f64, _int],
    "ddf_halo_mode": [_P, _intp],
    "ddf_solve_bicgstabl": [_P, _P, _P, _int, _f64, _f64, _int, _intp, _f64p, _intp],
    "ddf_vec_lincomb": [_P, _P, _f64, _int, _f64p, _PP],
    "ddf_wave_step_bytes": [_P, _i64p, _i64p],
    "ddf_wave_format": [_P, _P],
    "ddf_comm_unique_id": [_P],
    "ddf_comm_init": [_P, _int, _int, _P],
    "ddf_comm_destroy": [_P],
    "ddf_mesh_set_halo": [_P, _i64, _i64, _i64],
    "ddf_mesh_set_halo_lists": [_P, _i64, _i64, _int, C.POINTER(C.c_int32), _i64p, _i64p, _i64p, _i64p],
    "ddf_wave_leapfrog_dist": [_P, _P, _P, _f64, _int],
    "ddf_halo_exchange": [_P, _P],
    "ddf_halo_exchange_k": [_P, _P, _int],
    "ddf_comm_allreduce": [_P, _f64p, _int],
}
NON_STATUS = {"ddf_last_error": (C.c_char_p, []), "ddf_version": (_int, [])}


def load():
    """Load the shared library (idempotent).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DDFError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or ddf.jl_b200/csrc/build.sh).  ddf.jl_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _int
    for name, (res, argtypes) in NON_STATUS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = res
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        msg = _lib.ddf_last_error()
        raise DDFError(msg.decode("utf-8", "replace") if msg else f"libddf_b200 status {status}")


def call(name: str, *args):
    lib = load()
    check(getattr(lib, name)(*args))

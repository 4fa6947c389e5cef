"""Host-side mirror of DDF.jl's Manifold / Fun / Op types over libddf_b200.

Reference interfaces mirrored (paths under /root/reference):
  Manifold  src/Manifolds.jl:37-101, getters :434-690
  Fun       src/Funs.jl:16-31, arithmetic :156-188
  Op        src/Ops.jl:18-34, arithmetic :178-276
Device memory is owned by the library; these classes hold opaque handles.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import DDFError, call

Pr = True    # PrimalDual: `Pr`
Dl = False   # PrimalDual: `Dl`  (`!Pr == Dl`)


def _P(P) -> int:
    return 1 if P else 0


def _f64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class Context:
    """One GPU == one context == one process rank."""

    def __init__(self, device: int = 0):
        _lib.load()
        self._h = C.c_void_p()
        call("ddf_ctx_create", int(device), C.byref(self._h))
        self.device = device
        self.rank = 0
        self.nranks = 1
        # manifolds, cochains and operators hold pointers into this context: close() is deferred until the last of
        # them is gone (their finalizers call _release)
        self._live = 0
        self._closing = False

    def sync(self):
        call("ddf_sync", self._h)

    def bytes_in_use(self) -> int:
        out = C.c_int64()
        call("ddf_ctx_bytes_in_use", self._h, C.byref(out))
        return out.value

    def trim(self):
        """Synchronise and hand the library's cached device memory back to the driver."""
        call("ddf_ctx_trim", self._h)

    def launch_count(self) -> int:
        out = C.c_int64()
        call("ddf_ctx_launch_count", self._h, C.byref(out))
        return out.value

    def tic(self):
        call("ddf_timer_tic", self._h)

    def toc(self) -> float:
        out = C.c_double()
        call("ddf_timer_toc", self._h, C.byref(out))
        return out.value

    def event_record(self, slot: int):
        call("ddf_event_record", self._h, int(slot))

    def event_elapsed(self, a: int, b: int) -> float:
        out = C.c_double()
        call("ddf_event_elapsed", self._h, int(a), int(b), C.byref(out))
        return out.value

    def flush_l2(self):
        call("ddf_flush_l2", self._h)

    # -- NCCL communicator (one rank per process) --------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        _lib.load()
        buf = C.create_string_buffer(128)
        call("ddf_comm_unique_id", buf)
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        call("ddf_comm_init", self._h, int(nranks), int(rank), buf)
        self.rank, self.nranks = rank, nranks

    def allreduce(self, value: float, op: str = "sum") -> float:
        v = C.c_double(value)
        call("ddf_comm_allreduce", self._h, C.byref(v), {"sum": 0, "max": 1, "min": 2}[op])
        return v.value

    def _retain(self):
        self._live += 1

    def _release(self):
        self._live -= 1
        if self._closing and self._live <= 0:
            self.close()

    def close(self):
        """Destroy the context; deferred while manifolds / cochains / operators created on it are alive."""
        self._closing = True
        if self._h and self._live <= 0:
            _lib.load().ddf_ctx_destroy(self._h)
            self._h = C.c_void_p()


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def set_default_context(ctx: Context):
    global _default_ctx
    _default_ctx = ctx


class Manifold:
    """Device-resident `Manifold{D,C,Float64}` (src/Manifolds.jl:37-101)."""

    def __init__(self, handle, ctx: Context, name: str):
        self._h = handle
        self.ctx = ctx
        self.name = name
        ctx._retain()
        D, Cc = C.c_int(), C.c_int()
        call("ddf_mesh_dims", self._h, C.byref(D), C.byref(Cc))
        self.D, self.C = D.value, Cc.value
        # options of the outer constructor (Manifolds.jl:321-323)
        self.dualkind = "CircumcentricDuals"
        self.use_weighted_duals = True
        # Ops hold raw pointers into this mesh's device tables (diagonals, boundary tables): count the live ones so
        # that close() cannot leave them dangling; Fun / Op also keep `self` alive through their .manifold reference
        self._live_ops = 0
        self._closing = False

    @classmethod
    def from_simplices(cls, name: str, simplices: np.ndarray, coords: np.ndarray,
                       weights: Optional[np.ndarray] = None, ctx: Optional[Context] = None,
                       use_weighted_duals: bool = True) -> "Manifold":
        """Outer constructor `Manifold(name, simplicesD, coords0, weights;
        optimize_mesh=false)` (Manifolds.jl:319-430).  `simplices` is the 0-based
        (n_D, D+1) vertex table; it crosses the ABI in the reference's storage
        (1-based Int64 CSC of a `One` matrix)."""
        ctx = ctx or default_context()
        simplices = np.sort(np.ascontiguousarray(simplices, dtype=np.int64), axis=1)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        V, Cc = coords.shape
        n, N = simplices.shape
        D = N - 1
        colptr = 1 + N * np.arange(n + 1, dtype=np.int64)
        rowval = np.ascontiguousarray(simplices.reshape(-1) + 1)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        h = C.c_void_p()
        call("ddf_mesh_create", ctx._h, D, Cc, V, n, _i64p(colptr), _i64p(rowval), _f64p(coords),
             _f64p(w) if w is not None else None, C.byref(h))
        m = cls(h, ctx, name)
        m.use_weighted_duals = use_weighted_duals
        call("ddf_mesh_assemble", h)        # the face loop of Manifolds.jl:361-368
        return m

    # -- getters -----------------------------------------------------------------
    def nsimplices(self, R: int) -> int:
        out = C.c_int64()
        call("ddf_mesh_nsimplices", self._h, int(R), C.byref(out))
        return out.value

    def get_simplices(self, R: int) -> np.ndarray:
        """get_simplices(mfd, R) as a 0-based (n_R, R+1) table (Manifolds.jl:434-441)."""
        colptr, rowval = self.get_simplices_csc(R)
        return (rowval - 1).reshape(-1, R + 1)

    def get_simplices_csc(self, R: int):
        n = self.nsimplices(R)
        colptr = np.empty(n + 1, dtype=np.int64)
        rowval = np.empty(n * (R + 1), dtype=np.int64)
        call("ddf_mesh_get_simplices_csc", self._h, int(R), _i64p(colptr), _i64p(rowval))
        return colptr, rowval

    def get_boundaries_csc(self, R: int):
        """get_boundaries(mfd, R).op in the reference's storage (Manifolds.jl:443-450)."""
        n = self.nsimplices(R)
        colptr = np.empty(n + 1, dtype=np.int64)
        rowval = np.empty(n * (R + 1), dtype=np.int64)
        nzval = np.empty(n * (R + 1), dtype=np.int8)
        call("ddf_mesh_get_boundary_csc", self._h, int(R), _i64p(colptr), _i64p(rowval),
             nzval.ctypes.data_as(C.POINTER(C.c_int8)))
        return colptr, rowval, nzval

    def get_boundaries(self, R: int):
        import scipy.sparse as sp
        colptr, rowval, nzval = self.get_boundaries_csc(R)
        return sp.csc_matrix((nzval, rowval - 1, colptr - 1),
                             shape=(self.nsimplices(R - 1), self.nsimplices(R)))

    def get_lookup(self, Ri: int, Rj: int):
        """get_lookup(mfd, Ri, Rj) (Manifolds.jl:455-498) as a scipy CSC pattern matrix (n_Ri x n_Rj, values 1)."""
        import scipy.sparse as sp
        nnz = C.c_int64()
        call("ddf_mesh_get_lookup_csc", self._h, int(Ri), int(Rj), C.byref(nnz), None, None)
        colptr = np.empty(self.nsimplices(Rj) + 1, dtype=np.int64)
        rowval = np.empty(nnz.value, dtype=np.int64)
        call("ddf_mesh_get_lookup_csc", self._h, int(Ri), int(Rj), C.byref(nnz), _i64p(colptr), _i64p(rowval))
        return sp.csc_matrix((np.ones(nnz.value, dtype=np.int8), rowval - 1, colptr - 1),
                             shape=(self.nsimplices(Ri), self.nsimplices(Rj)))

    def refine_coords(self) -> np.ndarray:
        """refine_coords(get_lookup(mfd, 0, 1), get_coords(mfd)) (Meshing.jl:127-146), computed on the device."""
        out = np.empty((self.nsimplices(0) + self.nsimplices(1), self.C))
        call("ddf_mesh_refine_coords", self._h, _f64p(out))
        return out

    def _geometry(self):
        call("ddf_mesh_geometry", self._h, 1, 1 if self.use_weighted_duals else 0)

    def get_coords(self, R: int = 0) -> np.ndarray:
        if R == 0:
            out = np.empty((self.nsimplices(0), self.C))
            call("ddf_mesh_get_coords", self._h, _f64p(out))
            return out
        raise DDFError("get_coords(R>0) (barycentres) is not on the operator path")

    def get_weights(self) -> np.ndarray:
        out = np.empty(self.nsimplices(0))
        call("ddf_mesh_get_weights", self._h, _f64p(out))
        return out

    def get_volumes(self, R: int) -> np.ndarray:
        self._geometry()
        out = np.empty(self.nsimplices(R))
        call("ddf_mesh_get_volumes", self._h, int(R), _f64p(out))
        return out

    def get_dualvolumes(self, R: int) -> np.ndarray:
        self._geometry()
        out = np.empty(self.nsimplices(R))
        call("ddf_mesh_get_dualvolumes", self._h, int(R), _f64p(out))
        return out

    def get_dualcoords(self, R: int) -> np.ndarray:
        self._geometry()
        out = np.empty((self.nsimplices(R), self.C))
        call("ddf_mesh_get_dualcoords", self._h, int(R), _f64p(out))
        return out

    def get_isboundary(self, R: int) -> np.ndarray:
        out = np.empty(self.nsimplices(R), dtype=np.uint8)
        call("ddf_mesh_get_isboundary", self._h, int(R), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.astype(bool)

    def owned_rows(self, R: int):
        """(lo, hi): the local R-simplices this rank owns as one row range (everything on a mesh without a halo)."""
        lo, hi = C.c_int64(), C.c_int64()
        call("ddf_mesh_owned_rows", self._h, int(R), C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def set_halo(self, plane: int, ghost_lo: int, ghost_hi: int):
        call("ddf_mesh_set_halo", self._h, int(plane), int(ghost_lo), int(ghost_hi))

    def set_halo_lists(self, part):
        """Ghost lists of a mesh partitioned by vertex ranges (`ddf.dist.partition_mesh`)."""
        peers = np.ascontiguousarray(part.peers, dtype=np.int32)
        arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in (part.send_ptr, part.send_idx, part.recv_ptr, part.recv_idx)]
        call("ddf_mesh_set_halo_lists", self._h, int(part.own_lo), int(part.own_hi), int(peers.shape[0]),
             peers.ctypes.data_as(C.POINTER(C.c_int32)), *[_i64p(a) for a in arrs])

    def close(self):
        """Release the device tables.  Operators built on this manifold borrow its tables, so while any of them is
        alive the release is deferred to the moment the last one is dropped (Op.__del__)."""
        self._closing = True
        if self._h and self._live_ops <= 0:
            _lib.load().ddf_mesh_destroy(self._h)
            self._h = C.c_void_p()
            self.ctx._release()

    def __del__(self):
        try:
            if self._h and _lib._lib is not None and self.ctx._h:
                _lib._lib.ddf_mesh_destroy(self._h)
                self._h = None
                self.ctx._release()
        except Exception:
            pass


class Fun:
    """Cochain `Fun{D,P,R,C,S,Float64}` (src/Funs.jl:16-31) with device values."""

    def __init__(self, manifold: Manifold, P: bool, R: int, values=None, _handle=None):
        self.manifold = manifold
        self.P = bool(P)
        self.R = int(R)
        n = manifold.nsimplices(R)
        self._ctx = manifold.ctx
        if _handle is not None:
            self._h = _handle
        else:
            self._h = C.c_void_p()
            call("ddf_vec_create", manifold.ctx._h, n, C.byref(self._h))
        self._ctx._retain()
        if _handle is None and values is not None:
            self.upload(values)

    def __len__(self):
        return self.manifold.nsimplices(self.R)

    def upload(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.shape != (len(self),):
            raise DDFError(f"Fun: expected {len(self)} values, got {v.shape}")
        call("ddf_vec_upload", self._h, _f64p(v))
        return self

    def upload_async(self, pinned: np.ndarray):
        """Asynchronous upload from a page-locked float64 array that the caller keeps alive until ctx.sync()."""
        assert pinned.dtype == np.float64 and pinned.flags.c_contiguous and pinned.shape == (len(self),)
        call("ddf_vec_upload_async", self._h, _f64p(pinned))
        return self

    def download_async(self, pinned: np.ndarray):
        """Asynchronous download into a page-locked float64 array; valid after ctx.sync()."""
        assert pinned.dtype == np.float64 and pinned.flags.c_contiguous and pinned.shape == (len(self),)
        call("ddf_vec_download_async", self._h, _f64p(pinned))
        return pinned

    @property
    def values(self) -> np.ndarray:
        out = np.empty(len(self))
        call("ddf_vec_download", self._h, _f64p(out))
        return out

    def _like(self) -> "Fun":
        return Fun(self.manifold, self.P, self.R)

    def _check(self, g: "Fun"):
        if g.manifold is not self.manifold or g.P != self.P or g.R != self.R:
            raise DDFError("Fun arithmetic: manifold / P / R mismatch (Funs.jl:164-172)")

    # vector space (Funs.jl:156-188)
    def __pos__(self):
        return self * 1.0

    def __neg__(self):
        return self * -1.0

    def __add__(self, g):
        self._check(g)
        z = self._like()
        call("ddf_vec_axpby", 1.0, self._h, 1.0, g._h, z._h)
        return z

    def __sub__(self, g):
        self._check(g)
        z = self._like()
        call("ddf_vec_axpby", 1.0, self._h, -1.0, g._h, z._h)
        return z

    def __mul__(self, a):
        z = self._like()
        call("ddf_vec_axpby", float(a), self._h, 0.0, None, z._h)
        return z

    __rmul__ = __mul__

    def __truediv__(self, a):
        # f / a divides on the device (Funs.jl:186-188), keeping the division's rounding
        z = self._like()
        call("ddf_vec_div", self._h, float(a), z._h)
        return z

    def copy(self):
        z = self._like()
        call("ddf_vec_copy", z._h, self._h)
        return z

    def norm(self, p=2) -> float:
        kind = {1: 1, 2: 2, np.inf: 0, float("inf"): 0, "inf": 0}[p]
        out = C.c_double()
        call("ddf_vec_norm", self._h, kind, C.byref(out))
        return out.value

    def vdot(self, g) -> float:
        """Plain Euclidean dot of the value vectors, `dot(f.values, g.values)` (no weights)."""
        out = C.c_double()
        call("ddf_vec_dot", self._h, g._h, C.byref(out))
        return out.value

    def dot(self, g) -> float:
        """dot(f, g) of the reference (ManifoldOps.jl:154-164): weighted by 1 ./ vol_D for Fun{D,Pr,D} and by
        1 ./ dualvol_0 for Fun{D,Dl,0}; the reference defines no other method, so other (P, R) raise."""
        self._check(g)
        self.manifold._geometry()
        out = C.c_double()
        call("ddf_vec_wdot", self.manifold._h, _P(self.P), self.R, self._h, g._h, C.byref(out))
        return out.value

    def sum(self) -> float:
        out = C.c_double()
        call("ddf_vec_sum", self._h, C.byref(out))
        return out.value

    def __eq__(self, g):
        return (isinstance(g, Fun) and g.manifold is self.manifold and g.P == self.P and g.R == self.R
                and np.array_equal(self.values, g.values))

    __hash__ = None

    def __del__(self):
        try:
            if self._h and _lib._lib is not None and self._ctx._h:
                _lib._lib.ddf_vec_destroy(self._h)
                self._h = None
                self._ctx._release()
        except Exception:
            pass


def zero_fun(manifold: Manifold, P: bool, R: int) -> Fun:
    """zero(Fun{D,P,R,C,S,T}, mfd) (Funs.jl:132-136)."""
    return Fun(manifold, P, R)


def unit_fun(manifold: Manifold, P: bool, R: int, n: int) -> Fun:
    """unit(Fun, mfd, n), n 1-based like the reference (Funs.jl:148-154)."""
    v = np.zeros(manifold.nsimplices(R))
    v[n - 1] = 1.0
    return Fun(manifold, P, R, v)


def rand_fun(manifold: Manifold, P: bool, R: int, seed: int = 0) -> Fun:
    """rand(Fun, mfd) (Funs.jl:80-84): uniform [0,1), generated on the device."""
    f = Fun(manifold, P, R)
    call("ddf_vec_rand", f._h, C.c_uint64(seed))
    return f


class Op:
    """`Op{D,P1,R1,P2,R2,Float64}` (src/Ops.jl:18-34): maps (P2,R2)-cochains to
    (P1,R1)-cochains."""

    def __init__(self, manifold: Manifold, P1, R1, P2, R2, handle):
        self.manifold = manifold
        self.P1, self.R1, self.P2, self.R2 = bool(P1), int(R1), bool(P2), int(R2)
        self._h = handle
        manifold._live_ops += 1
        self._ctx = manifold.ctx
        self._ctx._retain()

    @property
    def shape(self):
        a, b = C.c_int64(), C.c_int64()
        call("ddf_op_shape", self._h, C.byref(a), C.byref(b))
        return (a.value, b.value)

    def _new(self, fn, *args, P1=None, R1=None, P2=None, R2=None):
        h = C.c_void_p()
        call(fn, *args, C.byref(h))
        return Op(self.manifold, self.P1 if P1 is None else P1, self.R1 if R1 is None else R1,
                  self.P2 if P2 is None else P2, self.R2 if R2 is None else R2, h)

    # Op * Fun (Ops.jl:273-276) and Op * Op (Ops.jl:228-232)
    def __mul__(self, other):
        if isinstance(other, Fun):
            if other.manifold is not self.manifold:
                raise DDFError("Op*Fun: different manifolds (Ops.jl:274)")
            if (other.P, other.R) != (self.P2, self.R2):
                raise DDFError(f"Op*Fun: operator acts on (P={self.P2},R={self.R2}) cochains, got (P={other.P},R={other.R})")
            y = Fun(self.manifold, self.P1, self.R1)
            call("ddf_op_apply", self._h, other._h, y._h)
            return y
        if isinstance(other, Op):
            if other.manifold is not self.manifold:
                raise DDFError("Op*Op: different manifolds (Ops.jl:230)")
            if (other.P1, other.R1) != (self.P2, self.R2):
                raise DDFError("Op*Op: inner (P,R) types differ")
            return self._new("ddf_op_compose", self._h, other._h, P2=other.P2, R2=other.R2)
        return self._new("ddf_op_scale", float(other), self._h)

    def __rmul__(self, a):
        return self._new("ddf_op_scale", float(a), self._h)

    def __truediv__(self, a):
        # A / a divides the assembled values (Ops.jl:210-212): not the same bits as (1/a) * A unless 1/a is exact
        return self._new("ddf_op_div", self._h, float(a))

    def __neg__(self):
        return self._new("ddf_op_scale", -1.0, self._h)

    def __pos__(self):
        return self

    def _check(self, B):
        if B.manifold is not self.manifold or (B.P1, B.R1, B.P2, B.R2) != (self.P1, self.R1, self.P2, self.R2):
            raise DDFError("Op +/-: type mismatch (Ops.jl:186-196)")

    def __add__(self, B):
        self._check(B)
        return self._new("ddf_op_add", self._h, 1.0, B._h)

    def __sub__(self, B):
        self._check(B)
        return self._new("ddf_op_add", self._h, -1.0, B._h)

    def restrict_rows(self, lo: int, hi: int) -> "Op":
        """The same operator writing rows [lo, hi) only (leaf operators: deriv(Pr,k), hodge, invhodge, isboundary) --
        what a slab rank applies to the simplices it owns."""
        return self._new("ddf_op_restrict_rows", self._h, int(lo), int(hi))

    def adjoint(self):
        """adjoint(A) (Ops.jl:258-260)."""
        return self._new("ddf_op_adjoint", self._h, P1=self.P2, R1=self.R2, P2=self.P1, R2=self.R1)

    @property
    def T(self):
        return self.adjoint()

    def apply(self, x: Fun, out: Fun) -> Fun:
        """mul!(out, A, x): apply into an existing cochain."""
        call("ddf_op_apply", self._h, x._h, out._h)
        return out

    def assemble(self) -> "Op":
        """This operator held the way the reference holds every Op: as its assembled sparse matrix (Ops.jl:25), with
        the constructor's dnz/chop after every product and sum.  `A.assemble() * f` is one row kernel."""
        return self._new("ddf_op_assemble", self._h)

    def to_csc(self):
        """Assembled `.values` with dnz/chop applied (Ops.jl:25,36-58) as a scipy CSC."""
        import scipy.sparse as sp
        nnz = C.c_int64()
        call("ddf_op_to_csc", self._h, C.byref(nnz), None, None, None)
        nr, nc = self.shape
        colptr = np.empty(nc + 1, dtype=np.int64)
        rowval = np.empty(nnz.value, dtype=np.int64)
        nzval = np.empty(nnz.value)
        call("ddf_op_to_csc", self._h, C.byref(nnz), _i64p(colptr), _i64p(rowval), _f64p(nzval))
        return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(nr, nc))

    def __del__(self):
        try:
            if self._h and _lib._lib is not None and self._ctx._h:
                _lib._lib.ddf_op_destroy(self._h)
                self._h = None
                m = self.manifold
                m._live_ops -= 1
                if m._closing and m._live_ops <= 0:
                    m.close()
                self._ctx._release()
        except Exception:
            pass


def op_from_csc(manifold: Manifold, P1, R1, P2, R2, A) -> Op:
    """Arbitrary Op from a scipy sparse matrix in the reference's storage
    (SparseMatrixCSC{Float64,Int})."""
    A = A.tocsc().astype(np.float64)
    A.sort_indices()
    colptr = A.indptr.astype(np.int64) + 1
    rowval = A.indices.astype(np.int64) + 1
    nzval = np.ascontiguousarray(A.data)
    h = C.c_void_p()
    call("ddf_op_from_csc", manifold.ctx._h, A.shape[0], A.shape[1], _i64p(colptr), _i64p(rowval),
         _f64p(nzval), C.byref(h))
    return Op(manifold, P1, R1, P2, R2, h)


def op_from_diag(manifold: Manifold, P1, R1, P2, R2, diag) -> Op:
    d = np.ascontiguousarray(diag, dtype=np.float64)
    h = C.c_void_p()
    call("ddf_op_from_diag", manifold.ctx._h, len(d), _f64p(d), C.byref(h))
    return Op(manifold, P1, R1, P2, R2, h)

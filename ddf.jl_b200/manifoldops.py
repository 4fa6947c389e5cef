"""DEC operators and manifold constructors, mirroring the reference's call surface.

  deriv / boundary / isboundary     src/ManifoldOps.jl:17-49
  hodge / invhodge                  src/ManifoldOps.jl:58-105
  coderiv / laplace                 src/ManifoldOps.jl:110-146
  integral / dot                    src/ManifoldOps.jl:150-164
  sample (R = 0)                    src/Continuum.jl:124-158
  simplex_manifold / refined_manifold / large_hypercube_manifold
                                    src/ManifoldConstructors.jl:30-74,521-529,196-284
Same names, argument meaning and error behaviour (the reference's `@assert`s on
R ranges become DDFError).  `Val(P), Val(R)` become plain arguments.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Optional, Sequence

import numpy as np

from ._lib import DDFError, call
from .core import Context, Dl, Fun, Manifold, Op, Pr, _P, default_context


def _op(fn: str, mfd: Manifold, P, R, P1, R1, P2, R2) -> Op:
    h = C.c_void_p()
    call(fn, mfd._h, _P(P), int(R), C.byref(h))
    return Op(mfd, P1, R1, P2, R2, h)


# ------------------------------------------------------------ topological operators
def boundary(P, R: int, mfd: Manifold) -> Op:
    """boundary(Val(P), Val(R), mfd) (ManifoldOps.jl:17-26)."""
    if P:
        return _op("ddf_op_boundary", mfd, P, R, Pr, R - 1, Pr, R)
    return _op("ddf_op_boundary", mfd, P, R, Dl, R + 1, Dl, R)


def isboundary(P, R: int, mfd: Manifold) -> Op:
    """isboundary(Val(Pr), Val(R), mfd) (ManifoldOps.jl:31-35)."""
    return _op("ddf_op_isboundary", mfd, P, R, Pr, R, Pr, R)


def deriv(P, R: int, mfd: Manifold) -> Op:
    """deriv(Val(P), Val(R), mfd) (ManifoldOps.jl:40-47)."""
    s = 1 if P else -1
    return _op("ddf_op_deriv", mfd, P, R, P, R + s, P, R)


# --------------------------------------------------------------- geometric operators
def hodge(P, R: int, mfd: Manifold) -> Op:
    """hodge(Val(P), Val(R), mfd) (ManifoldOps.jl:58-68, 84-91)."""
    mfd._geometry()
    return _op("ddf_op_hodge", mfd, P, R, not P, R, P, R)


def invhodge(P, R: int, mfd: Manifold) -> Op:
    """invhodge(Val(P), Val(R), mfd) (ManifoldOps.jl:71-82, 93-100)."""
    mfd._geometry()
    return _op("ddf_op_invhodge", mfd, P, R, not P, R, P, R)


def coderiv(P, R: int, mfd: Manifold) -> Op:
    """coderiv(Val(P), Val(R), mfd) (ManifoldOps.jl:110-121)."""
    mfd._geometry()
    dR = 1 if P else -1
    return _op("ddf_op_coderiv", mfd, P, R, P, R - dR, P, R)


def laplace(P, R: int, mfd: Manifold) -> Op:
    """laplace(Val(P), Val(R), mfd) (ManifoldOps.jl:128-144)."""
    mfd._geometry()
    return _op("ddf_op_laplace", mfd, P, R, P, R, P, R)


def one(P, R: int, mfd: Manifold, a: float = 1.0) -> Op:
    """one(Op{D,P,R,P,R}, mfd) (Ops.jl:216-220), optionally scaled."""
    h = C.c_void_p()
    call("ddf_op_identity", mfd._h, int(R), float(a), C.byref(h))
    return Op(mfd, P, R, P, R, h)


def zero(P1, R1, P2, R2, mfd: Manifold) -> Op:
    """zero(Op{D,P1,R1,P2,R2}, mfd) (Ops.jl:152-157)."""
    h = C.c_void_p()
    call("ddf_op_zero", mfd._h, int(R1), int(R2), C.byref(h))
    return Op(mfd, P1, R1, P2, R2, h)


# convenience forms acting on a Fun (ManifoldOps.jl:28,49,102-105,123,146)
def _apply(opfn, f: Fun) -> Fun:
    return opfn(f.P, f.R, f.manifold) * f


def boundary_of(f: Fun) -> Fun:
    return _apply(boundary, f)


def deriv_of(f: Fun) -> Fun:
    return _apply(deriv, f)


def hodge_of(f: Fun) -> Fun:
    return _apply(hodge, f)


def invhodge_of(f: Fun) -> Fun:
    return _apply(invhodge, f)


def coderiv_of(f: Fun) -> Fun:
    return _apply(coderiv, f)


def laplace_of(f: Fun) -> Fun:
    return _apply(laplace, f)


def integral(f: Fun) -> float:
    """integral(f) for (Pr, D)- and (Dl, 0)-cochains (ManifoldOps.jl:150-151)."""
    if not ((f.P and f.R == f.manifold.D) or ((not f.P) and f.R == 0)):
        raise DDFError("integral: defined for Fun{D,Pr,D} and Fun{D,Dl,0} (ManifoldOps.jl:150-151)")
    return f.sum()


# ------------------------------------------------------------------------- sampling
_FIELDS = {"wave": 0, "readme": 1, "poisson_rho": 2}


def sample(mfd: Manifold, f, P=Pr, R: int = 0) -> Fun:
    """sample(Fun{D,Pr,0,...}, f, mfd) (Continuum.jl:124-158; R = 0 only, like the
    reference's assert at :130).  `f` is one of the example fields evaluated on the
    device ("wave", "readme", "poisson_rho") or a callable coords(V,C) -> values(V)
    evaluated on the host (the reference also evaluates `f` on the host)."""
    if not P or R != 0:
        raise DDFError("sample: only P == Pr and R == 0 (Continuum.jl:127-130)")
    out = Fun(mfd, Pr, 0)
    if isinstance(f, str):
        call("ddf_vec_sample", mfd._h, _FIELDS[f], out._h)
    else:
        out.upload(np.asarray(f(mfd.get_coords()), dtype=np.float64))
    return out


# --------------------------------------------------------------------- constructors
def regular_simplex(D: int) -> np.ndarray:
    """Coordinates of the regular D-simplex with edge length 1
    (ManifoldConstructors.jl:50-74).  Host side: mesh *input*."""
    N = D + 1
    s = np.zeros((N, D))
    if D > 0:
        s0 = regular_simplex(D - 1)
        z = 1.0 if D == 1 else math.sqrt(1 - float(np.sum(s0[0] ** 2)))
        z0 = -z / (D + 1)
        for n in range(N - 1):
            s[n, :D - 1] = s0[n]
            s[n, D - 1] = z0
        s[N - 1, D - 1] = z + z0
    return s


def simplex_manifold(D: int, ctx: Optional[Context] = None) -> Manifold:
    """simplex_manifold(Val(D), Float64; optimize_mesh=false) (ManifoldConstructors.jl:30-40)."""
    coords = regular_simplex(D).reshape(D + 1, D)
    simp = np.arange(D + 1, dtype=np.int64)[None, :]
    return Manifold.from_simplices("simplex manifold", simp, coords, ctx=ctx)


def orthogonal_simplex_tables(D: int):
    """(simplices, coords, weights) of the orthogonal D-simplex with unit legs (ManifoldConstructors.jl:82-113):
    the origin and the D unit points; the origin carries the weight -1 + 1/D, the other vertices 0."""
    if D < 0:
        raise DDFError("orthogonal_simplex_manifold: D >= 0")
    coords = np.zeros((D + 1, D))
    coords[1:, :] = np.eye(D)
    weights = np.zeros(D + 1)
    if D > 0:
        weights[0] = -1 + 1.0 / D
    return np.arange(D + 1, dtype=np.int64)[None, :], coords, weights


def orthogonal_simplex_manifold(D: int, ctx: Optional[Context] = None) -> Manifold:
    """orthogonal_simplex_manifold(Val(D), Float64; optimize_mesh=false) (ManifoldConstructors.jl:82-96)."""
    simp, coords, weights = orthogonal_simplex_tables(D)
    return Manifold.from_simplices("orthogonal simplex manifold", simp, coords, weights, ctx=ctx)


def hypercube_tables(D: int):
    """(simplices, coords, weights) of the standard tesselation of the unit D-cube into D! simplices
    (ManifoldConstructors.jl:121-190): every monotone lattice path from corner 0...0 to corner 1...1 is one simplex;
    the paths are enumerated depth first, at every corner over the axes not yet taken in ascending order; a corner's
    vertex number is the binary number of its coordinates with axis 1 as the lowest bit; all weights are zero."""
    if D < 0:
        raise DDFError("hypercube_manifold: D >= 0")
    paths = []

    def walk(corner: int, verts):
        if len(verts) == D + 1:
            paths.append(list(verts))
            return
        for d in range(D):
            if not (corner >> d) & 1:
                nxt = corner | (1 << d)
                walk(nxt, verts + [nxt])

    walk(0, [0])
    assert len(paths) == math.factorial(D)
    simp = np.array(paths, dtype=np.int64).reshape(len(paths), D + 1)
    corners = np.arange(2 ** D, dtype=np.int64)
    coords = ((corners[:, None] >> np.arange(D, dtype=np.int64)[None, :]) & 1).astype(np.float64).reshape(2 ** D, D)
    return simp, coords, np.zeros(2 ** D)


def hypercube_manifold(D: int, ctx: Optional[Context] = None) -> Manifold:
    """hypercube_manifold(Val(D), Float64; optimize_mesh=false) (ManifoldConstructors.jl:121-154): the fixture of the
    reference's own test-funs.jl / test-ops.jl."""
    simp, coords, weights = hypercube_tables(D)
    return Manifold.from_simplices("hypercube manifold", simp, coords, weights, ctx=ctx)


def boundary_tables(nzval_rows: np.ndarray, faces: np.ndarray, coords: np.ndarray, weights: np.ndarray):
    """The tables of boundary_manifold (ManifoldConstructors.jl:549-598) from `keep_face = B * ones` (row sums of the
    top boundary matrix), the (D-1)-simplex table and the vertex data: the faces with a non-zero sum, their vertices
    renumbered in ascending order of the old ids."""
    keep = np.asarray(nzval_rows).reshape(-1)
    if np.any(np.abs(keep) > 1):
        raise DDFError("boundary_manifold: a face appears with total orientation beyond +-1 (ManifoldConstructors.jl:556)")
    F = np.asarray(faces, dtype=np.int64)[keep != 0]
    used = np.zeros(coords.shape[0], dtype=bool)
    used[F.reshape(-1)] = True
    old2new = np.cumsum(used) - 1
    return old2new[F], np.asarray(coords)[used], np.asarray(weights)[used]


def boundary_manifold(mfd0: Manifold) -> Manifold:
    """boundary_manifold(mfd0; optimize_mesh=false) (ManifoldConstructors.jl:549-598): the (D-1)-dimensional manifold
    of the faces whose boundary-matrix row does not cancel.  The face tables come from the device assembly; the
    selection is host-side mesh input like the reference's."""
    D = mfd0.D
    if D <= 0:
        raise DDFError("boundary_manifold: D > 0 (ManifoldConstructors.jl:550)")
    B = mfd0.get_boundaries(D)
    rows = np.asarray(B.astype(np.int64).sum(axis=1)).reshape(-1)
    faces = mfd0.get_simplices(D - 1) if D - 1 >= 1 else np.arange(mfd0.nsimplices(0), dtype=np.int64)[:, None]
    simp, coords, weights = boundary_tables(rows, faces, mfd0.get_coords(), mfd0.get_weights())
    return Manifold.from_simplices("boundary " + mfd0.name, simp, coords, weights, ctx=mfd0.ctx)


def boundary_simplex_manifold(D: int, ctx: Optional[Context] = None) -> Manifold:
    """boundary_simplex_manifold(Val(D), Float64) (ManifoldConstructors.jl:606-610): the boundary of the regular
    (D+1)-simplex -- a closed D-manifold embedded in D+1 coordinates."""
    return boundary_manifold(simplex_manifold(D + 1, ctx))


def delaunay_mesh(coords: np.ndarray) -> np.ndarray:
    """delaunay_mesh(coords) (Meshing.jl:19-49): the reference calls
    scipy.spatial.Delaunay through Delaunay.jl/PyCall; so does this."""
    from scipy.spatial import Delaunay
    if coords.shape[1] == 1:
        order = np.argsort(coords[:, 0], kind="stable")
        return np.stack([order[:-1], order[1:]], axis=1).astype(np.int64)
    return np.asarray(Delaunay(coords).simplices, dtype=np.int64)


def delaunay_hypercube_tables(D: int):
    """Coordinates of delaunay_hypercube_manifold (ManifoldConstructors.jl:372-393): the 2^D corners of a unit cube moved
    to [5/2, 7/2]^D ("re-map coordinates to avoid round-off errors"), first axis fastest; zero weights."""
    corners = np.arange(2 ** D, dtype=np.int64)
    coords = ((corners[:, None] >> np.arange(D, dtype=np.int64)[None, :]) & 1).astype(np.float64).reshape(2 ** D, D) + 2.5
    return coords, np.zeros(2 ** D)


def delaunay_hypercube_manifold(D: int, ctx: Optional[Context] = None) -> Manifold:
    """delaunay_hypercube_manifold(Val(D), Float64; optimize_mesh=false): Delaunay triangulation (scipy Qhull, as in the
    reference) of the cube's corners."""
    coords, weights = delaunay_hypercube_tables(D)
    return Manifold.from_simplices("delaunay hypercube manifold", delaunay_mesh(coords), coords, weights, ctx=ctx)


LARGE_DELAUNAY_DEFAULT_NELTS = {0: 1, 1: 1024, 2: 32, 3: 16, 4: 4, 5: 2}


def large_delaunay_hypercube_tables(D: int, nelts: Optional[int] = None):
    """Coordinates of large_delaunay_hypercube_manifold (ManifoldConstructors.jl:401-470): the (nelts+1)^D grid points
    i / nelts, first axis fastest, where for every axis d the coordinates of the LOWER axes d' < d that are not on the
    boundary are squeezed by nelts / (nelts + di), di = (1/2) / sqrt(d - 1), and shifted by di / (nelts + di) on the odd
    layers of axis d ("stagger every second point, except boundary points"); zero weights.  nelts must be a power of 2."""
    if nelts is None:
        nelts = LARGE_DELAUNAY_DEFAULT_NELTS[D]
    if nelts < 1 or (nelts & (nelts - 1)):
        raise DDFError("large_delaunay_hypercube_manifold: nelts must be a power of 2 (ManifoldConstructors.jl:413)")
    n1 = nelts + 1
    idx = np.indices((n1,) * D).reshape(D, -1, order="F").T if D > 0 else np.zeros((1, 0), dtype=np.int64)   # first axis fastest
    coords = idx.astype(np.float64) / nelts
    for d in range(1, D):                                    # 0-based axis d >= 1 (the reference's d = 2..D)
        di = 0.5 / math.sqrt(float(d))
        odd = (idx[:, d] & 1) == 1
        for dp in range(d):
            interior = (idx[:, dp] != 0) & (idx[:, dp] != nelts)
            x1 = coords[:, dp] * nelts / (nelts + di)
            x1 = np.where(odd, x1 + di / (nelts + di), x1)
            coords[:, dp] = np.where(interior, x1, coords[:, dp])
    return coords, np.zeros(coords.shape[0])


def large_delaunay_hypercube_manifold(D: int, nelts: Optional[int] = None, ctx: Optional[Context] = None) -> Manifold:
    """large_delaunay_hypercube_manifold(Val(D), Float64; nelts, optimize_mesh=false)."""
    coords, weights = large_delaunay_hypercube_tables(D, nelts)
    return Manifold.from_simplices("large delaunay hypercube manifold", delaunay_mesh(coords), coords, weights, ctx=ctx)


def refined_manifold(mfd0: Manifold) -> Manifold:
    """refined_manifold(mfd0; optimize_mesh=false) (ManifoldConstructors.jl:521-529):
    old vertices + edge midpoints in edge-id order (Meshing.jl:127-146, on the device), Delaunay (scipy Qhull on the
    host, exactly the reference's Delaunay.jl -> PyCall -> scipy backend), zero weights; the face assembly runs on the
    device."""
    if mfd0.D == 0:
        return mfd0
    newc = mfd0.refine_coords()          # device: old vertices + edge midpoints in edge-id order
    return Manifold.from_simplices("refined " + mfd0.name, delaunay_mesh(newc), newc, ctx=mfd0.ctx)


def refined_simplex_manifold(D: int, levels: int = 1, ctx: Optional[Context] = None) -> Manifold:
    m = simplex_manifold(D, ctx)
    for _ in range(levels):
        m = refined_manifold(m)
    return m


def large_hypercube_manifold(D: int, nelts: Optional[Sequence[int]] = None, ctx: Optional[Context] = None,
                             hdiv: int = 0, z0: int = 0, assemble: bool = True) -> Manifold:
    """large_hypercube_manifold(Val(D), Float64; nelts, optimize_mesh=false)
    (ManifoldConstructors.jl:196-284), generated on the device."""
    ctx = ctx or default_context()
    if nelts is None:
        nelts = (2,) * D
    ne = np.ascontiguousarray(nelts, dtype=np.int64)
    if len(ne) != D:
        raise DDFError("large_hypercube_manifold: nelts must have D entries")
    h = C.c_void_p()
    call("ddf_mesh_create_hypercube", ctx._h, int(D), ne.ctypes.data_as(C.POINTER(C.c_int64)), int(hdiv), int(z0),
         C.byref(h))
    m = Manifold(h, ctx, "large hypercube manifold")
    if assemble:
        call("ddf_mesh_assemble", h)
    return m


def deriv_format(mfd, k: int) -> str:
    """Storage of the rows of deriv(Pr,k) the apply kernels use on this mesh: "bits32" / "bits64" (one 4- / 8-byte
    word of index differences per row, field widths chosen per table, + one base per 32 rows), "packed" (the first
    face of a row is its prefix -- W-1 words per row + one base per 32 rows) or "explicit" (tables whose differences
    do not fit, e.g. top simplices given in a scattered order).  Same bits either way."""
    import ctypes as C
    from ._lib import call
    f = C.c_int()
    call("ddf_mesh_deriv_format", mfd._h, int(k), C.byref(f))
    return ("explicit", "packed", "bits32", "bits64")[f.value]


def deriv_index_bytes(fmt: str, k: int) -> float:
    """Index bytes per row of deriv(Pr,k) in storage `fmt` (see deriv_format)."""
    return {"explicit": 4.0 * (k + 2), "packed": 4.0 * (k + 1) + 0.125, "bits32": 4.125, "bits64": 8.125}[fmt]

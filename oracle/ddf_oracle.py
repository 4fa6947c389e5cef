"""CPU oracle for the DDF.jl discrete-exterior-calculus hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it, and only as the checker.  The product (``ddf.jl_b200``) never routes
through this file and fails loudly when its CUDA library is missing.

This is a numpy/scipy *restatement* of the reference's algorithm (the reference
is Julia and cannot be executed in this image: no ``julia`` binary).  Each
function cites the reference ``file:line`` it follows (paths under
``/root/reference``).  The hot arithmetic itself (``A.values * f.values.vec``)
lives in Julia's ``SparseArrays`` stdlib (Julia 1.5, ``Manifest.toml:899-901``),
which is not vendored; its published loop orders are restated in
``spmv_csc`` / ``spmv_adjoint_csc`` below and anchored on the reference's call
sites ``src/Ops.jl:273-276`` and ``src/SparseOps.jl:438-440``.

Pinning status: the reference holds NO golden vectors for d_k, the Hodge star,
the Laplacian or the wave step -- its tests are algebraic identities, counts and
closed-form known answers.  The oracle is pinned against every one of those
that touches this path (``tests/test_oracle_pins.py``): simplex counts, exact
dd=0 / boundary-boundary=0, closed-form simplex volumes, volume sums,
dual-volume sums, the hodge sign rule, |star' star - (+-1)| <= 10 eps,
|delta delta| bound, weighted-circumcentre equidistance, the hand-written
complexes of ``test/test-topologies.jl``; and against the one set of computed
values the reference stores, the weighted-circumcentre outputs of
``src/hypercube-weights.nb`` (``tests/golden/hypercube_weights.json``).
Element-wise float values beyond those are UNPINNED (parity = "GPU == this
restatement"), apart from hand-derived closed forms on the regular simplex.

Conventions: everything is 0-based internally; ``to_julia_csc`` converts to the
reference's storage (1-based Int64 ``colptr``/``rowval``).
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

Pr = True    # primal  (reference: `const Pr = true`-like PrimalDual, src/Manifolds.jl)
Dl = False   # dual

EPS = np.finfo(np.float64).eps
# chop threshold: abs2(f) < sqrt(eps^3)   (src/Ops.jl:52-58)
CHOP_ABS2 = math.sqrt(EPS ** 3)


# --------------------------------------------------------------------------- #
# Manifold container (src/Manifolds.jl:37-101)
# --------------------------------------------------------------------------- #
@dataclass
class Manifold:
    name: str
    D: int
    C: int
    coords: np.ndarray                      # (V, C) float64   _coords[0]
    weights: np.ndarray                     # (V,)   float64   _weights
    simplices: Dict[int, np.ndarray]        # R -> (n_R, R+1) int64, rows ascending
    boundaries: Dict[int, sp.csc_matrix]    # R -> (n_{R-1} x n_R) int8 CSC, sorted rows
    _cache: dict = field(default_factory=dict)

    def nsimplices(self, R: int) -> int:
        return int(self.simplices[R].shape[0])


def bitsign(b) -> int:
    """DifferentialForms.bitsign: (-1)^b  (used at ManifoldOps.jl:88,97,117)."""
    return -1 if (int(b) & 1) else 1


# --------------------------------------------------------------------------- #
# A1: calc_faces_boundaries  (src/Manifolds.jl:717-768)
# --------------------------------------------------------------------------- #
def calc_faces_boundaries(simplices: np.ndarray, nvertices: int
                          ) -> Tuple[np.ndarray, sp.csc_matrix]:
    """From the R-simplex vertex table (n x (R+1), rows ascending) build the
    (R-1)-simplex table and the boundary matrix d_R (n_{R-1} x n_R, Int8).

    Follows Manifolds.jl:725-734 (emit R+1 faces per simplex, face n leaves out
    the n-th vertex with parity (-1)^(n-1)), :735 (lexicographic sort by
    vertices), :745-762 (number unique faces in sorted order), :764-765
    (COO -> CSC)."""
    n, N = simplices.shape
    R = N - 1
    assert R >= 1
    if n == 0:
        faces = np.zeros((0, R), dtype=np.int64)
        return faces, sp.csc_matrix((0, 0), dtype=np.int8)
    face_list = []
    parent = []
    parity = []
    for k in range(N):                      # leave out vertex k (0-based)
        face_list.append(np.delete(simplices, k, axis=1))
        parent.append(np.arange(n, dtype=np.int64))
        parity.append(np.full(n, 1 if k % 2 == 0 else -1, dtype=np.int8))
    faces_all = np.concatenate(face_list, axis=0)
    parent = np.concatenate(parent)
    parity = np.concatenate(parity)
    # lexicographic order by vertex tuple (first vertex most significant)
    order = np.lexsort(faces_all.T[::-1])
    fs = faces_all[order]
    head = np.ones(len(fs), dtype=bool)
    head[1:] = np.any(fs[1:] != fs[:-1], axis=1)
    face_id = np.cumsum(head) - 1
    faces = fs[head]
    nfaces = faces.shape[0]
    B = sp.coo_matrix((parity[order], (face_id, parent[order])),
                      shape=(nfaces, n), dtype=np.int8).tocsc()
    B.sort_indices()
    return faces, B


def build_manifold(name: str, simplicesD: np.ndarray, coords: np.ndarray,
                   weights: Optional[np.ndarray] = None) -> Manifold:
    """Outer Manifold constructor with optimize_mesh=false
    (src/Manifolds.jl:319-430; the face loop is :361-368).  The sparse-column
    storage of the reference sorts the vertices of every simplex ascending
    (ManifoldConstructors.jl:274-280), restated by the row sort here."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    V, C = coords.shape
    simplicesD = np.sort(np.asarray(simplicesD, dtype=np.int64), axis=1)
    D = simplicesD.shape[1] - 1
    if weights is None:
        weights = np.zeros(V)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    simplices = {D: simplicesD}
    boundaries: Dict[int, sp.csc_matrix] = {}
    for R in range(D, 0, -1):
        nv = V
        faces, B = calc_faces_boundaries(simplices[R], nv)
        simplices[R - 1] = faces
        boundaries[R] = B
    if D == 0:
        simplices[0] = simplicesD
    return Manifold(name, D, C, coords, weights, simplices, boundaries)


def to_julia_csc(A: sp.csc_matrix):
    """Reference storage: 1-based Int64 colptr/rowval (+ nzval)."""
    A = A.tocsc()
    A.sort_indices()
    return (A.indptr.astype(np.int64) + 1, A.indices.astype(np.int64) + 1,
            A.data.copy())


def simplices_julia_csc(S: np.ndarray):
    """`_simplices[R]` as the reference stores it: n0 x n_R CSC of `One`
    (index-only), 1-based."""
    n, N = S.shape
    colptr = 1 + N * np.arange(n + 1, dtype=np.int64)
    rowval = (S.reshape(-1) + 1).astype(np.int64)
    return colptr, rowval


# --------------------------------------------------------------------------- #
# A12: synthetic-input generators
# --------------------------------------------------------------------------- #
def regular_simplex(D: int) -> np.ndarray:
    """ManifoldConstructors.jl:50-74: regular D-simplex, edge length 1."""
    N = D + 1
    s = np.zeros((N, D))
    if D > 0:
        s0 = regular_simplex(D - 1)
        if D == 1:
            z = 1.0
        else:
            z = math.sqrt(1 - float(np.sum(s0[0] ** 2)))
        z0 = -z / (D + 1)
        for n in range(N - 1):
            s[n, :D - 1] = s0[n]
            s[n, D - 1] = z0
        s[N - 1, D - 1] = z + z0
    return s


def simplex_manifold(D: int) -> Manifold:
    """ManifoldConstructors.jl:30-40."""
    coords = regular_simplex(D)
    simp = np.arange(D + 1, dtype=np.int64)[None, :]
    return build_manifold("simplex manifold", simp, coords.reshape(D + 1, D))


def orthogonal_simplex_manifold(D: int) -> Manifold:
    """ManifoldConstructors.jl:82-113 (weights: w[1] = -1 + 1/D)."""
    coords = np.zeros((D + 1, D))
    w = np.zeros(D + 1)
    if D > 0:
        w[0] = -1 + 1.0 / D
    for d in range(D):
        coords[d + 1, d] = 1.0
    simp = np.arange(D + 1, dtype=np.int64)[None, :]
    return build_manifold("orthogonal simplex manifold", simp, coords, w)


def _kuhn_permutation_simplices(D: int) -> List[List[Tuple[int, ...]]]:
    """create_hypercube / next_corner!  (ManifoldConstructors.jl:286-341):
    depth-first over axes d=1..D not yet set; one simplex per permutation, in
    lexicographic order of the axis sequence."""
    out: List[List[Tuple[int, ...]]] = []

    def rec(vertices, corner):
        if len(vertices) == D + 1:
            out.append(list(vertices))
            return
        for d in range(D):
            if not corner[d]:
                nc = list(corner)
                nc[d] = 1
                rec(vertices + [tuple(nc)], nc)

    rec([tuple([0] * D)], [0] * D)
    return out


def symmetric_hypercube(D: int) -> np.ndarray:
    """The 2^D-cell mirrored Kuhn block (ManifoldConstructors.jl:210-214,
    mirror_hypercube :343-354).  Returns (D! 2^D, D+1, D) int64."""
    simp = _kuhn_permutation_simplices(D)
    for direction in range(D):
        mirrored = []
        for s in simp:
            ms = []
            for v in s:
                v = list(v)
                v[direction] = -v[direction] + 2
                ms.append(tuple(v))
            mirrored.append(ms)
        simp = simp + mirrored
    return np.array(simp, dtype=np.int64).reshape(len(simp), D + 1, D)


LEVELWEIGHTS = {0: [0.0], 1: [0.0, 0.0], 2: [0.0, -1.0 / 6, 0.0],
                3: [0.0, -1.0 / 6, -1.0 / 6, 0.0],
                4: [0.0, -3.0 / 20, -1.0 / 5, -3.0 / 20, 0.0],
                5: [0.0, -2.0 / 15, -1.0 / 5, -1.0 / 5, -2.0 / 15, 0.0]}


def large_hypercube_tables(D: int, nelts: Sequence[int],
                           h_div: Optional[int] = None):
    """Simplex table / coords / weights of large_hypercube_manifold
    (ManifoldConstructors.jl:196-284) WITHOUT the face assembly.

    ``h_div``: the reference requires all nelts equal (:262) and divides the
    level weights by nelts[1]^2 and the coordinates by nelts.  For slab meshes
    (BASELINE config 5: unequal z extent at fixed h) ``h_div`` is the common
    divisor used for both (generalisation; equals nelts[0] when None)."""
    nelts = tuple(int(n) for n in nelts)
    assert len(nelts) == D and all(n >= 0 and n % 2 == 0 for n in nelts)
    sym = symmetric_hypercube(D)                       # (T, D+1, D)
    T = sym.shape[0]
    stride = np.array([int(np.prod([nelts[e] + 1 for e in range(d)]))
                       for d in range(D)], dtype=np.int64)
    half = [n // 2 for n in nelts]
    # CartesianIndices: first index fastest (Julia column-major), :229
    grids = np.meshgrid(*[np.arange(h, dtype=np.int64) for h in half],
                        indexing="ij")
    elts = np.stack([g.reshape(-1, order="F") for g in grids], axis=1) \
        if D > 0 else np.zeros((1, 0), dtype=np.int64)
    offs = 2 * elts                                    # (ne, D)
    verts = sym[None, :, :, :] + offs[:, None, None, :]   # (ne, T, D+1, D)
    ids = (verts * stride).sum(axis=3)                 # 0-based vertex ids
    simplices = ids.reshape(-1, D + 1)
    # coords: (elt.I - 1) ./ nelts with first index fastest (:238-243)
    gridsv = np.meshgrid(*[np.arange(n + 1, dtype=np.int64) for n in nelts],
                         indexing="ij")
    iv = np.stack([g.reshape(-1, order="F") for g in gridsv], axis=1) \
        if D > 0 else np.zeros((1, 0), dtype=np.int64)
    if h_div is None:
        assert D == 0 or all(n == nelts[0] for n in nelts)   # :262
        div = np.array(nelts, dtype=np.float64) if D > 0 else np.ones(0)
        wdiv = float(nelts[0]) if D > 0 else 1.0
    else:
        div = np.full(D, float(h_div))
        wdiv = float(h_div)
    coords = iv.astype(np.float64) / div
    lw = np.array(LEVELWEIGHTS[D], dtype=np.float64)
    if D > 0:
        lw = lw / (wdiv * wdiv)                        # levelweights ./= nelts[1]^2 (:263)
    level = (iv % 2).sum(axis=1) if D > 0 else np.zeros(1, dtype=np.int64)
    weights = lw[level]
    return np.sort(simplices, axis=1), coords, weights


def large_hypercube_manifold(D: int, nelts: Optional[Sequence[int]] = None,
                             h_div: Optional[int] = None) -> Manifold:
    """ManifoldConstructors.jl:196-284 with optimize_mesh=false."""
    if nelts is None:
        nelts = (2,) * D
    simp, coords, weights = large_hypercube_tables(D, nelts, h_div)
    return build_manifold("large hypercube manifold", simp, coords, weights)


def hypercube_manifold(D: int) -> Manifold:
    """ManifoldConstructors.jl:121-154 (single Kuhn cube, zero weights)."""
    perms = _kuhn_permutation_simplices(D)
    simp = np.array([[sum(v[d] << d for d in range(D)) for v in s]
                     for s in perms], dtype=np.int64).reshape(len(perms), D + 1)
    grids = np.meshgrid(*[np.arange(2) for _ in range(D)], indexing="ij")
    coords = np.stack([g.reshape(-1, order="F") for g in grids], axis=1
                      ).astype(np.float64) if D > 0 else np.zeros((1, 0))
    return build_manifold("hypercube manifold", simp, coords)


def refine_coords(edges: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """Meshing.jl:127-146: old vertices, then edge midpoints in edge-id order;
    midpoint = (x_a + x_b) / 2 with a<b (sum order of the sparse column)."""
    mid = (coords[edges[:, 0]] + coords[edges[:, 1]]) / 2
    return np.concatenate([coords, mid], axis=0)


def delaunay_mesh(coords: np.ndarray) -> np.ndarray:
    """Meshing.jl:19-49: Delaunay.jl 1.1.0 -> PyCall -> scipy.spatial.Delaunay
    (Manifest.toml:183-187).  scipy is importable here, so this is the same
    backend; simplex order is Qhull's."""
    from scipy.spatial import Delaunay
    V, C = coords.shape
    if C == 0:
        return np.zeros((1, 1), dtype=np.int64)
    if C == 1:
        order = np.argsort(coords[:, 0], kind="stable")
        return np.stack([order[:-1], order[1:]], axis=1).astype(np.int64)
    return np.asarray(Delaunay(coords).simplices, dtype=np.int64)


def refined_manifold(mfd0: Manifold) -> Manifold:
    """ManifoldConstructors.jl:521-529 with optimize_mesh=false, zero weights."""
    if mfd0.D == 0:
        return mfd0
    coords = refine_coords(mfd0.simplices[1], mfd0.coords)
    simp = delaunay_mesh(coords)
    return build_manifold("refined " + mfd0.name, simp, coords)


def refined_simplex_manifold(D: int, levels: int = 1) -> Manifold:
    """ManifoldConstructors.jl:537-541, repeated `levels` times
    (README.md:44-50, examples/poisson.jl:33-37)."""
    m = simplex_manifold(D)
    for _ in range(levels):
        m = refined_manifold(m)
    return m


def boundary_manifold(mfd0: Manifold) -> Manifold:
    """ManifoldConstructors.jl:549-598."""
    D = mfd0.D
    B = mfd0.boundaries[D]
    keep = np.asarray(B.astype(np.int64).sum(axis=1)).reshape(-1) != 0
    F = mfd0.simplices[D - 1][keep]
    used = np.zeros(mfd0.nsimplices(0), dtype=bool)
    used[F.reshape(-1)] = True
    old2new = np.cumsum(used) - 1
    simp = old2new[F]
    return build_manifold("boundary " + mfd0.name, simp, mfd0.coords[used],
                          mfd0.weights[used])


def manifold_from_simplices(name: str, simplices, nvertices: int,
                            C: Optional[int] = None) -> Manifold:
    """Purely topological complex (e.g. the hand-written complexes of
    test/test-topologies.jl:35-57) with dummy coordinates."""
    simp = np.asarray(simplices, dtype=np.int64)
    D = simp.shape[1] - 1
    C = D if C is None else C
    rng = np.random.default_rng(1234)
    return build_manifold(name, simp, rng.random((nvertices, C)))


# --------------------------------------------------------------------------- #
# A5: volumes  (src/Manifolds.jl:834-855, src/Algorithms.jl:102-147)
# --------------------------------------------------------------------------- #
def _gram_det_sqrt(Y: np.ndarray) -> np.ndarray:
    """norm(wedge(ys)) for Y (n, R, C): sqrt of the sum of squared RxR minors."""
    n, R, C = Y.shape
    if R == 0:
        return np.ones(n)
    tot = np.zeros(n)
    for cols in itertools.combinations(range(C), R):
        M = Y[:, :, list(cols)]
        tot = tot + np.linalg.det(M) ** 2 if R > 3 else tot + _small_det(M) ** 2
    return np.sqrt(tot)


def _small_det(M: np.ndarray) -> np.ndarray:
    R = M.shape[1]
    if R == 1:
        return M[:, 0, 0]
    if R == 2:
        return M[:, 0, 0] * M[:, 1, 1] - M[:, 0, 1] * M[:, 1, 0]
    if R == 3:
        return (M[:, 0, 0] * (M[:, 1, 1] * M[:, 2, 2] - M[:, 1, 2] * M[:, 2, 1])
                - M[:, 0, 1] * (M[:, 1, 0] * M[:, 2, 2] - M[:, 1, 2] * M[:, 2, 0])
                + M[:, 0, 2] * (M[:, 1, 0] * M[:, 2, 1] - M[:, 1, 1] * M[:, 2, 0]))
    return np.linalg.det(M)


def calc_volumes(mfd: Manifold, R: int) -> np.ndarray:
    """vol_R[i] = |wedge(x_n - x_1)| / R!, vol_0 = 1 (Manifolds.jl:838)."""
    key = ("vol", R)
    if key in mfd._cache:
        return mfd._cache[key]
    S = mfd.simplices[R]
    if R == 0:
        v = np.ones(S.shape[0])
    else:
        X = mfd.coords[S]                          # (n, R+1, C)
        Y = X[:, 1:, :] - X[:, :1, :]
        v = _gram_det_sqrt(Y) / math.factorial(R)
    mfd._cache[key] = v
    return v


# --------------------------------------------------------------------------- #
# A6: weighted circumcentres (src/Algorithms.jl:56-96, Manifolds.jl:896-913)
# --------------------------------------------------------------------------- #
def circumcentre(xs: np.ndarray, ws: Optional[np.ndarray] = None) -> np.ndarray:
    """Batched weighted circumcentre.  xs (n, N, C), ws (n, N).
    A = [2 x_i.x_j, 1; 1^T, 0], b = [x_i.x_i - w_i; 1], cc = sum c_i x_i."""
    n, N, C = xs.shape
    if ws is None:
        ws = np.zeros((n, N))
    A = np.zeros((n, N + 1, N + 1))
    A[:, :N, :N] = 2 * np.einsum("nic,njc->nij", xs, xs)
    A[:, :N, N] = 1
    A[:, N, :N] = 1
    b = np.zeros((n, N + 1))
    b[:, :N] = np.einsum("nic,nic->ni", xs, xs) - ws
    b[:, N] = 1
    c = np.linalg.solve(A, b[:, :, None])[:, :, 0]
    return np.einsum("ni,nic->nc", c[:, :N], xs)


def calc_dualcoords(mfd: Manifold, R: int, weighted: bool = True) -> np.ndarray:
    key = ("cc", R, weighted)
    if key in mfd._cache:
        return mfd._cache[key]
    S = mfd.simplices[R]
    if R == 0:
        cc = mfd.coords.copy()
    else:
        ws = mfd.weights[S] if weighted else None
        cc = circumcentre(mfd.coords[S], ws)
    mfd._cache[key] = cc
    return cc


def barycentre(xs: np.ndarray) -> np.ndarray:
    """Algorithms.jl:7-8: sum(xs)/length(xs), summed in vertex order."""
    acc = xs[:, 0, :].copy()
    for k in range(1, xs.shape[1]):
        acc = acc + xs[:, k, :]
    return acc / xs.shape[1]


# --------------------------------------------------------------------------- #
# A7: circumcentric dual volumes (src/Manifolds.jl:1160-1223, driver :644-690)
# --------------------------------------------------------------------------- #
def calc_dualvolumes(mfd: Manifold, R: int, weighted: bool = True) -> np.ndarray:
    key = ("dualvol", R, weighted)
    if key in mfd._cache:
        return mfd._cache[key]
    D = mfd.D
    if R == D:
        dv = np.ones(mfd.nsimplices(D))            # Manifolds.jl:655-656
    else:
        dv1 = calc_dualvolumes(mfd, R + 1, weighted)
        Si = mfd.simplices[R]
        Sj = mfd.simplices[R + 1]
        cci = calc_dualcoords(mfd, R, weighted)
        ccj = calc_dualcoords(mfd, R + 1, weighted)
        B = mfd.boundaries[R + 1].tocsr()          # rows = R-simplices i, cols = parents j
        B.sort_indices()
        i_of = np.repeat(np.arange(B.shape[0]), np.diff(B.indptr))
        j_of = B.indices
        Xi = mfd.coords[Si[i_of]]                  # (nnz, R+1, C)
        Xj = mfd.coords[Sj[j_of]]                  # (nnz, R+2, C)
        bcj = barycentre(Xj)
        v = bcj - Xi[:, 0, :]
        if R > 0:
            Y = Xi[:, 1:, :] - Xi[:, :1, :]        # (nnz, R, C) span of simplex i
            G = np.einsum("nrc,nsc->nrs", Y, Y)
            rhs = np.einsum("nrc,nc->nr", Y, v)
            coef = np.linalg.solve(G, rhs[:, :, None])[:, :, 0]
            v = v - np.einsum("nr,nrc->nc", coef, Y)
        nn = v / np.linalg.norm(v, axis=1)[:, None]
        h = np.einsum("nc,nc->n", ccj[j_of] - cci[i_of], nn)
        contrib = dv1[j_of] * h
        # voli += volj in parent order (ascending j), then / (D - R)
        dv = _segment_sum_sequential(contrib, B.indptr) / (D - R)
    mfd._cache[key] = dv
    return dv


def _segment_sum_sequential(vals: np.ndarray, indptr: np.ndarray) -> np.ndarray:
    """Row sums in storage order, left to right, starting from 0.0 -- the
    summation order of a sequential `for` loop (bit-exact, vectorised over rows
    by position-in-row)."""
    n = len(indptr) - 1
    out = np.zeros(n)
    lens = np.diff(indptr)
    if n == 0:
        return out
    for p in range(int(lens.max()) if len(lens) else 0):
        rows = np.nonzero(lens > p)[0]
        out[rows] = out[rows] + vals[indptr[rows] + p]
    return out


# --------------------------------------------------------------------------- #
# A9 (mask): calc_isboundary!  (src/Manifolds.jl:510-550)
# --------------------------------------------------------------------------- #
def calc_isboundary(mfd: Manifold, R: int) -> np.ndarray:
    """Boolean mask over R-simplices.  R = D-1: face appears an odd number of
    times in the columns of d_D (:520-534).  Lower R: sub-simplices of boundary
    (D-1)-faces through the lookup chain (:535-545)."""
    D = mfd.D
    assert 0 <= R < D
    BD = mfd.boundaries[D]
    cnt = np.bincount(BD.indices, minlength=BD.shape[0])
    isb_face = (cnt % 2) == 1
    if R == D - 1:
        return isb_face
    mask = np.zeros(mfd.nsimplices(R), dtype=bool)
    cur = isb_face
    for Rk in range(D - 1, R, -1):                 # walk down through |d_Rk|
        Bk = mfd.boundaries[Rk]
        cols = np.nonzero(cur)[0]
        nxt = np.zeros(Bk.shape[0], dtype=bool)
        sub = Bk[:, cols]
        nxt[np.unique(sub.indices)] = True
        cur = nxt
    mask = cur
    return mask


# --------------------------------------------------------------------------- #
# chop / dnz  (src/Ops.jl:36-58)
# --------------------------------------------------------------------------- #
def calc_lookup(mfd: Manifold, Ri: int, Rj: int) -> sp.csc_matrix:
    """get_lookup(mfd, Ri, Rj) (Manifolds.jl:455-498) as an n_Ri x n_Rj int8 pattern matrix: the reference's own chain
    -- Ri == 0: the simplex definitions (:479-481); Ri == Rj: identity (:482-485); Ri == Rj - 1: |boundaries[Rj]|
    (:486-488); Ri < Rj - 1: pattern of lookup(Ri, Rj-1) * lookup(Rj-1, Rj) (:489-493); else the transpose (:494-496)."""
    assert 0 <= Ri <= mfd.D and 0 <= Rj <= mfd.D
    key = ("lookup", Ri, Rj)
    if key in mfd._cache:
        return mfd._cache[key]
    ni, nj = mfd.nsimplices(Ri), mfd.nsimplices(Rj)
    if Ri == 0:
        S = mfd.simplices[Rj]
        A = sp.csc_matrix((np.ones(S.size, dtype=np.int8), S.reshape(-1), (Rj + 1) * np.arange(nj + 1)), shape=(ni, nj))
    elif Ri == Rj:
        A = sp.identity(ni, dtype=np.int8, format="csc")
    elif Ri == Rj - 1:
        A = abs(mfd.boundaries[Rj]).astype(np.int8).tocsc()
    elif Ri < Rj - 1:
        P = (calc_lookup(mfd, Ri, Rj - 1).astype(np.int64) @ calc_lookup(mfd, Rj - 1, Rj).astype(np.int64)).tocsc()
        P.data = np.ones_like(P.data)
        A = P.astype(np.int8)
    else:
        A = calc_lookup(mfd, Rj, Ri).T.tocsc()
    A.sort_indices()
    mfd._cache[key] = A
    return A


def chop_sparse(A: sp.spmatrix) -> sp.csc_matrix:
    """dnz for float sparse matrices: zero entries with abs2 < sqrt(eps^3),
    then dropzeros! (Ops.jl:44-46)."""
    A = A.tocsc().astype(np.float64)
    A.sort_indices()
    A.data = np.where(A.data * A.data < CHOP_ABS2, 0.0, A.data)
    A.eliminate_zeros()
    return A


# --------------------------------------------------------------------------- #
# Operators (src/ManifoldOps.jl)
# --------------------------------------------------------------------------- #
def boundary(P: bool, R: int, mfd: Manifold) -> sp.csc_matrix:
    """ManifoldOps.jl:17-26.  Pr: d_R (n_{R-1} x n_R).  Dl: d_{R+1}^T."""
    D = mfd.D
    if P == Pr:
        assert 0 < R <= D
        return mfd.boundaries[R].copy()
    assert 0 <= R < D
    return mfd.boundaries[R + 1].T.tocsc()


def deriv(P: bool, R: int, mfd: Manifold) -> sp.csc_matrix:
    """ManifoldOps.jl:40-47: adjoint(boundary(P, R+s)).  Returned as a plain
    integer matrix (n_{R+s} x n_R); the storage distinction (Adjoint-CSC vs
    CSC) only affects summation order, see spmv_* below."""
    s = 1 if P == Pr else -1
    assert 0 <= R <= mfd.D and 0 <= R + s <= mfd.D
    return boundary(P, R + s, mfd).T.tocsc()


def hodge_diag(P: bool, R: int, mfd: Manifold) -> np.ndarray:
    """ManifoldOps.jl:58-68 (Pr: dualvol ./ vol) and :84-91
    (Dl: (-1)^(R(D-R)) * vol ./ dualvol)."""
    vol = calc_volumes(mfd, R)
    dv = calc_dualvolumes(mfd, R)
    if P == Pr:
        return dv / vol
    return bitsign(R * (mfd.D - R)) * (vol / dv)


def invhodge_diag(P: bool, R: int, mfd: Manifold) -> np.ndarray:
    """ManifoldOps.jl:71-82 (Dl: vol ./ dualvol) and :93-100
    (Pr: (-1)^(R(D-R)) * dualvol ./ vol)."""
    vol = calc_volumes(mfd, R)
    dv = calc_dualvolumes(mfd, R)
    if P == Dl:
        return vol / dv
    return bitsign(R * (mfd.D - R)) * (dv / vol)


def coderiv(P: bool, R: int, mfd: Manifold) -> sp.csc_matrix:
    """ManifoldOps.jl:110-121, evaluated left to right exactly as written:
    ((bitsign(R) * invhodge(!P, R-dR)) * deriv(!P, R)) * hodge(P, R), with the
    Op constructor's chop after every product (Ops.jl:25,44-46).
    delta[i,e] = fl( (sgn * ih[i] * s_ie) * h[e] )."""
    dR = 1 if P == Pr else -1
    assert 0 <= R <= mfd.D and 0 <= R - dR <= mfd.D
    ih = bitsign(R) * invhodge_diag(not P, R - dR, mfd)
    d = deriv(not P, R, mfd).astype(np.float64)
    A = chop_sparse(sp.diags(ih).tocsc() @ d)      # products with +-1: exact
    h = hodge_diag(P, R, mfd)
    A = A.tocsc()
    A.data = A.data * h[np.repeat(np.arange(A.shape[1]), np.diff(A.indptr))]
    return chop_sparse(A)


def julia_spgemm(A: sp.csc_matrix, B: sp.csc_matrix) -> sp.csc_matrix:
    """SparseArrays.spmatmul (Gustavson, Julia 1.5): for each column j of B,
    for each stored B[k,j] (k ascending), for each stored A[i,k] (i ascending):
    first touch assigns, later touches accumulate -- sequential order in k.
    Restated with one python loop over the position-in-column (vectorised)."""
    A = A.tocsc(); B = B.tocsc()
    A.sort_indices(); B.sort_indices()
    # expand all (i,k,j) triples in (j, k, i) order
    colB = np.repeat(np.arange(B.shape[1]), np.diff(B.indptr))
    kB = B.indices
    lenA = np.diff(A.indptr)[kB]
    tj = np.repeat(colB, lenA)
    tvB = np.repeat(B.data.astype(np.float64), lenA)
    startA = np.repeat(A.indptr[kB], lenA)
    within = np.arange(len(tj)) - np.repeat(np.cumsum(lenA) - lenA, lenA)
    pa = startA + within
    ti = A.indices[pa]
    tv = A.data[pa].astype(np.float64) * tvB
    # sequential accumulation per (i,j) in encounter order
    key = tj.astype(np.int64) * A.shape[0] + ti
    order = np.argsort(key, kind="stable")
    ks = key[order]; vs = tv[order]
    head = np.ones(len(ks), dtype=bool); head[1:] = ks[1:] != ks[:-1]
    starts = np.nonzero(head)[0]
    indptr = np.append(starts, len(ks))
    lens = np.diff(indptr)
    out = vs[starts].copy()                         # first touch assigns
    for p in range(1, int(lens.max()) if len(lens) else 0):
        rows = np.nonzero(lens > p)[0]
        out[rows] = out[rows] + vs[indptr[rows] + p]
    uk = ks[starts]
    C = sp.coo_matrix((out, (uk % A.shape[0], uk // A.shape[0])),
                      shape=(A.shape[0], B.shape[1])).tocsc()
    C.sort_indices()
    return C


def laplace(P: bool, R: int, mfd: Manifold) -> sp.csc_matrix:
    """ManifoldOps.jl:128-144: zero + d_{R-dR} delta_R + delta_{R+dR} d_R, each
    product through Op*Op (Ops.jl:228-232) = SpGEMM then chop."""
    D = mfd.D
    dR = 1 if P == Pr else -1
    n = mfd.nsimplices(R)
    op = sp.csc_matrix((n, n), dtype=np.float64)
    if 0 <= R - dR <= D:
        t = chop_sparse(julia_spgemm(deriv(P, R - dR, mfd).astype(np.float64),
                                     coderiv(P, R, mfd)))
        op = chop_sparse(op + t)
    if 0 <= R + dR <= D:
        t = chop_sparse(julia_spgemm(coderiv(P, R + dR, mfd),
                                     deriv(P, R, mfd).astype(np.float64)))
        op = chop_sparse(op + t)
    return op


# --------------------------------------------------------------------------- #
# A4: SparseArrays mul! loop orders (Julia 1.5 stdlib; call sites Ops.jl:273-276)
# --------------------------------------------------------------------------- #
def spmv_adjoint_csc(parent: sp.csc_matrix, x: np.ndarray) -> np.ndarray:
    """y = adjoint(parent) * x, Julia order: for each column `col` of the parent
    tmp = 0; tmp += nzval[j]' * x[rowval[j]] (j ascending); y[col] = 0 + tmp.
    Separate multiply and add roundings (no FMA).  Bit-exact restatement."""
    P = parent.tocsc(); P.sort_indices()
    prod = P.data.astype(np.float64) * x[P.indices]
    tmp = _segment_sum_sequential(prod, P.indptr)
    return 0.0 + tmp


def spmv_csc(A: sp.csc_matrix, x: np.ndarray) -> np.ndarray:
    """y = A * x for a plain CSC, Julia order: y .= 0; for col ascending:
    y[rowval[j]] += nzval[j] * x[col].  Per output row this is a sequential sum
    over ascending columns, i.e. the row-major traversal of the CSR form."""
    R = A.tocsr(); R.sort_indices()
    prod = R.data.astype(np.float64) * x[R.indices]
    return _segment_sum_sequential(prod, R.indptr)


def apply_deriv(P: bool, R: int, mfd: Manifold, x: np.ndarray) -> np.ndarray:
    """deriv(Val(P),Val(R),mfd) * f with the storage the reference ends up
    with: primal d_R is Adjoint{CSC(d_{R+1})} (ManifoldOps.jl:46 via :20),
    dual d_R is adjoint(adjoint(CSC)) = plain CSC d_R (:25,46)."""
    if P == Pr:
        return spmv_adjoint_csc(mfd.boundaries[R + 1], x)
    return spmv_csc(mfd.boundaries[R], x)


def apply_diag(d: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Diagonal * Vector: y_i = d_i * x_i."""
    return d * x


# --------------------------------------------------------------------------- #
# A12: sample R=0 (src/Continuum.jl:124-158) and the example fields
# --------------------------------------------------------------------------- #
def sample0(mfd: Manifold, f) -> np.ndarray:
    return np.asarray(f(mfd.coords), dtype=np.float64)


def wave_u0(coords: np.ndarray) -> np.ndarray:
    """examples/wave.jl:42-47 at t=0: prod_d sinpi(x_d) (k=1)."""
    return np.prod(np.sin(np.pi * coords), axis=1)


def readme_field(coords: np.ndarray) -> np.ndarray:
    """README.md:63-65: sin(x1) * cos(x2)."""
    return np.sin(coords[:, 0]) * np.cos(coords[:, 1])


def poisson_rho(coords: np.ndarray) -> np.ndarray:
    """examples/poisson.jl:85-92: Gaussian, x0 = 0.1, sigma = 0.2."""
    D = coords.shape[1]
    sigma = 0.2
    r = np.linalg.norm(coords - 0.1, axis=1)
    rp = r / (math.sqrt(2.0) * sigma)
    return 1 / math.sqrt(2 * math.pi * sigma ** 2) ** D * np.exp(-rp ** 2)


# --------------------------------------------------------------------------- #
# A10: wave right-hand side and the leapfrog step
# --------------------------------------------------------------------------- #
def wave_operators(mfd: Manifold):
    """examples/wave.jl:82-85: B = isboundary(Pr,0), L = laplace(Pr,0)."""
    b = calc_isboundary(mfd, 0)
    L = laplace(Pr, 0, mfd)
    return b, L


def wave_F(mfd: Manifold) -> sp.csc_matrix:
    """examples/wave.jl:98-99: F = blockdiag(E-B, E-B) * [0 E; L 0]."""
    b, L = wave_operators(mfd)
    V = mfd.nsimplices(0)
    EB = sp.diags(1.0 - b.astype(np.float64)).tocsc()
    Z = sp.csc_matrix((V, V))
    F0 = sp.bmat([[Z, sp.identity(V, format="csc")], [L, Z]], format="csc")
    P = sp.bmat([[EB, Z], [Z, EB]], format="csc")
    F = julia_spgemm(P, F0)
    F.eliminate_zeros()
    return F


def wave_rhs(mfd: Manifold, u: np.ndarray, udot: np.ndarray):
    """examples/wave.jl:87-96: du = (E-B) udot, dudot = (E-B)(L u)."""
    b, L = wave_operators(mfd)
    m = 1.0 - b.astype(np.float64)
    return m * udot, m * spmv_csc(L, u)


def wave_dt(mfd: Manifold) -> float:
    """examples/wave.jl:106,118: dt = minimum(vol_1) / 2."""
    return float(calc_volumes(mfd, 1).min()) / 2


def wave_leapfrog(mfd: Manifold, u: np.ndarray, v: np.ndarray, dt: float,
                  nsteps: int):
    """north_star's fused step on the reference's operators (NOT in the
    reference, which integrates with Tsit5, wave.jl:134):
        v += dt * (1-b) .* (L u);   u += dt * (1-b) .* v
    (symplectic Euler / kick-drift leapfrog)."""
    b, L = wave_operators(mfd)
    m = 1.0 - b.astype(np.float64)
    u = u.copy(); v = v.copy()
    for _ in range(nsteps):
        v = v + dt * (m * spmv_csc(L, u))
        u = u + dt * (m * v)
    return u, v


# --------------------------------------------------------------------------- #
# N1: the integrator the reference actually runs -- fixed-step Tsit5
# (examples/wave.jl:134: solve(prob, Tsit5(); adaptive=false, dt=dt)).
# OrdinaryDiffEq 5.49.1 (Manifest.toml:711-715) is not vendored in the reference; this
# restates its published algorithm: the Tsitouras 5(4) tableau (Ch. Tsitouras, Comput.
# Math. Appl. 62 (2011) 770-775, Float64 coefficients as in OrdinaryDiffEq's
# tsit_tableaus.jl) and the in-place perform_step!:
#     tmp = uprev + dt*(a_s1 k1 + ... + a_s,s-1 k_{s-1});  k_s = f(tmp)
#     u   = uprev + dt*(a71 k1 + ... + a76 k6);  k7 = f(u) is k1 of the next step (FSAL)
# The reference evaluates these under @muladd (FMA contraction is hardware dependent), so
# element-wise parity with Julia is a 1e-12-class statement; here every product and sum is
# separately rounded, left to right, and the device kernel does exactly the same.
# tests/test_oracle_pins.py pins the tableau through the order conditions up to order 5.
# --------------------------------------------------------------------------- #
TSIT5_C = (0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0)
TSIT5_A = (
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
    (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774),
)


def lincomb(base: np.ndarray, dt: float, coefs, ks) -> np.ndarray:
    """base + dt*(c1 k1 + c2 k2 + ...), every product and sum separately rounded, left to right."""
    s = coefs[0] * ks[0]
    for c, k in zip(coefs[1:], ks[1:]):
        s = s + c * k
    return base + dt * s


def tsit5_steps(f, U, dt: float, nsteps: int):
    """`nsteps` fixed steps of Tsit5 on dU/dt = f(U) (autonomous), U a tuple of arrays."""
    k = [None] * 7
    k[0] = f(U)
    for _ in range(nsteps):
        for s in range(6):
            tmp = tuple(lincomb(U[c], dt, TSIT5_A[s], [k[j][c] for j in range(s + 1)]) for c in range(len(U)))
            k[s + 1] = f(tmp)
            if s == 5:
                U = tmp
        k[0] = k[6]
    return U


def wave_tsit5(mfd: Manifold, u: np.ndarray, udot: np.ndarray, dt: float, nsteps: int):
    """examples/wave.jl:98-100,134: dU = F U integrated with fixed-step Tsit5."""
    return tsit5_steps(lambda U: wave_rhs(mfd, U[0], U[1]), (u.copy(), udot.copy()), dt, nsteps)


# --------------------------------------------------------------------------- #
# N2: the Poisson system of examples/poisson.jl:124-227, block form as in the reference
# --------------------------------------------------------------------------- #
def poisson_u_exact(coords: np.ndarray) -> np.ndarray:
    """examples/poisson.jl:93-122 (D = 1, 2, 3): free-space potential of poisson_rho."""
    from scipy.special import erf, expi
    D = coords.shape[1]
    sigma = 0.2
    r = np.sqrt(np.sum((coords - 0.1) ** 2, axis=1))
    rp = r / (np.sqrt(2.0) * sigma)
    if D == 1:
        return sigma / np.sqrt(2 * np.pi) * (np.exp(-rp ** 2) - 1) + r / 2 * erf(rp)
    if D == 2:
        small = r ** 8 <= np.finfo(np.float64).eps
        rs = np.where(small, 1.0, rp)
        big = -1 / (4 * np.pi) * (-np.euler_gamma + expi(-rs ** 2) + np.log(1 / rs ** 2))
        ser = 1 / (4 * np.pi) * rp ** 2 - 1 / (16 * np.pi) * rp ** 4 + 1 / (72 * np.pi) * rp ** 6
        return np.where(small, ser, big)
    if D == 3:
        rs = np.where(r == 0, 1.0, r)
        return np.where(r == 0, -1 / (4 * np.pi) * 2 / (np.sqrt(2 * np.pi) * sigma), -1 / (4 * np.pi * rs) * erf(rp))
    raise ValueError("D <= 3")


def poisson_system(mfd: Manifold, rho: np.ndarray, u0: np.ndarray):
    """A', b' of examples/poisson.jl:172-201:
        A = [0 delta; -d 1],  b = [rho; 0],  B = [B0 0; 0 0],  c = [u0; 0],
        A' = (E - B) A + B,   b' = (E - B) b + B c."""
    n0, n1 = mfd.nsimplices(0), mfd.nsimplices(1)
    d = deriv(Pr, 0, mfd).astype(np.float64)
    delta = coderiv(Pr, 1, mfd)
    B0 = sp.diags(calc_isboundary(mfd, 0).astype(np.float64))
    A = sp.bmat([[None, delta], [-d, sp.identity(n1)]], format="csc")
    B = sp.bmat([[B0, None], [None, sp.csc_matrix((n1, n1))]], format="csc")
    E = sp.identity(n0 + n1, format="csc")
    b = np.concatenate([rho, np.zeros(n1)])
    c = np.concatenate([u0, np.zeros(n1)])
    Ap = ((E - B) @ A + B).tocsc()
    bp = (E - B) @ b + B @ c
    return Ap, bp


def poisson_jacobi(mfd: Manifold, u: np.ndarray, rho: np.ndarray, theta: float, nsweeps: int) -> np.ndarray:
    """north_star's fused Poisson step (NOT in the reference, which solves with UMFPACK / BiCGStab(l)): weighted Jacobi
    on L u = rho with the boundary values held fixed (the Dirichlet problem of poisson.jl:75-81),
        u <- u + theta * ((1-b) .* ((rho - L u) ./ diag(L))),   every operation separately rounded."""
    b, L = wave_operators(mfd)
    m = 1.0 - b.astype(np.float64)
    d = L.diagonal()
    dinv = np.where(d != 0.0, 1.0 / np.where(d != 0.0, d, 1.0), 0.0)
    u = u.copy()
    for _ in range(nsweeps):
        u = u + theta * (m * ((rho - spmv_csc(L, u)) * dinv))
    return u


def poisson_solve(mfd: Manifold, rho: np.ndarray, u0: np.ndarray):
    """x = A' \\ b' (poisson.jl:205-206; UMFPACK there, SuperLU here) split into (u, f)."""
    import scipy.sparse.linalg as spla
    Ap, bp = poisson_system(mfd, rho, u0)
    x = spla.spsolve(Ap, bp)
    n0 = mfd.nsimplices(0)
    return x[:n0], x[n0:]


# --------------------------------------------------------------------------- #
# invariant (src/Manifolds.jl:154-189)
# --------------------------------------------------------------------------- #
def invariant(mfd: Manifold) -> bool:
    D = mfd.D
    for R in range(0, D + 1):
        if mfd.simplices[R].shape[1] != R + 1:
            return False
    for R in range(1, D + 1):
        B = mfd.boundaries[R]
        if B.shape != (mfd.nsimplices(R - 1), mfd.nsimplices(R)):
            return False
        if not np.all(np.abs(B.data) == 1):
            return False
        SR, SR1 = mfd.simplices[R], mfd.simplices[R - 1]
        cols = np.repeat(np.arange(B.shape[1]), np.diff(B.indptr))
        fv = SR1[B.indices]                      # (nnz, R)
        pv = SR[cols]                            # (nnz, R+1)
        ok = (fv[:, :, None] == pv[:, None, :]).any(axis=2).all()
        if not ok:
            return False
    return True

"""Pins the CPU oracle (oracle/ddf_oracle.py + its C twin) against every known-answer
test, count, identity and hand-written fixture the reference's own tests hold for the
operator path.  The reference has no golden vectors for d_k / star_k / L;
what it has is restated here one for one, with the reference test cited.

CPU only (no GPU needed): runs in the `-m "not gpu"` tier.
"""
import itertools
import math

import os

import numpy as np
import pytest
import scipy.sparse as sp

import c_oracle
import ddf_oracle as o

EPS = np.finfo(float).eps


def binom(n, k):
    return math.comb(n, k)


# ----------------------------------------------------------------- test-manifolds.jl
@pytest.mark.parametrize("D", [0, 1, 2, 3, 4])
def test_simplex_manifold_counts_and_volumes(D):
    """test/test-manifolds.jl:20-49."""
    m = o.simplex_manifold(D)
    assert o.invariant(m)
    for R in range(D + 1):
        assert m.nsimplices(R) == binom(D + 1, R + 1)                       # :24-26
    for i, j in itertools.combinations(range(D + 1), 2):
        assert abs(np.linalg.norm(m.coords[i] - m.coords[j]) - 1) < 1e-14   # :30-32 edge length 1
    if D <= 3:
        for R in range(D + 1):
            vol = math.sqrt((R + 1) / 2 ** R) / math.factorial(R)           # :36
            assert np.allclose(o.calc_volumes(m, R), vol, rtol=1e-13)       # :37
            assert np.all(o.calc_dualvolumes(m, R) > 0)                     # :39
        vol = math.sqrt((D + 1) / 2 ** D) / math.factorial(D)
        assert np.all(o.calc_volumes(m, 0) == 1)                            # :45
        assert abs(o.calc_volumes(m, D).sum() - vol) < 1e-14                # :46
        assert abs(o.calc_dualvolumes(m, 0).sum() - vol) < 1e-13            # :47
        assert np.all(o.calc_dualvolumes(m, D) == 1)                        # :48


@pytest.mark.parametrize("D", [1, 2, 3])
def test_orthogonal_simplex(D):
    """test/test-manifolds.jl:51-76: sum vol = 1/D!, sum dualvol_0 = 1/D!."""
    m = o.orthogonal_simplex_manifold(D)
    vol = 1 / math.factorial(D)
    assert abs(o.calc_volumes(m, D).sum() - vol) < 1e-14
    assert abs(o.calc_dualvolumes(m, 0).sum() - vol) < 1e-13
    if D <= 2:
        for R in range(D + 1):
            assert np.all(o.calc_dualvolumes(m, R) > 0)


@pytest.mark.parametrize("D", [1, 2, 3])
def test_hypercube_manifold(D):
    """test/test-manifolds.jl:78-101."""
    m = o.hypercube_manifold(D)
    assert o.invariant(m)
    assert m.nsimplices(0) == 2 ** D and m.nsimplices(D) == math.factorial(D)
    assert abs(o.calc_volumes(m, D).sum() - 1) < 1e-14
    assert abs(o.calc_dualvolumes(m, 0).sum() - 1) < 1e-13


@pytest.mark.parametrize("D", [1, 2, 3, 4])
def test_large_hypercube_counts(D):
    """test/test-manifolds.jl:103-132 with n = 4."""
    n = 4
    m = o.large_hypercube_manifold(D, (n,) * D)
    assert o.invariant(m)                                                    # :108
    assert m.nsimplices(0) == (n + 1) ** D                                   # :109
    assert m.nsimplices(D) == n ** D * math.factorial(D)                     # :110
    if D <= 3:
        for R in range(D + 1):
            v, dv = o.calc_volumes(m, R), o.calc_dualvolumes(m, R)
            assert v.min() / v.max() >= 0.01                                 # :114
            assert np.all(dv > 0) and dv.min() / dv.max() >= 0.01            # :117-120
        assert np.all(o.calc_volumes(m, 0) == 1)                             # :125
        assert abs(o.calc_volumes(m, D).sum() - 1) < 1e-13                   # :126
        assert abs(o.calc_dualvolumes(m, 0).sum() - 1) < 1e-13               # :130
        assert np.all(o.calc_dualvolumes(m, D) == 1)                         # :131


def test_large_hypercube_closed_forms():
    """2D V=(n+1)^2, E=3n^2+2n, T=2n^2; 3D V=(n+1)^3, E=7n^3+9n^2+3n, F=12n^3+6n^2, T=6n^3."""
    for n in (2, 4, 6):
        m = o.large_hypercube_manifold(2, (n, n))
        assert [m.nsimplices(R) for R in range(3)] == [(n + 1) ** 2, 3 * n * n + 2 * n, 2 * n * n]
    for n in (2, 4):
        m = o.large_hypercube_manifold(3, (n,) * 3)
        assert [m.nsimplices(R) for R in range(4)] == [(n + 1) ** 3, 7 * n ** 3 + 9 * n * n + 3 * n,
                                                       12 * n ** 3 + 6 * n * n, 6 * n ** 3]


@pytest.mark.parametrize("D", [1, 2, 3])
def test_refined_simplex(D):
    """test/test-manifolds.jl:241-275."""
    m = o.refined_simplex_manifold(D)
    assert o.invariant(m)
    assert m.nsimplices(0) == binom(D + 1, 1) + binom(D + 1, 2)              # :245
    assert m.nsimplices(D) == 2 ** D                                         # :246
    for R in (0, D):
        vol = math.sqrt((R + 1) / 2 ** R) / (math.factorial(R) * 2 ** R)     # :252
        assert np.allclose(o.calc_volumes(m, R), vol, rtol=1e-12)
    vol = math.sqrt((D + 1) / 2 ** D) / math.factorial(D)
    assert abs(o.calc_volumes(m, D).sum() - vol) < 1e-13                     # :272
    assert abs(o.calc_dualvolumes(m, 0).sum() - vol) < 1e-13                 # :273


@pytest.mark.parametrize("D", [1, 2])
def test_boundary_simplex(D):
    """test/test-manifolds.jl:277-299: boundary of the (D+1)-simplex."""
    m = o.boundary_manifold(o.simplex_manifold(D + 1))
    assert o.invariant(m)
    for R in range(D + 1):
        assert m.nsimplices(R) == binom(D + 2, R + 1)                        # :282
        vol = math.sqrt((R + 1) / 2 ** R) / math.factorial(R)
        assert np.allclose(o.calc_volumes(m, R), vol, rtol=1e-13)            # :293


def test_config_sizes():
    """BASELINE config 1: 2D regular simplex refined 4x -> 153/408/256."""
    m = o.refined_simplex_manifold(2, 4)
    assert [m.nsimplices(R) for R in range(3)] == [153, 408, 256]


# ------------------------------------------------------------ test-manifoldops.jl
ZOO = [("simplex", 1), ("simplex", 2), ("simplex", 3), ("kuhn", 1), ("kuhn", 2), ("kuhn", 3), ("refined", 2),
       ("refined", 3), ("bsimplex", 2)]


def make(kind, D):
    if kind == "simplex":
        return o.simplex_manifold(D)
    if kind == "kuhn":
        return o.large_hypercube_manifold(D)
    if kind == "refined":
        return o.refined_simplex_manifold(D)
    return o.boundary_manifold(o.simplex_manifold(D + 1))


@pytest.mark.parametrize("kind,D", ZOO)
def test_boundary_boundary_and_dd_zero(kind, D):
    """iszero(b2*b1), iszero(d2*d1) exactly in integers (test-manifoldops.jl:38-46, 60-68)."""
    m = make(kind, D)
    for P in (o.Pr, o.Dl):
        dR = 1 if P else -1
        for R in range(D + 1):
            if 0 <= R - 2 * dR <= D:
                b21 = o.boundary(P, R - dR, m).astype(np.int64) @ o.boundary(P, R, m).astype(np.int64)
                assert b21.nnz == 0 or np.abs(b21.data).max() == 0
            if 0 <= R + 2 * dR <= D:
                d21 = o.deriv(P, R + dR, m).astype(np.int64) @ o.deriv(P, R, m).astype(np.int64)
                assert d21.nnz == 0 or np.abs(d21.data).max() == 0


@pytest.mark.parametrize("kind,D", ZOO)
def test_constant_sign_pattern(kind, D):
    """every column of d_R carries the pattern (-1)^(R+p) (p 0-based, rows ascending)."""
    m = make(kind, D)
    for R in range(1, D + 1):
        B = m.boundaries[R]
        assert np.all(np.diff(B.indptr) == R + 1)
        pat = B.data.reshape(-1, R + 1)
        expect = np.array([(-1) ** (R + p) for p in range(R + 1)], dtype=np.int8)
        assert np.all(pat == expect)


@pytest.mark.parametrize("kind,D", [("simplex", 2), ("simplex", 3), ("kuhn", 2), ("kuhn", 3), ("bsimplex", 2)])
def test_hodge_rules(kind, D):
    """Well-centred fixtures only (test-manifoldops.jl:86-88): star finite & nonzero (:91-94),
    invhodge == +-hodge sign rule (:95-98), |star' star - s 1| <= 10 eps (:99-105),
    |delta delta| <= |delta||delta| 10 eps (:116-126), entries of delta, Laplace finite & nonzero."""
    m = make(kind, D)
    for P in (o.Pr, o.Dl):
        dR = 1 if P else -1
        for R in range(D + 1):
            h, hp = o.hodge_diag(P, R, m), o.hodge_diag(not P, R, m)
            assert np.all(np.isfinite(h)) and np.all(h != 0)
            s = o.bitsign(0 if D % 2 else (R if P else D - R))
            assert np.array_equal(o.invhodge_diag(P, R, m), s * h)
            s2 = o.bitsign(R * (D - R))
            assert np.max(np.abs(hp * h - s2)) <= 10 * EPS
            if 0 <= R - 2 * dR <= D:
                d1, d2 = o.coderiv(P, R, m), o.coderiv(P, R - dR, m)
                n1 = abs(d1).sum(axis=1).max()
                n2 = abs(d2).sum(axis=1).max()
                dd = (d2 @ d1)
                assert (abs(dd).sum(axis=1).max() if dd.nnz else 0.0) <= n1 * n2 * 10 * EPS
            if 0 <= R - dR <= D:
                delta = o.coderiv(P, R, m)
                assert np.all(np.isfinite(delta.data)) and np.all(delta.data != 0)
            L = o.laplace(P, R, m)
            assert np.all(np.isfinite(L.data)) and np.all(L.data != 0)


def test_kuhn_hodge_fingerprints():
    """Closed-form values on the weighted 2D Kuhn mesh (SURVEY 8c fingerprints; exact rationals):
    star_1 in {1/6, 5/12, 5/6}; interior dualvol_0 * n^2 in {25/36, 47/36}."""
    n = 4
    m = o.large_hypercube_manifold(2, (n, n))
    h1 = np.unique(np.round(o.hodge_diag(o.Pr, 1, m), 13))
    assert np.allclose(h1, [1 / 6, 5 / 12, 5 / 6], rtol=1e-12)
    b = o.calc_isboundary(m, 0)
    dv0 = np.unique(np.round(o.calc_dualvolumes(m, 0)[~b] * n * n, 13))
    assert np.allclose(dv0, [25 / 36, 47 / 36], rtol=1e-12)
    m3 = o.large_hypercube_manifold(3, (4, 4, 4))
    mins = [o.calc_dualvolumes(m3, R).min() for R in range(4)]
    assert np.allclose(mins, [2.9975e-3, 2.2553e-3, 2.9463e-2, 1.0], rtol=1e-4)


# ------------------------------------------------------------ test-algorithms.jl
def test_volume_known_answers():
    """test/test-algorithms.jl:47-92: the orthogonal unit simplex has volume 1/(N-1)!,
    independent of vertex order; scaling one vertex by 5 scales the volume by 5."""
    for D in (1, 2, 3):
        for N in range(1, D + 2):
            x = np.zeros((N, D))
            for n_ in range(1, N):
                x[n_, n_ - 1] = 1.0
            m = o.build_manifold("t", np.arange(N)[None, :], x) if N == D + 1 else None
            Y = (x[1:] - x[:1])[None]
            assert abs(o._gram_det_sqrt(Y)[0] / math.factorial(N - 1) - 1 / math.factorial(N - 1)) < 1e-15
            if N >= 2:
                x2 = x.copy()
                x2[-1] *= 5
                Y2 = (x2[1:] - x2[:1])[None]
                assert abs(o._gram_det_sqrt(Y2)[0] / math.factorial(N - 1) - 5 / math.factorial(N - 1)) < 1e-14
                xp = x[::-1].copy()
                Yp = (xp[1:] - xp[:1])[None]
                assert abs(o._gram_det_sqrt(Yp)[0] - o._gram_det_sqrt(Y)[0]) < 1e-15
            del m


def test_circumcentre_properties():
    """test/test-algorithms.jl:7-45: (weighted) circumcentre is equidistant (|x_n-cc|^2 - w_n constant)
    and lies in the affine hull of the vertices."""
    rng = np.random.default_rng(0)
    for D in (1, 2, 3):
        for N in range(1, D + 2):
            xs = rng.random((50, N, D))
            ws = rng.random((50, N)) * 0.1
            for w in (None, ws):
                cc = o.circumcentre(xs, w)
                r2 = ((xs - cc[:, None, :]) ** 2).sum(axis=2) - (0 if w is None else w)
                assert np.max(np.abs(r2 - r2[:, :1])) < 1e-9
                # in the affine hull: residual of the least-squares fit of cc - x_1 on the edge vectors
                if N > 1:
                    for i in range(5):
                        Y = (xs[i, 1:] - xs[i, :1]).T
                        res = np.linalg.lstsq(Y, cc[i] - xs[i, 0], rcond=None)
                        fit = Y @ res[0]
                        assert np.max(np.abs(fit - (cc[i] - xs[i, 0]))) < 1e-9


# ------------------------------------------------------------ test-topologies.jl
def test_handwritten_complexes():
    """test/test-topologies.jl:33-57: Moebius strip (6/12/6) and tetrahedron surface (4/6/4)."""
    moebius = np.array([(1, 2, 4), (1, 4, 6), (4, 3, 6), (6, 3, 5), (3, 1, 5), (1, 2, 5)]) - 1
    m = o.manifold_from_simplices("moebius", moebius, 6, C=3)
    assert [m.nsimplices(R) for R in range(3)] == [6, 12, 6]
    assert (m.boundaries[1].astype(np.int64) @ m.boundaries[2].astype(np.int64)).nnz == 0 or \
        np.abs((m.boundaries[1].astype(np.int64) @ m.boundaries[2].astype(np.int64)).data).max() == 0
    tet = np.array([(1, 2, 3), (1, 3, 4), (1, 4, 2), (2, 3, 4)]) - 1
    m = o.manifold_from_simplices("tet surface", tet, 4, C=3)
    assert [m.nsimplices(R) for R in range(3)] == [4, 6, 4]
    z = m.boundaries[1].astype(np.int64) @ m.boundaries[2].astype(np.int64)
    assert z.nnz == 0 or np.abs(z.data).max() == 0
    # closed surface: no boundary edges; Moebius strip: one boundary circle of 6 edges
    assert not o.calc_isboundary(m, 1).any()
    mm = o.manifold_from_simplices("moebius", moebius, 6, C=3)
    assert o.calc_isboundary(mm, 1).sum() == 6


# ------------------------------------------------------- test-ops.jl / test-sparseops.jl
def test_op_fun_laws_exact():
    """(F*A)*f == F*(A*f), A*(f+g) == A*f + A*g with exact (integer-valued) data
    (test-ops.jl:229-260), adjoint laws (test-sparseops.jl:126-142)."""
    m = o.large_hypercube_manifold(3, (2, 2, 2))
    rng = np.random.default_rng(1)
    f = rng.integers(-9, 9, m.nsimplices(0)).astype(float)
    g = rng.integers(-9, 9, m.nsimplices(0)).astype(float)
    assert np.array_equal(o.apply_deriv(o.Pr, 0, m, f + g), o.apply_deriv(o.Pr, 0, m, f) + o.apply_deriv(o.Pr, 0, m, g))
    d10 = (o.deriv(o.Pr, 1, m).astype(np.int64) @ o.deriv(o.Pr, 0, m).astype(np.int64)).astype(float)
    lhs = d10 @ f if d10.nnz else np.zeros(m.nsimplices(2))
    assert np.array_equal(lhs, o.apply_deriv(o.Pr, 1, m, o.apply_deriv(o.Pr, 0, m, f)))
    A = o.deriv(o.Pr, 0, m)
    assert (A.T.T != A).nnz == 0
    h = rng.integers(-9, 9, m.nsimplices(1)).astype(float)
    assert np.dot(o.apply_deriv(o.Pr, 0, m, f), h) == np.dot(f, o.spmv_csc(m.boundaries[1], h))


def test_spmv_orders_agree_with_scipy():
    """The sequential-order restatements equal the mathematical product (to rounding)."""
    m = o.large_hypercube_manifold(3, (4, 4, 4))
    rng = np.random.default_rng(2)
    for k in range(3):
        x = rng.standard_normal(m.nsimplices(k))
        ref = m.boundaries[k + 1].T.astype(float) @ x
        assert np.allclose(o.apply_deriv(o.Pr, k, m, x), ref, rtol=0, atol=1e-13)
        xd = rng.standard_normal(m.nsimplices(k + 1))
        ref = m.boundaries[k + 1].astype(float) @ xd
        assert np.allclose(o.apply_deriv(o.Dl, k + 1, m, xd), ref, rtol=0, atol=1e-13)
    A = sp.random(40, 30, density=0.3, random_state=3, format="csc")
    B = sp.random(30, 20, density=0.3, random_state=4, format="csc")
    assert abs(o.julia_spgemm(A, B) - A @ B).max() < 1e-14


def test_c_twin_matches_numpy_oracle():
    """oracle/ddf_oracle.c == oracle/ddf_oracle.py bit for bit (assembly, SpMV orders, leapfrog)."""
    for D, n in ((2, 6), (3, 4)):
        m = o.large_hypercube_manifold(D, (n,) * D)
        for R in range(D, 0, -1):
            f, cp, rv, nz = c_oracle.calc_faces_boundaries(m.simplices[R])
            ocp, orv, onz = o.to_julia_csc(m.boundaries[R])
            assert np.array_equal(f, m.simplices[R - 1])
            assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz)
        rng = np.random.default_rng(D)
        for k in range(D):
            cp, rv, nz = o.to_julia_csc(m.boundaries[k + 1])
            x = rng.standard_normal(m.nsimplices(k))
            assert np.array_equal(c_oracle.spmv_adjoint_csc_i8(cp, rv, nz, x), o.apply_deriv(o.Pr, k, m, x))
            assert np.array_equal(c_oracle.spmv_adjoint_csc_i8(cp, rv, nz, x, omp=True), o.apply_deriv(o.Pr, k, m, x))
            xd = rng.standard_normal(m.nsimplices(k + 1))
            assert np.array_equal(c_oracle.spmv_csc_i8(m.nsimplices(k), cp, rv, nz, xd), o.spmv_csc(m.boundaries[k + 1], xd))
        L = o.laplace(o.Pr, 0, m)
        cp, rv, nz = o.to_julia_csc(L)
        u = o.wave_u0(m.coords)
        assert np.array_equal(c_oracle.spmv_csc_f64(L.shape[0], cp, rv, nz, u), o.spmv_csc(L, u))
        b = o.calc_isboundary(m, 0)
        dt = o.wave_dt(m)
        uu, vv = c_oracle.wave_leapfrog_csc(cp, rv, nz, 1.0 - b, u, np.zeros_like(u), dt, 4)
        ou, ov = o.wave_leapfrog(m, u, np.zeros_like(u), dt, 4)
        assert np.array_equal(uu, ou) and np.array_equal(vv, ov)


# --------------------------------------------------------------------- wave / poisson
def test_wave_operator_and_second_order():
    """examples/wave.jl:82-118 restated: F = blockdiag(E-B) [0 E; L 0]; the Laplacian of the
    eigenmode prod sin(pi x_d) is -D pi^2 u to second order (the 4th-order table in
    wave.jl:172-186 is stale and NOT used as a golden)."""
    errs = []
    for n in (4, 8, 16):
        m = o.large_hypercube_manifold(2, (n, n))
        b, L = o.wave_operators(m)
        u0 = o.wave_u0(m.coords)
        dt = o.wave_dt(m)
        assert dt == 1 / (2 * n)                                             # wave.jl:106,118
        # staggered leapfrog to t = 0.5: v^{-1/2} = v0 - dt/2 (1-b) L u0
        v = -(dt / 2) * ((1.0 - b) * o.spmv_csc(L, u0))
        nsteps = int(round(0.5 / dt))
        u, _ = o.wave_leapfrog(m, u0, v, dt, nsteps)
        exact = math.cos(math.pi * math.sqrt(2) * nsteps * dt) * u0
        errs.append(math.sqrt(np.sum((u - exact) ** 2) / len(u)))            # wave.jl:162
    # the solution error converges at second order (the local truncation error of the DEC
    # Laplacian on this mesh does not converge pointwise -- supraconvergence)
    assert errs[1] < errs[0] / 3 and errs[2] < errs[1] / 3, errs
    m = o.large_hypercube_manifold(2, (4, 4))
    F = o.wave_F(m)
    V = m.nsimplices(0)
    U = np.concatenate([o.wave_u0(m.coords), np.arange(V, dtype=float)])
    du, dv = o.wave_rhs(m, U[:V], U[V:])
    assert np.array_equal(o.spmv_csc(F, U), np.concatenate([du, dv]))


# ------------------------------------------------------------------ N1: Tsit5 tableau
def test_tsit5_tableau_order_conditions():
    """OrdinaryDiffEq is not vendored in the reference; the tableau restated in the oracle is pinned by the
    Runge-Kutta order conditions up to order 5 (17 conditions) and the row-sum conditions."""
    A = np.zeros((7, 7))
    for s_, row in enumerate(o.TSIT5_A):
        A[s_ + 1, :len(row)] = row
    c = np.array((0.0,) + o.TSIT5_C)
    b = A[6].copy()                         # FSAL: the weights are the last row
    A6, c6, b6 = A[:6, :6], c[:6], b[:6]
    assert np.allclose(A.sum(axis=1), c, atol=2e-15)
    e = np.ones(6)
    Ac = A6 @ c6
    conds = [
        (b6 @ e, 1), (b6 @ c6, 1 / 2), (b6 @ c6 ** 2, 1 / 3), (b6 @ Ac, 1 / 6),
        (b6 @ c6 ** 3, 1 / 4), (b6 @ (c6 * Ac), 1 / 8), (b6 @ (A6 @ c6 ** 2), 1 / 12), (b6 @ (A6 @ Ac), 1 / 24),
        (b6 @ c6 ** 4, 1 / 5), (b6 @ (c6 ** 2 * Ac), 1 / 10), (b6 @ (c6 * (A6 @ c6 ** 2)), 1 / 15),
        (b6 @ (c6 * (A6 @ Ac)), 1 / 30), (b6 @ (Ac * Ac), 1 / 20), (b6 @ (A6 @ c6 ** 3), 1 / 20),
        (b6 @ (A6 @ (c6 * Ac)), 1 / 40), (b6 @ (A6 @ (A6 @ c6 ** 2)), 1 / 60), (b6 @ (A6 @ (A6 @ Ac)), 1 / 120),
    ]
    for got, want in conds:
        assert abs(got - want) < 5e-15, (got, want)


def test_tsit5_fifth_order_on_oscillator():
    """u'' = -u integrated to t = 1: the error falls by ~2^5 per halving of dt."""
    def f(U):
        return (U[1], -U[0])
    errs = []
    for n in (8, 16, 32):
        u, v = o.tsit5_steps(f, (np.array([1.0]), np.array([0.0])), 1.0 / n, n)
        errs.append(abs(u[0] - np.cos(1.0)) + abs(v[0] + np.sin(1.0)))
    assert 24 < errs[0] / errs[1] < 40 and 24 < errs[1] / errs[2] < 40, errs


# ------------------------------------------------------------------ hand-derived known answer
def test_two_triangle_known_answer():
    """tests/golden/two_triangles.json (derived by hand from the reference's rules, see make_two_triangles.py;
    the same numbers are compiled into tests/c/abi_client.c): the oracle reproduces edges, signs and d0 x."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "two_triangles.json")))
    m = o.build_manifold("square", np.array(g["triangles"]) - 1, np.array(g["coords"], dtype=float))
    assert [m.nsimplices(R) for R in range(3)] == [4, 5, 2]
    assert (m.simplices[1] + 1).tolist() == g["edges"]
    for R, key in ((1, "boundary1"), (2, "boundary2")):
        cp, rv, nz = o.to_julia_csc(m.boundaries[R])
        for j, col in enumerate(g[key]):
            assert rv[cp[j] - 1:cp[j + 1] - 1].tolist() == col["rows"]
            assert nz[cp[j] - 1:cp[j + 1] - 1].tolist() == col["signs"]
    assert o.apply_deriv(o.Pr, 0, m, np.array(g["x"])).tolist() == g["d0x"]
    assert np.allclose(o.calc_volumes(m, 2), g["area_per_triangle"], atol=1e-15)


def test_regular_simplex_closed_forms():
    """tests/golden/regular_simplex.json: volumes, circumcentric dual volumes, Hodge stars and laplace(Pr,0) of ONE
    regular unit simplex in closed form (derived by hand in make_regular_simplex.py from the reference's definitions,
    not from this oracle) -- an element-wise known answer for star_k and L beyond the reference's identities."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regular_simplex.json")))
    for D in (2, 3):
        m = o.simplex_manifold(D)
        k = g[str(D)]
        for R in range(D + 1):
            assert np.allclose(o.calc_volumes(m, R), k["vol"][R], rtol=1e-14, atol=0), (D, R)
            assert np.allclose(o.calc_dualvolumes(m, R), k["dualvol"][R], rtol=1e-13, atol=0), (D, R)
            assert np.allclose(o.hodge_diag(o.Pr, R, m), k["hodge"][R], rtol=1e-13, atol=0), (D, R)
        L = o.laplace(o.Pr, 0, m).toarray()
        off = ~np.eye(D + 1, dtype=bool)
        assert np.allclose(L[off], k["laplace_offdiag"], rtol=1e-13) and np.allclose(np.diag(L), k["laplace_diag"], rtol=1e-13)


def test_weighted_circumcentre_reference_notebook():
    """src/hypercube-weights.nb (the reference's own Mathematica derivation of the level weights, outputs stored in the
    file; transcribed by tests/golden/make_hypercube_weights.py): weighted circumcentre and squared radius of the Kuhn
    path simplex in D = 1..5 -- with the solved weights the centre is the barycentre (Out[14], [21], [28], [35]), and for
    ARBITRARY weights it is the linear function of Out[12], [19], [26], [33].  Pins circumcentre() (Algorithms.jl:56-96)
    against values the reference holds.  The notebook's `Total[(x - c)^2 - w]` subtracts w from every component, i.e.
    its weights enter with a factor D (see the make script); the Julia code has no such factor, and the last block pins
    what the CODE yields for the generator's table (ManifoldConstructors.jl:253-271), worked out by hand."""
    import json
    from fractions import Fraction as F
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hypercube_weights.json")))

    def path_simplex(D):
        return np.array([[1.0 if d >= D - i else 0.0 for d in range(D)] for i in range(D + 1)])

    def power_radius2(xs, ws, c):
        r2 = ((xs - c) ** 2).sum(axis=1) - ws
        assert np.allclose(r2, r2[0], rtol=0, atol=1e-14)        # the same power distance from every vertex
        return r2[0]
    xs = path_simplex(1)
    c = o.circumcentre(xs[None], np.zeros((1, 2)))[0]
    assert np.allclose(c, 0.5) and abs(power_radius2(xs, np.zeros(2), c) - 0.25) < 1e-15
    rng = np.random.default_rng(3)
    for D in (2, 3, 4, 5):
        xs = path_simplex(D)
        k = g["solved"][str(D)]
        wnb = np.array([0.0] + [float(F(*w)) for w in k["w"]] + [0.0])       # the notebook's weights ...
        c = o.circumcentre(xs[None], (D * wnb)[None])[0]                     # ... enter its equation with a factor D
        want = np.array([float(F(*v)) for v in k["centre"]])
        assert np.allclose(c, want, rtol=0, atol=1e-15 * D), (D, c, want)
        assert np.allclose(want, xs.mean(axis=0))                 # = the barycentre (Out[10], [17], [24], [31])
        assert abs(power_radius2(xs, D * wnb, c) - float(F(*k["cr2"]))) < 1e-14, D
        if D in (2, 3):                                           # the generator's table holds the same numbers
            assert np.array_equal(wnb, np.array({2: [0.0, -1.0 / 6, 0.0], 3: [0.0, -1.0 / 6, -1.0 / 6, 0.0]}[D]))
        # arbitrary weights: the general solution
        M = np.array(g["general"][str(D)]["M"], dtype=float)
        for _ in range(5):
            w = rng.uniform(-0.3, 0.3, D - 1)
            wnb = np.concatenate([[0.0], w, [0.0]])
            c = o.circumcentre(xs[None], (D * wnb)[None])[0]
            assert np.allclose(c, 0.5 * (1 + D * (M @ w)), rtol=0, atol=1e-14), (D, w)
            cr2 = D / 4 * (1 + 2 * D * ((w ** 2).sum() - (w[:-1] * w[1:]).sum()))
            assert abs(power_radius2(xs, D * wnb, c) - cr2) < 1e-13, (D, w)
    # What the Julia CODE yields for the generator's D = 2 table (b_i = x_i.x_i - w_i, no factor D): subtracting the
    # equations pairwise gives 1 - 2 c_2 = w_2 and 1 - 2 c_1 = -w_2, i.e. (5/12, 7/12) -- 1/12 off the barycentre
    c = o.circumcentre(path_simplex(2)[None], np.array([[0.0, -1.0 / 6, 0.0]]))[0]
    assert np.allclose(c, [5.0 / 12, 7.0 / 12], rtol=0, atol=1e-15)


def test_poisson_exact_solution_reference_notebook():
    """examples/poisson.nb (the reference's Mathematica derivation, outputs stored in the file) holds the free-space
    potentials of the Gaussian charge that examples/poisson.jl:93-122 hard-codes: Out[5] (D=1), Out[13] (D=2),
    Out[20] (D=3) and their small-r series Out[6], Out[14].  The restated u_exact equals those closed forms."""
    from scipy.special import erf, expi
    sigma, x0 = 0.2, 0.1
    rng = np.random.default_rng(9)
    for D in (1, 2, 3):
        x = rng.uniform(-0.5, 1.5, (200, D))
        r = np.linalg.norm(x - x0, axis=1)
        if D == 1:      # Out[5]: (-1 + e^(-r^2/2s^2)) s / sqrt(2 pi) + r/2 Erf[r / (sqrt 2 s)]
            want = (-1 + np.exp(-r ** 2 / (2 * sigma ** 2))) * sigma / np.sqrt(2 * np.pi) + 0.5 * r * erf(r / (np.sqrt(2) * sigma))
        elif D == 2:    # Out[13]: -(-EulerGamma + Ei(-r^2/2s^2) + Log 2 - 2 Log r + 2 Log s) / (4 pi)
            want = -(-np.euler_gamma + expi(-r ** 2 / (2 * sigma ** 2)) + np.log(2) - 2 * np.log(r) + 2 * np.log(sigma)) / (4 * np.pi)
        else:           # Out[20]: -Erf[r / (sqrt 2 s)] / (4 pi r)
            want = -erf(r / (np.sqrt(2) * sigma)) / (4 * np.pi * r)
        got = o.poisson_u_exact(x)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-15), D
    # the small-r branch of D = 2 (poisson.jl:103-105) against the series Out[14]: r^2/(8 pi s^2) - r^4/(64 pi s^4) + O(r^6)
    x = np.full((1, 2), x0) + np.array([[3e-3, 0.0]])
    r = 3e-3
    ser = r ** 2 / (8 * np.pi * sigma ** 2) - r ** 4 / (64 * np.pi * sigma ** 4)
    assert abs(o.poisson_u_exact(x)[0] - ser) < 1e-9 * abs(ser) + (r / sigma) ** 6


def test_handwritten_cube_of_24_tetrahedra():
    """test/test-topologies.jl:78-107: the cube cut into 24 tetrahedra around its centre (8 corners, 6 face centres,
    the body centre; vertex names p000 ... p111 are 1..15 in the order written there).  Counts by Euler's formula,
    boundary o boundary = 0 (checkboundary2), every interior triangle shared by two tetrahedra, the boundary the 24
    triangles of the 6 faces."""
    p000, p002, p020, p022, p200, p202, p220, p222, p110, p112, p101, p121, p011, p211, p111 = range(15)
    tets = np.array([(p000, p020, p110, p111), (p020, p220, p110, p111), (p220, p200, p110, p111), (p200, p000, p110, p111),
                     (p002, p022, p112, p111), (p022, p222, p112, p111), (p222, p202, p112, p111), (p202, p002, p112, p111),
                     (p000, p002, p101, p111), (p002, p202, p101, p111), (p202, p200, p101, p111), (p200, p000, p101, p111),
                     (p020, p022, p121, p111), (p022, p222, p121, p111), (p222, p220, p121, p111), (p220, p020, p121, p111),
                     (p000, p002, p011, p111), (p002, p022, p011, p111), (p022, p020, p011, p111), (p020, p000, p011, p111),
                     (p200, p202, p211, p111), (p202, p222, p211, p111), (p222, p220, p211, p111), (p220, p200, p211, p111)])
    xyz = {"0": 0.0, "1": 0.5, "2": 1.0}
    names = ["000", "002", "020", "022", "200", "202", "220", "222", "110", "112", "101", "121", "011", "211", "111"]
    coords = np.array([[xyz[c] for c in nm] for nm in names])
    m = o.build_manifold("cube", tets, coords)
    n = [m.nsimplices(R) for R in range(4)]
    assert n[0] == 15 and n[3] == 24
    assert n[0] - n[1] + n[2] - n[3] == 1                      # Euler characteristic of a ball
    assert n[2] == (4 * 24 + 24) // 2 and n[1] == 50           # 24 boundary triangles; 12 cube edges + 24 face diagonals + 8 + 6
    for R in (2, 3):
        z = m.boundaries[R - 1].astype(np.int64) @ m.boundaries[R].astype(np.int64)
        assert z.nnz == 0 or np.abs(z.data).max() == 0
    per_face = np.asarray(abs(m.boundaries[3]).sum(axis=1)).reshape(-1)
    assert sorted(set(per_face.tolist())) == [1, 2] and int((per_face == 1).sum()) == 24
    assert o.calc_isboundary(m, 2).sum() == 24 and o.calc_isboundary(m, 0).sum() == 14     # all but the body centre
    assert abs(o.calc_volumes(m, 3).sum() - 1.0) < 1e-15 and np.allclose(o.calc_volumes(m, 3), 1.0 / 24)
    # the C twin and its linear-time variant give the same matrices
    import c_oracle
    tab = m.simplices[3]
    for R in (3, 2, 1):
        a = c_oracle.calc_faces_boundaries(tab)
        b = c_oracle.calc_faces_boundaries_linear(tab, 15)
        ref = o.to_julia_csc(m.boundaries[R])
        for x, y, zz in zip(a[1:], b[1:], ref):
            assert np.array_equal(x, y) and np.array_equal(x, zz)
        assert np.array_equal(a[0], m.simplices[R - 1]) and np.array_equal(b[0], m.simplices[R - 1])
        tab = a[0]


def test_linear_assembly_equals_literal_restatement():
    """oracle/ddf_oracle.c: the linear-time face ranking that builds the CPU baseline's n = 256 storage against the
    literal comparison-sort restatement of Manifolds.jl:717-768 and the numpy oracle -- Kuhn, Delaunay (scattered
    numbering, long buckets) and a fan with one very high-valence vertex; and the C Kuhn table against the oracle's."""
    import c_oracle
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(11)
    pts = rng.random((3000, 3))
    k = 3000
    fan = np.stack([np.zeros(k, dtype=np.int64), 1 + np.arange(k), 1 + (np.arange(k) + 1) % k], axis=1)[rng.permutation(k)]
    cases = [(o.large_hypercube_tables(3, (6, 6, 6))[0], 7 ** 3), (o.large_hypercube_tables(2, (10, 10))[0], 121),
             (np.sort(np.asarray(Delaunay(pts).simplices, dtype=np.int64), axis=1), 3000), (np.sort(fan, axis=1), k + 1)]
    for tab, V in cases:
        while tab.shape[1] >= 2:
            a = c_oracle.calc_faces_boundaries(tab)
            b = c_oracle.calc_faces_boundaries_linear(tab, V)
            faces, B = o.calc_faces_boundaries(tab, V)
            for x, y, zz in zip(a[1:], b[1:], o.to_julia_csc(B)):
                assert np.array_equal(x, y) and np.array_equal(x, zz)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[0], faces)
            tab = a[0]
            if tab.shape[1] == 1:
                break
    for n in (2, 6):
        assert np.array_equal(c_oracle.kuhn3_table(n, o.symmetric_hypercube(3)), o.large_hypercube_tables(3, (n,) * 3)[0])

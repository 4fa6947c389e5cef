"""Exact pins of the geometry and of the operators built on it: vol, weighted circumcentres,
circumcentric dual volumes, star_k and L = laplace(Pr,0) from tests/exact_geometry.py -- rational arithmetic + 50-digit
square roots, written from the Julia formulas independently of oracle/ddf_oracle.py -- against the oracle (CPU tier) and
against the device (GPU tier).  The reference's own tests pin the same quantities the same way where they can
(Rational{BigInt} in test/test-algorithms.jl; closed-form volumes in test-manifolds.jl:103-131).

Tolerance 1e-13 relative to the largest value of each array (the meshes are small: the conditioning factor
(extent/h)^2 of the reference's circumcentre system, see tests/test_gpu_round2.py, is <= 64 here)."""
import numpy as np
import pytest

import ddf_oracle as o
import exact_geometry as xg

TOL = 1e-13

MESHES = {
    "kuhn2n2": lambda: o.large_hypercube_manifold(2, (2, 2)),
    "kuhn2n4": lambda: o.large_hypercube_manifold(2, (4, 4)),
    "kuhn3n2": lambda: o.large_hypercube_manifold(3, (2, 2, 2)),
    "kuhn3n4": lambda: o.large_hypercube_manifold(3, (4, 4, 4)),
    "regular3": lambda: o.simplex_manifold(3),
    "orthogonal3": lambda: o.orthogonal_simplex_manifold(3),
    "refined2l2": lambda: o.refined_simplex_manifold(2, 2),
}
_EXACT = {}


def exact(kind):
    if kind not in _EXACT:
        om = MESHES[kind]()
        simp = {R: [tuple(int(v) for v in row) for row in om.simplices[R]] for R in range(om.D + 1)}
        vol, dualvol, cc = xg.geometry(om.D, simp, [tuple(map(float, p)) for p in om.coords], [float(w) for w in om.weights])
        _EXACT[kind] = (om, simp, vol, dualvol, cc)
    return _EXACT[kind]


def f64(v):
    return np.array([float(x) for x in v])


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def check(kind, get_vol, get_dualvol, get_cc, apply_L):
    om, simp, vol, dualvol, cc = exact(kind)
    D = om.D
    for R in range(D + 1):
        assert rel(get_vol(R), f64(vol[R])) <= TOL, f"vol R={R}"
        assert rel(get_dualvol(R), f64(dualvol[R])) <= TOL, f"dualvol R={R}"
        want = np.array([[float(c) for c in p] for p in cc[R]])
        ext = float(np.max(np.abs(om.coords))) or 1.0
        assert np.max(np.abs(get_cc(R) - want)) <= TOL * ext, f"circumcentre R={R}"
        star = f64([dv / v for dv, v in zip(dualvol[R], vol[R])])
        assert rel(get_dualvol(R) / get_vol(R), star) <= 2 * TOL, f"star R={R}"
    if apply_L is not None and kind not in ("regular3", "orthogonal3"):
        L = xg.laplace0(D, simp, vol, dualvol)
        V = om.nsimplices(0)
        A = np.zeros((V, V))
        for (i, j), l in L.items():
            A[i, j] = float(l)
        for col in (0, V // 2, V - 1):                      # columns of the assembled operator = L applied to unit vectors
            e = np.zeros(V)
            e[col] = 1.0
            assert np.max(np.abs(apply_L(e) - A[:, col])) <= 4 * TOL * np.max(np.abs(A)), f"L column {col}"


@pytest.mark.parametrize("kind", sorted(MESHES))
def test_oracle_geometry_and_laplacian_exact(kind):
    om = exact(kind)[0]
    L = o.laplace(o.Pr, 0, om) if kind not in ("regular3", "orthogonal3") else None
    check(kind, lambda R: o.calc_volumes(om, R), lambda R: o.calc_dualvolumes(om, R), lambda R: o.calc_dualcoords(om, R),
          (lambda x: o.spmv_csc(L, x)) if L is not None else None)


def test_kuhn_circumcentres_are_the_rationals_of_the_level_weights():
    """Rational pin in the reference's own style (Rational{BigInt}, test/test-algorithms.jl): with the level weights of
    ManifoldConstructors.jl:253-263 taken as the exact rationals -1/6 h^2 (levels 1 .. D-1) every weighted circumcentre
    of the Kuhn mesh is a small rational multiple of h, and the oracle's Float64 LU solve reproduces it to the last
    bits (D = 2, 3; n = 2)."""
    from fractions import Fraction as F
    for D in (2, 3):
        n = 2
        om = o.large_hypercube_manifold(D, (n,) * D)
        grid = np.rint(om.coords * n).astype(int)
        lw = [F(0)] + [F(-1, 6)] * (D - 1) + [F(0)]
        W = [lw[int((g % 2).sum())] / (n * n) for g in grid]
        assert np.allclose([float(w) for w in W], om.weights, rtol=1e-15, atol=0)
        X = [[F(int(c), n) for c in g] for g in grid]
        for R in range(1, D + 1):
            got = o.calc_dualcoords(om, R)
            for i, s in enumerate(om.simplices[R]):
                q = xg.circumcentre([X[v] for v in s], [W[v] for v in s])
                assert all(c.denominator <= 48 * n for c in q), (D, R, i, q)
                assert np.max(np.abs(got[i] - np.array([float(c) for c in q]))) <= 2e-15, (D, R, i, q)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", sorted(MESHES))
def test_device_geometry_and_laplacian_exact(ddf, ctx, kind):
    om = exact(kind)[0]
    dm = ddf.Manifold.from_simplices(kind, om.simplices[om.D], om.coords, om.weights, ctx=ctx)
    L = ddf.laplace(ddf.Pr, 0, dm) if kind not in ("regular3", "orthogonal3") else None
    check(kind, dm.get_volumes, dm.get_dualvolumes, lambda R: dm.get_dualcoords(R) if R > 0 else dm.get_coords(),
          (lambda x: (L * ddf.Fun(dm, ddf.Pr, 0, x)).values) if L is not None else None)
    del L
    dm.close()

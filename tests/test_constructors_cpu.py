"""CPU tier: the host-side mesh-input constructors of ddf.jl_b200/manifoldops.py against the oracle's restatements of
src/ManifoldConstructors.jl (the fixtures the reference's own tests run on: hypercube_manifold for test-funs.jl /
test-ops.jl, orthogonal_simplex_manifold and boundary_manifold for test-manifolds.jl).  Tables only -- no device."""
from importlib import import_module

import numpy as np
import pytest

import ddf_oracle as o

mo = import_module("ddf.jl_b200.manifoldops")


@pytest.mark.parametrize("D", range(6))
def test_hypercube_and_orthogonal_simplex_tables(D):
    simp, coords, weights = mo.hypercube_tables(D)
    om = o.hypercube_manifold(D)
    assert simp.shape == (int(np.prod(range(1, D + 1))), D + 1) and coords.shape == (2 ** D, D)    # D! simplices, 2^D corners
    if D > 0:
        assert np.array_equal(np.sort(simp, axis=1), om.simplices[D])        # same simplices in the same order
    assert np.array_equal(coords, om.coords) and np.array_equal(weights, om.weights)
    simp, coords, weights = mo.orthogonal_simplex_tables(D)
    om = o.orthogonal_simplex_manifold(D)
    assert np.array_equal(coords, om.coords) and np.array_equal(weights, om.weights)
    assert simp.tolist() == [list(range(D + 1))]


@pytest.mark.parametrize("D", [1, 2, 3, 4])
def test_boundary_tables_match_the_oracle(D):
    """boundary_simplex_manifold (test/test-manifolds.jl:277-299): the boundary of the regular D-simplex."""
    om = o.simplex_manifold(D)
    rows = np.asarray(om.boundaries[D].astype(np.int64).sum(axis=1)).reshape(-1)
    faces = om.simplices[D - 1] if D - 1 >= 1 else np.arange(om.nsimplices(0))[:, None]
    simp, coords, weights = mo.boundary_tables(rows, faces, om.coords, om.weights)
    ob = o.boundary_manifold(om)
    if D - 1 >= 1:
        assert np.array_equal(simp, ob.simplices[D - 1])
    assert np.array_equal(coords, ob.coords) and np.array_equal(weights, ob.weights)
    assert simp.shape == (D + 1, D)                      # D + 1 faces of D vertices each
    # the boundary of a manifold with boundary is closed: every (D-2)-face of it has exactly two parents
    if D - 1 >= 1:
        parents = np.asarray(abs(ob.boundaries[D - 1]).sum(axis=1)).reshape(-1)
        assert (parents == 2).all()


def test_boundary_of_an_unoriented_mesh_is_rejected_like_the_reference():
    """The reference asserts |B * ones| <= 1 (ManifoldConstructors.jl:556): simplices are stored with ascending vertices,
    so on a Kuhn cube the two parents of an interior face can carry the SAME sign and the row sums to +-2 -- the
    reference stops there, and so does the host mirror (DDFError)."""
    om = o.hypercube_manifold(2)
    rows = np.asarray(om.boundaries[2].astype(np.int64).sum(axis=1)).reshape(-1)
    assert np.abs(rows).max() == 2
    with pytest.raises(mo.DDFError):
        mo.boundary_tables(rows, om.simplices[1], om.coords, om.weights)


def _staggered_grid_literal(D, n):
    """ManifoldConstructors.jl:436-451 transcribed point by point (0-based CartesianIndex from 0 to nelts, first axis
    fastest; the loops over d and d' update x in place in that order)."""
    import math
    pts = []
    for flat in range((n + 1) ** D):
        i = [(flat // (n + 1) ** d) % (n + 1) for d in range(D)]
        x = [i[d] / n for d in range(D)]
        for d in range(1, D + 1):
            for dp in range(1, d):
                if i[dp - 1] in (0, n):
                    continue
                di = 0.5 / math.sqrt(d - 1)
                x1 = x[dp - 1] * n / (n + di)
                if i[d - 1] % 2 == 1:
                    x1 = x1 + di / (n + di)
                x[dp - 1] = x1
        pts.append(x)
    return np.array(pts).reshape(-1, D)


@pytest.mark.parametrize("D,n", [(1, 8), (2, 8), (3, 4), (4, 2)])
def test_delaunay_hypercube_tables(D, n):
    """large_delaunay_hypercube_manifold: the vectorised coordinates equal the literal loop, boundary points stay on
    the cube, and the Delaunay triangulation (scipy, the reference's backend) fills the unit cube (test-manifolds.jl
    checks sum(volumes) on every fixture).  delaunay_hypercube_manifold: the 2^D corners shifted by 5/2."""
    coords, weights = mo.large_delaunay_hypercube_tables(D, n)
    assert np.array_equal(coords, _staggered_grid_literal(D, n)) and not weights.any()
    assert coords.min() == 0.0 and coords.max() == 1.0
    m = o.build_manifold("ld", mo.delaunay_mesh(coords), coords, weights)
    assert abs(o.calc_volumes(m, D).sum() - 1.0) <= 1e-12
    c2, w2 = mo.delaunay_hypercube_tables(D)
    assert c2.shape == (2 ** D, D) and c2.min() == 2.5 and c2.max() == 3.5 and not w2.any()
    m2 = o.build_manifold("dh", mo.delaunay_mesh(c2), c2, w2)
    assert abs(o.calc_volumes(m2, D).sum() - 1.0) <= 1e-12
    with pytest.raises(mo.DDFError):
        mo.large_delaunay_hypercube_tables(D, 3)

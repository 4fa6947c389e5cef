"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on
the same inputs.  Integer / index work must be bit-exact; Float64 operator
arithmetic is bit-exact given identical geometry and within 1e-12 relative end to
end (north_star's tolerance; the geometry kernels evaluate the third-party
DifferentialForms / StaticArrays algebra in a different order).

Mirrors the reference's test strategy (test/test-manifolds.jl, test-manifoldops.jl,
test-ops.jl): manifold zoo x algebraic identities, plus element-wise comparison
with the oracle that the reference cannot offer.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import ddf_oracle as o

pytestmark = pytest.mark.gpu

RTOL = 1e-12   # north_star: "Float64 results must match ... within 1e-12 relative"


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0


def make_pair(ddf, ctx, kind):
    """(oracle manifold, device manifold) built from the same input."""
    if kind.startswith("kuhn") and kind != "kuhn3slab":
        D, n = int(kind[4]), int(kind.split("n")[-1])
        om = o.large_hypercube_manifold(D, (n,) * D)
        dm = ddf.large_hypercube_manifold(D, (n,) * D, ctx=ctx)      # on-device generator
        return om, dm
    if kind.startswith("refined"):
        D, lv = int(kind[7]), int(kind.split("l")[-1])
        om = o.refined_simplex_manifold(D, lv)
        dm = ddf.Manifold.from_simplices(om.name, om.simplices[om.D], om.coords, om.weights, ctx=ctx)
        return om, dm
    if kind.startswith("simplex"):
        D = int(kind[7])
        om = o.simplex_manifold(D)
        dm = ddf.Manifold.from_simplices(om.name, om.simplices[om.D], om.coords, om.weights, ctx=ctx)
        return om, dm
    if kind == "kuhn3slab":
        om = o.large_hypercube_manifold(3, (4, 4, 6), h_div=4)
        dm = ddf.large_hypercube_manifold(3, (4, 4, 6), ctx=ctx, hdiv=4)
        return om, dm
    raise KeyError(kind)


ZOO = ["kuhn1n4", "kuhn2n4", "kuhn2n16", "kuhn3n2", "kuhn3n4", "kuhn3n8", "refined2l1", "refined2l4", "refined3l1",
       "simplex1", "simplex2", "simplex3", "kuhn3slab"]
GEO_ZOO = ["kuhn1n4", "kuhn2n4", "kuhn2n16", "kuhn3n4", "kuhn3n8", "refined2l4", "simplex2", "simplex3", "kuhn3slab"]


@pytest.fixture(scope="module")
def pairs(ddf, ctx):
    cache = {}

    def get(kind):
        if kind not in cache:
            cache[kind] = make_pair(ddf, ctx, kind)
        return cache[kind]
    return get


# ------------------------------------------------------------------ A1/A2 assembly
@pytest.mark.parametrize("kind", ZOO)
def test_assembly_bit_exact(pairs, kind):
    om, dm = pairs(kind)
    for R in range(om.D + 1):
        assert dm.nsimplices(R) == om.nsimplices(R)
        colptr, rowval = dm.get_simplices_csc(R)
        ocp, orv = o.simplices_julia_csc(om.simplices[R])
        assert np.array_equal(colptr, ocp) and np.array_equal(rowval, orv), f"simplices R={R}"
    for R in range(1, om.D + 1):
        colptr, rowval, nzval = dm.get_boundaries_csc(R)
        ocp, orv, onz = o.to_julia_csc(om.boundaries[R])
        assert np.array_equal(colptr, ocp), f"boundary colptr R={R}"
        assert np.array_equal(rowval, orv), f"boundary rowval R={R}"
        assert np.array_equal(nzval, onz) and nzval.dtype == np.int8, f"boundary nzval R={R}"


def test_generator_matches_oracle_tables(ddf, ctx):
    for D, n in ((2, 6), (3, 4)):
        simp, coords, weights = o.large_hypercube_tables(D, (n,) * D)
        dm = ddf.large_hypercube_manifold(D, (n,) * D, ctx=ctx)
        assert np.array_equal(dm.get_simplices(D), simp)          # same simplex ORDER, not just the same set
        assert np.array_equal(dm.get_coords(), coords)
        assert np.array_equal(dm.get_weights(), weights)


def test_uploaded_equals_generated(ddf, ctx):
    simp, coords, weights = o.large_hypercube_tables(3, (4, 4, 4))
    a = ddf.Manifold.from_simplices("up", simp, coords, weights, ctx=ctx)
    b = ddf.large_hypercube_manifold(3, (4, 4, 4), ctx=ctx)
    for R in range(1, 4):
        for x, y in zip(a.get_boundaries_csc(R), b.get_boundaries_csc(R)):
            assert np.array_equal(x, y)


def test_dd_zero_on_device(ddf, ctx, pairs):
    """d_{k+1} d_k x == 0 exactly for integer-valued x (test-manifoldops.jl:60-68)."""
    for kind in ("kuhn3n4", "refined2l4", "simplex3"):
        om, dm = pairs(kind)
        rng = np.random.default_rng(0)
        for k in range(om.D - 1):
            x = ddf.Fun(dm, ddf.Pr, k, rng.integers(-50, 50, om.nsimplices(k)).astype(np.float64))
            y = ddf.deriv(ddf.Pr, k + 1, dm) * (ddf.deriv(ddf.Pr, k, dm) * x)
            assert not np.any(y.values)
        for k in range(om.D, 1, -1):
            x = ddf.Fun(dm, ddf.Pr, k, rng.integers(-50, 50, om.nsimplices(k)).astype(np.float64))
            y = ddf.boundary(ddf.Pr, k - 1, dm) * (ddf.boundary(ddf.Pr, k, dm) * x)
            assert not np.any(y.values)


# --------------------------------------------------------------------- A3/A4 apply
@pytest.mark.parametrize("kind", ZOO)
def test_deriv_apply_bit_exact(ddf, pairs, kind):
    om, dm = pairs(kind)
    rng = np.random.default_rng(1)
    for k in range(om.D):
        x = rng.standard_normal(om.nsimplices(k))
        y = (ddf.deriv(ddf.Pr, k, dm) * ddf.Fun(dm, ddf.Pr, k, x)).values
        assert np.array_equal(y, o.apply_deriv(o.Pr, k, om, x)), f"primal d_{k}"
    for k in range(1, om.D + 1):
        x = rng.standard_normal(om.nsimplices(k))
        y = (ddf.deriv(ddf.Dl, k, dm) * ddf.Fun(dm, ddf.Dl, k, x)).values
        assert np.array_equal(y, o.apply_deriv(o.Dl, k, om, x)), f"dual d_{k}"
        yb = (ddf.boundary(ddf.Pr, k, dm) * ddf.Fun(dm, ddf.Pr, k, x)).values
        assert np.array_equal(yb, o.spmv_csc(om.boundaries[k], x))
    for k in range(om.D):
        x = rng.standard_normal(om.nsimplices(k))
        yb = (ddf.boundary(ddf.Dl, k, dm) * ddf.Fun(dm, ddf.Dl, k, x)).values
        assert np.array_equal(yb, o.spmv_adjoint_csc(om.boundaries[k + 1], x))


def test_deriv_apply_staged_pipeline(ddf, ctx):
    """y = d_k x through the opt-in staged fixed-width TMA pipeline (csrc/fstage.cuh, tuning 100+10*NG+RPT; tables of
    >= 4096 rows) against the oracle and against the default explicit-index kernel -- bit for bit, including a
    ragged last window."""
    n = 20
    om = o.large_hypercube_manifold(3, (n, n, n))
    dm = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
    rng = np.random.default_rng(23)
    for k in range(3):
        x = rng.standard_normal(om.nsimplices(k))
        fx = ddf.Fun(dm, ddf.Pr, k, x)
        y = (ddf.deriv(ddf.Pr, k, dm) * fx).values
        assert np.array_equal(y, o.apply_deriv(o.Pr, k, om, x)), k
        for var in (114, 122, 111, 204, 202, 201):
            ddf._lib.call("ddf_set_tuning", b"fixed", var)
            try:
                y2 = (ddf.deriv(ddf.Pr, k, dm) * fx).values
            finally:
                ddf._lib.call("ddf_set_tuning", b"fixed", 0)
            assert np.array_equal(y, y2), (k, var)
    # a mesh without column locality: the staged build must refuse it and the explicit-index kernel takes over
    ddf._lib.call("ddf_set_tuning", b"fixed", 114)
    simp, coords, weights = o.large_hypercube_tables(3, (12, 12, 12))
    perm = rng.permutation(coords.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    om2 = o.build_manifold("shuffled", inv[simp], coords[perm], weights[perm])
    dm2 = ddf.Manifold.from_simplices("shuffled", inv[simp], coords[perm], weights[perm], ctx=ctx)
    for k in range(3):
        x = rng.standard_normal(om2.nsimplices(k))
        y = (ddf.deriv(ddf.Pr, k, dm2) * ddf.Fun(dm2, ddf.Pr, k, x)).values
        assert np.array_equal(y, o.apply_deriv(o.Pr, k, om2, x)), k
    ddf._lib.call("ddf_set_tuning", b"fixed", 0)


def test_prefix_packed_rows_match_explicit_table(ddf, ctx):
    """The default storage of d_k below the top level: the first face of every row is its prefix and prefix ids run
    non-decreasing down the rows (lexicographic numbering, Manifolds.jl:735-757), so a row keeps W-1 words.  Bit for
    bit the explicit-table kernel and the oracle, plain and with fused Hodge stars on either side; on ordered,
    shuffled, random-Delaunay and ragged meshes; the top level (rows in the caller's order) and meshes whose offsets
    do not fit (forced here with the 'packbits' knob) keep the explicit table."""
    rng = np.random.default_rng(77)
    meshes = {}
    for name, (D, n) in {"kuhn3": (3, 10), "kuhn2": (2, 37 * 2), "kuhn3odd": (3, 6)}.items():
        meshes[name] = o.large_hypercube_tables(D, (n,) * D)
    simp, coords, weights = o.large_hypercube_tables(3, (8, 8, 8))
    perm = rng.permutation(coords.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    meshes["shuffled"] = (inv[simp], coords[perm], weights[perm])
    pts = rng.random((400, 2))
    meshes["delaunay"] = (ddf.delaunay_mesh(pts), pts, np.zeros(len(pts)))
    for name, (s_, c_, w_) in meshes.items():
        om = o.build_manifold(name, s_, c_, w_)
        D = om.D
        for bits in (24, 3):
            ddf._lib.call("ddf_set_tuning", b"packbits", bits)
            try:
                dm = ddf.Manifold.from_simplices(name, s_, c_, w_, ctx=ctx)
                omg = with_device_geometry(om, dm)
                for k in range(D):
                    fmt = ddf.deriv_format(dm, k)
                    ddf._lib.call("ddf_set_tuning", b"fixed", 20)          # no bit-packed rows: the prefix-packed storage
                    try:
                        fmt20 = ddf.deriv_format(dm, k)
                    finally:
                        ddf._lib.call("ddf_set_tuning", b"fixed", 0)
                    if k == D - 1 or bits == 3:
                        # top level: rows in input order (only a mesh given in lexicographic order would prefix-pack)
                        assert fmt20 == "explicit" or (k == D - 1 and bits == 24), (name, k, fmt20)
                        assert fmt == "explicit" or bits == 24, (name, k, fmt)
                    else:
                        assert fmt20 == "packed", (name, k, bits)
                        assert fmt == ("packed" if k == 0 else "bits32"), (name, k, fmt)
                    x = rng.standard_normal(om.nsimplices(k))
                    fx = ddf.Fun(dm, ddf.Pr, k, x)
                    d = ddf.deriv(ddf.Pr, k, dm)
                    y = (d * fx).values
                    assert np.array_equal(y, o.apply_deriv(o.Pr, k, om, x)), (name, k, bits)
                    # fused stars: (star_{k+1} d_k) x and (d_k star_k^-1-typed diagonal) x against the unfused sequence
                    h1 = ddf.hodge(ddf.Pr, k + 1, dm)
                    y1 = ((h1 * d) * fx).values
                    assert np.array_equal(y1, (h1 * (d * fx)).values), (name, k, bits)
                    g = ddf.op_from_diag(dm, ddf.Pr, k, ddf.Pr, k, rng.standard_normal(om.nsimplices(k)))
                    y2 = ((d * g) * fx).values
                    assert np.array_equal(y2, (d * (g * fx)).values), (name, k, bits)
                    ddf._lib.call("ddf_set_tuning", b"fixed", 18)          # explicit table, same mapping
                    try:
                        assert np.array_equal((d * fx).values, y), (name, k, bits)
                        assert np.array_equal(((h1 * d) * fx).values, y1), (name, k, bits)
                    finally:
                        ddf._lib.call("ddf_set_tuning", b"fixed", 0)
                    # 19: packed d0 with lane = consecutive rows; 20: prefix-packed instead of bit-packed rows;
                    # 21, 22: other rows-per-lane counts of the bit-packed kernel
                    for var in (19, 20, 21, 22):
                        ddf._lib.call("ddf_set_tuning", b"fixed", var)
                        try:
                            assert np.array_equal((d * fx).values, y), (name, k, bits, var)
                            if var == 20:
                                assert np.array_equal(((h1 * d) * fx).values, y1), (name, k, bits)
                                assert np.array_equal(((d * g) * fx).values, y2), (name, k, bits)
                        finally:
                            ddf._lib.call("ddf_set_tuning", b"fixed", 0)
                del omg
                dm.close()
            finally:
                ddf._lib.call("ddf_set_tuning", b"packbits", 24)


def test_bit_packed_rows_match_explicit_table(ddf, ctx):
    """Default storage of d_1, d_2: ONE word of index differences per row (index 0 against the smallest index 0 of its
    32-row group, later indices against their predecessors), 4 or 8 bytes as the table's largest differences need.
    Every (row width, word size) combination, bit for bit the explicit-table kernel and the oracle, plain and with
    fused Hodge stars on either side: Kuhn cubes (d1: 4-byte words; d2 at the top level: 4-byte words on a small cube,
    8-byte words on a larger one), a vertex-shuffled cube (8-byte words for d2), a large random Delaunay triangulation
    in Qhull's triangle order (d1 at the top level: 8-byte words), a ragged row count (not a multiple of 32 * RPW)."""
    rng = np.random.default_rng(78)
    meshes = {}
    meshes["kuhn3n6"] = o.large_hypercube_tables(3, (6, 6, 6)) + ({1: "bits32", 2: "bits32"},)
    meshes["kuhn3n24"] = o.large_hypercube_tables(3, (24, 24, 24)) + ({1: "bits32", 2: "bits64"},)
    simp, coords, weights = o.large_hypercube_tables(3, (8, 8, 8))
    perm = rng.permutation(coords.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    meshes["shuffled"] = (inv[simp], coords[perm], weights[perm], {1: "bits32", 2: "bits64"})
    pts = rng.random((30000, 2))
    meshes["delaunay2"] = (ddf.delaunay_mesh(pts), pts, np.zeros(len(pts)), {1: "bits64"})
    seen = set()
    for name, (s_, c_, w_, expect) in meshes.items():
        om = o.build_manifold(name, s_, c_, w_)
        dm = ddf.Manifold.from_simplices(name, s_, c_, w_, ctx=ctx)
        for k in range(1, om.D):
            fmt = ddf.deriv_format(dm, k)
            assert fmt == expect[k], (name, k, fmt)
            seen.add((k + 2, fmt))
            x = rng.standard_normal(om.nsimplices(k))
            fx = ddf.Fun(dm, ddf.Pr, k, x)
            d = ddf.deriv(ddf.Pr, k, dm)
            y = (d * fx).values
            assert np.array_equal(y, o.apply_deriv(o.Pr, k, om, x)), (name, k)
            h1 = ddf.hodge(ddf.Pr, k + 1, dm)
            g = ddf.op_from_diag(dm, ddf.Pr, k, ddf.Pr, k, rng.standard_normal(om.nsimplices(k)))
            y1 = ((h1 * d) * fx).values
            y2 = ((d * g) * fx).values
            y3 = ((h1 * d * g) * fx).values
            assert np.array_equal(y1, (h1 * (d * fx)).values), (name, k)
            assert np.array_equal(y2, (d * (g * fx)).values), (name, k)
            assert np.array_equal(y3, (h1 * (d * (g * fx))).values), (name, k)
            for var in (18, 20, 21, 22):                     # explicit table; no bit-packed rows; other rows per lane
                ddf._lib.call("ddf_set_tuning", b"fixed", var)
                try:
                    assert np.array_equal((d * fx).values, y), (name, k, var)
                    assert np.array_equal(((h1 * d) * fx).values, y1), (name, k, var)
                    assert np.array_equal(((d * g) * fx).values, y2), (name, k, var)
                finally:
                    ddf._lib.call("ddf_set_tuning", b"fixed", 0)
        dm.close()
    assert seen == {(3, "bits32"), (3, "bits64"), (4, "bits32"), (4, "bits64")}, seen


def test_boundary_upper_runs_match_plain_rows(ddf, ctx):
    """Rows of the boundary matrix d_R (R < D): the cofaces that extend the row's simplex at the end are consecutive
    ids, come last, and tile the R-simplices in row order (lexicographic numbering), so they are stored as run offsets
    and -- one contiguous piece of x per 256-row window -- fetched by a single TMA bulk copy.  Bit for bit the plain
    SELL copy, the row-CSR and the oracle, with and without fused stars, on ordered, shuffled and Delaunay meshes."""
    rng = np.random.default_rng(91)
    meshes = {"kuhn3": o.large_hypercube_tables(3, (10, 10, 10)), "kuhn3odd": o.large_hypercube_tables(3, (6, 6, 6)),
              "kuhn2": o.large_hypercube_tables(2, (40, 40))}
    simp, coords, weights = o.large_hypercube_tables(3, (8, 8, 8))
    perm = rng.permutation(coords.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    meshes["shuffled"] = (inv[simp], coords[perm], weights[perm])
    pts = rng.random((300, 3))
    meshes["delaunay3"] = (ddf.delaunay_mesh(pts), pts, np.zeros(len(pts)))
    for name, (s_, c_, w_) in meshes.items():
        om = o.build_manifold(name, s_, c_, w_)
        D = om.D
        ref = {}
        for tune in (20, 22, 21, 0, 9):       # plain SELL; upper runs staged (forced for short rows) / global loop; default; CSR
            ddf._lib.call("ddf_set_tuning", b"fixed", tune)
            try:
                dm = ddf.Manifold.from_simplices(name, s_, c_, w_, ctx=ctx)
                rng2 = np.random.default_rng(5)
                for R in range(1, D + 1):
                    x = rng2.standard_normal(om.nsimplices(R))
                    fx = ddf.Fun(dm, ddf.Dl, R, x)
                    A = ddf.deriv(ddf.Dl, R, dm)                 # the matrix d_R as stored (ManifoldOps.jl:46)
                    y = (A * fx).values
                    g1 = ddf.op_from_diag(dm, ddf.Dl, R - 1, ddf.Dl, R - 1, rng2.standard_normal(om.nsimplices(R - 1)))
                    g2 = ddf.op_from_diag(dm, ddf.Dl, R, ddf.Dl, R, rng2.standard_normal(om.nsimplices(R)))
                    y2 = (((g1 * A) * g2) * fx).values           # fused diagonals on both sides
                    if tune == 20:
                        assert np.array_equal(y, o.apply_deriv(o.Dl, R, om, x)), (name, R)
                        assert np.array_equal(y2, (g1 * (A * (g2 * fx))).values), (name, R)
                        ref[R] = (y, y2)
                    else:
                        assert np.array_equal(y, ref[R][0]) and np.array_equal(y2, ref[R][1]), (name, R, tune)
                dm.close()
            finally:
                ddf._lib.call("ddf_set_tuning", b"fixed", 0)


def test_deriv_apply_ragged_sizes(ddf, ctx):
    """rows % 4 != 0 exercises the scalar tail of the 4-rows-per-thread kernel."""
    for n in (2, 6, 10):
        om = o.large_hypercube_manifold(2, (n, n))
        dm = ddf.large_hypercube_manifold(2, (n, n), ctx=ctx)
        rng = np.random.default_rng(n)
        for k in range(2):
            x = rng.standard_normal(om.nsimplices(k))
            y = (ddf.deriv(ddf.Pr, k, dm) * ddf.Fun(dm, ddf.Pr, k, x)).values
            assert np.array_equal(y, o.apply_deriv(o.Pr, k, om, x))


# ------------------------------------------------------------------- A5-A8 geometry
@pytest.mark.parametrize("kind", GEO_ZOO)
def test_geometry_parity(pairs, kind):
    om, dm = pairs(kind)
    for R in range(om.D + 1):
        assert relerr(dm.get_volumes(R), o.calc_volumes(om, R)) <= RTOL, f"vol R={R}"
        # positions: error relative to the mesh extent (a circumcentre may sit at the origin)
        ext = float(np.max(np.abs(om.coords))) or 1.0
        assert np.max(np.abs(dm.get_dualcoords(R) - o.calc_dualcoords(om, R))) <= 1e-12 * ext, f"dualcoords R={R}"
        dv, odv = dm.get_dualvolumes(R), o.calc_dualvolumes(om, R)
        assert relerr(dv, odv) <= RTOL, f"dualvol R={R}"
    assert np.all(dm.get_volumes(0) == 1) and np.all(dm.get_dualvolumes(om.D) == 1)


def test_weighted_circumcentre_reference_notebook(ddf, ctx):
    """tests/golden/hypercube_weights.json = the OUTPUT cells of the reference's src/hypercube-weights.nb: with the
    notebook's weights (which enter its equation with a factor D, see tests/golden/make_hypercube_weights.py) the
    weighted circumcentre of the Kuhn path simplex is its barycentre.  The device kernel on the same input, D = 2, 3."""
    import json
    from fractions import Fraction as F
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hypercube_weights.json")))
    for D in (2, 3):
        xs = np.array([[1.0 if d >= D - i else 0.0 for d in range(D)] for i in range(D + 1)])
        k = g["solved"][str(D)]
        w = D * np.array([0.0] + [float(F(*v)) for v in k["w"]] + [0.0])
        dm = ddf.Manifold.from_simplices("path simplex", np.arange(D + 1, dtype=np.int64)[None, :], xs, w, ctx=ctx)
        cc = dm.get_dualcoords(D)
        want = np.array([float(F(*v)) for v in k["centre"]])
        assert np.max(np.abs(cc.reshape(-1) - want)) <= RTOL, (D, cc, want)
        dm.close()


@pytest.mark.parametrize("kind", GEO_ZOO)
def test_isboundary_bit_exact(pairs, kind):
    om, dm = pairs(kind)
    for R in range(om.D):
        assert np.array_equal(dm.get_isboundary(R), o.calc_isboundary(om, R)), f"isboundary R={R}"


def with_device_geometry(om, dm):
    """Give the oracle the device's vol / dualvol so that the OPERATOR arithmetic
    can be compared bit for bit (geometry itself is compared in test_geometry_parity)."""
    for R in range(om.D + 1):
        om._cache[("vol", R)] = dm.get_volumes(R)
        om._cache[("dualvol", R, True)] = dm.get_dualvolumes(R)
    return om


@pytest.mark.parametrize("kind", ["kuhn2n4", "kuhn3n4", "refined2l4", "simplex3"])
def test_hodge_apply_bit_exact(ddf, pairs, kind):
    om, dm = pairs(kind)
    om = with_device_geometry(om, dm)
    rng = np.random.default_rng(2)
    for R in range(om.D + 1):
        x = rng.standard_normal(om.nsimplices(R))
        for P in (o.Pr, o.Dl):
            y = (ddf.hodge(P, R, dm) * ddf.Fun(dm, P, R, x)).values
            assert np.array_equal(y, o.hodge_diag(P, R, om) * x)
            y = (ddf.invhodge(P, R, dm) * ddf.Fun(dm, P, R, x)).values
            assert np.array_equal(y, o.invhodge_diag(P, R, om) * x)


# ---------------------------------------------------------- A9 coderiv / laplace
@pytest.mark.parametrize("kind", ["kuhn2n4", "kuhn3n4", "refined2l4", "simplex3", "kuhn3slab"])
def test_coderiv_assembled_bit_exact(ddf, pairs, kind):
    om, dm = pairs(kind)
    om = with_device_geometry(om, dm)
    for P in (o.Pr, o.Dl):
        dR = 1 if P else -1
        for R in range(om.D + 1):
            if not (0 <= R - dR <= om.D):
                continue
            A = ddf.coderiv(P, R, dm).to_csc()
            B = o.coderiv(P, R, om)
            assert A.shape == B.shape
            assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices), (P, R)
            assert np.array_equal(A.data, B.data), (P, R)


@pytest.mark.parametrize("kind", ["kuhn2n4", "kuhn2n16", "kuhn3n4", "refined2l4", "simplex3", "kuhn3slab"])
def test_laplace0_bit_exact(ddf, pairs, kind):
    om, dm = pairs(kind)
    om = with_device_geometry(om, dm)
    L = o.laplace(o.Pr, 0, om)
    A = ddf.laplace(ddf.Pr, 0, dm).to_csc()
    assert np.array_equal(A.indptr, L.indptr) and np.array_equal(A.indices, L.indices)
    assert np.array_equal(A.data, L.data)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(om.nsimplices(0))
    y = (ddf.laplace(ddf.Pr, 0, dm) * ddf.Fun(dm, ddf.Pr, 0, x)).values
    assert np.array_equal(y, o.spmv_csc(L, x))


@pytest.mark.parametrize("kind", ["kuhn2n4", "kuhn3n4", "refined2l4"])
def test_laplace_end_to_end_tolerance(ddf, pairs, kind):
    """Independent geometry on both sides: 1e-12-class agreement (north_star)."""
    om, dm = pairs(kind)
    om._cache.clear()
    L = o.laplace(o.Pr, 0, om)
    u = o.wave_u0(om.coords) if kind.startswith("kuhn") else o.readme_field(om.coords)
    y = (ddf.laplace(ddf.Pr, 0, dm) * ddf.Fun(dm, ddf.Pr, 0, u)).values
    assert relerr(y, o.spmv_csc(L, u)) <= RTOL


@pytest.mark.parametrize("kind", ["kuhn2n4", "kuhn3n4", "simplex3"])
def test_laplace_general_R(ddf, pairs, kind):
    """laplace(P,R) for R > 0 (both terms), lazily composed on the device, against the
    oracle's assembled operator (association differs: tolerance)."""
    om, dm = pairs(kind)
    om = with_device_geometry(om, dm)
    rng = np.random.default_rng(4)
    for P in (o.Pr, o.Dl):
        for R in range(om.D + 1):
            if P and R == 0:
                continue
            x = rng.standard_normal(om.nsimplices(R))
            y = (ddf.laplace(P, R, dm) * ddf.Fun(dm, P, R, x)).values
            ref = o.spmv_csc(o.laplace(P, R, om), x)
            assert relerr(y, ref) <= 1e-12, (P, R)


# ------------------------------------------------------------------ Op / Fun algebra
def test_op_fun_algebra(ddf, pairs):
    """Ring / vector-space laws of test-ops.jl:183-260 and test-funs.jl on the device."""
    om, dm = pairs("kuhn3n4")
    rng = np.random.default_rng(5)
    f = ddf.Fun(dm, ddf.Pr, 0, rng.integers(-9, 9, om.nsimplices(0)).astype(float))
    g = ddf.Fun(dm, ddf.Pr, 0, rng.integers(-9, 9, om.nsimplices(0)).astype(float))
    d0, d1 = ddf.deriv(ddf.Pr, 0, dm), ddf.deriv(ddf.Pr, 1, dm)
    assert (d0 * (f + g)) == (d0 * f + d0 * g)                  # A*(f+g) == A*f + A*g  (integers: exact)
    assert ((d1 * d0) * f) == (d1 * (d0 * f))                   # (F*A)*f == F*(A*f)
    assert ((2.0 * d0) * f) == (2.0 * (d0 * f))
    assert ((d0 + d0) * f) == ((2.0 * d0) * f)
    assert ((d0 - d0) * f).norm(np.inf) == 0
    assert (-(d0)) * f == (d0 * (-f))
    assert (f - g) + g == f and (f * 2.0) / 2.0 == f
    assert d0.adjoint().shape == (om.nsimplices(0), om.nsimplices(1))
    h = ddf.Fun(dm, ddf.Pr, 1, rng.integers(-9, 9, om.nsimplices(1)).astype(float))
    # <d0 f, h> == <f, d0' h> exactly on integers
    assert (d0 * f).vdot(h) == f.vdot(d0.adjoint() * h)      # plain Euclidean products (Fun.dot is the weighted one)
    E, B = ddf.one(ddf.Pr, 0, dm), ddf.isboundary(ddf.Pr, 0, dm)
    m = 1.0 - om_isb(om)
    assert np.array_equal(((E - B) * f).values, m * f.values)   # wave.jl:93
    with pytest.raises(ddf.DDFError):
        d1 * f                                                   # type mismatch (R)
    with pytest.raises(ddf.DDFError):
        ddf.deriv(ddf.Pr, om.D, dm)                              # @assert 0 <= R+s <= D


def om_isb(om):
    return o.calc_isboundary(om, 0).astype(np.float64)


def test_from_csc_bit_exact(ddf, pairs):
    om, dm = pairs("kuhn2n4")
    rng = np.random.default_rng(6)
    n0, n1 = om.nsimplices(0), om.nsimplices(1)
    A = sp.random(n1, n0, density=0.2, random_state=7, format="csc")
    op = ddf.op_from_csc(dm, ddf.Pr, 1, ddf.Pr, 0, A)
    x = rng.standard_normal(n0)
    y = (op * ddf.Fun(dm, ddf.Pr, 0, x)).values
    assert np.array_equal(y, o.spmv_csc(A, x))
    B = op.to_csc()
    A.sort_indices()
    assert np.array_equal(B.indptr, A.indptr) and np.array_equal(B.indices, A.indices) and np.array_equal(B.data, A.data)
    yt = (op.adjoint() * ddf.Fun(dm, ddf.Pr, 1, rng.standard_normal(n1))).values
    assert yt.shape == (n0,)


def test_export_deriv_and_hodge(ddf, pairs):
    om, dm = pairs("kuhn3n4")
    for k in range(om.D):
        A = ddf.deriv(ddf.Pr, k, dm).to_csc()
        B = o.deriv(o.Pr, k, om).astype(np.float64)
        B.sort_indices()
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)
        A = ddf.deriv(ddf.Dl, k + 1, dm).to_csc()
        B = o.deriv(o.Dl, k + 1, om).astype(np.float64)
        B.sort_indices()
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)


# ------------------------------------------------------- N3: assembled Op algebra (SpGEMM)
@pytest.mark.parametrize("kind", ["kuhn2n4", "kuhn3n4", "simplex3", "refined2l4"])
def test_assembled_products_bit_exact(ddf, pairs, kind):
    """Op*Op and Op+Op as ASSEMBLED matrices (Ops.jl:186-196,228-232: SparseArrays spmatmul, dnz/chop after every
    operation): laplace(P,R) for every (P,R), delta*delta, adjoints and user-composed products, exported with
    to_csc(), against the oracle's restatement of Julia's spmatmul order -- pattern and values bit for bit."""
    om, dm = pairs(kind)
    om = with_device_geometry(om, dm)
    D = om.D

    def same(A, B, what):
        B = B.tocsc()
        B.sort_indices()
        assert A.shape == B.shape, what
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices), what
        assert np.array_equal(A.data, B.data), what

    for P in (o.Pr, o.Dl):
        for R in range(D + 1):
            same(ddf.laplace(P, R, dm).to_csc(), o.laplace(P, R, om), ("laplace", P, R))
    for R in range(1, D):                                  # delta_R delta_{R+1} (test-manifoldops.jl:116-131)
        A = (ddf.coderiv(ddf.Pr, R, dm) * ddf.coderiv(ddf.Pr, R + 1, dm)).to_csc()
        B = o.chop_sparse(o.julia_spgemm(o.coderiv(o.Pr, R, om), o.coderiv(o.Pr, R + 1, om)))
        same(A, B, ("delta delta", R))
    d0, d1 = ddf.deriv(ddf.Pr, 0, dm), (ddf.deriv(ddf.Pr, 1, dm) if D >= 2 else None)
    od0 = o.deriv(o.Pr, 0, om).astype(np.float64)
    if d1 is not None:
        od1 = o.deriv(o.Pr, 1, om).astype(np.float64)
        assert (d1 * d0).to_csc().nnz == 0                 # dd = 0 exactly, every partial sum chopped away
        # (d1' * d1) * (2 * (d0 * d0')) + d0 * d0' : association, scalar, adjoint and sum
        A = ((d1.T * d1) * (2.0 * (d0 * d0.T)) + d0 * d0.T).to_csc()
        X = o.chop_sparse(o.julia_spgemm(od1.T.tocsc(), od1))
        Y = o.chop_sparse(o.julia_spgemm(od0, od0.T.tocsc()))
        Z = o.chop_sparse(o.julia_spgemm(X, o.chop_sparse(2.0 * Y)))
        same(A, o.chop_sparse(Z + Y), "composite")
    # star-weighted: (d0' * hodge1) * d0 versus d0' * (hodge1 * d0): different roundings, both reproduced
    h1 = ddf.hodge(ddf.Pr, 1, dm)
    oh1 = sp.diags(o.hodge_diag(o.Pr, 1, om)).tocsc()
    dd1 = ddf.deriv(ddf.Dl, 1, dm)                         # the matrix d0' typed as a dual derivative (ManifoldOps.jl:46)
    lhs = ((dd1 * h1) * d0).to_csc()
    same(lhs, o.chop_sparse(o.julia_spgemm(o.chop_sparse(od0.T.tocsc() @ oh1), od0)), "left assoc")
    rhs = (dd1 * (h1 * d0)).to_csc()
    same(rhs, o.chop_sparse(o.julia_spgemm(od0.T.tocsc(), o.chop_sparse(oh1 @ od0))), "right assoc")
    # the Dirichlet-projected Poisson operator of examples/poisson.jl:197 (u block)
    b0 = sp.diags(o.calc_isboundary(om, 0).astype(np.float64)).tocsc()
    e0 = sp.identity(om.nsimplices(0), format="csc")
    same(ddf.poisson_operator(dm).to_csc(), o.chop_sparse(o.chop_sparse((e0 - b0) @ o.laplace(o.Pr, 0, om)) + b0), "A'")


@pytest.mark.parametrize("D,n", [(2, 6), (2, 48), (3, 4), (3, 14)])
def test_assembled_apply_bit_exact(ddf, ctx, D, n):
    """Op.assemble(): the operator held as the reference holds it (its assembled, chopped sparse matrix, Ops.jl:25)
    and applied the way SparseArrays applies it (Ops.jl:273-276): same pattern and values as to_csc() of the
    expression, and A*x bit for bit the oracle's Julia-order sparse mat-vec of that matrix -- on meshes small enough
    for the plain row kernel and large enough for the SELL one."""
    dm = ddf.large_hypercube_manifold(D, (n,) * D, ctx=ctx)
    rng = np.random.default_rng(3 * D + n)
    ops = {"poisson": ddf.poisson_operator(dm), "laplace1": ddf.laplace(ddf.Pr, 1, dm),
           "delta1": ddf.coderiv(ddf.Pr, 1, dm), "d0": ddf.deriv(ddf.Pr, 0, dm), "star1": ddf.hodge(ddf.Pr, 1, dm),
           "dd": ddf.deriv(ddf.Pr, 1, dm) * ddf.deriv(ddf.Pr, 0, dm)}
    for name, A in ops.items():
        M = A.to_csc()
        B = A.assemble()
        assert B.shape == A.shape and (B.P1, B.R1, B.P2, B.R2) == (A.P1, A.R1, A.P2, A.R2)
        N = B.to_csc()
        assert np.array_equal(M.indptr, N.indptr) and np.array_equal(M.indices, N.indices), name
        assert np.array_equal(M.data, N.data), name
        x = rng.standard_normal(A.shape[1])
        y = (B * ddf.Fun(dm, A.P2, A.R2, x)).values
        assert np.array_equal(y, o.spmv_csc(M, x)), name
        # and it composes like any other operator
        z = ((2.0 * B) * ddf.Fun(dm, A.P2, A.R2, x)).values
        assert np.allclose(z, 2.0 * y, rtol=1e-15, atol=0)
    dm.close()


# ------------------------------------------------------------------ N2: Poisson solve
@pytest.mark.parametrize("kind", ["refined2l4", "kuhn2n16", "kuhn3n4", "kuhn3n8"])
def test_poisson_solve_matches_direct_solve(ddf, pairs, kind):
    """examples/poisson.jl:124-227: the device solve (BiCGStab(2) around (E-B) L + B, f = d u) against a sparse
    direct solve of the oracle's block system A' x = b' built exactly as the reference builds it."""
    om, dm = pairs(kind)
    om = with_device_geometry(om, dm)
    rho, u0 = o.poisson_rho(om.coords), o.poisson_u_exact(om.coords)
    ou, of = o.poisson_solve(om, rho, u0)
    u, f, info = ddf.poisson_solve(dm, ddf.Fun(dm, ddf.Pr, 0, rho), ddf.Fun(dm, ddf.Pr, 0, u0), reltol=1e-14,
                                   max_mv_products=20000)
    assert info["converged"], info
    assert relerr(u.values, ou) <= 1e-9, (relerr(u.values, ou), info)
    assert relerr(f.values, of) <= 1e-8
    b = o.calc_isboundary(om, 0)
    assert np.max(np.abs(u.values[b] - u0[b])) <= 1e-12 * np.abs(u0).max()      # Dirichlet rows
    # the same solve on the assembled A' (what the reference's solver is handed)
    ua, fa, infa = ddf.poisson_solve(dm, ddf.Fun(dm, ddf.Pr, 0, rho), ddf.Fun(dm, ddf.Pr, 0, u0), assembled=True,
                                     reltol=1e-14, max_mv_products=20000)
    assert infa["converged"] and relerr(ua.values, ou) <= 1e-9 and relerr(fa.values, of) <= 1e-8
    # the reference's own stopping rule (IterativeSolvers defaults) converges too
    _, _, out = ddf.poisson(om.D, 0, mfd=dm)
    assert out["converged"] and out["residual2"] < 1e-5 * max(1.0, np.abs(rho).max())


def test_bicgstabl_generic_operator(ddf, pairs):
    """The solver takes any Op: a shifted Hodge-Laplacian on edges, checked through the residual."""
    om, dm = pairs("kuhn2n16")
    A = ddf.laplace(ddf.Pr, 1, dm) - 50.0 * ddf.one(ddf.Pr, 1, dm)
    rng = np.random.default_rng(17)
    xs = ddf.Fun(dm, ddf.Pr, 1, rng.standard_normal(om.nsimplices(1)))
    b = A * xs
    x, info = ddf.bicgstabl(A, b, reltol=1e-13, max_mv_products=40000)
    assert info["converged"], info
    assert (A * x - b).norm(2) <= 1e-11 * b.norm(2)
    assert relerr(x.values, xs.values) <= 1e-7


@pytest.mark.parametrize("D,n", [(2, 24), (3, 8), (3, 14)])
def test_poisson_jacobi_bit_exact(ddf, ctx, D, n):
    """north_star kernel (3), the Poisson flavour of the fused step: weighted-Jacobi sweeps, one pass per sweep,
    bit for bit against the oracle for every row storage / kernel (persistent pipeline, one-block-per-window staged
    kernel, SELL, CSR); and the sweeps do reduce the residual."""
    om = o.large_hypercube_manifold(D, (n,) * D)
    theta = 2.0 / 3.0
    for tune in (0, 128, 16, 17):
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        try:
            dm = ddf.large_hypercube_manifold(D, (n,) * D, ctx=ctx)
            omg = with_device_geometry(om, dm)
            rho, u0 = o.poisson_rho(omg.coords), o.poisson_u_exact(omg.coords)
            b = o.calc_isboundary(omg, 0)
            start = np.where(b, u0, 0.0)                  # Dirichlet values on the boundary, zero inside
            want = o.poisson_jacobi(omg, start, rho, theta, 5)
            u = ddf.Fun(dm, ddf.Pr, 0, start)
            ddf.poisson_jacobi(dm, u, ddf.Fun(dm, ddf.Pr, 0, rho), theta, 5)
            assert np.array_equal(u.values, want), (D, n, tune, ddf.wave_format(dm))
            dm.close()
        finally:
            ddf._lib.call("ddf_set_tuning", b"wave", 0)
    L = o.laplace(o.Pr, 0, omg)

    def res(x):
        return np.linalg.norm((rho - o.spmv_csc(L, x))[~b])
    many = o.poisson_jacobi(omg, start, rho, theta, 60)
    assert res(many) < 0.7 * res(start)
    assert np.array_equal(many[b], u0[b])


# ------------------------------------------------------------------------ A10 wave
@pytest.mark.parametrize("kind", ["kuhn2n4", "kuhn2n16", "kuhn3n4", "kuhn3n8"])
def test_wave_rhs_and_leapfrog_bit_exact(ddf, pairs, kind):
    om, dm = pairs(kind)
    om = with_device_geometry(om, dm)
    u0 = o.wave_u0(om.coords)
    rng = np.random.default_rng(8)
    v0 = rng.standard_normal(len(u0))
    du, dv = ddf.wave_rhs(dm, ddf.Fun(dm, ddf.Pr, 0, u0), ddf.Fun(dm, ddf.Pr, 0, v0))
    odu, odv = o.wave_rhs(om, u0, v0)
    assert np.array_equal(du.values, odu) and np.array_equal(dv.values, odv)
    dt = ddf.wave_dt(dm)
    assert dt == o.wave_dt(om)
    u, v = ddf.Fun(dm, ddf.Pr, 0, u0), ddf.Fun(dm, ddf.Pr, 0, np.zeros_like(u0))
    ddf.wave_leapfrog(dm, u, v, dt, 7)
    ou, ov = o.wave_leapfrog(om, u0, np.zeros_like(u0), dt, 7)
    assert np.array_equal(u.values, ou) and np.array_equal(v.values, ov)


def test_wave_row_formats_agree(ddf, ctx):
    """The three storage formats of the L rows -- staged rows (TMA-staged gather, 16-bit slots, jagged
    slices; default), explicit-index SELL-32-256 (tuning 16) and CSR thread-per-row (tuning 16|1) -- give
    the same bits as the oracle, also on a mesh with shuffled vertex numbers (no column locality: the
    staged build must refuse it and fall back)."""
    n = 20        # 9261 vertices: a window of the shuffled mesh touches more columns than a staging buffer holds
    simp, coords, weights = o.large_hypercube_tables(3, (n, n, n))
    rng = np.random.default_rng(21)
    perm = rng.permutation(coords.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    cases = {"ordered": (simp, coords, weights), "shuffled": (inv[simp], coords[perm], weights[perm])}
    for name, (s_, c_, w_) in cases.items():
        om = o.build_manifold("m", s_, c_, w_)
        ref = None
        for tune in (0, 16, 17):
            ddf._lib.call("ddf_set_tuning", b"wave", tune)
            try:
                dm = ddf.Manifold.from_simplices("m", s_, c_, w_, ctx=ctx)
                omg = with_device_geometry(om, dm)
                u0 = rng.standard_normal(om.nsimplices(0))
                v0 = rng.standard_normal(om.nsimplices(0))
                u, v = ddf.Fun(dm, ddf.Pr, 0, u0), ddf.Fun(dm, ddf.Pr, 0, v0)
                dt = ddf.wave_dt(dm)
                ddf.wave_leapfrog(dm, u, v, dt, 3)
                ou, ov = o.wave_leapfrog(omg, u0, v0, dt, 3)
                assert np.array_equal(u.values, ou) and np.array_equal(v.values, ov), (name, tune)
                y = (ddf.laplace(ddf.Pr, 0, dm) * ddf.Fun(dm, ddf.Pr, 0, u0)).values
                assert np.array_equal(y, o.spmv_csc(o.laplace(o.Pr, 0, omg), u0)), (name, tune)
                fmt = ddf.wave_format(dm)
                if tune == 0:
                    assert (fmt == "staged") == (name == "ordered"), (name, fmt)
                elif tune == 17:
                    assert fmt != "staged"
                dm.close()
            finally:
                ddf._lib.call("ddf_set_tuning", b"wave", 0)


@pytest.mark.parametrize("kind", ["kuhn2n4", "kuhn2n16", "kuhn3n4"])
def test_wave_tsit5_bit_exact(ddf, pairs, kind):
    """N1: the reference's integrator (fixed-step Tsit5, wave.jl:134) on the device == the oracle's restatement
    of OrdinaryDiffEq's perform_step! (same rounding sequence: no contraction), including FSAL across steps."""
    om, dm = pairs(kind)
    om = with_device_geometry(om, dm)
    u0 = o.wave_u0(om.coords)
    rng = np.random.default_rng(12)
    v0 = rng.standard_normal(len(u0))
    dt = ddf.wave_dt(dm)
    u, v = ddf.Fun(dm, ddf.Pr, 0, u0), ddf.Fun(dm, ddf.Pr, 0, v0)
    ddf.wave_tsit5(dm, u, v, dt, 3)
    ou, ov = o.wave_tsit5(om, u0, v0, dt, 3)
    assert np.array_equal(u.values, ou) and np.array_equal(v.values, ov)
    # the stage broadcast on its own, in place and out of place
    ks = [ddf.Fun(dm, ddf.Pr, 0, rng.standard_normal(len(u0))) for _ in range(4)]
    cs = [0.25, -1.5, 3.0, 0.125]
    ref = o.lincomb(u0, dt, cs, [k.values for k in ks])
    base = ddf.Fun(dm, ddf.Pr, 0, u0)
    assert np.array_equal(ddf.lincomb(base, dt, cs, ks).values, ref)
    assert np.array_equal(ddf.lincomb(base, dt, cs, ks, out=base).values, ref)


@pytest.mark.parametrize("shape,tune", [((3, 40), 0), ((3, 40), 1 << 30), ((3, 40), 128), ((3, 12), 16), ((3, 12), 17),
                                        ((2, 96), 0), ((2, 96), 1 << 30), ((2, 96), 16), ((3, 72), 0)])
def test_wave_tsit5_fused_stages_equal_unfused(ddf, ctx, shape, tune):
    """ddf_wave_tsit5 against the unfused sequence of k_lincomb passes and wave right-hand sides (tuning bit 1 << 29) that
    test_wave_tsit5_bit_exact pins to the oracle: same bits from a random state, odd and even step counts, on every row
    storage.  Default on staged rows: THREE passes over L per step, two stages each (csrc/wave.cu tsit5_pair: su_l does
    not depend on k^v_{l-1}, so stages J and J+1 share one read of L); 1 << 30, the per-window staged kernel (128), SELL
    (16) and CSR (17): six fused single-stage launches (row sum + pointwise recomputation of the stage inputs)."""
    D, n = shape
    ddf._lib.call("ddf_set_tuning", b"wave", tune)
    try:
        m = ddf.large_hypercube_manifold(D, (n,) * D, ctx=ctx)
        dt = ddf.wave_dt(m)
        rng = np.random.default_rng(n + (tune & 0xffff))
        u0, v0 = rng.standard_normal(m.nsimplices(0)), rng.standard_normal(m.nsimplices(0))
        for steps in (1, 4, 5):
            ru, rv = ddf.Fun(m, ddf.Pr, 0, u0), ddf.Fun(m, ddf.Pr, 0, v0)
            ddf._lib.call("ddf_set_tuning", b"wave", tune | (1 << 29))
            ddf.wave_tsit5(m, ru, rv, dt, steps)
            ddf._lib.call("ddf_set_tuning", b"wave", tune)
            u, v = ddf.Fun(m, ddf.Pr, 0, u0), ddf.Fun(m, ddf.Pr, 0, v0)
            ctx.sync()
            l0 = ctx.launch_count()
            ddf.wave_tsit5(m, u, v, dt, steps)
            paired = tune == 0 and ddf.wave_format(m) == "staged"
            assert ctx.launch_count() - l0 == (3 * steps + 1 if paired else 6 * steps), (ctx.launch_count() - l0, steps)
            assert np.array_equal(u.values, ru.values) and np.array_equal(v.values, rv.values), (shape, tune, steps)
        assert ddf.wave_format(m) == ("staged" if tune in (0, 1 << 30, 128) else "sell")
        m.close()
    finally:
        ddf._lib.call("ddf_set_tuning", b"wave", 0)


def test_wave_tsit5_converges_to_exact_solution(ddf, ctx):
    """Wave.wave(Val(2), levels) with the reference's integrator: the error against
    cos(sqrt(2) pi t) prod sin(pi x) (wave.jl:150-163) falls at the same rate as with the leapfrog step
    (the spatial error of the DEC Laplacian dominates; the accuracy table of wave.jl:172-186 is stale)."""
    errs = []
    for lv in (3, 4, 5):
        e, _ = ddf.wave(2, lv, ctx=ctx, t1=0.5, integrator="tsit5")
        errs.append(e)
    assert errs[1] < errs[0] / 2.5 and errs[2] < errs[1] / 2.5, errs


def test_wave_staged_wide_planes(ddf, ctx):
    """A slab with 513 x 513 vertex planes (the cross-section of BASELINE config 5's n=512 cube): a window's
    columns span 2 planes = 526k vertex ids.  The staged build must accept it (no span limit) and agree bit for bit
    with the explicit-index SELL kernel."""
    res = {}
    for tune in (0, 16):
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        try:
            m = ddf.large_hypercube_manifold(3, (512, 512, 4), ctx=ctx, hdiv=512)
            u, v = ddf.sample(m, "wave"), ddf.Fun(m, ddf.Pr, 0)
            ddf.wave_leapfrog(m, u, v, ddf.wave_dt(m), 3)
            res[tune] = (ddf.wave_format(m), u.values, v.values)
            m.close()
        finally:
            ddf._lib.call("ddf_set_tuning", b"wave", 0)
    assert res[0][0] == "staged" and res[16][0] == "sell"
    assert np.array_equal(res[0][1], res[16][1]) and np.array_equal(res[0][2], res[16][2])
    assert np.abs(res[0][1]).max() > 0.01


def test_wave_pipeline_repeatable(ddf, ctx):
    """The persistent TMA pipeline shares its stage ring between consumer groups; a consumer that entered a
    parity wait one phase early once read a stage before its data had landed (rare, size dependent).  Guarded by
    the producer's `issued` counter now: repeated runs at a size with ~8400 windows must be bit-identical to each
    other and to the explicit-index SELL kernel."""
    n = 128
    ref = None
    for tune, reps in ((16, 1), (0, 6), (768, 3)):
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        try:
            m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
            u0 = ddf.sample(m, "wave")
            dt = ddf.wave_dt(m)
            for _ in range(reps):
                u, v = u0.copy(), ddf.Fun(m, ddf.Pr, 0)
                ddf.wave_leapfrog(m, u, v, dt, 8)
                out = (u.values, v.values)
                if ref is None:
                    ref = out
                assert np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1]), tune
            m.close()
        finally:
            ddf._lib.call("ddf_set_tuning", b"wave", 0)


@pytest.mark.parametrize("n", [64, 128])
def test_wavefront2_equals_single_steps(ddf, ctx, n):
    """Two leapfrog steps per launch (wavefront through L2, csrc/wavefront2.cuh, tuning bit 65536) against the
    one-step-per-launch kernel AND the oracle-checked explicit-index SELL kernel: same bits after an even and an odd
    number of steps, repeatably, from a random (non-smooth) state so that a stale or early read cannot hide."""
    m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)        # n=64: 1073 windows, lag 2; n=128: 8386 windows
    dt = ddf.wave_dt(m)
    rng = np.random.default_rng(n)
    u0, v0 = rng.standard_normal(m.nsimplices(0)), rng.standard_normal(m.nsimplices(0))
    for steps in (6, 7):
        ref_u, ref_v = ddf.Fun(m, ddf.Pr, 0, u0), ddf.Fun(m, ddf.Pr, 0, v0)
        ddf._lib.call("ddf_set_tuning", b"wave", 131072)          # one step per launch
        try:
            ddf.wave_leapfrog(m, ref_u, ref_v, dt, steps)
        finally:
            ddf._lib.call("ddf_set_tuning", b"wave", 0)
        ru, rv = ref_u.values, ref_v.values
        for slack in (0, 1, 4) if n >= 128 else (0, 0, 0):      # tuning bits 18-21 hold slack + 1 (extra iterations of lag)
            u, v = ddf.Fun(m, ddf.Pr, 0, u0), ddf.Fun(m, ddf.Pr, 0, v0)
            ctx.sync()
            l0 = ctx.launch_count()
            ddf._lib.call("ddf_set_tuning", b"wave", 65536 | ((slack + 1) << 18))
            try:
                ddf.wave_leapfrog(m, u, v, dt, steps)
            finally:
                ddf._lib.call("ddf_set_tuning", b"wave", 0)
            # the two-step kernel really ran: one launch per step pair (+ the row-reach kernel the first time)
            assert ctx.launch_count() - l0 <= steps // 2 + (steps % 2) + 1, (ctx.launch_count() - l0, steps)
            assert np.array_equal(u.values, ru) and np.array_equal(v.values, rv), (steps, slack)
    m.close()


def test_wave_row_formats_2d(ddf, ctx):
    """2D mesh (7 entries per row): staged rows in groups of 2 windows per pipeline stage (default), staged with one
    window per stage (tuning 4096), SELL with 16-bit column offsets (16), with 32-bit indices (16|2048) and the CSR
    kernel (17) all reproduce the oracle bit for bit."""
    n = 48
    om = o.large_hypercube_manifold(2, (n, n))
    rng = np.random.default_rng(22)
    u0 = rng.standard_normal(om.nsimplices(0))
    v0 = rng.standard_normal(om.nsimplices(0))
    want = {0: "staged", 4096: "staged", 16: "sell", 16 | 2048: "sell", 17: "sell"}   # 17: SELL storage, CSR kernel
    for tune, fmt in want.items():
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        try:
            dm = ddf.large_hypercube_manifold(2, (n, n), ctx=ctx)
            omg = with_device_geometry(om, dm)
            u, v = ddf.Fun(dm, ddf.Pr, 0, u0), ddf.Fun(dm, ddf.Pr, 0, v0)
            dt = ddf.wave_dt(dm)
            ddf.wave_leapfrog(dm, u, v, dt, 5)
            ou, ov = o.wave_leapfrog(omg, u0, v0, dt, 5)
            assert np.array_equal(u.values, ou) and np.array_equal(v.values, ov), tune
            assert ddf.wave_format(dm) == fmt, (tune, ddf.wave_format(dm))
            dm.close()
        finally:
            ddf._lib.call("ddf_set_tuning", b"wave", 0)


def test_wave_sample_and_convergence(ddf, ctx):
    """Device-sampled initial data and second-order convergence of the leapfrog
    solution (the oracle restatement of wave.jl is second order)."""
    errs = []
    for lv in (3, 4, 5):
        e, mfd = ddf.wave(2, lv, ctx=ctx, t1=0.5)
        errs.append(e)
    assert errs[1] < errs[0] / 2.5 and errs[2] < errs[1] / 2.5, errs
    m = ddf.large_hypercube_manifold(2, (8, 8), ctx=ctx)
    assert relerr(ddf.sample(m, "wave").values, o.wave_u0(m.get_coords())) < 1e-14
    assert relerr(ddf.sample(m, "readme").values, o.readme_field(m.get_coords())) < 1e-14
    assert relerr(ddf.sample(m, "poisson_rho").values, o.poisson_rho(m.get_coords())) < 1e-13


# --------------------------------------------------------- size-independent checks
def test_large_mesh_properties(ddf, ctx):
    """3D Kuhn n=64 (1.57e6 tets): closed-form counts, dd = 0, volume sums,
    d applied to coordinate functions, L(1) = 0 on interior rows."""
    n = 64
    m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
    assert [m.nsimplices(R) for R in range(4)] == [(n + 1) ** 3, 7 * n ** 3 + 9 * n ** 2 + 3 * n,
                                                   12 * n ** 3 + 6 * n ** 2, 6 * n ** 3]
    rng = np.random.default_rng(9)
    x = ddf.Fun(m, ddf.Pr, 0, rng.integers(-99, 99, m.nsimplices(0)).astype(float))
    d0, d1, d2 = (ddf.deriv(ddf.Pr, k, m) for k in range(3))
    assert (d1 * (d0 * x)).norm(np.inf) == 0
    y = ddf.Fun(m, ddf.Pr, 1, rng.integers(-99, 99, m.nsimplices(1)).astype(float))
    assert (d2 * (d1 * y)).norm(np.inf) == 0
    assert abs(m.get_volumes(3).sum() - 1) < 1e-12 and abs(m.get_dualvolumes(0).sum() - 1) < 1e-12
    for R in range(4):
        assert m.get_dualvolumes(R).min() > 0
    one = ddf.Fun(m, ddf.Pr, 0, np.ones(m.nsimplices(0)))
    L1 = (ddf.laplace(ddf.Pr, 0, m) * one).values
    scale = n * n
    assert np.max(np.abs(L1)) <= 1e-12 * scale * 64
    # linearity of the fused step: leapfrog(a u) == a leapfrog(u) for a power of two
    u0 = ddf.sample(m, "wave")
    u1, v1 = u0.copy(), ddf.Fun(m, ddf.Pr, 0)
    u2, v2 = u0 * 4.0, ddf.Fun(m, ddf.Pr, 0)
    dt = ddf.wave_dt(m)
    ddf.wave_leapfrog(m, u1, v1, dt, 5)
    ddf.wave_leapfrog(m, u2, v2, dt, 5)
    assert np.array_equal(u2.values, 4.0 * u1.values)


# ------------------------------------------------- BASELINE configs at their full sizes
def test_config2_poisson_mesh_full_size(ddf, ctx):
    """BASELINE config 2: 2D simplex refined 10x (525,825 / 1,574,400 / 1,048,576) -- Delaunay through
    scipy like the reference, assembly and L on the device.  Full-size checks: counts, dd = 0, sum of volumes,
    and L*rho / d0*rho against the oracle on the SAME mesh (bit-exact given the device geometry)."""
    dm = ddf.refined_simplex_manifold(2, 10, ctx=ctx)
    assert [dm.nsimplices(R) for R in range(3)] == [525825, 1574400, 1048576]
    rng = np.random.default_rng(31)
    x = ddf.Fun(dm, ddf.Pr, 0, rng.integers(-99, 99, dm.nsimplices(0)).astype(float))
    d0, d1 = ddf.deriv(ddf.Pr, 0, dm), ddf.deriv(ddf.Pr, 1, dm)
    assert (d1 * (d0 * x)).norm(np.inf) == 0
    area = np.sqrt(3) / 4                                    # regular simplex of edge length 1
    assert abs(dm.get_volumes(2).sum() - area) < 1e-12 and abs(dm.get_dualvolumes(0).sum() - area) < 1e-12
    # the oracle on the same tables (the triangle table is read back from the device: Qhull's order)
    om = o.build_manifold("c2", dm.get_simplices(2), dm.get_coords(), dm.get_weights())
    for R in (1, 2):
        cp, rv, nz = dm.get_boundaries_csc(R)
        ocp, orv, onz = o.to_julia_csc(om.boundaries[R])
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz)
    om = with_device_geometry(om, dm)
    rho = o.poisson_rho(om.coords)
    frho = ddf.Fun(dm, ddf.Pr, 0, rho)
    assert np.array_equal((d0 * frho).values, o.apply_deriv(o.Pr, 0, om, rho))
    y = (ddf.laplace(ddf.Pr, 0, dm) * frho).values
    assert np.array_equal(y, o.spmv_csc(o.laplace(o.Pr, 0, om), rho))


def test_config3_wave_2d_full_size(ddf, ctx):
    """BASELINE config 3: 2D Kuhn n=2048 (4,198,401 / 12,587,008 / 8,388,608), fused leapfrog.  Full-size
    checks: closed-form counts, the row format chosen for short rows, the step is linear (exact for a power of
    two), and the solution follows cos(sqrt(2) pi t) prod sin(pi x) (wave.jl:42-58,150-163)."""
    n = 2048
    m = ddf.large_hypercube_manifold(2, (n, n), ctx=ctx)
    assert [m.nsimplices(R) for R in range(3)] == [(n + 1) ** 2, 3 * n * n + 2 * n, 2 * n * n]
    assert ddf.wave_format(m) == "staged"        # short rows: two windows per pipeline stage
    u0 = ddf.sample(m, "wave")
    u1, v1 = u0.copy(), ddf.Fun(m, ddf.Pr, 0)
    u2, v2 = u0 * 8.0, ddf.Fun(m, ddf.Pr, 0)
    dt = ddf.wave_dt(m)
    assert dt == 1.0 / (2 * n)
    nsteps = 64
    ddf.wave_leapfrog(m, u1, v1, dt, nsteps)
    ddf.wave_leapfrog(m, u2, v2, dt, nsteps)
    a = u1.values
    assert np.array_equal(u2.values, 8.0 * a)
    exact = np.cos(np.sqrt(2.0) * np.pi * nsteps * dt) * u0.values
    assert np.max(np.abs(a - exact)) < 2e-4          # first-order in dt for this staggering, O(h^2) in space
    assert np.array_equal(a[m.get_isboundary(0).astype(bool)], u0.values[m.get_isboundary(0).astype(bool)])


def test_config4_3d_full_size(ddf, ctx):
    """BASELINE config 4: 3D Kuhn n=256 (16,974,593 / 118,031,104 / 201,719,808 / 100,663,296): assembly on the
    device, then exact identities on integer-valued cochains (d d = 0 is exact in Float64), d0 of a coordinate
    function, volume sums, row-format choice and linearity of the fused step."""
    n = 256
    m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
    assert [m.nsimplices(R) for R in range(4)] == [(n + 1) ** 3, 7 * n ** 3 + 9 * n ** 2 + 3 * n,
                                                   12 * n ** 3 + 6 * n ** 2, 6 * n ** 3]
    d0, d1, d2 = (ddf.deriv(ddf.Pr, k, m) for k in range(3))
    rng = np.random.default_rng(41)
    x = ddf.Fun(m, ddf.Pr, 0, rng.integers(-999, 999, m.nsimplices(0)).astype(np.float64))
    assert (d1 * (d0 * x)).norm(np.inf) == 0
    y = ddf.Fun(m, ddf.Pr, 1, rng.integers(-999, 999, m.nsimplices(1)).astype(np.float64))
    assert (d2 * (d1 * y)).norm(np.inf) == 0
    # storage of the rows at this size: d0 prefix-packed, d1 one 4-byte word, d2 (top level) one 8-byte word per row;
    # bit for bit the explicit-table kernels
    assert [ddf.deriv_format(m, k) for k in range(3)] == ["packed", "bits32", "bits64"]
    z = ddf.rand_fun(m, ddf.Pr, 2, seed=5)
    got = [(d1 * y).values[::3].copy(), (d2 * z).values[::3].copy()]
    ddf._lib.call("ddf_set_tuning", b"fixed", 18)
    try:
        assert np.array_equal((d1 * y).values[::3], got[0]) and np.array_equal((d2 * z).values[::3], got[1])
    finally:
        ddf._lib.call("ddf_set_tuning", b"fixed", 0)
    del z, got
    # d0 of the coordinate function x: every edge difference is 0 or +-h exactly (h = 2^-8)
    cx = ddf.Fun(m, ddf.Pr, 0, np.ascontiguousarray(m.get_coords()[:, 0]))
    e = (d0 * cx).values
    assert set(np.unique(e).tolist()) <= {0.0, 1.0 / n, -1.0 / n}
    assert abs(m.get_volumes(3).sum() - 1) < 1e-10
    assert ddf.wave_format(m) == "staged"
    u0 = ddf.sample(m, "wave")
    u1, v1 = u0.copy(), ddf.Fun(m, ddf.Pr, 0)
    u2, v2 = u0 * 4.0, ddf.Fun(m, ddf.Pr, 0)
    dt = ddf.wave_dt(m)
    ddf.wave_leapfrog(m, u1, v1, dt, 4)
    ddf.wave_leapfrog(m, u2, v2, dt, 4)
    assert (u2 - u1 * 4.0).norm(np.inf) == 0
    exact = u0 * float(np.cos(np.sqrt(3.0) * np.pi * 4 * dt))
    assert (u1 - exact).norm(np.inf) < 1e-3


def test_async_transfers_pipeline(ddf, ctx):
    """upload_async -> apply -> download_async on three streams gives the same bits as the blocking path,
    also when the same vectors are reused in the next round before the host synchronises."""
    import torch
    om = o.large_hypercube_manifold(3, (8, 8, 8))
    dm = ddf.large_hypercube_manifold(3, (8, 8, 8), ctx=ctx)
    d0, d1 = ddf.deriv(ddf.Pr, 0, dm), ddf.deriv(ddf.Pr, 1, dm)
    n0, n1, n2 = (om.nsimplices(k) for k in range(3))
    hx = torch.empty(n0, dtype=torch.float64).pin_memory().numpy()
    hy = torch.empty(n1, dtype=torch.float64).pin_memory().numpy()
    hz = torch.empty(n2, dtype=torch.float64).pin_memory().numpy()
    x, y, z = ddf.Fun(dm, ddf.Pr, 0), ddf.Fun(dm, ddf.Pr, 1), ddf.Fun(dm, ddf.Pr, 2)
    rng = np.random.default_rng(11)
    for rnd in range(4):
        xin = rng.integers(-99, 99, n0).astype(np.float64) if rnd % 2 else rng.standard_normal(n0)
        hx[:] = xin
        x.upload_async(hx)
        d0.apply(x, y)
        y.download_async(hy)
        d1.apply(y, z)
        z.download_async(hz)
        ctx.sync()
        assert np.array_equal(hy, o.apply_deriv(o.Pr, 0, om, xin))
        assert np.array_equal(hz, o.apply_deriv(o.Pr, 1, om, o.apply_deriv(o.Pr, 0, om, xin)))


def test_empty_and_tiny(ddf, ctx):
    """Edge cases of the reference's zoo: a single simplex per D and ragged tails."""
    for D in (1, 2, 3):
        om = o.simplex_manifold(D)
        dm = ddf.Manifold.from_simplices("s", om.simplices[D], om.coords, ctx=ctx)
        for R in range(D + 1):
            assert dm.nsimplices(R) == len(list(__import__("itertools").combinations(range(D + 1), R + 1)))
    with pytest.raises(ddf.DDFError):
        ddf.Manifold.from_simplices("bad", np.array([[0, 0, 1]]), np.zeros((2, 2)), ctx=ctx)   # repeated vertex
    with pytest.raises(ddf.DDFError):
        ddf.Manifold.from_simplices("bad", np.array([[0, 1, 5]]), np.zeros((3, 2)), ctx=ctx)   # out of range
    with pytest.raises(ddf.DDFError):
        ddf.large_hypercube_manifold(2, (3, 3), ctx=ctx)                                        # odd nelts


def test_plain_c_client(ddf, tmp_path):
    """The ABI from plain C (what Julia's ccall does): tests/c/abi_client.c is compiled with gcc against
    include/ddf_b200.h, linked to the in-tree libddf_b200.so and run on the GPU; it checks hand-derived
    tables, signs and values for the two-triangle square and the error convention."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "abi_client"
    libdir = os.path.dirname(ddf.LIB_PATH)
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "no C compiler"
    subprocess.check_call([cc, "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c", "abi_client.c"), "-o", str(exe),
                           "-L", libdir, "-l:libddf_b200.so", "-lm", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "abi_client ok" in out.stdout


def test_write_vtu_roundtrip(ddf, ctx, tmp_path):
    """examples/poisson.jl:268-295: the solution written as a VTK unstructured grid; parsed back with the
    standard library: points, connectivity, cell types and the point data arrive unchanged (17 significant digits)."""
    import xml.etree.ElementTree as ET
    m = ddf.large_hypercube_manifold(3, (2, 2, 2), ctx=ctx)
    u = ddf.sample(m, "wave")
    vol = ddf.Fun(m, ddf.Pr, 3, m.get_volumes(3))
    path = ddf.write_vtu(str(tmp_path / "wave3d"), m, {"u": u}, {"vol": vol})
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    assert int(piece.get("NumberOfPoints")) == 27 and int(piece.get("NumberOfCells")) == 48
    pts = np.array(piece.find("Points/DataArray").text.split(), dtype=float).reshape(-1, 3)
    assert np.array_equal(pts, m.get_coords())
    arrays = {a.get("Name"): a for a in piece.iter("DataArray")}
    conn = np.array(arrays["connectivity"].text.split(), dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(conn, m.get_simplices(3))
    assert set(arrays["types"].text.split()) == {"10"}
    assert np.array_equal(np.array(arrays["u"].text.split(), dtype=float), u.values)
    assert np.array_equal(np.array(arrays["vol"].text.split(), dtype=float), vol.values)


@pytest.mark.parametrize("D,npts,seed", [(2, 400, 0), (2, 3000, 1), (3, 300, 2), (3, 2500, 3)])
def test_random_delaunay_meshes(ddf, ctx, D, npts, seed):
    """Irregular input: the Delaunay triangulation of random points (vertex degrees 3..40, no numbering locality,
    ragged row lengths, boundary faces everywhere).  Assembly, every d_k / boundary apply, the Hodge stars, the
    assembled Laplacian, its apply and the fused wave step must still match the oracle bit for bit (the oracle is
    given the device geometry: slivers make circumcentres ill-conditioned, which is compared on the regular zoo)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(100 + seed)
    pts = rng.random((npts, D))
    simp = np.asarray(Delaunay(pts).simplices, dtype=np.int64)
    used = np.unique(simp)
    assert len(used) == npts                       # Qhull keeps every point of a random cloud
    om = o.build_manifold("delaunay", simp, pts)
    dm = ddf.Manifold.from_simplices("delaunay", simp, pts, ctx=ctx)
    for R in range(1, D + 1):
        cp, rv, nz = dm.get_boundaries_csc(R)
        ocp, orv, onz = o.to_julia_csc(om.boundaries[R])
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz), R
    om = with_device_geometry(om, dm)
    for k in range(D):
        x = rng.standard_normal(om.nsimplices(k))
        y = (ddf.deriv(ddf.Pr, k, dm) * ddf.Fun(dm, ddf.Pr, k, x)).values
        assert np.array_equal(y, o.apply_deriv(o.Pr, k, om, x)), ("d", k)
        z = rng.standard_normal(om.nsimplices(k + 1))
        w = (ddf.deriv(ddf.Dl, k + 1, dm) * ddf.Fun(dm, ddf.Dl, k + 1, z)).values      # the matrix d_{k+1} by rows
        assert np.array_equal(w, o.apply_deriv(o.Dl, k + 1, om, z)), ("dual d", k + 1)
    for R in range(D + 1):
        x = rng.standard_normal(om.nsimplices(R))
        assert np.array_equal((ddf.hodge(ddf.Pr, R, dm) * ddf.Fun(dm, ddf.Pr, R, x)).values,
                              o.hodge_diag(o.Pr, R, om) * x), ("hodge", R)
    L = o.laplace(o.Pr, 0, om)
    A = ddf.laplace(ddf.Pr, 0, dm).to_csc()
    assert np.array_equal(A.indptr, L.indptr) and np.array_equal(A.indices, L.indices) and np.array_equal(A.data, L.data)
    u0 = rng.standard_normal(npts)
    assert np.array_equal((ddf.laplace(ddf.Pr, 0, dm) * ddf.Fun(dm, ddf.Pr, 0, u0)).values, o.spmv_csc(L, u0))
    v0 = rng.standard_normal(npts)
    dt = 1e-3
    u, v = ddf.Fun(dm, ddf.Pr, 0, u0), ddf.Fun(dm, ddf.Pr, 0, v0)
    ddf.wave_leapfrog(dm, u, v, dt, 4)
    ou, ov = o.wave_leapfrog(om, u0, v0, dt, 4)
    assert np.array_equal(u.values, ou) and np.array_equal(v.values, ov)


@pytest.mark.parametrize("kind", ["simplex4", "simplex5", "kuhn4n2", "kuhn5n2"])
def test_topology_in_higher_dimensions(ddf, ctx, kind):
    """The reference's topology tests run up to D = 5 (test/test-manifolds.jl).  Assembly (128-bit face keys for the
    wider tuples), every boundary matrix, d_k applies through the W = 5, 6 fixed-width kernels and d d = 0 on the
    4- and 5-dimensional simplex and Kuhn cube; geometry is 2D/3D only (DDF_MAXC)."""
    D = int(kind[-3]) if kind.startswith("kuhn") else int(kind[-1])
    if kind.startswith("simplex"):
        simp = np.arange(D + 1, dtype=np.int64)[None, :]
        coords = o.regular_simplex(D)
    else:
        simp, coords, _ = o.large_hypercube_tables(D, (2,) * D)
    V = coords.shape[0]
    om = o.manifold_from_simplices(kind, simp, V, C=D)          # dummy coordinates: topology only
    dm = ddf.Manifold.from_simplices(kind, simp, np.ascontiguousarray(om.coords), ctx=ctx)
    assert [dm.nsimplices(R) for R in range(D + 1)] == [om.nsimplices(R) for R in range(D + 1)]
    for R in range(1, D + 1):
        cp, rv, nz = dm.get_boundaries_csc(R)
        ocp, orv, onz = o.to_julia_csc(om.boundaries[R])
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz), R
        assert np.array_equal(dm.get_simplices(R), om.simplices[R]), R
    rng = np.random.default_rng(7)
    prev = None
    for k in range(D):
        x = rng.integers(-9, 9, om.nsimplices(k)).astype(np.float64)
        d = ddf.deriv(ddf.Pr, k, dm)
        y = d * ddf.Fun(dm, ddf.Pr, k, x)
        assert np.array_equal(y.values, o.apply_deriv(o.Pr, k, om, x)), k
        if prev is not None:
            assert (d * prev).norm(np.inf) == 0                       # d_k d_{k-1} x = 0 exactly
        prev = y
        z = rng.integers(-9, 9, om.nsimplices(k + 1)).astype(np.float64)
        w = (ddf.boundary(ddf.Pr, k + 1, dm) * ddf.Fun(dm, ddf.Pr, k + 1, z)).values
        assert np.array_equal(w, om.boundaries[k + 1].astype(np.float64) @ z), ("boundary", k + 1)


def test_assembly_bucket_ranking_equals_radix_sort(ddf, ctx):
    """K1 has two paths: the bucket ranking (count / scatter by smallest vertex, warp-sorted buckets; default) and
    the global radix sort of the full tuples (tuning asm=1; also the fallback for wide tuples).  Every table they
    produce -- simplices, boundary CSC with signs, and through it the row-CSR -- must be identical, on a Kuhn
    mesh, an irregular Delaunay mesh (large, ragged buckets) and in 2D."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(77)
    pts3 = rng.random((4000, 3))
    cases = [("kuhn3", None, (40, 40, 40)), ("kuhn2", None, (300, 300)),
             ("delaunay3", (np.asarray(Delaunay(pts3).simplices, dtype=np.int64), pts3), None)]
    for name, tables, nelts in cases:
        res = []
        for asm in (0, 1):
            ddf._lib.call("ddf_set_tuning", b"asm", asm)
            try:
                if tables is None:
                    m = ddf.large_hypercube_manifold(len(nelts), nelts, ctx=ctx)
                else:
                    m = ddf.Manifold.from_simplices(name, tables[0], tables[1], ctx=ctx)
                D = m.D
                out = [m.get_simplices(R) for R in range(1, D + 1)] + [a for R in range(1, D + 1)
                                                                       for a in m.get_boundaries_csc(R)]
                rngx = np.random.default_rng(5)
                x = rngx.standard_normal(m.nsimplices(D))
                out.append((ddf.boundary(ddf.Pr, D, m) * ddf.Fun(m, ddf.Pr, D, x)).values)   # exercises the row-CSR
                res.append(out)
                m.close()
            finally:
                ddf._lib.call("ddf_set_tuning", b"asm", 0)
        assert len(res[0]) == len(res[1])
        for a, b in zip(res[0], res[1]):
            assert np.array_equal(a, b), name


def test_assembly_huge_bucket(ddf, ctx):
    """A fan of 300 triangles around vertex 0: 600 edge tuples share the smallest vertex, far more than a warp ranks
    in shared memory (BK_MAX = 128) -- the serial in-place path of k_bucket_sort; plus the same fan in 3D (cone of
    tets over a triangulated disc)."""
    M = 300
    ang = 2 * np.pi * np.arange(M) / M
    pts = np.concatenate([[[0.0, 0.0]], np.stack([np.cos(ang), np.sin(ang)], axis=1)])
    tris = np.array([[0, 1 + i, 1 + (i + 1) % M] for i in range(M)], dtype=np.int64)
    pts3 = np.concatenate([np.concatenate([pts, np.zeros((M + 1, 1))], axis=1), [[0.0, 0.0, 1.0]]])
    tets = np.array([[0, 1 + i, 1 + (i + 1) % M, M + 1] for i in range(M)], dtype=np.int64)
    for simp, coords in ((tris, pts), (tets, pts3)):
        om = o.build_manifold("fan", simp, coords)
        dm = ddf.Manifold.from_simplices("fan", simp, coords, ctx=ctx)
        D = om.D
        for R in range(1, D + 1):
            cp, rv, nz = dm.get_boundaries_csc(R)
            ocp, orv, onz = o.to_julia_csc(om.boundaries[R])
            assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz), R
            assert np.array_equal(dm.get_simplices(R), om.simplices[R]), R
        rng = np.random.default_rng(3)
        for k in range(D):
            x = rng.standard_normal(om.nsimplices(k))
            assert np.array_equal((ddf.deriv(ddf.Pr, k, dm) * ddf.Fun(dm, ddf.Pr, k, x)).values,
                                  o.apply_deriv(o.Pr, k, om, x))
            z = rng.standard_normal(om.nsimplices(k + 1))
            assert np.array_equal((ddf.boundary(ddf.Pr, k + 1, dm) * ddf.Fun(dm, ddf.Pr, k + 1, z)).values,
                                  om.boundaries[k + 1].astype(np.float64) @ z)

"""Exact restatement of the geometry behind the Hodge star (test infrastructure), written from the reference's formulas
and sharing no code with oracle/ddf_oracle.py:

  volume                 src/Algorithms.jl:102-118   vol = |wedge(x_k - x_1)| / R!        (here: sqrt(Gram determinant) / R!)
  weighted circumcentre  src/Algorithms.jl:56-96     [2 x_i.x_j 1; 1' 0] [c; .] = [x_i.x_i - w_i; 1],  cc = sum c_i x_i
  dual volumes           src/Manifolds.jl:1160-1223  dualvol_R[i] = sum_j dualvol_{R+1}[j] * h_ij / (D - R),
                         h_ij = s (cc_j - cc_i).n,  n = unit normal of i inside j,  s = sign((bc_j - x_i1).n)
  hodge                  src/ManifoldOps.jl:58-68    star_R = dualvol_R / vol_R
  laplace(Pr,0)          src/ManifoldOps.jl:110-144  L = -star_0^-1 d_0' star_1 d_0

Arithmetic: `fractions.Fraction` wherever the quantity is rational (Gram matrices, circumcentres, squared volumes,
squared heights -- the reference's own tests use Rational{BigInt} the same way, test/test-algorithms.jl) and 50-digit
`mpmath` for the square roots and the sums of square roots; inputs are the Float64 coordinates / weights taken as the
exact rationals they are.  Topology (vertex tables, cofaces) comes from the caller as integer tables.
"""
from fractions import Fraction as F
from math import factorial

import mpmath as mp

mp.mp.dps = 50


def _dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def _sub(a, b):
    return [x - y for x, y in zip(a, b)]


def solve_exact(A, b):
    """Gaussian elimination over the rationals (pivot = first non-zero)."""
    n = len(b)
    M = [list(map(F, row)) + [F(rhs)] for row, rhs in zip(A, b)]
    for c in range(n):
        p = next(r for r in range(c, n) if M[r][c] != 0)
        M[c], M[p] = M[p], M[c]
        piv = M[c][c]
        M[c] = [v / piv for v in M[c]]
        for r in range(n):
            if r != c and M[r][c] != 0:
                f = M[r][c]
                M[r] = [vr - f * vc for vr, vc in zip(M[r], M[c])]
    return [M[r][n] for r in range(n)]


def det_exact(A):
    n = len(A)
    M = [list(map(F, row)) for row in A]
    d = F(1)
    for c in range(n):
        p = next((r for r in range(c, n) if M[r][c] != 0), None)
        if p is None:
            return F(0)
        if p != c:
            M[c], M[p] = M[p], M[c]
            d = -d
        d *= M[c][c]
        for r in range(c + 1, n):
            f = M[r][c] / M[c][c]
            M[r] = [vr - f * vc for vr, vc in zip(M[r], M[c])]
    return d


def to_mp(q):
    return mp.mpf(q.numerator) / mp.mpf(q.denominator)


def volume_sq(xs):
    """vol^2 as a rational: Gram determinant / (R!)^2 (vol_0 = 1)."""
    R = len(xs) - 1
    if R == 0:
        return F(1)
    ys = [_sub(x, xs[0]) for x in xs[1:]]
    G = [[_dot(a, b) for b in ys] for a in ys]
    return det_exact(G) / F(factorial(R) ** 2)


def circumcentre(xs, ws):
    N = len(xs)
    A = [[2 * _dot(xs[i], xs[j]) for j in range(N)] + [F(1)] for i in range(N)] + [[F(1)] * N + [F(0)]]
    b = [_dot(xs[i], xs[i]) - ws[i] for i in range(N)] + [F(1)]
    c = solve_exact(A, b)
    C = len(xs[0])
    return [sum(c[i] * xs[i][d] for i in range(N)) for d in range(C)]


def height(xi, cci, xj, ccj):
    """h_ij of Manifolds.jl:1196-1212 as (sign, h^2 rational): simplex i (vertices xi) inside its coface j (vertices xj)."""
    ys = [_sub(x, xi[0]) for x in xi[1:]]
    extra = [x for x in xj if x not in xi]
    assert len(extra) == 1
    v = _sub(extra[0], xi[0])
    if ys:
        G = [[_dot(a, b) for b in ys] for a in ys]
        a = solve_exact(G, [_dot(y, v) for y in ys])
        n = [v[d] - sum(a[k] * ys[k][d] for k in range(len(ys))) for d in range(len(v))]
    else:
        n = v
    nn = _dot(n, n)
    assert nn > 0
    bc = [sum(x[d] for x in xj) / F(len(xj)) for d in range(len(v))]
    s = 1 if _dot(_sub(bc, xi[0]), n) > 0 else -1
    h0n = _dot(_sub(ccj, cci), n)              # (cc_j - cc_i).n  (times |n|)
    sign = s * (1 if h0n > 0 else (-1 if h0n < 0 else 0))
    return sign, h0n * h0n / nn


def geometry(D, simplices, coords, weights):
    """simplices: {R: list of vertex tuples (ascending)}, coords: list of float tuples, weights: list of floats.
    Returns (vol, dualvol) as {R: list of mpf}, and the circumcentres {R: list of rational vectors}."""
    X = [[F(c) for c in p] for p in coords]
    W = [F(w) for w in weights]
    vol, cc = {}, {}
    for R in range(D + 1):
        vol[R] = [mp.sqrt(to_mp(volume_sq([X[v] for v in s]))) for s in simplices[R]]
        cc[R] = [circumcentre([X[v] for v in s], [W[v] for v in s]) for s in simplices[R]]
    dualvol = {D: [mp.mpf(1)] * len(simplices[D])}
    for R in range(D - 1, -1, -1):
        index = {}
        for j, sj in enumerate(simplices[R + 1]):
            for k in range(len(sj)):
                index.setdefault(sj[:k] + sj[k + 1:], []).append(j)      # parents in ascending order
        out = []
        for i, si in enumerate(simplices[R]):
            acc = mp.mpf(0)
            for j in index.get(si, []):
                sj = simplices[R + 1][j]
                sign, h2 = height([X[v] for v in si], cc[R][i], [X[v] for v in sj], cc[R + 1][j])
                acc += dualvol[R + 1][j] * sign * mp.sqrt(to_mp(h2))
            out.append(acc / (D - R))
        dualvol[R] = out
    return vol, dualvol, cc


def laplace0(D, simplices, vol, dualvol):
    """L = laplace(Pr,0) as a dict {(i, j): mpf} from the exact star values: off-diagonal L[i,j] = (vol_0/dualvol_0)[i] *
    star_1[e] for every edge e = (i,j), diagonal = - sum of the row's off-diagonals."""
    L = {}
    for e, (a, b) in enumerate(simplices[1]):
        s1 = dualvol[1][e] / vol[1][e]
        for i, j in ((a, b), (b, a)):
            l = vol[0][i] / dualvol[0][i] * s1
            L[(i, j)] = L.get((i, j), mp.mpf(0)) + l
            L[(i, i)] = L.get((i, i), mp.mpf(0)) - l
    return L

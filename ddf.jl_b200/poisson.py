"""Poisson problem driver, mirroring `Poisson.poisson(Val(D), levels)` of examples/poisson.jl:27-262
on the device.

The reference writes the problem as the first-order system [0 delta; -d 1][u; f] = [rho; 0] and projects the
Dirichlet rows out: A' = (E-B) A + B, b' = (E-B) b + B c (poisson.jl:172-201).  The second block row is
f = d u on every edge (B has no f block), so the system is algebraically the same as

    ((E0 - B0) * laplace(Pr,0) + B0) u = (E0 - B0) rho + B0 u0,      f = d u,

which is what is solved here, with the operator composed from the drop-in Op algebra and the reference's
iterative solver (BiCGStab(2), IterativeSolvers defaults; `ddf_solve_bicgstabl`).  The reference switches
to UMFPACK `\\` for D <= 2; a sparse direct factorisation is outside the operator path (out of
scope), the Krylov solve is used for every D.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from ._lib import call
from .core import Fun, Manifold, Op, Pr
from .manifoldops import deriv, isboundary, laplace, one, refined_simplex_manifold


def bicgstabl(A: Op, b: Fun, x: Fun = None, l: int = 2, reltol: float = math.sqrt(np.finfo(np.float64).eps),
              abstol: float = 0.0, max_mv_products: int = None):
    """x = A \\ b by BiCGStab(l) on the device (IterativeSolvers.bicgstabl! defaults).  Returns
    (x, info) with info = dict(mv_products, residual, converged)."""
    if x is None:
        x = Fun(b.manifold, A.P1, A.R1)
    if max_mv_products is None:
        max_mv_products = A.shape[1]
    mv, res, conv = C.c_int(), C.c_double(), C.c_int()
    call("ddf_solve_bicgstabl", A._h, b._h, x._h, int(l), float(reltol), float(abstol), int(max_mv_products),
         C.byref(mv), C.byref(res), C.byref(conv))
    return x, {"mv_products": mv.value, "residual": res.value, "converged": bool(conv.value)}


def poisson_jacobi(mfd: Manifold, u: Fun, rho: Fun, theta: float = 2.0 / 3.0, nsweeps: int = 1) -> Fun:
    """In-place fused relaxation sweeps on L u = rho with the boundary values of u held fixed:
    u <- u + theta * ((1-b) .* ((rho - L u) ./ diag(L)))."""
    mfd._geometry()
    fn = "ddf_poisson_jacobi_dist" if mfd.ctx.nranks > 1 else "ddf_poisson_jacobi"
    call(fn, mfd._h, u._h, rho._h, float(theta), int(nsweeps))
    return u


def rho_exact(coords: np.ndarray) -> np.ndarray:
    """poisson.jl:85-92: Gaussian of width sigma = 0.2 centred at x0 = 0.1*1."""
    D = coords.shape[1]
    sigma = 0.2
    r = np.sqrt(np.sum((coords - 0.1) ** 2, axis=1))
    rp = r / (math.sqrt(2.0) * sigma)
    return 1.0 / math.sqrt(2 * math.pi * sigma ** 2) ** D * np.exp(-rp ** 2)


def u_exact(coords: np.ndarray) -> np.ndarray:
    """poisson.jl:93-122: the free-space potential of rho_exact (D = 1, 2, 3)."""
    from scipy.special import erf, expi
    D = coords.shape[1]
    sigma = 0.2
    r = np.sqrt(np.sum((coords - 0.1) ** 2, axis=1))
    rp = r / (math.sqrt(2.0) * sigma)
    eps = np.finfo(np.float64).eps
    if D == 1:
        return sigma / math.sqrt(2 * math.pi) * (np.exp(-rp ** 2) - 1) + r / 2 * erf(rp)
    if D == 2:
        small = r ** 8 <= eps
        rs = np.where(small, 1.0, rp)
        big = -1 / (4 * math.pi) * (-np.euler_gamma + expi(-rs ** 2) + np.log(1 / rs ** 2))
        ser = 1 / (4 * math.pi) * rp ** 2 - 1 / (16 * math.pi) * rp ** 4 + 1 / (72 * math.pi) * rp ** 6
        return np.where(small, ser, big)
    if D == 3:
        rs = np.where(r == 0, 1.0, r)
        return np.where(r == 0, -1 / (4 * math.pi) * 2 / (math.sqrt(2 * math.pi) * sigma),
                        -1 / (4 * math.pi * rs) * erf(rp))
    raise ValueError("u_exact: D <= 3")


def poisson_operator(mfd: Manifold) -> Op:
    """A' restricted to u: (E0 - B0) * laplace(Pr,0) + B0  (poisson.jl:172-200 with f eliminated)."""
    B0 = isboundary(Pr, 0, mfd)
    E0 = one(Pr, 0, mfd)
    return (E0 - B0) * laplace(Pr, 0, mfd) + B0


def poisson_solve(mfd: Manifold, rho: Fun, u0: Fun, assembled: bool = False, **solver_args):
    """Solve L u = rho in the interior, u = u0 on the boundary.  Returns (u, f = d u, info).
    `assembled=True` hands the solver the assembled, chopped A' the reference's solver sees (poisson.jl:203); the
    default applies the expression (E0-B0)*L + B0 through the fused Laplacian rows, which skips the sparse products
    of the assembly (tools/poisson_probe.py: 1 ms instead of 50-80 ms before a 30-110 ms solve)."""
    B0 = isboundary(Pr, 0, mfd)
    E0 = one(Pr, 0, mfd)
    A = poisson_operator(mfd)
    if assembled:
        A = A.assemble()
    b = (E0 - B0) * rho + B0 * u0
    u, info = bicgstabl(A, b, **solver_args)
    f = deriv(Pr, 0, mfd) * u
    return u, f, info


def poisson(D: int, levels: int, ctx=None, mfd: Manifold = None, **solver_args):
    """`Poisson.poisson(Val(D), levels)` (poisson.jl:27-262): refined simplex mesh, sampled rho and u0, solve,
    and the diagnostics the reference prints: residual norm and the 1-, 2- and max-norm errors."""
    if mfd is None:
        mfd = refined_simplex_manifold(D, levels, ctx=ctx)
    coords = mfd.get_coords()
    rho = Fun(mfd, Pr, 0, rho_exact(coords))
    u0 = Fun(mfd, Pr, 0, u_exact(coords))
    u, f, info = poisson_solve(mfd, rho, u0, **solver_args)
    B0 = isboundary(Pr, 0, mfd)
    E0 = one(Pr, 0, mfd)
    r = (E0 - B0) * (laplace(Pr, 0, mfd) * u - rho) + B0 * (u - u0)        # poisson.jl:240
    n = len(u)
    e = u - u0
    out = {"residual2": math.sqrt(r.norm(2) ** 2 / n), "e1": e.norm(1) / n, "e2": math.sqrt(e.norm(2) ** 2 / n),
           "einf": e.norm(np.inf), **info}
    return u, f, out

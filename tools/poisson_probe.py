#!/usr/bin/env python
"""Cost of the Poisson solve (examples/poisson.jl) on one GPU: BiCGStab(2) on the assembled Dirichlet system.
Each case runs twice; the second run has warm allocations.  Prints vertices, seconds, matrix-vector products."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddf.jl_b200 as ddf  # noqa: E402
from importlib import import_module  # noqa: E402

pz = import_module("ddf.jl_b200.poisson")

ctx = ddf.Context(0)
cases = [("2D refined level 6", lambda: ddf.refined_simplex_manifold(2, 6, ctx=ctx), 2, 6),
         ("2D refined level 8", lambda: ddf.refined_simplex_manifold(2, 8, ctx=ctx), 2, 8),
         ("3D Kuhn n=64", lambda: ddf.large_hypercube_manifold(3, (64, 64, 64), ctx=ctx), 3, 0),
         ("3D Kuhn n=128", lambda: ddf.large_hypercube_manifold(3, (128, 128, 128), ctx=ctx), 3, 0)]
for name, make, D, lv in cases:
    m = make()
    coords = m.get_coords()
    rho, u0 = ddf.Fun(m, ddf.Pr, 0, pz.rho_exact(coords)), ddf.Fun(m, ddf.Pr, 0, pz.u_exact(coords))
    B0, E0 = ddf.isboundary(ddf.Pr, 0, m), ddf.one(ddf.Pr, 0, m)
    b = (E0 - B0) * rho + B0 * u0
    for label, prep in (("expression", lambda A: A), ("assembled", lambda A: A.assemble())):
        for rep in range(2):
            ctx.sync()
            t0 = time.time()
            A = prep(ddf.poisson_operator(m))
            ctx.sync()
            t1 = time.time()
            u, info = ddf.bicgstabl(A, b)
            ctx.sync()
            t2 = time.time()
        print(f"{name} [{label}]: {m.nsimplices(0)} vertices, operator {1e3 * (t1 - t0):.1f} ms, solve {1e3 * (t2 - t1):.1f} ms, "
              f"{info['mv_products']} products, {1e3 * (t2 - t1) / info['mv_products']:.3f} ms/product, "
              f"converged={info['converged']}", flush=True)
    m.close()

#!/usr/bin/env python
"""A/B of the Tsit5 step at C4 (or argv[1]): three paired-stage passes over L (default), six fused single-stage launches
(tuning bit 1 << 30) and the unfused sequence of k_lincomb passes and right-hand sides (1 << 29); bits compared."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import ddf.jl_b200 as ddf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = ddf.Context(0)
m = ddf.large_hypercube_manifold(D, (n,) * D, ctx=ctx)
dt = ddf.wave_dt(m)
rng = np.random.default_rng(3)
u0, v0 = rng.standard_normal(m.nsimplices(0)), rng.standard_normal(m.nsimplices(0))
ref = None
for name, tune in (("unfused", 1 << 29), ("six fused single-stage launches", 1 << 30), ("three paired-stage passes over L", 0)):
    ddf._lib.call("ddf_set_tuning", b"wave", tune)
    u, v = ddf.Fun(m, ddf.Pr, 0, u0), ddf.Fun(m, ddf.Pr, 0, v0)
    ddf.wave_tsit5(m, u, v, dt, 2)
    got = (u.values, v.values)
    ref = ref or got
    same = np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    ts = []
    for _ in range(3):
        ctx.flush_l2()
        ctx.sync()
        ctx.tic()
        ddf.wave_tsit5(m, u, v, dt, 5)
        ts.append(ctx.toc() / 5)
    print(f"tsit5_ab D={D} n={n} {name}: {min(ts):.3f} ms per step, same bits: {same}", flush=True)
ddf._lib.call("ddf_set_tuning", b"wave", 0)

#!/usr/bin/env python
"""A/B of the fused wave step at C4 (or R2_N): one step per launch, two steps per launch, Tsit5 step, Poisson sweep;
bits compared against the explicit-index SELL kernel (tuning 16), which is pinned to the oracle by the parity tests."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import ddf.jl_b200 as ddf  # noqa: E402

n = int(os.environ.get("R2_N", "256"))
ctx = ddf.Context(0)
rng = np.random.default_rng(0)


def run(tune, steps=6):
    ddf._lib.call("ddf_set_tuning", b"wave", tune)
    m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
    dt = ddf.wave_dt(m)
    nv = m.nsimplices(0)
    r = np.random.default_rng(1)
    u0, v0 = r.standard_normal(nv), r.standard_normal(nv)
    u, v = ddf.Fun(m, ddf.Pr, 0, u0), ddf.Fun(m, ddf.Pr, 0, v0)
    ddf.wave_leapfrog(m, u, v, dt, steps)
    out = (u.values, v.values)
    ts = []
    for _ in range(3):
        ctx.flush_l2()
        ctx.sync()
        ctx.tic()
        ddf.wave_leapfrog(m, u, v, dt, 20)
        ts.append(ctx.toc() / 20)
    tu, tv = ddf.Fun(m, ddf.Pr, 0, u0), ddf.Fun(m, ddf.Pr, 0, v0)
    ddf.wave_tsit5(m, tu, tv, dt, 1)
    ctx.sync()
    ctx.tic()
    ddf.wave_tsit5(m, tu, tv, dt, 5)
    t5 = ctx.toc() / 5
    m.close()
    ddf._lib.call("ddf_set_tuning", b"wave", 0)
    return out, min(ts), t5


ref, t_sell, _ = run(16)
for name, tune in (("one step per launch", 131072), ("two steps per launch", 0)):
    got, t, t5 = run(tune)
    same = np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    print(f"wave_ab n={n} {name}: {t:.4f} ms per step, tsit5 {t5:.3f} ms per step, same bits as SELL: {same}", flush=True)
print(f"wave_ab n={n} SELL: {t_sell:.4f} ms per step", flush=True)

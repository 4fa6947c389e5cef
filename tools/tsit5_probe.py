#!/usr/bin/env python
"""Where a fixed-step Tsit5 wave step spends its time (3D Kuhn mesh, default n=256): the whole step, one right-hand
side, and the stage updates alone."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddf.jl_b200 as ddf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = ddf.Context(0)
m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
u, v = ddf.sample(m, "wave"), ddf.Fun(m, ddf.Pr, 0)
dt = ddf.wave_dt(m)
ddf.wave_tsit5(m, u, v, dt, 1)
ctx.sync()
ctx.tic()
ddf.wave_tsit5(m, u, v, dt, steps)
t = ctx.toc() / steps
print(f"tsit5 step: {t:.3f} ms", flush=True)
du, dv = ddf.Fun(m, ddf.Pr, 0), ddf.Fun(m, ddf.Pr, 0)
rhs = lambda: ddf._lib.call("ddf_wave_rhs", m._h, u._h, v._h, du._h, dv._h)  # noqa: E731
rhs()
ctx.sync()
ctx.tic()
for _ in range(10):
    rhs()
print(f"rhs: {ctx.toc() / 10:.3f} ms", flush=True)
ks = [ddf.Fun(m, ddf.Pr, 0) for _ in range(6)]
out = ddf.Fun(m, ddf.Pr, 0)
for nk in (1, 3, 6):
    ddf.lincomb(u, dt, [0.1] * nk, ks[:nk], out=out)
    ctx.sync()
    ctx.tic()
    for _ in range(10):
        ddf.lincomb(u, dt, [0.1] * nk, ks[:nk], out=out)
    tl = ctx.toc() / 10
    nb = (nk + 2) * 8 * len(u)
    print(f"lincomb nk={nk}: {tl:.3f} ms, {nb / tl * 1e-6:.0f} GB/s", flush=True)

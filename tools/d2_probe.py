#!/usr/bin/env python
"""Diagnostic: time y = d2 x after each stage of bench.py's setup (which stage changes it?)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddf.jl_b200 as ddf  # noqa: E402

ctx = ddf.Context(0)
n = 256


def t(fn, reps=20):
    fn()
    ctx.sync()
    ctx.tic()
    for _ in range(reps):
        fn()
    return round(ctx.toc() / reps, 4)


mfd = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx, assemble=False)
ddf._lib.call("ddf_mesh_assemble", mfd._h)
d2 = ddf.deriv(ddf.Pr, 2, mfd)
x2 = ddf.rand_fun(mfd, ddf.Pr, 2, seed=3)
y2 = ddf.Fun(mfd, ddf.Pr, 3)
print("after assemble:", t(lambda: d2.apply(x2, y2)), flush=True)
mfd._geometry()
print("after geometry:", t(lambda: d2.apply(x2, y2)), flush=True)
d = [ddf.deriv(ddf.Pr, k, mfd) for k in range(3)]
h = [ddf.hodge(ddf.Pr, k, mfd) for k in range(4)]
x = [ddf.rand_fun(mfd, ddf.Pr, k, seed=k + 1) for k in range(4)]
y = [ddf.Fun(mfd, ddf.Pr, k + 1) for k in range(3)]
z = [ddf.Fun(mfd, ddf.Dl, k) for k in range(4)]
print("after vectors (old x2,y2):", t(lambda: d2.apply(x2, y2)), " (bench's x[2],y[2]):", t(lambda: d[2].apply(x[2], y[2])), flush=True)
for _ in range(25):
    for k in range(3):
        d[k].apply(x[k], y[k])
    for k in range(4):
        h[k].apply(x[k], z[k])
print("after steps:", t(lambda: d2.apply(x2, y2)), t(lambda: d[2].apply(x[2], y[2])), flush=True)
print("d0,d1:", t(lambda: d[0].apply(x[0], y[0])), t(lambda: d[1].apply(x[1], y[1])), flush=True)


def hot():
    for _ in range(10):
        for k in range(3):
            d[k].apply(x[k], y[k])
        for k in range(4):
            h[k].apply(x[k], z[k])


# sustained-load A/B of the rows-per-lane variants (21: d1 RPW 8 / d2 RPW 4, 22: RPW 2) and of the prefetch distance
for var in (0, 21, 22, 0):
    ddf._lib.call("ddf_set_tuning", b"fixed", var)
    hot()
    print(f"sustained fixed={var}: d1 {t(lambda: d[1].apply(x[1], y[1]))} d2 {t(lambda: d[2].apply(x[2], y[2]))}", flush=True)
ddf._lib.call("ddf_set_tuning", b"fixed", 0)
for pf in (0, 148, 592, 2368):
    ddf._lib.call("ddf_set_tuning", b"prefetch", pf)
    hot()
    print(f"sustained prefetch={pf}: d0 {t(lambda: d[0].apply(x[0], y[0]))} d1 {t(lambda: d[1].apply(x[1], y[1]))} "
          f"d2 {t(lambda: d[2].apply(x[2], y[2]))}", flush=True)
ddf._lib.call("ddf_set_tuning", b"prefetch", 592)

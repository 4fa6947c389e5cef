#!/usr/bin/env python
"""Boundary / dual-derivative applies (rows of d_R^T type) at the BASELINE config-4 size: upper-run format (default)
against the plain SELL copy (tuning 20) and the row-CSR (tuning 9).  Each format needs its own mesh (the choice is
made at the first apply)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import ddf.jl_b200 as ddf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = 20
ctx = ddf.Context(0)
ref = None
for tune, name in ((20, "sell"), (21, "upper runs, global loads"), (0, "upper runs, TMA-staged"), (9, "csr")):
    ddf._lib.call("ddf_set_tuning", b"fixed", tune)
    m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
    cnt = [m.nsimplices(k) for k in range(4)]
    row, vals = {}, []
    for k in (1, 2, 3):
        A = ddf.deriv(ddf.Dl, k, m)
        x = ddf.rand_fun(m, ddf.Dl, k, seed=k + 7)
        y = ddf.Fun(m, ddf.Dl, k - 1)
        A.apply(x, y)
        ctx.sync()
        ctx.tic()
        for _ in range(reps):
            A.apply(x, y)
        ms = ctx.toc() / reps
        b = (k + 1) * cnt[k] * 4 + cnt[k - 1] * 12 + cnt[k] * 8
        row[f"dual_d{k}"] = (round(ms, 4), round(b / ms * 1e-6 / 6539.9, 3))
        vals.append(y.values[::7].copy())
    if ref is None:
        ref = vals
    else:
        assert all(np.array_equal(a, b_) for a, b_ in zip(ref, vals)), name
    print(f"{name}: {row}", flush=True)
    m.close()
ddf._lib.call("ddf_set_tuning", b"fixed", 0)

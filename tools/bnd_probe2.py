#!/usr/bin/env python
"""Boundary / dual-derivative apply on the rows of ~5 entries (rows = edges of a 3D mesh): A/B of the launch variants
that are chosen per apply (ddf_set_tuning("fixed", v)) on ONE mesh.  Usage: python tools/bnd_probe2.py [n] [v,v,...] [prefetch,prefetch,...] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import ddf.jl_b200 as ddf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 28, 0]
prefetch = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [592]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
ctx = ddf.Context(0)
m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
cnt = [m.nsimplices(k) for k in range(4)]
ref = {}
for tune, pf in [(t, p) for t in variants for p in prefetch]:
    ddf._lib.call("ddf_set_tuning", b"fixed", tune)
    ddf._lib.call("ddf_set_tuning", b"prefetch", pf)
    row = {}
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
        v = y.values[::7].copy()
        if k in ref:
            assert np.array_equal(ref[k], v), (tune, k)
        else:
            ref[k] = v
    print(f"fixed={tune} prefetch={pf}: {row}", flush=True)
ddf._lib.call("ddf_set_tuning", b"fixed", 0)
ddf._lib.call("ddf_set_tuning", b"prefetch", 592)      # the library default

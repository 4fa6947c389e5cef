#!/usr/bin/env python
"""A/B timing of kernel variants (ddf_set_tuning) at the BASELINE config-4 size.
Usage (on a B200): python tools/tune.py [--n 256] [--reps 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import ddf.jl_b200 as ddf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--fixed", default="0,1,2,3,4")
    ap.add_argument("--wave", default="0,1")
    args = ap.parse_args()
    ctx = ddf.Context(0)
    n = args.n
    m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
    m._geometry()
    cnt = [m.nsimplices(k) for k in range(4)]
    peak = 6539.9
    d = [ddf.deriv(ddf.Pr, k, m) for k in range(3)]
    x = [ddf.rand_fun(m, ddf.Pr, k, seed=k + 1) for k in range(3)]
    y = [ddf.Fun(m, ddf.Pr, k + 1) for k in range(3)]
    ref = None
    out = {}

    def timeit(fn):
        fn()
        ctx.sync()
        ctx.tic()
        for _ in range(args.reps):
            fn()
        return ctx.toc() / args.reps

    for var in [int(v) for v in args.fixed.split(",")]:
        ddf._lib.call("ddf_set_tuning", b"fixed", var)
        row = {}
        for k in range(3):
            ms = timeit(lambda k=k: d[k].apply(x[k], y[k]))
            b = cnt[k + 1] * (4 * (k + 2) + 8) + 8 * cnt[k]
            row[f"d{k}"] = (round(ms, 4), round(b / ms * 1e-6 / peak, 4))
        vals = [yy.values[:: max(1, len(yy) // 100000)].copy() for yy in y]
        if ref is None:
            ref = vals
        else:
            for a, b_ in zip(ref, vals):
                assert np.array_equal(a, b_), f"variant {var} differs"
        out[f"fixed{var}"] = row
        print(f"fixed variant {var}: {row}", flush=True)
    ddf._lib.call("ddf_set_tuning", b"fixed", 0)

    u0 = ddf.sample(m, "wave")
    dt = ddf.wave_dt(m)
    alg, fmt = ddf.wave_step_bytes(m)
    refu = None
    for var in [int(v) for v in args.wave.split(",")]:
        ddf._lib.call("ddf_set_tuning", b"wave", var)
        u, v = u0.copy(), ddf.Fun(m, ddf.Pr, 0)
        ddf.wave_leapfrog(m, u, v, dt, 3)
        ctx.sync()
        ctx.tic()
        ddf.wave_leapfrog(m, u, v, dt, args.reps)
        ms = ctx.toc() / args.reps
        uv = u.values[::97].copy()
        if refu is None:
            refu = uv
        else:
            assert np.array_equal(refu, uv), f"wave variant {var} differs"
        out[f"wave{var}"] = (round(ms, 4), round(alg / ms * 1e-6 / peak, 4), round(fmt / ms * 1e-6 / peak, 4))
        print(f"wave variant {var}: ms={ms:.4f} frac_alg={alg / ms * 1e-6 / peak:.4f} frac_fmt={fmt / ms * 1e-6 / peak:.4f}", flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

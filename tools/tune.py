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
    ap.add_argument("--torch", action="store_true", help="initialise torch's CUDA state first, like bench.py")
    ap.add_argument("--rebuild", action="store_true", help="build the mesh twice (the first one is closed), like bench.py")
    ap.add_argument("--extra", action="store_true", help="allocate bench.py's other work vectors too")
    ap.add_argument("--prefetch", default="", help="L2 prefetch distances (blocks) to time the d_k applies with")
    args = ap.parse_args()
    if args.torch:
        import torch
        torch.cuda.set_device(0)
        keep = torch.zeros(1 << 20, device="cuda")      # noqa: F841
    ctx = ddf.Context(0)
    n = args.n
    m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
    if args.rebuild:
        m.close()
        m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
    m._geometry()
    cnt = [m.nsimplices(k) for k in range(4)]
    extra = [ddf.hodge(ddf.Pr, k, m) for k in range(4)] if args.extra else None
    peak = 6539.9
    d = [ddf.deriv(ddf.Pr, k, m) for k in range(3)]
    x = [ddf.rand_fun(m, ddf.Pr, k, seed=k + 1) for k in range(3)]
    if args.extra:
        x.append(ddf.rand_fun(m, ddf.Pr, 3, seed=4))
    y = [ddf.Fun(m, ddf.Pr, k + 1) for k in range(3)]
    if args.extra:
        extra += [ddf.Fun(m, ddf.Dl, k) for k in range(4)]
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
    for pf in [int(v) for v in args.prefetch.split(",") if v]:
        ddf._lib.call("ddf_set_tuning", b"prefetch", pf)
        row = {}
        for k in range(3):
            ms = timeit(lambda k=k: d[k].apply(x[k], y[k]))
            b = cnt[k + 1] * (4 * (k + 2) + 8) + 8 * cnt[k]
            row[f"d{k}"] = (round(ms, 4), round(b / ms * 1e-6 / peak, 4))
        for a, b_ in zip(ref, [yy.values[:: max(1, len(yy) // 100000)].copy() for yy in y]):
            assert np.array_equal(a, b_), f"prefetch {pf} differs"
        print(f"prefetch {pf}: {row}", flush=True)
    ddf._lib.call("ddf_set_tuning", b"prefetch", 592)      # the library default

    # boundary / dual derivative (row-CSR vs SELL-32): y = d_k^T-type applies
    dd = [ddf.deriv(ddf.Dl, k, m) for k in (1, 2, 3)]
    xd = [ddf.rand_fun(m, ddf.Dl, k, seed=k + 7) for k in (1, 2, 3)]
    yd = [ddf.Fun(m, ddf.Dl, k - 1) for k in (1, 2, 3)]
    refd = None
    for var, name in ((0, "sell"), (9, "csr")):
        ddf._lib.call("ddf_set_tuning", b"fixed", var)
        row = {}
        for n_, k in enumerate((1, 2, 3)):
            ms = timeit(lambda n_=n_: dd[n_].apply(xd[n_], yd[n_]))
            b = (k + 1) * cnt[k] * 4 + cnt[k - 1] * 12 + cnt[k] * 8
            row[f"dual_d{k}"] = (round(ms, 4), round(b / ms * 1e-6 / peak, 4))
        vals = [yy.values[::997].copy() for yy in yd]
        if refd is None:
            refd = vals
        else:
            for a, b_ in zip(refd, vals):
                assert np.array_equal(a, b_), "sell != csr"
        out[f"dual_{name}"] = row
        print(f"dual deriv {name}: {row}", flush=True)
    ddf._lib.call("ddf_set_tuning", b"fixed", 0)

    u0 = ddf.sample(m, "wave")
    dt = ddf.wave_dt(m)
    alg, fmt = ddf.wave_step_bytes(m)
    refu = None
    for var in [int(v) for v in args.wave.split(",")]:
        ddf._lib.call("ddf_set_tuning", b"wave", var)
        if var & 12 or var == 0:     # build-time flags: fresh mesh so that the L rows are rebuilt
            m2 = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
            m2._geometry()
        else:
            m2 = m
        u, v = ddf.sample(m2, "wave"), ddf.Fun(m2, ddf.Pr, 0)
        alg, fmt = ddf.wave_step_bytes(m2)
        ddf.wave_leapfrog(m2, u, v, dt, 3)
        alg, fmt = ddf.wave_step_bytes(m2)
        ctx.sync()
        ctx.tic()
        ddf.wave_leapfrog(m2, u, v, dt, args.reps)
        ms = ctx.toc() / args.reps
        uv = u.values[::97].copy()
        if refu is None:
            refu = uv
        else:
            assert np.array_equal(refu, uv), f"wave variant {var} differs"
        out[f"wave{var}"] = (round(ms, 4), round(alg / ms * 1e-6 / peak, 4), round(fmt / ms * 1e-6 / peak, 4))
        if m2 is not m:
            m2.close()
        print(f"wave variant {var}: ms={ms:.4f} fmtGB={fmt/1e9:.3f} frac_alg={alg / ms * 1e-6 / peak:.4f} frac_fmt={fmt / ms * 1e-6 / peak:.4f}", flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

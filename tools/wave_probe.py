#!/usr/bin/env python
"""Fused leapfrog step: sweep of the producer's L2 prefetch distance (ddf_set_tuning("wavepf", p)) on the 3D Kuhn
n=256 mesh (BASELINE config 4 size) and the 2D Kuhn n=2048 mesh (config 3).  Results must be bit-identical.
Usage: python tools/wave_probe.py [p,p,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import ddf.jl_b200 as ddf  # noqa: E402

pfs = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 2, 3, 4, 5, 6, 7, 0]
ctx = ddf.Context(0)
for D, n, reps in ((3, 256, 100), (2, 2048, 400)):
    m = ddf.large_hypercube_manifold(D, (n,) * D, ctx=ctx)
    m._geometry()
    dt = ddf.wave_dt(m)
    ref = None
    for pf in pfs:
        ddf._lib.call("ddf_set_tuning", b"wavepf", pf)
        u, v = ddf.sample(m, "wave"), ddf.Fun(m, ddf.Pr, 0)
        ddf.wave_leapfrog(m, u, v, dt, 5)
        alg, fmt = ddf.wave_step_bytes(m)
        ctx.sync()
        ctx.tic()
        ddf.wave_leapfrog(m, u, v, dt, reps)
        ms = ctx.toc() / reps
        uv = u.values[::31].copy()
        if ref is None:
            ref = uv
        assert np.array_equal(ref, uv), (D, pf)
        print(f"D={D} n={n} wavepf={pf}: {ms:.4f} ms/step  frac_alg={alg / ms * 1e-6 / 6539.9:.3f} "
              f"fmt {fmt / ms * 1e-6:.0f} GB/s", flush=True)
    m.close()
ddf._lib.call("ddf_set_tuning", b"wavepf", 0)

# diagonal apply (star_k): sweep of the L2 prefetch distance of its x and diagonal streams
m = ddf.large_hypercube_manifold(3, (256, 256, 256), ctx=ctx)
m._geometry()
for pf in (0, 148, 592, 1184, 2368, 0):
    ddf._lib.call("ddf_set_tuning", b"diagpf", pf)
    row = []
    for k in range(4):
        h = ddf.hodge(ddf.Pr, k, m)
        x = ddf.rand_fun(m, ddf.Pr, k, seed=k)
        z = ddf.Fun(m, ddf.Dl, k)
        h.apply(x, z)
        ctx.sync()
        ctx.tic()
        for _ in range(20):
            h.apply(x, z)
        ms = ctx.toc() / 20
        row.append((round(ms, 4), round(24 * m.nsimplices(k) / ms * 1e-6)))
    print(f"diagpf={pf}: star0..3 (ms, GB/s) {row}", flush=True)
ddf._lib.call("ddf_set_tuning", b"diagpf", 0)

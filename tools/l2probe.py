#!/usr/bin/env python
"""How fast is the fused wave step when its operands are L2-resident?  (upper bound for temporal blocking)
Usage (on a B200): python tools/l2probe.py [--ns 32,48,64,96,128,192,256]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddf.jl_b200 as ddf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ns", default="32,48,64,96,128,192,256")
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--dim", type=int, default=3, help="mesh dimension (2: BASELINE config 3 at --ns 2048)")
    ap.add_argument("--tunes", default="0", help="comma list of ddf_set_tuning('wave', x) values to A/B")
    args = ap.parse_args()
    ctx = ddf.Context(0)
    for n, tune in [(int(s), int(t)) for s in args.ns.split(",") for t in args.tunes.split(",")]:
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        m = ddf.large_hypercube_manifold(args.dim, (n,) * args.dim, ctx=ctx)
        u, v = ddf.sample(m, "wave"), ddf.Fun(m, ddf.Pr, 0)
        dt = ddf.wave_dt(m)
        ddf.wave_leapfrog(m, u, v, dt, 5)
        alg, fmt = ddf.wave_step_bytes(m)
        ctx.sync()
        ctx.tic()
        ddf.wave_leapfrog(m, u, v, dt, args.reps)
        ms = ctx.toc() / args.reps
        V = m.nsimplices(0)
        peak = 6539.9
        print(f"alg={alg/1e6:.1f}MB frac_alg={alg / ms * 1e-6 / peak:.3f} ", end="")
        print(f"tune={tune} {ddf.wave_format(m)} n={n} V={V} fmt={fmt/1e6:.1f}MB ms={ms:.4f} ns/vertex={ms*1e6/V:.4f} fmtGB/s={fmt/ms*1e-6:.0f}", flush=True)
        m.close()


if __name__ == "__main__":
    main()

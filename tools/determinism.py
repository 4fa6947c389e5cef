#!/usr/bin/env python
"""Run the fused wave step repeatedly from identical data and compare bit for bit, per row-kernel variant.
Usage (on a B200): python tools/determinism.py [--n 256] [--steps 4] [--reps 6] [--tunes 0,128,16]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import ddf.jl_b200 as ddf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--tunes", default="0,128,16")
    args = ap.parse_args()
    ctx = ddf.Context(0)
    ref = None
    for tune in [int(t) for t in args.tunes.split(",")]:
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        m = ddf.large_hypercube_manifold(3, (args.n,) * 3, ctx=ctx)
        u0 = ddf.sample(m, "wave")
        dt = ddf.wave_dt(m)
        outs = []
        for r in range(args.reps):
            u, v = u0.copy(), ddf.Fun(m, ddf.Pr, 0)
            ddf.wave_leapfrog(m, u, v, dt, args.steps)
            outs.append((u.values, v.values))
        same = [np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) for o in outs]
        bad = [int(np.sum(outs[0][0] != o[0])) for o in outs]
        if ref is None:
            ref = outs[0]
        cross = np.array_equal(ref[0], outs[0][0]) and np.array_equal(ref[1], outs[0][1])
        print(f"tune={tune} format={ddf.wave_format(m)} repeatable={same} differing_u={bad} equals_first_variant={cross}", flush=True)
        if not all(same):
            k = same.index(False)
            idx = np.nonzero(outs[0][0] != outs[k][0])[0]
            print("   first differing rows:", idx[:10], "windows:", np.unique(idx // 256)[:10], "count windows", len(np.unique(idx // 256)))
        m.close()


if __name__ == "__main__":
    main()

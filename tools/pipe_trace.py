#!/usr/bin/env python
"""Where does the time of the persistent L-rows pipeline go?  Needs a library built with -DDDF_PIPE_TRACE
(block 0 records %globaltimer when the producer arms a stage, when a consumer group sees the data, and when it
releases the stage).  Usage: python tools/pipe_trace.py <trace_lib.so> [--n 256] [--dim 3]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--dim", type=int, default=3)
    args = ap.parse_args()
    import ddf.jl_b200._lib as L
    L.LIB_PATH = os.path.abspath(args.lib)
    import ddf.jl_b200 as ddf
    ctx = ddf.Context(0)
    m = ddf.large_hypercube_manifold(args.dim, (args.n,) * args.dim, ctx=ctx)
    u, v = ddf.sample(m, "wave"), ddf.Fun(m, ddf.Pr, 0)
    dt = ddf.wave_dt(m)
    ddf.wave_leapfrog(m, u, v, dt, 5)
    ctx.sync()
    lib = C.CDLL(L.LIB_PATH)
    buf = (C.c_ulonglong * (5 * 1024))()
    assert lib.ddf_debug_pipe_trace(buf) == 0
    t = np.array(buf, dtype=np.int64).reshape(5, 1024)
    n = int(np.sum(t[0] > 0))
    n = min(n, int(np.sum(t[2] > 0)))
    issue, ready, done = t[0][:n], t[1][:n], t[2][:n]
    t0 = issue[0]
    print(f"{n} groups of windows traced on block 0; total {(done[-1] - t0) / 1e3:.1f} us -> {(done[-1] - t0) / n:.0f} ns per group")
    fill = ready - issue
    cons = done - ready
    print(f"fill (armed -> data seen by the consumer): median {np.median(fill):.0f} ns, p10 {np.percentile(fill, 10):.0f}, p90 {np.percentile(fill, 90):.0f}")
    print(f"consume (data seen -> stage released):      median {np.median(cons):.0f} ns, p10 {np.percentile(cons, 10):.0f}, p90 {np.percentile(cons, 90):.0f}")
    offs, rows = t[3][:n], t[4][:n]
    print(f"  of which: slice offsets {np.median(offs - ready):.0f} ns, row sums (warp with the longest rows) "
          f"{np.median(rows - offs):.0f} ns, epilogue + stores {np.median(done - rows):.0f} ns")
    gap = np.diff(issue)
    print(f"producer arm interval:                      median {np.median(gap):.0f} ns, p10 {np.percentile(gap, 10):.0f}, p90 {np.percentile(gap, 90):.0f}")
    for S in (3,):
        w = issue[S:] - done[:-S]
        print(f"arm[it] - release[it-{S}] (how long a free stage waited for the producer; <0 would be a bug): median {np.median(w):.0f} ns, p10 {np.percentile(w, 10):.0f}")
    inflight = [(np.sum((issue <= x) & (ready > x))) for x in issue[10:-10]]
    print(f"stages in flight when the producer arms one: mean {np.mean(inflight):.2f}")
    for k in range(8, 16):
        print(f"  it={k}: armed +{issue[k]-t0} ready +{ready[k]-t0} released +{done[k]-t0}")


if __name__ == "__main__":
    main()

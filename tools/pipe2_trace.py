#!/usr/bin/env python
"""Time line of the two-step wavefront pipeline (k_rows_pipe2) on block 0; needs a library built with -DDDF_PIPE_TRACE
(tools/build_trace.sh).  Stamps per task p: 7 producer reaches the task, 5 producer starts waiting for a free stage,
0 stage armed, 6 bulk requests issued, 3 consumer group reaches the task, 1 data visible, 2 stage released.
Usage: python tools/pipe2_trace.py <trace_lib.so> [n] [slack]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import ddf.jl_b200._lib as L  # noqa: E402

L.LIB_PATH = os.path.abspath(sys.argv[1])
import ddf.jl_b200 as ddf  # noqa: E402

n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
slack = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = ddf.Context(0)
m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
u, v = ddf.sample(m, "wave"), ddf.Fun(m, ddf.Pr, 0)
dt = ddf.wave_dt(m)
ddf._lib.call("ddf_set_tuning", b"wave", int(os.environ.get("W2_TUNE", str(65536 | ((slack + 1) << 18)))))
ddf.wave_leapfrog(m, u, v, dt, 4)
ctx.sync()
lib = C.CDLL(L.LIB_PATH)
buf = (C.c_ulonglong * (8 * 1024))()
assert lib.ddf_debug_pipe_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(8, 1024)
nt = int(min(np.sum(t[0] > 0), np.sum(t[2] > 0), 900))
lo, hi = 100, nt - 20           # steady state: both kinds alternate
reach, wait_free, armed, issued, creach, ready, done = (t[k][:nt] for k in (7, 5, 0, 6, 3, 1, 2))
print(f"{nt} tasks traced; steady-state task interval {(armed[hi] - armed[lo]) / (hi - lo):.0f} ns")
for kind, name in ((0, "even tasks"), (1, "odd tasks")):
    sel = np.arange(lo, hi)[(np.arange(lo, hi) % 2) == kind]
    def med(x):
        return f"{np.median(x):.0f} (p10 {np.percentile(x, 10):.0f}, p90 {np.percentile(x, 90):.0f})"
    print(name)
    print("  producer: dependency wait", med(wait_free[sel] - reach[sel]), "| free-stage wait", med(armed[sel] - wait_free[sel]),
          "| issue", med(issued[sel] - armed[sel]))
    print("  fill (armed -> data visible)", med(ready[sel] - armed[sel]), "| consumer waited for data", med(ready[sel] - creach[sel]))
    print("  consume", med(done[sel] - ready[sel]), "| consumer idle before the task", med(creach[sel][1:] - done[sel - 2][1:]))
if os.environ.get("W2_VGTRACE"):
    ld0, ld1 = t[4][:nt], t[5][:nt]
    sel = np.arange(lo, hi)
    print("VG consumer: consume (data visible -> released)", np.median(done[sel] - ready[sel]), "| released -> loads start",
          np.median(ld0[sel] - done[sel]), "| issuing the loads", np.median(ld1[sel] - ld0[sel]))
for k in range(lo, lo + 8):
    b = reach[lo]
    print(f"  task {k}: reach +{reach[k]-b} waitfree +{wait_free[k]-b} armed +{armed[k]-b} issued +{issued[k]-b} | creach +{creach[k]-b} ready +{ready[k]-b} done +{done[k]-b}")

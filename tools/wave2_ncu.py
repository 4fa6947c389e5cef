#!/usr/bin/env python
"""Driver for an ncu capture of the fused wave step: runs a few steps with the one-step kernel (k_rows_pipe) and with
the two-step wavefront (k_rows_pipe2) at 3D Kuhn n (default 256).  Usage under ncu:
  ncu --set full --clock-control none -k regex:k_rows_pipe -c 4 -o gpurun_out/r02_wave python tools/wave2_ncu.py 256 2"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddf.jl_b200 as ddf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
slack = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = ddf.Context(0)
m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
u, v = ddf.sample(m, "wave"), ddf.Fun(m, ddf.Pr, 0)
dt = ddf.wave_dt(m)
ddf._lib.call("ddf_set_tuning", b"wave", 131072)
ddf.wave_leapfrog(m, u, v, dt, 2)
ctx.sync()
ddf._lib.call("ddf_set_tuning", b"wave", 65536 | ((slack + 1) << 18))
ddf.wave_leapfrog(m, u, v, dt, 4)
ctx.sync()
print("done")

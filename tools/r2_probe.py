#!/usr/bin/env python
"""Round-2 probes on a B200 (one call, several small measurements; results to stdout and gpurun_out/r2_probe.json):
  geom   relative differences of the device geometry against the oracle per rank, device arrays saved for offline study
  asm    wall time of create + ddf_mesh_assemble at C4, first and second call (DDF_DEBUG breakdown on stderr)
  wave   fused leapfrog step at C4: one step per launch against the two-step wavefront (same bits?)
Usage: python tools/r2_probe.py geom asm wave"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import ddf.jl_b200 as ddf  # noqa: E402

what = sys.argv[1:] or ["geom", "asm", "wave"]
out = {}
ctx = ddf.Context(0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)


def relerr(a, b):
    s = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b)) / s)


if "geom" in what:
    import ddf_oracle as o
    res = {}
    save = {}
    for name, mk in (("kuhn2n16", lambda: (o.large_hypercube_manifold(2, (16, 16)), None)),
                     ("kuhn3n8", lambda: (o.large_hypercube_manifold(3, (8,) * 3), None)),
                     ("kuhn3n32", lambda: (o.large_hypercube_manifold(3, (32,) * 3), None)),
                     ("refined2l4", lambda: (o.refined_simplex_manifold(2, 4), None)),
                     ("simplex3", lambda: (o.simplex_manifold(3), None))):
        om, _ = mk()
        dm = ddf.Manifold.from_simplices(name, om.simplices[om.D], om.coords, om.weights, ctx=ctx)
        row = {}
        for R in range(om.D + 1):
            dv, odv = dm.get_dualvolumes(R), o.calc_dualvolumes(om, R)
            v, ov = dm.get_volumes(R), o.calc_volumes(om, R)
            cc, occ = dm.get_dualcoords(R), o.calc_dualcoords(om, R)
            row[R] = {"vol": relerr(v, ov), "dualvol": relerr(dv, odv), "cc": float(np.max(np.abs(cc - occ))),
                      "dualvol_elementwise": float(np.max(np.abs(dv - odv) / np.maximum(np.abs(odv), 1e-300)))}
            save[f"{name}_dv{R}"] = dv
            save[f"{name}_cc{R}"] = cc
            save[f"{name}_v{R}"] = v
        res[name] = row
        print(name, json.dumps(row), flush=True)
        dm.close()
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "r2_geom_device.npz"), **save)
    out["geom"] = res

if "asm" in what:
    n = 256
    rows = []
    for rep in range(3):
        ctx.sync()
        t0 = time.perf_counter()
        m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx, assemble=False)
        ctx.sync()
        t1 = time.perf_counter()
        ddf._lib.call("ddf_mesh_assemble", m._h)
        ctx.sync()
        t2 = time.perf_counter()
        rows.append({"create_ms": (t1 - t0) * 1e3, "assemble_ms": (t2 - t1) * 1e3})
        print("asm", rows[-1], flush=True)
        m.close()
        ctx.sync()
    out["asm"] = rows

if "wave" in what:
    n = int(os.environ.get("R2_N", "256"))
    m = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx)
    dt = ddf.wave_dt(m)
    rng = np.random.default_rng(0)
    u0, v0 = rng.standard_normal(m.nsimplices(0)), rng.standard_normal(m.nsimplices(0))
    res = {}
    ref = None
    cases = [("single", 131072)] + [(f"two_step_slack{sl}_np2_pf{pf}", 65536 | ((sl + 1) << 18) | (2 << 22) | ((pf + 1) << 25))
                                    for pf in (0, 2, 3, 5) for sl in (3, 4)]
    for name, tune in cases:
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        u, v = ddf.Fun(m, ddf.Pr, 0, u0), ddf.Fun(m, ddf.Pr, 0, v0)
        ddf.wave_leapfrog(m, u, v, dt, 6)
        ctx.sync()
        got = (u.values, v.values)
        if ref is None:
            ref = got
        same = bool(np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]))
        times = []
        for _ in range(5):
            ctx.flush_l2()
            ctx.sync()
            ctx.tic()
            ddf.wave_leapfrog(m, u, v, dt, 20)
            times.append(ctx.toc() / 20)
        res[name] = {"ms_per_step": times, "same_bits": same}
        print("wave", name, res[name], flush=True)
    ddf._lib.call("ddf_set_tuning", b"wave", 0)
    alg, fmt = ddf.wave_step_bytes(m)
    res["algorithmic_bytes"], res["format_bytes"] = alg, fmt
    out["wave"] = res
    m.close()

json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_probe.json"), "w"), indent=1)
print("done", flush=True)

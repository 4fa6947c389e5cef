#!/usr/bin/env python
"""bench.py -- headline benchmark of the DEC operator hot path (BASELINE.json metric:
"cochain DOF updates/sec and achieved HBM GB/s for d_k / star_k SpMV").

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (restated, see below)

A "step" = one pass of the hot path over one batch of synthetic cochains on the 3D
Kuhn mesh of BASELINE config 4 (large_hypercube n=256: 1.0066e8 tets):
    y1 = d0 x0,  y2 = d1 x1,  y3 = d2 x2,  z_k = star_k x_k  (k = 0..3)
N > 1: weak scaling -- one C4-sized z-slab per GPU (+2 ghost cell layers per interior
face); the d_k/star_k applies need no collective, the fused wave step
(reported in "wave") exchanges one ghost plane per neighbour per step over NCCL.

The reference arm: DDF.jl is Julia and cannot run here (no julia binary, no network),
so `--impl reference` times oracle/ddf_oracle.c -- a C restatement of Julia
SparseArrays' `mul!` loops on the reference's own storage (Int64 CSC + Int8/Float64
values), single-threaded like Julia's sparse mul! -- on a bounded sample (3D Kuhn
n=64) of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

print_json = None
METRIC = "cochain DOF updates/sec (d_k + star_k apply, 3D Kuhn mesh)"
UNIT = "DOF/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def dk_bytes(n_rows, n_cols, k):
    """Algorithmic bytes of y = d_k x: rows*(4(k+2)+8) + 8*cols."""
    return n_rows * (4 * (k + 2) + 8) + 8 * n_cols


def owned_counts(ddf, ctx, nelts_xy, slab, local_counts):
    """k-simplices owned by this rank = those whose smallest vertex lies in an owned plane.
    Counted from two tiny auxiliary meshes: one 2-layer slab and one in-plane 2D mesh."""
    if slab is None:
        return list(local_counts)
    two = ddf.large_hypercube_manifold(3, tuple(nelts_xy) + (2,), ctx=ctx, hdiv=nelts_xy[0])
    flat = ddf.large_hypercube_manifold(2, tuple(nelts_xy), ctx=ctx)
    c2 = [two.nsimplices(k) for k in range(4)]
    p = [flat.nsimplices(k) for k in range(3)] + [0]
    two.close()
    flat.close()

    def S(L, k):        # k-simplices of a slab with L cell layers
        return p[k] + (L // 2) * (c2[k] - p[k])
    L = slab.z_own_hi - slab.z_own_lo - (1 if slab.rank == slab.nranks - 1 else 0)
    out = []
    for k in range(4):
        own = S(L, k) - p[k]
        if slab.rank == slab.nranks - 1:
            own += p[k]
        out.append(own)
    return out


# ----------------------------------------------------------------------------- CPU arm
_CPU_SETUP = {}


def cpu_reference_setup(n_ref):
    """The reference's storage of BASELINE config 4 at n = n_ref on the host: boundaries[1..3] as 1-based Int64 CSC with
    Int8 values.  Built entirely on the CPU by the oracle's C twin (Kuhn table + linear-time face ranking, both held
    against the literal restatement in tests/test_oracle_pins.py) -- nothing of the product is involved."""
    if n_ref in _CPU_SETUP:
        return _CPU_SETUP[n_ref]
    import c_oracle
    import ddf_oracle as o
    t0 = time.time()
    tab = c_oracle.kuhn3_table(n_ref, o.symmetric_hypercube(3))
    V = (n_ref + 1) ** 3
    n = [V, 0, 0, tab.shape[0]]
    csc = {}
    for R in (3, 2, 1):
        faces, cp, rv, nz = c_oracle.calc_faces_boundaries_linear(tab, V)
        csc[R] = (cp, rv, nz)
        n[R - 1] = faces.shape[0]
        tab = faces
    del tab
    assert n[0] == V
    _CPU_SETUP[n_ref] = (n, csc, time.time() - t0)
    return _CPU_SETUP[n_ref]


def cpu_reference_pass(n_ref, warmup, steps, seconds_cap, threads_omp=False):
    """Time the C restatement of the reference's CPU path: 3D Kuhn n=n_ref, the same 7 operator applications per step
    on the reference's storage.  `warmup` untimed + up to `steps` timed passes (fewer if they would exceed
    `seconds_cap`).  Returns dict(value, unit, cores, kind, sample, ...)."""
    import c_oracle
    n, csc, t_asm = cpu_reference_setup(n_ref)
    rng = np.random.default_rng(0)
    x = [rng.random(nk) for nk in n]
    # the diagonal's values do not change the cost of Diagonal * Vector
    star = [rng.random(nk) + 0.5 for nk in n]
    dof = 2 * (n[1] + n[2] + n[3]) + n[0]

    def one_pass():
        for k in range(3):
            cp, rv, nz = csc[k + 1]
            c_oracle.spmv_adjoint_csc_i8(cp, rv, nz, x[k], omp=threads_omp)
        for k in range(4):
            c_oracle.diag_mul(star[k], x[k])

    t1 = time.time()
    one_pass()
    first = time.time() - t1
    for _ in range(max(0, min(warmup - 1, int(seconds_cap / 4 / max(first, 1e-9))))):
        one_pass()
    t1 = time.time()
    reps = 0
    while reps < steps:
        one_pass()
        reps += 1
        el = time.time() - t1
        if el >= seconds_cap:
            break
    el = time.time() - t1
    per = el / reps
    cores = os.cpu_count() if threads_omp else 1
    return {"value": dof / per, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"3D Kuhn n={n_ref} ({n[3]} tets, {dof} DOF/step), {reps} timed passes in {el:.1f}s on {cores} "
                      f"thread(s), C restatement of SparseArrays mul! on the reference's Int64-CSC/Int8 storage "
                      f"(storage built on the CPU in {t_asm:.1f}s)",
            "ms_per_step": per * 1e3, "dof_per_step": dof, "timed_passes": reps, "n": n_ref, "nsimplices": list(n)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import c_oracle
    c_oracle.load()
    # the SAME configuration as the product arm (n = args.n_ref, default = the product's n = 256): W warm-up passes and
    # K timed passes, each one whole step of the workload, capped so that the run ends within a few minutes
    single = cpu_reference_pass(args.n_ref, args.warmup, args.steps, seconds_cap=150.0)
    omp = cpu_reference_pass(args.n_ref, args.warmup, args.steps, seconds_cap=40.0, threads_omp=True)
    # Julia's SparseArrays mul! is single-threaded whatever JULIA_NUM_THREADS says: one thread IS all the
    # host threads the reference's own implementation can use, so the line's value is the 1-thread loop.  The
    # row-parallel OpenMP variant on every host core is reported beside it (not what DDF.jl does).
    res = single
    cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    cpu["allcores"] = {"value": omp["value"], "unit": UNIT, "cores": omp["cores"], "sample": omp["sample"],
                       "note": "row-parallel OpenMP restatement; the reference itself cannot thread this path"}
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.n_ref, 1, False, res["nsimplices"], res["dof_per_step"]),
            "timed_passes": res["timed_passes"],
            "reference_runnable": False,
            "note": "DDF.jl is Julia; no julia binary in the image -> C restatement of its SparseArrays mul! loops on the "
                    "reference's storage (Int64 CSC, Int8 values), 1 thread like Julia's sparse mul!",
            "cpu_baseline": cpu,
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print_json(line)


def workload_config(n, N, strong, cnt, dof_total):
    """`config` of the JSON line: identical for the product arm and the reference arm on the same n at N = 1."""
    wl = (f"BASELINE config 4: 3D Kuhn large_hypercube n={n}"
          + ((f", ONE cube cut into {N} z-slabs (config 5 strong scaling)" if strong else
              f" per GPU, z-slabs of {n * N} layers over {N} GPUs (config 5 weak scaling)") if N > 1 else "")
          + ": y=d_k x (k=0,1,2) and star_k x (k=0..3), Float64")
    return {"workload": wl, "nsimplices_local": [int(c) for c in cnt], "dof_per_step": int(dof_total),
            "l2_hygiene": "inputs larger than the last-level cache (index tables 0.9-2.4 GB, vectors 0.14-1.6 GB; "
                          "B200 L2 = 126 MB)",
            "parallelism": "none" if N == 1 else f"z-slabs x{N}, no collective in d_k/star_k apply; "
                                                 "peer-memory / NCCL halo in the wave step"}


# ----------------------------------------------------------------------------- GPU arm
def multi_gpu_parity(ddf, ctx, dist, rank, N, nxy=24, steps=5):
    """"bit-exact" when every rank's owned rows of u and v after `steps` fused leapfrog steps on the slab-decomposed
    mesh equal the undivided single-GPU run (computed on rank 0); else a description of the mismatch."""
    nelts = (nxy, nxy, 8 * N)
    mfd, slab = ddf.dist.slab_hypercube(3, nelts, ctx)
    total = (nxy + 1) ** 2 * (8 * N + 1)
    gu = np.random.default_rng(7).standard_normal(total)
    gv = np.random.default_rng(8).standard_normal(total)
    lo_g = slab.z_cell_lo * slab.plane
    nloc = slab.n_local_planes * slab.plane
    u = ddf.Fun(mfd, ddf.Pr, 0, gu[lo_g:lo_g + nloc])
    v = ddf.Fun(mfd, ddf.Pr, 0, gv[lo_g:lo_g + nloc])
    dt = ctx.allreduce(ddf.wave_dt(mfd), "min")
    ddf.wave_leapfrog(mfd, u, v, dt, 2)
    ddf.wave_leapfrog(mfd, u, v, dt, steps - 2)
    lo, hi = slab.own_rows
    parts = [None] * N
    dist.all_gather_object(parts, (slab.z_own_lo, u.values[lo:hi].copy(), v.values[lo:hi].copy()))
    verdict = [None]
    if rank == 0:
        full = ddf.large_hypercube_manifold(3, nelts, ctx=ctx, hdiv=nxy)
        fu, fv = ddf.Fun(full, ddf.Pr, 0, gu), ddf.Fun(full, ddf.Pr, 0, gv)
        ddf._lib.call("ddf_wave_leapfrog", full._h, fu._h, fv._h, float(dt), int(steps))
        ru, rv = fu.values, fv.values
        bad = 0
        rows = 0
        for zlo, pu, pv in parts:
            a = zlo * slab.plane
            bad += int(np.sum(pu != ru[a:a + len(pu)]) + np.sum(pv != rv[a:a + len(pv)]))
            rows += len(pu)
        ok = bad == 0 and rows == total and float(np.max(np.abs(ru))) > 0.1
        verdict[0] = "bit-exact" if ok else f"MISMATCH: {bad} of {2 * total} values differ ({rows} owned rows of {total})"
        del fu, fv
        full.close()
    dist.broadcast_object_list(verdict, src=0)
    mode = ddf.dist.halo_mode(mfd)
    del u, v
    mfd.close()
    return {"verdict": verdict[0], "mesh": f"3D Kuhn {nelts}, {steps} fused leapfrog steps, random state", "halo": mode}


def run_b200(args):
    import ddf.jl_b200 as ddf
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N = args.gpus
    if world != N:
        if N == 1 and world == 1:
            pass
        else:
            raise SystemExit(f"--gpus {N} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {N}")
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if N > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = ddf.Context(local_rank)
    ddf.set_default_context(ctx)
    if N > 1:
        ddf.dist.init_comm(ctx)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # N > 1: before anything is timed, the slab-decomposed fused step is held against the undivided mesh on one GPU,
    # bit for bit, from a random state that is non-zero on the interface planes (a small mesh: 24 x 24 x 8N cells)
    parity_n = None
    if N > 1:
        parity_n = multi_gpu_parity(ddf, ctx, dist, rank, N)

    n = args.n
    t_setup0 = time.time()
    slab = None
    if N == 1:
        mfd = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx, assemble=False)
    else:
        # weak scaling: one n^3 cube per GPU stacked along z; --strong: ONE n^3 cube cut into N slabs
        mfd, slab = ddf.dist.slab_hypercube(3, (n, n, n if args.strong else n * N), ctx, assemble=False)
    # assembly is timed on the SECOND mesh of the process: the first cudaMalloc of tens of GB on a fresh device
    # maps memory for hundreds of ms, which says nothing about the kernels (both numbers are reported)
    ctx.sync()
    ctx.tic()
    ddf._lib.call("ddf_mesh_assemble", mfd._h)
    asm_first_ms = ctx.toc()
    asm_ms = float("inf")
    for _ in range(2):                   # best of two further assemblies (the pool is warm now)
        mfd.close()
        if N == 1:
            mfd = ddf.large_hypercube_manifold(3, (n, n, n), ctx=ctx, assemble=False)
        else:
            mfd, slab = ddf.dist.slab_hypercube(3, (n, n, n if args.strong else n * N), ctx, assemble=False)
        ctx.sync()
        ctx.tic()
        ddf._lib.call("ddf_mesh_assemble", mfd._h)
        asm_ms = min(asm_ms, ctx.toc())
    mfd._geometry()
    ctx.sync()
    cnt = [mfd.nsimplices(k) for k in range(4)]
    own = owned_counts(ddf, ctx, (n, n), slab, cnt)
    d = [ddf.deriv(ddf.Pr, k, mfd) for k in range(3)]
    h = [ddf.hodge(ddf.Pr, k, mfd) for k in range(4)]
    if slab is not None:
        # a slab rank applies d_k / star_k to the rows it OWNS (one contiguous range per rank by the lexicographic
        # numbering); the rows of ghost simplices belong to the neighbour rank and are neither computed nor counted
        rng_ = [mfd.owned_rows(k) for k in range(4)]
        assert [hi_ - lo_ for lo_, hi_ in rng_] == [int(c) for c in own], (rng_, own)
        d = [d[k].restrict_rows(*rng_[k + 1]) for k in range(3)]
        h = [h[k].restrict_rows(*rng_[k]) for k in range(4)]
    x = [ddf.rand_fun(mfd, ddf.Pr, k, seed=k + 1) for k in range(4)]
    y = [ddf.Fun(mfd, ddf.Pr, k + 1) for k in range(3)]
    z = [ddf.Fun(mfd, ddf.Dl, k) for k in range(4)]
    setup_s = time.time() - t_setup0

    def step(record=None):
        for k in range(3):
            if record is not None and k == 1:
                ctx.event_record(record)
            d[k].apply(x[k], y[k])
            if record is not None and k == 1:
                ctx.event_record(record + 1)
        for k in range(4):
            h[k].apply(x[k], z[k])

    K, W = args.steps, args.warmup
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("DDF_BENCH_NO_SAMPLER"):     # (diagnostic switch: is the polling itself visible?)
        sampler.start()          # sampled from the warm-up to the end of the timed wave steps
    for _ in range(W):
        step()
    barrier()
    l0 = ctx.launch_count()
    ctx.tic()
    for s in range(K):
        step(record=2 * s)
    ms_total = ctx.toc()
    barrier()
    launches = ctx.launch_count() - l0
    ms_step = max_over_ranks(ms_total / K)
    d1_ms = float(np.mean([ctx.event_elapsed(2 * s, 2 * s + 1) for s in range(K)]))
    d1_ms = max_over_ranks(d1_ms)
    dof_local = 2 * (own[1] + own[2] + own[3]) + own[0]
    dof_total = sum_over_ranks(float(dof_local))
    value = dof_total / (ms_step * 1e-3)

    # per-operator breakdown (each timed alone, back to back, inputs >> L2)
    breakdown = {}
    peak, peak_src = load_peaks()

    def time_op(fn, reps):
        fn()
        ctx.sync()
        ctx.tic()
        for _ in range(reps):
            fn()
        return ctx.toc() / reps
    for k in range(3):
        ms = time_op(lambda k=k: d[k].apply(x[k], y[k]), K)
        b = dk_bytes(cnt[k + 1], cnt[k], k)
        fmt = ddf.deriv_format(mfd, k)
        # bytes the stored format moves: bit-packed rows keep ONE 4- / 8-byte word per row, prefix-packed rows k+1
        # index words, both + one base per 32 rows
        fb = b - int((4 * (k + 2) - ddf.deriv_index_bytes(fmt, k)) * cnt[k + 1])
        breakdown[f"d{k}"] = {"ms": ms, "GBps": b / ms * 1e-6, "frac": b / ms * 1e-6 / peak, "bytes": b,
                              "row_format": fmt, "format_bytes": int(fb), "format_GBps": fb / ms * 1e-6,
                              "format_frac": fb / ms * 1e-6 / peak}
    for k in range(4):
        ms = time_op(lambda k=k: h[k].apply(x[k], z[k]), K)
        b = 24 * cnt[k]
        breakdown[f"star{k}"] = {"ms": ms, "GBps": b / ms * 1e-6, "frac": b / ms * 1e-6 / peak, "bytes": b}

    # fused wave step (leapfrog; halo exchange over NCCL when N > 1)
    u = ddf.sample(mfd, "wave")
    v = ddf.Fun(mfd, ddf.Pr, 0)
    dt = ddf.wave_dt(mfd)
    if N > 1:
        dt = ctx.allreduce(dt, "min")
    if args.wave_tune:
        ddf._lib.call("ddf_set_tuning", b"wave", int(args.wave_tune))
    ddf.wave_leapfrog(mfd, u, v, dt, W)
    barrier()
    ctx.tic()
    ddf.wave_leapfrog(mfd, u, v, dt, K)
    wave_ms = max_over_ranks(ctx.toc() / K)
    barrier()
    alg, fmt = ddf.wave_step_bytes(mfd)
    if slab is not None:     # owned share of the local slab
        frac_own = own[0] / cnt[0]
        alg, fmt = int(alg * frac_own), int(fmt * frac_own)
    unorm = u.norm(np.inf)
    wave = {"ms_per_step": wave_ms, "vertex_updates_per_s": sum_over_ranks(float(own[0])) / (wave_ms * 1e-3),
            "algorithmic_bytes": alg, "format_bytes": fmt, "GBps": alg / wave_ms * 1e-6,
            "frac": alg / wave_ms * 1e-6 / peak, "dt": dt, "u_inf_after": unorm, "row_format": ddf.wave_format(mfd),
            "halo": ddf.dist.halo_mode(mfd) if N > 1 else "none"}
    # BASELINE config 3: 2D scalar wave on the n=2048 Kuhn mesh (8.39e6 triangles), fused leapfrog step, 1 GPU
    wave2d = None
    if N == 1 and not args.no_wave2d:
        m2 = ddf.large_hypercube_manifold(2, (2048, 2048), ctx=ctx)
        u2, v2 = ddf.sample(m2, "wave"), ddf.Fun(m2, ddf.Pr, 0)
        dt2 = ddf.wave_dt(m2)
        ddf.wave_leapfrog(m2, u2, v2, dt2, 20)
        ctx.sync()
        ctx.tic()
        ddf.wave_leapfrog(m2, u2, v2, dt2, 200)
        ms2 = ctx.toc() / 200
        alg2, fmt2 = ddf.wave_step_bytes(m2)
        wave2d = {"workload": "BASELINE config 3: 2D Kuhn n=2048, fused leapfrog step", "ms_per_step": ms2,
                  "vertex_updates_per_s": m2.nsimplices(0) / (ms2 * 1e-3), "algorithmic_bytes": alg2,
                  "format_bytes": fmt2, "GBps": alg2 / ms2 * 1e-6, "frac": alg2 / ms2 * 1e-6 / peak,
                  "row_format": ddf.wave_format(m2), "steps": 200}
        # the integrator examples/wave.jl runs on this very configuration (levels = 11): fixed-step Tsit5
        ddf.wave_tsit5(m2, u2, v2, dt2, 2)
        ctx.sync()
        ctx.tic()
        ddf.wave_tsit5(m2, u2, v2, dt2, 50)
        wave2d["tsit5_ms_per_step"] = ctx.toc() / 50
        m2.close()
    # north_star kernel (3), Poisson flavour: one fused weighted-Jacobi sweep on the same mesh
    # algorithmic bytes (SURVEY 8d counting): E*(2*4+8) + V*(u r/w 16 + rho 8 + 1/diag 8 + mask 1)
    poisson_step = None
    if N == 1 and not args.no_extras:
        up_, rho_ = ddf.sample(mfd, "wave"), ddf.sample(mfd, "poisson_rho")
        ddf.poisson_jacobi(mfd, up_, rho_, 2.0 / 3.0, 5)
        ctx.sync()
        ctx.tic()
        ddf.poisson_jacobi(mfd, up_, rho_, 2.0 / 3.0, K)
        pj = ctx.toc() / K
        algp = cnt[1] * 16 + cnt[0] * 33
        poisson_step = {"ms_per_sweep": pj, "algorithmic_bytes": int(algp), "GBps": algp / pj * 1e-6,
                        "frac": algp / pj * 1e-6 / peak, "vertex_updates_per_s": cnt[0] / (pj * 1e-3)}
        del up_, rho_
    # deriv(Val(Dl), Val(k)) = the boundary matrices as plain CSC (ManifoldOps.jl:40-47): y = d_k^dual x, rows = (k-1)-simplices.
    # Bytes counted like SURVEY 8(d) counts d_k: int32 index per entry, row offset + y per row, x once.
    dual = None
    if N == 1 and not args.no_extras:
        dual = {}
        for k in (1, 2, 3):
            A = ddf.deriv(ddf.Dl, k, mfd)
            xd, yd = ddf.rand_fun(mfd, ddf.Dl, k, seed=k + 7), ddf.Fun(mfd, ddf.Dl, k - 1)
            ms = time_op(lambda: A.apply(xd, yd), K)
            b = (k + 1) * cnt[k] * 4 + cnt[k - 1] * 12 + cnt[k] * 8
            dual[f"dual_d{k}"] = {"ms": ms, "rows": cnt[k - 1], "bytes": int(b), "GBps": b / ms * 1e-6,
                                  "frac": b / ms * 1e-6 / peak}
            del A, xd, yd
    # the reference's own integrator on the same mesh: fixed-step Tsit5 (wave.jl:134), 6 right-hand sides per step
    tsit5 = None
    if not args.no_extras:
        # at N > 1 on the slab meshes: the ghost planes of the two new stage inputs travel after every pass (one grouped
        # ncclSend/ncclRecv per pass), timed like the wave step (barrier, max over ranks)
        ut, vt = ddf.sample(mfd, "wave"), ddf.Fun(mfd, ddf.Pr, 0)
        ddf.wave_tsit5(mfd, ut, vt, dt, 1)
        ctx.sync()
        barrier()
        ctx.tic()
        ddf.wave_tsit5(mfd, ut, vt, dt, 5)
        t5 = max_over_ranks(ctx.toc() / 5)
        barrier()
        # bytes of one Tsit5 step in this library's format: THREE reads of the staged rows of L (two stages share one
        # read: csrc/wave.cu tsit5_pair; the unfused integrator needs six) + the 29 vertex-vector passes of the three
        # paired-stage kernels (gathered stage inputs, base state, stored k's, outputs)
        _, fmt_b = ddf.wave_step_bytes(mfd)
        t5_bytes = 3 * (fmt_b - 32 * cnt[0]) + 29 * 8 * cnt[0]
        tsit5 = {"ms_per_step": t5, "rhs_per_step": 6, "launches_per_step": 3,
                 "vertex_updates_per_s": sum_over_ranks(float(own[0])) / (t5 * 1e-3),
                 "format_bytes": int(t5_bytes), "GBps": t5_bytes / t5 * 1e-6, "frac": t5_bytes / t5 * 1e-6 / peak,
                 "reads_of_L_per_step": 3}
        del ut, vt
    # BASELINE config 2: the Laplacian of examples/poisson.jl on the 2D simplex refined 10x (1,048,576 triangles;
    # Delaunay through scipy like the reference): y = L rho
    poisson2d = None
    if N == 1 and not args.no_extras:
        t0 = time.time()
        mp_ = ddf.refined_simplex_manifold(2, 10, ctx=ctx)
        t_mesh = time.time() - t0
        rho = ddf.sample(mp_, "poisson_rho")
        Lp = ddf.laplace(ddf.Pr, 0, mp_)
        yp = ddf.Fun(mp_, ddf.Pr, 0)
        Lp.apply(rho, yp)
        ctx.sync()
        ctx.tic()
        for _ in range(200):
            Lp.apply(rho, yp)
        msp = ctx.toc() / 200
        np2 = [mp_.nsimplices(k) for k in range(3)]
        poisson2d = {"workload": "BASELINE config 2: 2D simplex refined 10x, y = laplace(Pr,0) * rho", "nsimplices": np2,
                     "ms_per_apply": msp, "vertex_updates_per_s": np2[0] / (msp * 1e-3),
                     "row_format": ddf.wave_format(mp_), "mesh_build_s": t_mesh,
                     "note": "38 MB working set: L2-resident, launch-latency bound"}
        del rho, yp, Lp
        mp_.close()
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        hx = [torch.empty(c, dtype=torch.float64).pin_memory() for c in cnt]
        for k in range(4):
            hx[k].copy_(torch.from_numpy(np.random.default_rng(k).random(cnt[k])))
        hy = [torch.empty(cnt[k + 1], dtype=torch.float64).pin_memory() for k in range(3)]
        hz = [torch.empty(c, dtype=torch.float64).pin_memory() for c in cnt]
        nx, ny, nz_ = [t.numpy() for t in hx], [t.numpy() for t in hy], [t.numpy() for t in hz]

        def e2e_step():
            # uploads, kernels and downloads are queued on three streams (H2D / compute / D2H);
            # the library orders them per vector; ctx.sync() ends the step with every result on the host
            for k in range(4):
                x[k].upload_async(nx[k])
            for k in range(3):
                d[k].apply(x[k], y[k])
                y[k].download_async(ny[k])
            for k in range(4):
                h[k].apply(x[k], z[k])
                z[k].download_async(nz_[k])
            ctx.sync()
        e2e_step()
        barrier()
        Ke = max(1, min(K, args.e2e_steps))
        t0 = time.perf_counter()
        for _ in range(Ke):
            e2e_step()
        ctx.sync()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) / Ke * 1e3)
        barrier()
        h2d = 8 * sum(cnt)
        d2h = 8 * (cnt[1] + cnt[2] + cnt[3] + sum(cnt))
        # What the platform's host link can do at this N: the same pinned buffers copied up and down CONCURRENTLY on two
        # streams by every rank at once (no kernels) -- the ceiling of any end-to-end path on this box.  The e2e step
        # above is bound by it (D2H carries twice the bytes of H2D), and on a multi-GPU box the ranks share host links /
        # root ports, so the ceiling per rank drops with N: `frac_of_platform` separates that from the library.
        platform = None
        try:
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            nb = min(int(cnt[2]), 1 << 27)                       # up to 1 GiB per direction
            dsrc = torch.empty(nb, dtype=torch.float64, device="cuda")
            ddst = torch.empty(nb, dtype=torch.float64, device="cuda")
            hsrc, hdst = hx[2][:nb], hz[2][:nb]

            def duplex(reps):
                for _ in range(reps):
                    with torch.cuda.stream(s_up):
                        ddst.copy_(hsrc, non_blocking=True)
                    with torch.cuda.stream(s_dn):
                        hdst.copy_(dsrc, non_blocking=True)
                torch.cuda.synchronize()
            duplex(1)
            barrier()
            t0 = time.perf_counter()
            duplex(4)
            dt_p = max_over_ranks(time.perf_counter() - t0)
            barrier()
            per_dir = 4 * nb * 8 / dt_p * 1e-9                    # GB/s per direction per rank, all ranks at once
            t_plat = max(h2d, d2h) / (per_dir * 1e9) * 1e3        # ms this step's bytes need at that rate (duplex)
            platform = {"duplex_GBps_per_direction_per_rank": per_dir, "ranks_at_once": N, "ms_for_this_steps_bytes": t_plat,
                        "frac_of_platform": t_plat / e2e_ms,
                        "how": "torch pinned<->device copy_ of 1 GiB up and 1 GiB down concurrently on two streams, all ranks at once"}
            del dsrc, ddst
        except Exception as e:      # the probe must never break the bench
            platform = {"error": repr(e)}
        e2e = {"value": dof_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "platform": platform,
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": Ke,
               "api": "ddf_vec_upload_async -> ddf_op_apply -> ddf_vec_download_async -> ddf_sync "
                      "(pinned host buffers; H2D, compute and D2H on separate streams)"}

    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu:
        # bounded sample of the SAME workload (same n): a few whole passes, 1 thread (what Julia's sparse mul! uses),
        # and the all-core OpenMP variant beside it
        c1 = cpu_reference_pass(args.n_ref, 1, 8, seconds_cap=14.0)
        ca = cpu_reference_pass(args.n_ref, 1, 8, seconds_cap=6.0, threads_omp=True)
        cpu = {k: c1[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["allcores"] = {"value": ca["value"], "unit": UNIT, "cores": ca["cores"], "sample": ca["sample"]}

    if rank == 0:
        d1_bytes = dk_bytes(cnt[2], cnt[1], 1)
        d1_fmt = breakdown["d1"]["row_format"]
        d1_kernel = {"bits32": "k_apply_bits_rows<3,4,4>", "bits64": "k_apply_bits_rows<3,8,4>",
                     "packed": "k_apply_packed_rows<3,4>", "explicit": "k_apply_fixed_rows<3,4>"}[d1_fmt]
        d1_note = {"bits32": ", bit-packed 4-byte rows)", "bits64": ", bit-packed 8-byte rows)",
                   "packed": ", prefix-packed rows)", "explicit": ")"}[d1_fmt]
        d1_format_bytes = breakdown["d1"]["format_bytes"]
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch); scaled by the
        # algorithmic bytes when this run's slab differs slightly from the captured n=256 cube
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("n") == n:
                e = tj[d1_kernel]
                traffic = e["traffic_bytes"] * d1_bytes / e["algorithmic_bytes"]
                traffic_src = tj["source"]
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.strong and N > 1 else "weak",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, N, args.strong, cnt, dof_total),
            # `frac` is the PHYSICAL fraction: bytes the kernel really moves (its stored format = what ncu's
            # dram__bytes shows, `traffic`) / time / measured copy peak.  SURVEY 8(d)'s algorithmic count (explicit int32
            # indices) is larger than what the bit-packed rows store, so its fraction can exceed 1: kept as
            # `frac_algorithmic` beside it.
            "roofline": {"bound": "hbm", "kernel": d1_kernel + " (y = d1 x" + d1_note,
                         "achieved": d1_format_bytes / d1_ms * 1e-6,
                         "peak": peak, "unit": "GB/s", "frac": d1_format_bytes / d1_ms * 1e-6 / peak, "traffic": traffic,
                         "traffic_unit": "bytes per launch (ncu dram read+write)", "traffic_source": traffic_src,
                         "peak_source": peak_src, "bytes_per_launch": int(d1_format_bytes),
                         "bytes_counted": "stored format (bit-packed index word + y per row, x once) = ncu DRAM traffic",
                         "ms_per_launch": d1_ms,
                         "algorithmic_bytes_per_launch": int(d1_bytes),
                         "achieved_algorithmic": d1_bytes / d1_ms * 1e-6,
                         "frac_algorithmic": d1_bytes / d1_ms * 1e-6 / peak,
                         "frac_of_nominal_8TBs": d1_format_bytes / d1_ms * 1e-6 / 8000.0},
            # the one path with an exchange step: the fused wave step (per-N value for the scaling table)
            "wave_ms_per_step": wave["ms_per_step"], "wave_vertex_updates_per_s": wave["vertex_updates_per_s"],
            "wave_frac": wave["frac"], "wave_halo": wave["halo"], "parity_n": parity_n,
            # the integrator the reference itself runs (examples/wave.jl:134), same mesh / slabs: per-N value
            "wave_tsit5_ms_per_step": tsit5["ms_per_step"] if tsit5 else None,
            "breakdown": breakdown, "wave": wave, "wave2d": wave2d, "poisson_step": poisson_step, "dual_deriv": dual, "wave_tsit5": tsit5, "poisson2d": poisson2d,
            "assembly_ms": asm_ms,
            # assembly of d_3, d_2, d_1 on the device from the tet table (bit-exact; no roofline target):
            # face tuples generated, sorted and ranked per second
            "assembly": {"ms": asm_ms, "first_call_ms": asm_first_ms, "tuples": int(sum((R + 1) * cnt[R] for R in (1, 2, 3))),
                         "tuples_per_s": sum((R + 1) * cnt[R] for R in (1, 2, 3)) / (asm_ms * 1e-3)},
            "setup_s": setup_s,
            "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches), "clocks": clocks,
            "hbm_bytes_in_use": ctx.bytes_in_use(),
        }
        print_json(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    # libraries (NCCL_DEBUG=VERSION, torchrun banners) may write to fd 1: keep the real stdout for the ONE JSON
    # line and send everything else to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global print_json

    def print_json(line):
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--cells", dest="n", type=int, default=256,
                    help="cells per axis (per GPU along z); use --cells under torchrun (its parser rejects --n)")
    ap.add_argument("--n-ref", type=int, default=0, help="cells per axis of the CPU arm (default: the same n as the product arm)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-wave2d", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the Tsit5 and config-2 Poisson entries")
    ap.add_argument("--strong", action="store_true", help="N > 1: cut ONE n^3 cube into N slabs (default: weak scaling)")
    ap.add_argument("--wave-tune", type=int, default=0, help="ddf_set_tuning('wave', x) for A/B runs (1024: NCCL halo)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.n_ref <= 0:
        args.n_ref = args.n
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

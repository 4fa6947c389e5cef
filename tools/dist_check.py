#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
the slab-decomposed fused wave step with NCCL ghost-plane exchange must reproduce the
single-GPU run of the undivided mesh bit for bit (owned rows).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/dist_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import ddf.jl_b200 as ddf  # noqa: E402


def global_field(seed, slab, nelts, D):
    """This slab's part of ONE random vertex field of the undivided mesh (same numbers on every rank for a given
    seed): non-symmetric and non-zero on the interface planes, unlike prod sin(pi x) whose zeros sit exactly there."""
    total = int(np.prod([n + 1 for n in nelts]))
    g = np.random.default_rng(seed).standard_normal(total)
    lo = slab.z_cell_lo * slab.plane
    return g[lo:lo + slab.n_local_planes * slab.plane].copy(), g


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ddf.Context(local)
    ddf.dist.init_comm(ctx)
    ok = True
    # the last case is large enough per rank (3200 row windows) for TWO steps per launch with the deep halo: 10 steps in
    # calls of 1, 2 and 7 = single step, one pair, three pairs + a single step
    cases = ((3, (8, 8, 8 * world), 9), (2, (16, 16 * world), 12), (3, (16, 16, 4 * world + 4), 5),
             (3, (48, 48, 16 * world), 7), (3, (96, 96, 80 * world), 10))
    for tune, (D, nelts, nsteps) in [(t, c) for t in (0, 1024) for c in cases]:
        ddf._lib.call("ddf_set_tuning", b"wave", tune)       # 0: fused peer-memory halo, 1024: NCCL send/recv
        mfd, slab = ddf.dist.slab_hypercube(D, nelts, ctx)
        lu, gu = global_field(100 + D, slab, nelts, D)
        lv, gv = global_field(200 + D, slab, nelts, D)
        u = ddf.Fun(mfd, ddf.Pr, 0, lu)
        v = ddf.Fun(mfd, ddf.Pr, 0, lv)
        dt = ctx.allreduce(ddf.wave_dt(mfd), "min")
        # several calls with odd and even step counts: the ping-pong buffers swap roles between calls
        done = 0
        launches = []
        for part in (1, 2, nsteps - 3):
            ctx.sync()
            l0 = ctx.launch_count()
            ddf.wave_leapfrog(mfd, u, v, dt, part)
            ctx.sync()
            launches.append(ctx.launch_count() - l0)
            done += part
        assert done == nsteps
        mode = ddf.dist.halo_mode(mfd)
        lo, hi = slab.own_rows
        mine = (slab.z_own_lo, u.values[lo:hi].copy(), v.values[lo:hi].copy(), u.norm(np.inf))
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            full = ddf.large_hypercube_manifold(D, nelts, ctx=ctx, hdiv=nelts[0])
            fu = ddf.Fun(full, ddf.Pr, 0, gu)
            fv = ddf.Fun(full, ddf.Pr, 0, gv)
            fdt = ddf.wave_dt(full)
            ddf.wave_leapfrog_single = None
            ddf._lib.call("ddf_wave_leapfrog", full._h, fu._h, fv._h, float(dt), int(nsteps))
            ru, rv = fu.values, fv.values
            plane = slab.plane
            same = fdt == dt
            for zlo, pu, pv, _ in sorted(parts, key=lambda t: t[0]):
                a = zlo * plane
                same = same and np.array_equal(pu, ru[a:a + len(pu)]) and np.array_equal(pv, rv[a:a + len(pv)])
            print(f"dist_check halo={mode} D={D} nelts={nelts} world={world} steps={nsteps}: "
                  f"{'BIT-EXACT' if same else 'MISMATCH'}  |u|inf={np.max(np.abs(ru)):.6f}  launches per call {launches}", flush=True)
            ok = ok and same
            full.close()
        mfd.close()
    # fused Poisson relaxation sweeps on slabs, both halo protocols
    for tune, (D, nelts, nsw) in [(t, c) for t in (0, 1024) for c in ((3, (8, 8, 8 * world), 7), (3, (32, 32, 8 * world), 4))]:
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        mfd, slab = ddf.dist.slab_hypercube(D, nelts, ctx)
        lu, gu = global_field(300 + D, slab, nelts, D)
        lr, gr = global_field(400 + D, slab, nelts, D)
        u, rho = ddf.Fun(mfd, ddf.Pr, 0, lu), ddf.Fun(mfd, ddf.Pr, 0, lr)
        for part in (1, 2, nsw - 3):
            ddf.poisson_jacobi(mfd, u, rho, 2.0 / 3.0, part)
        mode = ddf.dist.halo_mode(mfd)
        lo, hi = slab.own_rows
        parts = [None] * world
        dist.all_gather_object(parts, (slab.z_own_lo, u.values[lo:hi].copy()))
        if rank == 0:
            full = ddf.large_hypercube_manifold(D, nelts, ctx=ctx, hdiv=nelts[0])
            fu, frho = ddf.Fun(full, ddf.Pr, 0, gu), ddf.Fun(full, ddf.Pr, 0, gr)
            ddf._lib.call("ddf_poisson_jacobi", full._h, fu._h, frho._h, 2.0 / 3.0, int(nsw))
            ru = fu.values
            same = True
            for zlo, pu in sorted(parts, key=lambda t: t[0]):
                a = zlo * slab.plane
                same = same and np.array_equal(pu, ru[a:a + len(pu)])
            print(f"dist_check jacobi halo={mode} D={D} nelts={nelts} world={world} sweeps={nsw}: "
                  f"{'BIT-EXACT' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
            full.close()
        mfd.close()
    ddf._lib.call("ddf_set_tuning", b"wave", 0)
    # the reference's integrator on slabs: fixed-step Tsit5 with a ghost-plane exchange before every right-hand side
    # (paired stages: peer-memory halo pass by pass -- tuning 0 -- or one grouped ncclSend/ncclRecv per pass -- 1024;
    # several calls, so the step counter of the halo protocol runs across calls and mixes with the leapfrog's)
    for tune, (D, nelts, nsteps) in [(t, c) for t in (0, 1024) for c in ((3, (8, 8, 8 * world), 3), (2, (16, 16 * world), 4),
                                                                         (3, (24, 24, 8 * world + 2), 4))]:
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        mfd, slab = ddf.dist.slab_hypercube(D, nelts, ctx)
        lu, gu = global_field(500 + D, slab, nelts, D)
        lv, gv = global_field(600 + D, slab, nelts, D)
        u, v = ddf.Fun(mfd, ddf.Pr, 0, lu), ddf.Fun(mfd, ddf.Pr, 0, lv)
        dt = ctx.allreduce(ddf.wave_dt(mfd), "min")
        ddf.wave_tsit5(mfd, u, v, dt, 1)
        ddf.wave_tsit5(mfd, u, v, dt, nsteps - 1)
        t5mode = ddf.dist.halo_mode(mfd)
        lo, hi = slab.own_rows
        parts = [None] * world
        dist.all_gather_object(parts, (slab.z_own_lo, u.values[lo:hi].copy(), v.values[lo:hi].copy()))
        if rank == 0:
            full = ddf.large_hypercube_manifold(D, nelts, ctx=ctx, hdiv=nelts[0])
            fu, fv = ddf.Fun(full, ddf.Pr, 0, gu), ddf.Fun(full, ddf.Pr, 0, gv)
            ctx.nranks_save = None
            ddf._lib.call("ddf_wave_tsit5", full._h, fu._h, fv._h, float(dt), int(nsteps))
            ru, rv = fu.values, fv.values
            same = True
            for zlo, pu, pv in sorted(parts, key=lambda t: t[0]):
                a = zlo * slab.plane
                same = same and np.array_equal(pu, ru[a:a + len(pu)]) and np.array_equal(pv, rv[a:a + len(pv)])
            print(f"dist_check tsit5 halo={t5mode} D={D} nelts={nelts} world={world} steps={nsteps}: {'BIT-EXACT' if same else 'MISMATCH'}",
                  flush=True)
            ok = ok and same
            full.close()
        mfd.close()
    # the two integrators interleaved on ONE slab mesh: leapfrog, Tsit5, leapfrog again -- the step counter and the
    # landing parities of the peer-memory protocol run through both (and through the NCCL fallback)
    for tune in (0, 1024):
        ddf._lib.call("ddf_set_tuning", b"wave", tune)
        D, nelts = 3, (16, 16, 8 * world + 2)
        mfd, slab = ddf.dist.slab_hypercube(D, nelts, ctx)
        lu, gu = global_field(650 + D, slab, nelts, D)
        lv, gv = global_field(660 + D, slab, nelts, D)
        u, v = ddf.Fun(mfd, ddf.Pr, 0, lu), ddf.Fun(mfd, ddf.Pr, 0, lv)
        dt = ctx.allreduce(ddf.wave_dt(mfd), "min")
        ddf.wave_leapfrog(mfd, u, v, dt, 3)
        ddf.wave_tsit5(mfd, u, v, dt, 2)
        ddf.wave_leapfrog(mfd, u, v, dt, 2)
        ddf.wave_tsit5(mfd, u, v, dt, 1)
        mmode = ddf.dist.halo_mode(mfd)
        lo, hi = slab.own_rows
        parts = [None] * world
        dist.all_gather_object(parts, (slab.z_own_lo, u.values[lo:hi].copy(), v.values[lo:hi].copy()))
        if rank == 0:
            full = ddf.large_hypercube_manifold(D, nelts, ctx=ctx, hdiv=nelts[0])
            fu, fv = ddf.Fun(full, ddf.Pr, 0, gu), ddf.Fun(full, ddf.Pr, 0, gv)
            ddf._lib.call("ddf_wave_leapfrog", full._h, fu._h, fv._h, float(dt), 3)
            ddf._lib.call("ddf_wave_tsit5", full._h, fu._h, fv._h, float(dt), 2)
            ddf._lib.call("ddf_wave_leapfrog", full._h, fu._h, fv._h, float(dt), 2)
            ddf._lib.call("ddf_wave_tsit5", full._h, fu._h, fv._h, float(dt), 1)
            ru, rv = fu.values, fv.values
            same = True
            for zlo, pu, pv in sorted(parts, key=lambda t: t[0]):
                a = zlo * slab.plane
                same = same and np.array_equal(pu, ru[a:a + len(pu)]) and np.array_equal(pv, rv[a:a + len(pv)])
            print(f"dist_check leapfrog+tsit5 interleaved halo={mmode} D={D} nelts={nelts} world={world}: "
                  f"{'BIT-EXACT' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
            full.close()
        mfd.close()
    ddf._lib.call("ddf_set_tuning", b"wave", 0)
    # ghost exchange of cochains of every rank: a globally consistent R-cochain (a function of the simplex's
    # global vertex coordinates), ghost entries wiped, exchanged, must come back bit for bit
    ddf._lib.call("ddf_set_tuning", b"wave", 0)
    for D, nelts in ((3, (8, 8, 8 * world)), (2, (16, 16 * world)), (3, (24, 24, 6 * world + 2))):
        mfd, slab = ddf.dist.slab_hypercube(D, nelts, ctx)
        coords = mfd.get_coords()
        g = np.sin(coords @ np.arange(1, D + 1) * 3.7) + coords[:, -1] ** 2
        same_all = True
        for R in range(D + 1):
            full = g.copy() if R == 0 else g[mfd.get_simplices(R)].sum(axis=1) * (R + 0.5)
            own = ddf.dist.owned_mask(mfd, slab, R)
            wiped = np.where(own, full, -777.0)
            f = ddf.Fun(mfd, ddf.Pr, R, wiped)
            ddf.dist.halo_exchange(mfd, f)
            got = f.values
            same = bool(np.array_equal(got, full))
            same_all = same_all and same
            if not same:
                bad = np.nonzero(got != full)[0]
                print(f"rank {rank}: halo_exchange R={R} D={D}: {len(bad)} of {len(full)} entries differ "
                      f"({int((~own).sum())} ghosts)", flush=True)
        res = [None] * world
        dist.all_gather_object(res, same_all)
        if rank == 0:
            print(f"dist_check halo_exchange_k D={D} nelts={nelts} world={world} ranks 0..{D}: "
                  f"{'BIT-EXACT' if all(res) else 'MISMATCH'}", flush=True)
            ok = ok and all(res)
        mfd.close()
    # global numbering + sharded d_k / star_k apply: every rank numbers its owned simplices into the
    # undivided mesh's id space, applies its local d_R / star_R to the matching entries of ONE global cochain, and the
    # owned rows of all ranks together must be the undivided mesh's tables and results bit for bit
    for D, nelts in ((3, (8, 8, 8 * world)), (2, (16, 16 * world)), (3, (20, 20, 6 * world + 2))):
        mfd, slab = ddf.dist.slab_hypercube(D, nelts, ctx)
        voff = slab.z_cell_lo * slab.plane                       # local -> global vertex id
        gids = [ddf.dist.global_ids(mfd, slab, R) for R in range(D + 1)]
        totals = [ddf.dist.owned_offsets(mfd, slab, R)[1] for R in range(D + 1)]
        payload = {"tables": [], "d": [], "star": []}
        for R in range(D + 1):
            own = ddf.dist.owned_mask(mfd, slab, R)
            tab = (np.arange(mfd.nsimplices(0))[:, None] if R == 0 else mfd.get_simplices(R)) + voff
            payload["tables"].append((gids[R][own], tab[own]))
            xg = np.random.default_rng(40 + R).standard_normal(totals[R])      # the same global cochain on every rank
            x = ddf.Fun(mfd, ddf.Pr, R, xg[gids[R]])
            ys = (ddf.hodge(ddf.Pr, R, mfd) * x).values
            payload["star"].append((gids[R][own], ys[own]))
            if R < D:
                own1 = ddf.dist.owned_mask(mfd, slab, R + 1)
                yd = (ddf.deriv(ddf.Pr, R, mfd) * x).values
                payload["d"].append((gids[R + 1][own1], yd[own1]))
        parts = [None] * world
        dist.all_gather_object(parts, payload)
        if rank == 0:
            full = ddf.large_hypercube_manifold(D, nelts, ctx=ctx, hdiv=nelts[0])
            same = all(totals[R] == full.nsimplices(R) for R in range(D + 1))
            for R in range(D + 1):
                ftab = np.arange(full.nsimplices(0))[:, None] if R == 0 else full.get_simplices(R)
                xg = np.random.default_rng(40 + R).standard_normal(full.nsimplices(R))
                fx = ddf.Fun(full, ddf.Pr, R, xg)
                fs = (ddf.hodge(ddf.Pr, R, full) * fx).values
                fd = (ddf.deriv(ddf.Pr, R, full) * fx).values if R < D else None
                seen = np.zeros(full.nsimplices(R), dtype=np.int64)
                seen1 = np.zeros(full.nsimplices(R + 1), dtype=np.int64) if R < D else None
                for p in parts:
                    g, t = p["tables"][R]
                    same = same and np.array_equal(ftab[g], t)
                    np.add.at(seen, g, 1)
                    g, y = p["star"][R]
                    same = same and np.array_equal(fs[g], y)
                    if R < D:
                        g, y = p["d"][R]
                        same = same and np.array_equal(fd[g], y)
                        np.add.at(seen1, g, 1)
                same = same and bool((seen == 1).all()) and (seen1 is None or bool((seen1 == 1).all()))   # a partition
            print(f"dist_check global numbering + sharded d_k/star_k D={D} nelts={nelts} world={world}: "
                  f"{'BIT-EXACT' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
            full.close()
        mfd.close()
    # meshes of ANY origin partitioned by vertex ranges (ddf.dist.partition_mesh + ddf_mesh_set_halo_lists): Delaunay
    # meshes in Qhull's scattered simplex order, vertices numbered along one axis; fused leapfrog, Poisson sweeps, Tsit5
    # and the sharded d_0 / star_0 / star_1 apply on owned rows, each against the undivided mesh on one GPU
    from scipy.spatial import Delaunay
    for tag, npts, dim in (("delaunay2d", 30000, 2), ("delaunay3d", 9000, 3)):
        rng = np.random.default_rng(70 + dim)
        pts = rng.random((npts, dim))
        pts = pts[np.argsort(pts[:, -1])]
        simp = np.sort(Delaunay(pts).simplices.astype(np.int64), axis=1)
        wts = np.zeros(npts)
        mfd, part = ddf.dist.distribute_mesh(tag, simp, pts, wts, ctx)
        l2g = part.local_to_global
        gu, gv, gr = (np.random.default_rng(k + dim).standard_normal(npts) for k in (700, 800, 900))
        lo, hi = part.own_lo, part.own_hi
        assert mfd.owned_rows(0) == (lo, hi)
        dt = 1e-4
        u, v = ddf.Fun(mfd, ddf.Pr, 0, gu[l2g]), ddf.Fun(mfd, ddf.Pr, 0, gv[l2g])
        for chunk in (1, 2, 4):
            ddf.wave_leapfrog(mfd, u, v, dt, chunk)
        pj, rho = ddf.Fun(mfd, ddf.Pr, 0, gu[l2g]), ddf.Fun(mfd, ddf.Pr, 0, gr[l2g])
        for chunk in (1, 3):
            ddf.poisson_jacobi(mfd, pj, rho, 2.0 / 3.0, chunk)
        tu, tv = ddf.Fun(mfd, ddf.Pr, 0, gu[l2g]), ddf.Fun(mfd, ddf.Pr, 0, gv[l2g])
        ddf.wave_tsit5(mfd, tu, tv, dt, 3)
        # sharded applies: d_0 on the owned edges, star_0 / star_1 on the owned vertices / edges, keyed by global tuples
        e_lo, e_hi = mfd.owned_rows(1)
        edges = l2g[mfd.get_simplices(1)[e_lo:e_hi]]
        x0 = ddf.Fun(mfd, ddf.Pr, 0, gu[l2g])
        d0 = (ddf.deriv(ddf.Pr, 0, mfd) * x0).values[e_lo:e_hi]
        s0 = (ddf.hodge(ddf.Pr, 0, mfd) * x0).values[lo:hi]
        s1 = mfd.get_dualvolumes(1)[e_lo:e_hi] / mfd.get_volumes(1)[e_lo:e_hi]
        mine = dict(glo=int(part.cuts[rank]), u=u.values[lo:hi], v=v.values[lo:hi], pj=pj.values[lo:hi], tu=tu.values[lo:hi],
                    tv=tv.values[lo:hi], edges=edges, d0=d0, s0=s0, s1=s1)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            full = ddf.Manifold.from_simplices(tag, simp, pts, wts, ctx=ctx)
            full._geometry()
            fu, fv = ddf.Fun(full, ddf.Pr, 0, gu), ddf.Fun(full, ddf.Pr, 0, gv)
            ddf._lib.call("ddf_wave_leapfrog", full._h, fu._h, fv._h, float(dt), 7)
            fp, frho = ddf.Fun(full, ddf.Pr, 0, gu), ddf.Fun(full, ddf.Pr, 0, gr)
            ddf._lib.call("ddf_poisson_jacobi", full._h, fp._h, frho._h, 2.0 / 3.0, 4)
            ftu, ftv = ddf.Fun(full, ddf.Pr, 0, gu), ddf.Fun(full, ddf.Pr, 0, gv)
            ddf._lib.call("ddf_wave_tsit5", full._h, ftu._h, ftv._h, float(dt), 3)
            fx = ddf.Fun(full, ddf.Pr, 0, gu)
            fd0 = (ddf.deriv(ddf.Pr, 0, full) * fx).values
            fs0 = (ddf.hodge(ddf.Pr, 0, full) * fx).values
            fs1 = full.get_dualvolumes(1) / full.get_volumes(1)
            fedges = full.get_simplices(1)
            same, eoff = True, 0
            for q in sorted(parts, key=lambda t: t["glo"]):
                a, n = q["glo"], len(q["u"])
                for key, ref in (("u", fu.values), ("v", fv.values), ("pj", fp.values), ("tu", ftu.values),
                                 ("tv", ftv.values), ("s0", fs0)):
                    same = same and np.array_equal(q[key], ref[a:a + n])
                ne = len(q["edges"])
                same = same and np.array_equal(q["edges"], fedges[eoff:eoff + ne])       # owned edges: contiguous, in order
                same = same and np.array_equal(q["d0"], fd0[eoff:eoff + ne]) and np.array_equal(q["s1"], fs1[eoff:eoff + ne])
                eoff += ne
            same = same and eoff == full.nsimplices(1)
            print(f"dist_check general partition {tag} nverts={npts} world={world} (leapfrog, jacobi, tsit5, d0/star on owned rows): "
                  f"{'BIT-EXACT' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
            full.close()
        mfd.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

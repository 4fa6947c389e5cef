"""CPU tier: host-side logic of the multi-GPU path (slab partition + halo protocol),
exercised with world_size 2 and 3 on the gloo backend.  The device side replaces the
gloo send/recv by ncclSend/ncclRecv on ghost planes (ddf.jl_b200/csrc/comm.cu); the
plane bookkeeping tested here is the same Python code the GPU path uses.

The emulation applies the oracle's Laplacian rows on each rank's slab mesh (owned rows
only) and exchanges ghost planes, and must reproduce the undivided oracle bit for bit.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_everything():
    import ddf.jl_b200 as ddf
    for nz, G in ((8, 1), (8, 2), (12, 3), (16, 4), (512, 8), (2048, 8)):
        plane = 25
        slabs = ddf.dist.partition_slabs(nz, G, plane)
        assert slabs[0].z_own_lo == 0 and slabs[-1].z_own_hi == nz + 1
        for a, b in zip(slabs, slabs[1:]):
            assert a.z_own_hi == b.z_own_lo                     # owned planes tile [0, nz]
        for s in slabs:
            assert s.z_cell_lo % 2 == 0 and s.nz_cells % 2 == 0  # Kuhn mirror parity (ManifoldConstructors.jl:208)
            assert s.ghost_lo == (2 if s.rank > 0 else 0)
            assert s.ghost_hi == (3 if s.rank < G - 1 else 0)
            lo, hi = s.own_rows
            assert (hi - lo) == (s.z_own_hi - s.z_own_lo) * plane
            assert s.z_cell_lo + s.ghost_lo == s.z_own_lo
    with pytest.raises(ValueError):
        ddf.dist.partition_slabs(7, 2, 4)
    with pytest.raises(ValueError):
        ddf.dist.partition_slabs(4, 3, 4)


def _worker(rank, world, port, D, nxy, nz, nsteps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ddf_oracle as o
    from importlib import import_module
    dd = import_module("ddf.jl_b200.dist")
    nelts = (nxy,) * (D - 1) + (nz,)
    plane = (nxy + 1) ** (D - 1)
    slab = dd.partition_slabs(nz, world, plane)[rank]
    # local slab mesh with global coordinates / weights (what ddf_mesh_create_hypercube builds on the device)
    simp, coords, weights = o.large_hypercube_tables(D, (nxy,) * (D - 1) + (slab.nz_cells,), h_div=nxy)
    coords = coords.copy()
    coords[:, -1] += slab.z_cell_lo / nxy
    # level parity of the weights follows the GLOBAL z index; z_cell_lo is even so local parity == global parity
    m = o.build_manifold("slab", simp, coords, weights)
    b, L = o.wave_operators(m)
    mask = 1.0 - b.astype(np.float64)
    lo, hi = slab.own_rows
    u = o.wave_u0(m.coords)
    v = np.zeros_like(u)
    dt = 1.0 / (2 * nxy)
    Lr = L.tocsr()
    Lr.sort_indices()
    for _ in range(nsteps):
        prod = Lr.data * u[Lr.indices]
        lu = o._segment_sum_sequential(prod, Lr.indptr)
        vn = v.copy()
        un = u.copy()
        vn[lo:hi] = v[lo:hi] + dt * (mask[lo:hi] * lu[lo:hi])
        un[lo:hi] = u[lo:hi] + dt * (mask[lo:hi] * vn[lo:hi])
        # halo: my lowest owned plane -> rank-1's first upper ghost plane; my highest owned -> rank+1's last lower ghost
        reqs = []
        recv_lo = torch.empty(plane, dtype=torch.float64)
        recv_hi = torch.empty(plane, dtype=torch.float64)
        if slab.ghost_lo:
            reqs.append(dist.isend(torch.from_numpy(un[lo:lo + plane].copy()), rank - 1))
            reqs.append(dist.irecv(recv_lo, rank - 1))
        if slab.ghost_hi:
            reqs.append(dist.isend(torch.from_numpy(un[hi - plane:hi].copy()), rank + 1))
            reqs.append(dist.irecv(recv_hi, rank + 1))
        for r in reqs:
            r.wait()
        if slab.ghost_lo:
            un[lo - plane:lo] = recv_lo.numpy()
        if slab.ghost_hi:
            un[hi:hi + plane] = recv_hi.numpy()
        u, v = un, vn
    q.put((rank, slab.z_own_lo, slab.z_own_hi, u[lo:hi].copy(), v[lo:hi].copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,D,nxy,nz", [(2, 2, 4, 8), (2, 3, 4, 8), (3, 3, 2, 12)])
def test_slab_wave_equals_undivided(world, D, nxy, nz):
    """World-size-N leapfrog on slab meshes + ghost-plane exchange == the undivided oracle, bit for bit.
    (Geometry and Laplacian rows of owned vertices are identical because the full star is local.)"""
    import ddf_oracle as o
    nsteps = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 7 + D
    procs = [ctx.Process(target=_worker, args=(r, world, port, D, nxy, nz, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = o.large_hypercube_manifold(D, (nxy,) * (D - 1) + (nz,), h_div=nxy)
    ou, ov = o.wave_leapfrog(m, o.wave_u0(m.coords), np.zeros(m.nsimplices(0)), 1.0 / (2 * nxy), nsteps)
    plane = (nxy + 1) ** (D - 1)
    for rank, zlo, zhi, u, v in sorted(res):
        assert np.array_equal(u, ou[zlo * plane:zhi * plane]), f"rank {rank} u"
        assert np.array_equal(v, ov[zlo * plane:zhi * plane]), f"rank {rank} v"


def test_peer_memory_halo_protocol_model():
    """Model of the fused peer-memory halo of csrc/comm.cu (leapfrog_dist_p2p): every rank writes its interface
    planes into the NEIGHBOURS' landing areas (parity k&1), release-stores the step number k into their flag
    words, and before step k waits for flags >= k-1 and copies landing[(k-1)&1] into its ghost planes.  Three
    ranks run as threads with random delays (any interleaving the flags allow); the owned rows must still equal
    the undivided oracle bit for bit -- i.e. the protocol has neither a RAW nor a WAR hazard."""
    import random
    import threading
    import time
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ddf_oracle as o
    from importlib import import_module
    dd = import_module("ddf.jl_b200.dist")
    D, nxy, nz, world, nsteps = 3, 2, 12, 3, 9
    calls = (1, 2, 6)                                   # several calls: the ghost copy is skipped on a call's first step
    plane = (nxy + 1) ** (D - 1)
    slabs = dd.partition_slabs(nz, world, plane)
    flags = [[0, 0] for _ in range(world)]              # [from lower, from upper]
    landing = [np.zeros((2, 2, plane)) for _ in range(world)]   # [from lower / from upper][parity][P]
    results = [None] * world
    errors = []

    def rank_main(rank):
        try:
            rnd = random.Random(100 + rank)
            slab = slabs[rank]
            simp, coords, weights = o.large_hypercube_tables(D, (nxy,) * (D - 1) + (slab.nz_cells,), h_div=nxy)
            coords = coords.copy()
            coords[:, -1] += slab.z_cell_lo / nxy
            m = o.build_manifold("slab", simp, coords, weights)
            b, L = o.wave_operators(m)
            mask = 1.0 - b.astype(np.float64)
            lo, hi = slab.own_rows
            has_lo, has_hi = slab.ghost_lo > 0, slab.ghost_hi > 0
            u = o.wave_u0(m.coords)
            v = np.zeros_like(u)
            dt = 1.0 / (2 * nxy)
            Lr = L.tocsr()
            Lr.sort_indices()
            k = 0

            def wait_copy(uvec, need, copy, par):
                t0 = time.time()
                while (has_lo and flags[rank][0] < need) or (has_hi and flags[rank][1] < need):
                    if time.time() - t0 > 60:
                        raise RuntimeError("flag timeout")
                    time.sleep(0)
                if copy:
                    if has_lo:
                        uvec[lo - plane:lo] = landing[rank][0][par]
                    if has_hi:
                        uvec[hi:hi + plane] = landing[rank][1][par]

            for nst in calls:
                for s in range(nst):
                    k += 1
                    time.sleep(rnd.random() * 0.003)
                    wait_copy(u, k - 1, s > 0, (k - 1) & 1)
                    lu = o._segment_sum_sequential(Lr.data * u[Lr.indices], Lr.indptr)
                    vn, un = v.copy(), u.copy()
                    vn[lo:hi] = v[lo:hi] + dt * (mask[lo:hi] * lu[lo:hi])
                    un[lo:hi] = u[lo:hi] + dt * (mask[lo:hi] * vn[lo:hi])
                    time.sleep(rnd.random() * 0.003)
                    if has_lo:                           # I am the UPPER neighbour of rank-1
                        landing[rank - 1][1][k & 1][:] = un[lo:lo + plane]
                    if has_hi:                           # ... and the LOWER neighbour of rank+1
                        landing[rank + 1][0][k & 1][:] = un[hi - plane:hi]
                    if has_lo:
                        flags[rank - 1][1] = k
                    if has_hi:
                        flags[rank + 1][0] = k
                    u, v = un, vn
                wait_copy(u, k, True, k & 1)             # leave the call with valid ghost planes
            results[rank] = (slab.z_own_lo, slab.z_own_hi, u[lo:hi].copy(), v[lo:hi].copy())
        except Exception as e:                            # pragma: no cover
            errors.append((rank, repr(e)))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    m = o.large_hypercube_manifold(D, (nxy,) * (D - 1) + (nz,), h_div=nxy)
    ou, ov = o.wave_leapfrog(m, o.wave_u0(m.coords), np.zeros(m.nsimplices(0)), 1.0 / (2 * nxy), nsteps)
    for zlo, zhi, u, v in results:
        assert np.array_equal(u, ou[zlo * plane:zhi * plane])
        assert np.array_equal(v, ov[zlo * plane:zhi * plane])


def test_rank_k_ghost_lists_match_by_position():
    """Host model of ddf_halo_exchange_k (csrc/comm.cu): with ownership = owner of the smallest vertex and the
    four list predicates of k_halo_flags, the i-th simplex of a rank's send list is the i-th simplex of the
    neighbour's receive list (same GLOBAL vertex tuple) -- because local ids are lexicographic ranks and slabs
    differ by a constant vertex shift.  Checked for every rank R on 3 slabs of a 3D and a 2D Kuhn mesh."""
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ddf_oracle as o
    from importlib import import_module
    dd = import_module("ddf.jl_b200.dist")
    for D, nxy, nz, world in ((3, 2, 14, 3), (2, 4, 12, 3)):
        plane = (nxy + 1) ** (D - 1)
        slabs = dd.partition_slabs(nz, world, plane)
        meshes = []
        for s in slabs:
            simp, coords, weights = o.large_hypercube_tables(D, (nxy,) * (D - 1) + (s.nz_cells,), h_div=nxy)
            meshes.append(o.build_manifold("slab", simp, coords, weights))

        def lists(rank, R):
            s, m = slabs[rank], meshes[rank]
            tab = np.arange(m.nsimplices(0))[:, None] if R == 0 else m.simplices[R]
            pmin, pmax = tab[:, 0] // plane, tab[:, -1] // plane
            nplanes = s.n_local_planes
            own_hi = nplanes - s.ghost_hi
            peer_lo_ghost_hi = slabs[rank - 1].ghost_hi if rank > 0 else 0
            peer_hi_ghost_lo = slabs[rank + 1].ghost_lo if rank < world - 1 else 0
            send_lo = np.nonzero((s.ghost_lo > 0) & (pmin >= s.ghost_lo) & (pmax <= s.ghost_lo + peer_lo_ghost_hi - 1))[0]
            send_hi = np.nonzero((s.ghost_hi > 0) & (pmin >= own_hi - peer_hi_ghost_lo) & (pmin < own_hi))[0]
            recv_lo = np.nonzero((s.ghost_lo > 0) & (pmin < s.ghost_lo))[0]
            recv_hi = np.nonzero((s.ghost_hi > 0) & (pmin >= own_hi))[0]
            glob = tab + s.z_cell_lo * plane                       # global vertex ids
            return {k: glob[v] for k, v in (("send_lo", send_lo), ("send_hi", send_hi), ("recv_lo", recv_lo),
                                            ("recv_hi", recv_hi))}, tab.shape[0]
        for R in range(D + 1):
            L = [lists(r, R) for r in range(world)]
            for r in range(world - 1):
                a, b = L[r][0], L[r + 1][0]
                assert len(a["send_hi"]) > 0 and len(b["send_lo"]) > 0
                assert np.array_equal(a["send_hi"], b["recv_lo"]), (D, R, r)
                assert np.array_equal(b["send_lo"], a["recv_hi"]), (D, R, r)
            # every ghost simplex is in exactly one receive list, no owned simplex is
            for r in range(world):
                s, m = slabs[r], meshes[r]
                tab = np.arange(m.nsimplices(0))[:, None] if R == 0 else m.simplices[R]
                lo, hi = s.own_rows
                nghost = int(np.sum((tab[:, 0] < lo) | (tab[:, 0] >= hi)))
                assert len(L[r][0]["recv_lo"]) + len(L[r][0]["recv_hi"]) == nghost


def test_global_numbering_of_owned_simplices():
    """Host model of ddf.dist.global_ids / owned_offsets: with ownership = owner of the smallest vertex,
    every rank's owned R-simplices are ONE contiguous range of the undivided mesh's ids (Manifolds.jl:735-757:
    lexicographic ranks), in rank order, and inside the range the local order is the global order.  So
    global id = exclusive scan of the owned counts + rank among the owned -- checked by comparing vertex tuples with
    the undivided oracle mesh, for every R, on 3 slabs of a 3D and a 2D Kuhn mesh; the owned sets partition the mesh."""
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ddf_oracle as o
    from importlib import import_module
    dd = import_module("ddf.jl_b200.dist")
    for D, nxy, nz, world in ((3, 2, 14, 3), (2, 4, 12, 3), (3, 4, 8, 2)):
        plane = (nxy + 1) ** (D - 1)
        full = o.large_hypercube_manifold(D, (nxy,) * (D - 1) + (nz,), h_div=nxy)
        slabs = dd.partition_slabs(nz, world, plane)
        for R in range(D + 1):
            ftab = np.arange(full.nsimplices(0))[:, None] if R == 0 else full.simplices[R]
            offset = 0
            for s in slabs:
                simp, coords, weights = o.large_hypercube_tables(D, (nxy,) * (D - 1) + (s.nz_cells,), h_div=nxy)
                m = o.build_manifold("slab", simp, coords, weights)
                tab = np.arange(m.nsimplices(0))[:, None] if R == 0 else m.simplices[R]
                lo, hi = s.own_rows
                own = (tab[:, 0] >= lo) & (tab[:, 0] < hi)
                glob = tab[own] + s.z_cell_lo * plane
                cnt = int(own.sum())
                assert np.array_equal(glob, ftab[offset:offset + cnt]), (D, R, s.rank)
                offset += cnt
            assert offset == ftab.shape[0], (D, R)


# ------------------------------------------------------------------------------------------------------------------
# General meshes partitioned by vertex ranges (ddf.dist.partition_mesh; device side ddf_mesh_set_halo_lists)
# ------------------------------------------------------------------------------------------------------------------
def _general_meshes():
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ddf_oracle as o
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(11)
    pts = rng.random((400, 2))
    pts = pts[np.argsort(pts[:, 0])]                     # a numbering with locality: sorted along x
    yield "delaunay2d", np.sort(Delaunay(pts).simplices.astype(np.int64), axis=1), pts, np.zeros(len(pts))
    pts3 = rng.random((250, 3))
    pts3 = pts3[np.argsort(pts3[:, 2])]
    yield "delaunay3d", np.sort(Delaunay(pts3).simplices.astype(np.int64), axis=1), pts3, np.zeros(len(pts3))
    simp, coords, weights = o.large_hypercube_tables(3, (4, 4, 4))
    yield "kuhn3", simp, coords, weights                 # non-zero weights


def test_general_partition_owned_rows_equal_undivided():
    """Closed star of a vertex range, renumbered monotonically: the owned rows of the boundary mask, star_0 and
    L = laplace(Pr,0) (values AND order of the entries) equal the undivided oracle's bit for bit; the ghost lists
    of a pair match by position; every ghost is received exactly once."""
    import ddf_oracle as o
    from importlib import import_module
    dd = import_module("ddf.jl_b200.dist")
    for name, simp, coords, weights in _general_meshes():
        full = o.build_manifold(name, simp, coords, weights)
        fb, fL = o.wave_operators(full)
        fL = fL.tocsr()
        fL.sort_indices()
        nv = coords.shape[0]
        for world in (2, 3, 5):
            parts = dd.partition_mesh(simp, nv, world)
            assert sum(p.own_hi - p.own_lo for p in parts) == nv
            for p in parts:
                l2g = p.local_to_global
                assert np.array_equal(l2g[p.own_lo:p.own_hi], np.arange(p.cuts[p.rank], p.cuts[p.rank + 1]))
                assert np.array_equal(l2g[p.simplices], simp[p.top_global])
                m = o.build_manifold(name, p.simplices, coords[l2g], weights[l2g])
                b, L = o.wave_operators(m)
                L = L.tocsr()
                L.sort_indices()
                for i in range(p.own_lo, p.own_hi):
                    g = l2g[i]
                    assert b[i] == fb[g]
                    a0, a1 = L.indptr[i], L.indptr[i + 1]
                    g0, g1 = fL.indptr[g], fL.indptr[g + 1]
                    assert np.array_equal(l2g[L.indices[a0:a1]], fL.indices[g0:g1])
                    assert np.array_equal(L.data[a0:a1], fL.data[g0:g1]), (name, world, p.rank, i)
                # ghosts: each exactly once
                nghost = p.nverts - (p.own_hi - p.own_lo)
                assert p.recv_idx.shape[0] == nghost and np.unique(p.recv_idx).shape[0] == nghost
                assert ((p.recv_idx < p.own_lo) | (p.recv_idx >= p.own_hi)).all()
                assert ((p.send_idx >= p.own_lo) & (p.send_idx < p.own_hi)).all()
            for p in parts:
                for k, q in enumerate(p.peers):
                    other = parts[q]
                    kk = list(other.peers).index(p.rank)
                    mine = p.local_to_global[p.send_idx[p.send_ptr[k]:p.send_ptr[k + 1]]]
                    theirs = other.local_to_global[other.recv_idx[other.recv_ptr[kk]:other.recv_ptr[kk + 1]]]
                    assert np.array_equal(mine, theirs)


def _general_worker(rank, world, port, nsteps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ddf_oracle as o
    from importlib import import_module
    dd = import_module("ddf.jl_b200.dist")
    name, simp, coords, weights = next(iter(_general_meshes()))
    part = dd.partition_mesh(simp, coords.shape[0], world, ranks=[rank])[0]
    l2g = part.local_to_global
    m = o.build_manifold(name, part.simplices, coords[l2g], weights[l2g])
    b, L = o.wave_operators(m)
    mask = 1.0 - b.astype(np.float64)
    lo, hi = part.own_lo, part.own_hi
    rng = np.random.default_rng(5)
    u = rng.standard_normal(coords.shape[0])[l2g]
    v = rng.standard_normal(coords.shape[0])[l2g]
    dt = 1e-3
    Lr = L.tocsr()
    Lr.sort_indices()
    for _ in range(nsteps):
        lu = o._segment_sum_sequential(Lr.data * u[Lr.indices], Lr.indptr)
        vn, un = v.copy(), u.copy()
        vn[lo:hi] = v[lo:hi] + dt * (mask[lo:hi] * lu[lo:hi])
        un[lo:hi] = u[lo:hi] + dt * (mask[lo:hi] * vn[lo:hi])
        reqs, bufs = [], []
        for k, peer in enumerate(part.peers):
            s = part.send_idx[part.send_ptr[k]:part.send_ptr[k + 1]]
            r = part.recv_idx[part.recv_ptr[k]:part.recv_ptr[k + 1]]
            if s.size:
                reqs.append(dist.isend(torch.from_numpy(un[s].copy()), int(peer)))
            if r.size:
                buf = torch.empty(r.size, dtype=torch.float64)
                bufs.append((r, buf))
                reqs.append(dist.irecv(buf, int(peer)))
        for r in reqs:
            r.wait()
        for r, buf in bufs:
            un[r] = buf.numpy()
        u, v = un, vn
    q.put((rank, int(part.cuts[rank]), int(part.cuts[rank + 1]), u[lo:hi].copy(), v[lo:hi].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_general_partition_wave_equals_undivided():
    """World-size-2 leapfrog on a vertex-range partition of a 2D Delaunay mesh (gloo send/recv along the ghost lists)
    == the undivided oracle, bit for bit, from a random state."""
    import ddf_oracle as o
    world, nsteps = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_general_worker, args=(r, world, port, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    name, simp, coords, weights = next(iter(_general_meshes()))
    m = o.build_manifold(name, simp, coords, weights)
    rng = np.random.default_rng(5)
    u0 = rng.standard_normal(coords.shape[0])
    v0 = rng.standard_normal(coords.shape[0])
    ou, ov = o.wave_leapfrog(m, u0, v0, 1e-3, nsteps)
    for rank, glo, ghi, u, v in sorted(res):
        assert np.array_equal(u, ou[glo:ghi]), f"rank {rank} u"
        assert np.array_equal(v, ov[glo:ghi]), f"rank {rank} v"


def test_deep_halo_two_steps_per_exchange_model():
    """Host model of the deep halo of csrc/comm.cu (steps_dist_p2p, two leapfrog steps per launch on slabs): a rank runs
    step n on its owned rows PLUS one ghost plane per interior side (the slabs carry two ghost cell layers, so the
    Laplacian rows, masks and geometry of that plane are complete locally), step n+1 on the owned rows, and only then
    exchanges: the two outermost owned planes of u and the outermost plane of v per neighbour.  Lockstep emulation of
    three ranks on the oracle's slab meshes; owned rows must equal the undivided oracle bit for bit, also with a single
    step (one-plane exchange) mixed in."""
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ddf_oracle as o
    from importlib import import_module
    dd = import_module("ddf.jl_b200.dist")
    D, nxy, nz, world = 3, 2, 12, 3
    plane = (nxy + 1) ** (D - 1)
    slabs = dd.partition_slabs(nz, world, plane)
    full = o.large_hypercube_manifold(D, (nxy,) * (D - 1) + (nz,), h_div=nxy)
    rng = np.random.default_rng(8)
    gu, gv = rng.standard_normal(full.nsimplices(0)), rng.standard_normal(full.nsimplices(0))
    dt = 1.0 / (2 * nxy)
    ranks = []
    for s in slabs:
        simp, coords, weights = o.large_hypercube_tables(D, (nxy,) * (D - 1) + (s.nz_cells,), h_div=nxy)
        coords = coords.copy()
        coords[:, -1] += s.z_cell_lo / nxy
        m = o.build_manifold("slab", simp, coords, weights)
        b, L = o.wave_operators(m)
        Lr = L.tocsr()
        Lr.sort_indices()
        g0 = s.z_cell_lo * plane
        n = m.nsimplices(0)
        lo, hi = s.own_rows
        ranks.append(dict(s=s, Lr=Lr, mask=1.0 - b.astype(np.float64), u=gu[g0:g0 + n].copy(), v=gv[g0:g0 + n].copy(),
                          lo=lo, hi=hi, has_lo=s.ghost_lo > 0, has_hi=s.ghost_hi > 0))

    def step(r, a, b_):                                   # one leapfrog step on rows [a, b_)
        lu = o._segment_sum_sequential(r["Lr"].data * r["u"][r["Lr"].indices], r["Lr"].indptr)
        vn, un = r["v"].copy(), r["u"].copy()
        vn[a:b_] = r["v"][a:b_] + dt * (r["mask"][a:b_] * lu[a:b_])
        un[a:b_] = r["u"][a:b_] + dt * (r["mask"][a:b_] * vn[a:b_])
        r["u"], r["v"] = un, vn

    def exchange(planes_u, planes_v):
        snap = [(r["u"].copy(), r["v"].copy()) for r in ranks]
        for k, r in enumerate(ranks):
            lo, hi, P = r["lo"], r["hi"], plane
            if r["has_lo"]:                               # from rank k-1: its top owned planes
                pu, pv = snap[k - 1]
                phi = ranks[k - 1]["hi"]
                r["u"][lo - planes_u * P:lo] = pu[phi - planes_u * P:phi]
                if planes_v:
                    r["v"][lo - P:lo] = pv[phi - P:phi]
            if r["has_hi"]:                               # from rank k+1: its lowest owned planes
                pu, pv = snap[k + 1]
                plo = ranks[k + 1]["lo"]
                r["u"][hi:hi + planes_u * P] = pu[plo:plo + planes_u * P]
                if planes_v:
                    r["v"][hi:hi + P] = pv[plo:plo + P]

    def pair():
        for r in ranks:
            P = plane
            step(r, r["lo"] - (P if r["has_lo"] else 0), r["hi"] + (P if r["has_hi"] else 0))
            step(r, r["lo"], r["hi"])
        exchange(2, 1)

    def single():
        for r in ranks:
            step(r, r["lo"], r["hi"])
        exchange(1, 0)

    nsteps = 0
    exchange(2, 1)                                        # the refresh at the start of a pair segment
    for _ in range(3):
        pair()
        nsteps += 2
    single()
    nsteps += 1
    exchange(2, 1)                                        # a new pair segment after a single step
    pair()
    nsteps += 2
    ou, ov = o.wave_leapfrog(full, gu, gv, dt, nsteps)
    for r in ranks:
        s = r["s"]
        a, b_ = s.z_own_lo * plane, s.z_own_hi * plane
        assert np.array_equal(r["u"][r["lo"]:r["hi"]], ou[a:b_]), s.rank
        assert np.array_equal(r["v"][r["lo"]:r["hi"]], ov[a:b_]), s.rank

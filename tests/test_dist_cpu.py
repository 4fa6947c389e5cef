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

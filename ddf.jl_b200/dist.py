"""Multi-GPU layer: one process per GPU, slabs along the slowest grid axis, ghost
planes exchanged with ncclSend/ncclRecv over NVLink.

The reference has no parallel code at all ("Parallelism strategies:
none"); this partitions the path's independent units (rows = simplices) by
simplex ranges.  Host-side logic here is pure Python so that it can be tested
with the gloo backend on CPU (tests/test_dist_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

GHOST_CELLS = 2   # ghost cell layers per interior slab face: even, so the mirrored Kuhn blocks keep their parity


@dataclass
class Slab:
    """This rank's part of a `large_hypercube` mesh cut along the last axis."""
    rank: int
    nranks: int
    z_own_lo: int          # first owned vertex plane (global index)
    z_own_hi: int          # one past the last owned vertex plane
    z_cell_lo: int         # first local cell layer (global index, even)
    z_cell_hi: int         # one past the last local cell layer
    plane: int             # vertices per plane
    ghost_lo: int          # ghost vertex planes below the owned range
    ghost_hi: int          # ghost vertex planes above the owned range

    @property
    def nz_cells(self) -> int:
        return self.z_cell_hi - self.z_cell_lo

    @property
    def n_local_planes(self) -> int:
        return self.nz_cells + 1

    @property
    def own_rows(self):
        return self.ghost_lo * self.plane, (self.n_local_planes - self.ghost_hi) * self.plane


def partition_slabs(nz_cells: int, nranks: int, plane: int) -> List[Slab]:
    """Cut `nz_cells` cell layers (even) into `nranks` contiguous slabs of even
    thickness.  Rank r owns vertex planes [Z_r, Z_{r+1}); the last rank also owns the
    top plane.  Every interior slab face is padded by GHOST_CELLS cell layers so the
    full star (all incident cells) of every owned vertex is local: geometry and the
    Laplacian rows of owned vertices are then identical to the undivided mesh."""
    if nz_cells % 2:
        raise ValueError("nz_cells must be even (ManifoldConstructors.jl:208)")
    blocks = nz_cells // 2
    if nranks > blocks:
        raise ValueError(f"cannot cut {nz_cells} cell layers into {nranks} slabs of even thickness")
    base, extra = divmod(blocks, nranks)
    cuts = [0]
    for r in range(nranks):
        cuts.append(cuts[-1] + 2 * (base + (1 if r < extra else 0)))
    slabs = []
    for r in range(nranks):
        lo, hi = cuts[r], cuts[r + 1]
        g_lo = GHOST_CELLS if r > 0 else 0
        g_hi = GHOST_CELLS if r < nranks - 1 else 0
        own_hi = hi if r < nranks - 1 else hi + 1           # the last rank owns the top vertex plane
        slabs.append(Slab(rank=r, nranks=nranks, z_own_lo=lo, z_own_hi=own_hi, z_cell_lo=lo - g_lo,
                          z_cell_hi=hi + g_hi, plane=plane, ghost_lo=g_lo,
                          ghost_hi=(g_hi + 1) if r < nranks - 1 else 0))
    return slabs


def init_comm(ctx, group=None):
    """Create the NCCL communicator of `ctx` from an initialised torch.distributed
    process group (any backend; only used to broadcast the 128-byte unique id)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        ctx.rank, ctx.nranks = 0, 1
        return
    if rank == 0:
        uid = ctx.comm_unique_id()
    else:
        uid = bytes(128)
    obj = [uid]
    dist.broadcast_object_list(obj, src=0, group=group)
    ctx.comm_init(world, rank, obj[0])


def slab_hypercube(D: int, nelts: Sequence[int], ctx, rank: Optional[int] = None, nranks: Optional[int] = None,
                   assemble: bool = True):
    """This rank's slab of large_hypercube_manifold(Val(D); nelts) as a device mesh
    with global coordinates and weights (hdiv = nelts[0]; the reference requires all
    nelts equal, ManifoldConstructors.jl:262 -- slabs relax that along the last axis
    only).  Returns (Manifold, Slab)."""
    from .manifoldops import large_hypercube_manifold
    rank = ctx.rank if rank is None else rank
    nranks = ctx.nranks if nranks is None else nranks
    plane = int(np.prod([n + 1 for n in nelts[:-1]])) if D > 1 else 1
    slab = partition_slabs(int(nelts[-1]), nranks, plane)[rank]
    local = list(nelts[:-1]) + [slab.nz_cells]
    mfd = large_hypercube_manifold(D, local, ctx=ctx, hdiv=int(nelts[0]), z0=slab.z_cell_lo, assemble=assemble)
    mfd.set_halo(plane, slab.ghost_lo, slab.ghost_hi)
    return mfd, slab


def halo_mode(mfd) -> str:
    """How the last distributed wave step on `mfd` moved its ghost planes: "p2p" (fused peer-memory stores from
    the row kernel + step flags), "nccl" (ncclSend/ncclRecv fallback) or "none" (not run yet)."""
    import ctypes as C
    from ._lib import call
    mode = C.c_int()
    call("ddf_halo_mode", mfd._h, C.byref(mode))
    return ("none", "p2p", "nccl")[mode.value]


def halo_exchange(mfd, f):
    """Fill the ghost entries of the cochain `f` (any rank R) with the owners' values; in place."""
    from ._lib import call
    call("ddf_halo_exchange_k", mfd._h, f._h, int(f.R))
    return f


def owned_mask(mfd, slab: "Slab", R: int) -> np.ndarray:
    """True for the local R-simplices this rank owns (smallest vertex in an owned plane)."""
    lo, hi = slab.own_rows
    first = np.arange(mfd.nsimplices(0)) if R == 0 else mfd.get_simplices(R)[:, 0]
    return (first >= lo) & (first < hi)


def owned_offsets(mfd, slab: "Slab", R: int, counts: Optional[Sequence[int]] = None):
    """(offset, total): this rank's owned R-simplices are the global ids [offset, offset + count) of the undivided
    mesh (a simplex belongs to the rank owning its smallest vertex; face ids are lexicographic ranks and
    slabs own contiguous vertex ranges, so every rank's share is one contiguous id range in rank order).  `counts`
    (owned count per rank) may be passed by a caller that already gathered them; otherwise they are summed over the
    ranks with the context's all-reduce."""
    mine = int(owned_mask(mfd, slab, R).sum())
    if counts is None:
        counts = [int(round(mfd.ctx.allreduce(float(mine if r == slab.rank else 0), "sum"))) if slab.nranks > 1 else mine
                  for r in range(slab.nranks)]
    return int(sum(counts[:slab.rank])), int(sum(counts))


def global_ids(mfd, slab: "Slab", R: int, counts: Optional[Sequence[int]] = None) -> np.ndarray:
    """Global id (0-based, the numbering `Manifold` gives the undivided mesh, Manifolds.jl:735-757) of every local
    R-simplex.  Owned simplices: offset + rank among the owned ones (the local order is the global order, because the
    local -> global vertex map is monotone); ghosts: their owners' ids, fetched with the rank-R ghost exchange."""
    from .core import Fun, Pr
    own = owned_mask(mfd, slab, R)
    offset, total = owned_offsets(mfd, slab, R, counts)
    if total >= 2 ** 53:
        raise ValueError("global ids beyond 2^53 do not survive the Float64 ghost exchange")
    gid = np.where(own, offset + np.cumsum(own) - 1, -1).astype(np.float64)
    if slab.nranks > 1:
        f = Fun(mfd, Pr, R, gid)
        halo_exchange(mfd, f)
        gid = f.values
    out = gid.astype(np.int64)
    if (out < 0).any():
        raise RuntimeError(f"{int((out < 0).sum())} local {R}-simplices have no owner among the neighbouring slabs")
    return out


# ----------------------------------------------------------------------------------------------------------------
# General meshes (uploaded, refined, Delaunay): partition by contiguous ranges of the global vertex numbering.
# north_star: "large refined meshes are partitioned by simplex ranges".  A simplex of any rank belongs to the rank
# owning its smallest vertex, and face ids are lexicographic ranks (Manifolds.jl:735-757), so contiguous vertex ranges
# give every rank contiguous ranges of simplices of every rank below the top one.
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class MeshPart:
    """One rank's share of a mesh partitioned by vertex ranges (`partition_mesh`)."""
    rank: int
    nranks: int
    cuts: np.ndarray            # nranks + 1 global vertex cuts: rank r owns [cuts[r], cuts[r+1])
    local_to_global: np.ndarray  # ascending global ids of the local vertices (owned and ghost)
    top_global: np.ndarray       # ascending rows of the global top-level table that are local
    simplices: np.ndarray        # local top-level table in LOCAL vertex ids, rows in the global order
    own_lo: int                  # owned local vertices are the contiguous local range [own_lo, own_hi)
    own_hi: int
    peers: np.ndarray            # ranks this one exchanges with, ascending
    send_ptr: np.ndarray         # len(peers) + 1
    send_idx: np.ndarray         # local ids of owned vertices that peers[k] holds as ghosts, ascending global id
    recv_ptr: np.ndarray
    recv_idx: np.ndarray         # local ids of this rank's ghosts owned by peers[k], ascending global id

    @property
    def nverts(self) -> int:
        return int(self.local_to_global.shape[0])


def vertex_cuts(nverts: int, nranks: int) -> np.ndarray:
    """Balanced contiguous vertex ranges."""
    return (np.arange(nranks + 1, dtype=np.int64) * int(nverts)) // int(nranks)


def partition_mesh(simplices: np.ndarray, nverts: int, nranks: int, cuts: Optional[Sequence[int]] = None,
                   ranks: Optional[Sequence[int]] = None) -> List[MeshPart]:
    """Cut a mesh given by its 0-based top-level table (n_D, D+1) into `nranks` parts by vertex ranges.

    Rank r's local mesh is the CLOSED STAR of its owned vertices -- every top-level simplex with at least one owned
    vertex -- renumbered by the monotone map global -> local, rows kept in the global order.  Then every simplex of any
    rank that contains an owned vertex is local together with ALL its cofaces, local face ids keep the global relative
    order (lexicographic ranks under a monotone renumbering), and so the volumes, circumcentric dual volumes (sums
    over ascending parents, Manifolds.jl:1187), star values, boundary flags and rows of L = laplace(Pr, 0) of the
    owned vertices are bit-identical to the undivided mesh's.  Ghost vertices (local, not owned) are refreshed from
    their owners through the index lists returned here; both sides of a pair list the shared vertices in ascending
    global id, so the i-th value sent is the i-th value received.

    The vertex numbering should have locality (a space-filling or generator order); with a random numbering the
    closed stars cover most of the mesh.  `ranks` limits the returned parts (each rank usually wants its own)."""
    simplices = np.sort(np.ascontiguousarray(simplices, dtype=np.int64), axis=1)
    cuts = vertex_cuts(nverts, nranks) if cuts is None else np.asarray(cuts, dtype=np.int64)
    if cuts.shape != (nranks + 1,) or cuts[0] != 0 or cuts[-1] != nverts or (np.diff(cuts) < 0).any():
        raise ValueError("cuts must be nranks + 1 non-decreasing vertex ids from 0 to nverts")
    if simplices.size and (simplices.min() < 0 or simplices.max() >= nverts):
        raise ValueError("simplex table names a vertex outside 0..nverts")
    # local vertex sets of every rank (needed by its neighbours for their send lists)
    local_rows, local_verts = [], []
    for r in range(nranks):
        touch = ((simplices >= cuts[r]) & (simplices < cuts[r + 1])).any(axis=1)
        rows = np.nonzero(touch)[0]
        local_rows.append(rows)
        verts = np.unique(simplices[rows])
        # owned vertices that no simplex touches still belong to the rank (the reference keeps isolated vertices)
        local_verts.append(np.union1d(verts, np.arange(cuts[r], cuts[r + 1], dtype=np.int64)))
    parts = []
    for r in (range(nranks) if ranks is None else ranks):
        l2g = local_verts[r]
        own_lo = int(np.searchsorted(l2g, cuts[r]))
        own_hi = int(np.searchsorted(l2g, cuts[r + 1]))
        ghost_g = np.concatenate([l2g[:own_lo], l2g[own_hi:]])
        ghost_l = np.concatenate([np.arange(own_lo), np.arange(own_hi, l2g.shape[0])]).astype(np.int64)
        owner = np.searchsorted(cuts, ghost_g, side="right") - 1
        # an empty range can share its cut with the next one: the owner is the rank whose range really holds the vertex
        peers, send_ptr, send_idx, recv_ptr, recv_idx = [], [0], [], [0], []
        for p in range(nranks):
            if p == r:
                continue
            rcv = ghost_l[owner == p]                                        # ascending local == ascending global
            mine_on_p = local_verts[p][(local_verts[p] >= cuts[r]) & (local_verts[p] < cuts[r + 1])]
            snd = np.searchsorted(l2g, mine_on_p)
            if rcv.size == 0 and snd.size == 0:
                continue
            peers.append(p)
            send_idx.append(snd)
            recv_idx.append(rcv)
            send_ptr.append(send_ptr[-1] + snd.size)
            recv_ptr.append(recv_ptr[-1] + rcv.size)
        cat = lambda xs: np.concatenate(xs).astype(np.int64) if xs else np.zeros(0, dtype=np.int64)
        parts.append(MeshPart(rank=r, nranks=nranks, cuts=cuts, local_to_global=l2g, top_global=local_rows[r],
                              simplices=np.searchsorted(l2g, simplices[local_rows[r]]), own_lo=own_lo, own_hi=own_hi,
                              peers=np.asarray(peers, dtype=np.int32), send_ptr=np.asarray(send_ptr, dtype=np.int64),
                              send_idx=cat(send_idx), recv_ptr=np.asarray(recv_ptr, dtype=np.int64),
                              recv_idx=cat(recv_idx)))
    return parts


def distribute_mesh(name: str, simplices: np.ndarray, coords: np.ndarray, weights: Optional[np.ndarray], ctx,
                    cuts: Optional[Sequence[int]] = None, use_weighted_duals: bool = True):
    """This rank's part of `Manifold(name, simplices, coords, weights)` as a device mesh with its ghost lists set
    (`ddf_mesh_set_halo_lists`).  Every rank passes the same global tables.  Returns (Manifold, MeshPart);
    `ddf.wave.leapfrog_dist`, `poisson_jacobi_dist`, `tsit5` and `halo_exchange` (0-cochains) then work on it exactly
    as on a generator slab, and `Manifold.owned_rows(R)` gives the owned row range of every rank below the top one."""
    from .core import Manifold
    part = partition_mesh(simplices, coords.shape[0], ctx.nranks, cuts, ranks=[ctx.rank])[0]
    w = None if weights is None else np.asarray(weights, dtype=np.float64)[part.local_to_global]
    mfd = Manifold.from_simplices(name, part.simplices, np.asarray(coords, dtype=np.float64)[part.local_to_global], w,
                                  ctx=ctx, use_weighted_duals=use_weighted_duals)
    mfd.set_halo_lists(part)
    return mfd, part

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

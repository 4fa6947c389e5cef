"""Independent restatement of the hot path in torch tensor algebra (test infrastructure, shared by the CPU and GPU tests).

Written from the reference's definitions (src/Manifolds.jl:717-768, src/ManifoldOps.jl:17-146, Julia SparseArrays'
`mul!` loop orders) WITHOUT using oracle/ or the product: tests/test_independent_cpu.py holds it against the numpy
oracle on small meshes (so the two restatements pin each other), tests/test_gpu_fullsize.py uses it on the tables the
device exports at the BASELINE sizes, where the numpy oracle is too slow.  Works on any torch device (DEVICE).
"""
import math

import numpy as np

DEVICE = "cpu"


def _torch():
    import torch
    return torch


def _dev(a, dtype=None):
    """numpy array -> cuda tensor (int32 for index tables to halve the footprint)."""
    torch = _torch()
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEVICE)
    return t if dtype is None else t.to(dtype)


class Tables:
    """Everything the checks need, exported through the C ABI (or by the oracle, same storage) once per mesh and held as
    torch tensors.  `m` offers nsimplices(R), get_simplices_csc(R), get_boundaries_csc(R) in the reference's storage
    (1-based Int64 CSC)."""

    def __init__(self, m):
        torch = _torch()
        self.D = m.D
        self.n = [m.nsimplices(R) for R in range(m.D + 1)]
        self.S = {0: torch.arange(self.n[0], device=DEVICE, dtype=torch.int32)[:, None]}
        self.B, self.Bsign = {}, {}
        for R in range(1, m.D + 1):
            colptr, rowval = m.get_simplices_csc(R)
            assert colptr[0] == 1 and np.array_equal(np.diff(colptr[:1024]), np.full(min(1023, len(colptr) - 1), R + 1))
            assert colptr[-1] == 1 + (R + 1) * self.n[R]
            self.S[R] = _dev(rowval, _torch().int64).sub_(1).to(torch.int32).view(-1, R + 1)
            del colptr, rowval
            colptr, rowval, nzval = m.get_boundaries_csc(R)
            cp = _dev(colptr)
            want = torch.arange(self.n[R] + 1, device=DEVICE, dtype=torch.int64) * (R + 1) + 1
            assert torch.equal(cp, want), f"boundary colptr R={R}"      # every column holds exactly R+1 entries
            del cp, want, colptr
            self.B[R] = _dev(rowval, torch.int64).sub_(1).to(torch.int32).view(-1, R + 1)
            self.Bsign[R] = _dev(nzval).view(-1, R + 1)
            del rowval, nzval
            torch.cuda.empty_cache()


def lex_strictly_increasing(T):
    """rows of the (n, w) table strictly increasing in lexicographic order"""
    torch = _torch()
    if T.shape[0] < 2:
        return True
    a, b = T[:-1], T[1:]
    lt = torch.zeros(a.shape[0], dtype=torch.bool, device=T.device)
    eq = torch.ones(a.shape[0], dtype=torch.bool, device=T.device)
    for c in range(T.shape[1]):
        lt |= eq & (a[:, c] < b[:, c])
        eq &= a[:, c] == b[:, c]
    return bool(lt.all().item())


def prove_assembly(tb):
    """The structural proof of A1 (see the module docstring).  Returns nothing; asserts."""
    torch = _torch()
    D = tb.D
    for R in range(D, 0, -1):
        S, F, B, sg = tb.S[R], tb.S[R - 1], tb.B[R], tb.Bsign[R]
        nR, nF = tb.n[R], tb.n[R - 1]
        # vertices of every simplex strictly ascending (the sparse-column storage of the reference)
        if R >= 1:
            assert bool((S[:, 1:] > S[:, :-1]).all().item()), f"simp[{R}] rows not ascending"
        # the face table is sorted and duplicate-free: ids are lexicographic ranks among unique faces
        if R - 1 >= 1:
            assert lex_strictly_increasing(F), f"simp[{R - 1}] not strictly lexicographically increasing"
        # rows ascending inside every column (SparseMatrixCSC invariant)
        assert bool((B[:, 1:] > B[:, :-1]).all().item()), f"boundary[{R}] rows not ascending in a column"
        assert int(B.min().item()) >= 0 and int(B.max().item()) < nF
        # entry p of column j is the face S[j] minus ONE vertex q; sign (-1)^q (Manifolds.jl:730-732: the face that
        # leaves out the n-th vertex, n 1-based, has parity (-1)^(n-1))
        omitted_sum = torch.zeros(nR, dtype=torch.int32, device=DEVICE)
        omitted_mask = torch.zeros(nR, dtype=torch.int32, device=DEVICE)
        for p in range(R + 1):
            T = F[B[:, p].long()]                                     # (nR, R) tuple of the named face
            found = torch.zeros(nR, dtype=torch.bool, device=DEVICE)
            for q in range(R + 1):
                keep = [c for c in range(R + 1) if c != q]
                match = (S[:, keep] == T).all(dim=1)
                want_sign = 1 if q % 2 == 0 else -1
                assert bool((sg[match, p] == want_sign).all().item()), f"sign of boundary[{R}] entry {p} (omits {q})"
                found |= match
                omitted_sum += match.to(torch.int32) * q
                omitted_mask += match.to(torch.int32) << q
            assert bool(found.all().item()), f"boundary[{R}] entry {p} is not a face of its column's simplex"
            del T, found
        # the R+1 entries of a column omit R+1 DIFFERENT vertices: all faces of the simplex are present
        assert bool((omitted_mask == (1 << (R + 1)) - 1).all().item()), f"boundary[{R}]: a column misses a face"
        # every face is the face of something (the table holds nothing but faces)
        seen = torch.zeros(nF, dtype=torch.bool, device=DEVICE)
        seen[B.reshape(-1).long()] = True
        assert bool(seen.all().item()), f"simp[{R - 1}] holds a tuple that is no face of any {R}-simplex"
        del seen, omitted_sum, omitted_mask
        torch.cuda.empty_cache()


def deriv_apply(tb, k, x):
    """y = adjoint(boundaries[k+1]) * x in SparseArrays order (ManifoldOps.jl:46; per column: tmp = 0;
    tmp += nzval[j]' * x[rowval[j]] for j ascending; y = 0 + tmp)."""
    torch = _torch()
    B, sg = tb.B[k + 1], tb.Bsign[k + 1]
    tmp = torch.zeros(B.shape[0], dtype=torch.float64, device=DEVICE)
    for p in range(k + 2):
        tmp = tmp + sg[:, p].to(torch.float64) * x[B[:, p].long()]
    return 0.0 + tmp


class RowLists:
    """Row-major view of boundaries[R] (n_{R-1} rows): entries of a row in ascending column order -- what the CSC
    scatter loop `for col: y[rowval[j]] += nzval[j] * x[col]` adds into y[row], in that order."""

    def __init__(self, tb, R):
        torch = _torch()
        B, sg = tb.B[R], tb.Bsign[R]
        nR, nF = tb.n[R], tb.n[R - 1]
        rows = B.reshape(-1)
        # CSC order = column-major: stable sort by row keeps the columns ascending inside a row
        order = torch.sort(rows.long(), stable=True).indices
        self.col = (order // (R + 1)).to(torch.int32)
        self.sign = sg.reshape(-1)[order].to(torch.float64)
        del order
        cnt = torch.bincount(rows.long(), minlength=nF)
        self.len = cnt.to(torch.int32)
        self.start = torch.cumsum(cnt, 0) - cnt
        self.maxlen = int(cnt.max().item())
        self.nrows = nF


def boundary_apply(rl, x):
    """y = boundaries[R] * x (plain CSC mul!): fill!(y, 0); columns ascending; y[row] += nzval * x[col]."""
    torch = _torch()
    y = torch.zeros(rl.nrows, dtype=torch.float64, device=DEVICE)
    for p in range(rl.maxlen):
        act = torch.nonzero(rl.len > p).squeeze(1)
        e = rl.start[act] + p
        y[act] = y[act] + rl.sign[e] * x[rl.col[e].long()]
    return y


def laplace_rows(tb, rl1, ih0, star1):
    """L = laplace(Pr,0) = coderiv(Pr,1) * deriv(Pr,0) (ManifoldOps.jl:110-144) entry by entry with the reference's
    roundings: delta[i,e] = fl(fl(-ih0[i] * d1[i,e]) * star1[e]) (left to right, the first product exact),
    L[i,j] = delta[i,e] * d0[e,j] (exact, +-1), L[i,i] = the spmatmul accumulation over the incident edges in
    ascending order (first touch assigns).  Returns per (vertex, incident edge) entry: neighbour, value; and the
    diagonal."""
    torch = _torch()
    E = tb.S[1]
    e = rl1.col.long()                                     # incident edge of every (row-major) entry
    rows = torch.repeat_interleave(torch.arange(rl1.nrows, device=DEVICE), rl1.len.long())
    other = torch.where(E[e, 0].long() == rows, E[e, 1].long(), E[e, 0].long())
    l = ih0[rows] * star1[e]                               # |delta[i,e]| = L[i,j], one rounding
    assert bool((l * l >= math.sqrt(np.finfo(np.float64).eps ** 3)).all().item()), "an entry would be chopped (Ops.jl:52-58)"
    diag = torch.zeros(rl1.nrows, dtype=torch.float64, device=DEVICE)
    for p in range(rl1.maxlen):
        act = torch.nonzero(rl1.len > p).squeeze(1)
        q = rl1.start[act] + p
        diag[act] = (-l[q]) if p == 0 else diag[act] + (-l[q])
    nlower = torch.zeros(rl1.nrows, dtype=torch.int32, device=DEVICE)
    nlower.index_add_(0, rows, (other < rows).to(torch.int32))
    return other, l, diag, nlower


def laplace_apply(rl1, other, l, diag, nlower, u):
    """y = L * u with L in CSC: per row the products in ascending column order = lower neighbours, the diagonal,
    upper neighbours (edge ids are lexicographic, so incident edges ascend with the neighbour)."""
    torch = _torch()
    y = torch.zeros(rl1.nrows, dtype=torch.float64, device=DEVICE)
    idx = torch.arange(rl1.nrows, device=DEVICE)
    for p in range(rl1.maxlen + 1):
        # position p of row i: entry p if p < nlower, the diagonal if p == nlower, entry p-1 after it
        is_diag = nlower == p
        act_d = torch.nonzero(is_diag).squeeze(1)
        y[act_d] = y[act_d] + diag[act_d] * u[act_d]
        ent = torch.where(nlower > p, p, p - 1)
        act = torch.nonzero((~is_diag) & (ent < rl1.len) & (ent >= 0)).squeeze(1)
        q = rl1.start[act] + ent[act]
        y[act] = y[act] + l[q] * u[other[q]]
    del idx
    return y


def boundary_vertex_mask(tb):
    """isboundary(Pr,0) (Manifolds.jl:510-550): a (D-1)-face is on the boundary iff it appears an odd number of
    times in the columns of boundaries[D]; lower ranks = the subsimplices of those faces."""
    torch = _torch()
    D = tb.D
    cnt = torch.bincount(tb.B[D].reshape(-1).long(), minlength=tb.n[D - 1])
    bface = torch.nonzero(cnt % 2 == 1).squeeze(1)
    mask = torch.zeros(tb.n[0], dtype=torch.bool, device=DEVICE)
    mask[tb.S[D - 1][bface].reshape(-1).long()] = True
    return mask



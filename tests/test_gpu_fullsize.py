"""Oracle-INDEPENDENT proofs at BASELINE sizes (config 4: 3D Kuhn n=256, 1.0066e8 tetrahedra; config 3: 2D n=2048).

The oracle (numpy) finishes n=64 in seconds but not n=256, so at full size the product is checked against a SECOND,
independent restatement written here in torch tensor algebra from the reference's definitions -- it shares no code with
oracle/ or with the product's kernels and works only from what the C ABI exports in the reference's own storage
(`_simplices[R]`, `_boundaries[R]` as 1-based Int64 CSC, volumes, dual volumes):

  * A1 (Manifolds.jl:717-768) is PROVEN, not sampled: the face table of every rank is strictly lexicographically
    increasing (=> unique and sorted), every boundary entry names the face whose vertex tuple is the parent's tuple
    minus one vertex, with the sign (-1)^(position of the omitted vertex) (:730-732), every column holds all its R+1
    faces in ascending row order, and every face is referenced.  Those properties determine colptr / rowval / nzval
    uniquely, so the exported matrices ARE the reference's.
  * A4: y = d_k x (adjoint-CSC loop), y = boundary_k x (CSC scatter loop), y = star_k x, L = laplace(Pr,0) and the
    fused leapfrog step are re-evaluated in Julia's SparseArrays order (one rounding per multiply and per add, entries
    in ascending order) and compared BIT FOR BIT on random Float64 cochains.

torch is used as plumbing for big-array arithmetic only (elementwise IEEE ops, gathers, stable sorts).
"""
import math

import numpy as np
import pytest

import independent as ind
from independent import (RowLists, Tables, _dev, _torch, boundary_apply, boundary_vertex_mask, deriv_apply,
                         laplace_apply, laplace_rows, prove_assembly)

pytestmark = pytest.mark.gpu


def run_full_checks(ddf, ctx, D, n, steps=4):
    torch = _torch()
    ind.DEVICE = "cuda"
    m = ddf.large_hypercube_manifold(D, (n,) * D, ctx=ctx)
    tb = Tables(m)
    # ---- A1: structural proof
    prove_assembly(tb)
    # ---- A4: d_k x bit for bit
    g = torch.Generator(device="cuda").manual_seed(1234 + n)
    for k in range(D):
        x = torch.rand(tb.n[k], dtype=torch.float64, device="cuda", generator=g) - 0.5
        want = deriv_apply(tb, k, x).cpu().numpy()
        got = (ddf.deriv(ddf.Pr, k, m) * ddf.Fun(m, ddf.Pr, k, x.cpu().numpy())).values
        assert np.array_equal(got, want), f"d_{k} x at full size"
        del x, want, got
    # ---- A4: boundary_R x / dual derivative (CSC scatter order) bit for bit
    rls = {}
    for R in range(1, D + 1):
        rl = RowLists(tb, R)
        x = torch.rand(tb.n[R], dtype=torch.float64, device="cuda", generator=g) - 0.5
        want = boundary_apply(rl, x).cpu().numpy()
        xf = x.cpu().numpy()
        got = (ddf.boundary(ddf.Pr, R, m) * ddf.Fun(m, ddf.Pr, R, xf)).values
        assert np.array_equal(got, want), f"boundary_{R} x at full size"
        got = (ddf.deriv(ddf.Dl, R, m) * ddf.Fun(m, ddf.Dl, R, xf)).values
        assert np.array_equal(got, want), f"deriv(Dl,{R}) x at full size"
        if R == 1:
            rls[1] = rl
        del x, want, got, xf
        torch.cuda.empty_cache()
    # ---- A8: star_R x = (dualvol ./ vol) .* x bit for bit from the exported geometry
    vol = {R: _dev(m.get_volumes(R)) for R in range(D + 1)}
    dvol = {R: _dev(m.get_dualvolumes(R)) for R in range(D + 1)}
    for R in range(D + 1):
        x = torch.rand(tb.n[R], dtype=torch.float64, device="cuda", generator=g) - 0.5
        want = ((dvol[R] / vol[R]) * x).cpu().numpy()
        got = (ddf.hodge(ddf.Pr, R, m) * ddf.Fun(m, ddf.Pr, R, x.cpu().numpy())).values
        assert np.array_equal(got, want), f"star_{R} x at full size"
        del x, want, got
    # closed forms of the Kuhn mesh: simplex volumes h^R * sqrt(#unit steps..)/R! are the same for every top simplex
    h = 1.0 / n
    vtop = vol[D]
    assert float((vtop - h ** D / math.factorial(D)).abs().max().item()) <= 1e-12 * h ** D
    assert abs(float(vtop.sum().item()) - 1.0) <= 1e-10 and abs(float(dvol[0].sum().item()) - 1.0) <= 1e-10
    # ---- A9/A10: L rows, L u and the fused leapfrog step bit for bit
    ih0 = vol[0] / dvol[0]
    star1 = dvol[1] / vol[1]
    other, l, diag, nlower = laplace_rows(tb, rls[1], ih0, star1)
    u = torch.rand(tb.n[0], dtype=torch.float64, device="cuda", generator=g) - 0.5
    v = torch.rand(tb.n[0], dtype=torch.float64, device="cuda", generator=g) - 0.5
    want = laplace_apply(rls[1], other, l, diag, nlower, u).cpu().numpy()
    got = (ddf.laplace(ddf.Pr, 0, m) * ddf.Fun(m, ddf.Pr, 0, u.cpu().numpy())).values
    assert np.array_equal(got, want), "L u at full size"
    bmask = boundary_vertex_mask(tb)
    assert np.array_equal(m.get_isboundary(0), bmask.cpu().numpy()), "isboundary(Pr,0) at full size"
    mk = (~bmask).to(torch.float64)
    dt = ddf.wave_dt(m)
    assert dt == float(vol[1].min().item()) / 2                       # wave.jl:106,118
    fu, fv = ddf.Fun(m, ddf.Pr, 0, u.cpu().numpy()), ddf.Fun(m, ddf.Pr, 0, v.cpu().numpy())
    ddf.wave_leapfrog(m, fu, fv, dt, steps)
    for _ in range(steps):
        v = v + dt * (mk * laplace_apply(rls[1], other, l, diag, nlower, u))
        u = u + dt * (mk * v)
    assert np.array_equal(fu.values, u.cpu().numpy()) and np.array_equal(fv.values, v.cpu().numpy()), "leapfrog at full size"
    # ---- N1: fixed-step Tsit5 (examples/wave.jl:134) restated here stage by stage in perform_step!'s order -- stage
    # broadcasts with separately rounded products and sums, left to right, FSAL across steps -- against ddf_wave_tsit5,
    # which at this size runs THREE paired-stage passes over L per step (csrc/wave.cu tsit5_pair): bit for bit
    A = ((0.161,),
         (-0.008480655492356989, 0.335480655492357),
         (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
         (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
         (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
         (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774))

    def rhs(U):
        return mk * U[1], mk * laplace_apply(rls[1], other, l, diag, nlower, U[0])

    def broadcast(base, coefs, ks):
        acc = coefs[0] * ks[0]
        for c, kk in zip(coefs[1:], ks[1:]):
            acc = acc + c * kk
        return base + dt * acc

    tu = torch.rand(tb.n[0], dtype=torch.float64, device="cuda", generator=g) - 0.5
    tv = torch.rand(tb.n[0], dtype=torch.float64, device="cuda", generator=g) - 0.5
    gu, gv = ddf.Fun(m, ddf.Pr, 0, tu.cpu().numpy()), ddf.Fun(m, ddf.Pr, 0, tv.cpu().numpy())
    tsteps = 2
    ddf.wave_tsit5(m, gu, gv, dt, tsteps)
    U = (tu, tv)
    ks = [rhs(U)]
    for _ in range(tsteps):
        for st in range(6):
            tmp = tuple(broadcast(U[c], A[st], [kk[c] for kk in ks]) for c in range(2))
            ks.append(rhs(tmp))
        U = tmp
        ks = [ks[6]]
    assert np.array_equal(gu.values, U[0].cpu().numpy()) and np.array_equal(gv.values, U[1].cpu().numpy()), "Tsit5 at full size"
    del tu, tv, U, ks, tmp
    del tb, rls, vol, dvol, other, l, diag, nlower, u, v
    torch.cuda.empty_cache()
    m.close()


def test_independent_restatement_small_3d(ddf, ctx):
    """The torch restatement itself on a mesh the oracle also covers (tests/test_gpu_parity.py compares the same
    device results with the oracle there): 3D Kuhn n=16."""
    run_full_checks(ddf, ctx, 3, 16)


def test_independent_restatement_small_2d(ddf, ctx):
    run_full_checks(ddf, ctx, 2, 64)


def test_config3_2d_full_size_values(ddf, ctx):
    """BASELINE config 3 (2D Kuhn n=2048: 4,198,401 / 12,587,008 / 8,388,608) by VALUE, not by identities."""
    run_full_checks(ddf, ctx, 2, 2048)


def test_config4_3d_full_size_values(ddf, ctx):
    """BASELINE config 4 (3D Kuhn n=256: 16,974,593 / 118,031,104 / 201,719,808 / 100,663,296): the assembly is
    proven structurally and every apply on the path plus the fused wave step is compared bit for bit with the
    independent restatement.  Exercises at size: vertex keys above 2^24, the 24-bit prefix offsets, the `bits64` d_2
    words, 513^2... (257^2) planes."""
    n = 256
    run_full_checks(ddf, ctx, 3, n)

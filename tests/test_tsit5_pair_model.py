"""CPU tier: host model of the paired Tsit5 stages of csrc/wave.cu (tsit5_pair / k_rows_pipe<5,2>).

Claim: in dU = F U with F = [0, M; M L, 0] (examples/wave.jl:98-100) the stage input su_l of the fixed-step Tsit5 of
examples/wave.jl:134 depends on k^v_1 .. k^v_{l-2} only -- not on k^v_{l-1} -- so stages J and J+1 can be evaluated from
ONE pass over L, a step is three passes (1,2), (3,4), (5,6), and when every stage input is formed with the operations
of the unfused sequence in the same order the trajectory is identical bit for bit.  The model below restates exactly
what the device epilogue does (per row: recompute sv_l, k^u_l from the base state and the stored k^v) in numpy and
compares it with the oracle's restatement of OrdinaryDiffEq's perform_step! (oracle/ddf_oracle.py tsit5_steps)."""
import numpy as np
import pytest

import ddf_oracle as o

A = o.TSIT5_A


def _lin(base, dt, coefs, ks):
    s = coefs[0] * ks[0]
    for c, k in zip(coefs[1:], ks[1:]):
        s = s + c * k
    return base + dt * s


def paired_tsit5(L_apply, m, u, v, dt, nsteps):
    """Three passes over L per step; L_apply(x) = L x with the oracle's summation order; m = 1 - boundary mask."""
    u, v = u.copy(), v.copy()
    su2 = u + dt * (A[0][0] * (m * v))                     # k_tsit5_first
    for _ in range(nsteps):
        kv = [None] * 6
        ku = [None] * 6

        def sv(l):                                         # stage input of udot, l = 2..7: needs kv_1 .. kv_{l-1}
            return _lin(v, dt, A[l - 2], kv[:l - 1])

        def su(l):                                         # stage input of u, l = 2..7: needs ku_1 .. ku_{l-1}
            return _lin(u, dt, A[l - 2], ku[:l - 1])

        ku[0] = m * v
        # pass (1,2): gathers u and su_2
        kv[0], kv[1] = m * L_apply(u), m * L_apply(su2)
        ku[1], ku[2] = m * sv(2), m * sv(3)
        su3, su4 = su(3), su(4)
        # pass (3,4)
        kv[2], kv[3] = m * L_apply(su3), m * L_apply(su4)
        ku[3], ku[4] = m * sv(4), m * sv(5)
        su5, su6 = su(5), su(6)
        # pass (5,6): new state, su_2 of the next step
        kv[4], kv[5] = m * L_apply(su5), m * L_apply(su6)
        ku[5] = m * sv(6)
        un, vn = su(7), sv(7)
        su2 = un + dt * (A[0][0] * (m * vn))
        u, v = un, vn
    return u, v


@pytest.mark.parametrize("D,n", [(2, 8), (3, 4)])
def test_paired_stages_equal_perform_step(D, n):
    mfd = o.large_hypercube_manifold(D, (n,) * D)
    b, L = o.wave_operators(mfd)
    m = 1.0 - b.astype(np.float64)
    rng = np.random.default_rng(D * 10 + n)
    u0, v0 = rng.standard_normal(mfd.nsimplices(0)), rng.standard_normal(mfd.nsimplices(0))
    dt = float(np.min(o.calc_volumes(mfd, 1))) / 2
    for nsteps in (1, 3):
        ou, ov = o.wave_tsit5(mfd, u0, v0, dt, nsteps)
        pu, pv = paired_tsit5(lambda x: o.spmv_csc(L, x), m, u0, v0, dt, nsteps)
        assert np.array_equal(pu, ou) and np.array_equal(pv, ov), (D, n, nsteps)


def test_stage_input_does_not_depend_on_previous_kv():
    """The structural fact itself: perturbing k^v_{l-1} leaves su_l unchanged (and changes sv_l)."""
    rng = np.random.default_rng(3)
    nv = 50
    u, v, m = rng.standard_normal(nv), rng.standard_normal(nv), (rng.random(nv) > 0.2).astype(np.float64)
    dt = 0.01
    kv = [rng.standard_normal(nv) for _ in range(6)]

    def inputs(kvs, l):
        ku = [m * v]
        for q in range(2, l):
            ku.append(m * _lin(v, dt, A[q - 2], kvs[:q - 1]))
        return _lin(u, dt, A[l - 2], ku[:l - 1]), _lin(v, dt, A[l - 2], kvs[:l - 1])

    for l in range(3, 8):
        su_a, sv_a = inputs(kv, l)
        kv2 = list(kv)
        kv2[l - 2] = kv2[l - 2] + 1.0                       # k^v_{l-1}
        su_b, sv_b = inputs(kv2, l)
        assert np.array_equal(su_a, su_b) and not np.array_equal(sv_a, sv_b), l

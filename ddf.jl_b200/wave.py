"""Scalar wave equation driver, mirroring `Wave.wave(Val(D), levels)` of
examples/wave.jl:25-170 on the device.

The reference integrates dU = F U with fixed-step Tsit5 (wave.jl:134).  This
module provides (i) the drop-in right-hand side `wave_rhs` (= `mul!(dU, F, U)`,
wave.jl:98-100) and (ii) the fused leapfrog step north_star asks for, on the same
operators (L, 1-B) and the same dt = min edge length / 2 (wave.jl:106,118).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from ._lib import call
from .core import Fun, Manifold, Pr
from .manifoldops import large_hypercube_manifold, sample


def wave_dt(mfd: Manifold) -> float:
    mfd._geometry()
    out = C.c_double()
    call("ddf_wave_dt", mfd._h, C.byref(out))
    return out.value


def wave_rhs(mfd: Manifold, u: Fun, udot: Fun):
    """(du, dudot) = F * [u; udot]  (examples/wave.jl:87-100)."""
    mfd._geometry()
    du, dudot = Fun(mfd, Pr, 0), Fun(mfd, Pr, 0)
    call("ddf_wave_rhs", mfd._h, u._h, udot._h, du._h, dudot._h)
    return du, dudot


def wave_leapfrog(mfd: Manifold, u: Fun, v: Fun, dt: float, nsteps: int):
    """In-place fused leapfrog: v += dt (1-b).(L u); u += dt (1-b).v, `nsteps` times."""
    mfd._geometry()
    if mfd.ctx.nranks > 1:
        call("ddf_wave_leapfrog_dist", mfd._h, u._h, v._h, float(dt), int(nsteps))
    else:
        call("ddf_wave_leapfrog", mfd._h, u._h, v._h, float(dt), int(nsteps))
    return u, v


def wave_tsit5(mfd: Manifold, u: Fun, udot: Fun, dt: float, nsteps: int):
    """In-place fixed-step Tsit5 on dU = F U -- the integrator of examples/wave.jl:134."""
    mfd._geometry()
    call("ddf_wave_tsit5", mfd._h, u._h, udot._h, float(dt), int(nsteps))
    return u, udot


def lincomb(base: Fun, dt: float, coefs, ks, out: Fun = None) -> Fun:
    """out = base + dt*(c1 k1 + c2 k2 + ...) in one pass (the Runge-Kutta stage broadcast)."""
    out = out if out is not None else base._like()
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    arr = (C.c_void_p * len(ks))(*[k._h for k in ks])
    call("ddf_vec_lincomb", out._h, base._h, float(dt), len(ks), coefs.ctypes.data_as(C.POINTER(C.c_double)), arr)
    return out


def wave_format(mfd: Manifold) -> str:
    """Storage of the assembled L rows the row kernels use on this mesh: "staged" (TMA-staged gather,
    16-bit slots, jagged slices), "sell" (explicit-index SELL-32-256) or "csr"."""
    mfd._geometry()
    f = C.c_int()
    call("ddf_wave_format", mfd._h, C.byref(f))
    return ("csr", "sell", "staged")[f.value]


def wave_step_bytes(mfd: Manifold):
    a, b = C.c_int64(), C.c_int64()
    call("ddf_wave_step_bytes", mfd._h, C.byref(a), C.byref(b))
    return a.value, b.value


def wave(D: int, levels: int, ctx=None, t1: float = 0.875, integrator: str = "leapfrog"):
    """`Wave.wave(Val(D), levels)` with the leapfrog integrator: mesh n = 2^levels
    (wave.jl:32-34), u0 = prod sinpi(x_d), udot0 = 0 (wave.jl:42-58), integrate to
    t1 (the last `saveat` time of wave.jl:116 is 0.875) and return the error norm
    sqrt(|u - u_exact|^2 / length) of wave.jl:156-163."""
    n = 2 ** levels
    mfd = large_hypercube_manifold(D, (n,) * D, ctx=ctx)
    u = sample(mfd, "wave")
    v = Fun(mfd, Pr, 0)
    dt = wave_dt(mfd)
    nsteps = int(round(t1 / dt))
    omega = math.sqrt(D)
    if integrator == "tsit5":            # what examples/wave.jl:134 runs
        wave_tsit5(mfd, u, v, dt, nsteps)
        t = nsteps * dt
        exact = math.cos(math.pi * omega * t) * np.prod(np.sin(np.pi * mfd.get_coords()), axis=1)
        e = u.values - exact
        return float(np.sqrt(np.sum(e * e) / len(e))), mfd
    # leapfrog staggering: the step is v^{n+1/2} = v^{n-1/2} + dt L u^n; u^{n+1} = u^n + dt v^{n+1/2},
    # so start from v^{-1/2} = v0 - dt/2 (1-b) L u0 (second order overall)
    _, dv0 = wave_rhs(mfd, u, v)
    v = v - dv0 * (dt / 2)
    wave_leapfrog(mfd, u, v, dt, nsteps)
    t = nsteps * dt
    coords = mfd.get_coords()
    exact = math.cos(math.pi * omega * t) * np.prod(np.sin(np.pi * coords), axis=1)
    e = u.values - exact
    return float(np.sqrt(np.sum(e * e) / len(e))), mfd

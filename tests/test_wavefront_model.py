"""Host model of the PLANNED two-step wavefront of the fused wave step: persistent blocks walk the
windows in order; iteration `it` of block b runs the step-n task of window it*B + b and then the step-(n+1) task of the
window `lag` iterations back.  With the u ping-pong (step n: U0 -> U1, step n+1: U1 -> U0) a step-(n+1) task of window w
may start once every step-n window <= w + P has finished (P = windows a row's neighbours reach); blocks publish one
counter per iteration.  The model runs the schedule on real threads with random delays and checks (a) the result equals
two sequential steps bit for bit, (b) the smallest deadlock-free lag is ceil(P / B) + 1 iterations, (c) the same with the
kernel's block structure (one producer, a ring of stages, two consumer groups that finish out of order -- the reason the
producer checks nstages counters, csrc/wavefront2.cuh).  No GPU needed."""
import random
import threading
import time

import numpy as np
import pytest


def step_rows(uin, v, w, P, coef, dt):
    """One leapfrog step on window w (a window = one row here): v += dt * sum_j coef_j u[w + j]; u' = u + dt * v."""
    W = len(uin)
    acc = 0.0
    for j in range(-P, P + 1):
        if 0 <= w + j < W:
            acc = acc + coef[j + P] * uin[w + j]
    vn = v[w] + dt * acc
    return uin[w] + dt * vn, vn


def sequential(u, v, P, coef, dt):
    for _ in range(2):
        un, vn = u.copy(), v.copy()
        for w in range(len(u)):
            un[w], vn[w] = step_rows(u, v, w, P, coef, dt)
        u, v = un, vn
    return u, v


def wavefront(u, v, P, coef, dt, B, lag, seed, timeout=20.0):
    W = len(u)
    nit = -(-W // B)
    U0, U1, V = u.copy(), np.zeros_like(u), v.copy()
    done = [0] * (nit + 1)
    active = [min(B, W - it * B) for it in range(nit)]
    lock = threading.Lock()
    deadline = time.time() + timeout
    failed = []

    def block(b):
        rng = random.Random(seed * 1000 + b)
        for it in range(nit + lag):
            wa = it * B + b
            if it < nit and wa < W:                         # step n: reads U0 (everything), writes U1[wa], V[wa]
                U1[wa], V[wa] = step_rows(U0, V, wa, P, coef, dt)
                time.sleep(rng.random() * 2e-4)
                with lock:
                    done[it] += 1
            wb = (it - lag) * B + b
            if it >= lag and wb < W:                        # step n+1: needs step n finished on every window <= wb + P
                need = min(nit - 1, (min(W - 1, wb + P)) // B)
                while True:
                    with lock:
                        ok = done[need] == active[need]
                    if ok:
                        break
                    if time.time() > deadline:
                        failed.append(("deadlock", b, it, need))
                        return
                    time.sleep(1e-5)
                U0[wb], V[wb] = step_rows(U1, V, wb, P, coef, dt)
                time.sleep(rng.random() * 2e-4)
    threads = [threading.Thread(target=block, args=(b,)) for b in range(B)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return U0, V, failed


@pytest.mark.parametrize("W,B,P", [(64, 4, 5), (50, 7, 3), (40, 8, 9)])
def test_two_step_wavefront_schedule(W, B, P):
    rng = np.random.default_rng(W + B + P)
    u, v = rng.standard_normal(W), rng.standard_normal(W)
    coef = rng.standard_normal(2 * P + 1)
    want_u, want_v = sequential(u, v, P, coef, 0.01)
    lag = -(-P // B) + 1                                    # ceil(P / B) + 1 iterations
    for seed in range(3):
        got_u, got_v, failed = wavefront(u, v, P, coef, 0.01, B, lag, seed)
        assert not failed, failed
        assert np.array_equal(got_u, want_u) and np.array_equal(got_v, want_v)


def test_wavefront_needs_its_lag():
    """With no lag a block waits for step-n windows it has not run itself: the schedule cannot finish."""
    rng = np.random.default_rng(1)
    W, B, P = 32, 4, 5
    u, v = rng.standard_normal(W), rng.standard_normal(W)
    coef = rng.standard_normal(2 * P + 1)
    _, _, failed = wavefront(u, v, P, coef, 0.01, B, 0, 0, timeout=1.0)
    assert failed and failed[0][0] == "deadlock"


def wavefront_groups(u, v, P, coef, dt, B, lag, seed, nstages=3, NG=2, timeout=30.0):
    """The same schedule with the kernel's block structure: per block ONE producer that arms tasks in order (it waits for
    a free ring stage, and for step-(n+1) tasks for the `done` counters of the needed iteration AND the nstages - 1
    iterations before it) and NG consumer groups that take the armed tasks alternately and may finish out of order."""
    W = len(u)
    nit = -(-W // B)
    U0, U1, V = u.copy(), np.zeros_like(u), v.copy()
    done = [0] * (nit + 1)
    active = [min(B, W - it * B) for it in range(nit)]
    lock = threading.Lock()
    deadline = time.time() + timeout
    failed = []

    def tasks_of(b):
        out = []
        for it in range(nit + lag):
            wa = it * B + b
            if it < nit and wa < W:
                out.append((0, it, wa))
            wb = (it - lag) * B + b
            if it >= lag and wb < W:
                out.append((1, it - lag, wb))
        return out

    def run_block(b):
        tasks = tasks_of(b)
        armed = [threading.Event() for _ in tasks]
        released = [threading.Event() for _ in tasks]

        def producer():
            for pit, (kind, it, w) in enumerate(tasks):
                if pit >= nstages:                                  # the ring: stage reuse needs the older task's release
                    while not released[pit - nstages].wait(0.01):
                        if time.time() > deadline:
                            failed.append(("ring", b, pit))
                            return
                if kind == 1:
                    need = min(nit - 1, min(W - 1, w + P) // B)
                    for q in range(need, max(-1, need - nstages), -1):
                        while True:
                            with lock:
                                ok = done[q] == active[q]
                            if ok:
                                break
                            if time.time() > deadline:
                                failed.append(("deadlock", b, pit, q))
                                return
                            time.sleep(1e-5)
                armed[pit].set()

        def consumer(grp):
            rng = random.Random(seed * 7919 + b * 31 + grp)
            for pit, (kind, it, w) in enumerate(tasks):
                if pit % NG != grp:
                    continue
                while not armed[pit].wait(0.01):
                    if time.time() > deadline or failed:
                        return
                time.sleep(rng.random() * 3e-4 * (1 + 3 * grp))      # one group is systematically slower
                if kind == 0:
                    U1[w], V[w] = step_rows(U0, V, w, P, coef, dt)
                else:
                    U0[w], V[w] = step_rows(U1, V, w, P, coef, dt)
                released[pit].set()
                if kind == 0:
                    with lock:
                        done[it] += 1
        ths = [threading.Thread(target=producer)] + [threading.Thread(target=consumer, args=(g,)) for g in range(NG)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    blocks = [threading.Thread(target=run_block, args=(b,)) for b in range(B)]
    for t in blocks:
        t.start()
    for t in blocks:
        t.join()
    return U0, V, failed


@pytest.mark.parametrize("W,B,P", [(48, 4, 5), (45, 5, 2)])
def test_two_step_wavefront_with_consumer_groups(W, B, P):
    rng = np.random.default_rng(W * 3 + B + P)
    u, v = rng.standard_normal(W), rng.standard_normal(W)
    coef = rng.standard_normal(2 * P + 1)
    want_u, want_v = sequential(u, v, P, coef, 0.01)
    lag = -(-P // B) + 1
    for seed in range(2):
        got_u, got_v, failed = wavefront_groups(u, v, P, coef, 0.01, B, lag, seed)
        assert not failed, failed
        assert np.array_equal(got_u, want_u) and np.array_equal(got_v, want_v)

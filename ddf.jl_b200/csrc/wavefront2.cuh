// wavefront2.cuh -- two fused leapfrog steps per launch (included by wave.cu; selected by ddf_wave_leapfrog, see there).
// The schedule is model-checked on host threads (tests/test_wavefront_model.py).
//
// Two fused leapfrog steps per launch with a wavefront through L2.  The persistent blocks of k_rows_pipe walk the
// windows in order; here iteration `it` of a block runs the step-n task of window it*B + b and then the step-(n+1)
// task of the window `lag` iterations back, so the entries of a window are read from HBM once (evict_last) and from
// L2 the second time (evict_first).  u ping-pong: step n reads U0 = a.u and writes U1 = a.out1, step n+1 reads U1 and
// writes U0; v = a.out0 is updated in place by both.  A step-(n+1) task of window w may be staged once every
// step-n window <= w + reach has finished: blocks count their finished step-n tasks per iteration in `done[it]`
// (threadfence + one atomic per task); a WATCHER warp per block follows the counters in order and publishes "every
// iteration <= k is finished" in shared memory, so the producer warp -- whose time per task is the budget of the whole
// pipeline -- never waits for an L2 round trip: it compares a shared-memory word before it arms a step-(n+1) stage.
// With lag >= ceil(reach / B) + 1 a block only ever waits for iterations it has armed itself (no deadlock while all
// blocks are co-resident: one block per SM, grid <= SM count); `slack` more iterations of lag let the blocks drift
// apart by that much before anyone waits (the price: lag * B windows of entries must stay in L2).
#pragma once

#define W2_RING 16             // > the most step-n tasks of one block that can be in flight (ring stages)

struct Wave2Task {
  int kind;        // 0 = step n, 1 = step n+1
  int64_t it;      // the task's own iteration (window = w0 + b + it * stride)
  int64_t win;
};

// The order in which ONE block runs its tasks; producer and consumers enumerate it identically.
struct Wave2Enum {
  int64_t it = 0;
  int phase = 0;
  int64_t nit, lag, w0, w1, stride, b;
  __device__ Wave2Enum(int64_t nit_, int64_t lag_, int64_t w0_, int64_t w1_, int64_t stride_, int64_t b_)
      : nit(nit_), lag(lag_), w0(w0_), w1(w1_), stride(stride_), b(b_) {}
  __device__ bool next(Wave2Task& t) {
    for (;;) {
      if (it >= nit + lag) return false;
      if (phase == 0) {
        phase = 1;
        const int64_t win = w0 + b + it * stride;
        if (it < nit && win < w1) { t.kind = 0; t.it = it; t.win = win; return true; }
      } else {
        phase = 0;
        const int64_t itb = it - lag;
        it++;
        if (itb >= 0) {
          const int64_t win = w0 + b + itb * stride;
          if (win < w1) { t.kind = 1; t.it = itb; t.win = win; return true; }
        }
      }
    }
  }
};

// MODE 2: leapfrog (a.out0 = v in/out, a.dt);  MODE 3: Poisson relaxation sweep (a.rho, a.dinv, a.theta; no v)
template <int MODE, int NG, int NP>
__global__ void __launch_bounds__(NG * PIPE_GROUP + 32 * (NP + 3), 1)
k_rows_pipe2(RowArgs a, const uint32_t* __restrict__ tab, const unsigned char* __restrict__ packed,
             const unsigned char* __restrict__ meta, int64_t w0, int64_t w1, PipeGeom g, int lag, int reachw,
             uint32_t* __restrict__ done, int pf) {
  extern __shared__ __align__(128) unsigned char pipe_smem[];
  __shared__ __align__(8) uint64_t full[PIPE_MAX_STAGES], empty[PIPE_MAX_STAGES], tfull[PIPE_TD];
  __shared__ __align__(16) uint32_t tabS[PIPE_TD * STAGE_TAB];
  __shared__ uint32_t issued[PIPE_MAX_STAGES];
  __shared__ int known_done;                         // every step-n iteration <= known_done is finished on ALL blocks
  // stored[it % W2_RING] = it + 1 once THIS block's step-n task of iteration `it` has issued all its stores (set by
  // its consumer group behind a group barrier); the publisher warp turns that into the device-wide count
  __shared__ int stored[W2_RING];
  __shared__ int prod_it;                            // the latest step-n iteration a producer has armed (prefetch throttle)
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid < PIPE_MAX_STAGES) issued[tid] = 0u;
  if (tid < W2_RING) stored[tid] = 0;
  if (tid == 0) {
    known_done = -1;
    prod_it = 0;
    for (int d = 0; d < PIPE_TD; d++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tfull[d])) : "memory");
    for (int s = 0; s < g.nstages; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(PIPE_GROUP / 32) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t stride = gridDim.x;
  const int64_t nit = (w1 - w0 + stride - 1) / stride;
  if (warp == NG * PIPE_GROUP / 32 + NP) {
    // ------------------------------------------------------------ watcher: iteration q is finished when all its
    // windows (one per block, fewer in the last iteration) have counted themselves in done[q]
    if (lane == 0) {
      for (int64_t q = 0; q < nit; q++) {
        const int64_t left = (w1 - w0) - q * stride;
        const uint32_t active = (uint32_t)(left < stride ? left : stride);
        unsigned long long t0 = 0;
        for (;;) {
          uint32_t seen;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done + q) : "memory");
          if (seen >= active) break;
          // never hang the device: the blocks are co-resident by construction, so a wait of seconds means a broken
          // schedule -- abort the launch (the host sees a launch failure) instead of spinning for ever
          unsigned long long t1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t0 == 0) t0 = t1;
          else if (t1 - t0 > 4000000000ull) __trap();
          __nanosleep(200);
        }
        asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(&known_done)), "r"((int)q) : "memory");
      }
    }
    return;
  }
  if (warp == NG * PIPE_GROUP / 32 + NP + 2) {
    // ------------------------------------------------------------ L2 prefetcher: the first read of a window's entries
    // comes from HBM and takes ~2 us from arming to data; with three stages in shared memory that latency bounds the
    // pipeline.  This warp runs `pf` iterations ahead of the producers and asks L2 for the entries + meta records
    // of the step-n windows (its own table loads may take their time: nobody waits for this warp).
    if (lane == 0 && pf > 0) {
      const uint64_t pol_keep = make_policy_evict_last();
      for (int64_t q = 0; q < nit; q++) {
        const int64_t win = w0 + blockIdx.x + q * stride;
        if (win >= w1) break;
        for (;;) {
          int seen;
          asm volatile("ld.relaxed.cta.shared.s32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&prod_it)) : "memory");
          if ((int64_t)seen + pf >= q) break;
          __nanosleep(200);
        }
        const uint32_t e0 = __ldg(tab + win * STAGE_TAB + 2), nent = __ldg(tab + win * STAGE_TAB + 3);
        if (nent)
          asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(packed + (size_t)e0 * 10),
                       "r"(nent * 10u), "l"(pol_keep) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(meta + win * g.K * STAGE_META),
                     "r"((uint32_t)g.K * STAGE_META), "l"(pol_keep) : "memory");
      }
    }
    return;
  }
  if (warp == NG * PIPE_GROUP / 32 + NP + 1) {
    // ------------------------------------------------------------ publisher: the device-scope fence that makes a
    // finished task's stores visible costs a round trip to L2; it is paid here, off the consumers' path.  The
    // consumers' stores happen-before their group barrier, the barrier before the shared-memory release below, and
    // this thread's fence is cumulative over what it has observed (PTX memory model, causality order).
    if (lane == 0) {
      for (int64_t q = 0; q < nit; q++) {
        if (w0 + blockIdx.x + q * stride >= w1) break;
        for (;;) {
          int seen;
          asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&stored[q % W2_RING])) : "memory");
          if (seen >= (int)q + 1) break;          // slots only grow: a later task in the slot implies this one
          __nanosleep(32);
        }
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(done + q) : "memory");
      }
    }
    return;
  }
  double* const U0 = const_cast<double*>(a.u);       // step n input, step n+1 output
  double* const U1 = a.out1;                         // step n output, step n+1 input
  double* const V = a.out0;                          // MODE 2 only
  const double* const AUX = MODE == 3 ? a.rho : a.out0;   // what travels in the stage's "v" slot
  if (warp >= NG * PIPE_GROUP / 32 && warp < NG * PIPE_GROUP / 32 + NP) {
    // ------------------------------------------------------------ NP producer warps: warp p arms the tasks p mod NP
    // (one warp spends ~1.3 us per task on its ~7 bulk requests and the bookkeeping around them -- more than the
    // consumers need for a window; the stages, the table ring slots (NP divides PIPE_TD) and the dependency check of a
    // task belong to exactly one of the two, so they never touch the same barrier phase)
    const uint32_t ppar = warp - NG * PIPE_GROUP / 32;
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    Wave2Enum ahead(nit, lag, w0, w1, stride, blockIdx.x), cur(nit, lag, w0, w1, stride, blockIdx.x);
    Wave2Task t;
    int64_t aidx = 0;                                  // tasks taken from `ahead` so far
    if (lane == 0) {
      for (int d = 0; d < PIPE_TD; d++) {
        Wave2Task ta;
        const bool ok = ahead.next(ta);
        aidx++;
        if (ok && (uint32_t)(d % NP) == ppar) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tfull[d])), "r"(STAGE_TAB * 4) : "memory");
          bulk_g2s(tabS + d * STAGE_TAB, tab + ta.win * STAGE_TAB, STAGE_TAB * 4, &tfull[d], pol_keep);
        }
      }
    }
    for (int64_t pit = 0; cur.next(t); pit++) {
      if ((uint32_t)(pit % NP) != ppar) continue;
      const int td = (int)(pit % PIPE_TD);
      stage_wait(&tfull[td], (uint32_t)(pit / PIPE_TD) & 1u);
      const uint32_t* wt = tabS + td * STAGE_TAB;
      const uint32_t nruns = wt[0], total = wt[1], e0 = wt[2], nent = wt[3];
      const uint32_t start = lane < STAGE_RMAX ? wt[4 + 2 * lane] : 0u, off = lane < STAGE_RMAX ? wt[5 + 2 * lane] : 0u;
      const uint32_t end = lane + 1 < nruns ? wt[5 + 2 * (lane + 1)] : total;
      __syncwarp();
      if (lane == 0) {                                                      // refill this table slot with task pit + PIPE_TD
        Wave2Task ta;
        bool ok = true;
        while (ok && aidx <= pit + PIPE_TD) { ok = ahead.next(ta); aidx++; }
        if (ok) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tfull[td])), "r"(STAGE_TAB * 4) : "memory");
          bulk_g2s(tabS + td * STAGE_TAB, tab + ta.win * STAGE_TAB, STAGE_TAB * 4, &tfull[td], pol_keep);
        }
      }
      PIPE_TRACE(7, pit);
      if (t.kind == 1) {
        // every step-n window <= win + reachw must be finished (on any block): the watcher says up to which iteration
        int64_t need = ((t.win - w0) + reachw) / stride;
        if (need > nit - 1) need = nit - 1;
        for (;;) {
          int seen;
          asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&known_done)) : "memory");
          if ((int64_t)seen >= need) break;
          __nanosleep(32);
        }
        __syncwarp();
        asm volatile("fence.proxy.async.global;" ::: "memory");             // generic-proxy stores of U1 / V -> our bulk reads
      }
      const int s = (int)(pit % g.nstages);
      const uint32_t j = (uint32_t)(pit / g.nstages);
      PIPE_TRACE(5, pit);
      if (j > 0) {
        // The previous use of this stage was armed by ANOTHER producer warp.  A parity wait entered before that use
        // has even been armed would look at the phase before it and pass at once (same aliasing as on the consumer
        // side): first see the previous use armed, then wait for its release.
        for (;;) {
          uint32_t seen;
          asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&issued[s])) : "memory");
          if (seen >= (uint32_t)(pit - g.nstages) + 1u) break;
        }
        stage_wait(&empty[s], (j - 1) & 1u);
      }
      unsigned char* sb = pipe_smem + (size_t)s * g.stage_bytes;
      const int64_t r0 = t.win * g.K * DDF_SELL_WIN;
      const uint32_t nr = (uint32_t)min((int64_t)g.K * DDF_SELL_WIN, g.nrows - r0);
      const uint32_t vbytes = ((nr + 1u) & ~1u) * 8u;
      const uint32_t mbytes = (uint32_t)g.K * STAGE_META;
      const uint32_t tx = total * 8u + nent * 10u + mbytes + 2u * vbytes;
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(tx) : "memory");
        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(&issued[s])), "r"((uint32_t)pit + 1u) : "memory");
        if (t.kind == 0) asm volatile("st.relaxed.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(&prod_it)), "r"((int)t.it) : "memory");
      }
      PIPE_TRACE(0, pit);
      __syncwarp();
      const double* uin = t.kind ? U1 : U0;
      // first touch of a window's entries keeps them in L2 for the second step, the second touch lets them go
      const uint64_t pol_ent = t.kind ? pol_stream : pol_keep;
      if (lane < nruns) bulk_g2s(sb + (size_t)off * 8, uin + start, (end - off) * 8u, &full[s], pol_keep);
      if (lane == 31 && nent) bulk_g2s(sb + g.o_ent, packed + (size_t)e0 * 10, nent * 10u, &full[s], pol_ent);
      if (lane == 30) bulk_g2s(sb + g.o_meta, meta + t.win * g.K * STAGE_META, mbytes, &full[s], pol_ent);
      if (lane == 29) {
        bulk_g2s(sb + g.o_v, AUX + r0, vbytes, &full[s], MODE == 3 ? pol_keep : pol_stream);
        bulk_g2s(sb + g.o_u, uin + r0, vbytes, &full[s], pol_keep);
      }
      PIPE_TRACE(6, pit);
    }
  } else {
    // ------------------------------------------------------------ consumer groups: group k takes tasks k, k + NG, ...
    const uint32_t grp = tid / PIPE_GROUP;
    const uint32_t gtid = tid % PIPE_GROUP, gwarp = gtid >> 5;
    Wave2Enum cur(nit, lag, w0, w1, stride, blockIdx.x);
    Wave2Task t;
    for (int64_t pit = 0; cur.next(t); pit++) {
      if ((uint32_t)(pit % NG) != grp) continue;
      const int s = (int)(pit % g.nstages);
      const uint32_t j = (uint32_t)(pit / g.nstages);
      for (;;) {                          // the producer has armed stage s for THIS task
        uint32_t seen;
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&issued[s])) : "memory");
        if (seen >= (uint32_t)pit + 1u) break;
      }
      PIPE_TRACE(3, pit);
      stage_wait(&full[s], j & 1u);
      PIPE_TRACE(1, pit);
      const unsigned char* sb = pipe_smem + (size_t)s * g.stage_bytes;
      const double* stage = reinterpret_cast<const double*>(sb);
      double* const uout = t.kind ? U0 : U1;
      uint32_t rec = 0;
      for (int kk = 0; kk < g.K; kk++) {
        const unsigned char* metaS = sb + g.o_meta + kk * STAGE_META;
        const uint32_t nentS = *reinterpret_cast<const uint32_t*>(metaS);
        const double* valS = reinterpret_cast<const double*>(sb + g.o_ent + rec);
        const uint16_t* slotS = reinterpret_cast<const uint16_t*>(sb + g.o_ent + rec + (size_t)nentS * 8);
        rec += nentS * 10u;
        const uint8_t* lenS = metaS + STAGE_META_HDR + DDF_SELL_WIN;
        const uint32_t mylen = lenS[gtid];
        const uint32_t p = (metaS + STAGE_META_HDR)[gtid];
        const uint32_t part = reinterpret_cast<const uint32_t*>(metaS)[4 + gwarp];
        const uint32_t maxlen = __shfl_sync(0xffffffffu, mylen, 0);
        const uint32_t pl = (uint32_t)kk * DDF_SELL_WIN + p;
        const int64_t i = t.win * g.K * DDF_SELL_WIN + pl;
        // deep halo (a.plane > 0, MODE 2): the step-(n+1) tasks cover the owned rows only
        const int64_t r0 = (MODE == 2 && t.kind && a.plane > 0) ? a.row0b : a.row0;
        const int64_t r1 = (MODE == 2 && t.kind && a.plane > 0) ? a.row1b : a.row1;
        double dinv_i = 0.0;                                 // MODE 3: requested now, needed after the row sum
        if (MODE == 3 && i >= r0 && i < r1) dinv_i = __ldg(a.dinv + i);
        const double acc = pipe_row_sum<DDF_PIPE_UNR>(valS, slotS, stage, part, lane, mylen, maxlen);
        if (i >= r0 && i < r1) {
          const double m = (metaS + STAGE_META_HDR + 2 * DDF_SELL_WIN)[p] ? 0.0 : 1.0;
          const double ev = reinterpret_cast<const double*>(sb + g.o_v)[pl];
          if (MODE == 3) {
            uout[i] = jacobi_update(reinterpret_cast<const double*>(sb + g.o_u)[pl], acc, ev, dinv_i, m, a.theta);
          } else {
            const double vn = __dadd_rn(ev, __dmul_rn(a.dt, __dmul_rn(m, acc)));
            V[i] = vn;
            const double un = __dadd_rn(reinterpret_cast<const double*>(sb + g.o_u)[pl], __dmul_rn(a.dt, __dmul_rn(m, vn)));
            uout[i] = un;
            if (MODE == 2 && t.kind && a.plane > 0) {
              // the state after BOTH steps of my outermost owned planes goes straight into the neighbours' landing arenas
              const int64_t P = a.plane;
              if (a.peer_lo) {
                if (i < r0 + P) { a.peer_lo[i - r0] = un; a.peer_lo3[i - r0] = vn; }
                else if (i < r0 + 2 * P) a.peer_lo2[i - r0 - P] = un;
              }
              if (a.peer_hi) {
                if (i >= r1 - P) { a.peer_hi[i - (r1 - P)] = un; a.peer_hi3[i - (r1 - P)] = vn; }
                else if (i >= r1 - 2 * P) a.peer_hi2[i - (r1 - 2 * P)] = un;
              }
            }
          }
        }
      }
      __syncwarp();
      PIPE_TRACE(2, pit);
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      if (t.kind == 0) {
        // hand the task to the publisher: all stores of the group issued (group barrier), then one shared-memory flag
        asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)grp), "r"(PIPE_GROUP) : "memory");
        if (gtid == 0)
          asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(&stored[t.it % W2_RING])), "r"((int)t.it + 1) : "memory");
      }
    }
  }
}

// largest |column - row| of the assembled L rows (how far a row reaches), for the wavefront's dependency distance
__global__ void k_row_reach(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr, int64_t nv,
                            unsigned long long* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long r = 0;
  if (i < nv)
    for (uint32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
      const long long d = (long long)nbr[p] - (long long)i;
      const unsigned long long ad = (unsigned long long)(d < 0 ? -d : d);
      r = ad > r ? ad : r;
    }
  r = __reduce_max_sync(0xffffffffu, (unsigned)r);      // local dimensions stay below 2^31
  if ((threadIdx.x & 31) == 0 && r > *reinterpret_cast<volatile unsigned long long*>(out)) atomicMax(out, r);
}

// Two leapfrog steps in one launch: on return u (U0) holds the state after both steps, utmp the one in between.
// kind 0: leapfrog (aux = v in/out, coef = dt);  kind 1: Poisson relaxation sweep (aux = rho read-only, coef = theta)
// deep halo of a slab (comm.cu): row ranges of the two steps and where the outermost owned planes go
struct Pipe2Deep {
  int64_t row0, row1;      // step n: owned rows + one ghost plane on either side that has one
  int64_t row0b, row1b;    // step n + 1: owned rows
  int64_t plane;
  double* peers[6];        // lo: u nearest plane, u second plane, v nearest plane; hi: the same
};

// geometry and schedule of the two-step launch on this mesh; *ok = false when the wavefront cannot pay
struct Pipe2Plan {
  PipeGeom g;
  int64_t g0, g1, nit;
  int grid, lag, reachw;
};
static int pipe2_plan(ddf_mesh* m, cudaStream_t stream, Pipe2Plan* pl, bool* ok) {
  *ok = false;
  const StageBufs& sb = m->lstage;
  if (!sb.ok) return 0;
  if (!pipe_geometry(m, &pl->g)) return 0;
  const int64_t nv = m->n[0];
  const int64_t gw = (int64_t)sb.K * DDF_SELL_WIN;
  pl->g0 = 0;
  pl->g1 = (nv + gw - 1) / gw;
  pl->grid = (int)std::min<int64_t>(m->ctx->sm_count, pl->g1 - pl->g0);
  if (m->wave2_reach < 0) {
    DevBuf<unsigned long long> r;
    DDF_CHECK(r.alloc(m->ctx, 1));
    DDF_CUDA(cudaMemsetAsync(r.p, 0, 8, stream));
    int gr;
    DDF_CHECK(grid_for(nv, 256, &gr));
    k_row_reach<<<gr, 256, 0, stream>>>(m->adj_ptr.p, m->adj_nbr.p, nv, r.p);
    DDF_LAUNCH_CHECK(m->ctx);
    unsigned long long h = 0;
    DDF_CUDA(cudaMemcpyAsync(&h, r.p, 8, cudaMemcpyDeviceToHost, stream));
    DDF_CUDA(cudaStreamSynchronize(stream));
    m->wave2_reach = (int64_t)h;
  }
  pl->reachw = (int)(m->wave2_reach / gw) + 2;
  // tuning bits 18-21 of "wave": extra iterations of lag (default 3: measured best at C4 together with prefetch distance 2, profiles/r02_tuning.md)
  const int slack_t = (m->ctx->tune.wave >> 18) & 15;
  const int slack = slack_t ? slack_t - 1 : 3;
  pl->lag = (pl->reachw + pl->grid - 1) / pl->grid + 1 + slack;
  pl->nit = (pl->g1 - pl->g0 + pl->grid - 1) / pl->grid;
  // worth it only when the wavefront is short against the sweep: the windows between the two steps must stay in L2
  // (lag * grid windows of entries) and the first / last `lag` iterations run one step only
  if (pl->nit < 4 * (int64_t)pl->lag) return 0;
  const int64_t held = (int64_t)pl->lag * pl->grid * (int64_t)sb.max_entries * 10;
  if (held > (int64_t)64 << 20) return 0;
  *ok = true;
  return 0;
}

static int launch_pipe2(ddf_mesh* m, cudaStream_t stream, int kind, double* u, double* aux, double* utmp, double coef, bool* ran,
                        const Pipe2Deep* deep = nullptr) {
  *ran = false;
  const StageBufs& sb = m->lstage;
  if (!sb.ok || (m->ctx->nranks > 1 && !deep)) return 0;
  Pipe2Plan pl;
  bool ok = false;
  DDF_CHECK(pipe2_plan(m, stream, &pl, &ok));
  if (!ok) return 0;
  PipeGeom g = pl.g;
  const int64_t nv = m->n[0], g0 = pl.g0, g1 = pl.g1, nit = pl.nit;
  const int grid = pl.grid, lag = pl.lag, reachw = pl.reachw;
  if ((int64_t)m->wave2_done.n < nit) DDF_CHECK(m->wave2_done.alloc(m->ctx, (size_t)nit));
  DDF_CUDA(cudaMemsetAsync(m->wave2_done.p, 0, (size_t)nit * 4, stream));
  RowArgs a{};
  a.u = u;
  a.out1 = utmp;
  a.row0 = 0;
  a.row1 = nv;
  a.isb = m->isbnd[0].p;
  if (deep) {
    a.row0 = deep->row0; a.row1 = deep->row1; a.row0b = deep->row0b; a.row1b = deep->row1b; a.plane = deep->plane;
    a.peer_lo = deep->peers[0]; a.peer_lo2 = deep->peers[1]; a.peer_lo3 = deep->peers[2];
    a.peer_hi = deep->peers[3]; a.peer_hi2 = deep->peers[4]; a.peer_hi3 = deep->peers[5];
  }
  if (kind == 0) { a.out0 = aux; a.dt = coef; }
  else { a.rho = aux; a.dinv = m->ldinv.p; a.theta = coef; }
  const size_t smem_p = (size_t)g.nstages * g.stage_bytes;
  if (getenv("DDF_DEBUG"))
    fprintf(stderr, "[ddf] k_rows_pipe2: %d stages x %u B (slots %u, entries %u), grid %d, reach %lld rows = %d windows, lag %d, nit %lld\n",
            g.nstages, g.stage_bytes, sb.max_slots, sb.max_entries, grid, (long long)m->wave2_reach, reachw, lag, (long long)nit);
  const int np = (m->ctx->tune.wave >> 22) & 7;      // tuning bits 22-24: producer warps (0 = default)
  // COOPERATIVE launch: the blocks wait for each other's step-n windows, so all of them must be resident at once.
  // One block per SM and grid <= SM count make that true on an idle device; the cooperative launch makes the driver
  // guarantee it (or refuse the launch) when other work shares the GPU.  A refused launch falls back to one step per
  // launch (*ran stays false).
  const uint32_t* tab_p = sb.tab.p;
  const unsigned char *packed_p = sb.packed.p, *meta_p = sb.meta.p;
  int64_t g0v = g0, g1v = g1;
  int lagv = lag, reachv = reachw;
  uint32_t* done_p = m->wave2_done.p;
  const int pf_t = (m->ctx->tune.wave >> 25) & 15;    // tuning bits 25-28: L2 prefetch distance + 1 (0 = default)
  int pfv = pf_t ? pf_t - 1 : 2;
  void* kargs[] = {&a, &tab_p, &packed_p, &meta_p, &g0v, &g1v, &g, &lagv, &reachv, &done_p, &pfv};
#define DDF_PIPE2_LAUNCH(MODEV, NPV)                                                                                   \
  do {                                                                                                                 \
    DDF_CUDA(cudaFuncSetAttribute(k_rows_pipe2<MODEV, 2, NPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p)); \
    cudaError_t le = cudaLaunchCooperativeKernel((const void*)k_rows_pipe2<MODEV, 2, NPV>, dim3(grid),                 \
                                                 dim3(2 * PIPE_GROUP + 32 * (NPV + 3)), kargs, smem_p, stream);         \
    if (le != cudaSuccess) { cudaGetLastError(); return 0; }                                                           \
  } while (0)
  if (kind == 1) DDF_PIPE2_LAUNCH(3, 2);
  else if (np == 1) DDF_PIPE2_LAUNCH(2, 1);
  else if (np == 4) DDF_PIPE2_LAUNCH(2, 4);
  else DDF_PIPE2_LAUNCH(2, 2);
#undef DDF_PIPE2_LAUNCH
  DDF_LAUNCH_CHECK(m->ctx);
  *ran = true;
  return 0;
}

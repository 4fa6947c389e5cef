// fstage.cuh -- staged FIXED-WIDTH rows: y = d_k x through a persistent TMA pipeline.
//
// d_k (the adjoint of the boundary matrix, ManifoldOps.jl:46) has exactly W = k+2 entries per
// row, +-1 with a constant sign pattern; the explicit-index kernel k_apply_fixed<W> spends
// 4 B per entry on column indices and its memory-pipe slots on 8-byte gathers of x.  Because
// face ids are lexicographic ranks of vertex tuples (Manifolds.jl:735-757) and every
// generator numbers vertices with locality, the columns touched by 1024 consecutive rows
// form a few contiguous runs.  This format stores per 1024-row window
//   * the run table {nruns, nslots, -, -, (start_col, slot_off) x <= 30} (256 B, as stage.cuh);
//     runs start on even columns, have even length, and absorb gaps of <= FSTAGE_GAP columns;
//   * per entry a 16-bit slot into the staged copy of x: 2 B instead of 4 B.
// k_fixed_pipe<W>: one block per SM; a producer warp bulk-copies the runs of x and the
// window's slot stream into a shared-memory ring (cp.async.bulk + mbarrier), consumer groups
// gather from shared memory, sum with the implicit signs in the reference's order
// (tmp = 0; tmp += v_j x[c_j], j ascending) and store y coalesced.
// A table whose windows need more than 30 runs or too many slots is rejected (state 2) and the
// explicit-index kernel is used.
#pragma once
#include "stage.cuh"

#define FSTAGE_WIN_MAX 1024          // rows per window: 256 * RPT, RPT = rows per consumer thread (A/B knob)
#define FSTAGE_GAP 15u
#define FSTAGE_MAX_SLOTS 6144u           // 48 KB of staged Float64 per window

static __global__ void __launch_bounds__(1024)
k_fstage_build(const int32_t* __restrict__ tab, int W, int64_t nrows, int win, uint32_t* __restrict__ wtab_all,
               uint16_t* __restrict__ slot, uint32_t* __restrict__ status) {
  __shared__ uint32_t keys[4096];
  __shared__ uint32_t rstart[STAGE_RMAX + 1], roff[STAGE_RMAX + 1];
  __shared__ uint32_t s_nruns, s_total, s_bad;
  const int t = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * win;
  const int64_t nr = min((int64_t)win, nrows - r0);
  const int nent = (int)nr * W;
  for (int e = t; e < 4096; e += 1024) keys[e] = e < nent ? (uint32_t)tab[r0 * W + e] : 0xFFFFFFFFu;
  __syncthreads();
  // bitonic sort of 4096 keys
  for (int k = 2; k <= 4096; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < 4096; i += 1024) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (t == 0) {
    uint32_t nruns = 0, total = 0, bad = 0;
    uint32_t cur_start = 0, cur_end = 0;
    bool open = false;
    for (int e = 0; e < nent; e++) {
      const uint32_t c = keys[e];
      if (open && c <= cur_end + FSTAGE_GAP) {
        const uint32_t ne = (c + 2u) & ~1u;
        if (ne > cur_end) cur_end = ne;
      } else {
        if (open) {
          if (nruns < STAGE_RMAX) { rstart[nruns] = cur_start; roff[nruns] = total; }
          total += cur_end - cur_start;
          nruns++;
        }
        cur_start = c & ~1u;
        cur_end = (c + 2u) & ~1u;
        open = true;
      }
    }
    if (open) {
      if (nruns < STAGE_RMAX) { rstart[nruns] = cur_start; roff[nruns] = total; }
      total += cur_end - cur_start;
      nruns++;
    }
    if (nruns > STAGE_RMAX || total > FSTAGE_MAX_SLOTS) bad = 1;
    s_nruns = nruns; s_total = total; s_bad = bad;
    uint32_t* wt = wtab_all + (int64_t)blockIdx.x * STAGE_TAB;
    wt[0] = bad ? 0u : nruns;
    wt[1] = bad ? 0u : total;
    wt[2] = 0u;
    wt[3] = (uint32_t)nent;
    if (!bad)
      for (uint32_t k = 0; k < nruns; k++) { wt[4 + 2 * k] = rstart[k]; wt[5 + 2 * k] = roff[k]; }
    atomicMax(status, bad ? 0xFFFFFFFFu : total);
  }
  __syncthreads();
  if (s_bad) return;
  const uint32_t nruns = s_nruns;
  for (int e = t; e < win * W; e += 1024) {
    uint16_t sl = 0;
    if (e < nent) {
      const uint32_t c = (uint32_t)tab[r0 * W + e];
      uint32_t lo = 0, hi = nruns;                 // last run with start <= c
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (rstart[mid] <= c) lo = mid; else hi = mid;
      }
      sl = (uint16_t)(roff[lo] + (c - rstart[lo]));
    }
    slot[r0 * W + e] = sl;
  }
}

static int fstage_build(ddf_ctx* ctx, const int32_t* tab, int W, int64_t nrows, int win, FStageBufs& out) {
  out.release();
  out.state = 2;
  if (W < 2 || W > 4 || nrows < 4096) return 0;     // tiny tables: the explicit-index kernel is fine
  const int64_t nwin = (nrows + win - 1) / win;
  DevBuf<uint32_t> status;
  DDF_CHECK(status.alloc(ctx, 1));
  DDF_CUDA(cudaMemsetAsync(status.p, 0, 4, ctx->stream));
  DDF_CHECK(out.tab.alloc(ctx, nwin * STAGE_TAB));
  DDF_CHECK(out.slot.alloc(ctx, nwin * win * W + 8));
  k_fstage_build<<<(int)nwin, 1024, 0, ctx->stream>>>(tab, W, nrows, win, out.tab.p, out.slot.p, status.p);
  DDF_LAUNCH_CHECK(ctx);
  uint32_t st = 0;
  DDF_CUDA(cudaMemcpyAsync(&st, status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (st == 0xFFFFFFFFu) { out.release(); out.state = 2; return 0; }
  out.max_slots = st;
  out.W = W;
  out.win = win;
  out.nwin = nwin;
  out.state = 1;
  return 0;
}

struct FPipeGeom {
  uint32_t o_slot, stage_bytes;
  int nstages;
  int64_t nrows;
};

template <int W, int NG, int RPT>
__global__ void __launch_bounds__(NG * 256 + 32, 1)
k_fixed_pipe(const uint32_t* __restrict__ tab, const uint16_t* __restrict__ slot, const double* __restrict__ x,
             double* __restrict__ y, int first_negative, int64_t nwin, FPipeGeom g) {
  extern __shared__ __align__(128) unsigned char pipe_smem[];
  __shared__ __align__(8) uint64_t full[8], empty[8], tfull[PIPE_TD];
  __shared__ __align__(16) uint32_t tabS[PIPE_TD * STAGE_TAB];
  __shared__ uint32_t issued[8];       // see k_rows_pipe: guards the parity wait of consumer groups that share the ring
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid < 8) issued[tid] = 0u;
  if (tid == 0) {
    for (int d = 0; d < PIPE_TD; d++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tfull[d])) : "memory");
    for (int s = 0; s < g.nstages; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&empty[s])) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t stride = gridDim.x;
  constexpr int WIN = 256 * RPT;
  constexpr uint32_t SLOT_BYTES = WIN * W * 2;
  if (warp == NG * 8) {
    // ------------------------------------------------------------ producer warp
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    int64_t win = blockIdx.x;
    if (lane == 0) {
      for (int d = 0; d < PIPE_TD; d++) {
        const int64_t w = win + d * stride;
        if (w < nwin) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tfull[d])), "r"(STAGE_TAB * 4) : "memory");
          bulk_g2s(tabS + d * STAGE_TAB, tab + w * STAGE_TAB, STAGE_TAB * 4, &tfull[d], pol_stream);
        }
      }
    }
    for (int64_t it = 0; win < nwin; it++, win += stride) {
      const int td = (int)(it % PIPE_TD);
      stage_wait(&tfull[td], (uint32_t)(it / PIPE_TD) & 1u);
      const uint32_t* wt = tabS + td * STAGE_TAB;
      const uint32_t nruns = wt[0], total = wt[1];
      const uint32_t start = lane < STAGE_RMAX ? wt[4 + 2 * lane] : 0u, off = lane < STAGE_RMAX ? wt[5 + 2 * lane] : 0u;
      const uint32_t end = lane + 1 < nruns ? wt[5 + 2 * (lane + 1)] : total;
      __syncwarp();
      if (lane == 0 && win + PIPE_TD * stride < nwin) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tfull[td])), "r"(STAGE_TAB * 4) : "memory");
        bulk_g2s(tabS + td * STAGE_TAB, tab + (win + PIPE_TD * stride) * STAGE_TAB, STAGE_TAB * 4, &tfull[td], pol_stream);
      }
      const int s = (int)(it % g.nstages);
      const uint32_t j = (uint32_t)(it / g.nstages);
      if (j > 0) stage_wait(&empty[s], (j - 1) & 1u);
      unsigned char* sb = pipe_smem + (size_t)s * g.stage_bytes;
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(total * 8u + SLOT_BYTES) : "memory");
        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(&issued[s])), "r"((uint32_t)it + 1u) : "memory");
      }
      __syncwarp();
      if (lane < nruns) bulk_g2s(sb + (size_t)off * 8, x + start, (end - off) * 8u, &full[s], pol_keep);
      if (lane == 31) bulk_g2s(sb + g.o_slot, slot + win * (WIN * W), SLOT_BYTES, &full[s], pol_stream);
    }
  } else {
    // ------------------------------------------------------------ consumer groups: RPT rows per thread
    const uint64_t pol_stream = make_policy_evict_first();
    const uint32_t grp = tid >> 8, gtid = tid & 255u;
    int64_t it = grp;
    for (int64_t win = blockIdx.x + grp * stride; win < nwin; it += NG, win += NG * stride) {
      const int s = (int)(it % g.nstages);
      const uint32_t j = (uint32_t)(it / g.nstages);
      for (;;) {                          // the producer has armed stage s for THIS iteration
        uint32_t seen;
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&issued[s])) : "memory");
        if (seen >= (uint32_t)it + 1u) break;
      }
      stage_wait(&full[s], j & 1u);
      const unsigned char* sb = pipe_smem + (size_t)s * g.stage_bytes;
      const double* stage = reinterpret_cast<const double*>(sb);
      const uint16_t* slotS = reinterpret_cast<const uint16_t*>(sb + g.o_slot);
      const int64_t r0 = win * WIN;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        const uint32_t r = gtid + 256u * q;
        uint32_t sl[W];
        if (W == 4) {
          const uint2 v = *reinterpret_cast<const uint2*>(slotS + 4 * r);
          sl[0] = v.x & 0xffffu; sl[1] = v.x >> 16; sl[2] = v.y & 0xffffu; sl[W - 1] = v.y >> 16;
        } else if (W == 2) {
          const uint32_t v = *reinterpret_cast<const uint32_t*>(slotS + 2 * r);
          sl[0] = v & 0xffffu; sl[1] = v >> 16;
        } else {
#pragma unroll
          for (int p = 0; p < W; p++) sl[p] = slotS[W * r + p];
        }
        double a = 0.0;                                   // tmp = zero(eltype(C))
#pragma unroll
        for (int p = 0; p < W; p++) {
          const double v = stage[sl[p]];
          const bool neg = ((p & 1) != 0) != (first_negative != 0);
          a = neg ? __dsub_rn(a, v) : __dadd_rn(a, v);
        }
        if (r0 + r < g.nrows) st_hint_f64(y + r0 + r, a, pol_stream);
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
  }
}

template <int W, int NG, int RPT>
static int launch_fixed_pipe_v(ddf_ctx* ctx, const FStageBufs& fs, int64_t nrows, int first_negative, const double* x,
                               double* y) {
  FPipeGeom g;
  g.o_slot = (fs.max_slots * 8u + 15u) & ~15u;
  g.stage_bytes = (g.o_slot + 256 * RPT * W * 2 + 127u) & ~127u;
  g.nstages = (int)((227u * 1024u - 4096u) / g.stage_bytes);
  if (g.nstages > 8) g.nstages = 8;
  if (g.nstages < 2) return -1;
  g.nrows = nrows;
  const size_t smem = (size_t)g.nstages * g.stage_bytes;
  static size_t attr = 0;
  if (attr < smem) {
    DDF_CUDA(cudaFuncSetAttribute(k_fixed_pipe<W, NG, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  static bool said = false;
  if (!said && getenv("DDF_DEBUG")) {
    fprintf(stderr, "[ddf] k_fixed_pipe<%d,%d,%d>: %d stages x %u B (slots %u)\n", W, NG, RPT, g.nstages, g.stage_bytes,
            fs.max_slots);
    said = true;
  }
  const int grid = (int)std::min<int64_t>(ctx->sm_count, fs.nwin);
  k_fixed_pipe<W, NG, RPT><<<grid, NG * 256 + 32, smem, ctx->stream>>>(fs.tab.p, fs.slot.p, x, y, first_negative, fs.nwin, g);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

// variant = 10*NG + RPT must match the window size the table was built with (fs.win = 256*RPT)
template <int W>
static int launch_fixed_pipe(ddf_ctx* ctx, const FStageBufs& fs, int ng, int64_t nrows, int first_negative,
                             const double* x, double* y) {
  const int rpt = fs.win / 256;
  if (ng == 1 && rpt == 1) return launch_fixed_pipe_v<W, 1, 1>(ctx, fs, nrows, first_negative, x, y);
  if (ng == 1 && rpt == 2) return launch_fixed_pipe_v<W, 1, 2>(ctx, fs, nrows, first_negative, x, y);
  if (ng == 1 && rpt == 4) return launch_fixed_pipe_v<W, 1, 4>(ctx, fs, nrows, first_negative, x, y);
  if (ng == 2 && rpt == 1) return launch_fixed_pipe_v<W, 2, 1>(ctx, fs, nrows, first_negative, x, y);
  if (ng == 2 && rpt == 2) return launch_fixed_pipe_v<W, 2, 2>(ctx, fs, nrows, first_negative, x, y);
  return launch_fixed_pipe_v<W, 2, 4>(ctx, fs, nrows, first_negative, x, y);
}

// ---------------------------------------------------------------------------------------------
// Tiled variant (no TMA): one block per window of 256*RPT rows.  All 256 threads copy the window's
// runs of x into shared memory with coalesced 16-byte loads (every run starts on an even column), their
// slot words are already in flight, then each thread gathers its RPT rows from shared memory.
// Unlike the pipeline above the loads are issued by 256 threads, not by one producer warp, so the
// per-request cost of cp.async.bulk does not apply; latency is hidden by 4-5 resident blocks per SM.
__device__ __forceinline__ double2 ld_keep_f64x2(const double2* p, uint64_t pol) {
  double2 r;
  asm("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
  return r;
}

template <int W, int RPT>
__global__ void __launch_bounds__(256)
k_fixed_tiled(const uint32_t* __restrict__ tab, const uint16_t* __restrict__ slot, const double* __restrict__ x,
              double* __restrict__ y, int first_negative, int64_t nrows) {
  extern __shared__ __align__(16) double tile[];
  __shared__ uint32_t wt[STAGE_TAB];
  constexpr int WIN = 256 * RPT;
  const uint32_t tid = threadIdx.x;
  const int64_t win = blockIdx.x;
  const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
  if (tid < STAGE_TAB) wt[tid] = (uint32_t)ld_stream_int(reinterpret_cast<const int*>(tab) + win * STAGE_TAB + tid);
  // this thread's slot words (rows tid, tid+256, ...), requested before the tile is filled
  uint32_t sl[RPT][W];
  const uint16_t* ws = slot + win * (int64_t)(WIN * W);
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    const uint32_t r = tid + 256u * q;
    if (W == 4) {
      const int2 v = ld_stream_int2(reinterpret_cast<const int2*>(ws + 4 * r));
      sl[q][0] = (uint32_t)v.x & 0xffffu; sl[q][1] = (uint32_t)v.x >> 16;
      sl[q][2] = (uint32_t)v.y & 0xffffu; sl[q][W - 1] = (uint32_t)v.y >> 16;
    } else if (W == 2) {
      const uint32_t v = (uint32_t)ld_stream_int(reinterpret_cast<const int*>(ws + 2 * r));
      sl[q][0] = v & 0xffffu; sl[q][1] = v >> 16;
    } else {
#pragma unroll
      for (int p = 0; p < W; p++) sl[q][p] = __ldg(ws + W * r + p);
    }
  }
  __syncthreads();
  const uint32_t nruns = wt[0], total = wt[1];
  for (uint32_t k = 0; k < nruns; k++) {
    const uint32_t start = wt[4 + 2 * k], off = wt[5 + 2 * k];
    const uint32_t len = (k + 1 < nruns ? wt[5 + 2 * (k + 1)] : total) - off;
    for (uint32_t i = 2 * tid; i < len; i += 512)      // asynchronous 16-byte copies: no register waits on the load
      asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(tile + off + i)),
                   "l"(x + start + i), "l"(pol_keep)
                   : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const int64_t r0 = win * WIN;
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    const uint32_t r = tid + 256u * q;
    double a = 0.0;                                   // tmp = zero(eltype(C))
#pragma unroll
    for (int p = 0; p < W; p++) {
      const double v = tile[sl[q][p]];
      const bool neg = ((p & 1) != 0) != (first_negative != 0);
      a = neg ? __dsub_rn(a, v) : __dadd_rn(a, v);
    }
    if (r0 + r < nrows) st_hint_f64(y + r0 + r, a, pol_stream);
  }
}

template <int W, int RPT>
static int launch_fixed_tiled_v(ddf_ctx* ctx, const FStageBufs& fs, int64_t nrows, int first_negative, const double* x,
                                double* y) {
  const size_t smem = ((size_t)fs.max_slots * 8 + 15) & ~(size_t)15;
  static size_t attr = 0;
  if (smem > 48 * 1024 && attr < smem) {
    DDF_CUDA(cudaFuncSetAttribute(k_fixed_tiled<W, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  if (fs.nwin > 0x7fffffffLL) DDF_FAIL("grid too large");
  k_fixed_tiled<W, RPT><<<(int)fs.nwin, 256, smem, ctx->stream>>>(fs.tab.p, fs.slot.p, x, y, first_negative, nrows);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}
template <int W>
static int launch_fixed_tiled(ddf_ctx* ctx, const FStageBufs& fs, int64_t nrows, int first_negative, const double* x,
                              double* y) {
  const int rpt = fs.win / 256;
  if (rpt == 1) return launch_fixed_tiled_v<W, 1>(ctx, fs, nrows, first_negative, x, y);
  if (rpt == 2) return launch_fixed_tiled_v<W, 2>(ctx, fs, nrows, first_negative, x, y);
  return launch_fixed_tiled_v<W, 4>(ctx, fs, nrows, first_negative, x, y);
}

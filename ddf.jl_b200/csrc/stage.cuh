// stage.cuh -- "staged rows": jagged sliced storage with 16-bit staged column slots.
//
// The row kernels of the vertex Laplacian spend their bytes on 4-byte column indices and
// their memory-pipe slots on 8-byte random gathers of u.  On a mesh whose vertex numbering
// has locality (every generator of the reference: ManifoldConstructors.jl:196-284 numbers
// the Kuhn grid lexicographically), the columns touched by a window of 256 consecutive rows
// form a few contiguous runs (9 on the 3D Kuhn mesh: the x-rows (dy,dz) around the window).
// This format stores, per 256-row window,
//   * a run table  {nruns, nslots, first entry, entries, (start_col_k, slot_off_k)...}  -- runs start on even
//     columns, have even length and absorb gaps of <= STAGE_GAP untouched columns, so each is one
//     16-byte-aligned TMA bulk copy
//     (cp.async.bulk global -> shared, completion on an mbarrier) of u into a staging buffer;
//   * per stored entry a 16-bit slot into that buffer + the 8-byte value: 10 B instead of 12 B, packed per
//     window as one record [values][slots] (one bulk copy), plus an 800-byte meta record (perm, len, mask);
//   * rows sorted by decreasing length inside the window (perm8, len8 in sorted order) and
//     every slice of 32 sorted rows stored as jagged diagonals: "column" q holds the entries
//     of the cnt_q = #{rows with len > q} leading lanes, packed -- no padding at all.
// The gather then reads shared memory; HBM sees every u element once per window that needs
// it (L2 hits after the first), every entry exactly once, fully coalesced.
// The order of the entries inside a row is unchanged: sums are bit-identical to the CSR form.
//
// Windows whose columns do not fit (more than STAGE_RMAX runs, more than STAGE_MAX_SLOTS
// slots, more than STAGE_KEYS entries, a row longer than 255) make the whole build report
// "unstageable"; the caller then keeps the explicit-index SELL kernel.
#pragma once
#include <cub/cub.cuh>

#include "mesh.cuh"

#define STAGE_RMAX 30                    // runs per window
#define STAGE_TAB (4 + 2 * STAGE_RMAX)    // table words per window (256 B: one 16-byte-aligned bulk copy)
#define STAGE_KEYS 8192                  // most entries of a 256-row window the builder sorts
#define STAGE_GAP 7u                     // untouched columns a run may absorb
#define STAGE_META 816                   // bytes of the per-window meta record
#define STAGE_META_HDR 48                // {u32 nent, 3 pad, u32 slice offset x 8} then perm8, len8, mask
#define STAGE_MAX_SLOTS 6080             // 47.5 KB of staged Float64 per block (static shared-memory limit)

// ---- pass A: sort the rows of a window by decreasing length (stable), slice entry totals
static __global__ void __launch_bounds__(DDF_SELL_WIN)
k_stage_sort(const uint32_t* __restrict__ rowptr, int64_t nrows, uint8_t* __restrict__ perm8,
             uint8_t* __restrict__ len8, uint32_t* __restrict__ slice_total, uint32_t* __restrict__ status) {
  __shared__ uint32_t len[DDF_SELL_WIN];
  __shared__ uint32_t slen[DDF_SELL_WIN];
  const int t = threadIdx.x;
  const int64_t row = (int64_t)blockIdx.x * DDF_SELL_WIN + t;
  const uint32_t mylen = row < nrows ? rowptr[row + 1] - rowptr[row] : 0u;
  if (mylen > 255u) atomicMax(status, 0xFFFFFFFFu);
  len[t] = mylen;
  __syncthreads();
  int rank = 0;
  for (int k = 0; k < DDF_SELL_WIN; k++) {
    const uint32_t lk = len[k];
    rank += (lk > mylen) || (lk == mylen && k < t);
  }
  const int64_t base = (int64_t)blockIdx.x * DDF_SELL_WIN;
  perm8[base + rank] = (uint8_t)t;
  len8[base + rank] = (uint8_t)(mylen > 255u ? 255u : mylen);
  slen[rank] = mylen;
  __syncthreads();
  if (t == 0) {
    // per-slice entry totals; the window's range is padded to a multiple of 8 entries so that its value
    // (8 B) and slot (2 B) streams start on 16-byte boundaries (one TMA bulk copy each, k_rows_pipe)
    uint32_t wtot = 0;
    for (int sl = 0; sl < DDF_SELL_WIN / 32; sl++) {
      uint32_t s = 0;
      for (int k = 0; k < 32; k++) s += slen[32 * sl + k];
      if (sl == DDF_SELL_WIN / 32 - 1) s += (8u - ((wtot + s) & 7u)) & 7u;
      wtot += s;
      slice_total[(int64_t)blockIdx.x * (DDF_SELL_WIN / 32) + sl] = s;
    }
    atomicMax(status + 1, wtot);
  }
}

// ---- pass B: per GROUP of K consecutive windows (K = 1 for long rows; K > 1 lets short rows -- 2D meshes --
// amortise the per-group bulk requests of the persistent pipeline): sort the touched columns -> runs and
// slots shared by the group; each window's entries into its own jagged record.
// dynamic shared memory: STAGE_KEYS sorted column indices
static __global__ void __launch_bounds__(DDF_SELL_WIN)
k_stage_fill(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
             int64_t nrows, int K, const uint8_t* __restrict__ perm8, const uint8_t* __restrict__ len8,
             const uint32_t* __restrict__ slice_ptr, const uint8_t* __restrict__ isb, uint32_t* __restrict__ tab,
             unsigned char* __restrict__ packed, unsigned char* __restrict__ meta,
             uint32_t* __restrict__ status /* [0] = max slots or ~0, [2] = max entries of a group */) {
  extern __shared__ uint32_t keys[];                                   // STAGE_KEYS
  __shared__ uint32_t rstart[STAGE_RMAX + 1], roff[STAGE_RMAX + 1];
  __shared__ uint32_t s_nruns, s_bad;
  const int t = threadIdx.x;
  const int64_t gbase = min((int64_t)blockIdx.x * K * DDF_SELL_WIN, nrows);
  const int64_t gend = min(gbase + (int64_t)K * DDF_SELL_WIN, nrows);
  const uint32_t e_lo = rowptr[gbase], e_hi = rowptr[gend];
  const uint32_t nent = e_hi - e_lo;
  uint32_t* wtab = tab + (int64_t)blockIdx.x * STAGE_TAB;
  const int64_t s0 = (int64_t)blockIdx.x * K * (DDF_SELL_WIN / 32);
  const uint32_t first_entry = slice_ptr[s0];
  const uint32_t stored = slice_ptr[s0 + (int64_t)K * (DDF_SELL_WIN / 32)] - first_entry;
  if (nent > (uint32_t)STAGE_KEYS) {
    if (t == 0) { wtab[0] = 0u; wtab[1] = 0u; wtab[2] = first_entry; wtab[3] = stored; atomicMax(status, 0xFFFFFFFFu); }
    return;
  }
  for (uint32_t e = t; e < (uint32_t)STAGE_KEYS; e += DDF_SELL_WIN) keys[e] = e < nent ? (uint32_t)col[e_lo + e] : 0xFFFFFFFFu;
  __syncthreads();
  for (int k = 2; k <= STAGE_KEYS; k <<= 1) {                           // bitonic sort
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < STAGE_KEYS; i += DDF_SELL_WIN) {
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
    // maximal runs of touched columns; runs start on even columns and have even length (16-byte TMA
    // granularity) and absorb gaps of up to STAGE_GAP untouched columns
    uint32_t nruns = 0, total = 0, cur_start = 0, cur_end = 0;
    bool open = false;
    for (uint32_t e = 0; e < nent; e++) {
      const uint32_t c = keys[e];
      if (open && c <= cur_end + STAGE_GAP) {
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
    const bool bad = nruns > (uint32_t)STAGE_RMAX || total > (uint32_t)STAGE_MAX_SLOTS;
    s_nruns = nruns;
    s_bad = bad ? 1u : 0u;
    wtab[0] = bad ? 0u : nruns;
    wtab[1] = bad ? 0u : total;
    wtab[2] = first_entry;                                              // the group's entry range
    wtab[3] = stored;
    if (!bad)
      for (uint32_t k = 0; k < nruns; k++) { wtab[4 + 2 * k] = rstart[k]; wtab[5 + 2 * k] = roff[k]; }
    atomicMax(status, bad ? 0xFFFFFFFFu : total);
    atomicMax(status + 2, stored);
  }
  __syncthreads();
  if (s_bad) return;
  const uint32_t nruns = s_nruns;
  const uint32_t lane = (uint32_t)t & 31u;
  for (int kk = 0; kk < K; kk++) {
    const int64_t win = (int64_t)blockIdx.x * K + kk;
    const int64_t wbase = win * DDF_SELL_WIN;
    const uint32_t wfirst = slice_ptr[win * (DDF_SELL_WIN / 32)];
    const uint32_t wstored = slice_ptr[(win + 1) * (DDF_SELL_WIN / 32)] - wfirst;
    // the window's record: values, then slots
    double* sval = reinterpret_cast<double*>(packed + (size_t)wfirst * 10);
    uint16_t* slot = reinterpret_cast<uint16_t*>(packed + (size_t)wfirst * 10 + (size_t)wstored * 8);
    // entries -> jagged slices (same addressing as the row kernels)
    const int64_t r = wbase + t;                       // sorted position
    const int64_t i = wbase + perm8[r];
    const uint32_t mylen = len8[r];
    const uint32_t maxlen = __shfl_sync(0xffffffffu, mylen, 0);
    uint32_t off = slice_ptr[r >> 5] - wfirst;         // entry offset inside the window's record
    const uint32_t rlo = i < nrows ? rowptr[i] : 0u;
    {
      unsigned char* wm = meta + (size_t)win * STAGE_META;
      if (t == 0) *reinterpret_cast<uint32_t*>(wm) = wstored;
      // entry offset of every slice inside the window's record (header words 4..11)
      if (t < DDF_SELL_WIN / 32) reinterpret_cast<uint32_t*>(wm)[4 + t] = slice_ptr[win * (DDF_SELL_WIN / 32) + t] - wfirst;
      wm[STAGE_META_HDR + t] = perm8[r];
      wm[STAGE_META_HDR + DDF_SELL_WIN + t] = (unsigned char)mylen;
      wm[STAGE_META_HDR + 2 * DDF_SELL_WIN + t] = (isb && wbase + t < nrows) ? isb[wbase + t] : 0;
    }
    for (uint32_t q = 0; q < maxlen; q++) {
      const bool act = q < mylen;
      const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, act));
      if (act) {
        const uint32_t c = (uint32_t)col[rlo + q];
        uint32_t lo = 0, hi = nruns;                   // last run with start <= c
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (rstart[mid] <= c) lo = mid; else hi = mid;
        }
        slot[(size_t)off + lane] = (uint16_t)(roff[lo] + (c - rstart[lo]));
        sval[(size_t)off + lane] = val[rlo + q];
      }
      off += cnt;
    }
  }
}

static int stage_build(ddf_ctx* ctx, const uint32_t* rowptr, const int32_t* col, const double* val, int64_t nrows,
                       int64_t nnz, const uint8_t* isb, int K, StageBufs& out) {
  out.release();
  if (K < 1 || K > 4) DDF_FAIL("stage_build: 1 <= K <= 4");
  const int64_t nwin_real = (nrows + DDF_SELL_WIN - 1) / DDF_SELL_WIN;
  const int64_t ngroups = (nwin_real + K - 1) / K;
  const int64_t nwin = ngroups * K;                     // windows incl. the empty ones that complete the last group
  const int64_t nslices = nwin * (DDF_SELL_WIN / 32);
  if (nwin == 0 || nnz == 0) return 0;
  DevBuf<uint32_t> tot, status;
  DevBuf<uint8_t> temp;
  DDF_CHECK(tot.alloc(ctx, nslices + 1));
  DDF_CHECK(status.alloc(ctx, 4));
  DDF_CHECK(out.ptr.alloc(ctx, nslices + 1));
  DDF_CHECK(out.perm8.alloc(ctx, nwin * DDF_SELL_WIN));
  DDF_CHECK(out.len8.alloc(ctx, nwin * DDF_SELL_WIN));
  DDF_CUDA(cudaMemsetAsync(status.p, 0, 16, ctx->stream));
  DDF_CUDA(cudaMemsetAsync(tot.p + nslices, 0, 4, ctx->stream));
  k_stage_sort<<<(int)nwin, DDF_SELL_WIN, 0, ctx->stream>>>(rowptr, nrows, out.perm8.p, out.len8.p, tot.p, status.p);
  DDF_LAUNCH_CHECK(ctx);
  size_t tb = 0;
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, tot.p, out.ptr.p, nslices + 1, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, tot.p, out.ptr.p, nslices + 1, ctx->stream));
  ctx->launches++;
  uint32_t stv[4] = {0, 0, 0, 0}, total = 0;
  DDF_CUDA(cudaMemcpyAsync(stv, status.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaMemcpyAsync(&total, out.ptr.p + nslices, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (stv[0] == 0xFFFFFFFFu) { out.release(); return 0; }     // a row longer than 255 entries
  out.entries = total;
  DDF_CHECK(out.tab.alloc(ctx, ngroups * STAGE_TAB));
  DDF_CHECK(out.packed.alloc(ctx, (size_t)total * 10 + 64));
  DDF_CHECK(out.meta.alloc(ctx, (size_t)nwin * STAGE_META));
  // padding entries are never read by a row, but they travel in the bulk copies: keep them defined
  DDF_CUDA(cudaMemsetAsync(out.packed.p, 0, (size_t)total * 10 + 64, ctx->stream));
  DDF_CUDA(cudaMemsetAsync(out.meta.p, 0, (size_t)nwin * STAGE_META, ctx->stream));
  const size_t smem = STAGE_KEYS * 4;
  k_stage_fill<<<(int)ngroups, DDF_SELL_WIN, smem, ctx->stream>>>(rowptr, col, val, nrows, K, out.perm8.p, out.len8.p,
                                                                 out.ptr.p, isb, out.tab.p, out.packed.p, out.meta.p,
                                                                 status.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaMemcpyAsync(stv, status.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (stv[0] == 0xFFFFFFFFu) { out.release(); return 0; }     // some group does not fit the staging limits
  out.max_slots = stv[0];
  out.max_entries = stv[2];                                   // of a group of K windows
  out.K = K;
  out.ngroups = ngroups;
  out.nnz = nnz;
  out.ok = true;
  return 0;
}

// ---- device side of the staged gather: issue the window's bulk copies, wait for them
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// warp 0 of the block calls this (all 32 lanes); `mbar` was initialised with count 1
__device__ __forceinline__ void stage_issue(const uint32_t* __restrict__ wtab, const double* __restrict__ u,
                                            double* stage, uint64_t* mbar, uint64_t policy) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t nruns = wtab[0], total = wtab[1];
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(total * 8u) : "memory");
  }
  __syncwarp();
  if (lane < nruns) {
    const uint32_t start = wtab[4 + 2 * lane], off = wtab[5 + 2 * lane];
    const uint32_t end = lane + 1 < nruns ? wtab[5 + 2 * (lane + 1)] : total;
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(stage + off)),
        "l"(u + start), "r"((end - off) * 8u), "r"(smem_u32(mbar)), "l"(policy)
        : "memory");
  }
}

#define PIPE_TD 8               // depth of the run-table ring of the persistent pipelines (even: wavefront2 splits it by task parity)

// one TMA bulk copy global -> shared, completion counted on `mbar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* mbar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(mbar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ void stage_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(mbar)),
      "r"(parity)
      : "memory");
}

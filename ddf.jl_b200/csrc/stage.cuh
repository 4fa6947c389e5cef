// stage.cuh -- "staged rows": jagged sliced storage with 16-bit staged column slots.
//
// The row kernels of the vertex Laplacian spend their bytes on 4-byte column indices and
// their memory-pipe slots on 8-byte random gathers of u.  On a mesh whose vertex numbering
// has locality (every generator of the reference: ManifoldConstructors.jl:196-284 numbers
// the Kuhn grid lexicographically), the columns touched by a window of 256 consecutive rows
// form a few contiguous runs (9 on the 3D Kuhn mesh: the x-rows (dy,dz) around the window).
// This format stores, per 256-row window,
//   * a run table  {nruns, nslots, first entry, entries, (start_col_k, slot_off_k)...}  -- runs start on even
//     columns and have even length, so each is one 16-byte-aligned TMA bulk copy
//     (cp.async.bulk global -> shared, completion on an mbarrier) of u into a staging buffer;
//   * per stored entry a 16-bit slot into that buffer + the 8-byte value: 10 B instead of 12 B;
//   * rows sorted by decreasing length inside the window (perm8, len8 in sorted order) and
//     every slice of 32 sorted rows stored as jagged diagonals: "column" q holds the entries
//     of the cnt_q = #{rows with len > q} leading lanes, packed -- no padding at all.
// The gather then reads shared memory; HBM sees every u element once per window that needs
// it (L2 hits after the first), every entry exactly once, fully coalesced.
// The order of the entries inside a row is unchanged: sums are bit-identical to the CSR form.
//
// Windows whose columns do not fit (more than STAGE_RMAX runs, more than STAGE_MAX_SLOTS
// slots, column span >= 2^18, a row longer than 255) make the whole build report
// "unstageable"; the caller then keeps the explicit-index SELL kernel.
#pragma once
#include <cub/cub.cuh>

#include "mesh.cuh"

#define STAGE_RMAX 30                    // runs per window
#define STAGE_TAB (4 + 2 * STAGE_RMAX)    // table words per window (256 B: one 16-byte-aligned bulk copy)
#define STAGE_BITS (1 << 18)             // bitmap span in columns
#define STAGE_WORDS (STAGE_BITS / 32)
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

// ---- pass B: per window: bitmap of touched columns -> runs, slots; entries into jagged slices
// dynamic shared memory: STAGE_WORDS bitmap words + STAGE_WORDS 16-bit slot prefixes
static __global__ void __launch_bounds__(DDF_SELL_WIN)
k_stage_fill(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
             int64_t nrows, const uint8_t* __restrict__ perm8, const uint8_t* __restrict__ len8,
             const uint32_t* __restrict__ slice_ptr, uint32_t* __restrict__ tab, uint16_t* __restrict__ slot,
             double* __restrict__ sval, uint32_t* __restrict__ status /* [0] = max slots or ~0 */) {
  extern __shared__ uint32_t sm_dyn[];
  uint32_t* bm = sm_dyn;                                               // STAGE_WORDS
  uint16_t* pre = reinterpret_cast<uint16_t*>(sm_dyn + STAGE_WORDS);   // STAGE_WORDS slot prefixes
  __shared__ uint32_t s_min, s_max, s_bad;
  __shared__ uint32_t s_runs[DDF_SELL_WIN], s_slots[DDF_SELL_WIN];     // per-thread counts -> exclusive prefixes
  const int t = threadIdx.x;
  const int64_t wbase = (int64_t)blockIdx.x * DDF_SELL_WIN;
  const int64_t row = wbase + t;
  const uint32_t lo = row < nrows ? rowptr[row] : 0u, hi = row < nrows ? rowptr[row + 1] : 0u;
  if (t == 0) { s_min = 0xFFFFFFFFu; s_max = 0u; s_bad = 0u; }
  for (int w = t; w < STAGE_WORDS; w += DDF_SELL_WIN) bm[w] = 0u;
  __syncthreads();
  if (hi > lo) {                         // columns ascend inside a row
    atomicMin(&s_min, (uint32_t)col[lo]);
    atomicMax(&s_max, (uint32_t)col[hi - 1]);
  }
  __syncthreads();
  uint32_t* wtab = tab + (int64_t)blockIdx.x * STAGE_TAB;
  if (s_min == 0xFFFFFFFFu) {            // a window of empty rows
    if (t == 0) { wtab[0] = 0u; wtab[1] = 0u; wtab[2] = slice_ptr[(int64_t)blockIdx.x * (DDF_SELL_WIN / 32)]; wtab[3] = 0u; }
    return;
  }
  const uint32_t cbase = s_min & ~63u;
  if (s_max - cbase >= (uint32_t)STAGE_BITS) {
    if (t == 0) { wtab[0] = 0u; wtab[1] = 0u; atomicMax(status, 0xFFFFFFFFu); }
    return;
  }
  for (uint32_t p = lo; p < hi; p++) {
    const uint32_t c = (uint32_t)col[p] - cbase;
    atomicOr(&bm[c >> 5], 1u << (c & 31u));
  }
  __syncthreads();
  // pairs: even start, even length (16-byte TMA granularity).  Each thread owns 32 consecutive words.
  uint32_t nslots = 0, nruns = 0;
  constexpr int WPT = STAGE_WORDS / DDF_SELL_WIN;
  for (int k = 0; k < WPT; k++) {
    const int w = t * WPT + k;
    uint32_t b = bm[w];
    b |= ((b | (b >> 1)) & 0x55555555u) * 3u;
    bm[w] = b;
    nslots += __popc(b);
  }
  __syncthreads();
  for (int k = 0; k < WPT; k++) {
    const int w = t * WPT + k;
    const uint32_t b = bm[w];
    const uint32_t prev = w > 0 ? bm[w - 1] >> 31 : 0u;
    nruns += __popc(b & ~((b << 1) | prev));
  }
  // block exclusive scan of (runs, slots)
  s_runs[t] = nruns;
  s_slots[t] = nslots;
  __syncthreads();
  if (t == 0) {
    uint32_t ar = 0, as = 0;
    for (int k = 0; k < DDF_SELL_WIN; k++) {
      const uint32_t r = s_runs[k], s = s_slots[k];
      s_runs[k] = ar;
      s_slots[k] = as;
      ar += r;
      as += s;
    }
    wtab[0] = ar;
    wtab[1] = as;
    wtab[2] = slice_ptr[(int64_t)blockIdx.x * (DDF_SELL_WIN / 32)];                  // the window's entry range
    wtab[3] = slice_ptr[(int64_t)(blockIdx.x + 1) * (DDF_SELL_WIN / 32)] - wtab[2];
    if (ar > (uint32_t)STAGE_RMAX || as > (uint32_t)STAGE_MAX_SLOTS) { s_bad = 1u; atomicMax(status, 0xFFFFFFFFu); }
    else atomicMax(status, as);
  }
  __syncthreads();
  if (s_bad) return;
  {
    uint32_t r = s_runs[t], s = s_slots[t];
    for (int k = 0; k < WPT; k++) {
      const int w = t * WPT + k;
      const uint32_t b = bm[w];
      const uint32_t prev = w > 0 ? bm[w - 1] >> 31 : 0u;
      pre[w] = (uint16_t)s;
      uint32_t starts = b & ~((b << 1) | prev);
      while (starts) {
        const int bit = __ffs(starts) - 1;
        starts &= starts - 1;
        wtab[4 + 2 * r] = cbase + 32u * (uint32_t)w + (uint32_t)bit;
        wtab[5 + 2 * r] = s + __popc(b & ((1u << bit) - 1u));
        r++;
      }
      s += __popc(b);
    }
  }
  __syncthreads();
  // entries -> jagged slices (same addressing as the row kernels)
  const int64_t r = wbase + t;                       // sorted position
  const int64_t i = wbase + perm8[r];
  const uint32_t mylen = len8[r];
  const uint32_t maxlen = __shfl_sync(0xffffffffu, mylen, 0);
  const uint32_t lane = (uint32_t)t & 31u;
  uint32_t off = slice_ptr[r >> 5];
  const uint32_t rlo = i < nrows ? rowptr[i] : 0u;
  for (uint32_t q = 0; q < maxlen; q++) {
    const bool act = q < mylen;
    const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, act));
    if (act) {
      const uint32_t c = (uint32_t)col[rlo + q] - cbase;
      slot[(size_t)off + lane] = (uint16_t)(pre[c >> 5] + __popc(bm[c >> 5] & ((1u << (c & 31u)) - 1u)));
      sval[(size_t)off + lane] = val[rlo + q];
    }
    off += cnt;
  }
}

static int stage_build(ddf_ctx* ctx, const uint32_t* rowptr, const int32_t* col, const double* val, int64_t nrows,
                       int64_t nnz, StageBufs& out) {
  out.release();
  const int64_t nwin = (nrows + DDF_SELL_WIN - 1) / DDF_SELL_WIN;
  const int64_t nslices = nwin * (DDF_SELL_WIN / 32);
  if (nwin == 0 || nnz == 0) return 0;
  DevBuf<uint32_t> tot, status;
  DevBuf<uint8_t> temp;
  DDF_CHECK(tot.alloc(ctx, nslices + 1));
  DDF_CHECK(status.alloc(ctx, 2));
  DDF_CHECK(out.ptr.alloc(ctx, nslices + 1));
  DDF_CHECK(out.perm8.alloc(ctx, nwin * DDF_SELL_WIN));
  DDF_CHECK(out.len8.alloc(ctx, nwin * DDF_SELL_WIN));
  DDF_CUDA(cudaMemsetAsync(status.p, 0, 8, ctx->stream));
  DDF_CUDA(cudaMemsetAsync(tot.p + nslices, 0, 4, ctx->stream));
  k_stage_sort<<<(int)nwin, DDF_SELL_WIN, 0, ctx->stream>>>(rowptr, nrows, out.perm8.p, out.len8.p, tot.p, status.p);
  DDF_LAUNCH_CHECK(ctx);
  size_t tb = 0;
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, tot.p, out.ptr.p, nslices + 1, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, tot.p, out.ptr.p, nslices + 1, ctx->stream));
  ctx->launches++;
  uint32_t st = 0, stv[2] = {0, 0}, total = 0;
  DDF_CUDA(cudaMemcpyAsync(stv, status.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaMemcpyAsync(&total, out.ptr.p + nslices, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (stv[0] == 0xFFFFFFFFu) { out.release(); return 0; }     // a row longer than 255 entries
  out.max_entries = stv[1];
  out.entries = total;
  DDF_CHECK(out.tab.alloc(ctx, nwin * STAGE_TAB));
  DDF_CHECK(out.slot.alloc(ctx, (size_t)total + 8));
  DDF_CHECK(out.val.alloc(ctx, (size_t)total + 8));
  // padding entries are never read by a row, but they travel in the bulk copies: keep them defined
  DDF_CUDA(cudaMemsetAsync(out.slot.p, 0, ((size_t)total + 8) * 2, ctx->stream));
  DDF_CUDA(cudaMemsetAsync(out.val.p, 0, ((size_t)total + 8) * 8, ctx->stream));
  const size_t smem = STAGE_WORDS * 4 + STAGE_WORDS * 2;
  static bool attr_set = false;
  if (!attr_set) {
    DDF_CUDA(cudaFuncSetAttribute(k_stage_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  k_stage_fill<<<(int)nwin, DDF_SELL_WIN, smem, ctx->stream>>>(rowptr, col, val, nrows, out.perm8.p, out.len8.p, out.ptr.p,
                                                              out.tab.p, out.slot.p, out.val.p, status.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaMemcpyAsync(&st, status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (st == 0xFFFFFFFFu) { out.release(); return 0; }     // some window does not fit the staging limits
  out.max_slots = st;
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

#define PIPE_TD 8               // depth of the run-table ring of the persistent pipelines

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

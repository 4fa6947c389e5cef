// sell.cuh -- sliced-ELL with local stencil sorting (SELL-32-256) for variable-length row
// operators, with optional lossless compression of the index stream.
//
// Rows are grouped in windows of 256 (one thread block).  Inside a window the rows are
// stably sorted by (decreasing length, stencil signature) -- perm8: sorted position -> row
// offset in the window, one byte per row -- so that each slice of 32 sorted rows holds rows
// of equal length and, on meshes with a repeating stencil, of identical column offsets
// j - i.  (On the mirrored Kuhn meshes the vertex degree and stencil alternate with the
// vertex parity; without the sort every slice would be padded to the maximum.)
// Inside a slice the entries are stored column-major: entry q of sorted row r lives at
// ptr[r / 32] + q*32 + (r & 31); short rows are padded with DDF_SELL_PAD.  One thread per
// row then reads fully coalesced 128-byte (index) / 256-byte (value) requests.
//
// Index compression (compress = true): for every slice and column q < 32 where all 32 rows
// hold an entry with the same offset j - i, the 32 indices are replaced by that ONE offset
// (umask bit q set); other columns stay explicit.  The index stream of a slice is the
// concatenation over q of (1 word | 32 words); a column's position follows from popc(umask).
// On structured meshes this removes ~85-97 % of the index bytes; on irregular meshes it
// degrades to the explicit form plus 8 bytes per slice.
//
// The order of the entries inside a row is unchanged: the per-row sequential sums are
// bit-identical to the CSR form.
#pragma once
#include <cub/cub.cuh>

#include "mesh.cuh"   // SellBufs, DDF_SELL_PAD, DDF_SELL_WIN

// length and stencil signature (hash of the column offsets) of every row
static __global__ void k_sell_sig(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col,
                                  int64_t nrows, uint32_t* __restrict__ sig) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  uint32_t h = 2166136261u;
  for (uint32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
    const uint32_t d = col[p] - (uint32_t)i;
    h = (h ^ d) * 16777619u;
    h ^= h >> 15;
  }
  sig[i] = h;
}

// one block per window: stable sort by (decreasing row length, signature), slice widths
static __global__ void __launch_bounds__(DDF_SELL_WIN)
k_sell_sort(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ sig, int64_t nrows,
            uint8_t* __restrict__ perm8, uint32_t* __restrict__ w32 /* 8 per window */) {
  __shared__ uint32_t len[DDF_SELL_WIN];
  __shared__ uint32_t sg[DDF_SELL_WIN];
  __shared__ uint32_t slen[DDF_SELL_WIN];
  const int t = threadIdx.x;
  const int64_t row = (int64_t)blockIdx.x * DDF_SELL_WIN + t;
  const uint32_t mylen = row < nrows ? rowptr[row + 1] - rowptr[row] : 0u;
  const uint32_t mysig = (sig && row < nrows) ? sig[row] : 0u;
  len[t] = mylen;
  sg[t] = mysig;
  __syncthreads();
  int rank = 0;
  for (int k = 0; k < DDF_SELL_WIN; k++) {
    const uint32_t lk = len[k], sk = sg[k];
    rank += (lk > mylen) || (lk == mylen && (sk < mysig || (sk == mysig && k < t)));
  }
  perm8[(int64_t)blockIdx.x * DDF_SELL_WIN + rank] = (uint8_t)t;
  slen[rank] = mylen;
  __syncthreads();
  if (t < DDF_SELL_WIN / 32) w32[(int64_t)blockIdx.x * (DDF_SELL_WIN / 32) + t] = 32u * slen[32 * t];   // first of a slice is its max
}

template <bool HAS_VAL>
static __global__ void __launch_bounds__(DDF_SELL_WIN)
k_sell_fill(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col, const double* __restrict__ val,
            const uint32_t* __restrict__ sell_ptr, const uint8_t* __restrict__ perm8, int64_t nrows,
            uint32_t* __restrict__ scol, double* __restrict__ sval) {
  const int64_t r = (int64_t)blockIdx.x * DDF_SELL_WIN + threadIdx.x;      // sorted position
  const int64_t row = (int64_t)blockIdx.x * DDF_SELL_WIN + perm8[r];
  const int64_t s = r >> 5;
  const uint32_t base = sell_ptr[s], w = (sell_ptr[s + 1] - base) >> 5;
  const uint32_t lo = row < nrows ? rowptr[row] : 0, len = row < nrows ? rowptr[row + 1] - lo : 0;
  for (uint32_t q = 0; q < w; q++) {
    const size_t pos = (size_t)base + (size_t)q * 32 + (r & 31);
    scol[pos] = q < len ? col[lo + q] : DDF_SELL_PAD;
    if (HAS_VAL) sval[pos] = q < len ? val[lo + q] : 0.0;
  }
}

// one warp per slice: which columns have a uniform offset j - i, and the compressed stream size
static __global__ void __launch_bounds__(DDF_SELL_WIN)
k_sell_uniform(const uint32_t* __restrict__ sell_ptr, const uint8_t* __restrict__ perm8,
               const uint32_t* __restrict__ scol, int64_t nslices, uint32_t* __restrict__ umask,
               uint32_t* __restrict__ isize) {
  const int64_t s = ((int64_t)blockIdx.x * DDF_SELL_WIN + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s > nslices) return;
  if (s == nslices) { if (lane == 0) isize[s] = 0; return; }
  const int64_t r = s * 32 + lane;
  const int64_t i = (r & ~(int64_t)(DDF_SELL_WIN - 1)) + perm8[r];
  const uint32_t base = sell_ptr[s], w = (sell_ptr[s + 1] - base) >> 5;
  uint32_t um = 0, words = 0;
  for (uint32_t q = 0; q < w; q++) {
    const uint32_t j = scol[(size_t)base + (size_t)q * 32 + lane];
    const uint32_t d = j - (uint32_t)i;
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
    const bool uni = q < 32 && __all_sync(0xffffffffu, j != DDF_SELL_PAD && d == d0);
    if (uni) um |= 1u << q;
    words += uni ? 1u : 32u;
  }
  if (lane == 0) { umask[s] = um; isize[s] = words; }
}

static __global__ void __launch_bounds__(DDF_SELL_WIN)
k_sell_pack(const uint32_t* __restrict__ sell_ptr, const uint8_t* __restrict__ perm8,
            const uint32_t* __restrict__ scol, const uint32_t* __restrict__ umask, const uint32_t* __restrict__ iptr,
            int64_t nslices, uint32_t* __restrict__ cidx) {
  const int64_t s = ((int64_t)blockIdx.x * DDF_SELL_WIN + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int64_t r = s * 32 + lane;
  const int64_t i = (r & ~(int64_t)(DDF_SELL_WIN - 1)) + perm8[r];
  const uint32_t base = sell_ptr[s], w = (sell_ptr[s + 1] - base) >> 5;
  const uint32_t um = umask[s];
  uint32_t off = iptr[s];
  for (uint32_t q = 0; q < w; q++) {
    const uint32_t j = scol[(size_t)base + (size_t)q * 32 + lane];
    if (q < 32 && ((um >> q) & 1u)) {
      if (lane == 0) cidx[off] = j - (uint32_t)i;
      off += 1;
    } else {
      cidx[off + lane] = j;
      off += 32;
    }
  }
}

// explicit 32-bit columns -> 16-bit offsets column - row (0x8000 = padding); *bad is set when an offset does not fit
static __global__ void __launch_bounds__(DDF_SELL_WIN)
k_sell_to16(const uint32_t* __restrict__ sell_ptr, const uint8_t* __restrict__ perm8, const uint32_t* __restrict__ scol,
            uint16_t* __restrict__ col16, uint32_t* __restrict__ bad) {
  const int64_t r = (int64_t)blockIdx.x * DDF_SELL_WIN + threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * DDF_SELL_WIN + perm8[r];
  const int64_t s = r >> 5;
  const uint32_t base = sell_ptr[s] + (uint32_t)(r & 31), w = (sell_ptr[s + 1] - sell_ptr[s]) >> 5;
  for (uint32_t q = 0; q < w; q++) {
    const size_t pos = (size_t)base + (size_t)q * 32;
    const uint32_t j = scol[pos];
    uint16_t d = 0x8000u;
    if (j != DDF_SELL_PAD) {
      const int64_t off = (int64_t)j - i;
      if (off < -32767 || off > 32767) *bad = 1u;
      d = (uint16_t)(int16_t)off;
    }
    col16[pos] = d;
  }
}

// position of column q in the compressed index stream of a slice
__device__ __forceinline__ uint32_t sell_col_offset(uint32_t um, uint32_t q) {
  const uint32_t before = q >= 32 ? um : (um & ((1u << q) - 1u));
  return 32u * q - 31u * (uint32_t)__popc(before);
}

// Build SELL-32-256 from a row-CSR (col may carry a sign in bit 31; val may be null).
static int sell_build(ddf_ctx* ctx, const uint32_t* rowptr, const uint32_t* col, const double* val, int64_t nrows,
                      SellBufs& out, bool compress, bool sort_sig = true, bool try16 = false) {
  const int64_t nwin = (nrows + DDF_SELL_WIN - 1) / DDF_SELL_WIN;
  const int64_t nslices = nwin * (DDF_SELL_WIN / 32);
  DevBuf<uint32_t> w32, sig;
  DevBuf<uint8_t> temp;
  out.release();
  DDF_CHECK(w32.alloc(ctx, nslices + 1));
  DDF_CHECK(out.ptr.alloc(ctx, nslices + 1));
  DDF_CHECK(out.perm8.alloc(ctx, nwin * DDF_SELL_WIN));
  if (nwin == 0) { out.padded = 0; DDF_CUDA(cudaMemsetAsync(out.ptr.p, 0, 4, ctx->stream)); return 0; }
  DDF_CHECK(sig.alloc(ctx, nrows));
  int grid;
  DDF_CHECK(grid_for(nrows, 256, &grid));
  k_sell_sig<<<grid, 256, 0, ctx->stream>>>(rowptr, col, nrows, sig.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaMemsetAsync(w32.p + nslices, 0, 4, ctx->stream));
  k_sell_sort<<<(int)nwin, DDF_SELL_WIN, 0, ctx->stream>>>(rowptr, sort_sig ? sig.p : nullptr, nrows, out.perm8.p, w32.p);
  DDF_LAUNCH_CHECK(ctx);
  size_t tb = 0;
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, w32.p, out.ptr.p, nslices + 1, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, w32.p, out.ptr.p, nslices + 1, ctx->stream));
  ctx->launches++;
  uint32_t tot = 0;
  DDF_CUDA(cudaMemcpyAsync(&tot, out.ptr.p + nslices, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  out.padded = tot;
  out.index_words = tot;
  DDF_CHECK(out.col.alloc(ctx, tot));
  if (val) DDF_CHECK(out.val.alloc(ctx, tot));
  if (!tot) return 0;
  if (val) k_sell_fill<true><<<(int)nwin, DDF_SELL_WIN, 0, ctx->stream>>>(rowptr, col, val, out.ptr.p, out.perm8.p, nrows, out.col.p, out.val.p);
  else k_sell_fill<false><<<(int)nwin, DDF_SELL_WIN, 0, ctx->stream>>>(rowptr, col, nullptr, out.ptr.p, out.perm8.p, nrows, out.col.p, nullptr);
  DDF_LAUNCH_CHECK(ctx);
  if (try16 && !compress) {          // 16-bit offsets when the matrix is banded enough (|col - row| <= 32767)
    DevBuf<uint32_t> bad;
    DDF_CHECK(bad.alloc(ctx, 1));
    DDF_CUDA(cudaMemsetAsync(bad.p, 0, 4, ctx->stream));
    DDF_CHECK(out.col16.alloc(ctx, tot));
    k_sell_to16<<<(int)nwin, DDF_SELL_WIN, 0, ctx->stream>>>(out.ptr.p, out.perm8.p, out.col.p, out.col16.p, bad.p);
    DDF_LAUNCH_CHECK(ctx);
    uint32_t hb = 0;
    DDF_CUDA(cudaMemcpyAsync(&hb, bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hb) out.col16.release();
    else { out.col.release(); out.idx16 = true; out.index_words = (tot + 1) / 2; }
  }
  if (!compress) return 0;
  // ---- index compression
  DevBuf<uint32_t> isize;
  DDF_CHECK(isize.alloc(ctx, nslices + 1));
  DDF_CHECK(out.umask.alloc(ctx, nslices));
  DDF_CHECK(out.iptr.alloc(ctx, nslices + 1));
  const int wgrid = (int)ceil_div64((nslices + 1) * 32, DDF_SELL_WIN);
  k_sell_uniform<<<wgrid, DDF_SELL_WIN, 0, ctx->stream>>>(out.ptr.p, out.perm8.p, out.col.p, nslices, out.umask.p, isize.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, isize.p, out.iptr.p, nslices + 1, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, isize.p, out.iptr.p, nslices + 1, ctx->stream));
  ctx->launches++;
  uint32_t words = 0;
  DDF_CUDA(cudaMemcpyAsync(&words, out.iptr.p + nslices, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  DDF_CHECK(out.cidx.alloc(ctx, words));
  k_sell_pack<<<wgrid, DDF_SELL_WIN, 0, ctx->stream>>>(out.ptr.p, out.perm8.p, out.col.p, out.umask.p, out.iptr.p, nslices, out.cidx.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  out.col.release();                 // the explicit index array is no longer needed
  out.index_words = words;
  out.compressed = true;
  return 0;
}

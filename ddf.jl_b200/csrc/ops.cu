// ops.cu -- DEC operators (src/ManifoldOps.jl), the Op ring (src/Ops.jl) and
// Op*Fun (src/Ops.jl:273-276 -> SparseArrays `mul!`), as device kernels.
//
// Kernels (all HBM-bound, Float64, integer +-1 coefficients never stored):
//   K2 k_apply_fixed<W>   y = d_k x for the primal derivative d_k = adjoint(d_{k+1}
//        boundary) (ManifoldOps.jl:46).  The reference holds Adjoint{Int8,CSC}; the
//        parent CSC's columns are the rows of d_k, all of width W = k+2 with the
//        constant sign pattern (-1)^(k+1+p).  One thread owns 4 consecutive rows:
//        W 128-bit streaming index loads, 4W independent gathers in flight, the
//        sum per row in the reference's order (tmp = 0; tmp += v_j x[c_j], j ascending),
//        two 128-bit streaming stores.  Bit-exact with the oracle.
//   K3 k_apply_csrsign    y = d_R x for boundary / dual derivative (plain CSC in the
//        reference: zero y, scatter by ascending column == per-row sequential sum over
//        ascending parents).  Row-CSR with the sign in bit 31 of the column index.
//   K4 diagonal Hodge stars fused as prologue (x[c]*pre[c]) / epilogue (post[r]*sum)
//        of K2/K3, or k_apply_diag when they stand alone.
//   k_apply_csrf64        arbitrary assembled Float64 operators (ddf_op_from_csc),
//        separately rounded multiply and add in ascending-column order (bit-exact
//        with SparseArrays' CSC `mul!`).
#include <cub/cub.cuh>

#include <algorithm>
#include <memory>

#include "mesh.cuh"
#include "sell.cuh"
#include "stage.cuh"      // bulk_g2s / stage_wait / smem_u32 (TMA staging helpers)

// Tuning (DdfTune in the ctx).  prefetch: blocks ahead a block prefetches index words into L2 (0 = off); measured at
// C4 (profiles/r01_tuning.md) any distance from 148 to 2368 blocks gives the same gain; 4 x 148 keeps the request
// well ahead of the block scheduler.
int ddf_laplace0_apply(ddf_mesh* m, const double* x, double* y);   // wave.cu

enum OpKind { OPK_ZERO = 0, OPK_FIXED, OPK_CSRSIGN, OPK_DIAG, OPK_CSRF64, OPK_LAPLACE0, OPK_CHAIN, OPK_SUM };

struct OpNode;
typedef std::shared_ptr<OpNode> OpRef;

struct OpNode {
  OpKind kind = OPK_ZERO;
  ddf_ctx* ctx = nullptr;
  ddf_mesh* mesh = nullptr;       // borrowed (FIXED / CSRSIGN / LAPLACE0 / mesh diagonals)
  int64_t nrows = 0, ncols = 0;
  // FIXED / CSRSIGN: boundary level R (tables live in the mesh)
  int R = 0;
  // DIAG: y = sign * diag .* x ; diag == nullptr means the identity scaled by `scale`
  const double* diag = nullptr;
  DevBuf<double> diag_own;
  double scale = 1.0;             // DIAG identity scale / CHAIN scalar
  // CSRF64
  DevBuf<uint32_t> rowptr;
  DevBuf<int32_t> col;
  DevBuf<double> val;
  int64_t nnz = 0;
  SellBufs sell;                  // CSRF64: SELL-32-256 copy of the rows, built at the first apply of a large operator
  int sell_state = 0;             // 0 not tried, 1 built, 2 not worth it / failed
  // CHAIN (factors left to right) / SUM (coef, term)
  std::vector<OpRef> factors;
  std::vector<double> coefs;
  // scratch vectors reused across applies
  std::vector<std::unique_ptr<DevBuf<double>>> temps;
  // The expression as the caller wrote it (association kept): `factors` above is flattened for the lazy
  // apply, but the ASSEMBLED matrix of the reference is rounded and chopped after every binary product /
  // sum (Ops.jl:25,228-232), so ddf_op_to_csc evaluates this tree instead.
  std::vector<OpRef> tree_ops;       // CHAIN: operands left to right ; SUM: terms
  std::vector<double> tree_coefs;    // SUM: coefficient of every term
  double tree_scale = 1.0;           // CHAIN: scalar applied to the assembled product (a * A, Ops.jl:240-243)
  OpRef adj_of;                      // this node is adjoint(adj_of)
  // ddf_op_restrict_rows: apply writes rows [row_lo, row_hi) only (row_hi < 0: all rows).  Slab meshes apply d_k /
  // star_k to the rows they OWN; the ghost rows belong to the neighbour rank.
  int64_t row_lo = 0, row_hi = -1;
};

struct ddf_op {
  OpRef node;
};

// shallow clone of a leaf node (payload pointers are borrowed; the source is kept
// alive through `factors`)
static OpRef clone_leaf(const OpRef& src) {
  OpRef o = std::make_shared<OpNode>();
  o->kind = src->kind; o->ctx = src->ctx; o->mesh = src->mesh;
  o->nrows = src->nrows; o->ncols = src->ncols; o->R = src->R;
  o->diag = src->diag; o->scale = src->scale;
  o->factors.push_back(src);
  return o;
}

static ddf_op* wrap(OpRef n) {
  ddf_op* o = new ddf_op();
  o->node = n;
  return o;
}

// ======================================================================= kernels
template <int W, bool PRE, bool POST>
__global__ void __launch_bounds__(256)
k_apply_fixed(const int32_t* __restrict__ tab, int64_t nrows, int first_negative, const double* __restrict__ x,
              const double* __restrict__ pre, const double* __restrict__ post, double alpha,
              double* __restrict__ y) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = t * 4;
  const uint64_t pol_keep = make_policy_evict_last();     // gathered x: keep in L2
  const uint64_t pol_stream = make_policy_evict_first();  // index table / outputs: touched once
  if (r0 >= nrows) return;
  if (r0 + 3 < nrows) {
    int idx[4 * W];
    const int4* tp = reinterpret_cast<const int4*>(tab) + t * W;
#pragma unroll
    for (int q = 0; q < W; q++) {
      int4 v = ld_l1_int4_hint(tp + q, pol_stream);   // L1-allocating: the W loads of a thread share sectors
      idx[4 * q + 0] = v.x; idx[4 * q + 1] = v.y; idx[4 * q + 2] = v.z; idx[4 * q + 3] = v.w;
    }
    double xv[4 * W];
    // evict_last on the gathered x pays for W <= 3 (d1: 84 -> 97 % of the copy peak); with 16 gathers per thread
    // (W = 4) the plain read-only load is 2-3 % faster and moves the same DRAM bytes (profiles/r01_tuning.md)
#pragma unroll
    for (int q = 0; q < 4 * W; q++) xv[q] = W >= 4 ? __ldg(x + idx[q]) : ld_keep_f64(x + idx[q], pol_keep);
    if (PRE) {
#pragma unroll
      for (int q = 0; q < 4 * W; q++) xv[q] = __dmul_rn(ld_keep_f64(pre + idx[q], pol_keep), xv[q]);
    }
    double acc[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      double a = 0.0;                                   // tmp = zero(eltype(C))
#pragma unroll
      for (int p = 0; p < W; p++) {
        const bool neg = ((p & 1) != 0) != (first_negative != 0);
        a = neg ? __dsub_rn(a, xv[r * W + p]) : __dadd_rn(a, xv[r * W + p]);
      }
      acc[r] = a;
    }
    if (POST) {
      const double2 p0 = ld_stream_f64x2_hint(reinterpret_cast<const double2*>(post + r0), pol_stream);
      const double2 p1 = ld_stream_f64x2_hint(reinterpret_cast<const double2*>(post + r0 + 2), pol_stream);
      acc[0] = __dmul_rn(p0.x, acc[0]); acc[1] = __dmul_rn(p0.y, acc[1]);
      acc[2] = __dmul_rn(p1.x, acc[2]); acc[3] = __dmul_rn(p1.y, acc[3]);
    }
    if (alpha != 1.0) {
#pragma unroll
      for (int r = 0; r < 4; r++) acc[r] = __dmul_rn(alpha, acc[r]);
    }
    st_hint_f64x2(reinterpret_cast<double2*>(y + r0), acc[0], acc[1], pol_stream);
    st_hint_f64x2(reinterpret_cast<double2*>(y + r0 + 2), acc[2], acc[3], pol_stream);
  } else {
    for (int64_t r = r0; r < nrows; r++) {
      double a = 0.0;
#pragma unroll
      for (int p = 0; p < W; p++) {
        const int c = tab[r * W + p];
        double v = x[c];
        if (PRE) v = __dmul_rn(pre[c], v);
        const bool neg = ((p & 1) != 0) != (first_negative != 0);
        a = neg ? __dsub_rn(a, v) : __dadd_rn(a, v);
      }
      if (POST) a = __dmul_rn(post[r], a);
      if (alpha != 1.0) a = __dmul_rn(alpha, a);
      y[r] = a;
    }
  }
}

// Row-CSR with sign bit; one thread per row, sequential ascending-parent sum.
template <bool PRE, bool POST>
__global__ void __launch_bounds__(256)
k_apply_csrsign(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col, int64_t nrows,
                const double* __restrict__ x, const double* __restrict__ pre, const double* __restrict__ post,
                double alpha, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const uint32_t lo = rowptr[i], hi = rowptr[i + 1];
  double a = 0.0;
  for (uint32_t p = lo; p < hi; p++) {
    const uint32_t cw = __ldg(col + p);
    const uint32_t c = cw & 0x7fffffffu;
    double v = __ldg(x + c);
    if (PRE) v = __dmul_rn(__ldg(pre + c), v);
    a = (cw >> 31) ? __dsub_rn(a, v) : __dadd_rn(a, v);
  }
  if (POST) a = __dmul_rn(post[i], a);
  if (alpha != 1.0) a = __dmul_rn(alpha, a);
  y[i] = a;
}

// Rows of at most two entries (the boundary matrix of the TOP simplices by rows: a face has one or two parents).
// The row-CSR costs two dependent loads (offsets, then entries); here a row is one (entry0, entry1) pair --
// entry1 = 0xFFFFFFFF when absent -- loaded with a single coalesced 8-byte access; lane = consecutive rows.
__global__ void k_build_pairs(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col, int64_t nrows,
                              uint32_t* __restrict__ pair, uint32_t* __restrict__ too_long) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const uint32_t lo = rowptr[i], hi = rowptr[i + 1];
  if (hi - lo > 2u) { *too_long = 1u; return; }
  pair[2 * i] = hi > lo ? col[lo] : 0xFFFFFFFFu;
  pair[2 * i + 1] = hi > lo + 1 ? col[lo + 1] : 0xFFFFFFFFu;
}

template <bool PRE, bool POST, int RPW>
__device__ __forceinline__ void
pairsign_body(const uint32_t* __restrict__ pair, int64_t nrows, const double* __restrict__ x,
              const double* __restrict__ pre, const double* __restrict__ post, double alpha, double* __restrict__ y,
              int pf_blocks) {
  const uint64_t pol_keep = make_policy_evict_last();
  const uint64_t pol_stream = make_policy_evict_first();
  if (pf_blocks && threadIdx.x == 0) {       // pairs of the block that runs pf_blocks later: 256 * RPW rows, contiguous
    const int64_t rp = ((int64_t)blockIdx.x + pf_blocks) * (256 * RPW);
    if (rp + 256 * RPW <= nrows) bulk_prefetch_l2(pair + 2 * rp, 256 * RPW * 8);
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t r0 = warp * (32 * RPW) + lane;
  if (warp * (32 * RPW) >= nrows) return;
  int2 cw[RPW];
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const int64_t r = r0 + 32 * k;
    cw[k] = make_int2(-1, -1);
    if (r < nrows) cw[k] = ld_stream_int2(reinterpret_cast<const int2*>(pair) + r);
  }
  double xv[RPW][2];
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const uint32_t c0 = (uint32_t)cw[k].x, c1 = (uint32_t)cw[k].y;
    xv[k][0] = c0 != 0xFFFFFFFFu ? ld_keep_f64(x + (c0 & 0x7fffffffu), pol_keep) : 0.0;
    xv[k][1] = c1 != 0xFFFFFFFFu ? ld_keep_f64(x + (c1 & 0x7fffffffu), pol_keep) : 0.0;
    if (PRE) {
      if (c0 != 0xFFFFFFFFu) xv[k][0] = __dmul_rn(ld_keep_f64(pre + (c0 & 0x7fffffffu), pol_keep), xv[k][0]);
      if (c1 != 0xFFFFFFFFu) xv[k][1] = __dmul_rn(ld_keep_f64(pre + (c1 & 0x7fffffffu), pol_keep), xv[k][1]);
    }
  }
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const int64_t r = r0 + 32 * k;
    if (r >= nrows) continue;
    const uint32_t c0 = (uint32_t)cw[k].x, c1 = (uint32_t)cw[k].y;
    double a = 0.0;
    if (c0 != 0xFFFFFFFFu) a = (c0 >> 31) ? __dsub_rn(a, xv[k][0]) : __dadd_rn(a, xv[k][0]);
    if (c1 != 0xFFFFFFFFu) a = (c1 >> 31) ? __dsub_rn(a, xv[k][1]) : __dadd_rn(a, xv[k][1]);
    if (POST) a = __dmul_rn(ld_stream_f64(post + r), a);
    if (alpha != 1.0) a = __dmul_rn(alpha, a);
    st_hint_f64(y + r, a, pol_stream);
  }
}

template <bool PRE, bool POST, int RPW>
__global__ void __launch_bounds__(256)
k_apply_pairsign(const uint32_t* __restrict__ pair, int64_t nrows, const double* __restrict__ x,
                 const double* __restrict__ pre, const double* __restrict__ post, double alpha, double* __restrict__ y,
                 int pf_blocks = 0) {
  pairsign_body<PRE, POST, RPW>(pair, nrows, x, pre, post, alpha, y, pf_blocks);
}
// The unfused apply capped at 32 registers (8 resident blocks per SM instead of 6): 0.652 -> 0.599 ms at C4, no spills.
// (`__launch_bounds__(256, 1)` is NOT the same as `__launch_bounds__(256)`: with it ptxas spends more registers and
// the kernel slows to 0.77 ms -- hence two wrappers instead of a template parameter.)
template <int RPW>
__global__ void __launch_bounds__(256, 8)
k_apply_pairsign32(const uint32_t* __restrict__ pair, int64_t nrows, const double* __restrict__ x, double* __restrict__ y,
                   int pf_blocks) {
  pairsign_body<false, false, RPW>(pair, nrows, x, nullptr, nullptr, 1.0, y, pf_blocks);
}

// The same operator from the SELL-32 copy of the row-CSR (sell.cuh): coalesced index loads.
// UP: the SELL copy holds only the LOWER entries of a row; its upper entries -- the cofaces (sigma, c) that extend
// the row's simplex sigma at the end -- are the contiguous ids [up[i], up[i+1]) and all carry the sign (-1)^R
// (upper_build below).  They come last in ascending-column order, so the sum keeps the reference's order.
// STAGE: the upper runs of the 256 rows of a window are ONE contiguous piece of x (and of pre): thread 0 fetches it
// with a single TMA bulk copy per array while all threads gather their lower entries; the upper sums then read
// shared memory.  `cap` = doubles of shared memory per array (the longest window run of the mesh, rounded).
template <bool PRE, bool POST, int BATCH = 8, bool UP = false, bool STAGE = false, int CB = BATCH>
__global__ void __launch_bounds__(256)
k_apply_sellsign(const uint32_t* __restrict__ sell_ptr, const uint8_t* __restrict__ perm8,
                 const uint32_t* __restrict__ scol, int64_t nrows, const double* __restrict__ x,
                 const double* __restrict__ pre, const double* __restrict__ post, double alpha,
                 double* __restrict__ y, const uint32_t* __restrict__ up = nullptr, int up_negative = 0,
                 uint32_t cap = 0, int pf_blocks = 0) {
  if (pf_blocks && threadIdx.x == 0 && blockIdx.x + pf_blocks < gridDim.x) {
    // index words of the window that runs pf_blocks later (slice offsets are multiples of 32 entries = 128 bytes)
    const int64_t sp = ((int64_t)blockIdx.x + pf_blocks) * (DDF_SELL_WIN / 32);
    const uint32_t pa = sell_ptr[sp], pb = sell_ptr[sp + DDF_SELL_WIN / 32];
    if (pb > pa) bulk_prefetch_l2(scol + pa, (pb - pa) * 4u);
  }
  const int64_t r = (int64_t)blockIdx.x * DDF_SELL_WIN + threadIdx.x;        // sorted position
  const int64_t i = (int64_t)blockIdx.x * DDF_SELL_WIN + perm8[r];           // the row it holds
  const uint64_t pol_stream = make_policy_evict_first();
  const uint64_t pol_keep = make_policy_evict_last();
  extern __shared__ __align__(16) double sx[];
  __shared__ uint64_t mbar;
  uint32_t a0 = 0, staged = 0;
  if (STAGE) {
    const int64_t w0 = (int64_t)blockIdx.x * DDF_SELL_WIN;
    const int64_t w1 = w0 + DDF_SELL_WIN < nrows ? w0 + DDF_SELL_WIN : nrows;
    const uint32_t ua = up[w0], ub = up[w1];
    a0 = ua & ~1u;                                        // 16-byte aligned start, even count
    staged = (ub - a0 + 1u) & ~1u;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      if (staged) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(staged * (PRE ? 16u : 8u)) : "memory");
        bulk_g2s(sx, x + a0, staged * 8u, &mbar, pol_keep);
        if (PRE) bulk_g2s(sx + cap, pre + a0, staged * 8u, &mbar, pol_keep);
      }
    }
    __syncthreads();                                      // the barrier is initialised before anyone waits on it
  }
  const int64_t s = r >> 5;
  const uint32_t base = sell_ptr[s] + (uint32_t)(r & 31);
  const uint32_t w = (sell_ptr[s + 1] - sell_ptr[s]) >> 5;
  double a = 0.0;
  // CB index words are fetched at once, their gathers issued in groups of BATCH (CB > BATCH: the index loads of a
  // row of up to CB entries are ONE round trip although only BATCH gathered values are held at a time)
  for (uint32_t q0 = 0; q0 < w; q0 += CB) {
    uint32_t cw[CB];
#pragma unroll
    for (int q = 0; q < CB; q++)
      cw[q] = (q0 + q < w) ? (uint32_t)ld_stream_int(reinterpret_cast<const int*>(scol) + (size_t)base + (size_t)(q0 + q) * 32)
                           : DDF_SELL_PAD;
#pragma unroll
    for (int g = 0; g < CB; g += BATCH) {
      if (g > 0 && q0 + g >= w) break;
      double xv[BATCH];
#pragma unroll
      for (int q = 0; q < BATCH; q++) {
        const uint32_t c = cw[g + q] != DDF_SELL_PAD ? (cw[g + q] & 0x7fffffffu) : 0u;
        xv[q] = ld_keep_f64(x + c, pol_keep);
        if (PRE) xv[q] = __dmul_rn(ld_keep_f64(pre + c, pol_keep), xv[q]);
      }
#pragma unroll
      for (int q = 0; q < BATCH; q++)
        if (cw[g + q] != DDF_SELL_PAD) a = (cw[g + q] >> 31) ? __dsub_rn(a, xv[q]) : __dadd_rn(a, xv[q]);
    }
  }
  if (STAGE && staged) stage_wait(&mbar, 0);
  if (i >= nrows) return;
  if (UP && STAGE) {
    const uint32_t u0 = up[i] - a0, u1 = up[i + 1] - a0;
    for (uint32_t q = u0; q < u1; q++) {
      double v = sx[q];
      if (PRE) v = __dmul_rn(sx[cap + q], v);
      a = up_negative ? __dsub_rn(a, v) : __dadd_rn(a, v);
    }
  } else if (UP) {
    const uint32_t u0 = up[i], u1 = up[i + 1];
    for (uint32_t q0 = u0; q0 < u1; q0 += BATCH) {
      double xv[BATCH];
#pragma unroll
      for (int q = 0; q < BATCH; q++) {
        const uint32_t c = q0 + q < u1 ? q0 + q : u0;
        xv[q] = ld_keep_f64(x + c, pol_keep);
        if (PRE) xv[q] = __dmul_rn(ld_keep_f64(pre + c, pol_keep), xv[q]);
      }
#pragma unroll
      for (int q = 0; q < BATCH; q++)
        if (q0 + q < u1) a = up_negative ? __dsub_rn(a, xv[q]) : __dadd_rn(a, xv[q]);
    }
  }
  if (POST) a = __dmul_rn(post[i], a);
  if (alpha != 1.0) a = __dmul_rn(alpha, a);
  st_hint_f64(y + i, a, pol_stream);
}

// y = scale * (d .* x): pure streaming (24 B per element).  Four elements per thread as two 128-bit accesses per
// array, read-once loads (L1 no-allocate), evict-first stores; the tail and unaligned views take the scalar path.
__global__ void __launch_bounds__(256)
k_apply_diag(const double* __restrict__ d, double sign_or_scale, const double* __restrict__ x, double* __restrict__ y,
             int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    const double2 x0 = ld_stream_f64x2(reinterpret_cast<const double2*>(x + i));
    const double2 x1 = ld_stream_f64x2(reinterpret_cast<const double2*>(x + i + 2));
    double v[4] = {x0.x, x0.y, x1.x, x1.y};
    if (d) {
      const double2 d0 = ld_stream_f64x2(reinterpret_cast<const double2*>(d + i));
      const double2 d1 = ld_stream_f64x2(reinterpret_cast<const double2*>(d + i + 2));
      v[0] = __dmul_rn(d0.x, v[0]); v[1] = __dmul_rn(d0.y, v[1]);
      v[2] = __dmul_rn(d1.x, v[2]); v[3] = __dmul_rn(d1.y, v[3]);
    }
    if (sign_or_scale != 1.0) {
#pragma unroll
      for (int q = 0; q < 4; q++) v[q] = __dmul_rn(sign_or_scale, v[q]);
    }
    st_stream_f64x2(reinterpret_cast<double2*>(y + i), v[0], v[1]);
    st_stream_f64x2(reinterpret_cast<double2*>(y + i + 2), v[2], v[3]);
  } else {
    for (int64_t k = i; k < n; k++) {
      double w = d ? __dmul_rn(d[k], x[k]) : x[k];
      if (sign_or_scale != 1.0) w = __dmul_rn(sign_or_scale, w);
      y[k] = w;
    }
  }
}

__global__ void k_apply_csrf64(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                               const double* __restrict__ val, int64_t nrows, const double* __restrict__ x,
                               double alpha, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const uint32_t lo = rowptr[i], hi = rowptr[i + 1];
  double a = 0.0;
  for (uint32_t p = lo; p < hi; p++) a = __dadd_rn(a, __dmul_rn(__ldg(val + p), __ldg(x + __ldg(col + p))));
  if (alpha != 1.0) a = __dmul_rn(alpha, a);
  y[i] = a;
}

// The same rows from SELL-32-256 storage (sell.cuh): one thread per row, a warp's index/value loads coalesced,
// 8 gathers in flight.  Entries of a row keep their ascending-column order, so the result is bit-identical to
// k_apply_csrf64.
__global__ void __launch_bounds__(DDF_SELL_WIN)
k_apply_sellf64(const uint32_t* __restrict__ sell_ptr, const uint8_t* __restrict__ perm8, const uint32_t* __restrict__ scol,
                const double* __restrict__ sval, int64_t nrows, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t r = (int64_t)blockIdx.x * DDF_SELL_WIN + threadIdx.x;        // sorted position
  const int64_t i = (int64_t)blockIdx.x * DDF_SELL_WIN + perm8[r];           // the row it holds
  const int64_t s = r >> 5;
  const uint32_t base = sell_ptr[s] + (uint32_t)(r & 31);
  const uint32_t w = (sell_ptr[s + 1] - sell_ptr[s]) >> 5;
  double acc = 0.0;
  for (uint32_t q0 = 0; q0 < w; q0 += 8) {
    uint32_t j[8];
    double l[8], xv[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const bool ok = q0 + q < w;                // warp-uniform
      const size_t pos = (size_t)base + (size_t)(q0 + q) * 32;
      j[q] = ok ? (uint32_t)ld_stream_int(reinterpret_cast<const int*>(scol) + pos) : DDF_SELL_PAD;
      l[q] = ok ? ld_stream_f64(sval + pos) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) xv[q] = __ldg(x + (j[q] != DDF_SELL_PAD ? j[q] : 0u));
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (j[q] != DDF_SELL_PAD) acc = __dadd_rn(acc, __dmul_rn(l[q], xv[q]));
  }
  if (i < nrows) y[i] = acc;
}

// ------------------------------------------------------------- tuning variants
// Experimental variants of the plain (no fused diagonal) fixed-width kernel, selected
// with ddf_set_tuning("fixed", v).  All produce bit-identical results.
//   1: L2 eviction-priority hints (gathers evict_last, streams evict_first)
//   2: 1 + index loads allocate in L1
//   3: 1 + warp-coalesced index loads transposed through shared memory
//   4: 1 + persistent blocks with the next tile's indices prefetched
// Storage / launch variants of the default paths (the value means one thing for y = d_k x and another for the
// boundary / dual-derivative apply; an operator only looks at its own):
//   y = d_k x:        18 explicit tables; 19 prefix-packed d0 with lane = consecutive rows; 20 prefix-packed instead
//                     of bit-packed rows; 21 / 22 bit-packed rows with more / fewer rows per lane
//   boundary apply:   9 row-CSR; 17 8 gathers per round on short rows; 20 plain SELL copy (no upper runs); 21 upper
//                     runs read from global memory; 22 upper runs for short rows too; 28 index words 4 at a time; 32 pair table with the 40-register build

template <int W, int XM = 0>
__device__ __forceinline__ void fixed_rows4(const int* idx, const double* __restrict__ x, uint64_t pol_keep,
                                            uint64_t pol_stream, int first_negative, double* __restrict__ y,
                                            int64_t r0) {
  double xv[4 * W];
#pragma unroll
  for (int q = 0; q < 4 * W; q++) {
    if (XM == 0) xv[q] = ld_keep_f64(x + idx[q], pol_keep);
    else if (XM == 1) xv[q] = __ldg(x + idx[q]);                                   // no eviction hint
    else asm("ld.global.nc.L2::256B.f64 %0, [%1];" : "=d"(xv[q]) : "l"(x + idx[q]));   // fetch whole lines into L2
  }
  double acc[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    double a = 0.0;
#pragma unroll
    for (int p = 0; p < W; p++) {
      const bool neg = ((p & 1) != 0) != (first_negative != 0);
      a = neg ? __dsub_rn(a, xv[r * W + p]) : __dadd_rn(a, xv[r * W + p]);
    }
    acc[r] = a;
  }
  st_hint_f64x2(reinterpret_cast<double2*>(y + r0), acc[0], acc[1], pol_stream);
  st_hint_f64x2(reinterpret_cast<double2*>(y + r0 + 2), acc[2], acc[3], pol_stream);
}

template <int W>
__device__ __forceinline__ void fixed_tail(const int32_t* __restrict__ tab, const double* __restrict__ x,
                                           int first_negative, double* __restrict__ y, int64_t r0, int64_t nrows) {
  for (int64_t r = r0; r < nrows; r++) {
    double a = 0.0;
#pragma unroll
    for (int p = 0; p < W; p++) {
      const double v = x[tab[r * W + p]];
      const bool neg = ((p & 1) != 0) != (first_negative != 0);
      a = neg ? __dsub_rn(a, v) : __dadd_rn(a, v);
    }
    y[r] = a;
  }
}

template <int W, int VAR>
__global__ void __launch_bounds__(256)
k_apply_fixed_x(const int32_t* __restrict__ tab, int64_t nrows, int first_negative, const double* __restrict__ x,
                double* __restrict__ y) {
  const uint64_t pol_keep = make_policy_evict_last();
  const uint64_t pol_stream = make_policy_evict_first();
  if (VAR == 1 || VAR == 2 || VAR == 5 || VAR == 6) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r0 = t * 4;
    if (r0 >= nrows) return;
    if (r0 + 3 < nrows) {
      int idx[4 * W];
      const int4* tp = reinterpret_cast<const int4*>(tab) + t * W;
#pragma unroll
      for (int q = 0; q < W; q++) {
        int4 v = VAR == 1 ? ld_stream_int4_hint(tp + q, pol_stream) : ld_l1_int4_hint(tp + q, pol_stream);
        idx[4 * q + 0] = v.x; idx[4 * q + 1] = v.y; idx[4 * q + 2] = v.z; idx[4 * q + 3] = v.w;
      }
      fixed_rows4<W, VAR == 5 ? 1 : (VAR == 6 ? 2 : 0)>(idx, x, pol_keep, pol_stream, first_negative, y, r0);
    } else {
      fixed_tail<W>(tab, x, first_negative, y, r0, nrows);
    }
  } else if (VAR == 3) {
    // a warp owns 128 consecutive rows = 32*W int4 of indices: W fully coalesced 512-byte loads,
    // transposed through shared memory so that lane l ends up with rows 4l..4l+3
    __shared__ int4 sm[8][32 * W];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp_row0 = ((int64_t)blockIdx.x * 8 + w) * 128;
    if (warp_row0 >= nrows) return;
    const int64_t r0 = warp_row0 + lane * 4;
    if (warp_row0 + 127 < nrows) {
      const int4* tp = reinterpret_cast<const int4*>(tab) + (warp_row0 / 4) * W;
#pragma unroll
      for (int q = 0; q < W; q++) sm[w][q * 32 + lane] = ld_stream_int4_hint(tp + q * 32 + lane, pol_stream);
      __syncwarp();
      int idx[4 * W];
#pragma unroll
      for (int q = 0; q < W; q++) {
        int4 v = sm[w][lane * W + q];
        idx[4 * q + 0] = v.x; idx[4 * q + 1] = v.y; idx[4 * q + 2] = v.z; idx[4 * q + 3] = v.w;
      }
      fixed_rows4<W>(idx, x, pol_keep, pol_stream, first_negative, y, r0);
    } else if (r0 < nrows) {
      fixed_tail<W>(tab, x, first_negative, y, r0, r0 + 4 < nrows ? r0 + 4 : nrows);
    }
  } else {
    // VAR 4: persistent blocks; the indices of the next tile are in flight while this tile gathers
    const int64_t ntiles = (nrows / 4);                 // full 4-row groups
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int4 nxt[W];
    if (t < ntiles) {
      const int4* tp = reinterpret_cast<const int4*>(tab) + t * W;
#pragma unroll
      for (int q = 0; q < W; q++) nxt[q] = ld_stream_int4_hint(tp + q, pol_stream);
    }
    for (; t < ntiles; t += stride) {
      int idx[4 * W];
#pragma unroll
      for (int q = 0; q < W; q++) {
        idx[4 * q + 0] = nxt[q].x; idx[4 * q + 1] = nxt[q].y; idx[4 * q + 2] = nxt[q].z; idx[4 * q + 3] = nxt[q].w;
      }
      if (t + stride < ntiles) {
        const int4* tp = reinterpret_cast<const int4*>(tab) + (t + stride) * W;
#pragma unroll
        for (int q = 0; q < W; q++) nxt[q] = ld_stream_int4_hint(tp + q, pol_stream);
      }
      fixed_rows4<W>(idx, x, pol_keep, pol_stream, first_negative, y, t * 4);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) fixed_tail<W>(tab, x, first_negative, y, ntiles * 4, nrows);
  }
}

// VAR 7: lane = consecutive rows.  A warp owns 32*RPW consecutive rows; in every gather instruction the 32 lanes
// read the x entries of 32 CONSECUTIVE rows (lexicographically adjacent simplices share faces, so the lanes hit
// few distinct sectors -> fewer L1 wavefronts per instruction than the rows-4t..4t+3-per-thread mapping).
template <int W, int RPW, bool PRE = false, bool POST = false>
__global__ void __launch_bounds__(256)
k_apply_fixed_rows(const int32_t* __restrict__ tab, int64_t nrows, int first_negative, const double* __restrict__ x,
                   double* __restrict__ y, const double* __restrict__ pre = nullptr,
                   const double* __restrict__ post = nullptr, double alpha = 1.0) {
  const uint64_t pol_keep = make_policy_evict_last();
  const uint64_t pol_stream = make_policy_evict_first();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t r0 = warp * (32 * RPW) + lane;
  if (warp * (32 * RPW) >= nrows) return;
  int idx[RPW][W];
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const int64_t r = r0 + 32 * k;
    if (W == 2) {                     // one 8-byte load per row, 256 B per warp
      int2 v = make_int2(0, 0);
      if (r < nrows) v = ld_stream_int2(reinterpret_cast<const int2*>(tab) + r);
      idx[k][0] = v.x; idx[k][W - 1] = v.y;
    } else if (W == 4) {              // one 16-byte load per row, 512 B per warp
      int4 v = make_int4(0, 0, 0, 0);
      if (r < nrows) v = ld_stream_int4_hint(reinterpret_cast<const int4*>(tab) + r, pol_stream);
      idx[k][0] = v.x; idx[k][1] = v.y; idx[k][W - 2] = v.z; idx[k][W - 1] = v.w;
    } else {                          // 12-byte rows: three L1-allocating loads that share their lines
#pragma unroll
      for (int p = 0; p < W; p++) idx[k][p] = r < nrows ? ld_keep_s32(tab + r * W + p, pol_stream) : 0;
    }
  }
  double xv[RPW][W];
#pragma unroll
  for (int k = 0; k < RPW; k++)
#pragma unroll
    for (int p = 0; p < W; p++) xv[k][p] = ld_keep_f64(x + idx[k][p], pol_keep);
  if (PRE) {                          // fused diagonal on the input side: fl(pre[c] * x[c]) like the unfused sequence
#pragma unroll
    for (int k = 0; k < RPW; k++)
#pragma unroll
      for (int p = 0; p < W; p++) xv[k][p] = __dmul_rn(ld_keep_f64(pre + idx[k][p], pol_keep), xv[k][p]);
  }
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const int64_t r = r0 + 32 * k;
    double a = 0.0;
#pragma unroll
    for (int p = 0; p < W; p++) {
      const bool neg = ((p & 1) != 0) != (first_negative != 0);
      a = neg ? __dsub_rn(a, xv[k][p]) : __dadd_rn(a, xv[k][p]);
    }
    if (r < nrows) {
      if (POST) a = __dmul_rn(ld_stream_f64(post + r), a);
      if (alpha != 1.0) a = __dmul_rn(alpha, a);
      st_hint_f64(y + r, a, pol_stream);
    }
  }
}

// ---------------------------------------------------------------- prefix-packed rows of d_k
// Row tau = (v_0 .. v_k+1) of d_k holds the ids of its faces in ascending order, and the FIRST of them -- the face
// that omits the last vertex -- is tau's prefix (v_0 .. v_k).  Ids are lexicographic ranks (Manifolds.jl:735-757), so
// down the rows that first column is non-decreasing, and the second one (the face (v_0 .. v_k-1, v_k+1), which
// shares the prefix's first k vertices) lies a short way after it.  The packed table stores, per row, W-1 words
// instead of W: word 0 = (idx0 - base) << 24 | (idx1 - idx0) with one `base` per group of 32 rows, the remaining
// indices as they are: 4 / 8 / 12 instead of 8 / 12 / 16 bytes of index data per row of d0 / d1 / d2.  The gathers,
// their order and the arithmetic are those of k_apply_fixed_rows, so results are bit-identical.  A mesh whose rows
// do not pack (a delta beyond 8 bits or an offset beyond 24) keeps the explicit table.
template <int W>
__global__ void __launch_bounds__(256)
k_pack_fixed(const int32_t* __restrict__ tab, int64_t nrows, uint32_t max_offset, uint32_t* __restrict__ words,
             uint32_t* __restrict__ base, uint32_t* __restrict__ bad) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const uint32_t b = (uint32_t)tab[(r & ~(int64_t)31) * W];
  const uint32_t i0 = (uint32_t)tab[r * W], i1 = (uint32_t)tab[r * W + 1];
  if ((r & 31) == 0) base[r >> 5] = b;
  const uint32_t d = i0 - b, o = i1 - i0;
  if (i0 < b || d > 255u || i1 <= i0 || o > max_offset) { atomicOr(bad, 1u); return; }
  words[r * (W - 1)] = (d << 24) | o;
#pragma unroll
  for (int p = 2; p < W; p++) words[r * (W - 1) + p - 1] = (uint32_t)tab[r * W + p];
}

// lane = consecutive rows, RPW groups of 32 rows per warp (the mapping of k_apply_fixed_rows)
template <int W, int RPW, bool PRE, bool POST>
__global__ void __launch_bounds__(256)
k_apply_packed_rows(const uint32_t* __restrict__ words, const uint32_t* __restrict__ base, int64_t nrows,
                    int first_negative, const double* __restrict__ x, double* __restrict__ y,
                    const double* __restrict__ pre, const double* __restrict__ post, double alpha) {
  const uint64_t pol_keep = make_policy_evict_last();
  const uint64_t pol_stream = make_policy_evict_first();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t r0 = warp * (32 * RPW) + lane;
  if (warp * (32 * RPW) >= nrows) return;
  int idx[RPW][W];
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const int64_t r = r0 + 32 * k;
    const bool ok = r < nrows;
    const uint32_t b = ok ? __ldg(base + warp * RPW + k) : 0u;
    uint32_t w0 = 0;
    if constexpr (W == 2) {
      if (ok) w0 = (uint32_t)ld_stream_int(reinterpret_cast<const int*>(words) + r);
    } else if constexpr (W == 3) {    // one 8-byte load per row, 256 B per warp
      int2 v = make_int2(0, 0);
      if (ok) v = ld_stream_int2(reinterpret_cast<const int2*>(words) + r);
      w0 = (uint32_t)v.x; idx[k][W - 1] = v.y;
    } else {                          // 12-byte rows: three L1-allocating loads that share their lines
      int v[W - 1];
#pragma unroll
      for (int p = 0; p < W - 1; p++) v[p] = ok ? ld_keep_s32(reinterpret_cast<const int*>(words) + r * (W - 1) + p, pol_stream) : 0;
      w0 = (uint32_t)v[0];
#pragma unroll
      for (int p = 2; p < W; p++) idx[k][p] = v[p - 1];
    }
    idx[k][0] = (int)(b + (w0 >> 24));
    idx[k][1] = idx[k][0] + (int)(w0 & 0xffffffu);
  }
  double xv[RPW][W];
#pragma unroll
  for (int k = 0; k < RPW; k++)
#pragma unroll
    for (int p = 0; p < W; p++) xv[k][p] = ld_keep_f64(x + idx[k][p], pol_keep);
  if (PRE) {
#pragma unroll
    for (int k = 0; k < RPW; k++)
#pragma unroll
      for (int p = 0; p < W; p++) xv[k][p] = __dmul_rn(ld_keep_f64(pre + idx[k][p], pol_keep), xv[k][p]);
  }
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const int64_t r = r0 + 32 * k;
    double a = 0.0;
#pragma unroll
    for (int p = 0; p < W; p++) {
      const bool neg = ((p & 1) != 0) != (first_negative != 0);
      a = neg ? __dsub_rn(a, xv[k][p]) : __dadd_rn(a, xv[k][p]);
    }
    if (r < nrows) {
      if (POST) a = __dmul_rn(ld_stream_f64(post + r), a);
      if (alpha != 1.0) a = __dmul_rn(alpha, a);
      st_hint_f64(y + r, a, pol_stream);
    }
  }
}

// d0 (W = 2): four consecutive rows per thread = ONE 16-byte load of their four words, 16-byte stores
template <bool PRE, bool POST>
__global__ void __launch_bounds__(256)
k_apply_packed_quad(const uint32_t* __restrict__ words, const uint32_t* __restrict__ base, int64_t nrows,
                    int first_negative, const double* __restrict__ x, double* __restrict__ y,
                    const double* __restrict__ pre, const double* __restrict__ post, double alpha, int pf_blocks) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = t * 4;
  if (pf_blocks && threadIdx.x == 0) {       // words of the block that runs pf_blocks later: 1024 rows, contiguous
    const int64_t rp = ((int64_t)blockIdx.x + pf_blocks) * 1024;
    if (rp + 1024 <= nrows) bulk_prefetch_l2(words + rp, 4096);
  }
  if (r0 >= nrows) return;
  const uint64_t pol_keep = make_policy_evict_last();
  const uint64_t pol_stream = make_policy_evict_first();
  const uint32_t b = __ldg(base + (r0 >> 5));
  uint32_t w[4] = {0, 0, 0, 0};
  if (r0 + 3 < nrows) {
    const int4 v = ld_stream_int4_hint(reinterpret_cast<const int4*>(words) + t, pol_stream);
    w[0] = (uint32_t)v.x; w[1] = (uint32_t)v.y; w[2] = (uint32_t)v.z; w[3] = (uint32_t)v.w;
  } else {
    for (int q = 0; q < 4 && r0 + q < nrows; q++) w[q] = words[r0 + q];
  }
  int idx[8];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    idx[2 * q] = (int)(b + (w[q] >> 24));
    idx[2 * q + 1] = idx[2 * q] + (int)(w[q] & 0xffffffu);
  }
  double xv[8];
#pragma unroll
  for (int q = 0; q < 8; q++) xv[q] = ld_keep_f64(x + idx[q], pol_keep);
  if (PRE) {
#pragma unroll
    for (int q = 0; q < 8; q++) xv[q] = __dmul_rn(ld_keep_f64(pre + idx[q], pol_keep), xv[q]);
  }
  double acc[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    double a = 0.0;
    a = first_negative ? __dsub_rn(a, xv[2 * q]) : __dadd_rn(a, xv[2 * q]);
    a = first_negative ? __dadd_rn(a, xv[2 * q + 1]) : __dsub_rn(a, xv[2 * q + 1]);
    acc[q] = a;
  }
  if (r0 + 3 < nrows) {
    if (POST) {
      const double2 p0 = ld_stream_f64x2_hint(reinterpret_cast<const double2*>(post + r0), pol_stream);
      const double2 p1 = ld_stream_f64x2_hint(reinterpret_cast<const double2*>(post + r0 + 2), pol_stream);
      acc[0] = __dmul_rn(p0.x, acc[0]); acc[1] = __dmul_rn(p0.y, acc[1]);
      acc[2] = __dmul_rn(p1.x, acc[2]); acc[3] = __dmul_rn(p1.y, acc[3]);
    }
    if (alpha != 1.0) {
#pragma unroll
      for (int q = 0; q < 4; q++) acc[q] = __dmul_rn(alpha, acc[q]);
    }
    st_hint_f64x2(reinterpret_cast<double2*>(y + r0), acc[0], acc[1], pol_stream);
    st_hint_f64x2(reinterpret_cast<double2*>(y + r0 + 2), acc[2], acc[3], pol_stream);
  } else {
    for (int q = 0; q < 4 && r0 + q < nrows; q++) {
      double a = acc[q];
      if (POST) a = __dmul_rn(post[r0 + q], a);
      if (alpha != 1.0) a = __dmul_rn(alpha, a);
      y[r0 + q] = a;
    }
  }
}

static int packed_build(ddf_ctx* ctx, const int32_t* tab, int W, int64_t nrows, PackedBufs& pb) {
  const DdfTune& tn = ctx->tune;
  pb.release();
  pb.state = 2;
  if (W < 2 || W > 4 || nrows == 0) return 0;
  DevBuf<uint32_t> bad;
  DDF_CHECK(bad.alloc(ctx, 1));
  DDF_CUDA(cudaMemsetAsync(bad.p, 0, 4, ctx->stream));
  DDF_CHECK(pb.words.alloc(ctx, (size_t)nrows * (W - 1)));
  DDF_CHECK(pb.base.alloc(ctx, (size_t)ceil_div64(nrows, 32)));
  int grid;
  DDF_CHECK(grid_for(nrows, 256, &grid));
  const uint32_t max_off = (1u << tn.packbits) - 1u;      // 24 bits; the tuning knob only narrows it (tests)
  if (W == 2) k_pack_fixed<2><<<grid, 256, 0, ctx->stream>>>(tab, nrows, max_off, pb.words.p, pb.base.p, bad.p);
  else if (W == 3) k_pack_fixed<3><<<grid, 256, 0, ctx->stream>>>(tab, nrows, max_off, pb.words.p, pb.base.p, bad.p);
  else k_pack_fixed<4><<<grid, 256, 0, ctx->stream>>>(tab, nrows, max_off, pb.words.p, pb.base.p, bad.p);
  DDF_LAUNCH_CHECK(ctx);
  uint32_t hb = 0;
  DDF_CUDA(cudaMemcpyAsync(&hb, bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hb) { pb.words.release(); pb.base.release(); return 0; }
  pb.state = 1;
  return 0;
}

static int launch_packed(ddf_ctx* ctx, const PackedBufs& pb, int W, int64_t nrows, int first_negative, const double* x,
                         const double* pre, const double* post, double alpha, double* y, int64_t row0 = 0) {
  const DdfTune& tn = ctx->tune;
  const uint32_t *pw = pb.words.p + row0 * (W - 1), *pbase = pb.base.p + row0 / 32;     // row range: see launch_bits_w
  y += row0;
  if (post) post += row0;
  const int g4 = (int)ceil_div64(ceil_div64(nrows, 4), 256), g3 = (int)ceil_div64(ceil_div64(nrows, 3), 256);
  const bool quad = W == 2 && tn.fixed != 19;
#define DDF_PACKED_CASE(PRE_, POST_)                                                                                 \
  do {                                                                                                               \
    if (quad) k_apply_packed_quad<PRE_, POST_><<<g4, 256, 0, ctx->stream>>>(pw, pbase, nrows, first_negative, x, y, pre, post, alpha, tn.prefetch); \
    else if (W == 2) k_apply_packed_rows<2, 4, PRE_, POST_><<<g4, 256, 0, ctx->stream>>>(pw, pbase, nrows, first_negative, x, y, pre, post, alpha); \
    else if (W == 3) k_apply_packed_rows<3, 4, PRE_, POST_><<<g4, 256, 0, ctx->stream>>>(pw, pbase, nrows, first_negative, x, y, pre, post, alpha); \
    else k_apply_packed_rows<4, 3, PRE_, POST_><<<g3, 256, 0, ctx->stream>>>(pw, pbase, nrows, first_negative, x, y, pre, post, alpha);             \
  } while (0)
  if (pre && post) DDF_PACKED_CASE(true, true);
  else if (pre) DDF_PACKED_CASE(true, false);
  else if (post) DDF_PACKED_CASE(false, true);
  else DDF_PACKED_CASE(false, false);
#undef DDF_PACKED_CASE
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

// ---------------------------------------------------------------- bit-packed rows of d_k
// The face ids of a row ascend, the later ones lie a bounded way after the earlier ones (faces of one simplex share
// vertices, ids are lexicographic ranks, Manifolds.jl:735-757), and rows that are stored next to each other start
// near each other.  So a row is stored as ONE word of differences -- idx0 against the smallest idx0 of its 32-row
// group, idx_p against idx_{p-1} -- with field widths chosen per table from the largest differences it holds
// (k_bits_scan): 4 bytes per row when they sum to <= 32 bits, 8 bytes when <= 64, instead of 4 (W-1) / 4 W.
// 3D Kuhn n=256: d1 rows (6+4+19 bits) take 4 bytes instead of 8, d2 rows at the top level (23+2+5+20 bits, rows in
// the generator's order) 8 instead of 16.  Same gathers in the same order as k_apply_fixed_rows => identical bits.
struct BitFmt {
  int s1, s2, s3;                 // shifts of the fields idx1 - idx0, idx2 - idx1, idx3 - idx2
  uint32_t m0, m1, m2, m3;        // field masks
};

template <int W>
__global__ void __launch_bounds__(256)
k_bits_scan(const int32_t* __restrict__ tab, int64_t nrows, uint32_t* __restrict__ base, uint32_t* __restrict__ maxd,
            uint32_t* __restrict__ bad) {
  // Persistent grid: a warp walks aligned groups of 32 rows with a grid stride and keeps the running maxima of the
  // W difference fields in registers; one block-level reduction and W atomics per BLOCK at the end.  (One atomic --
  // or even one filtered read -- per warp on the same four words is an L2 hot spot: 13 ms, then 5 ms at C4 for a
  // 2.4 GB table that streams in under a millisecond.)
  __shared__ uint32_t smax[8][W];
  __shared__ uint32_t sbad;
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) sbad = 0u;
  __syncthreads();
  uint32_t mx[W];
#pragma unroll
  for (int p = 0; p < W; p++) mx[p] = 0u;
  bool unsorted = false;
  const int64_t ngroups = (nrows + 31) >> 5;
  for (int64_t grp = (int64_t)blockIdx.x * 8 + w; grp < ngroups; grp += (int64_t)gridDim.x * 8) {
    const int64_t r = grp * 32 + lane;
    const bool ok = r < nrows;
    uint32_t idx[W];
#pragma unroll
    for (int p = 0; p < W; p++) idx[p] = ok ? (uint32_t)tab[r * W + p] : 0xffffffffu;
    const uint32_t b = __reduce_min_sync(0xffffffffu, idx[0]);
    if (lane == 0) base[grp] = b;
    if (ok) {
      mx[0] = max(mx[0], idx[0] - b);
#pragma unroll
      for (int p = 1; p < W; p++) {
        mx[p] = max(mx[p], idx[p] - idx[p - 1]);
        unsorted |= idx[p] <= idx[p - 1];
      }
    }
  }
#pragma unroll
  for (int p = 0; p < W; p++) {
    const uint32_t m = __reduce_max_sync(0xffffffffu, mx[p]);
    if (lane == 0) smax[w][p] = m;
  }
  if (unsorted) atomicOr(&sbad, 1u);
  __syncthreads();
  if (threadIdx.x < W) {
    uint32_t m = 0u;
    for (int k = 0; k < 8; k++) m = max(m, smax[k][threadIdx.x]);
    atomicMax(maxd + threadIdx.x, m);
  }
  if (threadIdx.x == 0 && sbad) atomicOr(bad, 1u);
}

template <int W, int WB>
__global__ void __launch_bounds__(256)
k_bits_pack(const int32_t* __restrict__ tab, int64_t nrows, const uint32_t* __restrict__ base, BitFmt f,
            uint32_t* __restrict__ words) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  uint32_t idx[W];
#pragma unroll
  for (int p = 0; p < W; p++) idx[p] = (uint32_t)tab[r * W + p];
  uint64_t w = (uint64_t)(idx[0] - base[r >> 5]);
  w |= (uint64_t)(idx[1] - idx[0]) << f.s1;
  if constexpr (W > 2) w |= (uint64_t)(idx[2] - idx[1]) << f.s2;
  if constexpr (W > 3) w |= (uint64_t)(idx[3] - idx[2]) << f.s3;
  if (WB == 4) words[r] = (uint32_t)w;
  else reinterpret_cast<uint2*>(words)[r] = make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}

// lane = consecutive rows, RPW groups of 32 rows per warp (the mapping of k_apply_fixed_rows)
template <int W, int WB, int RPW, bool PRE, bool POST>
__global__ void __launch_bounds__(256)
k_apply_bits_rows(const uint32_t* __restrict__ words, const uint32_t* __restrict__ base, BitFmt f, int64_t nrows,
                  int first_negative, const double* __restrict__ x, double* __restrict__ y,
                  const double* __restrict__ pre, const double* __restrict__ post, double alpha, int pf_blocks) {
  const uint64_t pol_keep = make_policy_evict_last();
  const uint64_t pol_stream = make_policy_evict_first();
  if (pf_blocks && threadIdx.x == 0) {       // words of the block that runs pf_blocks later: 256 * RPW rows, contiguous
    const int64_t rp = ((int64_t)blockIdx.x + pf_blocks) * (256 * RPW);
    if (rp + 256 * RPW <= nrows) bulk_prefetch_l2(reinterpret_cast<const char*>(words) + rp * WB, 256 * RPW * WB);
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t r0 = warp * (32 * RPW) + lane;
  if (warp * (32 * RPW) >= nrows) return;
  uint32_t idx[RPW][W];
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const int64_t r = r0 + 32 * k;
    const bool ok = r < nrows;
    const uint32_t b = ok ? __ldg(base + warp * RPW + k) : 0u;
    if constexpr (WB == 4) {          // one 4-byte word per row, 128 B per warp
      const uint32_t w = ok ? (uint32_t)ld_stream_int(reinterpret_cast<const int*>(words) + r) : 0u;
      idx[k][0] = b + (w & f.m0);
      idx[k][1] = idx[k][0] + ((w >> f.s1) & f.m1);
      if constexpr (W > 2) idx[k][2] = idx[k][1] + ((w >> f.s2) & f.m2);
      if constexpr (W > 3) idx[k][3] = idx[k][2] + ((w >> f.s3) & f.m3);
    } else {                          // one 8-byte word per row, 256 B per warp
      int2 v = make_int2(0, 0);
      if (ok) v = ld_stream_int2(reinterpret_cast<const int2*>(words) + r);
      const uint64_t w = (uint64_t)(uint32_t)v.x | ((uint64_t)(uint32_t)v.y << 32);
      idx[k][0] = b + ((uint32_t)w & f.m0);
      idx[k][1] = idx[k][0] + ((uint32_t)(w >> f.s1) & f.m1);
      if constexpr (W > 2) idx[k][2] = idx[k][1] + ((uint32_t)(w >> f.s2) & f.m2);
      if constexpr (W > 3) idx[k][3] = idx[k][2] + ((uint32_t)(w >> f.s3) & f.m3);
    }
  }
  double xv[RPW][W];
#pragma unroll
  for (int k = 0; k < RPW; k++)
#pragma unroll
    for (int p = 0; p < W; p++) xv[k][p] = ld_keep_f64(x + idx[k][p], pol_keep);
  if (PRE) {
#pragma unroll
    for (int k = 0; k < RPW; k++)
#pragma unroll
      for (int p = 0; p < W; p++) xv[k][p] = __dmul_rn(ld_keep_f64(pre + idx[k][p], pol_keep), xv[k][p]);
  }
#pragma unroll
  for (int k = 0; k < RPW; k++) {
    const int64_t r = r0 + 32 * k;
    double a = 0.0;
#pragma unroll
    for (int p = 0; p < W; p++) {
      const bool neg = ((p & 1) != 0) != (first_negative != 0);
      a = neg ? __dsub_rn(a, xv[k][p]) : __dadd_rn(a, xv[k][p]);
    }
    if (r < nrows) {
      if (POST) a = __dmul_rn(ld_stream_f64(post + r), a);
      if (alpha != 1.0) a = __dmul_rn(alpha, a);
      st_hint_f64(y + r, a, pol_stream);
    }
  }
}

static BitFmt bits_fmt(const BitRows& br) {
  BitFmt f;
  f.s1 = br.b[0]; f.s2 = br.b[0] + br.b[1]; f.s3 = br.b[0] + br.b[1] + br.b[2];
  auto mask = [](int bits) { return bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u); };
  f.m0 = mask(br.b[0]); f.m1 = mask(br.b[1]); f.m2 = mask(br.b[2]); f.m3 = mask(br.b[3]);
  return f;
}

static int bits_build(ddf_ctx* ctx, const int32_t* tab, int W, int64_t nrows, BitRows& br) {
  br.release();
  br.state = 2;
  if (W < 3 || W > 4 || nrows == 0) return 0;
  DevBuf<uint32_t> stat;                       // maxd[0..3], bad
  DDF_CHECK(stat.alloc(ctx, 5));
  DDF_CUDA(cudaMemsetAsync(stat.p, 0, 5 * 4, ctx->stream));
  DDF_CHECK(br.base.alloc(ctx, (size_t)ceil_div64(nrows, 32)));
  int grid;
  DDF_CHECK(grid_for(nrows, 256, &grid));
  const int sgrid = std::min(grid, ctx->sm_count * 8);          // persistent: 8 resident blocks per SM
  if (W == 3) k_bits_scan<3><<<sgrid, 256, 0, ctx->stream>>>(tab, nrows, br.base.p, stat.p, stat.p + 4);
  else k_bits_scan<4><<<sgrid, 256, 0, ctx->stream>>>(tab, nrows, br.base.p, stat.p, stat.p + 4);
  DDF_LAUNCH_CHECK(ctx);
  uint32_t h[5] = {0, 0, 0, 0, 0};
  DDF_CUDA(cudaMemcpyAsync(h, stat.p, 5 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  int total = 0;
  for (int p = 0; p < 4; p++) {
    int bits = 1;
    while (bits < 32 && (h[p] >> bits) != 0) bits++;
    br.b[p] = p < W ? bits : 0;
    total += br.b[p];
  }
  if (h[4] || total > 64) { br.base.release(); return 0; }
  br.wbytes = total <= 32 ? 4 : 8;
  DDF_CHECK(br.words.alloc(ctx, (size_t)nrows * (br.wbytes / 4)));
  const BitFmt f = bits_fmt(br);
  if (W == 3 && br.wbytes == 4) k_bits_pack<3, 4><<<grid, 256, 0, ctx->stream>>>(tab, nrows, br.base.p, f, br.words.p);
  else if (W == 3) k_bits_pack<3, 8><<<grid, 256, 0, ctx->stream>>>(tab, nrows, br.base.p, f, br.words.p);
  else if (br.wbytes == 4) k_bits_pack<4, 4><<<grid, 256, 0, ctx->stream>>>(tab, nrows, br.base.p, f, br.words.p);
  else k_bits_pack<4, 8><<<grid, 256, 0, ctx->stream>>>(tab, nrows, br.base.p, f, br.words.p);
  DDF_LAUNCH_CHECK(ctx);
  br.state = 1;
  return 0;
}

template <int W, int WB>
static void launch_bits_w(ddf_ctx* ctx, const BitRows& br, const BitFmt& f, int64_t nrows, int first_negative,
                          const double* x, const double* pre, const double* post, double alpha, double* y,
                          int64_t row0 = 0) {
  const DdfTune& tn = ctx->tune;
  constexpr int RPW = W == 3 ? 4 : 3;
  // a row range [row0, row0 + nrows): row0 is a multiple of 32, so the kernel sees shifted tables (the base word of
  // a 32-row group, the per-row words, y and the per-row output diagonal all move by row0; x and the input diagonal
  // are indexed by column and stay)
  const int g = (int)ceil_div64(ceil_div64(nrows, RPW), 256);
  const uint32_t *w = br.words.p + row0 * (WB / 4), *b = br.base.p + row0 / 32;
  y += row0;
  if (post) post += row0;
  if (!pre && !post && (tn.fixed == 21 || tn.fixed == 22)) {      // A/B: other rows-per-lane counts
    constexpr int RA = W == 3 ? 8 : 4, RB = 2;
    if (tn.fixed == 21)
      k_apply_bits_rows<W, WB, RA, false, false><<<(int)ceil_div64(ceil_div64(nrows, RA), 256), 256, 0, ctx->stream>>>(w, b, f, nrows, first_negative, x, y, pre, post, alpha, tn.prefetch);
    else
      k_apply_bits_rows<W, WB, RB, false, false><<<(int)ceil_div64(ceil_div64(nrows, RB), 256), 256, 0, ctx->stream>>>(w, b, f, nrows, first_negative, x, y, pre, post, alpha, tn.prefetch);
    return;
  }
  if (pre && post) k_apply_bits_rows<W, WB, RPW, true, true><<<g, 256, 0, ctx->stream>>>(w, b, f, nrows, first_negative, x, y, pre, post, alpha, tn.prefetch);
  else if (pre) k_apply_bits_rows<W, WB, RPW, true, false><<<g, 256, 0, ctx->stream>>>(w, b, f, nrows, first_negative, x, y, pre, post, alpha, tn.prefetch);
  else if (post) k_apply_bits_rows<W, WB, RPW, false, true><<<g, 256, 0, ctx->stream>>>(w, b, f, nrows, first_negative, x, y, pre, post, alpha, tn.prefetch);
  else k_apply_bits_rows<W, WB, RPW, false, false><<<g, 256, 0, ctx->stream>>>(w, b, f, nrows, first_negative, x, y, pre, post, alpha, tn.prefetch);
}

static int launch_bits(ddf_ctx* ctx, const BitRows& br, int W, int64_t nrows, int first_negative, const double* x,
                       const double* pre, const double* post, double alpha, double* y, int64_t row0 = 0) {
  const BitFmt f = bits_fmt(br);
  if (W == 3 && br.wbytes == 4) launch_bits_w<3, 4>(ctx, br, f, nrows, first_negative, x, pre, post, alpha, y, row0);
  else if (W == 3) launch_bits_w<3, 8>(ctx, br, f, nrows, first_negative, x, pre, post, alpha, y, row0);
  else if (br.wbytes == 4) launch_bits_w<4, 4>(ctx, br, f, nrows, first_negative, x, pre, post, alpha, y, row0);
  else launch_bits_w<4, 8>(ctx, br, f, nrows, first_negative, x, pre, post, alpha, y, row0);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

// Which storage the rows of d_{R-1} (the table of level R) use on this mesh, built on first use:
// 0 = explicit table, 1 = prefix-packed, 2 = bit-packed 4-byte rows, 3 = bit-packed 8-byte rows.
// Tuning: fixed 18 = explicit only, 20 = no bit-packed rows; a narrowed "packbits" (tests) also switches them off.
static int dk_rows_format(ddf_mesh* m, int R, int* fmt) {
  const DdfTune& tn = m->ctx->tune;
  const int W = R + 1;
  *fmt = 0;
  if (W < 2 || W > 4 || m->n[R] == 0) return 0;
  if (!(tn.fixed == 0 || (tn.fixed >= 19 && tn.fixed <= 22))) return 0;
  const bool allow_bits = W >= 3 && tn.fixed != 20 && tn.packbits == 24;
  BitRows& br = m->bbits[R];
  PackedBufs& pb = m->bpack[R];
  if (allow_bits) {
    if (br.state == 0) DDF_CHECK(bits_build(m->ctx, m->bnd_table(R), W, m->n[R], br));
    if (br.state == 1 && (br.wbytes == 4 || W == 4)) { *fmt = br.wbytes == 4 ? 2 : 3; return 0; }
  }
  if (pb.state == 0) DDF_CHECK(packed_build(m->ctx, m->bnd_table(R), W, m->n[R], pb));
  if (pb.state == 1) { *fmt = 1; return 0; }
  if (allow_bits && br.state == 1) *fmt = 3;      // W = 3 rows that do not prefix-pack (top level): 8 instead of 12 bytes
  return 0;
}

template <int W>
static void launch_fixed_x(ddf_ctx* ctx, int var, const int32_t* tab, int64_t nrows, int first_negative,
                           const double* x, double* y) {
  const int grid4 = (int)ceil_div64(ceil_div64(nrows, 4), 256);
  switch (var) {
    case 1: k_apply_fixed_x<W, 1><<<grid4, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    case 2: k_apply_fixed_x<W, 2><<<grid4, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    case 3: k_apply_fixed_x<W, 3><<<(int)ceil_div64(nrows, 1024), 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    case 7: k_apply_fixed_rows<W, 4><<<grid4, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    case 8: k_apply_fixed_rows<W, 8><<<(int)ceil_div64(ceil_div64(nrows, 8), 256), 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    case 10: k_apply_fixed_rows<W, 2><<<(int)ceil_div64(ceil_div64(nrows, 2), 256), 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    case 11: k_apply_fixed_rows<W, 6><<<(int)ceil_div64(ceil_div64(nrows, 6), 256), 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    case 12: k_apply_fixed_rows<W, 3><<<(int)ceil_div64(ceil_div64(nrows, 3), 256), 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    case 5: k_apply_fixed_x<W, 5><<<grid4, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    case 6: k_apply_fixed_x<W, 6><<<grid4, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y); break;
    default: {
      int g = ctx->sm_count * 8;
      if (g > grid4) g = grid4;
      k_apply_fixed_x<W, 4><<<g, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y);
    }
  }
}

// ================================================================ kernel launchers
static int launch_fixed(ddf_ctx* ctx, const int32_t* tab, int W, int64_t nrows, int first_negative, const double* x,
                        const double* pre, const double* post, double alpha, double* y, int64_t row0 = 0) {
  const DdfTune& tn = ctx->tune;
  if (nrows == 0) return 0;
  tab += row0 * W;                         // row range: see launch_bits_w
  y += row0;
  if (post) post += row0;
  int grid;
  DDF_CHECK(grid_for(ceil_div64(nrows, 4), 256, &grid));
  if (tn.fixed > 0 && tn.fixed < 16 && tn.fixed != 9 && !pre && !post && alpha == 1.0 && W >= 2 && W <= 4) {
    if (W == 2) launch_fixed_x<2>(ctx, tn.fixed, tab, nrows, first_negative, x, y);
    else if (W == 3) launch_fixed_x<3>(ctx, tn.fixed, tab, nrows, first_negative, x, y);
    else launch_fixed_x<4>(ctx, tn.fixed, tab, nrows, first_negative, x, y);
    DDF_LAUNCH_CHECK(ctx);
    return 0;
  }
  // default for the plain d_1 / d_2 apply: lane = consecutive rows (k_apply_fixed_rows).  Measured at C4
  // (profiles/r01_tuning.md): d1 0.786 -> 0.733 ms, d2 0.769 -> 0.564 ms; d0 (W = 2) is faster with 4 rows per thread.
  if ((tn.fixed == 0 || tn.fixed == 18) && (W == 3 || W == 4)) {
    const int g3 = (int)ceil_div64(ceil_div64(nrows, 3), 256);
#define DDF_ROWS_CASE(PRE_, POST_)                                                                                   \
  do {                                                                                                               \
    if (W == 3) k_apply_fixed_rows<3, 4, PRE_, POST_><<<grid, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y, pre, post, alpha); \
    else k_apply_fixed_rows<4, 3, PRE_, POST_><<<g3, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, y, pre, post, alpha);          \
  } while (0)
    if (pre && post) DDF_ROWS_CASE(true, true);
    else if (pre) DDF_ROWS_CASE(true, false);
    else if (post) DDF_ROWS_CASE(false, true);
    else DDF_ROWS_CASE(false, false);
#undef DDF_ROWS_CASE
    DDF_LAUNCH_CHECK(ctx);
    return 0;
  }
#define DDF_FIXED_CASE(WW)                                                                                          \
  case WW:                                                                                                          \
    if (pre && post) k_apply_fixed<WW, true, true><<<grid, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, pre, post, alpha, y); \
    else if (pre) k_apply_fixed<WW, true, false><<<grid, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, pre, post, alpha, y);   \
    else if (post) k_apply_fixed<WW, false, true><<<grid, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, pre, post, alpha, y);  \
    else k_apply_fixed<WW, false, false><<<grid, 256, 0, ctx->stream>>>(tab, nrows, first_negative, x, pre, post, alpha, y);           \
    break;
  switch (W) {
    DDF_FIXED_CASE(2)
    DDF_FIXED_CASE(3)
    DDF_FIXED_CASE(4)
    DDF_FIXED_CASE(5)
    DDF_FIXED_CASE(6)
    default: DDF_FAIL("apply: unsupported fixed width %d", W);
  }
#undef DDF_FIXED_CASE
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

static int launch_csrsign(ddf_ctx* ctx, const uint32_t* rowptr, const uint32_t* col, int64_t nrows, const double* x,
                          const double* pre, const double* post, double alpha, double* y) {
  if (nrows == 0) return 0;
  int grid;
  DDF_CHECK(grid_for(nrows, 256, &grid));
  if (pre && post) k_apply_csrsign<true, true><<<grid, 256, 0, ctx->stream>>>(rowptr, col, nrows, x, pre, post, alpha, y);
  else if (pre) k_apply_csrsign<true, false><<<grid, 256, 0, ctx->stream>>>(rowptr, col, nrows, x, pre, post, alpha, y);
  else if (post) k_apply_csrsign<false, true><<<grid, 256, 0, ctx->stream>>>(rowptr, col, nrows, x, pre, post, alpha, y);
  else k_apply_csrsign<false, false><<<grid, 256, 0, ctx->stream>>>(rowptr, col, nrows, x, pre, post, alpha, y);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

static int launch_diag(ddf_ctx* ctx, const double* d, double s, const double* x, double* y, int64_t n) {
  if (n == 0) return 0;
  int grid;
  DDF_CHECK(grid_for(ceil_div64(n, 4), 256, &grid));
  k_apply_diag<<<grid, 256, 0, ctx->stream>>>(d, s, x, y, n);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

// ======================================================== mesh diagonal payloads
__global__ void k_ratio(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __ddiv_rn(a[i], b[i]);
}
__global__ void k_mask_to_double(const uint8_t* __restrict__ m, double* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = m[i] ? 1.0 : 0.0;
}

int ddf_mesh_require_star(ddf_mesh* m, int R) {
  DDF_CHECK(ddf_mesh_require_geometry(m));
  if (m->star[R].p || m->n[R] == 0) return 0;
  DDF_CHECK(m->star[R].alloc(m->ctx, m->n[R]));
  DDF_CHECK(m->istar[R].alloc(m->ctx, m->n[R]));
  int grid;
  DDF_CHECK(grid_for(m->n[R], 256, &grid));
  k_ratio<<<grid, 256, 0, m->ctx->stream>>>(m->dualvol[R].p, m->vol[R].p, m->star[R].p, m->n[R]);    // dualvol ./ vol
  DDF_LAUNCH_CHECK(m->ctx);
  k_ratio<<<grid, 256, 0, m->ctx->stream>>>(m->vol[R].p, m->dualvol[R].p, m->istar[R].p, m->n[R]);   // vol ./ dualvol
  DDF_LAUNCH_CHECK(m->ctx);
  return 0;
}

static int bitsign(int b) { return (b & 1) ? -1 : 1; }

// ============================================================== op constructors
static OpRef make_node(ddf_mesh* m, OpKind k, int64_t nr, int64_t nc) {
  OpRef n = std::make_shared<OpNode>();
  n->kind = k;
  n->ctx = m->ctx;
  n->mesh = m;
  n->nrows = nr;
  n->ncols = nc;
  return n;
}

// the matrix d_R (n_{R-1} x n_R) as an operator: row-CSR with sign bits
static OpRef node_boundary_matrix(ddf_mesh* m, int R) {
  OpRef n = make_node(m, OPK_CSRSIGN, m->n[R - 1], m->n[R]);
  n->R = R;
  return n;
}
// its transpose (n_R x n_{R-1}): fixed-width rows
static OpRef node_boundary_transpose(ddf_mesh* m, int R) {
  OpRef n = make_node(m, OPK_FIXED, m->n[R], m->n[R - 1]);
  n->R = R;
  return n;
}

extern "C" int ddf_op_boundary(ddf_mesh* m, int P, int R, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  DDF_CHECK(ddf_mesh_require_assembled(m));
  if (P == DDF_PR) {
    if (!(0 < R && R <= m->D)) DDF_FAIL("boundary(Pr,R): need 0 < R <= D (ManifoldOps.jl:18)");
    *out = wrap(node_boundary_matrix(m, R));
  } else {
    if (!(0 <= R && R < m->D)) DDF_FAIL("boundary(Dl,R): need 0 <= R < D (ManifoldOps.jl:23)");
    *out = wrap(node_boundary_transpose(m, R + 1));
  }
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_deriv(ddf_mesh* m, int P, int R, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  DDF_CHECK(ddf_mesh_require_assembled(m));
  const int s = P == DDF_PR ? 1 : -1;
  if (!(0 <= R && R <= m->D) || !(0 <= R + s && R + s <= m->D))
    DDF_FAIL("deriv(P,R): need 0 <= R <= D and 0 <= R+s <= D (ManifoldOps.jl:43-44)");
  // adjoint(boundary(P, R+s)): Pr -> transpose of d_{R+1}; Dl -> adjoint(transpose d_R) = d_R
  if (P == DDF_PR) *out = wrap(node_boundary_transpose(m, R + 1));
  else *out = wrap(node_boundary_matrix(m, R));
  return 0;
  DDF_TRY_END
}

static int node_hodge(ddf_mesh* m, int P, int R, bool inverse, OpRef* out) {
  if (!(0 <= R && R <= m->D)) DDF_FAIL("hodge/invhodge: need 0 <= R <= D");
  DDF_CHECK(ddf_mesh_require_star(m, R));
  OpRef n = make_node(m, OPK_DIAG, m->n[R], m->n[R]);
  // hodge(Pr) = dualvol./vol ; invhodge(Dl) = vol./dualvol ;
  // hodge(Dl) = s*invhodge(Dl) ; invhodge(Pr) = s*hodge(Pr), s = (-1)^(R(D-R))  (ManifoldOps.jl:58-100)
  const int s = bitsign(R * (m->D - R));
  if (!inverse) {
    if (P == DDF_PR) { n->diag = m->star[R].p; n->scale = 1.0; }
    else { n->diag = m->istar[R].p; n->scale = s; }
  } else {
    if (P == DDF_DL) { n->diag = m->istar[R].p; n->scale = 1.0; }
    else { n->diag = m->star[R].p; n->scale = s; }
  }
  *out = n;
  return 0;
}

extern "C" int ddf_op_hodge(ddf_mesh* m, int P, int R, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpRef n;
  DDF_CHECK(node_hodge(m, P, R, false, &n));
  *out = wrap(n);
  return 0;
  DDF_TRY_END
}
extern "C" int ddf_op_invhodge(ddf_mesh* m, int P, int R, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpRef n;
  DDF_CHECK(node_hodge(m, P, R, true, &n));
  *out = wrap(n);
  return 0;
  DDF_TRY_END
}

static OpRef chain_of(std::vector<OpRef> f, double scalar) {
  OpRef n = std::make_shared<OpNode>();
  n->kind = OPK_CHAIN;
  n->ctx = f[0]->ctx;
  n->mesh = f[0]->mesh;
  n->nrows = f.front()->nrows;
  n->ncols = f.back()->ncols;
  n->scale = scalar;
  n->tree_ops = f;
  n->tree_scale = scalar;
  // flatten nested chains
  for (auto& g : f) {
    if (g->kind == OPK_CHAIN) {
      n->scale *= g->scale;
      for (auto& h : g->factors) n->factors.push_back(h);
    } else {
      n->factors.push_back(g);
    }
  }
  return n;
}

static int node_coderiv(ddf_mesh* m, int P, int R, OpRef* out) {
  const int dR = P == DDF_PR ? 1 : -1;
  if (!(0 <= R && R <= m->D) || !(0 <= R - dR && R - dR <= m->D))
    DDF_FAIL("coderiv(P,R): need 0 <= R <= D and 0 <= R-dR <= D (ManifoldOps.jl:114-115)");
  DDF_CHECK(ddf_mesh_require_assembled(m));
  // bitsign(R) * invhodge(!P, R-dR) * deriv(!P, R) * hodge(P, R)   (ManifoldOps.jl:117-120)
  OpRef ih, d, h;
  DDF_CHECK(node_hodge(m, P == DDF_PR ? DDF_DL : DDF_PR, R - dR, true, &ih));
  DDF_CHECK(node_hodge(m, P, R, false, &h));
  if (P == DDF_PR) d = node_boundary_matrix(m, R);          // deriv(Dl,R) = d_R   (n_{R-1} x n_R)
  else d = node_boundary_transpose(m, R + 1);                // deriv(Pr,R) = d_{R+1}^T
  OpRef ihs = clone_leaf(ih);                                // fold bitsign(R) into the left diagonal (exact)
  ihs->factors.clear();
  ihs->scale = ih->scale * bitsign(R);
  *out = chain_of({ihs, d, h}, 1.0);
  return 0;
}

extern "C" int ddf_op_coderiv(ddf_mesh* m, int P, int R, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpRef n;
  DDF_CHECK(node_coderiv(m, P, R, &n));
  *out = wrap(n);
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_laplace(ddf_mesh* m, int P, int R, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  if (!(0 <= R && R <= m->D)) DDF_FAIL("laplace(P,R): need 0 <= R <= D");
  DDF_CHECK(ddf_mesh_require_assembled(m));
  const int dR = P == DDF_PR ? 1 : -1;
  if (P == DDF_PR && R == 0 && m->D >= 1) {
    // delta_1 d_0 = -star0^-1 d0^T star1 d0 : fused vertex-adjacency kernel (wave.cu)
    DDF_CHECK(ddf_mesh_require_wave(m));
    *out = wrap(make_node(m, OPK_LAPLACE0, m->n[0], m->n[0]));
    return 0;
  }
  OpRef sum = std::make_shared<OpNode>();
  sum->kind = OPK_SUM;
  sum->ctx = m->ctx;
  sum->mesh = m;
  sum->nrows = sum->ncols = m->n[R];
  if (0 <= R - dR && R - dR <= m->D) {          // deriv(P, R-dR) * coderiv(P, R)
    OpRef cd;
    DDF_CHECK(node_coderiv(m, P, R, &cd));
    OpRef d = P == DDF_PR ? node_boundary_transpose(m, R) /* d_{R-1} = d_R^T table */
                          : node_boundary_matrix(m, R + 1) /* deriv(Dl,R+1) = d_{R+1} */;
    sum->factors.push_back(chain_of({d, cd}, 1.0));
    sum->coefs.push_back(1.0);
  }
  if (0 <= R + dR && R + dR <= m->D) {          // coderiv(P, R+dR) * deriv(P, R)
    OpRef cd;
    DDF_CHECK(node_coderiv(m, P, R + dR, &cd));
    OpRef d = P == DDF_PR ? node_boundary_transpose(m, R + 1) : node_boundary_matrix(m, R);
    sum->factors.push_back(chain_of({cd, d}, 1.0));
    sum->coefs.push_back(1.0);
  }
  if (sum->factors.empty()) { sum->kind = OPK_ZERO; }
  sum->tree_ops = sum->factors;          // op = zero; op += d*delta; op += delta*d  (ManifoldOps.jl:131-143)
  sum->tree_coefs = sum->coefs;
  *out = wrap(sum);
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_isboundary(ddf_mesh* m, int P, int R, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  if (P != DDF_PR) DDF_FAIL("isboundary: only defined for Pr (ManifoldOps.jl:31)");
  if (!(0 <= R && R < m->D)) DDF_FAIL("isboundary(Pr,R): need 0 <= R < D (ManifoldOps.jl:32)");
  DDF_CHECK(ddf_mesh_require_isboundary(m));
  if (!m->isbnd_d[R].p && m->n[R]) {
    DDF_CHECK(m->isbnd_d[R].alloc(m->ctx, m->n[R]));
    int grid;
    DDF_CHECK(grid_for(m->n[R], 256, &grid));
    k_mask_to_double<<<grid, 256, 0, m->ctx->stream>>>(m->isbnd[R].p, m->isbnd_d[R].p, m->n[R]);
    DDF_LAUNCH_CHECK(m->ctx);
  }
  OpRef n = make_node(m, OPK_DIAG, m->n[R], m->n[R]);
  n->diag = m->isbnd_d[R].p;
  *out = wrap(n);
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_identity(ddf_mesh* m, int R, double a, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  if (!(0 <= R && R <= m->D)) DDF_FAIL("identity: R out of range");
  if (R != 0 && R != m->D) DDF_CHECK(ddf_mesh_require_assembled(m));
  OpRef n = make_node(m, OPK_DIAG, m->n[R], m->n[R]);
  n->diag = nullptr;
  n->scale = a;
  *out = wrap(n);
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_zero(ddf_mesh* m, int R1, int R2, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  if (!(0 <= R1 && R1 <= m->D && 0 <= R2 && R2 <= m->D)) DDF_FAIL("zero: R out of range");
  DDF_CHECK(ddf_mesh_require_assembled(m));
  *out = wrap(make_node(m, OPK_ZERO, m->n[R1], m->n[R2]));
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_from_diag(ddf_ctx* ctx, int64_t n, const double* diag, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpRef o = std::make_shared<OpNode>();
  o->kind = OPK_DIAG;
  o->ctx = ctx;
  o->nrows = o->ncols = n;
  DDF_CHECK(o->diag_own.alloc(ctx, n));
  if (n) {
    DDF_CUDA(cudaMemcpyAsync(o->diag_own.p, diag, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  o->diag = o->diag_own.p;
  *out = wrap(o);
  return 0;
  DDF_TRY_END
}

// ---------------------------------------------------- CSC (reference storage) -> CSR
__global__ void k_expand_cols(const int64_t* __restrict__ colptr, int64_t ncols, int32_t* __restrict__ colidx) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  for (int64_t p = colptr[j] - 1; p < colptr[j + 1] - 1; p++) colidx[p] = (int32_t)j;
}
__global__ void k_rows_to_keys(const int64_t* __restrict__ rowval, int64_t nnz, uint32_t* __restrict__ keys,
                               uint32_t* __restrict__ idx) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) { keys[p] = (uint32_t)(rowval[p] - 1); idx[p] = (uint32_t)p; }
}
__global__ void k_gather_csr(const uint32_t* __restrict__ idx, const int32_t* __restrict__ colidx,
                             const double* __restrict__ nzval, int64_t nnz, int32_t* __restrict__ col,
                             double* __restrict__ val) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) { col[p] = colidx[idx[p]]; val[p] = nzval[idx[p]]; }
}
// rowptr[r] = first position whose key >= r (keys sorted)
__global__ void k_lower_bound(const uint32_t* __restrict__ keys, int64_t nnz, int64_t nrows, uint32_t* __restrict__ rowptr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nrows) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < (uint32_t)r) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = (uint32_t)lo;
}

extern "C" int ddf_op_from_csc(ddf_ctx* ctx, int64_t nrows, int64_t ncols, const int64_t* colptr,
                               const int64_t* rowval, const double* nzval, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  if (nrows >= 0x7fffffffLL || ncols >= 0x7fffffffLL) DDF_FAIL("from_csc: dimension >= 2^31");
  const int64_t nnz = colptr[ncols] - 1;
  if (nnz < 0 || nnz >= 0xffffffffLL) DDF_FAIL("from_csc: nnz %lld unsupported", (long long)nnz);
  DDF_CUDA(cudaSetDevice(ctx->device));
  OpRef o = std::make_shared<OpNode>();
  o->kind = OPK_CSRF64;
  o->ctx = ctx;
  o->nrows = nrows;
  o->ncols = ncols;
  o->nnz = nnz;
  DDF_CHECK(o->rowptr.alloc(ctx, nrows + 1));
  DDF_CHECK(o->col.alloc(ctx, nnz));
  DDF_CHECK(o->val.alloc(ctx, nnz));
  if (nnz == 0) {
    DDF_CUDA(cudaMemsetAsync(o->rowptr.p, 0, (nrows + 1) * 4, ctx->stream));
    *out = wrap(o);
    return 0;
  }
  DevBuf<int64_t> dcolptr, drowval;
  DevBuf<double> dval;
  DevBuf<int32_t> colidx;
  DevBuf<uint32_t> k0, k1, i0, i1;
  DevBuf<uint8_t> temp;
  DDF_CHECK(dcolptr.alloc(ctx, ncols + 1));
  DDF_CHECK(drowval.alloc(ctx, nnz));
  DDF_CHECK(dval.alloc(ctx, nnz));
  DDF_CHECK(colidx.alloc(ctx, nnz));
  DDF_CHECK(k0.alloc(ctx, nnz)); DDF_CHECK(k1.alloc(ctx, nnz));
  DDF_CHECK(i0.alloc(ctx, nnz)); DDF_CHECK(i1.alloc(ctx, nnz));
  DDF_CUDA(cudaMemcpyAsync(dcolptr.p, colptr, (ncols + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  DDF_CUDA(cudaMemcpyAsync(drowval.p, rowval, nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
  DDF_CUDA(cudaMemcpyAsync(dval.p, nzval, nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
  int grid;
  DDF_CHECK(grid_for(ncols, 256, &grid));
  k_expand_cols<<<grid, 256, 0, ctx->stream>>>(dcolptr.p, ncols, colidx.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CHECK(grid_for(nnz, 256, &grid));
  k_rows_to_keys<<<grid, 256, 0, ctx->stream>>>(drowval.p, nnz, k0.p, i0.p);
  DDF_LAUNCH_CHECK(ctx);
  int bits = 1;
  while (((int64_t)1 << bits) < nrows) bits++;
  cub::DoubleBuffer<uint32_t> dk(k0.p, k1.p), di(i0.p, i1.p);
  size_t tb = 0;
  DDF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, di, nnz, 0, bits, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, di, nnz, 0, bits, ctx->stream));   // stable: columns stay ascending
  ctx->launches += (bits + 7) / 8 + 1;
  k_gather_csr<<<grid, 256, 0, ctx->stream>>>(di.Current(), colidx.p, dval.p, nnz, o->col.p, o->val.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CHECK(grid_for(nrows + 1, 256, &grid));
  k_lower_bound<<<grid, 256, 0, ctx->stream>>>(dk.Current(), nnz, nrows, o->rowptr.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = wrap(o);
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_destroy(ddf_op* op) {
  if (!op) return 0;
  if (op->node && op->node->ctx) cudaStreamSynchronize(op->node->ctx->stream);
  delete op;
  return 0;
}
extern "C" int ddf_op_shape(ddf_op* op, int64_t* nrows, int64_t* ncols) {
  *nrows = op->node->nrows;
  *ncols = op->node->ncols;
  return 0;
}

// ==================================================================== Op algebra
extern "C" int ddf_op_compose(ddf_op* A, ddf_op* B, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  if (A->node->ncols != B->node->nrows)
    DDF_FAIL("compose: inner dimensions differ (%lld vs %lld)", (long long)A->node->ncols, (long long)B->node->nrows);
  if (A->node->kind == OPK_ZERO || B->node->kind == OPK_ZERO) {
    OpRef z = std::make_shared<OpNode>();
    z->kind = OPK_ZERO; z->ctx = A->node->ctx; z->mesh = A->node->mesh;
    z->nrows = A->node->nrows; z->ncols = B->node->ncols;
    *out = wrap(z);
    return 0;
  }
  *out = wrap(chain_of({A->node, B->node}, 1.0));
  return 0;
  DDF_TRY_END
}

__global__ void k_diag_combine(const double* __restrict__ a, double sa, const double* __restrict__ b, double sb,
                               double coef, double* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // Diagonal(a) + coef*Diagonal(b); identity payloads have null pointers
  double va = a ? __dmul_rn(sa, a[i]) : sa;
  double vb = b ? __dmul_rn(sb, b[i]) : sb;
  out[i] = __dadd_rn(va, __dmul_rn(coef, vb));
}

extern "C" int ddf_op_add(ddf_op* A, double b, ddf_op* B, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpRef a = A->node, bb = B->node;
  if (a->nrows != bb->nrows || a->ncols != bb->ncols) DDF_FAIL("add: shapes differ");
  if (a->kind == OPK_DIAG && bb->kind == OPK_DIAG) {     // Diagonal +- Diagonal stays diagonal (exact same arithmetic)
    OpRef o = std::make_shared<OpNode>();
    o->kind = OPK_DIAG; o->ctx = a->ctx; o->mesh = a->mesh; o->nrows = a->nrows; o->ncols = a->ncols;
    DDF_CHECK(o->diag_own.alloc(a->ctx, a->nrows));
    if (a->nrows) {
      int grid;
      DDF_CHECK(grid_for(a->nrows, 256, &grid));
      k_diag_combine<<<grid, 256, 0, a->ctx->stream>>>(a->diag, a->scale, bb->diag, bb->scale, b, o->diag_own.p, a->nrows);
      DDF_LAUNCH_CHECK(a->ctx);
    }
    o->diag = o->diag_own.p;
    o->scale = 1.0;
    // keep payload owners alive
    o->factors.push_back(a);
    o->factors.push_back(bb);
    *out = wrap(o);
    return 0;
  }
  OpRef s = std::make_shared<OpNode>();
  s->kind = OPK_SUM; s->ctx = a->ctx; s->mesh = a->mesh; s->nrows = a->nrows; s->ncols = a->ncols;
  auto push = [&](OpRef t, double c) {
    if (t->kind == OPK_ZERO) return;
    if (t->kind == OPK_SUM) {
      for (size_t i = 0; i < t->factors.size(); i++) { s->factors.push_back(t->factors[i]); s->coefs.push_back(c * t->coefs[i]); }
    } else { s->factors.push_back(t); s->coefs.push_back(c); }
  };
  push(a, 1.0);
  push(bb, b);
  s->tree_ops = {a, bb};
  s->tree_coefs = {1.0, b};
  if (s->factors.empty()) s->kind = OPK_ZERO;
  *out = wrap(s);
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_scale(double a, ddf_op* A, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpRef n = A->node;
  if (n->kind == OPK_ZERO) { *out = wrap(n); return 0; }
  if (n->kind == OPK_DIAG && (n->diag == nullptr || a == 1.0 || a == -1.0)) {
    OpRef o = clone_leaf(n);          // keeps an owned payload alive
    o->scale = n->scale * a;          // exact for +-1 and for the scaled identity
    *out = wrap(o);
    return 0;
  }
  if (n->kind == OPK_DIAG) {          // a * Diagonal(d) = Diagonal(a*d), one rounding per entry like the reference
    OpRef o = std::make_shared<OpNode>();
    o->kind = OPK_DIAG; o->ctx = n->ctx; o->mesh = n->mesh; o->nrows = n->nrows; o->ncols = n->ncols;
    DDF_CHECK(o->diag_own.alloc(n->ctx, n->nrows));
    if (n->nrows) {
      int grid;
      DDF_CHECK(grid_for(n->nrows, 256, &grid));
      k_diag_combine<<<grid, 256, 0, n->ctx->stream>>>(nullptr, 0.0, n->diag, n->scale, a, o->diag_own.p, n->nrows);
      DDF_LAUNCH_CHECK(n->ctx);
    }
    o->diag = o->diag_own.p;
    o->factors.push_back(n);
    *out = wrap(o);
    return 0;
  }
  *out = wrap(chain_of({n}, a));
  return 0;
  DDF_TRY_END
}

static int transpose_csrf64(OpRef src, OpRef* out);

static int adjoint_node(OpRef n, OpRef* out) {
  switch (n->kind) {
    case OPK_ZERO: {
      OpRef o = clone_leaf(n);
      o->factors.clear();
      std::swap(o->nrows, o->ncols);
      *out = o;
      return 0;
    }
    case OPK_FIXED: *out = node_boundary_matrix(n->mesh, n->R); return 0;
    case OPK_CSRSIGN: *out = node_boundary_transpose(n->mesh, n->R); return 0;
    case OPK_DIAG: *out = n; return 0;
    case OPK_CSRF64: return transpose_csrf64(n, out);
    case OPK_LAPLACE0: ddf_set_error("adjoint of the fused Laplacian is not available; use coderiv/deriv factors"); return 1;
    case OPK_CHAIN: {
      std::vector<OpRef> f;
      for (auto it = n->factors.rbegin(); it != n->factors.rend(); ++it) {
        OpRef a;
        DDF_CHECK(adjoint_node(*it, &a));
        f.push_back(a);
      }
      *out = chain_of(f, n->scale);
      (*out)->tree_ops.clear();
      (*out)->adj_of = n;
      return 0;
    }
    case OPK_SUM: {
      OpRef s = std::make_shared<OpNode>();
      s->kind = OPK_SUM; s->ctx = n->ctx; s->mesh = n->mesh; s->nrows = n->ncols; s->ncols = n->nrows;
      for (size_t i = 0; i < n->factors.size(); i++) {
        OpRef a;
        DDF_CHECK(adjoint_node(n->factors[i], &a));
        s->factors.push_back(a);
        s->coefs.push_back(n->coefs[i]);
      }
      s->adj_of = n;
      *out = s;
      return 0;
    }
  }
  return 1;
}

extern "C" int ddf_op_adjoint(ddf_op* A, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpRef o;
  DDF_CHECK(adjoint_node(A->node, &o));
  *out = wrap(o);
  return 0;
  DDF_TRY_END
}

// ====================================================================== Op * Fun
static double* temp_of(OpNode* n, size_t slot, int64_t len) {
  while (n->temps.size() <= slot) n->temps.emplace_back(new DevBuf<double>());
  DevBuf<double>* b = n->temps[slot].get();
  if ((int64_t)b->n < len) {
    if (b->alloc(n->ctx, (size_t)len)) return nullptr;
  }
  return b->p;
}

static int apply_node(OpNode* n, const double* x, double* y);

static bool is_sparse_leaf(const OpRef& f) { return f->kind == OPK_FIXED || f->kind == OPK_CSRSIGN; }

// sparse leaf with optional fused diagonals.  A diagonal payload with scale != 1
// cannot be fused bit-identically unless the scale is +-1 (sign flips commute with
// rounding), which is all the DEC operators ever produce.
// ---------------------------------------------------------------- upper runs of the boundary rows
// Row sigma of the boundary matrix d_R (R < D) lists the R-simplices that contain sigma, ascending.  Those that
// extend sigma at the END, (sigma, c) with c above every vertex of sigma, share the prefix sigma: in the lexicographic
// numbering (Manifolds.jl:735-757) they are consecutive ids, they are the LAST entries of the row, and the runs of
// consecutive rows tile the R-simplices in order -- every R-simplex is the upper coface of exactly one row, its prefix.
// So a row keeps its lower entries explicitly and its upper ones as [up[i], up[i+1]): a third (rows = edges) to a half
// (rows = vertices) of the index data goes away, and those gathers no longer wait for an index load.
__global__ void k_upper_count(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col, const int32_t* __restrict__ tab,
                              int W, int64_t nrows, int up_negative, uint32_t* __restrict__ nlow, uint32_t* __restrict__ nup,
                              uint32_t* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const uint32_t lo = rowptr[i], hi = rowptr[i + 1];
  uint32_t n = 0;
  bool ok = true;
  for (uint32_t p = hi; p > lo; p--) {                     // trailing entries whose prefix is this row
    const uint32_t e = col[p - 1], c = e & 0x7fffffffu;
    if ((uint32_t)tab[(int64_t)c * W] != (uint32_t)i) break;
    if ((int)(e >> 31) != up_negative) ok = false;
    if (n > 0 && (col[p] & 0x7fffffffu) != c + 1u) ok = false;     // consecutive ids
    n++;
  }
  for (uint32_t p = lo; p + n < hi; p++)                   // and no other entry has it
    if ((uint32_t)tab[(int64_t)(col[p] & 0x7fffffffu) * W] == (uint32_t)i) ok = false;
  if (!ok) atomicOr(bad, 1u);
  nlow[i] = hi - lo - n;
  nup[i] = n;
}
__global__ void k_upper_fill(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col, int64_t nrows,
                             const uint32_t* __restrict__ lowptr, const uint32_t* __restrict__ up,
                             uint32_t* __restrict__ lowcol, uint32_t* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const uint32_t lo = rowptr[i], hi = rowptr[i + 1], l0 = lowptr[i], nl = lowptr[i + 1] - l0;
  for (uint32_t q = 0; q < nl; q++) lowcol[l0 + q] = col[lo + q];
  if (lo + nl < hi && (col[lo + nl] & 0x7fffffffu) != up[i]) atomicOr(bad, 1u);      // the runs tile the columns in row order
  if (hi - lo - nl != up[i + 1] - up[i]) atomicOr(bad, 1u);
}

// bsell[R] <- SELL copy of the lower entries, bup[R] <- run offsets; returns with state 4 on success, 0 otherwise
static int upper_build(ddf_mesh* m, int R, int64_t nrows) {
  ddf_ctx* ctx = m->ctx;
  const int W = R + 1;
  const int64_t nnz = (int64_t)W * m->n[R];
  DevBuf<uint32_t> nlow, nup, lowptr, lowcol, bad;
  DevBuf<uint8_t> temp;
  DDF_CHECK(nlow.alloc(ctx, nrows + 1));
  DDF_CHECK(nup.alloc(ctx, nrows + 1));
  DDF_CHECK(lowptr.alloc(ctx, nrows + 1));
  DDF_CHECK(m->bup[R].alloc(ctx, nrows + 1));
  DDF_CHECK(bad.alloc(ctx, 1));
  DDF_CUDA(cudaMemsetAsync(bad.p, 0, 4, ctx->stream));
  DDF_CUDA(cudaMemsetAsync(nlow.p + nrows, 0, 4, ctx->stream));
  DDF_CUDA(cudaMemsetAsync(nup.p + nrows, 0, 4, ctx->stream));
  int grid;
  DDF_CHECK(grid_for(nrows, 256, &grid));
  k_upper_count<<<grid, 256, 0, ctx->stream>>>(m->brow_ptr[R].p, m->bcol[R].p, m->bnd_table(R), W, nrows, R & 1, nlow.p, nup.p, bad.p);
  DDF_LAUNCH_CHECK(ctx);
  size_t tb = 0;
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, nlow.p, lowptr.p, nrows + 1, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, nlow.p, lowptr.p, nrows + 1, ctx->stream));
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, nup.p, m->bup[R].p, nrows + 1, ctx->stream));
  ctx->launches += 2;
  uint32_t tot[2] = {0, 0};
  DDF_CUDA(cudaMemcpyAsync(&tot[0], lowptr.p + nrows, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaMemcpyAsync(&tot[1], m->bup[R].p + nrows, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  bool ok = (int64_t)tot[0] + tot[1] == nnz && (int64_t)tot[1] == m->n[R];
  if (ok) {
    DDF_CHECK(lowcol.alloc(ctx, tot[0] ? tot[0] : 1));
    k_upper_fill<<<grid, 256, 0, ctx->stream>>>(m->brow_ptr[R].p, m->bcol[R].p, nrows, lowptr.p, m->bup[R].p, lowcol.p, bad.p);
    DDF_LAUNCH_CHECK(ctx);
    uint32_t hb = 0;
    DDF_CUDA(cudaMemcpyAsync(&hb, bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
    ok = hb == 0;
  }
  if (ok) {
    DDF_CHECK(sell_build(ctx, lowptr.p, lowcol.p, nullptr, nrows, m->bsell[R], false, false));
    ok = m->bsell[R].padded <= (int64_t)tot[0] + tot[0] / 4 + 32 * DDF_SELL_WIN;
  }
  if (!ok) { m->bsell[R].release(); m->bup[R].release(); return 0; }
  // longest upper run of a 256-row window (+2 for the aligned start / even count): shared memory of the staged kernel
  {
    const int64_t nwin = ceil_div64(nrows, DDF_SELL_WIN);
    std::vector<uint32_t> h(nwin + 1);
    DDF_CUDA(cudaMemcpy2DAsync(h.data(), 4, m->bup[R].p, (size_t)DDF_SELL_WIN * 4, 4, nwin, cudaMemcpyDeviceToHost, ctx->stream));
    DDF_CUDA(cudaMemcpyAsync(&h[nwin], m->bup[R].p + nrows, 4, cudaMemcpyDeviceToHost, ctx->stream));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
    uint32_t mx = 0;
    for (int64_t w = 0; w < nwin; w++) mx = std::max(mx, h[w + 1] - h[w]);
    m->bup_cap[R] = (mx + 3u) & ~1u;
  }
  m->bsell_state[R] = 4;
  return 0;
}

static int apply_sparse(OpNode* f, const double* x, const OpNode* pre, const OpNode* post, double alpha, double* y) {
  const DdfTune& tn = f->ctx->tune;
  ddf_mesh* m = f->mesh;
  const double* pre_d = pre ? pre->diag : nullptr;
  const double* post_d = post ? post->diag : nullptr;
  double a = alpha;
  if (pre) a *= pre->scale;
  if (post) a *= post->scale;
  if (f->kind == OPK_FIXED) {
    const int W = f->R + 1;
    // default: prefix-packed rows (4 bytes less index data per row); tuning 18 keeps the explicit table
    // row range (ddf_op_restrict_rows), rounded outward to whole 1024-row blocks so that every storage keeps its
    // 32-row groups and vector widths; the few extra rows are correct values of rows next to the range
    int64_t r0 = 0, nr = f->nrows;
    if (f->row_hi >= 0) {
      r0 = f->row_lo / 1024 * 1024;
      const int64_t r1 = std::min<int64_t>((f->row_hi + 1023) / 1024 * 1024, f->nrows);
      nr = r1 > r0 ? r1 - r0 : 0;
      if (nr == 0) return 0;
    }
    {
      int fmt = 0;
      DDF_CHECK(dk_rows_format(m, f->R, &fmt));
      if (fmt >= 2) return launch_bits(f->ctx, m->bbits[f->R], W, nr, (f->R & 1), x, pre_d, post_d, a, y, r0);
      if (fmt == 1) return launch_packed(f->ctx, m->bpack[f->R], W, nr, (f->R & 1), x, pre_d, post_d, a, y, r0);
    }
    return launch_fixed(f->ctx, m->bnd_table(f->R), W, nr, (f->R & 1), x, pre_d, post_d, a, y, r0);
  }
  // boundary / dual derivative: SELL-32 copy of the row-CSR when its padding is small, else CSR
  const int R = f->R;
  if (m->bsell_state[R] == 0 && f->nrows > 0) {
    const int64_t nnz = (int64_t)(R + 1) * m->n[R];
    if (nnz < 3 * f->nrows) {
      m->bsell_state[R] = 2;           // rows of 1-2 entries (faces of the top simplices): pair table when every row fits
      DevBuf<uint32_t> flag;
      DDF_CHECK(flag.alloc(f->ctx, 1));
      DDF_CUDA(cudaMemsetAsync(flag.p, 0, 4, f->ctx->stream));
      DDF_CHECK(m->bpair[R].alloc(f->ctx, 2 * (size_t)f->nrows));
      int gb;
      DDF_CHECK(grid_for(f->nrows, 256, &gb));
      k_build_pairs<<<gb, 256, 0, f->ctx->stream>>>(m->brow_ptr[R].p, m->bcol[R].p, f->nrows, m->bpair[R].p, flag.p);
      DDF_LAUNCH_CHECK(f->ctx);
      uint32_t too_long = 0;
      DDF_CUDA(cudaMemcpyAsync(&too_long, flag.p, 4, cudaMemcpyDeviceToHost, f->ctx->stream));
      DDF_CUDA(cudaStreamSynchronize(f->ctx->stream));
      if (too_long) m->bpair[R].release(); else m->bsell_state[R] = 3;
    } else {
      // length-only sort: rows of equal length stay in row order (consecutive y stores, measured 20 % faster
      // than the stencil-signature sort on the Kuhn mesh)
      // lower entries + upper runs: pays for long rows (vertex rows of a 3D mesh, 14 entries: 0.484 -> 0.436 ms at C4);
      // rows of ~5 entries lose to the plain SELL copy (1.21 -> 1.39-1.51 ms), tuning 22 forces it for them anyway
      if (R < m->D && tn.fixed != 20 && (nnz >= 8 * f->nrows || tn.fixed == 22 || tn.fixed == 21))
        DDF_CHECK(upper_build(m, R, f->nrows));
      if (m->bsell_state[R] == 0) {
        DDF_CHECK(sell_build(f->ctx, m->brow_ptr[R].p, m->bcol[R].p, nullptr, f->nrows, m->bsell[R], false, false));
        if (m->bsell[R].padded <= nnz + nnz / 8) m->bsell_state[R] = 1;
        else { m->bsell_state[R] = 2; m->bsell[R].release(); }
      }
    }
  }
  if (m->bsell_state[R] == 3 && tn.fixed != 9) {
    const int grid = (int)ceil_div64(ceil_div64(f->nrows, 4), 256);
    const uint32_t* pp = m->bpair[R].p;
    if (pre_d && post_d) k_apply_pairsign<true, true, 4><<<grid, 256, 0, f->ctx->stream>>>(pp, f->nrows, x, pre_d, post_d, a, y, tn.prefetch);
    else if (pre_d) k_apply_pairsign<true, false, 4><<<grid, 256, 0, f->ctx->stream>>>(pp, f->nrows, x, pre_d, post_d, a, y, tn.prefetch);
    else if (post_d) k_apply_pairsign<false, true, 4><<<grid, 256, 0, f->ctx->stream>>>(pp, f->nrows, x, pre_d, post_d, a, y, tn.prefetch);
    else if (a == 1.0 && tn.fixed != 32) k_apply_pairsign32<4><<<grid, 256, 0, f->ctx->stream>>>(pp, f->nrows, x, y, tn.prefetch);
    else k_apply_pairsign<false, false, 4><<<grid, 256, 0, f->ctx->stream>>>(pp, f->nrows, x, pre_d, post_d, a, y, tn.prefetch);   // tuning 32: the 40-register build
    DDF_LAUNCH_CHECK(f->ctx);
    return 0;
  }
  if (m->bsell_state[R] == 4 && tn.fixed != 9) {
    const int grid = (int)ceil_div64(f->nrows, DDF_SELL_WIN);
    const SellBufs& sb = m->bsell[R];
    const uint32_t* up = m->bup[R].p;
    const bool short_rows = (int64_t)(R + 1) * m->n[R] < 8 * f->nrows;
    // staged upper runs: one TMA bulk copy per window and array, when the longest window run fits ~24 KB of shared
    // memory (8 blocks per SM stay resident) and x / pre are 16-byte aligned; tuning 21 keeps the global-memory loop
    const uint32_t cap = m->bup_cap[R];
    const size_t smem = (size_t)cap * 8 * (pre_d ? 2 : 1);
    const bool stage = tn.fixed != 21 && smem <= 24 * 1024 + (pre_d ? 12 * 1024 : 0) && ((uintptr_t)x & 15) == 0 &&
                       (!pre_d || ((uintptr_t)pre_d & 15) == 0);
#define DDF_SELLUP(PRE_, POST_)                                                                                     \
  do {                                                                                                              \
    if (stage && short_rows) k_apply_sellsign<PRE_, POST_, 4, true, true><<<grid, DDF_SELL_WIN, smem, f->ctx->stream>>>(sb.ptr.p, sb.perm8.p, sb.col.p, f->nrows, x, pre_d, post_d, a, y, up, R & 1, cap, tn.prefetch); \
    else if (stage) k_apply_sellsign<PRE_, POST_, 8, true, true><<<grid, DDF_SELL_WIN, smem, f->ctx->stream>>>(sb.ptr.p, sb.perm8.p, sb.col.p, f->nrows, x, pre_d, post_d, a, y, up, R & 1, cap, tn.prefetch);            \
    else if (short_rows) k_apply_sellsign<PRE_, POST_, 4, true><<<grid, DDF_SELL_WIN, 0, f->ctx->stream>>>(sb.ptr.p, sb.perm8.p, sb.col.p, f->nrows, x, pre_d, post_d, a, y, up, R & 1, 0, tn.prefetch); \
    else k_apply_sellsign<PRE_, POST_, 8, true><<<grid, DDF_SELL_WIN, 0, f->ctx->stream>>>(sb.ptr.p, sb.perm8.p, sb.col.p, f->nrows, x, pre_d, post_d, a, y, up, R & 1, 0, tn.prefetch);            \
  } while (0)
    if (pre_d && post_d) DDF_SELLUP(true, true);
    else if (pre_d) DDF_SELLUP(true, false);
    else if (post_d) DDF_SELLUP(false, true);
    else DDF_SELLUP(false, false);
#undef DDF_SELLUP
    DDF_LAUNCH_CHECK(f->ctx);
    return 0;
  }
  if (m->bsell_state[R] == 1 && tn.fixed != 9) {
    const int grid = (int)ceil_div64(f->nrows, DDF_SELL_WIN);
    const SellBufs& sb = m->bsell[R];
    // short rows (edges -> faces: ~5 entries): 4 gathers per round keep the kernel at 32 registers / full occupancy;
    // the index words of up to 8 entries are fetched in ONE round trip (1.230 -> 1.136 ms at C4; tuning 28: 4 at a time)
    const bool short_rows = (int64_t)(R + 1) * m->n[R] < 8 * f->nrows && tn.fixed != 17;
#define DDF_SELLSIGN(PRE_, POST_)                                                                                   \
  do {                                                                                                              \
    if (short_rows && tn.fixed == 28) k_apply_sellsign<PRE_, POST_, 4><<<grid, DDF_SELL_WIN, 0, f->ctx->stream>>>(sb.ptr.p, sb.perm8.p, sb.col.p, f->nrows, x, pre_d, post_d, a, y, nullptr, 0, 0, tn.prefetch); \
    else if (short_rows) k_apply_sellsign<PRE_, POST_, 4, false, false, 8><<<grid, DDF_SELL_WIN, 0, f->ctx->stream>>>(sb.ptr.p, sb.perm8.p, sb.col.p, f->nrows, x, pre_d, post_d, a, y, nullptr, 0, 0, tn.prefetch); \
    else k_apply_sellsign<PRE_, POST_, 8><<<grid, DDF_SELL_WIN, 0, f->ctx->stream>>>(sb.ptr.p, sb.perm8.p, sb.col.p, f->nrows, x, pre_d, post_d, a, y, nullptr, 0, 0, tn.prefetch);            \
  } while (0)
    if (pre_d && post_d) DDF_SELLSIGN(true, true);
    else if (pre_d) DDF_SELLSIGN(true, false);
    else if (post_d) DDF_SELLSIGN(false, true);
    else DDF_SELLSIGN(false, false);
#undef DDF_SELLSIGN
    DDF_LAUNCH_CHECK(f->ctx);
    return 0;
  }
  return launch_csrsign(f->ctx, m->brow_ptr[R].p, m->bcol[R].p, f->nrows, x, pre_d, post_d, a, y);
}

static bool fusable_diag(const OpRef& f) {
  return f->kind == OPK_DIAG && f->diag != nullptr && (f->scale == 1.0 || f->scale == -1.0);
}

static int apply_chain(OpNode* n, const double* x, double* y) {
  // evaluate right to left; stage = [post-diag] sparse [pre-diag]  or a lone factor
  const int nf = (int)n->factors.size();
  int i = nf - 1;
  const double* cur = x;
  size_t slot = 0;
  while (i >= 0) {
    int next_i;
    const OpNode *pre = nullptr, *post = nullptr;
    OpNode* core;
    if (fusable_diag(n->factors[i]) && i - 1 >= 0 && is_sparse_leaf(n->factors[i - 1])) {
      pre = n->factors[i].get();
      core = n->factors[i - 1].get();
      next_i = i - 2;
    } else {
      core = n->factors[i].get();
      next_i = i - 1;
    }
    if ((core->kind == OPK_FIXED || core->kind == OPK_CSRSIGN) && next_i >= 0 && fusable_diag(n->factors[next_i])) {
      post = n->factors[next_i].get();
      next_i--;
    }
    const bool last = next_i < 0;
    const double alpha = last ? n->scale : 1.0;
    double* dst = last ? y : temp_of(n, slot++ & 1, core->nrows);
    if (!dst) return 1;
    if (core->kind == OPK_FIXED || core->kind == OPK_CSRSIGN) {
      DDF_CHECK(apply_sparse(core, cur, pre, post, alpha, dst));
    } else {
      DDF_CHECK(apply_node(core, cur, dst));
      if (alpha != 1.0) DDF_CHECK(launch_diag(n->ctx, nullptr, alpha, dst, dst, core->nrows));
    }
    cur = dst;
    i = next_i;
  }
  return 0;
}

__global__ void k_axpy_inplace(double c, const double* __restrict__ t, double* __restrict__ y, int64_t n, int first) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = c == 1.0 ? t[i] : __dmul_rn(c, t[i]);
  y[i] = first ? v : __dadd_rn(y[i], v);
}

static int apply_node(OpNode* n, const double* x, double* y) {
  const DdfTune& tn = n->ctx->tune;
  ddf_ctx* ctx = n->ctx;
  switch (n->kind) {
    case OPK_ZERO:
      if (n->nrows) DDF_CUDA(cudaMemsetAsync(y, 0, n->nrows * sizeof(double), ctx->stream));
      return 0;
    case OPK_FIXED:
    case OPK_CSRSIGN: return apply_sparse(n, x, nullptr, nullptr, 1.0, y);
    case OPK_DIAG: {
      if (n->row_hi < 0) return launch_diag(ctx, n->diag, n->scale, x, y, n->nrows);
      const int64_t r0 = n->row_lo / 4 * 4;                      // 128-bit accesses: keep the 32-byte phase of the vectors
      const int64_t r1 = std::min<int64_t>((n->row_hi + 3) / 4 * 4, n->nrows);
      return r1 > r0 ? launch_diag(ctx, n->diag ? n->diag + r0 : nullptr, n->scale, x + r0, y + r0, r1 - r0) : 0;
    }
    case OPK_CSRF64: {
      if (!n->nrows) return 0;
      if (n->sell_state == 0) {                  // large operators: SELL copy of the rows, once
        n->sell_state = 2;
        if (n->nrows >= 4 * DDF_SELL_WIN && n->nnz >= 2 * n->nrows && !(tn.fixed == 9) &&
            sell_build(ctx, n->rowptr.p, reinterpret_cast<const uint32_t*>(n->col.p), n->val.p, n->nrows, n->sell,
                       false, true, false) == 0 &&
            n->sell.padded <= 2 * n->nnz)
          n->sell_state = 1;
        else
          n->sell.release();
      }
      if (n->sell_state == 1) {
        const int64_t nwin = ceil_div64(n->nrows, DDF_SELL_WIN);
        k_apply_sellf64<<<(int)nwin, DDF_SELL_WIN, 0, ctx->stream>>>(n->sell.ptr.p, n->sell.perm8.p, n->sell.col.p,
                                                                     n->sell.val.p, n->nrows, x, y);
        DDF_LAUNCH_CHECK(ctx);
        return 0;
      }
      int grid;
      DDF_CHECK(grid_for(n->nrows, 256, &grid));
      k_apply_csrf64<<<grid, 256, 0, ctx->stream>>>(n->rowptr.p, n->col.p, n->val.p, n->nrows, x, 1.0, y);
      DDF_LAUNCH_CHECK(ctx);
      return 0;
    }
    case OPK_LAPLACE0: return ddf_laplace0_apply(n->mesh, x, y);
    case OPK_CHAIN: return apply_chain(n, x, y);
    case OPK_SUM: {
      // terms are evaluated into a scratch vector and accumulated in order
      for (size_t t = 0; t < n->factors.size(); t++) {
        double* tmp = temp_of(n, 2, n->nrows);
        if (!tmp && n->nrows) return 1;
        DDF_CHECK(apply_node(n->factors[t].get(), x, tmp));
        if (n->nrows) {
          int grid;
          DDF_CHECK(grid_for(n->nrows, 256, &grid));
          k_axpy_inplace<<<grid, 256, 0, ctx->stream>>>(n->coefs[t], tmp, y, n->nrows, t == 0 ? 1 : 0);
          DDF_LAUNCH_CHECK(ctx);
        }
      }
      return 0;
    }
  }
  return 1;
}

// An operator that writes rows [row0, row1) only -- what a rank of a slab-decomposed mesh OWNS (rows of
// ghost simplices belong to the neighbour and are never used).  Leaf operators of the hot path: deriv(Pr,k) and the
// diagonal operators; the range may be widened to the storage's block size (rows next to the range, correct values).
extern "C" int ddf_op_restrict_rows(ddf_op* A, int64_t row0, int64_t row1, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpRef n = A->node;
  if (n->kind != OPK_FIXED && n->kind != OPK_DIAG)
    DDF_FAIL("restrict_rows: supported for deriv(Pr,k) and diagonal operators (hodge, invhodge, isboundary)");
  if (row0 < 0 || row1 < row0 || row1 > n->nrows) DDF_FAIL("restrict_rows: need 0 <= row0 <= row1 <= %lld", (long long)n->nrows);
  OpRef o = clone_leaf(n);
  o->row_lo = row0;
  o->row_hi = row1;
  *out = wrap(o);
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_apply(ddf_op* A, ddf_vec* x, ddf_vec* y) {
  DDF_TRY_BEGIN
  OpNode* n = A->node.get();
  if (x->n != n->ncols) DDF_FAIL("apply: x has length %lld, operator has %lld columns", (long long)x->n, (long long)n->ncols);
  if (y->n != n->nrows) DDF_FAIL("apply: y has length %lld, operator has %lld rows", (long long)y->n, (long long)n->nrows);
  if (x == y) DDF_FAIL("apply: x and y must not alias");
  DDF_CHECK(ddf_vec_settle(x));
  DDF_CHECK(ddf_vec_settle(y));
  return apply_node(n, x->d.p, y->d.p);
  DDF_TRY_END
}

// ============================================================ assembled export
// Materialise an operator as a device COO list in CSC order (col ascending, row
// ascending within a column) with the Op constructor's chop (Ops.jl:44-58).
struct Coo {
  DevBuf<int32_t> row, col;
  DevBuf<double> val;
  int64_t nnz = 0;
  bool is_diag = false;     // a Diagonal in the reference: products of Diagonals stay Diagonal and are never chopped
};

// d_R^T (FIXED, n_R x n_{R-1}): CSC columns are faces -> parents = the row-CSR of d_R
__global__ void k_coo_from_rowcsr(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colw, int64_t nfaces,
                                  const double* __restrict__ dl, double sl, const double* __restrict__ dr, double sr,
                                  int32_t* __restrict__ row, int32_t* __restrict__ col, double* __restrict__ val,
                                  int transpose) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfaces) return;
  for (uint32_t p = rowptr[f]; p < rowptr[f + 1]; p++) {
    const uint32_t cw = colw[p];
    const int32_t j = (int32_t)(cw & 0x7fffffffu);
    double v = (cw >> 31) ? -1.0 : 1.0;
    // matrix entry M[j, f] (transpose == 1: M = d_R^T) ; left diag indexed by row, right by column
    const int32_t r = transpose ? j : (int32_t)f, c = transpose ? (int32_t)f : j;
    if (dl || sl != 1.0) { v = __dmul_rn(dl ? __dmul_rn(sl, dl[r]) : sl, v); if (v * v < DDF_CHOP_ABS2) v = 0.0; }
    if (dr || sr != 1.0) v = __dmul_rn(v, dr ? __dmul_rn(sr, dr[c]) : sr);
    row[p] = r; col[p] = c; val[p] = v;
  }
}
// d_R (CSRSIGN, n_{R-1} x n_R): CSC columns are simplices -> faces = the fixed-width table
__global__ void k_coo_from_table(const int32_t* __restrict__ tab, int64_t nsimp, int W, int first_negative,
                                 const double* __restrict__ dl, double sl, const double* __restrict__ dr, double sr,
                                 int32_t* __restrict__ row, int32_t* __restrict__ col, double* __restrict__ val) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nsimp) return;
  for (int p = 0; p < W; p++) {
    const int32_t f = tab[j * W + p];
    double v = (((p & 1) != 0) != (first_negative != 0)) ? -1.0 : 1.0;
    if (dl || sl != 1.0) { v = __dmul_rn(dl ? __dmul_rn(sl, dl[f]) : sl, v); if (v * v < DDF_CHOP_ABS2) v = 0.0; }
    if (dr || sr != 1.0) v = __dmul_rn(v, dr ? __dmul_rn(sr, dr[j]) : sr);
    row[j * W + p] = f; col[j * W + p] = (int32_t)j; val[j * W + p] = v;
  }
}
__global__ void k_coo_diag(const double* __restrict__ d, double s, int64_t n, int32_t* __restrict__ row,
                           int32_t* __restrict__ col, double* __restrict__ val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  row[i] = col[i] = (int32_t)i;
  val[i] = d ? __dmul_rn(s, d[i]) : s;
}
__global__ void k_chop_flags(const double* __restrict__ val, int64_t nnz, double thr, uint8_t* __restrict__ keep) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) keep[p] = (val[p] * val[p] < thr) ? 0 : 1;    // abs2(f) < eps34sq  (Ops.jl:52)
}

static int coo_alloc(ddf_ctx* ctx, Coo& c, int64_t nnz) {
  c.nnz = nnz;
  DDF_CHECK(c.row.alloc(ctx, nnz));
  DDF_CHECK(c.col.alloc(ctx, nnz));
  DDF_CHECK(c.val.alloc(ctx, nnz));
  return 0;
}

static int coo_chop(ddf_ctx* ctx, Coo& c) {
  if (c.nnz == 0) return 0;
  DevBuf<uint8_t> keep, temp;
  DevBuf<int64_t> nsel;
  Coo o;
  DDF_CHECK(keep.alloc(ctx, c.nnz));
  DDF_CHECK(nsel.alloc(ctx, 1));
  DDF_CHECK(coo_alloc(ctx, o, c.nnz));
  int grid;
  DDF_CHECK(grid_for(c.nnz, 256, &grid));
  const double thr = DDF_CHOP_ABS2;
  k_chop_flags<<<grid, 256, 0, ctx->stream>>>(c.val.p, c.nnz, thr, keep.p);
  DDF_LAUNCH_CHECK(ctx);
  size_t tb = 0, tb2 = 0, tb3 = 0;
  DDF_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, c.row.p, keep.p, o.row.p, nsel.p, c.nnz, ctx->stream));
  DDF_CUDA(cub::DeviceSelect::Flagged(nullptr, tb2, c.val.p, keep.p, o.val.p, nsel.p, c.nnz, ctx->stream));
  tb3 = tb > tb2 ? tb : tb2;
  DDF_CHECK(temp.alloc(ctx, tb3));
  DDF_CUDA(cub::DeviceSelect::Flagged(temp.p, tb3, c.row.p, keep.p, o.row.p, nsel.p, c.nnz, ctx->stream));
  DDF_CUDA(cub::DeviceSelect::Flagged(temp.p, tb3, c.col.p, keep.p, o.col.p, nsel.p, c.nnz, ctx->stream));
  DDF_CUDA(cub::DeviceSelect::Flagged(temp.p, tb3, c.val.p, keep.p, o.val.p, nsel.p, c.nnz, ctx->stream));
  ctx->launches += 3;
  int64_t h = 0;
  DDF_CUDA(cudaMemcpyAsync(&h, nsel.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  o.nnz = h;
  c.row.swap(o.row); c.col.swap(o.col); c.val.swap(o.val);
  c.nnz = h;
  return 0;
}

__global__ void k_csr_to_coo(const uint32_t* __restrict__ rowptr, int64_t nrows, int32_t* __restrict__ row) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  for (uint32_t p = rowptr[i]; p < rowptr[i + 1]; p++) row[p] = (int32_t)i;
}
__global__ void k_gather_i32(const uint32_t* __restrict__ idx, const int32_t* __restrict__ src, int64_t n, int32_t* __restrict__ dst) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = src[idx[p]];
}
__global__ void k_gather_f64(const uint32_t* __restrict__ idx, const double* __restrict__ src, int64_t n, double* __restrict__ dst) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = src[idx[p]];
}
__global__ void k_iota_u32(uint32_t* __restrict__ p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}
__global__ void k_i32_to_u32(const int32_t* __restrict__ s, uint32_t* __restrict__ d, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (uint32_t)s[i];
}

// stable sort of a row-major COO by column -> CSC order
static int coo_sort_by_col(ddf_ctx* ctx, Coo& c, int64_t ncols) {
  if (c.nnz == 0) return 0;
  DevBuf<uint32_t> k0, k1, i0, i1;
  DevBuf<uint8_t> temp;
  DDF_CHECK(k0.alloc(ctx, c.nnz)); DDF_CHECK(k1.alloc(ctx, c.nnz));
  DDF_CHECK(i0.alloc(ctx, c.nnz)); DDF_CHECK(i1.alloc(ctx, c.nnz));
  int grid;
  DDF_CHECK(grid_for(c.nnz, 256, &grid));
  k_i32_to_u32<<<grid, 256, 0, ctx->stream>>>(c.col.p, k0.p, c.nnz);
  DDF_LAUNCH_CHECK(ctx);
  k_iota_u32<<<grid, 256, 0, ctx->stream>>>(i0.p, c.nnz);
  DDF_LAUNCH_CHECK(ctx);
  int bits = 1;
  while (((int64_t)1 << bits) < ncols) bits++;
  cub::DoubleBuffer<uint32_t> dk(k0.p, k1.p), di(i0.p, i1.p);
  size_t tb = 0;
  DDF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, di, c.nnz, 0, bits, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, di, c.nnz, 0, bits, ctx->stream));
  ctx->launches += (bits + 7) / 8 + 1;
  Coo o;
  DDF_CHECK(coo_alloc(ctx, o, c.nnz));
  k_gather_i32<<<grid, 256, 0, ctx->stream>>>(di.Current(), c.row.p, c.nnz, o.row.p);
  k_gather_i32<<<grid, 256, 0, ctx->stream>>>(di.Current(), c.col.p, c.nnz, o.col.p);
  k_gather_f64<<<grid, 256, 0, ctx->stream>>>(di.Current(), c.val.p, c.nnz, o.val.p);
  ctx->launches += 3;
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  c.row.swap(o.row); c.col.swap(o.col); c.val.swap(o.val);
  return 0;
}

static int transpose_csrf64(OpRef src, OpRef* out) {
  ddf_ctx* ctx = src->ctx;
  OpRef o = std::make_shared<OpNode>();
  o->kind = OPK_CSRF64; o->ctx = ctx; o->mesh = src->mesh;
  o->nrows = src->ncols; o->ncols = src->nrows; o->nnz = src->nnz;
  DDF_CHECK(o->rowptr.alloc(ctx, o->nrows + 1));
  DDF_CHECK(o->col.alloc(ctx, o->nnz));
  DDF_CHECK(o->val.alloc(ctx, o->nnz));
  if (o->nnz == 0) {
    DDF_CUDA(cudaMemsetAsync(o->rowptr.p, 0, (o->nrows + 1) * 4, ctx->stream));
    *out = o;
    return 0;
  }
  Coo c;
  DDF_CHECK(coo_alloc(ctx, c, src->nnz));
  int grid;
  DDF_CHECK(grid_for(src->nrows, 256, &grid));
  k_csr_to_coo<<<grid, 256, 0, ctx->stream>>>(src->rowptr.p, src->nrows, c.row.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaMemcpyAsync(c.col.p, src->col.p, src->nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  DDF_CUDA(cudaMemcpyAsync(c.val.p, src->val.p, src->nnz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  DDF_CHECK(coo_sort_by_col(ctx, c, src->ncols));
  // transposed CSR: rows = old columns (sorted), cols = old rows
  DDF_CUDA(cudaMemcpyAsync(o->col.p, c.row.p, o->nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  DDF_CUDA(cudaMemcpyAsync(o->val.p, c.val.p, o->nnz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  DevBuf<uint32_t> keys;
  DDF_CHECK(keys.alloc(ctx, o->nnz));
  DDF_CHECK(grid_for(o->nnz, 256, &grid));
  k_i32_to_u32<<<grid, 256, 0, ctx->stream>>>(c.col.p, keys.p, o->nnz);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CHECK(grid_for(o->nrows + 1, 256, &grid));
  k_lower_bound<<<grid, 256, 0, ctx->stream>>>(keys.p, o->nnz, o->nrows, o->rowptr.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = o;
  return 0;
}

// ----------------------------------------------------------------------------- N3: assembled Op algebra
// SparseArrays' A*B (spmatmul): for every column j of B, for k in B[:,j] ascending, for i in A[:,k]
// ascending: the first touch of C[i,j] assigns fl(A[i,k]*B[k,j]), later touches add -- i.e. a left fold
// over k ascending.  Here: expand every product in exactly that order, stable radix sort by (j, i)
// (ties keep the expansion order), left fold of every run.  Bit-identical values, CSC-ordered output.
__global__ void k_count_products(const int32_t* __restrict__ brow, int64_t bnnz, const uint32_t* __restrict__ acolptr,
                                 uint32_t* __restrict__ cnt) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < bnnz) { const int32_t k = brow[p]; cnt[p] = acolptr[k + 1] - acolptr[k]; }
  if (p == bnnz) cnt[p] = 0;
}
__global__ void k_expand_products(const int32_t* __restrict__ brow, const int32_t* __restrict__ bcol,
                                  const double* __restrict__ bval, int64_t bnnz, const uint32_t* __restrict__ acolptr,
                                  const int32_t* __restrict__ arow, const double* __restrict__ aval,
                                  const uint32_t* __restrict__ off, unsigned long long* __restrict__ key,
                                  double* __restrict__ val) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= bnnz) return;
  const int32_t k = brow[p];
  const unsigned long long j = (unsigned long long)(uint32_t)bcol[p] << 32;
  const double b = bval[p];
  uint32_t o = off[p];
  for (uint32_t q = acolptr[k]; q < acolptr[k + 1]; q++, o++) {
    key[o] = j | (uint32_t)arow[q];
    val[o] = __dmul_rn(aval[q], b);
  }
}
__global__ void k_key_heads(const unsigned long long* __restrict__ key, int64_t n, uint8_t* __restrict__ head) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) head[p] = (p == 0 || key[p] != key[p - 1]) ? 1 : 0;
}
// one thread per run head: sequential left fold of the run, output at the run's rank
__global__ void k_fold_runs(const unsigned long long* __restrict__ key, const double* __restrict__ val,
                            const uint8_t* __restrict__ head, const uint32_t* __restrict__ rank, int64_t n,
                            int32_t* __restrict__ row, int32_t* __restrict__ col, double* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n || !head[p]) return;
  double acc = val[p];
  for (int64_t q = p + 1; q < n && !head[q]; q++) acc = __dadd_rn(acc, val[q]);
  const uint32_t r = rank[p];
  row[r] = (int32_t)(key[p] & 0xffffffffull);
  col[r] = (int32_t)(key[p] >> 32);
  out[r] = acc;
}
__global__ void k_scale_vals(double* __restrict__ v, int64_t n, double a) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) v[p] = __dmul_rn(a, v[p]);
}
__global__ void k_pack_keys(const int32_t* __restrict__ row, const int32_t* __restrict__ col, int64_t n,
                            unsigned long long* __restrict__ key) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) key[p] = ((unsigned long long)(uint32_t)col[p] << 32) | (uint32_t)row[p];
}

static void coo_move(Coo& dst, Coo& src) {
  dst.row.swap(src.row); dst.col.swap(src.col); dst.val.swap(src.val);
  dst.nnz = src.nnz; dst.is_diag = src.is_diag;
}
static int coo_scale(ddf_ctx* ctx, Coo& c, double a) {
  if (a == 1.0 || c.nnz == 0) return 0;
  int grid;
  DDF_CHECK(grid_for(c.nnz, 256, &grid));
  k_scale_vals<<<grid, 256, 0, ctx->stream>>>(c.val.p, c.nnz, a);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}
static int bits_for(int64_t n) { int b = 1; while (((int64_t)1 << b) < n) b++; return b; }

// (key, val) in expansion order -> sorted by key (stable), runs folded left to right -> Coo in CSC order
static int coo_from_keyed(ddf_ctx* ctx, DevBuf<unsigned long long>& key, DevBuf<double>& val, int64_t n, int64_t ncols,
                          Coo& out) {
  if (n == 0) return coo_alloc(ctx, out, 0);
  if (n > 0xFFFFFFF0LL) DDF_FAIL("assembled product: %lld partial products exceed the 32-bit staging index", (long long)n);
  DevBuf<unsigned long long> key2;
  DevBuf<uint32_t> i0, i1, rank;
  DevBuf<uint8_t> temp, head;
  DevBuf<double> sval;
  DDF_CHECK(key2.alloc(ctx, n)); DDF_CHECK(i0.alloc(ctx, n)); DDF_CHECK(i1.alloc(ctx, n));
  int grid;
  DDF_CHECK(grid_for(n, 256, &grid));
  k_iota_u32<<<grid, 256, 0, ctx->stream>>>(i0.p, n);
  DDF_LAUNCH_CHECK(ctx);
  cub::DoubleBuffer<unsigned long long> dk(key.p, key2.p);
  cub::DoubleBuffer<uint32_t> di(i0.p, i1.p);
  size_t tb = 0;
  const int end_bit = 32 + bits_for(ncols);
  DDF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, di, n, 0, end_bit, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, di, n, 0, end_bit, ctx->stream));
  ctx->launches += (end_bit + 7) / 8 + 1;
  DDF_CHECK(sval.alloc(ctx, n));
  k_gather_f64<<<grid, 256, 0, ctx->stream>>>(di.Current(), val.p, n, sval.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CHECK(head.alloc(ctx, n));
  DDF_CHECK(rank.alloc(ctx, n + 1));
  k_key_heads<<<grid, 256, 0, ctx->stream>>>(dk.Current(), n, head.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, head.p, rank.p, n, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, head.p, rank.p, n, ctx->stream));
  ctx->launches++;
  uint32_t last_rank = 0;
  uint8_t last_head = 0;
  DDF_CUDA(cudaMemcpyAsync(&last_rank, rank.p + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaMemcpyAsync(&last_head, head.p + n - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t nout = (int64_t)last_rank + (last_head ? 1 : 0);
  DDF_CHECK(coo_alloc(ctx, out, nout));
  k_fold_runs<<<grid, 256, 0, ctx->stream>>>(dk.Current(), sval.p, head.p, rank.p, n, out.row.p, out.col.p, out.val.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// column pointers of a CSC-ordered Coo
static int coo_colptr(ddf_ctx* ctx, const Coo& c, int64_t ncols, DevBuf<uint32_t>& cp) {
  DevBuf<uint32_t> keys;
  DDF_CHECK(cp.alloc(ctx, ncols + 1));
  int grid;
  if (c.nnz) {
    DDF_CHECK(keys.alloc(ctx, c.nnz));
    DDF_CHECK(grid_for(c.nnz, 256, &grid));
    k_i32_to_u32<<<grid, 256, 0, ctx->stream>>>(c.col.p, keys.p, c.nnz);
    DDF_LAUNCH_CHECK(ctx);
  }
  DDF_CHECK(grid_for(ncols + 1, 256, &grid));
  k_lower_bound<<<grid, 256, 0, ctx->stream>>>(keys.p, c.nnz, ncols, cp.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// C = A * B, both CSC-ordered; A is nrowsA x inner, B is inner x ncolsB
static int coo_spgemm(ddf_ctx* ctx, const Coo& A, int64_t nrowsA, int64_t inner, const Coo& B, int64_t ncolsB, Coo& C) {
  (void)nrowsA;
  if (A.nnz == 0 || B.nnz == 0) return coo_alloc(ctx, C, 0);
  DevBuf<uint32_t> acp, cnt, off;
  DevBuf<uint8_t> temp;
  DDF_CHECK(coo_colptr(ctx, A, inner, acp));
  DDF_CHECK(cnt.alloc(ctx, B.nnz + 1));
  DDF_CHECK(off.alloc(ctx, B.nnz + 1));
  int grid;
  DDF_CHECK(grid_for(B.nnz + 1, 256, &grid));
  k_count_products<<<grid, 256, 0, ctx->stream>>>(B.row.p, B.nnz, acp.p, cnt.p);
  DDF_LAUNCH_CHECK(ctx);
  size_t tb = 0;
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, off.p, B.nnz + 1, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, cnt.p, off.p, B.nnz + 1, ctx->stream));
  ctx->launches++;
  uint32_t total = 0;
  DDF_CUDA(cudaMemcpyAsync(&total, off.p + B.nnz, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (total == 0) return coo_alloc(ctx, C, 0);
  DevBuf<unsigned long long> key;
  DevBuf<double> val;
  DDF_CHECK(key.alloc(ctx, total));
  DDF_CHECK(val.alloc(ctx, total));
  DDF_CHECK(grid_for(B.nnz, 256, &grid));
  k_expand_products<<<grid, 256, 0, ctx->stream>>>(B.row.p, B.col.p, B.val.p, B.nnz, acp.p, A.row.p, A.val.p, off.p, key.p,
                                                   val.p);
  DDF_LAUNCH_CHECK(ctx);
  return coo_from_keyed(ctx, key, val, total, ncolsB, C);
}

// C = A + B over the union pattern (A's value first in every sum)
static int coo_add(ddf_ctx* ctx, const Coo& A, const Coo& B, int64_t nrows, int64_t ncols, Coo& C) {
  (void)nrows;
  const int64_t n = A.nnz + B.nnz;
  if (n == 0) return coo_alloc(ctx, C, 0);
  DevBuf<unsigned long long> key;
  DevBuf<double> val;
  DDF_CHECK(key.alloc(ctx, n));
  DDF_CHECK(val.alloc(ctx, n));
  int grid;
  if (A.nnz) {
    DDF_CHECK(grid_for(A.nnz, 256, &grid));
    k_pack_keys<<<grid, 256, 0, ctx->stream>>>(A.row.p, A.col.p, A.nnz, key.p);
    DDF_LAUNCH_CHECK(ctx);
    DDF_CUDA(cudaMemcpyAsync(val.p, A.val.p, A.nnz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (B.nnz) {
    DDF_CHECK(grid_for(B.nnz, 256, &grid));
    k_pack_keys<<<grid, 256, 0, ctx->stream>>>(B.row.p, B.col.p, B.nnz, key.p + A.nnz);
    DDF_LAUNCH_CHECK(ctx);
    DDF_CUDA(cudaMemcpyAsync(val.p + A.nnz, B.val.p, B.nnz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return coo_from_keyed(ctx, key, val, n, ncols, C);
}

// in-place transpose of a CSC-ordered Coo (result CSC-ordered again); nrows = rows of the input
static int coo_transpose(ddf_ctx* ctx, Coo& c, int64_t nrows) {
  c.row.swap(c.col);                     // entries (j, i); still sorted by the OLD column = new row
  if (c.nnz == 0) return 0;
  DevBuf<unsigned long long> key;
  DevBuf<double> val;
  DDF_CHECK(key.alloc(ctx, c.nnz));
  DDF_CHECK(val.alloc(ctx, c.nnz));
  int grid;
  DDF_CHECK(grid_for(c.nnz, 256, &grid));
  k_pack_keys<<<grid, 256, 0, ctx->stream>>>(c.row.p, c.col.p, c.nnz, key.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaMemcpyAsync(val.p, c.val.p, c.nnz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  Coo out;
  const bool d = c.is_diag;
  DDF_CHECK(coo_from_keyed(ctx, key, val, c.nnz, nrows, out));
  coo_move(c, out);
  c.is_diag = d;
  return 0;
}

int ddf_laplace0_to_coo(ddf_mesh* m, DevBuf<int32_t>& row, DevBuf<int32_t>& col, DevBuf<double>& val, int64_t* nnz);  // wave.cu

// materialise node -> Coo in CSC order, chopped
static int materialise(OpNode* n, Coo& c) {
  ddf_ctx* ctx = n->ctx;
  ddf_mesh* m = n->mesh;
  int grid;
  if (n->adj_of) {                      // adjoint(A).values = transpose of the assembled A
    DDF_CHECK(materialise(n->adj_of.get(), c));
    return coo_transpose(ctx, c, n->adj_of->nrows);
  }
  // pattern [DIAG] sparse [DIAG] (also a bare sparse leaf / bare diagonal)
  const OpNode *dl = nullptr, *dr = nullptr, *core = nullptr;
  if (n->kind == OPK_FIXED || n->kind == OPK_CSRSIGN) core = n;
  else if (n->kind == OPK_CHAIN && n->scale == 1.0) {
    std::vector<OpNode*> f;
    for (auto& g : n->factors) f.push_back(g.get());
    size_t a = 0, b = f.size();
    if (a < b && f[a]->kind == OPK_DIAG) dl = f[a++];
    if (a < b && f[b - 1]->kind == OPK_DIAG && b - 1 > a) dr = f[--b];
    if (b - a == 1 && (f[a]->kind == OPK_FIXED || f[a]->kind == OPK_CSRSIGN)) core = f[a];
    else { dl = dr = nullptr; }
  }
  if (core) {
    const int R = core->R;
    const int64_t nnz = m->n[R] * (R + 1);
    DDF_CHECK(coo_alloc(ctx, c, nnz));
    if (nnz) {
      const double* pl = dl ? dl->diag : nullptr; const double sl = dl ? dl->scale : 1.0;
      const double* pr = dr ? dr->diag : nullptr; const double sr = dr ? dr->scale : 1.0;
      if (core->kind == OPK_FIXED) {
        DDF_CHECK(grid_for(m->n[R - 1], 256, &grid));
        k_coo_from_rowcsr<<<grid, 256, 0, ctx->stream>>>(m->brow_ptr[R].p, m->bcol[R].p, m->n[R - 1], pl, sl, pr, sr,
                                                         c.row.p, c.col.p, c.val.p, 1);
      } else {
        DDF_CHECK(grid_for(m->n[R], 256, &grid));
        k_coo_from_table<<<grid, 256, 0, ctx->stream>>>(m->bnd_table(R), m->n[R], R + 1, R & 1, pl, sl, pr, sr, c.row.p,
                                                        c.col.p, c.val.p);
      }
      DDF_LAUNCH_CHECK(ctx);
    }
    // NB: the reference chops after each product; with |entries| = |dl*dr| a second chop
    // after the first product can only differ when |dl| alone is below threshold -- apply both.
    return coo_chop(ctx, c);
  }
  if (n->kind == OPK_DIAG) {
    DDF_CHECK(coo_alloc(ctx, c, n->nrows));
    if (n->nrows) {
      DDF_CHECK(grid_for(n->nrows, 256, &grid));
      k_coo_diag<<<grid, 256, 0, ctx->stream>>>(n->diag, n->scale, n->nrows, c.row.p, c.col.p, c.val.p);
      DDF_LAUNCH_CHECK(ctx);
    }
    c.is_diag = true;
    return 0;     // Diagonal is not an AbstractSparseMatrix: dnz leaves it alone (Ops.jl:40)
  }
  if (n->kind == OPK_ZERO) return coo_alloc(ctx, c, 0);
  if (n->kind == OPK_CSRF64) {
    DDF_CHECK(coo_alloc(ctx, c, n->nnz));
    if (n->nnz) {
      DDF_CHECK(grid_for(n->nrows, 256, &grid));
      k_csr_to_coo<<<grid, 256, 0, ctx->stream>>>(n->rowptr.p, n->nrows, c.row.p);
      DDF_LAUNCH_CHECK(ctx);
      DDF_CUDA(cudaMemcpyAsync(c.col.p, n->col.p, n->nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      DDF_CUDA(cudaMemcpyAsync(c.val.p, n->val.p, n->nnz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      DDF_CHECK(coo_sort_by_col(ctx, c, n->ncols));
    }
    return coo_chop(ctx, c);
  }
  if (n->kind == OPK_LAPLACE0) {
    DDF_CHECK(ddf_laplace0_to_coo(m, c.row, c.col, c.val, &c.nnz));      // row order
    DDF_CHECK(coo_sort_by_col(ctx, c, n->ncols));                          // stable: rows stay ascending inside a column
    return coo_chop(ctx, c);
  }
  if (n->kind == OPK_CHAIN && !n->tree_ops.empty()) {
    // ((o1 * o2) * o3) ... : SparseArrays spmatmul + dnz/chop after every product (Ops.jl:228-232)
    DDF_CHECK(materialise(n->tree_ops[0].get(), c));
    int64_t nrows = n->tree_ops[0]->nrows;
    for (size_t t = 1; t < n->tree_ops.size(); t++) {
      Coo b, prod;
      DDF_CHECK(materialise(n->tree_ops[t].get(), b));
      DDF_CHECK(coo_spgemm(ctx, c, nrows, n->tree_ops[t]->nrows, b, n->tree_ops[t]->ncols, prod));
      prod.is_diag = c.is_diag && b.is_diag;
      if (!prod.is_diag) DDF_CHECK(coo_chop(ctx, prod));
      coo_move(c, prod);
    }
    if (n->tree_scale != 1.0) {
      DDF_CHECK(coo_scale(ctx, c, n->tree_scale));
      if (!c.is_diag) DDF_CHECK(coo_chop(ctx, c));
    }
    return 0;
  }
  if (n->kind == OPK_SUM && !n->tree_ops.empty()) {
    // A + b*B: map(+, A, B) over the union pattern, then dnz/chop (Ops.jl:186-196)
    DDF_CHECK(materialise(n->tree_ops[0].get(), c));
    DDF_CHECK(coo_scale(ctx, c, n->tree_coefs[0]));
    for (size_t t = 1; t < n->tree_ops.size(); t++) {
      Coo b, sum;
      DDF_CHECK(materialise(n->tree_ops[t].get(), b));
      DDF_CHECK(coo_scale(ctx, b, n->tree_coefs[t]));
      DDF_CHECK(coo_add(ctx, c, b, n->nrows, n->ncols, sum));
      sum.is_diag = c.is_diag && b.is_diag;
      if (!sum.is_diag) DDF_CHECK(coo_chop(ctx, sum));
      coo_move(c, sum);
    }
    return 0;
  }
  ddf_set_error("to_csc: this operator has no assembled form (kind %d)", (int)n->kind);
  return 1;
}

// The reference stores EVERY Op as its assembled sparse matrix (Ops.jl:25) and `A * f` is SparseArrays' mul! on it
// (Ops.jl:262-266): this is that representation on the device -- the tree is materialised with the reference's
// roundings and chops, then held as rows (ascending columns) so that an apply is one kernel.
static int rows_from_coo(ddf_ctx* ctx, OpNode* n, Coo& c, ddf_op** out) {
  DDF_CHECK(coo_transpose(ctx, c, n->nrows));   // sorted by (row, column): c.col = row, c.row = column
  OpRef o = std::make_shared<OpNode>();
  o->kind = OPK_CSRF64; o->ctx = ctx; o->mesh = n->mesh;
  o->nrows = n->nrows; o->ncols = n->ncols; o->nnz = c.nnz;
  DDF_CHECK(o->rowptr.alloc(ctx, o->nrows + 1));
  DDF_CHECK(o->col.alloc(ctx, c.nnz));
  DDF_CHECK(o->val.alloc(ctx, c.nnz));
  int grid;
  DevBuf<uint32_t> keys;
  if (c.nnz) {
    DDF_CHECK(keys.alloc(ctx, c.nnz));
    DDF_CHECK(grid_for(c.nnz, 256, &grid));
    k_i32_to_u32<<<grid, 256, 0, ctx->stream>>>(c.col.p, keys.p, c.nnz);
    DDF_LAUNCH_CHECK(ctx);
    DDF_CUDA(cudaMemcpyAsync(o->col.p, c.row.p, c.nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    DDF_CUDA(cudaMemcpyAsync(o->val.p, c.val.p, c.nnz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  DDF_CHECK(grid_for(o->nrows + 1, 256, &grid));
  k_lower_bound<<<grid, 256, 0, ctx->stream>>>(keys.p, c.nnz, o->nrows, o->rowptr.p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = wrap(o);
  return 0;
}

extern "C" int ddf_op_assemble(ddf_op* A, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpNode* n = A->node.get();
  Coo c;
  DDF_CHECK(materialise(n, c));
  return rows_from_coo(n->ctx, n, c, out);
  DDF_TRY_END
}

__global__ void k_div_vals(double* __restrict__ v, int64_t n, double a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = __ddiv_rn(v[i], a);
}
__global__ void k_diag_div(const double* __restrict__ d, double s, double a, double* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __ddiv_rn(d ? __dmul_rn(s, d[i]) : s, a);
}

// A / a (Ops.jl:210-212): `Op(A.manifold, A.values / a)` -- every stored value DIVIDED by a (one rounding; (1/a) * A
// rounds twice unless 1/a is exact), then the constructor's dnz/chop.  A power-of-two divisor stays lazy (exact either
// way); anything else is held like the reference holds it, as the assembled matrix.
extern "C" int ddf_op_div(ddf_op* A, double a, ddf_op** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  OpNode* n = A->node.get();
  ddf_ctx* ctx = n->ctx;
  int e = 0;
  const double mant = frexp(a, &e);
  if ((mant == 0.5 || mant == -0.5) && e > -1000 && e < 1000) return ddf_op_scale(1.0 / a, A, out);
  if (n->kind == OPK_ZERO) { *out = wrap(A->node); return 0; }
  int grid;
  if (n->kind == OPK_DIAG) {           // Diagonal(d) / a = Diagonal(d ./ a)
    OpRef o = std::make_shared<OpNode>();
    o->kind = OPK_DIAG; o->ctx = ctx; o->mesh = n->mesh; o->nrows = n->nrows; o->ncols = n->ncols;
    DDF_CHECK(o->diag_own.alloc(ctx, n->nrows));
    if (n->nrows) {
      DDF_CHECK(grid_for(n->nrows, 256, &grid));
      k_diag_div<<<grid, 256, 0, ctx->stream>>>(n->diag, n->scale, a, o->diag_own.p, n->nrows);
      DDF_LAUNCH_CHECK(ctx);
    }
    o->diag = o->diag_own.p;
    o->factors.push_back(A->node);
    *out = wrap(o);
    return 0;
  }
  Coo c;
  DDF_CHECK(materialise(n, c));
  if (c.nnz) {
    DDF_CHECK(grid_for(c.nnz, 256, &grid));
    k_div_vals<<<grid, 256, 0, ctx->stream>>>(c.val.p, c.nnz, a);
    DDF_LAUNCH_CHECK(ctx);
  }
  if (!c.is_diag) DDF_CHECK(coo_chop(ctx, c));
  return rows_from_coo(ctx, n, c, out);
  DDF_TRY_END
}

extern "C" int ddf_mesh_deriv_format(ddf_mesh* m, int k, int* format) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_assembled(m));
  if (k < 0 || k >= m->D) DDF_FAIL("deriv_format: need 0 <= k < D");
  const int R = k + 1, W = R + 1;
  (void)W;
  DDF_CHECK(dk_rows_format(m, R, format));
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_op_to_csc(ddf_op* A, int64_t* nnz, int64_t* colptr, int64_t* rowval, double* nzval) {
  DDF_TRY_BEGIN
  OpNode* n = A->node.get();
  ddf_ctx* ctx = n->ctx;
  Coo c;
  DDF_CHECK(materialise(n, c));
  *nnz = c.nnz;
  if (!colptr) return 0;
  // colptr by lower bound over the sorted column keys
  DevBuf<uint32_t> keys, cp;
  DDF_CHECK(cp.alloc(ctx, n->ncols + 1));
  int grid;
  if (c.nnz) {
    DDF_CHECK(keys.alloc(ctx, c.nnz));
    DDF_CHECK(grid_for(c.nnz, 256, &grid));
    k_i32_to_u32<<<grid, 256, 0, ctx->stream>>>(c.col.p, keys.p, c.nnz);
    DDF_LAUNCH_CHECK(ctx);
  }
  DDF_CHECK(grid_for(n->ncols + 1, 256, &grid));
  k_lower_bound<<<grid, 256, 0, ctx->stream>>>(keys.p, c.nnz, n->ncols, cp.p);
  DDF_LAUNCH_CHECK(ctx);
  std::vector<uint32_t> hcp(n->ncols + 1);
  std::vector<int32_t> hrow(c.nnz);
  DDF_CUDA(cudaMemcpyAsync(hcp.data(), cp.p, (n->ncols + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (c.nnz) {
    DDF_CUDA(cudaMemcpyAsync(hrow.data(), c.row.p, c.nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DDF_CUDA(cudaMemcpyAsync(nzval, c.val.p, c.nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int64_t j = 0; j <= n->ncols; j++) colptr[j] = (int64_t)hcp[j] + 1;
  for (int64_t p = 0; p < c.nnz; p++) rowval[p] = (int64_t)hrow[p] + 1;
  return 0;
  DDF_TRY_END
}

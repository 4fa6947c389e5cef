// wave.cu -- K6: the vertex Laplacian L = laplace(Val(Pr),Val(0)) = delta_1 d_0
// (src/ManifoldOps.jl:128-144 with :110-121) and the scalar wave step built on it
// (examples/wave.jl:82-100).
//
// The reference assembles L by two diagonal scalings and one SpGEMM
// (Ops.jl:228-232), chopping after every product (Ops.jl:44-58):
//     delta[i,e] = fl((-ih0[i] * s_ie) * star1[e]),   L[i,j] = delta[i,e] * d0[e,j],
//     L[i,i]    = sequential sum over incident edges e ascending of delta[i,e]*d0[e,i].
// k_build_laplace0 reproduces exactly those roundings into a row-CSR with the
// diagonal split off.  Because edge ids are lexicographic ranks of (a,b), the
// incident edges of vertex i in id order are its neighbours in ascending vertex
// order -- the reference's CSC `mul!` (zero y; scatter by ascending column) is
// therefore the per-row sum: incoming neighbours, diagonal, outgoing neighbours,
// with separately rounded multiply and add.  All kernels below use that order, so
// L*u, the wave right-hand side and the leapfrog step are bit-exact with the oracle.
#include "mesh.cuh"

int ddf_mesh_require_star(ddf_mesh* m, int R);   // ops.cu
int ddf_reduce(ddf_ctx* c, int mode, const double* x, const double* y, int64_t n, double* out);   // core.cu

__global__ void k_build_laplace0(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colw,
                                 const int32_t* __restrict__ edges, const double* __restrict__ ih0,
                                 const double* __restrict__ star1, int64_t nv, int32_t* __restrict__ nbr,
                                 double* __restrict__ lval, double* __restrict__ ldiag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  const double ih = ih0[i];
  double diag = 0.0;
  bool first = true;
  const bool ih_kept = !(ih * ih < DDF_CHOP_ABS2);   // chop of (bitsign * invhodge) * deriv  entries (+-ih)
  for (uint32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
    const uint32_t cw = colw[p];
    const uint32_t e = cw & 0x7fffffffu;
    const bool i_is_first = (cw >> 31) != 0;        // d_1[a,e] = -1
    nbr[p] = i_is_first ? edges[2 * (size_t)e + 1] : edges[2 * (size_t)e];
    double l = ih_kept ? __dmul_rn(ih, star1[e]) : 0.0;   // |delta[i,e]|
    if (l * l < DDF_CHOP_ABS2) l = 0.0;                   // chop(delta), then chop(L) on the same magnitude
    lval[p] = l;                                           // L[i,j] = +l for both orientations
    if (l != 0.0) {                                        // diagonal: first touch assigns, later touches add
      diag = first ? -l : __dadd_rn(diag, -l);
      first = false;
    }
  }
  if (diag * diag < DDF_CHOP_ABS2) diag = 0.0;
  ldiag[i] = diag;
}

int ddf_mesh_require_wave(ddf_mesh* m) {
  if (m->has_wave) return 0;
  if (m->D < 1) DDF_FAIL("laplace(Pr,0): needs D >= 1");
  DDF_CHECK(ddf_mesh_require_star(m, 0));
  DDF_CHECK(ddf_mesh_require_star(m, 1));
  DDF_CHECK(ddf_mesh_require_isboundary(m));
  ddf_ctx* ctx = m->ctx;
  const int64_t nv = m->n[0], ne = m->n[1];
  DDF_CHECK(m->adj_nbr.alloc(ctx, 2 * ne));
  DDF_CHECK(m->adj_l.alloc(ctx, 2 * ne));
  DDF_CHECK(m->ldiag.alloc(ctx, nv));
  if (nv) {
    int grid;
    DDF_CHECK(grid_for(nv, 256, &grid));
    k_build_laplace0<<<grid, 256, 0, ctx->stream>>>(m->brow_ptr[1].p, m->bcol[1].p, m->simp[1].p, m->istar[0].p,
                                                    m->star[1].p, nv, m->adj_nbr.p, m->adj_l.p, m->ldiag.p);
    DDF_LAUNCH_CHECK(ctx);
  }
  m->has_wave = true;
  return 0;
}

// (L u)_i in the reference's summation order
__device__ __forceinline__ double row_laplace(int64_t i, const uint32_t* __restrict__ rowptr,
                                              const int32_t* __restrict__ nbr, const double* __restrict__ lval,
                                              const double* __restrict__ ldiag, const double* __restrict__ u,
                                              double ui) {
  const uint32_t lo = rowptr[i], hi = rowptr[i + 1];
  const double d = ldiag[i];
  double acc = 0.0;
  bool diag_done = (d == 0.0);
  for (uint32_t p = lo; p < hi; p++) {
    const int32_t j = __ldg(nbr + p);
    const double l = __ldg(lval + p);
    if (!diag_done && j > i) {
      acc = __dadd_rn(acc, __dmul_rn(d, ui));
      diag_done = true;
    }
    if (l != 0.0) acc = __dadd_rn(acc, __dmul_rn(l, __ldg(u + j)));
  }
  if (!diag_done) acc = __dadd_rn(acc, __dmul_rn(d, ui));
  return acc;
}

__global__ void __launch_bounds__(256)
k_laplace0(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr, const double* __restrict__ lval,
           const double* __restrict__ ldiag, const double* __restrict__ u, double* __restrict__ y, int64_t nv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  y[i] = row_laplace(i, rowptr, nbr, lval, ldiag, u, u[i]);
}

int ddf_laplace0_apply(ddf_mesh* m, const double* x, double* y) {
  DDF_CHECK(ddf_mesh_require_wave(m));
  if (!m->n[0]) return 0;
  int grid;
  DDF_CHECK(grid_for(m->n[0], 256, &grid));
  k_laplace0<<<grid, 256, 0, m->ctx->stream>>>(m->brow_ptr[1].p, m->adj_nbr.p, m->adj_l.p, m->ldiag.p, x, y, m->n[0]);
  DDF_LAUNCH_CHECK(m->ctx);
  return 0;
}

// assembled L as COO in CSC order (column j: incoming rows, j, outgoing rows)
__global__ void k_laplace0_coo(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                               const double* __restrict__ lval, const double* __restrict__ ldiag,
                               const uint32_t* __restrict__ colw, int64_t nv, int32_t* __restrict__ row,
                               int32_t* __restrict__ col, double* __restrict__ val) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nv) return;
  // L[i,j] for the neighbours i of j equals fl(ih0[i]*star1[e]) = the value stored in ROW i for edge e;
  // find it through the edge: row i stores it at the position of e in i's incidence list.
  size_t out = (size_t)rowptr[j] + (size_t)j;
  bool diag_done = false;
  for (uint32_t p = rowptr[j]; p < rowptr[j + 1]; p++) {
    const int32_t i = nbr[p];
    if (!diag_done && i > j) {
      row[out] = (int32_t)j; col[out] = (int32_t)j; val[out] = ldiag[j]; out++;
      diag_done = true;
    }
    const uint32_t e = colw[p] & 0x7fffffffu;
    double v = 0.0;
    for (uint32_t q = rowptr[i]; q < rowptr[i + 1]; q++)
      if ((colw[q] & 0x7fffffffu) == e) { v = lval[q]; break; }
    row[out] = i; col[out] = (int32_t)j; val[out] = v; out++;
  }
  if (!diag_done) { row[out] = (int32_t)j; col[out] = (int32_t)j; val[out] = ldiag[j]; }
}

int ddf_laplace0_to_coo(ddf_mesh* m, DevBuf<int32_t>& row, DevBuf<int32_t>& col, DevBuf<double>& val, int64_t* nnz) {
  DDF_CHECK(ddf_mesh_require_wave(m));
  ddf_ctx* ctx = m->ctx;
  const int64_t n = 2 * m->n[1] + m->n[0];
  *nnz = n;
  DDF_CHECK(row.alloc(ctx, n));
  DDF_CHECK(col.alloc(ctx, n));
  DDF_CHECK(val.alloc(ctx, n));
  if (!m->n[0]) return 0;
  int grid;
  DDF_CHECK(grid_for(m->n[0], 256, &grid));
  k_laplace0_coo<<<grid, 256, 0, ctx->stream>>>(m->brow_ptr[1].p, m->adj_nbr.p, m->adj_l.p, m->ldiag.p, m->bcol[1].p,
                                                m->n[0], row.p, col.p, val.p);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

// examples/wave.jl:87-100: du = (E-B) udot ; dudot = (E-B) (L u)
__global__ void __launch_bounds__(256)
k_wave_rhs(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr, const double* __restrict__ lval,
           const double* __restrict__ ldiag, const uint8_t* __restrict__ isb, const double* __restrict__ u,
           const double* __restrict__ udot, double* __restrict__ du, double* __restrict__ dudot, int64_t nv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  const double m = isb[i] ? 0.0 : 1.0;
  const double lu = row_laplace(i, rowptr, nbr, lval, ldiag, u, u[i]);
  du[i] = __dmul_rn(m, udot[i]);
  dudot[i] = __dmul_rn(m, lu);
}

extern "C" int ddf_wave_rhs(ddf_mesh* m, ddf_vec* u, ddf_vec* udot, ddf_vec* du, ddf_vec* dudot) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_wave(m));
  const int64_t nv = m->n[0];
  if (u->n != nv || udot->n != nv || du->n != nv || dudot->n != nv) DDF_FAIL("wave_rhs: vectors must have length n0=%lld", (long long)nv);
  if (du == u || dudot == u) DDF_FAIL("wave_rhs: outputs must not alias u");
  if (!nv) return 0;
  int grid;
  DDF_CHECK(grid_for(nv, 256, &grid));
  k_wave_rhs<<<grid, 256, 0, m->ctx->stream>>>(m->brow_ptr[1].p, m->adj_nbr.p, m->adj_l.p, m->ldiag.p, m->isbnd[0].p,
                                               u->d.p, udot->d.p, du->d.p, dudot->d.p, nv);
  DDF_LAUNCH_CHECK(m->ctx);
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_wave_dt(ddf_mesh* m, double* dt) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_geometry(m));
  double mn = 0;
  DDF_CHECK(ddf_reduce(m->ctx, 5, m->vol[1].p, nullptr, m->n[1], &mn));
  *dt = mn / 2;                                    // dt = minimum(vol_1) / 2  (wave.jl:106,118)
  return 0;
  DDF_TRY_END
}

// Fused leapfrog step (north_star; not in the reference, which uses Tsit5):
//   v' = v + dt * ((1-b) .* (L u));   u' = u + dt * ((1-b) .* v')
// One pass: u is read (own value + gathered neighbours), u' goes to the ping-pong
// buffer so neighbours still see the old u.
__global__ void __launch_bounds__(256)
k_leapfrog(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr, const double* __restrict__ lval,
           const double* __restrict__ ldiag, const uint8_t* __restrict__ isb, const double* __restrict__ u,
           double* __restrict__ v, double* __restrict__ unew, double dt, int64_t row0, int64_t row1) {
  const int64_t i = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row1) return;
  const double ui = u[i];
  const double m = isb[i] ? 0.0 : 1.0;
  const double lu = row_laplace(i, rowptr, nbr, lval, ldiag, u, ui);
  const double vn = __dadd_rn(v[i], __dmul_rn(dt, __dmul_rn(m, lu)));
  v[i] = vn;
  unew[i] = __dadd_rn(ui, __dmul_rn(dt, __dmul_rn(m, vn)));
}

// rows [row0, row1) of one leapfrog step on `stream`
int ddf_leapfrog_launch(ddf_mesh* m, cudaStream_t stream, const double* u, double* v, double* unew, double dt,
                        int64_t row0, int64_t row1) {
  if (row1 <= row0) return 0;
  int grid;
  DDF_CHECK(grid_for(row1 - row0, 256, &grid));
  k_leapfrog<<<grid, 256, 0, stream>>>(m->brow_ptr[1].p, m->adj_nbr.p, m->adj_l.p, m->ldiag.p, m->isbnd[0].p, u, v,
                                       unew, dt, row0, row1);
  DDF_LAUNCH_CHECK(m->ctx);
  return 0;
}

extern "C" int ddf_wave_leapfrog(ddf_mesh* m, ddf_vec* u, ddf_vec* v, double dt, int nsteps) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_wave(m));
  const int64_t nv = m->n[0];
  if (u->n != nv || v->n != nv) DDF_FAIL("wave_leapfrog: vectors must have length n0=%lld", (long long)nv);
  if (u == v) DDF_FAIL("wave_leapfrog: u and v must not alias");
  if (!nv || nsteps <= 0) return 0;
  if ((int64_t)m->utmp.n != nv) DDF_CHECK(m->utmp.alloc(m->ctx, nv));
  for (int s = 0; s < nsteps; s++) {
    DDF_CHECK(ddf_leapfrog_launch(m, m->ctx->stream, u->d.p, v->d.p, m->utmp.p, dt, 0, nv));
    u->d.swap(m->utmp);          // O(1) ping-pong: the vector handle now owns the new state
  }
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_wave_step_bytes(ddf_mesh* m, int64_t* algorithmic, int64_t* format_bytes) {
  DDF_CHECK(ddf_mesh_require_assembled(m));
  const int64_t E = m->n[1], V = m->n[0];
  *algorithmic = E * 16 + V * 41;                  // E*(2*4+8) + V*(8*5+1)
  *format_bytes = 2 * E * 12 + V * (4 + 8 + 8 * 4 + 1);   // what this format streams per step
  return 0;
}

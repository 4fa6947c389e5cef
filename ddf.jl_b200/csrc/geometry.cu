// geometry.cu -- K5: per-simplex geometry feeding the diagonal Hodge stars.
//   volumes            src/Manifolds.jl:834-855 -> src/Algorithms.jl:102-147
//   weighted circumcentres   src/Manifolds.jl:896-913 -> src/Algorithms.jl:56-96
//   circumcentric dual volumes  src/Manifolds.jl:1160-1223 (driver :644-690)
//   boundary masks     src/Manifolds.jl:510-550
// One thread per simplex, Float64 throughout; these run once per mesh and are
// tiny next to the operator applications.
#include "mesh.cuh"

template <int C>
struct VecC {
  double v[C > 0 ? C : 1];
};

template <int C>
__device__ __forceinline__ VecC<C> load_pt(const double* __restrict__ coords, int32_t i) {
  VecC<C> r;
#pragma unroll
  for (int c = 0; c < C; c++) r.v[c] = coords[(size_t)i * C + c];
  return r;
}
template <int C>
__device__ __forceinline__ double dotC(const VecC<C>& a, const VecC<C>& b) {
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; c++) s += a.v[c] * b.v[c];
  return s;
}
template <int C>
__device__ __forceinline__ VecC<C> subC(const VecC<C>& a, const VecC<C>& b) {
  VecC<C> r;
#pragma unroll
  for (int c = 0; c < C; c++) r.v[c] = a.v[c] - b.v[c];
  return r;
}

// |wedge(y_1..y_R)| in C dims = sqrt(sum of squared RxR minors)
template <int R, int C>
__device__ __forceinline__ double wedge_norm(const VecC<C>* y) {
  if (R == 1) return sqrt(dotC<C>(y[0], y[0]));
  if (R == 2) {
    if (C == 2) return fabs(y[0].v[0] * y[1].v[1] - y[0].v[1] * y[1].v[0]);
    double cx = y[0].v[1 % C] * y[1].v[2 % C] - y[0].v[2 % C] * y[1].v[1 % C];
    double cy = y[0].v[2 % C] * y[1].v[0] - y[0].v[0] * y[1].v[2 % C];
    double cz = y[0].v[0] * y[1].v[1 % C] - y[0].v[1 % C] * y[1].v[0];
    return sqrt(cx * cx + cy * cy + cz * cz);
  }
  // R == 3, C == 3
  double det = y[0].v[0] * (y[1].v[1 % C] * y[2 % R].v[2 % C] - y[1].v[2 % C] * y[2 % R].v[1 % C]) -
               y[0].v[1 % C] * (y[1].v[0] * y[2 % R].v[2 % C] - y[1].v[2 % C] * y[2 % R].v[0]) +
               y[0].v[2 % C] * (y[1].v[0] * y[2 % R].v[1 % C] - y[1].v[1 % C] * y[2 % R].v[0]);
  return fabs(det);
}

template <int R, int C>
__global__ void k_volumes(const int32_t* __restrict__ simp, const double* __restrict__ coords, int64_t n,
                          double* __restrict__ vol) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  VecC<C> x0 = load_pt<C>(coords, simp[i * (R + 1)]);
  VecC<C> y[R];
#pragma unroll
  for (int k = 0; k < R; k++) y[k] = subC<C>(load_pt<C>(coords, simp[i * (R + 1) + k + 1]), x0);
  double f = 1.0;
#pragma unroll
  for (int k = 2; k <= R; k++) f *= k;
  vol[i] = wedge_norm<R, C>(y) / f;            // vol /= factorial(N-1)  (Algorithms.jl:116)
}

// Weighted circumcentre: solve the (N+1)x(N+1) system of Algorithms.jl:88-94 by
// Gaussian elimination with partial pivoting (StaticArrays `\` is an LU solve).
template <int R, int C>
__global__ void k_circumcentres(const int32_t* __restrict__ simp, const double* __restrict__ coords,
                                const double* __restrict__ weights, int weighted, int64_t n,
                                double* __restrict__ cc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int N = R + 1, M = N + 1;
  VecC<C> x[N];
  double w[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    int32_t v = simp[i * N + k];
    x[k] = load_pt<C>(coords, v);
    w[k] = weighted ? weights[v] : 0.0;
  }
  double A[M][M + 1];
#pragma unroll
  for (int a = 0; a < N; a++) {
#pragma unroll
    for (int b = 0; b < N; b++) A[a][b] = 2 * dotC<C>(x[a], x[b]);
    A[a][N] = 1.0;
    A[a][M] = dotC<C>(x[a], x[a]) - w[a];
  }
#pragma unroll
  for (int b = 0; b < N; b++) A[N][b] = 1.0;
  A[N][N] = 0.0;
  A[N][M] = 1.0;
#pragma unroll
  for (int col = 0; col < M; col++) {
    int piv = col;
    double best = fabs(A[col][col]);
#pragma unroll
    for (int r = col + 1; r < M; r++) {
      double a = fabs(A[r][col]);
      if (a > best) { best = a; piv = r; }
    }
#pragma unroll
    for (int r = col + 1; r < M; r++) {
      if (r == piv) {
#pragma unroll
        for (int c2 = 0; c2 <= M; c2++) { double t = A[col][c2]; A[col][c2] = A[r][c2]; A[r][c2] = t; }
      }
    }
    double inv = 1.0 / A[col][col];
#pragma unroll
    for (int r = col + 1; r < M; r++) {
      double f = A[r][col] * inv;
#pragma unroll
      for (int c2 = col; c2 <= M; c2++) A[r][c2] -= f * A[col][c2];
    }
  }
  double sol[M];
#pragma unroll
  for (int r = M - 1; r >= 0; r--) {
    double s = A[r][M];
#pragma unroll
    for (int c2 = r + 1; c2 < M; c2++) s -= A[r][c2] * sol[c2];
    sol[r] = s / A[r][r];
  }
#pragma unroll
  for (int c = 0; c < C; c++) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < N; k++) s += sol[k] * x[k].v[c];
    cc[(size_t)i * C + c] = s;
  }
}

// Dual volumes of R-simplices from those of (R+1)-simplices (Manifolds.jl:1177-1221):
//   dualvol_R[i] = 1/(D-R) * sum_{j in parents(i), ascending} dualvol_{R+1}[j] * h_ij
//   h_ij = (cc_j - cc_i) . nhat,  nhat = unit normal of i inside j pointing to the
//   barycentre of j (the reference's s = sign((bc_j - x_i1).nhat) folds into that).
template <int R, int C>
__global__ void k_dualvolumes(const int32_t* __restrict__ simpR /* n x (R+1); null for R=0 */,
                              const int32_t* __restrict__ simpR1 /* (R+1)-simplices, R+2 wide */,
                              const uint32_t* __restrict__ prow, const uint32_t* __restrict__ pcol,
                              const double* __restrict__ coords, const double* __restrict__ ccR,
                              const double* __restrict__ ccR1, const double* __restrict__ dv1, int D, int64_t n,
                              double* __restrict__ dv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  VecC<C> xi[R + 1];
#pragma unroll
  for (int k = 0; k <= R; k++) xi[k] = load_pt<C>(coords, R == 0 ? (int32_t)i : simpR[i * (R + 1) + k]);
  VecC<C> cci = load_pt<C>(ccR, (int32_t)i);
  VecC<C> y[R > 0 ? R : 1];
#pragma unroll
  for (int k = 0; k < R; k++) y[k] = subC<C>(xi[k + 1], xi[0]);
  double acc = 0.0;
  for (uint32_t p = prow[i]; p < prow[i + 1]; p++) {
    const uint32_t j = pcol[p] & 0x7fffffffu;
    VecC<C> bc;
#pragma unroll
    for (int c = 0; c < C; c++) bc.v[c] = 0;
#pragma unroll
    for (int k = 0; k <= R + 1; k++) {
      VecC<C> xj = load_pt<C>(coords, simpR1[(size_t)j * (R + 2) + k]);
#pragma unroll
      for (int c = 0; c < C; c++) bc.v[c] = k == 0 ? xj.v[c] : bc.v[c] + xj.v[c];
    }
#pragma unroll
    for (int c = 0; c < C; c++) bc.v[c] /= (R + 2);          // barycentre (Algorithms.jl:7)
    VecC<C> v = subC<C>(bc, xi[0]);
    if (R == 1) {
      double f = dotC<C>(v, y[0]) / dotC<C>(y[0], y[0]);
#pragma unroll
      for (int c = 0; c < C; c++) v.v[c] -= f * y[0].v[c];
    } else if (R == 2) {
      double g00 = dotC<C>(y[0], y[0]), g01 = dotC<C>(y[0], y[1 % (R > 0 ? R : 1)]),
             g11 = dotC<C>(y[1 % (R > 0 ? R : 1)], y[1 % (R > 0 ? R : 1)]);
      double r0 = dotC<C>(v, y[0]), r1 = dotC<C>(v, y[1 % (R > 0 ? R : 1)]);
      double det = g00 * g11 - g01 * g01;
      double a = (r0 * g11 - r1 * g01) / det, b = (g00 * r1 - g01 * r0) / det;
#pragma unroll
      for (int c = 0; c < C; c++) v.v[c] -= a * y[0].v[c] + b * y[1 % (R > 0 ? R : 1)].v[c];
    }
    double nv = sqrt(dotC<C>(v, v));
    VecC<C> ccj = load_pt<C>(ccR1, (int32_t)j);
    VecC<C> dcc = subC<C>(ccj, cci);
    double h = dotC<C>(dcc, v) / nv;
    acc += dv1[j] * h;                                        // voli += b * h  (:1214-1217)
  }
  dv[i] = acc / (double)(D - R);                              // :1220
}

__global__ void k_fill_d(double* __restrict__ p, int64_t n, double a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = a;
}

template <int R, int C>
static int geometry_level(ddf_mesh* m) {
  ddf_ctx* ctx = m->ctx;
  const int64_t n = m->n[R];
  DDF_CHECK(m->vol[R].alloc(ctx, n));
  DDF_CHECK(m->cc[R].alloc(ctx, (size_t)n * C));
  if (n == 0) return 0;
  int grid;
  DDF_CHECK(grid_for(n, 128, &grid));
  k_volumes<R, C><<<grid, 128, 0, ctx->stream>>>(m->simp[R].p, m->coords.p, n, m->vol[R].p);
  DDF_LAUNCH_CHECK(ctx);
  k_circumcentres<R, C><<<grid, 128, 0, ctx->stream>>>(m->simp[R].p, m->coords.p, m->weights.p, m->weighted ? 1 : 0, n,
                                                      m->cc[R].p);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

template <int R, int C>
static int dualvol_level(ddf_mesh* m) {
  ddf_ctx* ctx = m->ctx;
  const int64_t n = m->n[R];
  DDF_CHECK(m->dualvol[R].alloc(ctx, n));
  if (n == 0) return 0;
  int grid;
  DDF_CHECK(grid_for(n, 128, &grid));
  k_dualvolumes<R, C><<<grid, 128, 0, ctx->stream>>>(R == 0 ? nullptr : m->simp[R].p, m->simp[R + 1].p,
                                                    m->brow_ptr[R + 1].p, m->bcol[R + 1].p, m->coords.p,
                                                    R == 0 ? m->coords.p : m->cc[R].p, m->cc[R + 1].p,
                                                    m->dualvol[R + 1].p, m->D, n, m->dualvol[R].p);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

template <int C>
static int geometry_all(ddf_mesh* m) {
  ddf_ctx* ctx = m->ctx;
  const int D = m->D;
  // R = 0: vol = 1 (Manifolds.jl:838), dual coords = coords (:903)
  DDF_CHECK(m->vol[0].alloc(ctx, m->n[0]));
  if (m->n[0]) {
    int grid;
    DDF_CHECK(grid_for(m->n[0], 256, &grid));
    k_fill_d<<<grid, 256, 0, ctx->stream>>>(m->vol[0].p, m->n[0], 1.0);
    DDF_LAUNCH_CHECK(ctx);
  }
  if (D >= 1) DDF_CHECK((geometry_level<1, C>(m)));
  if (C >= 2 && D >= 2) DDF_CHECK((geometry_level<(C >= 2 ? 2 : 1), C>(m)));
  if (C >= 3 && D >= 3) DDF_CHECK((geometry_level<(C >= 3 ? 3 : 1), C>(m)));
  // dualvol_D = 1 (Manifolds.jl:655-656), then downward recursion
  DDF_CHECK(m->dualvol[D].alloc(ctx, m->n[D]));
  if (m->n[D]) {
    int grid;
    DDF_CHECK(grid_for(m->n[D], 256, &grid));
    k_fill_d<<<grid, 256, 0, ctx->stream>>>(m->dualvol[D].p, m->n[D], 1.0);
    DDF_LAUNCH_CHECK(ctx);
  }
  if (C >= 3 && D >= 3) DDF_CHECK((dualvol_level<(C >= 3 ? 2 : 0), C>(m)));
  if (C >= 2 && D >= 2) DDF_CHECK((dualvol_level<(C >= 2 ? 1 : 0), C>(m)));
  if (D >= 1) DDF_CHECK((dualvol_level<0, C>(m)));
  return 0;
}

extern "C" int ddf_mesh_geometry(ddf_mesh* m, int dualkind, int use_weighted_duals) {
  DDF_TRY_BEGIN
  if (dualkind != DDF_DUAL_CIRCUMCENTRIC)
    DDF_FAIL("geometry: only circumcentric duals are implemented (the reference's barycentric path is dead code, "
             "Manifolds.jl:649-652)");
  if (m->has_geometry && m->weighted == (use_weighted_duals != 0)) return 0;
  DDF_CHECK(ddf_mesh_require_assembled(m));
  if (m->D > 3 || m->C > 3) DDF_FAIL("geometry: implemented for D <= C <= 3 (got D=%d C=%d)", m->D, m->C);
  DDF_CUDA(cudaSetDevice(m->ctx->device));
  m->weighted = use_weighted_duals != 0;
  m->has_wave = false;
  int rc;
  switch (m->C) {
    case 0: rc = 0;
      DDF_CHECK(m->vol[0].alloc(m->ctx, m->n[0]));
      DDF_CHECK(m->dualvol[0].alloc(m->ctx, m->n[0]));
      if (m->n[0]) {
        k_fill_d<<<1, 256, 0, m->ctx->stream>>>(m->vol[0].p, m->n[0], 1.0);
        k_fill_d<<<1, 256, 0, m->ctx->stream>>>(m->dualvol[0].p, m->n[0], 1.0);
      }
      break;
    case 1: rc = geometry_all<1>(m); break;
    case 2: rc = geometry_all<2>(m); break;
    default: rc = geometry_all<3>(m); break;
  }
  if (rc) return rc;
  DDF_CUDA(cudaStreamSynchronize(m->ctx->stream));
  m->has_geometry = true;
  return 0;
  DDF_TRY_END
}

int ddf_mesh_require_geometry(ddf_mesh* m) {
  if (!m->has_geometry) return ddf_mesh_geometry(m, DDF_DUAL_CIRCUMCENTRIC, m->weighted ? 1 : 0);
  return 0;
}

static int download_d(ddf_mesh* m, const double* p, size_t cnt, double* out) {
  if (!cnt) return 0;
  DDF_CUDA(cudaMemcpyAsync(out, p, cnt * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return 0;
}
extern "C" int ddf_mesh_get_volumes(ddf_mesh* m, int R, double* out) {
  if (R < 0 || R > m->D) DDF_FAIL("get_volumes: R out of range");
  DDF_CHECK(ddf_mesh_require_geometry(m));
  return download_d(m, m->vol[R].p, m->n[R], out);
}
extern "C" int ddf_mesh_get_dualvolumes(ddf_mesh* m, int R, double* out) {
  if (R < 0 || R > m->D) DDF_FAIL("get_dualvolumes: R out of range");
  DDF_CHECK(ddf_mesh_require_geometry(m));
  return download_d(m, m->dualvol[R].p, m->n[R], out);
}
extern "C" int ddf_mesh_get_dualcoords(ddf_mesh* m, int R, double* out) {
  if (R < 0 || R > m->D) DDF_FAIL("get_dualcoords: R out of range");
  DDF_CHECK(ddf_mesh_require_geometry(m));
  return download_d(m, R == 0 ? m->coords.p : m->cc[R].p, (size_t)m->n[R] * m->C, out);
}

// ------------------------------------------------------------- boundary masks
// R = D-1: a face is on the boundary iff it appears an odd number of times in the
// columns of d_D (Manifolds.jl:520-534) = odd row length of the row-CSR.
__global__ void k_isbnd_top(const uint32_t* __restrict__ prow, int64_t n, uint8_t* __restrict__ mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = (uint8_t)((prow[i + 1] - prow[i]) & 1u);
}
// lower R: faces of marked (R+1)-simplices (lookup chain, Manifolds.jl:535-545)
__global__ void k_isbnd_down(const uint8_t* __restrict__ mask_up, const int32_t* __restrict__ bnd, int N, int64_t n_up,
                             uint8_t* __restrict__ mask) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_up || !mask_up[j]) return;
  for (int p = 0; p < N; p++) mask[bnd[j * N + p]] = 1;
}

int ddf_mesh_require_isboundary(ddf_mesh* m) {
  if (m->has_isbnd) return 0;
  DDF_CHECK(ddf_mesh_require_assembled(m));
  ddf_ctx* ctx = m->ctx;
  const int D = m->D;
  if (D == 0) { m->has_isbnd = true; return 0; }
  for (int R = D - 1; R >= 0; R--) {
    DDF_CHECK(m->isbnd[R].alloc(ctx, m->n[R]));
    if (m->n[R] == 0) continue;
    int grid;
    if (R == D - 1) {
      DDF_CHECK(grid_for(m->n[R], 256, &grid));
      k_isbnd_top<<<grid, 256, 0, ctx->stream>>>(m->brow_ptr[D].p, m->n[R], m->isbnd[R].p);
    } else {
      DDF_CUDA(cudaMemsetAsync(m->isbnd[R].p, 0, m->n[R], ctx->stream));
      DDF_CHECK(grid_for(m->n[R + 1], 256, &grid));
      k_isbnd_down<<<grid, 256, 0, ctx->stream>>>(m->isbnd[R + 1].p, m->bnd_table(R + 1), R + 2, m->n[R + 1],
                                                  m->isbnd[R].p);
    }
    DDF_LAUNCH_CHECK(ctx);
  }
  m->has_isbnd = true;
  return 0;
}

extern "C" int ddf_mesh_get_isboundary(ddf_mesh* m, int R, uint8_t* out) {
  DDF_TRY_BEGIN
  if (R < 0 || R >= m->D) DDF_FAIL("get_isboundary: need 0 <= R < D");
  DDF_CHECK(ddf_mesh_require_isboundary(m));
  if (!m->n[R]) return 0;
  DDF_CUDA(cudaMemcpyAsync(out, m->isbnd[R].p, m->n[R], cudaMemcpyDeviceToHost, m->ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return 0;
  DDF_TRY_END
}

// ------------------------------------------------------------------ sample R=0
// sample(Fun{D,Pr,0}, f, mfd) (src/Continuum.jl:124-158) for the example fields.
template <int C>
__global__ void k_sample(const double* __restrict__ coords, int64_t n, int kind, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[C > 0 ? C : 1];
#pragma unroll
  for (int c = 0; c < C; c++) x[c] = coords[i * C + c];
  double r;
  if (kind == 0) {                       // prod_d sinpi(x_d)            (examples/wave.jl:45, t = 0)
    r = 1.0;
#pragma unroll
    for (int c = 0; c < C; c++) r *= sinpi(x[c]);
  } else if (kind == 1) {                // sin(x1) cos(x2)              (README.md:63-65)
    r = sin(x[0]) * cos(x[C > 1 ? 1 : 0]);
  } else {                               // Gaussian rho                 (examples/poisson.jl:85-92)
    const double sigma = 0.2;
    double r2 = 0;
#pragma unroll
    for (int c = 0; c < C; c++) r2 += (x[c] - 0.1) * (x[c] - 0.1);
    const double rp = sqrt(r2) / (sqrt(2.0) * sigma);
    double nrm = 1.0;
    const double s = sqrt(2 * 3.141592653589793 * sigma * sigma);
#pragma unroll
    for (int c = 0; c < C; c++) nrm *= s;
    r = 1.0 / nrm * exp(-rp * rp);
  }
  out[i] = r;
}

extern "C" int ddf_vec_sample(ddf_mesh* m, int kind, ddf_vec* out) {
  DDF_TRY_BEGIN
  if (out->n != m->n[0]) DDF_FAIL("vec_sample: output must be a vertex cochain (n0=%lld)", (long long)m->n[0]);
  if (kind < 0 || kind > 2) DDF_FAIL("vec_sample: unknown field kind %d", kind);
  if (kind == 1 && m->C < 2) DDF_FAIL("vec_sample: field 1 needs C >= 2");
  if (!m->n[0]) return 0;
  DDF_CHECK(ddf_vec_settle(out));
  int grid;
  DDF_CHECK(grid_for(m->n[0], 256, &grid));
  switch (m->C) {
    case 1: k_sample<1><<<grid, 256, 0, m->ctx->stream>>>(m->coords.p, m->n[0], kind, out->d.p); break;
    case 2: k_sample<2><<<grid, 256, 0, m->ctx->stream>>>(m->coords.p, m->n[0], kind, out->d.p); break;
    case 3: k_sample<3><<<grid, 256, 0, m->ctx->stream>>>(m->coords.p, m->n[0], kind, out->d.p); break;
    default: DDF_FAIL("vec_sample: C=%d unsupported", m->C);
  }
  DDF_LAUNCH_CHECK(m->ctx);
  return 0;
  DDF_TRY_END
}

/* ddf_oracle.c -- C twin of the CPU oracle (oracle/ddf_oracle.py) for the loops
 * whose wall time is reported as the CPU baseline.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs; never by the product.
 *
 * These are restatements of Julia 1.5 SparseArrays `mul!` (stdlib, not vendored in
 * /root/reference; Manifest.toml:899-901) on the reference's exact storage --
 * 1-based Int64 colptr/rowval, Int8 or Float64 nzval -- anchored on the call sites
 * src/Ops.jl:273-276 and src/SparseOps.jl:438-440, and of calc_faces_boundaries
 * (src/Manifolds.jl:717-768).  Parity pinning: see the header of ddf_oracle.py.
 * Built with -ffp-contract=off: Julia rounds the multiply and the add separately.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* y = adjoint(A) * x, A::SparseMatrixCSC{Int8,Int}: primal deriv (ManifoldOps.jl:46).
 * for col: tmp = 0; for j in nzrange: tmp += nzval[j]' * x[rowval[j]]; y[col] = 0 + tmp */
void ddf_spmv_adjoint_csc_i8(int64_t ncols, const int64_t* colptr, const int64_t* rowval,
                             const int8_t* nzval, const double* x, double* y) {
  for (int64_t col = 0; col < ncols; col++) {
    double tmp = 0.0;
    for (int64_t j = colptr[col] - 1; j < colptr[col + 1] - 1; j++)
      tmp += (double)nzval[j] * x[rowval[j] - 1];
    y[col] = 0.0 + tmp;
  }
}

/* the same loop, rows distributed over OpenMP threads (a cross-check row of
 * BASELINE.md; Julia's own mul! is single-threaded) */
void ddf_spmv_adjoint_csc_i8_omp(int64_t ncols, const int64_t* colptr, const int64_t* rowval,
                                 const int8_t* nzval, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t col = 0; col < ncols; col++) {
    double tmp = 0.0;
    for (int64_t j = colptr[col] - 1; j < colptr[col + 1] - 1; j++)
      tmp += (double)nzval[j] * x[rowval[j] - 1];
    y[col] = 0.0 + tmp;
  }
}

/* y = A * x, A::SparseMatrixCSC{Int8,Int}: boundary / dual deriv.
 * fill!(y, 0); for col: ax = x[col]; for j: y[rowval[j]] += nzval[j] * ax */
void ddf_spmv_csc_i8(int64_t nrows, int64_t ncols, const int64_t* colptr, const int64_t* rowval,
                     const int8_t* nzval, const double* x, double* y) {
  memset(y, 0, (size_t)nrows * sizeof(double));
  for (int64_t col = 0; col < ncols; col++) {
    const double ax = x[col];
    for (int64_t j = colptr[col] - 1; j < colptr[col + 1] - 1; j++)
      y[rowval[j] - 1] += (double)nzval[j] * ax;
  }
}

/* y = A * x, A::SparseMatrixCSC{Float64,Int}: assembled delta, L, F (wave.jl:98-100) */
void ddf_spmv_csc_f64(int64_t nrows, int64_t ncols, const int64_t* colptr, const int64_t* rowval,
                      const double* nzval, const double* x, double* y) {
  memset(y, 0, (size_t)nrows * sizeof(double));
  for (int64_t col = 0; col < ncols; col++) {
    const double ax = x[col];
    for (int64_t j = colptr[col] - 1; j < colptr[col + 1] - 1; j++)
      y[rowval[j] - 1] += nzval[j] * ax;
  }
}

/* y = Diagonal(d) * x (hodge, ManifoldOps.jl:58-68) */
void ddf_diag_mul(int64_t n, const double* d, const double* x, double* y) {
  for (int64_t i = 0; i < n; i++) y[i] = d[i] * x[i];
}

/* leapfrog step on the assembled L (CSC) and the boundary mask (oracle wave_leapfrog) */
void ddf_wave_leapfrog_csc(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                           const double* mask /* 1-b */, double* u, double* v, double* tmp, double dt, int nsteps) {
  for (int s = 0; s < nsteps; s++) {
    ddf_spmv_csc_f64(n, n, colptr, rowval, nzval, u, tmp);
    for (int64_t i = 0; i < n; i++) v[i] = v[i] + dt * (mask[i] * tmp[i]);
    for (int64_t i = 0; i < n; i++) u[i] = u[i] + dt * (mask[i] * v[i]);
  }
}

/* ------------------------------------------------------------------ assembly */
typedef struct {
  int64_t v[5];   /* up to D = 5 vertices of a face */
  int64_t parent;
  int8_t parity;
} face_t;

static int g_nv;
static int face_cmp(const void* a, const void* b) {
  const face_t* fa = (const face_t*)a;
  const face_t* fb = (const face_t*)b;
  for (int k = 0; k < g_nv; k++) {
    if (fa->v[k] < fb->v[k]) return -1;
    if (fa->v[k] > fb->v[k]) return 1;
  }
  /* keep equal faces in parent order (the reference's sort is stable) */
  return (fa->parent > fb->parent) - (fa->parent < fb->parent);
}

/* calc_faces_boundaries (Manifolds.jl:717-768) on a 0-based (n x N) ascending vertex
 * table.  Outputs: faces (nfaces x (N-1)), and the boundary matrix as CSC with 1-based
 * Int64 colptr/rowval and Int8 nzval ((N) entries per column).  Returns nfaces. */
int64_t ddf_calc_faces_boundaries(int64_t n, int N, const int64_t* simplices, int64_t* faces_out,
                                  int64_t* colptr, int64_t* rowval, int8_t* nzval) {
  const int R = N - 1;
  const int64_t nt = n * N;
  face_t* fl = (face_t*)malloc((size_t)nt * sizeof(face_t));
  int64_t* face_of = (int64_t*)malloc((size_t)nt * sizeof(int64_t));   /* face id per (parent, omitted k) */
  for (int64_t j = 0; j < n; j++)
    for (int k = 0; k < N; k++) {
      face_t* f = &fl[j * N + k];
      int q = 0;
      for (int m = 0; m < N; m++)
        if (m != k) f->v[q++] = simplices[j * N + m];
      f->parent = j * N + k;          /* encodes (parent, k) */
      f->parity = (k % 2 == 0) ? 1 : -1;
    }
  g_nv = R;
  qsort(fl, (size_t)nt, sizeof(face_t), face_cmp);
  int64_t nfaces = 0;
  for (int64_t i = 0; i < nt; i++) {
    int isnew = (i == 0);
    if (!isnew)
      for (int k = 0; k < R; k++)
        if (fl[i].v[k] != fl[i - 1].v[k]) { isnew = 1; break; }
    if (isnew) {
      for (int k = 0; k < R; k++) faces_out[nfaces * R + k] = fl[i].v[k];
      nfaces++;
    }
    face_of[fl[i].parent] = nfaces - 1;
  }
  /* sparse(bI, bJ, bV): column j holds its N faces with rows ascending */
  for (int64_t j = 0; j <= n; j++) colptr[j] = 1 + j * N;
  for (int64_t j = 0; j < n; j++) {
    /* faces omitting k descending have ascending ids; insertion-sort to be general */
    int64_t ids[6];
    int8_t par[6];
    for (int k = 0; k < N; k++) { ids[k] = face_of[j * N + k]; par[k] = (k % 2 == 0) ? 1 : -1; }
    for (int a = 1; a < N; a++) {
      int64_t x = ids[a]; int8_t p = par[a]; int b = a - 1;
      while (b >= 0 && ids[b] > x) { ids[b + 1] = ids[b]; par[b + 1] = par[b]; b--; }
      ids[b + 1] = x; par[b + 1] = p;
    }
    for (int k = 0; k < N; k++) { rowval[j * N + k] = ids[k] + 1; nzval[j * N + k] = par[k]; }
  }
  free(fl);
  free(face_of);
  return nfaces;
}

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

/* ------------------------------------------------- assembly, linear-time variant (same results)
 * calc_faces_boundaries (Manifolds.jl:717-768) numbers the faces by the lexicographic rank of their vertex tuples.
 * The comparison sort above follows the reference literally and takes minutes at 1e9 tuples; this variant gets the
 * SAME ranks in linear time: the rank is decided first by the smallest vertex, so the tuples are counted and
 * scattered into one bucket per smallest vertex, every (small) bucket is sorted by (rest of the tuple, tuple id),
 * and the ranks follow from a prefix sum of the distinct tuples per bucket.  Used to build the reference's storage
 * for the CPU baseline at the full BASELINE size (n = 256: 1.24e9 tuples); tests/test_oracle_pins.py holds it
 * against the literal version.  The rest of a tuple is packed into 64 bits (vbits per vertex): returns -1 when it
 * does not fit (callers then use the comparison sort). */
typedef struct { uint64_t key; uint32_t tup; } bk_t;
static int bk_cmp(const void* a, const void* b) {
  const bk_t* x = (const bk_t*)a; const bk_t* y = (const bk_t*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return (x->tup > y->tup) - (x->tup < y->tup);
}
int64_t ddf_calc_faces_boundaries_linear(int64_t n, int N, int64_t nvertices, const int64_t* simplices,
                                         int64_t* faces_out, int64_t* colptr, int64_t* rowval, int8_t* nzval) {
  const int R = N - 1;
  int vbits = 1;
  while (((int64_t)1 << vbits) < nvertices) vbits++;
  if ((N - 2) * vbits > 64 || n * N >= 0xffffffffLL) return -1;
  const int64_t nt = n * N;
  uint32_t* start = (uint32_t*)calloc((size_t)nvertices + 2, sizeof(uint32_t));
  uint32_t* cursor = (uint32_t*)malloc(((size_t)nvertices + 1) * sizeof(uint32_t));
  bk_t* bk = (bk_t*)malloc((size_t)nt * sizeof(bk_t));
  uint32_t* face_of = (uint32_t*)malloc((size_t)nt * sizeof(uint32_t));
  uint32_t* ucnt = (uint32_t*)calloc((size_t)nvertices + 1, sizeof(uint32_t));
  if (!start || !cursor || !bk || !face_of || !ucnt) { free(start); free(cursor); free(bk); free(face_of); free(ucnt); return -2; }
  /* count: N-1 faces keep the smallest vertex, one face starts at the second */
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; j++) {
    __atomic_fetch_add(&start[simplices[j * N] + 1], (uint32_t)(N - 1), __ATOMIC_RELAXED);
    __atomic_fetch_add(&start[simplices[j * N + 1] + 1], 1u, __ATOMIC_RELAXED);
  }
  for (int64_t v = 0; v < nvertices; v++) start[v + 1] += start[v];
  memcpy(cursor, start, ((size_t)nvertices + 1) * sizeof(uint32_t));
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; j++) {
    const int64_t* s = simplices + j * N;
    for (int k = 0; k < N; k++) {            /* the face that leaves out vertex k */
      const int first = k == 0 ? 1 : 0;
      uint64_t key = 0;
      for (int q = first + 1; q < N; q++)
        if (q != k) key = (key << vbits) | (uint64_t)s[q];
      const uint32_t p = __atomic_fetch_add(&cursor[s[first]], 1u, __ATOMIC_RELAXED);
      bk[p].key = key;
      bk[p].tup = (uint32_t)(j * N + k);
    }
  }
  /* sort every bucket by (rest, tuple id): the order the reference's stable sort gives equal faces */
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t v = 0; v < nvertices; v++) {
    const uint32_t lo = start[v], sz = start[v + 1] - lo;
    if (sz == 0) continue;
    bk_t* b = bk + lo;
    if (sz <= 48) {
      for (uint32_t i = 1; i < sz; i++) {
        const bk_t x = b[i];
        uint32_t k = i;
        while (k > 0 && (b[k - 1].key > x.key || (b[k - 1].key == x.key && b[k - 1].tup > x.tup))) { b[k] = b[k - 1]; k--; }
        b[k] = x;
      }
    } else {
      qsort(b, sz, sizeof(bk_t), bk_cmp);
    }
    uint32_t u = 1;
    for (uint32_t i = 1; i < sz; i++) u += b[i].key != b[i - 1].key;
    ucnt[v] = u;
  }
  int64_t nfaces = 0;
  for (int64_t v = 0; v < nvertices; v++) { const uint32_t u = ucnt[v]; ucnt[v] = (uint32_t)nfaces; nfaces += u; }
  const uint64_t vmask = vbits >= 64 ? ~0ull : (((uint64_t)1 << vbits) - 1);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t v = 0; v < nvertices; v++) {
    const uint32_t lo = start[v], sz = start[v + 1] - lo;
    int64_t f = (int64_t)ucnt[v] - 1;
    for (uint32_t i = 0; i < sz; i++) {
      if (i == 0 || bk[lo + i].key != bk[lo + i - 1].key) {
        f++;
        uint64_t key = bk[lo + i].key;
        faces_out[f * R] = v;
        for (int q = R - 1; q >= 1; q--) { faces_out[f * R + q] = (int64_t)(key & vmask); key >>= vbits; }
      }
      face_of[bk[lo + i].tup] = (uint32_t)f;
    }
  }
  /* sparse(bI, bJ, bV): column j holds its N faces with rows ascending */
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j <= n; j++) colptr[j] = 1 + j * N;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; j++) {
    int64_t ids[8];
    int8_t par[8];
    for (int k = 0; k < N; k++) { ids[k] = face_of[j * N + k]; par[k] = (k % 2 == 0) ? 1 : -1; }
    for (int a = 1; a < N; a++) {
      int64_t x = ids[a]; int8_t p = par[a]; int b = a - 1;
      while (b >= 0 && ids[b] > x) { ids[b + 1] = ids[b]; par[b + 1] = par[b]; b--; }
      ids[b + 1] = x; par[b + 1] = p;
    }
    for (int k = 0; k < N; k++) { rowval[j * N + k] = ids[k] + 1; nzval[j * N + k] = par[k]; }
  }
  free(start); free(cursor); free(bk); free(face_of); free(ucnt);
  return nfaces;
}

/* large_hypercube_manifold (ManifoldConstructors.jl:196-284) simplex table for D = 3 with equal nelts: the mirrored
 * Kuhn block of 2x2x2 cells (48 tetrahedra, given by the caller as 48 x 4 x 3 corner offsets in {0,1,2}, computed by
 * the oracle's own generator) repeated over the grid in the reference's order (blocks with the first axis fastest),
 * vertex id = i0 + (n+1) i1 + (n+1)^2 i2, rows sorted ascending.  Only here so that the n = 256 table (1.0e8 rows) need
 * not be built by numpy. */
void ddf_kuhn3_table(int64_t n, const int8_t* block /* 48*4*3 */, int64_t* out /* 6 n^3 x 4 */) {
  const int64_t nb = n / 2, s1 = n + 1, s2 = (n + 1) * (n + 1);
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t b2 = 0; b2 < nb; b2++)
    for (int64_t b1 = 0; b1 < nb; b1++)
      for (int64_t b0 = 0; b0 < nb; b0++) {
        const int64_t blk = b0 + nb * (b1 + nb * b2);
        for (int t = 0; t < 48; t++) {
          int64_t v[4];
          for (int q = 0; q < 4; q++) {
            const int8_t* o = block + (t * 4 + q) * 3;
            v[q] = (2 * b0 + o[0]) + s1 * (2 * b1 + o[1]) + s2 * (2 * b2 + o[2]);
          }
          for (int a = 1; a < 4; a++) { int64_t x = v[a]; int b = a - 1; while (b >= 0 && v[b] > x) { v[b + 1] = v[b]; b--; } v[b + 1] = x; }
          for (int q = 0; q < 4; q++) out[(blk * 48 + t) * 4 + q] = v[q];
        }
      }
}

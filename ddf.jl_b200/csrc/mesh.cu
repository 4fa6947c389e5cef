// mesh.cu -- device manifold: upload, on-device large_hypercube generator and
// the assembly of the signed boundary matrices d_R (calc_faces_boundaries,
// src/Manifolds.jl:717-768, driven for R = D..1 as at :361-368).
//
// Assembly per R (bit-exact with the reference's pattern, numbering and signs):
//   K1a  k_gen_faces     one thread per (simplex j, omitted vertex k): pack the R
//                        remaining (ascending) vertex ids into one radix key, first
//                        vertex most significant  (Manifolds.jl:725-734)
//   K1b  radix sort      LSD radix sort of (key, tuple index) over R*ceil(log2 V)
//                        bits = the lexicographic `sort!(facelist; by=vertices)` (:735)
//   K1c  head flags + inclusive scan = unique-face numbering in sorted order (:745-757)
//   K1d  k_scatter_faces writes the face table simp[R-1], the fixed-width column
//                        table bnd[R] (face ids ascending; sign implicit) and the
//                        row-CSR (brow_ptr/bcol, parent | sign bit) in one pass
//                        (:759-765; the two `sparse()` calls become direct writes
//                        because the sort is stable and rows/columns come out ordered)
#include <algorithm>
#include <chrono>
#include <cub/cub.cuh>

#include "mesh.cuh"

// ------------------------------------------------------------------ utilities
__global__ void k_i64_to_i32_rows(const int64_t* __restrict__ rowval, int32_t* __restrict__ out,
                                  int64_t nsimp, int N, int64_t nvertices, int* __restrict__ bad) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nsimp) return;
  int32_t v[DDF_MAXD + 1];
  for (int k = 0; k < N; k++) {
    int64_t r = rowval[j * N + k] - 1;   // 1-based -> 0-based
    if (r < 0 || r >= nvertices) { atomicExch(bad, 1); r = 0; }
    v[k] = (int32_t)r;
  }
  // CSC columns are stored ascending; sort defensively (insertion sort, N <= 6)
  for (int a = 1; a < N; a++) {
    int32_t x = v[a];
    int b = a - 1;
    while (b >= 0 && v[b] > x) { v[b + 1] = v[b]; b--; }
    v[b + 1] = x;
  }
  for (int k = 1; k < N; k++)
    if (v[k] == v[k - 1]) atomicExch(bad, 2);
  for (int k = 0; k < N; k++) out[j * N + k] = v[k];
}

extern "C" int ddf_mesh_create(ddf_ctx* ctx, int D, int C, int64_t nvertices, int64_t nsimplices,
                               const int64_t* colptr, const int64_t* rowval, const double* coords,
                               const double* weights, ddf_mesh** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  if (D < 0 || D > DDF_MAXD) DDF_FAIL("mesh_create: D=%d unsupported (0..%d)", D, DDF_MAXD);
  if (C < D) DDF_FAIL("mesh_create: need D <= C (D=%d, C=%d)", D, C);
  if (nvertices < 0 || nsimplices < 0) DDF_FAIL("mesh_create: negative sizes");
  if (nvertices >= 0x7fffffffLL || nsimplices >= 0x7fffffffLL)
    DDF_FAIL("mesh_create: local dimension >= 2^31 (V=%lld, n_D=%lld); partition the mesh",
             (long long)nvertices, (long long)nsimplices);
  const int N = D + 1;
  for (int64_t j = 0; j <= nsimplices; j++)
    if (colptr[j] != 1 + j * N)
      DDF_FAIL("mesh_create: column %lld of simplicesD does not hold D+1=%d vertices", (long long)j, N);
  DDF_CUDA(cudaSetDevice(ctx->device));
  ddf_mesh* m = new ddf_mesh();
  m->ctx = ctx;
  m->D = D;
  m->C = C;
  m->n[0] = nvertices;
  m->n[D] = D == 0 ? nvertices : nsimplices;
  int rc = 0;
  do {
    if ((rc = m->coords.alloc(ctx, (size_t)nvertices * C))) break;
    if ((rc = m->weights.alloc(ctx, (size_t)nvertices))) break;
    if (nvertices) {
      if (C) cudaMemcpyAsync(m->coords.p, coords, (size_t)nvertices * C * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
      if (weights) cudaMemcpyAsync(m->weights.p, weights, (size_t)nvertices * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
      else cudaMemsetAsync(m->weights.p, 0, (size_t)nvertices * sizeof(double), ctx->stream);
    }
    if (D >= 1 && nsimplices > 0) {
      DevBuf<int64_t> tmp;
      DevBuf<int> bad;
      if ((rc = tmp.alloc(ctx, (size_t)nsimplices * N))) break;
      if ((rc = bad.alloc(ctx, 1))) break;
      if ((rc = m->simp[D].alloc(ctx, (size_t)nsimplices * N))) break;
      cudaMemcpyAsync(tmp.p, rowval, (size_t)nsimplices * N * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
      cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream);
      int grid;
      if ((rc = grid_for(nsimplices, 256, &grid))) break;
      k_i64_to_i32_rows<<<grid, 256, 0, ctx->stream>>>(tmp.p, m->simp[D].p, nsimplices, N, nvertices, bad.p);
      ctx->launches++;
      int hbad = 0;
      cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
      cudaError_t e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) { ddf_set_error("mesh_create: %s", cudaGetErrorString(e)); rc = 1; break; }
      if (hbad == 1) { ddf_set_error("mesh_create: vertex index out of range in simplicesD"); rc = 1; break; }
      if (hbad == 2) { ddf_set_error("mesh_create: repeated vertex inside a simplex"); rc = 1; break; }
    } else {
      cudaError_t e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) { ddf_set_error("mesh_create: %s", cudaGetErrorString(e)); rc = 1; break; }
    }
  } while (0);
  if (rc) { delete m; return rc; }
  *out = m;
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_mesh_destroy(ddf_mesh* m) {
  if (!m) return 0;
  ddf_sync(m->ctx);                 // every library stream: the tables go back to the pool in compute-stream order
  ddf_halo_release(m);
  delete m;
  return 0;
}

// ------------------------------------------- on-device large_hypercube generator
// Host: build the mirrored Kuhn block exactly as ManifoldConstructors.jl:210-214,
// 286-354 does (depth-first over unset axes, then mirror per direction).
static void kuhn_rec(int D, std::vector<int8_t>& out, std::vector<int8_t>& verts, std::vector<int8_t>& corner) {
  if ((int)verts.size() == (D + 1) * D) {
    out.insert(out.end(), verts.begin(), verts.end());
    return;
  }
  for (int d = 0; d < D; d++) {
    if (!corner[d]) {
      std::vector<int8_t> nc = corner;
      nc[d] = 1;
      std::vector<int8_t> nv = verts;
      nv.insert(nv.end(), nc.begin(), nc.end());
      kuhn_rec(D, out, nv, nc);
    }
  }
}
static std::vector<int8_t> symmetric_hypercube(int D) {
  std::vector<int8_t> simp, verts(D, 0), corner(D, 0);
  if (D == 0) { return std::vector<int8_t>(); }
  kuhn_rec(D, simp, verts, corner);
  const int per = (D + 1) * D;
  for (int dir = 0; dir < D; dir++) {
    size_t cnt = simp.size() / per;
    std::vector<int8_t> mir(simp);
    for (size_t s = 0; s < cnt; s++)
      for (int k = 0; k <= D; k++) mir[s * per + k * D + dir] = (int8_t)(2 - mir[s * per + k * D + dir]);
    simp.insert(simp.end(), mir.begin(), mir.end());
  }
  return simp;
}

struct HyperParams {
  int D;
  int T;                       // simplices per 2^D block
  int64_t half[DDF_MAXD];      // nelts / 2
  int64_t stride[DDF_MAXD];    // vertex strides
  int64_t np1[DDF_MAXD];       // nelts + 1
  double div[DDF_MAXD];        // coordinate divisor per axis
  double lw[DDF_MAXD + 1];     // level weights already divided by n^2
  int64_t z0;
};

template <int D>
__global__ void k_hypercube_simplices(HyperParams hp, const int8_t* __restrict__ sym, int32_t* __restrict__ out,
                                      int64_t nsimp) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nsimp) return;
  int64_t elt = j / hp.T;
  int s = (int)(j - elt * hp.T);
  int64_t off[D];
#pragma unroll
  for (int d = 0; d < D; d++) {          // CartesianIndices: first index fastest (:229)
    off[d] = 2 * (elt % hp.half[d]);
    elt /= hp.half[d];
  }
  int32_t v[D + 1];
  const int8_t* sp = sym + (size_t)s * (D + 1) * D;
#pragma unroll
  for (int k = 0; k <= D; k++) {
    int64_t id = 0;
#pragma unroll
    for (int d = 0; d < D; d++) id += (sp[k * D + d] + off[d]) * hp.stride[d];
    v[k] = (int32_t)id;
  }
#pragma unroll
  for (int a = 1; a <= D; a++) {         // ascending vertex order (:274-280)
    int32_t x = v[a];
    int b = a - 1;
    while (b >= 0 && v[b] > x) { v[b + 1] = v[b]; b--; }
    v[b + 1] = x;
  }
#pragma unroll
  for (int k = 0; k <= D; k++) out[j * (D + 1) + k] = v[k];
}

template <int D>
__global__ void k_hypercube_vertices(HyperParams hp, double* __restrict__ coords, double* __restrict__ weights,
                                     int64_t nv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  int64_t r = i;
  int level = 0;
#pragma unroll
  for (int d = 0; d < D; d++) {
    int64_t id = r % hp.np1[d];
    r /= hp.np1[d];
    if (d == D - 1) id += hp.z0;
    level += (int)(id & 1);                               // :269
    coords[i * D + d] = (double)id / hp.div[d];           // (elt.I - 1) ./ nelts (:241)
  }
  weights[i] = hp.lw[level];                              // :270
}

extern "C" int ddf_mesh_create_hypercube(ddf_ctx* ctx, int D, const int64_t* nelts, int64_t hdiv, int64_t z0,
                                         ddf_mesh** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  if (D < 1 || D > 4) DDF_FAIL("create_hypercube: D=%d unsupported on device (1..4)", D);
  for (int d = 0; d < D; d++)
    if (nelts[d] <= 0 || nelts[d] % 2) DDF_FAIL("create_hypercube: nelts must be positive and even (ManifoldConstructors.jl:208)");
  if (hdiv == 0)
    for (int d = 1; d < D; d++)
      if (nelts[d] != nelts[0]) DDF_FAIL("create_hypercube: all nelts must be equal unless hdiv is given (ManifoldConstructors.jl:262)");
  if (z0 % 2) DDF_FAIL("create_hypercube: z0 must be even");
  static const double LW[6][6] = {{0}, {0, 0}, {0, -1.0 / 6, 0}, {0, -1.0 / 6, -1.0 / 6, 0},
                                  {0, -3.0 / 20, -1.0 / 5, -3.0 / 20, 0},
                                  {0, -2.0 / 15, -1.0 / 5, -1.0 / 5, -2.0 / 15, 0}};
  DDF_CUDA(cudaSetDevice(ctx->device));
  HyperParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.D = D;
  hp.z0 = z0;
  int64_t nv = 1, nblocks = 1;
  for (int d = 0; d < D; d++) {
    hp.half[d] = nelts[d] / 2;
    hp.np1[d] = nelts[d] + 1;
    hp.stride[d] = nv;
    nv *= nelts[d] + 1;
    nblocks *= nelts[d] / 2;
    hp.div[d] = hdiv ? (double)hdiv : (double)nelts[d];
  }
  double wdiv = hdiv ? (double)hdiv : (double)nelts[0];
  for (int l = 0; l <= D; l++) hp.lw[l] = LW[D][l] / (wdiv * wdiv);
  std::vector<int8_t> sym = symmetric_hypercube(D);
  hp.T = (int)(sym.size() / ((D + 1) * D));
  int64_t nsimp = nblocks * hp.T;
  if (nv >= 0x7fffffffLL || nsimp >= 0x7fffffffLL)
    DDF_FAIL("create_hypercube: local dimension >= 2^31 (V=%lld, n_D=%lld); partition the mesh", (long long)nv, (long long)nsimp);
  ddf_mesh* m = new ddf_mesh();
  m->ctx = ctx;
  m->D = D;
  m->C = D;
  m->n[0] = nv;
  m->n[D] = nsimp;
  m->is_hypercube = true;
  for (int d = 0; d < D; d++) m->nelts[d] = nelts[d];
  m->hdiv = hdiv;
  m->z0 = z0;
  int rc = 0;
  do {
    DevBuf<int8_t> dsym;
    if ((rc = dsym.alloc(ctx, sym.size()))) break;
    if ((rc = m->coords.alloc(ctx, (size_t)nv * D))) break;
    if ((rc = m->weights.alloc(ctx, (size_t)nv))) break;
    if ((rc = m->simp[D].alloc(ctx, (size_t)nsimp * (D + 1)))) break;
    cudaMemcpyAsync(dsym.p, sym.data(), sym.size(), cudaMemcpyHostToDevice, ctx->stream);
    int gs, gv;
    if ((rc = grid_for(nsimp, 256, &gs))) break;
    if ((rc = grid_for(nv, 256, &gv))) break;
    switch (D) {
      case 1: k_hypercube_simplices<1><<<gs, 256, 0, ctx->stream>>>(hp, dsym.p, m->simp[D].p, nsimp);
              k_hypercube_vertices<1><<<gv, 256, 0, ctx->stream>>>(hp, m->coords.p, m->weights.p, nv); break;
      case 2: k_hypercube_simplices<2><<<gs, 256, 0, ctx->stream>>>(hp, dsym.p, m->simp[D].p, nsimp);
              k_hypercube_vertices<2><<<gv, 256, 0, ctx->stream>>>(hp, m->coords.p, m->weights.p, nv); break;
      case 3: k_hypercube_simplices<3><<<gs, 256, 0, ctx->stream>>>(hp, dsym.p, m->simp[D].p, nsimp);
              k_hypercube_vertices<3><<<gv, 256, 0, ctx->stream>>>(hp, m->coords.p, m->weights.p, nv); break;
      default: k_hypercube_simplices<4><<<gs, 256, 0, ctx->stream>>>(hp, dsym.p, m->simp[D].p, nsimp);
               k_hypercube_vertices<4><<<gv, 256, 0, ctx->stream>>>(hp, m->coords.p, m->weights.p, nv); break;
    }
    ctx->launches += 2;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ddf_set_error("create_hypercube: %s", cudaGetErrorString(e)); rc = 1; break; }
  } while (0);
  if (rc) { delete m; return rc; }
  *out = m;
  return 0;
  DDF_TRY_END
}

// -------------------------------------------------------------------- assembly
template <typename KeyT>
struct KeyBits;
template <> struct KeyBits<uint32_t> { static constexpr int bits = 32; };
template <> struct KeyBits<uint64_t> { static constexpr int bits = 64; };
template <> struct KeyBits<unsigned __int128> { static constexpr int bits = 128; };

// K1a: one thread per simplex; emits its N = R+1 faces.  Tuple index t = j*N + k.
template <typename KeyT, int N>
__global__ void k_gen_faces(const int32_t* __restrict__ simp, int64_t nsimp, int vbits, KeyT* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nsimp) return;
  int32_t v[N];
#pragma unroll
  for (int k = 0; k < N; k++) v[k] = simp[j * N + k];
#pragma unroll
  for (int k = 0; k < N; k++) {          // leave out vertex k (Manifolds.jl:728-733)
    KeyT key = 0;
#pragma unroll
    for (int q = 0; q < N; q++)
      if (q != k) key = (key << vbits) | (KeyT)(uint32_t)v[q];
    keys[j * N + k] = key;
    vals[j * N + k] = (uint32_t)(j * N + k);
  }
}

template <typename KeyT>
struct HeadFlag {
  const KeyT* keys;
  __host__ __device__ uint32_t operator()(int64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u; }
};

// K1d: one thread per sorted tuple.
template <typename KeyT, int N>
__global__ void k_scatter_faces(const KeyT* __restrict__ keys, const uint32_t* __restrict__ vals,
                                const uint32_t* __restrict__ fid_incl, int64_t ntuples, int vbits,
                                int32_t* __restrict__ faces /* nfaces x (N-1), may be null when N-1 == 1 */,
                                int32_t* __restrict__ bnd /* nsimp x N */, uint32_t* __restrict__ brow_ptr,
                                uint32_t* __restrict__ bcol) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntuples) return;
  const uint32_t f = fid_incl[i] - 1;               // 0-based face id = lexicographic rank
  const uint32_t t = vals[i];
  const uint32_t j = t / N, k = t - j * N;
  // column j of d_R: rows ascending <=> omitted vertex descending:
  // the face omitting vertex k sits at position N-1-k.
  if (bnd) bnd[(size_t)j * N + (N - 1 - k)] = (int32_t)f;
  bcol[i] = j | ((k & 1u) ? 0x80000000u : 0u);      // parity (-1)^k  (Manifolds.jl:731)
  const bool head = (i == 0) || (fid_incl[i - 1] != fid_incl[i]);
  if (head) {
    brow_ptr[f] = (uint32_t)i;
    if (faces) {
      KeyT key = keys[i];
      const KeyT mask = (((KeyT)1) << vbits) - 1;
#pragma unroll
      for (int q = N - 2; q >= 0; q--) {
        faces[(size_t)f * (N - 1) + q] = (int32_t)(uint32_t)(key & mask);
        key >>= vbits;
      }
    }
  }
  if (i == ntuples - 1) brow_ptr[f + 1] = (uint32_t)ntuples;
}

template <typename KeyT, int N>
static int assemble_level(ddf_mesh* m, int vbits) {
  constexpr int R = N - 1;
  ddf_ctx* ctx = m->ctx;
  const int64_t nsimp = m->n[R];
  const int64_t ntuples = nsimp * N;
  if (ntuples >= 0xffffffffLL) DDF_FAIL("assemble: (R+1)*n_R = %lld >= 2^32 tuples; partition the mesh", (long long)ntuples);
  DevBuf<KeyT> keys0, keys1;
  DevBuf<uint32_t> vals0, vals1, fid;
  DevBuf<uint8_t> temp;
  DDF_CHECK(keys0.alloc(ctx, ntuples));
  DDF_CHECK(keys1.alloc(ctx, ntuples));
  DDF_CHECK(vals0.alloc(ctx, ntuples));
  DDF_CHECK(vals1.alloc(ctx, ntuples));
  int grid;
  DDF_CHECK(grid_for(nsimp, 256, &grid));
  k_gen_faces<KeyT, N><<<grid, 256, 0, ctx->stream>>>(m->simp[R].p, nsimp, vbits, keys0.p, vals0.p);
  DDF_LAUNCH_CHECK(ctx);
  cub::DoubleBuffer<KeyT> dk(keys0.p, keys1.p);
  cub::DoubleBuffer<uint32_t> dv(vals0.p, vals1.p);
  size_t tbytes = 0;
  const int end_bit = R * vbits;
  DDF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tbytes, dk, dv, ntuples, 0, end_bit, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tbytes));
  DDF_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tbytes, dk, dv, ntuples, 0, end_bit, ctx->stream));
  ctx->launches += (end_bit + 7) / 8 + 2;
  const KeyT* skeys = dk.Current();
  const uint32_t* svals = dv.Current();
  // free the alternate buffers early
  if (skeys == keys0.p) keys1.release(); else keys0.release();
  if (svals == vals0.p) vals1.release(); else vals0.release();
  temp.release();
  DDF_CHECK(fid.alloc(ctx, ntuples));
  {
    HeadFlag<KeyT> hf{skeys};
    cub::CountingInputIterator<int64_t> cnt(0);
    cub::TransformInputIterator<uint32_t, HeadFlag<KeyT>, cub::CountingInputIterator<int64_t>> it(cnt, hf);
    size_t sbytes = 0;
    DDF_CUDA(cub::DeviceScan::InclusiveSum(nullptr, sbytes, it, fid.p, ntuples, ctx->stream));
    DDF_CHECK(temp.alloc(ctx, sbytes));
    DDF_CUDA(cub::DeviceScan::InclusiveSum(temp.p, sbytes, it, fid.p, ntuples, ctx->stream));
    ctx->launches += 2;
  }
  uint32_t nfaces = 0;
  DDF_CUDA(cudaMemcpyAsync(&nfaces, fid.p + (ntuples - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (R == 1 && (int64_t)nfaces != m->n[0])
    DDF_FAIL("assemble: %lld of %lld vertices are not used by any edge; the reference numbers 0-simplices by rank "
             "among used vertices (Manifolds.jl:745-757), which is inconsistent with coords -- remove unused vertices",
             (long long)(m->n[0] - nfaces), (long long)m->n[0]);
  m->n[R - 1] = nfaces;
  if (R - 1 >= 1) DDF_CHECK(m->simp[R - 1].alloc(ctx, (size_t)nfaces * R));
  if (R >= 2) DDF_CHECK(m->bnd[R].alloc(ctx, (size_t)ntuples));
  DDF_CHECK(m->brow_ptr[R].alloc(ctx, (size_t)nfaces + 1));
  DDF_CHECK(m->bcol[R].alloc(ctx, (size_t)ntuples));
  DDF_CHECK(grid_for(ntuples, 256, &grid));
  k_scatter_faces<KeyT, N><<<grid, 256, 0, ctx->stream>>>(skeys, svals, fid.p, ntuples, vbits,
                                                           R - 1 >= 1 ? m->simp[R - 1].p : nullptr,
                                                           R >= 2 ? m->bnd[R].p : nullptr, m->brow_ptr[R].p,
                                                           m->bcol[R].p);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// -------------------------------------------------------------- assembly, bucket ranking (default)
// The lexicographic rank of a face (a < b < c ...) is decided first by its smallest vertex a.  So instead of
// radix-sorting 75-bit tuples globally (20 CUB passes over 16-byte keys at C4), faces are
//   1. counted per smallest vertex (one atomicAdd per simplex and bucket),
//   2. scattered into their vertex's bucket as (rest-of-tuple packed in 64 bits, tuple id),
//   3. sorted INSIDE each bucket by (rest, tuple id) -- a warp per bucket, bitonic network in shared memory; buckets
//      hold the few dozen faces above one vertex -- and their distinct faces counted,
//   4. ranked: face id = (distinct faces of all smaller vertices) + rank inside the bucket; the same pass writes
//      faces, the column table and the row-CSR exactly like k_scatter_faces.
// Ties (the same face from several parents) are ordered by tuple id = parent id, which is what the stable radix
// sort of the reference path produces, so every table is bit-identical (tests/test_gpu_parity.py compares both
// paths with the oracle).  Used when the rest of the tuple fits 64 bits; otherwise the radix path below.

#define BK_MAX 128       // largest bucket ranked by one warp in shared memory; larger ones go to k_bucket_sort_big

template <int N>
__global__ void k_bucket_count(const int32_t* __restrict__ simp, int64_t nsimp, uint32_t* __restrict__ cnt) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nsimp) return;
  atomicAdd(&cnt[simp[j * N]], (uint32_t)(N - 1));     // the N-1 faces that keep the smallest vertex
  atomicAdd(&cnt[simp[j * N + 1]], 1u);                // the face that omits it
}

template <int N>
__global__ void k_bucket_scatter(const int32_t* __restrict__ simp, int64_t nsimp, int vbits, uint32_t* __restrict__ cursor,
                                 unsigned long long* __restrict__ rest, uint32_t* __restrict__ tup) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nsimp) return;
  int32_t v[N];
#pragma unroll
  for (int k = 0; k < N; k++) v[k] = simp[j * N + k];
  const uint32_t base = atomicAdd(&cursor[v[0]], (uint32_t)(N - 1));
#pragma unroll
  for (int k = 1; k < N; k++) {                        // leave out vertex k; the bucket is v[0]
    unsigned long long key = 0;
#pragma unroll
    for (int q = 1; q < N; q++)
      if (q != k) key = (key << vbits) | (unsigned long long)(uint32_t)v[q];
    rest[base + (k - 1)] = key;
    tup[base + (k - 1)] = (uint32_t)(j * N + k);
  }
  {                                                    // leave out vertex 0; the bucket is v[1]
    unsigned long long key = 0;
#pragma unroll
    for (int q = 2; q < N; q++) key = (key << vbits) | (unsigned long long)(uint32_t)v[q];
    const uint32_t p = atomicAdd(&cursor[v[1]], 1u);
    rest[p] = key;
    tup[p] = (uint32_t)(j * N);
  }
}

__device__ __forceinline__ bool bk_less(unsigned long long ra, uint32_t ta, unsigned long long rb, uint32_t tb) {
  return ra < rb || (ra == rb && ta < tb);
}

// warp per bucket: sort by (rest, tuple id), count the distinct faces
// Rank by counting, EL elements per lane: ONE walk over the bucket's shared-memory copy for all of a lane's elements.
// NARROW (the rest of the tuple fits 32 bits: every level but the first of a 3D mesh): rest and tuple id form ONE
// 64-bit key, a comparison is one 64-bit compare; otherwise (rest, tuple id) lexicographically.  The distinct faces of
// the bucket are counted AFTER the ranking from the sorted keys (a head differs from its predecessor) instead of by a
// second counter inside the O(size^2) walk.
template <int EL, bool NARROW>
__device__ __forceinline__ uint32_t bucket_rank_walk(unsigned long long* __restrict__ sk, const uint32_t* __restrict__ st,
                                                     uint32_t sz, uint32_t lo, int lane, unsigned long long* __restrict__ rest,
                                                     uint32_t* __restrict__ tup) {
  unsigned long long myk[EL];
  uint32_t myt[EL], rank[EL];
#pragma unroll
  for (int e = 0; e < EL; e++) {
    const uint32_t i = (uint32_t)e * 32u + (uint32_t)lane;
    myk[e] = i < sz ? sk[i] : ~0ull;
    myt[e] = (!NARROW && i < sz) ? st[i] : ~0u;
    rank[e] = 0;
  }
  for (uint32_t j = 0; j < sz; j++) {
    const unsigned long long kj = sk[j];
    const uint32_t tj = NARROW ? 0u : st[j];
#pragma unroll
    for (int e = 0; e < EL; e++) rank[e] += NARROW ? (kj < myk[e]) : ((kj < myk[e]) || (kj == myk[e] && tj < myt[e]));
  }
  __syncwarp();                                           // every lane has finished reading the unsorted bucket
#pragma unroll
  for (int e = 0; e < EL; e++) {
    const uint32_t i = (uint32_t)e * 32u + (uint32_t)lane;
    if (i < sz) {
      const unsigned long long r = NARROW ? (myk[e] >> 32) : myk[e];
      rest[lo + rank[e]] = r;
      tup[lo + rank[e]] = NARROW ? (uint32_t)myk[e] : myt[e];
      sk[rank[e]] = r;                                    // sorted rests, for the head count
    }
  }
  __syncwarp();
  uint32_t u = 0;
#pragma unroll
  for (int e = 0; e < EL; e++) {
    const uint32_t i = (uint32_t)e * 32u + (uint32_t)lane;
    const bool head = i < sz && (i == 0 || sk[i] != sk[i - 1]);       // first (smallest parent) of its face
    u += __popc(__ballot_sync(0xffffffffu, head));
  }
  return u;
}

template <bool NARROW>
__global__ void __launch_bounds__(256)
k_bucket_sort(const uint32_t* __restrict__ start, int64_t nbuckets, unsigned long long* __restrict__ rest,
              uint32_t* __restrict__ tup, uint32_t* __restrict__ ucnt, uint32_t* __restrict__ big /* [0] = count, then ids */) {
  __shared__ unsigned long long sr[8][BK_MAX];
  __shared__ uint32_t st[NARROW ? 1 : 8][NARROW ? 1 : BK_MAX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * 8 + w;
  if (b >= nbuckets) return;
  const uint32_t lo = start[b], sz = start[b + 1] - lo;
  if (sz == 0) { if (lane == 0) ucnt[b] = 0u; return; }
  if (sz > BK_MAX) {                                  // a high-valence vertex: whole blocks sort it (k_bucket_sort_big)
    if (lane == 0) big[1 + atomicAdd(&big[0], 1u)] = (uint32_t)b;
    return;
  }
  // every lane compares its element(s) with all sz elements of the bucket (broadcast reads of shared memory, no
  // barriers) -- cheaper than a bitonic network for the few dozen faces above one vertex
  for (uint32_t i = lane; i < sz; i += 32) {
    if (NARROW) {
      sr[w][i] = (rest[lo + i] << 32) | (unsigned long long)tup[lo + i];
    } else {
      sr[w][i] = rest[lo + i];
      st[NARROW ? 0 : w][i] = tup[lo + i];
    }
  }
  __syncwarp();
  // (most buckets of the triangle level hold 33..64 tuples: two elements per lane in one walk)
  uint32_t u;
  switch ((sz + 31u) >> 5) {                             // warp-uniform: elements per lane in this bucket
    case 1: u = bucket_rank_walk<1, NARROW>(sr[w], st[NARROW ? 0 : w], sz, lo, lane, rest, tup); break;
    case 2: u = bucket_rank_walk<2, NARROW>(sr[w], st[NARROW ? 0 : w], sz, lo, lane, rest, tup); break;
    case 3: u = bucket_rank_walk<3, NARROW>(sr[w], st[NARROW ? 0 : w], sz, lo, lane, rest, tup); break;
    default: u = bucket_rank_walk<4, NARROW>(sr[w], st[NARROW ? 0 : w], sz, lo, lane, rest, tup); break;
  }
  if (lane == 0) ucnt[b] = u;
}

// Buckets above BK_MAX tuples (a vertex of very high valence): one block of 1024 threads per bucket sorts it in
// place in global memory with a bitonic network written with ascending comparators only ("flip" then "disperse"
// steps), so that any length works -- positions beyond the bucket act as +infinity and are never touched.  Sorted by
// (rest, tuple id) like the warp path; then the distinct faces are counted.  O(n log^2 n / 1024) per bucket.
__global__ void __launch_bounds__(1024)
k_bucket_sort_big(const uint32_t* __restrict__ start, unsigned long long* __restrict__ rest, uint32_t* __restrict__ tup,
                  uint32_t* __restrict__ ucnt, const uint32_t* __restrict__ big) {
  __shared__ uint32_t wsum[32];
  const uint32_t nbig = big[0];
  for (uint32_t q = blockIdx.x; q < nbig; q += gridDim.x) {
    const uint32_t b = big[1 + q];
    const uint32_t lo = start[b], sz = start[b + 1] - lo;
    unsigned long long* r = rest + lo;
    uint32_t* t = tup + lo;
    uint32_t P = 1;
    while (P < sz) P <<= 1;
    for (uint32_t k = 2; k <= P; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        const bool flip = j == (k >> 1);
        for (uint32_t h = threadIdx.x; h < (P >> 1); h += blockDim.x) {
          // h-th comparator of this step: lower index i has bit j clear
          const uint32_t i = ((h & ~(j - 1u)) << 1) | (h & (j - 1u));
          const uint32_t p = flip ? (i ^ (k - 1u)) : (i | j);
          if (p < sz) {
            const unsigned long long ri = r[i], rp = r[p];
            const uint32_t ti = t[i], tp = t[p];
            if (bk_less(rp, tp, ri, ti)) { r[i] = rp; t[i] = tp; r[p] = ri; t[p] = ti; }
          }
        }
        __syncthreads();
      }
    }
    uint32_t u = 0;
    for (uint32_t i = threadIdx.x; i < sz; i += blockDim.x) u += (i == 0 || r[i] != r[i - 1]) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = u;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += wsum[w];
      ucnt[b] = tot;
    }
    __syncthreads();
  }
}

// warp per bucket: face ids and every output table (the bucket variant of k_scatter_faces)
template <int N>
__global__ void __launch_bounds__(256)
k_bucket_rank(const uint32_t* __restrict__ start, const uint32_t* __restrict__ fbase, int64_t nbuckets,
              const unsigned long long* __restrict__ rest, const uint32_t* __restrict__ tup, int vbits,
              int32_t* __restrict__ faces, int32_t* __restrict__ bnd, uint32_t* __restrict__ brow_ptr,
              uint32_t* __restrict__ bcol, uint32_t ntuples, uint32_t nfaces) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * 8 + w;
  if (b == 0 && lane == 0) brow_ptr[nfaces] = ntuples;
  if (b >= nbuckets) return;
  const uint32_t lo = start[b], sz = start[b + 1] - lo;
  uint32_t running = 0;
  const uint32_t fb = fbase[b];
  for (uint32_t i0 = 0; i0 < sz; i0 += 32) {
    const uint32_t i = i0 + lane;
    const bool valid = i < sz;
    unsigned long long r = 0;
    uint32_t t = 0;
    bool head = false;
    if (valid) {
      r = rest[lo + i];
      t = tup[lo + i];
      head = i == 0 || r != rest[lo + i - 1];
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, head);
    if (valid) {
      const uint32_t f = fb + running + __popc(mask & (0xffffffffu >> (31 - lane))) - 1u;
      const uint32_t gpos = lo + i;
      const uint32_t j = t / N, k = t - j * N;
      if (bnd) bnd[(size_t)j * N + (N - 1 - k)] = (int32_t)f;
      bcol[gpos] = j | ((k & 1u) ? 0x80000000u : 0u);
      if (head) {
        brow_ptr[f] = gpos;
        if (faces) {
          const unsigned long long vm = (1ull << vbits) - 1ull;
          faces[(size_t)f * (N - 1)] = (int32_t)b;
#pragma unroll
          for (int q = N - 2; q >= 1; q--) {
            faces[(size_t)f * (N - 1) + q] = (int32_t)(uint32_t)(r & vm);
            r >>= vbits;
          }
        }
      }
    }
    running += __popc(mask);
  }
}

// ---------------------------------------------------------------- exclusive prefix sum (uint32), hand-written
// Three passes: per-tile sums, a single-block scan of the tile sums, per-tile scan + offset.  Tiles of 4096.
#define SCAN_TILE 4096
__global__ void __launch_bounds__(256) k_scan_tile_sums(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ sums) {
  __shared__ uint32_t ws[8];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  uint32_t acc = 0;
  for (int k = threadIdx.x; k < SCAN_TILE; k += 256) acc += base + k < n ? in[base + k] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; w++) t += ws[w];
    sums[blockIdx.x] = t;
  }
}
// one block: exclusive scan of `m` tile sums in place (m is a few thousand)
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ sums, int64_t m) {
  __shared__ uint32_t part[1024];
  const int64_t per = (m + 1023) / 1024;
  const int64_t lo = threadIdx.x * per, hi = min(lo + per, m);
  uint32_t acc = 0;
  for (int64_t k = lo; k < hi; k++) acc += sums[k];
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int k = 0; k < 1024; k++) { const uint32_t v = part[k]; part[k] = run; run += v; }
  }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (int64_t k = lo; k < hi; k++) { const uint32_t v = sums[k]; sums[k] = run; run += v; }
}
__global__ void __launch_bounds__(256) k_scan_tiles(const uint32_t* __restrict__ in, int64_t n, const uint32_t* __restrict__ sums,
                                                   uint32_t* __restrict__ out) {
  __shared__ uint32_t ws[8];
  __shared__ uint32_t carry;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = sums[blockIdx.x];
  __syncthreads();
  for (int c = 0; c < SCAN_TILE; c += 256) {            // 16 chunks of 256, scanned in order
    const int64_t i = base + c + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int k = 0; k < w; k++) woff += ws[k];
    const uint32_t c0 = carry;
    if (i < n) out[i] = c0 + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 255) carry = c0 + woff + incl;
    __syncthreads();
  }
}
// out[i] = in[0] + ... + in[i-1]; `sums` holds ceil(n / SCAN_TILE) words of scratch
static int scan_exclusive_u32(ddf_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, uint32_t* sums) {
  if (n <= 0) return 0;
  const int64_t tiles = ceil_div64(n, SCAN_TILE);
  if (tiles > 0x7fffffffLL) DDF_FAIL("scan: too many tiles");
  k_scan_tile_sums<<<(int)tiles, 256, 0, ctx->stream>>>(in, n, sums);
  DDF_LAUNCH_CHECK(ctx);
  k_scan_sums<<<1, 1024, 0, ctx->stream>>>(sums, tiles);
  DDF_LAUNCH_CHECK(ctx);
  k_scan_tiles<<<(int)tiles, 256, 0, ctx->stream>>>(in, n, sums, out);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

// one grow-only device workspace for all levels of an assembly: cudaMalloc / cudaFree of several GB cost
// milliseconds per GB, more than the kernels they would serve
struct AsmWork {
  DevBuf<unsigned char> buf;
  int ensure(ddf_ctx* ctx, size_t bytes) {
    if (buf.n >= bytes) return 0;
    buf.release();
    return buf.alloc(ctx, bytes);
  }
};
// the workspace of the ddf_mesh_assemble in flight is held by its ctx (ctx->asm_work), not by a process global

// bytes of workspace one level of the bucket ranking needs (layout in assemble_level_bucket)
static size_t bucket_ws_bytes(int64_t ntuples, int64_t nv) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t tb = (size_t)ceil_div64(nv + 1, 4096) * 4;
  return up((size_t)ntuples * 8) + up((size_t)ntuples * 4) + 5 * up((size_t)(nv + 1) * 4) + up(tb) +
         up(((size_t)ntuples / BK_MAX + 2) * 4);
}

template <int N>
static int assemble_level_bucket(ddf_mesh* m, int vbits) {
  constexpr int R = N - 1;
  ddf_ctx* ctx = m->ctx;
  const int64_t nsimp = m->n[R], nv = m->n[0];
  const int64_t ntuples = nsimp * N;
  if (ntuples >= 0xffffffffLL) DDF_FAIL("assemble: (R+1)*n_R = %lld >= 2^32 tuples; partition the mesh", (long long)ntuples);
  // workspace layout: rest[ntuples] u64 | tup[ntuples] u32 | cnt, start, cursor, ucnt, fbase [nv+1] u32 each | scan temp
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t tb = (size_t)ceil_div64(nv + 1, SCAN_TILE) * 4;       // scratch of the hand-written scan
  const size_t o_rest = 0, o_tup = up((size_t)ntuples * 8), o_cnt = o_tup + up((size_t)ntuples * 4);
  const size_t vsz = up((size_t)(nv + 1) * 4);
  const size_t o_start = o_cnt + vsz, o_cursor = o_start + vsz, o_ucnt = o_cursor + vsz, o_fbase = o_ucnt + vsz;
  const size_t o_temp = o_fbase + vsz, o_big = o_temp + up(tb);
  const size_t total = o_big + up(((size_t)ntuples / BK_MAX + 2) * 4);      // ids of the buckets above BK_MAX tuples
  AsmWork local;
  AsmWork* wk = ctx->asm_work ? static_cast<AsmWork*>(ctx->asm_work) : &local;
  const bool dbg = getenv("DDF_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_a = now();
  DDF_CHECK(wk->ensure(ctx, total));
  const double t_b = now();
  unsigned char* base = wk->buf.p;
  unsigned long long* rest = reinterpret_cast<unsigned long long*>(base + o_rest);
  uint32_t* tup = reinterpret_cast<uint32_t*>(base + o_tup);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(base + o_cnt);
  uint32_t* start = reinterpret_cast<uint32_t*>(base + o_start);
  uint32_t* cursor = reinterpret_cast<uint32_t*>(base + o_cursor);
  uint32_t* ucnt = reinterpret_cast<uint32_t*>(base + o_ucnt);
  uint32_t* fbase = reinterpret_cast<uint32_t*>(base + o_fbase);
  uint32_t* temp = reinterpret_cast<uint32_t*>(base + o_temp);
  uint32_t* big = reinterpret_cast<uint32_t*>(base + o_big);
  DDF_CUDA(cudaMemsetAsync(big, 0, 4, ctx->stream));
  DDF_CUDA(cudaMemsetAsync(cnt, 0, (nv + 1) * 4, ctx->stream));
  int gs, gb;
  DDF_CHECK(grid_for(nsimp, 256, &gs));
  DDF_CHECK(grid_for(ceil_div64(nv, 8) * 256, 256, &gb));
  k_bucket_count<N><<<gs, 256, 0, ctx->stream>>>(m->simp[R].p, nsimp, cnt);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CHECK(scan_exclusive_u32(ctx, cnt, start, nv + 1, temp));
  DDF_CUDA(cudaMemcpyAsync(cursor, start, (nv + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  k_bucket_scatter<N><<<gs, 256, 0, ctx->stream>>>(m->simp[R].p, nsimp, vbits, cursor, rest, tup);
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaMemsetAsync(ucnt + nv, 0, 4, ctx->stream));
  // the rest of a face tuple (its N - 2 larger vertices) fits 32 bits at every level but the first of a 3D mesh
  if ((N - 2) * vbits <= 32) k_bucket_sort<true><<<gb, 256, 0, ctx->stream>>>(start, nv, rest, tup, ucnt, big);
  else k_bucket_sort<false><<<gb, 256, 0, ctx->stream>>>(start, nv, rest, tup, ucnt, big);
  DDF_LAUNCH_CHECK(ctx);
  k_bucket_sort_big<<<2 * ctx->sm_count, 1024, 0, ctx->stream>>>(start, rest, tup, ucnt, big);   // no-op without big buckets
  DDF_LAUNCH_CHECK(ctx);
  DDF_CHECK(scan_exclusive_u32(ctx, ucnt, fbase, nv + 1, temp));
  uint32_t nfaces = 0;
  DDF_CUDA(cudaMemcpyAsync(&nfaces, fbase + nv, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (R == 1 && (int64_t)nfaces != m->n[0])
    DDF_FAIL("assemble: %lld of %lld vertices are not used by any edge; the reference numbers 0-simplices by rank "
             "among used vertices (Manifolds.jl:745-757), which is inconsistent with coords -- remove unused vertices",
             (long long)(m->n[0] - nfaces), (long long)m->n[0]);
  m->n[R - 1] = nfaces;
  const double t_c = now();
  if (R - 1 >= 1) DDF_CHECK(m->simp[R - 1].alloc(ctx, (size_t)nfaces * R));
  if (R >= 2) DDF_CHECK(m->bnd[R].alloc(ctx, (size_t)ntuples));
  DDF_CHECK(m->brow_ptr[R].alloc(ctx, (size_t)nfaces + 1));
  DDF_CHECK(m->bcol[R].alloc(ctx, (size_t)ntuples));
  const double t_d = now();
  if (dbg)
    fprintf(stderr, "[ddf] assemble R=%d: workspace %.1f ms (%.2f GB), count..sort %.1f ms, output alloc %.1f ms\n", R,
            t_b - t_a, total * 1e-9, t_c - t_b, t_d - t_c);
  k_bucket_rank<N><<<gb, 256, 0, ctx->stream>>>(start, fbase, nv, rest, tup, vbits,
                                               R - 1 >= 1 ? m->simp[R - 1].p : nullptr, R >= 2 ? m->bnd[R].p : nullptr,
                                               m->brow_ptr[R].p, m->bcol[R].p, (uint32_t)ntuples, nfaces);
  DDF_LAUNCH_CHECK(ctx);
  if (!ctx->asm_work) DDF_CUDA(cudaStreamSynchronize(ctx->stream));     // the local workspace is freed on return
  return 0;
}

template <int N>
static int assemble_level_dispatch(ddf_mesh* m, int vbits) {
  if (!m->ctx->tune.asm_path && (N - 2) * vbits <= 64) return assemble_level_bucket<N>(m, vbits);
  const int bits = (N - 1) * vbits;
  if (bits <= 32) return assemble_level<uint32_t, N>(m, vbits);
  if (bits <= 64) return assemble_level<uint64_t, N>(m, vbits);
  if (bits <= 128) return assemble_level<unsigned __int128, N>(m, vbits);
  ddf_set_error("assemble: face key needs %d bits (> 128)", bits);
  return 1;
}

extern "C" int ddf_mesh_assemble(ddf_mesh* m) {
  DDF_TRY_BEGIN
  if (m->assembled) return 0;
  DDF_CUDA(cudaSetDevice(m->ctx->device));
  int vbits = 1;
  while (((int64_t)1 << vbits) < m->n[0]) vbits++;
  AsmWork work;                          // shared by all levels, released (after a sync) when this call returns
  struct WorkScope {
    ddf_ctx* c;
    ~WorkScope() { cudaStreamSynchronize(c->stream); c->asm_work = nullptr; }
  } scope{m->ctx};
  m->ctx->asm_work = &work;
  {
    // One allocation for all levels: the lower levels of a mesh usually hold MORE tuples than the top one (3D Kuhn:
    // 4.0e8 / 6.1e8 / 2.4e8), and growing the workspace in the middle of the assembly costs a free + an allocation
    // of several GB.  Face counts are predicted from "a face is shared by two simplices" (+3 %); a mesh that beats
    // the prediction just takes the grow path.
    size_t need = 0;
    int64_t nR = m->n[m->D];
    for (int R = m->D; R >= 1 && nR > 0; R--) {
      const int64_t ntuples = nR * (R + 1);
      need = std::max(need, bucket_ws_bytes(ntuples, m->n[0]));
      nR = (int64_t)((double)ntuples / 2 * 1.03) + 1024;       // predicted n_{R-1}
    }
    if (!m->ctx->tune.asm_path && need) DDF_CHECK(work.ensure(m->ctx, need));
  }
  for (int R = m->D; R >= 1; R--) {
    if (m->n[R] == 0) {                 // empty manifold: every level is empty
      m->n[R - 1] = (R - 1 == 0) ? m->n[0] : 0;
      if (m->n[0] != 0 && R == 1) DDF_FAIL("assemble: vertices without simplices");
      DDF_CHECK(m->brow_ptr[R].alloc(m->ctx, (size_t)m->n[R - 1] + 1));
      if (m->n[R - 1] + 1) DDF_CUDA(cudaMemsetAsync(m->brow_ptr[R].p, 0, (m->n[R - 1] + 1) * 4, m->ctx->stream));
      continue;
    }
    int rc;
    switch (R) {
      case 1: rc = assemble_level_dispatch<2>(m, vbits); break;
      case 2: rc = assemble_level_dispatch<3>(m, vbits); break;
      case 3: rc = assemble_level_dispatch<4>(m, vbits); break;
      case 4: rc = assemble_level_dispatch<5>(m, vbits); break;
      default: rc = assemble_level_dispatch<6>(m, vbits); break;
    }
    if (rc) return rc;
  }
  m->assembled = true;
  return 0;
  DDF_TRY_END
}

int ddf_mesh_require_assembled(ddf_mesh* m) {
  if (!m->assembled) return ddf_mesh_assemble(m);
  return 0;
}

extern "C" int ddf_mesh_nsimplices(ddf_mesh* m, int R, int64_t* out) {
  if (R < 0 || R > m->D) DDF_FAIL("nsimplices: R=%d out of range 0..%d", R, m->D);
  if (R != m->D && R != 0) DDF_CHECK(ddf_mesh_require_assembled(m));
  *out = m->n[R];
  return 0;
}
extern "C" int ddf_mesh_dims(ddf_mesh* m, int* D, int* C) { *D = m->D; *C = m->C; return 0; }

// ---------------------------------------------------------------------- export
__global__ void k_export_rows(const int32_t* __restrict__ tab, int64_t count, int64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (int64_t)tab[i] + 1;
}

static int export_table(ddf_mesh* m, const int32_t* tab, int64_t nrows, int N, int64_t* colptr, int64_t* rowval) {
  ddf_ctx* ctx = m->ctx;
  for (int64_t j = 0; j <= nrows; j++) colptr[j] = 1 + j * N;
  int64_t count = nrows * N;
  if (count == 0) return 0;
  if (!tab) {                                       // R = 0: the identity table
    for (int64_t i = 0; i < count; i++) rowval[i] = i + 1;
    return 0;
  }
  const int64_t CH = (int64_t)1 << 26;              // 64 Mi entries (512 MB) per chunk
  DevBuf<int64_t> tmp;
  DDF_CHECK(tmp.alloc(ctx, (size_t)(count < CH ? count : CH)));
  for (int64_t off = 0; off < count; off += CH) {
    int64_t c = count - off < CH ? count - off : CH;
    int grid;
    DDF_CHECK(grid_for(c, 256, &grid));
    k_export_rows<<<grid, 256, 0, ctx->stream>>>(tab + off, c, tmp.p);
    DDF_LAUNCH_CHECK(ctx);
    DDF_CUDA(cudaMemcpyAsync(rowval + off, tmp.p, c * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

extern "C" int ddf_mesh_get_simplices_csc(ddf_mesh* m, int R, int64_t* colptr, int64_t* rowval) {
  DDF_TRY_BEGIN
  if (R < 0 || R > m->D) DDF_FAIL("get_simplices: R=%d out of range", R);
  if (R != m->D) DDF_CHECK(ddf_mesh_require_assembled(m));
  return export_table(m, R == 0 ? nullptr : m->simp[R].p, m->n[R], R + 1, colptr, rowval);
  DDF_TRY_END
}

extern "C" int ddf_mesh_get_boundary_csc(ddf_mesh* m, int R, int64_t* colptr, int64_t* rowval, int8_t* nzval) {
  DDF_TRY_BEGIN
  if (R < 1 || R > m->D) DDF_FAIL("get_boundary: R=%d out of range 1..%d", R, m->D);
  DDF_CHECK(ddf_mesh_require_assembled(m));
  DDF_CHECK(export_table(m, m->bnd_table(R), m->n[R], R + 1, colptr, rowval));
  const int N = R + 1;
  for (int64_t j = 0; j < m->n[R]; j++)
    for (int p = 0; p < N; p++) nzval[j * N + p] = ((R + p) & 1) ? -1 : 1;   // sign (-1)^(R+p)
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_mesh_get_coords(ddf_mesh* m, double* out) {
  size_t cnt = (size_t)m->n[0] * m->C;
  if (!cnt) return 0;
  DDF_CUDA(cudaMemcpyAsync(out, m->coords.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return 0;
}
extern "C" int ddf_mesh_get_weights(ddf_mesh* m, double* out) {
  if (!m->n[0]) return 0;
  DDF_CUDA(cudaMemcpyAsync(out, m->weights.p, m->n[0] * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return 0;
}

// lookup.cu -- SURVEY 8(f) N4: incidence lookup tables and edge-midpoint refinement on the device.
//   get_lookup(mfd, Ri, Rj)   src/Manifolds.jl:455-498   n_Ri x n_Rj pattern matrix (`One` values) of "Ri-simplex i is a
//                             sub-simplex (Ri < Rj) / super-simplex (Ri > Rj) of Rj-simplex j"; Ri == 0: the vertex tables,
//                             Ri == Rj: the identity, Ri == Rj - 1: |boundaries[Rj]|, Ri < Rj - 1: chains of those, Ri > Rj:
//                             the transpose
//   refine_coords             src/Meshing.jl:127-146     old vertices, then the edge midpoints in edge-id order
// The reference builds the chains by sparse products of pattern matrices; here a sub-simplex is FOUND: the vertex table
// of every rank below D is sorted lexicographically (face ids are lexicographic ranks, Manifolds.jl:735-757), so the id
// of a vertex subset is a binary search.  The transpose direction is one stable radix sort of (sub-simplex, simplex)
// pairs.  Setup-time code, not on the per-step path.
#include <cub/cub.cuh>

#include "mesh.cuh"

template <int NI>
__device__ __forceinline__ int64_t find_tuple(const int32_t* __restrict__ tab, int64_t n, const int32_t* key) {
  int64_t lo = 0, hi = n;                      // first row >= key
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    bool less = false;
#pragma unroll
    for (int q = 0; q < NI; q++) {
      const int32_t v = tab[mid * NI + q];
      if (v != key[q]) { less = v < key[q]; break; }
    }
    if (less) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// column j of lookup(Ri, Rj), Ri < Rj: the ids of all (Ri+1)-subsets of simplex j, ascending (subsets enumerated in
// lexicographic order of positions = lexicographic order of the vertex tuples = ascending ids)
template <int NI>
__global__ void k_lookup_down(const int32_t* __restrict__ simpJ, int NJ, int64_t nJ, const int32_t* __restrict__ simpI,
                              int64_t nI, int ncomb, const uint8_t* __restrict__ comb /* ncomb x NI positions */,
                              int32_t* __restrict__ out, int* __restrict__ bad) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nJ) return;
  for (int c = 0; c < ncomb; c++) {
    int32_t key[NI];
#pragma unroll
    for (int q = 0; q < NI; q++) key[q] = simpJ[j * NJ + comb[c * NI + q]];
    int64_t id;
    if (NI == 1) id = key[0];                    // the vertex table is the identity
    else {
      id = find_tuple<NI>(simpI, nI, key);
      bool hit = id < nI;
      if (hit)
#pragma unroll
        for (int q = 0; q < NI; q++) hit = hit && simpI[id * NI + q] == key[q];
      if (!hit) { atomicExch(bad, 1); id = 0; }
    }
    out[j * ncomb + c] = (int32_t)id;
  }
}

__global__ void k_pairs_keys(const int32_t* __restrict__ down, int64_t nJ, int ncomb, unsigned long long* __restrict__ keys) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nJ * ncomb) return;
  keys[e] = ((unsigned long long)(uint32_t)down[e] << 32) | (unsigned long long)(uint32_t)(e / ncomb);
}
__global__ void k_keys_split(const unsigned long long* __restrict__ keys, int64_t n, int64_t ncols, int64_t* __restrict__ rowval,
                             unsigned long long* __restrict__ colcount) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  rowval[e] = (int64_t)(keys[e] & 0xffffffffull) + 1;
  atomicAdd(&colcount[(keys[e] >> 32) + 1], 1ull);
}
__global__ void k_i32_to_i64p1(const int32_t* __restrict__ a, int64_t n, int64_t* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = (int64_t)a[e] + 1;
}

static void combos(int n, int k, std::vector<uint8_t>& out) {
  std::vector<int> idx(k);
  for (int i = 0; i < k; i++) idx[i] = i;
  for (;;) {
    for (int i = 0; i < k; i++) out.push_back((uint8_t)idx[i]);
    int i = k - 1;
    while (i >= 0 && idx[i] == n - k + i) i--;
    if (i < 0) break;
    idx[i]++;
    for (int q = i + 1; q < k; q++) idx[q] = idx[q - 1] + 1;
  }
}

// down table of lookup(Ri, Rj), Ri < Rj: n_Rj x C(Rj+1, Ri+1) ids
static int lookup_down(ddf_mesh* m, int Ri, int Rj, DevBuf<int32_t>& down, int* ncomb_out) {
  ddf_ctx* ctx = m->ctx;
  std::vector<uint8_t> comb;
  combos(Rj + 1, Ri + 1, comb);
  const int ncomb = (int)(comb.size() / (Ri + 1));
  *ncomb_out = ncomb;
  const int64_t nJ = m->n[Rj];
  DDF_CHECK(down.alloc(ctx, (size_t)nJ * ncomb));
  if (!nJ) return 0;
  DevBuf<uint8_t> dcomb;
  DevBuf<int> bad;
  DDF_CHECK(dcomb.alloc(ctx, comb.size()));
  DDF_CHECK(bad.alloc(ctx, 1));
  DDF_CUDA(cudaMemcpyAsync(dcomb.p, comb.data(), comb.size(), cudaMemcpyHostToDevice, ctx->stream));
  DDF_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
  int grid;
  DDF_CHECK(grid_for(nJ, 128, &grid));
  const int32_t* sj = m->simp[Rj].p;
  const int32_t* si = Ri >= 1 ? m->simp[Ri].p : nullptr;
#define DDF_LOOKUP(NI) k_lookup_down<NI><<<grid, 128, 0, ctx->stream>>>(sj, Rj + 1, nJ, si, m->n[Ri], ncomb, dcomb.p, down.p, bad.p)
  switch (Ri + 1) {
    case 1: DDF_LOOKUP(1); break;
    case 2: DDF_LOOKUP(2); break;
    case 3: DDF_LOOKUP(3); break;
    case 4: DDF_LOOKUP(4); break;
    case 5: DDF_LOOKUP(5); break;
    default: DDF_FAIL("lookup: Ri = %d not supported", Ri);
  }
#undef DDF_LOOKUP
  DDF_LAUNCH_CHECK(ctx);
  int hbad = 0;
  DDF_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hbad) DDF_FAIL("lookup(%d,%d): a sub-simplex is missing from the rank-%d table (mesh not assembled consistently)", Ri, Rj, Ri);
  return 0;
}

// get_lookup(mfd, Ri, Rj).op as the reference stores it: index-only CSC (n_Ri x n_Rj), 1-based Int64.
// Call with colptr == NULL to query nnz.
extern "C" int ddf_mesh_get_lookup_csc(ddf_mesh* m, int Ri, int Rj, int64_t* nnz, int64_t* colptr, int64_t* rowval) {
  DDF_TRY_BEGIN
  if (Ri < 0 || Ri > m->D || Rj < 0 || Rj > m->D) DDF_FAIL("get_lookup: need 0 <= Ri, Rj <= D (Manifolds.jl:455-456)");
  DDF_CHECK(ddf_mesh_require_assembled(m));
  ddf_ctx* ctx = m->ctx;
  DDF_CUDA(cudaSetDevice(ctx->device));
  const int64_t nI = m->n[Ri], nJ = m->n[Rj];
  if (Ri == Rj) {                                        // identity (:482-485)
    *nnz = nI;
    if (!colptr) return 0;
    for (int64_t j = 0; j <= nJ; j++) colptr[j] = j + 1;
    for (int64_t j = 0; j < nJ; j++) rowval[j] = j + 1;
    return 0;
  }
  if (Ri < Rj) {                                         // simplex definitions, |boundaries|, chains (:479-493)
    int ncomb = 0;
    DevBuf<int32_t> down;
    // only the count is needed for the query
    int64_t c = 1;
    for (int q = 0; q < Ri + 1; q++) c = c * (Rj + 1 - q) / (q + 1);
    *nnz = nJ * c;
    if (!colptr) return 0;
    DDF_CHECK(lookup_down(m, Ri, Rj, down, &ncomb));
    DevBuf<int64_t> r64;
    DDF_CHECK(r64.alloc(ctx, (size_t)nJ * ncomb));
    if (nJ) {
      int grid;
      DDF_CHECK(grid_for(nJ * ncomb, 256, &grid));
      k_i32_to_i64p1<<<grid, 256, 0, ctx->stream>>>(down.p, nJ * ncomb, r64.p);
      DDF_LAUNCH_CHECK(ctx);
      DDF_CUDA(cudaMemcpyAsync(rowval, r64.p, (size_t)nJ * ncomb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t j = 0; j <= nJ; j++) colptr[j] = 1 + j * ncomb;
    return 0;
  }
  // Ri > Rj: transpose of lookup(Rj, Ri) (:494-496): column j = the Ri-simplices that contain Rj-simplex j, ascending
  {
    int64_t c = 1;
    for (int q = 0; q < Rj + 1; q++) c = c * (Ri + 1 - q) / (q + 1);
    *nnz = nI * c;
    if (!colptr) return 0;
    int ncomb = 0;
    DevBuf<int32_t> down;
    DDF_CHECK(lookup_down(m, Rj, Ri, down, &ncomb));     // n_Ri x ncomb ids of Rj-simplices
    const int64_t n = nI * ncomb;
    DevBuf<unsigned long long> keys, keys2, cnt;
    DevBuf<uint8_t> temp;
    DevBuf<int64_t> r64;
    DDF_CHECK(keys.alloc(ctx, n));
    DDF_CHECK(keys2.alloc(ctx, n));
    DDF_CHECK(cnt.alloc(ctx, nJ + 1));
    DDF_CHECK(r64.alloc(ctx, n));
    DDF_CUDA(cudaMemsetAsync(cnt.p, 0, (nJ + 1) * 8, ctx->stream));
    if (n) {
      int grid;
      DDF_CHECK(grid_for(n, 256, &grid));
      k_pairs_keys<<<grid, 256, 0, ctx->stream>>>(down.p, nI, ncomb, keys.p);
      DDF_LAUNCH_CHECK(ctx);
      size_t tb = 0;
      DDF_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys.p, keys2.p, n, 0, 64, ctx->stream));
      DDF_CHECK(temp.alloc(ctx, tb));
      DDF_CUDA(cub::DeviceRadixSort::SortKeys(temp.p, tb, keys.p, keys2.p, n, 0, 64, ctx->stream));
      ctx->launches++;
      k_keys_split<<<grid, 256, 0, ctx->stream>>>(keys2.p, n, nJ, r64.p, cnt.p);
      DDF_LAUNCH_CHECK(ctx);
      DDF_CUDA(cudaMemcpyAsync(rowval, r64.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    std::vector<unsigned long long> hc(nJ + 1);
    DDF_CUDA(cudaMemcpyAsync(hc.data(), cnt.p, (nJ + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
    colptr[0] = 1;
    for (int64_t j = 0; j < nJ; j++) colptr[j + 1] = colptr[j] + (int64_t)hc[j + 1];
    return 0;
  }
  DDF_TRY_END
}

// refine_coords (Meshing.jl:127-146): out[0..V) = coords, out[V + e] = (x_a + x_b) / 2 for edge e = (a, b), a < b --
// the reference sums the two vertices of the sparse column in ascending order and divides by 2
template <int C>
__global__ void k_refine_coords(const double* __restrict__ coords, int64_t nv, const int32_t* __restrict__ edges, int64_t ne,
                                double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv + ne) return;
  if (i < nv) {
#pragma unroll
    for (int c = 0; c < C; c++) out[i * C + c] = coords[i * C + c];
  } else {
    const int64_t e = i - nv;
    const int32_t a = edges[2 * e], b = edges[2 * e + 1];
#pragma unroll
    for (int c = 0; c < C; c++) out[i * C + c] = __ddiv_rn(__dadd_rn(coords[(size_t)a * C + c], coords[(size_t)b * C + c]), 2.0);
  }
}

extern "C" int ddf_mesh_refine_coords(ddf_mesh* m, double* out /* (n_0 + n_1) * C */) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_assembled(m));
  if (m->D < 1) DDF_FAIL("refine_coords: needs D >= 1");
  ddf_ctx* ctx = m->ctx;
  DDF_CUDA(cudaSetDevice(ctx->device));
  const int64_t nv = m->n[0], ne = m->n[1], tot = nv + ne;
  if (!tot || !m->C) return 0;
  DevBuf<double> buf;
  DDF_CHECK(buf.alloc(ctx, (size_t)tot * m->C));
  int grid;
  DDF_CHECK(grid_for(tot, 256, &grid));
  switch (m->C) {
    case 1: k_refine_coords<1><<<grid, 256, 0, ctx->stream>>>(m->coords.p, nv, m->simp[1].p, ne, buf.p); break;
    case 2: k_refine_coords<2><<<grid, 256, 0, ctx->stream>>>(m->coords.p, nv, m->simp[1].p, ne, buf.p); break;
    case 3: k_refine_coords<3><<<grid, 256, 0, ctx->stream>>>(m->coords.p, nv, m->simp[1].p, ne, buf.p); break;
    default: DDF_FAIL("refine_coords: C = %d not supported (C <= 3)", m->C);
  }
  DDF_LAUNCH_CHECK(ctx);
  DDF_CUDA(cudaMemcpyAsync(out, buf.p, (size_t)tot * m->C * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
  DDF_TRY_END
}

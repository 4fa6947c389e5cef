// wave.cu -- K6: the vertex Laplacian L = laplace(Val(Pr),Val(0)) = delta_1 d_0
// (src/ManifoldOps.jl:128-144 with :110-121) and the scalar wave step built on it
// (examples/wave.jl:82-100).
//
// The reference assembles L by two diagonal scalings and one SpGEMM
// (Ops.jl:228-232), chopping after every product (Ops.jl:44-58):
//     delta[i,e] = fl((-ih0[i] * s_ie) * star1[e]),   L[i,j] = delta[i,e] * d0[e,j],
//     L[i,i]    = sequential sum over incident edges e ascending of delta[i,e]*d0[e,i].
// k_build_laplace0 reproduces exactly those roundings into a row-CSR (adj_ptr,
// adj_nbr, adj_l) with the diagonal entry stored inline at its sorted position and
// chopped entries removed.  Because edge ids are lexicographic ranks of (a,b), the
// incident edges of vertex i in id order are its neighbours in ascending vertex
// order -- the reference's CSC `mul!` (zero y; scatter by ascending column) is
// therefore the per-row sequential sum over the stored entries, with separately
// rounded multiply and add.  All kernels below sum in that order, so L*u, the wave
// right-hand side and the leapfrog step are bit-exact with the oracle.
//
// Row kernels are warp-cooperative ("row streaming"): a warp owns 32 consecutive
// rows; their entries are one contiguous range of the CSR arrays, loaded with fully
// coalesced 128/256-byte requests, every lane issues its u-gathers back to back
// (14 independent gathers per lane on the 3D Kuhn mesh), the products go to shared
// memory and each lane then adds up its own row in order.
#include <algorithm>
#include <cub/cub.cuh>

#include "mesh.cuh"
#include "sell.cuh"
#include "stage.cuh"

int ddf_mesh_require_star(ddf_mesh* m, int R);   // ops.cu
int ddf_reduce(ddf_ctx* c, int mode, const double* x, const double* y, int64_t n, double* out);   // core.cu

// windows per pipeline group the builder tries first for short rows (2D meshes); 0 = SELL with 16-bit offsets
#define DDF_STAGE_K_SHORT 2

#ifndef DDF_PIPE_UNR
#define DDF_PIPE_UNR 4          // unroll factor of the consumers' row sum
#endif
#ifndef PIPE_PF
#define PIPE_PF 2               // L2 prefetch distance (iterations) of the entries in the Tsit5 stage kernels; 0 = off.
                                // Measured at C4 (paired stages): 0 / 2 / 3 / 5 -> 2.27 / 2.05 / 2.06 / 2.12 ms per step; the
                                // same prefetch in the HBM-bound one-step leapfrog kernel LOSES (0.482 -> 0.546 ms): Tsit5 only
#endif
#define PIPE_GROUP 256          // threads of one consumer group (= rows of a window)
#define PIPE_MAX_STAGES 6
// geometry of one stage of the persistent pipeline (k_rows_pipe below)
struct PipeGeom {
  uint32_t o_ent, o_v, o_u, o_meta, stage_bytes;   // byte offsets inside a stage (x16): entries records, v, u, meta
  uint32_t o_run2;                                 // MODE 5: the staged runs of the SECOND gathered vector (0 otherwise)
  int nstages;
  int K;                                           // windows per group (= per stage)
  int64_t nrows;
};

// pair: two gathered vectors per stage and no v / u slices (the paired Tsit5 stages, MODE 5)
static int pipe_geometry(const ddf_mesh* m, PipeGeom* g, bool pair = false) {
  const StageBufs& sb = m->lstage;
  auto up = [](uint32_t x, uint32_t a) { return (x + a - 1) / a * a; };
  g->K = sb.K;
  g->o_run2 = pair ? up(sb.max_slots * 8u, 16) : 0u;
  g->o_ent = (pair ? 2u : 1u) * up(sb.max_slots * 8u, 16);
  g->o_v = g->o_ent + up(sb.max_entries * 10u, 16);
  g->o_u = g->o_v + (pair ? 0u : (uint32_t)sb.K * DDF_SELL_WIN * 8);
  g->o_meta = g->o_u + (pair ? 0u : (uint32_t)sb.K * DDF_SELL_WIN * 8);
  g->stage_bytes = up(g->o_meta + (uint32_t)sb.K * STAGE_META, 128);
  const uint32_t budget = 227u * 1024u - 4096u;
  g->nstages = (int)(budget / g->stage_bytes);
  if (g->nstages > PIPE_MAX_STAGES) g->nstages = PIPE_MAX_STAGES;
  g->nrows = m->n[0];
  return g->nstages >= 2;
}

// pass 1 (count) / pass 2 (fill) of the assembly of L rows
template <bool FILL>
__global__ void k_build_laplace0(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ colw,
                                 const int32_t* __restrict__ edges, const double* __restrict__ ih0,
                                 const double* __restrict__ star1, int64_t nv, uint32_t* __restrict__ counts,
                                 const uint32_t* __restrict__ adj_ptr, int32_t* __restrict__ nbr,
                                 double* __restrict__ lval) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  const double ih = ih0[i];
  const bool ih_kept = !(ih * ih < DDF_CHOP_ABS2);   // chop of the (bitsign * invhodge) * deriv entries (+-ih)
  // the diagonal: first touch assigns, later touches add (SparseArrays spmatmul), edges ascending
  double diag = 0.0;
  bool first = true;
  uint32_t kept = 0;
  for (uint32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
    const uint32_t e = colw[p] & 0x7fffffffu;
    double l = ih_kept ? __dmul_rn(ih, star1[e]) : 0.0;     // |delta[i,e]| = L[i,j]
    if (l * l < DDF_CHOP_ABS2) l = 0.0;                     // chop(delta) and chop(L) act on the same magnitude
    if (l != 0.0) {
      diag = first ? -l : __dadd_rn(diag, -l);
      first = false;
      kept++;
    }
  }
  const bool diag_kept = !first && !(diag * diag < DDF_CHOP_ABS2);
  if (!FILL) {
    counts[i] = kept + (diag_kept ? 1u : 0u);
    return;
  }
  uint32_t out = adj_ptr[i];
  bool diag_done = !diag_kept;
  for (uint32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
    const uint32_t cw = colw[p];
    const uint32_t e = cw & 0x7fffffffu;
    const bool i_is_first = (cw >> 31) != 0;                // d_1[a,e] = -1
    const int32_t j = i_is_first ? edges[2 * (size_t)e + 1] : edges[2 * (size_t)e];
    double l = ih_kept ? __dmul_rn(ih, star1[e]) : 0.0;
    if (l * l < DDF_CHOP_ABS2) l = 0.0;
    if (!diag_done && j > i) { nbr[out] = (int32_t)i; lval[out] = diag; out++; diag_done = true; }
    if (l != 0.0) { nbr[out] = j; lval[out] = l; out++; }
  }
  if (!diag_done) { nbr[out] = (int32_t)i; lval[out] = diag; }
}

int ddf_mesh_require_wave(ddf_mesh* m) {
  if (m->has_wave) return 0;
  if (m->D < 1) DDF_FAIL("laplace(Pr,0): needs D >= 1");
  DDF_CHECK(ddf_mesh_require_star(m, 0));
  DDF_CHECK(ddf_mesh_require_star(m, 1));
  DDF_CHECK(ddf_mesh_require_isboundary(m));
  ddf_ctx* ctx = m->ctx;
  const int64_t nv = m->n[0];
  DDF_CHECK(m->adj_ptr.alloc(ctx, nv + 1));
  if (nv == 0) {
    DDF_CUDA(cudaMemsetAsync(m->adj_ptr.p, 0, 4, ctx->stream));
    m->adj_nnz = 0;
    m->has_wave = true;
    return 0;
  }
  int grid;
  DDF_CHECK(grid_for(nv, 256, &grid));
  DevBuf<uint32_t> counts;
  DevBuf<uint8_t> temp;
  DDF_CHECK(counts.alloc(ctx, nv + 1));
  DDF_CUDA(cudaMemsetAsync(counts.p + nv, 0, 4, ctx->stream));
  k_build_laplace0<false><<<grid, 256, 0, ctx->stream>>>(m->brow_ptr[1].p, m->bcol[1].p, m->simp[1].p, m->istar[0].p,
                                                         m->star[1].p, nv, counts.p, nullptr, nullptr, nullptr);
  DDF_LAUNCH_CHECK(ctx);
  size_t tb = 0;
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.p, m->adj_ptr.p, nv + 1, ctx->stream));
  DDF_CHECK(temp.alloc(ctx, tb));
  DDF_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, counts.p, m->adj_ptr.p, nv + 1, ctx->stream));
  ctx->launches++;
  uint32_t nnz = 0;
  DDF_CUDA(cudaMemcpyAsync(&nnz, m->adj_ptr.p + nv, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  m->adj_nnz = nnz;
  DDF_CHECK(m->adj_nbr.alloc(ctx, nnz));
  DDF_CHECK(m->adj_l.alloc(ctx, nnz));
  k_build_laplace0<true><<<grid, 256, 0, ctx->stream>>>(m->brow_ptr[1].p, m->bcol[1].p, m->simp[1].p, m->istar[0].p,
                                                        m->star[1].p, nv, nullptr, m->adj_ptr.p, m->adj_nbr.p,
                                                        m->adj_l.p);
  DDF_LAUNCH_CHECK(ctx);
  // Two storage candidates for the row kernels:
  //  * staged rows (stage.cuh): 10 B per entry, no padding, everything streamed by the persistent TMA pipeline.  Its
  //    cost is ~1 us per 256-row WINDOW (bulk-request rate), so it pays when a window carries enough bytes: rows of
  //    >= 10 entries on average (3D meshes).  Tuning bit 16 disables it.
  //  * SELL-32-256 with explicit indices -- as 16-bit offsets column - row (10 B per entry) when the matrix is
  //    banded enough, else 32-bit (12 B per entry).  Short rows (2D meshes: 7 entries) are faster here
  //    (profiles/r01_tuning.md).  Tuning bit 2048 forces 32-bit indices.
  const bool long_rows = m->adj_nnz >= 10 * nv;
  const int tune_wave = ctx->tune.wave;
  // tuning bits 13-14: windows per pipeline group for SHORT rows (0 = use SELL); bit 4096 = legacy "force staged, K = 1"
  const int k_short = (tune_wave & 4096) ? 1 : ((tune_wave >> 13) & 3) ? ((tune_wave >> 13) & 3) + 1 : DDF_STAGE_K_SHORT;
  if (!(tune_wave & 16) && (long_rows || k_short > 0)) {
    for (int K = long_rows ? 1 : k_short; K >= 1; K--) {
      DDF_CHECK(stage_build(ctx, m->adj_ptr.p, m->adj_nbr.p, m->adj_l.p, nv, m->adj_nnz, m->isbnd[0].p, K, m->lstage));
      if (m->lstage.ok && K > 1) {                 // a group must leave room for >= 3 stages
        PipeGeom g;
        if (!pipe_geometry(m, &g) || g.nstages < 3) m->lstage.release();
      }
      if (m->lstage.ok) {
        m->has_wave = true;
        return 0;
      }
    }
  }
  DDF_CHECK(sell_build(ctx, m->adj_ptr.p, reinterpret_cast<const uint32_t*>(m->adj_nbr.p), m->adj_l.p, nv, m->lsell,
                       (tune_wave & 4) != 0, (tune_wave & 8) != 0, !(tune_wave & 2048)));
  m->wave_sell = m->lsell.padded <= m->adj_nnz + m->adj_nnz / 8;
  if (!m->wave_sell) m->lsell.release();
  m->has_wave = true;
  return 0;
}

// ---------------------------------------------------------------- row kernels
// MODE 0: y = L u            MODE 1: wave rhs (du, dudot)          MODE 2: leapfrog step
struct RowArgs {
  const uint32_t* rowptr;
  const int32_t* nbr;
  const double* lval;
  const uint8_t* isb;
  const double* u;
  const double* udot;   // MODE 1 input
  double* out0;         // MODE 0: y ; MODE 1: du    ; MODE 2: v (in/out)
  double* out1;         //            MODE 1: dudot ; MODE 2: unew
  double dt;
  int64_t row0, row1;
  // fused halo (comm.cu, MODE 2): the new u of the interface planes is ALSO stored straight into the
  // neighbours' ghost planes through NVLink peer mappings -- rows [row0, send_lo_end) to peer_lo[i - row0],
  // rows [send_hi_beg, row1) to peer_hi[i - send_hi_beg]
  double* peer_lo;
  double* peer_hi;
  int64_t send_lo_end, send_hi_beg;
  // MODE 3 (fused Poisson relaxation): u' = u + theta * ((1-b) .* ((rho - L u) .* dinv)),  dinv = 1 ./ diag(L)
  const double* rho;
  const double* dinv;
  double theta;
  // MODE 4 (one fused Tsit5 stage, ddf_wave_tsit5): the row sum is L * su_J (a.u = su_J); the epilogue forms
  // k^v_J = (1-b) .* (L su_J) and the NEXT stage input su_{J+1} -- and, for J = 6, the new udot -- pointwise from the
  // step's base state and the stored k^v_1 .. k^v_{J-1} (see tsit5_stage below)
  const double* kv[5];
  const double* base_u;
  const double* base_v;
  double* out_kv;       // k^v_J (null for J = 6: nobody reads it)
  double* out_su;       // su_{J+1}; for J = 6 the new u (may be base_u: every thread reads its own row first)
  double* out_sv;       // J = 6: the new udot (may be base_v)
  int J;
  // MODE 5 (TWO Tsit5 stages per pass over L, tsit5_pair below): J = 1, 3, 5; a.u = su_J and u2 = su_{J+1} are gathered
  // with the same entries; out_kv / out_kv2 = k^v_J / k^v_{J+1} (J < 5), out_su / out_su2 = su_{J+2} / su_{J+3};
  // for J = 5: out_su = new u (may be base_u), out_sv = new udot (may be base_v), out_su2 = su_2 of the NEXT step
  const double* u2;
  double* out_kv2;
  double* out_su2;
  double* peer_lo2;     // fused halo of the SECOND stored stage input (out_su2), like peer_lo / peer_hi for out_su
  double* peer_hi2;
  // two leapfrog steps per launch on a slab (wavefront2.cuh, deep halo): the step-n tasks cover rows [row0, row1) = the
  // owned rows + one ghost plane on either side, the step-(n+1) tasks the owned rows [row0b, row1b) and store the two
  // outermost owned planes of the new u (peer_lo / peer_lo2 towards rank-1, peer_hi / peer_hi2 towards rank+1; the
  // plane next to the interface first) and the outermost plane of the new v (peer_lo3 / peer_hi3) into the neighbours
  int64_t row0b, row1b, plane;
  double* peer_lo3;
  double* peer_hi3;
};

// Tsitouras 5(4) tableau (OrdinaryDiffEq's Tsit5ConstantCache; pinned by its order conditions in tests/test_oracle_pins.py)
__constant__ double c_tsit5_a[6][6] = {
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};

// One row of a fused Tsit5 stage.  The unfused sequence (six k_lincomb passes + six wave right-hand sides per step)
// computes, per stage l, the input (su_l, sv_l) = U + dt * (a_l1 k_1 + ... ) with every product and sum rounded
// separately, left to right, and k_l = ((1-b) .* sv_l, (1-b) .* (L su_l)).  Only k^v_l = (1-b) .* (L su_l) needs other
// rows; k^u_l = m * sv_l and sv_l itself are functions of THIS row's base state and k^v_1 .. k^v_{l-1}.  So a stage
// kernel recomputes them in registers with the same operations in the same order -- identical bits, and the 2 x 6
// lincomb passes over up to 8 vectors disappear (the stored state per step shrinks from 16 vectors to 7).
template <int J>
__device__ __forceinline__ void tsit5_stage(double m, double lu, double (&kv)[6], double u0, double v0, double dt,
                                            double& kv_new, double& su_next, double& sv_next) {
  kv[J - 1] = kv_new = __dmul_rn(m, lu);
  double ku[6];
  ku[0] = __dmul_rn(m, v0);
#pragma unroll
  for (int l = 2; l <= J; l++) {
    double s = __dmul_rn(c_tsit5_a[l - 2][0], kv[0]);
#pragma unroll
    for (int q = 1; q <= l - 2; q++) s = __dadd_rn(s, __dmul_rn(c_tsit5_a[l - 2][q], kv[q]));
    ku[l - 1] = __dmul_rn(m, __dadd_rn(v0, __dmul_rn(dt, s)));
  }
  double s = __dmul_rn(c_tsit5_a[J - 1][0], ku[0]);
#pragma unroll
  for (int q = 1; q <= J - 1; q++) s = __dadd_rn(s, __dmul_rn(c_tsit5_a[J - 1][q], ku[q]));
  su_next = __dadd_rn(u0, __dmul_rn(dt, s));
  sv_next = 0.0;
  if (J == 6) {
    double t = __dmul_rn(c_tsit5_a[5][0], kv[0]);
#pragma unroll
    for (int q = 1; q <= 5; q++) t = __dadd_rn(t, __dmul_rn(c_tsit5_a[5][q], kv[q]));
    sv_next = __dadd_rn(v0, __dmul_rn(dt, t));
  }
}

// operands of row i requested up front (their latency overlaps the row sum), finished after it
struct Tsit5Row {
  double kv[6], u0, v0;
};
__device__ __forceinline__ void tsit5_load(const RowArgs& a, int64_t i, Tsit5Row& r) {
#pragma unroll
  for (int q = 0; q < 5; q++) r.kv[q] = q < a.J - 1 ? __ldg(a.kv[q] + i) : 0.0;
  r.kv[5] = 0.0;
  r.u0 = a.base_u[i];       // plain loads: for J = 6 these vectors are also written by this kernel (own row only)
  r.v0 = a.base_v[i];
}
__device__ __forceinline__ void tsit5_finish(const RowArgs& a, int64_t i, double m, double lu, Tsit5Row& r) {
  double kn, su, sv;
  switch (a.J) {
    case 1: tsit5_stage<1>(m, lu, r.kv, r.u0, r.v0, a.dt, kn, su, sv); break;
    case 2: tsit5_stage<2>(m, lu, r.kv, r.u0, r.v0, a.dt, kn, su, sv); break;
    case 3: tsit5_stage<3>(m, lu, r.kv, r.u0, r.v0, a.dt, kn, su, sv); break;
    case 4: tsit5_stage<4>(m, lu, r.kv, r.u0, r.v0, a.dt, kn, su, sv); break;
    case 5: tsit5_stage<5>(m, lu, r.kv, r.u0, r.v0, a.dt, kn, su, sv); break;
    default: tsit5_stage<6>(m, lu, r.kv, r.u0, r.v0, a.dt, kn, su, sv); break;
  }
  if (a.out_kv) a.out_kv[i] = kn;
  a.out_su[i] = su;
  if (a.J == 6) a.out_sv[i] = sv;
}

// TWO stages per pass over L.  In dU = F U with F = [0, M; M L, 0] the stage input su_l depends on k^u_q = M sv_q
// (q < l) only, and sv_q on k^v_r (r < q): su_l needs k^v_1 .. k^v_{l-2}, NOT k^v_{l-1}.  So su_J and su_{J+1} are both
// known before stage J runs, the two row sums L su_J and L su_{J+1} share one read of the entries of L, and a Tsit5 step
// is THREE passes over L instead of six: (1,2) -> su_3, su_4;  (3,4) -> su_5, su_6;  (5,6) -> the new state and su_2 of
// the next step.  Every stage input / derivative is formed by the same operations in the same order as the unfused
// sequence: same bits.
template <int J>
__device__ __forceinline__ void tsit5_pair(double m, double lu_a, double lu_b, double (&kv)[6], double u0, double v0, double dt,
                                           double& kv_a, double& kv_b, double& o1, double& o2, double& o3) {
  kv[J - 1] = kv_a = __dmul_rn(m, lu_a);
  kv[J] = kv_b = __dmul_rn(m, lu_b);
  constexpr int LMAX = J + 2 < 6 ? J + 2 : 6;           // k^u_1 .. k^u_LMAX are needed
  double ku[6];
  ku[0] = __dmul_rn(m, v0);
#pragma unroll
  for (int l = 2; l <= LMAX; l++) {
    double s = __dmul_rn(c_tsit5_a[l - 2][0], kv[0]);
#pragma unroll
    for (int q = 1; q <= l - 2; q++) s = __dadd_rn(s, __dmul_rn(c_tsit5_a[l - 2][q], kv[q]));
    ku[l - 1] = __dmul_rn(m, __dadd_rn(v0, __dmul_rn(dt, s)));
  }
  // su_l = u0 + dt * (a_{l,1} k^u_1 + ... + a_{l,l-1} k^u_{l-1}),  tableau row l - 2
  {
    constexpr int l = J + 2;
    double s = __dmul_rn(c_tsit5_a[l - 2][0], ku[0]);
#pragma unroll
    for (int q = 1; q <= l - 2; q++) s = __dadd_rn(s, __dmul_rn(c_tsit5_a[l - 2][q], ku[q]));
    o1 = __dadd_rn(u0, __dmul_rn(dt, s));
  }
  o3 = 0.0;
  if (J < 5) {
    constexpr int l = J < 5 ? J + 3 : 7;
    double s = __dmul_rn(c_tsit5_a[l - 2][0], ku[0]);
#pragma unroll
    for (int q = 1; q <= l - 2; q++) s = __dadd_rn(s, __dmul_rn(c_tsit5_a[l - 2][q], ku[q]));
    o2 = __dadd_rn(u0, __dmul_rn(dt, s));
  } else {
    // o1 is the new u; o2 = the new udot; o3 = su_2 of the next step = u' + dt * (a_21 * (m * udot'))
    double t = __dmul_rn(c_tsit5_a[5][0], kv[0]);
#pragma unroll
    for (int q = 1; q <= 5; q++) t = __dadd_rn(t, __dmul_rn(c_tsit5_a[5][q], kv[q]));
    o2 = __dadd_rn(v0, __dmul_rn(dt, t));
    o3 = __dadd_rn(o1, __dmul_rn(dt, __dmul_rn(c_tsit5_a[0][0], __dmul_rn(m, o2))));
  }
}
// fused halo (comm.cu): interface-plane rows are ALSO stored into the neighbours' landing arenas over NVLink
__device__ __forceinline__ void peer_store(const RowArgs& a, int64_t i, double w) {
  if (a.peer_lo && i < a.send_lo_end) a.peer_lo[i - a.row0] = w;
  if (a.peer_hi && i >= a.send_hi_beg) a.peer_hi[i - a.send_hi_beg] = w;
}
__device__ __forceinline__ void peer_store2(const RowArgs& a, int64_t i, double w) {
  if (a.peer_lo2 && i < a.send_lo_end) a.peer_lo2[i - a.row0] = w;
  if (a.peer_hi2 && i >= a.send_hi_beg) a.peer_hi2[i - a.send_hi_beg] = w;
}
__device__ __forceinline__ void tsit5_pair_finish(const RowArgs& a, int64_t i, double m, double lu_a, double lu_b, Tsit5Row& r) {
  double ka, kb, o1, o2, o3;
  if (a.J == 1) tsit5_pair<1>(m, lu_a, lu_b, r.kv, r.u0, r.v0, a.dt, ka, kb, o1, o2, o3);
  else if (a.J == 3) tsit5_pair<3>(m, lu_a, lu_b, r.kv, r.u0, r.v0, a.dt, ka, kb, o1, o2, o3);
  else tsit5_pair<5>(m, lu_a, lu_b, r.kv, r.u0, r.v0, a.dt, ka, kb, o1, o2, o3);
  if (a.J < 5) {
    a.out_kv[i] = ka;
    a.out_kv2[i] = kb;
    a.out_su[i] = o1;
    a.out_su2[i] = o2;
    peer_store(a, i, o1);
    peer_store2(a, i, o2);
  } else {
    a.out_su[i] = o1;
    a.out_sv[i] = o2;
    a.out_su2[i] = o3;
    peer_store(a, i, o1);
    peer_store2(a, i, o3);
  }
}
// su_2 of the first step: u + dt * (a_21 * ((1-b) .* udot))
__global__ void k_tsit5_first(const double* __restrict__ u, const double* __restrict__ udot, const uint8_t* __restrict__ isb,
                              double dt, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = isb[i] ? 0.0 : 1.0;
  out[i] = __dadd_rn(u[i], __dmul_rn(dt, __dmul_rn(c_tsit5_a[0][0], __dmul_rn(m, udot[i]))));
}

// the epilogue of MODE 3, shared by every row kernel: one rounding per operation, in this order
__device__ __forceinline__ double jacobi_update(double u_i, double lu, double rho_i, double dinv_i, double m, double theta) {
  const double c = __dmul_rn(__dsub_rn(rho_i, lu), dinv_i);
  return __dadd_rn(u_i, __dmul_rn(theta, __dmul_rn(m, c)));
}


template <int MODE>
__device__ __forceinline__ void row_epilogue(const RowArgs& a, int64_t i, double lu, uint64_t pol_stream) {
  if (MODE == 0) {
    st_hint_f64(a.out0 + i, lu, pol_stream);
  } else if (MODE == 1) {
    const double m = a.isb[i] ? 0.0 : 1.0;
    a.out0[i] = __dmul_rn(m, a.udot[i]);
    a.out1[i] = __dmul_rn(m, lu);
  } else if (MODE == 3) {
    const double m = a.isb[i] ? 0.0 : 1.0;
    const double w = jacobi_update(a.u[i], lu, a.rho[i], a.dinv[i], m, a.theta);
    a.out1[i] = w;
    peer_store(a, i, w);
  } else if (MODE == 4) {
    Tsit5Row r;
    tsit5_load(a, i, r);
    tsit5_finish(a, i, a.isb[i] ? 0.0 : 1.0, lu, r);
  } else {
    const double m = a.isb[i] ? 0.0 : 1.0;
    const double vn = __dadd_rn(a.out0[i], __dmul_rn(a.dt, __dmul_rn(m, lu)));
    a.out0[i] = vn;
    const double w = __dadd_rn(a.u[i], __dmul_rn(a.dt, __dmul_rn(m, vn)));
    a.out1[i] = w;
    peer_store(a, i, w);
  }
}

// variant 1: one thread per row (baseline)
template <int MODE>
__global__ void __launch_bounds__(256) k_rows_thread(RowArgs a) {
  const int64_t i = a.row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.row1) return;
  const uint64_t pol_stream = make_policy_evict_first();
  const uint64_t pol_keep = make_policy_evict_last();
  double acc = 0.0;
  const uint32_t lo = a.rowptr[i], hi = a.rowptr[i + 1];
  // batches of 8 entries: 8 index + 8 value loads, then 8 gathers in flight, then the ordered sum
  for (uint32_t p = lo; p < hi; p += 8) {
    int j[8];
    double l[8], uu[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const bool ok = p + q < hi;
      j[q] = ok ? __ldg(a.nbr + p + q) : (int)i;
      l[q] = ok ? __ldg(a.lval + p + q) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) uu[q] = ld_keep_f64(a.u + j[q], pol_keep);
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (p + q < hi) acc = __dadd_rn(acc, __dmul_rn(l[q], uu[q]));
  }
  row_epilogue<MODE>(a, i, acc, pol_stream);
}

// SELL-32: one thread per row, coalesced index/value loads, batches of 8 gathers in flight
// COMPRESSED: indices come from the per-slice stream (uniform-offset columns hold one word)
template <int MODE, bool COMPRESSED, bool IDX16 = false>
__global__ void __launch_bounds__(256)
k_rows_sell(RowArgs a, const uint32_t* __restrict__ sell_ptr, const uint8_t* __restrict__ perm8,
            const uint32_t* __restrict__ snbr, const double* __restrict__ sl, const uint32_t* __restrict__ umask,
            const uint32_t* __restrict__ iptr, int64_t first_window) {
  const int64_t win = first_window + blockIdx.x;                            // one block = one sorted window
  const int64_t r = win * DDF_SELL_WIN + threadIdx.x;                        // sorted position
  const int64_t i = win * DDF_SELL_WIN + perm8[r];                           // the row it holds
  const uint64_t pol_stream = make_policy_evict_first();
  const uint64_t pol_keep = make_policy_evict_last();
  const int64_t s = r >> 5;
  const uint32_t lane = (uint32_t)(r & 31);
  const uint32_t base = sell_ptr[s] + lane;
  const uint32_t w = (sell_ptr[s + 1] - sell_ptr[s]) >> 5;
  const uint32_t um = COMPRESSED ? umask[s] : 0u;
  const uint32_t ib = COMPRESSED ? iptr[s] : 0u;
  double acc = 0.0;
  for (uint32_t q0 = 0; q0 < w; q0 += 8) {
    uint32_t j[8];
    double l[8], uu[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const uint32_t c = q0 + q;
      const bool ok = c < w;                     // warp-uniform
      const size_t pos = (size_t)base + (size_t)c * 32;
      if (COMPRESSED) {
        const bool uni = c < 32 && ((um >> c) & 1u);                      // warp-uniform
        const uint32_t word = ok ? (uint32_t)ld_stream_int(reinterpret_cast<const int*>(snbr) + ib +
                                                           sell_col_offset(um, c) + (uni ? 0u : lane))
                                 : DDF_SELL_PAD;
        j[q] = (ok && uni) ? (uint32_t)i + word : word;
      } else if (IDX16) {                        // 16-bit offsets column - row, 0x8000 = padding
        const uint16_t d = ok ? __ldg(reinterpret_cast<const uint16_t*>(snbr) + pos) : (uint16_t)0x8000u;
        j[q] = d != 0x8000u ? (uint32_t)((int64_t)i + (int16_t)d) : DDF_SELL_PAD;
      } else {
        j[q] = ok ? (uint32_t)ld_stream_int(reinterpret_cast<const int*>(snbr) + pos) : DDF_SELL_PAD;
      }
      l[q] = ok ? ld_stream_f64(sl + pos) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) uu[q] = ld_keep_f64(a.u + (j[q] != DDF_SELL_PAD ? j[q] : 0u), pol_keep);
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (j[q] != DDF_SELL_PAD) acc = __dadd_rn(acc, __dmul_rn(l[q], uu[q]));
  }
  if (i >= a.row0 && i < a.row1) row_epilogue<MODE>(a, i, acc, pol_stream);
}

// staged rows (stage.cuh): one block = one 256-row window.  Warp 0 issues the window's TMA bulk
// copies of u into shared memory; meanwhile every thread prefetches its first batch of
// (slot, value) pairs and its epilogue operands; the gather itself reads shared memory.
// A slice is walked in SEGMENTS of constant active-lane count (rows are sorted by decreasing
// length, so the active lanes are always a prefix): inside a segment a row's entries are
// nact apart -- no per-entry ballots.  Slices of equal-length rows (all but <= 2-3 per window)
// are a single segment of stride 32.
template <int BATCH>
__device__ __forceinline__ double stage_row_sum(const uint16_t* __restrict__ slot, const double* __restrict__ sl,
                                                const double* stage, uint64_t* mbar, uint32_t off, uint32_t lane,
                                                uint32_t mylen, uint32_t maxlen) {
  double acc = 0.0;
  bool staged = false;
  const uint16_t* ps = slot + (size_t)off + lane;
  const double* pv = sl + (size_t)off + lane;
  uint32_t q = 0;
  while (q < maxlen) {                                                     // warp-uniform control flow throughout
    const uint32_t nact = __popc(__ballot_sync(0xffffffffu, q < mylen));   // lanes [0, nact) still hold entries
    const uint32_t qend = __shfl_sync(0xffffffffu, mylen, nact - 1);       // the shortest of them ends the segment
    const bool act = lane < nact;
    for (; q < qend; q += BATCH) {
      uint32_t sidx[BATCH];
      double l[BATCH];
#pragma unroll
      for (int k = 0; k < BATCH; k++) {
        const bool ok = act && q + k < qend;
        sidx[k] = ok ? (uint32_t)__ldg(ps + k * nact) : 0xFFFFFFFFu;
        l[k] = ok ? ld_stream_f64(pv + k * nact) : 0.0;
      }
      if (!staged) { stage_wait(mbar, 0); staged = true; }
#pragma unroll
      for (int k = 0; k < BATCH; k++)
        if (sidx[k] != 0xFFFFFFFFu) acc = __dadd_rn(acc, __dmul_rn(l[k], stage[sidx[k]]));
      const uint32_t adv = min((uint32_t)BATCH, qend - q) * nact;
      ps += adv;
      pv += adv;
    }
    q = qend;
  }
  if (!staged) stage_wait(mbar, 0);       // never leave while bulk copies into this block's shared memory are in flight
  return acc;
}

template <int MODE, int BATCH>
__global__ void __launch_bounds__(DDF_SELL_WIN)
k_rows_stage(RowArgs a, const uint32_t* __restrict__ slice_ptr, const uint8_t* __restrict__ perm8,
             const uint8_t* __restrict__ len8, const uint32_t* __restrict__ tab, const unsigned char* __restrict__ packed,
             int64_t first_window) {
  extern __shared__ __align__(16) double stage[];
  __shared__ __align__(8) uint64_t mbar;
  const int64_t win = first_window + blockIdx.x;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) stage_issue(tab + win * STAGE_TAB, a.u, stage, &mbar, make_policy_evict_last());
  const int64_t r = win * DDF_SELL_WIN + threadIdx.x;                        // sorted position
  const int64_t i = win * DDF_SELL_WIN + perm8[r];                           // the row it holds
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t mylen = len8[r];
  const uint32_t maxlen = __shfl_sync(0xffffffffu, mylen, 0);                // rows are sorted by decreasing length
  // the window's record: values at byte 10*e0, slots behind them; entry offsets are relative to e0
  const uint32_t ent0 = __ldg(tab + win * STAGE_TAB + 2), stored = __ldg(tab + win * STAGE_TAB + 3);
  const double* sl = reinterpret_cast<const double*>(packed + (size_t)ent0 * 10);
  const uint16_t* slot = reinterpret_cast<const uint16_t*>(packed + (size_t)ent0 * 10 + (size_t)stored * 8);
  const uint32_t off = slice_ptr[r >> 5] - ent0;
  const bool mine = i >= a.row0 && i < a.row1;
  // epilogue operands, requested before the row sum so that their latency overlaps it
  double e0 = 0.0, e1 = 0.0;
  uint8_t eb = 0;
  if (MODE != 0 && mine) eb = a.isb[i];
  if (MODE == 1 && mine) e0 = a.udot[i];
  if (MODE == 2 && mine) { e0 = a.out0[i]; e1 = a.u[i]; }
  double e2 = 0.0;
  if (MODE == 3 && mine) { e0 = a.rho[i]; e1 = a.u[i]; e2 = a.dinv[i]; }
  const double lu = stage_row_sum<BATCH>(slot, sl, stage, &mbar, off, lane, mylen, maxlen);
  if (!mine) return;
  if (MODE == 0) {
    st_hint_f64(a.out0 + i, lu, make_policy_evict_first());
  } else if (MODE == 1) {
    const double m = eb ? 0.0 : 1.0;
    a.out0[i] = __dmul_rn(m, e0);
    a.out1[i] = __dmul_rn(m, lu);
  } else if (MODE == 3) {
    const double w = jacobi_update(e1, lu, e0, e2, eb ? 0.0 : 1.0, a.theta);
    a.out1[i] = w;
    peer_store(a, i, w);
  } else if (MODE == 4) {
    Tsit5Row r;
    tsit5_load(a, i, r);
    tsit5_finish(a, i, eb ? 0.0 : 1.0, lu, r);
  } else {
    const double m = eb ? 0.0 : 1.0;
    const double vn = __dadd_rn(e0, __dmul_rn(a.dt, __dmul_rn(m, lu)));
    a.out0[i] = vn;
    const double w = __dadd_rn(e1, __dmul_rn(a.dt, __dmul_rn(m, vn)));
    a.out1[i] = w;
    peer_store(a, i, w);
  }
}

// ---------------------------------------------------------------- staged rows, persistent TMA pipeline
// One block per SM walks windows blockIdx.x, blockIdx.x + gridDim.x, ...  A producer warp streams
// EVERYTHING a window needs into a ring of shared-memory stages with cp.async.bulk (completion
// on the stage's "full" mbarrier): the runs of u, the window's value and slot streams (contiguous:
// stage_build pads a window's entry range to 16-byte boundaries), its slice of v / udot, u, the
// boundary mask and the perm/len bytes.  Eight consumer warps compute a window out of shared
// memory only and release the stage through its "empty" mbarrier.  No register ever waits on
// HBM, so the bytes in flight per SM are (stages - 1) x ~50 KB regardless of occupancy.

#ifdef DDF_PIPE_TRACE
// block 0 records %globaltimer at: producer armed stage (0), consumer saw data (1), consumer released stage (2)
__device__ unsigned long long g_pipe_trace[8 * 1024];
#define PIPE_TRACE(kind, it)                                                        \
  do {                                                                              \
    if (blockIdx.x == 0 && (it) < 1024 && (threadIdx.x & 31) == 0) {                \
      unsigned long long t_;                                                        \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                        \
      if ((kind) != 0 && (threadIdx.x & 255) != 0) {} else g_pipe_trace[(kind) * 1024 + (it)] = t_; \
    }                                                                               \
  } while (0)
#else
#define PIPE_TRACE(kind, it)
#endif

// The row sum of the pipeline consumers: a row's entries live in the stage's record as jagged diagonals (segments of
// constant active-lane count, entries of a row `nact` apart); the adds are sequential in entry order.  Unrolled by 4
// with a peeled remainder: fully unrolled, predicated blocks of 8 or 16 entries (all loads of a block up front) were
// measured SLOWER at C4 (two steps per launch 0.454 -> 0.499 / 0.610 ms, profiles/r02_tuning.md).
template <int UNR>
__device__ __forceinline__ double pipe_row_sum(const double* __restrict__ valS, const uint16_t* __restrict__ slotS,
                                               const double* __restrict__ stage, uint32_t part, uint32_t lane,
                                               uint32_t mylen, uint32_t maxlen) {
  uint32_t e = part + lane, q = 0;
  double acc = 0.0;
  while (q < maxlen) {                                                   // segments of constant active-lane count
    const uint32_t nact = __popc(__ballot_sync(0xffffffffu, q < mylen));
    const uint32_t qend = __shfl_sync(0xffffffffu, mylen, nact - 1);
    if (lane < nact) {
#pragma unroll UNR
      for (uint32_t k = q; k < qend; k++, e += nact) acc = __dadd_rn(acc, __dmul_rn(valS[e], stage[slotS[e]]));
    }
    q = qend;
  }
  return acc;
}

// the same walk with TWO gathered vectors (stage and stage2 share the slots): two independent sums, each in entry order
template <int UNR>
__device__ __forceinline__ void pipe_row_sum2(const double* __restrict__ valS, const uint16_t* __restrict__ slotS,
                                              const double* __restrict__ stage, const double* __restrict__ stage2,
                                              uint32_t part, uint32_t lane, uint32_t mylen, uint32_t maxlen, double& acc_a,
                                              double& acc_b) {
  uint32_t e = part + lane, q = 0;
  double a0 = 0.0, a1 = 0.0;
  while (q < maxlen) {
    const uint32_t nact = __popc(__ballot_sync(0xffffffffu, q < mylen));
    const uint32_t qend = __shfl_sync(0xffffffffu, mylen, nact - 1);
    if (lane < nact) {
#pragma unroll UNR
      for (uint32_t k = q; k < qend; k++, e += nact) {
        const double v = valS[e];
        const uint32_t sl = slotS[e];
        a0 = __dadd_rn(a0, __dmul_rn(v, stage[sl]));
        a1 = __dadd_rn(a1, __dmul_rn(v, stage2[sl]));
      }
    }
    q = qend;
  }
  acc_a = a0;
  acc_b = a1;
}

template <int MODE, int NG>
__global__ void __launch_bounds__(NG * PIPE_GROUP + 32, 1)
k_rows_pipe(RowArgs a, const uint32_t* __restrict__ tab, const unsigned char* __restrict__ packed,
            const unsigned char* __restrict__ meta, int64_t w0, int64_t w1, PipeGeom g) {
  extern __shared__ __align__(128) unsigned char pipe_smem[];
  __shared__ __align__(8) uint64_t full[PIPE_MAX_STAGES], empty[PIPE_MAX_STAGES], tfull[PIPE_TD];
  __shared__ __align__(16) uint32_t tabS[PIPE_TD * STAGE_TAB];
  // issued[s] = 1 + the last iteration the producer has armed stage s for.  A consumer group must see its own
  // iteration here BEFORE it waits on full[s] by parity: with NG groups sharing the ring, the previous use of the
  // stage belongs to another group, and a parity wait entered a whole phase early would pass at once.
  __shared__ uint32_t issued[PIPE_MAX_STAGES];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid < PIPE_MAX_STAGES) issued[tid] = 0u;
  if (tid == 0) {
    for (int d = 0; d < PIPE_TD; d++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tfull[d])) : "memory");
    for (int s = 0; s < g.nstages; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(PIPE_GROUP / 32) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t stride = gridDim.x;
  if (warp == NG * PIPE_GROUP / 32) {
    // ------------------------------------------------------------ producer warp
    // The windows' run tables (256 B each) are themselves bulk-copied into a small ring PIPE_TD
    // iterations ahead, so the producer never waits on a global load.
    const uint64_t pol_keep = make_policy_evict_last(), pol_stream = make_policy_evict_first();
    int64_t win = w0 + blockIdx.x;
    if (lane == 0) {
      for (int d = 0; d < PIPE_TD; d++) {
        const int64_t w = win + d * stride;
        if (w < w1) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tfull[d])), "r"(STAGE_TAB * 4) : "memory");
          bulk_g2s(tabS + d * STAGE_TAB, tab + w * STAGE_TAB, STAGE_TAB * 4, &tfull[d], pol_stream);
        }
      }
    }
    for (int64_t it = 0; win < w1; it++, win += stride) {
      const int td = (int)(it % PIPE_TD);
      stage_wait(&tfull[td], (uint32_t)(it / PIPE_TD) & 1u);
      const uint32_t* wt = tabS + td * STAGE_TAB;
      const uint32_t nruns = wt[0], total = wt[1], e0 = wt[2], nent = wt[3];
      const uint32_t start = lane < STAGE_RMAX ? wt[4 + 2 * lane] : 0u, off = lane < STAGE_RMAX ? wt[5 + 2 * lane] : 0u;
      const uint32_t end = lane + 1 < nruns ? wt[5 + 2 * (lane + 1)] : total;
      __syncwarp();
      if (lane == 0 && win + PIPE_TD * stride < w1) {                 // refill this table slot
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tfull[td])), "r"(STAGE_TAB * 4) : "memory");
        bulk_g2s(tabS + td * STAGE_TAB, tab + (win + PIPE_TD * stride) * STAGE_TAB, STAGE_TAB * 4, &tfull[td], pol_stream);
      }
      if (PIPE_PF > 0 && (MODE == 4 || MODE == 5) && win + PIPE_PF * stride < w1) {
        // ask L2 for the entries + meta records of the window PIPE_PF iterations ahead (its table arrived long ago): the
        // fill of that stage then costs an L2 round trip instead of an HBM one -- it matters where few stages fit
        const int tp = (int)((it + PIPE_PF) % PIPE_TD);
        stage_wait(&tfull[tp], (uint32_t)((it + PIPE_PF) / PIPE_TD) & 1u);
        if (lane == 0) {
          const uint32_t pe0 = tabS[tp * STAGE_TAB + 2], pn = tabS[tp * STAGE_TAB + 3];
          if (pn) bulk_prefetch_l2(packed + (size_t)pe0 * 10, pn * 10u);
          bulk_prefetch_l2(meta + (win + PIPE_PF * stride) * g.K * STAGE_META, (uint32_t)g.K * STAGE_META);
        }
      }
      const int s = (int)(it % g.nstages);
      const uint32_t j = (uint32_t)(it / g.nstages);
      if (j > 0) stage_wait(&empty[s], (j - 1) & 1u);
      unsigned char* sb = pipe_smem + (size_t)s * g.stage_bytes;
      // `win` counts GROUPS of K consecutive 256-row windows here (K = 1 on 3D meshes)
      const int64_t r0 = win * g.K * DDF_SELL_WIN;
      const uint32_t nr = (uint32_t)min((int64_t)g.K * DDF_SELL_WIN, g.nrows - r0);
      const uint32_t vbytes = ((nr + 1u) & ~1u) * 8u;
      const uint32_t mbytes = (uint32_t)g.K * STAGE_META;
      // requests per group: the runs of u, ONE range of entries records, ONE range of meta records, v (or udot), u
      uint32_t tx = total * 8u + nent * 10u + mbytes;
      if (MODE == 5) tx += total * 8u;
      if (MODE == 1) tx += vbytes;
      if (MODE == 2 || MODE == 3) tx += 2u * vbytes;
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(tx) : "memory");
        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(&issued[s])), "r"((uint32_t)it + 1u) : "memory");
      }
      PIPE_TRACE(0, it);
      __syncwarp();
      if (lane < nruns) bulk_g2s(sb + (size_t)off * 8, a.u + start, (end - off) * 8u, &full[s], pol_keep);
      if (MODE == 5 && lane < nruns) bulk_g2s(sb + g.o_run2 + (size_t)off * 8, a.u2 + start, (end - off) * 8u, &full[s], pol_keep);
      if (lane == 31 && nent) bulk_g2s(sb + g.o_ent, packed + (size_t)e0 * 10, nent * 10u, &full[s], pol_stream);
      if (lane == 30) bulk_g2s(sb + g.o_meta, meta + win * g.K * STAGE_META, mbytes, &full[s], pol_stream);
      if (lane == 29 && MODE != 0 && MODE != 4 && MODE != 5) {
        bulk_g2s(sb + g.o_v, (MODE == 1 ? a.udot : (MODE == 3 ? a.rho : a.out0)) + r0, vbytes, &full[s], pol_stream);
        if (MODE == 2 || MODE == 3) bulk_g2s(sb + g.o_u, a.u + r0, vbytes, &full[s], pol_keep);
      }
    }
  } else {
    // ------------------------------------------------------------ consumer groups: thread = sorted row position;
    // group k takes this block's iterations k, k + NG, ...
    const uint64_t pol_stream = make_policy_evict_first();
    const uint32_t grp = tid / PIPE_GROUP;
    const uint32_t gtid = tid % PIPE_GROUP, gwarp = gtid >> 5;
    int64_t it = grp;
    for (int64_t win = w0 + blockIdx.x + grp * stride; win < w1; it += NG, win += NG * stride) {
      const int s = (int)(it % g.nstages);
      const uint32_t j = (uint32_t)(it / g.nstages);
      for (;;) {                          // the producer has armed stage s for THIS iteration
        uint32_t seen;
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&issued[s])) : "memory");
        if (seen >= (uint32_t)it + 1u) break;
      }
      stage_wait(&full[s], j & 1u);
      PIPE_TRACE(1, it);
      const unsigned char* sb = pipe_smem + (size_t)s * g.stage_bytes;
      const double* stage = reinterpret_cast<const double*>(sb);
      uint32_t rec = 0;                                        // byte offset of window kk's record in the stage
      for (int kk = 0; kk < g.K; kk++) {
        const unsigned char* metaS = sb + g.o_meta + kk * STAGE_META;
        const uint32_t nentS = *reinterpret_cast<const uint32_t*>(metaS);
        const double* valS = reinterpret_cast<const double*>(sb + g.o_ent + rec);
        const uint16_t* slotS = reinterpret_cast<const uint16_t*>(sb + g.o_ent + rec + (size_t)nentS * 8);
        rec += nentS * 10u;
        const uint8_t* lenS = metaS + STAGE_META_HDR + DDF_SELL_WIN;
        const uint32_t mylen = lenS[gtid];
        const uint32_t p = (metaS + STAGE_META_HDR)[gtid];
        double dinv_i = 0.0;                                 // MODE 3: requested now, needed after the row sum
        if (MODE == 3) {
          const int64_t ip = win * g.K * DDF_SELL_WIN + (int64_t)kk * DDF_SELL_WIN + p;
          if (ip >= a.row0 && ip < a.row1) dinv_i = __ldg(a.dinv + ip);
        }
        Tsit5Row trow;                                       // MODE 4: likewise
        if (MODE == 4 || MODE == 5) {
          const int64_t ip = win * g.K * DDF_SELL_WIN + (int64_t)kk * DDF_SELL_WIN + p;
          if (ip >= a.row0 && ip < a.row1) tsit5_load(a, ip, trow);
        }
        // entry offset of this slice inside the window's record: precomputed in the meta header
        const uint32_t part = reinterpret_cast<const uint32_t*>(metaS)[4 + gwarp];
        const uint32_t maxlen = __shfl_sync(0xffffffffu, mylen, 0);
        PIPE_TRACE(3, it);
        double acc, acc2 = 0.0;
        if (MODE == 5) pipe_row_sum2<DDF_PIPE_UNR>(valS, slotS, stage, reinterpret_cast<const double*>(sb + g.o_run2), part, lane, mylen, maxlen, acc, acc2);
        else acc = pipe_row_sum<DDF_PIPE_UNR>(valS, slotS, stage, part, lane, mylen, maxlen);
        PIPE_TRACE(4, it);
        const uint32_t pl = (uint32_t)kk * DDF_SELL_WIN + p;                  // row offset inside the group
        const int64_t i = win * g.K * DDF_SELL_WIN + pl;
        // what the epilogue needs from the stage goes to registers first; after the LAST window of the group the stage
        // is handed back to the producer BEFORE the epilogue (global loads consumed, arithmetic, stores) runs
        const double m = (MODE != 0 && (metaS + STAGE_META_HDR + 2 * DDF_SELL_WIN)[p]) ? 0.0 : 1.0;
        double e0 = 0.0, eu = 0.0;
        if (MODE == 1 || MODE == 2 || MODE == 3) e0 = reinterpret_cast<const double*>(sb + g.o_v)[pl];
        if (MODE == 2 || MODE == 3) eu = reinterpret_cast<const double*>(sb + g.o_u)[pl];
        if (kk == g.K - 1) {
          __syncwarp();
          PIPE_TRACE(2, it);
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        if (i >= a.row0 && i < a.row1) {
          if (MODE == 0) {
            st_hint_f64(a.out0 + i, acc, pol_stream);
          } else if (MODE == 4) {
            tsit5_finish(a, i, m, acc, trow);
          } else if (MODE == 5) {
            tsit5_pair_finish(a, i, m, acc, acc2, trow);
          } else if (MODE == 1) {
            a.out0[i] = __dmul_rn(m, e0);
            a.out1[i] = __dmul_rn(m, acc);
          } else if (MODE == 3) {
            const double w = jacobi_update(eu, acc, e0, dinv_i, m, a.theta);
            a.out1[i] = w;
            peer_store(a, i, w);
          } else {
            const double vn = __dadd_rn(e0, __dmul_rn(a.dt, __dmul_rn(m, acc)));
            a.out0[i] = vn;
            const double w = __dadd_rn(eu, __dmul_rn(a.dt, __dmul_rn(m, vn)));
            a.out1[i] = w;
            peer_store(a, i, w);
          }
        }
      }
    }
  }
}

#ifdef DDF_PIPE_TRACE
extern "C" __attribute__((visibility("default"))) int ddf_debug_pipe_trace(unsigned long long* out /* 8*1024 */) {
  return cudaMemcpyFromSymbol(out, g_pipe_trace, sizeof(unsigned long long) * 8 * 1024) == cudaSuccess ? 0 : 1;
}
#endif


template <int MODE>
static int launch_rows(ddf_mesh* m, cudaStream_t stream, RowArgs a) {
  const int64_t n = a.row1 - a.row0;
  if (n <= 0) return 0;
  a.rowptr = m->adj_ptr.p;
  a.nbr = m->adj_nbr.p;
  a.lval = m->adj_l.p;
  a.isb = m->isbnd[0].p;
  const int tune_wave = m->ctx->tune.wave;
  if (m->lstage.ok && (tune_wave & 3) == 0) {          // default: staged rows
    const int64_t w0 = a.row0 / DDF_SELL_WIN, w1 = (a.row1 + DDF_SELL_WIN - 1) / DDF_SELL_WIN;
    if (w1 - w0 > 0x7fffffffLL) DDF_FAIL("grid too large");
    const StageBufs& sb = m->lstage;
    PipeGeom g;
    // groups of K windows: the pipeline is the only kernel that understands K > 1
    const int64_t gw = (int64_t)sb.K * DDF_SELL_WIN;
    const int64_t g0 = a.row0 / gw, g1 = (a.row1 + gw - 1) / gw;
    const bool fits = pipe_geometry(m, &g);
    if (sb.K > 1 && !fits) DDF_FAIL("staged rows: a group of %d windows does not fit the shared-memory ring", sb.K);
    if (sb.K > 1 || (!(tune_wave & 128) && fits && w1 - w0 >= 4)) {          // persistent TMA pipeline
      const size_t smem_p = (size_t)g.nstages * g.stage_bytes;
      const int grid = (int)std::min<int64_t>(m->ctx->sm_count, g1 - g0);
      const int ng = (tune_wave >> 8) & 3;                     // A/B knob: consumer groups per block
      // the dynamic shared-memory limit is a per-device function attribute: set it on every launch (a host-side
      // table update, no device work) instead of caching it in a process-wide static
#define DDF_PIPE_LAUNCH(NGV)                                                                                        \
  do {                                                                                                              \
    DDF_CUDA(cudaFuncSetAttribute(k_rows_pipe<MODE, NGV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p)); \
    k_rows_pipe<MODE, NGV><<<grid, NGV * PIPE_GROUP + 32, smem_p, stream>>>(a, sb.tab.p, sb.packed.p, sb.meta.p,    \
                                                                           g0, g1, g);                               \
  } while (0)
      if (ng == 1) DDF_PIPE_LAUNCH(1);
      else if (ng == 3) DDF_PIPE_LAUNCH(3);
      else DDF_PIPE_LAUNCH(2);
#undef DDF_PIPE_LAUNCH
      DDF_LAUNCH_CHECK(m->ctx);
      return 0;
    }
    const size_t smem = (size_t)sb.max_slots * 8;
    if (smem > 48 * 1024 - 64) DDF_FAIL("staged rows: %zu bytes of staging exceed the static limit", smem);
    const int batch = (tune_wave >> 5) & 3;          // A/B knob: gathers in flight per thread
    if (batch == 1)
      k_rows_stage<MODE, 4><<<(int)(w1 - w0), DDF_SELL_WIN, smem, stream>>>(a, sb.ptr.p, sb.perm8.p, sb.len8.p, sb.tab.p,
                                                                           sb.packed.p, w0);
    else if (batch == 2)
      k_rows_stage<MODE, 6><<<(int)(w1 - w0), DDF_SELL_WIN, smem, stream>>>(a, sb.ptr.p, sb.perm8.p, sb.len8.p, sb.tab.p,
                                                                           sb.packed.p, w0);
    else if (batch == 3)
      k_rows_stage<MODE, 12><<<(int)(w1 - w0), DDF_SELL_WIN, smem, stream>>>(a, sb.ptr.p, sb.perm8.p, sb.len8.p, sb.tab.p,
                                                                            sb.packed.p, w0);
    else
      k_rows_stage<MODE, 8><<<(int)(w1 - w0), DDF_SELL_WIN, smem, stream>>>(a, sb.ptr.p, sb.perm8.p, sb.len8.p, sb.tab.p,
                                                                           sb.packed.p, w0);
  } else if (m->wave_sell && (tune_wave & 3) == 0) {   // explicit-index SELL-32 rows
    const int64_t w0 = a.row0 / DDF_SELL_WIN, w1 = (a.row1 + DDF_SELL_WIN - 1) / DDF_SELL_WIN;
    if (w1 - w0 > 0x7fffffffLL) DDF_FAIL("grid too large");
    const SellBufs& sb = m->lsell;
    if (sb.compressed)
      k_rows_sell<MODE, true><<<(int)(w1 - w0), DDF_SELL_WIN, 0, stream>>>(a, sb.ptr.p, sb.perm8.p, sb.cidx.p, sb.val.p,
                                                                            sb.umask.p, sb.iptr.p, w0);
    else if (sb.idx16)
      k_rows_sell<MODE, false, true><<<(int)(w1 - w0), DDF_SELL_WIN, 0, stream>>>(
          a, sb.ptr.p, sb.perm8.p, reinterpret_cast<const uint32_t*>(sb.col16.p), sb.val.p, nullptr, nullptr, w0);
    else
      k_rows_sell<MODE, false><<<(int)(w1 - w0), DDF_SELL_WIN, 0, stream>>>(a, sb.ptr.p, sb.perm8.p, sb.col.p, sb.val.p,
                                                                             nullptr, nullptr, w0);
  } else {                                // CSR, one thread per row
    int grid;
    DDF_CHECK(grid_for(n, 256, &grid));
    k_rows_thread<MODE><<<grid, 256, 0, stream>>>(a);
  }
  DDF_LAUNCH_CHECK(m->ctx);
  return 0;
}

int ddf_laplace0_apply(ddf_mesh* m, const double* x, double* y) {
  DDF_CHECK(ddf_mesh_require_wave(m));
  RowArgs a{};
  a.u = x;
  a.out0 = y;
  a.row0 = 0;
  a.row1 = m->n[0];
  return launch_rows<0>(m, m->ctx->stream, a);
}

// assembled L as COO in ROW order (row i: its stored entries, columns ascending); the caller sorts by column.
// (Taking column j's entries from row j's neighbour list -- "the pattern is symmetric" -- loses L[i,j] when the chop
// removed L[j,i] but kept L[i,j]: the two are different products ih0[i] * star1[e], ih0[j] * star1[e].)
__global__ void k_laplace0_coo(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                               const double* __restrict__ lval, int64_t nv, int32_t* __restrict__ row,
                               int32_t* __restrict__ col, double* __restrict__ val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  for (uint32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
    row[p] = (int32_t)i; col[p] = nbr[p]; val[p] = lval[p];
  }
}

int ddf_laplace0_to_coo(ddf_mesh* m, DevBuf<int32_t>& row, DevBuf<int32_t>& col, DevBuf<double>& val, int64_t* nnz) {
  DDF_CHECK(ddf_mesh_require_wave(m));
  ddf_ctx* ctx = m->ctx;
  const int64_t n = m->adj_nnz;
  *nnz = n;
  DDF_CHECK(row.alloc(ctx, n));
  DDF_CHECK(col.alloc(ctx, n));
  DDF_CHECK(val.alloc(ctx, n));
  if (!m->n[0]) return 0;
  int grid;
  DDF_CHECK(grid_for(m->n[0], 256, &grid));
  k_laplace0_coo<<<grid, 256, 0, ctx->stream>>>(m->adj_ptr.p, m->adj_nbr.p, m->adj_l.p, m->n[0], row.p, col.p, val.p);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

// examples/wave.jl:87-100: du = (E-B) udot ; dudot = (E-B) (L u)
extern "C" int ddf_wave_rhs(ddf_mesh* m, ddf_vec* u, ddf_vec* udot, ddf_vec* du, ddf_vec* dudot) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_wave(m));
  const int64_t nv = m->n[0];
  if (u->n != nv || udot->n != nv || du->n != nv || dudot->n != nv) DDF_FAIL("wave_rhs: vectors must have length n0=%lld", (long long)nv);
  if (du == u || dudot == u) DDF_FAIL("wave_rhs: outputs must not alias u");
  DDF_CHECK(ddf_vec_settle(u));
  DDF_CHECK(ddf_vec_settle(udot));
  DDF_CHECK(ddf_vec_settle(du));
  DDF_CHECK(ddf_vec_settle(dudot));
  RowArgs a{};
  a.u = u->d.p;
  a.udot = udot->d.p;
  a.out0 = du->d.p;
  a.out1 = dudot->d.p;
  a.row0 = 0;
  a.row1 = nv;
  return launch_rows<1>(m, m->ctx->stream, a);
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
// buffer so neighbours still see the old u.  Rows [row0, row1) on `stream`.
// One fused step on rows [row0, row1) for the distributed drivers of comm.cu.
// kind 0: leapfrog (aux = v, in/out; coef = dt)    kind 1: Poisson relaxation sweep (aux = rho; coef = theta)
int ddf_step_launch(ddf_mesh* m, cudaStream_t stream, int kind, const double* u, double* aux, double* unew, double coef,
                    int64_t row0, int64_t row1, double* peer_lo, double* peer_hi, int64_t plane) {
  RowArgs a{};
  a.u = u;
  a.out1 = unew;
  a.row0 = row0;
  a.row1 = row1;
  a.peer_lo = peer_lo;
  a.peer_hi = peer_hi;
  a.send_lo_end = row0 + plane;
  a.send_hi_beg = row1 - plane;
  if (kind == 0) {
    a.out0 = aux;
    a.dt = coef;
    return launch_rows<2>(m, stream, a);
  }
  if ((int64_t)m->ldinv.n != m->n[0]) DDF_FAIL("step_launch: 1/diag(L) not built");
  a.rho = aux;
  a.dinv = m->ldinv.p;
  a.theta = coef;
  return launch_rows<3>(m, stream, a);
}

#include "wavefront2.cuh"      // two leapfrog steps per launch with a wavefront through L2
#ifndef DDF_WAVE2_DEFAULT
#define DDF_WAVE2_DEFAULT 1    // measured at C4: 0.451 ms per step against 0.490 ms with one step per launch (profiles/r02_tuning.md)
#endif

// ---- for the slab driver of comm.cu: two leapfrog steps per launch with a DEEP halo (step n on the owned rows + one
// ghost plane per side, step n+1 on the owned rows; the state after both steps of the outermost owned planes is stored
// into the neighbours' landing arenas).  feasible: this rank's verdict (the ranks agree on it before using it).
int ddf_pipe2_feasible(ddf_mesh* m, int* ok) {
  *ok = 0;
  const int tw = m->ctx->tune.wave;
  if (!((tw & 65536) || (DDF_WAVE2_DEFAULT && !(tw & 131072) && (tw & 0xffff) == 0))) return 0;
  DDF_CHECK(ddf_mesh_require_wave(m));
  Pipe2Plan pl;
  bool good = false;
  DDF_CHECK(pipe2_plan(m, m->ctx->stream, &pl, &good));
  *ok = good ? 1 : 0;
  return 0;
}
int ddf_leapfrog_pair_launch(ddf_mesh* m, cudaStream_t stream, double* u, double* v, double* utmp, double dt, int64_t row0,
                             int64_t row1, int64_t row0b, int64_t row1b, int64_t plane, double* const* peers6, int* ran) {
  Pipe2Deep d;
  d.row0 = row0; d.row1 = row1; d.row0b = row0b; d.row1b = row1b; d.plane = plane;
  for (int q = 0; q < 6; q++) d.peers[q] = peers6[q];
  bool r = false;
  DDF_CHECK(launch_pipe2(m, stream, 0, u, v, utmp, dt, &r, &d));
  *ran = r ? 1 : 0;
  return 0;
}

int ddf_leapfrog_launch(ddf_mesh* m, cudaStream_t stream, const double* u, double* v, double* unew, double dt,
                        int64_t row0, int64_t row1) {
  RowArgs a{};
  a.u = u;
  a.out0 = v;
  a.out1 = unew;
  a.dt = dt;
  a.row0 = row0;
  a.row1 = row1;
  return launch_rows<2>(m, stream, a);
}

extern "C" int ddf_wave_leapfrog(ddf_mesh* m, ddf_vec* u, ddf_vec* v, double dt, int nsteps) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_wave(m));
  const int64_t nv = m->n[0];
  if (u->n != nv || v->n != nv) DDF_FAIL("wave_leapfrog: vectors must have length n0=%lld", (long long)nv);
  if (u == v) DDF_FAIL("wave_leapfrog: u and v must not alias");
  if (!nv || nsteps <= 0) return 0;
  DDF_CHECK(ddf_vec_settle(u));
  DDF_CHECK(ddf_vec_settle(v));
  if ((int64_t)m->utmp.n != nv) DDF_CHECK(m->utmp.alloc(m->ctx, nv));
  int s = 0;
  // two steps per launch (wavefront2.cuh; the state ends up in u again: no swap).  Tuning bit 65536 forces it on,
  // 131072 forces one step per launch; launch_pipe2 itself declines meshes where the wavefront cannot pay
  const int tw = m->ctx->tune.wave;
  if ((tw & 65536) || (DDF_WAVE2_DEFAULT && !(tw & 131072) && (tw & 0xffff) == 0)) {
    for (bool ran = true; ran && s + 1 < nsteps;) {
      DDF_CHECK(launch_pipe2(m, m->ctx->stream, 0, u->d.p, v->d.p, m->utmp.p, dt, &ran));
      if (ran) s += 2;
    }
  }
  for (; s < nsteps; s++) {
    DDF_CHECK(ddf_leapfrog_launch(m, m->ctx->stream, u->d.p, v->d.p, m->utmp.p, dt, 0, nv));
    u->d.swap(m->utmp);          // O(1) ping-pong: the vector handle now owns the new state
  }
  return 0;
  DDF_TRY_END
}

// ---------------------------------------------------------------- N1: fixed-step Tsit5 on the device
// examples/wave.jl:134 integrates dU = F U with `Tsit5(); adaptive=false, dt`.  This is
// OrdinaryDiffEq's in-place perform_step! on device vectors: six stage updates (one fused
// k_lincomb pass each, core.cu) and six right-hand sides (k_rows_* MODE 1) per step, the
// seventh derivative re-used as k1 of the next step (FSAL).  Tableau: Tsitouras 5(4),
// pinned by the order conditions in tests/test_oracle_pins.py.
int ddf_lincomb_launch(ddf_ctx* ctx, const double* base, double dt, int nk, const double* coef, const double* const* ks,
                       double* out, int64_t n);   // core.cu

int ddf_halo_exchange_raw(ddf_mesh* m, double* u);     // comm.cu
int ddf_halo_exchange_raw2(ddf_mesh* m, double* a, double* b);
int ddf_p2p_usable(ddf_mesh* m, int* usable);           // comm.cu: the fused peer-memory halo, pass by pass
int ddf_p2p_pass_begin(ddf_mesh* m, double* in_a, double* in_b, int copy, double** peers, int64_t* own_lo, int64_t* own_hi);
int ddf_p2p_pass_end(ddf_mesh* m);
int ddf_p2p_finish(ddf_mesh* m, double* a, double* b);

static const double TSIT5_A[6][6] = {
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};

static int wave_rhs_launch(ddf_mesh* m, const double* u, const double* udot, double* du, double* dudot) {
  RowArgs a{};
  a.u = u;
  a.udot = udot;
  a.out0 = du;
  a.out1 = dudot;
  a.row0 = 0;
  a.row1 = m->n[0];
  return launch_rows<1>(m, m->ctx->stream, a);
}

// The unfused sequence (one k_lincomb pass per stage and component + one wave right-hand side per stage): kept as the
// cross-check of the fused stages (ddf_set_tuning("wave", 1 << 29)); same bits.
static int tsit5_unfused(ddf_mesh* m, ddf_vec* u, ddf_vec* udot, double dt, int nsteps, bool dist) {
  ddf_ctx* ctx = m->ctx;
  const int64_t nv = m->n[0];
  // work space: k1..k7 and tmp, each for (u, udot): 16 vertex vectors, kept with the mesh between calls
  const size_t nvp = ((size_t)nv + 1) & ~(size_t)1;       // even stride: every sub-vector stays 16-byte aligned (TMA, 128-bit loads)
  if (m->tsit_work.n != 16 * nvp) DDF_CHECK(m->tsit_work.alloc(ctx, 16 * nvp));
  double* kp[7][2];
  for (int s = 0; s < 7; s++)
    for (int c = 0; c < 2; c++) kp[s][c] = m->tsit_work.p + (size_t)(2 * s + c) * nvp;
  struct { double* p; } tmp[2] = {{m->tsit_work.p + 14 * nvp}, {m->tsit_work.p + 15 * nvp}};
  double* U[2] = {u->d.p, udot->d.p};
  if (dist) DDF_CHECK(ddf_halo_exchange_raw(m, U[0]));
  DDF_CHECK(wave_rhs_launch(m, U[0], U[1], kp[0][0], kp[0][1]));
  for (int step = 0; step < nsteps; step++) {
    for (int s = 0; s < 6; s++) {
      // stages 1..5 go to tmp; the last one IS the new state, written over U after its inputs are consumed
      double* dst[2] = {tmp[0].p, tmp[1].p};
      for (int c = 0; c < 2; c++) {
        const double* ks[6];
        for (int j = 0; j <= s; j++) ks[j] = kp[j][c];
        DDF_CHECK(ddf_lincomb_launch(ctx, U[c], dt, s + 1, TSIT5_A[s], ks, s == 5 ? U[c] : dst[c], nv));
      }
      double* su = s == 5 ? U[0] : dst[0];
      const double* sv = s == 5 ? U[1] : dst[1];
      if (dist) DDF_CHECK(ddf_halo_exchange_raw(m, su));
      DDF_CHECK(wave_rhs_launch(m, su, sv, kp[s + 1][0], kp[s + 1][1]));
    }
    for (int c = 0; c < 2; c++) std::swap(kp[0][c], kp[6][c]);     // FSAL
  }
  return 0;
}

// Three launches of k_rows_pipe<5,2> per step (two stages each); tuning bit 1 << 30 selects one stage per launch.
static int tsit5_paired(ddf_mesh* m, ddf_vec* u, ddf_vec* udot, double dt, int nsteps, bool dist, const PipeGeom& g) {
  ddf_ctx* ctx = m->ctx;
  const int64_t nv = m->n[0];
  const StageBufs& sb = m->lstage;
  const size_t nvp = ((size_t)nv + 1) & ~(size_t)1;       // even stride: every sub-vector stays 16-byte aligned (TMA)
  // work space: k^v_1 .. k^v_4 and four stage inputs (su_2 | su_3, su_4 | su_5, su_6 rotate through them)
  if (m->tsit_work.n != 8 * nvp) DDF_CHECK(m->tsit_work.alloc(ctx, 8 * nvp));
  double* kvp[4];
  double* X[4];
  for (int q = 0; q < 4; q++) { kvp[q] = m->tsit_work.p + (size_t)q * nvp; X[q] = m->tsit_work.p + (size_t)(4 + q) * nvp; }
  const int64_t gw = (int64_t)sb.K * DDF_SELL_WIN;
  const int64_t g1 = (nv + gw - 1) / gw;
  const size_t smem_p = (size_t)g.nstages * g.stage_bytes;
  const int grid = (int)std::min<int64_t>(ctx->sm_count, g1);
  DDF_CUDA(cudaFuncSetAttribute(k_rows_pipe<5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
  // slab meshes: the fused peer-memory halo when the ranks can map each other (comm.cu), else grouped ncclSend/ncclRecv
  int p2p = 0;
  if (dist) DDF_CHECK(ddf_p2p_usable(m, &p2p));
  bool first_pass = true;
  auto pass = [&](int J, double* ga, double* gb, double* o1, double* o2, double* o3) -> int {
    RowArgs a{};
    a.isb = m->isbnd[0].p;
    a.u = ga;
    a.u2 = gb;
    for (int q = 0; q < J - 1; q++) a.kv[q] = kvp[q];
    a.base_u = u->d.p;
    a.base_v = udot->d.p;
    a.J = J;
    a.dt = dt;
    a.row0 = 0;
    a.row1 = nv;
    int64_t gg0 = 0, gg1 = g1;
    if (p2p) {
      // wait for the neighbours' previous pass, take over what they landed (not before the first pass of a call: the
      // ghost planes of its inputs come from the NCCL exchange below), then run the owned rows only
      double* peers[4];
      DDF_CHECK(ddf_p2p_pass_begin(m, ga, gb, first_pass ? 0 : 1, peers, &a.row0, &a.row1));
      first_pass = false;
      a.peer_lo = peers[0]; a.peer_hi = peers[1]; a.peer_lo2 = peers[2]; a.peer_hi2 = peers[3];
      a.send_lo_end = a.row0 + m->halo_plane;
      a.send_hi_beg = a.row1 - m->halo_plane;
      gg0 = a.row0 / gw;
      gg1 = (a.row1 + gw - 1) / gw;
    }
    if (J < 5) { a.out_kv = kvp[J - 1]; a.out_kv2 = kvp[J]; a.out_su = o1; a.out_su2 = o2; }
    else { a.out_su = o1; a.out_sv = o2; a.out_su2 = o3; }
    const int gridj = (int)std::min<int64_t>(ctx->sm_count, gg1 - gg0);
    k_rows_pipe<5, 2><<<gridj, 2 * PIPE_GROUP + 32, smem_p, ctx->stream>>>(a, sb.tab.p, sb.packed.p, sb.meta.p, gg0, gg1, g);
    DDF_LAUNCH_CHECK(ctx);
    if (p2p) DDF_CHECK(ddf_p2p_pass_end(m));
    return 0;
  };
  int gridp;
  DDF_CHECK(grid_for(nv, 256, &gridp));
  if (dist) DDF_CHECK(ddf_halo_exchange_raw(m, u->d.p));
  k_tsit5_first<<<gridp, 256, 0, ctx->stream>>>(u->d.p, udot->d.p, m->isbnd[0].p, dt, nv, X[0]);
  DDF_LAUNCH_CHECK(ctx);
  if (dist) DDF_CHECK(ddf_halo_exchange_raw(m, X[0]));
  int c = 0;                                               // X[c] holds su_2 of the current step
  for (int step = 0; step < nsteps; step++) {
    double *s2 = X[c], *s3 = X[(c + 1) & 3], *s4 = X[(c + 2) & 3], *s5 = X[(c + 3) & 3], *s6 = X[c], *n2 = X[(c + 1) & 3];
    DDF_CHECK(pass(1, u->d.p, s2, s3, s4, nullptr));
    if (dist && !p2p) DDF_CHECK(ddf_halo_exchange_raw2(m, s3, s4));
    DDF_CHECK(pass(3, s3, s4, s5, s6, nullptr));           // su_6 overwrites su_2, which nobody reads any more
    if (dist && !p2p) DDF_CHECK(ddf_halo_exchange_raw2(m, s5, s6));
    DDF_CHECK(pass(5, s5, s6, u->d.p, udot->d.p, n2));     // in place: a row reads its own u, udot before it stores them
    if (dist && !p2p) DDF_CHECK(ddf_halo_exchange_raw2(m, u->d.p, n2));
    c = (c + 1) & 3;
  }
  if (p2p && nsteps > 0) DDF_CHECK(ddf_p2p_finish(m, u->d.p, nullptr));   // leave with valid ghost planes of u
  return 0;
}

extern "C" int ddf_wave_tsit5(ddf_mesh* m, ddf_vec* u, ddf_vec* udot, double dt, int nsteps) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_wave(m));
  ddf_ctx* ctx = m->ctx;
  const int64_t nv = m->n[0];
  if (u->n != nv || udot->n != nv) DDF_FAIL("wave_tsit5: vectors must have length n0=%lld", (long long)nv);
  if (u == udot) DDF_FAIL("wave_tsit5: u and udot must not alias");
  if (!nv || nsteps <= 0) return 0;
  DDF_CHECK(ddf_vec_settle(u));
  DDF_CHECK(ddf_vec_settle(udot));
  // partitioned meshes (one rank per GPU): the ghost entries of the state every right-hand side reads are refreshed
  // from their owners first; ghost rows of the derivatives are scratch and never reach an owned row
  const bool dist = ctx->nranks > 1 && (m->halo_plane > 0 || m->halo_gen);
  if (ctx->tune.wave & (1 << 29)) {
    DDF_CHECK(tsit5_unfused(m, u, udot, dt, nsteps, dist));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  // Paired stages (default on meshes whose rows are staged): THREE passes over L per step, see tsit5_pair.
  {
    PipeGeom gp;
    const int tw = ctx->tune.wave;
    if (m->lstage.ok && (tw & 3) == 0 && !(tw & 128) && !(tw & (1 << 30)) && pipe_geometry(m, &gp, true)) {
      DDF_CHECK(tsit5_paired(m, u, udot, dt, nsteps, dist, gp));
      DDF_CUDA(cudaStreamSynchronize(ctx->stream));
      return 0;
    }
  }
  // Fused stages: SIX launches per step.  Stage kernel J gathers su_J (su_1 = u), stores k^v_J and the next stage
  // input su_{J+1}; kernel 6 stores the new state over (u, udot).  FSAL is implicit: k_1 of the next step is what its
  // kernel 1 computes from the new u.  Work space: k^v_1..k^v_5 and two stage inputs (ping-pong) = 7 vertex vectors.
  const size_t nvp = ((size_t)nv + 1) & ~(size_t)1;       // even stride: every sub-vector stays 16-byte aligned (TMA)
  if (m->tsit_work.n != 7 * nvp) DDF_CHECK(m->tsit_work.alloc(ctx, 7 * nvp));
  double* kvp[5];
  for (int q = 0; q < 5; q++) kvp[q] = m->tsit_work.p + (size_t)q * nvp;
  double* su[2] = {m->tsit_work.p + 5 * nvp, m->tsit_work.p + 6 * nvp};
  if (dist) DDF_CHECK(ddf_halo_exchange_raw(m, u->d.p));
  for (int step = 0; step < nsteps; step++) {
    for (int J = 1; J <= 6; J++) {
      RowArgs a{};
      a.u = J == 1 ? u->d.p : su[J & 1];                   // su_2 -> su[0], su_3 -> su[1], ...
      for (int q = 0; q < J - 1; q++) a.kv[q] = kvp[q];
      a.base_u = u->d.p;
      a.base_v = udot->d.p;
      a.out_kv = J < 6 ? kvp[J - 1] : nullptr;
      a.out_su = J < 6 ? su[(J + 1) & 1] : u->d.p;
      a.out_sv = J < 6 ? nullptr : udot->d.p;
      a.J = J;
      a.dt = dt;
      a.row0 = 0;
      a.row1 = nv;
      DDF_CHECK(launch_rows<4>(m, ctx->stream, a));
      if (dist) DDF_CHECK(ddf_halo_exchange_raw(m, a.out_su));
    }
  }
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
  DDF_TRY_END
}

// 1 ./ diag(L) from the assembled rows (0 where the diagonal entry was chopped away)
__global__ void k_laplace0_dinv(const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                                const double* __restrict__ lval, int64_t nv, double* __restrict__ dinv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  double d = 0.0;
  for (uint32_t p = rowptr[i]; p < rowptr[i + 1]; p++)
    if (nbr[p] == (int32_t)i) d = lval[p];
  dinv[i] = d != 0.0 ? __ddiv_rn(1.0, d) : 0.0;
}

int ddf_mesh_require_dinv(ddf_mesh* m) {
  DDF_CHECK(ddf_mesh_require_wave(m));
  const int64_t nv = m->n[0];
  if ((int64_t)m->ldinv.n == nv) return 0;
  DDF_CHECK(m->ldinv.alloc(m->ctx, nv));
  if (!nv) return 0;
  int grid;
  DDF_CHECK(grid_for(nv, 256, &grid));
  k_laplace0_dinv<<<grid, 256, 0, m->ctx->stream>>>(m->adj_ptr.p, m->adj_nbr.p, m->adj_l.p, nv, m->ldinv.p);
  DDF_LAUNCH_CHECK(m->ctx);
  return 0;
}

// Fused Poisson relaxation step (north_star kernel (3), "wave/Poisson step"): weighted Jacobi on
//   L u = rho in the interior, u fixed on the boundary   (examples/poisson.jl:75-81, 172-201 with f eliminated)
//   u' = u + theta * ((1-b) .* ((rho - L u) ./ diag(L)))
// one pass per sweep through the same row kernels as the wave step (u ping-pong, rho and u staged by TMA).
extern "C" int ddf_poisson_jacobi(ddf_mesh* m, ddf_vec* u, ddf_vec* rho, double theta, int nsweeps) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_wave(m));
  ddf_ctx* ctx = m->ctx;
  const int64_t nv = m->n[0];
  if (u->n != nv || rho->n != nv) DDF_FAIL("poisson_jacobi: vectors must have length n0=%lld", (long long)nv);
  if (u == rho) DDF_FAIL("poisson_jacobi: u and rho must not alias");
  if (!nv || nsweeps <= 0) return 0;
  DDF_CHECK(ddf_vec_settle(u));
  DDF_CHECK(ddf_vec_settle(rho));
  DDF_CHECK(ddf_mesh_require_dinv(m));
  if ((int64_t)m->utmp.n != nv) DDF_CHECK(m->utmp.alloc(ctx, nv));
  int s = 0;
  {
    // two sweeps per launch with the wavefront through L2 (wavefront2.cuh), like the leapfrog step
    const int tw = ctx->tune.wave;
    if ((tw & 65536) || (DDF_WAVE2_DEFAULT && !(tw & 131072) && (tw & 0xffff) == 0)) {
      for (bool ran = true; ran && s + 1 < nsweeps;) {
        DDF_CHECK(launch_pipe2(m, ctx->stream, 1, u->d.p, rho->d.p, m->utmp.p, theta, &ran));
        if (ran) s += 2;
      }
    }
  }
  for (; s < nsweeps; s++) {
    RowArgs a{};
    a.u = u->d.p;
    a.out1 = m->utmp.p;
    a.rho = rho->d.p;
    a.dinv = m->ldinv.p;
    a.theta = theta;
    a.row0 = 0;
    a.row1 = nv;
    DDF_CHECK(launch_rows<3>(m, ctx->stream, a));
    u->d.swap(m->utmp);
  }
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_wave_format(ddf_mesh* m, int* format) {
  DDF_TRY_BEGIN
  DDF_CHECK(ddf_mesh_require_wave(m));
  *format = m->lstage.ok ? 2 : (m->wave_sell ? 1 : 0);
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_wave_step_bytes(ddf_mesh* m, int64_t* algorithmic, int64_t* format_bytes) {
  DDF_CHECK(ddf_mesh_require_assembled(m));
  const int64_t E = m->n[1], V = m->n[0];
  *algorithmic = E * 16 + V * 41;                  // E*(2*4+8) + V*(8*5+1)
  // what this format streams per step: 12 B per stored entry + u, v (r/w), u', mask per vertex + row offsets
  if (m->has_wave && m->lstage.ok)   // 10 B per entry, no padding; per vertex u (staged once), v r/w, u'; per window meta + run table
    *format_bytes = m->lstage.entries * 10 + V * (8 * 4) + (V / 256 + 1) * (STAGE_META + STAGE_TAB * 4);
  else if (m->has_wave && m->wave_sell)
    *format_bytes = m->lsell.padded * 8 + m->lsell.index_words * 4 + V * (8 * 4 + 1 + 1) + (V / 32 + 1) * 12;   // index_words: 16-bit offsets count as half words
  else *format_bytes = (m->has_wave ? m->adj_nnz : 2 * E + V) * 12 + V * (4 + 8 * 4 + 1);
  return 0;
}

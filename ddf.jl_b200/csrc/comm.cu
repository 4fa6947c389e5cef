// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink 5.
// libnccl is resolved lazily with dlopen so the single-GPU path has no NCCL
// dependency and so that, inside a torch process, the already-loaded
// libnccl.so.2 (torch's bundled 2.28.9) is the one that gets used.
//
// The only exchange step of this path is the ghost-plane halo of a vertex cochain
// (SURVEY 8(e)): slabs along the slowest grid axis, neighbours rank-1 / rank+1,
// ncclSend/ncclRecv inside one group, issued on a second stream so that the
// interior rows of the fused wave step overlap the transfer.
#include <dlfcn.h>
#include <nccl.h>

#include <cub/cub.cuh>

#include <algorithm>
#include <initializer_list>
#include <vector>

#include "mesh.cuh"

int ddf_pipe2_feasible(ddf_mesh* m, int* ok);           // wave.cu
int ddf_leapfrog_pair_launch(ddf_mesh* m, cudaStream_t stream, double* u, double* v, double* utmp, double dt, int64_t row0,
                             int64_t row1, int64_t row0b, int64_t row1b, int64_t plane, double* const* peers6, int* ran);
int ddf_step_launch(ddf_mesh* m, cudaStream_t stream, int kind, const double* u, double* aux, double* unew, double coef,
                    int64_t row0, int64_t row1, double* peer_lo, double* peer_hi, int64_t plane);           // wave.cu
int ddf_mesh_require_dinv(ddf_mesh* m);                                                                      // wave.cu

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) DDF_FAIL("cannot dlopen libnccl.so.2: %s", dlerror());
#define DDF_SYM(field, name)                                         \
  *(void**)(&g_nccl.field) = dlsym(h, name);                         \
  if (!g_nccl.field) DDF_FAIL("libnccl: missing symbol %s", name);
  DDF_SYM(GetUniqueId, "ncclGetUniqueId")
  DDF_SYM(CommInitRank, "ncclCommInitRank")
  DDF_SYM(CommDestroy, "ncclCommDestroy")
  DDF_SYM(Send, "ncclSend")
  DDF_SYM(Recv, "ncclRecv")
  DDF_SYM(GroupStart, "ncclGroupStart")
  DDF_SYM(GroupEnd, "ncclGroupEnd")
  DDF_SYM(AllReduce, "ncclAllReduce")
  DDF_SYM(GetErrorString, "ncclGetErrorString")
#undef DDF_SYM
  g_nccl.handle = h;
  return 0;
}

#define DDF_NCCL(expr)                                                                            \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess) {                                                                      \
      ddf_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

extern "C" int ddf_comm_unique_id(void* id128) {
  DDF_TRY_BEGIN
  DDF_CHECK(nccl_load());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  DDF_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_comm_init(ddf_ctx* ctx, int nranks, int rank, const void* id128) {
  DDF_TRY_BEGIN
  if (ctx->nccl_comm) DDF_FAIL("comm_init: communicator already initialised");
  if (nranks < 1 || rank < 0 || rank >= nranks) DDF_FAIL("comm_init: bad rank %d of %d", rank, nranks);
  ctx->nranks = nranks;
  ctx->rank = rank;
  if (nranks == 1) return 0;
  DDF_CHECK(nccl_load());
  DDF_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  DDF_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = (void*)comm;
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_comm_destroy(ddf_ctx* ctx) {
  if (ctx->nccl_comm && g_nccl.CommDestroy) {
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  return 0;
}

extern "C" int ddf_comm_allreduce(ddf_ctx* ctx, double* value, int op) {
  DDF_TRY_BEGIN
  if (ctx->nranks == 1) return 0;
  if (!ctx->nccl_comm) DDF_FAIL("comm_allreduce: communicator not initialised");
  ncclRedOp_t rop = op == 0 ? ncclSum : (op == 1 ? ncclMax : ncclMin);
  ctx->red_host[0] = *value;
  DDF_CUDA(cudaMemcpyAsync(ctx->red_dev, ctx->red_host, 8, cudaMemcpyHostToDevice, ctx->stream));
  DDF_NCCL(g_nccl.AllReduce(ctx->red_dev, ctx->red_dev, 1, ncclDouble, rop, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  DDF_CUDA(cudaMemcpyAsync(ctx->red_host, ctx->red_dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  *value = ctx->red_host[0];
  return 0;
  DDF_TRY_END
}

void ddf_halo_gen_release(ddf_mesh* m);
extern "C" int ddf_mesh_set_halo(ddf_mesh* m, int64_t plane, int64_t ghost_lo, int64_t ghost_hi) {
  if (plane <= 0 || ghost_lo < 0 || ghost_hi < 0 || (ghost_lo + ghost_hi + 1) * plane > m->n[0] || m->n[0] % plane)
    DDF_FAIL("set_halo: inconsistent slab description (plane=%lld, lo=%lld, hi=%lld, V=%lld)", (long long)plane,
             (long long)ghost_lo, (long long)ghost_hi, (long long)m->n[0]);
  ddf_halo_gen_release(m);            // a slab description replaces a vertex-range partition and vice versa
  m->halo_plane = plane;
  m->halo_lo = ghost_lo;
  m->halo_hi = ghost_hi;
  return 0;
}

// =====================================================================================
// General halo: a mesh of ANY origin (uploaded, refined, Delaunay) partitioned by contiguous ranges of the global
// vertex numbering (north_star: "large refined meshes are partitioned by simplex ranges").  The host side
// (ddf.dist.partition_mesh) gives every rank the closed star of its owned vertices, renumbered monotonically, so the
// local face ranks keep the global order and every owned row (geometry, star values, L) is bit-identical to the
// undivided mesh; ghost vertices get their values from their owners through explicit index lists, one grouped
// ncclSend/ncclRecv per exchange, peers in any number and order.
// =====================================================================================
struct HaloGeneral {
  std::vector<int> peers;
  std::vector<int64_t> send_ptr, recv_ptr;      // npeers + 1 offsets into the index lists
  DevBuf<uint32_t> send_idx, recv_idx;          // local vertex ids
  DevBuf<double> send_buf, recv_buf;
};
__global__ void k_gen_pack(const uint32_t* __restrict__ idx, const double* __restrict__ src, int64_t n, double* __restrict__ dst) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = src[idx[p]];
}
__global__ void k_gen_unpack(const uint32_t* __restrict__ idx, const double* __restrict__ src, int64_t n, double* __restrict__ dst) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[idx[p]] = src[p];
}
void ddf_halo_gen_release(ddf_mesh* m) {
  delete (HaloGeneral*)m->halo_gen;
  m->halo_gen = nullptr;
}
// own_lo / own_hi: the owned local vertex range; peers[k] exchanges send_idx[send_ptr[k] .. send_ptr[k+1]) (values this
// rank owns and peer k holds as ghosts) against recv_idx[recv_ptr[k] .. recv_ptr[k+1]) (this rank's ghosts owned by
// peer k); both sides list a pair's vertices in ascending GLOBAL id, so positions match.  0-based local ids.
extern "C" int ddf_mesh_set_halo_lists(ddf_mesh* m, int64_t own_lo, int64_t own_hi, int npeers, const int32_t* peers,
                                       const int64_t* send_ptr, const int64_t* send_idx, const int64_t* recv_ptr,
                                       const int64_t* recv_idx) {
  DDF_TRY_BEGIN
  ddf_ctx* ctx = m->ctx;
  if (own_lo < 0 || own_hi < own_lo || own_hi > m->n[0]) DDF_FAIL("set_halo_lists: owned range [%lld, %lld) outside 0..%lld", (long long)own_lo, (long long)own_hi, (long long)m->n[0]);
  if (npeers < 0) DDF_FAIL("set_halo_lists: npeers < 0");
  // validate everything before the mesh's current description is replaced
  for (int k = 0; k < npeers; k++)
    if (peers[k] < 0 || peers[k] >= ctx->nranks || peers[k] == ctx->rank || send_ptr[k + 1] < send_ptr[k] || recv_ptr[k + 1] < recv_ptr[k])
      DDF_FAIL("set_halo_lists: bad peer %d", k);
  const int64_t ns = npeers ? send_ptr[npeers] : 0, nr = npeers ? recv_ptr[npeers] : 0;
  std::vector<uint32_t> hs(ns), hr(nr);
  for (int64_t q = 0; q < ns; q++) {
    if (send_idx[q] < own_lo || send_idx[q] >= own_hi) DDF_FAIL("set_halo_lists: send entry %lld is not an owned vertex", (long long)q);
    hs[q] = (uint32_t)send_idx[q];
  }
  for (int64_t q = 0; q < nr; q++) {
    if (recv_idx[q] < 0 || recv_idx[q] >= m->n[0] || (recv_idx[q] >= own_lo && recv_idx[q] < own_hi)) DDF_FAIL("set_halo_lists: recv entry %lld is not a ghost vertex", (long long)q);
    hr[q] = (uint32_t)recv_idx[q];
  }
  ddf_halo_gen_release(m);
  HaloGeneral* hg = new HaloGeneral();
  m->halo_gen = hg;
  m->gen_own_lo = own_lo;
  m->gen_own_hi = own_hi;
  m->halo_plane = 0;
  hg->peers.assign(peers, peers + npeers);
  hg->send_ptr.assign(send_ptr, send_ptr + npeers + 1);
  hg->recv_ptr.assign(recv_ptr, recv_ptr + npeers + 1);
  DDF_CHECK(hg->send_idx.alloc(ctx, ns));
  DDF_CHECK(hg->recv_idx.alloc(ctx, nr));
  DDF_CHECK(hg->send_buf.alloc(ctx, ns));
  DDF_CHECK(hg->recv_buf.alloc(ctx, nr));
  if (ns) DDF_CUDA(cudaMemcpyAsync(hg->send_idx.p, hs.data(), ns * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (nr) DDF_CUDA(cudaMemcpyAsync(hg->recv_idx.p, hr.data(), nr * 4, cudaMemcpyHostToDevice, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
  DDF_TRY_END
}

static int halo_exchange_general(ddf_mesh* m, double* u, cudaStream_t stream) {
  ddf_ctx* ctx = m->ctx;
  HaloGeneral* hg = (HaloGeneral*)m->halo_gen;
  if (ctx->nranks == 1 || hg->peers.empty()) return 0;
  if (!ctx->nccl_comm) DDF_FAIL("halo: communicator not initialised (ddf_comm_init)");
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  const int np = (int)hg->peers.size();
  const int64_t ns = hg->send_ptr[np], nr = hg->recv_ptr[np];
  int grid;
  if (ns) {
    DDF_CHECK(grid_for(ns, 256, &grid));
    k_gen_pack<<<grid, 256, 0, stream>>>(hg->send_idx.p, u, ns, hg->send_buf.p);
    DDF_LAUNCH_CHECK(ctx);
  }
  DDF_NCCL(g_nccl.GroupStart());
  for (int k = 0; k < np; k++) {
    const int64_t s0 = hg->send_ptr[k], s1 = hg->send_ptr[k + 1], r0 = hg->recv_ptr[k], r1 = hg->recv_ptr[k + 1];
    if (s1 > s0) DDF_NCCL(g_nccl.Send(hg->send_buf.p + s0, s1 - s0, ncclDouble, hg->peers[k], comm, stream));
    if (r1 > r0) DDF_NCCL(g_nccl.Recv(hg->recv_buf.p + r0, r1 - r0, ncclDouble, hg->peers[k], comm, stream));
  }
  DDF_NCCL(g_nccl.GroupEnd());
  ctx->launches++;
  if (nr) {
    DDF_CHECK(grid_for(nr, 256, &grid));
    k_gen_unpack<<<grid, 256, 0, stream>>>(hg->recv_idx.p, hg->recv_buf.p, nr, u);
    DDF_LAUNCH_CHECK(ctx);
  }
  return 0;
}

// rows of the rank-R tables this rank OWNS (smallest vertex in an owned plane), as one range [lo, hi):
// below the top rank the tables are sorted lexicographically, so the owned rows are the lower-bound range of the first
// column; the top-level table is in the caller's order, where the range is measured (and must be contiguous, which it is
// for the generator's block order -- otherwise the whole table is returned).
__global__ void k_owned_top(const int32_t* __restrict__ simp, int N, int64_t n, int32_t vlo, int32_t vhi,
                            unsigned long long* __restrict__ st /* min, max, count */) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool own = j < n && simp[j * N] >= vlo && simp[j * N] < vhi;
  const unsigned mask = __ballot_sync(0xffffffffu, own);
  if (!mask) return;
  if ((threadIdx.x & 31) == (unsigned)(__ffs(mask) - 1)) atomicMin(st, (unsigned long long)j);
  if ((threadIdx.x & 31) == (unsigned)(31 - __clz(mask))) atomicMax(st + 1, (unsigned long long)j);
  if ((threadIdx.x & 31) == (unsigned)(__ffs(mask) - 1)) atomicAdd(st + 2, (unsigned long long)__popc(mask));
}

extern "C" int ddf_mesh_owned_rows(ddf_mesh* m, int R, int64_t* lo, int64_t* hi) {
  DDF_TRY_BEGIN
  if (R < 0 || R > m->D) DDF_FAIL("owned_rows: R=%d out of range", R);
  DDF_CHECK(ddf_mesh_require_assembled(m));
  ddf_ctx* ctx = m->ctx;
  DDF_CUDA(cudaSetDevice(ctx->device));
  const int64_t nv = m->n[0], n = m->n[R];
  *lo = 0;
  *hi = n;
  if ((m->halo_plane <= 0 && !m->halo_gen) || n == 0) return 0;        // no halo: everything is owned
  const int64_t vlo = m->halo_gen ? m->gen_own_lo : m->halo_lo * m->halo_plane;
  const int64_t vhi = m->halo_gen ? m->gen_own_hi : nv - m->halo_hi * m->halo_plane;
  if (R == 0) { *lo = vlo; *hi = vhi; return 0; }
  const int N = R + 1;
  if (R < m->D) {
    auto lower_bound = [&](int64_t v, int64_t* out) -> int {   // first row whose smallest vertex is >= v
      int64_t a = 0, b = n;
      while (a < b) {
        const int64_t mid = (a + b) >> 1;
        int32_t first = 0;
        DDF_CUDA(cudaMemcpyAsync(&first, m->simp[R].p + mid * N, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        DDF_CUDA(cudaStreamSynchronize(ctx->stream));
        if (first < v) a = mid + 1; else b = mid;
      }
      *out = a;
      return 0;
    };
    DDF_CHECK(lower_bound(vlo, lo));
    DDF_CHECK(lower_bound(vhi, hi));
    return 0;
  }
  DevBuf<unsigned long long> st;
  DDF_CHECK(st.alloc(ctx, 3));
  const unsigned long long init[3] = {~0ull, 0ull, 0ull};
  DDF_CUDA(cudaMemcpyAsync(st.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  int grid;
  DDF_CHECK(grid_for(n, 256, &grid));
  k_owned_top<<<grid, 256, 0, ctx->stream>>>(m->simp[R].p, N, n, (int32_t)vlo, (int32_t)vhi, st.p);
  DDF_LAUNCH_CHECK(ctx);
  unsigned long long h[3] = {0, 0, 0};
  DDF_CUDA(cudaMemcpyAsync(h, st.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h[2] == 0) { *lo = 0; *hi = 0; return 0; }
  if (h[1] - h[0] + 1 == h[2]) { *lo = (int64_t)h[0]; *hi = (int64_t)h[1] + 1; }
  return 0;
  DDF_TRY_END
}

// exchange on `stream`: my lowest owned plane -> rank-1 (into its first upper ghost plane),
// my highest owned plane -> rank+1 (into its last lower ghost plane), and the converse receives.
static int halo_exchange_on(ddf_mesh* m, double* u, cudaStream_t stream) {
  ddf_ctx* ctx = m->ctx;
  if (m->halo_gen) return halo_exchange_general(m, u, stream);
  if (ctx->nranks == 1) return 0;
  if (!ctx->nccl_comm) DDF_FAIL("halo: communicator not initialised (ddf_comm_init)");
  if (!m->halo_plane) DDF_FAIL("halo: slab not described (ddf_mesh_set_halo)");
  const int64_t P = m->halo_plane, V = m->n[0];
  const int64_t own_lo = m->halo_lo * P, own_hi = V - m->halo_hi * P;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  DDF_NCCL(g_nccl.GroupStart());
  if (m->halo_lo > 0) {
    DDF_NCCL(g_nccl.Send(u + own_lo, P, ncclDouble, ctx->rank - 1, comm, stream));
    DDF_NCCL(g_nccl.Recv(u + own_lo - P, P, ncclDouble, ctx->rank - 1, comm, stream));
  }
  if (m->halo_hi > 0) {
    DDF_NCCL(g_nccl.Send(u + own_hi - P, P, ncclDouble, ctx->rank + 1, comm, stream));
    DDF_NCCL(g_nccl.Recv(u + own_hi, P, ncclDouble, ctx->rank + 1, comm, stream));
  }
  DDF_NCCL(g_nccl.GroupEnd());
  ctx->launches += (m->halo_lo > 0) + (m->halo_hi > 0);
  return 0;
}

// raw-pointer variant for the integrators of wave.cu (vertex cochain on the compute stream)
int ddf_halo_exchange_raw(ddf_mesh* m, double* u) { return halo_exchange_on(m, u, m->ctx->stream); }

// two vertex cochains at once (the two stage inputs a paired Tsit5 pass produces): on slabs ONE grouped
// ncclSend/ncclRecv carries the interface planes of both
int ddf_halo_exchange_raw2(ddf_mesh* m, double* a, double* b) {
  ddf_ctx* ctx = m->ctx;
  cudaStream_t stream = ctx->stream;
  m->halo_last = 2;
  if (m->halo_gen || ctx->nranks == 1 || !ctx->nccl_comm || !m->halo_plane) {
    DDF_CHECK(halo_exchange_on(m, a, stream));
    return halo_exchange_on(m, b, stream);
  }
  const int64_t P = m->halo_plane, V = m->n[0];
  const int64_t own_lo = m->halo_lo * P, own_hi = V - m->halo_hi * P;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  DDF_NCCL(g_nccl.GroupStart());
  for (double* u : {a, b}) {
    if (m->halo_lo > 0) {
      DDF_NCCL(g_nccl.Send(u + own_lo, P, ncclDouble, ctx->rank - 1, comm, stream));
      DDF_NCCL(g_nccl.Recv(u + own_lo - P, P, ncclDouble, ctx->rank - 1, comm, stream));
    }
    if (m->halo_hi > 0) {
      DDF_NCCL(g_nccl.Send(u + own_hi - P, P, ncclDouble, ctx->rank + 1, comm, stream));
      DDF_NCCL(g_nccl.Recv(u + own_hi, P, ncclDouble, ctx->rank + 1, comm, stream));
    }
  }
  DDF_NCCL(g_nccl.GroupEnd());
  ctx->launches += (m->halo_lo > 0) + (m->halo_hi > 0);
  return 0;
}

extern "C" int ddf_halo_exchange(ddf_mesh* m, ddf_vec* u) {
  DDF_TRY_BEGIN
  if (u->n != m->n[0]) DDF_FAIL("halo_exchange: vector must be a vertex cochain");
  DDF_CHECK(ddf_vec_settle(u));
  return halo_exchange_on(m, u->d.p, m->ctx->stream);
  DDF_TRY_END
}

// =====================================================================================
// Fused halo over NVLink peer memory.
//
// The NCCL path (ddf_wave_leapfrog_dist below) needs three row-kernel launches and one send/recv kernel per step.
// Here the compute of a step is ONE launch: the row kernel stores the new u of its two
// interface planes straight into a landing area in the neighbours' memory (a per-mesh arena
// mapped once through CUDA IPC), so the transfer rides along with the compute, row by row.
// Neighbours stay within one step of each other through a counter in the same arena:
//   * after step k a 1-thread kernel release-stores k into both neighbours' flag words;
//   * before step k a small kernel waits until both of MY flag words hold >= k-1 -- the
//     neighbour has then finished step k-1, its stores into my landing area (parity (k-1)&1)
//     have arrived -- and copies them into the ghost planes of u.  The neighbour overwrites
//     that parity again in its step k+1, which it cannot start before I signal step k.
// The handshake (IPC handle of the arena, plane size, step counter) travels over the NCCL
// communicator on the first distributed step of a mesh.  If IPC is unavailable on any rank,
// every rank falls back to the NCCL path (agreed through an all-reduce).
// Arena layout: 4 x u64 {flag from lower, flag from upper, error, magic} | pad to 256 B |
//               landing[from lower][parity 0,1][P] | landing[from upper][parity 0,1][P]
// =====================================================================================
struct HaloRecord {                    // 128 bytes, exchanged with both neighbours
  cudaIpcMemHandle_t arena;
  unsigned long long step;
  long long plane;
  unsigned long long magic;
  int ok;
  char pad[128 - 64 - 3 * 8 - 4];
};
static_assert(sizeof(HaloRecord) == 128, "HaloRecord is 128 bytes");

struct HaloP2P {
  DevBuf<unsigned char> arena;
  unsigned char* lo_base = nullptr;    // mapped arena of rank-1 / rank+1
  unsigned char* hi_base = nullptr;
  unsigned long long* flags() { return reinterpret_cast<unsigned long long*>(arena.p); }
  // slot 0: the vector of the fused leapfrog / Poisson steps (the plane next to the interface) and the FIRST output of a
  // paired Tsit5 pass; slot 1: its second output, or the second plane of u of a two-step leapfrog launch; slot 2: the
  // plane of v of a two-step leapfrog launch
  static double* landing(unsigned char* base, int from_upper, int parity, int64_t P, int slot = 0) {
    return reinterpret_cast<double*>(base + 256) + (((int64_t)slot * 2 + from_upper) * 2 + parity) * P;
  }
};

// every block waits for the neighbours' step flags, then copies its share of the landed planes into u
__global__ void __launch_bounds__(256)
k_halo_wait_copy(unsigned long long* flags, unsigned long long need_lo, unsigned long long need_hi,
                 const double* __restrict__ land_lo, double* __restrict__ ghost_lo, const double* __restrict__ land_hi,
                 double* __restrict__ ghost_hi, int64_t P) {
  __shared__ int bad;
  if (threadIdx.x == 0) {
    bad = 0;
    if (*(volatile unsigned long long*)(flags + 2)) bad = 1;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!bad) {
      unsigned long long a, b;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(flags) : "memory");
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(b) : "l"(flags + 1) : "memory");
      if (a >= need_lo && b >= need_hi) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) { flags[2] = 1ull; bad = 1; }     // 20 s: the neighbour is gone
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (bad) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
    if (land_lo) ghost_lo[i] = __ldcv(land_lo + i);     // written by a peer: do not trust a cached line
    if (land_hi) ghost_hi[i] = __ldcv(land_hi + i);
  }
}

__global__ void k_halo_signal(unsigned long long* to_lower, unsigned long long* to_upper, unsigned long long step) {
  __threadfence_system();
  if (to_lower) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(to_lower), "l"(step) : "memory");
  if (to_upper) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(to_upper), "l"(step) : "memory");
}

void ddf_halo_lists_release(ddf_mesh* m);
void ddf_halo_release(ddf_mesh* m) {
  ddf_halo_lists_release(m);
  ddf_halo_gen_release(m);
  HaloP2P* hp = (HaloP2P*)m->halo_p2p;
  if (!hp) return;
  if (hp->lo_base) cudaIpcCloseMemHandle(hp->lo_base);
  if (hp->hi_base) cudaIpcCloseMemHandle(hp->hi_base);
  delete hp;
  m->halo_p2p = nullptr;
}

// first distributed step of a mesh: allocate the arena, exchange its handle with both neighbours, map theirs
static int halo_handshake_local(ddf_mesh* m, int* ok_out) {
  *ok_out = 0;
  ddf_ctx* ctx = m->ctx;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  const int64_t P = m->halo_plane;
  HaloP2P* hp = new HaloP2P();
  m->halo_p2p = hp;
  // a dedicated allocation of whole 2 MiB pages: cudaMalloc never shares those with other buffers, so the
  // IPC handle maps exactly this arena at offset 0
  const size_t need = 256 + 12 * (size_t)P * sizeof(double);      // 3 slots x 2 sides x 2 parities
  const size_t bytes = (need + (2u << 20) - 1) / (2u << 20) * (2u << 20);
  hp->arena.plain = true;                 // cudaMalloc: pool memory cannot be exported with cudaIpcGetMemHandle
  DDF_CHECK(hp->arena.alloc(ctx, bytes));
  DDF_CUDA(cudaMemsetAsync(hp->arena.p, 0, 256, ctx->stream));
  const unsigned long long magic = 0xDDF0B200ull * 1315423911ull + (unsigned long long)ctx->rank;
  DDF_CUDA(cudaMemcpyAsync(hp->flags() + 3, &magic, 8, cudaMemcpyHostToDevice, ctx->stream));
  HaloRecord* host = reinterpret_cast<HaloRecord*>(ctx->red_host);      // 3 records in the pinned scratch
  HaloRecord* dev = reinterpret_cast<HaloRecord*>(ctx->red_dev);
  memset(&host[0], 0, 3 * sizeof(HaloRecord));
  int ok = 1;
  if (cudaIpcGetMemHandle(&host[0].arena, hp->arena.p) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  host[0].step = m->halo_step;
  host[0].plane = P;
  host[0].magic = magic;
  host[0].ok = ok;
  DDF_CUDA(cudaMemcpyAsync(&dev[0], &host[0], sizeof(HaloRecord), cudaMemcpyHostToDevice, ctx->stream));
  DDF_NCCL(g_nccl.GroupStart());
  if (m->halo_lo > 0) {
    DDF_NCCL(g_nccl.Send(&dev[0], sizeof(HaloRecord), ncclChar, ctx->rank - 1, comm, ctx->stream));
    DDF_NCCL(g_nccl.Recv(&dev[1], sizeof(HaloRecord), ncclChar, ctx->rank - 1, comm, ctx->stream));
  }
  if (m->halo_hi > 0) {
    DDF_NCCL(g_nccl.Send(&dev[0], sizeof(HaloRecord), ncclChar, ctx->rank + 1, comm, ctx->stream));
    DDF_NCCL(g_nccl.Recv(&dev[2], sizeof(HaloRecord), ncclChar, ctx->rank + 1, comm, ctx->stream));
  }
  DDF_NCCL(g_nccl.GroupEnd());
  DDF_CUDA(cudaMemcpyAsync(&host[1], &dev[1], 2 * sizeof(HaloRecord), cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  const HaloRecord lo = host[1], hi = host[2];
  auto map = [&](const HaloRecord& r, unsigned char** base) {
    if (!r.ok || r.plane != P || r.step != m->halo_step) { ok = 0; return; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, r.arena, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; return; }
    *base = (unsigned char*)p;
    unsigned long long seen = 0;                       // the mapping must show the exporter's magic word
    if (cudaMemcpy(&seen, (unsigned long long*)p + 3, 8, cudaMemcpyDeviceToHost) != cudaSuccess || seen != r.magic) {
      cudaGetLastError();
      ok = 0;
    }
  };
  if (m->halo_lo > 0) map(lo, &hp->lo_base);
  if (m->halo_hi > 0 && ok) map(hi, &hp->hi_base);
  *ok_out = ok;
  return 0;
}

// Every rank reaches the agreeing all-reduce, also the one whose local part failed on the host (a CUDA / NCCL error,
// an allocation that did not fit): its neighbours would otherwise wait in that all-reduce for ever.  A local failure
// counts as "peer memory unusable" and the step falls back to ncclSend/ncclRecv on all ranks.
static int halo_handshake(ddf_mesh* m, int* usable) {
  int ok = 0;
  const int rc = halo_handshake_local(m, &ok);
  if (rc != 0) { ok = 0; cudaGetLastError(); }
  double agreed = ok ? 1.0 : 0.0;
  DDF_CHECK(ddf_comm_allreduce(m->ctx, &agreed, 2));    // min over ranks
  *usable = agreed > 0.5;
  return 0;
}


// kind 0: leapfrog (aux = v, coef = dt); kind 1: Poisson relaxation sweep (aux = rho, coef = theta)
static int steps_dist_p2p(ddf_mesh* m, int kind, ddf_vec* u, ddf_vec* v, double dt, int nsteps) {
  ddf_ctx* ctx = m->ctx;
  HaloP2P* hp = (HaloP2P*)m->halo_p2p;
  const int64_t nv = m->n[0], P = m->halo_plane;
  const int64_t own_lo = m->halo_lo * P, own_hi = nv - m->halo_hi * P;
  unsigned long long* flags = hp->flags();
  const bool has_lo = m->halo_lo > 0, has_hi = m->halo_hi > 0;
  const int cgrid = (int)std::min<int64_t>(32, ceil_div64(P, 256));
  unsigned long long err = 0;
  DDF_CUDA(cudaMemcpyAsync(&err, flags + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err) DDF_FAIL("halo: timed out waiting for a neighbour's step flag in an earlier call");
  int s = 0;
  // ---- two steps per launch with a deep halo (leapfrog only; every rank must be able to: agreed once per mesh)
  if (kind == 0 && nsteps >= 2 && m->wave2_dist_state == 0) {
    int ok = 0;
    const int rc = ddf_pipe2_feasible(m, &ok);
    if (rc != 0) { ok = 0; cudaGetLastError(); }
    // the deep halo needs two ghost planes per interior side and two owned planes to send
    if ((has_lo && m->halo_lo < 2) || (has_hi && m->halo_hi < 2) || own_hi - own_lo < 2 * P) ok = 0;
    double agreed = ok ? 1.0 : 0.0;
    DDF_CHECK(ddf_comm_allreduce(ctx, &agreed, 2));     // min over ranks
    m->wave2_dist_state = agreed > 0.5 ? 1 : 2;
    m->wave2_dist_tune = ctx->tune.wave;
  }
  if (kind == 0 && nsteps >= 2 && m->wave2_dist_state != 0 && m->wave2_dist_tune != ctx->tune.wave) {
    int ok = 0;                                          // the tuning changed since the agreement: ask again
    const int rc = ddf_pipe2_feasible(m, &ok);
    if (rc != 0) { ok = 0; cudaGetLastError(); }
    if ((has_lo && m->halo_lo < 2) || (has_hi && m->halo_hi < 2) || own_hi - own_lo < 2 * P) ok = 0;
    double agreed = ok ? 1.0 : 0.0;
    DDF_CHECK(ddf_comm_allreduce(ctx, &agreed, 2));
    m->wave2_dist_state = agreed > 0.5 ? 1 : 2;
    m->wave2_dist_tune = ctx->tune.wave;
  }
  if (kind == 0 && m->wave2_dist_state == 1 && nsteps >= 2) {
    {
      // the ghost planes the pairs read -- two of u, one of v -- are refreshed from their owners once per call
      ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
      double *uu = u->d.p, *vv = v->d.p;
      DDF_NCCL(g_nccl.GroupStart());
      if (has_lo) {
        DDF_NCCL(g_nccl.Send(uu + own_lo, 2 * P, ncclDouble, ctx->rank - 1, comm, ctx->stream));
        DDF_NCCL(g_nccl.Recv(uu + own_lo - 2 * P, 2 * P, ncclDouble, ctx->rank - 1, comm, ctx->stream));
        DDF_NCCL(g_nccl.Send(vv + own_lo, P, ncclDouble, ctx->rank - 1, comm, ctx->stream));
        DDF_NCCL(g_nccl.Recv(vv + own_lo - P, P, ncclDouble, ctx->rank - 1, comm, ctx->stream));
      }
      if (has_hi) {
        DDF_NCCL(g_nccl.Send(uu + own_hi - 2 * P, 2 * P, ncclDouble, ctx->rank + 1, comm, ctx->stream));
        DDF_NCCL(g_nccl.Recv(uu + own_hi, 2 * P, ncclDouble, ctx->rank + 1, comm, ctx->stream));
        DDF_NCCL(g_nccl.Send(vv + own_hi - P, P, ncclDouble, ctx->rank + 1, comm, ctx->stream));
        DDF_NCCL(g_nccl.Recv(vv + own_hi, P, ncclDouble, ctx->rank + 1, comm, ctx->stream));
      }
      DDF_NCCL(g_nccl.GroupEnd());
      ctx->launches += (has_lo ? 1 : 0) + (has_hi ? 1 : 0);
      bool first = true;
      while (nsteps - s >= 2) {
        const unsigned long long k = ++m->halo_step;
        const int par_in = (int)((k - 1) & 1), par_out = (int)(k & 1);
        uu = u->d.p;
        // wait for the neighbours' previous event; after the first pair take over the three planes they landed
        for (int slot = 0; slot < 3; slot++) {
          double* dst_lo = slot == 0 ? uu + own_lo - P : (slot == 1 ? uu + own_lo - 2 * P : vv + own_lo - P);
          double* dst_hi = slot == 0 ? uu + own_hi : (slot == 1 ? uu + own_hi + P : vv + own_hi);
          k_halo_wait_copy<<<cgrid, 256, 0, ctx->stream>>>(
              flags, has_lo ? k - 1 : 0ull, has_hi ? k - 1 : 0ull,
              !first && has_lo ? HaloP2P::landing(hp->arena.p, 0, par_in, P, slot) : nullptr, dst_lo,
              !first && has_hi ? HaloP2P::landing(hp->arena.p, 1, par_in, P, slot) : nullptr, dst_hi, P);
          DDF_LAUNCH_CHECK(ctx);
          if (first) break;                              // the one launch has waited; nothing to copy yet
        }
        double* peers[6];
        for (int slot = 0; slot < 3; slot++) {
          peers[slot] = has_lo ? HaloP2P::landing(hp->lo_base, 1, par_out, P, slot) : nullptr;
          peers[3 + slot] = has_hi ? HaloP2P::landing(hp->hi_base, 0, par_out, P, slot) : nullptr;
        }
        int ran = 0;
        DDF_CHECK(ddf_leapfrog_pair_launch(m, ctx->stream, uu, vv, m->utmp.p, dt, own_lo - (has_lo ? P : 0),
                                           own_hi + (has_hi ? P : 0), own_lo, own_hi, P, peers, &ran));
        if (!ran) DDF_FAIL("distributed step: the two-step launch all ranks agreed on was refused on this one");
        k_halo_signal<<<1, 1, 0, ctx->stream>>>(has_lo ? reinterpret_cast<unsigned long long*>(hp->lo_base) + 1 : nullptr,
                                                has_hi ? reinterpret_cast<unsigned long long*>(hp->hi_base) + 0 : nullptr, k);
        DDF_LAUNCH_CHECK(ctx);
        s += 2;
        first = false;
      }
    }
  }
  for (; s < nsteps; s++) {
    const unsigned long long k = ++m->halo_step;
    const int par_in = (int)((k - 1) & 1), par_out = (int)(k & 1);
    // the ghost planes are valid on entry (caller's contract); later steps take what the neighbours landed
    const bool copy = s > 0;
    double* uo = u->d.p;
    k_halo_wait_copy<<<cgrid, 256, 0, ctx->stream>>>(
        flags, has_lo ? k - 1 : 0ull, has_hi ? k - 1 : 0ull,
        copy && has_lo ? HaloP2P::landing(hp->arena.p, 0, par_in, P) : nullptr, uo + own_lo - P,
        copy && has_hi ? HaloP2P::landing(hp->arena.p, 1, par_in, P) : nullptr, uo + own_hi, P);
    DDF_LAUNCH_CHECK(ctx);
    // I am the UPPER neighbour of rank-1 and the LOWER neighbour of rank+1
    DDF_CHECK(ddf_step_launch(m, ctx->stream, kind, uo, v->d.p, m->utmp.p, dt, own_lo, own_hi,
                              has_lo ? HaloP2P::landing(hp->lo_base, 1, par_out, P) : nullptr,
                              has_hi ? HaloP2P::landing(hp->hi_base, 0, par_out, P) : nullptr, P));
    k_halo_signal<<<1, 1, 0, ctx->stream>>>(has_lo ? reinterpret_cast<unsigned long long*>(hp->lo_base) + 1 : nullptr,
                                            has_hi ? reinterpret_cast<unsigned long long*>(hp->hi_base) + 0 : nullptr, k);
    DDF_LAUNCH_CHECK(ctx);
    u->d.swap(m->utmp);
  }
  // leave with valid ghost planes: what the neighbours landed in their last step
  const unsigned long long k = m->halo_step;
  k_halo_wait_copy<<<cgrid, 256, 0, ctx->stream>>>(
      flags, has_lo ? k : 0ull, has_hi ? k : 0ull, has_lo ? HaloP2P::landing(hp->arena.p, 0, (int)(k & 1), P) : nullptr,
      u->d.p + own_lo - P, has_hi ? HaloP2P::landing(hp->arena.p, 1, (int)(k & 1), P) : nullptr, u->d.p + own_hi, P);
  DDF_LAUNCH_CHECK(ctx);
  // a wait that timed out inside THIS call left stale ghost planes behind: report it now, not in the next call
  DDF_CUDA(cudaMemcpyAsync(&err, flags + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err) DDF_FAIL("halo: timed out (20 s) waiting for a neighbour's step flag; the ghost planes are stale");
  return 0;
}

// ---- the same protocol, pass by pass, for the paired Tsit5 stages of wave.cu: a pass gathers TWO vertex vectors and
// stores two (slots 0 / 1 of the landing arenas).  begin: wait for the neighbours' flags >= k - 1 and (copy != 0) move what
// they landed in their pass k - 1 into the ghost planes of this pass's inputs; the caller then runs the pass on the owned
// rows with the returned peer pointers and calls end (flag = k).  finish leaves valid ghost planes in a / b.
int ddf_p2p_usable(ddf_mesh* m, int* usable) {
  ddf_ctx* ctx = m->ctx;
  *usable = 0;
  if (ctx->nranks == 1 || !m->halo_plane || m->halo_gen || (ctx->tune.wave & 1024)) return 0;
  if (m->halo_p2p_state == 0) {
    int ok = 0;
    DDF_CHECK(halo_handshake(m, &ok));
    m->halo_p2p_state = ok ? 1 : 2;
  }
  if (m->halo_p2p_state != 1) return 0;
  unsigned long long err = 0;
  DDF_CUDA(cudaMemcpyAsync(&err, ((HaloP2P*)m->halo_p2p)->flags() + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err) DDF_FAIL("halo: timed out waiting for a neighbour's step flag in an earlier call");
  *usable = 1;
  return 0;
}
int ddf_p2p_pass_begin(ddf_mesh* m, double* in_a, double* in_b, int copy, double** peers /* lo_a, hi_a, lo_b, hi_b */,
                       int64_t* own_lo_out, int64_t* own_hi_out) {
  ddf_ctx* ctx = m->ctx;
  HaloP2P* hp = (HaloP2P*)m->halo_p2p;
  const int64_t nv = m->n[0], P = m->halo_plane;
  const int64_t own_lo = m->halo_lo * P, own_hi = nv - m->halo_hi * P;
  const bool has_lo = m->halo_lo > 0, has_hi = m->halo_hi > 0;
  const int cgrid = (int)std::min<int64_t>(32, ceil_div64(P, 256));
  const unsigned long long k = ++m->halo_step;
  const int par_in = (int)((k - 1) & 1), par_out = (int)(k & 1);
  double* in[2] = {in_a, in_b};
  for (int slot = 0; slot < 2; slot++) {
    k_halo_wait_copy<<<cgrid, 256, 0, ctx->stream>>>(
        hp->flags(), has_lo ? k - 1 : 0ull, has_hi ? k - 1 : 0ull,
        copy && has_lo ? HaloP2P::landing(hp->arena.p, 0, par_in, P, slot) : nullptr, in[slot] + own_lo - P,
        copy && has_hi ? HaloP2P::landing(hp->arena.p, 1, par_in, P, slot) : nullptr, in[slot] + own_hi, P);
    DDF_LAUNCH_CHECK(ctx);
    if (!copy) break;                       // the first launch has waited; nothing to copy
  }
  // I am the UPPER neighbour of rank-1 and the LOWER neighbour of rank+1
  for (int slot = 0; slot < 2; slot++) {
    peers[2 * slot + 0] = has_lo ? HaloP2P::landing(hp->lo_base, 1, par_out, P, slot) : nullptr;
    peers[2 * slot + 1] = has_hi ? HaloP2P::landing(hp->hi_base, 0, par_out, P, slot) : nullptr;
  }
  *own_lo_out = own_lo;
  *own_hi_out = own_hi;
  m->halo_last = 1;
  return 0;
}
int ddf_p2p_pass_end(ddf_mesh* m) {
  ddf_ctx* ctx = m->ctx;
  HaloP2P* hp = (HaloP2P*)m->halo_p2p;
  const bool has_lo = m->halo_lo > 0, has_hi = m->halo_hi > 0;
  k_halo_signal<<<1, 1, 0, ctx->stream>>>(has_lo ? reinterpret_cast<unsigned long long*>(hp->lo_base) + 1 : nullptr,
                                          has_hi ? reinterpret_cast<unsigned long long*>(hp->hi_base) + 0 : nullptr, m->halo_step);
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}
int ddf_p2p_finish(ddf_mesh* m, double* a, double* b) {
  ddf_ctx* ctx = m->ctx;
  HaloP2P* hp = (HaloP2P*)m->halo_p2p;
  const int64_t nv = m->n[0], P = m->halo_plane;
  const int64_t own_lo = m->halo_lo * P, own_hi = nv - m->halo_hi * P;
  const bool has_lo = m->halo_lo > 0, has_hi = m->halo_hi > 0;
  const int cgrid = (int)std::min<int64_t>(32, ceil_div64(P, 256));
  const unsigned long long k = m->halo_step;
  double* out[2] = {a, b};
  for (int slot = 0; slot < 2; slot++) {
    if (!out[slot]) continue;
    k_halo_wait_copy<<<cgrid, 256, 0, ctx->stream>>>(
        hp->flags(), has_lo ? k : 0ull, has_hi ? k : 0ull,
        has_lo ? HaloP2P::landing(hp->arena.p, 0, (int)(k & 1), P, slot) : nullptr, out[slot] + own_lo - P,
        has_hi ? HaloP2P::landing(hp->arena.p, 1, (int)(k & 1), P, slot) : nullptr, out[slot] + own_hi, P);
    DDF_LAUNCH_CHECK(ctx);
  }
  unsigned long long err = 0;
  DDF_CUDA(cudaMemcpyAsync(&err, hp->flags() + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err) DDF_FAIL("halo: timed out (20 s) waiting for a neighbour's step flag; the ghost planes are stale");
  return 0;
}

// =====================================================================================
// Ghost exchange of a cochain of ANY rank ("rows owned locally need x on ghost
// k-simplices owned by r+-1").  A k-simplex is owned by the rank that owns its smallest vertex.
// The slabs of two neighbours overlap in whole cell layers and local vertex ids differ by a
// constant, so -- local simplex ids being lexicographic ranks of vertex tuples
// (Manifolds.jl:735-757) -- the simplices of the overlap appear in the SAME relative order on
// both ranks.  Four index lists per rank R therefore suffice (built once per mesh and R by a
// flag + stream-compaction pass over the vertex table):
//   recv_lo  ghosts below:  min vertex in the lower ghost planes
//   recv_hi  ghosts above:  min vertex in the upper ghost planes
//   send_lo  what rank-1 holds of mine as ITS upper ghosts: min vertex owned, every vertex inside
//            rank-1's top planes (its ghost_hi, exchanged once)
//   send_hi  what rank+1 holds as ITS lower ghosts: min vertex in my last (its ghost_lo) owned planes
// and the i-th entry of my send list is the i-th entry of the neighbour's recv list.
// Exchange = gather -> ncclSend/ncclRecv in one group -> scatter.
// =====================================================================================
struct HaloLists {
  struct PerRank {
    bool built = false;
    DevBuf<uint32_t> idx[4];           // 0 send_lo, 1 send_hi, 2 recv_lo, 3 recv_hi
    int64_t cnt[4] = {0, 0, 0, 0};
    DevBuf<double> buf[4];
  };
  PerRank r[DDF_MAXD + 1];
  int64_t peer_lo_ghost_hi = 0, peer_hi_ghost_lo = 0;
  bool peers_known = false;
};

void ddf_halo_lists_release(ddf_mesh* m) {
  delete (HaloLists*)m->halo_lists;
  m->halo_lists = nullptr;
}

// flags[c][i] = 1 when simplex i belongs to list c
__global__ void k_halo_flags(const int32_t* __restrict__ simp, int N, int64_t n, int64_t P, int64_t nplanes,
                             int64_t ghost_lo, int64_t ghost_hi, int64_t peer_lo_ghost_hi, int64_t peer_hi_ghost_lo,
                             uint8_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t vmin = simp ? simp[i * N] : i, vmax = simp ? simp[i * N + N - 1] : i;
  const int64_t pmin = vmin / P, pmax = vmax / P;
  const int64_t own_hi = nplanes - ghost_hi;         // first upper ghost plane
  flags[0 * n + i] = (ghost_lo > 0 && pmin >= ghost_lo && pmax <= ghost_lo + peer_lo_ghost_hi - 1) ? 1 : 0;
  flags[1 * n + i] = (ghost_hi > 0 && pmin >= own_hi - peer_hi_ghost_lo && pmin < own_hi) ? 1 : 0;
  flags[2 * n + i] = (ghost_lo > 0 && pmin < ghost_lo) ? 1 : 0;
  flags[3 * n + i] = (ghost_hi > 0 && pmin >= own_hi) ? 1 : 0;
}
__global__ void k_pack_f64(const uint32_t* __restrict__ idx, const double* __restrict__ src, int64_t n, double* __restrict__ dst) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = src[idx[p]];
}
__global__ void k_unpack_f64(const uint32_t* __restrict__ idx, const double* __restrict__ src, int64_t n, double* __restrict__ dst) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[idx[p]] = src[p];
}

static int halo_lists_build(ddf_mesh* m, int R) {
  ddf_ctx* ctx = m->ctx;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  if (!m->halo_lists) m->halo_lists = new HaloLists();
  HaloLists* hl = (HaloLists*)m->halo_lists;
  const int64_t P = m->halo_plane, nplanes = m->n[0] / P;
  // how many of my planes each neighbour keeps as ghosts (once per mesh)
  if (!hl->peers_known) {
    long long* host = reinterpret_cast<long long*>(ctx->red_host);
    long long* dev = reinterpret_cast<long long*>(ctx->red_dev);
    host[0] = m->halo_lo;
    host[1] = m->halo_hi;
    host[2] = host[3] = host[4] = host[5] = 0;
    DDF_CUDA(cudaMemcpyAsync(dev, host, 48, cudaMemcpyHostToDevice, ctx->stream));
    DDF_NCCL(g_nccl.GroupStart());
    if (m->halo_lo > 0) {
      DDF_NCCL(g_nccl.Send(dev, 2, ncclInt64, ctx->rank - 1, comm, ctx->stream));
      DDF_NCCL(g_nccl.Recv(dev + 2, 2, ncclInt64, ctx->rank - 1, comm, ctx->stream));
    }
    if (m->halo_hi > 0) {
      DDF_NCCL(g_nccl.Send(dev, 2, ncclInt64, ctx->rank + 1, comm, ctx->stream));
      DDF_NCCL(g_nccl.Recv(dev + 4, 2, ncclInt64, ctx->rank + 1, comm, ctx->stream));
    }
    DDF_NCCL(g_nccl.GroupEnd());
    DDF_CUDA(cudaMemcpyAsync(host, dev, 48, cudaMemcpyDeviceToHost, ctx->stream));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
    hl->peer_lo_ghost_hi = host[3];          // rank-1's ghost_hi
    hl->peer_hi_ghost_lo = host[4];          // rank+1's ghost_lo
    hl->peers_known = true;
  }
  HaloLists::PerRank& pr = hl->r[R];
  const int64_t n = m->n[R];
  DevBuf<uint8_t> flags, temp;
  DevBuf<int64_t> nsel;
  DDF_CHECK(flags.alloc(ctx, 4 * (size_t)n));
  DDF_CHECK(nsel.alloc(ctx, 1));
  if (n) {
    int grid;
    DDF_CHECK(grid_for(n, 256, &grid));
    k_halo_flags<<<grid, 256, 0, ctx->stream>>>(R >= 1 ? m->simp[R].p : nullptr, R + 1, n, P, nplanes, m->halo_lo, m->halo_hi,
                                                hl->peer_lo_ghost_hi, hl->peer_hi_ghost_lo, flags.p);
    DDF_LAUNCH_CHECK(ctx);
  }
  for (int c = 0; c < 4; c++) {
    DevBuf<uint32_t> all;
    DDF_CHECK(all.alloc(ctx, n));
    int64_t h = 0;
    if (n) {
      cub::CountingInputIterator<uint32_t> it(0);
      size_t tb = 0;
      DDF_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, it, flags.p + (size_t)c * n, all.p, nsel.p, n, ctx->stream));
      DDF_CHECK(temp.alloc(ctx, tb));
      DDF_CUDA(cub::DeviceSelect::Flagged(temp.p, tb, it, flags.p + (size_t)c * n, all.p, nsel.p, n, ctx->stream));
      ctx->launches++;
      DDF_CUDA(cudaMemcpyAsync(&h, nsel.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
      DDF_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    pr.cnt[c] = h;
    DDF_CHECK(pr.idx[c].alloc(ctx, h));
    DDF_CHECK(pr.buf[c].alloc(ctx, h));
    if (h) DDF_CUDA(cudaMemcpyAsync(pr.idx[c].p, all.p, h * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  // the neighbour's recv list must be as long as my send list
  {
    long long* host = reinterpret_cast<long long*>(ctx->red_host);
    long long* dev = reinterpret_cast<long long*>(ctx->red_dev);
    host[0] = pr.cnt[0]; host[1] = pr.cnt[1]; host[2] = host[3] = -1;
    DDF_CUDA(cudaMemcpyAsync(dev, host, 32, cudaMemcpyHostToDevice, ctx->stream));
    DDF_NCCL(g_nccl.GroupStart());
    if (m->halo_lo > 0) {
      DDF_NCCL(g_nccl.Send(dev, 1, ncclInt64, ctx->rank - 1, comm, ctx->stream));        // my send_lo count
      DDF_NCCL(g_nccl.Recv(dev + 2, 1, ncclInt64, ctx->rank - 1, comm, ctx->stream));    // its send_hi count
    }
    if (m->halo_hi > 0) {
      DDF_NCCL(g_nccl.Send(dev + 1, 1, ncclInt64, ctx->rank + 1, comm, ctx->stream));    // my send_hi count
      DDF_NCCL(g_nccl.Recv(dev + 3, 1, ncclInt64, ctx->rank + 1, comm, ctx->stream));    // its send_lo count
    }
    DDF_NCCL(g_nccl.GroupEnd());
    DDF_CUDA(cudaMemcpyAsync(host, dev, 32, cudaMemcpyDeviceToHost, ctx->stream));
    DDF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (m->halo_lo > 0 && host[2] != pr.cnt[2])
      DDF_FAIL("halo_exchange_k(R=%d): rank %d sends %lld simplices, %lld ghosts expected below", R, ctx->rank - 1, host[2], (long long)pr.cnt[2]);
    if (m->halo_hi > 0 && host[3] != pr.cnt[3])
      DDF_FAIL("halo_exchange_k(R=%d): rank %d sends %lld simplices, %lld ghosts expected above", R, ctx->rank + 1, host[3], (long long)pr.cnt[3]);
  }
  pr.built = true;
  return 0;
}

extern "C" int ddf_halo_exchange_k(ddf_mesh* m, ddf_vec* x, int R) {
  DDF_TRY_BEGIN
  ddf_ctx* ctx = m->ctx;
  if (R < 0 || R > m->D) DDF_FAIL("halo_exchange_k: 0 <= R <= D");
  if (x->n != m->n[R]) DDF_FAIL("halo_exchange_k: vector has length %lld, the mesh has %lld %d-simplices", (long long)x->n, (long long)m->n[R], R);
  if (m->halo_gen) {
    // vertex-range partition of a general mesh: ghost lists exist for vertices only
    if (R != 0) DDF_FAIL("halo_exchange_k: a mesh partitioned by vertex ranges exchanges 0-cochains only (R = %d)", R);
    DDF_CHECK(ddf_vec_settle(x));
    return halo_exchange_general(m, x->d.p, ctx->stream);
  }
  if (ctx->nranks == 1) return 0;
  if (!ctx->nccl_comm) DDF_FAIL("halo: communicator not initialised (ddf_comm_init)");
  if (!m->halo_plane) DDF_FAIL("halo: slab not described (ddf_mesh_set_halo)");
  DDF_CHECK(ddf_mesh_require_assembled(m));
  DDF_CHECK(ddf_vec_settle(x));
  if (!m->halo_lists || !((HaloLists*)m->halo_lists)->r[R].built) DDF_CHECK(halo_lists_build(m, R));
  HaloLists::PerRank& pr = ((HaloLists*)m->halo_lists)->r[R];
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  int grid;
  for (int c = 0; c < 2; c++)
    if (pr.cnt[c]) {
      DDF_CHECK(grid_for(pr.cnt[c], 256, &grid));
      k_pack_f64<<<grid, 256, 0, ctx->stream>>>(pr.idx[c].p, x->d.p, pr.cnt[c], pr.buf[c].p);
      DDF_LAUNCH_CHECK(ctx);
    }
  DDF_NCCL(g_nccl.GroupStart());
  if (m->halo_lo > 0) {
    DDF_NCCL(g_nccl.Send(pr.buf[0].p, pr.cnt[0], ncclDouble, ctx->rank - 1, comm, ctx->stream));
    DDF_NCCL(g_nccl.Recv(pr.buf[2].p, pr.cnt[2], ncclDouble, ctx->rank - 1, comm, ctx->stream));
  }
  if (m->halo_hi > 0) {
    DDF_NCCL(g_nccl.Send(pr.buf[1].p, pr.cnt[1], ncclDouble, ctx->rank + 1, comm, ctx->stream));
    DDF_NCCL(g_nccl.Recv(pr.buf[3].p, pr.cnt[3], ncclDouble, ctx->rank + 1, comm, ctx->stream));
  }
  DDF_NCCL(g_nccl.GroupEnd());
  ctx->launches += (m->halo_lo > 0) + (m->halo_hi > 0);
  for (int c = 2; c < 4; c++)
    if (pr.cnt[c]) {
      DDF_CHECK(grid_for(pr.cnt[c], 256, &grid));
      k_unpack_f64<<<grid, 256, 0, ctx->stream>>>(pr.idx[c].p, pr.buf[c].p, pr.cnt[c], x->d.p);
      DDF_LAUNCH_CHECK(ctx);
    }
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_halo_mode(ddf_mesh* m, int* mode) {
  *mode = m->halo_last;
  return 0;
}

static int steps_dist(ddf_mesh* m, int kind, ddf_vec* u, ddf_vec* v, double dt, int nsteps) {
  ddf_ctx* ctx = m->ctx;
  DDF_CHECK(ddf_mesh_require_wave(m));
  if (kind == 1) DDF_CHECK(ddf_mesh_require_dinv(m));
  const int64_t nv = m->n[0];
  if (u->n != nv || v->n != nv) DDF_FAIL("distributed step: vectors must have length n0");
  if (u == v) DDF_FAIL("distributed step: the two vectors must not alias");
  if (!m->halo_plane && !m->halo_gen) DDF_FAIL("distributed step: partition not described (ddf_mesh_set_halo / ddf_mesh_set_halo_lists)");
  DDF_CHECK(ddf_vec_settle(u));
  DDF_CHECK(ddf_vec_settle(v));
  if ((int64_t)m->utmp.n != nv) {
    DDF_CHECK(m->utmp.alloc(ctx, nv));
    // ghost planes that are never written must still hold finite data
    DDF_CUDA(cudaMemcpyAsync(m->utmp.p, u->d.p, nv * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (m->halo_gen) {
    // general partition: the owned rows in one launch, then the list-driven ghost exchange of the new u
    m->halo_last = 2;
    for (int s = 0; s < nsteps; s++) {
      DDF_CHECK(ddf_step_launch(m, ctx->stream, kind, u->d.p, v->d.p, m->utmp.p, dt, m->gen_own_lo, m->gen_own_hi, nullptr,
                                nullptr, 0));
      DDF_CHECK(halo_exchange_general(m, m->utmp.p, ctx->stream));
      u->d.swap(m->utmp);
    }
    return 0;
  }
  const int64_t P = m->halo_plane;
  const int64_t own_lo = m->halo_lo * P, own_hi = nv - m->halo_hi * P;
  if (nsteps <= 0) return 0;
  // fused peer-memory halo (default when every rank can map its neighbours; tuning bit 1024 forces NCCL)
  if (ctx->nranks > 1 && !(ctx->tune.wave & 1024)) {
    if (m->halo_p2p_state == 0) {
      int usable = 0;
      DDF_CHECK(halo_handshake(m, &usable));
      m->halo_p2p_state = usable ? 1 : 2;
    }
    if (m->halo_p2p_state == 1) {
      m->halo_last = 1;
      return steps_dist_p2p(m, kind, u, v, dt, nsteps);
    }
  }
  m->halo_last = 2;
  for (int s = 0; s < nsteps; s++) {
    const double* uo = u->d.p;
    double* un = m->utmp.p;
    // interface planes first, so that their transfer overlaps the interior rows
    int64_t lo_end = own_lo, hi_beg = own_hi;
    if (m->halo_lo > 0) {
      lo_end = own_lo + P < own_hi ? own_lo + P : own_hi;
      DDF_CHECK(ddf_step_launch(m, ctx->stream, kind, uo, v->d.p, un, dt, own_lo, lo_end, nullptr, nullptr, P));
    }
    if (m->halo_hi > 0) {
      hi_beg = own_hi - P > lo_end ? own_hi - P : lo_end;
      DDF_CHECK(ddf_step_launch(m, ctx->stream, kind, uo, v->d.p, un, dt, hi_beg, own_hi, nullptr, nullptr, P));
    }
    DDF_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    DDF_CHECK(ddf_step_launch(m, ctx->stream, kind, uo, v->d.p, un, dt, lo_end, hi_beg, nullptr, nullptr, P));
    DDF_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
    DDF_CHECK(halo_exchange_on(m, un, ctx->comm_stream));
    DDF_CUDA(cudaEventRecord(ctx->ev_b, ctx->comm_stream));
    DDF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
    u->d.swap(m->utmp);
  }
  return 0;
}

extern "C" int ddf_wave_leapfrog_dist(ddf_mesh* m, ddf_vec* u, ddf_vec* v, double dt, int nsteps) {
  DDF_TRY_BEGIN
  if (m->ctx->nranks == 1 && !m->halo_plane && !m->halo_gen) return ddf_wave_leapfrog(m, u, v, dt, nsteps);
  return steps_dist(m, 0, u, v, dt, nsteps);
  DDF_TRY_END
}

// the fused Poisson relaxation sweep on slab meshes: same halo protocols as the wave step (peer-memory stores of the
// interface planes + step flags, or NCCL send/recv), u ping-pong, rho read-only
extern "C" int ddf_poisson_jacobi_dist(ddf_mesh* m, ddf_vec* u, ddf_vec* rho, double theta, int nsweeps) {
  DDF_TRY_BEGIN
  if (m->ctx->nranks == 1 && !m->halo_plane && !m->halo_gen) return ddf_poisson_jacobi(m, u, rho, theta, nsweeps);
  return steps_dist(m, 1, u, rho, theta, nsweeps);
  DDF_TRY_END
}

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

#include "mesh.cuh"

int ddf_leapfrog_launch(ddf_mesh* m, cudaStream_t stream, const double* u, double* v, double* unew, double dt,
                        int64_t row0, int64_t row1);   // wave.cu

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

extern "C" int ddf_mesh_set_halo(ddf_mesh* m, int64_t plane, int64_t ghost_lo, int64_t ghost_hi) {
  if (plane <= 0 || ghost_lo < 0 || ghost_hi < 0 || (ghost_lo + ghost_hi + 1) * plane > m->n[0] || m->n[0] % plane)
    DDF_FAIL("set_halo: inconsistent slab description (plane=%lld, lo=%lld, hi=%lld, V=%lld)", (long long)plane,
             (long long)ghost_lo, (long long)ghost_hi, (long long)m->n[0]);
  m->halo_plane = plane;
  m->halo_lo = ghost_lo;
  m->halo_hi = ghost_hi;
  return 0;
}

// exchange on `stream`: my lowest owned plane -> rank-1 (into its first upper ghost plane),
// my highest owned plane -> rank+1 (into its last lower ghost plane), and the converse receives.
static int halo_exchange_on(ddf_mesh* m, double* u, cudaStream_t stream) {
  ddf_ctx* ctx = m->ctx;
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

extern "C" int ddf_halo_exchange(ddf_mesh* m, ddf_vec* u) {
  DDF_TRY_BEGIN
  if (u->n != m->n[0]) DDF_FAIL("halo_exchange: vector must be a vertex cochain");
  DDF_CHECK(ddf_vec_settle(u));
  return halo_exchange_on(m, u->d.p, m->ctx->stream);
  DDF_TRY_END
}

extern "C" int ddf_wave_leapfrog_dist(ddf_mesh* m, ddf_vec* u, ddf_vec* v, double dt, int nsteps) {
  DDF_TRY_BEGIN
  ddf_ctx* ctx = m->ctx;
  if (ctx->nranks == 1 && !m->halo_plane) return ddf_wave_leapfrog(m, u, v, dt, nsteps);
  DDF_CHECK(ddf_mesh_require_wave(m));
  const int64_t nv = m->n[0];
  if (u->n != nv || v->n != nv) DDF_FAIL("wave_leapfrog_dist: vectors must have length n0");
  if (!m->halo_plane) DDF_FAIL("wave_leapfrog_dist: slab not described (ddf_mesh_set_halo)");
  DDF_CHECK(ddf_vec_settle(u));
  DDF_CHECK(ddf_vec_settle(v));
  if ((int64_t)m->utmp.n != nv) {
    DDF_CHECK(m->utmp.alloc(ctx, nv));
    // ghost planes that are never written must still hold finite data
    DDF_CUDA(cudaMemcpyAsync(m->utmp.p, u->d.p, nv * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  const int64_t P = m->halo_plane;
  const int64_t own_lo = m->halo_lo * P, own_hi = nv - m->halo_hi * P;
  for (int s = 0; s < nsteps; s++) {
    const double* uo = u->d.p;
    double* un = m->utmp.p;
    // interface planes first, so that their transfer overlaps the interior rows
    int64_t lo_end = own_lo, hi_beg = own_hi;
    if (m->halo_lo > 0) { lo_end = own_lo + P < own_hi ? own_lo + P : own_hi; DDF_CHECK(ddf_leapfrog_launch(m, ctx->stream, uo, v->d.p, un, dt, own_lo, lo_end)); }
    if (m->halo_hi > 0) { hi_beg = own_hi - P > lo_end ? own_hi - P : lo_end; DDF_CHECK(ddf_leapfrog_launch(m, ctx->stream, uo, v->d.p, un, dt, hi_beg, own_hi)); }
    DDF_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    DDF_CHECK(ddf_leapfrog_launch(m, ctx->stream, uo, v->d.p, un, dt, lo_end, hi_beg));
    DDF_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
    DDF_CHECK(halo_exchange_on(m, un, ctx->comm_stream));
    DDF_CUDA(cudaEventRecord(ctx->ev_b, ctx->comm_stream));
    DDF_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
    u->d.swap(m->utmp);
  }
  return 0;
  DDF_TRY_END
}

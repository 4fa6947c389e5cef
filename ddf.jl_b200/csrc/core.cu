// core.cu -- context, error state, device allocation accounting, cochain
// vectors (Fun.values, src/Funs.jl:16-31) with their axpy-class arithmetic
// (src/Funs.jl:156-188, src/SparseOps.jl:108-119) and reductions
// (ManifoldOps.jl:150-164; norms used at examples/wave.jl:127,162).
#include "common.cuh"
#include "mesh.cuh"

#include <algorithm>
#include <mutex>

static thread_local char g_err[1024] = "";

// Tuning selectors live in the context (DdfTune).  ddf_set_tuning -- the process-wide convenience the probes and
// tests use -- writes the defaults new contexts start from and every live context, under a lock; a thread that
// drives its own context and wants its own variant calls ddf_ctx_set_tuning.
static std::mutex g_ctx_mu;
static std::vector<ddf_ctx*> g_ctxs;
static DdfTune g_tune_default;

static int tune_set(DdfTune& t, const char* key, int value) {
  if (!strcmp(key, "fixed")) t.fixed = value;
  else if (!strcmp(key, "wave")) t.wave = value;
  else if (!strcmp(key, "asm")) t.asm_path = value;
  else if (!strcmp(key, "packbits")) t.packbits = value >= 1 && value <= 24 ? value : 24;
  else if (!strcmp(key, "prefetch")) t.prefetch = value >= 0 ? value : 0;
  else DDF_FAIL("set_tuning: unknown key %s", key);
  return 0;
}
extern "C" int ddf_set_tuning(const char* key, int value) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DDF_CHECK(tune_set(g_tune_default, key, value));
  for (ddf_ctx* c : g_ctxs) tune_set(c->tune, key, value);
  return 0;
}
extern "C" int ddf_ctx_set_tuning(ddf_ctx* ctx, const char* key, int value) { return tune_set(ctx->tune, key, value); }

void ddf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* ddf_last_error(void) { return g_err; }
extern "C" int ddf_version(void) { return 100; }

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync / cudaFreeAsync on the compute stream)
// with an unlimited release threshold: a freed table goes back to the pool, not to the driver, so building and
// dropping multi-GB tables (assembly workspace, operator storages, cochains) costs no cudaMalloc / cudaFree round
// trip -- and cudaFree's implicit device synchronisation is gone.  Buffers are used on the library's other streams
// only behind event dependencies on the compute stream, and every destroy path synchronises before it frees.
// `plain` = cudaMalloc (memory that is exported through CUDA IPC must not come from a pool).
int ddf_dev_alloc(ddf_ctx* ctx, void** p, size_t bytes, bool plain) {
  *p = nullptr;
  if (bytes == 0) return 0;
  bytes = (bytes + 15) & ~(size_t)15;      // 16-byte TMA granules may read the pad element of an odd-length vector
  cudaError_t e = plain ? cudaMalloc(p, bytes) : cudaMallocAsync(p, bytes, ctx->stream);
  if (e != cudaSuccess && !plain) {        // out of memory with cached blocks in the pool: give them back and retry
    cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    e = cudaMallocAsync(p, bytes, ctx->stream);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    ddf_set_error("device allocation of %zu bytes failed: %s (in use: %lld bytes)", bytes,
                  cudaGetErrorString(e), (long long)ctx->bytes_in_use);
    return 1;
  }
  ctx->bytes_in_use += (int64_t)bytes;
  return 0;
}

void ddf_dev_free(ddf_ctx* ctx, void* p, size_t bytes, bool plain) {
  if (!p) return;
  if (plain || !ctx) cudaFree(p);
  else cudaFreeAsync(p, ctx->stream);
  if (ctx) ctx->bytes_in_use -= (int64_t)((bytes + 15) & ~(size_t)15);
}

// give the cached blocks of the pool back to the driver (e.g. before another allocator in the process needs them)
extern "C" int ddf_ctx_trim(ddf_ctx* c) {
  DDF_CUDA(cudaSetDevice(c->device));
  DDF_CHECK(ddf_sync(c));
  cudaMemPool_t pool;
  DDF_CUDA(cudaDeviceGetDefaultMemPool(&pool, c->device));
  DDF_CUDA(cudaMemPoolTrimTo(pool, 0));
  return 0;
}

extern "C" int ddf_ctx_create(int device, ddf_ctx** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    DDF_FAIL("no CUDA device available (%s); libddf_b200 has no CPU fallback",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0 || device >= ndev) DDF_FAIL("device %d out of range (have %d)", device, ndev);
  cudaDeviceProp prop;
  DDF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    DDF_FAIL("device %d is sm_%d%d; libddf_b200 is built for sm_100a only (no fallback)", device,
             prop.major, prop.minor);
  DDF_CUDA(cudaSetDevice(device));
  {
    cudaMemPool_t pool;
    DDF_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;            // freed blocks stay cached in the pool
    DDF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  ddf_ctx* c = new ddf_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  DDF_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  DDF_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  DDF_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
  DDF_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  DDF_CUDA(cudaEventCreate(&c->ev_tic));
  DDF_CUDA(cudaEventCreate(&c->ev_toc));
  DDF_CUDA(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
  DDF_CUDA(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
  DDF_CHECK(ddf_dev_alloc(c, (void**)&c->red_dev, 4096 * sizeof(double)));
  DDF_CUDA(cudaMallocHost((void**)&c->red_host, 4096 * sizeof(double)));
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    c->tune = g_tune_default;
    g_ctxs.push_back(c);
  }
  *out = c;
  return 0;
  DDF_TRY_END
}

extern "C" int ddf_ctx_destroy(ddf_ctx* c) {
  if (!c) return 0;
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    g_ctxs.erase(std::remove(g_ctxs.begin(), g_ctxs.end(), c), g_ctxs.end());
  }
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  ddf_comm_destroy(c);
  if (c->flush_buf) ddf_dev_free(c, c->flush_buf, c->flush_bytes);
  ddf_dev_free(c, c->red_dev, 4096 * sizeof(double));
  cudaStreamSynchronize(c->stream);          // the stream-ordered frees above
  cudaFreeHost(c->red_host);
  cudaEventDestroy(c->ev_tic);
  cudaEventDestroy(c->ev_toc);
  cudaEventDestroy(c->ev_a);
  cudaEventDestroy(c->ev_b);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->comm_stream);
  cudaStreamDestroy(c->h2d_stream);
  cudaStreamDestroy(c->d2h_stream);
  delete c;
  return 0;
}

extern "C" int ddf_sync(ddf_ctx* c) {
  DDF_CUDA(cudaSetDevice(c->device));
  DDF_CUDA(cudaStreamSynchronize(c->stream));
  DDF_CUDA(cudaStreamSynchronize(c->comm_stream));
  DDF_CUDA(cudaStreamSynchronize(c->h2d_stream));
  DDF_CUDA(cudaStreamSynchronize(c->d2h_stream));
  return 0;
}

extern "C" int ddf_ctx_bytes_in_use(ddf_ctx* c, int64_t* out) { *out = c->bytes_in_use; return 0; }
extern "C" int ddf_ctx_launch_count(ddf_ctx* c, int64_t* out) { *out = c->launches; return 0; }

extern "C" int ddf_timer_tic(ddf_ctx* c) {
  DDF_CUDA(cudaEventRecord(c->ev_tic, c->stream));
  return 0;
}
extern "C" int ddf_timer_toc(ddf_ctx* c, double* ms) {
  DDF_CUDA(cudaEventRecord(c->ev_toc, c->stream));
  DDF_CUDA(cudaEventSynchronize(c->ev_toc));
  float f = 0;
  DDF_CUDA(cudaEventElapsedTime(&f, c->ev_tic, c->ev_toc));
  *ms = (double)f;
  return 0;
}

extern "C" int ddf_event_record(ddf_ctx* c, int slot) {
  if (slot < 0 || slot >= 4096) DDF_FAIL("event_record: slot out of range");
  if (c->events.size() <= (size_t)slot) c->events.resize(slot + 1, nullptr);
  if (!c->events[slot]) DDF_CUDA(cudaEventCreate(&c->events[slot]));
  DDF_CUDA(cudaEventRecord(c->events[slot], c->stream));
  return 0;
}
extern "C" int ddf_event_elapsed(ddf_ctx* c, int a, int b, double* ms) {
  if (a < 0 || b < 0 || (size_t)a >= c->events.size() || (size_t)b >= c->events.size() || !c->events[a] || !c->events[b])
    DDF_FAIL("event_elapsed: slot not recorded");
  DDF_CUDA(cudaEventSynchronize(c->events[b]));
  float f = 0;
  DDF_CUDA(cudaEventElapsedTime(&f, c->events[a], c->events[b]));
  *ms = (double)f;
  return 0;
}

__global__ void k_fill_u32(uint32_t* p, size_t n, uint32_t v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v + (uint32_t)i;
}

extern "C" int ddf_flush_l2(ddf_ctx* c) {
  DDF_TRY_BEGIN
  if (!c->flush_buf) {
    c->flush_bytes = (size_t)512 << 20;   // 512 MiB > 126 MB L2
    DDF_CHECK(ddf_dev_alloc(c, &c->flush_buf, c->flush_bytes));
  }
  k_fill_u32<<<c->sm_count * 8, 256, 0, c->stream>>>((uint32_t*)c->flush_buf, c->flush_bytes / 4, 1u);
  DDF_LAUNCH_CHECK(c);
  return 0;
  DDF_TRY_END
}

// ------------------------------------------------------------------ vectors
extern "C" int ddf_vec_create(ddf_ctx* ctx, int64_t n, ddf_vec** out) {
  DDF_TRY_BEGIN
  *out = nullptr;
  if (n < 0) DDF_FAIL("negative vector length");
  ddf_vec* v = new ddf_vec();
  v->ctx = ctx;
  v->n = n;
  if (v->d.alloc(ctx, (size_t)n)) { delete v; return 1; }
  if (n) DDF_CUDA(cudaMemsetAsync(v->d.p, 0, n * sizeof(double), ctx->stream));
  *out = v;
  return 0;
  DDF_TRY_END
}
extern "C" int ddf_vec_destroy(ddf_vec* v) {
  if (!v) return 0;
  cudaStreamSynchronize(v->ctx->stream);
  if (v->up_pending) cudaEventSynchronize(v->ev_up);
  if (v->down_pending) cudaEventSynchronize(v->ev_down);
  if (v->ev_up) cudaEventDestroy(v->ev_up);
  if (v->ev_down) cudaEventDestroy(v->ev_down);
  delete v;
  return 0;
}
extern "C" int ddf_vec_length(ddf_vec* v, int64_t* out) { *out = v->n; return 0; }

extern "C" int ddf_vec_upload(ddf_vec* v, const double* host) {
  if (v->n == 0) return 0;
  DDF_CHECK(ddf_vec_settle(v));
  DDF_CUDA(cudaMemcpyAsync(v->d.p, host, v->n * sizeof(double), cudaMemcpyHostToDevice, v->ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(v->ctx->stream));   // host pointer is borrowed for the call only
  return 0;
}
extern "C" int ddf_vec_download(ddf_vec* v, double* host) {
  if (v->n == 0) return 0;
  DDF_CHECK(ddf_vec_settle(v));
  DDF_CUDA(cudaMemcpyAsync(host, v->d.p, v->n * sizeof(double), cudaMemcpyDeviceToHost, v->ctx->stream));
  DDF_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return 0;
}

// ---- asynchronous transfers on dedicated copy streams (full-duplex PCIe, overlapped with compute).
// Ordering is tracked per vector: kernels that touch v first wait for its pending copies.
int ddf_vec_settle(ddf_vec* v) {
  ddf_ctx* c = v->ctx;
  if (v->up_pending) { DDF_CUDA(cudaStreamWaitEvent(c->stream, v->ev_up, 0)); v->up_pending = false; }
  if (v->down_pending) { DDF_CUDA(cudaStreamWaitEvent(c->stream, v->ev_down, 0)); v->down_pending = false; }
  return 0;
}
extern "C" int ddf_vec_upload_async(ddf_vec* v, const double* host) {
  if (v->n == 0) return 0;
  ddf_ctx* c = v->ctx;
  if (!v->ev_up) DDF_CUDA(cudaEventCreateWithFlags(&v->ev_up, cudaEventDisableTiming));
  if (!v->ev_down) DDF_CUDA(cudaEventCreateWithFlags(&v->ev_down, cudaEventDisableTiming));
  // kernels already queued on the compute stream may still read v
  DDF_CUDA(cudaEventRecord(c->ev_a, c->stream));
  DDF_CUDA(cudaStreamWaitEvent(c->h2d_stream, c->ev_a, 0));
  if (v->down_pending) DDF_CUDA(cudaStreamWaitEvent(c->h2d_stream, v->ev_down, 0));
  DDF_CUDA(cudaMemcpyAsync(v->d.p, host, v->n * sizeof(double), cudaMemcpyHostToDevice, c->h2d_stream));
  DDF_CUDA(cudaEventRecord(v->ev_up, c->h2d_stream));
  v->up_pending = true;
  return 0;
}
extern "C" int ddf_vec_download_async(ddf_vec* v, double* host) {
  if (v->n == 0) return 0;
  ddf_ctx* c = v->ctx;
  if (!v->ev_up) DDF_CUDA(cudaEventCreateWithFlags(&v->ev_up, cudaEventDisableTiming));
  if (!v->ev_down) DDF_CUDA(cudaEventCreateWithFlags(&v->ev_down, cudaEventDisableTiming));
  // the producing kernel is on the compute stream
  DDF_CUDA(cudaEventRecord(c->ev_b, c->stream));
  DDF_CUDA(cudaStreamWaitEvent(c->d2h_stream, c->ev_b, 0));
  if (v->up_pending) DDF_CUDA(cudaStreamWaitEvent(c->d2h_stream, v->ev_up, 0));
  DDF_CUDA(cudaMemcpyAsync(host, v->d.p, v->n * sizeof(double), cudaMemcpyDeviceToHost, c->d2h_stream));
  DDF_CUDA(cudaEventRecord(v->ev_down, c->d2h_stream));
  v->down_pending = true;
  return 0;
}

__global__ void k_fill(double* __restrict__ p, int64_t n, double a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = a;
}
extern "C" int ddf_vec_fill(ddf_vec* v, double a) {
  if (v->n == 0) return 0;
  DDF_CHECK(ddf_vec_settle(v));
  int grid;
  DDF_CHECK(grid_for(v->n, 256, &grid));
  k_fill<<<grid, 256, 0, v->ctx->stream>>>(v->d.p, v->n, a);
  DDF_LAUNCH_CHECK(v->ctx);
  return 0;
}
extern "C" int ddf_vec_copy(ddf_vec* dst, ddf_vec* src) {
  if (dst->n != src->n) DDF_FAIL("vec_copy: length mismatch %lld vs %lld", (long long)dst->n, (long long)src->n);
  if (dst->n == 0) return 0;
  DDF_CHECK(ddf_vec_settle(dst));
  DDF_CHECK(ddf_vec_settle(src));
  DDF_CUDA(cudaMemcpyAsync(dst->d.p, src->d.p, dst->n * sizeof(double), cudaMemcpyDeviceToDevice, dst->ctx->stream));
  return 0;
}

// z = a x + b y, two doubles per thread (128-bit accesses)
__global__ void k_axpby(double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                        double* __restrict__ z, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    double2 xv = *reinterpret_cast<const double2*>(x + i);
    double2 r;
    if (y) {
      double2 yv = *reinterpret_cast<const double2*>(y + i);
      // separate roundings like Julia's broadcast a*x + b*y (no contraction)
      r.x = __dadd_rn(__dmul_rn(a, xv.x), __dmul_rn(b, yv.x));
      r.y = __dadd_rn(__dmul_rn(a, xv.y), __dmul_rn(b, yv.y));
    } else {
      r.x = __dmul_rn(a, xv.x);
      r.y = __dmul_rn(a, xv.y);
    }
    *reinterpret_cast<double2*>(z + i) = r;
  } else if (i < n) {
    z[i] = y ? __dadd_rn(__dmul_rn(a, x[i]), __dmul_rn(b, y[i])) : __dmul_rn(a, x[i]);
  }
}
extern "C" int ddf_vec_axpby(double a, ddf_vec* x, double b, ddf_vec* y, ddf_vec* z) {
  if (x->n != z->n || (y && y->n != z->n)) DDF_FAIL("vec_axpby: length mismatch");
  if (z->n == 0) return 0;
  DDF_CHECK(ddf_vec_settle(x));
  if (y) DDF_CHECK(ddf_vec_settle(y));
  DDF_CHECK(ddf_vec_settle(z));
  int grid;
  DDF_CHECK(grid_for(ceil_div64(z->n, 2), 256, &grid));
  k_axpby<<<grid, 256, 0, z->ctx->stream>>>(a, x->d.p, b, y ? y->d.p : nullptr, z->d.p, z->n);
  DDF_LAUNCH_CHECK(z->ctx);
  return 0;
}

// out = base + dt * (c_0 k_0 + c_1 k_1 + ...): the stage update of an explicit Runge-Kutta
// step (OrdinaryDiffEq's `@.. tmp = uprev + dt*(a31*k1 + a32*k2)`), every product and sum
// separately rounded, left to right.  One pass: N+1 reads, one write per element.
struct LincombArgs {
  const double* k[8];
  double c[8];
};
template <int N>
__global__ void __launch_bounds__(256) k_lincomb(const double* base, double dt, LincombArgs a, double* out,
                                                 int64_t n) {   // out may alias base (same index, same thread)
  // four elements per thread as two 128-bit accesses per array, read-once loads, like k_apply_diag
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    double s[4];
    {
      const double2 k0 = ld_stream_f64x2(reinterpret_cast<const double2*>(a.k[0] + i));
      const double2 k1 = ld_stream_f64x2(reinterpret_cast<const double2*>(a.k[0] + i + 2));
      s[0] = __dmul_rn(a.c[0], k0.x); s[1] = __dmul_rn(a.c[0], k0.y);
      s[2] = __dmul_rn(a.c[0], k1.x); s[3] = __dmul_rn(a.c[0], k1.y);
    }
#pragma unroll
    for (int j = 1; j < N; j++) {
      const double2 k0 = ld_stream_f64x2(reinterpret_cast<const double2*>(a.k[j] + i));
      const double2 k1 = ld_stream_f64x2(reinterpret_cast<const double2*>(a.k[j] + i + 2));
      s[0] = __dadd_rn(s[0], __dmul_rn(a.c[j], k0.x)); s[1] = __dadd_rn(s[1], __dmul_rn(a.c[j], k0.y));
      s[2] = __dadd_rn(s[2], __dmul_rn(a.c[j], k1.x)); s[3] = __dadd_rn(s[3], __dmul_rn(a.c[j], k1.y));
    }
    // base is read with a plain load: it may be `out` itself
    const double2 b0 = *reinterpret_cast<const double2*>(base + i);
    const double2 b1 = *reinterpret_cast<const double2*>(base + i + 2);
    st_stream_f64x2(reinterpret_cast<double2*>(out + i), __dadd_rn(b0.x, __dmul_rn(dt, s[0])), __dadd_rn(b0.y, __dmul_rn(dt, s[1])));
    st_stream_f64x2(reinterpret_cast<double2*>(out + i + 2), __dadd_rn(b1.x, __dmul_rn(dt, s[2])), __dadd_rn(b1.y, __dmul_rn(dt, s[3])));
  } else {
    for (int64_t q = i; q < n; q++) {
      double s = __dmul_rn(a.c[0], a.k[0][q]);
#pragma unroll
      for (int j = 1; j < N; j++) s = __dadd_rn(s, __dmul_rn(a.c[j], a.k[j][q]));
      out[q] = __dadd_rn(base[q], __dmul_rn(dt, s));
    }
  }
}

int ddf_lincomb_launch(ddf_ctx* ctx, const double* base, double dt, int nk, const double* coef, const double* const* ks,
                       double* out, int64_t n) {
  if (nk < 1 || nk > 8) DDF_FAIL("lincomb: 1..8 terms supported, got %d", nk);
  if (n == 0) return 0;
  LincombArgs a{};
  for (int j = 0; j < nk; j++) { a.k[j] = ks[j]; a.c[j] = coef[j]; }
  int grid;
  DDF_CHECK(grid_for(ceil_div64(n, 4), 256, &grid));
  switch (nk) {
    case 1: k_lincomb<1><<<grid, 256, 0, ctx->stream>>>(base, dt, a, out, n); break;
    case 2: k_lincomb<2><<<grid, 256, 0, ctx->stream>>>(base, dt, a, out, n); break;
    case 3: k_lincomb<3><<<grid, 256, 0, ctx->stream>>>(base, dt, a, out, n); break;
    case 4: k_lincomb<4><<<grid, 256, 0, ctx->stream>>>(base, dt, a, out, n); break;
    case 5: k_lincomb<5><<<grid, 256, 0, ctx->stream>>>(base, dt, a, out, n); break;
    case 6: k_lincomb<6><<<grid, 256, 0, ctx->stream>>>(base, dt, a, out, n); break;
    case 7: k_lincomb<7><<<grid, 256, 0, ctx->stream>>>(base, dt, a, out, n); break;
    default: k_lincomb<8><<<grid, 256, 0, ctx->stream>>>(base, dt, a, out, n); break;
  }
  DDF_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" int ddf_vec_lincomb(ddf_vec* out, ddf_vec* base, double dt, int nk, const double* coef, ddf_vec** ks) {
  DDF_TRY_BEGIN
  if (nk < 1 || nk > 8) DDF_FAIL("vec_lincomb: 1..8 terms supported, got %d", nk);
  if (base->n != out->n) DDF_FAIL("vec_lincomb: length mismatch");
  const double* kp[8];
  DDF_CHECK(ddf_vec_settle(out));
  DDF_CHECK(ddf_vec_settle(base));
  for (int j = 0; j < nk; j++) {
    if (ks[j]->n != out->n) DDF_FAIL("vec_lincomb: length mismatch in term %d", j);
    if (ks[j] == out) DDF_FAIL("vec_lincomb: out must not alias a term");
    DDF_CHECK(ddf_vec_settle(ks[j]));
    kp[j] = ks[j]->d.p;
  }
  return ddf_lincomb_launch(out->ctx, base->d.p, dt, nk, coef, kp, out->d.p, out->n);
  DDF_TRY_END
}

__global__ void k_mul(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ z, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = x[i] * y[i];
}
extern "C" int ddf_vec_mul(ddf_vec* x, ddf_vec* y, ddf_vec* z) {
  if (x->n != z->n || y->n != z->n) DDF_FAIL("vec_mul: length mismatch");
  if (z->n == 0) return 0;
  DDF_CHECK(ddf_vec_settle(x));
  DDF_CHECK(ddf_vec_settle(y));
  DDF_CHECK(ddf_vec_settle(z));
  int grid;
  DDF_CHECK(grid_for(z->n, 256, &grid));
  k_mul<<<grid, 256, 0, z->ctx->stream>>>(x->d.p, y->d.p, z->d.p, z->n);
  DDF_LAUNCH_CHECK(z->ctx);
  return 0;
}

__global__ void k_div(const double* __restrict__ x, double a, double* __restrict__ z, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = __ddiv_rn(x[i], a);
}
extern "C" int ddf_vec_div(ddf_vec* x, double a, ddf_vec* z) {
  if (x->n != z->n) DDF_FAIL("vec_div: length mismatch");
  if (z->n == 0) return 0;
  DDF_CHECK(ddf_vec_settle(x));
  DDF_CHECK(ddf_vec_settle(z));
  int grid;
  DDF_CHECK(grid_for(z->n, 256, &grid));
  k_div<<<grid, 256, 0, z->ctx->stream>>>(x->d.p, a, z->d.p, z->n);
  DDF_LAUNCH_CHECK(z->ctx);
  return 0;
}

// counter-based uniform [0,1): splitmix64 of (seed, i) -> 53-bit mantissa
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void k_rand(double* __restrict__ p, int64_t n, uint64_t seed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + (uint64_t)i);
    p[i] = (double)(r >> 11) * (1.0 / 9007199254740992.0);
  }
}
extern "C" int ddf_vec_rand(ddf_vec* v, uint64_t seed) {
  if (v->n == 0) return 0;
  DDF_CHECK(ddf_vec_settle(v));
  int grid;
  DDF_CHECK(grid_for(v->n, 256, &grid));
  k_rand<<<grid, 256, 0, v->ctx->stream>>>(v->d.p, v->n, seed);
  DDF_LAUNCH_CHECK(v->ctx);
  return 0;
}

// ------------------------------------------------------------------ reductions
// Two-stage deterministic reduction: fixed grid, each block reduces a contiguous
// chunk in a fixed order, the (<= 1024) partials are combined on the host in
// index order.  mode: 0 sum(x), 1 sum|x|, 2 sum x^2, 3 max|x|, 4 sum x*y, 5 min(x),
// 6 sum (x * (1/w)) * y  (the weighted inner product of ManifoldOps.jl:154-164: dot(f, Diagonal(1 ./ w), g))
template <int MODE>
__global__ void k_reduce(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                         double* __restrict__ partial, const double* __restrict__ wgt = nullptr) {
  __shared__ double sm[32];
  double acc = (MODE == 5) ? INFINITY : 0.0;
  int64_t per = (n + gridDim.x - 1) / gridDim.x;
  int64_t lo = (int64_t)blockIdx.x * per;
  int64_t hi = lo + per < n ? lo + per : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    double v = x[i];
    if (MODE == 0) acc += v;
    else if (MODE == 1) acc += fabs(v);
    else if (MODE == 2) acc += v * v;
    else if (MODE == 3) acc = fmax(acc, fabs(v));
    else if (MODE == 4) acc += v * y[i];
    else if (MODE == 6) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(v, __ddiv_rn(1.0, wgt[i])), y[i]));
    else acc = fmin(acc, v);
  }
  for (int off = 16; off > 0; off >>= 1) {
    double o = __shfl_down_sync(0xffffffffu, acc, off);
    if (MODE == 3) acc = fmax(acc, o);
    else if (MODE == 5) acc = fmin(acc, o);
    else acc += o;
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sm[w] = acc;
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    acc = lane < nw ? sm[lane] : ((MODE == 5) ? INFINITY : 0.0);
    for (int off = 16; off > 0; off >>= 1) {
      double o = __shfl_down_sync(0xffffffffu, acc, off);
      if (MODE == 3) acc = fmax(acc, o);
      else if (MODE == 5) acc = fmin(acc, o);
      else acc += o;
    }
    if (lane == 0) partial[blockIdx.x] = acc;
  }
}

static int reduce_impl(ddf_ctx* c, int mode, const double* x, const double* y, int64_t n, double* out, const double* w) {
  if (n == 0) { *out = (mode == 5) ? INFINITY : 0.0; return 0; }
  int grid = (int)(ceil_div64(n, 4096) < 1024 ? ceil_div64(n, 4096) : 1024);
  switch (mode) {
    case 0: k_reduce<0><<<grid, 256, 0, c->stream>>>(x, y, n, c->red_dev); break;
    case 1: k_reduce<1><<<grid, 256, 0, c->stream>>>(x, y, n, c->red_dev); break;
    case 2: k_reduce<2><<<grid, 256, 0, c->stream>>>(x, y, n, c->red_dev); break;
    case 3: k_reduce<3><<<grid, 256, 0, c->stream>>>(x, y, n, c->red_dev); break;
    case 4: k_reduce<4><<<grid, 256, 0, c->stream>>>(x, y, n, c->red_dev); break;
    case 6: k_reduce<6><<<grid, 256, 0, c->stream>>>(x, y, n, c->red_dev, w); break;
    default: k_reduce<5><<<grid, 256, 0, c->stream>>>(x, y, n, c->red_dev); break;
  }
  DDF_LAUNCH_CHECK(c);
  DDF_CUDA(cudaMemcpyAsync(c->red_host, c->red_dev, grid * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  DDF_CUDA(cudaStreamSynchronize(c->stream));
  double acc = (mode == 5) ? INFINITY : 0.0;
  for (int i = 0; i < grid; i++) {
    if (mode == 3) acc = acc > c->red_host[i] ? acc : c->red_host[i];
    else if (mode == 5) acc = acc < c->red_host[i] ? acc : c->red_host[i];
    else acc += c->red_host[i];
  }
  *out = acc;
  return 0;
}

int ddf_reduce(ddf_ctx* c, int mode, const double* x, const double* y, int64_t n, double* out) {
  return reduce_impl(c, mode, x, y, n, out, nullptr);
}

// All inner products <v_r, v_c>, r <= c, of up to 8 vectors in one pass and one host synchronisation (the
// Gram matrix of BiCGStab(l)'s minimal-residual step, solve.cu).  Each block owns a contiguous range and
// writes its npairs partial sums; the host adds the blocks in index order, so the result is deterministic.
struct GramPtrs {
  const double* p[8];
};
template <int NV>
__global__ void k_gram(GramPtrs vs, int64_t n, double* __restrict__ partial) {
  constexpr int NP = NV * (NV + 1) / 2;
  __shared__ double sm[8][NP];
  double acc[NP];
#pragma unroll
  for (int k = 0; k < NP; k++) acc[k] = 0.0;
  int64_t per = (n + gridDim.x - 1) / gridDim.x;
  int64_t lo = (int64_t)blockIdx.x * per;
  int64_t hi = lo + per < n ? lo + per : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    double x[NV];
#pragma unroll
    for (int r = 0; r < NV; r++) x[r] = vs.p[r][i];
    int k = 0;
#pragma unroll
    for (int r = 0; r < NV; r++)
#pragma unroll
      for (int c = r; c < NV; c++) acc[k++] += x[r] * x[c];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NP; k++) {
    double a = acc[k];
    for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
    if (lane == 0) sm[w][k] = a;
  }
  __syncthreads();
  if (threadIdx.x < NP) {
    double a = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); q++) a += sm[q][threadIdx.x];
    partial[(int64_t)blockIdx.x * NP + threadIdx.x] = a;
  }
}

// out[r * nv + c] = <v_r, v_c> (symmetric, both triangles filled)
int ddf_vec_gram(ddf_ctx* c, int nv, ddf_vec* const* vs, double* out) {
  if (nv < 1 || nv > 8) DDF_FAIL("vec_gram: 1 <= nv <= 8, got %d", nv);
  GramPtrs g{};
  const int64_t n = vs[0]->n;
  for (int r = 0; r < nv; r++) {
    if (vs[r]->n != n) DDF_FAIL("vec_gram: length mismatch");
    DDF_CHECK(ddf_vec_settle(vs[r]));
    g.p[r] = vs[r]->d.p;
  }
  const int np = nv * (nv + 1) / 2;
  int grid = (int)(ceil_div64(n, 2048) < 4096 / np ? ceil_div64(n, 2048) : 4096 / np);
  if (grid < 1) grid = 1;
  switch (nv) {
    case 1: k_gram<1><<<grid, 256, 0, c->stream>>>(g, n, c->red_dev); break;
    case 2: k_gram<2><<<grid, 256, 0, c->stream>>>(g, n, c->red_dev); break;
    case 3: k_gram<3><<<grid, 256, 0, c->stream>>>(g, n, c->red_dev); break;
    case 4: k_gram<4><<<grid, 256, 0, c->stream>>>(g, n, c->red_dev); break;
    case 5: k_gram<5><<<grid, 256, 0, c->stream>>>(g, n, c->red_dev); break;
    case 6: k_gram<6><<<grid, 256, 0, c->stream>>>(g, n, c->red_dev); break;
    case 7: k_gram<7><<<grid, 256, 0, c->stream>>>(g, n, c->red_dev); break;
    default: k_gram<8><<<grid, 256, 0, c->stream>>>(g, n, c->red_dev); break;
  }
  DDF_LAUNCH_CHECK(c);
  DDF_CUDA(cudaMemcpyAsync(c->red_host, c->red_dev, (size_t)grid * np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  DDF_CUDA(cudaStreamSynchronize(c->stream));
  int k = 0;
  for (int r = 0; r < nv; r++)
    for (int cc = r; cc < nv; cc++, k++) {
      double a = 0.0;
      for (int b = 0; b < grid; b++) a += c->red_host[(size_t)b * np + k];
      out[r * nv + cc] = a;
      out[cc * nv + r] = a;
    }
  return 0;
}

extern "C" int ddf_vec_norm(ddf_vec* v, int kind, double* out) {
  double r = 0;
  DDF_CHECK(ddf_vec_settle(v));
  if (kind == DDF_NORM_INF) { DDF_CHECK(ddf_reduce(v->ctx, 3, v->d.p, nullptr, v->n, &r)); }
  else if (kind == DDF_NORM_1) { DDF_CHECK(ddf_reduce(v->ctx, 1, v->d.p, nullptr, v->n, &r)); }
  else if (kind == DDF_NORM_2) { DDF_CHECK(ddf_reduce(v->ctx, 2, v->d.p, nullptr, v->n, &r)); r = sqrt(r); }
  else DDF_FAIL("vec_norm: unknown kind %d", kind);
  *out = r;
  return 0;
}
extern "C" int ddf_vec_dot(ddf_vec* x, ddf_vec* y, double* out) {
  if (x->n != y->n) DDF_FAIL("vec_dot: length mismatch");
  DDF_CHECK(ddf_vec_settle(x));
  DDF_CHECK(ddf_vec_settle(y));
  return ddf_reduce(x->ctx, 4, x->d.p, y->d.p, x->n, out);
}
// dot(f, g) of the reference (ManifoldOps.jl:154-164): defined for Fun{D,Pr,D} (weights 1 ./ vol_D) and
// Fun{D,Dl,0} (weights 1 ./ dualvol_0) only; one pass, the 1/w division fused into the read of the weight.
extern "C" int ddf_vec_wdot(ddf_mesh* m, int P, int R, ddf_vec* f, ddf_vec* g, double* out) {
  DDF_TRY_BEGIN
  const bool prD = (P == DDF_PR && R == m->D), dl0 = (P == DDF_DL && R == 0);
  if (!prD && !dl0) DDF_FAIL("dot: defined for Fun{D,Pr,D} and Fun{D,Dl,0} only (ManifoldOps.jl:154-164), got P=%d R=%d", P, R);
  DDF_CHECK(ddf_mesh_require_geometry(m));
  if (f->n != m->n[R] || g->n != m->n[R]) DDF_FAIL("dot: cochain length %lld / %lld, mesh has %lld %d-simplices", (long long)f->n, (long long)g->n, (long long)m->n[R], R);
  DDF_CHECK(ddf_vec_settle(f));
  DDF_CHECK(ddf_vec_settle(g));
  const double* w = prD ? m->vol[R].p : m->dualvol[0].p;
  return reduce_impl(m->ctx, 6, f->d.p, g->d.p, f->n, out, w);
  DDF_TRY_END
}
extern "C" int ddf_vec_sum(ddf_vec* x, double* out) {
  DDF_CHECK(ddf_vec_settle(x));
  return ddf_reduce(x->ctx, 0, x->d.p, nullptr, x->n, out);
}

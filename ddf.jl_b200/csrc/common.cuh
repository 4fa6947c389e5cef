// common.cuh -- shared host-side plumbing of libddf_b200 (handles, error state,
// device allocation accounting, launch counting).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <new>
#include <string>
#include <vector>

#include "../../include/ddf_b200.h"

#define DDF_MAXD 5            // topology (assembly, d_k) up to D = 5 like the reference tests
#define DDF_MAXC 3            // geometry kernels are written for C <= 3 (BASELINE configs are 2D/3D)

// chop threshold of the Op constructor: abs2(f) < sqrt(eps(Float64)^3)  (src/Ops.jl:52,57)
#define DDF_CHOP_ABS2 3.308722450212111e-24

void ddf_set_error(const char* fmt, ...);

#define DDF_FAIL(...)              \
  do {                             \
    ddf_set_error(__VA_ARGS__);    \
    return 1;                      \
  } while (0)

#define DDF_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) {                                                  \
      ddf_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,        \
                    cudaGetErrorString(_e));                                  \
      return 1;                                                               \
    }                                                                         \
  } while (0)

#define DDF_CHECK(expr)            \
  do {                             \
    int _r = (expr);               \
    if (_r != 0) return _r;        \
  } while (0)

#define DDF_TRY_BEGIN try {
#define DDF_TRY_END                                        \
  }                                                        \
  catch (const std::exception& e) {                        \
    ddf_set_error("exception: %s", e.what());              \
    return 1;                                              \
  }                                                        \
  catch (...) {                                            \
    ddf_set_error("unknown exception");                    \
    return 1;                                              \
  }

// kernel-variant selectors (ddf_set_tuning / ddf_ctx_set_tuning): per context, so that distinct contexts can be
// driven from distinct threads; every variant returns bit-identical results
struct DdfTune {
  int fixed = 0;       // storage / launch variant of the d_k and boundary applies (ops.cu)
  int wave = 0;        // row-kernel variant of the fused steps (wave.cu), bit 1024: NCCL halo (comm.cu)
  int asm_path = 0;    // 1: radix-sort assembly instead of bucket ranking (mesh.cu)
  int packbits = 24;   // offset bits of the prefix-packed rows (tests narrow it to reach the explicit tables)
  int prefetch = 592;  // blocks ahead a block asks L2 for the index words of a later block (0 = off)
};

struct ddf_ctx {
  int device = -1;
  DdfTune tune;
  void* asm_work = nullptr;          // workspace of the ddf_mesh_assemble in flight on this ctx (mesh.cu)
  int sm_count = 0;
  cudaStream_t stream = nullptr;     // compute stream
  cudaStream_t comm_stream = nullptr;  // halo / NCCL stream
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // asynchronous host copies (one per PCIe direction)
  cudaEvent_t ev_tic = nullptr, ev_toc = nullptr, ev_a = nullptr, ev_b = nullptr;
  std::vector<cudaEvent_t> events;   // numbered timing events (ddf_event_record)
  int64_t bytes_in_use = 0;
  int64_t launches = 0;
  // scratch for reductions
  double* red_dev = nullptr;         // small device buffer
  double* red_host = nullptr;        // pinned
  // L2 flush buffer
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // NCCL (dlopen'ed lazily)
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
};

int ddf_dev_alloc(ddf_ctx* ctx, void** p, size_t bytes, bool plain = false);
void ddf_dev_free(ddf_ctx* ctx, void* p, size_t bytes, bool plain = false);

// RAII-ish device buffer tied to a ctx (not thread safe; one ctx per thread)
template <typename T>
struct DevBuf {
  ddf_ctx* ctx = nullptr;
  T* p = nullptr;
  size_t n = 0;
  bool plain = false;        // cudaMalloc instead of the stream-ordered pool (memory exported through CUDA IPC)
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  int alloc(ddf_ctx* c, size_t count) {
    release();
    ctx = c;
    n = count;
    if (count == 0) { p = nullptr; return 0; }
    return ddf_dev_alloc(c, (void**)&p, count * sizeof(T), plain);
  }
  void release() {
    if (p) ddf_dev_free(ctx, p, n * sizeof(T), plain);
    p = nullptr;
    n = 0;
  }
  void swap(DevBuf& o) {
    std::swap(ctx, o.ctx); std::swap(p, o.p); std::swap(n, o.n); std::swap(plain, o.plain);
  }
  size_t bytes() const { return n * sizeof(T); }
};

struct ddf_vec {
  ddf_ctx* ctx = nullptr;
  DevBuf<double> d;
  int64_t n = 0;
  // pending asynchronous copies (ddf_vec_upload_async / ddf_vec_download_async)
  cudaEvent_t ev_up = nullptr, ev_down = nullptr;
  bool up_pending = false, down_pending = false;
};
// make the compute stream wait for v's pending copies (call before any kernel touches v)
int ddf_vec_settle(ddf_vec* v);

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid helper: 1-D launch with 64-bit element counts
static inline int grid_for(int64_t n, int per_block, int* grid) {
  int64_t g = ceil_div64(n, per_block);
  if (g > 0x7fffffffLL) { ddf_set_error("grid too large"); return 1; }
  *grid = (int)(g > 0 ? g : 1);
  return 0;
}

#define DDF_LAUNCH_CHECK(ctx)                                                   \
  do {                                                                          \
    (ctx)->launches++;                                                          \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) {                                                    \
      ddf_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,      \
                    cudaGetErrorString(_e));                                    \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

// ---------------------------------------------------------------- PTX helpers
// streaming (read-once) 128-bit index load: bypass L1 allocation so the L1 stays
// free for the x-gather.
__device__ __forceinline__ int4 ld_stream_int4(const int4* p) {
  int4 r;
  asm("ld.global.nc.L1::no_allocate.L2::128B.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int2 ld_stream_int2(const int2* p) {
  int2 r;
  asm("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];"
               : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_stream_int(const int* p) {
  int r;
  asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double r;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_stream_f64x2(const double2* p) {
  double2 r;
  asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
// L2 eviction-priority policies (createpolicy, sm_80+): the gathered cochain x is
// re-used by neighbouring rows -> evict_last; index tables and outputs are touched
// once -> evict_first, so they do not push x out of the 126 MB L2.
__device__ __forceinline__ uint64_t make_policy_evict_last() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t make_policy_evict_first() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double ld_keep_f64(const double* p, uint64_t pol) {
  double r;
  asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ int ld_keep_s32(const int* p, uint64_t pol) {
  int r;
  asm("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ int4 ld_stream_int4_hint(const int4* p, uint64_t pol) {
  int4 r;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ int4 ld_l1_int4_hint(const int4* p, uint64_t pol) {
  int4 r;
  asm("ld.global.nc.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ double2 ld_stream_f64x2_hint(const double2* p, uint64_t pol) {
  double2 r;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
               : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_hint_f64x2(double2* p, double a, double b, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" :: "l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint_f64(double* p, double a, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" :: "l"(p), "d"(a), "l"(pol) : "memory");
}
// L2 prefetch of a contiguous piece of global memory (address and size multiples of 16 bytes): a block asks for the
// index words a LATER block will read, so that block's first load hits L2 instead of waiting on HBM.
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
// streaming (write-once) stores: evict-first so the outputs do not push the
// gathered x out of L2.
__device__ __forceinline__ void st_stream_f64x2(double2* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_stream_f64(double* p, double a) {
  asm volatile("st.global.cs.f64 [%0], %1;" :: "l"(p), "d"(a) : "memory");
}

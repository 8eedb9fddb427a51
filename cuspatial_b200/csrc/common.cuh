// common.cuh -- shared host/device utilities of the B200 spatial-join library.
// Hand-written sm_100a CUDA; no Thrust/CUB anywhere in this directory.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <exception>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cuspatial_b200.h"

namespace bsj {

using u8  = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i32 = int32_t;
using i64 = int64_t;

// SM count of the current device (148 on a B200: 2 dies x 74 SMs), queried once per device.
int num_sms();

// One-time per-DEVICE kernel configuration (cudaFuncSetAttribute applies to the current device
// only): runs `configure` the first time it is called with `slot` on each device, thread-safe.
// Slots: 0 radix sort, 1 bbox join, 2 partition, 3 pip.
void configure_once_per_device(int slot, void (*configure)());

// ---------------------------------------------------------------------------------------------
// Errors.  logic_error wording follows the reference (cpp/include/cuspatial/error.hpp:76-79):
//   "cuSpatial failure at: <file>:<line>: <reason>"
// ---------------------------------------------------------------------------------------------
struct error : public std::exception {
  int code;
  std::string msg;
  error(int c, std::string m) : code(c), msg(std::move(m)) {}
  const char* what() const noexcept override { return msg.c_str(); }
};

#define BSJ_STR2(x) #x
#define BSJ_STR(x) BSJ_STR2(x)
#define BSJ_EXPECTS(cond, reason)                                                       \
  do {                                                                                  \
    if (!(cond))                                                                        \
      throw ::bsj::error(BSJ_INVALID_ARGUMENT,                                          \
                         std::string("cuSpatial failure at: " __FILE__                  \
                                     ":" BSJ_STR(__LINE__) ": ") + (reason));           \
  } while (0)

#define BSJ_CUDA_TRY(call)                                                              \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      cudaGetLastError();                                                               \
      throw ::bsj::error(e_ == cudaErrorMemoryAllocation ? BSJ_OUT_OF_MEMORY            \
                                                         : BSJ_CUDA_ERROR,              \
                         std::string("CUDA error at: " __FILE__ ":" BSJ_STR(__LINE__)   \
                                     ": ") + cudaGetErrorName(e_) + " " +               \
                           cudaGetErrorString(e_));                                     \
    }                                                                                   \
  } while (0)

extern std::atomic<u64> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }
#define BSJ_CHECK_LAUNCH()                 \
  do {                                     \
    ::bsj::count_launch();                 \
    BSJ_CUDA_TRY(cudaGetLastError());      \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Stream-ordered temporaries from the device's default mempool (cudaMallocAsync).  The pool's
// release threshold is raised once so that memory stays cached between calls: after the first
// call an allocation is a few hundred ns of host time and no device synchronisation.
// ---------------------------------------------------------------------------------------------
void ensure_pool_configured();

template <typename T>
class dev_buf {
 public:
  dev_buf() = default;
  dev_buf(size_t n, cudaStream_t s) { alloc(n, s); }
  dev_buf(dev_buf&& o) noexcept : p_(o.p_), n_(o.n_), s_(o.s_) { o.p_ = nullptr; o.n_ = 0; }
  dev_buf& operator=(dev_buf&& o) noexcept
  {
    if (this != &o) {
      release();
      p_ = o.p_; n_ = o.n_; s_ = o.s_;
      o.p_ = nullptr; o.n_ = 0;
    }
    return *this;
  }
  dev_buf(dev_buf const&)            = delete;
  dev_buf& operator=(dev_buf const&) = delete;
  ~dev_buf() { release(); }

  void alloc(size_t n, cudaStream_t s)
  {
    release();
    ensure_pool_configured();
    s_ = s;
    n_ = n;
    if (n) BSJ_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&p_), n * sizeof(T), s));
  }
  void release()
  {
    if (p_) cudaFreeAsync(p_, s_);
    p_ = nullptr;
    n_ = 0;
  }
  T* get() const { return p_; }
  size_t size() const { return n_; }
  T* detach()
  {
    T* p = p_;
    p_   = nullptr;
    n_   = 0;
    return p;
  }

 private:
  T* p_{nullptr};
  size_t n_{0};
  cudaStream_t s_{nullptr};
};

// Output columns: allocated through the caller's bsj_allocator (the reference's `mr`) or, when it
// is NULL, with cudaMallocAsync (released by bsj_free).
struct out_alloc {
  const bsj_allocator* mr;
  cudaStream_t stream;
  std::vector<std::pair<void*, size_t>> live;  // for unwinding on error

  out_alloc(const bsj_allocator* m, cudaStream_t s) : mr(m), stream(s) {}
  void* raw(size_t bytes)
  {
    if (bytes == 0) return nullptr;
    void* p = nullptr;
    if (mr && mr->allocate) {
      p = mr->allocate(bytes, (bsj_stream_t)stream, mr->ctx);
      if (!p) throw error(BSJ_OUT_OF_MEMORY, "output allocator returned NULL");
    } else {
      ensure_pool_configured();
      BSJ_CUDA_TRY(cudaMallocAsync(&p, bytes, stream));
    }
    live.emplace_back(p, bytes);
    return p;
  }
  template <typename T>
  T* get(size_t n)
  {
    return static_cast<T*>(raw(n * sizeof(T)));
  }
  void commit() { live.clear(); }
  ~out_alloc()
  {
    for (auto& pr : live) {
      if (mr && mr->deallocate)
        mr->deallocate(pr.first, pr.second, (bsj_stream_t)stream, mr->ctx);
      else if (!(mr && mr->allocate))
        cudaFreeAsync(pr.first, stream);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Optional per-stage profiling (CUDA events on the call's stream).
// ---------------------------------------------------------------------------------------------
struct stage_timer {
  cudaStream_t s;
  bool on;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  explicit stage_timer(cudaStream_t stream);
  void mark(const char* name);
  void finish();  // stream must be synchronised; publishes into the thread-local profile
  ~stage_timer();
};

// mark on the innermost live stage_timer of this thread (no-op when profiling is off)
void prof_mark(const char* name);

inline int div_up(u64 a, u64 b) { return (int)((a + b - 1) / b); }

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt()
{
  u32 m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ void st_relaxed_u64(u64* p, u64 v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p)
{
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// streaming (evict-first) loads/stores for data touched exactly once
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p)
{
  return __ldcs(p);
}
template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v)
{
  __stcs(p, v);
}

__device__ __forceinline__ u32 warp_inclusive_scan(u32 v)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= (u32)o) v += t;
  }
  return v;
}

// Decoupled look-back descriptor: [63:32] tag (0 = not ready), [31:0] value.
// Tags are unique per (use, pass) so one zero-initialised array serves several passes.
__device__ __forceinline__ u64 lb_pack(u32 tag, u32 v) { return ((u64)tag << 32) | v; }

// Single-value chained scan step for tile `tile` (called by ONE thread): publish `aggregate`,
// return the exclusive prefix over tiles [0, tile).
__device__ __forceinline__ u32 lookback_exclusive(u64* desc, u32 tile, u32 aggregate, u32 tag_agg,
                                                  u32 tag_pre)
{
  if (tile == 0) {
    st_relaxed_u64(desc, lb_pack(tag_pre, aggregate));
    return 0;
  }
  st_relaxed_u64(desc + tile, lb_pack(tag_agg, aggregate));
  u32 excl = 0;
  i64 t    = (i64)tile - 1;
  while (true) {
    u64 v    = ld_relaxed_u64(desc + t);
    u32 flag = (u32)(v >> 32);
    if (flag == tag_pre) {
      excl += (u32)v;
      break;
    }
    if (flag == tag_agg) {
      excl += (u32)v;
      --t;
      continue;
    }
    __nanosleep(20);
  }
  st_relaxed_u64(desc + tile, lb_pack(tag_pre, excl + aggregate));
  return excl;
}
#endif  // __CUDACC__

}  // namespace bsj

// Minimal stand-in for the part of RMM the reference's header-only hot path uses.
// TEST INFRASTRUCTURE ONLY (used to compile the reference's own headers from
// /root/reference into oracle/_ref/); never part of the product path.
// Two flavours: host (g++ + Thrust OpenMP backend) and CUDA (nvcc).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#if defined(__CUDACC__)
#include <cuda_runtime_api.h>
#endif

namespace rmm {
#if defined(__CUDACC__)
using shim_stream_t = cudaStream_t;
#else
using shim_stream_t = void*;
#endif
class cuda_stream_view {
 public:
  constexpr cuda_stream_view() = default;
  constexpr cuda_stream_view(shim_stream_t s) : s_(s) {}
  constexpr shim_stream_t value() const { return s_; }
  constexpr operator shim_stream_t() const { return s_; }
  void synchronize() const
  {
#if defined(__CUDACC__)
    cudaStreamSynchronize(s_);
#endif
  }

 private:
  shim_stream_t s_{};
};
static constexpr cuda_stream_view cuda_stream_default{};

struct out_of_memory : public std::bad_alloc {
  const char* what() const noexcept override { return "shim out_of_memory"; }
};
}  // namespace rmm

// Host-flavour only (force-included by oracle/Makefile's ref_host rule): the few CUDA global
// functions the reference's header-only path calls unqualified from "device" code.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>

// CUDA's global min/max overloads for floating point are fminf/fmin, fmaxf/fmax.
inline float min(float a, float b) { return std::fmin(a, b); }
inline double min(double a, double b) { return std::fmin(a, b); }
inline float max(float a, float b) { return std::fmax(a, b); }
inline double max(double a, double b) { return std::fmax(a, b); }

// detail/utility/zero_data.cuh calls cudaMemsetAsync(dst, 0, bytes, stream.value()); the host
// flavour's stream handle is a void* and its "device" memory is host memory.
inline int shim_memset_async(void* dst, int value, std::size_t bytes, void*)
{
  std::memset(dst, value, bytes);
  return 0;
}
#define cudaMemsetAsync shim_memset_async

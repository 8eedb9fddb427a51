#pragma once
#include <rmm/cuda_stream_view.hpp>
#include <thrust/execution_policy.h>
#if defined(__CUDACC__)
#include <thrust/system/cuda/execution_policy.h>
#else
#include <thrust/system/omp/execution_policy.h>
#endif
namespace rmm {
#if defined(__CUDACC__)
inline auto exec_policy(cuda_stream_view s = {}) { return thrust::cuda::par_nosync.on(s.value()); }
inline auto exec_policy_nosync(cuda_stream_view s = {}) { return thrust::cuda::par_nosync.on(s.value()); }
#else
inline auto exec_policy(cuda_stream_view = {}) { return thrust::omp::par; }
inline auto exec_policy_nosync(cuda_stream_view = {}) { return thrust::omp::par; }
#endif
}  // namespace rmm

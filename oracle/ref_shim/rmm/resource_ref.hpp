#pragma once
#include <rmm/cuda_stream_view.hpp>
namespace rmm {
// The reference only passes this through to device_uvector; one global resource.
struct device_async_resource_ref {
  void* allocate(std::size_t bytes, cuda_stream_view s) const
  {
    if (bytes == 0) return nullptr;
#if defined(__CUDACC__)
    void* p = nullptr;
    if (cudaMallocAsync(&p, bytes, s.value()) != cudaSuccess) {
      cudaGetLastError();
      throw out_of_memory{};
    }
    return p;
#else
    void* p = std::malloc(bytes);
    if (!p) throw out_of_memory{};
    return p;
#endif
  }
  void deallocate(void* p, cuda_stream_view s) const
  {
    if (!p) return;
#if defined(__CUDACC__)
    cudaFreeAsync(p, s.value());
#else
    std::free(p);
#endif
  }
};
}  // namespace rmm

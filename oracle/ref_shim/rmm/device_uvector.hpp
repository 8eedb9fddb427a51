#pragma once
#include <rmm/cuda_stream_view.hpp>
#include <rmm/mr/device/per_device_resource.hpp>
#include <rmm/resource_ref.hpp>
#include <thrust/device_ptr.h>
#include <utility>

namespace rmm {
// Uninitialised, stream-ordered, move-only typed buffer (subset of rmm::device_uvector).
template <typename T>
class device_uvector {
 public:
  using value_type      = T;
  using size_type       = std::size_t;
  using pointer         = T*;
  using const_pointer   = T const*;
  using iterator        = T*;
  using const_iterator  = T const*;
  using reference       = T&;
  using const_reference = T const&;

  device_uvector(std::size_t n, cuda_stream_view s, device_async_resource_ref mr = {})
    : mr_(mr), stream_(s)
  {
    p_ = static_cast<T*>(mr_.allocate(n * sizeof(T), s));
    n_ = cap_ = n;
  }
  device_uvector(device_uvector&& o) noexcept
    : p_(o.p_), n_(o.n_), cap_(o.cap_), mr_(o.mr_), stream_(o.stream_)
  {
    o.p_ = nullptr;
    o.n_ = o.cap_ = 0;
  }
  device_uvector& operator=(device_uvector&& o) noexcept
  {
    if (this != &o) {
      mr_.deallocate(p_, stream_);
      p_ = o.p_; n_ = o.n_; cap_ = o.cap_; mr_ = o.mr_; stream_ = o.stream_;
      o.p_ = nullptr; o.n_ = o.cap_ = 0;
    }
    return *this;
  }
  device_uvector(device_uvector const&)            = delete;
  device_uvector& operator=(device_uvector const&) = delete;
  ~device_uvector() { mr_.deallocate(p_, stream_); }

  T* data() noexcept { return p_; }
  T const* data() const noexcept { return p_; }
  T* begin() noexcept { return p_; }
  T const* begin() const noexcept { return p_; }
  T const* cbegin() const noexcept { return p_; }
  T* end() noexcept { return p_ + n_; }
  T const* end() const noexcept { return p_ + n_; }
  T const* cend() const noexcept { return p_ + n_; }
  std::size_t size() const noexcept { return n_; }
  std::int64_t ssize() const noexcept { return static_cast<std::int64_t>(n_); }
  std::size_t capacity() const noexcept { return cap_; }
  bool is_empty() const noexcept { return n_ == 0; }
  cuda_stream_view stream() const noexcept { return stream_; }

  void reserve(std::size_t new_cap, cuda_stream_view s)
  {
    if (new_cap <= cap_) return;
    realloc_to(new_cap, s);
  }
  void resize(std::size_t n, cuda_stream_view s)
  {
    if (n > cap_) realloc_to(n, s);
    n_ = n;
  }
  void shrink_to_fit(cuda_stream_view s)
  {
    if (n_ != cap_) realloc_to(n_, s);
  }
  void set_element_async(std::size_t i, T const& v, cuda_stream_view s)
  {
#if defined(__CUDACC__)
    cudaMemcpyAsync(p_ + i, &v, sizeof(T), cudaMemcpyHostToDevice, s.value());
#else
    p_[i] = v;
#endif
  }
  void set_element_to_zero_async(std::size_t i, cuda_stream_view s)
  {
#if defined(__CUDACC__)
    cudaMemsetAsync(p_ + i, 0, sizeof(T), s.value());
#else
    std::memset(p_ + i, 0, sizeof(T));
#endif
  }
  T element(std::size_t i, cuda_stream_view s) const
  {
#if defined(__CUDACC__)
    T v;
    cudaMemcpyAsync(&v, p_ + i, sizeof(T), cudaMemcpyDeviceToHost, s.value());
    cudaStreamSynchronize(s.value());
    return v;
#else
    return p_[i];
#endif
  }
  // shim-only: hand the allocation to the caller (freed with the resource's deallocate)
  T* shim_release() noexcept
  {
    T* p = p_;
    p_   = nullptr;
    n_ = cap_ = 0;
    return p;
  }
  T front_element(cuda_stream_view s) const { return element(0, s); }
  T back_element(cuda_stream_view s) const { return element(n_ - 1, s); }

 private:
  void realloc_to(std::size_t new_cap, cuda_stream_view s)
  {
    T* q             = static_cast<T*>(mr_.allocate(new_cap * sizeof(T), s));
    std::size_t keep = n_ < new_cap ? n_ : new_cap;
    if (keep) {
#if defined(__CUDACC__)
      cudaMemcpyAsync(q, p_, keep * sizeof(T), cudaMemcpyDeviceToDevice, s.value());
#else
      std::memcpy(q, p_, keep * sizeof(T));
#endif
    }
    mr_.deallocate(p_, s);
    p_   = q;
    cap_ = new_cap;
  }
  T* p_{nullptr};
  std::size_t n_{0}, cap_{0};
  device_async_resource_ref mr_{};
  cuda_stream_view stream_{};
};
}  // namespace rmm

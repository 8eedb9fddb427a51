// bulk_store_probe.cu -- how fast can one SM issue SMALL 1-D bulk stores (cp.async.bulk
// shared -> global)?  Decides whether the radix sort's per-digit runs (128 B on average) can
// leave the SM through the TMA engine instead of the LSU pipe.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/bulk_store_probe.cu -o gpurun_variants/bulk_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each CTA owns a 64 KB tile in shared memory and writes it out as (65536 / bytes) runs
template <int MODE>  // 0: bulk stores, one run per thread-iteration; 1: plain 128-bit STG
__global__ void __launch_bounds__(512, 2) probe(uint32_t* __restrict__ dst, int bytes, size_t tile_stride_words)
{
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* s = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 16384; i += 512) s[i] = i + blockIdx.x;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  int const runs = 65536 / bytes;
  uint32_t* const base = dst + (size_t)blockIdx.x * tile_stride_words;
  if (MODE == 0) {
    for (int r = threadIdx.x; r < runs; r += 512) {
      // scatter the runs: run r goes to slot (r * 37) % runs of this tile's region
      int const slot = (int)(((long long)r * 37) % runs);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                     base + (size_t)slot * (bytes / 4)),
                   "r"(smem_u32(s + (size_t)r * (bytes / 4))), "r"(bytes)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else {
    int const vec_per_run = bytes / 16;
    for (int v = threadIdx.x; v < 4096; v += 512) {
      int const r = v / vec_per_run, o = v % vec_per_run;
      int const slot = (int)(((long long)r * 37) % runs);
      reinterpret_cast<uint4*>(base + (size_t)slot * (bytes / 4))[o] = reinterpret_cast<const uint4*>(s)[v];
    }
  }
}

int main()
{
  size_t const tiles = 12208;  // 100 M keys x 8 B / 64 KB
  uint32_t* d;
  cudaMalloc(&d, tiles * 65536);
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 2; ++mode)
    for (int bytes : {64, 128, 256, 512, 2048, 16384}) {
      float best = 1e9f;
      for (int it = 0; it < 5; ++it) {
        cudaEventRecord(a);
        if (mode == 0) probe<0><<<tiles, 512, 65536>>>(d, bytes, 16384);
        else probe<1><<<tiles, 512, 65536>>>(d, bytes, 16384);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it) best = ms < best ? ms : best;
      }
      printf("%s run=%5d B: %.3f ms for 0.8 GB (%.0f GB/s)  [%s]\n", mode == 0 ? "bulk" : "stg ", bytes, best,
             0.8 / best * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}

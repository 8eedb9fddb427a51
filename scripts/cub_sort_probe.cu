// Context measurement only (not part of the product, which uses no CUB): how fast is
// cub::DeviceRadixSort::SortPairs (onesweep) on this GPU for the bench's sort shape --
// 100 M (u32 key, u32 value) pairs, key bits [0, 30)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/cub_sort_probe.cu -o gpurun_variants/cub_sort_probe
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void fill(uint32_t* k, uint32_t* v, size_t n)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) {
    uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
    k[i] = (uint32_t)z & 0x3FFFFFFFu;
    v[i] = (uint32_t)i;
  }
}

int main(int argc, char** argv)
{
  size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
  uint32_t *k0, *k1, *v0, *v1;
  cudaMalloc(&k0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, n, 0, 30);
  void* tmp; cudaMalloc(&tmp, tmp_bytes);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int bits : {30, 32}) {
    float best = 1e9f, sum = 0;
    for (int it = 0; it < 8; ++it) {
      fill<<<(unsigned)((n + 255) / 256), 256>>>(k0, v0, n);
      cudaEventRecord(a);
      cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, bits);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (it >= 3) { sum += ms; if (ms < best) best = ms; }
    }
    printf("cub SortPairs n=%zu bits=%d: avg %.3f ms best %.3f ms (%s)\n", n, bits, sum / 5, best,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

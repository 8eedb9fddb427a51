// sort_bench.cu -- A/B harness for the hand-written onesweep radix sort (cuspatial_b200/csrc/
// radix_sort.cu).  Not part of the product: compiles radix_sort.cu with a ranking variant chosen
// on the command line, checks it against std::stable_sort on the host, and times it on the
// bench's sort shape (100 M (u32 key, u32 iota) pairs, 30-bit keys, all 32 key bits sorted).
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -DBSJ_SORT_RANK=2 \
//        scripts/sort_bench.cu cuspatial_b200/csrc/radix_sort.cu -o gpurun_variants/sort_rank2
#include "../cuspatial_b200/csrc/radix_sort.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

namespace bsj {
std::atomic<u64> g_launch_count{0};
void ensure_pool_configured() {}
void prof_mark(const char*) {}
int num_sms()
{
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, 0);
  return n;
}
void configure_once_per_device(int, void (*f)()) { f(); }
}  // namespace bsj

using namespace bsj;
#ifndef BSJ_SORT_RANK
#define BSJ_SORT_RANK_PRINT -1 /* library default */
#else
#define BSJ_SORT_RANK_PRINT BSJ_SORT_RANK
#endif

__global__ void fill(u32* k, size_t n, u32 mask, int mode)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) {
    uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
    u32 v = (u32)z & mask;
    if (mode == 1) v = (v % 1000u) * 7919u;         // heavy duplicates
    if (mode == 2) v = (u32)(i / 3) & mask;          // presorted runs of equal keys
    k[i] = v;
  }
}

static bool run_once(size_t n, u32 mask, int mode, int bits, bool check, float* ms_out)
{
  cudaStream_t s = 0;
  u32 *k0, *k1, *v0, *v1;
  cudaMalloc(&k0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
  sort_workspace ws;
  ws.alloc(n, s);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  bool in_a = true;
  int const iters = check ? 1 : 6;
  std::vector<u32> hk;
  for (int it = 0; it < iters; ++it) {
    fill<<<(unsigned)((n + 255) / 256), 256>>>(k0, n, mask, mode);
    if (check && it == 0) {
      hk.resize(n);
      cudaMemcpy(hk.data(), k0, n * 4, cudaMemcpyDeviceToHost);
    }
    sort_workspace_reset(ws, s);
    sort_histogram(k0, n, 0, bits, ws, s);
    cudaEventRecord(a, s);
    sort_passes(k0, v0, true, k1, v1, n, 0, bits, ws, s, &in_a);
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it >= 1 || iters == 1) best = std::min(best, ms);
  }
  *ms_out = best;
  bool ok = cudaGetLastError() == cudaSuccess;
  if (check && ok) {
    std::vector<u32> gk(n), gv(n), idx(n);
    cudaMemcpy(gk.data(), in_a ? k0 : k1, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(gv.data(), in_a ? v0 : v1, n * 4, cudaMemcpyDeviceToHost);
    std::iota(idx.begin(), idx.end(), 0u);
    u32 const km = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    std::stable_sort(idx.begin(), idx.end(),
                     [&](u32 x, u32 y) { return (hk[x] & km) < (hk[y] & km); });
    for (size_t i = 0; i < n; ++i)
      if (gv[i] != idx[i] || gk[i] != hk[idx[i]]) {
        printf("  MISMATCH at %zu: got (%u,%u) want (%u,%u)\n", i, gk[i], gv[i], hk[idx[i]], idx[i]);
        ok = false;
        break;
      }
  }
  cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1);
  cudaEventDestroy(a); cudaEventDestroy(b);
  return ok;
}

int main(int argc, char** argv)
{
  size_t const n_big = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
  bool all_ok = true;
  float ms;
  // correctness: random / duplicates / presorted, sizes that are not tile multiples, 1..4 passes
  struct { size_t n; u32 mask; int mode; int bits; } cases[] = {
    {1, 0x3FFFFFFFu, 0, 32}, {1000, 0x3FFFFFFFu, 0, 32}, {8192, 0xFFFFFFFFu, 0, 32},
    {8193, 0x3FFFFFFFu, 1, 32}, {3000001, 0x3FFFFFFFu, 0, 30}, {3000001, 0xFFFFFFFFu, 0, 32},
    {2000003, 0x3FFFFFFFu, 1, 32}, {2000003, 0x3FFFFFFFu, 2, 32}, {1500000, 0xFFFFu, 0, 16},
    {1500000, 0x1FFu, 0, 9},
  };
  bool const skip_checks = argc > 2;  // timing-only runs
  for (auto& c : cases) {
    if (skip_checks) break;
    bool ok = run_once(c.n, c.mask, c.mode, c.bits, true, &ms);
    printf("check n=%zu mask=%08x mode=%d bits=%d: %s\n", c.n, c.mask, c.mode, c.bits,
           ok ? "ok" : "FAILED");
    all_ok = all_ok && ok;
  }
  for (int bits : {32, 30}) {
    bool ok  = run_once(n_big, 0x3FFFFFFFu, 0, bits, false, &ms);
    int const p = passes_for_bits(0, bits);
    printf("RANK=%d radix_bits=%d n=%zu bits=%d: best %.3f ms total, %.3f ms/pass (%s)\n",
           BSJ_SORT_RANK_PRINT, kRadixBits, n_big, bits, ms, ms / p, ok ? "ok" : "CUDA ERROR");
    all_ok = all_ok && ok;
  }
  return all_ok ? 0 : 1;
}

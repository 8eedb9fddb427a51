// prim_bench.cu -- per-SM throughput of the warp-level primitives the radix-sort ranking is built
// from (shared-memory atomics, random LDS/STS, votes, match.any, shuffles), measured on the GPU
// this runs on.  Design aid only, not part of the product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/prim_bench.cu -o gpurun_variants/prim_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 512;  // 16 warps, 2 CTAs/SM like the sort kernel
constexpr int kIters   = 256;

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int OP>
__global__ void __launch_bounds__(kThreads, 2) prim_kernel(uint32_t* out, long long* cycles)
{
  __shared__ uint32_t s[16 * 256 * 2];
  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 16 * 256 * 2; i += kThreads) s[i] = 0;
  __syncthreads();
  uint32_t* const w32  = s + warp * 256;                                   // 256 x u32 per warp
  uint2* const w64     = reinterpret_cast<uint2*>(s) + warp * 256;        // 256 x u64 per warp
  uint32_t acc         = 0;
  uint32_t r           = hash32(tid * 2654435761u + blockIdx.x);
  long long const t0   = clock64();
#pragma unroll 4
  for (int it = 0; it < kIters; ++it) {
    r = r * 1664525u + 1013904223u;
    uint32_t const d = (r >> 12) & 0xFFu;  // random 8-bit digit
    if (OP == 0) {                          // ATOMS.OR random word
      atomicOr(&w32[d], 1u << lane);
    } else if (OP == 1) {                   // LDS.32 random
      acc += w32[d];
    } else if (OP == 2) {                   // STS.32 random
      w32[d] = r;
    } else if (OP == 3) {                   // LDS.64 random
      uint2 v = w64[d];
      acc += v.x + v.y;
    } else if (OP == 4) {                   // 8 votes -> peers mask
      uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        bool const bit = (d >> b) & 1u;
        uint32_t const m = __ballot_sync(0xFFFFFFFFu, bit);
        peers &= bit ? m : ~m;
      }
      acc += peers;
    } else if (OP == 5) {                   // match.any on the 8-bit digit
      acc += __match_any_sync(0xFFFFFFFFu, d);
    } else if (OP == 6) {                   // shuffle with a random source lane
      acc += __shfl_sync(0xFFFFFFFFu, r, d & 31);
    } else if (OP == 7) {                   // ATOMS.ADD with return, random word
      acc += atomicAdd(&w32[d], 1u);
    } else if (OP == 8) {                   // match.any on a 4-bit value
      acc += __match_any_sync(0xFFFFFFFFu, d & 15u);
    } else if (OP == 9) {                   // ATOMS.OR on the odd words of 64-bit entries (round-1 layout)
      atomicOr(&w64[d].y, 1u << lane);
    } else if (OP == 10) {                  // STS.64 random
      w64[d] = make_uint2(r, r);
    } else if (OP == 11) {                  // one vote
      acc += __ballot_sync(0xFFFFFFFFu, d & 1u);
    } else if (OP == 12) {                  // ATOMS.OR without conflicts (lane-private word)
      atomicOr(&w32[(d & 0xE0u) | lane], 1u << lane);
    } else if (OP == 13) {                  // baseline: the loop alone
      acc += d;
    } else if (OP == 20) {                  // one full ranking round, split count / match arrays
      uint32_t* const wc = w32;
      uint32_t* const wm = s + 16 * 256 + warp * 256;
      atomicOr(&wm[d], 1u << lane);
      __syncwarp();
      uint32_t const peers = wm[d];
      uint32_t const cnt   = wc[d];
      __syncwarp();
      uint32_t const below = peers & ((1u << lane) - 1u);
      if (below == 0) {
        wc[d] = cnt + __popc(peers);
        wm[d] = 0u;
      }
      acc += cnt + __popc(below);
      __syncwarp();
    } else if (OP == 21) {                  // placement: offset lookup + 64-bit store at a random slot
      uint32_t const pos = (w32[d] + r) & 4095u;
      reinterpret_cast<uint2*>(s)[pos] = make_uint2(r, acc);
    } else if (OP == 22) {                  // scatter read: linear 64-bit load + digit base lookup
      uint2 const e = reinterpret_cast<uint2*>(s)[(it * 32 + lane + warp * 64) & 4095];
      acc += w32[(e.x >> 8) & 0xFFu] + e.y;
    }
  }
  long long const t1 = clock64();
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * kThreads + tid] = acc + s[tid];
}

template <int OP>
void run(const char* name, uint32_t* out, long long* cyc, int sms)
{
  int const grid = sms * 2;
  prim_kernel<OP><<<grid, kThreads>>>(out, cyc);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  prim_kernel<OP><<<grid, kThreads>>>(out, cyc);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  long long h[8];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  // warp-instructions per SM = 2 CTAs x 16 warps x kIters
  double const per_sm_cycles = (double)h[0];
  printf("%-34s %8.1f us  %9.0f clk  -> %6.2f clk per warp-op per SM (32 warps resident)\n", name,
         ms * 1e3, per_sm_cycles, per_sm_cycles / (2.0 * 16 * kIters));
}

int main()
{
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)sms * 2 * kThreads * 4);
  cudaMalloc(&cyc, (size_t)sms * 2 * 8);
  run<13>("loop only", out, cyc, sms);
  run<0>("ATOMS.OR random u32", out, cyc, sms);
  run<9>("ATOMS.OR random, odd words of u64", out, cyc, sms);
  run<12>("ATOMS.OR conflict-free", out, cyc, sms);
  run<7>("ATOMS.ADD(return) random u32", out, cyc, sms);
  run<1>("LDS.32 random", out, cyc, sms);
  run<2>("STS.32 random", out, cyc, sms);
  run<3>("LDS.64 random", out, cyc, sms);
  run<10>("STS.64 random", out, cyc, sms);
  run<11>("1 vote", out, cyc, sms);
  run<4>("8 votes + combine (peer mask)", out, cyc, sms);
  run<5>("match.any 8-bit", out, cyc, sms);
  run<8>("match.any 4-bit", out, cyc, sms);
  run<6>("shfl random lane", out, cyc, sms);
  run<20>("ranking round (split arrays)", out, cyc, sms);
  run<21>("placement (lookup + STS.64 random)", out, cyc, sms);
  run<22>("scatter read (LDS.64 + base lookup)", out, cyc, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

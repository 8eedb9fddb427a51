// scan.cu -- single-pass exclusive scan with 64-bit accumulation.  See scan.cuh.
#include "scan.cuh"

namespace bsj {
namespace {

// Descriptor for a 64-bit chained scan: two words written with one 128-bit... kept simple instead:
// value (62 bits) and a 2-bit state packed in one u64: [63:62] state (0 none, 1 aggregate,
// 2 inclusive prefix), [61:0] value.  Sums on this path are < 2^62.
__device__ __forceinline__ u64 d_pack(u64 state, u64 v) { return (state << 62) | v; }

__global__ void __launch_bounds__(kScanBlock)
scan_kernel(const u32* __restrict__ in, u64* __restrict__ out, u64 n, u64* __restrict__ total,
            u64* __restrict__ desc, u32* __restrict__ ticket)
{
  __shared__ u64 s_warp[kScanBlock / 32];
  __shared__ u64 s_base;
  __shared__ u32 s_tile;
  int const tid = threadIdx.x;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  u64 const tile = s_tile;
  u64 const base = tile * kScanTile + (u64)tid * kScanIPT;

  u32 v[kScanIPT];
  u64 local = 0;
#pragma unroll
  for (int i = 0; i < kScanIPT; ++i) {
    v[i] = base + i < n ? in[base + i] : 0u;
    local += v[i];
  }
  // warp inclusive scan of 64-bit thread sums
  u64 incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u64 const t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  u64 wbase = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < kScanBlock / 32; ++w) {
    u64 const sw = s_warp[w];
    if (w < (tid >> 5)) wbase += sw;
    block_total += sw;
  }
  if (tid == 0) {
    u64 excl = 0;
    if (tile == 0) {
      st_relaxed_u64(desc, d_pack(2, block_total));
    } else {
      st_relaxed_u64(desc + tile, d_pack(1, block_total));
      i64 t = (i64)tile - 1;
      while (true) {
        u64 const d  = ld_relaxed_u64(desc + t);
        u64 const st = d >> 62;
        if (st == 2) {
          excl += d & ((1ull << 62) - 1);
          break;
        }
        if (st == 1) {
          excl += d & ((1ull << 62) - 1);
          --t;
          continue;
        }
        __nanosleep(20);
      }
      st_relaxed_u64(desc + tile, d_pack(2, excl + block_total));
    }
    s_base = excl;
    if ((tile + 1) * kScanTile >= n) *total = excl + block_total;
  }
  __syncthreads();
  u64 run = s_base + wbase + incl - local;
#pragma unroll
  for (int i = 0; i < kScanIPT; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

}  // namespace

void exclusive_scan_u32_to_u64(const u32* in, u64* out, u64 n, u64* total, cudaStream_t s)
{
  if (n == 0) {
    BSJ_CUDA_TRY(cudaMemsetAsync(total, 0, sizeof(u64), s));
    return;
  }
  u64 const tiles = (n + kScanTile - 1) / kScanTile;
  dev_buf<u64> desc(tiles, s);
  dev_buf<u32> ticket(1, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(desc.get(), 0, tiles * sizeof(u64), s));
  BSJ_CUDA_TRY(cudaMemsetAsync(ticket.get(), 0, sizeof(u32), s));
  scan_kernel<<<(unsigned)tiles, kScanBlock, 0, s>>>(in, out, n, total, desc.get(), ticket.get());
  BSJ_CHECK_LAUNCH();
}

}  // namespace bsj

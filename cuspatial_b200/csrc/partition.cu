// partition.cu -- Morton-range sharding of a point set across GPUs (new work: the reference is
// single-GPU).
//
// Each rank (1) computes the reference's Morton key of every local point
// (detail/index/construction/phase_1.cuh:78-85, same arithmetic as quadtree.cu) together with a
// histogram of the keys' leading bits, (2) after the ranks agreed on key-range splitters (an
// all-reduce of the histograms, done by the host layer over NCCL), STABLY partitions its points by
// destination rank: one pass, tile by tile -- per-destination ballots give the in-tile rank, a
// decoupled look-back over tile descriptors gives the tile's offset inside each destination
// bucket -- writing x, y and the global point id into contiguous per-destination send buffers.
// Stability (ascending global id inside a bucket) is what keeps the tie order of the reference's
// stable sort after the exchange.
#include "common.cuh"

namespace bsj {

// implemented in quadtree.cu
template <typename T>
void launch_point_keys(const void* x, const void* y, u64 n, double x_min, double x_max,
                       double y_min, double y_max, double scale, int max_depth, u32* keys,
                       cudaStream_t s);

namespace {

constexpr int kMaxRanks  = 32;
constexpr int kPartBlock = 256;
constexpr int kPartIPT   = 8;
constexpr int kPartTile  = kPartBlock * kPartIPT;

struct splitters_t {
  u32 key[kMaxRanks];  // rank r owns keys in [key[r-1], key[r]); key[R-1] unused
  int n_ranks;
};
// Per-destination output bases.  On the multi-GPU path these are PEER pointers (symmetric memory
// mapped over NVLink): the partition kernel's stores ARE the all-to-all exchange.
template <typename T>
struct dests_t {
  T* x[kMaxRanks];
  T* y[kMaxRanks];
  u32* gid[kMaxRanks];
};

// histogram of the keys' leading bits; bins are privatised in shared memory when they fit
constexpr int kHistSmemBins = 8192;
__global__ void __launch_bounds__(512)
key_histogram_kernel(const u32* __restrict__ keys, u64 n, int shift, u32 n_bins,
                     u32* __restrict__ bins)
{
  __shared__ u32 s_bins[kHistSmemBins];
  bool const use_smem = n_bins <= (u32)kHistSmemBins;
  if (use_smem) {
    for (u32 i = threadIdx.x; i < n_bins; i += blockDim.x) s_bins[i] = 0;
    __syncthreads();
  }
  u64 const stride = (u64)gridDim.x * blockDim.x;
  u64 const nvec   = (reinterpret_cast<uintptr_t>(keys) & 15) == 0 ? n / 4 : 0;
  for (u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    uint4 const k = __ldcs(reinterpret_cast<const uint4*>(keys) + v);
    if (use_smem) {
      atomicAdd(&s_bins[k.x >> shift], 1u); atomicAdd(&s_bins[k.y >> shift], 1u);
      atomicAdd(&s_bins[k.z >> shift], 1u); atomicAdd(&s_bins[k.w >> shift], 1u);
    } else {
      atomicAdd(&bins[k.x >> shift], 1u); atomicAdd(&bins[k.y >> shift], 1u);
      atomicAdd(&bins[k.z >> shift], 1u); atomicAdd(&bins[k.w >> shift], 1u);
    }
  }
  for (u64 i = nvec * 4 + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    u32 const b = keys[i] >> shift;
    if (use_smem) atomicAdd(&s_bins[b], 1u); else atomicAdd(&bins[b], 1u);
  }
  if (use_smem) {
    __syncthreads();
    for (u32 i = threadIdx.x; i < n_bins; i += blockDim.x)
      if (s_bins[i]) atomicAdd(&bins[i], s_bins[i]);
  }
}

// second-level histogram: only keys whose leading bits (key >> shift1) equal one of the target
// bins are counted, by their next bits ((key >> shift2) & (n_sub - 1)).  Used to place the rank
// splitters INSIDE a heavy first-level bin (clustered data).
struct targets_t {
  u32 bin[kMaxRanks];
  int n;
};
__global__ void __launch_bounds__(512)
key_subhistogram_kernel(const u32* __restrict__ keys, u64 n, int shift1, targets_t tg, int shift2,
                        u32 n_sub, u32* __restrict__ bins)
{
  extern __shared__ u32 s_sub[];
  u32 const total = (u32)tg.n * n_sub;
  for (u32 i = threadIdx.x; i < total; i += blockDim.x) s_sub[i] = 0;
  __syncthreads();
  u64 const stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    u32 const k = __ldcs(keys + i);
    u32 const b = k >> shift1;
    for (int t = 0; t < tg.n; ++t)
      if (b == tg.bin[t]) atomicAdd(&s_sub[(u32)t * n_sub + ((k >> shift2) & (n_sub - 1))], 1u);
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < total; i += blockDim.x)
    if (s_sub[i]) atomicAdd(&bins[i], s_sub[i]);
}

__device__ __forceinline__ int dest_of(u32 key, const splitters_t& sp)
{
  int d = 0;
  for (int r = 0; r + 1 < sp.n_ranks; ++r) d += key >= sp.key[r];
  return d;
}

// descriptors: [tile][kMaxRanks] u64 {tag, value}
// The tile is first re-ordered by destination in shared memory so that what leaves the SM (over
// NVLink for remote destinations) are contiguous per-destination runs, not 8-byte scatters.
template <typename T>
struct part_smem {
  T x[kPartTile];
  T y[kPartTile];
  u32 gid[kPartTile];
  u32 warp_cnt[kPartBlock / 32][kMaxRanks];
  u32 warp_off[kPartBlock / 32][kMaxRanks];
  u32 bin_start[kMaxRanks + 1];  // exclusive scan of the tile's per-destination counts
  u32 base[kMaxRanks];           // tile's first slot inside each destination bucket
  u32 tile;
};

template <typename T>
__global__ void __launch_bounds__(kPartBlock)
partition_kernel(const u32* __restrict__ keys, const T* __restrict__ x, const T* __restrict__ y,
                 u32 n, u32 gid_base, splitters_t sp, dests_t<T> dst, u64* __restrict__ desc,
                 u32* __restrict__ ticket)
{
  extern __shared__ __align__(16) unsigned char part_smem_raw[];
  part_smem<T>& sm = *reinterpret_cast<part_smem<T>*>(part_smem_raw);
  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int const R   = sp.n_ranks;
  if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
  __syncthreads();
  u32 const tile = sm.tile;
  // blocked-by-warp layout keeps the original order: warp w owns items [w*IPT*32, (w+1)*IPT*32)
  u32 const tile_base = tile * kPartTile;
  u32 const warp_base = tile_base + warp * (kPartIPT * 32);
  u32 const valid     = min((u32)kPartTile, n - tile_base);
  u32 const lt        = lanemask_lt();

  int dest[kPartIPT];
  u32 rank[kPartIPT];
  u32 cnt = 0;  // lane r < R accumulates this warp's count for destination r
#pragma unroll
  for (int i = 0; i < kPartIPT; ++i) {
    u32 const idx = warp_base + i * 32 + lane;
    dest[i]       = idx < n ? dest_of(__ldcs(keys + idx), sp) : -1;
    rank[i]       = 0;
    for (int r = 0; r < R; ++r) {
      u32 const m      = __ballot_sync(0xffffffffu, dest[i] == r);
      u32 const before = __shfl_sync(0xffffffffu, cnt, r);
      if (dest[i] == r) rank[i] = before + __popc(m & lt);
      if (lane == r) cnt += __popc(m);
    }
  }
  if (lane < R) sm.warp_cnt[warp][lane] = cnt;
  __syncthreads();
  // per-destination totals, warp offsets and the chained scan over tiles (column = destination)
  if (tid < R) {
    u32 total = 0;
    for (int w = 0; w < kPartBlock / 32; ++w) {
      sm.warp_off[w][tid] = total;
      total += sm.warp_cnt[w][tid];
    }
    sm.bin_start[tid + 1] = total;  // turned into a prefix below
    u64* const col = desc + tid;
    u32 excl       = 0;
    if (tile == 0) {
      st_relaxed_u64(col, lb_pack(3u, total));
    } else {
      st_relaxed_u64(col + (u64)tile * kMaxRanks, lb_pack(2u, total));
      i64 t = (i64)tile - 1;
      while (true) {
        u64 const v    = ld_relaxed_u64(col + (u64)t * kMaxRanks);
        u32 const flag = (u32)(v >> 32);
        if (flag == 3u) {
          excl += (u32)v;
          break;
        }
        if (flag == 2u) {
          excl += (u32)v;
          --t;
        }
      }
      st_relaxed_u64(col + (u64)tile * kMaxRanks, lb_pack(3u, excl + total));
    }
    sm.base[tid] = excl;
  }
  __syncthreads();
  if (tid == 0) {
    sm.bin_start[0] = 0;
    for (int r = 0; r < R; ++r) sm.bin_start[r + 1] += sm.bin_start[r];
  }
  __syncthreads();
  // stage the tile ordered by (destination, original order)
#pragma unroll
  for (int i = 0; i < kPartIPT; ++i) {
    if (dest[i] >= 0) {
      u32 const idx  = warp_base + i * 32 + lane;
      int const d    = dest[i];
      u32 const slot = sm.bin_start[d] + sm.warp_off[warp][d] + rank[i];
      sm.x[slot]     = __ldcs(x + idx);
      sm.y[slot]     = __ldcs(y + idx);
      sm.gid[slot]   = gid_base + idx;
    }
  }
  __syncthreads();
  // coalesced per-destination runs out of shared memory
  for (u32 j = tid; j < valid; j += kPartBlock) {
    int d = 0;
    while (d + 1 < R && j >= sm.bin_start[d + 1]) ++d;
    u32 const o   = sm.base[d] + (j - sm.bin_start[d]);
    dst.x[d][o]   = sm.x[j];
    dst.y[d][o]   = sm.y[j];
    dst.gid[d][o] = sm.gid[j];
  }
}

template <typename T>
void partition_t(const u32* keys, const void* x, const void* y, u64 n, u32 gid_base,
                 const u32* h_splitters, int n_ranks, void* const* dst_x, void* const* dst_y,
                 u32* const* dst_gid, cudaStream_t s)
{
  splitters_t sp{};
  sp.n_ranks = n_ranks;
  for (int r = 0; r + 1 < n_ranks; ++r) sp.key[r] = h_splitters[r];
  dests_t<T> dst{};
  for (int r = 0; r < n_ranks; ++r) {
    dst.x[r]   = (T*)dst_x[r];
    dst.y[r]   = (T*)dst_y[r];
    dst.gid[r] = dst_gid[r];
  }
  u32 const tiles = (u32)div_up(n, kPartTile);
  dev_buf<u64> desc((size_t)tiles * kMaxRanks, s);
  dev_buf<u32> ticket(1, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(desc.get(), 0, desc.size() * sizeof(u64), s));
  BSJ_CUDA_TRY(cudaMemsetAsync(ticket.get(), 0, sizeof(u32), s));
  configure_once_per_device(2, [] {  // per device, not per process
    BSJ_CUDA_TRY(cudaFuncSetAttribute(partition_kernel<float>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(part_smem<float>)));
    BSJ_CUDA_TRY(cudaFuncSetAttribute(partition_kernel<double>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(part_smem<double>)));
  });
  partition_kernel<T><<<tiles, kPartBlock, sizeof(part_smem<T>), s>>>(
    keys, (const T*)x, (const T*)y, (u32)n, gid_base, sp, dst, desc.get(), ticket.get());
  BSJ_CHECK_LAUNCH();
  BSJ_CUDA_TRY(cudaStreamSynchronize(s));
}

}  // namespace

void point_keys_histogram_impl(const void* x, const void* y, int dtype, u64 n, double x_min,
                               double x_max, double y_min, double y_max, double scale,
                               int max_depth, int hist_shift, u32* keys, u32* bins, u64 n_bins,
                               cudaStream_t s)
{
  if (n == 0) return;
  int const d = std::max(0, std::min(15, max_depth));
  if (dtype == BSJ_FLOAT32)
    launch_point_keys<float>(x, y, n, x_min, x_max, y_min, y_max, scale, d, keys, s);
  else
    launch_point_keys<double>(x, y, n, x_min, x_max, y_min, y_max, scale, d, keys, s);
  if (bins) {
    BSJ_EXPECTS(hist_shift >= 0 && hist_shift < 32 && (0xFFFFFFFFull >> hist_shift) < n_bins,
                "histogram does not cover the key range");
    int const grid = (int)std::min<u64>((u64)num_sms() * 2, (u64)div_up(n, 2048));
    key_histogram_kernel<<<std::max(grid, 1), 512, 0, s>>>(keys, n, hist_shift, (u32)n_bins, bins);
    BSJ_CHECK_LAUNCH();
  }
  BSJ_CUDA_TRY(cudaStreamSynchronize(s));
}

void key_subhistogram_impl(const u32* keys, u64 n, int shift1, const u32* h_targets, int n_targets,
                           int shift2, u32 n_sub, u32* bins, cudaStream_t s)
{
  BSJ_EXPECTS(n_targets >= 0 && n_targets <= kMaxRanks, "too many target bins");
  BSJ_EXPECTS(n_sub >= 1 && (n_sub & (n_sub - 1)) == 0 && (u64)n_targets * n_sub <= 12288,
              "sub-histogram does not fit shared memory");
  BSJ_EXPECTS(shift1 >= shift2 && shift1 < 32 && shift2 >= 0, "invalid histogram shifts");
  if (n == 0 || n_targets == 0) return;
  targets_t tg{};
  tg.n = n_targets;
  for (int t = 0; t < n_targets; ++t) tg.bin[t] = h_targets[t];
  int const grid = (int)std::min<u64>((u64)num_sms() * 2, (u64)div_up(n, 2048));
  key_subhistogram_kernel<<<std::max(grid, 1), 512, (size_t)n_targets * n_sub * sizeof(u32), s>>>(
    keys, n, shift1, tg, shift2, n_sub, bins);
  BSJ_CHECK_LAUNCH();
  BSJ_CUDA_TRY(cudaStreamSynchronize(s));
}

void partition_points_impl(const u32* keys, const void* x, const void* y, int dtype, u64 n,
                           u32 gid_base, const u32* h_splitters, int n_ranks,
                           void* const* dst_x, void* const* dst_y, u32* const* dst_gid,
                           cudaStream_t s)
{
  BSJ_EXPECTS(n_ranks >= 1 && n_ranks <= kMaxRanks, "unsupported number of ranks");
  BSJ_EXPECTS(n < 0xFFFFFFFFull, "number of points must fit uint32 indices");
  if (n == 0) return;
  if (dtype == BSJ_FLOAT32)
    partition_t<float>(keys, x, y, n, gid_base, h_splitters, n_ranks, dst_x, dst_y, dst_gid, s);
  else
    partition_t<double>(keys, x, y, n, gid_base, h_splitters, n_ranks, dst_x, dst_y, dst_gid, s);
}

}  // namespace bsj

// partition.cu -- Morton-range sharding of a point set across GPUs (new work: the reference is
// single-GPU).
//
// Every rank
//   (1) computes the reference's Morton key of each local point
//       (detail/index/construction/phase_1.cuh:78-85, same arithmetic as quadtree.cu) together
//       with a histogram of the keys' leading bits;
//   (2) after the histograms have been summed over the ranks (one all-reduce, done by the host
//       layer on device buffers), derives the key-range splitters ON THE DEVICE -- two levels: the
//       first-level bin that contains each rank boundary, then, from a second histogram of the
//       next key bits inside those bins, the sub-bin where the boundary falls -- and its own send
//       counts per destination; after an all-gather of those counts, its write offsets inside
//       every destination's receive buffer.  All of it lives in one small device struct
//       (bsj_shard_plan): no host round trip between the key histogram and the exchange;
//   (3) STABLY partitions (key, global id) by destination rank: one pass, tile by tile -- votes
//       give the in-tile rank, a decoupled look-back over tile descriptors gives the tile's offset
//       inside each destination bucket -- and writes each per-destination run straight into the
//       destination GPU's receive buffer (peer memory over NVLink): the stores ARE the all-to-all.
//       Runs leave the SM as 1-D bulk copies (cp.async.bulk shared -> global, the TMA engine)
//       between scalar head/tail elements.  Only 8 bytes per point cross NVLink: the owner sorts
//       the keys it receives and pulls coordinates on demand (pip.cu, coordinate segments) for the
//       few points whose finest cell is touched by a polygon edge.
// Stability (ascending global id inside a bucket) is what keeps the tie order of the reference's
// stable sort after the exchange.
#include "common.cuh"

namespace bsj {

// implemented in quadtree.cu
template <typename T>
void launch_point_keys(const void* x, const void* y, u64 n, double x_min, double x_max,
                       double y_min, double y_max, double scale, int max_depth, u32* keys,
                       u32* point_flags, cudaStream_t s, u32* lead_bins, int bin_shift,
                       u32 n_bins);

namespace {

constexpr int kMaxRanks = BSJ_MAX_RANKS;
static_assert(kMaxRanks == 32, "plan kernels use one warp lane per rank");

// ---------------------------------------------------------------------------------------------
// histogram of the keys' leading bits; bins are privatised in shared memory when they fit
// ---------------------------------------------------------------------------------------------
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

// second-level histogram: only keys whose leading bits (key >> hist_shift) equal one of the
// plan's target bins are counted, by their next bits.  Used to place the rank splitters INSIDE a
// heavy first-level bin (clustered data).  Target bins are read from the device plan.
__global__ void __launch_bounds__(512)
key_subhistogram_kernel(const u32* __restrict__ keys, u64 n, const bsj_shard_plan* __restrict__ plan,
                        u32* __restrict__ bins)
{
  extern __shared__ u32 s_sub[];
  __shared__ u32 s_target[kMaxRanks];
  u32 const nt = plan->n_targets, n_sub = plan->n_sub;
  int const shift1 = (int)plan->hist_shift, shift2 = (int)plan->sub_shift;
  u32 const total = nt * n_sub;
  if (threadIdx.x < nt) s_target[threadIdx.x] = plan->target_bin[threadIdx.x];
  for (u32 i = threadIdx.x; i < total; i += blockDim.x) s_sub[i] = 0;
  __syncthreads();
  u64 const stride = (u64)gridDim.x * blockDim.x;
  auto tally = [&](u32 k) {
    u32 const b = k >> shift1;
    for (u32 t = 0; t < nt; ++t)
      if (b == s_target[t]) atomicAdd(&s_sub[t * n_sub + ((k >> shift2) & (n_sub - 1))], 1u);
  };
  // 128-bit loads, two in flight per thread: the kernel only has to stream the keys
  u64 const nvec = (reinterpret_cast<uintptr_t>(keys) & 15) == 0 ? n / 4 : 0;
  u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (; v + stride < nvec; v += 2 * stride) {
    uint4 const a = __ldcs(reinterpret_cast<const uint4*>(keys) + v);
    uint4 const c = __ldcs(reinterpret_cast<const uint4*>(keys) + v + stride);
    tally(a.x); tally(a.y); tally(a.z); tally(a.w);
    tally(c.x); tally(c.y); tally(c.z); tally(c.w);
  }
  for (; v < nvec; v += stride) {
    uint4 const a = __ldcs(reinterpret_cast<const uint4*>(keys) + v);
    tally(a.x); tally(a.y); tally(a.z); tally(a.w);
  }
  for (u64 i = nvec * 4 + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    tally(__ldcs(keys + i));
  __syncthreads();
  for (u32 i = threadIdx.x; i < total; i += blockDim.x)
    if (s_sub[i]) atomicAdd(&bins[i], s_sub[i]);
}

// ---------------------------------------------------------------------------------------------
// the sharding plan, level 1: which first-level bin holds each rank boundary
// (host restatement: cuspatial_b200/multi_gpu.py refine_splitters)
// ---------------------------------------------------------------------------------------------
struct rank_sizes_t {
  u32 n[kMaxRanks];
};

__global__ void __launch_bounds__(1024)
plan_level1_kernel(const u32* __restrict__ global_hist, u32 n_bins, rank_sizes_t sizes, u32 n_ranks,
                   u32 rank, u32 hist_shift, u32 sub_shift, u32 n_sub, bsj_shard_plan* plan)
{
  __shared__ u32 s_csum[kHistSmemBins];  // inclusive prefix sums of the global histogram
  __shared__ u32 s_warp[32];
  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // block-wide inclusive scan, each thread owns a contiguous chunk
  u32 const per = (n_bins + 1023) / 1024;
  u32 const lo = min((u32)tid * per, n_bins), hi = min(lo + per, n_bins);
  u32 sum = 0;
  for (u32 i = lo; i < hi; ++i) sum += global_hist[i];
  u32 const incl = warp_inclusive_scan(sum);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  u32 base = 0;
  for (int w = 0; w < warp; ++w) base += s_warp[w];
  u32 run = base + incl - sum;
  for (u32 i = lo; i < hi; ++i) {
    run += global_hist[i];
    s_csum[i] = run;
  }
  __syncthreads();
  u64 const total = n_bins ? s_csum[n_bins - 1] : 0;
  if (tid == 0) {
    plan->n_ranks = n_ranks; plan->rank = rank;
    plan->hist_shift = hist_shift; plan->sub_shift = sub_shift; plan->n_sub = n_sub;
    plan->status = 0;
    u32 g = 0;
    for (u32 r = 0; r < n_ranks; ++r) {
      plan->gid_base[r] = g;
      g += sizes.n[r];
    }
    plan->gid_base[n_ranks] = g;
  }
  if ((u32)tid + 1 < n_ranks) {
    u32 const r      = tid + 1;
    u64 const target = (total * r + n_ranks - 1) / n_ranks;
    // first bin whose inclusive prefix reaches the target
    u32 a = 0, b = n_bins;
    while (a < b) {
      u32 const m = a + ((b - a) >> 1);
      if ((u64)s_csum[m] < target) a = m + 1; else b = m;
    }
    u32 const bin    = min(a, n_bins - 1);
    u64 const before = (u64)s_csum[bin] - global_hist[bin];
    plan->bound_bin[tid]     = bin;
    plan->bound_missing[tid] = target > before ? (u32)(target - before) : 0u;
  }
  __syncthreads();
  if (tid == 0) {  // distinct boundary bins, ascending (boundary bins are non-decreasing)
    u32 nt = 0;
    for (u32 i = 0; i + 1 < n_ranks; ++i)
      if (i == 0 || plan->bound_bin[i] != plan->bound_bin[i - 1])
        plan->target_bin[nt++] = plan->bound_bin[i];
    plan->n_targets = nt;
  }
}

// level 2: the splitters (inside the boundary bins, from the summed sub-histograms) and this
// rank's send counts per destination (from its own histograms)
// (host restatement: multi_gpu.py splitters_from_subhist + send_counts_for)
__global__ void __launch_bounds__(1024)
plan_level2_kernel(const u32* __restrict__ local_hist, u32 n_bins, const u32* __restrict__ local_sub,
                   const u32* __restrict__ global_sub, bsj_shard_plan* plan)
{
  __shared__ u32 s_split[kMaxRanks];
  __shared__ u32 s_cnt[kMaxRanks];
  __shared__ u32 s_target[kMaxRanks];
  int const tid = threadIdx.x;
  u32 const R = plan->n_ranks, nt = plan->n_targets, n_sub = plan->n_sub;
  u32 const shift = plan->hist_shift, shift2 = plan->sub_shift;
  if (tid < kMaxRanks) {
    s_cnt[tid]    = 0;
    s_target[tid] = (u32)tid < nt ? plan->target_bin[tid] : 0xFFFFFFFFu;
  }
  __syncthreads();
  if ((u32)tid + 1 < R) {
    u32 const b = plan->bound_bin[tid], missing = plan->bound_missing[tid];
    u32 t = 0;
    while (t + 1 < nt && s_target[t] != b) ++t;
    // first sub-bin whose inclusive prefix reaches `missing` (none needed when nothing is missing)
    i64 j = -1;
    if (missing > 0) {
      u64 run = 0;
      j       = (i64)n_sub - 1;
      for (u32 k = 0; k < n_sub; ++k) {
        run += global_sub[t * n_sub + k];
        if (run >= missing) {
          j = k;
          break;
        }
      }
    }
    u64 const v = ((u64)b << shift) + ((u64)(j + 1) << shift2);  // first key of the next rank
    s_split[tid] = (u32)min(v, (u64)0xFFFFFFFFull);
  }
  __syncthreads();
  if (tid == 0) {
    for (u32 i = 1; i + 1 < R; ++i) s_split[i] = max(s_split[i], s_split[i - 1]);
    for (u32 i = 0; i + 1 < R; ++i) plan->splitter[i] = s_split[i];
    for (u32 i = R > 0 ? R - 1 : 0; i < (u32)kMaxRanks; ++i) plan->splitter[i] = 0xFFFFFFFFu;
  }
  __syncthreads();
  auto owner_of = [&](u64 first_key) {
    u32 d = 0;
    for (u32 r = 0; r + 1 < R; ++r) d += (u64)s_split[r] <= first_key;
    return d;
  };
  for (u32 i = tid; i < n_bins; i += blockDim.x) {
    u32 const c = local_hist[i];
    if (c == 0) continue;
    bool is_target = false;
    for (u32 t = 0; t < nt; ++t) is_target = is_target || s_target[t] == i;
    if (!is_target) atomicAdd(&s_cnt[owner_of((u64)i << shift)], c);
  }
  for (u32 i = tid; i < nt * n_sub; i += blockDim.x) {
    u32 const c = local_sub[i];
    if (c == 0) continue;
    u32 const t = i / n_sub, j = i % n_sub;
    atomicAdd(&s_cnt[owner_of(((u64)s_target[t] << shift) + ((u64)j << shift2))], c);
  }
  __syncthreads();
  if (tid < kMaxRanks) plan->send_count[tid] = (u32)tid < R ? s_cnt[tid] : 0u;
}

// after the all-gather of the send counts: where this rank's bucket starts inside every
// destination's receive buffer, and how much every rank receives
__global__ void plan_finalize_kernel(const u32* __restrict__ counts_matrix, u64 capacity,
                                     bsj_shard_plan* plan)
{
  u32 const R = plan->n_ranks, me = plan->rank;
  u32 const d = threadIdx.x;
  bool over   = false;
  if (d < R) {
    u64 off = 0, tot = 0;
    for (u32 s = 0; s < R; ++s) {
      u32 const c = counts_matrix[s * R + d];  // rank s sends c points to rank d
      if (s < me) off += c;
      tot += c;
    }
    plan->send_offset[d] = (u32)off;
    plan->recv_total[d]  = (u32)min(tot, (u64)0xFFFFFFFFull);
    over                 = tot > capacity;
  }
  if (__any_sync(0xffffffffu, over) && d == 0) plan->status = 1;
}

// ---------------------------------------------------------------------------------------------
// fused partition + all-to-all of (key, global id)
// ---------------------------------------------------------------------------------------------
constexpr int kPartBlock = 256;  // 4 CTAs per SM: a CTA idles while its peer stores complete
constexpr int kPartIPT   = 16;
constexpr int kPartTile  = kPartBlock * kPartIPT;       // 8192 keys
constexpr int kPartSlots = kPartTile + 8 * kMaxRanks;  // + alignment padding per destination

struct key_dests_t {  // per-destination receive buffers (peer pointers over NVLink)
  u32* key[kMaxRanks];
  u32* gid[kMaxRanks];
};

struct part_smem {
  alignas(16) u32 key[kPartSlots];
  alignas(16) u32 gid[kPartSlots];
  u32 warp_cnt[kPartBlock / 32][kMaxRanks];
  u32 warp_off[kPartBlock / 32][kMaxRanks];
  u32 slot0[kMaxRanks];   // first shared-memory slot of the destination's run (co-aligned)
  u32 count[kMaxRanks];   // elements of this tile going to the destination
  u32 gfirst[kMaxRanks];  // element index inside the destination buffer where the run starts
  u32 split[kMaxRanks];
};

__device__ __forceinline__ u32 smem_addr(const void* p)
{
  return (u32)__cvta_generic_to_shared(p);
}

// descriptors: [tile][kMaxRanks] u64 {tag, value}
template <bool BULK>
__global__ void __launch_bounds__(kPartBlock, 4)
partition_keys_kernel(const u32* __restrict__ keys, u32 n, const bsj_shard_plan* __restrict__ plan,
                      key_dests_t dst, u64* __restrict__ desc)
{
  extern __shared__ __align__(16) unsigned char part_smem_raw[];
  part_smem& sm = *reinterpret_cast<part_smem*>(part_smem_raw);
  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tile id = blockIdx.x: CTAs of a 1-D grid start in index order, so every predecessor the
  // look-back below can wait for is already resident or finished
  u32 const tile = blockIdx.x;
  // blocked-by-warp layout keeps the original order: warp w owns items [w*IPT*32, (w+1)*IPT*32)
  u32 const tile_base = tile * kPartTile;
  u32 const warp_base = tile_base + warp * (kPartIPT * 32);
  u32 const lt        = lanemask_lt();

  u32 key[kPartIPT];
  u32 dr[kPartIPT];  // destination << 16 | rank among the warp's earlier items of that destination
  u32 cnt = 0;       // lane r < R accumulates this warp's count for destination r
#pragma unroll
  for (int i = 0; i < kPartIPT; ++i) {  // the keys are in flight while the plan is read
    u32 const idx = warp_base + i * 32 + lane;
    key[i]        = idx < n ? __ldcs(keys + idx) : 0u;
  }
  if (plan->status != 0) return;  // receive capacity exceeded: the host re-plans
  int const R = (int)plan->n_ranks;
  if (tid < kMaxRanks) sm.split[tid] = plan->splitter[tid];
  u32 const gid0 = plan->gid_base[plan->rank];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kPartIPT; ++i) {
    u32 const idx = warp_base + i * 32 + lane;
    int d         = 0;
    for (int r = 0; r + 1 < R; ++r) d += key[i] >= sm.split[r];
    if (idx >= n) d = -1;
    u32 rk = 0;
    for (int r = 0; r < R; ++r) {
      u32 const m      = __ballot_sync(0xffffffffu, d == r);
      u32 const before = __shfl_sync(0xffffffffu, cnt, r);
      if (d == r) rk = before + __popc(m & lt);
      if (lane == r) cnt += __popc(m);
    }
    dr[i] = d < 0 ? 0xFFFFFFFFu : ((u32)d << 16) | rk;
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
    sm.count[tid]  = total;
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
    sm.gfirst[tid] = plan->send_offset[tid] + excl;
  }
  __syncthreads();
  if (tid == 0) {
    // runs are laid out one after the other, each starting at a slot that is congruent (mod 4
    // elements = 16 bytes) to its destination index, so that the aligned body of a run is
    // 16-byte aligned in shared AND in global memory: one bulk copy moves it
    u32 s = 0;
    for (int r = 0; r < R; ++r) {
      s           = ((s + 3u) & ~3u) + (sm.gfirst[r] & 3u);
      sm.slot0[r] = s;
      s += sm.count[r];
    }
  }
  __syncthreads();
  // stage the tile ordered by (destination, original order)
#pragma unroll
  for (int i = 0; i < kPartIPT; ++i) {
    if (dr[i] != 0xFFFFFFFFu) {
      int const d    = (int)(dr[i] >> 16);
      u32 const slot = sm.slot0[d] + sm.warp_off[warp][d] + (dr[i] & 0xFFFFu);
      sm.key[slot]   = key[i];
      sm.gid[slot]   = gid0 + warp_base + i * 32 + lane;
    }
  }
  if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (BULK) {
    // warp r moves destination r's run: scalar head and tail, the 16-byte-aligned body as two
    // bulk copies (keys, ids) issued by one lane
    for (int r = warp; r < R; r += kPartBlock / 32) {
      u32 const c = sm.count[r], g = sm.gfirst[r], s0 = sm.slot0[r];
      if (c == 0) continue;
      u32 const head = min((4u - (g & 3u)) & 3u, c);
      u32 const body = (c - head) & ~3u;
      u32 const tail = c - head - body;
      if ((u32)lane < head) {
        dst.key[r][g + lane] = sm.key[s0 + lane];
        dst.gid[r][g + lane] = sm.gid[s0 + lane];
      }
      if ((u32)lane < tail) {
        dst.key[r][g + head + body + lane] = sm.key[s0 + head + body + lane];
        dst.gid[r][g + head + body + lane] = sm.gid[s0 + head + body + lane];
      }
      if (lane == 0 && body) {
        u32 const bytes = body * 4u;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                       dst.key[r] + g + head),
                     "r"(smem_addr(&sm.key[s0 + head])), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                       dst.gid[r] + g + head),
                     "r"(smem_addr(&sm.gid[s0 + head])), "r"(bytes)
                     : "memory");
      }
    }
    if (lane == 0) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    // coalesced per-destination runs out of shared memory, 128-bit stores on the aligned body
    for (int r = 0; r < R; ++r) {
      u32 const c = sm.count[r], g = sm.gfirst[r], s0 = sm.slot0[r];
      u32 const head = min((4u - (g & 3u)) & 3u, c);
      u32 const nvec = (c - head) >> 2;
      if ((u32)tid < head) {
        dst.key[r][g + tid] = sm.key[s0 + tid];
        dst.gid[r][g + tid] = sm.gid[s0 + tid];
      }
      for (u32 v = tid; v < nvec; v += kPartBlock) {
        u32 const o = head + v * 4;
        *reinterpret_cast<uint4*>(dst.key[r] + g + o) = *reinterpret_cast<const uint4*>(&sm.key[s0 + o]);
        *reinterpret_cast<uint4*>(dst.gid[r] + g + o) = *reinterpret_cast<const uint4*>(&sm.gid[s0 + o]);
      }
      u32 const o = head + nvec * 4 + tid;
      if (o < c) {
        dst.key[r][g + o] = sm.key[s0 + o];
        dst.gid[r][g + o] = sm.gid[s0 + o];
      }
    }
  }
}

}  // namespace

void point_keys_histogram_impl(const void* x, const void* y, int dtype, u64 n, double x_min,
                               double x_max, double y_min, double y_max, double scale,
                               int max_depth, int hist_shift, u32* keys, u32* bins, u64 n_bins,
                               u32* point_flags, cudaStream_t s)
{
  if (n == 0) return;
  int const d = std::max(0, std::min(15, max_depth));
  bool const fused = bins != nullptr && n_bins <= (u64)kHistSmemBins;
  if (bins)
    BSJ_EXPECTS(hist_shift >= 0 && hist_shift < 32 && (0xFFFFFFFFull >> hist_shift) < n_bins,
                "histogram does not cover the key range");
  // keys and the leading-bit histogram in ONE pass over the coordinates when the bins fit the
  // encode kernel's shared memory (always the case for the sharding plan's 8192 bins)
  if (dtype == BSJ_FLOAT32)
    launch_point_keys<float>(x, y, n, x_min, x_max, y_min, y_max, scale, d, keys, point_flags, s,
                             fused ? bins : nullptr, hist_shift, (u32)n_bins);
  else
    launch_point_keys<double>(x, y, n, x_min, x_max, y_min, y_max, scale, d, keys, point_flags, s,
                              fused ? bins : nullptr, hist_shift, (u32)n_bins);
  if (bins && !fused) {
    int const grid = (int)std::min<u64>((u64)num_sms() * 4, (u64)div_up(n, 2048));
    key_histogram_kernel<<<std::max(grid, 1), 512, 0, s>>>(keys, n, hist_shift, (u32)n_bins, bins);
    BSJ_CHECK_LAUNCH();
  }
}

void shard_plan_level1_impl(const u32* global_hist, u64 n_bins, const u32* h_rank_sizes,
                            int n_ranks, int rank, int hist_shift, int sub_shift, u32 n_sub,
                            bsj_shard_plan* plan, cudaStream_t s)
{
  BSJ_EXPECTS(n_ranks >= 1 && n_ranks <= kMaxRanks && rank >= 0 && rank < n_ranks,
              "unsupported number of ranks");
  BSJ_EXPECTS(n_bins >= 1 && n_bins <= (u64)kHistSmemBins, "histogram too large for the planner");
  BSJ_EXPECTS(n_sub >= 1 && (n_sub & (n_sub - 1)) == 0 &&
                (u64)std::max(n_ranks - 1, 1) * n_sub <= 12288,
              "sub-histogram does not fit shared memory");
  BSJ_EXPECTS(hist_shift >= sub_shift && hist_shift < 32 && sub_shift >= 0 &&
                (1u << (hist_shift - sub_shift)) == n_sub,
              "invalid histogram shifts");
  rank_sizes_t sz{};
  u64 total = 0;
  for (int r = 0; r < n_ranks; ++r) {
    sz.n[r] = h_rank_sizes[r];
    total += h_rank_sizes[r];
  }
  BSJ_EXPECTS(total < 0xFFFFC000ull, "total number of points must fit uint32 global indices");
  plan_level1_kernel<<<1, 1024, 0, s>>>(global_hist, (u32)n_bins, sz, (u32)n_ranks, (u32)rank,
                                        (u32)hist_shift, (u32)sub_shift, n_sub, plan);
  BSJ_CHECK_LAUNCH();
}

void shard_subhistogram_impl(const u32* keys, u64 n, const bsj_shard_plan* plan, int n_ranks,
                             u32 n_sub, u32* bins, cudaStream_t s)
{
  if (n == 0 || n_ranks <= 1) return;
  int const grid = (int)std::min<u64>((u64)num_sms() * 4, (u64)div_up(n, 4096));
  key_subhistogram_kernel<<<std::max(grid, 1), 512, (size_t)(n_ranks - 1) * n_sub * sizeof(u32),
                            s>>>(keys, n, plan, bins);
  BSJ_CHECK_LAUNCH();
}

void shard_plan_level2_impl(const u32* local_hist, u64 n_bins, const u32* local_sub,
                            const u32* global_sub, bsj_shard_plan* plan, cudaStream_t s)
{
  plan_level2_kernel<<<1, 1024, 0, s>>>(local_hist, (u32)n_bins, local_sub, global_sub, plan);
  BSJ_CHECK_LAUNCH();
}

void shard_plan_finalize_impl(const u32* counts_matrix, u64 capacity, bsj_shard_plan* plan,
                              cudaStream_t s)
{
  plan_finalize_kernel<<<1, 32, 0, s>>>(counts_matrix, capacity, plan);
  BSJ_CHECK_LAUNCH();
}

void partition_keys_impl(const u32* keys, u64 n, const bsj_shard_plan* plan, int n_ranks,
                         u32* const* dst_key, u32* const* dst_gid, int use_bulk_copy,
                         cudaStream_t s)
{
  BSJ_EXPECTS(n_ranks >= 1 && n_ranks <= kMaxRanks, "unsupported number of ranks");
  BSJ_EXPECTS(n < 0xFFFFFFFFull, "number of points must fit uint32 indices");
  if (n == 0) return;
  key_dests_t dst{};
  for (int r = 0; r < n_ranks; ++r) {
    BSJ_EXPECTS((reinterpret_cast<uintptr_t>(dst_key[r]) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dst_gid[r]) & 15) == 0,
                "receive buffers must be 16-byte aligned");
    dst.key[r] = dst_key[r];
    dst.gid[r] = dst_gid[r];
  }
  u32 const tiles = (u32)div_up(n, kPartTile);
  dev_buf<u64> desc((size_t)tiles * kMaxRanks, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(desc.get(), 0, desc.size() * sizeof(u64), s));
  configure_once_per_device(2, [] {  // per device, not per process
    BSJ_CUDA_TRY(cudaFuncSetAttribute(partition_keys_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(part_smem)));
    BSJ_CUDA_TRY(cudaFuncSetAttribute(partition_keys_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(part_smem)));
  });
  if (use_bulk_copy)
    partition_keys_kernel<true><<<tiles, kPartBlock, sizeof(part_smem), s>>>(
      keys, (u32)n, plan, dst, desc.get());
  else
    partition_keys_kernel<false><<<tiles, kPartBlock, sizeof(part_smem), s>>>(
      keys, (u32)n, plan, dst, desc.get());
  BSJ_CHECK_LAUNCH();
}

}  // namespace bsj

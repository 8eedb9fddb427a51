// radix_sort.cu -- onesweep LSD radix sort (u32 key, u32 value), sm_100a.  See radix_sort.cuh.
#include "radix_sort.cuh"

// Ranking scheme of the per-warp stable ranking (the shared-memory-pipe cost of the kernel):
//   0  match word (atomicOr of the lane bit) packed next to the running count, one 64-bit entry
//      per (warp, digit)                                       [round-1 form, kept for A/B runs]
//   1  same, count and match word in separate 32-bit arrays (all 32 banks instead of 16 pairs)
//   2  peers found with eight warp votes (no shared-memory traffic, no atomics); only the
//      running count lives in shared memory (one 32-bit load + one leader store per round)
//   3  as 1, but only the lowest peer touches the running count (atomicAdd returning the old
//      value) and hands it to its peers with a shuffle: one load less per round   [default]
#ifndef BSJ_SORT_RANK
#define BSJ_SORT_RANK 3
#endif
// Tile id: 0 = blockIdx.x (CTAs of a 1-D grid are dispatched in index order, so every
// predecessor a look-back can wait for is already resident or done -- what cub::DeviceScan has
// always relied on); 1 = dynamic ticket (one global atomic + a barrier before the first load)
#ifndef BSJ_SORT_TICKET
#define BSJ_SORT_TICKET 0
#endif
// Tile load: 1 = the tile's keys and values arrive in shared memory as two 1-D bulk copies (TMA
// engine, cp.async.bulk + mbarrier) issued by one thread at kernel entry -- no load instructions
// through the LSU pipe, and the values' HBM latency hides behind the whole ranking phase without
// holding registers; 0 = per-thread streaming loads.
#ifndef BSJ_SORT_TMA
#define BSJ_SORT_TMA 1
#endif

namespace bsj {

namespace {

constexpr int kWarps = kSortBlock / 32;

// shared-memory layout of one onesweep CTA (dynamic)
struct sort_smem {
  uint2 kv[kSortTile];  // tile-sorted {key, value} slots: one 64-bit store / load per element
#if BSJ_SORT_RANK == 0
  // per-warp, per-digit {x: running count -> exclusive warp offset, y: match mask of the
  // current round}
  uint2 whist[kWarps * kRadixDigits];
#else
  u32 wcnt[kWarps * kRadixDigits];   // running count -> exclusive warp offset (+ bin start)
#if BSJ_SORT_RANK == 1 || BSJ_SORT_RANK == 3
  u32 wmask[kWarps * kRadixDigits];  // match mask of the current round
#endif
#endif
  u32 bin_start[kRadixDigits];       // exclusive scan of the tile's digit totals
  u32 gbase[kRadixDigits];           // global destination of (digit, slot j): gbase[d] + j
  u32 warp_sums[kWarps];
  u32 gsums[kWarps];
  u32 tile;
  alignas(8) u64 mbar;  // completion barrier of the tile's bulk loads
};

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, u32 bytes, u64* bar)
{
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
      smem_u32(dst_smem)),
    "l"(src), "r"(bytes), "r"(smem_u32(bar))
    : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity)
{
  u32 done = 0;
  while (!done) {
    asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// generic histogram (used when the producer of the keys did not already fuse it)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) histogram_kernel(const u32* __restrict__ keys, u64 n,
                                                        int begin_bit, int passes,
                                                        u32* __restrict__ hist)
{
  __shared__ u32 s_hist[kMaxPasses * kRadixDigits];
  for (int i = threadIdx.x; i < kMaxPasses * kRadixDigits; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  u64 const stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    u32 const k = ld_stream(keys + i);
    for (int p = 0; p < passes; ++p)
      atomicAdd(&s_hist[p * kRadixDigits + ((k >> (begin_bit + p * kRadixBits)) & 0xFF)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * kRadixDigits; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// ---------------------------------------------------------------------------------------------
// one onesweep pass: rank -> look-back -> scatter
// `digit_counts` are the RAW digit counts of this pass (the histogram row): every CTA turns them
// into exclusive offsets itself (256 loads that hit L2), which removed the separate scan launch.
// ---------------------------------------------------------------------------------------------
template <bool IOTA>
__global__ void __launch_bounds__(kSortBlock, BSJ_SORT_MINBLOCKS)
onesweep_kernel(const u32* __restrict__ keys_in, const u32* __restrict__ vals_in,
                u32* __restrict__ keys_out, u32* __restrict__ vals_out, u32 n, int shift,
                const u32* __restrict__ digit_counts, u64* __restrict__ lookback,
                u32* __restrict__ ticket, u32 tag_agg, u32 tag_pre)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  sort_smem& sm = *reinterpret_cast<sort_smem*>(smem_raw);

  int const tid  = threadIdx.x;
  int const lane = tid & 31;
  int const warp = tid >> 5;

#if BSJ_SORT_TICKET
  if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
  __syncthreads();
  u32 const tile = sm.tile;
#else
  u32 const tile = blockIdx.x;
#endif
  u32 const tile_base = tile * (u32)kSortTile;
  u32 const valid     = min((u32)kSortTile, n - tile_base);
  u32 const warp_base = tile_base + warp * (kSortIPT * 32);

  // ---- tile load.  Full tiles of 16-byte-aligned inputs: two bulk copies into the (still
  // unused) slot array -- keys into its first half, values into its second half.
  u32* const skeys = reinterpret_cast<u32*>(sm.kv);
  u32* const svals = skeys + kSortTile;
#if BSJ_SORT_TMA
  bool const bulk = valid == (u32)kSortTile &&
                    ((reinterpret_cast<uintptr_t>(keys_in) |
                      (IOTA ? (uintptr_t)0 : reinterpret_cast<uintptr_t>(vals_in))) & 15) == 0;
  if (bulk && tid == 0) {
    mbar_init(&sm.mbar, 1);
    mbar_expect_tx(&sm.mbar, (IOTA ? 1u : 2u) * (u32)kSortTile * 4u);
    bulk_load(skeys, keys_in + tile_base, (u32)kSortTile * 4u, &sm.mbar);
    if (!IOTA) bulk_load(svals, vals_in + tile_base, (u32)kSortTile * 4u, &sm.mbar);
  }
#else
  constexpr bool bulk = false;
#endif
#if BSJ_SORT_RANK == 0
  for (int i = tid; i < kWarps * kRadixDigits; i += kSortBlock) sm.whist[i] = make_uint2(0u, 0u);
#else
  for (int i = tid; i < kWarps * kRadixDigits; i += kSortBlock) {
    sm.wcnt[i] = 0u;
#if BSJ_SORT_RANK == 1 || BSJ_SORT_RANK == 3
    sm.wmask[i] = 0u;
#endif
  }
#endif
  u32 key[kSortIPT];
  if (!bulk) {
    // warp-striped: item i of this lane is warp_base + i*32 + lane
#pragma unroll
    for (int i = 0; i < kSortIPT; ++i) {
      u32 const idx = warp_base + i * 32 + lane;
      key[i]        = idx < n ? ld_stream(keys_in + idx) : 0xFFFFFFFFu;
    }
  }
  __syncthreads();  // counters zeroed; the barrier's initialisation is visible to every waiter
  if (bulk) {
    mbar_wait(&sm.mbar, 0);
#pragma unroll
    for (int i = 0; i < kSortIPT; ++i) key[i] = skeys[warp * (kSortIPT * 32) + i * 32 + lane];
  }
  // this pass's global digit count (turned into an exclusive offset below)
  u32 gcount = 0;
  if (tid < kRadixDigits) gcount = __ldg(digit_counts + tid);

  // ---- per-warp stable ranking: rank[i] = number of keys with the same digit that precede
  // item i inside this warp's 16 x 32 items
  u32 const lt = lanemask_lt();
  unsigned short rank[kSortIPT];
#if BSJ_SORT_RANK == 0
  // Lanes holding the same digit find each other through a shared-memory match word (atomicOr of
  // the lane bit): the word sits next to the digit's running count, so ONE 64-bit load gives a
  // lane both its peer mask and the count of earlier items; the lowest peer then bumps the count
  // and clears the mask for the next round.
  uint2* const wh = sm.whist + warp * kRadixDigits;
  u32 const mybit = 1u << lane;
#pragma unroll
  for (int i = 0; i < kSortIPT; ++i) {
    u32 const d = (key[i] >> shift) & 0xFFu;
    atomicOr(&wh[d].y, mybit);
    __syncwarp();
    uint2 const cm = wh[d];  // {count before this round, peers of this round}
    __syncwarp();
    u32 const below = cm.y & lt;
    if (below == 0) wh[d] = make_uint2(cm.x + __popc(cm.y), 0u);
    rank[i] = (unsigned short)(cm.x + __popc(below));
    __syncwarp();
  }
#elif BSJ_SORT_RANK == 1
  u32* const wc = sm.wcnt + warp * kRadixDigits;
  u32* const wm = sm.wmask + warp * kRadixDigits;
  u32 const mybit = 1u << lane;
#pragma unroll
  for (int i = 0; i < kSortIPT; ++i) {
    u32 const d = (key[i] >> shift) & 0xFFu;
    atomicOr(&wm[d], mybit);
    __syncwarp();
    u32 const peers = wm[d];
    u32 const cnt   = wc[d];
    __syncwarp();
    u32 const below = peers & lt;
    if (below == 0) {
      wc[d] = cnt + __popc(peers);
      wm[d] = 0u;
    }
    rank[i] = (unsigned short)(cnt + __popc(below));
    __syncwarp();
  }
#elif BSJ_SORT_RANK == 3
  // as 1, but only the lowest peer touches the running count (atomicAdd returning the old
  // value) and hands it to its peers with a shuffle: one shared-memory load less per round
  u32* const wc = sm.wcnt + warp * kRadixDigits;
  u32* const wm = sm.wmask + warp * kRadixDigits;
  u32 const mybit = 1u << lane;
#pragma unroll
  for (int i = 0; i < kSortIPT; ++i) {
    u32 const d = (key[i] >> shift) & 0xFFu;
    atomicOr(&wm[d], mybit);
    __syncwarp();
    u32 const peers = wm[d];
    u32 const below = peers & lt;
    u32 cnt = 0;
    __syncwarp();
    if (below == 0) {
      cnt   = atomicAdd(&wc[d], __popc(peers));
      wm[d] = 0u;
    }
    cnt     = __shfl_sync(0xFFFFFFFFu, cnt, __ffs(peers) - 1);
    rank[i] = (unsigned short)(cnt + __popc(below));
    __syncwarp();  // the cleared match word is ordered before the next round's atomics
  }
#else
  // Peers by eight votes: after bit b the mask keeps the lanes whose digit agrees with mine on
  // bits 0..b.  No shared-memory traffic and no atomics for the match; the running count costs
  // one 32-bit load (peers read the same word: broadcast) and one store by the lowest peer.
  u32* const wc = sm.wcnt + warp * kRadixDigits;
#pragma unroll
  for (int i = 0; i < kSortIPT; ++i) {
    u32 const d = (key[i] >> shift) & 0xFFu;
    u32 peers   = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < kRadixBits; ++b) {
      bool const bit = (d >> b) & 1u;
      u32 const m    = __ballot_sync(0xFFFFFFFFu, bit);
      peers &= bit ? m : ~m;
    }
    u32 const below = peers & lt;
    u32 const cnt   = wc[d];
    __syncwarp();
    if (below == 0) wc[d] = cnt + __popc(peers);
    rank[i] = (unsigned short)(cnt + __popc(below));
    __syncwarp();
  }
#endif
  __syncthreads();

  // ---- digit totals over warps (thread d < 256 owns digit d), exclusive scan over digits;
  // the global digit counts are scanned in the same sweep
#if BSJ_SORT_RANK == 0
#define BSJ_WCNT(w, d) sm.whist[(w) * kRadixDigits + (d)].x
#else
#define BSJ_WCNT(w, d) sm.wcnt[(w) * kRadixDigits + (d)]
#endif
  u32 total = 0;
  if (tid < kRadixDigits) {
    u32 sum = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      u32 const c      = BSJ_WCNT(w, tid);
      BSJ_WCNT(w, tid) = sum;
      sum += c;
    }
    total = sum;
  }
  u32 const incl  = warp_inclusive_scan(total);
  u32 const gincl = warp_inclusive_scan(gcount);
  if (lane == 31) {
    sm.warp_sums[warp] = incl;
    sm.gsums[warp]     = gincl;
  }
  __syncthreads();
  // The tile's digit totals are PUBLISHED before the in-tile permutation and the look-back is
  // resolved after it: successors see this tile's aggregate as early as possible and this tile
  // hides its own wait for its predecessors behind the shared-memory placement.
  u32 goffset = 0;  // exclusive global offset of digit `tid`
  if (tid < kRadixDigits) {
    u64* const col = lookback + tid;
    u32 base = 0, gb = 0;
    for (int w = 0; w < warp; ++w) {
      base += sm.warp_sums[w];
      gb += sm.gsums[w];
    }
    u32 const bin_start = base + incl - total;
    goffset             = gb + gincl - gcount;
    sm.bin_start[tid]   = bin_start;

    // padding keys (0xFFFFFFFF) of the last tile land at the end of digit 255: not counted
    u32 my_total = total;
    if (tid == kRadixDigits - 1) my_total -= ((u32)kSortTile - valid);
    sm.gbase[tid] = my_total;  // parked here until the look-back below
    if (tile == 0)
      st_relaxed_u64(col, lb_pack(tag_pre, my_total));
    else
      st_relaxed_u64(col + (u64)tile * kRadixDigits, lb_pack(tag_agg, my_total));
    // fold the digit's tile offset into every warp's offset: one random lookup per element
    // in the placement below instead of two
#pragma unroll
    for (int w = 0; w < kWarps; ++w) BSJ_WCNT(w, tid) += bin_start;
  }
  __syncthreads();

  // ---- place {key, value} at their tile-sorted slot in shared memory
  if (IOTA) {
#pragma unroll
    for (int i = 0; i < kSortIPT; ++i) {
      u32 const d   = (key[i] >> shift) & 0xFFu;
      u32 const pos = BSJ_WCNT(warp, d) + rank[i];
      sm.kv[pos]    = make_uint2(key[i], warp_base + i * 32 + lane);
    }
  } else {
    u32 v[kSortIPT];
    if (bulk) {
      // the values have been sitting in the second half of the slot array since kernel entry;
      // everyone takes its own before anyone overwrites the slots
#pragma unroll
      for (int i = 0; i < kSortIPT; ++i) v[i] = svals[warp * (kSortIPT * 32) + i * 32 + lane];
      __syncthreads();
    } else {
#pragma unroll
      for (int i = 0; i < kSortIPT; ++i) {
        u32 const idx = warp_base + i * 32 + lane;
        v[i]          = idx < n ? ld_stream(vals_in + idx) : 0u;
      }
    }
#pragma unroll
    for (int i = 0; i < kSortIPT; ++i) {
      u32 const d   = (key[i] >> shift) & 0xFFu;
      u32 const pos = BSJ_WCNT(warp, d) + rank[i];
      sm.kv[pos]    = make_uint2(key[i], v[i]);
    }
  }
#undef BSJ_WCNT

  // ---- decoupled look-back on this digit's column of tile descriptors
  if (tid < kRadixDigits) {
    u64* const col     = lookback + tid;
    u32 const my_total = sm.gbase[tid];
    u32 excl           = 0;
    if (tile != 0) {
      i64 t = (i64)tile - 1;
      while (true) {
        u64 const v    = ld_relaxed_u64(col + (u64)t * kRadixDigits);
        u32 const flag = (u32)(v >> 32);
        if (flag == tag_pre) {
          excl += (u32)v;
          break;
        }
        if (flag == tag_agg) {
          excl += (u32)v;
          --t;
        }
      }
      st_relaxed_u64(col + (u64)tile * kRadixDigits, lb_pack(tag_pre, excl + my_total));
    }
    sm.gbase[tid] = goffset + excl - sm.bin_start[tid];
  }
  __syncthreads();

  // ---- scatter: consecutive slots of one digit go to consecutive global addresses
#pragma unroll
  for (int i = 0; i < kSortIPT; ++i) {
    u32 const j = i * kSortBlock + tid;
    if (j < valid) {
      uint2 const e = sm.kv[j];
      u32 const dst = sm.gbase[(e.x >> shift) & 0xFFu] + j;
      keys_out[dst] = e.x;
      vals_out[dst] = e.y;
    }
  }
}

// cudaFuncSetAttribute applies to the current device only: configured once per device
void set_kernel_attrs()
{
  configure_once_per_device(0, [] {
    BSJ_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(sort_smem)));
    BSJ_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(sort_smem)));
  });
}

}  // namespace

void sort_workspace::alloc(u64 n, cudaStream_t s)
{
  num_tiles = (u32)div_up(n, kSortTile);
  hist.alloc(kMaxPasses * kRadixDigits, s);
  tickets.alloc(kMaxPasses, s);
  lookback.alloc((size_t)num_tiles * kRadixDigits, s);
}

void sort_workspace_reset(sort_workspace& ws, cudaStream_t s)
{
  BSJ_CUDA_TRY(cudaMemsetAsync(ws.hist.get(), 0, ws.hist.size() * sizeof(u32), s));
  BSJ_CUDA_TRY(cudaMemsetAsync(ws.tickets.get(), 0, ws.tickets.size() * sizeof(u32), s));
  BSJ_CUDA_TRY(cudaMemsetAsync(ws.lookback.get(), 0, ws.lookback.size() * sizeof(u64), s));
}

void sort_histogram(const u32* keys, u64 n, int begin_bit, int end_bit, sort_workspace& ws,
                    cudaStream_t s)
{
  int const passes = passes_for_bits(begin_bit, end_bit);
  int const grid   = (int)std::min<u64>((u64)num_sms() * 4, (u64)div_up(n, 512));
  histogram_kernel<<<grid, 512, 0, s>>>(keys, n, begin_bit, passes, ws.hist.get());
  BSJ_CHECK_LAUNCH();
}

void sort_passes(u32* keys_a, u32* vals_a, bool iota_values, u32* keys_b, u32* vals_b, u64 n,
                 int begin_bit, int end_bit, sort_workspace& ws, cudaStream_t s,
                 bool* result_in_a, const char* profile_label)
{
  set_kernel_attrs();
  int const passes = passes_for_bits(begin_bit, end_bit);
  bool in_a = true;
  for (int p = 0; p < passes; ++p) {
    u32* kin  = in_a ? keys_a : keys_b;
    u32* vin  = in_a ? vals_a : vals_b;
    u32* kout = in_a ? keys_b : keys_a;
    u32* vout = in_a ? vals_b : vals_a;
    u32 const tag_agg = 2u * (p + 1), tag_pre = 2u * (p + 1) + 1u;
    int const shift   = begin_bit + p * kRadixBits;
    if (p == 0 && iota_values) {
      onesweep_kernel<true><<<ws.num_tiles, kSortBlock, sizeof(sort_smem), s>>>(
        kin, nullptr, kout, vout, (u32)n, shift, ws.hist.get() + p * kRadixDigits,
        ws.lookback.get(), ws.tickets.get() + p, tag_agg, tag_pre);
    } else {
      onesweep_kernel<false><<<ws.num_tiles, kSortBlock, sizeof(sort_smem), s>>>(
        kin, vin, kout, vout, (u32)n, shift, ws.hist.get() + p * kRadixDigits,
        ws.lookback.get(), ws.tickets.get() + p, tag_agg, tag_pre);
    }
    BSJ_CHECK_LAUNCH();
    prof_mark(profile_label);
    in_a = !in_a;
  }
  *result_in_a = in_a;
}

// Three-buffer form: passes ping-pong between the input and a temporary, the last one writes the
// caller's output buffers.  The INPUT buffers are scratch (overwritten when there are >= 3
// passes).  Used where the sorted result must land in freshly allocated output columns while the
// input lives in a buffer the sort may not keep (the multi-GPU receive buffers).
void sort_passes_to(u32* keys_in, u32* vals_in, u32* keys_tmp, u32* vals_tmp, u32* keys_out,
                    u32* vals_out, u64 n, int begin_bit, int end_bit, sort_workspace& ws,
                    cudaStream_t s, const char* profile_label)
{
  set_kernel_attrs();
  int const passes = passes_for_bits(begin_bit, end_bit);
  for (int p = 0; p < passes; ++p) {
    bool const from_in = (p % 2) == 0;                 // X -> Y -> X -> ... , last -> Z
    bool const last    = p == passes - 1;
    u32* kin  = from_in ? keys_in : keys_tmp;
    u32* vin  = from_in ? vals_in : vals_tmp;
    u32* kout = last ? keys_out : (from_in ? keys_tmp : keys_in);
    u32* vout = last ? vals_out : (from_in ? vals_tmp : vals_in);
    u32 const tag_agg = 2u * (p + 1), tag_pre = 2u * (p + 1) + 1u;
    onesweep_kernel<false><<<ws.num_tiles, kSortBlock, sizeof(sort_smem), s>>>(
      kin, vin, kout, vout, (u32)n, begin_bit + p * kRadixBits, ws.hist.get() + p * kRadixDigits,
      ws.lookback.get(), ws.tickets.get() + p, tag_agg, tag_pre);
    BSJ_CHECK_LAUNCH();
    prof_mark(profile_label);
  }
}

}  // namespace bsj

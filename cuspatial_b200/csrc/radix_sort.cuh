// radix_sort.cuh -- hand-written onesweep LSD radix sort of (u32 key, u32 value) pairs.
//
// Replaces thrust::stable_sort_by_key on the reference's hot path
// (cpp/include/cuspatial/detail/index/construction/phase_1.cuh:92,
//  cpp/include/cuspatial/detail/join/quadtree_bbox_filtering.cuh:175-179).
// Stable; 8-bit digits; one histogram read of the keys for all passes; per pass ONE kernel that
// ranks a tile in shared memory, resolves its global digit offsets with a decoupled look-back
// over tile descriptors and scatters through shared memory so that writes leave the SM as
// contiguous per-digit runs.  Algorithmic traffic per element: 4 B (histogram) + 16 B per pass.
#pragma once
#include "common.cuh"

namespace bsj {

constexpr int kRadixBits   = 8;
constexpr int kRadixDigits = 1 << kRadixBits;
constexpr int kMaxPasses   = 4;

#ifndef BSJ_SORT_BLOCK
#define BSJ_SORT_BLOCK 512
#endif
#ifndef BSJ_SORT_IPT
#define BSJ_SORT_IPT 16
#endif
#ifndef BSJ_SORT_MINBLOCKS
#define BSJ_SORT_MINBLOCKS 2
#endif
constexpr int kSortBlock = BSJ_SORT_BLOCK;  // threads per CTA
constexpr int kSortIPT   = BSJ_SORT_IPT;    // keys per thread
constexpr int kSortTile  = kSortBlock * kSortIPT;

// Temporary storage for one sort of n elements.
struct sort_workspace {
  dev_buf<u32> hist;      // [kMaxPasses][256] digit counts, then exclusive offsets
  dev_buf<u64> lookback;  // [tiles][256] descriptors (tag-reused across passes)
  dev_buf<u32> tickets;   // [kMaxPasses] dynamic tile counters
  u32 num_tiles{0};
  void alloc(u64 n, cudaStream_t s);
};

inline int passes_for_bits(int begin_bit, int end_bit)
{
  return (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
}

// Zero hist/tickets/lookback (stream-ordered).
void sort_workspace_reset(sort_workspace& ws, cudaStream_t s);

// Accumulate the digit histograms of `keys` (all passes of [begin_bit, end_bit)) into ws.hist.
// quadtree.cu fuses this into its Morton-encode kernel instead and skips this call.
void sort_histogram(const u32* keys, u64 n, int begin_bit, int end_bit, sort_workspace& ws,
                    cudaStream_t s);

// Run the passes. keys_a/vals_a hold the input (vals_a == nullptr => values are 0..n-1 and are
// never read), keys_b/vals_b are scratch of the same size. Returns in *result_in_a whether the
// sorted data ended in (keys_a, vals_a) (true) or (keys_b, vals_b) (false).
// If vals_a == nullptr the caller must still provide a writable vals_a buffer when more than one
// pass runs (it becomes a ping-pong target).
void sort_passes(u32* keys_a, u32* vals_a, bool iota_values, u32* keys_b, u32* vals_b, u64 n,
                 int begin_bit, int end_bit, sort_workspace& ws, cudaStream_t s,
                 bool* result_in_a, const char* profile_label = "onesweep_pass");

// Same passes with three buffers: input (scratch, overwritten) <-> temporary, last pass into the
// output buffers.  Needs end_bit > begin_bit.
void sort_passes_to(u32* keys_in, u32* vals_in, u32* keys_tmp, u32* vals_tmp, u32* keys_out,
                    u32* vals_out, u64 n, int begin_bit, int end_bit, sort_workspace& ws,
                    cudaStream_t s, const char* profile_label = "onesweep_pass");

}  // namespace bsj

// scan.cuh -- single-pass device-wide exclusive scan (decoupled look-back), u32 values.
// Replaces the thrust::exclusive_scan / inclusive_scan calls on the reference's hot path.
#pragma once
#include "common.cuh"

namespace bsj {

constexpr int kScanBlock = 256;
constexpr int kScanIPT   = 8;
constexpr int kScanTile  = kScanBlock * kScanIPT;

// out[i] = sum_{j<i} in[i]  (u64 accumulation, u64 output); *total = sum of all (device pointer).
// `in == out` aliasing is not allowed (different element sizes).
void exclusive_scan_u32_to_u64(const u32* in, u64* out, u64 n, u64* total, cudaStream_t s);

}  // namespace bsj

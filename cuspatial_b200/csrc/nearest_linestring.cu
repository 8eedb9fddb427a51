// nearest_linestring.cu -- quadtree_point_to_nearest_linestring and linestring_bounding_boxes
// (the same quadtree + bbox filter, a different refinement).
//
// Replaces (rapidsai/cuspatial 25.06):
//   cpp/include/cuspatial/detail/join/quadtree_point_to_nearest_linestring.cuh:150-314
//   cpp/include/cuspatial/detail/algorithm/point_linestring_distance.cuh:33-53
//   cpp/include/cuspatial/detail/bounding_boxes.cuh:96-134 (linestring_bounding_boxes)
//
// The reference enumerates every (point, linestring) candidate of every (linestring, quadrant)
// pair through a "transposed" index mapping (two binary searches + a div/mod per candidate) so
// that one point's candidates are consecutive, evaluates the distance inside a transform
// iterator and runs a CUB reduce_by_key + scatter over all candidates.  Here one warp owns one
// quadrant (a run of consecutive pairs with the same quadrant offset), a lane owns one point of
// that quadrant for the whole run, the segments of each candidate linestring are read once per
// warp (warp-uniform loads, each vertex fetched once and carried to the next segment), and the
// reference's selection rule is folded in registers: no candidate list, no reduce, no scatter.
//
// Exactness: distances use the reference's operation order with the FMA contraction nvcc applies
// to it under its default flags (dot(a,b) = fma(a.x, b.x, a.y*b.y), checked against the SASS of
// the reference build and its golden vectors), IEEE division and square root.  The selection rule
// (:283-300: a zero distance loses to anything, ties go to the smaller linestring id) is
// associative for non-NaN distances, so the in-order fold equals the reference's tree reduction.
// (A point within rounding of a segment can get d0 - r < 0 and hence a NaN distance for that
// linestring; with a NaN candidate the reference's own answer depends on the shape of CUB's
// reduction tree.  The fold here is the sequential reduce_by_key answer, deterministic.)
// Points in no candidate quadrant keep distance 0 like the reference; their two index columns are
// uninitialised memory there and 0 here.
#include "common.cuh"

#include <algorithm>
#include <limits>

namespace bsj {

namespace {

template <typename T>
struct fpl;
template <>
struct fpl<float> {
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ float max() { return 3.402823466e+38f; }
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};
template <>
struct fpl<double> {
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ double max() { return 1.7976931348623157e+308; }
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
};

// vec_2d.hpp:166-170 as nvcc contracts it
template <typename T>
__device__ __forceinline__ T dot2(T ax, T ay, T bx, T by)
{
  return fpl<T>::fma(ax, bx, fpl<T>::mul(ay, by));
}

constexpr int kNlBlock = 128;

template <typename T>
__global__ void __launch_bounds__(kNlBlock, 6)
nearest_linestring_kernel(const u32* __restrict__ pair_line, const u32* __restrict__ pair_quad,
                          u64 n_pairs, const u32* __restrict__ length,
                          const u32* __restrict__ offset, const u32* __restrict__ point_indices,
                          const T* __restrict__ px, const T* __restrict__ py, u64 n_points,
                          const u32* __restrict__ line_offsets, const T* __restrict__ lx,
                          const T* __restrict__ ly, u32* __restrict__ out_point,
                          u32* __restrict__ out_line, T* __restrict__ out_dist)
{
  u32 const lane    = lane_id();
  u64 const warp    = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  u64 const n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 j = warp; j < n_pairs; j += n_warps) {
    u32 const quad = __ldg(pair_quad + j);
    u32 const qoff = __ldg(offset + quad);
    // only the first pair of a run of equal quadrant offsets works (:44-60)
    if (j > 0 && __ldg(offset + __ldg(pair_quad + j - 1)) == qoff) continue;
    u32 nl = 1;
    while (j + nl < n_pairs && __ldg(offset + __ldg(pair_quad + j + nl)) == qoff) ++nl;
    u32 const np = __ldg(length + quad);

    for (u32 base = 0; base < np; base += 32) {
      u32 const i       = base + lane;
      bool const active = i < np && (u64)qoff + i < n_points;
      u32 const pos     = qoff + i;
      T x = 0, y = 0;
      if (active) {
        u32 const pid = __ldg(point_indices + pos);
        x             = __ldg(px + pid);
        y             = __ldg(py + pid);
      }
      T best_d      = 0;
      u32 best_line = 0;
      for (u32 k = 0; k < nl; ++k) {
        u32 const line = __ldg(pair_line + j + k);
        u32 const v0 = __ldg(line_offsets + line), v1 = __ldg(line_offsets + line + 1);
        T dsq = fpl<T>::max();
        if (v1 > v0 + 1) {
          T ax = __ldg(lx + v0), ay = __ldg(ly + v0);
          T v1px = fpl<T>::sub(x, ax), v1py = fpl<T>::sub(y, ay);
          T d0 = dot2<T>(v1px, v1py, v1px, v1py);
          for (u32 s = v0 + 1; s < v1; ++s) {
            T const bx = __ldg(lx + s), by = __ldg(ly + s);
            T const v2px = fpl<T>::sub(x, bx), v2py = fpl<T>::sub(y, by);
            T const d1 = dot2<T>(v2px, v2py, v2px, v2py);
            T const ex = fpl<T>::sub(bx, ax), ey = fpl<T>::sub(by, ay);
            T const d2 = dot2<T>(ex, ey, ex, ey);
            T const d3 = dot2<T>(v1px, v1py, ex, ey);
            // reference: r = d3*d3/d2; d = (d3 <= 0 || r >= d2) ? min(d0, d1) : d0 - r.
            // The IEEE division is the most expensive operation here and its result only
            // matters when the projection falls inside the segment: d3 <= 0 decides without it,
            // and d3*d3 > d2*d2 (with a margin far above the rounding of the two products)
            // implies r >= d2 -- rounding is monotone and d2 is representable.  NaN, overflow
            // and underflow fail the comparison and take the literal path.
            T d;
            if (d3 <= (T)0) {
              d = fmin(d0, d1);
            } else {
              T const m = fpl<T>::mul(d3, d3);
              if (m > fpl<T>::mul(fpl<T>::mul(d2, d2), (T)1.00001)) {
                d = fmin(d0, d1);
              } else {
                T const r = fpl<T>::div(m, d2);
                d         = (r >= d2) ? fmin(d0, d1) : fpl<T>::sub(d0, r);
              }
            }
            dsq = fmin(dsq, d);
            ax = bx; ay = by; v1px = v2px; v1py = v2py; d0 = d1;
          }
        }
        T const d = fpl<T>::sqrt(dsq);
        if (k == 0) {
          best_d    = d;
          best_line = line;
        } else if (best_d == (T)0) {  // :288-289 zero on the left: take the right
          best_d    = d;
          best_line = line;
        } else if (d == (T)0) {       // :290-291 zero on the right: keep the left
        } else if (best_d == d) {     // :293-297 tie: smaller linestring id
          if (!(best_line < line)) best_line = line;
        } else if (!(best_d < d)) {   // :299
          best_d    = d;
          best_line = line;
        }
      }
      if (active) {
        out_point[pos] = pos;
        out_line[pos]  = best_line;
        out_dist[pos]  = best_d;
      }
    }
  }
}

// detail/bounding_boxes.cuh:36-60,96-134: per-linestring min/max of (v - r, v + r); warp per line
template <typename T>
__global__ void __launch_bounds__(128)
line_bbox_kernel(const u32* __restrict__ line_offsets, u32 n_lines, const T* __restrict__ lx,
                 const T* __restrict__ ly, u32 n_verts, T r, T* __restrict__ ox0,
                 T* __restrict__ oy0, T* __restrict__ ox1, T* __restrict__ oy1)
{
  u32 const p    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  u32 const lane = lane_id();
  if (p >= n_lines) return;
  u32 const v0 = line_offsets[p];
  u32 const v1 = min(line_offsets[p + 1], n_verts);
  T xmin = fpl<T>::inf(), ymin = fpl<T>::inf(), xmax = -fpl<T>::inf(), ymax = -fpl<T>::inf();
  for (u32 i = v0 + lane; i < v1; i += 32) {
    T const ax = lx[i], ay = ly[i];
    xmin = fmin(xmin, fpl<T>::sub(ax, r)); ymin = fmin(ymin, fpl<T>::sub(ay, r));
    xmax = fmax(xmax, ax + r);             ymax = fmax(ymax, ay + r);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if (lane == 0) {
    ox0[p] = xmin; oy0[p] = ymin; ox1[p] = xmax; oy1[p] = ymax;
  }
}

template <typename T>
void nearest_t(const u32* pair_line, const u32* pair_quad, u64 n_pairs, const u32* length,
               const u32* offset, const u32* point_indices, const void* px, const void* py,
               u64 n_points, const u32* line_offsets, const void* lx, const void* ly,
               cudaStream_t s, u32* out_point, u32* out_line, void* out_dist)
{
  stage_timer tm(s);
  // :264 distances start at zero; the index columns are uninitialised in the reference, 0 here
  BSJ_CUDA_TRY(cudaMemsetAsync(out_point, 0, n_points * sizeof(u32), s));
  BSJ_CUDA_TRY(cudaMemsetAsync(out_line, 0, n_points * sizeof(u32), s));
  BSJ_CUDA_TRY(cudaMemsetAsync(out_dist, 0, n_points * sizeof(T), s));
  int const grid = (int)std::min<u64>((u64)num_sms() * 16, div_up(n_pairs * 32, (u64)kNlBlock));
  nearest_linestring_kernel<T><<<std::max(grid, 1), kNlBlock, 0, s>>>(
    pair_line, pair_quad, n_pairs, length, offset, point_indices, (const T*)px, (const T*)py,
    n_points, line_offsets, (const T*)lx, (const T*)ly, out_point, out_line, (T*)out_dist);
  BSJ_CHECK_LAUNCH();
  tm.mark("nearest_linestring");
  tm.finish();  // results are ready in stream order; no trailing host synchronisation
}

}  // namespace

void quadtree_point_to_nearest_linestring_impl(
  const u32* pair_line, const u32* pair_quad, u64 n_pairs, const u32* length, const u32* offset,
  u64 num_nodes, const u32* point_indices, const void* px, const void* py, int dtype, u64 n_points,
  const u32* line_offsets, u64 n_line_offsets, const void* lx, const void* ly, u64 n_verts,
  cudaStream_t s, u32* out_point, u32* out_line, void* out_dist, u64* out_rows)
{
  (void)n_verts;
  *out_rows = 0;
  // cpp/src/join/quadtree_point_to_nearest_linestring.cu:176-184: empty in, empty table out
  if (n_pairs == 0 || num_nodes == 0 || n_points == 0 || n_line_offsets == 0) return;
  BSJ_EXPECTS(out_point && out_line && out_dist, "output columns must not be NULL");
  if (dtype == BSJ_FLOAT32)
    nearest_t<float>(pair_line, pair_quad, n_pairs, length, offset, point_indices, px, py,
                     n_points, line_offsets, lx, ly, s, out_point, out_line, out_dist);
  else
    nearest_t<double>(pair_line, pair_quad, n_pairs, length, offset, point_indices, px, py,
                      n_points, line_offsets, lx, ly, s, out_point, out_line, out_dist);
  *out_rows = n_points;
}

void linestring_bounding_boxes_impl(const u32* line_offsets, u64 n_line_offsets, const void* lx,
                                    const void* ly, int dtype, u64 n_verts, double r,
                                    cudaStream_t s, void* x0, void* y0, void* x1, void* y1)
{
  if (n_line_offsets < 2 || n_verts == 0) return;
  u32 const n_lines = (u32)(n_line_offsets - 1);
  if (dtype == BSJ_FLOAT32)
    line_bbox_kernel<float><<<div_up((u64)n_lines * 32, 128), 128, 0, s>>>(
      line_offsets, n_lines, (const float*)lx, (const float*)ly, (u32)n_verts, (float)r,
      (float*)x0, (float*)y0, (float*)x1, (float*)y1);
  else
    line_bbox_kernel<double><<<div_up((u64)n_lines * 32, 128), 128, 0, s>>>(
      line_offsets, n_lines, (const double*)lx, (const double*)ly, (u32)n_verts, r, (double*)x0,
      (double*)y0, (double*)x1, (double*)y1);
  BSJ_CHECK_LAUNCH();
}

}  // namespace bsj

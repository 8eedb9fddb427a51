// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Thin C-ABI driver around the UNMODIFIED header-only implementation of the reference
// (rapidsai/cuspatial 25.06, headers read in place from /root/reference/cpp/include).
// It plays the role of the reference's cuDF column layer, which cannot be built offline:
//   - quadtree_on_points        : cpp/src/indexing/point_quadtree.cu:64-119
//   - join_quadtree_and_bboxes  : cpp/src/join/quadtree_bbox_filtering.cu:40-80
//   - quadtree_point_in_polygon : cpp/src/join/quadtree_point_in_polygon.cu:45-120
//   - point_in_polygon (bitmask): cpp/src/point_in_polygon/point_in_polygon.cu:52-98
//   - pairwise_point_in_polygon : cpp/src/point_in_polygon/point_in_polygon.cu:52-98,172-190
//   - polygon_bounding_boxes    : cpp/src/bounding_boxes/polygon_bounding_boxes.cu:132-161
//   - linestring_bounding_boxes : cpp/src/bounding_boxes/linestring_bounding_boxes.cu:60-120
//   - quadtree_point_to_nearest_linestring : cpp/src/join/quadtree_point_to_nearest_linestring.cu:47-112
// i.e. cast the double parameters to T, wrap raw columns into the reference's iterators,
// call the header-only entry point, hand the result arrays back.
//
// Two build flavours (oracle/Makefile):
//   host : g++ -x c++  -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_OMP  (Thrust's OpenMP
//          backend runs the reference's own algorithms on the host cores; all pointers
//          are host pointers).  -> oracle/_ref/libcuspatial_ref_host.so
//   cuda : nvcc -x cu  -arch=sm_100a  (the reference's Thrust/CUB kernels on the GPU; all
//          data pointers are device pointers).       -> oracle/_ref/libcuspatial_ref_cuda.so
//
// FP note for the host flavour: it is compiled with -mfma -ffp-contract=fast so that gcc
// contracts `v_min + (k+1)*level_scale` (detail/join/intersection.cuh:113-119) into an FMA
// exactly where nvcc does under the reference's default flags; the PIP predicate has no
// mul+add shape, so nothing contracts there (same as the reference SASS: DMUL/DADD only).

#include <thrust/execution_policy.h>

#include <cuspatial/bounding_boxes.cuh>
#include <cuspatial/geometry/box.hpp>
#include <cuspatial/geometry/vec_2d.hpp>
#include <cuspatial/iterator_factory.cuh>
#include <cuspatial/point_in_polygon.cuh>
#include <cuspatial/point_quadtree.cuh>
#include <cuspatial/range/multilinestring_range.cuh>
#include <cuspatial/range/multipoint_range.cuh>
#include <cuspatial/range/multipolygon_range.cuh>

#include <cuspatial/spatial_join.cuh>

#include <thrust/iterator/counting_iterator.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace {

thread_local std::string g_err;

#if defined(__CUDACC__)
constexpr bool kCuda = true;
#else
constexpr bool kCuda = false;
#endif

// Hand a result uvector's allocation to the caller, who releases it with ref_free().
template <typename T>
void* release(rmm::device_uvector<T>& v, uint64_t* n)
{
  *n = v.size();
  if (v.size() == 0) return nullptr;
  return v.shim_release();
}

// CUDA flavour: keep freed blocks in the stream-ordered pool between calls, as the RMM pool
// resource the reference normally runs on does (otherwise every call pays cudaMalloc/cudaFree).
void init_once()
{
#if defined(__CUDACC__)
  static bool done = false;
  if (!done) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done = true;
  }
#endif
}

template <typename T>
int quadtree_on_points_t(T const* x, T const* y, uint64_t n, double x_min, double x_max,
                         double y_min, double y_max, double scale, int8_t max_depth,
                         int32_t max_size, void** out, uint64_t* out_n)
{
  rmm::cuda_stream_view stream{};
  auto points        = cuspatial::make_vec_2d_iterator(x, y);
  auto [idx, tree]   = cuspatial::quadtree_on_points(
    points, points + n,
    cuspatial::vec_2d<T>{static_cast<T>(x_min), static_cast<T>(y_min)},
    cuspatial::vec_2d<T>{static_cast<T>(x_max), static_cast<T>(y_max)},
    static_cast<T>(scale), max_depth, max_size, stream, rmm::mr::get_current_device_resource());
  stream.synchronize();
  uint64_t q = 0;
  out[0] = release(idx, &out_n[0]);
  out[1] = release(tree.key, &q);
  out[2] = release(tree.level, &q);
  out[3] = release(tree.internal_node_flag, &q);
  out[4] = release(tree.length, &q);
  out[5] = release(tree.offset, &q);
  out_n[1] = q;
  return 0;
}

template <typename T>
int join_t(uint32_t const* key, uint8_t const* level, bool const* internal, uint32_t const* length,
           uint32_t const* offset, uint64_t q, T const* bx0, T const* by0, T const* bx1,
           T const* by1, uint64_t n_boxes, double x_min, double y_min, double scale,
           int8_t max_depth, void** out, uint64_t* out_n)
{
  rmm::cuda_stream_view stream{};
  auto bbox_min = cuspatial::make_vec_2d_iterator(bx0, by0);
  auto bbox_max = cuspatial::make_vec_2d_iterator(bx1, by1);
  auto bbox_itr = cuspatial::make_box_iterator(bbox_min, bbox_max);
  cuspatial::point_quadtree_ref tree(key, key + q, level, internal, length, offset);
  auto [bbox_offset, quad_offset] = cuspatial::join_quadtree_and_bounding_boxes(
    tree, bbox_itr, bbox_itr + n_boxes,
    cuspatial::vec_2d<T>{static_cast<T>(x_min), static_cast<T>(y_min)}, static_cast<T>(scale),
    max_depth, stream, rmm::mr::get_current_device_resource());
  stream.synchronize();
  out[0] = release(bbox_offset, &out_n[0]);
  out[1] = release(quad_offset, &out_n[0]);
  return 0;
}

template <typename T>
int qpip_t(uint32_t const* pair_poly, uint32_t const* pair_quad, uint64_t n_pairs,
           uint32_t const* key, uint8_t const* level, bool const* internal, uint32_t const* length,
           uint32_t const* offset, uint64_t q, uint32_t const* point_indices, T const* px,
           T const* py, uint64_t n_points, uint32_t const* poly_offsets, uint64_t n_poly_offsets,
           uint32_t const* ring_offsets, uint64_t n_ring_offsets, T const* vx, T const* vy,
           uint64_t n_verts, void** out, uint64_t* out_n)
{
  rmm::cuda_stream_view stream{};
  cuspatial::point_quadtree_ref tree(key, key + q, level, internal, length, offset);
  // same construction as cpp/src/join/quadtree_point_in_polygon.cu:68-76
  auto multipolygons = cuspatial::multipolygon_range(
    thrust::make_counting_iterator(0), thrust::make_counting_iterator((int)n_poly_offsets),
    poly_offsets, poly_offsets + n_poly_offsets, ring_offsets, ring_offsets + n_ring_offsets,
    cuspatial::make_vec_2d_iterator(vx, vy), cuspatial::make_vec_2d_iterator(vx + n_verts, vy + n_verts));
  auto [poly_idx, point_idx] = cuspatial::quadtree_point_in_polygon(
    pair_poly, pair_poly + n_pairs, pair_quad, tree, point_indices, point_indices + n_points,
    cuspatial::make_vec_2d_iterator(px, py), multipolygons, stream,
    rmm::mr::get_current_device_resource());
  stream.synchronize();
  out[0] = release(poly_idx, &out_n[0]);
  out[1] = release(point_idx, &out_n[0]);
  return 0;
}

template <typename T>
int pip_t(T const* px, T const* py, uint64_t n_points, int32_t const* poly_offsets,
          uint64_t n_poly_offsets, int32_t const* ring_offsets, uint64_t n_ring_offsets,
          T const* vx, T const* vy, uint64_t n_verts, int32_t* out_mask)
{
  rmm::cuda_stream_view stream{};
  // same construction as cpp/src/point_in_polygon/point_in_polygon.cu:71-89
  auto points_begin = cuspatial::make_vec_2d_iterator(px, py);
  auto multipoints  = cuspatial::make_multipoint_range(
    n_points, thrust::make_counting_iterator(0), n_points, points_begin);
  auto polygon_size = n_poly_offsets - 1;
  auto multipolygons = cuspatial::make_multipolygon_range(
    polygon_size, thrust::make_counting_iterator(0), polygon_size, poly_offsets,
    n_ring_offsets - 1, ring_offsets, n_verts, cuspatial::make_vec_2d_iterator(vx, vy));
  cuspatial::point_in_polygon(multipoints, multipolygons, out_mask, stream);
  stream.synchronize();
  return 0;
}

template <typename T>
int pairwise_pip_t(T const* px, T const* py, uint64_t n_points, int32_t const* poly_offsets,
                   uint64_t n_poly_offsets, int32_t const* ring_offsets, uint64_t n_ring_offsets,
                   T const* vx, T const* vy, uint64_t n_verts, uint8_t* out)
{
  rmm::cuda_stream_view stream{};
  // cpp/src/point_in_polygon/point_in_polygon.cu:122-125 (column-layer check)
  if (n_points != (n_poly_offsets ? n_poly_offsets - 1 : 0))
    throw std::logic_error("Must pass in the same number of points as polygons.");
  // same construction as cpp/src/point_in_polygon/point_in_polygon.cu:71-93
  auto points_begin = cuspatial::make_vec_2d_iterator(px, py);
  auto multipoints  = cuspatial::make_multipoint_range(
    n_points, thrust::make_counting_iterator(0), n_points, points_begin);
  auto polygon_size = n_poly_offsets - 1;
  auto multipolygons = cuspatial::make_multipolygon_range(
    polygon_size, thrust::make_counting_iterator(0), polygon_size, poly_offsets,
    n_ring_offsets - 1, ring_offsets, n_verts, cuspatial::make_vec_2d_iterator(vx, vy));
  cuspatial::pairwise_point_in_polygon(multipoints, multipolygons, out, stream);
  stream.synchronize();
  return 0;
}

template <typename T>
int poly_bbox_t(uint32_t const* poly_offsets, uint64_t n_poly_offsets, uint32_t const* ring_offsets,
                uint64_t n_ring_offsets, T const* vx, T const* vy, uint64_t n_verts, T expansion,
                T* x0, T* y0, T* x1, T* y1)
{
  rmm::cuda_stream_view stream{};
  uint64_t n_poly = n_poly_offsets - 1;
  rmm::device_uvector<cuspatial::box<T>> boxes(n_poly, stream);
  auto pts = cuspatial::make_vec_2d_iterator(vx, vy);
  cuspatial::polygon_bounding_boxes(poly_offsets, poly_offsets + n_poly_offsets, ring_offsets,
                                    ring_offsets + n_ring_offsets, pts, pts + n_verts,
                                    boxes.begin(), expansion, stream);
  stream.synchronize();
  std::vector<cuspatial::box<T>> h(n_poly);
#if defined(__CUDACC__)
  cudaMemcpy(h.data(), boxes.data(), n_poly * sizeof(cuspatial::box<T>), cudaMemcpyDeviceToHost);
#else
  std::memcpy(h.data(), boxes.data(), n_poly * sizeof(cuspatial::box<T>));
#endif
  // bbox outputs are always HOST arrays (tiny), in both flavours
  for (uint64_t i = 0; i < n_poly; ++i) {
    x0[i] = h[i].v1.x; y0[i] = h[i].v1.y; x1[i] = h[i].v2.x; y1[i] = h[i].v2.y;
  }
  return 0;
}

template <typename T>
int nearest_linestring_t(uint32_t const* pair_line, uint32_t const* pair_quad, uint64_t n_pairs,
                         uint32_t const* key, uint8_t const* level, bool const* internal,
                         uint32_t const* length, uint32_t const* offset, uint64_t q,
                         uint32_t const* point_indices, T const* px, T const* py,
                         uint64_t n_points, uint32_t const* line_offsets, uint64_t n_line_offsets,
                         T const* lx, T const* ly, uint64_t n_verts, void** out, uint64_t* out_n)
{
  rmm::cuda_stream_view stream{};
  cuspatial::point_quadtree_ref tree(key, key + q, level, internal, length, offset);
  // same construction as cpp/src/join/quadtree_point_to_nearest_linestring.cu:70-76
  auto linestrings = cuspatial::multilinestring_range(
    thrust::make_counting_iterator(0), thrust::make_counting_iterator((int)n_line_offsets),
    line_offsets, line_offsets + n_line_offsets, cuspatial::make_vec_2d_iterator(lx, ly),
    cuspatial::make_vec_2d_iterator(lx + n_verts, ly + n_verts));
  auto [point_idx, line_idx, dist] = cuspatial::quadtree_point_to_nearest_linestring(
    pair_line, pair_line + n_pairs, pair_quad, tree, point_indices, point_indices + n_points,
    cuspatial::make_vec_2d_iterator(px, py), linestrings, stream,
    rmm::mr::get_current_device_resource());
  stream.synchronize();
  out[0] = release(point_idx, &out_n[0]);
  out[1] = release(line_idx, &out_n[0]);
  out[2] = release(dist, &out_n[0]);
  return 0;
}

template <typename T>
int line_bbox_t(uint32_t const* line_offsets, uint64_t n_line_offsets, T const* lx, T const* ly,
                uint64_t n_verts, T expansion, T* x0, T* y0, T* x1, T* y1)
{
  rmm::cuda_stream_view stream{};
  uint64_t n_lines = n_line_offsets - 1;
  rmm::device_uvector<cuspatial::box<T>> boxes(n_lines, stream);
  auto pts = cuspatial::make_vec_2d_iterator(lx, ly);
  cuspatial::linestring_bounding_boxes(line_offsets, line_offsets + n_line_offsets, pts,
                                       pts + n_verts, boxes.begin(), expansion, stream);
  stream.synchronize();
  std::vector<cuspatial::box<T>> h(n_lines);
#if defined(__CUDACC__)
  cudaMemcpy(h.data(), boxes.data(), n_lines * sizeof(cuspatial::box<T>), cudaMemcpyDeviceToHost);
#else
  std::memcpy(h.data(), boxes.data(), n_lines * sizeof(cuspatial::box<T>));
#endif
  for (uint64_t i = 0; i < n_lines; ++i) {
    x0[i] = h[i].v1.x; y0[i] = h[i].v1.y; x1[i] = h[i].v2.x; y1[i] = h[i].v2.y;
  }
  return 0;
}

template <typename F>
int guarded(F&& f)
{
  try {
    init_once();
    return f();
  } catch (std::exception const& e) {
    g_err = e.what();
    return 1;
  }
}

}  // namespace

extern "C" {

int ref_is_cuda() { return kCuda ? 1 : 0; }
const char* ref_last_error() { return g_err.c_str(); }

void ref_free(void* p)
{
#if defined(__CUDACC__)
  cudaFreeAsync(p, 0);
#else
  std::free(p);
#endif
}

// out[6] = {point_indices, key, level, is_internal, length, offset}; out_n[2] = {n, q}
int ref_quadtree_on_points(void const* x, void const* y, int dtype, uint64_t n, double x_min,
                           double x_max, double y_min, double y_max, double scale, int max_depth,
                           int max_size, void** out, uint64_t* out_n)
{
  return guarded([&] {
    return dtype == 0
             ? quadtree_on_points_t<float>((float const*)x, (float const*)y, n, x_min, x_max, y_min,
                                           y_max, scale, (int8_t)max_depth, max_size, out, out_n)
             : quadtree_on_points_t<double>((double const*)x, (double const*)y, n, x_min, x_max,
                                            y_min, y_max, scale, (int8_t)max_depth, max_size, out,
                                            out_n);
  });
}

// out[2] = {bbox_offset, quad_offset}; out_n[1] = {p}
int ref_join_quadtree_and_bounding_boxes(uint32_t const* key, uint8_t const* level,
                                         uint8_t const* internal, uint32_t const* length,
                                         uint32_t const* offset, uint64_t q, void const* bx0,
                                         void const* by0, void const* bx1, void const* by1,
                                         int dtype, uint64_t n_boxes, double x_min, double y_min,
                                         double scale, int max_depth, void** out, uint64_t* out_n)
{
  return guarded([&] {
    return dtype == 0 ? join_t<float>(key, level, (bool const*)internal, length, offset, q,
                                      (float const*)bx0, (float const*)by0, (float const*)bx1,
                                      (float const*)by1, n_boxes, x_min, y_min, scale,
                                      (int8_t)max_depth, out, out_n)
                      : join_t<double>(key, level, (bool const*)internal, length, offset, q,
                                       (double const*)bx0, (double const*)by0, (double const*)bx1,
                                       (double const*)by1, n_boxes, x_min, y_min, scale,
                                       (int8_t)max_depth, out, out_n);
  });
}

// out[2] = {polygon_index, point_index}; out_n[1] = {h}
int ref_quadtree_point_in_polygon(uint32_t const* pair_poly, uint32_t const* pair_quad,
                                  uint64_t n_pairs, uint32_t const* key, uint8_t const* level,
                                  uint8_t const* internal, uint32_t const* length,
                                  uint32_t const* offset, uint64_t q,
                                  uint32_t const* point_indices, void const* px, void const* py,
                                  int dtype, uint64_t n_points, uint32_t const* poly_offsets,
                                  uint64_t n_poly_offsets, uint32_t const* ring_offsets,
                                  uint64_t n_ring_offsets, void const* vx, void const* vy,
                                  uint64_t n_verts, void** out, uint64_t* out_n)
{
  return guarded([&] {
    return dtype == 0
             ? qpip_t<float>(pair_poly, pair_quad, n_pairs, key, level, (bool const*)internal,
                             length, offset, q, point_indices, (float const*)px, (float const*)py,
                             n_points, poly_offsets, n_poly_offsets, ring_offsets, n_ring_offsets,
                             (float const*)vx, (float const*)vy, n_verts, out, out_n)
             : qpip_t<double>(pair_poly, pair_quad, n_pairs, key, level, (bool const*)internal,
                              length, offset, q, point_indices, (double const*)px,
                              (double const*)py, n_points, poly_offsets, n_poly_offsets,
                              ring_offsets, n_ring_offsets, (double const*)vx, (double const*)vy,
                              n_verts, out, out_n);
  });
}

int ref_point_in_polygon(void const* px, void const* py, int dtype, uint64_t n_points,
                         int32_t const* poly_offsets, uint64_t n_poly_offsets,
                         int32_t const* ring_offsets, uint64_t n_ring_offsets, void const* vx,
                         void const* vy, uint64_t n_verts, int32_t* out_mask)
{
  return guarded([&] {
    return dtype == 0
             ? pip_t<float>((float const*)px, (float const*)py, n_points, poly_offsets,
                            n_poly_offsets, ring_offsets, n_ring_offsets, (float const*)vx,
                            (float const*)vy, n_verts, out_mask)
             : pip_t<double>((double const*)px, (double const*)py, n_points, poly_offsets,
                             n_poly_offsets, ring_offsets, n_ring_offsets, (double const*)vx,
                             (double const*)vy, n_verts, out_mask);
  });
}

int ref_pairwise_point_in_polygon(void const* px, void const* py, int dtype, uint64_t n_points,
                                  int32_t const* poly_offsets, uint64_t n_poly_offsets,
                                  int32_t const* ring_offsets, uint64_t n_ring_offsets,
                                  void const* vx, void const* vy, uint64_t n_verts, uint8_t* out)
{
  return guarded([&] {
    if (n_points == 0 && n_poly_offsets <= 1) return 0;
    return dtype == 0
             ? pairwise_pip_t<float>((float const*)px, (float const*)py, n_points, poly_offsets,
                                     n_poly_offsets, ring_offsets, n_ring_offsets,
                                     (float const*)vx, (float const*)vy, n_verts, out)
             : pairwise_pip_t<double>((double const*)px, (double const*)py, n_points, poly_offsets,
                                      n_poly_offsets, ring_offsets, n_ring_offsets,
                                      (double const*)vx, (double const*)vy, n_verts, out);
  });
}

// out[3] = {point_index u32, linestring_index u32, distance T}; out_n[1] = {n_points}
int ref_quadtree_point_to_nearest_linestring(
  uint32_t const* pair_line, uint32_t const* pair_quad, uint64_t n_pairs, uint32_t const* key,
  uint8_t const* level, uint8_t const* internal, uint32_t const* length, uint32_t const* offset,
  uint64_t q, uint32_t const* point_indices, void const* px, void const* py, int dtype,
  uint64_t n_points, uint32_t const* line_offsets, uint64_t n_line_offsets, void const* lx,
  void const* ly, uint64_t n_verts, void** out, uint64_t* out_n)
{
  return guarded([&] {
    return dtype == 0
             ? nearest_linestring_t<float>(pair_line, pair_quad, n_pairs, key, level,
                                           (bool const*)internal, length, offset, q, point_indices,
                                           (float const*)px, (float const*)py, n_points,
                                           line_offsets, n_line_offsets, (float const*)lx,
                                           (float const*)ly, n_verts, out, out_n)
             : nearest_linestring_t<double>(pair_line, pair_quad, n_pairs, key, level,
                                            (bool const*)internal, length, offset, q,
                                            point_indices, (double const*)px, (double const*)py,
                                            n_points, line_offsets, n_line_offsets,
                                            (double const*)lx, (double const*)ly, n_verts, out,
                                            out_n);
  });
}

// x0..y1 are HOST output arrays of n_line_offsets-1 elements of T.
int ref_linestring_bounding_boxes(uint32_t const* line_offsets, uint64_t n_line_offsets,
                                  void const* lx, void const* ly, int dtype, uint64_t n_verts,
                                  double expansion, void* x0, void* y0, void* x1, void* y1)
{
  return guarded([&] {
    return dtype == 0 ? line_bbox_t<float>(line_offsets, n_line_offsets, (float const*)lx,
                                           (float const*)ly, n_verts, (float)expansion,
                                           (float*)x0, (float*)y0, (float*)x1, (float*)y1)
                      : line_bbox_t<double>(line_offsets, n_line_offsets, (double const*)lx,
                                            (double const*)ly, n_verts, expansion, (double*)x0,
                                            (double*)y0, (double*)x1, (double*)y1);
  });
}

// x0..y1 are HOST output arrays of n_poly_offsets-1 elements of T.
int ref_polygon_bounding_boxes(uint32_t const* poly_offsets, uint64_t n_poly_offsets,
                               uint32_t const* ring_offsets, uint64_t n_ring_offsets,
                               void const* vx, void const* vy, int dtype, uint64_t n_verts,
                               double expansion, void* x0, void* y0, void* x1, void* y1)
{
  return guarded([&] {
    return dtype == 0
             ? poly_bbox_t<float>(poly_offsets, n_poly_offsets, ring_offsets, n_ring_offsets,
                                  (float const*)vx, (float const*)vy, n_verts, (float)expansion,
                                  (float*)x0, (float*)y0, (float*)x1, (float*)y1)
             : poly_bbox_t<double>(poly_offsets, n_poly_offsets, ring_offsets, n_ring_offsets,
                                   (double const*)vx, (double const*)vy, n_verts, expansion,
                                   (double*)x0, (double*)y0, (double*)x1, (double*)y1);
  });
}

}  // extern "C"

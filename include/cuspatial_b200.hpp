// cuspatial_b200.hpp -- C++ drop-in over the C ABI: the reference's function names, parameter
// order and output column order/dtypes for the quadtree point-in-polygon path, on raw device
// columns instead of cudf::column_view (cuDF is not required on this path).
//
//   cuspatial::quadtree_on_points              cpp/include/cuspatial/point_quadtree.hpp:68-78
//   cuspatial::join_quadtree_and_bounding_boxes  cpp/include/cuspatial/spatial_join.hpp:66-75
//   cuspatial::quadtree_point_in_polygon         cpp/include/cuspatial/spatial_join.hpp:116-126
//   cuspatial::point_in_polygon                  cpp/include/cuspatial/point_in_polygon.hpp:75-82
//   cuspatial::pairwise_point_in_polygon         cpp/include/cuspatial/point_in_polygon.hpp:124-131
//   cuspatial::quadtree_point_to_nearest_linestring  cpp/include/cuspatial/spatial_join.hpp:166-175
//   cuspatial::linestring_bounding_boxes         cpp/include/cuspatial/bounding_boxes.hpp:52-57
//
// Header only; link against libcuspatial_b200.so.  Errors are rethrown as the reference does:
// std::logic_error for CUSPATIAL_EXPECTS conditions (cuspatial::logic_error derives from it,
// cpp/include/cuspatial/error.hpp:40-43), std::runtime_error for CUDA failures, std::bad_alloc.
#pragma once

#include "cuspatial_b200.h"

#include <cstdint>
#include <new>
#include <stdexcept>
#include <string>
#include <utility>

namespace cuspatial_b200 {

/// Non-owning typed device column (the analogue of cudf::column_view for this path).
template <typename T>
struct column_view {
  const T* data{nullptr};
  uint64_t size{0};
};

/// Owning device column released with bsj_free (the analogue of std::unique_ptr<cudf::column>).
template <typename T>
class column {
 public:
  column() = default;
  column(T* p, uint64_t n, bsj_stream_t s) : p_(p), n_(n), s_(s) {}
  column(column&& o) noexcept : p_(o.p_), n_(o.n_), s_(o.s_) { o.p_ = nullptr; o.n_ = 0; }
  column& operator=(column&& o) noexcept
  {
    if (this != &o) {
      reset();
      p_ = o.p_; n_ = o.n_; s_ = o.s_;
      o.p_ = nullptr; o.n_ = 0;
    }
    return *this;
  }
  column(const column&)            = delete;
  column& operator=(const column&) = delete;
  ~column() { reset(); }
  T* data() const { return p_; }
  uint64_t size() const { return n_; }
  column_view<T> view() const { return {p_, n_}; }

 private:
  void reset()
  {
    if (p_) bsj_free(p_, s_);
    p_ = nullptr;
    n_ = 0;
  }
  T* p_{nullptr};
  uint64_t n_{0};
  bsj_stream_t s_{nullptr};
};

/// The reference's quadtree table: key, level, is_internal_node, length, offset.
struct quadtree_table {
  column<uint32_t> key;
  column<uint8_t> level;
  column<uint8_t> is_internal_node;  // BOOL8
  column<uint32_t> length;
  column<uint32_t> offset;
  bsj_grid grid{};  // optional acceleration hint for quadtree_point_in_polygon
  column<uint32_t> sorted_keys;  // backing store of grid.sorted_keys
};

struct pair_table {  // (bbox_offset, quad_offset) or (polygon_index, point_index)
  column<uint32_t> first;
  column<uint32_t> second;
};

namespace detail {
inline void check(int rc)
{
  if (rc == BSJ_SUCCESS) return;
  std::string msg = bsj_last_error();
  if (rc == BSJ_INVALID_ARGUMENT) throw std::logic_error(msg);
  if (rc == BSJ_OUT_OF_MEMORY) throw std::bad_alloc();
  throw std::runtime_error(msg);
}
template <typename T>
constexpr int dtype_of()
{
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "float or double");
  return sizeof(T) == 4 ? BSJ_FLOAT32 : BSJ_FLOAT64;
}
}  // namespace detail

/// cuspatial::quadtree_on_points(x, y, x_min, x_max, y_min, y_max, scale, max_depth, max_size, mr)
template <typename T>
std::pair<column<uint32_t>, quadtree_table> quadtree_on_points(column_view<T> x, column_view<T> y,
                                                               double x_min, double x_max,
                                                               double y_min, double y_max,
                                                               double scale, int8_t max_depth,
                                                               int32_t max_size,
                                                               bsj_stream_t stream = nullptr)
{
  if (x.size != y.size) throw std::logic_error("x and y columns must have the same length");
  bsj_quadtree t{};
  detail::check(bsj_quadtree_on_points(x.data, y.data, detail::dtype_of<T>(), x.size, x_min, x_max,
                                       y_min, y_max, scale, max_depth, max_size, nullptr, stream,
                                       &t));
  quadtree_table q;
  q.key              = column<uint32_t>(t.key, t.num_nodes, stream);
  q.level            = column<uint8_t>(t.level, t.num_nodes, stream);
  q.is_internal_node = column<uint8_t>(t.is_internal_node, t.num_nodes, stream);
  q.length           = column<uint32_t>(t.length, t.num_nodes, stream);
  q.offset           = column<uint32_t>(t.offset, t.num_nodes, stream);
  q.grid             = t.grid;
  q.sorted_keys      = column<uint32_t>(t.sorted_keys, t.num_points, stream);
  return {column<uint32_t>(t.point_indices, t.num_points, stream), std::move(q)};
}

/// cuspatial::join_quadtree_and_bounding_boxes(quadtree, bbox, x_min, x_max, y_min, y_max, scale,
/// max_depth, mr); bbox = (x_min, y_min, x_max, y_max) columns.
template <typename T>
pair_table join_quadtree_and_bounding_boxes(const quadtree_table& quadtree, column_view<T> bbox_x_min,
                                            column_view<T> bbox_y_min, column_view<T> bbox_x_max,
                                            column_view<T> bbox_y_max, double x_min, double x_max,
                                            double y_min, double y_max, double scale,
                                            int8_t max_depth, bsj_stream_t stream = nullptr)
{
  bsj_pairs p{};
  detail::check(bsj_join_quadtree_and_bounding_boxes(
    quadtree.key.data(), quadtree.level.data(), quadtree.is_internal_node.data(),
    quadtree.length.data(), quadtree.offset.data(), quadtree.key.size(), bbox_x_min.data,
    bbox_y_min.data, bbox_x_max.data, bbox_y_max.data, detail::dtype_of<T>(), bbox_x_min.size,
    x_min, x_max, y_min, y_max, scale, max_depth, nullptr, stream, &p));
  return {column<uint32_t>(p.first, p.size, stream), column<uint32_t>(p.second, p.size, stream)};
}

/// cuspatial::quadtree_point_in_polygon(poly_quad_pairs, quadtree, point_indices, point_x, point_y,
/// poly_offsets, ring_offsets, poly_points_x, poly_points_y, mr)
template <typename T>
pair_table quadtree_point_in_polygon(const pair_table& poly_quad_pairs,
                                     const quadtree_table& quadtree,
                                     column_view<uint32_t> point_indices, column_view<T> point_x,
                                     column_view<T> point_y, column_view<uint32_t> poly_offsets,
                                     column_view<uint32_t> ring_offsets,
                                     column_view<T> poly_points_x, column_view<T> poly_points_y,
                                     bsj_stream_t stream = nullptr)
{
  if (point_indices.size != point_x.size || point_x.size != point_y.size)
    throw std::logic_error("number of points must be the same for both x and y columns");
  if (poly_points_x.size != poly_points_y.size)
    throw std::logic_error("numbers of vertices must be the same for both x and y columns");
  bsj_pairs p{};
  detail::check(bsj_quadtree_point_in_polygon_ex(
    poly_quad_pairs.first.data(), poly_quad_pairs.second.data(), poly_quad_pairs.first.size(),
    quadtree.key.data(), quadtree.level.data(), quadtree.is_internal_node.data(),
    quadtree.length.data(), quadtree.offset.data(), quadtree.key.size(), point_indices.data,
    point_x.data, point_y.data, detail::dtype_of<T>(), point_x.size, poly_offsets.data,
    poly_offsets.size, ring_offsets.data, ring_offsets.size, poly_points_x.data,
    poly_points_y.data, poly_points_x.size, quadtree.grid.valid ? &quadtree.grid : nullptr,
    nullptr, stream, &p));
  return {column<uint32_t>(p.first, p.size, stream), column<uint32_t>(p.second, p.size, stream)};
}

/// cuspatial::point_in_polygon(test_points_x, test_points_y, poly_offsets, poly_ring_offsets,
/// poly_points_x, poly_points_y, mr) -> INT32 bitmask column (caller-allocated here).
template <typename T>
void point_in_polygon(column_view<T> test_points_x, column_view<T> test_points_y,
                      column_view<int32_t> poly_offsets, column_view<int32_t> poly_ring_offsets,
                      column_view<T> poly_points_x, column_view<T> poly_points_y, int32_t* out_mask,
                      bsj_stream_t stream = nullptr)
{
  if (test_points_x.size != test_points_y.size || poly_points_x.size != poly_points_y.size)
    throw std::logic_error("All points must have both x and y values");
  detail::check(bsj_point_in_polygon(test_points_x.data, test_points_y.data, detail::dtype_of<T>(),
                                     test_points_x.size, poly_offsets.data, poly_offsets.size,
                                     poly_ring_offsets.data, poly_ring_offsets.size,
                                     poly_points_x.data, poly_points_y.data, poly_points_x.size,
                                     stream, out_mask));
}

/// cuspatial::pairwise_point_in_polygon(test_points_x, test_points_y, poly_offsets,
/// poly_ring_offsets, poly_points_x, poly_points_y, mr) -> UINT8 column (caller-allocated here).
template <typename T>
void pairwise_point_in_polygon(column_view<T> test_points_x, column_view<T> test_points_y,
                               column_view<int32_t> poly_offsets,
                               column_view<int32_t> poly_ring_offsets,
                               column_view<T> poly_points_x, column_view<T> poly_points_y,
                               uint8_t* out_flags, bsj_stream_t stream = nullptr)
{
  if (test_points_x.size != test_points_y.size || poly_points_x.size != poly_points_y.size)
    throw std::logic_error("All points must have both x and y values");
  detail::check(bsj_pairwise_point_in_polygon(
    test_points_x.data, test_points_y.data, detail::dtype_of<T>(), test_points_x.size,
    poly_offsets.data, poly_offsets.size, poly_ring_offsets.data, poly_ring_offsets.size,
    poly_points_x.data, poly_points_y.data, poly_points_x.size, stream, out_flags));
}

/// Result of quadtree_point_to_nearest_linestring: (point_index, linestring_index, distance).
template <typename T>
struct nearest_table {
  column<uint32_t> point_index;
  column<uint32_t> linestring_index;
  column<T> distance;
};

/// cuspatial::quadtree_point_to_nearest_linestring(linestring_quad_pairs, quadtree, point_indices,
/// point_x, point_y, linestring_offsets, linestring_points_x, linestring_points_y, mr).
/// The three output columns are caller-allocated (n_points rows each); returns the row count
/// (n_points, or 0 for empty inputs as the reference returns an empty table).
template <typename T>
uint64_t quadtree_point_to_nearest_linestring(
  const pair_table& linestring_quad_pairs, const quadtree_table& quadtree,
  column_view<uint32_t> point_indices, column_view<T> point_x, column_view<T> point_y,
  column_view<uint32_t> linestring_offsets, column_view<T> linestring_points_x,
  column_view<T> linestring_points_y, uint32_t* out_point_index, uint32_t* out_linestring_index,
  T* out_distance, bsj_stream_t stream = nullptr)
{
  if (point_indices.size != point_x.size || point_x.size != point_y.size)
    throw std::logic_error("number of points must be the same for both x and y columns");
  if (linestring_points_x.size != linestring_points_y.size)
    throw std::logic_error("numbers of vertices must be the same for both x and y columns");
  uint64_t rows = 0;
  detail::check(bsj_quadtree_point_to_nearest_linestring(
    linestring_quad_pairs.first.data(), linestring_quad_pairs.second.data(),
    linestring_quad_pairs.first.size(), quadtree.key.data(), quadtree.level.data(),
    quadtree.is_internal_node.data(), quadtree.length.data(), quadtree.offset.data(),
    quadtree.key.size(), point_indices.data, point_x.data, point_y.data, detail::dtype_of<T>(),
    point_x.size, linestring_offsets.data, linestring_offsets.size, linestring_points_x.data,
    linestring_points_y.data, linestring_points_x.size, stream, out_point_index,
    out_linestring_index, out_distance, &rows));
  return rows;
}

/// cuspatial::linestring_bounding_boxes(linestring_offsets, x, y, expansion_radius, mr) ->
/// (x_min, y_min, x_max, y_max), caller-allocated columns of linestring_offsets.size-1 rows.
template <typename T>
void linestring_bounding_boxes(column_view<uint32_t> linestring_offsets, column_view<T> x,
                               column_view<T> y, double expansion_radius, T* out_x_min,
                               T* out_y_min, T* out_x_max, T* out_y_max,
                               bsj_stream_t stream = nullptr)
{
  if (x.size != y.size) throw std::logic_error("x and y must be the same size");
  detail::check(bsj_linestring_bounding_boxes(linestring_offsets.data, linestring_offsets.size,
                                              x.data, y.data, detail::dtype_of<T>(), x.size,
                                              expansion_radius, stream, out_x_min, out_y_min,
                                              out_x_max, out_y_max));
}

}  // namespace cuspatial_b200

// api.cu -- the C ABI (include/cuspatial_b200.h): validation, error translation, allocation
// plumbing.  The reference-side checks are cited next to each condition.
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace bsj {

std::atomic<u64> g_launch_count{0};

// implemented in quadtree.cu / bbox_join.cu / pip.cu
void quadtree_on_points_impl(const void* x, const void* y, int dtype, u64 n, double x_min,
                             double x_max, double y_min, double y_max, double scale,
                             int max_depth, int max_size, const bsj_allocator* mr, cudaStream_t s,
                             bsj_quadtree* out);
void join_quadtree_and_bounding_boxes_impl(const u32* key, const u8* level, const u8* internal,
                                           const u32* length, const u32* offset, u64 q,
                                           const void* bx0, const void* by0, const void* bx1,
                                           const void* by1, int dtype, u64 n_boxes, double x_min,
                                           double x_max, double y_min, double y_max, double scale,
                                           int max_depth, const bsj_allocator* mr, cudaStream_t s,
                                           bsj_pairs* out);
void quadtree_point_in_polygon_impl(const u32* pair_poly, const u32* pair_quad, u64 n_pairs,
                                    const u32* key, const u8* level, const u8* internal,
                                    const u32* length, const u32* offset, u64 num_nodes,
                                    const u32* point_indices, const void* px, const void* py,
                                    int dtype, u64 n_points, const u32* poly_offsets,
                                    u64 n_poly_offsets, const u32* ring_offsets,
                                    u64 n_ring_offsets, const void* vx, const void* vy,
                                    u64 n_verts, const bsj_grid* grid, const bsj_allocator* mr,
                                    cudaStream_t s, bsj_pairs* out);
void quadtree_point_in_polygon_compact_impl(
  const u32* pair_poly, const u32* pair_quad, u64 n_pairs, const u32* key, const u8* level,
  const u8* internal, const u32* length, const u32* offset, u64 num_nodes,
  const u32* point_indices, const void* px, const void* py, int dtype, u64 n_points,
  const u32* poly_offsets, u64 n_poly_offsets, const u32* ring_offsets, u64 n_ring_offsets,
  const void* vx, const void* vy, u64 n_verts, const bsj_grid* grid,
  const bsj_coord_segments* segs, const bsj_allocator* mr, cudaStream_t s, bsj_pip_compact* c);
void expand_pip_compact_impl(const u32* pair_poly, const bsj_pip_compact* c, u32 position_base,
                             u32* out_poly, u32* out_point, cudaStream_t s);
void point_in_polygon_impl(const void* px, const void* py, int dtype, u64 n_points,
                           const i32* poly_offsets, u64 n_poly_offsets, const i32* ring_offsets,
                           u64 n_ring_offsets, const void* vx, const void* vy, u64 n_verts,
                           cudaStream_t s, i32* out_mask);
void pairwise_point_in_polygon_impl(const void* px, const void* py, int dtype, u64 n_points,
                                    const i32* poly_offsets, u64 n_poly_offsets,
                                    const i32* ring_offsets, u64 n_ring_offsets, const void* vx,
                                    const void* vy, u64 n_verts, cudaStream_t s, u8* out);
void quadtree_point_to_nearest_linestring_impl(
  const u32* pair_line, const u32* pair_quad, u64 n_pairs, const u32* length, const u32* offset,
  u64 num_nodes, const u32* point_indices, const void* px, const void* py, int dtype, u64 n_points,
  const u32* line_offsets, u64 n_line_offsets, const void* lx, const void* ly, u64 n_verts,
  cudaStream_t s, u32* out_point, u32* out_line, void* out_dist, u64* out_rows);
void linestring_bounding_boxes_impl(const u32* line_offsets, u64 n_line_offsets, const void* lx,
                                    const void* ly, int dtype, u64 n_verts, double r,
                                    cudaStream_t s, void* x0, void* y0, void* x1, void* y1);
void polygon_bounding_boxes_impl(const u32* poly_offsets, u64 n_poly_offsets,
                                 const u32* ring_offsets, u64 n_ring_offsets, const void* vx,
                                 const void* vy, int dtype, u64 n_verts, double r, cudaStream_t s,
                                 void* x0, void* y0, void* x1, void* y1);

void point_keys_histogram_impl(const void* x, const void* y, int dtype, u64 n, double x_min,
                               double x_max, double y_min, double y_max, double scale,
                               int max_depth, int hist_shift, u32* keys, u32* bins, u64 n_bins,
                               u32* point_flags, cudaStream_t s);
void shard_plan_level1_impl(const u32* global_hist, u64 n_bins, const u32* h_rank_sizes,
                            int n_ranks, int rank, int hist_shift, int sub_shift, u32 n_sub,
                            bsj_shard_plan* plan, cudaStream_t s);
void shard_subhistogram_impl(const u32* keys, u64 n, const bsj_shard_plan* plan, int n_ranks,
                             u32 n_sub, u32* bins, cudaStream_t s);
void shard_plan_level2_impl(const u32* local_hist, u64 n_bins, const u32* local_sub,
                            const u32* global_sub, bsj_shard_plan* plan, cudaStream_t s);
void shard_plan_finalize_impl(const u32* counts_matrix, u64 capacity, bsj_shard_plan* plan,
                              cudaStream_t s);
void partition_keys_impl(const u32* keys, u64 n, const bsj_shard_plan* plan, int n_ranks,
                         u32* const* dst_key, u32* const* dst_gid, int use_bulk_copy,
                         cudaStream_t s);
void quadtree_on_keys_impl(u32* keys, u32* values, u64 n, const bsj_grid* g, int max_size,
                           const bsj_allocator* mr, cudaStream_t s, bsj_quadtree* out);

namespace {
thread_local std::string t_err;
thread_local std::vector<std::pair<std::string, float>> t_profile;
std::atomic<int> g_profiling{0};
thread_local stage_timer* t_timer = nullptr;
std::mutex g_pool_mutex;
bool g_pool_done[64] = {};
constexpr int kMaxDevices = 64, kConfigSlots = 8;
std::mutex g_config_mutex;
std::atomic<bool> g_config_done[kConfigSlots][kMaxDevices];
std::atomic<int> g_sm_count[kMaxDevices];
}  // namespace

int num_sms()
{
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int n = g_sm_count[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 148;
  }
  g_sm_count[dev].store(n, std::memory_order_relaxed);
  return n;
}

void configure_once_per_device(int slot, void (*configure)())
{
  int dev = 0;
  BSJ_CUDA_TRY(cudaGetDevice(&dev));
  bool const tracked = dev >= 0 && dev < kMaxDevices && slot >= 0 && slot < kConfigSlots;
  if (tracked && g_config_done[slot][dev].load(std::memory_order_acquire)) return;
  std::lock_guard<std::mutex> lk(g_config_mutex);
  if (tracked && g_config_done[slot][dev].load(std::memory_order_relaxed)) return;
  configure();  // untracked device ordinals are simply configured on every call
  if (tracked) g_config_done[slot][dev].store(true, std::memory_order_release);
}

void ensure_pool_configured()
{
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  if (g_pool_done[dev]) return;
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  if (g_pool_done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = ~0ull;  // keep freed blocks cached in the pool
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  // The refinement gathers 8-byte coordinates at random: ask L2 to fetch 32-byte sectors from
  // HBM instead of promoting every miss to a larger block (measured: see DESIGN.md section 5).
  {
    size_t gran = 32;
    if (const char* e = std::getenv("BSJ_L2_FETCH_GRANULARITY")) gran = (size_t)std::atoi(e);
    if (gran == 32 || gran == 64 || gran == 128)
      cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    cudaGetLastError();
  }
  g_pool_done[dev] = true;
}

stage_timer::stage_timer(cudaStream_t stream) : s(stream), on(g_profiling.load() != 0)
{
  if (on) {
    mark("begin");
    t_timer = this;
  }
}
void prof_mark(const char* name)
{
  if (t_timer) t_timer->mark(name);
}
void stage_timer::mark(const char* name)
{
  if (!on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, s);
  marks.emplace_back(name, e);
}
void stage_timer::finish()
{
  if (!on) return;
  cudaStreamSynchronize(s);
  for (size_t i = 1; i < marks.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
    t_profile.emplace_back(marks[i].first, ms);
  }
}
stage_timer::~stage_timer()
{
  if (t_timer == this) t_timer = nullptr;
  for (auto& m : marks) cudaEventDestroy(m.second);
}

namespace {
// NVTX range over an entry point (the reference's CUSPATIAL_FUNC_RANGE,
// cpp/include/cuspatial/detail/nvtx/ranges.hpp:25-46): header-only NVTX 3, a no-op unless a
// profiler has injected itself.
struct func_range {
  explicit func_range(const char* name) { nvtxRangePushA(name); }
  ~func_range() { nvtxRangePop(); }
  func_range(func_range const&)            = delete;
  func_range& operator=(func_range const&) = delete;
};
#define BSJ_FUNC_RANGE() ::bsj::func_range bsj_nvtx_range_(__func__)

template <typename F>
int guarded(F&& f)
{
  try {
    t_err.clear();
    f();
    return BSJ_SUCCESS;
  } catch (error const& e) {
    t_err = e.msg;
    return e.code;
  } catch (std::bad_alloc const&) {
    t_err = "host allocation failed";
    return BSJ_OUT_OF_MEMORY;
  } catch (std::exception const& e) {
    t_err = e.what();
    return BSJ_CUDA_ERROR;
  }
}
void check_dtype(int dtype)
{
  // cpp/src/indexing/point_quadtree.cu:55-62
  BSJ_EXPECTS(dtype == BSJ_FLOAT32 || dtype == BSJ_FLOAT64,
              "Only floating-point types are supported");
}
}  // namespace
}  // namespace bsj

using namespace bsj;

extern "C" {

int bsj_quadtree_on_points(const void* x, const void* y, int dtype, uint64_t n, double x_min,
                           double x_max, double y_min, double y_max, double scale,
                           int8_t max_depth, int32_t max_size, const bsj_allocator* mr,
                           bsj_stream_t stream, bsj_quadtree* out)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(out != nullptr, "output struct must not be NULL");
    *out = bsj_quadtree{};
    check_dtype(dtype);
    BSJ_EXPECTS(n == 0 || (x != nullptr && y != nullptr),
                "x and y columns must have the same length");  // point_quadtree.cu:166
    quadtree_on_points_impl(x, y, dtype, n, x_min, x_max, y_min, y_max, scale, max_depth, max_size,
                            mr, (cudaStream_t)stream, out);
  });
}

int bsj_join_quadtree_and_bounding_boxes(const uint32_t* key, const uint8_t* level,
                                         const uint8_t* is_internal_node, const uint32_t* length,
                                         const uint32_t* offset, uint64_t num_nodes,
                                         const void* bbox_x_min, const void* bbox_y_min,
                                         const void* bbox_x_max, const void* bbox_y_max, int dtype,
                                         uint64_t n_boxes, double x_min, double x_max, double y_min,
                                         double y_max, double scale, int8_t max_depth,
                                         const bsj_allocator* mr, bsj_stream_t stream,
                                         bsj_pairs* out)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(out != nullptr, "output struct must not be NULL");
    *out = bsj_pairs{};
    check_dtype(dtype);
    // quadtree_bbox_filtering.cu:100-101 (a table is passed as its columns here)
    BSJ_EXPECTS(num_nodes == 0 || (key && level && is_internal_node && length && offset),
                "quadtree table must have 5 columns");
    BSJ_EXPECTS(n_boxes == 0 || (bbox_x_min && bbox_y_min && bbox_x_max && bbox_y_max),
                "bbox table must have 4 columns");
    join_quadtree_and_bounding_boxes_impl(key, level, is_internal_node, length, offset, num_nodes,
                                          bbox_x_min, bbox_y_min, bbox_x_max, bbox_y_max, dtype,
                                          n_boxes, x_min, x_max, y_min, y_max, scale, max_depth, mr,
                                          (cudaStream_t)stream, out);
  });
}

int bsj_quadtree_point_in_polygon_ex(const uint32_t* pair_poly, const uint32_t* pair_quad,
                                     uint64_t n_pairs, const uint32_t* key, const uint8_t* level,
                                     const uint8_t* is_internal_node, const uint32_t* length,
                                     const uint32_t* offset, uint64_t num_nodes,
                                     const uint32_t* point_indices, const void* point_x,
                                     const void* point_y, int dtype, uint64_t n_points,
                                     const uint32_t* poly_offsets, uint64_t n_poly_offsets,
                                     const uint32_t* ring_offsets, uint64_t n_ring_offsets,
                                     const void* poly_points_x, const void* poly_points_y,
                                     uint64_t n_poly_points, const bsj_grid* grid,
                                     const bsj_allocator* mr, bsj_stream_t stream, bsj_pairs* out)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(out != nullptr, "output struct must not be NULL");
    *out = bsj_pairs{};
    check_dtype(dtype);
    // quadtree_point_in_polygon.cu:154-169 (sizes/types are implied by the flat signature)
    BSJ_EXPECTS(n_pairs == 0 || (pair_poly && pair_quad),
                "a quadrant-polygon table must have 2 columns");
    BSJ_EXPECTS(num_nodes == 0 || (length && offset), "a quadtree table must have 5 columns");
    BSJ_EXPECTS(n_points == 0 || (point_indices && point_x && point_y),
                "number of points must be the same for both x and y columns");
    BSJ_EXPECTS(n_poly_points == 0 || (poly_points_x && poly_points_y),
                "numbers of vertices must be the same for both x and y columns");
    quadtree_point_in_polygon_impl(pair_poly, pair_quad, n_pairs, key, level, is_internal_node,
                                   length, offset, num_nodes, point_indices, point_x, point_y,
                                   dtype, n_points, poly_offsets, n_poly_offsets, ring_offsets,
                                   n_ring_offsets, poly_points_x, poly_points_y, n_poly_points,
                                   grid, mr, (cudaStream_t)stream, out);
  });
}

int bsj_quadtree_point_in_polygon(const uint32_t* pair_poly, const uint32_t* pair_quad,
                                  uint64_t n_pairs, const uint32_t* key, const uint8_t* level,
                                  const uint8_t* is_internal_node, const uint32_t* length,
                                  const uint32_t* offset, uint64_t num_nodes,
                                  const uint32_t* point_indices, const void* point_x,
                                  const void* point_y, int dtype, uint64_t n_points,
                                  const uint32_t* poly_offsets, uint64_t n_poly_offsets,
                                  const uint32_t* ring_offsets, uint64_t n_ring_offsets,
                                  const void* poly_points_x, const void* poly_points_y,
                                  uint64_t n_poly_points, const bsj_allocator* mr,
                                  bsj_stream_t stream, bsj_pairs* out)
{
  return bsj_quadtree_point_in_polygon_ex(pair_poly, pair_quad, n_pairs, key, level,
                                          is_internal_node, length, offset, num_nodes,
                                          point_indices, point_x, point_y, dtype, n_points,
                                          poly_offsets, n_poly_offsets, ring_offsets,
                                          n_ring_offsets, poly_points_x, poly_points_y,
                                          n_poly_points, nullptr, mr, stream, out);
}

int bsj_quadtree_point_in_polygon_compact(
  const uint32_t* pair_poly, const uint32_t* pair_quad, uint64_t n_pairs, const uint32_t* key,
  const uint8_t* level, const uint8_t* is_internal_node, const uint32_t* length,
  const uint32_t* offset, uint64_t num_nodes, const uint32_t* point_indices, const void* point_x,
  const void* point_y, int dtype, uint64_t n_points, const uint32_t* poly_offsets,
  uint64_t n_poly_offsets, const uint32_t* ring_offsets, uint64_t n_ring_offsets,
  const void* poly_points_x, const void* poly_points_y, uint64_t n_poly_points,
  const bsj_grid* grid, const bsj_allocator* mr, bsj_stream_t stream, bsj_pip_compact* out)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(out != nullptr, "output struct must not be NULL");
    *out = bsj_pip_compact{};
    check_dtype(dtype);
    BSJ_EXPECTS(n_pairs == 0 || (pair_poly && pair_quad),
                "a quadrant-polygon table must have 2 columns");
    BSJ_EXPECTS(num_nodes == 0 || (length && offset), "a quadtree table must have 5 columns");
    BSJ_EXPECTS(n_points == 0 || (point_indices && point_x && point_y),
                "number of points must be the same for both x and y columns");
    quadtree_point_in_polygon_compact_impl(
      pair_poly, pair_quad, n_pairs, key, level, is_internal_node, length, offset, num_nodes,
      point_indices, point_x, point_y, dtype, n_points, poly_offsets, n_poly_offsets,
      ring_offsets, n_ring_offsets, poly_points_x, poly_points_y, n_poly_points, grid, nullptr, mr,
      (cudaStream_t)stream, out);
  });
}

int bsj_quadtree_point_in_polygon_compact_seg(
  const uint32_t* pair_poly, const uint32_t* pair_quad, uint64_t n_pairs, const uint32_t* key,
  const uint8_t* level, const uint8_t* is_internal_node, const uint32_t* length,
  const uint32_t* offset, uint64_t num_nodes, const uint32_t* point_indices,
  const bsj_coord_segments* segments, int dtype, uint64_t n_points, const uint32_t* poly_offsets,
  uint64_t n_poly_offsets, const uint32_t* ring_offsets, uint64_t n_ring_offsets,
  const void* poly_points_x, const void* poly_points_y, uint64_t n_poly_points,
  const bsj_grid* grid, const bsj_allocator* mr, bsj_stream_t stream, bsj_pip_compact* out)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(out != nullptr, "output struct must not be NULL");
    *out = bsj_pip_compact{};
    check_dtype(dtype);
    BSJ_EXPECTS(n_pairs == 0 || (pair_poly && pair_quad),
                "a quadrant-polygon table must have 2 columns");
    BSJ_EXPECTS(num_nodes == 0 || (length && offset), "a quadtree table must have 5 columns");
    BSJ_EXPECTS(segments != nullptr && segments->n_segments >= 1 &&
                  segments->n_segments <= BSJ_MAX_RANKS,
                "coordinate segments must be given");
    BSJ_EXPECTS(n_points == 0 || point_indices, "point indices must not be NULL");
    quadtree_point_in_polygon_compact_impl(
      pair_poly, pair_quad, n_pairs, key, level, is_internal_node, length, offset, num_nodes,
      point_indices, nullptr, nullptr, dtype, n_points, poly_offsets, n_poly_offsets, ring_offsets,
      n_ring_offsets, poly_points_x, poly_points_y, n_poly_points, grid, segments, mr,
      (cudaStream_t)stream, out);
  });
}

int bsj_expand_pip_compact(const uint32_t* pair_poly, const bsj_pip_compact* c,
                           uint32_t position_base, bsj_stream_t stream,
                           uint32_t* out_polygon_index, uint32_t* out_point_index)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(c != nullptr, "compact result must not be NULL");
    BSJ_EXPECTS(c->n_hits == 0 || (pair_poly && out_polygon_index && out_point_index),
                "output columns must not be NULL");
    expand_pip_compact_impl(pair_poly, c, position_base, out_polygon_index, out_point_index,
                            (cudaStream_t)stream);
  });
}

int bsj_point_in_polygon(const void* point_x, const void* point_y, int dtype, uint64_t n_points,
                         const int32_t* poly_offsets, uint64_t n_poly_offsets,
                         const int32_t* ring_offsets, uint64_t n_ring_offsets,
                         const void* poly_points_x, const void* poly_points_y,
                         uint64_t n_poly_points, bsj_stream_t stream, int32_t* out_mask)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    check_dtype(dtype);
    // point_in_polygon.cu:118-121
    BSJ_EXPECTS(n_points == 0 || (point_x && point_y && out_mask),
                "All points must have both x and y values");
    point_in_polygon_impl(point_x, point_y, dtype, n_points, poly_offsets, n_poly_offsets,
                          ring_offsets, n_ring_offsets, poly_points_x, poly_points_y, n_poly_points,
                          (cudaStream_t)stream, out_mask);
  });
}

int bsj_pairwise_point_in_polygon(const void* point_x, const void* point_y, int dtype,
                                  uint64_t n_points, const int32_t* poly_offsets,
                                  uint64_t n_poly_offsets, const int32_t* ring_offsets,
                                  uint64_t n_ring_offsets, const void* poly_points_x,
                                  const void* poly_points_y, uint64_t n_poly_points,
                                  bsj_stream_t stream, uint8_t* out_flags)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    check_dtype(dtype);
    // point_in_polygon.cu:108-110
    BSJ_EXPECTS(n_points == 0 || (point_x && point_y && out_flags),
                "All points must have both x and y values");
    pairwise_point_in_polygon_impl(point_x, point_y, dtype, n_points, poly_offsets, n_poly_offsets,
                                   ring_offsets, n_ring_offsets, poly_points_x, poly_points_y,
                                   n_poly_points, (cudaStream_t)stream, out_flags);
  });
}

int bsj_quadtree_point_to_nearest_linestring(
  const uint32_t* pair_linestring, const uint32_t* pair_quad, uint64_t n_pairs,
  const uint32_t* key, const uint8_t* level, const uint8_t* is_internal_node,
  const uint32_t* length, const uint32_t* offset, uint64_t num_nodes,
  const uint32_t* point_indices, const void* point_x, const void* point_y, int dtype,
  uint64_t n_points, const uint32_t* linestring_offsets, uint64_t n_linestring_offsets,
  const void* linestring_points_x, const void* linestring_points_y, uint64_t n_linestring_points,
  bsj_stream_t stream, uint32_t* out_point_index, uint32_t* out_linestring_index,
  void* out_distance, uint64_t* out_rows)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    (void)key; (void)level; (void)is_internal_node;
    BSJ_EXPECTS(out_rows != nullptr, "output row count must not be NULL");
    *out_rows = 0;
    check_dtype(dtype);
    // quadtree_point_to_nearest_linestring.cu:161-174 (sizes/types implied by the flat signature)
    BSJ_EXPECTS(n_pairs == 0 || (pair_linestring && pair_quad),
                "a quadrant-linestring table must have 2 columns");
    BSJ_EXPECTS(num_nodes == 0 || (length && offset), "a quadtree table must have 5 columns");
    BSJ_EXPECTS(n_points == 0 || (point_indices && point_x && point_y),
                "number of points must be the same for both x and y columns");
    BSJ_EXPECTS(n_linestring_points == 0 || (linestring_points_x && linestring_points_y),
                "numbers of vertices must be the same for both x and y columns");
    quadtree_point_to_nearest_linestring_impl(
      pair_linestring, pair_quad, n_pairs, length, offset, num_nodes, point_indices, point_x,
      point_y, dtype, n_points, linestring_offsets, n_linestring_offsets, linestring_points_x,
      linestring_points_y, n_linestring_points, (cudaStream_t)stream, out_point_index,
      out_linestring_index, out_distance, out_rows);
  });
}

int bsj_linestring_bounding_boxes(const uint32_t* linestring_offsets,
                                  uint64_t n_linestring_offsets, const void* points_x,
                                  const void* points_y, int dtype, uint64_t n_points,
                                  double expansion_radius, bsj_stream_t stream, void* out_x_min,
                                  void* out_y_min, void* out_x_max, void* out_y_max)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    check_dtype(dtype);
    // linestring_bounding_boxes.cu:135-141
    BSJ_EXPECTS(expansion_radius >= 0, "expansion radius must be greater or equal than 0");
    BSJ_EXPECTS(n_points >= 2 * (n_linestring_offsets ? n_linestring_offsets - 1 : 0),
                "all linestrings must have at least 2 vertices");
    linestring_bounding_boxes_impl(linestring_offsets, n_linestring_offsets, points_x, points_y,
                                   dtype, n_points, expansion_radius, (cudaStream_t)stream,
                                   out_x_min, out_y_min, out_x_max, out_y_max);
  });
}

int bsj_polygon_bounding_boxes(const uint32_t* poly_offsets, uint64_t n_poly_offsets,
                               const uint32_t* ring_offsets, uint64_t n_ring_offsets,
                               const void* poly_points_x, const void* poly_points_y, int dtype,
                               uint64_t n_poly_points, double expansion_radius,
                               bsj_stream_t stream, void* out_x_min, void* out_y_min,
                               void* out_x_max, void* out_y_max)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    check_dtype(dtype);
    // polygon_bounding_boxes.cu:144-151
    BSJ_EXPECTS(expansion_radius >= 0, "expansion radius must be greater or equal than 0");
    polygon_bounding_boxes_impl(poly_offsets, n_poly_offsets, ring_offsets, n_ring_offsets,
                                poly_points_x, poly_points_y, dtype, n_poly_points,
                                expansion_radius, (cudaStream_t)stream, out_x_min, out_y_min,
                                out_x_max, out_y_max);
  });
}

int bsj_point_keys_histogram(const void* x, const void* y, int dtype, uint64_t n, double x_min,
                             double x_max, double y_min, double y_max, double scale,
                             int8_t max_depth, int hist_shift, uint32_t* keys, uint32_t* bins,
                             uint64_t n_bins, uint32_t* point_flags, bsj_stream_t stream)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    check_dtype(dtype);
    BSJ_EXPECTS(n == 0 || (x && y && keys), "x and y columns must have the same length");
    point_keys_histogram_impl(x, y, dtype, n, x_min, x_max, y_min, y_max, scale, max_depth,
                              hist_shift, keys, bins, n_bins, point_flags, (cudaStream_t)stream);
  });
}

int bsj_shard_plan_level1(const uint32_t* global_hist, uint64_t n_bins,
                          const uint32_t* host_rank_sizes, int n_ranks, int rank, int hist_shift,
                          int sub_shift, uint32_t n_sub, bsj_shard_plan* plan, bsj_stream_t stream)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(global_hist && host_rank_sizes && plan, "plan inputs must not be NULL");
    shard_plan_level1_impl(global_hist, n_bins, host_rank_sizes, n_ranks, rank, hist_shift,
                           sub_shift, n_sub, plan, (cudaStream_t)stream);
  });
}

int bsj_shard_subhistogram(const uint32_t* keys, uint64_t n, const bsj_shard_plan* plan,
                           int n_ranks, uint32_t n_sub, uint32_t* bins, bsj_stream_t stream)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(n == 0 || (keys && bins && plan), "keys, bins and plan must not be NULL");
    BSJ_EXPECTS(n_ranks >= 1 && n_ranks <= BSJ_MAX_RANKS && n_sub >= 1 &&
                  (uint64_t)(n_ranks > 1 ? n_ranks - 1 : 1) * n_sub <= 12288,
                "sub-histogram does not fit shared memory");
    shard_subhistogram_impl(keys, n, plan, n_ranks, n_sub, bins, (cudaStream_t)stream);
  });
}

int bsj_shard_plan_level2(const uint32_t* local_hist, uint64_t n_bins, const uint32_t* local_sub,
                          const uint32_t* global_sub, bsj_shard_plan* plan, bsj_stream_t stream)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(local_hist && local_sub && global_sub && plan, "plan inputs must not be NULL");
    shard_plan_level2_impl(local_hist, n_bins, local_sub, global_sub, plan, (cudaStream_t)stream);
  });
}

int bsj_shard_plan_finalize(const uint32_t* counts_matrix, uint64_t capacity, bsj_shard_plan* plan,
                            bsj_stream_t stream)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(counts_matrix && plan, "plan inputs must not be NULL");
    shard_plan_finalize_impl(counts_matrix, capacity, plan, (cudaStream_t)stream);
  });
}

int bsj_partition_keys(const uint32_t* keys, uint64_t n, const bsj_shard_plan* plan, int n_ranks,
                       uint32_t* const* dst_key, uint32_t* const* dst_gid, int use_bulk_copy,
                       bsj_stream_t stream)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(plan && dst_key && dst_gid, "destination pointer tables must not be NULL");
    partition_keys_impl(keys, n, plan, n_ranks, dst_key, dst_gid, use_bulk_copy,
                        (cudaStream_t)stream);
  });
}

int bsj_quadtree_on_keys(uint32_t* keys, uint32_t* values, uint64_t n, const bsj_grid* grid,
                         int32_t max_size, const bsj_allocator* mr, bsj_stream_t stream,
                         bsj_quadtree* out)
{
  BSJ_FUNC_RANGE();
  return guarded([&] {
    BSJ_EXPECTS(out != nullptr, "output struct must not be NULL");
    *out = bsj_quadtree{};
    BSJ_EXPECTS(n == 0 || (keys && values), "keys and values must not be NULL");
    quadtree_on_keys_impl(keys, values, n, grid, max_size, mr, (cudaStream_t)stream, out);
  });
}

void bsj_free(void* ptr, bsj_stream_t stream)
{
  if (ptr) cudaFreeAsync(ptr, (cudaStream_t)stream);
}
void bsj_free_quadtree(bsj_quadtree* t, bsj_stream_t stream)
{
  if (!t) return;
  bsj_free(t->point_indices, stream);
  bsj_free(t->key, stream);
  bsj_free(t->level, stream);
  bsj_free(t->is_internal_node, stream);
  bsj_free(t->length, stream);
  bsj_free(t->offset, stream);
  bsj_free(t->sorted_keys, stream);
  *t = bsj_quadtree{};
}
void bsj_free_pairs(bsj_pairs* p, bsj_stream_t stream)
{
  if (!p) return;
  bsj_free(p->first, stream);
  bsj_free(p->second, stream);
  *p = bsj_pairs{};
}

const char* bsj_last_error(void) { return t_err.c_str(); }
const char* bsj_version(void) { return "cuspatial_b200 0.1 (sm_100a)"; }
uint64_t bsj_kernel_launch_count(void) { return g_launch_count.load(); }

void bsj_set_profiling(int enabled)
{
  g_profiling.store(enabled);
  t_profile.clear();
}
int bsj_get_profile(const char** names, float* millis, int capacity)
{
  static thread_local std::vector<std::string> keep;
  keep.clear();
  int n = 0;
  for (auto& e : t_profile) {
    if (n >= capacity) break;
    keep.push_back(e.first);
    ++n;
  }
  for (int i = 0; i < n; ++i) {
    names[i]  = keep[i].c_str();
    millis[i] = t_profile[i].second;
  }
  t_profile.clear();
  return n;
}

}  // extern "C"

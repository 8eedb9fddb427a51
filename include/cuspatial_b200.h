/*
 * cuspatial_b200.h -- C ABI of the B200-native quadtree point-in-polygon spatial join.
 *
 * Drop-in boundary for ONE hot path of rapidsai/cuspatial (25.06):
 *     quadtree_on_points -> join_quadtree_and_bounding_boxes -> quadtree_point_in_polygon
 *     (+ the non-indexed bitmask point_in_polygon, + polygon_bounding_boxes as input producer)
 *
 * Every entry point states the reference interface it replaces (paths relative to the reference
 * tree).  The reference's column layer takes cudf::column_view / table_view; here a column is a
 * plain device pointer + length and a table is its columns passed one by one, in the reference's
 * column order.  Output columns have the reference's dtypes:
 *     point_indices UINT32 | key UINT32, level UINT8, is_internal_node BOOL8(1 byte), length UINT32,
 *     offset UINT32 | (bbox_offset, quad_offset) UINT32 | (polygon_index, point_index) UINT32 |
 *     bitmask INT32.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers.  `dtype`: 0 = float32, 1 = float64 (the only
 *     coordinate types the reference dispatches on, cpp/src/indexing/point_quadtree.cu:61).
 *   - `stream` is a cudaStream_t (NULL = default stream, which is what the reference uses:
 *     rmm::cuda_stream_default).  Calls are re-entrant and stateless like the reference; sizes
 *     written to the result structs are exact on return, the column CONTENTS are ready in stream
 *     order on `stream` (as with the reference's Thrust calls: no trailing host synchronisation).
 *   - `mr` plays the role of the reference's `rmm::device_async_resource_ref mr`: OUTPUT columns
 *     are allocated through it (so a host framework can hand in its own allocator, e.g. a torch
 *     caching-allocator callback).  NULL = library default (cudaMallocAsync on `stream`); such
 *     buffers are released with bsj_free().  Temporaries always come from the library's pool.
 *   - Return value: BSJ_SUCCESS, or an error code with a thread-local message from
 *     bsj_last_error().  BSJ_INVALID_ARGUMENT corresponds to the reference's cuspatial::logic_error
 *     (CUSPATIAL_EXPECTS, cpp/include/cuspatial/error.hpp:76-79) with the same wording;
 *     BSJ_CUDA_ERROR to cuspatial::cuda_error; BSJ_OUT_OF_MEMORY to rmm::out_of_memory.
 *   - Empty inputs return empty outputs (NULL pointers, size 0) and BSJ_SUCCESS, never an error
 *     (point_quadtree.cu:167-177, quadtree_bbox_filtering.cu:108-114,
 *     quadtree_point_in_polygon.cu:171-178).
 */
#ifndef CUSPATIAL_B200_H
#define CUSPATIAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSJ_SUCCESS 0
#define BSJ_INVALID_ARGUMENT 1
#define BSJ_CUDA_ERROR 2
#define BSJ_OUT_OF_MEMORY 3

#define BSJ_FLOAT32 0
#define BSJ_FLOAT64 1

/* Opaque stream handle: pass a cudaStream_t. */
typedef void* bsj_stream_t;

/* Output allocator, the analogue of rmm::device_async_resource_ref. `allocate` must return device
 * memory usable on `stream` (or NULL on failure -> BSJ_OUT_OF_MEMORY). */
typedef struct bsj_allocator {
  void* (*allocate)(size_t bytes, bsj_stream_t stream, void* ctx);
  void (*deallocate)(void* ptr, size_t bytes, bsj_stream_t stream, void* ctx);
  void* ctx;
} bsj_allocator;

/* How the quadtree's cells map to coordinates: filled by bsj_quadtree_on_points, optionally handed
 * back to bsj_quadtree_point_in_polygon_ex so that quadrants lying entirely inside or outside a
 * polygon are decided from their cell rectangle without gathering their points.  Purely an
 * acceleration hint: results are identical with or without it. */
typedef struct bsj_grid {
  int32_t valid;           /* 0 = unknown (hint ignored)                                        */
  int32_t max_depth;       /* clamped depth the keys were built with                            */
  double min_x, min_y;     /* area-of-interest corners and scale exactly as used for the keys   */
  double max_x, max_y;     /* (values of the coordinate type, widened to double)                */
  double scale;
  int32_t has_nan;         /* some coordinate was NaN (such points key into row/column 0)       */
  int32_t has_out_of_bbox; /* some point lay outside the box (all keyed to the last cell)       */
  const uint32_t* sorted_keys; /* optional: Morton keys of the points in sorted order (device);
                                  lets the refinement decide most points of boundary quadrants
                                  from their finest cell instead of gathering coordinates    */
  uint64_t n_sorted_keys;
} bsj_grid;

/* Result of bsj_quadtree_on_points == the reference's
 * std::pair<std::unique_ptr<cudf::column>, std::unique_ptr<cudf::table>>
 * (cpp/include/cuspatial/point_quadtree.hpp:68-78; column order cpp/src/indexing/point_quadtree.cu:92-118). */
typedef struct bsj_quadtree {
  uint32_t* point_indices;   /* UINT32[num_points]: sorted position -> original point index   */
  uint64_t num_points;
  uint32_t* key;             /* UINT32[num_nodes]                                              */
  uint8_t* level;            /* UINT8 [num_nodes]                                              */
  uint8_t* is_internal_node; /* BOOL8 [num_nodes] (one byte, 0/1)                              */
  uint32_t* length;          /* UINT32[num_nodes]: #children (internal) or #points (leaf)      */
  uint32_t* offset;          /* UINT32[num_nodes]: first child row (internal) / first point pos */
  uint64_t num_nodes;
  bsj_grid grid;             /* not part of the reference's result: optional hint, see bsj_grid */
  uint32_t* sorted_keys;     /* UINT32[num_points], owned like the columns above; grid.sorted_keys
                                points at it.  Free it (or drop the hint) when not needed.     */
} bsj_quadtree;

/* A two-column UINT32 table: (bbox_offset, quad_offset) or (polygon_index, point_index). */
typedef struct bsj_pairs {
  uint32_t* first;
  uint32_t* second;
  uint64_t size;
} bsj_pairs;

/*
 * Replaces cuspatial::quadtree_on_points
 *   (cpp/include/cuspatial/point_quadtree.hpp:68-78, cpp/src/indexing/point_quadtree.cu:154-180).
 * Same argument order and meaning; `x`,`y` are the two coordinate columns (length n each).
 * Clamping as the reference: max_size >= 1, 0 <= max_depth <= 15,
 * scale >= max(dx,dy)/((1<<max_depth)+2) computed in the coordinate type
 * (cpp/include/cuspatial/detail/point_quadtree.cuh:259-268).
 */
int bsj_quadtree_on_points(const void* x, const void* y, int dtype, uint64_t n, double x_min,
                           double x_max, double y_min, double y_max, double scale,
                           int8_t max_depth, int32_t max_size, const bsj_allocator* mr,
                           bsj_stream_t stream, bsj_quadtree* out);

/*
 * Replaces cuspatial::join_quadtree_and_bounding_boxes
 *   (cpp/include/cuspatial/spatial_join.hpp:66-75, cpp/src/join/quadtree_bbox_filtering.cu:90-126).
 * quadtree = its 5 columns; bbox = 4 columns (x_min, y_min, x_max, y_max) of n_boxes rows.
 * Errors (same conditions/wording as quadtree_bbox_filtering.cu:100-106): scale <= 0,
 * !(x_min < x_max && y_min < y_max), !(0 < max_depth < 16).
 * Output rows are ordered exactly as the reference's: by quadtree.offset[quad] ascending, ties by
 * bbox index ascending (cpp/include/cuspatial/detail/join/quadtree_bbox_filtering.cuh:166-180).
 */
int bsj_join_quadtree_and_bounding_boxes(const uint32_t* key, const uint8_t* level,
                                         const uint8_t* is_internal_node, const uint32_t* length,
                                         const uint32_t* offset, uint64_t num_nodes,
                                         const void* bbox_x_min, const void* bbox_y_min,
                                         const void* bbox_x_max, const void* bbox_y_max, int dtype,
                                         uint64_t n_boxes, double x_min, double x_max, double y_min,
                                         double y_max, double scale, int8_t max_depth,
                                         const bsj_allocator* mr, bsj_stream_t stream,
                                         bsj_pairs* out /* first=bbox_offset, second=quad_offset */);

/*
 * Replaces cuspatial::quadtree_point_in_polygon
 *   (cpp/include/cuspatial/spatial_join.hpp:116-126, cpp/src/join/quadtree_point_in_polygon.cu:143-191).
 * poly_quad_pairs = (pair_poly, pair_quad); quadtree = 5 columns; point_indices, point_x, point_y of
 * n_points rows; poly_offsets has n_polygons+1 entries into ring_offsets, ring_offsets has
 * n_rings+1 entries into the vertex columns (GeoArrow; read as uint32 like the reference, :71-74).
 * Output (polygon_index, point_index): rows in (pair order, point order within the quadrant);
 * point_index is the position in point_indices (spatial_join.hpp:112-113).
 */
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
                                  bsj_stream_t stream,
                                  bsj_pairs* out /* first=polygon_index, second=point_index */);

/* Same as bsj_quadtree_point_in_polygon plus the optional cell-geometry hint (`grid` may be NULL or
 * have valid == 0).  The hint must describe the quadtree passed in, and `point_x/point_y` must be
 * the points that quadtree was built from -- which the reference's contract requires anyway. */
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
                                     const bsj_allocator* mr, bsj_stream_t stream, bsj_pairs* out);

/*
 * Compact form of a quadtree_point_in_polygon result (no reference analogue; used by the
 * multi-GPU merge, which all-gathers this instead of the expanded rows).  For pair j:
 * quadrant points are sorted positions pair_offset[j] .. +pair_length[j]-1; pair_class[j] is
 * 0 (no point inside), 1 (every point inside) or 2 (bit l of the ballot words starting at
 * mask_words[pair_word_base[j]] tells whether point l is inside); pair_hits[j] rows start at row
 * pair_row_base[j] of the expanded table.
 */
typedef struct bsj_pip_compact {
  uint32_t* pair_offset;
  uint32_t* pair_length;
  uint32_t* pair_hits;
  uint8_t* pair_class;
  uint64_t* pair_word_base;
  uint64_t* pair_row_base;
  uint32_t* mask_words;
  uint64_t n_pairs, n_words, n_hits;
} bsj_pip_compact;

/* Same inputs as bsj_quadtree_point_in_polygon_ex; buffers of `out` come from `mr`. */
int bsj_quadtree_point_in_polygon_compact(
  const uint32_t* pair_poly, const uint32_t* pair_quad, uint64_t n_pairs, const uint32_t* key,
  const uint8_t* level, const uint8_t* is_internal_node, const uint32_t* length,
  const uint32_t* offset, uint64_t num_nodes, const uint32_t* point_indices, const void* point_x,
  const void* point_y, int dtype, uint64_t n_points, const uint32_t* poly_offsets,
  uint64_t n_poly_offsets, const uint32_t* ring_offsets, uint64_t n_ring_offsets,
  const void* poly_points_x, const void* poly_points_y, uint64_t n_poly_points,
  const bsj_grid* grid, const bsj_allocator* mr, bsj_stream_t stream, bsj_pip_compact* out);

/* Expand a compact result into rows: out_polygon_index/out_point_index must hold c->n_hits rows;
 * point_index = sorted position + position_base (the first global position of the producing
 * rank's key range; 0 on a single GPU). */
int bsj_expand_pip_compact(const uint32_t* pair_poly, const bsj_pip_compact* c,
                           uint32_t position_base, bsj_stream_t stream,
                           uint32_t* out_polygon_index, uint32_t* out_point_index);

/*
 * Replaces cuspatial::point_in_polygon (bitmask form)
 *   (cpp/include/cuspatial/point_in_polygon.hpp:75-82, cpp/src/point_in_polygon/point_in_polygon.cu:153-171).
 * Offsets are int32 (cudf::size_type, :72-73). At most 31 polygons
 * (cpp/include/cuspatial/detail/point_in_polygon.cuh:93-94). out_mask: caller-allocated INT32[n_points];
 * bit i of out_mask[p] is set iff point p is inside polygon i.
 * From 2^20 points on, the call builds a grid of cell classes over the polygons' common box
 * (2^6..2^10 cells per side, chosen so that cells x polygons stays below a quarter of the
 * points; BSJ_BITMASK_GRID_LOG2 overrides, 0 = off) and a
 * point costs one lookup plus the exact test for the polygons its cell could not decide; the
 * result does not depend on the grid.  One host synchronisation (size of the edge index).
 */
int bsj_point_in_polygon(const void* point_x, const void* point_y, int dtype, uint64_t n_points,
                         const int32_t* poly_offsets, uint64_t n_poly_offsets,
                         const int32_t* ring_offsets, uint64_t n_ring_offsets,
                         const void* poly_points_x, const void* poly_points_y,
                         uint64_t n_poly_points, bsj_stream_t stream, int32_t* out_mask);

/*
 * Replaces cuspatial::pairwise_point_in_polygon
 *   (cpp/include/cuspatial/point_in_polygon.hpp:124-131, cpp/src/point_in_polygon/point_in_polygon.cu:172-190;
 *    header form cpp/include/cuspatial/detail/point_in_polygon.cuh:104-145): point i is tested
 * against polygon i only. n_points must equal n_poly_offsets-1 (point_in_polygon.cu:122-125).
 * out_flags: caller-allocated UINT8[n_points], 1 iff point i is inside polygon i (boundary => 0).
 */
int bsj_pairwise_point_in_polygon(const void* point_x, const void* point_y, int dtype,
                                  uint64_t n_points, const int32_t* poly_offsets,
                                  uint64_t n_poly_offsets, const int32_t* ring_offsets,
                                  uint64_t n_ring_offsets, const void* poly_points_x,
                                  const void* poly_points_y, uint64_t n_poly_points,
                                  bsj_stream_t stream, uint8_t* out_flags);

/*
 * Replaces cuspatial::quadtree_point_to_nearest_linestring
 *   (cpp/include/cuspatial/spatial_join.hpp:166-175, cpp/src/join/quadtree_point_to_nearest_linestring.cu:147-196;
 *    header form cpp/include/cuspatial/detail/join/quadtree_point_to_nearest_linestring.cuh:150-314).
 * linestring_quad_pairs = (pair_linestring, pair_quad) as returned by
 * bsj_join_quadtree_and_bounding_boxes on the linestring bounding boxes (rows ordered by quadrant
 * offset); linestring_offsets has n_linestrings+1 entries into the vertex columns.
 * Outputs are caller-allocated columns of n_points rows: out_point_index[i] = i (the position in
 * point_indices), out_linestring_index[i] = nearest linestring among those paired with the
 * point's quadrant, out_distance[i] (same type as the coordinates). *out_rows = n_points, or 0 when
 * any input is empty (the reference returns an empty table, :176-184). Rows of points whose
 * quadrant is in no pair hold distance 0 and index 0 (the reference: distance 0, indices
 * uninitialised).
 */
int bsj_quadtree_point_to_nearest_linestring(
  const uint32_t* pair_linestring, const uint32_t* pair_quad, uint64_t n_pairs,
  const uint32_t* key, const uint8_t* level, const uint8_t* is_internal_node,
  const uint32_t* length, const uint32_t* offset, uint64_t num_nodes,
  const uint32_t* point_indices, const void* point_x, const void* point_y, int dtype,
  uint64_t n_points, const uint32_t* linestring_offsets, uint64_t n_linestring_offsets,
  const void* linestring_points_x, const void* linestring_points_y, uint64_t n_linestring_points,
  bsj_stream_t stream, uint32_t* out_point_index, uint32_t* out_linestring_index,
  void* out_distance, uint64_t* out_rows);

/*
 * Replaces cuspatial::linestring_bounding_boxes
 *   (cpp/include/cuspatial/bounding_boxes.hpp:52-57, cpp/src/bounding_boxes/linestring_bounding_boxes.cu:127-160;
 *    header form cpp/include/cuspatial/detail/bounding_boxes.cuh:96-134).
 * Outputs: 4 caller-allocated columns of n_linestring_offsets-1 rows (x_min, y_min, x_max, y_max),
 * each expanded by `expansion_radius`.
 */
int bsj_linestring_bounding_boxes(const uint32_t* linestring_offsets,
                                  uint64_t n_linestring_offsets, const void* points_x,
                                  const void* points_y, int dtype, uint64_t n_points,
                                  double expansion_radius, bsj_stream_t stream, void* out_x_min,
                                  void* out_y_min, void* out_x_max, void* out_y_max);

/*
 * Replaces cuspatial::polygon_bounding_boxes
 *   (cpp/include/cuspatial/bounding_boxes.hpp, cpp/src/bounding_boxes/polygon_bounding_boxes.cu:132-161):
 * the producer of the bbox table the join consumes. Outputs: 4 caller-allocated columns of
 * n_poly_offsets-1 rows (x_min, y_min, x_max, y_max), each expanded by `expansion_radius`.
 */
int bsj_polygon_bounding_boxes(const uint32_t* poly_offsets, uint64_t n_poly_offsets,
                               const uint32_t* ring_offsets, uint64_t n_ring_offsets,
                               const void* poly_points_x, const void* poly_points_y, int dtype,
                               uint64_t n_poly_points, double expansion_radius,
                               bsj_stream_t stream, void* out_x_min, void* out_y_min,
                               void* out_x_max, void* out_y_max);

/*
 * Multi-GPU (no reference analogue: the reference is single-GPU).  Points
 * are sharded across ranks by contiguous Morton-key range.  One process per GPU; the host layer
 * (cuspatial_b200/multi_gpu.py, torch.distributed over NCCL) only sums the histograms and gathers
 * the send counts -- splitters, send counts and write offsets are computed by kernels into one
 * device struct, and the partition kernel's stores ARE the exchange (peer memory over NVLink).
 * What crosses NVLink is 8 bytes per point, (key, global id): the owner sorts the keys it receives
 * (bsj_quadtree_on_keys) and the refinement reads coordinates on demand through peer pointers
 * (bsj_coord_segments) for the few points whose finest cell is touched by a polygon edge.
 */
#define BSJ_MAX_RANKS 32

/* Device-resident sharding plan (all fields written by the bsj_shard_plan_* kernels). */
typedef struct bsj_shard_plan {
  uint32_t n_ranks, rank, hist_shift, sub_shift;
  uint32_t n_sub, n_targets;
  uint32_t status;                         /* 0 ok, 1 a receive total exceeds the capacity        */
  uint32_t reserved;
  uint32_t gid_base[BSJ_MAX_RANKS + 1];    /* first global point id of every rank, then the total */
  uint32_t bound_bin[BSJ_MAX_RANKS];       /* boundary r+1: first-level bin it falls into ...     */
  uint32_t bound_missing[BSJ_MAX_RANKS];   /* ... and points still missing when that bin starts   */
  uint32_t target_bin[BSJ_MAX_RANKS];      /* distinct boundary bins (second-level histogram rows) */
  uint32_t splitter[BSJ_MAX_RANKS];        /* rank r owns keys in [splitter[r-1], splitter[r])    */
  uint32_t send_count[BSJ_MAX_RANKS];      /* points this rank sends to every destination         */
  uint32_t send_offset[BSJ_MAX_RANKS];     /* where its bucket starts in the destination's buffer  */
  uint32_t recv_total[BSJ_MAX_RANKS];      /* points every rank receives in total                 */
} bsj_shard_plan;

/* Coordinates spread over several arrays: point ids first_id[s] .. first_id[s+1]-1 live in
 * (x[s], y[s]) -- one segment per rank, the pointers may be peer-GPU memory. */
typedef struct bsj_coord_segments {
  int32_t n_segments;
  uint32_t first_id[BSJ_MAX_RANKS + 1];
  const void* x[BSJ_MAX_RANKS];
  const void* y[BSJ_MAX_RANKS];
} bsj_coord_segments;

/* keys[i] = the reference's Morton key of point i
 *   (cpp/include/cuspatial/detail/index/construction/phase_1.cuh:78-85); if bins != NULL,
 *   bins[keys[i] >> hist_shift] += 1 (accumulated; n_bins > max key >> hist_shift); if
 *   point_flags != NULL (device), bit 0 is OR-ed in when a point lies outside the box and bit 1
 *   when a coordinate is NaN.  Stream-ordered, no host synchronisation. */
int bsj_point_keys_histogram(const void* x, const void* y, int dtype, uint64_t n, double x_min,
                             double x_max, double y_min, double y_max, double scale,
                             int8_t max_depth, int hist_shift, uint32_t* keys, uint32_t* bins,
                             uint64_t n_bins, uint32_t* point_flags, bsj_stream_t stream);
/* Plan, level 1: from the SUMMED first-level histogram and the (host-known) rank sizes, the
 * first-level bin holding each rank boundary.  n_sub = 1 << (hist_shift - sub_shift). */
int bsj_shard_plan_level1(const uint32_t* global_hist, uint64_t n_bins,
                          const uint32_t* host_rank_sizes, int n_ranks, int rank, int hist_shift,
                          int sub_shift, uint32_t n_sub, bsj_shard_plan* plan, bsj_stream_t stream);
/* bins[t * n_sub + ((key >> sub_shift) & (n_sub-1))] += 1 for every key whose first-level bin is
 * plan->target_bin[t]; bins must hold (n_ranks-1) * n_sub zero-initialised counters. */
int bsj_shard_subhistogram(const uint32_t* keys, uint64_t n, const bsj_shard_plan* plan,
                           int n_ranks, uint32_t n_sub, uint32_t* bins, bsj_stream_t stream);
/* Plan, level 2: splitters from the SUMMED second-level histogram; this rank's send counts from
 * its own two histograms. */
int bsj_shard_plan_level2(const uint32_t* local_hist, uint64_t n_bins, const uint32_t* local_sub,
                          const uint32_t* global_sub, bsj_shard_plan* plan, bsj_stream_t stream);
/* counts_matrix[s * n_ranks + d] = points rank s sends to rank d (the all-gathered send counts):
 * fills send_offset / recv_total, sets status = 1 if a receive total exceeds `capacity`. */
int bsj_shard_plan_finalize(const uint32_t* counts_matrix, uint64_t capacity, bsj_shard_plan* plan,
                            bsj_stream_t stream);
/* Stable partition of (key, gid_base[rank] + index) by destination rank, written straight into
 * the destinations' receive buffers: dst_key / dst_gid are HOST arrays of n_ranks DEVICE pointers
 * (peer memory over NVLink), 16-byte aligned.  The aligned body of every per-destination run
 * leaves the SM as a bulk copy (cp.async.bulk, use_bulk_copy != 0) or as 128-bit stores.  Does
 * nothing when plan->status != 0. */
int bsj_partition_keys(const uint32_t* keys, uint64_t n, const bsj_shard_plan* plan, int n_ranks,
                       uint32_t* const* dst_key, uint32_t* const* dst_gid, int use_bulk_copy,
                       bsj_stream_t stream);
/* bsj_quadtree_on_points for keys computed elsewhere: stable sort of (keys, values) and the same
 * tree rows; out->point_indices = values in sorted order.  `grid` gives the key geometry (and the
 * out-of-box / NaN flags of the encode).  keys / values are SCRATCH: overwritten. */
int bsj_quadtree_on_keys(uint32_t* keys, uint32_t* values, uint64_t n, const bsj_grid* grid,
                         int32_t max_size, const bsj_allocator* mr, bsj_stream_t stream,
                         bsj_quadtree* out);
/* bsj_quadtree_point_in_polygon_compact with the point coordinates given as segments indexed by
 * the (global) ids stored in point_indices; n_points = number of sorted positions. */
int bsj_quadtree_point_in_polygon_compact_seg(
  const uint32_t* pair_poly, const uint32_t* pair_quad, uint64_t n_pairs, const uint32_t* key,
  const uint8_t* level, const uint8_t* is_internal_node, const uint32_t* length,
  const uint32_t* offset, uint64_t num_nodes, const uint32_t* point_indices,
  const bsj_coord_segments* segments, int dtype, uint64_t n_points, const uint32_t* poly_offsets,
  uint64_t n_poly_offsets, const uint32_t* ring_offsets, uint64_t n_ring_offsets,
  const void* poly_points_x, const void* poly_points_y, uint64_t n_poly_points,
  const bsj_grid* grid, const bsj_allocator* mr, bsj_stream_t stream, bsj_pip_compact* out);

/* Release a buffer the library allocated with its default allocator (mr == NULL). */
void bsj_free(void* ptr, bsj_stream_t stream);
void bsj_free_quadtree(bsj_quadtree* tree, bsj_stream_t stream);
void bsj_free_pairs(bsj_pairs* pairs, bsj_stream_t stream);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* bsj_last_error(void);

/* Library/version string, and the number of kernels this library has launched in this process
 * (used by bench.py for its `gpu_launches` claim). */
const char* bsj_version(void);
uint64_t bsj_kernel_launch_count(void);

/* Per-stage device timings (milliseconds, CUDA events on `stream`) of the last call of each entry
 * point on this thread, when profiling was enabled with bsj_set_profiling(1). Keys are
 * NUL-terminated names; returns the number of entries written (<= capacity). */
void bsj_set_profiling(int enabled);
int bsj_get_profile(const char** names, float* millis, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* CUSPATIAL_B200_H */

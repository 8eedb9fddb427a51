// oracle.cpp -- CPU restatement of the reference's indexed point-in-polygon join.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product path
// (cuspatial_b200/) never does and fails loudly without its CUDA library.
//
// Parity status: PINNED.  This restatement is checked (tests/test_oracle.py) against
//   (1) every golden vector the reference's own tests hold for the path
//       (tests/golden/cuspatial_golden.json, harvested by tests/golden/harvest_golden.py), and
//   (2) the reference's own header-only implementation compiled in place from
//       /root/reference/cpp/include for the host (oracle/_ref/libcuspatial_ref_host.so,
//       recipe in oracle/Makefile) on randomised inputs, for float and double.
//
// Every function cites the reference file:line it restates (rapidsai/cuspatial 25.06).
// Plain C++17 + OpenMP, host pointers only, same C ABI shape as oracle/ref_driver.cpp
// (prefix orc_ instead of ref_) so one ctypes wrapper drives both.
//
// Compile with -ffp-contract=off: the only fused multiply-add on the path is the explicit
// std::fma in node_bounds().

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <type_traits>
#include <vector>

namespace {

thread_local std::string g_err;

// ---------------------------------------------------------------------------------------
// floating_point.cuh:96-130  float_equal: 4-ULP comparison on the biased integer images.
// ---------------------------------------------------------------------------------------
template <typename T>
struct bits_of;
template <>
struct bits_of<float> {
  using type = uint32_t;
};
template <>
struct bits_of<double> {
  using type = uint64_t;
};

template <typename T>
inline bool float_equal(T a, T b)
{
  using B = typename bits_of<T>::type;
  if (std::isnan(a) || std::isnan(b)) return false;
  B ia, ib;
  std::memcpy(&ia, &a, sizeof(T));
  std::memcpy(&ib, &b, sizeof(T));
  B const sign = B(1) << (sizeof(B) * 8 - 1);
  B const ba   = (ia & sign) ? (B)(~ia + 1) : (B)(ia | sign);  // signmagnitude_to_biased :103-108
  B const bb   = (ib & sign) ? (B)(~ib + 1) : (B)(ib | sign);
  return ba >= bb ? (ba - bb) <= 4 : (bb - ba) <= 4;  // default_max_ulp = 4
}

// ---------------------------------------------------------------------------------------
// is_point_in_polygon.cuh:46-101  crossings-multiply with on-edge => false.
// ring_offsets[r]..ring_offsets[r+1] are the vertices of ring r (closing vertex included
// or not); poly rings are [ring_first, ring_last).
// ---------------------------------------------------------------------------------------
template <typename T, typename O>
inline bool is_point_in_polygon(T px, T py, O const* ring_offsets, int64_t ring_first,
                                int64_t ring_last, T const* vx, T const* vy)
{
  bool within  = false;
  bool on_edge = false;
  for (int64_t r = ring_first; r < ring_last; ++r) {
    int64_t const v0 = (int64_t)ring_offsets[r];
    int64_t const v1 = (int64_t)ring_offsets[r + 1];
    if (v1 <= v0) continue;  // empty ring: the reference would read out of bounds; skip
    T bx    = vx[v1 - 1];    // last_segment.v2 == last vertex (:53-55)
    T by    = vy[v1 - 1];
    bool y0 = by > py;
    for (int64_t i = v0; i < v1; ++i) {
      T const ax   = vx[i];
      T const ay   = vy[i];
      T const run  = bx - ax;
      T const rise = by - ay;
      // degenerate segment: skipped WITHOUT advancing b (:65-66)
      if (float_equal(run, T(0)) && float_equal(rise, T(0))) continue;
      T const rise_to_point = py - ay;
      T const run_to_point  = px - ax;
      // point-on-edge test (:71-81); two separately rounded products
      if (float_equal(run * rise_to_point, run_to_point * rise)) {
        T minx = ax, maxx = bx;
        if (minx > maxx) std::swap(minx, maxx);
        if (minx <= px && px <= maxx) {
          on_edge = true;
          break;
        }
      }
      bool const y1 = ay > py;
      if (y1 != y0) {
        T const lhs = (px - ax) * rise;
        T const rhs = run * rise_to_point;
        if ((lhs < rhs) != y1) within = !within;
      }
      bx = ax;
      by = ay;
      y0 = y1;
    }
    if (on_edge) {
      within = false;
      break;
    }
  }
  return within;
}

// ---------------------------------------------------------------------------------------
// z_order.cuh:62-94  bit dilation; arithmetic form of the lookup tables.
// ---------------------------------------------------------------------------------------
inline uint32_t dilate16(uint32_t v)
{
  v &= 0xFFFFu;
  v = (v | (v << 8)) & 0x00FF00FFu;
  v = (v | (v << 4)) & 0x0F0F0F0Fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
inline uint32_t undilate16(uint32_t v)
{
  v &= 0x55555555u;
  v = (v | (v >> 1)) & 0x33333333u;
  v = (v | (v >> 2)) & 0x0F0F0F0Fu;
  v = (v | (v >> 4)) & 0x00FF00FFu;
  v = (v | (v >> 8)) & 0x0000FFFFu;
  return v;
}
inline uint32_t z_order(uint32_t x, uint32_t y) { return (dilate16(y) << 1) | dilate16(x); }
inline uint32_t z_order_x(uint32_t k) { return undilate16(k); }
inline uint32_t z_order_y(uint32_t k) { return undilate16(k >> 1); }

// static_cast<uint16_t>(T) as nvcc compiles it for the device: cvt.rzi.u32.fXX (saturating,
// NaN -> 0) followed by & 0xFFFF.  phase_1.cuh:83-84
template <typename T>
inline uint32_t device_u16(T v)
{
  uint32_t u;
  if (std::isnan(v) || v <= T(0))
    u = 0;
  else if (v >= T(4294967295.0))
    u = 0xFFFFFFFFu;
  else
    u = (uint32_t)v;  // truncation toward zero
  return u & 0xFFFFu;
}

template <typename T>
void* to_buf(std::vector<T> const& v, uint64_t* n)
{
  *n = v.size();
  if (v.empty()) return nullptr;
  void* p = std::malloc(v.size() * sizeof(T));
  std::memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

// stable LSD radix sort of (key, payload) -- semantics of thrust::stable_sort_by_key
template <typename P>
void stable_sort_by_key(std::vector<uint32_t>& key, std::vector<P>& val)
{
  size_t const n = key.size();
  std::vector<uint32_t> k2(n);
  std::vector<P> v2(n);
  for (int pass = 0; pass < 4; ++pass) {
    int const sh = pass * 8;
    size_t hist[257] = {0};
    for (size_t i = 0; i < n; ++i) hist[((key[i] >> sh) & 0xFF) + 1]++;
    bool trivial = false;
    for (int b = 0; b < 256; ++b)
      if (hist[b + 1] == n) trivial = true;
    if (trivial) continue;
    for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
    for (size_t i = 0; i < n; ++i) {
      size_t const d = hist[(key[i] >> sh) & 0xFF]++;
      k2[d]          = key[i];
      v2[d]          = val[i];
    }
    key.swap(k2);
    val.swap(v2);
  }
}

// ---------------------------------------------------------------------------------------
// quadtree_on_points: detail/point_quadtree.cuh:238-272 (clamps), phase_1.cuh:60-95 (keys +
// stable sort), phase_1.cuh:108-181,256-381 (full levels bottom-up), phase_2.cuh:186-345 +
// detail/point_quadtree.cuh:43-188 (prune, flags, offsets, lengths).
// ---------------------------------------------------------------------------------------
template <typename T>
int quadtree_on_points_t(T const* x, T const* y, uint64_t n, double x_min_d, double x_max_d,
                         double y_min_d, double y_max_d, double scale_d, int max_depth_i,
                         int max_size, void** out, uint64_t* out_n)
{
  for (int i = 0; i < 6; ++i) out[i] = nullptr;
  out_n[0] = out_n[1] = 0;
  if (n == 0) return 0;  // point_quadtree.cu:167-177, detail/point_quadtree.cuh:249-257

  // column API casts to T first (point_quadtree.cu:82-84) ...
  T const x1 = (T)x_min_d, x2 = (T)x_max_d, y1 = (T)y_min_d, y2 = (T)y_max_d;
  T scale = (T)scale_d;
  // ... then the header API orders and clamps in T (detail/point_quadtree.cuh:259-268)
  T const min_x = std::min(x1, x2), min_y = std::min(y1, y2);
  T const max_x = std::max(x1, x2), max_y = std::max(y1, y2);
  max_size      = std::max(1, max_size);
  int8_t md8    = (int8_t)max_depth_i;
  int const d   = std::max<int>(0, std::min<int>(15, md8));
  scale         = std::max(scale, std::max(max_x - min_x, max_y - min_y) / T((1 << d) + 2));

  // phase_1.cuh:73-85
  std::vector<uint32_t> keys(n), idx(n);
  uint32_t const oob_key = (uint32_t)((1 << (2 * d)) - 1);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    T const px = x[i], py = y[i];
    uint32_t k;
    if (px < min_x || px > max_x || py < min_y || py > max_y)
      k = oob_key;
    else
      k = z_order(device_u16((px - min_x) / scale), device_u16((py - min_y) / scale));
    keys[i] = k;
    idx[i]  = (uint32_t)i;  // thrust::sequence :89
  }
  stable_sort_by_key(keys, idx);  // :92

  // bottom-level quads = reduce_by_key(keys, 1) (phase_1.cuh:279-285)
  std::vector<std::vector<uint32_t>> lkey(std::max(d, 1)), lcnt(std::max(d, 1)),
    lchild(std::max(d, 1));
  int const bottom = std::max(d, 1) - 1;
  {
    auto& k = lkey[bottom];
    auto& c = lcnt[bottom];
    for (uint64_t i = 0; i < n; ++i) {
      if (i == 0 || keys[i] != keys[i - 1]) {
        k.push_back(keys[i]);
        c.push_back(1);
      } else {
        c.back()++;
      }
    }
    lchild[bottom].assign(k.size(), 0);
  }
  // parent levels: reduce_by_key on key>>2 with (point_count, 1) sums (phase_1.cuh:161-173,
  // utilities.cuh:26-34). Level d-1 is the bottom; level 0 the children of the root.
  uint64_t num_parent_nodes_signed_base = 0;  // nodes in levels 0..d-2
  for (int L = bottom - 1; L >= 0; --L) {
    auto const& ck = lkey[L + 1];
    auto const& cc = lcnt[L + 1];
    auto& k        = lkey[L];
    auto& c        = lcnt[L];
    auto& ch       = lchild[L];
    for (size_t i = 0; i < ck.size(); ++i) {
      uint32_t const pk = ck[i] >> 2;
      if (i == 0 || pk != (ck[i - 1] >> 2)) {
        k.push_back(pk);
        c.push_back(cc[i]);
        ch.push_back(1);
      } else {
        c.back() += cc[i];
        ch.back()++;
      }
    }
    num_parent_nodes_signed_base += k.size();
  }

  std::vector<uint32_t> okey, olen, ooff;
  std::vector<uint8_t> olevel, ointernal;

  if (d <= 1) {
    // num_parent_nodes <= 0 (phase_1.cuh:350): leaf-only tree, detail/point_quadtree.cuh:155-188
    okey = lkey[bottom];
    olen = lcnt[bottom];
    olevel.assign(okey.size(), 0);
    ointernal.assign(okey.size(), 0);
    ooff.resize(okey.size());
    uint32_t acc = 0;
    for (size_t i = 0; i < okey.size(); ++i) {
      ooff[i] = acc;
      acc += olen[i];
    }
  } else {
    // reverse_tree_levels (phase_1.cuh:197-238): concatenate level 0 .. d-1, then
    // remove_unqualified_quads (phase_2.cuh:236-296): a node below level 0 survives iff its
    // parent's point count is > max_size; the parent of child j at level L is found by the
    // running child-count scan (compute_parent_positions :186-221) == running parent index.
    std::vector<uint32_t> kcnt, kchild;  // per kept node: point count, child count
    std::vector<std::vector<uint8_t>> keep(d);
    keep[0].assign(lkey[0].size(), 1);
    for (int L = 1; L < d; ++L) {
      keep[L].assign(lkey[L].size(), 0);
      size_t child = 0;
      for (size_t p = 0; p < lkey[L - 1].size(); ++p) {
        for (uint32_t c = 0; c < lchild[L - 1][p]; ++c, ++child) {
          // children of a removed parent have an even smaller-or-equal parent count chain;
          // the reference tests only the immediate parent's count (phase_2.cuh:268-276)
          keep[L][child] = lcnt[L - 1][p] > (uint32_t)max_size;
        }
      }
    }
    for (int L = 0; L < d; ++L) {
      for (size_t i = 0; i < lkey[L].size(); ++i) {
        if (!keep[L][i]) continue;
        okey.push_back(lkey[L][i]);
        olevel.push_back((uint8_t)L);
        kcnt.push_back(lcnt[L][i]);
        kchild.push_back(lchild[L][i]);
        // construct_non_leaf_indicator (phase_2.cuh:308-345): nodes above the bottom level
        // are internal iff count > max_size; bottom level is always leaf
        ointernal.push_back((L < d - 1 && lcnt[L][i] > (uint32_t)max_size) ? 1 : 0);
      }
    }
    size_t const q = okey.size();
    // leaf first-point positions (phase_2.cuh:105-184): leaves ordered by their key shifted to
    // the bottom level, exclusive scan of their point counts.
    std::vector<uint32_t> fkey, leaf_row;
    for (size_t i = 0; i < q; ++i) {
      if (!ointernal[i]) {
        fkey.push_back(okey[i] << (2 * ((d - 1) - olevel[i])));  // flatten_point_keys :83-96
        leaf_row.push_back((uint32_t)i);
      }
    }
    stable_sort_by_key(fkey, leaf_row);
    std::vector<uint32_t> point_pos(q, 0);
    {
      uint32_t acc = 0;
      for (size_t j = 0; j < leaf_row.size(); ++j) {
        point_pos[leaf_row[j]] = acc;
        acc += kcnt[leaf_row[j]];
      }
    }
    // child positions: exclusive scan of child counts (leaves zeroed) with init = number of
    // level-0 nodes (detail/point_quadtree.cuh:88-101)
    olen.resize(q);
    ooff.resize(q);
    uint32_t acc = (uint32_t)lkey[0].size();
    for (size_t i = 0; i < q; ++i) {
      if (ointernal[i]) {
        ooff[i] = acc;
        olen[i] = kchild[i];
        acc += kchild[i];
      } else {
        ooff[i] = point_pos[i];
        olen[i] = kcnt[i];
      }
    }
  }
  (void)num_parent_nodes_signed_base;

  uint64_t q = 0;
  out[0]     = to_buf(idx, &out_n[0]);
  out[1]     = to_buf(okey, &q);
  out[2]     = to_buf(olevel, &q);
  out[3]     = to_buf(ointernal, &q);
  out[4]     = to_buf(olen, &q);
  out[5]     = to_buf(ooff, &q);
  out_n[1]   = q;
  return 0;
}

// ---------------------------------------------------------------------------------------
// join_quadtree_and_bounding_boxes: quadtree_bbox_filtering.cuh:35-188,
// intersection.cuh:94-128 (bounds + classification), traversal.cuh:63-145 (descent).
// ---------------------------------------------------------------------------------------
template <typename T>
int join_t(uint32_t const* key, uint8_t const* level, uint8_t const* internal,
           uint32_t const* length, uint32_t const* offset, uint64_t q, T const* bx0, T const* by0,
           T const* bx1, T const* by1, uint64_t n_boxes, double x_min, double y_min,
           double scale_d, int max_depth, void** out, uint64_t* out_n)
{
  out[0] = out[1] = nullptr;
  out_n[0]        = 0;
  if (q == 0 || n_boxes == 0) return 0;  // quadtree_bbox_filtering.cu:108-114
  T const vminx = (T)x_min, vminy = (T)y_min, scale = (T)scale_d;  // :67-68

  // intersection.cuh:104-127.  nvcc contracts v_min + k*level_scale into an FMA under the
  // reference's default flags -- restated with an explicit std::fma.
  auto classify = [&](uint32_t node, uint32_t box) -> int {
    uint32_t const k  = key[node];
    int const lv      = level[node];
    T const kx        = (T)z_order_x(k);
    T const ky        = (T)z_order_y(k);
    T const ls        = scale * (T)(1 << (max_depth - 1 - lv));
    T const nx0       = std::fma(kx, ls, vminx);
    T const ny0       = std::fma(ky, ls, vminy);
    T const nx1       = std::fma(kx + T(1), ls, vminx);
    T const ny1       = std::fma(ky + T(1), ls, vminy);
    if (nx0 > bx1[box] || nx1 < bx0[box] || ny0 > by1[box] || ny1 < by0[box]) return 2;  // none
    return internal[node] ? 1 : 0;  // quad : leaf
  };

  // number of level-0 nodes (:53-56)
  uint64_t top = 0;
  for (uint64_t i = 0; i < q; ++i) top += (level[i] == 0);

  std::vector<uint32_t> cur_node, cur_box, out_node, out_box;
  // level 0: pair i -> (node i % top, box i / top)  (:90-111)
  for (uint64_t i = 0; i < top * n_boxes; ++i) {
    uint32_t const node = (uint32_t)(i % top), box = (uint32_t)(i / top);
    int const t = classify(node, box);
    if (t == 0) {
      out_node.push_back(node);
      out_box.push_back(box);
    } else if (t == 1) {
      cur_node.push_back(node);
      cur_box.push_back(box);
    }
  }
  // descend (:116-163)
  for (int lv = 1; lv < max_depth && !cur_node.empty(); ++lv) {
    std::vector<uint32_t> nn, nb;
    for (size_t i = 0; i < cur_node.size(); ++i) {  // traversal.cuh:73-138
      uint32_t const p = cur_node[i];
      for (uint32_t c = 0; c < length[p]; ++c) {
        nn.push_back(offset[p] + c);
        nb.push_back(cur_box[i]);
      }
    }
    cur_node.clear();
    cur_box.clear();
    for (size_t i = 0; i < nn.size(); ++i) {
      int const t = classify(nn[i], nb[i]);
      if (t == 0) {
        out_node.push_back(nn[i]);
        out_box.push_back(nb[i]);
      } else if (t == 1) {
        cur_node.push_back(nn[i]);
        cur_box.push_back(nb[i]);
      }
    }
  }
  // stable sort by quadtree.offset[node] (:166-180)
  std::vector<uint32_t> order(out_node.size());
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    return offset[out_node[a]] < offset[out_node[b]];
  });
  std::vector<uint32_t> rb(order.size()), rn(order.size());
  for (size_t i = 0; i < order.size(); ++i) {
    rb[i] = out_box[order[i]];
    rn[i] = out_node[order[i]];
  }
  out[0] = to_buf(rb, &out_n[0]);
  out[1] = to_buf(rn, &out_n[0]);
  return 0;
}

// ---------------------------------------------------------------------------------------
// quadtree_point_in_polygon: detail/join/quadtree_point_in_polygon.cuh:41-94,104-217.
// Row order = (pair order, local point order); point_index is the position in point_indices.
// ---------------------------------------------------------------------------------------
template <typename T>
int qpip_t(uint32_t const* pair_poly, uint32_t const* pair_quad, uint64_t n_pairs,
           uint32_t const* length, uint32_t const* offset, uint64_t q,
           uint32_t const* point_indices, T const* px, T const* py, uint64_t n_points,
           uint32_t const* poly_offsets, uint64_t n_poly_offsets, uint32_t const* ring_offsets,
           uint64_t n_ring_offsets, T const* vx, T const* vy, void** out, uint64_t* out_n)
{
  out[0] = out[1] = nullptr;
  out_n[0]        = 0;
  if (n_pairs == 0 || q == 0 || n_points == 0 || n_poly_offsets == 0) return 0;  // .cu:171-178
  std::vector<std::vector<uint32_t>> hits(n_pairs);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t j = 0; j < (int64_t)n_pairs; ++j) {
    uint32_t const poly = pair_poly[j], quad = pair_quad[j];
    // polygons[poly][0]: rings poly_offsets[poly] .. poly_offsets[poly+1] (.cu:68-76)
    int64_t const r0 = poly_offsets[poly], r1 = poly_offsets[poly + 1];
    auto& h          = hits[j];
    for (uint32_t l = 0; l < length[quad]; ++l) {
      uint32_t const pos = offset[quad] + l;  // :66
      uint32_t const pi  = point_indices[pos];
      if (is_point_in_polygon<T, uint32_t>(px[pi], py[pi], ring_offsets, r0, r1, vx, vy))
        h.push_back(pos);
    }
  }
  std::vector<uint32_t> rp, rq;
  for (uint64_t j = 0; j < n_pairs; ++j)
    for (uint32_t pos : hits[j]) {
      rp.push_back(pair_poly[j]);
      rq.push_back(pos);
    }
  out[0] = to_buf(rp, &out_n[0]);
  out[1] = to_buf(rq, &out_n[0]);
  return 0;
}

// detail/point_in_polygon.cuh:43-66,69-102  bitmask over <= 31 polygons
template <typename T>
int pip_t(T const* px, T const* py, uint64_t n_points, int32_t const* poly_offsets,
          uint64_t n_poly_offsets, int32_t const* ring_offsets, uint64_t n_ring_offsets,
          T const* vx, T const* vy, int32_t* out_mask)
{
  int64_t const n_poly = n_poly_offsets ? (int64_t)n_poly_offsets - 1 : 0;
  if (n_poly > 31) {
    g_err = "Number of polygons cannot exceed 31";
    return 1;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n_points; ++i) {
    int32_t m = 0;
    for (int64_t p = 0; p < n_poly; ++p)
      m |= (int32_t)is_point_in_polygon<T, int32_t>(px[i], py[i], ring_offsets, poly_offsets[p],
                                                    poly_offsets[p + 1], vx, vy)
           << p;
    out_mask[i] = m;
  }
  return 0;
}

// detail/point_in_polygon.cuh:104-145  pairwise: point i against polygon i, uint8 result
template <typename T>
int pairwise_pip_t(T const* px, T const* py, uint64_t n_points, int32_t const* poly_offsets,
                   uint64_t n_poly_offsets, int32_t const* ring_offsets, T const* vx, T const* vy,
                   uint8_t* out)
{
  // cpp/src/point_in_polygon/point_in_polygon.cu:122-125
  if (n_points != (n_poly_offsets ? n_poly_offsets - 1 : 0)) {
    g_err = "Must pass in the same number of points as polygons.";
    return 1;
  }
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t)n_points; ++i)
    out[i] = is_point_in_polygon<T, int32_t>(px[i], py[i], ring_offsets, poly_offsets[i],
                                             poly_offsets[i + 1], vx, vy)
               ? 1
               : 0;
  return 0;
}

// vec_2d.hpp:166-170 dot(a,b) = a.x*b.x + a.y*b.y.  nvcc (default -fmad=true) contracts it to
// fma(a.x, b.x, a.y*b.y) -- checked in the SASS of oracle/_ref/libcuspatial_ref_cuda.so
// (FMUL/DMUL of the y terms first, then one FFMA/DFMA); restated explicitly because this file is
// compiled with -ffp-contract=off.
template <typename T>
inline T dot2(T ax, T ay, T bx, T by)
{
  return std::fma(ax, bx, ay * by);
}

// detail/algorithm/point_linestring_distance.cuh:33-53
template <typename T>
inline T point_linestring_distance(T px, T py, T const* lx, T const* ly, uint32_t v0, uint32_t v1)
{
  T distance_squared = std::numeric_limits<T>::max();
  for (uint32_t i = v0; i + 1 < v1; ++i) {  // linestring_ref: segments (v[i], v[i+1])
    T const ax = lx[i], ay = ly[i], bx = lx[i + 1], by = ly[i + 1];
    T const v1px = px - ax, v1py = py - ay, v2px = px - bx, v2py = py - by;
    T const d0 = dot2(v1px, v1py, v1px, v1py);
    T const d1 = dot2(v2px, v2py, v2px, v2py);
    T const ex = bx - ax, ey = by - ay;
    T const d2 = dot2(ex, ey, ex, ey);        // segment::length2
    T const d3 = dot2(v1px, v1py, ex, ey);    // proj2
    T const r  = d3 * d3 / d2;
    T const d  = (d3 <= 0 || r >= d2) ? std::fmin(d0, d1) : d0 - r;
    distance_squared = std::fmin(distance_squared, d);
  }
  return std::sqrt(distance_squared);
}

// detail/join/quadtree_point_to_nearest_linestring.cuh:150-314, literally: the candidate list is
// enumerated in the reference's "transposed" order (:64-92) so that the candidates of one point
// are consecutive, reduced by key with the reference's selection rule (:283-300: a zero distance
// loses to anything, ties go to the smaller linestring id), then scattered to the point's sorted
// position (:309-316).  Points that belong to no candidate quadrant keep distance 0 (:264); the
// reference leaves their two index columns uninitialised -- here they are 0.
template <typename T>
int nearest_linestring_t(uint32_t const* pair_line, uint32_t const* pair_quad, uint64_t n_pairs,
                         uint32_t const* length, uint32_t const* offset, uint64_t q,
                         uint32_t const* point_indices, T const* px, T const* py,
                         uint64_t n_points, uint32_t const* line_offsets, uint64_t n_line_offsets,
                         T const* lx, T const* ly, void** out, uint64_t* out_n)
{
  out[0] = out[1] = out[2] = nullptr;
  out_n[0]                 = 0;
  // cpp/src/join/quadtree_point_to_nearest_linestring.cu:176-184
  if (n_pairs == 0 || q == 0 || n_points == 0 || n_line_offsets == 0) return 0;
  std::vector<uint32_t> quad_len(n_pairs), quad_off(n_pairs), local(n_pairs + 1, 0);
  for (uint64_t j = 0; j < n_pairs; ++j) {
    quad_len[j]  = length[pair_quad[j]];
    quad_off[j]  = offset[pair_quad[j]];
    local[j + 1] = local[j] + quad_len[j];  // :176-183 (uint32 arithmetic like IndexType)
  }
  uint32_t const total = local[n_pairs];
  std::vector<uint32_t> out_point(n_points, 0), out_line(n_points, 0);
  std::vector<T> out_dist(n_points, (T)0);

  struct cand { uint32_t point, line; T d; };
  auto candidate = [&](uint32_t g) {
    // get_quad_and_local_point_indices.cuh:37-44
    uint32_t const j  = (uint32_t)(std::upper_bound(local.begin(), local.end(), g) - local.begin()) - 1;
    uint32_t const lp = g - local[j];
    // :44-60 run of pairs with the same quadrant offset around j
    uint32_t const lhs = (uint32_t)(std::lower_bound(quad_off.begin(), quad_off.begin() + j, quad_off[j]) - quad_off.begin());
    uint32_t const rhs = (uint32_t)(std::upper_bound(quad_off.begin() + j, quad_off.end(), quad_off[j]) - quad_off.begin());
    uint32_t const local_line = j - lhs, n_lines = rhs - lhs;
    // :76-91
    uint32_t const t     = local_line * quad_len[j] + lp;
    uint32_t const point = t / n_lines + quad_off[j];
    uint32_t const pj    = t % n_lines + (j - local_line);
    uint32_t const line  = pair_line[pj];
    uint32_t const pi    = point_indices[point];
    return cand{point, line,
                point_linestring_distance<T>(px[pi], py[pi], lx, ly, line_offsets[line],
                                             line_offsets[line + 1])};
  };
  auto select = [](cand const& l, cand const& r) {  // :283-300
    if (l.d == (T)0) return r;
    if (r.d == (T)0) return l;
    if (l.d == r.d) return l.line < r.line ? l : r;
    return l.d < r.d ? l : r;
  };
  uint32_t g = 0;
  while (g < total) {  // reduce_by_key over equal consecutive point ids, then scatter
    cand acc = candidate(g++);
    while (g < total) {
      cand const c = candidate(g);
      if (c.point != acc.point) break;
      cand sel  = select(acc, c);
      sel.point = acc.point;
      acc       = sel;
      ++g;
    }
    if (acc.point < n_points) {
      out_point[acc.point] = acc.point;
      out_line[acc.point]  = acc.line;
      out_dist[acc.point]  = acc.d;
    }
  }
  out[0] = to_buf(out_point, &out_n[0]);
  out[1] = to_buf(out_line, &out_n[0]);
  out[2] = to_buf(out_dist, &out_n[0]);
  return 0;
}

// detail/bounding_boxes.cuh:96-134  per-linestring min/max of (v - r, v + r)
template <typename T>
int line_bbox_t(uint32_t const* line_offsets, uint64_t n_line_offsets, T const* lx, T const* ly,
                uint64_t n_verts, T r, T* x0, T* y0, T* x1, T* y1)
{
  if (n_line_offsets < 2 || n_verts == 0) return 0;
  for (uint64_t p = 0; p + 1 < n_line_offsets; ++p) {
    T ax = std::numeric_limits<T>::infinity(), ay = ax, bx = -ax, by = -ax;
    for (uint64_t i = line_offsets[p]; i < line_offsets[p + 1] && i < n_verts; ++i) {
      ax = std::fmin(ax, lx[i] - r); ay = std::fmin(ay, ly[i] - r);
      bx = std::fmax(bx, lx[i] + r); by = std::fmax(by, ly[i] + r);
    }
    x0[p] = ax; y0[p] = ay; x1[p] = bx; y1[p] = by;
  }
  return 0;
}

// detail/bounding_boxes.cuh:36-60,136-184  per-polygon min/max of (v - r, v + r)
template <typename T>
int poly_bbox_t(uint32_t const* poly_offsets, uint64_t n_poly_offsets,
                uint32_t const* ring_offsets, uint64_t n_ring_offsets, T const* vx, T const* vy,
                uint64_t n_verts, T r, T* x0, T* y0, T* x1, T* y1)
{
  if (n_poly_offsets < 2) return 0;
  for (uint64_t p = 0; p + 1 < n_poly_offsets; ++p) {
    uint64_t const vb = ring_offsets[poly_offsets[p]];
    uint64_t const ve = ring_offsets[poly_offsets[p + 1]];
    T lx = std::numeric_limits<T>::max(), ly = lx, hx = std::numeric_limits<T>::lowest(), hy = hx;
    bool first = true;
    for (uint64_t i = vb; i < ve && i < n_verts; ++i) {
      T const ax = vx[i] - r, ay = vy[i] - r, cx = vx[i] + r, cy = vy[i] + r;
      if (first) {
        lx = ax; ly = ay; hx = cx; hy = cy;
        first = false;
      } else {
        lx = std::min(lx, ax); ly = std::min(ly, ay);
        hx = std::max(hx, cx); hy = std::max(hy, cy);
      }
    }
    x0[p] = lx; y0[p] = ly; x1[p] = hx; y1[p] = hy;
  }
  return 0;
}

}  // namespace

extern "C" {

int orc_is_cuda() { return 0; }
const char* orc_last_error() { return g_err.c_str(); }
void orc_free(void* p) { std::free(p); }

int orc_quadtree_on_points(void const* x, void const* y, int dtype, uint64_t n, double x_min,
                           double x_max, double y_min, double y_max, double scale, int max_depth,
                           int max_size, void** out, uint64_t* out_n)
{
  return dtype == 0
           ? quadtree_on_points_t<float>((float const*)x, (float const*)y, n, x_min, x_max, y_min,
                                         y_max, scale, max_depth, max_size, out, out_n)
           : quadtree_on_points_t<double>((double const*)x, (double const*)y, n, x_min, x_max,
                                          y_min, y_max, scale, max_depth, max_size, out, out_n);
}

int orc_join_quadtree_and_bounding_boxes(uint32_t const* key, uint8_t const* level,
                                         uint8_t const* internal, uint32_t const* length,
                                         uint32_t const* offset, uint64_t q, void const* bx0,
                                         void const* by0, void const* bx1, void const* by1,
                                         int dtype, uint64_t n_boxes, double x_min, double y_min,
                                         double scale, int max_depth, void** out, uint64_t* out_n)
{
  return dtype == 0
           ? join_t<float>(key, level, internal, length, offset, q, (float const*)bx0,
                           (float const*)by0, (float const*)bx1, (float const*)by1, n_boxes, x_min,
                           y_min, scale, max_depth, out, out_n)
           : join_t<double>(key, level, internal, length, offset, q, (double const*)bx0,
                            (double const*)by0, (double const*)bx1, (double const*)by1, n_boxes,
                            x_min, y_min, scale, max_depth, out, out_n);
}

int orc_quadtree_point_in_polygon(uint32_t const* pair_poly, uint32_t const* pair_quad,
                                  uint64_t n_pairs, uint32_t const* key, uint8_t const* level,
                                  uint8_t const* internal, uint32_t const* length,
                                  uint32_t const* offset, uint64_t q,
                                  uint32_t const* point_indices, void const* px, void const* py,
                                  int dtype, uint64_t n_points, uint32_t const* poly_offsets,
                                  uint64_t n_poly_offsets, uint32_t const* ring_offsets,
                                  uint64_t n_ring_offsets, void const* vx, void const* vy,
                                  uint64_t n_verts, void** out, uint64_t* out_n)
{
  (void)key; (void)level; (void)internal; (void)n_verts;
  return dtype == 0
           ? qpip_t<float>(pair_poly, pair_quad, n_pairs, length, offset, q, point_indices,
                           (float const*)px, (float const*)py, n_points, poly_offsets,
                           n_poly_offsets, ring_offsets, n_ring_offsets, (float const*)vx,
                           (float const*)vy, out, out_n)
           : qpip_t<double>(pair_poly, pair_quad, n_pairs, length, offset, q, point_indices,
                            (double const*)px, (double const*)py, n_points, poly_offsets,
                            n_poly_offsets, ring_offsets, n_ring_offsets, (double const*)vx,
                            (double const*)vy, out, out_n);
}

int orc_point_in_polygon(void const* px, void const* py, int dtype, uint64_t n_points,
                         int32_t const* poly_offsets, uint64_t n_poly_offsets,
                         int32_t const* ring_offsets, uint64_t n_ring_offsets, void const* vx,
                         void const* vy, uint64_t n_verts, int32_t* out_mask)
{
  (void)n_verts;
  return dtype == 0
           ? pip_t<float>((float const*)px, (float const*)py, n_points, poly_offsets,
                          n_poly_offsets, ring_offsets, n_ring_offsets, (float const*)vx,
                          (float const*)vy, out_mask)
           : pip_t<double>((double const*)px, (double const*)py, n_points, poly_offsets,
                           n_poly_offsets, ring_offsets, n_ring_offsets, (double const*)vx,
                           (double const*)vy, out_mask);
}

int orc_pairwise_point_in_polygon(void const* px, void const* py, int dtype, uint64_t n_points,
                                  int32_t const* poly_offsets, uint64_t n_poly_offsets,
                                  int32_t const* ring_offsets, uint64_t n_ring_offsets,
                                  void const* vx, void const* vy, uint64_t n_verts, uint8_t* out)
{
  (void)n_verts;
  (void)n_ring_offsets;
  return dtype == 0 ? pairwise_pip_t<float>((float const*)px, (float const*)py, n_points,
                                            poly_offsets, n_poly_offsets, ring_offsets,
                                            (float const*)vx, (float const*)vy, out)
                    : pairwise_pip_t<double>((double const*)px, (double const*)py, n_points,
                                             poly_offsets, n_poly_offsets, ring_offsets,
                                             (double const*)vx, (double const*)vy, out);
}

int orc_quadtree_point_to_nearest_linestring(
  uint32_t const* pair_line, uint32_t const* pair_quad, uint64_t n_pairs, uint32_t const* key,
  uint8_t const* level, uint8_t const* internal, uint32_t const* length, uint32_t const* offset,
  uint64_t q, uint32_t const* point_indices, void const* px, void const* py, int dtype,
  uint64_t n_points, uint32_t const* line_offsets, uint64_t n_line_offsets, void const* lx,
  void const* ly, uint64_t n_verts, void** out, uint64_t* out_n)
{
  (void)key; (void)level; (void)internal; (void)n_verts;
  return dtype == 0
           ? nearest_linestring_t<float>(pair_line, pair_quad, n_pairs, length, offset, q,
                                         point_indices, (float const*)px, (float const*)py,
                                         n_points, line_offsets, n_line_offsets, (float const*)lx,
                                         (float const*)ly, out, out_n)
           : nearest_linestring_t<double>(pair_line, pair_quad, n_pairs, length, offset, q,
                                          point_indices, (double const*)px, (double const*)py,
                                          n_points, line_offsets, n_line_offsets,
                                          (double const*)lx, (double const*)ly, out, out_n);
}

int orc_linestring_bounding_boxes(uint32_t const* line_offsets, uint64_t n_line_offsets,
                                  void const* lx, void const* ly, int dtype, uint64_t n_verts,
                                  double expansion, void* x0, void* y0, void* x1, void* y1)
{
  return dtype == 0 ? line_bbox_t<float>(line_offsets, n_line_offsets, (float const*)lx,
                                         (float const*)ly, n_verts, (float)expansion, (float*)x0,
                                         (float*)y0, (float*)x1, (float*)y1)
                    : line_bbox_t<double>(line_offsets, n_line_offsets, (double const*)lx,
                                          (double const*)ly, n_verts, expansion, (double*)x0,
                                          (double*)y0, (double*)x1, (double*)y1);
}

int orc_polygon_bounding_boxes(uint32_t const* poly_offsets, uint64_t n_poly_offsets,
                               uint32_t const* ring_offsets, uint64_t n_ring_offsets,
                               void const* vx, void const* vy, int dtype, uint64_t n_verts,
                               double expansion, void* x0, void* y0, void* x1, void* y1)
{
  return dtype == 0
           ? poly_bbox_t<float>(poly_offsets, n_poly_offsets, ring_offsets, n_ring_offsets,
                                (float const*)vx, (float const*)vy, n_verts, (float)expansion,
                                (float*)x0, (float*)y0, (float*)x1, (float*)y1)
           : poly_bbox_t<double>(poly_offsets, n_poly_offsets, ring_offsets, n_ring_offsets,
                                 (double const*)vx, (double const*)vy, n_verts, expansion,
                                 (double*)x0, (double*)y0, (double*)x1, (double*)y1);
}

}  // extern "C"

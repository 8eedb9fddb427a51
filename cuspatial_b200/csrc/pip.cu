// pip.cu -- B200-native point-in-polygon refinement
// (replaces cuspatial::quadtree_point_in_polygon and the bitmask cuspatial::point_in_polygon).
//
// Reference behaviour restated:
//   cpp/include/cuspatial/detail/join/quadtree_point_in_polygon.cuh:41-94,104-235 (candidate
//   enumeration in (pair, local point) order, stable compaction),
//   detail/algorithm/is_point_in_polygon.cuh:46-101 (crossings-multiply, on-edge => false),
//   detail/utility/floating_point.cuh:96-130 (4-ULP float_equal),
//   detail/point_in_polygon.cuh:43-102 (bitmask over <= 31 polygons).
//
// Design (not a port).  The reference tests every candidate against EVERY vertex of its polygon
// after a ~19-step binary search and two dependent random gathers per candidate.  Here:
//   * stage 1 (pip_classify_kernel) decides whole quadrants from their cell rectangle through a
//     per-polygon y-slab edge index -- no point is read for them;
//   * stage 2 works on (pair, point tile) UNITS of the undecided pairs, one warp per unit drawn
//     from tickets: with the sorted Morton keys at hand (pip_eval_cells_kernel) a point is decided
//     from the centre of its finest cell unless an edge touches that cell, and only touched points
//     are gathered and evaluated exactly; without the keys (pip_eval_kernel) the tile's points are
//     gathered once (256-point tiles, 8 per lane, coordinates in registers);
//   * per (tile, polygon) the warp scans the polygon's edges 32 at a time and keeps only the
//     edges that can influence a point of the tile -- the predicate is
//         inside = XOR_e crossing(e)  AND NOT  OR_e on_edge(e),
//     crossing(e) needs the point's y inside the edge's y-range, on_edge(e) needs the point within
//     a few ULP of the segment (or the edge vertical and x equal) -- so edges whose (slightly
//     widened) y-range misses the tile's y-range, and that are not vertical edges inside the
//     tile's x-range, are skipped EXACTLY.  Surviving edges are broadcast by shuffle and evaluated
//     with the reference's own arithmetic, operation for operation (separately rounded products,
//     no FMA: __dmul_rn/__dsub_rn);
//   * results leave as 32-bit ballot words in (pair, point) order; a 64-bit scan of the per-pair
//     hit counts gives every pair its output offset and a second kernel expands the words into
//     (polygon_index, point_index) rows -- the reference's stable copy_if order, exact output size,
//     no worst-case buffer (quadtree_point_in_polygon.cuh:195-196) and no OOM retry.
//   The exact-skipping argument needs products that neither underflow nor overflow; polygons or
//   point tiles with coordinates outside [2^-200, 2^200] (fp32: [2^-30, 2^30]), NaN or Inf fall
//   back to evaluating every edge in the reference's order (pip_reference below).
#include "scan.cuh"

#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace bsj {

namespace {

template <typename T>
struct fpp;
template <>
struct fpp<float> {
  using bits_t = u32;
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ u32 bits(float a) { return __float_as_uint(a); }
  static __host__ __device__ constexpr float tiny() { return 9.313225746154785e-10f; }  // 2^-30
  static __host__ __device__ constexpr float huge() { return 1073741824.0f; }           // 2^30
  static __host__ __device__ constexpr float eps() { return 6.103515625e-05f; }         // 2^-14
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};
template <>
struct fpp<double> {
  using bits_t = u64;
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ u64 bits(double a) { return (u64)__double_as_longlong(a); }
  static __host__ __device__ constexpr double tiny() { return 6.223015277861142e-61; }   // 2^-200
  static __host__ __device__ constexpr double huge() { return 1.6069380442589903e+60; }  // 2^200
  static __host__ __device__ constexpr double eps() { return 9.094947017729282e-13; }    // 2^-40
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
};

// floating_point.cuh:118-130 (max_ulp = 4)
template <typename T>
__device__ __forceinline__ bool float_equal(T a, T b)
{
  using B = typename fpp<T>::bits_t;
  if (a != a || b != b) return false;
  B const sign = (B)1 << (sizeof(B) * 8 - 1);
  B const ia = fpp<T>::bits(a), ib = fpp<T>::bits(b);
  B const ba = (ia & sign) ? (B)(~ia + 1) : (B)(ia | sign);
  B const bb = (ib & sign) ? (B)(~ib + 1) : (B)(ib | sign);
  return ba >= bb ? (ba - bb) <= 4 : (bb - ba) <= 4;
}

// |c| is zero or a "comfortable" normal number (see header comment)
template <typename T>
__device__ __forceinline__ bool comfy(T c)
{
  T const a = fabs(c);
  return a == (T)0 || (a >= fpp<T>::tiny() && a <= fpp<T>::huge());  // false for NaN/Inf
}
template <typename T>
__device__ __forceinline__ bool comfy_delta(T d)  // edge run/rise
{
  T const a = fabs(d);
  return a == (T)0 || (a >= fpp<T>::tiny() && a <= (T)2 * fpp<T>::huge());
}

// is_point_in_polygon.cuh:46-101, operation for operation (fallback path, any input).
template <typename T>
__device__ bool pip_reference(T px, T py, const u32* __restrict__ ring_offsets, u32 r0, u32 r1,
                              const T* __restrict__ vx, const T* __restrict__ vy)
{
  bool within  = false;
  bool on_edge = false;
  for (u32 r = r0; r < r1; ++r) {
    u32 const v0 = ring_offsets[r], v1 = ring_offsets[r + 1];
    if (v1 <= v0) continue;  // the reference would read out of bounds
    T bx    = __ldg(vx + v1 - 1);
    T by    = __ldg(vy + v1 - 1);
    bool y0 = by > py;
    for (u32 i = v0; i < v1; ++i) {
      T const ax = __ldg(vx + i), ay = __ldg(vy + i);
      T const run  = fpp<T>::sub(bx, ax);
      T const rise = fpp<T>::sub(by, ay);
      if (float_equal(run, (T)0) && float_equal(rise, (T)0)) continue;  // b NOT advanced (:65-66)
      T const rtp  = fpp<T>::sub(py, ay);
      T const rntp = fpp<T>::sub(px, ax);
      T const u    = fpp<T>::mul(run, rtp);
      T const v    = fpp<T>::mul(rntp, rise);
      if (float_equal(u, v)) {
        T const lo = ax > bx ? bx : ax, hi = ax > bx ? ax : bx;
        if (lo <= px && px <= hi) {
          on_edge = true;
          break;
        }
      }
      bool const y1 = ay > py;
      if (y1 != y0) {
        if ((v < u) != y1) within = !within;
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

// Per-polygon facts computed once per call.
template <typename T>
struct poly_meta {
  T xmin, ymin, xmax, ymax;
  T inv_h;  // n_slabs / (ymax - ymin)
  u32 ring_begin, ring_end;
  u32 safe;      // every coordinate / edge delta "comfy": exact edge skipping is allowed
  u32 n_verts;   // vertices over all rings
  u32 n_slabs;   // y-slab edge index (0 = none: unsafe polygon)
  u32 slab_base; // first slab of this polygon in the global slab arrays
  u32 n_vertical, vert_begin;  // vertical edges (x-only rule of the reference) and their list
};

template <typename T>
struct edge_rec {  // edge b -> a of the reference's walk (b = previous vertex of the ring)
  T ax, ay, bx, by;
};

constexpr u32 kFirstFlag = 0x80000000u;
constexpr u32 kMaxSlabs  = 2048;

// floor((y - ymin) * inv_h) clamped to the polygon's slabs; monotone in y, which is all the
// build/query consistency needs (two overlapping y-intervals map to overlapping slab intervals)
template <typename T>
__device__ __forceinline__ u32 slab_of(T y, const poly_meta<T>& m)
{
  T const f = (y - m.ymin) * m.inv_h;
  if (!(f > (T)0)) return 0u;
  if (f >= (T)m.n_slabs) return m.n_slabs - 1;
  return (u32)f;
}

// one warp per polygon
template <typename T>
__global__ void __launch_bounds__(128)
poly_meta_kernel(const u32* __restrict__ poly_offsets, u32 n_poly,
                 const u32* __restrict__ ring_offsets, u32 n_rings, const T* __restrict__ vx,
                 const T* __restrict__ vy, u32 n_verts, poly_meta<T>* __restrict__ meta)
{
  u32 const p    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  u32 const lane = lane_id();
  if (p >= n_poly) return;
  u32 const r0 = poly_offsets[p], r1 = poly_offsets[p + 1];
  T xmin = fpp<T>::inf(), ymin = fpp<T>::inf(), xmax = -fpp<T>::inf(), ymax = -fpp<T>::inf();
  bool safe = r0 <= r1 && r1 <= n_rings;
  u32 nv = 0, nvert = 0;
  if (safe) {
    for (u32 r = r0; r < r1; ++r) {
      u32 const v0 = ring_offsets[r], v1 = ring_offsets[r + 1];
      if (v1 > n_verts || v0 > v1) {
        safe = false;
        break;
      }
      nv += v1 - v0;
      for (u32 i = v0 + lane; i < v1; i += 32) {
        u32 const pr = i == v0 ? v1 - 1 : i - 1;
        T const ax = vx[i], ay = vy[i], bx = vx[pr], by = vy[pr];
        xmin = fmin(xmin, ax); xmax = fmax(xmax, ax);
        ymin = fmin(ymin, ay); ymax = fmax(ymax, ay);
        safe = safe && comfy(ax) && comfy(ay) && comfy_delta(fpp<T>::sub(bx, ax)) &&
               comfy_delta(fpp<T>::sub(by, ay));
        nvert += (ax == bx && ay != by);
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    nvert += __shfl_xor_sync(0xffffffffu, nvert, o);
  }
  safe = __all_sync(0xffffffffu, safe);
  if (lane == 0) {
    poly_meta<T> m;
    m.xmin = xmin; m.ymin = ymin; m.xmax = xmax; m.ymax = ymax;
    m.ring_begin = r0; m.ring_end = r1; m.safe = safe ? 1u : 0u;
    m.n_verts = nv;
    // about two edges per slab: a quadrant-sized query then touches a handful of entries
    u32 ns = 0;
    if (safe && nv > 0) {
      ns = 1;
      while (ns < kMaxSlabs && ns * 2 < nv) ns <<= 1;
    }
    T const h = ymax - ymin;
    m.n_slabs = ns;
    m.inv_h   = (ns && h > (T)0) ? (T)ns / h : (T)0;
    m.slab_base = 0; m.n_vertical = safe ? nvert : 0u; m.vert_begin = 0;
    meta[p] = m;
  }
}

// exclusive scans over the polygons (slab bases, vertical-list bases); one block.
// totals[0] = total slabs, totals[1] = total vertical edges
template <typename T>
__global__ void __launch_bounds__(1024)
poly_scan_kernel(poly_meta<T>* __restrict__ meta, u32 n_poly, u32* __restrict__ totals)
{
  __shared__ u32 s_a[32], s_b[32];
  __shared__ u32 carry_a, carry_b;
  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_a = carry_b = 0;
  __syncthreads();
  for (u32 base = 0; base < n_poly; base += 1024) {
    u32 const i = base + tid;
    u32 const a = i < n_poly ? meta[i].n_slabs : 0u, b = i < n_poly ? meta[i].n_vertical : 0u;
    u32 const ia = warp_inclusive_scan(a), ib = warp_inclusive_scan(b);
    if (lane == 31) { s_a[warp] = ia; s_b[warp] = ib; }
    __syncthreads();
    u32 wa = 0, wb = 0, ta = 0, tb = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) { wa += s_a[w]; wb += s_b[w]; }
      ta += s_a[w]; tb += s_b[w];
    }
    if (i < n_poly) {
      meta[i].slab_base  = carry_a + wa + ia - a;
      meta[i].vert_begin = carry_b + wb + ib - b;
    }
    __syncthreads();
    if (tid == 0) { carry_a += ta; carry_b += tb; }
    __syncthreads();
  }
  if (tid == 0) { totals[0] = carry_a; totals[1] = carry_b; }
}

// Edge records + y-slab index, one CTA per polygon (a 4096-vertex polygon would otherwise keep one
// warp busy long after the 200-vertex ones are done).  FILL = false: write edge records, count
// slab entries and list vertical edges; FILL = true: write the entries (order inside a slab is
// arbitrary -- crossings XOR and on-edge ORs commute, the result does not depend on it).
constexpr int kSlabBlock = 256;
template <typename T, bool FILL>
__global__ void __launch_bounds__(kSlabBlock)
slab_build_kernel(const poly_meta<T>* __restrict__ meta, u32 n_poly,
                  const u32* __restrict__ ring_offsets, const T* __restrict__ vx,
                  const T* __restrict__ vy, edge_rec<T>* __restrict__ edges,
                  u32* __restrict__ slab_count, const u32* __restrict__ slab_start,
                  u32* __restrict__ entries, u32* __restrict__ vert_edges,
                  u32* __restrict__ vert_cursor)
{
  u32 const p = blockIdx.x;
  if (p >= n_poly) return;
  poly_meta<T> const m = meta[p];
  if (m.n_slabs == 0) return;
  for (u32 r = m.ring_begin; r < m.ring_end; ++r) {
    u32 const v0 = ring_offsets[r], v1 = ring_offsets[r + 1];
    for (u32 i = v0 + threadIdx.x; i < v1; i += kSlabBlock) {
      u32 const pr = i == v0 ? v1 - 1 : i - 1;
      T const ax = vx[i], ay = vy[i], bx = vx[pr], by = vy[pr];
      if (!FILL) edges[i] = edge_rec<T>{ax, ay, bx, by};
      if (ax == bx && ay == by) continue;  // degenerate: skipped by the reference
      T const dl  = fpp<T>::eps() * fmax(fabs(ay), fabs(by));
      u32 const s0 = slab_of<T>(fmin(ay, by) - dl, m), s1 = slab_of<T>(fmax(ay, by) + dl, m);
      for (u32 sl = s0; sl <= s1; ++sl) {
        if (!FILL) {
          atomicAdd(&slab_count[m.slab_base + sl], 1u);
        } else {
          u32 const at = atomicAdd(&slab_count[m.slab_base + sl], 1u);
          entries[slab_start[m.slab_base + sl] + at] = i | (sl == s0 ? kFirstFlag : 0u);
        }
      }
      if (!FILL && ax == bx) vert_edges[m.vert_begin + atomicAdd(&vert_cursor[p], 1u)] = i;
    }
  }
}

template <typename T>
struct edge_index;
template <typename T>
struct polygon_index;  // owning builder, defined after the kernels

template <typename T>
struct edge_index {  // device view of the per-call polygon edge index
  const edge_rec<T>* edges;
  const u32* slab_start;
  const u32* entries;
  const u32* vert_edges;
};

// Where the coordinates of point `id` live.  Single GPU: two columns.  Multi-GPU: the points of a
// key range stay on the ranks that own them and only (key, id) pairs were exchanged; `id` is then
// a GLOBAL id and the coordinates are read from the owning rank's columns through peer pointers
// (NVLink) -- only for the few points that need the exact predicate.
template <typename T, bool SEG>
struct coord_source;
template <typename T>
struct coord_source<T, false> {
  const T* x;
  const T* y;
  u32 n_ids;  // ids are valid below this
  __device__ __forceinline__ void load(u32 id, T& ox, T& oy) const
  {
    ox = __ldg(x + id);
    oy = __ldg(y + id);
  }
};
template <typename T>
struct coord_source<T, true> {
  u32 n_ids;
  int n_seg;  // ids first_id[s] .. first_id[s+1]-1 live in segment s
  u32 first_id[BSJ_MAX_RANKS + 1];
  const T* sx[BSJ_MAX_RANKS];
  const T* sy[BSJ_MAX_RANKS];
  __device__ __forceinline__ void load(u32 id, T& ox, T& oy) const
  {
    int sgm = 0;
    while (sgm + 1 < n_seg && id >= first_id[sgm + 1]) ++sgm;
    u32 const j = id - first_id[sgm];
    ox = sx[sgm][j];  // possibly a peer GPU's memory: plain (coherent) loads
    oy = sy[sgm][j];
  }
};

// ---------------------------------------------------------------------------------------------
// Whole-quadrant classification from the cell rectangle (optional bsj_grid hint).
//
// A quadrant at (level, key) holds exactly the points whose cell index, floor((x-min)/scale)
// computed in T, falls in its key range; its points therefore lie in the cell rectangle widened by
// a rounding margin.  If no polygon edge comes near that rectangle (edge bounding box, widened by
// the 4-ULP on-edge tolerance, disjoint from it) and no vertical edge has its x inside the
// rectangle's x-range (the reference rejects such points at ANY y), then every crossing decision
// is sign-certain and identical for all points of the rectangle: the quadrant is uniformly inside
// or outside and the answer is the predicate of the rectangle's centre.  No point is read.
// ---------------------------------------------------------------------------------------------
struct grid_info {
  int valid;
  int max_depth;
  int has_oob;
  double min_x, min_y, scale;
  double margin_x, margin_y;  // rounding margin of the point -> cell assignment
  const u32* sorted_keys;     // Morton keys of the points in sorted order (optional)
};

constexpr int kClsOutside = 0, kClsInside = 1, kClsBoundary = 2;

__device__ __forceinline__ u32 undilate16p(u32 v)
{
  v &= 0x55555555u;
  v = (v | (v >> 1)) & 0x33333333u;
  v = (v | (v >> 2)) & 0x0F0F0F0Fu;
  v = (v | (v >> 4)) & 0x00FF00FFu;
  v = (v | (v >> 8)) & 0x0000FFFFu;
  return v;
}

// What the edge pass of one (quadrant, polygon) test needs.
template <typename T>
struct quad_query {
  double ex0, ex1, ey0, ey1;   // the cell rectangle widened by the rounding margin
  T cx, cy;                    // its centre
  u32 kbeg, kmid, kend;        // slab entries [kbeg, kend); beyond kmid only those flagged "first"
  u32 n_vertical, vert_begin;  // the polygon's vertical edges
};

// Part 1 (per thread, no cooperation): the rectangle, the exact box rejection, the slab range.
// Returns kClsOutside / kClsBoundary when that already decides, -1 when the edges must be looked at.
template <typename T>
__device__ __forceinline__ int quadrant_setup(const grid_info& g, u32 key, u32 level,
                                              const poly_meta<T>& m, const edge_index<T>& ix,
                                              quad_query<T>& q)
{
  int const sh = g.max_depth - 1 - (int)level;
  if (!g.valid || !m.safe || sh < 0 || m.n_slabs == 0) return kClsBoundary;
  // the last cell also receives every out-of-box point, whatever its coordinates
  if (g.has_oob && level < 16 && key == ((1u << (2 * (level + 1))) - 1u)) return kClsBoundary;
  double const ls = g.scale * (double)(1u << sh);
  double const kx = (double)undilate16p(key), ky = (double)undilate16p(key >> 1);
  double const x0 = g.min_x + kx * ls, x1 = g.min_x + (kx + 1.0) * ls;
  double const y0 = g.min_y + ky * ls, y1 = g.min_y + (ky + 1.0) * ls;
  q.ex0 = x0 - g.margin_x; q.ex1 = x1 + g.margin_x;
  q.ey0 = y0 - g.margin_y; q.ey1 = y1 + g.margin_y;
  double const eps = (double)fpp<T>::eps();
  {
    double const dx = eps * fmax(fabs((double)m.xmin), fabs((double)m.xmax));
    double const dy = eps * fmax(fabs((double)m.ymin), fabs((double)m.ymax));
    if (q.ex1 < (double)m.xmin - dx || q.ex0 > (double)m.xmax + dx ||
        q.ey1 < (double)m.ymin - dy || q.ey0 > (double)m.ymax + dy)
      return kClsOutside;
  }
  q.cx = (T)(0.5 * (x0 + x1));
  q.cy = (T)(0.5 * (y0 + y1));
  // edges whose (tolerance-widened) y-range meets the rectangle's: slabs q0..q1 of the index.
  // An edge listed in several slabs is taken once: in slab q0, or where it is flagged "first".
  u32 const q0 = slab_of<T>((T)q.ey0, m), q1 = slab_of<T>((T)q.ey1, m);
  q.kbeg = __ldg(ix.slab_start + m.slab_base + q0);
  q.kmid = __ldg(ix.slab_start + m.slab_base + q0 + 1);
  q.kend = __ldg(ix.slab_start + m.slab_base + q1 + 1);
  q.n_vertical = m.n_vertical;
  q.vert_begin = m.vert_begin;
  return -1;
}

// Part 2 (cooperative over groups of W consecutive lanes, `q` uniform inside a group; every lane
// of the warp must call it, `active` false for groups with nothing to do): the group's lanes
// stride the candidate edges; all lanes of a group return the same class.
template <typename T, int W>
__device__ __forceinline__ int quadrant_edges(const quad_query<T>& q, bool active,
                                              const edge_index<T>& ix)
{
  double const eps = (double)fpp<T>::eps();
  u32 const lane = lane_id(), sub = lane % W, gsh = (lane / W) * W;
  u32 const gm   = W == 32 ? 0xffffffffu : ((1u << (W & 31)) - 1u);
  bool near  = false;
  u32 cross  = 0;
  if (active) {
    for (u32 k = q.kbeg + sub; k < q.kend; k += W) {
      u32 const ent = __ldg(ix.entries + k);
      if (k >= q.kmid && !(ent & kFirstFlag)) continue;
      edge_rec<T> const e = ix.edges[ent & ~kFirstFlag];
      double const d = eps * fmax(fmax(fabs((double)e.ax), fabs((double)e.bx)),
                                  fmax(fabs((double)e.ay), fabs((double)e.by)));
      double const lx = fmin((double)e.ax, (double)e.bx) - d, hx = fmax((double)e.ax, (double)e.bx) + d;
      if (hx < q.ex0) continue;  // wholly left of the rectangle: cannot touch it, never toggles
      bool const f1 = e.ay > q.cy, f0 = e.by > q.cy;
      if (lx > q.ex1) {  // wholly right: cannot touch it, toggles exactly on a y-straddle
        cross ^= (u32)(f1 != f0);
        continue;
      }
      double const ly = fmin((double)e.ay, (double)e.by) - d, hy = fmax((double)e.ay, (double)e.by) + d;
      if (!(hy < q.ey0 || ly > q.ey1)) {
        // the edge's (tolerance-widened) box meets the rectangle: does its LINE pass through it?
        // f = (v - u) of the reference is linear, so |f(centre)| <= |rise| hw + |run| hh bounds it
        // over the rectangle; the relative allowance covers the 4-ULP on-edge band and the
        // rounding of the reference's products (same test as pip_eval_cells_kernel's `touch`)
        double const eax = (double)e.ax, eay = (double)e.ay;
        double const run = (double)e.bx - eax, rise = (double)e.by - eay;
        double const hw = 0.5 * (q.ex1 - q.ex0), hh = 0.5 * (q.ey1 - q.ey0);
        double const ddx = 0.5 * (q.ex0 + q.ex1) - eax, ddy = 0.5 * (q.ey0 + q.ey1) - eay;
        double const f   = ddx * rise - run * ddy;
        double const allow = sizeof(T) == 4 ? 2e-6 : 1e-9;
        double const thr = (fabs(rise) * hw + fabs(run) * hh) * 1.000001 +
                           allow * ((fabs(ddx) + hw) * fabs(rise) + fabs(run) * (fabs(ddy) + hh));
        near = near || fabs(f) <= thr;
      }
      if (f1 != f0) {
        T const u = fpp<T>::mul(fpp<T>::sub(e.bx, e.ax), fpp<T>::sub(q.cy, e.ay));
        T const v = fpp<T>::mul(fpp<T>::sub(q.cx, e.ax), fpp<T>::sub(e.by, e.ay));
        cross ^= (u32)((v < u) != f1);
      }
    }
    // vertical edges reject points with the same x at ANY y
    for (u32 k = sub; k < q.n_vertical; k += W) {
      T const ax = ix.edges[__ldg(ix.vert_edges + q.vert_begin + k)].ax;
      near = near || ((double)ax >= q.ex0 && (double)ax <= q.ex1);
    }
  }
  u32 const nb = (__ballot_sync(0xffffffffu, near) >> gsh) & gm;
  u32 const cb = (__ballot_sync(0xffffffffu, cross & 1u) >> gsh) & gm;
  if (nb) return kClsBoundary;
  return (__popc(cb) & 1) ? kClsInside : kClsOutside;
}

// warp-cooperative, warp-uniform arguments; all lanes return the same class
template <typename T>
__device__ int classify_quadrant(const grid_info& g, u32 key, u32 level, const poly_meta<T>& m,
                                 const edge_index<T>& ix)
{
  quad_query<T> q;
  int const c = quadrant_setup<T>(g, key, level, m, ix, q);
  return c >= 0 ? c : quadrant_edges<T, 32>(q, true, ix);
}

// ---------------------------------------------------------------------------------------------
// pair bookkeeping
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pair_prep_kernel(const u32* __restrict__ pair_quad, u32 n_pairs, const u32* __restrict__ length,
                 const u32* __restrict__ offset, u32 num_nodes, u32* __restrict__ words,
                 u32* __restrict__ pair_len, u32* __restrict__ pair_off)
{
  u32 const j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pairs) return;
  u32 const q   = pair_quad[j];
  u32 const len = q < num_nodes ? length[q] : 0u;
  pair_len[j]   = len;
  pair_off[j]   = q < num_nodes ? offset[q] : 0u;
  words[j]      = len / 32 + ((len & 31) != 0);
}

__global__ void __launch_bounds__(256)
narrow_u64_kernel(const u64* __restrict__ in, u32* __restrict__ out, u32 n)
{
  u32 const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (u32)in[i];
}

// ---------------------------------------------------------------------------------------------
// evaluation: one warp per (pair, point tile) unit
// ---------------------------------------------------------------------------------------------
constexpr int kPipWarps = 4;
constexpr u32 kEvalSlices = 32;
constexpr int kPPL      = 8;          // points per lane
constexpr int kPipTile  = 32 * kPPL;  // points per tile

// exact predicate of one point through the y-slab index (defined with the bitmask kernel below)
template <typename T>
__device__ bool pip_indexed(T px, T py, const poly_meta<T>& m, const edge_index<T>& ix);

// Stage 1 of the evaluation: one warp per (polygon, quadrant) pair decides whole quadrants from
// their cell rectangle (classify_quadrant).  Pairs are independent, so the grid keeps every SM
// full of warps and the dependent index/edge loads of one pair hide behind the others.  Quadrants
// that need their points are appended (once) to a work list for stage 2.
template <typename T>
__global__ void __launch_bounds__(256)
pip_classify_kernel(const u32* __restrict__ pair_poly, const u32* __restrict__ pair_quad,
                    u32 n_pairs, const u32* __restrict__ length,
                    const u32* __restrict__ offset, u32 num_nodes,
                    u32 n_points, const poly_meta<T>* __restrict__ meta, u32 n_poly,
                    int force_reference, const u32* __restrict__ node_key,
                    const u8* __restrict__ node_level, grid_info grid, edge_index<T> ix,
                    u8* __restrict__ cls, u32* __restrict__ hits,
                    u32* __restrict__ run_list, u32* __restrict__ tile_list,
                    u32* __restrict__ run_count, u32 tile_points, u32 list_capacity, u32 group)
{
  // A warp takes `group` (<= 32) consecutive pairs at a time.  Lane l does everything of pair l that needs no
  // cooperation -- the chain pair -> quadrant -> (length, offset, key, level), polygon -> record
  // -> rectangle test -> slab range -- so that those dependent loads run 32 pairs wide instead
  // of one pair per warp; only the pairs whose edges must be looked at are then handed to the
  // whole warp, one after the other, through shuffles (two dependent loads left per pair).
  u32 const lane   = lane_id();
  u32 const warps  = (gridDim.x * blockDim.x) >> 5;
  u64 const groups = ((u64)n_pairs + group - 1) / group;
  for (u64 grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < groups; grp += warps) {
    u64 const j = grp * group + lane;
    bool const mine = lane < group && j < n_pairs;
    int c       = kClsOutside;
    u32 nvalid = 0, qlen = 0;
    quad_query<T> q{};
    bool need = false;
    if (mine) {
      u32 const quad = pair_quad[j];
      if (quad < num_nodes) {
        u32 const len = length[quad], off = offset[quad];
        qlen          = len;
        nvalid        = off < n_points ? min(len, n_points - off) : 0u;
        u32 const poly = pair_poly[j];
        if (len != 0 && poly < n_poly) {
          if (grid.valid && force_reference != 1) {
            c    = quadrant_setup<T>(grid, node_key[quad], node_level[quad], meta[poly], ix, q);
            need = c < 0;
          } else {
            c = kClsBoundary;
          }
        }
      }
    }
    // four pairs at a time, eight lanes each: a pair typically has a handful of candidate
    // edges, and four independent entry -> edge load chains are in flight per warp
    // (a pair with MANY candidate edges -- a large quadrant of a sparse tree against a detailed
    // polygon -- gets the whole warp instead)
    constexpr int W = 8, NG = 32 / W;
    bool const wide = need && (q.kend - q.kbeg > 64u || q.n_vertical > 64u);
    u32 big = __ballot_sync(0xffffffffu, wide);
    while (big) {
      int const src = __ffs(big) - 1;
      big &= big - 1;
      quad_query<T> b;
      b.ex0 = __shfl_sync(0xffffffffu, q.ex0, src); b.ex1 = __shfl_sync(0xffffffffu, q.ex1, src);
      b.ey0 = __shfl_sync(0xffffffffu, q.ey0, src); b.ey1 = __shfl_sync(0xffffffffu, q.ey1, src);
      b.cx  = __shfl_sync(0xffffffffu, q.cx, src);  b.cy  = __shfl_sync(0xffffffffu, q.cy, src);
      b.kbeg = __shfl_sync(0xffffffffu, q.kbeg, src);
      b.kmid = __shfl_sync(0xffffffffu, q.kmid, src);
      b.kend = __shfl_sync(0xffffffffu, q.kend, src);
      b.n_vertical = __shfl_sync(0xffffffffu, q.n_vertical, src);
      b.vert_begin = __shfl_sync(0xffffffffu, q.vert_begin, src);
      int const r = quadrant_edges<T, 32>(b, true, ix);
      if ((int)lane == src) c = r;
    }
    u32 todo = __ballot_sync(0xffffffffu, need && !wide);
    u32 const lt = lanemask_lt();
    while (todo) {
      u32 const grp_id = lane / W;
      u32 const pick   = __fns(todo, 0, (int)grp_id + 1);  // lane of the group's pair, or ~0
      bool const active = pick < 32u;
      int const src     = active ? (int)pick : 0;
      quad_query<T> b;
      b.ex0 = __shfl_sync(0xffffffffu, q.ex0, src); b.ex1 = __shfl_sync(0xffffffffu, q.ex1, src);
      b.ey0 = __shfl_sync(0xffffffffu, q.ey0, src); b.ey1 = __shfl_sync(0xffffffffu, q.ey1, src);
      b.cx  = __shfl_sync(0xffffffffu, q.cx, src);  b.cy  = __shfl_sync(0xffffffffu, q.cy, src);
      b.kbeg = __shfl_sync(0xffffffffu, q.kbeg, src);
      b.kmid = __shfl_sync(0xffffffffu, q.kmid, src);
      b.kend = __shfl_sync(0xffffffffu, q.kend, src);
      b.n_vertical = __shfl_sync(0xffffffffu, q.n_vertical, src);
      b.vert_begin = __shfl_sync(0xffffffffu, q.vert_begin, src);
      int const r = quadrant_edges<T, W>(b, active, ix);
      // the owner of the k-th pending pair (k < NG) takes the result of lane group k
      u32 const k  = __popc(todo & lt);
      int const rr = __shfl_sync(0xffffffffu, r, (int)(min(k, (u32)NG - 1u) * W));
      if (((todo >> lane) & 1u) && k < (u32)NG) c = rr;
#pragma unroll
      for (int t = 0; t < NG; ++t) todo &= todo - 1;
    }
    // Stage 2 works on (pair, point tile) units: bounded work per unit (one polygon against one
    // tile) whatever the data looks like -- a dense bottom-level leaf that cannot split (many
    // tiles) or a large sparse quadrant under hundreds of polygon boxes (many pairs) spreads over
    // many warps instead of serialising in one.
    u32 n_tiles = 0;
    if (mine) {
      cls[j]  = (u8)c;
      hits[j] = c == kClsInside ? nvalid : 0u;  // boundary pairs: stage 2 adds its counts
      if (c == kClsBoundary) n_tiles = max(1u, qlen / tile_points + (qlen % tile_points != 0));
    }
    // one atomic per group reserves the units of all its boundary pairs
    u32 const incl = warp_inclusive_scan(n_tiles);
    u32 const tot  = __shfl_sync(0xffffffffu, incl, 31);
    if (tot == 0) continue;
    u32 base = 0;
    if (lane == 31) base = atomicAdd(run_count, tot);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - n_tiles;
    u32 bnd = __ballot_sync(0xffffffffu, n_tiles != 0);
    while (bnd) {
      int const src = __ffs(bnd) - 1;
      bnd &= bnd - 1;
      u32 const nt = __shfl_sync(0xffffffffu, n_tiles, src);
      u32 const b0 = __shfl_sync(0xffffffffu, base, src);
      u32 const jj = (u32)(grp * group) + (u32)src;
      for (u32 t = lane; t < nt; t += 32)
        if (b0 + t < list_capacity) {
          run_list[b0 + t]  = jj;
          tile_list[b0 + t] = t;
        }
    }
  }
}

// Stage 2 without the sorted keys: one warp per (pair, tile) unit gathers the tile's points and
// tests them against the pair's polygon with exact edge skipping.
template <typename T, bool SEG>
__global__ void __launch_bounds__(kPipWarps * 32)
pip_eval_kernel(const u32* __restrict__ pair_poly, const u32* __restrict__ pair_quad,
                const u32* __restrict__ run_list,
                const u32* __restrict__ tile_list, u32 list_capacity,
                const u32* __restrict__ run_count, const u32* __restrict__ length,
                const u32* __restrict__ offset, const u32* __restrict__ point_indices,
                u32 n_points, coord_source<T, SEG> const pts,
                const poly_meta<T>* __restrict__ meta, u32 n_poly,
                const u32* __restrict__ ring_offsets, const T* __restrict__ vx,
                const T* __restrict__ vy, const u64* __restrict__ wbase,
                u32* __restrict__ mask_words, u32* __restrict__ hits, u32* __restrict__ ticket,
                u32 n_slices, int force_reference, const u8* __restrict__ cls,
                edge_index<T> ix)
{
  u32 const lane   = lane_id();
  u32 const n_list = min(*run_count, list_capacity);
  // units are drawn from n_slices <= kEvalSlices counters (a single counter serialises the whole
  // grid's atomics in L2)
  u32 const slice = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) % n_slices;
  ticket += slice;

  while (true) {
    u32 t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    // slice s owns units s, s + n_slices, s + 2 n_slices, ...: neighbouring units (similar cost:
    // they come from neighbouring pairs) spread over all slices
    u64 const slot64 = (u64)t * n_slices + slice;
    if (slot64 >= n_list) break;
    u32 const slot = (u32)slot64;
    u32 const j0 = run_list[slot], j1 = j0 + 1;  // one (pair, tile) unit per slot
    u32 const quad = pair_quad[j0];
    u32 const len = length[quad], off = offset[quad];

    {
      u32 const base = tile_list[slot] * kPipTile;
      // ---- gather this tile's points once (quadtree_point_in_polygon.cuh:66,170)
      u32 idx[kPPL];
      u32 valid = 0;
#pragma unroll
      for (int i = 0; i < kPPL; ++i) {
        u32 const l = base + i * 32 + lane;
        bool const ok = l < len && off + l < n_points;
        valid |= (u32)ok << i;
        idx[i] = ok ? __ldcs(point_indices + off + l) : 0xFFFFFFFFu;
      }
      T x[kPPL], y[kPPL];
#pragma unroll
      for (int i = 0; i < kPPL; ++i) {
        bool const ok = idx[i] < pts.n_ids;
        if (!ok) valid &= ~(1u << i);
        x[i] = (T)0;
        y[i] = (T)0;
        if (ok) pts.load(idx[i], x[i], y[i]);
      }
      // invalid slots mirror the tile's first point so they neither widen the bbox nor trap
      T const fx = __shfl_sync(0xffffffffu, x[0], 0), fy = __shfl_sync(0xffffffffu, y[0], 0);
      T tx0 = fx, tx1 = fx, ty0 = fy, ty1 = fy;
      bool pts_ok = true;
#pragma unroll
      for (int i = 0; i < kPPL; ++i) {
        if (!((valid >> i) & 1u)) {
          x[i] = fx;
          y[i] = fy;
        }
        tx0 = fmin(tx0, x[i]); tx1 = fmax(tx1, x[i]);
        ty0 = fmin(ty0, y[i]); ty1 = fmax(ty1, y[i]);
        pts_ok = pts_ok && comfy(x[i]) && comfy(y[i]);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        tx0 = fmin(tx0, __shfl_xor_sync(0xffffffffu, tx0, o));
        ty0 = fmin(ty0, __shfl_xor_sync(0xffffffffu, ty0, o));
        tx1 = fmax(tx1, __shfl_xor_sync(0xffffffffu, tx1, o));
        ty1 = fmax(ty1, __shfl_xor_sync(0xffffffffu, ty1, o));
      }
      bool const tile_safe = __all_sync(0xffffffffu, pts_ok) && force_reference != 1;
      u32 const tile_words = (min(len - base, (u32)kPipTile) + 31) / 32;  // <= 8

      for (u32 j = j0; j < j1; ++j) {
        if (cls[j] != kClsBoundary) continue;  // settled from the cell rectangle by stage 1
        u32 const poly = pair_poly[j];
        u32 inside     = 0;  // bit i: point i of this lane is inside
        if (poly < n_poly) {
          poly_meta<T> const m = meta[poly];
          if (tile_safe && m.safe) {
            T const mx = fpp<T>::eps() * fmax(fabs(m.xmin), fabs(m.xmax));
            bool const miss = ty1 < m.ymin || ty0 >= m.ymax || tx1 < m.xmin - mx ||
                              tx0 > m.xmax + mx;
            if (!miss) {
              u32 within = 0, onedge = 0;
              // evaluate one edge (held by lane `src`) against this lane's points, with the
              // reference's own arithmetic (is_point_in_polygon.cuh:60-96)
              auto eval_edges = [&](u32 em, T ax, T ay, T bx, T by) {
                while (em) {
                  int const src = __ffs(em) - 1;
                  em &= em - 1;
                  T const eax = __shfl_sync(0xffffffffu, ax, src);
                  T const eay = __shfl_sync(0xffffffffu, ay, src);
                  T const ebx = __shfl_sync(0xffffffffu, bx, src);
                  T const eby = __shfl_sync(0xffffffffu, by, src);
                  T const run  = fpp<T>::sub(ebx, eax);
                  T const rise = fpp<T>::sub(eby, eay);
                  T const lo = fmin(eax, ebx), hi = fmax(eax, ebx);
#pragma unroll
                  for (int i = 0; i < kPPL; ++i) {
                    T const rtp  = fpp<T>::sub(y[i], eay);
                    T const rntp = fpp<T>::sub(x[i], eax);
                    T const u    = fpp<T>::mul(run, rtp);
                    T const v    = fpp<T>::mul(rntp, rise);
                    if (lo <= x[i] && x[i] <= hi) {
                      if (float_equal(u, v)) onedge |= 1u << i;
                    }
                    bool const y1 = eay > y[i], y0 = eby > y[i];
                    bool const cross = (y1 != y0) && ((v < u) != y1);
                    within ^= (u32)cross << i;
                  }
                }
              };
              u32 kbeg = 0, kmid = 0, kend = 0;
              if (m.n_slabs) {
                // candidate edges from the y-slab index (slabs covering the tile's y-range)
                u32 const q0 = slab_of<T>(ty0, m), q1 = slab_of<T>(ty1, m);
                kbeg = __ldg(ix.slab_start + m.slab_base + q0);
                kmid = __ldg(ix.slab_start + m.slab_base + q0 + 1);
                kend = __ldg(ix.slab_start + m.slab_base + q1 + 1);
              }
              if (m.n_slabs && kend - kbeg > 64u) {
                // a tile spanning most of the polygon's height (large, sparse quadrant): each
                // point through its own slab instead of every edge over the whole warp
#pragma unroll
                for (int i = 0; i < kPPL; ++i)
                  if ((valid >> i) & 1u) within |= (u32)pip_indexed<T>(x[i], y[i], m, ix) << i;
              } else if (m.n_slabs) {
                for (u32 k0 = kbeg; k0 < kend; k0 += 32) {
                  u32 const k = k0 + lane;
                  T ax = 0, ay = 0, bx = 0, by = 0;
                  bool rel = false;
                  if (k < kend) {
                    u32 const ent = __ldg(ix.entries + k);
                    if (k < kmid || (ent & kFirstFlag)) {
                      edge_rec<T> const e = ix.edges[ent & ~kFirstFlag];
                      ax = e.ax; ay = e.ay; bx = e.bx; by = e.by;
                      T const ylo = fmin(ay, by), yhi = fmax(ay, by);
                      T const dl  = fpp<T>::eps() * fmax(fabs(ay), fabs(by));
                      rel = ty1 >= ylo - dl && ty0 <= yhi + dl;
                    }
                  }
                  eval_edges(__ballot_sync(0xffffffffu, rel), ax, ay, bx, by);
                }
                // vertical edges inside the tile's x-range whose y-range misses the tile
                for (u32 k0 = 0; k0 < m.n_vertical; k0 += 32) {
                  u32 const k = k0 + lane;
                  T ax = 0, ay = 0, bx = 0, by = 0;
                  bool rel = false;
                  if (k < m.n_vertical) {
                    edge_rec<T> const e = ix.edges[__ldg(ix.vert_edges + m.vert_begin + k)];
                    ax = e.ax; ay = e.ay; bx = e.bx; by = e.by;
                    T const ylo = fmin(ay, by), yhi = fmax(ay, by);
                    T const dl  = fpp<T>::eps() * fmax(fabs(ay), fabs(by));
                    bool const yrel = ty1 >= ylo - dl && ty0 <= yhi + dl;  // taken above
                    rel = !yrel && tx0 <= ax && ax <= tx1;
                  }
                  eval_edges(__ballot_sync(0xffffffffu, rel), ax, ay, bx, by);
                }
              } else {
                for (u32 ring = m.ring_begin; ring < m.ring_end; ++ring) {
                  u32 const v0 = ring_offsets[ring], v1 = ring_offsets[ring + 1];
                  u32 const nv = v1 - v0;
                  for (u32 c0 = 0; c0 < nv; c0 += 32) {
                    u32 const e    = c0 + lane;
                    bool const has = e < nv;
                    T ax = 0, ay = 0, bx = 0, by = 0;
                    bool rel = false;
                    if (has) {
                      u32 const pr = e == 0 ? nv - 1 : e - 1;
                      ax = __ldg(vx + v0 + e);  ay = __ldg(vy + v0 + e);
                      bx = __ldg(vx + v0 + pr); by = __ldg(vy + v0 + pr);
                      T const ylo = fmin(ay, by), yhi = fmax(ay, by);
                      T const dl  = fpp<T>::eps() * fmax(fabs(ay), fabs(by));
                      bool const yrel = ty1 >= ylo - dl && ty0 <= yhi + dl;
                      bool const vert = ax == bx && tx0 <= ax && ax <= tx1;
                      rel = !(ax == bx && ay == by) && (yrel || vert);
                    }
                    eval_edges(__ballot_sync(0xffffffffu, rel), ax, ay, bx, by);
                  }
                }
              }
              inside = within & ~onedge;
            }
          } else {
#pragma unroll
            for (int i = 0; i < kPPL; ++i)  // unrolled: x[]/y[] must stay in registers
              if ((valid >> i) & 1u)
                inside |= (u32)pip_reference<T>(x[i], y[i], ring_offsets, m.ring_begin,
                                                m.ring_end, vx, vy)
                          << i;
          }
        }
        inside &= valid;
        // ---- ballot words in (pair, local point) order: word i covers points base+32i .. +31
        u32 mine = 0, cnt = 0;
#pragma unroll
        for (int i = 0; i < kPPL; ++i) {
          u32 const w = __ballot_sync(0xffffffffu, (inside >> i) & 1u);
          if (lane == (u32)i) mine = w;
          cnt += __popc(w);
        }
        if (lane < tile_words) mask_words[wbase[j] + base / 32 + lane] = mine;
        if (lane == 0 && cnt) atomicAdd(&hits[j], cnt);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stage 2, cell-centre form (needs the sorted Morton keys of the points, part of the bsj_grid
// hint).  A point's key names its finest cell, a square of side `scale` that contains the point
// (plus the rounding margin of the key computation).  If no edge of the polygon passes through
// that (margin-widened) square and no vertical edge has its x inside the square's x-range, the
// point and the square's centre lie in the same face of the polygon's edge arrangement, every
// crossing comparison of the reference is sign-certain for both, and the reference's answer for
// the point equals the plain crossing parity of the CENTRE -- which needs the 4-byte key, read
// coalesced, instead of two random 8-byte gathers that cost ~115 B of HBM traffic each.  Only the
// points whose square is touched by an edge (a fraction of a percent) are gathered and put
// through the reference's own arithmetic.  Out-of-box points share the last key whatever their
// coordinates: they always take the exact path.
// ---------------------------------------------------------------------------------------------
constexpr int kCellPPL  = 4;
constexpr int kCellTile = 32 * kCellPPL;
// (tile, polygon) combinations with more candidate edges than this are evaluated point by point
// through the y-slab index instead of edge by edge over the whole warp (a tile spanning the whole
// height of a very detailed polygon).  Measured with the right-side shortcut of the edge-by-edge
// form in place (ms of this kernel at thresholds 64 / 1024 / never): configs[3] at 125 M points
// 4.04 / 2.97 / 3.30, at 1 G 8.0 / 7.6 / 7.8; configs[4] at 250 M 63 / 23.9 / 22.6; configs[1]
// 0.59 either way.
constexpr u32 kCoopEdges = 1024;


template <typename T, bool SEG>
__global__ void __launch_bounds__(kPipWarps * 32, 4)
pip_eval_cells_kernel(const u32* __restrict__ pair_poly, const u32* __restrict__ pair_quad,
                      const u32* __restrict__ run_list,
                      const u32* __restrict__ tile_list, u32 list_capacity,
                      const u32* __restrict__ run_count, const u32* __restrict__ length,
                      const u32* __restrict__ offset, const u32* __restrict__ point_indices,
                      u32 n_points, coord_source<T, SEG> const pts,
                      const poly_meta<T>* __restrict__ meta, u32 n_poly,
                      const u32* __restrict__ ring_offsets, const T* __restrict__ vx,
                      const T* __restrict__ vy, const u64* __restrict__ wbase,
                      u32* __restrict__ mask_words, u32* __restrict__ hits,
                      u32* __restrict__ ticket, u32 n_slices, const u8* __restrict__ cls,
                      edge_index<T> ix,
                      grid_info grid, const u32* __restrict__ sorted_keys, u32 coop_edges)
{
  u32 const lane   = lane_id();
  u32 const n_list = min(*run_count, list_capacity);
  double const hw  = 0.5 * grid.scale + grid.margin_x;  // half extents of a widened finest cell
  double const hh  = 0.5 * grid.scale + grid.margin_y;
  u32 const oob_key = grid.max_depth >= 16 ? 0xFFFFFFFFu : ((1u << (2 * grid.max_depth)) - 1u);
  double const eps  = (double)fpp<T>::eps();
  // relative allowance on the centre's line function f = v - u: the reference evaluates u and v
  // in T (products rounded separately) and calls them equal within 4 ULP -- 5e-7 |u| for float
  double const allow = sizeof(T) == 4 ? 2e-6 : 1e-9;
  // units are drawn from n_slices <= kEvalSlices counters (a single counter serialises the whole
  // grid's atomics in L2)
  u32 const slice = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) % n_slices;
  ticket += slice;

  while (true) {
    u32 t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    // slice s owns units s, s + n_slices, s + 2 n_slices, ...: neighbouring units (similar cost:
    // they come from neighbouring pairs) spread over all slices
    u64 const slot64 = (u64)t * n_slices + slice;
    if (slot64 >= n_list) break;
    u32 const slot = (u32)slot64;
    u32 const j0 = run_list[slot], j1 = j0 + 1;  // one (pair, tile) unit per slot
    u32 const quad = pair_quad[j0];
    u32 const len = length[quad], off = offset[quad];

    {
      u32 const base = tile_list[slot] * kCellTile;
      u32 valid = 0, oob = 0;
      double cx[kCellPPL], cy[kCellPPL];
      u32 ixmin = 0xFFFFFFFFu, ixmax = 0u, iymin = 0xFFFFFFFFu, iymax = 0u;
#pragma unroll
      for (int i = 0; i < kCellPPL; ++i) {
        u32 const l   = base + i * 32 + lane;
        bool const ok = l < len && off + l < n_points;
        u32 const k   = ok ? __ldcs(sorted_keys + off + l) : 0u;
        valid |= (u32)ok << i;
        oob |= (u32)(ok && grid.has_oob && k == oob_key) << i;
        u32 const ix = undilate16p(k), iy = undilate16p(k >> 1);
        if (ok) {
          ixmin = min(ixmin, ix); ixmax = max(ixmax, ix);
          iymin = min(iymin, iy); iymax = max(iymax, iy);
        }
        cx[i] = grid.min_x + ((double)ix + 0.5) * grid.scale;
        cy[i] = grid.min_y + ((double)iy + 0.5) * grid.scale;
      }
      // tile extent (cell squares of the valid points): the centre is monotone in the cell
      // index, so the extremes come from four integer warp reductions
      ixmin = __reduce_min_sync(0xffffffffu, ixmin); ixmax = __reduce_max_sync(0xffffffffu, ixmax);
      iymin = __reduce_min_sync(0xffffffffu, iymin); iymax = __reduce_max_sync(0xffffffffu, iymax);
      bool const any_pt = ixmin <= ixmax;
      double const tx0 = any_pt ? grid.min_x + ((double)ixmin + 0.5) * grid.scale - hw : 1e300;
      double const tx1 = any_pt ? grid.min_x + ((double)ixmax + 0.5) * grid.scale + hw : -1e300;
      double const ty0 = any_pt ? grid.min_y + ((double)iymin + 0.5) * grid.scale - hh : 1e300;
      double const ty1 = any_pt ? grid.min_y + ((double)iymax + 0.5) * grid.scale + hh : -1e300;
      // real coordinates, gathered lazily and only for the points that need them
      T xr[kCellPPL], yr[kCellPPL];
      u32 loaded = 0, unsafe_pt = 0;
      auto load_points = [&](u32 want) {
        want &= valid & ~loaded;
#pragma unroll
        for (int i = 0; i < kCellPPL; ++i)
          if ((want >> i) & 1u) {
            u32 const id = __ldg(point_indices + off + base + i * 32 + lane);
            if (id < pts.n_ids) {
              pts.load(id, xr[i], yr[i]);
              if (!(comfy(xr[i]) && comfy(yr[i]))) unsafe_pt |= 1u << i;
            } else {
              valid &= ~(1u << i);
            }
          }
        loaded |= want;
      };
      u32 const tile_words = (min(len - base, (u32)kCellTile) + 31) / 32;  // <= kCellPPL

      for (u32 j = j0; j < j1; ++j) {
        if (cls[j] != kClsBoundary) continue;
        u32 const poly = pair_poly[j];
        u32 inside     = 0;
        if (poly < n_poly) {
          poly_meta<T> const m = meta[poly];
          if (!(m.safe && m.n_slabs)) {
            // unusual polygon (NaN/Inf/extreme magnitudes): the literal reference loop
            load_points(0xFFFFFFFFu);
#pragma unroll
            for (int i = 0; i < kCellPPL; ++i)
              if ((valid >> i) & 1u)
                inside |= (u32)pip_reference<T>(xr[i], yr[i], ring_offsets, m.ring_begin,
                                                m.ring_end, vx, vy) << i;
          } else {
            double const dx = eps * fmax(fabs((double)m.xmin), fabs((double)m.xmax));
            double const dy = eps * fmax(fabs((double)m.ymin), fabs((double)m.ymax));
            bool const miss = tx1 < (double)m.xmin - dx || tx0 > (double)m.xmax + dx ||
                              ty1 < (double)m.ymin - dy || ty0 > (double)m.ymax + dy;
            if (!miss) {
              // relevant edges: slabs covering the tile's y-range, taken once each
              T const qy0 = (T)ty0, qy1 = (T)ty1;
              u32 const q0 = slab_of<T>(qy0, m), q1 = slab_of<T>(qy1, m);
              u32 const kbeg = __ldg(ix.slab_start + m.slab_base + q0);
              u32 const kmid = __ldg(ix.slab_start + m.slab_base + q0 + 1);
              u32 const kend = __ldg(ix.slab_start + m.slab_base + q1 + 1);
              auto fetch = [&](u32 k, T& ax, T& ay, T& bx, T& by) -> bool {
                if (k >= kend) return false;
                u32 const ent = __ldg(ix.entries + k);
                if (k >= kmid && !(ent & kFirstFlag)) return false;
                edge_rec<T> const e = ix.edges[ent & ~kFirstFlag];
                ax = e.ax; ay = e.ay; bx = e.bx; by = e.by;
                double const dl = eps * fmax(fabs((double)ay), fabs((double)by));
                // an edge wholly to the LEFT of the tile (beyond the tolerance) can neither be
                // touched nor toggle a crossing: for a point right of a straddling edge the
                // reference's test never toggles (sign-certain), see DESIGN.md section 5
                double const dxl = eps * fmax(fabs((double)ax), fabs((double)bx));
                return ty1 >= fmin((double)ay, (double)by) - dl &&
                       ty0 <= fmax((double)ay, (double)by) + dl &&
                       fmax((double)ax, (double)bx) + dxl >= tx0;
              };
              if (kend - kbeg > coop_edges) {
                // ---- point-by-point form (many candidate edges: a large, sparse quadrant)
                u32 cross = 0, near = oob;
#pragma unroll
                for (int i = 0; i < kCellPPL; ++i) {
                  if (!((valid >> i) & 1u)) continue;
                  double const pcx = cx[i], pcy = cy[i];
                  // the point's cell clear of the polygon's (tolerance-widened) box: no edge can
                  // touch it and the crossing parity out there is even
                  if (pcx + hw < (double)m.xmin - dx || pcx - hw > (double)m.xmax + dx ||
                      pcy + hh < (double)m.ymin - dy || pcy - hh > (double)m.ymax + dy)
                    continue;
                  u32 const s0 = slab_of<T>((T)(pcy - hh), m), s1 = slab_of<T>((T)(pcy + hh), m);
                  u32 const pb = __ldg(ix.slab_start + m.slab_base + s0);
                  u32 const pm = __ldg(ix.slab_start + m.slab_base + s0 + 1);
                  u32 const pe = __ldg(ix.slab_start + m.slab_base + s1 + 1);
                  for (u32 k = pb; k < pe; ++k) {
                    u32 const ent = __ldg(ix.entries + k);
                    if (k >= pm && !(ent & kFirstFlag)) continue;
                    edge_rec<T> const e = ix.edges[ent & ~kFirstFlag];
                    double const eax = (double)e.ax, eay = (double)e.ay;
                    double const ebx = (double)e.bx, eby = (double)e.by;
                    double const d =
                      eps * fmax(fmax(fabs(eax), fabs(ebx)), fmax(fabs(eay), fabs(eby)));
                    if (fmax(eax, ebx) + d < pcx - hw) continue;  // wholly left: cannot matter
                    double const run = ebx - eax, rise = eby - eay;
                    double const ddx = pcx - eax, ddy = pcy - eay;
                    double const f   = ddx * rise - run * ddy;
                    double const thr = (fabs(rise) * hw + fabs(run) * hh) * 1.000001 +
                                       allow * (fabs(ddx) * fabs(rise) + fabs(run) * fabs(ddy));
                    bool const touch = fabs(f) <= thr && pcx >= fmin(eax, ebx) - d - hw &&
                                       pcx <= fmax(eax, ebx) + d + hw &&
                                       pcy >= fmin(eay, eby) - d - hh &&
                                       pcy <= fmax(eay, eby) + d + hh;
                    near |= (u32)touch << i;
                    bool const y1 = eay > pcy, y0 = eby > pcy;
                    cross ^= (u32)((y1 != y0) && ((f < 0.0) != y1)) << i;
                  }
                  for (u32 k = 0; k < m.n_vertical; ++k) {  // x-only rule of vertical edges
                    double const ax = (double)ix.edges[__ldg(ix.vert_edges + m.vert_begin + k)].ax;
                    near |= (u32)(ax >= pcx - hw && ax <= pcx + hw) << i;
                  }
                }
                near &= valid;
                inside = cross & ~near;
                if (near) {  // touched points: their real coordinates, the reference's arithmetic
                  load_points(near);
                  near &= valid;
                  inside &= ~near;
#pragma unroll
                  for (int i = 0; i < kCellPPL; ++i)
                    if ((near >> i) & 1u) {
                      bool const in = ((unsafe_pt >> i) & 1u)
                                        ? pip_reference<T>(xr[i], yr[i], ring_offsets, m.ring_begin,
                                                           m.ring_end, vx, vy)
                                        : pip_indexed<T>(xr[i], yr[i], m, ix);
                      inside |= (u32)in << i;
                    }
                }
              } else {
              // ---- pass 1: crossing parity of the cell centres + "an edge touches my cell"
              u32 cross = 0, near = oob;
              for (u32 k0 = kbeg; k0 < kend; k0 += 32) {
                T ax = 0, ay = 0, bx = 0, by = 0;
                bool const rel = fetch(k0 + lane, ax, ay, bx, by);
                u32 em = __ballot_sync(0xffffffffu, rel);
                while (em) {
                  int const src = __ffs(em) - 1;
                  em &= em - 1;
                  double const eax = (double)__shfl_sync(0xffffffffu, ax, src);
                  double const eay = (double)__shfl_sync(0xffffffffu, ay, src);
                  double const ebx = (double)__shfl_sync(0xffffffffu, bx, src);
                  double const eby = (double)__shfl_sync(0xffffffffu, by, src);
                  double const run = ebx - eax, rise = eby - eay;
                  double const d   = eps * fmax(fmax(fabs(eax), fabs(ebx)), fmax(fabs(eay), fabs(eby)));
                  if (fmin(eax, ebx) - d > tx1) {
                    // the edge lies wholly to the RIGHT of every cell square of the tile (beyond
                    // its tolerance): it touches none of them, and a point left of both
                    // endpoints crosses it exactly when its y is straddled -- for rise > 0 the
                    // straddle means y1 = 0 and the point is on the f < 0 side, for rise < 0
                    // y1 = 1 and f > 0: (f < 0) != y1 holds either way
#pragma unroll
                    for (int i = 0; i < kCellPPL; ++i) {
                      bool const y1 = eay > cy[i], y0 = eby > cy[i];
                      cross ^= (u32)(y1 != y0) << i;
                    }
                    continue;
                  }
                  // per-edge constants (warp-uniform): the edge's extent widened by its tolerance
                  // AND by the cell half-size, and the |f| threshold.  The rounding allowance of f
                  // uses the largest |ddx|, |ddy| any cell centre of this tile can have.
                  double const lx = fmin(eax, ebx) - d - hw, hx = fmax(eax, ebx) + d + hw;
                  double const ly = fmin(eay, eby) - d - hh, hy = fmax(eay, eby) + d + hh;
                  double const mdx = fmax(fabs(tx0 - eax), fabs(tx1 - eax));
                  double const mdy = fmax(fabs(ty0 - eay), fabs(ty1 - eay));
                  double const thr = (fabs(rise) * hw + fabs(run) * hh) * 1.000001 +
                                     allow * (mdx * fabs(rise) + fabs(run) * mdy);
#pragma unroll
                  for (int i = 0; i < kCellPPL; ++i) {
                    double const ddx = cx[i] - eax, ddy = cy[i] - eay;
                    double const f   = ddx * rise - run * ddy;  // (v - u) of the reference
                    // the edge's line passes through the widened cell square, within the edge's
                    // own (tolerance-widened) extent?
                    bool const touch = fabs(f) <= thr && cx[i] >= lx && cx[i] <= hx &&
                                       cy[i] >= ly && cy[i] <= hy;
                    near |= (u32)touch << i;
                    bool const y1 = eay > cy[i], y0 = eby > cy[i];
                    cross ^= (u32)((y1 != y0) && ((f < 0.0) != y1)) << i;
                  }
                }
              }
              for (u32 k = 0; k < m.n_vertical; ++k) {  // x-only rule of vertical edges
                double const ax = (double)ix.edges[__ldg(ix.vert_edges + m.vert_begin + k)].ax;
#pragma unroll
                for (int i = 0; i < kCellPPL; ++i)
                  near |= (u32)(ax >= cx[i] - hw && ax <= cx[i] + hw) << i;
              }
              near &= valid;
              inside = cross & ~near;
              // ---- pass 2: the touched points, with their real coordinates and the reference's
              // own arithmetic over the same edges
              if (__any_sync(0xffffffffu, near != 0)) {
                load_points(near);
                near &= valid;
                u32 const exact = near & ~unsafe_pt;
                u32 within = 0, onedge = 0;
                auto eval_edges = [&](u32 em, T ax, T ay, T bx, T by) {
                  while (em) {
                    int const src = __ffs(em) - 1;
                    em &= em - 1;
                    T const eax = __shfl_sync(0xffffffffu, ax, src);
                    T const eay = __shfl_sync(0xffffffffu, ay, src);
                    T const ebx = __shfl_sync(0xffffffffu, bx, src);
                    T const eby = __shfl_sync(0xffffffffu, by, src);
                    if (!exact) continue;
                    T const run  = fpp<T>::sub(ebx, eax);
                    T const rise = fpp<T>::sub(eby, eay);
                    T const lo = fmin(eax, ebx), hi = fmax(eax, ebx);
#pragma unroll
                    for (int i = 0; i < kCellPPL; ++i)
                      if ((exact >> i) & 1u) {
                        T const rtp  = fpp<T>::sub(yr[i], eay);
                        T const rntp = fpp<T>::sub(xr[i], eax);
                        T const u    = fpp<T>::mul(run, rtp);
                        T const v    = fpp<T>::mul(rntp, rise);
                        if (lo <= xr[i] && xr[i] <= hi) {
                          if (float_equal(u, v)) onedge |= 1u << i;
                        }
                        bool const y1 = eay > yr[i], y0 = eby > yr[i];
                        within ^= (u32)((y1 != y0) && ((v < u) != y1)) << i;
                      }
                  }
                };
                // the exact path needs every edge whose y-range meets the POINT (a subset of the
                // tile's list) plus vertical edges at the point's x: both are covered by the
                // tile-level lists below, exactly as in pip_eval_kernel
                T rx0 = fpp<T>::inf(), rx1 = -fpp<T>::inf();
#pragma unroll
                for (int i = 0; i < kCellPPL; ++i)
                  if ((exact >> i) & 1u) { rx0 = fmin(rx0, xr[i]); rx1 = fmax(rx1, xr[i]); }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                  rx0 = fmin(rx0, __shfl_xor_sync(0xffffffffu, rx0, o));
                  rx1 = fmax(rx1, __shfl_xor_sync(0xffffffffu, rx1, o));
                }
                for (u32 k0 = kbeg; k0 < kend; k0 += 32) {
                  T ax = 0, ay = 0, bx = 0, by = 0;
                  bool const rel = fetch(k0 + lane, ax, ay, bx, by);
                  eval_edges(__ballot_sync(0xffffffffu, rel), ax, ay, bx, by);
                }
                for (u32 k0 = 0; k0 < m.n_vertical; k0 += 32) {
                  u32 const k = k0 + lane;
                  T ax = 0, ay = 0, bx = 0, by = 0;
                  bool rel = false;
                  if (k < m.n_vertical) {
                    edge_rec<T> const e = ix.edges[__ldg(ix.vert_edges + m.vert_begin + k)];
                    ax = e.ax; ay = e.ay; bx = e.bx; by = e.by;
                    double const dl = eps * fmax(fabs((double)ay), fabs((double)by));
                    bool const yrel = ty1 >= fmin((double)ay, (double)by) - dl &&
                                      ty0 <= fmax((double)ay, (double)by) + dl;  // taken above
                    rel = !yrel && rx0 <= ax && ax <= rx1;
                  }
                  eval_edges(__ballot_sync(0xffffffffu, rel), ax, ay, bx, by);
                }
                inside = (inside & ~near) | (within & ~onedge & exact);
                if (unsafe_pt & near) {
#pragma unroll
                  for (int i = 0; i < kCellPPL; ++i)
                    if (((unsafe_pt & near) >> i) & 1u)
                      inside |= (u32)pip_reference<T>(xr[i], yr[i], ring_offsets, m.ring_begin,
                                                      m.ring_end, vx, vy) << i;
                }
              }
              }  // warp-cooperative form
            }
          }
        }
        inside &= valid;
        u32 mine = 0, cnt = 0;
#pragma unroll
        for (int i = 0; i < kCellPPL; ++i) {
          u32 const w = __ballot_sync(0xffffffffu, (inside >> i) & 1u);
          if (lane == (u32)i) mine = w;
          cnt += __popc(w);
        }
        if (lane < tile_words) mask_words[wbase[j] + base / 32 + lane] = mine;
        if (lane == 0 && cnt) atomicAdd(&hits[j], cnt);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// expansion of the compact result into (polygon_index, point_index) rows.  A warp draws groups of
// consecutive pairs (see the kernel); whole-quadrant hits are a 128-bit identity fill, ballot
// words are decoded through a per-warp shared-memory buffer (lane l lists the set bits of word l
// at its prefix position) and leave with one lane per OUTPUT row, i.e. coalesced.
// ---------------------------------------------------------------------------------------------
// Measured on B200 (configs[1], 1.15 GB of rows): streaming (__stcs) stores 0.36 ms, plain stores
// 0.33 ms (round 1); this form 0.23 ms = 4.8 TB/s, the same at 4-16 CTAs per SM and group sizes
// 8-32 (profiles/r2_notebook.md).
template <typename P, typename V>
__device__ __forceinline__ void emit_store(P* p, V v)
{
  *p = v;
}
#ifndef BSJ_EMIT_STORE
#define BSJ_EMIT_STORE emit_store
#endif
#ifndef BSJ_EMIT_GRID_MULT
#define BSJ_EMIT_GRID_MULT 8
#endif
constexpr int kEmitWarps = 8;
constexpr u32 kEmitSlices = 64;
__global__ void __launch_bounds__(kEmitWarps * 32)
pip_emit_kernel(const u32* __restrict__ pair_poly, const u32* __restrict__ pair_off,
                const u32* __restrict__ pair_len, u32 n_pairs, const u64* __restrict__ wbase,
                const u64* __restrict__ obase, const u32* __restrict__ hits,
                const u32* __restrict__ mask_words, const u8* __restrict__ cls, u32 position_base,
                u32* __restrict__ out_poly, u32* __restrict__ out_point, u32 group,
                u32* __restrict__ ticket, u32 n_slices)
{
  // per-warp staging of the hit positions of one 32-word group (boundary pairs)
  // (one slot of padding per 32: lane l lists word l's bits from slot 32 l on when the words are
  // full -- the common case -- which would put all lanes on two banks)
  __shared__ unsigned short s_pos[kEmitWarps][1024 + 32];
  u32 const lane = lane_id();
  unsigned short* const pos = s_pos[threadIdx.x >> 5];
  // A warp takes `group` (<= 32) consecutive pairs at a time: lane l fetches the records of pair
  // g*group + l -- seven coalesced loads in ONE latency round for the whole group instead of two
  // dependent rounds per pair (the records may sit in a peer GPU's memory: the multi-GPU merge
  // expands every rank's compact result in place) -- and the pairs with rows are then handed to
  // the whole warp one by one through shuffles.  Groups are handed out through a ticket (pairs
  // differ in cost by orders of magnitude: a whole-quadrant fill vs. ballot words to decode), the
  // next ticket being drawn while the current group is written.
  // (n_slices <= kEmitSlices counters, each shared by the warps whose index is congruent to
  // it: ten thousand warps drawing from ONE address serialise in L2 -- 20 % of this kernel's
  // stall samples with a single counter)
  u64 const n_groups  = ((u64)n_pairs + group - 1) / group;
  u32 const slice = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) % n_slices;
  ticket += slice;
  u32 g32 = 0;
  if (lane == 0) g32 = atomicAdd(ticket, 1u);
  g32 = __shfl_sync(0xffffffffu, g32, 0);
  // slice s owns groups s, s + n_slices, s + 2 n_slices, ...
  while ((u64)g32 * n_slices + slice < n_groups) {
    u64 const g = (u64)g32 * n_slices + slice;
    u32 next = 0;
    if (lane == 0) next = atomicAdd(ticket, 1u);
    u64 const jl = g * group + lane;
    u32 nh_l = 0, poly_l = 0, len_l = 0, off_l = 0, cls_l = 0;
    u64 wb_l = 0, o_l = 0;
    if (lane < group && jl < n_pairs) {
      nh_l   = hits[jl];
      poly_l = pair_poly[jl];
      len_l  = pair_len[jl];
      off_l  = pair_off[jl];
      cls_l  = cls[jl];
      wb_l   = wbase[jl];
      o_l    = obase[jl];
    }
    u32 live = __ballot_sync(0xffffffffu, nh_l != 0);
    while (live) {
      int const src = __ffs(live) - 1;
      live &= live - 1;
      u32 const nh   = __shfl_sync(0xffffffffu, nh_l, src);
      u32 const poly = __shfl_sync(0xffffffffu, poly_l, src);
      u32 const len  = __shfl_sync(0xffffffffu, len_l, src);
      u32 const off  = __shfl_sync(0xffffffffu, off_l, src) + position_base;
      u32 const c8   = __shfl_sync(0xffffffffu, cls_l, src);
      u64 const wb   = __shfl_sync(0xffffffffu, wb_l, src);
      u64 o          = __shfl_sync(0xffffffffu, o_l, src);
      u32 const words = len / 32 + ((len & 31) != 0);
      if (c8 == kClsInside) {  // whole quadrant inside: rows are (poly, off .. off+nh-1)
        u32* const op = out_poly + o;
        u32* const oq = out_point + o;
        uintptr_t const ph = reinterpret_cast<uintptr_t>(op) & 15;
        if (ph == (reinterpret_cast<uintptr_t>(oq) & 15)) {
          // 128-bit stores between a scalar head and tail
          u32 const head = min((u32)(((16 - ph) & 15) >> 2), nh);
          if (lane < head) {
            BSJ_EMIT_STORE(op + lane, poly);
            BSJ_EMIT_STORE(oq + lane, off + lane);
          }
          u32 const nvec = (nh - head) >> 2;
          uint4* const vp = reinterpret_cast<uint4*>(op + head);
          uint4* const vq = reinterpret_cast<uint4*>(oq + head);
          for (u32 v = lane; v < nvec; v += 32) {
            u32 const p0 = off + head + v * 4;
            BSJ_EMIT_STORE(vp + v, make_uint4(poly, poly, poly, poly));
            BSJ_EMIT_STORE(vq + v, make_uint4(p0, p0 + 1, p0 + 2, p0 + 3));
          }
          u32 const r = head + nvec * 4 + lane;
          if (r < nh) {
            BSJ_EMIT_STORE(op + r, poly);
            BSJ_EMIT_STORE(oq + r, off + r);
          }
        } else {
          for (u32 r = lane; r < nh; r += 32) {
            BSJ_EMIT_STORE(op + r, poly);
            BSJ_EMIT_STORE(oq + r, off + r);
          }
        }
        continue;
      }
      for (u32 w0 = 0; w0 < words; w0 += 32) {
        u32 w          = w0 + lane < words ? __ldcs(mask_words + wb + w0 + lane) : 0u;
        u32 const c    = __popc(w);
        u32 const incl = warp_inclusive_scan(c);
        u32 const tot  = __shfl_sync(0xffffffffu, incl, 31);
        if (tot == 1024) {  // every candidate of the group is a hit: identity mapping
          for (u32 r = lane; r < 1024; r += 32) {
            BSJ_EMIT_STORE(out_poly + o + r, poly);
            BSJ_EMIT_STORE(out_point + o + r, off + w0 * 32 + r);
          }
        } else if (tot) {
          // lane l lists the set bits of word l at its place in the staging buffer, then the
          // warp writes the rows with one lane per OUTPUT row (coalesced)
          u32 k = incl - c;
          while (w) {
            pos[k + (k >> 5)] = (unsigned short)(lane * 32 + (__ffs(w) - 1));
            ++k;
            w &= w - 1;
          }
          __syncwarp();
          for (u32 r = lane; r < tot; r += 32) {
            BSJ_EMIT_STORE(out_poly + o + r, poly);
            BSJ_EMIT_STORE(out_point + o + r, off + w0 * 32 + pos[r + (r >> 5)]);
          }
          __syncwarp();
        }
        o += tot;
      }
    }
    g32 = __shfl_sync(0xffffffffu, next, 0);
  }
}

// ---------------------------------------------------------------------------------------------
// bitmask point_in_polygon (<= 31 polygons): one thread per point
// ---------------------------------------------------------------------------------------------
// Exact predicate of ONE point against one indexed polygon: only the edges of the point's y-slab
// (those whose tolerance-widened y-range contains y) can cross or lie under the point; a vertical
// edge with the point's x rejects it at any y.  Same per-edge arithmetic as the reference.
template <typename T>
__device__ bool pip_indexed(T px, T py, const poly_meta<T>& m, const edge_index<T>& ix)
{
  for (u32 k = 0; k < m.n_vertical; ++k)
    if (ix.edges[__ldg(ix.vert_edges + m.vert_begin + k)].ax == px) return false;
  u32 const sl   = slab_of<T>(py, m);
  u32 const kbeg = __ldg(ix.slab_start + m.slab_base + sl);
  u32 const kend = __ldg(ix.slab_start + m.slab_base + sl + 1);
  bool within = false;
  for (u32 k = kbeg; k < kend; ++k) {
    edge_rec<T> const e = ix.edges[__ldg(ix.entries + k) & ~kFirstFlag];
    T const run  = fpp<T>::sub(e.bx, e.ax);
    T const rise = fpp<T>::sub(e.by, e.ay);
    T const rtp  = fpp<T>::sub(py, e.ay);
    T const rntp = fpp<T>::sub(px, e.ax);
    T const u    = fpp<T>::mul(run, rtp);
    T const v    = fpp<T>::mul(rntp, rise);
    if (fmin(e.ax, e.bx) <= px && px <= fmax(e.ax, e.bx)) {
      if (float_equal(u, v)) return false;  // on an edge: outside, whatever the crossings say
    }
    bool const y1 = e.ay > py, y0 = e.by > py;
    if (y1 != y0 && ((v < u) != y1)) within = !within;
  }
  return within;
}

// ---------------------------------------------------------------------------------------------
// bitmask point_in_polygon, large point sets: a uniform grid of cell classes over the polygons'
// common bounding box.  The argument is the one of classify_quadrant: a cell rectangle (widened
// by the rounding margin of the point -> cell assignment) that no tolerance-widened edge box
// touches and no vertical edge's x falls into is uniformly inside or outside a polygon, and the
// answer is the predicate of its centre.  One table entry per cell = {bits of the polygons the
// cell is inside of, bits of the polygons that need the exact test}; a point then costs one
// 8-byte lookup plus the exact slab test for the (few) undecided polygons of its cell.
// ---------------------------------------------------------------------------------------------
struct cell_grid {
  double min_x, min_y, scale, inv_scale, margin_x, margin_y;
  int log2_cells;  // cells per side = 1 << log2_cells (square cells)
  // a box that contains every polygon's rejection box (x widened by more than any polygon's own
  // tolerance): a comfy point outside it is rejected by every polygon.  Valid only when all
  // polygons are safe (an unsafe polygon takes the reference loop for every point).
  int union_valid;
  double ux0, ux1, uy0, uy1;
};

__device__ __forceinline__ u32 dilate16p(u32 v)
{
  v &= 0x0000FFFFu;
  v = (v | (v << 8)) & 0x00FF00FFu;
  v = (v | (v << 4)) & 0x0F0F0F0Fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

// one warp per cell; lane p pre-tests polygon p's box against the widened cell, the warp then
// classifies the cell against the polygons that passed, one at a time
template <typename T>
__global__ void __launch_bounds__(256)
bitmask_grid_kernel(cell_grid cg, const poly_meta<T>* __restrict__ meta, u32 n_poly,
                    edge_index<T> ix, uint2* __restrict__ cells)
{
  __shared__ poly_meta<T> s_meta[31];
  for (u32 i = threadIdx.x; i < n_poly; i += blockDim.x) s_meta[i] = meta[i];
  __syncthreads();
  u32 const lane = lane_id();
  u32 const G    = 1u << cg.log2_cells;
  u64 const c    = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= (u64)G * G) return;
  u32 const cx = (u32)c & (G - 1), cy = (u32)(c >> cg.log2_cells);
  grid_info g{};
  g.valid = 1; g.max_depth = cg.log2_cells; g.has_oob = 0;
  g.min_x = cg.min_x; g.min_y = cg.min_y; g.scale = cg.scale;
  g.margin_x = cg.margin_x; g.margin_y = cg.margin_y; g.sorted_keys = nullptr;
  u32 const key   = dilate16p(cx) | (dilate16p(cy) << 1);
  u32 const level = (u32)cg.log2_cells - 1;
  bool cand = false;
  if (lane < n_poly) {
    poly_meta<T> const& m = s_meta[lane];
    double const eps = (double)fpp<T>::eps();
    double const x0 = cg.min_x + (double)cx * cg.scale - cg.margin_x;
    double const x1 = cg.min_x + ((double)cx + 1.0) * cg.scale + cg.margin_x;
    double const y0 = cg.min_y + (double)cy * cg.scale - cg.margin_y;
    double const y1 = cg.min_y + ((double)cy + 1.0) * cg.scale + cg.margin_y;
    double const dx = eps * fmax(fabs((double)m.xmin), fabs((double)m.xmax));
    double const dy = eps * fmax(fabs((double)m.ymin), fabs((double)m.ymax));
    // same rejection as classify_quadrant's own (which stays the authority for the survivors)
    cand = !m.safe || m.n_slabs == 0 ||
           !(x1 < (double)m.xmin - dx || x0 > (double)m.xmax + dx || y1 < (double)m.ymin - dy ||
             y0 > (double)m.ymax + dy);
  }
  u32 todo = __ballot_sync(0xffffffffu, cand);
  u32 inside = 0, boundary = 0;
  while (todo) {
    int const p = __ffs(todo) - 1;
    todo &= todo - 1;
    int const cls = classify_quadrant<T>(g, key, level, s_meta[p], ix);
    inside |= (u32)(cls == kClsInside) << p;
    boundary |= (u32)(cls == kClsBoundary) << p;
  }
  if (lane == 0) cells[c] = make_uint2(inside, boundary);
}

// exact predicate of point (x, y) against polygon p, after the exact box rejections
template <typename T>
__device__ __forceinline__ bool bitmask_exact(T x, T y, bool p_ok, const poly_meta<T>& m,
                                              const u32* __restrict__ ring_offsets,
                                              const T* __restrict__ vx, const T* __restrict__ vy,
                                              const edge_index<T>& ix)
{
  if (p_ok && m.safe && m.n_slabs) return pip_indexed<T>(x, y, m, ix);
  return pip_reference<T>(x, y, ring_offsets, m.ring_begin, m.ring_end, vx, vy);
}
// no edge can straddle y outside [ymin, ymax); x beyond the widened extent decides every crossing
// comparison with certainty (parity even => outside).  Only for comfy points and safe polygons.
template <typename T>
__device__ __forceinline__ bool bitmask_rejects(T x, T y, const poly_meta<T>& m)
{
  T const mx = fpp<T>::eps() * fmax(fabs(m.xmin), fabs(m.xmax));
  return y < m.ymin || y >= m.ymax || x < m.xmin - mx || x > m.xmax + mx;
}

// A CTA takes kBmChunk points at a time, in three phases, so that the expensive exact tests run
// on converged warps (one thread per UNDECIDED (point, polygon) item) instead of a few lanes of
// every warp of the streaming loop (measured: 4 of 32 lanes active):
//   A  stream the points: cell lookup -> decided bits into the chunk's mask array (shared
//      memory), undecided (point, polygon) items into a shared-memory queue;
//   B  threads stride over the queue: exact test of one item each, result bit OR-ed into the
//      mask array;
//   C  the masks leave as coalesced 128-bit stores.
// A full queue is not an error: the item is evaluated on the spot.
constexpr int kBmBlock = 256;
constexpr int kBmPPT   = 16;                   // points per thread and chunk
constexpr int kBmChunk = kBmBlock * kBmPPT;   // 4096 points
constexpr int kBmBatch = 8;                   // points whose loads are in flight together
constexpr int kBmQueue = 6144;                // queue slots (item = polygon << 12 | point)

template <typename T>
__global__ void __launch_bounds__(kBmBlock, 3)
pip_bitmask_kernel(const T* __restrict__ px, const T* __restrict__ py, u64 n_points,
                   const poly_meta<T>* __restrict__ meta, u32 n_poly,
                   const u32* __restrict__ ring_offsets, const T* __restrict__ vx,
                   const T* __restrict__ vy, i32* __restrict__ out, int force_reference,
                   edge_index<T> ix, cell_grid cg, const uint2* __restrict__ cells)
{
  __shared__ poly_meta<T> s_meta[31];
  __shared__ __align__(16) u32 s_mask[kBmChunk];
  __shared__ u32 s_queue[kBmQueue];
  __shared__ u32 s_count;
  int const tid = threadIdx.x;
  for (u32 i = tid; i < n_poly; i += kBmBlock) s_meta[i] = meta[i];
  u32 const all_polys = n_poly >= 32 ? 0xFFFFFFFFu : ((1u << n_poly) - 1u);
  double const G      = (double)(1u << cg.log2_cells);
  u64 const n_chunks  = (n_points + kBmChunk - 1) / kBmChunk;
  for (u64 chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    u64 const base = chunk * kBmChunk;
    u32 const cnt  = (u32)min((u64)kBmChunk, n_points - base);
    if (tid == 0) s_count = 0;
    __syncthreads();  // also: s_meta loaded, previous chunk's masks written out
    // ---- phase A.  The coordinates of kBmBatch points are requested before the first one is
    // used: the kernel was bound by the latency of these loads (51 % of its stall samples sat on
    // the first use of x/y at 24 warps per SM)
    for (int k0 = 0; k0 < kBmPPT; k0 += kBmBatch) {
      T bx[kBmBatch], by[kBmBatch];
#pragma unroll
      for (int u = 0; u < kBmBatch; ++u) {
        u32 const j = (k0 + u) * kBmBlock + tid;
        bx[u] = j < cnt ? __ldcs(px + base + j) : (T)0;
        by[u] = j < cnt ? __ldcs(py + base + j) : (T)0;
      }
#pragma unroll
      for (int u = 0; u < kBmBatch; ++u) {
      u32 const j = (k0 + u) * kBmBlock + tid;
      if (j >= cnt) break;
      T const x = bx[u], y = by[u];
      bool const p_ok = comfy(x) && comfy(y) && !force_reference;
      u32 mask = 0, todo = all_polys;
      if (p_ok && cg.union_valid &&
          ((double)y < cg.uy0 || (double)y >= cg.uy1 || (double)x < cg.ux0 || (double)x > cg.ux1)) {
        todo = 0;  // outside every polygon's box
      } else if (p_ok && cells) {
        double const fx = ((double)x - cg.min_x) * cg.inv_scale;
        double const fy = ((double)y - cg.min_y) * cg.inv_scale;
        if (fx >= 0.0 && fy >= 0.0 && fx < G && fy < G) {
          uint2 const c = __ldg(cells + (((u64)(u32)fy << cg.log2_cells) + (u32)fx));
          mask = c.x;
          todo = c.y;
        }
      }
      if (p_ok && todo) {
        // exact box rejections first: they empty `todo` for most points outside the grid
        u32 keep = 0;
        for (u32 t = todo; t;) {
          int const p = __ffs(t) - 1;
          t &= t - 1;
          poly_meta<T> const& m = s_meta[p];
          if (!(m.safe && bitmask_rejects<T>(x, y, m))) keep |= 1u << p;
        }
        todo = keep;
      }
      if (todo) {
        u32 const n   = __popc(todo);
        u32 const pos = atomicAdd(&s_count, n);
        if (pos + n <= (u32)kBmQueue) {
          u32 q = pos;
          for (u32 t = todo; t; t &= t - 1) s_queue[q++] = ((u32)(__ffs(t) - 1) << 12) | j;
        } else {  // queue full: evaluate here; the slots claimed inside the queue become no-ops
          for (u32 q = pos; q < (u32)kBmQueue; ++q) s_queue[q] = 0xFFFFFFFFu;
          for (u32 t = todo; t; t &= t - 1) {
            int const p = __ffs(t) - 1;
            mask |= (u32)bitmask_exact<T>(x, y, p_ok, s_meta[p], ring_offsets, vx, vy, ix) << p;
          }
        }
      }
      s_mask[j] = mask;
      }
    }
    __syncthreads();
    // ---- phase B
    u32 const n_items = min(s_count, (u32)kBmQueue);
    for (u32 q = tid; q < n_items; q += kBmBlock) {
      u32 const item = s_queue[q];
      if (item == 0xFFFFFFFFu) continue;
      u32 const j = item & 0xFFFu, p = item >> 12;
      T const x = __ldg(px + base + j), y = __ldg(py + base + j);
      bool const p_ok = comfy(x) && comfy(y) && !force_reference;
      if (bitmask_exact<T>(x, y, p_ok, s_meta[p], ring_offsets, vx, vy, ix))
        atomicOr(&s_mask[j], 1u << p);
    }
    __syncthreads();
    // ---- phase C
    if (cnt == (u32)kBmChunk && (reinterpret_cast<uintptr_t>(out + base) & 15) == 0) {
      for (int v = tid; v < kBmChunk / 4; v += kBmBlock)
        __stcs(reinterpret_cast<uint4*>(out + base) + v, reinterpret_cast<const uint4*>(s_mask)[v]);
    } else {
      for (u32 j = tid; j < cnt; j += kBmBlock) __stcs(out + base + j, (i32)s_mask[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pairwise point_in_polygon (detail/point_in_polygon.cuh:104-145): point i against polygon i.
// One warp per pair, lanes stride the ring's edges (coalesced vertex loads), crossing parity by
// ballot/popc.  The reference walks the ring sequentially and keeps `b` unchanged over a
// degenerate segment (is_point_in_polygon.cuh:65-66), so the lane-parallel form (b = previous
// vertex) is exact only if no segment of the polygon is degenerate -- then no `continue` ever
// fires and b is always the previous vertex.  Otherwise lane 0 replays the reference loop.
// On-edge anywhere => false, independent of order (the reference breaks out with within=false).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pip_pairwise_kernel(const T* __restrict__ px, const T* __restrict__ py, u64 n_pairs,
                    const u32* __restrict__ poly_offsets, const u32* __restrict__ ring_offsets,
                    const T* __restrict__ vx, const T* __restrict__ vy, u8* __restrict__ out)
{
  u32 const lane    = lane_id();
  u64 const warp    = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  u64 const n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 i = warp; i < n_pairs; i += n_warps) {
    T const x = __ldg(px + i), y = __ldg(py + i);
    u32 const r0 = __ldg(poly_offsets + i), r1 = __ldg(poly_offsets + i + 1);
    u32 flips = 0;
    bool degenerate = false, on_edge = false;
    for (u32 r = r0; r < r1; ++r) {
      u32 const v0 = __ldg(ring_offsets + r), v1 = __ldg(ring_offsets + r + 1);
      for (u32 e = v0 + lane; e < v1; e += 32) {
        u32 const pe = e == v0 ? v1 - 1 : e - 1;
        T const ax = __ldg(vx + e), ay = __ldg(vy + e);
        T const bx = __ldg(vx + pe), by = __ldg(vy + pe);
        T const run  = fpp<T>::sub(bx, ax);
        T const rise = fpp<T>::sub(by, ay);
        if (float_equal(run, (T)0) && float_equal(rise, (T)0)) degenerate = true;
        T const rtp  = fpp<T>::sub(y, ay);
        T const rntp = fpp<T>::sub(x, ax);
        T const u    = fpp<T>::mul(run, rtp);
        T const v    = fpp<T>::mul(rntp, rise);
        if (float_equal(u, v)) {
          T const lo = ax > bx ? bx : ax, hi = ax > bx ? ax : bx;
          if (lo <= x && x <= hi) on_edge = true;
        }
        bool const y1 = ay > y, y0 = by > y;
        if (y1 != y0 && ((v < u) != y1)) flips ^= 1u;
      }
    }
    bool const any_degenerate = __any_sync(0xffffffffu, degenerate);
    bool const any_on_edge    = __any_sync(0xffffffffu, on_edge);
    u32 const parity          = __popc(__ballot_sync(0xffffffffu, flips & 1u)) & 1u;
    if (lane == 0) {
      bool hit = !any_on_edge && parity;
      if (any_degenerate) hit = pip_reference<T>(x, y, ring_offsets, r0, r1, vx, vy);
      out[i] = hit ? 1 : 0;
    }
  }
}

// per-polygon bounding boxes (detail/bounding_boxes.cuh:36-60,136-184): min/max of (v -+ r)
template <typename T>
__global__ void __launch_bounds__(128)
poly_bbox_kernel(const u32* __restrict__ poly_offsets, u32 n_poly,
                 const u32* __restrict__ ring_offsets, const T* __restrict__ vx,
                 const T* __restrict__ vy, u32 n_verts, T r, T* __restrict__ ox0,
                 T* __restrict__ oy0, T* __restrict__ ox1, T* __restrict__ oy1)
{
  u32 const p    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  u32 const lane = lane_id();
  if (p >= n_poly) return;
  u32 const v0 = ring_offsets[poly_offsets[p]];
  u32 const v1 = min(ring_offsets[poly_offsets[p + 1]], n_verts);
  T xmin = fpp<T>::inf(), ymin = fpp<T>::inf(), xmax = -fpp<T>::inf(), ymax = -fpp<T>::inf();
  for (u32 i = v0 + lane; i < v1; i += 32) {
    T const ax = vx[i], ay = vy[i];
    xmin = fmin(xmin, fpp<T>::sub(ax, r)); ymin = fmin(ymin, fpp<T>::sub(ay, r));
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

// Builds (and owns) the per-call polygon metadata + y-slab edge index.  `finish()` needs one host
// round trip (the number of index entries); callers merge it with a synchronisation they need
// anyway by calling begin() early and finish() after their own cudaStreamSynchronize.
template <typename T>
struct polygon_index {
  dev_buf<poly_meta<T>> meta;
  dev_buf<edge_rec<T>> edges;
  dev_buf<u32> idx_totals, vert_cursor, slab_count, vert_edges, slab_start, entries;
  dev_buf<u64> slab_start64, entry_total;
  u32 n_poly{0}, total_slabs{0};
  int poly_grid{1};
  const u32* ring_offsets{nullptr};
  const T* vx{nullptr};
  const T* vy{nullptr};
  u64 h_entries{0};

  void begin(const u32* poly_offsets, u64 n_poly_offsets, const u32* ring_off, u64 n_ring_offsets,
             const void* vx_, const void* vy_, u64 n_verts, cudaStream_t s)
  {
    n_poly       = (u32)(n_poly_offsets ? n_poly_offsets - 1 : 0);
    ring_offsets = ring_off;
    vx           = (const T*)vx_;
    vy           = (const T*)vy_;
    meta.alloc(std::max<u32>(n_poly, 1), s);
    idx_totals.alloc(2, s);
    edges.alloc(std::max<u64>(n_verts, 1), s);
    vert_cursor.alloc(std::max<u32>(n_poly, 1), s);
    poly_grid = div_up((u64)std::max<u32>(n_poly, 1) * 32, 128);
    poly_meta_kernel<T><<<poly_grid, 128, 0, s>>>(
      poly_offsets, n_poly, ring_offsets, (u32)(n_ring_offsets ? n_ring_offsets - 1 : 0), vx, vy,
      (u32)n_verts, meta.get());
    BSJ_CHECK_LAUNCH();
    poly_scan_kernel<T><<<1, 1024, 0, s>>>(meta.get(), n_poly, idx_totals.get());
    BSJ_CHECK_LAUNCH();
    BSJ_CUDA_TRY(cudaMemsetAsync(vert_cursor.get(), 0, vert_cursor.size() * sizeof(u32), s));
    // sizes bounded without a host round trip: a polygon gets at most as many slabs as vertices
    // (and at least one), and at most one vertical-list entry per vertex
    total_slabs = (u32)std::min<u64>(n_verts + n_poly, 0xFFFFFFF0ull);
    slab_count.alloc(total_slabs + 1, s);
    vert_edges.alloc(std::max<u64>(n_verts, 1), s);
    slab_start64.alloc(total_slabs + 1, s);
    entry_total.alloc(1, s);
    slab_start.alloc(total_slabs + 1, s);
    BSJ_CUDA_TRY(cudaMemsetAsync(slab_count.get(), 0, slab_count.size() * sizeof(u32), s));
    slab_build_kernel<T, false><<<std::max<u32>(n_poly, 1), kSlabBlock, 0, s>>>(
      meta.get(), n_poly, ring_offsets, vx, vy, edges.get(), slab_count.get(), nullptr, nullptr,
      vert_edges.get(), vert_cursor.get());
    BSJ_CHECK_LAUNCH();
    exclusive_scan_u32_to_u64(slab_count.get(), slab_start64.get(), total_slabs + 1,
                              entry_total.get(), s);
    BSJ_CUDA_TRY(cudaMemcpyAsync(&h_entries, entry_total.get(), sizeof(u64),
                                 cudaMemcpyDeviceToHost, s));
  }
  // the stream must have been synchronised after begin()
  edge_index<T> finish(cudaStream_t s)
  {
    BSJ_EXPECTS(h_entries < 0xFFFFFFFFull, "polygon edge index too large");
    entries.alloc(std::max<u64>(h_entries, 1), s);
    narrow_u64_kernel<<<div_up(total_slabs + 1, 256), 256, 0, s>>>(slab_start64.get(),
                                                                   slab_start.get(),
                                                                   total_slabs + 1);
    BSJ_CHECK_LAUNCH();
    BSJ_CUDA_TRY(cudaMemsetAsync(slab_count.get(), 0, slab_count.size() * sizeof(u32), s));
    slab_build_kernel<T, true><<<std::max<u32>(n_poly, 1), kSlabBlock, 0, s>>>(
      meta.get(), n_poly, ring_offsets, vx, vy, edges.get(), slab_count.get(), slab_start.get(),
      entries.get(), vert_edges.get(), vert_cursor.get());
    BSJ_CHECK_LAUNCH();
    return edge_index<T>{edges.get(), slab_start.get(), entries.get(), vert_edges.get()};
  }
};

int force_reference_mode()
{
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("BSJ_PIP_REFERENCE_LOOP");
    v             = (e && e[0] == '1') ? 1 : (e && e[0] == '2') ? 2 : 0;
  }
  return v;
}

// Everything up to (and including) the evaluation: per-pair hit counts, classes, ballot words and
// output row offsets -- the COMPACT form of the result (include/cuspatial_b200.h,
// bsj_pip_compact).  Expanding it into (polygon_index, point_index) rows is a separate step so
// that a multi-GPU caller can all-gather the compact form (tens of MB) instead of the rows (GBs).
template <typename T>
void qpip_compact_t(const u32* pair_poly, const u32* pair_quad, u64 n_pairs, const u32* length,
                    const u32* offset, u64 num_nodes, const u32* point_indices, const void* px,
                    const void* py, u64 n_points, const u32* poly_offsets, u64 n_poly_offsets,
                    const u32* ring_offsets, u64 n_ring_offsets, const void* vx, const void* vy,
                    u64 n_verts, const u32* node_key, const u8* node_level, const bsj_grid* grid,
                    const bsj_coord_segments* segs, out_alloc& oa, cudaStream_t s,
                    bsj_pip_compact* c)
{
  u32 const n_poly = (u32)(n_poly_offsets - 1);
  coord_source<T, false> pts{};
  pts.x = (const T*)px;
  pts.y = (const T*)py;
  pts.n_ids = (u32)n_points;
  coord_source<T, true> spts{};
  bool const segmented = segs && segs->n_segments > 0;
  if (segmented) {
    BSJ_EXPECTS(segs->n_segments <= BSJ_MAX_RANKS, "too many coordinate segments");
    spts.n_seg = segs->n_segments;
    for (int i = 0; i < segs->n_segments; ++i) {
      spts.first_id[i] = segs->first_id[i];
      spts.sx[i]       = (const T*)segs->x[i];
      spts.sy[i]       = (const T*)segs->y[i];
    }
    spts.first_id[segs->n_segments] = segs->first_id[segs->n_segments];
    spts.n_ids                      = segs->first_id[segs->n_segments];
  }
  grid_info gi{};
  {
    static int no_grid = -1;
    if (no_grid < 0) {
      const char* e = std::getenv("BSJ_PIP_NO_GRID");
      no_grid       = (e && e[0] == '1') ? 1 : 0;
    }
    // NaN coordinates pass the reference's box test and key into row/column 0, so with NaNs
    // present a cell rectangle no longer bounds its points: the hint is dropped.
    if (grid && grid->valid && !grid->has_nan && !no_grid && node_key && node_level &&
        grid->scale > 0 && grid->max_depth >= 1 && grid->max_depth <= 15) {
      gi.valid     = 1;
      gi.max_depth = grid->max_depth;
      gi.has_oob   = grid->has_out_of_bbox;
      gi.min_x = grid->min_x; gi.min_y = grid->min_y; gi.scale = grid->scale;
      // |x - min| / scale is computed with two roundings of relative size 2^-53 (2^-24): a point
      // lies at most ~3 ulp(extent) outside its cell; take 2^-40 (2^-18) of the coordinate range
      double const cm = sizeof(T) == 8 ? 9.094947017729282e-13 : 3.814697265625e-06;
      gi.margin_x = cm * (std::fabs(grid->min_x) + std::fabs(grid->max_x) +
                          std::fabs(grid->max_x - grid->min_x));
      gi.margin_y = cm * (std::fabs(grid->min_y) + std::fabs(grid->max_y) +
                          std::fabs(grid->max_y - grid->min_y));
      static int no_keys = -1;
      if (no_keys < 0) {
        const char* e = std::getenv("BSJ_PIP_NO_KEYS");
        no_keys       = (e && e[0] == '1') ? 1 : 0;
      }
      gi.sorted_keys = (grid->sorted_keys && grid->n_sorted_keys == n_points && !no_keys)
                         ? grid->sorted_keys : nullptr;
    }
  }
  polygon_index<T> pidx;
  pidx.begin(poly_offsets, n_poly_offsets, ring_offsets, n_ring_offsets, vx, vy, n_verts, s);
  auto& meta = pidx.meta;
  c->n_pairs        = n_pairs;
  c->pair_offset    = oa.get<u32>(n_pairs);
  c->pair_length    = oa.get<u32>(n_pairs);
  c->pair_hits      = oa.get<u32>(n_pairs);
  c->pair_class     = oa.get<u8>(n_pairs);
  c->pair_word_base = oa.get<u64>(n_pairs);
  c->pair_row_base  = oa.get<u64>(n_pairs);
  dev_buf<u32> words(n_pairs, s);
  dev_buf<u64> totals(4, s);
  dev_buf<u32> ticket(kEvalSlices, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(ticket.get(), 0, kEvalSlices * sizeof(u32), s));
  auto eval_slices = [](int grid_dim) {  // every slice needs at least one warp
    return (u32)std::min<u64>(kEvalSlices, (u64)std::max(grid_dim, 1) * kPipWarps);
  };
  pair_prep_kernel<<<div_up(n_pairs, 256), 256, 0, s>>>(pair_quad, (u32)n_pairs, length, offset,
                                                        (u32)num_nodes, words.get(),
                                                        c->pair_length, c->pair_offset);
  BSJ_CHECK_LAUNCH();
  exclusive_scan_u32_to_u64(words.get(), c->pair_word_base, n_pairs, totals.get() + 0, s);
  u64 h_tot[1] = {0};
  BSJ_CUDA_TRY(cudaMemcpyAsync(h_tot, totals.get(), sizeof(u64), cudaMemcpyDeviceToHost, s));
  BSJ_CUDA_TRY(cudaStreamSynchronize(s));
  u64 const total_words = h_tot[0];
  edge_index<T> ix = pidx.finish(s);
  prof_mark("pair_prep");

  c->n_words    = total_words;
  c->mask_words = oa.get<u32>(std::max<u64>(total_words, 1));
  // words of settled pairs are never read; zeroing keeps malformed pair tables deterministic
  BSJ_CUDA_TRY(cudaMemsetAsync(c->mask_words, 0, std::max<u64>(total_words, 1) * sizeof(u32), s));
  {
    bool const cells = force_reference_mode() == 0 && gi.valid && gi.sorted_keys;
    u32 const tile_points = cells ? (u32)kCellTile : (u32)kPipTile;
    // (pair, tile) units of the boundary pairs: at most every tile of every pair
    u64 const list_cap = total_words * 32 / tile_points + n_pairs + 2;
    dev_buf<u32> run_list(list_cap, s), tile_list(list_cap, s), run_count(1, s);
    BSJ_CUDA_TRY(cudaMemsetAsync(run_count.get(), 0, sizeof(u32), s));
    // pairs per warp visit: 32 when there are plenty of pairs, fewer for small tables so that
    // the work still spreads over the whole GPU
    u32 cgroup = 32;
    while (cgroup > 2 && (u64)n_pairs / cgroup < (u64)num_sms() * 64) cgroup >>= 1;
    int const cgrid = (int)std::min<u64>((u64)num_sms() * 16,
                                         (u64)div_up(div_up(n_pairs, (u64)cgroup), 8));
    pip_classify_kernel<T><<<std::max(cgrid, 1), 256, 0, s>>>(
      pair_poly, pair_quad, (u32)n_pairs, length, offset, (u32)num_nodes, (u32)n_points,
      meta.get(), n_poly, force_reference_mode(), node_key, node_level, gi, ix, c->pair_class,
      c->pair_hits, run_list.get(), tile_list.get(), run_count.get(), tile_points,
      (u32)std::min<u64>(list_cap, 0xFFFFFFFFull), cgroup);
    BSJ_CHECK_LAUNCH();
    prof_mark("pip_classify");
    u32 const list_cap32 = (u32)std::min<u64>(list_cap, 0xFFFFFFFFull);
    if (cells) {
      int const grid_dim = (int)std::min<u64>((u64)num_sms() * 8, (u64)div_up(list_cap, kPipWarps));
      static int coop_env = -1;
      if (coop_env < 0) {
        const char* e = std::getenv("BSJ_PIP_COOP_EDGES");
        coop_env = e ? std::atoi(e) : (int)kCoopEdges;
      }
      u32 const coop_edges = (u32)coop_env;
      auto launch = [&](auto const& src) {
        constexpr bool SEG = std::is_same<std::decay_t<decltype(src)>, coord_source<T, true>>::value;
        pip_eval_cells_kernel<T, SEG><<<std::max(grid_dim, 1), kPipWarps * 32, 0, s>>>(
          pair_poly, pair_quad, run_list.get(), tile_list.get(), list_cap32,
          run_count.get(), length, offset, point_indices, (u32)n_points, src, meta.get(), n_poly,
          ring_offsets, (const T*)vx, (const T*)vy, c->pair_word_base, c->mask_words,
          c->pair_hits, ticket.get(), eval_slices(grid_dim), c->pair_class, ix, gi,
          gi.sorted_keys, coop_edges);
      };
      if (segmented) launch(spts); else launch(pts);
      BSJ_CHECK_LAUNCH();
    } else if (force_reference_mode() != 2) {  // 2: timing experiment, classification only
      int const grid_dim = (int)std::min<u64>((u64)num_sms() * 8, (u64)div_up(list_cap, kPipWarps));
      auto launch = [&](auto const& src) {
        constexpr bool SEG = std::is_same<std::decay_t<decltype(src)>, coord_source<T, true>>::value;
        pip_eval_kernel<T, SEG><<<std::max(grid_dim, 1), kPipWarps * 32, 0, s>>>(
          pair_poly, pair_quad, run_list.get(), tile_list.get(), list_cap32,
          run_count.get(), length, offset, point_indices, (u32)n_points, src, meta.get(), n_poly,
          ring_offsets, (const T*)vx, (const T*)vy, c->pair_word_base, c->mask_words,
          c->pair_hits, ticket.get(), eval_slices(grid_dim), force_reference_mode(),
          c->pair_class, ix);
      };
      if (segmented) launch(spts); else launch(pts);
      BSJ_CHECK_LAUNCH();
    }
  }
  prof_mark("pip_eval");
  exclusive_scan_u32_to_u64(c->pair_hits, c->pair_row_base, n_pairs, totals.get() + 2, s);
  u64 h_hits = 0;
  BSJ_CUDA_TRY(cudaMemcpyAsync(&h_hits, totals.get() + 2, sizeof(u64), cudaMemcpyDeviceToHost, s));
  BSJ_CUDA_TRY(cudaStreamSynchronize(s));
  c->n_hits = h_hits;
}

void expand_compact(const u32* pair_poly, const bsj_pip_compact* c, u32 position_base,
                    u32* out_poly, u32* out_point, cudaStream_t s)
{
  if (c->n_hits == 0 || c->n_pairs == 0) return;
  // pairs per warp-visit: 32 when there are plenty of pairs, fewer (>= 4) when the table is
  // small, so that every warp still draws several groups
  static int mult = 0, group_min = 0, group_factor = 0;
  if (!mult) {  // tuning knobs (A/B runs)
    const char* e = std::getenv("BSJ_EMIT_MULT");
    mult = e ? std::max(1, std::atoi(e)) : BSJ_EMIT_GRID_MULT;
    e = std::getenv("BSJ_EMIT_GROUP_MIN");
    group_min = e ? std::max(1, std::atoi(e)) : 8;
    e = std::getenv("BSJ_EMIT_GROUP_FACTOR");
    group_factor = e ? std::max(1, std::atoi(e)) : 8;
  }
  u64 const n_warps = (u64)num_sms() * mult * kEmitWarps;
  u32 group = 32;
  while (group > (u32)group_min && div_up(c->n_pairs, (u64)group) < (u64)group_factor * n_warps)
    group >>= 1;
  int const grid_dim = (int)std::min<u64>((u64)num_sms() * mult,
                                          (u64)div_up(div_up(c->n_pairs, (u64)group), kEmitWarps));
  dev_buf<u32> ticket(kEmitSlices, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(ticket.get(), 0, kEmitSlices * sizeof(u32), s));
  // every slice needs at least one warp
  u32 const n_slices = (u32)std::min<u64>(kEmitSlices, (u64)std::max(grid_dim, 1) * kEmitWarps);
  pip_emit_kernel<<<std::max(grid_dim, 1), kEmitWarps * 32, 0, s>>>(
    pair_poly, c->pair_offset, c->pair_length, (u32)c->n_pairs, c->pair_word_base,
    c->pair_row_base, c->pair_hits, c->mask_words, c->pair_class, position_base, out_poly,
    out_point, group, ticket.get(), n_slices);
  BSJ_CHECK_LAUNCH();
  prof_mark("pip_emit");
}

template <typename T>
void pip_bitmask_t(const void* px, const void* py, u64 n_points, const u32* poly_offsets,
                   u64 n_poly_offsets, const u32* ring_offsets, u64 n_ring_offsets,
                   const void* vx, const void* vy, u64 n_verts, cudaStream_t s, i32* out)
{
  stage_timer tm(s);
  u32 const n_poly = (u32)(n_poly_offsets ? n_poly_offsets - 1 : 0);
  polygon_index<T> pidx;
  pidx.begin(poly_offsets, n_poly_offsets, ring_offsets, n_ring_offsets, vx, vy, n_verts, s);
  // the polygon boxes come back with the index totals (same synchronisation): they size the grid
  std::vector<poly_meta<T>> h_meta(n_poly);
  if (n_poly)
    BSJ_CUDA_TRY(cudaMemcpyAsync(h_meta.data(), pidx.meta.get(), n_poly * sizeof(poly_meta<T>),
                                 cudaMemcpyDeviceToHost, s));
  BSJ_CUDA_TRY(cudaStreamSynchronize(s));
  edge_index<T> ix = pidx.finish(s);
  prof_mark("polygon_index");

  // cell-class grid (see bitmask_grid_kernel): worth building when the points outnumber the
  // cells by far (measured at 100 M points x 31 polygons: 1.47 ms with 2^9 cells per side,
  // 1.56 ms with 2^10 -- the finer grid costs 0.14 ms more to build)
  cell_grid cg{};
  dev_buf<uint2> cells;
  if (n_poly && force_reference_mode() == 0) {
    bool all_safe = true;
    double ux0 = INFINITY, uy0 = INFINITY, ux1 = -INFINITY, uy1 = -INFINITY, mx = 0;
    for (auto const& m : h_meta) {
      all_safe = all_safe && m.safe;
      ux0 = std::min(ux0, (double)m.xmin); ux1 = std::max(ux1, (double)m.xmax);
      uy0 = std::min(uy0, (double)m.ymin); uy1 = std::max(uy1, (double)m.ymax);
      mx = std::max(mx, std::max(std::fabs((double)m.xmin), std::fabs((double)m.xmax)));
    }
    if (all_safe && std::isfinite(ux0) && std::isfinite(ux1) && std::isfinite(uy0) &&
        std::isfinite(uy1)) {
      // the kernel's per-polygon rejection widens x by eps * max|x| evaluated in T (possibly
      // contracted to an FMA): 4x that bound is outside every polygon's own threshold
      double const wx = 4.0 * (double)fpp<T>::eps() * mx + 1e-300;
      cg.union_valid = 1;
      cg.ux0 = ux0 - wx; cg.ux1 = ux1 + wx; cg.uy0 = uy0; cg.uy1 = uy1;
    }
  }
  // resolution: classifying a cell costs about as much as a few points, and every polygon whose
  // box meets the cell is classified -- keep cells x polygons below a quarter of the points
  int log2_cells = 0;
  if (n_points >= (1ull << 20)) {
    double const budget = (double)n_points / (4.0 * (double)std::max<u32>(n_poly, 1));
    log2_cells = std::max(6, std::min(10, (int)std::floor(0.5 * std::log2(std::max(budget, 1.0)))));
  }
  if (const char* e = std::getenv("BSJ_BITMASK_GRID_LOG2")) log2_cells = std::atoi(e);
  log2_cells = std::min(log2_cells, 12);
  if (log2_cells >= 1 && n_poly && force_reference_mode() == 0) {
    double x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
    for (auto const& m : h_meta) {
      if (!m.safe || m.n_slabs == 0) continue;
      x0 = std::min(x0, (double)m.xmin); x1 = std::max(x1, (double)m.xmax);
      y0 = std::min(y0, (double)m.ymin); y1 = std::max(y1, (double)m.ymax);
    }
    double const ext = std::max(x1 - x0, y1 - y0);
    if (std::isfinite(x0) && std::isfinite(y0) && std::isfinite(ext) && ext > 0) {
      double const G = (double)(1u << log2_cells);
      cg.min_x = x0; cg.min_y = y0;
      cg.scale = ext / G * (1.0 + 1e-12);
      cg.inv_scale = 1.0 / cg.scale;
      cg.log2_cells = log2_cells;
      // The cell index is computed in double from the (exactly converted) coordinate: three
      // roundings of 2^-53, i.e. a point lies at most ~1e-12 cells outside its cell; the cell's
      // centre is rounded to T.  The margin covers both with room to spare.
      double const rel = sizeof(T) == 8 ? 9.094947017729282e-13 /* 2^-40 */
                                        : 2.384185791015625e-07 /* 2^-22 */;
      cg.margin_x = rel * (std::fabs(x0) + std::fabs(x0 + ext) + ext) + 1e-9 * cg.scale;
      cg.margin_y = rel * (std::fabs(y0) + std::fabs(y0 + ext) + ext) + 1e-9 * cg.scale;
      u64 const n_cells = 1ull << (2 * log2_cells);
      cells.alloc(n_cells, s);
      bitmask_grid_kernel<T><<<(unsigned)div_up(n_cells * 32, 256), 256, 0, s>>>(
        cg, pidx.meta.get(), n_poly, ix, cells.get());
      BSJ_CHECK_LAUNCH();
      prof_mark("bitmask_grid");
    }
  }
  int const grid = (int)std::min<u64>((u64)num_sms() * 3, (u64)div_up(n_points, kBmChunk));
  pip_bitmask_kernel<T><<<std::max(grid, 1), kBmBlock, 0, s>>>(
    (const T*)px, (const T*)py, n_points, pidx.meta.get(), n_poly, ring_offsets, (const T*)vx,
    (const T*)vy, out, force_reference_mode(), ix, cg, cells.get());
  BSJ_CHECK_LAUNCH();
  tm.mark("pip_bitmask");
  tm.finish();  // results are ready in stream order; no trailing host synchronisation
}

}  // namespace

namespace {
struct temp_allocator {  // compact buffers that live only for the duration of one call
  bsj_allocator a;
  cudaStream_t s;
  static void* alloc(size_t bytes, bsj_stream_t st, void*)
  {
    void* p = nullptr;
    ensure_pool_configured();
    return cudaMallocAsync(&p, bytes, (cudaStream_t)st) == cudaSuccess ? p : nullptr;
  }
  static void dealloc(void* p, size_t, bsj_stream_t st, void*) { cudaFreeAsync(p, (cudaStream_t)st); }
  explicit temp_allocator(cudaStream_t st) : a{&alloc, &dealloc, nullptr}, s(st) {}
};
}  // namespace

void quadtree_point_in_polygon_compact_impl(
  const u32* pair_poly, const u32* pair_quad, u64 n_pairs, const u32* key, const u8* level,
  const u8* internal, const u32* length, const u32* offset, u64 num_nodes,
  const u32* point_indices, const void* px, const void* py, int dtype, u64 n_points,
  const u32* poly_offsets, u64 n_poly_offsets, const u32* ring_offsets, u64 n_ring_offsets,
  const void* vx, const void* vy, u64 n_verts, const bsj_grid* grid,
  const bsj_coord_segments* segs, const bsj_allocator* mr, cudaStream_t s, bsj_pip_compact* c)
{
  (void)internal;
  *c = bsj_pip_compact{};
  // empty inputs: cpp/src/join/quadtree_point_in_polygon.cu:171-178
  if (n_pairs == 0 || num_nodes == 0 || n_points == 0 || n_poly_offsets == 0) return;
  BSJ_EXPECTS(n_pairs < 0xFFFFFFFFull && num_nodes < 0xFFFFFFFFull && n_points <= 0xFFFFFFFFull,
              "table too large");
  BSJ_EXPECTS(n_ring_offsets >= 1 || n_poly_offsets <= 1, "ring offsets must not be empty");
  stage_timer tm(s);
  out_alloc oa(mr, s);
  if (dtype == BSJ_FLOAT32)
    qpip_compact_t<float>(pair_poly, pair_quad, n_pairs, length, offset, num_nodes, point_indices,
                          px, py, n_points, poly_offsets, n_poly_offsets, ring_offsets,
                          n_ring_offsets, vx, vy, n_verts, key, level, grid, segs, oa, s, c);
  else
    qpip_compact_t<double>(pair_poly, pair_quad, n_pairs, length, offset, num_nodes,
                           point_indices, px, py, n_points, poly_offsets, n_poly_offsets,
                           ring_offsets, n_ring_offsets, vx, vy, n_verts, key, level, grid, segs,
                           oa, s, c);
  tm.finish();
  oa.commit();
}

void expand_pip_compact_impl(const u32* pair_poly, const bsj_pip_compact* c, u32 position_base,
                             u32* out_poly, u32* out_point, cudaStream_t s)
{
  stage_timer tm(s);
  expand_compact(pair_poly, c, position_base, out_poly, out_point, s);
  tm.finish();  // results are ready in stream order; no trailing host synchronisation
}

void quadtree_point_in_polygon_impl(const u32* pair_poly, const u32* pair_quad, u64 n_pairs,
                                    const u32* key, const u8* level, const u8* internal,
                                    const u32* length, const u32* offset, u64 num_nodes,
                                    const u32* point_indices, const void* px, const void* py,
                                    int dtype, u64 n_points, const u32* poly_offsets,
                                    u64 n_poly_offsets, const u32* ring_offsets,
                                    u64 n_ring_offsets, const void* vx, const void* vy,
                                    u64 n_verts, const bsj_grid* grid, const bsj_allocator* mr,
                                    cudaStream_t s, bsj_pairs* out)
{
  (void)internal;
  *out = bsj_pairs{};
  if (n_pairs == 0 || num_nodes == 0 || n_points == 0 || n_poly_offsets == 0) return;
  BSJ_EXPECTS(n_pairs < 0xFFFFFFFFull && num_nodes < 0xFFFFFFFFull && n_points <= 0xFFFFFFFFull,
              "table too large");
  BSJ_EXPECTS(n_ring_offsets >= 1 || n_poly_offsets <= 1, "ring offsets must not be empty");
  stage_timer tm(s);
  temp_allocator tmp(s);
  out_alloc scratch(&tmp.a, s);  // the compact form is a temporary here: freed on return
  bsj_pip_compact c{};
  if (dtype == BSJ_FLOAT32)
    qpip_compact_t<float>(pair_poly, pair_quad, n_pairs, length, offset, num_nodes, point_indices,
                          px, py, n_points, poly_offsets, n_poly_offsets, ring_offsets,
                          n_ring_offsets, vx, vy, n_verts, key, level, grid, nullptr, scratch, s,
                          &c);
  else
    qpip_compact_t<double>(pair_poly, pair_quad, n_pairs, length, offset, num_nodes,
                           point_indices, px, py, n_points, poly_offsets, n_poly_offsets,
                           ring_offsets, n_ring_offsets, vx, vy, n_verts, key, level, grid,
                           nullptr, scratch, s, &c);
  out_alloc oa(mr, s);
  out->size = c.n_hits;
  if (c.n_hits) {
    out->first  = oa.get<u32>(c.n_hits);
    out->second = oa.get<u32>(c.n_hits);
    expand_compact(pair_poly, &c, 0u, out->first, out->second, s);
  }
  tm.finish();  // results are ready in stream order; no trailing host synchronisation
  oa.commit();
  // `scratch` is not committed: its destructor releases the compact buffers
}

void point_in_polygon_impl(const void* px, const void* py, int dtype, u64 n_points,
                           const i32* poly_offsets, u64 n_poly_offsets, const i32* ring_offsets,
                           u64 n_ring_offsets, const void* vx, const void* vy, u64 n_verts,
                           cudaStream_t s, i32* out_mask)
{
  // detail/point_in_polygon.cuh:93-94
  BSJ_EXPECTS(n_poly_offsets == 0 || n_poly_offsets - 1 <= 31, "Number of polygons cannot exceed 31");
  if (n_points == 0) return;
  // offsets are non-negative int32: reinterpreting as uint32 is value preserving
  if (dtype == BSJ_FLOAT32)
    pip_bitmask_t<float>(px, py, n_points, (const u32*)poly_offsets, n_poly_offsets,
                         (const u32*)ring_offsets, n_ring_offsets, vx, vy, n_verts, s, out_mask);
  else
    pip_bitmask_t<double>(px, py, n_points, (const u32*)poly_offsets, n_poly_offsets,
                          (const u32*)ring_offsets, n_ring_offsets, vx, vy, n_verts, s, out_mask);
}

void pairwise_point_in_polygon_impl(const void* px, const void* py, int dtype, u64 n_points,
                                    const i32* poly_offsets, u64 n_poly_offsets,
                                    const i32* ring_offsets, u64 n_ring_offsets, const void* vx,
                                    const void* vy, u64 n_verts, cudaStream_t s, u8* out)
{
  // cpp/src/point_in_polygon/point_in_polygon.cu:122-125
  BSJ_EXPECTS(n_points == (n_poly_offsets ? n_poly_offsets - 1 : 0),
              "Must pass in the same number of points as polygons.");
  if (n_points == 0) return;
  (void)n_ring_offsets;
  (void)n_verts;
  int const grid = (int)std::min<u64>((u64)num_sms() * 8, (u64)div_up(n_points * 32, 256));
  // offsets are non-negative int32: reinterpreting as uint32 is value preserving
  if (dtype == BSJ_FLOAT32)
    pip_pairwise_kernel<float><<<std::max(grid, 1), 256, 0, s>>>(
      (const float*)px, (const float*)py, n_points, (const u32*)poly_offsets,
      (const u32*)ring_offsets, (const float*)vx, (const float*)vy, out);
  else
    pip_pairwise_kernel<double><<<std::max(grid, 1), 256, 0, s>>>(
      (const double*)px, (const double*)py, n_points, (const u32*)poly_offsets,
      (const u32*)ring_offsets, (const double*)vx, (const double*)vy, out);
  BSJ_CHECK_LAUNCH();
}

void polygon_bounding_boxes_impl(const u32* poly_offsets, u64 n_poly_offsets,
                                 const u32* ring_offsets, u64 n_ring_offsets, const void* vx,
                                 const void* vy, int dtype, u64 n_verts, double r, cudaStream_t s,
                                 void* x0, void* y0, void* x1, void* y1)
{
  if (n_poly_offsets < 2 || n_ring_offsets < 2 || n_verts == 0) return;
  u32 const n_poly = (u32)(n_poly_offsets - 1);
  if (dtype == BSJ_FLOAT32)
    poly_bbox_kernel<float><<<div_up((u64)n_poly * 32, 128), 128, 0, s>>>(
      poly_offsets, n_poly, ring_offsets, (const float*)vx, (const float*)vy, (u32)n_verts,
      (float)r, (float*)x0, (float*)y0, (float*)x1, (float*)y1);
  else
    poly_bbox_kernel<double><<<div_up((u64)n_poly * 32, 128), 128, 0, s>>>(
      poly_offsets, n_poly, ring_offsets, (const double*)vx, (const double*)vy, (u32)n_verts, r,
      (double*)x0, (double*)y0, (double*)x1, (double*)y1);
  BSJ_CHECK_LAUNCH();
}

}  // namespace bsj

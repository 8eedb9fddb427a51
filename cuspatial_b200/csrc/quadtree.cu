// quadtree.cu -- B200-native quadtree builder (replaces cuspatial::quadtree_on_points).
//
// Reference behaviour restated: cpp/include/cuspatial/detail/point_quadtree.cuh:238-272 (clamps),
// detail/index/construction/phase_1.cuh:60-95 (Morton keys + stable sort),
// phase_1.cuh:108-381 + phase_2.cuh:56-345 + detail/point_quadtree.cuh:43-188 (tree arrays).
//
// Design (not a port): the reference materialises every non-empty cell of every level bottom-up
// (~2.5 N nodes, ~18 GB transient for 100 M points) and prunes afterwards.  Here the tree is built
// TOP-DOWN from the sorted keys: a node is a key range [start, start+cnt) of the sorted array, its
// children are found with binary searches inside that range, and only nodes that survive the
// max_size rule are ever created.  Rows come out directly in the reference's (level, key) order
// with children contiguous, so `offset`/`length` need no post-pass.  Work is O(Q log(N/Q)) instead
// of O(N * depth); the N-sized work is the Morton encode (+ fused digit histograms) and the
// onesweep sort.
#include "radix_sort.cuh"

namespace bsj {

namespace {

// z_order.cuh:62-77 -- arithmetic dilation instead of the reference's lookup tables
__device__ __forceinline__ u32 dilate16(u32 v)
{
  v &= 0xFFFFu;
  v = (v | (v << 8)) & 0x00FF00FFu;
  v = (v | (v << 4)) & 0x0F0F0F0Fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

template <typename T>
struct fp;
template <>
struct fp<float> {
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ u32 to_u32(float v) { return __float2uint_rz(v); }
  static __device__ __forceinline__ float trunc_(float v) { return truncf(v); }
  // |a*inv - a/s| < 2^-22 * 65536 = 2^-6 for quotients below 2^16
  static __host__ __device__ constexpr float guard() { return 0.03125f; }
};
template <>
struct fp<double> {
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ u32 to_u32(double v) { return __double2uint_rz(v); }
  static __device__ __forceinline__ double trunc_(double v) { return trunc(v); }
  // |a*inv - a/s| < 2^-51 * 65536 = 2^-35 for quotients below 2^16
  static __host__ __device__ constexpr double guard() { return 9.313225746154785e-10; }  // 2^-30
};

// uint16 cell index == static_cast<uint16_t>(a / scale) of the reference (phase_1.cuh:83-84), bit
// for bit.  The IEEE quotient is only needed when it could land on the other side of an integer
// than the cheap product a * (1/scale): both are within `guard` of the exact quotient, so when the
// product's fractional part is farther than `guard` from 0 and 1 (and the quotient is in range)
// the truncations agree.  Otherwise (probability ~2^-29 for fp64, ~6 % for fp32; NaN; huge
// values) the true division is evaluated, out of line so the hot loop stays short.
// static_cast<uint16_t>(T) compiles to cvt.rzi.u32 (saturating, NaN -> 0) followed by & 0xFFFF,
// restated explicitly; IEEE division (the reference builds with the default -prec-div=true).
template <typename T>
__device__ __noinline__ u32 cell_index_exact(T a, T scale, T coord, u32* point_flags)
{
  if (coord != coord && !(*(volatile u32*)point_flags & 2u)) atomicOr(point_flags, 2u);  // a NaN
  return fp<T>::to_u32(fp<T>::div(a, scale)) & 0xFFFFu;
}

template <typename T>
__device__ __forceinline__ u32 cell_index(T coord, T lo, T scale, T inv_scale, u32* point_flags)
{
  T const a = fp<T>::sub(coord, lo);
  if constexpr (sizeof(T) == 4) {
    // fp32: the guard band would be 2^-5 wide (6 % of the coordinates, i.e. some lane of nearly
    // every warp), and the IEEE fp32 division is only ~10 instructions: always divide.
    if (coord != coord && !(*(volatile u32*)point_flags & 2u)) atomicOr(point_flags, 2u);
    return fp<T>::to_u32(fp<T>::div(a, scale)) & 0xFFFFu;
  }
  T const q = fp<T>::mul(a, inv_scale);
  T const t = fp<T>::trunc_(q);
  T const f = q - t;
  // (q < 65536 and f > guard imply q >= 0: a negative q has f <= 0)
  if (q < (T)65536 && f > fp<T>::guard() && f < (T)1 - fp<T>::guard()) return (u32)(int)t;
  return cell_index_exact<T>(a, scale, coord, point_flags);
}

// phase_1.cuh:78-85
template <typename T>
__device__ __forceinline__ u32 point_key(T x, T y, T min_x, T min_y, T max_x, T max_y, T scale,
                                         T inv_scale, u32 oob_key, u32& flags, u32* point_flags)
{
  if (x < min_x || x > max_x || y < min_y || y > max_y) {
    flags |= 1u;
    return oob_key;
  }
  u32 const ix = cell_index<T>(x, min_x, scale, inv_scale, point_flags);
  u32 const iy = cell_index<T>(y, min_y, scale, inv_scale, point_flags);
  return (dilate16(iy) << 1) | dilate16(ix);
}

// ---------------------------------------------------------------------------------------------
// K1: Morton encode fused with the radix digit histograms of all sort passes.
// 128-bit coordinate loads, 128/64-bit key stores, shared-memory histograms.
// Algorithmic bytes per point: 2*sizeof(T) read + 4 written.
// The pass count is a template parameter: the tally unrolls to one shift/mask + one shared
// atomic with an immediate offset per pass.
// ---------------------------------------------------------------------------------------------
constexpr int kLeadBins = 8192;
template <typename T, int PASSES>
__global__ void __launch_bounds__(512, 2)
encode_hist_kernel(const T* __restrict__ x, const T* __restrict__ y, u64 n, T min_x, T min_y,
                   T max_x, T max_y, T scale, u32 oob_key, u32* __restrict__ keys,
                   u32* __restrict__ hist, u32* __restrict__ point_flags, int bin_shift,
                   u32 n_bins)
{
  u32 flags = 0;  // bit 0: a point outside the box (bit 1, a NaN coordinate, is set out of line)
  T const inv_scale = (T)1 / scale;
  constexpr int V = 16 / sizeof(T);  // points per 128-bit load
  // PASSES == 0: keys only; PASSES < 0: ONE histogram of the keys' leading bits (key >>
  // bin_shift, n_bins <= kLeadBins bins) -- the multi-GPU sharding plan's first level
  constexpr int kHistWords = PASSES < 0 ? kLeadBins : (PASSES ? PASSES : 1) * kRadixDigits;
  __shared__ u32 s_hist[kHistWords];
  int const hist_len = PASSES < 0 ? (int)n_bins : PASSES * kRadixDigits;
  for (int i = threadIdx.x; i < hist_len; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();

  u64 const nvec   = n / V;
  u64 const stride = (u64)gridDim.x * blockDim.x;
  bool const aligned =
    ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
    (reinterpret_cast<uintptr_t>(keys) & (4 * V - 1)) == 0;

  auto tally = [&](u32 k) {
    if constexpr (PASSES < 0) {
      atomicAdd(&s_hist[k >> bin_shift], 1u);
    } else {
#pragma unroll
      for (int p = 0; p < PASSES; ++p)
        atomicAdd(&s_hist[p * kRadixDigits + ((k >> (p * kRadixBits)) & 0xFFu)], 1u);
    }
  };
  auto key_of = [&](T px, T py) {
    return point_key<T>(px, py, min_x, min_y, max_x, max_y, scale, inv_scale, oob_key, flags,
                        point_flags);
  };

  if (aligned) {
    // two 128-bit vectors per coordinate in flight per thread: the kernel is bound by load
    // latency at 2 CTAs/SM (ncu: 59 % of the stall samples on the first use of x/y), so the
    // bytes in flight per SM are what sets its bandwidth
    auto emit = [&](u64 v, const T* xs, const T* ys) {
      u32 ks[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        ks[j] = key_of(xs[j], ys[j]);
        tally(ks[j]);
      }
      if constexpr (V == 2)
        *reinterpret_cast<uint2*>(keys + v * V) = make_uint2(ks[0], ks[1]);
      else
        *reinterpret_cast<uint4*>(keys + v * V) = make_uint4(ks[0], ks[1], ks[2], ks[3]);
    };
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (; v + 2 * stride < nvec; v += 3 * stride) {
      T xa[V], ya[V], xb[V], yb[V], xc[V], yc[V];
      *reinterpret_cast<int4*>(xa) = __ldcs(reinterpret_cast<const int4*>(x) + v);
      *reinterpret_cast<int4*>(ya) = __ldcs(reinterpret_cast<const int4*>(y) + v);
      *reinterpret_cast<int4*>(xb) = __ldcs(reinterpret_cast<const int4*>(x) + v + stride);
      *reinterpret_cast<int4*>(yb) = __ldcs(reinterpret_cast<const int4*>(y) + v + stride);
      *reinterpret_cast<int4*>(xc) = __ldcs(reinterpret_cast<const int4*>(x) + v + 2 * stride);
      *reinterpret_cast<int4*>(yc) = __ldcs(reinterpret_cast<const int4*>(y) + v + 2 * stride);
      emit(v, xa, ya);
      emit(v + stride, xb, yb);
      emit(v + 2 * stride, xc, yc);
    }
    for (; v < nvec; v += stride) {
      T xa[V], ya[V];
      *reinterpret_cast<int4*>(xa) = __ldcs(reinterpret_cast<const int4*>(x) + v);
      *reinterpret_cast<int4*>(ya) = __ldcs(reinterpret_cast<const int4*>(y) + v);
      emit(v, xa, ya);
    }
    // tail
    for (u64 i = nvec * V + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      u32 const k = key_of(x[i], y[i]);
      tally(k);
      keys[i] = k;
    }
  } else {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      u32 const k = key_of(x[i], y[i]);
      tally(k);
      keys[i] = k;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hist_len; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
  if (flags) atomicOr(point_flags, flags);
}

// ---------------------------------------------------------------------------------------------
// Tree construction state kept on the device (no host round-trips between levels).
// ---------------------------------------------------------------------------------------------
struct tree_state {
  u32 level_begin[17];
  u32 level_end[17];
  u32 overflow;
  u32 point_flags;  // written by the encode kernel
};

// first index in [lo, hi) whose key is >= target (target is 64-bit: wide keys cannot overflow it)
__device__ __forceinline__ u32 lower_bound_key(const u32* __restrict__ keys, u32 lo, u32 hi,
                                               u64 target)
{
  while (lo < hi) {
    u32 const mid = lo + ((hi - lo) >> 1);
    if ((u64)__ldg(keys + mid) < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

// warp-cooperative 32-ary lower bound (short dependent-load chain for the huge top-level ranges)
__device__ __forceinline__ u32 warp_lower_bound_key(const u32* __restrict__ keys, u32 lo, u32 hi,
                                                    u64 target)
{
  u32 const lane = lane_id();
  while (hi - lo > 32) {
    u64 const span  = (u64)(hi - lo);
    u32 const probe = lo + (u32)((span * (lane + 1)) / 33);  // 32 interior probes
    bool const less = (u64)__ldg(keys + probe) < target;
    u32 const m     = __ballot_sync(0xffffffffu, less);
    int const c     = __popc(m);  // probes [0,c) are < target (keys sorted)
    u32 const nlo   = c == 0 ? lo : __shfl_sync(0xffffffffu, probe, c - 1) + 1;
    u32 const nhi   = c == 32 ? hi : __shfl_sync(0xffffffffu, probe, c);
    lo              = nlo;
    hi              = nhi;
  }
  u32 const idx   = lo + lane;
  bool const less = idx < hi && (u64)__ldg(keys + idx) < target;
  return lo + __popc(__ballot_sync(0xffffffffu, less));
}

// Level 0 (children of the implicit root): distinct values of key >> shift0, found by one warp.
// For max_depth <= 1 these rows are the whole (leaf-only) tree, detail/point_quadtree.cuh:155-188.
__global__ void level0_kernel(const u32* __restrict__ keys, u32 n, int shift0, u32 cap,
                              u32* __restrict__ okey, u8* __restrict__ olevel,
                              u8* __restrict__ ointernal, u32* __restrict__ olength,
                              u32* __restrict__ ooffset, tree_state* st)
{
  u32 pos = 0, rows = 0;
  while (pos < n) {
    u32 const k     = __ldg(keys + pos) >> shift0;
    u64 const next  = ((u64)k + 1) << shift0;
    u32 const end   = next > 0xFFFFFFFFull ? n : warp_lower_bound_key(keys, pos, n, next);
    if (lane_id() == 0) {
      if (rows < cap) {
        okey[rows]      = k;
        olevel[rows]    = 0;
        ointernal[rows] = 0;
        olength[rows]   = end - pos;
        ooffset[rows]   = pos;
      } else {
        st->overflow = 1;
      }
    }
    ++rows;
    pos = end;
  }
  if (lane_id() == 0) {
    st->level_begin[0] = 0;
    st->level_end[0]   = min(rows, cap);
  }
}

// Expand level L -> L+1.  One thread per level-L node; nodes with more than max_size points
// become internal and emit their non-empty children (ordered by key) right after the children of
// all earlier nodes: chained scan over tiles in ticket order.  Rows of level L+1 start at
// level_end[L].  phase_2.cuh:256-281 (prune rule), :321-341 (internal flag),
// detail/point_quadtree.cuh:88-133 (child offsets / lengths).
constexpr int kExpandBlock = 256;
__global__ void __launch_bounds__(kExpandBlock)
expand_level_kernel(const u32* __restrict__ keys, int L, int max_depth, u32 max_size, u32 cap,
                    u32* __restrict__ okey, u8* __restrict__ olevel, u8* __restrict__ ointernal,
                    u32* __restrict__ olength, u32* __restrict__ ooffset, tree_state* st,
                    u64* __restrict__ lookback, u32* __restrict__ ticket, u32* __restrict__ done)
{
  __shared__ u32 s_tile, s_base, s_warp_sums[kExpandBlock / 32];
  u32 const lbeg      = st->level_begin[L];
  u32 const lend      = st->level_end[L];
  u32 const count     = lend - lbeg;
  u32 const num_tiles = (count + kExpandBlock - 1) / kExpandBlock;
  u32 const tag_agg = 2u * (L + 1), tag_pre = 2u * (L + 1) + 1u;
  int const shift   = 2 * (max_depth - 2 - L);  // level L+1 cell key = sorted key >> shift
  int const tid     = threadIdx.x;

  while (true) {
    __syncthreads();
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    u32 const tile = s_tile;
    if (tile >= num_tiles) break;

    u32 const row = lbeg + tile * kExpandBlock + tid;
    u32 k = 0, cnt = 0, start = 0, b[5] = {0, 0, 0, 0, 0};
    u32 nchild = 0;
    if (row < lend) {
      k     = okey[row];
      cnt   = olength[row];
      start = ooffset[row];
      if (cnt > max_size) {
        b[0] = start;
        b[4] = start + cnt;
        // three independent binary searches advanced in lock step: three loads in flight per
        // round instead of three dependent searches one after the other
        u32 lo1 = start, lo2 = start, lo3 = start, hi1 = b[4], hi2 = b[4], hi3 = b[4];
        u64 const t1 = (((u64)k << 2) + 1) << shift, t2 = (((u64)k << 2) + 2) << shift,
                  t3 = (((u64)k << 2) + 3) << shift;
        while (lo1 < hi1 || lo2 < hi2 || lo3 < hi3) {
          u32 const m1 = lo1 + ((hi1 - lo1) >> 1), m2 = lo2 + ((hi2 - lo2) >> 1),
                    m3 = lo3 + ((hi3 - lo3) >> 1);
          u64 const k1 = lo1 < hi1 ? (u64)__ldg(keys + m1) : 0;
          u64 const k2 = lo2 < hi2 ? (u64)__ldg(keys + m2) : 0;
          u64 const k3 = lo3 < hi3 ? (u64)__ldg(keys + m3) : 0;
          if (lo1 < hi1) { if (k1 < t1) lo1 = m1 + 1; else hi1 = m1; }
          if (lo2 < hi2) { if (k2 < t2) lo2 = m2 + 1; else hi2 = m2; }
          if (lo3 < hi3) { if (k3 < t3) lo3 = m3 + 1; else hi3 = m3; }
        }
        b[1] = lo1; b[2] = lo2; b[3] = lo3;
#pragma unroll
        for (int c = 0; c < 4; ++c) nchild += (b[c + 1] > b[c]);
      }
    }
    // block exclusive scan of nchild
    u32 const incl = warp_inclusive_scan(nchild);
    if ((tid & 31) == 31) s_warp_sums[tid >> 5] = incl;
    __syncthreads();
    u32 wbase = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kExpandBlock / 32; ++w) {
      u32 const s = s_warp_sums[w];
      if (w < (tid >> 5)) wbase += s;
      block_total += s;
    }
    if (tid == 0) s_base = lookback_exclusive(lookback, tile, block_total, tag_agg, tag_pre);
    __syncthreads();
    u32 const first_child = lend + s_base + wbase + incl - nchild;

    if (nchild) {
      ointernal[row] = 1;
      olength[row]   = nchild;
      ooffset[row]   = first_child;
      u32 r          = first_child;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (b[c + 1] > b[c]) {
          if (r < cap) {
            okey[r]      = (k << 2) + c;
            olevel[r]    = (u8)(L + 1);
            ointernal[r] = 0;
            olength[r]   = b[c + 1] - b[c];
            ooffset[r]   = b[c];
          } else {
            st->overflow = 1;
          }
          ++r;
        }
      }
    }
    // the last tile knows the level's total number of children
    if (tid == 0 && tile == num_tiles - 1) {
      st->level_begin[L + 1] = lend;
      st->level_end[L + 1]   = min(lend + s_base + block_total, cap);
    }
  }
  // empty level: propagate an empty range so deeper launches do nothing
  if (num_tiles == 0 && blockIdx.x == 0 && tid == 0) {
    st->level_begin[L + 1] = lend;
    st->level_end[L + 1]   = lend;
  }
  (void)done;
}

// Warp-per-node form of expand_level_kernel for the first levels, where there are few nodes but
// each spans a huge key range: the three child boundaries are searched CONCURRENTLY by three
// 10-lane groups, 11-ary (8 dependent loads for a 100 M range instead of 27).
constexpr int kExpandWarpNodes = kExpandBlock / 32;
constexpr u64 kWarpLevelNodes  = 2048;  // levels with at most this many cells use the warp kernel
__global__ void __launch_bounds__(kExpandBlock)
expand_level_warp_kernel(const u32* __restrict__ keys, int L, int max_depth, u32 max_size, u32 cap,
                         u32* __restrict__ okey, u8* __restrict__ olevel,
                         u8* __restrict__ ointernal, u32* __restrict__ olength,
                         u32* __restrict__ ooffset, tree_state* st, u64* __restrict__ lookback,
                         u32* __restrict__ ticket)
{
  __shared__ u32 s_tile, s_base, s_nchild[kExpandWarpNodes];
  u32 const lbeg      = st->level_begin[L];
  u32 const lend      = st->level_end[L];
  u32 const count     = lend - lbeg;
  u32 const num_tiles = (count + kExpandWarpNodes - 1) / kExpandWarpNodes;
  u32 const tag_agg = 2u * (L + 1), tag_pre = 2u * (L + 1) + 1u;
  int const shift   = 2 * (max_depth - 2 - L);
  int const tid = threadIdx.x, warp = tid >> 5;
  u32 const lane = lane_id();
  u32 const grp  = min(lane / 10u, 2u), li = lane - grp * 10u;  // lanes 30,31 shadow group 2
  bool const active_lane = lane < 30;

  while (true) {
    __syncthreads();
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    u32 const tile = s_tile;
    if (tile >= num_tiles) break;
    u32 const row = lbeg + tile * kExpandWarpNodes + warp;
    u32 k = 0, cnt = 0, start = 0, b1 = 0, b2 = 0, b3 = 0, nchild = 0;
    if (row < lend) {
      k     = okey[row];
      cnt   = olength[row];
      start = ooffset[row];
      if (cnt > max_size) {
        u64 const target = (((u64)k << 2) + grp + 1) << shift;
        u32 lo = start, hi = start + cnt;
        while (__any_sync(0xffffffffu, hi - lo > 10)) {
          bool less = false;
          u32 probe = lo;
          if (hi - lo > 10) {
            probe = lo + (u32)(((u64)(hi - lo) * (li + 1)) / 11);  // 10 interior probes
            less  = active_lane && (u64)__ldg(keys + probe) < target;
          }
          u32 const m  = (__ballot_sync(0xffffffffu, less) >> (10 * grp)) & 0x3FFu;
          int const c  = __popc(m);  // probes [0,c) of my group are < target
          u32 const pl = __shfl_sync(0xffffffffu, probe, grp * 10 + max(c - 1, 0));
          u32 const ph = __shfl_sync(0xffffffffu, probe, grp * 10 + min(c, 9));
          if (hi - lo > 10) {
            u32 const nlo = c == 0 ? lo : pl + 1;
            u32 const nhi = c == 10 ? hi : ph;
            lo = nlo;
            hi = nhi;
          }
        }
        // final step: at most 10 candidates per group, one per lane
        u32 const idx   = lo + li;
        bool const less = active_lane && idx < hi && (u64)__ldg(keys + idx) < target;
        u32 const m     = (__ballot_sync(0xffffffffu, less) >> (10 * grp)) & 0x3FFu;
        u32 const bound = lo + __popc(m);
        b1 = __shfl_sync(0xffffffffu, bound, 0);
        b2 = __shfl_sync(0xffffffffu, bound, 10);
        b3 = __shfl_sync(0xffffffffu, bound, 20);
        nchild = (b1 > start) + (b2 > b1) + (b3 > b2) + (start + cnt > b3);
      }
    }
    if (lane == 0) s_nchild[warp] = nchild;
    __syncthreads();
    u32 wbase = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kExpandWarpNodes; ++w) {
      if (w < warp) wbase += s_nchild[w];
      block_total += s_nchild[w];
    }
    if (tid == 0) s_base = lookback_exclusive(lookback, tile, block_total, tag_agg, tag_pre);
    __syncthreads();
    if (nchild && lane == 0) {
      u32 const first_child = lend + s_base + wbase;
      ointernal[row] = 1;
      olength[row]   = nchild;
      ooffset[row]   = first_child;
      u32 const b[5] = {start, b1, b2, b3, start + cnt};
      u32 r          = first_child;
      for (int c = 0; c < 4; ++c) {
        if (b[c + 1] > b[c]) {
          if (r < cap) {
            okey[r]      = (k << 2) + c;
            olevel[r]    = (u8)(L + 1);
            ointernal[r] = 0;
            olength[r]   = b[c + 1] - b[c];
            ooffset[r]   = b[c];
          } else {
            st->overflow = 1;
          }
          ++r;
        }
      }
    }
    if (tid == 0 && tile == num_tiles - 1) {
      st->level_begin[L + 1] = lend;
      st->level_end[L + 1]   = min(lend + s_base + block_total, cap);
    }
  }
  if (num_tiles == 0 && blockIdx.x == 0 && tid == 0) {
    st->level_begin[L + 1] = lend;
    st->level_end[L + 1]   = lend;
  }
}

// one launch instead of five device-to-device copies
__global__ void __launch_bounds__(256)
copy_tree_kernel(const u32* __restrict__ k, const u8* __restrict__ l, const u8* __restrict__ f,
                 const u32* __restrict__ n, const u32* __restrict__ o, u32 q, u32* __restrict__ ok,
                 u8* __restrict__ ol, u8* __restrict__ of, u32* __restrict__ on,
                 u32* __restrict__ oo)
{
  u32 const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q) {
    ok[i] = k[i]; ol[i] = l[i]; of[i] = f[i]; on[i] = n[i]; oo[i] = o[i];
  }
}

template <typename T>
void launch_encode(const void* x, const void* y, u64 n, double x_min, double x_max, double y_min,
                   double y_max, double scale_d, int max_depth, int passes, u32* keys, u32* hist,
                   u32* point_flags, bsj_grid* grid, cudaStream_t s, int bin_shift = 0,
                   u32 n_bins = 0)
{
  // the column API casts to T (cpp/src/indexing/point_quadtree.cu:82-84), the header API then
  // orders the corners and clamps scale in T (detail/point_quadtree.cuh:259-268)
  T const x1 = (T)x_min, x2 = (T)x_max, y1 = (T)y_min, y2 = (T)y_max;
  T const min_x = std::min(x1, x2), min_y = std::min(y1, y2);
  T const max_x = std::max(x1, x2), max_y = std::max(y1, y2);
  T const scale = std::max((T)scale_d, std::max(max_x - min_x, max_y - min_y) /
                                         (T)((1 << max_depth) + 2));
  u32 const oob_key = (u32)((1 << (2 * max_depth)) - 1);
  int const V       = 16 / sizeof(T);
  int const nblk    = (int)std::min<u64>((u64)num_sms() * 4, (u64)div_up(div_up(n, V), 512));
  auto launch = [&](auto pass_tag) {
    constexpr int P = decltype(pass_tag)::value;
    encode_hist_kernel<T, P><<<std::max(nblk, 1), 512, 0, s>>>(
      (const T*)x, (const T*)y, n, min_x, min_y, max_x, max_y, scale, oob_key, keys, hist,
      point_flags, bin_shift, n_bins);
  };
  switch (passes) {
    case -1: launch(std::integral_constant<int, -1>{}); break;
    case 0: launch(std::integral_constant<int, 0>{}); break;
    case 1: launch(std::integral_constant<int, 1>{}); break;
    case 2: launch(std::integral_constant<int, 2>{}); break;
    case 3: launch(std::integral_constant<int, 3>{}); break;
    default: launch(std::integral_constant<int, 4>{}); break;
  }
  BSJ_CHECK_LAUNCH();
  grid->valid     = 1;
  grid->max_depth = max_depth;
  grid->min_x = min_x; grid->min_y = min_y; grid->max_x = max_x; grid->max_y = max_y;
  grid->scale = scale;
}

}  // namespace

// Keys only (no histograms): used by the multi-GPU partitioner (partition.cu).
// `point_flags` (device, may be NULL): bit 0 is OR-ed in when a point lies outside the box, bit 1
// when a coordinate is NaN.
template <typename T>
void launch_point_keys(const void* x, const void* y, u64 n, double x_min, double x_max,
                       double y_min, double y_max, double scale, int max_depth, u32* keys,
                       u32* point_flags, cudaStream_t s, u32* lead_bins, int bin_shift,
                       u32 n_bins)
{
  dev_buf<u32> flags;
  if (!point_flags) {
    flags.alloc(1, s);
    BSJ_CUDA_TRY(cudaMemsetAsync(flags.get(), 0, sizeof(u32), s));
    point_flags = flags.get();
  }
  bsj_grid g{};
  // with `lead_bins` (n_bins <= 8192 counters of key >> bin_shift) the histogram is fused in
  bool const fused = lead_bins != nullptr && n_bins <= (u32)kLeadBins;
  launch_encode<T>(x, y, n, x_min, x_max, y_min, y_max, scale, max_depth,
                   /*passes=*/fused ? -1 : 0, keys, fused ? lead_bins : nullptr, point_flags, &g, s,
                   bin_shift, n_bins);
}
template void launch_point_keys<float>(const void*, const void*, u64, double, double, double,
                                       double, double, int, u32*, u32*, cudaStream_t, u32*, int,
                                       u32);
template void launch_point_keys<double>(const void*, const void*, u64, double, double, double,
                                        double, double, int, u32*, u32*, cudaStream_t, u32*, int,
                                        u32);

namespace {
// Tree rows from the sorted keys, then the exact-size copy-out.  `st->point_flags` carries the
// out-of-box / NaN flags of the encode (or of the caller, quadtree_on_keys).
void finish_tree(const u32* sorted_keys, u32* out_keys, u32* out_idx, u64 n, int d, u32 max_size,
                 tree_state* st_dev, bsj_grid grid, out_alloc& oa, stage_timer& tm, cudaStream_t s,
                 bsj_quadtree* out)
{
  // ---- tree rows. Capacity: every node below level 0 has a parent with > max_size points, so a
  // level holds at most 4*floor(N/(max_size+1)) nodes, and never more than N or (2^(L+1)+3)^2
  // cells (cell indices reach 2^d + 2 under the reference's scale clamp).
  u64 cap = 64;
  {
    u64 const per_level = std::min<u64>(n, 4 * (n / ((u64)max_size + 1)));
    for (int L = 1; L < d; ++L) {
      u64 const side = (2ull << L) + 3;
      cap += std::min(per_level, side * side);
    }
  }
  dev_buf<u32> tkey(cap, s), tlen(cap, s), toff(cap, s);
  dev_buf<u8> tlevel(cap, s), tint(cap, s);
  // one look-back descriptor per tile of the largest level: 256-node tiles in the thread kernel,
  // 8-node tiles in the warp kernel (which only runs on levels with <= kWarpLevelNodes nodes)
  u32 const max_tiles = (u32)std::max<u64>(div_up(cap, kExpandBlock),
                                           div_up(std::min<u64>(cap, kWarpLevelNodes), kExpandWarpNodes)) + 1;
  dev_buf<u64> lb(max_tiles, s);
  dev_buf<u32> tickets(16, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(lb.get(), 0, max_tiles * sizeof(u64), s));
  BSJ_CUDA_TRY(cudaMemsetAsync(tickets.get(), 0, 16 * sizeof(u32), s));

  int const shift0 = d >= 1 ? 2 * (d - 1) : 0;
  level0_kernel<<<1, 32, 0, s>>>(sorted_keys, (u32)n, shift0, (u32)cap, tkey.get(), tlevel.get(),
                                 tint.get(), tlen.get(), toff.get(), st_dev);
  BSJ_CHECK_LAUNCH();
  for (int L = 0; L + 1 < d; ++L) {
    // level L holds at most min(cap, (2^(L+1)+3)^2) nodes; size the grid for that
    u64 const geo  = ((2ull << L) + 3) * ((2ull << L) + 3);
    int const grid = (int)std::min<u64>((u64)num_sms() * 8,
                                        std::max<u64>(1, div_up(std::min<u64>(geo, cap), kExpandBlock)));
    if (geo <= kWarpLevelNodes) {  // few, huge nodes: warp-cooperative child search
      int const wgrid = (int)std::max<u64>(1, div_up(std::min<u64>(geo, cap), kExpandWarpNodes));
      expand_level_warp_kernel<<<wgrid, kExpandBlock, 0, s>>>(
        sorted_keys, L, d, max_size, (u32)cap, tkey.get(), tlevel.get(), tint.get(), tlen.get(),
        toff.get(), st_dev, lb.get(), tickets.get() + L);
    } else {
      expand_level_kernel<<<grid, kExpandBlock, 0, s>>>(
        sorted_keys, L, d, max_size, (u32)cap, tkey.get(), tlevel.get(), tint.get(), tlen.get(),
        toff.get(), st_dev, lb.get(), tickets.get() + L, nullptr);
    }
    BSJ_CHECK_LAUNCH();
  }
  tm.mark("tree_levels");

  tree_state h{};
  BSJ_CUDA_TRY(cudaMemcpyAsync(&h, st_dev, sizeof(h), cudaMemcpyDeviceToHost, s));
  BSJ_CUDA_TRY(cudaStreamSynchronize(s));
  if (h.overflow) throw error(BSJ_CUDA_ERROR, "quadtree node capacity exceeded (internal error)");
  int const last = d >= 2 ? d - 1 : 0;
  u64 const q    = h.level_end[last];

  grid.sorted_keys      = out_keys;
  grid.n_sorted_keys    = n;
  out->sorted_keys      = out_keys;
  grid.has_out_of_bbox  = (h.point_flags & 1u) ? 1 : 0;
  grid.has_nan          = (h.point_flags & 2u) ? 1 : 0;
  out->grid             = grid;
  out->point_indices    = out_idx;
  out->num_points       = n;
  out->num_nodes        = q;
  out->key              = oa.get<u32>(q);
  out->level            = oa.get<u8>(q);
  out->is_internal_node = oa.get<u8>(q);
  out->length           = oa.get<u32>(q);
  out->offset           = oa.get<u32>(q);
  if (q) {
    copy_tree_kernel<<<div_up(q, 256), 256, 0, s>>>(tkey.get(), tlevel.get(), tint.get(),
                                                   tlen.get(), toff.get(), (u32)q, out->key,
                                                   out->level, out->is_internal_node, out->length,
                                                   out->offset);
    BSJ_CHECK_LAUNCH();
  }
  tm.mark("finalize");
  tm.finish();  // results are ready in stream order; no trailing host synchronisation
  oa.commit();
}
}  // namespace

// Host orchestration.  One stream synchronisation at the end (to learn the node count).
void quadtree_on_points_impl(const void* x, const void* y, int dtype, u64 n, double x_min,
                             double x_max, double y_min, double y_max, double scale,
                             int max_depth_in, int max_size_in, const bsj_allocator* mr,
                             cudaStream_t s, bsj_quadtree* out)
{
  *out = bsj_quadtree{};
  if (n == 0) return;  // point_quadtree.cu:167-177
  BSJ_EXPECTS(n < 0xFFFFC000ull, "number of points must fit uint32 indices");

  u32 const max_size = (u32)std::max(1, max_size_in);                   // :264
  int const d        = std::max(0, std::min(15, max_depth_in));         // :266
  stage_timer tm(s);

  // key width: in-bbox cell indices are <= 2^d + 2 (scale clamp), i.e. d+2 bits per axis
  int const key_bits = std::min(32, 2 * (d + 2));
  int const passes   = passes_for_bits(0, key_bits);

  out_alloc oa(mr, s);
  // the sorted permutation ends in `idx_a` or `idx_b` depending on pass parity; make the final
  // target the caller-visible output buffer
  u32* out_idx = oa.get<u32>(n);
  // likewise the sorted keys end in side A (even number of passes) or B: that side is allocated
  // through the output allocator and handed out as part of the bsj_grid hint
  bool const even_passes = (passes % 2) == 0;
  u32* out_keys = oa.get<u32>(n);
  dev_buf<u32> keys_tmp(n, s), idx_tmp(n, s);
  struct { u32* p; u32* get() const { return p; } } keys_a{even_passes ? out_keys : keys_tmp.get()},
    keys_b{even_passes ? keys_tmp.get() : out_keys};
  sort_workspace ws;
  ws.alloc(n, s);
  sort_workspace_reset(ws, s);

  dev_buf<tree_state> st(1, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(st.get(), 0, sizeof(tree_state), s));
  bsj_grid grid{};
  if (dtype == BSJ_FLOAT32)
    launch_encode<float>(x, y, n, x_min, x_max, y_min, y_max, scale, d, passes, keys_a.get(),
                         ws.hist.get(), &st.get()->point_flags, &grid, s);
  else
    launch_encode<double>(x, y, n, x_min, x_max, y_min, y_max, scale, d, passes, keys_a.get(),
                          ws.hist.get(), &st.get()->point_flags, &grid, s);
  tm.mark("encode_hist");

  // ping-pong so that the last pass writes into out_idx: with P passes the result lands in
  // side A when P is even, side B when odd.
  bool const even = (passes % 2) == 0;
  u32* vals_a     = even ? out_idx : idx_tmp.get();
  u32* vals_b     = even ? idx_tmp.get() : out_idx;
  bool in_a       = true;
  sort_passes(keys_a.get(), vals_a, /*iota=*/true, keys_b.get(), vals_b, n, 0, key_bits, ws, s,
              &in_a);
  const u32* sorted_keys = in_a ? keys_a.get() : keys_b.get();

  finish_tree(sorted_keys, out_keys, out_idx, n, d, max_size, st.get(), grid, oa, tm, s, out);
}

// Quadtree from Morton keys that were computed elsewhere (multi-GPU: the keys a rank RECEIVES for
// its key range, with the points' global ids as sort payload).  Same stable sort and the same tree
// rows as quadtree_on_points on the corresponding points; out->point_indices = the payload in
// sorted order.  `keys` and `values` are scratch: the sort overwrites them.
void quadtree_on_keys_impl(u32* keys, u32* values, u64 n, const bsj_grid* g, int max_size_in,
                           const bsj_allocator* mr, cudaStream_t s, bsj_quadtree* out)
{
  *out = bsj_quadtree{};
  if (n == 0) return;
  BSJ_EXPECTS(n < 0xFFFFC000ull, "number of points must fit uint32 indices");
  BSJ_EXPECTS(g != nullptr && g->valid, "key geometry (bsj_grid) must be given");
  u32 const max_size = (u32)std::max(1, max_size_in);
  int const d        = std::max(0, std::min(15, (int)g->max_depth));
  stage_timer tm(s);
  int const key_bits = std::min(32, 2 * (d + 2));
  out_alloc oa(mr, s);
  u32* out_idx  = oa.get<u32>(n);
  u32* out_keys = oa.get<u32>(n);
  dev_buf<u32> keys_tmp(n, s), vals_tmp(n, s);
  sort_workspace ws;
  ws.alloc(n, s);
  sort_workspace_reset(ws, s);
  sort_histogram(keys, n, 0, key_bits, ws, s);
  tm.mark("key_hist");
  sort_passes_to(keys, values, keys_tmp.get(), vals_tmp.get(), out_keys, out_idx, n, 0, key_bits,
                 ws, s);
  dev_buf<tree_state> st(1, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(st.get(), 0, sizeof(tree_state), s));
  u32 const flags = (g->has_out_of_bbox ? 1u : 0u) | (g->has_nan ? 2u : 0u);
  BSJ_CUDA_TRY(cudaMemcpyAsync(&st.get()->point_flags, &flags, sizeof(u32), cudaMemcpyHostToDevice,
                               s));
  bsj_grid grid = *g;
  finish_tree(out_keys, out_keys, out_idx, n, d, max_size, st.get(), grid, oa, tm, s, out);
}

}  // namespace bsj

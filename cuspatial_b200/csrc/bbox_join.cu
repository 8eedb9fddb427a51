// bbox_join.cu -- B200-native bounding-box x quadtree filter
// (replaces cuspatial::join_quadtree_and_bounding_boxes).
//
// Reference behaviour restated: cpp/include/cuspatial/detail/join/quadtree_bbox_filtering.cuh:35-188
// (level-synchronous BFS, final stable sort by quadtree.offset[node]),
// detail/join/intersection.cuh:94-128 (node bounds + overlap classification),
// detail/join/traversal.cuh:63-145 (descent to children).
//
// Design (not a port): the reference runs ~12 Thrust launches and >= 3 host synchronisations per
// level (latency bound).  Here ONE persistent kernel walks the tree level by level: the (box, node)
// work items of a level sit in a device queue, every warp takes 32 of them at a time (one 128-bit
// node load per lane from a packed copy of the tree), appends leaf hits to the global pair list
// and the children of internal hits to the next level's queue with one warp-aggregated atomic
// each, and a grid-wide barrier separates the levels.  Work is balanced at 32-item granularity
// whatever the shape of the tree (clustered data: a few very deep subtrees) or the number of
// boxes.  The reference's output order (offset[node] ascending, ties by box index -- Appendix A.2
// of SURVEY.md) is then restored by two stable radix sorts (box, then leaf offset) of the small
// pair list, which also makes the result independent of the atomics' arrival order.
#include "radix_sort.cuh"

namespace bsj {

namespace {

template <typename T>
struct fpj;
template <>
struct fpj<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
};
template <>
struct fpj<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
};

// z_order.cuh:80-94 (arithmetic form)
__device__ __forceinline__ u32 undilate16(u32 v)
{
  v &= 0x55555555u;
  v = (v | (v >> 1)) & 0x33333333u;
  v = (v | (v >> 2)) & 0x0F0F0F0Fu;
  v = (v | (v >> 4)) & 0x00FF00FFu;
  v = (v | (v >> 8)) & 0x0000FFFFu;
  return v;
}

struct join_state {
  unsigned long long cursor0;  // level-0 items handed out so far (boxes x level-0 nodes: 64-bit)
  u32 n_top;        // number of level-0 nodes (quadtree_bbox_filtering.cuh:53-56)
  u32 n_hits;       // leaf hits found by the traversal
  u32 overflow;     // a level's queue did not fit its buffer
  u32 barrier;      // grid barrier arrivals (monotonic)
  u32 count[17];    // items queued for level L
  u32 cursor[17];   // items of level L handed out so far
};

// SoA tree -> one 16-byte record per node; counts level-0 nodes on the way.
__global__ void __launch_bounds__(256)
pack_tree_kernel(const u32* __restrict__ key, const u8* __restrict__ level,
                 const u8* __restrict__ internal, const u32* __restrict__ length,
                 const u32* __restrict__ offset, u32 q, uint4* __restrict__ nodes, join_state* st)
{
  u32 const i = blockIdx.x * blockDim.x + threadIdx.x;
  bool top    = false;
  if (i < q) {
    u32 const lv = level[i];
    nodes[i]     = make_uint4(key[i], lv | ((u32)(internal[i] != 0) << 8), length[i], offset[i]);
    top          = lv == 0;
  }
  u32 const m = __ballot_sync(0xffffffffu, top);
  if (lane_id() == 0 && m) atomicAdd(&st->n_top, (u32)__popc(m));
}

constexpr int kJoinBlock = 256;

// all CTAs of the (co-resident) grid: arrive, then wait for everybody
__device__ __forceinline__ void grid_barrier(u32* counter, u32 generation)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    u32 const target = (generation + 1) * gridDim.x;
    while (*(volatile u32*)counter < target) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kJoinBlock)
traverse_kernel(const uint4* __restrict__ nodes, const T* __restrict__ bx0,
                const T* __restrict__ by0, const T* __restrict__ bx1, const T* __restrict__ by1,
                u32 n_boxes, T vmin_x, T vmin_y, T scale, int max_depth,
                u32* __restrict__ out_box, u32* __restrict__ out_node, u32 capacity,
                join_state* st, uint2* __restrict__ queue_a, uint2* __restrict__ queue_b,
                u32 q_capacity)
{
  u32 const lane = lane_id();
  u32 const lt   = lanemask_lt();
  u32 const n_top = st->n_top;
  // level 0: every (box, level-0 node) combination, box-major (quadtree_bbox_filtering.cuh:94-102)
  u64 n_in = (u64)n_boxes * n_top;
  for (int level = 0; level < max_depth && n_in > 0; ++level) {
    const uint2* const q_in = (level & 1) ? queue_b : queue_a;
    uint2* const q_out      = (level & 1) ? queue_a : queue_b;
    while (true) {
      u64 base = 0;
      if (lane == 0) base = level == 0 ? atomicAdd(&st->cursor0, 32ull)
                                       : (u64)atomicAdd(&st->cursor[level], 32u);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= n_in) break;
      u64 const i     = base + lane;
      bool const have = i < n_in;
      u32 box = 0, node = 0;
      if (have) {
        if (level == 0) {
          box  = (u32)(i / n_top);
          node = (u32)(i % n_top);
        } else {
          uint2 const it = q_in[i];
          box  = it.x;
          node = it.y;
        }
      }
      bool leaf_hit = false;
      u32 nchild = 0, child0 = 0;
      if (have) {
        T const qx0 = __ldg(bx0 + box), qy0 = __ldg(by0 + box);
        T const qx1 = __ldg(bx1 + box), qy1 = __ldg(by1 + box);
        uint4 const nd   = __ldg(nodes + node);
        u32 const lv     = nd.y & 0xFFu;
        bool const inner = (nd.y >> 8) & 1u;
        // intersection.cuh:104-127; the adds/multiplies below are the ones nvcc emits for the
        // reference under its default flags: one multiply for level_scale, FMAs for the bounds.
        T const kx  = (T)undilate16(nd.x);
        T const ky  = (T)undilate16(nd.x >> 1);
        int const sh = max(0, max_depth - 1 - (int)lv);  // lv > max_depth-1 is UB in the reference
        T const ls   = fpj<T>::mul(scale, (T)(1 << sh));
        T const nx0 = fpj<T>::fma(kx, ls, vmin_x);
        T const ny0 = fpj<T>::fma(ky, ls, vmin_y);
        T const nx1 = fpj<T>::fma(fpj<T>::add(kx, (T)1), ls, vmin_x);
        T const ny1 = fpj<T>::fma(fpj<T>::add(ky, (T)1), ls, vmin_y);
        bool const miss = (nx0 > qx1) || (nx1 < qx0) || (ny0 > qy1) || (ny1 < qy0);
        if (!miss) {
          if (!inner) {
            leaf_hit = true;
          } else if (level + 1 < max_depth) {  // quadtree_bbox_filtering.cuh:116 loop bound
            nchild = min(nd.z, 4u);
            child0 = nd.w;
          }
        }
      }
      // ---- leaf hits: one atomic per warp step
      u32 const m = __ballot_sync(0xffffffffu, leaf_hit);
      if (m) {
        u32 obase = 0;
        if (lane == 0) obase = atomicAdd(&st->n_hits, (u32)__popc(m));
        obase = __shfl_sync(0xffffffffu, obase, 0);
        if (leaf_hit) {
          u32 const o = obase + __popc(m & lt);
          if (o < capacity) {
            out_box[o]  = box;
            out_node[o] = node;
          }
        }
      }
      // ---- children of internal hits: the next level's queue
      u32 const incl  = warp_inclusive_scan(nchild);
      u32 const total = __shfl_sync(0xffffffffu, incl, 31);
      if (total) {
        u32 qbase = 0;
        if (lane == 0) qbase = atomicAdd(&st->count[level + 1], total);
        qbase = __shfl_sync(0xffffffffu, qbase, 0);
        u64 const at = (u64)qbase + incl - nchild;
        if (at + nchild <= (u64)q_capacity) {
          for (u32 c = 0; c < nchild; ++c) q_out[at + c] = make_uint2(box, child0 + c);
        } else if (nchild) {
          st->overflow = 1;
        }
      }
    }
    grid_barrier(&st->barrier, (u32)level);
    n_in = min(st->count[level + 1], q_capacity);
  }
}

__global__ void __launch_bounds__(256)
gather_offsets_kernel(const u32* __restrict__ node, const uint4* __restrict__ nodes, u32 p,
                      u32* __restrict__ keys)
{
  u32 const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p) keys[i] = __ldg(nodes + node[i]).w;
}

__global__ void __launch_bounds__(256)
gather_pairs_kernel(const u32* __restrict__ perm, const u32* __restrict__ box,
                    const u32* __restrict__ node, u32 p, u32* __restrict__ out_box,
                    u32* __restrict__ out_node)
{
  u32 const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p) {
    u32 const j = perm[i];
    out_box[i]  = box[j];
    out_node[i] = node[j];
  }
}

int bits_for(u64 max_value)
{
  int b = 1;
  while (b < 32 && (max_value >> b)) ++b;
  return b;
}

template <typename T>
void join_impl_t(const u32* key, const u8* level, const u8* internal, const u32* length,
                 const u32* offset, u64 q, const void* bx0, const void* by0, const void* bx1,
                 const void* by1, u64 n_boxes, double x_min, double y_min, double scale,
                 int max_depth, const bsj_allocator* mr, cudaStream_t s, bsj_pairs* out)
{
  stage_timer tm(s);
  dev_buf<uint4> nodes(q, s);
  dev_buf<join_state> st(1, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(st.get(), 0, sizeof(join_state), s));
  pack_tree_kernel<<<div_up(q, 256), 256, 0, s>>>(key, level, internal, length, offset, (u32)q,
                                                  nodes.get(), st.get());
  BSJ_CHECK_LAUNCH();

  // The grid barrier needs every CTA resident at once: size the grid from the occupancy.
  int per_sm = 1;
  BSJ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, traverse_kernel<T>,
                                                            kJoinBlock, 0));
  int const grid = std::max(1, std::min(per_sm, 4) * num_sms());
  // optimistic capacities (pair list and per-level queues); a second traversal runs only if
  // they were too small
  u64 capacity = std::min<u64>(std::max<u64>(std::max<u64>(1u << 20, n_boxes * 64), q * 3),
                               0xFFFFFFF0ull);
  dev_buf<u32> hit_box, hit_node;
  dev_buf<uint2> queue_a, queue_b;
  join_state h{};
  for (int attempt = 0; attempt < 3; ++attempt) {
    hit_box.alloc(capacity, s);
    hit_node.alloc(capacity, s);
    queue_a.alloc(capacity, s);
    queue_b.alloc(capacity, s);
    traverse_kernel<T><<<grid, kJoinBlock, 0, s>>>(
      nodes.get(), (const T*)bx0, (const T*)by0, (const T*)bx1, (const T*)by1, (u32)n_boxes,
      (T)x_min, (T)y_min, (T)scale, max_depth, hit_box.get(), hit_node.get(), (u32)capacity,
      st.get(), queue_a.get(), queue_b.get(), (u32)capacity);
    BSJ_CHECK_LAUNCH();
    BSJ_CUDA_TRY(cudaMemcpyAsync(&h, st.get(), sizeof(h), cudaMemcpyDeviceToHost, s));
    BSJ_CUDA_TRY(cudaStreamSynchronize(s));
    u64 need = h.n_hits;
    for (int L = 0; L < 17; ++L) need = std::max<u64>(need, h.count[L]);
    if (!h.overflow && need <= capacity) break;
    if (attempt == 2 || capacity >= 0xFFFFFFF0ull)
      throw error(BSJ_INVALID_ARGUMENT,
                  "quadtree traversal work list overflow (malformed quadtree table?)");
    // a truncated level under-reports the levels below it: grow generously
    capacity = std::min<u64>(std::max<u64>(need, capacity) * 4, 0xFFFFFFF0ull);
    u32 const n_top = h.n_top;
    BSJ_CUDA_TRY(cudaMemsetAsync(st.get(), 0, sizeof(join_state), s));
    BSJ_CUDA_TRY(cudaMemcpyAsync(&st.get()->n_top, &n_top, sizeof(u32), cudaMemcpyHostToDevice, s));
  }
  tm.mark("traverse");
  u64 const p = h.n_hits;
  out_alloc oa(mr, s);
  out->size = p;
  if (p == 0) {
    tm.finish();
    return;
  }
  out->first  = oa.get<u32>(p);
  out->second = oa.get<u32>(p);

  // ---- order: stable by box, then stable by offset[node]  ==  (offset, box) lexicographic
  dev_buf<u32> k2(capacity, s), v2(capacity, s);
  sort_workspace ws;
  ws.alloc(p, s);
  bool in_a = true;
  {
    int const bits = bits_for(n_boxes ? n_boxes - 1 : 0);
    sort_workspace_reset(ws, s);
    sort_histogram(hit_box.get(), p, 0, bits, ws, s);
    sort_passes(hit_box.get(), hit_node.get(), false, k2.get(), v2.get(), p, 0, bits, ws, s, &in_a,
                "pair_sort_pass");
  }
  u32* box_sorted  = in_a ? hit_box.get() : k2.get();
  u32* node_sorted = in_a ? hit_node.get() : v2.get();
  u32* spare_k     = in_a ? k2.get() : hit_box.get();
  u32* spare_v     = in_a ? v2.get() : hit_node.get();
  dev_buf<u32> okeys(p, s), perm_b(p, s);
  gather_offsets_kernel<<<div_up(p, 256), 256, 0, s>>>(node_sorted, nodes.get(), (u32)p,
                                                       okeys.get());
  BSJ_CHECK_LAUNCH();
  {
    sort_workspace_reset(ws, s);
    sort_histogram(okeys.get(), p, 0, 32, ws, s);
    // values = iota; spare_v / perm_b are the value ping-pong buffers
    sort_passes(okeys.get(), spare_v, true, spare_k, perm_b.get(), p, 0, 32, ws, s, &in_a,
                "pair_sort_pass");
  }
  u32* perm = in_a ? spare_v : perm_b.get();
  gather_pairs_kernel<<<div_up(p, 256), 256, 0, s>>>(perm, box_sorted, node_sorted, (u32)p,
                                                     out->first, out->second);
  BSJ_CHECK_LAUNCH();
  tm.mark("order_pairs");
  tm.finish();  // results are ready in stream order; no trailing host synchronisation
  oa.commit();
}

}  // namespace

void join_quadtree_and_bounding_boxes_impl(const u32* key, const u8* level, const u8* internal,
                                           const u32* length, const u32* offset, u64 q,
                                           const void* bx0, const void* by0, const void* bx1,
                                           const void* by1, int dtype, u64 n_boxes, double x_min,
                                           double x_max, double y_min, double y_max, double scale,
                                           int max_depth, const bsj_allocator* mr, cudaStream_t s,
                                           bsj_pairs* out)
{
  *out = bsj_pairs{};
  // cpp/src/join/quadtree_bbox_filtering.cu:100-106
  BSJ_EXPECTS(scale > 0, "scale must be positive");
  BSJ_EXPECTS(x_min < x_max && y_min < y_max, "invalid bounding box (x_min, x_max, y_min, y_max)");
  BSJ_EXPECTS(max_depth > 0 && max_depth < 16, "maximum depth must be positive and less than 16");
  if (q == 0 || n_boxes == 0) return;  // :108-114
  BSJ_EXPECTS(q < 0xFFFFFFFFull && n_boxes < 0xFFFFFFFFull, "table too large");
  if (dtype == BSJ_FLOAT32)
    join_impl_t<float>(key, level, internal, length, offset, q, bx0, by0, bx1, by1, n_boxes, x_min,
                       y_min, scale, max_depth, mr, s, out);
  else
    join_impl_t<double>(key, level, internal, length, offset, q, bx0, by0, bx1, by1, n_boxes,
                        x_min, y_min, scale, max_depth, mr, s, out);
}

}  // namespace bsj

// bbox_join.cu -- B200-native bounding-box x quadtree filter
// (replaces cuspatial::join_quadtree_and_bounding_boxes).
//
// Reference behaviour restated: cpp/include/cuspatial/detail/join/quadtree_bbox_filtering.cuh:35-188
// (level-synchronous BFS, final stable sort by quadtree.offset[node]),
// detail/join/intersection.cuh:94-128 (node bounds + overlap classification),
// detail/join/traversal.cuh:63-145 (descent to children).
//
// Design (not a port): the reference runs ~12 Thrust launches and >= 3 host synchronisations per
// level (latency bound).  Here ONE kernel walks the tree: one warp per bounding box keeps a LIFO
// work list of node indices in shared memory, tests 32 nodes per step (one 128-bit node load per
// lane from a packed copy of the tree), appends leaf hits to the global pair list with one
// warp-aggregated atomic per step and pushes the children of internal hits back with a warp scan.
// The reference's output order (offset[node] ascending, ties by box index -- Appendix A.2 of
// SURVEY.md) is then restored by two stable radix sorts (box, then leaf offset) of the small pair
// list, which also makes the result independent of the atomic's arrival order.
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
  u32 n_top;     // number of level-0 nodes (quadtree_bbox_filtering.cuh:53-56)
  u32 n_hits;    // leaf hits found by the traversal
  u32 overflow;  // work-list overflow (malformed tree)
  u32 n_seeds;   // (box, node) sub-traversals queued by the seeding pass
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

constexpr int kJoinWarps    = 4;
constexpr int kJoinStackCap = 4096;  // u32 entries per warp

template <typename T>
__global__ void __launch_bounds__(kJoinWarps * 32)
traverse_kernel(const uint4* __restrict__ nodes, const T* __restrict__ bx0,
                const T* __restrict__ by0, const T* __restrict__ bx1, const T* __restrict__ by1,
                u32 n_boxes, T vmin_x, T vmin_y, T scale, int max_depth,
                u32* __restrict__ out_box, u32* __restrict__ out_node, u32 capacity,
                join_state* st, const u32* __restrict__ seed_box,
                const u32* __restrict__ seed_node, int stop_level, u32* __restrict__ q_box,
                u32* __restrict__ q_node, u32 q_capacity)
{
  extern __shared__ u32 s_stack_all[];
  int const warp      = threadIdx.x >> 5;
  u32 const lane      = lane_id();
  u32 const lt        = lanemask_lt();
  u32* const stack    = s_stack_all + warp * kJoinStackCap;
  u32 const n_top     = min(st->n_top, (u32)kJoinStackCap);
  u32 const num_warps = gridDim.x * kJoinWarps;

  // Two launches of this kernel.  Seeding pass (seed_box == nullptr): one warp per (bounding
  // box, level-0 node) descends to `stop_level`; internal nodes hit there are not expanded but
  // queued as (box, node) seeds.  Main pass: one warp per seed traverses that subtree.  A box then
  // spreads over as many warps as it overlaps level-`stop_level` cells instead of serialising its
  // whole (latency-bound) traversal in one warp.  stop_level < 0: single pass, no queue.
  bool const seeded = seed_box != nullptr;
  u64 const n_units = seeded ? (u64)min(st->n_seeds, q_capacity) : (u64)n_boxes * max(n_top, 1u);
  for (u64 unit = (u64)blockIdx.x * kJoinWarps + warp; unit < n_units; unit += num_warps) {
    u32 const box = seeded ? __ldg(seed_box + unit) : (u32)(unit / max(n_top, 1u));
    T const qx0 = __ldg(bx0 + box), qy0 = __ldg(by0 + box);
    T const qx1 = __ldg(bx1 + box), qy1 = __ldg(by1 + box);
    if (lane == 0) stack[0] = seeded ? __ldg(seed_node + unit) : (u32)(unit % max(n_top, 1u));
    u32 sp = (seeded || n_top) ? 1u : 0u;
    __syncwarp();
    while (sp > 0) {
      u32 const take  = min(sp, 32u);
      u32 const base  = sp - take;
      bool const have = lane < take;
      u32 const node  = have ? stack[base + lane] : 0u;
      sp              = base;
      __syncwarp();

      bool leaf_hit = false, seed_hit = false;
      u32 nchild = 0, child0 = 0;
      if (have) {
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
          } else if ((int)lv + 1 < max_depth) {  // quadtree_bbox_filtering.cuh:116 loop bound
            if (!seeded && (int)lv == stop_level) {
              seed_hit = true;
            } else {
              nchild = nd.z;
              child0 = nd.w;
            }
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
      // ---- seeds for the main pass
      u32 const sm_ = __ballot_sync(0xffffffffu, seed_hit);
      if (sm_) {
        u32 qbase = 0;
        if (lane == 0) qbase = atomicAdd(&st->n_seeds, (u32)__popc(sm_));
        qbase = __shfl_sync(0xffffffffu, qbase, 0);
        if (seed_hit) {
          u32 const o = qbase + __popc(sm_ & lt);
          if (o < q_capacity) {
            q_box[o]  = box;
            q_node[o] = node;
          } else {
            st->overflow = 1;  // more level-k nodes than 4^(k+1) per box: malformed tree
          }
        }
      }
      // ---- children of internal hits go back on the work list
      u32 const incl  = warp_inclusive_scan(nchild);
      u32 const total = __shfl_sync(0xffffffffu, incl, 31);
      if (total) {
        if (sp + total > (u32)kJoinStackCap) {
          if (lane == 0) st->overflow = 1;
          sp = 0;  // abandon this box; the host reports the error
        } else {
          u32 const at = sp + incl - nchild;
          for (u32 c = 0; c < nchild; ++c) stack[at + c] = child0 + c;
          sp += total;
        }
      }
      __syncwarp();
    }
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
  configure_once_per_device(1, [] {  // per device, not per process
    BSJ_CUDA_TRY(cudaFuncSetAttribute(traverse_kernel<float>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kJoinWarps * kJoinStackCap * 4));
    BSJ_CUDA_TRY(cudaFuncSetAttribute(traverse_kernel<double>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kJoinWarps * kJoinStackCap * 4));
  });
  dev_buf<uint4> nodes(q, s);
  dev_buf<join_state> st(1, s);
  BSJ_CUDA_TRY(cudaMemsetAsync(st.get(), 0, sizeof(join_state), s));
  pack_tree_kernel<<<div_up(q, 256), 256, 0, s>>>(key, level, internal, length, offset, (u32)q,
                                                  nodes.get(), st.get());
  BSJ_CHECK_LAUNCH();

  // optimistic capacity; a second traversal runs only if it was too small
  u64 capacity = std::max<u64>(1u << 20, n_boxes * 64);
  dev_buf<u32> hit_box, hit_node;
  join_state h{};
  int const grid = (int)std::min<u64>((u64)num_sms() * 3, (u64)div_up(n_boxes * 4, kJoinWarps));
  // seeding level k: at most 4^(k+1) level-k nodes per box; keep the queue below 2^24 entries
  int stop_level = n_boxes <= (1u << 16) ? 3 : n_boxes <= (1u << 18) ? 2 : 1;
  if (stop_level + 2 >= max_depth) stop_level = -1;  // shallow tree: one pass does it all
  u64 const q_cap = stop_level >= 0 ? n_boxes << (2 * (stop_level + 1)) : 0;
  dev_buf<u32> q_box(std::max<u64>(q_cap, 1), s), q_node(std::max<u64>(q_cap, 1), s);
  for (int attempt = 0; attempt < 2; ++attempt) {
    hit_box.alloc(capacity, s);
    hit_node.alloc(capacity, s);
    traverse_kernel<T><<<grid, kJoinWarps * 32, kJoinWarps * kJoinStackCap * 4, s>>>(
      nodes.get(), (const T*)bx0, (const T*)by0, (const T*)bx1, (const T*)by1, (u32)n_boxes,
      (T)x_min, (T)y_min, (T)scale, max_depth, hit_box.get(), hit_node.get(), (u32)capacity,
      st.get(), nullptr, nullptr, stop_level, q_box.get(), q_node.get(), (u32)q_cap);
    BSJ_CHECK_LAUNCH();
    if (stop_level >= 0) {
      traverse_kernel<T><<<num_sms() * 3, kJoinWarps * 32, kJoinWarps * kJoinStackCap * 4, s>>>(
        nodes.get(), (const T*)bx0, (const T*)by0, (const T*)bx1, (const T*)by1, (u32)n_boxes,
        (T)x_min, (T)y_min, (T)scale, max_depth, hit_box.get(), hit_node.get(), (u32)capacity,
        st.get(), q_box.get(), q_node.get(), -1, nullptr, nullptr, (u32)q_cap);
      BSJ_CHECK_LAUNCH();
    }
    BSJ_CUDA_TRY(cudaMemcpyAsync(&h, st.get(), sizeof(h), cudaMemcpyDeviceToHost, s));
    BSJ_CUDA_TRY(cudaStreamSynchronize(s));
    if (h.overflow)
      throw error(BSJ_INVALID_ARGUMENT,
                  "quadtree traversal work list overflow (malformed quadtree table?)");
    if (h.n_hits <= capacity) break;
    capacity = h.n_hits;
    BSJ_CUDA_TRY(cudaMemsetAsync(&st.get()->n_hits, 0, sizeof(u32), s));
    BSJ_CUDA_TRY(cudaMemsetAsync(&st.get()->n_seeds, 0, sizeof(u32), s));
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

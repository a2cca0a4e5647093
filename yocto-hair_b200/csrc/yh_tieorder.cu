// The reference's leaf order, built on the device.
//
// When two primitives of a shape are hit at exactly the same t, the reference keeps the one its own BVH visits last
// (yh_scene.cuh "tie order").  That order is the leaf order of ITS tree: build_bvh / split_middle
// (yocto_pathtrace.cpp:564-650) - centroid bounds, largest axis, std::partition about the spatial middle, median when
// that fails, leaves of <= 4 - and it includes the permutation libstdc++'s std::partition leaves inside each half.
// Round 1 replayed those splits on host threads (~40 ms per million primitives, in every process of a multi-GPU
// run); here the same tree is grown level by level on the GPU, a few milliseconds per million primitives:
//
//   per level, all open ranges ("segments") at once
//     bounds of the centres per segment            atomics (warp-aggregated), exact min / max
//     axis, middle, predicate centre[axis] < middle  the reference's float arithmetic (this TU: --fmad=false)
//     std::partition's permutation                 bidirectional std::partition swaps the i-th misplaced element
//                                                  from the left with the i-th misplaced from the right: two prefix
//                                                  sums give every misplaced element its rank, the pairs swap in place
//     children > 4 primitives open the next level
//
// Output: ref_pos[e] = position of element e in that order, tie_nodes[k] = {mid, axis, child0, child1} (child < 0: a
// leaf) as reference_visits_later() walks them, and the exact root box (the union of the primitive bounds) the
// top-level tree over the objects is built from.  tests/test_tieorder_gpu.py compares the order with the plain-C port's
// restatement of the reference builder, element for element.
#include <cub/cub.cuh>

#include "yh_scene.cuh"

namespace yh {

__device__ __forceinline__ vec3f ldv3(const float* p, int i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

// float min / max through integer atomics: non-negative floats order like signed ints, negative ones inversely like
// unsigned ints
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
  if (v >= 0) atomicMin((int*)a, __float_as_int(v));
  else atomicMax((unsigned*)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
  if (v >= 0) atomicMax((int*)a, __float_as_int(v));
  else atomicMin((unsigned*)a, __float_as_uint(v));
}
// min / max of (lo, hi) into box[0..2] / box[3..5]: one set of atomics per warp when the whole warp adds to the same box
__device__ __forceinline__ void box_add(float* box, bool live, bool uniform, vec3f lo, vec3f hi) {
  if (uniform) {
    if (!live) lo = {flt_max, flt_max, flt_max}, hi = {-flt_max, -flt_max, -flt_max};
    for (int d = 16; d > 0; d >>= 1) {
      lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, d)), lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, d));
      lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, d)), hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, d));
      hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, d)), hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, d));
    }
    live = (threadIdx.x & 31) == 0;
  }
  if (!live) return;
  atomic_min_f(box + 0, lo.x), atomic_min_f(box + 1, lo.y), atomic_min_f(box + 2, lo.z);
  atomic_max_f(box + 3, hi.x), atomic_max_f(box + 4, hi.y), atomic_max_f(box + 5, hi.z);
}

// item = {centre.xyz, primitive id}: centre = center(bounds) as init_bvh computes it (:715-742, yocto_math.h:3008)
__global__ void __launch_bounds__(256) k_tie_items(int kind, const int* elems, const float* pos, const float* radius, int n,
    float4* items, int* seg_of, float* root_box) {
  int    i    = blockIdx.x * blockDim.x + threadIdx.x;
  bool   live = i < n;
  bbox3f b    = {{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
  if (live) {
    if (kind == 2) {
      int a = elems[i];
      b     = point_bounds(ldv3(pos, a), radius[a]);
    } else if (kind == 1) {
      int a = elems[2 * i], c = elems[2 * i + 1];
      b     = line_bounds(ldv3(pos, a), ldv3(pos, c), radius[a], radius[c]);
    } else {
      b = triangle_bounds(ldv3(pos, elems[3 * i]), ldv3(pos, elems[3 * i + 1]), ldv3(pos, elems[3 * i + 2]));
    }
    vec3f c    = (b.min + b.max) / 2.0f;
    items[i]   = make_float4(c.x, c.y, c.z, __int_as_float(i));
    seg_of[i]  = n > kRefLeafMax ? 0 : -1;
  }
  box_add(root_box, live, true, b.min, b.max);
}

// per-segment scratch of one level
struct tie_segments {
  int*   start;   // first position
  int*   end;     // one past the last
  int*   node;    // index of the segment's node in tie_nodes
  float* box;     // 6 per segment: bounds of the centres
  int*   axis;
  float* middle;
  int*   count;   // elements with centre[axis] < middle
  int*   mid;     // split position
  int*   nopart;  // 1: the range is halved by index, nothing moves
  int*   nchild;  // children that open a segment on the next level (0..2), then its exclusive scan
};

__global__ void __launch_bounds__(256) k_tie_seg_reset(tie_segments s, int num) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= num) return;
  for (int c = 0; c < 3; c++) s.box[6 * k + c] = flt_max, s.box[6 * k + 3 + c] = -flt_max;
  s.count[k] = 0;
}
__global__ void __launch_bounds__(256) k_tie_seg_bounds(const float4* items, const int* seg_of, int n, tie_segments s) {
  int  i    = blockIdx.x * blockDim.x + threadIdx.x;
  int  seg  = i < n ? seg_of[i] : -1;
  int  seg0 = __shfl_sync(0xffffffffu, seg, 0);
  bool uniform = __all_sync(0xffffffffu, seg == seg0 || seg < 0) && seg0 >= 0;
  if (!uniform && seg < 0) return;
  float4 it = seg >= 0 ? items[i] : make_float4(0, 0, 0, 0);
  vec3f  c  = {it.x, it.y, it.z};
  box_add(s.box + 6 * (uniform ? seg0 : seg), seg >= 0, uniform, c, c);
}
// split_middle, :564-595, up to the partition
__global__ void __launch_bounds__(256) k_tie_seg_decide(tie_segments s, int num) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= num) return;
  vec3f cmin = {s.box[6 * k], s.box[6 * k + 1], s.box[6 * k + 2]}, cmax = {s.box[6 * k + 3], s.box[6 * k + 4], s.box[6 * k + 5]};
  vec3f csize = cmax - cmin;
  int   axis = 0, nopart = 0;
  if (csize == vec3f{0, 0, 0}) {
    nopart = 1;
  } else {
    if (csize.x >= csize.y && csize.x >= csize.z) axis = 0;
    if (csize.y >= csize.x && csize.y >= csize.z) axis = 1;
    if (csize.z >= csize.x && csize.z >= csize.y) axis = 2;
  }
  vec3f cc = (cmin + cmax) / 2.0f;
  s.axis[k] = axis, s.middle[k] = axis == 0 ? cc.x : (axis == 1 ? cc.y : cc.z), s.nopart[k] = nopart;
}
__device__ __forceinline__ bool tie_pred(const float4& it, int axis, float middle) {
  return (axis == 0 ? it.x : (axis == 1 ? it.y : it.z)) < middle;
}
__global__ void __launch_bounds__(256) k_tie_count(const float4* items, const int* seg_of, int n, tie_segments s) {
  int i   = blockIdx.x * blockDim.x + threadIdx.x;
  int seg = i < n ? seg_of[i] : -1;
  bool yes = seg >= 0 && !s.nopart[seg] && tie_pred(items[i], s.axis[seg], s.middle[seg]);
  // one atomic per (warp, segment)
  unsigned peers = __match_any_sync(0xffffffffu, seg);
  unsigned votes = __ballot_sync(0xffffffffu, yes) & peers;
  if (seg >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1 && votes) atomicAdd(s.count + seg, __popc(votes));
}
__global__ void __launch_bounds__(256) k_tie_seg_mid(tie_segments s, int num) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= num) return;
  int start = s.start[k], end = s.end[k], mid = start + s.count[k];
  if (s.nopart[k] || mid == start || mid == end) mid = (start + end) / 2, s.nopart[k] = 1;  // :592, nothing was swapped
  s.mid[k] = mid;
}
// misplaced elements: left of mid with the predicate false, right of it with the predicate true
__global__ void __launch_bounds__(256) k_tie_misplaced(const float4* items, const int* seg_of, int n, tie_segments s, int* m_left, int* m_right) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int l = 0, r = 0;
  if (i < n) {
    int seg = seg_of[i];
    if (seg >= 0 && !s.nopart[seg]) {
      bool yes = tie_pred(items[i], s.axis[seg], s.middle[seg]);
      l = i < s.mid[seg] && !yes, r = i >= s.mid[seg] && yes;
    }
  }
  m_left[i] = l, m_right[i] = r;
}
__global__ void __launch_bounds__(256) k_tie_positions(const int* m_left, const int* m_right, const int* scan_left, const int* scan_right, int n,
    int* pos_left, int* pos_right) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (m_left[i]) pos_left[scan_left[i]] = i;
  if (m_right[i]) pos_right[scan_right[i]] = i;
}
// libstdc++'s std::partition on bidirectional iterators: the j-th misplaced element from the left changes places with
// the j-th misplaced element from the right
__global__ void __launch_bounds__(256) k_tie_swap(float4* items, const int* seg_of, const int* m_left, const int* scan_left, const int* scan_right,
    const int* pos_right, int n, tie_segments s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !m_left[i]) return;
  int seg = seg_of[i], start = s.start[seg], mid = s.mid[seg];
  int j = scan_left[i] - scan_left[start], pairs = scan_left[mid] - scan_left[start];
  int partner = pos_right[scan_right[start] + pairs - 1 - j];
  float4 a = items[i], b = items[partner];
  items[i] = b, items[partner] = a;
}
__global__ void __launch_bounds__(256) k_tie_children(tie_segments s, int num) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > num) return;
  s.nchild[k] = k < num ? (s.mid[k] - s.start[k] > kRefLeafMax) + (s.end[k] - s.mid[k] > kRefLeafMax) : 0;
}
__global__ void __launch_bounds__(256) k_tie_emit(tie_segments s, int num, const int* child_off, int node_base, tie_segments next, int4* nodes) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= num) return;
  int start = s.start[k], mid = s.mid[k], end = s.end[k];
  int c0 = mid - start > kRefLeafMax, c1 = end - mid > kRefLeafMax, off = child_off[k];
  if (c0) next.start[off] = start, next.end[off] = mid, next.node[off] = node_base + off;
  if (c1) next.start[off + c0] = mid, next.end[off + c0] = end, next.node[off + c0] = node_base + off + c0;
  nodes[s.node[k]] = make_int4(mid, s.axis[k], c0 ? node_base + off : -1, c1 ? node_base + off + c0 : -1);
}
__global__ void __launch_bounds__(256) k_tie_assign(int* seg_of, int n, tie_segments s, const int* child_off) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int seg = seg_of[i];
  if (seg < 0) return;
  int start = s.start[seg], mid = s.mid[seg], end = s.end[seg];
  int c0 = mid - start > kRefLeafMax, c1 = end - mid > kRefLeafMax, off = child_off[seg];
  seg_of[i] = i < mid ? (c0 ? off : -1) : (c1 ? off + c0 : -1);
}
__global__ void __launch_bounds__(256) k_tie_ref_pos(const float4* items, int n, int* ref_pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ref_pos[__float_as_int(items[i].w)] = i;
}

yh_status build_tie_order(host_shape& sh) {
  const int n = sh.num_elems();
  sh.ref_root_bbox = {{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
  sh.d_ref_pos.release(), sh.d_tie_nodes.release();
  if (n == 0) return YH_OK;
  const int B = 256;
  auto      G = [&](int c) { return (c + B - 1) / B; };
  const int max_seg = n / (kRefLeafMax + 1) + 2;  // segments of one level are disjoint ranges of > 4 elements

  dev_buffer<float4> items;
  dev_buffer<int>    seg_of, m_left, m_right, scan_left, scan_right, pos_left, pos_right, ibuf, child_off;
  dev_buffer<float>  fbuf, root_box;
  dev_buffer<int4>   nodes;
  dev_buffer<char>   tmp;
  YH_CUDA(items.alloc(n));
  YH_CUDA(seg_of.alloc(n));
  YH_CUDA(root_box.alloc(6));
  {
    float init[6] = {flt_max, flt_max, flt_max, -flt_max, -flt_max, -flt_max};
    YH_CUDA(cudaMemcpy(root_box.ptr, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  k_tie_items<<<G(n), B>>>(sh.kind(), sh.d_elems.ptr, sh.d_positions.ptr, sh.d_radius.ptr, n, items.ptr, seg_of.ptr, root_box.ptr);
  YH_CUDA(cudaGetLastError());
  {
    float box[6];
    YH_CUDA(cudaMemcpy(box, root_box.ptr, sizeof(box), cudaMemcpyDeviceToHost));
    sh.ref_root_bbox = {{box[0], box[1], box[2]}, {box[3], box[4], box[5]}};
  }
  YH_CUDA(sh.d_ref_pos.alloc(n));
  int num_nodes = 0;
  if (n > kRefLeafMax) {
    YH_CUDA(m_left.alloc((size_t)n + 1));
    YH_CUDA(m_right.alloc((size_t)n + 1));
    YH_CUDA(scan_left.alloc((size_t)n + 1));
    YH_CUDA(scan_right.alloc((size_t)n + 1));
    YH_CUDA(pos_left.alloc(n));
    YH_CUDA(pos_right.alloc(n));
    YH_CUDA(nodes.alloc((size_t)n));  // every node has > 4 elements and the ranges of a level are disjoint
    // two sets of per-segment arrays (this level / the next): 9 int arrays + child_off, 8 floats
    YH_CUDA(ibuf.alloc((size_t)2 * 9 * max_seg));
    YH_CUDA(fbuf.alloc((size_t)2 * 7 * max_seg));
    YH_CUDA(child_off.alloc((size_t)max_seg + 1));
    auto segs = [&](int which) {
      int*   ip = ibuf.ptr + (size_t)which * 9 * max_seg;
      float* fp = fbuf.ptr + (size_t)which * 7 * max_seg;
      return tie_segments{ip, ip + max_seg, ip + 2 * max_seg, fp, ip + 3 * max_seg, fp + 6 * max_seg, ip + 4 * max_seg, ip + 5 * max_seg,
          ip + 6 * max_seg, ip + 7 * max_seg};
    };
    size_t scan_bytes = 0, scan_bytes2 = 0;
    YH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, m_left.ptr, scan_left.ptr, n + 1));
    YH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes2, child_off.ptr, child_off.ptr, max_seg + 1));
    YH_CUDA(tmp.alloc(std::max(scan_bytes, scan_bytes2)));
    scan_bytes = tmp.count;
    int cur = 0, num = 1;
    {
      int first[3] = {0, n, 0};  // the root segment: [0, n), node 0
      tie_segments s0 = segs(0);
      YH_CUDA(cudaMemcpy(s0.start, &first[0], 4, cudaMemcpyHostToDevice));
      YH_CUDA(cudaMemcpy(s0.end, &first[1], 4, cudaMemcpyHostToDevice));
      YH_CUDA(cudaMemcpy(s0.node, &first[2], 4, cudaMemcpyHostToDevice));
    }
    num_nodes = 1;
    for (int level = 0; num > 0; level++) {
      if (level > 4096) return fail(YH_ERR_STATE, "tie order: the reference tree of this shape is deeper than 4096 levels");
      tie_segments s = segs(cur), nx = segs(cur ^ 1);
      k_tie_seg_reset<<<G(num), B>>>(s, num);
      k_tie_seg_bounds<<<G(n), B>>>(items.ptr, seg_of.ptr, n, s);
      k_tie_seg_decide<<<G(num), B>>>(s, num);
      k_tie_count<<<G(n), B>>>(items.ptr, seg_of.ptr, n, s);
      k_tie_seg_mid<<<G(num), B>>>(s, num);
      k_tie_misplaced<<<G(n + 1), B>>>(items.ptr, seg_of.ptr, n, s, m_left.ptr, m_right.ptr);
      YH_CUDA(cub::DeviceScan::ExclusiveSum(tmp.ptr, scan_bytes, m_left.ptr, scan_left.ptr, n + 1));
      YH_CUDA(cub::DeviceScan::ExclusiveSum(tmp.ptr, scan_bytes, m_right.ptr, scan_right.ptr, n + 1));
      k_tie_positions<<<G(n), B>>>(m_left.ptr, m_right.ptr, scan_left.ptr, scan_right.ptr, n, pos_left.ptr, pos_right.ptr);
      k_tie_swap<<<G(n), B>>>(items.ptr, seg_of.ptr, m_left.ptr, scan_left.ptr, scan_right.ptr, pos_right.ptr, n, s);
      k_tie_children<<<G(num + 1), B>>>(s, num);
      YH_CUDA(cub::DeviceScan::ExclusiveSum(tmp.ptr, scan_bytes, s.nchild, child_off.ptr, num + 1));
      k_tie_emit<<<G(num), B>>>(s, num, child_off.ptr, num_nodes, nx, nodes.ptr);
      k_tie_assign<<<G(n), B>>>(seg_of.ptr, n, s, child_off.ptr);
      YH_CUDA(cudaGetLastError());
      int opened = 0;
      YH_CUDA(cudaMemcpy(&opened, child_off.ptr + num, 4, cudaMemcpyDeviceToHost));
      num_nodes += opened;
      num = opened, cur ^= 1;
    }
  }
  k_tie_ref_pos<<<G(n), B>>>(items.ptr, n, sh.d_ref_pos.ptr);
  YH_CUDA(cudaGetLastError());
  if (num_nodes > 0) {  // keep only the nodes that exist
    YH_CUDA(sh.d_tie_nodes.alloc(num_nodes));
    YH_CUDA(cudaMemcpy(sh.d_tie_nodes.ptr, nodes.ptr, sizeof(int4) * (size_t)num_nodes, cudaMemcpyDeviceToDevice));
  }
  YH_CUDA(cudaDeviceSynchronize());
  return YH_OK;
}

}  // namespace yh

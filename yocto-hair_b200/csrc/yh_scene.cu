// Scene upload, GPU LBVH build over line / triangle primitives, light tables.
// Replaces the bodies of init_bvh (yocto_pathtrace.cpp:652-818) and init_lights
// (:1695-1740) behind the C ABI.
#include <cub/cub.cuh>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstring>
#include <future>

#include "yh_scene.cuh"

namespace yh {

// ---------------------------------------------------------------- kernels --

// One reference = one (primitive, sub-span) pair with its padded bounds.
struct ref_bounds {
  float4 lo;  // min.xyz, -
  float4 hi;  // max.xyz, -
};

__device__ __forceinline__ vec3f ldv(const float* p, int i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

// Number of references a primitive gets: long thin segments are bounded by up
// to max_split boxes along their length (aspect target: span <= 8 diameters).
__device__ __forceinline__ int split_count(const vec3f& p0, const vec3f& p1, float r0, float r1, int max_split) {
  if (max_split <= 1) return 1;
  float len = length(p1 - p0);
  float dia = fmaxf(r0 + r1, 1e-12f);  // mean diameter
  int   k   = (int)ceilf(len / (8.0f * dia));
  return k < 1 ? 1 : (k > max_split ? max_split : k);
}

// kind: 0 triangles, 1 lines, 2 points (host_shape::kind())
__global__ void k_count_refs(int kind, const int* elems, const float* pos, const float* radius, int n,
    int max_split, int* counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int k = 1;
  if (kind == 1) {
    int a = elems[2 * i], b = elems[2 * i + 1];
    k     = split_count(ldv(pos, a), ldv(pos, b), radius[a], radius[b], max_split);
  }
  counts[i] = k;
}

// Padding keeps every box conservative with respect to the rounding of the
// exact primitive test (the hit point of an accepted intersection always lies
// inside the padded box of the sub-span that contains its segment parameter).
// `scale` = largest |coordinate| of the shape: the traversal's fused slab test
// needs the pad to dominate ~1e-6 * scale (rounding of -o/d for |o| <= 4 scale, of
// the grid decode, and of the quantisation itself; yh_traverse.cuh).
__device__ __forceinline__ void pad_box(vec3f& lo, vec3f& hi, float scale) {
  float eps = 2e-6f * scale + 1e-30f;
  lo        = lo - eps;
  hi        = hi + eps;
}

__global__ void k_make_refs(int kind, const int* elems, const float* pos, const float* radius, int n,
    const int* offsets, const int* counts, ref_bounds* refs, int* ref_prim, vec3f cmin, vec3f cinv, float scale,
    unsigned long long* keys, int* ref_ids) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int first = offsets[i], k = counts[i];
  if (kind == 2) {
    int    a = elems[i];
    bbox3f box = point_bounds(ldv(pos, a), radius[a]);
    pad_box(box.min, box.max, scale);
    refs[first]     = {make_float4(box.min.x, box.min.y, box.min.z, 0), make_float4(box.max.x, box.max.y, box.max.z, 0)};
    ref_prim[first] = i;
  } else if (kind == 1) {
    int   a = elems[2 * i], b = elems[2 * i + 1];
    vec3f p0 = ldv(pos, a), p1 = ldv(pos, b);
    float r0 = radius[a], r1 = radius[b];
    for (int s = 0; s < k; s++) {
      float s0 = (float)s / (float)k, s1 = (float)(s + 1) / (float)k;
      vec3f pa = p0 + (p1 - p0) * s0, pb = p0 + (p1 - p0) * s1;
      float ra = r0 * (1 - s0) + r1 * s0, rb = r0 * (1 - s1) + r1 * s1;
      // the interpolated radius can exceed both ends' by rounding only; widen a hair
      ra         = ra * 1.000001f;
      rb         = rb * 1.000001f;
      bbox3f box = k == 1 ? line_bounds(p0, p1, r0, r1) : line_bounds(pa, pb, ra, rb);
      pad_box(box.min, box.max, scale);
      refs[first + s]     = {make_float4(box.min.x, box.min.y, box.min.z, 0), make_float4(box.max.x, box.max.y, box.max.z, 0)};
      ref_prim[first + s] = i;
    }
  } else {
    int    a = elems[3 * i], b = elems[3 * i + 1], c = elems[3 * i + 2];
    bbox3f box = triangle_bounds(ldv(pos, a), ldv(pos, b), ldv(pos, c));
    pad_box(box.min, box.max, scale);
    refs[first]     = {make_float4(box.min.x, box.min.y, box.min.z, 0), make_float4(box.max.x, box.max.y, box.max.z, 0)};
    ref_prim[first] = i;
  }
  // 63-bit Morton key of each reference's box centre
  for (int s = 0; s < k; s++) {
    ref_bounds r = refs[first + s];
    float cx = ((r.lo.x + r.hi.x) * 0.5f - cmin.x) * cinv.x;
    float cy = ((r.lo.y + r.hi.y) * 0.5f - cmin.y) * cinv.y;
    float cz = ((r.lo.z + r.hi.z) * 0.5f - cmin.z) * cinv.z;
    auto  q  = [](float v) -> unsigned long long {
      v = fminf(fmaxf(v * 2097152.0f, 0.0f), 2097151.0f);
      unsigned long long x = (unsigned long long)v;  // 21 bits
      x = (x | x << 32) & 0x1f00000000ffffull;
      x = (x | x << 16) & 0x1f0000ff0000ffull;
      x = (x | x << 8) & 0x100f00f00f00f00full;
      x = (x | x << 4) & 0x10c30c30c30c30c3ull;
      x = (x | x << 2) & 0x1249249249249249ull;
      return x;
    };
    keys[first + s]    = (q(cx) << 2) | (q(cy) << 1) | q(cz);
    ref_ids[first + s] = first + s;
  }
}

// Karras 2012, "Maximizing parallelism in the construction of BVHs": one thread
// per internal node of the binary radix tree over the sorted keys.
__device__ __forceinline__ int lcp(const unsigned long long* keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}
// child encoding inside the build: >= 0 internal node, < 0 leaf ~position
__global__ void k_hierarchy(const unsigned long long* keys, int n, int* left, int* right, int* first,
    int* last, int* parent_inner, int* parent_leaf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d    = (lcp(keys, n, i, i + 1) - lcp(keys, n, i, i - 1)) < 0 ? -1 : 1;
  int dmin = lcp(keys, n, i, i - d);
  int lmax = 2;
  while (lcp(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (lcp(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j     = i + l * d;
  int dnode = lcp(keys, n, i, j);
  int s     = 0;
  int t     = l;
  do {
    t = (t + 1) >> 1;
    if (lcp(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  int gamma = i + s * d + min(d, 0);
  int lo = min(i, j), hi = max(i, j);
  first[i] = lo;
  last[i]  = hi;
  if (lo == gamma) {
    left[i]            = ~gamma;
    parent_leaf[gamma] = i;
  } else {
    left[i]             = gamma;
    parent_inner[gamma] = i;
  }
  if (hi == gamma + 1) {
    right[i]               = ~(gamma + 1);
    parent_leaf[gamma + 1] = i;
  } else {
    right[i]                = gamma + 1;
    parent_inner[gamma + 1] = i;
  }
  if (i == 0) parent_inner[0] = -1;
}

// Bottom-up bounds: the second thread to reach a node merges its children.
__global__ void k_refit(int n, const int* sorted_ref, const ref_bounds* refs, const int* left,
    const int* right, const int* parent_inner, const int* parent_leaf, int* flags, ref_bounds* node_box) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int node = parent_leaf[j];
  while (node >= 0) {
    __threadfence();
    if (atomicAdd(&flags[node], 1) == 0) return;
    __threadfence();
    int        l = left[node], r = right[node];
    // child boxes were written by other SMs: read them past L1 (.cg), a line of
    // node_box may sit there from before its neighbour was written
    auto       ld = [&](int c) -> ref_bounds {
      if (c < 0) return refs[sorted_ref[~c]];
      return {__ldcg(&node_box[c].lo), __ldcg(&node_box[c].hi)};
    };
    ref_bounds a = ld(l), b = ld(r);
    ref_bounds m;
    m.lo = make_float4(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z), 0);
    m.hi = make_float4(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z), 0);
    node_box[node] = m;
    node           = parent_inner[node];
  }
}

// Tree rotations (Kensler 2008) over the radix tree, bottom-up like k_refit: when both subtrees of a node are
// final, the thread that arrived second tries the four child <-> grandchild swaps and keeps the one that shrinks the
// surface area of the affected inner child the most (the node's own box does not change).  Everything a rotation
// touches lies inside the node's subtree, which no other thread works on any more.
__device__ __forceinline__ float box_area(const ref_bounds& b) {
  float dx = b.hi.x - b.lo.x, dy = b.hi.y - b.lo.y, dz = b.hi.z - b.lo.z;
  return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ ref_bounds box_merge(const ref_bounds& a, const ref_bounds& b) {
  ref_bounds m;
  m.lo = make_float4(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z), 0);
  m.hi = make_float4(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z), 0);
  return m;
}
__global__ void k_rotate(int n, const int* sorted_ref, const ref_bounds* refs, int* left, int* right, int* parent_inner,
    int* parent_leaf, int* flags, ref_bounds* node_box, int* rotations) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int  node = __ldcg(parent_leaf + j);
  auto box  = [&](int c) -> ref_bounds {
    if (c < 0) return refs[sorted_ref[~c]];
    return {__ldcg(&node_box[c].lo), __ldcg(&node_box[c].hi)};
  };
  auto set_parent = [&](int c, int p) {
    if (c < 0) parent_leaf[~c] = p;
    else parent_inner[c] = p;
  };
  while (node >= 0) {
    __threadfence();
    if (atomicAdd(&flags[node], 1) == 0) return;
    __threadfence();
    int   child[2] = {__ldcg(left + node), __ldcg(right + node)};
    float best_gain = 0;
    int   best_side = -1, best_opt = 0;
    ref_bounds best_box{};
    for (int side = 0; side < 2; side++) {
      int A = child[side], B = child[side ^ 1];  // A moves down into B, a child of B moves up
      if (B < 0) continue;
      int        b0 = __ldcg(left + B), b1 = __ldcg(right + B);
      ref_bounds boxA = box(A);
      float      areaB = box_area(box(B));
      ref_bounds m0 = box_merge(boxA, box(b1));  // swap A with b0: B' = {A, b1}
      ref_bounds m1 = box_merge(box(b0), boxA);  // swap A with b1: B' = {b0, A}
      float g0 = areaB - box_area(m0), g1 = areaB - box_area(m1);
      if (g0 > best_gain) best_gain = g0, best_side = side, best_opt = 0, best_box = m0;
      if (g1 > best_gain) best_gain = g1, best_side = side, best_opt = 1, best_box = m1;
    }
    // both children inner: the two other pairings of the four grandchildren {a, b | c, d} -> {a, c | b, d}, {a, d | b, c}
    if (child[0] >= 0 && child[1] >= 0) {
      int L = child[0], Rn = child[1];
      int a = __ldcg(left + L), b = __ldcg(right + L), c = __ldcg(left + Rn), d = __ldcg(right + Rn);
      ref_bounds ba = box(a), bb = box(b), bc = box(c), bd = box(d);
      float      cost0 = box_area(box(L)) + box_area(box(Rn));
      ref_bounds ac = box_merge(ba, bc), bdm = box_merge(bb, bd), ad = box_merge(ba, bd), bcm = box_merge(bb, bc);
      float      g1 = cost0 - (box_area(ac) + box_area(bdm)), g2 = cost0 - (box_area(ad) + box_area(bcm));
      if (g1 > best_gain && g1 >= g2) {
        right[L] = c, left[Rn] = b;
        set_parent(c, L), set_parent(b, Rn);
        node_box[L] = ac, node_box[Rn] = bdm;
        atomicAdd(rotations, 1);
        best_side = -1;
      } else if (g2 > best_gain) {
        right[L] = d, right[Rn] = b;  // L' = {a, d}, R' = {c, b}
        set_parent(d, L), set_parent(b, Rn);
        node_box[L] = ad, node_box[Rn] = bcm;
        atomicAdd(rotations, 1);
        best_side = -1;
      }
    }
    if (best_side >= 0) {
      int A = child[best_side], B = child[best_side ^ 1];
      int b0 = __ldcg(left + B), b1 = __ldcg(right + B);
      int up = best_opt == 0 ? b0 : b1;  // the grandchild that takes A's place under `node`
      if (best_side == 0) left[node] = up;
      else right[node] = up;
      if (best_opt == 0) left[B] = A;
      else right[B] = A;
      set_parent(up, node), set_parent(A, B);
      node_box[B] = best_box;
      atomicAdd(rotations, 1);
    }
    node = __ldcg(parent_inner + node);
  }
}

// Depth of the deepest leaf (inner nodes above it): the traversal stack holds one entry per level.
__global__ void k_max_depth(int n, const int* parent_inner, const int* parent_leaf, int* max_depth) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int depth = 0;
  for (int node = parent_leaf[j]; node >= 0; node = parent_inner[node]) depth++;
  atomicMax(max_depth, depth);
}

// Quantises one box onto the shape's 16-bit grid, outwards, after padding it by `pad` grid steps
// per axis (fold_pad_steps(): the traversal's plane decode moves all planes of an axis together
// by up to about half a step, yh_traverse.cuh slab_consts).
__device__ __forceinline__ void quantise(const ref_bounds& b, const vec3f& g, const vec3f& inv_s, const vec3f& pad,
    unsigned q[6]) {
  auto dn = [](float v) { return (unsigned)fminf(fmaxf(floorf(v), 0.0f), 65535.0f); };
  auto up = [](float v) { return (unsigned)fminf(fmaxf(ceilf(v), 0.0f), 65535.0f); };
  q[0] = dn((b.lo.x - g.x) * inv_s.x - pad.x), q[1] = dn((b.lo.y - g.y) * inv_s.y - pad.y), q[2] = dn((b.lo.z - g.z) * inv_s.z - pad.z);
  q[3] = up((b.hi.x - g.x) * inv_s.x + pad.x), q[4] = up((b.hi.y - g.y) * inv_s.y + pad.y), q[5] = up((b.hi.z - g.z) * inv_s.z + pad.z);
}

// Writes the packed tree (yh_scene.cuh "shape BVH"): one thread per inner node
// of the radix tree emits its own 32-byte record at the slot its parent
// reserved for it, and the primitive records of its leaf children next to it.
__global__ void k_emit_packed(int kind, int n_inner, const int* left, const int* right, const int* parent_inner,
    const int* sorted_ref, const ref_bounds* refs, const ref_bounds* node_box, const int* ref_prim, const int* elems,
    const float* pos, const float* radius, vec3f g, vec3f inv_s, vec3f pad, uint4* slots, int* slot_elem) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inner) return;
  int my_slot = 0;
  if (i > 0) {
    int p   = parent_inner[i];
    my_slot = 1 + 2 * p + (right[p] == i ? 1 : 0);
  }
  int      child[2] = {left[i], right[i]};
  unsigned q[2][6];
  unsigned flags = 0;
  for (int c = 0; c < 2; c++) {
    int ch = child[c];
    if (ch < 0) {
      int sr = sorted_ref[~ch];
      quantise(refs[sr], g, inv_s, pad, q[c]);
      flags |= 0x80000000u >> c;
      // the leaf child is the primitive itself
      int    prim = ref_prim[sr];
      size_t slot = (size_t)(1 + 2 * i + c);
      if (kind == 1) {
        int a = elems[2 * prim], b = elems[2 * prim + 1];
        slots[2 * slot]     = make_uint4(__float_as_uint(pos[3 * a]), __float_as_uint(pos[3 * a + 1]), __float_as_uint(pos[3 * a + 2]), __float_as_uint(radius[a]));
        slots[2 * slot + 1] = make_uint4(__float_as_uint(pos[3 * b]), __float_as_uint(pos[3 * b + 1]), __float_as_uint(pos[3 * b + 2]), __float_as_uint(radius[b]));
      } else if (kind == 2) {
        int a = elems[prim];
        slots[2 * slot]     = make_uint4(__float_as_uint(pos[3 * a]), __float_as_uint(pos[3 * a + 1]), __float_as_uint(pos[3 * a + 2]), __float_as_uint(radius[a]));
        slots[2 * slot + 1] = make_uint4(0, 0, 0, 0);
      } else {
        slots[2 * slot]     = make_uint4((unsigned)prim, 0, 0, 0);
        slots[2 * slot + 1] = make_uint4(0, 0, 0, 0);
      }
      slot_elem[slot] = prim;
    } else {
      quantise(node_box[ch], g, inv_s, pad, q[c]);
    }
  }
  // one word per axis: lo | hi << 16
  slots[2 * (size_t)my_slot] = make_uint4(q[0][0] | (q[0][3] << 16), q[0][1] | (q[0][4] << 16), q[0][2] | (q[0][5] << 16),
      q[1][0] | (q[1][3] << 16));
  slots[2 * (size_t)my_slot + 1] = make_uint4(q[1][1] | (q[1][4] << 16), q[1][2] | (q[1][5] << 16), (unsigned)(1 + 2 * i), flags);
}

// Triangle vertices by element (48 bytes): leaf slots of triangle shapes hold the element id.
__global__ void k_emit_triangles(int n, const int* elems, const float* pos, float4* tris) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int   a = elems[3 * j], b = elems[3 * j + 1], c = elems[3 * j + 2];
  vec3f p0 = ldv(pos, a), p1 = ldv(pos, b), p2 = ldv(pos, c);
  tris[3 * (size_t)j]     = make_float4(p0.x, p0.y, p0.z, 0);
  tris[3 * (size_t)j + 1] = make_float4(p1.x, p1.y, p1.z, 0);
  tris[3 * (size_t)j + 2] = make_float4(p2.x, p2.y, p2.z, 0);
}

// ------------------------------------------------------------ host driver --

// One split of the reference's builder (build_bvh / split_middle, yocto_pathtrace.cpp:564-650): spatial-middle split
// of the centroid bounds along their largest axis with std::partition, median index when that fails.  Used on the
// host for the top-level tree over the OBJECTS (build_scene_tree); the per-shape trees, whose leaf order breaks
// exact-t ties, are grown on the device (yh_tieorder.cu).
struct ref_prim {
  vec3f center;
  int   primitive;
};
// one split of split_middle (yocto_pathtrace.cpp:564-595) on prims[start, end)
static void reference_split(ref_prim* prims, int start, int end, int& mid, int& axis) {
  axis = 0, mid = (start + end) / 2;
  vec3f cmin = {flt_max, flt_max, flt_max}, cmax = {-flt_max, -flt_max, -flt_max};
  for (int i = start; i < end; i++) cmin = min3(cmin, prims[i].center), cmax = max3(cmax, prims[i].center);
  vec3f csize = cmax - cmin;
  if (csize == vec3f{0, 0, 0}) return;
  if (csize.x >= csize.y && csize.x >= csize.z) axis = 0;
  if (csize.y >= csize.x && csize.y >= csize.z) axis = 1;
  if (csize.z >= csize.x && csize.z >= csize.y) axis = 2;
  vec3f cc     = (cmin + cmax) / 2.0f;
  float middle = axis == 0 ? cc.x : (axis == 1 ? cc.y : cc.z);
  int   ax     = axis;
  auto  key    = [ax](const ref_prim& p) { return ax == 0 ? p.center.x : (ax == 1 ? p.center.y : p.center.z); };
  mid = (int)(std::partition(prims + start, prims + end, [&](const ref_prim& p) { return key(p) < middle; }) - prims);
  if (mid == start || mid == end) mid = (start + end) / 2;
}
// ---- quality builder (yh_bvh_params::quality = 1): binned SAH, top down, on the host ----------
// Same outputs as k_hierarchy (left / right / parents over inner nodes 0 .. R-2, root = 0, leaf
// children encoded as ~position) plus the permutation of the references, so the device kernels
// k_refit and k_emit_packed finish the build exactly as for the LBVH.  Inner nodes are numbered
// in preorder (a subtree over m references owns m - 1 consecutive ids), which lets independent
// subtrees be built by different host threads.
struct sah_topology {
  std::vector<int> order, left, right, parent_inner, parent_leaf;
};
struct sah_bin {
  vec3f lo{flt_max, flt_max, flt_max}, hi{-flt_max, -flt_max, -flt_max};
  int   count = 0;
  void  add(const ref_bounds& b) {
    lo = min3(lo, vec3f{b.lo.x, b.lo.y, b.lo.z}), hi = max3(hi, vec3f{b.hi.x, b.hi.y, b.hi.z});
    count++;
  }
  void merge(const sah_bin& o) { lo = min3(lo, o.lo), hi = max3(hi, o.hi), count += o.count; }
  float area() const {
    if (count == 0) return 0;
    vec3f d = hi - lo;
    return d.x * d.y + d.y * d.z + d.z * d.x;
  }
};
constexpr int kSahBins = 16;
static inline vec3f ref_centre(const ref_bounds& b) { return {(b.lo.x + b.hi.x) * 0.5f, (b.lo.y + b.hi.y) * 0.5f, (b.lo.z + b.hi.z) * 0.5f}; }
static inline float axis_of(const vec3f& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// best binned split of order[s, e): returns the partition point (order is permuted)
static int sah_split(const ref_bounds* refs, int* order, int s, int e) {
  int   m = e - s;
  vec3f cmin = {flt_max, flt_max, flt_max}, cmax = {-flt_max, -flt_max, -flt_max};
  auto  scan_bounds = [&](int a, int b, vec3f& lo, vec3f& hi) {
    for (int i = a; i < b; i++) {
      vec3f c = ref_centre(refs[order[i]]);
      lo = min3(lo, c), hi = max3(hi, c);
    }
  };
  const int nthreads = m > (1 << 18) ? 8 : 1;  // the top ranges are binned by several threads
  if (nthreads == 1) {
    scan_bounds(s, e, cmin, cmax);
  } else {
    std::vector<std::future<std::pair<vec3f, vec3f>>> fs;
    for (int t = 0; t < nthreads; t++)
      fs.push_back(std::async(std::launch::async, [&, t]() {
        vec3f lo = {flt_max, flt_max, flt_max}, hi = {-flt_max, -flt_max, -flt_max};
        scan_bounds(s + (long long)m * t / nthreads, s + (long long)m * (t + 1) / nthreads, lo, hi);
        return std::make_pair(lo, hi);
      }));
    for (auto& f : fs) {
      auto r = f.get();
      cmin = min3(cmin, r.first), cmax = max3(cmax, r.second);
    }
  }
  vec3f ext = cmax - cmin;
  if (!(ext.x > 0) && !(ext.y > 0) && !(ext.z > 0)) return (s + e) / 2;
  vec3f scale = {ext.x > 0 ? kSahBins / ext.x : 0.0f, ext.y > 0 ? kSahBins / ext.y : 0.0f, ext.z > 0 ? kSahBins / ext.z : 0.0f};
  auto  bin_of = [&](const vec3f& c, int a) {
    int b = (int)((axis_of(c, a) - axis_of(cmin, a)) * axis_of(scale, a));
    return b < 0 ? 0 : (b >= kSahBins ? kSahBins - 1 : b);
  };
  using bin_set = std::array<std::array<sah_bin, kSahBins>, 3>;
  auto fill = [&](int a, int b, bin_set& bins) {
    for (int i = a; i < b; i++) {
      const ref_bounds& r = refs[order[i]];
      vec3f c = ref_centre(r);
      for (int ax = 0; ax < 3; ax++) bins[ax][bin_of(c, ax)].add(r);
    }
  };
  bin_set bins;
  if (nthreads == 1) {
    fill(s, e, bins);
  } else {
    std::vector<std::future<bin_set>> fs;
    for (int t = 0; t < nthreads; t++)
      fs.push_back(std::async(std::launch::async, [&, t]() {
        bin_set local;
        fill(s + (long long)m * t / nthreads, s + (long long)m * (t + 1) / nthreads, local);
        return local;
      }));
    for (auto& f : fs) {
      bin_set local = f.get();
      for (int ax = 0; ax < 3; ax++)
        for (int b = 0; b < kSahBins; b++) bins[ax][b].merge(local[ax][b]);
    }
  }
  float best = flt_max;
  int   best_axis = -1, best_bin = 0;
  for (int ax = 0; ax < 3; ax++) {
    if (!(axis_of(ext, ax) > 0)) continue;
    float   right_cost[kSahBins];
    sah_bin acc;
    for (int b = kSahBins - 1; b > 0; b--) {
      acc.merge(bins[ax][b]);
      right_cost[b] = acc.area() * acc.count;
    }
    sah_bin accl;
    for (int b = 0; b + 1 < kSahBins; b++) {
      accl.merge(bins[ax][b]);
      if (accl.count == 0 || accl.count == m) continue;
      float cost = accl.area() * accl.count + right_cost[b + 1];
      if (cost < best) best = cost, best_axis = ax, best_bin = b;
    }
  }
  if (best_axis < 0) return (s + e) / 2;
  int mid = (int)(std::partition(order + s, order + e,
                      [&](int id) { return bin_of(ref_centre(refs[id]), best_axis) <= best_bin; }) -
                  order);
  if (mid == s || mid == e) mid = (s + e) / 2;
  return mid;
}
// Tree depth at which the builder stops looking for the best split and halves the range instead: the
// traversal stack holds one entry per level (kStackShared + kStackLocal = 96, yh_traverse.cuh), and inputs
// such as geometrically spaced primitives make every SAH split peel off ONE reference.  Below this depth
// the subtree is balanced, so the total depth stays under kSahMaxDepth + log2(references) < 96.
constexpr int kSahMaxDepth = 64;
static void sah_build_range(const ref_bounds* refs, sah_topology& t, int node, int s, int e, int depth, int tree_depth) {
  struct item { int node, s, e, level; };
  std::vector<item>              todo{{node, s, e, tree_depth}};
  std::vector<std::future<void>> spawned;
  while (!todo.empty()) {
    item it = todo.back();
    todo.pop_back();
    int mid = it.level >= kSahMaxDepth ? (it.s + it.e) / 2 : sah_split(refs, t.order.data(), it.s, it.e);
    int ml = mid - it.s, mr = it.e - mid;
    int child[2] = {ml == 1 ? ~it.s : it.node + 1, mr == 1 ? ~mid : it.node + ml};
    t.left[it.node] = child[0], t.right[it.node] = child[1];
    int range[2][2] = {{it.s, mid}, {mid, it.e}};
    for (int c = 0; c < 2; c++) {
      if (child[c] < 0) {
        t.parent_leaf[~child[c]] = it.node;
        continue;
      }
      t.parent_inner[child[c]] = it.node;
      int cnt = range[c][1] - range[c][0];
      if (c == 0 && depth < 6 && cnt > (1 << 14) && (it.e - it.s) - cnt > (1 << 14)) {  // big sibling pair: fork
        int cn = child[c], cs = range[c][0], ce = range[c][1], cl = it.level + 1;
        spawned.push_back(std::async(std::launch::async, [refs, &t, cn, cs, ce, depth, cl]() { sah_build_range(refs, t, cn, cs, ce, depth + 1, cl); }));
      } else {
        todo.push_back({child[c], range[c][0], range[c][1], it.level + 1});
      }
    }
    if (!spawned.empty() && depth < 6) depth++;  // forks deepen the task tree, not the loop
  }
  for (auto& f : spawned) f.get();
}
static void sah_build(const std::vector<ref_bounds>& refs, sah_topology& t) {
  int R = (int)refs.size();
  t.order.resize(R);
  for (int i = 0; i < R; i++) t.order[i] = i;
  t.left.assign(R - 1, 0), t.right.assign(R - 1, 0), t.parent_inner.assign(R - 1, -1), t.parent_leaf.assign(R, -1);
  sah_build_range(refs.data(), t, 0, 0, R, 0, 0);
}

static yh_status build_shape_bvh(host_shape& sh, int max_split, int quality) {
  bool lines = sh.is_lines();
  int  kind  = sh.kind();
  int  n     = sh.num_elems();
  sh.root_leaf = 0, sh.num_refs = 0, sh.num_nodes = 0, sh.tiny_count = 0;
  sh.root_bbox = {{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
  if (n == 0) return YH_OK;
  // YH_BUILD_TIMING=1: phase times of this build on stderr (host clock; each mark synchronises the device)
  const bool timing = getenv("YH_BUILD_TIMING") != nullptr;
  auto       t_last = std::chrono::steady_clock::now();
  auto       mark   = [&](const char* what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "  bvh %-28s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  if (!lines) max_split = 1;
  const int B = 256;
  auto      G = [&](int c) { return (c + B - 1) / B; };

  // references
  dev_buffer<int> counts, offsets;
  YH_CUDA(counts.alloc(n));
  YH_CUDA(offsets.alloc(n));
  k_count_refs<<<G(n), B>>>(kind, sh.d_elems.ptr, sh.d_positions.ptr, sh.d_radius.ptr, n, max_split, counts.ptr);
  YH_CUDA(cudaGetLastError());
  size_t            tmp_bytes = 0;
  dev_buffer<char>  tmp;
  YH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.ptr, offsets.ptr, n));
  YH_CUDA(tmp.alloc(tmp_bytes));
  YH_CUDA(cub::DeviceScan::ExclusiveSum(tmp.ptr, tmp_bytes, counts.ptr, offsets.ptr, n));
  int last_off = 0, last_cnt = 0;
  YH_CUDA(cudaMemcpy(&last_off, offsets.ptr + (n - 1), 4, cudaMemcpyDeviceToHost));
  YH_CUDA(cudaMemcpy(&last_cnt, counts.ptr + (n - 1), 4, cudaMemcpyDeviceToHost));
  int R = last_off + last_cnt;
  if (R >= (1 << 28)) return fail(YH_ERR_UNSUPPORTED, "shape has %d BVH references (limit 2^28)", R);

  mark("count + scan references");
  // centre bounds for Morton quantisation (host pass over the vertices)
  vec3f cmin = {flt_max, flt_max, flt_max}, cmax = {-flt_max, -flt_max, -flt_max};
  float rmax = 0;
  {
    const size_t nv = sh.positions.size() / 3;
    const int    nt = nv > (1u << 16) ? 8 : 1;
    struct part { vec3f lo, hi; float r; };
    std::vector<std::future<part>> parts;
    for (int t = 0; t < nt; t++)
      parts.push_back(std::async(nt > 1 ? std::launch::async : std::launch::deferred, [&sh, nv, nt, t]() {
        part q{{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}, 0};
        for (size_t v = nv * t / nt; v < nv * (t + 1) / nt; v++) {
          vec3f p = {sh.positions[3 * v], sh.positions[3 * v + 1], sh.positions[3 * v + 2]};
          q.lo = min3(q.lo, p), q.hi = max3(q.hi, p);
        }
        for (size_t v = sh.radius.size() * t / nt; v < sh.radius.size() * (t + 1) / nt; v++) q.r = fmax_(q.r, fabs_(sh.radius[v]));
        return q;
      }));
    for (auto& f : parts) {
      part q = f.get();
      cmin = min3(cmin, q.lo), cmax = max3(cmax, q.hi), rmax = fmax_(rmax, q.r);
    }
  }
  sh.scale = fmax_(fmax_(fmax_(fabs_(cmin.x), fabs_(cmax.x)), fmax_(fabs_(cmin.y), fabs_(cmax.y))),
                 fmax_(fabs_(cmin.z), fabs_(cmax.z))) + rmax;
  vec3f ext  = cmax - cmin;
  vec3f cinv = {ext.x > 0 ? 1.0f / ext.x : 0.0f, ext.y > 0 ? 1.0f / ext.y : 0.0f, ext.z > 0 ? 1.0f / ext.z : 0.0f};

  mark("host bounds of the vertices");
  dev_buffer<ref_bounds>         refs;
  dev_buffer<int>                ref_prim, ids_in, ids_out;
  dev_buffer<unsigned long long> keys_in, keys_out;
  YH_CUDA(refs.alloc(R));
  YH_CUDA(ref_prim.alloc(R));
  YH_CUDA(ids_in.alloc(R));
  YH_CUDA(ids_out.alloc(R));
  YH_CUDA(keys_in.alloc(R));
  YH_CUDA(keys_out.alloc(R));
  k_make_refs<<<G(n), B>>>(kind, sh.d_elems.ptr, sh.d_positions.ptr, sh.d_radius.ptr, n, offsets.ptr,
      counts.ptr, refs.ptr, ref_prim.ptr, cmin, cinv, sh.scale, keys_in.ptr, ids_in.ptr);
  YH_CUDA(cudaGetLastError());
  tmp_bytes = 0;
  YH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in.ptr, keys_out.ptr, ids_in.ptr, ids_out.ptr, R, 0, 63));
  YH_CUDA(tmp.alloc(tmp_bytes));
  YH_CUDA(cub::DeviceRadixSort::SortPairs(tmp.ptr, tmp_bytes, keys_in.ptr, keys_out.ptr, ids_in.ptr, ids_out.ptr, R, 0, 63));

  mark("references + Morton sort");
  // grid of the packed boxes: 16 bits per axis over the (padded) bounds of the shape
  float pad  = 2e-6f * sh.scale + 1e-30f;
  vec3f glo  = cmin - (rmax + pad), ghi = cmax + (rmax + pad);
  auto  step = [](float lo, float hi) { return fmax_((hi - lo) / 65534.0f, 1e-30f); };
  // The boxes are padded by fold_pad_steps() grid steps and rounded outwards AFTER that; a box that touches the
  // bounds of the shape must not lose that padding to the clamp at the ends of the grid (a ray grazing the outside
  // face of a shape within half a step - every ray, for a flat shape such as a quad light - would miss it), so the grid
  // extends past the padded bounds by the padding in space, 0.54 step + 0.53 (2 scale + extent) 2^-23, plus a step
  // for the rounding, with some slack.
  auto margin = [&](float lo, float hi) { return 2.0f * step(lo, hi) + 0.6f * (2.0f * sh.scale + (hi - lo)) / 8388608.0f; };
  vec3f gm = {margin(glo.x, ghi.x), margin(glo.y, ghi.y), margin(glo.z, ghi.z)};
  glo = glo - gm, ghi = ghi + gm;
  sh.grid_origin = glo;
  sh.grid_step   = {step(glo.x, ghi.x), step(glo.y, ghi.y), step(glo.z, ghi.z)};
  vec3f inv_step = {1.0f / sh.grid_step.x, 1.0f / sh.grid_step.y, 1.0f / sh.grid_step.z};
  // Padding in grid steps that covers the traversal's folded plane decode t = fma(2^23 + q, a, b - 2^23 a):
  // b - 2^23 a is rounded once, to half an ulp of a number of magnitude 2^23 |a| + |b|, where |b| <= (distance
  // of the box origin from the grid origin, at most 2 scale + extent) / step * |a|.  In steps: ulp <= 2 * 2^-24 *
  // (2^23 + |b| / |a|), doubled once more when the sum crosses a power of two => pad = 0.5 * (1 + |b/a| 2^-23) * 1.06.
  auto fold_pad_steps = [&](float lo, float hi, float inv) {
    float reach = (2.0f * sh.scale + (hi - lo)) * inv;  // |b / a| bound, in steps
    return 0.53f * (1.0f + reach / 8388608.0f) + 0.01f;
  };
  vec3f pad_steps = {fold_pad_steps(glo.x, ghi.x, inv_step.x), fold_pad_steps(glo.y, ghi.y, inv_step.y),
      fold_pad_steps(glo.z, ghi.z, inv_step.z)};
  sh.num_refs    = R;
  if (kind == 0) {
    YH_CUDA(sh.d_prims.alloc((size_t)n * 3));
    k_emit_triangles<<<G(n), B>>>(n, sh.d_elems.ptr, sh.d_positions.ptr, sh.d_prims.ptr);
    YH_CUDA(cudaGetLastError());
  }
  size_t num_slots = 2 * (size_t)R - 1;
  YH_CUDA(sh.d_slots.alloc(2 * num_slots));
  YH_CUDA(sh.d_slot_elem.alloc(num_slots));
  YH_CUDA(cudaMemset(sh.d_slot_elem.ptr, 0xff, sh.d_slot_elem.bytes()));
  sh.num_nodes = (int)num_slots;

  if (R == 1) {  // the whole shape is one leaf: slot 0 is the primitive
    std::vector<ref_bounds> h(1);
    YH_CUDA(cudaMemcpy(h.data(), refs.ptr, sizeof(ref_bounds), cudaMemcpyDeviceToHost));
    sh.root_bbox = {{h[0].lo.x, h[0].lo.y, h[0].lo.z}, {h[0].hi.x, h[0].hi.y, h[0].hi.z}};
    sh.root_leaf = 1;
    uint4 rec[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (lines) {
      int a = sh.lines[0], b = sh.lines[1];
      float v[8] = {sh.positions[3 * a], sh.positions[3 * a + 1], sh.positions[3 * a + 2], sh.radius[a],
          sh.positions[3 * b], sh.positions[3 * b + 1], sh.positions[3 * b + 2], sh.radius[b]};
      memcpy(rec, v, 32);
    } else if (kind == 2) {
      int a = sh.points[0];
      float v[8] = {sh.positions[3 * a], sh.positions[3 * a + 1], sh.positions[3 * a + 2], sh.radius[a], 0, 0, 0, 0};
      memcpy(rec, v, 32);
    }
    int zero = 0;
    YH_CUDA(cudaMemcpy(sh.d_slots.ptr, rec, 32, cudaMemcpyHostToDevice));
    YH_CUDA(cudaMemcpy(sh.d_slot_elem.ptr, &zero, 4, cudaMemcpyHostToDevice));
    YH_CUDA(cudaStreamSynchronize(0));
    return YH_OK;
  }
  sh.root_leaf = 0;

  // hierarchy + bounds
  int             ni = R - 1;
  dev_buffer<int> left, right, first, last, parent_inner, parent_leaf, flags;
  dev_buffer<ref_bounds> node_box;
  YH_CUDA(left.alloc(ni));
  YH_CUDA(right.alloc(ni));
  YH_CUDA(first.alloc(ni));
  YH_CUDA(last.alloc(ni));
  YH_CUDA(parent_inner.alloc(ni));
  YH_CUDA(parent_leaf.alloc(R));
  YH_CUDA(flags.alloc(ni));
  YH_CUDA(node_box.alloc(ni));
  YH_CUDA(cudaMemset(flags.ptr, 0, flags.bytes()));
  if (quality >= 1) {  // binned SAH topology from the host builder; ids_out becomes its permutation
    std::vector<ref_bounds> href(R);
    YH_CUDA(cudaMemcpy(href.data(), refs.ptr, sizeof(ref_bounds) * (size_t)R, cudaMemcpyDeviceToHost));
    sah_topology topo;
    sah_build(href, topo);
    YH_CUDA(cudaMemcpy(left.ptr, topo.left.data(), 4 * (size_t)ni, cudaMemcpyHostToDevice));
    YH_CUDA(cudaMemcpy(right.ptr, topo.right.data(), 4 * (size_t)ni, cudaMemcpyHostToDevice));
    YH_CUDA(cudaMemcpy(parent_inner.ptr, topo.parent_inner.data(), 4 * (size_t)ni, cudaMemcpyHostToDevice));
    YH_CUDA(cudaMemcpy(parent_leaf.ptr, topo.parent_leaf.data(), 4 * (size_t)R, cudaMemcpyHostToDevice));
    YH_CUDA(cudaMemcpy(ids_out.ptr, topo.order.data(), 4 * (size_t)R, cudaMemcpyHostToDevice));
  } else {
    k_hierarchy<<<G(ni), B>>>(keys_out.ptr, R, left.ptr, right.ptr, first.ptr, last.ptr, parent_inner.ptr, parent_leaf.ptr);
    YH_CUDA(cudaGetLastError());
  }
  k_refit<<<G(R), B>>>(R, ids_out.ptr, refs.ptr, left.ptr, right.ptr, parent_inner.ptr, parent_leaf.ptr, flags.ptr, node_box.ptr);
  YH_CUDA(cudaGetLastError());
  // quality passes over the radix tree: 3 rounds of k_rotate (0.3 ms each at a million segments) take 4.6 % off
  // the node visits per ray on C3 (YH_ROTATE_PASSES overrides the count; each round can deepen a path by one level,
  // which the traversal stack - 96 entries against a radix-tree depth of at most 91 - has room for)
  if (quality == 0) {
    int passes = 3;
    if (const char* e = getenv("YH_ROTATE_PASSES")) passes = std::min(4, std::max(0, atoi(e)));
    dev_buffer<int> rotations;
    YH_CUDA(rotations.alloc(1));
    // a rotation moves one subtree a level down; k_max_depth checks the result against the traversal stack and,
    // should an input ever drive a path past the limit, the build falls back to the plain radix tree
    const int kDepthLimit = 92;  // kStackShared + kStackLocal = 96 entries (yh_traverse.cuh)
    for (int pass = 0; pass < passes; pass++) {
      YH_CUDA(cudaMemset(flags.ptr, 0, flags.bytes()));
      YH_CUDA(cudaMemset(rotations.ptr, 0, 4));
      k_rotate<<<G(R), B>>>(R, ids_out.ptr, refs.ptr, left.ptr, right.ptr, parent_inner.ptr, parent_leaf.ptr, flags.ptr,
          node_box.ptr, rotations.ptr);
      YH_CUDA(cudaGetLastError());
      if (timing) {
        int nrot = 0;
        YH_CUDA(cudaMemcpy(&nrot, rotations.ptr, 4, cudaMemcpyDeviceToHost));
        fprintf(stderr, "  bvh rotation pass %d: %d rotations\n", pass, nrot);
      }
    }
    if (passes > 0) {
      int depth = 0;
      YH_CUDA(cudaMemset(rotations.ptr, 0, 4));
      k_max_depth<<<G(R), B>>>(R, parent_inner.ptr, parent_leaf.ptr, rotations.ptr);
      YH_CUDA(cudaMemcpy(&depth, rotations.ptr, 4, cudaMemcpyDeviceToHost));
      if (timing) fprintf(stderr, "  bvh depth after the rotations: %d\n", depth);
      if (depth > kDepthLimit) {  // rebuild the untouched radix tree
        YH_CUDA(cudaMemset(flags.ptr, 0, flags.bytes()));
        k_hierarchy<<<G(ni), B>>>(keys_out.ptr, R, left.ptr, right.ptr, first.ptr, last.ptr, parent_inner.ptr, parent_leaf.ptr);
        k_refit<<<G(R), B>>>(R, ids_out.ptr, refs.ptr, left.ptr, right.ptr, parent_inner.ptr, parent_leaf.ptr, flags.ptr, node_box.ptr);
        YH_CUDA(cudaGetLastError());
      }
    }
  }
  k_emit_packed<<<G(ni), B>>>(kind, ni, left.ptr, right.ptr, parent_inner.ptr, ids_out.ptr, refs.ptr, node_box.ptr,
      ref_prim.ptr, sh.d_elems.ptr, sh.d_positions.ptr, sh.d_radius.ptr, sh.grid_origin, inv_step, pad_steps, sh.d_slots.ptr,
      sh.d_slot_elem.ptr);
  YH_CUDA(cudaGetLastError());
  mark("hierarchy + refit + emit");
  ref_bounds rb;
  YH_CUDA(cudaMemcpy(&rb, node_box.ptr, sizeof(rb), cudaMemcpyDeviceToHost));
  sh.root_bbox = {{rb.lo.x, rb.lo.y, rb.lo.z}, {rb.hi.x, rb.hi.y, rb.hi.z}};
  YH_CUDA(cudaDeviceSynchronize());
  mark("root box");
  // a tree of at most kTinyRefs leaves (the quad of an area light): its leaf slots, so that the light-pdf rays test the
  // primitives directly instead of walking boxes around two triangles (intersect_shape)
  sh.tiny_count = 0;
  if (R <= kTinyRefs) {
    std::vector<uint4> slots(2 * num_slots);
    YH_CUDA(cudaMemcpy(slots.data(), sh.d_slots.ptr, slots.size() * sizeof(uint4), cudaMemcpyDeviceToHost));
    std::vector<int> open = {0};
    while (!open.empty() && sh.tiny_count <= kTinyRefs) {
      int node = open.back();
      open.pop_back();
      uint4 hi = slots[2 * (size_t)node + 1];  // words 4-7: w6 = first child, w7 bits 31 / 30 = child 0 / 1 is a leaf
      int   child = (int)hi.z;
      for (int c = 1; c >= 0; c--) {
        bool leaf = c == 0 ? (int)hi.w < 0 : (int)(hi.w << 1) < 0;
        if (child + c < 0 || child + c >= (int)num_slots) { sh.tiny_count = kTinyRefs + 1; break; }
        if (leaf) { if (sh.tiny_count < kTinyRefs) sh.tiny_slots[sh.tiny_count] = child + c; sh.tiny_count++; }
        else open.push_back(child + c);
      }
    }
    if (sh.tiny_count != R) sh.tiny_count = 0;  // not the expected tree: keep the general walk
  }
  return YH_OK;
}

// per-material part of eval_hair_brdf + the denominators of mp(), with the device's own arithmetic
__global__ void k_material_hair_consts(dev_material* materials, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dev_material& m  = materials[i];
  hair_material hm = {m.sigma_a, m.beta_m, m.beta_n, m.alpha, m.eta, m.color, m.eumelanin, m.pheomelanin};
  hair_brdf     b{};
  hair_brdf_material_part(hm, b);
  m.hair_proto = b;
  m.hair_k = make_hair_consts(b);
}

// transform_bbox(frame, bbox), yocto_math.h:3172-3182: the exact form the reference feeds to its top-level builder
static bbox3f transform_bbox_exact(const frame3f& f, const bbox3f& b) {
  bbox3f out = {{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
  for (int c = 0; c < 8; c++) {  // corner order of the reference: x slowest ... any order gives the same min / max
    vec3f p = {c & 4 ? b.max.x : b.min.x, c & 2 ? b.max.y : b.min.y, c & 1 ? b.max.z : b.min.z};
    vec3f q = transform_point(f, p);
    out.min = min3(out.min, q);
    out.max = max3(out.max, q);
  }
  return out;
}

// transform_bbox(frame, bbox), yocto_math.h:3172-3182, plus padding
static bbox3f transform_bbox_padded(const frame3f& f, const bbox3f& b) {
  bbox3f out = {{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
  for (int c = 0; c < 8; c++) {
    vec3f p = {c & 4 ? b.max.x : b.min.x, c & 2 ? b.max.y : b.min.y, c & 1 ? b.max.z : b.min.z};
    vec3f q = transform_point(f, p);
    out.min = min3(out.min, q);
    out.max = max3(out.max, q);
  }
  float m   = std::max({std::fabs(out.min.x), std::fabs(out.min.y), std::fabs(out.min.z), std::fabs(out.max.x),
      std::fabs(out.max.y), std::fabs(out.max.z)});
  float eps = 4e-6f * m + 1e-30f;
  out.min   = out.min - eps;
  out.max   = out.max + eps;
  return out;
}

}  // namespace yh

using namespace yh;

static yh_status upload_tables(yh_scene* sc);

yh::dev_scene yh_scene::view() const {
  dev_scene v{};
  v.shapes = d_shapes.ptr, v.materials = d_materials.ptr, v.objects = d_objects.ptr;
  v.textures = d_textures.ptr, v.environments = d_environments.ptr, v.lights = d_lights.ptr;
  v.num_shapes = (int)shapes.size(), v.num_materials = (int)materials.size();
  v.num_objects = (int)objects.size(), v.num_textures = (int)textures.size();
  v.num_environments = (int)environments.size(), v.num_lights = (int)lights.size();
  v.scene_nodes = d_scene_nodes.ptr, v.scene_prims = d_scene_prims.ptr, v.num_scene_nodes = (int)d_scene_nodes.count / 2;
  return v;
}
// The reference's top-level tree (init_bvh :793-814 -> build_bvh :601-650) over the objects: primitive = object,
// bounds = transform_bbox(frame, root box of the shape's tree) or the invalid box for an empty shape, split_middle,
// leaves of <= 4.  Only topology, split axes and leaf order are the reference's; the node boxes are unions of OUR
// padded object bounds.
static void build_scene_tree(const yh_scene* sc, std::vector<float4>& nodes, std::vector<int>& prims_out) {
  int n = (int)sc->objects.size();
  nodes.clear(), prims_out.clear();
  if (n == 0) return;
  std::vector<ref_prim> prims(n);
  for (int i = 0; i < n; i++) {
    const host_shape* s = sc->shapes[sc->objects[i].shape];
    bbox3f b = {{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};  // invalidb3f
    if (s->num_elems() > 0) b = transform_bbox_exact(sc->objects[i].frame, s->ref_root_bbox);
    prims[i] = {(b.min + b.max) / 2.0f, i};
  }
  struct item { int node, start, end; };
  std::vector<item> todo{{0, 0, n}};
  nodes.resize(2);
  while (!todo.empty()) {
    item it = todo.back();
    todo.pop_back();
    bbox3f box = {{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
    for (int i = it.start; i < it.end; i++) {
      const bbox3f& ob = sc->objects[prims[i].primitive].bbox;
      if (ob.min.x <= ob.max.x) box = {min3(box.min, ob.min), max3(box.max, ob.max)};
    }
    int a, b;
    if (it.end - it.start > kRefLeafMax) {
      int mid, axis;
      reference_split(prims.data(), it.start, it.end, mid, axis);
      a = (int)nodes.size() / 2, b = 0x100 | axis;
      nodes.resize(nodes.size() + 4);
      todo.push_back({a, it.start, mid});
      todo.push_back({a + 1, mid, it.end});
    } else {
      a = it.start, b = (it.end - it.start) << 16;
    }
    auto bits = [](int v) { float f; memcpy(&f, &v, 4); return f; };
    nodes[2 * it.node]     = make_float4(box.min.x, box.min.y, box.min.z, bits(a));
    nodes[2 * it.node + 1] = make_float4(box.max.x, box.max.y, box.max.z, bits(b));
  }
  prims_out.resize(n);
  for (int i = 0; i < n; i++) prims_out[i] = prims[i].primitive;
}

yh_status yh_scene::sync_scene_tree() const {
  std::lock_guard<std::mutex> lock(tree_mutex);
  if (!scene_tree_pending) return YH_OK;
  auto self = const_cast<yh_scene*>(this);
  std::vector<float4> nodes;
  std::vector<int>    prims;
  build_scene_tree(this, nodes, prims);
  YH_CUDA(self->d_scene_nodes.upload(nodes.data(), nodes.size()));
  YH_CUDA(self->d_scene_prims.upload(prims.data(), prims.size()));
  scene_tree_pending = false;
  return YH_OK;
}

yh_scene::~yh_scene() {
  for (auto s : shapes) delete s;
  for (auto t : tex_dev) delete t;
  for (auto c : light_cdfs) delete c;
  for (auto g : light_guides) delete g;
}

static frame3f to_frame(const float* f) {
  return {{f[0], f[1], f[2]}, {f[3], f[4], f[5]}, {f[6], f[7], f[8]}, {f[9], f[10], f[11]}};
}
static vec3f to_vec(const float* f) { return {f[0], f[1], f[2]}; }

static yh_status upload_tables(yh_scene* sc) {
  std::vector<dev_shape> ds;
  for (auto s : sc->shapes) {
    dev_shape d{};
    d.slots = s->d_slots.ptr, d.tris = s->d_prims.ptr, d.slot_elem = s->d_slot_elem.ptr;
    d.ref_pos = s->d_ref_pos.ptr, d.tie_nodes = s->d_tie_nodes.ptr, d.num_tie_nodes = (int)s->d_tie_nodes.count;
    d.root_leaf = s->root_leaf, d.tiny_count = s->tiny_count, d.num_refs = s->num_refs;
    for (int t = 0; t < 4; t++) d.tiny_slots[t] = s->tiny_slots[t];
    d.is_lines = s->is_lines() ? 1 : 0, d.is_points = s->is_points() ? 1 : 0;
    d.scale = s->scale, d.grid_origin = s->grid_origin, d.grid_step = s->grid_step;
    d.has_normals = s->normals.empty() ? 0 : 1;
    d.positions = s->d_positions.ptr, d.normals = s->d_normals.ptr, d.elems = s->d_elems.ptr;
    d.texcoords = s->texcoords.empty() ? nullptr : s->d_texcoords.ptr;
    d.num_elems = s->num_elems();
    ds.push_back(d);
  }
  YH_CUDA(sc->d_shapes.upload(ds.data(), ds.size()));
  YH_CUDA(sc->d_materials.upload(sc->materials.data(), sc->materials.size()));
  if (!sc->materials.empty()) {
    k_material_hair_consts<<<((int)sc->materials.size() + 63) / 64, 64>>>(sc->d_materials.ptr, (int)sc->materials.size());
    YH_CUDA(cudaGetLastError());
    YH_CUDA(cudaDeviceSynchronize());
  }
  YH_CUDA(sc->d_objects.upload(sc->objects.data(), sc->objects.size()));
  YH_CUDA(sc->d_textures.upload(sc->textures.data(), sc->textures.size()));
  YH_CUDA(sc->d_environments.upload(sc->environments.data(), sc->environments.size()));
  YH_CUDA(sc->d_lights.upload(sc->lights.data(), sc->lights.size()));
  return YH_OK;
}

extern "C" {

yh_status yh_scene_create(const yh_scene_desc* d, yh_scene** out) {
  YH_REQUIRE_DEVICE();
  if (!d || !out) return fail(YH_ERR_INVALID, "yh_scene_create: null argument");
  *out = nullptr;
  // ---- validate: anything outside the hot-path scope fails loudly ----
  for (int i = 0; i < d->num_materials; i++) {
    auto& m = d->materials[i];
    int texs[] = {m.emission_tex, m.color_tex, m.specular_tex, m.metallic_tex, m.roughness_tex,
        m.transmission_tex, m.spectint_tex, m.scattering_tex, m.opacity_tex, m.normal_tex};
    for (int t : texs)
      if (t >= d->num_textures) return fail(YH_ERR_INVALID, "material %d: texture index out of range", i);
  }
  for (int i = 0; i < d->num_objects; i++) {
    auto& o = d->objects[i];
    if (o.shape < 0 || o.shape >= d->num_shapes || o.material < 0 || o.material >= d->num_materials)
      return fail(YH_ERR_INVALID, "object %d: shape/material index out of range", i);
    // On a line shape the hair lobe replaces eval / sample / pdf of all surface lobes (brdf.hair, :474-490), so
    // specular / metallic have no effect - unless they make is_delta() (:495) true: with diffuse == 0 and roughness == 0
    // the reference runs its delta-SURFACE branch on hair.  The hair kernel carries no surface lobes, so that corner
    // (and refraction, which would open a volume behind a fibre, :1458-1467) is refused.
    auto& m = d->materials[o.material];
    if (d->shapes[o.shape].num_lines > 0 && d->shapes[o.shape].num_points == 0) {
      if (m.transmission != 0)
        return fail(YH_ERR_UNSUPPORTED, "object %d: transmission on a line shape is not built", i);
      if ((m.specular != 0 || m.metallic != 0) && (m.roughness == 0 || m.roughness_tex >= 0))
        return fail(YH_ERR_UNSUPPORTED, "object %d: specular / metallic with zero or textured roughness on a line shape is not built", i);
    }
  }
  for (int i = 0; i < d->num_environments; i++)
    if (d->environments[i].emission_tex >= d->num_textures)
      return fail(YH_ERR_INVALID, "environment %d: texture index out of range", i);
  for (int i = 0; i < d->num_shapes; i++) {
    auto& s = d->shapes[i];
    // the reference dispatches points > lines > triangles in its traversal but triangles > lines > points when it
    // shades (:232-310, :894-925): a shape with two kinds is something its loaders never produce
    if ((s.num_lines > 0) + (s.num_triangles > 0) + (s.num_points > 0) > 1)
      return fail(YH_ERR_INVALID, "shape %d mixes points / lines / triangles", i);
    if ((s.num_lines || s.num_points) && s.num_radius != s.num_positions)
      return fail(YH_ERR_INVALID, "shape %d: points and lines need one radius per vertex", i);
    for (int k = 0; k < s.num_points; k++)
      if (s.points[k] < 0 || s.points[k] >= s.num_positions) return fail(YH_ERR_INVALID, "shape %d: point index out of range", i);
    if (s.num_normals && s.num_normals != s.num_positions)
      return fail(YH_ERR_INVALID, "shape %d: normals/positions size mismatch", i);
    if (s.num_texcoords && s.num_texcoords != s.num_positions)
      return fail(YH_ERR_INVALID, "shape %d: texcoords/positions size mismatch", i);
    for (int k = 0; k < 2 * s.num_lines; k++)
      if (s.lines[k] < 0 || s.lines[k] >= s.num_positions) return fail(YH_ERR_INVALID, "shape %d: line index out of range", i);
    for (int k = 0; k < 3 * s.num_triangles; k++)
      if (s.triangles[k] < 0 || s.triangles[k] >= s.num_positions)
        return fail(YH_ERR_INVALID, "shape %d: triangle index out of range", i);
  }
  auto sc = new yh_scene{};
  cudaGetDevice(&sc->device);
  auto bail = [&](yh_status st) {
    delete sc;
    return st;
  };
  for (int i = 0; i < d->num_shapes; i++) {
    auto& s  = d->shapes[i];
    auto  hs = new host_shape{};
    sc->shapes.push_back(hs);
    hs->positions.assign(s.positions, s.positions + 3 * (size_t)s.num_positions);
    if (s.num_normals) hs->normals.assign(s.normals, s.normals + 3 * (size_t)s.num_normals);
    if (s.num_radius) hs->radius.assign(s.radius, s.radius + s.num_radius);
    if (s.num_texcoords) hs->texcoords.assign(s.texcoords, s.texcoords + 2 * (size_t)s.num_texcoords);
    if (s.num_points) hs->points.assign(s.points, s.points + s.num_points);
    if (s.num_lines) hs->lines.assign(s.lines, s.lines + 2 * (size_t)s.num_lines);
    if (s.num_triangles) hs->triangles.assign(s.triangles, s.triangles + 3 * (size_t)s.num_triangles);
    if (hs->d_positions.upload(hs->positions.data(), hs->positions.size()) != cudaSuccess ||
        hs->d_normals.upload(hs->normals.data(), hs->normals.size()) != cudaSuccess ||
        hs->d_radius.upload(hs->radius.data(), hs->radius.size()) != cudaSuccess ||
        hs->d_texcoords.upload(hs->texcoords.data(), hs->texcoords.size()) != cudaSuccess ||
        hs->d_elems.upload(hs->elems().data(), hs->elems().size()) != cudaSuccess)
      return bail(fail(YH_ERR_CUDA, "shape %d: device upload failed: %s", i, cudaGetErrorString(cudaGetLastError())));
  }
  for (int i = 0; i < d->num_textures; i++) {
    auto& t = d->textures[i];
    if (!t.colorf || t.width <= 0 || t.height <= 0) return bail(fail(YH_ERR_INVALID, "texture %d is empty", i));
    sc->tex_host.emplace_back(t.colorf, t.colorf + 3 * (size_t)t.width * t.height);
    auto buf = new dev_buffer<float>{};
    sc->tex_dev.push_back(buf);
    if (buf->upload(t.colorf, 3 * (size_t)t.width * t.height) != cudaSuccess)
      return bail(fail(YH_ERR_CUDA, "texture %d: device upload failed", i));
    sc->textures.push_back({buf->ptr, t.width, t.height});
  }
  for (int i = 0; i < d->num_materials; i++) {
    auto&        m = d->materials[i];
    dev_material dm{};
    dm.emission = to_vec(m.emission), dm.color = to_vec(m.color), dm.opacity = m.opacity, dm.thin = m.thin;
    dm.specular = m.specular, dm.metallic = m.metallic, dm.roughness = m.roughness, dm.ior = m.ior, dm.transmission = m.transmission;
    dm.scattering = to_vec(m.scattering), dm.scanisotropy = m.scanisotropy, dm.trdepth = m.trdepth;
    if (!m.thin && m.transmission != 0) sc->has_volumes = true;  // has_volume(), yocto_pathtrace.cpp:529-531
    dm.emission_tex = m.emission_tex, dm.color_tex = m.color_tex, dm.specular_tex = m.specular_tex;
    dm.metallic_tex = m.metallic_tex, dm.roughness_tex = m.roughness_tex, dm.transmission_tex = m.transmission_tex;
    dm.opacity_tex = m.opacity_tex, dm.scattering_tex = m.scattering_tex, dm.normal_tex = m.normal_tex;
    dm.any_tex = (m.emission_tex >= 0 || m.color_tex >= 0 || m.specular_tex >= 0 || m.metallic_tex >= 0 || m.roughness_tex >= 0 ||
                  m.transmission_tex >= 0 || m.opacity_tex >= 0 || m.scattering_tex >= 0) ? 1 : 0;
    dm.eumelanin = m.eumelanin, dm.pheomelanin = m.pheomelanin, dm.sigma_a = to_vec(m.sigma_a);
    dm.beta_m = m.beta_m, dm.beta_n = m.beta_n, dm.alpha = m.alpha, dm.eta = m.eta;
    sc->materials.push_back(dm);
  }
  for (int i = 0; i < d->num_objects; i++) {
    dev_object o{};
    o.frame     = to_frame(d->objects[i].frame);
    o.inv_frame = inverse_nonrigid(o.frame);  // inverse(frame, non_rigid_frames = true), :1012
    o.shape = d->objects[i].shape, o.material = d->objects[i].material;
    sc->objects.push_back(o);
  }
  for (int i = 0; i < d->num_environments; i++) {
    dev_environment e{};
    e.frame     = to_frame(d->environments[i].frame);
    e.inv_frame = inverse_rigid(e.frame);  // inverse(environment->frame), :539
    e.emission  = to_vec(d->environments[i].emission);
    e.tex       = d->environments[i].emission_tex;
    sc->environments.push_back(e);
  }
  yh_status st = upload_tables(sc);
  if (st != YH_OK) return bail(st);
  *out = sc;
  return YH_OK;
}

void yh_scene_destroy(yh_scene* scene) { delete scene; }

yh_status yh_scene_init_bvh(yh_scene* sc, const yh_bvh_params* params) {
  YH_REQUIRE_DEVICE();
  if (!sc) return fail(YH_ERR_INVALID, "yh_scene_init_bvh: scene is null");
  int max_split = params && params->max_split > 0 ? params->max_split : kDefaultSplit;
  if (max_split > 8) max_split = 8;
  sc->stats = yh_bvh_stats{};
  for (auto s : sc->shapes) {
    yh_status st = build_shape_bvh(*s, max_split, params ? params->quality : 0);
    if (st != YH_OK) return st;
    st = build_tie_order(*s);  // the reference's leaf order (exact-t ties) and its exact root box, on the device
    if (st != YH_OK) return st;
    sc->stats.num_primitives += s->num_elems();
    sc->stats.num_references += s->num_refs;
    sc->stats.num_nodes += s->num_nodes;
    sc->stats.node_bytes += (int64_t)s->d_slots.bytes();
    sc->stats.prim_bytes += (int64_t)s->d_prims.bytes() + (int64_t)s->d_slot_elem.bytes();
  }
  for (auto& o : sc->objects) {
    auto s = sc->shapes[o.shape];
    o.bbox = s->num_refs ? transform_bbox_padded(o.frame, s->root_bbox)
                         : bbox3f{{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
  }
  sc->world_bbox = {{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
  for (auto& o : sc->objects)
    if (o.bbox.min.x <= o.bbox.max.x) sc->world_bbox = {min3(sc->world_bbox.min, o.bbox.min), max3(sc->world_bbox.max, o.bbox.max)};
  if (!(sc->world_bbox.min.x <= sc->world_bbox.max.x)) sc->world_bbox = {{0, 0, 0}, {0, 0, 0}};
  sc->bvh_built = true;
  {
    std::lock_guard<std::mutex> lock(sc->tree_mutex);
    sc->scene_tree_pending = true;  // its node boxes are unions of the bounds just computed
  }
  yh_status st = sc->sync_scene_tree();
  if (st != YH_OK) return st;
  return upload_tables(sc);
}

yh_status yh_scene_reference_order(const yh_scene* sc, int shape, int* order, int64_t* count) {
  YH_REQUIRE_DEVICE();
  if (!sc || shape < 0 || shape >= (int)sc->shapes.size() || !count) return fail(YH_ERR_INVALID, "yh_scene_reference_order: bad argument");
  if (!sc->bvh_built) return fail(YH_ERR_STATE, "yh_scene_reference_order: call yh_scene_init_bvh first");
  const host_shape* s = sc->shapes[shape];
  int               n = s->num_elems();
  *count = n;
  if (!order || n == 0) return YH_OK;
  std::vector<int> ref_pos(n);
  YH_CUDA(cudaMemcpy(ref_pos.data(), s->d_ref_pos.ptr, 4 * (size_t)n, cudaMemcpyDeviceToHost));
  for (int e = 0; e < n; e++) order[ref_pos[e]] = e;
  return YH_OK;
}

yh_status yh_scene_bvh_stats(const yh_scene* sc, yh_bvh_stats* stats) {
  if (!sc || !stats) return fail(YH_ERR_INVALID, "yh_scene_bvh_stats: null argument");
  if (!sc->bvh_built) return fail(YH_ERR_STATE, "yh_scene_bvh_stats: call yh_scene_init_bvh first");
  *stats = sc->stats;
  return YH_OK;
}

// init_lights, yocto_pathtrace.cpp:1695-1740: one light per emissive triangle
// object (cumulative triangle areas) and per emissive environment (cumulative
// max(rgb) * sin(theta) per texel).  The running float sums are sequential in
// the reference, so they are formed on the host in the same order.
yh_status yh_scene_init_lights(yh_scene* sc) {
  YH_REQUIRE_DEVICE();
  if (!sc) return fail(YH_ERR_INVALID, "yh_scene_init_lights: scene is null");
  for (auto c : sc->light_cdfs) delete c;
  for (auto g : sc->light_guides) delete g;
  sc->light_cdfs.clear();
  sc->light_guides.clear();
  sc->lights.clear();
  for (int oi = 0; oi < (int)sc->objects.size(); oi++) {
    auto& o = sc->objects[oi];
    auto& m = sc->materials[o.material];
    if (m.emission == vec3f{0, 0, 0}) continue;
    auto s = sc->shapes[o.shape];
    if (s->triangles.empty()) continue;
    int                nt = (int)s->triangles.size() / 3;
    std::vector<float> cdf(nt);
    for (int idx = 0; idx < nt; idx++) {
      auto  p  = [&](int k) { int v = s->triangles[3 * idx + k]; return vec3f{s->positions[3 * v], s->positions[3 * v + 1], s->positions[3 * v + 2]}; };
      vec3f p0 = p(0), p1 = p(1), p2 = p(2);
      cdf[idx] = length(cross(p1 - p0, p2 - p0)) / 2;  // triangle_area, yocto_math.h:3306
      if (idx) cdf[idx] += cdf[idx - 1];
    }
    auto buf = new dev_buffer<float>{};
    sc->light_cdfs.push_back(buf);
    YH_CUDA(buf->upload(cdf.data(), cdf.size()));
    sc->lights.push_back({oi, -1, buf->ptr, nt, nullptr, 0, 0.0f, 0.0f, nullptr});
  }
  for (int ei = 0; ei < (int)sc->environments.size(); ei++) {
    auto& e = sc->environments[ei];
    if (e.emission == vec3f{0, 0, 0}) continue;
    dev_light l{-1, ei, nullptr, 0, nullptr, 0, 0.0f, 0.0f, nullptr};
    if (e.tex >= 0) {
      int                w = sc->textures[e.tex].width, h = sc->textures[e.tex].height;
      auto&              px = sc->tex_host[e.tex];
      std::vector<float> cdf((size_t)w * h);
      // sin(theta) depends on the row only: one libm call per row instead of one per texel (the same float either way);
      // the running sum stays sequential, in the reference's order (:1731-1735)
      std::vector<float> row_sin(h);
      for (int iy = 0; iy < h; iy++) row_sin[iy] = std::sin((iy + 0.5f) * pif / h);
      float run = 0;
      for (size_t i = 0; i < cdf.size(); i++) {
        float mx = fmax_(fmax_(px[3 * i], px[3 * i + 1]), px[3 * i + 2]);
        float v  = mx * row_sin[i / w];
        run      = i ? v + run : v;
        cdf[i]   = run;
      }
      auto buf = new dev_buffer<float>{};
      sc->light_cdfs.push_back(buf);
      YH_CUDA(buf->upload(cdf.data(), cdf.size()));
      l.cdf = buf->ptr, l.cdf_size = (int)cdf.size();
      std::vector<float> row_angle(h);
      for (int j = 0; j < h; j++)
        row_angle[j] = div_(2 * pif, (float)w) * div_(pif, (float)h) * sinf_(div_(pif * (j + 0.5f), (float)h));
      auto abuf = new dev_buffer<float>{};
      sc->light_cdfs.push_back(abuf);
      YH_CUDA(abuf->upload(row_angle.data(), row_angle.size()));
      l.row_angle = abuf->ptr;
      // guide table (one bucket per texel)
      float back = cdf.back();
      int   G    = (int)cdf.size();
      if (back > 0 && std::isfinite(back)) {
        float            gw = back / (float)G;
        std::vector<int> guide((size_t)G + 1);
        size_t           j = 0;
        for (int b = 0; b <= G; b++) {  // thresholds grow with b: one sweep = upper_bound per bucket
          float t = (float)b * gw;
          while (j < cdf.size() && !(t < cdf[j])) j++;
          guide[b] = (int)j;
        }
        auto gbuf = new dev_buffer<int>{};
        sc->light_guides.push_back(gbuf);
        YH_CUDA(gbuf->upload(guide.data(), guide.size()));
        l.guide = gbuf->ptr, l.guide_size = G, l.guide_w = gw, l.guide_inv_w = 1.0f / gw;
      }
    }
    sc->lights.push_back(l);
  }
  sc->lights_built = true;
  return upload_tables(sc);
}

}  // extern "C"

// Closest-hit traversal of the two-level scene: objects (linear, bounds
// culled) -> shape LBVH (yh_scene.cuh layout).  Restates intersect_scene_bvh /
// intersect_shape_bvh (yocto_pathtrace.cpp:821-1028): the ray is taken to
// object space with inverse(frame, non_rigid = true), the direction is NOT
// renormalised (t is preserved), and every primitive test is the reference's
// own arithmetic, so object / element / uv / distance agree bit for bit.  The
// box tests are free to differ: boxes are padded, hence conservative.
#pragma once
#include "yh_scene.cuh"

namespace yh {

#ifndef YH_STACK_SHARED
#define YH_STACK_SHARED 16
#endif
constexpr int kStackShared = YH_STACK_SHARED;  // entries per thread kept in shared memory
constexpr int kStackLocal  = 96 - kStackShared;  // overflow (LBVH chains can be deep)
constexpr int kTraceBlock  = 128;

struct hit_record {
  int   object, element;
  vec2f uv;
  float distance;
  bool  hit;
};

// Per-thread traversal stack: the first kStackShared entries are interleaved
// in shared memory (entry e of thread t at smem[e * blockDim.x + t], so a warp
// touches 32 consecutive banks), deeper entries spill to local memory.
struct trav_stack {
  int* smem;
  int  local[kStackLocal];
  int  sp;
  __device__ __forceinline__ void push(int v) {
    if (sp < kStackShared) smem[sp * kTraceBlock] = v;
    else local[sp - kStackShared] = v;
    sp++;
  }
  __device__ __forceinline__ int pop() {
    sp--;
    return sp < kStackShared ? smem[sp * kTraceBlock] : local[sp - kStackShared];
  }
};

#ifdef YH_TRAVERSAL_STATS
#define YH_STAT(x) x
#else
#define YH_STAT(x)
#endif

// True when the reference's traversal (intersect_shape_bvh,
// yocto_pathtrace.cpp:866-925) reaches element eb AFTER element ea for a ray
// with these direction signs.  Its DFS pops the child on the ray's near side
// of the split axis first (:879-886) and walks a leaf in primitive order, and
// a hit at t == tmax replaces the stored one, so among exact ties the element
// visited last is the one it returns.
__device__ __noinline__ bool reference_visits_later(const dev_shape& sh, int ea, int eb, unsigned dsign) {
  int pa = __ldg(sh.ref_pos + ea), pb = __ldg(sh.ref_pos + eb);
  if (pa == pb) return false;
  int  lo = min(pa, pb), hi = max(pa, pb);
  bool hi_later = true;  // same leaf: position order
  int  node     = sh.num_tie_nodes > 0 ? 0 : -1;
  while (node >= 0) {
    int4 n = __ldg(sh.tie_nodes + node);
    if (hi < n.x) node = n.z;
    else if (lo >= n.x) node = n.w;
    else {  // lowest common ancestor: lo in child 0, hi in child 1
      hi_later = ((dsign >> n.y) & 1u) == 0;  // positive direction: child 0 first
      break;
    }
  }
  return (pb == hi) == hi_later;
}

// Certain-miss test for a line primitive, evaluated in fast arithmetic (FMA,
// approximate sqrt) BEFORE the reference's own test.  intersect_line
// (yocto_math.h:3426-3469) accepts only when a point of the ray line and a
// point of the segment line are closer than max(r0, r1); both are at least
// the line-line distance L = |w . n| / |n|, n = u x v, apart.  So
//     L > max(r0, r1) + (rounding of either evaluation)
// proves the reference rejects.  The slack covers the reference's rounding of
// pr / pl (a few ulp of the coordinates) and this evaluation's own (n is only
// trusted when the lines are more than ~1.8 degrees apart).  A primitive that is
// not provably missed takes the exact path, so hit records are unchanged.
__device__ __forceinline__ bool line_certain_miss(const vec3f& o, const vec3f& u, float uu, float mo,
    const float4& a, const float4& b) {
  float vx = b.x - a.x, vy = b.y - a.y, vz = b.z - a.z;
  float wx = o.x - a.x, wy = o.y - a.y, wz = o.z - a.z;
  float nx = __fmaf_rn(u.y, vz, -u.z * vy), ny = __fmaf_rn(u.z, vx, -u.x * vz), nz = __fmaf_rn(u.x, vy, -u.y * vx);
  float nn = __fmaf_rn(nx, nx, __fmaf_rn(ny, ny, nz * nz));
  float vv = __fmaf_rn(vx, vx, __fmaf_rn(vy, vy, vz * vz));
  float ww = __fmaf_rn(wx, wx, __fmaf_rn(wy, wy, wz * wz));
  float wn = __fmaf_rn(wx, nx, __fmaf_rn(wy, ny, wz * nz));
  float m1  = mo + fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(b.x) + fabsf(b.y) + fabsf(b.z);
  float tol = __fmaf_rn(fmaxf(a.w, b.w), 1.01f, __fmaf_rn(2e-5f, __fsqrt_ru(ww), 4e-6f * m1));
  return nn > 1e-3f * uu * vv && wn * wn > nn * tol * tol;
}

constexpr int kTraverseDone = (int)0x80000000;  // never a leaf code: those are ~x with x < 2^31

// Traverses one shape with an object-space ray.  On a hit updates ray.tmax,
// ref (leaf-order reference index), uv, and returns true.  while-while form:
// every lane first walks inner nodes until it holds a leaf (or is done), then
// the warp tests its leaves together, so the long primitive test does not
// serialise against the short node step.
//
// Box test: slab distances as ONE FMA per plane, plane * (1/d) + (-o/d).  The
// rounding of -o/d displaces a plane by <= 6e-8 |o| in space; the builder pads
// every box by 4e-7 * (largest coordinate of the shape), which covers it for
// |o| <= 4 * that scale.  Rays that start farther away (FUSED = false) use the
// (plane - o) * (1/d) form, exact in the difference like the reference's
// intersect_bbox (yocto_math.h:3544-3554).  A zero direction component is
// replaced by +-1e-20 in the fused form (1/0 = inf would turn the FMA into
// inf - inf); the slab then still rejects origins outside it and passes
// origins inside, as the reference does.
template <bool FUSED>
__device__ __forceinline__ bool intersect_shape_impl(const dev_shape& sh, ray3f& ray, trav_stack& st, int& ref,
    vec2f& uv, bool find_any, unsigned long long* nodes_visited, unsigned long long* prims_tested) {
  bool  hit  = false;
  // ray_dsign, :862-863
  unsigned dsign = ((1.0f / ray.d.x) < 0 ? 1u : 0u) | ((1.0f / ray.d.y) < 0 ? 2u : 0u) | ((1.0f / ray.d.z) < 0 ? 4u : 0u);
  auto     safe  = [](float d) { return FUSED && fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d; };
  float idx = 1.0f / safe(ray.d.x), idy = 1.0f / safe(ray.d.y), idz = 1.0f / safe(ray.d.z);
  float ox = FUSED ? -ray.o.x * idx : ray.o.x, oy = FUSED ? -ray.o.y * idy : ray.o.y,
        oz = FUSED ? -ray.o.z * idz : ray.o.z;
  // slab distance of plane p along one axis
  auto slab = [](float p, float id, float o) { return FUSED ? __fmaf_rn(p, id, o) : (p - o) * id; };
  float uu = __fmaf_rn(ray.d.x, ray.d.x, __fmaf_rn(ray.d.y, ray.d.y, ray.d.z * ray.d.z));
  float mo = fabsf(ray.o.x) + fabsf(ray.o.y) + fabsf(ray.o.z);
  const float4* __restrict__ nodes = sh.nodes;
  const float4* __restrict__ prims = sh.prims;
  const bool is_lines = sh.is_lines != 0;
  int base = st.sp;
  int cur  = sh.root;
  while (true) {
    while (cur >= 0) {
      const float4 n0 = __ldg(nodes + 4 * (size_t)cur), n1 = __ldg(nodes + 4 * (size_t)cur + 1),
                   n2 = __ldg(nodes + 4 * (size_t)cur + 2), n3 = __ldg(nodes + 4 * (size_t)cur + 3);
      YH_STAT(if (nodes_visited)(*nodes_visited)++;)
      // child 0: min = n0.xyz, max = n0.w n1.x n1.y ; child 1: min = n1.z n1.w n2.x, max = n2.yzw
      float a0 = slab(n0.x, idx, ox), b0 = slab(n0.w, idx, ox);
      float a1 = slab(n0.y, idy, oy), b1 = slab(n1.x, idy, oy);
      float a2 = slab(n0.z, idz, oz), b2 = slab(n1.y, idz, oz);
      float tn0 = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fmaxf(fminf(a2, b2), ray.tmin));
      float tf0 = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fminf(fmaxf(a2, b2), ray.tmax));
      float c0 = slab(n1.z, idx, ox), d0 = slab(n2.y, idx, ox);
      float c1 = slab(n1.w, idy, oy), d1 = slab(n2.z, idy, oy);
      float c2 = slab(n2.x, idz, oz), d2 = slab(n2.w, idz, oz);
      float tn1 = fmaxf(fmaxf(fminf(c0, d0), fminf(c1, d1)), fmaxf(fminf(c2, d2), ray.tmin));
      float tf1 = fminf(fminf(fmaxf(c0, d0), fmaxf(c1, d1)), fminf(fmaxf(c2, d2), ray.tmax));
      // generous slack on both ends (padding + this keep the test conservative)
      bool h0 = tn0 * 0.9999995f <= tf0 * 1.0000005f, h1 = tn1 * 0.9999995f <= tf1 * 1.0000005f;
      int  ch0 = __float_as_int(n3.x), ch1 = __float_as_int(n3.y);
      if (h0 && h1) {
        bool swap = tn1 < tn0;
        st.push(swap ? ch0 : ch1);
        cur = swap ? ch1 : ch0;
      } else if (h0) {
        cur = ch0;
      } else if (h1) {
        cur = ch1;
      } else {
        cur = st.sp == base ? kTraverseDone : st.pop();
      }
    }
    if (cur == kTraverseDone) break;
    {
      int code = ~cur, first = code >> 3, count = (code & 7) + 1;
      if (is_lines) {
        for (int k = 0; k < count; k++) {
          const float4 a = __ldg(prims + 2 * (size_t)(first + k)), b = __ldg(prims + 2 * (size_t)(first + k) + 1);
          YH_STAT(if (prims_tested)(*prims_tested)++;)
          if (line_certain_miss(ray.o, ray.d, uu, mo, a, b)) continue;
          vec2f uvc;
          float tc;
          if (intersect_line(ray, {a.x, a.y, a.z}, {b.x, b.y, b.z}, a.w, b.w, uvc, tc)) {
            if (hit && tc == ray.tmax &&
                !reference_visits_later(sh, __ldg(sh.ref_elem + ref), __ldg(sh.ref_elem + first + k), dsign))
              continue;  // exact tie that the reference resolves the other way
            hit = true, ref = first + k, uv = uvc, ray.tmax = tc;
          }
        }
      } else {
        for (int k = 0; k < count; k++) {
          const float4 a = __ldg(prims + 3 * (size_t)(first + k)), b = __ldg(prims + 3 * (size_t)(first + k) + 1),
                       c = __ldg(prims + 3 * (size_t)(first + k) + 2);
          YH_STAT(if (prims_tested)(*prims_tested)++;)
          vec2f uvc;
          float tc;
          if (intersect_triangle(ray, {a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}, uvc, tc)) {
            if (hit && tc == ray.tmax &&
                !reference_visits_later(sh, __ldg(sh.ref_elem + ref), __ldg(sh.ref_elem + first + k), dsign))
              continue;
            hit = true, ref = first + k, uv = uvc, ray.tmax = tc;
          }
        }
      }
      if (find_any && hit) {
        st.sp = base;
        return true;
      }
    }
    cur = st.sp == base ? kTraverseDone : st.pop();
  }
  return hit;
}

template <bool FUSED>
__device__ __noinline__ bool intersect_shape_far(const dev_shape& sh, ray3f& ray, trav_stack& st, int& ref, vec2f& uv,
    bool find_any, unsigned long long* nodes_visited, unsigned long long* prims_tested) {
  return intersect_shape_impl<FUSED>(sh, ray, st, ref, uv, find_any, nodes_visited, prims_tested);
}
__device__ __forceinline__ bool intersect_shape(const dev_shape& sh, ray3f& ray, trav_stack& st, int& ref,
    vec2f& uv, bool find_any, unsigned long long* nodes_visited, unsigned long long* prims_tested) {
  if (sh.num_refs == 0) return false;
  float mo = fmaxf(fmaxf(fabsf(ray.o.x), fabsf(ray.o.y)), fabsf(ray.o.z));
  if (!(mo <= 4.0f * sh.scale))  // far origin (or NaN): rare, kept out of line
    return intersect_shape_far<false>(sh, ray, st, ref, uv, find_any, nodes_visited, prims_tested);
  return intersect_shape_impl<true>(sh, ray, st, ref, uv, find_any, nodes_visited, prims_tested);
}

// conservative ray / box test for the object level
__device__ __forceinline__ bool hits_box(const ray3f& ray, const bbox3f& b) {
  float idx = 1.0f / ray.d.x, idy = 1.0f / ray.d.y, idz = 1.0f / ray.d.z;
  float a0 = (b.min.x - ray.o.x) * idx, b0 = (b.max.x - ray.o.x) * idx;
  float a1 = (b.min.y - ray.o.y) * idy, b1 = (b.max.y - ray.o.y) * idy;
  float a2 = (b.min.z - ray.o.z) * idz, b2 = (b.max.z - ray.o.z) * idz;
  float tn = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fmaxf(fminf(a2, b2), ray.tmin));
  float tf = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fminf(fmaxf(a2, b2), ray.tmax));
  return tn * 0.999999f <= tf * 1.000001f;
}

// intersect_instance_bvh, yocto_pathtrace.cpp:1031-1037
__device__ __forceinline__ bool intersect_instance(const dev_scene& sc, int object, const ray3f& ray,
    trav_stack& st, int& element, vec2f& uv, float& distance, bool find_any,
    unsigned long long* nv = nullptr, unsigned long long* pt = nullptr) {
  const dev_object& ob = sc.objects[object];
  const dev_shape&  sh = sc.shapes[ob.shape];
  ray3f             r  = {transform_point(ob.inv_frame, ray.o), transform_vector(ob.inv_frame, ray.d), ray.tmin, ray.tmax};
  int               ref = -1;
  if (!intersect_shape(sh, r, st, ref, uv, find_any, nv, pt)) return false;
  distance = r.tmax;
  element  = __ldg(sh.ref_elem + ref);
  return true;
}

// intersect_scene_bvh, yocto_pathtrace.cpp:934-1028
__device__ __forceinline__ hit_record intersect_scene(const dev_scene& sc, ray3f ray, trav_stack& st,
    bool find_any, unsigned long long* nv = nullptr, unsigned long long* pt = nullptr) {
  hit_record h;
  h.object = -1, h.element = -1, h.uv = {0, 0}, h.distance = 0, h.hit = false;
  for (int o = 0; o < sc.num_objects; o++) {
    if (!hits_box(ray, sc.objects[o].bbox)) continue;
    int   element;
    vec2f uv;
    float dist;
    if (intersect_instance(sc, o, ray, st, element, uv, dist, find_any, nv, pt)) {
      h.hit = true, h.object = o, h.element = element, h.uv = uv, h.distance = dist;
      ray.tmax = dist;
      if (find_any) return h;
    }
  }
  return h;
}


// --------------------------------------------------------------------------
// Persistent closest-hit tracer: dynamic ray fetch, a pure node loop, parked
// leaves.
//
// Traversal lengths of hair rays are wildly uneven (a ray that leaves the hair
// volume ends after a handful of nodes, one that grazes a lock visits hundreds),
// and so is the number of node steps between two leaves.  A warp that takes 32
// rays and alternates "all node steps / all leaf tests" runs its node loop at
// ~7 of 32 lanes (ncu, profiles/r01_*).  Here
//   * a warp keeps running and refills the lanes whose ray has finished from a
//     global cursor (finished lanes also write their hit records together);
//   * the hot loop is ONE node step per lane per iteration; a leaf child whose
//     box is hit is not tested on the spot but parked in a small per-lane list;
//   * when enough lanes hold parked leaves (or too few can still walk), the warp
//     tests them together: the certain-miss test first, then, for the few
//     survivors, the reference's exact arithmetic.
// Primitive tests therefore run later, and in another order, than a plain
// depth-first walk would run them.  That is legal because the hit record does
// not depend on the order: the closest t wins and an exact tie goes to
// reference_visits_later(); a later test only means t_max shrinks later, i.e.
// some extra node steps.
//
// IO supplies rays and consumes hit records:
//   bool load(int index, int& handle, ray3f& world_ray)
//   void reload(int handle, vec3f& o, vec3f& d)        (world origin / direction)
//   void store(int handle, const hit_record&)
constexpr int kRefillIdle = 8;   // refill once this many lanes of the warp have no ray
constexpr int kParkSlots  = 4;   // parked leaves per lane
constexpr int kParkLanes  = 12;  // test parked leaves once this many lanes hold some
constexpr int kWalkLanes  = 12;  // ... or once fewer lanes than this can still take a node step

// Per-lane state that the node loop does not touch lives in shared memory
// (word w of thread t at cold[w * kTraceBlock + t]: conflict-free), so the hot
// loop fits in few registers and more warps are resident to hide the latency
// of the dependent node fetches.
enum cold_word {
  CW_OX, CW_OY, CW_OZ, CW_DX, CW_DY, CW_DZ,   // object-space ray
  CW_HANDLE, CW_OBJ, CW_REF, CW_U, CW_V,      // ray handle, object index, best reference of this shape, its uv
  CW_BOBJ, CW_BELEM, CW_BU, CW_BV, CW_BDIST,  // best hit over the objects closed so far
  CW_FLAGS,                                    // bit 0 shape_hit, 1 best.hit, 2 is_lines, 4..6 dsign
  CW_PRIMS_LO, CW_PRIMS_HI, CW_SHAPE_LO, CW_SHAPE_HI,
  CW_COUNT
};
constexpr int kColdWords = CW_COUNT;

template <class IO>
__device__ __forceinline__ void trace_persistent(const dev_scene& sc, IO& io, int n, int* cursor, int* smem_stack,
    int* smem_park, int* smem_cold, bool find_any, unsigned long long* nodes_visited,
    unsigned long long* prims_tested) {
  const unsigned full = 0xffffffffu;
  const int      lane = threadIdx.x & 31;
  trav_stack     st;
  st.smem = smem_stack;
  st.sp   = 0;
  volatile int* cold = smem_cold;
  auto cw  = [&](int w) -> volatile int& { return cold[w * kTraceBlock]; };
  auto cwf = [&](int w) -> float { return __int_as_float(cold[w * kTraceBlock]); };
  auto put = [&](int w, float v) { cold[w * kTraceBlock] = __float_as_int(v); };
  auto ptr = [&](int w) -> const void* {
    return (const void*)(((unsigned long long)(unsigned)cold[(w + 1) * kTraceBlock] << 32) | (unsigned)cold[w * kTraceBlock]);
  };
  auto put_ptr = [&](int w, const void* p) {
    cold[w * kTraceBlock]       = (int)(unsigned)((unsigned long long)p);
    cold[(w + 1) * kTraceBlock] = (int)(unsigned)((unsigned long long)p >> 32);
  };
  // registers of the node loop
  bool  have = false, finished = false, exhausted = false;
  int   cur = kTraverseDone, npark = 0;
  float tmin = 0, tmax = 0;
  float idx = 0, idy = 0, idz = 0, ox = 0, oy = 0, oz = 0;
  const float4* __restrict__ nodes = nullptr;

  while (true) {
    // ---- 1. lanes without a ray: write finished records, fetch new rays
    unsigned idle = __ballot_sync(full, !have);
    if (__popc(idle) >= kRefillIdle || idle == full) {
      if (finished) {
        hit_record best;
        best.object = cw(CW_BOBJ), best.element = cw(CW_BELEM), best.uv = {cwf(CW_BU), cwf(CW_BV)};
        best.distance = cwf(CW_BDIST), best.hit = (cw(CW_FLAGS) & 2) != 0;
        io.store(cw(CW_HANDLE), best);
        finished = false;
      }
      if (!exhausted) {
        int cnt = __popc(idle), base = 0;
        if (lane == 0) base = atomicAdd(cursor, cnt);
        base = __shfl_sync(full, base, 0);
        if (!have) {
          int   index = base + __popc(idle & ((1u << lane) - 1));
          ray3f wr;
          int   handle;
          if (index < n && io.load(index, handle, wr)) {
            have = true, cur = kTraverseDone, st.sp = 0, npark = 0;
            tmin = wr.tmin, tmax = wr.tmax;
            cw(CW_HANDLE) = handle, cw(CW_OBJ) = -1, cw(CW_FLAGS) = 0;
            cw(CW_BOBJ) = -1, cw(CW_BELEM) = -1, put(CW_BU, 0), put(CW_BV, 0), put(CW_BDIST, 0);
          }
        }
        exhausted = base + cnt >= n;
        idle      = __ballot_sync(full, !have);
      }
      if (idle == full) break;  // nothing in flight and nothing left to fetch
    }

    // ---- 2. parked leaves
    {
      bool     can_walk = have && cur >= 0 && npark <= kParkSlots - 2;
      unsigned pm = __ballot_sync(full, npark > 0);
      unsigned wm = __ballot_sync(full, can_walk);
      if (npark > 0 && (__popc(pm) >= kParkLanes || __popc(wm) < kWalkLanes)) {
        vec3f o = {cwf(CW_OX), cwf(CW_OY), cwf(CW_OZ)}, d = {cwf(CW_DX), cwf(CW_DY), cwf(CW_DZ)};
        float uu = __fmaf_rn(d.x, d.x, __fmaf_rn(d.y, d.y, d.z * d.z));
        float mo = fabsf(o.x) + fabsf(o.y) + fabsf(o.z);
        int   flags = cw(CW_FLAGS), ref = cw(CW_REF);
        bool  shape_hit = flags & 1, is_lines = (flags & 4) != 0;
        unsigned dsign = (flags >> 4) & 7;
        vec2f uv = {cwf(CW_U), cwf(CW_V)};
        const float4* __restrict__ prims = (const float4*)ptr(CW_PRIMS_LO);
        const dev_shape* shp = (const dev_shape*)ptr(CW_SHAPE_LO);
        ray3f ray = {o, d, tmin, tmax};
        for (int e = 0; e < npark; e++) {
          int code = ~smem_park[e * kTraceBlock], first = code >> 3, count = (code & 7) + 1;
          for (int k = 0; k < count; k++) {
            int   r = first + k;
            vec2f uvc;
            float tc;
            bool  got;
            YH_STAT(if (prims_tested)(*prims_tested)++;)
            if (is_lines) {
              const float4 a = __ldg(prims + 2 * (size_t)r), b = __ldg(prims + 2 * (size_t)r + 1);
              if (line_certain_miss(o, d, uu, mo, a, b)) continue;
              got = intersect_line(ray, {a.x, a.y, a.z}, {b.x, b.y, b.z}, a.w, b.w, uvc, tc);
            } else {
              const float4 a = __ldg(prims + 3 * (size_t)r), b = __ldg(prims + 3 * (size_t)r + 1),
                           c = __ldg(prims + 3 * (size_t)r + 2);
              got = intersect_triangle(ray, {a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}, uvc, tc);
            }
            if (got && !(shape_hit && tc == ray.tmax &&
                           !reference_visits_later(*shp, __ldg(shp->ref_elem + ref), __ldg(shp->ref_elem + r), dsign)))
              shape_hit = true, ref = r, uv = uvc, ray.tmax = tc;
          }
        }
        tmax  = ray.tmax;
        npark = 0;
        if (shape_hit) {
          cw(CW_FLAGS) = flags | 1, cw(CW_REF) = ref, put(CW_U, uv.x), put(CW_V, uv.y);
          if (find_any) st.sp = 0, cur = kTraverseDone;
        }
      }
    }

    // ---- 3. lanes between shapes: close the finished one, open the next object
    if (have && cur == kTraverseDone && npark == 0) {
      int flags = cw(CW_FLAGS), obj = cw(CW_OBJ), handle = cw(CW_HANDLE);
      if (flags & 1) {
        const dev_shape* shp = (const dev_shape*)ptr(CW_SHAPE_LO);
        cw(CW_BOBJ) = obj, cw(CW_BELEM) = __ldg(shp->ref_elem + cw(CW_REF));
        cw(CW_BU) = cw(CW_U), cw(CW_BV) = cw(CW_V), put(CW_BDIST, tmax);
        flags = (flags & ~1) | 2;
        if (find_any) obj = sc.num_objects;
      }
      vec3f wo, wd;
      io.reload(handle, wo, wd);
      bool opened = false;
      while (++obj < sc.num_objects) {
        const dev_object& ob = sc.objects[obj];
        ray3f             wr = {wo, wd, tmin, tmax};
        if (!hits_box(wr, ob.bbox)) continue;
        const dev_shape& sh = sc.shapes[ob.shape];
        if (sh.num_refs == 0) continue;
        vec3f o = transform_point(ob.inv_frame, wo), d = transform_vector(ob.inv_frame, wd);
        float mo = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
        if (!(mo <= 4.0f * sh.scale)) {  // far origin: whole shape at once, out of line (rare)
          ray3f r = {o, d, tmin, tmax};
          int   rr = -1;
          vec2f ruv;
          if (intersect_shape_far<false>(sh, r, st, rr, ruv, find_any, nodes_visited, prims_tested)) {
            cw(CW_BOBJ) = obj, cw(CW_BELEM) = __ldg(sh.ref_elem + rr), put(CW_BU, ruv.x), put(CW_BV, ruv.y);
            put(CW_BDIST, r.tmax);
            tmax = r.tmax, flags |= 2;
            if (find_any) break;
          }
          continue;
        }
        auto safe = [](float v) { return fabsf(v) < 1e-20f ? copysignf(1e-20f, v) : v; };
        unsigned dsign = ((1.0f / d.x) < 0 ? 1u : 0u) | ((1.0f / d.y) < 0 ? 2u : 0u) | ((1.0f / d.z) < 0 ? 4u : 0u);
        idx = 1.0f / safe(d.x), idy = 1.0f / safe(d.y), idz = 1.0f / safe(d.z);
        ox = -o.x * idx, oy = -o.y * idy, oz = -o.z * idz;
        put(CW_OX, o.x), put(CW_OY, o.y), put(CW_OZ, o.z), put(CW_DX, d.x), put(CW_DY, d.y), put(CW_DZ, d.z);
        put_ptr(CW_PRIMS_LO, sh.prims), put_ptr(CW_SHAPE_LO, &sh);
        flags = (flags & 3) | (sh.is_lines ? 4 : 0) | (int)(dsign << 4);
        nodes = sh.nodes;
        cur   = sh.root;
        cw(CW_REF) = -1;
        if (cur < 0) {  // the whole shape is one leaf
          smem_park[0] = cur, npark = 1, cur = kTraverseDone;
        }
        opened = true;
        break;
      }
      cw(CW_FLAGS) = flags, cw(CW_OBJ) = obj;
      if (!opened) have = false, finished = true;
    }

    // ---- 4. one node step; leaf children whose box is hit are parked
    if (cur >= 0 && npark <= kParkSlots - 2) {
      const float4 n0 = __ldg(nodes + 4 * (size_t)cur), n1 = __ldg(nodes + 4 * (size_t)cur + 1),
                   n2 = __ldg(nodes + 4 * (size_t)cur + 2), n3 = __ldg(nodes + 4 * (size_t)cur + 3);
      YH_STAT(if (nodes_visited)(*nodes_visited)++;)
      float a0 = __fmaf_rn(n0.x, idx, ox), b0 = __fmaf_rn(n0.w, idx, ox);
      float a1 = __fmaf_rn(n0.y, idy, oy), b1 = __fmaf_rn(n1.x, idy, oy);
      float a2 = __fmaf_rn(n0.z, idz, oz), b2 = __fmaf_rn(n1.y, idz, oz);
      float tn0 = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fmaxf(fminf(a2, b2), tmin));
      float tf0 = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fminf(fmaxf(a2, b2), tmax));
      float c0 = __fmaf_rn(n1.z, idx, ox), d0 = __fmaf_rn(n2.y, idx, ox);
      float c1 = __fmaf_rn(n1.w, idy, oy), d1 = __fmaf_rn(n2.z, idy, oy);
      float c2 = __fmaf_rn(n2.x, idz, oz), d2 = __fmaf_rn(n2.w, idz, oz);
      float tn1 = fmaxf(fmaxf(fminf(c0, d0), fminf(c1, d1)), fmaxf(fminf(c2, d2), tmin));
      float tf1 = fminf(fminf(fmaxf(c0, d0), fmaxf(c1, d1)), fminf(fmaxf(c2, d2), tmax));
      bool h0 = tn0 * 0.9999995f <= tf0 * 1.0000005f, h1 = tn1 * 0.9999995f <= tf1 * 1.0000005f;
      int  ch0 = __float_as_int(n3.x), ch1 = __float_as_int(n3.y);
      if (h0 && ch0 < 0) smem_park[npark++ * kTraceBlock] = ch0, h0 = false;
      if (h1 && ch1 < 0) smem_park[npark++ * kTraceBlock] = ch1, h1 = false;
      if (h0 && h1) {
        bool swap = tn1 < tn0;
        st.push(swap ? ch0 : ch1);
        cur = swap ? ch1 : ch0;
      } else if (h0) {
        cur = ch0;
      } else if (h1) {
        cur = ch1;
      } else {
        cur = st.sp == 0 ? kTraverseDone : st.pop();
      }
    }
  }
}

}  // namespace yh

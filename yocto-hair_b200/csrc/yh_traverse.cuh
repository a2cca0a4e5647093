// Closest-hit traversal of the two-level scene: objects (the reference's own
// top-level tree, so objects are visited in its order) -> shape LBVH
// (yh_scene.cuh layout).  Restates intersect_scene_bvh /
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
constexpr int kStackShared = YH_STACK_SHARED;    // entries per thread kept in shared memory
constexpr int kStackLocal  = 96 - kStackShared;  // overflow (LBVH chains can be deep)
constexpr int kTraceBlock  = 128;

// One 32-byte slot of the packed tree, fetched with ONE 256-bit load
// (LDG.E.256.CONSTANT on sm_100a; the data is read-only for the kernel).
struct slot_words {
  unsigned w0, w1, w2, w3, w4, w5, w6, w7;
};
__device__ __forceinline__ slot_words ld_slot(const uint4* __restrict__ slots, int slot) {
  slot_words s;
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(s.w0), "=r"(s.w1), "=r"(s.w2), "=r"(s.w3), "=r"(s.w4), "=r"(s.w5), "=r"(s.w6), "=r"(s.w7)
      : "l"(slots + 2 * (size_t)slot));
  return s;
}

// leaf flags of an inner slot (w7): sign-bit tests are one instruction each
__device__ __forceinline__ bool slot_leaf0(unsigned w7) { return (int)w7 < 0; }
__device__ __forceinline__ bool slot_leaf1(unsigned w7) { return (int)(w7 << 1) < 0; }

struct hit_record {
  int   object, element;
  vec2f uv;
  float distance;
  bool  hit;
};

// Per-thread traversal stack: the first kStackShared entries are interleaved
// in shared memory (entry e of thread t at smem[e * blockDim.x + t], so a warp
// touches 32 consecutive banks), deeper entries spill to local memory.
// The stack pointer is kept as the shared-memory ADDRESS of the next free entry (`top`), so a
// push is one st.shared + one add and a pop one add + one ld.shared; nearly every node step of
// a warp has some lane pushing and some lane popping, so these few instructions are paid by
// the whole warp every step.  The overflow array is not a member and is touched only through
// out-of-line helpers: a dynamically indexed member would put the whole struct into local
// memory (seen in the SASS of the first version), and an inlined overflow path costs
// predicated-off issue slots in the hot loop.
constexpr unsigned kStackStride = kTraceBlock * 4u;  // bytes between two entries of one thread
// (extern "C": node_control() calls them by name from PTX)
extern "C" __device__ __noinline__ void yh_stack_overflow_store(int* local, unsigned index, int v) { local[index] = v; }
extern "C" __device__ __noinline__ int  yh_stack_overflow_load(const int* local, unsigned index) { return local[index]; }
struct trav_stack {
  unsigned base;   // shared-state-space byte address of this thread's entry 0
  unsigned top;    // ... of its next free entry
  unsigned limit;  // ... of the first entry that no longer fits in shared memory
  int*     local;  // kStackLocal ints of the caller's frame
  __device__ __forceinline__ void init(int* shared_entry0, int* local_array) {
    base  = (unsigned)__cvta_generic_to_shared(shared_entry0);
    top   = base;
    limit = base + kStackShared * kStackStride;
    local = local_array;
  }
  // `base` is the floor of the shape walk in progress; the entries between floor() and base belong to the walk over
  // the top-level tree (scene nodes and objects still to visit), which only runs while no shape is open
  __device__ __forceinline__ unsigned floor() const { return limit - kStackShared * kStackStride; }
  __device__ __forceinline__ bool     empty() const { return top == base; }
  __device__ __forceinline__ void     clear() { top = base; }
  __device__ __forceinline__ unsigned mark() const { return top; }
  __device__ __forceinline__ void push(int v) {
    if (top < limit) asm volatile("st.shared.s32 [%0], %1;" ::"r"(top), "r"(v) : "memory");
    else yh_stack_overflow_store(local, (top - limit) / kStackStride, v);
    top += kStackStride;
  }
  __device__ __forceinline__ int pop() {
    top -= kStackStride;
    int v;
    if (top < limit) asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(top) : "memory");
    else v = yh_stack_overflow_load(local, (top - limit) / kStackStride);
    return v;
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
__device__ __forceinline__ bool reference_visits_later(const dev_shape& sh, int ea, int eb, unsigned dsign) {
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

// The primitive in leaf slot `slot` against the ray: certain-miss test, then
// the reference's exact test.
// kind: 1 lines, 2 points, 0 triangles
__device__ __forceinline__ bool test_leaf_slot(const dev_shape& sh, int kind, const ray3f& ray, float uu, float mo,
    int slot, vec2f& uv, float& t) {
  const slot_words s = ld_slot(sh.slots, slot);
  if (kind == 2)
    return intersect_point(ray, {__uint_as_float(s.w0), __uint_as_float(s.w1), __uint_as_float(s.w2)}, __uint_as_float(s.w3), uv, t);
  if (kind == 1) {
    const float4 a = {__uint_as_float(s.w0), __uint_as_float(s.w1), __uint_as_float(s.w2), __uint_as_float(s.w3)};
    const float4 b = {__uint_as_float(s.w4), __uint_as_float(s.w5), __uint_as_float(s.w6), __uint_as_float(s.w7)};
    if (line_certain_miss(ray.o, ray.d, uu, mo, a, b)) return false;
    return intersect_line(ray, {a.x, a.y, a.z}, {b.x, b.y, b.z}, a.w, b.w, uv, t);
  }
  const float4* tri = sh.tris + 3 * (size_t)s.w0;
  const float4  a = __ldg(tri), b = __ldg(tri + 1), c = __ldg(tri + 2);
  return intersect_triangle(ray, {a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}, uv, t);
}

// Folds a leaf hit into the running closest hit of this shape: the smaller t
// wins (the exact test already rejected t > tmax), an exact tie goes to the
// element the reference visits last.
__device__ __forceinline__ void accept_hit(const dev_shape& sh, unsigned dsign, int slot, const vec2f& uvc, float tc,
    bool& hit, int& ref, vec2f& uv, float& tmax) {
  if (hit && tc == tmax &&
      !reference_visits_later(sh, __ldg(sh.slot_elem + ref), __ldg(sh.slot_elem + slot), dsign))
    return;
  hit = true, ref = slot, uv = uvc, tmax = tc;
}

// Box tests.  Boxes live on the shape's 16-bit grid, plane = grid_origin + q *
// grid_step, so a slab distance is ONE FMA per plane,
//     t = q * (step / d) + (grid_origin - o) / d,
// with both factors per-ray constants.  Their rounding displaces a plane by
// <~ 1e-6 * scale in space as long as |o| <= ~2 * scale (scale = largest
// |coordinate| of the shape); the builder pads every box by 2e-6 * scale before
// quantising outwards, which covers it.  A ray that starts farther away is
// first moved along itself to its point nearest to the object-space origin
// (t0 >= 0): if even that point is more than 2 * scale away the ray cannot
// touch the shape, otherwise box distances are measured from there and compared
// with [tmin - t0, tmax - t0].  Primitive tests always see the original ray.
struct slab_consts {
  float    ax, ay, az, bx, by, bz;
  unsigned nx, ny, nz, fx, fy, fz;  // PRMT selectors of the entry (near) and exit (far) plane of each axis
};
// Decoding a plane costs ONE instruction before the FMA and none on the conversion pipe: a box word
// holds the lo and hi grid coordinate of one axis (lo | hi << 16); PRMT picks the half this ray
// enters (or leaves) through - the selector depends only on the sign of the direction - and
// drops it into the mantissa of the float 2^23 + q, exactly.  The 2^23 is folded into the slab
// offset, b' = b - 2^23 a, so  t = fma(2^23 + q, a, b').  The rounding of b' (half an ulp of
// ~2^23 |a|, i.e. about half a grid step) moves ALL planes of that axis by the same amount for
// this ray; the builder pads every box by that much (yh_scene.cu, fold_pad_steps).
constexpr unsigned kSelLo = 0x7610u, kSelHi = 0x7632u;
__device__ __forceinline__ float plane_float(unsigned w, unsigned sel) {
  unsigned m;  // prmt directly: __byte_perm() masks the selector first (one LOP3 per plane)
  asm("prmt.b32 %0, %1, 0x4b000000, %2;" : "=r"(m) : "r"(w), "r"(sel));
  return __uint_as_float(m);
}
struct shape_ray {
  slab_consts k;
  float       t0;     // box distances are measured from o + t0 * d
  bool        valid;  // false: the ray cannot touch the shape
};
// A zero direction component is replaced by +-1e-20 (1/0 = inf would turn the
// FMA into inf - inf); the slab then still rejects origins outside it and
// passes origins inside, as the reference's intersect_bbox does.
__device__ __forceinline__ shape_ray make_shape_ray(const dev_shape& sh, const vec3f& o, const vec3f& d) {
  shape_ray r;
  r.t0    = 0;
  r.valid = true;
  vec3f bo = o;
  float lim = 2.0f * sh.scale;
  if (!(fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z)) <= lim)) {
    float uu = __fmaf_rn(d.x, d.x, __fmaf_rn(d.y, d.y, d.z * d.z));
    r.t0     = fmaxf(0.0f, -(o.x * d.x + o.y * d.y + o.z * d.z) / fmaxf(uu, 1e-30f));
    bo       = {__fmaf_rn(d.x, r.t0, o.x), __fmaf_rn(d.y, r.t0, o.y), __fmaf_rn(d.z, r.t0, o.z)};
    // nearest point still outside the ball of radius 2 * scale (> sqrt(3) * scale): no contact
    r.valid = bo.x * bo.x + bo.y * bo.y + bo.z * bo.z <= 1.0001f * lim * lim;
  }
  auto  safe = [](float v) { return fabsf(v) < 1e-20f ? copysignf(1e-20f, v) : v; };
  float idx = 1.0f / safe(d.x), idy = 1.0f / safe(d.y), idz = 1.0f / safe(d.z);
  const float ax = sh.grid_step.x * idx, ay = sh.grid_step.y * idy, az = sh.grid_step.z * idz;
  const float bx = __fmaf_rn(sh.grid_origin.x, idx, -bo.x * idx), by = __fmaf_rn(sh.grid_origin.y, idy, -bo.y * idy),
              bz = __fmaf_rn(sh.grid_origin.z, idz, -bo.z * idz);
  r.k = {ax, ay, az, __fmaf_rn(-8388608.0f, ax, bx), __fmaf_rn(-8388608.0f, ay, by), __fmaf_rn(-8388608.0f, az, bz),
      idx < 0 ? kSelHi : kSelLo, idy < 0 ? kSelHi : kSelLo, idz < 0 ? kSelHi : kSelLo,
      idx < 0 ? kSelLo : kSelHi, idy < 0 ? kSelLo : kSelHi, idz < 0 ? kSelLo : kSelHi};
  return r;
}
// [tmin, tmax] seen from the shifted origin, widened by the rounding of t0
__device__ __forceinline__ float shifted_tmin(float tmin, float t0) { return tmin - t0 - 4e-6f * t0; }
__device__ __forceinline__ float shifted_tmax(float tmax, float t0) { return tmax - t0 + 4e-6f * t0; }

// One child box (three words, one per axis: lo | hi << 16) against [tlo, thi]: entry distance and
// far limit; the box is hit iff tn <= lim.  lim carries a relative slack of 5e-7 on both ends,
// outward-going also for the negative distances of a shifted origin:
// tn - tf <= 1e-6 * max(|tn|, |tf|) is implied by tn - 5e-7 |tn| <= tf + 5e-7 |tf|.
__device__ __forceinline__ void child_dist(unsigned wx, unsigned wy, unsigned wz, const slab_consts& k, float tlo,
    float thi, float& tn, float& lim) {
  const float nx = __fmaf_rn(plane_float(wx, k.nx), k.ax, k.bx), fx = __fmaf_rn(plane_float(wx, k.fx), k.ax, k.bx);
  const float ny = __fmaf_rn(plane_float(wy, k.ny), k.ay, k.by), fy = __fmaf_rn(plane_float(wy, k.fy), k.ay, k.by);
  const float nz = __fmaf_rn(plane_float(wz, k.nz), k.az, k.bz), fz = __fmaf_rn(plane_float(wz, k.fz), k.az, k.bz);
  tn             = fmaxf(fmaxf(nx, ny), fmaxf(nz, tlo));
  const float tf = fminf(fminf(fx, fy), fminf(fz, thi));
  lim            = __fmaf_rn(fmaxf(fabsf(tn), fabsf(tf)), 1e-6f, tf);
}
__device__ __forceinline__ void child_dists(const slot_words& s, const slab_consts& k, float tlo, float thi, float& tn0,
    float& lim0, float& tn1, float& lim1) {
  child_dist(s.w0, s.w1, s.w2, k, tlo, thi, tn0, lim0);
  child_dist(s.w3, s.w4, s.w5, k, tlo, thi, tn1, lim1);
}
// both child boxes of an inner slot against [tlo, thi]
__device__ __forceinline__ void child_boxes(const slot_words& s, const slab_consts& k, float tlo, float thi,
    bool& h0, bool& h1, float& tn0, float& tn1) {
  float lim0, lim1;
  child_dists(s, k, tlo, thi, tn0, lim0, tn1, lim1);
  h0 = tn0 <= lim0, h1 = tn1 <= lim1;
}

// What follows the two box tests of a node step, written in PTX so that it stays predicated:
// nearly every step of a warp has SOME lane parking a leaf, some lane pushing and some lane
// popping, so whatever these paths cost is paid by the whole warp on every step (the compiler's
// version of the same logic, with its branches and select chains, was 56 instructions; this is
// about 40).  In C++ terms:
//     h0 = tn0 <= lim0, h1 = tn1 <= lim1; leaf0 = w7 bit 31, leaf1 = w7 bit 30; ch0 = w6, ch1 = w6 + 1
//     if (h0 && leaf0) *park_top++ = ch0;   if (h1 && leaf1) *park_top++ = ch1;
//     in0 = h0 && !leaf0, in1 = h1 && !leaf1; second = in1 && (!in0 || tn1 < tn0)
//     if (in0 && in1) push(second ? ch0 : ch1)
//     cur = (in0 || in1) ? (second ? ch1 : ch0) : (stack empty ? kTraverseDone : pop())
// park_top and the stack top are shared-memory addresses, entries kStackStride bytes apart; stack
// entries beyond `limit` live in the local overflow array (rare: generic stores / loads behind a branch).
__device__ __forceinline__ void node_control(float tn0, float lim0, float tn1, float lim1, unsigned w6, unsigned w7,
    int& cur, unsigned& park_top, trav_stack& st) {
  asm volatile(
      "{\n\t"
      ".reg .pred h0, h1, l0, l1, pa, pb, i0, i1, both, any, sec, fit, q, slow;\n\t"
      ".reg .s32 ch1, nr, fr;\n\t"
      ".reg .u32 t;\n\t"
      "setp.le.f32 h0, %5, %6;\n\t"
      "setp.le.f32 h1, %7, %8;\n\t"
      "setp.lt.s32 l0, %10, 0;\n\t"
      "and.b32 t, %10, 0x40000000;\n\t"
      "setp.ne.u32 l1, t, 0;\n\t"
      "add.s32 ch1, %9, 1;\n\t"
      "and.pred pa, h0, l0;\n\t"
      "and.pred pb, h1, l1;\n\t"
      "and.pred i0, h0, !l0;\n\t"
      "and.pred i1, h1, !l1;\n\t"
      "@pa st.shared.s32 [%1], %9;\n\t"
      "@pa add.u32 %1, %1, %12;\n\t"
      "@pb st.shared.s32 [%1], ch1;\n\t"
      "@pb add.u32 %1, %1, %12;\n\t"
      "setp.lt.f32 sec, %7, %5;\n\t"
      "or.pred sec, sec, !i0;\n\t"
      "and.pred sec, sec, i1;\n\t"
      "selp.s32 nr, ch1, %9, sec;\n\t"
      "selp.s32 fr, %9, ch1, sec;\n\t"
      "and.pred both, i0, i1;\n\t"
      "or.pred any, i0, i1;\n\t"
      // push the farther child
      "setp.lt.u32 fit, %2, %4;\n\t"
      "and.pred q, both, fit;\n\t"
      "and.pred slow, both, !fit;\n\t"
      "@q st.shared.s32 [%2], fr;\n\t"
      "@!slow bra PUSHED;\n\t"
      "{\n\t"
      ".param .b64 p0;\n\t"
      ".param .b32 p1;\n\t"
      ".param .b32 p2;\n\t"
      "sub.u32 t, %2, %4;\n\t"
      "shr.u32 t, t, 9;\n\t"
      "st.param.b64 [p0], %11;\n\t"
      "st.param.b32 [p1], t;\n\t"
      "st.param.b32 [p2], fr;\n\t"
      "call yh_stack_overflow_store, (p0, p1, p2);\n\t"
      "}\n\t"
      "PUSHED:\n\t"
      "@both add.u32 %2, %2, %12;\n\t"
      // next node: the nearer child, else the top of the stack, else done
      "mov.s32 %0, nr;\n\t"
      "@!any mov.s32 %0, 0x80000000;\n\t"
      "setp.ne.u32 q, %2, %3;\n\t"
      "and.pred q, q, !any;\n\t"
      "@q sub.u32 %2, %2, %12;\n\t"
      "setp.lt.u32 fit, %2, %4;\n\t"
      "and.pred slow, q, !fit;\n\t"
      "and.pred q, q, fit;\n\t"
      "@q ld.shared.s32 %0, [%2];\n\t"
      "@!slow bra POPPED;\n\t"
      "{\n\t"
      ".param .b64 p0;\n\t"
      ".param .b32 p1;\n\t"
      ".param .b32 r0;\n\t"
      "sub.u32 t, %2, %4;\n\t"
      "shr.u32 t, t, 9;\n\t"
      "st.param.b64 [p0], %11;\n\t"
      "st.param.b32 [p1], t;\n\t"
      "call (r0), yh_stack_overflow_load, (p0, p1);\n\t"
      "ld.param.b32 %0, [r0];\n\t"
      "}\n\t"
      "POPPED:\n\t"
      "}"
      : "+r"(cur), "+r"(park_top), "+r"(st.top)
      : "r"(st.base), "r"(st.limit), "f"(tn0), "f"(lim0), "f"(tn1), "f"(lim1), "r"(w6), "r"(w7), "l"(st.local),
        "n"(kStackStride)
      : "memory");
}

__device__ __forceinline__ unsigned ray_dsign(const vec3f& d) {  // ray_dsign, :862-863
  return ((1.0f / d.x) < 0 ? 1u : 0u) | ((1.0f / d.y) < 0 ? 2u : 0u) | ((1.0f / d.z) < 0 ? 4u : 0u);
}

constexpr int kTraverseDone = (int)0x80000000;

// Traverses one shape with an object-space ray (depth first, leaves tested on
// the spot).  On a hit updates ray.tmax, ref (leaf slot), uv, and returns true.
// Used for the light-pdf rays against one object (sample_lights_pdf); the path
// rays go through trace_persistent() below.
__device__ __forceinline__ bool intersect_shape(const dev_shape& sh, ray3f& ray, trav_stack& st, int& ref, vec2f& uv,
    bool find_any, unsigned long long* nodes_visited, unsigned long long* prims_tested) {
  if (sh.num_refs == 0) return false;
  bool     hit      = false;
  int      kind     = sh.is_lines ? 1 : (sh.is_points ? 2 : 0);
  unsigned dsign    = ray_dsign(ray.d);
  float    uu = __fmaf_rn(ray.d.x, ray.d.x, __fmaf_rn(ray.d.y, ray.d.y, ray.d.z * ray.d.z));
  float    mo = fabsf(ray.o.x) + fabsf(ray.o.y) + fabsf(ray.o.z);
  auto leaf = [&](int slot) {
    vec2f uvc;
    float tc;
    YH_STAT(if (prims_tested)(*prims_tested)++;)
    if (test_leaf_slot(sh, kind, ray, uu, mo, slot, uvc, tc)) accept_hit(sh, dsign, slot, uvc, tc, hit, ref, uv, ray.tmax);
  };
  if (sh.root_leaf) {
    leaf(0);
    return hit;
  }
  if (sh.tiny_count) {  // a handful of primitives (the two triangles of a light quad): the boxes only cull, the exact
    for (int i = 0; i < sh.tiny_count; i++) {  // tests decide, and the closest hit does not depend on their order
      leaf(sh.tiny_slots[i]);
      if (find_any && hit) return true;
    }
    return hit;
  }
  shape_ray sr = make_shape_ray(sh, ray.o, ray.d);
  if (!sr.valid) return false;
  const unsigned base = st.mark();
  int            cur  = 0;
  while (true) {
    const slot_words s = ld_slot(sh.slots, cur);
    YH_STAT(if (nodes_visited)(*nodes_visited)++;)
    bool  h0, h1;
    float tn0, tn1;
    child_boxes(s, sr.k, shifted_tmin(ray.tmin, sr.t0), shifted_tmax(ray.tmax, sr.t0), h0, h1, tn0, tn1);
    int ch0 = (int)s.w6, ch1 = (int)s.w6 + 1;
    if (h0 && slot_leaf0(s.w7)) leaf(ch0), h0 = false;
    if (h1 && slot_leaf1(s.w7)) leaf(ch1), h1 = false;
    if (find_any && hit) {
      st.top = base;
      return true;
    }
    if (h0 && h1) {
      bool swap = tn1 < tn0;
      st.push(swap ? ch0 : ch1);
      cur = swap ? ch1 : ch0;
    } else if (h0) {
      cur = ch0;
    } else if (h1) {
      cur = ch1;
    } else {
      if (st.mark() == base) break;
      cur = st.pop();
    }
  }
  return hit;
}

// conservative ray / box test for the object level (inv = 1 / direction, formed once per ray)
__device__ __forceinline__ bool hits_box(const ray3f& ray, const vec3f& inv, const bbox3f& b) {
  float a0 = (b.min.x - ray.o.x) * inv.x, b0 = (b.max.x - ray.o.x) * inv.x;
  float a1 = (b.min.y - ray.o.y) * inv.y, b1 = (b.max.y - ray.o.y) * inv.y;
  float a2 = (b.min.z - ray.o.z) * inv.z, b2 = (b.max.z - ray.o.z) * inv.z;
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
  element  = __ldg(sh.slot_elem + ref);
  return true;
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
//   * the hot loop is ONE node step per lane per iteration: one 256-bit load,
//     two box tests; a leaf child whose box is hit is not tested on the spot
//     but parked in a small per-lane list;
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
#ifndef YH_REFILL_IDLE
#define YH_REFILL_IDLE 8
#endif
#ifndef YH_PARK_SLOTS
#define YH_PARK_SLOTS 4
#endif
#ifndef YH_PARK_LANES
#define YH_PARK_LANES 12
#endif
#ifndef YH_WALK_LANES
#define YH_WALK_LANES 12
#endif
#ifndef YH_OPEN_LANES
#define YH_OPEN_LANES 6
#endif
#ifndef YH_INNER_STEPS
#define YH_INNER_STEPS 8
#endif
#ifndef YH_INNER_LANES
#define YH_INNER_LANES 8
#endif
constexpr int kRefillIdle = YH_REFILL_IDLE;  // refill once this many lanes of the warp have no ray
constexpr int kParkSlots  = YH_PARK_SLOTS;   // parked leaves per lane
constexpr int kParkLanes  = YH_PARK_LANES;   // test parked leaves once this many lanes hold some
constexpr int kWalkLanes  = YH_WALK_LANES;   // ... or once fewer lanes than this can still take a node step
constexpr int kOpenLanes  = YH_OPEN_LANES;   // open / close shapes once this many lanes wait for it (or too few walk)
constexpr int kInnerSteps = YH_INNER_STEPS;  // node steps taken back to back while kInnerLanes lanes can walk
constexpr int kInnerLanes = YH_INNER_LANES;
#ifndef YH_VOTE_EVERY
#define YH_VOTE_EVERY 8
#endif
constexpr int kVoteEvery = YH_VOTE_EVERY;    // node steps between two votes of the inner loop

// Per-lane state that the node loop does not touch lives in shared memory
// (word w of thread t at cold[w * kTraceBlock + t]: conflict-free), so the hot
// loop fits in few registers and more warps are resident to hide the latency
// of the dependent node fetches.
enum cold_word {
  CW_OX, CW_OY, CW_OZ, CW_DX, CW_DY, CW_DZ,   // object-space ray
  CW_TMIN, CW_TMAX, CW_T0,                    // its range; box distances are measured from o + t0 d
  CW_HANDLE, CW_OBJ, CW_REF, CW_U, CW_V,      // ray handle, object index, best leaf slot of this shape, its uv
  CW_BOBJ, CW_BELEM, CW_BU, CW_BV, CW_BDIST,  // best hit over the objects closed so far
  CW_FLAGS,                                    // bit 0 shape_hit, 1 best.hit, 2 is_lines, 3 is_points, 4..6 dsign
  CW_SHAPE_LO, CW_SHAPE_HI,
  CW_COUNT
};
constexpr int kColdWords = CW_COUNT;

// Rays carried over from one launch to the next (wavefront only).  A launch ends with a tail: the queue is empty and
// a few long rays (hundreds of node steps, against ~70 on average) keep a few lanes of a few warps busy while the SMs
// drain - measured at a third of the traversal time of a 0.8 M-ray launch.  Instead, once a warp knows the queue is
// exhausted and has no more than kSuspendLanes rays left, it writes their complete walk state (cold words, node,
// parked leaves, stack) to a record and exits; the next launch of this group picks the records up FIRST, next to its
// fresh rays, so the long rays run inside the bulk of that launch.  A suspended path simply takes part in the
// shading one wavefront iteration later; nothing about its result changes.
#ifndef YH_SUSPEND_LANES
#define YH_SUSPEND_LANES 4
#endif
constexpr int kSuspendLanes = YH_SUSPEND_LANES;
// ... and only rays that have walked at least this many bursts of kInnerSteps node steps in this launch are handed over,
// so every launch moves every ray forward (a launch of a few rays would otherwise pass them on untouched)
#ifndef YH_SUSPEND_MIN_BURSTS
#define YH_SUSPEND_MIN_BURSTS 12
#endif
constexpr int kSuspendMinBursts = YH_SUSPEND_MIN_BURSTS;
constexpr int kSuspendEmpty = (int)0x80000001;  // SW_CUR of a record that holds no ray
constexpr int kSuspendWords = 128;  // cold words, cur, park count + 4 entries, stack count, floor count, <= 96 entries
struct suspend_io {
  const int* in_records;   // written by the previous launch
  int        n_in;
  int*       out_records;  // read by the next launch
  int*       out_count;
  int        capacity;
};
enum suspend_word { SW_CUR = CW_COUNT, SW_PARK_COUNT, SW_PARK0, SW_STACK_COUNT = SW_PARK0 + YH_PARK_SLOTS, SW_BASE_COUNT, SW_STACK0 };
static_assert(SW_STACK0 + 96 <= kSuspendWords, "a suspend record holds the whole stack");

// CAN_SUSPEND is a template parameter, not a run-time test: the mere presence of the hand-over code in the loop costs
// the traversal 5 % (instruction cache; measured on C3 with the buffer switched off), so launches that do not use it
// run an instantiation without it.
template <bool CAN_SUSPEND, class IO>
__device__ __forceinline__ void trace_persistent(const dev_scene& sc, IO& io, int n, int* cursor, int* smem_stack,
    int* smem_park, int* smem_cold, bool find_any, unsigned long long* nodes_visited,
    unsigned long long* prims_tested, const suspend_io* sus = nullptr) {
  const unsigned full = 0xffffffffu;
  const int      lane = threadIdx.x & 31;
  int            stack_overflow[kStackLocal];
  trav_stack     st;
  st.init(smem_stack, stack_overflow);
  volatile int* cold = smem_cold;
  auto cw  = [&](int w) -> volatile int& { return cold[w * kTraceBlock]; };
  auto cwf = [&](int w) -> float { return __int_as_float(cold[w * kTraceBlock]); };
  auto put = [&](int w, float v) { cold[w * kTraceBlock] = __float_as_int(v); };
  auto shape_ptr = [&]() -> const dev_shape* {
    return (const dev_shape*)(((unsigned long long)(unsigned)cold[CW_SHAPE_HI * kTraceBlock] << 32) |
                              (unsigned)cold[CW_SHAPE_LO * kTraceBlock]);
  };
  // registers of the node loop
  bool  have = false, finished = false, exhausted = false;
  int   cur = kTraverseDone;
  // parked leaves: shared-memory address of the next free entry of this lane's list (entries kStackStride apart)
  const unsigned park_base = (unsigned)__cvta_generic_to_shared(smem_park);
  const unsigned park_walk = park_base + (kParkSlots - 2) * kStackStride;  // a node step may park two more up to here
  unsigned       park_top  = park_base;
  float tlo = 0, thi = 0;  // [tmin, tmax] as seen from the (possibly shifted) box origin
  slab_consts k = {0, 0, 0, 0, 0, 0, kSelLo, kSelLo, kSelLo, kSelHi, kSelHi, kSelHi};
  const uint4* __restrict__ slots = nullptr;

  const int n_resume = CAN_SUSPEND ? sus->n_in : 0;  // work items [0, n_resume) are suspended rays, the rest the queue
  const int n_items  = n + n_resume;
  bool      keep_on  = false;  // the record buffer is full: finish in place
  int       bursts   = 0;      // node-step bursts this lane's ray has taken in this launch
  while (true) {
    // ---- 0. the queue is exhausted and few lanes still hold a ray: hand them to the next launch
    if (CAN_SUSPEND && exhausted && !keep_on) {
      const bool     ripe = have && bursts >= kSuspendMinBursts;
      const unsigned busy = __ballot_sync(full, have), out = __ballot_sync(full, ripe);
      const int      nb   = __popc(out);
      if (nb > 0 && __popc(busy) <= kSuspendLanes) {
        // Records are reserved with one atomic per warp and never given back (a roll-back would leave holes under
        // contention).  A reservation that does not fit marks the part of it that lies inside the buffer as empty -
        // the reader skips such records - and the warp finishes its rays in place.
        int first = 0;
        if (lane == 0) first = atomicAdd(sus->out_count, nb);
        first = __shfl_sync(full, first, 0);
        const int mine = first + __popc(out & ((1u << lane) - 1));
        if (first + nb > sus->capacity) {
          keep_on = true;
          if (ripe && mine < sus->capacity) sus->out_records[(size_t)mine * kSuspendWords + SW_CUR] = kSuspendEmpty;
        } else if (ripe) {
          int* rec = sus->out_records + (size_t)mine * kSuspendWords;
          for (int w = 0; w < CW_COUNT; w++) rec[w] = cold[w * kTraceBlock];
          rec[SW_CUR] = cur;
          const int pc = (int)((park_top - park_base) / kStackStride);
          rec[SW_PARK_COUNT] = pc;
          for (int i = 0; i < pc; i++) rec[SW_PARK0 + i] = smem_park[i * kTraceBlock];
          const int ns = (int)((st.top - st.floor()) / kStackStride);
          rec[SW_STACK_COUNT] = ns, rec[SW_BASE_COUNT] = (int)((st.base - st.floor()) / kStackStride);
          for (int e = 0; e < ns; e++) rec[SW_STACK0 + e] = e < kStackShared ? smem_stack[e * kTraceBlock] : stack_overflow[e - kStackShared];
          have = false;
        }
      }
    }
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
          if (CAN_SUSPEND && index < n_resume) {  // a ray suspended by the previous launch: restore its walk where it stopped
            const int* rec = sus->in_records + (size_t)index * kSuspendWords;
            const int  stored = __ldcs(rec + SW_CUR);
            if (stored != kSuspendEmpty) {  // (an empty record: a reservation that did not fit, see step 0)
              for (int w = 0; w < CW_COUNT; w++) cold[w * kTraceBlock] = __ldcs(rec + w);
              have = true, cur = stored, bursts = 0;
              const int pc = __ldcs(rec + SW_PARK_COUNT);
              for (int i = 0; i < pc; i++) smem_park[i * kTraceBlock] = __ldcs(rec + SW_PARK0 + i);
              park_top = park_base + pc * kStackStride;
              const int ns = __ldcs(rec + SW_STACK_COUNT);
              st.base = st.top = st.floor();
              for (int e = 0; e < ns; e++) st.push(__ldcs(rec + SW_STACK0 + e));
              st.base = st.floor() + __ldcs(rec + SW_BASE_COUNT) * kStackStride;
              if (cur >= 0 || pc > 0) {  // a shape is open: its per-ray constants follow from the stored ray
                const dev_shape& sh = *shape_ptr();
                shape_ray sr = make_shape_ray(sh, {cwf(CW_OX), cwf(CW_OY), cwf(CW_OZ)}, {cwf(CW_DX), cwf(CW_DY), cwf(CW_DZ)});
                k = sr.k, slots = sh.slots;
                tlo = shifted_tmin(cwf(CW_TMIN), cwf(CW_T0)), thi = shifted_tmax(cwf(CW_TMAX), cwf(CW_T0));
              }
            }
          } else if (index < n_items && io.load(index - n_resume, handle, wr)) {
            have = true, cur = kTraverseDone, park_top = park_base, bursts = 0;
            // the walk over the top-level tree starts at its root; a tree that is ONE leaf (<= 4 objects, the hair
            // scenes) is unrolled here: its objects in order, no box test of its own
            st.base = st.top = st.floor();
            if (sc.num_scene_nodes == 1) {
              const int count = __float_as_int(__ldg(&sc.scene_nodes[1].w)) >> 16;
              for (int k = count - 1; k >= 0; k--) st.push(~__ldg(sc.scene_prims + k));
            } else if (sc.num_scene_nodes > 1) {
              st.push(0);
            }
            st.base = st.top;
            put(CW_TMIN, wr.tmin), put(CW_TMAX, wr.tmax);
            cw(CW_HANDLE) = handle, cw(CW_OBJ) = -1, cw(CW_FLAGS) = 0;
            cw(CW_BOBJ) = -1, cw(CW_BELEM) = -1, put(CW_BU, 0), put(CW_BV, 0), put(CW_BDIST, 0);
          }
        }
        exhausted = base + cnt >= n_items;
        idle      = __ballot_sync(full, !have);
      }
      if (idle == full) break;  // nothing in flight and nothing left to fetch
    }

    // ---- 2. parked leaves
    {
      bool     can_walk = have && cur >= 0 && park_top <= park_walk;
      unsigned pm = __ballot_sync(full, park_top != park_base);
      unsigned wm = __ballot_sync(full, can_walk);
      if (park_top != park_base && (__popc(pm) >= kParkLanes || __popc(wm) < kWalkLanes)) {
        ray3f ray = {{cwf(CW_OX), cwf(CW_OY), cwf(CW_OZ)}, {cwf(CW_DX), cwf(CW_DY), cwf(CW_DZ)}, cwf(CW_TMIN),
            cwf(CW_TMAX)};
        float uu = __fmaf_rn(ray.d.x, ray.d.x, __fmaf_rn(ray.d.y, ray.d.y, ray.d.z * ray.d.z));
        float mo = fabsf(ray.o.x) + fabsf(ray.o.y) + fabsf(ray.o.z);
        int   flags = cw(CW_FLAGS), ref = cw(CW_REF);
        bool  shape_hit = flags & 1;
        int   kind = (flags & 4) ? 1 : ((flags & 8) ? 2 : 0);
        unsigned dsign = (flags >> 4) & 7;
        vec2f uv = {cwf(CW_U), cwf(CW_V)};
        const dev_shape& sh = *shape_ptr();
        for (unsigned pe = park_base; pe != park_top; pe += kStackStride) {
          int slot;
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(slot) : "r"(pe) : "memory");
          vec2f uvc;
          float tc;
          YH_STAT(if (prims_tested)(*prims_tested)++;)
          if (test_leaf_slot(sh, kind, ray, uu, mo, slot, uvc, tc))
            accept_hit(sh, dsign, slot, uvc, tc, shape_hit, ref, uv, ray.tmax);
        }
        park_top = park_base;
        if (shape_hit) {
          put(CW_TMAX, ray.tmax);
          thi = shifted_tmax(ray.tmax, cwf(CW_T0));
          cw(CW_FLAGS) = flags | 1, cw(CW_REF) = ref, put(CW_U, uv.x), put(CW_V, uv.y);
          if (find_any) st.clear(), cur = kTraverseDone;
        }
      }
    }

    // ---- 3. lanes between shapes: close the finished one, open the next object.  This block is
    // long and few lanes need it at a time, so it waits until kOpenLanes lanes do (or few can walk).
    bool between = have && cur == kTraverseDone && park_top == park_base;
    if (kOpenLanes > 1) {
      unsigned bm = __ballot_sync(full, between);
      unsigned wk = __ballot_sync(full, have && cur >= 0 && park_top <= park_walk);
      if (__popc(bm) < kOpenLanes && __popc(wk) >= kWalkLanes) between = false;
    }
    if (between) {
      int   flags = cw(CW_FLAGS), obj = cw(CW_OBJ), handle = cw(CW_HANDLE);
      float tmin = cwf(CW_TMIN), tmax = cwf(CW_TMAX);
      if (flags & 1) {
        cw(CW_BOBJ) = obj, cw(CW_BELEM) = __ldg(shape_ptr()->slot_elem + cw(CW_REF));
        cw(CW_BU) = cw(CW_U), cw(CW_BV) = cw(CW_V), put(CW_BDIST, tmax);
        flags = (flags & ~1) | 2;
        if (find_any) st.base = st.top = st.floor();  // any hit ends the walk
      }
      vec3f wo, wd;
      io.reload(handle, wo, wd);
      const vec3f winv = {1.0f / wd.x, 1.0f / wd.y, 1.0f / wd.z};
      bool opened = false;
      // intersect_scene_bvh, :960-1028: depth-first over the top-level tree, the child on the ray's near side of the
      // split axis first, the objects of a leaf in order.  Entries >= 0 are nodes, < 0 objects (~index).
      while (st.base != st.floor()) {
        st.top = st.base;
        const int entry = st.pop();
        st.base = st.top;
        ray3f wr = {wo, wd, tmin, tmax};
        if (entry >= 0) {
          const float4 lo = __ldg(&sc.scene_nodes[2 * entry]), hi = __ldg(&sc.scene_nodes[2 * entry + 1]);
          if (!hits_box(wr, winv, {{lo.x, lo.y, lo.z}, {hi.x, hi.y, hi.z}})) continue;
          const int a = __float_as_int(lo.w), b = __float_as_int(hi.w);
          if (b & 0x100) {
            const int  axis = b & 3;
            const float dinv = axis == 0 ? winv.x : (axis == 1 ? winv.y : winv.z);
            if (dinv < 0) st.push(a), st.push(a + 1);  // ray_dsign[node.axis], :983-989
            else st.push(a + 1), st.push(a);
          } else {
            for (int k = (b >> 16) - 1; k >= 0; k--) st.push(~__ldg(sc.scene_prims + a + k));
          }
          st.base = st.top;
          continue;
        }
        obj = ~entry;
        const dev_object& ob = sc.objects[obj];
        if (!hits_box(wr, winv, ob.bbox)) continue;
        const dev_shape& sh = sc.shapes[ob.shape];
        if (sh.num_refs == 0) continue;
        vec3f     o = transform_point(ob.inv_frame, wo), d = transform_vector(ob.inv_frame, wd);
        shape_ray sr = make_shape_ray(sh, o, d);
        if (!sr.valid) continue;
        k = sr.k;
        tlo = shifted_tmin(tmin, sr.t0), thi = shifted_tmax(tmax, sr.t0);
        put(CW_OX, o.x), put(CW_OY, o.y), put(CW_OZ, o.z), put(CW_DX, d.x), put(CW_DY, d.y), put(CW_DZ, d.z);
        put(CW_T0, sr.t0);
        cw(CW_SHAPE_LO) = (int)(unsigned)((unsigned long long)&sh);
        cw(CW_SHAPE_HI) = (int)(unsigned)((unsigned long long)&sh >> 32);
        flags = (flags & 3) | (sh.is_lines ? 4 : 0) | (sh.is_points ? 8 : 0) | (int)(ray_dsign(d) << 4);
        slots = sh.slots;
        cw(CW_REF) = -1;
        if (sh.root_leaf) smem_park[0] = 0, park_top = park_base + kStackStride, cur = kTraverseDone;  // the whole shape is one leaf
        else cur = 0;
        opened = true;
        break;
      }
      cw(CW_FLAGS) = flags, cw(CW_OBJ) = obj;
      if (!opened) have = false, finished = true;
    }

    // ---- 4. node steps; leaf children whose box is hit are parked.  Two steps between two votes:
    // the vote (and its branch) costs as much as a tenth of a step.
    auto node_step = [&]() {
      const slot_words s = ld_slot(slots, cur);
      YH_STAT(if (nodes_visited)(*nodes_visited)++;)
      float tn0, lim0, tn1, lim1;
      child_dists(s, k, tlo, thi, tn0, lim0, tn1, lim1);
      node_control(tn0, lim0, tn1, lim1, s.w6, s.w7, cur, park_top, st);
    };
#if YH_INNER_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
    bursts += (cur >= 0) ? 1 : 0;
    for (int inner = 0; inner < kInnerSteps; inner += kVoteEvery) {
      if (inner > 0 && __popc(__ballot_sync(full, cur >= 0 && park_top <= park_walk)) < kInnerLanes) break;
#pragma unroll
      for (int v = 0; v < kVoteEvery; v++)
        if (cur >= 0 && park_top <= park_walk) node_step();
    }
  }
}

}  // namespace yh

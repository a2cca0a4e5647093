// fp32 vector math of the hot path, operation order as in the reference's
// libs/yocto/yocto_math.h (cited per function).  Compiled with --fmad=false:
// the oracle is an x86-64 build without FMA, and bit-for-bit agreement of
// hit records and BSDF values needs the same roundings.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define YH_HD __host__ __device__ __forceinline__
// out of line on the device: the shading kernel would otherwise inline ~270 KB
// of code (8 copies of mp(), dozens of expf_/logf_) and stall on instruction
// fetch (ncu: stall_no_inst was its top stall reason)
#define YH_HD_CALL __host__ __device__ inline __noinline__
#else
#define YH_HD inline
#define YH_HD_CALL inline
#endif

namespace yh {

struct vec2f { float x, y; };
struct vec3f { float x, y, z; };
struct vec4f { float x, y, z, w; };
struct frame3f { vec3f x, y, z, o; };  // yocto_math.h:932-947
struct bbox3f { vec3f min, max; };

constexpr float pif     = 3.14159265358979323846f;  // yocto_math.h:214
constexpr float flt_max = 3.402823466e+38f;
constexpr float flt_eps = 1.192092896e-07f;
constexpr float ray_eps = 1e-4f;                    // yocto_math.h:1106

// scalar helpers, yocto_math.h:1778-1803.  min/max keep the reference's
// `(a < b) ? a : b` form (argument order matters for NaN).
YH_HD float fabs_(float a) { return a < 0 ? -a : a; }
YH_HD float fmin_(float a, float b) { return (a < b) ? a : b; }
YH_HD float fmax_(float a, float b) { return (a > b) ? a : b; }
YH_HD float clampf(float a, float lo, float hi) { return fmin_(fmax_(a, lo), hi); }
YH_HD int   clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }
YH_HD float sqrt_(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return __builtin_sqrtf(a);
#endif
}
YH_HD float div_(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
YH_HD float sqr(float a) { return a * a; }
YH_HD bool  isfinite_(float a) { return fabs_(a) <= flt_max; }  // false for NaN/Inf
YH_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
YH_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

// vec3f operators, yocto_math.h:1960-2120
YH_HD vec3f operator-(const vec3f& a) { return {-a.x, -a.y, -a.z}; }
YH_HD vec3f operator+(const vec3f& a, const vec3f& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
YH_HD vec3f operator+(const vec3f& a, float b) { return {a.x + b, a.y + b, a.z + b}; }
YH_HD vec3f operator-(const vec3f& a, const vec3f& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
YH_HD vec3f operator-(const vec3f& a, float b) { return {a.x - b, a.y - b, a.z - b}; }
YH_HD vec3f operator-(float a, const vec3f& b) { return {a - b.x, a - b.y, a - b.z}; }
YH_HD vec3f operator*(const vec3f& a, const vec3f& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
YH_HD vec3f operator*(const vec3f& a, float b) { return {a.x * b, a.y * b, a.z * b}; }
YH_HD vec3f operator*(float a, const vec3f& b) { return {a * b.x, a * b.y, a * b.z}; }
YH_HD vec3f operator/(const vec3f& a, const vec3f& b) { return {div_(a.x, b.x), div_(a.y, b.y), div_(a.z, b.z)}; }
YH_HD vec3f operator/(const vec3f& a, float b) { return {div_(a.x, b), div_(a.y, b), div_(a.z, b)}; }
YH_HD vec3f operator/(float a, const vec3f& b) { return {div_(a, b.x), div_(a, b.y), div_(a, b.z)}; }
YH_HD bool  operator==(const vec3f& a, const vec3f& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
YH_HD bool  nonzero(const vec3f& a) { return !(a.x == 0 && a.y == 0 && a.z == 0); }  // vec3f::operator bool, :1846

YH_HD float dot(const vec3f& a, const vec3f& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // :2029
YH_HD vec3f cross(const vec3f& a, const vec3f& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
YH_HD float length(const vec3f& a) { return sqrt_(dot(a, a)); }
YH_HD vec3f normalize(const vec3f& a) {  // :2037 (returns the input when its length is 0)
  float l = length(a);
  return (l != 0) ? a / l : a;
}
YH_HD vec3f orthonormalize(const vec3f& a, const vec3f& b) { return normalize(a - b * dot(a, b)); }  // :2056
YH_HD float max3(const vec3f& a) { return fmax_(fmax_(a.x, a.y), a.z); }  // :2092
// x / 2^k written as x * 2^-k: both are the correctly rounded value of the same real number, so every bit agrees for
// every x (zero, denormal, infinite included); the generic division is 8 instructions and a range check, this is one.
YH_HD float div_pow2(float x, float inv_pow2) { return x * inv_pow2; }
YH_HD float mean3(const vec3f& a) { return div_((a.x + a.y + a.z), 3.0f); }
YH_HD bool  isfinite3(const vec3f& a) { return isfinite_(a.x) && isfinite_(a.y) && isfinite_(a.z); }
YH_HD vec3f min3(const vec3f& a, const vec3f& b) { return {fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)}; }
YH_HD vec3f max3(const vec3f& a, const vec3f& b) { return {fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)}; }
YH_HD float luminance(const vec3f& a) {  // :3737
  return (0.2126f * a.x + 0.7152f * a.y + 0.0722f * a.z);
}

// frames, yocto_math.h:3136-3160, :2877-2903
YH_HD vec3f transform_point(const frame3f& a, const vec3f& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.o; }
YH_HD vec3f transform_vector(const frame3f& a, const vec3f& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
YH_HD vec3f transform_direction(const frame3f& a, const vec3f& b) { return normalize(transform_vector(a, b)); }
// transform_normal(frame, n) with non_rigid=false is transform_direction (:3150)
YH_HD frame3f transpose_rot(const frame3f& a) {  // inverse(frame, non_rigid=false) without the origin
  return {{a.x.x, a.y.x, a.z.x}, {a.x.y, a.y.y, a.z.y}, {a.x.z, a.y.z, a.z.z}, {0, 0, 0}};
}
YH_HD frame3f inverse_rigid(const frame3f& a) {  // :2882-2885
  frame3f m = transpose_rot(a);
  m.o       = -(m.x * a.o.x + m.y * a.o.y + m.z * a.o.z);
  return m;
}
// inverse(frame, non_rigid=true): adjoint / determinant (:2735-2741, :2878-2881)
YH_HD frame3f inverse_nonrigid(const frame3f& a) {
  float det  = dot(a.x, cross(a.y, a.z));
  vec3f r0   = cross(a.y, a.z), r1 = cross(a.z, a.x), r2 = cross(a.x, a.y);
  float idet = div_(1.0f, det);
  // adjoint = transpose({r0,r1,r2}); columns of the inverse:
  frame3f m;
  m.x = vec3f{r0.x, r1.x, r2.x} * idet;
  m.y = vec3f{r0.y, r1.y, r2.y} * idet;
  m.z = vec3f{r0.z, r1.z, r2.z} * idet;
  m.o = -(m.x * a.o.x + m.y * a.o.y + m.z * a.o.z);
  return m;
}
YH_HD frame3f frame_fromzx(const vec3f& z_, const vec3f& x_) {  // :2898-2903 with o = 0
  vec3f z = normalize(z_);
  vec3f x = orthonormalize(x_, z);
  vec3f y = normalize(cross(z, x));
  return {x, y, z, {0, 0, 0}};
}
// basis_fromz, yocto_math.h:2743-2752 (returned as a frame with o = 0)
YH_HD frame3f basis_fromz(const vec3f& v) {
  vec3f z    = normalize(v);
  float sign = (f2u(z.z) >> 31) ? -1.0f : 1.0f;  // copysignf(1, z.z)
  float a    = div_(-1.0f, (sign + z.z));
  float b    = z.x * z.y * a;
  vec3f x    = {1.0f + sign * z.x * z.x * a, sign * b, -sign * z.x};
  vec3f y    = {b, sign + z.y * z.y * a, -z.y};
  return {x, y, z, {0, 0, 0}};
}

}  // namespace yh

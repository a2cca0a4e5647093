// Device restatement of the libm entry points the reference binary calls on
// the hot path (nm of the oracle build: expf logf sinf cosf sincosf atan2f
// acosf fmodf float; asin and sinh DOUBLE).
//
// The oracle links glibc 2.39 (Ubuntu GLIBC 2.39-0ubuntu8.5):
//   expf / logf / sinf / cosf  = Arm optimized-routines algorithms (S. Nagy),
//       sysdeps/ieee754/flt-32/{e_expf,e_logf,s_sinf,s_cosf,s_sincosf}.[ch]:
//       double arithmetic over small tables, result rounded once to float.
//       Tables come verbatim from the distribution's libm.a through
//       tools/gen_libm_tables.py (yh_libm_tables.inc).
//   atan2f / atanf / acosf     = Sun fdlibm float code,
//       sysdeps/ieee754/flt-32/{e_atan2f,s_atanf,e_acosf}.c: float arithmetic,
//       so it must run without FMA contraction (this TU: --fmad=false).
//   asin (double)              = CUDA's double routine; the call site rounds to
//       float and the floats agree with glibc's over the whole float domain.
//   sinh (double)              = glibc's e_sinh.c / s_expm1.c (fdlibm) restated in
//       double arithmetic without contraction: the double itself is glibc's.
// Why: mp() sums terms of magnitude 1/v (up to ~3000) in fp32, so a 1-ulp
// difference in logf moves the BSDF by ~1e-4 relative.  1e-5 parity needs the
// same bits out of libm, not merely "accurate" functions.
#pragma once
#include "yh_math.cuh"

namespace yh {

namespace libm_host {
#define YH_LIBM_CONST static const
#include "yh_libm_tables.inc"
#undef YH_LIBM_CONST
}  // namespace libm_host
#if defined(__CUDACC__)
namespace libm_dev {
#define YH_LIBM_CONST static __device__ const
#include "yh_libm_tables.inc"
#undef YH_LIBM_CONST
}  // namespace libm_dev
#endif

#if defined(__CUDA_ARCH__)
#define YH_TAB(name) libm_dev::name
#else
#define YH_TAB(name) libm_host::name
#endif

YH_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  union { double d; uint64_t u; } c; c.d = d; return c.u;
#endif
}
YH_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  union { double d; uint64_t u; } c; c.u = u; return c.d;
#endif
}
// double ops without contraction (nvcc may fuse double a*b+c even when
// --fmad=false only governs float; intrinsics pin the rounding on the device)
YH_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
YH_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
// a * b + c as glibc's running variant computes it.  On an x86-64 CPU with
// FMA+AVX2 (every current server part) glibc's ifunc picks __expf_fma /
// __logf_fma / __sinf_fma / __cosf_fma, in which GCC contracted EVERY
// multiply-add of these routines into vfmadd (checked by disassembling
// libm-2.39.a: e_expf-fma.o, e_logf-fma.o, s_sinf-fma.o, s_cosf-fma.o).
// YH_LIBM_FMA=0 restates the baseline (sse2) variant instead.
#ifndef YH_LIBM_FMA
#define YH_LIBM_FMA 1
#endif
YH_HD double dfma(double a, double b, double c) {
#if YH_LIBM_FMA
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
#else
  return dadd(dmul(a, b), c);
#endif
}

// glibc e_expf.c (__expf)
YH_HD_CALL float expf_(float x) {
  uint32_t abstop = (f2u(x) >> 20) & 0x7ff;
  if (abstop >= 0x42b) {  // top12(88.0f): |x| >= 88 or NaN
    if (f2u(x) == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8) return x + x;
    if (x > 0x1.62e42ep6f) return u2f(0x7f800000u);  // overflow
    if (x < -0x1.9fe368p6f) return 0.0f;             // underflow
  }
  double xd = (double)x;
  double kd = dfma(YH_TAB(yh_expf_invln2_scaled), xd, YH_TAB(yh_expf_shift));
  uint64_t ki = d2u(kd);
  kd          = dadd(kd, -YH_TAB(yh_expf_shift));
  double r    = dfma(YH_TAB(yh_expf_invln2_scaled), xd, -kd);
  uint64_t t  = YH_TAB(yh_exp2f_tab)[ki % 32];
  t += ki << (52 - 5);
  double s  = u2d(t);
  double zz = dfma(YH_TAB(yh_expf_poly_scaled)[0], r, YH_TAB(yh_expf_poly_scaled)[1]);
  double r2 = dmul(r, r);
  double y  = dfma(YH_TAB(yh_expf_poly_scaled)[2], r, 1.0);
  y         = dfma(zz, r2, y);
  y         = dmul(y, s);
  return (float)y;
}

// glibc e_logf.c (__logf)
YH_HD_CALL float logf_(float x) {
  uint32_t ix = f2u(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return u2f(0xff800000u);  // log(+-0) = -inf
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return u2f(0x7fc00000u);
    ix = f2u(x * 0x1p23f);  // subnormal
    ix -= 23u << 23;
  }
  uint32_t tmp = ix - 0x3f330000u;
  int      i   = (tmp >> (23 - 4)) % 16;
  int      k   = (int32_t)tmp >> 23;
  uint32_t iz  = ix - (tmp & 0xff800000u);
  double invc  = YH_TAB(yh_logf_tab)[2 * i];
  double logc  = YH_TAB(yh_logf_tab)[2 * i + 1];
  double z     = (double)u2f(iz);
  double r     = dfma(z, invc, -1.0);
  double y0    = dfma((double)k, YH_TAB(yh_logf_ln2), logc);
  double r2    = dmul(r, r);
  double y     = dfma(YH_TAB(yh_logf_poly)[1], r, YH_TAB(yh_logf_poly)[2]);
  y            = dfma(YH_TAB(yh_logf_poly)[0], r2, y);
  y            = dfma(y, r2, dadd(y0, r));
  return (float)y;
}

// glibc s_sincosf.h: table layout on x86-64 is sign[4], hpi_inv*2^24, hpi,
// then {c0,c1},{s1,c2},{s2,c3},{s3,c4} (sysdeps/x86_64/fpu/s_sincosf_data.c).
struct sincos_tab {
  const double* p;
  YH_HD double sign(int i) const { return p[i]; }
  YH_HD double hpi_inv() const { return p[4]; }
  YH_HD double hpi() const { return p[5]; }
  YH_HD double c0() const { return p[6]; }
  YH_HD double c1() const { return p[7]; }
  YH_HD double s1() const { return p[8]; }
  YH_HD double c2() const { return p[9]; }
  YH_HD double s2() const { return p[10]; }
  YH_HD double c3() const { return p[11]; }
  YH_HD double s3() const { return p[12]; }
  YH_HD double c4() const { return p[13]; }
};
YH_HD float sinf_poly(double x, double x2, const sincos_tab& p, int n) {
  if ((n & 1) == 0) {
    double x3 = dmul(x, x2);
    double s1 = dfma(x2, p.s3(), p.s2());
    double x7 = dmul(x3, x2);
    double s  = dfma(x3, p.s1(), x);
    return (float)dfma(x7, s1, s);
  } else {
    double x4 = dmul(x2, x2);
    double c2 = dfma(x2, p.c4(), p.c3());
    double c1 = dfma(x2, p.c1(), p.c0());
    double x6 = dmul(x4, x2);
    double c  = dfma(x4, p.c2(), c1);
    return (float)dfma(x6, c2, c);
  }
}
YH_HD double reduce_fast(double x, const sincos_tab& p, int* np) {
  double r = dmul(x, p.hpi_inv());
  int    n = ((int32_t)r + 0x800000) >> 24;
  *np      = n;
  return dfma(-(double)n, p.hpi(), x);
}
YH_HD double reduce_large(uint32_t xi, int* np) {
  const uint32_t* arr   = &YH_TAB(yh_inv_pio4)[(xi >> 26) & 15];
  int             shift = (xi >> 23) & 7;
  uint64_t        n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  res0 = (uint64_t)(xi * arr[0]);
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  double x = (double)(int64_t)res0;
  *np      = (int)n;
  return dmul(x, 0x1.921FB54442D18p-62);
}
// sinf (which = 0) / cosf (which = 1): glibc s_sinf.c / s_cosf.c
YH_HD_CALL float sincosf_(float y, int which) {
  double     x = (double)y;
  int        n;
  sincos_tab p{&YH_TAB(yh_sincosf_tab)[0]};
  uint32_t   top = (f2u(y) >> 20) & 0x7ff;
  if (top < 0x3f4) {  // |y| < pi/4 (abstop12(pio4) = 0x3f4)
    double s = dmul(x, x);
    if (top < 0x398) return which ? 1.0f : y;  // |y| < 2^-12
    return sinf_poly(x, s, p, which);
  } else if (top < 0x42f) {  // |y| < 120
    x        = reduce_fast(x, p, &n);
    double s = p.sign(n & 3);
    if (n & 2) p.p = &YH_TAB(yh_sincosf_tab)[14];
    return sinf_poly(dmul(x, s), dmul(x, x), p, n ^ which);
  } else if (top < 0x7f8) {
    uint32_t xi   = f2u(y);
    int      sign = xi >> 31;
    x             = reduce_large(xi, &n);
    double s      = p.sign((n + sign) & 3);
    if ((n + sign) & 2) p.p = &YH_TAB(yh_sincosf_tab)[14];
    return sinf_poly(dmul(x, s), dmul(x, x), p, n ^ which);
  }
  return u2f(0x7fc00000u);
}
YH_HD float sinf_(float x) { return sincosf_(x, 0); }
YH_HD float cosf_(float x) { return sincosf_(x, 1); }

// fdlibm s_atanf.c (__atanf), float arithmetic
YH_HD float atanf_(float x) {
  const float atanhi[4] = {u2f(0x3eed6338u), u2f(0x3f490fdau), u2f(0x3f7b985eu), u2f(0x3fc90fdau)};
  const float atanlo[4] = {u2f(0x31ac3769u), u2f(0x33222168u), u2f(0x33140fb4u), u2f(0x33a22168u)};
  const float aT0 = u2f(0x3eaaaaabu), aT1 = u2f(0xbe4ccccdu), aT2 = u2f(0x3e124925u),
              aT3 = u2f(0xbde38e38u), aT4 = u2f(0x3dba2e6eu), aT5 = u2f(0xbd9d8795u),
              aT6 = u2f(0x3d886b35u), aT7 = u2f(0xbd6ef16bu), aT8 = u2f(0x3d4bda59u),
              aT9 = u2f(0xbd15a221u), aT10 = u2f(0x3c8569d7u);
  int32_t hx = (int32_t)f2u(x);
  int32_t ix = hx & 0x7fffffff;
  int     id;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return x + x;
    if (hx > 0) return atanhi[3] + atanlo[3];
    return -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = u2f((uint32_t)ix);  // fabsf
    if (ix < 0x3f980000) {      // |x| < 1.1875
      if (ix < 0x3f300000) {    // 7/16 <= |x| < 11/16
        id = 0;
        x  = div_((2.0f * x - 1.0f), (2.0f + x));
      } else {                  // 11/16 <= |x| < 19/16
        id = 1;
        x  = div_((x - 1.0f), (x + 1.0f));
      }
    } else {
      if (ix < 0x401c0000) {    // |x| < 2.4375
        id = 2;
        x  = div_((x - 1.5f), (1.0f + 1.5f * x));
      } else {                  // 2.4375 <= |x| < 2^25
        id = 3;
        x  = div_(-1.0f, x);
      }
    }
  }
  float z  = x * x;
  float w  = z * z;
  float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
  float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
  if (id < 0) return x - x * (s1 + s2);
  z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}

// fdlibm e_atan2f.c (__ieee754_atan2f)
YH_HD_CALL float atan2f_(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = u2f(0x3f490fdbu), pi_o_2 = u2f(0x3fc90fdbu),
              pi = u2f(0x40490fdbu), pi_lo = u2f(0xb3bbbd2eu);
  int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
  int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return atanf_(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0:
      case 1: return y;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi + tiny;
        default: return -pi - tiny;
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  int   k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_(u2f(f2u(div_(y, x)) & 0x7fffffffu));  // fabsf
  switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

// fdlibm e_acosf.c (__ieee754_acosf)
YH_HD_CALL float acosf_(float x) {
  const float pi = u2f(0x40490fdau), pio2_hi = u2f(0x3fc90fdau), pio2_lo = u2f(0x33a22168u),
              pS0 = u2f(0x3e2aaaabu), pS1 = u2f(0xbea6b090u), pS2 = u2f(0x3e4e0aa8u),
              pS3 = u2f(0xbd241146u), pS4 = u2f(0x3a4f7f04u), pS5 = u2f(0x3811ef08u),
              qS1 = u2f(0xc019d139u), qS2 = u2f(0x4001572du), qS3 = u2f(0xbf303361u),
              qS4 = u2f(0x3d9dc62eu);
  int32_t hx = (int32_t)f2u(x);
  int32_t ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) {
    if (hx > 0) return 0.0f;
    return pi + 2.0f * pio2_lo;
  } else if (ix > 0x3f800000) {
    return u2f(0x7fc00000u);
  }
  if (ix < 0x3f000000) {  // |x| < 0.5
    if (ix <= 0x32800000) return pio2_hi + pio2_lo;
    float z = x * x;
    float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    float r = div_(p, q);
    return pio2_hi - (x - (pio2_lo - r * x));
  } else if (hx < 0) {  // x < -0.5
    float z = (1.0f + x) * 0.5f;
    float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    float s = sqrt_(z);
    float r = div_(p, q);
    float w = r * s - pio2_lo;
    return pi - 2.0f * (s + w);
  } else {  // x > 0.5
    float z  = (1.0f - x) * 0.5f;
    float s  = sqrt_(z);
    float df = u2f(f2u(s) & 0xfffff000u);
    float c  = div_((z - df * df), (s + df));
    float p  = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    float q  = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    float r  = div_(p, q);
    float w  = r * s + c;
    return 2.0f * (df + w);
  }
}

// fmodf is exact in IEEE arithmetic, so any correct implementation agrees.
YH_HD float fmodf_(float a, float b) {
#if defined(__CUDA_ARCH__)
  return ::fmodf(a, b);
#else
  return __builtin_fmodf(a, b);
#endif
}

// double asin: CUDA's routine; every call site rounds the result to float, and the floats agree with glibc's for EVERY
// float argument in [-1, 1] (tests/test_hair_gpu.py::test_double_asin_and_sinh_over_their_whole_float_domains).
YH_HD double asin_d(double x) { return ::asin(x); }

// double sinh as glibc computes it (sysdeps/ieee754/dbl-64/e_sinh.c over s_expm1.c, the fdlibm routines; glibc 2.39
// has no FMA variant of either, so every multiply-add is two roundings).  mp() divides by sinh(1 / v) * 2 v in double
// (yocto_extension.cpp:205-206); CUDA's own ::sinh differs from glibc's in the last bits of the double for 1.4 % of
// the arguments, which can move the float quotient across a rounding boundary once in ~10^10 calls.  With the same
// operations in the same order the denominator carries glibc's bits for every float v (checked exhaustively over
// (0.1, 110], host and device).
YH_HD uint32_t d_hi(double x) { return (uint32_t)(d2u(x) >> 32); }
YH_HD double   d_with_hi(double x, uint32_t hi) { return u2d((d2u(x) & 0xffffffffull) | ((uint64_t)hi << 32)); }
YH_HD double   ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
// s_expm1.c.  On a CPU with FMA + AVX2 glibc's ifunc picks __expm1_fma, in which GCC contracted the multiply-adds
// marked dfma below (read off the disassembly of libm-2.39: vfmadd / vfnmadd / vfmsub at exactly these places; k, the
// sums of differences and the exponent arithmetic are not fused).  YH_LIBM_FMA = 0 gives the baseline variant.
// e_sinh.c itself has no FMA variant.
YH_HD double expm1_d(double x) {
  const double one = 1.0, huge = 1.0e+300, o_threshold = 7.09782712893383973096e+02, ln2_hi = 6.93147180369123816490e-01,
               ln2_lo = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00, Q1 = -3.33333333333331316428e-02,
               Q2 = 1.58730158725481460165e-03, Q3 = -7.93650757867487942473e-05, Q4 = 4.00821782732936239552e-06,
               Q5 = -2.01099218183624371326e-07;
  double   y, hi, lo, c = 0, t, e, hxs, hfx, r1, h2, h4, R1, R2, R3;
  int      k;
  uint32_t hx  = d_hi(x);
  uint32_t xsb = hx & 0x80000000u;
  hx &= 0x7fffffffu;
  if (hx >= 0x4043687Au) {  // |x| >= 56 ln2
    if (hx >= 0x40862E42u) {  // |x| >= 709.78
      if (hx >= 0x7ff00000u) {
        if (((hx & 0xfffffu) | (uint32_t)d2u(x)) != 0) return dadd(x, x);  // NaN
        return xsb == 0 ? x : -1.0;                                       // exp(+-inf) - 1
      }
      if (x > o_threshold) return dmul(huge, huge);
    }
    if (xsb != 0) return -1.0;  // x < -56 ln2 (the reference raises inexact here)
  }
  if (hx > 0x3fd62e42u) {      // |x| > 0.5 ln2
    if (hx < 0x3FF0A2B2u) {    // and |x| < 1.5 ln2
      if (xsb == 0) hi = dadd(x, -ln2_hi), lo = ln2_lo, k = 1;
      else hi = dadd(x, ln2_hi), lo = -ln2_lo, k = -1;
    } else {
      k  = (int)dadd(dmul(invln2, x), xsb == 0 ? 0.5 : -0.5);
      t  = (double)k;
      hi = dfma(-t, ln2_hi, x);
      lo = dmul(t, ln2_lo);
    }
    x = dadd(hi, -lo);
    c = dadd(dadd(hi, -x), -lo);
  } else if (hx < 0x3c900000u) {  // |x| < 2^-54
    t = dadd(huge, x);
    return dadd(x, -dadd(t, -dadd(huge, x)));
  } else {
    k = 0;
  }
  hfx = dmul(0.5, x);
  hxs = dmul(x, hfx);
  R1  = dfma(hxs, Q1, one);
  h2  = dmul(hxs, hxs);
  R2  = dfma(hxs, Q3, Q2);
  h4  = dmul(h2, h2);
  R3  = dfma(hxs, Q5, Q4);
  r1  = dfma(h4, R3, dfma(h2, R2, R1));
  t   = dfma(-r1, hfx, 3.0);
  e   = dmul(hxs, ddiv(dadd(r1, -t), dfma(-x, t, 6.0)));
  if (k == 0) return dadd(x, -dfma(x, e, -hxs));
  e = dfma(x, dadd(e, -c), -c);
  e = dadd(e, -hxs);
  if (k == -1) return dfma(0.5, dadd(x, -e), -0.5);
  if (k == 1) {
    if (x < -0.25) return dmul(-2.0, dadd(e, -dadd(x, 0.5)));
    return dfma(2.0, dadd(x, -e), one);
  }
  if (k <= -2 || k > 56) {
    y = dadd(one, -dadd(e, -x));
    y = d_with_hi(y, d_hi(y) + ((uint32_t)k << 20));
    return dadd(y, -one);
  }
  t = one;
  if (k < 20) {
    t = d_with_hi(t, 0x3ff00000u - (0x200000u >> k));  // 1 - 2^-k
    y = dadd(t, -dadd(e, -x));
    y = d_with_hi(y, d_hi(y) + ((uint32_t)k << 20));
  } else {
    t = d_with_hi(t, (uint32_t)(0x3ff - k) << 20);  // 2^-k
    y = dadd(x, -dadd(e, t));
    y = dadd(y, one);
    y = d_with_hi(y, d_hi(y) + ((uint32_t)k << 20));
  }
  return y;
}
YH_HD_CALL double sinh_d(double x) {  // e_sinh.c
  const double one = 1.0, shuge = 1.0e307;
  uint32_t jx = d_hi(x), ix = jx & 0x7fffffffu;
  if (ix >= 0x7ff00000u) return dadd(x, x);
  double h = (jx & 0x80000000u) ? -0.5 : 0.5;
  double ax = u2d(d2u(x) & 0x7fffffffffffffffull);
  if (ix < 0x40360000u) {  // |x| < 22
    if (ix < 0x3e300000u && dadd(shuge, x) > one) return x;  // |x| < 2^-28
    double t = expm1_d(ax);
    if (ix < 0x3ff00000u) return dmul(h, dadd(dmul(2.0, t), -ddiv(dmul(t, t), dadd(t, one))));
    return dmul(h, dadd(t, ddiv(t, dadd(t, one))));
  }
  // |x| >= 22 does not occur on this path (x = 1 / v, v > 0.1); CUDA's routine keeps the function total
#if defined(__CUDA_ARCH__)
  return ::sinh(x);
#else
  return __builtin_sinh(x);
#endif
}

}  // namespace yh

/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Plain-C restatement of the reference's hair path-tracing hot path
 * (dsforza96/yocto-hair), written from the reference's algorithm; every
 * function cites the reference lines it follows (paths relative to the
 * reference checkout: ext = libs/yocto_extension/yocto_extension.cpp,
 * pt = libs/yocto_pathtrace/yocto_pathtrace.cpp, ym = libs/yocto/yocto_math.h).
 *
 * Scope = SURVEY.md section 8a: PCG32, the pbrt hair BSDF and its four
 * statistical self-tests, line / triangle / box tests, the middle-split BVH
 * (build + depth-first walk, two levels), camera, diffuse + hair lobes,
 * environment / area lights, trace_path (MIS path shader), trace_sample,
 * init_state; and the material model of the triangles that share a scene
 * with the hair: specular / metal / thin transmission /
 * refraction lobes and their delta forms, opacity, homogeneous volumes, float
 * RGB material textures and normal maps (byte textures are expanded by the
 * caller, as for the CUDA library).
 *
 * Arithmetic: fp32, no FMA contraction (-ffp-contract=off), glibc libm, with
 * the reference's two double-precision calls (asin, sinh) kept in double.
 *
 * Parity: PINNED.  tests/test_oracle_cpu.py checks this file bit for bit
 * against oracle/_ref/libyh_ref.so (the unmodified reference sources compiled
 * by oracle/Makefile) and against tests/golden/ (generated from that build).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load the
 * library built from this file.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../include/yocto_hair_b200.h"

#define PIF 3.14159265358979323846f /* ym:214 */
#define FLT_MAXF 3.402823466e+38f
#define FLT_EPSF 1.192092896e-07f
#define RAY_EPS 1e-4f /* ym:1106 */

typedef struct { float x, y; } v2;
typedef struct { float x, y, z; } v3;
typedef struct { v3 x, y, z, o; } fr3; /* ym:932-947 */
typedef struct { v3 lo, hi; } box3;

/* ---- scalar / vector helpers (ym:1778-2120); min/max keep the (a<b)?a:b form */
static inline float fminr(float a, float b) { return (a < b) ? a : b; }
static inline float fmaxr(float a, float b) { return (a > b) ? a : b; }
static inline float clampr(float a, float lo, float hi) { return fminr(fmaxr(a, lo), hi); }
static inline int   clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }
static inline float absr(float a) { return a < 0 ? -a : a; }
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdiv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 vs(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vds(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 vadds(v3 a, float s) { return V(a.x + s, a.y + s, a.z + s); }
static inline v3 vsubs(v3 a, float s) { return V(a.x - s, a.y - s, a.z - s); }
static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* ym:2029 */
static inline v3 cross3(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float len3(v3 a) { return sqrtf(dot3(a, a)); }
static inline v3 norm3(v3 a) { float l = len3(a); return (l != 0) ? vds(a, l) : a; } /* ym:2037 */
static inline v3 orthonorm(v3 a, v3 b) { return norm3(vsub(a, vs(b, dot3(a, b)))); } /* ym:2056 */
static inline float max3(v3 a) { return fmaxr(fmaxr(a.x, a.y), a.z); }              /* ym:2092 */
static inline int   nonzero3(v3 a) { return !(a.x == 0 && a.y == 0 && a.z == 0); }   /* ym:1846 */
static inline int   finite3(v3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }
static inline v3 vmin(v3 a, v3 b) { return V(fminr(a.x, b.x), fminr(a.y, b.y), fminr(a.z, b.z)); }
static inline v3 vmax(v3 a, v3 b) { return V(fmaxr(a.x, b.x), fmaxr(a.y, b.y), fmaxr(a.z, b.z)); }
static inline float lum(v3 a) { return 0.2126f * a.x + 0.7152f * a.y + 0.0722f * a.z; } /* ym:3737 */
static inline v3 lerp3(v3 a, v3 b, float u) { return vadd(vs(a, 1 - u), vs(b, u)); }       /* interpolate_line */
static inline float comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* frames, ym:3136-3160, :2877-2903 */
static inline v3 xf_vector(const fr3* f, v3 b) { return vadd(vadd(vs(f->x, b.x), vs(f->y, b.y)), vs(f->z, b.z)); }
static inline v3 xf_point(const fr3* f, v3 b) { return vadd(xf_vector(f, b), f->o); }
static inline v3 xf_direction(const fr3* f, v3 b) { return norm3(xf_vector(f, b)); }
static fr3 frame_from(const float* p) {
  fr3 f = {V(p[0], p[1], p[2]), V(p[3], p[4], p[5]), V(p[6], p[7], p[8]), V(p[9], p[10], p[11])};
  return f;
}
static fr3 inverse_frame(const fr3* a, int non_rigid) { /* ym:2735-2741, :2877-2885 */
  fr3 m;
  if (non_rigid) {
    float det  = dot3(a->x, cross3(a->y, a->z));
    v3    r0 = cross3(a->y, a->z), r1 = cross3(a->z, a->x), r2 = cross3(a->x, a->y);
    float idet = 1 / det;
    m.x = vs(V(r0.x, r1.x, r2.x), idet), m.y = vs(V(r0.y, r1.y, r2.y), idet), m.z = vs(V(r0.z, r1.z, r2.z), idet);
  } else {
    m.x = V(a->x.x, a->y.x, a->z.x), m.y = V(a->x.y, a->y.y, a->z.y), m.z = V(a->x.z, a->y.z, a->z.z);
  }
  m.o = vneg(vadd(vadd(vs(m.x, a->o.x), vs(m.y, a->o.y)), vs(m.z, a->o.z)));
  return m;
}
static fr3 frame_fromzx(v3 z_, v3 x_) { /* ym:2898-2903, o = 0 */
  fr3 f;
  f.z = norm3(z_), f.x = orthonorm(x_, f.z), f.y = norm3(cross3(f.z, f.x)), f.o = V(0, 0, 0);
  return f;
}
static fr3 basis_fromz(v3 v) { /* ym:2743-2752 */
  fr3   f;
  v3    z = norm3(v);
  float sign = copysignf(1.0f, z.z), a = -1.0f / (sign + z.z), b = z.x * z.y * a;
  f.x = V(1.0f + sign * z.x * z.x * a, sign * b, -sign * z.x);
  f.y = V(b, sign + z.y * z.y * a, -z.y);
  f.z = z, f.o = V(0, 0, 0);
  return f;
}

/* ================================================================= PCG32 == */
typedef struct { uint64_t state, inc; } rng_t;                 /* ym:1396-1401 */
static uint32_t rng_next(rng_t* r) {                            /* ym:1410-1416 */
  uint64_t old = r->state;
  r->state     = old * 6364136223846793005ULL + r->inc;
  uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
  return (xs >> rot) | (xs << ((~rot + 1u) & 31));
}
static rng_t rng_make(uint64_t seed, uint64_t seq) {            /* ym:1419-1427 */
  rng_t r = {0, (seq << 1u) | 1u};
  rng_next(&r);
  r.state += seed;
  rng_next(&r);
  return r;
}
static int   rng_1i(rng_t* r, int n) { return (int)(rng_next(r) % (uint32_t)n); } /* ym:1430 */
static float rng_1f(rng_t* r) {                                                   /* ym:1431-1441 */
  union { uint32_t u; float f; } c;
  c.u = (rng_next(r) >> 9) | 0x3f800000u;
  return c.f - 1.0f;
}
static v2 rng_2f(rng_t* r) { v2 v; v.x = rng_1f(r); v.y = rng_1f(r); return v; }   /* ym:1442-1447 */

/* ============================================================= hair BSDF == */
#define PMAX 3 /* ext header :84 */
typedef struct {
  v3    sigma_a;
  float alpha, eta, h, v[PMAX + 1], s, sin2k[3], cos2k[3], gamma_o;
  fr3   w2b;
} hbrdf;

static inline float safe_asin(float x) { return (float)asin((double)clampr(x, -1.0f, 1.0f)); } /* ext:111, double asin */
static inline float safe_sqrt(float x) { return sqrtf(fmaxr(0.0f, x)); }                     /* ext:113 */
static inline float sqrf(float x) { return x * x; }
static float powi(float v, int n) { /* ext:95-109: pow<n> = pow<n/2>^2 * pow<n&1> */
  if (n == 0) return 1;
  if (n == 1) return v;
  float h = powi(v, n / 2);
  return h * h * powi(v, n & 1);
}

static hbrdf hair_brdf_make(const yh_hair_material* m, float v, v3 normal, v3 tangent) { /* ext:127-177 */
  hbrdf b;
  memset(&b, 0, sizeof b);
  v3 sa = V(m->sigma_a[0], m->sigma_a[1], m->sigma_a[2]), col = V(m->color[0], m->color[1], m->color[2]);
  if (nonzero3(sa)) {
    b.sigma_a = sa;
  } else if (nonzero3(col)) { /* sigma_a_from_reflectance, ext:121-125 */
    float bn = m->beta_n;
    float den = 5.969f - 0.215f * bn + 2.532f * sqrf(bn) - 10.73f * powi(bn, 3) + 5.574f * powi(bn, 4) + 0.245f * powi(bn, 5);
    v3 q = vds(V(logf(col.x), logf(col.y), logf(col.z)), den);
    b.sigma_a = vmul(q, q);
  } else if (m->eumelanin || m->pheomelanin) { /* sigma_a_from_concentration, ext:115-119 */
    b.sigma_a = vadd(vs(V(0.419f, 0.697f, 1.37f), m->eumelanin), vs(V(0.187f, 0.4f, 1.05f), m->pheomelanin));
  }
  float bm = m->beta_m, bn = m->beta_n;
  b.alpha = m->alpha, b.eta = m->eta;
  b.h       = -1 + 2 * v;
  b.gamma_o = safe_asin(b.h);
  b.v[0]    = sqrf(0.726f * bm + 0.812f * sqrf(bm) + 3.7f * powi(bm, 20));
  b.v[1]    = 0.25f * b.v[0];
  b.v[2]    = 4 * b.v[0];
  b.v[3]    = b.v[2];
  b.s       = 0.626657069f * (0.265f * bn + 1.194f * sqrf(bn) + 5.372f * powi(bn, 22)); /* sqrt(pi/8) */
  b.sin2k[0] = sinf(PIF / 180 * b.alpha);
  b.cos2k[0] = safe_sqrt(1 - sqrf(b.sin2k[0]));
  for (int i = 1; i < 3; i++) {
    b.sin2k[i] = 2 * b.cos2k[i - 1] * b.sin2k[i - 1];
    b.cos2k[i] = sqrf(b.cos2k[i - 1]) - sqrf(b.sin2k[i - 1]);
  }
  fr3 f = frame_fromzx(normal, tangent);
  b.w2b = inverse_frame(&f, 0);
  return b;
}

static float bessel_i0(float x) { /* ext:179-192 */
  float   val = 0, x2i = 1;
  int64_t ifact = 1;
  int     i4 = 1;
  for (int i = 0; i < 10; i++) {
    if (i > 1) ifact *= i;
    val += x2i / (float)(i4 * ifact * ifact);
    x2i *= x * x;
    i4 *= 4;
  }
  return val;
}
static float log_i0(float x) { /* ext:194-199 */
  if (x > 12) return x + 0.5f * (-logf(2 * PIF) + logf(1 / x) + 1 / (8 * x));
  return logf(bessel_i0(x));
}
static float hair_mp(float cti, float cto, float sti, float sto, float v) { /* ext:201-207 */
  float a = cti * cto / v, b = sti * sto / v;
  if (v <= 0.1f) return expf(log_i0(a) - b - 1 / v + 0.6931f + logf(1 / (2 * v)));
  return (float)((double)(expf(-b) * bessel_i0(a)) / (sinh((double)(1 / v)) * 2 * (double)v)); /* double sinh */
}
static float fresnel_cos(float eta, float cos_theta) { /* ym:4215-4235 with dot = cos_theta */
  float cosw = absr(0.0f * 0.0f + 0.0f * 0.0f + 1.0f * cos_theta);
  float sin2 = 1 - cosw * cosw, eta2 = eta * eta, cos2t = 1 - sin2 / eta2;
  if (cos2t < 0) return 1;
  float t0 = sqrtf(cos2t), t1 = eta * t0, t2 = eta * cosw;
  float rs = (cosw - t1) / (cosw + t1), rp = (t0 - t2) / (t0 + t2);
  return (rs * rs + rp * rp) / 2;
}
static void hair_ap(float cto, float eta, float h, v3 T, v3 ap[PMAX + 1]) { /* ext:209-230 */
  float cgo = safe_sqrt(1 - h * h), ct = cto * cgo;
  float f = fresnel_cos(eta, ct);
  ap[0] = V(f, f, f);
  ap[1] = vs(T, sqrf(1 - f));
  for (int p = 2; p < PMAX; p++) ap[p] = vs(vmul(ap[p - 1], T), f);
  ap[PMAX] = vdiv(vmul(vs(ap[PMAX - 1], f), T), vsub(V(1, 1, 1), vs(T, f)));
}
static float hair_phi(int p, float go, float gt) { return 2 * p * gt - 2 * go + p * PIF; } /* ext:232-234 */
static float logistic(float x, float s) {                                                   /* ext:236-239 */
  x = absr(x);
  return expf(-x / s) / (s * sqrf(1 + expf(-x / s)));
}
static float logistic_cdf(float x, float s) { return 1 / (1 + expf(-x / s)); } /* ext:241 */
static float hair_np(float phi, int p, float s, float go, float gt) {          /* ext:247-253 */
  float dphi = phi - hair_phi(p, go, gt);
  while (dphi > PIF) dphi -= 2 * PIF;
  while (dphi < -PIF) dphi += 2 * PIF;
  return logistic(dphi, s) / (logistic_cdf(PIF, s) - logistic_cdf(-PIF, s));
}
static void hair_tilt(const hbrdf* b, int p, float sto, float cto, float* stop, float* ctop) { /* ext:297-322 */
  if (p == 0) {
    *stop = sto * b->cos2k[1] - cto * b->sin2k[1], *ctop = cto * b->cos2k[1] + sto * b->sin2k[1];
  } else if (p == 1) {
    *stop = sto * b->cos2k[0] + cto * b->sin2k[0], *ctop = cto * b->cos2k[0] - sto * b->sin2k[0];
  } else if (p == 2) {
    *stop = sto * b->cos2k[2] + cto * b->sin2k[2], *ctop = cto * b->cos2k[2] - sto * b->sin2k[2];
  } else {
    *stop = sto, *ctop = cto;
  }
}
static v3 exp_neg_scaled(v3 sigma_a, float k) { /* exp(-sigma_a * k) */
  v3 a = vs(vneg(sigma_a), k);
  return V(expf(a.x), expf(a.y), expf(a.z));
}

static v3 hair_eval(const hbrdf* b, v3 wo_, v3 wi_) { /* ext:255-336 */
  v3    wo = xf_direction(&b->w2b, wo_), wi = xf_direction(&b->w2b, wi_);
  float sto = wo.x, cto = safe_sqrt(1 - sqrf(sto)), phio = atan2f(wo.z, wo.y);
  float sti = wi.x, cti = safe_sqrt(1 - sqrf(sti)), phii = atan2f(wi.z, wi.y);
  float stt = sto / b->eta, ctt = safe_sqrt(1 - sqrf(stt));
  float etap = sqrtf(b->eta * b->eta - sqrf(sto)) / cto;
  float sgt = b->h / etap, cgt = safe_sqrt(1 - sqrf(sgt)), gt = safe_asin(sgt);
  v3    T = exp_neg_scaled(b->sigma_a, 2 * cgt / ctt);
  float phi = phii - phio;
  v3    ap[PMAX + 1];
  hair_ap(cto, b->eta, b->h, T, ap);
  v3 fsum = V(0, 0, 0);
  for (int p = 0; p < PMAX; p++) {
    float stop, ctop;
    hair_tilt(b, p, sto, cto, &stop, &ctop);
    ctop = absr(ctop);
    fsum = vadd(fsum, vs(vs(ap[p], hair_mp(cti, ctop, sti, stop, b->v[p])), hair_np(phi, p, b->s, b->gamma_o, gt)));
  }
  fsum = vadd(fsum, vds(vs(ap[PMAX], hair_mp(cti, cto, sti, sto, b->v[PMAX])), 2 * PIF));
  return fsum;
}

static uint32_t compact1by1(uint32_t x) { /* ext:339-351 */
  x &= 0x55555555;
  x = (x ^ (x >> 1)) & 0x33333333;
  x = (x ^ (x >> 2)) & 0x0f0f0f0f;
  x = (x ^ (x >> 4)) & 0x00ff00ff;
  x = (x ^ (x >> 8)) & 0x0000ffff;
  return x;
}
static v2 demux_float(float f) { /* ext:353-357 */
  uint64_t v = (uint64_t)(f * (float)(1ull << 32));
  v2 r = {compact1by1((uint32_t)v) / (float)(1 << 16), compact1by1((uint32_t)(v >> 1)) / (float)(1 << 16)};
  return r;
}
static float sample_trimmed_logistic(float u, float s, float a, float b) { /* ext:359-363 */
  float k = logistic_cdf(b, s) - logistic_cdf(a, s);
  float x = -s * logf(1 / (u * k + logistic_cdf(a, s)) - 1);
  return clampr(x, a, b);
}
static void hair_ap_pdf(const hbrdf* b, float cto, float pdf[PMAX + 1]) { /* ext:365-397 */
  float sto = safe_sqrt(1 - cto * cto);
  float stt = sto / b->eta, ctt = safe_sqrt(1 - sqrf(stt));
  float etap = sqrtf(b->eta * b->eta - sqrf(sto)) / cto;
  float sgt = b->h / etap, cgt = safe_sqrt(1 - sqrf(sgt));
  v3    T = exp_neg_scaled(b->sigma_a, 2 * cgt / ctt);
  v3    ap[PMAX + 1];
  hair_ap(cto, b->eta, b->h, T, ap);
  float sum = 0.0f;
  for (int i = 0; i <= PMAX; i++) sum += lum(ap[i]);
  for (int i = 0; i <= PMAX; i++) pdf[i] = lum(ap[i]) / sum;
}
static v3 hair_sample(const hbrdf* b, v3 wo_, v2 rn) { /* ext:399-479 */
  v3    wo = xf_direction(&b->w2b, wo_);
  float sto = wo.x, cto = safe_sqrt(1 - sqrf(sto)), phio = atan2f(wo.z, wo.y);
  v2    u0 = demux_float(rn.x), u1 = demux_float(rn.y);
  float appdf[PMAX + 1];
  hair_ap_pdf(b, cto, appdf);
  int p;
  for (p = 0; p < PMAX; p++) {
    if (u0.x < appdf[p]) break;
    u0.x -= appdf[p];
  }
  float stop, ctop;
  hair_tilt(b, p, sto, cto, &stop, &ctop);
  u1.x = fmaxr(u1.x, 1e-5f);
  float ct = 1 + b->v[p] * logf(u1.x + (1 - u1.x) * expf(-2 / b->v[p]));
  float st = safe_sqrt(1 - sqrf(ct));
  float cphi = cosf(2 * PIF * u1.y);
  float sti = -ct * stop + st * cphi * ctop;
  float cti = safe_sqrt(1 - sqrf(sti));
  float etap = sqrtf(b->eta * b->eta - sqrf(sto)) / cto;
  float sgt = b->h / etap, gt = safe_asin(sgt);
  float dphi;
  if (p < PMAX) dphi = hair_phi(p, b->gamma_o, gt) + sample_trimmed_logistic(u0.y, b->s, -PIF, PIF);
  else dphi = 2 * PIF * u0.y;
  float phii = phio + dphi;
  v3    wi = V(sti, cti * cosf(phii), cti * sinf(phii));
  fr3   inv = inverse_frame(&b->w2b, 0);
  return xf_direction(&inv, wi);
}
static float hair_pdf(const hbrdf* b, v3 wo_, v3 wi_) { /* ext:481-551 */
  v3    wo = xf_direction(&b->w2b, wo_), wi = xf_direction(&b->w2b, wi_);
  float sto = wo.x, cto = safe_sqrt(1 - sqrf(sto)), phio = atan2f(wo.z, wo.y);
  float sti = wi.x, cti = safe_sqrt(1 - sqrf(sti)), phii = atan2f(wi.z, wi.y);
  float etap = sqrtf(b->eta * b->eta - sqrf(sto)) / cto;
  float sgt = b->h / etap, gt = safe_asin(sgt);
  float appdf[PMAX + 1];
  hair_ap_pdf(b, cto, appdf);
  float phi = phii - phio, pdf = 0.0f;
  for (int p = 0; p < PMAX; p++) {
    float stop, ctop;
    hair_tilt(b, p, sto, cto, &stop, &ctop);
    ctop = absr(ctop);
    pdf += hair_mp(cti, ctop, sti, stop, b->v[p]) * appdf[p] * hair_np(phi, p, b->s, b->gamma_o, gt);
  }
  pdf += hair_mp(cti, cto, sti, sto, b->v[PMAX]) * appdf[PMAX] * (1 / (2 * PIF));
  return pdf;
}

static void pack_brdf(const hbrdf* b, yh_hair_brdf* o) {
  memcpy(o->sigma_a, &b->sigma_a, 12);
  o->alpha = b->alpha, o->eta = b->eta, o->h = b->h;
  memcpy(o->v, b->v, 16);
  o->s = b->s;
  memcpy(o->sin_2k_alpha, b->sin2k, 12);
  memcpy(o->cos_2k_alpha, b->cos2k, 12);
  o->gamma_o = b->gamma_o;
  memcpy(o->world_to_brdf, &b->w2b, 48);
}
static hbrdf unpack_brdf(const yh_hair_brdf* o) {
  hbrdf b;
  memcpy(&b.sigma_a, o->sigma_a, 12);
  b.alpha = o->alpha, b.eta = o->eta, b.h = o->h;
  memcpy(b.v, o->v, 16);
  b.s = o->s;
  memcpy(b.sin2k, o->sin_2k_alpha, 12);
  memcpy(b.cos2k, o->cos_2k_alpha, 12);
  b.gamma_o = o->gamma_o;
  memcpy(&b.w2b, o->world_to_brdf, 48);
  return b;
}
static v3 ld3(const float* p, int64_t i) { return V(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

void yho_hair_brdf_batch(const yh_hair_material* m, const float* v, const float* n, const float* t, int64_t count,
    yh_hair_brdf* out) {
  for (int64_t i = 0; i < count; i++) {
    hbrdf b = hair_brdf_make(&m[i], v[i], ld3(n, i), ld3(t, i));
    pack_brdf(&b, &out[i]);
  }
}
void yho_hair_eval_batch(const yh_hair_brdf* b, const float* wo, const float* wi, int64_t count, float* out) {
  for (int64_t i = 0; i < count; i++) {
    hbrdf hb = unpack_brdf(&b[i]);
    v3    f = hair_eval(&hb, ld3(wo, i), ld3(wi, i));
    memcpy(out + 3 * i, &f, 12);
  }
}
void yho_hair_sample_batch(const yh_hair_brdf* b, const float* wo, const float* rn, int64_t count, float* out) {
  for (int64_t i = 0; i < count; i++) {
    hbrdf hb = unpack_brdf(&b[i]);
    v2    r = {rn[2 * i], rn[2 * i + 1]};
    v3    w = hair_sample(&hb, ld3(wo, i), r);
    memcpy(out + 3 * i, &w, 12);
  }
}
void yho_hair_pdf_batch(const yh_hair_brdf* b, const float* wo, const float* wi, int64_t count, float* out) {
  for (int64_t i = 0; i < count; i++) {
    hbrdf hb = unpack_brdf(&b[i]);
    out[i]   = hair_pdf(&hb, ld3(wo, i), ld3(wi, i));
  }
}

/* sample_sphere, ym:4847-4852 */
static v3 sample_sphere(v2 ruv) {
  float z = 2 * ruv.y - 1, r = sqrtf(clampr(1 - z * z, 0.0f, 1.0f)), phi = 2 * PIF * ruv.x;
  return V(r * cosf(phi), r * sinf(phi), z);
}

/* The four statistical self-tests, ext:555-693.  0 = passed, 1 = failed. */
static yh_hair_material selftest_material(float bm, float bn) {
  yh_hair_material m;
  memset(&m, 0, sizeof m);
  m.beta_m = bm, m.beta_n = bn, m.alpha = 0, m.eta = 1.55f;
  return m;
}
int yho_hair_selftest(int which) {
  rng_t rng = rng_make(199382389514ULL, 1);
  v3    nz = V(0, 0, 1), tx = V(1, 0, 0);
  if (which == 0 || which == 1) { /* white_furnace_test :555-587, white_furnace_sampled_test :589-619 */
    v3 wo = sample_sphere(rng_2f(&rng));
    for (float bm = 0.1f; bm < 1.0f; bm += 0.2f)
      for (float bn = 0.1f; bn < 1.0f; bn += 0.2f) {
        v3  sum = V(0, 0, 0);
        int count = 300000;
        for (int i = 0; i < count; i++) {
          float h = rng_1f(&rng);
          if (which == 0 && h == 0) h += FLT_EPSF;
          yh_hair_material m = selftest_material(bm, bn);
          hbrdf b = hair_brdf_make(&m, h, nz, tx);
          if (which == 0) {
            v3 wi = sample_sphere(rng_2f(&rng));
            sum   = vadd(sum, hair_eval(&b, wo, wi));
          } else {
            v3    wi = hair_sample(&b, wo, rng_2f(&rng));
            v3    f = hair_eval(&b, wo, wi);
            float pdf = hair_pdf(&b, wo, wi);
            if (pdf > 0) sum = vadd(sum, vds(f, pdf));
          }
        }
        float avg = which == 0 ? lum(sum) / (count * (1 / (4 * PIF))) : lum(sum) / (count);
        float lo = which == 0 ? 0.95f : 0.99f, hi = which == 0 ? 1.05f : 1.01f;
        if (!(avg >= lo && avg <= hi)) return 1;
      }
    return 0;
  }
  if (which == 2) { /* sampling_weights_test :621-652 */
    for (float bm = 0.1f; bm < 1.0f; bm += 0.2f)
      for (float bn = 0.4f; bn < 1.0f; bn += 0.2f)
        for (int i = 0; i < 10000; i++) {
          float h = rng_1f(&rng);
          yh_hair_material m = selftest_material(bm, bn);
          hbrdf b = hair_brdf_make(&m, h, nz, tx);
          v3    wo = sample_sphere(rng_2f(&rng));
          v3    wi = hair_sample(&b, wo, rng_2f(&rng));
          v3    f = hair_eval(&b, wo, wi);
          float pdf = hair_pdf(&b, wo, wi);
          if (pdf > 0 && !(lum(f) / pdf >= 0.999f && lum(f) / pdf <= 1.001f)) return 1;
        }
    return 0;
  }
  if (which == 3) { /* sampling_consistency_test :654-693 */
    for (float bm = 0.2f; bm < 1.0f; bm += 0.2f)
      for (float bn = 0.4f; bn < 1.0f; bn += 0.2f) {
        const int count = 64 * 1024;
        v3 wo = sample_sphere(rng_2f(&rng)), fi = V(0, 0, 0), fu = V(0, 0, 0);
        for (int i = 0; i < count; i++) {
          float h = rng_1f(&rng);
          yh_hair_material m = selftest_material(bm, bn);
          hbrdf b = hair_brdf_make(&m, h, nz, tx);
          v2    u = rng_2f(&rng);
          v3    wi = hair_sample(&b, wo, u);
          v3    f = hair_eval(&b, wo, wi);
          float pdf = hair_pdf(&b, wo, wi);
          if (pdf > 0) fi = vadd(fi, vds(vs(f, wi.z * wi.z), count * pdf));
          wi = sample_sphere(u);
          fu = vadd(fu, vds(vs(hair_eval(&b, wo, wi), wi.z * wi.z), count * (1 / (4 * PIF))));
        }
        float err = absr(lum(fi) - lum(fu)) / lum(fu);
        if (err >= 0.05f) return 1;
      }
    return 0;
  }
  return 2;
}

void yho_rng_floats(uint64_t seed, uint64_t seq, int n, float* out) {
  rng_t r = rng_make(seed, seq);
  for (int i = 0; i < n; i++) out[i] = rng_1f(&r);
}
void yho_libm_batch(int fn, const float* a, const float* b, int64_t n, float* out) {
  for (int64_t i = 0; i < n; i++) {
    switch (fn) {
      case 0: out[i] = expf(a[i]); break;
      case 1: out[i] = logf(a[i]); break;
      case 2: out[i] = sinf(a[i]); break;
      case 3: out[i] = cosf(a[i]); break;
      case 4: out[i] = atan2f(a[i], b[i]); break;
      case 5: out[i] = acosf(a[i]); break;
      case 6: out[i] = (float)asin((double)a[i]); break;
      case 7: out[i] = (float)(1.0 / (sinh((double)(1 / a[i])) * 2 * a[i])); break;
      case 8: out[i] = fmodf(a[i], b[i]); break;
      case 9: case 10: case 11: { /* the double denominator of mp() (ye:205-206) as an exact three-float expansion */
        double den = sinh((double)(1 / a[i])) * 2 * a[i];
        float  hi = (float)den, mid = (float)(den - (double)hi), lo = (float)(den - (double)hi - (double)mid);
        out[i] = fn == 9 ? hi : (fn == 10 ? mid : lo);
      } break;
      default: out[i] = 0;
    }
  }
}
int yho_hardware_concurrency(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* ========================================================= ray primitives == */
typedef struct { v3 o, d; float tmin, tmax; } ray_t;

static int hit_point(const ray_t* ray, v3 p, float r, v2* uv, float* dist) { /* intersect_point, ym:3405-3423 */
  v3    w = vsub(p, ray->o);
  float t = dot3(w, ray->d) / dot3(ray->d, ray->d);
  if (t < ray->tmin || t > ray->tmax) return 0;
  v3 rp = vadd(ray->o, vs(ray->d, t));
  v3 prp = vsub(p, rp);
  if (dot3(prp, prp) > r * r) return 0;
  uv->x = 0, uv->y = 0;
  *dist = t;
  return 1;
}
static int hit_line(const ray_t* ray, v3 p0, v3 p1, float r0, float r1, v2* uv, float* dist) { /* ym:3426-3469 */
  v3    u = ray->d, v = vsub(p1, p0), w = vsub(ray->o, p0);
  float a = dot3(u, u), b = dot3(u, v), c = dot3(v, v), d = dot3(u, w), e = dot3(v, w);
  float det = a * c - b * b;
  if (det == 0) return 0;
  float t = (b * e - c * d) / det, s = (a * e - b * d) / det;
  if (t < ray->tmin || t > ray->tmax) return 0;
  s = clampr(s, (float)0, (float)1);
  v3    pr = vadd(ray->o, vs(ray->d, t)), pl = vadd(p0, vs(vsub(p1, p0), s)), prl = vsub(pr, pl);
  float d2 = dot3(prl, prl), r = r0 * (1 - s) + r1 * s;
  if (d2 > r * r) return 0;
  uv->x = s, uv->y = sqrtf(d2) / r;
  *dist = t;
  return 1;
}
static int hit_triangle(const ray_t* ray, v3 p0, v3 p1, v3 p2, v2* uv, float* dist) { /* ym:3472-3505 */
  v3    e1 = vsub(p1, p0), e2 = vsub(p2, p0), pvec = cross3(ray->d, e2);
  float det = dot3(e1, pvec);
  if (det == 0) return 0;
  float inv = 1.0f / det;
  v3    tvec = vsub(ray->o, p0);
  float u = dot3(tvec, pvec) * inv;
  if (u < 0 || u > 1) return 0;
  v3    qvec = cross3(tvec, e1);
  float v = dot3(ray->d, qvec) * inv;
  if (v < 0 || u + v > 1) return 0;
  float t = dot3(e2, qvec) * inv;
  if (t < ray->tmin || t > ray->tmax) return 0;
  uv->x = u, uv->y = v;
  *dist = t;
  return 1;
}
static int hit_box(const ray_t* ray, v3 dinv, const box3* b) { /* ym:3544-3554 */
  v3    a = vmul(vsub(b->lo, ray->o), dinv), c = vmul(vsub(b->hi, ray->o), dinv);
  v3    tmn = vmin(a, c), tmx = vmax(a, c);
  float t0 = fmaxr(max3(tmn), ray->tmin);
  float t1 = fminr(fminr(fminr(tmx.x, tmx.y), tmx.z), ray->tmax);
  t1 *= 1.00000024f;
  return t0 <= t1;
}

/* ==================================================================== BVH == */
typedef struct { box3 box; int start; short num; char internal, axis; } node_t; /* pt header :243-249 */
typedef struct { node_t* nodes; int num_nodes; int* prims; int num_prims; } tree_t;
typedef struct { box3 box; v3 center; int prim; } bprim_t;                       /* pt:557-561 */
static const box3 kInvalidBox = {{FLT_MAXF, FLT_MAXF, FLT_MAXF}, {-FLT_MAXF, -FLT_MAXF, -FLT_MAXF}};
static box3 box_merge(box3 a, box3 b) { box3 r = {vmin(a.lo, b.lo), vmax(a.hi, b.hi)}; return r; }
static box3 box_add(box3 a, v3 p) { box3 r = {vmin(a.lo, p), vmax(a.hi, p)}; return r; }
static v3   box_center(box3 a) { return vds(vadd(a.lo, a.hi), 2); } /* ym:3008 */

/* std::partition as libstdc++ runs it on bidirectional iterators (the order of
 * the primitives inside each half is part of the reference's leaf order). */
static int partition_less(bprim_t* p, int first, int last, int axis, float middle) {
  for (;;) {
    for (;;) {
      if (first == last) return first;
      if (comp(p[first].center, axis) < middle) first++;
      else break;
    }
    last--;
    for (;;) {
      if (first == last) return first;
      if (!(comp(p[last].center, axis) < middle)) last--;
      else break;
    }
    bprim_t t = p[first];
    p[first] = p[last], p[last] = t;
    first++;
  }
}
static void split_middle(bprim_t* p, int start, int end, int* mid, int* axis) { /* pt:564-595 */
  *axis = 0, *mid = (start + end) / 2;
  box3 cb = kInvalidBox;
  for (int i = start; i < end; i++) cb = box_add(cb, p[i].center);
  v3 cs = vsub(cb.hi, cb.lo);
  if (cs.x == 0 && cs.y == 0 && cs.z == 0) return;
  if (cs.x >= cs.y && cs.x >= cs.z) *axis = 0;
  if (cs.y >= cs.x && cs.y >= cs.z) *axis = 1;
  if (cs.z >= cs.x && cs.z >= cs.y) *axis = 2;
  *mid = partition_less(p, start, end, *axis, comp(box_center(cb), *axis));
  if (*mid == start || *mid == end) *mid = (start + end) / 2;
}
static tree_t build_tree(bprim_t* p, int n) { /* pt:601-650: breadth-first, leaves of <= 4 */
  tree_t t = {NULL, 0, NULL, n};
  if (n == 0) return t; /* nodes stays empty: "bvh->nodes.empty()" */
  t.nodes = (node_t*)calloc((size_t)n * 2 + 1, sizeof(node_t));
  int (*queue)[3] = malloc(sizeof(int[3]) * ((size_t)n * 2 + 1));
  int qh = 0, qt = 0;
  queue[qt][0] = 0, queue[qt][1] = 0, queue[qt][2] = n, qt++;
  t.num_nodes = 1;
  while (qh < qt) {
    int id = queue[qh][0], start = queue[qh][1], end = queue[qh][2];
    qh++;
    node_t* nd = &t.nodes[id];
    nd->box = kInvalidBox;
    for (int i = start; i < end; i++) nd->box = box_merge(nd->box, p[i].box);
    if (end - start > 4) {
      int mid, axis;
      split_middle(p, start, end, &mid, &axis);
      nd->internal = 1, nd->axis = (char)axis, nd->num = 2, nd->start = t.num_nodes;
      t.num_nodes += 2;
      queue[qt][0] = nd->start, queue[qt][1] = start, queue[qt][2] = mid, qt++;
      queue[qt][0] = nd->start + 1, queue[qt][1] = mid, queue[qt][2] = end, qt++;
    } else {
      nd->internal = 0, nd->num = (short)(end - start), nd->start = start;
    }
  }
  free(queue);
  t.prims = (int*)malloc(sizeof(int) * (size_t)n);
  for (int i = 0; i < n; i++) t.prims[i] = p[i].prim;
  return t;
}

/* ================================================================== scene == */
typedef struct {
  v3 *pos, *nrm;
  v2* tex;
  float* rad;
  int *lines, *tris, *points;
  int   npos, nnrm, ntex, nlines, ntris, npoints; /* one kind per shape */
  tree_t bvh;
} shape_t;
typedef struct {
  v3 emission, color; float opacity; int thin; yh_hair_material hair;
  float specular, metallic, roughness, ior, transmission; /* surface lobes, pt header :293-307 */
  v3    scattering; float scanisotropy, trdepth;          /* volume behind a refractive surface */
  int   emission_tex, color_tex, specular_tex, metallic_tex, roughness_tex, transmission_tex, opacity_tex, scattering_tex, normal_tex;
} mat_t;
typedef struct { fr3 frame; int shape, material; } obj_t;
typedef struct { v3* px; int w, h; } tex_t;
typedef struct { fr3 frame; v3 emission; int tex; } env_t;
typedef struct { int object, env; float* cdf; int ncdf; } light_t; /* pt header :383-387 */
typedef struct {
  shape_t* shapes; mat_t* mats; obj_t* objs; tex_t* texs; env_t* envs; light_t* lights;
  int      nshapes, nmats, nobjs, ntexs, nenvs, nlights;
  tree_t   bvh;
} scene_t;

static void* dup_mem(const void* p, size_t bytes) {
  void* r = malloc(bytes ? bytes : 1);
  if (bytes) memcpy(r, p, bytes);
  return r;
}
static void init_shape_bvh(shape_t* s) { /* pt:721-752 */
  int      n = s->npoints ? s->npoints : s->nlines ? s->nlines : s->ntris;
  bprim_t* p = (bprim_t*)malloc(sizeof(bprim_t) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; i++) {
    if (s->npoints) { /* point_bounds(p, r), ym:3030-3033, pt:715-721 */
      int a = s->points[i];
      v3  lo = vsubs(s->pos[a], s->rad[a]), hi = vadds(s->pos[a], s->rad[a]);
      p[i].box.lo = vmin(lo, hi), p[i].box.hi = vmax(lo, hi);
    } else if (s->nlines) { /* line_bounds, ym:3037 */
      int a = s->lines[2 * i], b = s->lines[2 * i + 1];
      p[i].box.lo = vmin(vsubs(s->pos[a], s->rad[a]), vsubs(s->pos[b], s->rad[b]));
      p[i].box.hi = vmax(vadds(s->pos[a], s->rad[a]), vadds(s->pos[b], s->rad[b]));
    } else { /* triangle_bounds, ym:3041 */
      v3 a = s->pos[s->tris[3 * i]], b = s->pos[s->tris[3 * i + 1]], c = s->pos[s->tris[3 * i + 2]];
      p[i].box.lo = vmin(a, vmin(b, c)), p[i].box.hi = vmax(a, vmax(b, c));
    }
    p[i].center = box_center(p[i].box);
    p[i].prim   = i;
  }
  s->bvh = build_tree(p, n);
  free(p);
}
static box3 transform_box(const fr3* f, box3 b) { /* ym:3172-3182 */
  box3 r = kInvalidBox;
  for (int c = 0; c < 8; c++) {
    v3 corner = V(c & 4 ? b.hi.x : b.lo.x, c & 2 ? b.hi.y : b.lo.y, c & 1 ? b.hi.z : b.lo.z);
    r = box_add(r, xf_point(f, corner));
  }
  return r;
}
static void init_scene_bvh(scene_t* sc) { /* pt:793-814 */
  bprim_t* p = (bprim_t*)malloc(sizeof(bprim_t) * (size_t)(sc->nobjs ? sc->nobjs : 1));
  for (int i = 0; i < sc->nobjs; i++) {
    shape_t* s = &sc->shapes[sc->objs[i].shape];
    p[i].box    = s->bvh.num_nodes == 0 ? kInvalidBox : transform_box(&sc->objs[i].frame, s->bvh.nodes[0].box);
    p[i].center = box_center(p[i].box);
    p[i].prim   = i;
  }
  sc->bvh = build_tree(p, sc->nobjs);
  free(p);
}
static void init_lights(scene_t* sc) { /* pt:1695-1740 */
  sc->lights  = (light_t*)calloc((size_t)(sc->nobjs + sc->nenvs + 1), sizeof(light_t));
  sc->nlights = 0;
  for (int o = 0; o < sc->nobjs; o++) {
    mat_t*   m = &sc->mats[sc->objs[o].material];
    shape_t* s = &sc->shapes[sc->objs[o].shape];
    if (!nonzero3(m->emission) || s->ntris == 0) continue;
    light_t* l = &sc->lights[sc->nlights++];
    l->object = o, l->env = -1, l->ncdf = s->ntris;
    l->cdf = (float*)malloc(sizeof(float) * (size_t)s->ntris);
    for (int i = 0; i < s->ntris; i++) { /* triangle_area, ym:3306 */
      v3 p0 = s->pos[s->tris[3 * i]], p1 = s->pos[s->tris[3 * i + 1]], p2 = s->pos[s->tris[3 * i + 2]];
      l->cdf[i] = len3(cross3(vsub(p1, p0), vsub(p2, p0))) / 2;
      if (i) l->cdf[i] += l->cdf[i - 1];
    }
  }
  for (int e = 0; e < sc->nenvs; e++) {
    env_t* env = &sc->envs[e];
    if (!nonzero3(env->emission)) continue;
    light_t* l = &sc->lights[sc->nlights++];
    l->object = -1, l->env = e, l->cdf = NULL, l->ncdf = 0;
    if (env->tex >= 0) {
      tex_t* t = &sc->texs[env->tex];
      l->ncdf = t->w * t->h;
      l->cdf  = (float*)malloc(sizeof(float) * (size_t)l->ncdf);
      for (int i = 0; i < l->ncdf; i++) {
        float th = ((i / t->w) + 0.5f) * PIF / t->h;
        l->cdf[i] = max3(t->px[i]) * sinf(th);
        if (i) l->cdf[i] += l->cdf[i - 1];
      }
    }
  }
}

void* yho_scene_create(const yh_scene_desc* d) {
  scene_t* sc = (scene_t*)calloc(1, sizeof(scene_t));
  sc->nshapes = d->num_shapes, sc->nmats = d->num_materials, sc->nobjs = d->num_objects;
  sc->ntexs = d->num_textures, sc->nenvs = d->num_environments;
  sc->shapes = (shape_t*)calloc((size_t)sc->nshapes + 1, sizeof(shape_t));
  for (int i = 0; i < sc->nshapes; i++) {
    const yh_shape_desc* s = &d->shapes[i];
    shape_t* t = &sc->shapes[i];
    t->npos = s->num_positions, t->nnrm = s->num_normals, t->nlines = s->num_lines, t->ntris = s->num_triangles;
    t->pos = (v3*)dup_mem(s->positions, 12 * (size_t)t->npos);
    t->nrm = (v3*)dup_mem(s->normals, 12 * (size_t)t->nnrm);
    t->ntex = s->num_texcoords;
    t->tex = (v2*)dup_mem(s->texcoords, 8 * (size_t)t->ntex);
    t->rad = (float*)dup_mem(s->radius, 4 * (size_t)s->num_radius);
    t->lines = (int*)dup_mem(s->lines, 8 * (size_t)t->nlines);
    t->tris = (int*)dup_mem(s->triangles, 12 * (size_t)t->ntris);
    t->npoints = s->num_points;
    t->points = (int*)dup_mem(s->points, 4 * (size_t)t->npoints);
    if (t->nlines) t->ntris = 0; /* the reference tests lines before triangles only after points; keep one kind */
    if (t->npoints) t->nlines = 0, t->ntris = 0;
    init_shape_bvh(t);
  }
  sc->texs = (tex_t*)calloc((size_t)sc->ntexs + 1, sizeof(tex_t));
  for (int i = 0; i < sc->ntexs; i++) {
    sc->texs[i].w = d->textures[i].width, sc->texs[i].h = d->textures[i].height;
    sc->texs[i].px = (v3*)dup_mem(d->textures[i].colorf, 12 * (size_t)sc->texs[i].w * sc->texs[i].h);
  }
  sc->mats = (mat_t*)calloc((size_t)sc->nmats + 1, sizeof(mat_t));
  for (int i = 0; i < sc->nmats; i++) {
    const yh_material_desc* m = &d->materials[i];
    mat_t* t = &sc->mats[i];
    t->emission = V(m->emission[0], m->emission[1], m->emission[2]);
    t->color = V(m->color[0], m->color[1], m->color[2]);
    t->opacity = m->opacity, t->thin = m->thin;
    t->specular = m->specular, t->metallic = m->metallic, t->roughness = m->roughness, t->ior = m->ior;
    t->transmission = m->transmission, t->scanisotropy = m->scanisotropy, t->trdepth = m->trdepth;
    t->scattering = V(m->scattering[0], m->scattering[1], m->scattering[2]);
    t->emission_tex = m->emission_tex, t->color_tex = m->color_tex, t->specular_tex = m->specular_tex;
    t->metallic_tex = m->metallic_tex, t->roughness_tex = m->roughness_tex, t->transmission_tex = m->transmission_tex;
    t->opacity_tex = m->opacity_tex, t->scattering_tex = m->scattering_tex, t->normal_tex = m->normal_tex;
    memcpy(t->hair.sigma_a, m->sigma_a, 12);
    t->hair.beta_m = m->beta_m, t->hair.beta_n = m->beta_n, t->hair.alpha = m->alpha, t->hair.eta = m->eta;
    memcpy(t->hair.color, m->color, 12);
    t->hair.eumelanin = m->eumelanin, t->hair.pheomelanin = m->pheomelanin;
  }
  sc->objs = (obj_t*)calloc((size_t)sc->nobjs + 1, sizeof(obj_t));
  for (int i = 0; i < sc->nobjs; i++) {
    sc->objs[i].frame = frame_from(d->objects[i].frame);
    sc->objs[i].shape = d->objects[i].shape, sc->objs[i].material = d->objects[i].material;
  }
  sc->envs = (env_t*)calloc((size_t)sc->nenvs + 1, sizeof(env_t));
  for (int i = 0; i < sc->nenvs; i++) {
    sc->envs[i].frame = frame_from(d->environments[i].frame);
    sc->envs[i].emission = V(d->environments[i].emission[0], d->environments[i].emission[1], d->environments[i].emission[2]);
    sc->envs[i].tex = d->environments[i].emission_tex;
  }
  init_scene_bvh(sc);
  init_lights(sc);
  return sc;
}
/* shape->bvh->primitives after build_bvh (pt:746-752): the leaf order of the reference's tree */
int yho_shape_order(void* h, int shape, int* order) {
  scene_t* sc = (scene_t*)h;
  if (!sc || shape < 0 || shape >= sc->nshapes) return -1;
  const tree_t* t = &sc->shapes[shape].bvh;
  if (order) memcpy(order, t->prims, sizeof(int) * (size_t)t->num_prims);
  return t->num_prims;
}
void yho_scene_destroy(void* h) {
  scene_t* sc = (scene_t*)h;
  if (!sc) return;
  for (int i = 0; i < sc->nshapes; i++) {
    shape_t* s = &sc->shapes[i];
    free(s->pos), free(s->nrm), free(s->tex), free(s->rad), free(s->lines), free(s->tris), free(s->points), free(s->bvh.nodes), free(s->bvh.prims);
  }
  for (int i = 0; i < sc->ntexs; i++) free(sc->texs[i].px);
  for (int i = 0; i < sc->nlights; i++) free(sc->lights[i].cdf);
  free(sc->shapes), free(sc->texs), free(sc->mats), free(sc->objs), free(sc->envs), free(sc->lights);
  free(sc->bvh.nodes), free(sc->bvh.prims), free(sc);
}

/* ============================================================== traversal == */
static int walk_shape(const shape_t* s, const ray_t* ray_, int* element, v2* uv, float* distance, int find_any) {
  /* pt:848-931: stack[128] DFS, near child first by the sign of the split axis */
  const tree_t* t = &s->bvh;
  if (t->num_nodes == 0) return 0;
  int   stack[128], top = 0, hit = 0;
  ray_t ray = *ray_;
  v3    dinv = V(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);
  int   dsign[3] = {dinv.x < 0, dinv.y < 0, dinv.z < 0};
  stack[top++] = 0;
  while (top) {
    const node_t* nd = &t->nodes[stack[--top]];
    if (!hit_box(&ray, dinv, &nd->box)) continue;
    if (nd->internal) {
      if (dsign[(int)nd->axis]) stack[top++] = nd->start, stack[top++] = nd->start + 1;
      else stack[top++] = nd->start + 1, stack[top++] = nd->start;
    } else if (s->npoints) { /* pt:894-903 */
      for (int i = nd->start; i < nd->start + nd->num; i++) {
        int e = t->prims[i], a = s->points[e];
        if (hit_point(&ray, s->pos[a], s->rad[a], uv, distance)) hit = 1, *element = e, ray.tmax = *distance;
      }
    } else if (s->nlines) {
      for (int i = nd->start; i < nd->start + nd->num; i++) {
        int e = t->prims[i], a = s->lines[2 * e], b = s->lines[2 * e + 1];
        if (hit_line(&ray, s->pos[a], s->pos[b], s->rad[a], s->rad[b], uv, distance)) hit = 1, *element = e, ray.tmax = *distance;
      }
    } else {
      for (int i = nd->start; i < nd->start + nd->num; i++) {
        int e = t->prims[i];
        if (hit_triangle(&ray, s->pos[s->tris[3 * e]], s->pos[s->tris[3 * e + 1]], s->pos[s->tris[3 * e + 2]], uv, distance))
          hit = 1, *element = e, ray.tmax = *distance;
      }
    }
    if (find_any && hit) return hit;
  }
  return hit;
}
static ray_t ray_to_object(const obj_t* ob, const ray_t* ray) { /* transform_ray(inverse(frame, true), ray), pt:1012 */
  fr3   inv = inverse_frame(&ob->frame, 1);
  ray_t r = {xf_point(&inv, ray->o), xf_vector(&inv, ray->d), ray->tmin, ray->tmax};
  return r;
}
static int walk_scene(const scene_t* sc, const ray_t* ray_, int* object, int* element, v2* uv, float* distance, int find_any) {
  /* pt:960-1028 */
  const tree_t* t = &sc->bvh;
  if (t->num_nodes == 0) return 0;
  int   stack[128], top = 0, hit = 0;
  ray_t ray = *ray_;
  v3    dinv = V(1 / ray.d.x, 1 / ray.d.y, 1 / ray.d.z);
  int   dsign[3] = {dinv.x < 0, dinv.y < 0, dinv.z < 0};
  stack[top++] = 0;
  while (top) {
    const node_t* nd = &t->nodes[stack[--top]];
    if (!hit_box(&ray, dinv, &nd->box)) continue;
    if (nd->internal) {
      if (dsign[(int)nd->axis]) stack[top++] = nd->start, stack[top++] = nd->start + 1;
      else stack[top++] = nd->start + 1, stack[top++] = nd->start;
    } else {
      for (int i = nd->start; i < nd->start + nd->num; i++) {
        const obj_t* ob = &sc->objs[t->prims[i]];
        ray_t inv = ray_to_object(ob, &ray);
        if (walk_shape(&sc->shapes[ob->shape], &inv, element, uv, distance, find_any))
          hit = 1, *object = t->prims[i], ray.tmax = *distance;
      }
    }
    if (find_any && hit) return hit;
  }
  return hit;
}

typedef struct { const scene_t* sc; const yh_ray* rays; yh_hit* hits; int64_t b, e; int find_any; } isect_job;
static void* isect_worker(void* arg) {
  isect_job* j = (isect_job*)arg;
  for (int64_t i = j->b; i < j->e; i++) {
    const yh_ray* r = &j->rays[i];
    ray_t  ray = {V(r->o[0], r->o[1], r->o[2]), V(r->d[0], r->d[1], r->d[2]), r->tmin, r->tmax};
    yh_hit h = {-1, -1, {0, 0}, 0, 0}; /* intersection3f{} defaults, pt header :441-447 */
    v2     uv = {0, 0};
    h.hit = walk_scene(j->sc, &ray, &h.object, &h.element, &uv, &h.distance, j->find_any);
    h.uv[0] = uv.x, h.uv[1] = uv.y;
    j->hits[i] = h;
  }
  return NULL;
}
void yho_intersect_batch(void* h, const yh_ray* rays, int64_t n, int find_any, yh_hit* hits, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  isect_job jobs[256];
  for (int t = 0; t < nthreads; t++) {
    isect_job j = {(const scene_t*)h, rays, hits, n * t / nthreads, n * (t + 1) / nthreads, find_any};
    jobs[t] = j;
    if (nthreads == 1) isect_worker(&jobs[t]);
    else pthread_create(&th[t], NULL, isect_worker, &jobs[t]);
  }
  if (nthreads > 1)
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ================================================================ shading == */
static v3 eval_position(const scene_t* sc, int object, int element, v2 uv) { /* pt:232-250 */
  const obj_t* ob = &sc->objs[object];
  const shape_t* s = &sc->shapes[ob->shape];
  if (s->ntris) {
    v3 p0 = s->pos[s->tris[3 * element]], p1 = s->pos[s->tris[3 * element + 1]], p2 = s->pos[s->tris[3 * element + 2]];
    return xf_point(&ob->frame, vadd(vadd(vs(p0, 1 - uv.x - uv.y), vs(p1, uv.x)), vs(p2, uv.y)));
  } else if (s->nlines) {
    return xf_point(&ob->frame, lerp3(s->pos[s->lines[2 * element]], s->pos[s->lines[2 * element + 1]], uv.x));
  } else if (s->npoints) {
    return xf_point(&ob->frame, s->pos[s->points[element]]);
  }
  return V(0, 0, 0);
}
static v3 eval_element_normal(const scene_t* sc, int object, int element) { /* pt:253-269 */
  const obj_t* ob = &sc->objs[object];
  const shape_t* s = &sc->shapes[ob->shape];
  if (s->ntris) {
    v3 p0 = s->pos[s->tris[3 * element]], p1 = s->pos[s->tris[3 * element + 1]], p2 = s->pos[s->tris[3 * element + 2]];
    return xf_direction(&ob->frame, norm3(cross3(vsub(p1, p0), vsub(p2, p0))));
  } else if (s->nlines) {
    return xf_direction(&ob->frame, norm3(vsub(s->pos[s->lines[2 * element + 1]], s->pos[s->lines[2 * element]])));
  } else if (s->npoints) {
    return V(0, 0, 1);
  }
  return V(0, 0, 0);
}
static v3 eval_normal(const scene_t* sc, int object, int element, v2 uv) { /* pt:272-292 */
  const obj_t* ob = &sc->objs[object];
  const shape_t* s = &sc->shapes[ob->shape];
  if (s->nnrm == 0) return eval_element_normal(sc, object, element);
  if (s->ntris) {
    v3 n0 = s->nrm[s->tris[3 * element]], n1 = s->nrm[s->tris[3 * element + 1]], n2 = s->nrm[s->tris[3 * element + 2]];
    return xf_direction(&ob->frame, norm3(vadd(vadd(vs(n0, 1 - uv.x - uv.y), vs(n1, uv.x)), vs(n2, uv.y))));
  } else if (s->nlines) {
    return xf_direction(&ob->frame, norm3(lerp3(s->nrm[s->lines[2 * element]], s->nrm[s->lines[2 * element + 1]], uv.x)));
  } else if (s->npoints) {
    return xf_direction(&ob->frame, norm3(s->nrm[s->points[element]]));
  }
  return V(0, 0, 0);
}
static v3 eval_texture(const scene_t* sc, int tex, v2 uv) { /* pt:166-200, float RGB, bilinear, tiled */
  if (tex < 0) return V(1, 1, 1);
  const tex_t* t = &sc->texs[tex];
  float s = fmodf(uv.x, 1.0f) * t->w;
  if (s < 0) s += t->w;
  float tt = fmodf(uv.y, 1.0f) * t->h;
  if (tt < 0) tt += t->h;
  int   i = clampi((int)s, 0, t->w - 1), j = clampi((int)tt, 0, t->h - 1);
  int   ii = (i + 1) % t->w, jj = (j + 1) % t->h;
  float u = s - i, v = tt - j;
#define TX(a, b) t->px[(size_t)(b) * t->w + (a)]
  v3 r = vadd(vadd(vadd(vs(vs(TX(i, j), 1 - u), 1 - v), vs(vs(TX(i, jj), 1 - u), v)), vs(vs(TX(ii, j), u), 1 - v)),
      vs(vs(TX(ii, jj), u), v));
#undef TX
  return r;
}
static v2 env_texcoord(const env_t* env, v3 dir) {
  fr3 inv = inverse_frame(&env->frame, 0);
  v3  wl = xf_direction(&inv, dir);
  v2  tc = {atan2f(wl.z, wl.x) / (2 * PIF), acosf(clampr(wl.y, -1.0f, 1.0f)) / PIF};
  if (tc.x < 0) tc.x += 1;
  return tc;
}
static v3 eval_environment(const scene_t* sc, v3 dir) { /* pt:536-547 */
  v3 em = V(0, 0, 0);
  for (int e = 0; e < sc->nenvs; e++) {
    v2 tc = env_texcoord(&sc->envs[e], dir);
    em    = vadd(em, vmul(sc->envs[e].emission, eval_texture(sc, sc->envs[e].tex, tc)));
  }
  return em;
}
static int sample_discrete_cdf(const float* cdf, int n, float r) { /* ym:4953-4958, std::upper_bound */
  float back = cdf[n - 1];
  r = clampr(r * back, (float)0, back - (float)0.00001);
  int lo = 0, len = n;
  while (len > 0) {
    int half = len >> 1;
    if (!(r < cdf[lo + half])) lo += half + 1, len -= half + 1;
    else len = half;
  }
  return clampi(lo, 0, n - 1);
}
static v3 sample_lights(const scene_t* sc, v3 position, float rl, float rel, v2 ruv) { /* pt:1283-1308 */
  int            id = clampi((int)(rl * sc->nlights), 0, sc->nlights - 1);
  const light_t* l = &sc->lights[id];
  if (l->object >= 0) {
    int element = sample_discrete_cdf(l->cdf, l->ncdf, rel);
    v2  uv = {1 - sqrtf(ruv.x), ruv.y * sqrtf(ruv.x)};
    return norm3(vsub(eval_position(sc, l->object, element, uv), position));
  }
  const env_t* env = &sc->envs[l->env];
  if (env->tex >= 0) {
    const tex_t* t = &sc->texs[env->tex];
    int idx = sample_discrete_cdf(l->cdf, l->ncdf, rel);
    v2  uv = {(idx % t->w + 0.5f) / t->w, (idx / t->w + 0.5f) / t->h};
    return xf_direction(&env->frame, V(cosf(uv.x * 2 * PIF) * sinf(uv.y * PIF), cosf(uv.y * PIF), sinf(uv.x * 2 * PIF) * sinf(uv.y * PIF)));
  }
  return sample_sphere(ruv);
}
static float sample_lights_pdf(const scene_t* sc, v3 position, v3 direction) { /* pt:1311-1358 */
  float pdf = 0.0f;
  for (int i = 0; i < sc->nlights; i++) {
    const light_t* l = &sc->lights[i];
    if (l->object >= 0) {
      float lpdf = 0.0f;
      v3    next = position;
      for (int bounce = 0; bounce < 100; bounce++) {
        const obj_t* ob = &sc->objs[l->object];
        ray_t ray = {next, direction, RAY_EPS, FLT_MAXF};
        ray_t inv = ray_to_object(ob, &ray);
        int   element = -1;
        v2    uv = {0, 0};
        float dist = 0;
        if (!walk_shape(&sc->shapes[ob->shape], &inv, &element, &uv, &dist, 0)) break;
        v3    lp = eval_position(sc, l->object, element, uv), ln = eval_element_normal(sc, l->object, element);
        float area = l->cdf[l->ncdf - 1];
        v3    dp = vsub(lp, position);
        lpdf += dot3(dp, dp) / (absr(dot3(ln, direction)) * area);
        next = vadd(lp, vs(direction, 1e-3f));
      }
      pdf += lpdf;
    } else {
      const env_t* env = &sc->envs[l->env];
      if (env->tex >= 0) {
        const tex_t* t = &sc->texs[env->tex];
        v2    tc = env_texcoord(env, direction);
        int   ti = clampi((int)(tc.x * t->w), 0, t->w - 1), tj = clampi((int)(tc.y * t->h), 0, t->h - 1);
        int   idx = tj * t->w + ti;
        float prob = (idx == 0 ? l->cdf[0] : l->cdf[idx] - l->cdf[idx - 1]) / l->cdf[l->ncdf - 1];
        float angle = (2 * PIF / t->w) * (PIF / t->h) * sinf(PIF * (tj + 0.5f) / t->h);
        pdf += prob / angle;
      } else {
        pdf += 1 / (4 * PIF);
      }
    }
  }
  pdf *= (float)1 / (float)sc->nlights;
  return pdf;
}

/* ================================================================ renderer == */
typedef struct { float accum[4]; int samples; rng_t rng; uint64_t seq; } pixel_t; /* pt header :419-423 (+ seq: init_state's sequence id) */
typedef struct {
  const scene_t* sc;
  fr3 cam_frame; float lens, film[2], focus, aperture;
  yh_trace_params params;
  int w, h;
  pixel_t* px;
  float*   render;
  volatile int next_row;
  /* NOT in the reference: the sample-sharded "counter mode" of the CUDA library
   * (include/yocto_hair_b200.h, yh_state_set_sample_shard), restated here so the
   * device path can be checked bit for bit.  Sample s of a pixel draws from
   * make_rng(seed, seq_pixel + (s << 31)); this state renders s = srank + k * sworld. */
  int counter, srank, sworld;
} state_t;

static ray_t camera_ray(const state_t* st, int i, int j, v2 puv, v2 luv) { /* pt:211-229 */
  v2 uv = {((float)i + puv.x) / (float)st->w, ((float)j + puv.y) / (float)st->h};
  float r = sqrtf(luv.y), phi = 2 * PIF * luv.x; /* sample_disk, ym:4897-4901 */
  v2 lens = {cosf(phi) * r, sinf(phi) * r};
  v3 q = V(st->film[0] * (0.5f - uv.x), st->film[1] * (uv.y - 0.5f), st->lens);
  v3 dc = vneg(norm3(q));
  v3 e = V(lens.x * st->aperture / 2, lens.y * st->aperture / 2, 0);
  v3 p = vds(vs(dc, st->focus), absr(dc.z));
  v3 d = norm3(vsub(p, e));
  ray_t ray = {xf_point(&st->cam_frame, e), xf_direction(&st->cam_frame, d), RAY_EPS, FLT_MAXF};
  return ray;
}

/* trace_path pt:1380-1511, trace_naive pt:1514-1573, trace_eyelight pt:1576-1633 and trace_normal
 * pt:1636-1651 (selected by params->shader as get_trace_shader_func does, pt:1662-1673), for the
 * lobes of SURVEY.md 8a (hair on lines, diffuse on triangles). */

static v2 eval_texcoord(const scene_t* sc, int object, int element, v2 uv) { /* pt:294-309 */
  const shape_t* s = &sc->shapes[sc->objs[object].shape];
  if (s->ntex == 0) return uv;
  v2 r;
  if (s->ntris) {
    v2    t0 = s->tex[s->tris[3 * element]], t1 = s->tex[s->tris[3 * element + 1]], t2 = s->tex[s->tris[3 * element + 2]];
    float w = 1 - uv.x - uv.y;
    r.x = t0.x * w + t1.x * uv.x + t2.x * uv.y, r.y = t0.y * w + t1.y * uv.x + t2.y * uv.y;
    return r;
  }
  if (s->npoints) return s->tex[s->points[element]];
  v2 t0 = s->tex[s->lines[2 * element]], t1 = s->tex[s->lines[2 * element + 1]];
  r.x = t0.x * (1 - uv.x) + t1.x * uv.x, r.y = t0.y * (1 - uv.x) + t1.y * uv.x;
  return r;
}
static v3 eval_normalmap(const scene_t* sc, int object, int element, v2 uv, int normal_tex) { /* pt:312-347 */
  const obj_t*   ob = &sc->objs[object];
  const shape_t* s = &sc->shapes[ob->shape];
  v3 normal = eval_normal(sc, object, element, uv);
  v2 texcoord = eval_texcoord(sc, object, element, uv);
  if (!s->ntris) return normal;
  v3 t2x = vs(eval_texture(sc, normal_tex, texcoord), 2);
  v3 nm = V(-1 + t2x.x, -1 + t2x.y, -1 + t2x.z);
  v3 tu = V(0, 0, 0), tv = V(0, 0, 0);
  if (s->ntex) { /* triangle_tangents_fromuv, ym:3360-3383 */
    int a = s->tris[3 * element], b = s->tris[3 * element + 1], c = s->tris[3 * element + 2];
    v3  p = vsub(s->pos[b], s->pos[a]), q = vsub(s->pos[c], s->pos[a]);
    v2  uv0 = s->tex[a], uv1 = s->tex[b], uv2 = s->tex[c];
    float sx = uv1.x - uv0.x, sy = uv2.x - uv0.x, tx = uv1.y - uv0.y, ty = uv2.y - uv0.y;
    float d = sx * ty - sy * tx;
    v3 eu = V(1, 0, 0), ev = V(0, 1, 0);
    if (d != 0) {
      eu = vds(V(ty * p.x - tx * q.x, ty * p.y - tx * q.y, ty * p.z - tx * q.z), d);
      ev = vds(V(sx * q.x - sy * p.x, sx * q.y - sy * p.y, sx * q.z - sy * p.z), d);
    }
    tu = xf_direction(&ob->frame, eu), tv = xf_direction(&ob->frame, ev);
  }
  v3  fx = orthonorm(tu, normal);
  v3  fy = norm3(cross3(normal, fx));
  int flip_v = dot3(fy, tv) < 0;
  nm.y *= flip_v ? 1.0f : -1.0f;
  return norm3(vadd(vadd(vs(fx, nm.x), vs(fy, nm.y)), vs(normal, nm.z)));
}

/* ---- surface lobes, refraction, volumes (ym:4215-4812; pt:405-531, :1069-1280, :1360-1378) ---------------------- */
typedef struct {
  v3 diffuse, specular, metal, transmission, refraction;
  float roughness, ior;
  v3 meta, metak;
  float diffuse_pdf, specular_pdf, metal_pdf, transmission_pdf, refraction_pdf;
} sbrdf; /* pt:372-393 without the hair block */
typedef struct { v3 density, scatter; float anisotropy; } vsdf_t; /* pt:498-502 */

static float fresnel_dielectric(float eta, v3 normal, v3 outgoing) { /* ym:4215-4235 */
  float cosw = absr(dot3(normal, outgoing));
  float sin2 = 1 - cosw * cosw, eta2 = eta * eta;
  float cos2t = 1 - sin2 / eta2;
  if (cos2t < 0) return 1;
  float t0 = sqrtf(cos2t), t1 = eta * t0, t2 = eta * cosw;
  float rs = (cosw - t1) / (cosw + t1), rp = (t0 - t2) / (t0 + t2);
  return (rs * rs + rp * rp) / 2;
}
static float conductor1(float e, float k, float cosw, float cos2, float sin2) {
  float eta2 = e * e, etak2 = k * k;
  float t0 = eta2 - etak2 - sin2;
  float a2plusb2 = sqrtf(t0 * t0 + 4 * eta2 * etak2);
  float t1 = a2plusb2 + cos2;
  float a  = sqrtf((a2plusb2 + t0) / 2);
  float t2 = 2 * a * cosw;
  float rs = (t1 - t2) / (t1 + t2);
  float t3 = cos2 * a2plusb2 + sin2 * sin2;
  float t4 = t2 * sin2;
  float rp = rs * (t3 - t4) / (t3 + t4);
  return (rp + rs) / 2;
}
static v3 fresnel_conductor(v3 eta, v3 etak, v3 normal, v3 outgoing) { /* ym:4238-4263, per component */
  float cosw = dot3(normal, outgoing);
  if (cosw <= 0) return V(0, 0, 0);
  cosw = clampr(cosw, -1.0f, 1.0f);
  float cos2 = cosw * cosw, sin2 = clampr(1 - cos2, 0.0f, 1.0f);
  return V(conductor1(eta.x, etak.x, cosw, cos2, sin2), conductor1(eta.y, etak.y, cosw, cos2, sin2), conductor1(eta.z, etak.z, cosw, cos2, sin2));
}
static float refl_to_eta1(float r) { r = clampr(r, 0.0f, 0.99f); return (1 + sqrtf(r)) / (1 - sqrtf(r)); }
static v3 reflectivity_to_eta(v3 r) { return V(refl_to_eta1(r.x), refl_to_eta1(r.y), refl_to_eta1(r.z)); } /* ym:4270-4273 */
static v3 reflect3(v3 w, v3 n) { return vadd(vneg(w), vs(n, 2 * dot3(n, w))); }                           /* ym:2059-2061 */
static v3 refract3(v3 w, v3 n, float inv_eta) {                                                           /* ym:2062-2067 */
  float cosine = dot3(n, w);
  float k = 1 + inv_eta * inv_eta * (cosine * cosine - 1);
  if (k < 0) return V(0, 0, 0);
  return vadd(vs(vneg(w), inv_eta), vs(n, inv_eta * cosine - sqrtf(k)));
}
static float mf_distribution(float roughness, v3 normal, v3 halfway) { /* ym:4307-4322, ggx */
  float cosine = dot3(normal, halfway);
  if (cosine <= 0) return 0;
  float r2 = roughness * roughness, c2 = cosine * cosine;
  return r2 / (PIF * (c2 * r2 + 1 - c2) * (c2 * r2 + 1 - c2));
}
static float mf_shadowing1(float roughness, v3 normal, v3 halfway, v3 direction) { /* ym:4325-4343 */
  float cosine = dot3(normal, direction), cosineh = dot3(halfway, direction);
  if (cosine * cosineh <= 0) return 0;
  float r2 = roughness * roughness, c2 = cosine * cosine;
  return 2 * absr(cosine) / (absr(cosine) + sqrtf(c2 - r2 * c2 + r2));
}
static float mf_shadowing(float roughness, v3 normal, v3 halfway, v3 outgoing, v3 incoming) { /* ym:4346-4351 */
  return mf_shadowing1(roughness, normal, halfway, outgoing) * mf_shadowing1(roughness, normal, halfway, incoming);
}
static v3 sample_mf(float roughness, v3 normal, v2 rn) { /* ym:4354-4367 */
  float phi = 2 * PIF * rn.x;
  float theta = atanf(roughness * sqrtf(rn.y / (1 - rn.y)));
  fr3   basis = basis_fromz(normal);
  return xf_direction(&basis, V(cosf(phi) * sinf(theta), sinf(phi) * sinf(theta), cosf(theta)));
}
static float sample_mf_pdf(float roughness, v3 normal, v3 halfway) { /* ym:4370-4375 */
  float cosine = dot3(normal, halfway);
  if (cosine < 0) return 0;
  return mf_distribution(roughness, normal, halfway) * cosine;
}
static v3 bcast(float v) { return V(v, v, v); }
static v3 eval_mf_reflection(float ior, float roughness, v3 normal, v3 outgoing, v3 incoming) { /* ym:4441-4451 */
  if (dot3(normal, incoming) <= 0 || dot3(normal, outgoing) <= 0) return V(0, 0, 0);
  v3    halfway = norm3(vadd(incoming, outgoing));
  float F = fresnel_dielectric(ior, halfway, incoming), D = mf_distribution(roughness, normal, halfway);
  float G = mf_shadowing(roughness, normal, halfway, outgoing, incoming);
  return bcast(1.0f * F * D * G / (4 * dot3(normal, outgoing) * dot3(normal, incoming)) * dot3(normal, incoming));
}
static v3 eval_mf_reflection_metal(v3 eta, v3 etak, float roughness, v3 normal, v3 outgoing, v3 incoming) { /* ym:4454-4464 */
  if (dot3(normal, incoming) <= 0 || dot3(normal, outgoing) <= 0) return V(0, 0, 0);
  v3    halfway = norm3(vadd(incoming, outgoing));
  v3    F = fresnel_conductor(eta, etak, halfway, incoming);
  float D = mf_distribution(roughness, normal, halfway), G = mf_shadowing(roughness, normal, halfway, outgoing, incoming);
  return vs(vds(vs(vs(F, D), G), 4 * dot3(normal, outgoing) * dot3(normal, incoming)), dot3(normal, incoming));
}
static v3 eval_mf_transmission(float roughness, v3 normal, v3 outgoing, v3 incoming) { /* ym:4467-4480 */
  if (dot3(normal, incoming) >= 0 || dot3(normal, outgoing) <= 0) return V(0, 0, 0);
  v3    reflected = reflect3(vneg(incoming), normal);
  v3    halfway = norm3(vadd(reflected, outgoing));
  float D = mf_distribution(roughness, normal, halfway), G = mf_shadowing(roughness, normal, halfway, outgoing, reflected);
  return bcast(1.0f * D * G / (4 * dot3(normal, outgoing) * dot3(normal, reflected)) * (dot3(normal, reflected)));
}
static v3 eval_mf_refraction(float ior, float roughness, v3 normal, v3 outgoing, v3 incoming) { /* ym:4483-4514 */
  int   entering = dot3(normal, outgoing) >= 0;
  v3    up_normal = entering ? normal : vneg(normal);
  float rel_ior = entering ? ior : (1 / ior);
  if (dot3(normal, incoming) * dot3(normal, outgoing) >= 0) {
    v3    halfway = norm3(vadd(incoming, outgoing));
    float F = fresnel_dielectric(rel_ior, halfway, outgoing), D = mf_distribution(roughness, up_normal, halfway);
    float G = mf_shadowing(roughness, up_normal, halfway, outgoing, incoming);
    return bcast(1.0f * F * D * G / absr(4 * dot3(normal, outgoing) * dot3(normal, incoming)) * absr(dot3(normal, incoming)));
  } else {
    v3    halfway = vs(vneg(norm3(vadd(vs(incoming, rel_ior), outgoing))), entering ? 1.0f : -1.0f);
    float F = fresnel_dielectric(rel_ior, halfway, outgoing), D = mf_distribution(roughness, up_normal, halfway);
    float G = mf_shadowing(roughness, up_normal, halfway, outgoing, incoming);
    float t = rel_ior * dot3(halfway, incoming) + dot3(halfway, outgoing); /* pow(t, 2): powf folded to t * t at -O1+ */
    return bcast(1.0f * absr((dot3(outgoing, halfway) * dot3(incoming, halfway)) / (dot3(outgoing, normal) * dot3(incoming, normal))) *
                 (1 - F) * D * G / (t * t) * absr(dot3(normal, incoming)));
  }
}
static v3 sample_mf_reflection(float roughness, v3 normal, v3 outgoing, v2 rn) { /* ym:4530-4546 */
  if (dot3(normal, outgoing) <= 0) return V(0, 0, 0);
  return reflect3(outgoing, sample_mf(roughness, normal, rn));
}
static v3 sample_mf_transmission(float roughness, v3 normal, v3 outgoing, v2 rn) { /* ym:4549-4559 */
  if (dot3(normal, outgoing) <= 0) return V(0, 0, 0);
  v3 reflected = reflect3(outgoing, sample_mf(roughness, normal, rn));
  return vneg(reflect3(reflected, normal));
}
static v3 sample_mf_refraction(float ior, float roughness, v3 normal, v3 outgoing, float rnl, v2 rn) { /* ym:4562-4573 */
  int entering = dot3(normal, outgoing) >= 0;
  v3  up_normal = entering ? normal : vneg(normal);
  v3  halfway = sample_mf(roughness, up_normal, rn);
  if (rnl < fresnel_dielectric(entering ? ior : (1 / ior), halfway, outgoing)) return reflect3(outgoing, halfway);
  return refract3(outgoing, halfway, entering ? (1 / ior) : ior);
}
static float sample_mf_reflection_pdf(float roughness, v3 normal, v3 outgoing, v3 incoming) { /* ym:4590-4608 */
  if (dot3(normal, incoming) <= 0 || dot3(normal, outgoing) <= 0) return 0;
  v3 halfway = norm3(vadd(outgoing, incoming));
  return sample_mf_pdf(roughness, normal, halfway) / (4 * absr(dot3(outgoing, halfway)));
}
static float sample_mf_transmission_pdf(float roughness, v3 normal, v3 outgoing, v3 incoming) { /* ym:4611-4620 */
  if (dot3(normal, incoming) >= 0 || dot3(normal, outgoing) <= 0) return 0;
  v3 reflected = reflect3(vneg(incoming), normal);
  v3 halfway = norm3(vadd(reflected, outgoing));
  return sample_mf_pdf(roughness, normal, halfway) / (4 * absr(dot3(outgoing, halfway)));
}
static float sample_mf_refraction_pdf(float ior, float roughness, v3 normal, v3 outgoing, v3 incoming) { /* ym:4623-4641 */
  int   entering = dot3(normal, outgoing) >= 0;
  v3    up_normal = entering ? normal : vneg(normal);
  float rel_ior = entering ? ior : (1 / ior);
  if (dot3(normal, incoming) * dot3(normal, outgoing) >= 0) {
    v3 halfway = norm3(vadd(incoming, outgoing));
    return fresnel_dielectric(rel_ior, halfway, outgoing) * sample_mf_pdf(roughness, up_normal, halfway) / (4 * absr(dot3(outgoing, halfway)));
  } else {
    v3    halfway = vs(vneg(norm3(vadd(vs(incoming, rel_ior), outgoing))), entering ? 1.0f : -1.0f);
    float t = rel_ior * dot3(halfway, incoming) + dot3(halfway, outgoing);
    return (1 - fresnel_dielectric(rel_ior, halfway, outgoing)) * sample_mf_pdf(roughness, up_normal, halfway) * absr(dot3(halfway, outgoing)) / (t * t);
  }
}
static v3 eval_delta_refraction(float ior, v3 normal, v3 outgoing, v3 incoming) { /* ym:4664-4678 */
  if (absr(ior - 1) < 1e-3f) return dot3(normal, incoming) * dot3(normal, outgoing) <= 0 ? V(1, 1, 1) : V(0, 0, 0);
  int   entering = dot3(normal, outgoing) >= 0;
  v3    up_normal = entering ? normal : vneg(normal);
  float rel_ior = entering ? ior : (1 / ior);
  if (dot3(normal, incoming) * dot3(normal, outgoing) >= 0) return bcast(1.0f * fresnel_dielectric(rel_ior, up_normal, outgoing));
  return bcast(1.0f * (1 / (rel_ior * rel_ior)) * (1 - fresnel_dielectric(rel_ior, up_normal, outgoing)));
}
static v3 sample_delta_refraction(float ior, v3 normal, v3 outgoing, float rnl) { /* ym:4706-4717 */
  if (absr(ior - 1) < 1e-3f) return vneg(outgoing);
  int   entering = dot3(normal, outgoing) >= 0;
  v3    up_normal = entering ? normal : vneg(normal);
  float rel_ior = entering ? ior : (1 / ior);
  if (rnl < fresnel_dielectric(rel_ior, up_normal, outgoing)) return reflect3(outgoing, up_normal);
  return refract3(outgoing, up_normal, 1 / rel_ior);
}
static float sample_delta_refraction_pdf(float ior, v3 normal, v3 outgoing, v3 incoming) { /* ym:4741-4754 */
  if (absr(ior - 1) < 1e-3f) return dot3(normal, incoming) * dot3(normal, outgoing) < 0 ? 1.0f : 0.0f;
  int   entering = dot3(normal, outgoing) >= 0;
  v3    up_normal = entering ? normal : vneg(normal);
  float rel_ior = entering ? ior : (1 / ior);
  if (dot3(normal, incoming) * dot3(normal, outgoing) >= 0) return fresnel_dielectric(rel_ior, up_normal, outgoing);
  return (1 - fresnel_dielectric(rel_ior, up_normal, outgoing));
}
/* eval_brdf, pt:405-471, from the (already textured) material factors; thin = material.thin || !material.transmission */
static sbrdf eval_sbrdf(v3 base, float specular, float metallic, float roughness, float ior, float transmission, int thin,
    v3 normal, v3 outgoing) {
  sbrdf b;
  v3    weight = V(1, 1, 1);
  b.metal = vs(weight, metallic);
  weight  = vs(weight, 1 - metallic);
  b.refraction = thin ? V(0, 0, 0) : vs(weight, transmission);
  weight       = vs(weight, 1 - (thin ? 0.0f : transmission));
  b.specular = vs(weight, specular);
  weight     = vs(weight, 1 - specular * fresnel_dielectric(ior, outgoing, normal));
  b.transmission = thin ? vmul(vs(weight, transmission), base) : V(0, 0, 0);
  weight         = vs(weight, 1 - (thin ? transmission : 0.0f));
  b.diffuse = vmul(weight, base);
  b.meta = reflectivity_to_eta(base), b.metak = V(0, 0, 0);
  b.roughness = roughness * roughness, b.ior = ior;
  if (nonzero3(b.diffuse) || b.roughness) b.roughness = clampr(b.roughness, 0.03f * 0.03f, 1.0f);
  if (!nonzero3(b.specular) && !nonzero3(b.metal) && !nonzero3(b.transmission) && !nonzero3(b.refraction)) b.roughness = 1;
  b.diffuse_pdf = max3(b.diffuse);
  b.specular_pdf = max3(vs(b.specular, fresnel_dielectric(b.ior, normal, outgoing)));
  b.metal_pdf = max3(vmul(b.metal, fresnel_conductor(b.meta, b.metak, normal, outgoing)));
  b.transmission_pdf = max3(b.transmission);
  b.refraction_pdf = max3(b.refraction);
  float pdf_sum = b.diffuse_pdf + b.specular_pdf + b.metal_pdf + b.transmission_pdf + b.refraction_pdf;
  if (pdf_sum) {
    b.diffuse_pdf /= pdf_sum, b.specular_pdf /= pdf_sum, b.metal_pdf /= pdf_sum;
    b.transmission_pdf /= pdf_sum, b.refraction_pdf /= pdf_sum;
  }
  return b;
}
static v3 s_eval_brdfcos(const sbrdf* b, v3 normal, v3 outgoing, v3 incoming) { /* pt:1069-1103 */
  if (!b->roughness) return V(0, 0, 0);
  v3 r = V(0, 0, 0);
  if (nonzero3(b->diffuse)) {
    float ni = dot3(normal, incoming), no = dot3(normal, outgoing);
    r = vadd(r, vmul(b->diffuse, (ni <= 0 || no <= 0) ? V(0, 0, 0) : vs(vds(V(1, 1, 1), PIF), ni)));
  }
  if (nonzero3(b->specular)) r = vadd(r, vmul(b->specular, eval_mf_reflection(b->ior, b->roughness, normal, outgoing, incoming)));
  if (nonzero3(b->metal)) r = vadd(r, vmul(b->metal, eval_mf_reflection_metal(b->meta, b->metak, b->roughness, normal, outgoing, incoming)));
  if (nonzero3(b->transmission)) r = vadd(r, vmul(b->transmission, eval_mf_transmission(b->roughness, normal, outgoing, incoming)));
  if (nonzero3(b->refraction)) r = vadd(r, vmul(b->refraction, eval_mf_refraction(b->ior, b->roughness, normal, outgoing, incoming)));
  return r;
}
static v3 s_eval_delta(const sbrdf* b, v3 normal, v3 outgoing, v3 incoming) { /* pt:1105-1129 */
  if (b->roughness) return V(0, 0, 0);
  v3    r = V(0, 0, 0);
  float ni = dot3(normal, incoming), no = dot3(normal, outgoing);
  if (nonzero3(b->specular) && !nonzero3(b->refraction))
    r = vadd(r, vmul(b->specular, bcast((ni <= 0 || no <= 0) ? 0.0f : 1.0f * fresnel_dielectric(b->ior, normal, outgoing))));
  if (nonzero3(b->metal)) r = vadd(r, vmul(b->metal, (ni <= 0 || no <= 0) ? V(0, 0, 0) : fresnel_conductor(b->meta, b->metak, normal, outgoing)));
  if (nonzero3(b->transmission)) r = vadd(r, vmul(b->transmission, (ni >= 0 || no <= 0) ? V(0, 0, 0) : V(1, 1, 1)));
  if (nonzero3(b->refraction)) r = vadd(r, vmul(b->refraction, eval_delta_refraction(b->ior, normal, outgoing, incoming)));
  return r;
}
static v3 s_sample_brdfcos(const sbrdf* b, v3 normal, v3 outgoing, float rnl, v2 rn) { /* pt:1132-1176 */
  if (!b->roughness) return V(0, 0, 0);
  float cdf = 0.0f;
  if (b->diffuse_pdf) {
    cdf += b->diffuse_pdf;
    if (rnl < cdf) { /* sample_diffuse_reflection -> sample_hemisphere_cos, ym:4515, :4873-4879 */
      if (dot3(normal, outgoing) <= 0) return V(0, 0, 0);
      float z = sqrtf(rn.y), r = sqrtf(1 - z * z), phi = 2 * PIF * rn.x;
      fr3   basis = basis_fromz(normal);
      return xf_direction(&basis, V(r * cosf(phi), r * sinf(phi), z));
    }
  }
  if (b->specular_pdf && !b->refraction_pdf) {
    cdf += b->specular_pdf;
    if (rnl < cdf) return sample_mf_reflection(b->roughness, normal, outgoing, rn);
  }
  if (b->metal_pdf) {
    cdf += b->metal_pdf;
    if (rnl < cdf) return sample_mf_reflection(b->roughness, normal, outgoing, rn);
  }
  if (b->transmission_pdf) {
    cdf += b->transmission_pdf;
    if (rnl < cdf) return sample_mf_transmission(b->roughness, normal, outgoing, rn);
  }
  if (b->refraction_pdf) {
    cdf += b->refraction_pdf;
    if (rnl < cdf) return sample_mf_refraction(b->ior, b->roughness, normal, outgoing, rnl, rn);
  }
  return V(0, 0, 0);
}
static v3 s_sample_delta(const sbrdf* b, v3 normal, v3 outgoing, float rnl) { /* pt:1178-1215 */
  if (b->roughness) return V(0, 0, 0);
  float cdf = 0.0f;
  cdf += b->diffuse_pdf;
  if (b->specular_pdf && !b->refraction_pdf) {
    cdf += b->specular_pdf;
    if (rnl < cdf) return (dot3(normal, outgoing) <= 0) ? V(0, 0, 0) : reflect3(outgoing, normal);
  }
  if (b->metal_pdf) {
    cdf += b->metal_pdf;
    if (rnl < cdf) return (dot3(normal, outgoing) <= 0) ? V(0, 0, 0) : reflect3(outgoing, normal);
  }
  if (b->transmission_pdf) {
    cdf += b->transmission_pdf;
    if (rnl < cdf) return (dot3(normal, outgoing) <= 0) ? V(0, 0, 0) : vneg(outgoing);
  }
  if (b->refraction_pdf) {
    cdf += b->refraction_pdf;
    if (rnl < cdf) return sample_delta_refraction(b->ior, normal, outgoing, rnl);
  }
  return V(0, 0, 0);
}
static float s_sample_brdfcos_pdf(const sbrdf* b, v3 normal, v3 outgoing, v3 incoming) { /* pt:1218-1255 */
  if (!b->roughness) return 0;
  float pdf = 0.0f;
  if (b->diffuse_pdf) {
    float ni = dot3(normal, incoming), no = dot3(normal, outgoing);
    pdf += b->diffuse_pdf * ((ni <= 0 || no <= 0) ? 0 : ni / PIF);
  }
  if (b->specular_pdf && !b->refraction_pdf) pdf += b->specular_pdf * sample_mf_reflection_pdf(b->roughness, normal, outgoing, incoming);
  if (b->metal_pdf) pdf += b->metal_pdf * sample_mf_reflection_pdf(b->roughness, normal, outgoing, incoming);
  if (b->transmission_pdf) pdf += b->transmission_pdf * sample_mf_transmission_pdf(b->roughness, normal, outgoing, incoming);
  if (b->refraction_pdf) pdf += b->refraction_pdf * sample_mf_refraction_pdf(b->ior, b->roughness, normal, outgoing, incoming);
  return pdf;
}
static float s_sample_delta_pdf(const sbrdf* b, v3 normal, v3 outgoing, v3 incoming) { /* pt:1257-1280 */
  if (b->roughness) return 0;
  float pdf = 0.0f;
  float ni = dot3(normal, incoming), no = dot3(normal, outgoing);
  if (b->specular_pdf && !b->refraction_pdf) pdf += b->specular_pdf * ((ni <= 0 || no <= 0) ? 0.0f : 1.0f);
  if (b->metal_pdf) pdf += b->metal_pdf * ((ni <= 0 || no <= 0) ? 0.0f : 1.0f);
  if (b->transmission_pdf) pdf += b->transmission_pdf * ((ni >= 0 || no <= 0) ? 0.0f : 1.0f);
  if (b->refraction_pdf) pdf += b->refraction_pdf * sample_delta_refraction_pdf(b->ior, normal, outgoing, incoming);
  return pdf;
}
/* volumes: ym:4757-4812, pt:1360-1378 */
static float sample_transmittance(v3 density, float max_distance, float rl, float rd) {
  int   channel = clampi((int)(rl * 3), 0, 2);
  float dc = comp(density, channel);
  float distance = (dc == 0) ? FLT_MAXF : -logf(1 - rd) / dc;
  return fminr(distance, max_distance);
}
static v3 eval_transmittance(v3 density, float distance) {
  v3 a = vs(vneg(density), distance);
  return V(expf(a.x), expf(a.y), expf(a.z));
}
static float sample_transmittance_pdf(v3 density, float distance, float max_distance) {
  if (distance < max_distance) {
    v3 e = vmul(density, eval_transmittance(density, distance));
    return (e.x + e.y + e.z) / 3;
  }
  v3 e = eval_transmittance(density, max_distance);
  return (e.x + e.y + e.z) / 3;
}
static float eval_phasefunction(float anisotropy, v3 outgoing, v3 incoming) {
  float cosine = -dot3(outgoing, incoming);
  float denom = 1 + anisotropy * anisotropy - 2 * anisotropy * cosine;
  return (1 - anisotropy * anisotropy) / (4 * PIF * denom * sqrtf(denom));
}
static v3 sample_phasefunction(float anisotropy, v3 outgoing, v2 rn) {
  float cos_theta = 0.0f;
  if (absr(anisotropy) < 1e-3f) {
    cos_theta = 1 - 2 * rn.y;
  } else {
    float square = (1 - anisotropy * anisotropy) / (1 + anisotropy - 2 * anisotropy * rn.y);
    cos_theta = (1 + anisotropy * anisotropy - square * square) / (2 * anisotropy);
  }
  float sin_theta = sqrtf(fmaxr(0.0f, 1 - cos_theta * cos_theta));
  float phi = 2 * PIF * rn.x;
  fr3   basis = basis_fromz(vneg(outgoing));
  return xf_vector(&basis, V(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta));
}

/* diagnostics (yho_debug_pixel): every path ray of the samples traced by this thread and what it hit */
typedef struct { yh_ray* rays; yh_hit* hits; int max, count; } path_log_t;
static __thread path_log_t* g_path_log = NULL;
static void log_path_ray(const ray_t* ray, int hit, int object, int element, v2 uv, float dist) {
  path_log_t* l = g_path_log;
  if (!l || l->count >= l->max) return;
  yh_ray r = {{ray->o.x, ray->o.y, ray->o.z}, {ray->d.x, ray->d.y, ray->d.z}, ray->tmin, ray->tmax};
  yh_hit h = {object, element, {uv.x, uv.y}, dist, hit};
  l->rays[l->count] = r, l->hits[l->count] = h, l->count++;
}

static void trace_path(const scene_t* sc, ray_t ray, rng_t* rng, const yh_trace_params* params, float out[4]) {
  v3  radiance = V(0, 0, 0), weight = V(1, 1, 1);
  int hit = 0;
  const int shader = params->shader;
  int max_bounces = params->bounces;
  if (shader == YH_SHADER_EYELIGHT) max_bounces = params->bounces > 4 ? params->bounces : 4; /* max(bounces, 4) */
  if (shader == YH_SHADER_NORMAL) max_bounces = 1;
  vsdf_t vol; /* the volume stack never holds more than one entry, pt:1460-1467 */
  int    in_vol = 0;
  memset(&vol, 0, sizeof(vol));
  for (int bounce = 0; bounce < max_bounces; bounce++) {
    int   object = -1, element = -1;
    v2    uv = {0, 0};
    float dist = 0;
    int found = walk_scene(sc, &ray, &object, &element, &uv, &dist, 0);
    if (g_path_log) log_path_ray(&ray, found, object, element, uv, dist);
    if (!found) {
      radiance = vadd(radiance, vmul(weight, eval_environment(sc, ray.d)));
      if (shader == YH_SHADER_NORMAL) hit = 1; /* trace_normal returns alpha 1 on a miss too */
      break;
    }
    v3 outgoing = vneg(ray.d);
    /* transmittance through the current volume (trace_path only), pt:1403-1414 */
    int scatter_event = 0;
    if (shader == YH_SHADER_PATH && in_vol) {
      float rd = rng_1f(rng), rl = rng_1f(rng); /* arguments evaluated right to left */
      float distance = sample_transmittance(vol.density, dist, rl, rd);
      weight = vmul(weight, vds(eval_transmittance(vol.density, distance), sample_transmittance_pdf(vol.density, distance, dist)));
      scatter_event = distance < dist;
      dist = distance;
    }
    if (scatter_event) { /* pt:1470-1497 */
      v3 position = vadd(ray.o, vs(ray.d, dist));
      hit = 1;
      v3 incoming;
      if (rng_1f(rng) < 0.5f) {
        v2    rn = rng_2f(rng);
        float rnl = rng_1f(rng);
        (void)rnl;
        incoming = nonzero3(vol.density) ? sample_phasefunction(vol.anisotropy, outgoing, rn) : V(0, 0, 0);
      } else {
        v2    ruv = rng_2f(rng);
        float rel = rng_1f(rng), rl = rng_1f(rng);
        incoming = sample_lights(sc, position, rl, rel, ruv);
      }
      v3    f = nonzero3(vol.density) ? vs(vmul(vol.scatter, vol.density), eval_phasefunction(vol.anisotropy, outgoing, incoming)) : V(0, 0, 0);
      float spdf = nonzero3(vol.density) ? eval_phasefunction(vol.anisotropy, outgoing, incoming) : 0;
      weight = vmul(weight, vds(f, 0.5f * spdf + 0.5f * sample_lights_pdf(sc, position, incoming)));
      ray_t nr = {position, incoming, RAY_EPS, FLT_MAXF};
      ray = nr;
    } else {
    const obj_t*   ob = &sc->objs[object];
    const shape_t* sh = &sc->shapes[ob->shape];
    const mat_t*   mat = &sc->mats[ob->material];
    int   is_hair = sh->nlines != 0;
    v3    position = eval_position(sc, object, element, uv);
    v3    normal; /* eval_shading_normal, pt:350-369 */
    if (sh->npoints) {
      normal = vneg(outgoing); /* pt:364-365 */
    } else if (!is_hair) {
      normal = mat->normal_tex >= 0 ? eval_normalmap(sc, object, element, uv, mat->normal_tex) : eval_normal(sc, object, element, uv);
      if (mat->thin) normal = dot3(normal, outgoing) >= 0 ? normal : vneg(normal);
    } else {
      normal = orthonorm(outgoing, eval_normal(sc, object, element, uv));
    }
    if (shader == YH_SHADER_NORMAL) { /* pt:1645-1650 */
      radiance = vadds(vs(normal, 0.5f), 0.5f);
      hit = 1;
      break;
    }
    /* eval_emission pt:397-402 and the factors of eval_brdf pt:409-424 (a missing texture is {1,1,1}; the
     * transmission factor comes from the slot the caller filled with the reference's choice, its emission texture) */
    v2    texcoord = eval_texcoord(sc, object, element, uv);
    v3    emission = vmul(mat->emission, eval_texture(sc, mat->emission_tex, texcoord));
    v3    base = vmul(mat->color, eval_texture(sc, mat->color_tex, texcoord));
    float specular = mat->specular * eval_texture(sc, mat->specular_tex, texcoord).x;
    float metallic = mat->metallic * eval_texture(sc, mat->metallic_tex, texcoord).x;
    float roughness = mat->roughness * eval_texture(sc, mat->roughness_tex, texcoord).x;
    float transmission = mat->transmission * eval_texture(sc, mat->transmission_tex, texcoord).x;
    v3    otex = eval_texture(sc, mat->opacity_tex, texcoord);
    float opacity = mat->opacity * ((otex.x + otex.y + otex.z) / 3);
    int   thin = mat->thin || !mat->transmission;
    if (opacity > 0.999f) opacity = 1;
    sbrdf sb;
    hbrdf hb;
    memset(&sb, 0, sizeof(sb));
    if (is_hair) hb = hair_brdf_make(&mat->hair, uv.y, normal, eval_normal(sc, object, element, uv));
    else sb = eval_sbrdf(base, specular, metallic, roughness, mat->ior, transmission, thin, normal, outgoing);
    int delta = !is_hair && !sb.roughness; /* is_delta, pt:495 */
    if (opacity < 1 && rng_1f(rng) >= opacity) { /* pt:1429-1433 */
      ray_t nr = {vadd(position, vs(ray.d, 1e-2f)), ray.d, RAY_EPS, FLT_MAXF};
      ray = nr;
      bounce -= 1;
      continue;
    }
    hit      = 1;
    radiance = vadd(radiance, vmul(weight, emission));
    v3 incoming;
    if (shader == YH_SHADER_EYELIGHT) { /* pt:1617-1626: light at the eye; delta surfaces are followed */
      v3 f = is_hair ? hair_eval(&hb, outgoing, outgoing) : s_eval_brdfcos(&sb, normal, outgoing, outgoing);
      radiance = vadd(radiance, vmul(vs(weight, PIF), f));
      if (!delta) break;
      incoming = s_sample_delta(&sb, normal, outgoing, rng_1f(rng));
      weight = vmul(weight, vds(s_eval_delta(&sb, normal, outgoing, incoming), s_sample_delta_pdf(&sb, normal, outgoing, incoming)));
      if ((weight.x == 0 && weight.y == 0 && weight.z == 0) || !finite3(weight)) break;
      ray_t nr = {position, incoming, RAY_EPS, FLT_MAXF};
      ray = nr;
      continue;
    }
    if (delta) { /* pt:1452-1456 */
      incoming = s_sample_delta(&sb, normal, outgoing, rng_1f(rng));
      weight = vmul(weight, vds(s_eval_delta(&sb, normal, outgoing, incoming), s_sample_delta_pdf(&sb, normal, outgoing, incoming)));
    } else {
      /* MIS one-sample model, pt:1441-1451 (trace_naive always samples the BSDF, pt:1548-1551).
       * GCC evaluates call arguments right to left: rand2f first. */
      if (shader == YH_SHADER_NAIVE || rng_1f(rng) < 0.5f) {
        v2    rn = rng_2f(rng);
        float rnl = rng_1f(rng);
        incoming = is_hair ? hair_sample(&hb, outgoing, rn) : s_sample_brdfcos(&sb, normal, outgoing, rnl, rn);
      } else {
        v2    ruv = rng_2f(rng);
        float rel = rng_1f(rng), rl = rng_1f(rng);
        incoming = sample_lights(sc, position, rl, rel, ruv);
      }
      v3    brdfcos;
      float brdf_pdf;
      if (is_hair) brdfcos = hair_eval(&hb, outgoing, incoming), brdf_pdf = hair_pdf(&hb, outgoing, incoming);
      else brdfcos = s_eval_brdfcos(&sb, normal, outgoing, incoming), brdf_pdf = s_sample_brdfcos_pdf(&sb, normal, outgoing, incoming);
      if (shader == YH_SHADER_NAIVE) {
        weight = vmul(weight, vds(brdfcos, brdf_pdf));
      } else {
        float light_pdf = sample_lights_pdf(sc, position, incoming);
        weight = vmul(weight, vds(brdfcos, 0.5f * brdf_pdf + 0.5f * light_pdf));
      }
    }
    /* update the volume stack, pt:1458-1467 (trace_path only) */
    if (shader == YH_SHADER_PATH && !mat->thin && mat->transmission && dot3(normal, outgoing) * dot3(normal, incoming) < 0) {
      if (!in_vol) { /* eval_vsdf, pt:505-526 */
        v3 scattering = vmul(mat->scattering, eval_texture(sc, mat->scattering_tex, texcoord));
        vol.density = V(0, 0, 0);
        if (transmission && !thin) {
          v3 l = V(logf(clampr(base.x, 0.0001f, 1.0f)), logf(clampr(base.y, 0.0001f, 1.0f)), logf(clampr(base.z, 0.0001f, 1.0f)));
          vol.density = vds(vneg(l), mat->trdepth);
        }
        vol.scatter = scattering, vol.anisotropy = mat->scanisotropy;
        in_vol = 1;
      } else {
        in_vol = 0;
      }
    }
    ray_t nr = {position, incoming, RAY_EPS, FLT_MAXF};
    ray = nr;
    }
    if ((weight.x == 0 && weight.y == 0 && weight.z == 0) || !finite3(weight)) break;
    if (bounce > 3) { /* russian roulette, pt:1503-1507 */
      float rr = fminr((float)0.99, max3(weight));
      if (rng_1f(rng) >= rr) break;
      weight = vs(weight, 1 / rr);
    }
  }
  out[0] = radiance.x, out[1] = radiance.y, out[2] = radiance.z, out[3] = hit ? 1.0f : 0.0f;
}

static void trace_pixel(state_t* st, int i, int j) { /* trace_sample, pt:1676-1689 */
  pixel_t* px = &st->px[(size_t)j * st->w + i];
  if (st->counter) {
    uint64_t s = (uint64_t)st->srank + (uint64_t)st->sworld * (uint64_t)px->samples;
    px->rng = rng_make(st->params.seed, px->seq + (s << 31));
  }
  v2 luv = rng_2f(&px->rng); /* right-to-left argument evaluation: lens sample first */
  v2 puv = rng_2f(&px->rng);
  ray_t ray = camera_ray(st, i, j, puv, luv);
  float s[4];
  trace_path(st->sc, ray, &px->rng, &st->params, s);
  v3 c = V(s[0], s[1], s[2]);
  if (!finite3(c)) c = V(0, 0, 0);
  if (max3(c) > st->params.clamp) c = vs(c, st->params.clamp / max3(c));
  px->accum[0] += c.x, px->accum[1] += c.y, px->accum[2] += c.z, px->accum[3] += s[3];
  px->samples += 1;
  float* r = st->render + 4 * ((size_t)j * st->w + i);
  for (int k = 0; k < 4; k++) r[k] = px->accum[k] / px->samples;
}
static void* row_worker(void* arg) { /* parallel_for, pt:1954-1970: dynamic rows */
  state_t* st = (state_t*)arg;
  for (;;) {
    int j = __sync_fetch_and_add(&st->next_row, 1);
    if (j >= st->h) break;
    for (int i = 0; i < st->w; i++) trace_pixel(st, i, j);
  }
  return NULL;
}

/* Diagnostics: traces `nsamples` MORE samples of pixel (i, j) of this state (the pixel's accumulation and RNG advance as
 * in a normal pass) and records the path rays with their hit records; returns the number of records written. */
int yho_debug_pixel(void* h, int i, int j, int nsamples, yh_ray* rays, yh_hit* hits, int max_records) {
  state_t*   st = (state_t*)h;
  path_log_t log = {rays, hits, max_records, 0};
  g_path_log = &log;
  for (int s = 0; s < nsamples; s++) trace_pixel(st, i, j);
  g_path_log = NULL;
  return log.count;
}

void* yho_state_create(void* scene, const yh_camera_desc* c, const yh_trace_params* p) { /* init_state, pt:1931-1946 */
  state_t* st = (state_t*)calloc(1, sizeof(state_t));
  st->sc = (const scene_t*)scene;
  st->cam_frame = frame_from(c->frame);
  st->lens = c->lens, st->film[0] = c->film[0], st->film[1] = c->film[1], st->focus = c->focus, st->aperture = c->aperture;
  st->params = *p;
  if (c->film[0] > c->film[1]) st->w = p->resolution, st->h = (int)roundf(p->resolution * c->film[1] / c->film[0]);
  else st->w = (int)roundf(p->resolution * c->film[0] / c->film[1]), st->h = p->resolution;
  size_t n = (size_t)st->w * st->h;
  st->px = (pixel_t*)calloc(n ? n : 1, sizeof(pixel_t));
  st->render = (float*)calloc(n ? n : 1, 16);
  rng_t master = rng_make(1301081, 1);
  for (size_t i = 0; i < n; i++) {
    st->px[i].seq = (uint64_t)(rng_1i(&master, 1 << 31) / 2 + 1);
    st->px[i].rng = rng_make(p->seed, st->px[i].seq);
  }
  st->counter = 0, st->srank = 0, st->sworld = 1;
  return st;
}
void yho_state_destroy(void* h) {
  state_t* st = (state_t*)h;
  if (!st) return;
  free(st->px), free(st->render), free(st);
}
void yho_state_set_sample_shard(void* h, int rank, int world) { /* counter mode, see state_t */
  state_t* st = (state_t*)h;
  st->counter = 1, st->srank = rank, st->sworld = world;
  size_t n = (size_t)st->w * st->h;
  for (size_t i = 0; i < n; i++) {
    memset(st->px[i].accum, 0, 16);
    st->px[i].samples = 0;
    st->px[i].rng = rng_make(st->params.seed, st->px[i].seq);
  }
  memset(st->render, 0, 16 * n);
}
void yho_state_size(void* h, int* w, int* hgt) { *w = ((state_t*)h)->w, *hgt = ((state_t*)h)->h; }
double yho_trace_samples(void* h, int nsamples) { /* trace_samples, pt:1992-2007, nsamples passes */
  state_t* st = (state_t*)h;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int nthreads = st->params.noparallel ? 1 : yho_hardware_concurrency();
  if (nthreads > 256) nthreads = 256;
  for (int s = 0; s < nsamples; s++) {
    st->next_row = 0;
    if (nthreads == 1) {
      row_worker(st);
    } else {
      pthread_t th[256];
      for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, row_worker, st);
      for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
void yho_state_download(void* h, float* render, float* accum, int* samples, uint64_t* rng) {
  state_t* st = (state_t*)h;
  size_t   n = (size_t)st->w * st->h;
  if (render) memcpy(render, st->render, 16 * n);
  for (size_t i = 0; i < n; i++) {
    if (accum) memcpy(accum + 4 * i, st->px[i].accum, 16);
    if (samples) samples[i] = st->px[i].samples;
    if (rng) rng[2 * i] = st->px[i].rng.state, rng[2 * i + 1] = st->px[i].rng.inc;
  }
}

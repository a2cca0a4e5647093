// pbrt hair BSDF (Chiang et al. 2016), B200 restatement of the reference's
// libs/yocto_extension/yocto_extension.cpp.  fp32, same operation order as the
// reference (TU compiled with --fmad=false), transcendental functions from
// yh_libm.cuh so values agree with the x86-64 oracle bit for bit.
#pragma once
#include "yh_libm.cuh"

namespace yh {

constexpr int p_max = 3;  // yocto_extension.h:84

// yocto_extension.h:86-95
struct hair_material {
  vec3f sigma_a;
  float beta_m, beta_n, alpha, eta;
  vec3f color;
  float eumelanin, pheomelanin;
};
// yocto_extension.h:97-113 (30 floats, same field order as yh_hair_brdf)
struct hair_brdf {
  vec3f   sigma_a;
  float   alpha, eta, h;
  float   v[p_max + 1];
  float   s;
  float   sin_2k_alpha[3];
  float   cos_2k_alpha[3];
  float   gamma_o;
  frame3f world_to_brdf;
};

YH_HD float safe_asin(float x) { return (float)asin_d((double)clampf(x, -1.0f, 1.0f)); }  // :111
YH_HD float safe_sqrt(float x) { return sqrt_(fmax_(0.0f, x)); }                         // :113
YH_HD vec3f exp3(const vec3f& a) { return {expf_(a.x), expf_(a.y), expf_(a.z)}; }
YH_HD vec3f log3(const vec3f& a) { return {logf_(a.x), logf_(a.y), logf_(a.z)}; }

// pow<n>(v), yocto_extension.cpp:95-109 (the factors of 1 are exact)
YH_HD float pow2_(float v) { return v * v; }
YH_HD float pow3_(float v) { return v * v * v; }
YH_HD float pow4_(float v) { float n2 = pow2_(v); return n2 * n2; }
YH_HD float pow5_(float v) { float n2 = pow2_(v); return n2 * n2 * v; }
YH_HD float pow10_(float v) { float n2 = pow5_(v); return n2 * n2; }
YH_HD float pow11_(float v) { float n2 = pow5_(v); return n2 * n2 * v; }
YH_HD float pow20_(float v) { float n2 = pow10_(v); return n2 * n2; }
YH_HD float pow22_(float v) { float n2 = pow11_(v); return n2 * n2; }

// :115-119
YH_HD vec3f sigma_a_from_concentration(float ce, float cp) {
  return ce * vec3f{0.419f, 0.697f, 1.37f} + cp * vec3f{0.187f, 0.4f, 1.05f};
}
// :121-125
YH_HD vec3f sigma_a_from_reflectance(const vec3f& c, float beta_n) {
  float den = 5.969f - 0.215f * beta_n + 2.532f * sqr(beta_n) - 10.73f * pow3_(beta_n) +
              5.574f * pow4_(beta_n) + 0.245f * pow5_(beta_n);
  vec3f q = log3(c) / den;
  return q * q;
}

// eval_hair_brdf, yocto_extension.cpp:127-177, in two parts: what depends on the MATERIAL only (absorption, the
// longitudinal variances v[], the azimuthal scale s, the scale-tilt terms: :131-146, :151-167) and what depends on the
// hit (h, gamma_o, the frame: :148-149, :169-175).  The wavefront evaluates the first part once per material
// (dev_material::hair_proto, k_material_hair_consts) instead of once per vertex - same operations, same bits.
YH_HD void hair_brdf_material_part(const hair_material& material, hair_brdf& brdf) {
  brdf.sigma_a = {0, 0, 0};
  if (nonzero(material.sigma_a)) {
    brdf.sigma_a = material.sigma_a;
  } else if (nonzero(material.color)) {
    brdf.sigma_a = sigma_a_from_reflectance(material.color, material.beta_n);
  } else if (material.eumelanin != 0 || material.pheomelanin != 0) {
    brdf.sigma_a = sigma_a_from_concentration(material.eumelanin, material.pheomelanin);
  }
  float beta_m = material.beta_m;
  float beta_n = material.beta_n;
  brdf.alpha   = material.alpha;
  brdf.eta     = material.eta;
  brdf.v[0]    = sqr(0.726f * beta_m + 0.812f * sqr(beta_m) + 3.7f * pow20_(beta_m));
  brdf.v[1]    = 0.25f * brdf.v[0];
  brdf.v[2]    = 4 * brdf.v[0];
  brdf.v[3]    = brdf.v[2];
  brdf.s       = 0.626657069f * (0.265f * beta_n + 1.194f * sqr(beta_n) + 5.372f * pow22_(beta_n));
  brdf.sin_2k_alpha[0] = sinf_(div_(pif, 180.0f) * brdf.alpha);
  brdf.cos_2k_alpha[0] = safe_sqrt(1 - sqr(brdf.sin_2k_alpha[0]));
  for (int i = 1; i < 3; i++) {
    brdf.sin_2k_alpha[i] = 2 * brdf.cos_2k_alpha[i - 1] * brdf.sin_2k_alpha[i - 1];
    brdf.cos_2k_alpha[i] = sqr(brdf.cos_2k_alpha[i - 1]) - sqr(brdf.sin_2k_alpha[i - 1]);
  }
}
YH_HD void hair_brdf_hit_part(hair_brdf& brdf, float v, const vec3f& normal, const vec3f& tangent) {
  brdf.h             = -1 + 2 * v;
  brdf.gamma_o       = safe_asin(brdf.h);
  brdf.world_to_brdf = inverse_rigid(frame_fromzx(normal, tangent));
}
YH_HD hair_brdf eval_hair_brdf(const hair_material& material, float v, const vec3f& normal,
    const vec3f& tangent) {
  hair_brdf brdf;
  hair_brdf_material_part(material, brdf);
  hair_brdf_hit_part(brdf, v, normal, tangent);
  return brdf;
}

// 1 / (4^i (i!)^2) denominators of i0(), yocto_extension.cpp:179-192: the
// reference forms them in int64 and converts to float at the division.
YH_HD float i0(float x) {
  const float den[10] = {1.0f, 4.0f, 64.0f, 2304.0f, 147456.0f, 14745600.0f, 2123366400.0f,
      416179814400.0f, 106542032486400.0f, 34519618525593600.0f};
  float val = 0;
  float x2i = 1;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    // 1, 4 and 64 are powers of two: the quotient is the exact product (div_pow2)
    val += i == 0 ? x2i : i == 1 ? div_pow2(x2i, 0.25f) : i == 2 ? div_pow2(x2i, 0.015625f) : div_(x2i, den[i]);
    x2i *= x * x;
  }
  return val;
}
// :194-199.  log(2 pi) is the same call with the same argument every time: log_2pi carries its value.
YH_HD float log_i0(float x, float log_2pi) {
  if (x > 12)
    return x + 0.5f * (-log_2pi + logf_(div_(1.0f, x)) + div_(1.0f, (8 * x)));
  else
    return logf_(i0(x));
}
YH_HD float log_i0(float x) { return log_i0(x, logf_(2 * pif)); }

// What the lobes compute from the MATERIAL's numbers alone (the variances v[] and the logistic scale s), gathered so
// that the wavefront forms each value once per material (dev_material::hair_k, k_material_hair_consts) and not once per
// lobe per vertex.  Every field is the result of the very call the reference makes at that place, with the same
// argument, so using the stored value changes no bit:
//   mp(), :201-207      sinh(1 / v) * 2 * v   (double: float / (double*int*float)), 1 / v, log(1 / (2 v))
//   log_i0(), :196      log(2 pi)
//   trimmed_logistic(), :243-245 and sample_trimmed_logistic(), :359-363, always called with [a, b] = [-pi, pi]:
//                       logistic_cdf(-pi, s) and logistic_cdf(pi, s) - logistic_cdf(-pi, s)
//   sample_hair_scattering(), :436   exp(-2 / v[p])
struct hair_consts {
  double mp_den[p_max + 1];      // read only where v[p] > 0.1
  float  inv_v[p_max + 1];       // read only where v[p] <= 0.1, like the next
  float  log_inv_2v[p_max + 1];
  float  exp_m2_inv_v[p_max + 1];
  float  log_2pi;
  float  cdf_lo, cdf_width;
};
YH_HD double mp_den(float v) { return sinh_d((double)div_(1.0f, v)) * 2.0 * (double)v; }
YH_HD_CALL float mp(float cos_theta_i, float cos_theta_o, float sin_theta_i, float sin_theta_o, float v, double den,
    float inv_v, float log_inv_2v, float log_2pi) {
  float a = div_(cos_theta_i * cos_theta_o, v);
  float b = div_(sin_theta_i * sin_theta_o, v);
  if (v <= 0.1f) return expf_(log_i0(a, log_2pi) - b - inv_v + 0.6931f + log_inv_2v);
  return (float)((double)(expf_(-b) * i0(a)) / den);
}
YH_HD float mp(float cos_theta_i, float cos_theta_o, float sin_theta_i, float sin_theta_o, float v, int p,
    const hair_consts& k) {
  return mp(cos_theta_i, cos_theta_o, sin_theta_i, sin_theta_o, v, k.mp_den[p], k.inv_v[p], k.log_inv_2v[p], k.log_2pi);
}
YH_HD float mp(float cos_theta_i, float cos_theta_o, float sin_theta_i, float sin_theta_o, float v) {
  bool small = v <= 0.1f;
  return mp(cos_theta_i, cos_theta_o, sin_theta_i, sin_theta_o, v, small ? 0.0 : mp_den(v),
      small ? div_(1.0f, v) : 0.0f, small ? logf_(div_(1.0f, (2 * v))) : 0.0f, logf_(2 * pif));
}

// fresnel_dielectric(eta, {0,0,1}, {0,0,cos_theta}), yocto_math.h:4215-4235
YH_HD float fresnel_dielectric_cos(float eta, float cos_theta) {
  float cosw = fabs_(0.0f * 0.0f + 0.0f * 0.0f + 1.0f * cos_theta);
  float sin2  = 1 - cosw * cosw;
  float eta2  = eta * eta;
  float cos2t = 1 - div_(sin2, eta2);
  if (cos2t < 0) return 1;
  float t0 = sqrt_(cos2t);
  float t1 = eta * t0;
  float t2 = eta * cosw;
  float rs = div_((cosw - t1), (cosw + t1));
  float rp = div_((t0 - t2), (t0 + t2));
  return div_pow2((rs * rs + rp * rp), 0.5f);
}

// ap(), :209-230
// The Fresnel term depends on (cos_theta_o, eta, h) only - not on the transmittance T - so the two ap() calls of one
// shading point (eval :291 and compute_ap_pdf :391, which differ in T alone) share it: ap_fresnel() once, ap_from() twice.
YH_HD float ap_fresnel(float cos_theta_o, float eta, float h) {
  float cos_gamma_o = safe_sqrt(1 - h * h);
  float cos_theta   = cos_theta_o * cos_gamma_o;
  return fresnel_dielectric_cos(eta, cos_theta);
}
YH_HD void ap_from(float f, const vec3f& T, vec3f out[p_max + 1]) {
  out[0]            = {f, f, f};
  out[1]            = sqr(1 - f) * T;
  for (int p = 2; p < p_max; p++) out[p] = out[p - 1] * T * f;
  out[p_max] = out[p_max - 1] * f * T / (vec3f{1.f, 1.f, 1.f} - T * f);
}
YH_HD void ap(float cos_theta_o, float eta, float h, const vec3f& T, vec3f out[p_max + 1]) {
  ap_from(ap_fresnel(cos_theta_o, eta, h), T, out);
}

YH_HD float phi_p(int p, float gamma_o, float gamma_t) {  // :232-234
  return 2 * p * gamma_t - 2 * gamma_o + p * pif;
}
YH_HD float logistic(float x, float s) {  // :236-239
  x = fabs_(x);
  float e = expf_(div_(-x, s));  // the reference writes the call twice; same argument, same value
  return div_(e, (s * sqr(1 + e)));
}
YH_HD float logistic_cdf(float x, float s) { return div_(1.0f, (1 + expf_(div_(-x, s)))); }  // :241
YH_HD float trimmed_logistic(float x, float s, float a, float b) {  // :243-245
  return div_(logistic(x, s), (logistic_cdf(b, s) - logistic_cdf(a, s)));
}
YH_HD float np_dphi(float phi, int p, float gamma_o, float gamma_t) {
  float dphi = phi - phi_p(p, gamma_o, gamma_t);
  while (dphi > pif) dphi -= 2 * pif;
  while (dphi < -pif) dphi += 2 * pif;
  return dphi;
}
YH_HD_CALL float np(float phi, int p, float s, float gamma_o, float gamma_t) {  // :247-253
  return trimmed_logistic(np_dphi(phi, p, gamma_o, gamma_t), s, -pif, pif);
}
// the same with the denominator of trimmed_logistic(., s, -pi, pi) taken from the material's constants
YH_HD_CALL float np(float phi, int p, float s, float gamma_o, float gamma_t, float cdf_width) {
  return div_(logistic(np_dphi(phi, p, gamma_o, gamma_t), s), cdf_width);
}

// hair scale tilt for lobe p, :297-322 (shared by eval / sample / pdf)
YH_HD void tilt(const hair_brdf& brdf, int p, float sin_theta_o, float cos_theta_o,
    float& sin_theta_op, float& cos_theta_op) {
  if (p == 0) {
    sin_theta_op = sin_theta_o * brdf.cos_2k_alpha[1] - cos_theta_o * brdf.sin_2k_alpha[1];
    cos_theta_op = cos_theta_o * brdf.cos_2k_alpha[1] + sin_theta_o * brdf.sin_2k_alpha[1];
  } else if (p == 1) {
    sin_theta_op = sin_theta_o * brdf.cos_2k_alpha[0] + cos_theta_o * brdf.sin_2k_alpha[0];
    cos_theta_op = cos_theta_o * brdf.cos_2k_alpha[0] - sin_theta_o * brdf.sin_2k_alpha[0];
  } else if (p == 2) {
    sin_theta_op = sin_theta_o * brdf.cos_2k_alpha[2] + cos_theta_o * brdf.sin_2k_alpha[2];
    cos_theta_op = cos_theta_o * brdf.cos_2k_alpha[2] - sin_theta_o * brdf.sin_2k_alpha[2];
  } else {
    sin_theta_op = sin_theta_o;
    cos_theta_op = cos_theta_o;
  }
}

// compute_ap_pdf, :365-397: note sin_theta_o = safe_sqrt(1 - cos^2) here, which
// is |sin_theta_o| up to rounding, so Ap is evaluated again for the pdf.
YH_HD void hair_ap_pdf(const hair_brdf& brdf, float cos_theta_o, float f, float ap_pdf[p_max + 1]) {
  float sin_theta_o = safe_sqrt(1 - cos_theta_o * cos_theta_o);
  float sin_theta_t = div_(sin_theta_o, brdf.eta);
  float cos_theta_t = safe_sqrt(1 - sqr(sin_theta_t));
  float etap        = div_(sqrt_(brdf.eta * brdf.eta - sqr(sin_theta_o)), cos_theta_o);
  float sin_gamma_t = div_(brdf.h, etap);
  float cos_gamma_t = safe_sqrt(1 - sqr(sin_gamma_t));
  vec3f T           = exp3(-brdf.sigma_a * div_(2 * cos_gamma_t, cos_theta_t));
  vec3f a[p_max + 1];
  ap_from(f, T, a);
  float sum_y = 0.0f;
  for (int i = 0; i <= p_max; i++) sum_y += luminance(a[i]);
  for (int i = 0; i <= p_max; i++) ap_pdf[i] = div_(luminance(a[i]), sum_y);
}

// Terms that depend on the outgoing direction only; eval, pdf and sample of
// one shading point share them (the reference recomputes them per call with
// identical arithmetic, so sharing does not change any bit).
struct hair_wo_terms {
  float sin_theta_o, cos_theta_o, phi_o;
  float gamma_t;
  vec3f ap[p_max + 1];
  float ap_pdf[p_max + 1];  // compute_ap_pdf(brdf, cos_theta_o), :365-397
};
YH_HD hair_wo_terms hair_prepare_wo(const hair_brdf& brdf, const vec3f& outgoing_) {
  hair_wo_terms t;
  vec3f outgoing = transform_direction(brdf.world_to_brdf, outgoing_);
  t.sin_theta_o  = outgoing.x;
  t.cos_theta_o  = safe_sqrt(1 - sqr(t.sin_theta_o));
  t.phi_o        = atan2f_(outgoing.z, outgoing.y);
  // :281-291 — eval derives sin_theta_t from sin_theta_o; compute_ap_pdf (:365-397)
  // re-derives sin_theta_o from cos_theta_o: see hair_ap_for_pdf below.
  float sin_theta_t = div_(t.sin_theta_o, brdf.eta);
  float cos_theta_t = safe_sqrt(1 - sqr(sin_theta_t));
  float etap        = div_(sqrt_(brdf.eta * brdf.eta - sqr(t.sin_theta_o)), t.cos_theta_o);
  float sin_gamma_t = div_(brdf.h, etap);
  float cos_gamma_t = safe_sqrt(1 - sqr(sin_gamma_t));
  t.gamma_t         = safe_asin(sin_gamma_t);
  vec3f T           = exp3(-brdf.sigma_a * div_(2 * cos_gamma_t, cos_theta_t));
  float f = ap_fresnel(t.cos_theta_o, brdf.eta, brdf.h);
  ap_from(f, T, t.ap);
  hair_ap_pdf(brdf, t.cos_theta_o, f, t.ap_pdf);
  return t;
}
// eval_hair_scattering, :255-336
YH_HD vec3f eval_hair_scattering(const hair_brdf& brdf, const hair_wo_terms& t, const vec3f& incoming_) {
  vec3f incoming    = transform_direction(brdf.world_to_brdf, incoming_);
  float sin_theta_i = incoming.x;
  float cos_theta_i = safe_sqrt(1 - sqr(sin_theta_i));
  float phi_i       = atan2f_(incoming.z, incoming.y);
  float phi         = phi_i - t.phi_o;
  vec3f fsum        = {0, 0, 0};
  for (int p = 0; p < p_max; p++) {
    float sin_theta_op, cos_theta_op;
    tilt(brdf, p, t.sin_theta_o, t.cos_theta_o, sin_theta_op, cos_theta_op);
    cos_theta_op = fabs_(cos_theta_op);
    fsum = fsum + mp(cos_theta_i, cos_theta_op, sin_theta_i, sin_theta_op, brdf.v[p]) * t.ap[p] *
                      np(phi, p, brdf.s, brdf.gamma_o, t.gamma_t);
  }
  fsum = fsum + mp(cos_theta_i, t.cos_theta_o, sin_theta_i, t.sin_theta_o, brdf.v[p_max]) * t.ap[p_max] /
                    (2 * pif);
  return fsum;
}
YH_HD vec3f eval_hair_scattering(const hair_brdf& brdf, const vec3f& outgoing, const vec3f& incoming) {
  return eval_hair_scattering(brdf, hair_prepare_wo(brdf, outgoing), incoming);
}

// compact1by1 / demux_float, :339-357
YH_HD uint32_t compact1by1(uint32_t x) {
  x &= 0x55555555;
  x = (x ^ (x >> 1)) & 0x33333333;
  x = (x ^ (x >> 2)) & 0x0f0f0f0f;
  x = (x ^ (x >> 4)) & 0x00ff00ff;
  x = (x ^ (x >> 8)) & 0x0000ffff;
  return x;
}
YH_HD vec2f demux_float(float f) {
  uint64_t v = (uint64_t)(f * 4294967296.0f);
  uint32_t b0 = compact1by1((uint32_t)v), b1 = compact1by1((uint32_t)(v >> 1));
  return {div_pow2((float)b0, 1.0f / 65536.0f), div_pow2((float)b1, 1.0f / 65536.0f)};
}
YH_HD float sample_trimmed_logistic(float u, float s, float a, float b, float cdf_lo, float cdf_width) {
  float x = -s * logf_(div_(1.0f, (u * cdf_width + cdf_lo)) - 1);
  return clampf(x, a, b);
}
YH_HD float sample_trimmed_logistic(float u, float s, float a, float b) {  // :359-363
  return sample_trimmed_logistic(u, s, a, b, logistic_cdf(a, s), logistic_cdf(b, s) - logistic_cdf(a, s));
}
YH_HD hair_consts make_hair_consts(const hair_brdf& brdf) {
  hair_consts k;
  for (int p = 0; p <= p_max; p++) {
    float v = brdf.v[p];
    bool small = v <= 0.1f;
    k.mp_den[p]       = small ? 0.0 : mp_den(v);
    k.inv_v[p]        = small ? div_(1.0f, v) : 0.0f;
    k.log_inv_2v[p]   = small ? logf_(div_(1.0f, (2 * v))) : 0.0f;
    k.exp_m2_inv_v[p] = expf_(div_(-2.0f, v));
  }
  k.log_2pi   = logf_(2 * pif);
  k.cdf_lo    = logistic_cdf(-pif, brdf.s);
  k.cdf_width = logistic_cdf(pif, brdf.s) - k.cdf_lo;
  return k;
}

// sample_hair_scattering, :399-479
YH_HD vec3f sample_hair_scattering(const hair_brdf& brdf, const hair_wo_terms& t, const vec2f& rn,
    const hair_consts& k) {
  vec2f u0 = demux_float(rn.x), u1 = demux_float(rn.y);
  const float* ap_pdf = t.ap_pdf;
  int p = 0;
  for (p = 0; p < p_max; p++) {
    if (u0.x < ap_pdf[p]) break;
    u0.x -= ap_pdf[p];
  }
  float sin_theta_op, cos_theta_op;
  tilt(brdf, p, t.sin_theta_o, t.cos_theta_o, sin_theta_op, cos_theta_op);
  u1.x            = fmax_(u1.x, 1e-5f);
  float vp        = brdf.v[p];
  float cos_theta = 1 + vp * logf_(u1.x + (1 - u1.x) * k.exp_m2_inv_v[p]);
  float sin_theta = safe_sqrt(1 - sqr(cos_theta));
  float cos_phi   = cosf_(2 * pif * u1.y);
  float sin_theta_i = -cos_theta * sin_theta_op + sin_theta * cos_phi * cos_theta_op;
  float cos_theta_i = safe_sqrt(1 - sqr(sin_theta_i));
  float dphi;
  if (p < p_max)
    dphi = phi_p(p, brdf.gamma_o, t.gamma_t) + sample_trimmed_logistic(u0.y, brdf.s, -pif, pif, k.cdf_lo, k.cdf_width);
  else
    dphi = 2 * pif * u0.y;
  float phi_i    = t.phi_o + dphi;
  vec3f incoming = {sin_theta_i, cos_theta_i * cosf_(phi_i), cos_theta_i * sinf_(phi_i)};
  // inverse(world_to_brdf) is the rigid inverse again = its transpose (:478)
  return transform_direction(transpose_rot(brdf.world_to_brdf), incoming);
}
YH_HD vec3f sample_hair_scattering(const hair_brdf& brdf, const hair_wo_terms& t, const vec2f& rn) {
  return sample_hair_scattering(brdf, t, rn, make_hair_consts(brdf));
}
YH_HD vec3f sample_hair_scattering(const hair_brdf& brdf, const vec3f& outgoing, const vec2f& rn) {
  return sample_hair_scattering(brdf, hair_prepare_wo(brdf, outgoing), rn);
}

// sample_hair_scattering_pdf, :481-551
YH_HD float sample_hair_scattering_pdf(const hair_brdf& brdf, const hair_wo_terms& t, const vec3f& incoming_) {
  vec3f incoming    = transform_direction(brdf.world_to_brdf, incoming_);
  float sin_theta_i = incoming.x;
  float cos_theta_i = safe_sqrt(1 - sqr(sin_theta_i));
  float phi_i       = atan2f_(incoming.z, incoming.y);
  const float* ap_pdf = t.ap_pdf;
  float phi = phi_i - t.phi_o;
  float pdf = 0.0f;
  for (int p = 0; p < p_max; p++) {
    float sin_theta_op, cos_theta_op;
    tilt(brdf, p, t.sin_theta_o, t.cos_theta_o, sin_theta_op, cos_theta_op);
    cos_theta_op = fabs_(cos_theta_op);
    pdf += mp(cos_theta_i, cos_theta_op, sin_theta_i, sin_theta_op, brdf.v[p]) * ap_pdf[p] *
           np(phi, p, brdf.s, brdf.gamma_o, t.gamma_t);
  }
  pdf += mp(cos_theta_i, t.cos_theta_o, sin_theta_i, t.sin_theta_o, brdf.v[p_max]) * ap_pdf[p_max] *
         div_(1.0f, (2 * pif));
  return pdf;
}
// eval_hair_scattering (:255-336) and sample_hair_scattering_pdf (:481-551) at
// the SAME incoming direction, as trace_path asks for them (:1449-1451).  Both
// sums run over the same Mp / Np factors; each factor is formed once, with the
// reference's arithmetic, and feeds both sums in the reference's order, so f
// and pdf carry the bits the two separate calls produce.
YH_HD void eval_hair_scattering_and_pdf(const hair_brdf& brdf, const hair_wo_terms& t, const vec3f& incoming_,
    vec3f& f, float& pdf, const hair_consts& k) {
  vec3f incoming    = transform_direction(brdf.world_to_brdf, incoming_);
  float sin_theta_i = incoming.x;
  float cos_theta_i = safe_sqrt(1 - sqr(sin_theta_i));
  float phi_i       = atan2f_(incoming.z, incoming.y);
  float phi         = phi_i - t.phi_o;
  f   = {0, 0, 0};
  pdf = 0.0f;
  for (int p = 0; p < p_max; p++) {
    float sin_theta_op, cos_theta_op;
    tilt(brdf, p, t.sin_theta_o, t.cos_theta_o, sin_theta_op, cos_theta_op);
    cos_theta_op = fabs_(cos_theta_op);
    float m = mp(cos_theta_i, cos_theta_op, sin_theta_i, sin_theta_op, brdf.v[p], p, k);
    float n = np(phi, p, brdf.s, brdf.gamma_o, t.gamma_t, k.cdf_width);
    f = f + m * t.ap[p] * n;
    pdf += m * t.ap_pdf[p] * n;
  }
  float m = mp(cos_theta_i, t.cos_theta_o, sin_theta_i, t.sin_theta_o, brdf.v[p_max], p_max, k);
  f = f + m * t.ap[p_max] / (2 * pif);
  pdf += m * t.ap_pdf[p_max] * div_(1.0f, (2 * pif));
}
YH_HD float sample_hair_scattering_pdf(const hair_brdf& brdf, const vec3f& outgoing, const vec3f& incoming) {
  return sample_hair_scattering_pdf(brdf, hair_prepare_wo(brdf, outgoing), incoming);
}

}  // namespace yh

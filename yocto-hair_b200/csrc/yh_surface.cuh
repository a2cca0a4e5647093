// Surface lobes of the reference's material model for the triangles that share a scene with the
// hair: diffuse, specular (GGX microfacet, dielectric Fresnel), metal (conductor
// Fresnel), thin-sheet transmission, and their delta forms for roughness 0.  Restates eval_brdf
// (yocto_pathtrace.cpp:405-471), eval_brdfcos / sample_brdfcos / sample_brdfcos_pdf and the delta
// trio (:1069-1280) and the lobe functions of yocto_math.h:4215-4760, same operation order, so
// renders stay bit-identical to the reference.  Refraction (thin = false with transmission) and the
// homogeneous volume it opens (eval_vsdf :505-526, trace_path :1403-1414, :1458-1497) are here too.
// Textures multiply the material factors before these functions see them (shade_vertex, yh_trace.cu).
#pragma once
#include "yh_libm.cuh"

namespace yh {

// yocto_pathtrace.cpp:372-393 without the hair block
struct surface_brdf {
  vec3f diffuse, specular, metal, transmission, refraction;
  float roughness, ior;
  vec3f meta, metak;
  float diffuse_pdf, specular_pdf, metal_pdf, transmission_pdf, refraction_pdf;
};

// fresnel_dielectric, yocto_math.h:4215-4235 (symmetric in its two directions)
YH_HD float fresnel_dielectric(float eta, const vec3f& normal, const vec3f& outgoing) {
  float cosw  = fabs_(dot(normal, outgoing));
  float sin2  = 1 - cosw * cosw;
  float eta2  = eta * eta;
  float cos2t = 1 - div_(sin2, eta2);
  if (cos2t < 0) return 1;  // tir
  float t0 = sqrt_(cos2t);
  float t1 = eta * t0;
  float t2 = eta * cosw;
  float rs = div_((cosw - t1), (cosw + t1));
  float rp = div_((t0 - t2), (t0 + t2));
  return div_pow2((rs * rs + rp * rp), 0.5f);
}
// fresnel_conductor, :4238-4263: the reference works on vec3f, every operation is per component
YH_HD vec3f fresnel_conductor(const vec3f& eta, const vec3f& etak, const vec3f& normal, const vec3f& outgoing) {
  float cosw = dot(normal, outgoing);
  if (cosw <= 0) return {0, 0, 0};
  cosw       = clampf(cosw, -1.0f, 1.0f);
  float cos2 = cosw * cosw;
  float sin2 = clampf(1 - cos2, 0.0f, 1.0f);
  auto  one  = [&](float e, float k) {
    float eta2 = e * e, etak2 = k * k;
    float t0       = eta2 - etak2 - sin2;
    float a2plusb2 = sqrt_(t0 * t0 + 4 * eta2 * etak2);
    float t1       = a2plusb2 + cos2;
    float a        = sqrt_(div_pow2((a2plusb2 + t0), 0.5f));
    float t2       = 2 * a * cosw;
    float rs       = div_((t1 - t2), (t1 + t2));
    float t3       = cos2 * a2plusb2 + sin2 * sin2;
    float t4       = t2 * sin2;
    float rp       = div_(rs * (t3 - t4), (t3 + t4));
    return div_pow2((rp + rs), 0.5f);
  };
  return {one(eta.x, etak.x), one(eta.y, etak.y), one(eta.z, etak.z)};
}
// reflectivity_to_eta, :4270-4273
YH_HD vec3f reflectivity_to_eta(const vec3f& reflectivity) {
  auto one = [](float r) {
    r = clampf(r, 0.0f, 0.99f);
    return div_((1 + sqrt_(r)), (1 - sqrt_(r)));
  };
  return {one(reflectivity.x), one(reflectivity.y), one(reflectivity.z)};
}
// reflect, :2059-2061
YH_HD vec3f reflect(const vec3f& w, const vec3f& n) { return -w + 2 * dot(n, w) * n; }
// refract, :2062-2067
YH_HD vec3f refract(const vec3f& w, const vec3f& n, float inv_eta) {
  float cosine = dot(n, w);
  float k      = 1 + inv_eta * inv_eta * (cosine * cosine - 1);
  if (k < 0) return {0, 0, 0};  // tir
  return -w * inv_eta + (inv_eta * cosine - sqrt_(k)) * n;
}

// microfacet_distribution (GGX), :4307-4322
YH_HD float microfacet_distribution(float roughness, const vec3f& normal, const vec3f& halfway) {
  float cosine = dot(normal, halfway);
  if (cosine <= 0) return 0;
  float roughness2 = roughness * roughness;
  float cosine2    = cosine * cosine;
  return div_(roughness2, (pif * (cosine2 * roughness2 + 1 - cosine2) * (cosine2 * roughness2 + 1 - cosine2)));
}
// microfacet_shadowing1 / microfacet_shadowing (GGX), :4325-4351
YH_HD float microfacet_shadowing1(float roughness, const vec3f& normal, const vec3f& halfway, const vec3f& direction) {
  float cosine  = dot(normal, direction);
  float cosineh = dot(halfway, direction);
  if (cosine * cosineh <= 0) return 0;
  float roughness2 = roughness * roughness;
  float cosine2    = cosine * cosine;
  return div_(2 * fabs_(cosine), (fabs_(cosine) + sqrt_(cosine2 - roughness2 * cosine2 + roughness2)));
}
YH_HD float microfacet_shadowing(float roughness, const vec3f& normal, const vec3f& halfway, const vec3f& outgoing,
    const vec3f& incoming) {
  return microfacet_shadowing1(roughness, normal, halfway, outgoing) *
         microfacet_shadowing1(roughness, normal, halfway, incoming);
}
// sample_microfacet (GGX, not the visible-normal variant), :4354-4367
YH_HD vec3f sample_microfacet(float roughness, const vec3f& normal, const vec2f& rn) {
  float phi   = 2 * pif * rn.x;
  float theta = atanf_(roughness * sqrt_(div_(rn.y, (1 - rn.y))));
  vec3f local = {cosf_(phi) * sinf_(theta), sinf_(phi) * sinf_(theta), cosf_(theta)};
  return transform_direction(basis_fromz(normal), local);
}
// sample_microfacet_pdf, :4370-4375
YH_HD float sample_microfacet_pdf(float roughness, const vec3f& normal, const vec3f& halfway) {
  float cosine = dot(normal, halfway);
  if (cosine < 0) return 0;
  return microfacet_distribution(roughness, normal, halfway) * cosine;
}

// eval_microfacet_reflection (dielectric), :4441-4451; vec3f{1} * F * D * G is the same scalar three times
YH_HD vec3f eval_microfacet_reflection(float ior, float roughness, const vec3f& normal, const vec3f& outgoing,
    const vec3f& incoming) {
  if (dot(normal, incoming) <= 0 || dot(normal, outgoing) <= 0) return {0, 0, 0};
  vec3f halfway = normalize(incoming + outgoing);
  float F       = fresnel_dielectric(ior, halfway, incoming);
  float D       = microfacet_distribution(roughness, normal, halfway);
  float G       = microfacet_shadowing(roughness, normal, halfway, outgoing, incoming);
  float v       = div_(1.0f * F * D * G, (4 * dot(normal, outgoing) * dot(normal, incoming))) * dot(normal, incoming);
  return {v, v, v};
}
// eval_microfacet_reflection (conductor), :4454-4464
YH_HD vec3f eval_microfacet_reflection(const vec3f& eta, const vec3f& etak, float roughness, const vec3f& normal,
    const vec3f& outgoing, const vec3f& incoming) {
  if (dot(normal, incoming) <= 0 || dot(normal, outgoing) <= 0) return {0, 0, 0};
  vec3f halfway = normalize(incoming + outgoing);
  vec3f F       = fresnel_conductor(eta, etak, halfway, incoming);
  float D       = microfacet_distribution(roughness, normal, halfway);
  float G       = microfacet_shadowing(roughness, normal, halfway, outgoing, incoming);
  return F * D * G / (4 * dot(normal, outgoing) * dot(normal, incoming)) * dot(normal, incoming);
}
// eval_microfacet_transmission (thin sheet), :4467-4480
YH_HD vec3f eval_microfacet_transmission(float roughness, const vec3f& normal, const vec3f& outgoing, const vec3f& incoming) {
  if (dot(normal, incoming) >= 0 || dot(normal, outgoing) <= 0) return {0, 0, 0};
  vec3f reflected = reflect(-incoming, normal);
  vec3f halfway   = normalize(reflected + outgoing);
  float D         = microfacet_distribution(roughness, normal, halfway);
  float G         = microfacet_shadowing(roughness, normal, halfway, outgoing, reflected);
  float v         = div_(1.0f * D * G, (4 * dot(normal, outgoing) * dot(normal, reflected))) * (dot(normal, reflected));
  return {v, v, v};
}
// eval_microfacet_refraction, :4483-4514.  pow(x, 2) is yocto::math::pow(float, float) = powf(x, 2.0f), which
// GCC folds to x * x at -O1 and above (the oracle build is -O3).
YH_HD vec3f eval_microfacet_refraction(float ior, float roughness, const vec3f& normal, const vec3f& outgoing,
    const vec3f& incoming) {
  bool  entering  = dot(normal, outgoing) >= 0;
  vec3f up_normal = entering ? normal : -normal;
  float rel_ior   = entering ? ior : div_(1.0f, ior);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0) {
    vec3f halfway = normalize(incoming + outgoing);
    float F       = fresnel_dielectric(rel_ior, halfway, outgoing);
    float D       = microfacet_distribution(roughness, up_normal, halfway);
    float G       = microfacet_shadowing(roughness, up_normal, halfway, outgoing, incoming);
    float v = div_(1.0f * F * D * G, fabs_(4 * dot(normal, outgoing) * dot(normal, incoming))) * fabs_(dot(normal, incoming));
    return {v, v, v};
  } else {
    vec3f halfway = -normalize(rel_ior * incoming + outgoing) * (entering ? 1.0f : -1.0f);
    float F       = fresnel_dielectric(rel_ior, halfway, outgoing);
    float D       = microfacet_distribution(roughness, up_normal, halfway);
    float G       = microfacet_shadowing(roughness, up_normal, halfway, outgoing, incoming);
    float t       = rel_ior * dot(halfway, incoming) + dot(halfway, outgoing);
    float v = div_(1.0f * fabs_(div_((dot(outgoing, halfway) * dot(incoming, halfway)), (dot(outgoing, normal) * dot(incoming, normal)))) *
                       (1 - F) * D * G,
                  t * t) *
              fabs_(dot(normal, incoming));
    return {v, v, v};
  }
}
// sample_microfacet_refraction, :4562-4573
YH_HD vec3f sample_microfacet_refraction(float ior, float roughness, const vec3f& normal, const vec3f& outgoing, float rnl,
    const vec2f& rn) {
  bool  entering  = dot(normal, outgoing) >= 0;
  vec3f up_normal = entering ? normal : -normal;
  vec3f halfway   = sample_microfacet(roughness, up_normal, rn);
  if (rnl < fresnel_dielectric(entering ? ior : div_(1.0f, ior), halfway, outgoing)) return reflect(outgoing, halfway);
  return refract(outgoing, halfway, entering ? div_(1.0f, ior) : ior);
}
// sample_microfacet_refraction_pdf, :4623-4641
YH_HD float sample_microfacet_refraction_pdf(float ior, float roughness, const vec3f& normal, const vec3f& outgoing,
    const vec3f& incoming) {
  bool  entering  = dot(normal, outgoing) >= 0;
  vec3f up_normal = entering ? normal : -normal;
  float rel_ior   = entering ? ior : div_(1.0f, ior);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0) {
    vec3f halfway = normalize(incoming + outgoing);
    return div_(fresnel_dielectric(rel_ior, halfway, outgoing) * sample_microfacet_pdf(roughness, up_normal, halfway),
        (4 * fabs_(dot(outgoing, halfway))));
  } else {
    vec3f halfway = -normalize(rel_ior * incoming + outgoing) * (entering ? 1.0f : -1.0f);
    float num     = (1 - fresnel_dielectric(rel_ior, halfway, outgoing)) * sample_microfacet_pdf(roughness, up_normal, halfway) *
                fabs_(dot(halfway, outgoing));
    float t = rel_ior * dot(halfway, incoming) + dot(halfway, outgoing);
    return div_(num, t * t);
  }
}
// eval_delta_refraction / sample_delta_refraction / sample_delta_refraction_pdf, :4664-4678, :4706-4717, :4741-4754
YH_HD vec3f eval_delta_refraction(float ior, const vec3f& normal, const vec3f& outgoing, const vec3f& incoming) {
  if (fabs_(ior - 1) < 1e-3f) return dot(normal, incoming) * dot(normal, outgoing) <= 0 ? vec3f{1, 1, 1} : vec3f{0, 0, 0};
  bool  entering  = dot(normal, outgoing) >= 0;
  vec3f up_normal = entering ? normal : -normal;
  float rel_ior   = entering ? ior : div_(1.0f, ior);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0) {
    float f = 1.0f * fresnel_dielectric(rel_ior, up_normal, outgoing);
    return {f, f, f};
  }
  float f = 1.0f * div_(1.0f, (rel_ior * rel_ior)) * (1 - fresnel_dielectric(rel_ior, up_normal, outgoing));
  return {f, f, f};
}
YH_HD vec3f sample_delta_refraction(float ior, const vec3f& normal, const vec3f& outgoing, float rnl) {
  if (fabs_(ior - 1) < 1e-3f) return -outgoing;
  bool  entering  = dot(normal, outgoing) >= 0;
  vec3f up_normal = entering ? normal : -normal;
  float rel_ior   = entering ? ior : div_(1.0f, ior);
  if (rnl < fresnel_dielectric(rel_ior, up_normal, outgoing)) return reflect(outgoing, up_normal);
  return refract(outgoing, up_normal, div_(1.0f, rel_ior));
}
YH_HD float sample_delta_refraction_pdf(float ior, const vec3f& normal, const vec3f& outgoing, const vec3f& incoming) {
  if (fabs_(ior - 1) < 1e-3f) return dot(normal, incoming) * dot(normal, outgoing) < 0 ? 1.0f : 0.0f;
  bool  entering  = dot(normal, outgoing) >= 0;
  vec3f up_normal = entering ? normal : -normal;
  float rel_ior   = entering ? ior : div_(1.0f, ior);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0) return fresnel_dielectric(rel_ior, up_normal, outgoing);
  return (1 - fresnel_dielectric(rel_ior, up_normal, outgoing));
}

// sample_microfacet_reflection (both overloads) / _transmission, :4530-4559
YH_HD vec3f sample_microfacet_reflection(float roughness, const vec3f& normal, const vec3f& outgoing, const vec2f& rn) {
  if (dot(normal, outgoing) <= 0) return {0, 0, 0};
  return reflect(outgoing, sample_microfacet(roughness, normal, rn));
}
YH_HD vec3f sample_microfacet_transmission(float roughness, const vec3f& normal, const vec3f& outgoing, const vec2f& rn) {
  if (dot(normal, outgoing) <= 0) return {0, 0, 0};
  vec3f halfway   = sample_microfacet(roughness, normal, rn);
  vec3f reflected = reflect(outgoing, halfway);
  return -reflect(reflected, normal);
}
// sample_microfacet_reflection_pdf (both overloads) / _transmission_pdf, :4590-4620
YH_HD float sample_microfacet_reflection_pdf(float roughness, const vec3f& normal, const vec3f& outgoing,
    const vec3f& incoming) {
  if (dot(normal, incoming) <= 0 || dot(normal, outgoing) <= 0) return 0;
  vec3f halfway = normalize(outgoing + incoming);
  return div_(sample_microfacet_pdf(roughness, normal, halfway), (4 * fabs_(dot(outgoing, halfway))));
}
YH_HD float sample_microfacet_transmission_pdf(float roughness, const vec3f& normal, const vec3f& outgoing,
    const vec3f& incoming) {
  if (dot(normal, incoming) >= 0 || dot(normal, outgoing) <= 0) return 0;
  vec3f reflected = reflect(-incoming, normal);
  vec3f halfway   = normalize(reflected + outgoing);
  float d         = sample_microfacet_pdf(roughness, normal, halfway);
  return div_(d, (4 * fabs_(dot(outgoing, halfway))));
}

// The material at a surface point, eval_brdf :405-471 with every texture at its {1,1,1} default.
struct surface_material {
  vec3f color;
  float specular, metallic, roughness, ior, transmission;
  bool  thin;  // material.thin || !material.transmission (:425: the UNTEXTURED transmission decides)
};
YH_HD surface_brdf eval_surface_brdf(const surface_material& m, const vec3f& normal, const vec3f& outgoing) {
  surface_brdf brdf;
  vec3f base   = m.color;
  vec3f weight = {1, 1, 1};
  bool  thin   = m.thin;
  brdf.metal   = weight * m.metallic;
  weight       = weight * (1 - m.metallic);
  brdf.refraction = thin ? vec3f{0, 0, 0} : (weight * m.transmission);
  weight          = weight * (1 - (thin ? 0.0f : m.transmission));
  brdf.specular = weight * m.specular;
  weight        = weight * (1 - m.specular * fresnel_dielectric(m.ior, outgoing, normal));
  brdf.transmission = thin ? (weight * m.transmission * base) : vec3f{0, 0, 0};
  weight            = weight * (1 - (thin ? m.transmission : 0.0f));
  brdf.diffuse      = weight * base;
  brdf.meta         = reflectivity_to_eta(base);
  brdf.metak        = {0, 0, 0};
  brdf.roughness    = m.roughness * m.roughness;
  brdf.ior          = m.ior;
  if (nonzero(brdf.diffuse) || brdf.roughness != 0) brdf.roughness = clampf(brdf.roughness, 0.03f * 0.03f, 1.0f);
  if (!nonzero(brdf.specular) && !nonzero(brdf.metal) && !nonzero(brdf.transmission) && !nonzero(brdf.refraction))
    brdf.roughness = 1;
  brdf.diffuse_pdf      = max3(brdf.diffuse);
  brdf.specular_pdf     = max3(brdf.specular * fresnel_dielectric(brdf.ior, normal, outgoing));
  brdf.metal_pdf        = max3(brdf.metal * fresnel_conductor(brdf.meta, brdf.metak, normal, outgoing));
  brdf.transmission_pdf = max3(brdf.transmission);
  brdf.refraction_pdf   = max3(brdf.refraction);
  float pdf_sum = brdf.diffuse_pdf + brdf.specular_pdf + brdf.metal_pdf + brdf.transmission_pdf + brdf.refraction_pdf;
  if (pdf_sum != 0) {
    brdf.diffuse_pdf      = div_(brdf.diffuse_pdf, pdf_sum);
    brdf.specular_pdf     = div_(brdf.specular_pdf, pdf_sum);
    brdf.metal_pdf        = div_(brdf.metal_pdf, pdf_sum);
    brdf.transmission_pdf = div_(brdf.transmission_pdf, pdf_sum);
    brdf.refraction_pdf   = div_(brdf.refraction_pdf, pdf_sum);
  }
  return brdf;
}
YH_HD bool is_delta(const surface_brdf& brdf) { return brdf.roughness == 0; }  // :495

// eval_brdfcos, :1069-1103 (surface branch)
YH_HD vec3f eval_brdfcos(const surface_brdf& brdf, const vec3f& normal, const vec3f& outgoing, const vec3f& incoming) {
  if (brdf.roughness == 0) return {0, 0, 0};
  vec3f brdfcos = {0, 0, 0};
  if (nonzero(brdf.diffuse)) {  // eval_diffuse_reflection, yocto_math.h:4427
    float ni = dot(normal, incoming), no = dot(normal, outgoing);
    brdfcos  = brdfcos + brdf.diffuse * ((ni <= 0 || no <= 0) ? vec3f{0, 0, 0} : vec3f{1, 1, 1} / pif * ni);
  }
  if (nonzero(brdf.specular))
    brdfcos = brdfcos + brdf.specular * eval_microfacet_reflection(brdf.ior, brdf.roughness, normal, outgoing, incoming);
  if (nonzero(brdf.metal))
    brdfcos = brdfcos + brdf.metal * eval_microfacet_reflection(brdf.meta, brdf.metak, brdf.roughness, normal, outgoing, incoming);
  if (nonzero(brdf.transmission))
    brdfcos = brdfcos + brdf.transmission * eval_microfacet_transmission(brdf.roughness, normal, outgoing, incoming);
  if (nonzero(brdf.refraction))
    brdfcos = brdfcos + brdf.refraction * eval_microfacet_refraction(brdf.ior, brdf.roughness, normal, outgoing, incoming);
  return brdfcos;
}
// eval_delta, :1105-1129
YH_HD vec3f eval_delta(const surface_brdf& brdf, const vec3f& normal, const vec3f& outgoing, const vec3f& incoming) {
  if (brdf.roughness != 0) return {0, 0, 0};
  vec3f brdfcos = {0, 0, 0};
  float ni = dot(normal, incoming), no = dot(normal, outgoing);
  if (nonzero(brdf.specular) && !nonzero(brdf.refraction)) {  // eval_delta_reflection, yocto_math.h:4644
    float f = (ni <= 0 || no <= 0) ? 0.0f : 1.0f * fresnel_dielectric(brdf.ior, normal, outgoing);
    brdfcos = brdfcos + brdf.specular * vec3f{f, f, f};
  }
  if (nonzero(brdf.metal))
    brdfcos = brdfcos + brdf.metal * ((ni <= 0 || no <= 0) ? vec3f{0, 0, 0} : fresnel_conductor(brdf.meta, brdf.metak, normal, outgoing));
  if (nonzero(brdf.transmission))
    brdfcos = brdfcos + brdf.transmission * ((ni >= 0 || no <= 0) ? vec3f{0, 0, 0} : vec3f{1, 1, 1});
  if (nonzero(brdf.refraction)) brdfcos = brdfcos + brdf.refraction * eval_delta_refraction(brdf.ior, normal, outgoing, incoming);
  return brdfcos;
}
// sample_hemisphere_cos(normal, ruv), yocto_math.h:4873-4879
YH_HD vec3f sample_hemisphere_cos(const vec3f& normal, const vec2f& ruv) {
  float z   = sqrt_(ruv.y);
  float r   = sqrt_(1 - z * z);
  float phi = 2 * pif * ruv.x;
  vec3f local = {r * cosf_(phi), r * sinf_(phi), z};
  return transform_direction(basis_fromz(normal), local);
}
// sample_brdfcos, :1132-1176
YH_HD vec3f sample_brdfcos(const surface_brdf& brdf, const vec3f& normal, const vec3f& outgoing, float rnl, const vec2f& rn) {
  if (brdf.roughness == 0) return {0, 0, 0};
  float cdf = 0.0f;
  if (brdf.diffuse_pdf != 0) {
    cdf += brdf.diffuse_pdf;
    if (rnl < cdf) {  // sample_diffuse_reflection, yocto_math.h:4515
      if (dot(normal, outgoing) <= 0) return {0, 0, 0};
      return sample_hemisphere_cos(normal, rn);
    }
  }
  if (brdf.specular_pdf != 0 && brdf.refraction_pdf == 0) {
    cdf += brdf.specular_pdf;
    if (rnl < cdf) return sample_microfacet_reflection(brdf.roughness, normal, outgoing, rn);
  }
  if (brdf.metal_pdf != 0) {
    cdf += brdf.metal_pdf;
    if (rnl < cdf) return sample_microfacet_reflection(brdf.roughness, normal, outgoing, rn);
  }
  if (brdf.transmission_pdf != 0) {
    cdf += brdf.transmission_pdf;
    if (rnl < cdf) return sample_microfacet_transmission(brdf.roughness, normal, outgoing, rn);
  }
  if (brdf.refraction_pdf != 0) {
    cdf += brdf.refraction_pdf;
    if (rnl < cdf) return sample_microfacet_refraction(brdf.ior, brdf.roughness, normal, outgoing, rnl, rn);
  }
  return {0, 0, 0};
}
// sample_delta, :1178-1215 (cdf starts at the diffuse weight)
YH_HD vec3f sample_delta(const surface_brdf& brdf, const vec3f& normal, const vec3f& outgoing, float rnl) {
  if (brdf.roughness != 0) return {0, 0, 0};
  float cdf = 0.0f;
  cdf += brdf.diffuse_pdf;
  if (brdf.specular_pdf != 0 && brdf.refraction_pdf == 0) {
    cdf += brdf.specular_pdf;
    if (rnl < cdf) return (dot(normal, outgoing) <= 0) ? vec3f{0, 0, 0} : reflect(outgoing, normal);
  }
  if (brdf.metal_pdf != 0) {
    cdf += brdf.metal_pdf;
    if (rnl < cdf) return (dot(normal, outgoing) <= 0) ? vec3f{0, 0, 0} : reflect(outgoing, normal);
  }
  if (brdf.transmission_pdf != 0) {
    cdf += brdf.transmission_pdf;
    if (rnl < cdf) return (dot(normal, outgoing) <= 0) ? vec3f{0, 0, 0} : -outgoing;
  }
  if (brdf.refraction_pdf != 0) {
    cdf += brdf.refraction_pdf;
    if (rnl < cdf) return sample_delta_refraction(brdf.ior, normal, outgoing, rnl);
  }
  return {0, 0, 0};
}
// sample_brdfcos_pdf, :1218-1255
YH_HD float sample_brdfcos_pdf(const surface_brdf& brdf, const vec3f& normal, const vec3f& outgoing, const vec3f& incoming) {
  if (brdf.roughness == 0) return 0;
  float pdf = 0.0f;
  if (brdf.diffuse_pdf != 0) {  // sample_diffuse_reflection_pdf -> sample_hemisphere_cos_pdf, yocto_math.h:4572, :4880
    float ni = dot(normal, incoming), no = dot(normal, outgoing);
    pdf += brdf.diffuse_pdf * ((ni <= 0 || no <= 0) ? 0.0f : div_(ni, pif));
  }
  if (brdf.specular_pdf != 0 && brdf.refraction_pdf == 0)
    pdf += brdf.specular_pdf * sample_microfacet_reflection_pdf(brdf.roughness, normal, outgoing, incoming);
  if (brdf.metal_pdf != 0) pdf += brdf.metal_pdf * sample_microfacet_reflection_pdf(brdf.roughness, normal, outgoing, incoming);
  if (brdf.transmission_pdf != 0)
    pdf += brdf.transmission_pdf * sample_microfacet_transmission_pdf(brdf.roughness, normal, outgoing, incoming);
  if (brdf.refraction_pdf != 0)
    pdf += brdf.refraction_pdf * sample_microfacet_refraction_pdf(brdf.ior, brdf.roughness, normal, outgoing, incoming);
  return pdf;
}
// sample_delta_pdf, :1257-1280
YH_HD float sample_delta_pdf(const surface_brdf& brdf, const vec3f& normal, const vec3f& outgoing, const vec3f& incoming) {
  if (brdf.roughness != 0) return 0;
  float pdf = 0.0f;
  float ni = dot(normal, incoming), no = dot(normal, outgoing);
  if (brdf.specular_pdf != 0 && brdf.refraction_pdf == 0) pdf += brdf.specular_pdf * ((ni <= 0 || no <= 0) ? 0.0f : 1.0f);
  if (brdf.metal_pdf != 0) pdf += brdf.metal_pdf * ((ni <= 0 || no <= 0) ? 0.0f : 1.0f);
  if (brdf.transmission_pdf != 0) pdf += brdf.transmission_pdf * ((ni >= 0 || no <= 0) ? 0.0f : 1.0f);
  if (brdf.refraction_pdf != 0) pdf += brdf.refraction_pdf * sample_delta_refraction_pdf(brdf.ior, normal, outgoing, incoming);
  return pdf;
}

// ---- homogeneous volumes behind a refractive surface -------------------------------------------
// vsdf, yocto_pathtrace.cpp:498-502; the reference's volume stack never holds more than one entry
// (:1460-1467 pushes only when it is empty, otherwise pops), so the path carries one vsdf and a flag.
struct vsdf {
  vec3f density, scatter;
  float anisotropy;
};
// eval_vsdf, :505-526 (every texture at its {1,1,1} default); called only when has_volume() holds
YH_HD vsdf eval_vsdf(const vec3f& color, float transmission, bool thin, const vec3f& scattering, float scanisotropy, float trdepth) {
  vsdf v;
  v.density = {0, 0, 0};
  if (transmission != 0 && !thin) {
    vec3f l   = {logf_(clampf(color.x, 0.0001f, 1.0f)), logf_(clampf(color.y, 0.0001f, 1.0f)), logf_(clampf(color.z, 0.0001f, 1.0f))};
    v.density = -l / trdepth;
  }
  v.scatter    = scattering;
  v.anisotropy = scanisotropy;
  return v;
}
// sample_transmittance / sample_transmittance_pdf / eval_transmittance, yocto_math.h:4757-4781
YH_HD float sample_transmittance(const vec3f& density, float max_distance, float rl, float rd) {
  int   channel  = clampi((int)(rl * 3), 0, 2);
  float dc       = channel == 0 ? density.x : (channel == 1 ? density.y : density.z);
  float distance = (dc == 0) ? flt_max : div_(-logf_(1 - rd), dc);
  return fmin_(distance, max_distance);
}
YH_HD vec3f eval_transmittance(const vec3f& density, float distance) {
  vec3f a = -density * distance;
  return {expf_(a.x), expf_(a.y), expf_(a.z)};
}
YH_HD float sample_transmittance_pdf(const vec3f& density, float distance, float max_distance) {
  if (distance < max_distance) {
    vec3f e = density * eval_transmittance(density, distance);
    return div_((e.x + e.y + e.z), 3.0f);
  }
  vec3f e = eval_transmittance(density, max_distance);
  return div_((e.x + e.y + e.z), 3.0f);
}
// eval_phasefunction / sample_phasefunction (Henyey-Greenstein), yocto_math.h:4784-4812
YH_HD float eval_phasefunction(float anisotropy, const vec3f& outgoing, const vec3f& incoming) {
  float cosine = -dot(outgoing, incoming);
  float denom  = 1 + anisotropy * anisotropy - 2 * anisotropy * cosine;
  return div_((1 - anisotropy * anisotropy), (4 * pif * denom * sqrt_(denom)));
}
YH_HD vec3f sample_phasefunction(float anisotropy, const vec3f& outgoing, const vec2f& rn) {
  float cos_theta = 0.0f;
  if (fabs_(anisotropy) < 1e-3f) {
    cos_theta = 1 - 2 * rn.y;
  } else {
    float square = div_((1 - anisotropy * anisotropy), (1 + anisotropy - 2 * anisotropy * rn.y));
    cos_theta    = div_((1 + anisotropy * anisotropy - square * square), (2 * anisotropy));
  }
  float sin_theta = sqrt_(fmax_(0.0f, 1 - cos_theta * cos_theta));
  float phi       = 2 * pif * rn.x;
  vec3f local     = {sin_theta * cosf_(phi), sin_theta * sinf_(phi), cos_theta};
  return transform_vector(basis_fromz(-outgoing), local);  // mat3f * vec3f: not renormalised
}
// eval_scattering / sample_scattering / sample_scattering_pdf, yocto_pathtrace.cpp:1360-1378
YH_HD vec3f eval_scattering(const vsdf& v, const vec3f& outgoing, const vec3f& incoming) {
  if (!nonzero(v.density)) return {0, 0, 0};
  return v.scatter * v.density * eval_phasefunction(v.anisotropy, outgoing, incoming);
}
YH_HD vec3f sample_scattering(const vsdf& v, const vec3f& outgoing, const vec2f& rn) {
  if (!nonzero(v.density)) return {0, 0, 0};
  return sample_phasefunction(v.anisotropy, outgoing, rn);
}
YH_HD float sample_scattering_pdf(const vsdf& v, const vec3f& outgoing, const vec3f& incoming) {
  if (!nonzero(v.density)) return 0;
  return eval_phasefunction(v.anisotropy, outgoing, incoming);
}

}  // namespace yh

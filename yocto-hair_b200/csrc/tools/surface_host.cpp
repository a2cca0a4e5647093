// Host build of csrc/yh_surface.cuh (the surface lobes, refraction and volume functions of the device):
// a CPU-only test compares them, function by function and bit for bit, with the reference's own
// yocto_math.h functions called through oracle/ref_driver.cpp::ref_surface_batch (tests/test_surface_host.py).
//
// One item = 20 floats: normal[3] outgoing[3] incoming[3] rn[2] rnl ior roughness eta[3] etak[3] anisotropy.
// fn selects the function (same numbering on both sides); up to 3 floats out per item.
#include <math.h>
#include <string.h>

#include "yh_surface.cuh"

using namespace yh;

extern "C" void yh_host_surface_batch(int fn, long n, const float* in, float* out) {
  for (long k = 0; k < n; k++) {
    const float* p = in + 20 * k;
    vec3f nrm = {p[0], p[1], p[2]}, o = {p[3], p[4], p[5]}, i = {p[6], p[7], p[8]};
    vec2f rn  = {p[9], p[10]};
    float rnl = p[11], ior = p[12], rough = p[13], anis = p[19];
    vec3f eta = {p[14], p[15], p[16]}, etak = {p[17], p[18], p[11]};
    vec3f v = {0, 0, 0};
    switch (fn) {
      case 0: v.x = fresnel_dielectric(ior, nrm, o); break;
      case 1: v = fresnel_conductor(eta, etak, nrm, o); break;
      case 2: v = reflectivity_to_eta(eta); break;
      case 3: v.x = microfacet_distribution(rough, nrm, i); break;
      case 4: v.x = microfacet_shadowing(rough, nrm, normalize(o + i), o, i); break;
      case 5: v = sample_microfacet(rough, nrm, rn); break;
      case 6: v = eval_microfacet_reflection(ior, rough, nrm, o, i); break;
      case 7: v = eval_microfacet_reflection(eta, etak, rough, nrm, o, i); break;
      case 8: v = eval_microfacet_transmission(rough, nrm, o, i); break;
      case 9: v = eval_microfacet_refraction(ior, rough, nrm, o, i); break;
      case 10: v = sample_microfacet_reflection(rough, nrm, o, rn); break;
      case 11: v = sample_microfacet_transmission(rough, nrm, o, rn); break;
      case 12: v = sample_microfacet_refraction(ior, rough, nrm, o, rnl, rn); break;
      case 13: v.x = sample_microfacet_reflection_pdf(rough, nrm, o, i); break;
      case 14: v.x = sample_microfacet_transmission_pdf(rough, nrm, o, i); break;
      case 15: v.x = sample_microfacet_refraction_pdf(ior, rough, nrm, o, i); break;
      case 16: v = eval_delta_refraction(ior, nrm, o, i); break;
      case 17: v = sample_delta_refraction(ior, nrm, o, rnl); break;
      case 18: v.x = sample_delta_refraction_pdf(ior, nrm, o, i); break;
      case 19: v.x = eval_phasefunction(anis, o, i); break;
      case 20: v = sample_phasefunction(anis, o, rn); break;
      case 21: v.x = sample_transmittance(eta, rough * 10, rn.x, rn.y); break;
      case 22: v.x = sample_transmittance_pdf(eta, rn.x * 5, rough * 10); break;
      case 23: v = eval_transmittance(eta, rn.x * 5); break;
      case 24: v = sample_hemisphere_cos(nrm, rn); break;
      case 25: v = refract(o, nrm, ior); break;
    }
    out[3 * k] = v.x, out[3 * k + 1] = v.y, out[3 * k + 2] = v.z;
  }
}

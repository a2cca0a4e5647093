// TEST INFRASTRUCTURE — not part of the product.
//
// C-callable driver around the UNMODIFIED reference sources under
// /root/reference (compiled where they lie by oracle/Makefile into
// oracle/_ref/libyh_ref.so).  It feeds the reference's own C++ API
// (yocto::pathtrace / yocto::extension) from the same flattened scene
// description the CUDA library takes (include/yocto_hair_b200.h), so tests can
// run both on identical inputs.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg may load this library.
#include <yocto_pathtrace/yocto_pathtrace.h>

#include <chrono>
#include <cstring>
#include <thread>

#include "../include/yocto_hair_b200.h"

namespace ptr = yocto::pathtrace;
namespace ext = yocto::extension;
namespace ym  = yocto::math;
using ym::frame3f;
using ym::vec2f;
using ym::vec2i;
using ym::vec3f;
using ym::vec3i;

// trace_sample is a non-static symbol of yocto_pathtrace.cpp:1676 that the
// header does not declare; declaring it here lets the driver time one pixel.
namespace yocto::pathtrace {
ym::vec4f trace_sample(ptr::state* state, const ptr::scene* scene,
    const ptr::camera* camera, const vec2i& ij, const trace_params& params);
}

static vec3f v3(const float* p) { return {p[0], p[1], p[2]}; }
static frame3f fr(const float* p) {
  return frame3f{v3(p), v3(p + 3), v3(p + 6), v3(p + 9)};
}

static ext::hair_brdf unpack(const yh_hair_brdf& b) {
  auto r    = ext::hair_brdf{};
  r.sigma_a = v3(b.sigma_a);
  r.alpha   = b.alpha;
  r.eta     = b.eta;
  r.h       = b.h;
  for (int i = 0; i < 4; i++) r.v[i] = b.v[i];
  r.s             = b.s;
  r.sin_2k_alpha  = v3(b.sin_2k_alpha);
  r.cos_2k_alpha  = v3(b.cos_2k_alpha);
  r.gamma_o       = b.gamma_o;
  r.world_to_brdf = fr(b.world_to_brdf);
  return r;
}
static void pack(const ext::hair_brdf& r, yh_hair_brdf& b) {
  memcpy(b.sigma_a, &r.sigma_a, 12);
  b.alpha = r.alpha;
  b.eta   = r.eta;
  b.h     = r.h;
  for (int i = 0; i < 4; i++) b.v[i] = r.v[i];
  b.s = r.s;
  memcpy(b.sin_2k_alpha, &r.sin_2k_alpha, 12);
  memcpy(b.cos_2k_alpha, &r.cos_2k_alpha, 12);
  b.gamma_o = r.gamma_o;
  memcpy(b.world_to_brdf, &r.world_to_brdf, 48);
}

struct ref_scene {
  ptr::scene*               scene = nullptr;
  std::vector<ptr::camera*> cams;
};
struct ref_state {
  ref_scene*        rs = nullptr;
  ptr::camera*      camera = nullptr;
  ptr::state*       state  = nullptr;
  ptr::trace_params params;
};

extern "C" {

int ref_hardware_concurrency(void) {
  return (int)std::thread::hardware_concurrency();
}

// ---- hair BSDF (yocto_extension.cpp:127-551) ------------------------------
void ref_hair_brdf_batch(const yh_hair_material* m, const float* v,
    const float* n, const float* t, int64_t count, yh_hair_brdf* out) {
  for (int64_t i = 0; i < count; i++) {
    auto mat        = ext::hair_material{};
    mat.sigma_a     = v3(m[i].sigma_a);
    mat.beta_m      = m[i].beta_m;
    mat.beta_n      = m[i].beta_n;
    mat.alpha       = m[i].alpha;
    mat.eta         = m[i].eta;
    mat.color       = v3(m[i].color);
    mat.eumelanin   = m[i].eumelanin;
    mat.pheomelanin = m[i].pheomelanin;
    pack(ext::eval_hair_brdf(mat, v[i], v3(n + 3 * i), v3(t + 3 * i)), out[i]);
  }
}
void ref_hair_eval_batch(const yh_hair_brdf* b, const float* wo,
    const float* wi, int64_t count, float* out) {
  for (int64_t i = 0; i < count; i++) {
    auto f = ext::eval_hair_scattering(unpack(b[i]), v3(wo + 3 * i), v3(wi + 3 * i));
    memcpy(out + 3 * i, &f, 12);
  }
}
void ref_hair_sample_batch(const yh_hair_brdf* b, const float* wo,
    const float* rn, int64_t count, float* out) {
  for (int64_t i = 0; i < count; i++) {
    auto w = ext::sample_hair_scattering(
        unpack(b[i]), v3(wo + 3 * i), vec2f{rn[2 * i], rn[2 * i + 1]});
    memcpy(out + 3 * i, &w, 12);
  }
}
void ref_hair_pdf_batch(const yh_hair_brdf* b, const float* wo,
    const float* wi, int64_t count, float* out) {
  for (int64_t i = 0; i < count; i++)
    out[i] = ext::sample_hair_scattering_pdf(
        unpack(b[i]), v3(wo + 3 * i), v3(wi + 3 * i));
}
// the reference's four self-tests (yocto_extension.cpp:555-693); 0 = passed
int ref_hair_selftest(int which) {
  try {
    switch (which) {
      case 0: ext::white_furnace_test(); break;
      case 1: ext::white_furnace_sampled_test(); break;
      case 2: ext::sampling_weights_test(); break;
      case 3: ext::sampling_consistency_test(); break;
      default: return 2;
    }
  } catch (const std::exception&) {
    return 1;
  }
  return 0;
}

// ---- PCG32 (yocto_math.h:1396-1442) ---------------------------------------
void ref_rng_floats(uint64_t seed, uint64_t seq, int n, float* out) {
  auto rng = ym::make_rng(seed, seq);
  for (int i = 0; i < n; i++) out[i] = ym::rand1f(rng);
}

// ---- libm as the reference binary sees it ----------------------------------
void ref_libm_batch(int fn, const float* a, const float* b, int64_t n, float* out) {
  for (int64_t i = 0; i < n; i++) {
    switch (fn) {
      case 0: out[i] = ym::exp(a[i]); break;
      case 1: out[i] = ym::log(a[i]); break;
      case 2: out[i] = ym::sin(a[i]); break;
      case 3: out[i] = ym::cos(a[i]); break;
      case 4: out[i] = ym::atan2(a[i], b[i]); break;
      case 5: out[i] = ym::acos(a[i]); break;
      case 6: out[i] = (float)::asin((double)a[i]); break;
      case 7: out[i] = (float)(1.0 / (::sinh((double)(1 / a[i])) * 2 * a[i])); break;
      case 8: out[i] = ym::fmod(a[i], b[i]); break;
      case 9: case 10: case 11: { /* the double denominator of mp() (ye:205-206) as an exact three-float expansion */
        double den = ::sinh((double)(1 / a[i])) * 2 * a[i];
        float  hi = (float)den, mid = (float)(den - (double)hi), lo = (float)(den - (double)hi - (double)mid);
        out[i] = fn == 9 ? hi : (fn == 10 ? mid : lo);
      } break;
      default: out[i] = 0;
    }
  }
}

// ---- the surface lobes / refraction / volume functions of yocto_math.h, one by one ------------
// Same item layout and numbering as yocto-hair_b200/csrc/tools/surface_host.cpp.
void ref_surface_batch(int fn, int64_t n, const float* in, float* out) {
  using namespace ym;
  for (int64_t k = 0; k < n; k++) {
    const float* p   = in + 20 * k;
    auto         nrm = vec3f{p[0], p[1], p[2]}, o = vec3f{p[3], p[4], p[5]}, i = vec3f{p[6], p[7], p[8]};
    auto         rn  = vec2f{p[9], p[10]};
    auto         rnl = p[11], ior = p[12], rough = p[13], anis = p[19];
    auto         eta = vec3f{p[14], p[15], p[16]}, etak = vec3f{p[17], p[18], p[11]};
    auto         v   = vec3f{0, 0, 0};
    switch (fn) {
      case 0: v.x = fresnel_dielectric(ior, nrm, o); break;
      case 1: v = fresnel_conductor(eta, etak, nrm, o); break;
      case 2: v = reflectivity_to_eta(eta); break;
      case 3: v.x = microfacet_distribution(rough, nrm, i); break;
      case 4: v.x = microfacet_shadowing(rough, nrm, normalize(o + i), o, i); break;
      case 5: v = sample_microfacet(rough, nrm, rn); break;
      case 6: v = eval_microfacet_reflection(ior, rough, nrm, o, i); break;
      case 7: v = eval_microfacet_reflection(eta, etak, rough, nrm, o, i); break;
      case 8: v = eval_microfacet_transmission(ior, rough, nrm, o, i); break;
      case 9: v = eval_microfacet_refraction(ior, rough, nrm, o, i); break;
      case 10: v = sample_microfacet_reflection(ior, rough, nrm, o, rn); break;
      case 11: v = sample_microfacet_transmission(ior, rough, nrm, o, rn); break;
      case 12: v = sample_microfacet_refraction(ior, rough, nrm, o, rnl, rn); break;
      case 13: v.x = sample_microfacet_reflection_pdf(ior, rough, nrm, o, i); break;
      case 14: v.x = sample_microfacet_transmission_pdf(ior, rough, nrm, o, i); break;
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

// ---- scene (yscenetrace.cpp:49-197 equivalent, from the flattened desc) ----
void* ref_scene_create(const yh_scene_desc* d) {
  auto rs   = new ref_scene{};
  rs->scene = new ptr::scene{};
  auto sc   = rs->scene;
  auto texs = std::vector<ptr::texture*>{};
  for (int i = 0; i < d->num_textures; i++) {
    auto& t   = d->textures[i];
    auto  tex = ptr::add_texture(sc);
    auto  img = yocto::image::image<vec3f>{{t.width, t.height}};
    memcpy(img.data(), t.colorf, sizeof(float) * 3 * t.width * t.height);
    ptr::set_texture(tex, img);
    texs.push_back(tex);
  }
  auto tx = [&](int i) -> ptr::texture* { return i < 0 ? nullptr : texs[i]; };
  auto shapes = std::vector<ptr::shape*>{};
  for (int i = 0; i < d->num_shapes; i++) {
    auto& s  = d->shapes[i];
    auto  sh = ptr::add_shape(sc);
    auto  p3 = [](const float* p, int n) {
      auto v = std::vector<vec3f>(n);
      if (n) memcpy(v.data(), p, 12 * (size_t)n);
      return v;
    };
    ptr::set_positions(sh, p3(s.positions, s.num_positions));
    ptr::set_normals(sh, p3(s.normals, s.num_normals));
    if (s.num_texcoords) {
      auto v = std::vector<vec2f>(s.num_texcoords);
      memcpy(v.data(), s.texcoords, 8 * (size_t)s.num_texcoords);
      ptr::set_texcoords(sh, v);
    }
    if (s.num_radius)
      ptr::set_radius(sh, std::vector<float>(s.radius, s.radius + s.num_radius));
    if (s.num_points) ptr::set_points(sh, std::vector<int>(s.points, s.points + s.num_points));
    if (s.num_lines) {
      auto v = std::vector<vec2i>(s.num_lines);
      memcpy(v.data(), s.lines, 8 * (size_t)s.num_lines);
      ptr::set_lines(sh, v);
    }
    if (s.num_triangles) {
      auto v = std::vector<vec3i>(s.num_triangles);
      memcpy(v.data(), s.triangles, 12 * (size_t)s.num_triangles);
      ptr::set_triangles(sh, v);
    }
    shapes.push_back(sh);
  }
  auto mats = std::vector<ptr::material*>{};
  for (int i = 0; i < d->num_materials; i++) {
    auto& m  = d->materials[i];
    auto  mt = ptr::add_material(sc);
    ptr::set_emission(mt, v3(m.emission), tx(m.emission_tex));
    ptr::set_color(mt, v3(m.color), tx(m.color_tex));
    ptr::set_specular(mt, m.specular, tx(m.specular_tex));
    ptr::set_ior(mt, m.ior);
    ptr::set_metallic(mt, m.metallic, tx(m.metallic_tex));
    ptr::set_transmission(
        mt, m.transmission, m.thin != 0, m.trdepth, tx(m.transmission_tex));
    ptr::set_roughness(mt, m.roughness, tx(m.roughness_tex));
    ptr::set_opacity(mt, m.opacity, tx(m.opacity_tex));
    ptr::set_thin(mt, m.thin != 0);
    ptr::set_scattering(
        mt, v3(m.scattering), m.scanisotropy, tx(m.scattering_tex));
    ptr::set_normalmap(mt, tx(m.normal_tex));
    ptr::set_eumelanin(mt, m.eumelanin);
    ptr::set_pheomelanin(mt, m.pheomelanin);
    ptr::set_sigma_a(mt, v3(m.sigma_a));
    ptr::set_beta_m(mt, m.beta_m);
    ptr::set_beta_n(mt, m.beta_n);
    ptr::set_alpha(mt, m.alpha);
    ptr::set_eta(mt, m.eta);
    mats.push_back(mt);
  }
  for (int i = 0; i < d->num_objects; i++) {
    auto& o  = d->objects[i];
    auto  ob = ptr::add_object(sc);
    ptr::set_frame(ob, fr(o.frame));
    ptr::set_shape(ob, shapes[o.shape]);
    ptr::set_material(ob, mats[o.material]);
  }
  for (int i = 0; i < d->num_environments; i++) {
    auto& e  = d->environments[i];
    auto  en = ptr::add_environment(sc);
    ptr::set_frame(en, fr(e.frame));
    ptr::set_emission(en, v3(e.emission), tx(e.emission_tex));
  }
  auto params = ptr::trace_params{};
  ptr::init_bvh(sc, params);
  ptr::init_lights(sc, params, {});
  return rs;
}
void ref_scene_destroy(void* h) {
  auto rs = (ref_scene*)h;
  delete rs->scene;
  delete rs;
}
// shape->bvh->primitives: the leaf order of the reference's own tree over this shape (build_bvh, :601-650, :746-752)
int ref_shape_order(void* h, int shape, int* order) {
  auto rs = (ref_scene*)h;
  if (shape < 0 || shape >= (int)rs->scene->shapes.size()) return -1;
  auto& prims = rs->scene->shapes[shape]->bvh->primitives;
  if (order) memcpy(order, prims.data(), sizeof(int) * prims.size());
  return (int)prims.size();
}
int64_t ref_scene_bvh_nodes(void* h, int shape) {
  auto rs = (ref_scene*)h;
  if (shape < 0) return (int64_t)rs->scene->bvh->nodes.size();
  return (int64_t)rs->scene->shapes[shape]->bvh->nodes.size();
}

// intersect_scene_bvh(scene, ray, find_any, true), yocto_pathtrace.cpp:1039
void ref_intersect_batch(void* h, const yh_ray* rays, int64_t n, int find_any,
    yh_hit* hits, int nthreads) {
  auto rs = (ref_scene*)h;
  auto work = [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; i++) {
      auto& r   = rays[i];
      auto  ray = ym::ray3f{v3(r.o), v3(r.d), r.tmin, r.tmax};
      auto  is  = ptr::intersect_scene_bvh(rs->scene, ray, find_any != 0, true);
      hits[i]   = {is.object, is.element, {is.uv.x, is.uv.y}, is.distance,
          is.hit ? 1 : 0};
    }
  };
  if (nthreads <= 1) return work(0, n);
  auto th = std::vector<std::thread>{};
  for (int t = 0; t < nthreads; t++)
    th.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
  for (auto& t : th) t.join();
}

// ---- render state (init_state :1931, trace_samples :1992) -------------------
void* ref_state_create(
    void* h, const yh_camera_desc* c, const yh_trace_params* p) {
  auto rs = (ref_scene*)h;
  auto st = new ref_state{};
  st->rs  = rs;
  auto cam      = ptr::add_camera(rs->scene);
  cam->frame    = fr(c->frame);
  cam->lens     = c->lens;
  cam->film     = {c->film[0], c->film[1]};
  cam->focus    = c->focus;
  cam->aperture = c->aperture;
  st->camera    = cam;
  st->params.resolution = p->resolution;
  st->params.shader     = (ptr::shader_type)p->shader;
  st->params.samples    = p->samples;
  st->params.bounces    = p->bounces;
  st->params.clamp      = p->clamp;
  st->params.seed       = p->seed;
  st->params.noparallel = p->noparallel != 0;
  st->params.pratio     = p->pratio;
  st->state             = new ptr::state{};
  ptr::init_state(st->state, rs->scene, cam, st->params);
  return st;
}
void ref_state_destroy(void* h) {
  auto st = (ref_state*)h;
  delete st->state;
  delete st;
}
void ref_state_size(void* h, int* w, int* hgt) {
  auto st = (ref_state*)h;
  *w      = st->state->render.size().x;
  *hgt    = st->state->render.size().y;
}
// returns seconds spent inside the trace_samples loop
double ref_trace_samples(void* h, int nsamples) {
  auto st = (ref_state*)h;
  auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < nsamples; s++)
    ptr::trace_samples(st->state, st->rs->scene, st->camera, st->params);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}
// one trace_sample call on one pixel (yocto_pathtrace.cpp:1676)
void ref_trace_pixel(void* h, int i, int j, float* rgba) {
  auto st = (ref_state*)h;
  auto v  = ptr::trace_sample(
      st->state, st->rs->scene, st->camera, {i, j}, st->params);
  memcpy(rgba, &v, 16);
}
void ref_state_download(void* h, float* render, float* accum, int* samples,
    uint64_t* rng) {
  auto st = (ref_state*)h;
  auto n  = st->state->render.count();
  if (render) memcpy(render, st->state->render.data(), 16 * n);
  for (size_t i = 0; i < n; i++) {
    auto& px = st->state->pixels.data()[i];
    if (accum) memcpy(accum + 4 * i, &px.accumulated, 16);
    if (samples) samples[i] = px.samples;
    if (rng) {
      rng[2 * i]     = px.rng.state;
      rng[2 * i + 1] = px.rng.inc;
    }
  }
}

}  // extern "C"

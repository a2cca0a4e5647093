// yh_loadcheck scene.json [--fast-only]
// Loads a scene with the reference's sceneio::load_scene and with pathtrace_b200::load_scene
// (bulk PLY reader + restated JSON schema), compares EVERY field of the two models bit for
// bit and prints one JSON line with the two load times.  Test + timing tool (tests/test_dropin_cpu.py).
#include <yocto/yocto_sceneio.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>

#include "yh_sceneio.h"

namespace sio = yocto::sceneio;
namespace gpu = yocto::pathtrace_b200;

template <typename T>
static bool same_bits(const std::vector<T>& a, const std::vector<T>& b) {
  return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(T)) == 0);
}
template <typename T>
static bool same_image(const yocto::image::image<T>& a, const yocto::image::image<T>& b) {
  return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.count() * sizeof(T)) == 0);
}
template <typename T>
static bool same_pod(const T& a, const T& b) { return memcmp(&a, &b, sizeof(T)) == 0; }

static int differences = 0;
#define CHECK(cond, what)                                         \
  do {                                                            \
    if (!(cond)) {                                                \
      differences++;                                              \
      if (differences <= 20) fprintf(stderr, "differs: %s\n", std::string(what).c_str()); \
    }                                                             \
  } while (0)

template <typename T>
static int index_of(const std::vector<T*>& v, const T* p) {
  for (size_t i = 0; i < v.size(); i++)
    if (v[i] == p) return (int)i;
  return -1;
}

int main(int argc, const char** argv) {
  if (argc < 2) return fprintf(stderr, "usage: yh_loadcheck scene.json [--fast-only]\n"), 2;
  bool fast_only = argc > 2 && std::string(argv[2]) == "--fast-only";
  auto clock = [] { return std::chrono::steady_clock::now(); };
  auto secs  = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
  std::string error;
  auto        t0 = clock();
  auto        b  = std::make_unique<sio::model>();
  if (!gpu::load_scene(argv[1], b.get(), error)) return fprintf(stderr, "fast loader: %s\n", error.c_str()), 1;
  auto t1 = clock();
  size_t segments = 0, vertices = 0;
  for (auto s : b->shapes) segments += s->lines.size(), vertices += s->positions.size();
  if (fast_only) {
    printf("{\"fast_s\": %.4f, \"segments\": %zu, \"vertices\": %zu}\n", secs(t0, t1), segments, vertices);
    return 0;
  }
  auto a = std::make_unique<sio::model>();
  if (!sio::load_scene(argv[1], a.get(), error)) return fprintf(stderr, "reference loader: %s\n", error.c_str()), 1;
  auto t2 = clock();

  CHECK(a->name == b->name && a->copyright == b->copyright, "scene name");
  CHECK(a->cameras.size() == b->cameras.size() && a->shapes.size() == b->shapes.size() && a->textures.size() == b->textures.size() &&
            a->materials.size() == b->materials.size() && a->objects.size() == b->objects.size() &&
            a->environments.size() == b->environments.size() && a->instances.size() == b->instances.size() &&
            a->subdivs.size() == b->subdivs.size(),
      "element counts");
  if (differences == 0) {
    for (size_t i = 0; i < a->cameras.size(); i++) {
      auto x = a->cameras[i], y = b->cameras[i];
      CHECK(x->name == y->name && same_pod(x->frame, y->frame) && x->orthographic == y->orthographic && same_pod(x->lens, y->lens) &&
                same_pod(x->film, y->film) && same_pod(x->aspect, y->aspect) && same_pod(x->focus, y->focus) &&
                same_pod(x->aperture, y->aperture),
          "camera " + x->name);
    }
    for (size_t i = 0; i < a->textures.size(); i++) {
      auto x = a->textures[i], y = b->textures[i];
      CHECK(x->name == y->name && same_image(x->colorf, y->colorf) && same_image(x->colorb, y->colorb) &&
                same_image(x->scalarf, y->scalarf) && same_image(x->scalarb, y->scalarb),
          "texture " + x->name);
    }
    for (size_t i = 0; i < a->shapes.size(); i++) {
      auto x = a->shapes[i], y = b->shapes[i];
      CHECK(x->name == y->name && same_bits(x->points, y->points) && same_bits(x->lines, y->lines) &&
                same_bits(x->triangles, y->triangles) && same_bits(x->quads, y->quads) && same_bits(x->positions, y->positions) &&
                same_bits(x->normals, y->normals) && same_bits(x->texcoords, y->texcoords) && same_bits(x->colors, y->colors) &&
                same_bits(x->radius, y->radius) && same_bits(x->tangents, y->tangents),
          "shape " + x->name);
    }
    for (size_t i = 0; i < a->instances.size(); i++)
      CHECK(a->instances[i]->name == b->instances[i]->name && same_bits(a->instances[i]->frames, b->instances[i]->frames), "instance");
    for (size_t i = 0; i < a->subdivs.size(); i++) {
      auto x = a->subdivs[i], y = b->subdivs[i];
      CHECK(x->name == y->name && same_bits(x->quadspos, y->quadspos) && same_bits(x->quadsnorm, y->quadsnorm) &&
                same_bits(x->quadstexcoord, y->quadstexcoord) && same_bits(x->positions, y->positions) &&
                same_bits(x->normals, y->normals) && same_bits(x->texcoords, y->texcoords),
          "subdiv " + x->name);
    }
    for (size_t i = 0; i < a->materials.size(); i++) {
      auto x = a->materials[i], y = b->materials[i];
      bool scalars = x->name == y->name && same_pod(x->emission, y->emission) && same_pod(x->color, y->color) &&
                     same_pod(x->specular, y->specular) && same_pod(x->roughness, y->roughness) && same_pod(x->metallic, y->metallic) &&
                     same_pod(x->ior, y->ior) && same_pod(x->spectint, y->spectint) && same_pod(x->coat, y->coat) &&
                     same_pod(x->transmission, y->transmission) && same_pod(x->translucency, y->translucency) &&
                     same_pod(x->scattering, y->scattering) && same_pod(x->scanisotropy, y->scanisotropy) &&
                     same_pod(x->trdepth, y->trdepth) && same_pod(x->opacity, y->opacity) && same_pod(x->displacement, y->displacement) &&
                     x->thin == y->thin && same_pod(x->eumelanin, y->eumelanin) && same_pod(x->pheomelanin, y->pheomelanin) &&
                     same_pod(x->sigma_a, y->sigma_a) && same_pod(x->beta_m, y->beta_m) && same_pod(x->beta_n, y->beta_n) &&
                     same_pod(x->alpha, y->alpha) && same_pod(x->eta, y->eta) && x->subdivisions == y->subdivisions && x->smooth == y->smooth;
      auto tx = [&](sio::texture* p, sio::texture* q) { return index_of(a->textures, p) == index_of(b->textures, q); };
      bool texs = tx(x->emission_tex, y->emission_tex) && tx(x->color_tex, y->color_tex) && tx(x->specular_tex, y->specular_tex) &&
                  tx(x->metallic_tex, y->metallic_tex) && tx(x->roughness_tex, y->roughness_tex) &&
                  tx(x->transmission_tex, y->transmission_tex) && tx(x->translucency_tex, y->translucency_tex) &&
                  tx(x->spectint_tex, y->spectint_tex) && tx(x->scattering_tex, y->scattering_tex) && tx(x->coat_tex, y->coat_tex) &&
                  tx(x->opacity_tex, y->opacity_tex) && tx(x->normal_tex, y->normal_tex) && tx(x->displacement_tex, y->displacement_tex);
      CHECK(scalars && texs, "material " + x->name);
    }
    for (size_t i = 0; i < a->objects.size(); i++) {
      auto x = a->objects[i], y = b->objects[i];
      CHECK(x->name == y->name && same_pod(x->frame, y->frame) && index_of(a->shapes, x->shape) == index_of(b->shapes, y->shape) &&
                index_of(a->materials, x->material) == index_of(b->materials, y->material) &&
                index_of(a->instances, x->instance) == index_of(b->instances, y->instance) &&
                index_of(a->subdivs, x->subdiv) == index_of(b->subdivs, y->subdiv),
          "object " + x->name);
    }
    for (size_t i = 0; i < a->environments.size(); i++) {
      auto x = a->environments[i], y = b->environments[i];
      CHECK(x->name == y->name && same_pod(x->frame, y->frame) && same_pod(x->emission, y->emission) &&
                index_of(a->textures, x->emission_tex) == index_of(b->textures, y->emission_tex),
          "environment " + x->name);
    }
  }
  printf("{\"equal\": %s, \"differences\": %d, \"fast_s\": %.4f, \"reference_s\": %.4f, \"segments\": %zu, \"vertices\": %zu}\n",
      differences == 0 ? "true" : "false", differences, secs(t0, t1), secs(t1, t2), segments, vertices);
  return differences == 0 ? 0 : 1;
}

// yscenetrace on the B200 path: the reference's offline renderer command line
// (apps/yscenetrace/yscenetrace.cpp:199-278: same flags, same JSON + PLY loading
// through the reference's own yocto_sceneio, same image writers) with the
// per-sample hot path running in the CUDA library behind include/yocto_hair_b200.h.
//
//   yscenetrace_b200 [--loader fast|reference] [--spp-batch N] [--seed S] [--device D] [--gpus N] [--shard tiles|samples]
//                    [--camera NAME] [-r RES] [-s SAMPLES] [-t path] [-b BOUNCES]
//                    [--clamp C] [--save-batch] [-o out.pfm] scene.json
//
// The product binary has ONE backend, the CUDA library: there is no CPU path in it.  The same source compiled
// with -DYH_AB_TOOL gives host/_build/yscenetrace_ab, a TEST tool whose extra `--backend cpu` runs the reference's
// own functions on the same scene for the A/B comparisons of tests/test_dropin_gpu.py.
#include <yocto/yocto_commonio.h>
#include <yocto/yocto_image.h>
#include <yocto/yocto_sceneio.h>
#include <yocto/yocto_shape.h>
#include <yocto_pathtrace/yocto_pathtrace.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <future>
#include <thread>
#include <map>
#include <memory>

#include "yh_sceneio.h"
#include "yocto_pathtrace_b200.h"

using namespace std::string_literals;
namespace cli = yocto::commonio;
namespace ptr = yocto::pathtrace;
namespace gpu = yocto::pathtrace_b200;
namespace sio = yocto::sceneio;
namespace shp = yocto::shape;

// sceneio::model -> pathtrace::scene through the reference's public setters
// (the role of init_scene in apps/yscenetrace/yscenetrace.cpp:49-197).
static ptr::camera* build_trace_scene(ptr::scene* scene, const sio::model* io, const sio::camera* wanted) {
  std::map<const void*, ptr::texture*>  textures{{nullptr, nullptr}};
  std::map<const void*, ptr::material*> materials{{nullptr, nullptr}};
  std::map<const void*, ptr::shape*>    shapes{{nullptr, nullptr}};
  ptr::camera* chosen = nullptr;
  for (auto c : io->cameras) {
    auto camera = ptr::add_camera(scene);
    ptr::set_frame(camera, c->frame);
    ptr::set_lens(camera, c->lens, c->aspect, c->film);
    ptr::set_focus(camera, c->aperture, c->focus);
    if (c == wanted) chosen = camera;
  }
  for (auto t : io->textures) {
    auto texture = textures[t] = ptr::add_texture(scene);
    if (!t->colorf.empty()) ptr::set_texture(texture, t->colorf);
    else if (!t->colorb.empty()) ptr::set_texture(texture, t->colorb);
    else if (!t->scalarf.empty()) ptr::set_texture(texture, t->scalarf);
    else if (!t->scalarb.empty()) ptr::set_texture(texture, t->scalarb);
  }
  for (auto m : io->materials) {
    auto material = materials[m] = ptr::add_material(scene);
    ptr::set_emission(material, m->emission, textures.at(m->emission_tex));
    ptr::set_color(material, m->color, textures.at(m->color_tex));
    ptr::set_specular(material, m->specular, textures.at(m->specular_tex));
    ptr::set_metallic(material, m->metallic, textures.at(m->metallic_tex));
    ptr::set_roughness(material, m->roughness, textures.at(m->roughness_tex));
    ptr::set_transmission(material, m->transmission, m->thin, m->trdepth, textures.at(m->transmission_tex));
    ptr::set_scattering(material, m->scattering, m->scanisotropy, textures.at(m->scattering_tex));
    ptr::set_opacity(material, m->opacity, textures.at(m->opacity_tex));
    ptr::set_normalmap(material, textures.at(m->normal_tex));
    ptr::set_ior(material, m->ior);
    ptr::set_thin(material, m->thin);
    ptr::set_sigma_a(material, m->sigma_a);
    ptr::set_eumelanin(material, m->eumelanin);
    ptr::set_pheomelanin(material, m->pheomelanin);
    ptr::set_beta_m(material, m->beta_m);
    ptr::set_beta_n(material, m->beta_n);
    ptr::set_alpha(material, m->alpha);
    ptr::set_eta(material, m->eta);
  }
  for (auto s : io->shapes) {
    auto shape = shapes[s] = ptr::add_shape(scene);
    ptr::set_positions(shape, s->positions);
    ptr::set_normals(shape, s->normals);
    ptr::set_texcoords(shape, s->texcoords);
    ptr::set_radius(shape, s->radius);
    ptr::set_points(shape, s->points);
    ptr::set_lines(shape, s->lines);
    ptr::set_triangles(shape, s->quads.empty() ? s->triangles : shp::quads_to_triangles(s->quads));
  }
  // subdivision surfaces: control cages go in as shapes with the subdiv fields set; the reference's own
  // init_subdivs() tesselates and displaces them ON THE HOST before either backend builds its BVH
  // (apps/yscenetrace/yscenetrace.cpp:120-130, :155-163, :243)
  std::map<const void*, ptr::shape*> subdivs{{nullptr, nullptr}};
  for (auto sd : io->subdivs) {
    auto shape = subdivs[sd] = ptr::add_shape(scene);
    ptr::set_subdiv_quadspos(shape, sd->quadspos);
    ptr::set_subdiv_quadstexcoord(shape, sd->quadstexcoord);
    ptr::set_subdiv_positions(shape, sd->positions);
    ptr::set_subdiv_texcoords(shape, sd->texcoords);
  }
  for (auto o : io->objects) {
    auto frames = o->instance ? o->instance->frames : std::vector<yocto::math::frame3f>{yocto::math::identity3x4f};
    for (auto& f : frames) {
      auto object = ptr::add_object(scene);
      ptr::set_frame(object, o->instance ? f * o->frame : o->frame);
      if (o->shape) ptr::set_shape(object, shapes.at(o->shape));
      if (o->subdiv) {
        ptr::set_shape(object, subdivs.at(o->subdiv));
        ptr::set_subdiv_subdivision(subdivs.at(o->subdiv), o->material->subdivisions, true);
        ptr::set_subdiv_displacement(subdivs.at(o->subdiv), o->material->displacement, textures.at(o->material->displacement_tex));
      }
      ptr::set_material(object, materials.at(o->material));
    }
  }
  for (auto e : io->environments) {
    auto environment = ptr::add_environment(scene);
    ptr::set_frame(environment, e->frame);
    ptr::set_emission(environment, e->emission, textures.at(e->emission_tex));
  }
  return chosen;
}

// ---- the interactive consumer, headless -----------------------------------------
// What apps/ysceneitraces/ysceneitraces.cpp:255-302 (reset_display) does to the tracer, without the window: a
// preview pass at resolution / pratio, a worker thread that adds one sample per pass and polls a stop flag, a
// camera edit that stops the worker and resets the state, then a render to completion.  Both backends run the
// same sequence; the files it writes (<dir>/preview.pfm, edited.pfm) must agree between them, and on the device
// the interrupted state must equal a fresh render of exactly the passes that completed.
template <typename Backend>
static int interactive_session(const Backend& be, ptr::scene* scene, ptr::camera* camera, const ptr::trace_params& params,
    const std::string& dir, int stop_after, bool check_interrupted) {
  auto save = [&](const std::string& name, const yocto::image::image<yocto::math::vec4f>& img) {
    auto error = ""s;
    if (!save_image(dir + "/" + name, img, error)) cli::print_fatal(error);
  };
  // reset_display: preview (:268-283)
  auto pstate = std::make_unique<ptr::state>();
  auto pprms  = params;
  pprms.resolution /= params.pratio;
  pprms.samples = 1;
  be.init_state(pstate.get(), scene, camera, pprms);
  be.trace_samples(pstate.get(), scene, camera, pprms, nullptr);
  save("preview.pfm", pstate->render);
  // reset_display: the render worker (:286-300), stopped from this thread after `stop_after` passes
  auto state = std::make_unique<ptr::state>();
  be.init_state(state.get(), scene, camera, params);
  std::atomic<bool> stop{false};
  std::atomic<int>  passes{0};
  auto worker = std::async(std::launch::async, [&]() {
    for (auto sample = 0; sample < params.samples; sample++) {
      if (stop) return;
      be.trace_samples(state.get(), scene, camera, params, &stop);
      if (stop) return;  // a pass that saw the flag did not count (the reference abandons it mid-image)
      passes++;
    }
  });
  while (passes < stop_after && passes < params.samples) std::this_thread::sleep_for(std::chrono::microseconds(200));
  stop = true;
  worker.get();
  int completed = passes;
  printf("interactive: stopped after %d of %d passes\n", completed, params.samples);
  if (check_interrupted) {  // device passes are atomic: the interrupted image is a clean k-sample render, where k
                            // counts the pass that was in flight when the flag rose (it ran to its end)
    completed = be.samples_done(state.get());
    auto fresh = std::make_unique<ptr::state>();
    be.init_state(fresh.get(), scene, camera, params);
    for (int s = 0; s < completed; s++) be.trace_samples(fresh.get(), scene, camera, params, nullptr);
    bool same = completed == 0 || memcmp(fresh->render.data(), state->render.data(), sizeof(yocto::math::vec4f) * fresh->render.count()) == 0;
    printf("interactive: interrupted state %s a fresh %d-sample render\n", same ? "equals" : "DIFFERS FROM", completed);
    if (!same) return 3;
  }
  // camera edit -> reset_display again -> render to completion
  camera->frame.o = camera->frame.o + camera->frame.x * 0.35f + camera->frame.y * 0.1f;
  be.init_state(state.get(), scene, camera, params);
  stop = false;
  for (auto sample = 0; sample < params.samples; sample++) be.trace_samples(state.get(), scene, camera, params, &stop);
  save("edited.pfm", state->render);
  return 0;
}
struct device_backend {
  int samples_done(ptr::state* st) const {
    gpu::sync_pixels(st);
    return st->pixels.empty() ? 0 : st->pixels[{0, 0}].samples;
  }
  void init_state(ptr::state* st, const ptr::scene* sc, const ptr::camera* cam, const ptr::trace_params& p) const {
    gpu::init_state(st, sc, cam, p);
  }
  void trace_samples(ptr::state* st, const ptr::scene* sc, const ptr::camera* cam, const ptr::trace_params& p,
      std::atomic<bool>* stop) const {
    if (stop) gpu::trace_samples(st, sc, cam, p, stop);
    else gpu::trace_samples(st, sc, cam, p);
  }
};
#ifdef YH_AB_TOOL
struct reference_backend {
  int samples_done(ptr::state* st) const { return st->pixels.empty() ? 0 : st->pixels[{0, 0}].samples; }
  void init_state(ptr::state* st, const ptr::scene* sc, const ptr::camera* cam, const ptr::trace_params& p) const {
    ptr::init_state(st, sc, cam, p);
  }
  void trace_samples(ptr::state* st, const ptr::scene* sc, const ptr::camera* cam, const ptr::trace_params& p,
      std::atomic<bool>* stop) const {
    if (stop) ptr::trace_samples(st, sc, cam, p, stop);
    else ptr::trace_samples(st, sc, cam, p);
  }
};
#endif

int main(int argc, const char* argv[]) {
  auto params      = ptr::trace_params{};
  auto save_batch  = false;
  auto camera_name = ""s;
  auto imfilename  = "out.hdr"s;
  auto filename    = "scene.json"s;
  auto backend     = "cuda"s;
  auto spp_batch   = 16;
  auto device      = 0;
  auto seed        = (int)ptr::default_seed;
  auto loader      = "fast"s;
  auto session_dir = ""s;
  auto stop_after  = 2;
  auto gpus        = 1;
  auto shard       = "tiles"s;

  auto cli = cli::make_cli("yscenetrace_b200", "Offline hair path tracing on B200");
#ifdef YH_AB_TOOL
  add_option(cli, "--backend", backend, "cuda (B200 library) or cpu (the reference's own code; test tool only).");
#endif
  add_option(cli, "--spp-batch", spp_batch, "Samples per pixel per device call.");
  add_option(cli, "--loader", loader, "fast (bulk PLY reader, yh_sceneio) or reference (yocto_sceneio).");
  add_option(cli, "--device", device, "CUDA device (the first one with --gpus).");
  add_option(cli, "--gpus", gpus, "GPUs of this node driven by this process (scene replicated, image sharded, summed over NVLink).");
  add_option(cli, "--shard", shard, "With --gpus: tiles (reference RNG streams, bit-identical image) or samples (counter-mode streams).");
  add_option(cli, "--interactive-check", session_dir, "Run the headless interactive session and write its images here.");
  add_option(cli, "--stop-after", stop_after, "Interactive session: passes before the simulated camera edit.");
  add_option(cli, "--seed", seed, "Render seed (trace_params::seed).");
  add_option(cli, "--camera", camera_name, "Camera name.");
  add_option(cli, "--resolution,-r", params.resolution, "Image resolution.");
  add_option(cli, "--samples,-s", params.samples, "Number of samples.");
  add_option(cli, "--shader,-t", params.shader, "Shader type.", ptr::shader_names);
  add_option(cli, "--bounces,-b", params.bounces, "Maximum number of bounces.");
  add_option(cli, "--clamp", params.clamp, "Final pixel clamping.");
  add_option(cli, "--save-batch", save_batch, "Save images progressively");
  add_option(cli, "--output-image,-o", imfilename, "Image filename");
  add_option(cli, "scene", filename, "Scene filename", true);
  parse_cli(cli, argc, argv);
  params.seed = (uint64_t)(unsigned)seed;
  if (backend != "cuda" && backend != "cpu") cli::print_fatal("unknown backend " + backend);
  const bool on_gpu = backend == "cuda";

  auto ioscene = std::make_unique<sio::model>();
  auto ioerror = ""s;
  if (loader != "fast" && loader != "reference") cli::print_fatal("unknown loader " + loader);
  auto tl0    = std::chrono::steady_clock::now();
  auto loaded = loader == "fast" ? gpu::load_scene(filename, ioscene.get(), ioerror, cli::print_progress)
                                 : sio::load_scene(filename, ioscene.get(), ioerror, cli::print_progress);
  if (!loaded) cli::print_fatal(ioerror);
  printf("loader %s: %.3f s\n", loader.c_str(), std::chrono::duration<double>(std::chrono::steady_clock::now() - tl0).count());
  auto scene  = std::make_unique<ptr::scene>();
  auto camera = build_trace_scene(scene.get(), ioscene.get(), get_camera(ioscene.get(), camera_name));
  ioscene.reset();

  auto clock = [] { return std::chrono::steady_clock::now(); };
  auto secs  = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
  auto state = std::make_unique<ptr::state>();
  auto t0    = clock();
  try {
    ptr::init_subdivs(scene.get(), params, cli::print_progress);  // host tesselation, the reference's own (both backends)
    if (on_gpu) {
      if (shard != "tiles" && shard != "samples") cli::print_fatal("unknown shard mode " + shard);
      auto opts   = gpu::device_options{};
      opts.device = device;
      if (gpus > 1) {
        for (int k = 0; k < gpus; k++) opts.devices.push_back(device + k);
        opts.sample_shard = shard == "samples";
      }
      gpu::set_device_options(opts);
      gpu::init_bvh(scene.get(), params, cli::print_progress);
      gpu::init_lights(scene.get(), params, cli::print_progress);
      gpu::init_state(state.get(), scene.get(), camera, params);
    } else {
#ifdef YH_AB_TOOL
      ptr::init_bvh(scene.get(), params, cli::print_progress);
      ptr::init_lights(scene.get(), params, cli::print_progress);
      ptr::init_state(state.get(), scene.get(), camera, params);
#endif
    }
    if (!session_dir.empty()) {
      int rc = 0;
      if (on_gpu) rc = interactive_session(device_backend{}, scene.get(), camera, params, session_dir, stop_after, true);
#ifdef YH_AB_TOOL
      else rc = interactive_session(reference_backend{}, scene.get(), camera, params, session_dir, stop_after, false);
#endif
      if (on_gpu) gpu::release(scene.get());
      return rc;
    }
    auto t1 = clock();
    if (save_batch || !on_gpu) spp_batch = 1;
    for (auto sample = 0; sample < params.samples;) {
      cli::print_progress("render image", sample, params.samples);
      auto n = std::min(spp_batch, params.samples - sample);
      if (on_gpu) gpu::trace_samples(state.get(), scene.get(), camera, params, n);
#ifdef YH_AB_TOOL
      else ptr::trace_samples(state.get(), scene.get(), camera, params);
#endif
      sample += n;
      if (save_batch) {
        auto dot  = imfilename.rfind('.');  // <name>-s<k>.<ext>, as the reference names its batches
        auto stem = dot == std::string::npos ? imfilename : imfilename.substr(0, dot);
        auto ext  = dot == std::string::npos ? ""s : imfilename.substr(dot);
        if (!save_image(stem + "-s" + std::to_string(sample - 1) + ext, state->render, ioerror)) cli::print_fatal(ioerror);
      }
    }
    cli::print_progress("render image", params.samples, params.samples);
    auto   t2      = clock();
    double samples = (double)state->render.count() * params.samples;
    printf("backend %s: setup %.3f s, render %.3f s, %.3f Msamples/s\n", backend.c_str(), secs(t0, t1), secs(t1, t2),
        samples / secs(t1, t2) / 1e6);
  } catch (const std::exception& e) {
    cli::print_fatal(e.what());
  }
  if (!save_image(imfilename, state->render, ioerror)) cli::print_fatal(ioerror);
  if (on_gpu) gpu::release(scene.get());
  return 0;
}

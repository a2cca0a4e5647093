// Host shim over the C ABI; see yocto_pathtrace_b200.h.
#include "yocto_pathtrace_b200.h"

#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "../../include/yocto_hair_b200.h"

namespace yocto::pathtrace_b200 {

namespace {

// device copies of one ptr::scene: a scene on one device, or a group of replicas (device_options::devices)
struct scene_record {
  yh_scene* scene = nullptr;
  yh_group* group = nullptr;
};
struct state_record {
  yh_state*          handle = nullptr;  // one device ...
  yh_group*          group  = nullptr;  // ... or the group the states live in
  const ptr::scene*  scene  = nullptr;
  ptr::camera        camera{};
};
std::mutex                                            table_mutex;
std::unordered_map<const ptr::scene*, scene_record>   scene_table;
std::unordered_map<const ptr::state*, state_record>   state_table;
device_options                                        options{};

void check(yh_status status, const char* what) {
  if (status == YH_OK) return;
  const char* msg = yh_last_error();
  throw std::runtime_error(std::string(what) + ": " + (msg ? msg : "unknown error"));
}
void put_frame(float* out, const math::frame3f& f) { memcpy(out, &f, 48); }

scene_record find_scene_record(const ptr::scene* scene) {
  std::lock_guard<std::mutex> lock(table_mutex);
  auto it = scene_table.find(scene);
  if (it == scene_table.end()) throw std::runtime_error("yocto_pathtrace_b200: call init_bvh(scene) first");
  return it->second;
}
// the scene handle of one device (a group: its first replica, made current)
yh_scene* find_scene(const ptr::scene* scene) {
  scene_record rec = find_scene_record(scene);
  if (!rec.group) return rec.scene;
  const yh_scene* first = nullptr;
  int             device = 0;
  check(yh_group_scene(rec.group, 0, &first, &device), "yh_group_scene");
  check(yh_set_device(device), "yh_set_device");
  return const_cast<yh_scene*>(first);
}
state_record& find_state(const ptr::state* state) {
  std::lock_guard<std::mutex> lock(table_mutex);
  auto it = state_table.find(state);
  if (it == state_table.end()) throw std::runtime_error("yocto_pathtrace_b200: call init_state(state) first");
  return it->second;
}

// ptr::scene -> yh_scene_desc.  Vertex arrays are passed in place (vec3f is 3
// packed floats); textures are expanded to linear float RGB with the
// reference's own conversion (lookup_texture, yocto_pathtrace.cpp:148-164).
scene_record upload(const ptr::scene* scene, const yh_bvh_params& bvh) {
  auto index_of = [](const auto& items, const auto* item) -> int {
    if (!item) return -1;
    for (size_t i = 0; i < items.size(); i++)
      if (items[i] == item) return (int)i;
    throw std::runtime_error("yocto_pathtrace_b200: scene element is not owned by the scene");
  };
  std::vector<yh_shape_desc> shapes(scene->shapes.size());
  for (size_t i = 0; i < shapes.size(); i++) {
    auto  s = scene->shapes[i];
    auto& d = shapes[i];
    d = yh_shape_desc{};
    d.positions = (const float*)s->positions.data(), d.num_positions = (int)s->positions.size();
    d.normals = (const float*)s->normals.data(), d.num_normals = (int)s->normals.size();
    d.texcoords = (const float*)s->texcoords.data(), d.num_texcoords = (int)s->texcoords.size();
    d.radius = s->radius.data(), d.num_radius = (int)s->radius.size();
    // the reference dispatches triangles before lines for shading but lines before triangles
    // for intersection; a shape with both is not something its loaders produce
    d.lines = (const int*)s->lines.data(), d.num_lines = (int)s->lines.size();
    d.triangles = (const int*)s->triangles.data(), d.num_triangles = (int)s->triangles.size();
    d.points = s->points.data(), d.num_points = (int)s->points.size();
  }
  std::vector<std::vector<float>> texels(scene->textures.size());
  std::vector<yh_texture_desc>    textures(scene->textures.size());
  for (size_t i = 0; i < textures.size(); i++) {
    auto t = scene->textures[i];
    auto size = !t->colorf.empty() ? t->colorf.size() : !t->colorb.empty() ? t->colorb.size()
              : !t->scalarf.empty() ? t->scalarf.size() : t->scalarb.size();
    auto& px = texels[i];
    px.resize(3 * (size_t)size.x * size.y);
    for (int j = 0; j < size.y; j++)
      for (int k = 0; k < size.x; k++) {
        auto c = math::vec3f{1, 1, 1};
        if (!t->colorf.empty()) c = t->colorf[{k, j}];
        else if (!t->colorb.empty()) c = math::srgb_to_rgb(math::byte_to_float(t->colorb[{k, j}]));
        else if (!t->scalarf.empty()) c = math::vec3f{t->scalarf[{k, j}]};
        else if (!t->scalarb.empty()) c = math::srgb_to_rgb(math::byte_to_float(math::vec3b{t->scalarb[{k, j}]}));
        memcpy(&px[3 * ((size_t)j * size.x + k)], &c, 12);
      }
    textures[i] = {px.data(), size.x, size.y};
  }
  // Slots the reference evaluates with ldr_as_linear = true (specular, metallic, roughness, opacity and the
  // transmission factor, yocto_pathtrace.cpp:413-424) need the BYTE textures expanded without the sRGB curve:
  // such a texture gets a second, linear copy the first time one of those slots refers to it.
  std::map<int, int> linear_copy;
  auto linear_of = [&](int index) -> int {
    if (index < 0) return -1;
    auto t = scene->textures[index];
    if (t->colorb.empty() && t->scalarb.empty()) return index;  // float textures have one reading
    if (!t->colorf.empty()) return index;                      // lookup_texture prefers colorf (:149-150)
    auto it = linear_copy.find(index);
    if (it != linear_copy.end()) return it->second;
    auto size = !t->colorb.empty() ? t->colorb.size() : t->scalarb.size();
    texels.emplace_back(3 * (size_t)size.x * size.y);
    auto& px = texels.back();
    for (int j = 0; j < size.y; j++)
      for (int k = 0; k < size.x; k++) {
        auto c = !t->colorb.empty() ? math::byte_to_float(t->colorb[{k, j}]) : math::byte_to_float(math::vec3b{t->scalarb[{k, j}]});
        memcpy(&px[3 * ((size_t)j * size.x + k)], &c, 12);
      }
    textures.push_back({px.data(), size.x, size.y});
    return linear_copy[index] = (int)textures.size() - 1;
  };
  std::vector<yh_material_desc> materials(scene->materials.size());
  for (size_t i = 0; i < materials.size(); i++) {
    auto  m = scene->materials[i];
    auto& d = materials[i];
    check(yh_default_material(&d), "yh_default_material");
    memcpy(d.emission, &m->emission, 12), memcpy(d.color, &m->color, 12), memcpy(d.spectint, &m->spectint, 12);
    memcpy(d.scattering, &m->scattering, 12), memcpy(d.sigma_a, &m->sigma_a, 12);
    d.specular = m->specular, d.roughness = m->roughness, d.metallic = m->metallic, d.ior = m->ior;
    d.transmission = m->transmission, d.scanisotropy = m->scanisotropy, d.trdepth = m->trdepth, d.opacity = m->opacity;
    d.thin = m->thin ? 1 : 0;
    d.eumelanin = m->eumelanin, d.pheomelanin = m->pheomelanin;
    d.beta_m = m->beta_m, d.beta_n = m->beta_n, d.alpha = m->alpha, d.eta = m->eta;
    d.emission_tex = index_of(scene->textures, m->emission_tex), d.color_tex = index_of(scene->textures, m->color_tex);
    d.specular_tex  = linear_of(index_of(scene->textures, m->specular_tex));
    d.metallic_tex  = linear_of(index_of(scene->textures, m->metallic_tex));
    d.roughness_tex = linear_of(index_of(scene->textures, m->roughness_tex));
    // the reference multiplies transmission by its EMISSION texture (:420-421), read as linear
    d.transmission_tex = linear_of(index_of(scene->textures, m->emission_tex));
    d.spectint_tex = -1;  // never evaluated by the reference's tracer
    d.scattering_tex = index_of(scene->textures, m->scattering_tex);
    d.opacity_tex = linear_of(index_of(scene->textures, m->opacity_tex)), d.normal_tex = linear_of(index_of(scene->textures, m->normal_tex));
  }
  std::vector<yh_object_desc> objects(scene->objects.size());
  for (size_t i = 0; i < objects.size(); i++) {
    put_frame(objects[i].frame, scene->objects[i]->frame);
    objects[i].shape    = index_of(scene->shapes, scene->objects[i]->shape);
    objects[i].material = index_of(scene->materials, scene->objects[i]->material);
  }
  std::vector<yh_environment_desc> environments(scene->environments.size());
  for (size_t i = 0; i < environments.size(); i++) {
    put_frame(environments[i].frame, scene->environments[i]->frame);
    memcpy(environments[i].emission, &scene->environments[i]->emission, 12);
    environments[i].emission_tex = index_of(scene->textures, scene->environments[i]->emission_tex);
  }
  yh_scene_desc desc{shapes.data(), textures.data(), materials.data(), objects.data(), environments.data(),
      (int)shapes.size(), (int)textures.size(), (int)materials.size(), (int)objects.size(), (int)environments.size()};
  scene_record rec{};
  if (options.devices.size() > 1) {  // replicas on several GPUs: scene + BVH + light tables in one call
    check(yh_group_create(&desc, &bvh, options.devices.data(), (int)options.devices.size(), &rec.group), "yh_group_create");
  } else {
    check(yh_scene_create(&desc, &rec.scene), "yh_scene_create");
  }
  return rec;
}

yh_trace_params to_params(const ptr::trace_params& p) {
  return {p.resolution, (int)p.shader, p.samples, p.bounces, p.clamp, p.seed, p.noparallel ? 1 : 0, p.pratio};
}

}  // namespace

void set_device_options(const device_options& o) { options = o; }

void init_bvh(ptr::scene* scene, const ptr::trace_params& params, ptr::progress_callback progress_cb) {
  if (progress_cb) progress_cb("build bvh", 0, 2);
  check(yh_set_device(options.devices.size() == 1 ? options.devices[0] : options.device), "yh_set_device");
  release(scene);
  yh_bvh_params bp{};
  bp.max_split = options.max_split;
  scene_record rec = upload(scene, bp);
  if (progress_cb) progress_cb("build bvh", 1, 2);
  if (rec.scene) {
    yh_status st = yh_scene_init_bvh(rec.scene, &bp);
    if (st != YH_OK) {
      std::string msg = yh_last_error();
      yh_scene_destroy(rec.scene);
      throw std::runtime_error("yh_scene_init_bvh: " + msg);
    }
  }
  {
    std::lock_guard<std::mutex> lock(table_mutex);
    scene_table[scene] = rec;
  }
  if (progress_cb) progress_cb("build bvh", 2, 2);
}

void init_lights(ptr::scene* scene, const ptr::trace_params& params, ptr::progress_callback progress_cb) {
  if (progress_cb) progress_cb("build light", 0, 1);
  scene_record rec = find_scene_record(scene);
  if (rec.scene) check(yh_scene_init_lights(rec.scene), "yh_scene_init_lights");  // a group built them in yh_group_create
  if (progress_cb) progress_cb("build light", 1, 1);
}

void init_state(ptr::state* state, const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params) {
  release(state);
  yh_camera_desc cam{};
  put_frame(cam.frame, camera->frame);
  cam.lens = camera->lens, cam.film[0] = camera->film.x, cam.film[1] = camera->film.y;
  cam.focus = camera->focus, cam.aperture = camera->aperture;
  yh_trace_params p = to_params(params);
  yh_state*       handle = nullptr;
  scene_record    rec = find_scene_record(scene);
  int w = 0, h = 0;
  if (rec.group) {
    check(yh_group_init_state(rec.group, &cam, &p, options.sample_shard ? YH_SHARD_SAMPLES : YH_SHARD_TILES), "yh_group_init_state");
    check(yh_group_size(rec.group, &w, &h), "yh_group_size");
  } else {
    check(yh_state_create(rec.scene, &cam, &p, &handle), "yh_state_create");
    if (options.world > 1) check(yh_state_set_shard(handle, options.rank, options.world), "yh_state_set_shard");
    check(yh_state_size(handle, &w, &h), "yh_state_size");
  }
  state->pixels.assign({w, h}, ptr::pixel{});
  state->render.assign({w, h}, math::zero4f);
  std::lock_guard<std::mutex> lock(table_mutex);
  state_table[state] = {handle, rec.group, scene, *camera};
}

void trace_samples(ptr::state* state, const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params,
    int num_samples, std::atomic<bool>* stop) {
  auto& rec = find_state(state);
  if (rec.scene != scene) throw std::runtime_error("yocto_pathtrace_b200: state was initialised for another scene");
  if (memcmp(&rec.camera, camera, sizeof(ptr::camera)) != 0)
    throw std::runtime_error("yocto_pathtrace_b200: the camera changed since init_state; call init_state again");
  if (rec.group) {
    // sample sharding deals whole samples to the GPUs: a pass is one sample per GPU
    int step = 1;
    if (options.sample_shard) check(yh_group_num_devices(rec.group, &step), "yh_group_num_devices");
    if (stop) {
      for (int s = 0; s < num_samples && !*stop; s += step) check(yh_group_trace_samples(rec.group, step, nullptr), "yh_group_trace_samples");
    } else {
      check(yh_group_trace_samples(rec.group, (num_samples + step - 1) / step * step, nullptr), "yh_group_trace_samples");
    }
    check(yh_group_download(rec.group, (float*)state->render.data(), nullptr, nullptr), "yh_group_download");
    return;
  }
  if (stop) {  // the reference polls `stop` per image row inside a 1-spp pass; here between 1-spp passes
    for (int s = 0; s < num_samples && !*stop; s++) check(yh_trace_samples(rec.handle, 1, nullptr), "yh_trace_samples");
  } else {
    check(yh_trace_samples(rec.handle, num_samples, nullptr), "yh_trace_samples");
  }
  check(yh_state_download(rec.handle, (float*)state->render.data(), nullptr, nullptr), "yh_state_download");
}
void trace_samples(ptr::state* state, const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params) {
  trace_samples(state, scene, camera, params, 1, nullptr);
}
void trace_samples(ptr::state* state, const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params,
    std::atomic<bool>* stop) {
  trace_samples(state, scene, camera, params, 1, stop);
}

img::image<vec4f> trace_image(const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params,
    ptr::progress_callback progress_cb, image_callback image_cb) {
  auto state = ptr::state{};
  pathtrace_b200::init_state(&state, scene, camera, params);
  const int batch = image_cb ? 1 : 16;  // a per-sample image callback forces per-sample downloads
  for (int done = 0; done < params.samples;) {
    if (progress_cb) progress_cb("trace image", done, params.samples);
    int n = std::min(batch, params.samples - done);
    pathtrace_b200::trace_samples(&state, scene, camera, params, n, nullptr);
    done += n;
    if (image_cb) image_cb(state.render, done, params.samples);
  }
  if (progress_cb) progress_cb("trace image", params.samples, params.samples);
  auto render = state.render;
  release(&state);
  return render;
}

ptr::intersection3f intersect_scene_bvh(const ptr::scene* scene, const ray3f& ray, bool find_any, bool non_rigid_frames) {
  if (!non_rigid_frames) throw std::runtime_error("yocto_pathtrace_b200: rigid-only inverse frames are not built on the device");
  yh_ray r{{ray.o.x, ray.o.y, ray.o.z}, {ray.d.x, ray.d.y, ray.d.z}, ray.tmin, ray.tmax};
  yh_hit h{};
  check(yh_intersect_batch(find_scene(scene), &r, 1, find_any ? 1 : 0, &h), "yh_intersect_batch");
  auto isec = ptr::intersection3f{};
  isec.hit  = h.hit != 0;
  if (isec.hit) isec.object = h.object, isec.element = h.element, isec.uv = {h.uv[0], h.uv[1]}, isec.distance = h.distance;
  return isec;
}

void sync_pixels(ptr::state* state) {
  auto&  rec = find_state(state);
  size_t n   = state->pixels.count();
  std::vector<float>    accum(4 * n);
  std::vector<int>      samples(n);
  std::vector<uint64_t> rng(2 * n);
  if (rec.group) {  // totals over the group; the per-pixel generators stay on their devices (rng reads 0)
    check(yh_group_download(rec.group, nullptr, accum.data(), samples.data()), "yh_group_download");
  } else {
    check(yh_state_download(rec.handle, nullptr, accum.data(), samples.data()), "yh_state_download");
    check(yh_state_download_rng(rec.handle, rng.data()), "yh_state_download_rng");
  }
  for (size_t i = 0; i < n; i++) {
    auto& px = state->pixels.data()[i];
    memcpy(&px.accumulated, &accum[4 * i], 16);
    px.samples   = samples[i];
    px.rng.state = rng[2 * i], px.rng.inc = rng[2 * i + 1];
  }
}

trace_stats last_stats(const ptr::state* state) {
  yh_trace_stats s{};
  auto&          rec = find_state(state);
  if (rec.group) check(yh_group_stats(rec.group, &s, nullptr), "yh_group_stats");
  else check(yh_state_stats(rec.handle, &s), "yh_state_stats");
  return {s.samples, s.rays, s.light_rays, s.launches, s.device_ms};
}

void release(const ptr::state* state) {
  std::lock_guard<std::mutex> lock(table_mutex);
  auto it = state_table.find(state);
  if (it == state_table.end()) return;
  if (it->second.handle) yh_state_destroy(it->second.handle);  // a group owns its states
  state_table.erase(it);
}
void release(const ptr::scene* scene) {
  std::lock_guard<std::mutex> lock(table_mutex);
  for (auto it = state_table.begin(); it != state_table.end();)
    if (it->second.scene == scene) {
      if (it->second.handle) yh_state_destroy(it->second.handle);
      it = state_table.erase(it);
    } else {
      ++it;
    }
  auto it = scene_table.find(scene);
  if (it == scene_table.end()) return;
  if (it->second.group) yh_group_destroy(it->second.group);
  if (it->second.scene) yh_scene_destroy(it->second.scene);
  scene_table.erase(it);
}

}  // namespace yocto::pathtrace_b200

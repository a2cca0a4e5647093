// See yh_sceneio.h.  Compiled against the reference tree (its sceneio::model, image loaders and
// the vendored nlohmann json header); nothing of the reference is copied here.
#include "yh_sceneio.h"

#include <climits>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <yocto/ext/json.hpp>
#include <yocto/yocto_image.h>
#include <yocto/yocto_ply.h>
#include <yocto/yocto_shape.h>

#include <array>
#include <cstring>
#include <filesystem>
#include <future>
#include <map>

// fix-ups the reference applies at the end of load_json_scene (yocto_sceneio.cpp:1408-1412); they
// have external linkage in the reference's translation unit (:367-440) but are not in its header
namespace yocto::sceneio {
void add_cameras(model* scene);
void add_radius(model* scene, float radius);
void add_materials(model* scene);
void trim_memory(model* scene);
}  // namespace yocto::sceneio

namespace yocto::pathtrace_b200 {

using json    = nlohmann::json;
namespace sfs = std::filesystem;
using yocto::math::frame3f;
using yocto::math::mat3f;

// ------------------------------------------------------------------ PLY ------------------------

namespace {

enum class ptype { i8, i16, i32, i64, u8, u16, u32, u64, f32, f64, bad };
struct pprop {
  std::string name;
  ptype       type = ptype::bad;
  bool        list = false;
};
struct pelem {
  std::string        name;
  size_t             count = 0;
  std::vector<pprop> props;
};

ptype type_of(const std::string& t) {
  static const std::map<std::string, ptype> names = {{"char", ptype::i8}, {"short", ptype::i16}, {"int", ptype::i32},
      {"long", ptype::i64}, {"uchar", ptype::u8}, {"ushort", ptype::u16}, {"uint", ptype::u32}, {"ulong", ptype::u64},
      {"float", ptype::f32}, {"double", ptype::f64}, {"int8", ptype::i8}, {"int16", ptype::i16}, {"int32", ptype::i32},
      {"int64", ptype::i64}, {"uint8", ptype::u8}, {"uint16", ptype::u16}, {"uint32", ptype::u32}, {"uint64", ptype::u64},
      {"float32", ptype::f32}, {"float64", ptype::f64}};
  auto it = names.find(t);
  return it == names.end() ? ptype::bad : it->second;
}
size_t size_of(ptype t) {
  switch (t) {
    case ptype::i8: case ptype::u8: return 1;
    case ptype::i16: case ptype::u16: return 2;
    case ptype::i32: case ptype::u32: case ptype::f32: return 4;
    case ptype::i64: case ptype::u64: case ptype::f64: return 8;
    default: return 0;
  }
}

struct mapped_file {
  const unsigned char* data = nullptr;
  size_t               size = 0;
  int                  fd   = -1;
  ~mapped_file() {
    if (data) munmap((void*)data, size);
    if (fd >= 0) close(fd);
  }
  bool open(const std::string& filename) {
    fd = ::open(filename.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) return false;
    size = (size_t)st.st_size;
    if (size == 0) return true;
    void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    if (p == MAP_FAILED) return false;
    madvise(p, size, MADV_SEQUENTIAL);
    data = (const unsigned char*)p;
    return true;
  }
};

// Header: "ply" / format / comment / obj_info / element / property ... / end_header, one per line.
// ok = false: malformed; fast = false: well-formed but outside the bulk path.
struct pheader {
  std::vector<pelem> elems;
  size_t             body = 0;
  bool               ok = false, fast = true;
};
pheader parse_header(const mapped_file& f) {
  pheader h;
  size_t  pos = 0;
  bool    first = true, ended = false;
  while (pos < f.size && !ended) {
    size_t eol = pos;
    while (eol < f.size && f.data[eol] != '\n') eol++;
    std::string line((const char*)f.data + pos, eol - pos);
    pos = eol + 1;
    if (auto c = line.find('#'); c != std::string::npos) line.resize(c);
    std::vector<std::string> tok;
    for (size_t i = 0; i < line.size();) {
      while (i < line.size() && isspace((unsigned char)line[i])) i++;
      size_t j = i;
      while (j < line.size() && !isspace((unsigned char)line[j])) j++;
      if (j > i) tok.emplace_back(line.substr(i, j - i));
      i = j;
    }
    if (tok.empty()) continue;
    if (first) {
      if (tok[0] != "ply") return h;
      first = false;
      continue;
    }
    if (tok[0] == "format") {
      if (tok.size() < 2) return h;
      if (tok[1] == "ascii" || tok[1] == "binary_big_endian") h.fast = false;
      else if (tok[1] != "binary_little_endian") return h;
    } else if (tok[0] == "comment" || tok[0] == "obj_info") {
    } else if (tok[0] == "element") {
      if (tok.size() < 3) return h;
      pelem e;
      e.name = tok[1];
      try {
        e.count = (size_t)std::stoull(tok[2]);
      } catch (...) {
        return h;
      }
      h.elems.push_back(e);
    } else if (tok[0] == "property") {
      if (h.elems.empty() || tok.size() < 3) return h;
      pprop p;
      if (tok[1] == "list") {
        if (tok.size() < 5) return h;
        if (type_of(tok[2]) != ptype::u8) return h;  // the reference accepts uchar list sizes only (yocto_ply.h:581)
        p.list = true, p.type = type_of(tok[3]), p.name = tok[4];
      } else {
        p.type = type_of(tok[1]), p.name = tok[2];
      }
      if (p.type == ptype::bad) return h;
      h.elems.back().props.push_back(p);
    } else if (tok[0] == "end_header") {
      ended = true;
    } else {
      return h;
    }
  }
  if (!ended) return h;
  h.body = pos;
  h.ok   = true;
  return h;
}

inline float load_f32(const unsigned char* p) {
  float v;
  memcpy(&v, p, 4);
  return v;
}
inline int load_i32(const unsigned char* p) {
  int v;
  memcpy(&v, p, 4);
  return v;
}

}  // namespace

bool load_shape_ply(const std::string& filename, std::vector<int>& points, std::vector<vec2i>& lines,
    std::vector<vec3i>& triangles, std::vector<vec4i>& quads, std::vector<vec3f>& positions,
    std::vector<vec3f>& normals, std::vector<vec2f>& texcoords, std::vector<vec3f>& colors,
    std::vector<float>& radius, std::string& error, bool& handled, bool flip_texcoord) {
  points = {}, lines = {}, triangles = {}, quads = {}, positions = {}, normals = {}, texcoords = {}, colors = {}, radius = {};
  handled = true;
  mapped_file file;
  if (!file.open(filename)) {
    error = filename + ": file not found";
    return false;
  }
  pheader h = parse_header(file);
  if (!h.ok) {
    error = filename + ": parse error";
    return false;
  }
  // the bulk path: scalar elements whose wanted properties are float, list elements with one int list
  auto wanted_vertex = [](const std::string& n) {
    static const char* names[] = {"x", "y", "z", "nx", "ny", "nz", "u", "v", "s", "t", "red", "green", "blue", "radius"};
    for (auto w : names)
      if (n == w) return true;
    return false;
  };
  for (auto& e : h.elems) {
    bool any_list = false;
    for (auto& p : e.props) any_list |= p.list;
    if (any_list) {
      if (e.props.size() != 1) h.fast = false;
      else if (e.props[0].type != ptype::i32 && e.props[0].type != ptype::u32) h.fast = false;
    } else if (e.name == "vertex") {
      for (auto& p : e.props)
        if (wanted_vertex(p.name) && p.type != ptype::f32) h.fast = false;
    }
  }
  if (!h.fast) {
    handled = false;
    return false;
  }
  auto truncated = [&]() {
    error = filename + ": read error";
    return false;
  };

  const unsigned char* cur = file.data + h.body;
  const unsigned char* end = file.data + file.size;
  std::vector<unsigned char> face_sizes, line_sizes;
  std::vector<int>           face_idx, line_idx;
  bool                       has_face = false, has_line = false, has_point = false;
  for (auto& e : h.elems) {
    if (!e.props.empty() && e.props[0].list) {
      // records: one size byte, then that many 32-bit indices
      bool is_face = e.name == "face", is_line = e.name == "line", is_point = e.name == "point";
      bool keep = e.props[0].name == "vertex_indices" && (is_face || is_line || is_point);
      auto& sizes = is_face ? face_sizes : line_sizes;
      auto& idx   = is_face ? face_idx : (is_line ? line_idx : points);
      if (keep) {
        (is_face ? has_face : (is_line ? has_line : has_point)) = true;
        if (!is_point) sizes.reserve(e.count);
        idx.reserve(e.count * 3);
      }
      for (size_t r = 0; r < e.count; r++) {
        if (cur >= end) return truncated();
        unsigned n = *cur++;
        if ((size_t)(end - cur) < 4 * (size_t)n) return truncated();
        if (keep) {
          if (!is_point) sizes.push_back((unsigned char)n);
          size_t at = idx.size();
          idx.resize(at + n);
          memcpy(idx.data() + at, cur, 4 * (size_t)n);
        }
        cur += 4 * (size_t)n;
      }
      continue;
    }
    size_t stride = 0;
    for (auto& p : e.props) stride += size_of(p.type);
    if ((size_t)(end - cur) < stride * e.count) return truncated();
    if (e.name == "vertex") {
      std::map<std::string, size_t> off;
      size_t                         o = 0;
      for (auto& p : e.props) {
        if (!off.count(p.name)) off[p.name] = o;  // first of a repeated name, as get_property does
        o += size_of(p.type);
      }
      auto has = [&](std::initializer_list<const char*> names) {
        for (auto n : names)
          if (!off.count(n)) return false;
        return true;
      };
      const size_t n = e.count;
      auto vec3s = [&](const char* a, const char* b, const char* c, std::vector<vec3f>& out) {
        if (!has({a, b, c})) return;
        size_t oa = off[a], ob = off[b], oc = off[c];
        out.resize(n);
        const unsigned char* row = cur;
        for (size_t i = 0; i < n; i++, row += stride) out[i] = {load_f32(row + oa), load_f32(row + ob), load_f32(row + oc)};
      };
      vec3s("x", "y", "z", positions);
      vec3s("nx", "ny", "nz", normals);
      vec3s("red", "green", "blue", colors);
      const char *tu = has({"u"}) ? "u" : "s", *tv = has({"u"}) ? "v" : "t";  // get_texcoords, yocto_ply.h:1083-1094
      if (has({tu, tv})) {
        size_t oa = off[tu], ob = off[tv];
        texcoords.resize(n);
        const unsigned char* row = cur;
        for (size_t i = 0; i < n; i++, row += stride) {
          float v      = load_f32(row + ob);
          texcoords[i] = {load_f32(row + oa), flip_texcoord ? 1 - v : v};
        }
      }
      if (has({"radius"})) {
        size_t oa = off["radius"];
        radius.resize(n);
        const unsigned char* row = cur;
        for (size_t i = 0; i < n; i++, row += stride) radius[i] = load_f32(row + oa);
      }
    }
    cur += stride * e.count;
  }

  // faces -> quads if any record has four vertices, else triangle fans (yocto_ply.h:1104-1144, :1164-1170)
  if (has_face) {
    bool any_quad = false;
    for (auto s : face_sizes) any_quad |= s == 4;
    size_t at = 0;
    if (any_quad) {
      quads.reserve(face_sizes.size());
      for (auto s : face_sizes) {
        if (s == 4) quads.push_back({face_idx[at], face_idx[at + 1], face_idx[at + 2], face_idx[at + 3]});
        else
          for (int c = 2; c < s; c++) quads.push_back({face_idx[at], face_idx[at + c - 1], face_idx[at + c], face_idx[at + c]});
        at += s;
      }
    } else {
      triangles.reserve(face_sizes.size());
      for (auto s : face_sizes) {
        for (int c = 2; c < s; c++) triangles.push_back({face_idx[at], face_idx[at + c - 1], face_idx[at + c]});
        at += s;
      }
    }
  }
  // polylines -> segments (yocto_ply.h:1145-1160)
  if (has_line) {
    size_t total = 0;
    for (auto s : line_sizes) total += s > 0 ? s - 1 : 0;
    lines.reserve(total);
    size_t at = 0;
    for (auto s : line_sizes) {
      for (int c = 1; c < s; c++) lines.push_back({line_idx[at + c - 1], line_idx[at + c]});
      at += s;
    }
  }
  (void)has_point;
  if (positions.empty()) {
    error = filename + ": empty shape";
    return false;
  }
  return true;
}

// ------------------------------------------------------------------ JSON -----------------------

namespace {

template <size_t N>
bool floats(const json& j, float* out) {
  if (!j.is_array() || j.size() != N) return false;
  for (size_t i = 0; i < N; i++) {
    if (!j[i].is_number()) return false;
    out[i] = j[i].get<float>();
  }
  return true;
}
// a missing key keeps the default; a present key of the wrong shape is an error
bool value(const json& e, const char* key, float& v) {
  if (!e.contains(key)) return true;
  if (!e[key].is_number()) return false;
  v = e[key].get<float>();
  return true;
}
bool value(const json& e, const char* key, int& v) {
  if (!e.contains(key)) return true;
  if (!e[key].is_number()) return false;
  v = e[key].get<int>();
  return true;
}
bool value(const json& e, const char* key, bool& v) {
  if (!e.contains(key)) return true;
  if (!e[key].is_boolean()) return false;
  v = e[key].get<bool>();
  return true;
}
bool value(const json& e, const char* key, std::string& v) {
  if (!e.contains(key)) return true;
  if (!e[key].is_string()) return false;
  v = e[key].get<std::string>();
  return true;
}
bool value(const json& e, const char* key, vec3f& v) { return !e.contains(key) || floats<3>(e[key], &v.x); }
bool value(const json& e, const char* key, frame3f& v) { return !e.contains(key) || floats<12>(e[key], &v.x.x); }
bool value(const json& e, const char* key, mat3f& v) { return !e.contains(key) || floats<9>(e[key], &v.x.x); }

// name -> element, created on first use; creation order is the order of scene->shapes / textures
template <typename T, typename Add>
bool named(const json& e, const char* key, T*& out, std::map<std::string, T*>& seen, std::vector<std::string>& order, Add add) {
  std::string name;
  if (!value(e, key, name)) return false;
  if (name.empty()) return true;
  auto it = seen.find(name);
  if (it == seen.end()) {
    it = seen.emplace(name, add(name)).first;
    order.push_back(name);
  }
  out = it->second;
  return true;
}

sfs::path resolve(const std::string& scene_file, const std::string& name, const char* group,
    std::initializer_list<const char*> extensions) {
  auto dir = sfs::path(scene_file).parent_path() / group;
  for (auto ext : extensions)
    if (sfs::exists(dir / (name + ext))) return dir / (name + ext);
  return dir / (name + *extensions.begin());
}

bool load_texture_file(const std::string& path, sio::texture* t, bool scalar, std::string& error) {
  bool hdr = yocto::image::is_hdr_filename(path);
  if (scalar) return hdr ? yocto::image::load_image(path, t->scalarf, error) : yocto::image::load_image(path, t->scalarb, error);
  return hdr ? yocto::image::load_image(path, t->colorf, error) : yocto::image::load_image(path, t->colorb, error);
}

}  // namespace

bool load_scene(const std::string& filename, sio::model* scene, std::string& error, sio::progress_callback progress_cb) {
  auto ext = sfs::path(filename).extension().string();
  if (ext != ".json" && ext != ".JSON") return sio::load_scene(filename, scene, error, progress_cb);
  if (progress_cb) progress_cb("load scene", 0, 2);

  json js;
  {
    mapped_file f;
    if (!f.open(filename)) {
      error = filename + ": file not found";
      return false;
    }
    try {
      js = json::parse((const char*)f.data, (const char*)f.data + f.size);
    } catch (std::exception&) {
      error = filename + ": parse error in json";
      return false;
    }
  }
  auto bad = [&](const std::string& what) {
    error = filename + ": parse error in " + what;
    return false;
  };

  std::map<std::string, sio::texture*>  ctex, stex;
  std::map<std::string, sio::shape*>    shapes;
  std::map<std::string, sio::subdiv*>   subdivs;
  std::map<std::string, sio::instance*> instances;
  std::map<std::string, sio::material*> materials;
  std::vector<std::string>              ctex_order, stex_order, shape_order, subdiv_order, instance_order;
  auto color_tex  = [&](const json& e, const char* key, sio::texture*& out) {
    return named(e, key, out, ctex, ctex_order, [&](const std::string& n) { return sio::add_texture(scene, n); });
  };
  auto scalar_tex = [&](const json& e, const char* key, sio::texture*& out) {
    return named(e, key, out, stex, stex_order, [&](const std::string& n) { return sio::add_texture(scene, n); });
  };

  if (js.contains("asset") && !value(js["asset"], "copyright", scene->copyright)) return bad("asset");
  if (js.contains("cameras")) {
    for (auto& [name, e] : js["cameras"].items()) {  // nlohmann objects iterate in key order
      auto c  = sio::add_camera(scene);
      c->name = name;
      if (!value(e, "frame", c->frame) || !value(e, "orthographic", c->orthographic) || !value(e, "lens", c->lens) ||
          !value(e, "aspect", c->aspect) || !value(e, "film", c->film) || !value(e, "focus", c->focus) ||
          !value(e, "aperture", c->aperture))
        return bad("camera " + name);
      if (e.contains("lookat")) {  // {from, to, up}: overrides the frame, focus on the target
        mat3f l = yocto::math::identity3x3f;
        if (!value(e, "lookat", l)) return bad("camera " + name);
        c->frame = yocto::math::lookat_frame(l.x, l.y, l.z);
        c->focus = yocto::math::length(l.x - l.y);
      }
    }
  }
  if (js.contains("environments")) {
    for (auto& [name, e] : js["environments"].items()) {
      auto env  = sio::add_environment(scene);
      env->name = name;
      if (!value(e, "frame", env->frame) || !value(e, "emission", env->emission) || !color_tex(e, "emission_tex", env->emission_tex))
        return bad("environment " + name);
      if (e.contains("lookat")) {
        mat3f l = yocto::math::identity3x3f;
        if (!value(e, "lookat", l)) return bad("environment " + name);
        env->frame = yocto::math::lookat_frame(l.x, l.y, l.z, true);
      }
    }
  }
  if (js.contains("materials")) {
    for (auto& [name, e] : js["materials"].items()) {
      auto m  = sio::add_material(scene);
      m->name = name;
      bool ok = value(e, "eumelanin", m->eumelanin) && value(e, "pheomelanin", m->pheomelanin) &&
                value(e, "sigma_a", m->sigma_a) && value(e, "beta_m", m->beta_m) && value(e, "beta_n", m->beta_n) &&
                value(e, "alpha", m->alpha) && value(e, "eta", m->eta) && value(e, "emission", m->emission) &&
                value(e, "color", m->color) && value(e, "metallic", m->metallic) && value(e, "specular", m->specular) &&
                value(e, "roughness", m->roughness) && value(e, "coat", m->coat) && value(e, "transmission", m->transmission) &&
                value(e, "translucency", m->translucency) && value(e, "thin", m->thin) && value(e, "ior", m->ior) &&
                value(e, "trdepth", m->trdepth) && value(e, "scattering", m->scattering) &&
                value(e, "scanisotropy", m->scanisotropy) && value(e, "opacity", m->opacity) &&
                value(e, "displacement", m->displacement) &&
                // texture creation order = this order (it fixes the order of scene->textures)
                color_tex(e, "emission_tex", m->emission_tex) && color_tex(e, "color_tex", m->color_tex) &&
                scalar_tex(e, "metallic_tex", m->metallic_tex) && scalar_tex(e, "specular_tex", m->specular_tex) &&
                scalar_tex(e, "transmission_tex", m->transmission_tex) && scalar_tex(e, "translucency_tex", m->translucency_tex) &&
                scalar_tex(e, "roughness_tex", m->roughness_tex) && color_tex(e, "scattering_tex", m->scattering_tex) &&
                scalar_tex(e, "opacity_tex", m->opacity_tex) && color_tex(e, "normal_tex", m->normal_tex) &&
                scalar_tex(e, "displacement_tex", m->displacement_tex) && value(e, "subdivisions", m->subdivisions) &&
                value(e, "smooth", m->smooth);
      if (!ok) return bad("material " + name);
      materials[name] = m;
    }
  }
  if (js.contains("objects")) {
    for (auto& [name, e] : js["objects"].items()) {
      auto o  = sio::add_object(scene);
      o->name = name;
      if (!value(e, "frame", o->frame)) return bad("object " + name);
      if (e.contains("lookat")) {
        mat3f l = yocto::math::identity3x3f;
        if (!value(e, "lookat", l)) return bad("object " + name);
        o->frame = yocto::math::lookat_frame(l.x, l.y, l.z, true);
      }
      std::string mat;
      if (!value(e, "material", mat)) return bad("object " + name);
      if (e.contains("material")) {
        if (mat.empty()) o->material = nullptr;
        else if (!materials.count(mat)) {
          error = filename + ": missing material " + mat;
          return false;
        } else o->material = materials[mat];
      }
      bool ok = named(e, "shape", o->shape, shapes, shape_order, [&](const std::string& n) { return sio::add_shape(scene, n); }) &&
                named(e, "subdiv", o->subdiv, subdivs, subdiv_order, [&](const std::string& n) { return sio::add_subdiv(scene, n); }) &&
                named(e, "instance", o->instance, instances, instance_order,
                    [&](const std::string& n) { return sio::add_instance(scene, n); });
      if (!ok) return bad("object " + name);
    }
  }

  int total = 2 + (int)(shapes.size() + subdivs.size() + ctex.size() + stex.size() + instances.size()), done = 1;
  if (progress_cb) progress_cb("load scene", done++, total);
  auto dependent = [&](const std::string& e) {
    error = filename + ": error in " + e;
    return false;
  };

  // shapes: independent files, read concurrently (bulk PLY path, else the reference's loader)
  std::vector<std::future<std::string>> jobs;
  for (auto& name : shape_order) {
    auto s    = shapes[name];
    auto path = resolve(filename, name, "shapes", {".ply", ".obj"}).string();
    jobs.push_back(std::async(std::launch::async, [s, path]() -> std::string {
      std::string err;
      auto pext = sfs::path(path).extension().string();
      if (pext == ".ply" || pext == ".PLY") {
        bool handled = false;
        if (load_shape_ply(path, s->points, s->lines, s->triangles, s->quads, s->positions, s->normals, s->texcoords,
                s->colors, s->radius, err, handled))
          return "";
        if (handled) return err;
      }
      if (!yocto::shape::load_shape(path, s->points, s->lines, s->triangles, s->quads, s->positions, s->normals,
              s->texcoords, s->colors, s->radius, err))
        return err;
      return "";
    }));
  }
  // textures, on this thread meanwhile
  std::string tex_error;
  for (auto& name : ctex_order) {
    if (progress_cb) progress_cb("load texture", done++, total);
    auto path = resolve(filename, name, "textures", {".hdr", ".exr", ".png", ".jpg"}).string();
    if (tex_error.empty() && !load_texture_file(path, ctex[name], false, tex_error)) break;
  }
  for (auto& name : stex_order) {
    if (progress_cb) progress_cb("load texture", done++, total);
    auto path = resolve(filename, name, "textures", {".hdr", ".exr", ".png", ".jpg"}).string();
    if (tex_error.empty() && !load_texture_file(path, stex[name], true, tex_error)) break;
  }
  std::string shape_error;
  for (auto& j : jobs) {
    if (progress_cb) progress_cb("load shape", done++, total);
    auto e = j.get();
    if (shape_error.empty()) shape_error = e;
  }
  if (!shape_error.empty()) return dependent(shape_error);
  if (!tex_error.empty()) return dependent(tex_error);
  for (auto& name : subdiv_order) {
    if (progress_cb) progress_cb("load subdiv", done++, total);
    auto sd   = subdivs[name];
    auto path = resolve(filename, name, "subdivs", {".obj"}).string();
    std::string err;
    if (!yocto::shape::load_fvshape(path, sd->quadspos, sd->quadsnorm, sd->quadstexcoord, sd->positions, sd->normals, sd->texcoords, err))
      return dependent(err);
  }
  for (auto& name : instance_order) {  // instances/<name>.ply: element "instance", 12 float properties per frame
    if (progress_cb) progress_cb("load instance", done++, total);
    auto path = resolve(filename, name, "instances", {".ply"}).string();
    auto pext = sfs::path(path).extension().string();
    if (pext != ".ply" && pext != ".PLY") return dependent(path + ": unknown format");
    yocto::ply::model ply;
    std::string       err;
    if (!yocto::ply::load_ply(path, &ply, err)) return dependent(err);
    yocto::ply::get_values(&ply, "instance", {"xx", "xy", "xz", "yx", "yy", "yz", "zx", "zy", "zz", "ox", "oy", "oz"},
        instances[name]->frames);
  }

  if (scene->name.empty()) scene->name = sfs::path(filename).stem().string();
  sio::add_cameras(scene);
  sio::add_radius(scene, 0.001f);
  sio::add_materials(scene);
  sio::trim_memory(scene);
  if (progress_cb) progress_cb("load scene", total, total);
  return true;
}

}  // namespace yocto::pathtrace_b200

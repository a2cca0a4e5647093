// Shading-side device functions of trace_path (yocto_pathtrace.cpp:1380-1511):
// shading point set-up, emission, diffuse + hair lobes, environment lookup,
// light sampling and its pdf.  Same arithmetic as the reference, file:line
// cited per function.
#pragma once
#include "yh_hair.cuh"
#include "yh_rng.cuh"
#include "yh_surface.cuh"
#include "yh_traverse.cuh"

namespace yh {

__device__ __forceinline__ vec3f ldp(const float* p, int i) {
  return {__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)};
}

// eval_position, yocto_pathtrace.cpp:232-250 (lines: on the fibre AXIS)
__device__ __forceinline__ vec3f eval_position(const dev_scene& sc, int object, int element, const vec2f& uv) {
  const dev_object& ob = sc.objects[object];
  const dev_shape&  sh = sc.shapes[ob.shape];
  if (sh.is_points) return transform_point(ob.frame, ldp(sh.positions, __ldg(sh.elems + element)));
  if (!sh.is_lines) {
    int   a = __ldg(sh.elems + 3 * element), b = __ldg(sh.elems + 3 * element + 1), c = __ldg(sh.elems + 3 * element + 2);
    vec3f p = ldp(sh.positions, a) * (1 - uv.x - uv.y) + ldp(sh.positions, b) * uv.x + ldp(sh.positions, c) * uv.y;
    return transform_point(ob.frame, p);
  }
  int   a = __ldg(sh.elems + 2 * element), b = __ldg(sh.elems + 2 * element + 1);
  vec3f p = ldp(sh.positions, a) * (1 - uv.x) + ldp(sh.positions, b) * uv.x;
  return transform_point(ob.frame, p);
}
// eval_element_normal, :253-269
__device__ __forceinline__ vec3f eval_element_normal(const dev_scene& sc, int object, int element) {
  const dev_object& ob = sc.objects[object];
  const dev_shape&  sh = sc.shapes[ob.shape];
  if (sh.is_points) return {0, 0, 1};  // :264-265
  if (!sh.is_lines) {
    int   a = __ldg(sh.elems + 3 * element), b = __ldg(sh.elems + 3 * element + 1), c = __ldg(sh.elems + 3 * element + 2);
    vec3f p0 = ldp(sh.positions, a), p1 = ldp(sh.positions, b), p2 = ldp(sh.positions, c);
    return transform_direction(ob.frame, normalize(cross(p1 - p0, p2 - p0)));
  }
  int a = __ldg(sh.elems + 2 * element), b = __ldg(sh.elems + 2 * element + 1);
  return transform_direction(ob.frame, normalize(ldp(sh.positions, b) - ldp(sh.positions, a)));
}
// eval_normal, :272-292 (transform_normal with non_rigid = false normalises)
__device__ __forceinline__ vec3f eval_normal(const dev_scene& sc, int object, int element, const vec2f& uv) {
  const dev_object& ob = sc.objects[object];
  const dev_shape&  sh = sc.shapes[ob.shape];
  if (!sh.has_normals) return eval_element_normal(sc, object, element);
  if (sh.is_points) return transform_direction(ob.frame, normalize(ldp(sh.normals, __ldg(sh.elems + element))));  // :286-288
  if (!sh.is_lines) {
    int   a = __ldg(sh.elems + 3 * element), b = __ldg(sh.elems + 3 * element + 1), c = __ldg(sh.elems + 3 * element + 2);
    vec3f n = ldp(sh.normals, a) * (1 - uv.x - uv.y) + ldp(sh.normals, b) * uv.x + ldp(sh.normals, c) * uv.y;
    return transform_direction(ob.frame, normalize(n));
  }
  int   a = __ldg(sh.elems + 2 * element), b = __ldg(sh.elems + 2 * element + 1);
  vec3f n = ldp(sh.normals, a) * (1 - uv.x) + ldp(sh.normals, b) * uv.x;
  return transform_direction(ob.frame, normalize(n));
}

// lookup_texture / eval_texture, :148-200 (float RGB textures, bilinear, tiled)
__device__ __forceinline__ vec3f lookup_texture(const dev_texture& t, int i, int j) {
  const float* p = t.pixels + 3 * ((size_t)j * t.width + i);
  return {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
}
__device__ __forceinline__ vec3f eval_texture(const dev_scene& sc, int tex, const vec2f& uv) {
  if (tex < 0) return {1, 1, 1};
  const dev_texture& t = sc.textures[tex];
  float s = fmodf_(uv.x, 1.0f) * t.width;
  if (s < 0) s += t.width;
  float tt = fmodf_(uv.y, 1.0f) * t.height;
  if (tt < 0) tt += t.height;
  int   i = clampi((int)s, 0, t.width - 1), j = clampi((int)tt, 0, t.height - 1);
  int   ii = (i + 1) % t.width, jj = (j + 1) % t.height;
  float u = s - i, v = tt - j;
  return lookup_texture(t, i, j) * (1 - u) * (1 - v) + lookup_texture(t, i, jj) * (1 - u) * v +
         lookup_texture(t, ii, j) * u * (1 - v) + lookup_texture(t, ii, jj) * u * v;
}

// eval_texcoord, :294-309 (interpolate_triangle / interpolate_line of the vertex texcoords, the element uv without them)
__device__ __forceinline__ vec2f eval_texcoord(const dev_scene& sc, int object, int element, const vec2f& uv) {
  const dev_shape& sh = sc.shapes[sc.objects[object].shape];
  if (!sh.texcoords) return uv;
  auto ldt = [&](int i) -> vec2f { return {__ldg(sh.texcoords + 2 * i), __ldg(sh.texcoords + 2 * i + 1)}; };
  if (sh.is_points) return ldt(__ldg(sh.elems + element));  // :306-307
  if (!sh.is_lines) {
    int   a = __ldg(sh.elems + 3 * element), b = __ldg(sh.elems + 3 * element + 1), c = __ldg(sh.elems + 3 * element + 2);
    vec2f t0 = ldt(a), t1 = ldt(b), t2 = ldt(c);
    float w = 1 - uv.x - uv.y;
    return {t0.x * w + t1.x * uv.x + t2.x * uv.y, t0.y * w + t1.y * uv.x + t2.y * uv.y};
  }
  int   a = __ldg(sh.elems + 2 * element), b = __ldg(sh.elems + 2 * element + 1);
  vec2f t0 = ldt(a), t1 = ldt(b);
  return {t0.x * (1 - uv.x) + t1.x * uv.x, t0.y * (1 - uv.x) + t1.y * uv.x};
}

// eval_normalmap, :326-347, with eval_element_tangents :312-324 and triangle_tangents_fromuv (yocto_math.h:3360-3383):
// the tangent-space normal of the texture in the frame built from the uv tangents and the shading normal
__device__ __forceinline__ vec3f eval_normalmap(const dev_scene& sc, int object, int element, const vec2f& uv, int normal_tex) {
  const dev_object& ob = sc.objects[object];
  const dev_shape&  sh = sc.shapes[ob.shape];
  vec3f normal   = eval_normal(sc, object, element, uv);
  vec2f texcoord = eval_texcoord(sc, object, element, uv);
  if (sh.is_lines) return normal;
  vec3f t2 = 2 * eval_texture(sc, normal_tex, texcoord);
  vec3f nm = {-1 + t2.x, -1 + t2.y, -1 + t2.z};  // -1 + 2 * texture (operator+(float, vec3f), yocto_math.h:1970)
  vec3f tu = {0, 0, 0}, tv = {0, 0, 0};  // a shape without texcoords has no tangents (`return {}`)
  if (sh.texcoords) {
    int   a = __ldg(sh.elems + 3 * element), b = __ldg(sh.elems + 3 * element + 1), c = __ldg(sh.elems + 3 * element + 2);
    vec3f p0 = ldp(sh.positions, a), p1 = ldp(sh.positions, b), p2 = ldp(sh.positions, c);
    auto  ldt = [&](int i) -> vec2f { return {__ldg(sh.texcoords + 2 * i), __ldg(sh.texcoords + 2 * i + 1)}; };
    vec2f uv0 = ldt(a), uv1 = ldt(b), uv2 = ldt(c);
    vec3f p = p1 - p0, q = p2 - p0;
    vec2f s = {uv1.x - uv0.x, uv2.x - uv0.x}, t = {uv1.y - uv0.y, uv2.y - uv0.y};
    float d = s.x * t.y - s.y * t.x;
    vec3f eu = {1, 0, 0}, ev = {0, 1, 0};
    if (d != 0) {
      eu = vec3f{t.y * p.x - t.x * q.x, t.y * p.y - t.x * q.y, t.y * p.z - t.x * q.z} / d;
      ev = vec3f{s.x * q.x - s.y * p.x, s.x * q.y - s.y * p.y, s.x * q.z - s.y * p.z} / d;
    }
    tu = transform_direction(ob.frame, eu), tv = transform_direction(ob.frame, ev);
  }
  vec3f fx = orthonormalize(tu, normal);
  vec3f fy = normalize(cross(normal, fx));
  bool  flip_v = dot(fy, tv) < 0;
  nm.y *= flip_v ? 1.0f : -1.0f;
  return normalize(fx * nm.x + fy * nm.y + normal * nm.z);
}

// eval_environment, :536-547
__device__ __forceinline__ vec3f eval_environment(const dev_scene& sc, const vec3f& dir) {
  vec3f emission = {0, 0, 0};
  for (int e = 0; e < sc.num_environments; e++) {
    const dev_environment& env = sc.environments[e];
    vec3f tex = {1, 1, 1};
    if (env.tex >= 0) {
      vec3f wl       = transform_direction(env.inv_frame, dir);
      vec2f texcoord = {div_(atan2f_(wl.z, wl.x), (2 * pif)), div_(acosf_(clampf(wl.y, -1.0f, 1.0f)), pif)};
      if (texcoord.x < 0) texcoord.x += 1;
      tex = eval_texture(sc, env.tex, texcoord);
    }
    emission = emission + env.emission * tex;
  }
  return emission;
}

// sample_discrete_cdf, yocto_math.h:4953-4958 (std::upper_bound).  With a guide
// table the search runs over the few elements between guide[b - 1] and
// guide[b + 2] (b = bucket of r): every element before the window is <= r and
// the element at its end is > r, so the first element > r is the same one the
// full search finds.
__device__ __forceinline__ int sample_discrete_cdf(const dev_light& light, float r) {
  const float* cdf = light.cdf;
  int   n    = light.cdf_size;
  float back = __ldg(cdf + n - 1);
  r          = clampf(r * back, 0.0f, back - 0.00001f);
  int lo = 0, len = n;
  if (light.guide) {
    int b  = clampi((int)(r * light.guide_inv_w), 0, light.guide_size - 1);
    int b0 = b > 0 ? b - 1 : 0, b1 = b + 2 < light.guide_size ? b + 2 : light.guide_size;
    // the margins of one bucket absorb the rounding of r * guide_inv_w; checked anyway
    if ((float)b0 * light.guide_w <= r && r <= (float)b1 * light.guide_w) {
      lo  = __ldg(light.guide + b0);
      len = __ldg(light.guide + b1) - lo;
    }
  }
  while (len > 0) {  // first element > r
    int half = len >> 1;
    if (!(r < __ldg(cdf + lo + half))) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  return clampi(lo, 0, n - 1);
}

// sample_lights, yocto_pathtrace.cpp:1283-1308
__device__ __forceinline__ vec3f sample_lights(const dev_scene& sc, const vec3f& position, float rl, float rel,
    const vec2f& ruv) {
  int              light_id = clampi((int)(rl * sc.num_lights), 0, sc.num_lights - 1);  // sample_uniform
  const dev_light& light    = sc.lights[light_id];
  if (light.object >= 0) {
    int   element = sample_discrete_cdf(light, rel);
    vec2f uv      = {1 - sqrt_(ruv.x), ruv.y * sqrt_(ruv.x)};  // sample_triangle, yocto_math.h:4911
    vec3f lpos    = eval_position(sc, light.object, element, uv);
    return normalize(lpos - position);
  }
  const dev_environment& env = sc.environments[light.environment];
  if (env.tex >= 0) {
    const dev_texture& t = sc.textures[env.tex];
    int   idx = sample_discrete_cdf(light, rel);
    vec2f uv  = {div_((idx % t.width + 0.5f), (float)t.width), div_((idx / t.width + 0.5f), (float)t.height)};
    return transform_direction(env.frame, {cosf_(uv.x * 2 * pif) * sinf_(uv.y * pif), cosf_(uv.y * pif),
                                              sinf_(uv.x * 2 * pif) * sinf_(uv.y * pif)});
  }
  // sample_sphere, yocto_math.h:4847-4852
  float z   = 2 * ruv.y - 1;
  float r   = sqrt_(clampf(1 - z * z, 0.0f, 1.0f));
  float phi = 2 * pif * ruv.x;
  return {r * cosf_(phi), r * sinf_(phi), z};
}

// sample_lights_pdf, :1311-1358.  Area lights re-intersect the light object up
// to 100 times along the direction (intersect_instance_bvh).
__device__ __forceinline__ float sample_lights_pdf(const dev_scene& sc, const vec3f& position,
    const vec3f& direction, trav_stack& st, unsigned int& light_rays) {
  float pdf = 0.0f;
  for (int l = 0; l < sc.num_lights; l++) {
    const dev_light& light = sc.lights[l];
    if (light.object >= 0) {
      float lpdf          = 0.0f;
      vec3f next_position = position;
      for (int bounce = 0; bounce < 100; bounce++) {
        int   element;
        vec2f uv;
        float dist;
        light_rays++;
        ray3f r = {next_position, direction, ray_eps, flt_max};
        if (!intersect_instance(sc, light.object, r, st, element, uv, dist, false)) break;
        vec3f lposition = eval_position(sc, light.object, element, uv);
        vec3f lnormal   = eval_element_normal(sc, light.object, element);
        float area      = __ldg(light.cdf + light.cdf_size - 1);
        vec3f dp        = lposition - position;
        lpdf += div_(dot(dp, dp), (fabs_(dot(lnormal, direction)) * area));
        next_position = lposition + direction * 1e-3f;
      }
      pdf += lpdf;
    } else {
      const dev_environment& env = sc.environments[light.environment];
      if (env.tex >= 0) {
        const dev_texture& t  = sc.textures[env.tex];
        vec3f              wl = transform_direction(env.inv_frame, direction);
        vec2f texcoord = {div_(atan2f_(wl.z, wl.x), (2 * pif)), div_(acosf_(clampf(wl.y, -1.0f, 1.0f)), pif)};
        if (texcoord.x < 0) texcoord.x += 1;
        int   i    = clampi((int)(texcoord.x * t.width), 0, t.width - 1);
        int   j    = clampi((int)(texcoord.y * t.height), 0, t.height - 1);
        int   idx  = j * t.width + i;
        float c    = idx == 0 ? __ldg(light.cdf) : __ldg(light.cdf + idx) - __ldg(light.cdf + idx - 1);
        float prob = div_(c, __ldg(light.cdf + light.cdf_size - 1));
        pdf += div_(prob, __ldg(light.row_angle + j));  // the texel's solid angle, a function of the row (init_lights)
      } else {
        pdf += div_(1.0f, (4 * pif));
      }
    }
  }
  pdf *= div_(1.0f, (float)sc.num_lights);  // sample_uniform_pdf
  return pdf;
}

}  // namespace yh

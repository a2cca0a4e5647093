// Ray / primitive tests of the hot path, operation order as in the reference's
// libs/yocto/yocto_math.h (needs --fmad=false: hit ids and distances are
// compared bit-for-bit with the x86-64 oracle).
#pragma once
#include "yh_math.cuh"

namespace yh {

struct ray3f {
  vec3f o, d;
  float tmin, tmax;
};

// intersect_point, yocto_math.h:3405-3423 (approximate: closest approach of the ray to the centre).  uv = {0, 0}.
YH_HD bool intersect_point(const ray3f& ray, const vec3f& p, float r, vec2f& uv, float& dist) {
  vec3f w = p - ray.o;
  float t = div_(dot(w, ray.d), dot(ray.d, ray.d));
  if (t < ray.tmin || t > ray.tmax) return false;
  vec3f rp  = ray.o + ray.d * t;
  vec3f prp = p - rp;
  if (dot(prp, prp) > r * r) return false;
  uv   = {0, 0};
  dist = t;
  return true;
}

// intersect_line, yocto_math.h:3426-3469.  Closest approach of the ray line and
// the segment line; not a true cylinder test.  uv = {s, sqrt(d2)/r}.
YH_HD bool intersect_line(const ray3f& ray, const vec3f& p0, const vec3f& p1, float r0, float r1,
    vec2f& uv, float& dist) {
  vec3f u = ray.d;
  vec3f v = p1 - p0;
  vec3f w = ray.o - p0;
  float a = dot(u, u), b = dot(u, v), c = dot(v, v), d = dot(u, w), e = dot(v, w);
  float det = a * c - b * b;
  if (det == 0) return false;
  float t = div_((b * e - c * d), det);
  float s = div_((a * e - b * d), det);
  if (t < ray.tmin || t > ray.tmax) return false;
  s         = clampf(s, 0.0f, 1.0f);
  vec3f pr  = ray.o + ray.d * t;
  vec3f pl  = p0 + (p1 - p0) * s;
  vec3f prl = pr - pl;
  float d2  = dot(prl, prl);
  float r   = r0 * (1 - s) + r1 * s;
  if (d2 > r * r) return false;
  uv   = {s, div_(sqrt_(d2), r)};
  dist = t;
  return true;
}

// intersect_triangle, yocto_math.h:3472-3505 (Moller-Trumbore, no epsilon)
YH_HD bool intersect_triangle(const ray3f& ray, const vec3f& p0, const vec3f& p1, const vec3f& p2,
    vec2f& uv, float& dist) {
  vec3f edge1 = p1 - p0;
  vec3f edge2 = p2 - p0;
  vec3f pvec  = cross(ray.d, edge2);
  float det   = dot(edge1, pvec);
  if (det == 0) return false;
  float inv_det = div_(1.0f, det);
  vec3f tvec    = ray.o - p0;
  float u       = dot(tvec, pvec) * inv_det;
  if (u < 0 || u > 1) return false;
  vec3f qvec = cross(tvec, edge1);
  float v    = dot(ray.d, qvec) * inv_det;
  if (v < 0 || u + v > 1) return false;
  float t = dot(edge2, qvec) * inv_det;
  if (t < ray.tmin || t > ray.tmax) return false;
  uv   = {u, v};
  dist = t;
  return true;
}

// point_bounds / line_bounds / triangle_bounds, yocto_math.h:3030-3048
YH_HD bbox3f point_bounds(const vec3f& p, float r) { return {min3(p - r, p + r), max3(p - r, p + r)}; }
YH_HD bbox3f line_bounds(const vec3f& p0, const vec3f& p1, float r0, float r1) {
  return {min3(p0 - r0, p1 - r1), max3(p0 + r0, p1 + r1)};
}
YH_HD bbox3f triangle_bounds(const vec3f& p0, const vec3f& p1, const vec3f& p2) {
  return {min3(p0, min3(p1, p2)), max3(p0, max3(p1, p2))};
}

}  // namespace yh

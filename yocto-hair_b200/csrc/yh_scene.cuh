// Device-side scene layout (what lives in HBM) and its host owner.
//
//   shape BVH    : binary LBVH packed into 32-byte SLOTS, one 256-bit load each
//                  (LDG.E.256 on sm_100a).  A slot is either
//                    an inner node: the boxes of BOTH children on the shape's 16-bit
//                      grid (box = grid_origin + q * grid_step, rounded outwards)
//                        w0 = c0.lo.x | c0.hi.x << 16   w1 = c0.lo.y | c0.hi.y << 16
//                        w2 = c0.lo.z | c0.hi.z << 16   w3 = c1.lo.x | c1.hi.x << 16
//                        w4 = c1.lo.y | c1.hi.y << 16   w5 = c1.lo.z | c1.hi.z << 16
//                        w6 = slot of child 0 (child 1 is the next slot)
//                        w7 = bit 31: child 0 is a leaf, bit 30: child 1 is a leaf
//                    or a leaf = ONE primitive reference:
//                        line     : {p0.xyz, r0, p1.xyz, r1}  (the segment itself)
//                        point    : {p.xyz, r, -, -, -, -}
//                        triangle : w0 = element id (vertices in tris[3 * id ..])
//                  Siblings share a 64-byte line, a leaf needs no second fetch, and
//                  the whole tree of 1M segments is 64-106 MB instead of 160 MB.
//                  slot_elem[slot] = element id of a leaf slot.  A long segment may be
//                  referenced by several leaves (each reference bounds a sub-span of
//                  it) — the test always runs on the whole segment, so hit records
//                  do not depend on splitting.
//   tie order    : when two primitives of one shape are hit at EXACTLY the same
//                  t, the reference keeps the one its own BVH visits last
//                  (intersect_line accepts t == tmax, yocto_math.h:3450).  So
//                  the topology of the reference's tree (split_middle,
//                  yocto_pathtrace.cpp:564-595), grown on the device by
//                  yh_tieorder.cu, is kept beside ours: ref_pos[e] =
//                  position of element e in the reference's leaf order, and
//                  tie_nodes[k] = {mid, axis, left, right} of its internal
//                  nodes (child < 0: leaf).  Only ties read them.
//   shading data : the reference's own arrays (positions, normals = hair
//                  tangents, lines / triangles) for eval_position / eval_normal.
#pragma once
#include <mutex>
#include <vector>

#include "yh_common.cuh"
#include "yh_hair.cuh"
#include "yh_intersect.cuh"

namespace yh {

constexpr int kDefaultSplit = 4;  // default yh_bvh_params::max_split
constexpr int kTinyRefs = 4;    // shapes of up to this many leaves are tested without their boxes (light quads)
constexpr int kRefLeafMax = 4;  // the reference's bvh_max_prims (yocto_pathtrace.cpp:598)

struct dev_shape {
  const uint4*  slots;      // 2 uint4 per slot
  const float4* tris;       // triangle shapes: 3 float4 per element
  const int*    slot_elem;  // element id of each leaf slot
  const int*    ref_pos;    // element id -> position in the reference's leaf order
  const int4*   tie_nodes;  // internal nodes of the reference's tree: mid, axis, left, right
  int           num_tie_nodes;
  int           root_leaf;  // the tree is a single leaf (slot 0)
  int           tiny_count;     // 2..kTinyRefs leaves in all: their slots, tested directly by intersect_shape
  int           tiny_slots[4];
  vec3f         grid_origin, grid_step;
  int           num_refs;
  float         scale;      // largest |coordinate| of the shape: boxes are padded by 2e-6 * scale
  int           is_lines;
  int           is_points;  // point shape (intersect_point; shaded with the surface lobes, normal = -outgoing)
  int           has_normals;
  const float*  positions;  // xyz
  const float*  normals;    // xyz (hair tangents for lines)
  const float*  texcoords;  // uv per vertex, or null (eval_texcoord then returns the element uv)
  const int*    elems;      // points: 1 int, lines: 2 ints, triangles: 3 ints
  int           num_elems;
};

struct dev_material {
  vec3f emission;
  vec3f color;
  float opacity;
  int   thin;
  float specular, metallic, roughness, ior, transmission;  // surface lobes (yh_surface.cuh)
  vec3f scattering;                                        // volume behind a refractive surface (!thin && transmission)
  float scanisotropy, trdepth;
  // float RGB textures (index into dev_scene::textures or -1), eval_brdf / eval_emission / eval_vsdf
  int emission_tex, color_tex, specular_tex, metallic_tex, roughness_tex, transmission_tex, opacity_tex, scattering_tex;
  int normal_tex;  // tangent-space normal map (eval_normalmap), triangles only
  int any_tex;  // some slot is set: the shading code then evaluates eval_texcoord
  // hair block (yocto_pathtrace.h:310-316)
  float eumelanin, pheomelanin;
  vec3f sigma_a;
  float beta_m, beta_n, alpha, eta;
  // what eval_hair_brdf derives from the material alone (hair_brdf_material_part) and what the lobes derive from that
  // (hair_consts: the denominators of mp(), the normalisation of the trimmed logistic, ...), filled on the DEVICE by
  // k_material_hair_consts after every upload of the table
  hair_brdf   hair_proto;
  hair_consts hair_k;
};

struct dev_object {
  frame3f frame;
  frame3f inv_frame;  // inverse(frame, non_rigid = true)
  bbox3f  bbox;       // world-space bounds of the shape root (padded)
  int     shape, material;
};

struct dev_texture {
  const float* pixels;  // rgb float
  int          width, height;
};

struct dev_environment {
  frame3f frame;
  frame3f inv_frame;  // inverse(frame) = rigid
  vec3f   emission;
  int     tex;
};

// trace lights, yocto_pathtrace.h:383-387
struct dev_light {
  int          object;       // -1 when this is an environment light
  int          environment;  // -1 when this is an object light
  const float* cdf;
  int          cdf_size;
  // guide table of the cdf: guide[b] = first index with cdf > b * guide_w (b = 0 .. guide_size);
  // narrows the reference's std::upper_bound to a few elements, same result
  const int*   guide;
  int          guide_size;
  float        guide_w, guide_inv_w;
  // textured environments: the solid angle of a texel of row j, sample_lights_pdf's `angle` (:1347-1349), which
  // depends on the row alone - formed once per row with the same operations instead of once per vertex
  const float* row_angle;
};

struct dev_scene {
  // Top-level tree over the objects: the REFERENCE's topology (build_bvh over the objects' world bounds with
  // split_middle, leaves of <= 4 objects, yocto_pathtrace.cpp:597-650, :793-814), so the traversal visits objects in
  // the reference's order and an exact-t tie between two objects goes to the one it would keep.  Node k = two float4:
  //   {box.min.xyz, a}, {box.max.xyz, b}   inner: a = first child (the second follows), b = 0x100 | split axis
  //                                        leaf : a = first entry in scene_prims,       b = object count << 16
  // Boxes are the unions of the objects' padded bounds (conservative).
  const float4*          scene_nodes;
  const int*             scene_prims;
  int                    num_scene_nodes;
  const dev_shape*       shapes;
  const dev_material*    materials;
  const dev_object*      objects;
  const dev_texture*     textures;
  const dev_environment* environments;
  const dev_light*       lights;
  int num_shapes, num_materials, num_objects, num_textures, num_environments, num_lights;
};

// Host-side copy of one shape (kept for light CDFs and BVH set-up).
struct host_shape {
  std::vector<float> positions, normals, radius, texcoords;
  std::vector<int>   points, lines, triangles;
  dev_buffer<float>  d_positions, d_normals, d_radius, d_texcoords;
  dev_buffer<int>    d_elems;
  dev_buffer<uint4>  d_slots;
  dev_buffer<float4> d_prims;  // triangle vertices by element
  dev_buffer<int>    d_slot_elem, d_ref_pos;
  dev_buffer<int4>   d_tie_nodes;
  int                root_leaf = 0, num_refs = 0, num_nodes = 0;
  int                tiny_count = 0, tiny_slots[4] = {0, 0, 0, 0};
  float              scale = 0;
  vec3f              grid_origin{0, 0, 0}, grid_step{1, 1, 1};
  bbox3f             root_bbox{{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
  // one primitive kind per shape; the reference's traversal looks at points, then lines, then triangles (:894-925)
  int  num_elems() const { return is_points() ? (int)points.size() : is_lines() ? (int)lines.size() / 2 : (int)triangles.size() / 3; }
  bool is_points() const { return !points.empty(); }
  bool is_lines() const { return points.empty() && !lines.empty(); }
  int  kind() const { return is_points() ? 2 : is_lines() ? 1 : 0; }  // 0 triangles, 1 lines, 2 points
  const std::vector<int>& elems() const { return is_points() ? points : is_lines() ? lines : triangles; }
  // exact root box of the reference's tree (union of the primitive bounds), from yh_tieorder.cu: the top-level tree
  // over the objects is built from it
  bbox3f ref_root_bbox{{flt_max, flt_max, flt_max}, {-flt_max, -flt_max, -flt_max}};
};

}  // namespace yh

struct yh_scene {
  std::vector<yh::host_shape*>      shapes;
  std::vector<yh::dev_material>     materials;
  std::vector<yh::dev_object>       objects;
  std::vector<yh::dev_environment>  environments;
  std::vector<std::vector<float>>   tex_host;  // rgb float
  std::vector<yh::dev_texture>      textures;
  std::vector<yh::dev_buffer<float>*> tex_dev;
  std::vector<yh::dev_light>        lights;
  std::vector<yh::dev_buffer<float>*> light_cdfs;
  std::vector<yh::dev_buffer<int>*>   light_guides;
  yh::dev_buffer<yh::dev_shape>       d_shapes;
  yh::dev_buffer<yh::dev_material>    d_materials;
  yh::dev_buffer<yh::dev_object>      d_objects;
  yh::dev_buffer<yh::dev_texture>     d_textures;
  yh::dev_buffer<yh::dev_environment> d_environments;
  yh::dev_buffer<yh::dev_light>       d_lights;
  yh::dev_buffer<float4>              d_scene_nodes;  // top-level tree (dev_scene::scene_nodes)
  yh::dev_buffer<int>                 d_scene_prims;
  mutable bool                        scene_tree_pending = true;  // (re)built by sync_scene_tree in init_bvh
  bool bvh_built = false, lights_built = false;
  bool has_volumes = false;  // some material is refractive (!thin && transmission): paths carry a volume state
  int  device = 0;
  yh::bbox3f   world_bbox{{0, 0, 0}, {0, 0, 0}};  // union of the objects' padded bounds (yh_scene_init_bvh)
  yh_bvh_stats stats{};
  yh::dev_scene view() const;
  // (re)builds the top-level tree over the objects after yh_scene_init_bvh
  yh_status sync_scene_tree() const;
  mutable std::mutex tree_mutex;
  ~yh_scene();
};

namespace yh {
// yh_tieorder.cu: fills sh.d_ref_pos / sh.d_tie_nodes / sh.ref_root_bbox (the reference's leaf order, on the device)
yh_status build_tie_order(host_shape& sh);
}

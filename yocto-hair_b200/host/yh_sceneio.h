// Scene files of the hair path as first-class inputs: the reference's scene
// JSON schema (libs/yocto/yocto_sceneio.cpp:1064-1418) and its hair PLY layout
// (vertex {x y z nx ny nz radius} float, line {list uchar int vertex_indices};
// libs/yocto/yocto_ply.h:491-748, :1077-1163) read in bulk.
//
// The reference reads a binary PLY with one fread() per VALUE (yocto_ply.h:690-745): 60 M
// calls for a 5 M-segment head of hair.  load_shape_ply() maps the file once and de-interleaves
// the vertex block and the list records from memory; load_scene() fills the reference's own
// sceneio::model through its public add_* functions, so everything downstream (the app's
// build_trace_scene, the shim, the C ABI) is unchanged.  Both produce exactly what the
// reference's loaders produce (host/yh_loadcheck.cpp compares every field); files outside the
// fast path (ascii / big-endian PLY, non-float vertex properties, OBJ) go through the reference's
// own load_shape.
#pragma once
#include <yocto/yocto_sceneio.h>

#include <string>
#include <vector>

namespace yocto::pathtrace_b200 {

namespace sio = yocto::sceneio;
using yocto::math::vec2f;
using yocto::math::vec2i;
using yocto::math::vec3f;
using yocto::math::vec3i;
using yocto::math::vec4i;

// Bulk reader for binary little-endian PLY shapes; same outputs as shape::load_shape
// (yocto_shape.cpp:4829-4880).  Returns false with `handled` = false when the file is outside
// the fast path (the caller then uses the reference's loader), false with `handled` = true and
// `error` set on a malformed file.
bool load_shape_ply(const std::string& filename, std::vector<int>& points, std::vector<vec2i>& lines,
    std::vector<vec3i>& triangles, std::vector<vec4i>& quads, std::vector<vec3f>& positions,
    std::vector<vec3f>& normals, std::vector<vec2f>& texcoords, std::vector<vec3f>& colors,
    std::vector<float>& radius, std::string& error, bool& handled, bool flip_texcoord = true);

// load_scene for ".json" scenes with the reference's semantics (field defaults, lookat, name ->
// shapes/<name>.{ply,obj}, textures/<name>.{hdr,exr,png,jpg}, instances/<name>.ply, missing
// camera / radius / material fix-ups).  Other scene formats are delegated to sceneio::load_scene.
bool load_scene(const std::string& filename, sio::model* scene, std::string& error,
    sio::progress_callback progress_cb = {});

}  // namespace yocto::pathtrace_b200

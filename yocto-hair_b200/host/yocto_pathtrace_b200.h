// Host shim: the reference's `yocto::pathtrace` entry points of the hair hot path,
// same signatures, running on a B200 through the C ABI (include/yocto_hair_b200.h).
//
// It compiles against the UNMODIFIED reference headers (libs/yocto_pathtrace/
// yocto_pathtrace.h) and works on the reference's own `ptr::scene` / `ptr::state`
// structs, so an application swaps backends by changing a namespace:
//
//     namespace ptr = yocto::pathtrace;         // scene structs, setters: unchanged
//     namespace gpu = yocto::pathtrace_b200;    // this file
//     gpu::init_bvh(scene, params);             // was ptr::init_bvh     (yocto_pathtrace.cpp:755-818)
//     gpu::init_lights(scene, params);          // was ptr::init_lights  (:1695-1740)
//     gpu::init_state(state, scene, camera, params);   // was ptr::init_state (:1931-1946)
//     gpu::trace_samples(state, scene, camera, params);// was ptr::trace_samples (:1992-2026)
//
// Device copies hang off a side table keyed by the scene / state pointer
// (SURVEY.md 8b); release() drops them.  Errors of the C ABI surface as
// std::runtime_error, the exception type the reference's own path throws
// (get_trace_shader_func, yocto_pathtrace.cpp:1669).  There is no CPU fallback.
#pragma once
#include <yocto_pathtrace/yocto_pathtrace.h>

#include <atomic>
#include <functional>
#include <vector>

namespace yocto::pathtrace_b200 {

namespace ptr = yocto::pathtrace;
namespace img = yocto::image;
using math::ray3f;
using math::vec4f;

// BVH knobs of the device builder (0 = library default); set before init_bvh.
struct device_options {
  int device    = 0;  // CUDA device of this process (one process per GPU)
  int max_split = 0;  // references per long segment
  int rank = 0, world = 1;  // image-tile shard rendered by this process
  // Two or more entries: ONE process drives these GPUs (yh_group_*): the scene is replicated on each, the image is
  // sharded over them and summed on the first one over NVLink.  Tiles (default) keep the reference's RNG streams -
  // the image is bit-identical to one GPU and to the CPU reference; sample_shard = true gives every GPU all pixels
  // for its own sample indices (counter-mode streams, full-width wavefront, statistically equivalent image).
  std::vector<int> devices;
  bool             sample_shard = false;
};
void set_device_options(const device_options& options);

// Uploads the scene and builds the GPU LBVH.  Replaces ptr::init_bvh.
void init_bvh(ptr::scene* scene, const ptr::trace_params& params, ptr::progress_callback progress_cb = {});
// Light tables on the device.  Replaces ptr::init_lights (scene->lights stays empty:
// only the tracer reads it).
void init_lights(ptr::scene* scene, const ptr::trace_params& params, ptr::progress_callback progress_cb = {});
// Image size + one PCG32 stream per pixel, on the device.  Replaces ptr::init_state.
// The camera is bound here (the reference's callers re-run init_state after every camera edit).
void init_state(ptr::state* state, const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params);

// One more sample per pixel; state->render = accumulated / samples is valid on return.
void trace_samples(ptr::state* state, const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params);
void trace_samples(ptr::state* state, const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params,
    std::atomic<bool>* stop);
// Batched form: `num_samples` samples per pixel in one device call (the same
// samples, in the same per-pixel order, as num_samples calls above), one download.
void trace_samples(ptr::state* state, const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params,
    int num_samples, std::atomic<bool>* stop = nullptr);

// Upstream convenience (libs/yocto/yocto_trace.h:275): params.samples samples per pixel.
using image_callback = std::function<void(const img::image<vec4f>& render, int current, int total)>;
img::image<vec4f> trace_image(const ptr::scene* scene, const ptr::camera* camera, const ptr::trace_params& params,
    ptr::progress_callback progress_cb = {}, image_callback image_cb = {});

// ptr::intersect_scene_bvh on the device BVH (one ray; use the C ABI's yh_intersect_batch for many).
ptr::intersection3f intersect_scene_bvh(const ptr::scene* scene, const ray3f& ray, bool find_any = false,
    bool non_rigid_frames = true);

// Fills state->pixels (accumulated, samples, rng) from the device, for callers that read them.
void sync_pixels(ptr::state* state);
// Counters of the last trace_samples call on this state.
struct trace_stats { long long samples = 0, rays = 0, light_rays = 0, launches = 0; float device_ms = 0; };
trace_stats last_stats(const ptr::state* state);

// Drops the device copies of a scene / state (also run at process exit).
void release(const ptr::scene* scene);
void release(const ptr::state* state);

}  // namespace yocto::pathtrace_b200

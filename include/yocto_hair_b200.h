/*
 * yocto_hair_b200.h — C ABI of the B200-native hair path-tracing hot path.
 *
 * Drop-in boundary for the per-sample path of dsforza96/yocto-hair
 * (`trace_samples -> trace_sample -> trace_path`, ray/line BVH intersection,
 * pbrt hair BSDF).  The reference has no FFI: its seam is the C++ free-function
 * API of `namespace yocto::pathtrace` / `yocto::extension`.  Every entry point
 * below names the reference function whose body it replaces (file:line are
 * relative to the reference checkout); INTEGRATION.md shows the C++ binding a
 * maintainer adds to the reference tree.
 *
 * Conventions
 *  - plain C, POD structs, host pointers unless a name ends in `_dev`;
 *  - every call returns a yh_status; on failure `yh_last_error()` holds a
 *    message (thread local).  Nothing throws across this boundary;
 *  - all host arrays passed to `yh_scene_create` are copied; the caller may
 *    free them as soon as the call returns (the reference's setters copy their
 *    vectors too, libs/yocto_pathtrace/yocto_pathtrace.cpp:2125-2145);
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    fails with YH_ERR_NO_DEVICE.
 */
#ifndef YOCTO_HAIR_B200_H_
#define YOCTO_HAIR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YH_ABI_VERSION 2

typedef int yh_status;
enum {
  YH_OK              = 0,
  YH_ERR_INVALID     = 1, /* bad argument / inconsistent description        */
  YH_ERR_UNSUPPORTED = 2, /* reference feature outside the hot-path scope   */
  YH_ERR_NO_DEVICE   = 3, /* no usable CUDA device                          */
  YH_ERR_CUDA        = 4, /* a CUDA runtime call failed                     */
  YH_ERR_STATE       = 5  /* call order violated (e.g. trace before bvh)    */
};

/* ---------------------------------------------------------------- scene -- */

/* yocto::pathtrace::shape, libs/yocto_pathtrace/yocto_pathtrace.h:335-366.
 * Lines carry per-vertex tangents in `normals` and per-vertex `radius`.     */
typedef struct yh_shape_desc {
  const float* positions; /* 3 * num_positions                              */
  const float* normals;   /* 3 * num_normals   (0 or num_positions)         */
  const float* texcoords; /* 2 * num_texcoords (0 or num_positions)         */
  const float* radius;    /* num_radius        (0 or num_positions)         */
  const int*   lines;     /* 2 * num_lines                                  */
  const int*   triangles; /* 3 * num_triangles                              */
  int num_positions, num_normals, num_texcoords, num_radius;
  int num_lines, num_triangles;
  /* point shapes (yocto_pathtrace.h:341, intersect_point yocto_math.h:3405): one
   * vertex index per point, per-vertex radius; shaded with the surface lobes and
   * the normal -outgoing (eval_shading_normal :364).  A shape holds ONE kind of
   * primitive (points, lines or triangles).                                   */
  int        num_points;
  const int* points; /* num_points                                           */
} yh_shape_desc;

/* yocto::pathtrace::texture, yocto_pathtrace.h:282-287 (float RGB only).    */
typedef struct yh_texture_desc {
  const float* colorf; /* 3 * width * height, row-major, row 0 on top       */
  int width, height;
} yh_texture_desc;

/* yocto::pathtrace::material, yocto_pathtrace.h:293-329.  Texture slots are
 * indices into yh_scene_desc::textures or -1.  Built on the device: the hair
 * block on line shapes; on triangles emission, color (diffuse), specular,
 * metallic, roughness, ior, transmission (thin sheet, or refraction with
 * thin = 0 and the volume given by scattering / scanisotropy / trdepth) and
 * opacity (eval_brdf / eval_vsdf, yocto_pathtrace.cpp:405-531), each times the
 * float RGB texture in its slot (x channel for the scalar factors, mean for
 * opacity).  transmission_tex scales `transmission`: the reference reads its
 * EMISSION texture for that (:420-421), so callers that want its behaviour
 * pass the emission texture here (the C++ shim does).  Byte (LDR) textures
 * are expanded by the caller: sRGB-decoded for emission / color / scattering,
 * linear for the other slots (normal_tex: tangent-space normal map on
 * triangles, eval_normalmap :326-347).  spectint_tex is never evaluated.
 * On a line shape the hair lobe replaces eval / sample / pdf of every surface
 * lobe (brdf.hair, :474-490), but is_delta (:495) still reads the surface
 * roughness: yh_scene_create refuses (YH_ERR_UNSUPPORTED) a line shape whose
 * material has transmission, or specular / metallic with roughness 0 or a
 * roughness texture (the reference would then run its delta-surface branch on
 * hair); specular / metallic with a plain non-zero roughness are accepted and,
 * as in the reference, have no effect on hair.                                */
typedef struct yh_material_desc {
  float emission[3];
  float color[3];
  float specular, roughness, metallic, ior;
  float spectint[3];
  float transmission;
  float scattering[3];
  float scanisotropy, trdepth, opacity;
  int   thin;
  /* hair block, yocto_pathtrace.h:310-316 */
  float eumelanin, pheomelanin;
  float sigma_a[3];
  float beta_m, beta_n, alpha, eta;
  int emission_tex, color_tex, specular_tex, metallic_tex, roughness_tex;
  int transmission_tex, spectint_tex, scattering_tex, opacity_tex, normal_tex;
} yh_material_desc;

/* yocto::pathtrace::object, yocto_pathtrace.h:369-373.  frame = x,y,z,o.    */
typedef struct yh_object_desc {
  float frame[12];
  int   shape, material;
} yh_object_desc;

/* yocto::pathtrace::environment, yocto_pathtrace.h:376-380.                 */
typedef struct yh_environment_desc {
  float frame[12];
  float emission[3];
  int   emission_tex;
} yh_environment_desc;

/* yocto::pathtrace::camera, yocto_pathtrace.h:272-278.                      */
typedef struct yh_camera_desc {
  float frame[12];
  float lens;
  float film[2];
  float focus;
  float aperture;
} yh_camera_desc;

typedef struct yh_scene_desc {
  const yh_shape_desc*       shapes;
  const yh_texture_desc*     textures;
  const yh_material_desc*    materials;
  const yh_object_desc*      objects;
  const yh_environment_desc* environments;
  int num_shapes, num_textures, num_materials, num_objects, num_environments;
} yh_scene_desc;

/* yocto::pathtrace::trace_params, yocto_pathtrace.h:188-197.                */
enum { YH_SHADER_NAIVE = 0, YH_SHADER_PATH = 1, YH_SHADER_EYELIGHT = 2,
       YH_SHADER_NORMAL = 3 };
typedef struct yh_trace_params {
  int      resolution; /* 720                                               */
  int      shader;     /* YH_SHADER_*: naive, path (default), eyelight, normal */
  int      samples;    /* 512 (informative; the caller drives the loop)     */
  int      bounces;    /* 8                                                 */
  float    clamp;      /* 100                                               */
  uint64_t seed;       /* 961748941                                         */
  int      noparallel; /* ignored on the device                             */
  int      pratio;     /* 8, interactive preview only                       */
} yh_trace_params;

typedef struct yh_scene yh_scene; /* opaque, owns the device copies         */
typedef struct yh_state yh_state; /* opaque, owns accumulation + RNG state  */

/* Library / device ------------------------------------------------------- */
int         yh_abi_version(void);
const char* yh_last_error(void);
yh_status   yh_device_count(int* count);
yh_status   yh_set_device(int device); /* one process per GPU: call first   */
yh_status   yh_default_params(yh_trace_params* params);
yh_status   yh_default_material(yh_material_desc* material);

/* Flattens + uploads a scene.  Replaces the io->ptr copy in
 * apps/yscenetrace/yscenetrace.cpp:49-197 plus the device upload.           */
yh_status yh_scene_create(const yh_scene_desc* desc, yh_scene** scene);
void      yh_scene_destroy(yh_scene* scene);

/* Replaces init_bvh(scene, params), yocto_pathtrace.cpp:755-818 (and the
 * per-shape init_bvh :652-753).  Morton LBVH over line/triangle primitives,
 * built on the device.  `max_split` > 1 lets the builder reference one long
 * segment from up to `max_split` boxes (0 = library default).               */
typedef struct yh_bvh_params {
  int max_split;
  int leaf_size; /* ignored: a leaf holds exactly one reference            */
  int quality;   /* 0 = Morton LBVH on the device (fast build, default);
                    1 = binned-SAH topology built on the host by all cores
                        (slower build, fewer node visits per ray: for long renders) */
  int reserved[5];
} yh_bvh_params;
yh_status yh_scene_init_bvh(yh_scene* scene, const yh_bvh_params* params);

/* Replaces init_lights(scene, params), yocto_pathtrace.cpp:1695-1740.       */
yh_status yh_scene_init_lights(yh_scene* scene);

/* BVH statistics of the built scene (nodes, references, bytes).             */
typedef struct yh_bvh_stats {
  int64_t num_primitives, num_references, num_nodes, node_bytes, prim_bytes;
} yh_bvh_stats;
yh_status yh_scene_bvh_stats(const yh_scene* scene, yh_bvh_stats* stats);

/* Diagnostics: the order in which the REFERENCE's BVH over shape `shape` holds
 * its primitives (shape->bvh->primitives after build_bvh, yocto_pathtrace.cpp:
 * 601-650, :746-752).  The device grows that tree too - only to break exact-t
 * ties the way the reference does.  `order` (may be null) receives *count
 * element indices.                                                           */
yh_status yh_scene_reference_order(const yh_scene* scene, int shape, int* order,
    int64_t* count);

/* Intersection ------------------------------------------------------------
 * yocto::math::ray3f (libs/yocto/yocto_math.h:1106) and
 * yocto::pathtrace::intersection3f (yocto_pathtrace.h:441-447).             */
typedef struct yh_ray {
  float o[3];
  float d[3];
  float tmin, tmax;
} yh_ray;
typedef struct yh_hit {
  int   object, element;
  float uv[2];
  float distance;
  int   hit;
} yh_hit;
/* Batched intersect_scene_bvh(scene, ray, find_any, non_rigid_frames=true),
 * yocto_pathtrace.cpp:934-1053.  Host buffers.                              */
yh_status yh_intersect_batch(const yh_scene* scene, const yh_ray* rays,
    int64_t num_rays, int find_any, yh_hit* hits);
/* Same, buffers already on the device (timing without PCIe).  When `ms` is
 * non-null it receives the kernel time measured with CUDA events.           */
yh_status yh_intersect_batch_dev(const yh_scene* scene, const yh_ray* rays_dev,
    int64_t num_rays, int find_any, yh_hit* hits_dev, float* ms);

/* Rendering ---------------------------------------------------------------
 * yh_state_create replaces init_state, yocto_pathtrace.cpp:1931-1946: image
 * size from the film aspect, one PCG32 stream per pixel.                    */
yh_status yh_state_create(const yh_scene* scene, const yh_camera_desc* camera,
    const yh_trace_params* params, yh_state** state);
void      yh_state_destroy(yh_state* state);
yh_status yh_state_size(const yh_state* state, int* width, int* height);

/* Restrict this state to the pixels of image tiles owned by `rank` of
 * `world` (32x32 tiles, round-robin).  Every pixel keeps its own RNG stream,
 * so the union over ranks is bit-identical to a single-device render.       */
yh_status yh_state_set_shard(yh_state* state, int rank, int world);
/* Sample-range sharding (the alternative north_star names): this state renders
 * EVERY pixel but only the samples with index s = rank, rank + world, ... of each
 * pixel.  Needs per-sample random streams, so it switches the state to COUNTER
 * MODE: sample s of a pixel draws from its own PCG32 stream
 *     make_rng(params.seed, seq_pixel + (s << 31))      (yocto_math.h:1419-1427)
 * where seq_pixel is the reference's per-pixel sequence of init_state (:1943).
 * Sample 0 therefore equals the reference's first sample bit for bit; later
 * samples are statistically equivalent (the reference continues ONE stream per
 * pixel, which makes a pixel's samples strictly sequential).  The per-rank
 * accumulation / count buffers add up (yh_state_buffers_dev + one reduce) to the
 * world = 1 counter-mode image up to fp32 summation order.  world = 1 is valid
 * (counter mode on one device).  Resets the accumulation.                     */
yh_status yh_state_set_sample_shard(yh_state* state, int rank, int world);
/* Counter mode only: `slots_per_pixel` path slots per pixel (0 = as many as fill
 * the device, <= 8).  Samples of one pixel are independent in counter mode, so
 * lane k of a pixel traces the sample ordinals k, k + S, ... concurrently with
 * the other lanes - the wavefront is S times wider, which is what a small image
 * (or a GPU's share of one) needs to fill 148 SMs.  Every lane adds into its own
 * partial image; the partials are folded into the accumulation buffer in lane
 * order after each yh_trace_samples call, so results are deterministic (they
 * equal the one-slot image up to fp32 summation order).  Resets the accumulation. */
yh_status yh_state_set_sample_slots(yh_state* state, int slots_per_pixel);
/* The ownership map behind yh_state_set_shard, host only (needs no device):
 * writes the image pixel indices (j * width + i) owned by `rank` of `world`,
 * in slot order, to `pixels` (may be null) and their number to `count`.
 * Replaces the row scheduler of parallel_for, yocto_pathtrace.cpp:1954-1989. */
yh_status yh_shard_pixels(int width, int height, int rank, int world,
    int* pixels, int64_t* count);

/* Replaces `num_samples` consecutive calls of trace_samples(state, scene,
 * camera, params), yocto_pathtrace.cpp:1992-2026: every owned pixel gains
 * exactly `num_samples` samples, drawn in the reference's per-pixel order.
 * `stop` may be null; it is polled between wavefront iterations
 * (reference: std::atomic<bool>* stop, :2009-2026).  Once it reads non-zero no
 * pixel starts another sample; the samples already in flight are finished, so
 * every pixel still holds WHOLE samples of its stream (per-pixel counts may then
 * differ by the samples not started: read them with yh_state_download) and the
 * state can be traced further.  yh_trace_stats::samples reports what was traced. */
yh_status yh_trace_samples(yh_state* state, int num_samples,
    const volatile int* stop);

/* Counters of the last yh_trace_samples call.                               */
typedef struct yh_trace_stats {
  int64_t samples;      /* trace_sample calls                               */
  int64_t rays;         /* intersect_scene_bvh calls issued by trace_path   */
  int64_t light_rays;   /* intersect_instance_bvh calls (sample_lights_pdf) */
  int64_t iterations;   /* wavefront iterations                             */
  int64_t launches;     /* kernels launched                                 */
  float   device_ms;    /* CUDA-event time of the whole call                */
  float   traverse_ms;  /* CUDA-event time spent in the traversal kernel    */
  float   shade_ms;     /* CUDA-event time spent in the shading kernels     */
  int64_t nodes_visited, prims_tested; /* only in a stats build (else 0)    */
} yh_trace_stats;
yh_status yh_state_stats(const yh_state* state, yh_trace_stats* stats);
/* When enabled, yh_trace_samples brackets the traversal and the shading
 * kernels of every iteration with CUDA events on the state's stream and fills
 * traverse_ms / shade_ms (also enabled by YH_PROFILE_KERNELS=1).            */
yh_status yh_state_set_profiling(yh_state* state, int enabled);

/* Copies results to the host.  Any pointer may be null.
 *  render : 4 floats/pixel = accumulated / samples  (state->render)
 *  accum  : 4 floats/pixel = pixel.accumulated
 *  samples: 1 int/pixel    = pixel.samples
 * Row-major, row 0 on top (img::image, libs/yocto/yocto_image.h:613-615).   */
yh_status yh_state_download(const yh_state* state, float* render, float* accum,
    int* samples);
/* Device pointers of the accumulation (float4/pixel) and sample-count
 * (int/pixel) buffers, for a collective over NVLink.                        */
yh_status yh_state_buffers_dev(yh_state* state, float** accum_dev,
    int** samples_dev);
/* render = accumulated / samples (:1688) on device buffers that did not come
 * out of one state - the sum a collective left on the root rank.  Synchronous. */
yh_status yh_resolve_dev(const float* accum_dev, const int* samples_dev,
    int64_t num_pixels, float* render_dev);
/* Dump / restore of the per-pixel RNG (2 x uint64 per pixel: state, inc).   */
yh_status yh_state_download_rng(const yh_state* state, uint64_t* rng);

/* Several GPUs of one node in ONE process ---------------------------------
 * Replaces, across the GPUs of a box, what parallel_for does across host
 * threads for the reference (yocto_pathtrace.cpp:1954-1989): the scene is
 * replicated in every GPU's HBM, each GPU renders its shard with its own state,
 * and yh_group_download sums the fp32 accumulation / count buffers on the first
 * GPU and resolves the image in ONE kernel that loads the peers' buffers over
 * NVLink (peer access; copies without it), or with ncclReduce (YH_REDUCE_NCCL:
 * libnccl.so.2 is loaded on demand, communicators from ncclCommInitAll).
 *   yh_group_create      = yh_scene_create + yh_scene_init_bvh + yh_scene_init_lights
 *                          on every device (devices == NULL: 0 .. num_devices - 1; a
 *                          device listed k times carries k ranks), side by side
 *   yh_group_init_state  = init_state (:1931-1946) on every device + the shard:
 *                          YH_SHARD_TILES   yh_state_set_shard: image tiles, the reference's
 *                                           RNG streams; the summed image is bit-identical
 *                                           to one GPU and to the reference
 *                          YH_SHARD_SAMPLES yh_state_set_sample_shard: every GPU renders all
 *                                           pixels for its sample indices (counter mode)
 *   yh_group_trace_samples  `num_samples` more samples per pixel over the whole group
 *                          (a multiple of the device count for YH_SHARD_SAMPLES)
 *   yh_group_download    totals over the group: render / accum / samples as yh_state_download
 * The calling thread's current device is preserved.                          */
typedef struct yh_group yh_group;
enum { YH_SHARD_TILES = 0, YH_SHARD_SAMPLES = 1 };
enum { YH_REDUCE_PEER = 0, YH_REDUCE_NCCL = 1 };
yh_status yh_group_create(const yh_scene_desc* desc, const yh_bvh_params* bvh,
    const int* devices, int num_devices, yh_group** group);
void      yh_group_destroy(yh_group* group);
yh_status yh_group_num_devices(const yh_group* group, int* num_devices);
/* The replica of rank `rank` and the device it lives on (make it current with
 * yh_set_device before handing the scene to yh_intersect_batch).             */
yh_status yh_group_scene(const yh_group* group, int rank, const yh_scene** scene, int* device);
yh_status yh_group_init_state(yh_group* group, const yh_camera_desc* camera,
    const yh_trace_params* params, int shard);
yh_status yh_group_set_reduce(yh_group* group, int mode);
yh_status yh_group_size(const yh_group* group, int* width, int* height);
yh_status yh_group_trace_samples(yh_group* group, int num_samples,
    const volatile int* stop);
/* Summed counters of the last yh_group_trace_samples (device_ms = the slowest
 * device) and the CUDA-event time of the last reduce + resolve.              */
yh_status yh_group_stats(const yh_group* group, yh_trace_stats* stats, float* reduce_ms);
yh_status yh_group_download(yh_group* group, float* render, float* accum,
    int* samples);

/* Hair BSDF, batched (test + measurement entry points) -------------------- */

/* yocto::extension::hair_material, libs/yocto_extension/yocto_extension.h:86-95 */
typedef struct yh_hair_material {
  float sigma_a[3];
  float beta_m, beta_n, alpha, eta;
  float color[3];
  float eumelanin, pheomelanin;
} yh_hair_material;

/* yocto::extension::hair_brdf, yocto_extension.h:97-113 (30 floats).        */
typedef struct yh_hair_brdf {
  float sigma_a[3];
  float alpha, eta, h;
  float v[4];
  float s;
  float sin_2k_alpha[3];
  float cos_2k_alpha[3];
  float gamma_o;
  float world_to_brdf[12]; /* frame3f x,y,z,o */
} yh_hair_brdf;

/* eval_hair_brdf(material, v, normal, tangent), yocto_extension.cpp:127-177 */
yh_status yh_hair_brdf_batch(const yh_hair_material* materials, const float* v,
    const float* normals, const float* tangents, int64_t n, yh_hair_brdf* brdfs);
/* eval_hair_scattering(brdf, outgoing, incoming), yocto_extension.cpp:255-336 */
yh_status yh_hair_eval_batch(const yh_hair_brdf* brdfs, const float* outgoing,
    const float* incoming, int64_t n, float* f_rgb);
/* sample_hair_scattering(brdf, outgoing, rn), yocto_extension.cpp:399-479   */
yh_status yh_hair_sample_batch(const yh_hair_brdf* brdfs, const float* outgoing,
    const float* rn, int64_t n, float* incoming);
/* sample_hair_scattering_pdf(brdf, outgoing, incoming), yocto_extension.cpp:481-551 */
yh_status yh_hair_pdf_batch(const yh_hair_brdf* brdfs, const float* outgoing,
    const float* incoming, int64_t n, float* pdf);
/* Fused eval + pdf + sample on device-resident tuples; `ms` = kernel time.
 * Inputs: brdfs_dev[n], wo_dev[3n], wi_dev[3n], rn_dev[2n].
 * Outputs: f_dev[3n], pdf_dev[n], ws_dev[3n].                               */
yh_status yh_hair_fused_dev(const yh_hair_brdf* brdfs_dev, const float* wo_dev,
    const float* wi_dev, const float* rn_dev, int64_t n, float* f_dev,
    float* pdf_dev, float* ws_dev, float* ms);

/* The four statistical self-tests of yocto_extension.cpp:555-693 run on the
 * device BSDF.  `which`: 0 white_furnace, 1 white_furnace_sampled,
 * 2 sampling_weights, 3 sampling_consistency.  `worst` receives the value
 * farthest from the pass band.  Returns YH_OK and *passed = 1 on success.    */
yh_status yh_hair_selftest(int which, int* passed, float* worst);

/* PCG32: first `n` rand1f of make_rng(seed, seq), yocto_math.h:1396-1442.    */
yh_status yh_rng_floats(uint64_t seed, uint64_t seq, int n, float* out);

/* libm replicas evaluated on the device (parity diagnostics).
 * fn: 0 expf 1 logf 2 sinf 3 cosf 4 atan2f(x=a,y=b) 5 acosf 6 asin(double)->float
 *     7 (float)(1 / (sinh(1/a) * 2a)) in double  8 fmodf(a, b)
 *     9 / 10 / 11: the double sinh(1/a) * 2a of mp() as an exact hi / mid / lo float expansion */
yh_status yh_libm_batch(int fn, const float* a, const float* b, int64_t n,
    float* out);

/* Raw device memory helpers for harnesses that have no CUDA binding.        */
yh_status yh_dev_alloc(void** ptr, int64_t bytes);
yh_status yh_dev_free(void* ptr);
yh_status yh_dev_upload(void* dst_dev, const void* src_host, int64_t bytes);
yh_status yh_dev_download(void* dst_host, const void* src_dev, int64_t bytes);
yh_status yh_dev_synchronize(void);

#ifdef __cplusplus
}
#endif
#endif /* YOCTO_HAIR_B200_H_ */

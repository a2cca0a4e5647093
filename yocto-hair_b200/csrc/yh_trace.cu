// Wavefront path integrator: replaces trace_samples -> trace_sample ->
// trace_path (yocto_pathtrace.cpp:1380-1511, :1676-1689, :1992-2026).
//
// One path "slot" per owned pixel.  The reference keeps ONE PCG32 stream per
// pixel and consumes it sample after sample, so a pixel's samples are strictly
// sequential; the parallelism is across pixels.  A slot that finishes a sample
// immediately regenerates the next camera ray for its pixel (until the call's
// quota is met), so every lane stays busy until the tail of the call.
//
// Per iteration:
//   k_intersect        persistent warps fetch the active slots' rays: closest hit, the hit record goes to the slot and
//                      the slot's bit is set in the bitmap of the hit's kind (miss / surface / hair).
//   k_enumerate_kinds  bitmaps -> the three shading queues, in slot order inside runs of 1024 slots.
//   k_shade_miss       environment radiance, finish sample, regenerate.
//   k_shade<false>     surfaces and points (all material lobes, volumes with VOL).
//   k_shade<true>      hair BSDF: sample / eval / pdf, MIS with the lights.
// Each shading kernel appends surviving or regenerated slots to the next
// iteration's active queue (one warp-aggregated atomic per warp).
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "yh_shade.cuh"

namespace yh {

struct dev_camera {
  frame3f frame;
  float   lens;
  vec2f   film;
  float   focus, aperture;
};

// Path state, structure of arrays indexed by slot.
struct path_state {
  float4*             ray_o;     // o.xyz, -
  float4*             ray_d;     // d.xyz, -
  float4*             weight;    // weight.xyz, hit alpha
  float4*             radiance;  // radiance.xyz, bounce (int bits)
  int4*               hit;       // object, element, u bits, v bits
  unsigned long long* rng_state;
  unsigned long long* rng_inc;
  unsigned long long* rng_base;   // increment of the pixel's reference stream (counter mode re-seeds from it)
  int*                remaining;  // samples still to start for this slot in this call
  int*                pixel;      // image pixel index j * width + i
  float4*             accum;      // image layout, pixel.accumulated
  int*                samples;    // image layout, pixel.samples
  // only when the scene has a refractive material (else null): the one-entry volume stack of trace_path
  float4*             vol_density;  // vsdf.density.xyz, vsdf.anisotropy
  float4*             vol_scatter;  // vsdf.scatter.xyz, 1 while the path is inside the volume (0: stack empty)
  float*              hit_dist;     // intersection.distance of the current vertex
  // only with several slots per pixel (counter mode, yh_state_set_sample_slots; else null): lane k of a pixel draws the
  // sample ordinals k, k + S, ... and adds into its own partial image, folded into accum in lane order after the call
  int*                ordinal;          // next sample ordinal of the slot
  float4*             partial;          // S x image
  int*                partial_samples;  // S x image
  int                 num_pixels;
};
// pixel[slot] = image pixel | lane << 24
constexpr int kPixelMask = 0x00ffffff;

struct queues {
  int* active[2];  // slot ids
  int* miss;
  int* surface;
  int* hair;
  // counts: [0] active cur, [1] active next, [2] miss, [3] surface, [4] hair
  int* count;
  // The traversal does not append to the three shading queues: it sets bit `slot` in the bitmap of the hit's kind
  // (kind_bits + kind * kind_words), and k_enumerate_kinds turns the bitmaps into the queues, 1024 slots at a time.
  // The queues are then sorted by slot inside every run of up to 1024 entries, so the lanes of a shading warp read
  // and write NEIGHBOURING path-state records (a queue in completion order scatters them: one 16-byte record per
  // 32-byte sector).  Measured with the hair queue sorted by a radix sort: k_shade<hair> -15 % on C3.
  unsigned* kind_bits;
  int       kind_words;  // words per bitmap, a multiple of 32
  // statistics: [0] rays, [1] light rays, [2] samples, [3] nodes, [4] prims
  unsigned long long* stats;
};

struct trace_consts {
  int   width, height, bounces;
  float clamp;
  // counter mode (yh_state_set_sample_shard): sample s of a pixel draws from its own PCG32
  // stream make_rng(seed, seq_pixel + (s << 31)); this state renders s = srank, srank + sworld, ...
  int                counter, srank, sworld;
  int                slots;  // slots per pixel (1 unless yh_state_set_sample_slots)
  unsigned long long seed;
  int                shader;  // YH_SHADER_*: trace_naive / trace_path / trace_eyelight / trace_normal (:1380-1657)
};

// Path state is touched once per wavefront iteration and is larger than L2
// (104 B per pixel): stream it (evict-first) so the BVH stays L2-resident.
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }
template <typename T>
__device__ __forceinline__ void st_stream(T* p, decltype(T{}) v) { __stcs(p, v); }

// warp-aggregated queue append
__device__ __forceinline__ void queue_push(int* queue, int* counter, bool pred, int value) {
  unsigned mask = __ballot_sync(__activemask(), pred);
  if (!pred) return;
  int lane   = threadIdx.x & 31;
  int leader = __ffs(mask) - 1;
  int base   = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(mask));
  base = __shfl_sync(mask, base, leader);
  queue[base + __popc(mask & ((1u << lane) - 1))] = value;
}

// sample_camera / eval_camera, yocto_pathtrace.cpp:211-229.  GCC evaluates the
// two rand2f arguments of the call at :1680 right to left: the LENS sample is
// drawn first, the pixel sample second (verified against the oracle build).
__device__ __forceinline__ ray3f camera_ray(const dev_camera& cam, const trace_consts& tc, int pixel, rng_state& rng) {
  vec2f luv = rand2f(rng);
  vec2f puv = rand2f(rng);
  int   i = pixel % tc.width, j = pixel / tc.width;
  vec2f uv = {div_((float)i + puv.x, (float)tc.width), div_((float)j + puv.y, (float)tc.height)};
  vec3f q  = {cam.film.x * (0.5f - uv.x), cam.film.y * (uv.y - 0.5f), cam.lens};
  vec3f dc = -normalize(q);
  vec3f e  = {0, 0, 0};
  if (cam.aperture != 0) {  // sample_disk, yocto_math.h:4897-4901
    float r = sqrt_(luv.y), phi = 2 * pif * luv.x;
    e = {div_pow2(cosf_(phi) * r * cam.aperture, 0.5f), div_pow2(sinf_(phi) * r * cam.aperture, 0.5f), 0};
  }
  vec3f p = dc * cam.focus / fabs_(dc.z);
  vec3f d = normalize(p - e);
  return {transform_point(cam.frame, e), transform_direction(cam.frame, d), ray_eps, flt_max};
}

__device__ __forceinline__ void start_sample(const path_state& ps, const dev_camera& cam, const trace_consts& tc,
    int slot, rng_state& rng, int samples_done) {
  if (tc.counter) {
    // make_rng(seed, seq) (yocto_math.h:1419-1427) in closed form: inc = (seq << 1) | 1,
    // state = (inc + seed) * mult + inc; seq = seq_pixel + (s << 31) shifts inc by s << 32
    if (ps.ordinal) {  // this slot's own ordinals: k, k + S, ...
      samples_done = ld_stream(&ps.ordinal[slot]);
      st_stream(&ps.ordinal[slot], samples_done + tc.slots);
    }
    unsigned long long sidx = (unsigned long long)tc.srank + (unsigned long long)tc.sworld * (unsigned long long)samples_done;
    rng.inc   = ld_stream(&ps.rng_base[slot]) + (sidx << 32);
    rng.state = (rng.inc + tc.seed) * 6364136223846793005ULL + rng.inc;
    st_stream(&ps.rng_inc[slot], rng.inc);
  }
  ray3f ray       = camera_ray(cam, tc, ld_stream(&ps.pixel[slot]) & kPixelMask, rng);
  st_stream(&ps.ray_o[slot], make_float4(ray.o.x, ray.o.y, ray.o.z, 0));
  st_stream(&ps.ray_d[slot], make_float4(ray.d.x, ray.d.y, ray.d.z, 0));
  st_stream(&ps.weight[slot], make_float4(1, 1, 1, 0));
  st_stream(&ps.radiance[slot], make_float4(0, 0, 0, __int_as_float(0)));
  if (ps.vol_scatter) st_stream(&ps.vol_scatter[slot], make_float4(0, 0, 0, 0));  // a new path starts outside every volume
}

// trace_sample tail, :1683-1688, then regenerate or retire the slot.
__device__ __forceinline__ bool finish_sample(const path_state& ps, const dev_camera& cam, const trace_consts& tc,
    int slot, vec3f radiance, float alpha, rng_state& rng) {
  if (!isfinite3(radiance)) radiance = {0, 0, 0};
  float m = max3(radiance);
  if (m > tc.clamp) radiance = radiance * div_(tc.clamp, m);
  int     px      = ld_stream(&ps.pixel[slot]);
  float4* accum   = ps.accum;
  int*    samples = ps.samples;
  if (ps.partial) {  // lane (px >> 24) of the pixel owns one image of the partial stack
    size_t lane = (size_t)(px >> 24) * (size_t)ps.num_pixels;
    accum = ps.partial + lane, samples = ps.partial_samples + lane, px &= kPixelMask;
  }
  float4 a  = ld_stream(&accum[px]);
  a.x += radiance.x, a.y += radiance.y, a.z += radiance.z, a.w += alpha;
  st_stream(&accum[px], a);
  int done = ld_stream(&samples[px]) + 1;
  st_stream(&samples[px], done);
  int left = ld_stream(&ps.remaining[slot]);
  if (left > 0) {
    st_stream(&ps.remaining[slot], left - 1);
    start_sample(ps, cam, tc, slot, rng, done);
    return true;
  }
  return false;
}

__global__ void __launch_bounds__(256) k_start(path_state ps, queues q, dev_camera cam, trace_consts tc,
    int num_slots, int num_samples, int ordinal_base) {
  int  slot = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = slot < num_slots && num_samples > 0;
  if (live && ps.ordinal) {  // lane k of a pixel takes the ordinals base + k, base + k + S, ... < base + num_samples
    int lane    = ld_stream(&ps.pixel[slot]) >> 24;
    num_samples = num_samples > lane ? (num_samples - lane + tc.slots - 1) / tc.slots : 0;
    st_stream(&ps.ordinal[slot], ordinal_base + lane);
    live = num_samples > 0;
  }
  if (live) {
    rng_state rng = {ld_stream(&ps.rng_state[slot]), ld_stream(&ps.rng_inc[slot])};
    if (tc.bounces <= 0 && (tc.shader == YH_SHADER_PATH || tc.shader == YH_SHADER_NAIVE)) {
      // the path loop (:1395, :1524) never runs: black sample, alpha 0, camera draws still consumed
      int  px = ld_stream(&ps.pixel[slot]) & kPixelMask;
      int* samples = ps.partial ? ps.partial_samples + (size_t)(ld_stream(&ps.pixel[slot]) >> 24) * (size_t)ps.num_pixels : ps.samples;
      int  done = ld_stream(&samples[px]);
      for (int s = 0; s < num_samples; s++) {
        if (tc.counter) {  // same re-seed as start_sample
          unsigned long long ord = ps.ordinal ? (unsigned long long)(ordinal_base + (ld_stream(&ps.pixel[slot]) >> 24) + s * tc.slots)
                                              : (unsigned long long)(done + s);
          unsigned long long sidx = (unsigned long long)tc.srank + (unsigned long long)tc.sworld * ord;
          rng.inc   = ld_stream(&ps.rng_base[slot]) + (sidx << 32);
          rng.state = (rng.inc + tc.seed) * 6364136223846793005ULL + rng.inc;
          st_stream(&ps.rng_inc[slot], rng.inc);
        }
        camera_ray(cam, tc, px, rng);
      }
      st_stream(&samples[px], done + num_samples);
      live = false;
    } else {
      st_stream(&ps.remaining[slot], num_samples - 1);
      start_sample(ps, cam, tc, slot, rng, ld_stream(&ps.samples[ld_stream(&ps.pixel[slot]) & kPixelMask]));
    }
    st_stream(&ps.rng_state[slot], rng.state);
  }
  queue_push(q.active[0], q.count + 0, live, slot);
}

// Ray source / hit sink of the wavefront for trace_persistent().
struct wavefront_io {  // by value: references to kernel parameters would force them into local memory
  const dev_scene  sc;
  const path_state ps;
  const queues     q;
  const int*       active;
  __device__ __forceinline__ bool load(int index, int& slot, ray3f& r) const {
    slot     = active[index];
    float4 o = ld_stream(&ps.ray_o[slot]), d = ld_stream(&ps.ray_d[slot]);
    r        = {{o.x, o.y, o.z}, {d.x, d.y, d.z}, ray_eps, flt_max};
    return true;
  }
  __device__ __forceinline__ void reload(int slot, vec3f& o, vec3f& d) const {
    float4 o4 = ld_stream(&ps.ray_o[slot]), d4 = ld_stream(&ps.ray_d[slot]);
    o = {o4.x, o4.y, o4.z}, d = {d4.x, d4.y, d4.z};
  }
  __device__ __forceinline__ void store(int slot, const hit_record& h) const {
    st_stream(&ps.hit[slot], make_int4(h.object, h.element, __float_as_int(h.uv.x), __float_as_int(h.uv.y)));
    if (ps.hit_dist) st_stream(&ps.hit_dist[slot], h.distance);
    int kind     = !h.hit ? 0 : (sc.shapes[sc.objects[h.object].shape].is_lines ? 2 : 1);
    atomicOr(q.kind_bits + (size_t)kind * q.kind_words + (slot >> 5), 1u << (slot & 31));  // result unused: a RED
  }
};

// Closest hit for every slot of the active queue, then the slot id goes to the
// miss / surface / hair queue.  Persistent warps, q.count[5] is the ray cursor.
#ifndef YH_INTERSECT_BLOCKS
#define YH_INTERSECT_BLOCKS 7
#endif
template <bool CAN_SUSPEND>
__global__ void __launch_bounds__(kTraceBlock, YH_INTERSECT_BLOCKS) k_intersect(dev_scene sc, path_state ps, queues q, int cur,
    const int* active, suspend_io sus) {
  __shared__ int smem_stack[kStackShared * kTraceBlock];
  int            n = q.count[cur];
  if (CAN_SUSPEND) {  // rays carried over: read this launch's share, write the next one's (counts [6 + cur], [7 - cur])
    sus.n_in      = min(q.count[6 + cur], sus.capacity);  // (reservations past the end of the buffer hold nothing)
    sus.out_count = q.count + 6 + (cur ^ 1);
  }
  unsigned long long nv = 0, pt = 0;
  __shared__ int smem_park[kParkSlots * kTraceBlock];
  __shared__ int smem_cold[kColdWords * kTraceBlock];
  wavefront_io   io{sc, ps, q, active};
  trace_persistent<CAN_SUSPEND>(sc, io, n, q.count + 5, smem_stack + threadIdx.x, smem_park + threadIdx.x, smem_cold + threadIdx.x, false,
      &nv, &pt, &sus);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(q.stats + 0, (unsigned long long)n);
#ifdef YH_TRAVERSAL_STATS
  atomicAdd(q.stats + 3, nv);
  atomicAdd(q.stats + 4, pt);
#endif
}

// Bitmaps of the traversal -> shading queues (queues::kind_bits).  One warp per run of 32 words (1024 slots) and kind
// (blockIdx.y): lane l owns word l, the warp reserves the run's entries with one atomic and every lane writes the
// slots of its word in ascending order behind those of the lanes before it; the words are cleared on the way, so the
// bitmaps are empty again when the round's shading kernels start.
__global__ void __launch_bounds__(128) k_enumerate_kinds(queues q) {
  const int kind = blockIdx.y, lane = threadIdx.x & 31;
  unsigned* bits = q.kind_bits + (size_t)kind * q.kind_words;
  int*      queue = kind == 0 ? q.miss : (kind == 1 ? q.surface : q.hair);
  const int runs = q.kind_words >> 5;
  for (int run = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); run < runs; run += gridDim.x * (blockDim.x >> 5)) {
    unsigned* wp   = bits + run * 32 + lane;
    unsigned  word = *wp;
    if (word) *wp = 0;
    const int c = __popc(word);
    int       inc = c;
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    if (total == 0) continue;
    int base = 0;
    if (lane == 31) base = atomicAdd(q.count + 2 + kind, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    int out = base + inc - c, first = (run * 32 + lane) * 32;
    while (word) {
      queue[out++] = first + __ffs(word) - 1;
      word &= word - 1;
    }
  }
}

// miss: radiance += weight * eval_environment, break (:1398-1401)
__global__ void __launch_bounds__(256) k_shade_miss(dev_scene sc, path_state ps, queues q, dev_camera cam,
    trace_consts tc, int next) {
  int n = q.count[2];
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    int  idx  = base + threadIdx.x;
    bool live = idx < n, again = false;
    int  slot = live ? q.miss[idx] : 0;
    if (live) {
      float4    d = ld_stream(&ps.ray_d[slot]), w = ld_stream(&ps.weight[slot]), r = ld_stream(&ps.radiance[slot]);
      vec3f     radiance = vec3f{r.x, r.y, r.z} + vec3f{w.x, w.y, w.z} * eval_environment(sc, {d.x, d.y, d.z});
      rng_state rng      = {ld_stream(&ps.rng_state[slot]), ld_stream(&ps.rng_inc[slot])};
      again              = finish_sample(ps, cam, tc, slot, radiance, tc.shader == YH_SHADER_NORMAL ? 1.0f : w.w, rng);
      st_stream(&ps.rng_state[slot], rng.state);
    }
    queue_push(q.active[next], q.count + next, again, slot);
  }
}

// The body of trace_path for one surface vertex, :1417-1507.  HAIR selects the
// lobe set at compile time so the two kernels stay divergence-free inside.
template <bool HAIR, int SHADER, bool VOL>
__device__ __forceinline__ bool shade_vertex(const dev_scene& sc, const path_state& ps, const dev_camera& cam,
    const trace_consts& tc, int slot, trav_stack& st, unsigned int& light_rays) {
  float4    o4 = ld_stream(&ps.ray_o[slot]), d4 = ld_stream(&ps.ray_d[slot]), w4 = ld_stream(&ps.weight[slot]), r4 = ld_stream(&ps.radiance[slot]);
  int4      h4 = ld_stream(&ps.hit[slot]);
  rng_state rng = {ld_stream(&ps.rng_state[slot]), ld_stream(&ps.rng_inc[slot])};
  vec3f     ray_d = {d4.x, d4.y, d4.z}, weight = {w4.x, w4.y, w4.z}, radiance = {r4.x, r4.y, r4.z};
  int       bounce = __float_as_int(r4.w);
  int       object = h4.x, element = h4.y;
  vec2f     uv = {__int_as_float(h4.z), __int_as_float(h4.w)};
  const dev_object&   ob  = sc.objects[object];
  const dev_material& mat = sc.materials[ob.material];

  vec3f outgoing = -ray_d;
  // Inside a volume (trace_path only, :1403-1414): sample a distance along the ray against the transmittance; a
  // distance short of the surface makes this vertex a scattering event in the medium (:1470-1497).
  vsdf vol;
  bool has_vol = false;
  if (VOL) {
    float4 vs4 = ld_stream(&ps.vol_scatter[slot]);
    has_vol    = vs4.w != 0;
    if (has_vol) {
      float4 vd4 = ld_stream(&ps.vol_density[slot]);
      vol        = {{vd4.x, vd4.y, vd4.z}, {vs4.x, vs4.y, vs4.z}, vd4.w};
      float hit_distance = ld_stream(&ps.hit_dist[slot]);
      float rd = rand1f(rng);  // arguments are evaluated right to left in the oracle build: rd, then rl
      float rl = rand1f(rng);
      float distance = sample_transmittance(vol.density, hit_distance, rl, rd);
      weight = weight * (eval_transmittance(vol.density, distance) / sample_transmittance_pdf(vol.density, distance, hit_distance));
      if (distance < hit_distance) {
        vec3f ray_o = {o4.x, o4.y, o4.z};
        vec3f vpos  = ray_o + ray_d * distance;
        vec3f vin;
        if (rand1f(rng) < 0.5f) {
          vec2f rn  = rand2f(rng);
          float rnl = rand1f(rng);
          (void)rnl;
          vin = sample_scattering(vol, outgoing, rn);
        } else {
          vec2f ruv = rand2f(rng);
          float rel = rand1f(rng);
          float rl2 = rand1f(rng);
          vin       = sample_lights(sc, vpos, rl2, rel, ruv);
        }
        float light_pdf = sample_lights_pdf(sc, vpos, vin, st, light_rays);
        weight = weight * (eval_scattering(vol, outgoing, vin) / (0.5f * sample_scattering_pdf(vol, outgoing, vin) + 0.5f * light_pdf));
        st_stream(&ps.ray_o[slot], make_float4(vpos.x, vpos.y, vpos.z, 0));
        st_stream(&ps.ray_d[slot], make_float4(vin.x, vin.y, vin.z, 0));
        bool vfinished = false;
        if ((weight == vec3f{0, 0, 0}) || !isfinite3(weight)) vfinished = true;
        if (!vfinished && bounce > 3) {
          float rr_prob = fmin_(0.99f, max3(weight));
          if (rand1f(rng) >= rr_prob) vfinished = true;
          else weight = weight * div_(1.0f, rr_prob);
        }
        bounce += 1;
        if (bounce >= tc.bounces) vfinished = true;
        bool vagain;
        if (vfinished) {
          vagain = finish_sample(ps, cam, tc, slot, radiance, 1.0f, rng);
        } else {
          st_stream(&ps.weight[slot], make_float4(weight.x, weight.y, weight.z, 1.0f));
          st_stream(&ps.radiance[slot], make_float4(radiance.x, radiance.y, radiance.z, __int_as_float(bounce)));
          vagain = true;
        }
        st_stream(&ps.rng_state[slot], rng.state);
        return vagain;
      }
    }
  }
  vec3f position = eval_position(sc, object, element, uv);
  // eval_shading_normal, :350-369
  vec3f normal, tangent;
  if (HAIR) {
    tangent = eval_normal(sc, object, element, uv);
    normal  = orthonormalize(outgoing, tangent);
  } else if (sc.shapes[ob.shape].is_points) {
    normal = -outgoing;  // :364-365
  } else {
    normal = mat.normal_tex >= 0 ? eval_normalmap(sc, object, element, uv, mat.normal_tex) : eval_normal(sc, object, element, uv);
    if (mat.thin) normal = dot(normal, outgoing) >= 0 ? normal : -normal;
  }
  // material factors x textures, eval_emission :397-402 and eval_brdf :409-424 (a missing texture evaluates to
  // {1,1,1}, and x * 1 is exact, so untextured materials take the cheap branch).  The transmission factor comes from
  // the slot the caller put in transmission_tex - the reference reads its EMISSION texture there (:420-421).
  vec3f emission = mat.emission, base = mat.color, scattering = mat.scattering;
  float specular = mat.specular, metallic = mat.metallic, roughness = mat.roughness, transmission = mat.transmission;
  float opacity_mean = div_(1.0f + 1.0f + 1.0f, 3.0f);
  if (mat.any_tex) {
    vec2f texcoord = eval_texcoord(sc, object, element, uv);
    emission     = mat.emission * eval_texture(sc, mat.emission_tex, texcoord);
    base         = mat.color * eval_texture(sc, mat.color_tex, texcoord);
    specular     = mat.specular * eval_texture(sc, mat.specular_tex, texcoord).x;
    metallic     = mat.metallic * eval_texture(sc, mat.metallic_tex, texcoord).x;
    roughness    = mat.roughness * eval_texture(sc, mat.roughness_tex, texcoord).x;
    transmission = mat.transmission * eval_texture(sc, mat.transmission_tex, texcoord).x;
    scattering   = mat.scattering * eval_texture(sc, mat.scattering_tex, texcoord);
    opacity_mean = mean3(eval_texture(sc, mat.opacity_tex, texcoord));
  }

  // opacity, :423-424, :453, :1429-1433 (mean of the {1,1,1} texture default is exactly 1)
  float opacity = mat.opacity * opacity_mean;
  if (opacity > 0.999f) opacity = 1;
  bool finished = false;
  if (SHADER == YH_SHADER_NORMAL) {  // trace_normal, :1636-1651: the shading normal as a colour, alpha 1
    radiance = normal * 0.5f + 0.5f;
    w4.w     = 1.0f;
    finished = true;
  } else if (opacity < 1 && rand1f(rng) >= opacity) {
    vec3f no = position + ray_d * 1e-2f;
    st_stream(&ps.ray_o[slot], make_float4(no.x, no.y, no.z, 0));  // same direction, bounce not counted
  } else {
    float alpha = 1.0f;
    radiance    = radiance + weight * emission;  // :1437

    // eval_brdf, :405-492: the hair lobe for line shapes (HAIR), the surface lobes of yh_surface.cuh
    // for triangles; a surface with roughness 0 is a delta brdf (:495) and takes the sample_delta branch
    vec3f         incoming;
    vec3f         brdfcos;
    float         brdf_pdf;
    hair_brdf     hb;
    hair_wo_terms ht;
    surface_brdf  sb;
    bool          delta = false;
    if (HAIR) {
      hb = mat.hair_proto;  // eval_hair_brdf, :474-490: the material's part comes from the table, the hit's part here
      hair_brdf_hit_part(hb, uv.y, normal, tangent);
      ht = hair_prepare_wo(hb, outgoing);
    } else {
      surface_material sm = {base, specular, metallic, roughness, mat.ior, transmission, mat.thin != 0 || mat.transmission == 0};
      sb    = eval_surface_brdf(sm, normal, outgoing);
      delta = is_delta(sb);
    }
    bool continues = true;  // the path goes on to `incoming` (false: this vertex ended the sample)
    if (SHADER == YH_SHADER_EYELIGHT) {  // trace_eyelight, :1576-1633: the light sits at the eye
      vec3f f;
      if (HAIR) {
        float unused_pdf;
        eval_hair_scattering_and_pdf(hb, ht, outgoing, f, unused_pdf, mat.hair_k);
      } else {
        f = eval_brdfcos(sb, normal, outgoing, outgoing);
      }
      radiance = radiance + weight * pif * f;
      if (!delta) {
        finished = true, continues = false;  // :1618
      } else {  // mirrors and glass sheets are followed, :1619-1626 (no Russian roulette, bound max(bounces, 4))
        float rnl = rand1f(rng);
        incoming  = sample_delta(sb, normal, outgoing, rnl);
        weight    = weight * (eval_delta(sb, normal, outgoing, incoming) / sample_delta_pdf(sb, normal, outgoing, incoming));
        if ((weight == vec3f{0, 0, 0}) || !isfinite3(weight)) finished = true;
        bounce += 1;
        if (bounce >= (tc.bounces > 4 ? tc.bounces : 4)) finished = true;
      }
    } else {
    if (delta) {  // :1452-1456, trace_naive :1552-1556
      float rnl = rand1f(rng);
      incoming  = sample_delta(sb, normal, outgoing, rnl);
      weight    = weight * (eval_delta(sb, normal, outgoing, incoming) / sample_delta_pdf(sb, normal, outgoing, incoming));
    } else {
    // MIS one-sample model, :1441-1451 (trace_naive, :1548-1551, always samples the BSDF).  Argument
    // evaluation order of the oracle build (GCC, right to left): rand2f before rand1f, see camera_ray().
    if (SHADER == YH_SHADER_NAIVE || rand1f(rng) < 0.5f) {
      vec2f rn  = rand2f(rng);
      float rnl = rand1f(rng);
      if (HAIR) incoming = sample_hair_scattering(hb, ht, rn, mat.hair_k);  // :1133-1135
      else incoming = sample_brdfcos(sb, normal, outgoing, rnl, rn);
    } else {
      vec2f ruv = rand2f(rng);
      float rel = rand1f(rng);
      float rl  = rand1f(rng);
      incoming  = sample_lights(sc, position, rl, rel, ruv);
    }
    if (HAIR) {
      eval_hair_scattering_and_pdf(hb, ht, incoming, brdfcos, brdf_pdf, mat.hair_k);  // :1071-1073, :1219-1221
    } else {
      brdfcos  = eval_brdfcos(sb, normal, outgoing, incoming);
      brdf_pdf = sample_brdfcos_pdf(sb, normal, outgoing, incoming);
    }
    if (SHADER == YH_SHADER_NAIVE) {
      weight = weight * (brdfcos / brdf_pdf);
    } else {
      float light_pdf = sample_lights_pdf(sc, position, incoming, st, light_rays);
      weight          = weight * (brdfcos / (0.5f * brdf_pdf + 0.5f * light_pdf));
    }
    }

    // :1500-1507, then the loop condition bounce < params.bounces
    if ((weight == vec3f{0, 0, 0}) || !isfinite3(weight)) finished = true;
    if (!finished && bounce > 3) {
      float rr_prob = fmin_(0.99f, max3(weight));
      if (rand1f(rng) >= rr_prob) finished = true;
      else weight = weight * div_(1.0f, rr_prob);
    }
    bounce += 1;
    if (bounce >= tc.bounces) finished = true;
    }
    // update the volume stack, :1458-1467: crossing a refractive surface enters its volume or leaves the current one
    if (VOL && !HAIR && !mat.thin && mat.transmission != 0 && dot(normal, outgoing) * dot(normal, incoming) < 0) {
      if (!has_vol) {
        vsdf v = eval_vsdf(base, transmission, mat.thin != 0 || mat.transmission == 0, scattering, mat.scanisotropy, mat.trdepth);
        st_stream(&ps.vol_density[slot], make_float4(v.density.x, v.density.y, v.density.z, v.anisotropy));
        st_stream(&ps.vol_scatter[slot], make_float4(v.scatter.x, v.scatter.y, v.scatter.z, 1.0f));
      } else {
        st_stream(&ps.vol_scatter[slot], make_float4(0, 0, 0, 0));
      }
    }
    if (continues) {
      st_stream(&ps.ray_o[slot], make_float4(position.x, position.y, position.z, 0));
      st_stream(&ps.ray_d[slot], make_float4(incoming.x, incoming.y, incoming.z, 0));
    }
    w4.w = alpha;
  }
  bool again;
  if (finished) {
    again = finish_sample(ps, cam, tc, slot, radiance, w4.w, rng);
  } else {
    st_stream(&ps.weight[slot], make_float4(weight.x, weight.y, weight.z, w4.w));
    st_stream(&ps.radiance[slot], make_float4(radiance.x, radiance.y, radiance.z, __int_as_float(bounce)));
    again             = true;
  }
  st_stream(&ps.rng_state[slot], rng.state);
  return again;
}

// one instantiation per (primitive kind, shader): the shader is a compile-time constant so the
// path shader's code carries none of the other shaders' branches
template <bool HAIR, int SHADER, bool VOL>
#ifndef YH_SHADE_BLOCKS
#define YH_SHADE_BLOCKS 4
#endif
__global__ void __launch_bounds__(kTraceBlock, YH_SHADE_BLOCKS) k_shade(dev_scene sc, path_state ps, queues q, dev_camera cam,
    trace_consts tc, int next) {
  __shared__ int smem_stack[kStackShared * kTraceBlock];
  int            n     = q.count[HAIR ? 4 : 3];
  const int*     queue = HAIR ? q.hair : q.surface;
  unsigned int   light_rays = 0;
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    int  idx  = base + threadIdx.x;
    bool live = idx < n, again = false;
    int  slot = live ? queue[idx] : 0;
    if (live) {
      int        stack_overflow[kStackLocal];
      trav_stack st;
      st.init(smem_stack + threadIdx.x, stack_overflow);
      again   = shade_vertex<HAIR, SHADER, VOL>(sc, ps, cam, tc, slot, st, light_rays);
    }
    queue_push(q.active[next], q.count + next, again, slot);
  }
  if (light_rays) atomicAdd(q.stats + 1, (unsigned long long)light_rays);
}

// stop request: no slot starts another sample; the samples in flight finish, so every pixel still holds whole samples
// of its stream.  cancelled += the samples that will not be traced.
__global__ void k_cancel_remaining(path_state ps, int num_slots, unsigned long long* cancelled) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  int left = 0;
  if (slot < num_slots) {
    left = ld_stream(&ps.remaining[slot]);
    if (left > 0) st_stream(&ps.remaining[slot], 0);
  }
  for (int d = 16; d > 0; d >>= 1) left += __shfl_down_sync(0xffffffffu, left, d);
  if ((threadIdx.x & 31) == 0 && left > 0) atomicAdd(cancelled, (unsigned long long)left);
}

// before the traversal of a round that reads queue `next ^ 1`: its outputs start empty (the next active queue, the three
// shading queues, the ray cursor and the suspend records the traversal will write, count[6 + next])
__global__ void k_reset_counts(int* count, int next) {
  count[next] = 0;
  count[2] = count[3] = count[4] = count[5] = 0;
  count[6 + next] = 0;
}

// init_state, yocto_pathtrace.cpp:1931-1946: pixel k takes the k-th draw of the
// master stream make_rng(1301081): seq = rand1i(rng, 1 << 31) / 2 + 1.
__host__ __device__ inline rng_state rng_skip(rng_state rng, unsigned long long delta) {
  unsigned long long cur_mult = 6364136223846793005ULL, cur_plus = rng.inc, acc_mult = 1u, acc_plus = 0u;
  while (delta > 0) {
    if (delta & 1) {
      acc_mult *= cur_mult;
      acc_plus = acc_plus * cur_mult + cur_plus;
    }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  rng.state = acc_mult * rng.state + acc_plus;
  return rng;
}
__global__ void k_init_slots(path_state ps, const int* slot_pixel, int num_slots, unsigned long long seed) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= num_slots) return;
  int       px     = slot_pixel[slot] & kPixelMask;
  rng_state master = rng_skip(make_rng(1301081), (unsigned long long)px);
  uint32_t  r      = advance_rng(master) & 0x7fffffffu;  // % (1 << 31)
  rng_state rng    = make_rng(seed, (uint64_t)(r / 2 + 1));
  st_stream(&ps.rng_state[slot], rng.state);
  st_stream(&ps.rng_inc[slot], rng.inc);
  st_stream(&ps.rng_base[slot], rng.inc);
  st_stream(&ps.pixel[slot], slot_pixel[slot]);
  st_stream(&ps.remaining[slot], 0);
}

// several slots per pixel: the lanes' partial images into the accumulation buffer, in lane order (deterministic)
__global__ void __launch_bounds__(256) k_fold_partials(float4* accum, int* samples, float4* partial, int* partial_samples, int n, int lanes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = accum[i];
  int    s = samples[i];
  for (int k = 0; k < lanes; k++) {
    size_t j = (size_t)k * n + i;
    float4 b = partial[j];
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    s += partial_samples[j];
    partial[j]         = make_float4(0, 0, 0, 0);
    partial_samples[j] = 0;
  }
  accum[i]   = a;
  samples[i] = s;
}

__global__ void k_resolve(const float4* accum, const int* samples, float4* render, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = accum[i];
  int    s = samples[i];
  // pixel.accumulated / pixel.samples, :1688 (render stays 0 before the first sample)
  render[i] = s > 0 ? make_float4(div_(a.x, (float)s), div_(a.y, (float)s), div_(a.z, (float)s), div_(a.w, (float)s))
                    : make_float4(0, 0, 0, 0);
}

struct batch_io {
  const yh_ray* rays;
  yh_hit*       hits;
  __device__ __forceinline__ bool load(int index, int& handle, ray3f& r) const {
    yh_ray y = rays[index];
    handle   = index;
    r        = {{y.o[0], y.o[1], y.o[2]}, {y.d[0], y.d[1], y.d[2]}, y.tmin, y.tmax};
    return true;
  }
  __device__ __forceinline__ void reload(int handle, vec3f& o, vec3f& d) const {
    const yh_ray& y = rays[handle];
    o = {y.o[0], y.o[1], y.o[2]}, d = {y.d[0], y.d[1], y.d[2]};
  }
  __device__ __forceinline__ void store(int handle, const hit_record& h) const {
    hits[handle] = {h.object, h.element, {h.uv.x, h.uv.y}, h.distance, h.hit ? 1 : 0};
  }
};
__global__ void __launch_bounds__(kTraceBlock) k_intersect_batch(dev_scene sc, const yh_ray* rays, int n,
    int find_any, yh_hit* hits, int* cursor) {
  __shared__ int smem_stack[kStackShared * kTraceBlock];
  __shared__ int smem_park[kParkSlots * kTraceBlock];
  __shared__ int smem_cold[kColdWords * kTraceBlock];
  batch_io       io{rays, hits};
  trace_persistent<false>(sc, io, n, cursor, smem_stack + threadIdx.x, smem_park + threadIdx.x, smem_cold + threadIdx.x,
      find_any != 0, nullptr, nullptr);
}

}  // namespace yh

using namespace yh;

// one resident wave of kTraceBlock-thread blocks (persistent kernels)
static int persistent_grid(const void* kernel) {
  int per_sm = 0;
  // Shared memory and L1 share 228 KB per SM.  The node fetches of the traversal hit L1 about
  // half of the time, so the carve-out is a trade between resident blocks and L1 capacity
  // (measured: the largest carve-out costs 30 % of the traversal rate).  YH_CARVEOUT (percent)
  // overrides the driver's choice for experiments.
  if (const char* c = getenv("YH_CARVEOUT")) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c));
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTraceBlock, 0) != cudaSuccess || per_sm < 1)
    per_sm = 4;
  return per_sm * sm_count();
}

// One wavefront group: a contiguous range of path slots with its own queues, counters and
// stream.  The traversal kernel of one group ends with a tail (a few long rays keep a few
// warps busy while most SMs idle: ncu showed SMs active 74 % of the kernel's duration), so the
// slots are split into groups whose kernel sequences run on separate streams: the tail of one
// group's traversal overlaps the bulk of another group's work.  Groups own disjoint pixels.
// Pinned 64-byte read-back buffers are pooled for the life of the process: cudaMallocHost /
// cudaFreeHost synchronise the device and were seen to take up to a second on a busy context.
static std::mutex        pinned_mutex;
static std::vector<int*> pinned_free;
static int* pinned_acquire() {
  std::lock_guard<std::mutex> lock(pinned_mutex);
  if (!pinned_free.empty()) {
    int* p = pinned_free.back();
    pinned_free.pop_back();
    return p;
  }
  int* p = nullptr;
  if (cudaMallocHost((void**)&p, 64) != cudaSuccess) return nullptr;
  return p;
}
static void pinned_release(int* p) {
  std::lock_guard<std::mutex> lock(pinned_mutex);
  pinned_free.push_back(p);
}

struct wave_group {
  int                            first_slot = 0, num_slots = 0;
  dev_buffer<int>                q_active0, q_active1, q_miss, q_surface, q_hair, q_count;
  dev_buffer<unsigned>           kind_bits;  // 3 bitmaps of kind_words words (queues::kind_bits)
  int                            kind_words = 0;
  dev_buffer<int>                sus_records[2];  // rays carried over between launches (suspend_io), ping-pong
  int                            sus_capacity = 0;
  dev_buffer<unsigned long long> stats;
  cudaStream_t                   stream      = nullptr;
  cudaEvent_t                    ev_done     = nullptr;
  int*                           host_counts = nullptr;  // pinned
  int                            cur = 0;
  bool                           done = false;
  queues qs() { return {{q_active0.ptr, q_active1.ptr}, q_miss.ptr, q_surface.ptr, q_hair.ptr, q_count.ptr, kind_bits.ptr, kind_words, stats.ptr}; }
  ~wave_group() {
    if (host_counts) pinned_release(host_counts);
    if (ev_done) cudaEventDestroy(ev_done);
    if (stream) cudaStreamDestroy(stream);
  }
};

struct yh_state {
  const yh_scene* scene = nullptr;
  dev_camera      camera{};
  yh_trace_params params{};
  int             width = 0, height = 0, num_slots = 0;
  int             rank = 0, world = 1;
  bool            counter = false;          // per-sample streams (yh_state_set_sample_shard)
  int             srank = 0, sworld = 1;    // sample indices srank, srank + sworld, ... are rendered here
  int             slots_per_pixel = 1;      // counter mode only (yh_state_set_sample_slots)
  int             ordinal_base = 0;         // sample ordinals handed out so far (several slots per pixel)
  int             num_owned = 0;            // owned pixels (num_slots = num_owned * slots_per_pixel)
  dev_buffer<int>    ordinal, partial_samples;
  dev_buffer<float4> partial;
  dev_buffer<float4>             ray_o, ray_d, weight, radiance, accum, render, vol_density, vol_scatter;
  dev_buffer<float>              hit_dist;
  dev_buffer<int4>               hit;
  dev_buffer<unsigned long long> rng_state, rng_inc, rng_base;
  dev_buffer<int>                remaining, pixel, samples;
  std::vector<wave_group*>       groups;
  cudaStream_t    stream      = nullptr;  // set-up, download
  cudaEvent_t     ev_begin = nullptr, ev_end = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  bool            profile = false;
  yh_trace_stats  last{};
  // path state of the slots [first, ...): kernels index it with group-local slot ids
  path_state      ps(int first = 0) {
    return {ray_o.ptr + first, ray_d.ptr + first, weight.ptr + first, radiance.ptr + first, hit.ptr + first,
        rng_state.ptr + first, rng_inc.ptr + first, rng_base.ptr + first, remaining.ptr + first, pixel.ptr + first,
        accum.ptr, samples.ptr, vol_density.ptr ? vol_density.ptr + first : nullptr,
        vol_scatter.ptr ? vol_scatter.ptr + first : nullptr, hit_dist.ptr ? hit_dist.ptr + first : nullptr,
        ordinal.ptr ? ordinal.ptr + first : nullptr, partial.ptr, partial_samples.ptr, width * height};
  }
  void drain() {  // nothing of this state is in flight afterwards
    for (auto g : groups)
      if (g->stream) cudaStreamSynchronize(g->stream);
    if (stream) cudaStreamSynchronize(stream);
  }
  void clear_groups() {
    drain();
    for (auto g : groups) delete g;
    groups.clear();
  }
  ~yh_state() {
    clear_groups();
    for (auto e : ev_pool) cudaEventDestroy(e);
    if (ev_begin) cudaEventDestroy(ev_begin);
    if (ev_end) cudaEventDestroy(ev_end);
    if (stream) cudaStreamDestroy(stream);
  }
};

// Slot order: 32x32 tiles round-robin over ranks; inside a tile 8x4 pixel
// blocks, so one warp starts with a compact bundle of camera rays.  The owned tiles are
// dealt round-robin to `ngroups` wavefront groups (group g = slots [first[g], first[g + 1])),
// which interleaves the groups over the image and balances their load.
static std::vector<int> owned_pixels(int w, int h, int rank, int world, int ngroups = 1, std::vector<int>* first = nullptr) {
  std::vector<int> px;
  const int        T   = 32;
  int              tx  = (w + T - 1) / T, ty = (h + T - 1) / T;
  if (first) first->assign(1, 0);
  for (int g = 0; g < ngroups; g++) {
    int owned = 0;
    for (int t = 0; t < tx * ty; t++) {
      if (t % world != rank) continue;
      if (owned++ % ngroups != g) continue;
      int x0 = (t % tx) * T, y0 = (t / tx) * T;
      for (int by = 0; by < T; by += 4)
        for (int bx = 0; bx < T; bx += 8)
          for (int y = 0; y < 4; y++)
            for (int x = 0; x < 8; x++) {
              int i = x0 + bx + x, j = y0 + by + y;
              if (i < w && j < h) px.push_back(j * w + i);
            }
    }
    if (first) first->push_back((int)px.size());
  }
  return px;
}

#ifndef YH_WAVE_GROUPS
#define YH_WAVE_GROUPS 2
#endif
constexpr long long kOneGroupBelowSlots = 800000;
static yh_status assign_slots(yh_state* st) {
  int ngroups = YH_WAVE_GROUPS;
  if (const char* e = getenv("YH_WAVE_GROUPS")) ngroups = atoi(e);
  int S = st->counter ? st->slots_per_pixel : 1;
  if ((size_t)st->width * st->height > (size_t)kPixelMask) S = 1;  // the lane lives in the top byte of the pixel word
  // Two groups overlap one group's shading with the other's traversal, but each then runs half as wide.  Measured
  // (profiles/r02_s4_exp_wave_groups.txt): below ~0.8 M slots one group wins (291 k: +2 %, 518 k: +3-4 %), above it two
  // do (1.64 M: +7 %); three or four never.
  long long slots_estimate = (long long)st->width * st->height / st->world * S;
  if (slots_estimate < kOneGroupBelowSlots && !getenv("YH_WAVE_GROUPS")) ngroups = 1;
  if (ngroups < 1) ngroups = 1;
  if (ngroups > 8) ngroups = 8;
  std::vector<int> first;
  std::vector<int> px = owned_pixels(st->width, st->height, st->rank, st->world, ngroups, &first);
  st->num_owned = (int)px.size();
  st->slots_per_pixel = S;
  st->ordinal_base    = 0;
  st->ordinal.release(), st->partial.release(), st->partial_samples.release();
  if (S > 1) {  // lane-major inside a group: lane k of all its pixels, then lane k + 1 (warps keep compact pixel bundles)
    std::vector<int> wide, wfirst(1, 0);
    wide.reserve(px.size() * S);
    for (int g = 0; g < ngroups; g++) {
      for (int k = 0; k < S; k++)
        for (int i = first[g]; i < first[g + 1]; i++) wide.push_back(px[i] | (k << 24));
      wfirst.push_back((int)wide.size());
    }
    px.swap(wide), first.swap(wfirst);
    size_t np = (size_t)st->width * st->height;
    YH_CUDA(st->ordinal.alloc(px.size() ? px.size() : 1));
    YH_CUDA(st->partial.alloc(np * S));
    YH_CUDA(st->partial_samples.alloc(np * S));
    YH_CUDA(cudaMemset(st->partial.ptr, 0, st->partial.bytes()));
    YH_CUDA(cudaMemset(st->partial_samples.ptr, 0, st->partial_samples.bytes()));
    YH_CUDA(cudaDeviceSynchronize());
  }
  st->num_slots       = (int)px.size();
  size_t n            = px.size() ? px.size() : 1;
  YH_CUDA(st->ray_o.alloc(n));
  YH_CUDA(st->ray_d.alloc(n));
  YH_CUDA(st->weight.alloc(n));
  YH_CUDA(st->radiance.alloc(n));
  YH_CUDA(st->hit.alloc(n));
  YH_CUDA(st->rng_state.alloc(n));
  YH_CUDA(st->rng_inc.alloc(n));
  YH_CUDA(st->rng_base.alloc(n));
  YH_CUDA(st->remaining.alloc(n));
  YH_CUDA(st->pixel.alloc(n));
  if (st->scene->has_volumes) {
    YH_CUDA(st->vol_density.alloc(n));
    YH_CUDA(st->vol_scatter.alloc(n));
    YH_CUDA(st->hit_dist.alloc(n));
  }
  st->clear_groups();
  for (int g = 0; g < ngroups; g++) {
    auto grp = new wave_group{};
    st->groups.push_back(grp);
    grp->first_slot = first[g], grp->num_slots = first[g + 1] - first[g];
    size_t m        = grp->num_slots ? grp->num_slots : 1;
    YH_CUDA(grp->q_active0.alloc(m));
    YH_CUDA(grp->q_active1.alloc(m));
    YH_CUDA(grp->q_miss.alloc(m));
    YH_CUDA(grp->q_surface.alloc(m));
    YH_CUDA(grp->q_hair.alloc(m));
    grp->kind_words = (int)((m + 1023) / 1024) * 32;
    YH_CUDA(grp->kind_bits.alloc((size_t)3 * grp->kind_words));
    YH_CUDA(cudaMemset(grp->kind_bits.ptr, 0, grp->kind_bits.bytes()));
    YH_CUDA(grp->q_count.alloc(8));
    YH_CUDA(grp->stats.alloc(8));
    {  // Suspend records (yh_traverse.cuh): the long rays of a launch's tail go to the next launch.  It pays when a
       // launch is small against the device - the tail is then a large part of it (C2: +9 %) - and costs about a
       // percent when launches are large (C3, C5: more, smaller wavefront iterations), so groups above kSuspendMaxSlots
       // slots run the plain kernel.  YH_SUSPEND_CAP (records per group, 0 = off) overrides.
      const size_t kSuspendMaxSlots = 600000;
      int cap = m <= kSuspendMaxSlots ? 1 << 15 : 0;
      if (const char* e = getenv("YH_SUSPEND_CAP")) cap = std::max(0, atoi(e));
      grp->sus_capacity = (int)std::min<size_t>((size_t)cap, m);
      if (grp->sus_capacity > 0) {
        YH_CUDA(grp->sus_records[0].alloc((size_t)grp->sus_capacity * kSuspendWords));
        YH_CUDA(grp->sus_records[1].alloc((size_t)grp->sus_capacity * kSuspendWords));
      }
    }

    YH_CUDA(cudaStreamCreateWithFlags(&grp->stream, cudaStreamNonBlocking));
    YH_CUDA(cudaEventCreateWithFlags(&grp->ev_done, cudaEventDisableTiming));
    grp->host_counts = pinned_acquire();
    if (!grp->host_counts) return fail(YH_ERR_CUDA, "pinned host allocation failed");
  }
  if (st->num_slots) {
    dev_buffer<int> d_px;
    YH_CUDA(d_px.upload(px.data(), px.size()));
    k_init_slots<<<(st->num_slots + 255) / 256, 256, 0, st->stream>>>(st->ps(), d_px.ptr, st->num_slots, st->params.seed);
    YH_CUDA(cudaGetLastError());
    YH_CUDA(cudaStreamSynchronize(st->stream));
  }
  return YH_OK;
}

extern "C" {

yh_status yh_state_create(const yh_scene* scene, const yh_camera_desc* camera, const yh_trace_params* params,
    yh_state** out) {
  YH_REQUIRE_DEVICE();
  if (!scene || !camera || !params || !out) return fail(YH_ERR_INVALID, "yh_state_create: null argument");
  *out = nullptr;
  if (params->shader < YH_SHADER_NAIVE || params->shader > YH_SHADER_NORMAL)
    return fail(YH_ERR_INVALID, "shader %d: sampler unknown", params->shader);
  if (!scene->bvh_built || !scene->lights_built)
    return fail(YH_ERR_STATE, "yh_state_create: call yh_scene_init_bvh and yh_scene_init_lights first");
  // only trace_path reads the lights (sample_lights / sample_lights_pdf, :1441-1451); the other shaders render
  // emitter-less scenes as the reference does
  if (scene->lights.empty() && params->shader == YH_SHADER_PATH)
    return fail(YH_ERR_INVALID, "scene has no lights (the reference's sample_lights is undefined then); use the eyelight, naive or normal shader");
  if (params->resolution <= 0 || params->bounces < 0) return fail(YH_ERR_INVALID, "bad trace params");
  auto st    = new yh_state{};
  st->scene  = scene;
  st->params = *params;
  st->camera = {{{camera->frame[0], camera->frame[1], camera->frame[2]}, {camera->frame[3], camera->frame[4], camera->frame[5]},
                    {camera->frame[6], camera->frame[7], camera->frame[8]}, {camera->frame[9], camera->frame[10], camera->frame[11]}},
      camera->lens, {camera->film[0], camera->film[1]}, camera->focus, camera->aperture};
  // image size, :1933-1939
  float fx = camera->film[0], fy = camera->film[1];
  if (fx > fy) {
    st->width  = params->resolution;
    st->height = (int)std::round(params->resolution * fy / fx);
  } else {
    st->width  = (int)std::round(params->resolution * fx / fy);
    st->height = params->resolution;
  }
  auto bail = [&](yh_status s) {
    delete st;
    return s;
  };
  if (st->width <= 0 || st->height <= 0) return bail(fail(YH_ERR_INVALID, "empty image"));
  size_t np = (size_t)st->width * st->height;
  if (cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&st->ev_begin) != cudaSuccess ||
      cudaEventCreate(&st->ev_end) != cudaSuccess)
    return bail(fail(YH_ERR_CUDA, "stream/event creation failed"));
  if (st->accum.alloc(np) != cudaSuccess || st->render.alloc(np) != cudaSuccess || st->samples.alloc(np) != cudaSuccess)
    return bail(fail(YH_ERR_CUDA, "device allocation failed"));
  cudaMemset(st->accum.ptr, 0, st->accum.bytes());
  cudaMemset(st->render.ptr, 0, st->render.bytes());
  cudaMemset(st->samples.ptr, 0, st->samples.bytes());
  cudaDeviceSynchronize();  // the state's streams do not synchronise with the legacy default stream
  yh_status s = assign_slots(st);
  if (s != YH_OK) return bail(s);
  const char* prof = getenv("YH_PROFILE_KERNELS");
  st->profile      = prof && prof[0] == '1';
  *out             = st;
  return YH_OK;
}

void yh_state_destroy(yh_state* state) { delete state; }

yh_status yh_state_size(const yh_state* st, int* width, int* height) {
  if (!st || !width || !height) return fail(YH_ERR_INVALID, "yh_state_size: null argument");
  *width = st->width, *height = st->height;
  return YH_OK;
}

yh_status yh_state_set_shard(yh_state* st, int rank, int world) {
  YH_REQUIRE_DEVICE();
  if (!st || world <= 0 || rank < 0 || rank >= world) return fail(YH_ERR_INVALID, "yh_state_set_shard: bad rank/world");
  st->rank = rank, st->world = world;
  st->counter = false, st->srank = 0, st->sworld = 1, st->slots_per_pixel = 1;
  cudaMemset(st->accum.ptr, 0, st->accum.bytes());
  cudaMemset(st->samples.ptr, 0, st->samples.bytes());
  cudaDeviceSynchronize();
  return assign_slots(st);
}

yh_status yh_state_set_sample_shard(yh_state* st, int rank, int world) {
  YH_REQUIRE_DEVICE();
  if (!st || world <= 0 || rank < 0 || rank >= world) return fail(YH_ERR_INVALID, "yh_state_set_sample_shard: bad rank/world");
  st->rank = 0, st->world = 1;  // every pixel is rendered here ...
  st->counter = true, st->srank = rank, st->sworld = world;  // ... for the sample indices of this rank
  cudaMemset(st->accum.ptr, 0, st->accum.bytes());
  cudaMemset(st->samples.ptr, 0, st->samples.bytes());
  cudaDeviceSynchronize();
  return assign_slots(st);
}

yh_status yh_state_set_sample_slots(yh_state* st, int slots_per_pixel) {
  YH_REQUIRE_DEVICE();
  if (!st || slots_per_pixel < 0 || slots_per_pixel > 64) return fail(YH_ERR_INVALID, "yh_state_set_sample_slots: bad argument");
  if (!st->counter) return fail(YH_ERR_STATE, "yh_state_set_sample_slots: counter mode only (call yh_state_set_sample_shard first)");
  if (slots_per_pixel == 0) {  // enough paths in flight to fill the device: ~2 M
    long long np = (long long)st->width * st->height;
    slots_per_pixel = (int)std::min<long long>(8, std::max<long long>(1, (2000000 + np - 1) / std::max<long long>(1, np)));
  }
  st->slots_per_pixel = slots_per_pixel;
  st->drain();
  cudaMemset(st->accum.ptr, 0, st->accum.bytes());
  cudaMemset(st->samples.ptr, 0, st->samples.bytes());
  cudaDeviceSynchronize();
  return assign_slots(st);
}

yh_status yh_shard_pixels(int width, int height, int rank, int world, int* pixels, int64_t* count) {
  if (width <= 0 || height <= 0 || world <= 0 || rank < 0 || rank >= world || !count)
    return fail(YH_ERR_INVALID, "yh_shard_pixels: bad argument");
  std::vector<int> px = owned_pixels(width, height, rank, world);
  if (pixels) memcpy(pixels, px.data(), px.size() * sizeof(int));
  *count = (int64_t)px.size();
  return YH_OK;
}

yh_status yh_state_set_profiling(yh_state* st, int enabled) {
  if (!st) return fail(YH_ERR_INVALID, "yh_state_set_profiling: state is null");
  st->profile = enabled != 0;
  return YH_OK;
}

static yh_status trace_samples_impl(yh_state* st, int num_samples, const volatile int* stop);
yh_status yh_trace_samples(yh_state* st, int num_samples, const volatile int* stop) {
  YH_REQUIRE_DEVICE();
  yh_status status = trace_samples_impl(st, num_samples, stop);
  // An error return leaves kernels and copies in flight on the groups' (non-blocking) streams; buffers are handed
  // back to the pool through the legacy stream, which does not order against them.  Drain before the caller can
  // destroy or re-shard the state.
  if (status != YH_OK && st) st->drain();
  return status;
}
static yh_status trace_samples_impl(yh_state* st, int num_samples, const volatile int* stop) {
  if (!st || num_samples < 0) return fail(YH_ERR_INVALID, "yh_trace_samples: bad argument");
  st->last = yh_trace_stats{};
  if (num_samples == 0 || st->num_slots == 0) return YH_OK;
  dev_scene    sc = st->scene->view();
  trace_consts tc = {st->width, st->height, st->params.bounces, st->params.clamp, st->counter ? 1 : 0, st->srank, st->sworld,
      st->slots_per_pixel, (unsigned long long)st->params.seed, st->params.shader};
  int  sms    = sm_count();
  bool prof   = st->profile;  // per-kernel timing: every group runs on ONE stream, nothing overlaps
  int  grid_t = sms * 8, grid_s = sms * 4;
  int  grid_i = persistent_grid((const void*)k_intersect<false>);
  // the groups share the device: each group's persistent traversal grid takes its share of the resident blocks
  int  ngroups = (int)st->groups.size();
  int  grid_div = 1;
  if (const char* e = getenv("YH_GROUP_GRID_DIV")) grid_div = std::max(1, atoi(e));
  int  grid_ig = prof ? grid_i : std::max(sms, grid_i / std::min(grid_div, ngroups));
  auto stream_of = [&](wave_group* g) { return prof ? st->groups[0]->stream : g->stream; };
  size_t ev_used = 0;
  auto   ev = [&](cudaStream_t s) -> cudaEvent_t {
    if (ev_used == st->ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      st->ev_pool.push_back(e);
    }
    cudaEvent_t e = st->ev_pool[ev_used++];
    cudaEventRecord(e, s);
    return e;
  };
  // begin: the groups' streams start after ev_begin (recorded on the first group's stream)
  cudaStream_t s0 = st->groups[0]->stream;
  YH_CUDA(cudaEventRecord(st->ev_begin, s0));
  int64_t launches = 0, iterations = 0, rounds = 0;
  // YH_DEBUG_SYNC=1: synchronise after every launch and name the kernel that faulted
  static const bool debug_sync = getenv("YH_DEBUG_SYNC") != nullptr;
#define YH_DEBUG_KERNEL(name, stream_)                                                                              \
  do {                                                                                                               \
    if (debug_sync) {                                                                                                \
      cudaError_t e_ = cudaStreamSynchronize(stream_);                                                               \
      if (e_ != cudaSuccess) return fail(YH_ERR_CUDA, "%s failed in round %lld: %s", name, (long long)rounds, cudaGetErrorString(e_)); \
    }                                                                                                                \
  } while (0)
  for (auto g : st->groups) {
    cudaStream_t s = stream_of(g);
    if (s != s0) YH_CUDA(cudaStreamWaitEvent(s, st->ev_begin, 0));
    queues q = g->qs();
    YH_CUDA(cudaMemsetAsync(q.count, 0, 8 * sizeof(int), s));
    YH_CUDA(cudaMemsetAsync(q.stats, 0, 8 * sizeof(unsigned long long), s));
    // empty after every complete round; cleared anyway so that a call that ended in an error leaves nothing behind
    YH_CUDA(cudaMemsetAsync(q.kind_bits, 0, (size_t)3 * q.kind_words * sizeof(unsigned), s));
    g->cur = 0, g->done = g->num_slots == 0;
    if (g->done) continue;
    k_start<<<(g->num_slots + 255) / 256, 256, 0, s>>>(st->ps(g->first_slot), q, st->camera, tc, g->num_slots, num_samples, st->ordinal_base);
    YH_DEBUG_KERNEL("k_start", s);
    launches++;
  }
  // a scene without triangles / points (or without lines) never fills that kind's queue: its kernel is not launched
  bool any_lines = false, any_surface = false;
  for (auto& o : st->scene->objects) (st->scene->shapes[o.shape]->is_lines() ? any_lines : any_surface) = true;
  int64_t min_iters = (num_samples + st->slots_per_pixel - 1) / st->slots_per_pixel;  // every sample of a slot needs at least one iteration
  bool    stopped   = false;
  while (true) {
    bool any = false;
    for (auto g : st->groups) {
      if (g->done) continue;
      any             = true;
      cudaStream_t s  = stream_of(g);
      path_state   ps = st->ps(g->first_slot);
      queues       q  = g->qs();
      int          next = g->cur ^ 1;
      k_reset_counts<<<1, 1, 0, s>>>(q.count, next);
      if (prof) ev(s);
      if (g->sus_capacity > 0) {
        suspend_io sus = {g->sus_records[g->cur].ptr, 0, g->sus_records[next].ptr, nullptr, g->sus_capacity};
        k_intersect<true><<<grid_ig, kTraceBlock, 0, s>>>(sc, ps, q, g->cur, q.active[g->cur], sus);
      } else {
        k_intersect<false><<<grid_ig, kTraceBlock, 0, s>>>(sc, ps, q, g->cur, q.active[g->cur], suspend_io{});
      }
      YH_DEBUG_KERNEL("k_intersect", s);
      k_enumerate_kinds<<<dim3((unsigned)std::max(1, std::min(q.kind_words / 128, sms * 4)), 3), 128, 0, s>>>(q);
      YH_DEBUG_KERNEL("k_enumerate_kinds", s);
      if (prof) ev(s);
      k_shade_miss<<<grid_s, 256, 0, s>>>(sc, ps, q, st->camera, tc, next);
      YH_DEBUG_KERNEL("k_shade_miss", s);
      switch (tc.shader) {
        case YH_SHADER_NAIVE:
          if (any_surface) k_shade<false, YH_SHADER_NAIVE, false><<<grid_s, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
          if (any_lines) k_shade<true, YH_SHADER_NAIVE, false><<<grid_t, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
          break;
        case YH_SHADER_EYELIGHT:
          if (any_surface) k_shade<false, YH_SHADER_EYELIGHT, false><<<grid_s, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
          if (any_lines) k_shade<true, YH_SHADER_EYELIGHT, false><<<grid_t, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
          break;
        case YH_SHADER_NORMAL:
          if (any_surface) k_shade<false, YH_SHADER_NORMAL, false><<<grid_s, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
          if (any_lines) k_shade<true, YH_SHADER_NORMAL, false><<<grid_t, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
          break;
        default:
          if (st->scene->has_volumes) {  // only trace_path follows volumes; the volume-free kernels stay as they were
            if (any_surface) k_shade<false, YH_SHADER_PATH, true><<<grid_s, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
            if (any_lines) k_shade<true, YH_SHADER_PATH, true><<<grid_t, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
          } else {
            if (any_surface) k_shade<false, YH_SHADER_PATH, false><<<grid_s, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
            if (any_lines) k_shade<true, YH_SHADER_PATH, false><<<grid_t, kTraceBlock, 0, s>>>(sc, ps, q, st->camera, tc, next);
          }
      }
      YH_DEBUG_KERNEL("k_shade", s);
      if (prof) ev(s);
      launches += 4 + (any_surface ? 1 : 0) + (any_lines ? 1 : 0);
      iterations++;
      g->cur = next;
    }
    if (!any) break;
    rounds++;
    if (rounds >= min_iters && (rounds % 4 == 0 || rounds < min_iters + 4)) {
      for (auto g : st->groups)
        if (!g->done) YH_CUDA(cudaMemcpyAsync(g->host_counts, g->q_count.ptr, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream_of(g)));
      for (auto g : st->groups) {
        if (g->done) continue;
        YH_CUDA(cudaStreamSynchronize(stream_of(g)));
        if (g->host_counts[g->cur] == 0 && g->host_counts[6 + g->cur] == 0) g->done = true;  // no active path, no suspended ray
      }
    }
    if (stop && *stop && !stopped) {  // finish the samples in flight, start no new ones
      stopped = true;
      for (auto g : st->groups)
        if (!g->done && g->num_slots > 0) {
          k_cancel_remaining<<<(g->num_slots + 255) / 256, 256, 0, stream_of(g)>>>(st->ps(g->first_slot), g->num_slots, g->stats.ptr + 5);
          launches++;
        }
      min_iters = rounds;  // poll the counts from now on
    }
    if (rounds > (int64_t)num_samples * ((st->params.bounces > 4 ? st->params.bounces : 4) + 2) * 64 + 1024)
      return fail(YH_ERR_STATE, "yh_trace_samples: wavefront did not drain after %lld rounds", (long long)rounds);
  }
  // end: the first group's stream waits for the others, then ev_end
  for (auto g : st->groups) {
    cudaStream_t s = stream_of(g);
    if (s == s0) continue;
    YH_CUDA(cudaEventRecord(g->ev_done, s));
    YH_CUDA(cudaStreamWaitEvent(s0, g->ev_done, 0));
  }
  if (st->partial.ptr) {  // the lanes' partial images -> accumulation buffer, in lane order
    int np = st->width * st->height;
    k_fold_partials<<<(np + 255) / 256, 256, 0, s0>>>(st->accum.ptr, st->samples.ptr, st->partial.ptr, st->partial_samples.ptr, np, st->slots_per_pixel);
    launches++;
    st->ordinal_base += num_samples;  // also after a stop request: ordinals are never handed out twice
  }
  YH_CUDA(cudaEventRecord(st->ev_end, s0));
  YH_CUDA(cudaEventSynchronize(st->ev_end));
  for (auto g : st->groups) YH_CUDA(cudaStreamSynchronize(g->stream));
  YH_CUDA(cudaGetLastError());
  unsigned long long tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (auto g : st->groups) {
    unsigned long long hs[8];
    YH_CUDA(cudaMemcpy(hs, g->stats.ptr, sizeof(hs), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; i++) tot[i] += hs[i];
  }
  float ms = 0;
  YH_CUDA(cudaEventElapsedTime(&ms, st->ev_begin, st->ev_end));
  st->last.samples    = (int64_t)st->num_owned * num_samples - (int64_t)tot[5];  // less the samples a stop request cancelled
  st->last.rays       = (int64_t)tot[0];
  st->last.light_rays = (int64_t)tot[1];
  st->last.iterations = iterations;
  st->last.launches   = launches;
  st->last.device_ms  = ms;
  st->last.nodes_visited = (int64_t)tot[3];
  st->last.prims_tested  = (int64_t)tot[4];
  if (prof) {
    float tt = 0, ts = 0;
    for (size_t i = 0; i + 2 < ev_used; i += 3) {  // three events per group iteration
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, st->ev_pool[i], st->ev_pool[i + 1]);
      cudaEventElapsedTime(&b, st->ev_pool[i + 1], st->ev_pool[i + 2]);
      tt += a, ts += b;
    }
    st->last.traverse_ms = tt, st->last.shade_ms = ts;
  }
  return YH_OK;
}

yh_status yh_state_stats(const yh_state* st, yh_trace_stats* stats) {
  if (!st || !stats) return fail(YH_ERR_INVALID, "yh_state_stats: null argument");
  *stats = st->last;
  return YH_OK;
}

yh_status yh_state_download(const yh_state* cst, float* render, float* accum, int* samples) {
  YH_REQUIRE_DEVICE();
  if (!cst) return fail(YH_ERR_INVALID, "yh_state_download: state is null");
  auto   st = const_cast<yh_state*>(cst);
  size_t np = (size_t)st->width * st->height;
  if (render) {
    k_resolve<<<(int)((np + 255) / 256), 256, 0, st->stream>>>(st->accum.ptr, st->samples.ptr, st->render.ptr, (int)np);
    YH_CUDA(cudaGetLastError());
    YH_CUDA(cudaMemcpyAsync(render, st->render.ptr, np * 16, cudaMemcpyDeviceToHost, st->stream));
  }
  if (accum) YH_CUDA(cudaMemcpyAsync(accum, st->accum.ptr, np * 16, cudaMemcpyDeviceToHost, st->stream));
  if (samples) YH_CUDA(cudaMemcpyAsync(samples, st->samples.ptr, np * 4, cudaMemcpyDeviceToHost, st->stream));
  YH_CUDA(cudaStreamSynchronize(st->stream));
  return YH_OK;
}

yh_status yh_resolve_dev(const float* accum_dev, const int* samples_dev, int64_t num_pixels, float* render_dev) {
  YH_REQUIRE_DEVICE();
  if (num_pixels < 0 || num_pixels > 0x7fffffff || (num_pixels > 0 && (!accum_dev || !samples_dev || !render_dev)))
    return fail(YH_ERR_INVALID, "yh_resolve_dev: bad argument");
  if (num_pixels == 0) return YH_OK;
  k_resolve<<<(int)((num_pixels + 255) / 256), 256>>>((const float4*)accum_dev, samples_dev, (float4*)render_dev, (int)num_pixels);
  YH_CUDA(cudaGetLastError());
  YH_CUDA(cudaStreamSynchronize(0));
  return YH_OK;
}

yh_status yh_state_buffers_dev(yh_state* st, float** accum_dev, int** samples_dev) {
  if (!st) return fail(YH_ERR_INVALID, "yh_state_buffers_dev: state is null");
  if (accum_dev) *accum_dev = (float*)st->accum.ptr;
  if (samples_dev) *samples_dev = st->samples.ptr;
  return YH_OK;
}

yh_status yh_state_download_rng(const yh_state* st, uint64_t* rng) {
  YH_REQUIRE_DEVICE();
  if (!st || !rng) return fail(YH_ERR_INVALID, "yh_state_download_rng: null argument");
  // image layout: 2 x uint64 per pixel; pixels this rank does not own stay 0
  size_t                          np = (size_t)st->width * st->height;
  std::vector<unsigned long long> s(st->num_slots), inc(st->num_slots);
  std::vector<int>                px(st->num_slots);
  YH_CUDA(cudaMemcpy(s.data(), st->rng_state.ptr, 8 * (size_t)st->num_slots, cudaMemcpyDeviceToHost));
  YH_CUDA(cudaMemcpy(inc.data(), st->rng_inc.ptr, 8 * (size_t)st->num_slots, cudaMemcpyDeviceToHost));
  YH_CUDA(cudaMemcpy(px.data(), st->pixel.ptr, 4 * (size_t)st->num_slots, cudaMemcpyDeviceToHost));
  memset(rng, 0, np * 16);
  for (int i = 0; i < st->num_slots; i++) {  // (several slots per pixel: the last lane's generator)
    rng[2 * (size_t)(px[i] & kPixelMask)]     = s[i];
    rng[2 * (size_t)(px[i] & kPixelMask) + 1] = inc[i];
  }
  return YH_OK;
}

yh_status yh_intersect_batch_dev(const yh_scene* scene, const yh_ray* rays_dev, int64_t n, int find_any,
    yh_hit* hits_dev, float* ms) {
  YH_REQUIRE_DEVICE();
  if (!scene || n < 0 || (n > 0 && (!rays_dev || !hits_dev))) return fail(YH_ERR_INVALID, "yh_intersect_batch_dev: bad argument");
  if (!scene->bvh_built) return fail(YH_ERR_STATE, "yh_intersect_batch: call yh_scene_init_bvh first");
  if (n == 0) return YH_OK;
  cudaEvent_t e0, e1;
  YH_CUDA(cudaEventCreate(&e0));
  YH_CUDA(cudaEventCreate(&e1));
  const int64_t   chunk = 1ll << 30;  // the ray cursor is a 32-bit counter
  int64_t         nchunks = (n + chunk - 1) / chunk;
  dev_buffer<int> cursor;
  YH_CUDA(cursor.alloc((size_t)nchunks));
  YH_CUDA(cudaMemset(cursor.ptr, 0, cursor.bytes()));
  int grid = persistent_grid((const void*)k_intersect_batch);
  YH_CUDA(cudaEventRecord(e0));
  for (int64_t c = 0; c < nchunks; c++) {
    int64_t m = n - c * chunk < chunk ? n - c * chunk : chunk;
    int64_t g = (m + kTraceBlock - 1) / kTraceBlock;
    k_intersect_batch<<<(int)(g < grid ? g : grid), kTraceBlock>>>(scene->view(), rays_dev + c * chunk, (int)m, find_any,
        hits_dev + c * chunk, cursor.ptr + c);
  }
  YH_CUDA(cudaEventRecord(e1));
  YH_CUDA(cudaEventSynchronize(e1));
  YH_CUDA(cudaGetLastError());
  float t = 0;
  YH_CUDA(cudaEventElapsedTime(&t, e0, e1));
  if (ms) *ms = t;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return YH_OK;
}

yh_status yh_intersect_batch(const yh_scene* scene, const yh_ray* rays, int64_t n, int find_any, yh_hit* hits) {
  YH_REQUIRE_DEVICE();
  if (!scene || n < 0 || (n > 0 && (!rays || !hits))) return fail(YH_ERR_INVALID, "yh_intersect_batch: bad argument");
  if (n == 0) return YH_OK;
  dev_buffer<yh_ray> dr;
  dev_buffer<yh_hit> dh;
  YH_CUDA(dr.upload(rays, n));
  YH_CUDA(dh.alloc(n));
  yh_status st = yh_intersect_batch_dev(scene, dr.ptr, n, find_any, dh.ptr, nullptr);
  if (st != YH_OK) return st;
  YH_CUDA(cudaMemcpy(hits, dh.ptr, dh.bytes(), cudaMemcpyDeviceToHost));
  return YH_OK;
}

}  // extern "C"

#include <cstdlib>
// Batched hair-BSDF, PCG32 and libm entry points of the C ABI (test and
// measurement surface for rows E1-E15 / M1 of SURVEY.md section 8a).
#include <cmath>
#include <cstring>
#include <vector>

#include "yh_common.cuh"
#include "yh_hair.cuh"
#include "yh_rng.cuh"

namespace yh {

static_assert(sizeof(hair_brdf) == sizeof(yh_hair_brdf), "hair_brdf layout");
static_assert(sizeof(hair_material) == sizeof(yh_hair_material), "hair_material layout");

std::string& last_error_storage() {
  static thread_local std::string s;
  return s;
}
yh_status fail(yh_status code, const char* fmt, ...) {
  char    buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_storage() = buf;
  return code;
}
bool require_device() {
  int         n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    fail(YH_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
        e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return false;
  }
  return true;
}
void pool_keep_memory() {
  static thread_local int done_for = -1;
  int                     dev      = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;  // never trim on synchronise: the pool is the allocator's cache
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done_for = dev;
}
bool plain_allocations() {
  static const bool plain = getenv("YH_NO_POOL") != nullptr;
  return plain;
}
int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int                     dev        = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached;
}

__device__ __forceinline__ vec3f ld3(const float* p, int64_t i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
__device__ __forceinline__ void  st3(float* p, int64_t i, const vec3f& v) {
  p[3 * i]     = v.x;
  p[3 * i + 1] = v.y;
  p[3 * i + 2] = v.z;
}

__global__ void k_hair_brdf(const hair_material* m, const float* v, const float* n, const float* t,
    int64_t count, hair_brdf* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = eval_hair_brdf(m[i], v[i], ld3(n, i), ld3(t, i));
}
__global__ void k_hair_eval(const hair_brdf* b, const float* wo, const float* wi, int64_t count, float* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    st3(out, i, eval_hair_scattering(b[i], ld3(wo, i), ld3(wi, i)));
}
__global__ void k_hair_sample(const hair_brdf* b, const float* wo, const float* rn, int64_t count, float* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    st3(out, i, sample_hair_scattering(b[i], ld3(wo, i), vec2f{rn[2 * i], rn[2 * i + 1]}));
}
__global__ void k_hair_pdf(const hair_brdf* b, const float* wo, const float* wi, int64_t count, float* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = sample_hair_scattering_pdf(b[i], ld3(wo, i), ld3(wi, i));
}
// what one shaded hair vertex costs in trace_path: sample + eval + pdf sharing
// the outgoing-direction terms
__global__ void k_hair_fused(const hair_brdf* b, const float* wo, const float* wi, const float* rn,
    int64_t count, float* f, float* pdf, float* ws) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    hair_brdf     br = b[i];
    hair_wo_terms t  = hair_prepare_wo(br, ld3(wo, i));
    vec3f         w  = ld3(wi, i);
    st3(ws, i, sample_hair_scattering(br, t, vec2f{rn[2 * i], rn[2 * i + 1]}));
    st3(f, i, eval_hair_scattering(br, t, w));
    pdf[i] = sample_hair_scattering_pdf(br, t, w);
  }
}
__global__ void k_libm(int fn, const float* a, const float* b, int64_t count, float* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    float r = 0;
    switch (fn) {
      case 0: r = expf_(a[i]); break;
      case 1: r = logf_(a[i]); break;
      case 2: r = sinf_(a[i]); break;
      case 3: r = cosf_(a[i]); break;
      case 4: r = atan2f_(a[i], b[i]); break;
      case 5: r = acosf_(a[i]); break;
      case 6: r = (float)asin_d((double)a[i]); break;
      case 7: r = (float)(1.0 / (sinh_d((double)div_(1.0f, a[i])) * 2.0 * (double)a[i])); break;
      case 8: r = fmodf_(a[i], b[i]); break;
      case 9: case 10: case 11: {  // the DOUBLE denominator of mp(), sinh(1 / v) * 2 * v, as an exact three-float expansion
        double den = sinh_d((double)div_(1.0f, a[i])) * 2.0 * (double)a[i];
        float  hi = (float)den, mid = (float)(den - (double)hi), lo = (float)(den - (double)hi - (double)mid);
        r = fn == 9 ? hi : (fn == 10 ? mid : lo);
      } break;
    }
    out[i] = r;
  }
}
__global__ void k_rng(uint64_t seed, uint64_t seq, int n, float* out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    rng_state rng = make_rng(seed, seq);
    for (int i = 0; i < n; i++) out[i] = rand1f(rng);
  }
}

// ---- the four statistical self-tests, yocto_extension.cpp:555-693 -----------
// The reference draws everything from ONE sequential PCG32 stream.  Each
// iteration consumes a fixed number of values, so thread i jumps the stream
// ahead (PCG's O(log n) LCG skip) and reproduces the reference's inputs.
__host__ __device__ inline rng_state rng_skip(rng_state rng, uint64_t delta) {
  uint64_t cur_mult = 6364136223846793005ULL, cur_plus = rng.inc, acc_mult = 1u, acc_plus = 0u;
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
__device__ inline vec3f sample_sphere(const vec2f& ruv) {  // yocto_math.h:4847-4852
  float z   = 2 * ruv.y - 1;
  float r   = sqrt_(clampf(1 - z * z, 0.0f, 1.0f));
  float phi = 2 * pif * ruv.x;
  return {r * cosf_(phi), r * sinf_(phi), z};
}
__device__ inline hair_brdf test_brdf(float beta_m, float beta_n, float h) {
  hair_material mat{{0, 0, 0}, beta_m, beta_n, 0.0f, 1.55f, {0, 0, 0}, 0, 0};
  return eval_hair_brdf(mat, h, {0, 0, 1}, {1, 0, 0});
}
struct selftest_cfg {
  float    beta_m, beta_n;
  uint64_t skip;  // stream position of the config's first draw
  vec3f    wo;
};
// which: 0 furnace, 1 furnace sampled, 2 weights, 3 consistency.  acc[4*cfg+..]
__global__ void k_selftest(int which, const selftest_cfg* cfgs, int count, rng_state base, double* acc) {
  int cfg_id       = blockIdx.y;
  selftest_cfg cfg = cfgs[cfg_id];
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    int       per = which == 2 ? 5 : 3;
    rng_state rng = rng_skip(base, cfg.skip + (uint64_t)per * (uint64_t)i);
    float     h   = rand1f(rng);
    if (which == 0) {
      if (h == 0) h += flt_eps;
      hair_brdf b  = test_brdf(cfg.beta_m, cfg.beta_n, h);
      vec3f     wi = sample_sphere(rand2f(rng));
      s0 += luminance(eval_hair_scattering(b, cfg.wo, wi));
    } else if (which == 1 || which == 2) {
      hair_brdf b  = test_brdf(cfg.beta_m, cfg.beta_n, h);
      vec3f     wo = which == 2 ? sample_sphere(rand2f(rng)) : cfg.wo;
      vec3f     wi = sample_hair_scattering(b, wo, rand2f(rng));
      vec3f     f  = eval_hair_scattering(b, wo, wi);
      float    pdf = sample_hair_scattering_pdf(b, wo, wi);
      if (pdf > 0) {
        float w = div_(luminance(f), pdf);
        if (which == 1) s0 += w;
        else {  // track the worst sample weight below / above 1 and failures
          s0 = fmax(s0, (double)fabs_(w - 1.0f));
          if (!(w >= 0.999f && w <= 1.001f)) s1 += 1;
        }
      }
    } else {
      hair_brdf b  = test_brdf(cfg.beta_m, cfg.beta_n, h);
      vec2f     u  = rand2f(rng);
      vec3f     wi = sample_hair_scattering(b, cfg.wo, u);
      vec3f     f  = eval_hair_scattering(b, cfg.wo, wi);
      float    pdf = sample_hair_scattering_pdf(b, cfg.wo, wi);
      if (pdf > 0) s0 += (double)luminance(f * (wi.z * wi.z) / (count * pdf));
      wi = sample_sphere(u);
      s1 += (double)luminance(eval_hair_scattering(b, cfg.wo, wi) * (wi.z * wi.z) /
                              (count * div_(1.0f, 4 * pif)));
    }
  }
  // warp shuffle reduction, then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    if (which == 2) s0 = fmax(s0, __shfl_down_sync(0xffffffffu, s0, o));
    else s0 += __shfl_down_sync(0xffffffffu, s0, o);
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
  }
  (void)s2; (void)s3;
  if ((threadIdx.x & 31) == 0) {
    if (which == 2) {
      // max via CAS on the bit pattern (values are non-negative)
      unsigned long long* p = (unsigned long long*)&acc[4 * cfg_id];
      unsigned long long  v = (unsigned long long)__double_as_longlong(s0);
      atomicMax(p, v);
    } else {
      atomicAdd(&acc[4 * cfg_id], s0);
    }
    atomicAdd(&acc[4 * cfg_id + 1], s1);
  }
}

static int grid_for(int64_t n, int block) {
  int64_t g   = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename T>
static cudaError_t up(dev_buffer<T>& b, const void* host, size_t n) {
  return b.upload((const T*)host, n);
}

}  // namespace yh

using namespace yh;

extern "C" {

int         yh_abi_version(void) { return YH_ABI_VERSION; }
const char* yh_last_error(void) { return last_error_storage().c_str(); }
yh_status   yh_device_count(int* count) {
  if (!count) return fail(YH_ERR_INVALID, "count is null");
  int         n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  *count        = e == cudaSuccess ? n : 0;
  return YH_OK;
}
yh_status yh_set_device(int device) {
  YH_REQUIRE_DEVICE();
  YH_CUDA(cudaSetDevice(device));
  return YH_OK;
}
yh_status yh_default_params(yh_trace_params* p) {
  if (!p) return fail(YH_ERR_INVALID, "params is null");
  *p = {720, YH_SHADER_PATH, 512, 8, 100.0f, 961748941ull, 0, 8};
  return YH_OK;
}
yh_status yh_default_material(yh_material_desc* m) {
  if (!m) return fail(YH_ERR_INVALID, "material is null");
  *m             = yh_material_desc{};
  m->ior         = 1.5f;
  m->spectint[0] = m->spectint[1] = m->spectint[2] = 1;
  m->trdepth                                       = 0.01f;
  m->opacity                                       = 1;
  m->beta_m = m->beta_n = 0.3f;
  m->alpha              = 2;
  m->eta                = 1.55f;
  m->emission_tex = m->color_tex = m->specular_tex = m->metallic_tex = m->roughness_tex = -1;
  m->transmission_tex = m->spectint_tex = m->scattering_tex = m->opacity_tex = m->normal_tex = -1;
  return YH_OK;
}

yh_status yh_hair_brdf_batch(const yh_hair_material* materials, const float* v, const float* normals,
    const float* tangents, int64_t n, yh_hair_brdf* brdfs) {
  YH_REQUIRE_DEVICE();
  if (n < 0 || (n > 0 && (!materials || !v || !normals || !tangents || !brdfs)))
    return fail(YH_ERR_INVALID, "yh_hair_brdf_batch: null argument");
  if (n == 0) return YH_OK;
  dev_buffer<hair_material> dm;
  dev_buffer<float>         dv, dn, dt;
  dev_buffer<hair_brdf>     db;
  YH_CUDA(up(dm, materials, n));
  YH_CUDA(up(dv, v, n));
  YH_CUDA(up(dn, normals, 3 * n));
  YH_CUDA(up(dt, tangents, 3 * n));
  YH_CUDA(db.alloc(n));
  k_hair_brdf<<<grid_for(n, 128), 128>>>(dm.ptr, dv.ptr, dn.ptr, dt.ptr, n, db.ptr);
  YH_CUDA(cudaGetLastError());
  YH_CUDA(cudaMemcpy(brdfs, db.ptr, db.bytes(), cudaMemcpyDeviceToHost));
  return YH_OK;
}

static yh_status hair_batch3(int kind, const yh_hair_brdf* brdfs, const float* wo, const float* second,
    int64_t n, float* out) {
  YH_REQUIRE_DEVICE();
  if (n < 0 || (n > 0 && (!brdfs || !wo || !second || !out)))
    return fail(YH_ERR_INVALID, "hair batch: null argument");
  if (n == 0) return YH_OK;
  dev_buffer<hair_brdf> db;
  dev_buffer<float>     dwo, ds, dout;
  YH_CUDA(up(db, brdfs, n));
  YH_CUDA(up(dwo, wo, 3 * n));
  YH_CUDA(up(ds, second, (kind == 1 ? 2 : 3) * n));
  YH_CUDA(dout.alloc((kind == 2 ? 1 : 3) * n));
  int g = grid_for(n, 128);
  if (kind == 0) k_hair_eval<<<g, 128>>>(db.ptr, dwo.ptr, ds.ptr, n, dout.ptr);
  if (kind == 1) k_hair_sample<<<g, 128>>>(db.ptr, dwo.ptr, ds.ptr, n, dout.ptr);
  if (kind == 2) k_hair_pdf<<<g, 128>>>(db.ptr, dwo.ptr, ds.ptr, n, dout.ptr);
  YH_CUDA(cudaGetLastError());
  YH_CUDA(cudaMemcpy(out, dout.ptr, dout.bytes(), cudaMemcpyDeviceToHost));
  return YH_OK;
}
yh_status yh_hair_eval_batch(const yh_hair_brdf* b, const float* wo, const float* wi, int64_t n, float* f) {
  return hair_batch3(0, b, wo, wi, n, f);
}
yh_status yh_hair_sample_batch(const yh_hair_brdf* b, const float* wo, const float* rn, int64_t n, float* wi) {
  return hair_batch3(1, b, wo, rn, n, wi);
}
yh_status yh_hair_pdf_batch(const yh_hair_brdf* b, const float* wo, const float* wi, int64_t n, float* pdf) {
  return hair_batch3(2, b, wo, wi, n, pdf);
}
yh_status yh_hair_fused_dev(const yh_hair_brdf* brdfs_dev, const float* wo_dev, const float* wi_dev,
    const float* rn_dev, int64_t n, float* f_dev, float* pdf_dev, float* ws_dev, float* ms) {
  YH_REQUIRE_DEVICE();
  if (n <= 0) return fail(YH_ERR_INVALID, "yh_hair_fused_dev: n <= 0");
  cudaEvent_t e0, e1;
  YH_CUDA(cudaEventCreate(&e0));
  YH_CUDA(cudaEventCreate(&e1));
  YH_CUDA(cudaEventRecord(e0));
  k_hair_fused<<<grid_for(n, 128), 128>>>((const hair_brdf*)brdfs_dev, wo_dev, wi_dev, rn_dev, n, f_dev,
      pdf_dev, ws_dev);
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

yh_status yh_rng_floats(uint64_t seed, uint64_t seq, int n, float* out) {
  YH_REQUIRE_DEVICE();
  if (n < 0 || (n > 0 && !out)) return fail(YH_ERR_INVALID, "yh_rng_floats: bad argument");
  if (n == 0) return YH_OK;
  dev_buffer<float> d;
  YH_CUDA(d.alloc(n));
  k_rng<<<1, 32>>>(seed, seq, n, d.ptr);
  YH_CUDA(cudaGetLastError());
  YH_CUDA(cudaMemcpy(out, d.ptr, d.bytes(), cudaMemcpyDeviceToHost));
  return YH_OK;
}

yh_status yh_libm_batch(int fn, const float* a, const float* b, int64_t n, float* out) {
  YH_REQUIRE_DEVICE();
  if (fn < 0 || fn > 11 || n < 0 || (n > 0 && (!a || !out))) return fail(YH_ERR_INVALID, "yh_libm_batch: bad argument");
  if (n == 0) return YH_OK;
  dev_buffer<float> da, dbb, dout;
  YH_CUDA(up(da, a, n));
  YH_CUDA(up(dbb, b ? b : a, n));
  YH_CUDA(dout.alloc(n));
  k_libm<<<grid_for(n, 256), 256>>>(fn, da.ptr, dbb.ptr, n, dout.ptr);
  YH_CUDA(cudaGetLastError());
  YH_CUDA(cudaMemcpy(out, dout.ptr, dout.bytes(), cudaMemcpyDeviceToHost));
  return YH_OK;
}

yh_status yh_hair_selftest(int which, int* passed, float* worst) {
  YH_REQUIRE_DEVICE();
  if (which < 0 || which > 3 || !passed) return fail(YH_ERR_INVALID, "yh_hair_selftest: bad argument");
  // host replay of the reference's loop structure (float loop counters included)
  rng_state base  = make_rng(199382389514ull);
  uint64_t  pos   = 0;
  auto      draw2 = [&](uint64_t at) {
    rng_state r = rng_skip(base, at);
    return rand2f(r);
  };
  auto host_sample_sphere = [](vec2f ruv) {
    float z = 2 * ruv.y - 1, r = sqrt_(clampf(1 - z * z, 0.0f, 1.0f)), phi = 2 * pif * ruv.x;
    return vec3f{r * cosf_(phi), r * sinf_(phi), z};
  };
  std::vector<selftest_cfg> cfgs;
  int   count   = which == 2 ? 10000 : (which == 3 ? 64 * 1024 : 300000);
  int   per     = which == 2 ? 5 : 3;
  vec3f wo      = {0, 0, 1};
  if (which == 0 || which == 1) {
    wo = host_sample_sphere(draw2(pos));
    pos += 2;
  }
  float bm0 = which == 3 ? 0.2f : 0.1f, bn0 = (which >= 2) ? 0.4f : 0.1f;
  for (float beta_m = bm0; beta_m < 1.0f; beta_m += 0.2f)
    for (float beta_n = bn0; beta_n < 1.0f; beta_n += 0.2f) {
      if (which == 3) {
        wo = host_sample_sphere(draw2(pos));
        pos += 2;
      }
      cfgs.push_back({beta_m, beta_n, pos, wo});
      pos += (uint64_t)per * count;
    }
  dev_buffer<selftest_cfg> dc;
  dev_buffer<double>       dacc;
  YH_CUDA(dc.upload(cfgs.data(), cfgs.size()));
  YH_CUDA(dacc.alloc(4 * cfgs.size()));
  YH_CUDA(cudaMemset(dacc.ptr, 0, dacc.bytes()));
  dim3 grid((count + 127) / 128 > 592 ? 592 : (count + 127) / 128, (unsigned)cfgs.size());
  k_selftest<<<grid, 128>>>(which, dc.ptr, count, base, dacc.ptr);
  YH_CUDA(cudaGetLastError());
  std::vector<double> acc(4 * cfgs.size());
  YH_CUDA(cudaMemcpy(acc.data(), dacc.ptr, dacc.bytes(), cudaMemcpyDeviceToHost));
  bool  ok = true;
  float w  = 0;
  for (size_t c = 0; c < cfgs.size(); c++) {
    double a0 = acc[4 * c], a1 = acc[4 * c + 1];
    if (which == 0) {  // avg = luminance(sum) / (count * 1/(4 pi)) in [0.95, 1.05]
      double avg = a0 / (count * (1.0 / (4 * 3.14159265358979323846)));
      if (!(avg >= 0.95 && avg <= 1.05)) ok = false;
      if (fabs(avg - 1) > w) w = (float)fabs(avg - 1);
    } else if (which == 1) {  // avg in [0.99, 1.01]
      double avg = a0 / count;
      if (!(avg >= 0.99 && avg <= 1.01)) ok = false;
      if (fabs(avg - 1) > w) w = (float)fabs(avg - 1);
    } else if (which == 2) {  // every weight in [0.999, 1.001]
      double worst_dev;
      memcpy(&worst_dev, &acc[4 * c], 8);
      if (a1 > 0) ok = false;
      if (worst_dev > w) w = (float)worst_dev;
    } else {  // |importance - uniform| / uniform < 0.05
      double err = fabs(a0 - a1) / a1;
      if (!(err < 0.05)) ok = false;
      if (err > w) w = (float)err;
    }
  }
  *passed = ok ? 1 : 0;
  if (worst) *worst = w;
  return YH_OK;
}

yh_status yh_dev_alloc(void** ptr, int64_t bytes) {
  YH_REQUIRE_DEVICE();
  if (!ptr || bytes < 0) return fail(YH_ERR_INVALID, "yh_dev_alloc: bad argument");
  YH_CUDA(cudaMalloc(ptr, (size_t)bytes));
  return YH_OK;
}
yh_status yh_dev_free(void* ptr) {
  YH_CUDA(cudaFree(ptr));
  return YH_OK;
}
yh_status yh_dev_upload(void* dst_dev, const void* src_host, int64_t bytes) {
  YH_REQUIRE_DEVICE();
  YH_CUDA(cudaMemcpy(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice));
  YH_CUDA(cudaStreamSynchronize(0));  // pageable source: the DMA may still be in flight when cudaMemcpy returns
  return YH_OK;
}
yh_status yh_dev_download(void* dst_host, const void* src_dev, int64_t bytes) {
  YH_REQUIRE_DEVICE();
  YH_CUDA(cudaMemcpy(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost));
  return YH_OK;
}
yh_status yh_dev_synchronize(void) {
  YH_REQUIRE_DEVICE();
  YH_CUDA(cudaDeviceSynchronize());
  return YH_OK;
}

}  // extern "C"

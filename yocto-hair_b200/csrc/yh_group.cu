// Several GPUs of one node behind the C ABI, in ONE process (yh_group_*).
//
// The reference spreads the image over host threads (parallel_for, yocto_pathtrace.cpp:1954-1989: rows
// handed out by an atomic counter).  Here the image is spread over the GPUs of a box: the scene is replicated in
// every GPU's HBM (a million segments are ~170 MB), every GPU renders its shard of the work with its own
// yh_state - image tiles (bit-identical to one GPU and to the reference) or sample ranges (counter-mode streams,
// full-width wavefront on every GPU) - and the only exchange is at the end: the fp32 accumulation and count buffers
// of all GPUs are summed on the first one.  That sum is ONE kernel on the root GPU which loads the peers' buffers
// straight over NVLink (peer access to their stream-ordered pools), adds them in rank order and writes totals and
// the resolved image (accumulated / samples) in the same pass - the collective and the resolve are fused, nothing
// is staged.  Without peer access the buffers are copied first (cudaMemcpyPeerAsync); YH_REDUCE_NCCL selects an
// ncclReduce over communicators from ncclCommInitAll instead (libnccl is loaded on demand, the library does not
// link against it).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "yh_scene.cuh"

namespace yh {

constexpr int kMaxGroup = 16;
struct peer_table {
  const float4* accum[kMaxGroup];
  const int*    samples[kMaxGroup];
  int           n;
};

// totals = sum over ranks (rank order, so the result does not depend on timing); render = accumulated / samples
// (yocto_pathtrace.cpp:1688).  One float4 + one int per pixel and rank: the kernel streams at NVLink speed.
__global__ void __launch_bounds__(256) k_group_reduce(peer_table t, float4* total_accum, int* total_samples, float4* render, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = __ldcs(t.accum[0] + i);
  int    s = __ldcs(t.samples[0] + i);
  for (int r = 1; r < t.n; r++) {
    float4 b = __ldcs(t.accum[r] + i);
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    s += __ldcs(t.samples[r] + i);
  }
  total_accum[i]   = a;
  total_samples[i] = s;
  float fs         = (float)s;
  render[i] = s > 0 ? make_float4(__fdiv_rn(a.x, fs), __fdiv_rn(a.y, fs), __fdiv_rn(a.z, fs), __fdiv_rn(a.w, fs))
                    : make_float4(0, 0, 0, 0);
}
__global__ void __launch_bounds__(256) k_group_resolve(const float4* total_accum, const int* total_samples, float4* render, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a  = total_accum[i];
  int    s  = total_samples[i];
  float  fs = (float)s;
  render[i] = s > 0 ? make_float4(__fdiv_rn(a.x, fs), __fdiv_rn(a.y, fs), __fdiv_rn(a.z, fs), __fdiv_rn(a.w, fs))
                    : make_float4(0, 0, 0, 0);
}

// libnccl, loaded on first use
struct nccl_api {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*)                                                        = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                                          = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t)   = nullptr;
  ncclResult_t (*GroupStart)()                                                                                     = nullptr;
  ncclResult_t (*GroupEnd)()                                                                                       = nullptr;
  const char* (*GetErrorString)(ncclResult_t)                                                                      = nullptr;
  bool load(std::string& why) {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) {
      why = dlerror() ? dlerror() : "libnccl.so.2 not found";
      return false;
    }
    auto sym = [&](const char* n) { return dlsym(lib, n); };
    CommInitAll    = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy    = (decltype(CommDestroy))sym("ncclCommDestroy");
    Reduce         = (decltype(Reduce))sym("ncclReduce");
    GroupStart     = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd       = (decltype(GroupEnd))sym("ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !Reduce || !GroupStart || !GroupEnd || !GetErrorString) {
      why = "libnccl lacks an expected symbol";
      dlclose(lib), lib = nullptr;
      return false;
    }
    return true;
  }
};
static nccl_api g_nccl;

struct group_member {
  int          device = 0;
  yh_scene*    scene  = nullptr;
  yh_state*    state  = nullptr;
  cudaStream_t stream = nullptr;  // copies / ncclReduce of this member
  float4*      stage_accum   = nullptr;  // on the ROOT device, only without peer access
  int*         stage_samples = nullptr;
  bool         peer_direct   = false;  // the root kernel may load this member's buffers
  ncclComm_t   comm          = nullptr;
};

}  // namespace yh

using namespace yh;

struct yh_group {
  std::vector<group_member> m;
  int     width = 0, height = 0;
  int     shard = YH_SHARD_TILES, reduce = YH_REDUCE_PEER;
  bool    have_state = false;
  float4* total_accum = nullptr;  // root device
  int*    total_samples = nullptr;
  float4* render = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  yh_trace_stats last{};
  float   reduce_ms = 0;
  int     root() const { return m[0].device; }
};

namespace {

// runs fn(i) for every member on its own host thread with that member's device current; the first failure wins
yh_status on_all_members(yh_group* g, const std::function<yh_status(int)>& fn, int first = 0) {
  int                      n = (int)g->m.size();
  std::vector<yh_status>   status(n, YH_OK);
  std::vector<std::string> message(n);
  std::vector<std::thread> threads;
  for (int i = first; i < n; i++)
    threads.emplace_back([&, i]() {
      if (cudaSetDevice(g->m[i].device) != cudaSuccess) {
        status[i]  = YH_ERR_CUDA;
        message[i] = "cudaSetDevice failed";
        return;
      }
      status[i] = fn(i);
      if (status[i] != YH_OK) message[i] = yh_last_error();
    });
  for (auto& t : threads) t.join();
  for (int i = first; i < n; i++)
    if (status[i] != YH_OK) return fail(status[i], "device %d: %s", g->m[i].device, message[i].c_str());
  return YH_OK;
}

void free_image_buffers(yh_group* g) {
  cudaSetDevice(g->root());
  for (auto& mem : g->m) {
    if (mem.stage_accum) cudaFree(mem.stage_accum), mem.stage_accum = nullptr;
    if (mem.stage_samples) cudaFree(mem.stage_samples), mem.stage_samples = nullptr;
  }
  if (g->total_accum) cudaFree(g->total_accum), g->total_accum = nullptr;
  if (g->total_samples) cudaFree(g->total_samples), g->total_samples = nullptr;
  if (g->render) cudaFree(g->render), g->render = nullptr;
}

}  // namespace

extern "C" {

yh_status yh_group_create(const yh_scene_desc* desc, const yh_bvh_params* bvh, const int* devices, int num_devices,
    yh_group** out) {
  YH_REQUIRE_DEVICE();
  if (!desc || !out || num_devices < 1 || num_devices > kMaxGroup) return fail(YH_ERR_INVALID, "yh_group_create: bad argument");
  *out = nullptr;
  int have = 0;
  YH_CUDA(cudaGetDeviceCount(&have));
  std::vector<int> devs(num_devices);
  for (int i = 0; i < num_devices; i++) {
    devs[i] = devices ? devices[i] : i;
    if (devs[i] < 0 || devs[i] >= have) return fail(YH_ERR_INVALID, "yh_group_create: device %d of %d", devs[i], have);
  }
  int prev = 0;
  cudaGetDevice(&prev);
  auto g = new yh_group{};
  g->m.resize(num_devices);
  for (int i = 0; i < num_devices; i++) g->m[i].device = devs[i];
  auto bail = [&](yh_status s) {
    std::string msg = yh_last_error();
    yh_group_destroy(g);
    cudaSetDevice(prev);
    last_error_storage() = msg;
    return s;
  };
  // every device uploads and builds its own replica, side by side (the whole build runs on the device)
  yh_bvh_params bp{};
  if (bvh) bp = *bvh;
  yh_status s = on_all_members(g, [&](int i) -> yh_status {
    yh_status c = yh_scene_create(desc, &g->m[i].scene);
    if (c != YH_OK) return c;
    yh_status b = yh_scene_init_bvh(g->m[i].scene, &bp);
    if (b != YH_OK) return b;
    b = yh_scene_init_lights(g->m[i].scene);
    if (b != YH_OK) return b;
    YH_CUDA(cudaStreamCreateWithFlags(&g->m[i].stream, cudaStreamNonBlocking));
    return YH_OK;
  });
  if (s != YH_OK) return bail(s);
  // let the root load the peers' stream-ordered allocations (the states' accumulation buffers)
  cudaSetDevice(devs[0]);
  g->m[0].peer_direct = true;
  for (int i = 1; i < num_devices; i++) {
    if (devs[i] == devs[0]) {  // a device may be listed more than once (every entry is one rank)
      g->m[i].peer_direct = true;
      continue;
    }
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, devs[0], devs[i]) != cudaSuccess || !can) continue;
    cudaError_t e = cudaDeviceEnablePeerAccess(devs[i], 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
      cudaGetLastError();
      continue;
    }
    cudaGetLastError();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, devs[i]) != cudaSuccess) continue;
    cudaMemAccessDesc access{};
    access.location.type = cudaMemLocationTypeDevice;
    access.location.id   = devs[0];
    access.flags         = cudaMemAccessFlagsProtReadWrite;
    if (cudaMemPoolSetAccess(pool, &access, 1) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    g->m[i].peer_direct = true;
  }
  if (cudaEventCreate(&g->ev0) != cudaSuccess || cudaEventCreate(&g->ev1) != cudaSuccess)
    return bail(fail(YH_ERR_CUDA, "event creation failed"));
  cudaSetDevice(prev);
  *out = g;
  return YH_OK;
}

void yh_group_destroy(yh_group* g) {
  if (!g) return;
  int prev = 0;
  cudaGetDevice(&prev);
  if (!g->m.empty()) free_image_buffers(g);
  for (auto& mem : g->m) {
    cudaSetDevice(mem.device);
    if (mem.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(mem.comm);
    if (mem.state) yh_state_destroy(mem.state);
    if (mem.scene) yh_scene_destroy(mem.scene);
    if (mem.stream) cudaStreamDestroy(mem.stream);
  }
  if (!g->m.empty()) {
    cudaSetDevice(g->root());
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
  }
  cudaSetDevice(prev);
  delete g;
}

yh_status yh_group_num_devices(const yh_group* g, int* n) {
  if (!g || !n) return fail(YH_ERR_INVALID, "yh_group_num_devices: null argument");
  *n = (int)g->m.size();
  return YH_OK;
}

yh_status yh_group_scene(const yh_group* g, int rank, const yh_scene** scene, int* device) {
  if (!g || rank < 0 || rank >= (int)g->m.size()) return fail(YH_ERR_INVALID, "yh_group_scene: bad argument");
  if (scene) *scene = g->m[rank].scene;
  if (device) *device = g->m[rank].device;
  return YH_OK;
}

yh_status yh_group_init_state(yh_group* g, const yh_camera_desc* camera, const yh_trace_params* params, int shard) {
  YH_REQUIRE_DEVICE();
  if (!g || !camera || !params) return fail(YH_ERR_INVALID, "yh_group_init_state: null argument");
  if (shard != YH_SHARD_TILES && shard != YH_SHARD_SAMPLES) return fail(YH_ERR_INVALID, "yh_group_init_state: unknown shard mode %d", shard);
  int prev = 0;
  cudaGetDevice(&prev);
  int       n = (int)g->m.size();
  yh_status s = on_all_members(g, [&](int i) -> yh_status {
    if (g->m[i].state) yh_state_destroy(g->m[i].state), g->m[i].state = nullptr;
    yh_status c = yh_state_create(g->m[i].scene, camera, params, &g->m[i].state);
    if (c != YH_OK) return c;
    if (shard == YH_SHARD_SAMPLES) return yh_state_set_sample_shard(g->m[i].state, i, n);
    return n > 1 ? yh_state_set_shard(g->m[i].state, i, n) : YH_OK;
  });
  g->have_state = false;
  if (s != YH_OK) {
    cudaSetDevice(prev);
    return s;
  }
  g->shard = shard;
  yh_state_size(g->m[0].state, &g->width, &g->height);
  size_t np = (size_t)g->width * g->height;
  free_image_buffers(g);
  cudaSetDevice(g->root());
  auto done = [&](yh_status st) {
    cudaSetDevice(prev);
    return st;
  };
  if (cudaMalloc((void**)&g->total_accum, np * 16) != cudaSuccess || cudaMalloc((void**)&g->total_samples, np * 4) != cudaSuccess ||
      cudaMalloc((void**)&g->render, np * 16) != cudaSuccess)
    return done(fail(YH_ERR_CUDA, "yh_group_init_state: device allocation failed"));
  for (int i = 1; i < n; i++)
    if (!g->m[i].peer_direct &&
        (cudaMalloc((void**)&g->m[i].stage_accum, np * 16) != cudaSuccess || cudaMalloc((void**)&g->m[i].stage_samples, np * 4) != cudaSuccess))
      return done(fail(YH_ERR_CUDA, "yh_group_init_state: staging allocation failed"));
  g->have_state = true;
  g->last = yh_trace_stats{};
  return done(YH_OK);
}

yh_status yh_group_set_reduce(yh_group* g, int mode) {
  if (!g || (mode != YH_REDUCE_PEER && mode != YH_REDUCE_NCCL)) return fail(YH_ERR_INVALID, "yh_group_set_reduce: bad argument");
  if (mode == YH_REDUCE_NCCL && !g->m[0].comm) {
    std::string why;
    if (!g_nccl.load(why)) return fail(YH_ERR_UNSUPPORTED, "NCCL is not available: %s", why.c_str());
    int                     n = (int)g->m.size();
    std::vector<ncclComm_t> comms(n);
    std::vector<int>        devs(n);
    for (int i = 0; i < n; i++) devs[i] = g->m[i].device;
    int prev = 0;
    cudaGetDevice(&prev);
    ncclResult_t r = g_nccl.CommInitAll(comms.data(), n, devs.data());
    cudaSetDevice(prev);
    if (r != ncclSuccess) return fail(YH_ERR_CUDA, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
    for (int i = 0; i < n; i++) g->m[i].comm = comms[i];
  }
  g->reduce = mode;
  return YH_OK;
}

yh_status yh_group_size(const yh_group* g, int* width, int* height) {
  if (!g || !width || !height) return fail(YH_ERR_INVALID, "yh_group_size: null argument");
  if (!g->have_state) return fail(YH_ERR_STATE, "yh_group_size: call yh_group_init_state first");
  *width = g->width, *height = g->height;
  return YH_OK;
}

yh_status yh_group_trace_samples(yh_group* g, int num_samples, const volatile int* stop) {
  YH_REQUIRE_DEVICE();
  if (!g || num_samples < 0) return fail(YH_ERR_INVALID, "yh_group_trace_samples: bad argument");
  if (!g->have_state) return fail(YH_ERR_STATE, "yh_group_trace_samples: call yh_group_init_state first");
  int n   = (int)g->m.size();
  int per = num_samples;
  if (g->shard == YH_SHARD_SAMPLES) {
    if (num_samples % n) return fail(YH_ERR_INVALID, "sample sharding: num_samples (%d) must be a multiple of the %d devices", num_samples, n);
    per = num_samples / n;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  std::vector<yh_trace_stats> stats(n);
  yh_status s = on_all_members(g, [&](int i) -> yh_status {
    yh_status t = yh_trace_samples(g->m[i].state, per, stop);
    if (t != YH_OK) return t;
    return yh_state_stats(g->m[i].state, &stats[i]);
  });
  cudaSetDevice(prev);
  if (s != YH_OK) return s;
  yh_trace_stats sum{};
  for (auto& t : stats) {
    sum.samples += t.samples, sum.rays += t.rays, sum.light_rays += t.light_rays;
    sum.iterations += t.iterations, sum.launches += t.launches;
    sum.device_ms = std::max(sum.device_ms, t.device_ms);  // the devices run side by side
    sum.traverse_ms = std::max(sum.traverse_ms, t.traverse_ms), sum.shade_ms = std::max(sum.shade_ms, t.shade_ms);
    sum.nodes_visited += t.nodes_visited, sum.prims_tested += t.prims_tested;
  }
  g->last = sum;
  return YH_OK;
}

yh_status yh_group_stats(const yh_group* g, yh_trace_stats* stats, float* reduce_ms) {
  if (!g) return fail(YH_ERR_INVALID, "yh_group_stats: group is null");
  if (stats) *stats = g->last;
  if (reduce_ms) *reduce_ms = g->reduce_ms;
  return YH_OK;
}

yh_status yh_group_download(yh_group* g, float* render, float* accum, int* samples) {
  YH_REQUIRE_DEVICE();
  if (!g) return fail(YH_ERR_INVALID, "yh_group_download: group is null");
  if (!g->have_state) return fail(YH_ERR_STATE, "yh_group_download: call yh_group_init_state first");
  int prev = 0;
  cudaGetDevice(&prev);
  int    n  = (int)g->m.size();
  size_t np = (size_t)g->width * g->height;
  std::vector<float*> acc(n);
  std::vector<int*>   smp(n);
  for (int i = 0; i < n; i++) {
    yh_status s = yh_state_buffers_dev(g->m[i].state, &acc[i], &smp[i]);
    if (s != YH_OK) return s;
  }
  auto done = [&](yh_status st) {
    cudaSetDevice(prev);
    return st;
  };
#define YH_CUDA_G(call)                                                                                        \
  do {                                                                                                         \
    cudaError_t err_ = (call);                                                                                 \
    if (err_ != cudaSuccess) return done(fail(YH_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(err_))); \
  } while (0)
  YH_CUDA_G(cudaSetDevice(g->root()));
  cudaStream_t rs = g->m[0].stream;
  int grid = (int)((np + 255) / 256);
  if (g->reduce == YH_REDUCE_NCCL) {
    YH_CUDA_G(cudaEventRecord(g->ev0, rs));
    for (int pass = 0; pass < 2; pass++) {
      ncclResult_t r = g_nccl.GroupStart();
      for (int i = 0; i < n && r == ncclSuccess; i++)
        r = pass == 0 ? g_nccl.Reduce(acc[i], g->total_accum, np * 4, ncclFloat, ncclSum, 0, g->m[i].comm, g->m[i].stream)
                      : g_nccl.Reduce(smp[i], g->total_samples, np, ncclInt, ncclSum, 0, g->m[i].comm, g->m[i].stream);
      ncclResult_t e = g_nccl.GroupEnd();
      if (r == ncclSuccess) r = e;
      if (r != ncclSuccess) return done(fail(YH_ERR_CUDA, "ncclReduce failed: %s", g_nccl.GetErrorString(r)));
    }
    YH_CUDA_G(cudaSetDevice(g->root()));
    k_group_resolve<<<grid, 256, 0, rs>>>(g->total_accum, g->total_samples, g->render, (int)np);
    YH_CUDA_G(cudaGetLastError());
    YH_CUDA_G(cudaEventRecord(g->ev1, rs));
    for (int i = 1; i < n; i++) {
      YH_CUDA_G(cudaSetDevice(g->m[i].device));
      YH_CUDA_G(cudaStreamSynchronize(g->m[i].stream));
    }
    YH_CUDA_G(cudaSetDevice(g->root()));
  } else {
    peer_table t{};
    t.n = n;
    YH_CUDA_G(cudaEventRecord(g->ev0, rs));
    for (int i = 0; i < n; i++) {
      if (g->m[i].peer_direct) {
        t.accum[i] = (const float4*)acc[i], t.samples[i] = smp[i];
      } else {  // no peer access: copy first
        YH_CUDA_G(cudaMemcpyPeerAsync(g->m[i].stage_accum, g->root(), acc[i], g->m[i].device, np * 16, rs));
        YH_CUDA_G(cudaMemcpyPeerAsync(g->m[i].stage_samples, g->root(), smp[i], g->m[i].device, np * 4, rs));
        t.accum[i] = g->m[i].stage_accum, t.samples[i] = g->m[i].stage_samples;
      }
    }
    k_group_reduce<<<grid, 256, 0, rs>>>(t, g->total_accum, g->total_samples, g->render, (int)np);
    YH_CUDA_G(cudaGetLastError());
    YH_CUDA_G(cudaEventRecord(g->ev1, rs));
  }
  if (render) YH_CUDA_G(cudaMemcpyAsync(render, g->render, np * 16, cudaMemcpyDeviceToHost, rs));
  if (accum) YH_CUDA_G(cudaMemcpyAsync(accum, g->total_accum, np * 16, cudaMemcpyDeviceToHost, rs));
  if (samples) YH_CUDA_G(cudaMemcpyAsync(samples, g->total_samples, np * 4, cudaMemcpyDeviceToHost, rs));
  YH_CUDA_G(cudaStreamSynchronize(rs));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, g->ev0, g->ev1) == cudaSuccess) g->reduce_ms = ms;
#undef YH_CUDA_G
  return done(YH_OK);
}

}  // extern "C"

// vb_trace.cu -- see vb_trace.cuh.
#include <atomic>
#include <mutex>
#include <vector>

#include "vb_common.cuh"
#include "vb_trace.cuh"

namespace {
const char* kNames[VB_K_COUNT] = {"get_pixel", "get_geometry", "ctx_to_nhwc", "lift_pool_fwd", "lift_plan",
                                  "lift_pool_bwd", "pack_cam_volume", "march_fwd", "bev_fwd", "march_bwd",
                                  "unpack_bev_bwd", "misc"};
std::atomic<long long> g_launches[VB_K_COUNT];
std::atomic<int> g_enabled{0};
std::mutex g_mu;
struct Pair { cudaEvent_t a, b; int id; };
std::vector<Pair> g_pairs;       // completed begin/end pairs awaiting collection
std::vector<cudaEvent_t> g_open[VB_K_COUNT];
}  // namespace

void vb_trace_begin(int id, cudaStream_t st, int launches) {
  g_launches[id].fetch_add(launches, std::memory_order_relaxed);
  if (!g_enabled.load(std::memory_order_relaxed)) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  std::lock_guard<std::mutex> lk(g_mu);
  g_open[id].push_back(e);
}

void vb_trace_end(int id, cudaStream_t st) {
  if (!g_enabled.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_open[id].empty()) return;
  cudaEvent_t a = g_open[id].back();
  g_open[id].pop_back();
  cudaEvent_t b;
  if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return; }
  cudaEventRecord(b, st);
  g_pairs.push_back({a, b, id});
}

extern "C" int vb200_trace_num_kernels(void) { return VB_K_COUNT; }
extern "C" const char* vb200_trace_kernel_name(int id) { return (id >= 0 && id < VB_K_COUNT) ? kNames[id] : ""; }

extern "C" long long vb200_launch_count(int id) {
  if (id >= 0 && id < VB_K_COUNT) return g_launches[id].load();
  long long s = 0;
  for (int i = 0; i < VB_K_COUNT; ++i) s += g_launches[i].load();
  return s;
}

extern "C" int vb200_trace_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& p : g_pairs) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  g_pairs.clear();
  for (auto& v : g_open) { for (auto e : v) cudaEventDestroy(e); v.clear(); }
  g_enabled.store(on ? 1 : 0);
  return VB200_OK;
}

// Synchronises the recorded events. total_ms / launches: arrays of vb200_trace_num_kernels() entries.
extern "C" int vb200_trace_collect(double* total_ms, long long* launches) {
  if (!total_ms || !launches) return VB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < VB_K_COUNT; ++i) { total_ms[i] = 0.0; launches[i] = 0; }
  for (auto& p : g_pairs) {
    if (cudaEventSynchronize(p.b) != cudaSuccess) return VB200_ERR_CUDA;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) != cudaSuccess) return VB200_ERR_CUDA;
    total_ms[p.id] += ms;
    launches[p.id] += 1;
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  g_pairs.clear();
  return VB200_OK;
}

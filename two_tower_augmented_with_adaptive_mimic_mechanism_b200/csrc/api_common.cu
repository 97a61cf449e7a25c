#include "common.cuh"
#include <cstring>
#include <atomic>

namespace ttam {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace ttam

extern "C" const char* ttam_last_error(void) { return ttam::g_err; }
extern "C" int ttam_version(void) { return 100; }
extern "C" int64_t ttam_launch_count(void) { return (int64_t)ttam::g_launches.load(std::memory_order_relaxed); }

__global__ void ttam_probe_kernel(int* out) { *out = 100; }

extern "C" int ttam_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    ttam::set_error("no CUDA device");
    return 0;
  }
  cudaFuncAttributes attr;
  if (cudaFuncGetAttributes(&attr, ttam_probe_kernel) != cudaSuccess) {
    cudaGetLastError();
    ttam::set_error("libttam.so holds sm_100a code only; this device cannot run it");
    return 0;
  }
  return 1;
}

// Shared helpers for libttam.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/ttam.h"

namespace ttam {

void set_error(const char* fmt, ...);
void count_launch();  // every kernel launch of this library is counted (ttam_launch_count)

#define TTAM_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::ttam::set_error(__VA_ARGS__);             \
      return TTAM_EINVAL;                         \
    }                                             \
  } while (0)

#define TTAM_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::ttam::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return TTAM_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define TTAM_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    ::ttam::count_launch();                                                               \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      ::ttam::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return TTAM_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// streaming (read-once) 16-byte load that does not allocate in L1
__device__ __forceinline__ float4 ld_f4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 (counter-based RNG for dropout)
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0u, c3 = 0u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Bernoulli keep decision of element `e` of a dropout mask: uniform from one 32-bit lane of Philox(seed, e/4)
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t e, float p) {
  const uint4 r = philox4x32(seed, e >> 2);
  const uint32_t bits = (e & 3) == 0 ? r.x : (e & 3) == 1 ? r.y : (e & 3) == 2 ? r.z : r.w;
  const float u = (float)(bits >> 8) * (1.0f / 16777216.0f);
  return u >= p;
}

}  // namespace ttam

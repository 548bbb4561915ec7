// Shared host/device helpers for libnasr_ctc.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "nasr_ctc.h"

namespace nasr {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define NASR_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      nasr::set_error(__VA_ARGS__);          \
      return NASR_ERR_INVALID_ARGUMENT;      \
    }                                        \
  } while (0)

#define NASR_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      nasr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__,   \
                      __LINE__);                                                          \
      return NASR_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

// SM count of the current device (cached per device: a process may drive several)
inline cudaError_t device_sm_count(int* out) {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int n = (dev >= 0 && dev < 64) ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (!n) {
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  }
  *out = n;
  return cudaSuccess;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: set it once per (kernel, device).  (Setting it at every
// launch costs ~0.2 ms of host time per call -- measured: it made nasr_edit_distance host-bound.)
template <auto Kernel>
inline cudaError_t ensure_max_dynamic_smem(int bytes) {
  static std::atomic<int> done[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && done[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev].store(bytes, std::memory_order_release);
  return e;
}

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace nasr

// Shared helpers for the speecht_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/speecht_b200.h"

#define ST_API extern "C" __attribute__((visibility("default")))

void st_set_error(const char* fmt, ...);

#define ST_CHECK_ARG(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      st_set_error(__VA_ARGS__);               \
      return ST_ERR_INVALID_ARG;               \
    }                                          \
  } while (0)

#define ST_CUDA_LAUNCH_CHECK(name)                                                  \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      st_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__));    \
      return ST_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define ST_CUDA_CALL(expr)                                                          \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      st_set_error("%s failed: %s", #expr, cudaGetErrorString(e__));                \
      return ST_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

static inline cudaStream_t st_cu(st_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

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

static inline int st_num_sms() {
  // per device: one process may drive several GPUs
  static int sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (sms[dev] == 0) {
    if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

// Shared helpers for libe4s_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/e4s_b200.h"

namespace e4s {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(E4S_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return E4S_OK;
}

#define E4S_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return ::e4s::fail(E4S_ERR_ARG, __VA_ARGS__);   \
  } while (0)

// Per-device one-time state (cudaFuncSetAttribute is per device; a process may drive several GPUs): index of the calling thread's
// current device, clamped to the table size.
constexpr int E4S_MAX_DEVICES = 64;
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < E4S_MAX_DEVICES ? dev : E4S_MAX_DEVICES - 1;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// torch legacy 'nearest' source index: min(floor(dst * (float)in/out), in-1)
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
  if (in_size == out_size) return dst;
  float scale = (float)in_size / (float)out_size;
  int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}

}  // namespace e4s

// Library-wide state: thread-local error text, launch counter, device query.
#include <stdarg.h>

#include "common.cuh"

namespace e4s {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace e4s

extern "C" const char* e4s_last_error(void) { return e4s::g_err; }
extern "C" int e4s_sizeof_conv(void) { return (int)sizeof(E4SConv); }
extern "C" int64_t e4s_launch_count(void) { return e4s::g_launches.load(); }

extern "C" int e4s_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e4s::fail(E4S_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return e4s::fail(E4S_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return E4S_OK;
}

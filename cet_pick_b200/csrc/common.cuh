// Shared helpers for libcetpick_sm100a.so (sm_100a only; no multi-arch dispatch).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "../../include/cetpick.h"

namespace cetpick {

extern thread_local int64_t g_launches;        // kernels enqueued by the current API call
extern thread_local std::string g_cuda_err;    // text of the last CUDA failure on this thread

inline int cuda_fail(cudaError_t e, const char* what) {
  g_cuda_err = std::string(what) + ": " + cudaGetErrorString(e);
  return CETPICK_ERR_CUDA;
}

#define CETPICK_CUDA(call)                                            \
  do {                                                                \
    cudaError_t e__ = (call);                                         \
    if (e__ != cudaSuccess) return ::cetpick::cuda_fail(e__, #call);  \
  } while (0)

#define CETPICK_LAUNCH_CHECK()                                                 \
  do {                                                                         \
    ++::cetpick::g_launches;                                                   \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess) return ::cetpick::cuda_fail(e__, "kernel launch"); \
  } while (0)

inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  return dev;
}

inline int num_sms() {
  static int n[64] = {};                       // per device: one process may hold plans on several GPUs
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

// Function attributes (opt-in shared memory size) belong to a device context: set them once PER DEVICE.
struct DeviceOnce {
  unsigned long long done = 0;                 // bit d: the guarded block already ran on device d
  bool first() {
    const unsigned long long bit = 1ull << current_device();
    if (done & bit) return false;
    done |= bit;
    return true;
  }
};

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// float -> bf16 bits, round to nearest even (host side weight packing)
inline uint16_t f2bf_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Kernel launch with explicit configuration (decode.cu).  Programmatic dependent launch was tried for
// the ~19 short kernels of one decode and measured SLOWER on B200 (0.70 -> 1.0-1.18 ms at 512x1024x1024,
// profiles/r1g): dependents that become resident early and park in griddepcontrol.wait take issue and
// memory slots from the streaming kernel; plain stream order is kept.
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

}  // namespace cetpick

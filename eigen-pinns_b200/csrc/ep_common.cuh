// Shared helpers for the eigenpinns_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/eigenpinns_b200.h"

namespace ep {

void set_error(const char* fmt, ...);

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return EP_ERR_CUDA;
}

#define EP_CUDA_CHECK(call)                                   \
  do {                                                        \
    cudaError_t _e = (call);                                  \
    if (_e != cudaSuccess) return ep::cuda_fail(_e, #call);   \
  } while (0)

#define EP_LAUNCH_CHECK(name)                                 \
  do {                                                        \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) return ep::cuda_fail(_e, name);    \
  } while (0)

#define EP_REQUIRE(cond, msg)                                 \
  do {                                                        \
    if (!(cond)) { ep::set_error("%s: %s", __func__, msg); return EP_ERR_INVALID; } \
  } while (0)

inline cudaStream_t as_stream(ep_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();   // cached
int tune_flag(int key);          // experiment switches set through ep_tune_set (capi.cu)
void set_tune_flag(int key, int value);

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

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

}  // namespace ep

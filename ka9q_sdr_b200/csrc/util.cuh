// Small shared helpers: warp/block reductions (shuffle based), error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

namespace k9 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// thread-local last error text, exported through ka9q_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

#define K9_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      k9::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return -1;                                                                        \
    }                                                                                   \
  } while (0)

}  // namespace k9

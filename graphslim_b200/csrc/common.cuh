// Shared helpers for the graphslim_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/graphslim_b200.h"

namespace gs {

extern std::atomic<int64_t> g_launches;
void set_error(const char* what, cudaError_t e);
void set_error_msg(const char* what);

inline int finish_launch(const char* name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error(name, e);
    return (int)e;
  }
  return GS_OK;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

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
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// chunk index of flat row k for np.array_split style boundaries (nchunk is tiny)
__device__ __forceinline__ int chunk_of(int64_t k, int nchunk, const int64_t* __restrict__ off) {
  int c = 0;
  while (c + 1 < nchunk && k >= off[c + 1]) ++c;
  return c;
}

}  // namespace gs

#define GS_REQUIRE(cond)                                   \
  do {                                                     \
    if (!(cond)) {                                         \
      gs::set_error_msg("invalid argument: " #cond);       \
      return GS_EINVAL;                                    \
    }                                                      \
  } while (0)

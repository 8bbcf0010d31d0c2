// Shared helpers for libguidedattn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "guided_attn.h"

namespace ga {

// ---- error reporting (thread-local string behind ga_last_error) -------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define GA_CHECK_ARG(cond, ...)                        \
  do {                                                 \
    if (!(cond)) return ga::fail(GA_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define GA_CHECK_ALIGN(ptr, bytes, name)                                                              \
  do {                                                                                                \
    if ((reinterpret_cast<uintptr_t>(ptr) % (bytes)) != 0)                                            \
      return ga::fail(GA_ERR_ALIGNMENT, "%s must be %d-byte aligned", name, (int)(bytes));           \
  } while (0)

// Checks the launch only (never synchronises).
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(GA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return GA_OK;
}

inline size_t dtype_size(int dtype) { return dtype == GA_F32 ? 4 : 2; }

// ---- device helpers -----------------------------------------------------------------------------------------------
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

// 32-bit "word" of packed elements: 1 float or 2 halves / bfloat16s.
template <typename T> struct Word;
template <> struct Word<float> {
  static constexpr int E = 1;
  static __device__ __forceinline__ float2 unpack(uint32_t w) { return make_float2(__uint_as_float(w), 0.f); }
  static __device__ __forceinline__ uint32_t pack(float a, float) { return __float_as_uint(a); }
};
template <> struct Word<__half> {
  static constexpr int E = 2;
  static __device__ __forceinline__ float2 unpack(uint32_t w) {
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <> struct Word<__nv_bfloat16> {
  static constexpr int E = 2;
  static __device__ __forceinline__ float2 unpack(uint32_t w) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

}  // namespace ga

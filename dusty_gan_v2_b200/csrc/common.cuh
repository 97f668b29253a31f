// Shared helpers for libdusty_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dusty_b200.h"

namespace dusty {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int num_sms();

#define DUSTY_CHECK_ARG(cond, msg)                                   \
  do {                                                               \
    if (!(cond)) {                                                   \
      dusty::set_error("%s: %s", __func__, msg);                     \
      return DUSTY_EINVAL;                                           \
    }                                                                \
  } while (0)

#define DUSTY_LAUNCH_CHECK()                                                        \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      dusty::set_error("%s: CUDA launch failed: %s", __func__, cudaGetErrorString(e__)); \
      return DUSTY_ECUDA;                                                           \
    }                                                                               \
    dusty::count_launch();                                                          \
  } while (0)

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// 16-byte vector of T
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ float get(int i) const { return (&raw.x)[i]; }
  __device__ __forceinline__ void set(int i, float v) { (&raw.x)[i] = v; }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ float get(int i) const {
    uint32_t w = (&raw.x)[i >> 1];
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
  }
  __device__ __forceinline__ void set(int i, float v) {
    uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
    uint32_t &w = (&raw.x)[i >> 1];
    w = (i & 1) ? ((w & 0x0000ffffu) | (b << 16)) : ((w & 0xffff0000u) | b);
  }
};

// Pair access (elements 2i, 2i+1) for the packed fp32 pipe of sm_100 (FFMA2 / FADD2 / FMUL2: two
// fp32 operations per issued instruction).  The stencils are issue-bound, not HBM-bound, when
// written one element at a time (~27 instructions per bf16 output element in blur4_cl).
__device__ __forceinline__ float2 get2(const Vec16<float> &v, int i) {
  return i ? make_float2(v.raw.z, v.raw.w) : make_float2(v.raw.x, v.raw.y);
}
__device__ __forceinline__ void set2(Vec16<float> &v, int i, float2 p) {
  if (i) { v.raw.z = p.x; v.raw.w = p.y; } else { v.raw.x = p.x; v.raw.y = p.y; }
}
__device__ __forceinline__ float2 get2(const Vec16<__nv_bfloat16> &v, int i) {
  const uint32_t w = (&v.raw.x)[i];
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ void set2(Vec16<__nv_bfloat16> &v, int i, float2 p) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(p.x, p.y);      // one F2FP for the pair
  (&v.raw.x)[i] = *reinterpret_cast<const uint32_t *>(&b);
}
__device__ __forceinline__ float2 fma2(float k, float2 v, float2 acc) {
  return __ffma2_rn(make_float2(k, k), v, acc);
}
__device__ __forceinline__ float2 mul2(float k, float2 v) { return __fmul2_rn(make_float2(k, k), v); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

// 8-byte vector of T (register-lean variant for the sliding-window stencils)
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static constexpr int N = 2;
  float2 raw;
  __device__ __forceinline__ float get(int i) const { return i ? raw.y : raw.x; }
  __device__ __forceinline__ void set(int i, float v) { if (i) raw.y = v; else raw.x = v; }
};
template <> struct Vec8<__nv_bfloat16> {
  static constexpr int N = 4;
  uint2 raw;
  __device__ __forceinline__ float get(int i) const {
    uint32_t w = (i >> 1) ? raw.y : raw.x;
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
  }
  __device__ __forceinline__ void set(int i, float v) {
    uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
    uint32_t &w = (i >> 1) ? raw.y : raw.x;
    w = (i & 1) ? ((w & 0x0000ffffu) | (b << 16)) : ((w & 0xffff0000u) | b);
  }
};
template <typename T> __device__ __forceinline__ Vec8<T> ld8(const T *p) {
  Vec8<T> v;
  v.raw = *reinterpret_cast<const decltype(v.raw) *>(p);
  return v;
}
template <typename T> __device__ __forceinline__ void st8(T *p, const Vec8<T> &v) {
  *reinterpret_cast<decltype(v.raw) *>(p) = v.raw;
}

template <typename T> __device__ __forceinline__ Vec16<T> ld16(const T *p) {
  Vec16<T> v;
  v.raw = *reinterpret_cast<const decltype(v.raw) *>(p);
  return v;
}
template <typename T> __device__ __forceinline__ void st16(T *p, const Vec16<T> &v) {
  *reinterpret_cast<decltype(v.raw) *>(p) = v.raw;
}
// streaming variants (read-once / write-once data: keep L2 for reusable tensors)
template <typename T> __device__ __forceinline__ Vec16<T> ld16_stream(const T *p) {
  Vec16<T> v;
  v.raw = __ldcs(reinterpret_cast<const decltype(v.raw) *>(p));
  return v;
}
template <typename T> __device__ __forceinline__ void st16_stream(T *p, const Vec16<T> &v) {
  __stcs(reinterpret_cast<decltype(v.raw) *>(p), v.raw);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float *red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) v = warp_sum(v);
  return v;
}

__device__ __forceinline__ uint32_t smem_addr_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace dusty

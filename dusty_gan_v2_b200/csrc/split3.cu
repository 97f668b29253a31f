// g1: fp32 mode on the tensor cores.  An fp32 contraction sum_k a_k * b_k is evaluated on the
// bf16 UMMA path (kind::f16, fp32 accumulation in TMEM) by splitting every operand into two
// bf16 terms, a = a_hi + a_lo (+ a residual below 2^-17 |a|), and CONCATENATING the three
// significant products along the contracted axis:
//     sum_k a_k b_k  ~=  sum_k (a_hi b_hi + a_hi b_lo + a_lo b_hi)
//                     =  [a_hi | a_hi | a_lo] . [b_hi | b_lo | b_hi]          (3K-long dot product)
// so the existing tcgen05 kernels run unchanged on operands with a three times longer K axis and
// an fp32 epilogue; the dropped lo*lo term and the residuals are ~2^-16 relative per product,
// two orders below the rtol 1e-3 the fp32 mode is held to.  (kind::tf32 was the other candidate:
// 3xTF32 needs the same three products, and tf32 operands cannot be MN-major under the 128-byte
// swizzle -- linear_tc.cu -- which the generator's pixel-contiguous activations are.)
//
// dusty_split_bf16x3: src fp32 viewed as [outer][K][inner] through element strides; dst bf16
// [outer][3K][inner] through its own strides; part p of the K axis holds
//     pattern 0 (the "a" side): hi, hi, lo         pattern 1 (the "b" side): hi, lo, hi
// With K = 1 and inner = numel the three parts are three whole-tensor copies (batch
// concatenation for the weight-gradient kernels, whose contracted axis is batch x pixels).
// Replaces nothing in the reference (its fp32 path is cuDNN / cuBLAS fp32); serves
// gans/models/ops/common.py:187-210 and ops/style.py:68-126 in fp32 mode.
#include "common.cuh"

namespace dusty {
namespace {

struct SplitArgs {
  long long outer, K, inner;
  long long s_o, s_k, s_i;       // source strides (elements)
  long long d_o, d_k, d_i;       // destination strides (elements)
  int pattern;
  int k_fast;                    // 1: thread index runs over k fastest (source k-contiguous)
};

__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float *__restrict__ src,
                                                           __nv_bfloat16 *__restrict__ dst,
                                                           const SplitArgs a, long long total) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total) return;
  long long o, k, i;
  if (a.k_fast) {
    k = t % a.K;
    const long long r = t / a.K;
    i = r % a.inner;
    o = r / a.inner;
  } else {
    i = t % a.inner;
    const long long r = t / a.inner;
    k = r % a.K;
    o = r / a.K;
  }
  const float v = __ldg(src + o * a.s_o + k * a.s_k + i * a.s_i);
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16 *d = dst + o * a.d_o + k * a.d_k + i * a.d_i;
  const long long part = a.K * a.d_k;
  d[0] = hi;
  d[part] = a.pattern == 0 ? hi : lo;
  d[2 * part] = a.pattern == 0 ? lo : hi;
}

// 8 consecutive elements of the contiguous axis per thread (two float4 loads, three 16-byte
// stores): `vec_axis` 0 = the k axis is contiguous in source and destination, 1 = the inner axis
__global__ void __launch_bounds__(256) split_bf16x3_vec_kernel(const float *__restrict__ src,
                                                               __nv_bfloat16 *__restrict__ dst,
                                                               const SplitArgs a, long long total8,
                                                               int vec_axis) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= total8) return;
  long long o, k, i;
  if (vec_axis == 0) {
    const long long K8 = a.K >> 3;
    k = (t % K8) << 3;
    const long long r = t / K8;
    i = r % a.inner;
    o = r / a.inner;
  } else {
    const long long I8 = a.inner >> 3;
    i = (t % I8) << 3;
    const long long r = t / I8;
    k = r % a.K;
    o = r / a.K;
  }
  const float *s = src + o * a.s_o + k * a.s_k + i * a.s_i;
  const float4 v0 = __ldg(reinterpret_cast<const float4 *>(s));
  const float4 v1 = __ldg(reinterpret_cast<const float4 *>(s) + 1);
  const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
    const __nv_bfloat162 hp = __halves2bfloat162(h0, h1);
    const __nv_bfloat162 lp = __floats2bfloat162_rn(v[2 * j] - __bfloat162float(h0), v[2 * j + 1] - __bfloat162float(h1));
    hi[j] = *reinterpret_cast<const uint32_t *>(&hp);
    lo[j] = *reinterpret_cast<const uint32_t *>(&lp);
  }
  const uint4 H = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  const uint4 L = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  __nv_bfloat16 *d = dst + o * a.d_o + k * a.d_k + i * a.d_i;
  const long long part = a.K * a.d_k;
  *reinterpret_cast<uint4 *>(d) = H;
  *reinterpret_cast<uint4 *>(d + part) = a.pattern == 0 ? H : L;
  *reinterpret_cast<uint4 *>(d + 2 * part) = a.pattern == 0 ? L : H;
}

}  // namespace
}  // namespace dusty

using namespace dusty;

extern "C" int dusty_split_bf16x3(const float *src, void *dst, long long outer, long long K,
                                  long long inner, const long long *src_strides,
                                  const long long *dst_strides, int pattern, void *stream) {
  DUSTY_CHECK_ARG(src && dst && src_strides && dst_strides, "null pointer");
  DUSTY_CHECK_ARG(outer >= 1 && K >= 1 && inner >= 1, "empty tensor");
  DUSTY_CHECK_ARG(pattern == 0 || pattern == 1, "pattern: 0 = hi,hi,lo  1 = hi,lo,hi");
  SplitArgs a;
  a.outer = outer; a.K = K; a.inner = inner;
  a.s_o = src_strides[0]; a.s_k = src_strides[1]; a.s_i = src_strides[2];
  a.d_o = dst_strides[0]; a.d_k = dst_strides[1]; a.d_i = dst_strides[2];
  a.pattern = pattern;
  a.k_fast = (a.s_k == 1 && K > 1) ? 1 : 0;
  const long long total = outer * K * inner;
  // vector path: a contiguous axis of a multiple of 8 elements, every other offset 16-byte aligned
  const bool al = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  const long long part = K * a.d_k;
  int vec_axis = -1;
  if (al && a.s_k == 1 && a.d_k == 1 && K % 8 == 0 && a.s_o % 4 == 0 && a.d_o % 8 == 0 &&
      (inner == 1 || (a.s_i % 4 == 0 && a.d_i % 8 == 0)))
    vec_axis = 0;
  else if (al && a.s_i == 1 && a.d_i == 1 && inner % 8 == 0 && a.s_o % 4 == 0 && a.d_o % 8 == 0 &&
           part % 8 == 0 && (K == 1 || (a.s_k % 4 == 0 && a.d_k % 8 == 0)))
    vec_axis = 1;
  if (vec_axis >= 0) {
    const long long total8 = total / 8;
    const long long blocks8 = (total8 + 255) / 256;
    DUSTY_CHECK_ARG(blocks8 <= 0x7fffffff, "tensor too large");
    split_bf16x3_vec_kernel<<<(unsigned)blocks8, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16 *)dst, a,
                                                                                 total8, vec_axis);
    DUSTY_LAUNCH_CHECK();
    return DUSTY_OK;
  }
  const long long blocks = (total + 255) / 256;
  DUSTY_CHECK_ARG(blocks <= 0x7fffffff, "tensor too large");
  split_bf16x3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16 *)dst, a, total);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

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
  const long long blocks = (total + 255) / 256;
  DUSTY_CHECK_ARG(blocks <= 0x7fffffff, "tensor too large");
  split_bf16x3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16 *)dst, a, total);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

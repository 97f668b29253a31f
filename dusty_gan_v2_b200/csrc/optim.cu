// f2: multi-tensor Adam with the generator's EMA folded in, and multi-tensor copy / scale
// (gradient bucket packing, EMA buffer copies).  Replaces torch.optim.Adam(fused=True) +
// torch._foreach_lerp_ / _foreach_copy_ behind gans/trainer.py:30-41 (ema_inplace),
// 128-171 (optimisers).  One launch per group of up to 32 tensors: a CTA looks its tensor up
// in a chunk-prefix table carried in the kernel parameters.
#include "common.cuh"

namespace dusty {
namespace {

constexpr int kMaxTensors = 32;
constexpr int kChunk = 4096;            // elements per CTA (256 threads x 4 float4)

struct MultiArgs {
  void *a[kMaxTensors];                 // param        | dst
  const void *b[kMaxTensors];           // grad         | src
  void *c[kMaxTensors];                 // exp_avg
  void *d[kMaxTensors];                 // exp_avg_sq
  void *e[kMaxTensors];                 // ema param or NULL
  long long n[kMaxTensors];
  int first_chunk[kMaxTensors + 1];
  int count;
};

__device__ __forceinline__ int find_tensor(const MultiArgs &m, int chunk) {
  int lo = 0, hi = m.count - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (m.first_chunk[mid] <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Adam as torch.optim.Adam (no weight decay, no amsgrad):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// then, when an EMA tensor is attached, ema += ema_w * (p - ema)   (lerp towards the NEW weight)
__global__ void __launch_bounds__(256)
multi_adam_kernel(const __grid_constant__ MultiArgs m, float lr, float b1, float b2, float eps,
                  float bc1, float rsqrt_bc2, float ema_w, float grad_scale) {
  const int t = find_tensor(m, blockIdx.x);
  const long long base = (long long)(blockIdx.x - m.first_chunk[t]) * kChunk;
  const long long n = m.n[t];
  float *p = (float *)m.a[t];
  const float *g = (const float *)m.b[t];
  float *ea = (float *)m.c[t], *es = (float *)m.d[t], *em = (float *)m.e[t];
  const float step = lr / bc1;
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)ea | (uintptr_t)es | (uintptr_t)em) & 15) == 0);
#pragma unroll
  for (int i = 0; i < kChunk / (256 * 4); ++i) {
    const long long o = base + (long long)(i * 256 + threadIdx.x) * 4;
    if (o >= n) break;
    if (vec && o + 4 <= n) {
      float4 pv = *reinterpret_cast<float4 *>(p + o);
      const float4 gv = *reinterpret_cast<const float4 *>(g + o);
      float4 av = *reinterpret_cast<float4 *>(ea + o), sv = *reinterpret_cast<float4 *>(es + o);
      float *pp = &pv.x, *aa = &av.x, *ss = &sv.x;
      const float *gg = &gv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = gg[k] * grad_scale;
        aa[k] = b1 * aa[k] + (1.f - b1) * gk;
        ss[k] = b2 * ss[k] + (1.f - b2) * gk * gk;
        pp[k] -= step * aa[k] / (sqrtf(ss[k]) * rsqrt_bc2 + eps);
      }
      *reinterpret_cast<float4 *>(p + o) = pv;
      *reinterpret_cast<float4 *>(ea + o) = av;
      *reinterpret_cast<float4 *>(es + o) = sv;
      if (em) {
        float4 ev = *reinterpret_cast<float4 *>(em + o);
        ev.x += ema_w * (pv.x - ev.x); ev.y += ema_w * (pv.y - ev.y);
        ev.z += ema_w * (pv.z - ev.z); ev.w += ema_w * (pv.w - ev.w);
        *reinterpret_cast<float4 *>(em + o) = ev;
      }
    } else {
      for (long long q = o; q < n && q < o + 4; ++q) {
        const float gk = g[q] * grad_scale;
        const float a = b1 * ea[q] + (1.f - b1) * gk;
        const float s = b2 * es[q] + (1.f - b2) * gk * gk;
        const float pn = p[q] - step * a / (sqrtf(s) * rsqrt_bc2 + eps);
        ea[q] = a; es[q] = s; p[q] = pn;
        if (em) em[q] += ema_w * (pn - em[q]);
      }
    }
  }
}

// dst = src * scale (fp32), tensor by tensor: packing gradients into a flat bucket and back,
// EMA buffer copies
__global__ void __launch_bounds__(256)
multi_copy_kernel(const __grid_constant__ MultiArgs m, float scale) {
  const int t = find_tensor(m, blockIdx.x);
  const long long base = (long long)(blockIdx.x - m.first_chunk[t]) * kChunk;
  const long long n = m.n[t];
  float *dst = (float *)m.a[t];
  const float *src = (const float *)m.b[t];
  const bool vec = ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0);
#pragma unroll
  for (int i = 0; i < kChunk / (256 * 4); ++i) {
    const long long o = base + (long long)(i * 256 + threadIdx.x) * 4;
    if (o >= n) break;
    if (vec && o + 4 <= n) {
      float4 v = *reinterpret_cast<const float4 *>(src + o);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      *reinterpret_cast<float4 *>(dst + o) = v;
    } else {
      for (long long q = o; q < n && q < o + 4; ++q) dst[q] = src[q] * scale;
    }
  }
}

template <typename Launch>
int for_groups(int count, const long long *n, Launch launch) {
  int at = 0;
  while (at < count) {
    MultiArgs m;
    m.count = 0;
    int chunks = 0;
    while (at < count && m.count < kMaxTensors) {
      if (n[at] > 0) {
        m.n[m.count] = n[at];
        m.first_chunk[m.count] = chunks;
        chunks += (int)((n[at] + kChunk - 1) / kChunk);
        launch(m, m.count, at, false);
        ++m.count;
      }
      ++at;
    }
    if (m.count == 0) continue;
    m.first_chunk[m.count] = chunks;
    for (int i = m.count; i < kMaxTensors; ++i) {
      m.a[i] = nullptr; m.b[i] = nullptr; m.c[i] = nullptr; m.d[i] = nullptr; m.e[i] = nullptr; m.n[i] = 0;
      m.first_chunk[i + 1] = chunks;
    }
    launch(m, chunks, -1, true);
  }
  return 0;
}

}  // namespace
}  // namespace dusty

using namespace dusty;

extern "C" int dusty_multi_adam(void *const *params, const void *const *grads, void *const *exp_avg,
                                void *const *exp_avg_sq, void *const *ema, const long long *numel,
                                int count, float lr, float beta1, float beta2, float eps, int step,
                                float ema_weight, float grad_scale, void *stream) {
  DUSTY_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel, "null pointer");
  DUSTY_CHECK_ARG(count >= 0 && step >= 1, "bad count / step");
  cudaStream_t st = (cudaStream_t)stream;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  bool bad = false;
  for_groups(count, numel, [&](MultiArgs &m, int slot_or_chunks, int src, bool go) {
    if (!go) {
      m.a[slot_or_chunks] = params[src]; m.b[slot_or_chunks] = grads[src];
      m.c[slot_or_chunks] = exp_avg[src]; m.d[slot_or_chunks] = exp_avg_sq[src];
      m.e[slot_or_chunks] = ema ? ema[src] : nullptr;
      if (!params[src] || !grads[src] || !exp_avg[src] || !exp_avg_sq[src]) bad = true;
      return;
    }
    if (bad) return;
    multi_adam_kernel<<<slot_or_chunks, 256, 0, st>>>(m, lr, beta1, beta2, eps, (float)bc1,
                                                      (float)(1.0 / sqrt(bc2)), ema_weight, grad_scale);
    count_launch();
  });
  DUSTY_CHECK_ARG(!bad, "null tensor pointer");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("dusty_multi_adam: CUDA launch failed: %s", cudaGetErrorString(e));
    return DUSTY_ECUDA;
  }
  return DUSTY_OK;
}

extern "C" int dusty_multi_copy(void *const *dst, const void *const *src, const long long *numel,
                                int count, float scale, void *stream) {
  DUSTY_CHECK_ARG(dst && src && numel && count >= 0, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  bool bad = false;
  for_groups(count, numel, [&](MultiArgs &m, int slot_or_chunks, int s, bool go) {
    if (!go) {
      m.a[slot_or_chunks] = dst[s]; m.b[slot_or_chunks] = src[s];
      m.c[slot_or_chunks] = nullptr; m.d[slot_or_chunks] = nullptr; m.e[slot_or_chunks] = nullptr;
      if (!dst[s] || !src[s]) bad = true;
      return;
    }
    if (bad) return;
    multi_copy_kernel<<<slot_or_chunks, 256, 0, st>>>(m, scale);
    count_launch();
  });
  DUSTY_CHECK_ARG(!bad, "null tensor pointer");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("dusty_multi_copy: CUDA launch failed: %s", cudaGetErrorString(e));
    return DUSTY_ECUDA;
  }
  return DUSTY_OK;
}

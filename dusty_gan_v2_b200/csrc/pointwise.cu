// Memory-bound pointwise / reduction kernels of the hot path:
//   a10 Gumbel-sigmoid raydrop (fwd/bwd), a14 range image -> point cloud (+ valid count),
//   a12 minibatch stddev (fwd/bwd), a13 row-wise sum of squares (R1, EMA statistic),
//   a7  circular fractional un-shift.
// All use coalesced (vectorised where layout allows) HBM access and warp-shuffle
// reductions with one atomic per CTA.
#include <cfloat>

#include "common.cuh"

namespace dusty {

// ------------------------------------------------------------------ a10 raydrop
// eps = FLT_EPSILON, tiny = FLT_MIN: torch.distributions clamp_probs / SigmoidTransform.
__global__ void __launch_bounds__(256)
raydrop_fwd_kernel(const float *__restrict__ logit, const float *__restrict__ image,
                   const float *__restrict__ u, float *__restrict__ mask,
                   float *__restrict__ image_out, float *__restrict__ dsoft,
                   int *__restrict__ count, int64_t n, float rconst, float inv_temp) {
  __shared__ float red[32];
  const float eps = FLT_EPSILON;
  float local = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float l = logit[i];
    float p = 1.f / (1.f + expf(-l));
    const bool p_clamped = (p < eps) || (p > 1.f - eps);
    p = fminf(fmaxf(p, eps), 1.f - eps);
    const float uu = fminf(fmaxf(u[i], eps), 1.f - eps);
    const float y = ((logf(uu) - log1pf(-uu)) + logf(p) - log1pf(-p)) * inv_temp;
    float s = 1.f / (1.f + expf(-y));
    const bool s_clamped = (s < FLT_MIN) || (s > 1.f - eps);
    s = fminf(fmaxf(s, FLT_MIN), 1.f - eps);
    const float hard = (s > 0.5f) ? 1.f : 0.f;
    mask[i] = hard;
    const float im = image[i];
    // torch.lerp(image, const, 1 - mask) with an exactly 0/1 weight
    image_out[i] = (hard != 0.f) ? im : rconst;
    // d soft / d logit: sigmoid'(y)/T * d logit(p)/d l  (= 1 unless a clamp is active)
    dsoft[i] = (p_clamped || s_clamped) ? 0.f : s * (1.f - s) * inv_temp;
    local += hard;
  }
  if (count != nullptr) {
    local = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(count, (int)(local + 0.5f));
  }
}

__global__ void __launch_bounds__(256)
raydrop_bwd_kernel(const float *__restrict__ g_out, const float *__restrict__ g_mask,
                   const float *__restrict__ image, const float *__restrict__ mask,
                   const float *__restrict__ dsoft, float *__restrict__ g_logit,
                   float *__restrict__ g_image, int64_t n, float rconst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float go = g_out ? g_out[i] : 0.f;
    float gm = go * (image[i] - rconst);  // d image_out / d mask = image - const
    if (g_mask) gm += g_mask[i];
    g_logit[i] = gm * dsoft[i];
    g_image[i] = go * mask[i];
  }
}

// ------------------------------------------------------------------ a14 projection
template <int LAYOUT>
__global__ void __launch_bounds__(256)
project_kernel(const float *__restrict__ x, const float *__restrict__ trig, float *__restrict__ out,
               unsigned long long *__restrict__ valid_count, int64_t HW, int64_t total,
               float min_depth, float lo_thr, float hi_thr, float tol) {
  __shared__ float red[32];
  float local = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, p = i - b * HW;
    const float v = x[i];
    const float inv = __fdiv_rn(v, min_depth);
    const bool ok = (v > tol) && (inv >= lo_thr) && (inv <= hi_thr) && (inv > 0.f);
    const float valid = ok ? 1.f : 0.f;
    const float depth = __fmul_rn(__fdiv_rn(1.f, __fadd_rn(inv, tol)), valid);
    const float ce = __ldg(trig + p), se = __ldg(trig + HW + p);
    const float ca = __ldg(trig + 2 * HW + p), sa = __ldg(trig + 3 * HW + p);
    const float dc = __fmul_rn(depth, ce);
    const float px = __fmul_rn(dc, ca), py = __fmul_rn(dc, sa), pz = __fmul_rn(depth, se);
    if (LAYOUT == 0) {
      float *o = out + b * 3 * HW + p;
      o[0] = px; o[HW] = py; o[2 * HW] = pz;
    } else {
      float *o = out + i * 3;
      o[0] = px; o[1] = py; o[2] = pz;
    }
    local += valid;
  }
  if (valid_count != nullptr) {
    local = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(valid_count, (unsigned long long)(local + 0.5f));
  }
}

// ------------------------------------------------------------------ a12 minibatch stddev
// x viewed as [G, M, CHW] with M = B/G; stat[m] = mean_chw sqrt(var_g + alpha).
template <typename T>
__global__ void __launch_bounds__(256)
mbstd_stat_kernel(const T *__restrict__ x, float *__restrict__ stat, int G, int M, int64_t CHW,
                  float alpha) {
  __shared__ float red[32];
  const int m = blockIdx.y;
  float local = 0.f;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < CHW;
       j += (int64_t)gridDim.x * blockDim.x) {
    float mean = 0.f;
    for (int g = 0; g < G; ++g) mean += to_f(x[((int64_t)g * M + m) * CHW + j]);
    mean /= (float)G;
    float var = 0.f;
    for (int g = 0; g < G; ++g) {
      const float d = to_f(x[((int64_t)g * M + m) * CHW + j]) - mean;
      var = fmaf(d, d, var);
    }
    local += sqrtf(var / (float)G + alpha);
  }
  local = block_sum(local, red);
  if (threadIdx.x == 0) atomicAdd(stat + m, local / (float)CHW);
}

template <typename T>
__global__ void __launch_bounds__(256)
mbstd_write_kernel(const T *__restrict__ x, const float *__restrict__ stat, T *__restrict__ y,
                   int B, int C, int64_t HW, int M) {
  const int64_t total = (int64_t)B * (C + 1) * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / ((C + 1) * HW);
    const int64_t r = i - b * (C + 1) * HW;
    y[i] = (r < C * HW) ? x[b * C * HW + r] : from_f<T>(stat[b % M]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
mbstd_dstat_kernel(const T *__restrict__ dy, float *__restrict__ dstat, int B, int C, int64_t HW,
                   int M) {
  __shared__ float red[32];
  const int m = blockIdx.x;
  float local = 0.f;
  const int G = B / M;
  for (int64_t j = threadIdx.x; j < (int64_t)G * HW; j += blockDim.x) {
    const int g = (int)(j / HW);
    const int64_t p = j - (int64_t)g * HW;
    local += to_f(dy[((int64_t)(g * M + m) * (C + 1) + C) * HW + p]);
  }
  local = block_sum(local, red);
  if (threadIdx.x == 0) dstat[m] = local;
}

template <typename T>
__global__ void __launch_bounds__(256)
mbstd_bwd_kernel(const T *__restrict__ dy, const T *__restrict__ x, const float *__restrict__ dstat,
                 T *__restrict__ dx, int G, int M, int C, int64_t HW, float alpha) {
  const int m = blockIdx.y;
  const int64_t CHW = (int64_t)C * HW;
  const float coef = dstat[m] / ((float)CHW * (float)G);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < CHW;
       j += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    float mean = 0.f;
    for (int g = 0; g < G; ++g) {
      v[g] = to_f(x[((int64_t)g * M + m) * CHW + j]);
      mean += v[g];
    }
    mean /= (float)G;
    float var = 0.f;
    for (int g = 0; g < G; ++g) var = fmaf(v[g] - mean, v[g] - mean, var);
    const float inv_sd = rsqrtf(var / (float)G + alpha);
    for (int g = 0; g < G; ++g) {
      const int64_t b = (int64_t)g * M + m;
      const float gpass = dy ? to_f(dy[b * (C + 1) * HW + j]) : 0.f;
      dx[b * CHW + j] = from_f<T>(gpass + coef * (v[g] - mean) * inv_sd);
    }
  }
}

// ------------------------------------------------------------------ a13 sum of squares
template <typename T>
__global__ void __launch_bounds__(256)
sumsq_rows_kernel(const T *__restrict__ x, float *__restrict__ out, int64_t cols, int64_t chunk) {
  __shared__ float red[32];
  constexpr int V = Vec16<T>::N;
  const int64_t row = blockIdx.y;
  const T *xp = x + row * cols;
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  int64_t hi = lo + chunk;
  if (hi > cols) hi = cols;
  float acc = 0.f;
  const bool vec = (cols % V == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  if (vec) {
    for (int64_t j = lo + (int64_t)threadIdx.x * V; j < hi; j += (int64_t)blockDim.x * V) {
      Vec16<T> v = ld16(xp + j);
#pragma unroll
      for (int k = 0; k < V; ++k) acc = fmaf(v.get(k), v.get(k), acc);
    }
  } else {
    for (int64_t j = lo + threadIdx.x; j < hi; j += blockDim.x) {
      const float v = to_f(xp[j]);
      acc = fmaf(v, v, acc);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out + row, acc);
}

// ------------------------------------------------------------------ a7 circular shift
__global__ void __launch_bounds__(256)
circ_shift_kernel(const float *__restrict__ v, const float *__restrict__ shift01,
                  float *__restrict__ out, int C, int H, int W, float scale, int adjoint,
                  int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % W);
    const int64_t row = i / W;
    const int b = (int)(row / ((int64_t)C * H));
    const float tw = shift01[b] * (float)W;
    const float fl = floorf(tw);
    const float f = tw - fl;
    const int n = ((int)fl) % W;
    const float *r = v + row * W;
    float val;
    if (!adjoint) {
      int a = j + n; if (a >= W) a -= W;
      int c = a + 1; if (c >= W) c -= W;
      val = (1.f - f) * r[a] + f * r[c];
    } else {
      int a = j - n; if (a < 0) a += W;
      int c = a - 1; if (c < 0) c += W;
      val = (1.f - f) * r[a] + f * r[c];
    }
    out[i] = val * scale;
  }
}

// ema <- lerp(ema, (sum_a + rep_b * sum_b) / numel, weight): ModConv2d's running input power
__global__ void ema_lerp_kernel(float *ema, const float *sum_a, const float *sum_b, float rep_b,
                                float inv_numel, float weight) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float total = (sum_a ? *sum_a : 0.f) + (sum_b ? rep_b * *sum_b : 0.f);
    const float var = total * inv_numel;
    const float e = *ema;
    *ema = e + weight * (var - e);
  }
}

static unsigned flat_grid(int64_t total, int per_sm = 8) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace dusty

using namespace dusty;

extern "C" int dusty_gumbel_raydrop_fwd(const float *logit, const float *image, const float *u,
                                        float *mask, float *image_out, float *dsoft, int *count,
                                        int64_t n, float rconst, float temperature, void *stream) {
  DUSTY_CHECK_ARG(logit && image && u && mask && image_out && dsoft, "null pointer");
  DUSTY_CHECK_ARG(temperature > 0.f, "temperature must be positive");
  if (n <= 0) return DUSTY_OK;
  raydrop_fwd_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(
      logit, image, u, mask, image_out, dsoft, count, n, rconst, 1.f / temperature);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_gumbel_raydrop_bwd(const float *g_out, const float *g_mask, const float *image,
                                        const float *mask, const float *dsoft, float *g_logit,
                                        float *g_image, int64_t n, float rconst, void *stream) {
  DUSTY_CHECK_ARG(image && mask && dsoft && g_logit && g_image, "null pointer");
  if (n <= 0) return DUSTY_OK;
  raydrop_bwd_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(
      g_out, g_mask, image, mask, dsoft, g_logit, g_image, n, rconst);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_point_project(const float *x, const float *trig, float *out,
                                   long long *valid_count, int B, int64_t HW, float min_depth,
                                   float max_depth, float tol, int layout, void *stream) {
  DUSTY_CHECK_ARG(x && trig && out, "null pointer");
  DUSTY_CHECK_ARG(B >= 1 && HW >= 1, "bad shape");
  DUSTY_CHECK_ARG(max_depth > min_depth && min_depth > 0.f, "bad depth range");
  DUSTY_CHECK_ARG(layout == 0 || layout == 1, "layout must be 0 (map) or 1 (set)");
  const int64_t total = (int64_t)B * HW;
  // thresholds are python doubles rounded to fp32 when compared with an fp32 tensor
  const float lo_thr = (float)(1.0 / (double)max_depth), hi_thr = (float)(1.0 / (double)min_depth);
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == 0)
    project_kernel<0><<<flat_grid(total), 256, 0, st>>>(
        x, trig, out, (unsigned long long *)valid_count, HW, total, min_depth, lo_thr, hi_thr, tol);
  else
    project_kernel<1><<<flat_grid(total), 256, 0, st>>>(
        x, trig, out, (unsigned long long *)valid_count, HW, total, min_depth, lo_thr, hi_thr, tol);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_minibatch_std_fwd(const void *x, void *y, float *stat, int B, int C,
                                       int64_t HW, int group, float alpha, int dtype,
                                       void *stream) {
  DUSTY_CHECK_ARG(x && stat, "null pointer");
  DUSTY_CHECK_ARG(B >= 1 && C >= 1 && HW >= 1 && group >= 1, "bad shape");
  const int G = B < group ? B : group;
  DUSTY_CHECK_ARG(B % G == 0 && G <= 8, "batch must be divisible by the group (<= 8)");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  const int M = B / G;
  const int64_t CHW = (int64_t)C * HW;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(stat, 0, sizeof(float) * M, st) != cudaSuccess) return DUSTY_ECUDA;
  if (y == nullptr) {       // statistic only (any per-sample element order)
    int64_t bx1 = (CHW + 255) / 256;
    if (bx1 > 64) bx1 = 64;
    dim3 grid1((unsigned)bx1, (unsigned)M);
    if (dtype == DUSTY_F32)
      mbstd_stat_kernel<float><<<grid1, 256, 0, st>>>((const float *)x, stat, G, M, CHW, alpha);
    else
      mbstd_stat_kernel<__nv_bfloat16><<<grid1, 256, 0, st>>>((const __nv_bfloat16 *)x, stat, G, M,
                                                              CHW, alpha);
    DUSTY_LAUNCH_CHECK();
    return DUSTY_OK;
  }
  int64_t bx = (CHW + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, (unsigned)M);
  const int64_t total = (int64_t)B * (C + 1) * HW;
  if (dtype == DUSTY_F32) {
    mbstd_stat_kernel<float><<<grid, 256, 0, st>>>((const float *)x, stat, G, M, CHW, alpha);
    mbstd_write_kernel<float><<<flat_grid(total), 256, 0, st>>>((const float *)x, stat, (float *)y,
                                                                B, C, HW, M);
  } else {
    mbstd_stat_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, stat, G, M,
                                                           CHW, alpha);
    mbstd_write_kernel<__nv_bfloat16><<<flat_grid(total), 256, 0, st>>>(
        (const __nv_bfloat16 *)x, stat, (__nv_bfloat16 *)y, B, C, HW, M);
  }
  DUSTY_LAUNCH_CHECK();
  count_launch(1);
  return DUSTY_OK;
}

extern "C" int dusty_minibatch_std_bwd(const void *dy, const void *x, void *dx, float *dstat,
                                       int B, int C, int64_t HW, int group, float alpha, int dtype,
                                       void *stream) {
  DUSTY_CHECK_ARG(x && dx && dstat, "null pointer");
  DUSTY_CHECK_ARG(B >= 1 && C >= 1 && HW >= 1 && group >= 1, "bad shape");
  const int G = B < group ? B : group;
  DUSTY_CHECK_ARG(B % G == 0 && G <= 8, "batch must be divisible by the group (<= 8)");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  const int M = B / G;
  const int64_t CHW = (int64_t)C * HW;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t bx = (CHW + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, (unsigned)M);
  if (dy == nullptr) {      // statistic-only variant: dstat is the incoming gradient
    if (dtype == DUSTY_F32)
      mbstd_bwd_kernel<float><<<grid, 256, 0, st>>>(nullptr, (const float *)x, dstat, (float *)dx, G,
                                                    M, C, HW, alpha);
    else
      mbstd_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(nullptr, (const __nv_bfloat16 *)x, dstat,
                                                            (__nv_bfloat16 *)dx, G, M, C, HW, alpha);
    DUSTY_LAUNCH_CHECK();
    return DUSTY_OK;
  }
  if (dtype == DUSTY_F32) {
    mbstd_dstat_kernel<float><<<M, 256, 0, st>>>((const float *)dy, dstat, B, C, HW, M);
    mbstd_bwd_kernel<float><<<grid, 256, 0, st>>>((const float *)dy, (const float *)x, dstat,
                                                  (float *)dx, G, M, C, HW, alpha);
  } else {
    mbstd_dstat_kernel<__nv_bfloat16><<<M, 256, 0, st>>>((const __nv_bfloat16 *)dy, dstat, B, C, HW,
                                                         M);
    mbstd_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        (const __nv_bfloat16 *)dy, (const __nv_bfloat16 *)x, dstat, (__nv_bfloat16 *)dx, G, M, C, HW,
        alpha);
  }
  DUSTY_LAUNCH_CHECK();
  count_launch(1);
  return DUSTY_OK;
}

extern "C" int dusty_sumsq_rows(const void *x, float *out, int64_t rows, int64_t cols,
                                int accumulate, int dtype, void *stream) {
  DUSTY_CHECK_ARG(x && out, "null pointer");
  DUSTY_CHECK_ARG(rows >= 1 && rows <= 65535 && cols >= 1, "bad shape");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate && cudaMemsetAsync(out, 0, sizeof(float) * rows, st) != cudaSuccess)
    return DUSTY_ECUDA;
  // enough CTAs to fill the machine even for a single row
  int64_t want = ((int64_t)num_sms() * 4 + rows - 1) / rows;
  int64_t chunk = (cols + want - 1) / want;
  const int64_t min_chunk = 256 * 8 * 2;
  if (chunk < min_chunk) chunk = min_chunk;
  chunk = (chunk + 7) / 8 * 8;
  dim3 grid((unsigned)((cols + chunk - 1) / chunk), (unsigned)rows);
  if (dtype == DUSTY_F32)
    sumsq_rows_kernel<float><<<grid, 256, 0, st>>>((const float *)x, out, cols, chunk);
  else
    sumsq_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, out, cols,
                                                           chunk);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_circular_shift(const float *v, const float *shift01, float *out, int B, int C,
                                    int H, int W, float scale, int adjoint, void *stream) {
  DUSTY_CHECK_ARG(v && shift01 && out, "null pointer");
  DUSTY_CHECK_ARG(B >= 1 && C >= 1 && H >= 1 && W >= 1, "bad shape");
  const int64_t total = (int64_t)B * C * H * W;
  circ_shift_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(v, shift01, out, C, H, W,
                                                                       scale, adjoint, total);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_ema_lerp(float *ema_var, const float *sum_a, const float *sum_b, float rep_b,
                              float inv_numel, float weight, void *stream) {
  DUSTY_CHECK_ARG(ema_var && (sum_a || sum_b), "null pointer");
  ema_lerp_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ema_var, sum_a, sum_b, rep_b, inv_numel, weight);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

// ------------------------------------------------------------------ a11: EqualLR weight preparation
// Conv2d + EqualLR (gans/models/ops/common.py:158-184 of the reference) multiplies by
// 1/sqrt(fan_in); folded into the weight that is scale, cast and NCHW -> NHWC filter layout:
// three ATen kernels per convolution per pass (and four on the way back).  One kernel each way:
//   fwd: out[o][rs][c] (bf16 / fp32) = w[o][c][rs] (fp32 master) * scale
//   adj: gw[o][c][rs] (fp32)         = g[o][rs][c] or g[o][c][rs] (bf16 / fp32) * scale
namespace dusty {
template <typename TO>
__global__ void __launch_bounds__(256)
weight_prep_fwd_kernel(const float *__restrict__ w, TO *__restrict__ out, TO *__restrict__ out_tco,
                       int64_t n, int O, int C, int RS, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t t = i / C;
    const int rs = (int)(t % RS);
    const int64_t o = t / RS;
    out[i] = from_f<TO>(__ldg(w + (o * C + c) * RS + rs) * scale);
  }
  if (out_tco == nullptr) return;
  // second layout for the data-gradient kernels: [tap][c][o], o contiguous
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    const int64_t t = i / O;
    const int c = (int)(t % C);
    const int64_t rs = t / C;
    out_tco[i] = from_f<TO>(__ldg(w + ((int64_t)o * C + c) * RS + rs) * scale);
  }
}
template <typename TI>
__global__ void __launch_bounds__(256)
weight_prep_adj_kernel(const TI *__restrict__ g, float *__restrict__ gw, int64_t n, int C, int RS,
                       float scale, int nhwc) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int rs = (int)(i % RS);
    const int64_t t = i / RS;
    const int c = (int)(t % C);
    const int64_t o = t / C;
    const int64_t src = nhwc ? (o * RS + rs) * C + c : i;
    gw[i] = to_f(g[src]) * scale;
  }
}
}  // namespace dusty

extern "C" int dusty_weight_prep(const float *w, void *out, void *out_tco, int O, int C, int RS,
                                 float scale, int out_dtype, void *stream) {
  DUSTY_CHECK_ARG(w && out, "null pointer");
  DUSTY_CHECK_ARG(O >= 1 && C >= 1 && RS >= 1, "bad shape");
  DUSTY_CHECK_ARG(out_dtype == DUSTY_F32 || out_dtype == DUSTY_BF16, "bad dtype");
  const int64_t n = (int64_t)O * C * RS;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == DUSTY_F32)
    weight_prep_fwd_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(w, (float *)out, (float *)out_tco, n, O, C, RS,
                                                                    scale);
  else
    weight_prep_fwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
        w, (__nv_bfloat16 *)out, (__nv_bfloat16 *)out_tco, n, O, C, RS, scale);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_weight_prep_adj(const void *g, float *gw, int O, int C, int RS, float scale,
                                     int g_dtype, int g_nhwc, void *stream) {
  DUSTY_CHECK_ARG(g && gw, "null pointer");
  DUSTY_CHECK_ARG(O >= 1 && C >= 1 && RS >= 1, "bad shape");
  DUSTY_CHECK_ARG(g_dtype == DUSTY_F32 || g_dtype == DUSTY_BF16, "bad dtype");
  const int64_t n = (int64_t)O * C * RS;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_dtype == DUSTY_F32)
    weight_prep_adj_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float *)g, gw, n, C, RS, scale, g_nhwc);
  else
    weight_prep_adj_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16 *)g, gw, n, C, RS,
                                                                           scale, g_nhwc);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

// filter gradient as the wgrad kernel produces it, fp32 [RS][C][O], -> OHWI [O][RS][C] in the
// activation dtype (the layout / dtype of the prepared filter whose gradient it is): one
// kernel instead of a permuted cast + a layout copy
namespace dusty {
template <typename TO>
__global__ void __launch_bounds__(256)
filter_rsco_to_ohwi_kernel(const float *__restrict__ src, TO *__restrict__ dst, int64_t n, int O, int C,
                           int RS) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t t = i / C;
    const int rs = (int)(t % RS);
    const int64_t o = t / RS;
    dst[i] = from_f<TO>(__ldg(src + ((int64_t)rs * C + c) * O + o));
  }
}
}  // namespace dusty

extern "C" int dusty_filter_rsco_to_ohwi(const float *src, void *dst, int O, int C, int RS, int dst_dtype,
                                         void *stream) {
  DUSTY_CHECK_ARG(src && dst, "null pointer");
  DUSTY_CHECK_ARG(O >= 1 && C >= 1 && RS >= 1, "bad shape");
  DUSTY_CHECK_ARG(dst_dtype == DUSTY_F32 || dst_dtype == DUSTY_BF16, "bad dtype");
  const int64_t n = (int64_t)O * C * RS;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dst_dtype == DUSTY_F32)
    filter_rsco_to_ohwi_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(src, (float *)dst, n, O, C, RS);
  else
    filter_rsco_to_ohwi_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(src, (__nv_bfloat16 *)dst, n, O, C, RS);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

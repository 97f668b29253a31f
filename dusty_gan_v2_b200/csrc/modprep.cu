// a1: per-sample effective weights of ModConv2d (style.py:72-103), fused.
//
//   w'  = scale*W / max|scale*W|        (demod)      else scale*W
//   s'  = s / max_i|s_i| + 1            (demod)      else s + 1          s = mod(style)
//   t   = w'[o,i] * s'[b,i]
//   d   = rsqrt(sum_i t^2 + 1e-8)       (demod)      else 1
//   wb  = t * d * g,    g = 1 / (sqrt(ema_var) + 1e-8)
//
// The reference spends ~15 ATen launches on this per layer and autograd ~30 more on the way
// back; here it is 2 launches forward and 5 backward with the analytic gradient (including
// the inf-norm pre-normalisations, whose gradient lands on the arg-max element).  All math
// fp32; wb is written in the activation dtype of the contraction kernel.
#include "common.cuh"

namespace dusty {

// stats layout (fp32): [0,B) smax | [B] wmax | [B+1] g | [B+2, B+2+B*O) d
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256)
modprep_stats_kernel(const float *__restrict__ slin, const float *__restrict__ W,
                     const float *__restrict__ ema_var, float *__restrict__ stats, int B, int O,
                     int I, float scale, int demod) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float m = 0.f;
  if ((int)blockIdx.x < B) {
    if (demod) for (int i = threadIdx.x; i < I; i += blockDim.x) m = fmaxf(m, fabsf(slin[(int64_t)blockIdx.x * I + i]));
  } else {
    const int64_t n = (int64_t)O * I;
    const int nb = gridDim.x - B;
    if (demod)
      for (int64_t i = (int64_t)(blockIdx.x - B) * blockDim.x + threadIdx.x; i < n; i += (int64_t)nb * blockDim.x)
        m = fmaxf(m, fabsf(W[i] * scale));
  }
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  if (wid == 0) {
    m = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    m = warp_max(m);
    if (lane == 0) {
      if ((int)blockIdx.x < B) stats[blockIdx.x] = demod ? m : 1.f;
      else if (demod) atomicMax(reinterpret_cast<int *>(stats + B), __float_as_int(m));  // m >= 0
      if (blockIdx.x == 0) stats[B + 1] = ema_var ? 1.f / (sqrtf(*ema_var) + 1e-8f) : 1.f;
    }
  }
}

struct PrepCtx {
  const float *slin, *W, *stats;
  int B, O, I, demod;
  float scale;
  // optional per-sample rotation of the Fourier columns (azimuth-shift identity, see header):
  // rot[b, f] = cos(psi_bf), rot[b, F + f] = sin(psi_bf); columns [C1, C1+F) are the sin block,
  // [C1+F, C1+2F) the cos block.  rot == nullptr: none.
  const float *rot;
  int C1, F;
  // backward only: ema_var read at backward time (the forward prepared the weights WITHOUT the
  // EMA normaliser, which the contraction's epilogue applied; d(loss)/d(wb * g) arrives here)
  const float *g_late;
  __device__ __forceinline__ bool rotated(int i) const { return rot != nullptr && i >= C1; }
  __device__ __forceinline__ float smax(int b) const { return stats[b]; }
  __device__ __forceinline__ float wmax() const { return demod ? stats[B] : 1.f; }
  __device__ __forceinline__ float g() const {
    return g_late ? 1.f / (sqrtf(__ldg(g_late)) + 1e-8f) : stats[B + 1];
  }
  __device__ __forceinline__ float d(int b, int o) const { return stats[B + 2 + (int64_t)b * O + o]; }
  __device__ __forceinline__ float wp(int o, int i, float inv_wmax) const {
    return W[(int64_t)o * I + i] * scale * inv_wmax;
  }
  __device__ __forceinline__ float sp(int b, int i, float inv_smax) const {
    return slin[(int64_t)b * I + i] * inv_smax + 1.f;
  }
};

// one warp per (b, o) row
template <typename T>
__global__ void __launch_bounds__(128)
modprep_wb_kernel(PrepCtx c, float *__restrict__ stats_w, T *__restrict__ wb) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (o >= c.O) return;
  const float inv_w = 1.f / c.wmax(), inv_s = 1.f / c.smax(b);
  float ss = 0.f;
  for (int i = lane; i < c.I; i += 32) {
    const float t = c.wp(o, i, inv_w) * c.sp(b, i, inv_s);
    ss = fmaf(t, t, ss);
  }
  ss = warp_sum(ss);
  const float d = c.demod ? rsqrtf(ss + 1e-8f) : 1.f;
  if (lane == 0) stats_w[c.B + 2 + (int64_t)b * c.O + o] = d;
  const float dg = d * c.g();
  T *row = wb + ((int64_t)b * c.O + o) * c.I;
  if (c.rot == nullptr) {
    for (int i = lane; i < c.I; i += 32) row[i] = from_f<T>(c.wp(o, i, inv_w) * c.sp(b, i, inv_s) * dg);
  } else {
    // rotation preserves sum t^2, so d above is unaffected
    for (int i = lane; i < c.C1 + c.F; i += 32) {
      const float ts = c.wp(o, i, inv_w) * c.sp(b, i, inv_s) * dg;
      if (i < c.C1) {
        row[i] = from_f<T>(ts);
      } else {
        const int f = i - c.C1;
        const float tc = c.wp(o, i + c.F, inv_w) * c.sp(b, i + c.F, inv_s) * dg;
        const float cs = c.rot[(int64_t)b * 2 * c.F + f], sn = c.rot[(int64_t)b * 2 * c.F + c.F + f];
        row[i] = from_f<T>(ts * cs - tc * sn);
        row[i + c.F] = from_f<T>(ts * sn + tc * cs);
      }
    }
  }
}

// incoming gradient w.r.t. the UN-rotated weights (transpose of the rotation above)
__device__ __forceinline__ float prep_gw(const PrepCtx &c, const float *__restrict__ gwb, int b,
                                         int o, int i) {
  const float *row = gwb + ((int64_t)b * c.O + o) * c.I;
  if (!c.rotated(i)) return row[i];
  const int r = i - c.C1;
  const bool is_cos = r >= c.F;
  const int f = is_cos ? r - c.F : r;
  const float gs = row[c.C1 + f], gc = row[c.C1 + c.F + f];
  const float cs = c.rot[(int64_t)b * 2 * c.F + f], sn = c.rot[(int64_t)b * 2 * c.F + c.F + f];
  return is_cos ? (gc * cs - gs * sn) : (gs * cs + gc * sn);
}

// c[b,o] = sum_i gwb * g * t      (demod only); one warp per (b, o) row
__global__ void __launch_bounds__(128)
modprep_c_kernel(PrepCtx c, const float *__restrict__ gwb, float *__restrict__ cbo) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (o >= c.O) return;
  const float inv_w = 1.f / c.wmax(), inv_s = 1.f / c.smax(b);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int i = lane;
  for (; i + 96 < c.I; i += 128) {      // four independent load groups in flight
    acc0 = fmaf(prep_gw(c, gwb, b, o, i), c.wp(o, i, inv_w) * c.sp(b, i, inv_s), acc0);
    acc1 = fmaf(prep_gw(c, gwb, b, o, i + 32), c.wp(o, i + 32, inv_w) * c.sp(b, i + 32, inv_s), acc1);
    acc2 = fmaf(prep_gw(c, gwb, b, o, i + 64), c.wp(o, i + 64, inv_w) * c.sp(b, i + 64, inv_s), acc2);
    acc3 = fmaf(prep_gw(c, gwb, b, o, i + 96), c.wp(o, i + 96, inv_w) * c.sp(b, i + 96, inv_s), acc3);
  }
  for (; i < c.I; i += 32)
    acc0 = fmaf(prep_gw(c, gwb, b, o, i), c.wp(o, i, inv_w) * c.sp(b, i, inv_s), acc0);
  const float acc = warp_sum((acc0 + acc1) + (acc2 + acc3));
  if (lane == 0) cbo[(int64_t)b * c.O + o] = acc * c.g();
}

// dt[b,o,i] = u*d - c*d^3*t  (demod)   |   u  (otherwise),   u = gwb * g
__device__ __forceinline__ float prep_dt(const PrepCtx &c, float gw, float t, float d, float cbo) {
  const float u = gw * c.g();
  return c.demod ? (u * d - cbo * d * d * d * t) : u;
}

constexpr int kRedY = 8;     // threads along the reduced axis per block (block = 32 x kRedY)

// ds'[b,i] = sum_o dt * w'.  Block = 32 columns i x 8 slices of the o axis; the serial
// version (one thread per (b, i), O dependent iterations) was pure load latency.
__global__ void __launch_bounds__(32 * kRedY)
modprep_ds_kernel(PrepCtx c, const float *__restrict__ gwb, const float *__restrict__ cbo,
                  float *__restrict__ dsp, float *__restrict__ sums) {
  __shared__ float part[kRedY][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * 32 + tx;
  const int b = blockIdx.y;
  float acc = 0.f;
  if (i < c.I) {
    const float inv_w = 1.f / c.wmax(), inv_s = 1.f / c.smax(b);
    const float sp = c.sp(b, i, inv_s);
#pragma unroll 4
    for (int o = ty; o < c.O; o += kRedY) {
      const float wp = c.wp(o, i, inv_w);
      const float dt = prep_dt(c, prep_gw(c, gwb, b, o, i), wp * sp, c.d(b, o),
                               c.demod ? cbo[(int64_t)b * c.O + o] : 0.f);
      acc = fmaf(dt, wp, acc);
    }
  }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int k = 1; k < kRedY; ++k) acc += part[k][tx];
    float dot = 0.f;
    if (i < c.I) {
      dsp[(int64_t)b * c.I + i] = acc;
      dot = acc * c.slin[(int64_t)b * c.I + i];
    }
    if (c.demod) {       // sum_j ds'_j s_j, needed by the inf-norm term of the finish pass
      dot = warp_sum(dot);
      if (tx == 0) atomicAdd(sums + b, dot);
    }
  }
}

// dw'[o,i] = sum_b dt * s'.  Block = 32 columns i x 8 slices of the batch axis.
__global__ void __launch_bounds__(32 * kRedY)
modprep_dw_kernel(PrepCtx c, const float *__restrict__ gwb, const float *__restrict__ cbo,
                  float *__restrict__ dwp, float *__restrict__ sums) {
  __shared__ float part[kRedY][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * 32 + tx;
  const int o = blockIdx.y;
  float acc = 0.f;
  if (i < c.I) {
    const float inv_w = 1.f / c.wmax();
    const float wp = c.wp(o, i, inv_w);
#pragma unroll 4
    for (int b = ty; b < c.B; b += kRedY) {
      const float sp = c.sp(b, i, 1.f / c.smax(b));
      const float dt = prep_dt(c, prep_gw(c, gwb, b, o, i), wp * sp, c.d(b, o),
                               c.demod ? cbo[(int64_t)b * c.O + o] : 0.f);
      acc = fmaf(dt, sp, acc);
    }
  }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int k = 1; k < kRedY; ++k) acc += part[k][tx];
    float dot = 0.f;
    if (i < c.I) {
      dwp[(int64_t)o * c.I + i] = acc;
      dot = acc * c.W[(int64_t)o * c.I + i] * c.scale;
    }
    if (c.demod) {
      dot = warp_sum(dot);
      if (tx == 0) atomicAdd(sums + c.B, dot);
    }
  }
}

// Elementwise finish over the B*I style entries and the O*I weight entries.
// v = x / max|x| (+1):  dx_i = dv_i / m  -  [|x_i| == m] * sign(x_i) * (sum_j dv_j x_j) / m^2
// (the arg-max element is recognised by equality with the stored maximum).
__global__ void __launch_bounds__(256)
modprep_finish_kernel(PrepCtx c, const float *__restrict__ dsp, const float *__restrict__ dwp,
                      const float *__restrict__ sums, float *__restrict__ dslin,
                      float *__restrict__ dW) {
  const int64_t ns = (int64_t)c.B * c.I, nw = (int64_t)c.O * c.I;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ns + nw;
       t += (int64_t)gridDim.x * blockDim.x) {
    const bool is_w = t >= ns;
    const int64_t i = is_w ? t - ns : t;
    const float pre = is_w ? c.scale : 1.f;
    const float dv = is_w ? dwp[i] : dsp[i];
    float *dst = is_w ? dW + i : dslin + i;
    if (!c.demod) {
      *dst = dv * pre;
      continue;
    }
    const int b = is_w ? 0 : (int)(i / c.I);
    const float m = is_w ? c.wmax() : c.smax(b);
    const float sum = is_w ? sums[c.B] : sums[b];
    const float xv = (is_w ? c.W[i] : c.slin[i]) * pre;
    const float inv_m = 1.f / m;
    float g = dv * inv_m;
    if (fabsf(xv) == m) g -= (xv >= 0.f ? 1.f : -1.f) * sum * inv_m * inv_m;
    *dst = g * pre;
  }
}

}  // namespace dusty

using namespace dusty;

extern "C" int dusty_modprep_fwd(const float *slin, const float *weight, const float *ema_var,
                                 void *wb, float *stats, int B, int O, int I, float scale,
                                 int demod, int wdtype, const float *rot, int C1, int F,
                                 void *stream) {
  DUSTY_CHECK_ARG(slin && weight && wb && stats, "null pointer");
  DUSTY_CHECK_ARG(rot == nullptr || (C1 >= 0 && F >= 1 && C1 + 2 * F == I), "bad rotation layout");
  DUSTY_CHECK_ARG(B >= 1 && B <= 65535 && O >= 1 && I >= 1, "bad shape");
  DUSTY_CHECK_ARG(wdtype == DUSTY_F32 || wdtype == DUSTY_BF16, "bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(stats + B, 0, sizeof(float), st) != cudaSuccess) return DUSTY_ECUDA;
  int nw = (int)(((int64_t)O * I + 256 * 16 - 1) / (256 * 16));
  if (nw > 64) nw = 64;
  if (nw < 1) nw = 1;
  modprep_stats_kernel<<<B + nw, 256, 0, st>>>(slin, weight, ema_var, stats, B, O, I, scale, demod);
  PrepCtx c{slin, weight, stats, B, O, I, demod, scale, rot, C1, F, nullptr};
  dim3 grid((unsigned)((O + 3) / 4), (unsigned)B);
  if (wdtype == DUSTY_F32) modprep_wb_kernel<float><<<grid, 128, 0, st>>>(c, stats, (float *)wb);
  else modprep_wb_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(c, stats, (__nv_bfloat16 *)wb);
  DUSTY_LAUNCH_CHECK();
  count_launch(1);
  return DUSTY_OK;
}

extern "C" int dusty_modprep_bwd(const float *gwb, const float *slin, const float *weight,
                                 const float *stats, float *dslin, float *dweight, float *work,
                                 int B, int O, int I, float scale, int demod, const float *rot,
                                 int C1, int F, const float *ema_late, void *stream) {
  DUSTY_CHECK_ARG(gwb && slin && weight && stats && dslin && dweight && work, "null pointer");
  DUSTY_CHECK_ARG(rot == nullptr || (C1 >= 0 && F >= 1 && C1 + 2 * F == I), "bad rotation layout");
  DUSTY_CHECK_ARG(B >= 1 && B <= 65535 && O >= 1 && O <= 65535 && I >= 1, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  PrepCtx c{slin, weight, stats, B, O, I, demod, scale, rot, C1, F, ema_late};
  float *cbo = work;                         // [B*O]
  float *dsp = work + (int64_t)B * O;        // [B*I]
  float *dwp = dsp + (int64_t)B * I;         // [O*I]
  float *sums = dwp + (int64_t)O * I;        // [B + 1]
  if (cudaMemsetAsync(sums, 0, sizeof(float) * (B + 1), st) != cudaSuccess) return DUSTY_ECUDA;
  if (demod) {
    dim3 grid((unsigned)((O + 3) / 4), (unsigned)B);
    modprep_c_kernel<<<grid, 128, 0, st>>>(c, gwb, cbo);
    count_launch(1);
  }
  const dim3 rblock(32, kRedY);
  modprep_ds_kernel<<<dim3((unsigned)((I + 31) / 32), (unsigned)B), rblock, 0, st>>>(c, gwb, cbo, dsp, sums);
  modprep_dw_kernel<<<dim3((unsigned)((I + 31) / 32), (unsigned)O), rblock, 0, st>>>(c, gwb, cbo, dwp, sums);
  {
    int64_t nb = (((int64_t)B + O) * I + 255) / 256;
    if (nb > 1184) nb = 1184;
    modprep_finish_kernel<<<(unsigned)nb, 256, 0, st>>>(c, dsp, dwp, sums, dslin, dweight);
  }
  DUSTY_LAUNCH_CHECK();
  count_launch(2);
  return DUSTY_OK;
}

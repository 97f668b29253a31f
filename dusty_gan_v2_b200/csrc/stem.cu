// a11 (first three layers of the discriminator, gans/models/dusty_v2.py:352-354 of the
// reference): BlurVH -> 1x1 convolution 2 -> O (EqualLR, no bias) -> FusedLeakyReLU(O).
//
// As separate ops this is five passes over the [B, O, 64, 512] tensor (conv output, bias_act
// read + write, NCHW -> NHWC conversion read + write) in front of a two-channel contraction
// that no tensor-core kernel wants (K = 2).  Here it is ONE pass: each thread evaluates the two
// blurred values of its pixel from the 1-channel input (6 cached loads), forms 8 output
// channels (w[o,0]*v + w[o,1]*h + b[o], leaky ReLU, gain) and writes them as one 16-byte NHWC
// vector.  Algorithmic bytes: 4 B read + 2*O B written per pixel.
//
// Backward, also one pass over (dy, y): gate from the saved output, then per thread
//   d_pre[o] = dy[o] * (y[o] > 0 ? 1 : alpha) * gain
//   db[o] += d_pre[o];  dW[o,0] += d_pre[o] * v;  dW[o,1] += d_pre[o] * h     (block-reduced, atomics)
//   dv = sum_o w[o,0] d_pre[o],  dh = sum_o w[o,1] d_pre[o]                 (shuffle over the O/8 lanes)
// and a small second kernel applies the adjoint of the two 3-tap blurs to (dv, dh).
#include "common.cuh"

namespace dusty {

struct StemTaps { float k[3]; };

// blurred pair of pixel (y, x): v = vertical [k0 k1 k2] with clamped rows, h = horizontal with
// circular columns (common.py:141-155: channel 0 = blur over H, channel 1 = blur over W)
template <typename TX>
__device__ __forceinline__ void stem_vh(const TX *__restrict__ img, int y, int x, int H, int W,
                                        const StemTaps &t, float &v, float &h) {
  const int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
  const int xm = x > 0 ? x - 1 : W - 1, xp = x < W - 1 ? x + 1 : 0;
  const float c = to_f(img[(int64_t)y * W + x]);
  v = t.k[0] * to_f(img[(int64_t)ym * W + x]) + t.k[1] * c + t.k[2] * to_f(img[(int64_t)yp * W + x]);
  h = t.k[0] * to_f(img[(int64_t)y * W + xm]) + t.k[1] * c + t.k[2] * to_f(img[(int64_t)y * W + xp]);
}

// thread -> (pixel, group of 8 output channels); OG = O / 8 lanes share a pixel
template <typename TX>
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const TX *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                __nv_bfloat16 *__restrict__ y, int64_t n_pix, int H, int W, int OG, StemTaps t,
                float alpha, float scale) {
  const int og = threadIdx.x % OG;
  float w0[8], w1[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    w0[j] = __ldg(w + (og * 8 + j) * 2);
    w1[j] = __ldg(w + (og * 8 + j) * 2 + 1);
    b[j] = bias != nullptr ? __ldg(bias + og * 8 + j) : 0.f;
  }
  // Each block owns a contiguous range of image rows (b, y) and walks it row by row, 256 / OG
  // pixels at a time: no integer divisions per pixel (the first version spent ~40 of its ~120
  // instructions per vector on p / (H*W) and r / W and was issue-bound: ncu 63 % issue slots
  // busy at 18 % of DRAM bandwidth).
  const int ppb = blockDim.x / OG;
  const int rows_total = (int)(n_pix / W);
  const int rpb = (rows_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int row0 = (int)blockIdx.x * rpb;
  const int row1 = min(row0 + rpb, rows_total);
  int bi = row0 / H, yy = row0 - bi * H;
  const int xl = threadIdx.x / OG;
  for (int row = row0; row < row1; ++row) {
    const TX *img = x + (int64_t)bi * H * W;
    __nv_bfloat16 *yrow = y + ((int64_t)row * W) * OG * 8;
    for (int xx = xl; xx < W; xx += ppb) {
      float v, h;
      stem_vh(img, yy, xx, H, W, t, v, h);
      Vec16<__nv_bfloat16> o;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a0 = fmaf(w0[2 * j], v, fmaf(w1[2 * j], h, b[2 * j]));
        float a1 = fmaf(w0[2 * j + 1], v, fmaf(w1[2 * j + 1], h, b[2 * j + 1]));
        a0 = (a0 > 0.f ? a0 : a0 * alpha) * scale;
        a1 = (a1 > 0.f ? a1 : a1 * alpha) * scale;
        set2(o, j, make_float2(a0, a1));
      }
      st16_stream(yrow + ((int64_t)xx * OG + og) * 8, o);
    }
    if (++yy == H) { yy = 0; ++bi; }
  }
}

template <typename TX>
__global__ void __launch_bounds__(256, 3)
stem_bwd_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ y,
                const TX *__restrict__ x, const float *__restrict__ w, float *__restrict__ dvh,
                float *__restrict__ dwb, int64_t n_pix, int H, int W, int OG, StemTaps t, float alpha,
                float scale) {
  __shared__ float red[8][24 * 8];                      // [warp][og (<= 8)][24]
  const int og = threadIdx.x % OG;
  float w0[8], w1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    w0[j] = __ldg(w + (og * 8 + j) * 2);
    w1[j] = __ldg(w + (og * 8 + j) * 2 + 1);
  }
  float sb[8], s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sb[j] = s0[j] = s1[j] = 0.f;
  // same row-range walk as the forward kernel (no per-pixel divisions); every lane of a warp
  // runs the same trip counts (the shuffles below need full warps)
  const int ppb = blockDim.x / OG;
  const int rows_total = (int)(n_pix / W);
  const int rpb = (rows_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int row0 = (int)blockIdx.x * rpb;
  const int row1 = min(row0 + rpb, rows_total);
  int bi = row0 / H, yy = row0 - bi * H;
  const int xl = threadIdx.x / OG;
  const int chunks = (W + ppb - 1) / ppb;
  for (int row = row0; row < row1; ++row) {
    const TX *img = x + (int64_t)bi * H * W;
    const int64_t rbase = (int64_t)row * W;
    for (int ch = 0; ch < chunks; ++ch) {
      const int xx = xl + ch * ppb;
      const bool live = xx < W;
      float dv = 0.f, dh = 0.f;
      if (live) {
        const Vec16<__nv_bfloat16> g = ld16_stream(dy + ((rbase + xx) * OG + og) * 8);
        const Vec16<__nv_bfloat16> o = ld16_stream(y + ((rbase + xx) * OG + og) * 8);
        float v, h;
        stem_vh(img, yy, xx, H, W, t, v, h);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gp = g.get(j) * (o.get(j) > 0.f ? 1.f : alpha) * scale;
          sb[j] += gp;
          s0[j] = fmaf(gp, v, s0[j]);
          s1[j] = fmaf(gp, h, s1[j]);
          dv = fmaf(w0[j], gp, dv);
          dh = fmaf(w1[j], gp, dh);
        }
      }
      if (dvh != nullptr) {                             // sum over the OG lanes of this pixel
        for (int m = 1; m < OG; m <<= 1) {
          dv += __shfl_xor_sync(0xffffffffu, dv, m);
          dh += __shfl_xor_sync(0xffffffffu, dh, m);
        }
        if (live && og == 0) {
          const int64_t hw = (int64_t)H * W;
          const int64_t r = (int64_t)yy * W + xx;
          dvh[((int64_t)bi * 2) * hw + r] = dv;
          dvh[((int64_t)bi * 2 + 1) * hw + r] = dh;
        }
      }
    }
    if (++yy == H) { yy = 0; ++bi; }
  }
  // reduce the 24 sums over lanes with equal og, then over warps, then one atomic per value
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    for (int m = OG; m < 32; m <<= 1) {
      sb[j] += __shfl_xor_sync(0xffffffffu, sb[j], m);
      s0[j] += __shfl_xor_sync(0xffffffffu, s0[j], m);
      s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], m);
    }
  }
  if (lane < OG) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[wid][lane * 24 + j] = s0[j];
      red[wid][lane * 24 + 8 + j] = s1[j];
      red[wid][lane * 24 + 16 + j] = sb[j];
    }
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < OG * 24; i += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < nw; ++q) s += red[q][i];
    const int g8 = i / 24, k = i % 24, j = k & 7, which = k >> 3;   // which: 0 dW0, 1 dW1, 2 db
    atomicAdd(dwb + (g8 * 8 + j) * 3 + which, s);
  }
}

// dx = blur_v^T(dv) + blur_h^T(dh) on the [B, 2, H, W] fp32 gradient of the blurred pair
__global__ void __launch_bounds__(256)
stem_dx_kernel(const float *__restrict__ dvh, float *__restrict__ dx, int64_t n_pix, int H, int W,
               StemTaps t) {
  const int64_t hw = (int64_t)H * W;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bi = p / hw;
    const int r = (int)(p - bi * hw);
    const int yy = r / W, xx = r - yy * W;
    const float *dv = dvh + (bi * 2) * hw, *dh = dv + hw;
    // vertical: out[n] = k0 x[c(n-1)] + k1 x[n] + k2 x[c(n+1)]  =>  x[m] collects every (n, tap)
    // whose clamped source row is m
    float a = t.k[1] * dv[(int64_t)yy * W + xx];
    if (yy + 1 <= H - 1) a += t.k[0] * dv[(int64_t)(yy + 1) * W + xx];   // n = m+1, tap n-1
    if (yy - 1 >= 0) a += t.k[2] * dv[(int64_t)(yy - 1) * W + xx];       // n = m-1, tap n+1
    if (yy == 0) a += t.k[0] * dv[xx];                                   // n = 0, tap -1 clamps to 0
    if (yy == H - 1) a += t.k[2] * dv[(int64_t)(H - 1) * W + xx];        // n = H-1, tap H clamps
    const int xm = xx > 0 ? xx - 1 : W - 1, xp = xx < W - 1 ? xx + 1 : 0;
    // horizontal (circular): x[m] <- k0 dh[m+1] + k1 dh[m] + k2 dh[m-1]
    a += t.k[0] * dh[(int64_t)yy * W + xp] + t.k[1] * dh[(int64_t)yy * W + xx] +
         t.k[2] * dh[(int64_t)yy * W + xm];
    dx[p] = a;
  }
}

static unsigned stem_grid(int64_t n_pix, int ppb) {
  int64_t blocks = (n_pix + ppb - 1) / ppb;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace dusty

using namespace dusty;

static bool stem_shape_ok(int B, int H, int W, int O) {
  const int og = O / 8;
  if ((int64_t)B * H * W * (og > 0 ? og : 1) >= 0x7fffffffLL) return false;   // 32-bit pixel indices
  return B >= 1 && H >= 2 && W >= 2 && O >= 8 && O % 8 == 0 && og <= 8 && (og & (og - 1)) == 0;
}

extern "C" int dusty_stem_fwd(const void *x, const float *w, const float *bias, void *y, int B, int H,
                              int W, int O, float k0, float k1, float k2, float alpha, float scale,
                              int x_dtype, void *stream) {
  DUSTY_CHECK_ARG(x && w && y, "null pointer");
  DUSTY_CHECK_ARG(stem_shape_ok(B, H, W, O), "O must be 8, 16, 32 or 64; H, W >= 2");
  DUSTY_CHECK_ARG(x_dtype == DUSTY_F32 || x_dtype == DUSTY_BF16, "bad dtype");
  DUSTY_CHECK_ARG(aligned16(y), "y must be 16-byte aligned");
  const int OG = O / 8;
  const int64_t n_pix = (int64_t)B * H * W;
  StemTaps t{{k0, k1, k2}};
  cudaStream_t st = (cudaStream_t)stream;
  int64_t gblocks = (int64_t)B * H;                  // one or more image rows per block
  if (gblocks > (int64_t)num_sms() * 16) gblocks = (int64_t)num_sms() * 16;
  const unsigned grid = (unsigned)gblocks;
  if (x_dtype == DUSTY_F32)
    stem_fwd_kernel<float><<<grid, 256, 0, st>>>((const float *)x, w, bias, (__nv_bfloat16 *)y, n_pix, H, W,
                                                 OG, t, alpha, scale);
  else
    stem_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, w, bias,
                                                         (__nv_bfloat16 *)y, n_pix, H, W, OG, t, alpha, scale);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_stem_bwd(const void *dy, const void *y, const void *x, const float *w, float *dvh,
                              float *dwb, int B, int H, int W, int O, float k0, float k1, float k2,
                              float alpha, float scale, int x_dtype, void *stream) {
  DUSTY_CHECK_ARG(dy && y && x && w && dwb, "null pointer");
  DUSTY_CHECK_ARG(stem_shape_ok(B, H, W, O), "O must be 8, 16, 32 or 64; H, W >= 2");
  DUSTY_CHECK_ARG(x_dtype == DUSTY_F32 || x_dtype == DUSTY_BF16, "bad dtype");
  DUSTY_CHECK_ARG(aligned16(dy) && aligned16(y), "dy / y must be 16-byte aligned");
  const int OG = O / 8;
  const int64_t n_pix = (int64_t)B * H * W;
  StemTaps t{{k0, k1, k2}};
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(dwb, 0, sizeof(float) * O * 3, st) != cudaSuccess) {
    set_error("dusty_stem_bwd: memset failed");
    return DUSTY_ECUDA;
  }
  int64_t blocks = (int64_t)num_sms() * 6;           // row ranges; few blocks keep the atomics cheap
  if (blocks > (int64_t)B * H) blocks = (int64_t)B * H;
  if (x_dtype == DUSTY_F32)
    stem_bwd_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(
        (const __nv_bfloat16 *)dy, (const __nv_bfloat16 *)y, (const float *)x, w, dvh, dwb, n_pix, H, W, OG, t,
        alpha, scale);
  else
    stem_bwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
        (const __nv_bfloat16 *)dy, (const __nv_bfloat16 *)y, (const __nv_bfloat16 *)x, w, dvh, dwb, n_pix, H, W,
        OG, t, alpha, scale);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_stem_dx(const float *dvh, float *dx, int B, int H, int W, float k0, float k1,
                             float k2, void *stream) {
  DUSTY_CHECK_ARG(dvh && dx, "null pointer");
  DUSTY_CHECK_ARG(B >= 1 && H >= 2 && W >= 2, "bad shape");
  const int64_t n_pix = (int64_t)B * H * W;
  StemTaps t{{k0, k1, k2}};
  stem_dx_kernel<<<stem_grid(n_pix, 256), 256, 0, (cudaStream_t)stream>>>(dvh, dx, n_pix, H, W, t);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

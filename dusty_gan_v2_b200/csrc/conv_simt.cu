// a11 / a15: dense 2-D convolution family on the CUDA cores (fp32 accumulation): forward, data
// gradient (= transposed convolution) and filter gradient for ANY layout (element strides),
// filter size, stride and zero padding, fp32 or bf16 tensors.  This is the path of fp32 parity
// mode (exact fp32 FMAs: no TF32 rounding between the mirror and the CPU oracle) and of every
// shape outside the tcgen05 kernels' domain (conv_tc.cu: bf16 NHWC, channel counts that are
// multiples of 8) -- e.g. the 1- and 2-channel first layers, the 513-channel epilogue
// convolution, the 4x4 (transposed) convolutions of the vanilla / dusty_v1 baselines in fp32.
// The product never calls a library convolution.
//
// One implicit-GEMM kernel, C[m, n] = sum_k A(m, k) * B(n, k), 64 x 64 output tile per CTA,
// 16-deep k slices staged in shared memory, 4 x 4 outputs per thread:
//   fprop  m = (b, oh, ow)  n = o          k = (r, s, c)     A = x window,  B = w
//   dgrad  m = (b, ih, iw)  n = c          k = (r, s, o)     A = dy taps,   B = w
//   wgrad  m = o            n = (c, r, s)  k = (b, oh, ow)   A = dy,        B = x window
#include "common.cuh"

namespace dusty {
namespace {

struct SimtConv {
  int B, C, H, W, O, Ho, Wo, R, S, sh, sw, ph, pw;
  long long x_sb, x_sc, x_sh, x_sw;      // element strides of x  [B, C, H, W]
  long long y_sb, y_sc, y_sh, y_sw;      // ... of y / dy          [B, O, Ho, Wo]
  long long w_so, w_sc, w_sr, w_ss;      // ... of w               [O, C, R, S]
  long long M, N, K;
  float scale;
};

constexpr int kTile = 64, kSlice = 16;

template <typename T, int MODE>
__device__ __forceinline__ float fetch_a(const SimtConv &p, const T *__restrict__ x,
                                         const T *__restrict__ dy, long long m, long long k) {
  if (m >= p.M || k >= p.K) return 0.f;
  if (MODE == 0) {                                   // x[b, oh*sh + r - ph, ow*sw + s - pw, c]
    const int ow = (int)(m % p.Wo);
    long long t = m / p.Wo;
    const int oh = (int)(t % p.Ho), b = (int)(t / p.Ho);
    const int c = (int)(k % p.C);
    const int rs = (int)(k / p.C);
    const int r = rs / p.S, s = rs % p.S;
    const int ih = oh * p.sh + r - p.ph, iw = ow * p.sw + s - p.pw;
    if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) return 0.f;
    return to_f(x[b * p.x_sb + c * p.x_sc + ih * p.x_sh + iw * p.x_sw]);
  } else if (MODE == 1) {                            // dy[b, (ih + ph - r)/sh, (iw + pw - s)/sw, o]
    const int iw = (int)(m % p.W);
    long long t = m / p.W;
    const int ih = (int)(t % p.H), b = (int)(t / p.H);
    const int o = (int)(k % p.O);
    const int rs = (int)(k / p.O);
    const int r = rs / p.S, s = rs % p.S;
    const int nh = ih + p.ph - r, nw = iw + p.pw - s;
    if (nh < 0 || nw < 0 || nh % p.sh || nw % p.sw) return 0.f;
    const int oh = nh / p.sh, ow = nw / p.sw;
    if (oh >= p.Ho || ow >= p.Wo) return 0.f;
    return to_f(dy[b * p.y_sb + o * p.y_sc + oh * p.y_sh + ow * p.y_sw]);
  } else {                                           // dy[b, oh, ow, m]
    const int ow = (int)(k % p.Wo);
    long long t = k / p.Wo;
    const int oh = (int)(t % p.Ho), b = (int)(t / p.Ho);
    return to_f(dy[b * p.y_sb + m * p.y_sc + oh * p.y_sh + ow * p.y_sw]);
  }
}

template <typename T, int MODE>
__device__ __forceinline__ float fetch_b(const SimtConv &p, const T *__restrict__ x,
                                         const T *__restrict__ w, long long n, long long k) {
  if (n >= p.N || k >= p.K) return 0.f;
  if (MODE == 0) {                                   // w[n, c, r, s], k = (r, s, c)
    const int c = (int)(k % p.C);
    const int rs = (int)(k / p.C);
    return to_f(w[n * p.w_so + c * p.w_sc + (rs / p.S) * p.w_sr + (rs % p.S) * p.w_ss]);
  } else if (MODE == 1) {                            // w[o, n, r, s], k = (r, s, o)
    const int o = (int)(k % p.O);
    const int rs = (int)(k / p.O);
    return to_f(w[o * p.w_so + n * p.w_sc + (rs / p.S) * p.w_sr + (rs % p.S) * p.w_ss]);
  } else {                                           // x window of (c, r, s) = n at pixel k
    const int s = (int)(n % p.S);
    long long t = n / p.S;
    const int r = (int)(t % p.R), c = (int)(t / p.R);
    const int ow = (int)(k % p.Wo);
    long long u = k / p.Wo;
    const int oh = (int)(u % p.Ho), b = (int)(u / p.Ho);
    const int ih = oh * p.sh + r - p.ph, iw = ow * p.sw + s - p.pw;
    if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) return 0.f;
    return to_f(x[b * p.x_sb + c * p.x_sc + ih * p.x_sh + iw * p.x_sw]);
  }
}

// TO: output element type (T for fprop / dgrad, float for wgrad)
template <typename T, typename TO, int MODE>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const T *__restrict__ x, const T *__restrict__ dy, const T *__restrict__ w,
                 TO *__restrict__ out, const SimtConv p, int k_splits) {
  __shared__ float As[kSlice][kTile + 1];
  __shared__ float Bs[kSlice][kTile + 1];
  const long long m0 = (long long)blockIdx.x * kTile, n0 = (long long)blockIdx.y * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16 threads, 4 x 4 outputs each
  // split-K (wgrad: K = all pixels): this CTA's slice range
  const long long slices = (p.K + kSlice - 1) / kSlice;
  const long long per = (slices + k_splits - 1) / k_splits;
  const long long s_begin = (long long)blockIdx.z * per;
  const long long s_end = s_begin + per < slices ? s_begin + per : slices;
  float acc[4][4] = {};
  const int lm = threadIdx.x & 63, lk = threadIdx.x >> 6;       // loader: row lm, k rows lk, lk+4, ..
  for (long long sl = s_begin; sl < s_end; ++sl) {
    const long long k0 = sl * kSlice;
#pragma unroll
    for (int i = 0; i < kSlice / 4; ++i) {
      const int kk = lk + 4 * i;
      As[kk][lm] = fetch_a<T, MODE>(p, x, dy, m0 + lm, k0 + kk);
      Bs[kk][lm] = fetch_b<T, MODE>(p, x, w, n0 + lm, k0 + kk);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSlice; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      const float v = acc[i][j] * p.scale;
      if (MODE == 0) {
        const int ow = (int)(m % p.Wo);
        long long t = m / p.Wo;
        const int oh = (int)(t % p.Ho), b = (int)(t / p.Ho);
        out[b * p.y_sb + n * p.y_sc + oh * p.y_sh + ow * p.y_sw] = from_f<TO>(v);
      } else if (MODE == 1) {
        const int iw = (int)(m % p.W);
        long long t = m / p.W;
        const int ih = (int)(t % p.H), b = (int)(t / p.H);
        out[b * p.x_sb + n * p.x_sc + ih * p.x_sh + iw * p.x_sw] = from_f<TO>(v);
      } else {                                         // fp32 [O, C, R, S] contiguous
        float *o = reinterpret_cast<float *>(out) + m * p.N + n;
        if (k_splits > 1) atomicAdd(o, v); else *o = v;
      }
    }
  }
}

// ---- pointwise (1x1, unit stride, no padding) convolutions with a handful of input channels:
// the discriminator stem's 2 -> 32 convolution, which the R1 step differentiates twice through
// the single ops (functional._Stem's composite).  With K = C <= 4 the tiled kernel above loads
// 16 x 64 slices to use a 2 x 64 corner; these are plain streaming passes over the pixels.
// channel-contiguous (NHWC) y / dy: 16-byte accesses along o (same arithmetic, same order)
template <typename T>
__device__ __forceinline__ bool pw_vec_ok(const SimtConv &p, const void *y) {
  constexpr int V = Vec16<T>::N;
  return p.y_sc == 1 && p.O % V == 0 && p.y_sb % V == 0 && p.y_sh % V == 0 && p.y_sw % V == 0 &&
         (reinterpret_cast<uintptr_t>(y) & 15) == 0;
}

template <typename T, int C>
__global__ void pw_fprop_kernel(const T *__restrict__ x, const T *__restrict__ w, T *__restrict__ y,
                                const SimtConv p) {
  __shared__ float ws[64 * C];
  for (int i = threadIdx.x; i < p.O * C; i += blockDim.x)
    ws[i] = to_f(w[(i / C) * p.w_so + (i % C) * p.w_sc]);
  __syncthreads();
  const long long P = (long long)p.B * p.H * p.W;
  const int groups = (p.O + 7) / 8;
  const bool vec = pw_vec_ok<T>(p, y) && p.O % 8 == 0;
  for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < P * groups;
       it += (long long)gridDim.x * blockDim.x) {
    const long long pix = it / groups;
    const int o0 = (int)(it % groups) * 8;
    const int wq = (int)(pix % p.W);
    const long long t = pix / p.W;
    const int h = (int)(t % p.H), b = (int)(t / p.H);
    float xv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) xv[c] = to_f(x[b * p.x_sb + c * p.x_sc + h * p.x_sh + wq * p.x_sw]);
    T *yp = y + b * p.y_sb + h * p.y_sh + wq * p.y_sw;
    if (vec) {                                         // O % 8 == 0 here: the whole group exists
      constexpr int V = Vec16<T>::N;
#pragma unroll
      for (int j0 = 0; j0 < 8; j0 += V) {
        Vec16<T> out;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) acc = fmaf(xv[c], ws[(o0 + j0 + j) * C + c], acc);
          out.set(j, to_f(from_f<T>(acc * p.scale)));
        }
        st16(yp + o0 + j0, out);
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = o0 + j;
      if (o >= p.O) break;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) acc = fmaf(xv[c], ws[o * C + c], acc);
      yp[o * p.y_sc] = from_f<T>(acc * p.scale);
    }
  }
}

template <typename T, int C>
__global__ void pw_dgrad_kernel(const T *__restrict__ dy, const T *__restrict__ w, T *__restrict__ dx,
                                const SimtConv p) {
  __shared__ float ws[64 * C];
  for (int i = threadIdx.x; i < p.O * C; i += blockDim.x)
    ws[i] = to_f(w[(i / C) * p.w_so + (i % C) * p.w_sc]);
  __syncthreads();
  const long long P = (long long)p.B * p.H * p.W;
  const bool vec = pw_vec_ok<T>(p, dy);
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < P;
       pix += (long long)gridDim.x * blockDim.x) {
    const int wq = (int)(pix % p.W);
    const long long t = pix / p.W;
    const int h = (int)(t % p.H), b = (int)(t / p.H);
    const T *gp = dy + b * p.y_sb + h * p.y_sh + wq * p.y_sw;
    float acc[C] = {};
    if (vec) {
      constexpr int V = Vec16<T>::N;
      for (int o0 = 0; o0 < p.O; o0 += V) {
        const Vec16<T> gv = ld16(gp + o0);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float g = gv.get(j);
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(g, ws[(o0 + j) * C + c], acc[c]);
        }
      }
    } else {
      for (int o = 0; o < p.O; ++o) {
        const float g = to_f(gp[o * p.y_sc]);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(g, ws[o * C + c], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
      dx[b * p.x_sb + c * p.x_sc + h * p.x_sh + wq * p.x_sw] = from_f<T>(acc[c] * p.scale);
  }
}

// dw[o, c] = sum_pixels dy[pixel, o] * x[pixel, c]; blockIdx.y picks a 16-wide group of o
template <typename T, int C>
__global__ void __launch_bounds__(256)
pw_wgrad_kernel(const T *__restrict__ dy, const T *__restrict__ x, float *__restrict__ dw, const SimtConv p) {
  const int o0 = blockIdx.y * 16;
  const long long P = (long long)p.B * p.H * p.W;
  const bool vec = pw_vec_ok<T>(p, dy) && p.O % 16 == 0;
  float acc[16][C] = {};
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < P;
       pix += (long long)gridDim.x * blockDim.x) {
    const int wq = (int)(pix % p.W);
    const long long t = pix / p.W;
    const int h = (int)(t % p.H), b = (int)(t / p.H);
    float xv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) xv[c] = to_f(x[b * p.x_sb + c * p.x_sc + h * p.x_sh + wq * p.x_sw]);
    const T *gp = dy + b * p.y_sb + h * p.y_sh + wq * p.y_sw;
    if (vec) {                                          // O % 16 == 0: no partial group
      constexpr int V = Vec16<T>::N;
#pragma unroll
      for (int j0 = 0; j0 < 16; j0 += V) {
        const Vec16<T> gv = ld16(gp + o0 + j0);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float g = gv.get(j);
#pragma unroll
          for (int c = 0; c < C; ++c) acc[j0 + j][c] = fmaf(g, xv[c], acc[j0 + j][c]);
        }
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float g = (o0 + j < p.O) ? to_f(gp[(o0 + j) * p.y_sc]) : 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) acc[j][c] = fmaf(g, xv[c], acc[j][c]);
    }
  }
  __shared__ float red[8][16 * C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 16; ++j)
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float v = acc[j][c];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) red[warp][j * C + c] = v;
    }
  __syncthreads();
  if (threadIdx.x < 16 * C) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    const int j = threadIdx.x / C, c = threadIdx.x % C;
    if (o0 + j < p.O) atomicAdd(dw + (long long)(o0 + j) * C + c, v * p.scale);
  }
}

template <typename T, int C>
int launch_pointwise(int mode, const void *x, const void *dy, const void *w, void *out, const SimtConv &p,
                     cudaStream_t st) {
  const long long P = (long long)p.B * p.H * p.W;
  const int blocks = (int)(P / 256 + 1 < 16LL * num_sms() ? P / 256 + 1 : 16LL * num_sms());
  if (mode == 0) {
    pw_fprop_kernel<T, C><<<blocks, 256, 0, st>>>((const T *)x, (const T *)w, (T *)out, p);
  } else if (mode == 1) {
    pw_dgrad_kernel<T, C><<<blocks, 256, 0, st>>>((const T *)dy, (const T *)w, (T *)out, p);
  } else {
    if (cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(p.O * C), st) != cudaSuccess) {
      set_error("dusty_conv2d_simt: memset failed");
      return DUSTY_ECUDA;
    }
    const int bx = blocks < 4 * num_sms() ? blocks : 4 * num_sms();
    pw_wgrad_kernel<T, C><<<dim3((unsigned)bx, (unsigned)((p.O + 15) / 16)), 256, 0, st>>>(
        (const T *)dy, (const T *)x, (float *)out, p);
  }
  return 0;
}

template <typename T>
int launch_simt(int mode, const void *x, const void *dy, const void *w, void *out, const SimtConv &p,
                cudaStream_t st) {
  if (p.R == 1 && p.S == 1 && p.sh == 1 && p.sw == 1 && p.ph == 0 && p.pw == 0 && p.C <= 4 && p.O <= 64 &&
      p.Ho == p.H && p.Wo == p.W) {
    switch (p.C) {
      case 1: return launch_pointwise<T, 1>(mode, x, dy, w, out, p, st);
      case 2: return launch_pointwise<T, 2>(mode, x, dy, w, out, p, st);
      case 3: return launch_pointwise<T, 3>(mode, x, dy, w, out, p, st);
      default: return launch_pointwise<T, 4>(mode, x, dy, w, out, p, st);
    }
  }
  const long long gm = (p.M + kTile - 1) / kTile, gn = (p.N + kTile - 1) / kTile;
  if (gm > 0x7fffffffLL || gn > 65535) {
    set_error("dusty_conv2d_simt: problem too large for the grid");
    return DUSTY_EINVAL;
  }
  int splits = 1;
  if (mode == 2) {                                    // few output tiles, long reduction
    const long long want = (4LL * num_sms() + gm * gn - 1) / (gm * gn);
    const long long slices = (p.K + kSlice - 1) / kSlice;
    splits = (int)(want < 1 ? 1 : (want > slices ? slices : want));
    if (splits > 1024) splits = 1024;
    if (splits > 1 &&
        cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(p.M * p.N), st) != cudaSuccess) {
      set_error("dusty_conv2d_simt: memset failed");
      return DUSTY_ECUDA;
    }
  }
  dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)splits);
  const T *xp = (const T *)x, *dp = (const T *)dy, *wp = (const T *)w;
  if (mode == 0) conv_simt_kernel<T, T, 0><<<grid, 256, 0, st>>>(xp, dp, wp, (T *)out, p, 1);
  else if (mode == 1) conv_simt_kernel<T, T, 1><<<grid, 256, 0, st>>>(xp, dp, wp, (T *)out, p, 1);
  else conv_simt_kernel<T, float, 2><<<grid, 256, 0, st>>>(xp, dp, wp, (float *)out, p, splits);
  return 0;
}

}  // namespace
}  // namespace dusty

using namespace dusty;

// mode 0: y = conv2d(x, w) * scale; mode 1: dx = conv_transpose2d(dy, w) * scale (the data
// gradient); mode 2: dw (fp32, [O, C, R, S] contiguous) = filter gradient * scale.
// Tensors are addressed through element strides in logical (b, c, h, w) / (o, c, r, s) order:
// NCHW, NHWC and OHWI memory all work without copies.  dtype: element type of x, dy, w and of
// the output of modes 0 / 1.
extern "C" int dusty_conv2d_simt(int mode, const void *x, const void *dy, const void *w, void *out,
                                 int B, int C, int H, int W, int O, int Ho, int Wo, int R, int S,
                                 int stride_h, int stride_w, int pad_h, int pad_w,
                                 const long long *x_strides, const long long *y_strides,
                                 const long long *w_strides, float scale, int dtype, void *stream) {
  DUSTY_CHECK_ARG(mode >= 0 && mode <= 2, "mode: 0 fprop, 1 dgrad, 2 wgrad");
  DUSTY_CHECK_ARG(out && x_strides && y_strides && w_strides, "null pointer");
  DUSTY_CHECK_ARG((mode == 1 || x) && (mode == 0 || dy) && (mode == 2 || w), "missing operand");
  DUSTY_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && O > 0 && Ho > 0 && Wo > 0 && R > 0 && S > 0, "empty tensor");
  DUSTY_CHECK_ARG(stride_h >= 1 && stride_w >= 1 && pad_h >= 0 && pad_w >= 0, "bad stride / padding");
  // fprop / wgrad: every window lies inside the zero-padded input (the data gradient needs no
  // such bound: positions no window reaches receive zero, a smaller dx is a cropped one)
  DUSTY_CHECK_ARG(mode == 1 || ((long long)(Ho - 1) * stride_h + R <= H + 2LL * pad_h &&
                                (long long)(Wo - 1) * stride_w + S <= W + 2LL * pad_w),
                  "output larger than the convolution produces");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  SimtConv p;
  p.B = B; p.C = C; p.H = H; p.W = W; p.O = O; p.Ho = Ho; p.Wo = Wo; p.R = R; p.S = S;
  p.sh = stride_h; p.sw = stride_w; p.ph = pad_h; p.pw = pad_w;
  p.x_sb = x_strides[0]; p.x_sc = x_strides[1]; p.x_sh = x_strides[2]; p.x_sw = x_strides[3];
  p.y_sb = y_strides[0]; p.y_sc = y_strides[1]; p.y_sh = y_strides[2]; p.y_sw = y_strides[3];
  p.w_so = w_strides[0]; p.w_sc = w_strides[1]; p.w_sr = w_strides[2]; p.w_ss = w_strides[3];
  p.scale = scale;
  if (mode == 0) { p.M = (long long)B * Ho * Wo; p.N = O; p.K = (long long)R * S * C; }
  else if (mode == 1) { p.M = (long long)B * H * W; p.N = C; p.K = (long long)R * S * O; }
  else { p.M = O; p.N = (long long)C * R * S; p.K = (long long)B * Ho * Wo; }
  const int rc = dtype == DUSTY_F32 ? launch_simt<float>(mode, x, dy, w, out, p, (cudaStream_t)stream)
                                    : launch_simt<__nv_bfloat16>(mode, x, dy, w, out, p, (cudaStream_t)stream);
  if (rc) return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

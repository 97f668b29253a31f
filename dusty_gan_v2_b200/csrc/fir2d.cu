// a4/a5: polyphase FIR family -- upfirdn2d, Resample (up2 / down2 / blur), BlurVH,
// filter2d and Pad are all instances of
//   out[n,my,mx] = sum_{ty,tx} K[ty,tx] * XU(my*dy + ty - py0, mx*dx + tx - px0)
// with XU the zero-inserted, boundary-extended input.  The boundary extension
// (circular in W for the LiDAR ring, replicate / reflect in H) is folded into the index
// math, so neither the padded nor the zero-inserted tensor is ever materialised.
//
// Forward is a gather over taps that hit the input lattice (taps are stepped by `up`, so
// no multiply-by-zero work); the adjoint is a gather as well (each input element visits
// the extended coordinates that alias onto it), so there are no atomics anywhere.
// Taps live in shared memory; reads of x go through L1 (neighbouring threads share
// almost all of their footprint), writes are coalesced.
#include <climits>

#include "common.cuh"

namespace dusty {

constexpr int kMaxTaps = 1024;

struct FirParams {
  int kh, kw, flip;
  int in_h, in_w, out_h, out_w;
  int up_y, up_x, down_y, down_x;
  int pad_y0, pad_x0;
  int mode_y, mode_x;
  int lo_y, hi_y, lo_x, hi_x;  // range of extended input coordinates ever touched
};

__host__ __device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q;
}
__device__ __forceinline__ int posmod(int a, int b) {
  int r = a % b;
  return r < 0 ? r + b : r;
}

// extended coordinate -> stored index, or -1 when it reads as zero
__device__ __forceinline__ int bmap(int i, int n, int mode) {
  if (mode == DUSTY_PAD_CIRCULAR) {
    if (i < 0) { i += n; if (i < 0) i = posmod(i, n); }
    else if (i >= n) { i -= n; if (i >= n) i = posmod(i, n); }
    return i;
  }
  if (mode == DUSTY_PAD_REPLICATE) return i < 0 ? 0 : (i >= n ? n - 1 : i);
  if (mode == DUSTY_PAD_REFLECT) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return (i < 0 || i >= n) ? -1 : i;
  }
  return (i < 0 || i >= n) ? -1 : i;
}

__device__ __forceinline__ float tap_at(const float *sk, const FirParams &p, int ty, int tx) {
  return p.flip ? sk[(p.kh - 1 - ty) * p.kw + (p.kw - 1 - tx)] : sk[ty * p.kw + tx];
}

// grid = (ceil(out_w / 256), out_h, min(N, 65535)): no per-element div/mod
// UPX / UPY: compile-time up factors (1 or 2) so the polyphase index math needs no integer
// division; 0 = run-time value (generic).
template <typename T, int UPX, int UPY>
__global__ void __launch_bounds__(256)
fir2d_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, const float *__restrict__ taps,
                 FirParams p, int64_t N) {
  __shared__ float sk[kMaxTaps];
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) sk[i] = taps[i];
  __syncthreads();
  const int mx = blockIdx.x * blockDim.x + threadIdx.x;
  const int my = blockIdx.y;
  if (mx >= p.out_w) return;
  const int64_t in_plane = (int64_t)p.in_h * p.in_w;
  const int64_t out_plane = (int64_t)p.out_h * p.out_w;
  const int by = my * p.down_y - p.pad_y0;
  const int bx = mx * p.down_x - p.pad_x0;
  const int up_y = UPY ? UPY : p.up_y, up_x = UPX ? UPX : p.up_x;
  const int ty0 = UPY == 1 ? 0 : (UPY == 2 ? (by & 1) : posmod(-by, up_y));
  const int tx0 = UPX == 1 ? 0 : (UPX == 2 ? (bx & 1) : posmod(-bx, up_x));
  for (int64_t n = blockIdx.z; n < N; n += gridDim.z) {
    const T *xp = x + n * in_plane;
    float acc = 0.f;
    for (int ty = ty0; ty < p.kh; ty += up_y) {
      const int qy = by + ty;                     // exact multiple of up_y
      const int iy = bmap(UPY == 1 ? qy : (UPY == 2 ? (qy >> 1) : qy / up_y), p.in_h, p.mode_y);
      if (iy < 0) continue;
      const T *row = xp + (int64_t)iy * p.in_w;
      float racc = 0.f;
      for (int tx = tx0; tx < p.kw; tx += up_x) {
        const int qx = bx + tx;
        const int ix = bmap(UPX == 1 ? qx : (UPX == 2 ? (qx >> 1) : qx / up_x), p.in_w, p.mode_x);
        if (ix < 0) continue;
        racc = fmaf(tap_at(sk, p, ty, tx), to_f(row[ix]), racc);
      }
      acc += racc;
    }
    y[n * out_plane + (int64_t)my * p.out_w + mx] = from_f<T>(acc);
  }
}

// c-th extended coordinate that aliases onto stored index i (INT_MIN = none / exhausted).
__device__ __forceinline__ int preimage(int c, int i, int n, int mode, int lo, int hi) {
  switch (mode) {
    case DUSTY_PAD_CIRCULAR: {
      int start = i - floordiv(i - lo, n) * n;  // smallest e >= lo, e == i (mod n)
      int e = start + c * n;
      return e <= hi ? e : INT_MIN;
    }
    case DUSTY_PAD_REPLICATE: {
      if (n == 1) {
        int s = lo < 0 ? lo : 0;
        int e = s + c;
        return e <= (hi > 0 ? hi : 0) ? e : INT_MIN;
      }
      if (i == 0) {
        int s = lo < 0 ? lo : 0;
        int e = s + c;
        return e <= 0 ? e : INT_MIN;
      }
      if (i == n - 1) {
        int e = n - 1 + c;
        return e <= (hi > n - 1 ? hi : n - 1) ? e : INT_MIN;
      }
      return c == 0 ? i : INT_MIN;
    }
    case DUSTY_PAD_REFLECT: {
      if (c == 0) return i;
      if (c == 1) return (i > 0 && -i >= lo) ? -i : INT_MIN + 1;      // +1: "skip, keep going"
      if (c == 2) return (i < n - 1 && 2 * (n - 1) - i <= hi) ? 2 * (n - 1) - i : INT_MIN;
      return INT_MIN;
    }
    default:
      return c == 0 ? i : INT_MIN;
  }
}

template <typename T, int DNX, int DNY>
__global__ void __launch_bounds__(256)
fir2d_adj_kernel(const T *__restrict__ dy, T *__restrict__ dx, const float *__restrict__ taps,
                 FirParams p, int64_t N) {
  __shared__ float sk[kMaxTaps];
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) sk[i] = taps[i];
  __syncthreads();
  const int ix = blockIdx.x * blockDim.x + threadIdx.x;
  const int iy = blockIdx.y;
  if (ix >= p.in_w) return;
  const int64_t out_plane = (int64_t)p.out_h * p.out_w;
  const int64_t in_plane = (int64_t)p.in_h * p.in_w;
  for (int64_t n = blockIdx.z; n < N; n += gridDim.z) {
    const T *gp = dy + n * out_plane;
    float acc = 0.f;
    for (int cy = 0;; ++cy) {
      const int ey = preimage(cy, iy, p.in_h, p.mode_y, p.lo_y, p.hi_y);
      if (ey == INT_MIN) break;
      if (ey == INT_MIN + 1) continue;
      const int ny = ey * p.up_y + p.pad_y0;  // = my*down_y + ty
      const int down_y = DNY ? DNY : p.down_y, down_x = DNX ? DNX : p.down_x;
      for (int ty = (DNY == 1 ? 0 : (DNY == 2 ? (ny & 1) : posmod(ny, down_y))); ty < p.kh; ty += down_y) {
        const int my = DNY == 1 ? (ny - ty) : (DNY == 2 ? ((ny - ty) >> 1) : (ny - ty) / down_y);
        if (ny - ty < 0 || my >= p.out_h) continue;
        const T *grow = gp + (int64_t)my * p.out_w;
        for (int cx = 0;; ++cx) {
          const int ex = preimage(cx, ix, p.in_w, p.mode_x, p.lo_x, p.hi_x);
          if (ex == INT_MIN) break;
          if (ex == INT_MIN + 1) continue;
          const int nx = ex * p.up_x + p.pad_x0;
          for (int tx = (DNX == 1 ? 0 : (DNX == 2 ? (nx & 1) : posmod(nx, down_x))); tx < p.kw; tx += down_x) {
            const int mx = DNX == 1 ? (nx - tx) : (DNX == 2 ? ((nx - tx) >> 1) : (nx - tx) / down_x);
            if (nx - tx < 0 || mx >= p.out_w) continue;
            acc = fmaf(tap_at(sk, p, ty, tx), to_f(grow[mx]), acc);
          }
        }
      }
    }
    dx[n * in_plane + (int64_t)iy * p.in_w + ix] = from_f<T>(acc);
  }
}

static int fill_params(FirParams &p, int kh, int kw, int flip, int in_h, int in_w, int out_h,
                       int out_w, int up_y, int up_x, int down_y, int down_x, int pad_y0,
                       int pad_x0, int mode_y, int mode_x) {
  if (kh < 1 || kw < 1 || kh * kw > kMaxTaps) {
    set_error("fir2d: kh*kw must be in [1, %d]", kMaxTaps);
    return DUSTY_EINVAL;
  }
  if (in_h < 1 || in_w < 1 || out_h < 0 || out_w < 0 || up_y < 1 || up_x < 1 || down_y < 1 ||
      down_x < 1) {
    set_error("fir2d: bad geometry");
    return DUSTY_EINVAL;
  }
  if (mode_y < 0 || mode_y > 3 || mode_x < 0 || mode_x > 3) {
    set_error("fir2d: bad boundary mode");
    return DUSTY_EINVAL;
  }
  p.kh = kh; p.kw = kw; p.flip = flip ? 1 : 0;
  p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
  p.up_y = up_y; p.up_x = up_x; p.down_y = down_y; p.down_x = down_x;
  p.pad_y0 = pad_y0; p.pad_x0 = pad_x0; p.mode_y = mode_y; p.mode_x = mode_x;
  p.lo_y = floordiv(-pad_y0, up_y);
  p.hi_y = floordiv((out_h - 1) * down_y + kh - 1 - pad_y0, up_y);
  p.lo_x = floordiv(-pad_x0, up_x);
  p.hi_x = floordiv((out_w - 1) * down_x + kw - 1 - pad_x0, up_x);
  // a single reflection must be enough
  if (mode_y == DUSTY_PAD_REFLECT && (p.lo_y <= -in_h || p.hi_y >= 2 * in_h - 1)) {
    set_error("fir2d: reflect padding larger than the input (H)");
    return DUSTY_EINVAL;
  }
  if (mode_x == DUSTY_PAD_REFLECT && (p.lo_x <= -in_w || p.hi_x >= 2 * in_w - 1)) {
    set_error("fir2d: reflect padding larger than the input (W)");
    return DUSTY_EINVAL;
  }
  return 0;
}

static dim3 grid_for(int w, int h, int64_t N) {
  return dim3((unsigned)((w + 255) / 256), (unsigned)h, (unsigned)(N > 65535 ? 65535 : N));
}

}  // namespace dusty

using namespace dusty;

extern "C" int dusty_fir2d(const void *x, void *y, const float *taps, int kh, int kw, int flip,
                           int64_t N, int in_h, int in_w, int out_h, int out_w, int up_y,
                           int up_x, int down_y, int down_x, int pad_y0, int pad_x0, int mode_y,
                           int mode_x, int dtype, void *stream) {
  DUSTY_CHECK_ARG(x && y && taps, "null pointer");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  FirParams p;
  int rc = fill_params(p, kh, kw, flip, in_h, in_w, out_h, out_w, up_y, up_x, down_y, down_x,
                       pad_y0, pad_x0, mode_y, mode_x);
  if (rc) return rc;
  if (N <= 0 || out_h <= 0 || out_w <= 0) return DUSTY_OK;
  DUSTY_CHECK_ARG(out_h <= 65535, "out_h too large");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid = grid_for(out_w, out_h, N);
#define FIR_FWD(T, UX, UY) \
  fir2d_fwd_kernel<T, UX, UY><<<grid, 256, 0, st>>>((const T *)x, (T *)y, taps, p, N)
#define FIR_FWD_DISPATCH(T)                                   \
  do {                                                        \
    if (up_x == 1 && up_y == 1) FIR_FWD(T, 1, 1);             \
    else if (up_x == 2 && up_y == 1) FIR_FWD(T, 2, 1);        \
    else if (up_x == 1 && up_y == 2) FIR_FWD(T, 1, 2);        \
    else if (up_x == 2 && up_y == 2) FIR_FWD(T, 2, 2);        \
    else FIR_FWD(T, 0, 0);                                    \
  } while (0)
  if (dtype == DUSTY_F32) FIR_FWD_DISPATCH(float);
  else FIR_FWD_DISPATCH(__nv_bfloat16);
#undef FIR_FWD_DISPATCH
#undef FIR_FWD
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_fir2d_adj(const void *dy, void *dx, const float *taps, int kh, int kw,
                               int flip, int64_t N, int in_h, int in_w, int out_h, int out_w,
                               int up_y, int up_x, int down_y, int down_x, int pad_y0, int pad_x0,
                               int mode_y, int mode_x, int dtype, void *stream) {
  DUSTY_CHECK_ARG(dy && dx && taps, "null pointer");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  FirParams p;
  int rc = fill_params(p, kh, kw, flip, in_h, in_w, out_h, out_w, up_y, up_x, down_y, down_x,
                       pad_y0, pad_x0, mode_y, mode_x);
  if (rc) return rc;
  if (N <= 0) return DUSTY_OK;
  DUSTY_CHECK_ARG(in_h <= 65535, "in_h too large");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid = grid_for(in_w, in_h, N);
#define FIR_ADJ(T, DX, DY) \
  fir2d_adj_kernel<T, DX, DY><<<grid, 256, 0, st>>>((const T *)dy, (T *)dx, taps, p, N)
#define FIR_ADJ_DISPATCH(T)                                       \
  do {                                                            \
    if (down_x == 1 && down_y == 1) FIR_ADJ(T, 1, 1);             \
    else if (down_x == 2 && down_y == 1) FIR_ADJ(T, 2, 1);        \
    else if (down_x == 1 && down_y == 2) FIR_ADJ(T, 1, 2);        \
    else if (down_x == 2 && down_y == 2) FIR_ADJ(T, 2, 2);        \
    else FIR_ADJ(T, 0, 0);                                        \
  } while (0)
  if (dtype == DUSTY_F32) FIR_ADJ_DISPATCH(float);
  else FIR_ADJ_DISPATCH(__nv_bfloat16);
#undef FIR_ADJ_DISPATCH
#undef FIR_ADJ
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_upfirdn2d(const void *x, const float *kernel, void *y, int64_t major,
                               int in_h, int in_w, int kh, int kw, int up_x, int up_y, int down_x,
                               int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                               int dtype, void *stream) {
  DUSTY_CHECK_ARG(down_x >= 1 && down_y >= 1, "bad down factor");
  const int out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) / down_y;
  const int out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) / down_x;
  DUSTY_CHECK_ARG(out_h >= 0 && out_w >= 0, "negative output size");
  return dusty_fir2d(x, y, kernel, kh, kw, 1, major, in_h, in_w, out_h, out_w, up_y, up_x, down_y,
                     down_x, pad_y0, pad_x0, DUSTY_PAD_ZERO, DUSTY_PAD_ZERO, dtype, stream);
}

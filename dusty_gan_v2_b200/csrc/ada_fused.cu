// f1: AdaptiveAugment as a device-side op (gans/augment/adaptive_augment.py:271-291, 386-545).
//
// The geometric policies the shipped configs enable (x / y flips, integer and fractional
// translations, the y scale) are AXIS-ALIGNED: the inverse transform is
//     G_inv = [[ax, 0, tx], [0, dy, ty], [0, 0, 1]],
// so every stage of the reference's pipeline
//     pad (circular W / reflect H) -> 2x up (12-tap SYM6, x then y) -> bilinear affine warp ->
//     2x down (x then y) -> colour gain / offset
// is separable, and the whole geometric part is  out_b = A_y(b) X_b A_x(b)^T  with two per-sample
// 1-D operators.  TWO launches instead of seven (pad, up-x, up-y, warp, down-x, down-y, gain): a
// row pass (one warp per image row, the row and its upsampled / warped forms in shared memory) and
// a column pass (one CTA per 32-column tile) that also applies the colour transform; no padded /
// upsampled intermediates in HBM.
//
// Padding.  The reference pads by a data-dependent amount (the batch maximum of what the sampled
// transforms need, get_padding, clamped to W-1 / H-1; it synchronises the host to read it).  A
// sample never reads beyond its own need plus a 6-pixel margin, so any sufficient padding gives
// the same result; here the padding is FIXED at the clamp (W-1 / H-1 per side, evaluated as index
// arithmetic, nothing is materialised), which also reproduces the clamped cases exactly: zero
// beyond the padded range, and the FIR's own edge effects at the end of that range.
//
// 1-D conventions (n = W or H; Np = 3n - 2 padded, Nu = 2 Np upsampled, No = 2 (n + 6) warped):
//   P[m]  = X[map(m - (n - 1))]                    map: circular (x) / reflect (y)
//   U[u]  = sum_t k[11-t] xu[u + t - 6]            xu = zero-inserted P       (upfirdn2d up = 2, pad 6 / 5)
//   V[j]  = (1-l) U[j0] + l U[j0+1]                xp = ((th0 xn + th2 + 1) Nu - 1) / 2, xn = (2j+1)/No - 1
//   Y[x]  = sum_t k[t] V[2x + t + 1]               (upfirdn2d down = 2, pad -1 / -1, flipped taps)
//   th0 = a No / Nu,  th2 = (0.5 a + 2 t - 0.5) 2 / Nu        (a, t) = (ax, tx) or (dy, ty)
// Only the part of U the interpolation reads is evaluated (a staged range of ~No entries, once
// each).  The adjoint runs the stages backwards; its one data-dependent scatter (through the
// interpolation) is two shared-memory atomics per warped sample into the staged range of dU, the
// fixed stencils are gathers.
//
// dusty_ada_sample draws the transforms ON THE DEVICE (Philox, one subsequence per sample, a
// device-resident call counter), so the whole augmentation can live inside a CUDA graph.
#include <curand_kernel.h>

#include "common.cuh"

namespace dusty {
namespace {

constexpr int kTaps = 12;
// SYM6 taps as compile-time constants: inside the unrolled loops they become FMA immediates
// (indexed through __constant__ memory the two polyphase parities of a warp serialised)
__host__ __device__ constexpr float sym6(int t) {
  return t == 0 ? 0.015404109327027373f : t == 1 ? 0.0034907120842174702f : t == 2 ? -0.11799011114819057f
       : t == 3 ? -0.048311742585633f : t == 4 ? 0.4910559419267466f : t == 5 ? 0.787641141030194f
       : t == 6 ? 0.3379294217276218f : t == 7 ? -0.07263752278646252f : t == 8 ? -0.021060292512300564f
       : t == 9 ? 0.04472490177066578f : t == 10 ? 0.0017677118642428036f : -0.007800708325034148f;
}

constexpr int kAdaWarps = 8;
constexpr int kAdaThreads = kAdaWarps * 32;
constexpr int kParams = 8;           // ax, tx, dy, ty, gain, offset, -, -
constexpr int kColTile = 32;         // columns per CTA of the column pass

struct Axis {
  int n, Np, Nu, No;
  float th0, th2;
  bool circular;
  const int *i0_tab;                 // per warped position j: first source tap, interpolation weight
  const float *lam_tab;
  int ulo, ulen;                     // staged range of the upsampled signal: U[ulo, ulo + ulen)
  __device__ __forceinline__ void set(int n_, float a, float t, bool circ) {
    n = n_; Np = 3 * n_ - 2; Nu = 2 * Np; No = 2 * (n_ + 6);
    th0 = a * (float)No / (float)Nu;
    th2 = (0.5f * a + 2.f * t - 0.5f) * 2.f / (float)Nu;
    circular = circ;
  }
  // padded position m in [0, Np) -> signal index; m - (n-1) lies in [-(n-1), 2n-2]: one wrap
  __device__ __forceinline__ int src_index(int m) const {
    const int i = m - (n - 1);
    if (circular) return i < 0 ? i + n : (i >= n ? i - n : i);
    return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i);
  }
  __device__ __forceinline__ void coord(int j, int &i0, float &lam) const {
    const float xn = (2.f * (float)j + 1.f) / (float)No - 1.f;
    const float xs = th0 * xn + th2;
    const float xp = ((xs + 1.f) * (float)Nu - 1.f) * 0.5f;
    const float f0 = floorf(xp);
    lam = xp - f0;
    // far out-of-range coordinates (huge translations) must not overflow the int conversion
    i0 = (int)fminf(fmaxf(f0, -4.f), (float)Nu + 4.f);
  }
  // tables for all warped positions (block-wide) and the range of U they touch (the map is affine
  // in j, hence monotone: the extremes sit at the two ends)
  __device__ __forceinline__ void build(int *i0s, float *lams, int ucap) {
    for (int j = threadIdx.x; j < No; j += blockDim.x) {
      int i0;
      float lam;
      coord(j, i0, lam);
      i0s[j] = i0;
      lams[j] = lam;
    }
    __syncthreads();
    i0_tab = i0s;
    lam_tab = lams;
    const int a = i0s[0], b = i0s[No - 1];
    const int lo = max(0, min(a, b)), hi = min(Nu - 1, max(a, b) + 1);
    ulo = lo;
    ulen = max(0, min(hi - lo + 1, ucap));
  }
};

// U[u] = sum_t k[11-t] xu[u + t - 6]: the six taps that meet non-zero samples of the zero-inserted
// padded signal (PAR = parity of u)
template <int PAR>
__device__ __forceinline__ float up_tap(const float *sig, int stride, const Axis &ax, int u) {
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < kTaps / 2; ++r) {
    const int q = u + PAR + 2 * r - 6;             // even
    if (q >= 0 && q < 2 * ax.Np) acc = fmaf(sym6(kTaps - 1 - (PAR + 2 * r)), sig[ax.src_index(q >> 1) * stride], acc);
  }
  return acc;
}
__device__ __forceinline__ float up_direct(const float *sig, int stride, const Axis &ax, int u) {
  if (u < 0 || u >= ax.Nu) return 0.f;
  return (u & 1) ? up_tap<1>(sig, stride, ax, u) : up_tap<0>(sig, stride, ax, u);
}

// forward 1-D pipeline over one signal (one warp): sig (n, stride) -> sig, through this warp's
// buffers U (staged range of the upsampled signal) and V (No warped samples)
__device__ __forceinline__ void pipe_fwd(float *sig, int stride, const Axis &ax, float *U, float *V, int lane) {
  // 1. the part of the upsampled signal the interpolation reads, each entry once (lanes take
  //    pairs: one even and one odd polyphase branch per lane, no divergence)
  const int base = ax.ulo & ~1;
  for (int k2 = lane; 2 * k2 < ax.ulen + 1; k2 += 32) {
    const int ue = base + 2 * k2, uo = ue + 1;
    if (ue >= ax.ulo && ue < ax.ulo + ax.ulen) U[ue - ax.ulo] = ue < ax.Nu ? up_tap<0>(sig, stride, ax, ue) : 0.f;
    if (uo >= ax.ulo && uo < ax.ulo + ax.ulen) U[uo - ax.ulo] = uo < ax.Nu ? up_tap<1>(sig, stride, ax, uo) : 0.f;
  }
  __syncwarp();
  auto U_at = [&](int u) -> float {
    const int r = u - ax.ulo;
    if (r >= 0 && r < ax.ulen) return U[r];
    return up_direct(sig, stride, ax, u);          // beyond the staged range (|scale| > 1): direct
  };
  for (int j = lane; j < ax.No; j += 32) {
    const int i0 = ax.i0_tab[j];
    const float lam = ax.lam_tab[j];
    V[j] = fmaf(lam, U_at(i0 + 1), (1.f - lam) * U_at(i0));
  }
  __syncwarp();
  for (int x = lane; x < ax.n; x += 32) {
    float y = 0.f;
#pragma unroll
    for (int t = 0; t < kTaps; ++t) y = fmaf(sym6(t), V[2 * x + t + 1], y);
    sig[x * stride] = y;
  }
  __syncwarp();
}

// adjoint: sig holds the gradient w.r.t. Y on entry, w.r.t. X on exit.  The interpolation's
// scatter goes into the staged range of dU (shared-memory atomics, two per warped sample), the
// fixed stencils are gathers.
__device__ __forceinline__ void pipe_adj(float *sig, int stride, const Axis &ax, float *U, float *V, int lane) {
  for (int j = lane; j < ax.No; j += 32) {
    float dv = 0.f;
#pragma unroll
    for (int t = 0; t < kTaps; ++t) {
      const int e = j - 1 - t;
      if (e >= 0 && !(e & 1) && (e >> 1) < ax.n) dv = fmaf(sym6(t), sig[(e >> 1) * stride], dv);
    }
    V[j] = dv;
  }
  for (int r = lane; r < ax.ulen; r += 32) U[r] = 0.f;
  __syncwarp();
  for (int x = lane; x < ax.n; x += 32) sig[x * stride] = 0.f;
  __syncwarp();
  for (int j = lane; j < ax.No; j += 32) {
    const float dv = V[j];
    const int i0 = ax.i0_tab[j];
    const float lam = ax.lam_tab[j];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int u = i0 + s;
      const float w = (s ? lam : 1.f - lam) * dv;
      if (u < 0 || u >= ax.Nu || w == 0.f) continue;
      const int r = u - ax.ulo;
      if (r >= 0 && r < ax.ulen) {
        atomicAdd(&U[r], w);
      } else {                                      // beyond the staged range: straight to the signal
#pragma unroll
        for (int t = 0; t < kTaps; ++t) {
          const int q = u + t - 6;
          if (!(q & 1) && q >= 0 && q < 2 * ax.Np) atomicAdd(&sig[ax.src_index(q >> 1) * stride], w * sym6(kTaps - 1 - t));
        }
      }
    }
  }
  __syncwarp();
  // dX[x] += sum over the (at most three) padded positions m that read x, and the twelve u = 2m - t + 6
  for (int x = lane; x < ax.n; x += 32) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int m;
      bool ok;
      if (ax.circular) {
        m = x + (ax.n - 1) + (c - 1) * ax.n;       // x - 1, x + n - 1, x + 2n - 1
        ok = m >= 0 && m < ax.Np;
      } else {
        // reflect: i = x (always), i = -x (x >= 1), i = 2(n-1) - x (x <= n-2)
        const int i = c == 0 ? x : (c == 1 ? -x : 2 * (ax.n - 1) - x);
        ok = c == 0 || (c == 1 ? x >= 1 : x <= ax.n - 2);
        m = i + (ax.n - 1);
        ok = ok && m >= 0 && m < ax.Np;
      }
      if (!ok) continue;
#pragma unroll
      for (int t = 0; t < kTaps; ++t) {
        const int u = 2 * m - t + 6;
        const int r = u - ax.ulo;
        if (u >= 0 && u < ax.Nu && r >= 0 && r < ax.ulen) acc = fmaf(sym6(kTaps - 1 - t), U[r], acc);
      }
    }
    sig[x * stride] += acc;
  }
  __syncwarp();
}

// Row pass (x, circular): one warp per image row, 8 rows of ONE sample per CTA.
//   forward: dst row = pipeline(src row);  adjoint: the same on gradients (in place allowed)
template <bool ADJ>
__global__ void __launch_bounds__(kAdaThreads)
ada_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, const float *__restrict__ params,
                int H, int W, int ucap) {
  extern __shared__ float smem[];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *pr = params + (int64_t)b * kParams;
  Axis ax;
  ax.set(W, pr[0], pr[1], true);
  float *lams = smem;
  int *i0s = reinterpret_cast<int *>(smem + ax.No);
  float *wbuf = smem + 2 * ax.No + warp * (W + ucap + ax.No);
  ax.build(i0s, lams, ucap);
  const int r = blockIdx.x * kAdaWarps + warp;
  if (r >= H) return;
  float *sig = wbuf, *U = wbuf + W, *V = U + ucap;
  const float *s = src + ((int64_t)b * H + r) * W;
  for (int x = lane; x < W; x += 32) sig[x] = s[x];
  __syncwarp();
  if (ADJ) pipe_adj(sig, 1, ax, U, V, lane);
  else pipe_fwd(sig, 1, ax, U, V, lane);
  float *d = dst + ((int64_t)b * H + r) * W;
  for (int x = lane; x < W; x += 32) d[x] = sig[x];
}

// Column pass (y, reflect) + colour: one CTA per 32-column tile of one sample, the tile in shared
// memory (pitch 33: conflict-free column walks), one warp per column at a time.
//   forward (mode 0 / 2): dst = gain * pipeline(src) (+ offset);  adjoint: dst = pipeline^T(gain * src)
template <bool ADJ>
__global__ void __launch_bounds__(kAdaThreads)
ada_cols_kernel(const float *__restrict__ src, float *__restrict__ dst, const float *__restrict__ params,
                int H, int W, int ucap, int no_offset) {
  extern __shared__ float smem[];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *pr = params + (int64_t)b * kParams;
  const float gain = pr[4], offs = no_offset ? 0.f : pr[5];
  Axis ay;
  ay.set(H, pr[2], pr[3], false);
  constexpr int pitch = kColTile + 1;
  float *lams = smem;
  int *i0s = reinterpret_cast<int *>(smem + ay.No);
  float *tile = smem + 2 * ay.No;
  float *wbuf = tile + H * pitch + warp * (ucap + ay.No);
  const int c0 = blockIdx.x * kColTile;
  const float *s = src + (int64_t)b * H * W;
  for (int i = threadIdx.x; i < H * kColTile; i += kAdaThreads) {
    const int y = i / kColTile, c = i - y * kColTile;
    const float v = (c0 + c < W) ? s[(int64_t)y * W + c0 + c] : 0.f;
    tile[y * pitch + c] = ADJ ? v * gain : v;       // colour adjoint: d(out)/d(img') = gain
  }
  ay.build(i0s, lams, ucap);                        // (ends with a block barrier)
  float *U = wbuf, *V = wbuf + ucap;
  for (int c = warp; c < kColTile && c0 + c < W; c += kAdaWarps) {
    if (ADJ) pipe_adj(tile + c, pitch, ay, U, V, lane);
    else pipe_fwd(tile + c, pitch, ay, U, V, lane);
  }
  __syncthreads();
  float *d = dst + (int64_t)b * H * W;
  for (int i = threadIdx.x; i < H * kColTile; i += kAdaThreads) {
    const int y = i / kColTile, c = i - y * kColTile;
    if (c0 + c < W) {
      const float v = tile[y * pitch + c];
      d[(int64_t)y * W + c0 + c] = ADJ ? v : fmaf(v, gain, offs);
    }
  }
}

// ---- sampling (adaptive_augment.py:386-469 restricted to the axis-aligned policies) -----------
struct AdaPolicy {
  float lr_flip, ud_flip, int_trans, iso_scale, frac_trans;
  float brightness, contrast, luma_flip, hue, saturation;
  float h_trans;
};

struct M4 {
  float m[4][4];
};
__device__ __forceinline__ M4 m4_eye() {
  M4 r;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1.f : 0.f;
  return r;
}
__device__ __forceinline__ M4 m4_mul(const M4 &a, const M4 &b) {
  M4 r;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) s = fmaf(a.m[i][k], b.m[k][j], s);
      r.m[i][j] = s;
    }
  return r;
}

__global__ void __launch_bounds__(1024)
ada_sample_kernel(float *__restrict__ params, const float *__restrict__ p_ptr, unsigned long long seed,
                  unsigned long long *__restrict__ counter, int B, int H, int W, AdaPolicy pol) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long call = *counter;
  if (b < B) {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)b, call * 64ull, &st);
    const float p = fminf(fmaxf(*p_ptr, 0.f), 1.f);
    auto gate = [&](float mul) { return curand_uniform(&st) < fminf(fmaxf(p * mul, 0.f), 1.f); };
    auto coin = [&]() { return (curand(&st) & 1u) ? 1.f : 0.f; };
    // forward transform G = [[a, 0, tx], [0, d, ty]]; candidates are composed on the left
    float a = 1.f, tx = 0.f, d = 1.f, ty = 0.f;
    auto compose = [&](float ma, float mtx, float md, float mty) {
      tx = ma * tx + mtx; a *= ma;
      ty = md * ty + mty; d *= md;
    };
    if (pol.lr_flip > 0.f) {
      const float s = 1.f - 2.f * coin();
      if (gate(pol.lr_flip)) compose(s, 0.f, 1.f, 0.f);
    }
    if (pol.ud_flip > 0.f) {
      const float s = 1.f - 2.f * coin();
      if (gate(pol.ud_flip)) compose(1.f, 0.f, s, 0.f);
    }
    if (pol.int_trans > 0.f) {
      const float t0 = (curand_uniform(&st) - 0.5f) * 0.25f, t1 = (curand_uniform(&st) - 0.5f) * 0.25f;
      if (gate(pol.int_trans)) compose(1.f, rintf(t1 * W), 1.f, rintf(t0 * H) * pol.h_trans);
    }
    if (pol.iso_scale > 0.f) {
      const float s = expf(curand_normal(&st) * 0.2f * 0.6931471805599453f);
      if (gate(pol.iso_scale)) compose(1.f, 0.f, s, 0.f);       // the mirror / reference scale y only
    }
    if (pol.frac_trans > 0.f) {
      const float t0 = curand_normal(&st) * 0.125f, t1 = curand_normal(&st) * 0.125f;
      if (gate(pol.frac_trans)) compose(1.f, t1 * W, 1.f, t0 * H * pol.h_trans);
    }
    // colour (4x4 homogeneous, luma axis v = (1,1,1,0)/sqrt3); one-channel images use the mean row
    M4 C = m4_eye();
    const float v = 0.5773502691896258f;
    if (pol.brightness > 0.f) {
      const float bb = curand_normal(&st) * 0.2f;
      if (gate(pol.brightness)) {
        M4 m = m4_eye();
        m.m[0][3] = m.m[1][3] = m.m[2][3] = bb;
        C = m4_mul(m, C);
      }
    }
    if (pol.contrast > 0.f) {
      const float c = expf(curand_normal(&st) * 0.5f * 0.6931471805599453f);
      if (gate(pol.contrast)) {
        M4 m = m4_eye();
        m.m[0][0] = m.m[1][1] = m.m[2][2] = c;
        C = m4_mul(m, C);
      }
    }
    if (pol.luma_flip > 0.f) {
      const float i = coin();
      if (gate(pol.luma_flip)) {
        M4 m = m4_eye();
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c2 = 0; c2 < 3; ++c2) m.m[r][c2] -= 2.f * v * v * i;
        C = m4_mul(m, C);
      }
    }
    if (pol.hue > 0.f) {
      const float th = (curand_uniform(&st) * 2.f - 1.f) * 3.14159265358979f;
      if (gate(pol.hue)) {
        float sn, cs;
        sincosf(th, &sn, &cs);
        const float cross[3][3] = {{0.f, -v, v}, {v, 0.f, -v}, {-v, v, 0.f}};
        M4 m = m4_eye();
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c2 = 0; c2 < 3; ++c2)
            m.m[r][c2] = cs * (r == c2 ? 1.f : 0.f) + sn * cross[r][c2] + (1.f - cs) * v * v;
        C = m4_mul(m, C);
      }
    }
    if (pol.saturation > 0.f) {
      const float s = expf(curand_normal(&st) * 0.6931471805599453f);
      if (gate(pol.saturation)) {
        M4 m;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) {
            const float o = (r < 3 && c2 < 3) ? v * v : 0.f;
            m.m[r][c2] = o + ((r == c2 ? 1.f : 0.f) - o) * s;
          }
        C = m4_mul(m, C);
      }
    }
    float gain = 0.f, offs = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      gain += (C.m[r][0] + C.m[r][1] + C.m[r][2]) * (1.f / 3.f);
      offs += C.m[r][3] * (1.f / 3.f);
    }
    float *o = params + (int64_t)b * kParams;
    o[0] = 1.f / a; o[1] = -tx / a;                  // inverse of [[a, tx], [0, 1]]
    o[2] = 1.f / d; o[3] = -ty / d;
    o[4] = gain; o[5] = offs; o[6] = 0.f; o[7] = 0.f;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter = call + 1;
}

}  // namespace
}  // namespace dusty

using namespace dusty;

static int rows_ucap(int W) { return 2 * (W + 6) + 16; }          // |x scale| <= 1 staged, beyond: direct
static int cols_ucap(int H) { return 4 * (H + 6) + 16; }          // |y scale| <= 2 staged
static long long rows_smem(int W) {
  const long long No = 2LL * (W + 6);
  return (2 * No + (long long)kAdaWarps * (W + rows_ucap(W) + No)) * (long long)sizeof(float);
}
static long long cols_smem(int H) {
  const long long No = 2LL * (H + 6);
  return (2 * No + (long long)H * (kColTile + 1) + (long long)kAdaWarps * (cols_ucap(H) + No)) * (long long)sizeof(float);
}

extern "C" long long dusty_ada_apply_smem(int H, int W) {
  const long long a = rows_smem(W), b = cols_smem(H);
  return a > b ? a : b;
}

extern "C" int dusty_ada_apply(const float *img, float *out, const float *params, int B, int H, int W,
                               int mode, void *stream) {
  DUSTY_CHECK_ARG(img && out && params, "null pointer");
  DUSTY_CHECK_ARG(B >= 1 && B <= 65535 && H >= 2 && W >= 2, "bad shape");
  DUSTY_CHECK_ARG(mode >= 0 && mode <= 2, "mode: 0 forward, 1 adjoint, 2 forward without the colour offset");
  DUSTY_CHECK_ARG(img != out, "in-place operation is not supported");
  const long long sr = rows_smem(W), sc = cols_smem(H);
  if (sr > 227 * 1024 || sc > 227 * 1024) {
    set_error("dusty_ada_apply: a %d x %d image does not fit the shared-memory buffers", H, W);
    return DUSTY_EUNSUPPORTED;
  }
  static long long configured[4] = {0, 0, 0, 0};
  auto reserve = [&](int slot, const void *fn, long long bytes) -> bool {
    if (bytes <= configured[slot]) return true;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return false;
    configured[slot] = bytes;
    return true;
  };
  if (!reserve(0, (const void *)ada_rows_kernel<false>, sr) || !reserve(1, (const void *)ada_rows_kernel<true>, sr) ||
      !reserve(2, (const void *)ada_cols_kernel<false>, sc) || !reserve(3, (const void *)ada_cols_kernel<true>, sc)) {
    set_error("dusty_ada_apply: cannot reserve shared memory");
    return DUSTY_ECUDA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grows((unsigned)((H + kAdaWarps - 1) / kAdaWarps), (unsigned)B);
  const dim3 gcols((unsigned)((W + kColTile - 1) / kColTile), (unsigned)B);
  if (mode != 1) {
    // rows (img -> out), then columns + colour in place (a CTA owns its tile)
    ada_rows_kernel<false><<<grows, kAdaThreads, (size_t)sr, st>>>(img, out, params, H, W, rows_ucap(W));
    ada_cols_kernel<false><<<gcols, kAdaThreads, (size_t)sc, st>>>(out, out, params, H, W, cols_ucap(H), mode == 2 ? 1 : 0);
  } else {
    ada_cols_kernel<true><<<gcols, kAdaThreads, (size_t)sc, st>>>(img, out, params, H, W, cols_ucap(H), 0);
    ada_rows_kernel<true><<<grows, kAdaThreads, (size_t)sr, st>>>(out, out, params, H, W, rows_ucap(W));
  }
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_ada_sample(float *params, const float *p, unsigned long long seed,
                                unsigned long long *counter, int B, int H, int W, const float *policy,
                                void *stream) {
  DUSTY_CHECK_ARG(params && p && counter && policy, "null pointer");
  DUSTY_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "bad shape");
  AdaPolicy pol;
  pol.lr_flip = policy[0]; pol.ud_flip = policy[1]; pol.int_trans = policy[2]; pol.iso_scale = policy[3];
  pol.frac_trans = policy[4]; pol.brightness = policy[5]; pol.contrast = policy[6]; pol.luma_flip = policy[7];
  pol.hue = policy[8]; pol.saturation = policy[9]; pol.h_trans = policy[10];
  // one block: the call counter is advanced after every sample has read it
  DUSTY_CHECK_ARG(B <= 1024, "at most 1024 samples per call");
  ada_sample_kernel<<<1, B > 128 ? 1024 : 128, 0, (cudaStream_t)stream>>>(
      params, p, seed, counter, B, H, W, pol);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

// a4: fast path of the Resample family for the 4-tap separable windows the models use
// ([1,3,3,1]): same-size blur (D residual blocks) and 2x upsampling (G synthesis blocks),
// circular in W (LiDAR ring), edge-replicated in H, forward and exact adjoint.
//
// HBM-bound design: one thread owns one 16-byte column group (8 bf16 / 4 fp32 pixels) of
// one image and slides down a strip of rows.  Per input row it does ONE coalesced 16-byte
// load plus the few halo pixels (L1 hits: the neighbouring threads load them as their main
// vector), runs the horizontal 4-tap pass in registers, and keeps the last rows of that
// pass in a register window for the vertical pass, so every input element is fetched from
// DRAM once and the horizontally filtered intermediate never exists in memory.  Stores are
// full 16-byte vectors.  No atomics in the adjoints: each thread gathers the extended rows
// that fold onto its output row.
//
// Closed forms (taps k0..k3; verified against the reference):
//   blur  : out[n]    = k0 x[n-2] + k1 x[n-1] + k2 x[n] + k3 x[n+1]
//   up2   : out[2m]   = k0 x[m-1] + k2 x[m],   out[2m+1] = k1 x[m] + k3 x[m+1]
//   blur^T: g[e]      = k0 d[e+2] + k1 d[e+1] + k2 d[e] + k3 d[e-1]
//   up2^T : g[e]      = k0 d[2e+2] + k1 d[2e+1] + k2 d[2e] + k3 d[2e-1]
// with x extended circularly (W) / clamped (H), d zero outside its range in H and circular
// in W, and the adjoint rows e in [-2, H] (blur) or [-1, H] (up2) folded by clamping.
#include "common.cuh"

namespace dusty {

struct Taps4 { float k[4]; };

template <typename T> struct RowIO {
  static constexpr int V = Vec16<T>::N;
  // dst[j] = row[(x0 - L + j) mod W], j in [0, L + V + R)
  template <int L, int R>
  static __device__ __forceinline__ void load(const T *__restrict__ row, int x0, int W, float *dst) {
    Vec16<T> v = ld16(row + x0);
#pragma unroll
    for (int j = 0; j < V; ++j) dst[L + j] = v.get(j);
#pragma unroll
    for (int j = 0; j < L; ++j) {
      int c = x0 - L + j;
      if (c < 0) c += W;
      dst[j] = to_f(row[c]);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int c = x0 + V + j;
      if (c >= W) c -= W;
      dst[L + V + j] = to_f(row[c]);
    }
  }
  static __device__ __forceinline__ void store(T *__restrict__ row, int x0, const float *src) {
    Vec16<T> v;
#pragma unroll
    for (int j = 0; j < V; ++j) v.set(j, src[j]);
    st16(row + x0, v);
  }
};

// ------------------------------------------------------------------ blur forward
// thread -> (image n, vector column); strip of output rows [y0, y1)
template <typename T>
__global__ void __launch_bounds__(128)
blur4_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, Taps4 t, int H, int W, int strip,
                 int64_t n_threads) {
  constexpr int V = RowIO<T>::V;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_threads) return;
  const int vpr = W / V;
  const int vc = (int)(tid % vpr);
  const int64_t n = tid / vpr;
  const int x0 = vc * V;
  const int y0 = blockIdx.y * strip;
  const int y1 = min(y0 + strip, H);
  const T *img = x + n * (int64_t)H * W;
  T *out = y + n * (int64_t)H * W;
  float a[V], b[V], c[V], d[V];   // horizontally filtered rows r-2, r-1, r, r+1
  auto hpass = [&](int r, float *dst) {
    r = r < 0 ? 0 : (r >= H ? H - 1 : r);
    float s[V + 3];
    RowIO<T>::template load<2, 1>(img + (int64_t)r * W, x0, W, s);
#pragma unroll
    for (int j = 0; j < V; ++j)
      dst[j] = fmaf(t.k[3], s[j + 3], fmaf(t.k[2], s[j + 2], fmaf(t.k[1], s[j + 1], t.k[0] * s[j])));
  };
  hpass(y0 - 2, a);
  hpass(y0 - 1, b);
  hpass(y0, c);
  for (int r = y0; r < y1; ++r) {
    hpass(r + 1, d);
    float o[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      o[j] = fmaf(t.k[3], d[j], fmaf(t.k[2], c[j], fmaf(t.k[1], b[j], t.k[0] * a[j])));
      a[j] = b[j]; b[j] = c[j]; c[j] = d[j];
    }
    RowIO<T>::store(out + (int64_t)r * W, x0, o);
  }
}

// ------------------------------------------------------------------ blur adjoint
template <typename T>
__global__ void __launch_bounds__(128)
blur4_adj_kernel(const T *__restrict__ dy, T *__restrict__ dx, Taps4 t, int H, int W, int strip,
                 int64_t n_threads) {
  constexpr int V = RowIO<T>::V;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_threads) return;
  const int vpr = W / V;
  const int vc = (int)(tid % vpr);
  const int64_t n = tid / vpr;
  const int x0 = vc * V;
  const int y0 = blockIdx.y * strip;
  const int y1 = min(y0 + strip, H);
  const T *g = dy + n * (int64_t)H * W;
  T *out = dx + n * (int64_t)H * W;
  // horizontal transposed pass of gradient row r (zero outside [0,H)):
  //   Gh[x] = k0 d[x+2] + k1 d[x+1] + k2 d[x] + k3 d[x-1]
  auto hpass = [&](int r, float *dst) {
    if (r < 0 || r >= H) {
#pragma unroll
      for (int j = 0; j < V; ++j) dst[j] = 0.f;
      return;
    }
    float s[V + 3];
    RowIO<T>::template load<1, 2>(g + (int64_t)r * W, x0, W, s);   // s[j] = d[x0 - 1 + j]
#pragma unroll
    for (int j = 0; j < V; ++j)
      dst[j] = fmaf(t.k[0], s[j + 3], fmaf(t.k[1], s[j + 2], fmaf(t.k[2], s[j + 1], t.k[3] * s[j])));
  };
  // g(e) = k0 Gh[e+2] + k1 Gh[e+1] + k2 Gh[e] + k3 Gh[e-1]; output row i sums g over the
  // extended rows that clamp onto it: {i} plus {-2,-1} for i == 0 and {H} for i == H-1.
  const int e_lo = (y0 == 0) ? -2 : y0;
  const int e_hi = (y1 == H) ? H : y1 - 1;
  float a[V], b[V], c[V], d[V];   // Gh rows e-1, e, e+1, e+2
  hpass(e_lo - 1, a);
  hpass(e_lo, b);
  hpass(e_lo + 1, c);
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  for (int e = e_lo; e <= e_hi; ++e) {
    hpass(e + 2, d);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      acc[j] += fmaf(t.k[0], d[j], fmaf(t.k[1], c[j], fmaf(t.k[2], b[j], t.k[3] * a[j])));
      a[j] = b[j]; b[j] = c[j]; c[j] = d[j];
    }
    const int i = e < 0 ? 0 : (e >= H ? H - 1 : e);
    const int i_next = (e + 1) < 0 ? 0 : ((e + 1) >= H ? H - 1 : (e + 1));
    if (e == e_hi || i_next != i) {
      RowIO<T>::store(out + (int64_t)i * W, x0, acc);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = 0.f;
    }
  }
}

// ------------------------------------------------------------------ up2 forward
// thread -> (image, input vector column); strip of INPUT rows [r0, r1) -> output rows 2r, 2r+1
// SUMSQ: also accumulate sum(y^2) (fp32, of the un-rounded outputs) into *sumsq -- the statistic
// ModConv2d's EMA normaliser takes of its input (style.py:99-102), which for conv1 of a synthesis
// block IS this kernel's output: one atomicAdd per CTA instead of a second pass over 4N elements.
template <typename T, bool SUMSQ>
__global__ void __launch_bounds__(128)
up2_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, Taps4 t, int H, int W, int strip,
               int64_t n_threads, float *__restrict__ sumsq) {
  constexpr int V = RowIO<T>::V;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float ss = 0.f;
  if (tid < n_threads) {
  const int vpr = W / V;
  const int vc = (int)(tid % vpr);
  const int64_t n = tid / vpr;
  const int x0 = vc * V;
  const int r0 = blockIdx.y * strip;
  const int r1 = min(r0 + strip, H);
  const int W2 = 2 * W;
  const T *img = x + n * (int64_t)H * W;
  T *out = y + n * (int64_t)(2 * H) * W2;
  // horizontal pass of input row r -> 2V outputs
  auto hpass = [&](int r, float *dst) {
    r = r < 0 ? 0 : (r >= H ? H - 1 : r);
    float s[V + 2];
    RowIO<T>::template load<1, 1>(img + (int64_t)r * W, x0, W, s);   // s[j] = x[x0 - 1 + j]
#pragma unroll
    for (int j = 0; j < V; ++j) {
      dst[2 * j] = fmaf(t.k[2], s[j + 1], t.k[0] * s[j]);
      dst[2 * j + 1] = fmaf(t.k[3], s[j + 2], t.k[1] * s[j + 1]);
    }
  };
  float a[2 * V], b[2 * V], c[2 * V];   // rows r-1, r, r+1
  hpass(r0 - 1, a);
  hpass(r0, b);
  for (int r = r0; r < r1; ++r) {
    hpass(r + 1, c);
    float o[2 * V];
#pragma unroll
    for (int j = 0; j < 2 * V; ++j) {
      o[j] = fmaf(t.k[2], b[j], t.k[0] * a[j]);
      if (SUMSQ) ss = fmaf(o[j], o[j], ss);
    }
    RowIO<T>::store(out + (int64_t)(2 * r) * W2, 2 * x0, o);
    RowIO<T>::store(out + (int64_t)(2 * r) * W2, 2 * x0 + V, o + V);
#pragma unroll
    for (int j = 0; j < 2 * V; ++j) {
      o[j] = fmaf(t.k[3], c[j], t.k[1] * b[j]);
      if (SUMSQ) ss = fmaf(o[j], o[j], ss);
      a[j] = b[j]; b[j] = c[j];
    }
    RowIO<T>::store(out + (int64_t)(2 * r + 1) * W2, 2 * x0, o);
    RowIO<T>::store(out + (int64_t)(2 * r + 1) * W2, 2 * x0 + V, o + V);
  }
  }
  if (SUMSQ) {
    __shared__ float part[4];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(sumsq, (part[0] + part[1]) + (part[2] + part[3]));
  }
}

// ------------------------------------------------------------------ up2 adjoint
// dy: [N, 2H, 2W] -> dx: [N, H, W]; thread -> (image, dx vector column), strip of dx rows
template <typename T>
__global__ void __launch_bounds__(128)
up2_adj_kernel(const T *__restrict__ dy, T *__restrict__ dx, Taps4 t, int H, int W, int strip,
               int64_t n_threads) {
  constexpr int V = RowIO<T>::V;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_threads) return;
  const int vpr = W / V;
  const int vc = (int)(tid % vpr);
  const int64_t n = tid / vpr;
  const int x0 = vc * V;
  const int y0 = blockIdx.y * strip;
  const int y1 = min(y0 + strip, H);
  const int W2 = 2 * W, H2 = 2 * H;
  const T *g = dy + n * (int64_t)H2 * W2;
  T *out = dx + n * (int64_t)H * W;
  // Gh[row][ix] = k0 d[2ix+2] + k1 d[2ix+1] + k2 d[2ix] + k3 d[2ix-1]   (circular in 2W)
  auto hpass = [&](int r, float *dst) {
    if (r < 0 || r >= H2) {
#pragma unroll
      for (int j = 0; j < V; ++j) dst[j] = 0.f;
      return;
    }
    const T *row = g + (int64_t)r * W2;
    float s[2 * V + 2];                       // s[j] = d[2*x0 - 1 + j]
    {
      Vec16<T> v0 = ld16(row + 2 * x0), v1 = ld16(row + 2 * x0 + V);
#pragma unroll
      for (int j = 0; j < V; ++j) { s[1 + j] = v0.get(j); s[1 + V + j] = v1.get(j); }
      int cl = 2 * x0 - 1; if (cl < 0) cl += W2;
      int cr = 2 * x0 + 2 * V; if (cr >= W2) cr -= W2;
      s[0] = to_f(row[cl]);
      s[2 * V + 1] = to_f(row[cr]);
    }
#pragma unroll
    for (int j = 0; j < V; ++j)
      dst[j] = fmaf(t.k[0], s[2 * j + 3], fmaf(t.k[1], s[2 * j + 2], fmaf(t.k[2], s[2 * j + 1], t.k[3] * s[2 * j])));
  };
  // g(e) = k0 Gh[2e+2] + k1 Gh[2e+1] + k2 Gh[2e] + k3 Gh[2e-1], e in [-1, H] folded by clamp
  const int e_lo = (y0 == 0) ? -1 : y0;
  const int e_hi = (y1 == H) ? H : y1 - 1;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  float pm1[V], p0[V], p1[V], p2[V];
  hpass(2 * e_lo - 1, pm1);
  hpass(2 * e_lo, p0);
  for (int e = e_lo; e <= e_hi; ++e) {
    hpass(2 * e + 1, p1);
    hpass(2 * e + 2, p2);
#pragma unroll
    for (int j = 0; j < V; ++j)
      acc[j] += fmaf(t.k[0], p2[j], fmaf(t.k[1], p1[j], fmaf(t.k[2], p0[j], t.k[3] * pm1[j])));
    const int i = e < 0 ? 0 : (e >= H ? H - 1 : e);
    const int i_next = (e + 1) < 0 ? 0 : ((e + 1) >= H ? H - 1 : (e + 1));
    if (e == e_hi || i_next != i) {
      RowIO<T>::store(out + (int64_t)i * W, x0, acc);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) { pm1[j] = p1[j]; p0[j] = p2[j]; }
  }
}

template <typename T>
static int launch_resample4(const void *x, void *y, Taps4 t, int64_t N, int H, int W, int up,
                            int adjoint, cudaStream_t st, float *sumsq = nullptr) {
  constexpr int V = Vec16<T>::N;
  const int64_t n_threads = N * (W / V);
  // strips: enough CTAs to fill the GPU, at least 8 rows per strip to amortise the halo
  int strip = H;
  const int64_t ctas_x = (n_threads + 127) / 128;
  while (strip > 8 && ctas_x * ((H + strip - 1) / strip) < (int64_t)num_sms() * 8) strip = (strip + 1) / 2;
  dim3 grid((unsigned)ctas_x, (unsigned)((H + strip - 1) / strip));
  const T *xp = (const T *)x;
  T *yp = (T *)y;
  if (up == 1 && !adjoint) blur4_fwd_kernel<T><<<grid, 128, 0, st>>>(xp, yp, t, H, W, strip, n_threads);
  else if (up == 1) blur4_adj_kernel<T><<<grid, 128, 0, st>>>(xp, yp, t, H, W, strip, n_threads);
  else if (!adjoint && sumsq) up2_fwd_kernel<T, true><<<grid, 128, 0, st>>>(xp, yp, t, H, W, strip, n_threads, sumsq);
  else if (!adjoint) up2_fwd_kernel<T, false><<<grid, 128, 0, st>>>(xp, yp, t, H, W, strip, n_threads, nullptr);
  else up2_adj_kernel<T><<<grid, 128, 0, st>>>(xp, yp, t, H, W, strip, n_threads);
  return 0;
}

}  // namespace dusty

using namespace dusty;

extern "C" int dusty_resample4(const void *x, void *y, float k0, float k1, float k2, float k3,
                               int64_t N, int H, int W, int up, int adjoint, int dtype,
                               void *stream) {
  DUSTY_CHECK_ARG(x && y, "null pointer");
  DUSTY_CHECK_ARG(up == 1 || up == 2, "up must be 1 (blur) or 2 (2x upsample)");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  DUSTY_CHECK_ARG(N >= 1 && H >= 2 && W >= 4, "bad shape");
  const int V = dtype == DUSTY_F32 ? 4 : 8;
  DUSTY_CHECK_ARG(W % V == 0, "W must be a multiple of the 16-byte vector width");
  DUSTY_CHECK_ARG(aligned16(x) && aligned16(y), "tensors must be 16-byte aligned");
  Taps4 t;
  t.k[0] = k0; t.k[1] = k1; t.k[2] = k2; t.k[3] = k3;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = dtype == DUSTY_F32 ? launch_resample4<float>(x, y, t, N, H, W, up, adjoint, st)
                              : launch_resample4<__nv_bfloat16>(x, y, t, N, H, W, up, adjoint, st);
  if (rc) return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_up2_sumsq(const void *x, void *y, float *sumsq, float k0, float k1, float k2,
                               float k3, int64_t N, int H, int W, int dtype, void *stream) {
  DUSTY_CHECK_ARG(x && y && sumsq, "null pointer");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  DUSTY_CHECK_ARG(N >= 1 && H >= 2 && W >= 4, "bad shape");
  const int V = dtype == DUSTY_F32 ? 4 : 8;
  DUSTY_CHECK_ARG(W % V == 0, "W must be a multiple of the 16-byte vector width");
  DUSTY_CHECK_ARG(aligned16(x) && aligned16(y), "tensors must be 16-byte aligned");
  Taps4 t;
  t.k[0] = k0; t.k[1] = k1; t.k[2] = k2; t.k[3] = k3;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = dtype == DUSTY_F32 ? launch_resample4<float>(x, y, t, N, H, W, 2, 0, st, sumsq)
                              : launch_resample4<__nv_bfloat16>(x, y, t, N, H, W, 2, 0, st, sumsq);
  if (rc) return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

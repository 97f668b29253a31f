// a4/a11: ring padding for the discriminator's 3x3 convolutions (Pad, common.py:10-24):
// circular / replicate / reflect in W, replicate / reflect in H, forward and exact adjoint.
// Pure data movement: each thread produces two adjacent output elements (one 4-byte store
// for bf16 when the row pitch allows), index math hoisted per row (2-D grid, no div/mod).
#include "common.cuh"

namespace dusty {

struct PadParams {
  int H, W, Ho, Wo, pt, pb, pl, pr, mode_y, mode_x;
};

__device__ __forceinline__ int pad_map(int i, int n, int mode) {
  if (mode == DUSTY_PAD_CIRCULAR) return i < 0 ? i + n : (i >= n ? i - n : i);
  if (mode == DUSTY_PAD_REFLECT) return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i);
  return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

template <typename T>
__global__ void __launch_bounds__(128)
pad2d_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, PadParams p, int64_t N) {
  const int ox = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int oy = blockIdx.y;
  if (ox >= p.Wo) return;
  const int iy = pad_map(oy - p.pt, p.H, p.mode_y);
  const int ix0 = pad_map(ox - p.pl, p.W, p.mode_x);
  const bool two = ox + 1 < p.Wo;
  const int ix1 = two ? pad_map(ox + 1 - p.pl, p.W, p.mode_x) : 0;
  const int64_t in_plane = (int64_t)p.H * p.W, out_plane = (int64_t)p.Ho * p.Wo;
  const bool pair_store = two && sizeof(T) == 2 && (p.Wo % 2 == 0);
  // four planes per iteration: eight independent loads in flight per thread before the stores
  // (one plane per thread and 400 k tiny blocks ran at 0.16 of the HBM peak)
  const int64_t zs = gridDim.z;
  for (int64_t n = blockIdx.z; n < N; n += 4 * zs) {
    T v0[4], v1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t nn = n + u * zs;
      if (nn < N) {
        const T *row = x + nn * in_plane + (int64_t)iy * p.W;
        v0[u] = row[ix0];
        if (two) v1[u] = row[ix1];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t nn = n + u * zs;
      if (nn >= N) break;
      T *orow = y + nn * out_plane + (int64_t)oy * p.Wo + ox;
      if (two) {
        if (pair_store) {
          uint32_t w = (uint32_t)(*reinterpret_cast<const uint16_t *>(&v0[u])) |
                       ((uint32_t)(*reinterpret_cast<const uint16_t *>(&v1[u])) << 16);
          *reinterpret_cast<uint32_t *>(orow) = w;
        } else {
          orow[0] = v0[u];
          orow[1] = v1[u];
        }
      } else {
        orow[0] = v0[u];
      }
    }
  }
}

// candidate output coordinates that read input coordinate i: the interior one plus halo ones
__device__ __forceinline__ int pad_preimages(int i, int n, int p0, int p1, int mode, int *out) {
  int c = 0;
  out[c++] = i + p0;
  if (mode == DUSTY_PAD_CIRCULAR) {
    if (i >= n - p0) out[c++] = i - (n - p0);         // left halo
    if (i < p1) out[c++] = p0 + n + i;                // right halo
  } else if (mode == DUSTY_PAD_REFLECT) {
    if (i >= 1 && i <= p0) out[c++] = p0 - i;
    if (i <= n - 2 && i >= n - 1 - p1) out[c++] = p0 + 2 * (n - 1) - i;
  } else {
    if (i == 0) for (int j = 0; j < p0; ++j) out[c++] = j;
    if (i == n - 1) for (int j = 0; j < p1; ++j) out[c++] = p0 + n + j;
  }
  return c;
}

constexpr int kMaxPadFast = 4;   // halo width handled by the fast kernels

template <typename T>
__global__ void __launch_bounds__(128)
pad2d_adj_kernel(const T *__restrict__ dy, T *__restrict__ dx, PadParams p, int64_t N) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x;
  const int iy = blockIdx.y;
  if (ix >= p.W) return;
  int ys[2 + 2 * kMaxPadFast], xs[2 + 2 * kMaxPadFast];
  const int ny = pad_preimages(iy, p.H, p.pt, p.pb, p.mode_y, ys);
  const int nx = pad_preimages(ix, p.W, p.pl, p.pr, p.mode_x, xs);
  const int64_t in_plane = (int64_t)p.H * p.W, out_plane = (int64_t)p.Ho * p.Wo;
  const bool simple = ny == 1 && nx == 1;       // interior element: one pre-image
  const int64_t zs = gridDim.z;
  for (int64_t n = blockIdx.z; n < N; n += 4 * zs) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t nn = n + u * zs;
      if (nn >= N) break;
      const T *g = dy + nn * out_plane;
      if (simple) {
        acc[u] = to_f(g[(int64_t)ys[0] * p.Wo + xs[0]]);
      } else {
        for (int a = 0; a < ny; ++a)
          for (int b = 0; b < nx; ++b) acc[u] += to_f(g[(int64_t)ys[a] * p.Wo + xs[b]]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t nn = n + u * zs;
      if (nn >= N) break;
      dx[nn * in_plane + (int64_t)iy * p.W + ix] = from_f<T>(acc[u]);
    }
  }
}

}  // namespace dusty

using namespace dusty;

extern "C" int dusty_pad2d(const void *x, void *y, int64_t N, int H, int W, int pt, int pb, int pl,
                           int pr, int mode_y, int mode_x, int adjoint, int dtype, void *stream) {
  DUSTY_CHECK_ARG(x && y, "null pointer");
  DUSTY_CHECK_ARG(N >= 1 && H >= 1 && W >= 1, "bad shape");
  DUSTY_CHECK_ARG(pt >= 0 && pb >= 0 && pl >= 0 && pr >= 0, "pads must be non-negative");
  DUSTY_CHECK_ARG(pt <= kMaxPadFast && pb <= kMaxPadFast && pl <= kMaxPadFast && pr <= kMaxPadFast,
                  "pad wider than the fast path supports (use dusty_fir2d)");
  DUSTY_CHECK_ARG(pt < H && pb < H && pl < W && pr < W, "pad must be smaller than the image");
  DUSTY_CHECK_ARG(mode_y == DUSTY_PAD_REPLICATE || mode_y == DUSTY_PAD_REFLECT, "bad mode_y");
  DUSTY_CHECK_ARG(mode_x >= DUSTY_PAD_CIRCULAR && mode_x <= DUSTY_PAD_REFLECT, "bad mode_x");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  PadParams p{H, W, H + pt + pb, W + pl + pr, pt, pb, pl, pr, mode_y, mode_x};
  DUSTY_CHECK_ARG(p.Ho <= 65535, "image too tall");
  cudaStream_t st = (cudaStream_t)stream;
  // planes are strided over grid.z, four per thread
  const unsigned gz = (unsigned)(N >= 64 ? (N + 7) / 8 > 65535 ? 65535 : (N + 7) / 8 : N);
  if (!adjoint) {
    dim3 grid((unsigned)((p.Wo + 255) / 256), (unsigned)p.Ho, gz);
    if (dtype == DUSTY_F32) pad2d_fwd_kernel<float><<<grid, 128, 0, st>>>((const float *)x, (float *)y, p, N);
    else pad2d_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, p, N);
  } else {
    dim3 grid((unsigned)((W + 127) / 128), (unsigned)H, gz);
    if (dtype == DUSTY_F32) pad2d_adj_kernel<float><<<grid, 128, 0, st>>>((const float *)x, (float *)y, p, N);
    else pad2d_adj_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, p, N);
  }
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

// f4: LiDAR scan -> range image on the device.  Replaces the numba `scatter` loop and the
// torchvision NEAREST resize of gans/datasets/kitti.py:216-220,275-279,317-370: the reference
// sorts the points by decreasing depth and writes them one by one into a [H, W, 6] image, so
// every cell ends up holding its NEAREST point; then it keeps every (W / W_out)-th column and
// multiplies by the validity channel.
//   pass 1: one thread per point, atomicMin of (depth bits << 32 | point index) on its cell
//   pass 2: one thread per OUTPUT pixel: winner of cell (h, w * step) -> 6 channels, CHW, masked
// The integer cell coordinates and the float32 depths come from the host (numpy, the reference's
// own arithmetic: the float32 arctan2 decides columns and the float32 norm decides which point is
// nearest; only the same library reproduces them bit for bit).
#include "common.cuh"

namespace dusty {
namespace {

__global__ void scan_keys_kernel(const float *__restrict__ depth, const int *__restrict__ cell_h,
                                 const int *__restrict__ cell_w, unsigned long long *__restrict__ keys,
                                 int N, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int h = cell_h[i];
  const int w = cell_w[i];
  if (h < 0) h += H;                                   // numpy-style wrap of the reference's ring index -1
  if (h < 0 || h >= H || w < 0 || w >= W) return;
  const float d = depth[i];                            // host float32 norm (decides ties exactly)
  const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
  atomicMin(keys + (size_t)h * W + w, key);            // depth >= 0: float order == integer order
}

// out: [6, H, W_out] = (x, y, z, reflectance, depth, mask) * mask
__global__ void scan_gather_kernel(const float *__restrict__ pts, const unsigned long long *__restrict__ keys,
                                   float *__restrict__ out, int H, int W, int W_out, int step,
                                   float min_depth, float max_depth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W_out) return;
  const int h = i / W_out, wo = i % W_out;
  const unsigned long long key = keys[(size_t)h * W + (size_t)wo * step];
  float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (key != ~0ull) {
    const unsigned p = (unsigned)(key & 0xffffffffu);
    const float d = __uint_as_float((unsigned)(key >> 32));
    const float m = (d >= min_depth && d <= max_depth) ? 1.f : 0.f;
    v[0] = pts[4 * p] * m; v[1] = pts[4 * p + 1] * m; v[2] = pts[4 * p + 2] * m; v[3] = pts[4 * p + 3] * m;
    v[4] = d * m; v[5] = m;
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) out[(size_t)c * H * W_out + i] = v[c];
}

}  // namespace
}  // namespace dusty

using namespace dusty;

extern "C" int dusty_scan_project(const float *points, const float *depth, const int *cell_h,
                                  const int *cell_w, unsigned long long *keys, float *out, int N, int H, int W,
                                  int W_out, float min_depth, float max_depth, void *stream) {
  DUSTY_CHECK_ARG(points && depth && cell_h && cell_w && keys && out, "null pointer");
  DUSTY_CHECK_ARG(N >= 0 && H > 0 && W > 0 && W_out > 0 && W % W_out == 0, "W must be a multiple of W_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * (size_t)H * W, st) != cudaSuccess) {
    set_error("dusty_scan_project: memset failed");
    return DUSTY_ECUDA;
  }
  if (N > 0) {
    scan_keys_kernel<<<(N + 255) / 256, 256, 0, st>>>(depth, cell_h, cell_w, keys, N, H, W);
    DUSTY_LAUNCH_CHECK();
  }
  scan_gather_kernel<<<(H * W_out + 255) / 256, 256, 0, st>>>(points, keys, out, H, W, W_out, W / W_out, min_depth,
                                                            max_depth);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

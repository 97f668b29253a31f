// a5/a6: the band-limited geometric pipeline of AdaptiveAugment
// (adaptive_augment.py:496-535): SYM6 2x up-sampling in x then y, bilinear affine warp,
// 2x down-sampling in x then y -- as lean single-axis FIR kernels plus one warp kernel.
//
//  * fir1d: zero-padded polyphase FIR along ONE axis (upfirdn2d with a [1,k] or [k,1]
//    kernel), up/down factors compile-time (no integer division), flipped taps staged once
//    in shared memory; the x variant also stages its input row segment (+halo) in shared
//    memory, the y variant reads whole rows coalesced (neighbouring output rows re-use them
//    through L1).  The adjoint of such an op is the same op with up<->down swapped and the
//    taps reversed (UpFirDn2dBackward, upfirdn2d.py:20-59), so one kernel serves all orders.
//  * affine_warp: out = bilinear(img, theta * [x_n, y_n, 1]) with zero padding,
//    align_corners=False -- F.affine_grid + F.grid_sample fused: the sampling grid
//    (B x H x W x 2 floats, 150 MB at B=128) and the batched [HW,3]x[3,2] GEMM that builds it
//    never exist.  The adjoint scatters with red.global.add.f32.
#include "common.cuh"

namespace dusty {

constexpr int kMaxTaps1d = 64;

struct Fir1d {
  int n_in, n_out;     // extent along the filtered axis
  int other;           // extent of the other image axis
  int k, p0, flip;
};

constexpr int kFirR = 4;      // outputs per thread (independent accumulators: loads in flight)
constexpr int kFirRows = 4;   // image rows per block of the x kernel

// ---- along x: grid = (ceil(n_out / (256 * R)), ceil(rows / kFirRows), N).  R is picked so that
// one block spans a whole row when it can (ADA's rows are 512 - 1048 outputs: with R fixed at
// 4 every second block held 24 live outputs), and a block walks kFirRows rows so that the tap
// load and the block launch are amortised -- the first version launched 20-40 k blocks of one
// short row segment each and ran at 12-30 % of HBM on these 1-channel images.
template <int U, int D, int R>
__global__ void __launch_bounds__(256)
fir1d_x_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ taps,
               Fir1d p) {
  constexpr int kOut = 256 * R;
  __shared__ float sk[kMaxTaps1d];
  __shared__ float sx[kOut * D / U + kMaxTaps1d + 4];
  if (threadIdx.x < p.k) sk[threadIdx.x] = taps[p.flip ? p.k - 1 - threadIdx.x : threadIdx.x];
  const int m0 = blockIdx.x * kOut;
  // input span needed by outputs [m0, m0 + kOut): q in [m0*D - p0, (m0+kOut-1)*D - p0 + k - 1]
  const int q_lo = m0 * D - p.p0;
  const int i_lo = (q_lo >= 0 ? q_lo : q_lo - (U - 1)) / U;          // floor(q_lo / U)
  const int span = ((kOut - 1) * D + p.k - 1) / U + 2;
  int base[R], t0[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int m = m0 + threadIdx.x + 256 * r;
    base[r] = m * D - p.p0;                   // q = base + t
    t0[r] = (U == 1) ? 0 : ((-base[r]) & (U - 1));
  }
  const int row_end = min((int)(blockIdx.y + 1) * kFirRows, p.other);
  for (int rowi = blockIdx.y * kFirRows; rowi < row_end; ++rowi) {
    const float *row = x + ((int64_t)blockIdx.z * p.other + rowi) * p.n_in;
    __syncthreads();                          // previous row's reads of sx are done (and sk is set)
    for (int j = threadIdx.x; j < span; j += 256) {
      const int i = i_lo + j;
      sx[j] = (i >= 0 && i < p.n_in) ? __ldg(row + i) : 0.f;
    }
    __syncthreads();
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    for (int t = 0; t < p.k; t += U) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int tt = t + t0[r];
        if (tt < p.k) {
          const int q = base[r] + tt;           // multiple of U
          const int i = (U == 1) ? q : (q >> 1);
          acc[r] = fmaf(sk[tt], sx[i - i_lo], acc[r]);
        }
      }
    }
    float *orow = y + ((int64_t)blockIdx.z * p.other + rowi) * p.n_out;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int m = m0 + threadIdx.x + 256 * r;
      if (m < p.n_out) orow[m] = acc[r];
    }
  }
}

// ---- along y: threads are a flat index over (row group, column): grid = (ceil(other * groups /
// 256), 1, N), so no block is a mostly-empty column remainder; thread = one column, kFirR
// consecutive output rows
template <int U, int D>
__global__ void __launch_bounds__(256)
fir1d_y_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ taps,
               Fir1d p) {
  __shared__ float sk[kMaxTaps1d];
  if (threadIdx.x < p.k) sk[threadIdx.x] = taps[p.flip ? p.k - 1 - threadIdx.x : threadIdx.x];
  __syncthreads();
  const int groups = (p.n_out + kFirR - 1) / kFirR;
  const int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (gid >= (int64_t)groups * p.other) return;
  const int grp = (int)(gid / p.other);
  const int col = (int)(gid - (int64_t)grp * p.other);
  const float *img = x + (int64_t)blockIdx.z * p.n_in * p.other + col;
  float acc[kFirR];
  int base[kFirR], t0[kFirR];
#pragma unroll
  for (int r = 0; r < kFirR; ++r) {
    const int m = grp * kFirR + r;
    base[r] = m * D - p.p0;
    t0[r] = (U == 1) ? 0 : ((-base[r]) & (U - 1));
    acc[r] = 0.f;
  }
  for (int t = 0; t < p.k; t += U) {
#pragma unroll
    for (int r = 0; r < kFirR; ++r) {
      const int tt = t + t0[r];
      const int q = base[r] + tt;
      const int i = (U == 1) ? q : (q >> 1);
      if (tt < p.k && i >= 0 && i < p.n_in) acc[r] = fmaf(sk[tt], __ldg(img + (int64_t)i * p.other), acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < kFirR; ++r) {
    const int m = grp * kFirR + r;
    if (m < p.n_out) y[((int64_t)blockIdx.z * p.n_out + m) * p.other + col] = acc[r];
  }
}

// ---- register-window variants for the shapes AdaptiveAugment runs: 12 taps (SYM6), 2x up OR 2x
// down.  A thread owns G consecutive outputs along the filtered axis and the input window they
// share, loaded straight from global memory (neighbouring windows overlap: L1 / L2 hits); every
// tap / window index is a compile-time constant, so there is no shared memory, no barrier, no
// per-tap predicate, and the grid is a flat index (fine-grained tail).  The reference's own
// upfirdn2d kernel (sm_100 build, oracle/_ref) ran these 1-channel images 1.5x faster than the
// generic kernels above -- see profiles/ for the timings of all three.
//   out[m] = sum_t tk[t] xz[m*D - p0 + t],  xz[q] = x[q / U] if q % U == 0 and in range else 0
template <int U, int D, int K, int G>
struct FirWin {
  static constexpr int NW = (U == 2) ? (G / 2 + K / 2 + 1) : (D * (G - 1) + K);
  // b = m_s*D - p0 (first q of the group's first output); returns the first input index
  __device__ static __forceinline__ int first(int b) { return U == 2 ? (b >> 1) : b; }
  __device__ static __forceinline__ void run(const float (&tk)[K], const float (&win)[NW], int b,
                                             float (&acc)[G]) {
#pragma unroll
    for (int r = 0; r < G; ++r) acc[r] = 0.f;
    if (U == 2) {
      if (b & 1) {
#pragma unroll
        for (int r = 0; r < G; ++r)
#pragma unroll
          for (int t = 0; t < K; ++t)
            if (((1 + r + t) & 1) == 0) acc[r] = fmaf(tk[t], win[(1 + r + t) >> 1], acc[r]);
      } else {
#pragma unroll
        for (int r = 0; r < G; ++r)
#pragma unroll
          for (int t = 0; t < K; ++t)
            if (((r + t) & 1) == 0) acc[r] = fmaf(tk[t], win[(r + t) >> 1], acc[r]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < G; ++r)
#pragma unroll
        for (int t = 0; t < K; ++t) acc[r] = fmaf(tk[t], win[D * r + t], acc[r]);
    }
  }
};

template <int U, int D, int K, int G>
__global__ void __launch_bounds__(256)
fir1d_x_win_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ taps,
                   Fir1d p, int groups, int total, int vec_ok) {
  using FW = FirWin<U, D, K, G>;
  const int gid = blockIdx.x * 256 + threadIdx.x;
  if (gid >= total) return;
  const int grp = gid % groups;
  const int row = gid / groups;                 // image * other + row
  float tk[K];
#pragma unroll
  for (int t = 0; t < K; ++t) tk[t] = __ldg(taps + (p.flip ? K - 1 - t : t));
  const int m_s = grp * G;
  const int b = m_s * D - p.p0;
  const int i0 = FW::first(b);
  const float *xr = x + (int64_t)row * p.n_in;
  float win[FW::NW];
#pragma unroll
  for (int j = 0; j < FW::NW; ++j) {
    const int i = i0 + j;
    win[j] = (i >= 0 && i < p.n_in) ? __ldg(xr + i) : 0.f;
  }
  float acc[G];
  FW::run(tk, win, b, acc);
  float *yr = y + (int64_t)row * p.n_out + m_s;
  if (vec_ok && m_s + G <= p.n_out) {
#pragma unroll
    for (int r = 0; r < G; r += 4)
      *reinterpret_cast<float4 *>(yr + r) = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
  } else {
#pragma unroll
    for (int r = 0; r < G; ++r)
      if (m_s + r < p.n_out) yr[r] = acc[r];
  }
}

// along y: thread = (image, group of G output rows, column), column fastest (coalesced rows)
template <int U, int D, int K, int G>
__global__ void __launch_bounds__(256)
fir1d_y_win_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ taps,
                   Fir1d p, int groups, int total) {
  using FW = FirWin<U, D, K, G>;
  const int gid = blockIdx.x * 256 + threadIdx.x;
  if (gid >= total) return;
  const int col = gid % p.other;
  const int q = gid / p.other;
  const int grp = q % groups;
  const int n = q / groups;
  float tk[K];
#pragma unroll
  for (int t = 0; t < K; ++t) tk[t] = __ldg(taps + (p.flip ? K - 1 - t : t));
  const int m_s = grp * G;
  const int b = m_s * D - p.p0;
  const int i0 = FW::first(b);
  const float *img = x + (int64_t)n * p.n_in * p.other + col;
  float win[FW::NW];
#pragma unroll
  for (int j = 0; j < FW::NW; ++j) {
    const int i = i0 + j;
    win[j] = (i >= 0 && i < p.n_in) ? __ldg(img + (int64_t)i * p.other) : 0.f;
  }
  float acc[G];
  FW::run(tk, win, b, acc);
  float *out = y + ((int64_t)n * p.n_out + m_s) * p.other + col;
#pragma unroll
  for (int r = 0; r < G; ++r)
    if (m_s + r < p.n_out) out[(int64_t)r * p.other] = acc[r];
}

// ------------------------------------------------------------------ affine warp
// theta: [N, 2, 3] row-major.  grid = (ceil(Wo/256), Ho, N*C)
constexpr int kWarpR = 4;     // output rows per thread

template <bool ADJ>
__global__ void __launch_bounds__(256)
affine_warp_kernel(const float *__restrict__ src, float *__restrict__ dst,
                   const float *__restrict__ theta, int C, int Hi, int Wi, int Ho, int Wo) {
  const int ox = blockIdx.x * 256 + threadIdx.x;
  if (ox >= Wo) return;
  const int nc = blockIdx.z, n = nc / C;
  const float *th = theta + (int64_t)n * 6;
  const float t0 = th[0], t1 = th[1], t2 = th[2], t3 = th[3], t4 = th[4], t5 = th[5];
  const int64_t in_base = (int64_t)nc * Hi * Wi;
  // base grid of affine_grid (align_corners=False): (2j + 1) / W - 1
  const float xn = (2.f * ox + 1.f) / Wo - 1.f;
  float w[kWarpR][4];
  int64_t off[kWarpR][4];
  bool ok[kWarpR][4];
#pragma unroll
  for (int r = 0; r < kWarpR; ++r) {
    const int oy = blockIdx.y * kWarpR + r;
    const float yn = (2.f * oy + 1.f) / Ho - 1.f;
    const float gx = t0 * xn + t1 * yn + t2;
    const float gy = t3 * xn + t4 * yn + t5;
    // grid_sample un-normalisation (align_corners=False)
    const float ix = ((gx + 1.f) * Wi - 1.f) * 0.5f, iy = ((gy + 1.f) * Hi - 1.f) * 0.5f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float ax = ix - fx0, ay = iy - fy0;
    const int x0 = (int)fx0, y0 = (int)fy0;
    w[r][0] = (1.f - ax) * (1.f - ay); w[r][1] = ax * (1.f - ay);
    w[r][2] = (1.f - ax) * ay;         w[r][3] = ax * ay;
    const bool vx0 = x0 >= 0 && x0 < Wi, vx1 = x0 + 1 >= 0 && x0 + 1 < Wi;
    const bool vy0 = y0 >= 0 && y0 < Hi, vy1 = y0 + 1 >= 0 && y0 + 1 < Hi;
    const bool row_ok = oy < Ho;
    ok[r][0] = row_ok && vy0 && vx0; ok[r][1] = row_ok && vy0 && vx1;
    ok[r][2] = row_ok && vy1 && vx0; ok[r][3] = row_ok && vy1 && vx1;
    off[r][0] = (int64_t)y0 * Wi + x0;       off[r][1] = off[r][0] + 1;
    off[r][2] = (int64_t)(y0 + 1) * Wi + x0; off[r][3] = off[r][2] + 1;
  }
  if (!ADJ) {
    const float *im = src + in_base;
    float v[kWarpR][4];
#pragma unroll
    for (int r = 0; r < kWarpR; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[r][j] = ok[r][j] ? __ldg(im + off[r][j]) : 0.f;
#pragma unroll
    for (int r = 0; r < kWarpR; ++r) {
      const int oy = blockIdx.y * kWarpR + r;
      if (oy < Ho) {
        float o = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o = fmaf(w[r][j], v[r][j], o);
        dst[((int64_t)nc * Ho + oy) * Wo + ox] = o;
      }
    }
  } else {
    float *gi = dst + in_base;                 // zero-filled by the caller
#pragma unroll
    for (int r = 0; r < kWarpR; ++r) {
      const int oy = blockIdx.y * kWarpR + r;
      if (oy >= Ho) continue;
      const float g = src[((int64_t)nc * Ho + oy) * Wo + ox];   // gradient w.r.t. the warped image
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ok[r][j]) atomicAdd(gi + off[r][j], w[r][j] * g);
    }
  }
}

}  // namespace dusty

using namespace dusty;

extern "C" int dusty_fir1d(const float *x, float *y, const float *taps, int k, int flip, int64_t N,
                           int in_h, int in_w, int axis, int up, int down, int pad0, int pad1,
                           void *stream) {
  DUSTY_CHECK_ARG(x && y && taps, "null pointer");
  DUSTY_CHECK_ARG(k >= 1 && k <= kMaxTaps1d, "tap count must be in [1, 64]");
  DUSTY_CHECK_ARG((up == 1 || up == 2) && (down == 1 || down == 2), "up/down must be 1 or 2");
  DUSTY_CHECK_ARG(axis == 0 || axis == 1, "axis must be 0 (y) or 1 (x)");
  DUSTY_CHECK_ARG(N >= 1 && N <= 65535 && in_h >= 1 && in_w >= 1, "bad shape");
  const int n_in = axis ? in_w : in_h, other = axis ? in_h : in_w;
  const int n_out = (n_in * up + pad0 + pad1 - k + down) / down;
  DUSTY_CHECK_ARG(n_out >= 1 && n_out <= (axis ? 0x7fffffff : 65535) && other <= (axis ? 65535 : 0x7fffffff),
                  "bad output size");
  Fir1d p{n_in, n_out, other, k, pad0, flip ? 1 : 0};
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_X(R)                                                                                   \
  do {                                                                                                  \
    dim3 grid((unsigned)((n_out + 256 * R - 1) / (256 * R)), (unsigned)((other + kFirRows - 1) / kFirRows), \
              (unsigned)N);                                                                             \
    if (up == 1 && down == 1) fir1d_x_kernel<1, 1, R><<<grid, 256, 0, st>>>(x, y, taps, p);             \
    else if (up == 2 && down == 1) fir1d_x_kernel<2, 1, R><<<grid, 256, 0, st>>>(x, y, taps, p);        \
    else if (up == 1 && down == 2) fir1d_x_kernel<1, 2, R><<<grid, 256, 0, st>>>(x, y, taps, p);        \
    else fir1d_x_kernel<2, 2, R><<<grid, 256, 0, st>>>(x, y, taps, p);                                  \
  } while (0)
  // AdaptiveAugment's passes (12-tap SYM6, 2x up or 2x down): register-window kernels
  constexpr int kG = 8;
  const int64_t groups_w = (n_out + kG - 1) / kG;
  const int64_t total_w = (int64_t)N * other * groups_w;
  // (not for 2x decimation along x: there a thread's window starts 16 floats after its
  // neighbour's, every load instruction touches 32 different sectors, and the row-staging kernel
  // below measured faster: 35 vs 40 us at [64,1,140,1036])
  if (k == 12 && ((up == 2 && down == 1) || (up == 1 && down == 2 && axis == 0)) && total_w < 0x7fffffff &&
      (int64_t)N * other < 0x7fffffff) {
    const unsigned blocks = (unsigned)((total_w + 255) / 256);
    if (axis == 1) {
      const int vec_ok = (n_out % 4 == 0) && aligned16(y);
      if (up == 2)
        fir1d_x_win_kernel<2, 1, 12, kG><<<blocks, 256, 0, st>>>(x, y, taps, p, (int)groups_w, (int)total_w, vec_ok);
      else
        fir1d_x_win_kernel<1, 2, 12, kG><<<blocks, 256, 0, st>>>(x, y, taps, p, (int)groups_w, (int)total_w, vec_ok);
    } else {
      if (up == 2)
        fir1d_y_win_kernel<2, 1, 12, kG><<<blocks, 256, 0, st>>>(x, y, taps, p, (int)groups_w, (int)total_w);
      else
        fir1d_y_win_kernel<1, 2, 12, kG><<<blocks, 256, 0, st>>>(x, y, taps, p, (int)groups_w, (int)total_w);
    }
    DUSTY_LAUNCH_CHECK();
    return DUSTY_OK;
  }
  if (axis == 1) {
    // outputs per thread: the smallest of {2, 3, 4, 5} that lets one block span the row, else 4
    if (n_out <= 512) LAUNCH_X(2);
    else if (n_out <= 768) LAUNCH_X(3);
    else if (n_out <= 1024 || n_out > 1280) LAUNCH_X(4);
    else LAUNCH_X(5);
  } else {
    const int64_t groups = (n_out + kFirR - 1) / kFirR;
    dim3 grid((unsigned)((groups * other + 255) / 256), 1u, (unsigned)N);
    if (up == 1 && down == 1) fir1d_y_kernel<1, 1><<<grid, 256, 0, st>>>(x, y, taps, p);
    else if (up == 2 && down == 1) fir1d_y_kernel<2, 1><<<grid, 256, 0, st>>>(x, y, taps, p);
    else if (up == 1 && down == 2) fir1d_y_kernel<1, 2><<<grid, 256, 0, st>>>(x, y, taps, p);
    else fir1d_y_kernel<2, 2><<<grid, 256, 0, st>>>(x, y, taps, p);
  }
#undef LAUNCH_X
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_affine_warp(const float *src, float *dst, const float *theta, int N, int C,
                                 int Hi, int Wi, int Ho, int Wo, int adjoint, void *stream) {
  DUSTY_CHECK_ARG(src && dst && theta, "null pointer");
  DUSTY_CHECK_ARG(N >= 1 && C >= 1 && (int64_t)N * C <= 65535 && Hi >= 1 && Wi >= 1 && Ho >= 1 &&
                      Ho <= 65535 && Wo >= 1, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((Wo + 255) / 256), (unsigned)((Ho + kWarpR - 1) / kWarpR), (unsigned)(N * C));
  if (!adjoint) {
    affine_warp_kernel<false><<<grid, 256, 0, st>>>(src, dst, theta, C, Hi, Wi, Ho, Wo);
  } else {
    if (cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)N * C * Hi * Wi, st) != cudaSuccess)
      return DUSTY_ECUDA;
    affine_warp_kernel<true><<<grid, 256, 0, st>>>(src, dst, theta, C, Hi, Wi, Ho, Wo);
  }
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

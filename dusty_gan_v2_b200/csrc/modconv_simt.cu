// a1: modulated 1x1 convolution -- SIMT (CUDA-core, exact fp32 FMA) implementation.
// This is the fp32 parity path (rtol 1e-3 against the oracle needs true fp32 products,
// SURVEY 7.3-3); the bf16 production path is the tcgen05 kernel in modconv_tc.cu.
//
// Data layout: activations NCHW => per sample X_b is [K, P] with the pixel index
// contiguous; per-sample effective weights wb_b are [O, K] row-major.  The channel axis of
// X is the concatenation of two sources (upsampled features x1, Fourier features x2) that
// are never concatenated in memory; x2 may be shared by the whole batch (B2 == 1).
#include "common.cuh"

namespace dusty {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T> struct Ld4 {};
template <> struct Ld4<float> {
  static __device__ __forceinline__ void ld(const float *p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct Ld4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16 *p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2 *>(p);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
};

struct GemmNN {
  // A(m,k) = a[b*a_bs + m*a_ms + k*a_ks]
  const void *a; int64_t a_bs, a_ms, a_ks;
  // B(k,n): k < K1 from b1 (batch stride b1_bs), else from b2 (batch stride b2_bs, 0 if shared)
  const void *b1; const void *b2; int64_t b1_bs, b2_bs; int K1;
  void *c; int64_t c_bs;
  const float *bias;
  int M, K; int64_t N;
  int act; float alpha, scale;
  // optional per-row EMA normalisers (heads: one ModConv2d per output row, each with its own
  // ema_var): row r of the weights is multiplied by 1 / (sqrt(*ema_rows[r]) + 1e-8) as it is
  // staged in shared memory (small-O kernels only)
  const float *ema_rows[4];
  __device__ __forceinline__ float row_scale(int r) const {
    return (r < 4 && ema_rows[r]) ? 1.f / (sqrtf(__ldg(ema_rows[r])) + 1e-8f) : 1.f;
  }
};

// C[b] (M x N) = A[b] (M x K) * B[b] (K x N), fp32 accumulate, fused bias + lrelu epilogue.
template <typename T, typename TA, bool A_KCONTIG>
__global__ void __launch_bounds__(256)
gemm_nn_kernel(GemmNN g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * BM;
  const int64_t n0 = (int64_t)blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const TA *A = (const TA *)g.a + (int64_t)b * g.a_bs;
  const T *B1 = (const T *)g.b1 + (int64_t)b * g.b1_bs;
  const T *B2 = (const T *)g.b2 + (int64_t)b * g.b2_bs;
  const bool n_vec = (g.N % 4 == 0);
  float acc[4][4] = {};

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (A_KCONTIG) {
      const int m = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gm = m0 + m, gk = k0 + kq + i;
        As[kq + i][m] = (gm < g.M && gk < g.K) ? to_f(A[(int64_t)gm * g.a_ms + gk]) : 0.f;
      }
    } else {
      const int k = tid >> 4, mq = (tid & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gm = m0 + mq + i, gk = k0 + k;
        As[k][mq + i] = (gm < g.M && gk < g.K) ? to_f(A[(int64_t)gm * g.a_ms + (int64_t)gk * g.a_ks]) : 0.f;
      }
    }
    // ---- B tile -> Bs[k][n]
    {
      const int k = tid >> 4, nq = (tid & 15) * 4;
      const int gk = k0 + k;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gk < g.K) {
        const T *src = (gk < g.K1) ? (B1 + (int64_t)gk * g.N) : (B2 + (int64_t)(gk - g.K1) * g.N);
        const int64_t gn = n0 + nq;
        if (n_vec && gn + 4 <= g.N) {
          Ld4<T>::ld(src + gn, v);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (gn + i < g.N) v[i] = to_f(src[gn + i]);
        }
      }
      *reinterpret_cast<float4 *>(&Bs[k][nq]) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
  T *C = (T *)g.c + (int64_t)b * g.c_bs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
    const float bv = g.bias ? g.bias[gm] : 0.f;
    const int64_t gn = n0 + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (gn + j >= g.N) continue;
      float v = acc[i][j] + bv;
      if (g.act == 3) v = (v > 0.f) ? v : v * g.alpha;
      C[(int64_t)gm * g.N + gn + j] = from_f<T>(v * g.scale);
    }
  }
}

// Heads: O <= 4 outputs.  Pure streaming read of x (memory bound); weights in smem.  Block =
// PG pixel groups (16 bytes of pixels each) x KS slices of the channel axis: at the coarse
// levels there are few pixels and many channels, and one thread walking all of K is pure load
// latency.  Partial sums meet in shared memory; slice 0 applies bias / activation and stores.
template <typename T, typename TA, int OMAX>
__global__ void __launch_bounds__(256)
small_o_kernel(GemmNN g, int PG) {
  extern __shared__ float sw[];  // [M][K] weights, then [256][OMAX * V] partials
  constexpr int V = Vec16<T>::N;
  float *part = sw + g.M * g.K;
  const int b = blockIdx.y;
  const int KS = blockDim.x / PG;
  const int tx = threadIdx.x % PG, ty = threadIdx.x / PG;
  const TA *A = (const TA *)g.a + (int64_t)b * g.a_bs;
  for (int i = threadIdx.x; i < g.M * g.K; i += blockDim.x) {
    const int m = i / g.K, k = i - m * g.K;
    sw[i] = to_f(A[(int64_t)m * g.a_ms + (int64_t)k * g.a_ks]) * g.row_scale(m);
  }
  __syncthreads();
  const int64_t p0 = ((int64_t)blockIdx.x * PG + tx) * V;
  const bool active = p0 < g.N;
  const T *B1 = (const T *)g.b1 + (int64_t)b * g.b1_bs;
  const T *B2 = (const T *)g.b2 + (int64_t)b * g.b2_bs;
  const bool full = active && (g.N % V == 0) && (p0 + V <= g.N) &&
                    ((reinterpret_cast<uintptr_t>(B1) | reinterpret_cast<uintptr_t>(B2)) & 15u) == 0;
  float acc[OMAX][V] = {};
  auto src_of = [&](int k) {
    return ((k < g.K1) ? (B1 + (int64_t)k * g.N) : (B2 + (int64_t)(k - g.K1) * g.N)) + p0;
  };
  auto fma_k = [&](int k, const float *v) {
#pragma unroll
    for (int m = 0; m < OMAX; ++m) {
      if (m < g.M) {
        const float w = sw[m * g.K + k];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[m][i] = fmaf(w, v[i], acc[m][i]);
      }
    }
  };
  if (active) {
    int k = ty;
    if (full) {
      for (; k + 3 * KS < g.K; k += 4 * KS) {
        Vec16<T> v0 = ld16_stream(src_of(k)), v1 = ld16_stream(src_of(k + KS));
        Vec16<T> v2 = ld16_stream(src_of(k + 2 * KS)), v3 = ld16_stream(src_of(k + 3 * KS));
        float f[V];
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] = v0.get(i);
        fma_k(k, f);
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] = v1.get(i);
        fma_k(k + KS, f);
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] = v2.get(i);
        fma_k(k + 2 * KS, f);
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] = v3.get(i);
        fma_k(k + 3 * KS, f);
      }
    }
    for (; k < g.K; k += KS) {
      const T *src = src_of(k);
      float f[V];
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = (p0 + i < g.N) ? to_f(src[i]) : 0.f;
      fma_k(k, f);
    }
  }
  if (KS > 1) {
    if (ty > 0) {
#pragma unroll
      for (int m = 0; m < OMAX; ++m)
#pragma unroll
        for (int i = 0; i < V; ++i) part[(m * V + i) * 256 + threadIdx.x] = acc[m][i];
    }
    __syncthreads();
    if (ty > 0) return;
    for (int s2 = 1; s2 < KS; ++s2) {
#pragma unroll
      for (int m = 0; m < OMAX; ++m)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[m][i] += part[(m * V + i) * 256 + s2 * PG + tx];
    }
  }
  if (!active) return;
  T *C = (T *)g.c + (int64_t)b * g.c_bs;
#pragma unroll
  for (int m = 0; m < OMAX; ++m) {
    if (m >= g.M) continue;
    const float bv = g.bias ? g.bias[m] : 0.f;
    float o[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float v = acc[m][i] + bv;
      if (g.act == 3) v = (v > 0.f) ? v : v * g.alpha;
      o[i] = v * g.scale;
    }
    T *dst = C + (int64_t)m * g.N + p0;
    if (full && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
      Vec16<T> ov;
#pragma unroll
      for (int i = 0; i < V; ++i) ov.set(i, o[i]);
      st16(dst, ov);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) if (p0 + i < g.N) dst[i] = from_f<T>(o[i]);
    }
  }
}

// dX of the heads: dx[b, c, p] = sum_{o < O} wb[b, o, c] * dy[b, o, p], O <= 4.  Streaming
// write of C x P per sample; each thread keeps its 16-byte pixel group of dY in registers and
// walks a chunk of the channel axis.
template <typename T, typename TA, int OMAX>
__global__ void __launch_bounds__(256)
small_o_dx_kernel(GemmNN g, int c_chunk) {
  extern __shared__ float sw[];  // [O][c_chunk]
  constexpr int V = Vec16<T>::N;
  const int b = blockIdx.y;
  const int c0 = blockIdx.z * c_chunk;
  const int nc = min(c_chunk, g.M - c0);
  const TA *A = (const TA *)g.a + (int64_t)b * g.a_bs;
  for (int i = threadIdx.x; i < g.K * nc; i += blockDim.x) {
    const int o = i / nc, c = i - o * nc;
    sw[o * c_chunk + c] = to_f(A[(int64_t)(c0 + c) * g.a_ms + (int64_t)o * g.a_ks]) * g.row_scale(o);
  }
  __syncthreads();
  const int64_t p0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (p0 >= g.N) return;
  const T *DY = (const T *)g.b1 + (int64_t)b * g.b1_bs;
  T *C = (T *)g.c + (int64_t)b * g.c_bs;
  const bool full = (g.N % V == 0) && (p0 + V <= g.N) &&
                    ((reinterpret_cast<uintptr_t>(DY) | reinterpret_cast<uintptr_t>(C)) & 15u) == 0;
  float gv[OMAX][V];
#pragma unroll
  for (int o = 0; o < OMAX; ++o) {
    if (o < g.K) {
      if (full) {
        Vec16<T> v = ld16_stream(DY + (int64_t)o * g.N + p0);
#pragma unroll
        for (int i = 0; i < V; ++i) gv[o][i] = v.get(i);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) gv[o][i] = (p0 + i < g.N) ? to_f(DY[(int64_t)o * g.N + p0 + i]) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) gv[o][i] = 0.f;
    }
  }
  for (int c = 0; c < nc; ++c) {
    float o8[V];
#pragma unroll
    for (int i = 0; i < V; ++i) o8[i] = 0.f;
#pragma unroll
    for (int o = 0; o < OMAX; ++o) {
      if (o < g.K) {
        const float w = sw[o * c_chunk + c];
#pragma unroll
        for (int i = 0; i < V; ++i) o8[i] = fmaf(w, gv[o][i], o8[i]);
      }
    }
    T *dst = C + (int64_t)(c0 + c) * g.N + p0;
    if (full) {
      Vec16<T> ov;
#pragma unroll
      for (int i = 0; i < V; ++i) ov.set(i, o8[i]);
      st16(dst, ov);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) if (p0 + i < g.N) dst[i] = from_f<T>(o8[i]);
    }
  }
}

struct GemmNT {
  const void *dy;  // [B, O, P]
  const void *x1; const void *x2; int64_t x1_bs, x2_bs; int K1;
  float *dw;       // [B, O, K]
  int O, K; int64_t P;
};

// dw[b] (O x K) = dY[b] (O x P) * X[b]^T (P x K): both operands are pixel-contiguous.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_nt_kernel(GemmNT g) {
  __shared__ __align__(16) float As[BK][BM + 4];  // [p][o]
  __shared__ __align__(16) float Bs[BK][BN + 4];  // [p][k]
  const int b = blockIdx.z;
  const int o0 = blockIdx.y * BM, k0 = blockIdx.x * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const T *DY = (const T *)g.dy + (int64_t)b * g.O * g.P;
  const T *X1 = (const T *)g.x1 + (int64_t)b * g.x1_bs;
  const T *X2 = (const T *)g.x2 + (int64_t)b * g.x2_bs;
  const bool p_vec = (g.P % 4 == 0);
  float acc[4][4] = {};
  const int r = tid >> 2, pq = (tid & 3) * 4;  // row within the tile, 4 consecutive pixels
  for (int64_t p0 = 0; p0 < g.P; p0 += BK) {
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int go = o0 + r;
      if (go < g.O) {
        const T *src = DY + (int64_t)go * g.P + p0 + pq;
        if (p_vec && p0 + pq + 4 <= g.P) Ld4<T>::ld(src, v);
        else
          for (int i = 0; i < 4; ++i) if (p0 + pq + i < g.P) v[i] = to_f(src[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[pq + i][r] = v[i];
    }
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int gk = k0 + r;
      if (gk < g.K) {
        const T *src = ((gk < g.K1) ? (X1 + (int64_t)gk * g.P) : (X2 + (int64_t)(gk - g.K1) * g.P)) + p0 + pq;
        if (p_vec && p0 + pq + 4 <= g.P) Ld4<T>::ld(src, v);
        else
          for (int i = 0; i < 4; ++i) if (p0 + pq + i < g.P) v[i] = to_f(src[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[pq + i][r] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < BK; ++p) {
      const float4 av = *reinterpret_cast<const float4 *>(&As[p][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4 *>(&Bs[p][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float *DW = g.dw + (int64_t)b * g.O * g.K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int go = o0 + ty * 4 + i;
    if (go >= g.O) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gk = k0 + tx * 4 + j;
      if (gk < g.K) DW[(int64_t)go * g.K + gk] = acc[i][j];
    }
  }
}

// dW for the heads (O <= 4): a streaming reduction over pixels.  Each CTA owns KB input
// channels of one sample; every thread strides over the pixel axis with 4-wide loads, the
// O x KB partial sums are reduced by warp shuffles + one smem pass.  x is read once.
constexpr int kSmallKB = 8;
template <typename T, int OMAX>
__global__ void __launch_bounds__(256)
small_o_dw_kernel(GemmNT g) {
  __shared__ float red[8][kSmallKB * OMAX];
  const int b = blockIdx.y;
  const int k0 = blockIdx.x * kSmallKB;
  const T *DY = (const T *)g.dy + (int64_t)b * g.O * g.P;
  const T *X1 = (const T *)g.x1 + (int64_t)b * g.x1_bs;
  const T *X2 = (const T *)g.x2 + (int64_t)b * g.x2_bs;
  float acc[kSmallKB][OMAX] = {};
  const bool p_vec = (g.P % 4 == 0);
  // grid.z slices of the pixel axis (multiples of 1024 pixels); partial sums meet in dw by
  // atomics when there is more than one slice (the host zero-fills dw then)
  const int64_t p_per = (((g.P + gridDim.z - 1) / gridDim.z) + 1023) / 1024 * 1024;
  const int64_t p_lo = (int64_t)blockIdx.z * p_per;
  const int64_t p_hi = p_lo + p_per < g.P ? p_lo + p_per : g.P;
  for (int64_t p = p_lo + (int64_t)threadIdx.x * 4; p < p_hi; p += 256 * 4) {
    float gv[OMAX][4];
#pragma unroll
    for (int o = 0; o < OMAX; ++o) {
      if (o < g.O) {
        if (p_vec) Ld4<T>::ld(DY + (int64_t)o * g.P + p, gv[o]);
        else for (int i = 0; i < 4; ++i) gv[o][i] = (p + i < g.P) ? to_f(DY[(int64_t)o * g.P + p + i]) : 0.f;
      } else {
        gv[o][0] = gv[o][1] = gv[o][2] = gv[o][3] = 0.f;
      }
    }
#pragma unroll
    for (int kk = 0; kk < kSmallKB; ++kk) {
      const int k = k0 + kk;
      if (k >= g.K) break;
      const T *src = ((k < g.K1) ? (X1 + (int64_t)k * g.P) : (X2 + (int64_t)(k - g.K1) * g.P)) + p;
      float xv[4] = {0.f, 0.f, 0.f, 0.f};
      if (p_vec) Ld4<T>::ld(src, xv);
      else for (int i = 0; i < 4; ++i) if (p + i < g.P) xv[i] = to_f(src[i]);
#pragma unroll
      for (int o = 0; o < OMAX; ++o)
        acc[kk][o] += xv[0] * gv[o][0] + xv[1] * gv[o][1] + xv[2] * gv[o][2] + xv[3] * gv[o][3];
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int kk = 0; kk < kSmallKB; ++kk)
#pragma unroll
    for (int o = 0; o < OMAX; ++o) {
      const float v = warp_sum(acc[kk][o]);
      if (lane == 0) red[wid][kk * OMAX + o] = v;
    }
  __syncthreads();
  if (threadIdx.x < kSmallKB * OMAX) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    const int kk = threadIdx.x / OMAX, o = threadIdx.x % OMAX;
    if (k0 + kk < g.K && o < g.O) {
      float *dst = g.dw + ((int64_t)b * g.O + o) * g.K + k0 + kk;
      if (gridDim.z > 1) atomicAdd(dst, v);
      else *dst = v;
    }
  }
}

// bf16 specialisation for the heads (O <= 2): 16-byte loads (8 pixels per thread per channel)
// with the next sweep's 10 vectors requested before the current ones are reduced.  The generic
// kernel above is latency-bound there (ncu: 17 % issue slots, 19 warps stalled on the long
// scoreboard per issue, 12 % of DRAM bandwidth): 80 bytes in flight per thread, consumed at once.
template <int ON>
__global__ void __launch_bounds__(256)
heads_dw_bf16_kernel(GemmNT g) {
  __shared__ float red[8][kSmallKB * ON];
  const int b = blockIdx.y;
  const int k0 = blockIdx.x * kSmallKB;
  const __nv_bfloat16 *DY = (const __nv_bfloat16 *)g.dy + (int64_t)b * g.O * g.P;
  const __nv_bfloat16 *xrow[kSmallKB];
  bool kval[kSmallKB];
#pragma unroll
  for (int kk = 0; kk < kSmallKB; ++kk) {
    const int k = k0 + kk;
    kval[kk] = k < g.K;
    const int kc = kval[kk] ? k : 0;
    xrow[kk] = (kc < g.K1) ? (const __nv_bfloat16 *)g.x1 + (int64_t)b * g.x1_bs + (int64_t)kc * g.P
                           : (const __nv_bfloat16 *)g.x2 + (int64_t)b * g.x2_bs + (int64_t)(kc - g.K1) * g.P;
  }
  const int64_t p_per = (((g.P + gridDim.z - 1) / gridDim.z) + 2047) / 2048 * 2048;
  const int64_t p_lo = (int64_t)blockIdx.z * p_per;
  const int64_t p_hi = p_lo + p_per < g.P ? p_lo + p_per : g.P;
  float acc[kSmallKB][ON] = {};
  uint4 gn[ON], xn[kSmallKB];
  auto fetch = [&](int64_t p) {
#pragma unroll
    for (int o = 0; o < ON; ++o)
      gn[o] = o < g.O ? __ldcs(reinterpret_cast<const uint4 *>(DY + (int64_t)o * g.P + p)) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int kk = 0; kk < kSmallKB; ++kk) xn[kk] = __ldcs(reinterpret_cast<const uint4 *>(xrow[kk] + p));
  };
  int64_t p = p_lo + (int64_t)threadIdx.x * 8;
  if (p < p_hi) fetch(p);
  for (; p < p_hi; p += 256 * 8) {
    uint4 gc[ON], xc[kSmallKB];
#pragma unroll
    for (int o = 0; o < ON; ++o) gc[o] = gn[o];
#pragma unroll
    for (int kk = 0; kk < kSmallKB; ++kk) xc[kk] = xn[kk];
    if (p + 256 * 8 < p_hi) fetch(p + 256 * 8);
    float gv[ON][8];
#pragma unroll
    for (int o = 0; o < ON; ++o) {
      const uint32_t w[4] = {gc[o].x, gc[o].y, gc[o].z, gc[o].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        gv[o][2 * i] = __uint_as_float(w[i] << 16);
        gv[o][2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int kk = 0; kk < kSmallKB; ++kk) {
      const uint32_t w[4] = {xc[kk].x, xc[kk].y, xc[kk].z, xc[kk].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x0 = __uint_as_float(w[i] << 16), x1 = __uint_as_float(w[i] & 0xffff0000u);
#pragma unroll
        for (int o = 0; o < ON; ++o) acc[kk][o] = fmaf(x1, gv[o][2 * i + 1], fmaf(x0, gv[o][2 * i], acc[kk][o]));
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int kk = 0; kk < kSmallKB; ++kk)
#pragma unroll
    for (int o = 0; o < ON; ++o) {
      const float v = warp_sum(acc[kk][o]);
      if (lane == 0) red[wid][kk * ON + o] = v;
    }
  __syncthreads();
  if (threadIdx.x < kSmallKB * ON) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    const int kk = threadIdx.x / ON, o = threadIdx.x % ON;
    if (kval[0] && k0 + kk < g.K && o < g.O) {
      float *dst = g.dw + ((int64_t)b * g.O + o) * g.K + k0 + kk;
      if (gridDim.z > 1) atomicAdd(dst, v);
      else *dst = v;
    }
  }
}

template <typename T, typename TA>
static int run_nn(const GemmNN &g, int B, bool a_kcontig, cudaStream_t st) {
  if (g.M <= 4 && (size_t)g.M * g.K * sizeof(float) <= 48 * 1024) {
    constexpr int V = Vec16<T>::N;
    // pixel groups per block: 32 (8 channel slices) unless the channel axis is short
    const int64_t groups = (g.N + V - 1) / V;
    // one slice (PG = 256) when the pixel axis alone fills the machine; otherwise halve PG
    // (double the channel slices) while a slice keeps >= 8 channels
    int PG = 256;
    while (PG > 8 && ((groups + PG - 1) / PG) * B < 4 * (int64_t)num_sms() &&
           g.K / (256 / (PG / 2)) >= 8)
      PG >>= 1;
    dim3 grid((unsigned)((groups + PG - 1) / PG), (unsigned)B);
    const size_t smem = ((size_t)g.M * g.K + 256 * 4 * V) * sizeof(float);
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(small_o_kernel<T, TA, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           96 * 1024);
      configured = true;
    }
    small_o_kernel<T, TA, 4><<<grid, 256, smem, st>>>(g, PG);
    return 0;
  }
  dim3 grid((unsigned)((g.N + BN - 1) / BN), (unsigned)((g.M + BM - 1) / BM), (unsigned)B);
  if (a_kcontig) gemm_nn_kernel<T, TA, true><<<grid, 256, 0, st>>>(g);
  else gemm_nn_kernel<T, TA, false><<<grid, 256, 0, st>>>(g);
  return 0;
}

static bool set_rows(GemmNN &g, const float *const *ema_rows, int rows, bool small_o) {
  for (int r = 0; r < 4; ++r) g.ema_rows[r] = nullptr;
  if (!ema_rows) return true;
  if (!small_o || rows > 4) return false;
  for (int r = 0; r < rows; ++r) g.ema_rows[r] = ema_rows[r];
  return true;
}

int modconv_fwd_simt(const void *wb, const void *x1, const void *x2, const float *bias, void *y,
                     int B, int O, int C1, int C2, int B2, int64_t P, int act, float alpha,
                     float scale, int dtype, int wdtype, cudaStream_t st, const float *const *ema_rows) {
  GemmNN g;
  const int K = C1 + C2;
  if (!set_rows(g, ema_rows, O, O <= 4 && (size_t)O * K * sizeof(float) <= 48 * 1024)) {
    set_error("modconv_fwd_simt: ema_rows needs the small-O kernel (O <= 4)");
    return DUSTY_EUNSUPPORTED;
  }
  g.a = wb; g.a_bs = (int64_t)O * K; g.a_ms = K; g.a_ks = 1;
  g.b1 = x1; g.b2 = x2; g.b1_bs = (int64_t)C1 * P; g.b2_bs = (B2 == 1) ? 0 : (int64_t)C2 * P;
  g.K1 = C1; g.c = y; g.c_bs = (int64_t)O * P; g.bias = bias;
  g.M = O; g.K = K; g.N = P; g.act = act; g.alpha = alpha; g.scale = scale;
  if (dtype == DUSTY_F32) {
    if (wdtype == DUSTY_F32) return run_nn<float, float>(g, B, true, st);
    return run_nn<float, __nv_bfloat16>(g, B, true, st);
  }
  if (wdtype == DUSTY_F32) return run_nn<__nv_bfloat16, float>(g, B, true, st);
  return run_nn<__nv_bfloat16, __nv_bfloat16>(g, B, true, st);
}

int modconv_bwd_dx_simt(const void *wb, const void *dy, void *dx1, int B, int O, int C1, int K,
                        int64_t P, int dtype, int wdtype, cudaStream_t st, const float *const *ema_rows) {
  GemmNN g;
  if (!set_rows(g, ema_rows, O, O <= 4)) {
    set_error("modconv_bwd_dx_simt: ema_rows needs the small-O kernel (O <= 4)");
    return DUSTY_EUNSUPPORTED;
  }
  // A(m = k, kk = o) = wb[b, o, k]
  g.a = wb; g.a_bs = (int64_t)O * K; g.a_ms = 1; g.a_ks = K;
  g.b1 = dy; g.b2 = dy; g.b1_bs = (int64_t)O * P; g.b2_bs = 0; g.K1 = O;
  g.c = dx1; g.c_bs = (int64_t)C1 * P; g.bias = nullptr;
  g.M = C1; g.K = O; g.N = P; g.act = 1; g.alpha = 0.f; g.scale = 1.f;
  if (O <= 4) {
    const int V = dtype == DUSTY_F32 ? 4 : 8;
    const int64_t groups = (P + V - 1) / V;
    // enough CTAs to fill the machine: split the channel axis when there are few pixels
    int c_chunk = C1;
    while (c_chunk > 16 && ((groups + 255) / 256) * B * ((C1 + c_chunk - 1) / c_chunk) < 2 * num_sms())
      c_chunk = (c_chunk + 1) / 2;
    dim3 grid((unsigned)((groups + 255) / 256), (unsigned)B, (unsigned)((C1 + c_chunk - 1) / c_chunk));
    const size_t smem = (size_t)O * c_chunk * sizeof(float);
    if (dtype == DUSTY_F32) {
      if (wdtype == DUSTY_F32) small_o_dx_kernel<float, float, 4><<<grid, 256, smem, st>>>(g, c_chunk);
      else small_o_dx_kernel<float, __nv_bfloat16, 4><<<grid, 256, smem, st>>>(g, c_chunk);
    } else {
      if (wdtype == DUSTY_F32) small_o_dx_kernel<__nv_bfloat16, float, 4><<<grid, 256, smem, st>>>(g, c_chunk);
      else small_o_dx_kernel<__nv_bfloat16, __nv_bfloat16, 4><<<grid, 256, smem, st>>>(g, c_chunk);
    }
    return 0;
  }
  dim3 grid((unsigned)((g.N + BN - 1) / BN), (unsigned)((g.M + BM - 1) / BM), (unsigned)B);
  if (dtype == DUSTY_F32) {
    if (wdtype == DUSTY_F32) gemm_nn_kernel<float, float, false><<<grid, 256, 0, st>>>(g);
    else gemm_nn_kernel<float, __nv_bfloat16, false><<<grid, 256, 0, st>>>(g);
  } else {
    if (wdtype == DUSTY_F32) gemm_nn_kernel<__nv_bfloat16, float, false><<<grid, 256, 0, st>>>(g);
    else gemm_nn_kernel<__nv_bfloat16, __nv_bfloat16, false><<<grid, 256, 0, st>>>(g);
  }
  return 0;
}

int modconv_bwd_dw_simt(const void *dy, const void *x1, const void *x2, float *dwb, int B, int O,
                        int C1, int C2, int B2, int64_t P, int dtype, cudaStream_t st) {
  GemmNT g;
  g.dy = dy; g.x1 = x1; g.x2 = x2; g.x1_bs = (int64_t)C1 * P;
  g.x2_bs = (B2 == 1) ? 0 : (int64_t)C2 * P; g.K1 = C1; g.dw = dwb;
  g.O = O; g.K = C1 + C2; g.P = P;
  if (O <= 4) {
    const int kt = (g.K + kSmallKB - 1) / kSmallKB;
    int ps = 1;
    while (ps < 32 && (int64_t)kt * B * ps < 4 * num_sms() && P / (ps * 2) >= 2048) ps *= 2;
    if (ps > 1 &&
        cudaMemsetAsync(dwb, 0, sizeof(float) * (size_t)B * O * g.K, st) != cudaSuccess)
      return DUSTY_ECUDA;
    dim3 grid((unsigned)kt, (unsigned)B, (unsigned)ps);
    if (g.O <= 2) {      // the two 1-channel heads: half the accumulators, twice the resident warps
      const bool vec8 = dtype == DUSTY_BF16 && P % 8 == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(x1) & 15) == 0 && (reinterpret_cast<uintptr_t>(x2) & 15) == 0;
      if (vec8) heads_dw_bf16_kernel<2><<<grid, 256, 0, st>>>(g);
      else if (dtype == DUSTY_F32) small_o_dw_kernel<float, 2><<<grid, 256, 0, st>>>(g);
      else small_o_dw_kernel<__nv_bfloat16, 2><<<grid, 256, 0, st>>>(g);
      return 0;
    }
    if (dtype == DUSTY_F32) small_o_dw_kernel<float, 4><<<grid, 256, 0, st>>>(g);
    else small_o_dw_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>(g);
    return 0;
  }
  dim3 grid((unsigned)((g.K + BN - 1) / BN), (unsigned)((g.O + BM - 1) / BM), (unsigned)B);
  if (dtype == DUSTY_F32) gemm_nt_kernel<float><<<grid, 256, 0, st>>>(g);
  else gemm_nt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(g);
  return 0;
}

}  // namespace dusty

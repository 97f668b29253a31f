// a3: fused bias + leaky-ReLU, forward / backward (+ bias-gradient reduction).
// HBM-bound: 16-byte vector accesses, one channel lookup per vector, no smem needed for
// the elementwise part; the bias gradient is reduced warp-shuffle -> smem -> one atomic
// per CTA and channel.
#include "common.cuh"

namespace dusty {

template <int ACT, int GRAD>
__device__ __forceinline__ float act_apply(float x, float ref, float alpha, float scale) {
  if (GRAD == 2) return 0.f;
  if (ACT == 3) {
    const float gate = (GRAD == 0) ? x : ref;
    return ((gate > 0.f) ? x : x * alpha) * scale;
  }
  return x * scale;
}

// Vector kernel: inner % VEC == 0 so a vector never straddles two channels.
template <typename T, int ACT, int GRAD>
__global__ void __launch_bounds__(256)
bias_act_vec_kernel(const T *__restrict__ x, const T *__restrict__ bias, const T *__restrict__ ref,
                    T *__restrict__ y, int64_t n_vec, int C, int64_t inner_vec, float alpha,
                    float scale) {
  constexpr int V = Vec16<T>::N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    Vec16<T> vx = ld16_stream(x + i * V);
    Vec16<T> vr;
    if (GRAD == 1) vr = ld16_stream(ref + i * V);
    float b = 0.f;
    if (bias != nullptr) b = to_f(bias[(int)((i / inner_vec) % C)]);
    Vec16<T> vy;
#pragma unroll
    for (int j = 0; j < V; ++j)
      vy.set(j, act_apply<ACT, GRAD>(vx.get(j) + b, GRAD == 1 ? vr.get(j) : 0.f, alpha, scale));
    st16(y + i * V, vy);
  }
}

template <typename T, int ACT, int GRAD>
__global__ void __launch_bounds__(256)
bias_act_scalar_kernel(const T *__restrict__ x, const T *__restrict__ bias,
                       const T *__restrict__ ref, T *__restrict__ y, int64_t n, int C,
                       int64_t inner, float alpha, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = to_f(x[i]);
    if (bias != nullptr) v += to_f(bias[(int)((i / inner) % C)]);
    const float r = (GRAD == 1) ? to_f(ref[i]) : 0.f;
    y[i] = from_f<T>(act_apply<ACT, GRAD>(v, r, alpha, scale));
  }
}

template <typename T, int ACT, int GRAD>
static int launch_bias_act(const void *x, const void *bias, const void *ref, void *y, int64_t n,
                           int C, int64_t inner, float alpha, float scale, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  const bool vec = (inner % V == 0) && aligned16(x) && aligned16(y) && (GRAD != 1 || aligned16(ref));
  const int threads = 256;
  const int64_t work = vec ? n / V : n;
  int64_t blocks = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (vec)
    bias_act_vec_kernel<T, ACT, GRAD><<<(unsigned)blocks, threads, 0, st>>>(
        (const T *)x, (const T *)bias, (const T *)ref, (T *)y, n / V, C, inner / V, alpha, scale);
  else
    bias_act_scalar_kernel<T, ACT, GRAD><<<(unsigned)blocks, threads, 0, st>>>(
        (const T *)x, (const T *)bias, (const T *)ref, (T *)y, n, C, inner, alpha, scale);
  return 0;
}

// Backward over rows r = n*C + c of `inner` contiguous elements.
// grid = (rows, chunks); each CTA handles one chunk of one row, so its partial bias
// gradient belongs to a single channel.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
bias_act_bwd_kernel(const T *__restrict__ dy, const T *__restrict__ out, T *__restrict__ dx,
                    float *__restrict__ db, int C, int64_t inner, int64_t chunk, float alpha,
                    float scale) {
  __shared__ float red[32];
  constexpr int V = VEC ? Vec16<T>::N : 1;
  const int64_t row = blockIdx.x;
  const int64_t base = row * inner;
  const int64_t lo = (int64_t)blockIdx.y * chunk;
  int64_t hi = lo + chunk;
  if (hi > inner) hi = inner;
  float acc = 0.f;
  if (VEC) {
    for (int64_t j = lo + (int64_t)threadIdx.x * V; j < hi; j += (int64_t)blockDim.x * V) {
      Vec16<T> g = ld16_stream(dy + base + j);
      Vec16<T> o = ld16_stream(out + base + j);
      Vec16<T> r;
#pragma unroll
      for (int k = 0; k < Vec16<T>::N; ++k) {
        const float gv = g.get(k);
        const float v = ((o.get(k) > 0.f) ? gv : gv * alpha) * scale;
        r.set(k, v);
        acc += v;
      }
      st16(dx + base + j, r);
    }
  } else {
    for (int64_t j = lo + threadIdx.x; j < hi; j += blockDim.x) {
      const float gv = to_f(dy[base + j]);
      const float v = ((to_f(out[base + j]) > 0.f) ? gv : gv * alpha) * scale;
      dx[base + j] = from_f<T>(v);
      acc += v;
    }
  }
  if (db != nullptr) {
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(db + (int)(row % C), acc);
  }
}

// inner == 1 ([N, C] activations of the D epilogue): one thread per channel column.
template <typename T>
__global__ void bias_act_bwd_2d_kernel(const T *__restrict__ dy, const T *__restrict__ out,
                                       T *__restrict__ dx, float *__restrict__ db, int64_t N,
                                       int C, float alpha, float scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int64_t n = 0; n < N; ++n) {
    const int64_t i = n * C + c;
    const float gv = to_f(dy[i]);
    const float v = ((to_f(out[i]) > 0.f) ? gv : gv * alpha) * scale;
    dx[i] = from_f<T>(v);
    acc += v;
  }
  if (db != nullptr) atomicAdd(db + c, acc);
}

template <typename T>
static int launch_bias_act_bwd(const void *dy, const void *out, void *dx, float *db, int64_t N,
                               int C, int64_t inner, float alpha, float scale, cudaStream_t st) {
  if (inner == 1) {
    bias_act_bwd_2d_kernel<T><<<(C + 127) / 128, 128, 0, st>>>((const T *)dy, (const T *)out,
                                                               (T *)dx, db, N, C, alpha, scale);
    return 0;
  }
  constexpr int V = Vec16<T>::N;
  const bool vec = (inner % V == 0) && aligned16(dy) && aligned16(out) && aligned16(dx);
  const int64_t rows = N * C;
  int64_t chunk = 256 * V * 4;  // 4 vectors per thread
  if (chunk > inner) chunk = ((inner + V - 1) / V) * V;
  const int64_t chunks = (inner + chunk - 1) / chunk;
  if (rows > 0x7fffffff || chunks > 65535) {
    set_error("dusty_bias_act_bwd: tensor too large (rows %lld, chunks %lld)", (long long)rows,
              (long long)chunks);
    return DUSTY_EUNSUPPORTED;
  }
  dim3 grid((unsigned)rows, (unsigned)chunks);
  if (vec)
    bias_act_bwd_kernel<T, true><<<grid, 256, 0, st>>>((const T *)dy, (const T *)out, (T *)dx, db,
                                                       C, inner, chunk, alpha, scale);
  else
    bias_act_bwd_kernel<T, false><<<grid, 256, 0, st>>>((const T *)dy, (const T *)out, (T *)dx,
                                                        db, C, inner, chunk, alpha, scale);
  return 0;
}

}  // namespace dusty

using namespace dusty;

extern "C" int dusty_bias_act(const void *x, const void *bias, const void *ref, void *y,
                              int64_t n_elem, int C, int64_t inner, int act, int grad, float alpha,
                              float scale, int dtype, void *stream) {
  DUSTY_CHECK_ARG(x && y, "null tensor");
  DUSTY_CHECK_ARG(act == 1 || act == 3, "act must be 1 (linear) or 3 (lrelu)");
  DUSTY_CHECK_ARG(grad >= 0 && grad <= 2, "grad must be 0, 1 or 2");
  DUSTY_CHECK_ARG(grad != 1 || act != 3 || ref != nullptr, "grad=1 needs ref");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  DUSTY_CHECK_ARG(inner >= 1 && (bias == nullptr || C >= 1), "bad shape");
  if (n_elem == 0) return DUSTY_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (C < 1) C = 1;
#define DISPATCH(T)                                                                           \
  do {                                                                                        \
    if (act == 1 || grad == 2) {                                                              \
      if (grad == 2) launch_bias_act<T, 1, 2>(x, bias, ref, y, n_elem, C, inner, alpha, scale, st); \
      else launch_bias_act<T, 1, 0>(x, bias, ref, y, n_elem, C, inner, alpha, scale, st);     \
    } else if (grad == 0)                                                                     \
      launch_bias_act<T, 3, 0>(x, bias, ref, y, n_elem, C, inner, alpha, scale, st);          \
    else                                                                                      \
      launch_bias_act<T, 3, 1>(x, bias, ref, y, n_elem, C, inner, alpha, scale, st);          \
  } while (0)
  if (dtype == DUSTY_F32) DISPATCH(float);
  else DISPATCH(__nv_bfloat16);
#undef DISPATCH
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_bias_act_bwd(const void *dy, const void *out, void *dx, float *db, int64_t N,
                                  int C, int64_t inner, float alpha, float scale, int dtype,
                                  void *stream) {
  DUSTY_CHECK_ARG(dy && out && dx, "null tensor");
  DUSTY_CHECK_ARG(N >= 0 && C >= 1 && inner >= 1, "bad shape");
  DUSTY_CHECK_ARG(dtype == DUSTY_F32 || dtype == DUSTY_BF16, "bad dtype");
  if (N == 0) return DUSTY_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = (dtype == DUSTY_F32)
               ? launch_bias_act_bwd<float>(dy, out, dx, db, N, C, inner, alpha, scale, st)
               : launch_bias_act_bwd<__nv_bfloat16>(dy, out, dx, db, N, C, inner, alpha, scale, st);
  if (rc) return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

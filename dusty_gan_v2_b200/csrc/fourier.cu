// a2: Fourier features of the laser angles, and the angle pyramid step.
// Stand-alone producer: out[b, f, p] = sin(fh*el + fw*az + phi), out[b, F+f, p] = cos(..).
// Write-bound (8 B read vs 2F*sizeof(T) B written per pixel): each thread keeps 4 pixels'
// angles in registers and streams FCHUNK frequencies, 16-byte coalesced stores.
// Full-range sincosf (|arg| reaches ~1.2e3 rad) -- no fast-math intrinsics.
#include "common.cuh"

namespace dusty {

constexpr int kFChunk = 8;

template <typename T>
__global__ void __launch_bounds__(256)
fourier_kernel(const float *__restrict__ angle, const float *__restrict__ freqs,
               const float *__restrict__ phase, T *__restrict__ out, int F, int64_t P) {
  const int b = blockIdx.z;
  const int f0 = blockIdx.y * kFChunk;
  const int64_t p0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (p0 >= P) return;
  const float *el = angle + (int64_t)b * 2 * P;
  const float *az = el + P;
  float e[4], a[4];
  const bool full = (p0 + 4 <= P) && ((P & 3) == 0);
  if (full) {
    float4 ev = *reinterpret_cast<const float4 *>(el + p0);
    float4 av = *reinterpret_cast<const float4 *>(az + p0);
    e[0] = ev.x; e[1] = ev.y; e[2] = ev.z; e[3] = ev.w;
    a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      e[j] = (p0 + j < P) ? el[p0 + j] : 0.f;
      a[j] = (p0 + j < P) ? az[p0 + j] : 0.f;
    }
  }
  T *ob = out + (int64_t)b * 2 * F * P;
#pragma unroll 2
  for (int f = f0; f < f0 + kFChunk && f < F; ++f) {
    const float fh = __ldg(freqs + 2 * f), fw = __ldg(freqs + 2 * f + 1), ph = __ldg(phase + f);
    float s[4], c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // (el*fh + az*fw) + phase: a 1x1 conv over the 2 angle channels, then the bias
      const float arg = __fadd_rn(__fadd_rn(__fmul_rn(e[j], fh), __fmul_rn(a[j], fw)), ph);
      sincosf(arg, &s[j], &c[j]);
    }
    T *os = ob + (int64_t)f * P + p0;
    T *oc = ob + (int64_t)(F + f) * P + p0;
    if (full) {
      if (sizeof(T) == 4) {
        __stcs(reinterpret_cast<float4 *>(os), make_float4(s[0], s[1], s[2], s[3]));
        __stcs(reinterpret_cast<float4 *>(oc), make_float4(c[0], c[1], c[2], c[3]));
      } else {
        __nv_bfloat162 s01 = __floats2bfloat162_rn(s[0], s[1]), s23 = __floats2bfloat162_rn(s[2], s[3]);
        __nv_bfloat162 c01 = __floats2bfloat162_rn(c[0], c[1]), c23 = __floats2bfloat162_rn(c[2], c[3]);
        uint2 sv, cv;
        sv.x = *reinterpret_cast<uint32_t *>(&s01); sv.y = *reinterpret_cast<uint32_t *>(&s23);
        cv.x = *reinterpret_cast<uint32_t *>(&c01); cv.y = *reinterpret_cast<uint32_t *>(&c23);
        *reinterpret_cast<uint2 *>(os) = sv;
        *reinterpret_cast<uint2 *>(oc) = cv;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (p0 + j < P) { os[j] = from_f<T>(s[j]); oc[j] = from_f<T>(c[j]); }
    }
  }
}

// cat(sin,cos) -> FIR [1,3,3,1]/8 per axis, stride 2, circular W / replicate H -> atan2.
__global__ void __launch_bounds__(256)
angle_down2_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W,
                   int64_t total) {
  const int Ho = H / 2, Wo = W / 2;
  const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int mx = (int)(idx % Wo);
    const int64_t t = idx / Wo;
    const int my = (int)(t % Ho);
    const int64_t bc = t / Ho;  // b*2 + channel
    const float *src = in + bc * (int64_t)H * W;
    float ss = 0.f, cc = 0.f;
#pragma unroll
    for (int ty = 0; ty < 4; ++ty) {
      int iy = 2 * my + ty - 1;
      iy = iy < 0 ? 0 : (iy >= H ? H - 1 : iy);
      float rs = 0.f, rc = 0.f;
#pragma unroll
      for (int tx = 0; tx < 4; ++tx) {
        int ix = 2 * mx + tx - 1;
        ix = ix < 0 ? ix + W : (ix >= W ? ix - W : ix);
        float s, c;
        sincosf(src[(int64_t)iy * W + ix], &s, &c);
        rs = fmaf(k[tx], s, rs);
        rc = fmaf(k[tx], c, rc);
      }
      ss = fmaf(k[ty], rs, ss);
      cc = fmaf(k[ty], rc, cc);
    }
    out[idx] = atan2f(ss, cc);
  }
}

}  // namespace dusty

using namespace dusty;

extern "C" int dusty_fourier(const float *angle, const float *freqs, const float *phase, void *out,
                             int Ba, int F, int64_t P, int out_dtype, void *stream) {
  DUSTY_CHECK_ARG(angle && freqs && phase && out, "null pointer");
  DUSTY_CHECK_ARG(Ba >= 1 && F >= 1 && P >= 1, "bad shape");
  DUSTY_CHECK_ARG(out_dtype == DUSTY_F32 || out_dtype == DUSTY_BF16, "bad dtype");
  DUSTY_CHECK_ARG(Ba <= 65535, "batch too large");
  DUSTY_CHECK_ARG(aligned16(angle) && aligned16(out), "angle/out must be 16-byte aligned");
  dim3 grid((unsigned)((P + 1023) / 1024), (unsigned)((F + kFChunk - 1) / kFChunk), (unsigned)Ba);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == DUSTY_F32)
    fourier_kernel<float><<<grid, 256, 0, st>>>(angle, freqs, phase, (float *)out, F, P);
  else
    fourier_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(angle, freqs, phase, (__nv_bfloat16 *)out,
                                                        F, P);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_angle_down2(const float *angle_in, float *angle_out, int Ba, int H, int W,
                                 void *stream) {
  DUSTY_CHECK_ARG(angle_in && angle_out, "null pointer");
  DUSTY_CHECK_ARG(Ba >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "bad shape");
  const int64_t total = (int64_t)Ba * 2 * (H / 2) * (W / 2);
  int64_t blocks = (total + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  angle_down2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(angle_in, angle_out, H, W,
                                                                        total);
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

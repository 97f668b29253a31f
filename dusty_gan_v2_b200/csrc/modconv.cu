// a1: C-ABI entry points of the modulated 1x1 contraction; dispatch between the SIMT fp32
// kernel (exact parity path) and the tcgen05 tensor-core kernel (bf16 production path).
#include "common.cuh"

namespace dusty {
int modconv_fwd_simt(const void *wb, const void *x1, const void *x2, const float *bias, void *y,
                     int B, int O, int C1, int C2, int B2, int64_t P, int act, float alpha,
                     float scale, int dtype, int wdtype, cudaStream_t st, const float *const *ema_rows);
int modconv_bwd_dx_simt(const void *wb, const void *dy, void *dx1, int B, int O, int C1, int K,
                        int64_t P, int dtype, int wdtype, cudaStream_t st, const float *const *ema_rows);
int modconv_bwd_dw_simt(const void *dy, const void *x1, const void *x2, float *dwb, int B, int O,
                        int C1, int C2, int B2, int64_t P, int dtype, cudaStream_t st);
// tensor-core path (modconv_tc.cu); returns DUSTY_EUNSUPPORTED when the shape does not fit
int modconv_fwd_tc(const void *wb, const void *x1, const void *x2, const float *bias, void *y,
                   int B, int O, int C1, int C2, int B2, int64_t P, int act, float alpha,
                   float scale, bool batch_fused, cudaStream_t st, bool out_f32, const float *ema, float *sumsq);
bool modconv_fwd_tc_supported(int B, int O, int C1, int C2, int B2, int64_t P);
int modconv_dx_tc(const void *wb, const void *dy, void *dx1, int B, int O, int C1, int K, int64_t P,
                  cudaStream_t st, bool out_f32, const float *ema);
bool modconv_dx_tc_supported(int B, int O, int C1, int K, int64_t P);
int modconv_dw_tc(const void *dy, const void *x1, const void *x2, float *dwb, int B, int O, int C1,
                  int C2, int B2, int64_t P, cudaStream_t st, int64_t ld);
bool modconv_dw_tc_supported(int B, int O, int C1, int C2, int B2, int64_t P);
}  // namespace dusty

using namespace dusty;

static bool dtype_ok(int d) { return d == DUSTY_F32 || d == DUSTY_BF16; }

extern "C" int dusty_modconv_fwd(const void *wb, const void *x1, const void *x2, const float *bias,
                                 void *y, int B, int O, int C1, int C2, int B2, int64_t P, int act,
                                 float alpha, float scale, int dtype, int wdtype, int impl,
                                 const float *ema_var, const float *const *ema_rows, float *sumsq,
                                 void *stream) {
  DUSTY_CHECK_ARG(wb && y, "null pointer");
  DUSTY_CHECK_ARG(!(ema_var && ema_rows), "ema_var and ema_rows are exclusive");
  DUSTY_CHECK_ARG(B >= 1 && B <= 65535 && O >= 1 && C1 >= 0 && C2 >= 0 && C1 + C2 >= 1 && P >= 1,
                  "bad shape");
  DUSTY_CHECK_ARG((C1 == 0 || x1) && (C2 == 0 || x2), "missing source tensor");
  DUSTY_CHECK_ARG(C2 == 0 || B2 == B || B2 == 1, "x2 batch must be B or 1");
  DUSTY_CHECK_ARG(act == 1 || act == 3, "act must be 1 or 3");
  DUSTY_CHECK_ARG(dtype_ok(dtype) && dtype_ok(wdtype), "bad dtype");
  DUSTY_CHECK_ARG(impl >= 0 && impl <= 4,
                  "impl must be 0 (auto), 1 (simt), 2 (tcgen05), 3 (tcgen05, per-sample tiles only) or "
                  "4 (tcgen05, fp32 output)");
  cudaStream_t st = (cudaStream_t)stream;
  if (C1 == 0) x1 = x2;  // keep pointers valid for address arithmetic
  if (C2 == 0) { x2 = x1; B2 = B; }
  const bool tc_ok = dtype == DUSTY_BF16 && wdtype == DUSTY_BF16 &&
                     modconv_fwd_tc_supported(B, O, C1, C2, B2, P);
  if (impl >= 2 && !tc_ok) {
    set_error("dusty_modconv_fwd: tcgen05 path does not support this shape/dtype");
    return DUSTY_EUNSUPPORTED;
  }
  int rc;
  const bool use_tc = impl >= 2 || (impl == 0 && tc_ok);
  if ((ema_var && !use_tc) || (ema_rows && use_tc) || (sumsq && (!use_tc || impl == 4))) {
    set_error("dusty_modconv_fwd: ema_var / sumsq belong to the tcgen05 (bf16 output) epilogue, ema_rows to "
              "the small-O kernel");
    return DUSTY_EUNSUPPORTED;
  }
  if (use_tc)
    rc = modconv_fwd_tc(wb, x1, x2, bias, y, B, O, C1, C2, B2, P, act, alpha, scale, impl < 3, st,
                        impl == 4, ema_var, sumsq);
  else
    rc = modconv_fwd_simt(wb, x1, x2, bias, y, B, O, C1, C2, B2, P, act, alpha, scale, dtype,
                          wdtype, st, ema_rows);
  if (rc) return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_modconv_bwd_dx(const void *wb, const void *dy, void *dx1, int B, int O, int C1,
                                    int K, int64_t P, int dtype, int wdtype, int impl,
                                    const float *ema_var, const float *const *ema_rows, void *stream) {
  DUSTY_CHECK_ARG(wb && dy && dx1, "null pointer");
  DUSTY_CHECK_ARG(!(ema_var && ema_rows), "ema_var and ema_rows are exclusive");
  DUSTY_CHECK_ARG(B >= 1 && B <= 65535 && O >= 1 && C1 >= 1 && K >= C1 && P >= 1, "bad shape");
  DUSTY_CHECK_ARG(dtype_ok(dtype) && dtype_ok(wdtype), "bad dtype");
  DUSTY_CHECK_ARG(impl >= 0 && impl <= 4, "impl must be 0 (auto), 1 (simt), 2 / 3 (tcgen05) or 4 (tcgen05, fp32 output)");
  const bool tc_ok = dtype == DUSTY_BF16 && wdtype == DUSTY_BF16 && modconv_dx_tc_supported(B, O, C1, K, P);
  if (impl >= 2 && !tc_ok) {
    set_error("dusty_modconv_bwd_dx: tcgen05 path does not support this shape/dtype");
    return DUSTY_EUNSUPPORTED;
  }
  int rc;
  const bool use_tc = impl >= 2 || (impl == 0 && tc_ok);
  if ((ema_var && !use_tc) || (ema_rows && use_tc)) {
    set_error("dusty_modconv_bwd_dx: ema_var is applied by the tcgen05 epilogue, ema_rows by the small-O kernel");
    return DUSTY_EUNSUPPORTED;
  }
  if (use_tc)
    rc = modconv_dx_tc(wb, dy, dx1, B, O, C1, K, P, (cudaStream_t)stream, impl == 4, ema_var);
  else
    rc = modconv_bwd_dx_simt(wb, dy, dx1, B, O, C1, K, P, dtype, wdtype, (cudaStream_t)stream, ema_rows);
  if (rc) return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_modconv_bwd_dw(const void *dy, const void *x1, const void *x2, float *dwb,
                                    int B, int O, int C1, int C2, int B2, int64_t P, int dtype,
                                    int impl, long long dw_ld, void *stream) {
  DUSTY_CHECK_ARG(dy && dwb, "null pointer");
  DUSTY_CHECK_ARG(dw_ld == 0 || dw_ld >= C1 + C2, "dw_ld: row pitch of dwb, 0 = dense");
  DUSTY_CHECK_ARG(B >= 1 && B <= 65535 && O >= 1 && C1 >= 0 && C2 >= 0 && C1 + C2 >= 1 && P >= 1,
                  "bad shape");
  DUSTY_CHECK_ARG((C1 == 0 || x1) && (C2 == 0 || x2), "missing source tensor");
  DUSTY_CHECK_ARG(C2 == 0 || B2 == B || B2 == 1, "x2 batch must be B or 1");
  DUSTY_CHECK_ARG(dtype_ok(dtype), "bad dtype");
  DUSTY_CHECK_ARG(impl >= 0 && impl <= 3, "impl must be 0 (auto), 1 (simt), 2 or 3 (tcgen05)");
  if (C1 == 0) x1 = x2;
  if (C2 == 0) { x2 = x1; B2 = B; }
  const bool tc_ok = dtype == DUSTY_BF16 && modconv_dw_tc_supported(B, O, C1, C2, B2, P);
  if (impl >= 2 && !tc_ok) {
    set_error("dusty_modconv_bwd_dw: tcgen05 path does not support this shape/dtype");
    return DUSTY_EUNSUPPORTED;
  }
  int rc;
  const bool dw_tc = impl >= 2 || (impl == 0 && tc_ok);
  if (dw_ld > C1 + C2 && !dw_tc) {
    set_error("dusty_modconv_bwd_dw: a strided result needs the tcgen05 path");
    return DUSTY_EUNSUPPORTED;
  }
  if (dw_tc)
    rc = modconv_dw_tc(dy, x1, x2, dwb, B, O, C1, C2, B2, P, (cudaStream_t)stream, dw_ld);
  else
    rc = modconv_bwd_dw_simt(dy, x1, x2, dwb, B, O, C1, C2, B2, P, dtype, (cudaStream_t)stream);
  if (rc) return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

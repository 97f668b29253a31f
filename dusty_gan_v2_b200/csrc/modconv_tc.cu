// a1: tcgen05 tensor-core implementation of the modulated 1x1 contraction (bf16 operands,
// fp32 accumulation in TMEM).  Placeholder until the kernel lands: reports "unsupported" so
// the dispatcher keeps using the SIMT kernel.
#include "common.cuh"

namespace dusty {
bool modconv_fwd_tc_supported(int, int, int, int, int, int64_t) { return false; }
int modconv_fwd_tc(const void *, const void *, const void *, const float *, void *, int, int, int,
                   int, int, int64_t, int, float, float, cudaStream_t) {
  set_error("modconv_fwd_tc: not built");
  return DUSTY_EUNSUPPORTED;
}
}  // namespace dusty

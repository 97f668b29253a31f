// Experiment: does a K-major SWIZZLE_128B UMMA A-operand descriptor whose start address is
// shifted by a whole number of 128-byte rows (not a multiple of the 8-row swizzle atom) read
// rows [shift, shift+128) of a tile that TMA wrote with the same swizzle?
// D[m][n] = sum_k A[m + shift][k] * I[n][k]  (B = identity)  ->  D[m][n] == X[m + shift][n].
// Variants: base_offset field 0, or ((start >> 7) & 7).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../dusty_gan_v2_b200/csrc \
//        umma_rowshift.cu ../../build/csrc/core.o -o /tmp/umma_rowshift -lcuda
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
using namespace dusty;

constexpr int ROWS = 144;  // 128 + 16 rows of slack
__global__ void __launch_bounds__(128)
probe(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_i,
      float *out, int shift, int use_base_offset, int sw64) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *a = smem;                      // ROWS x 128 B (or 64 B)
  uint8_t *b = smem + 32768;              // 64 x 128 B
  uint64_t *bar = (uint64_t *)(smem + 49152);
  uint64_t *done = bar + 1;
  uint32_t *slot = (uint32_t *)(done + 1);
  const int rowb = sw64 ? 64 : 128;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, ROWS * rowb + 64 * rowb);
    tma_load_3d(a, &map_x, bar, 0, 0, 0);
    tma_load_3d(b, &map_i, bar, 0, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const int N = sw64 ? 32 : 64;
    const uint32_t idesc = make_idesc(128, N, false, false);
    const int ksteps = sw64 ? 2 : 4;
    for (int k16 = 0; k16 < ksteps; ++k16) {
      const uint32_t sa = smem_u32(a) + shift * rowb + k16 * 32;
      uint64_t ad = make_desc(sa, 16, 8 * rowb);
      uint64_t bd = make_desc(smem_u32(b) + k16 * 32, 16, 8 * rowb);
      if (sw64) {  // layout type 4 = SWIZZLE_64B
        ad = (ad & ~((uint64_t)7 << 61)) | ((uint64_t)4 << 61);
        bd = (bd & ~((uint64_t)7 << 61)) | ((uint64_t)4 << 61);
      }
      if (use_base_offset) ad |= (uint64_t)((sa >> 7) & 7) << 49;
      umma_bf16(tmem, ad, bd, idesc, k16 > 0);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp * 32 + lane;
  const int N = sw64 ? 32 : 64;
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[row * 64 + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

static bool mk(CUtensorMap *m, void *p, int cols, int rows, int boxr, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t str[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * rows * 2};
  cuuint32_t box[3] = {(cuuint32_t)cols, (cuuint32_t)boxr, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main() {
  for (int sw64 = 0; sw64 < 2; ++sw64) {
    const int K = sw64 ? 32 : 64;
    std::vector<__nv_bfloat16> hx(ROWS * K), hi(64 * K);
    for (int r = 0; r < ROWS; ++r) for (int c = 0; c < K; ++c) hx[r * K + c] = __float2bfloat16((float)(r * 64 + c) / 8.f);
    for (int r = 0; r < 64; ++r) for (int c = 0; c < K; ++c) hi[r * K + c] = __float2bfloat16(r == c ? 1.f : 0.f);
    __nv_bfloat16 *dx, *di; float *dout;
    cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&di, hi.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(di, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap mx, mi;
    const CUtensorMapSwizzle sw = sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    if (!mk(&mx, dx, K, ROWS, ROWS, sw) || !mk(&mi, di, K, 64, 64, sw)) { printf("map failed\n"); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    for (int ubo = 0; ubo < 2; ++ubo)
      for (int shift = 0; shift <= 9; ++shift) {
        cudaMemset(dout, 0, 128 * 64 * 4);
        probe<<<1, 128, 60 * 1024>>>(mx, mi, dout, shift, ubo, sw64);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sw64=%d ubo=%d shift=%d: CUDA error %s\n", sw64, ubo, shift, cudaGetErrorString(e)); return 2; }
        std::vector<float> ho(128 * 64);
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0; const int N = sw64 ? 32 : 64;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
          const float want = __bfloat162float(hx[(m + shift) * K + n]);
          if (ho[m * 64 + n] != want) ++bad;
        }
        printf("swizzle=%s base_offset_field=%d shift=%d rows: mismatches=%d\n", sw64 ? "64B" : "128B", ubo, shift, bad);
      }
  }
  return 0;
}

// Experiment: issue rate of tcgen05.mma (cta_group::1, kind::f16, M = 128) from shared memory
// as a function of N and of the A operand's major-ness.  One CTA per SM, one thread issues
// `iters` UMMAs back to back on a fixed (zeroed) operand buffer, then commits and waits;
// cycles per UMMA = (clock after the commit barrier - clock before the first issue) / iters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../dusty_gan_v2_b200/csrc \
//        umma_rate.cu ../../build/csrc/core.o -o ../../build/umma_rate -lcuda
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
using namespace dusty;

__global__ void __launch_bounds__(128)
rate(long long *out, int N, int a_mn, int b_mn, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *a = smem;                      // 128 x 64 bf16 = 16 KiB
  uint8_t *b = smem + 16384;              // 256 x 64 bf16 = 32 KiB
  uint64_t *done = (uint64_t *)(smem + 49152);
  uint32_t *slot = (uint32_t *)(done + 1);
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(done, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(slot, 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = make_idesc(128, N, a_mn != 0, b_mn != 0);
    const uint32_t d_hi = desc_hi(1024, 2);
    const uint32_t a_lo = a_mn ? desc_lo(smem_u32(a), 8192) : desc_lo(smem_u32(a), 16);
    const uint32_t b_lo = b_mn ? desc_lo(smem_u32(b), 8192) : desc_lo(smem_u32(b), 16);
    const uint32_t a_step = a_mn ? (2048 >> 4) : 2, b_step = b_mn ? (2048 >> 4) : 2;
    long long t0 = clock64();
    if (elect_one_sync()) {
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16)
          umma_bf16_lh(tmem, a_lo + k16 * a_step, d_hi, b_lo + k16 * b_step, d_hi, idesc, 1u);
      }
      umma_commit(done);
    }
    __syncwarp();
    mbar_wait(done, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
  long long *d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200);
  const int iters = 512;
  for (int ctas : {1, 148})
    for (int a_mn = 0; a_mn < 2; ++a_mn)
      for (int b_mn = 0; b_mn < 2; ++b_mn)
        for (int N : {32, 64, 128, 256}) {
          rate<<<ctas, 128, 51200>>>(d, N, a_mn, b_mn, iters);
          rate<<<ctas, 128, 51200>>>(d, N, a_mn, b_mn, iters);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          std::vector<long long> h(ctas);
          cudaMemcpy(h.data(), d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (auto v : h) mx = v > mx ? v : mx;
          const double cyc = (double)mx / (iters * 4);
          printf("{\"ctas\": %d, \"a_mn\": %d, \"b_mn\": %d, \"N\": %d, \"cycles_per_umma\": %.1f, \"flop_per_cycle\": %.0f}\n",
                 ctas, a_mn, b_mn, N, cyc, 2.0 * 128 * N * 16 / cyc);
        }
  return 0;
}

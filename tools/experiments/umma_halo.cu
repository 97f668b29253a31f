// Experiment: what does the MMA stream of the halo-resident convolution cost per UMMA?
// Emulates conv_halo_tc_kernel's issue pattern on static (zeroed) buffers: a patch of pixel rows
// (ROWB = 64 or 128 bytes, K-major, 64B / 128B swizzle), 9 taps read as row-shifted views of it,
// N = BN output channels, (a) alone, (b) with four other warps streaming tcgen05.ld from the
// second accumulator (the epilogue's traffic), (c) with a TMA-sized bulk copy landing in a
// second buffer meanwhile.  cycles per UMMA = (commit wait - first issue) / UMMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../dusty_gan_v2_b200/csrc \
//        umma_halo.cu ../../build/csrc/core.o -o ../../build/umma_halo -lcuda
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
using namespace dusty;

template <int ROWB, int BN>
__global__ void __launch_bounds__(192)
halo_rate(long long *out, int PW, int MB, int iters, int with_ld, int variant) {
  const int unrolled = 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t kLayout = ROWB == 64 ? 4 : 2;
  constexpr int kWTap = BN * ROWB;
  uint8_t *w = smem;                               // 9 taps
  uint8_t *p = smem + ((9 * kWTap + 1023) & ~1023);  // patch: (MB*128 + 2*PW + 2) rows
  const int patch_bytes = ((MB * 128 + 2 * PW + 2) * ROWB + 1023) & ~1023;
  uint64_t *done = (uint64_t *)(p + patch_bytes);
  uint64_t *dummy = done + 1;
  uint32_t *slot = (uint32_t *)(done + 10);
  volatile int *stop = (volatile int *)(slot + 1);
  for (int i = threadIdx.x; i < (int)((uint8_t *)done - smem) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(done, 1); for (int i = 0; i < 8; ++i) mbar_init(&dummy[i], 1); fence_barrier_init(); *stop = 0; }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(slot, 2 * BN < 32 ? 32 : 2 * BN);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc(128, BN, false, false);
    const uint32_t d_hi = desc_hi(8 * ROWB, kLayout);
    const uint32_t p_lo0 = desc_lo(smem_u32(p), 16);
    const uint32_t w_lo0 = desc_lo(smem_u32(w), 16);
    constexpr uint32_t kRow16 = ROWB >> 4;
    long long t0 = clock64();
    int blk = 0;
    for (int it = 0; it < iters; ++it)
      for (int mb = 0; mb < MB; ++mb, ++blk) {
        if ((variant & 8) && blk >= 4) mbar_wait(&dummy[(blk - 4) & 7], (uint32_t)(((blk - 4) >> 3) & 1));
        if (variant & 1) tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t tacc = tmem + ((variant & 4) ? (uint32_t)((blk & 1) * BN) : 0u);
          uint32_t a_row = p_lo0 + (uint32_t)(mb * 128) * kRow16;
          uint32_t first = 0, w_lo = w_lo0;
          int s2 = 0;
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int k16 = 0; k16 < ROWB / 32; ++k16) {
              umma_bf16_lh(tacc, a_row + (uint32_t)s2 * kRow16 + k16 * 2, d_hi, w_lo + k16 * 2, d_hi, idesc, first);
              first = 1u;
            }
            w_lo += kWTap >> 4;
            if (++s2 == 3) { s2 = 0; a_row += (uint32_t)PW * kRow16; }
          }
          if (variant & 2) umma_commit(&dummy[blk & 7]);
        }
        __syncwarp();
      }
    if (elect_one_sync()) umma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    long long t1 = clock64();
    *stop = 1;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp >= 2 && with_ld) {
    // epilogue-like traffic on accumulator 1 (lanes of this warp's quarter)
    const uint32_t tacc = tmem + (uint32_t)BN + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0;
    long long n = 0;
    const long long t0 = clock64();
    while (!*stop) {
      uint32_t r[16];
      tmem_ld16(tacc, r);
      tmem_ld_wait();
      sink += r[0];
      ++n;
    }
    const long long t1 = clock64();
    if (sink == 0x12345678u) out[0] = 0;
    if (warp == 2 && (threadIdx.x & 31) == 0) out[148 + blockIdx.x] = (t1 - t0) / (n > 0 ? n : 1);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 2 * BN < 32 ? 32 : 2 * BN); }
}

template <int ROWB, int BN>
void run(long long *d, const char *tag) {
  const int smem = 160 * 1024;
  cudaFuncSetAttribute(halo_rate<ROWB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 64, MB = 4;
  for (int PW : {64})
    for (int unrolled : {7})
      for (int with_ld = 0; with_ld < 2; ++with_ld) {
        for (int rep = 0; rep < 2; ++rep) halo_rate<ROWB, BN><<<148, 192, smem>>>(d, PW, MB, iters, with_ld, unrolled);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return; }
        std::vector<long long> h(296);
        cudaMemcpy(h.data(), d, 296 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        if (with_ld) printf("  cycles per tcgen05.ld.x16 + wait (warp 2, CTA 0): %lld\n", h[148]);
        const int ummas = iters * MB * 9 * (ROWB / 32);
        printf("{\"kernel\": \"%s\", \"PW\": %d, \"variant(1=fence,2=commit,4=rotate acc)\": %d, \"with_tmem_ld\": %d, \"cycles_per_umma\": %.1f}\n", tag, PW,
               unrolled, with_ld, (double)mx / ummas);
      }
}

int main() {
  long long *d;
  cudaMalloc(&d, 296 * sizeof(long long));
  run<64, 32>(d, "C32_N32_sw64");
  run<64, 64>(d, "C32_N64_sw64");
  run<128, 64>(d, "C64_N64_sw128");
  return 0;
}

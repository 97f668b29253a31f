// a11: the dense linears of the discriminator epilogue (gans/models/dusty_v2.py:382-384,
// EqualLR(Linear(65536 -> 512)) and EqualLR(Linear(512 -> 1))) on own kernels.
//
// C[M, N] (+)= alpha * A[M, K] * B[N, K]^T, fp32 accumulation and output, on tcgen05:
//   dusty_gemm_tf32  kind::tf32 straight on fp32 operands, K-major (K contiguous) -- the forward
//                    y[B, 512] = x[B, 65536] * W[512, 65536]^T streams the 134 MB fp32 master
//                    weight from HBM once, with no cast / scale pass over it; split-K.
//   dusty_gemm_bf16  kind::f16 on bf16 operands of either major-ness, so the two gradients read
//                    their tensors in place:  dx[B, 65536] = dy[B, 512] * W (W as an MN-major B
//                    operand),  dW[512, 65536] = dy^T * x (both MN-major, K = B).
//   (Transposed 32-bit operands need the 32-byte-atom swizzle family; with the plain 128-byte
//   swizzle kind::tf32 returned zeros for MN-major operands on sm_100a -- tools/debug/gemm_dbg.py
//   -- hence bf16 for the gradients, as in the rest of the low-precision path.)
// Tile: 128 x BN accumulator in TMEM, K = 32 fp32 (one 128-byte swizzle row) per stage, four
// UMMAs (K = 8) per stage; warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue:
// TMEM -> registers -> swizzled staging tile -> ONE TMA store (or TMA reduce-add for split-K) per
// 32 columns.
//
// dusty_gemm_simt: the 512 -> 1 head and its gradients (tiny GEMMs, any strides): CUDA cores, fp32.
#include "tc_common.cuh"

namespace dusty {
namespace {

constexpr int kGM = 128;                 // UMMA M
constexpr int kGThreads = 64 + 128;
// One stage holds one 128-byte swizzle row of K per operand row: 32 fp32 (kind::tf32, UMMA K = 8)
// or 64 bf16 (kind::f16, UMMA K = 16) -- four UMMAs of 32 bytes of K per stage either way.
template <typename T> struct Elem;
template <> struct Elem<float> { static constexpr int kPerRow = 32; static constexpr bool kTf32 = true; };
template <> struct Elem<__nv_bfloat16> { static constexpr int kPerRow = 64; static constexpr bool kTf32 = false; };

struct GemmMaps {
  CUtensorMap a, b, c;
};

struct GemmParams {
  int kb_total, kb_per_split;            // 32-wide k blocks
  int reduce;                            // 1: add into C (split-K), 0: store
  float alpha;
};

__device__ __forceinline__ void umma_tf32_lh(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo,
                                             uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D = F32 (bit 4), A = B = TF32 (format 2 at bits 7 and 10), major-ness bits 15 / 16
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}

template <typename T, int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGThreads, 1)
gemm_tc_kernel(const __grid_constant__ GemmMaps maps, const GemmParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kGK = Elem<T>::kPerRow;                     // K elements per stage
  constexpr int kABytes = kGM * 128;                        // 16 KiB
  constexpr int kBBytes = BN * 128;
  constexpr int kMnBlock = kGK * 128;                       // MN-major: [kGK k rows][128 bytes of M / N]
  constexpr int kStage = kGM * 32 * 4;                      // staging: [128 rows][32 fp32]
  uint8_t *a_base = smem;
  uint8_t *b_base = smem + STAGES * kABytes;
  uint8_t *stage = b_base + STAGES * kBBytes;
  uint64_t *full = (uint64_t *)(stage + kStage);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;
  uint32_t *tmem_slot = (uint32_t *)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kGM, n0 = blockIdx.y * BN;
  const int kb_begin = blockIdx.z * prm.kb_per_split;
  const int kb_end = min(kb_begin + prm.kb_per_split, prm.kb_total);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    RingPos r;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      const int s = r.s;
      mbar_wait(&empty[s], r.ph ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&full[s], kABytes + kBBytes);
        uint8_t *ad = a_base + s * kABytes, *bd = b_base + s * kBBytes;
        if (A_MN) {                                         // [kGK k rows][128 bytes of m] blocks
#pragma unroll
          for (int j = 0; j < kGM / kGK; ++j)
            tma_load_2d(ad + j * kMnBlock, &maps.a, &full[s], m0 + kGK * j, kb * kGK);
        } else {
          tma_load_2d(ad, &maps.a, &full[s], kb * kGK, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / kGK; ++j)
            tma_load_2d(bd + j * kMnBlock, &maps.b, &full[s], n0 + kGK * j, kb * kGK);
        } else {
          tma_load_2d(bd, &maps.b, &full[s], kb * kGK, n0);
        }
      }
      __syncwarp();
      r.template advance<STAGES>();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = Elem<T>::kTf32 ? make_idesc_tf32(kGM, BN, A_MN, B_MN)
                                              : make_idesc(kGM, BN, A_MN, B_MN);
    // K-major: 32 bytes per UMMA k step inside the swizzle row, SBO = 8 rows (1 KiB).  MN-major
    // (bf16 only): a k step is 16 rows of 128 bytes (2 KiB), LBO = next 64-element block along
    // M / N, SBO = next 8 k rows (1 KiB).
    constexpr uint32_t kStepMn = (kGK / 4) * 128;           // bytes of one UMMA k step, MN-major
    const uint32_t d_hi = desc_hi(1024, 2);
    const uint32_t a_lo0 = desc_lo(smem_u32(a_base), A_MN ? kMnBlock : 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(b_base), B_MN ? kMnBlock : 16);
    constexpr uint32_t a_step = A_MN ? (kStepMn >> 4) : 2, b_step = B_MN ? (kStepMn >> 4) : 2;
    RingPos r;
    bool ready = false;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      const int s = r.s;
      if (!ready) mbar_wait(&full[s], r.ph);
      tc_fence_after();
      RingPos nx = r;
      nx.template advance<STAGES>();
      ready = kb + 1 < kb_end && mbar_try_wait(&full[nx.s], nx.ph);
      if (elect_one_sync()) {
        const uint32_t a_lo = a_lo0 + (uint32_t)s * (kABytes >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)s * (kBBytes >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acc = (kb > kb_begin || ks > 0) ? 1u : 0u;
          if (Elem<T>::kTf32) umma_tf32_lh(tmem_acc, a_lo + ks * a_step, d_hi, b_lo + ks * b_step, d_hi, idesc, acc);
          else umma_bf16_lh(tmem_acc, a_lo + ks * a_step, d_hi, b_lo + ks * b_step, d_hi, idesc, acc);
        }
        umma_commit(&empty[s]);
        if (kb == kb_end - 1) umma_commit(acc_full);
      }
      __syncwarp();
      r = nx;
    }
  } else if (kb_end > kb_begin) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool issuer = threadIdx.x == 64;
    const uint32_t stage_u32 = smem_u32(stage);
    const uint32_t srow = stage_u32 + (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)row & 7u;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[2][16];
      tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v[0]);
      tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c + 16), v[1]);
      tmem_ld_wait();
      if (issuer) tma_store_wait_read<0>();
      named_bar_sync(1, 128);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {                      // 16-byte chunks of the 128-byte row
        const uint32_t *src = &v[ch >> 2][(ch & 3) * 4];
        const uint32_t addr = srow + (((uint32_t)ch ^ swz) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                     "r"(__float_as_uint(__uint_as_float(src[0]) * prm.alpha)),
                     "r"(__float_as_uint(__uint_as_float(src[1]) * prm.alpha)),
                     "r"(__float_as_uint(__uint_as_float(src[2]) * prm.alpha)),
                     "r"(__float_as_uint(__uint_as_float(src[3]) * prm.alpha))
                     : "memory");
      }
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (issuer) {
        if (prm.reduce) tma_reduce_add_2d(&maps.c, stage_u32, n0 + c, m0);
        else tma_store_2d(&maps.c, stage_u32, n0 + c, m0);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, BN < 32 ? 32 : BN);
  }
}

bool make_map2(CUtensorMap *m, const void *ptr, bool bf16, uint64_t d0, uint64_t d1, uint64_t stride1_elems,
               uint32_t box0, uint32_t box1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {stride1_elems * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t es[2] = {1, 1};
  return enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
             const_cast<void *>(ptr), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T, int BN, bool A_MN, bool B_MN>
int launch_gemm(const GemmMaps &maps, const GemmParams &prm, int mt, int nt, int splits, cudaStream_t st) {
  constexpr int STAGES = BN == 256 ? 4 : 5;
  constexpr int smem = STAGES * (kGM * 128 + BN * 128) + kGM * 128 + 256 + 1024;
  static bool configured = false;
  if (int rc = set_smem(gemm_tc_kernel<T, BN, STAGES, A_MN, B_MN>, smem, &configured)) return rc;
  gemm_tc_kernel<T, BN, STAGES, A_MN, B_MN><<<dim3((unsigned)mt, (unsigned)nt, (unsigned)splits), kGThreads, smem, st>>>(
      maps, prm);
  return 0;
}

// C fp32 [M, N] (+)= alpha * A * B^T; operands fp32 (K-major only) or bf16 (either major)
int gemm_dispatch(const char *who, bool bf16, const void *a, const void *b, float *c, int M, int N, int K,
                  int a_mn, int b_mn, long long lda, long long ldb, long long ldc, float alpha, int accumulate,
                  cudaStream_t st) {
  const int kGK = bf16 ? 64 : 32;
  const int BN = N > 128 ? 256 : (N > 64 ? 128 : 64);
  GemmMaps maps;
  bool ok;
  // K-major: dims (K, rows), one box of kGK k x 128 / BN rows.  MN-major: dims (rows, K), boxes of
  // kGK rows x kGK k.
  ok = a_mn ? make_map2(&maps.a, a, bf16, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kGK, kGK)
            : make_map2(&maps.a, a, bf16, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kGK, kGM);
  ok = ok && (b_mn ? make_map2(&maps.b, b, bf16, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, kGK, kGK)
                   : make_map2(&maps.b, b, bf16, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, kGK, (uint32_t)BN));
  ok = ok && make_map2(&maps.c, c, false, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32, kGM);
  if (!ok) {
    set_error("%s: cuTensorMapEncodeTiled failed", who);
    return DUSTY_ECUDA;
  }
  const int mt = (M + kGM - 1) / kGM, nt = (N + BN - 1) / BN;
  GemmParams prm;
  prm.kb_total = (K + kGK - 1) / kGK;
  // split-K when the output tiles cannot fill the machine (the 65536-deep forward GEMM has two)
  int splits = 1;
  if (mt * nt < num_sms()) {
    splits = (num_sms() + mt * nt - 1) / (mt * nt);
    if (splits > prm.kb_total / 8) splits = prm.kb_total / 8 > 0 ? prm.kb_total / 8 : 1;
  }
  prm.kb_per_split = (prm.kb_total + splits - 1) / splits;
  splits = (prm.kb_total + prm.kb_per_split - 1) / prm.kb_per_split;
  prm.reduce = (splits > 1 || accumulate) ? 1 : 0;
  prm.alpha = alpha;
  if (splits > 1 && !accumulate) {
    if (cudaMemset2DAsync(c, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st) != cudaSuccess) {
      set_error("%s: memset failed", who);
      return DUSTY_ECUDA;
    }
  }
#define DUSTY_GEMM_MAJORS(T, bn)                                                                    \
  (a_mn ? (b_mn ? launch_gemm<T, bn, true, true>(maps, prm, mt, nt, splits, st)                     \
                : launch_gemm<T, bn, true, false>(maps, prm, mt, nt, splits, st))                   \
        : (b_mn ? launch_gemm<T, bn, false, true>(maps, prm, mt, nt, splits, st)                    \
                : launch_gemm<T, bn, false, false>(maps, prm, mt, nt, splits, st)))
  if (bf16) {
    if (BN == 256) return DUSTY_GEMM_MAJORS(__nv_bfloat16, 256);
    if (BN == 128) return DUSTY_GEMM_MAJORS(__nv_bfloat16, 128);
    return DUSTY_GEMM_MAJORS(__nv_bfloat16, 64);
  }
#undef DUSTY_GEMM_MAJORS
  if (BN == 256) return launch_gemm<float, 256, false, false>(maps, prm, mt, nt, splits, st);
  if (BN == 128) return launch_gemm<float, 128, false, false>(maps, prm, mt, nt, splits, st);
  return launch_gemm<float, 64, false, false>(maps, prm, mt, nt, splits, st);
}

// ---- tiny GEMMs (the 512 -> 1 head and its gradients): CUDA cores, fp32, any strides.
// C[m, n] = alpha * sum_k A[m, k] * B[n, k]; one warp per output when K is long, else one thread.
__global__ void gemm_simt_warp_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                      float *__restrict__ c, int M, int N, int K, long long a_sm, long long a_sk,
                                      long long b_sn, long long b_sk, long long c_sm, long long c_sn, float alpha) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)M * N) return;
  const int m = (int)(wid / N), n = (int)(wid % N);
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(a[m * a_sm + k * a_sk], b[n * b_sn + k * b_sk], acc);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) c[m * c_sm + n * c_sn] = acc * alpha;
}
__global__ void gemm_simt_thread_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                        float *__restrict__ c, int M, int N, int K, long long a_sm, long long a_sk,
                                        long long b_sn, long long b_sk, long long c_sm, long long c_sn,
                                        float alpha) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);             // n fastest: coalesced when B / C are n-contiguous
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(a[m * a_sm + k * a_sk], b[n * b_sn + k * b_sk], acc);
  c[m * c_sm + n * c_sn] = acc * alpha;
}

}  // namespace
}  // namespace dusty

using namespace dusty;

extern "C" int dusty_gemm_tf32(const float *a, const float *b, float *c, int M, int N, int K, long long lda,
                               long long ldb, long long ldc, float alpha, int accumulate, void *stream) {
  DUSTY_CHECK_ARG(a && b && c, "null pointer");
  DUSTY_CHECK_ARG(get_encode() != nullptr, "cuTensorMapEncodeTiled unavailable");
  DUSTY_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty problem");
  DUSTY_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(c), "16-byte alignment");
  DUSTY_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 && ldc >= N && lda >= K && ldb >= K,
                  "leading dimensions: multiples of 4, at least the row length");
  if (int rc = gemm_dispatch("dusty_gemm_tf32", false, a, b, c, M, N, K, 0, 0, lda, ldb, ldc, alpha, accumulate,
                             (cudaStream_t)stream))
    return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_gemm_bf16(const void *a, const void *b, float *c, int M, int N, int K, int a_mn,
                               int b_mn, long long lda, long long ldb, long long ldc, float alpha,
                               int accumulate, void *stream) {
  DUSTY_CHECK_ARG(a && b && c, "null pointer");
  DUSTY_CHECK_ARG(get_encode() != nullptr, "cuTensorMapEncodeTiled unavailable");
  DUSTY_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty problem");
  DUSTY_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(c), "16-byte alignment");
  DUSTY_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0 && ldc >= N, "leading dimensions: 16-byte multiples");
  DUSTY_CHECK_ARG(lda >= (a_mn ? M : K) && ldb >= (b_mn ? N : K), "leading dimension too small");
  if (int rc = gemm_dispatch("dusty_gemm_bf16", true, a, b, c, M, N, K, a_mn, b_mn, lda, ldb, ldc, alpha,
                             accumulate, (cudaStream_t)stream))
    return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_gemm_simt(const float *a, const float *b, float *c, int M, int N, int K,
                               long long a_sm, long long a_sk, long long b_sn, long long b_sk,
                               long long c_sm, long long c_sn, float alpha, void *stream) {
  DUSTY_CHECK_ARG(a && b && c, "null pointer");
  DUSTY_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty problem");
  // one warp per output for long dot products, one thread per output for K < 64 (e.g. the
  // weight gradient of the 65536 -> 512 linear at batch 8: 33.5 M outputs of 8 terms)
  DUSTY_CHECK_ARG((long long)M * N <= (K >= 64 ? (1LL << 24) : (1LL << 31)),
                  "dusty_gemm_simt is for small outputs (<= 16 M elements of K >= 64 terms)");
  cudaStream_t st = (cudaStream_t)stream;
  const long long outs = (long long)M * N;
  if (K >= 64) {
    gemm_simt_warp_kernel<<<(unsigned)((outs * 32 + 255) / 256), 256, 0, st>>>(a, b, c, M, N, K, a_sm, a_sk, b_sn,
                                                                               b_sk, c_sm, c_sn, alpha);
  } else {
    gemm_simt_thread_kernel<<<(unsigned)((outs + 255) / 256), 256, 0, st>>>(a, b, c, M, N, K, a_sm, a_sk, b_sn, b_sk,
                                                                            c_sm, c_sn, alpha);
  }
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

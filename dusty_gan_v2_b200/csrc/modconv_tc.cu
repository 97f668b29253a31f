// a1: tcgen05 tensor-core implementation of the modulated 1x1 contraction (bf16 operands,
// fp32 accumulation in TMEM) -- forward.
//
//   Y[b, o, p] = act( sum_k wb[b, o, k] * X(b, k, p) + bias[o] ) * scale
//
// Mapping onto UMMA (cta_group::1, kind::f16):
//   M = 128 pixels        A tile = X_b[k0:k0+64, p0:p0+128]   pixel-contiguous  -> MN-major
//   N = BN out channels   B tile = wb_b[n0:n0+BN, k0:k0+64]   k-contiguous      -> K-major
//   K = 64 channels per pipeline stage = 4 x (UMMA_K = 16)
//   D = fp32 [128 lanes x BN columns] in TMEM
// Putting pixels on M keeps the instruction shape full (M = 128) even where O shrinks to 32
// at the 64x512 level, and lets NCHW activations be consumed as they are: the A operand is
// a "transposed" (MN-major) shared-memory tile, which TMA writes directly with the 128-byte
// swizzle the UMMA descriptor expects -- two boxes of [64 channels x 64 pixels] per stage.
// The K axis is the concatenation [features | Fourier features]; each 64-channel block
// comes from one of two tensor maps, and the Fourier map may be batch-shared (L2 resident).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
// (one elected thread), warps 2..5 = epilogue (tcgen05.ld 32x32b: one pixel row per thread,
// bias + leaky-ReLU fused, coalesced stores along the pixel axis).
#include "tc_common.cuh"

namespace dusty {

constexpr int kBM = 128;          // pixels per tile (UMMA M)
constexpr int kBK = 64;           // channels per stage (one 128-byte swizzle atom of bf16)
constexpr int kABytes = kBM * kBK * 2;   // 16 KiB
constexpr int kThreads = 192;

struct TcParams {
  int MT, NT, total_tiles, tiles_per_cta;   // tile schedule (pixel tiles, channel tiles)
  int O, C1, K, B2;       // K = C1 + C2 rounded up to kBK by TMA zero fill
  int64_t P;
  const float *bias;
  __nv_bfloat16 *y;
  int act;
  float alpha, scale;
};

// Multi-tile ("persistent") kernel.  Each CTA owns a contiguous range of output tiles
// (128 pixels x BN channels); the TMA producer streams the k-blocks of successive tiles
// through one shared-memory ring without draining it between tiles, the MMA thread
// alternates between TWO TMEM accumulators, and the epilogue warps drain accumulator i
// while the MMAs of tile i+1 are already running.  This matters most where the layer is
// thin (64x512 level: K = 32..64 per tile, i.e. one or two k-blocks): a one-tile CTA there
// is all prologue and epilogue.
//
// B_MN == false: forward (B = wb tile [BN out-channels x 64 k], K-major)
// B_MN == true : dX      (A = dY, contraction over out-channels; B = wb tile
//                         [64 out-channels x BN in-channels], in-channel contiguous = MN-major)
template <int BN, int STAGES, bool B_MN>
__global__ void __launch_bounds__(kThreads)
modconv_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_x1,
                      const __grid_constant__ CUtensorMap map_x2,
                      const __grid_constant__ CUtensorMap map_w, TcParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kBBytes = BN * kBK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  uint8_t *a_base = smem;
  uint8_t *b_base = smem + STAGES * kABytes;
  uint64_t *full = (uint64_t *)(smem + STAGES * kStageBytes);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;      // [2]
  uint64_t *acc_empty = acc_full + 2;       // [2]
  uint32_t *tmem_slot = (uint32_t *)(acc_empty + 2);
  uint8_t *epi_base = smem + STAGES * kStageBytes + 128;   // 2 x 8 KiB staging, 16-B aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (prm.K + kBK - 1) / kBK;
  const int t_begin = blockIdx.x * prm.tiles_per_cta;
  const int t_end = min(t_begin + prm.tiles_per_cta, prm.total_tiles);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);          // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (sample b, channel tile n0, pixel tile p0); pixel tiles of a sample are adjacent
  auto decode = [&](int tile, int &b, int &n0, int &p0) {
    const int mt = tile % prm.MT;
    const int r = tile / prm.MT;
    p0 = mt * kBM;
    n0 = (r % prm.NT) * BN;
    b = r / prm.NT;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        int b, n0, p0;
        decode(tile, b, n0, p0);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], kStageBytes);
          const int c0 = kb * kBK;
          uint8_t *a_dst = a_base + s * kABytes;
          if (c0 < prm.C1) {
            tma_load_3d(a_dst, &map_x1, &full[s], p0, c0, b);
            tma_load_3d(a_dst + kABytes / 2, &map_x1, &full[s], p0 + 64, c0, b);
          } else {
            const int bb = prm.B2 == 1 ? 0 : b;
            tma_load_3d(a_dst, &map_x2, &full[s], p0, c0 - prm.C1, bb);
            tma_load_3d(a_dst + kABytes / 2, &map_x2, &full[s], p0 + 64, c0 - prm.C1, bb);
          }
          if (!B_MN) {
            tma_load_3d(b_base + s * kBBytes, &map_w, &full[s], c0, n0, b);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_3d(b_base + s * kBBytes + j * 8192, &map_w, &full[s], n0 + 64 * j, c0, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN, true, B_MN);
      // A (MN-major, SW128): 16 channel rows = 2 KiB per UMMA_K step; LBO = next 64-pixel block
      // (8 KiB), SBO = next group of 8 channel rows (1 KiB).  B K-major (SW128): 32 bytes per
      // UMMA_K step inside the swizzle atom, SBO = next group of 8 out-channel rows; B MN-major
      // (dX): same geometry as A.  High words are constant, low words advance by adds.
      const uint32_t d_hi = desc_hi(1024, 2);
      const uint32_t a_lo0 = desc_lo(smem_u32(a_base), kABytes / 2);
      const uint32_t b_lo0 = desc_lo(smem_u32(b_base), B_MN ? 8192 : 16);
      constexpr uint32_t kBStep = (B_MN ? 2048 : 32) >> 4;
      int it = 0, lt = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
        const int a = lt & 1;
        mbar_wait(&acc_empty[a], ((lt >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)s * (kABytes >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)s * (kBBytes >> 4);
#pragma unroll
          for (int k16 = 0; k16 < kBK / 16; ++k16)
            umma_bf16_lh(tmem_acc, a_lo + k16 * (2048 >> 4), d_hi, b_lo + k16 * kBStep, d_hi, idesc,
                         (kb > 0 || k16 > 0) ? 1u : 0u);
          umma_commit(&empty[s]);            // smem slot reusable once these MMAs retire
        }
        umma_commit(&acc_full[a]);           // accumulator of this tile complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM -> registers (one pixel row per thread) -> bias/act -> bf16 -> shared-memory
    // staging tile [32 channels][128 pixels] -> 16-byte global stores along the pixel axis
    // (a direct store from the accumulator layout would be 2 bytes per lane).
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;           // pixel row inside the tile
    const int et = threadIdx.x - 64;         // 0..127 among the epilogue threads
    __nv_bfloat16 *stage = reinterpret_cast<__nv_bfloat16 *>(epi_base);   // 2 x [32][128]
    int lt = 0, chunk_id = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
      int b, n0, p0;
      decode(tile, b, n0, p0);
      const int a = lt & 1;
      mbar_wait(&acc_full[a], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN) + ((uint32_t)(q * 32) << 16);
      __nv_bfloat16 *yt = prm.y + (int64_t)b * prm.O * prm.P + p0;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32, ++chunk_id) {
        __nv_bfloat16 *buf = stage + (chunk_id & 1) * (32 * kBM);
        uint32_t r0[16], r1[16];
        tmem_ld16(tmem_acc + (uint32_t)c, r0);
        tmem_ld16(tmem_acc + (uint32_t)(c + 16), r1);
        tmem_ld_wait();
        if (c + 32 >= BN) {                  // accumulator fully read: hand it back to the MMA
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int o = n0 + c + j;
          float v = __uint_as_float(j < 16 ? r0[j] : r1[j - 16]);
          if (prm.bias && o < prm.O) v += __ldg(prm.bias + o);
          if (prm.act == 3) v = v > 0.f ? v : v * prm.alpha;
          buf[j * kBM + row] = __float2bfloat16_rn(v * prm.scale);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // 32 channel rows x 256 bytes = 512 vectors of 16 bytes, 4 per thread
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int v = i * 128 + et;
          const int ch = v >> 4, seg = v & 15;
          const int o = n0 + c + ch;
          if (o < prm.O) {
            const uint4 val = *reinterpret_cast<const uint4 *>(buf + ch * kBM + seg * 8);
            *reinterpret_cast<uint4 *>(yt + (int64_t)o * prm.P + seg * 8) = val;
          }
        }
        // the other staging buffer is used next; this one is rewritten two chunks later,
        // after the next bar.sync has ordered these reads before those writes
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ dW kernel
// dwb[b, o, k] = sum_p dY[b, o, p] * X(b, k, p): both operands are pixel-contiguous, i.e.
// K-major for a contraction over pixels.  M = 128 input channels (two 64-channel boxes,
// each from the feature or the Fourier tensor map), N = BN out-channels, K = 64 pixels per
// stage; the accumulator row is the in-channel index, so the fp32 result is stored with the
// in-channel axis contiguous (coalesced), no atomics.
struct DwParams {
  int O, C1, K, B2;
  int64_t P;
  float *dw;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads)
modconv_dw_tc_kernel(const __grid_constant__ CUtensorMap map_x1,
                     const __grid_constant__ CUtensorMap map_x2,
                     const __grid_constant__ CUtensorMap map_g, DwParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kBBytes = BN * kBK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  uint8_t *a_base = smem;
  uint8_t *b_base = smem + STAGES * kABytes;
  uint64_t *full = (uint64_t *)(smem + STAGES * kStageBytes);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;
  uint32_t *tmem_slot = (uint32_t *)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM;       // in-channel tile
  const int n0 = blockIdx.y * BN;        // out-channel tile
  const int b = blockIdx.z;
  const int num_pb = (int)(prm.P / kBK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int bb = prm.B2 == 1 ? 0 : b;
      for (int pb = 0; pb < num_pb; ++pb) {
        const int s = pb % STAGES;
        const uint32_t ph = (pb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], kStageBytes);
        uint8_t *a_dst = a_base + s * kABytes;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c = m0 + 64 * half;
          if (c < prm.C1) tma_load_3d(a_dst + half * 8192, &map_x1, &full[s], pb * kBK, c, b);
          else tma_load_3d(a_dst + half * 8192, &map_x2, &full[s], pb * kBK, c - prm.C1, bb);
        }
        tma_load_3d(b_base + s * kBBytes, &map_g, &full[s], pb * kBK, n0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN, false, false);
      const uint32_t d_hi = desc_hi(1024, 2);
      const uint32_t a_lo0 = desc_lo(smem_u32(a_base), 16);
      const uint32_t b_lo0 = desc_lo(smem_u32(b_base), 16);
      for (int pb = 0; pb < num_pb; ++pb) {
        const int s = pb % STAGES;
        const uint32_t ph = (pb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)s * (kABytes >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)s * (kBBytes >> 4);
#pragma unroll
        for (int k16 = 0; k16 < kBK / 16; ++k16)
          umma_bf16_lh(tmem_acc, a_lo + k16 * 2, d_hi, b_lo + k16 * 2, d_hi, idesc,
                       (pb > 0 || k16 > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int k = m0 + q * 32 + lane;      // in-channel index of this thread's accumulator row
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float *dwb = prm.dw + (int64_t)b * prm.O * prm.K + k;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int o = n0 + c + j;
        if (o < prm.O && k < prm.K) dwb[(int64_t)o * prm.K] = __uint_as_float(r[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, BN);
  }
}

// 3-D bf16 tensor [d2, d1, d0] (d0 contiguous), box [1, box1, box0], 128-byte swizzle
static bool make_map3(CUtensorMap *m, const void *ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint32_t box0, uint32_t box1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int STAGES>
constexpr int tc_smem_bytes() {
  // stage ring + barrier block (128 B) + epilogue staging (2 x 32 x 128 bf16) + alignment slack
  return STAGES * (kABytes + BN * kBK * 2) + 128 + 2 * 32 * kBM * 2 + 1024;
}

bool modconv_fwd_tc_supported(int B, int O, int C1, int C2, int B2, int64_t P) {
  if (get_encode() == nullptr) return false;
  if (O < 32 || O % 16 != 0) return false;               // heads (O <= 4) use the streaming kernel
  if (C1 % kBK != 0 && C2 != 0) return false;            // a K block must not straddle the sources
  if (P % kBM != 0 || P > 0x7fffffff) return false;
  if ((C1 % 8) || (C2 % 8)) return false;                // 16-byte global strides for TMA
  (void)B; (void)B2;
  return true;
}

bool modconv_dx_tc_supported(int B, int O, int C1, int K, int64_t P) {
  if (get_encode() == nullptr) return false;
  if (O % 8 || K % 8 || C1 % 8 || C1 < 32) return false;
  if (P % kBM != 0 || P > 0x7fffffff) return false;
  (void)B;
  return true;
}

bool modconv_dw_tc_supported(int B, int O, int C1, int C2, int B2, int64_t P) {
  if (get_encode() == nullptr) return false;
  if (O < 32 || O % 16 != 0) return false;
  if (C1 % kBK != 0 && C2 != 0) return false;            // 64-channel boxes must not straddle
  if (P % kBK != 0 || P > 0x7fffffff) return false;
  if ((C1 % 8) || (C2 % 8)) return false;
  (void)B; (void)B2;
  return true;
}

template <int BN, int STAGES, bool B_MN>
static int launch_tc(const CUtensorMap &mx1, const CUtensorMap &mx2, const CUtensorMap &mw,
                     TcParams prm, int B, cudaStream_t st) {
  constexpr int smem = tc_smem_bytes<BN, STAGES>();
  static bool configured = false;
  if (int rc = set_smem(modconv_fwd_tc_kernel<BN, STAGES, B_MN>, smem, &configured)) return rc;
  prm.MT = (int)(prm.P / kBM);
  prm.NT = (prm.O + BN - 1) / BN;
  const int64_t total = (int64_t)prm.MT * prm.NT * B;
  if (total > 0x7fffffff) {
    set_error("modconv_tc: too many tiles");
    return DUSTY_EUNSUPPORTED;
  }
  prm.total_tiles = (int)total;
  // CTAs that can be co-resident: 2 per SM unless the stage ring fills shared memory
  const int resident = num_sms() * ((2 * smem <= 225 * 1024 && 2 * BN * 2 <= 512) ? 2 : 1);
  int ctas = prm.total_tiles < resident ? prm.total_tiles : resident;
  prm.tiles_per_cta = (prm.total_tiles + ctas - 1) / ctas;
  ctas = (prm.total_tiles + prm.tiles_per_cta - 1) / prm.tiles_per_cta;
  modconv_fwd_tc_kernel<BN, STAGES, B_MN><<<ctas, kThreads, smem, st>>>(mx1, mx2, mw, prm);
  return 0;
}

int modconv_fwd_tc(const void *wb, const void *x1, const void *x2, const float *bias, void *y,
                   int B, int O, int C1, int C2, int B2, int64_t P, int act, float alpha,
                   float scale, cudaStream_t st) {
  const int K = C1 + C2;
  const int BN = O >= 256 ? 256 : (O >= 128 ? 128 : (O >= 64 ? 64 : 32));
  CUtensorMap mx1, mx2, mw;
  const bool ok1 = make_map3(&mx1, x1, (uint64_t)P, (uint64_t)(C1 ? C1 : C2), (uint64_t)(C1 ? B : B2),
                             64, kBK);
  const bool ok2 = make_map3(&mx2, x2, (uint64_t)P, (uint64_t)(C2 ? C2 : C1), (uint64_t)(C2 ? B2 : B),
                             64, kBK);
  const bool ok3 = make_map3(&mw, wb, (uint64_t)K, (uint64_t)O, (uint64_t)B, kBK, (uint32_t)BN);
  if (!(ok1 && ok2 && ok3)) {
    set_error("modconv_fwd_tc: cuTensorMapEncodeTiled failed");
    return DUSTY_ECUDA;
  }
  TcParams prm;
  prm.O = O; prm.C1 = C1; prm.K = K; prm.B2 = B2; prm.P = P; prm.bias = bias;
  prm.y = (__nv_bfloat16 *)y; prm.act = act; prm.alpha = alpha; prm.scale = scale;
  switch (BN) {
    case 256: return launch_tc<256, 4, false>(mx1, mx2, mw, prm, B, st);
    case 128: return launch_tc<128, 5, false>(mx1, mx2, mw, prm, B, st);
    case 64: return launch_tc<64, 3, false>(mx1, mx2, mw, prm, B, st);
    default: return launch_tc<32, 4, false>(mx1, mx2, mw, prm, B, st);
  }
}

// dX1[b, c, p] = sum_o wb[b, o, c] * dY[b, o, p], c < C1
int modconv_dx_tc(const void *wb, const void *dy, void *dx1, int B, int O, int C1, int K, int64_t P,
                  cudaStream_t st) {
  const int BN = C1 > 128 ? 256 : (C1 > 64 ? 128 : 64);
  CUtensorMap mg, mw;
  const bool ok1 = make_map3(&mg, dy, (uint64_t)P, (uint64_t)O, (uint64_t)B, 64, kBK);
  // wb viewed with the in-channel axis innermost: box = [64 out-channels x 64 in-channels]
  const bool ok2 = make_map3(&mw, wb, (uint64_t)K, (uint64_t)O, (uint64_t)B, 64, kBK);
  if (!(ok1 && ok2)) {
    set_error("modconv_dx_tc: cuTensorMapEncodeTiled failed");
    return DUSTY_ECUDA;
  }
  TcParams prm;
  prm.O = C1;            // output channels of this contraction = input channels of the layer
  prm.C1 = O;            // the whole contraction axis (out-channels) comes from the dY map
  prm.K = O; prm.B2 = B; prm.P = P; prm.bias = nullptr;
  prm.y = (__nv_bfloat16 *)dx1; prm.act = 1; prm.alpha = 0.f; prm.scale = 1.f;
  switch (BN) {
    case 256: return launch_tc<256, 4, true>(mg, mg, mw, prm, B, st);
    case 128: return launch_tc<128, 5, true>(mg, mg, mw, prm, B, st);
    default: return launch_tc<64, 3, true>(mg, mg, mw, prm, B, st);
  }
}

template <int BN, int STAGES>
static int launch_dw(const CUtensorMap &mx1, const CUtensorMap &mx2, const CUtensorMap &mg,
                     const DwParams &prm, int B, cudaStream_t st) {
  constexpr int smem = STAGES * (kABytes + BN * kBK * 2) + 128 + 1024;   // no epilogue staging
  static bool configured = false;
  if (int rc = set_smem(modconv_dw_tc_kernel<BN, STAGES>, smem, &configured)) return rc;
  dim3 grid((unsigned)((prm.K + kBM - 1) / kBM), (unsigned)((prm.O + BN - 1) / BN), (unsigned)B);
  modconv_dw_tc_kernel<BN, STAGES><<<grid, kThreads, smem, st>>>(mx1, mx2, mg, prm);
  return 0;
}

int modconv_dw_tc(const void *dy, const void *x1, const void *x2, float *dwb, int B, int O, int C1,
                  int C2, int B2, int64_t P, cudaStream_t st) {
  const int BN = O >= 256 ? 256 : (O >= 128 ? 128 : (O >= 64 ? 64 : 32));
  CUtensorMap mx1, mx2, mg;
  const bool ok1 = make_map3(&mx1, x1, (uint64_t)P, (uint64_t)(C1 ? C1 : C2), (uint64_t)(C1 ? B : B2),
                             kBK, 64);
  const bool ok2 = make_map3(&mx2, x2, (uint64_t)P, (uint64_t)(C2 ? C2 : C1), (uint64_t)(C2 ? B2 : B),
                             kBK, 64);
  const bool ok3 = make_map3(&mg, dy, (uint64_t)P, (uint64_t)O, (uint64_t)B, kBK, (uint32_t)BN);
  if (!(ok1 && ok2 && ok3)) {
    set_error("modconv_dw_tc: cuTensorMapEncodeTiled failed");
    return DUSTY_ECUDA;
  }
  DwParams prm;
  prm.O = O; prm.C1 = C1; prm.K = C1 + C2; prm.B2 = B2; prm.P = P; prm.dw = dwb;
  switch (BN) {
    case 256: return launch_dw<256, 4>(mx1, mx2, mg, prm, B, st);
    case 128: return launch_dw<128, 4>(mx1, mx2, mg, prm, B, st);
    case 64: return launch_dw<64, 4>(mx1, mx2, mg, prm, B, st);
    default: return launch_dw<32, 5>(mx1, mx2, mg, prm, B, st);
  }
}

}  // namespace dusty

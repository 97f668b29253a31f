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
#include <stdlib.h>

#include "tc_common.cuh"

namespace dusty {

// DUSTY_TC_PREFETCH=<tiles> overrides the L2 prefetch distance (0 disables); experiments only
static int g_tc_prefetch = [] {
  const char *e = getenv("DUSTY_TC_PREFETCH");
  return e ? atoi(e) : 0;   // default off: measured slower (L4 conv1 102 -> 132 us)
}();

constexpr int kBM = 128;          // pixels per tile (UMMA M)
constexpr int kBK = 64;           // channels per stage (one 128-byte swizzle atom of bf16)
constexpr int kABytes = kBM * kBK * 2;   // 16 KiB
constexpr int kThreads = 192;

struct TcParams {
  int MT, NT, total_tiles, tiles_per_cta;   // tile schedule (pixel tiles, channel tiles)
  int pf, pf_all;         // L2 prefetch distance in tiles (0: off); pf_all: also for n-tiles > 0
  int O, C1, K, B2;       // K = C1 + C2 rounded up to kBK by TMA zero fill
  int64_t P;
  const float *bias;
  __nv_bfloat16 *y;
  int act;
  float alpha, scale;
  int out_f32;            // 1: fp32 output (split-bf16 operands of the fp32 mode, split3.cu)
  const float *ema;       // optional device scalar ema_var: accumulator * 1 / (sqrt(ema_var) + 1e-8)
  float *sumsq;           // optional: += sum of the squares of the STORED (bf16-rounded) outputs
};


// ------------------------------------------------------------------ epilogue (forward kernels)
// Eight epilogue warps in two groups of four (one warp per TMEM lane quarter in each group);
// the groups take alternate 32-column chunks of the accumulators, so two chunks are always in
// flight per SM and the latency chain of one (tcgen05.ld -> bias/act -> shared-memory
// transpose -> store) hides behind the other.  A chunk leaves as ONE TMA store of a dense
// [32 channels][128 pixels] bf16 tile issued by one thread: the warps never touch global
// memory, rows past the channel count are clipped by the tensor map.  (The first version --
// four warps, per-element __ldg of the bias, generic ld/st through the staging tile and
// 16-byte global stores by every thread -- took ~2000 cycles per chunk and left the MMA
// thread waiting on acc_empty: ncu source page, profiles/r01_ncu_modconv_epilogue.txt.)
constexpr int kFwdThreads = 64 + 256;
constexpr int kBiasFloats = 1024;         // bias copy in shared memory (zero padded)
constexpr int kEpiBytes = 2 * 32 * kBM * 2 + kBiasFloats * 4;   // 2 staging tiles + bias

struct EpiThread {
  int g, q, lane, row;
  bool issuer;
  uint32_t stage_u32;
  const float *bias_s;
  float pre;              // accumulator factor ahead of the bias: ModConv2d's EMA normaliser
};

// ema: optional device scalar (ModConv2d.ema_var, style.py:99-103).  The reference divides the
// per-sample weights by sqrt(ema_var) + 1e-8; applying that scalar to the accumulator instead
// makes the weights independent of the activation statistics, so all layers' weights can be
// prepared ahead of (and concurrently with) the activation chain.
__device__ __forceinline__ EpiThread epi_setup(uint8_t *epi_base, const float *bias, int n_bias,
                                               const float *ema = nullptr) {
  EpiThread e;
  e.pre = ema ? 1.f / (sqrtf(__ldg(ema)) + 1e-8f) : 1.f;
  const int warp = threadIdx.x >> 5;
  e.lane = threadIdx.x & 31;
  e.g = (warp - 2) >> 2;
  e.q = warp & 3;
  e.row = e.q * 32 + e.lane;
  const int et = threadIdx.x - 64;                      // 0..255 among the epilogue threads
  e.issuer = (et & 127) == 0;
  e.stage_u32 = smem_u32(epi_base) + (uint32_t)e.g * (32 * kBM * 2);
  float *bs = reinterpret_cast<float *>(epi_base + 2 * 32 * kBM * 2);
  for (int i = et; i < kBiasFloats; i += 256) bs[i] = (bias != nullptr && i < n_bias) ? __ldg(bias + i) : 0.f;
  named_bar_sync(3, 256);
  e.bias_s = bs;
  return e;
}

// one atomic per epilogue warp and CTA (persistent CTAs: a few thousand adds per launch)
__device__ __forceinline__ void epi_flush_sumsq(float ss, float *out, int lane) {
  if (out == nullptr) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) atomicAdd(out, ss);
}

// Drains the 32-column chunks of one accumulator that belong to this thread's group.
// gc: running chunk counter (same sequence in every epilogue thread); bias_at(c): index of the
// bias entry of accumulator column c; (p0, row0, b): TMA coordinates of the tile's first
// output row.
template <typename BiasAt>
__device__ __forceinline__ void epi_drain_tile(const EpiThread &e, uint32_t tmem_acc, int BN, int &gc,
                                               uint64_t *acc_empty_bar, BiasAt bias_at, int act,
                                               float alpha, float scale, const CUtensorMap *map_y,
                                               int p0, int row0, int b, bool out_f32, float &ss) {
  const int nch = BN >> 5;
  int my_last = nch - 1;                               // last chunk of this tile this group reads
  if (((gc + my_last) & 1) != e.g) --my_last;
  if (my_last < 0) {                                   // nothing to read: hand the accumulator back
    tc_fence_before();
    __syncwarp();
    if (e.lane == 0) mbar_arrive(acc_empty_bar);
  }
#pragma unroll 1
  for (int ci = 0; ci < nch; ++ci) {
    if (((gc + ci) & 1) != e.g) continue;
    const int c = ci << 5;
    uint32_t r0[16], r1[16];
    tmem_ld16(tmem_acc + (uint32_t)c, r0);
    tmem_ld16(tmem_acc + (uint32_t)(c + 16), r1);
    tmem_ld_wait();
    if (ci == my_last) {
      tc_fence_before();
      __syncwarp();
      if (e.lane == 0) mbar_arrive(acc_empty_bar);
    }
    const float *bp = e.bias_s + bias_at(c);
    if (out_f32) {
      // fp32 output: the same 8 KiB staging tile holds [16 channels][128 pixels] fp32, so a
      // 32-column chunk leaves as two TMA stores
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (e.issuer) tma_store_wait_read<0>();
        named_bar_sync(1 + e.g, 128);
        const uint32_t dst = e.stage_u32 + (uint32_t)e.row * 4;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v = fmaf(__uint_as_float(hf == 0 ? r0[j] : r1[j]), e.pre, bp[hf * 16 + j]);
          if (act == 3) v = v > 0.f ? v : v * alpha;
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + (uint32_t)j * (kBM * 4)), "r"(__float_as_uint(v * scale))
                       : "memory");
        }
        fence_proxy_async();
        named_bar_sync(1 + e.g, 128);
        if (e.issuer) {
          tma_store_3d(map_y, e.stage_u32, p0, row0 + c + hf * 16, b);
          tma_store_commit();
        }
      }
      continue;
    }
    uint32_t h[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float v0 = fmaf(__uint_as_float(j < 8 ? r0[2 * j] : r1[2 * j - 16]), e.pre, bp[2 * j]);
      float v1 = fmaf(__uint_as_float(j < 8 ? r0[2 * j + 1] : r1[2 * j - 15]), e.pre, bp[2 * j + 1]);
      if (act == 3) {
        v0 = v0 > 0.f ? v0 : v0 * alpha;
        v1 = v1 > 0.f ? v1 : v1 * alpha;
      }
      const __nv_bfloat162 pk = __floats2bfloat162_rn(v0 * scale, v1 * scale);
      h[j] = *reinterpret_cast<const uint32_t *>(&pk);
      // statistic of the tensor as the next layer will read it (ModConv2d's ema_var, style.py:99-102)
      const float2 f = __bfloat1622float2(pk);
      ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
    }
    if (e.issuer) tma_store_wait_read<0>();            // previous store has left the staging tile
    named_bar_sync(1 + e.g, 128);
    const uint32_t dst = e.stage_u32 + (uint32_t)e.row * 2;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      st_shared_b16(dst + (uint32_t)(2 * j) * (kBM * 2), (uint16_t)(h[j] & 0xffff));
      st_shared_b16(dst + (uint32_t)(2 * j + 1) * (kBM * 2), (uint16_t)(h[j] >> 16));
    }
    fence_proxy_async();
    named_bar_sync(1 + e.g, 128);
    if (e.issuer) {
      tma_store_3d(map_y, e.stage_u32, p0, row0 + c, b);
      tma_store_commit();
    }
  }
  gc += nch;
}

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
__global__ void __launch_bounds__(kFwdThreads)
modconv_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_x1,
                      const __grid_constant__ CUtensorMap map_x2,
                      const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_y, TcParams prm) {
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

  const int warp = threadIdx.x >> 5;
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
      mbar_init(&acc_empty[a], 8);          // one arrival per epilogue warp
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
    // ===================== TMA producer (warp-uniform loop, elected lane issues) ============
    RingPos r;
    for (int tile = t_begin; tile < t_end; ++tile) {
      int b, n0, p0;
      decode(tile, b, n0, p0);
      if (prm.pf > 0 && tile + prm.pf < t_end) {      // a later tile's HBM-resident operands -> L2
        int bn, n0n, p0n;
        decode(tile + prm.pf, bn, n0n, p0n);
        if ((n0n == 0 || prm.pf_all) && elect_one_sync())
          for (int kb = 0; kb < num_kb; ++kb) {
            const int c0 = kb * kBK;
            if (c0 < prm.C1) {
              tma_prefetch_3d(&map_x1, p0n, c0, bn);
              tma_prefetch_3d(&map_x1, p0n + 64, c0, bn);
            } else if (prm.B2 != 1) {
              tma_prefetch_3d(&map_x2, p0n, c0 - prm.C1, bn);
              tma_prefetch_3d(&map_x2, p0n + 64, c0 - prm.C1, bn);
            }
          }
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = r.s;
        mbar_wait(&empty[s], r.ph ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&full[s], kStageBytes);
          const int c0 = kb * kBK;
          uint8_t *a_dst = a_base + s * kABytes;
          if (c0 < prm.C1) {
            tma_load_3d_hint(a_dst, &map_x1, &full[s], p0, c0, b, kEvictFirst);
            tma_load_3d_hint(a_dst + kABytes / 2, &map_x1, &full[s], p0 + 64, c0, b, kEvictFirst);
          } else {
            const int bb = prm.B2 == 1 ? 0 : b;
            const uint64_t pol = prm.B2 == 1 ? kEvictLast : kEvictFirst;
            tma_load_3d_hint(a_dst, &map_x2, &full[s], p0, c0 - prm.C1, bb, pol);
            tma_load_3d_hint(a_dst + kABytes / 2, &map_x2, &full[s], p0 + 64, c0 - prm.C1, bb, pol);
          }
          if (!B_MN) {
            tma_load_3d(b_base + s * kBBytes, &map_w, &full[s], c0, n0, b);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_3d(b_base + s * kBBytes + j * 8192, &map_w, &full[s], n0 + 64 * j, c0, b);
          }
        }
        __syncwarp();
        r.template advance<STAGES>();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elected lane issues) =============
    constexpr uint32_t idesc = make_idesc(kBM, BN, true, B_MN);
    // A (MN-major, SW128): 16 channel rows = 2 KiB per UMMA_K step; LBO = next 64-pixel block
    // (8 KiB), SBO = next group of 8 channel rows (1 KiB).  B K-major (SW128): 32 bytes per
    // UMMA_K step inside the swizzle atom, SBO = next group of 8 out-channel rows; B MN-major
    // (dX): same geometry as A.  High words are constant, low words advance by adds.
    const uint32_t d_hi = desc_hi(1024, 2);
    const uint32_t a_lo0 = desc_lo(smem_u32(a_base), kABytes / 2);
    const uint32_t b_lo0 = desc_lo(smem_u32(b_base), B_MN ? 8192 : 16);
    constexpr uint32_t kBStep = (B_MN ? 2048 : 32) >> 4;
    RingPos r;
    int lt = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
      const int a = lt & 1;
      mbar_wait(&acc_empty[a], ((lt >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = r.s;
        mbar_wait(&full[s], r.ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)s * (kABytes >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)s * (kBBytes >> 4);
#pragma unroll
          for (int k16 = 0; k16 < kBK / 16; ++k16)
            umma_bf16_lh(tmem_acc, a_lo + k16 * (2048 >> 4), d_hi, b_lo + k16 * kBStep, d_hi, idesc,
                         (kb > 0 || k16 > 0) ? 1u : 0u);
          umma_commit(&empty[s]);            // smem slot reusable once these MMAs retire
        }
        __syncwarp();
        r.template advance<STAGES>();
      }
      if (elect_one_sync()) umma_commit(&acc_full[a]);   // accumulator of this tile complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const EpiThread e = epi_setup(epi_base, prm.bias, prm.O, prm.ema);
    int lt = 0, gc = 0;
    float ss = 0.f;
    for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
      int b, n0, p0;
      decode(tile, b, n0, p0);
      const int a = lt & 1;
      mbar_wait(&acc_full[a], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN) + ((uint32_t)(e.q * 32) << 16);
      epi_drain_tile(e, tmem_acc, BN, gc, &acc_empty[a], [&](int c) { return n0 + c; }, prm.act,
                     prm.alpha, prm.scale, &map_y, p0, n0, b, prm.out_f32 != 0, ss);
    }
    if (e.issuer) tma_store_wait_read<0>();
    epi_flush_sumsq(ss, prm.sumsq, e.lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ batch-fused forward
// Batch-shared Fourier block (B2 == 1): the Fourier half of the contraction does not depend
// on the sample on its A side,
//   Y[(b,o), p] = sum_{k<C1} wb[b,o,k] X1[b,k,p]  +  sum_{k<C2} wb[(b,o), C1+k] PE[k,p],
// so its second term is ONE dense GEMM  [P x C2] . [C2 x (B*O)]  over the weight matrix
// viewed as [(B*O), K].  A tile is 128 pixels x BN = NS*O columns (NS samples side by side,
// BN <= 256): the Fourier k-blocks run as full-width 128 x BN x 16 UMMAs (the A tile is read
// from shared memory once for NS samples instead of once per sample, and the instruction is
// tensor-bound instead of shared-memory-bound at O = 32), while the per-sample feature
// k-blocks are 128 x O x 16 UMMAs into the column slice [j*O, (j+1)*O) of the same
// accumulator.  The two kinds of stage are interleaved in the ring (P X P X ...) so that the
// producer's look-ahead always spans a long Fourier stage; the very first UMMA of a tile is a
// Fourier one and overwrites all BN columns, everything after accumulates.
struct ShParams {
  int MT, NT, total_tiles, tiles_per_cta;
  int O, C1, C2, NS, BN, pf;
  int64_t P;
  const float *bias;
  __nv_bfloat16 *y;
  int act;
  float alpha, scale;
  const float *ema;       // as TcParams::ema
  float *sumsq;           // as TcParams::sumsq
};

constexpr int kShBBytes = 256 * kBK * 2;          // B slot: up to 256 weight rows x 64 k
constexpr int kShStages = 4;

__global__ void __launch_bounds__(kFwdThreads)
modconv_fwd_shared_tc_kernel(const __grid_constant__ CUtensorMap map_x1,
                             const __grid_constant__ CUtensorMap map_pe,
                             const __grid_constant__ CUtensorMap map_ws,
                             const __grid_constant__ CUtensorMap map_wp,
                             const __grid_constant__ CUtensorMap map_y, ShParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int STAGES = kShStages;
  constexpr int kStageBytes = kABytes + kShBBytes;
  uint8_t *a_base = smem;
  uint8_t *b_base = smem + STAGES * kABytes;
  uint64_t *full = (uint64_t *)(smem + STAGES * kStageBytes);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;      // [2]
  uint64_t *acc_empty = acc_full + 2;       // [2]
  uint32_t *tmem_slot = (uint32_t *)(acc_empty + 2);
  uint8_t *epi_base = smem + STAGES * kStageBytes + 128;

  const int warp = threadIdx.x >> 5;
  const int kb1 = prm.C1 / kBK;             // feature k-blocks per sample
  const int nX = prm.NS * kb1;              // per-sample stages of a tile
  const int nP = prm.C2 / kBK;              // Fourier stages of a tile
  // two (sample, k-block) units share one ring slot when their weight tiles are <= 8 KiB each:
  // slot = [A0 16K | B0 8K | B1 8K | A1 16K], so the ring carries ~45 KiB of its 48 per slot
  const int pack = (prm.O <= 64 && (nX & 1) == 0) ? 2 : 1;
  const int nXs = nX / pack;
  const int t_begin = blockIdx.x * prm.tiles_per_cta;
  const int t_end = min(t_begin + prm.tiles_per_cta, prm.total_tiles);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, elected lane issues) ============
    RingPos r;
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int p0 = (tile % prm.MT) * kBM;
      const int nt = tile / prm.MT;
      if (prm.pf && tile + 1 < t_end && elect_one_sync()) {   // next tile's activations -> L2
        const int p0n = ((tile + 1) % prm.MT) * kBM, ntn = (tile + 1) / prm.MT;
        for (int j = 0; j < prm.NS; ++j)
          for (int kbx = 0; kbx < kb1; ++kbx) {
            tma_prefetch_3d(&map_x1, p0n, kbx * kBK, ntn * prm.NS + j);
            tma_prefetch_3d(&map_x1, p0n + 64, kbx * kBK, ntn * prm.NS + j);
          }
      }
      int j = 0, kbx = 0, acc = 0;              // next per-sample unit; interleave accumulator
      for (int pi = 0; pi < nP; ++pi) {
        {
          const int s = r.s;
          mbar_wait(&empty[s], r.ph ^ 1);
          if (elect_one_sync()) {
            mbar_expect_tx(&full[s], kABytes + prm.BN * kBK * 2);
            uint8_t *a_dst = a_base + s * kABytes;
            tma_load_3d_hint(a_dst, &map_pe, &full[s], p0, pi * kBK, 0, kEvictLast);
            tma_load_3d_hint(a_dst + kABytes / 2, &map_pe, &full[s], p0 + 64, pi * kBK, 0, kEvictLast);
            tma_load_3d_hint(b_base + s * kShBBytes, &map_wp, &full[s], prm.C1 + pi * kBK, nt * prm.BN, 0,
                             kEvictLast);
          }
          __syncwarp();
          r.advance<STAGES>();
        }
        for (acc += nXs; acc >= nP; acc -= nP) {
          const int s = r.s;
          mbar_wait(&empty[s], r.ph ^ 1);
          const bool leader = elect_one_sync();
          if (leader) mbar_expect_tx(&full[s], pack * (kABytes + prm.O * kBK * 2));
          for (int u = 0; u < pack; ++u) {
            if (leader) {
              const int b = nt * prm.NS + j, c0 = kbx * kBK;
              uint8_t *a_dst = u ? b_base + s * kShBBytes + 16384 : a_base + s * kABytes;
              tma_load_3d_hint(a_dst, &map_x1, &full[s], p0, c0, b, kEvictFirst);
              tma_load_3d_hint(a_dst + kABytes / 2, &map_x1, &full[s], p0 + 64, c0, b, kEvictFirst);
              tma_load_3d_hint(b_base + s * kShBBytes + u * 8192, &map_ws, &full[s], c0, 0, b, kEvictLast);
            }
            if (++kbx == kb1) { kbx = 0; ++j; }
          }
          __syncwarp();
          r.advance<STAGES>();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elected lane issues) =============
    const uint32_t idesc_p = make_idesc(kBM, prm.BN, true, false);
    const uint32_t idesc_x = make_idesc(kBM, prm.O, true, false);
    const uint32_t d_hi = desc_hi(1024, 2);
    const uint32_t a_lo0 = desc_lo(smem_u32(a_base), kABytes / 2);
    const uint32_t b_lo0 = desc_lo(smem_u32(b_base), 16);
    // second unit of a packed slot: its A tile sits in the upper half of the weight slot
    const uint32_t a1_lo0 = a_lo0 + (uint32_t)((b_base - a_base) >> 4) + (16384 >> 4);
    RingPos r;
    int lt = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
      const int a = lt & 1;
      mbar_wait(&acc_empty[a], ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(a * 256);
      int j = 0, kbx = 0, acc = 0;
      for (int pi = 0; pi < nP; ++pi) {
        {
          const int s = r.s;
          mbar_wait(&full[s], r.ph);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)s * (kABytes >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)s * (kShBBytes >> 4);
#pragma unroll
            for (int k16 = 0; k16 < kBK / 16; ++k16)
              umma_bf16_lh(tmem_acc, a_lo + k16 * (2048 >> 4), d_hi, b_lo + k16 * 2, d_hi, idesc_p,
                           (pi > 0 || k16 > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
          }
          __syncwarp();
          r.advance<STAGES>();
        }
        for (acc += nXs; acc >= nP; acc -= nP) {
          const int s = r.s;
          mbar_wait(&full[s], r.ph);
          tc_fence_after();
          const bool leader = elect_one_sync();
          for (int u = 0; u < pack; ++u) {
            if (leader) {
              const uint32_t a_lo = u ? a1_lo0 + (uint32_t)s * (kShBBytes >> 4)
                                      : a_lo0 + (uint32_t)s * (kABytes >> 4);
              const uint32_t bu_lo = b_lo0 + (uint32_t)s * (kShBBytes >> 4) + (uint32_t)u * (8192 >> 4);
              const uint32_t tmem_x = tmem_acc + (uint32_t)(j * prm.O);
#pragma unroll
              for (int k16 = 0; k16 < kBK / 16; ++k16)
                umma_bf16_lh(tmem_x, a_lo + k16 * (2048 >> 4), d_hi, bu_lo + k16 * 2, d_hi, idesc_x, 1u);
            }
            if (++kbx == kb1) { kbx = 0; ++j; }
          }
          if (leader) umma_commit(&empty[s]);
          __syncwarp();
          r.advance<STAGES>();
        }
      }
      if (elect_one_sync()) umma_commit(&acc_full[a]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // y viewed as [(B*O), P]: accumulator column c of column tile nt is row nt*BN + c of that
    // matrix, its bias is bias[(nt*BN + c) % O] (O % 32 == 0: a chunk never straddles samples)
    const EpiThread e = epi_setup(epi_base, prm.bias, prm.O, prm.ema);
    int lt = 0, gc = 0;
    float ss = 0.f;
    for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
      const int p0 = (tile % prm.MT) * kBM;
      const int n0 = (tile / prm.MT) * prm.BN;
      const int a = lt & 1;
      mbar_wait(&acc_full[a], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(a * 256) + ((uint32_t)(e.q * 32) << 16);
      epi_drain_tile(e, tmem_acc, prm.BN, gc, &acc_empty[a], [&](int c) { return (n0 + c) % prm.O; },
                     prm.act, prm.alpha, prm.scale, &map_y, p0, n0, 0, false, ss);
    }
    if (e.issuer) tma_store_wait_read<0>();
    epi_flush_sumsq(ss, prm.sumsq, e.lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
  int64_t ld;            // row pitch of dw (floats): K, or the pitch of a wider [B, O, ld] tensor
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
    const int bb = prm.B2 == 1 ? 0 : b;
    RingPos r;
    for (int pb = 0; pb < num_pb; ++pb) {
      const int s = r.s;
      mbar_wait(&empty[s], r.ph ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&full[s], kStageBytes);
        uint8_t *a_dst = a_base + s * kABytes;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c = m0 + 64 * half;
          // stream-once activations must not evict the batch-shared Fourier block, which every
          // sample's CTAs re-read (without hints: 2.6 GB of DRAM reads for 0.44 GB algorithmic)
          if (c < prm.C1)
            tma_load_3d_hint(a_dst + half * 8192, &map_x1, &full[s], pb * kBK, c, b, kEvictFirst);
          else
            tma_load_3d_hint(a_dst + half * 8192, &map_x2, &full[s], pb * kBK, c - prm.C1, bb,
                             prm.B2 == 1 ? kEvictLast : kEvictFirst);
        }
        tma_load_3d_hint(b_base + s * kBBytes, &map_g, &full[s], pb * kBK, n0, b, kEvictFirst);
      }
      __syncwarp();
      r.template advance<STAGES>();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(kBM, BN, false, false);
    const uint32_t d_hi = desc_hi(1024, 2);
    const uint32_t a_lo0 = desc_lo(smem_u32(a_base), 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(b_base), 16);
    RingPos r;
    for (int pb = 0; pb < num_pb; ++pb) {
      const int s = r.s;
      mbar_wait(&full[s], r.ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t a_lo = a_lo0 + (uint32_t)s * (kABytes >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)s * (kBBytes >> 4);
#pragma unroll
        for (int k16 = 0; k16 < kBK / 16; ++k16)
          umma_bf16_lh(tmem_acc, a_lo + k16 * 2, d_hi, b_lo + k16 * 2, d_hi, idesc,
                       (pb > 0 || k16 > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      __syncwarp();
      r.template advance<STAGES>();
    }
    if (elect_one_sync()) umma_commit(acc_full);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int k = m0 + q * 32 + lane;      // in-channel index of this thread's accumulator row
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float *dwb = prm.dw + (int64_t)b * prm.O * prm.ld + k;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int o = n0 + c + j;
        if (o < prm.O && k < prm.K) dwb[(int64_t)o * prm.ld] = __uint_as_float(r[j]);
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

// same tensor, dense (un-swizzled) box: destination of the epilogue's TMA stores
static bool make_map3_plain(CUtensorMap *m, const void *ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                            uint32_t box0, uint32_t box1, bool f32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const uint64_t esz = f32 ? 4 : 2;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * esz, d0 * d1 * esz};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void *>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int STAGES>
constexpr int tc_smem_bytes() {
  // stage ring + barrier block (128 B) + epilogue staging / bias copy + alignment slack
  return STAGES * (kABytes + BN * kBK * 2) + 128 + kEpiBytes + 1024;
}

bool modconv_fwd_tc_supported(int B, int O, int C1, int C2, int B2, int64_t P) {
  if (get_encode() == nullptr) return false;
  if (O < 32 || O % 16 != 0 || O > kBiasFloats - 256) return false;   // heads (O <= 4): streaming kernel
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
                     const CUtensorMap &my, TcParams prm, int B, cudaStream_t st) {
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
  // thin tiles (one or two k-blocks) finish in ~1-2 us: look two tiles ahead there
  prm.pf = g_tc_prefetch < 0 ? (prm.K <= 2 * kBK ? 2 : 1) : g_tc_prefetch;
  prm.pf_all = 0;
  modconv_fwd_tc_kernel<BN, STAGES, B_MN><<<ctas, kFwdThreads, smem, st>>>(mx1, mx2, mw, my, prm);
  return 0;
}

// samples placed side by side in one accumulator by the batch-fused kernel (0: not applicable)
int modconv_shared_group(int B, int O, int C1, int C2, int B2, int64_t P) {
  if (B2 != 1 || C1 <= 0 || C2 <= 0 || B < 2) return 0;
  // O = 128 (two samples per tile) measured no faster than per-sample tiles: not used
  if (C1 % kBK || C2 % kBK || O % 32 || O > 64 || P % kBM) return 0;
  int ns = 256 / O;
  while (ns > 1 && B % ns) --ns;
  return ns >= 2 ? ns : 0;
}

static int modconv_fwd_shared_tc(const void *wb, const void *x1, const void *pe, const float *bias,
                                 void *y, int B, int O, int C1, int C2, int NS, int64_t P, int act,
                                 float alpha, float scale, cudaStream_t st, const float *ema, float *sumsq) {
  const int K = C1 + C2;
  const int BN = NS * O;
  CUtensorMap mx1, mpe, mws, mwp;
  const bool ok1 = make_map3(&mx1, x1, (uint64_t)P, (uint64_t)C1, (uint64_t)B, 64, kBK);
  const bool ok2 = make_map3(&mpe, pe, (uint64_t)P, (uint64_t)C2, 1, 64, kBK);
  const bool ok3 = make_map3(&mws, wb, (uint64_t)K, (uint64_t)O, (uint64_t)B, kBK, (uint32_t)O);
  const bool ok4 = make_map3(&mwp, wb, (uint64_t)K, (uint64_t)B * O, 1, kBK, (uint32_t)BN);
  CUtensorMap my;
  const bool ok5 = make_map3_plain(&my, y, (uint64_t)P, (uint64_t)B * O, 1, kBM, 32);
  if (!(ok1 && ok2 && ok3 && ok4 && ok5)) {
    set_error("modconv_fwd_shared_tc: cuTensorMapEncodeTiled failed");
    return DUSTY_ECUDA;
  }
  constexpr int smem = kShStages * (kABytes + kShBBytes) + 128 + kEpiBytes + 1024;
  static bool configured = false;
  if (int rc = set_smem(modconv_fwd_shared_tc_kernel, smem, &configured)) return rc;
  ShParams prm;
  prm.O = O; prm.C1 = C1; prm.C2 = C2; prm.NS = NS; prm.BN = BN; prm.P = P; prm.bias = bias;
  prm.y = (__nv_bfloat16 *)y; prm.act = act; prm.alpha = alpha; prm.scale = scale;
  prm.ema = ema;
  prm.sumsq = sumsq;
  prm.MT = (int)(P / kBM);
  prm.NT = B / NS;
  prm.pf = g_tc_prefetch != 0;
  const int64_t total = (int64_t)prm.MT * prm.NT;
  if (total > 0x7fffffff) {
    set_error("modconv_tc: too many tiles");
    return DUSTY_EUNSUPPORTED;
  }
  prm.total_tiles = (int)total;
  int ctas = prm.total_tiles < num_sms() ? prm.total_tiles : num_sms();   // 512 TMEM columns: 1 CTA / SM
  prm.tiles_per_cta = (prm.total_tiles + ctas - 1) / ctas;
  ctas = (prm.total_tiles + prm.tiles_per_cta - 1) / prm.tiles_per_cta;
  modconv_fwd_shared_tc_kernel<<<ctas, kFwdThreads, smem, st>>>(mx1, mpe, mws, mwp, my, prm);
  return 0;
}

int modconv_fwd_tc(const void *wb, const void *x1, const void *x2, const float *bias, void *y,
                   int B, int O, int C1, int C2, int B2, int64_t P, int act, float alpha,
                   float scale, bool batch_fused, cudaStream_t st, bool out_f32, const float *ema, float *sumsq) {
  const int K = C1 + C2;
  if (batch_fused && !out_f32) {
    const int ns = modconv_shared_group(B, O, C1, C2, B2, P);
    if (ns) return modconv_fwd_shared_tc(wb, x1, x2, bias, y, B, O, C1, C2, ns, P, act, alpha, scale, st, ema, sumsq);
  }
  const int BN = O >= 256 ? 256 : (O >= 128 ? 128 : (O >= 64 ? 64 : 32));
  CUtensorMap mx1, mx2, mw;
  const bool ok1 = make_map3(&mx1, x1, (uint64_t)P, (uint64_t)(C1 ? C1 : C2), (uint64_t)(C1 ? B : B2),
                             64, kBK);
  const bool ok2 = make_map3(&mx2, x2, (uint64_t)P, (uint64_t)(C2 ? C2 : C1), (uint64_t)(C2 ? B2 : B),
                             64, kBK);
  const bool ok3 = make_map3(&mw, wb, (uint64_t)K, (uint64_t)O, (uint64_t)B, kBK, (uint32_t)BN);
  CUtensorMap my;
  const bool ok4 = make_map3_plain(&my, y, (uint64_t)P, (uint64_t)O, (uint64_t)B, kBM, out_f32 ? 16 : 32, out_f32);
  if (!(ok1 && ok2 && ok3 && ok4)) {
    set_error("modconv_fwd_tc: cuTensorMapEncodeTiled failed");
    return DUSTY_ECUDA;
  }
  TcParams prm;
  prm.O = O; prm.C1 = C1; prm.K = K; prm.B2 = B2; prm.P = P; prm.bias = bias;
  prm.y = (__nv_bfloat16 *)y; prm.act = act; prm.alpha = alpha; prm.scale = scale;
  prm.out_f32 = out_f32 ? 1 : 0;
  prm.ema = ema;
  prm.sumsq = out_f32 ? nullptr : sumsq;
  switch (BN) {
    case 256: return launch_tc<256, 4, false>(mx1, mx2, mw, my, prm, B, st);
    case 128: return launch_tc<128, 5, false>(mx1, mx2, mw, my, prm, B, st);
    case 64: return launch_tc<64, 3, false>(mx1, mx2, mw, my, prm, B, st);
    default: return launch_tc<32, 4, false>(mx1, mx2, mw, my, prm, B, st);
  }
}

// dX1[b, c, p] = sum_o wb[b, o, c] * dY[b, o, p], c < C1
int modconv_dx_tc(const void *wb, const void *dy, void *dx1, int B, int O, int C1, int K, int64_t P,
                  cudaStream_t st, bool out_f32, const float *ema) {
  const int BN = C1 > 128 ? 256 : (C1 > 64 ? 128 : 64);
  CUtensorMap mg, mw;
  const bool ok1 = make_map3(&mg, dy, (uint64_t)P, (uint64_t)O, (uint64_t)B, 64, kBK);
  // wb viewed with the in-channel axis innermost: box = [64 out-channels x 64 in-channels]
  const bool ok2 = make_map3(&mw, wb, (uint64_t)K, (uint64_t)O, (uint64_t)B, 64, kBK);
  CUtensorMap my;
  const bool ok3 = make_map3_plain(&my, dx1, (uint64_t)P, (uint64_t)C1, (uint64_t)B, kBM, out_f32 ? 16 : 32, out_f32);
  if (!(ok1 && ok2 && ok3)) {
    set_error("modconv_dx_tc: cuTensorMapEncodeTiled failed");
    return DUSTY_ECUDA;
  }
  TcParams prm;
  prm.O = C1;            // output channels of this contraction = input channels of the layer
  prm.C1 = O;            // the whole contraction axis (out-channels) comes from the dY map
  prm.K = O; prm.B2 = B; prm.P = P; prm.bias = nullptr;
  prm.y = (__nv_bfloat16 *)dx1; prm.act = 1; prm.alpha = 0.f; prm.scale = 1.f;
  prm.out_f32 = out_f32 ? 1 : 0;
  prm.ema = ema;
  prm.sumsq = nullptr;
  switch (BN) {
    case 256: return launch_tc<256, 4, true>(mg, mg, mw, my, prm, B, st);
    case 128: return launch_tc<128, 5, true>(mg, mg, mw, my, prm, B, st);
    default: return launch_tc<64, 3, true>(mg, mg, mw, my, prm, B, st);
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
                  int C2, int B2, int64_t P, cudaStream_t st, int64_t ld) {
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
  prm.ld = ld > 0 ? ld : C1 + C2;
  switch (BN) {
    case 256: return launch_dw<256, 4>(mx1, mx2, mg, prm, B, st);
    case 128: return launch_dw<128, 4>(mx1, mx2, mg, prm, B, st);
    case 64: return launch_dw<64, 4>(mx1, mx2, mg, prm, B, st);
    default: return launch_dw<32, 5>(mx1, mx2, mg, prm, B, st);
  }
}

}  // namespace dusty

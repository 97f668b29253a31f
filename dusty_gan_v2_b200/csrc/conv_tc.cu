// a11: dense 2-D convolutions of the discriminator trunk on tcgen05 (bf16 operands, fp32
// accumulation in TMEM), NHWC activations.  Implicit GEMM without an im2col buffer: the K axis
// of the contraction is walked as "groups" of 64-wide chunks, and every chunk is one TMA box
// of the activation tensor at shifted coordinates.
//
//   y[b, oh, ow, n] = sum_g sum_k  A_g[b, oh, ow, k] * wpk[g, n, k]
//
// Two ways of addressing A_g (chosen by the host):
//  * "window" (forward / wgrad of a valid convolution over an explicitly padded input): the S
//    taps of one filter row are CONTIGUOUS in NHWC memory -- x[b, h, w:w+S, :] is a run of
//    S*C elements -- so filter row r is ONE group with K_g = S*C "virtual channels", read
//    through a tensor map whose pixel strides overlap the window (stride C per pixel, the
//    convolution stride folded into the map strides).  3 groups for a 3x3 filter, each tap row
//    read once, no zero padding needed.
//  * "tap" (dgrad: correlation of dY with the flipped filter, zero padded): one group per
//    filter tap, K_g = channels of dY, coordinates shifted by the tap offset; out-of-range
//    rows/columns are zero-filled by TMA.  A strided dgrad is 4 such launches (one per output
//    parity class) writing through strided output pointers.
//
// UMMA mapping (cta_group::1, kind::f16): M = 128 output pixels (TH x TW patch of one sample),
// N = BN output channels, K = 64 per stage.  A tile = [128 pixels][64 k] K-major SW128 (what
// the TMA box [64, TW, TH, 1] lands), B tile = wpk[g, n0:n0+BN, k0:k0+64] K-major SW128.
// Same warp-specialised multi-tile pipeline as modconv_tc.cu (TMA warp, MMA warp, 4 epilogue
// warps, two TMEM accumulators).  The accumulator row is a pixel and its columns are the
// output channels, which are contiguous in NHWC: the epilogue stores 16-byte vectors straight
// from registers (bias + leaky-ReLU optional).
//
// wgrad: dwp[r, m, n] = sum_{b,oh,ow} x_window_r[b, oh, ow, m] * dY[b, oh, ow, n]; M = 128
// virtual channels (s, c) of filter row r, N = BN output channels, K = 64 pixels per stage,
// both operands MN-major (channel-contiguous).  Split over pixel ranges; partials in a
// caller-provided fp32 workspace, reduced by a second kernel.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dusty {

namespace {

constexpr int kCM = 128;                 // UMMA M
constexpr int kCK = 64;                  // K elements per stage
constexpr int kCABytes = kCM * kCK * 2;  // 16 KiB
constexpr int kCThreads = 64 + 256;   // TMA warp, MMA warp, 8 epilogue warps (fprop / dgrad)
constexpr int kMaxClasses = 4;            // output parity classes of a stride-2 data gradient
constexpr int kWgThreads = 192;        // wgrad: one epilogue pass per CTA, 4 warps
constexpr int kMaxGroups = 16;

// Role profiling (tools only; libdusty_b200_prof.so is built with -DDUSTY_ROLE_PROF): cycles the
// three warp roles spend waiting on each other, summed over CTAs.  slots: 0 producer waits for
// a free ring slot, 1 MMA warp waits for operands, 2 MMA warp waits for a free accumulator,
// 3 epilogue warp 2 waits for a finished accumulator, 4 lifetime of the producer warp, 5 of
// the MMA warp, 6 of epilogue warp 2, 7 CTAs.
#ifdef DUSTY_ROLE_PROF
__device__ unsigned long long s_role_prof[12];   // 8: MMA issue-loop cycles (halo), 9: blocks
// event trace of CTA 0 (halo kernel): (tag << 56) | (index << 40) | clock
__device__ unsigned long long s_trace[8192];
__device__ unsigned int s_trace_n;
// no atomics (an atomic with a return value stalls its thread for ~700 cycles): every role
// appends to its own quarter of the buffer with a private counter
__device__ __forceinline__ void trace_ev(unsigned role, unsigned &n, unsigned tag, unsigned idx) {
  if (blockIdx.x != 0 || n >= 2048) return;
  s_trace[role * 2048 + n++] = ((unsigned long long)tag << 56) | ((unsigned long long)(idx & 0xffff) << 40) |
                               ((unsigned long long)clock64() & 0xffffffffffull);
}
#define TRACE_DECL unsigned trace_n__ = 0
#define TRACE(role, tag, idx) trace_ev(role, trace_n__, tag, idx)
#define TRACE0(role, tag, idx) do { if ((threadIdx.x & 31) == 0) trace_ev(role, trace_n__, tag, idx); } while (0)
#define PROF_DECL long long prof_acc__[2] = {0, 0}; const long long prof_t0__ = clock64()
#define PROF_WAIT(i, stmt) do { const long long t__ = clock64(); stmt; prof_acc__[i] += clock64() - t__; } while (0)
#define PROF_FLUSH(slot_a, slot_b, slot_life)                                              \
  do {                                                                                     \
    if ((threadIdx.x & 31) == 0) {                                                         \
      atomicAdd(&s_role_prof[slot_a], (unsigned long long)prof_acc__[0]);                  \
      if ((slot_b) >= 0) atomicAdd(&s_role_prof[(slot_b) < 0 ? 0 : (slot_b)], (unsigned long long)prof_acc__[1]); \
      atomicAdd(&s_role_prof[slot_life], (unsigned long long)(clock64() - prof_t0__));     \
      if ((slot_life) == 4) atomicAdd(&s_role_prof[7], 1ull);                              \
    }                                                                                      \
  } while (0)
#else
#define PROF_DECL
#define PROF_WAIT(i, stmt) stmt
#define PROF_FLUSH(a, b, c)
#define TRACE_DECL
#define TRACE(role, tag, idx)
#define TRACE0(role, tag, idx)
#endif

struct ConvMaps {
  CUtensorMap a[4];
  CUtensorMap w;
  CUtensorMap y[kMaxClasses];              // output view of every class (TMA stores)
};

// A launch covers `ncls` classes (1 except for a strided data gradient, whose output parity
// classes each have their own taps, output origin and extent); a CTA walks tiles with the class
// as the fastest index, so every CTA sees the same mix of cheap and expensive classes.
struct ConvClass {
  int G;                                   // groups (taps / filter rows)
  int aw[kMaxGroups], ah[kMaxGroups];      // coordinate offset of group g's box
  int wtap[kMaxGroups];                    // index of group g's filter block in the weight tensor
  int H_out, W_out;                        // extent of this class's output view
  long long y_off;                         // element offset of its origin
};

struct ConvParams {
  int ncls, KC;                            // classes, 64-wide chunks per group
  ConvClass cls[kMaxClasses];
  int amap[kMaxGroups];                    // window mode: tensor map of group g (class 0 only)
  int TW, TH, tiles_w, tiles_h, NT, total_tiles, tiles_per_cta;
  int a_bytes;                             // bytes one activation box lands (TW * TH pixels x 64 k)
  int O;
  long long y_sb, y_sh, y_sw;              // element strides of the output view
  const float *bias;
  __nv_bfloat16 *y;
  int act;
  float alpha, scale;
};

// Pipeline (see the halo kernel's notes on the shallow UMMA queue): NACC accumulators in TMEM
// (4 up to BN = 128), the MMA warp probes the next stage's barrier before issuing the current
// stage's UMMAs, and the output leaves through a swizzled staging tile + ONE TMA store per
// 64-channel sub-tile (per-thread 16-byte stores at a 64..512-byte stride kept the LSU busy for
// longer than the MMAs of a thin tile).  BN <= 64: the two epilogue groups take alternate
// tiles; BN >= 128: they split a tile's 64-column sub-tiles.
// OutT = float: fp32 output (the fp32 mode's split-bf16 contractions, see split_bf16x3 in
// pointwise.cu) through 32-channel sub-tiles (one 128-byte swizzle row per pixel).
template <int BN, int STAGES, typename OutT>
__global__ void __launch_bounds__(kCThreads, BN == 32 ? 2 : 1)
conv_fwd_tc_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kBBytes = BN * kCK * 2;
  constexpr int kStageBytes = kCABytes + kBBytes;
  constexpr int NACC = 512 / BN >= 4 ? 4 : 2;
  constexpr bool kSplitTile = BN >= 128;                    // both groups work on every tile
  constexpr bool kF32 = sizeof(OutT) == 4;
  constexpr int kSubCh = kF32 ? 32 : (BN > 64 ? 64 : BN);   // channels of a staging sub-tile
  constexpr int kSubs = BN / kSubCh;
  constexpr int kSubBytes = kCM * kSubCh * (int)sizeof(OutT);
  uint8_t *a_base = smem;
  uint8_t *b_base = smem + STAGES * kCABytes;
  uint8_t *stage_base = smem + STAGES * kStageBytes;        // one staging sub-tile per group
  uint64_t *full = (uint64_t *)(stage_base + 2 * kSubBytes);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;
  uint64_t *acc_empty = acc_full + NACC;
  uint32_t *tmem_slot = (uint32_t *)(acc_empty + NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = blockIdx.x * prm.tiles_per_cta;
  const int t_end = min(t_begin + prm.tiles_per_cta, prm.total_tiles);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kSplitTile ? 8 : 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, NACC * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PROF_DECL;

  // tile -> (class, channel tile, sample, patch row, patch column); the classes and channel
  // tiles of a patch are adjacent (same activation patch: L2-resident across them)
  auto decode = [&](int tile, int &c, int &b, int &oh0, int &ow0, int &n0) {
    c = tile % prm.ncls;
    tile /= prm.ncls;
    n0 = (tile % prm.NT) * BN;
    int r = tile / prm.NT;
    ow0 = (r % prm.tiles_w) * prm.TW;
    r /= prm.tiles_w;
    oh0 = (r % prm.tiles_h) * prm.TH;
    b = r / prm.tiles_h;
  };

  if (warp == 0) {
    // warp-uniform loop, elected lane issues (see elect_one_sync)
    RingPos r;
    for (int tile = t_begin; tile < t_end; ++tile) {
      int ci, b, oh0, ow0, n0;
      decode(tile, ci, b, oh0, ow0, n0);
      const ConvClass &cl = prm.cls[ci];
      for (int g = 0; g < cl.G; ++g) {
        const CUtensorMap *am = &maps.a[prm.amap[g]];
        const int cw = ow0 + cl.aw[g], ch = oh0 + cl.ah[g];
        for (int kc = 0; kc < prm.KC; ++kc) {
          const int s = r.s;
          PROF_WAIT(0, mbar_wait(&empty[s], r.ph ^ 1));
          if (elect_one_sync()) {
            mbar_expect_tx(&full[s], prm.a_bytes + kBBytes);
            tma_load_4d(a_base + s * kCABytes, am, &full[s], kc * kCK, cw, ch, b);
            tma_load_3d(b_base + s * kBBytes, &maps.w, &full[s], kc * kCK, n0, cl.wtap[g]);
          }
          __syncwarp();
          r.template advance<STAGES>();
        }
      }
    }
    PROF_FLUSH(0, -1, 4);
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(kCM, BN, false, false);
    // K-major SW128 on both sides: 32 bytes per UMMA_K step inside the swizzle atom, SBO =
    // next group of 8 rows (1 KiB); constant high word, low word advanced by adds
    const uint32_t d_hi = desc_hi(1024, 2);
    const uint32_t a_lo0 = desc_lo(smem_u32(a_base), 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(b_base), 16);
    RingPos r, acc;
    bool ready = false;                  // stage r.s already known to be full (probed ahead)
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int a = acc.s;
      const int num_kb = prm.cls[tile % prm.ncls].G * prm.KC;
      PROF_WAIT(1, mbar_wait(&acc_empty[a], acc.ph ^ 1));
      const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = r.s;
        if (!ready) PROF_WAIT(0, mbar_wait(&full[s], r.ph));
        tc_fence_after();
        RingPos nx = r;
        nx.template advance<STAGES>();
        const bool more = kb + 1 < num_kb || tile + 1 < t_end;
        ready = more && mbar_try_wait(&full[nx.s], nx.ph);
        if (elect_one_sync()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)s * (kCABytes >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)s * (kBBytes >> 4);
#pragma unroll
          for (int k16 = 0; k16 < kCK / 16; ++k16)
            umma_bf16_lh(tmem_acc, a_lo + k16 * 2, d_hi, b_lo + k16 * 2, d_hi, idesc,
                         (kb > 0 || k16 > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (kb == num_kb - 1) umma_commit(&acc_full[a]);
        }
        __syncwarp();
        r = nx;
      }
      acc.template advance<NACC>();
    }
    PROF_FLUSH(1, 2, 5);
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool issuer = (warp - 2) % 4 == 0 && lane == 0;
    const bool lrelu = prm.act == 3;
    const float alpha = prm.alpha, scale = prm.scale;
    const uint32_t stage_u32 = smem_u32(stage_base) + (uint32_t)(grp * kSubBytes);
    constexpr int kRowBytes = kSubCh * (int)sizeof(OutT);   // 64 (64B swizzle) or 128
    constexpr int kChunks = kRowBytes / 16;                 // 16-byte chunks per staging row
    constexpr int kPerChunk = 16 / (int)sizeof(OutT);       // channels per chunk
    const uint32_t srow_addr = stage_u32 + (uint32_t)row * (uint32_t)kRowBytes;
    const uint32_t swz = kRowBytes == 64 ? (((uint32_t)row >> 1) & 3u) : ((uint32_t)row & 7u);
    int lt = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
      if (!kSplitTile && (lt & 1) != grp) continue;
      int ci, b, oh0, ow0, n0;
      decode(tile, ci, b, oh0, ow0, n0);
      const int a = lt % NACC;
      PROF_WAIT(0, mbar_wait(&acc_full[a], (lt / NACC) & 1));
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int sub = kSplitTile ? grp : 0; sub < kSubs; sub += kSplitTile ? 2 : 1) {
        uint32_t v[kSubCh / 16][16];
#pragma unroll
        for (int i = 0; i < kSubCh / 16; ++i) tmem_ld16(tmem_acc + (uint32_t)(sub * kSubCh + i * 16), v[i]);
        tmem_ld_wait();
        if (sub + (kSplitTile ? 2 : 1) >= kSubs) {          // this warp's last read of the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
        const int o0 = n0 + sub * kSubCh;
        if (o0 < prm.O) {
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(1 + grp, 128);
#pragma unroll
          for (int ch = 0; ch < kChunks; ++ch) {
            uint32_t w4[4];
            float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (prm.bias && o0 + ch * kPerChunk < prm.O) {
              const float4 b0 = __ldg(reinterpret_cast<const float4 *>(prm.bias + o0 + ch * kPerChunk));
              bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
              if (!kF32) {
                const float4 b1 = __ldg(reinterpret_cast<const float4 *>(prm.bias + o0 + ch * 8 + 4));
                bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (kF32) {
                const int e = ch * 4 + j;
                float v0 = __uint_as_float(v[e >> 4][e & 15]) + bb[j];
                if (lrelu) v0 = v0 > 0.f ? v0 : v0 * alpha;
                w4[j] = __float_as_uint(v0 * scale);
              } else {
                const int e = ch * 8 + 2 * j;
                float v0 = __uint_as_float(v[e >> 4][e & 15]) + bb[2 * j];
                float v1 = __uint_as_float(v[e >> 4][(e & 15) + 1]) + bb[2 * j + 1];
                if (lrelu) {
                  v0 = v0 > 0.f ? v0 : v0 * alpha;
                  v1 = v1 > 0.f ? v1 : v1 * alpha;
                }
                const __nv_bfloat162 pr = __floats2bfloat162_rn(v0 * scale, v1 * scale);
                w4[j] = *reinterpret_cast<const uint32_t *>(&pr);
              }
            }
            const uint32_t addr = srow_addr + (((uint32_t)ch ^ swz) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w4[0]), "r"(w4[1]), "r"(w4[2]),
                         "r"(w4[3])
                         : "memory");
          }
          fence_proxy_async();
          named_bar_sync(1 + grp, 128);
          if (issuer) {
            tma_store_4d(&maps.y[ci], stage_u32, o0, ow0, oh0, b);
            tma_store_commit();
          }
        }
      }
    }
    if (issuer) tma_store_wait_read<0>();
    if (warp == 2) PROF_FLUSH(3, -1, 6);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NACC * BN);
  }
}

// ------------------------------------------------------------------ CTA-pair implicit GEMM (C >= 128)
// The deep layers are GEMM-shaped (K = 1152 .. 4608) and bound by the shared-memory port: a
// 128 x N x 16 UMMA reads 4 KB of A and 32 N bytes of B, and the TMA writes of single-use
// operands cross the same 128 B / cycle -- 256 B / cycle of demand at N = 128, 192 at N = 256
// (measured: 47-57 % tensor-pipe activity for conv_fwd_tc_kernel<128 / 256>).  Here two CTAs of
// a cluster (the two SMs of a TPC) run ONE 256 x N UMMA per k-step (cta_group::2): each CTA
// lands its own 128-pixel A tile and HALF of the filter tile, the halves are exchanged by the
// hardware, so per CTA the port carries 32 KB + 32 KB per 512 MMA cycles at N = 256.
// Roles per CTA: warp 0 = TMA producer (both CTAs; all transaction bytes are counted on the
// LEADER's full barrier), warp 1 = MMA issuer (leader only; commits multicast to the barriers of
// both CTAs), warps 2-9 = epilogue (each CTA drains the 128 accumulator rows in its own TMEM
// and reports to the leader's acc_empty barrier).  The MMA warp probes the next stage's barrier
// BEFORE issuing the current stage's UMMAs (the tensor pipe's queue is a few instructions deep:
// a blocking wait between stages is idle time, see the halo kernel's notes).  Output through a
// swizzled staging tile and TMA stores, 64 channels at a time.
struct PairMaps {
  CUtensorMap a[4];
  CUtensorMap w;
  CUtensorMap y[kMaxClasses];
};

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kCThreads, 1)
conv_pair_tc_kernel(const __grid_constant__ PairMaps maps, const __grid_constant__ ConvParams prm,
                    int patches) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kBHalf = (BN / 2) * kCK * 2;                // this CTA's half of the filter tile
  constexpr int kStageBytes = kCABytes + kBHalf;
  constexpr int kStageTile = kCM * 64 * 2;                  // [128 rows][64 channels] staging tile
  uint8_t *a_base = smem;
  uint8_t *b_base = smem + STAGES * kCABytes;
  uint8_t *stage_base = smem + STAGES * kStageBytes;        // one staging tile per epilogue group
  uint64_t *full = (uint64_t *)(stage_base + 2 * kStageTile);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;
  uint64_t *acc_empty = acc_full + 2;
  uint32_t *tmem_slot = (uint32_t *)(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int t_begin = cluster_id * prm.tiles_per_cta;
  const int t_end = min(t_begin + prm.tiles_per_cta, prm.total_tiles);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);            // the leader's arrive.expect_tx (+ both CTAs' bytes)
      mbar_init(&empty[s], 1);           // the leader's multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);        // multicast commit
      mbar_init(&acc_empty[a], 16);      // 8 epilogue warps of each CTA (leader's copy is used)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // peer barriers initialised before anything remote
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PROF_DECL;

  // pair tile -> (class, channel tile, patch pair); this CTA's patch -> (sample, row, column)
  auto decode = [&](int pt, int &ci, int &n0, int &b, int &oh0, int &ow0, bool &valid) {
    ci = pt % prm.ncls;
    pt /= prm.ncls;
    n0 = (pt % prm.NT) * BN;
    const int patch = (pt / prm.NT) * 2 + (int)rank;
    valid = patch < patches;
    int r = patch;
    ow0 = (r % prm.tiles_w) * prm.TW;
    r /= prm.tiles_w;
    oh0 = (r % prm.tiles_h) * prm.TH;
    b = r / prm.tiles_h;                 // == B for the dummy patch of an odd count: zero-filled loads
  };

  if (warp == 0) {
    uint32_t lead_full[STAGES];
#pragma unroll
    for (int s = 0; s < STAGES; ++s) lead_full[s] = mapa_u32(smem_u32(&full[s]), 0);
    RingPos r;
    for (int pt = t_begin; pt < t_end; ++pt) {
      int ci, n0, b, oh0, ow0;
      bool valid;
      decode(pt, ci, n0, b, oh0, ow0, valid);
      const ConvClass &cl = prm.cls[ci];
      for (int g = 0; g < cl.G; ++g) {
        const CUtensorMap *am = &maps.a[prm.amap[g]];
        const int cw = ow0 + cl.aw[g], ch = oh0 + cl.ah[g];
        for (int kc = 0; kc < prm.KC; ++kc) {
          const int s = r.s;
          PROF_WAIT(0, mbar_wait(&empty[s], r.ph ^ 1));
          if (elect_one_sync()) {
            if (rank == 0) mbar_expect_tx(&full[s], 2 * (prm.a_bytes + kBHalf));
            uint32_t lf = lead_full[0];
#pragma unroll
            for (int i = 1; i < STAGES; ++i) lf = (s == i) ? lead_full[i] : lf;
            tma_load_4d_pair(a_base + s * kCABytes, am, lf, kc * kCK, cw, ch, b);
            tma_load_3d_pair(b_base + s * kBHalf, &maps.w, lf, kc * kCK, n0 + (int)rank * (BN / 2), cl.wtap[g]);
          }
          __syncwarp();
          r.template advance<STAGES>();
        }
      }
    }
    if (rank == 0) PROF_FLUSH(0, -1, 4);
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(2 * kCM, BN, false, false);
      const uint32_t d_hi = desc_hi(1024, 2);
      const uint32_t a_lo0 = desc_lo(smem_u32(a_base), 16);
      const uint32_t b_lo0 = desc_lo(smem_u32(b_base), 16);
      RingPos r;
      int lt = 0;
      bool ready = false;                // stage r.s already known to be full (probed ahead)
      for (int pt = t_begin; pt < t_end; ++pt, ++lt) {
        const int a = lt & 1;
        const int num_kb = prm.cls[pt % prm.ncls].G * prm.KC;
        PROF_WAIT(1, mbar_wait(&acc_empty[a], ((lt >> 1) & 1) ^ 1));
        const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int s = r.s;
          if (!ready) PROF_WAIT(0, mbar_wait(&full[s], r.ph));
          tc_fence_after();
          RingPos nx = r;
          nx.template advance<STAGES>();
          const bool more = kb + 1 < num_kb || pt + 1 < t_end;
          ready = more && mbar_try_wait(&full[nx.s], nx.ph);      // probe, consumed next iteration
          if (elect_one_sync()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)s * (kCABytes >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)s * (kBHalf >> 4);
#pragma unroll
            for (int k16 = 0; k16 < kCK / 16; ++k16)
              umma_bf16_lh_pair(tmem_acc, a_lo + k16 * 2, d_hi, b_lo + k16 * 2, d_hi, idesc,
                                (kb > 0 || k16 > 0) ? 1u : 0u);
            umma_commit_pair(&empty[s]);
            if (kb == num_kb - 1) umma_commit_pair(&acc_full[a]);
          }
          __syncwarp();
          r = nx;
        }
      }
      PROF_FLUSH(1, 2, 5);
    }
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool issuer = (warp - 2) % 4 == 0 && lane == 0;
    const uint32_t lead_acc_empty[2] = {mapa_u32(smem_u32(&acc_empty[0]), 0), mapa_u32(smem_u32(&acc_empty[1]), 0)};
    const uint32_t stage_u32 = smem_u32(stage_base) + (uint32_t)(grp * kStageTile);
    const uint32_t srow_addr = stage_u32 + (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)row & 7u;
    constexpr int kNChunks = BN / 64;
    int lt = 0;
    for (int pt = t_begin; pt < t_end; ++pt, ++lt) {
      int ci, n0, b, oh0, ow0;
      bool valid;
      decode(pt, ci, n0, b, oh0, ow0, valid);
      const int a = lt & 1;
      PROF_WAIT(0, mbar_wait(&acc_full[a], (lt >> 1) & 1));
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = grp; c < kNChunks; c += 2) {
        uint32_t v[4][16];
#pragma unroll
        for (int i = 0; i < 4; ++i) tmem_ld16(tmem_acc + (uint32_t)(c * 64 + i * 16), v[i]);
        tmem_ld_wait();
        if (c + 2 >= kNChunks) {           // this group's last chunk: the accumulator may be reused
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead_acc_empty[a]);
        }
        if (valid && n0 + c * 64 < prm.O) {
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(1 + grp, 128);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            uint32_t w4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = ch * 8 + 2 * j;
              const __nv_bfloat162 pr = __floats2bfloat162_rn(__uint_as_float(v[e >> 4][e & 15]) * prm.scale,
                                                              __uint_as_float(v[e >> 4][(e & 15) + 1]) * prm.scale);
              w4[j] = *reinterpret_cast<const uint32_t *>(&pr);
            }
            const uint32_t addr = srow_addr + (((uint32_t)ch ^ swz) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w4[0]), "r"(w4[1]), "r"(w4[2]),
                         "r"(w4[3])
                         : "memory");
          }
          fence_proxy_async();
          named_bar_sync(1 + grp, 128);
          if (issuer) {
            tma_store_4d(&maps.y[ci], stage_u32, n0 + c * 64, ow0, oh0, b);
            tma_store_commit();
          }
        }
      }
    }
    if (issuer) tma_store_wait_read<0>();
    if (warp == 2 && rank == 0) PROF_FLUSH(3, -1, 6);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // no CTA leaves while its peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ halo-resident 3x3 (C = 32 / 64)
// The thin, wide layers (32 / 64 channels at 64x512 / 32x256) are bound by operand delivery, not
// by the tensor pipe: an implicit GEMM that lands every filter tap as its own shared-memory tile
// moves each input pixel 9 times from L2.  Here every input pixel is landed ONCE: a CTA loads a
// (TH + R - 1) x PW pixel patch (rows of C channels = 64 or 128 bytes, K-major, 64B / 128B
// swizzle) with a single TMA box, and the A operand of tap (r, s) is the SAME buffer read
// through a descriptor whose start address is advanced by r * PW + s pixel rows.  That relies
// on the swizzle being a function of the absolute shared-memory address (start addresses that
// are not multiples of the 8-row swizzle atom, base-offset field 0) -- measured on sm_100a by
// tools/experiments/umma_rowshift.cu for both swizzle widths.  M = 128 consecutive "virtual
// pixels" of the patch (pitch PW); the PW - TW virtual pixels per row whose window crosses the
// row end are computed and discarded.  All R*S filter taps stay resident in shared memory.
// fprop of a valid conv: patch origin = tile origin.  dgrad (unit stride): the same kernel on
// dY with the patch origin moved by -(R-1), -(S-1) (TMA zero-fills outside dY) and the filter
// flipped / transposed by the host.
struct HaloMaps {
  CUtensorMap x, w, y;
};

struct HaloParams {
  int T, S;                 // taps, taps per filter row
  int flip;                 // read the filter blocks in reverse tap order (dgrad of a correlation)
  int PW, TH, TWo;          // patch pitch (pixels), output rows per tile, output columns per tile
  int MB;                   // 128-pixel M blocks per tile = TH * PW / 128
  int org_h, org_w;         // patch origin relative to the tile origin
  int tiles_w, tiles_h, total_tiles, tiles_per_cta;
  int H_out, W_out, O;
  int patch_bytes;          // bytes landed per patch
  long long y_off, y_sb, y_sh, y_sw;
  const float *bias;
  __nv_bfloat16 *y;
  int act;
  float alpha, scale;
};

// Pipeline.  The tensor pipe's queue is shallow (a handful of UMMAs: ~100-300 cycles of work at
// N <= 64), so EVERY cycle the issuing thread spends between UMMAs on barrier round trips idles
// it: with one acc_empty wait + fence + elect + commit per 128-pixel block the first version
// measured ~1290 cycles per block for 720 cycles of MMA work (tools/conv_roles.py,
// tools/debug/halo_trace.py; tools/experiments/umma_halo.cu shows the stream itself running at
// 40 cycles per 128x32x16 UMMA next to tcgen05.ld traffic).  Here the unit of synchronisation is
// the TILE: TMEM holds two SETS of MB accumulators (2 * MB * BN <= 512 columns), the MMA warp
// waits once per tile (patch landed, set drained), then issues all MB * T * C/16 UMMAs back to
// back with only a tcgen05.commit between blocks, so the epilogue of block 0 overlaps the MMAs
// of blocks 1.. and of the next tile's set.  The two epilogue groups take alternate blocks, each
// thread owning one pixel row (all BN channels: one tcgen05.wait, a full row of stores).
template <int ROWB, int BN, int NBUF>
__global__ void __launch_bounds__(kCThreads, 1)
conv_halo_tc_kernel(const __grid_constant__ HaloMaps maps, const __grid_constant__ HaloParams prm,
                    int buf_stride) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint64_t kLayout = ROWB == 64 ? 4 : 2;          // SWIZZLE_64B : SWIZZLE_128B
  constexpr int kWTap = BN * ROWB;                          // bytes of one filter tap [BN][C]
  constexpr int kSetBlocks = 256 / BN;                      // accumulators per set (MB <= this)
  constexpr int kSubCh = BN > 64 ? 64 : BN;                 // channels of one staging sub-tile
  constexpr int kSubs = BN / kSubCh;
  constexpr int kSubBytes = kCM * kSubCh * 2;               // [128 rows][kSubCh] swizzled
  uint8_t *w_base = smem;
  uint8_t *stage_base = smem + ((prm.T * kWTap + 1023) & ~1023);     // [2 groups][kSubs] sub-tiles
  uint8_t *p_base = stage_base + 2 * kSubs * kSubBytes;
  uint64_t *full = (uint64_t *)(p_base + NBUF * buf_stride);
  uint64_t *empty = full + NBUF;
  uint64_t *acc_full = empty + NBUF;                        // [2][kSetBlocks]
  uint64_t *set_empty = acc_full + 2 * kSetBlocks;          // [2]
  uint64_t *w_full = set_empty + 2;
  uint32_t *tmem_slot = (uint32_t *)(w_full + 1);
  float *bias_s = (float *)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);   // BN floats

  // warp roles: 0 = TMA producer, 1..8 = epilogue (two groups of four), 9 = MMA issuer
  constexpr int kMmaWarp = 9;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = blockIdx.x * prm.tiles_per_cta;
  const int t_end = min(t_begin + prm.tiles_per_cta, prm.total_tiles);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NBUF; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2 * kSetBlocks; ++a) mbar_init(&acc_full[a], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&set_empty[a], 8);      // every epilogue warp, once per tile
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + BN) {
    const int o = threadIdx.x - 32;
    bias_s[o] = (prm.bias && o < prm.O) ? prm.bias[o] : 0.f;
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  PROF_DECL;
  TRACE_DECL;

  auto decode = [&](int tile, int &b, int &oh0, int &ow0) {
    ow0 = (tile % prm.tiles_w) * prm.TWo;
    int r = tile / prm.tiles_w;
    oh0 = (r % prm.tiles_h) * prm.TH;
    b = r / prm.tiles_h;
  };

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(w_full, prm.T * kWTap);
      for (int t = 0; t < prm.T; ++t)
        tma_load_3d(w_base + t * kWTap, &maps.w, w_full, 0, 0, prm.flip ? prm.T - 1 - t : t);
    }
    __syncwarp();
    RingPos r;
    for (int tile = t_begin; tile < t_end; ++tile) {
      int b, oh0, ow0;
      decode(tile, b, oh0, ow0);
      const int s = r.s;
      PROF_WAIT(0, mbar_wait(&empty[s], r.ph ^ 1));
      if (elect_one_sync()) {
        mbar_expect_tx(&full[s], prm.patch_bytes);
        tma_load_4d(p_base + s * buf_stride, &maps.x, &full[s], 0, ow0 + prm.org_w, oh0 + prm.org_h, b);
        TRACE(0, 1, tile - t_begin);
      }
      __syncwarp();
      r.template advance<NBUF>();
    }
    PROF_FLUSH(0, -1, 4);
  } else if (warp == kMmaWarp) {
    constexpr uint32_t idesc = make_idesc(kCM, BN, false, false);
    mbar_wait(w_full, 0);
    const uint32_t d_hi = desc_hi(8 * ROWB, (uint32_t)kLayout);
    const uint32_t p_lo0 = desc_lo(smem_u32(p_base), 16);
    const uint32_t w_lo0 = desc_lo(smem_u32(w_base), 16);
    constexpr uint32_t kRow16 = ROWB >> 4;            // descriptor units per pixel row
    const int T = prm.T, S = prm.S, MB = prm.MB;
    const uint32_t row_step = (uint32_t)prm.PW * kRow16;
    RingPos r;
    int lt = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
      const int s = r.s, set = lt & 1;
      PROF_WAIT(0, mbar_wait(&full[s], r.ph));
      PROF_WAIT(1, mbar_wait(&set_empty[set], ((lt >> 1) & 1) ^ 1));
      TRACE0(1, 2, tile - t_begin);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t p_lo = p_lo0 + (uint32_t)s * (uint32_t)(buf_stride >> 4);
        uint32_t tmem_acc = tmem_base + (uint32_t)(set * 256);
        uint64_t *af = &acc_full[set * kSetBlocks];
        for (int mb = 0; mb < MB; ++mb) {
          uint32_t first = 0;
          // tap (r, s2): A = the patch read from pixel row mb*128 + r*PW + s2 onwards
          uint32_t a_row = p_lo + (uint32_t)(mb * kCM) * kRow16;
          uint32_t w_lo = w_lo0;
          int s2 = 0;
          for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int k16 = 0; k16 < ROWB / 32; ++k16) {
              umma_bf16_lh(tmem_acc, a_row + (uint32_t)s2 * kRow16 + k16 * 2, d_hi, w_lo + k16 * 2,
                           d_hi, idesc, first);
              first = 1u;
            }
            w_lo += kWTap >> 4;
            if (++s2 == S) { s2 = 0; a_row += row_step; }
          }
          umma_commit(af + mb);
          tmem_acc += BN;
        }
        umma_commit(&empty[s]);            // patch buffer reusable once these MMAs retire
        TRACE(1, 4, tile - t_begin);
      }
      __syncwarp();
      r.template advance<NBUF>();
    }
    PROF_FLUSH(1, 2, 5);
  } else {
    // Output path: registers -> swizzled staging tile in shared memory -> ONE TMA store per
    // block and 64-channel sub-tile.  (Per-thread 16-byte global stores of a pixel row each --
    // 32 lanes x 16 B at a 64..256-byte stride per instruction -- kept the LSU busy for ~1300
    // cycles per block: the epilogue, not the MMA stream, bounded the first versions.)
    const int q = warp & 3;                           // TMEM lane quarter of this warp
    const int grp = (warp - 1) >> 2;                  // epilogue group: blocks with (mb & 1) == grp ^ (lt & 1)
    const int row = q * 32 + lane;
    const int pw_shift = prm.PW == 64 ? 6 : 5;        // patch pitch is 32 or 64 pixels
    const int rows_pb = kCM >> pw_shift;              // image rows per 128-pixel block
    const bool lrelu = prm.act == 3;
    const bool has_bias = prm.bias != nullptr;
    const float alpha = prm.alpha, scale = prm.scale;
    const bool issuer = (q == 0 && lane == 0);
    const uint32_t stage_u32 = smem_u32(stage_base) + (uint32_t)(grp * kSubs * kSubBytes);
    // dense box order of the staging tile: row = (image row inside the block) * TWo + column
    const int hh_l = row >> pw_shift, ww = row & (prm.PW - 1);
    const bool col_ok = ww < prm.TWo;
    const uint32_t srow = (uint32_t)(hh_l * prm.TWo + ww);
    // 16-byte chunk c of staging row r lives at chunk c ^ swz(r) (64B / 128B TMA swizzle)
    constexpr int kChunks = kSubCh / 8;               // 16-byte chunks per row: 4 or 8
    const uint32_t swz = kSubCh == 32 ? ((srow >> 1) & 3u) : (srow & 7u);
    const uint32_t srow_addr = stage_u32 + srow * (uint32_t)(kSubCh * 2);
    int lt = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++lt) {
      int b, oh0, ow0;
      decode(tile, b, oh0, ow0);
      const int set = lt & 1;
      const uint32_t set_parity = (lt >> 1) & 1;
      // odd block counts: the group with the extra block alternates from tile to tile
      for (int mb = (grp ^ (lt & 1)); mb < prm.MB; mb += 2) {
        PROF_WAIT(0, mbar_wait(&acc_full[set * kSetBlocks + mb], set_parity));
        if (q == 0) TRACE0(2 + grp, 5, lt * 8 + mb);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(set * 256 + mb * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int sub = 0; sub < kSubs; ++sub) {
          uint32_t v[kSubCh / 16][16];
#pragma unroll
          for (int c = 0; c < kSubCh / 16; ++c) tmem_ld16(tmem_acc + (uint32_t)(sub * kSubCh + c * 16), v[c]);
          tmem_ld_wait();
          uint4 pk[kChunks];
#pragma unroll
          for (int ch = 0; ch < kChunks; ++ch) {
            uint32_t w4[4];
            // bias as two 16-byte broadcast loads per 8 columns: a broadcast LDS.32 costs a whole
            // shared-memory wavefront, and the UMMA operand reads saturate that port at N <= 64
            const float4 b0 = has_bias ? *reinterpret_cast<const float4 *>(bias_s + sub * kSubCh + ch * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b1 = has_bias ? *reinterpret_cast<const float4 *>(bias_s + sub * kSubCh + ch * 8 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = ch * 8 + 2 * j;           // column inside the sub-tile
              float v0 = __uint_as_float(v[e >> 4][e & 15]) + (j < 2 ? (j == 0 ? b0.x : b0.z) : (j == 2 ? b1.x : b1.z));
              float v1 = __uint_as_float(v[e >> 4][(e & 15) + 1]) + (j < 2 ? (j == 0 ? b0.y : b0.w) : (j == 2 ? b1.y : b1.w));
              if (lrelu) {
                v0 = v0 > 0.f ? v0 : v0 * alpha;
                v1 = v1 > 0.f ? v1 : v1 * alpha;
              }
              const __nv_bfloat162 pr = __floats2bfloat162_rn(v0 * scale, v1 * scale);
              w4[j] = *reinterpret_cast<const uint32_t *>(&pr);
            }
            pk[ch] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
          if (sub == 0) {                             // the previous store has left the staging tiles
            if (issuer) tma_store_wait_read<0>();
            named_bar_sync(1 + grp, 128);
          }
          if (col_ok) {
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
              const uint32_t a = srow_addr + (uint32_t)(sub * kSubBytes) + (((uint32_t)ch ^ swz) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[ch].x), "r"(pk[ch].y),
                           "r"(pk[ch].z), "r"(pk[ch].w)
                           : "memory");
            }
          }
        }
        fence_proxy_async();
        named_bar_sync(1 + grp, 128);
        if (issuer) {
#pragma unroll
          for (int sub = 0; sub < kSubs; ++sub)
            tma_store_4d(&maps.y, stage_u32 + (uint32_t)(sub * kSubBytes), sub * kSubCh, ow0,
                         oh0 + mb * rows_pb, b);
          tma_store_commit();
        }
        if (q == 0) TRACE0(2 + grp, 7, lt * 8 + mb);
      }
      // this warp is done reading the set's accumulators
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&set_empty[set]);
    }
    if (issuer) tma_store_wait_read<0>();             // staging must outlive the last store's read
    if (warp == 3) PROF_FLUSH(3, -1, 6);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ wgrad
struct WgMaps {
  CUtensorMap a[4];    // x window of filter row r
  CUtensorMap g;       // dY
  CUtensorMap o;       // fp32 result [R][S*C][O] (split-K partial tiles are TMA-reduced into it)
};

struct WgParams {
  int tiles_w, tiles_h, TW, TH;   // 64-pixel patches
  int total_pb, pb_per_split;
  int MT, NT;                     // 128-wide virtual-channel tiles, BN-wide out-channel tiles
  int SC, O;                      // S*C, out channels
  float *out;                     // [splits][R][SC][O], or the result itself when atomic
  long long split_stride;
  int atomic;                     // partial tiles are added into `out` with 16-byte reductions
  // operand geometry of one stage (bytes).  Row-resident mode (rows = 1, NR = 3, unit row
  // stride): ONE x patch of TH + 2 image rows is landed per stage (maps.a[3]) and filter row r
  // reads it from pixel row r * TW onwards -- the three row windows are the same pixels shifted
  // by whole image rows, a multiple of the 8-row swizzle atom when TW % 8 == 0 -- so x crosses
  // L2->SM (TH + 2) / TH times per stage instead of three times.
  int rows;
  int ksteps;                     // UMMA K steps (16 pixels each) per stage
  int a_half, a_rr, a_slot;       // between the two 64-element blocks / filter rows / stages of A
  int b_blk, b_slot;              // between 64-channel blocks / stages of dY
};

// NR = filter rows accumulated by one CTA.  NR == 1: one row per CTA (grid.z = R), every CTA
// streams its own copy of dY.  NR == 3 (3x3 filters of the thin / mid layers, O <= 128): the
// three row-windows of x share ONE landing of the dY tile per stage and feed three TMEM
// accumulators -- dY crosses L2->SM once instead of three times (the 32-channel 64x512 layers
// are bound by exactly that traffic: 1.6 GB per call for 0.54 GB of tensors).
template <int BN, int STAGES, int NR>
__global__ void __launch_bounds__(kWgThreads)
conv_wgrad_tc_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgParams prm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kTmemCols = NR * BN <= 32 ? 32 : (NR * BN <= 64 ? 64 : (NR * BN <= 128 ? 128 : (NR * BN <= 256 ? 256 : 512)));
  const int kASlot = prm.a_slot, kBBytes = prm.b_slot;
  const int kStageBytes = kASlot + kBBytes;
  uint8_t *a_base = smem;
  uint8_t *b_base = smem + STAGES * kASlot;
  uint64_t *full = (uint64_t *)(smem + STAGES * kStageBytes);
  uint64_t *empty = full + STAGES;
  uint64_t *acc_full = empty + STAGES;
  uint32_t *tmem_slot = (uint32_t *)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int m0 = (blockIdx.y % prm.MT) * kCM;
  const int n0 = (blockIdx.y / prm.MT) * BN;
  const int r_first = NR == 1 ? (int)blockIdx.z : 0;     // filter rows [r_first, r_first + NR)
  const int pb_begin = split * prm.pb_per_split;
  const int pb_end = min(pb_begin + prm.pb_per_split, prm.total_pb);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  PROF_DECL;

  if (warp == 0) {
    RingPos rp;
    for (int pb = pb_begin; pb < pb_end; ++pb) {
      const int s = rp.s;
      int t = pb;
      const int ow0 = (t % prm.tiles_w) * prm.TW;
      t /= prm.tiles_w;
      const int oh0 = (t % prm.tiles_h) * prm.TH;
      const int b = t / prm.tiles_h;
      PROF_WAIT(0, mbar_wait(&empty[s], rp.ph ^ 1));
      if (elect_one_sync()) {
        mbar_expect_tx(&full[s], kStageBytes);
        if (NR == 3 && prm.rows) {
          uint8_t *a_dst = a_base + s * kASlot;
          tma_load_4d(a_dst, &maps.a[3], &full[s], m0, ow0, oh0, b);
          tma_load_4d(a_dst + prm.a_half, &maps.a[3], &full[s], m0 + 64, ow0, oh0, b);
        } else {
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            const CUtensorMap *am = &maps.a[r_first + rr];
            uint8_t *a_dst = a_base + s * kASlot + rr * kCABytes;
            tma_load_4d(a_dst, am, &full[s], m0, ow0, oh0, b);
            tma_load_4d(a_dst + 8192, am, &full[s], m0 + 64, ow0, oh0, b);
          }
        }
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_4d(b_base + s * kBBytes + j * prm.b_blk, &maps.g, &full[s], n0 + 64 * j, ow0, oh0, b);
      }
      __syncwarp();
      rp.template advance<STAGES>();
    }
    PROF_FLUSH(0, -1, 4);
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(kCM, BN, true, true);
    // MN-major SW128: 16 pixel rows = 2 KiB per UMMA_K step; LBO = next 64-channel block
    // (8 KiB), SBO = next group of 8 pixel rows (1 KiB)
    const uint32_t d_hi = desc_hi(1024, 2);
    const uint32_t a_lo0 = desc_lo(smem_u32(a_base), (uint32_t)prm.a_half);
    const uint32_t b_lo0 = desc_lo(smem_u32(b_base), (uint32_t)prm.b_blk);
    const int ksteps = prm.ksteps;
    RingPos rp;
    for (int pb = pb_begin; pb < pb_end; ++pb) {
      const int s = rp.s;
      PROF_WAIT(0, mbar_wait(&full[s], rp.ph));
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t b_lo = b_lo0 + (uint32_t)s * (uint32_t)(kBBytes >> 4);
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
          const uint32_t a_lo = a_lo0 + (uint32_t)s * (uint32_t)(kASlot >> 4) + (uint32_t)rr * (uint32_t)(prm.a_rr >> 4);
#pragma unroll 4
          for (int k16 = 0; k16 < ksteps; ++k16)
            umma_bf16_lh(tmem_acc + (uint32_t)(rr * BN), a_lo + k16 * (2048 >> 4), d_hi,
                         b_lo + k16 * (2048 >> 4), d_hi, idesc, (pb > pb_begin || k16 > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
      rp.template advance<STAGES>();
    }
    if (elect_one_sync()) umma_commit(acc_full);
    __syncwarp();
    PROF_FLUSH(1, 2, 5);
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    PROF_WAIT(0, mbar_wait(acc_full, 0));
    tc_fence_after();
    const bool have = pb_end > pb_begin;
    const bool atomic = prm.atomic != 0;    // split-K partials meet in the (zeroed) result
    if (atomic) {
      // Partial tile -> swizzled staging tile (the operand ring is idle once acc_full has
      // fired) -> ONE TMA reduce-add per 32 columns.  (Per-thread 16-byte reductions -- 8192
      // per CTA at BN = 256 -- took half of the CTA's lifetime: tools/conv_roles.py.)
      if (have) {
        const int row = q * 32 + lane;
        const bool issuer = threadIdx.x == 64;
        const uint32_t stage_u32 = smem_u32(a_base);
        const uint32_t srow = stage_u32 + (uint32_t)row * 128u;
        const uint32_t swz = (uint32_t)row & 7u;
#pragma unroll 1
        for (int rr = 0; rr < NR; ++rr) {
#pragma unroll 1
          for (int c = 0; c < BN; c += 32) {
            uint32_t v[2][16];
            tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(rr * BN + c), v[0]);
            tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(rr * BN + c + 16), v[1]);
            tmem_ld_wait();
            if (issuer) tma_store_wait_read<0>();
            named_bar_sync(1, 128);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint32_t *src = &v[ch >> 2][(ch & 3) * 4];
              const uint32_t addr = srow + (((uint32_t)ch ^ swz) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(src[0]), "r"(src[1]),
                           "r"(src[2]), "r"(src[3])
                           : "memory");
            }
            fence_proxy_async();
            named_bar_sync(1, 128);
            if (issuer) {
              tma_reduce_add_3d(&maps.o, stage_u32, n0 + c, m0, r_first + rr);
              tma_store_commit();
            }
          }
        }
        if (issuer) tma_store_wait_read<0>();
      }
    } else {
#pragma unroll 1
      for (int rr = 0; rr < NR; ++rr) {
        float *op = prm.out + (long long)split * prm.split_stride +
                    ((long long)(r_first + rr) * prm.SC + m) * prm.O + n0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(rr * BN + c), v);
          tmem_ld_wait();
          if (m < prm.SC) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              if (n0 + c + j4 * 4 < prm.O) {
                float4 o4;
                o4.x = have ? __uint_as_float(v[j4 * 4 + 0]) : 0.f;
                o4.y = have ? __uint_as_float(v[j4 * 4 + 1]) : 0.f;
                o4.z = have ? __uint_as_float(v[j4 * 4 + 2]) : 0.f;
                o4.w = have ? __uint_as_float(v[j4 * 4 + 3]) : 0.f;
                *reinterpret_cast<float4 *>(op + c + j4 * 4) = o4;
              }
            }
          }
        }
      }
    }
    if (warp == 2) PROF_FLUSH(3, -1, 6);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, kTmemCols);
  }
}

__global__ void wgrad_reduce_kernel(const float *__restrict__ ws, float *__restrict__ out,
                                    long long n4, int splits, long long split_stride4) {
  const float4 *w4 = reinterpret_cast<const float4 *>(ws);
  float4 *o4 = reinterpret_cast<float4 *>(out);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 acc = w4[i];
    for (int s = 1; s < splits; ++s) {
      const float4 v = w4[i + s * split_stride4];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    o4[i] = acc;
  }
}

// 4-D bf16 tensor map, 128-byte swizzle, zero fill out of range
bool make_map4(CUtensorMap *m, const void *ptr, const uint64_t dims[4], const uint64_t strides_b[3],
               const uint32_t box[4]) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t s[3] = {strides_b[0], strides_b[1], strides_b[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(ptr), d, s, bx, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_map3w(CUtensorMap *m, const void *ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                uint32_t box0, uint32_t box1, uint64_t s1_elems = 0, uint64_t s2_elems = 0) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {(s1_elems ? s1_elems : d0) * 2, (s2_elems ? s2_elems : d0 * d1) * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(ptr), dims, strides, box,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_map_sw(CUtensorMap *m, const void *ptr, int rank, const uint64_t *dims,
                 const uint64_t *strides_b, const uint32_t *box, bool sw64, bool f32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t d[4];
  cuuint64_t s[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_b[i];
  return enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
             const_cast<void *>(ptr), d, s, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// one CTA per SM: it owns all 512 TMEM columns (two sets of 256 / BN accumulators)
template <int ROWB, int BN, int NBUF>
int launch_halo(const HaloMaps &maps, HaloParams prm, int buf_stride, cudaStream_t st) {
  const int smem = ((prm.T * BN * ROWB + 1023) & ~1023) + 2 * kCM * BN * 2 + NBUF * buf_stride + 1024 + 1024;
  static int configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(conv_halo_tc_kernel<ROWB, BN, NBUF>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("conv_halo_tc: cannot reserve %d bytes of shared memory", smem);
      return DUSTY_ECUDA;
    }
    configured = smem;
  }
  const int resident = num_sms();
  int ctas = prm.total_tiles < resident ? prm.total_tiles : resident;
  prm.tiles_per_cta = (prm.total_tiles + ctas - 1) / ctas;
  ctas = (prm.total_tiles + prm.tiles_per_cta - 1) / prm.tiles_per_cta;
  conv_halo_tc_kernel<ROWB, BN, NBUF><<<ctas, kCThreads, smem, st>>>(maps, prm, buf_stride);
  return 0;
}

template <int BN, int STAGES, typename OutT>
constexpr int conv_smem_bytes() {
  return STAGES * (kCABytes + BN * kCK * 2) +
         2 * kCM * (sizeof(OutT) == 4 ? 128 : (BN > 64 ? 128 : BN * 2)) + 256 + 1024;
}

template <int BN, int STAGES, typename OutT>
int launch_conv(const ConvMaps &maps, ConvParams prm, cudaStream_t st) {
  constexpr int smem = conv_smem_bytes<BN, STAGES, OutT>();
  static bool configured = false;
  if (int rc = set_smem(conv_fwd_tc_kernel<BN, STAGES, OutT>, smem, &configured)) return rc;
  // BN = 32 tiles are epilogue-bound (1-4 k-blocks of 160 tensor cycles per 8 KB of output):
  // two co-resident CTAs double the epilogue warps; wider tiles own the SM
  const int resident = num_sms() * (BN == 32 ? 2 : 1);
  int ctas = prm.total_tiles < resident ? prm.total_tiles : resident;
  prm.tiles_per_cta = (prm.total_tiles + ctas - 1) / ctas;
  ctas = (prm.total_tiles + prm.tiles_per_cta - 1) / prm.tiles_per_cta;
  conv_fwd_tc_kernel<BN, STAGES, OutT><<<ctas, kCThreads, smem, st>>>(maps, prm);
  return 0;
}

template <int BN, int STAGES, int NR>
int launch_wgrad(const WgMaps &maps, const WgParams &prm, int splits, int R, cudaStream_t st) {
  const int smem = STAGES * (prm.a_slot + prm.b_slot) + 256 + 1024;
  static int configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN, STAGES, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem) != cudaSuccess) {
      set_error("conv_wgrad_tc: cannot reserve %d bytes of shared memory", smem);
      return DUSTY_ECUDA;
    }
    configured = smem;
  }
  dim3 grid((unsigned)splits, (unsigned)(prm.MT * prm.NT), (unsigned)(NR == 1 ? R : 1));
  conv_wgrad_tc_kernel<BN, STAGES, NR><<<grid, kWgThreads, smem, st>>>(maps, prm);
  return 0;
}

// pixel patch of `n` pixels (n = 128 or 64): widest power-of-two row segment with the least
// padding waste over a W x H output
// Patch of AT MOST `n` pixels, any TW x TH (a TMA box need not be a power of two): the fewest
// tiles over a W x H output, ties broken towards fuller tiles.  The operand tile keeps its
// 128 rows; rows past TW * TH hold stale data whose products are never stored.  (A data
// gradient is taken on the padded grid -- 66 x 10, 130 x 18, parity classes of 33 x 5 -- where
// power-of-two patches waste up to half of every tile.)
void pick_patch_any(int W, int H, int n, int *TW, int *TH) {
  long long best_tiles = -1;
  int best_fill = 0;
  for (int tw = 1; tw <= n && tw <= 256; ++tw) {
    if (tw > W && tw != 1) break;
    int th = n / tw;
    if (th > H) th = H;
    if (th > 256) th = 256;
    if (th < 1) continue;
    // rows of a box land 8 to a swizzle atom; any count works, but keep boxes of >= 8 pixels
    const long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
    const int fill = tw * th;
    // ties: fuller tiles, then wider rows (long contiguous runs per box row)
    if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && fill >= best_fill)) {
      best_tiles = tiles;
      best_fill = fill;
      *TW = tw;
      *TH = th;
    }
  }
}

void pick_patch(int W, int H, int n, int *TW, int *TH) {
  long long best = -1;
  for (int tw = n; tw >= 1; tw >>= 1) {
    const int th = n / tw;
    const long long cover = (long long)((W + tw - 1) / tw) * tw * ((H + th - 1) / th) * th;
    if (best < 0 || cover < best) {
      best = cover;
      *TW = tw;
      *TH = th;
    }
  }
}

}  // namespace

}  // namespace dusty

using namespace dusty;

static bool pair_enabled() {
  static const bool off = [] { const char *e = getenv("DUSTY_CONV_PAIR"); return e && atoi(e) == 0; }();
  return !off;
}

template <int BN, int STAGES>
int launch_pair(const PairMaps &maps, ConvParams prm, int patches, cudaStream_t st) {
  constexpr int smem = STAGES * (kCABytes + (BN / 2) * kCK * 2) + 2 * kCM * 64 * 2 + 1024 + 1024;
  static bool configured = false;
  if (int rc = set_smem(conv_pair_tc_kernel<BN, STAGES>, smem, &configured)) return rc;
  int clusters = num_sms() / 2;
  if (prm.total_tiles < clusters) clusters = prm.total_tiles;
  prm.tiles_per_cta = (prm.total_tiles + clusters - 1) / clusters;
  clusters = (prm.total_tiles + prm.tiles_per_cta - 1) / prm.tiles_per_cta;
  conv_pair_tc_kernel<BN, STAGES><<<2 * clusters, kCThreads, smem, st>>>(maps, prm, patches);
  return 0;
}

// One launch of the implicit-GEMM kernel over `ncls` classes.  Host-side description of a class:
struct HostClass {
  int G;
  const int *dh, *dw, *wtap;
  int H_out, W_out;
  long long y_off;
};

static int conv_launch(const char *who, const void *x, const void *wpk, const float *bias, void *y,
                       int B, int H_in, int W_in, int C, int O, int mode, int ncls,
                       const HostClass *hc, int S, int stride_h, int stride_w, long long y_sb,
                       long long y_sh, long long y_sw, int act, float alpha, float scale,
                       long long w_sn, long long w_sg, int w_taps, cudaStream_t st, bool out_f32 = false) {
  const int Kg = mode == 1 ? S * C : C;
  ConvMaps maps;
  ConvParams prm;
  prm.ncls = ncls;
  prm.KC = (Kg + kCK - 1) / kCK;
  int H_max = 0, W_max = 0;
  for (int c = 0; c < ncls; ++c) {
    H_max = hc[c].H_out > H_max ? hc[c].H_out : H_max;
    W_max = hc[c].W_out > W_max ? hc[c].W_out : W_max;
  }
  pick_patch_any(W_max, H_max, kCM, &prm.TW, &prm.TH);
  prm.a_bytes = prm.TW * prm.TH * kCK * 2;
  const uint32_t box[4] = {(uint32_t)kCK, (uint32_t)prm.TW, (uint32_t)prm.TH, 1u};
  const __nv_bfloat16 *xb = (const __nv_bfloat16 *)x;
  bool ok = true;
  for (int g = 0; g < kMaxGroups; ++g) prm.amap[g] = 0;
  if (mode == 1) {
    const HostClass &h = hc[0];
    for (int g = 0; g < h.G; ++g) {
      const int dh = h.dh[g], dw = h.dw[g];
      if (!(dh >= 0 && dw >= 0 && (long long)(h.H_out - 1) * stride_h + dh < H_in &&
            (long long)(h.W_out - 1) * stride_w + dw + S <= W_in)) {
        set_error("%s: window outside the input", who);
        return DUSTY_EINVAL;
      }
      const uint64_t dims[4] = {(uint64_t)Kg, (uint64_t)h.W_out, (uint64_t)h.H_out, (uint64_t)B};
      const uint64_t strides[3] = {(uint64_t)stride_w * C * 2, (uint64_t)stride_h * W_in * C * 2,
                                   (uint64_t)H_in * W_in * C * 2};
      ok = ok && make_map4(&maps.a[g], xb + ((long long)dh * W_in + dw) * C, dims, strides, box);
      prm.amap[g] = g;
    }
    for (int g = h.G; g < 4; ++g) maps.a[g] = maps.a[0];
  } else {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W_in, (uint64_t)H_in, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W_in * C * 2, (uint64_t)H_in * W_in * C * 2};
    ok = make_map4(&maps.a[0], xb, dims, strides, box);
    for (int g = 1; g < 4; ++g) maps.a[g] = maps.a[0];
  }
  int Gw = w_taps;
  for (int c = 0; c < kMaxClasses; ++c) {
    ConvClass &cl = prm.cls[c];
    const HostClass &h = hc[c < ncls ? c : 0];
    cl.G = h.G; cl.H_out = h.H_out; cl.W_out = h.W_out; cl.y_off = h.y_off;
    for (int g = 0; g < kMaxGroups; ++g) {
      const bool in = g < h.G;
      cl.aw[g] = (in && mode == 0) ? h.dw[g] : 0;
      cl.ah[g] = (in && mode == 0) ? h.dh[g] : 0;
      cl.wtap[g] = in ? (h.wtap ? h.wtap[g] : g) : 0;
    }
    if (!h.wtap && c < ncls && h.G > Gw) Gw = h.G;
  }
  for (int c = 0; c < ncls; ++c)
    for (int g = 0; g < hc[c].G; ++g)
      if (prm.cls[c].wtap[g] < 0 || prm.cls[c].wtap[g] >= Gw) {
        set_error("%s: wtap out of range", who);
        return DUSTY_EINVAL;
      }
  const int BN = O > 128 ? 256 : (O > 64 ? 128 : (O > 32 ? 64 : 32));
  ok = ok && make_map3w(&maps.w, wpk, (uint64_t)Kg, (uint64_t)O, (uint64_t)Gw, kCK, (uint32_t)BN,
                        (uint64_t)w_sn, (uint64_t)w_sg);
  if (!ok) {
    set_error("%s: cuTensorMapEncodeTiled failed", who);
    return DUSTY_ECUDA;
  }
  prm.tiles_w = (W_max + prm.TW - 1) / prm.TW;
  prm.tiles_h = (H_max + prm.TH - 1) / prm.TH;
  prm.NT = (O + BN - 1) / BN;
  const long long total = (long long)prm.tiles_w * prm.tiles_h * prm.NT * B * ncls;
  if (total > 0x7fffffff) {
    set_error("%s: too many tiles", who);
    return DUSTY_EINVAL;
  }
  prm.total_tiles = (int)total;
  prm.O = O;
  prm.y_sb = y_sb; prm.y_sh = y_sh; prm.y_sw = y_sw;
  prm.bias = bias; prm.y = (__nv_bfloat16 *)y; prm.act = act; prm.alpha = alpha; prm.scale = scale;
  // deep layers: CTA pairs (plain output, no bias / activation epilogue)
  if (pair_enabled() && !out_f32 && BN >= 128 && Kg % kCK == 0 && prm.KC * hc[0].G >= 4 && bias == nullptr &&
      act == 1 && O % 8 == 0) {
    PairMaps pm;
    for (int g = 0; g < 4; ++g) pm.a[g] = maps.a[g];
    pm.w = maps.w;
    // the pair kernel lands HALF of the filter tile per CTA
    ok = make_map3w(&pm.w, wpk, (uint64_t)Kg, (uint64_t)O, (uint64_t)Gw, kCK, (uint32_t)(BN / 2),
                    (uint64_t)w_sn, (uint64_t)w_sg);
    const uint32_t ybox[4] = {64u, (uint32_t)prm.TW, (uint32_t)prm.TH, 1u};
    for (int c = 0; c < kMaxClasses; ++c) {
      const HostClass &h = hc[c < ncls ? c : 0];
      const uint64_t ydims[4] = {(uint64_t)O, (uint64_t)h.W_out, (uint64_t)h.H_out, (uint64_t)B};
      const uint64_t ystr[3] = {(uint64_t)y_sw * 2, (uint64_t)y_sh * 2, (uint64_t)y_sb * 2};
      ok = ok && make_map_sw(&pm.y[c], (const __nv_bfloat16 *)y + h.y_off, 4, ydims, ystr, ybox, false);
    }
    if (!ok) {
      set_error("%s: cuTensorMapEncodeTiled failed (pair kernel)", who);
      return DUSTY_ECUDA;
    }
    const int patches = prm.tiles_w * prm.tiles_h * B;
    prm.total_tiles = ((patches + 1) / 2) * prm.NT * ncls;
    return BN == 256 ? launch_pair<256, 5>(pm, prm, patches, st) : launch_pair<128, 7>(pm, prm, patches, st);
  }
  {
    const int sub_ch = out_f32 ? 32 : (BN > 64 ? 64 : BN);
    const uint64_t esz = out_f32 ? 4 : 2;
    const uint32_t ybox[4] = {(uint32_t)sub_ch, (uint32_t)prm.TW, (uint32_t)prm.TH, 1u};
    for (int c = 0; c < kMaxClasses; ++c) {
      const HostClass &h = hc[c < ncls ? c : 0];
      const uint64_t ydims[4] = {(uint64_t)O, (uint64_t)h.W_out, (uint64_t)h.H_out, (uint64_t)B};
      const uint64_t ystr[3] = {(uint64_t)y_sw * esz, (uint64_t)y_sh * esz, (uint64_t)y_sb * esz};
      ok = ok && make_map_sw(&maps.y[c], (const uint8_t *)y + h.y_off * (long long)esz, 4, ydims, ystr, ybox,
                             sub_ch * esz == 64, out_f32);
    }
    if (!ok) {
      set_error("%s: cuTensorMapEncodeTiled failed (output view)", who);
      return DUSTY_ECUDA;
    }
  }
  if (out_f32) {
    switch (BN) {
      case 256: return launch_conv<256, 3, float>(maps, prm, st);
      case 128: return launch_conv<128, 5, float>(maps, prm, st);
      case 64: return launch_conv<64, 6, float>(maps, prm, st);
      default: return launch_conv<32, 4, float>(maps, prm, st);
    }
  }
  switch (BN) {
    case 256: return launch_conv<256, 3, __nv_bfloat16>(maps, prm, st);
    case 128: return launch_conv<128, 5, __nv_bfloat16>(maps, prm, st);
    case 64: return launch_conv<64, 6, __nv_bfloat16>(maps, prm, st);
    default: return launch_conv<32, 4, __nv_bfloat16>(maps, prm, st);
  }
}

extern "C" int dusty_conv2d_tc(const void *x, const void *wpk, const float *bias, void *y, int B,
                               int H_in, int W_in, int C, int H_out, int W_out, int O, int mode,
                               int G, const int *tap_dh, const int *tap_dw, int S, int stride_h,
                               int stride_w, long long y_off, long long y_sb, long long y_sh,
                               long long y_sw, int act, float alpha, float scale, long long w_sn,
                               long long w_sg, const int *wtap, int w_taps, int out_dtype,
                               void *stream) {
  DUSTY_CHECK_ARG(x && wpk && y, "null pointer");
  DUSTY_CHECK_ARG(out_dtype == DUSTY_BF16 || out_dtype == DUSTY_F32, "output dtype: bf16 or fp32");
  DUSTY_CHECK_ARG(get_encode() != nullptr, "cuTensorMapEncodeTiled unavailable");
  DUSTY_CHECK_ARG(B > 0 && H_in > 0 && W_in > 0 && H_out > 0 && W_out > 0, "empty tensor");
  DUSTY_CHECK_ARG(C % 8 == 0 && O % 8 == 0, "channel counts must be multiples of 8");
  DUSTY_CHECK_ARG(G >= 1 && G <= kMaxGroups, "1..16 groups");
  DUSTY_CHECK_ARG(w_sn >= 0 && w_sg >= 0 && w_sn % 8 == 0 && w_sg % 8 == 0, "weight strides: multiples of 8");
  DUSTY_CHECK_ARG(w_taps >= 0 && (wtap == nullptr || w_taps >= 1), "w_taps: filter blocks in the weight tensor");
  DUSTY_CHECK_ARG(mode == 0 || mode == 1, "mode: 0 = tap, 1 = window");
  DUSTY_CHECK_ARG(mode == 1 ? (G <= 4 && S >= 1) : (stride_h == 1 && stride_w == 1),
                  "window mode: at most 4 filter rows; tap mode: unit stride");
  DUSTY_CHECK_ARG(aligned16(x) && aligned16(wpk) && aligned16(y), "16-byte alignment");
  DUSTY_CHECK_ARG((y_off % 8 == 0) && (y_sb % 8 == 0) && (y_sh % 8 == 0) && (y_sw % 8 == 0),
                  "output strides must keep 16-byte alignment");
  const HostClass hc = {G, tap_dh, tap_dw, wtap, H_out, W_out, y_off};
  if (int rc = conv_launch("dusty_conv2d_tc", x, wpk, bias, y, B, H_in, W_in, C, O, mode, 1, &hc, S,
                           stride_h, stride_w, y_sb, y_sh, y_sw, act, alpha, scale, w_sn, w_sg,
                           wtap ? w_taps : 0, (cudaStream_t)stream, out_dtype == DUSTY_F32))
    return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_conv2d_tc_classes(const void *x, const void *wpk, void *y, int B, int H_in,
                                       int W_in, int C, int O, int ncls, const int *cls_G,
                                       const int *tap_dh, const int *tap_dw, const int *wtap,
                                       const int *cls_H_out, const int *cls_W_out,
                                       const long long *cls_y_off, long long y_sb, long long y_sh,
                                       long long y_sw, long long w_sn, long long w_sg, int w_taps,
                                       int out_dtype, void *stream) {
  DUSTY_CHECK_ARG(out_dtype == DUSTY_BF16 || out_dtype == DUSTY_F32, "output dtype: bf16 or fp32");
  DUSTY_CHECK_ARG(x && wpk && y && cls_G && tap_dh && tap_dw && wtap && cls_H_out && cls_W_out && cls_y_off,
                  "null pointer");
  DUSTY_CHECK_ARG(get_encode() != nullptr, "cuTensorMapEncodeTiled unavailable");
  DUSTY_CHECK_ARG(B > 0 && H_in > 0 && W_in > 0, "empty tensor");
  DUSTY_CHECK_ARG(C % 8 == 0 && O % 8 == 0, "channel counts must be multiples of 8");
  DUSTY_CHECK_ARG(ncls >= 1 && ncls <= kMaxClasses, "1..4 classes");
  DUSTY_CHECK_ARG(w_sn >= 0 && w_sg >= 0 && w_sn % 8 == 0 && w_sg % 8 == 0, "weight strides: multiples of 8");
  DUSTY_CHECK_ARG(w_taps >= 1, "w_taps: filter blocks in the weight tensor");
  DUSTY_CHECK_ARG(aligned16(x) && aligned16(wpk) && aligned16(y), "16-byte alignment");
  DUSTY_CHECK_ARG((y_sb % 8 == 0) && (y_sh % 8 == 0) && (y_sw % 8 == 0), "output strides must keep 16-byte alignment");
  HostClass hc[kMaxClasses];
  int at = 0;
  for (int c = 0; c < ncls; ++c) {
    DUSTY_CHECK_ARG(cls_G[c] >= 1 && cls_G[c] <= kMaxGroups, "1..16 taps per class");
    DUSTY_CHECK_ARG(cls_H_out[c] > 0 && cls_W_out[c] > 0, "empty class");
    DUSTY_CHECK_ARG(cls_y_off[c] % 8 == 0, "class origin must keep 16-byte alignment");
    hc[c] = {cls_G[c], tap_dh + at, tap_dw + at, wtap + at, cls_H_out[c], cls_W_out[c], cls_y_off[c]};
    at += cls_G[c];
  }
  if (int rc = conv_launch("dusty_conv2d_tc_classes", x, wpk, nullptr, y, B, H_in, W_in, C, O, 0, ncls,
                           hc, 1, 1, 1, y_sb, y_sh, y_sw, 1, 0.f, 1.f, w_sn, w_sg, w_taps,
                           (cudaStream_t)stream, out_dtype == DUSTY_F32))
    return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

extern "C" int dusty_conv_trace(unsigned long long *out, int cap) {
#ifdef DUSTY_ROLE_PROF
  if (cap < 8192) return -1;
  if (cudaMemcpyFromSymbol(out, s_trace, sizeof(unsigned long long) * 8192) != cudaSuccess) return -1;
  static unsigned long long zeros[8192];
  cudaMemcpyToSymbol(s_trace, zeros, sizeof(zeros));
  return 8192;
#else
  (void)out; (void)cap;
  return -1;
#endif
}

extern "C" int dusty_conv_role_prof(double *out8, int reset) {     // 12 doubles
#ifdef DUSTY_ROLE_PROF
  unsigned long long h[12];
  if (cudaMemcpyFromSymbol(h, s_role_prof, sizeof(h)) != cudaSuccess) return DUSTY_ECUDA;
  for (int i = 0; i < 12; ++i) out8[i] = (double)h[i];
  if (reset) {
    unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(s_role_prof, z, sizeof(z)) != cudaSuccess) return DUSTY_ECUDA;
  }
  return DUSTY_OK;
#else
  (void)reset;
  for (int i = 0; i < 12; ++i) out8[i] = 0.0;
  return DUSTY_EUNSUPPORTED;
#endif
}

// three filter rows per CTA (shared dY tile): 3-row filters whose N tile is at most 128 wide
static bool wgrad_rows_merged(int O, int R) {
  static const bool off = [] { const char *e = getenv("DUSTY_WGRAD_ROWS"); return e && atoi(e) == 1; }();
  return !off && R == 3 && O <= 128;
}

static bool wgrad_rows_resident() {
  static const bool off = [] { const char *e = getenv("DUSTY_WGRAD_RESIDENT"); return e && atoi(e) == 0; }();
  return !off;
}

static int wgrad_splits(int B, int H_out, int W_out, int C, int O, int R, int S, bool rows) {
  int TW, TH;
  if (rows) {
    TW = W_out >= 32 ? 32 : ((W_out + 7) / 8) * 8;
    TH = 4;
  } else {
    pick_patch(W_out, H_out, kCK, &TW, &TH);
  }
  const long long total_pb = (long long)((W_out + TW - 1) / TW) * ((H_out + TH - 1) / TH) * B;
  const int BN = O > 128 ? 256 : (O > 64 ? 128 : 64);
  const int tile_ctas = ((S * C + kCM - 1) / kCM) * ((O + BN - 1) / BN) * (wgrad_rows_merged(O, R) ? 1 : R);
  long long splits = (2 * num_sms() + tile_ctas - 1) / tile_ctas;
  if (splits > total_pb) splits = total_pb;
  return splits < 1 ? 1 : (int)splits;
}

static bool wgrad_use_workspace() {
  static const bool on = [] { const char *e = getenv("DUSTY_WGRAD_WORKSPACE"); return e && atoi(e) != 0; }();
  return on;
}

extern "C" long long dusty_conv2d_wgrad_tc_workspace(int B, int H_out, int W_out, int C, int O,
                                                     int R, int S) {
  if (!wgrad_use_workspace()) return 0;      // split-K partials are reduced in place (atomics)
  const int splits = wgrad_splits(B, H_out, W_out, C, O, R, S, false);
  return splits > 1 ? (long long)R * S * C * O * splits : 0;
}

extern "C" int dusty_conv2d_wgrad_tc(const void *x, const void *dy, float *dwp, float *ws,
                                     long long ws_elems, int B, int H_in, int W_in, int C,
                                     int H_out, int W_out, int O, int R, int S, int stride_h,
                                     int stride_w, void *stream) {
  DUSTY_CHECK_ARG(x && dy && dwp, "null pointer");
  DUSTY_CHECK_ARG(get_encode() != nullptr, "cuTensorMapEncodeTiled unavailable");
  DUSTY_CHECK_ARG(C % 8 == 0 && O % 8 == 0, "channel counts must be multiples of 8");
  DUSTY_CHECK_ARG(R >= 1 && R <= 4 && S >= 1, "1..4 filter rows");
  DUSTY_CHECK_ARG((long long)(H_out - 1) * stride_h + R <= H_in &&
                      (long long)(W_out - 1) * stride_w + S <= W_in,
                  "window outside the input");
  DUSTY_CHECK_ARG(aligned16(x) && aligned16(dy) && aligned16(dwp), "16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int SC = S * C;
  WgMaps maps;
  WgParams prm;
  const int BN = O > 128 ? 256 : (O > 64 ? 128 : 64);
  // row-resident x patches: 3x3 with unit row stride, TW a multiple of 8 (row shifts stay on the
  // swizzle atom), four image rows per stage (+ 2 halo rows)
  const bool rows = wgrad_rows_merged(O, R) && wgrad_rows_resident() && stride_h == 1 && BN == 64;   // 3 x 64 KiB stages
  if (rows) {
    prm.TW = W_out >= 32 ? 32 : ((W_out + 7) / 8) * 8;
    prm.TH = 4;
  } else {
    pick_patch(W_out, H_out, kCK, &prm.TW, &prm.TH);
  }
  const int kpix = prm.TW * prm.TH;
  prm.rows = rows ? 1 : 0;
  prm.ksteps = kpix / 16;
  const int nr = wgrad_rows_merged(O, R) ? 3 : 1;
  prm.a_half = rows ? (prm.TH + 2) * prm.TW * 128 : 8192;
  prm.a_rr = rows ? prm.TW * 128 : kCABytes;
  prm.a_slot = rows ? 2 * prm.a_half : nr * kCABytes;
  prm.b_blk = kpix * 128;
  prm.b_slot = (BN / 64) * prm.b_blk;
  const uint32_t box[4] = {64u, (uint32_t)prm.TW, (uint32_t)prm.TH, 1u};
  const __nv_bfloat16 *xb = (const __nv_bfloat16 *)x;
  bool ok = true;
  for (int r = 0; r < R; ++r) {
    const uint64_t dims[4] = {(uint64_t)SC, (uint64_t)W_out, (uint64_t)H_out, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)stride_w * C * 2, (uint64_t)stride_h * W_in * C * 2,
                                 (uint64_t)H_in * W_in * C * 2};
    ok = ok && make_map4(&maps.a[r], xb + (long long)r * W_in * C, dims, strides, box);
  }
  for (int r = R; r < 4; ++r) maps.a[r] = maps.a[0];
  if (rows) {
    const uint64_t dims[4] = {(uint64_t)SC, (uint64_t)W_out, (uint64_t)(H_out + 2), (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)stride_w * C * 2, (uint64_t)W_in * C * 2, (uint64_t)H_in * W_in * C * 2};
    const uint32_t rbox[4] = {64u, (uint32_t)prm.TW, (uint32_t)(prm.TH + 2), 1u};
    ok = ok && make_map4(&maps.a[3], xb, dims, strides, rbox);
  }
  {
    const uint64_t dims[4] = {(uint64_t)O, (uint64_t)W_out, (uint64_t)H_out, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)O * 2, (uint64_t)W_out * O * 2, (uint64_t)H_out * W_out * O * 2};
    ok = ok && make_map4(&maps.g, dy, dims, strides, box);
  }
  {
    EncodeTiledFn enc = get_encode();
    cuuint64_t od[3] = {(cuuint64_t)O, (cuuint64_t)SC, (cuuint64_t)R};
    cuuint64_t os[2] = {(cuuint64_t)O * 4, (cuuint64_t)SC * O * 4};
    cuuint32_t ob[3] = {32, 128, 1}, oe[3] = {1, 1, 1};
    ok = ok && enc(&maps.o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dwp, od, os, ob, oe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  if (!ok) {
    set_error("dusty_conv2d_wgrad_tc: cuTensorMapEncodeTiled failed");
    return DUSTY_ECUDA;
  }
  prm.tiles_w = (W_out + prm.TW - 1) / prm.TW;
  prm.tiles_h = (H_out + prm.TH - 1) / prm.TH;
  const long long total_pb = (long long)prm.tiles_w * prm.tiles_h * B;
  DUSTY_CHECK_ARG(total_pb <= 0x7fffffff, "too many pixel blocks");
  prm.total_pb = (int)total_pb;
  prm.MT = (SC + kCM - 1) / kCM;
  prm.NT = (O + BN - 1) / BN;
  prm.SC = SC; prm.O = O;
  const long long n = (long long)R * SC * O;
  int splits = wgrad_splits(B, H_out, W_out, C, O, R, S, rows);
  // Split-K partials are ADDED into dwp with vector reductions (fp32, 16 bytes each) instead of
  // going through a [splits] workspace and a second kernel: the reduce pass cost 19 us per
  // convolution, 0.45 ms per training iteration.  (Summation order is then unordered: results
  // vary in the last fp32 bits from run to run.)  DUSTY_WGRAD_WORKSPACE=1 restores the
  // deterministic two-pass form when the caller provides the workspace.
  const bool atomic = !wgrad_use_workspace() && splits > 1;
  if (!atomic && splits > 1 && (ws == nullptr || ws_elems < n * splits)) {
    splits = ws ? (int)(ws_elems / n) : 1;
    if (splits < 1) splits = 1;
  }
  prm.pb_per_split = (prm.total_pb + splits - 1) / splits;
  splits = (prm.total_pb + prm.pb_per_split - 1) / prm.pb_per_split;
  prm.atomic = (atomic && splits > 1) ? 1 : 0;
  prm.out = prm.atomic ? dwp : (splits > 1 ? ws : dwp);
  prm.split_stride = prm.atomic ? 0 : n;
  if (prm.atomic && cudaMemsetAsync(dwp, 0, sizeof(float) * (size_t)n, st) != cudaSuccess) {
    set_error("dusty_conv2d_wgrad_tc: memset failed");
    return DUSTY_ECUDA;
  }
  int rc;
  if (wgrad_rows_merged(O, R)) {
    rc = BN == 128 ? launch_wgrad<128, 3, 3>(maps, prm, splits, R, st)
                   : launch_wgrad<64, 3, 3>(maps, prm, splits, R, st);
  } else {
    switch (BN) {
      case 256: rc = launch_wgrad<256, 4, 1>(maps, prm, splits, R, st); break;
      case 128: rc = launch_wgrad<128, 4, 1>(maps, prm, splits, R, st); break;
      default: rc = launch_wgrad<64, 4, 1>(maps, prm, splits, R, st); break;
    }
  }
  if (rc) return rc;
  DUSTY_LAUNCH_CHECK();
  if (splits > 1 && !prm.atomic) {
    const long long n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(ws, dwp, n4, splits, n / 4);
    DUSTY_LAUNCH_CHECK();
  }
  return DUSTY_OK;
}

extern "C" int dusty_conv2d_halo_supported(int C, int O, int R, int S) {
  const int BN = O > 64 ? 128 : (O > 32 ? 64 : 32);
  return (C == 32 || C == 64) && O % 8 == 0 && O >= 8 && O <= 128 && R >= 1 && R <= 3 && S >= 1 &&
         S <= 3 && R * S * BN * C * 2 <= 72 * 1024 && get_encode() != nullptr;
}

extern "C" int dusty_conv2d_halo_tc(const void *x, const void *wpk, const float *bias, void *y,
                                    int B, int H_in, int W_in, int C, int H_out, int W_out, int O,
                                    int R, int S, int org_h, int org_w, long long y_off,
                                    long long y_sb, long long y_sh, long long y_sw, int act,
                                    float alpha, float scale, long long w_sn, long long w_sg, int flip,
                                    void *stream) {
  DUSTY_CHECK_ARG(x && wpk && y, "null pointer");
  DUSTY_CHECK_ARG(w_sn >= 0 && w_sg >= 0 && w_sn % 8 == 0 && w_sg % 8 == 0, "weight strides: multiples of 8");
  DUSTY_CHECK_ARG(dusty_conv2d_halo_supported(C, O, R, S), "shape not supported by the halo kernel");
  DUSTY_CHECK_ARG(B > 0 && H_in > 0 && W_in > 0 && H_out > 0 && W_out > 0, "empty tensor");
  DUSTY_CHECK_ARG(aligned16(x) && aligned16(wpk) && aligned16(y), "16-byte alignment");
  DUSTY_CHECK_ARG((y_off % 8 == 0) && (y_sb % 8 == 0) && (y_sh % 8 == 0) && (y_sw % 8 == 0),
                  "output strides must keep 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int ROWB = C * 2;
  const int BN = O > 64 ? 128 : (O > 32 ? 64 : 32);
  HaloParams prm;
  prm.T = R * S; prm.S = S; prm.flip = flip ? 1 : 0;
  // patch pitch: 64 pixels, or 32 when that wastes fewer columns on a narrow image
  auto cover = [&](int pw) { const int two = pw - (S - 1); return (long long)((W_out + two - 1) / two) * pw; };
  prm.PW = cover(32) < cover(64) ? 32 : 64;
  prm.TWo = prm.PW - (S - 1);
  // rows per tile: multiples that make TH * PW % 128 == 0, bounded by the shared memory left
  // for two patches next to the resident filter and the staging tiles, by the accumulators of
  // one TMEM set, and chosen to load the fewest rows (tile rows + halo) over the image
  const int th_step = 128 / prm.PW;
  const int smem_left = 227 * 1024 - 2048 - ((prm.T * BN * ROWB + 1023) & ~1023) - 2 * kCM * BN * 2;
  const int patch_cap = smem_left / 2 - 1024 < 64 * 1024 ? smem_left / 2 - 1024 : 64 * 1024;
  const int mb_cap = 256 / BN;             // accumulators of one TMEM set
  int th = th_step;
  long long best_rows = -1;
  for (int t = th_step; t <= 16 && t * prm.PW / 128 <= mb_cap; t += th_step) {
    if ((t + R - 1) * prm.PW * ROWB + (S - 1) * ROWB > patch_cap || t + R - 1 > 256) break;
    const long long rows = (long long)((H_out + t - 1) / t) * (t + R - 1);
    if (best_rows < 0 || rows <= best_rows) { best_rows = rows; th = t; }
    if (t >= H_out) break;
  }
  DUSTY_CHECK_ARG((th + R - 1) * prm.PW * ROWB + (S - 1) * ROWB <= patch_cap, "filter too large for the halo kernel");
  prm.TH = th;
  prm.MB = prm.TH * prm.PW / 128;
  prm.org_h = org_h; prm.org_w = org_w;
  prm.tiles_w = (W_out + prm.TWo - 1) / prm.TWo;
  prm.tiles_h = (H_out + prm.TH - 1) / prm.TH;
  const long long total = (long long)prm.tiles_w * prm.tiles_h * B;
  DUSTY_CHECK_ARG(total <= 0x7fffffff, "too many tiles");
  prm.total_tiles = (int)total;
  prm.H_out = H_out; prm.W_out = W_out; prm.O = O;
  const int prow = prm.TH + R - 1;
  prm.patch_bytes = prow * prm.PW * ROWB;
  // the last taps of the last M block read up to S - 1 rows past the patch: keep them inside
  // the buffer (their values only reach discarded virtual pixels)
  const int buf_stride = (prm.patch_bytes + (S - 1) * ROWB + 1023) & ~1023;
  prm.y_off = y_off; prm.y_sb = y_sb; prm.y_sh = y_sh; prm.y_sw = y_sw;
  prm.bias = bias; prm.y = (__nv_bfloat16 *)y; prm.act = act; prm.alpha = alpha; prm.scale = scale;
  HaloMaps maps;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W_in, (uint64_t)H_in, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W_in * C * 2, (uint64_t)H_in * W_in * C * 2};
    const uint32_t box[4] = {(uint32_t)C, (uint32_t)prm.PW, (uint32_t)prow, 1u};
    const uint64_t wdims[3] = {(uint64_t)C, (uint64_t)O, (uint64_t)prm.T};
    const uint64_t wstr[2] = {(uint64_t)(w_sn ? w_sn : C) * 2, (uint64_t)(w_sg ? w_sg : (long long)O * C) * 2};
    const uint32_t wbox[3] = {(uint32_t)C, (uint32_t)BN, 1u};
    // output: NHWC view through the caller's strides; one box = the valid pixels of a block
    const int sub_ch = BN > 64 ? 64 : BN;
    const uint64_t ydims[4] = {(uint64_t)O, (uint64_t)W_out, (uint64_t)H_out, (uint64_t)B};
    const uint64_t ystr[3] = {(uint64_t)y_sw * 2, (uint64_t)y_sh * 2, (uint64_t)y_sb * 2};
    const uint32_t ybox[4] = {(uint32_t)sub_ch, (uint32_t)prm.TWo, (uint32_t)(kCM / prm.PW), 1u};
    if (!make_map_sw(&maps.x, x, 4, dims, strides, box, ROWB == 64) ||
        !make_map_sw(&maps.w, wpk, 3, wdims, wstr, wbox, ROWB == 64) ||
        !make_map_sw(&maps.y, (const __nv_bfloat16 *)y + y_off, 4, ydims, ystr, ybox, sub_ch == 32)) {
      set_error("dusty_conv2d_halo_tc: cuTensorMapEncodeTiled failed");
      return DUSTY_ECUDA;
    }
  }
  int rc;
  if (ROWB == 64) {
    if (BN == 32) rc = launch_halo<64, 32, 2>(maps, prm, buf_stride, st);
    else if (BN == 64) rc = launch_halo<64, 64, 2>(maps, prm, buf_stride, st);
    else rc = launch_halo<64, 128, 2>(maps, prm, buf_stride, st);
  } else {
    if (BN == 32) rc = launch_halo<128, 32, 2>(maps, prm, buf_stride, st);
    else if (BN == 64) rc = launch_halo<128, 64, 2>(maps, prm, buf_stride, st);
    else rc = launch_halo<128, 128, 2>(maps, prm, buf_stride, st);
  }
  if (rc) return rc;
  DUSTY_LAUNCH_CHECK();
  return DUSTY_OK;
}

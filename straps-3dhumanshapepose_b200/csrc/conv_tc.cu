// conv_tc.cu -- ResNet-18 encoder on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces the nn.Conv2d -> BatchNorm2d -> (+identity) -> ReLU clusters of the reference's
// models/resnet.py:61-77,201-216 in STRAPS_CONV_F16X3_TC mode.
//
// Numerics: the north-star bar is 1e-4 relative fp32, which a single TF32 or bf16 pass misses.
// Every operand is therefore carried as a 2-term fp16 split x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 11 + 11
// mantissa bits) and each K-block issues three kind::f16 MMAs with fp32 accumulation in TMEM:
//   A_hi.W_hi + A_hi.W_lo + A_lo.W_hi      (dropped lo.lo term and split residual: ~2^-22 relative).
// A first version used a bf16 split (8 + 8 bits): 4e-6 error per layer, 5e-5 at the features after 17 layers -- too
// close to the bar; fp16 has the mantissa and, with the two range guards below, the range:
//   * weights (BatchNorm scale folded in) are multiplied per output channel by a power of two that puts the row
//     maximum in [2^13, 2^14) so that hi AND lo stay fp16-normal; the exact inverse is applied in the epilogue;
//   * activations are clamped to +-65504 before the split (post-BatchNorm ResNet activations are O(1..100)); small
//     activations make `lo` fp16-subnormal, which costs at most 3e-8 ABSOLUTE error.
//
// HBM layout
//   activations   NHWC, two fp16 planes per tensor: [hi: B,H,W,C][lo: B,H,W,C]  (same bytes as fp32)
//   conv1 input   [B][262][264][24] fp16 hi/lo, zero halo of 3 pixels, channels padded 17/18 -> 24 (48 B / pixel):
//                 for one filter row kh the 7 taps x 24 channels of an output pixel are ONE contiguous run of 168
//                 elements, read as three 64-element K-chunks (the 8th "tap" of each run has zero weights).
//   weights       [Cout][K_eff] fp16 hi/lo, K ordered (kh, kw, ci) to match the A operand, eval-mode BatchNorm scale
//                 folded in before the split; BN shift stays fp32 and is added in the epilogue.
// Implicit GEMM: M = output pixels (128 per tile = whole output rows, so a tile is a TMA box of the NHWC input),
//   N = Cout tile (64 or 128), K = taps x Cin in blocks of 64 (one SWIZZLE_128B row).  Padding comes from TMA
//   out-of-bounds zero fill (negative start coordinates), stride 2 from the tensor map's element strides.
// Kernel: persistent, warp specialised -- warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread),
//   warps 2..5 = epilogue (tcgen05.ld -> +shift, +residual, ReLU -> fp16 split -> vector stores).  Shared-memory
//   ring of 3-4 stages, two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
// Roofline: tensor pipe; algorithmic FLOPs 6,180.1 MFLOP/body (C=17), issued = 3 passes (+ conv1 K padding).
#include "regressor.h"
#include <map>
#include <cstdlib>

namespace straps {

constexpr int XP_H = 262, XP_W = 264, XP_C = 24;   // padded conv1 input
constexpr int C1_KROW = 192;                       // K elements per filter row of conv1 (3 chunks of 64)
constexpr int TC_THREADS = 192;
// STRAPS_TC_EXPERIMENTS builds keep the in-situ timing switches of profiles/r01_conv_experiments.txt (p.debug); normal builds have
// no trace of them in the hot loops
// STRAPS_TC_DEBUG is a bit mask: 1 = no TMA traffic (the producer only arrives on the full barriers, MMAs run on stale operands),
// 2 = no MMAs (the issuer only commits), 4 = no epilogue global I/O.  3 / 5 / 6 leave ONE of the three parts running alone.
#ifdef STRAPS_TC_EXPERIMENTS
#define TC_EPI_IO(p) (((p).debug & 4) == 0)
#define TC_NO_TMA(p) (((p).debug & 1) != 0)
#define TC_NO_MMA(p) (((p).debug & 2) != 0)
#else
#define TC_EPI_IO(p) true
#define TC_NO_TMA(p) false
#define TC_NO_MMA(p) false
#endif
constexpr int BM_TC = 128;
constexpr int BK_TC = 64;
constexpr int TC_DEFAULT_BK = 64;   // K elements per stage of conv_tc_kernel unless STRAPS_TC_BK says otherwise

struct TcConvParams {
  int n_mtiles, n_ntiles, n_kblocks;
  int conv1;           // 1 = flattened (kw,c) runs of the padded input
  int cchunks;         // Cin / 64
  int kw_count;        // filter width
  int stride, pad;
  int hw_out;          // Hout*Wout
  int wout, th;        // tile = th rows of wout pixels (x nb images)
  int cout;
  long long m_total;   // B*Hout*Wout
  const float* shift;  // [cout]
  const float* unscale;   // [cout] exact inverse of the per-channel power-of-two weight scale
  // outputs
  float* out_f32;                  // NHWC fp32 (or null)
  __half* out_hi;           // NHWC split planes (or null)
  __half* out_lo;
  const __half* res_hi;     // residual (identity) planes or null
  const __half* res_lo;
  const float* res_f32;     // fp32 NHWC tensor added to the result (data-gradient accumulation) or null
  int relu;
  int debug;        // timing experiments only (STRAPS_TC_DEBUG bit mask): 1 = no TMA traffic, 2 = no MMAs, 4 = no epilogue global I/O
};

// BN = output-channel tile, MT = number of 128-pixel M-tiles that share one weight tile per K-block.
// Measured (tools/tc_probe.cu, profiles/r01_tc_probe.txt): an SS-mode tcgen05.mma M128xNx16 costs 48 / 64 / 128 cycles
// for N = 64 / 128 / 256, i.e. it streams its operands from shared memory at the full 128 B/clk for N <= 128, so
// the TMA writes of the next stages and the MMA operand reads share one 128 B/clk port and the kernel is
// SHARED-MEMORY-BANDWIDTH bound:  cycles per K-block ~ (MMA operand bytes + TMA bytes) / 128.
//   (64,1): 56 + 48 KB -> 812 cycles  (MMA alone 448)      (128,1): 80 + 64 KB -> 1125 cycles (MMA alone 768)
// Wider tiles -- (128,2), (256,1) -- need fewer bytes per FLOP but only fit 2 ring stages and a single TMEM
// accumulator stage; they measured 10-25 % SLOWER (profiles/r01_launches_tilecfg.txt) and are kept only as
// template instantiations for the next round (2-CTA pairs halve the weight bytes per SM).
// BK = K elements per pipeline stage: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows).  The ring is LATENCY bound -- measured
// L2 -> SM throughput = (STAGES - 1) * STAGE_BYTES / ~2600 cycles for every tile shape (profiles/r01_conv_experiments.txt) -- and the
// stage being consumed carries no load, so half-size stages put more of the same shared memory in flight:
//   (64,1):  4 x 48 KB -> 144 KB in flight;   9 x 24 KB -> 192 KB        (128,1): 3 x 64 KB -> 128 KB;   7 x 32 KB -> 192 KB
template <int BN, int MT, int BK = 64>
struct TcCfg {
  static_assert(BK == 64 || BK == 32, "BK: one swizzle row of 128 or 64 bytes");
  static constexpr int A_BYTES = BM_TC * BK * 2;        // one plane of one A tile
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = MT * 2 * A_BYTES + 2 * W_BYTES;
  static constexpr int MAX_STAGES = (BK == 64) ? 4 : 10;
  static constexpr int BUDGET = (BK == 64) ? 220 * 1024 : (232448 - 1024 - 256);
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > MAX_STAGES ? MAX_STAGES : (BUDGET / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static constexpr int TILE_COLS = 2 * BN;              // per M-tile: [hi.hi | lo terms]
  static constexpr int ACC_COLS = MT * TILE_COLS;       // per accumulator stage
  static constexpr int TSTAGES = (2 * ACC_COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = TSTAGES * ACC_COLS;
  static constexpr uint32_t DESC_HI = (BK == 64) ? UMMA_DESC_SW128_HI : UMMA_DESC_SW64_HI;
  static constexpr int C1_CHUNKS = C1_KROW / BK;        // K-blocks per filter row of conv1
  // conv1: the 168 real elements of a filter row end inside the last chunk; its all-padding K=16 steps are skipped
  static constexpr int C1_LAST_KSTEPS = (7 * XP_C - (C1_CHUNKS - 1) * BK + 15) / 16;
  static_assert(STAGES >= 2 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "bad tile configuration");
  static_assert(2 * STAGES * 8 + 40 <= 256, "barrier block too small");
};

// CL > 1: clusters of CL CTAs along M work on CL consecutive M-tiles of the SAME N-tile in lock step and share the weight tile:
// each CTA fetches 1/CL of its rows and multicasts them to the whole cluster (L2 -> SM weight traffic / CL).  A stage is reused
// only after EVERY CTA of the cluster has consumed it (empty barriers count CL arrivals: each issuer's commit is multicast).
// EPW = epilogue warps: 4 (one per TMEM lane quadrant) or 8 (two per quadrant, alternating 32-column chunks).  On the short-K layers
// the epilogue of a tile is as long as its main loop (layer1: MMAs alone 53 us of 68) and a single warp per scheduler hides no latency.
// PDL = launched with programmatic stream serialization (STRAPS_TC_PDL=1): the CTA lets the next kernel of the stream start as soon as
// this grid's CTAs have begun (griddepcontrol.launch_dependents), and does its own set-up (barriers, TMEM allocation, descriptor
// prefetch) BEFORE it waits for the previous kernel's results (griddepcontrol.wait) -- launch latency and prologue move under the
// previous kernel's tail.  Not yet run on hardware; the default instantiations carry neither instruction.
// M2 (STRAPS_TC_TMA2=1) = two TMA operations per stage instead of four, with NO change of the data in HBM: the hi and lo planes of a
// tensor are two allocations a fixed distance apart, so "plane" is simply one more (outermost) tensor-map dimension -- a 5-D box
// {64 ch, w, h, images, 2 planes} lands as [A_hi tile][A_lo tile] and a 3-D box {64 K, BN rows, 2 planes} as [W_hi][W_lo], exactly the
// stage layout the MMAs read.  map_a_hi / map_w_hi then carry the merged maps.  Tests the per-operation cost of item 3 in DESIGN.md 4.2.
template <int BN, int MT, int CL = 1, int BK = 64, int EPW = 4, bool PDL = false, bool M2 = false>
__global__ void __launch_bounds__(64 + 32 * EPW, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const TcConvParams p) {
  using Cfg = TcCfg<BN, MT, BK>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]
  uint64_t* tfull = bars + 2 * Cfg::STAGES;    // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  static_assert(CL == 1 || (MT == 1 && (CL == 2 || CL == 4) && BN / CL >= 8), "cluster variant: one M-tile per CTA");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int GROUP = (CL > 1) ? CL : MT;                 // M-tiles per work item (of the cluster, or of the CTA)
  const int n_groups = (p.n_mtiles + GROUP - 1) / GROUP;    // a work item = GROUP consecutive M-tiles x one N-tile
  const int n_items = n_groups * p.n_ntiles;
  const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int item0 = (CL > 1) ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
  const int item_step = (CL > 1) ? (int)(gridDim.x / CL) : (int)gridDim.x;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 32 * EPW); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();      // every CTA's barriers exist before any multicast load / remote commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if constexpr (PDL) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");      // every thread: activations, residuals and outputs are touched only after this
  }

  // The two issue roles are ONE thread each, so their cost is instruction LATENCY, not throughput: the first version of these
  // loops (div/mod per K-block, 64-bit descriptor arithmetic, `lane == 0` divergence that made the compiler wrap every
  // UTMALDG / UTCHMMA in an ELECT + BRA.U.ANY loop) ran ~125-150 dependent SASS instructions per K-block -- about as long as the
  // 448 tensor-pipe cycles of the K-block itself (in-situ: MMAs alone 810 cycles per K-block, TMA alone 870).  Both loops now
  // keep incremental counters, 32-bit shared addresses and descriptor words, and run under elect.sync.
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= TMA producer =================
      uint32_t st = 0, ph = 1;                       // ring stage and the parity to wait for on empty[st]
      const int nkb = p.n_kblocks, cchunks = p.cchunks, kwc = p.kw_count, pad = p.pad, stride = p.stride;
      const bool conv1 = p.conv1 != 0;
      for (int item = item0; item < n_items; item += item_step) {
        const int mg = item / p.n_ntiles, nt = item - mg * p.n_ntiles;
        int b0[MT], row0[MT];
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const long long pix0 = (long long)(mg * GROUP + crank + t) * BM_TC;   // tiles past the end land out of bounds -> zeros
          b0[t] = (int)(pix0 / p.hw_out);
          const int oh0 = (int)((pix0 % p.hw_out) / p.wout);
          row0[t] = conv1 ? oh0 : oh0 * stride - pad;
        }
        const int wrow = nt * BN;
        // K-block coordinates, advanced incrementally: conv1 (c0 = kh * row pitch + j * 64, j = 0..2), others (cc, kw, kh)
        int c0 = 0, c1 = conv1 ? 0 : -pad, dh = 0, sub = 0, kwi = 0, wk = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t sa = smem0 + st * Cfg::STAGE_BYTES;
          const uint32_t fb = full0 + st * 8;
          mbar_wait_u32(empty0 + st * 8, ph);
#ifdef STRAPS_TC_EXPERIMENTS
          if (TC_NO_TMA(p)) { mbar_arrive(&full[st]); if (++st == Cfg::STAGES) { st = 0; ph ^= 1; } continue; }
#endif
          mbar_expect_tx_u32(fb, Cfg::STAGE_BYTES);
          if constexpr (M2) {
            static_assert(!M2 || (CL == 1 && BK == 64), "merged-plane boxes: single-CTA SWIZZLE_128B stages only");
#pragma unroll
            for (int t = 0; t < MT; ++t) tma_load_5d_u32(sa + (2 * t) * Cfg::A_BYTES, &map_a_hi, fb, c0, c1, row0[t] + dh, b0[t], 0);
            tma_load_3d_u32(sa + MT * 2 * Cfg::A_BYTES, &map_w_hi, fb, wk, wrow, 0);
          } else {
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            tma_load_4d_u32(sa + (2 * t) * Cfg::A_BYTES, &map_a_hi, fb, c0, c1, row0[t] + dh, b0[t]);
            tma_load_4d_u32(sa + (2 * t + 1) * Cfg::A_BYTES, &map_a_lo, fb, c0, c1, row0[t] + dh, b0[t]);
          }
          }
          if constexpr (M2) {
          } else if constexpr (CL > 1) {
            // this CTA's 1/CL of the weight rows, delivered to every CTA of the cluster (the W maps have BN / CL-row boxes)
            const uint32_t wo = sa + MT * 2 * Cfg::A_BYTES + crank * (Cfg::W_BYTES / CL);
            tma_load_2d_mc_u32(wo, &map_w_hi, fb, wk, wrow + crank * (BN / CL), CMASK);
            tma_load_2d_mc_u32(wo + Cfg::W_BYTES, &map_w_lo, fb, wk, wrow + crank * (BN / CL), CMASK);
          } else {
            tma_load_2d_u32(sa + MT * 2 * Cfg::A_BYTES, &map_w_hi, fb, wk, wrow);
            tma_load_2d_u32(sa + MT * 2 * Cfg::A_BYTES + Cfg::W_BYTES, &map_w_lo, fb, wk, wrow);
          }
          wk += BK;
          if (++st == Cfg::STAGES) { st = 0; ph ^= 1; }
          if (conv1) {
            c0 += BK;
            if (++sub == Cfg::C1_CHUNKS) { sub = 0; c0 += XP_W * XP_C - C1_KROW; }     // next filter row of the padded input
          } else {
            c0 += BK;
            if (++sub == cchunks) {
              sub = 0; c0 = 0; ++c1;
              if (++kwi == kwc) { kwi = 0; c1 = -pad; ++dh; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer =================
      // W_hi and W_lo are adjacent in the stage, so for BN <= 128 ONE MMA of width 2*BN computes
      // A_hi.[W_hi ; W_lo] into [acc_hi | acc_lo] (A_hi is read from shared memory once instead of twice).
      // Two accumulators per tile: tensor-core fp32 accumulation truncates (measured -1e-8 relative per MMA, 5.6e-5
      // at the features with a single accumulator), so the small lo-terms get their own accumulator and only
      // K/16 additions happen at full magnitude; the two are summed in fp32 in the epilogue.
      constexpr uint32_t idesc = umma_idesc_f16(BM_TC, BN);
      constexpr uint32_t idesc_wide = umma_idesc_f16(BM_TC, (2 * BN <= 256) ? 2 * BN : BN);
      constexpr uint32_t STAGE16 = Cfg::STAGE_BYTES >> 4, A16 = Cfg::A_BYTES >> 4, W16 = Cfg::W_BYTES >> 4;
      const uint32_t desc0 = umma_desc_sw128_lo(smem0);           // low descriptor word of stage 0, A tile 0, hi plane
      const uint32_t tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      uint32_t st = 0, ph = 0, as = 0, aph = 1;
      const int nkb = p.n_kblocks;
      const bool conv1 = p.conv1 != 0;
      for (int item = item0; item < n_items; item += item_step) {
        mbar_wait_u32(tempty0 + as * 8, aph);
        tc_fence_after();
        const uint32_t acc = tmem_base + as * Cfg::ACC_COLS;
        int sub = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_u32(full0 + st * 8, ph);
          tc_fence_after();
          const uint32_t d0 = desc0 + st * STAGE16;
          const uint32_t w_hi = d0 + MT * 2 * A16, w_lo = w_hi + W16;
          // conv1: the last chunk of a filter row (192 elements for 7 taps x 24 channels = 168 real ones) ends in zero-weight
          // padding; its all-padding K=16 steps are skipped (11 of 12 MMAs per filter row)
          int ksteps = BK / 16;
          if (conv1) { if (++sub == Cfg::C1_CHUNKS) { sub = 0; ksteps = Cfg::C1_LAST_KSTEPS; } }
#ifdef STRAPS_TC_EXPERIMENTS
          if (TC_NO_MMA(p)) ksteps = 0;                                   // timing experiment: no MMAs
#endif
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            const uint32_t a_hi = d0 + (2 * t) * A16, a_lo = a_hi + A16;
            const uint32_t d_hi = acc + t * Cfg::TILE_COLS, d_lo = d_hi + BN;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (k < ksteps) {
                const uint32_t ko = k * 2;                 // +32 bytes along K inside the 128-byte swizzle row
                const uint32_t accum = (kb | k) != 0;
                if constexpr (2 * BN <= 256) {
                  umma_f16_lohi(d_hi, a_hi + ko, w_hi + ko, idesc_wide, accum, Cfg::DESC_HI);   // -> [acc_hi | acc_lo]
                  umma_f16_lohi(d_lo, a_lo + ko, w_hi + ko, idesc, 1, Cfg::DESC_HI);
                } else {
                  umma_f16_lohi(d_hi, a_hi + ko, w_hi + ko, idesc, accum, Cfg::DESC_HI);
                  umma_f16_lohi(d_lo, a_hi + ko, w_lo + ko, idesc, accum, Cfg::DESC_HI);
                  umma_f16_lohi(d_lo, a_lo + ko, w_hi + ko, idesc, 1, Cfg::DESC_HI);
                }
              }
            }
          }
          // frees the smem stage once these MMAs have read it (in every CTA of the cluster: peers multicast into it)
          if constexpr (CL > 1) umma_commit_mc_u32(empty0 + st * 8, CMASK);
          else umma_commit_u32(empty0 + st * 8);
          if (++st == Cfg::STAGES) { st = 0; ph ^= 1; }
        }
        umma_commit_u32(tfull0 + as * 8);           // accumulators complete -> epilogue
        if (++as == Cfg::TSTAGES) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps 2..5 (EPW = 8: 2..9) =================
    static_assert(EPW == 4 || EPW == 8, "one or two epilogue warps per TMEM lane quadrant");
    const int quad = warp & 3;                   // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    constexpr int NCHUNK = MT * (BN / 32);       // 32-column chunks per work item
    constexpr int CSTEP = EPW / 4;               // the warps of a quadrant take alternate chunks
    const int part = (EPW > 4) ? ((warp - 2) >> 2) : 0;
    uint32_t ti = 0;
    for (int item = item0; item < n_items; item += item_step, ++ti) {
      const int mg = item / p.n_ntiles, nt = item % p.n_ntiles;
      const uint32_t as = ti % Cfg::TSTAGES;
      // The residual (identity) tile does not depend on the accumulator: its loads are software-pipelined one chunk ahead, and
      // the first chunk is requested BEFORE waiting for the MMAs -- the in-situ experiment (profiles/r01_conv_experiments.txt)
      // showed the un-pipelined residual latency exposed on the short-K layers.
      uint4 rh[4], rl[4];
      auto fetch_residual = [&](int chunk) {
        const int t = chunk / (BN / 32), c0 = (chunk % (BN / 32)) * 32;
        const long long m = (long long)(mg * GROUP + crank + t) * BM_TC + row;
        if (p.res_hi && chunk < NCHUNK && m < p.m_total && TC_EPI_IO(p)) {
          const size_t o = (size_t)m * p.cout + (size_t)nt * BN + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            rh[q] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o) + q);
            rl[q] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + o) + q);
          }
        }
      };
      fetch_residual(part);
      mbar_wait(&tfull[as], (ti / Cfg::TSTAGES) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = part; chunk < NCHUNK; chunk += CSTEP) {
        const int t = chunk / (BN / 32), c0 = (chunk % (BN / 32)) * 32;
        const long long m = (long long)(mg * GROUP + crank + t) * BM_TC + row;
        const bool valid = m < p.m_total;
        const size_t obase = (size_t)m * p.cout + (size_t)nt * BN;
        uint32_t v[32], vl[32];
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + as * Cfg::ACC_COLS + t * Cfg::TILE_COLS + c0;
        tmem_ld_32x32(tacc, v);
        tmem_ld_32x32(tacc + BN, vl);
        tmem_ld_wait();
        float y[32];
        const float4* sh4 = reinterpret_cast<const float4*>(p.shift + nt * BN + c0);
        const float4* us4 = reinterpret_cast<const float4*>(p.unscale + nt * BN + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = __ldg(sh4 + q), u4 = __ldg(us4 + q);
          y[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]) + __uint_as_float(vl[q * 4 + 0]), u4.x, s4.x);
          y[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]) + __uint_as_float(vl[q * 4 + 1]), u4.y, s4.y);
          y[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]) + __uint_as_float(vl[q * 4 + 2]), u4.z, s4.z);
          y[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]) + __uint_as_float(vl[q * 4 + 3]), u4.w, s4.w);
        }
        if (valid && TC_EPI_IO(p)) {
          if (p.res_hi) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t hw[4] = {rh[q].x, rh[q].y, rh[q].z, rh[q].w}, lw[4] = {rl[q].x, rl[q].y, rl[q].z, rl[q].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                y[q * 8 + e * 2 + 0] += f16lo_to_f(hw[e]) + f16lo_to_f(lw[e]);
                y[q * 8 + e * 2 + 1] += f16hi_to_f(hw[e]) + f16hi_to_f(lw[e]);
              }
            }
          }
        }
        fetch_residual(chunk + CSTEP);    // in flight while this chunk is split and stored
        if (valid && TC_EPI_IO(p)) {
          if (p.res_f32) {
            const float4* r4 = reinterpret_cast<const float4*>(p.res_f32 + obase + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 a = r4[q];
              y[q * 4 + 0] += a.x; y[q * 4 + 1] += a.y; y[q * 4 + 2] += a.z; y[q * 4 + 3] += a.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaxf(y[i], 0.f);
          }
          if (p.out_f32) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + obase + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
          } else {
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              __half h0, l0, h1, l1;
              split_f16(y[2 * i], h0, l0);
              split_f16(y[2 * i + 1], h1, l1);
              ph[i] = pack_f16(h0, h1);
              pl[i] = pack_f16(l0, l1);
            }
            uint4* oh = reinterpret_cast<uint4*>(p.out_hi + obase + c0);
            uint4* ol = reinterpret_cast<uint4*>(p.out_lo + obase + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              oh[q] = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
              ol[q] = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();      // no CTA leaves while a peer may still multicast into it / commit to it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Halo variant for the stride-1 3x3 convolutions with Cin = Cout = BN (layer1: 64x64x64, layer2: 32x32x128) -- STRAPS_TC_HALO=1.
// Run on B200 once, at the very end of round 1 (tools/halo_check.py, profiles/r01_halo_check.json): CORRECT on the first run --
// layer1 bit-identical to conv_tc_kernel (one channel chunk: same K order), layer2 6e-7 (chunk-major instead of tap-major K order) --
// but NOT faster: encoder 1.584 ms shipped, 1.604 with layer1 on this kernel, 1.611 with layer2, 1.676 with both.  Halving the bytes
// TMA brings into the SM therefore does not shorten these layers: their epilogue (16-byte per-thread stores at a 128 / 256-byte
// stride) is at least an equal limit.  Kept off by default as the vehicle for the round-2 epilogue work.
//
// conv_tc_kernel fetches one 128-pixel A tile per filter tap: 9 x 32 KB per 64-channel chunk, and the kernel is paced by the bytes
// TMA brings INTO the SM.  Here the outputs of an image are enumerated on a virtual zero-padded raster of
// (H+2) x (W+2) positions, an M-tile is 128 CONSECUTIVE raster positions p0 .. p0+127, and per 64-channel chunk ONE box
// {64 channels, W+2 pixels, RH rows} is loaded from the UNPADDED NHWC planes, starting at pixel -1 / image row r0-2: TMA's
// out-of-bounds zero fill materialises the border, so shared memory holds raster rows r0 .. r0+RH-1, one 128-byte line per position.
// Tap (kh, kw) is the 128-line window starting at line (p0 - r0 (W+2)) + (kh-1)(W+2) + (kw-1): the same UMMA descriptor with a
// shifted start address.  Border positions are computed and discarded by the epilogue.  Bytes into the SM per image:
// layer1 13.8 -> 7.7 MB, layer2 9.2 -> 6.3 MB.  Two rings: AS halo stages (hi + lo plane) and WS weight stages (one per tap).
// ---------------------------------------------------------------------------------------------------------
struct HaloParams {
  int n_items;            // batch * tiles_per_image
  int tiles_per_image;    // tiles that contain at least one interior position
  int H, W;               // input = output size
  int cchunks;            // Cin / 64
  int cout;
  const float* shift;
  const float* unscale;
  __half* out_hi;
  __half* out_lo;
  const __half* res_hi;   // residual (identity) planes or null
  const __half* res_lo;
  int relu;
  int debug;              // STRAPS_TC_DEBUG bit mask (experiment builds only)
};

template <int BN, int RH, int WP>
struct HaloCfg {
  static constexpr int A_LINES = RH * WP;                                    // raster positions held per plane
  static constexpr int A_BOX_BYTES = A_LINES * 128;                           // what one TMA box delivers
  static constexpr int A_PLANE = (A_BOX_BYTES + 1023) / 1024 * 1024;          // planes start on swizzle-atom boundaries
  static constexpr int A_STAGE = 2 * A_PLANE;
  static constexpr int AS = 2;
  static constexpr int W_BYTES = BN * 128;
  static constexpr int W_STAGE = 2 * W_BYTES;                                 // [W_hi ; W_lo] adjacent: one wide MMA reads both
  static constexpr int WS = 3;
  static constexpr int W_OFF = AS * A_STAGE;
  static constexpr int SMEM_BYTES = AS * A_STAGE + WS * W_STAGE + 1024 + 256;
  static constexpr int TILE_COLS = 2 * BN;
  static constexpr int TMEM_COLS = 2 * TILE_COLS;
  // every tap window [start, start + 128) lies inside the box: the tile's positions and their 3x3 neighbourhoods span
  // 128 + 2 WP + 2 consecutive positions, which touch at most ceil(that / WP) + 1 raster rows
  static_assert(RH * WP >= 128 + 2 * WP + 2 + WP - 1, "halo box too small");
  static_assert(WP <= 256 && RH <= 256, "TMA box dimensions are limited to 256");
  static_assert(SMEM_BYTES <= 232448 && TMEM_COLS <= 512, "halo tile configuration does not fit");
};

template <int BN, int RH, int WP, int EPW = 4>
__global__ void __launch_bounds__(64 + 32 * EPW, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const HaloParams p) {
  using Cfg = HaloCfg<BN, RH, WP>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::AS * Cfg::A_STAGE + Cfg::WS * Cfg::W_STAGE);
  uint64_t* a_full = bars;                          // [AS]
  uint64_t* a_empty = a_full + Cfg::AS;             // [AS]
  uint64_t* w_full = a_empty + Cfg::AS;             // [WS]
  uint64_t* w_empty = w_full + Cfg::WS;             // [WS]
  uint64_t* tfull = w_empty + Cfg::WS;              // [2]
  uint64_t* tempty = tfull + 2;                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::AS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < Cfg::WS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 32 * EPW); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty), w_full0 = smem_u32(w_full), w_empty0 = smem_u32(w_empty);
  const int tpi = p.tiles_per_image, cchunks = p.cchunks;

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= TMA producer: per 64-channel chunk one halo box (2 planes), then its 9 weight tiles =================
      uint32_t as = 0, aph = 1, ws = 0, wph = 1;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int b = item / tpi, t = item - b * tpi;
        const int p0 = t * BM_TC;
        const int r0 = (p0 + WP - 1) / WP - 2;            // = floor((p0 - WP - 1) / WP): raster row of box row 0 (-2 for the first tile)
        for (int cc = 0; cc < cchunks; ++cc) {
          const uint32_t sa = smem0 + as * Cfg::A_STAGE, fb = a_full0 + as * 8;
          mbar_wait_u32(a_empty0 + as * 8, aph);
          if (TC_NO_TMA(p)) {
            mbar_arrive(&a_full[as]);
          } else {
            mbar_expect_tx_u32(fb, 2 * Cfg::A_BOX_BYTES);
            // image coordinates of the box origin: pixel -1 (left border), row r0 - 1 (raster row r = image row r - 1)
            tma_load_4d_u32(sa, &map_a_hi, fb, cc * 64, -1, r0 - 1, b);
            tma_load_4d_u32(sa + Cfg::A_PLANE, &map_a_lo, fb, cc * 64, -1, r0 - 1, b);
          }
          if (++as == Cfg::AS) { as = 0; aph ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t sw = smem0 + Cfg::W_OFF + ws * Cfg::W_STAGE, wb = w_full0 + ws * 8;
            mbar_wait_u32(w_empty0 + ws * 8, wph);
            if (TC_NO_TMA(p)) {
              mbar_arrive(&w_full[ws]);
            } else {
              mbar_expect_tx_u32(wb, Cfg::W_STAGE);
              const int wk = (tap * cchunks + cc) * BK_TC;  // K offset of (tap, chunk) in the [Cout][(kh, kw, ci)] weight rows
              tma_load_2d_u32(sw, &map_w_hi, wb, wk, 0);
              tma_load_2d_u32(sw + Cfg::W_BYTES, &map_w_lo, wb, wk, 0);
            }
            if (++ws == Cfg::WS) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer: 9 taps = 9 row-shifted windows of the same halo tile =================
      constexpr uint32_t idesc = umma_idesc_f16(BM_TC, BN);
      constexpr uint32_t idesc_wide = umma_idesc_f16(BM_TC, 2 * BN);
      static_assert(2 * BN <= 256, "the merged [W_hi ; W_lo] MMA needs N <= 256");
      const uint32_t tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      uint32_t as = 0, aph = 0, ws = 0, wph = 0, acs = 0, acph = 1;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int t = item % tpi;
        const int p0 = t * BM_TC;
        const int r0 = (p0 + WP - 1) / WP - 2;
        const int base = p0 - r0 * WP;                    // line of the tile's first position inside the box (>= WP + 1)
        mbar_wait_u32(tempty0 + acs * 8, acph);
        tc_fence_after();
        const uint32_t d_hi = tmem_base + acs * Cfg::TILE_COLS, d_lo = d_hi + BN;
        uint32_t accum = 0;
        for (int cc = 0; cc < cchunks; ++cc) {
          mbar_wait_u32(a_full0 + as * 8, aph);
          tc_fence_after();
          const uint32_t a0 = umma_desc_sw128_lo(smem0 + as * Cfg::A_STAGE);
          int start = base - WP - 1;                      // tap (0, 0)
          for (int kh = 0; kh < 3; ++kh, start += WP - 3)
            for (int kw = 0; kw < 3; ++kw, ++start) {
              mbar_wait_u32(w_full0 + ws * 8, wph);
              tc_fence_after();
              const uint32_t a_hi = a0 + (uint32_t)start * 8u;          // 128 bytes per line = 8 descriptor units
              const uint32_t a_lo = a_hi + (Cfg::A_PLANE >> 4);
              const uint32_t w_hi = umma_desc_sw128_lo(smem0 + Cfg::W_OFF + ws * Cfg::W_STAGE);
              if (!TC_NO_MMA(p)) {
#pragma unroll
                for (int k = 0; k < BK_TC / 16; ++k) {
                  umma_f16_lohi(d_hi, a_hi + 2 * k, w_hi + 2 * k, idesc_wide, accum);      // A_hi.[W_hi ; W_lo] -> [acc_hi | acc_lo]
                  umma_f16_lohi(d_lo, a_lo + 2 * k, w_hi + 2 * k, idesc, 1);               // A_lo.W_hi
                  accum = 1;
                }
              }
              umma_commit_u32(w_empty0 + ws * 8);
              if (++ws == Cfg::WS) { ws = 0; wph ^= 1; }
            }
          umma_commit_u32(a_empty0 + as * 8);
          if (++as == Cfg::AS) { as = 0; aph ^= 1; }
        }
        umma_commit_u32(tfull0 + acs * 8);
        if (++acs == 2) { acs = 0; acph ^= 1; }
      }
    }
  } else {
    // ================= epilogue: raster position -> pixel, border positions discarded =================
    static_assert(EPW == 4 || EPW == 8, "one or two epilogue warps per TMEM lane quadrant");
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    constexpr int NCHUNK = BN / 32;
    constexpr int CSTEP = EPW / 4;
    const int part = (EPW > 4) ? ((warp - 2) >> 2) : 0;
    uint32_t ti = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++ti) {
      const int b = item / tpi, t = item - b * tpi;
      const int pos = t * BM_TC + row;
      const int py = pos / WP, px = pos - py * WP;
      const bool valid = py >= 1 && py <= p.H && px >= 1 && px <= p.W;
      const size_t obase = (((size_t)b * p.H + (py - 1)) * p.W + (px - 1)) * p.cout;   // only used when valid
      const uint32_t acs = ti & 1;
      uint4 rh[4], rl[4];
      auto fetch_residual = [&](int chunk) {
        if (p.res_hi && chunk < NCHUNK && valid && TC_EPI_IO(p)) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            rh[q] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + obase + chunk * 32) + q);
            rl[q] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + obase + chunk * 32) + q);
          }
        }
      };
      fetch_residual(part);
      mbar_wait(&tfull[acs], (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = part; chunk < NCHUNK; chunk += CSTEP) {
        const int c0 = chunk * 32;
        uint32_t v[32], vl[32];
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + acs * Cfg::TILE_COLS + c0;
        tmem_ld_32x32(tacc, v);
        tmem_ld_32x32(tacc + BN, vl);
        tmem_ld_wait();
        float y[32];
        const float4* sh4 = reinterpret_cast<const float4*>(p.shift + c0);
        const float4* us4 = reinterpret_cast<const float4*>(p.unscale + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = __ldg(sh4 + q), u4 = __ldg(us4 + q);
          y[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]) + __uint_as_float(vl[q * 4 + 0]), u4.x, s4.x);
          y[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]) + __uint_as_float(vl[q * 4 + 1]), u4.y, s4.y);
          y[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]) + __uint_as_float(vl[q * 4 + 2]), u4.z, s4.z);
          y[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]) + __uint_as_float(vl[q * 4 + 3]), u4.w, s4.w);
        }
        if (valid && p.res_hi && TC_EPI_IO(p)) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t hw[4] = {rh[q].x, rh[q].y, rh[q].z, rh[q].w}, lw[4] = {rl[q].x, rl[q].y, rl[q].z, rl[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              y[q * 8 + e * 2 + 0] += f16lo_to_f(hw[e]) + f16lo_to_f(lw[e]);
              y[q * 8 + e * 2 + 1] += f16hi_to_f(hw[e]) + f16hi_to_f(lw[e]);
            }
          }
        }
        fetch_residual(chunk + CSTEP);
        if (valid && TC_EPI_IO(p)) {
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaxf(y[i], 0.f);
          }
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            __half h0, l0, h1, l1;
            split_f16(y[2 * i], h0, l0);
            split_f16(y[2 * i + 1], h1, l1);
            ph[i] = pack_f16(h0, h1);
            pl[i] = pack_f16(l0, l1);
          }
          uint4* oh = reinterpret_cast<uint4*>(p.out_hi + obase + c0);
          uint4* ol = reinterpret_cast<uint4*>(p.out_lo + obase + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            oh[q] = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
            ol[q] = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acs]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Halo kernel, second configuration (STRAPS_TC_HALO=2, layer1 only): MT = 2 consecutive 128-position tiles per work item share ONE
// halo box (7 raster rows) and every weight stage, with a 5-stage weight ring and a single halo stage.
// WRITTEN AFTER THE ROUND'S GPU BUDGET WAS SPENT; a separate kernel so that the hardware-verified conv_halo_kernel stays untouched.
//
// Why: conv_halo_kernel was correct but not faster.  Reading (a) of that result: per tile it still streams all
// 9 weight tiles (144 KB on layer1) through a ring that keeps only 2 x 16 KB in flight, and with ~2600 cycles of loaded TMA latency that
// alone is ~11,000 cycles per tile.  Here a weight stage feeds the MMAs of two tiles (72 KB of W per tile) and four stages are in
// flight while one is consumed; the price is a single-buffered halo box (118 KB), whose load is exposed once per two tiles.
// If reading (a) is right this is ~5,500 cycles per tile instead of ~9,700; if the epilogue is the limit (reading (b)) it is not
// faster -- one run decides.
// ---------------------------------------------------------------------------------------------------------
template <int BN, int RH, int WP, int MT, int NAS, int NWS>
struct Halo2Cfg {
  static constexpr int A_LINES = RH * WP;
  static constexpr int A_BOX_BYTES = A_LINES * 128;
  static constexpr int A_PLANE = (A_BOX_BYTES + 1023) / 1024 * 1024;
  static constexpr int A_STAGE = 2 * A_PLANE;
  static constexpr int AS = NAS;
  static constexpr int W_BYTES = BN * 128;
  static constexpr int W_STAGE = 2 * W_BYTES;
  static constexpr int WS = NWS;
  static constexpr int W_OFF = AS * A_STAGE;
  static constexpr int SMEM_BYTES = AS * A_STAGE + WS * W_STAGE + 1024 + 256;
  static constexpr int TILE_COLS = 2 * BN;
  static constexpr int ACC_COLS = MT * TILE_COLS;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(RH * WP >= MT * 128 + 2 * WP + 2 + WP - 1, "halo box too small");
  static_assert(WP <= 256 && RH <= 256, "TMA box dimensions are limited to 256");
  static_assert(SMEM_BYTES <= 232448 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "halo tile configuration does not fit");
  static_assert(2 * (AS + WS) * 8 + 40 <= 256, "barrier block too small");
};

template <int BN, int RH, int WP, int MT, int NAS, int NWS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_halo2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const HaloParams p) {
  using Cfg = Halo2Cfg<BN, RH, WP, MT, NAS, NWS>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::AS * Cfg::A_STAGE + Cfg::WS * Cfg::W_STAGE);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + Cfg::AS;
  uint64_t* w_full = a_empty + Cfg::AS;
  uint64_t* w_empty = w_full + Cfg::WS;
  uint64_t* tfull = w_empty + Cfg::WS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::AS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < Cfg::WS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty), w_full0 = smem_u32(w_full), w_empty0 = smem_u32(w_empty);
  const int ipi = p.tiles_per_image, cchunks = p.cchunks;     // here: ITEMS (groups of MT tiles) per image

  if (warp == 0) {
    if (elect_one_sync()) {
      uint32_t as = 0, aph = 1, ws = 0, wph = 1;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int b = item / ipi, it = item - b * ipi;
        const int p0 = it * (MT * BM_TC);
        const int r0 = (p0 + WP - 1) / WP - 2;
        for (int cc = 0; cc < cchunks; ++cc) {
          const uint32_t sa = smem0 + as * Cfg::A_STAGE, fb = a_full0 + as * 8;
          mbar_wait_u32(a_empty0 + as * 8, aph);
          if (TC_NO_TMA(p)) {
            mbar_arrive(&a_full[as]);
          } else {
            mbar_expect_tx_u32(fb, 2 * Cfg::A_BOX_BYTES);
            tma_load_4d_u32(sa, &map_a_hi, fb, cc * 64, -1, r0 - 1, b);
            tma_load_4d_u32(sa + Cfg::A_PLANE, &map_a_lo, fb, cc * 64, -1, r0 - 1, b);
          }
          if (++as == Cfg::AS) { as = 0; aph ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t sw = smem0 + Cfg::W_OFF + ws * Cfg::W_STAGE, wb = w_full0 + ws * 8;
            mbar_wait_u32(w_empty0 + ws * 8, wph);
            if (TC_NO_TMA(p)) {
              mbar_arrive(&w_full[ws]);
            } else {
              mbar_expect_tx_u32(wb, Cfg::W_STAGE);
              const int wk = (tap * cchunks + cc) * BK_TC;
              tma_load_2d_u32(sw, &map_w_hi, wb, wk, 0);
              tma_load_2d_u32(sw + Cfg::W_BYTES, &map_w_lo, wb, wk, 0);
            }
            if (++ws == Cfg::WS) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_f16(BM_TC, BN);
      constexpr uint32_t idesc_wide = umma_idesc_f16(BM_TC, 2 * BN);
      static_assert(2 * BN <= 256, "the merged [W_hi ; W_lo] MMA needs N <= 256");
      const uint32_t tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      uint32_t as = 0, aph = 0, ws = 0, wph = 0, acs = 0, acph = 1;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int it = item % ipi;
        const int p0 = it * (MT * BM_TC);
        const int r0 = (p0 + WP - 1) / WP - 2;
        const int base = p0 - r0 * WP;
        mbar_wait_u32(tempty0 + acs * 8, acph);
        tc_fence_after();
        const uint32_t acc = tmem_base + acs * Cfg::ACC_COLS;
        uint32_t accum = 0;
        for (int cc = 0; cc < cchunks; ++cc) {
          mbar_wait_u32(a_full0 + as * 8, aph);
          tc_fence_after();
          const uint32_t a0 = umma_desc_sw128_lo(smem0 + as * Cfg::A_STAGE);
          int start = base - WP - 1;
          for (int kh = 0; kh < 3; ++kh, start += WP - 3)
            for (int kw = 0; kw < 3; ++kw, ++start) {
              mbar_wait_u32(w_full0 + ws * 8, wph);
              tc_fence_after();
              const uint32_t w_hi = umma_desc_sw128_lo(smem0 + Cfg::W_OFF + ws * Cfg::W_STAGE);
              if (!TC_NO_MMA(p)) {
#pragma unroll
                for (int t = 0; t < MT; ++t) {
                  const uint32_t a_hi = a0 + (uint32_t)(start + t * BM_TC) * 8u;      // tile t = the next 128 lines of the same box
                  const uint32_t a_lo = a_hi + (Cfg::A_PLANE >> 4);
                  const uint32_t d_hi = acc + t * Cfg::TILE_COLS, d_lo = d_hi + BN;
#pragma unroll
                  for (int k = 0; k < BK_TC / 16; ++k) {
                    umma_f16_lohi(d_hi, a_hi + 2 * k, w_hi + 2 * k, idesc_wide, accum);
                    umma_f16_lohi(d_lo, a_lo + 2 * k, w_hi + 2 * k, idesc, 1);
                  }
                }
              }
              accum = 1;
              umma_commit_u32(w_empty0 + ws * 8);
              if (++ws == Cfg::WS) { ws = 0; wph ^= 1; }
            }
          umma_commit_u32(a_empty0 + as * 8);
          if (++as == Cfg::AS) { as = 0; aph ^= 1; }
        }
        umma_commit_u32(tfull0 + acs * 8);
        if (++acs == 2) { acs = 0; acph ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    constexpr int NCHUNK = MT * (BN / 32);
    uint32_t ti = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++ti) {
      const int b = item / ipi, it = item - b * ipi;
      const uint32_t acs = ti & 1;
      // raster position, validity and output offset of this thread's row in tile t of the item
      auto locate = [&](int t, bool& valid, size_t& obase) {
        const int pos = it * (MT * BM_TC) + t * BM_TC + row;
        const int py = pos / WP, px = pos - py * WP;
        valid = py >= 1 && py <= p.H && px >= 1 && px <= p.W;
        obase = (((size_t)b * p.H + (py - 1)) * p.W + (px - 1)) * p.cout;             // only used when valid
      };
      uint4 rh[4], rl[4];
      auto fetch_residual = [&](int chunk) {
        if (p.res_hi && chunk < NCHUNK && TC_EPI_IO(p)) {
          bool valid; size_t obase;
          locate(chunk / (BN / 32), valid, obase);
          if (valid) {
            const int c0 = (chunk % (BN / 32)) * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              rh[q] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + obase + c0) + q);
              rl[q] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + obase + c0) + q);
            }
          }
        }
      };
      fetch_residual(0);
      mbar_wait(&tfull[acs], (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < NCHUNK; ++chunk) {
        const int t = chunk / (BN / 32), c0 = (chunk % (BN / 32)) * 32;
        bool valid; size_t obase;
        locate(t, valid, obase);
        uint32_t v[32], vl[32];
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + acs * Cfg::ACC_COLS + t * Cfg::TILE_COLS + c0;
        tmem_ld_32x32(tacc, v);
        tmem_ld_32x32(tacc + BN, vl);
        tmem_ld_wait();
        float y[32];
        const float4* sh4 = reinterpret_cast<const float4*>(p.shift + c0);
        const float4* us4 = reinterpret_cast<const float4*>(p.unscale + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = __ldg(sh4 + q), u4 = __ldg(us4 + q);
          y[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]) + __uint_as_float(vl[q * 4 + 0]), u4.x, s4.x);
          y[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]) + __uint_as_float(vl[q * 4 + 1]), u4.y, s4.y);
          y[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]) + __uint_as_float(vl[q * 4 + 2]), u4.z, s4.z);
          y[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]) + __uint_as_float(vl[q * 4 + 3]), u4.w, s4.w);
        }
        if (valid && p.res_hi && TC_EPI_IO(p)) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t hw[4] = {rh[q].x, rh[q].y, rh[q].z, rh[q].w}, lw[4] = {rl[q].x, rl[q].y, rl[q].z, rl[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              y[q * 8 + e * 2 + 0] += f16lo_to_f(hw[e]) + f16lo_to_f(lw[e]);
              y[q * 8 + e * 2 + 1] += f16hi_to_f(hw[e]) + f16hi_to_f(lw[e]);
            }
          }
        }
        fetch_residual(chunk + 1);
        if (valid && TC_EPI_IO(p)) {
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaxf(y[i], 0.f);
          }
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            __half h0, l0, h1, l1;
            split_f16(y[2 * i], h0, l0);
            split_f16(y[2 * i + 1], h1, l1);
            ph[i] = pack_f16(h0, h1);
            pl[i] = pack_f16(l0, l1);
          }
          uint4* oh = reinterpret_cast<uint4*>(p.out_hi + obase + c0);
          uint4* ol = reinterpret_cast<uint4*>(p.out_lo + obase + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            oh[q] = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
            ol[q] = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acs]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// conv1 (7x7 / stride 2 / pad 3, Cout = 64) from a pixel-PAIR layout of the padded input -- STRAPS_TC_CONV1=s2d | s2d2.
// WRITTEN AFTER THE ROUND'S GPU BUDGET WAS SPENT: compiled, its index arithmetic replayed against conv2d on the CPU
// (tools/conv1_s2d_emulation.py); NOT yet run on hardware, and the default path never reaches it.  It applies to conv1 what
// conv_halo_kernel proved on hardware for the 3x3 layers (row-shifted SWIZZLE_128B descriptors over one TMA box).
//
// conv_tc_kernel reads conv1's A operand as 21 K-blocks of 128 x 64 elements per output row (672 KB of A + 336 KB of W per tile in
// 84 TMA operations) although an output row only depends on 7 input rows (7 x 262 pixels x 24 channels x 2 planes = 176 KB), and
// conv1 is paced by exactly that TMA stream (516 us with the MMAs off, 307 us with the TMA traffic off).  Here the
// padded input is stored as pixel PAIRS:  xs[b][ph][q][0..23] = padded pixel 2q, [24..47] = padded pixel 2q+1, one 128-byte line
// per pair (elements 48..63 unused; ph = h + 3, padded pixel = w + 3).  Output pixel ow and filter tap pair kw' = kw / 2 read pair
// ow + kw', so for one filter row kh the whole A operand of an output row is ONE box of 131 lines {pairs 0..130 of input row
// 2 oh + kh}, and tap pair kw' is the 128-line window that starts at line kw' (descriptor start address + kw' * 128 B).
// K inside a filter row is ordered (kw, c) = kw' * 48 + (kw & 1) * 24 + c -- exactly the order of the existing conv1 weight rows
// (C1_KROW = 192 per filter row, kw = 7 zero), so the packed weights are shared with conv_tc_kernel.  K step j = 0..10 of a filter
// row (16 elements; step 11 is all padding) reads A at window j / 3, byte offset (j % 3) * 32, and W at chunk j / 4, byte offset
// (j % 4) * 32.  W_hi and W_lo of conv1 are adjacent in HBM, so one 2-D box {64, 128} of the [2 x 64][1344] view brings the
// [W_hi ; W_lo] chunk of the wide MMA in one operation.
// Per output row: 7 x (2 x 16.4 KB A + 48 KB W) = 565 KB in 35 TMA operations (MT = 2: two output rows share the weights, 400 KB
// per row) against 1008 KB in 84.  Two rings: NA slots of one input row (hi + lo plane), NW slots of one weight chunk.
// ---------------------------------------------------------------------------------------------------------
constexpr int XS_H = 262, XS_PAIRS = 132;          // padded rows / pixel pairs per padded row of the pair layout
constexpr int S2D_LINES = 131;                     // pairs 0..130 serve output pixels 0..127 with tap pairs 0..3

struct S2dParams {
  int n_items;            // groups of MT output rows
  int n_rows;             // batch * 128 output rows
  uint32_t a_bytes;       // bytes ONE A box delivers (131 lines x bytes per pair line in HBM)
  const float* shift;
  const float* unscale;
  float* out;             // NHWC fp32 [B,128,128,64]
  __half* pool_hi;        // POOL variant: 3x3 / stride 2 max-pooled output, split planes [B,64,64,64] (the stem tensor is never written)
  __half* pool_lo;
  int relu;
  int debug;              // STRAPS_TC_DEBUG bit mask (experiment builds only)
};

template <int MT>
struct S2dCfg {
  static constexpr int A_PLANE = 17 * 1024;                         // 131 lines x 128 B = 16,768 B; planes start on swizzle atoms
  static constexpr int A_SLOT = 2 * A_PLANE;                        // hi + lo plane of one input row
  static constexpr int NA = (MT == 1) ? 3 : 4;
  static constexpr int W_SLOT = 128 * 128;                          // [W_hi (64 rows) ; W_lo (64 rows)] x 64 K elements
  static constexpr int NW = (MT == 1) ? 6 : 5;
  static constexpr int W_OFF = NA * A_SLOT;
  static constexpr int BAR_OFF = W_OFF + NW * W_SLOT;
  static constexpr int EDGE_BYTES = 2 * 4 * 2 * 32 * 4;             // POOL: [channel chunk][quadrant][array][32] floats of the warps' last pixels
  static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 256 + EDGE_BYTES;
  static constexpr int TILE_COLS = 128;                             // [acc_hi (64) | acc_lo (64)]
  static constexpr int ACC_COLS = MT * TILE_COLS;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(S2D_LINES * 128 <= A_PLANE, "A plane too small");
  static_assert(3 + 128 <= S2D_LINES, "the last tap window must lie inside the box");
  static_assert(NA % MT == 0, "the rows of one filter-row step occupy consecutive ring slots");
  static_assert(SMEM_BYTES <= 232448 && TMEM_COLS <= 512 && 2 * (NA + NW) * 8 + 40 <= 256, "conv1 pair-layout configuration does not fit");
};

// POOL (STRAPS_TC_CONV1=s2dp, MT = 2): the 3x3 / stride 2 / pad 1 max pool is fused into the epilogue and the 268 MB stem tensor is never
// written or re-read.  A work item is then one POOLED row pr of one image = conv rows 2 pr and 2 pr + 1; pooled row pr also needs conv
// row 2 pr - 1, the odd row of the item above, so every CTA walks a CONTIGUOUS range of items and carries the horizontally pooled odd
// row in registers from one item to the next (max commutes: pool(max(a, b, c)) = max(pool(a), pool(b), pool(c))).  The first item of
// a range that does not start an image is preceded by the item above it, computed only for its carry (1 extra item in ~28).
// Horizontal pooling: a thread owns one conv pixel (TMEM lane); even lanes combine lanes - 1 / + 1 with warp shuffles, and the one
// neighbour that lives in another warp (pixel 32 q - 1) travels through 256 bytes of shared memory per warp and channel chunk.
template <int MT, bool POOL = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv1_s2d_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_w, const S2dParams p) {
  using Cfg = S2dCfg<MT>;
  static_assert(!POOL || MT == 2, "the fused pool works on pairs of conv rows");
  // items of this CTA: strided over the grid, or (POOL) one contiguous range, preceded by the item above when the range starts inside an image
  int it0, it1, itstep, it_real;
  if constexpr (POOL) {
    const int per = (p.n_items + (int)gridDim.x - 1) / (int)gridDim.x;
    const int first = (int)blockIdx.x * per;
    it1 = min(first + per, p.n_items);
    it0 = first - (((first & 63) != 0 && first < it1) ? 1 : 0);
    it_real = first;
    itstep = 1;
  } else {
    it0 = (int)blockIdx.x; it1 = p.n_items; itstep = (int)gridDim.x; it_real = 0;
  }
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* a_full = bars;                          // [NA]
  uint64_t* a_empty = a_full + Cfg::NA;             // [NA]
  uint64_t* w_full = a_empty + Cfg::NA;             // [NW]
  uint64_t* w_empty = w_full + Cfg::NW;             // [NW]
  uint64_t* tfull = w_empty + Cfg::NW;              // [2]
  uint64_t* tempty = tfull + 2;                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < Cfg::NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty), w_full0 = smem_u32(w_full), w_empty0 = smem_u32(w_empty);

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= TMA producer: per filter row the input row of every tile of the item, then the 3 weight chunks =================
      uint32_t as = 0, aph = 1, ws = 0, wph = 1;
      for (int item = it0; item < it1; item += itstep) {
        const int r_first = item * MT;                    // output row index over (b, oh); 128 rows per image, MT divides 128
        const int b = r_first >> 7, oh0 = r_first & 127;
        for (int kh = 0; kh < 7; ++kh) {
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            const uint32_t sa = smem0 + as * Cfg::A_SLOT, fb = a_full0 + as * 8;
            mbar_wait_u32(a_empty0 + as * 8, aph);
            if (TC_NO_TMA(p)) {
              mbar_arrive(&a_full[as]);
            } else {
              mbar_expect_tx_u32(fb, 2 * p.a_bytes);
              tma_load_4d_u32(sa, &map_a_hi, fb, 0, 0, 2 * (oh0 + t) + kh, b);
              tma_load_4d_u32(sa + Cfg::A_PLANE, &map_a_lo, fb, 0, 0, 2 * (oh0 + t) + kh, b);
            }
            if (++as == Cfg::NA) { as = 0; aph ^= 1; }
          }
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            const uint32_t wb = w_full0 + ws * 8;
            mbar_wait_u32(w_empty0 + ws * 8, wph);
            if (TC_NO_TMA(p)) {
              mbar_arrive(&w_full[ws]);
            } else {
              mbar_expect_tx_u32(wb, Cfg::W_SLOT);
              tma_load_2d_u32(smem0 + Cfg::W_OFF + ws * Cfg::W_SLOT, &map_w, wb, kh * C1_KROW + c * 64, 0);
            }
            if (++ws == Cfg::NW) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer: 11 K steps per filter row, tap pairs = line-shifted windows of the row's box =================
      constexpr uint32_t idesc = umma_idesc_f16(BM_TC, 64);
      constexpr uint32_t idesc_wide = umma_idesc_f16(BM_TC, 128);
      const uint32_t tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      const uint32_t adesc0 = umma_desc_sw128_lo(smem0), wdesc0 = umma_desc_sw128_lo(smem0 + Cfg::W_OFF);
      uint32_t as = 0, aph = 0, ws = 0, wph = 0, acs = 0, acph = 1;
      for (int item = it0; item < it1; item += itstep) {
        mbar_wait_u32(tempty0 + acs * 8, acph);
        tc_fence_after();
        const uint32_t acc = tmem_base + acs * Cfg::ACC_COLS;
        uint32_t accum = 0;
        for (int kh = 0; kh < 7; ++kh) {
          // the MT input rows of this step sit in the consecutive slots as .. as + MT - 1 (NA is a multiple of MT: no wrap inside)
#pragma unroll
          for (int t = 0; t < MT; ++t) mbar_wait_u32(a_full0 + (as + t) * 8, aph);
          tc_fence_after();
          uint32_t wd = 0;
#pragma unroll
          for (int j = 0; j < 11; ++j) {
            if ((j & 3) == 0) {
              mbar_wait_u32(w_full0 + ws * 8, wph);
              tc_fence_after();
              wd = wdesc0 + ws * (Cfg::W_SLOT >> 4);
            }
            const uint32_t w_k = wd + (j & 3) * 2;                         // +32 bytes along K inside the weight chunk
            const uint32_t a_off = (j / 3) * 8 + (j % 3) * 2;              // window (j / 3) lines down, +32 bytes along K inside the pair
#pragma unroll
            for (int t = 0; t < MT; ++t) {
              const uint32_t a_hi = adesc0 + (as + t) * (Cfg::A_SLOT >> 4) + a_off, a_lo = a_hi + (Cfg::A_PLANE >> 4);
              const uint32_t d_hi = acc + t * Cfg::TILE_COLS, d_lo = d_hi + 64;
              if (!TC_NO_MMA(p)) {
                umma_f16_lohi(d_hi, a_hi, w_k, idesc_wide, accum);         // A_hi.[W_hi ; W_lo] -> [acc_hi | acc_lo]
                umma_f16_lohi(d_lo, a_lo, w_k, idesc, 1);                  // A_lo.W_hi
              }
            }
            accum = 1;
            if ((j & 3) == 3 || j == 10) {
              umma_commit_u32(w_empty0 + ws * 8);
              if (++ws == Cfg::NW) { ws = 0; wph ^= 1; }
            }
          }
#pragma unroll
          for (int t = 0; t < MT; ++t) umma_commit_u32(a_empty0 + (as + t) * 8);
          as += MT;
          if (as == Cfg::NA) { as = 0; aph ^= 1; }
        }
        umma_commit_u32(tfull0 + acs * 8);
        if (++acs == 2) { acs = 0; acph ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps 2..5: BatchNorm shift, ReLU, fp32 NHWC stores =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ti = 0;
    if constexpr (POOL) {
      float* edge = reinterpret_cast<float*>(smem + Cfg::BAR_OFF + 256);
      const float ninf = __int_as_float(0xff800000);
      const int pc = quad * 16 + (lane >> 1);                 // pooled column of an even lane
      float carry[2][32];                                     // even lanes: the pooled odd conv row above, per channel chunk
#pragma unroll
      for (int i = 0; i < 32; ++i) carry[0][i] = carry[1][i] = ninf;
      for (int item = it0; item < it1; ++item, ++ti) {
        const uint32_t acs = ti & 1;
        const bool top = (item & 63) == 0;                    // pooled row 0: no conv row above
        const bool store = item >= it_real && TC_EPI_IO(p);
        mbar_wait(&tfull[acs], (ti >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int c0 = c * 32;
          float y0[32], y1[32];
          {
            const float4* sh4 = reinterpret_cast<const float4*>(p.shift + c0);
            const float4* us4 = reinterpret_cast<const float4*>(p.unscale + c0);
            const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + acs * Cfg::ACC_COLS + c0;
            uint32_t v[32], vl[32];
            tmem_ld_32x32(tacc, v);
            tmem_ld_32x32(tacc + 64, vl);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 s4 = __ldg(sh4 + q), u4 = __ldg(us4 + q);
              y0[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]) + __uint_as_float(vl[q * 4 + 0]), u4.x, s4.x);
              y0[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]) + __uint_as_float(vl[q * 4 + 1]), u4.y, s4.y);
              y0[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]) + __uint_as_float(vl[q * 4 + 2]), u4.z, s4.z);
              y0[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]) + __uint_as_float(vl[q * 4 + 3]), u4.w, s4.w);
            }
            tmem_ld_32x32(tacc + Cfg::TILE_COLS, v);
            tmem_ld_32x32(tacc + Cfg::TILE_COLS + 64, vl);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 s4 = __ldg(sh4 + q), u4 = __ldg(us4 + q);
              y1[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]) + __uint_as_float(vl[q * 4 + 0]), u4.x, s4.x);
              y1[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]) + __uint_as_float(vl[q * 4 + 1]), u4.y, s4.y);
              y1[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]) + __uint_as_float(vl[q * 4 + 2]), u4.z, s4.z);
              y1[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]) + __uint_as_float(vl[q * 4 + 3]), u4.w, s4.w);
            }
          }
          if (c == 1) {                                       // every TMEM read of the item is done: the issuer may reuse the stage
            tc_fence_before();
            mbar_arrive(&tempty[acs]);
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { y0[i] = fmaxf(y0[i], 0.f); y1[i] = fmaxf(y1[i], 0.f); }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) y0[i] = fmaxf(y0[i], y1[i]);          // y0 = vertical max of the item's two conv rows
          // the warp's last pixel is the left neighbour of the next warp's first pooled column
          float* e = edge + ((c * 4 + quad) * 2) * 32;
          if (lane == 31) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              reinterpret_cast<float4*>(e)[q] = make_float4(y0[q * 4], y0[q * 4 + 1], y0[q * 4 + 2], y0[q * 4 + 3]);
              reinterpret_cast<float4*>(e + 32)[q] = make_float4(y1[q * 4], y1[q * 4 + 1], y1[q * 4 + 2], y1[q * 4 + 3]);
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps only
          const float* ep = e - 64;                           // quadrant quad - 1 (read by lane 0 of quadrants 1..3 only)
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float um = __shfl_up_sync(0xffffffffu, y0[i], 1), u1 = __shfl_up_sync(0xffffffffu, y1[i], 1);
            const float dm = __shfl_down_sync(0xffffffffu, y0[i], 1), d1 = __shfl_down_sync(0xffffffffu, y1[i], 1);
            if (lane == 0) {
              um = quad ? ep[i] : ninf;                       // pixel -1 is padding
              u1 = quad ? ep[32 + i] : ninf;
            }
            const float hm = fmaxf(fmaxf(um, y0[i]), dm), h1 = fmaxf(fmaxf(u1, y1[i]), d1);
            o[i] = top ? hm : fmaxf(hm, carry[c][i]);
            carry[c][i] = h1;
          }
          if (store && !(lane & 1)) {
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              __half h0, l0, h1, l1;
              split_f16(o[2 * i], h0, l0);
              split_f16(o[2 * i + 1], h1, l1);
              ph[i] = pack_f16(h0, h1);
              pl[i] = pack_f16(l0, l1);
            }
            const size_t ob = ((size_t)item * 64 + pc) * 64 + c0;            // item = b * 64 + pooled row
            uint4* oh = reinterpret_cast<uint4*>(p.pool_hi + ob);
            uint4* ol = reinterpret_cast<uint4*>(p.pool_lo + ob);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              oh[q] = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
              ol[q] = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
            }
          }
        }
      }
    } else
    for (int item = it0; item < it1; item += itstep, ++ti) {
      const uint32_t acs = ti & 1;
      mbar_wait(&tfull[acs], (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < MT * 2; ++chunk) {
        const int t = chunk >> 1, c0 = (chunk & 1) * 32;
        const int r = item * MT + t;
        const bool valid = r < p.n_rows;
        uint32_t v[32], vl[32];
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + acs * Cfg::ACC_COLS + t * Cfg::TILE_COLS + c0;
        tmem_ld_32x32(tacc, v);
        tmem_ld_32x32(tacc + 64, vl);
        tmem_ld_wait();
        float y[32];
        const float4* sh4 = reinterpret_cast<const float4*>(p.shift + c0);
        const float4* us4 = reinterpret_cast<const float4*>(p.unscale + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = __ldg(sh4 + q), u4 = __ldg(us4 + q);
          y[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]) + __uint_as_float(vl[q * 4 + 0]), u4.x, s4.x);
          y[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]) + __uint_as_float(vl[q * 4 + 1]), u4.y, s4.y);
          y[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]) + __uint_as_float(vl[q * 4 + 2]), u4.z, s4.z);
          y[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]) + __uint_as_float(vl[q * 4 + 3]), u4.w, s4.w);
        }
        if (valid && TC_EPI_IO(p)) {
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaxf(y[i], 0.f);
          }
          float4* o = reinterpret_cast<float4*>(p.out + ((size_t)r * BM_TC + row) * 64 + c0);
#pragma unroll
          for (int q = 0; q < 8; ++q) o[q] = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acs]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// x NCHW fp32 [B,C,256,256] -> pixel-pair planes xs[B][262][132][PITCH] (hi, lo) for conv1_s2d_kernel.  One CTA per (b, h) image row:
// the C channel rows are read with coalesced 1 KB requests, split to fp16 hi / lo and assembled in shared memory as the complete
// padded row (pairs 0..131: the 3-pixel left border, the interior, the right border and the unused tail of every line are written
// as zeros, so only the 6 border ROWS rely on the one-time memset), which then leaves as contiguous 16-byte stores.
// PITCH = 64 elements (128-byte lines, the TMA box is {64, 131}) or 48 (dense pairs, box {48, 131}: TMA then fills 96 of the 128 bytes
// of every shared-memory line -- smaller HBM footprint, to be confirmed on hardware).
template <int PITCH>
__global__ void __launch_bounds__(256) pack_input_s2d_kernel(const float* __restrict__ x, int C, __half* __restrict__ hi,
                                                             __half* __restrict__ lo) {
  constexpr int WPP = PITCH / 2;                                   // 32-bit words per pair line
  constexpr int ROW_WORDS = XS_PAIRS * WPP;
  __shared__ __align__(16) uint32_t srow[2][ROW_WORDS];
  const int h = blockIdx.x, b = blockIdx.y, w = threadIdx.x;
  for (int i = w; i < 2 * ROW_WORDS; i += 256) (&srow[0][0])[i] = 0u;
  const float* src = x + ((size_t)b * C * IMG + h) * IMG + w;
  float v[XP_C];
#pragma unroll
  for (int c = 0; c < XP_C; ++c) v[c] = (c < C) ? __ldg(src + (size_t)c * IMG * IMG) : 0.f;
  __syncthreads();
  const int pw = w + 3;                                            // padded pixel
  const int pair = pw >> 1;
  // PITCH 64: a pair line is 32 words = every bank once, so all lanes of a store would hit the same two banks; the 16-byte units of a
  // line are therefore rotated by the pair index in shared memory (4-way conflicts, as for the dense layout) and rotated back below
  auto phys = [&](int q, int word) { return (PITCH == 64) ? q * WPP + ((((word >> 2) + q) & 7) << 2) + (word & 3) : q * WPP + word; };
#pragma unroll
  for (int c = 0; c < XP_C; c += 2) {
    __half h0, l0, h1, l1;
    split_f16(v[c], h0, l0);
    split_f16(v[c + 1], h1, l1);
    const int o = phys(pair, (pw & 1) * (XP_C / 2) + c / 2);
    srow[0][o] = pack_f16(h0, h1);
    srow[1][o] = pack_f16(l0, l1);
  }
  __syncthreads();
  const size_t base = ((size_t)b * XS_H + h + 3) * XS_PAIRS * PITCH;      // 16-byte aligned for PITCH 64 and 48
  uint4* dh = reinterpret_cast<uint4*>(hi + base);
  uint4* dl = reinterpret_cast<uint4*>(lo + base);
  const uint4* sh = reinterpret_cast<const uint4*>(srow[0]);
  const uint4* sl = reinterpret_cast<const uint4*>(srow[1]);
  for (int i = w; i < ROW_WORDS / 4; i += 256) {
    const int si = (PITCH == 64) ? (i & ~7) + (((i & 7) + (i >> 3)) & 7) : i;   // 8 units per 128-byte line: unit u of pair q sits at (u + q) & 7
    dh[i] = sh[si];
    dl[i] = sl[si];
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster compute M = 256 output pixels together.  Each CTA stages its own
// 128-pixel A tile but only HALF of the weight tile (BN/2 rows); the pair's MMA (issued by the leader) reads the other half
// from the peer's shared memory, so per SM the weight bytes -- TMA writes and MMA operand reads -- halve.  That is the
// lever profiles/r01_tc_probe.txt points at for the shared-memory-bandwidth-bound BN = 128 layers.
//   full[s]   : leader's barrier; TMA of BOTH CTAs completes on it (cta_group::2 loads may signal the peer's barrier)
//   empty[s]  : one per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tfull[a]  : one per CTA (multicast commit); tempty[a]: leader's, 256 arrivals (epilogue threads of both CTAs)
// ---------------------------------------------------------------------------------------------------------
template <int BN>
struct TcCfg2 {
  static constexpr int A_BYTES = BM_TC * 128;
  static constexpr int WH_BYTES = (BN / 2) * 128;      // this CTA's half of one weight plane
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * WH_BYTES;
  static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 5 ? 5 : (200 * 1024 / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int ACC_COLS = 2 * BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS <= 512, "BN too large for the CTA-pair kernel");
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const TcConvParams p) {
  using Cfg = TcCfg2<BN>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_pairs = (p.n_mtiles + 1) / 2;
  const int n_items = n_pairs * p.n_ntiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 256); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer (both CTAs) =================
      uint32_t it = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int mg = item / p.n_ntiles, nt = item % p.n_ntiles;
        const long long pix0 = (long long)(mg * 2 + (int)rank) * BM_TC;
        const int b0 = (int)(pix0 / p.hw_out);
        const int oh0 = (int)((pix0 % p.hw_out) / p.wout);
        for (int kb = 0; kb < p.n_kblocks; ++kb, ++it) {
          const int st = it % Cfg::STAGES;
          mbar_wait(&empty[st], ((it / Cfg::STAGES) & 1) ^ 1);
          unsigned char* sa = smem + st * Cfg::STAGE_BYTES;
          const uint32_t full_leader = mapa_u32(smem_u32(&full[st]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full[st], 2 * Cfg::STAGE_BYTES);   // bytes landing in BOTH CTAs
          int c0, c1, c2;
          if (p.conv1) {
            const int kh = kb / 3, j = kb % 3;
            c0 = kh * (XP_W * XP_C) + j * 64; c1 = 0; c2 = oh0;
          } else {
            const int tap = kb / p.cchunks, cc = kb % p.cchunks;
            const int kh = tap / p.kw_count, kw = tap % p.kw_count;
            c0 = cc * 64; c1 = kw - p.pad; c2 = oh0 * p.stride + kh - p.pad;
          }
          tma_load_4d_2cta(sa, &map_a_hi, full_leader, c0, c1, c2, b0);
          tma_load_4d_2cta(sa + Cfg::A_BYTES, &map_a_lo, full_leader, c0, c1, c2, b0);
          const int wrow = nt * BN + (int)rank * (BN / 2);
          tma_load_2d_2cta(sa + 2 * Cfg::A_BYTES, &map_w_hi, full_leader, kb * BK_TC, wrow);
          tma_load_2d_2cta(sa + 2 * Cfg::A_BYTES + Cfg::WH_BYTES, &map_w_lo, full_leader, kb * BK_TC, wrow);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ================= MMA issuer (leader CTA only) =================
      constexpr uint32_t idesc = umma_idesc_f16(2 * BM_TC, BN);
      uint32_t it = 0, ti = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters, ++ti) {
        const uint32_t as = ti & 1;
        mbar_wait(&tempty[as], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_hi = tmem_base + as * Cfg::ACC_COLS, d_lo = d_hi + BN;
        for (int kb = 0; kb < p.n_kblocks; ++kb, ++it) {
          const int st = it % Cfg::STAGES;
          mbar_wait(&full[st], (it / Cfg::STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + st * Cfg::STAGE_BYTES);
          const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + Cfg::A_BYTES);
          const uint64_t w_hi = umma_desc_sw128(sa + 2 * Cfg::A_BYTES);
          const uint64_t w_lo = umma_desc_sw128(sa + 2 * Cfg::A_BYTES + Cfg::WH_BYTES);
#pragma unroll
          for (int k = 0; k < BK_TC / 16; ++k) {
            const uint64_t ko = (uint64_t)(k * 32 >> 4);
            const uint32_t first = (kb | k) != 0;
            umma_f16_2cta(d_hi, a_hi + ko, w_hi + ko, idesc, first);
            umma_f16_2cta(d_lo, a_hi + ko, w_lo + ko, idesc, first);
            umma_f16_2cta(d_lo, a_lo + ko, w_hi + ko, idesc, 1);
          }
          umma_commit_2cta(&empty[st]);
        }
        umma_commit_2cta(&tfull[as]);
      }
    }
  } else {
    // ================= epilogue warps 2..5 (both CTAs, each its own 128 rows) =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ti = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters, ++ti) {
      const int mg = item / p.n_ntiles, nt = item % p.n_ntiles;
      const uint32_t as = ti & 1;
      mbar_wait(&tfull[as], (ti >> 1) & 1);
      tc_fence_after();
      const long long m = (long long)(mg * 2 + (int)rank) * BM_TC + row;
      const bool valid = m < p.m_total;
      const size_t obase = (size_t)m * p.cout + (size_t)nt * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32], vl[32];
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + as * Cfg::ACC_COLS + c0;
        tmem_ld_32x32(tacc, v);
        tmem_ld_32x32(tacc + BN, vl);
        tmem_ld_wait();
        float y[32];
        const float4* sh4 = reinterpret_cast<const float4*>(p.shift + nt * BN + c0);
        const float4* us4 = reinterpret_cast<const float4*>(p.unscale + nt * BN + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = __ldg(sh4 + q), u4 = __ldg(us4 + q);
          y[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]) + __uint_as_float(vl[q * 4 + 0]), u4.x, s4.x);
          y[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]) + __uint_as_float(vl[q * 4 + 1]), u4.y, s4.y);
          y[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]) + __uint_as_float(vl[q * 4 + 2]), u4.z, s4.z);
          y[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]) + __uint_as_float(vl[q * 4 + 3]), u4.w, s4.w);
        }
        if (valid) {
          if (p.res_hi) {
            const uint4* rh = reinterpret_cast<const uint4*>(p.res_hi + obase + c0);
            const uint4* rl = reinterpret_cast<const uint4*>(p.res_lo + obase + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 h = __ldg(rh + q), l = __ldg(rl + q);
              const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                y[q * 8 + e * 2 + 0] += f16lo_to_f(hw[e]) + f16lo_to_f(lw[e]);
                y[q * 8 + e * 2 + 1] += f16hi_to_f(hw[e]) + f16hi_to_f(lw[e]);
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaxf(y[i], 0.f);
          }
          if (p.out_f32) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + obase + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
          } else {
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              __half h0, l0, h1, l1;
              split_f16(y[2 * i], h0, l0);
              split_f16(y[2 * i + 1], h1, l1);
              ph[i] = pack_f16(h0, h1);
              pl[i] = pack_f16(l0, l1);
            }
            uint4* oh = reinterpret_cast<uint4*>(p.out_hi + obase + c0);
            uint4* ol = reinterpret_cast<uint4*>(p.out_lo + obase + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              oh[q] = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
              ol[q] = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));     // the leader's accumulator-free barrier
    }
  }
  tc_fence_before();
  cluster_sync_all();          // nobody leaves (or frees TMEM) while the peer may still touch its shared memory / TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair kernel, second version (STRAPS_TC_PAIR=m | m128): merged wide MMA + lean issue loops.
// WRITTEN AFTER THE ROUND'S GPU BUDGET WAS SPENT: compiled and reviewed against conv_tc2_kernel (which is verified on hardware), NOT
// yet run; the default path never reaches it.
//
// Why conv_tc2_kernel lost 3-8 %: the shared-memory port model that predicts conv_tc_kernel's K-block time to within 5 % --
// cycles = (bytes the MMAs read from shared memory + bytes TMA writes into it) / 128 B/clk -- says per K-block and SM
//   conv_tc_kernel  (64,1):  wide A_hi 16 + [W_hi;W_lo] 16, narrow A_lo 16 + W_hi 8  = 56 KB read + 48 KB written = 104 KB
//   conv_tc2_kernel BN = 64: three N = 64 MMAs, each 16 KB of A + 4 KB of half-W     = 60 KB read + 40 KB written = 100 KB
// i.e. splitting the wide MMA back into two made the pair READ more than it saved in writes.  Here the pair issues the same two
// MMAs as conv_tc_kernel, with M = 256:
//   wide   A_hi . [W_hi ; W_lo]   N = 2 BN: CTA 0 holds the W_hi rows, CTA 1 the W_lo rows        (per SM: 16 + BN/8 KB read)
//   narrow A_lo . W_hi            N = BN:   each CTA holds its half of the W_hi rows (own slot)   (per SM: 16 + BN/16 KB read)
//   BN = 64: 44 KB read + 44 KB written = 88 KB (-15 %);  BN = 128: 56 + 56 = 112 KB against 144 KB (-22 %)
// and a stage shrinks to 44 / 56 KB, so the ring holds 5 / 4 stages instead of 4 / 3.  The producer and issuer loops are the lean ones
// of conv_tc_kernel (elect.sync, incremental K-block coordinates, 32-bit descriptor words); the epilogue software-pipelines the
// residual loads like conv_tc_kernel's.  Barriers as in conv_tc2_kernel.
// ---------------------------------------------------------------------------------------------------------
template <int BN>
struct TcCfg2m {
  static constexpr int A_BYTES = BM_TC * 128;                        // one plane of this CTA's A tile
  static constexpr int WB_BYTES = BN * 128;                          // wide-B slot: W_hi rows (rank 0) or W_lo rows (rank 1)
  static constexpr int NB_BYTES = (BN / 2) * 128;                    // narrow-B slot: this CTA's half of the W_hi rows
  static constexpr int STAGE_BYTES = 2 * A_BYTES + WB_BYTES + NB_BYTES;
  static constexpr int STAGES = (BN == 64) ? 5 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int ACC_COLS = 2 * BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(SMEM_BYTES <= 232448 && TMEM_COLS <= 512 && 2 * BN <= 256, "CTA-pair configuration does not fit");
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
conv_tc2m_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                 const __grid_constant__ CUtensorMap map_wh_hi, const TcConvParams p) {
  using Cfg = TcCfg2m<BN>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]  leader's: TMA of BOTH CTAs completes on it
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]  one per CTA, released by the leader's multicast commit
  uint64_t* tfull = bars + 2 * Cfg::STAGES;    // [2]       one per CTA (multicast commit)
  uint64_t* tempty = tfull + 2;                // [2]       leader's: 256 arrivals (epilogue threads of both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int n_pairs = (p.n_mtiles + 1) / 2;
  const int n_items = n_pairs * p.n_ntiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
    tma_prefetch_desc(&map_wh_hi);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 256); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= TMA producer (both CTAs): own A tile, own wide-B rows, own half of the narrow-B rows =================
      const uint32_t full_leader0 = mapa_u32(full0, 0);      // the leader's full[0] as a shared::cluster address (+ 8 per stage)
      const CUtensorMap* map_wide = rank ? &map_w_lo : &map_w_hi;
      uint32_t st = 0, ph = 1;
      const int nkb = p.n_kblocks, cchunks = p.cchunks, kwc = p.kw_count, pad = p.pad, stride = p.stride;
      const bool conv1 = p.conv1 != 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int mg = item / p.n_ntiles, nt = item - mg * p.n_ntiles;
        const long long pix0 = (long long)(mg * 2 + rank) * BM_TC;       // a tile past the end lands out of bounds -> zeros
        const int b0 = (int)(pix0 / p.hw_out);
        const int oh0 = (int)((pix0 % p.hw_out) / p.wout);
        const int row0 = conv1 ? oh0 : oh0 * stride - pad;
        const int wrow = nt * BN, nrow = wrow + rank * (BN / 2);
        int c0 = 0, c1 = conv1 ? 0 : -pad, dh = 0, sub = 0, kwi = 0, wk = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t sa = smem0 + st * Cfg::STAGE_BYTES;
          const uint32_t fb = full_leader0 + st * 8;
          mbar_wait_u32(empty0 + st * 8, ph);
          if (TC_NO_TMA(p)) {
            if (rank == 0) mbar_arrive(&full[st]);
          } else {
            if (rank == 0) mbar_expect_tx_u32(full0 + st * 8, 2 * Cfg::STAGE_BYTES);       // bytes landing in BOTH CTAs
            tma_load_4d_2cta_u32(sa, &map_a_hi, fb, c0, c1, row0 + dh, b0);
            tma_load_4d_2cta_u32(sa + Cfg::A_BYTES, &map_a_lo, fb, c0, c1, row0 + dh, b0);
            tma_load_2d_2cta_u32(sa + 2 * Cfg::A_BYTES, map_wide, fb, wk, wrow);
            tma_load_2d_2cta_u32(sa + 2 * Cfg::A_BYTES + Cfg::WB_BYTES, &map_wh_hi, fb, wk, nrow);
          }
          wk += BK_TC;
          if (++st == Cfg::STAGES) { st = 0; ph ^= 1; }
          if (conv1) {
            c0 += BK_TC;
            if (++sub == C1_KROW / BK_TC) { sub = 0; c0 += XP_W * XP_C - C1_KROW; }
          } else {
            c0 += BK_TC;
            if (++sub == cchunks) {
              sub = 0; c0 = 0; ++c1;
              if (++kwi == kwc) { kwi = 0; c1 = -pad; ++dh; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one_sync()) {
      // ================= MMA issuer (leader CTA only): M = 256 over both CTAs' A tiles =================
      constexpr uint32_t idesc = umma_idesc_f16(2 * BM_TC, BN);
      constexpr uint32_t idesc_wide = umma_idesc_f16(2 * BM_TC, 2 * BN);
      constexpr uint32_t STAGE16 = Cfg::STAGE_BYTES >> 4, A16 = Cfg::A_BYTES >> 4, WB16 = Cfg::WB_BYTES >> 4;
      constexpr int C1_LAST = (7 * XP_C - (C1_KROW / BK_TC - 1) * BK_TC + 15) / 16;
      const uint32_t desc0 = umma_desc_sw128_lo(smem0);
      const uint32_t tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      uint32_t st = 0, ph = 0, as = 0, aph = 1;
      const int nkb = p.n_kblocks;
      const bool conv1 = p.conv1 != 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        mbar_wait_u32(tempty0 + as * 8, aph);
        tc_fence_after();
        const uint32_t d_hi = tmem_base + as * Cfg::ACC_COLS, d_lo = d_hi + BN;
        int sub = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_u32(full0 + st * 8, ph);
          tc_fence_after();
          const uint32_t a_hi = desc0 + st * STAGE16, a_lo = a_hi + A16;
          const uint32_t wb = a_hi + 2 * A16, nb = wb + WB16;
          int ksteps = BK_TC / 16;
          if (conv1) { if (++sub == C1_KROW / BK_TC) { sub = 0; ksteps = C1_LAST; } }
          if (TC_NO_MMA(p)) ksteps = 0;
#pragma unroll
          for (int k = 0; k < BK_TC / 16; ++k) {
            if (k < ksteps) {
              const uint32_t ko = k * 2;
              umma_f16_2cta_lohi(d_hi, a_hi + ko, wb + ko, idesc_wide, (kb | k) != 0);     // -> [acc_hi | acc_lo] in both CTAs
              umma_f16_2cta_lohi(d_lo, a_lo + ko, nb + ko, idesc, 1);
            }
          }
          umma_commit_2cta_u32(empty0 + st * 8);      // frees the stage in both CTAs
          if (++st == Cfg::STAGES) { st = 0; ph ^= 1; }
        }
        umma_commit_2cta_u32(tfull0 + as * 8);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps 2..5 (both CTAs, each its own 128 rows) =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    constexpr int NCHUNK = BN / 32;
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(tempty), 0);
    uint32_t ti = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters, ++ti) {
      const int mg = item / p.n_ntiles, nt = item % p.n_ntiles;
      const uint32_t as = ti & 1;
      const long long m = (long long)(mg * 2 + rank) * BM_TC + row;
      const bool valid = m < p.m_total;
      const size_t obase = (size_t)m * p.cout + (size_t)nt * BN;
      uint4 rh[4], rl[4];
      auto fetch_residual = [&](int chunk) {
        if (p.res_hi && chunk < NCHUNK && valid && TC_EPI_IO(p)) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            rh[q] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + obase + chunk * 32) + q);
            rl[q] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + obase + chunk * 32) + q);
          }
        }
      };
      fetch_residual(0);
      mbar_wait(&tfull[as], (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < NCHUNK; ++chunk) {
        const int c0 = chunk * 32;
        uint32_t v[32], vl[32];
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + as * Cfg::ACC_COLS + c0;
        tmem_ld_32x32(tacc, v);
        tmem_ld_32x32(tacc + BN, vl);
        tmem_ld_wait();
        float y[32];
        const float4* sh4 = reinterpret_cast<const float4*>(p.shift + nt * BN + c0);
        const float4* us4 = reinterpret_cast<const float4*>(p.unscale + nt * BN + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = __ldg(sh4 + q), u4 = __ldg(us4 + q);
          y[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]) + __uint_as_float(vl[q * 4 + 0]), u4.x, s4.x);
          y[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]) + __uint_as_float(vl[q * 4 + 1]), u4.y, s4.y);
          y[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]) + __uint_as_float(vl[q * 4 + 2]), u4.z, s4.z);
          y[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]) + __uint_as_float(vl[q * 4 + 3]), u4.w, s4.w);
        }
        if (valid && p.res_hi && TC_EPI_IO(p)) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t hw[4] = {rh[q].x, rh[q].y, rh[q].z, rh[q].w}, lw[4] = {rl[q].x, rl[q].y, rl[q].z, rl[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              y[q * 8 + e * 2 + 0] += f16lo_to_f(hw[e]) + f16lo_to_f(lw[e]);
              y[q * 8 + e * 2 + 1] += f16hi_to_f(hw[e]) + f16hi_to_f(lw[e]);
            }
          }
        }
        fetch_residual(chunk + 1);
        if (valid && TC_EPI_IO(p)) {
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaxf(y[i], 0.f);
          }
          if (p.out_f32) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + obase + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
          } else {
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              __half h0, l0, h1, l1;
              split_f16(y[2 * i], h0, l0);
              split_f16(y[2 * i + 1], h1, l1);
              ph[i] = pack_f16(h0, h1);
              pl[i] = pack_f16(l0, l1);
            }
            uint4* oh = reinterpret_cast<uint4*>(p.out_hi + obase + c0);
            uint4* ol = reinterpret_cast<uint4*>(p.out_lo + obase + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              oh[q] = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
              ol[q] = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(tempty_leader0 + as * 8);     // the leader's accumulator-free barrier
    }
  }
  tc_fence_before();
  cluster_sync_all();          // nobody leaves (or frees TMEM) while the peer may still touch its shared memory / TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
  }
}

// x NCHW fp32 [B,C,256,256] -> padded NHWC split planes [B,262,264,24] (interior only; the halo stays zero).
// One CTA per (b, h) image row, one thread per pixel: the C channel rows are read with fully coalesced 1 KB
// requests (all loads in flight before the first use), split to fp16 hi/lo and staged through shared memory so
// that the two 12 KB output rows leave as contiguous 16-byte stores.  HBM-bound: 4*C + 96 bytes per pixel.
__global__ void __launch_bounds__(256) pack_input_tc_kernel(const float* __restrict__ x, int C, __half* __restrict__ hi,
                                                            __half* __restrict__ lo) {
  __shared__ __align__(16) uint32_t srow[2][IMG * XP_C / 2];   // [plane][256 px * 12 half2]
  const int h = blockIdx.x, b = blockIdx.y, w = threadIdx.x;
  const float* src = x + ((size_t)b * C * IMG + h) * IMG + w;
  float v[XP_C];
#pragma unroll
  for (int c = 0; c < XP_C; ++c) v[c] = (c < C) ? __ldg(src + (size_t)c * IMG * IMG) : 0.f;
#pragma unroll
  for (int c = 0; c < XP_C; c += 2) {
    __half h0, l0, h1, l1;
    split_f16(v[c], h0, l0);
    split_f16(v[c + 1], h1, l1);
    srow[0][w * (XP_C / 2) + c / 2] = pack_f16(h0, h1);
    srow[1][w * (XP_C / 2) + c / 2] = pack_f16(l0, l1);
  }
  __syncthreads();
  const size_t base = (((size_t)b * XP_H + h + 3) * XP_W + 3) * XP_C;   // 16-byte aligned (48 B per pixel)
  uint4* dh = reinterpret_cast<uint4*>(hi + base);
  uint4* dl = reinterpret_cast<uint4*>(lo + base);
  const uint4* sh = reinterpret_cast<const uint4*>(srow[0]);
  const uint4* sl = reinterpret_cast<const uint4*>(srow[1]);
  for (int i = w; i < IMG * XP_C * 2 / 16; i += 256) {
    dh[i] = sh[i];
    dl[i] = sl[i];
  }
}

// per output channel: power-of-two scale that puts max|w * bn_scale| into [2^13, 2^14), and its exact inverse
__global__ void w_rowscale_kernel(const float* __restrict__ w, const float* __restrict__ scale, int row_len,
                                  float* __restrict__ pscale, float* __restrict__ unscale) {
  __shared__ float red[256];
  const int n = blockIdx.x;
  float m = 0.f;
  for (int i = threadIdx.x; i < row_len; i += blockDim.x) m = fmaxf(m, fabsf(w[(size_t)n * row_len + i]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mx = red[0] * fabsf(scale[n]);
    int e = 0;
    if (mx > 0.f && isfinite(mx)) {
      frexpf(mx, &e);          // mx = f * 2^e, f in [0.5, 1)
      e = 14 - e;              // mx * 2^e in [2^13, 2^14)
      e = max(-60, min(60, e));
    }
    pscale[n] = ldexpf(1.f, e);
    unscale[n] = ldexpf(1.f, -e);
  }
}

// OIHW fp32 -> [Cout][K_eff] fp16 hi/lo with the BN scale and the power-of-two row scale folded in.
__global__ void pack_w_tc_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                 const float* __restrict__ pscale, int cout, int cin,
                                 int ks, int conv1, int k_eff, __half* __restrict__ hi,
                                 __half* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)cout * k_eff) return;
  const int n = (int)(i / k_eff), k = (int)(i % k_eff);
  float v = 0.f;
  if (conv1) {
    const int kh = k / C1_KROW, r = k % C1_KROW, kw = r / XP_C, c = r % XP_C;
    if (kw < ks && c < cin) v = w[(((size_t)n * cin + c) * ks + kh) * ks + kw];
  } else {
    const int tap = k / cin, c = k % cin, kh = tap / ks, kw = tap % ks;
    v = w[(((size_t)n * cin + c) * ks + kh) * ks + kw];
  }
  v = (v * scale[n]) * pscale[n];
  __half h, l;
  split_f16(v, h, l);
  hi[i] = h;
  lo[i] = l;
}

// 3x3/2 max pool: fp32 NHWC [B,128,128,64] -> split planes [B,64,64,64]
__global__ void maxpool_split_kernel(const float* __restrict__ in, int B, int H, int W, int C,
                                     __half* __restrict__ hi, __half* __restrict__ lo) {
  const int HO = H / 2, WO = W / 2, C4 = C / 4;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * HO * WO * C4) return;
  const int c4 = i % C4;
  size_t t = i / C4;
  const int ow = t % WO; t /= WO;
  const int oh = t % HO;
  const size_t b = t / HO;
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int dh = 0; dh < 3; ++dh) {
    const int ih = oh * 2 - 1 + dh;
    if (ih < 0 || ih >= H) continue;
#pragma unroll
    for (int dw = 0; dw < 3; ++dw) {
      const int iw = ow * 2 - 1 + dw;
      if (iw < 0 || iw >= W) continue;
      const float4 v = *reinterpret_cast<const float4*>(in + ((b * H + ih) * W + iw) * C + c4 * 4);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  __half h[4], l[4];
  split_f16(m.x, h[0], l[0]); split_f16(m.y, h[1], l[1]); split_f16(m.z, h[2], l[2]); split_f16(m.w, h[3], l[3]);
  const size_t o = ((b * HO + oh) * WO + ow) * C + c4 * 4;
  *reinterpret_cast<uint2*>(hi + o) = make_uint2(pack_f16(h[0], h[1]), pack_f16(h[2], h[3]));
  *reinterpret_cast<uint2*>(lo + o) = make_uint2(pack_f16(l[0], l[1]), pack_f16(l[2], l[3]));
}

// split NHWC planes -> fp32 NCHW (parity hook)
__global__ void split_to_nchw_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int C,
                                     int H, int W, float* __restrict__ x, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = i % W;
  size_t t = i / W;
  const int h = t % H; t /= H;
  const int c = t % C;
  const size_t b = t / C;
  const size_t s = ((b * H + h) * W + w) * C + c;
  x[i] = __half2float(hi[s]) + __half2float(lo[s]);
}
__global__ void f32_nhwc_to_nchw_kernel(const float* __restrict__ y, int C, int H, int W, float* __restrict__ x, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = i % W;
  size_t t = i / W;
  const int h = t % H; t /= H;
  const int c = t % C;
  const size_t b = t / C;
  x[i] = y[((b * H + h) * W + w) * C + c];
}
__global__ void avgpool_f32_kernel(const float* __restrict__ in, int B, int HW, int C, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C, b = i / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += in[((size_t)b * HW + p) * C + c];
  out[i] = s / (float)HW;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcLayerMaps {
  CUtensorMap a_hi, a_lo, w_hi, w_lo;
  CUtensorMap w2_hi, w2_lo;     // half-tile (BN/2 rows) boxes for the CTA-pair kernel and the 2-CTA multicast clusters
  CUtensorMap w4_hi, w4_lo;     // quarter-tile boxes for the 4-CTA multicast clusters
  // BK = 32 stages: boxes of 32 K elements, SWIZZLE_64B (w32h = half-tile rows for the 2-CTA multicast clusters)
  CUtensorMap a32_hi, a32_lo, w32_hi, w32_lo, w32h_hi, w32h_lo;
  // halo boxes {64 channels, W + 2 pixels, RH rows} of the unpadded planes for conv_halo_kernel (halo = 1 when the layer qualifies)
  CUtensorMap h_hi, h_lo;
  int halo;
  CUtensorMap h2_hi, h2_lo;     // 7-row boxes for conv_halo2_kernel (layer1: two tiles per item); halo2 = 1 when present
  int halo2;
  // merged-plane maps (STRAPS_TC_TMA2): A over {.., 2 planes} (5-D), W over {K, Cout, 2 planes} (3-D)
  CUtensorMap a5, w3;
  int has_merged;
};

struct TcState {
  EncodeTiledFn encode;
  int num_sms;
  __half* xp;            // padded conv1 input planes (hi then lo), for max_batch
  size_t xp_plane;              // elements per plane
  __half* wpool;         // packed weights
  float* rowscale;              // [sum cout] power-of-two weight row scales
  float* unscale;               // [sum cout] their inverses
  size_t ch_off[NCONV];         // offset of each conv in the two arrays above
  std::map<int, std::vector<TcLayerMaps>> maps;   // per batch size
  struct TcTrain* train;        // training-path state, allocated on first use
  // fork/join for the 1x1 downsample convolutions (independent of the block's first 3x3 convolution): launched on `side`,
  // their CTAs fill the SMs the 3x3 kernel's last, partial wave leaves idle
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  int ds_overlap;
  // pixel-pair input planes of conv1_s2d_kernel (STRAPS_TC_CONV1=s2d | s2d2): allocated on first use, for max_batch
  __half* xs;
  size_t xs_plane;
  int xs_pitch;                 // elements per pair line in HBM: 64 (default) or 48 (STRAPS_TC_S2D_PITCH=48)
  struct S2dMaps { CUtensorMap a_hi, a_lo, w; };
  std::map<int, S2dMaps> s2d_maps;
};

static inline __half* plane_hi(const straps_regressor* r, int buf) {
  return reinterpret_cast<__half*>(r->ws + r->bufs[buf].offset);
}
static inline __half* plane_lo(const straps_regressor* r, int buf) {
  const ActBuf& b = r->bufs[buf];
  return plane_hi(r, buf) + (size_t)r->max_batch * b.h * b.w * b.c;
}
static inline bool out_is_f32(const straps_regressor* r, int ci) {
  return ci == 0 || r->conv[ci].out_buf == r->buf_final;
}

int tc_create(straps_regressor* r) {
  TcState* t = new TcState();
  r->tc = t;
  t->xp = nullptr; t->wpool = nullptr; t->encode = nullptr; t->num_sms = 148; t->train = nullptr;
  t->side = nullptr; t->ev_fork = t->ev_join = nullptr;
  t->xs = nullptr; t->xs_plane = 0; t->xs_pitch = 64;
  { const char* e = getenv("STRAPS_TC_DS_OVERLAP"); t->ds_overlap = e ? atoi(e) : 1; }
  // The driver entry point is resolved at first use (no GPU / driver in the build container).
  t->xp_plane = (size_t)r->max_batch * XP_H * XP_W * XP_C;
  size_t welts = 0, nch = 0;
  for (int i = 0; i < NCONV; ++i) {
    ConvSpec& c = r->conv[i];
    c.k_eff = (i == 0) ? 7 * C1_KROW : c.ksize * c.ksize * c.cin;
    welts += 2 * (size_t)c.cout * c.k_eff;
    t->ch_off[i] = nch;
    nch += c.cout;
  }
  t->rowscale = t->unscale = nullptr;
  if (cudaMalloc(&t->xp, 2 * t->xp_plane * sizeof(__half)) != cudaSuccess ||
      cudaMalloc(&t->wpool, welts * sizeof(__half)) != cudaSuccess ||
      cudaMalloc(&t->rowscale, 2 * nch * sizeof(float)) != cudaSuccess) {
    set_error("tc_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  t->unscale = t->rowscale + nch;
  STRAPS_CUDA(cudaMemset(t->xp, 0, 2 * t->xp_plane * sizeof(__half)));
  __half* p = t->wpool;
  for (int i = 0; i < NCONV; ++i) {
    ConvSpec& c = r->conv[i];
    c.w_hi = p; p += (size_t)c.cout * c.k_eff;
    c.w_lo = p; p += (size_t)c.cout * c.k_eff;
  }
  int dev = 0;
  STRAPS_CUDA(cudaGetDevice(&dev));
  STRAPS_CUDA(cudaDeviceGetAttribute(&t->num_sms, cudaDevAttrMultiProcessorCount, dev));
  STRAPS_CUDA(cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking));
  STRAPS_CUDA(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
  STRAPS_CUDA(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
  return 0;
}

static void tc_train_free(TcState* t);

void tc_destroy(straps_regressor* r) {
  TcState* t = static_cast<TcState*>(r->tc);
  if (!t) return;
  tc_train_free(t);
  if (t->xp) cudaFree(t->xp);
  if (t->xs) cudaFree(t->xs);
  if (t->wpool) cudaFree(t->wpool);
  if (t->rowscale) cudaFree(t->rowscale);
  if (t->ev_fork) cudaEventDestroy(t->ev_fork);
  if (t->ev_join) cudaEventDestroy(t->ev_join);
  if (t->side) cudaStreamDestroy(t->side);
  delete t;
  r->tc = nullptr;
}

int tc_pack(straps_regressor* r, const float* const* conv_w, cudaStream_t st) {
  for (int i = 0; i < NCONV; ++i) {
    ConvSpec& c = r->conv[i];
    const size_t total = (size_t)c.cout * c.k_eff;
    TcState* t = static_cast<TcState*>(r->tc);
    w_rowscale_kernel<<<c.cout, 256, 0, st>>>(conv_w[i], c.scale, c.cin * c.ksize * c.ksize, t->rowscale + t->ch_off[i],
                                             t->unscale + t->ch_off[i]);
    STRAPS_LAUNCH_CHECK();
    pack_w_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        conv_w[i], c.scale, t->rowscale + t->ch_off[i], c.cout, c.cin, c.ksize, i == 0, c.k_eff, static_cast<__half*>(c.w_hi),
        static_cast<__half*>(c.w_lo));
    STRAPS_LAUNCH_CHECK();
  }
  return 0;
}

static int encode(TcState* t, CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUresult rc = t->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu, box %u %u %u %u)", (int)rc,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return 1;
  }
  return 0;
}

// geometry of one implicit-GEMM convolution as the kernel sees it (a forward conv, or a data gradient recast as one)
struct TcGeom {
  int cin, cout, ksize, stride, pad, hin, win, hout, wout, k_eff, conv1;
};
static TcGeom geom_fwd(const ConvSpec& c, int ci) {
  TcGeom g;
  g.cin = c.cin; g.cout = c.cout; g.ksize = c.ksize; g.stride = c.stride; g.pad = c.pad;
  g.hin = c.hin; g.win = c.win; g.hout = c.hout; g.wout = c.wout; g.k_eff = c.k_eff; g.conv1 = (ci == 0);
  return g;
}
// dX = conv(dY [zero-upsampled x2 when the forward conv has stride 2], taps flipped, Cin <-> Cout), stride 1, pad ks-1-pad
static TcGeom geom_dgrad(const ConvSpec& c) {
  TcGeom g;
  g.cin = c.cout; g.cout = c.cin; g.ksize = c.ksize; g.stride = 1; g.pad = c.ksize - 1 - c.pad;
  g.hin = c.hin; g.win = c.win; g.hout = c.hin; g.wout = c.win; g.k_eff = c.ksize * c.ksize * c.cout; g.conv1 = 0;
  return g;
}

// Kernel-selection switches (README: environment switches).  They are read from the environment ONCE PER ENCODER CALL / TRAINING STEP
// (tc_refresh_switches), not per launch -- a process may still change them between calls to compare variants (tools/halo_check.py).
// One host thread per GPU drives the library, so a plain global is enough; everything defaults to the shipped path.
struct TcSwitches {
  bool wide;          // STRAPS_TC_TILES=wide
  int debug;          // STRAPS_TC_DEBUG (experiment builds)
  int halo;           // STRAPS_TC_HALO: 0 off, 1 layer1 + layer2, 64 / 128 one of them, 2 = conv_halo2_kernel on layer1
  bool halo_e8;       // "...,8": two epilogue warps per quadrant in conv_halo_kernel
  char pair[8];       // STRAPS_TC_PAIR: "", "1", "all", "m", "m128"
  int epw;            // STRAPS_TC_EPI_WARPS
  int tma2;           // STRAPS_TC_TMA2
  int pdl;            // STRAPS_TC_PDL
  int conv1;          // STRAPS_TC_CONV1: 0 shipped, 1 s2d, 2 s2d2, 3 s2dp
};
static TcSwitches g_sw = {false, 0, 0, false, {0}, 4, 0, 0, 0};

static void tc_refresh_switches() {
  const char* e;
  e = getenv("STRAPS_TC_TILES");      g_sw.wide = e && e[0] == 'w';
  e = getenv("STRAPS_TC_DEBUG");      g_sw.debug = e ? atoi(e) : 0;
  e = getenv("STRAPS_TC_HALO");       g_sw.halo = e ? atoi(e) : 0; g_sw.halo_e8 = e && strstr(e, ",8") != nullptr;
  e = getenv("STRAPS_TC_PAIR");       memset(g_sw.pair, 0, sizeof(g_sw.pair)); if (e) strncpy(g_sw.pair, e, sizeof(g_sw.pair) - 1);
  e = getenv("STRAPS_TC_EPI_WARPS");  g_sw.epw = e ? atoi(e) : 4;
  e = getenv("STRAPS_TC_TMA2");       g_sw.tma2 = e ? atoi(e) : 0;
  e = getenv("STRAPS_TC_PDL");        g_sw.pdl = e ? atoi(e) : 0;
  e = getenv("STRAPS_TC_CONV1");
  g_sw.conv1 = (e && strncmp(e, "s2d", 3) == 0) ? (e[3] == 'p' ? 3 : e[3] == '2' ? 2 : 1) : 0;
}

// tile configuration per layer (see TcCfg): BN, MT
static void tile_cfg(int cout, int* bn, int* mt) {
  const bool wide = g_sw.wide;                    // "wide" selects the experimental (64,2)/(128,2)/(256,1) shapes
  if (cout == 64) { *bn = 64; *mt = wide ? 2 : 1; }
  else if (cout == 128) { *bn = 128; *mt = wide ? 2 : 1; }
  else if (cout == 256 && wide) { *bn = 256; *mt = 1; }
  else { *bn = 128; *mt = 1; }
}
static int tile_bn(int cout) { int bn, mt; tile_cfg(cout, &bn, &mt); return bn; }

static int ensure_encode(TcState* t) {
  if (t->encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  STRAPS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  STRAPS_CHECK(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
  t->encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

// layers conv_halo_kernel is instantiated for: 3x3 / stride 1 / pad 1, Cin = Cout = the N tile, 64x64x64 (layer1) or 32x32x128 (layer2)
static bool halo_geom(const TcGeom& c, int* rh) {
  if (c.conv1 || c.ksize != 3 || c.stride != 1 || c.pad != 1 || c.cin != c.cout || c.hin != c.win || c.hout != c.hin) return false;
  if (c.cout == 64 && c.hin == 64) { *rh = 5; return true; }
  if (c.cout == 128 && c.hin == 32) { *rh = 7; return true; }
  return false;
}

// tensor maps of one convolution: A over the split activation planes, W over [Cout][K_eff]
static int build_layer_maps(TcState* t, const TcGeom& c, int B, __half* a_hi, __half* a_lo, void* w_hi, void* w_lo, TcLayerMaps& out) {
  const int bn = tile_bn(c.cout);
  out.halo = 0;
  out.halo2 = 0;
  const long long wplane = (const char*)w_lo - (const char*)w_hi;           // the lo plane follows the hi plane in the weight pool
  {
    cuuint64_t dims[2] = {(cuuint64_t)c.k_eff, (cuuint64_t)c.cout};
    cuuint64_t str[1] = {(cuuint64_t)c.k_eff * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t es[2] = {1, 1};
    if (encode(t, &out.w_hi, w_hi, 2, dims, str, box, es)) return 1;
    if (encode(t, &out.w_lo, w_lo, 2, dims, str, box, es)) return 1;
    cuuint32_t box2[2] = {64, (cuuint32_t)((c.cout == 64 ? 64 : 128) / 2)};
    if (encode(t, &out.w2_hi, w_hi, 2, dims, str, box2, es)) return 1;
    if (encode(t, &out.w2_lo, w_lo, 2, dims, str, box2, es)) return 1;
    cuuint32_t box4[2] = {64, (cuuint32_t)((c.cout == 64 ? 64 : 128) / 4)};
    if (encode(t, &out.w4_hi, w_hi, 2, dims, str, box4, es)) return 1;
    if (encode(t, &out.w4_lo, w_lo, 2, dims, str, box4, es)) return 1;
    cuuint32_t b32[2] = {32, (cuuint32_t)bn}, b32h[2] = {32, (cuuint32_t)((c.cout == 64 ? 64 : 128) / 2)};
    if (encode(t, &out.w32_hi, w_hi, 2, dims, str, b32, es, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (encode(t, &out.w32_lo, w_lo, 2, dims, str, b32, es, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (encode(t, &out.w32h_hi, w_hi, 2, dims, str, b32h, es, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (encode(t, &out.w32h_lo, w_lo, 2, dims, str, b32h, es, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  }
  // Optional maps of switched-off variants must never take the shipped path down: a failed encode only clears the variant's flag.
  const long long aplane = (const char*)a_lo - (const char*)a_hi;
  const bool a5_ok = aplane > 0 && aplane % 16 == 0;
  out.has_merged = (a5_ok && wplane > 0 && wplane % 16 == 0) ? 1 : 0;
  if (out.has_merged) {
    cuuint64_t d3[3] = {(cuuint64_t)c.k_eff, (cuuint64_t)c.cout, 2};
    cuuint64_t s3[2] = {(cuuint64_t)c.k_eff * 2, (cuuint64_t)wplane};
    cuuint32_t b3[3] = {64, (cuuint32_t)bn, 2};
    cuuint32_t e3[3] = {1, 1, 1};
    if (encode(t, &out.w3, w_hi, 3, d3, s3, b3, e3)) out.has_merged = 0;
  }
  if (c.conv1) {
    // conv1: (flattened kw,c run | ow | oh | b) over the padded input, strides bake in the stride-2 sampling
    cuuint64_t dims[4] = {(cuuint64_t)6 * XP_W * XP_C + C1_KROW, 128, 128, (cuuint64_t)B};
    cuuint64_t str[3] = {2 * XP_C * 2, (cuuint64_t)2 * XP_W * XP_C * 2, (cuuint64_t)XP_H * XP_W * XP_C * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(t, &out.a_hi, a_hi, 4, dims, str, box, es)) return 1;
    if (encode(t, &out.a_lo, a_lo, 4, dims, str, box, es)) return 1;
    if (out.has_merged) {
      cuuint64_t d5[5] = {dims[0], dims[1], dims[2], dims[3], 2};
      cuuint64_t s5[4] = {str[0], str[1], str[2], (cuuint64_t)aplane};
      cuuint32_t b5[5] = {64, 128, 1, 1, 2};
      cuuint32_t e5[5] = {1, 1, 1, 1, 1};
      if (encode(t, &out.a5, a_hi, 5, d5, s5, b5, e5)) out.has_merged = 0;
    }
    box[0] = 32;
    if (encode(t, &out.a32_hi, a_hi, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (encode(t, &out.a32_lo, a_lo, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  } else {
    const int th = (c.hout * c.wout >= BM_TC) ? BM_TC / c.wout : c.hout;
    const int nb = BM_TC / (c.wout * th);
    cuuint64_t dims[4] = {(cuuint64_t)c.cin, (cuuint64_t)c.win, (cuuint64_t)c.hin, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)c.cin * 2, (cuuint64_t)c.win * c.cin * 2, (cuuint64_t)c.hin * c.win * c.cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(c.wout * c.stride), (cuuint32_t)(th * c.stride), (cuuint32_t)nb};
    cuuint32_t es[4] = {1, (cuuint32_t)c.stride, (cuuint32_t)c.stride, 1};
    if (encode(t, &out.a_hi, a_hi, 4, dims, str, box, es)) return 1;
    if (encode(t, &out.a_lo, a_lo, 4, dims, str, box, es)) return 1;
    if (out.has_merged) {
      cuuint64_t d5[5] = {dims[0], dims[1], dims[2], dims[3], 2};
      cuuint64_t s5[4] = {str[0], str[1], str[2], (cuuint64_t)aplane};
      cuuint32_t b5[5] = {box[0], box[1], box[2], box[3], 2};
      cuuint32_t e5[5] = {es[0], es[1], es[2], es[3], 1};
      if (encode(t, &out.a5, a_hi, 5, d5, s5, b5, e5)) out.has_merged = 0;
    }
    box[0] = 32;
    if (encode(t, &out.a32_hi, a_hi, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    if (encode(t, &out.a32_lo, a_lo, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    int rh = 0;
    if (halo_geom(c, &rh)) {
      cuuint32_t hbox[4] = {64, (cuuint32_t)(c.win + 2), (cuuint32_t)rh, 1};
      cuuint32_t hes[4] = {1, 1, 1, 1};
      out.halo = (encode(t, &out.h_hi, a_hi, 4, dims, str, hbox, hes) || encode(t, &out.h_lo, a_lo, 4, dims, str, hbox, hes)) ? 0 : 1;
      if (c.cout == 64) {
        hbox[2] = 7;
        out.halo2 = (encode(t, &out.h2_hi, a_hi, 4, dims, str, hbox, hes) || encode(t, &out.h2_lo, a_lo, 4, dims, str, hbox, hes)) ? 0 : 1;
      }
    }
  }
  return 0;
}

static int build_maps(straps_regressor* r, int B, std::vector<TcLayerMaps>& out) {
  TcState* t = static_cast<TcState*>(r->tc);
  if (ensure_encode(t)) return 1;
  out.resize(NCONV);
  for (int i = 0; i < NCONV; ++i) {
    const ConvSpec& c = r->conv[i];
    __half* a_hi = (i == 0) ? t->xp : plane_hi(r, c.in_buf);
    __half* a_lo = (i == 0) ? t->xp + t->xp_plane : plane_lo(r, c.in_buf);
    if (build_layer_maps(t, geom_fwd(c, i), B, a_hi, a_lo, c.w_hi, c.w_lo, out[i])) return 1;
  }
  return 0;
}

// the shipped tile shapes with two TMA operations per stage (see conv_tc_kernel: M2)
template <int BN>
static int launch_conv_tc_m2(const TcLayerMaps& m, const TcConvParams& p, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<BN, 1, 64>;
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, 1, 1, 64, 4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const int items = p.n_mtiles * p.n_ntiles;
  const int grid = items < num_sms ? items : num_sms;
  conv_tc_kernel<BN, 1, 1, 64, 4, false, true><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(m.a5, m.a5, m.w3, m.w3, p);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// the shipped tile shapes launched with programmatic stream serialization (see conv_tc_kernel: PDL)
template <int BN>
static int launch_conv_tc_pdl(const TcLayerMaps& m, const TcConvParams& p, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<BN, 1, 64>;
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, 1, 1, 64, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const int items = p.n_mtiles * p.n_ntiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(items < num_sms ? items : num_sms, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  STRAPS_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, 1, 1, 64, 4, true>, m.a_hi, m.a_lo, m.w_hi, m.w_lo, p));
  straps::count_launch();
  return 0;
}

template <int BN, int MT, int BK = 64, int EPW = 4>
static int launch_conv_tc(const TcLayerMaps& m, const TcConvParams& p, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<BN, MT, BK>;
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, MT, 1, BK, EPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const int items = ((p.n_mtiles + MT - 1) / MT) * p.n_ntiles;
  const int grid = items < num_sms ? items : num_sms;
  constexpr int threads = 64 + 32 * EPW;
  if (BK == 64) conv_tc_kernel<BN, MT, 1, BK, EPW><<<grid, threads, Cfg::SMEM_BYTES, st>>>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, p);
  else conv_tc_kernel<BN, MT, 1, BK, EPW><<<grid, threads, Cfg::SMEM_BYTES, st>>>(m.a32_hi, m.a32_lo, m.w32_hi, m.w32_lo, p);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// cluster variant: CL CTAs along M share the weight tile through TMA multicast
template <int BN, int CL, int BK = 64>
static int launch_conv_tc_cl(const TcLayerMaps& m, const TcConvParams& p, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<BN, 1, BK>;
  static_assert(BK == 64 || CL == 2, "BK = 32 has half-tile weight maps only");
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, 1, CL, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const int items = ((p.n_mtiles + CL - 1) / CL) * p.n_ntiles;
  const int clusters = items < num_sms / CL ? items : num_sms / CL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CL * clusters, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const CUtensorMap& ah = (BK == 64) ? m.a_hi : m.a32_hi;
  const CUtensorMap& al = (BK == 64) ? m.a_lo : m.a32_lo;
  const CUtensorMap& wh = (BK == 32) ? m.w32h_hi : (CL == 2) ? m.w2_hi : m.w4_hi;
  const CUtensorMap& wl = (BK == 32) ? m.w32h_lo : (CL == 2) ? m.w2_lo : m.w4_lo;
  STRAPS_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, 1, CL, BK>, ah, al, wh, wl, p));
  straps::count_launch();
  return 0;
}

template <int BN>
static int launch_conv_tc2(const TcLayerMaps& m, const TcConvParams& p, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg2<BN>;
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const int items = ((p.n_mtiles + 1) / 2) * p.n_ntiles;
  const int clusters = items < num_sms / 2 ? items : num_sms / 2;
  conv_tc2_kernel<BN><<<2 * clusters, TC_THREADS, Cfg::SMEM_BYTES, st>>>(m.a_hi, m.a_lo, m.w2_hi, m.w2_lo, p);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

template <int BN>
static int launch_conv_tc2m(const TcLayerMaps& m, const TcConvParams& p, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg2m<BN>;
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv_tc2m_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const int items = ((p.n_mtiles + 1) / 2) * p.n_ntiles;
  const int clusters = items < num_sms / 2 ? items : num_sms / 2;
  // full-tile weight boxes {64, BN} for the wide operand (W_hi for rank 0, W_lo for rank 1), half-tile boxes of W_hi for the narrow one
  conv_tc2m_kernel<BN><<<2 * clusters, TC_THREADS, Cfg::SMEM_BYTES, st>>>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, m.w2_hi, p);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// halo variant (conv_halo_kernel): one work item = one 128-position tile of one image's padded raster
template <int BN, int RH, int WP, int EPW = 4>
static int launch_conv_halo(const TcLayerMaps& m, const TcConvParams& p, const TcGeom& c, int B, int num_sms, cudaStream_t st) {
  using Cfg = HaloCfg<BN, RH, WP>;
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN, RH, WP, EPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  HaloParams h;
  memset(&h, 0, sizeof(h));
  h.H = c.hin; h.W = c.win;
  h.tiles_per_image = (c.hin * WP + c.win) / BM_TC + 1;      // the last tile that still holds interior position (H, W)
  h.n_items = B * h.tiles_per_image;
  h.cchunks = c.cin / 64;
  h.cout = c.cout;
  h.shift = p.shift; h.unscale = p.unscale;
  h.out_hi = p.out_hi; h.out_lo = p.out_lo;
  h.res_hi = p.res_hi; h.res_lo = p.res_lo;
  h.relu = p.relu;
  h.debug = p.debug;
  const int grid = h.n_items < num_sms ? h.n_items : num_sms;
  conv_halo_kernel<BN, RH, WP, EPW><<<grid, 64 + 32 * EPW, Cfg::SMEM_BYTES, st>>>(m.h_hi, m.h_lo, m.w_hi, m.w_lo, h);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// conv_halo2_kernel, layer1 configuration: two tiles per item, one halo stage of 7 raster rows, five weight stages
static int launch_conv_halo2(const TcLayerMaps& m, const TcConvParams& p, const TcGeom& c, int B, int num_sms, cudaStream_t st) {
  constexpr int BN = 64, RH = 7, WP = 66, MT = 2;
  using Cfg = Halo2Cfg<BN, RH, WP, MT, 1, 5>;
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv_halo2_kernel<BN, RH, WP, MT, 1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  HaloParams h;
  memset(&h, 0, sizeof(h));
  h.H = c.hin; h.W = c.win;
  const int tiles = (c.hin * WP + c.win) / BM_TC + 1;          // tiles that hold an interior position (34 on layer1)
  h.tiles_per_image = (tiles + MT - 1) / MT;                   // this kernel counts ITEMS of MT tiles
  h.n_items = B * h.tiles_per_image;
  h.cchunks = c.cin / 64;
  h.cout = c.cout;
  h.shift = p.shift; h.unscale = p.unscale;
  h.out_hi = p.out_hi; h.out_lo = p.out_lo;
  h.res_hi = p.res_hi; h.res_lo = p.res_lo;
  h.relu = p.relu;
  h.debug = p.debug;
  const int grid = h.n_items < num_sms ? h.n_items : num_sms;
  conv_halo2_kernel<BN, RH, WP, MT, 1, 5><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(m.h2_hi, m.h2_lo, m.w_hi, m.w_lo, h);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// p carries the pointers (shift, unscale, outputs, residuals, relu); the geometry fields are filled here
static int run_tc(TcState* t, const TcGeom& c, const TcLayerMaps& m, TcConvParams p, int B, cudaStream_t st) {
  const int bn = tile_bn(c.cout);
  p.m_total = (long long)B * c.hout * c.wout;
  p.n_mtiles = (int)((p.m_total + BM_TC - 1) / BM_TC);
  p.n_ntiles = c.cout / bn;
  p.conv1 = c.conv1;
  // STRAPS_TC_BK = 32 | 64: K elements per pipeline stage (see TcCfg); STRAPS_TC_MCAST = 2 | 4: clusters of that many CTAs share
  // each weight tile through TMA multicast
  static const int bk = [] { const char* e = getenv("STRAPS_TC_BK"); const int v = e ? atoi(e) : TC_DEFAULT_BK; return v == 32 ? 32 : 64; }();
  static const int mc = [] { const char* e = getenv("STRAPS_TC_MCAST"); return e ? atoi(e) : 0; }();
  int mt;
  { int bn2; tile_cfg(c.cout, &bn2, &mt); }
  const bool k32 = (bk == 32) && mt == 1 && bn <= 128;
  const int kb_elems = k32 ? 32 : BK_TC;
  p.n_kblocks = c.k_eff / kb_elems;
  p.cchunks = c.cin / kb_elems;
  p.kw_count = c.ksize;
  p.stride = c.stride; p.pad = c.pad;
  p.hw_out = c.hout * c.wout; p.wout = c.wout;
  p.th = (c.hout * c.wout >= BM_TC) ? BM_TC / c.wout : c.hout;
  p.cout = c.cout;
  p.debug = g_sw.debug;
  {
    // STRAPS_TC_HALO: "1" = layer1 and layer2 stride-1 3x3 convolutions through conv_halo_kernel, "64" / "128" = only the layers of
    // that width; "8" appended ("1,8") = two epilogue warps per quadrant.
    // Verified on B200 (profiles/r01_halo_check.json) but slower than conv_tc_kernel: off by default.
    const int sel = g_sw.halo;
    // "2": layer1 through conv_halo2_kernel (two tiles per item, deeper weight ring); not yet run on hardware
    if (sel == 2 && m.halo2 && p.out_hi && !p.out_f32 && !p.res_f32 && c.cout == 64 && bn == 64 && c.hin == 64)
      return launch_conv_halo2(m, p, c, B, t->num_sms, st);
    if (sel && sel != 2 && m.halo && p.out_hi && !p.out_f32 && !p.res_f32 && bn == c.cout && (sel == 1 || sel == c.cout)) {
      const bool e8 = g_sw.halo_e8;
      if (c.cout == 64)
        return e8 ? launch_conv_halo<64, 5, 66, 8>(m, p, c, B, t->num_sms, st) : launch_conv_halo<64, 5, 66>(m, p, c, B, t->num_sms, st);
      return e8 ? launch_conv_halo<128, 7, 34, 8>(m, p, c, B, t->num_sms, st) : launch_conv_halo<128, 7, 34>(m, p, c, B, t->num_sms, st);
    }
  }
  {
    // STRAPS_TC_PAIR: "128" = CTA pairs for the Cout >= 128 layers, "all" = every layer, unset/"0" = single-CTA kernels;
    // "m" / "m128" = the same selection with conv_tc2m_kernel (merged wide MMA, lean loops; not yet run on hardware)
    const char* e = g_sw.pair;
    if (e[0] == 'm' && (e[1] == '\0' || c.cout >= 128) && mt == 1 && bn <= 128 && !p.res_f32 && !k32)
      return bn == 64 ? launch_conv_tc2m<64>(m, p, t->num_sms, st) : launch_conv_tc2m<128>(m, p, t->num_sms, st);
    const bool pair = (e[0] == 'a') || (e[0] == '1' && c.cout >= 128);
    if (pair && mt == 1 && bn <= 128 && !p.res_f32 && !k32)
      return bn == 64 ? launch_conv_tc2<64>(m, p, t->num_sms, st) : launch_conv_tc2<128>(m, p, t->num_sms, st);
  }
  if (k32) {
    if (mc == 2) return bn == 64 ? launch_conv_tc_cl<64, 2, 32>(m, p, t->num_sms, st) : launch_conv_tc_cl<128, 2, 32>(m, p, t->num_sms, st);
    return bn == 64 ? launch_conv_tc<64, 1, 32>(m, p, t->num_sms, st) : launch_conv_tc<128, 1, 32>(m, p, t->num_sms, st);
  }
  if ((mc == 2 || mc == 4) && mt == 1 && bn <= 128) {
    if (bn == 64) return mc == 2 ? launch_conv_tc_cl<64, 2>(m, p, t->num_sms, st) : launch_conv_tc_cl<64, 4>(m, p, t->num_sms, st);
    return mc == 2 ? launch_conv_tc_cl<128, 2>(m, p, t->num_sms, st) : launch_conv_tc_cl<128, 4>(m, p, t->num_sms, st);
  }
  {
    // STRAPS_TC_EPI_WARPS = 8: two epilogue warps per TMEM lane quadrant.  Bit-identical on B200 but 4 % slower (encoder 1.651 vs
    // 1.584 ms, profiles/r01_halo_check.json): the epilogue is not short of warps to hide latency; the default stays 4.
    if (g_sw.epw == 8 && mt == 1 && bn <= 128)
      return bn == 64 ? launch_conv_tc<64, 1, 64, 8>(m, p, t->num_sms, st) : launch_conv_tc<128, 1, 64, 8>(m, p, t->num_sms, st);
  }
  {
    // STRAPS_TC_TMA2=1: merged-plane tensor maps, two TMA operations per stage (not yet run on hardware)
    if (g_sw.tma2 == 1 && mt == 1 && bn <= 128 && m.has_merged)
      return bn == 64 ? launch_conv_tc_m2<64>(m, p, t->num_sms, st) : launch_conv_tc_m2<128>(m, p, t->num_sms, st);
  }
  {
    // STRAPS_TC_PDL=1: programmatic dependent launch of the shipped tile shapes (not yet run on hardware)
    if (g_sw.pdl == 1 && mt == 1 && bn <= 128)
      return bn == 64 ? launch_conv_tc_pdl<64>(m, p, t->num_sms, st) : launch_conv_tc_pdl<128>(m, p, t->num_sms, st);
  }
  if (bn == 64) return mt == 2 ? launch_conv_tc<64, 2>(m, p, t->num_sms, st) : launch_conv_tc<64, 1>(m, p, t->num_sms, st);
  if (bn == 256) return launch_conv_tc<256, 1>(m, p, t->num_sms, st);
  if (mt == 2) return launch_conv_tc<128, 2>(m, p, t->num_sms, st);
  return launch_conv_tc<128, 1>(m, p, t->num_sms, st);
}

static int run_conv_tc(straps_regressor* r, const std::vector<TcLayerMaps>& maps, int ci, int B, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  const ConvSpec& c = r->conv[ci];
  TcConvParams p;
  memset(&p, 0, sizeof(p));
  p.shift = c.shift;
  p.unscale = t->unscale + t->ch_off[ci];
  p.relu = c.relu;
  if (out_is_f32(r, ci)) {
    p.out_f32 = act_ptr(r, c.out_buf);
  } else {
    p.out_hi = plane_hi(r, c.out_buf);
    p.out_lo = plane_lo(r, c.out_buf);
  }
  if (c.res_buf >= 0) {
    p.res_hi = plane_hi(r, c.res_buf);
    p.res_lo = plane_lo(r, c.res_buf);
  }
  return run_tc(t, geom_fwd(c, ci), maps[ci], p, B, st);
}

// conv1 through the pixel-pair layout (conv1_s2d_kernel): buffer + tensor maps on first use, then pack + convolution
static int s2d_prepare(straps_regressor* r, int B, const TcState::S2dMaps** out, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  if (ensure_encode(t)) return 1;
  const ConvSpec& c = r->conv[0];
  STRAPS_CHECK(c.cout == 64 && c.k_eff == 7 * C1_KROW && r->c_in <= XP_C, "conv1_s2d_kernel: unexpected conv1 geometry");
  STRAPS_CHECK(static_cast<__half*>(c.w_lo) == static_cast<__half*>(c.w_hi) + (size_t)c.cout * c.k_eff,
               "conv1_s2d_kernel: W_hi and W_lo of conv1 must be adjacent");
  if (!t->xs) {
    const char* e = getenv("STRAPS_TC_S2D_PITCH");
    t->xs_pitch = (e && atoi(e) == 48) ? 48 : 64;
    t->xs_plane = (size_t)r->max_batch * XS_H * XS_PAIRS * t->xs_pitch;
    STRAPS_CUDA(cudaMalloc(&t->xs, 2 * t->xs_plane * sizeof(__half)));
    STRAPS_CUDA(cudaMemsetAsync(t->xs, 0, 2 * t->xs_plane * sizeof(__half), st));      // the 6 border rows stay zero for good
  }
  auto it = t->s2d_maps.find(B);
  if (it == t->s2d_maps.end()) {
    TcState::S2dMaps m;
    const cuuint64_t pitch = (cuuint64_t)t->xs_pitch;
    cuuint64_t dims[4] = {pitch, (cuuint64_t)XS_PAIRS, (cuuint64_t)XS_H, (cuuint64_t)B};
    cuuint64_t str[3] = {pitch * 2, (cuuint64_t)XS_PAIRS * pitch * 2, (cuuint64_t)XS_H * XS_PAIRS * pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)pitch, (cuuint32_t)S2D_LINES, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(t, &m.a_hi, t->xs, 4, dims, str, box, es)) return 1;
    if (encode(t, &m.a_lo, t->xs + t->xs_plane, 4, dims, str, box, es)) return 1;
    cuuint64_t wdims[2] = {(cuuint64_t)c.k_eff, (cuuint64_t)2 * c.cout};          // rows 0..63 = W_hi, 64..127 = W_lo
    cuuint64_t wstr[1] = {(cuuint64_t)c.k_eff * 2};
    cuuint32_t wbox[2] = {64, 128};
    if (encode(t, &m.w, c.w_hi, 2, wdims, wstr, wbox, es)) return 1;
    it = t->s2d_maps.emplace(B, m).first;
  }
  *out = &it->second;
  return 0;
}

template <int MT, bool POOL = false>
static int launch_conv1_s2d(straps_regressor* r, const TcState::S2dMaps& m, int B, cudaStream_t st) {
  using Cfg = S2dCfg<MT>;
  TcState* t = static_cast<TcState*>(r->tc);
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv1_s2d_kernel<MT, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const ConvSpec& c = r->conv[0];
  S2dParams p;
  memset(&p, 0, sizeof(p));
  p.n_rows = B * 128;
  p.n_items = (p.n_rows + MT - 1) / MT;
  p.a_bytes = (uint32_t)(S2D_LINES * t->xs_pitch * 2);
  p.shift = c.shift;
  p.unscale = t->unscale + t->ch_off[0];
  p.out = act_ptr(r, c.out_buf);
  p.pool_hi = plane_hi(r, r->buf_pool);
  p.pool_lo = plane_lo(r, r->buf_pool);
  p.relu = c.relu;
  p.debug = g_sw.debug;
  const int grid = p.n_items < t->num_sms ? p.n_items : t->num_sms;
  conv1_s2d_kernel<MT, POOL><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(m.a_hi, m.a_lo, m.w, p);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

static int run_conv1_s2d(straps_regressor* r, const float* x, int B, int mt, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  const TcState::S2dMaps* m = nullptr;
  if (s2d_prepare(r, B, &m, st)) return 1;
  if (t->xs_pitch == 64) pack_input_s2d_kernel<64><<<dim3(IMG, B), 256, 0, st>>>(x, r->c_in, t->xs, t->xs + t->xs_plane);
  else pack_input_s2d_kernel<48><<<dim3(IMG, B), 256, 0, st>>>(x, r->c_in, t->xs, t->xs + t->xs_plane);
  STRAPS_LAUNCH_CHECK();
  if (mt == 3) return launch_conv1_s2d<2, true>(r, *m, B, st);          // two conv rows per item + fused max pool
  return mt == 2 ? launch_conv1_s2d<2>(r, *m, B, st) : launch_conv1_s2d<1>(r, *m, B, st);
}

int tc_encoder_forward(straps_regressor* r, const float* x, int B, float* feat, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  STRAPS_CHECK(t, "tc_encoder_forward: tensor-core state missing");
  tc_refresh_switches();
  auto it = t->maps.find(B);
  if (it == t->maps.end()) {
    std::vector<TcLayerMaps> v;
    if (build_maps(r, B, v)) return 1;
    it = t->maps.emplace(B, std::move(v)).first;
  }
  const std::vector<TcLayerMaps>& maps = it->second;
  // STRAPS_TC_CONV1 = "s2d" / "s2d2" / "s2dp": conv1 from the pixel-pair layout (one / two output rows per work item / two rows and the
  // max pool fused into the epilogue); all switches are refreshed from the environment at the top of this call.
  // NOT YET RUN ON HARDWARE (see conv1_s2d_kernel); unset = the shipped path.
  const bool s2d = g_sw.conv1 != 0;
  const bool fused_pool = g_sw.conv1 == 3;                      // "s2dp": the max pool runs in conv1's epilogue
  if (s2d) {
    if (run_conv1_s2d(r, x, B, g_sw.conv1, st)) return 1;
  } else {
    pack_input_tc_kernel<<<dim3(IMG, B), 256, 0, st>>>(x, r->c_in, t->xp, t->xp + t->xp_plane);
    STRAPS_LAUNCH_CHECK();
    if (run_conv_tc(r, maps, 0, B, st)) return 1;
  }
  if (!fused_pool) {
    const size_t n = (size_t)B * 64 * 64 * 16;
    maxpool_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(act_ptr(r, r->buf_stem), B, 128, 128, 64,
                                                                    plane_hi(r, r->buf_pool), plane_lo(r, r->buf_pool));
    STRAPS_LAUNCH_CHECK();
  }
  int i = 1;
  while (i < NCONV) {
    const bool ds = (i + 2 < NCONV) && r->conv[i + 2].ksize == 1;
    if (ds && t->ds_overlap) {
      // fork: the 3x3/2 convolution goes first on the caller's stream and takes every SM; the 1x1/2 downsample (same input,
      // own output buffer) follows on the side stream and starts on the SMs freed by the 3x3 kernel's partial last wave
      STRAPS_CUDA(cudaEventRecord(t->ev_fork, st));
      STRAPS_CUDA(cudaStreamWaitEvent(t->side, t->ev_fork, 0));
      if (run_conv_tc(r, maps, i, B, st)) return 1;
      if (run_conv_tc(r, maps, i + 2, B, t->side)) return 1;
      STRAPS_CUDA(cudaEventRecord(t->ev_join, t->side));
      STRAPS_CUDA(cudaStreamWaitEvent(st, t->ev_join, 0));      // join: conv2 adds the downsample output as its residual
    } else {
      if (run_conv_tc(r, maps, i, B, st)) return 1;
      if (ds && run_conv_tc(r, maps, i + 2, B, st)) return 1;
    }
    if (run_conv_tc(r, maps, i + 1, B, st)) return 1;
    i += ds ? 3 : 2;
  }
  avgpool_f32_kernel<<<ceil_div(B * 512, 256), 256, 0, st>>>(act_ptr(r, r->buf_final), B, 64, 512, feat);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

int tc_read_activation(straps_regressor* r, int buf, int batch, float* out, cudaStream_t st) {
  const ActBuf& b = r->bufs[buf];
  const size_t total = (size_t)batch * b.c * b.h * b.w;
  if (buf == r->buf_stem || buf == r->buf_final)
    f32_nhwc_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(act_ptr(r, buf), b.c, b.h, b.w, out, total);
  else
    split_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(plane_hi(r, buf), plane_lo(r, buf), b.c, b.h, b.w, out, total);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// weight gradient on the tensor cores
//   dW[co][(tap,ci)] = sum over output pixels of  X[pixel @ tap][ci] * dY[pixel][co]
// Both operands are read exactly as they lie in HBM -- NHWC, i.e. the reduction index (the pixel) is the SLOW one -- so both
// are MN-major UMMA operands: a TMA box {64 channels, 64 pixels} lands as 64 rows of 128 bytes, which IS the canonical
// SWIZZLE_128B MN-major atom layout (no transposes anywhere).  A = X patches (M = 128 = two 64-wide (tap, ci-chunk) columns of
// the weight matrix, the same 4-D boxes the forward uses, 64 pixels deep), B = dY (N = BN output channels), D = dW^T tile.
// 3-pass fp16 split as in the forward:  X_hi.[dY_hi | dY_lo]  (one MMA of width 2 BN)  +  X_lo.dY_hi , separate accumulators.
// dY is pre-scaled by a power of two (split_scaled_kernel); the inverse is applied by unpack_dw_tc_kernel.
// Split-K: an item = (M-tile, N-tile, range of <= 128 K-tiles of 64 pixels); partial tiles are added to dW[cout][k_eff] with
// coalesced fp32 reductions (a warp = 32 consecutive ci of one co).  Short K ranges also bound the truncation bias of the
// tensor-core accumulator (-1e-8 relative per MMA).
struct WgParams {
  int n_mtiles, n_ntiles, splits, ktiles, kt_per;
  int nchunks;         // 64-wide columns (tap, ci-chunk) of the weight matrix
  int conv1, cchunks, kw_count, stride, pad;
  int hw_out, wout;
  int k_eff;
  float* dw;           // [cout][k_eff], accumulated
};

template <int BN>
struct WgCfg {
  static constexpr int CHUNK_BYTES = 64 * 128;                      // 64 pixels x 64 channels fp16
  static constexpr int A_BYTES = 2 * CHUNK_BYTES;                   // per plane
  static constexpr int B_BYTES = (BN / 64) * CHUNK_BYTES;           // per plane
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;     // 48 KB (BN 64) / 64 KB (BN 128)
  static constexpr int STAGES = (BN == 64) ? 4 : 3;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int ACC_COLS = 2 * BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;                    // two accumulator stages
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                const __grid_constant__ CUtensorMap map_dy_hi, const __grid_constant__ CUtensorMap map_dy_lo, const WgParams p) {
  using Cfg = WgCfg<BN>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_mtiles * p.n_ntiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x_hi); tma_prefetch_desc(&map_x_lo); tma_prefetch_desc(&map_dy_hi); tma_prefetch_desc(&map_dy_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (split, nt, mt): consecutive CTAs share the K range (dY and X tiles hit in L2)
  auto decode = [&](int item, int& mt, int& nt, int& kt0, int& kt1) {
    mt = item % p.n_mtiles;
    const int r = item / p.n_mtiles;
    nt = r % p.n_ntiles;
    const int sp = r / p.n_ntiles;
    kt0 = sp * p.kt_per;
    kt1 = min(p.ktiles, kt0 + p.kt_per);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int mt, nt, kt0, kt1;
        decode(item, mt, nt, kt0, kt1);
        int c0[2], c1o[2], c2o[2];        // per chunk: channel / flattened-run coordinate and the tap offsets
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int q = min(2 * mt + j, p.nchunks - 1);
          if (p.conv1) {
            const int kh = q / 3, jj = q % 3;
            c0[j] = kh * (XP_W * XP_C) + jj * 64; c1o[j] = 0; c2o[j] = 0;
          } else {
            const int tap = q / p.cchunks, cc = q % p.cchunks;
            c0[j] = cc * 64; c1o[j] = tap % p.kw_count - p.pad; c2o[j] = tap / p.kw_count - p.pad;
          }
        }
        for (int kt = kt0; kt < kt1; ++kt, ++it) {
          const int st = it % Cfg::STAGES;
          mbar_wait(&empty[st], ((it / Cfg::STAGES) & 1) ^ 1);
          unsigned char* sa = smem + st * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full[st], Cfg::STAGE_BYTES);
          const long long pix0 = (long long)kt * 64;
          const int b0 = (int)(pix0 / p.hw_out), rem = (int)(pix0 % p.hw_out);
          const int oh0 = rem / p.wout, ow0 = rem % p.wout;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c1 = p.conv1 ? ow0 : ow0 * p.stride + c1o[j];
            const int c2 = p.conv1 ? oh0 : oh0 * p.stride + c2o[j];
            tma_load_4d(sa + j * Cfg::CHUNK_BYTES, &map_x_hi, &full[st], c0[j], c1, c2, b0);
            tma_load_4d(sa + Cfg::A_BYTES + j * Cfg::CHUNK_BYTES, &map_x_lo, &full[st], c0[j], c1, c2, b0);
          }
          unsigned char* sb = sa + 2 * Cfg::A_BYTES;
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) {
            tma_load_2d(sb + c * Cfg::CHUNK_BYTES, &map_dy_hi, &full[st], nt * BN + c * 64, kt * 64);
            tma_load_2d(sb + Cfg::B_BYTES + c * Cfg::CHUNK_BYTES, &map_dy_lo, &full[st], nt * BN + c * 64, kt * 64);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16_mn(128, BN);
      constexpr uint32_t idesc_wide = umma_idesc_f16_mn(128, 2 * BN);
      constexpr uint32_t lbo = Cfg::CHUNK_BYTES, sbo = 1024;     // chunk stride along M/N, 8-pixel atom stride along K
      uint32_t it = 0, ti = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ti) {
        int mt, nt, kt0, kt1;
        decode(item, mt, nt, kt0, kt1);
        const uint32_t as = ti & 1;
        mbar_wait(&tempty[as], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_hi = tmem_base + as * Cfg::ACC_COLS, d_lo = d_hi + BN;
        for (int kt = kt0; kt < kt1; ++kt, ++it) {
          const int st = it % Cfg::STAGES;
          mbar_wait(&full[st], (it / Cfg::STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + st * Cfg::STAGE_BYTES);
          const uint64_t x_hi = umma_desc_mn_sw128(sa, lbo, sbo);
          const uint64_t x_lo = umma_desc_mn_sw128(sa + Cfg::A_BYTES, lbo, sbo);
          const uint64_t dy = umma_desc_mn_sw128(sa + 2 * Cfg::A_BYTES, lbo, sbo);     // [dY_hi chunks | dY_lo chunks]
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ko = (uint64_t)(k * 2048 >> 4);       // 16 pixels = two 8-row swizzle atoms
            umma_f16(d_hi, x_hi + ko, dy + ko, idesc_wide, (kt != kt0) || (k != 0));
            umma_f16(d_lo, x_lo + ko, dy + ko, idesc, 1);
          }
          umma_commit(&empty[st]);
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;            // = chunk_local * 64 + ci_local
    uint32_t ti = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ti) {
      int mt, nt, kt0, kt1;
      decode(item, mt, nt, kt0, kt1);
      const uint32_t as = ti & 1;
      mbar_wait(&tfull[as], (ti >> 1) & 1);
      tc_fence_after();
      const int q = 2 * mt + (row >> 6);
      const bool valid = q < p.nchunks;
      float* dst = p.dw + (size_t)(nt * BN) * p.k_eff + (size_t)q * 64 + (row & 63);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32], vl[32];
        const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + as * Cfg::ACC_COLS + c0;
        tmem_ld_32x32(tacc, v);
        tmem_ld_32x32(tacc + BN, vl);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            atomicAdd(dst + (size_t)(c0 + j) * p.k_eff, __uint_as_float(v[j]) + __uint_as_float(vl[j]));
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// dW[cout][k_eff] (x dY scale) -> OIHW fp32
__global__ void unpack_dw_tc_kernel(const float* __restrict__ dw, const unsigned* __restrict__ maxbits, int cout, int cin, int ks, int conv1,
                                    int k_eff, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)cout * cin * ks * ks) return;
  int e = 0;
  const float mx = __uint_as_float(*maxbits);
  if (mx > 0.f && isfinite(mx)) { frexpf(mx, &e); e = max(-100, min(100, 14 - e)); }
  const int kw = (int)(i % ks);
  size_t t = i / ks;
  const int kh = (int)(t % ks); t /= ks;
  const int c = (int)(t % cin);
  const int n = (int)(t / cin);
  const int k = conv1 ? kh * C1_KROW + kw * XP_C + c : (kh * ks + kw) * cin + c;
  out[i] = dw[(size_t)n * k_eff + k] * ldexpf(1.f, -e);
}

// ------------------------------------------------------------------------------------------------
// training path on the tensor cores (called from train.cu)
//
// forward       : the same conv_tc_kernel with the BatchNorm scale NOT folded (train-mode BN needs the raw conv output for its
//                 batch statistics): weights re-split every step, output = fp32 NHWC `raw`, no shift / residual / ReLU.
// data gradient : dX = conv(dY, W') with W'[ci][(kh',kw',co)] = W[co][ci][ks-1-kh'][ks-1-kw'], stride 1, pad ks-1-pad -- again the
//                 same kernel.  A stride-2 forward conv becomes a stride-1 conv over dY zero-upsampled by 2 (4x redundant MMAs on
//                 6 small layers, 9 % of the data-gradient FLOPs, instead of a second kernel).  dY has no natural scale (1e-2 ..
//                 1e-8), so it is multiplied by the power of two that puts max|dY| into [2^13, 2^14) before the fp16 split; the
//                 exact inverse is folded into the per-channel epilogue scale.  The other branch's gradient (`add`) is summed in
//                 fp32 in the epilogue.
// ------------------------------------------------------------------------------------------------
struct TcTrain {
  __half* pool;
  std::vector<__half*> plane_hi, plane_lo;     // split planes per activation buffer (null: never a conv input)
  __half *wf_hi[NCONV], *wf_lo[NCONV];         // forward weights [cout][k_eff]
  __half *wd_hi[NCONV], *wd_lo[NCONV];         // data-gradient weights [cin][ks*ks*cout]
  float *wf_scale, *wf_unscale;                // [sum cout]
  float *wd_scale, *wd_unscale;                // [sum cin]
  size_t wd_off[NCONV];
  float* dg_unscale;                           // [512] wd_unscale / dY scale of the current data gradient
  float* ones; float* zeros;                   // [512]
  unsigned* dy_max;                            // bits of max |dY| (atomicMax target of bn_bwd_apply_kernel)
  __half *dy_hi, *dy_lo;                       // split (scaled, optionally zero-upsampled) dY of the current layer
  size_t dy_plane;                             // elements per plane
  float* dw_packed;                            // [cout][k_eff] fp32 accumulation target of the weight gradient (largest conv)
  size_t dw_elems;
  std::map<int, std::vector<TcLayerMaps>> fwd_maps, dg_maps;
  std::map<int, std::vector<TcLayerMaps>> wg_maps;   // a_* = X with 64-pixel boxes, w_* = dY [pixels][cout] boxes {64, 64}
};

static void tc_train_free(TcState* t) {
  if (!t->train) return;
  if (t->train->pool) cudaFree(t->train->pool);
  delete t->train;
  t->train = nullptr;
}

// per INPUT channel of the forward conv: power-of-two scale for the rows of the data-gradient weight matrix
__global__ void w_colscale_kernel(const float* __restrict__ w, int cout, int cin, int kk, float* __restrict__ pscale,
                                  float* __restrict__ unscale) {
  __shared__ float red[256];
  const int ci = blockIdx.x;
  float m = 0.f;
  for (int i = threadIdx.x; i < cout * kk; i += blockDim.x) m = fmaxf(m, fabsf(w[((size_t)(i / kk) * cin + ci) * kk + (i % kk)]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mx = red[0];
    int e = 0;
    if (mx > 0.f && isfinite(mx)) { frexpf(mx, &e); e = max(-60, min(60, 14 - e)); }
    pscale[ci] = ldexpf(1.f, e);
    unscale[ci] = ldexpf(1.f, -e);
  }
}
// OIHW fp32 -> [Cin][(kh',kw',co)] fp16 hi/lo with the taps flipped
__global__ void pack_w_dgrad_tc_kernel(const float* __restrict__ w, const float* __restrict__ pscale, int cout, int cin, int ks,
                                       __half* __restrict__ hi, __half* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k_eff = ks * ks * cout;
  if (i >= (size_t)cin * k_eff) return;
  const int ci = (int)(i / k_eff), k = (int)(i % k_eff);
  const int tap = k / cout, co = k % cout, kh = ks - 1 - tap / ks, kw = ks - 1 - tap % ks;
  const float v = w[(((size_t)co * cin + ci) * ks + kh) * ks + kw] * pscale[ci];
  __half h, l;
  split_f16(v, h, l);
  hi[i] = h;
  lo[i] = l;
}
// fp32 NHWC -> split planes, scaled by the power of two derived from *maxbits (null: no scaling); up = 1 writes pixel (h, w) to
// (2h, 2w) of a [B, 2H, 2W, C] tensor (zeroed beforehand).  Block 0 also publishes the epilogue scales of the data gradient.
__global__ void split_scaled_kernel(const float* __restrict__ src, const unsigned* __restrict__ maxbits, long long n4, int C, int H, int W,
                                    int up, __half* __restrict__ hi, __half* __restrict__ lo, const float* __restrict__ w_unscale,
                                    int n_unscale, float* __restrict__ out_unscale) {
  int e = 0;
  if (maxbits) {
    const float mx = __uint_as_float(*maxbits);
    if (mx > 0.f && isfinite(mx)) { frexpf(mx, &e); e = max(-100, min(100, 14 - e)); }
  }
  const float sc = ldexpf(1.f, e);
  if (blockIdx.x == 0 && out_unscale) {
    const float inv = ldexpf(1.f, -e);
    for (int c = threadIdx.x; c < n_unscale; c += blockDim.x) out_unscale[c] = w_unscale[c] * inv;
  }
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(src)[i];
  __half h[4], l[4];
  split_f16(v.x * sc, h[0], l[0]); split_f16(v.y * sc, h[1], l[1]); split_f16(v.z * sc, h[2], l[2]); split_f16(v.w * sc, h[3], l[3]);
  long long o = i * 4;
  if (up) {
    const int c = (int)(o % C);
    long long t = o / C;
    const int w = (int)(t % W); t /= W;
    const int hh = (int)(t % H);
    const long long b = t / H;
    o = (((b * 2 * H + 2 * hh) * 2 * W) + 2 * w) * C + c;
  }
  *reinterpret_cast<uint2*>(hi + o) = make_uint2(pack_f16(h[0], h[1]), pack_f16(h[2], h[3]));
  *reinterpret_cast<uint2*>(lo + o) = make_uint2(pack_f16(l[0], l[1]), pack_f16(l[2], l[3]));
}

static bool buf_is_conv_input(const straps_regressor* r, int buf) {
  for (int i = 1; i < NCONV; ++i)
    if (r->conv[i].in_buf == buf) return true;
  return false;
}

int tc_train_begin(straps_regressor* r, int B, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  STRAPS_CHECK(t, "tc_train_begin: tensor-core state missing");
  tc_refresh_switches();
  if (ensure_encode(t)) return 1;
  if (!t->train) {
    TcTrain* tt = new TcTrain();
    size_t el = 0;   // fp16 elements
    auto take = [&](size_t n) { size_t o = el; el += (n + 511) & ~(size_t)511; return o; };   // 1 KB granules
    const size_t nb = r->bufs.size();
    std::vector<size_t> off_p(nb, (size_t)-1);
    for (size_t b = 0; b < nb; ++b)
      if ((int)b != r->buf_xin && buf_is_conv_input(r, (int)b)) off_p[b] = take(2 * (size_t)r->max_batch * r->bufs[b].h * r->bufs[b].w * r->bufs[b].c);
    size_t off_wf[NCONV], off_wd[NCONV], nco = 0, nci = 0, dy_el = 0, dw_el = 0;
    for (int i = 0; i < NCONV; ++i) {
      const ConvSpec& c = r->conv[i];
      off_wf[i] = take(2 * (size_t)c.cout * c.k_eff);
      off_wd[i] = (i == 0) ? 0 : take(2 * (size_t)c.cin * c.ksize * c.ksize * c.cout);
      tt->wd_off[i] = nci;
      nco += c.cout; nci += (i == 0) ? 0 : c.cin;
      dy_el = std::max(dy_el, (size_t)r->max_batch * c.hout * c.wout * c.cout);
      if (i > 0) dy_el = std::max(dy_el, (size_t)r->max_batch * c.hin * c.win * c.cout);   // (upsampled) dY of this layer
      dw_el = std::max(dw_el, (size_t)c.cout * c.k_eff);
    }
    tt->dy_plane = dy_el;
    const size_t off_dy = take(2 * dy_el);
    const size_t n_f32 = 2 * nco + 2 * nci + 512 * 3 + 64 + dw_el;
    const size_t off_f = take(2 * n_f32);
    if (cudaMalloc(&tt->pool, el * sizeof(__half)) != cudaSuccess) {
      set_error("tc_train_begin: cudaMalloc of %zu bytes failed: %s", el * sizeof(__half), cudaGetErrorString(cudaGetLastError()));
      delete tt;
      return 1;
    }
    tt->plane_hi.assign(nb, nullptr); tt->plane_lo.assign(nb, nullptr);
    for (size_t b = 0; b < nb; ++b)
      if (off_p[b] != (size_t)-1) {
        tt->plane_hi[b] = tt->pool + off_p[b];
        tt->plane_lo[b] = tt->plane_hi[b] + (size_t)r->max_batch * r->bufs[b].h * r->bufs[b].w * r->bufs[b].c;
      }
    for (int i = 0; i < NCONV; ++i) {
      const ConvSpec& c = r->conv[i];
      tt->wf_hi[i] = tt->pool + off_wf[i]; tt->wf_lo[i] = tt->wf_hi[i] + (size_t)c.cout * c.k_eff;
      tt->wd_hi[i] = (i == 0) ? nullptr : tt->pool + off_wd[i];
      tt->wd_lo[i] = (i == 0) ? nullptr : tt->wd_hi[i] + (size_t)c.cin * c.ksize * c.ksize * c.cout;
    }
    tt->dy_hi = tt->pool + off_dy; tt->dy_lo = tt->dy_hi + dy_el;
    float* f = reinterpret_cast<float*>(tt->pool + off_f);
    tt->wf_scale = f; f += nco; tt->wf_unscale = f; f += nco;
    tt->wd_scale = f; f += nci; tt->wd_unscale = f; f += nci;
    tt->dg_unscale = f; f += 512; tt->ones = f; f += 512; tt->zeros = f; f += 512;
    tt->dy_max = reinterpret_cast<unsigned*>(f); f += 64;
    tt->dw_packed = f; tt->dw_elems = dw_el;
    std::vector<float> one(512, 1.f);
    STRAPS_CUDA(cudaMemcpy(tt->ones, one.data(), 512 * sizeof(float), cudaMemcpyHostToDevice));
    STRAPS_CUDA(cudaMemset(tt->zeros, 0, 512 * sizeof(float)));
    STRAPS_CUDA(cudaMemset(tt->dy_max, 0, sizeof(unsigned)));
    t->train = tt;
  }
  TcTrain* tt = t->train;
  // forward weights of this step (the optimiser has changed them): BN scale not folded
  for (int i = 0; i < NCONV; ++i) {
    const ConvSpec& c = r->conv[i];
    const size_t total = (size_t)c.cout * c.k_eff;
    w_rowscale_kernel<<<c.cout, 256, 0, st>>>(c.w_oihw, tt->ones, c.cin * c.ksize * c.ksize, tt->wf_scale + t->ch_off[i],
                                             tt->wf_unscale + t->ch_off[i]);
    STRAPS_LAUNCH_CHECK();
    pack_w_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(c.w_oihw, tt->ones, tt->wf_scale + t->ch_off[i], c.cout, c.cin, c.ksize,
                                                                     i == 0, c.k_eff, tt->wf_hi[i], tt->wf_lo[i]);
    STRAPS_LAUNCH_CHECK();
  }
  if (tt->fwd_maps.find(B) == tt->fwd_maps.end()) {
    std::vector<TcLayerMaps> fm(NCONV), dm(NCONV);
    for (int i = 0; i < NCONV; ++i) {
      const ConvSpec& c = r->conv[i];
      __half* a_hi = (i == 0) ? t->xp : tt->plane_hi[c.in_buf];
      __half* a_lo = (i == 0) ? t->xp + t->xp_plane : tt->plane_lo[c.in_buf];
      STRAPS_CHECK(a_hi && a_lo, "tc_train_begin: conv %d has no input planes", i);
      if (build_layer_maps(t, geom_fwd(c, i), B, a_hi, a_lo, tt->wf_hi[i], tt->wf_lo[i], fm[i])) return 1;
      if (i > 0 && build_layer_maps(t, geom_dgrad(c), B, tt->dy_hi, tt->dy_lo, tt->wd_hi[i], tt->wd_lo[i], dm[i])) return 1;
    }
    std::vector<TcLayerMaps> wm(NCONV);
    for (int i = 0; i < NCONV; ++i) {
      const ConvSpec& c = r->conv[i];
      __half* x_hi = (i == 0) ? t->xp : tt->plane_hi[c.in_buf];
      __half* x_lo = (i == 0) ? t->xp + t->xp_plane : tt->plane_lo[c.in_buf];
      cuuint32_t es1[4] = {1, 1, 1, 1};
      if (i == 0) {
        cuuint64_t dims[4] = {(cuuint64_t)6 * XP_W * XP_C + C1_KROW, 128, 128, (cuuint64_t)B};
        cuuint64_t str[3] = {2 * XP_C * 2, (cuuint64_t)2 * XP_W * XP_C * 2, (cuuint64_t)XP_H * XP_W * XP_C * 2};
        cuuint32_t box[4] = {64, 64, 1, 1};
        if (encode(t, &wm[i].a_hi, x_hi, 4, dims, str, box, es1)) return 1;
        if (encode(t, &wm[i].a_lo, x_lo, 4, dims, str, box, es1)) return 1;
      } else {
        const int wb = c.wout < 64 ? c.wout : 64, hb = 64 / wb;      // 64 output pixels = hb rows of wb
        STRAPS_CHECK(hb <= c.hout && (c.hout * c.wout) % 64 == 0, "weight-gradient tiling: unsupported output size %dx%d", c.hout, c.wout);
        cuuint64_t dims[4] = {(cuuint64_t)c.cin, (cuuint64_t)c.win, (cuuint64_t)c.hin, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)c.cin * 2, (cuuint64_t)c.win * c.cin * 2, (cuuint64_t)c.hin * c.win * c.cin * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(wb * c.stride), (cuuint32_t)(hb * c.stride), 1};
        cuuint32_t es[4] = {1, (cuuint32_t)c.stride, (cuuint32_t)c.stride, 1};
        if (encode(t, &wm[i].a_hi, x_hi, 4, dims, str, box, es)) return 1;
        if (encode(t, &wm[i].a_lo, x_lo, 4, dims, str, box, es)) return 1;
      }
      cuuint64_t ddims[2] = {(cuuint64_t)c.cout, (cuuint64_t)B * c.hout * c.wout};
      cuuint64_t dstr[1] = {(cuuint64_t)c.cout * 2};
      cuuint32_t dbox[2] = {64, 64};
      if (encode(t, &wm[i].w_hi, tt->dy_hi, 2, ddims, dstr, dbox, es1)) return 1;
      if (encode(t, &wm[i].w_lo, tt->dy_lo, 2, ddims, dstr, dbox, es1)) return 1;
    }
    tt->fwd_maps.emplace(B, std::move(fm));
    tt->dg_maps.emplace(B, std::move(dm));
    tt->wg_maps.emplace(B, std::move(wm));
  }
  return 0;
}

int tc_train_pack_input(straps_regressor* r, const float* x, int B, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  pack_input_tc_kernel<<<dim3(IMG, B), 256, 0, st>>>(x, r->c_in, t->xp, t->xp + t->xp_plane);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// padded split planes of the conv1 input -> fp32 NHWC [B,256,256,c_in_pad] (what the fp32 CUDA-core conv1 weight gradient reads);
// only used when fp32 gradients are requested after a tensor-core forward
__global__ void unsplit_input_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int c_pad, float* __restrict__ out, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % c_pad);
  size_t t = i / c_pad;
  const int w = (int)(t % IMG); t /= IMG;
  const int h = (int)(t % IMG);
  const size_t b = t / IMG;
  float v = 0.f;
  if (c < XP_C) {
    const size_t s = ((b * XP_H + h + 3) * XP_W + w + 3) * XP_C + c;
    v = __half2float(hi[s]) + __half2float(lo[s]);
  }
  out[i] = v;
}
int tc_train_unpack_input(straps_regressor* r, int B, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  const size_t total = (size_t)B * IMG * IMG * r->c_in_pad;
  unsplit_input_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(t->xp, t->xp + t->xp_plane, r->c_in_pad, act_ptr(r, r->buf_xin), total);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

__half* tc_train_plane(straps_regressor* r, int buf, int lo) {
  TcState* t = static_cast<TcState*>(r->tc);
  if (!t || !t->train || buf < 0) return nullptr;
  return lo ? t->train->plane_lo[buf] : t->train->plane_hi[buf];
}

int tc_train_split_act(straps_regressor* r, int buf, int B, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  TcTrain* tt = t->train;
  STRAPS_CHECK(tt && tt->plane_hi[buf], "tc_train_split_act: buffer %d has no planes", buf);
  const ActBuf& b = r->bufs[buf];
  const long long n4 = (long long)B * b.h * b.w * b.c / 4;
  split_scaled_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(act_ptr(r, buf), nullptr, n4, b.c, b.h, b.w, 0, tt->plane_hi[buf],
                                                                  tt->plane_lo[buf], nullptr, 0, nullptr);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

int tc_train_conv_fwd(straps_regressor* r, int ci, int B, float* raw, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  TcTrain* tt = t->train;
  const ConvSpec& c = r->conv[ci];
  TcConvParams p;
  memset(&p, 0, sizeof(p));
  p.shift = tt->zeros;
  p.unscale = tt->wf_unscale + t->ch_off[ci];
  p.out_f32 = raw;
  return run_tc(t, geom_fwd(c, ci), tt->fwd_maps.at(B)[ci], p, B, st);
}

int tc_train_pack_dgrad(straps_regressor* r, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  TcTrain* tt = t->train;
  STRAPS_CHECK(tt, "tc_train_pack_dgrad: no tensor-core training forward before this backward");
  for (int i = 1; i < NCONV; ++i) {
    const ConvSpec& c = r->conv[i];
    const int kk = c.ksize * c.ksize;
    w_colscale_kernel<<<c.cin, 256, 0, st>>>(c.w_oihw, c.cout, c.cin, kk, tt->wd_scale + tt->wd_off[i], tt->wd_unscale + tt->wd_off[i]);
    STRAPS_LAUNCH_CHECK();
    const size_t total = (size_t)c.cin * kk * c.cout;
    pack_w_dgrad_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(c.w_oihw, tt->wd_scale + tt->wd_off[i], c.cout, c.cin, c.ksize,
                                                                           tt->wd_hi[i], tt->wd_lo[i]);
    STRAPS_LAUNCH_CHECK();
  }
  return 0;
}

unsigned* tc_train_dy_max(straps_regressor* r) {
  TcState* t = static_cast<TcState*>(r->tc);
  return (t && t->train) ? t->train->dy_max : nullptr;
}

// dy: fp32 NHWC [B,hout,wout,cout] whose max |.| bits are in *dy_max  ->  scaled split planes (plain, or zero-upsampled x2 for the
// data gradient of a stride-2 conv); also publishes the data gradient's epilogue scales
int tc_train_split_dy(straps_regressor* r, int ci, int B, const float* dy, int up, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  TcTrain* tt = t->train;
  const ConvSpec& c = r->conv[ci];
  if (up) {
    STRAPS_CHECK(c.stride == 2 && c.hin == 2 * c.hout && c.win == 2 * c.wout, "tc_train_split_dy: unsupported stride");
    const size_t n = (size_t)B * c.hin * c.win * c.cout;
    STRAPS_CUDA(cudaMemsetAsync(tt->dy_hi, 0, n * sizeof(__half), st));
    STRAPS_CUDA(cudaMemsetAsync(tt->dy_lo, 0, n * sizeof(__half), st));
  }
  const long long n4 = (long long)B * c.hout * c.wout * c.cout / 4;
  split_scaled_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(dy, tt->dy_max, n4, c.cout, c.hout, c.wout, up, tt->dy_hi, tt->dy_lo,
                                                                  ci > 0 ? tt->wd_unscale + tt->wd_off[ci] : nullptr, ci > 0 ? c.cin : 0,
                                                                  ci > 0 ? tt->dg_unscale : nullptr);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// gin = dgrad(split dY) (+ add)
int tc_train_conv_dgrad(straps_regressor* r, int ci, int B, const float* add, float* gin, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  TcTrain* tt = t->train;
  const ConvSpec& c = r->conv[ci];
  TcConvParams p;
  memset(&p, 0, sizeof(p));
  p.shift = tt->zeros;
  p.unscale = tt->dg_unscale;
  p.out_f32 = gin;
  p.res_f32 = add;
  return run_tc(t, geom_dgrad(c), tt->dg_maps.at(B)[ci], p, B, st);
}

template <int BN>
static int launch_wgrad_tc(const TcLayerMaps& m, const WgParams& p, int num_sms, cudaStream_t st) {
  using Cfg = WgCfg<BN>;
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const int items = p.n_mtiles * p.n_ntiles * p.splits;
  wgrad_tc_kernel<BN><<<items < num_sms ? items : num_sms, TC_THREADS, Cfg::SMEM_BYTES, st>>>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, p);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// dw_oihw = wgrad(X planes of the forward, plain split dY of tc_train_split_dy)
int tc_train_conv_wgrad(straps_regressor* r, int ci, int B, float* dw_oihw, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  TcTrain* tt = t->train;
  const ConvSpec& c = r->conv[ci];
  const int bn = c.cout == 64 ? 64 : 128;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.nchunks = c.k_eff / 64;
  p.n_mtiles = (p.nchunks + 1) / 2;
  p.n_ntiles = c.cout / bn;
  p.ktiles = (int)((long long)B * c.hout * c.wout / 64);
  const int tiles = p.n_mtiles * p.n_ntiles;
  int splits = std::min(std::max(1, p.ktiles / 4), (2 * t->num_sms + tiles - 1) / tiles);
  int kt_max = 32;     // <= 128 MMAs per TMEM accumulation chain (the accumulator truncates); longer chains measured no faster
  { const char* e = getenv("STRAPS_WG_KT"); if (e && atoi(e) > 0) kt_max = atoi(e); }
  splits = std::max(splits, (p.ktiles + kt_max - 1) / kt_max);
  p.kt_per = (p.ktiles + splits - 1) / splits;
  p.splits = (p.ktiles + p.kt_per - 1) / p.kt_per;
  p.conv1 = (ci == 0);
  p.cchunks = c.cin / 64; p.kw_count = c.ksize; p.stride = c.stride; p.pad = c.pad;
  p.hw_out = c.hout * c.wout; p.wout = c.wout;
  p.k_eff = c.k_eff;
  p.dw = tt->dw_packed;
  STRAPS_CUDA(cudaMemsetAsync(tt->dw_packed, 0, (size_t)c.cout * c.k_eff * sizeof(float), st));
  const TcLayerMaps& m = tt->wg_maps.at(B)[ci];
  if (bn == 64 ? launch_wgrad_tc<64>(m, p, t->num_sms, st) : launch_wgrad_tc<128>(m, p, t->num_sms, st)) return 1;
  const size_t total = (size_t)c.cout * c.cin * c.ksize * c.ksize;
  unpack_dw_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(tt->dw_packed, tt->dy_max, c.cout, c.cin, c.ksize, ci == 0, c.k_eff, dw_oihw);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

}  // namespace straps

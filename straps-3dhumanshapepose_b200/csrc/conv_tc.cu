// conv_tc.cu -- ResNet-18 encoder on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces the nn.Conv2d -> BatchNorm2d -> (+identity) -> ReLU clusters of the reference's
// models/resnet.py:61-77,201-216 in STRAPS_CONV_F16X3_TC mode.
//
// Numerics: the north-star bar is 1e-4 relative fp32, which a single TF32 or bf16 pass misses.
// Every operand is therefore carried as a 2-term fp16 split x = hi + 2^-11 lo (hi = fp16(x), lo = fp16(2^11 (x - hi)): 11 + 11
// mantissa bits; split_f16 in regressor.h) and each K-block issues three kind::f16 MMAs with fp32 accumulation in TMEM:
//   A_hi.W_hi + A_hi.W_lo + A_lo.W_hi      (dropped lo.lo term and split residual: ~2^-22 relative).
// A first version used a bf16 split (8 + 8 bits): 4e-6 error per layer, 5e-5 at the features after 17 layers -- too
// close to the bar; fp16 has the mantissa and, with the two range guards below, the range:
//   * weights (BatchNorm scale folded in) are multiplied per output channel by a power of two that puts the row
//     maximum in [2^13, 2^14) so that hi AND lo stay fp16-normal; the exact inverse is applied in the epilogue;
//   * activations are clamped to +-65504 before the split (post-BatchNorm ResNet activations are O(1..100)); the low term is
//     stored scaled by 2^11, so it is fp16-normal wherever the high term is (tests/test_gpu_numeric_range.py drives layers
//     down to 1e-3 and up to 6e4; an unscaled lo went subnormal below 0.125 and cost 3e-8 ABSOLUTE).
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
  int epi_slab;     // outputs leave through per-warp shared-memory slabs as full 64 / 128-byte runs per pixel (default on)
};

// BN = output-channel tile (64 for the Cout = 64 layers, else 128); an M-tile is 128 output pixels.
// Measured (tools/tc_probe.cu, profiles/r01_tc_probe.txt): an SS-mode tcgen05.mma M128xNx16 costs 48 / 64 / 128 cycles
// for N = 64 / 128 / 256.  What else was tried on hardware and removed again (profiles/r01_conv_experiments_s2.txt,
// profiles/r02_first_call.json): wider tiles (128 x 2 M-tiles, BN = 256), half-size SWIZZLE_64B stages, weight multicast in
// clusters of 2 / 4, two CTA-pair kernels (cta_group::2), halo kernels for layer1 / layer2, two epilogue warps per TMEM quadrant,
// programmatic dependent launch, merged-plane tensor maps -- all bit-identical (one halo variant wrong), none faster.
template <int BN>
struct TcCfg {
  static constexpr int BK = 64;                         // K elements per stage: one SWIZZLE_128B row
  static constexpr int A_BYTES = BM_TC * BK * 2;        // one plane of the A tile
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
  static constexpr int STAGES = (220 * 1024 / STAGE_BYTES) > 4 ? 4 : (220 * 1024 / STAGE_BYTES);
  static constexpr int SLAB_OFF = STAGES * STAGE_BYTES + 256;     // 4 epilogue warps x 4 KB output slabs, then 4 x 4 KB residual slabs
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + 8 * 4096;
  static constexpr int ACC_COLS = 2 * BN;               // per accumulator stage: [hi.hi | lo terms]
  static constexpr int TMEM_COLS = 2 * ACC_COLS;        // two accumulator stages
  static constexpr int C1_CHUNKS = C1_KROW / BK;        // K-blocks per filter row of conv1 (im2col form)
  // conv1: the 168 real elements of a filter row end inside the last chunk; its all-padding K=16 steps are skipped
  static constexpr int C1_LAST_KSTEPS = (7 * XP_C - (C1_CHUNKS - 1) * BK + 15) / 16;
  static_assert(BN == 64 || BN == 128, "tile shapes of the ResNet-18 layers");
  static_assert(STAGES >= 2 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0 && SMEM_BYTES <= 232448, "bad tile configuration");
  static_assert(2 * STAGES * 8 + 40 <= 256, "barrier block too small");
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const TcConvParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int BK = Cfg::BK;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]
  uint64_t* tfull = bars + 2 * Cfg::STAGES;    // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_mtiles * p.n_ntiles;     // a work item = one M-tile x one N-tile
  const int item0 = (int)blockIdx.x, item_step = (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
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
        const long long pix0 = (long long)mg * BM_TC;   // tiles past the end land out of bounds -> zeros
        const int b0 = (int)(pix0 / p.hw_out);
        const int oh0 = (int)((pix0 % p.hw_out) / p.wout);
        const int row0 = conv1 ? oh0 : oh0 * stride - pad;
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
          tma_load_4d_u32(sa, &map_a_hi, fb, c0, c1, row0 + dh, b0);
          tma_load_4d_u32(sa + Cfg::A_BYTES, &map_a_lo, fb, c0, c1, row0 + dh, b0);
          tma_load_2d_u32(sa + 2 * Cfg::A_BYTES, &map_w_hi, fb, wk, wrow);
          tma_load_2d_u32(sa + 2 * Cfg::A_BYTES + Cfg::W_BYTES, &map_w_lo, fb, wk, wrow);
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
      // W_hi and W_lo are adjacent in the stage, so ONE MMA of width 2*BN computes A_hi.[W_hi ; W_lo] into [acc_hi | acc_lo]
      // (A_hi is read from shared memory once instead of twice).
      // Two accumulators per tile: tensor-core fp32 accumulation truncates (measured -1e-8 relative per MMA, 5.6e-5
      // at the features with a single accumulator), so the small lo-terms get their own accumulator and only
      // K/16 additions happen at full magnitude; the two are summed in fp32 in the epilogue.
      constexpr uint32_t idesc = umma_idesc_f16(BM_TC, BN);
      constexpr uint32_t idesc_wide = umma_idesc_f16(BM_TC, 2 * BN);
      constexpr uint32_t STAGE16 = Cfg::STAGE_BYTES >> 4, A16 = Cfg::A_BYTES >> 4;
      const uint32_t desc0 = umma_desc_sw128_lo(smem0);           // low descriptor word of stage 0, A tile, hi plane
      const uint32_t tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
      uint32_t st = 0, ph = 0, as = 0, aph = 1;
      const int nkb = p.n_kblocks;
      const bool conv1 = p.conv1 != 0;
      for (int item = item0; item < n_items; item += item_step) {
        mbar_wait_u32(tempty0 + as * 8, aph);
        tc_fence_after();
        const uint32_t d_hi = tmem_base + as * Cfg::ACC_COLS, d_lo = d_hi + BN;
        int sub = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_u32(full0 + st * 8, ph);
          tc_fence_after();
          const uint32_t a_hi = desc0 + st * STAGE16, a_lo = a_hi + A16;
          const uint32_t w_hi = a_hi + 2 * A16;
          // conv1: the last chunk of a filter row (192 elements for 7 taps x 24 channels = 168 real ones) ends in zero-weight
          // padding; its all-padding K=16 steps are skipped (11 of 12 MMAs per filter row)
          int ksteps = BK / 16;
          if (conv1) { if (++sub == Cfg::C1_CHUNKS) { sub = 0; ksteps = Cfg::C1_LAST_KSTEPS; } }
#ifdef STRAPS_TC_EXPERIMENTS
          if (TC_NO_MMA(p)) ksteps = 0;                                   // timing experiment: no MMAs
#endif
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (k < ksteps) {
              const uint32_t ko = k * 2;                 // +32 bytes along K inside the 128-byte swizzle row
              const uint32_t accum = (kb | k) != 0;
              umma_f16_lohi(d_hi, a_hi + ko, w_hi + ko, idesc_wide, accum, UMMA_DESC_SW128_HI);   // -> [acc_hi | acc_lo]
              umma_f16_lohi(d_lo, a_lo + ko, w_hi + ko, idesc, 1, UMMA_DESC_SW128_HI);
            }
          }
          umma_commit_u32(empty0 + st * 8);         // frees the smem stage once these MMAs have read it
          if (++st == Cfg::STAGES) { st = 0; ph ^= 1; }
        }
        umma_commit_u32(tfull0 + as * 8);           // accumulators complete -> epilogue
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps 2..5 =================
    const int quad = warp & 3;                   // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    constexpr int NCHUNK = BN / 32;              // 32-column chunks per work item
    uint4* slab = reinterpret_cast<uint4*>(smem + Cfg::SLAB_OFF) + (warp - 2) * 256;          // output slab of this warp (4 KB)
    uint4* rslab = reinterpret_cast<uint4*>(smem + Cfg::SLAB_OFF) + (4 + warp - 2) * 256;     // residual slab (hi 2 KB | lo 2 KB)
    uint32_t ti = 0;
    for (int item = item0; item < n_items; item += item_step, ++ti) {
      const int mg = item / p.n_ntiles, nt = item % p.n_ntiles;
      const uint32_t as = ti & 1;
      const long long m = (long long)mg * BM_TC + row;
      const bool valid = m < p.m_total;
      const size_t obase = (size_t)m * p.cout + (size_t)nt * BN;
      // The residual (identity) tile does not depend on the accumulator: its loads are software-pipelined one chunk ahead, and
      // the first chunk is requested BEFORE waiting for the MMAs -- the in-situ experiment (profiles/r01_conv_experiments.txt)
      // showed the un-pipelined residual latency exposed on the short-K layers.
      uint4 rh[4], rl[4];
      auto fetch_residual = [&](int chunk) {
        if (p.epi_slab) {
          // four lanes fetch one pixel's 64-byte run per plane straight into the warp's residual slab (cp.async, full sectors, no
          // registers held across the MMA wait); the owner reads its pixel back when the chunk is processed
          if (p.res_hi && chunk < NCHUNK && TC_EPI_IO(p)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int pix = (lane >> 2) + 8 * k, part = lane & 3;
              const long long mm = (long long)mg * BM_TC + quad * 32 + pix;
              if (mm < p.m_total) {
                const size_t o = (size_t)mm * p.cout + (size_t)nt * BN + chunk * 32 + part * 8;
                const int u = pix * 4 + (part ^ ((pix >> 1) & 3));
                cp_async_16(&rslab[u], p.res_hi + o);
                cp_async_16(&rslab[128 + u], p.res_lo + o);
              }
            }
          }
          cp_async_commit();
          return;
        }
        if (p.res_hi && chunk < NCHUNK && valid && TC_EPI_IO(p)) {
          const size_t o = obase + chunk * 32;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            rh[q] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + o) + q);
            rl[q] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + o) + q);
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
          y[q * 4 + 0] = fmaf(fmaf(__uint_as_float(vl[q * 4 + 0]), LO_UNSCALE, __uint_as_float(v[q * 4 + 0])), u4.x, s4.x);
          y[q * 4 + 1] = fmaf(fmaf(__uint_as_float(vl[q * 4 + 1]), LO_UNSCALE, __uint_as_float(v[q * 4 + 1])), u4.y, s4.y);
          y[q * 4 + 2] = fmaf(fmaf(__uint_as_float(vl[q * 4 + 2]), LO_UNSCALE, __uint_as_float(v[q * 4 + 2])), u4.z, s4.z);
          y[q * 4 + 3] = fmaf(fmaf(__uint_as_float(vl[q * 4 + 3]), LO_UNSCALE, __uint_as_float(v[q * 4 + 3])), u4.w, s4.w);
        }
        if (p.epi_slab && p.res_hi && TC_EPI_IO(p)) {       // this chunk's residual has landed in the slab: own pixel back into registers
          cp_async_wait_all();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int u = lane * 4 + (q ^ ((lane >> 1) & 3));
            rh[q] = rslab[u];
            rl[q] = rslab[128 + u];
          }
          __syncwarp();                                     // the next fetch may overwrite the slab
        }
        if (valid && TC_EPI_IO(p)) {
          if (p.res_hi) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t hw[4] = {rh[q].x, rh[q].y, rh[q].z, rh[q].w}, lw[4] = {rl[q].x, rl[q].y, rl[q].z, rl[q].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                y[q * 8 + e * 2 + 0] += fmaf(f16lo_to_f(lw[e]), LO_UNSCALE, f16lo_to_f(hw[e]));
                y[q * 8 + e * 2 + 1] += fmaf(f16hi_to_f(lw[e]), LO_UNSCALE, f16hi_to_f(hw[e]));
              }
            }
          }
        }
        fetch_residual(chunk + 1);        // in flight while this chunk is split and stored
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
            if (!p.epi_slab) {
              float4* o = reinterpret_cast<float4*>(p.out_f32 + obase + c0);
#pragma unroll
              for (int q = 0; q < 8; ++q) o[q] = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
            } else {
              // own pixel's 128 bytes into the warp's slab (unit u of pixel r at unit u ^ (r & 7): conflict free)
              float4* sf = reinterpret_cast<float4*>(slab);
#pragma unroll
              for (int q = 0; q < 8; ++q) sf[lane * 8 + (q ^ (lane & 7))] = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
            }
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
            if (!p.epi_slab) {
              uint4* oh = reinterpret_cast<uint4*>(p.out_hi + obase + c0);
              uint4* ol = reinterpret_cast<uint4*>(p.out_lo + obase + c0);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                oh[q] = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
                ol[q] = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
              }
            } else {
              // own pixel's 64 bytes per plane into the warp's slab (16-byte unit u of pixel r at unit u ^ ((r >> 1) & 3): conflict free)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int u = q ^ ((lane >> 1) & 3);
                slab[lane * 4 + u] = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
                slab[128 + lane * 4 + u] = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
              }
            }
          }
        }
        if (p.epi_slab && p.out_f32 && TC_EPI_IO(p)) {
          // fp32 output (training forward, data gradient): eight lanes write one pixel's 128-byte run
          __syncwarp();
          const float4* sf = reinterpret_cast<const float4*>(slab);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int pix = (lane >> 3) + 4 * k, part = lane & 7;
            const long long mm = (long long)mg * BM_TC + quad * 32 + pix;
            if (mm < p.m_total)
              *reinterpret_cast<float4*>(p.out_f32 + (size_t)mm * p.cout + (size_t)nt * BN + c0 + part * 4) = sf[pix * 8 + (part ^ (pix & 7))];
          }
          __syncwarp();
        }
        if (p.epi_slab && !p.out_f32 && TC_EPI_IO(p)) {
          // four lanes write one pixel's 64-byte run: every store instruction covers 16 FULL sectors (a thread storing its own
          // pixel's 16 bytes at a time writes 32 half sectors per instruction, each sector in two visits)
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int pix = (lane >> 2) + 8 * k, part = lane & 3;
            const long long mm = (long long)mg * BM_TC + quad * 32 + pix;
            if (mm < p.m_total) {
              const size_t o = (size_t)mm * p.cout + (size_t)nt * BN + c0 + part * 8;
              const int u = pix * 4 + (part ^ ((pix >> 1) & 3));
              *reinterpret_cast<uint4*>(p.out_hi + o) = slab[u];
              *reinterpret_cast<uint4*>(p.out_lo + o) = slab[128 + u];
            }
          }
          __syncwarp();
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


// ---------------------------------------------------------------------------------------------------------
// conv1 (7x7 / stride 2 / pad 3, Cout = 64) from a pixel-PAIR layout of the padded input -- the shipped stem since round 2
// (STRAPS_TC_CONV1 = "s2dp" default: max pool fused; "s2d": stem tensor written + separate pool; "im2col": the round-1 kernel).
// Verified on B200 bit-identical to the im2col form of conv_tc_kernel (profiles/r02_first_call.json) and 0.16 ms faster.
//
// conv_tc_kernel reads conv1's A operand as 21 K-blocks of 128 x 64 elements per output row (672 KB of A + 336 KB of W per tile in
// 84 TMA operations) although an output row only depends on 7 input rows, and it was paced by exactly that TMA stream (516 us with
// the MMAs off, 307 us with the TMA traffic off).  Here the padded input is stored as pixel PAIRS:
// xs[b][ph][q][0..23] = padded pixel 2q, [24..47] = padded pixel 2q+1 (ph = h + 3, padded pixel = w + 3), dense 96-byte pair lines
// in HBM; the TMA box {48 elements, 131 lines} lands as 131 SWIZZLE_128B lines of 128 bytes with 96 filled.  Output pixel ow and
// filter tap pair kw' = kw / 2 read pair ow + kw', so for one filter row kh the whole A operand of an output row is ONE box
// {pairs 0..130 of input row 2 oh + kh}, and tap pair kw' is the 128-line window that starts at line kw' (descriptor start address
// + kw' * 128 B -- what tools/shift_probe.cu established for SWIZZLE_128B descriptors).
// K inside a filter row is ordered (kw, c) = kw' * 48 + (kw & 1) * 24 + c -- exactly the order of the im2col conv1 weight rows
// (C1_KROW = 192 per filter row, kw = 7 zero), so the packed weights are shared with conv_tc_kernel.  K step j = 0..10 of a filter
// row (16 elements; step 11 is all padding) reads A at window j / 3, byte offset (j % 3) * 32, and W at chunk j / 4, byte offset
// (j % 4) * 32.  W_hi and W_lo of conv1 are adjacent in HBM, so one 2-D box {64, 128} of the [2 x 64][1344] view brings the
// [W_hi ; W_lo] chunk of the wide MMA in one operation.
// Per output row: 7 x (2 x 12.3 KB A + 48 KB W) = 508 KB in 35 TMA operations against 1008 KB in 84.  Two rings: NA slots of one
// input row (hi + lo plane), NW slots of one weight chunk.  (Two output rows per item sharing the weight chunks measured 0.06 ms
// SLOWER and was removed.)
// ---------------------------------------------------------------------------------------------------------
constexpr int XS_H = 262, XS_PAIRS = 132;          // padded rows / pixel pairs per padded row of the pair layout
constexpr int XS_PITCH = 48;                       // elements per pair line in HBM
constexpr int S2D_LINES = 131;                     // pairs 0..130 serve output pixels 0..127 with tap pairs 0..3

struct S2dParams {
  int n_rows;             // batch * 128 output rows = work items
  const float* shift;
  const float* unscale;
  float* out;             // !POOL: NHWC fp32 [B,128,128,64]
  __half* pool_hi;        // POOL: 3x3 / stride 2 max-pooled output, split planes [B,64,64,64] (the stem tensor is never written)
  __half* pool_lo;
  int relu;
  int debug;              // STRAPS_TC_DEBUG bit mask (experiment builds only)
};

struct S2dCfg {
  static constexpr int A_PLANE = 17 * 1024;                         // 131 lines x 128 B = 16,768 B; planes start on swizzle atoms
  static constexpr int A_SLOT = 2 * A_PLANE;                        // hi + lo plane of one input row
  static constexpr int A_TX = 2 * S2D_LINES * XS_PITCH * 2;         // bytes the two boxes of a slot deliver
  static constexpr int NA = 3;
  static constexpr int W_SLOT = 128 * 128;                          // [W_hi (64 rows) ; W_lo (64 rows)] x 64 K elements
  static constexpr int NW = 6;
  static constexpr int W_OFF = NA * A_SLOT;
  static constexpr int BAR_OFF = W_OFF + NW * W_SLOT;
  static constexpr int EDGE_BYTES = 2 * 4 * 32 * 4;                 // POOL: [channel chunk][quadrant][32] floats of the warps' last pixels
  static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 256 + EDGE_BYTES;
  static constexpr int ACC_COLS = 128;                              // [acc_hi (64) | acc_lo (64)]
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(S2D_LINES * 128 <= A_PLANE, "A plane too small");
  static_assert(3 + 128 <= S2D_LINES, "the last tap window must lie inside the box");
  static_assert(SMEM_BYTES <= 232448 && 2 * (NA + NW) * 8 + 40 <= 256, "conv1 pair-layout configuration does not fit");
};

// POOL: the 3x3 / stride 2 / pad 1 max pool is fused into the epilogue and the 268 MB stem tensor is never written or re-read.
// Pooled row pr = max over conv rows 2 pr - 1, 2 pr, 2 pr + 1 of their horizontally pooled pixels (max commutes), so every CTA walks
// a CONTIGUOUS, even-aligned range of conv rows and keeps the running vertical maximum in registers: an even row joins it, an odd
// row completes pooled row (oh - 1) / 2, stores it and becomes the "row above" of the next one.  A range that starts inside an
// image is preceded by the odd row above it, computed only for that carry (1 extra row in ~56).
// Horizontal pooling: a thread owns one conv pixel (TMEM lane); even lanes combine lanes - 1 / + 1 with warp shuffles, and the one
// neighbour that lives in another warp (pixel 32 q - 1) travels through 128 bytes of shared memory per warp and channel chunk.
template <bool POOL>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv1_s2d_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_w, const S2dParams p) {
  using Cfg = S2dCfg;
  // items (conv rows) of this CTA: strided over the grid, or (POOL) one contiguous even-aligned range, preceded by the row above
  int it0, it1, itstep, it_real;
  if constexpr (POOL) {
    int per = (p.n_rows + (int)gridDim.x - 1) / (int)gridDim.x;
    per += per & 1;
    const int first = (int)blockIdx.x * per;
    it1 = min(first + per, p.n_rows);
    it0 = first - (((first & 127) != 0 && first < it1) ? 1 : 0);
    it_real = first;
    itstep = 1;
  } else {
    it0 = (int)blockIdx.x; it1 = p.n_rows; itstep = (int)gridDim.x; it_real = 0;
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
      // ================= TMA producer: per filter row the input row of the item, then the 3 weight chunks =================
      uint32_t as = 0, aph = 1, ws = 0, wph = 1;
      for (int item = it0; item < it1; item += itstep) {
        const int b = item >> 7, oh = item & 127;         // output row index over (b, oh); 128 rows per image
        for (int kh = 0; kh < 7; ++kh) {
          {
            const uint32_t sa = smem0 + as * Cfg::A_SLOT, fb = a_full0 + as * 8;
            mbar_wait_u32(a_empty0 + as * 8, aph);
            if (TC_NO_TMA(p)) {
              mbar_arrive(&a_full[as]);
            } else {
              mbar_expect_tx_u32(fb, Cfg::A_TX);
              tma_load_4d_u32(sa, &map_a_hi, fb, 0, 0, 2 * oh + kh, b);
              tma_load_4d_u32(sa + Cfg::A_PLANE, &map_a_lo, fb, 0, 0, 2 * oh + kh, b);
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
        const uint32_t d_hi = tmem_base + acs * Cfg::ACC_COLS, d_lo = d_hi + 64;
        uint32_t accum = 0;
        for (int kh = 0; kh < 7; ++kh) {
          mbar_wait_u32(a_full0 + as * 8, aph);
          tc_fence_after();
          const uint32_t a_row = adesc0 + as * (Cfg::A_SLOT >> 4);
          uint32_t wd = 0;
#pragma unroll
          for (int j = 0; j < 11; ++j) {
            if ((j & 3) == 0) {
              mbar_wait_u32(w_full0 + ws * 8, wph);
              tc_fence_after();
              wd = wdesc0 + ws * (Cfg::W_SLOT >> 4);
            }
            const uint32_t w_k = wd + (j & 3) * 2;                         // +32 bytes along K inside the weight chunk
            const uint32_t a_hi = a_row + (j / 3) * 8 + (j % 3) * 2;       // window (j / 3) lines down, +32 bytes along K inside the pair
            const uint32_t a_lo = a_hi + (Cfg::A_PLANE >> 4);
            if (!TC_NO_MMA(p)) {
              umma_f16_lohi(d_hi, a_hi, w_k, idesc_wide, accum);           // A_hi.[W_hi ; W_lo] -> [acc_hi | acc_lo]
              umma_f16_lohi(d_lo, a_lo, w_k, idesc, 1);                    // A_lo.W_hi
            }
            accum = 1;
            if ((j & 3) == 3 || j == 10) {
              umma_commit_u32(w_empty0 + ws * 8);
              if (++ws == Cfg::NW) { ws = 0; wph ^= 1; }
            }
          }
          umma_commit_u32(a_empty0 + as * 8);
          if (++as == Cfg::NA) { as = 0; aph ^= 1; }
        }
        umma_commit_u32(tfull0 + acs * 8);
        if (++acs == 2) { acs = 0; acph ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps 2..5: BatchNorm shift, ReLU, then fp32 NHWC stores or the fused max pool =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ti = 0;
    // y[32] of one channel chunk of this thread's pixel: TMEM -> (acc_hi + acc_lo) * unscale + shift -> ReLU
    auto load_chunk = [&](uint32_t acs, int c0, float (&y)[32]) {
      uint32_t v[32], vl[32];
      const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + acs * Cfg::ACC_COLS + c0;
      tmem_ld_32x32(tacc, v);
      tmem_ld_32x32(tacc + 64, vl);
      tmem_ld_wait();
      const float4* sh4 = reinterpret_cast<const float4*>(p.shift + c0);
      const float4* us4 = reinterpret_cast<const float4*>(p.unscale + c0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 s4 = __ldg(sh4 + q), u4 = __ldg(us4 + q);
        y[q * 4 + 0] = fmaf(fmaf(__uint_as_float(vl[q * 4 + 0]), LO_UNSCALE, __uint_as_float(v[q * 4 + 0])), u4.x, s4.x);
        y[q * 4 + 1] = fmaf(fmaf(__uint_as_float(vl[q * 4 + 1]), LO_UNSCALE, __uint_as_float(v[q * 4 + 1])), u4.y, s4.y);
        y[q * 4 + 2] = fmaf(fmaf(__uint_as_float(vl[q * 4 + 2]), LO_UNSCALE, __uint_as_float(v[q * 4 + 2])), u4.z, s4.z);
        y[q * 4 + 3] = fmaf(fmaf(__uint_as_float(vl[q * 4 + 3]), LO_UNSCALE, __uint_as_float(v[q * 4 + 3])), u4.w, s4.w);
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = fmaxf(y[i], 0.f);
      }
    };
    if constexpr (POOL) {
      float* edge = reinterpret_cast<float*>(smem + Cfg::BAR_OFF + 256);
      const float ninf = __int_as_float(0xff800000);
      const int pc = quad * 16 + (lane >> 1);                 // pooled column of an even lane
      float vmax[2][32];                                      // even lanes: running vertical max of the pooled row in progress
#pragma unroll
      for (int i = 0; i < 32; ++i) vmax[0][i] = vmax[1][i] = ninf;
      for (int item = it0; item < it1; ++item, ++ti) {
        const uint32_t acs = ti & 1;
        const int oh = item & 127;
        const bool odd = (oh & 1) != 0;
        const bool store = odd && item >= it_real && TC_EPI_IO(p);
        mbar_wait(&tfull[acs], (ti >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int c0 = c * 32;
          float y[32];
          load_chunk(acs, c0, y);
          if (c == 1) {                                       // every TMEM read of the item is done: the issuer may reuse the stage
            tc_fence_before();
            mbar_arrive(&tempty[acs]);
          }
          // the warp's last pixel is the left neighbour of the next warp's first pooled column
          float* e = edge + (c * 4 + quad) * 32;
          if (lane == 31) {
#pragma unroll
            for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(e)[q] = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps only
          const float* ep = e - 32;                           // quadrant quad - 1 (read by lane 0 of quadrants 1..3 only)
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float up = __shfl_up_sync(0xffffffffu, y[i], 1);
            const float dn = __shfl_down_sync(0xffffffffu, y[i], 1);
            if (lane == 0) up = quad ? ep[i] : ninf;          // pixel -1 is padding
            const float h = fmaxf(fmaxf(up, y[i]), dn);
            if (odd) { o[i] = fmaxf(vmax[c][i], h); vmax[c][i] = h; }
            else { o[i] = 0.f; vmax[c][i] = (oh == 0) ? h : fmaxf(vmax[c][i], h); }
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
            const size_t ob = ((size_t)(item >> 1) * 64 + pc) * 64 + c0;         // item >> 1 = b * 64 + pooled row
            uint4* oh4 = reinterpret_cast<uint4*>(p.pool_hi + ob);
            uint4* ol4 = reinterpret_cast<uint4*>(p.pool_lo + ob);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              oh4[q] = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
              ol4[q] = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
            }
          }
        }
      }
    } else {
      for (int item = it0; item < it1; item += itstep, ++ti) {
        const uint32_t acs = ti & 1;
        mbar_wait(&tfull[acs], (ti >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float y[32];
          load_chunk(acs, c * 32, y);
          if (TC_EPI_IO(p)) {
            float4* o = reinterpret_cast<float4*>(p.out + ((size_t)item * BM_TC + row) * 64 + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty[acs]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// x NCHW fp32 [B,C,256,256] -> pixel-pair planes xs[B][262][132][48] (hi, lo) for conv1_s2d_kernel.  One CTA per (b, h) image row:
// the C channel rows are read with coalesced 1 KB requests, split to fp16 hi / lo and assembled in shared memory as the complete
// padded row (pairs 0..131: the 3-pixel left border, the interior and the right border are written as zeros, so only the 6 border
// ROWS rely on the one-time memset), which then leaves as contiguous 16-byte stores.  HBM-bound: 4*C + 2*48 bytes per pixel.
__global__ void __launch_bounds__(256) pack_input_s2d_kernel(const float* __restrict__ x, int C, __half* __restrict__ hi,
                                                             __half* __restrict__ lo) {
  constexpr int WPP = XS_PITCH / 2;                                // 32-bit words per pair line
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
#pragma unroll
  for (int c = 0; c < XP_C; c += 2) {
    __half h0, l0, h1, l1;
    split_f16(v[c], h0, l0);
    split_f16(v[c + 1], h1, l1);
    const int o = pair * WPP + (pw & 1) * (XP_C / 2) + c / 2;
    srow[0][o] = pack_f16(h0, h1);
    srow[1][o] = pack_f16(l0, l1);
  }
  __syncthreads();
  const size_t base = ((size_t)b * XS_H + h + 3) * XS_PAIRS * XS_PITCH;      // 16-byte aligned
  uint4* dh = reinterpret_cast<uint4*>(hi + base);
  uint4* dl = reinterpret_cast<uint4*>(lo + base);
  const uint4* sh = reinterpret_cast<const uint4*>(srow[0]);
  const uint4* sl = reinterpret_cast<const uint4*>(srow[1]);
  for (int i = w; i < ROW_WORDS / 4; i += 256) {
    dh[i] = sh[i];
    dl[i] = sl[i];
  }
}

// SURVEY.md 8f row N1 fused into the stem's input pack: the proxy representation [B, 1 + J, 256, 256] (binary silhouette + J Gaussian
// joint heat-maps, reference utils/label_conversions.py:48-55,90-127 assembled at train/train_synthetic_otf_rendering.py:178-182)
// is never materialised in fp32 -- this kernel writes conv1's pixel-pair planes straight from the part labels and the 2-D joints.
// Same values as binary_labels_kernel / heatmaps_paste_kernel followed by pack_input_s2d_kernel (bit-identical planes): channel 0 =
// (label != 0), channel 1 + j = the reference's (2 size)^2 window `table` pasted at the truncated joint position with its clipping
// rules, zero elsewhere.  One CTA per (b, h) image row, one thread per pixel; HBM: 4 B read + 2 * 96 B / 2 written per pixel.
__global__ void __launch_bounds__(256) pack_proxy_s2d_kernel(const float* __restrict__ seg, const float* __restrict__ joints2d, int J,
                                                             const float* __restrict__ table, int size, __half* __restrict__ hi,
                                                             __half* __restrict__ lo) {
  constexpr int WPP = XS_PITCH / 2;
  constexpr int ROW_WORDS = XS_PAIRS * WPP;
  __shared__ __align__(16) uint32_t srow[2][ROW_WORDS];
  __shared__ int win[XP_C][4];                                     // per joint: first column, width, table offset of this row (or width 0)
  const int h = blockIdx.x, b = blockIdx.y, w = threadIdx.x;
  for (int i = w; i < 2 * ROW_WORDS; i += 256) (&srow[0][0])[i] = 0u;
  if (w < J) {
    const float fx = joints2d[((size_t)b * J + w) * 2 + 0], fy = joints2d[((size_t)b * J + w) * 2 + 1];
    const int cx = (int)fx, cy = (int)fy;                          // Tensor.int(): truncation toward zero
    int ww = 0, hsx = 0, toff = 0;
    if (cx > -size && cy > -size && cx < IMG - 1 + size && cy < IMG - 1 + size) {
      const int hex = min(IMG - 1, cx + size), hsy = max(0, cy - size), hey = min(IMG - 1, cy + size);
      hsx = max(0, cx - size);
      const int gsx = max(0, size - cx), gex = min(2 * size, 2 * size - (size + cx - (IMG - 1)));
      const int gsy = max(0, size - cy), gey = min(2 * size, 2 * size - (size + cy - (IMG - 1)));
      const int wd = min(hex - hsx, gex - gsx), ht = min(hey - hsy, gey - gsy);
      if (h >= hsy && h < hsy + ht) { ww = wd; toff = (gsy + h - hsy) * 2 * size + gsx; }
    }
    win[w][0] = hsx; win[w][1] = ww; win[w][2] = toff;
  }
  float v[XP_C];
  v[0] = (__ldg(seg + ((size_t)b * IMG + h) * IMG + w) != 0.f) ? 1.f : 0.f;
  __syncthreads();
#pragma unroll
  for (int c = 1; c < XP_C; ++c) {
    float val = 0.f;
    if (c - 1 < J) {
      const int x0 = win[c - 1][0], wd = win[c - 1][1];
      if (w >= x0 && w < x0 + wd) val = __ldg(table + win[c - 1][2] + (w - x0));
    }
    v[c] = val;
  }
  const int pw = w + 3;
  const int pair = pw >> 1;
#pragma unroll
  for (int c = 0; c < XP_C; c += 2) {
    __half h0, l0, h1, l1;
    split_f16(v[c], h0, l0);
    split_f16(v[c + 1], h1, l1);
    const int o = pair * WPP + (pw & 1) * (XP_C / 2) + c / 2;
    srow[0][o] = pack_f16(h0, h1);
    srow[1][o] = pack_f16(l0, l1);
  }
  __syncthreads();
  const size_t base = ((size_t)b * XS_H + h + 3) * XS_PAIRS * XS_PITCH;
  uint4* dh = reinterpret_cast<uint4*>(hi + base);
  uint4* dl = reinterpret_cast<uint4*>(lo + base);
  const uint4* sh = reinterpret_cast<const uint4*>(srow[0]);
  const uint4* sl = reinterpret_cast<const uint4*>(srow[1]);
  for (int i = w; i < ROW_WORDS / 4; i += 256) {
    dh[i] = sh[i];
    dl[i] = sl[i];
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
  x[i] = unsplit_f16(hi[s], lo[s]);
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

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


struct TcLayerMaps {
  CUtensorMap a_hi, a_lo, w_hi, w_lo;
};

struct TcState {
  EncodeTiledFn encode;
  int num_sms;
  __half* xp;            // padded conv1 input planes of the im2col form (hi then lo), for max_batch: training path + STRAPS_TC_CONV1=im2col
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
  // pixel-pair input planes of conv1_s2d_kernel: allocated on first use, for max_batch
  __half* xs;
  size_t xs_plane;
  struct S2dMaps { CUtensorMap a_hi, a_lo, w; };
  std::map<int, S2dMaps> s2d_maps;
  bool stem_valid;              // the last inference forward wrote the fp32 stem tensor (false with the fused max pool)
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
  t->xs = nullptr; t->xs_plane = 0; t->stem_valid = false;
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

// Kernel-selection switches (README: environment switches), read from the environment ONCE PER ENCODER CALL / TRAINING STEP
// (tc_refresh_switches), not per launch -- a process may still change them between calls (the parity tests compare the stem kernels).
// One host thread per GPU drives the library, so a plain global is enough.
struct TcSwitches {
  int debug;          // STRAPS_TC_DEBUG (experiment builds)
  int epi_slab;       // outputs leave through per-warp shared-memory slabs as full runs per pixel (default; STRAPS_TC_EPI=direct: off)
  int conv1;          // STRAPS_TC_CONV1: 2 = "s2dp" pair layout + fused max pool (default), 1 = "s2d" pair layout, stem tensor written,
                      // 0 = "im2col" the round-1 form through conv_tc_kernel
};
static TcSwitches g_sw = {0, 1, 2};

static void tc_refresh_switches() {
  const char* e;
  e = getenv("STRAPS_TC_DEBUG");      g_sw.debug = e ? atoi(e) : 0;
  e = getenv("STRAPS_TC_EPI");        g_sw.epi_slab = (e && e[0] == 'd') ? 0 : 1;      // "direct": per-thread 16-byte stores (round 1)
  e = getenv("STRAPS_TC_CONV1");
  g_sw.conv1 = !e ? 2 : (strcmp(e, "im2col") == 0 ? 0 : strcmp(e, "s2d") == 0 ? 1 : 2);
}

// N tile per layer (see TcCfg)
static int tile_bn(int cout) { return cout == 64 ? 64 : 128; }

static int ensure_encode(TcState* t) {
  if (t->encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  STRAPS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  STRAPS_CHECK(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
  t->encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

// tensor maps of one convolution: A over the split activation planes, W over [Cout][K_eff]
static int build_layer_maps(TcState* t, const TcGeom& c, int B, __half* a_hi, __half* a_lo, void* w_hi, void* w_lo, TcLayerMaps& out) {
  const int bn = tile_bn(c.cout);
  {
    cuuint64_t dims[2] = {(cuuint64_t)c.k_eff, (cuuint64_t)c.cout};
    cuuint64_t str[1] = {(cuuint64_t)c.k_eff * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t es[2] = {1, 1};
    if (encode(t, &out.w_hi, w_hi, 2, dims, str, box, es)) return 1;
    if (encode(t, &out.w_lo, w_lo, 2, dims, str, box, es)) return 1;
  }
  if (c.conv1) {
    // conv1 (im2col form): (flattened kw,c run | ow | oh | b) over the padded input, strides bake in the stride-2 sampling
    cuuint64_t dims[4] = {(cuuint64_t)6 * XP_W * XP_C + C1_KROW, 128, 128, (cuuint64_t)B};
    cuuint64_t str[3] = {2 * XP_C * 2, (cuuint64_t)2 * XP_W * XP_C * 2, (cuuint64_t)XP_H * XP_W * XP_C * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(t, &out.a_hi, a_hi, 4, dims, str, box, es)) return 1;
    if (encode(t, &out.a_lo, a_lo, 4, dims, str, box, es)) return 1;
  } else {
    const int th = (c.hout * c.wout >= BM_TC) ? BM_TC / c.wout : c.hout;
    const int nb = BM_TC / (c.wout * th);
    cuuint64_t dims[4] = {(cuuint64_t)c.cin, (cuuint64_t)c.win, (cuuint64_t)c.hin, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)c.cin * 2, (cuuint64_t)c.win * c.cin * 2, (cuuint64_t)c.hin * c.win * c.cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(c.wout * c.stride), (cuuint32_t)(th * c.stride), (cuuint32_t)nb};
    cuuint32_t es[4] = {1, (cuuint32_t)c.stride, (cuuint32_t)c.stride, 1};
    if (encode(t, &out.a_hi, a_hi, 4, dims, str, box, es)) return 1;
    if (encode(t, &out.a_lo, a_lo, 4, dims, str, box, es)) return 1;
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

template <int BN>
static int launch_conv_tc(const TcLayerMaps& m, const TcConvParams& p, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const int items = p.n_mtiles * p.n_ntiles;
  const int grid = items < num_sms ? items : num_sms;
  conv_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(m.a_hi, m.a_lo, m.w_hi, m.w_lo, p);
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
  p.n_kblocks = c.k_eff / BK_TC;
  p.cchunks = c.cin / BK_TC;
  p.kw_count = c.ksize;
  p.stride = c.stride; p.pad = c.pad;
  p.hw_out = c.hout * c.wout; p.wout = c.wout;
  p.th = (c.hout * c.wout >= BM_TC) ? BM_TC / c.wout : c.hout;
  p.cout = c.cout;
  p.debug = g_sw.debug;
  p.epi_slab = g_sw.epi_slab;
  return bn == 64 ? launch_conv_tc<64>(m, p, t->num_sms, st) : launch_conv_tc<128>(m, p, t->num_sms, st);
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
    t->xs_plane = (size_t)r->max_batch * XS_H * XS_PAIRS * XS_PITCH;
    STRAPS_CUDA(cudaMalloc(&t->xs, 2 * t->xs_plane * sizeof(__half)));
    STRAPS_CUDA(cudaMemsetAsync(t->xs, 0, 2 * t->xs_plane * sizeof(__half), st));      // the 6 border rows stay zero for good
  }
  auto it = t->s2d_maps.find(B);
  if (it == t->s2d_maps.end()) {
    TcState::S2dMaps m;
    const cuuint64_t pitch = (cuuint64_t)XS_PITCH;
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

template <bool POOL>
static int launch_conv1_s2d(straps_regressor* r, const TcState::S2dMaps& m, int B, cudaStream_t st) {
  using Cfg = S2dCfg;
  TcState* t = static_cast<TcState*>(r->tc);
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(conv1_s2d_kernel<POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_once.done(attr_dev);
  }
  const ConvSpec& c = r->conv[0];
  S2dParams p;
  memset(&p, 0, sizeof(p));
  p.n_rows = B * 128;
  p.shift = c.shift;
  p.unscale = t->unscale + t->ch_off[0];
  p.out = act_ptr(r, c.out_buf);
  p.pool_hi = plane_hi(r, r->buf_pool);
  p.pool_lo = plane_lo(r, r->buf_pool);
  p.relu = c.relu;
  p.debug = g_sw.debug;
  const int grid = p.n_rows < t->num_sms ? p.n_rows : t->num_sms;
  conv1_s2d_kernel<POOL><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(m.a_hi, m.a_lo, m.w, p);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// where the stem reads its input from: the caller's fp32 NCHW tensor, or part labels + 2-D joints (the proxy representation is then
// generated inside the pack kernel)
struct StemInput {
  const float* x;            // [B, C, 256, 256] or null
  const float* seg;          // [B, 256, 256] part labels (non-zero = body) when x is null
  const float* joints2d;     // [B, J, 2]
  int num_joints;
  const float* table;        // (2 half_size)^2 Gaussian window of the reference
  int half_size;
};

static int run_conv1_s2d(straps_regressor* r, const StemInput& in, int B, bool pool, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  const TcState::S2dMaps* m = nullptr;
  if (s2d_prepare(r, B, &m, st)) return 1;
  if (in.x) pack_input_s2d_kernel<<<dim3(IMG, B), 256, 0, st>>>(in.x, r->c_in, t->xs, t->xs + t->xs_plane);
  else pack_proxy_s2d_kernel<<<dim3(IMG, B), 256, 0, st>>>(in.seg, in.joints2d, in.num_joints, in.table, in.half_size, t->xs,
                                                         t->xs + t->xs_plane);
  STRAPS_LAUNCH_CHECK();
  return pool ? launch_conv1_s2d<true>(r, *m, B, st) : launch_conv1_s2d<false>(r, *m, B, st);
}

static int tc_encoder_run(straps_regressor* r, const StemInput& in, int B, float* feat, cudaStream_t st);

int tc_encoder_forward(straps_regressor* r, const float* x, int B, float* feat, cudaStream_t st) {
  StemInput in = {x, nullptr, nullptr, 0, nullptr, 0};
  return tc_encoder_run(r, in, B, feat, st);
}

int tc_encoder_forward_from_labels(straps_regressor* r, const float* seg, const float* joints2d, int num_joints, const float* table,
                                   int half_size, int B, float* feat, cudaStream_t st) {
  STRAPS_CHECK(num_joints + 1 == r->c_in, "forward_from_labels: %d joints + the silhouette do not match the regressor's %d input channels",
               num_joints, r->c_in);
  STRAPS_CHECK(half_size > 0 && half_size <= 64, "forward_from_labels: bad Gaussian window half size %d", half_size);
  StemInput in = {nullptr, seg, joints2d, num_joints, table, half_size};
  return tc_encoder_run(r, in, B, feat, st);
}

static int tc_encoder_run(straps_regressor* r, const StemInput& in, int B, float* feat, cudaStream_t st) {
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
  // stem (models/resnet.py:202-206): pixel-pair conv1 with the max pool in its epilogue, unless STRAPS_TC_CONV1 asks for the stem tensor
  // ("s2d") or for the round-1 im2col form ("im2col") -- both bit-identical to the default, kept for the parity tests
  const bool fused_pool = g_sw.conv1 == 2;
  if (g_sw.conv1 != 0 || !in.x) {
    if (run_conv1_s2d(r, in, B, fused_pool, st)) return 1;
  } else {
    pack_input_tc_kernel<<<dim3(IMG, B), 256, 0, st>>>(in.x, r->c_in, t->xp, t->xp + t->xp_plane);
    STRAPS_LAUNCH_CHECK();
    if (run_conv_tc(r, maps, 0, B, st)) return 1;
  }
  t->stem_valid = !fused_pool;
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
  STRAPS_CHECK(buf != r->buf_stem || static_cast<TcState*>(r->tc)->stem_valid,
               "the stem tensor is not materialised when the max pool is fused into conv1: set STRAPS_TC_CONV1=s2d to inspect it");
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
// dY is pre-scaled by a power of two (split_scaled_kernel); the inverse is applied by unpack_dw_all_kernel.
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
            atomicAdd(dst + (size_t)(c0 + j) * p.k_eff, fmaf(__uint_as_float(vl[j]), LO_UNSCALE, __uint_as_float(v[j])));
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
// [cout][k_eff] fp32 accumulators (scaled by the power of two of max |dY|) -> OIHW, every convolution of a backward pass in one launch
struct DwUnpackTable {
  const float* src[NCONV];
  float* dst[NCONV];            // null: skipped
  int cout[NCONV], cin[NCONV], ks[NCONV], k_eff[NCONV];
  int blk_start[NCONV + 1];
  const unsigned* maxbits;      // [NCONV] bits of max |dY| per convolution
};
// one block per output channel: the packed row [k_eff] is read contiguously into shared memory and written out in OIHW order
__global__ void __launch_bounds__(256) unpack_dw_all_kernel(const __grid_constant__ DwUnpackTable t) {
  __shared__ float row[7 * 192 > 512 * 9 ? 7 * 192 : 512 * 9];
  int l = 0;
#pragma unroll 1
  while (l + 1 < NCONV && (int)blockIdx.x >= t.blk_start[l + 1]) ++l;
  const int n = (int)blockIdx.x - t.blk_start[l];
  const int ks = t.ks[l], cin = t.cin[l], k_eff = t.k_eff[l], kk = ks * ks;
  int e = 0;
  const float mx = __uint_as_float(t.maxbits[l]);
  if (mx > 0.f && isfinite(mx)) { frexpf(mx, &e); e = max(-100, min(100, 14 - e)); }
  const float inv = ldexpf(1.f, -e);
  const float* src = t.src[l] + (size_t)n * k_eff;
  for (int k = threadIdx.x; k < k_eff; k += 256) row[k] = src[k];
  __syncthreads();
  float* dst = t.dst[l] + (size_t)n * cin * kk;
  for (int i = threadIdx.x; i < cin * kk; i += 256) {
    const int c = i / kk, tap = i % kk;
    const int k = (l == 0) ? (tap / ks) * C1_KROW + (tap % ks) * XP_C + c : tap * cin + c;
    dst[i] = row[k] * inv;
  }
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
  float* dw_packed;                            // per conv [cout][k_eff] fp32 accumulation targets of the weight gradients (dw_off), zeroed
  size_t dw_elems;                             // once per backward; unpacked to OIHW by ONE launch at its end (tc_train_unpack_all)
  size_t dw_off[NCONV];
  float* dw_dst[NCONV];                        // where this backward's weight gradient of conv i goes (null: not computed on the tensor cores)
  std::map<int, std::vector<TcLayerMaps>> fwd_maps, dg_maps;
  std::map<int, std::vector<TcLayerMaps>> wg_maps;   // a_* = X with 64-pixel boxes, w_* = dY [pixels][cout] boxes {64, 64}
};

static void tc_train_free(TcState* t) {
  if (!t->train) return;
  if (t->train->pool) cudaFree(t->train->pool);
  delete t->train;
  t->train = nullptr;
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

// ---- all 20 convolutions' weights prepared in ONE launch per training step (round 1: w_rowscale + pack_w_tc per layer at the
// forward, w_colscale + pack_w_dgrad_tc per layer at the backward = 78 launches of 3-7 us for 11 M weights) ----
struct WPrepLayer {
  const float* w;               // OIHW fp32 (the parameter itself)
  __half* f_hi; __half* f_lo;   // forward operand [cout][k_eff]
  __half* d_hi; __half* d_lo;   // data-gradient operand [cin][(kh', kw', co)], null for conv1
  int cout, cin, ks, k_eff;
  int ch_off, wd_off;           // offsets into the forward row-scale / data-gradient row-scale arrays
};
struct WPrepTable {
  WPrepLayer L[NCONV];
  int row_start[NCONV + 1];     // blocks [row_start[i], row_start[i+1]) = output channels of conv i (forward operand rows) ...
  int col_start[NCONV + 1];     // ... followed by rows_total + [col_start[i], col_start[i+1]) = input channels of conv i (i >= 1)
  float* wf_scale; float* wf_unscale; float* wd_scale; float* wd_unscale;
};

__device__ __forceinline__ int wprep_find(const int* start, int b) {
  int i = 0;
#pragma unroll 1
  while (i + 1 < NCONV && b >= start[i + 1]) ++i;
  return i;
}
__device__ __forceinline__ void wprep_pow2(float mx, float* pscale, float* unscale) {
  int e = 0;
  if (mx > 0.f && isfinite(mx)) { frexpf(mx, &e); e = max(-60, min(60, 14 - e)); }     // mx * 2^e in [2^13, 2^14)
  *pscale = ldexpf(1.f, e);
  *unscale = ldexpf(1.f, -e);
}

// One block per output channel of a convolution (forward operand row) or per input channel (data-gradient operand row): the block
// brings everything its row is made of into shared memory with contiguous reads (an OIHW row is contiguous; the data-gradient row is
// Cout runs of ks*ks floats), takes the maximum for the power-of-two row scale, and writes the permuted, scaled, split row with
// contiguous stores.  (The first merged version gathered one float per thread from a 36-byte stride: 208 us for 11 M weights.)
constexpr int WPREP_MAX_ROW = 512 * 9;
__global__ void __launch_bounds__(256) pack_w_all_kernel(const __grid_constant__ WPrepTable t) {
  __shared__ float row[WPREP_MAX_ROW];
  __shared__ float red[8];
  __shared__ float s_scale;
  const int rows_total = t.row_start[NCONV];
  int b = blockIdx.x;
  const bool fwd = b < rows_total;
  int i, n;
  if (fwd) { i = wprep_find(t.row_start, b); n = b - t.row_start[i]; }
  else { b -= rows_total; i = wprep_find(t.col_start, b); n = b - t.col_start[i]; }
  const WPrepLayer& L = t.L[i];
  const int kk = L.ks * L.ks;
  const int len = fwd ? L.cin * kk : L.cout * kk;      // elements of this row
  float m = 0.f;
  for (int k = threadIdx.x; k < len; k += 256) {
    // forward: row[c * kk + tap] = w[n][c][tap];  data gradient: row[co * kk + tap] = w[co][n][tap]
    const float v = fwd ? L.w[(size_t)n * len + k] : L.w[((size_t)(k / kk) * L.cin + n) * kk + (k % kk)];
    row[k] = v;
    m = fmaxf(m, fabsf(v));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) mx = fmaxf(mx, red[k]);
    float ps, us;
    wprep_pow2(mx, &ps, &us);
    s_scale = ps;
    if (fwd) { t.wf_scale[L.ch_off + n] = ps; t.wf_unscale[L.ch_off + n] = us; }
    else { t.wd_scale[L.wd_off + n] = ps; t.wd_unscale[L.wd_off + n] = us; }
  }
  __syncthreads();
  const float sc = s_scale;
  if (fwd) {
    __half* hi = L.f_hi + (size_t)n * L.k_eff;
    __half* lo = L.f_lo + (size_t)n * L.k_eff;
    for (int e = threadIdx.x; e < L.k_eff; e += 256) {
      float v = 0.f;
      if (i == 0) {                                    // conv1: K ordered (kh, [kw, c] padded to C1_KROW)
        const int kh = e / C1_KROW, r = e % C1_KROW, kw = r / XP_C, c = r % XP_C;
        if (kw < L.ks && c < L.cin) v = row[c * kk + kh * L.ks + kw];
      } else {                                         // K ordered (tap, c)
        v = row[(e % L.cin) * kk + e / L.cin];
      }
      __half h, l;
      split_f16(v * sc, h, l);
      hi[e] = h;
      lo[e] = l;
    }
  } else {
    const int k_eff = kk * L.cout;                     // K ordered (flipped tap, co)
    __half* hi = L.d_hi + (size_t)n * k_eff;
    __half* lo = L.d_lo + (size_t)n * k_eff;
    for (int e = threadIdx.x; e < k_eff; e += 256) {
      const float v = row[(e % L.cout) * kk + (kk - 1 - e / L.cout)];
      __half h, l;
      split_f16(v * sc, h, l);
      hi[e] = h;
      lo[e] = l;
    }
  }
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
    size_t off_wf[NCONV], off_wd[NCONV], dw_off_tmp[NCONV], nco = 0, nci = 0, dy_el = 0, dw_el = 0;
    for (int i = 0; i < NCONV; ++i) {
      const ConvSpec& c = r->conv[i];
      off_wf[i] = take(2 * (size_t)c.cout * c.k_eff);
      off_wd[i] = (i == 0) ? 0 : take(2 * (size_t)c.cin * c.ksize * c.ksize * c.cout);
      tt->wd_off[i] = nci;
      nco += c.cout; nci += (i == 0) ? 0 : c.cin;
      dy_el = std::max(dy_el, (size_t)r->max_batch * c.hout * c.wout * c.cout);
      if (i > 0) dy_el = std::max(dy_el, (size_t)r->max_batch * c.hin * c.win * c.cout);   // (upsampled) dY of this layer
      dw_off_tmp[i] = dw_el;
      dw_el += (size_t)c.cout * c.k_eff;
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
    for (int i = 0; i < NCONV; ++i) { tt->dw_off[i] = dw_off_tmp[i]; tt->dw_dst[i] = nullptr; }
    std::vector<float> one(512, 1.f);
    STRAPS_CUDA(cudaMemcpy(tt->ones, one.data(), 512 * sizeof(float), cudaMemcpyHostToDevice));
    STRAPS_CUDA(cudaMemset(tt->zeros, 0, 512 * sizeof(float)));
    STRAPS_CUDA(cudaMemset(tt->dy_max, 0, 64 * sizeof(unsigned)));
    t->train = tt;
  }
  TcTrain* tt = t->train;
  // weights of this step (the optimiser has changed them; BN scale not folded): forward AND data-gradient operands of all 20
  // convolutions in one launch
  {
    WPrepTable wt;
    int rows = 0, cols = 0;
    for (int i = 0; i < NCONV; ++i) {
      const ConvSpec& c = r->conv[i];
      WPrepLayer& L = wt.L[i];
      L.w = c.w_oihw; L.f_hi = tt->wf_hi[i]; L.f_lo = tt->wf_lo[i]; L.d_hi = tt->wd_hi[i]; L.d_lo = tt->wd_lo[i];
      L.cout = c.cout; L.cin = c.cin; L.ks = c.ksize; L.k_eff = c.k_eff;
      L.ch_off = (int)t->ch_off[i]; L.wd_off = (i == 0) ? 0 : (int)tt->wd_off[i];
      wt.row_start[i] = rows; rows += c.cout;
      wt.col_start[i] = cols; cols += (i == 0) ? 0 : c.cin;
    }
    wt.row_start[NCONV] = rows; wt.col_start[NCONV] = cols;
    wt.wf_scale = tt->wf_scale; wt.wf_unscale = tt->wf_unscale; wt.wd_scale = tt->wd_scale; wt.wd_unscale = tt->wd_unscale;
    pack_w_all_kernel<<<rows + cols, 256, 0, st>>>(wt);
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
    v = unsplit_f16(hi[s], lo[s]);
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
  // start of a backward pass.  The data-gradient operands were packed with the forward ones (pack_w_all_kernel in tc_train_begin:
  // same weights); what is left is clearing the per-convolution max |dY| words and weight-gradient accumulators -- two memsets.
  TcState* t = static_cast<TcState*>(r->tc);
  STRAPS_CHECK(t && t->train, "tc_train_pack_dgrad: no tensor-core training forward before this backward");
  TcTrain* tt = t->train;
  STRAPS_CUDA(cudaMemsetAsync(tt->dy_max, 0, 64 * sizeof(unsigned), st));
  STRAPS_CUDA(cudaMemsetAsync(tt->dw_packed, 0, tt->dw_elems * sizeof(float), st));
  for (int i = 0; i < NCONV; ++i) tt->dw_dst[i] = nullptr;
  return 0;
}

unsigned* tc_train_dy_max(straps_regressor* r, int ci) {
  TcState* t = static_cast<TcState*>(r->tc);
  return (t && t->train) ? t->train->dy_max + ci : nullptr;
}

// where bn_backward's apply pass writes the scaled split planes of dX (= dY of conv ci), and what its scale kernel needs to publish the
// data gradient's epilogue unscales (row unscales of the packed weights x 2^-e)
void tc_train_dy_planes(straps_regressor* r, int ci, __half** hi, __half** lo) {
  TcTrain* tt = static_cast<TcState*>(r->tc)->train;
  (void)ci;
  *hi = tt->dy_hi; *lo = tt->dy_lo;
}
void tc_train_dgrad_scales(straps_regressor* r, int ci, const float** w_unscale, int* n, float** out_unscale) {
  TcTrain* tt = static_cast<TcState*>(r->tc)->train;
  if (ci > 0) { *w_unscale = tt->wd_unscale + tt->wd_off[ci]; *n = r->conv[ci].cin; *out_unscale = tt->dg_unscale; }
  else { *w_unscale = nullptr; *n = 0; *out_unscale = nullptr; }
}

// end of a backward pass: every weight gradient the tensor cores accumulated, [cout][k_eff] scaled -> OIHW, one launch
int tc_train_unpack_all(straps_regressor* r, cudaStream_t st) {
  TcState* t = static_cast<TcState*>(r->tc);
  TcTrain* tt = t ? t->train : nullptr;
  if (!tt) return 0;
  DwUnpackTable u;
  int blocks = 0;
  for (int i = 0; i < NCONV; ++i) {
    const ConvSpec& c = r->conv[i];
    u.src[i] = tt->dw_packed + tt->dw_off[i];
    u.dst[i] = tt->dw_dst[i];
    u.cout[i] = c.cout; u.cin[i] = c.cin; u.ks[i] = c.ksize; u.k_eff[i] = c.k_eff;
    u.blk_start[i] = blocks;
    if (tt->dw_dst[i]) blocks += c.cout;
    tt->dw_dst[i] = nullptr;
  }
  u.blk_start[NCONV] = blocks;
  u.maxbits = tt->dy_max;
  if (!blocks) return 0;
  unpack_dw_all_kernel<<<blocks, 256, 0, st>>>(u);
  STRAPS_LAUNCH_CHECK();
  return 0;
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
  split_scaled_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(dy, tt->dy_max + ci, n4, c.cout, c.hout, c.wout, up, tt->dy_hi, tt->dy_lo,
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
  p.dw = tt->dw_packed + tt->dw_off[ci];       // zeroed at the start of the backward pass (tc_train_pack_dgrad)
  const TcLayerMaps& m = tt->wg_maps.at(B)[ci];
  if (bn == 64 ? launch_wgrad_tc<64>(m, p, t->num_sms, st) : launch_wgrad_tc<128>(m, p, t->num_sms, st)) return 1;
  tt->dw_dst[ci] = dw_oihw;                    // unpacked by tc_train_unpack_all at the end of the pass
  return 0;
}

}  // namespace straps

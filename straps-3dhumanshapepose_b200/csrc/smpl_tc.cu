// smpl_tc.cu -- SMPL linear blend skinning with the blend shapes on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces, for batches >= 32, the CUDA-core lbs_kernel of smpl.cu on the path smplx.lbs.lbs (called from the
// reference's models/smpl_official.py:29).  lbs_kernel spends 621 FMA per (vertex, body) on the pose-corrective term
// pf[B,207] @ posedirs[207,20670] and is bound by instruction issue (2.35 warp instructions per FMA; round-1 review, weak #4):
// 0.03 of the HBM roofline at B = 4096.  That term -- and the shape blend and the template with it -- is a GEMM:
//
//   v_posed[b, (v,c)] = sum_k  Apk[(v,c), k] * Bm[b, k]      k = 0..206 pose feature (R_j - I), 207..216 betas, 217 the constant 1
//                                                            Apk = [posedirs ; shapedirs ; v_template], K padded to 256
//
//   T[b, v, (r,c4)]   = sum_j  W[v, j] * A[b, j, (r,c4)]     W = skinning weights, A = the 24 3x4 transforms of the kinematic chain
//   out[b, v, r]      = T[b,v,(r,0..2)] . v_posed[b,v,:] + T[b,v,(r,3)] + transl[b,r]
//
// Kernel 1 (smpl_chain_kernel, 4 bodies per CTA of 256 threads): rotations (Rodrigues or given), pre-reduced rest joints, the 24-joint
//   kinematic chain (the code of lbs_kernel's prologue, same arithmetic order) -> the 24 posed joints and the transforms A; writes BOTH
//   per-body tensor-core operands as ready-made swizzled shared-memory images: Bm (fp16 hi / lo, x 2^10, SWIZZLE_128B) and the
//   transposed transforms A^T[(body, column)][joint] (fp16 hi / lo, x 2^10, SWIZZLE_64B rows of 32 joint slots).
// Kernel 2, two GEMMs into TMEM (M = 128 vertices, 3-pass fp16 split A_hi.B_hi + A_hi.B_lo + A_lo.B_hi: 22 mantissa bits; the rows of Apk
//   and W are pre-scaled by a power of two so that both halves are fp16-normal), constants packed ONCE at create time as swizzled
//   images so that every operand arrives by a plain bulk-async copy (no tensor maps):
//   * lbs_tc3_kernel (default): persistent, software pipelined -- see its own comment below;
//   * lbs_tc_kernel (STRAPS_LBS_V=2): one CTA per (vertex tile, body group), blend GEMM through a 5-stage ring, then the transform
//     GEMM one output row at a time (N = 256), epilogue = T_r . v_posed from TMEM, output staged in shared memory.  Kept because it
//     is the simple form of the same arithmetic (bit-identical results) and the base line of profiles/r02_lbs_tc_ncu.txt.
//   The first version blended the transforms on the CUDA cores from shared memory and looked LDS-bound (B = 4096: 563 us); the real
//   critical path was a per-body __ldg of transl inside the row loop (see lbs_tc_kernel's epilogue).
// Roofline: HBM by contract (20,587,320 B + 84,664 B per body per launch); tensor work 3 x 2 x (224 x 20736 + 32 x 128 x 54 x 12) FLOP / body.
#include "smpl.h"
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace straps {

constexpr int TC_K = 256;                                   // padded K of the blend GEMM
constexpr int TC_KREAL = NPF + STRAPS_NUM_BETAS + 1;        // 218
constexpr int TC_KSTEPS = (TC_KREAL + 15) / 16;             // 14
constexpr int TC_CHUNKS = 4;                                // K chunks of 64 elements (one SWIZZLE_128B row each)
constexpr int TC_NB = 64;                                   // bodies per CTA = MMA N of the blend GEMM
constexpr int TC_A_IMG = TV * 128;                          // bytes of one [128 vertices x 64 K] fp16 image
constexpr int TC_B_IMG = TC_NB * 128;                       // bytes of one [64 bodies x 64 K] fp16 image
constexpr int TC_STAGES_PER_TILE = TC_CHUNKS * 3 * 2;       // (chunk, plane, hi | lo) = 24 A images per vertex tile
constexpr int TC_NA = 5;                                    // A ring stages
constexpr float TC_BSCALE = 1024.f;                         // Bm and the transforms are multiplied by 2^10 before their fp16 split
constexpr float TC_WSCALE = 16384.f;                        // skinning weights (<= 1) by 2^14
constexpr int TC_B_BYTES = TC_CHUNKS * 2 * TC_B_IMG;        // Bm images of one body group: 65,536 B
// skinning-transform GEMM: K = 32 joint slots (24 used) = one SWIZZLE_64B row of 64 bytes
constexpr int TC_TK = 32;
constexpr int TC_TN = TC_NB * 4;                            // 256 = (body, column of the 3x4 transform)
constexpr int TC_W_IMG = TV * 64;                           // [128 vertices x 32 joints] fp16: 8 KB
constexpr int TC_W_BYTES = 2 * TC_W_IMG;                    // hi + lo
constexpr int TC_AT_IMG = TC_TN * 64;                       // [256 x 32 joints] fp16: 16 KB
constexpr int TC_AT_BYTES = 2 * TC_AT_IMG;                  // hi + lo of one output row r
constexpr int TC_AT_GROUP = 3 * TC_AT_BYTES;                // 98,304 B per body group in HBM
constexpr int TC_OFF_B = TC_NA * TC_A_IMG;
constexpr int TC_OFF_W = TC_OFF_B + TC_B_BYTES;
constexpr int TC_OFF_AT = TC_OFF_W + TC_W_BYTES;            // two slots: rows 0 and 1 up front, row 2 re-uses slot 0
constexpr int TC_OFF_BAR = TC_OFF_AT + 2 * TC_AT_BYTES;
constexpr int TC_OFF_TR = TC_OFF_BAR + 256;                 // transl of the group's 64 bodies
constexpr int TC_SMEM = TC_OFF_TR + TC_NB * 3 * 4 + 1024;
constexpr int TC_THREADS = 320;                             // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int TC_EPI_THREADS = 256;
static_assert(TC_NB * TV * 3 * 4 <= TC_OFF_W, "output staging must fit below the transform operands");
constexpr int TC_COL_T = 3 * TC_NB;                         // TMEM: columns [0, 192) v_posed planes, [192, 448) T_r
static_assert(TC_SMEM <= 232448, "lbs_tc_kernel shared memory");
static_assert(TC_KREAL <= TC_K && TC_KSTEPS <= TC_CHUNKS * 4, "K padding");
static_assert(NJ <= TC_TK && TC_COL_T + TC_TN <= 512, "transform GEMM shape");

struct LbsTcArgs {
  const unsigned char* apk;     // [54][24] images of TC_A_IMG bytes
  const float* ainv;            // [54][3][128] 1 / (row scale * TC_BSCALE)
  const unsigned char* wk;      // [54][2] images of TC_W_IMG bytes
  const unsigned char* bimg;    // [groups][4][2] images of TC_B_IMG bytes
  const unsigned char* atimg;   // [groups][3][2] images of TC_AT_IMG bytes
  const float* transl;          // [B][3] or null
  int B;
  float* verts;
  float* save_vposed;           // or null
};

__device__ __forceinline__ void tc_rodrigues(const float r[3], float* R) {
  // smplx.lbs.batch_rodrigues: the epsilon goes inside the norm; same code as smpl.cu
  float x = r[0] + 1e-8f, y = r[1] + 1e-8f, z = r[2] + 1e-8f;
  float angle = sqrtf(x * x + y * y + z * z);
  float ax = r[0] / angle, ay = r[1] / angle, az = r[2] / angle;
  float c = cosf(angle), s = sinf(angle);
  float K[9] = {0.f, -az, ay, az, 0.f, -ax, -ay, ax, 0.f};
  float omc = 1.f - c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float kk = K[i * 3 + 0] * K[0 * 3 + j] + K[i * 3 + 1] * K[1 * 3 + j] + K[i * 3 + 2] * K[2 * 3 + j];
      R[i * 3 + j] = (i == j ? 1.f : 0.f) + s * K[i * 3 + j] + omc * kk;
    }
}

__device__ __forceinline__ void tc_split_store(unsigned char* hi_img, size_t lo_distance, size_t off, float v) {
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  *reinterpret_cast<__half*>(hi_img + off) = h;
  *reinterpret_cast<__half*>(hi_img + off + lo_distance) = l;
}

struct ChainArgs {
  const float* go;
  const float* bp;
  const float* betas;
  const float* transl;
  long long go_stride, bp_stride, betas_stride;
  int B, B_pad, pose2rot;
  float* joints;                // [B][90][3]: rows 0..23 written here
  unsigned char* bimg;
  unsigned char* atimg;
  int at_layout;                // 0: [r][hi|lo][256 rows] (lbs_tc_kernel)   1: [half][r][hi|lo][128 rows] (lbs_tc3_kernel)
  float* save_A;                // training: the transforms as fp32 [B][24][12] in the caller's tensor, or null
};

// 4 bodies per CTA of 256 threads: the kernel is a chain of short dependent phases (one CTA of 128 threads x 8 bodies took 18 us at any
// batch size, instruction-latency bound), so the work per thread is what counts
constexpr int CH_TB = 4;
constexpr int CH_THREADS = 256;

__global__ void __launch_bounds__(CH_THREADS) smpl_chain_kernel(const SmplDev m, const ChainArgs a) {
  __shared__ float sR[CH_TB][NJ][9];
  __shared__ float sJ[CH_TB][NJ][3];
  __shared__ float sG[CH_TB][NJ][12];
  __shared__ float sbeta[CH_TB][STRAPS_NUM_BETAS];
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * CH_TB;
  const int nb = max(0, min(CH_TB, a.B - b0));
  for (int i = tid; i < CH_TB * STRAPS_NUM_BETAS; i += CH_THREADS) {
    int b = i / STRAPS_NUM_BETAS, l = i % STRAPS_NUM_BETAS;
    sbeta[b][l] = (b < nb) ? a.betas[(size_t)(b0 + b) * a.betas_stride + l] : 0.f;
  }
  if (a.pose2rot) {
    for (int i = tid; i < CH_TB * NJ; i += CH_THREADS) {
      int b = i / NJ, j = i % NJ;
      float r[3] = {0.f, 0.f, 0.f};
      if (b < nb) {
        const float* src = (j == 0) ? a.go + (size_t)(b0 + b) * a.go_stride : a.bp + (size_t)(b0 + b) * a.bp_stride + (j - 1) * 3;
        r[0] = src[0]; r[1] = src[1]; r[2] = src[2];
      }
      tc_rodrigues(r, &sR[b][j][0]);
    }
  } else {
    // fully unrolled so that all of a thread's loads are in flight together: as a rolled loop every iteration waited for its own
    // L2 / HBM round trip (14 of them in a row -- most of this kernel's 20 us at any batch size)
    constexpr int NIT = (CH_TB * NJ * 9 + CH_THREADS - 1) / CH_THREADS;
    float val[NIT];
#pragma unroll
    for (int u = 0; u < NIT; ++u) {
      const int i = tid + u * CH_THREADS;
      const int b = i / (NJ * 9), r = i % (NJ * 9), j = r / 9, e = r % 9;
      float v = (e == 0 || e == 4 || e == 8) ? 1.f : 0.f;
      if (i < CH_TB * NJ * 9 && b < nb)
        v = (j == 0) ? a.go[(size_t)(b0 + b) * a.go_stride + e] : a.bp[(size_t)(b0 + b) * a.bp_stride + (j - 1) * 9 + e];
      val[u] = v;
    }
#pragma unroll
    for (int u = 0; u < NIT; ++u) {
      const int i = tid + u * CH_THREADS;
      if (i < CH_TB * NJ * 9) (&sR[0][0][0])[i] = val[u];
    }
  }
  // rest-joint regressors of this thread's outputs, fetched before the barrier for the same reason
  constexpr int NJT = (CH_TB * NJ * 3 + CH_THREADS - 1) / CH_THREADS;
  float jsv[NJT][STRAPS_NUM_BETAS], jtv[NJT];
#pragma unroll
  for (int u = 0; u < NJT; ++u) {
    const int i = tid + u * CH_THREADS, jc = (i % (NJ * 3));
#pragma unroll
    for (int l = 0; l < STRAPS_NUM_BETAS; ++l) jsv[u][l] = m.js[jc * STRAPS_NUM_BETAS + l];
    jtv[u] = m.jt[jc];
  }
  __syncthreads();
  // Bm = [R_j - I (j = 1..23) | betas | 1 | 0...] * 2^10, fp16 hi / lo, in the SWIZZLE_128B image of its body group
  for (int i = tid; i < CH_TB * TC_K; i += CH_THREADS) {
    const int b = i / TC_K, k = i % TC_K;
    const int gb = b0 + b;
    if (gb >= a.B_pad) continue;
    float v = 0.f;
    if (b < nb) {
      if (k < NPF) { const int e = k % 9; v = sR[b][1 + k / 9][e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f); }
      else if (k < NPF + STRAPS_NUM_BETAS) v = sbeta[b][k - NPF];
      else if (k == NPF + STRAPS_NUM_BETAS) v = 1.f;
    }
    const int g = gb / TC_NB, r = gb % TC_NB, kc = k >> 6, u = (k & 63) >> 3, e = k & 7;
    const size_t off = ((size_t)(g * TC_CHUNKS + kc) * 2) * TC_B_IMG + (size_t)r * 128 + (size_t)((u ^ (r & 7)) << 4) + (size_t)e * 2;
    tc_split_store(a.bimg, TC_B_IMG, off, v * TC_BSCALE);
  }
#pragma unroll
  for (int u = 0; u < NJT; ++u) {
    const int i = tid + u * CH_THREADS;
    if (i < CH_TB * NJ * 3) {
      const int b = i / (NJ * 3), jc = i % (NJ * 3);
      float acc = 0.f;
#pragma unroll
      for (int l = 0; l < STRAPS_NUM_BETAS; ++l) acc = fmaf(jsv[u][l], sbeta[b][l], acc);
      sJ[b][jc / 3][jc % 3] = jtv[u] + acc;
    }
  }
  __syncthreads();
  for (int lvl = 0; lvl < m.nlevels; ++lvl) {
    const int j0 = m.lvl_start[lvl], nj = m.lvl_start[lvl + 1] - j0;
    for (int i = tid; i < CH_TB * nj * 3; i += CH_THREADS) {
      int b = i / (nj * 3), q = i % (nj * 3), j = m.lvl_joint[j0 + q / 3], r = q % 3;
      int p = m.parents[j];
      const float* Rj = &sR[b][j][0];
      float* g = &sG[b][j][r * 4];
      if (p < 0) {
        g[0] = Rj[r * 3 + 0]; g[1] = Rj[r * 3 + 1]; g[2] = Rj[r * 3 + 2]; g[3] = sJ[b][j][r];
      } else {
        const float* gp = &sG[b][p][r * 4];
        float rel0 = sJ[b][j][0] - sJ[b][p][0];
        float rel1 = sJ[b][j][1] - sJ[b][p][1];
        float rel2 = sJ[b][j][2] - sJ[b][p][2];
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = gp[0] * Rj[0 * 3 + c] + gp[1] * Rj[1 * 3 + c] + gp[2] * Rj[2 * 3 + c];
        g[3] = gp[0] * rel0 + gp[1] * rel1 + gp[2] * rel2 + gp[3];
      }
    }
    __syncthreads();
  }
  // skinning transforms A = [G_rot | G_t - G_rot J_rest]: fp32 for the training path, and as the B operand of the transform GEMM --
  // image (group, r, hi | lo) = 256 rows n = (body in group) * 4 + column, 64 bytes each (32 joint slots), SWIZZLE_64B: the 16-byte
  // unit u of row n sits at physical unit u ^ ((n >> 1) & 3).  Joint slots 24..31 stay zero from the memset at allocation.
  for (int i = tid; i < CH_TB * NJ * 3; i += CH_THREADS) {
    int b = i / (NJ * 3), q = i % (NJ * 3), j = q / 3, r = q % 3;
    const int gb = b0 + b;
    if (gb >= a.B_pad) continue;
    const float* g = &sG[b][j][r * 4];
    const float t = g[3] - (g[0] * sJ[b][j][0] + g[1] * sJ[b][j][1] + g[2] * sJ[b][j][2]);
    const bool live = b < nb;
    const float row[4] = {live ? g[0] : 0.f, live ? g[1] : 0.f, live ? g[2] : 0.f, live ? t : 0.f};
    const int grp = gb / TC_NB, bl = gb % TC_NB;
    // layout 0: one image pair per output row r, 256 rows; layout 1: one pair per (half of the group, r), 128 rows
    unsigned char* img = a.atimg + (size_t)grp * TC_AT_GROUP +
                         (a.at_layout ? (size_t)((bl >> 5) * 3 + r) * (TC_AT_BYTES / 2) : (size_t)r * TC_AT_BYTES);
    const size_t lo_dist = a.at_layout ? TC_AT_IMG / 2 : TC_AT_IMG;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const int n = (a.at_layout ? (bl & 31) : bl) * 4 + c4;
      const size_t off = (size_t)n * 64 + (size_t)(((j >> 3) ^ ((n >> 1) & 3)) << 4) + (size_t)(j & 7) * 2;
      tc_split_store(img, lo_dist, off, row[c4] * TC_BSCALE);
    }
    if (live) {
      if (a.save_A) *reinterpret_cast<float4*>(a.save_A + ((size_t)gb * NJ + j) * 12 + r * 4) = make_float4(row[0], row[1], row[2], row[3]);
      const float tr = a.transl ? a.transl[(size_t)gb * 3 + r] : 0.f;
      a.joints[((size_t)gb * STRAPS_NUM_SUPERSET_JOINTS + j) * 3 + r] = g[3] + tr;
    }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) lbs_tc_kernel(const LbsTcArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);
  uint64_t* a_full = bars;                 // [TC_NA]
  uint64_t* a_empty = bars + TC_NA;        // [TC_NA]
  uint64_t* b_full = bars + 2 * TC_NA;     // Bm + W
  uint64_t* at_full = b_full + 1;          // [2] transform operand slots
  uint64_t* at_free = at_full + 2;         // slot 0 may be refilled (row 0's MMAs have completed)
  uint64_t* vp_full = at_free + 1;         // blend accumulators complete
  uint64_t* t_full = vp_full + 1;          // T_r complete
  uint64_t* t_empty = t_full + 1;          // every epilogue thread has read T_r
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, group = blockIdx.y;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < TC_NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(b_full, 1); mbar_init(&at_full[0], 1); mbar_init(&at_full[1], 1); mbar_init(at_free, 1);
    mbar_init(vp_full, 1); mbar_init(t_full, 1); mbar_init(t_empty, TC_EPI_THREADS);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem0 = smem_u32(smem);

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= producer: group operands up front, then the 24 A images of this vertex tile, then transform row 2 =================
      mbar_arrive_expect_tx(b_full, TC_B_BYTES + TC_W_BYTES);
      bulk_g2s(smem + TC_OFF_B, a.bimg + (size_t)group * TC_B_BYTES, TC_B_BYTES, b_full);
      bulk_g2s(smem + TC_OFF_W, a.wk + (size_t)tile * TC_W_BYTES, TC_W_BYTES, b_full);
      const unsigned char* at = a.atimg + (size_t)group * TC_AT_GROUP;
      for (int r = 0; r < 2; ++r) {
        mbar_arrive_expect_tx(&at_full[r], TC_AT_BYTES);
        bulk_g2s(smem + TC_OFF_AT + r * TC_AT_BYTES, at + (size_t)r * TC_AT_BYTES, TC_AT_BYTES, &at_full[r]);
      }
      const unsigned char* src = a.apk + (size_t)tile * TC_STAGES_PER_TILE * TC_A_IMG;
      // the whole 384 KB of this tile towards L2 at once: after the encoder's activations have gone through L2 the constants are in
      // HBM again, and five 16 KB stages in flight would fetch them in ~5 dependent round trips
      for (int s = 0; s < TC_STAGES_PER_TILE; s += 4) bulk_prefetch_l2(src + (size_t)s * TC_A_IMG, 4 * TC_A_IMG);
      uint32_t st = 0, ph = 1;
      for (int s = 0; s < TC_STAGES_PER_TILE; ++s) {
        mbar_wait(&a_empty[st], ph);
        mbar_arrive_expect_tx(&a_full[st], TC_A_IMG);
        bulk_g2s(smem + st * TC_A_IMG, src + (size_t)s * TC_A_IMG, TC_A_IMG, &a_full[st]);
        if (++st == TC_NA) { st = 0; ph ^= 1; }
      }
      mbar_wait(at_free, 0);
      mbar_arrive_expect_tx(&at_full[0], TC_AT_BYTES);
      bulk_g2s(smem + TC_OFF_AT, at + (size_t)2 * TC_AT_BYTES, TC_AT_BYTES, &at_full[0]);
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer =================
      // (i) blend: per K chunk and plane, A_hi.[B_hi, B_lo] then A_lo.B_hi into the plane's accumulator
      constexpr uint32_t idesc = umma_idesc_f16(TV, TC_NB);
      const uint32_t adesc0 = umma_desc_sw128_lo(smem0), bdesc0 = umma_desc_sw128_lo(smem0 + TC_OFF_B);
      mbar_wait(b_full, 0);
      tc_fence_after();
      uint32_t st = 0, ph = 0;
      for (int kc = 0; kc < TC_CHUNKS; ++kc) {
        const int nks = min(4, TC_KSTEPS - 4 * kc);
        const uint32_t b_hi = bdesc0 + (uint32_t)(kc * 2) * (TC_B_IMG >> 4), b_lo = b_hi + (TC_B_IMG >> 4);
        for (int c = 0; c < 3; ++c) {
          const uint32_t d = tmem_base + c * TC_NB;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&a_full[st], ph);
            tc_fence_after();
            const uint32_t ad = adesc0 + st * (TC_A_IMG >> 4);
            for (int ks = 0; ks < nks; ++ks) {
              const uint32_t ko = ks * 2;                                   // +32 bytes along K inside the swizzle row
              if (h == 0) {
                umma_f16_lohi(d, ad + ko, b_hi + ko, idesc, (kc | ks) != 0);
                umma_f16_lohi(d, ad + ko, b_lo + ko, idesc, 1);
              } else {
                umma_f16_lohi(d, ad + ko, b_hi + ko, idesc, 1);
              }
            }
            umma_commit(&a_empty[st]);
            if (++st == TC_NA) { st = 0; ph ^= 1; }
          }
        }
      }
      umma_commit(vp_full);
      // (ii) transforms, one output row at a time: W_hi.AT_hi + W_hi.AT_lo + W_lo.AT_hi, K = 2 steps of 16 joints (SWIZZLE_64B rows)
      constexpr uint32_t idesc_t = umma_idesc_f16(TV, TC_TN);
      const uint32_t w_hi = umma_desc_sw128_lo(smem0 + TC_OFF_W), w_lo = w_hi + (TC_W_IMG >> 4);
      const uint32_t dt = tmem_base + TC_COL_T;
      for (int r = 0; r < 3; ++r) {
        const int slot = r & 1;
        if (r > 0) { mbar_wait(t_empty, (r - 1) & 1); }
        mbar_wait(&at_full[slot], r >> 1);
        tc_fence_after();
        const uint32_t t_hi = umma_desc_sw128_lo(smem0 + TC_OFF_AT + slot * TC_AT_BYTES), t_lo = t_hi + (TC_AT_IMG >> 4);
#pragma unroll
        for (int ks = 0; ks < TC_TK / 16; ++ks) {
          const uint32_t ko = ks * 2;
          umma_f16_lohi(dt, w_hi + ko, t_hi + ko, idesc_t, ks != 0, UMMA_DESC_SW64_HI);
          umma_f16_lohi(dt, w_hi + ko, t_lo + ko, idesc_t, 1, UMMA_DESC_SW64_HI);
          umma_f16_lohi(dt, w_lo + ko, t_hi + ko, idesc_t, 1, UMMA_DESC_SW64_HI);
        }
        umma_commit(t_full);
        if (r == 0) umma_commit(at_free);
      }
    }
  }
  __syncwarp();

  // ================= epilogue, warps 2..9: thread = vertex row (TMEM lane) x 32 bodies =================
  // (the two issue warps stay out of it: the MMA issuer waits for the epilogue's t_empty arrivals between output rows)
  if (warp >= 2) {
  const int quad = warp & 3, half = (warp - 2) >> 2;    // a warp reads the TMEM lanes 32 (warp % 4) .. +31
  const int row = quad * 32 + lane;
  const int v = tile * TV + row;                        // < VPAD: the packed arrays are zero padded
  const bool vok = v < V;
  const int body0 = group * TC_NB + half * 32;          // first of this thread's 32 bodies
  const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
  // transl of the group's bodies once, into shared memory: a per-body __ldg inside the row loop (the first version) was a dependent
  // L2 round trip before every output -- 96 of them per thread, 25 k of the CTA's 49 k cycles (profiles/r02_lbs_tc_ncu.txt)
  float* s_tr = reinterpret_cast<float*>(smem + TC_OFF_TR);
  for (int i = threadIdx.x - 64; i < TC_NB * 3; i += TC_EPI_THREADS) {
    const int gb = group * TC_NB + i / 3;
    s_tr[i] = (a.transl && gb < a.B) ? a.transl[(size_t)gb * 3 + i % 3] : 0.f;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
  mbar_wait(vp_full, 0);
  tc_fence_after();
  float ainv[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) ainv[c] = a.ainv[(tile * 3 + c) * TV + row];
  if (a.save_vposed) {                                  // training path: v_posed itself, once
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t d[3][8];
#pragma unroll
      for (int c = 0; c < 3; ++c) tmem_ld_32x8(tlane + c * TC_NB + half * 32 + q * 8, d[c]);
      tmem_ld_wait();
#pragma unroll
      for (int bb = 0; bb < 8; ++bb) {
        const int gb = body0 + q * 8 + bb;
        if (gb < a.B && vok) {
          float* sv = a.save_vposed + ((size_t)gb * V + v) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) sv[c] = __uint_as_float(d[c][bb]) * ainv[c];
        }
      }
    }
  }
  // out_r = T_r . [v_posed ; 1]: the v_posed accumulators are re-read from TMEM for every row (24 registers instead of 96 held
  // across the loop; 10 warps leave 168 registers per thread), with 1 / (all scales) folded into three per-thread constants
  // output staging [64 bodies][128 vertices x 3] fp32 = 96 KB over the A ring and Bm: both are dead once vp_full has completed
  float* sout = reinterpret_cast<float*>(smem);
  constexpr float tscale = 1.f / (TC_WSCALE * TC_BSCALE);
  const float k0 = ainv[0] * tscale, k1 = ainv[1] * tscale, k2 = ainv[2] * tscale;
#pragma unroll 1
  for (int r = 0; r < 3; ++r) {
    mbar_wait(t_full, r & 1);
    tc_fence_after();
#pragma unroll
    for (int q = 0; q < 4; ++q) {                       // 8 bodies = 32 TMEM columns of T_r per load
      uint32_t t[32], d[3][8];
      tmem_ld_32x32(tlane + TC_COL_T + half * 128 + q * 32, t);
#pragma unroll
      for (int c = 0; c < 3; ++c) tmem_ld_32x8(tlane + c * TC_NB + half * 32 + q * 8, d[c]);
      tmem_ld_wait();
#pragma unroll
      for (int bb = 0; bb < 8; ++bb) {
        const int gb = body0 + q * 8 + bb;
        if (gb < a.B && vok) {
          const float tr = s_tr[(q * 8 + bb + half * 32) * 3 + r];
          float acc = fmaf(__uint_as_float(t[bb * 4 + 3]), tscale, tr);
          acc = fmaf(__uint_as_float(t[bb * 4 + 0]) * k0, __uint_as_float(d[0][bb]), acc);
          acc = fmaf(__uint_as_float(t[bb * 4 + 1]) * k1, __uint_as_float(d[1][bb]), acc);
          acc = fmaf(__uint_as_float(t[bb * 4 + 2]) * k2, __uint_as_float(d[2][bb]), acc);
          sout[(q * 8 + bb + half * 32) * (TV * 3) + row * 3 + r] = acc;        // lanes 3 words apart: conflict free
        }
      }
    }
    if (r < 2) {
      tc_fence_before();
      mbar_arrive(t_empty);
    }
  }
  // the tile's vertices of one body are 1536 contiguous bytes of the output: written from the staging buffer with full-width
  // stores.  (The first version stored each coordinate from its thread -- 4 bytes every 12 -- i.e. 12 partially written sectors
  // per warp instruction, 144 sector writes per body and tile instead of 48: B = 4096 took 582 us.)
  asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
  {
    const int n2 = (min(TV, V - tile * TV) * 3) / 2;            // float2 per body (384 or 318 floats: both even)
    for (int b = warp - 2; b < TC_NB; b += TC_EPI_THREADS / 32) {
      const int gb = group * TC_NB + b;
      if (gb >= a.B) break;
      const float2* src = reinterpret_cast<const float2*>(sout + b * (TV * 3));
      float2* dst = reinterpret_cast<float2*>(a.verts + ((size_t)gb * V + (size_t)tile * TV) * 3);      // 8-byte aligned (82,680 B per body)
      for (int i = lane; i < n2; i += 32) dst[i] = src[i];
    }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// lbs_tc3_kernel: the same two GEMMs as lbs_tc_kernel, persistent and software pipelined.
//
// What ncu showed for lbs_tc_kernel (one CTA per work item, profiles/r02_lbs_tc_ncu.txt): the tensor pipe was busy 23 % of the time;
// an item (128 vertices x 64 bodies) took 16 us of which 8.6 us was the blend phase WAITING FOR ITS OPERANDS -- 464 KB through 80 KB
// of ring = one L2 round trip (~1.5 us under load) per 80 KB, 54 GB/s per SM -- and the stream stood still during the transform
// phases, the epilogue, the copy-out and the next CTA's start-up.  Here:
//   * a CTA walks a contiguous range of items; the producer runs ahead across item boundaries, so the stream never stops;
//   * every operand is streamed (Bm per K chunk, the transform operand per phase): the A ring gets 112 KB instead of 80;
//   * v_posed accumulators are double buffered in TMEM (2 x 192 columns + 128 for T = 512), and because tcgen05.mma executes in
//     issue order the issuer INTERLEAVES: six chunks of the next item's blend GEMM alternate with the six transform phases
//     (half of the group x output row) of the current item, so the epilogue of phase p runs under blend chunk p + 1.
// Producer and issuer follow the same operand order, so a full ring can only wait on work that has already been issued.
constexpr int T3_NA = 7;                                    // A ring slots of 16 KB
constexpr int T3_BM_SLOT = 2 * TC_B_IMG;                    // Bm of one K chunk, hi + lo: 16 KB
constexpr int T3_AT_SLOT = TC_AT_BYTES / 2;                 // transform operand of one phase (32 bodies x 4 columns, hi + lo): 16 KB
constexpr int T3_OFF_BM = T3_NA * TC_A_IMG;                 // 2 slots
constexpr int T3_OFF_W = T3_OFF_BM + 2 * T3_BM_SLOT;        // 2 slots
constexpr int T3_OFF_AT = T3_OFF_W + 2 * TC_W_BYTES;        // 2 slots
constexpr int T3_OFF_BAR = T3_OFF_AT + 2 * T3_AT_SLOT;
constexpr int T3_OFF_TR = T3_OFF_BAR + 512;
constexpr int T3_OFF_ST = T3_OFF_TR + TC_NB * 3 * 4;         // per epilogue warp: two slabs of 32 vertices x 3 floats (output transposition)
constexpr int T3_EPI_WARPS = 16;                            // 4 TMEM lane quadrants x 4 body octets of a phase
constexpr int T3_EPI_THREADS = T3_EPI_WARPS * 32;
constexpr int T3_THREADS = 64 + T3_EPI_THREADS;             // warp 0 producer, warp 1 MMA issuer, warps 2..17 epilogue
constexpr int T3_SMEM = T3_OFF_ST + T3_EPI_WARPS * 2 * 384 + 1024;
constexpr int T3_TN = 128;                                  // transform GEMM N per phase
constexpr int T3_COL_T = 2 * 3 * TC_NB;                     // TMEM: [0,192) v_posed buffer 0, [192,384) buffer 1, [384,512) T
static_assert(T3_SMEM <= 232448 && T3_COL_T + T3_TN <= 512, "lbs_tc3_kernel resources");

struct LbsTc3Args {
  LbsTcArgs k;
  int n_items;                  // groups * NTILES, item = group * NTILES + tile
};

__global__ void __launch_bounds__(T3_THREADS, 1) lbs_tc3_kernel(const LbsTc3Args args) {
  const LbsTcArgs& a = args.k;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T3_OFF_BAR);
  uint64_t* a_full = bars;                    // [T3_NA]
  uint64_t* a_empty = a_full + T3_NA;         // [T3_NA]
  uint64_t* bm_full = a_empty + T3_NA;        // [2]
  uint64_t* bm_empty = bm_full + 2;           // [2]
  uint64_t* wk_full = bm_empty + 2;           // [2]
  uint64_t* wk_empty = wk_full + 2;           // [2]
  uint64_t* at_full = wk_empty + 2;           // [2]
  uint64_t* at_empty = at_full + 2;           // [2]
  uint64_t* vp_full = at_empty + 2;           // [2]
  uint64_t* vp_empty = vp_full + 2;           // [2]
  uint64_t* t_full = vp_empty + 2;
  uint64_t* t_empty = t_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i_beg = (int)((long long)blockIdx.x * args.n_items / gridDim.x);
  const int i_end = (int)((long long)(blockIdx.x + 1) * args.n_items / gridDim.x);

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < T3_NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bm_full[i], 1); mbar_init(&bm_empty[i], 1); mbar_init(&wk_full[i], 1); mbar_init(&wk_empty[i], 1);
      mbar_init(&at_full[i], 1); mbar_init(&at_empty[i], 1); mbar_init(&vp_full[i], 1); mbar_init(&vp_empty[i], T3_EPI_THREADS);
    }
    mbar_init(t_full, 1); mbar_init(t_empty, T3_EPI_THREADS);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem0 = smem_u32(smem);

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= producer =================
      uint32_t na = 0, nbm = 0, nwk = 0, nat = 0;            // loads issued so far into each ring
      auto blend_stage = [&](int item, int s) {
        const int group = item / NTILES, tile = item % NTILES;
        if (s % 6 == 0) {                                    // Bm of K chunk s / 6 goes first
          const uint32_t sl = nbm & 1;
          mbar_wait(&bm_empty[sl], ((nbm >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&bm_full[sl], T3_BM_SLOT);
          bulk_g2s(smem + T3_OFF_BM + sl * T3_BM_SLOT, a.bimg + ((size_t)group * TC_CHUNKS + s / 6) * T3_BM_SLOT, T3_BM_SLOT, &bm_full[sl]);
          ++nbm;
        }
        const uint32_t sl = na % T3_NA;
        mbar_wait(&a_empty[sl], ((na / T3_NA) & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[sl], TC_A_IMG);
        bulk_g2s(smem + sl * TC_A_IMG, a.apk + ((size_t)tile * TC_STAGES_PER_TILE + s) * TC_A_IMG, TC_A_IMG, &a_full[sl]);
        ++na;
      };
      if (i_beg < i_end)
        for (int s = 0; s < TC_STAGES_PER_TILE; ++s) blend_stage(i_beg, s);
      for (int item = i_beg; item < i_end; ++item) {
        const int group = item / NTILES, tile = item % NTILES;
        {
          const uint32_t sl = nwk & 1;
          mbar_wait(&wk_empty[sl], ((nwk >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&wk_full[sl], TC_W_BYTES);
          bulk_g2s(smem + T3_OFF_W + sl * TC_W_BYTES, a.wk + (size_t)tile * TC_W_BYTES, TC_W_BYTES, &wk_full[sl]);
          ++nwk;
        }
        for (int p = 0; p < 6; ++p) {
          if (item + 1 < i_end)
            for (int s = 4 * p; s < 4 * p + 4; ++s) blend_stage(item + 1, s);
          const uint32_t sl = nat & 1;
          mbar_wait(&at_empty[sl], ((nat >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&at_full[sl], T3_AT_SLOT);
          bulk_g2s(smem + T3_OFF_AT + sl * T3_AT_SLOT, a.atimg + (size_t)group * TC_AT_GROUP + (size_t)p * T3_AT_SLOT, T3_AT_SLOT, &at_full[sl]);
          ++nat;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer =================
      constexpr uint32_t idesc = umma_idesc_f16(TV, TC_NB);
      constexpr uint32_t idesc_t = umma_idesc_f16(TV, T3_TN);
      const uint32_t adesc0 = umma_desc_sw128_lo(smem0), bmdesc0 = umma_desc_sw128_lo(smem0 + T3_OFF_BM);
      const uint32_t wdesc0 = umma_desc_sw128_lo(smem0 + T3_OFF_W), atdesc0 = umma_desc_sw128_lo(smem0 + T3_OFF_AT);
      uint32_t na = 0, nbm = 0, nwk = 0, nat = 0, nt = 0;     // operands consumed so far / transform phases issued so far
      // one stage of the blend GEMM of the item whose accumulators are v_posed buffer `buf`
      auto blend_stage = [&](int s, uint32_t buf) {
        const int kc = s / 6, c = (s % 6) >> 1, h = s & 1;
        const uint32_t bsl = (nbm - (s % 6 == 0 ? 0 : 1)) & 1;      // the chunk's Bm slot: counted when its first stage is issued
        if (s % 6 == 0) {
          mbar_wait(&bm_full[nbm & 1], (nbm >> 1) & 1);
          ++nbm;
        }
        const uint32_t b_hi = bmdesc0 + bsl * (T3_BM_SLOT >> 4), b_lo = b_hi + (TC_B_IMG >> 4);
        const uint32_t sl = na % T3_NA;
        mbar_wait(&a_full[sl], (na / T3_NA) & 1);
        tc_fence_after();
        const uint32_t ad = adesc0 + sl * (TC_A_IMG >> 4);
        const uint32_t d = tmem_base + buf * (3 * TC_NB) + c * TC_NB;
        const int nks = min(4, TC_KSTEPS - 4 * kc);
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t ko = ks * 2;
          if (h == 0) {
            umma_f16_lohi(d, ad + ko, b_hi + ko, idesc, (kc | ks) != 0);
            umma_f16_lohi(d, ad + ko, b_lo + ko, idesc, 1);
          } else {
            umma_f16_lohi(d, ad + ko, b_hi + ko, idesc, 1);
          }
        }
        umma_commit(&a_empty[sl]);
        ++na;
        if (s % 6 == 5) umma_commit(&bm_empty[bsl]);
      };
      uint32_t nvp = 0;                                        // blend GEMMs started so far: buffer = nvp & 1
      if (i_beg < i_end) {
        // (buffer 0 is free at the start)
        for (int s = 0; s < TC_STAGES_PER_TILE; ++s) blend_stage(s, 0);
        umma_commit(&vp_full[0]);
        nvp = 1;
      }
      for (int item = i_beg; item < i_end; ++item) {
        const uint32_t wsl = nwk & 1;
        mbar_wait(&wk_full[wsl], (nwk >> 1) & 1);
        ++nwk;
        const uint32_t w_hi = wdesc0 + wsl * (TC_W_BYTES >> 4), w_lo = w_hi + (TC_W_IMG >> 4);
        const bool more = item + 1 < i_end;
        const uint32_t nbuf = nvp & 1;
        for (int p = 0; p < 6; ++p) {
          if (more) {
            if (p == 0 && nvp >= 2) {                          // the epilogue of two items ago has released this buffer
              mbar_wait(&vp_empty[nbuf], ((nvp >> 1) - 1) & 1);
              tc_fence_after();
            }
            for (int s = 4 * p; s < 4 * p + 4; ++s) blend_stage(s, nbuf);
            if (p == 5) umma_commit(&vp_full[nbuf]);
          }
          if (nt > 0) {
            mbar_wait(t_empty, (nt - 1) & 1);
            tc_fence_after();
          }
          const uint32_t asl = nat & 1;
          mbar_wait(&at_full[asl], (nat >> 1) & 1);
          tc_fence_after();
          const uint32_t t_hi = atdesc0 + asl * (T3_AT_SLOT >> 4), t_lo = t_hi + (T3_AT_SLOT >> 5);
          const uint32_t dt = tmem_base + T3_COL_T;
#pragma unroll
          for (int ks = 0; ks < TC_TK / 16; ++ks) {
            const uint32_t ko = ks * 2;
            umma_f16_lohi(dt, w_hi + ko, t_hi + ko, idesc_t, ks != 0, UMMA_DESC_SW64_HI);
            umma_f16_lohi(dt, w_hi + ko, t_lo + ko, idesc_t, 1, UMMA_DESC_SW64_HI);
            umma_f16_lohi(dt, w_lo + ko, t_hi + ko, idesc_t, 1, UMMA_DESC_SW64_HI);
          }
          umma_commit(t_full);
          umma_commit(&at_empty[asl]);
          ++nat; ++nt;
        }
        umma_commit(&wk_empty[wsl]);
        if (more) ++nvp;
      }
    }
  }
  __syncwarp();

  // ================= epilogue, warps 2..17: thread = vertex row (TMEM lane) x 8 bodies per phase =================
  if (warp >= 2) {
    const int quad = warp & 3, hh = (warp - 2) >> 2;    // a warp reads the TMEM lanes 32 (warp % 4) .. +31; hh = which 8 of a phase's 32 bodies
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    float* s_tr = reinterpret_cast<float*>(smem + T3_OFF_TR);
    float* slab = reinterpret_cast<float*>(smem + T3_OFF_ST) + (warp - 2) * 192;      // two slabs of 96 floats
    constexpr float tscale = 1.f / (TC_WSCALE * TC_BSCALE);
    uint32_t nt = 0, it = 0;
    for (int item = i_beg; item < i_end; ++item, ++it) {
      const int group = item / NTILES, tile = item % NTILES;
      const int v = tile * TV + row;
      const bool vok = v < V;
      // transl of the group's bodies (everybody is past the previous item's reads of it after the first barrier)
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      for (int i = threadIdx.x - 64; i < TC_NB * 3; i += T3_EPI_THREADS) {
        const int gb = group * TC_NB + i / 3;
        s_tr[i] = (a.transl && gb < a.B) ? a.transl[(size_t)gb * 3 + i % 3] : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EPI_THREADS) : "memory");
      float ainv[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) ainv[c] = a.ainv[(tile * 3 + c) * TV + row];
      const float k0 = ainv[0] * tscale, k1 = ainv[1] * tscale, k2 = ainv[2] * tscale;
      const int nf2 = (max(0, min(32, V - (tile * TV + quad * 32))) * 3) / 2;       // float2 of this warp's valid vertices (48 or 15)
      const uint32_t buf = it & 1;
      mbar_wait(&vp_full[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int bl0 = half * 32 + hh * 8;                   // first of this thread's 8 bodies inside the group
        uint32_t d[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c) tmem_ld_32x8(tlane + buf * (3 * TC_NB) + c * TC_NB + bl0, d[c]);
        tmem_ld_wait();
        if (a.save_vposed && vok) {
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            const int gb = group * TC_NB + bl0 + b;
            if (gb < a.B) {
              float* sv = a.save_vposed + ((size_t)gb * V + v) * 3;
#pragma unroll
              for (int c = 0; c < 3; ++c) sv[c] = __uint_as_float(d[c][b]) * ainv[c];
            }
          }
        }
        // The three output rows of a body are kept until the last one: a vertex's 12 bytes leave together, and a warp's 32 vertices
        // (384 contiguous bytes of the output) go through a per-warp slab so that the stores are full 8-byte-per-lane lines.  Storing
        // each coordinate as it was produced (4 bytes every 12: 36 partially written sectors per body and warp instead of 12 full
        // ones) made the epilogue the pacing stage -- every FFMA waited for the store queue to release its register
        // (profiles/r02_lbs_tc_ncu.txt).
        float o01[2][8];
#pragma unroll
        for (int r = 0; r < 3; ++r, ++nt) {
          mbar_wait(t_full, nt & 1);
          tc_fence_after();
          {
            uint32_t t[32];                                   // 8 bodies x 4 columns of T
            tmem_ld_32x32(tlane + T3_COL_T + hh * 32, t);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(t_empty);                             // T is in registers: the issuer may overwrite it
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) {
              const int b = bb, gb = group * TC_NB + bl0 + b;
              float acc = fmaf(__uint_as_float(t[bb * 4 + 3]), tscale, s_tr[(bl0 + b) * 3 + r]);
              acc = fmaf(__uint_as_float(t[bb * 4 + 0]) * k0, __uint_as_float(d[0][b]), acc);
              acc = fmaf(__uint_as_float(t[bb * 4 + 1]) * k1, __uint_as_float(d[1][b]), acc);
              acc = fmaf(__uint_as_float(t[bb * 4 + 2]) * k2, __uint_as_float(d[2][b]), acc);
              if (r == 0) o01[0][b] = acc;
              else if (r == 1) o01[1][b] = acc;
              else if (gb < a.B) {                            // warp-uniform
                float* sl = slab + (b & 1) * 96;
                sl[lane * 3 + 0] = o01[0][b]; sl[lane * 3 + 1] = o01[1][b]; sl[lane * 3 + 2] = acc;      // 3 words apart: conflict free
                __syncwarp();
                float2* dst = reinterpret_cast<float2*>(a.verts + ((size_t)gb * V + (size_t)tile * TV + quad * 32) * 3);
                const float2* src = reinterpret_cast<const float2*>(sl);
                if (lane < nf2) dst[lane] = src[lane];
                if (lane + 32 < nf2) dst[lane + 32] = src[lane + 32];
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&vp_empty[buf]);                            // this buffer's accumulators have been read
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace straps

using namespace straps;

// ---- host side: pack the constant operands once, run the two kernels ----
static inline uint16_t f2h_bits(float f) {          // round-to-nearest-even fp32 -> fp16 bits (cuda_fp16.h host path)
  const __half h = __float2half_rn(f);
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}
static inline float h2f_bits(uint16_t b) {
  __half h;
  memcpy(&h, &b, 2);
  return __half2float(h);
}

// Apk[(v, c), k] rows scaled to [2^13, 2^14), split into fp16 hi / lo, laid out as the SWIZZLE_128B images lbs_tc_kernel streams:
// image (tile, kc, c, h) = 128 rows of 128 bytes, 16-byte unit u of row r at physical unit u ^ (r & 7).
// Wk[v, j] = skinning weight * 2^14, hi / lo, image (tile, h) = 128 rows of 64 bytes (32 joint slots), SWIZZLE_64B: unit u of row r at
// physical unit u ^ ((r >> 1) & 3).
int straps::smpl_tc_pack(const float* v_template, const float* shapedirs, const float* posedirs, const float* lbs_weights,
                         std::vector<unsigned char>& apk, std::vector<float>& ainv, std::vector<unsigned char>& wk) {
  apk.assign((size_t)NTILES * TC_STAGES_PER_TILE * TC_A_IMG, 0);
  ainv.assign((size_t)NTILES * 3 * TV, 0.f);
  wk.assign((size_t)NTILES * TC_W_BYTES, 0);
  std::vector<float> rowv(TC_K);
  for (int v = 0; v < V; ++v) {
    const int tile = v / TV, r = v % TV;
    for (int c = 0; c < 3; ++c) {
      float mx = 0.f;
      for (int k = 0; k < TC_K; ++k) {
        float x = 0.f;
        if (k < NPF) x = posedirs[(size_t)k * V * 3 + v * 3 + c];
        else if (k < NPF + STRAPS_NUM_BETAS) x = shapedirs[((size_t)v * 3 + c) * STRAPS_NUM_BETAS + (k - NPF)];
        else if (k == NPF + STRAPS_NUM_BETAS) x = v_template[v * 3 + c];
        rowv[k] = x;
        mx = std::max(mx, std::fabs(x));
      }
      int e = 0;
      if (mx > 0.f) { std::frexp(mx, &e); e = 14 - e; }                        // mx * 2^e in [2^13, 2^14)
      const float sc = std::ldexp(1.f, e);
      ainv[((size_t)tile * 3 + c) * TV + r] = 1.f / (sc * TC_BSCALE);
      for (int k = 0; k < TC_K; ++k) {
        const float x = rowv[k] * sc;
        const uint16_t hb = f2h_bits(x);
        const uint16_t lb = f2h_bits(x - h2f_bits(hb));
        const int kc = k >> 6, u = (k & 63) >> 3, el = k & 7;
        const size_t img = ((size_t)tile * TC_STAGES_PER_TILE + (size_t)(kc * 3 + c) * 2) * TC_A_IMG;
        const size_t off = (size_t)r * 128 + (size_t)((u ^ (r & 7)) << 4) + (size_t)el * 2;
        memcpy(&apk[img + off], &hb, 2);
        memcpy(&apk[img + TC_A_IMG + off], &lb, 2);
      }
    }
    for (int j = 0; j < NJ; ++j) {
      const float x = lbs_weights[(size_t)v * NJ + j] * TC_WSCALE;
      const uint16_t hb = f2h_bits(x);
      const uint16_t lb = f2h_bits(x - h2f_bits(hb));
      const size_t img = (size_t)tile * TC_W_BYTES;
      const size_t off = (size_t)r * 64 + (size_t)(((j >> 3) ^ ((r >> 1) & 3)) << 4) + (size_t)(j & 7) * 2;
      memcpy(&wk[img + off], &hb, 2);
      memcpy(&wk[img + TC_W_IMG + off], &lb, 2);
    }
  }
  return 0;
}

int straps::smpl_tc_forward(straps_smpl* m, const float* global_orient, int64_t go_stride, const float* body_pose, int64_t bp_stride,
                            const float* betas, int64_t betas_stride, const float* transl, int batch, int pose2rot, float* vertices,
                            float* joints, float* save_vposed, float* save_A, cudaStream_t st) {
  const int groups = (batch + TC_NB - 1) / TC_NB;
  const int b_pad = groups * TC_NB;
  const size_t need = (size_t)groups * (TC_B_BYTES + TC_AT_GROUP);
  if (need > m->tc_scratch_bytes) {
    // grows with the largest batch seen (first call / warm-up; never inside a CUDA-graph capture of a warmed-up shape)
    STRAPS_CUDA(cudaStreamSynchronize(st));
    if (m->tc_scratch) cudaFree(m->tc_scratch);
    m->tc_scratch = nullptr;
    m->tc_scratch_bytes = 0;
    STRAPS_CUDA(cudaMalloc(&m->tc_scratch, need));
    STRAPS_CUDA(cudaMemsetAsync(m->tc_scratch, 0, need, st));        // the unused joint slots of the transform images stay zero
    m->tc_scratch_bytes = need;
  }
  unsigned char* bimg = static_cast<unsigned char*>(m->tc_scratch);
  unsigned char* atimg = bimg + (size_t)groups * TC_B_BYTES;
  if (atimg != m->tc_at_base) {
    // a smaller batch moves the boundary between the two regions: what was Bm data must not be read as joint slots 24..31
    STRAPS_CUDA(cudaMemsetAsync(atimg, 0, (size_t)groups * TC_AT_GROUP, st));
    m->tc_at_base = atimg;
  }
  ChainArgs ca;
  ca.go = global_orient; ca.bp = body_pose; ca.betas = betas; ca.transl = transl;
  ca.go_stride = go_stride; ca.bp_stride = bp_stride; ca.betas_stride = betas_stride;
  ca.B = batch; ca.B_pad = b_pad; ca.pose2rot = pose2rot; ca.joints = joints; ca.bimg = bimg; ca.atimg = atimg; ca.save_A = save_A;
  // STRAPS_LBS_V=2: the one-CTA-per-item kernel (lbs_tc_kernel); default: the persistent pipelined one (lbs_tc3_kernel)
  const char* ver = getenv("STRAPS_LBS_V");
  const bool v3 = !(ver && ver[0] == '2');
  ca.at_layout = v3 ? 1 : 0;
  smpl_chain_kernel<<<b_pad / CH_TB, CH_THREADS, 0, st>>>(m->d, ca);
  STRAPS_LAUNCH_CHECK();
  static PerDeviceOnce attr_once;
  static int num_sms[64];
  const int dev = current_device();
  if (attr_once.need(dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(lbs_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    STRAPS_CUDA(cudaFuncSetAttribute(lbs_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM));
    int n = 148;
    STRAPS_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev < 0 ? 0 : dev));
    if (dev >= 0 && dev < 64) num_sms[dev] = n;
    attr_once.done(dev);
  }
  LbsTcArgs la;
  la.apk = m->tc_apk; la.ainv = m->tc_ainv; la.wk = m->tc_wk; la.bimg = bimg; la.atimg = atimg;
  la.transl = transl; la.B = batch; la.verts = vertices; la.save_vposed = save_vposed;
  if (v3) {
    LbsTc3Args a3;
    a3.k = la; a3.n_items = groups * NTILES;
    const int sms = (dev >= 0 && dev < 64 && num_sms[dev] > 0) ? num_sms[dev] : 148;
    lbs_tc3_kernel<<<std::min(a3.n_items, sms), T3_THREADS, T3_SMEM, st>>>(a3);
  } else {
    lbs_tc_kernel<<<dim3(NTILES, groups), TC_THREADS, TC_SMEM, st>>>(la);
  }
  STRAPS_LAUNCH_CHECK();
  return 0;
}

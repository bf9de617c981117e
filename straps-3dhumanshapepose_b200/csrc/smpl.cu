// smpl.cu -- fused SMPL forward for sm_100a: (Rodrigues | rotmats) -> kinematic chain -> shape blend +
// pose-corrective blend + linear blend skinning -> 90-joint superset.
//
// Replaces smplx.SMPL.forward / smplx.lbs.lbs (third-party) and the extension in the
// reference's models/smpl_official.py:27-41 (S8), plus utils/rigid_transform_utils.py:27-41 and
// utils/cam_utils.py:5-26.
//
// Data layout in HBM (packed once by straps_smpl_create; VP3 = 6912*3, 6912 = 54 tiles of 128 vertices):
//   vt   [VP3]            v_template, (vertex, xyz) inner order, zero padded
//   sdir [10][VP3]        shapedirs, beta-major   (so a CTA's slab is one contiguous run per beta)
//   pdir [208][VP3]       posedirs, pose-feature-major (smplx layout, row 207 zero, rows 16-byte aligned)
//   jt [24*3], js [24*3][10]   rest joints pre-reduced through the joint regressor:
//                         J(beta) = J_regressor.v_template + (J_regressor.shapedirs).beta   (hard part 4)
//   widx/wval [6912][4]   the <= 4 non-zero skinning weights per vertex (dense [6912][24] fallback otherwise)
//   CSR of the 45 extra regressor rows + 21 vertex picks for the joint superset.
//
// Kernel 1 (lbs_kernel): CTA = 128 vertices x TB bodies, one thread per vertex, TB bodies register-tiled.
//   The posedirs slab [207][384] streams through a 3-stage shared-memory ring filled by 1-D bulk-async (TMA)
//   copies; the per-body 24 rigid transforms are rebuilt per CTA in shared memory (9 tree levels) while the
//   first slabs are in flight.  Algorithmic HBM bytes: 20,587,320 constants + 84,664 B/body;
//   the pose-corrective term costs 621 FMA per (vertex, body) and is the FP32-FMA bound at B >= 64.
// Kernel 2 (joints_kernel): one warp per (body, superset joint 24..89): picks are plain copies (bit exact),
//   regressed rows are CSR dot products over the just-written (L2 resident) vertices.
#include "smpl.h"
#include <cmath>
#include <cstdlib>

namespace straps {

struct LbsArgs {
  const float* go;
  const float* bp;
  const float* betas;
  const float* transl;
  long long go_stride, bp_stride, betas_stride;
  int B;
  int pose2rot;
  float* verts;
  float* joints;
  float* save_vposed;   // [B,6890,3] or null (training: input of lbs_bwd_kernel)
  float* save_A;        // [B,24,12]  or null
};

// Prologue scratch (global transforms G, rotations R, rest joints Jr) lives in ring stages 1..2, which are not
// filled until the prologue is over: 53 KB of shared memory per CTA instead of 71 KB -> 4 CTAs (16 warps) per SM.
template <int TB>
struct LbsScratch {
  float G[TB][NJ][12];
  float R[TB][NJ][9];
  float Jr[TB][NJ][3];
};
static_assert(sizeof(LbsScratch<8>) <= 2 * KC * ROWF * sizeof(float), "prologue scratch must fit in ring stages 1..2");

template <int TB>
struct LbsSmem {
  float pbuf[NSTAGE][KC][ROWF];   // posedirs ring; reused as the output staging tile
  float pf[NPF_PAD][TB];          // pose feature, body-minor (float4 broadcast loads)
  float4 A[TB][NJ][3];            // skinning transforms, rows [R | t]
  float beta[TB][STRAPS_NUM_BETAS];
  float tr[TB][4];
  uint64_t full[NSTAGE];
  uint64_t empty[NSTAGE];
};

__device__ __forceinline__ void rodrigues(const float r[3], float* R) {
  // smplx.lbs.batch_rodrigues: the epsilon goes inside the norm
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

template <int TB, bool SPARSE>
__global__ void __launch_bounds__(TV) lbs_kernel(const SmplDev m, const LbsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LbsSmem<TB>& s = *reinterpret_cast<LbsSmem<TB>*>(smem_raw);
  LbsScratch<TB>& x = *reinterpret_cast<LbsScratch<TB>*>(&s.pbuf[1][0][0]);
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int b0 = blockIdx.y * TB;
  const int nb = min(TB, a.B - b0);

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], TV / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const float* slab = m.pdir + (size_t)tile * ROWF;
  auto issue_chunk = [&](int chunk) {
    const int st = chunk % NSTAGE;
    mbar_arrive_expect_tx(&s.full[st], KC * ROWF * 4);
#pragma unroll
    for (int r = 0; r < KC; ++r)
      bulk_g2s(&s.pbuf[st][r][0], slab + (size_t)(chunk * KC + r) * VP3, ROWF * 4, &s.full[st]);
  };
  if (tid == 0) issue_chunk(0);

  // ---- prologue: rotations, pose feature, rest joints, kinematic chain (overlaps the first slab loads) ----
  for (int i = tid; i < TB * STRAPS_NUM_BETAS; i += TV) {
    int b = i / STRAPS_NUM_BETAS, l = i % STRAPS_NUM_BETAS;
    s.beta[b][l] = (b < nb) ? a.betas[(size_t)(b0 + b) * a.betas_stride + l] : 0.f;
  }
  for (int i = tid; i < TB * 3; i += TV) {
    int b = i / 3, c = i % 3;
    s.tr[b][c] = (b < nb && a.transl) ? a.transl[(size_t)(b0 + b) * 3 + c] : 0.f;
  }
  if (a.pose2rot) {
    for (int i = tid; i < TB * NJ; i += TV) {
      int b = i / NJ, j = i % NJ;
      float r[3] = {0.f, 0.f, 0.f};
      if (b < nb) {
        const float* src = (j == 0) ? a.go + (size_t)(b0 + b) * a.go_stride
                                    : a.bp + (size_t)(b0 + b) * a.bp_stride + (j - 1) * 3;
        r[0] = src[0]; r[1] = src[1]; r[2] = src[2];
      }
      rodrigues(r, &x.R[b][j][0]);
    }
  } else {
    for (int i = tid; i < TB * NJ * 9; i += TV) {
      int b = i / (NJ * 9), r = i % (NJ * 9), j = r / 9, e = r % 9;
      float v = (e == 0 || e == 4 || e == 8) ? 1.f : 0.f;
      if (b < nb)
        v = (j == 0) ? a.go[(size_t)(b0 + b) * a.go_stride + e]
                     : a.bp[(size_t)(b0 + b) * a.bp_stride + (j - 1) * 9 + e];
      x.R[b][j][e] = v;
    }
  }
  __syncthreads();
  for (int i = tid; i < NPF_PAD * TB; i += TV) {
    int k = i / TB, b = i % TB;
    float v = 0.f;
    if (k < NPF) {
      int e = k % 9;
      v = x.R[b][1 + k / 9][e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
    }
    s.pf[k][b] = v;
  }
  for (int i = tid; i < TB * NJ * 3; i += TV) {
    int b = i / (NJ * 3), jc = i % (NJ * 3);
    float acc = 0.f;
#pragma unroll
    for (int l = 0; l < STRAPS_NUM_BETAS; ++l) acc = fmaf(m.js[jc * STRAPS_NUM_BETAS + l], s.beta[b][l], acc);
    x.Jr[b][jc / 3][jc % 3] = m.jt[jc] + acc;
  }
  __syncthreads();
  for (int lvl = 0; lvl < m.nlevels; ++lvl) {
    const int j0 = m.lvl_start[lvl], nj = m.lvl_start[lvl + 1] - j0;
    for (int i = tid; i < TB * nj * 3; i += TV) {
      int b = i / (nj * 3), q = i % (nj * 3), j = m.lvl_joint[j0 + q / 3], r = q % 3;
      int p = m.parents[j];
      const float* Rj = &x.R[b][j][0];
      float* g = &x.G[b][j][r * 4];
      if (p < 0) {
        g[0] = Rj[r * 3 + 0]; g[1] = Rj[r * 3 + 1]; g[2] = Rj[r * 3 + 2]; g[3] = x.Jr[b][j][r];
      } else {
        const float* gp = &x.G[b][p][r * 4];
        float rel0 = x.Jr[b][j][0] - x.Jr[b][p][0];
        float rel1 = x.Jr[b][j][1] - x.Jr[b][p][1];
        float rel2 = x.Jr[b][j][2] - x.Jr[b][p][2];
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = gp[0] * Rj[0 * 3 + c] + gp[1] * Rj[1 * 3 + c] + gp[2] * Rj[2 * 3 + c];
        g[3] = gp[0] * rel0 + gp[1] * rel1 + gp[2] * rel2 + gp[3];
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < TB * NJ * 3; i += TV) {
    int b = i / (NJ * 3), q = i % (NJ * 3), j = q / 3, r = q % 3;
    const float* g = &x.G[b][j][r * 4];
    float t = g[3] - (g[0] * x.Jr[b][j][0] + g[1] * x.Jr[b][j][1] + g[2] * x.Jr[b][j][2]);
    s.A[b][j][r] = make_float4(g[0], g[1], g[2], t);
    if (tile == 0 && b < nb && a.save_A)
      *reinterpret_cast<float4*>(a.save_A + ((size_t)(b0 + b) * NJ + j) * 12 + r * 4) = make_float4(g[0], g[1], g[2], t);
    if (tile == 0 && b < nb) a.joints[((size_t)(b0 + b) * STRAPS_NUM_SUPERSET_JOINTS + j) * 3 + r] = g[3] + s.tr[b][r];
  }
  __syncthreads();
  if (tid == 0) issue_chunk(1);   // the scratch in stages 1..2 is dead from here on

  // ---- main loop: pose-corrective blend, acc[b][c] = sum_k pf[k][b] * posedirs[k][v*3+c] ----
  // The loop is bound by instruction issue (ncu: FMA pipe 24 %, issue slots 43 % busy), so for an even number of bodies it
  // uses Blackwell's packed fp32 FMA (fma.rn.f32x2 -> FFMA2): two bodies per instruction, each lane an IEEE fma, i.e.
  // bit-identical to the scalar form.  Per k: 12 FFMA2 + 3 MOV + 5 LDS instead of 24 FFMA + 5 LDS (TB = 8).
  float acc[TB][3];
  if constexpr (TB % 2 == 0) {
    unsigned long long acc2[TB / 2][3];
#pragma unroll
    for (int b = 0; b < TB / 2; ++b) acc2[b][0] = acc2[b][1] = acc2[b][2] = 0ull;
    for (int chunk = 0; chunk < NCHUNK; ++chunk) {
      const int st = chunk % NSTAGE;
      if (tid == 0 && chunk + 2 < NCHUNK) {
        const int nxt = chunk + 2;
        if (nxt >= NSTAGE) mbar_wait(&s.empty[nxt % NSTAGE], ((nxt / NSTAGE) - 1) & 1);
        issue_chunk(nxt);
      }
      mbar_wait(&s.full[st], (chunk / NSTAGE) & 1);
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const float* row = &s.pbuf[st][kk][tid * 3];
        unsigned long long pp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) asm("mov.b64 %0, {%1, %1};" : "=l"(pp[c]) : "f"(row[c]));
        const unsigned long long* w2 = reinterpret_cast<const unsigned long long*>(&s.pf[chunk * KC + kk][0]);
        unsigned long long w[TB / 2];
        if constexpr (TB % 4 == 0) {
#pragma unroll
          for (int q = 0; q < TB / 4; ++q) {
            const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(w2 + 2 * q);
            w[2 * q] = t.x; w[2 * q + 1] = t.y;
          }
        } else {
#pragma unroll
          for (int b = 0; b < TB / 2; ++b) w[b] = w2[b];
        }
#pragma unroll
        for (int b = 0; b < TB / 2; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[b][c]) : "l"(w[b]), "l"(pp[c]));
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&s.empty[st]);
    }
#pragma unroll
    for (int b = 0; b < TB / 2; ++b)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc[2 * b][c] = __uint_as_float((unsigned)(acc2[b][c] & 0xffffffffull));
        acc[2 * b + 1][c] = __uint_as_float((unsigned)(acc2[b][c] >> 32));
      }
  } else {
#pragma unroll
    for (int b = 0; b < TB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0.f;
    for (int chunk = 0; chunk < NCHUNK; ++chunk) {
      const int st = chunk % NSTAGE;
      if (tid == 0 && chunk + 2 < NCHUNK) {
        const int nxt = chunk + 2;
        if (nxt >= NSTAGE) mbar_wait(&s.empty[nxt % NSTAGE], ((nxt / NSTAGE) - 1) & 1);
        issue_chunk(nxt);
      }
      mbar_wait(&s.full[st], (chunk / NSTAGE) & 1);
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const float* row = &s.pbuf[st][kk][tid * 3];
        const float p0 = row[0], p1 = row[1], p2 = row[2];
        const float* pfk = &s.pf[chunk * KC + kk][0];
#pragma unroll
        for (int b = 0; b < TB; ++b) {
          acc[b][0] = fmaf(pfk[b], p0, acc[b][0]);
          acc[b][1] = fmaf(pfk[b], p1, acc[b][1]);
          acc[b][2] = fmaf(pfk[b], p2, acc[b][2]);
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&s.empty[st]);
    }
  }

  // ---- epilogue: shape blend, skinning ----
  const int v = tile * TV + tid;   // < VPAD, packed arrays are zero padded
  float vt3[3], sd[STRAPS_NUM_BETAS][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) vt3[c] = m.vt[v * 3 + c];
#pragma unroll
  for (int l = 0; l < STRAPS_NUM_BETAS; ++l)
#pragma unroll
    for (int c = 0; c < 3; ++c) sd[l][c] = m.sdir[(size_t)l * VP3 + v * 3 + c];

  int wj[4];
  float ww[4];
  if constexpr (SPARSE) {
    const int4 ji = *reinterpret_cast<const int4*>(m.widx + v * 4);
    const float4 jw = *reinterpret_cast<const float4*>(m.wval + v * 4);
    wj[0] = ji.x; wj[1] = ji.y; wj[2] = ji.z; wj[3] = ji.w;
    ww[0] = jw.x; ww[1] = jw.y; ww[2] = jw.z; ww[3] = jw.w;
  }
  __syncthreads();   // everyone is done with the posedirs ring -> reuse it as the output tile
  float* otile = &s.pbuf[0][0][0];   // [TB][384]
#pragma unroll
  for (int b = 0; b < TB; ++b) {
    float vp[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float bs = 0.f;
#pragma unroll
      for (int l = 0; l < STRAPS_NUM_BETAS; ++l) bs = fmaf(s.beta[b][l], sd[l][c], bs);
      vp[c] = acc[b][c] + (vt3[c] + bs);   // v_posed = pose_offsets + v_shaped
    }
    if (a.save_vposed && b < nb && v < V) {
      float* sv = a.save_vposed + ((size_t)(b0 + b) * V + v) * 3;
      sv[0] = vp[0]; sv[1] = vp[1]; sv[2] = vp[2];
    }
    float4 T0 = make_float4(0.f, 0.f, 0.f, 0.f), T1 = T0, T2 = T0;
    if constexpr (SPARSE) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 r0 = s.A[b][wj[e]][0], r1 = s.A[b][wj[e]][1], r2 = s.A[b][wj[e]][2];
        const float w = ww[e];
        T0.x = fmaf(w, r0.x, T0.x); T0.y = fmaf(w, r0.y, T0.y); T0.z = fmaf(w, r0.z, T0.z); T0.w = fmaf(w, r0.w, T0.w);
        T1.x = fmaf(w, r1.x, T1.x); T1.y = fmaf(w, r1.y, T1.y); T1.z = fmaf(w, r1.z, T1.z); T1.w = fmaf(w, r1.w, T1.w);
        T2.x = fmaf(w, r2.x, T2.x); T2.y = fmaf(w, r2.y, T2.y); T2.z = fmaf(w, r2.z, T2.z); T2.w = fmaf(w, r2.w, T2.w);
      }
    } else {
      for (int j = 0; j < NJ; ++j) {
        const float w = m.wdense[(size_t)v * NJ + j];
        const float4 r0 = s.A[b][j][0], r1 = s.A[b][j][1], r2 = s.A[b][j][2];
        T0.x = fmaf(w, r0.x, T0.x); T0.y = fmaf(w, r0.y, T0.y); T0.z = fmaf(w, r0.z, T0.z); T0.w = fmaf(w, r0.w, T0.w);
        T1.x = fmaf(w, r1.x, T1.x); T1.y = fmaf(w, r1.y, T1.y); T1.z = fmaf(w, r1.z, T1.z); T1.w = fmaf(w, r1.w, T1.w);
        T2.x = fmaf(w, r2.x, T2.x); T2.y = fmaf(w, r2.y, T2.y); T2.z = fmaf(w, r2.z, T2.z); T2.w = fmaf(w, r2.w, T2.w);
      }
    }
    otile[b * ROWF + tid * 3 + 0] = T0.x * vp[0] + T0.y * vp[1] + T0.z * vp[2] + T0.w + s.tr[b][0];
    otile[b * ROWF + tid * 3 + 1] = T1.x * vp[0] + T1.y * vp[1] + T1.z * vp[2] + T1.w + s.tr[b][1];
    otile[b * ROWF + tid * 3 + 2] = T2.x * vp[0] + T2.y * vp[1] + T2.z * vp[2] + T2.w + s.tr[b][2];
  }
  __syncthreads();
  const int nvalid2 = (min(TV, V - tile * TV) * 3) / 2;   // float2 count (always even number of floats)
  for (int b = 0; b < nb; ++b) {
    float2* dst = reinterpret_cast<float2*>(a.verts + (size_t)(b0 + b) * V * 3 + (size_t)tile * ROWF);
    const float2* src = reinterpret_cast<const float2*>(otile + b * ROWF);
    for (int i = tid; i < nvalid2; i += TV) dst[i] = src[i];
  }
}

// one warp per (body, superset joint 24..89).  (Eight bodies per warp -- CSR entries read once, eight gathers in flight -- measured
// SLOWER on B200, 71 vs 64 us at B = 4096: the kernel is bound by the scattered 12-byte gathers from a 338 MB tensor, not by latency.)
__global__ void __launch_bounds__(256) joints_kernel(const SmplDev m, const float* __restrict__ verts,
                                                     float* __restrict__ joints, int B) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int NOUT = STRAPS_NUM_EXTRA_PICKS + STRAPS_NUM_EXTRA_ROWS;   // 66
  if (warp >= B * NOUT) return;
  const int b = warp / NOUT, q = warp % NOUT;
  const float* vb = verts + (size_t)b * V * 3;
  float* out = joints + ((size_t)b * STRAPS_NUM_SUPERSET_JOINTS + NJ + q) * 3;
  if (q < STRAPS_NUM_EXTRA_PICKS) {
    if (lane < 3) out[lane] = vb[(size_t)m.pick_idx[q] * 3 + lane];   // plain copy: bit exact
    return;
  }
  const int r = q - STRAPS_NUM_EXTRA_PICKS;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int i = m.csr_ptr[r] + lane; i < m.csr_ptr[r + 1]; i += 32) {
    const float w = m.csr_val[i];
    const float* p = vb + (size_t)m.csr_idx[i] * 3;
    sx = fmaf(w, p[0], sx); sy = fmaf(w, p[1], sy); sz = fmaf(w, p[2], sz);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if (lane == 0) { out[0] = sx; out[1] = sy; out[2] = sz; }
}

__global__ void rot6d_kernel(const float* __restrict__ x, long long n, float* __restrict__ R) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // x.view(-1,3,2): a1 = x[0],x[2],x[4]; a2 = x[1],x[3],x[5]  (utils/rigid_transform_utils.py:35-37)
  const float* p = x + i * 6;
  float a1[3] = {p[0], p[2], p[4]}, a2[3] = {p[1], p[3], p[5]};
  float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);   // F.normalize eps
  float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
  float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
  float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
  float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
  float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
  float* o = R + i * 9;   // columns (b1, b2, b3)
#pragma unroll
  for (int r = 0; r < 3; ++r) { o[r * 3 + 0] = b1[r]; o[r * 3 + 1] = b2[r]; o[r * 3 + 2] = b3[r]; }
}

__global__ void ortho_kernel(const float* __restrict__ pts, const float* __restrict__ cam, long long cam_stride,
                             int B, int N, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  int b = i / N;
  const float s = cam[b * cam_stride + 0], tx = cam[b * cam_stride + 1], ty = cam[b * cam_stride + 2];
  out[i * 2 + 0] = s * (pts[(size_t)i * 3 + 0] + tx);
  out[i * 2 + 1] = s * (pts[(size_t)i * 3 + 1] + ty);
}

}  // namespace straps

using namespace straps;


template <typename T>
static int upload(straps_smpl* m, const std::vector<T>& h, const T** dst) {
  void* p = nullptr;
  size_t bytes = h.size() * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  STRAPS_CUDA(cudaMalloc(&p, bytes));
  m->allocs.push_back(p);
  if (!h.empty()) STRAPS_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dst = static_cast<const T*>(p);
  return 0;
}

extern "C" int straps_smpl_create(straps_smpl_t** out, const float* v_template, const float* shapedirs,
                                  const float* posedirs, const float* J_regressor, const float* lbs_weights,
                                  const int64_t* parents, const float* extra_regressors,
                                  const int64_t* extra_pick_idx) {
  STRAPS_CHECK(out && v_template && shapedirs && posedirs && J_regressor && lbs_weights && parents &&
                   extra_regressors && extra_pick_idx,
               "straps_smpl_create: null argument");
  straps_smpl* m = new straps_smpl();
  memset(&m->d, 0, sizeof(SmplDev));
  m->tc_apk = nullptr; m->tc_ainv = nullptr; m->tc_wk = nullptr; m->tc_at_base = nullptr; m->tc_scratch = nullptr; m->tc_scratch_bytes = 0;
  // --- kinematic tree -> levels
  int depth[NJ];
  for (int j = 0; j < NJ; ++j) {
    int p = (int)parents[j];
    if (j == 0) p = -1;
    if (!(p < j)) { delete m; STRAPS_CHECK(false, "straps_smpl_create: parents[%d]=%d must precede the joint", j, p); }
    m->d.parents[j] = p;
    depth[j] = (p < 0) ? 0 : depth[p] + 1;
  }
  int nl = 0, pos = 0;
  for (int j = 0; j < NJ; ++j) nl = depth[j] + 1 > nl ? depth[j] + 1 : nl;
  for (int l = 0; l < nl; ++l) {
    m->d.lvl_start[l] = pos;
    for (int j = 0; j < NJ; ++j)
      if (depth[j] == l) m->d.lvl_joint[pos++] = j;
  }
  m->d.lvl_start[nl] = pos;
  m->d.nlevels = nl;

  std::vector<float> vt(VP3, 0.f), sdir((size_t)STRAPS_NUM_BETAS * VP3, 0.f), pdir((size_t)NPF_PAD * VP3, 0.f);
  for (int i = 0; i < V * 3; ++i) vt[i] = v_template[i];
  for (int v = 0; v < V; ++v)
    for (int c = 0; c < 3; ++c)
      for (int l = 0; l < STRAPS_NUM_BETAS; ++l)
        sdir[(size_t)l * VP3 + v * 3 + c] = shapedirs[((size_t)v * 3 + c) * STRAPS_NUM_BETAS + l];
  for (int k = 0; k < NPF; ++k) memcpy(&pdir[(size_t)k * VP3], posedirs + (size_t)k * V * 3, sizeof(float) * V * 3);
  // --- pre-reduced rest joints (fp64 accumulate)
  std::vector<float> jt(NJ * 3), js(NJ * 3 * STRAPS_NUM_BETAS);
  for (int j = 0; j < NJ; ++j)
    for (int c = 0; c < 3; ++c) {
      double t = 0.0, sl[STRAPS_NUM_BETAS] = {0};
      for (int v = 0; v < V; ++v) {
        const double w = J_regressor[(size_t)j * V + v];
        if (w == 0.0) continue;
        t += w * v_template[v * 3 + c];
        for (int l = 0; l < STRAPS_NUM_BETAS; ++l) sl[l] += w * shapedirs[((size_t)v * 3 + c) * STRAPS_NUM_BETAS + l];
      }
      jt[j * 3 + c] = (float)t;
      for (int l = 0; l < STRAPS_NUM_BETAS; ++l) js[(j * 3 + c) * STRAPS_NUM_BETAS + l] = (float)sl[l];
    }
  // --- skinning weights: <= 4 non-zeros per vertex?
  std::vector<int> widx((size_t)VPAD * 4, 0);
  std::vector<float> wval((size_t)VPAD * 4, 0.f), wdense;
  int sparse4 = 1;
  for (int v = 0; v < V && sparse4; ++v) {
    int n = 0;
    for (int j = 0; j < NJ; ++j)
      if (lbs_weights[(size_t)v * NJ + j] != 0.f) {
        if (n == 4) { sparse4 = 0; break; }
        widx[v * 4 + n] = j;
        wval[v * 4 + n] = lbs_weights[(size_t)v * NJ + j];
        ++n;
      }
  }
  if (!sparse4) {
    wdense.assign((size_t)VPAD * NJ, 0.f);
    memcpy(wdense.data(), lbs_weights, sizeof(float) * V * NJ);
  }
  m->sparse4 = sparse4;
  // --- joint superset: picks + CSR of the 45 regressed rows
  std::vector<int> pick(STRAPS_NUM_EXTRA_PICKS), ptr(STRAPS_NUM_EXTRA_ROWS + 1, 0), idx;
  std::vector<float> val;
  for (int i = 0; i < STRAPS_NUM_EXTRA_PICKS; ++i) {
    if (extra_pick_idx[i] < 0 || extra_pick_idx[i] >= V) { delete m; STRAPS_CHECK(false, "straps_smpl_create: pick index out of range"); }
    pick[i] = (int)extra_pick_idx[i];
  }
  for (int r = 0; r < STRAPS_NUM_EXTRA_ROWS; ++r) {
    for (int v = 0; v < V; ++v) {
      float w = extra_regressors[(size_t)r * V + v];
      if (w != 0.f) { idx.push_back(v); val.push_back(w); }
    }
    ptr[r + 1] = (int)idx.size();
  }
  int rc = 0;
  rc |= upload(m, vt, &m->d.vt);
  rc |= upload(m, sdir, &m->d.sdir);
  rc |= upload(m, pdir, &m->d.pdir);
  rc |= upload(m, jt, &m->d.jt);
  rc |= upload(m, js, &m->d.js);
  rc |= upload(m, widx, &m->d.widx);
  rc |= upload(m, wval, &m->d.wval);
  rc |= upload(m, wdense, &m->d.wdense);
  rc |= upload(m, pick, &m->d.pick_idx);
  rc |= upload(m, ptr, &m->d.csr_ptr);
  rc |= upload(m, idx, &m->d.csr_idx);
  rc |= upload(m, val, &m->d.csr_val);
  if (!rc) {                     // operands of the tensor-core LBS (smpl_tc.cu; dense or sparse skinning weights alike)
    std::vector<unsigned char> apk, wk;
    std::vector<float> ainv;
    smpl_tc_pack(v_template, shapedirs, posedirs, lbs_weights, apk, ainv, wk);
    rc |= upload(m, apk, &m->tc_apk);
    rc |= upload(m, ainv, &m->tc_ainv);
    rc |= upload(m, wk, &m->tc_wk);
  }
  if (rc) { straps_smpl_destroy(m); return 1; }
  *out = m;
  return 0;
}

extern "C" void straps_smpl_destroy(straps_smpl_t* m) {
  if (!m) return;
  for (void* p : m->allocs) cudaFree(p);
  if (m->tc_scratch) cudaFree(m->tc_scratch);
  delete m;
}

extern "C" int straps_smpl_is_sparse4(const straps_smpl_t* m) { return m ? m->sparse4 : 0; }

template <int TB>
static int launch_lbs(const straps_smpl* m, const LbsArgs& a, cudaStream_t st) {
  dim3 grid(NTILES, ceil_div(a.B, TB));
  const size_t smem = sizeof(LbsSmem<TB>);
  if (m->sparse4) {
    STRAPS_CUDA(cudaFuncSetAttribute(lbs_kernel<TB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lbs_kernel<TB, true><<<grid, TV, smem, st>>>(m->d, a);
  } else {
    STRAPS_CUDA(cudaFuncSetAttribute(lbs_kernel<TB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lbs_kernel<TB, false><<<grid, TV, smem, st>>>(m->d, a);
  }
  STRAPS_LAUNCH_CHECK();
  return 0;
}

static int smpl_forward_impl(const straps_smpl_t* m, const float* global_orient, int64_t go_stride,
                             const float* body_pose, int64_t bp_stride, const float* betas,
                             int64_t betas_stride, const float* transl, int batch, int pose2rot,
                             float* vertices, float* joints, float* save_vposed, float* save_A, void* stream) {
  STRAPS_CHECK(m && global_orient && body_pose && betas && vertices && joints, "straps_smpl_forward: null argument");
  STRAPS_CHECK(batch > 0, "straps_smpl_forward: batch must be positive (got %d)", batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LbsArgs a;
  a.go = global_orient; a.bp = body_pose; a.betas = betas; a.transl = transl;
  a.go_stride = go_stride; a.bp_stride = bp_stride; a.betas_stride = betas_stride;
  a.B = batch; a.pose2rot = pose2rot; a.verts = vertices; a.joints = joints;
  a.save_vposed = save_vposed; a.save_A = save_A;
  // batches >= tc_min: blend shapes and skinning transforms on the tensor cores (smpl_tc.cu); STRAPS_LBS=simt keeps the CUDA-core
  // kernel, STRAPS_LBS=tc uses the tensor cores at every batch size (parity tests and tools/bench_lbs.py compare the two)
  {
    const char* e = getenv("STRAPS_LBS");
    const int tc_min = (e && e[0] == 't') ? 1 : 32;
    if (batch >= tc_min && m->tc_apk && !(e && e[0] == 's')) {
      if (smpl_tc_forward(const_cast<straps_smpl*>(m), global_orient, go_stride, body_pose, bp_stride, betas, betas_stride, transl, batch,
                          pose2rot, vertices, joints, save_vposed, save_A, st))
        return 1;
      const int warps = batch * (STRAPS_NUM_EXTRA_PICKS + STRAPS_NUM_EXTRA_ROWS);
      joints_kernel<<<ceil_div(warps * 32, 256), 256, 0, st>>>(m->d, vertices, joints, batch);
      STRAPS_LAUNCH_CHECK();
      return 0;
    }
  }
  // bodies per CTA: enough CTAs to fill 148 SMs at small batch, 8-way register tiling once the batch allows it
  int rc;
  if (batch >= 32) rc = launch_lbs<8>(m, a, st);
  else if (batch >= 12) rc = launch_lbs<4>(m, a, st);
  else if (batch >= 4) rc = launch_lbs<2>(m, a, st);
  else rc = launch_lbs<1>(m, a, st);
  if (rc) return rc;
  const int warps = batch * (STRAPS_NUM_EXTRA_PICKS + STRAPS_NUM_EXTRA_ROWS);
  joints_kernel<<<ceil_div(warps * 32, 256), 256, 0, st>>>(m->d, vertices, joints, batch);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_smpl_forward(const straps_smpl_t* m, const float* global_orient, int64_t go_stride,
                                   const float* body_pose, int64_t bp_stride, const float* betas,
                                   int64_t betas_stride, const float* transl, int batch, int pose2rot,
                                   float* vertices, float* joints, void* stream) {
  return smpl_forward_impl(m, global_orient, go_stride, body_pose, bp_stride, betas, betas_stride, transl, batch, pose2rot,
                           vertices, joints, nullptr, nullptr, stream);
}

extern "C" int straps_smpl_forward_train(const straps_smpl_t* m, const float* rotmats, const float* betas, int batch,
                                         float* vertices, float* joints, float* save_vposed, float* save_A, void* stream) {
  STRAPS_CHECK(rotmats && save_vposed && save_A, "straps_smpl_forward_train: null argument");
  return smpl_forward_impl(m, rotmats, 24 * 9, rotmats + 9, 24 * 9, betas, STRAPS_NUM_BETAS, nullptr, batch, 0, vertices, joints,
                           save_vposed, save_A, stream);
}

extern "C" int straps_rot6d_to_rotmat(const float* x6, int64_t n, float* R, void* stream) {
  STRAPS_CHECK(x6 && R, "straps_rot6d_to_rotmat: null argument");
  if (n <= 0) return 0;
  rot6d_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(x6, n, R);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_orthographic_project(const float* points, const float* cam, int64_t cam_stride, int batch,
                                           int npoints, float* out, void* stream) {
  STRAPS_CHECK(points && cam && out, "straps_orthographic_project: null argument");
  if (batch * npoints <= 0) return 0;
  ortho_kernel<<<ceil_div(batch * npoints, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(points, cam, cam_stride,
                                                                                           batch, npoints, out);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// on-device proxy-representation synthesis, the step right before the hot path
// (reference utils/label_conversions.py:48-55 and 90-127, called at train/...:178-182).
// ------------------------------------------------------------------------------------------------
namespace straps {
// one block per (b, joint): paste the (2*size)^2 truncated Gaussian window with the reference's clipping rules
__global__ void heatmaps_paste_kernel(const float* __restrict__ joints2d, const float* __restrict__ table, int size, int img_wh,
                                      float* __restrict__ out) {
  const int bj = blockIdx.x;
  const float fx = joints2d[bj * 2 + 0], fy = joints2d[bj * 2 + 1];
  const int cx = (int)fx, cy = (int)fy;                       // Tensor.int(): truncation toward zero
  if (!(cx > -size && cy > -size && cx < img_wh - 1 + size && cy < img_wh - 1 + size)) return;
  const int hsx = max(0, cx - size), hex = min(img_wh - 1, cx + size);
  const int hsy = max(0, cy - size), hey = min(img_wh - 1, cy + size);
  const int gsx = max(0, size - cx), gex = min(2 * size, 2 * size - (size + cx - (img_wh - 1)));
  const int gsy = max(0, size - cy), gey = min(2 * size, 2 * size - (size + cy - (img_wh - 1)));
  const int w = min(hex - hsx, gex - gsx), h = min(hey - hsy, gey - gsy);
  float* o = out + (size_t)bj * img_wh * img_wh;
  for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
    const int yy = i / w, xx = i % w;
    o[(size_t)(hsy + yy) * img_wh + hsx + xx] = table[(gsy + yy) * 2 * size + gsx + xx];
  }
}
__global__ void binary_labels_kernel(const float* __restrict__ in, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (in[i] != 0.f) ? 1.f : 0.f;
}
}  // namespace straps

extern "C" int straps_joints2d_to_heatmaps(const float* joints2d, int batch, int num_joints, int img_wh, int half_size,
                                           const float* table, float* heatmaps, void* stream) {
  STRAPS_CHECK(joints2d && table && heatmaps, "straps_joints2d_to_heatmaps: null argument");
  STRAPS_CHECK(batch > 0 && num_joints > 0 && img_wh > 0 && half_size > 0 && half_size <= 64, "straps_joints2d_to_heatmaps: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  STRAPS_CUDA(cudaMemsetAsync(heatmaps, 0, (size_t)batch * num_joints * img_wh * img_wh * sizeof(float), st));
  heatmaps_paste_kernel<<<batch * num_joints, 256, 0, st>>>(joints2d, table, half_size, img_wh, heatmaps);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_multiclass_to_binary(const float* labels, int64_t n, float* out, void* stream) {
  STRAPS_CHECK(labels && out, "straps_multiclass_to_binary: null argument");
  if (n <= 0) return 0;
  binary_labels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(labels, n, out);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

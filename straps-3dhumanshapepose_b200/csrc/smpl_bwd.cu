// smpl_bwd.cu -- backward of the fused SMPL forward (training path, BASELINE config 3), plus the backward of
// rot6d_to_rotmat and of the weak-perspective projection.
//
// Replaces the autograd graph torch builds for smplx.lbs.lbs + models/smpl_official.py:27-41 in the reference's
// training step (train/train_synthetic_otf_rendering.py:196-233); derivation in SURVEY.md Appendix B.
//
//   joints_bwd_kernel   g_verts += J_extra^T g_joints[45:90]  and  += g_joints[24:45] at the 21 picked vertices
//   lbs_bwd_kernel      per (128 vertices x TB bodies): T_v = sum_j w_vj A_j, g_vposed = T_v[:, :3]^T g_v,
//                       dA_j += w_vj g_v [v_posed;1]^T (shared-memory atomics -> global atomics),
//                       g_pf[k] = <g_vposed, posedirs[k]> and g_beta[l] = <g_vposed, shapedirs[l]> as warp reductions
//                       over the same TMA-streamed posedirs slab ring as the forward kernel
//   chain_bwd_kernel    one thread per body: reverse walk of the 24-joint kinematic tree, dA/g_posed_joints ->
//                       dR[24,3,3], d(rest joints) -> d_beta through the pre-reduced joint regressor, + g_pf -> dR[1:]
// fp32 throughout; reductions over vertices use atomics, so the summation order (not the value beyond ~1e-6) varies.
#include "smpl.h"

namespace straps {

__global__ void __launch_bounds__(256) joints_bwd_kernel(const SmplDev m, const float* __restrict__ g_joints,
                                                         float* __restrict__ gv, int B) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  constexpr int NOUT = STRAPS_NUM_EXTRA_PICKS + STRAPS_NUM_EXTRA_ROWS;
  if (warp >= B * NOUT) return;
  const int b = warp / NOUT, q = warp % NOUT;
  const float* g = g_joints + ((size_t)b * STRAPS_NUM_SUPERSET_JOINTS + NJ + q) * 3;
  float* gvb = gv + (size_t)b * V * 3;
  if (q < STRAPS_NUM_EXTRA_PICKS) {
    if (lane < 3) atomicAdd(gvb + (size_t)m.pick_idx[q] * 3 + lane, g[lane]);
    return;
  }
  const int r = q - STRAPS_NUM_EXTRA_PICKS;
  const float gx = g[0], gy = g[1], gz = g[2];
  for (int i = m.csr_ptr[r] + lane; i < m.csr_ptr[r + 1]; i += 32) {
    const float w = m.csr_val[i];
    float* p = gvb + (size_t)m.csr_idx[i] * 3;
    atomicAdd(p + 0, w * gx); atomicAdd(p + 1, w * gy); atomicAdd(p + 2, w * gz);
  }
}

struct LbsBwdArgs {
  const float* vposed;   // [B,6890,3] saved by the forward
  const float* A;        // [B,24,12]
  const float* gv;       // [B,6890,3] vertex gradients (joint contributions already folded in)
  float* dA;             // [B,24,12]  zero-initialised, atomically accumulated
  float* gpf;            // [B,208]    zero-initialised
  float* gbeta;          // [B,10]     zero-initialised
  int B;
};

template <int TB>
struct LbsBwdSmem {
  float pbuf[NSTAGE][KC][ROWF];
  float4 A[TB][NJ][3];
  float dA[TB][NJ][12];
  float gpf[TB][NPF_PAD];
  float gbeta[TB][STRAPS_NUM_BETAS];
  uint64_t full[NSTAGE];
  uint64_t empty[NSTAGE];
};

template <int TB, bool SPARSE>
__global__ void __launch_bounds__(TV) lbs_bwd_kernel(const SmplDev m, const LbsBwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LbsBwdSmem<TB>& s = *reinterpret_cast<LbsBwdSmem<TB>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, tile = blockIdx.x;
  const int b0 = blockIdx.y * TB, nb = min(TB, a.B - b0);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], TV / 32); }
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
  if (tid == 0) { issue_chunk(0); issue_chunk(1); }

  for (int i = tid; i < TB * NJ * 3; i += TV) {
    const int b = i / (NJ * 3), r = i % (NJ * 3);
    s.A[b][r / 3][r % 3] = (b < nb) ? *reinterpret_cast<const float4*>(a.A + ((size_t)(b0 + b) * NJ) * 12 + r * 4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = tid; i < TB * NJ * 12; i += TV) (&s.dA[0][0][0])[i] = 0.f;
  for (int i = tid; i < TB * NPF_PAD; i += TV) (&s.gpf[0][0])[i] = 0.f;
  for (int i = tid; i < TB * STRAPS_NUM_BETAS; i += TV) (&s.gbeta[0][0])[i] = 0.f;
  __syncthreads();

  const int v = tile * TV + tid;
  const bool vvalid = v < V;
  int wj[4];
  float ww[4];
  if constexpr (SPARSE) {
    const int4 ji = *reinterpret_cast<const int4*>(m.widx + v * 4);
    const float4 jw = *reinterpret_cast<const float4*>(m.wval + v * 4);
    wj[0] = ji.x; wj[1] = ji.y; wj[2] = ji.z; wj[3] = ji.w;
    ww[0] = jw.x; ww[1] = jw.y; ww[2] = jw.z; ww[3] = jw.w;
  }
  // ---- per vertex: g_vposed = T^T g, dA += w g [vp;1]^T
  float gvp[TB][3];
#pragma unroll
  for (int b = 0; b < TB; ++b) {
    float g[3] = {0.f, 0.f, 0.f}, vp[3] = {0.f, 0.f, 0.f};
    if (vvalid && b < nb) {
      const size_t o = ((size_t)(b0 + b) * V + v) * 3;
      g[0] = a.gv[o]; g[1] = a.gv[o + 1]; g[2] = a.gv[o + 2];
      vp[0] = a.vposed[o]; vp[1] = a.vposed[o + 1]; vp[2] = a.vposed[o + 2];
    }
    float T[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    const int nent = SPARSE ? 4 : NJ;
    for (int e = 0; e < nent; ++e) {
      const int j = SPARSE ? wj[e] : e;
      const float w = SPARSE ? ww[e] : m.wdense[(size_t)v * NJ + e];
      if (w == 0.f) continue;
      const float4 r0 = s.A[b][j][0], r1 = s.A[b][j][1], r2 = s.A[b][j][2];
      T[0][0] = fmaf(w, r0.x, T[0][0]); T[0][1] = fmaf(w, r0.y, T[0][1]); T[0][2] = fmaf(w, r0.z, T[0][2]);
      T[1][0] = fmaf(w, r1.x, T[1][0]); T[1][1] = fmaf(w, r1.y, T[1][1]); T[1][2] = fmaf(w, r1.z, T[1][2]);
      T[2][0] = fmaf(w, r2.x, T[2][0]); T[2][1] = fmaf(w, r2.y, T[2][1]); T[2][2] = fmaf(w, r2.z, T[2][2]);
      float* d = &s.dA[b][j][0];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float wg = w * g[r];
        if (wg != 0.f) {
          atomicAdd(d + r * 4 + 0, wg * vp[0]); atomicAdd(d + r * 4 + 1, wg * vp[1]);
          atomicAdd(d + r * 4 + 2, wg * vp[2]); atomicAdd(d + r * 4 + 3, wg);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) gvp[b][c] = T[0][c] * g[0] + T[1][c] * g[1] + T[2][c] * g[2];
  }
  // ---- g_beta[l] = sum_{v,c} g_vposed * shapedirs[l]  (v_shaped enters v_posed with coefficient 1)
  for (int l = 0; l < STRAPS_NUM_BETAS; ++l) {
    float sd[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) sd[c] = m.sdir[(size_t)l * VP3 + v * 3 + c];
#pragma unroll
    for (int b = 0; b < TB; ++b) {
      float t = gvp[b][0] * sd[0] + gvp[b][1] * sd[1] + gvp[b][2] * sd[2];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) atomicAdd(&s.gbeta[b][l], t);
    }
  }
  // ---- g_pf[k] = sum_{v,c} g_vposed * posedirs[k]
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
#pragma unroll
      for (int b = 0; b < TB; ++b) {
        float t = gvp[b][0] * p0 + gvp[b][1] * p1 + gvp[b][2] * p2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) atomicAdd(&s.gpf[b][chunk * KC + kk], t);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.empty[st]);
  }
  __syncthreads();
  for (int i = tid; i < nb * NJ * 12; i += TV) {
    const float x = (&s.dA[0][0][0])[i];
    if (x != 0.f) atomicAdd(a.dA + (size_t)b0 * NJ * 12 + i, x);
  }
  for (int i = tid; i < nb * NPF_PAD; i += TV) atomicAdd(a.gpf + (size_t)b0 * NPF_PAD + i, (&s.gpf[0][0])[i]);
  for (int i = tid; i < nb * STRAPS_NUM_BETAS; i += TV)
    atomicAdd(a.gbeta + (size_t)b0 * STRAPS_NUM_BETAS + i, (&s.gbeta[0][0])[i]);
}

// one thread per body
__global__ void chain_bwd_kernel(const SmplDev m, const float* __restrict__ R, const float* __restrict__ betas,
                                 const float* __restrict__ dA, const float* __restrict__ gpf,
                                 const float* __restrict__ g_joints, const float* __restrict__ gbeta_in, int B,
                                 float* __restrict__ dR, float* __restrict__ dbeta) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* Rb = R + (size_t)b * NJ * 9;
  float J[NJ][3], G[NJ][12], dG[NJ][12], dJ[NJ][3];
  for (int j = 0; j < NJ; ++j)
    for (int c = 0; c < 3; ++c) {
      float acc = 0.f;
      for (int l = 0; l < STRAPS_NUM_BETAS; ++l) acc = fmaf(m.js[(j * 3 + c) * STRAPS_NUM_BETAS + l], betas[(size_t)b * STRAPS_NUM_BETAS + l], acc);
      J[j][c] = m.jt[j * 3 + c] + acc;
      dJ[j][c] = 0.f;
    }
  // forward chain (G rows [R | t])
  for (int j = 0; j < NJ; ++j) {
    const int p = m.parents[j];
    const float* Rj = Rb + j * 9;
    for (int r = 0; r < 3; ++r) {
      if (p < 0) {
        G[j][r * 4 + 0] = Rj[r * 3]; G[j][r * 4 + 1] = Rj[r * 3 + 1]; G[j][r * 4 + 2] = Rj[r * 3 + 2]; G[j][r * 4 + 3] = J[j][r];
      } else {
        const float* gp = &G[p][r * 4];
        for (int c = 0; c < 3; ++c) G[j][r * 4 + c] = gp[0] * Rj[c] + gp[1] * Rj[3 + c] + gp[2] * Rj[6 + c];
        G[j][r * 4 + 3] = gp[0] * (J[j][0] - J[p][0]) + gp[1] * (J[j][1] - J[p][1]) + gp[2] * (J[j][2] - J[p][2]) + gp[3];
      }
    }
  }
  // seeds: A_j = [G.R | G.t - G.R J_j], posed_joint_j = G.t
  for (int j = 0; j < NJ; ++j) {
    const float* da = dA + ((size_t)b * NJ + j) * 12;
    const float* gj = g_joints + ((size_t)b * STRAPS_NUM_SUPERSET_JOINTS + j) * 3;
    for (int r = 0; r < 3; ++r) {
      const float dat = da[r * 4 + 3];
      for (int c = 0; c < 3; ++c) dG[j][r * 4 + c] = da[r * 4 + c] - dat * J[j][c];
      dG[j][r * 4 + 3] = dat + gj[r];
    }
    for (int c = 0; c < 3; ++c)
      dJ[j][c] -= G[j][0 * 4 + c] * da[3] + G[j][1 * 4 + c] * da[7] + G[j][2 * 4 + c] * da[11];
  }
  float* dRb = dR + (size_t)b * NJ * 9;
  for (int j = NJ - 1; j >= 0; --j) {
    const int p = m.parents[j];
    const float* Rj = Rb + j * 9;
    if (p < 0) {
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) dRb[j * 9 + r * 3 + c] = dG[j][r * 4 + c];
        dJ[j][r] += dG[j][r * 4 + 3];
      }
      continue;
    }
    float rel[3] = {J[j][0] - J[p][0], J[j][1] - J[p][1], J[j][2] - J[p][2]};
    // G_j.R = Gp.R R_j ; G_j.t = Gp.R rel + Gp.t
    for (int a2 = 0; a2 < 3; ++a2)
      for (int c = 0; c < 3; ++c) {
        float t = 0.f;
        for (int r = 0; r < 3; ++r) t += G[p][r * 4 + a2] * dG[j][r * 4 + c];   // (Gp.R^T dG_j.R)[a2][c]
        float pfg = 0.f;
        if (j >= 1) pfg = gpf[(size_t)b * NPF_PAD + (j - 1) * 9 + a2 * 3 + c];
        dRb[j * 9 + a2 * 3 + c] = t + pfg;
      }
    float drel[3];
    for (int a2 = 0; a2 < 3; ++a2) {
      float t = 0.f;
      for (int r = 0; r < 3; ++r) t += G[p][r * 4 + a2] * dG[j][r * 4 + 3];
      drel[a2] = t;
    }
    for (int r = 0; r < 3; ++r) {
      for (int a2 = 0; a2 < 3; ++a2) {
        float t = dG[j][r * 4 + 3] * rel[a2];
        for (int c = 0; c < 3; ++c) t += dG[j][r * 4 + c] * Rj[a2 * 3 + c];       // (dG_j.R R_j^T)[r][a2]
        dG[p][r * 4 + a2] += t;
      }
      dG[p][r * 4 + 3] += dG[j][r * 4 + 3];
    }
    for (int c = 0; c < 3; ++c) { dJ[j][c] += drel[c]; dJ[p][c] -= drel[c]; }
  }
  for (int l = 0; l < STRAPS_NUM_BETAS; ++l) {
    float t = gbeta_in[(size_t)b * STRAPS_NUM_BETAS + l];
    for (int j = 0; j < NJ; ++j)
      for (int c = 0; c < 3; ++c) t = fmaf(m.js[(j * 3 + c) * STRAPS_NUM_BETAS + l], dJ[j][c], t);
    dbeta[(size_t)b * STRAPS_NUM_BETAS + l] = t;
  }
}

// backward of utils/rigid_transform_utils.py:27-41
__global__ void rot6d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dR, long long n, float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = x + i * 6;
  const float* g = dR + i * 9;
  const float a1[3] = {p[0], p[2], p[4]}, a2[3] = {p[1], p[3], p[5]};
  const float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
  const float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
  const float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  const float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
  const float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
  const float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
  // R columns are (b1, b2, b3): g1 = dR[:,0], g2 = dR[:,1], g3 = dR[:,2]
  const float g1[3] = {g[0], g[3], g[6]}, g2[3] = {g[1], g[4], g[7]}, g3[3] = {g[2], g[5], g[8]};
  float db1[3], db2[3];
  // b3 = b1 x b2
  db1[0] = g1[0] + (b2[1] * g3[2] - b2[2] * g3[1]);
  db1[1] = g1[1] + (b2[2] * g3[0] - b2[0] * g3[2]);
  db1[2] = g1[2] + (b2[0] * g3[1] - b2[1] * g3[0]);
  db2[0] = g2[0] + (g3[1] * b1[2] - g3[2] * b1[1]);
  db2[1] = g2[1] + (g3[2] * b1[0] - g3[0] * b1[2]);
  db2[2] = g2[2] + (g3[0] * b1[1] - g3[1] * b1[0]);
  // b2 = u / |u|
  const float b2db2 = b2[0] * db2[0] + b2[1] * db2[1] + b2[2] * db2[2];
  const float du[3] = {(db2[0] - b2[0] * b2db2) / n2, (db2[1] - b2[1] * b2db2) / n2, (db2[2] - b2[2] * b2db2) / n2};
  // u = a2 - (b1.a2) b1
  const float b1du = b1[0] * du[0] + b1[1] * du[1] + b1[2] * du[2];
  const float da2[3] = {du[0] - b1[0] * b1du, du[1] - b1[1] * b1du, du[2] - b1[2] * b1du};
#pragma unroll
  for (int c = 0; c < 3; ++c) db1[c] += -a2[c] * b1du - d * du[c];
  // b1 = a1 / |a1|
  const float b1db1 = b1[0] * db1[0] + b1[1] * db1[1] + b1[2] * db1[2];
  const float da1[3] = {(db1[0] - b1[0] * b1db1) / n1, (db1[1] - b1[1] * b1db1) / n1, (db1[2] - b1[2] * b1db1) / n1};
  float* o = dx + i * 6;
  o[0] = da1[0]; o[2] = da1[1]; o[4] = da1[2];
  o[1] = da2[0]; o[3] = da2[1]; o[5] = da2[2];
}

// backward of utils/cam_utils.py:5-26: u = s (x + tx), v = s (y + ty)
__global__ void ortho_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ cam, long long cam_stride,
                                 const float* __restrict__ g, int B, int N, float* __restrict__ dpts, float* __restrict__ dcam) {
  const int b = blockIdx.x;
  const float s = cam[b * cam_stride + 0], tx = cam[b * cam_stride + 1], ty = cam[b * cam_stride + 2];
  float ds = 0.f, dtx = 0.f, dty = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const size_t i = (size_t)b * N + n;
    const float gu = g[i * 2], gv = g[i * 2 + 1];
    dpts[i * 3 + 0] = s * gu; dpts[i * 3 + 1] = s * gv; dpts[i * 3 + 2] = 0.f;
    ds += gu * (pts[i * 3] + tx) + gv * (pts[i * 3 + 1] + ty);
    dtx += s * gu; dty += s * gv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dtx += __shfl_xor_sync(0xffffffffu, dtx, o);
    dty += __shfl_xor_sync(0xffffffffu, dty, o);
  }
  if (threadIdx.x == 0) { dcam[b * 3 + 0] = ds; dcam[b * 3 + 1] = dtx; dcam[b * 3 + 2] = dty; }
}

}  // namespace straps

using namespace straps;

template <int TB>
static int launch_lbs_bwd(const straps_smpl* m, const LbsBwdArgs& a, cudaStream_t st) {
  dim3 grid(NTILES, ceil_div(a.B, TB));
  const size_t smem = sizeof(LbsBwdSmem<TB>);
  if (m->sparse4) {
    STRAPS_CUDA(cudaFuncSetAttribute(lbs_bwd_kernel<TB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lbs_bwd_kernel<TB, true><<<grid, TV, smem, st>>>(m->d, a);
  } else {
    STRAPS_CUDA(cudaFuncSetAttribute(lbs_bwd_kernel<TB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lbs_bwd_kernel<TB, false><<<grid, TV, smem, st>>>(m->d, a);
  }
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_smpl_backward(const straps_smpl_t* m, const float* rotmats, const float* betas,
                                    const float* v_posed, const float* A, const float* g_joints, float* gv_work,
                                    float* scratch, int batch, float* d_rotmats, float* d_betas, void* stream) {
  STRAPS_CHECK(m && rotmats && betas && v_posed && A && g_joints && gv_work && scratch && d_rotmats && d_betas,
               "straps_smpl_backward: null argument");
  STRAPS_CHECK(batch > 0, "straps_smpl_backward: batch must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nscratch = (size_t)batch * (NJ * 12 + NPF_PAD + STRAPS_NUM_BETAS);
  STRAPS_CUDA(cudaMemsetAsync(scratch, 0, nscratch * sizeof(float), st));
  const int warps = batch * (STRAPS_NUM_EXTRA_PICKS + STRAPS_NUM_EXTRA_ROWS);
  joints_bwd_kernel<<<ceil_div(warps * 32, 256), 256, 0, st>>>(m->d, g_joints, gv_work, batch);
  STRAPS_LAUNCH_CHECK();
  LbsBwdArgs a;
  a.vposed = v_posed; a.A = A; a.gv = gv_work;
  a.dA = scratch; a.gpf = scratch + (size_t)batch * NJ * 12; a.gbeta = a.gpf + (size_t)batch * NPF_PAD;
  a.B = batch;
  int rc = (batch >= 12) ? launch_lbs_bwd<4>(m, a, st) : (batch >= 4 ? launch_lbs_bwd<2>(m, a, st) : launch_lbs_bwd<1>(m, a, st));
  if (rc) return rc;
  chain_bwd_kernel<<<ceil_div(batch, 64), 64, 0, st>>>(m->d, rotmats, betas, a.dA, a.gpf, g_joints, a.gbeta, batch, d_rotmats, d_betas);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_rot6d_backward(const float* x6, const float* dR, int64_t n, float* dx6, void* stream) {
  STRAPS_CHECK(x6 && dR && dx6, "straps_rot6d_backward: null argument");
  if (n <= 0) return 0;
  rot6d_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(x6, dR, n, dx6);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_orthographic_project_backward(const float* points, const float* cam, int64_t cam_stride,
                                                    const float* g_out, int batch, int npoints, float* d_points,
                                                    float* d_cam, void* stream) {
  STRAPS_CHECK(points && cam && g_out && d_points && d_cam, "straps_orthographic_project_backward: null argument");
  if (batch * npoints <= 0) return 0;
  ortho_bwd_kernel<<<batch, 32, 0, static_cast<cudaStream_t>(stream)>>>(points, cam, cam_stride, g_out, batch, npoints, d_points, d_cam);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// ief.cu -- Iterative Error Feedback regressor as ONE launch for all iterations.
//
// Replaces IEFModule.forward (reference models/ief_module.py:48-64):
//     p0 = init.repeat(B,1);  3 x { state = [feat | p]; p += fc3(relu(fc2(relu(fc1(state))))) }
//
// B200 mapping: the fc stack is tiny (1.37 MFLOP/body/iteration, 2.74 MB of fp32 weights) and strictly
// layer-serial, so it is latency bound.  One thread-block CLUSTER of 8 CTAs owns 8 bodies: CTA `c` owns
// neurons [64c, 64c+64) of fc1/fc2 and outputs [20c, 20c+20) of fc3, streams only its slice of the
// (pre-transposed, L2-resident) weights with fully coalesced loads, and all-gathers the activations into
// every peer's shared memory through DSMEM stores + a cluster barrier.  The iteration-invariant half of
// fc1 (feat . W1[:, :512]^T + b1) is computed once and kept in shared memory.  Shared-weight layout in HBM:
//     w1t [669][512], w2t [512][512], w3t [512][160]  (input-major = transposed nn.Linear weights, fc3 padded).
#include "regressor.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace straps {

constexpr int CL = 8;          // CTAs per cluster
constexpr int TBI = 8;         // bodies per cluster
constexpr int NLOC = IEF_H / CL;            // 64 neurons per CTA
constexpr int OLOC = IEF_OUT_PAD / CL;      // 20 outputs per CTA
constexpr int IEF_THREADS = 256;
constexpr int KQ = IEF_THREADS / NLOC;      // 4 K-slices per neuron

struct IefSmem {
  float xs[IEF_IN][TBI];        // state, k-major / body-minor:  rows 0..511 feat, 512..668 params
  float h1[IEF_H][TBI];
  float h2[IEF_H][TBI];
  float base1[NLOC][TBI];       // b1 + W1[:, :512] . feat for this CTA's neurons
  float red[KQ][NLOC][TBI];     // K-slice partial sums (also used as [8][32][TBI] by fc3)
};

// acc[b] += sum_{k in [k0,k1)} wt[k*ld + col] * x[k][b].  The weight stream comes from L2 (~700 cycles away): 16
// independent loads are issued before the first FMA so one thread keeps 16 requests in flight.
__device__ __forceinline__ void dot_slice(const float* __restrict__ wt, int ld, int col, int k0, int k1,
                                          const float (*x)[TBI], float (&acc)[TBI]) {
  constexpr int U = 16;
  int k = k0;
  for (; k + U <= k1; k += U) {
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) w[u] = __ldg(wt + (size_t)(k + u) * ld + col);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 x0 = *reinterpret_cast<const float4*>(&x[k + u][0]);
      const float4 x1 = *reinterpret_cast<const float4*>(&x[k + u][4]);
      acc[0] = fmaf(w[u], x0.x, acc[0]); acc[1] = fmaf(w[u], x0.y, acc[1]);
      acc[2] = fmaf(w[u], x0.z, acc[2]); acc[3] = fmaf(w[u], x0.w, acc[3]);
      acc[4] = fmaf(w[u], x1.x, acc[4]); acc[5] = fmaf(w[u], x1.y, acc[5]);
      acc[6] = fmaf(w[u], x1.z, acc[6]); acc[7] = fmaf(w[u], x1.w, acc[7]);
    }
  }
  if (k < k1) {
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) w[u] = (k + u < k1) ? __ldg(wt + (size_t)(k + u) * ld + col) : 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (k + u < k1) {
#pragma unroll
        for (int b = 0; b < TBI; ++b) acc[b] = fmaf(w[u], x[k + u][b], acc[b]);
      }
    }
  }
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(IEF_THREADS)
ief_kernel(const float* __restrict__ feat, const float* __restrict__ init, const float* __restrict__ w1t,
           const float* __restrict__ b1, const float* __restrict__ w2t, const float* __restrict__ b2,
           const float* __restrict__ w3t, const float* __restrict__ b3, int B, int iters,
           float* __restrict__ params, float* __restrict__ saved) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  IefSmem& s = *reinterpret_cast<IefSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CL) * TBI;
  const int tid = threadIdx.x;
  const int nl = tid % NLOC, kq = tid / NLOC;

  // state: feat (transposed into k-major) + initial estimate
  for (int i = tid; i < STRAPS_FEAT_DIM * TBI; i += IEF_THREADS) {
    const int b = i / STRAPS_FEAT_DIM, k = i % STRAPS_FEAT_DIM;
    s.xs[k][b] = (b0 + b < B) ? feat[(size_t)(b0 + b) * STRAPS_FEAT_DIM + k] : 0.f;
  }
  for (int i = tid; i < STRAPS_IEF_PARAMS * TBI; i += IEF_THREADS) {
    const int k = i / TBI, b = i % TBI;
    s.xs[STRAPS_FEAT_DIM + k][b] = init[k];
  }
  __syncthreads();

  // iteration-invariant half of fc1
  {
    float acc[TBI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int kspan = STRAPS_FEAT_DIM / KQ;
    dot_slice(w1t, IEF_H, rank * NLOC + nl, kq * kspan, (kq + 1) * kspan, s.xs, acc);
#pragma unroll
    for (int b = 0; b < TBI; ++b) s.red[kq][nl][b] = acc[b];
    __syncthreads();
    for (int i = tid; i < NLOC * TBI; i += IEF_THREADS) {
      const int n = i / TBI, b = i % TBI;
      s.base1[n][b] = b1[rank * NLOC + n] + ((s.red[0][n][b] + s.red[1][n][b]) + (s.red[2][n][b] + s.red[3][n][b]));
    }
    __syncthreads();
  }

  // training: saved = [iters] x { p_k [B,157] | h1_k [B,512] | h2_k [B,512] }; CTA `rank` writes body b0+rank
  const size_t per_iter = (size_t)B * (STRAPS_IEF_PARAMS + 2 * IEF_H);
  const bool do_save = saved != nullptr && (b0 + rank < B);
  for (int it = 0; it < iters; ++it) {
    if (do_save)
      for (int k = tid; k < STRAPS_IEF_PARAMS; k += IEF_THREADS)
        saved[it * per_iter + (size_t)(b0 + rank) * STRAPS_IEF_PARAMS + k] = s.xs[STRAPS_FEAT_DIM + k][rank];
    // ---- fc1 (params half) + ReLU -> all-gather h1
    {
      float acc[TBI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int kspan = (STRAPS_IEF_PARAMS + KQ - 1) / KQ;   // 40
      const int k0 = STRAPS_FEAT_DIM + kq * kspan;
      const int k1 = min(IEF_IN, k0 + kspan);
      dot_slice(w1t, IEF_H, rank * NLOC + nl, k0, k1, s.xs, acc);
#pragma unroll
      for (int b = 0; b < TBI; ++b) s.red[kq][nl][b] = acc[b];
      __syncthreads();
      for (int i = tid; i < NLOC * TBI; i += IEF_THREADS) {
        const int n = i / TBI, b = i % TBI;
        const float v = fmaxf(s.base1[n][b] + ((s.red[0][n][b] + s.red[1][n][b]) + (s.red[2][n][b] + s.red[3][n][b])), 0.f);
        float* local = &s.h1[rank * NLOC + n][b];
#pragma unroll
        for (int peer = 0; peer < CL; ++peer) *cluster.map_shared_rank(local, peer) = v;
      }
      cluster.sync();
      if (do_save)
        for (int k = tid; k < IEF_H; k += IEF_THREADS)
          saved[it * per_iter + (size_t)B * STRAPS_IEF_PARAMS + (size_t)(b0 + rank) * IEF_H + k] = s.h1[k][rank];
    }
    // ---- fc2 + ReLU -> all-gather h2
    {
      float acc[TBI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int kspan = IEF_H / KQ;
      dot_slice(w2t, IEF_H, rank * NLOC + nl, kq * kspan, (kq + 1) * kspan, s.h1, acc);
#pragma unroll
      for (int b = 0; b < TBI; ++b) s.red[kq][nl][b] = acc[b];
      __syncthreads();
      for (int i = tid; i < NLOC * TBI; i += IEF_THREADS) {
        const int n = i / TBI, b = i % TBI;
        const float v = fmaxf(b2[rank * NLOC + n] + ((s.red[0][n][b] + s.red[1][n][b]) + (s.red[2][n][b] + s.red[3][n][b])), 0.f);
        float* local = &s.h2[rank * NLOC + n][b];
#pragma unroll
        for (int peer = 0; peer < CL; ++peer) *cluster.map_shared_rank(local, peer) = v;
      }
      cluster.sync();
      if (do_save)
        for (int k = tid; k < IEF_H; k += IEF_THREADS)
          saved[it * per_iter + (size_t)B * (STRAPS_IEF_PARAMS + IEF_H) + (size_t)(b0 + rank) * IEF_H + k] = s.h2[k][rank];
    }
    // ---- fc3: 20 outputs per CTA, 8 K-slices of 64;  p += delta -> all-gather the params rows of xs
    {
      float (*red3)[32][TBI] = reinterpret_cast<float (*)[32][TBI]>(&s.red[0][0][0]);   // [8][32][TBI] = 8 KB
      const int ol = tid % 32, ks = tid / 32;
      if (ol < OLOC) {
        float acc[TBI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        dot_slice(w3t, IEF_OUT_PAD, rank * OLOC + ol, ks * 64, (ks + 1) * 64, s.h2, acc);
#pragma unroll
        for (int b = 0; b < TBI; ++b) red3[ks][ol][b] = acc[b];
      }
      __syncthreads();
      for (int i = tid; i < OLOC * TBI; i += IEF_THREADS) {
        const int o = i / TBI, b = i % TBI, og = rank * OLOC + o;
        if (og < STRAPS_IEF_PARAMS) {
          float d = b3[og];
          float t = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) t += red3[q][o][b];
          d += t;
          float* local = &s.xs[STRAPS_FEAT_DIM + og][b];
          const float v = *local + d;
#pragma unroll
          for (int peer = 0; peer < CL; ++peer) *cluster.map_shared_rank(local, peer) = v;
        }
      }
      cluster.sync();
    }
  }
  // every CTA holds the full parameter block; CTA `rank` writes body b0+rank
  if (b0 + rank < B)
    for (int k = tid; k < STRAPS_IEF_PARAMS; k += IEF_THREADS)
      params[(size_t)(b0 + rank) * STRAPS_IEF_PARAMS + k] = s.xs[STRAPS_FEAT_DIM + k][rank];
}

// dst[k*ld_dst + n] = src[n*K + k]   (nn.Linear weight [N,K] -> input-major, zero padded to ld_dst columns)
__global__ void transpose_linear_kernel(const float* __restrict__ src, int N, int K, float* __restrict__ dst, int ld_dst) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, k = k0 + threadIdx.x;
    tile[r][threadIdx.x] = (n < N && k < K) ? src[(size_t)n * K + k] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, n = n0 + threadIdx.x;
    if (k < K && n < ld_dst) dst[(size_t)k * ld_dst + n] = tile[threadIdx.x][r];
  }
}

int ief_pack(straps_regressor* r, const float* const* fc_w, const float* const* fc_b, const float* init, cudaStream_t st) {
  dim3 blk(32, 8);
  transpose_linear_kernel<<<dim3(ceil_div(IEF_IN, 32), ceil_div(IEF_H, 32)), blk, 0, st>>>(fc_w[0], IEF_H, IEF_IN, r->w1t, IEF_H);
  STRAPS_LAUNCH_CHECK();
  transpose_linear_kernel<<<dim3(ceil_div(IEF_H, 32), ceil_div(IEF_H, 32)), blk, 0, st>>>(fc_w[1], IEF_H, IEF_H, r->w2t, IEF_H);
  STRAPS_LAUNCH_CHECK();
  transpose_linear_kernel<<<dim3(ceil_div(IEF_H, 32), ceil_div(IEF_OUT_PAD, 32)), blk, 0, st>>>(fc_w[2], STRAPS_IEF_PARAMS, IEF_H, r->w3t, IEF_OUT_PAD);
  STRAPS_LAUNCH_CHECK();
  STRAPS_CUDA(cudaMemcpyAsync(r->b1, fc_b[0], IEF_H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  STRAPS_CUDA(cudaMemcpyAsync(r->b2, fc_b[1], IEF_H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  STRAPS_CUDA(cudaMemsetAsync(r->b3, 0, IEF_OUT_PAD * sizeof(float), st));
  STRAPS_CUDA(cudaMemcpyAsync(r->b3, fc_b[2], STRAPS_IEF_PARAMS * sizeof(float), cudaMemcpyDeviceToDevice, st));
  STRAPS_CUDA(cudaMemcpyAsync(r->init, init, STRAPS_IEF_PARAMS * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int ief_launch_train(const straps_regressor* r, const float* feat, int batch, int iters, float* params, float* saved, cudaStream_t st) {
  const int nclusters = ceil_div(batch, TBI);
  const size_t smem = sizeof(IefSmem);
  STRAPS_CUDA(cudaFuncSetAttribute(ief_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ief_kernel<<<nclusters * CL, IEF_THREADS, smem, st>>>(feat, r->init, r->w1t, r->b1, r->w2t, r->b2, r->w3t, r->b3,
                                                       batch, iters, params, saved);
  STRAPS_LAUNCH_CHECK();
  return 0;
}
int ief_launch(const straps_regressor* r, const float* feat, int batch, int iters, float* params, cudaStream_t st) {
  return ief_launch_train(r, feat, batch, iters, params, nullptr, st);
}

}  // namespace straps

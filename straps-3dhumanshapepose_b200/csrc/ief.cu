// ief.cu -- Iterative Error Feedback regressor as ONE launch for all iterations.
//
// Replaces IEFModule.forward (reference models/ief_module.py:48-64):
//     p0 = init.repeat(B,1);  3 x { state = [feat | p]; p += fc3(relu(fc2(relu(fc1(state))))) }
//
// B200 mapping: the fc stack is tiny (1.37 MFLOP/body/iteration, 2.74 MB of fp32 weights) and strictly layer-serial, so it
// is latency bound (ncu, profiles/r01_small_kernels_ncu.txt: the first version streamed its weight slices from L2 in
// every iteration and spent its time in long-scoreboard stalls).  One thread-block CLUSTER of 16 CTAs (non-portable
// size, one cluster per GPC) owns 8 bodies: CTA c owns neurons [32c, 32c+32) of fc1/fc2 and outputs [10c, 10c+10) of fc3.
// Its weight slices (fc1 params-half 20 KB, fc2 64 KB, fc3 32 KB) are bulk-copied (TMA) into shared memory ONCE and stay
// resident for all iterations; the iteration-invariant half of fc1 (feat . W1[:, :512]^T + b1) is computed once from a slice
// that is streamed through the fc2 region before fc2's weights land there.  Activations are all-gathered into every peer's
// shared memory with DSMEM stores + a cluster barrier (3 per iteration).
// HBM layout (packed by ief_pack): per-CTA contiguous slices  w1f [16][512][32], w1p [16][157][32], w2 [16][512][32],
// w3 [16][512][16] (10 real outputs + 6 zero columns), input-major so that a warp reads 32 consecutive floats.
#include "regressor.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace straps {

constexpr int CL = 16;         // CTAs per cluster
constexpr int TBI = 8;         // bodies per cluster
constexpr int NLOC = IEF_H / CL;            // 32 neurons per CTA
constexpr int OLOC = IEF_OUT_PAD / CL;      // 10 outputs per CTA
constexpr int OPAD = 16;                    // padded fc3 slice width
constexpr int IEF_THREADS = 256;
constexpr int KQ = IEF_THREADS / NLOC;      // 8 K-slices per neuron (one per warp)

struct IefWeights {       // owned by ief.cu
  float *w1f, *w1p, *w2, *w3;
};

struct IefSmem {
  float xs[IEF_IN][TBI];        // state, k-major / body-minor:  rows 0..511 feat, 512..668 params
  float h1[IEF_H][TBI];
  float h2[IEF_H][TBI];
  float base1[NLOC][TBI];       // b1 + W1[:, :512] . feat for this CTA's neurons
  float red[KQ][NLOC][TBI];     // K-slice partial sums (fc3 views it as [16][16][TBI])
  float w1p[STRAPS_IEF_PARAMS][NLOC];
  float w2[IEF_H][NLOC];        // holds the w1f slice first
  float w3[IEF_H][OPAD];
  uint64_t bar[3];
};

// acc[b] += sum_{k in [k0,k1)} w[k][col] * x[k][b]   (w in shared memory, row stride LD)
template <int LD>
__device__ __forceinline__ void dot_smem(const float* __restrict__ w, int col, int k0, int k1, const float (*x)[TBI], float (&acc)[TBI]) {
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    const float wv = w[k * LD + col];
    const float4 x0 = *reinterpret_cast<const float4*>(&x[k][0]);
    const float4 x1 = *reinterpret_cast<const float4*>(&x[k][4]);
    acc[0] = fmaf(wv, x0.x, acc[0]); acc[1] = fmaf(wv, x0.y, acc[1]);
    acc[2] = fmaf(wv, x0.z, acc[2]); acc[3] = fmaf(wv, x0.w, acc[3]);
    acc[4] = fmaf(wv, x1.x, acc[4]); acc[5] = fmaf(wv, x1.y, acc[5]);
    acc[6] = fmaf(wv, x1.z, acc[6]); acc[7] = fmaf(wv, x1.w, acc[7]);
  }
}

__global__ void __launch_bounds__(IEF_THREADS)
ief_kernel(const float* __restrict__ feat, const float* __restrict__ init, const IefWeights wts, const float* __restrict__ b1,
           const float* __restrict__ b2, const float* __restrict__ b3, int B, int iters, float* __restrict__ params,
           float* __restrict__ saved) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  IefSmem& s = *reinterpret_cast<IefSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CL) * TBI;
  const int tid = threadIdx.x;
  const int nl = tid % NLOC, kq = tid / NLOC;

  if (tid == 0) {
    mbar_init(&s.bar[0], 1); mbar_init(&s.bar[1], 1); mbar_init(&s.bar[2], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&s.bar[0], IEF_H * NLOC * 4);
    bulk_g2s(&s.w2[0][0], wts.w1f + (size_t)rank * IEF_H * NLOC, IEF_H * NLOC * 4, &s.bar[0]);
    mbar_arrive_expect_tx(&s.bar[1], STRAPS_IEF_PARAMS * NLOC * 4);
    bulk_g2s(&s.w1p[0][0], wts.w1p + (size_t)rank * STRAPS_IEF_PARAMS * NLOC, STRAPS_IEF_PARAMS * NLOC * 4, &s.bar[1]);
    mbar_arrive_expect_tx(&s.bar[2], IEF_H * OPAD * 4);
    bulk_g2s(&s.w3[0][0], wts.w3 + (size_t)rank * IEF_H * OPAD, IEF_H * OPAD * 4, &s.bar[2]);
  }
  // state: feat (transposed into k-major) + initial estimate
  {
    // 16 independent loads per thread are issued before the first store (ncu: the naive loop serialised 16 L2 round trips)
    constexpr int PER = STRAPS_FEAT_DIM * TBI / IEF_THREADS;
    float fv[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = tid + j * IEF_THREADS, b = i / STRAPS_FEAT_DIM, k = i % STRAPS_FEAT_DIM;
      fv[j] = (b0 + b < B) ? __ldg(feat + (size_t)(b0 + b) * STRAPS_FEAT_DIM + k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = tid + j * IEF_THREADS;
      s.xs[i % STRAPS_FEAT_DIM][i / STRAPS_FEAT_DIM] = fv[j];
    }
  }
  for (int i = tid; i < STRAPS_IEF_PARAMS * TBI; i += IEF_THREADS) {
    const int k = i / TBI, b = i % TBI;
    s.xs[STRAPS_FEAT_DIM + k][b] = init[k];
  }
  __syncthreads();

  // iteration-invariant half of fc1 (its weight slice sits in the fc2 region for now)
  mbar_wait(&s.bar[0], 0);
  {
    float acc[TBI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    constexpr int span = STRAPS_FEAT_DIM / KQ;
    dot_smem<NLOC>(&s.w2[0][0], nl, kq * span, (kq + 1) * span, s.xs, acc);
#pragma unroll
    for (int b = 0; b < TBI; ++b) s.red[kq][nl][b] = acc[b];
    __syncthreads();
    {
      const int n = tid / TBI, b = tid % TBI;    // 256 threads = 32 neurons x 8 bodies
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < KQ; ++q) t += s.red[q][n][b];
      s.base1[n][b] = b1[rank * NLOC + n] + t;
    }
    __syncthreads();
  }
  if (tid == 0) {   // the fc2 region is free now: stream fc2's slice into it
    fence_proxy_async();
    mbar_arrive_expect_tx(&s.bar[0], IEF_H * NLOC * 4);
    bulk_g2s(&s.w2[0][0], wts.w2 + (size_t)rank * IEF_H * NLOC, IEF_H * NLOC * 4, &s.bar[0]);
  }
  mbar_wait(&s.bar[1], 0);

  // training: saved = [iters] x { p_k [B,157] | h1_k [B,512] | h2_k [B,512] }; CTA `rank` < 8 writes body b0+rank
  const size_t per_iter = (size_t)B * (STRAPS_IEF_PARAMS + 2 * IEF_H);
  const bool owner = rank < TBI && (b0 + rank < B);
  const bool do_save = saved != nullptr && owner;
  for (int it = 0; it < iters; ++it) {
    if (do_save)
      for (int k = tid; k < STRAPS_IEF_PARAMS; k += IEF_THREADS)
        saved[it * per_iter + (size_t)(b0 + rank) * STRAPS_IEF_PARAMS + k] = s.xs[STRAPS_FEAT_DIM + k][rank];
    // ---- fc1 (params half) + ReLU -> all-gather h1
    {
      float acc[TBI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      constexpr int span = (STRAPS_IEF_PARAMS + KQ - 1) / KQ;   // 20
      const int k0 = kq * span, k1 = min(STRAPS_IEF_PARAMS, k0 + span);
      dot_smem<NLOC>(&s.w1p[0][0], nl, k0, k1, s.xs + STRAPS_FEAT_DIM, acc);
#pragma unroll
      for (int b = 0; b < TBI; ++b) s.red[kq][nl][b] = acc[b];
      __syncthreads();
      {
        const int n = tid / TBI, b = tid % TBI;
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < KQ; ++q) t += s.red[q][n][b];
        const float v = fmaxf(s.base1[n][b] + t, 0.f);
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
    if (it == 0) mbar_wait(&s.bar[0], 1);
    {
      float acc[TBI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      constexpr int span = IEF_H / KQ;
      dot_smem<NLOC>(&s.w2[0][0], nl, kq * span, (kq + 1) * span, s.h1, acc);
#pragma unroll
      for (int b = 0; b < TBI; ++b) s.red[kq][nl][b] = acc[b];
      __syncthreads();
      {
        const int n = tid / TBI, b = tid % TBI;
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < KQ; ++q) t += s.red[q][n][b];
        const float v = fmaxf(b2[rank * NLOC + n] + t, 0.f);
        float* local = &s.h2[rank * NLOC + n][b];
#pragma unroll
        for (int peer = 0; peer < CL; ++peer) *cluster.map_shared_rank(local, peer) = v;
      }
      cluster.sync();
      if (do_save)
        for (int k = tid; k < IEF_H; k += IEF_THREADS)
          saved[it * per_iter + (size_t)B * (STRAPS_IEF_PARAMS + IEF_H) + (size_t)(b0 + rank) * IEF_H + k] = s.h2[k][rank];
    }
    // ---- fc3: 10 (padded 16) outputs per CTA, 16 K-slices of 32;  p += delta -> all-gather the params rows of xs
    if (it == 0) mbar_wait(&s.bar[2], 0);
    {
      float (*red3)[OPAD][TBI] = reinterpret_cast<float (*)[OPAD][TBI]>(&s.red[0][0][0]);   // [16][16][TBI]
      const int ol = tid % OPAD, ks = tid / OPAD;
      float acc[TBI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      dot_smem<OPAD>(&s.w3[0][0], ol, ks * 32, (ks + 1) * 32, s.h2, acc);
#pragma unroll
      for (int b = 0; b < TBI; ++b) red3[ks][ol][b] = acc[b];
      __syncthreads();
      if (tid < OLOC * TBI) {
        const int o = tid / TBI, b = tid % TBI, og = rank * OLOC + o;
        if (og < STRAPS_IEF_PARAMS) {
          float t = 0.f;
#pragma unroll
          for (int q = 0; q < 16; ++q) t += red3[q][o][b];
          float* local = &s.xs[STRAPS_FEAT_DIM + og][b];
          const float v = *local + (b3[og] + t);
#pragma unroll
          for (int peer = 0; peer < CL; ++peer) *cluster.map_shared_rank(local, peer) = v;
        }
      }
      cluster.sync();
    }
  }
  // every CTA holds the full parameter block; CTA `rank` < 8 writes body b0+rank
  if (owner)
    for (int k = tid; k < STRAPS_IEF_PARAMS; k += IEF_THREADS)
      params[(size_t)(b0 + rank) * STRAPS_IEF_PARAMS + k] = s.xs[STRAPS_FEAT_DIM + k][rank];
}

// dst[c][k][j] = (j < nloc && c*nloc+j < N) ? W[(c*nloc + j)*K + k0 + k] : 0      (W = nn.Linear weight [N,K])
__global__ void pack_slices_kernel(const float* __restrict__ W, int N, int K, int k0, int kcount, int nloc, int npad,
                                   float* __restrict__ dst) {
  const int total = CL * kcount * npad;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = i % npad;
  int t = i / npad;
  const int k = t % kcount, c = t / kcount;
  const int n = c * nloc + j;
  dst[i] = (j < nloc && n < N) ? W[(size_t)n * K + k0 + k] : 0.f;
}

static IefWeights* ief_state(const straps_regressor* r) { return static_cast<IefWeights*>(r->ief); }

int ief_create(straps_regressor* r) {
  IefWeights* w = new IefWeights();
  r->ief = w;
  const size_t n = (size_t)CL * (IEF_H * NLOC + STRAPS_IEF_PARAMS * NLOC + IEF_H * NLOC + IEF_H * OPAD);
  float* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(float)) != cudaSuccess) {
    set_error("ief_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    w->w1f = nullptr;
    return 1;
  }
  w->w1f = p; p += (size_t)CL * IEF_H * NLOC;
  w->w1p = p; p += (size_t)CL * STRAPS_IEF_PARAMS * NLOC;
  w->w2 = p; p += (size_t)CL * IEF_H * NLOC;
  w->w3 = p;
  return 0;
}

void ief_destroy(straps_regressor* r) {
  IefWeights* w = ief_state(r);
  if (!w) return;
  if (w->w1f) cudaFree(w->w1f);
  delete w;
  r->ief = nullptr;
}

int ief_pack(straps_regressor* r, const float* const* fc_w, const float* const* fc_b, const float* init, cudaStream_t st) {
  IefWeights* w = ief_state(r);
  auto pack = [&](const float* W, int N, int K, int k0, int kcount, int nloc, int npad, float* dst) {
    const int total = CL * kcount * npad;
    pack_slices_kernel<<<ceil_div(total, 256), 256, 0, st>>>(W, N, K, k0, kcount, nloc, npad, dst);
  };
  pack(fc_w[0], IEF_H, IEF_IN, 0, STRAPS_FEAT_DIM, NLOC, NLOC, w->w1f);
  STRAPS_LAUNCH_CHECK();
  pack(fc_w[0], IEF_H, IEF_IN, STRAPS_FEAT_DIM, STRAPS_IEF_PARAMS, NLOC, NLOC, w->w1p);
  STRAPS_LAUNCH_CHECK();
  pack(fc_w[1], IEF_H, IEF_H, 0, IEF_H, NLOC, NLOC, w->w2);
  STRAPS_LAUNCH_CHECK();
  pack(fc_w[2], STRAPS_IEF_PARAMS, IEF_H, 0, IEF_H, OLOC, OPAD, w->w3);
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
  static PerDeviceOnce attr_once;
  const int attr_dev = current_device();
  if (attr_once.need(attr_dev)) {
    STRAPS_CUDA(cudaFuncSetAttribute(ief_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    STRAPS_CUDA(cudaFuncSetAttribute(ief_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_once.done(attr_dev);
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(nclusters * CL);
  cfg.blockDim = dim3(IEF_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const IefWeights wts = *ief_state(r);
  const float* b1 = r->b1; const float* b2 = r->b2; const float* b3 = r->b3; const float* init = r->init;
  STRAPS_CUDA(cudaLaunchKernelEx(&cfg, ief_kernel, feat, init, wts, b1, b2, b3, batch, iters, params, saved));
  count_launch();
  return 0;
}

int ief_launch(const straps_regressor* r, const float* feat, int batch, int iters, float* params, cudaStream_t st) {
  return ief_launch_train(r, feat, batch, iters, params, nullptr, st);
}

}  // namespace straps

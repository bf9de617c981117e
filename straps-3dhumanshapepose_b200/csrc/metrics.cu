// metrics.cu -- SURVEY.md 8f row N4: the evaluation metrics of the training / validation loop on the device.
//
// Replaces the per-sample numpy loops of the reference and the four [B,6890,3] device->host copies that feed them:
//   utils/eval_utils.py:7-52   compute_similarity_transform (orthogonal Procrustes: 3x3 SVD, det fix, scale, translation)
//   utils/eval_utils.py:55-60  procrustes_analysis_batch
//   utils/eval_utils.py:63-85  scale_and_translation_transform_batch
//   metrics/train_loss_and_metrics_tracker.py:102-213  PVE / PVE-SC / PVE-PA / PVE-T* / MPJPE* sums, pose / shape MSE sums,
//                                                     2-D joint L2 error sum (utils/joints2d_utils.py:5-10 un-normalisation)
// points_metrics_kernel: one CTA per body, three passes over the body's two point sets (2 x 82 KB for 6890 vertices: the first
//   pass pulls them from HBM, the other two hit L1/L2): means -> centred moments (variances + 3x3 cross-covariance) -> one thread
//   solves the 3x3 problem in fp64 (Jacobi eigen-decomposition of K^T K) -> error sums, optionally the aligned points.
//   Sums are accumulated in fp64 and added with fp64 atomics to a device-resident accumulator, so a training loop never
//   synchronises on its metrics.  HBM-bound by contract: 2 * N * 12 bytes per body read once (+ 2 * N * 12 written when the
//   aligned points are requested).
#include "common.cuh"
#include "../../include/straps_b200.h"
#include "procrustes.h"

namespace straps {

constexpr int MT_THREADS = 256;

template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red /*[MT_THREADS/32][NV]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
  __syncthreads();   // protects `red` from the previous use
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp * NV + i] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < MT_THREADS / 32; ++w) s += red[w * NV + i];
    v[i] = s;          // every thread gets the total
  }
}

struct AlignSmem {
  double red[(MT_THREADS / 32) * 11];
  float mu1[3], mu2[3];
  float sc_div, sc_mul;     // scale-and-translation correction: ((p - mu1) / sc_div) * sc_mul + mu2
  float sR[9], t[3];        // Procrustes: sR p + t
};

// which: bit 0 = plain error, bit 1 = scale+translation corrected, bit 2 = Procrustes aligned
__global__ void __launch_bounds__(MT_THREADS)
points_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, int N, int which,
                      double* __restrict__ sums, float* __restrict__ out_sc, float* __restrict__ out_pa) {
  __shared__ AlignSmem s;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* P = pred + (size_t)b * N * 3;
  const float* T = target + (size_t)b * N * 3;
  const bool need_align = (which & 6) != 0;

  if (need_align) {
    // pass 1: means
    double m[6] = {0, 0, 0, 0, 0, 0};
    for (int i = tid; i < N; i += MT_THREADS) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { m[c] += P[i * 3 + c]; m[3 + c] += T[i * 3 + c]; }
    }
    block_sum<6>(m, s.red);
    const float mu1[3] = {(float)(m[0] / N), (float)(m[1] / N), (float)(m[2] / N)};
    const float mu2[3] = {(float)(m[3] / N), (float)(m[4] / N), (float)(m[5] / N)};
    // pass 2: centred moments  v[0] = sum |x1|^2, v[1] = sum |x2|^2, v[2 + 3a + c] = sum x1_a x2_c
    double v[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < N; i += MT_THREADS) {
      float x1[3], x2[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) { x1[c] = P[i * 3 + c] - mu1[c]; x2[c] = T[i * 3 + c] - mu2[c]; }
      v[0] += (double)(x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2]);
      v[1] += (double)(x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2]);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) v[2 + a * 3 + c] += (double)(x1[a] * x2[c]);
    }
    block_sum<11>(v, s.red);
    if (tid == 0) {
      for (int c = 0; c < 3; ++c) { s.mu1[c] = mu1[c]; s.mu2[c] = mu2[c]; }
      s.sc_div = (float)sqrt(v[0] / N);          // RMS distance from the mean
      s.sc_mul = (float)sqrt(v[1] / N);
      if (which & 4) {
        double K[3][3], R[3][3];
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 3; ++c) K[a][c] = v[2 + a * 3 + c];
        procrustes_rotation(K, R);
        double tr = 0.0;
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 3; ++c) tr += R[a][c] * K[c][a];
        const double scale = tr / v[0];
        for (int a = 0; a < 3; ++a) {
          double rm = 0.0;
          for (int c = 0; c < 3; ++c) { s.sR[a * 3 + c] = (float)(scale * R[a][c]); rm += R[a][c] * (m[c] / N); }
          s.t[a] = (float)(m[3 + a] / N - scale * rm);
        }
      }
    }
    __syncthreads();
  }

  // pass 3: error sums (and the aligned points when asked for)
  double e[3] = {0, 0, 0};
  for (int i = tid; i < N; i += MT_THREADS) {
    const float p0 = P[i * 3], p1 = P[i * 3 + 1], p2 = P[i * 3 + 2];
    const float t0 = T[i * 3], t1 = T[i * 3 + 1], t2 = T[i * 3 + 2];
    if (which & 1) {
      const float d0 = p0 - t0, d1 = p1 - t1, d2 = p2 - t2;
      e[0] += (double)sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    }
    if (which & 2) {
      const float q0 = (p0 - s.mu1[0]) / s.sc_div * s.sc_mul + s.mu2[0];
      const float q1 = (p1 - s.mu1[1]) / s.sc_div * s.sc_mul + s.mu2[1];
      const float q2 = (p2 - s.mu1[2]) / s.sc_div * s.sc_mul + s.mu2[2];
      if (out_sc) {
        float* o = out_sc + ((size_t)b * N + i) * 3;
        o[0] = q0; o[1] = q1; o[2] = q2;
      }
      const float d0 = q0 - t0, d1 = q1 - t1, d2 = q2 - t2;
      e[1] += (double)sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    }
    if (which & 4) {
      const float q0 = s.sR[0] * p0 + s.sR[1] * p1 + s.sR[2] * p2 + s.t[0];
      const float q1 = s.sR[3] * p0 + s.sR[4] * p1 + s.sR[5] * p2 + s.t[1];
      const float q2 = s.sR[6] * p0 + s.sR[7] * p1 + s.sR[8] * p2 + s.t[2];
      if (out_pa) {
        float* o = out_pa + ((size_t)b * N + i) * 3;
        o[0] = q0; o[1] = q1; o[2] = q2;
      }
      const float d0 = q0 - t0, d1 = q1 - t1, d2 = q2 - t2;
      e[2] += (double)sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    }
  }
  if (sums) {
    block_sum<3>(e, s.red);
    if (tid == 0) {
      if (which & 1) atomicAdd(&sums[0], e[0]);
      if (which & 2) atomicAdd(&sums[1], e[1]);
      if (which & 4) atomicAdd(&sums[2], e[2]);
    }
  }
}

// sum over rows of ||(pred + pred_add) * pred_mul - target||_2 (dim <= 4), or -- squared != 0 -- of the squared differences
__global__ void __launch_bounds__(MT_THREADS)
rows_metric_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long rows, int dim, float pred_add,
                   float pred_mul, int squared, double* __restrict__ sum) {
  __shared__ double red[MT_THREADS / 32];
  double acc[1] = {0.0};
  for (long long r = (long long)blockIdx.x * MT_THREADS + threadIdx.x; r < rows; r += (long long)gridDim.x * MT_THREADS) {
    float ss = 0.f;
    for (int c = 0; c < dim; ++c) {
      const float d = (pred[r * dim + c] + pred_add) * pred_mul - target[r * dim + c];
      ss += d * d;
    }
    acc[0] += (double)(squared ? ss : sqrtf(ss));
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) atomicAdd(sum, acc[0]);
}

__global__ void accumulate_kernel(const float* __restrict__ src, int n, double scale, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += scale * (double)src[i];
}

}  // namespace straps

using namespace straps;

extern "C" int straps_points_metrics(const float* pred, const float* target, int batch, int npoints, int which, double* sums,
                                     float* pred_sc, float* pred_pa, void* stream) {
  STRAPS_CHECK(batch >= 0 && npoints > 0 && which > 0 && which < 8, "straps_points_metrics: bad sizes / selector");
  STRAPS_CHECK(sums || pred_sc || pred_pa, "straps_points_metrics: no output requested");
  STRAPS_CHECK((!pred_sc || (which & 2)) && (!pred_pa || (which & 4)), "straps_points_metrics: aligned output without its selector bit");
  if (batch == 0) return 0;       // an empty batch has null data pointers
  STRAPS_CHECK(pred && target, "straps_points_metrics: null argument");
  points_metrics_kernel<<<batch, MT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(pred, target, npoints, which, sums, pred_sc,
                                                                                   pred_pa);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_rows_metric(const float* pred, const float* target, int64_t rows, int dim, float pred_add, float pred_mul,
                                  int squared, double* sum, void* stream) {
  STRAPS_CHECK(rows >= 0 && dim > 0 && sum, "straps_rows_metric: bad sizes / null sum");
  if (rows == 0) return 0;
  STRAPS_CHECK(pred && target, "straps_rows_metric: null argument");
  const long long blocks = (rows + MT_THREADS - 1) / MT_THREADS;
  rows_metric_kernel<<<(unsigned)(blocks < 592 ? blocks : 592), MT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, target, rows, dim, pred_add, pred_mul, squared, sum);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_accumulate(const float* src, int n, double scale, double* dst, void* stream) {
  STRAPS_CHECK(src && dst && n >= 0, "straps_accumulate: bad argument");
  if (n == 0) return 0;
  accumulate_kernel<<<ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(src, n, scale, dst);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// loss.cu -- fused homoscedastic-uncertainty multi-task loss, forward + backward seeds.
//
// Replaces HomoscedasticUncertaintyWeightedMultiTaskLoss.forward (reference losses/multi_task_loss.py:73-119) together with
// check_joints2d_visibility_torch (utils/joints2d_utils.py:23-33, folded in as the row mask) and the autograd graph of its
// five nn.MSELoss terms:    total = sum_t  MSE_t * exp(-s_t) + s_t,   t in {verts, joints2D, joints3D, shape, pose}
// with joints2D restricted to visible rows and its label mapped to [-1, 1] (2*label/256 - 1).
//
// HBM-bound: one pass over predictions + targets (2 x 5.3 MB of vertices at B=64 dominate) for the five sums, a one-thread
// finalise, and one elementwise pass that writes d(total)/d(prediction).  Sums are accumulated in fp64 (block partials in fp32).
#include "common.cuh"
#include "../../include/straps_b200.h"

namespace straps {

struct LossTask {
  const float* pred;
  const float* target;
  float* grad;          // d total / d pred (backward), same shape as pred
  long long n;          // elements
};

struct LossArgs {
  LossTask t[5];        // 0 verts, 1 joints2D [rows,2], 2 joints3D, 3 shape, 4 pose rotmats
  const unsigned char* vis;   // optional bool mask per joints2D row (1 = visible); null = compute from the label as the reference does
  int use_vis;          // 0 = all rows ('vis' key absent in labels)
  float img_wh;
  int on[5];
  int sum_reduction;    // 1 = reduction='sum'
};

__device__ __forceinline__ bool row_visible(const LossArgs& a, long long row) {
  if (!a.use_vis) return true;
  if (a.vis) return a.vis[row] != 0;
  const float x = a.t[1].target[row * 2], y = a.t[1].target[row * 2 + 1];
  return !(x > a.img_wh || y > a.img_wh || x < 0.f || y < 0.f);
}
__device__ __forceinline__ float j2d_label(const LossArgs& a, long long i) { return (2.0f * a.t[1].target[i]) / a.img_wh - 1.0f; }

// acc[0..4] = sum of squared differences, acc[5] = number of visible joints2D rows
__global__ void __launch_bounds__(256) loss_reduce_kernel(const LossArgs a, double* __restrict__ acc) {
  __shared__ float red[6][8];
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * blockDim.x, tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    if (!a.on[t]) continue;
    if (t == 1) {
      for (long long r = tid0; r < a.t[1].n / 2; r += stride) {
        if (!row_visible(a, r)) continue;
        const float d0 = a.t[1].pred[r * 2] - j2d_label(a, r * 2), d1 = a.t[1].pred[r * 2 + 1] - j2d_label(a, r * 2 + 1);
        s[1] += d0 * d0 + d1 * d1;
        s[5] += 1.f;
      }
    } else {
      for (long long i = tid0; i < a.t[t].n; i += stride) {
        const float d = a.t[t].pred[i] - a.t[t].target[i];
        s[t] += d * d;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(&acc[threadIdx.x], (double)t);
  }
}

// out[0] = total, out[1..5] = MSE_t * exp(-s_t) (0 for tasks that are off); coef[t] = 2 exp(-s_t) / N_t ; dlv[t] = d total / d s_t
__global__ void loss_finalize_kernel(const LossArgs a, const double* __restrict__ acc, const float* __restrict__ log_vars,
                                     float* __restrict__ out, float* __restrict__ coef, float* __restrict__ dlv) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float total = 0.f;
  for (int t = 0; t < 5; ++t) {
    out[1 + t] = 0.f; coef[t] = 0.f; dlv[t] = 0.f;
    if (!a.on[t]) continue;
    double n = (t == 1) ? 2.0 * acc[5] : (double)a.t[t].n;
    if (a.sum_reduction) n = 1.0;
    const float mse = (float)(acc[t] / n);       // empty joints2D selection -> 0/0 = NaN, like nn.MSELoss on an empty tensor
    const float s = log_vars[t], w = expf(-s);
    const float weighted = mse * w;
    out[1 + t] = weighted;
    total += weighted + s;
    coef[t] = (float)(2.0 / n) * w;
    dlv[t] = 1.f - weighted;
  }
  out[0] = total;
}

__global__ void __launch_bounds__(256) loss_backward_kernel(const LossArgs a, const float* __restrict__ coef, const float* __restrict__ g_total) {
  const float g = g_total ? g_total[0] : 1.f;
  const long long stride = (long long)gridDim.x * blockDim.x, tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    if (!a.on[t] || !a.t[t].grad) continue;
    const float c = coef[t] * g;
    if (t == 1) {
      for (long long r = tid0; r < a.t[1].n / 2; r += stride) {
        const bool v = row_visible(a, r);
        a.t[1].grad[r * 2] = v ? c * (a.t[1].pred[r * 2] - j2d_label(a, r * 2)) : 0.f;
        a.t[1].grad[r * 2 + 1] = v ? c * (a.t[1].pred[r * 2 + 1] - j2d_label(a, r * 2 + 1)) : 0.f;
      }
    } else {
      for (long long i = tid0; i < a.t[t].n; i += stride) a.t[t].grad[i] = c * (a.t[t].pred[i] - a.t[t].target[i]);
    }
  }
}

}  // namespace straps

using namespace straps;

// preds/targets/grads: 5 device pointers each in the order verts, joints2D, joints3D, shape_params, pose rotmats (null = task off;
// grads[t] may be null when no gradient is wanted).  counts[5] = element counts.  vis: optional dev uint8 [rows] mask, use_vis = 0
// disables the visibility selection.  log_vars dev [5].  scratch dev: 8 doubles + 16 floats (96 bytes).  Outputs:
// out dev [6] = (total, five weighted task losses), d_log_vars dev [5].  g_total dev [1] or null (= 1).
extern "C" int straps_multitask_loss(const float* const* preds, const float* const* targets, float* const* grads, const int64_t* counts,
                                     const unsigned char* vis, int use_vis, float img_wh, const float* log_vars, int sum_reduction,
                                     const float* g_total, void* scratch, float* out, float* d_log_vars, void* stream) {
  STRAPS_CHECK(preds && targets && grads && counts && log_vars && scratch && out && d_log_vars, "straps_multitask_loss: null argument");
  LossArgs a;
  memset(&a, 0, sizeof(a));
  long long maxn = 1;
  for (int t = 0; t < 5; ++t) {
    a.on[t] = preds[t] != nullptr;
    if (a.on[t]) STRAPS_CHECK(targets[t] && counts[t] > 0, "straps_multitask_loss: task %d has no target / elements", t);
    a.t[t].pred = preds[t]; a.t[t].target = targets[t]; a.t[t].grad = grads[t]; a.t[t].n = counts[t];
    if (counts[t] > maxn) maxn = counts[t];
  }
  a.vis = vis; a.use_vis = use_vis; a.img_wh = img_wh; a.sum_reduction = sum_reduction;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* acc = static_cast<double*>(scratch);
  float* coef = reinterpret_cast<float*>(acc + 8);
  STRAPS_CUDA(cudaMemsetAsync(acc, 0, 8 * sizeof(double), st));
  const int blocks = (int)std::min<long long>(592, (maxn + 255) / 256);
  loss_reduce_kernel<<<blocks, 256, 0, st>>>(a, acc);
  STRAPS_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 32, 0, st>>>(a, acc, log_vars, out, coef, d_log_vars);
  STRAPS_LAUNCH_CHECK();
  bool any_grad = false;
  for (int t = 0; t < 5; ++t) any_grad |= (a.on[t] && grads[t]);
  if (any_grad) {
    loss_backward_kernel<<<blocks, 256, 0, st>>>(a, coef, g_total);
    STRAPS_LAUNCH_CHECK();
  }
  return 0;
}

// synth.cu -- SURVEY.md 8f row N2: the target side of the synthetic training loop, on the device.
//
// Replaces, for train/train_synthetic_otf_rendering.py:121-145 of the reference:
//   smplx.lbs.batch_rodrigues (call sites augmentation/smpl_augmentation.py:55-58, train/...:190,301,322,
//     predict/predict_3D.py:134)                                                  -> batch_rodrigues_kernel
//   utils/cam_utils.py:40-71 perspective_project_torch                            -> perspective_kernel
//   the affine part of augmentation/smpl_augmentation.py:6-24 (shape sampling) and
//     augmentation/cam_augmentation.py:4-14 (camera translation sampling)        -> scale_shift_kernel
// The random draws themselves stay torch's generator (same seed => same stream as the reference on the same device);
// scale_shift_kernel rounds exactly like the reference's separate multiply / add ops (no FMA contraction), so the
// augmented parameters are bit-identical given the same draws.
// All three are one pass over a few KB: launch-latency bound, no roofline of their own.
#include "common.cuh"
#include "../../include/straps_b200.h"

namespace straps {

// angle = ||r + 1e-8|| (the epsilon goes inside the norm), axis = r / angle, R = I + sin.K + (1 - cos).K.K
__global__ void batch_rodrigues_kernel(const float* __restrict__ rv, long long n, float* __restrict__ R) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r0 = rv[i * 3 + 0], r1 = rv[i * 3 + 1], r2 = rv[i * 3 + 2];
  const float x = r0 + 1e-8f, y = r1 + 1e-8f, z = r2 + 1e-8f;
  const float angle = sqrtf(x * x + y * y + z * z);
  const float ax = r0 / angle, ay = r1 / angle, az = r2 / angle;
  const float c = cosf(angle), s = sinf(angle), omc = 1.f - c;
  const float K[9] = {0.f, -az, ay, az, 0.f, -ax, -ay, ax, 0.f};
  float* o = R + i * 9;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const float kk = K[a * 3 + 0] * K[0 * 3 + b] + K[a * 3 + 1] * K[1 * 3 + b] + K[a * 3 + 2] * K[2 * 3 + b];
      o[a * 3 + b] = (a == b ? 1.f : 0.f) + s * K[a * 3 + b] + omc * kk;
    }
}

// Gradient of batch_rodrigues_kernel (smplx builds this graph whenever an axis-angle pose requires grad: the reference's
// `smpl_model(betas=pred_shape)` call, train/train_synthetic_otf_rendering.py:206, runs with the module's own axis-angle
// parameters under autograd).  With u = r + 1e-8, t = |u|, a = r / t, K = skew(a), R = I + sin(t) K + (1 - cos(t)) K.K:
//   G = dL/dK = sin(t) dR + (1 - cos(t)) (dR K^T + K^T dR);   dL/da = (G21 - G12, G02 - G20, G10 - G01)
//   dL/dt = cos(t) <dR, K> + sin(t) <dR, K.K>;   dL/dr = dL/da / t + (dL/dt - <dL/da, r> / t^2) u / t
__global__ void batch_rodrigues_bwd_kernel(const float* __restrict__ rv, const float* __restrict__ dR, long long n,
                                           float* __restrict__ drv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r[3] = {rv[i * 3 + 0], rv[i * 3 + 1], rv[i * 3 + 2]};
  const float u[3] = {r[0] + 1e-8f, r[1] + 1e-8f, r[2] + 1e-8f};
  const float t = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  const float it = 1.f / t;
  const float ax = r[0] * it, ay = r[1] * it, az = r[2] * it;
  const float c = cosf(t), s = sinf(t), omc = 1.f - c;
  const float K[9] = {0.f, -az, ay, az, 0.f, -ax, -ay, ax, 0.f};
  float g[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) g[k] = dR[i * 9 + k];
  float KK[9], G[9];
  float dot_k = 0.f, dot_kk = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      KK[a * 3 + b] = K[a * 3 + 0] * K[0 * 3 + b] + K[a * 3 + 1] * K[1 * 3 + b] + K[a * 3 + 2] * K[2 * 3 + b];
      dot_k += g[a * 3 + b] * K[a * 3 + b];
      dot_kk += g[a * 3 + b] * KK[a * 3 + b];
    }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      // (dR K^T)_ab = sum_m dR_am K_bm ;  (K^T dR)_ab = sum_m K_ma dR_mb
      const float t1 = g[a * 3 + 0] * K[b * 3 + 0] + g[a * 3 + 1] * K[b * 3 + 1] + g[a * 3 + 2] * K[b * 3 + 2];
      const float t2 = K[0 * 3 + a] * g[0 * 3 + b] + K[1 * 3 + a] * g[1 * 3 + b] + K[2 * 3 + a] * g[2 * 3 + b];
      G[a * 3 + b] = s * g[a * 3 + b] + omc * (t1 + t2);
    }
  const float da[3] = {G[7] - G[5], G[2] - G[6], G[3] - G[1]};
  const float dt = c * dot_k + s * dot_kk;
  const float w = (dt - (da[0] * r[0] + da[1] * r[1] + da[2] * r[2]) * it * it) * it;
#pragma unroll
  for (int k = 0; k < 3; ++k) drv[i * 3 + k] = da[k] * it + w * u[k];
}

// q = R p + t;  q /= q.z (all three components, as the reference does);  out = (K q)[:2]
__global__ void perspective_kernel(const float* __restrict__ pts, const float* __restrict__ rot, const float* __restrict__ tr,
                                   const float* __restrict__ camK, int B, int N, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N;
  const float* Rm = rot + (size_t)b * 9;
  const float* Km = camK + (size_t)b * 9;
  const float px = pts[(size_t)i * 3 + 0], py = pts[(size_t)i * 3 + 1], pz = pts[(size_t)i * 3 + 2];
  const float qx = Rm[0] * px + Rm[1] * py + Rm[2] * pz + tr[b * 3 + 0];
  const float qy = Rm[3] * px + Rm[4] * py + Rm[5] * pz + tr[b * 3 + 1];
  const float qz = Rm[6] * px + Rm[7] * py + Rm[8] * pz + tr[b * 3 + 2];
  const float ux = qx / qz, uy = qy / qz, uz = qz / qz;
  out[(size_t)i * 2 + 0] = Km[0] * ux + Km[1] * uy + Km[2] * uz;
  out[(size_t)i * 2 + 1] = Km[3] * ux + Km[4] * uy + Km[5] * uz;
}

// out[r, c] = (noise[r, c] * mul[c] + add[c]) + base[r * base_stride + c], each operation rounded separately
__global__ void scale_shift_kernel(const float* __restrict__ noise, const float* __restrict__ mul, const float* __restrict__ add,
                                   const float* __restrict__ base, long long base_stride, int rows, int width,
                                   float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * width) return;
  const int r = i / width, c = i % width;
  float v = __fmul_rn(noise[i], mul[c]);
  if (add) v = __fadd_rn(v, add[c]);
  out[i] = __fadd_rn(v, base[(size_t)r * base_stride + c]);
}

}  // namespace straps

using namespace straps;

extern "C" int straps_batch_rodrigues(const float* rot_vecs, int64_t n, float* R, void* stream) {
  STRAPS_CHECK(rot_vecs && R, "straps_batch_rodrigues: null argument");
  if (n <= 0) return 0;
  batch_rodrigues_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(rot_vecs, n, R);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_batch_rodrigues_backward(const float* rot_vecs, const float* dR, int64_t n, float* d_rot_vecs, void* stream) {
  STRAPS_CHECK(rot_vecs && dR && d_rot_vecs, "straps_batch_rodrigues_backward: null argument");
  if (n <= 0) return 0;
  batch_rodrigues_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(rot_vecs, dR, n, d_rot_vecs);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_perspective_project(const float* points, const float* rotation, const float* translation,
                                          const float* cam_K, int batch, int npoints, float* out, void* stream) {
  STRAPS_CHECK(points && rotation && translation && cam_K && out, "straps_perspective_project: null argument");
  STRAPS_CHECK(batch >= 0 && npoints >= 0 && (long long)batch * npoints < (1ll << 31), "straps_perspective_project: bad sizes");
  if (batch * npoints == 0) return 0;
  perspective_kernel<<<ceil_div(batch * npoints, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(points, rotation, translation,
                                                                                                 cam_K, batch, npoints, out);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_scale_shift(const float* noise, const float* mul, const float* add, const float* base,
                                  int64_t base_stride, int rows, int width, float* out, void* stream) {
  STRAPS_CHECK(noise && mul && base && out, "straps_scale_shift: null argument");
  STRAPS_CHECK(rows >= 0 && width >= 0 && (long long)rows * width < (1ll << 31), "straps_scale_shift: bad sizes");
  if (rows * width == 0) return 0;
  scale_shift_kernel<<<ceil_div(rows * width, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(noise, mul, add, base, base_stride,
                                                                                              rows, width, out);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

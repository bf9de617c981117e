// train.cu -- training path of the regressor (BASELINE config 3): train-mode BatchNorm forward, full backward of the
// ResNet-18 encoder and of the IEF stack, and a fused Adam step on a flat parameter/gradient bucket.
//
// Replaces what autograd + optim.Adam do for the reference's step (train/train_synthetic_otf_rendering.py:186,230-233,
// run_train.py:200-201): BasicBlock.forward (models/resnet.py:61-77) in training mode, IEFModule.forward
// (models/ief_module.py:48-64) unrolled over its iterations with shared-weight gradient accumulation.
//
// This round the training kernels are fp32 CUDA-core code (correctness first, gradient-checked against the CPU
// oracle's autograd); the tensor-core data/weight gradients are next-round work.
//   conv forward / data gradient : conv_simt_kernel (regressor.cu), the data gradient is the same implicit GEMM with a
//                                  transposed gather over dY and weights repacked to [(kh,kw,co)][ci]
//   conv weight gradient          : wgrad_kernel, [tap x 64 ci] x [64 co] tiles, split over pixels, fp32 atomics
//   BatchNorm                     : two-pass batch statistics with fp64 accumulation (PyTorch's CPU kernel accumulates
//                                  in double), biased variance for normalisation, unbiased for the running estimate
//                                  (momentum 0.1, eps 1e-5); backward = 2 per-channel reductions + 1 elementwise pass
//   IEF backward                  : generic 32x32 SIMT GEMM (all four transpose combinations) + ReLU masks + column sums
// HBM layout: NHWC fp32 activations (regressor.cu); every conv keeps its pre-BN output `raw` for the backward.
#include "regressor.h"
#include <cstdlib>

namespace straps {

struct TrainState {
  int max_batch;
  float* pool;                 // one allocation
  size_t pool_bytes;
  float* raw[NCONV];           // pre-BN conv outputs
  std::vector<float*> gbuf;    // gradient wrt every activation buffer
  float* gmask;                // scratch: masked block-output gradient (largest block activation)
  float* draw;                 // scratch: gradient wrt a raw conv output (largest)
  float* mean[NCONV];
  float* invstd[NCONV];
  double* dstat;               // [2][512] fp64 accumulators
  float* fstat;                // [2][512] fp32 reductions (dbeta, dgamma) + [2][512] per-channel maxima (bits)
  float* ones; float* zeros;   // [512]
  float* w_dgrad[NCONV];       // [(kh,kw,co)][ci]
  float* dw_packed;            // [(kh,kw,ci_pad)][cout] scratch for the weight gradient (largest conv)
  int last_batch;
  int mode;                    // STRAPS_CONV_* of the last training forward (the backward follows it)
  int xin_valid;               // the fp32 NHWC copy of the input (CUDA-core conv1 weight gradient) matches the last forward
  bool draw_valid;             // the last bn_backward wrote the fp32 gradient `draw` (else only its split planes exist)
};

// ------------------------------------------------------------------------------------------------
// BatchNorm kernels (NHWC: channel = fastest index)
// ------------------------------------------------------------------------------------------------
// ONE pass over the pre-BN tensor: per channel S1 = sum(x - pivot), S2 = sum((x - pivot)^2) with pivot = the channel's first element
// (a sample of the distribution, so |mean - pivot| ~ sigma and S2 - S1^2 / n does not cancel).  Threads read float4 (4 channels) of
// 4 pixels per iteration, keep fp32 partials over at most 64 elements, fold them into fp64 registers, and the block adds its fp64
// totals to acc[0..C) / acc[512..512+C) with one atomic per channel.  (Round 1 ran two passes with scalar loads: 2.5 TB/s.)
// grid (chunks), block 256 = (C / 4 channel quads) x (1024 / C pixel lanes).
__global__ void __launch_bounds__(256) bn_stat_kernel(const float* __restrict__ x, long long npix, int C, double* __restrict__ acc) {
  __shared__ double red[2][256][4];
  const int tq = C >> 2;                       // threads per pixel
  const int lanes = 256 / tq;                  // pixels per block iteration
  const int q = threadIdx.x % tq, pl = threadIdx.x / tq;
  const float4 pv = *reinterpret_cast<const float4*>(x + 4 * q);
  double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  int pending = 0;
  const long long stride = (long long)gridDim.x * lanes;
  for (long long p = (long long)blockIdx.x * lanes + pl; p < npix; p += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long pp = p + u * stride;
      v[u] = (pp < npix) ? *reinterpret_cast<const float4*>(x + pp * C + 4 * q) : pv;      // pivot contributes zero
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float a = v[u].x - pv.x, b = v[u].y - pv.y, c = v[u].z - pv.z, d = v[u].w - pv.w;
      s1[0] += a; s1[1] += b; s1[2] += c; s1[3] += d;
      s2[0] = fmaf(a, a, s2[0]); s2[1] = fmaf(b, b, s2[1]); s2[2] = fmaf(c, c, s2[2]); s2[3] = fmaf(d, d, s2[3]);
    }
    if (++pending == 16) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { d1[k] += (double)s1[k]; d2[k] += (double)s2[k]; s1[k] = 0.f; s2[k] = 0.f; }
      pending = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { red[0][threadIdx.x][k] = d1[k] + (double)s1[k]; red[1][threadIdx.x][k] = d2[k] + (double)s2[k]; }
  __syncthreads();
  {                                            // C <= 256: one channel per thread; C = 512 takes two rounds
    for (int c = threadIdx.x; c < C; c += 256) {
      double t1 = 0, t2 = 0;
      for (int l = 0; l < lanes; ++l) { t1 += red[0][l * tq + (c >> 2)][c & 3]; t2 += red[1][l * tq + (c >> 2)][c & 3]; }
      atomicAdd(&acc[c], t1);
      atomicAdd(&acc[512 + c], t2);
    }
  }
}
// mean = pivot + S1 / n;  M2 = S2 - S1^2 / n (fp64);  biased variance normalises, the unbiased one feeds the running estimate
__global__ void bn_finalize_kernel(const double* __restrict__ acc, const float* __restrict__ x, long long npix, int C, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* __restrict__ rmean, float* __restrict__ rvar, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s1 = acc[c], s2 = acc[512 + c], n = (double)npix;
  const double mu = (double)x[c] + s1 / n;
  double m2 = s2 - s1 * s1 / n;
  if (m2 < 0) m2 = 0;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(m2 / n + 1e-5));
  if (rmean) {
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mu;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)(m2 / (n - 1.0));
  }
}
// y = (x - mean) * invstd * gamma + beta (+ res) (ReLU)
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ res,
                                int relu, long long n4, int C, float* __restrict__ y, __half* __restrict__ y_hi,
                                __half* __restrict__ y_lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)((i * 4) % C);
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
  const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
  float4 o = make_float4((v.x - mu.x) * is.x * g.x + b.x, (v.y - mu.y) * is.y * g.y + b.y, (v.z - mu.z) * is.z * g.z + b.z,
                         (v.w - mu.w) * is.w * g.w + b.w);
  if (res) { const float4 r = reinterpret_cast<const float4*>(res)[i]; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  reinterpret_cast<float4*>(y)[i] = o;
  if (y_hi) {   // tensor-core training path: the next conv reads the 2-term fp16 split
    __half h[4], l[4];
    split_f16(o.x, h[0], l[0]); split_f16(o.y, h[1], l[1]); split_f16(o.z, h[2], l[2]); split_f16(o.w, h[3], l[3]);
    reinterpret_cast<uint2*>(y_hi)[i] = make_uint2(pack_f16(h[0], h[1]), pack_f16(h[2], h[3]));
    reinterpret_cast<uint2*>(y_lo)[i] = make_uint2(pack_f16(l[0], l[1]), pack_f16(l[2], l[3]));
  }
}
// sums over pixels of dy and dy * xhat, with dy = gout * (act > 0) when act != null.  acc[0..C) = sum dy, acc[C..2C) = sum dy xhat.
// Same thread layout as bn_stat_kernel: float4 (4 channels) x 4 pixels per iteration, block = (C / 4) quads x (1024 / C) pixel lanes.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ gout, const float* __restrict__ act,
                                                            const float* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, long long npix, int C,
                                                            float* __restrict__ acc, unsigned* __restrict__ cmax) {
  __shared__ float red[2][256][4];
  const int tq = C >> 2, lanes = 256 / tq;
  const int q = threadIdx.x % tq, pl = threadIdx.x / tq;
  const float4 mu = *reinterpret_cast<const float4*>(mean + 4 * q), is = *reinterpret_cast<const float4*>(invstd + 4 * q);
  float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  float mdy[4] = {0, 0, 0, 0}, mxc[4] = {0, 0, 0, 0};       // max |dy| and max |x - mean| (bound of max |dx|: bn_bwd_scale_kernel)
  const long long stride = (long long)gridDim.x * lanes;
  for (long long p = (long long)blockIdx.x * lanes + pl; p < npix; p += 4 * stride) {
    float4 dy[4], xv[4], av[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long pp = p + u * stride;
      const bool in = pp < npix;
      const size_t o = (size_t)(in ? pp : 0) * C + 4 * q;
      dy[u] = in ? *reinterpret_cast<const float4*>(gout + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      xv[u] = *reinterpret_cast<const float4*>(x + o);
      av[u] = act ? *reinterpret_cast<const float4*>(act + o) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float d0 = (av[u].x > 0.f) ? dy[u].x : 0.f, d1 = (av[u].y > 0.f) ? dy[u].y : 0.f;
      const float d2 = (av[u].z > 0.f) ? dy[u].z : 0.f, d3 = (av[u].w > 0.f) ? dy[u].w : 0.f;
      s0[0] += d0; s0[1] += d1; s0[2] += d2; s0[3] += d3;
      s1[0] += d0 * (xv[u].x - mu.x) * is.x; s1[1] += d1 * (xv[u].y - mu.y) * is.y;
      s1[2] += d2 * (xv[u].z - mu.z) * is.z; s1[3] += d3 * (xv[u].w - mu.w) * is.w;
      mdy[0] = fmaxf(mdy[0], fabsf(d0)); mdy[1] = fmaxf(mdy[1], fabsf(d1)); mdy[2] = fmaxf(mdy[2], fabsf(d2)); mdy[3] = fmaxf(mdy[3], fabsf(d3));
      mxc[0] = fmaxf(mxc[0], fabsf(xv[u].x - mu.x)); mxc[1] = fmaxf(mxc[1], fabsf(xv[u].y - mu.y));
      mxc[2] = fmaxf(mxc[2], fabsf(xv[u].z - mu.z)); mxc[3] = fmaxf(mxc[3], fabsf(xv[u].w - mu.w));
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { red[0][threadIdx.x][k] = s0[k]; red[1][threadIdx.x][k] = s1[k]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t0 = 0.f, t1 = 0.f;
    for (int l = 0; l < lanes; ++l) { t0 += red[0][l * tq + (c >> 2)][c & 3]; t1 += red[1][l * tq + (c >> 2)][c & 3]; }
    atomicAdd(&acc[c], t0);
    atomicAdd(&acc[C + c], t1);
  }
  if (cmax) {                                  // non-negative floats order like their bit patterns
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) { red[0][threadIdx.x][k] = mdy[k]; red[1][threadIdx.x][k] = mxc[k]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float t0 = 0.f, t1 = 0.f;
      for (int l = 0; l < lanes; ++l) { t0 = fmaxf(t0, red[0][l * tq + (c >> 2)][c & 3]); t1 = fmaxf(t1, red[1][l * tq + (c >> 2)][c & 3]); }
      atomicMax(&cmax[c], __float_as_uint(t0));
      atomicMax(&cmax[C + c], __float_as_uint(t1));
    }
  }
}

// One block between the two passes: dbeta / dgamma to the caller's tensors, and -- tensor-core mode -- an upper bound of max |dx| from
// the per-channel sums and maxima, |dx| <= |gamma| invstd (max|dy| + |sum dy| / N + max|x - mean| invstd |sum dy xhat| / N).  Its bits go
// to *maxbits: the power of two that puts it into [2^13, 2^14) scales dX before the fp16 split (done by bn_bwd_apply_kernel itself
// instead of a separate pass over the tensor), is undone by the weight-gradient unpack, and -- folded into the data gradient's row
// unscales here -- by the data-gradient epilogue.  The bound is at most a few times the true maximum: two or three of fp16's thirty
// binades.
__global__ void __launch_bounds__(512) bn_bwd_scale_kernel(const float* __restrict__ acc, const unsigned* __restrict__ cmax,
                                                           const float* __restrict__ gamma, const float* __restrict__ invstd, int C, float inv_n,
                                                           float* __restrict__ dbeta, float* __restrict__ dgamma, unsigned* __restrict__ maxbits,
                                                           const float* __restrict__ w_unscale, int n_unscale, float* __restrict__ out_unscale) {
  __shared__ float red[16];
  __shared__ int s_e;
  const int c = threadIdx.x;
  float bound = 0.f;
  if (c < C) {
    const float s0 = acc[c], s1 = acc[C + c];
    dbeta[c] = s0;
    dgamma[c] = s1;
    if (maxbits) {
      const float is = invstd[c];
      bound = fabsf(gamma[c]) * is * (__uint_as_float(cmax[c]) + fabsf(s0) * inv_n + __uint_as_float(cmax[C + c]) * is * fabsf(s1) * inv_n);
    }
  }
  if (!maxbits) return;
  for (int o = 16; o > 0; o >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = bound;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int k = 1; k < 16; ++k) m = fmaxf(m, red[k]);
    *maxbits = __float_as_uint(m);
    int e = 0;
    if (m > 0.f && isfinite(m)) { frexpf(m, &e); e = max(-100, min(100, 14 - e)); }
    s_e = e;
  }
  __syncthreads();
  if (out_unscale) {
    const float inv = ldexpf(1.f, -s_e);
    for (int k = threadIdx.x; k < n_unscale; k += 512) out_unscale[k] = w_unscale[k] * inv;
  }
}
// dx = gamma * invstd * (dy - sum_dy/N - xhat * sum_dy_xhat/N) as fp32 (dx != null) and / or as scaled fp16 split planes (p_hi != null);
// optionally also writes the masked dy.
// 4 float4 per thread; n4 = npix * C / 4
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ gout, const float* __restrict__ act,
                                                           const float* __restrict__ x, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ acc, long long npix, int C, float* __restrict__ dx,
                                                           float* __restrict__ masked_out, const unsigned* __restrict__ maxbits,
                                                           __half* __restrict__ p_hi, __half* __restrict__ p_lo) {
  const long long n4 = npix * C / 4;
  const float inv_n = 1.f / (float)npix;
  float sc = 1.f;
  if (p_hi) {                                  // scaled fp16 split planes of dX for the tensor-core gradients (bound: bn_bwd_scale_kernel)
    int e = 0;
    const float mx = __uint_as_float(*maxbits);
    if (mx > 0.f && isfinite(mx)) { frexpf(mx, &e); e = max(-100, min(100, 14 - e)); }
    sc = ldexpf(1.f, e);
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const long long i = ((long long)blockIdx.x * 4 + it) * 256 + threadIdx.x;
    if (i >= n4) break;
    const int c = (int)((i * 4) % C);
    float4 dy = reinterpret_cast<const float4*>(gout)[i];
    if (act) {
      const float4 a = reinterpret_cast<const float4*>(act)[i];
      if (!(a.x > 0.f)) dy.x = 0.f;
      if (!(a.y > 0.f)) dy.y = 0.f;
      if (!(a.z > 0.f)) dy.z = 0.f;
      if (!(a.w > 0.f)) dy.w = 0.f;
    }
    if (masked_out) reinterpret_cast<float4*>(masked_out)[i] = dy;
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    const float4 s0 = *reinterpret_cast<const float4*>(acc + c), s1 = *reinterpret_cast<const float4*>(acc + C + c);
    float4 r;
    r.x = ga.x * is.x * (dy.x - s0.x * inv_n - ((xv.x - mu.x) * is.x) * s1.x * inv_n);
    r.y = ga.y * is.y * (dy.y - s0.y * inv_n - ((xv.y - mu.y) * is.y) * s1.y * inv_n);
    r.z = ga.z * is.z * (dy.z - s0.z * inv_n - ((xv.z - mu.z) * is.z) * s1.z * inv_n);
    r.w = ga.w * is.w * (dy.w - s0.w * inv_n - ((xv.w - mu.w) * is.w) * s1.w * inv_n);
    if (dx) reinterpret_cast<float4*>(dx)[i] = r;
    if (p_hi) {
      __half h[4], l[4];
      split_f16(r.x * sc, h[0], l[0]); split_f16(r.y * sc, h[1], l[1]); split_f16(r.z * sc, h[2], l[2]); split_f16(r.w * sc, h[3], l[3]);
      reinterpret_cast<uint2*>(p_hi)[i] = make_uint2(pack_f16(h[0], h[1]), pack_f16(h[2], h[3]));
      reinterpret_cast<uint2*>(p_lo)[i] = make_uint2(pack_f16(l[0], l[1]), pack_f16(l[2], l[3]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pooling backward
// ------------------------------------------------------------------------------------------------
// feat = mean over the 8x8 map; g wrt the map = dfeat / 64
__global__ void avgpool_bwd_kernel(const float* __restrict__ dfeat, int B, int HW, int C, float* __restrict__ g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW * C) return;
  const int c = (int)(i % C);
  const long long b = i / ((long long)HW * C);
  g[i] = dfeat[b * C + c] / (float)HW;
}
// 3x3/2 pad 1 max pool backward: the gradient of a window goes to its first maximum (scan order kh, kw)
__global__ void maxpool_bwd_kernel(const float* __restrict__ in, const float* __restrict__ gout, int B, int H, int W, int C,
                                   float* __restrict__ gin) {
  const int HO = H / 2, WO = W / 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HO * WO * C) return;
  const int c = (int)(i % C);
  long long t = i / C;
  const int ow = (int)(t % WO); t /= WO;
  const int oh = (int)(t % HO);
  const long long b = t / HO;
  float best = -INFINITY;
  long long arg = -1;
  for (int dh = 0; dh < 3; ++dh) {
    const int ih = oh * 2 - 1 + dh;
    if (ih < 0 || ih >= H) continue;
    for (int dw = 0; dw < 3; ++dw) {
      const int iw = ow * 2 - 1 + dw;
      if (iw < 0 || iw >= W) continue;
      const long long o = ((b * H + ih) * W + iw) * C + c;
      const float v = in[o];
      if (v > best) { best = v; arg = o; }
    }
  }
  if (arg >= 0) atomicAdd(&gin[arg], gout[i]);
}

// ------------------------------------------------------------------------------------------------
// convolution weight gradient:  dW[(tap,ci)][co] = sum_pixels A[pixel][(tap,ci)] * dY[pixel][co]
// grid (taps * ci_tiles, cout/64, splits); block 256: 16x16 threads, 4 ci x 4 co each; 16 pixels per step
// ------------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* in;   // NHWC [B,hin,win,cin]
  const float* dy;   // NHWC [B,hout,wout,cout]
  float* dw;         // [(kh,kw,cin)][cout], zero-initialised
  int B, hin, win, cin, hout, wout, cout, ks, stride, pad, ci_tile;
};
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64];
  const int ci_tiles = (a.cin + a.ci_tile - 1) / a.ci_tile;
  const int tap = blockIdx.x / ci_tiles, ci0 = (blockIdx.x % ci_tiles) * a.ci_tile;
  const int kh = tap / a.ks, kw = tap % a.ks;
  const int co0 = blockIdx.y * 64;
  const long long npix = (long long)a.B * a.hout * a.wout;
  const long long per = (npix + gridDim.z - 1) / gridDim.z;
  const long long p0 = (long long)blockIdx.z * per, p1 = min(npix, p0 + per);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // loaders: 16 pixels x 64 values = 256 float4 -> one float4 per thread for each operand
  const int lp = tid >> 4, l4 = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long pb = p0; pb < p1; pb += 16) {
    const long long p = pb + lp;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (p < p1) {
      const int ow = (int)(p % a.wout);
      const long long t = p / a.wout;
      const int oh = (int)(t % a.hout);
      const long long b = t / a.hout;
      const int ih = oh * a.stride - a.pad + kh, iw = ow * a.stride - a.pad + kw;
      if (ih >= 0 && ih < a.hin && iw >= 0 && iw < a.win && ci0 + l4 < a.cin && l4 < a.ci_tile)
        av = *reinterpret_cast<const float4*>(a.in + ((b * a.hin + ih) * a.win + iw) * a.cin + ci0 + l4);
      bv = *reinterpret_cast<const float4*>(a.dy + p * a.cout + co0 + l4);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lp][l4]) = av;
    *reinterpret_cast<float4*>(&Bs[lp][l4]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= a.cin || ty * 4 + i >= a.ci_tile) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&a.dw[((size_t)tap * a.cin + ci) * a.cout + co0 + tx * 4 + j], acc[i][j]);
  }
}
// [(kh,kw,ci_pad)][cout] -> OIHW
__global__ void unpack_w_kernel(const float* __restrict__ packed, int cout, int cin, int cin_pad, int ks, float* __restrict__ w) {
  const int total = cout * cin * ks * ks;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kw = i % ks;
  int t = i / ks;
  const int kh = t % ks; t /= ks;
  const int ci = t % cin;
  const int co = t / cin;
  w[i] = packed[((size_t)(kh * ks + kw) * cin_pad + ci) * cout + co];
}
// OIHW -> [(kh,kw,co)][ci] for the data gradient
__global__ void pack_w_dgrad_kernel(const float* __restrict__ w, int cout, int cin, int ks, float* __restrict__ out) {
  const int total = ks * ks * cout * cin;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = i % cin;
  int t = i / cin;
  const int co = t % cout; t /= cout;
  const int kw = t % ks, kh = t / ks;
  out[i] = w[(((size_t)co * cin + ci) * ks + kh) * ks + kw];
}

// ------------------------------------------------------------------------------------------------
// small GEMM for the IEF backward:  C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C,  row-major with leading dims
// ------------------------------------------------------------------------------------------------
// C (+)= op(A) op(B), fp32, 32x32 tiles.  gridDim.z > 1 splits K: every slice adds its partial tile with atomics, so C must hold the
// value to accumulate onto (zeros for a plain product).  The tile loads walk the CONTIGUOUS index of each operand with consecutive
// threads whatever the transposition (round 1 read transposed operands with a stride of lda / ldb floats between lanes).
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, int lda, int ta, const float* __restrict__ Bm, int ldb, int tb,
                                                   float* __restrict__ C, int ldc, int M, int N, int K, float beta, int k_per) {
  __shared__ float As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int kbeg = blockIdx.z * k_per, kend = min(K, kbeg + k_per);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = kbeg; k0 < kend; k0 += 32) {
    for (int i = threadIdx.x; i < 1024; i += 256) {
      const int lo = i & 31, hi = i >> 5;
      {                                        // As[r][c] = op(A)[m0 + r][k0 + c]
        const int r = ta ? lo : hi, c = ta ? hi : lo;
        const int m = m0 + r, k = k0 + c;
        As[r][c] = (m < M && k < kend) ? (ta ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]) : 0.f;
      }
      {                                        // Bs[r][c] = op(B)[k0 + r][n0 + c]
        const int r = tb ? lo : hi, c = tb ? hi : lo;
        const int kk = k0 + r, n = n0 + c;
        Bs[r][c] = (kk < kend && n < N) ? (tb ? Bm[(size_t)n * ldb + kk] : Bm[(size_t)kk * ldb + n]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[ty * 2][k], a1 = As[ty * 2 + 1][k], b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = m0 + ty * 2 + i, n = n0 + tx * 2 + j;
      if (m < M && n < N) {
        if (gridDim.z > 1) atomicAdd(&C[(size_t)m * ldc + n], acc[i][j]);
        else C[(size_t)m * ldc + n] = acc[i][j] + (beta != 0.f ? beta * C[(size_t)m * ldc + n] : 0.f);
      }
    }
}
// y[b][n] *= (h[b][n] > 0)
__global__ void relu_mask_kernel(float* __restrict__ y, const float* __restrict__ h, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(h[i] > 0.f)) y[i] = 0.f;
}
// out[n] += sum_b y[b][n]
__global__ void colsum_acc_kernel(const float* __restrict__ y, int B, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += y[(size_t)b * N + n];
  out[n] += s;
}
// dst[b][c] (ld) = a[b][c] + (add ? dst : 0)
__global__ void add_block_kernel(const float* __restrict__ a, int lda, float* __restrict__ dst, int ldd, int B, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  dst[(size_t)(i / N) * ldd + i % N] += a[(size_t)(i / N) * lda + i % N];
}

// Adam (torch.optim.Adam defaults, no weight decay / amsgrad) on a flat fp32 bucket
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2, float gscale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// The same update with the step count kept ON THE DEVICE (a captured CUDA graph replays with fixed kernel arguments, so the bias
// corrections cannot come from the host): adam_tick_kernel advances the counter, adam_dev_kernel derives 1 - beta^t from it.
__global__ void adam_tick_kernel(long long* __restrict__ step) { *step += 1; }
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, const long long* __restrict__ step, float lr, float b1, float b2, float eps, float gscale) {
  __shared__ float sbc[2];
  if (threadIdx.x == 0) {
    const float t = (float)*step;
    sbc[0] = 1.f - powf(b1, t);
    sbc[1] = 1.f - powf(b2, t);
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float bc1 = sbc[0], bc2 = sbc[1];
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

static int gemm(const float* A, int lda, int ta, const float* B, int ldb, int tb, float* C, int ldc, int M, int N, int K, float beta,
                cudaStream_t st) {
  // few tiles and a long K (the data gradients of the IEF layers: M = batch): split K over enough CTAs to fill the machine
  const int tiles = ceil_div(N, 32) * ceil_div(M, 32);
  int splits = 1;
  if (tiles < 120 && K >= 128 && (beta == 0.f || beta == 1.f)) splits = std::min(ceil_div(K, 32), std::max(1, 296 / tiles));
  const int k_per = ceil_div(ceil_div(K, splits), 32) * 32;
  splits = ceil_div(K, k_per);
  if (splits > 1 && beta == 0.f) {
    STRAPS_CHECK(ldc == N, "gemm: split-K needs a dense C to clear");
    STRAPS_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st));
  }
  gemm_kernel<<<dim3(ceil_div(N, 32), ceil_div(M, 32), splits), 256, 0, st>>>(A, lda, ta, B, ldb, tb, C, ldc, M, N, K, beta, k_per);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

static int train_ensure(straps_regressor* r) {
  if (r->train) return 0;
  TrainState* t = new TrainState();
  t->max_batch = r->max_batch;
  t->last_batch = 0;
  size_t fl = 0;   // floats
  auto take = [&](size_t n) { size_t o = fl; fl += (n + 63) & ~(size_t)63; return o; };
  size_t off_raw[NCONV], off_wd[NCONV], off_mean[NCONV], off_inv[NCONV];
  size_t max_act = 0, max_w = 0;
  for (int i = 0; i < NCONV; ++i) {
    const ConvSpec& c = r->conv[i];
    const size_t n = (size_t)r->max_batch * c.hout * c.wout * c.cout;
    off_raw[i] = take(n);
    off_wd[i] = take((size_t)c.ksize * c.ksize * c.cout * c.cin);
    off_mean[i] = take(512); off_inv[i] = take(512);
    max_act = n > max_act ? n : max_act;
    const size_t nw = (size_t)c.ksize * c.ksize * c.cin_pad * c.cout;
    max_w = nw > max_w ? nw : max_w;
  }
  std::vector<size_t> off_g(r->bufs.size());
  for (size_t i = 0; i < r->bufs.size(); ++i) {
    const ActBuf& b = r->bufs[i];
    off_g[i] = ((int)i == r->buf_xin) ? 0 : take((size_t)r->max_batch * b.h * b.w * b.c);
  }
  const size_t off_gmask = take((size_t)r->max_batch * 64 * 64 * 64);
  const size_t off_draw = take(max_act);
  const size_t off_dw = take(max_w);
  const size_t off_fstat = take(2048), off_ones = take(512), off_zeros = take(512);
  const size_t off_dstat = take(2048);   // 1024 doubles
  t->pool_bytes = fl * sizeof(float);
  if (cudaMalloc(&t->pool, t->pool_bytes) != cudaSuccess) {
    set_error("training workspace: cudaMalloc of %zu bytes failed: %s", t->pool_bytes, cudaGetErrorString(cudaGetLastError()));
    delete t;
    return 1;
  }
  for (int i = 0; i < NCONV; ++i) {
    t->raw[i] = t->pool + off_raw[i]; t->w_dgrad[i] = t->pool + off_wd[i];
    t->mean[i] = t->pool + off_mean[i]; t->invstd[i] = t->pool + off_inv[i];
  }
  t->gbuf.resize(r->bufs.size());
  for (size_t i = 0; i < r->bufs.size(); ++i) t->gbuf[i] = ((int)i == r->buf_xin) ? nullptr : t->pool + off_g[i];
  t->gmask = t->pool + off_gmask; t->draw = t->pool + off_draw; t->dw_packed = t->pool + off_dw;
  t->fstat = t->pool + off_fstat; t->ones = t->pool + off_ones; t->zeros = t->pool + off_zeros;
  t->dstat = reinterpret_cast<double*>(t->pool + off_dstat);
  std::vector<float> one(512, 1.f);
  STRAPS_CUDA(cudaMemcpy(t->ones, one.data(), 512 * sizeof(float), cudaMemcpyHostToDevice));
  STRAPS_CUDA(cudaMemset(t->zeros, 0, 512 * sizeof(float)));
  r->train = t;
  return 0;
}

const float* train_last_draw(const straps_regressor* r) {
  const TrainState* t = static_cast<const TrainState*>(r->train);
  return (t && t->draw_valid) ? t->draw : nullptr;      // null when the last gradient exists as split planes only
}

void train_destroy(straps_regressor* r) {
  TrainState* t = static_cast<TrainState*>(r->train);
  if (!t) return;
  if (t->pool) cudaFree(t->pool);
  delete t;
  r->train = nullptr;
}

static int conv_raw(straps_regressor* r, TrainState* t, int ci, int B, cudaStream_t st) {
  if (t->mode == STRAPS_CONV_F16X3_TC) return tc_train_conv_fwd(r, ci, B, t->raw[ci], st);
  const ConvSpec& c = r->conv[ci];
  ConvArgs a;
  a.in = act_ptr(r, c.in_buf); a.w = c.w_simt; a.scale = t->ones; a.shift = t->zeros; a.res = nullptr; a.out = t->raw[ci];
  a.B = B; a.hin = c.hin; a.win = c.win; a.cin = c.cin_pad; a.hout = c.hout; a.wout = c.wout; a.cout = c.cout;
  a.ks = c.ksize; a.stride = c.stride; a.pad = c.pad; a.relu = 0; a.transposed = 0;
  return launch_conv_simt(a, st);
}

static int bn_forward(straps_regressor* r, TrainState* t, int ci, int B, const float* res, int relu, int out_buf, int update_running,
                      cudaStream_t st) {
  const ConvSpec& c = r->conv[ci];
  float* out = act_ptr(r, out_buf);
  const bool tc = t->mode == STRAPS_CONV_F16X3_TC;
  __half* out_hi = tc ? tc_train_plane(r, out_buf, 0) : nullptr;
  __half* out_lo = tc ? tc_train_plane(r, out_buf, 1) : nullptr;
  const long long npix = (long long)B * c.hout * c.wout;
  const int lanes = 1024 / c.cout;                                      // pixels per block iteration of bn_stat_kernel
  const int chunks = (int)std::max<long long>(1, std::min<long long>(1184, npix / (16LL * lanes)));
  STRAPS_CUDA(cudaMemsetAsync(t->dstat, 0, 1024 * sizeof(double), st));
  bn_stat_kernel<<<chunks, 256, 0, st>>>(t->raw[ci], npix, c.cout, t->dstat);
  STRAPS_LAUNCH_CHECK();
  bn_finalize_kernel<<<ceil_div(c.cout, 128), 128, 0, st>>>(t->dstat, t->raw[ci], npix, c.cout, t->mean[ci], t->invstd[ci],
                                                           update_running ? c.rmean : nullptr, update_running ? c.rvar : nullptr, 0.1f);
  STRAPS_LAUNCH_CHECK();
  const long long n4 = npix * c.cout / 4;
  bn_apply_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(t->raw[ci], t->mean[ci], t->invstd[ci], c.gamma, c.beta, res, relu, n4,
                                                              c.cout, out, out_hi, out_lo);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// gout: gradient wrt the BN output (after the optional ReLU whose output is `act`); writes d(raw conv out) into t->draw,
// dgamma/dbeta into the caller's tensors, and (optionally) the ReLU-masked gout into masked_out.
static bool wgrad_on_tc() {
  const char* e = getenv("STRAPS_WGRAD");      // "simt" keeps the fp32 CUDA-core weight gradient in tensor-core mode (debugging)
  return !(e && e[0] == 's');
}

static int bn_backward(straps_regressor* r, TrainState* t, int ci, int B, const float* gout, const float* act, float* masked_out,
                       float* dgamma, float* dbeta, cudaStream_t st) {
  const ConvSpec& c = r->conv[ci];
  const long long npix = (long long)B * c.hout * c.wout;
  const int chunks = (int)std::max<long long>(1, std::min<long long>(1184, npix / (16LL * (1024 / c.cout))));
  STRAPS_CUDA(cudaMemsetAsync(t->fstat, 0, 2048 * sizeof(float), st));      // sums [2][512] and maxima [2][512]
  const bool tc = t->mode == STRAPS_CONV_F16X3_TC;
  // tensor-core gradients read dX as scaled fp16 split planes: written here by the apply pass (no separate split pass over the tensor).
  // The fp32 tensor is still written where something reads it: the zero-upsampled split of a stride-2 data gradient, the CUDA-core
  // weight gradient, the fp32 mode.
  const bool planes = tc && wgrad_on_tc();
  const bool need_f32 = !planes || c.stride == 2;
  unsigned* maxbits = tc ? tc_train_dy_max(r, ci) : nullptr;
  unsigned* cmax = reinterpret_cast<unsigned*>(t->fstat + 1024);
  bn_bwd_reduce_kernel<<<chunks, 256, 0, st>>>(gout, act, t->raw[ci], t->mean[ci], t->invstd[ci], npix, c.cout, t->fstat, tc ? cmax : nullptr);
  STRAPS_LAUNCH_CHECK();
  const float* w_unscale = nullptr;
  float* out_unscale = nullptr;
  int n_unscale = 0;
  __half *p_hi = nullptr, *p_lo = nullptr;
  if (tc) {
    tc_train_dgrad_scales(r, ci, &w_unscale, &n_unscale, &out_unscale);
    if (planes) tc_train_dy_planes(r, ci, &p_hi, &p_lo);
  }
  bn_bwd_scale_kernel<<<1, 512, 0, st>>>(t->fstat, cmax, c.gamma, t->invstd[ci], c.cout, 1.f / (float)npix, dbeta, dgamma, maxbits, w_unscale,
                                        n_unscale, out_unscale);
  STRAPS_LAUNCH_CHECK();
  const long long n4 = npix * c.cout / 4;
  bn_bwd_apply_kernel<<<(unsigned)((n4 + 1023) / 1024), 256, 0, st>>>(gout, act, t->raw[ci], t->mean[ci], t->invstd[ci], c.gamma, t->fstat, npix,
                                                                 c.cout, need_f32 ? t->draw : nullptr, masked_out, maxbits, p_hi, p_lo);
  STRAPS_LAUNCH_CHECK();
  t->draw_valid = need_f32;
  return 0;
}


static int conv_wgrad(straps_regressor* r, TrainState* t, int ci, int B, float* dw_oihw, cudaStream_t st) {
  if (t->mode == STRAPS_CONV_F16X3_TC) {
    // the plain split planes of dY (written by bn_backward's apply pass) serve the weight gradient and, for stride-1 convs, the data
    // gradient that follows; with the CUDA-core weight gradient they are made here from the fp32 tensor
    if (wgrad_on_tc()) return tc_train_conv_wgrad(r, ci, B, dw_oihw, st);
    if (tc_train_split_dy(r, ci, B, t->draw, 0, st)) return 1;
  }
  const ConvSpec& c = r->conv[ci];
  const size_t nw = (size_t)c.ksize * c.ksize * c.cin_pad * c.cout;
  STRAPS_CUDA(cudaMemsetAsync(t->dw_packed, 0, nw * sizeof(float), st));
  WgradArgs a;
  a.in = act_ptr(r, c.in_buf); a.dy = t->draw; a.dw = t->dw_packed;
  a.B = B; a.hin = c.hin; a.win = c.win; a.cin = c.cin_pad; a.hout = c.hout; a.wout = c.wout; a.cout = c.cout;
  a.ks = c.ksize; a.stride = c.stride; a.pad = c.pad; a.ci_tile = c.cin_pad < 64 ? c.cin_pad : 64;
  const int ci_tiles = (c.cin_pad + a.ci_tile - 1) / a.ci_tile;
  const int base = c.ksize * c.ksize * ci_tiles * (c.cout / 64);
  const long long npix = (long long)B * c.hout * c.wout;
  int splits = (int)std::max<long long>(1, std::min<long long>((npix + 255) / 256, (592 + base - 1) / base));
  wgrad_kernel<<<dim3(c.ksize * c.ksize * ci_tiles, c.cout / 64, splits), 256, 0, st>>>(a);
  STRAPS_LAUNCH_CHECK();
  const int total = c.cout * c.cin * c.ksize * c.ksize;
  unpack_w_kernel<<<ceil_div(total, 256), 256, 0, st>>>(t->dw_packed, c.cout, c.cin, c.cin_pad, c.ksize, dw_oihw);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

// gin[in_buf] (+)= dgrad(t->draw);  add != null -> that tensor is added (identity / other-branch gradient)
static int conv_dgrad(straps_regressor* r, TrainState* t, int ci, int B, const float* add, float* gin, cudaStream_t st) {
  if (t->mode == STRAPS_CONV_F16X3_TC) {   // always preceded by conv_wgrad of the same conv (same dY)
    if (r->conv[ci].stride == 2 && tc_train_split_dy(r, ci, B, t->draw, 1, st)) return 1;
    return tc_train_conv_dgrad(r, ci, B, add, gin, st);
  }
  const ConvSpec& c = r->conv[ci];
  ConvArgs a;
  a.in = t->draw; a.w = t->w_dgrad[ci]; a.scale = t->ones; a.shift = t->zeros; a.res = add; a.out = gin;
  a.B = B; a.hin = c.hout; a.win = c.wout; a.cin = c.cout; a.hout = c.hin; a.wout = c.win; a.cout = c.cin;
  a.ks = c.ksize; a.stride = c.stride; a.pad = c.pad; a.relu = 0; a.transposed = 1;
  return launch_conv_simt(a, st);
}

}  // namespace straps

using namespace straps;

extern "C" int straps_encoder_train_forward(straps_regressor_t* r, const float* x, int batch, int update_running_stats, int conv_mode,
                                            float* feat, void* stream) {
  STRAPS_CHECK(r && x && feat, "straps_encoder_train_forward: null argument");
  STRAPS_CHECK(r->loaded, "straps_encoder_train_forward: weights not loaded");
  STRAPS_CHECK(batch >= 2 && batch <= r->max_batch, "straps_encoder_train_forward: batch %d outside [2,%d]", batch, r->max_batch);
  if (train_ensure(r)) return 1;
  TrainState* t = static_cast<TrainState*>(r->train);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  STRAPS_CHECK(conv_mode == STRAPS_CONV_FP32_SIMT || conv_mode == STRAPS_CONV_F16X3_TC, "straps_encoder_train_forward: unknown conv_mode %d",
               conv_mode);
  t->last_batch = batch;
  t->mode = conv_mode;
  r->last_mode = STRAPS_CONV_FP32_SIMT;   // the activation buffers hold fp32 NHWC in both training modes
  if (conv_mode == STRAPS_CONV_F16X3_TC) {
    if (straps::tc_train_begin(r, batch, st)) return 1;
    if (straps::tc_train_pack_input(r, x, batch, st)) return 1;
    t->xin_valid = 0;      // rebuilt from the split planes if a CUDA-core weight gradient is requested
  } else {
    if (straps::ensure_packed(r, STRAPS_CONV_FP32_SIMT, 0, st)) return 1;
    if (straps::launch_nchw_to_nhwc(r, x, batch, st)) return 1;
    t->xin_valid = 1;
  }
  // stem: conv1 -> bn1 -> relu -> maxpool
  if (conv_raw(r, t, 0, batch, st)) return 1;
  if (bn_forward(r, t, 0, batch, nullptr, 1, r->buf_stem, update_running_stats, st)) return 1;
  if (straps::launch_maxpool(r, batch, st)) return 1;
  if (conv_mode == STRAPS_CONV_F16X3_TC && straps::tc_train_split_act(r, r->buf_pool, batch, st)) return 1;
  int i = 1;
  while (i < NCONV) {
    const bool ds = (i + 2 < NCONV) && r->conv[i + 2].ksize == 1;
    const ConvSpec &c1 = r->conv[i], &c2 = r->conv[i + 1];
    if (conv_raw(r, t, i, batch, st)) return 1;
    if (bn_forward(r, t, i, batch, nullptr, 1, c1.out_buf, update_running_stats, st)) return 1;
    if (ds) {
      if (conv_raw(r, t, i + 2, batch, st)) return 1;
      if (bn_forward(r, t, i + 2, batch, nullptr, 0, r->conv[i + 2].out_buf, update_running_stats, st)) return 1;
    }
    if (conv_raw(r, t, i + 1, batch, st)) return 1;
    if (bn_forward(r, t, i + 1, batch, act_ptr(r, c2.res_buf), 1, c2.out_buf, update_running_stats, st)) return 1;
    i += ds ? 3 : 2;
  }
  return straps::launch_avgpool(r, batch, feat, st);
}

// d_conv_w[20]: OIHW gradients (state_dict order); d_bn[40]: (dgamma, dbeta) per BatchNorm.  All PyTorch-owned, overwritten.
// The backward pass in two halves if the caller wishes (block_hi .. block_lo of the 8 BasicBlocks, last first; `first` = this call
// starts the pass, `stem` = it also ends it with the stem): a data-parallel caller starts the all-reduce of the layer4 + IEF gradients
// -- 76 % of the bucket, complete after the first two blocks -- while the rest of the pass runs.
extern "C" int straps_encoder_backward_range(straps_regressor_t* r, const float* dfeat, int batch, int conv_mode, float* const* d_conv_w,
                                             float* const* d_bn, int block_hi, int block_lo, int first, int stem, void* stream) {
  STRAPS_CHECK(r && dfeat && d_conv_w && d_bn, "straps_encoder_backward: null argument");
  STRAPS_CHECK(block_hi <= 7 && block_lo >= 0 && block_lo <= block_hi + 1, "straps_encoder_backward_range: blocks %d..%d", block_hi, block_lo);
  TrainState* t = static_cast<TrainState*>(r->train);
  STRAPS_CHECK(t && t->last_batch == batch, "straps_encoder_backward: no matching straps_encoder_train_forward (batch %d)", batch);
  STRAPS_CHECK(conv_mode == -1 || conv_mode == t->mode || (conv_mode == STRAPS_CONV_FP32_SIMT && t->mode == STRAPS_CONV_F16X3_TC),
               "straps_encoder_backward: conv_mode %d is not available after a forward in mode %d", conv_mode, t->mode);
  // the fp32 CUDA-core gradients can run on the activations a tensor-core forward saved (they are fp32 in both modes)
  struct ModeGuard { TrainState* t; int saved; ~ModeGuard() { t->mode = saved; } } guard{t, t->mode};
  if (conv_mode != -1) t->mode = conv_mode;
  const bool fp32_wgrad = t->mode == STRAPS_CONV_FP32_SIMT || !wgrad_on_tc();
  if (fp32_wgrad && !t->xin_valid) {   // the tensor-core forward did not write the fp32 NHWC input the CUDA-core conv1 weight gradient reads
    if (straps::tc_train_unpack_input(r, batch, static_cast<cudaStream_t>(stream))) return 1;
    t->xin_valid = 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (first && t->mode == STRAPS_CONV_F16X3_TC && straps::tc_train_pack_dgrad(r, st)) return 1;
  for (int i = 1; first && i < NCONV && t->mode != STRAPS_CONV_F16X3_TC; ++i) {   // data-gradient weight layout (conv1 needs none)
    const ConvSpec& c = r->conv[i];
    const int total = c.ksize * c.ksize * c.cout * c.cin;
    pack_w_dgrad_kernel<<<ceil_div(total, 256), 256, 0, st>>>(c.w_oihw, c.cout, c.cin, c.ksize, t->w_dgrad[i]);
    STRAPS_LAUNCH_CHECK();
  }
  if (first) {
    const long long n = (long long)batch * 64 * 512;
    avgpool_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dfeat, batch, 64, 512, t->gbuf[r->buf_final]);
    STRAPS_LAUNCH_CHECK();
  }
  // blocks in reverse
  int starts[8], nblk = 0;
  for (int i = 1; i < NCONV;) { starts[nblk++] = i; i += ((i + 2 < NCONV) && r->conv[i + 2].ksize == 1) ? 3 : 2; }
  for (int bi = std::min(nblk - 1, block_hi); bi >= block_lo; --bi) {
    const int i = starts[bi];
    const bool ds = (i + 2 < NCONV) && r->conv[i + 2].ksize == 1;
    const ConvSpec &c1 = r->conv[i], &c2 = r->conv[i + 1];
    const float* g_out = t->gbuf[c2.out_buf];
    float* g_in = t->gbuf[c1.in_buf];
    // conv2 branch: relu mask from the block output, bn2 backward (also materialises the masked gradient for the skip path)
    if (bn_backward(r, t, i + 1, batch, g_out, act_ptr(r, c2.out_buf), t->gmask, d_bn[2 * (i + 1)], d_bn[2 * (i + 1) + 1], st)) return 1;
    if (conv_wgrad(r, t, i + 1, batch, d_conv_w[i + 1], st)) return 1;
    if (conv_dgrad(r, t, i + 1, batch, nullptr, t->gbuf[c1.out_buf], st)) return 1;           // gradient wrt a = relu(bn1(.))
    // conv1 branch
    if (bn_backward(r, t, i, batch, t->gbuf[c1.out_buf], act_ptr(r, c1.out_buf), nullptr, d_bn[2 * i], d_bn[2 * i + 1], st)) return 1;
    if (conv_wgrad(r, t, i, batch, d_conv_w[i], st)) return 1;
    if (!ds) {
      if (conv_dgrad(r, t, i, batch, t->gmask, g_in, st)) return 1;                             // + identity gradient
    } else {
      if (conv_dgrad(r, t, i, batch, nullptr, g_in, st)) return 1;
      if (bn_backward(r, t, i + 2, batch, t->gmask, nullptr, nullptr, d_bn[2 * (i + 2)], d_bn[2 * (i + 2) + 1], st)) return 1;
      if (conv_wgrad(r, t, i + 2, batch, d_conv_w[i + 2], st)) return 1;
      if (conv_dgrad(r, t, i + 2, batch, g_in, g_in, st)) return 1;                             // accumulate the downsample branch
    }
  }
  // stem: maxpool backward -> bn1 (relu mask from the stem activation) -> conv1 weight gradient
  if (stem) {
    const size_t nstem = (size_t)batch * 128 * 128 * 64;
    STRAPS_CUDA(cudaMemsetAsync(t->gbuf[r->buf_stem], 0, nstem * sizeof(float), st));
    const long long n = (long long)batch * 64 * 64 * 64;
    maxpool_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(act_ptr(r, r->buf_stem), t->gbuf[r->buf_pool], batch, 128, 128, 64,
                                                                  t->gbuf[r->buf_stem]);
    STRAPS_LAUNCH_CHECK();
    if (bn_backward(r, t, 0, batch, t->gbuf[r->buf_stem], act_ptr(r, r->buf_stem), nullptr, d_bn[0], d_bn[1], st)) return 1;
    if (conv_wgrad(r, t, 0, batch, d_conv_w[0], st)) return 1;
  }
  if (t->mode == STRAPS_CONV_F16X3_TC && straps::tc_train_unpack_all(r, st)) return 1;     // the weight gradients computed by THIS call
  return 0;
}

extern "C" int straps_encoder_backward(straps_regressor_t* r, const float* dfeat, int batch, int conv_mode, float* const* d_conv_w,
                                       float* const* d_bn, void* stream) {
  return straps_encoder_backward_range(r, dfeat, batch, conv_mode, d_conv_w, d_bn, 7, 0, 1, 1, stream);
}

// saved: dev [iters] x { p_k [B,157] | h1_k [B,512] | h2_k [B,512] } written by straps_ief_forward_train.
// Outputs (overwritten): d_feat [B,512], d_fc_w[3] in nn.Linear layout, d_fc_b[3].  scratch: dev [B * (157 + 512 + 512 + 669)].
extern "C" int straps_ief_backward(straps_regressor_t* r, const float* feat, const float* saved, const float* d_params, int batch, int iters,
                                   float* d_feat, float* const* d_fc_w, float* const* d_fc_b, float* scratch, void* stream) {
  STRAPS_CHECK(r && feat && saved && d_params && d_feat && d_fc_w && d_fc_b && scratch, "straps_ief_backward: null argument");
  STRAPS_CHECK(r->loaded, "straps_ief_backward: weights not loaded");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = batch, P = STRAPS_IEF_PARAMS, H = IEF_H, IN = IEF_IN, F = STRAPS_FEAT_DIM;
  float* dp = scratch;                       // [B,157]
  float* dh2 = dp + (size_t)B * P;           // [B,512]
  float* dh1 = dh2 + (size_t)B * H;          // [B,512]
  float* dsx = dh1 + (size_t)B * H;          // [B,669]
  STRAPS_CUDA(cudaMemcpyAsync(dp, d_params, (size_t)B * P * sizeof(float), cudaMemcpyDeviceToDevice, st));
  STRAPS_CUDA(cudaMemsetAsync(d_feat, 0, (size_t)B * F * sizeof(float), st));
  STRAPS_CUDA(cudaMemsetAsync(d_fc_w[0], 0, (size_t)H * IN * sizeof(float), st));
  STRAPS_CUDA(cudaMemsetAsync(d_fc_w[1], 0, (size_t)H * H * sizeof(float), st));
  STRAPS_CUDA(cudaMemsetAsync(d_fc_w[2], 0, (size_t)P * H * sizeof(float), st));
  STRAPS_CUDA(cudaMemsetAsync(d_fc_b[0], 0, H * sizeof(float), st));
  STRAPS_CUDA(cudaMemsetAsync(d_fc_b[1], 0, H * sizeof(float), st));
  STRAPS_CUDA(cudaMemsetAsync(d_fc_b[2], 0, P * sizeof(float), st));
  const size_t per_iter = (size_t)B * (P + 2 * H);
  for (int k = iters - 1; k >= 0; --k) {
    const float* pk = saved + (size_t)k * per_iter;
    const float* h1 = pk + (size_t)B * P;
    const float* h2 = h1 + (size_t)B * H;
    // fc3: delta = h2 W3^T + b3 ; d(delta) = dp
    if (gemm(dp, P, 1, h2, H, 0, d_fc_w[2], H, P, H, B, 1.f, st)) return 1;                // dW3 += dp^T h2
    colsum_acc_kernel<<<ceil_div(P, 128), 128, 0, st>>>(dp, B, P, d_fc_b[2]);
    STRAPS_LAUNCH_CHECK();
    if (gemm(dp, P, 0, r->fc_w[2], H, 0, dh2, H, B, H, P, 0.f, st)) return 1;               // dh2 = dp W3
    relu_mask_kernel<<<ceil_div(B * H, 256), 256, 0, st>>>(dh2, h2, (long long)B * H);
    STRAPS_LAUNCH_CHECK();
    // fc2
    if (gemm(dh2, H, 1, h1, H, 0, d_fc_w[1], H, H, H, B, 1.f, st)) return 1;                // dW2 += dh2^T h1
    colsum_acc_kernel<<<ceil_div(H, 128), 128, 0, st>>>(dh2, B, H, d_fc_b[1]);
    STRAPS_LAUNCH_CHECK();
    if (gemm(dh2, H, 0, r->fc_w[1], H, 0, dh1, H, B, H, H, 0.f, st)) return 1;              // dh1 = dh2 W2
    relu_mask_kernel<<<ceil_div(B * H, 256), 256, 0, st>>>(dh1, h1, (long long)B * H);
    STRAPS_LAUNCH_CHECK();
    // fc1 on state = [feat | p_k]
    if (gemm(dh1, H, 1, feat, F, 0, d_fc_w[0], IN, H, F, B, 1.f, st)) return 1;             // dW1[:, :512] += dh1^T feat
    if (gemm(dh1, H, 1, pk, P, 0, d_fc_w[0] + F, IN, H, P, B, 1.f, st)) return 1;           // dW1[:, 512:] += dh1^T p_k
    colsum_acc_kernel<<<ceil_div(H, 128), 128, 0, st>>>(dh1, B, H, d_fc_b[0]);
    STRAPS_LAUNCH_CHECK();
    if (gemm(dh1, H, 0, r->fc_w[0], IN, 0, dsx, IN, B, IN, H, 0.f, st)) return 1;           // d(state) = dh1 W1
    add_block_kernel<<<ceil_div(B * F, 256), 256, 0, st>>>(dsx, IN, d_feat, F, B, F);       // d_feat += d(state)[:, :512]
    STRAPS_LAUNCH_CHECK();
    add_block_kernel<<<ceil_div(B * P, 256), 256, 0, st>>>(dsx + F, IN, dp, P, B, P);       // dp_k = dp_{k+1} + d(state)[:, 512:]
    STRAPS_LAUNCH_CHECK();
  }
  return 0;
}

// torch.optim.Adam semantics (defaults: betas (0.9, 0.999), eps 1e-8, no weight decay) on a flat fp32 bucket.
// grad_scale multiplies the gradient first (1/world_size after a summing all-reduce).  step >= 1.
extern "C" int straps_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int step, float lr,
                                float beta1, float beta2, float eps, float grad_scale, void* stream) {
  STRAPS_CHECK(params && grads && exp_avg && exp_avg_sq, "straps_adam_step: null argument");
  STRAPS_CHECK(step >= 1, "straps_adam_step: step must be >= 1");
  if (n <= 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                                        eps, bc1, bc2, grad_scale);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

extern "C" int straps_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t* step_counter,
                                    float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  STRAPS_CHECK(params && grads && exp_avg && exp_avg_sq && step_counter, "straps_adam_step_dev: null argument");
  if (n <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  adam_tick_kernel<<<1, 1, 0, st>>>(reinterpret_cast<long long*>(step_counter));
  STRAPS_LAUNCH_CHECK();
  adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, reinterpret_cast<const long long*>(step_counter),
                                                             lr, beta1, beta2, eps, grad_scale);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

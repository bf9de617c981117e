// regressor.cu -- ResNet-18 encoder plan + the fp32 CUDA-core convolution path + C-ABI glue.
//
// Replaces ResNet.forward (reference models/resnet.py:201-216, BasicBlock.forward 61-77) and
// SingleInputRegressor.forward (models/regressor.py:43-47).
//
// HBM layout: activations are NHWC (channel-minor) so that an implicit-GEMM K-slice (one filter tap x a run of
// input channels) is one contiguous run per output pixel; the NCHW fp32 input is repacked once into
// [B,256,256,Cpad] (Cpad = C rounded up to 8).  Weights are repacked from OIHW into [(kh,kw,ci)][Cout].
// Eval-mode BatchNorm is folded into per-channel scale/shift applied in the conv epilogue, together with the
// residual add and ReLU, so every activation is written once and read once (plus the residual read).
//
// This file holds the STRAPS_CONV_FP32_SIMT mode: a classic 128x64x8 register-tiled implicit GEMM on the FP32
// pipes (exact fp32 accumulation, used as the on-device cross-check of the tensor-core mode and wherever
// bit-stable fp32 ordering is wanted).  The headline mode, STRAPS_CONV_F16X3_TC, lives in conv_tc.cu.
#include "regressor.h"

namespace straps {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// x NCHW [B,C,H,W] -> y NHWC [B,H,W,CP] (channels >= C zero filled).  Block = (32 w, 8) handles one (b,h,32 w).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, int CP, int H, int W, float* __restrict__ y) {
  __shared__ float tile[32][33];
  const int w0 = blockIdx.x * 32, h = blockIdx.y, b = blockIdx.z;
  for (int c = threadIdx.y; c < CP; c += blockDim.y) {
    const int w = w0 + threadIdx.x;
    tile[c][threadIdx.x] = (c < C && w < W) ? x[(((size_t)b * C + c) * H + h) * W + w] : 0.f;
  }
  __syncthreads();
  float* dst = y + (((size_t)b * H + h) * W + w0) * CP;
  for (int i = threadIdx.y * 32 + threadIdx.x; i < 32 * CP; i += 32 * blockDim.y) {
    const int w = i / CP, c = i % CP;
    if (w0 + w < W) dst[i] = tile[c][w];
  }
}

// y NHWC [B,H,W,C] -> x NCHW (debug / parity hook)
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ y, int C, int H, int W, float* __restrict__ x, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = i % W;
  size_t t = i / W;
  const int h = t % H; t /= H;
  const int c = t % C;
  const size_t b = t / C;
  x[i] = y[((b * H + h) * W + w) * C + c];
}

// OIHW -> [(kh,kw,ci_pad)][cout]
__global__ void pack_w_simt_kernel(const float* __restrict__ w, int cout, int cin, int cin_pad, int ks, float* __restrict__ out) {
  const int total = ks * ks * cin_pad * cout;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = i % cout;
  int t = i / cout;
  const int ci = t % cin_pad;
  t /= cin_pad;
  const int kw = t % ks, kh = t / ks;
  out[i] = (ci < cin) ? w[(((size_t)co * cin + ci) * ks + kh) * ks + kw] : 0.f;
}

// scale = gamma / sqrt(var + eps), shift = beta - mean * scale    (models/resnet.py BatchNorm2d, eval mode)
__global__ void fold_bn_kernel(const float* g, const float* b, const float* m, const float* v, int n, float* scale, float* shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = g[i] / sqrtf(v[i] + 1e-5f);
  scale[i] = s;
  shift[i] = b[i] - m[i] * s;
}

constexpr int BM = 128, BN = 64, BK = 8, APAD = 4;

__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvArgs a) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int M = a.B * a.hout * a.wout;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // A loader: thread -> (pixel = tid/2, 4 channels at (tid%2)*4)
  const int lp = tid >> 1, lh = (tid & 1) * 4;
  const int lm = m0 + lp;
  const bool lvalid = lm < M;
  int lb = 0, loh = 0, low = 0;
  if (lvalid) {
    lb = lm / (a.hout * a.wout);
    const int r = lm % (a.hout * a.wout);
    loh = r / a.wout;
    low = r % a.wout;
  }
  const int cchunks = a.cin / BK;
  const int nk = a.ks * a.ks * cchunks;
  // compute mapping: 16 (n) x 16 (m) threads, 8 pixels x 4 couts each
  const int tn = tid & 15, tm = tid >> 4;

  float4 areg;
  float4 breg;
  auto gload = [&](int kc) {
    const int tap = kc / cchunks, ci0 = (kc % cchunks) * BK;
    const int kh = tap / a.ks, kw = tap % a.ks;
    int ih, iw;
    bool ok = lvalid;
    if (!a.transposed) {
      ih = loh * a.stride - a.pad + kh; iw = low * a.stride - a.pad + kw;
    } else {
      // data gradient: GEMM rows are INPUT pixels, the gathered tensor is dY at ((ih + pad - kh) / stride, ...)
      const int th = loh + a.pad - kh, tw = low + a.pad - kw;
      ok = ok && th >= 0 && tw >= 0 && (th % a.stride) == 0 && (tw % a.stride) == 0;
      ih = th / a.stride; iw = tw / a.stride;
    }
    areg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok && ih >= 0 && ih < a.hin && iw >= 0 && iw < a.win)
      areg = *reinterpret_cast<const float4*>(a.in + (((size_t)lb * a.hin + ih) * a.win + iw) * a.cin + ci0 + lh);
    if (tid < 128) {
      const int kr = tid >> 4, c4 = (tid & 15) * 4;
      breg = *reinterpret_cast<const float4*>(a.w + ((size_t)(tap * a.cin + ci0 + kr)) * a.cout + n0 + c4);
    }
  };
  auto sstore = [&](int buf) {
    As[buf][lh + 0][lp] = areg.x; As[buf][lh + 1][lp] = areg.y; As[buf][lh + 2][lp] = areg.z; As[buf][lh + 3][lp] = areg.w;
    if (tid < 128) *reinterpret_cast<float4*>(&Bs[buf][tid >> 4][(tid & 15) * 4]) = breg;
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nk) gload(kc + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kc + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  const int co = n0 + tn * 4;
  const float4 sc = *reinterpret_cast<const float4*>(a.scale + co);
  const float4 sh = *reinterpret_cast<const float4*>(a.shift + co);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tm * 8 + i;
    if (m >= M) continue;
    float4 y = make_float4(fmaf(acc[i][0], sc.x, sh.x), fmaf(acc[i][1], sc.y, sh.y), fmaf(acc[i][2], sc.z, sh.z),
                           fmaf(acc[i][3], sc.w, sh.w));
    if (a.res) {
      const float4 r = *reinterpret_cast<const float4*>(a.res + (size_t)m * a.cout + co);
      y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
    }
    if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
    *reinterpret_cast<float4*>(a.out + (size_t)m * a.cout + co) = y;
  }
}

// 3x3 stride-2 pad-1 max pool, NHWC fp32, one thread per (pixel, 4 channels)   (models/resnet.py:149,206)
__global__ void maxpool_nhwc_kernel(const float* __restrict__ in, int B, int H, int W, int C, float* __restrict__ out) {
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
  *reinterpret_cast<float4*>(out + ((b * HO + oh) * WO + ow) * C + c4 * 4) = m;
}

// global average pool NHWC [B,HW,C] -> [B,C]   (models/resnet.py:213-214)
__global__ void avgpool_nhwc_kernel(const float* __restrict__ in, int B, int HW, int C, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C, b = i / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += in[((size_t)b * HW + p) * C + c];
  out[i] = s / (float)HW;
}

}  // namespace straps

using namespace straps;

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
static int add_buf(straps_regressor* r, const char* name, int h, int w, int c) {
  ActBuf b;
  b.name = name; b.h = h; b.w = w; b.c = c;
  b.bytes = (size_t)r->max_batch * h * w * c * sizeof(float);
  b.bytes = (b.bytes + 1023) & ~(size_t)1023;
  b.offset = r->ws_bytes;
  r->ws_bytes += b.bytes;
  r->bufs.push_back(b);
  return (int)r->bufs.size() - 1;
}

static void set_conv(ConvSpec& c, const std::string& name, int cin, int cin_pad, int cout, int ks, int stride, int pad,
                     int hin, int in_buf, int out_buf, int res_buf, int relu) {
  c.name = name; c.cin = cin; c.cin_pad = cin_pad; c.cout = cout; c.ksize = ks; c.stride = stride; c.pad = pad;
  c.hin = c.win = hin;
  c.hout = c.wout = (hin + 2 * pad - ks) / stride + 1;
  c.in_buf = in_buf; c.out_buf = out_buf; c.res_buf = res_buf; c.relu = relu;
  c.w_simt = c.scale = c.shift = nullptr; c.w_hi = c.w_lo = nullptr; c.k_eff = 0;
}

extern "C" int straps_regressor_create(straps_regressor_t** out, int c_in, int max_batch) {
  STRAPS_CHECK(out, "straps_regressor_create: null out");
  STRAPS_CHECK(c_in >= 1 && c_in <= 32, "straps_regressor_create: c_in=%d outside [1,32]", c_in);
  STRAPS_CHECK(max_batch >= 1 && max_batch <= 4096, "straps_regressor_create: max_batch=%d outside [1,4096]", max_batch);
  straps_regressor* r = new straps_regressor();
  r->c_in = c_in; r->c_in_pad = (c_in + 7) / 8 * 8; r->max_batch = max_batch;
  r->ws = nullptr; r->ws_bytes = 0; r->wpool = nullptr; r->tc = nullptr; r->loaded = 0; r->last_mode = -1; r->dirty_bn = r->dirty_simt = r->dirty_tc = 1;
  r->train = nullptr; r->ief = nullptr;
  // activation buffers (NHWC)
  r->buf_xin = add_buf(r, "", IMG, IMG, 32);        // sized for the widest packed-input layout either mode uses
  r->buf_stem = add_buf(r, "stem", 128, 128, 64);
  r->buf_pool = add_buf(r, "pool", 64, 64, 64);
  int ci = 0;
  set_conv(r->conv[ci++], "conv1", c_in, r->c_in_pad, 64, 7, 2, 3, IMG, r->buf_xin, r->buf_stem, -1, 1);
  int cur = r->buf_pool, hw = 64, cin = 64;
  const int widths[4] = {64, 128, 256, 512};
  for (int L = 0; L < 4; ++L)
    for (int blk = 0; blk < 2; ++blk) {
      const int cout = widths[L];
      const int stride = (blk == 0 && L > 0) ? 2 : 1;
      const bool ds = (blk == 0 && L > 0);
      const int hout = hw / stride;
      char nm[32];
      snprintf(nm, sizeof(nm), "layer%d.%d", L + 1, blk);
      const int ba = add_buf(r, (std::string(nm) + ".a").c_str(), hout, hout, cout);
      int bds = -1;
      if (ds) bds = add_buf(r, (std::string(nm) + ".ds").c_str(), hout, hout, cout);
      const int bo = add_buf(r, nm, hout, hout, cout);
      set_conv(r->conv[ci++], std::string(nm) + ".conv1", cin, cin, cout, 3, stride, 1, hw, cur, ba, -1, 1);
      if (ds) {
        // order in the conv list follows the state_dict: conv1, conv2, downsample.0 -- keep that order
        set_conv(r->conv[ci++], std::string(nm) + ".conv2", cout, cout, cout, 3, 1, 1, hout, ba, bo, bds, 1);
        set_conv(r->conv[ci++], std::string(nm) + ".downsample.0", cin, cin, cout, 1, stride, 0, hw, cur, bds, -1, 0);
      } else {
        set_conv(r->conv[ci++], std::string(nm) + ".conv2", cout, cout, cout, 3, 1, 1, hout, ba, bo, cur, 1);
      }
      cur = bo; hw = hout; cin = cout;
    }
  r->buf_final = cur;
  // packed weights pool
  size_t wfloats = 0;
  for (int i = 0; i < NCONV; ++i) {
    const ConvSpec& c = r->conv[i];
    wfloats += (size_t)c.ksize * c.ksize * c.cin_pad * c.cout + 2 * (size_t)c.cout;
    wfloats = (wfloats + 63) & ~(size_t)63;
  }
  const size_t ief_floats = (size_t)IEF_IN * IEF_H + (size_t)IEF_H * IEF_H + (size_t)IEF_H * IEF_OUT_PAD + 2 * IEF_H + IEF_OUT_PAD + 256 +
                            (size_t)max_batch * STRAPS_FEAT_DIM;
  r->wpool_bytes = (wfloats + ief_floats) * sizeof(float);
  if (cudaMalloc(&r->wpool, r->wpool_bytes) != cudaSuccess || cudaMalloc(&r->ws, r->ws_bytes) != cudaSuccess) {
    set_error("straps_regressor_create: cudaMalloc of %zu + %zu bytes failed: %s", r->wpool_bytes, r->ws_bytes,
              cudaGetErrorString(cudaGetLastError()));
    straps_regressor_destroy(r);
    return 1;
  }
  float* p = r->wpool;
  for (int i = 0; i < NCONV; ++i) {
    ConvSpec& c = r->conv[i];
    float* start = p;
    c.w_simt = p; p += (size_t)c.ksize * c.ksize * c.cin_pad * c.cout;
    c.scale = p; p += c.cout;
    c.shift = p; p += c.cout;
    size_t used = (size_t)(p - start);
    p = start + ((used + 63) & ~(size_t)63);
  }
  r->w1t = p; p += (size_t)IEF_IN * IEF_H;
  r->w2t = p; p += (size_t)IEF_H * IEF_H;
  r->w3t = p; p += (size_t)IEF_H * IEF_OUT_PAD;
  r->b1 = p; p += IEF_H;
  r->b2 = p; p += IEF_H;
  r->b3 = p; p += IEF_OUT_PAD;
  r->init = p; p += 256;
  r->feat_scratch = p; p += (size_t)max_batch * STRAPS_FEAT_DIM;
  if (tc_create(r) || ief_create(r)) { straps_regressor_destroy(r); return 1; }
  *out = r;
  return 0;
}

extern "C" void straps_regressor_destroy(straps_regressor_t* r) {
  if (!r) return;
  tc_destroy(r);
  train_destroy(r);
  ief_destroy(r);
  if (r->ws) cudaFree(r->ws);
  if (r->wpool) cudaFree(r->wpool);
  delete r;
}

extern "C" const char* straps_regressor_conv_name(const straps_regressor_t* r, int i) {
  return (r && i >= 0 && i < NCONV) ? r->conv[i].name.c_str() : nullptr;
}

extern "C" size_t straps_regressor_workspace_bytes(const straps_regressor_t* r) { return r ? r->ws_bytes + r->wpool_bytes : 0; }

extern "C" int straps_regressor_load(straps_regressor_t* r, const float* const* conv_w, const float* const* bn,
                                     const float* const* fc_w, const float* const* fc_b, const float* init_params,
                                     void* stream) {
  STRAPS_CHECK(r && conv_w && bn && fc_w && fc_b && init_params, "straps_regressor_load: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < NCONV; ++i) {
    ConvSpec& c = r->conv[i];
    STRAPS_CHECK(conv_w[i] && bn[4 * i] && bn[4 * i + 1] && bn[4 * i + 2] && bn[4 * i + 3],
                 "straps_regressor_load: null tensor for conv %d (%s)", i, c.name.c_str());
    c.w_oihw = conv_w[i];
    c.gamma = bn[4 * i]; c.beta = bn[4 * i + 1];
    c.rmean = const_cast<float*>(bn[4 * i + 2]); c.rvar = const_cast<float*>(bn[4 * i + 3]);
  }
  for (int i = 0; i < 3; ++i) { r->fc_w[i] = fc_w[i]; r->fc_b[i] = fc_b[i]; }
  if (ief_pack(r, fc_w, fc_b, init_params, st)) return 1;
  // The convolution weights are only REFERENCED here: the packed inference copies (folded BatchNorm, fp32 / fp16-split layouts)
  // are rebuilt by the first forward that needs them, so a training loop -- which reloads after every optimiser step and packs
  // its own un-folded copies -- does not pay for them.  The tensors must stay alive and in place until the next load.
  r->dirty_bn = r->dirty_simt = r->dirty_tc = 1;
  r->loaded = 1;
  return 0;
}

namespace straps {
int ensure_packed(straps_regressor* r, int conv_mode, int need_folded_bn, cudaStream_t st) {
  if (need_folded_bn && r->dirty_bn) {
    for (int i = 0; i < NCONV; ++i) {
      ConvSpec& c = r->conv[i];
      fold_bn_kernel<<<ceil_div(c.cout, 128), 128, 0, st>>>(c.gamma, c.beta, c.rmean, c.rvar, c.cout, c.scale, c.shift);
      STRAPS_LAUNCH_CHECK();
    }
    r->dirty_bn = 0;
    r->dirty_tc = 1;      // the tensor-core copy folds the BN scale into the weights
  }
  if (conv_mode == STRAPS_CONV_FP32_SIMT && r->dirty_simt) {
    for (int i = 0; i < NCONV; ++i) {
      ConvSpec& c = r->conv[i];
      const int total = c.ksize * c.ksize * c.cin_pad * c.cout;
      pack_w_simt_kernel<<<ceil_div(total, 256), 256, 0, st>>>(c.w_oihw, c.cout, c.cin, c.cin_pad, c.ksize, c.w_simt);
      STRAPS_LAUNCH_CHECK();
    }
    r->dirty_simt = 0;
  }
  if (conv_mode == STRAPS_CONV_F16X3_TC && need_folded_bn && r->dirty_tc) {
    const float* w[NCONV];
    for (int i = 0; i < NCONV; ++i) w[i] = r->conv[i].w_oihw;
    if (tc_pack(r, w, st)) return 1;
    r->dirty_tc = 0;
  }
  return 0;
}
}  // namespace straps

namespace straps {
int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
  dim3 grid(ceil_div(a.B * a.hout * a.wout, BM), a.cout / BN);
  conv_simt_kernel<<<grid, 256, 0, st>>>(a);
  STRAPS_LAUNCH_CHECK();
  return 0;
}
}  // namespace straps

namespace straps {
int launch_nchw_to_nhwc(straps_regressor* r, const float* x, int B, cudaStream_t st) {
  nchw_to_nhwc_kernel<<<dim3(IMG / 32, IMG, B), dim3(32, 8), 0, st>>>(x, r->c_in, r->c_in_pad, IMG, IMG, act_ptr(r, r->buf_xin));
  STRAPS_LAUNCH_CHECK();
  return 0;
}
int launch_maxpool(straps_regressor* r, int B, cudaStream_t st) {
  const size_t n = (size_t)B * 64 * 64 * 16;
  maxpool_nhwc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(act_ptr(r, r->buf_stem), B, 128, 128, 64, act_ptr(r, r->buf_pool));
  STRAPS_LAUNCH_CHECK();
  return 0;
}
int launch_avgpool(straps_regressor* r, int B, float* feat, cudaStream_t st) {
  avgpool_nhwc_kernel<<<ceil_div(B * 512, 256), 256, 0, st>>>(act_ptr(r, r->buf_final), B, 64, 512, feat);
  STRAPS_LAUNCH_CHECK();
  return 0;
}
}  // namespace straps

static int run_conv_simt(const straps_regressor* r, const ConvSpec& c, int B, cudaStream_t st) {
  ConvArgs a;
  a.transposed = 0;
  a.in = act_ptr(r, c.in_buf); a.w = c.w_simt; a.scale = c.scale; a.shift = c.shift;
  a.res = c.res_buf >= 0 ? act_ptr(r, c.res_buf) : nullptr;
  a.out = act_ptr(r, c.out_buf);
  a.B = B; a.hin = c.hin; a.win = c.win; a.cin = c.cin_pad; a.hout = c.hout; a.wout = c.wout; a.cout = c.cout;
  a.ks = c.ksize; a.stride = c.stride; a.pad = c.pad; a.relu = c.relu;
  return launch_conv_simt(a, st);
}

static int encoder_forward_simt(straps_regressor* r, const float* x, int B, float* feat, cudaStream_t st) {
  if (launch_nchw_to_nhwc(r, x, B, st)) return 1;
  if (run_conv_simt(r, r->conv[0], B, st)) return 1;
  if (launch_maxpool(r, B, st)) return 1;
  int i = 1;
  while (i < NCONV) {
    // a block is conv1, conv2[, downsample]; the downsample branch must run before conv2 consumes it
    const bool ds = (i + 2 < NCONV) && r->conv[i + 2].ksize == 1;
    if (run_conv_simt(r, r->conv[i], B, st)) return 1;
    if (ds && run_conv_simt(r, r->conv[i + 2], B, st)) return 1;
    if (run_conv_simt(r, r->conv[i + 1], B, st)) return 1;
    i += ds ? 3 : 2;
  }
  return launch_avgpool(r, B, feat, st);
}

extern "C" int straps_encoder_forward(straps_regressor_t* r, const float* x, int batch, int conv_mode, float* feat,
                                      void* stream) {
  STRAPS_CHECK(r && x && feat, "straps_encoder_forward: null argument");
  STRAPS_CHECK(r->loaded, "straps_encoder_forward: weights not loaded (call straps_regressor_load first)");
  STRAPS_CHECK(batch >= 1 && batch <= r->max_batch, "straps_encoder_forward: batch %d outside [1,%d]", batch, r->max_batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  r->last_mode = conv_mode;
  STRAPS_CHECK(conv_mode == STRAPS_CONV_FP32_SIMT || conv_mode == STRAPS_CONV_F16X3_TC, "straps_encoder_forward: unknown conv_mode %d", conv_mode);
  if (ensure_packed(r, conv_mode, 1, st)) return 1;
  if (conv_mode == STRAPS_CONV_FP32_SIMT) return encoder_forward_simt(r, x, batch, feat, st);
  if (conv_mode == STRAPS_CONV_F16X3_TC) return tc_encoder_forward(r, x, batch, feat, st);
  STRAPS_CHECK(false, "straps_encoder_forward: unknown conv_mode %d", conv_mode);
}

extern "C" int straps_ief_forward(straps_regressor_t* r, const float* feat, int batch, int iters, float* params,
                                  void* stream) {
  STRAPS_CHECK(r && feat && params, "straps_ief_forward: null argument");
  STRAPS_CHECK(r->loaded, "straps_ief_forward: weights not loaded");
  STRAPS_CHECK(batch >= 1 && iters >= 0, "straps_ief_forward: bad batch/iters %d/%d", batch, iters);
  return ief_launch(r, feat, batch, iters, params, static_cast<cudaStream_t>(stream));
}

extern "C" int straps_ief_forward_train(straps_regressor_t* r, const float* feat, int batch, int iters, float* params, float* saved,
                                        void* stream) {
  STRAPS_CHECK(r && feat && params && saved, "straps_ief_forward_train: null argument");
  STRAPS_CHECK(r->loaded, "straps_ief_forward_train: weights not loaded");
  STRAPS_CHECK(batch >= 1 && iters >= 0, "straps_ief_forward_train: bad batch/iters %d/%d", batch, iters);
  return ief_launch_train(r, feat, batch, iters, params, saved, static_cast<cudaStream_t>(stream));
}

extern "C" int straps_regressor_forward(straps_regressor_t* r, const float* x, int batch, int conv_mode, int iters,
                                        float* feat_or_null, float* params, void* stream) {
  STRAPS_CHECK(r && x && params, "straps_regressor_forward: null argument");
  float* feat = feat_or_null ? feat_or_null : r->feat_scratch;
  int rc = straps_encoder_forward(r, x, batch, conv_mode, feat, stream);
  if (rc) return rc;
  return straps_ief_forward(r, feat, batch, iters, params, stream);
}

extern "C" int straps_regressor_forward_from_labels(straps_regressor_t* r, const float* seg_labels, const float* joints2d, int num_joints,
                                                    const float* table, int half_size, int batch, int iters, float* feat_or_null,
                                                    float* params, void* stream) {
  STRAPS_CHECK(r && seg_labels && joints2d && table && params, "straps_regressor_forward_from_labels: null argument");
  STRAPS_CHECK(r->loaded, "straps_regressor_forward_from_labels: weights not loaded (call straps_regressor_load first)");
  STRAPS_CHECK(batch >= 1 && batch <= r->max_batch, "straps_regressor_forward_from_labels: batch %d outside [1,%d]", batch, r->max_batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  r->last_mode = STRAPS_CONV_F16X3_TC;
  if (ensure_packed(r, STRAPS_CONV_F16X3_TC, 1, st)) return 1;
  float* feat = feat_or_null ? feat_or_null : r->feat_scratch;
  if (tc_encoder_forward_from_labels(r, seg_labels, joints2d, num_joints, table, half_size, batch, feat, st)) return 1;
  return straps_ief_forward(r, feat, batch, iters, params, stream);
}

extern "C" int straps_encoder_read_activation(straps_regressor_t* r, const char* name, int batch, float* out, int64_t* n,
                                              void* stream) {
  STRAPS_CHECK(r && name && out, "straps_encoder_read_activation: null argument");
  if (std::string(name) == "grad:conv1") {   // parity hook: dY of conv1 (the last raw-output gradient the encoder backward produced)
    const float* g = train_last_draw(r);
    STRAPS_CHECK(g, "straps_encoder_read_activation: no encoder backward has run");
    const size_t total = (size_t)batch * 64 * 128 * 128;
    if (n) *n = (int64_t)total;
    nhwc_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, 64, 128, 128, out, total);
    STRAPS_LAUNCH_CHECK();
    return 0;
  }
  int id = -1;
  for (size_t i = 0; i < r->bufs.size(); ++i)
    if (!r->bufs[i].name.empty() && r->bufs[i].name == name) id = (int)i;
  STRAPS_CHECK(id >= 0, "straps_encoder_read_activation: unknown activation '%s'", name);
  const ActBuf& b = r->bufs[id];
  const size_t total = (size_t)batch * b.c * b.h * b.w;
  if (n) *n = (int64_t)total;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (r->last_mode == STRAPS_CONV_F16X3_TC) return tc_read_activation(r, id, batch, out, st);
  nhwc_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(act_ptr(r, id), b.c, b.h, b.w, out, total);
  STRAPS_LAUNCH_CHECK();
  return 0;
}

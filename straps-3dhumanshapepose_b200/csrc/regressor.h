// regressor.h -- internal structures shared by the encoder / IEF translation units.
#pragma once
#include "common.cuh"
#include "../../include/straps_b200.h"
#include <cuda_fp16.h>
#include <vector>
#include <string>

namespace straps {

constexpr int IMG = 256;
constexpr int NCONV = 20;
constexpr int IEF_IN = STRAPS_FEAT_DIM + STRAPS_IEF_PARAMS;   // 669
constexpr int IEF_H = 512;
constexpr int IEF_OUT_PAD = 160;

// One convolution (+ folded BatchNorm [+ residual] [+ ReLU]) of the ResNet-18 encoder.
struct ConvSpec {
  int cin, cin_pad, cout, ksize, stride, pad;
  int hin, win, hout, wout;
  int in_buf, out_buf, res_buf;   // activation buffer ids (-1 = none)
  int relu;
  std::string name;
  // device pointers
  float* w_simt;    // [(kh,kw,ci_pad)][cout] fp32
  float* scale;     // [cout] gamma / sqrt(var + eps)
  float* shift;     // [cout] beta - mean * scale
  // training path (train.cu): the PyTorch-owned tensors of the last straps_regressor_load
  const float* w_oihw;
  const float* gamma;
  const float* beta;
  float* rmean;
  float* rvar;
  // tensor-core path (conv_tc.cu)
  void* w_hi;       // [cout][k_eff] fp16, BN scale folded
  void* w_lo;
  int k_eff;        // padded reduction length of the tensor-core layout
};

// arguments of the fp32 CUDA-core implicit-GEMM convolution (regressor.cu), also used for the data gradient
struct ConvArgs {
  const float* in;     // NHWC [B,hin,win,cin]   (the gathered tensor: activations, or dY when transposed)
  const float* w;      // [(kh,kw,cin)][cout]
  const float* scale;
  const float* shift;
  const float* res;    // NHWC [B,hout,wout,cout] added before the ReLU, or null
  float* out;          // NHWC [B,hout,wout,cout]
  int B, hin, win, cin, hout, wout, cout, ks, stride, pad, relu;
  int transposed;      // 1 = data-gradient gather (rows are input pixels, `in` is dY)
};

// 2-term fp16 split used by the tensor-core path (conv_tc.cu): x = hi + lo * 2^-11.
// The low term is stored SCALED by 2^11 (since round 2): |x - hi| <= 2^-11 |x|, so lo has the magnitude of x instead of sitting 11
// binades lower -- it stays fp16-NORMAL wherever hi is (a plain lo = fp16(x - hi) goes subnormal for |x| < 0.125 and costs 3e-8
// ABSOLUTE, i.e. 3e-5 relative on a layer whose activations peak at 1e-3: round-1 review) and cannot overflow.  Every lo plane
// (activations, weights, gradients) follows this convention, so both cross terms A_hi.W_lo and A_lo.W_hi carry the same factor 2^11
// and the epilogues combine the two accumulators as acc_hi + 2^-11 acc_lo.  For data that was normal before, results are unchanged
// bit for bit (power-of-two scalings are exact).
constexpr float LO_SCALE = 2048.f, LO_UNSCALE = 1.f / 2048.f;
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  lo = __float2half_rn(fminf(fmaxf((v - __half2float(hi)) * LO_SCALE, -65504.f), 65504.f));
}
// x back from its two halves
__device__ __forceinline__ float unsplit_f16(__half hi, __half lo) { return fmaf(__half2float(lo), LO_UNSCALE, __half2float(hi)); }
__device__ __forceinline__ uint32_t pack_f16(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ float f16lo_to_f(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
__device__ __forceinline__ float f16hi_to_f(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }

struct ActBuf {
  std::string name;   // "" = internal
  int h, w, c;        // NHWC per body
  size_t offset;      // bytes into the workspace (for max_batch bodies)
  size_t bytes;
};

}  // namespace straps

struct straps_regressor {
  int c_in, c_in_pad, max_batch;
  straps::ConvSpec conv[straps::NCONV];
  std::vector<straps::ActBuf> bufs;
  int buf_xin, buf_stem, buf_pool, buf_final;
  unsigned char* ws;       // activation workspace
  size_t ws_bytes;
  float* wpool;            // packed weights pool
  size_t wpool_bytes;
  // IEF
  float *w1t, *w2t, *w3t, *b1, *b2, *b3, *init, *feat_scratch;
  // tensor-core path state (opaque, owned by conv_tc.cu)
  void* tc;
  void* ief;               // IEF weight slices (ief.cu)
  void* train;             // training workspace (train.cu), allocated on first use
  const float* fc_w[3];    // PyTorch-owned IEF weights / biases of the last load (nn.Linear layout)
  const float* fc_b[3];
  int loaded;
  int last_mode;
  int dirty_bn, dirty_simt, dirty_tc;   // inference copies (folded BN, packed conv weights) are rebuilt lazily after a load
};

namespace straps {
int ief_launch(const straps_regressor* r, const float* feat, int batch, int iters, float* params, cudaStream_t st);
int ief_create(straps_regressor* r);
void ief_destroy(straps_regressor* r);
int ief_pack(straps_regressor* r, const float* const* fc_w, const float* const* fc_b, const float* init, cudaStream_t st);
int launch_conv_simt(const ConvArgs& a, cudaStream_t st);
int ensure_packed(straps_regressor* r, int conv_mode, int need_folded_bn, cudaStream_t st);   // lazy repack of the inference weight copies
int launch_nchw_to_nhwc(straps_regressor* r, const float* x, int B, cudaStream_t st);
int launch_maxpool(straps_regressor* r, int B, cudaStream_t st);
int launch_avgpool(straps_regressor* r, int B, float* feat, cudaStream_t st);
int ief_launch_train(const straps_regressor* r, const float* feat, int batch, int iters, float* params, float* saved, cudaStream_t st);
void train_destroy(straps_regressor* r);
const float* train_last_draw(const straps_regressor* r);   // scratch holding the last d(raw conv output) of the backward (conv1's)
// tensor-core encoder (conv_tc.cu)
int tc_create(straps_regressor* r);
void tc_destroy(straps_regressor* r);
int tc_pack(straps_regressor* r, const float* const* conv_w, cudaStream_t st);
int tc_encoder_forward(straps_regressor* r, const float* x, int batch, float* feat, cudaStream_t st);
int tc_encoder_forward_from_labels(straps_regressor* r, const float* seg, const float* joints2d, int num_joints, const float* table,
                                   int half_size, int batch, float* feat, cudaStream_t st);
int tc_read_activation(straps_regressor* r, int buf, int batch, float* out, cudaStream_t st);
// tensor-core training path (conv_tc.cu): forward convs with un-folded weights and data gradients as flipped-tap convolutions
int tc_train_begin(straps_regressor* r, int batch, cudaStream_t st);                      // workspace, forward weight pack, tensor maps
int tc_train_pack_input(straps_regressor* r, const float* x, int batch, cudaStream_t st);
int tc_train_unpack_input(straps_regressor* r, int batch, cudaStream_t st);               // conv1 input planes -> fp32 NHWC input buffer
__half* tc_train_plane(straps_regressor* r, int buf, int lo);                              // split planes of an activation buffer (or null)
int tc_train_split_act(straps_regressor* r, int buf, int batch, cudaStream_t st);         // fp32 activation buffer -> its split planes
int tc_train_conv_fwd(straps_regressor* r, int ci, int batch, float* raw, cudaStream_t st);
int tc_train_pack_dgrad(straps_regressor* r, cudaStream_t st);
int tc_train_split_dy(straps_regressor* r, int ci, int batch, const float* dy, int upsample, cudaStream_t st);   // scaled split of dY
int tc_train_conv_wgrad(straps_regressor* r, int ci, int batch, float* dw_oihw, cudaStream_t st);                // needs the plain split
unsigned* tc_train_dy_max(straps_regressor* r, int ci);                                    // device word of conv ci: bits of max |dY| (atomicMax target)
void tc_train_dy_planes(straps_regressor* r, int ci, __half** hi, __half** lo);              // split planes bn_backward writes dX into
void tc_train_dgrad_scales(straps_regressor* r, int ci, const float** w_unscale, int* n, float** out_unscale);
int tc_train_unpack_all(straps_regressor* r, cudaStream_t st);                             // end of a backward pass: all weight gradients -> OIHW
int tc_train_conv_dgrad(straps_regressor* r, int ci, int batch, const float* add, float* gin, cudaStream_t st);  // needs the split
inline float* act_ptr(const straps_regressor* r, int buf) { return reinterpret_cast<float*>(r->ws + r->bufs[buf].offset); }
}  // namespace straps

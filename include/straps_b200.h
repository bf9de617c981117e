/*
 * straps_b200.h -- C ABI of libstraps_b200.so (hand-written sm_100a CUDA for the STRAPS hot path).
 *
 * The reference (akashsengupta1997/STRAPS-3DHumanShapePose) has NO FFI / plugin layer: its boundary
 * is the Python module surface.  Each entry point below replaces the torch-op
 * cluster behind one reference interface, cited as file:line of the reference tree.  The Python
 * drop-in modules under straps-3dhumanshapepose_b200/{models,utils,losses}/ bind these with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers + explicit sizes; no torch / C++ types cross the ABI.
 *   - "dev" pointers are CUDA device pointers on the current device, "host" pointers are CPU memory.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Nothing synchronises
 *     except the *_create functions (they upload constants) and straps_*_destroy.
 *   - every function returns 0 on success, non-zero on failure; straps_last_error() then returns a
 *     thread-local message.  No exceptions cross the ABI.
 *   - tensors are fp32, C-contiguous unless a stride (in elements) is passed.
 */
#ifndef STRAPS_B200_H_
#define STRAPS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden; only the entry points below are exported */
#pragma GCC visibility push(default)

#define STRAPS_NUM_VERTS 6890
#define STRAPS_NUM_JOINTS 24
#define STRAPS_NUM_BETAS 10
#define STRAPS_NUM_EXTRA_PICKS 21   /* smplx VertexJointSelector picks */
#define STRAPS_NUM_EXTRA_ROWS 45    /* J_regressor_extra 9 + cocoplus 19 + h36m 17 (models/smpl_official.py:17-25) */
#define STRAPS_NUM_SUPERSET_JOINTS 90
#define STRAPS_IEF_PARAMS 157       /* 3 cam + 24*6 pose + 10 shape (models/regressor.py:25-26) */
#define STRAPS_FEAT_DIM 512

const char* straps_last_error(void);
int straps_abi_version(void);
/* Number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
unsigned long long straps_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * SMPL body model -- replaces smplx.SMPL.__init__/forward + models/smpl_official.py:15-41.
 * ------------------------------------------------------------------------------------------------ */
typedef struct straps_smpl straps_smpl_t;

/* Upload + pre-reduce the model constants (all HOST pointers, fp32 row-major):
 *   v_template [6890,3]; shapedirs [6890,3,10]; posedirs [207,20670] (smplx layout: pose-feature major,
 *   inner axis ordered (vertex, xyz)); J_regressor [24,6890]; lbs_weights [6890,24]; parents int64[24]
 *   (parents[0] = -1); extra_regressors [45,6890] = rows of J_regressor_extra, cocoplus, h36m stacked in
 *   that order; extra_pick_idx int64[21] vertex ids.  Synchronous. */
int straps_smpl_create(straps_smpl_t** out,
                       const float* v_template, const float* shapedirs, const float* posedirs,
                       const float* J_regressor, const float* lbs_weights, const int64_t* parents,
                       const float* extra_regressors, const int64_t* extra_pick_idx);
void straps_smpl_destroy(straps_smpl_t* m);
/* 1 if every vertex has <= 4 non-zero skinning weights (fast path), else 0 (dense path). */
int straps_smpl_is_sparse4(const straps_smpl_t* m);

/* SMPL forward (smplx lbs() + VertexJointSelector + the three extra regressors):
 *   pose2rot == 0: global_orient dev [B,1,3,3] (row stride go_stride floats), body_pose dev [B,23,3,3]
 *                  (row stride bp_stride) -- rotation matrices  (train/...:196-199).
 *   pose2rot == 1: global_orient dev [B,3], body_pose dev [B,69] axis-angle (models/smpl_official.py:29
 *                  default path, train/...:144,206).
 *   betas dev [B,10] (row stride betas_stride); transl dev [B,3] or NULL.
 * Outputs: vertices dev [B,6890,3]; joints dev [B,90,3] (24 posed + 21 picks + 9 + 19 + 17). */
int straps_smpl_forward(const straps_smpl_t* m,
                        const float* global_orient, int64_t go_stride,
                        const float* body_pose, int64_t bp_stride,
                        const float* betas, int64_t betas_stride,
                        const float* transl, int batch, int pose2rot,
                        float* vertices, float* joints, void* stream);

/* ---- training path of the SMPL side (backward of train/train_synthetic_otf_rendering.py:193-205) ---- */
/* Same as straps_smpl_forward with pose2rot == 0 and rotmats dev [B,24,3,3] contiguous, and additionally saves what
 * the backward needs: v_posed dev [B,6890,3] and the skinning transforms A dev [B,24,12]. */
int straps_smpl_forward_train(const straps_smpl_t* m, const float* rotmats, const float* betas, int batch,
                              float* vertices, float* joints, float* save_vposed, float* save_A, void* stream);
/* Backward: g_joints dev [B,90,3]; gv_work dev [B,6890,3] = a WRITABLE COPY of the vertex gradient (the joint
 * contributions are accumulated into it); scratch dev [B*(24*12+208+10)] floats.
 * Outputs: d_rotmats dev [B,24,3,3], d_betas dev [B,10]. */
int straps_smpl_backward(const straps_smpl_t* m, const float* rotmats, const float* betas, const float* v_posed,
                         const float* A, const float* g_joints, float* gv_work, float* scratch, int batch,
                         float* d_rotmats, float* d_betas, void* stream);
/* Backward of rot6d_to_rotmat: x6 dev [n,6], dR dev [n,3,3] -> dx6 dev [n,6]. */
int straps_rot6d_backward(const float* x6, const float* dR, int64_t n, float* dx6, void* stream);
/* Backward of the weak-perspective projection: g_out dev [B,N,2] -> d_points dev [B,N,3], d_cam dev [B,3]. */
int straps_orthographic_project_backward(const float* points, const float* cam, int64_t cam_stride,
                                         const float* g_out, int batch, int npoints, float* d_points, float* d_cam,
                                         void* stream);

/* utils/rigid_transform_utils.py:27-41 -- x6 dev [n,6] (interleaved a1/a2) -> R dev [n,3,3]. */
int straps_rot6d_to_rotmat(const float* x6, int64_t n, float* R, void* stream);

/* utils/cam_utils.py:5-26 -- points dev [B,N,3], cam dev [B,3] (row stride cam_stride) -> out dev [B,N,2]. */
int straps_orthographic_project(const float* points, const float* cam, int64_t cam_stride,
                                int batch, int npoints, float* out, void* stream);

/* ---- proxy-representation synthesis on the device (the step right before the path) ----
 * utils/label_conversions.py:90-127 -- joints2d dev [B,J,2] (pixels) -> heatmaps dev [B,J,img_wh,img_wh]: zero, with the
 * (2*half_size)^2 truncated Gaussian `table` (dev, row-major, computed by the caller exactly as the reference does)
 * pasted at the truncated-to-int joint position with the reference's clipping rules. */
int straps_joints2d_to_heatmaps(const float* joints2d, int batch, int num_joints, int img_wh, int half_size,
                                const float* table, float* heatmaps, void* stream);
/* utils/label_conversions.py:48-55 -- out[i] = (labels[i] != 0) ? 1 : 0  (fp32 in, fp32 out). */
int straps_multiclass_to_binary(const float* labels, int64_t n, float* out, void* stream);

/* ---- target side of the synthetic training loop (train/train_synthetic_otf_rendering.py:121-145) ----
 * smplx.lbs.batch_rodrigues (augmentation/smpl_augmentation.py:55-58, train/...:190,301,322, predict/predict_3D.py:134):
 * rot_vecs dev [n,3] axis-angle -> R dev [n,3,3]; angle = ||r + 1e-8||. */
int straps_batch_rodrigues(const float* rot_vecs, int64_t n, float* R, void* stream);
/* Its gradient (smplx builds this graph for axis-angle input, pose2rot=True: models/smpl_official.py:29 via the
 * `smpl_model(betas=pred_shape)` call of train/train_synthetic_otf_rendering.py:206): dR dev [n,3,3] -> d_rot_vecs dev [n,3]. */
int straps_batch_rodrigues_backward(const float* rot_vecs, const float* dR, int64_t n, float* d_rot_vecs, void* stream);
/* utils/cam_utils.py:40-71 perspective_project_torch: points dev [B,N,3], rotation dev [B,3,3], translation dev [B,3],
 * cam_K dev [B,3,3] -> out dev [B,N,2] = (K ((R p + t) / (R p + t).z))[:2]. */
int straps_perspective_project(const float* points, const float* rotation, const float* translation, const float* cam_K,
                               int batch, int npoints, float* out, void* stream);
/* The affine part of augmentation/smpl_augmentation.py:6-24 and augmentation/cam_augmentation.py:4-14:
 * out[r,c] = (noise[r,c] * mul[c] + add[c]) + base[r*base_stride + c], every operation rounded separately (no FMA), i.e.
 * bit-identical to the reference's torch ops on the same draws.  add may be NULL; base_stride 0 broadcasts one row. */
int straps_scale_shift(const float* noise, const float* mul, const float* add, const float* base, int64_t base_stride,
                       int rows, int width, float* out, void* stream);

/* ---- evaluation metrics on the device (metrics/train_loss_and_metrics_tracker.py:102-213,
 * utils/eval_utils.py:7-85) ----
 * Per body b: pred/target dev [B,N,3].  `which` selects (bit 0) sum_i ||p_i - t_i||, (bit 1) the same after the scale-and-
 * translation correction of eval_utils.py:63-85, (bit 2) after the similarity (Procrustes) alignment of eval_utils.py:7-52.
 * sums: dev double[3] ACCUMULATED into (fp64 atomics; slots of unselected metrics untouched) or NULL; pred_sc / pred_pa:
 * optional dev [B,N,3] outputs of the corrected / aligned points (what the reference's functions return). */
int straps_points_metrics(const float* pred, const float* target, int batch, int npoints, int which, double* sums,
                          float* pred_sc, float* pred_pa, void* stream);
/* *sum += sum over rows of ||(pred + pred_add) * pred_mul - target||_2 (squared == 0; rows of `dim` floats) or of the squared
 * differences (squared != 0): joints2D L2 error with utils/joints2d_utils.py:5-10 un-normalisation, pose / shape MSE sums. */
int straps_rows_metric(const float* pred, const float* target, int64_t rows, int dim, float pred_add, float pred_mul,
                       int squared, double* sum, void* stream);
/* dst[i] += scale * src[i], i < n (dev double accumulator; loss sums of the tracker without .item()). */
int straps_accumulate(const float* src, int n, double scale, double* dst, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Regressor -- replaces models/regressor.py:43-47 = ResNet.forward (models/resnet.py:201-216) +
 * IEFModule.forward (models/ief_module.py:48-64).
 * ------------------------------------------------------------------------------------------------ */
typedef struct straps_regressor straps_regressor_t;

/* Conv precision modes for the encoder. */
#define STRAPS_CONV_FP32_SIMT 0   /* fp32 FMA on CUDA cores (exact-order reference mode)           */
#define STRAPS_CONV_F16X3_TC 1   /* tcgen05 tensor cores, 2-term fp16 split, 3 MMA passes, fp32 acc */

/* Allocates packed-weight storage + activation workspace for batches up to max_batch on the current
 * device.  c_in = number of input channels (17 / 18 ...). */
int straps_regressor_create(straps_regressor_t** out, int c_in, int max_batch);
void straps_regressor_destroy(straps_regressor_t* r);
size_t straps_regressor_workspace_bytes(const straps_regressor_t* r);
/* Name of the i-th convolution (0..19) in the order straps_regressor_load expects: "conv1",
 * "layer1.0.conv1", "layer1.0.conv2", ..., "layer2.0.downsample.0", ... (state_dict order). */
const char* straps_regressor_conv_name(const straps_regressor_t* r, int i);

/* (Re)pack weights from PyTorch-owned device tensors in the reference state_dict layout:
 *   conv_w[20]  : OIHW fp32 conv weights in forward order (conv1, then per BasicBlock conv1, conv2,
 *                 [downsample.0])  -- models/resnet.py:145,177-199.
 *   bn[20*4]    : for each conv's BatchNorm: weight, bias, running_mean, running_var (fp32 [Cout]).
 *   fc_w[3], fc_b[3] : IEF fc1 [512,669], fc2 [512,512], fc3 [157,512] and biases.
 *   init_params : dev [157] initial estimate (models/ief_module.py:31).
 * The IEF weights are packed here; the convolution / BatchNorm tensors are REFERENCED, and their inference copies
 * (eval-mode BN folded into per-channel scale/shift, fp32 or fp16-split weight layouts) are built by the first forward
 * that needs them -- so every tensor must stay alive and in place until the next straps_regressor_load (the training
 * path reads and, for running_mean / running_var, writes them directly).  Asynchronous on `stream`. */
int straps_regressor_load(straps_regressor_t* r, const float* const* conv_w, const float* const* bn,
                          const float* const* fc_w, const float* const* fc_b, const float* init_params,
                          void* stream);

/* Encoder only: x dev [B,C,256,256] NCHW fp32 -> feat dev [B,512]. */
int straps_encoder_forward(straps_regressor_t* r, const float* x, int batch, int conv_mode,
                           float* feat, void* stream);
/* IEF only: feat dev [B,512] -> params dev [B,157] after `iters` iterations. */
int straps_ief_forward(straps_regressor_t* r, const float* feat, int batch, int iters,
                       float* params, void* stream);
/* Both: x -> params [B,157] (cam = [:, :3], pose6d = [:, 3:147], shape = [:, 147:]). */
int straps_regressor_forward(straps_regressor_t* r, const float* x, int batch, int conv_mode,
                             int iters, float* feat_or_null, float* params, void* stream);
/* The same with the proxy representation generated inside the stem's input pack (SURVEY.md 8f row N1 fused with the hot path):
 * the reference builds x = cat([binary(seg_labels), heatmaps(joints2d)]) in fp32 (utils/label_conversions.py:48-55,90-127,
 * train/train_synthetic_otf_rendering.py:178-182) and feeds it to the regressor; here x is never materialised.
 * seg_labels dev [B,256,256] fp32 part labels (non-zero = body), joints2d dev [B,num_joints,2] pixel coordinates,
 * table dev [(2*half_size)^2] = the reference's truncated Gaussian window (computed by the caller with the reference's own torch
 * ops so that the values are bit-identical).  num_joints + 1 must equal the handle's input channels.  Tensor-core mode only.
 * Bit-identical to straps_regressor_forward on the x the reference would have built. */
int straps_regressor_forward_from_labels(straps_regressor_t* r, const float* seg_labels, const float* joints2d, int num_joints,
                                         const float* table, int half_size, int batch, int iters, float* feat_or_null,
                                         float* params, void* stream);

/* ---- training path of the regressor (BASELINE config 3; reference train/...:186,230-233) ----
 * Train-mode forward of the encoder: BatchNorm uses batch statistics (biased variance), and -- when
 * update_running_stats != 0 -- updates the PyTorch-owned running_mean / running_var tensors given to
 * straps_regressor_load in place (momentum 0.1, unbiased variance).  Keeps every activation and pre-BN
 * convolution output in the handle's training workspace for straps_encoder_backward.
 * conv_mode: STRAPS_CONV_FP32_SIMT = fp32 CUDA-core convolutions (forward, data and weight gradients);
 *            STRAPS_CONV_F16X3_TC = forward convolutions, data gradients and weight gradients on the tensor cores
 *            (3-pass fp16 split, fp32-equivalent). */
int straps_encoder_train_forward(straps_regressor_t* r, const float* x, int batch, int update_running_stats,
                                 int conv_mode, float* feat, void* stream);
/* Backward of the last straps_encoder_train_forward: dfeat dev [B,512] ->
 *   d_conv_w[20] : OIHW fp32 weight gradients (same order as straps_regressor_load's conv_w), overwritten;
 *   d_bn[40]     : (d_weight, d_bias) of the 20 BatchNorms, overwritten.
 * conv_mode: -1 = the mode of the forward; STRAPS_CONV_FP32_SIMT after a tensor-core forward runs the fp32 CUDA-core
 * gradients on the SAME saved activations (the parity tests use this to compare the two arithmetic paths without
 * ReLU / max-pool flip noise); the tensor-core backward needs a tensor-core forward.  May be called repeatedly. */
int straps_encoder_backward(straps_regressor_t* r, const float* dfeat, int batch, int conv_mode,
                            float* const* d_conv_w, float* const* d_bn, void* stream);
/* The same pass in pieces: BasicBlocks block_hi .. block_lo (7 = layer4.1 ... 0 = layer1.0, processed last-first), `first` != 0 on the
 * call that starts the pass, `stem` != 0 on the call that ends it (max-pool / bn1 / conv1).  Only the gradients of the blocks covered
 * are written.  (straps_b200/parallel.py starts the all-reduce of layer4 + IEF between blocks 7..6 and 5..0.) */
int straps_encoder_backward_range(straps_regressor_t* r, const float* dfeat, int batch, int conv_mode,
                                  float* const* d_conv_w, float* const* d_bn, int block_hi, int block_lo, int first, int stem,
                                  void* stream);
/* IEF forward that also saves, per iteration, { p_k [B,157] | h1_k [B,512] | h2_k [B,512] } into
 * saved dev [iters * B * 1181] for straps_ief_backward. */
int straps_ief_forward_train(straps_regressor_t* r, const float* feat, int batch, int iters, float* params,
                             float* saved, void* stream);
/* IEF backward (shared-weight gradients accumulated over the iterations): d_params dev [B,157] ->
 * d_feat dev [B,512], d_fc_w[3] (nn.Linear layout), d_fc_b[3]; scratch dev [B * 1850] floats. */
int straps_ief_backward(straps_regressor_t* r, const float* feat, const float* saved, const float* d_params,
                        int batch, int iters, float* d_feat, float* const* d_fc_w, float* const* d_fc_b,
                        float* scratch, void* stream);
/* torch.optim.Adam step (run_train.py:200-201 defaults: no weight decay) on a flat fp32 bucket; the gradient is
 * multiplied by grad_scale first (1/world_size after the summing all-reduce).  step counts from 1. */
int straps_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int step,
                     float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
/* The same step with the step count on the device: *step_counter (dev int64) is incremented first and the bias corrections are
 * derived from it inside the kernel, so the call can be captured in a CUDA graph and replayed (a host-side `step` would be frozen
 * into the graph).  Same formula as straps_adam_step(step = the incremented value); beta^t is evaluated on the device. */
int straps_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t* step_counter,
                         float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* Fused multi-task loss, forward + backward seeds -- losses/multi_task_loss.py:73-119 (+ utils/joints2d_utils.py:23-33 as the
 * joints2D row mask).  preds / targets / grads: 5 dev pointers each in the order verts, joints2D [rows,2], joints3D, shape_params,
 * pose rotmats (preds[t] == NULL switches task t off; grads[t] may be NULL).  counts[5] = element counts.  vis: optional dev uint8
 * [rows] visibility mask (NULL + use_vis=1: computed from the pixel label exactly as the reference does; use_vis=0: all rows).
 * log_vars dev [5] in the same task order.  scratch dev 128 bytes.  out dev [6] = (total, 5 weighted task losses);
 * d_log_vars dev [5]; g_total dev [1] or NULL (= 1) scales the prediction gradients. */
int straps_multitask_loss(const float* const* preds, const float* const* targets, float* const* grads, const int64_t* counts,
                          const unsigned char* vis, int use_vis, float img_wh, const float* log_vars, int sum_reduction,
                          const float* g_total, void* scratch, float* out, float* d_log_vars, void* stream);

/* Debug/parity hook: copy a named intermediate activation (as NCHW fp32) of the last forward into
 * out (dev).  Names: "stem", "pool", "layer1.0" ... "layer4.1".  Returns element count via *n. */
int straps_encoder_read_activation(straps_regressor_t* r, const char* name, int batch, float* out,
                                   int64_t* n, void* stream);

#pragma GCC visibility pop

#ifdef __cplusplus
}
#endif
#endif /* STRAPS_B200_H_ */

"""Torch-tensor front ends of the C ABI (device memory and streams only -- the maths is in csrc/)."""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import StrapsError, check

CONV_FP32_SIMT = 0
CONV_F16X3_TC = 1
_MODE_NAMES = {'fp32_simt': CONV_FP32_SIMT, 'f16x3_tc': CONV_F16X3_TC}
DEFAULT_CONV_MODE = os.environ.get('STRAPS_CONV_MODE', 'f16x3_tc')


def conv_mode_id(mode):
    if isinstance(mode, int):
        return mode
    if mode not in _MODE_NAMES:
        raise ValueError('unknown conv mode %r (expected one of %s)' % (mode, sorted(_MODE_NAMES)))
    return _MODE_NAMES[mode]


def _need_cuda(t, name):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise StrapsError('%s must be a CUDA tensor: the B200 path has no CPU fallback' % name)
    if t.dtype != torch.float32:
        raise StrapsError('%s must be float32 (got %s)' % (name, t.dtype))


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def rot6d_to_rotmat(x):
    """[..., 6k] -> [n, 3, 3]  (reference utils/rigid_transform_utils.py:27-41)."""
    _need_cuda(x, 'x')
    x = x.contiguous().view(-1, 6)
    out = torch.empty((x.shape[0], 3, 3), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().straps_rot6d_to_rotmat(_p(x), x.shape[0], _p(out), _stream(x.device)), 'straps_rot6d_to_rotmat')
    return out


def orthographic_project(points3d, cam):
    """[B,N,3], [B,3] -> [B,N,2]  (reference utils/cam_utils.py:5-26)."""
    _need_cuda(points3d, 'points3D')
    _need_cuda(cam, 'cam_params')
    points3d = points3d.contiguous()
    if cam.stride(1) != 1:
        cam = cam.contiguous()
    B, N = points3d.shape[0], points3d.shape[1]
    out = torch.empty((B, N, 2), dtype=torch.float32, device=points3d.device)
    with torch.cuda.device(points3d.device):
        check(_lib.lib().straps_orthographic_project(_p(points3d), _p(cam), cam.stride(0), B, N, _p(out),
                                                     _stream(points3d.device)), 'straps_orthographic_project')
    return out


def rot6d_backward(x6, dR):
    """Gradient of rot6d_to_rotmat: x6 [n,6], dR [n,3,3] -> [n,6]."""
    _need_cuda(x6, 'x')
    _need_cuda(dR, 'grad')
    x6, dR = x6.contiguous(), dR.contiguous()
    dx = torch.empty_like(x6)
    with torch.cuda.device(x6.device):
        check(_lib.lib().straps_rot6d_backward(_p(x6), _p(dR), x6.shape[0], _p(dx), _stream(x6.device)), 'straps_rot6d_backward')
    return dx


def orthographic_project_backward(points3d, cam, g_out):
    """-> (d_points [B,N,3], d_cam [B,3])."""
    points3d, g_out = points3d.contiguous(), g_out.contiguous()
    if cam.stride(1) != 1:
        cam = cam.contiguous()
    B, N = points3d.shape[0], points3d.shape[1]
    dp = torch.empty_like(points3d)
    dc = torch.empty((B, 3), dtype=torch.float32, device=points3d.device)
    with torch.cuda.device(points3d.device):
        check(_lib.lib().straps_orthographic_project_backward(_p(points3d), _p(cam), cam.stride(0), _p(g_out), B, N, _p(dp), _p(dc),
                                                              _stream(points3d.device)), 'straps_orthographic_project_backward')
    return dp, dc


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
    """torch.optim.Adam update on flat fp32 CUDA buffers (in place)."""
    for t in (params, grads, exp_avg, exp_avg_sq):
        _need_cuda(t, 'adam buffer')
        if not t.is_contiguous():
            raise StrapsError('adam buffers must be contiguous')
    with torch.cuda.device(params.device):
        check(_lib.lib().straps_adam_step(_p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(), int(step), float(lr),
                                          float(betas[0]), float(betas[1]), float(eps), float(grad_scale), _stream(params.device)),
              'straps_adam_step')


def adam_step_dev(params, grads, exp_avg, exp_avg_sq, step_counter, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
    """adam_step with the step count in `step_counter` (CUDA int64, 1 element; incremented by the call): CUDA-graph capturable."""
    for t in (params, grads, exp_avg, exp_avg_sq):
        _need_cuda(t, 'adam buffer')
        if not t.is_contiguous():
            raise StrapsError('adam buffers must be contiguous')
    if not step_counter.is_cuda or step_counter.dtype != torch.int64 or step_counter.numel() != 1:
        raise StrapsError('adam_step_dev: step_counter must be a CUDA int64 tensor with one element')
    with torch.cuda.device(params.device):
        check(_lib.lib().straps_adam_step_dev(_p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(), _p(step_counter),
                                              float(lr), float(betas[0]), float(betas[1]), float(eps), float(grad_scale),
                                              _stream(params.device)), 'straps_adam_step_dev')


# ---- target side of the synthetic loop ----
def batch_rodrigues(rot_vecs):
    """[n,3] axis-angle -> [n,3,3]  (smplx.lbs.batch_rodrigues; reference call sites augmentation/smpl_augmentation.py:55-58)."""
    _need_cuda(rot_vecs, 'rot_vecs')
    if rot_vecs.dim() != 2 or rot_vecs.shape[1] != 3:
        raise StrapsError('batch_rodrigues expects [n,3], got %s' % (tuple(rot_vecs.shape),))
    r = rot_vecs.contiguous()
    out = torch.empty((r.shape[0], 3, 3), dtype=torch.float32, device=r.device)
    with torch.cuda.device(r.device):
        check(_lib.lib().straps_batch_rodrigues(_p(r), r.shape[0], _p(out), _stream(r.device)), 'straps_batch_rodrigues')
    return out


def batch_rodrigues_backward(rot_vecs, dR):
    """Gradient of batch_rodrigues: rot_vecs [n,3], dR [n,3,3] -> [n,3]."""
    _need_cuda(rot_vecs, 'rot_vecs')
    _need_cuda(dR, 'grad')
    r, g = rot_vecs.contiguous(), dR.contiguous()
    out = torch.empty_like(r)
    with torch.cuda.device(r.device):
        check(_lib.lib().straps_batch_rodrigues_backward(_p(r), _p(g), r.shape[0], _p(out), _stream(r.device)),
              'straps_batch_rodrigues_backward')
    return out


def perspective_project(points, rotation, translation, cam_K):
    """[B,N,3], [B,3,3], [B,3], [B,3,3] -> [B,N,2]  (reference utils/cam_utils.py:40-71)."""
    for n, t in (('points', points), ('rotation', rotation), ('translation', translation), ('cam_K', cam_K)):
        _need_cuda(t, n)
    B, N = points.shape[0], points.shape[1]
    if rotation.shape != (B, 3, 3) or translation.shape != (B, 3) or cam_K.shape != (B, 3, 3):
        raise StrapsError('perspective_project: expected rotation [B,3,3], translation [B,3], cam_K [B,3,3] for B=%d' % B)
    points, rotation, translation, cam_K = points.contiguous(), rotation.contiguous(), translation.contiguous(), cam_K.contiguous()
    out = torch.empty((B, N, 2), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().straps_perspective_project(_p(points), _p(rotation), _p(translation), _p(cam_K), B, N, _p(out),
                                                    _stream(points.device)), 'straps_perspective_project')
    return out


def scale_shift(noise, mul, add, base):
    """out[r,c] = (noise[r,c]*mul[c] + add[c]) + base[r or 0, c], each op rounded separately like the reference's torch ops.
    noise [R,W]; mul / add: [W] tensors or python scalars (add may be None); base [R,W] or [W]."""
    _need_cuda(noise, 'noise')
    noise = noise.contiguous()
    R, W = noise.shape
    dev = noise.device

    def vec(v):
        if v is None:
            return None
        if torch.is_tensor(v):
            v = v.to(device=dev, dtype=torch.float32).reshape(-1)
            return (v.expand(W) if v.numel() == 1 else v).contiguous()
        return torch.full((W,), float(v), dtype=torch.float32, device=dev)
    mul_t, add_t = vec(mul), vec(add)
    _need_cuda(base, 'base')
    base2 = base.reshape(-1, W)
    if base2.shape[0] not in (1, R):
        raise StrapsError('scale_shift: base has %d rows, expected 1 or %d' % (base2.shape[0], R))
    if base2.stride(1) != 1:
        base2 = base2.contiguous()
    stride = 0 if base2.shape[0] == 1 else base2.stride(0)
    out = torch.empty((R, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().straps_scale_shift(_p(noise), _p(mul_t), _p(add_t) if add_t is not None else None, _p(base2), stride, R, W,
                                            _p(out), _stream(dev)), 'straps_scale_shift')
    return out


# ---- evaluation metrics on the device ----
METRIC_PLAIN, METRIC_SC, METRIC_PA = 1, 2, 4


def points_metrics(pred, target, which, sums=None, want_sc=False, want_pa=False):
    """pred/target [B,N,3].  Accumulates the selected error sums into `sums` (CUDA float64 [3]: plain, scale-corrected,
    Procrustes-aligned) and/or returns the corrected / aligned copies of pred (reference utils/eval_utils.py:7-85)."""
    _need_cuda(pred, 'pred')
    _need_cuda(target, 'target')
    if pred.shape != target.shape or pred.dim() != 3 or pred.shape[2] != 3:
        raise StrapsError('points_metrics expects two [B,N,3] tensors, got %s and %s' % (tuple(pred.shape), tuple(target.shape)))
    pred, target = pred.contiguous(), target.contiguous()
    if sums is not None and (not sums.is_cuda or sums.dtype != torch.float64 or sums.numel() < 3 or not sums.is_contiguous()):
        raise StrapsError('points_metrics: sums must be a contiguous CUDA float64 tensor with >= 3 elements')
    out_sc = torch.empty_like(pred) if want_sc else None
    out_pa = torch.empty_like(pred) if want_pa else None
    with torch.cuda.device(pred.device):
        check(_lib.lib().straps_points_metrics(_p(pred), _p(target), pred.shape[0], pred.shape[1], int(which),
                                               _p(sums) if sums is not None else None, _p(out_sc) if want_sc else None,
                                               _p(out_pa) if want_pa else None, _stream(pred.device)), 'straps_points_metrics')
    return out_sc, out_pa


def rows_metric(pred, target, sum_slot, dim, pred_add=0.0, pred_mul=1.0, squared=False):
    """sum_slot (CUDA float64, 1 element view) += sum over rows of ||(pred+add)*mul - target|| (or squared differences)."""
    _need_cuda(pred, 'pred')
    _need_cuda(target, 'target')
    if pred.numel() != target.numel() or pred.numel() % dim:
        raise StrapsError('rows_metric: shapes %s / %s do not form rows of %d' % (tuple(pred.shape), tuple(target.shape), dim))
    pred, target = pred.contiguous(), target.contiguous()
    with torch.cuda.device(pred.device):
        check(_lib.lib().straps_rows_metric(_p(pred), _p(target), pred.numel() // dim, dim, float(pred_add), float(pred_mul),
                                            1 if squared else 0, _p(sum_slot), _stream(pred.device)), 'straps_rows_metric')


_JOINT_INDEX = {}


def select_joints(joints, *index_lists):
    """`joints[:, a, :][:, b, :]...` with the index lists of config.py (ALL_JOINTS_TO_COCO_MAP, ALL_JOINTS_TO_H36M_MAP, H36M_TO_J14;
    the reference indexes with the Python lists themselves, train/train_synthetic_otf_rendering.py:207-215) as ONE index_select with a
    cached device index: same values, but no host->device copy of the list per call (which breaks CUDA-graph capture) and an
    index_add backward instead of the sort-based index_put of advanced indexing."""
    key = (tuple(tuple(int(i) for i in l) for l in index_lists), joints.device)
    idx = _JOINT_INDEX.get(key)
    if idx is None:
        composed = list(key[0][0])
        for l in key[0][1:]:
            composed = [composed[i] for i in l]
        idx = torch.tensor(composed, dtype=torch.long, device=joints.device)
        _JOINT_INDEX[key] = idx
    return joints.index_select(1, idx)


def accumulate(src, scale, dst):
    """dst[:n] += scale * src  (src CUDA float32 [n], dst CUDA float64)."""
    _need_cuda(src, 'src')
    src = src.contiguous()
    with torch.cuda.device(src.device):
        check(_lib.lib().straps_accumulate(_p(src), src.numel(), float(scale), _p(dst), _stream(src.device)), 'straps_accumulate')


class SmplHandle(object):
    """Owns the device-resident packed SMPL constants of one CUDA device."""

    def __init__(self, device, v_template, shapedirs, posedirs, J_regressor, lbs_weights, parents,
                 extra_regressors, extra_pick_idx):
        f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))
        i64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int64))
        arrs = [f32(v_template), f32(shapedirs), f32(posedirs), f32(J_regressor), f32(lbs_weights), i64(parents),
                f32(extra_regressors), i64(extra_pick_idx)]
        shapes = [(6890, 3), (6890, 3, 10), (207, 20670), (24, 6890), (6890, 24), (24,), (45, 6890), (21,)]
        for a, s in zip(arrs, shapes):
            if tuple(a.shape) != s:
                raise StrapsError('SMPL constant has shape %s, expected %s' % (a.shape, s))
        self.device = torch.device(device)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_smpl_create(ctypes.byref(self._h), *[ctypes.c_void_p(a.ctypes.data) for a in arrs]),
                  'straps_smpl_create')
        self.sparse4 = bool(_lib.lib().straps_smpl_is_sparse4(self._h))

    def __del__(self):
        try:
            if self._h:
                _lib.lib().straps_smpl_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def forward(self, global_orient, body_pose, betas, transl, pose2rot):
        for n, t in (('global_orient', global_orient), ('body_pose', body_pose), ('betas', betas)):
            _need_cuda(t, n)
        B = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
        if pose2rot:
            go = global_orient.reshape(global_orient.shape[0], 3)
            bp = body_pose.reshape(body_pose.shape[0], 69)
        else:
            go = global_orient.reshape(global_orient.shape[0], 9)
            bp = body_pose.reshape(body_pose.shape[0], 23 * 9)

        def rows(t):
            # row-strided views are passed as is; anything else is made contiguous
            if t.stride(-1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
                t = t.contiguous()
            if t.shape[0] != B:
                t = t.expand(B, -1).contiguous()
            return t
        go, bp, be = rows(go), rows(bp), rows(betas.reshape(betas.shape[0], 10))
        tr = None
        if transl is not None:
            _need_cuda(transl, 'transl')
            tr = transl.reshape(-1, 3)
            if tr.shape[0] not in (1, B):
                # smplx broadcasts `vertices + transl.unsqueeze(1)`: a model built for another batch size fails there too
                raise StrapsError('SMPL: transl has batch %d but the call has batch %d -- construct SMPL(batch_size=%d) '
                                  '(reference run_train.py:109-110)' % (tr.shape[0], B, B))
            tr = tr.expand(B, -1).contiguous()
        verts = torch.empty((B, 6890, 3), dtype=torch.float32, device=betas.device)
        joints = torch.empty((B, 90, 3), dtype=torch.float32, device=betas.device)
        with torch.cuda.device(betas.device):
            check(_lib.lib().straps_smpl_forward(self._h, _p(go), go.stride(0), _p(bp), bp.stride(0), _p(be), be.stride(0),
                                                 _p(tr) if tr is not None else None, B, 1 if pose2rot else 0,
                                                 _p(verts), _p(joints), _stream(betas.device)), 'straps_smpl_forward')
        return verts, joints


    def forward_train(self, rotmats, betas):
        """rotmats [B,24,3,3], betas [B,10] -> (verts, joints, saved) with what straps_smpl_backward needs."""
        _need_cuda(rotmats, 'rotmats')
        _need_cuda(betas, 'betas')
        rotmats, betas = rotmats.contiguous(), betas.contiguous()
        B, dev = rotmats.shape[0], rotmats.device
        verts = torch.empty((B, 6890, 3), dtype=torch.float32, device=dev)
        joints = torch.empty((B, 90, 3), dtype=torch.float32, device=dev)
        vposed = torch.empty((B, 6890, 3), dtype=torch.float32, device=dev)
        A = torch.empty((B, 24, 12), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().straps_smpl_forward_train(self._h, _p(rotmats), _p(betas), B, _p(verts), _p(joints), _p(vposed), _p(A),
                                                       _stream(dev)), 'straps_smpl_forward_train')
        return verts, joints, (rotmats, betas, vposed, A)

    def backward(self, saved, g_verts, g_joints):
        rotmats, betas, vposed, A = saved
        B, dev = rotmats.shape[0], rotmats.device
        gv = g_verts.contiguous().clone()          # the kernel accumulates the joint contributions into it
        gj = g_joints.contiguous()
        scratch = torch.empty((B * (24 * 12 + 208 + 10),), dtype=torch.float32, device=dev)
        d_rot = torch.empty_like(rotmats)
        d_betas = torch.empty_like(betas)
        with torch.cuda.device(dev):
            check(_lib.lib().straps_smpl_backward(self._h, _p(rotmats), _p(betas), _p(vposed), _p(A), _p(gj), _p(gv), _p(scratch), B,
                                                  _p(d_rot), _p(d_betas), _stream(dev)), 'straps_smpl_backward')
        return d_rot, d_betas


class RegressorHandle(object):
    """Packed encoder + IEF weights and activation workspace of one CUDA device."""

    def __init__(self, device, c_in, max_batch):
        self.device = torch.device(device)
        self.c_in, self.max_batch = c_in, max_batch
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_regressor_create(ctypes.byref(self._h), c_in, max_batch), 'straps_regressor_create')
        self.conv_names = [_lib.lib().straps_regressor_conv_name(self._h, i).decode() for i in range(20)]
        self._keep = None
        self.generation = 0      # bumped by every call that overwrites the encoder workspace (see engine._RegressorTrain)

    def __del__(self):
        try:
            if self._h:
                _lib.lib().straps_regressor_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def workspace_bytes(self):
        return int(_lib.lib().straps_regressor_workspace_bytes(self._h))

    def load(self, conv_w, bn, fc_w, fc_b, init_params):
        """conv_w: 20 tensors (order = self.conv_names); bn: 20 x (weight, bias, running_mean, running_var)."""
        flat_bn = [t for quad in bn for t in quad]
        tensors = list(conv_w) + flat_bn + list(fc_w) + list(fc_b) + [init_params]
        keep = []
        for t in tensors:
            _need_cuda(t, 'weight')
            keep.append(t.detach().contiguous())
        arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        n = 0
        cw = keep[n:n + 20]; n += 20
        bb = keep[n:n + 80]; n += 80
        fw = keep[n:n + 3]; n += 3
        fb = keep[n:n + 3]; n += 3
        ip = keep[n]
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_regressor_load(self._h, arr(cw), arr(bb), arr(fw), arr(fb), _p(ip), _stream(self.device)),
                  'straps_regressor_load')
        self._keep = keep   # the pack kernels are asynchronous: keep the sources alive

    def encoder_forward(self, x, mode=DEFAULT_CONV_MODE):
        _need_cuda(x, 'input')
        x = x.contiguous()
        B = x.shape[0]
        if x.shape[1:] != (self.c_in, 256, 256):
            raise StrapsError('encoder input must be [B,%d,256,256], got %s' % (self.c_in, tuple(x.shape)))
        feat = torch.empty((B, 512), dtype=torch.float32, device=x.device)
        self.generation += 1
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_encoder_forward(self._h, _p(x), B, conv_mode_id(mode), _p(feat), _stream(self.device)),
                  'straps_encoder_forward')
        return feat

    def ief_forward(self, feat, iters=3):
        _need_cuda(feat, 'img_features')
        feat = feat.contiguous()
        params = torch.empty((feat.shape[0], 157), dtype=torch.float32, device=feat.device)
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_ief_forward(self._h, _p(feat), feat.shape[0], iters, _p(params), _stream(self.device)),
                  'straps_ief_forward')
        return params

    def forward(self, x, mode=DEFAULT_CONV_MODE, iters=3, out=None):
        _need_cuda(x, 'input')
        x = x.contiguous()
        B = x.shape[0]
        if x.shape[1:] != (self.c_in, 256, 256):
            raise StrapsError('regressor input must be [B,%d,256,256], got %s' % (self.c_in, tuple(x.shape)))
        params = out if out is not None else torch.empty((B, 157), dtype=torch.float32, device=x.device)
        self.generation += 1
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_regressor_forward(self._h, _p(x), B, conv_mode_id(mode), iters, None, _p(params),
                                                      _stream(self.device)), 'straps_regressor_forward')
        return params

    def forward_from_labels(self, seg_labels, joints2d, table, half_size, iters=3, out=None):
        """seg_labels [B,256,256], joints2d [B,J,2] (J + 1 = input channels) -> params [B,157]; the proxy representation is generated
        inside the stem's input pack (straps_regressor_forward_from_labels)."""
        _need_cuda(seg_labels, 'seg_labels')
        _need_cuda(joints2d, 'joints2D')
        seg_labels, joints2d = seg_labels.contiguous(), joints2d.contiguous()
        B, J = seg_labels.shape[0], joints2d.shape[1]
        if seg_labels.shape[1:] != (256, 256) or joints2d.shape != (B, J, 2) or J + 1 != self.c_in:
            raise StrapsError('forward_from_labels expects seg_labels [B,256,256] and joints2D [B,%d,2], got %s and %s'
                              % (self.c_in - 1, tuple(seg_labels.shape), tuple(joints2d.shape)))
        params = out if out is not None else torch.empty((B, 157), dtype=torch.float32, device=seg_labels.device)
        self.generation += 1
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_regressor_forward_from_labels(self._h, _p(seg_labels), _p(joints2d), J, _p(table), int(half_size), B, iters,
                                                                  None, _p(params), _stream(self.device)),
                  'straps_regressor_forward_from_labels')
        return params

    # ---- training path ----
    def encoder_train_forward(self, x, update_running_stats=True, mode=DEFAULT_CONV_MODE):
        _need_cuda(x, 'input')
        x = x.contiguous()
        B = x.shape[0]
        if x.shape[1:] != (self.c_in, 256, 256):
            raise StrapsError('encoder input must be [B,%d,256,256], got %s' % (self.c_in, tuple(x.shape)))
        feat = torch.empty((B, 512), dtype=torch.float32, device=x.device)
        self.generation += 1
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_encoder_train_forward(self._h, _p(x), B, 1 if update_running_stats else 0, conv_mode_id(mode), _p(feat),
                                                          _stream(self.device)), 'straps_encoder_train_forward')
        return feat

    def encoder_backward(self, dfeat, conv_shapes, bn_channels, mode=None, out_w=None, out_bn=None, blocks=None, dws=None, dbn=None):
        """-> (list of 20 OIHW weight gradients, list of 20 (d_weight, d_bias)).  mode None = the mode of the forward.
        out_w / out_bn: optional pre-allocated destinations (None entries are allocated here); every gradient is OVERWRITTEN.
        blocks = (block_hi, block_lo, first, stem): one piece of the pass (straps_encoder_backward_range)."""
        dfeat = dfeat.contiguous()
        B = dfeat.shape[0]
        new = lambda s: torch.empty(s, dtype=torch.float32, device=self.device)
        if dws is None:             # (a second call of a pass in pieces hands the first call's tensors back in)
            dws = [out_w[i] if out_w is not None and out_w[i] is not None else new(s) for i, s in enumerate(conv_shapes)]
            dbn = [((out_bn[i][0] if out_bn is not None and out_bn[i][0] is not None else new(c)),
                    (out_bn[i][1] if out_bn is not None and out_bn[i][1] is not None else new(c))) for i, c in enumerate(bn_channels)]
        flat_bn = [t for pair in dbn for t in pair]
        arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        hi, lo, first, stem = (7, 0, 1, 1) if blocks is None else blocks
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_encoder_backward_range(self._h, _p(dfeat), B, -1 if mode is None else conv_mode_id(mode), arr(dws),
                                                           arr(flat_bn), hi, lo, first, stem, _stream(self.device)),
                  'straps_encoder_backward_range')
        return dws, dbn

    def ief_forward_train(self, feat, iters=3):
        feat = feat.contiguous()
        B = feat.shape[0]
        params = torch.empty((B, 157), dtype=torch.float32, device=feat.device)
        saved = torch.empty((max(iters, 1) * B * (157 + 1024),), dtype=torch.float32, device=feat.device)
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_ief_forward_train(self._h, _p(feat), B, iters, _p(params), _p(saved), _stream(self.device)),
                  'straps_ief_forward_train')
        return params, saved

    def ief_backward(self, feat, saved, d_params, iters=3, out_w=None, out_b=None):
        feat, d_params = feat.contiguous(), d_params.contiguous()
        B, dev = feat.shape[0], feat.device
        d_feat = torch.empty((B, 512), dtype=torch.float32, device=dev)
        dw = [out_w[i] if out_w is not None and out_w[i] is not None else torch.empty(s, dtype=torch.float32, device=dev)
              for i, s in enumerate(((512, 669), (512, 512), (157, 512)))]
        db = [out_b[i] if out_b is not None and out_b[i] is not None else torch.empty(s, dtype=torch.float32, device=dev)
              for i, s in enumerate((512, 512, 157))]
        scratch = torch.empty((B * 1850,), dtype=torch.float32, device=dev)
        arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_ief_backward(self._h, _p(feat), _p(saved), _p(d_params), B, iters, _p(d_feat), arr(dw), arr(db),
                                                 _p(scratch), _stream(self.device)), 'straps_ief_backward')
        return d_feat, dw, db

    def read_activation(self, name, batch):
        shapes = {'stem': (64, 128, 128), 'pool': (64, 64, 64), 'grad:conv1': (64, 128, 128)}
        for L, (c, hw) in enumerate(((64, 64), (128, 32), (256, 16), (512, 8))):
            for blk in range(2):
                for suffix in ('', '.a', '.ds'):
                    shapes['layer%d.%d%s' % (L + 1, blk, suffix)] = (c, hw, hw)
        c, h, w = shapes[name]
        out = torch.empty((batch, c, h, w), dtype=torch.float32, device=self.device)
        n = ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            check(_lib.lib().straps_encoder_read_activation(self._h, name.encode(), batch, _p(out), ctypes.byref(n),
                                                            _stream(self.device)), 'straps_encoder_read_activation')
        assert n.value == out.numel()
        return out

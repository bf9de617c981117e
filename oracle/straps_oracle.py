"""ORACLE / TEST INFRASTRUCTURE -- CPU restatement of the STRAPS hot path in plain PyTorch fp32.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (straps-3dhumanshapepose_b200/) never does; it fails loudly when
its CUDA library is missing.

Every function cites the reference file:line it follows.  Pinning status:
  * encoder / IEF / rot6d / projection / loss: the reference's own modules import fine, so this
    restatement is PINNED against them in this container (tests/test_oracle_vs_reference.py, skipped
    where /root/reference is absent) and against fixtures generated from them
    (tests/golden/*.npz, made by oracle/gen_golden.py).
  * SMPL forward: the arithmetic lives in the absent third-party `smplx` package, so that part is
    PARITY UNPINNED (see oracle/smplx_shim/smplx/__init__.py); fixtures for it come from the
    reference's models/smpl_official.py running unchanged on top of the shim.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, 'smplx_shim')

NUM_IEF_PARAMS = 3 + 24 * 6 + 10          # models/regressor.py:25-26
ALL_JOINTS_TO_COCO_MAP = [24, 26, 25, 28, 27, 16, 17, 18, 19, 20, 21, 1, 2, 4, 5, 7, 8]   # config.py:27
ALL_JOINTS_TO_H36M_MAP = list(range(73, 90))                                               # config.py:28
H36M_TO_J14 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10]                            # config.py:31-32
REGRESSOR_IMG_WH = 256                                                                     # config.py:14

# (name prefix, cin, cout, stride, has_downsample) for the 8 BasicBlocks -- models/resnet.py:150-156,177-199
_BLOCKS = [('layer1.0', 64, 64, 1, False), ('layer1.1', 64, 64, 1, False),
           ('layer2.0', 64, 128, 2, True), ('layer2.1', 128, 128, 1, False),
           ('layer3.0', 128, 256, 2, True), ('layer3.1', 256, 256, 1, False),
           ('layer4.0', 256, 512, 2, True), ('layer4.1', 512, 512, 1, False)]


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
def conv_bn_names():
    """(conv_key, bn_prefix, cin(None=input), cout, k, stride, pad) in forward order."""
    out = [('conv1', 'bn1', None, 64, 7, 2, 3)]
    for name, cin, cout, stride, ds in _BLOCKS:
        out.append((name + '.conv1', name + '.bn1', cin, cout, 3, stride, 1))
        out.append((name + '.conv2', name + '.bn2', cout, cout, 3, 1, 1))
        if ds:
            out.append((name + '.downsample.0', name + '.downsample.1', cin, cout, 1, stride, 0))
    return out


def make_regressor_state(c_in, seed=0, randomize_bn=True):
    """A numpy-seeded regressor state_dict with the reference's 132 keys.

    Conv weights ~ kaiming-normal fan_out (models/resnet.py:160-162); BN affine/running stats are
    randomised (when asked) so that eval-mode BN is not the identity; IEF weights use nn.Linear's
    uniform bound with zero biases (models/ief_module.py:16-22).  Keys are prefixed like
    SingleInputRegressor's: image_encoder.* and ief_module.* (fc1/2/3 duplicated as ief_layers.0/2/4).
    """
    rng = np.random.RandomState(seed)
    sd = {}

    def t(a, dtype=np.float32):
        return torch.from_numpy(np.ascontiguousarray(a.astype(dtype)))

    for conv, bn, cin, cout, k, stride, pad in conv_bn_names():
        cin = c_in if cin is None else cin
        std = np.sqrt(2.0 / (cout * k * k))
        sd['image_encoder.' + conv + '.weight'] = t(rng.normal(0, std, (cout, cin, k, k)))
        if randomize_bn:
            sd['image_encoder.' + bn + '.weight'] = t(rng.uniform(0.6, 1.4, cout))
            sd['image_encoder.' + bn + '.bias'] = t(rng.normal(0, 0.1, cout))
            sd['image_encoder.' + bn + '.running_mean'] = t(rng.normal(0, 0.1, cout))
            sd['image_encoder.' + bn + '.running_var'] = t(rng.uniform(0.6, 1.4, cout))
        else:
            sd['image_encoder.' + bn + '.weight'] = torch.ones(cout)
            sd['image_encoder.' + bn + '.bias'] = torch.zeros(cout)
            sd['image_encoder.' + bn + '.running_mean'] = torch.zeros(cout)
            sd['image_encoder.' + bn + '.running_var'] = torch.ones(cout)
        sd['image_encoder.' + bn + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)
    for name, fin, fout in (('fc1', 512 + NUM_IEF_PARAMS, 512), ('fc2', 512, 512), ('fc3', 512, NUM_IEF_PARAMS)):
        bound = 1.0 / np.sqrt(fin)
        w = t(rng.uniform(-bound, bound, (fout, fin)))
        # small non-zero biases exercise the bias path (reference zero-inits them at construction only)
        b = t(rng.normal(0, 0.01, fout))
        sd['ief_module.%s.weight' % name] = w
        sd['ief_module.%s.bias' % name] = b
    for i, name in ((0, 'fc1'), (2, 'fc2'), (4, 'fc3')):
        sd['ief_module.ief_layers.%d.weight' % i] = sd['ief_module.%s.weight' % name]
        sd['ief_module.ief_layers.%d.bias' % i] = sd['ief_module.%s.bias' % name]
    return sd


def load_initial_params(mean_params_npz):
    """models/ief_module.py:33-46."""
    m = np.load(mean_params_npz)
    p = np.zeros(NUM_IEF_PARAMS)
    p[3:] = np.concatenate((m['pose'], m['shape']))
    p[0] = 0.9
    return torch.from_numpy(p.astype(np.float32)).float()


# ----------------------------------------------------------------------------------------------
# encoder (models/resnet.py:201-216, BasicBlock 61-77)
# ----------------------------------------------------------------------------------------------
def _bn(x, sd, prefix, train, stats_out):
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    if train:
        rm, rv = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm, rv, w, b, True, 0.1, 1e-5)
        if stats_out is not None:
            stats_out[prefix + '.running_mean'] = rm
            stats_out[prefix + '.running_var'] = rv
        return y
    return F.batch_norm(x, rm, rv, w, b, False, 0.1, 1e-5)


def encoder_forward(x, sd, train=False, taps=None, stats_out=None, prefix='image_encoder.'):
    """x [B,C,256,256] -> feat [B,512].  `taps` (dict) receives named intermediates (NCHW)."""
    g = lambda k: sd[prefix + k]
    sdp = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    y = F.conv2d(x, g('conv1.weight'), None, 2, 3)                         # resnet.py:202
    y = F.relu(_bn(y, sdp, 'bn1', train, stats_out))                      # :203-204
    tap('stem', y)
    y = F.max_pool2d(y, 3, 2, 1)                                          # :206
    tap('pool', y)
    for name, cin, cout, stride, ds in _BLOCKS:                            # :208-211
        idt = y
        o = F.conv2d(y, g(name + '.conv1.weight'), None, stride, 1)       # :64
        o = F.relu(_bn(o, sdp, name + '.bn1', train, stats_out))          # :65-66
        tap(name + '.a', o)
        o = F.conv2d(o, g(name + '.conv2.weight'), None, 1, 1)            # :68
        o = _bn(o, sdp, name + '.bn2', train, stats_out)                  # :69
        if ds:                                                            # :71-72
            idt = F.conv2d(y, g(name + '.downsample.0.weight'), None, stride, 0)
            idt = _bn(idt, sdp, name + '.downsample.1', train, stats_out)
        y = F.relu(o + idt)                                               # :74-75
        tap(name, y)
    return torch.flatten(F.adaptive_avg_pool2d(y, (1, 1)), 1)             # :213-214


# ----------------------------------------------------------------------------------------------
# IEF (models/ief_module.py:48-64)
# ----------------------------------------------------------------------------------------------
def ief_forward(feat, sd, init_params, iters=3, trace=None, prefix='ief_module.'):
    W1, b1 = sd[prefix + 'fc1.weight'], sd[prefix + 'fc1.bias']
    W2, b2 = sd[prefix + 'fc2.weight'], sd[prefix + 'fc2.bias']
    W3, b3 = sd[prefix + 'fc3.weight'], sd[prefix + 'fc3.bias']
    p = init_params.repeat([feat.shape[0], 1]).to(feat.dtype)
    for _ in range(iters):
        state = torch.cat([feat, p], dim=1)
        h = F.relu(F.linear(state, W1, b1))
        h = F.relu(F.linear(h, W2, b2))
        p = p + F.linear(h, W3, b3)
        if trace is not None:
            trace.append(p.clone())
    return p


def split_params(p):
    """models/ief_module.py:60-62 -> cam[B,3], pose6d[B,144], shape[B,10]."""
    return p[:, :3], p[:, 3:147], p[:, 147:]


# ----------------------------------------------------------------------------------------------
# rot6d (utils/rigid_transform_utils.py:27-41), projection (utils/cam_utils.py:5-26)
# ----------------------------------------------------------------------------------------------
def rot6d_to_rotmat(x):
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - torch.einsum('bi,bi->b', b1, a2).unsqueeze(-1) * b1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def orthographic_project(points3d, cam):
    s, tx, ty = cam[:, 0:1], cam[:, 1:2], cam[:, 2:3]
    return torch.stack([s * (points3d[:, :, 0] + tx), s * (points3d[:, :, 1] + ty)], dim=-1)


def joints2d_visibility(joints2d, img_wh=REGRESSOR_IMG_WH):
    """utils/joints2d_utils.py:23-33 (strict comparisons)."""
    vis = torch.ones(joints2d.shape[:2], dtype=torch.bool)
    vis[joints2d[:, :, 0] > img_wh] = 0
    vis[joints2d[:, :, 1] > img_wh] = 0
    vis[joints2d[:, :, 0] < 0] = 0
    vis[joints2d[:, :, 1] < 0] = 0
    return vis


# ----------------------------------------------------------------------------------------------
# SMPL (models/smpl_official.py:15-41 on top of the smplx restatement)
# ----------------------------------------------------------------------------------------------
class SmplOracle(object):
    """Loads `<assets>/additional/...` and evaluates the 90-joint SMPL forward on CPU."""

    def __init__(self, additional_dir, batch_size=1):
        if _SHIM not in sys.path:
            sys.path.insert(0, _SHIM)
        import smplx  # the shim
        self.smpl = smplx.SMPL(os.path.join(additional_dir, 'smpl'), batch_size=batch_size)
        ld = lambda n: torch.tensor(np.load(os.path.join(additional_dir, n)), dtype=torch.float32)
        self.J_extra = ld('J_regressor_extra.npy')              # smpl_official.py:17
        self.J_cocoplus = ld('cocoplus_regressor.npy')          # :18
        self.J_h36m = ld('J_regressor_h36m.npy')                # :19
        self.faces = self.smpl.faces_tensor

    def forward(self, betas=None, body_pose=None, global_orient=None, pose2rot=True, intermediates=None):
        from smplx.lbs import vertices2joints
        out = self.smpl(betas=betas, body_pose=body_pose, global_orient=global_orient, pose2rot=pose2rot)
        v = out.vertices
        joints = torch.cat([out.joints, vertices2joints(self.J_extra, v),          # :30-34
                            vertices2joints(self.J_cocoplus, v), vertices2joints(self.J_h36m, v)], dim=1)
        return v, joints

    def forward_rotmats(self, rotmats, betas):
        """rotmats [B,24,3,3] -> (vertices [B,6890,3], joints [B,90,3]); train/...:196-199."""
        return self.forward(betas=betas, body_pose=rotmats[:, 1:], global_orient=rotmats[:, 0:1], pose2rot=False)


# ----------------------------------------------------------------------------------------------
# loss (losses/multi_task_loss.py:73-119)
# ----------------------------------------------------------------------------------------------
LOSS_TASKS = ('verts', 'joints2D', 'joints3D', 'shape_params', 'pose_params')


def init_log_vars(init_loss_weights=None, eps=1e-6):
    """losses/multi_task_loss.py:30-44 -> dict task -> float32 scalar tensor."""
    out = {}
    for t in LOSS_TASKS:
        v = 0.0 if init_loss_weights is None else -np.log(init_loss_weights[t] + eps)
        out[t] = torch.tensor(v).float()
    return out


def multi_task_loss(labels, outputs, log_vars, losses_on=LOSS_TASKS, reduction='mean'):
    mse = lambda a, b: F.mse_loss(a, b, reduction=reduction)
    total, parts = 0.0, {}

    def add(task, raw):
        nonlocal total
        lv = log_vars[task]
        total = total + raw * torch.exp(-lv) + lv
        parts[task] = raw * torch.exp(-lv)

    if 'verts' in losses_on:
        add('verts', mse(outputs['verts'], labels['verts']))
    if 'joints2D' in losses_on:
        lab, pred = labels['joints2D'], outputs['joints2D']
        if 'vis' in labels:
            lab, pred = lab[labels['vis'], :], pred[labels['vis'], :]
        lab = (2.0 * lab) / REGRESSOR_IMG_WH - 1.0
        add('joints2D', mse(pred, lab))
    if 'joints3D' in losses_on:
        add('joints3D', mse(outputs['joints3D'], labels['joints3D']))
    if 'shape_params' in losses_on:
        add('shape_params', mse(outputs['shape_params'], labels['shape_params']))
    if 'pose_params' in losses_on:
        add('pose_params', mse(outputs['pose_params_rot_matrices'], labels['pose_params_rot_matrices']))
    return total, parts


# ----------------------------------------------------------------------------------------------
# the whole path (train/train_synthetic_otf_rendering.py:186-206 / predict/predict_3D.py:129-149)
# ----------------------------------------------------------------------------------------------
def regress_and_pose(x, sd, init_params, smpl, train=False, iters=3, stats_out=None):
    """x -> dict(cam, pose6d, shape, rotmats, vertices, joints, joints2d_coco, joints_h36mlsp)."""
    feat = encoder_forward(x, sd, train=train, stats_out=stats_out)
    p = ief_forward(feat, sd, init_params, iters)
    cam, pose6d, shape = split_params(p)
    R = rot6d_to_rotmat(pose6d.contiguous()).view(-1, 24, 3, 3)
    verts, joints = smpl.forward_rotmats(R, shape)
    j_h36m = joints[:, ALL_JOINTS_TO_H36M_MAP, :]
    j_coco = joints[:, ALL_JOINTS_TO_COCO_MAP, :]
    return {'feat': feat, 'params': p, 'cam': cam, 'pose6d': pose6d, 'shape': shape, 'rotmats': R,
            'vertices': verts, 'joints': joints, 'joints2d_coco': orthographic_project(j_coco, cam),
            'joints_h36mlsp': j_h36m[:, H36M_TO_J14, :]}


# ----------------------------------------------------------------------------------------------
# proxy-representation synthesis (utils/label_conversions.py:48-55, 90-127)
# ----------------------------------------------------------------------------------------------
def binary_labels(multiclass_labels):
    out = torch.zeros_like(multiclass_labels)
    out[multiclass_labels != 0] = 1
    return out


def joints2d_to_heatmaps(joints2d, img_wh, std=4):
    """numpy-style restatement of convert_2Djoints_to_gaussian_heatmaps_torch (integer window arithmetic spelled out)."""
    jr = joints2d.int()
    B, N = jr.shape[0], jr.shape[1]
    hm = torch.zeros((B, N, img_wh, img_wh), dtype=torch.float32)
    size = 2 * std
    x, y = torch.meshgrid(torch.linspace(-size, size, 2 * size), torch.linspace(-size, size, 2 * size), indexing='ij')
    d = torch.sqrt(x * x + y * y)
    g = torch.exp(-(d ** 2 / (2.0 * std ** 2)))
    for i in range(B):
        for j in range(N):
            cx, cy = int(jr[i, j, 0]), int(jr[i, j, 1])
            if cx > -size and cy > -size and cx < img_wh - 1 + size and cy < img_wh - 1 + size:
                hsx, hex_ = max(0, cx - size), min(img_wh - 1, cx + size)
                hsy, hey = max(0, cy - size), min(img_wh - 1, cy + size)
                gsx, gex = max(0, size - cx), min(2 * size, 2 * size - (size + cx - (img_wh - 1)))
                gsy, gey = max(0, size - cy), min(2 * size, 2 * size - (size + cy - (img_wh - 1)))
                hm[i, j, hsy:hey, hsx:hex_] = g[gsy:gey, gsx:gex]
    return hm


# ----------------------------------------------------------------------------------------------
# target side of the synthetic loop
# ----------------------------------------------------------------------------------------------
def batch_rodrigues(rot_vecs):
    """smplx.lbs.batch_rodrigues (restated in oracle/smplx_shim): [n,3] -> [n,3,3]."""
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    from smplx.lbs import batch_rodrigues as br
    return br(rot_vecs)


def perspective_project(points, rotation, translation, cam_K):
    """utils/cam_utils.py:56-71: rotate + translate, divide by depth, apply the intrinsics, drop the last coordinate."""
    q = torch.einsum('bij,bkj->bki', rotation, points) + translation.unsqueeze(1)
    q = q / q[:, :, -1].unsqueeze(-1)
    return torch.einsum('bij,bkj->bki', cam_K, q)[:, :, :-1]


def sample_shape_from_noise(noise, mean_shape, distribution, delta_betas_range=None, std_vector=None):
    """augmentation/smpl_augmentation.py:6-24 with the random draw passed in (`noise` = the rand / randn tensor)."""
    if distribution == 'uniform':
        l, h = delta_betas_range
        return ((h - l) * noise + l) + mean_shape
    return noise * std_vector + mean_shape


def cam_t_from_noise(mean_cam_t, noise_xy, noise_z, xy_std, delta_z_range):
    """augmentation/cam_augmentation.py:4-14 with the two draws (randn [bs,2], rand [bs]) passed in."""
    out = mean_cam_t.clone()
    out[:, :2] = mean_cam_t[:, :2] + noise_xy * xy_std
    l, h = delta_z_range
    out[:, 2] = mean_cam_t[:, 2] + ((h - l) * noise_z + l)
    return out


# ----------------------------------------------------------------------------------------------
# evaluation metrics (utils/eval_utils.py, metrics/train_loss_and_metrics_tracker.py:102-213)
# numpy float32 like the reference; the batch loop of the reference is vectorised with numpy's stacked SVD
# ----------------------------------------------------------------------------------------------
def procrustes_batch(S1, S2):
    """utils/eval_utils.py:7-60 for [B,N,3] arrays: S1 after the best similarity transform onto S2."""
    S1, S2 = np.asarray(S1), np.asarray(S2)
    A, Bm = np.swapaxes(S1, 1, 2), np.swapaxes(S2, 1, 2)                 # [B,3,N] as the reference works
    mu1, mu2 = A.mean(axis=2, keepdims=True), Bm.mean(axis=2, keepdims=True)
    X1, X2 = A - mu1, Bm - mu2
    var1 = np.sum(X1 ** 2, axis=(1, 2))
    K = X1 @ np.swapaxes(X2, 1, 2)
    U, s, Vh = np.linalg.svd(K)
    V = np.swapaxes(Vh, 1, 2)
    Z = np.tile(np.eye(3, dtype=K.dtype), (K.shape[0], 1, 1))
    Z[:, 2, 2] *= np.sign(np.linalg.det(U @ np.swapaxes(V, 1, 2)))
    R = V @ (Z @ np.swapaxes(U, 1, 2))
    scale = np.trace(R @ K, axis1=1, axis2=2) / var1
    t = mu2 - scale[:, None, None] * (R @ mu1)
    return np.swapaxes(scale[:, None, None] * (R @ A) + t, 1, 2)


def scale_translation_batch(P, T):
    """utils/eval_utils.py:63-85."""
    P, T = np.asarray(P), np.asarray(T)
    Pc = P - P.mean(axis=1, keepdims=True)
    p_scale = np.sqrt(np.sum(Pc ** 2, axis=(1, 2), keepdims=True) / P.shape[1])
    t_mean = T.mean(axis=1, keepdims=True)
    t_scale = np.sqrt(np.sum((T - t_mean) ** 2, axis=(1, 2), keepdims=True) / T.shape[1])
    return Pc / p_scale * t_scale + t_mean


def metric_sums(pred, target, img_wh=REGRESSOR_IMG_WH, pred_reposed=None, target_reposed=None):
    """The per-batch metric SUMS of metrics/train_loss_and_metrics_tracker.py:121-213 (all of them), as a dict of floats.
    pred / target: dicts of numpy arrays with the tracker's keys (verts, joints3D, joints2D, shape_params,
    pose_params_rot_matrices)."""
    l2 = lambda a, b: float(np.sum(np.linalg.norm(a - b, axis=-1)))
    out = {}
    for name, p, t in (('pves', pred['verts'], target['verts']), ('pve-ts', pred_reposed, target_reposed),
                       ('mpjpes', pred['joints3D'], target['joints3D'])):
        if p is None:
            continue
        out[name] = l2(p, t)
        out[name + '_sc'] = l2(scale_translation_batch(p, t), t)
        out[name + '_pa'] = l2(procrustes_batch(p, t), t)
    out['pose_mses'] = float(np.sum((pred['pose_params_rot_matrices'] - target['pose_params_rot_matrices']) ** 2))
    out['shape_mses'] = float(np.sum((pred['shape_params'] - target['shape_params']) ** 2))
    out['joints2D_l2es'] = l2((pred['joints2D'] + 1) * (img_wh / 2.0), target['joints2D'])
    return out

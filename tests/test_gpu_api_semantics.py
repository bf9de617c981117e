"""GPU: behaviours of the drop-in API that the reference's own loops rely on (round-1 review findings).

* gradients through the axis-angle input of SMPL (pose2rot=True, the default): `smpl_model(betas=pred_shape)` at
  train/train_synthetic_otf_rendering.py:206 runs with autograd enabled on the module's own axis-angle parameters;
* the training block train/...:186-232 replayed call for call;
* `regressor.train()` under torch.no_grad() uses batch statistics (nn.BatchNorm2d semantics, models/resnet.py:201-216);
* the module's `transl` parameter is applied on the autograd path and on the inference path alike;
* a second forward before backward() fails loudly instead of back-propagating through overwritten activations.
Tolerances: 1e-4 relative forward, 2e-4 gradients (as tests/test_gpu_train.py), 1e-5 for the small closed-form kernels."""
import os

import numpy as np
import pytest
import torch

import straps_oracle as O
from conftest import rel_err, RTOL, WEIGHT_SEED
from straps_b200 import synthetic_inputs

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GTOL = 2e-4


def _t(a, grad=False):
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).requires_grad_(grad)


def test_batch_rodrigues_backward_against_oracle_autograd():
    from straps_b200.autograd import BatchRodrigues
    rng = np.random.RandomState(0)
    r = rng.normal(0, 0.8, (500, 3)).astype(np.float32)
    r[:5] = 0.0                                    # the zero pose of the module's default parameters (angle = |1e-8| * sqrt(3))
    r[5] = (3.1, 0.0, 0.0)                         # close to pi
    g = rng.normal(0, 1, (500, 3, 3)).astype(np.float32)
    ro = _t(r, True)
    (O.batch_rodrigues(ro) * _t(g)).sum().backward()
    rg = _t(r).to(DEV).requires_grad_(True)
    R = BatchRodrigues.apply(rg)
    assert rel_err(R.detach().cpu().numpy(), O.batch_rodrigues(_t(r)).numpy()) < 1e-5
    (R * _t(g).to(DEV)).sum().backward()
    assert torch.isfinite(rg.grad).all()
    assert rel_err(rg.grad[5:].cpu().numpy(), ro.grad[5:].numpy()) < 1e-5
    # at exactly r = 0 both sides divide by 1.7e-8: compare loosely, finite is what matters for the unused default parameters
    assert rel_err(rg.grad[:5].cpu().numpy(), ro.grad[:5].numpy()) < 1e-2


@pytest.mark.parametrize('B', [2, 5])
def test_smpl_axis_angle_gradients_against_oracle_autograd(B, assets_root, smpl_oracle):
    import config
    from models.smpl_official import SMPL
    rng = np.random.RandomState(10 + B)
    betas = rng.normal(0, 1, (B, 10)).astype(np.float32)
    aa = rng.normal(0, 0.4, (B, 72)).astype(np.float32)
    gv = rng.normal(0, 1, (B, 6890, 3)).astype(np.float32)
    gj = rng.normal(0, 1, (B, 90, 3)).astype(np.float32)
    ao, bo = _t(aa, True), _t(betas, True)
    v, j = smpl_oracle.forward(betas=bo, body_pose=ao[:, 3:], global_orient=ao[:, :3], pose2rot=True)
    ((v * _t(gv)).sum() + (j * _t(gj)).sum()).backward()
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(DEV)
    ag, bg = _t(aa).to(DEV).requires_grad_(True), _t(betas).to(DEV).requires_grad_(True)
    out = smpl(body_pose=ag[:, 3:], global_orient=ag[:, :3], betas=bg)          # pose2rot=True is the default
    assert rel_err(out.vertices.detach().cpu().numpy(), v.detach().numpy()) < RTOL
    ((out.vertices * _t(gv).to(DEV)).sum() + (out.joints * _t(gj).to(DEV)).sum()).backward()
    assert rel_err(ag.grad.cpu().numpy(), ao.grad.numpy()) < GTOL
    assert rel_err(bg.grad.cpu().numpy(), bo.grad.numpy()) < GTOL


def test_smpl_betas_only_call_with_grad_enabled(assets_root, smpl_oracle):
    """train/...:206 `smpl_model(betas=pred_shape)`: default axis-angle parameters (requires_grad=True), grad mode on."""
    import config
    from models.smpl_official import SMPL
    B = 3
    rng = np.random.RandomState(3)
    betas = rng.normal(0, 1, (B, 10)).astype(np.float32)
    gv = rng.normal(0, 1, (B, 6890, 3)).astype(np.float32)
    bo = _t(betas, True)
    so = O.SmplOracle(os.path.join(assets_root, 'additional'), batch_size=B)
    v, _ = so.forward(betas=bo)
    (v * _t(gv)).sum().backward()
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(DEV)
    bg = _t(betas).to(DEV).requires_grad_(True)
    out = smpl(betas=bg)
    assert out.vertices.requires_grad
    assert rel_err(out.vertices.detach().cpu().numpy(), v.detach().numpy()) < RTOL
    (out.vertices * _t(gv).to(DEV)).sum().backward()
    assert rel_err(bg.grad.cpu().numpy(), bo.grad.numpy()) < GTOL
    assert smpl.body_pose.grad is not None and torch.isfinite(smpl.body_pose.grad).all()      # smplx builds this graph too


def test_transl_is_applied_on_both_paths(assets_root):
    import config
    from models.smpl_official import SMPL
    B = 2
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(DEV)
    with torch.no_grad():
        smpl.transl.copy_(torch.tensor([[0.1, -0.2, 0.3], [1.0, 2.0, -3.0]]))
    rng = np.random.RandomState(1)
    betas = _t(rng.normal(0, 1, (B, 10))).to(DEV)
    R = O.rot6d_to_rotmat(_t(rng.normal(0, 1, (B, 144)))).view(B, 24, 3, 3).to(DEV)
    with torch.no_grad():
        ref = smpl(body_pose=R[:, 1:], global_orient=R[:, :1], betas=betas, pose2rot=False)
    bg = betas.clone().requires_grad_(True)
    out = smpl(body_pose=R[:, 1:], global_orient=R[:, :1], betas=bg, pose2rot=False)
    assert out.vertices.requires_grad
    assert rel_err(out.vertices.detach().cpu().numpy(), ref.vertices.cpu().numpy()) < 1e-6
    assert rel_err(out.joints.detach().cpu().numpy(), ref.joints.cpu().numpy()) < 1e-6
    out.vertices.sum().backward()
    assert rel_err(smpl.transl.grad.cpu().numpy(), np.full((B, 3), 6890.0)) < 1e-6


def _regressor(C, sd, mode='f16x3_tc'):
    from models.regressor import SingleInputRegressor
    reg = SingleInputRegressor(C, 18, 3, conv_mode=mode)
    reg.load_state_dict(sd)
    return reg.to(DEV)


def test_train_mode_forward_under_no_grad_uses_batch_statistics(assets_root, additional_dir, smpl_oracle):
    C, B = 17, 4
    sd = O.make_regressor_state(C, seed=WEIGHT_SEED)
    init = O.load_initial_params(os.path.join(additional_dir, 'neutral_smpl_mean_params_6dpose.npz'))
    x = synthetic_inputs.make_proxy_batch(B, C, seed=13)
    stats = {}
    with torch.no_grad():
        o = O.regress_and_pose(_t(x), sd, init, smpl_oracle, train=True, stats_out=stats)
        o_eval = O.regress_and_pose(_t(x), sd, init, smpl_oracle, train=False)
    reg = _regressor(C, sd).train()
    with torch.no_grad():
        cam, pose, shape = reg(_t(x).to(DEV))
        feat = reg.image_encoder(_t(x).to(DEV))
    assert not cam.requires_grad
    assert rel_err(pose.cpu().numpy(), o['pose6d'].numpy()) < RTOL
    assert rel_err(shape.cpu().numpy(), o['shape'].numpy()) < RTOL
    assert rel_err(cam.cpu().numpy(), o['cam'].numpy()) < RTOL
    assert rel_err(pose.cpu().numpy(), o_eval['pose6d'].numpy()) > 10 * RTOL        # and it is NOT the running-statistics result
    assert feat.shape == (B, 512)
    new = reg.state_dict()
    assert int(new['image_encoder.bn1.num_batches_tracked']) == 2                   # two train-mode forwards
    assert int(new['image_encoder.layer4.1.bn2.num_batches_tracked']) == 2


def test_second_forward_before_backward_fails_loudly(assets_root):
    from straps_b200._lib import StrapsError
    C, B = 17, 2
    sd = O.make_regressor_state(C, seed=WEIGHT_SEED)
    reg = _regressor(C, sd).train()
    x = _t(synthetic_inputs.make_proxy_batch(B, C, seed=2)).to(DEV)
    cam1, pose1, shape1 = reg(x)
    cam2, pose2, shape2 = reg(x)
    with pytest.raises(StrapsError, match='overwritten'):
        pose1.sum().backward()
    pose2.sum().backward()                         # the latest forward still owns the workspace
    assert all(p.grad is not None for n, p in reg.named_parameters() if 'ief_layers' not in n)
    cam3, pose3, shape3 = reg(x)
    reg.eval()
    with torch.no_grad():
        reg(x)                                     # an inference forward overwrites the activations as well
    with pytest.raises(StrapsError, match='overwritten'):
        pose3.sum().backward()


def test_perspective_projection_refuses_silent_detach(assets_root):
    from straps_b200._lib import StrapsError
    from utils.cam_utils import perspective_project_torch
    B = 2
    pts = torch.randn(B, 5, 3, device=DEV, requires_grad=True)
    Rm = torch.eye(3, device=DEV).repeat(B, 1, 1)
    t = torch.tensor([[0., 0., 5.]], device=DEV).repeat(B, 1)
    with pytest.raises(StrapsError):
        perspective_project_torch(pts, Rm, t, focal_length=5000., img_wh=256)
    with torch.no_grad():
        out = perspective_project_torch(pts, Rm, t, focal_length=5000., img_wh=256)
    assert out.shape == (B, 5, 2)


def test_reference_training_block_verbatim(assets_root, additional_dir, smpl_oracle):
    """train/train_synthetic_otf_rendering.py:186-232 call for call (incl. the reposed `smpl_model(betas=pred_shape)` with grad
    enabled and `optimiser.zero_grad(); loss.backward(); optimiser.step()` with torch.optim.Adam over regressor + criterion)."""
    import config
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat
    from utils.cam_utils import orthographic_project_torch
    from utils.joints2d_utils import check_joints2d_visibility_torch
    from losses.multi_task_loss import HomoscedasticUncertaintyWeightedMultiTaskLoss as Loss
    C, B = 17, 4
    W = {'verts': 1.0, 'joints2D': 0.1, 'pose_params': 0.1, 'shape_params': 0.1, 'joints3D': 1.0}
    tasks = ['verts', 'joints2D', 'joints3D', 'pose_params', 'shape_params']
    sd = O.make_regressor_state(C, seed=WEIGHT_SEED)
    rng = np.random.RandomState(5)
    x = synthetic_inputs.make_proxy_batch(B, C, seed=21)
    t_betas = rng.normal(0, 1, (B, 10)).astype(np.float32)
    with torch.no_grad():
        t_R = O.rot6d_to_rotmat(_t(rng.normal(0, 1, (B, 144)))).view(B, 24, 3, 3)
        t_v, t_j = smpl_oracle.forward_rotmats(t_R, _t(t_betas))
    t_j2d = rng.uniform(-20, 276, (B, 17, 2)).astype(np.float32)
    device = DEV
    regressor = _regressor(C, sd)
    smpl_model = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(device)
    criterion = Loss(tasks, init_loss_weights=W).to(device)
    optimiser = torch.optim.Adam(list(regressor.parameters()) + list(criterion.parameters()), lr=1e-4)      # run_train.py:200-201
    before = {n: p.detach().clone() for n, p in regressor.named_parameters()}
    input = _t(x).to(device)
    target_glob_rotmats, target_pose_rotmats = t_R[:, :1].to(device), t_R[:, 1:].to(device)
    target_vertices, target_shape = t_v.to(device), _t(t_betas).to(device)
    target_joints_h36mlsp = t_j[:, O.ALL_JOINTS_TO_H36M_MAP][:, O.H36M_TO_J14].to(device)
    target_joints2d_coco = _t(t_j2d).to(device)

    regressor.train()
    pred_cam_wp, pred_pose, pred_shape = regressor(input)
    pred_pose_rotmats = rot6d_to_rotmat(pred_pose.contiguous()).view(-1, 24, 3, 3)
    pred_smpl_output = smpl_model(body_pose=pred_pose_rotmats[:, 1:],
                                  global_orient=pred_pose_rotmats[:, 0].unsqueeze(1),
                                  betas=pred_shape,
                                  pose2rot=False)
    pred_vertices = pred_smpl_output.vertices
    pred_joints_all = pred_smpl_output.joints
    pred_joints_h36m = pred_joints_all[:, config.ALL_JOINTS_TO_H36M_MAP, :]
    pred_joints_h36mlsp = pred_joints_h36m[:, config.H36M_TO_J14, :]
    pred_joints_coco = pred_joints_all[:, config.ALL_JOINTS_TO_COCO_MAP, :]
    pred_joints2d_coco = orthographic_project_torch(pred_joints_coco, pred_cam_wp)
    pred_reposed_smpl_output = smpl_model(betas=pred_shape)
    pred_reposed_vertices = pred_reposed_smpl_output.vertices
    target_pose_rotmats = torch.cat([target_glob_rotmats, target_pose_rotmats], dim=1)
    target_joints2d_vis_coco = check_joints2d_visibility_torch(target_joints2d_coco, config.REGRESSOR_IMG_WH)
    pred_dict_for_loss = {'joints2D': pred_joints2d_coco, 'verts': pred_vertices, 'shape_params': pred_shape,
                          'pose_params_rot_matrices': pred_pose_rotmats, 'joints3D': pred_joints_h36mlsp}
    target_dict_for_loss = {'joints2D': target_joints2d_coco, 'verts': target_vertices, 'shape_params': target_shape,
                            'pose_params_rot_matrices': target_pose_rotmats, 'joints3D': target_joints_h36mlsp,
                            'vis': target_joints2d_vis_coco}
    optimiser.zero_grad()
    loss, task_losses_dict = criterion(target_dict_for_loss, pred_dict_for_loss)
    loss.backward()
    optimiser.step()

    # the oracle's value of the same step
    init = O.load_initial_params(os.path.join(additional_dir, 'neutral_smpl_mean_params_6dpose.npz'))
    with torch.no_grad():
        o = O.regress_and_pose(_t(x), sd, init, smpl_oracle, train=True)
        labels_o = {'verts': t_v, 'joints2D': _t(t_j2d), 'joints3D': t_j[:, O.ALL_JOINTS_TO_H36M_MAP][:, O.H36M_TO_J14],
                    'shape_params': _t(t_betas), 'pose_params_rot_matrices': t_R, 'vis': O.joints2d_visibility(_t(t_j2d))}
        outs_o = {'verts': o['vertices'], 'joints2D': o['joints2d_coco'], 'joints3D': o['joints_h36mlsp'], 'shape_params': o['shape'],
                  'pose_params_rot_matrices': o['rotmats']}
        loss_o, _ = O.multi_task_loss(labels_o, outs_o, O.init_log_vars(W))
        so = O.SmplOracle(additional_dir, batch_size=B)
        rep_o, _ = so.forward(betas=o['shape'])
    assert rel_err(loss.detach().cpu().numpy(), loss_o.numpy()) < RTOL
    assert rel_err(pred_reposed_vertices.detach().cpu().numpy(), rep_o.numpy()) < RTOL
    assert pred_reposed_vertices.requires_grad
    moved = [n for n, p in regressor.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert len(moved) == len(before)            # Adam's first step moves every parameter by ~lr
    # and the next eval-mode forward sees the updated weights and running statistics
    regressor.eval()
    with torch.no_grad():
        cam_e, pose_e, shape_e = regressor(input)
    assert torch.isfinite(pose_e).all()

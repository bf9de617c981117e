"""GPU: a checkpoint in the reference's own format (tests/golden/checkpoint_ref.json pins it to the UNMODIFIED reference's run, see
oracle/checkpoint_oracle.py) resumes on the B200 path and the NEXT training step lands where the reference's would
(run_train.py:194-209, train/train_synthetic_otf_rendering.py:186-233)."""
import os

import numpy as np
import pytest
import torch

import straps_oracle as O
from conftest import SCRATCH, rel_err, RTOL

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_resume_reference_checkpoint_and_take_the_next_step(assets_root, additional_dir):
    import config
    import checkpoint_oracle as CK
    from test_checkpoint_cpu import _check_against_fixture
    from models.regressor import SingleInputRegressor
    from models.smpl_official import SMPL
    from utils.rigid_transform_utils import rot6d_to_rotmat
    from utils.cam_utils import orthographic_project_torch
    from utils.joints2d_utils import check_joints2d_visibility_torch
    from losses.multi_task_loss import HomoscedasticUncertaintyWeightedMultiTaskLoss as Loss
    from straps_b200.parallel import DataParallelAdam
    from utils.checkpoint_utils import resume_from_checkpoint
    ckpt, names = CK.build(additional_dir)
    _check_against_fixture(ckpt, names)
    after, _ = CK.build(additional_dir, steps=CK.STEPS + 1)            # where the reference lands one step later
    path = os.path.join(SCRATCH, 'reference_style_epoch3_gpu.tar')
    os.makedirs(SCRATCH, exist_ok=True)
    torch.save(ckpt, path)

    reg = SingleInputRegressor(CK.C_IN, 18, 3).to(DEV)
    crit = Loss(CK.LOSSES_ON, init_loss_weights=None).to(DEV)
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=CK.BATCH).to(DEV)
    opt = DataParallelAdam(list(reg.parameters()) + list(crit.parameters()), lr=1.0)
    resume_from_checkpoint(path, reg, opt, crit, map_location=DEV)
    opt.bucket.bump_versions()
    assert opt.step_count == CK.STEPS and len(opt.bucket.plist) == 69

    x, tg = CK.step_data(CK.STEPS, O.SmplOracle(additional_dir, batch_size=CK.BATCH))
    tg = {k: v.to(DEV) for k, v in tg.items()}
    tg['vis'] = check_joints2d_visibility_torch(tg['joints2D'], config.REGRESSOR_IMG_WH)
    reg.train()
    opt.zero_grad()
    cam, pose, shape = reg(x.to(DEV))
    R = rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
    out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
    outs = {'verts': out.vertices, 'joints2D': orthographic_project_torch(out.joints[:, config.ALL_JOINTS_TO_COCO_MAP, :], cam),
            'joints3D': out.joints[:, config.ALL_JOINTS_TO_H36M_MAP, :][:, config.H36M_TO_J14, :], 'shape_params': shape,
            'pose_params_rot_matrices': R}
    loss, parts = crit(tg, outs)
    assert set(parts) == set(CK.LOSSES_ON)
    loss.backward()
    opt.step()
    # Adam's update is lr * m / (sqrt(v) + eps): where a gradient element is ~0 the ratio is ill-conditioned, so the step is judged
    # in aggregate -- mean |update_gpu - update_reference| as a fraction of the learning rate -- and per tensor at 1e-4 of its value
    new = reg.state_dict()
    dev_sum, n = 0.0, 0
    for name in names[:66]:
        before, want, got = ckpt['model_state_dict'][name], after['model_state_dict'][name], new[name].cpu()
        assert rel_err(got.numpy(), want.numpy()) < RTOL, name
        dev_sum += float(((got - before) - (want - before)).abs().sum())
        n += got.numel()
    assert dev_sum / n < 0.02 * CK.LR, dev_sum / n / CK.LR
    for t in CK.CRITERION_ORDER:
        got, want = getattr(crit, t + '_log_var').detach().cpu(), after['criterion_state_dict'][t + '_log_var']
        if t in CK.LOSSES_ON:
            assert abs(float(got) - float(want)) < 0.02 * CK.LR, t
        else:
            assert torch.equal(got, want), t                         # frozen: untouched by the optimiser
    for k in ('image_encoder.bn1.running_mean', 'image_encoder.layer4.1.bn2.running_var'):
        assert rel_err(new[k].cpu().numpy(), after['model_state_dict'][k].numpy()) < 1e-4, k
    assert int(new['image_encoder.bn1.num_batches_tracked']) == CK.STEPS + 1

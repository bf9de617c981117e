"""ORACLE / TEST INFRASTRUCTURE -- tests/golden/checkpoint_ref.json from the UNMODIFIED reference (authoring container only).

    python oracle/gen_checkpoint_fixture.py

Runs the reference's own SingleInputRegressor + SMPL (on the smplx shim) + HomoscedasticUncertaintyWeightedMultiTaskLoss with a
SUBSET of losses switched on + torch.optim.Adam(list(regressor.parameters()) + list(criterion.parameters()), lr=1e-4)
(run_train.py:194-201) for checkpoint_oracle.STEPS training steps exactly as train/train_synthetic_otf_rendering.py:186-233 does,
builds the save dict of train/...:365-377, and
  1. asserts that oracle/checkpoint_oracle.build() reproduces EVERY tensor of it (so the tests may rebuild the checkpoint without
     the reference), and that the reference's parameter order is the one checkpoint_oracle assumes;
  2. writes the checkpoint's metadata (structure + float64 checksums) as the committed fixture.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, 'straps-3dhumanshapepose_b200'))

import ref_harness                      # noqa: E402
import straps_oracle as O               # noqa: E402
import checkpoint_oracle as CK          # noqa: E402
from straps_b200 import synthetic_assets   # noqa: E402


def main():
    root = os.path.join(REPO, 'tests', '_scratch', 'assets')
    synthetic_assets.write_synthetic_assets(root, seed=0)
    add = os.path.join(root, 'additional')
    ref = ref_harness.load_reference(root)
    smpl_oracle = O.SmplOracle(add, batch_size=CK.BATCH)
    sd = O.make_regressor_state(CK.C_IN, seed=CK.WEIGHT_SEED)
    with ref.cwd():
        regressor = ref.SingleInputRegressor(CK.C_IN, 18, 3)
        regressor.load_state_dict(sd)
        smpl_model = ref.SMPL(ref.config.SMPL_MODEL_DIR, batch_size=CK.BATCH)
        criterion = ref.Loss(CK.LOSSES_ON, init_loss_weights=CK.LOSS_WEIGHTS)
        params = list(regressor.parameters()) + list(criterion.parameters())
        optimiser = torch.optim.Adam(params, lr=CK.LR)
        regressor.train()
        for step in range(CK.STEPS):
            x, tg = CK.step_data(step, smpl_oracle)
            pred_cam_wp, pred_pose, pred_shape = regressor(x)
            pred_pose_rotmats = ref.rot6d_to_rotmat(pred_pose.contiguous()).view(-1, 24, 3, 3)
            out = smpl_model(body_pose=pred_pose_rotmats[:, 1:], global_orient=pred_pose_rotmats[:, 0].unsqueeze(1), betas=pred_shape,
                             pose2rot=False)
            j_all = out.joints
            pred = {'joints2D': ref.orthographic_project_torch(j_all[:, ref.config.ALL_JOINTS_TO_COCO_MAP, :], pred_cam_wp),
                    'verts': out.vertices, 'shape_params': pred_shape, 'pose_params_rot_matrices': pred_pose_rotmats,
                    'joints3D': j_all[:, ref.config.ALL_JOINTS_TO_H36M_MAP, :][:, ref.config.H36M_TO_J14, :]}
            tg['vis'] = ref.check_joints2d_visibility_torch(tg['joints2D'], ref.config.REGRESSOR_IMG_WH)
            optimiser.zero_grad()
            loss, _ = criterion(tg, pred)
            loss.backward()
            optimiser.step()
        ref_names = [n for n, _ in regressor.named_parameters()] + [n for n, _ in criterion.named_parameters()]
        ckpt = {'epoch': 3, 'best_epoch': 2, 'best_epoch_val_metrics': {'pves': 0.123, 'mpjpes_pa': 0.045},
                'model_state_dict': regressor.state_dict(), 'best_model_state_dict': regressor.state_dict(),
                'optimiser_state_dict': optimiser.state_dict(), 'criterion_state_dict': criterion.state_dict()}
    mine, names = CK.build(add)
    assert ref_names == names, 'parameter order differs from the reference: %s' % [(a, b) for a, b in zip(ref_names, names) if a != b][:5]
    assert set(mine['model_state_dict']) == set(ckpt['model_state_dict'])
    worst = 0.0
    for k, v in ckpt['model_state_dict'].items():
        assert torch.equal(v, mine['model_state_dict'][k]), k
    for k, v in ckpt['criterion_state_dict'].items():
        assert torch.equal(v, mine['criterion_state_dict'][k]), k
    a, b = ckpt['optimiser_state_dict'], mine['optimiser_state_dict']
    assert sorted(a['state']) == sorted(b['state']) and a['param_groups'] == b['param_groups'], (sorted(a['state']), sorted(b['state']))
    for i, st in a['state'].items():
        for f in ('exp_avg', 'exp_avg_sq'):
            assert torch.equal(st[f], b['state'][i][f]), (i, f)
        assert float(st['step']) == float(b['state'][i]['step'])
    meta = CK.metadata(ckpt)
    meta['param_names'] = ref_names
    meta['generator'] = 'oracle/gen_checkpoint_fixture.py: reference modules from /root/reference, %d steps, B=%d, losses_on=%s' % (
        CK.STEPS, CK.BATCH, CK.LOSSES_ON)
    out = os.path.join(REPO, 'tests', 'golden', 'checkpoint_ref.json')
    json.dump(meta, open(out, 'w'), indent=0, sort_keys=True)
    frozen = [i for i, n in enumerate(ref_names) if str(i) not in meta['optimiser_state']]
    print('reference checkpoint reproduced bit for bit by checkpoint_oracle.build(); %d tensors in the optimiser, no state for %s'
          % (len(ref_names), [ref_names[i] for i in frozen]))
    print(out, os.path.getsize(out), 'bytes')


if __name__ == '__main__':
    main()

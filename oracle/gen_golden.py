"""ORACLE / TEST INFRASTRUCTURE -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):   python oracle/gen_golden.py
The reference's own modules (models.regressor.SingleInputRegressor, models.smpl_official.SMPL on top of
oracle/smplx_shim, utils.*, losses.*) are imported through oracle/ref_harness.py and evaluated on seeded
synthetic assets / weights / inputs.  Inputs and weights are NOT stored (they are regenerated from the same
numpy seeds by the tests); float64 checksums of them are, so a drifted generator is detected.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, 'straps-3dhumanshapepose_b200'))

import ref_harness                      # noqa: E402
import straps_oracle as O               # noqa: E402
from golden_inputs import (synth_inputs, metrics_inputs, SYNTH_STD, SYNTH_RANGE, SYNTH_XY_STD, SYNTH_Z_RANGE,   # noqa: E402
                           ALL_METRICS, ALL_TASKS)
from straps_b200 import synthetic_assets, synthetic_inputs   # noqa: E402

GOLDEN = os.path.join(REPO, 'tests', 'golden')
ASSET_SEED, WEIGHT_SEED, INPUT_SEED = 0, 1, 3
LOSS_WEIGHTS = {'verts': 1.0, 'joints2D': 0.1, 'pose_params': 0.1, 'shape_params': 0.1, 'joints3D': 1.0}   # run_train.py:53-54


def smpl_inputs(batch, seed):
    """Config-1 style inputs: betas ~ N(0,1), rotations from random 6-D vectors, and axis-angle poses."""
    rng = np.random.RandomState(seed)
    betas = rng.normal(0, 1, (batch, 10)).astype(np.float32)
    pose6d = rng.normal(0, 1, (batch, 144)).astype(np.float32)
    aa = rng.normal(0, 0.4, (batch, 72)).astype(np.float32)
    return betas, pose6d, aa


def checksum(a):
    return float(np.asarray(a, dtype=np.float64).sum())


def synth_fixture(ref):
    I = {k: torch.from_numpy(v) for k, v in synth_inputs().items()}
    B = I['pose_aa'].shape[0]
    out = {'in_checksum': np.array([checksum(v.numpy()) for v in I.values()])}
    std_vec = torch.tensor([SYNTH_STD] * 10)
    for dist, seed in (('normal', 5), ('uniform', 6)):
        params = {'augment_shape': True, 'delta_betas_distribution': dist, 'delta_betas_range': SYNTH_RANGE,
                  'delta_betas_std_vector': std_vec}
        torch.manual_seed(seed)
        shape, pose_rm, glob_rm = ref.smpl_augmentation.augment_smpl(I['orig_shape'], I['pose_aa'][:, 3:], I['pose_aa'][:, :3],
                                                                     I['mean_shape'], params)
        torch.manual_seed(seed)            # the draw the call above consumed
        noise = torch.randn(B, 10) if dist == 'normal' else torch.rand(B, 10)
        out.update({'shape_%s' % dist: shape.numpy(), 'noise_%s' % dist: noise.numpy()})
    out.update(pose_rotmats=pose_rm.numpy(), glob_rotmats=glob_rm.numpy())
    torch.manual_seed(7)
    cam_t = ref.cam_augmentation.augment_cam_t(I['mean_cam_t'], xy_std=SYNTH_XY_STD, delta_z_range=SYNTH_Z_RANGE)
    torch.manual_seed(7)
    out.update(aug_cam_t=cam_t.numpy(), noise_xy=torch.randn(B, 2).numpy(), noise_z=torch.rand(B).numpy())
    K = torch.from_numpy(ref.get_intrinsics_matrix(256, 256, 5000.).astype(np.float32))[None].expand(B, -1, -1)
    out['proj'] = ref.perspective_project_torch(I['points'], I['cam_R'], I['cam_t'], cam_K=K).numpy()
    out['proj_default_K'] = ref.perspective_project_torch(I['points'], I['cam_R'], I['cam_t'], focal_length=5000., img_wh=256).numpy()
    return out


def metrics_fixture(ref):
    pred, target, pr, tr_, losses = metrics_inputs()
    B = pred['verts'].shape[0]
    tracker = ref.Tracker(ALL_TASKS, ALL_METRICS, 256, os.path.join(REPO, 'tests', '_scratch', 'ref_tracker_log.pkl'))
    tracker.initialise_loss_metric_sums()
    t = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}
    for split in ('train', 'val'):
        tracker.update_per_batch(split, torch.tensor(losses['total']), {k: torch.tensor(losses[k]) for k in ALL_TASKS}, t(pred), t(target),
                                 B, pred_reposed_vertices=torch.from_numpy(pr), target_reposed_vertices=torch.from_numpy(tr_))
    sums = dict(tracker.loss_metric_sums)
    tracker.update_per_epoch()
    out = {'in_checksum': np.array([checksum(pred['verts']), checksum(target['verts']), checksum(pr), checksum(tr_),
                                    checksum(pred['joints3D']), checksum(target['joints3D'])]),
           'sum_keys': np.array(sorted(sums)), 'sum_values': np.array([float(sums[k]) for k in sorted(sums)], dtype=np.float64),
           'history_keys': np.array(sorted(tracker.history)),
           'history_values': np.array([tracker.history[k][-1] if tracker.history[k] else np.nan for k in sorted(tracker.history)])}
    out['joints3D_pa'] = ref.eval_utils.procrustes_analysis_batch(pred['joints3D'], target['joints3D'])
    out['joints3D_sc'] = ref.eval_utils.scale_and_translation_transform_batch(pred['joints3D'], target['joints3D'])
    out['verts0_pa_first32'] = ref.eval_utils.compute_similarity_transform(pred['verts'][0], target['verts'][0])[:32]
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    root = os.path.join(REPO, 'tests', '_scratch', 'assets')
    synthetic_assets.write_synthetic_assets(root, seed=ASSET_SEED)
    ref = ref_harness.load_reference(root)
    torch.manual_seed(0)

    # ---- SMPL forward only (BASELINE config 1: B=4) -------------------------------------------------
    B = 4
    betas, pose6d, aa = smpl_inputs(B, 11)
    with ref.cwd(), torch.no_grad():
        smpl = ref.SMPL(ref.config.SMPL_MODEL_DIR, batch_size=B)
        R = ref.rot6d_to_rotmat(torch.from_numpy(pose6d)).view(-1, 24, 3, 3)
        out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=torch.from_numpy(betas), pose2rot=False)
        out_aa = smpl(body_pose=torch.from_numpy(aa[:, 3:]), global_orient=torch.from_numpy(aa[:, :3]),
                      betas=torch.from_numpy(betas))
        out_t = smpl(betas=torch.from_numpy(betas))          # T-pose call, train/...:206
    np.savez_compressed(os.path.join(GOLDEN, 'smpl_b4.npz'),
                        in_checksum=np.array([checksum(betas), checksum(pose6d), checksum(aa)]),
                        rotmats=R.numpy(), vertices=out.vertices.numpy(), joints=out.joints.numpy(),
                        vertices_aa=out_aa.vertices.numpy(), joints_aa=out_aa.joints.numpy(),
                        vertices_tpose=out_t.vertices.numpy(), joints_tpose=out_t.joints.numpy(),
                        faces=smpl.faces_tensor.numpy(), parents=smpl.parents.numpy())

    # ---- encoder + IEF + SMPL (BASELINE config 2 at B=2), C = 17 and 18 ------------------------------
    for C in (17, 18):
        B = 2
        sd = O.make_regressor_state(C, seed=WEIGHT_SEED)
        x = synthetic_inputs.make_proxy_batch(B, C, seed=INPUT_SEED)
        with ref.cwd(), torch.no_grad():
            reg = ref.SingleInputRegressor(C, 18, 3)
            reg.load_state_dict(sd)
            reg.eval()
            feat = reg.image_encoder(torch.from_numpy(x))
            cam, pose, shape = reg(torch.from_numpy(x))
            R = ref.rot6d_to_rotmat(pose.contiguous()).view(-1, 24, 3, 3)
            smpl = ref.SMPL(ref.config.SMPL_MODEL_DIR, batch_size=B)
            out = smpl(body_pose=R[:, 1:], global_orient=R[:, 0].unsqueeze(1), betas=shape, pose2rot=False)
            j_coco = out.joints[:, ref.config.ALL_JOINTS_TO_COCO_MAP, :]
            j2d = ref.orthographic_project_torch(j_coco, cam)
            j_lsp = out.joints[:, ref.config.ALL_JOINTS_TO_H36M_MAP, :][:, ref.config.H36M_TO_J14, :]
        wsum = checksum(np.concatenate([v.numpy().astype(np.float64).ravel() for k, v in sorted(sd.items())]))
        np.savez_compressed(os.path.join(GOLDEN, 'regressor_c%d_b2.npz' % C),
                            in_checksum=np.array([checksum(x), wsum]), feat=feat.numpy(), cam=cam.numpy(),
                            pose6d=pose.numpy(), shape=shape.numpy(), rotmats=R.numpy(), joints=out.joints.numpy(),
                            vertices=out.vertices.numpy(), joints2d_coco=j2d.numpy(), joints_h36mlsp=j_lsp.numpy())

    # ---- loss (losses/multi_task_loss.py) -----------------------------------------------------------
    rng = np.random.RandomState(21)
    B = 3
    mk = lambda *s: torch.from_numpy(rng.normal(0, 1, s).astype(np.float32))
    outputs = {'verts': mk(B, 6890, 3), 'joints2D': mk(B, 17, 2), 'joints3D': mk(B, 14, 3), 'shape_params': mk(B, 10),
               'pose_params_rot_matrices': mk(B, 24, 3, 3)}
    labels = {'verts': mk(B, 6890, 3), 'joints2D': torch.from_numpy(rng.uniform(-40, 300, (B, 17, 2)).astype(np.float32)),
              'joints3D': mk(B, 14, 3), 'shape_params': mk(B, 10), 'pose_params_rot_matrices': mk(B, 24, 3, 3)}
    labels['vis'] = ref.check_joints2d_visibility_torch(labels['joints2D'], 256)
    crit = ref.Loss(['verts', 'joints2D', 'joints3D', 'shape_params', 'pose_params'], init_loss_weights=LOSS_WEIGHTS)
    total, parts = crit(labels, outputs)
    np.savez_compressed(os.path.join(GOLDEN, 'loss_b3.npz'), total=total.detach().numpy(),
                        vis=labels['vis'].numpy(), log_vars=np.array([p.item() for p in crit.parameters()]),
                        **{'part_' + k: v.detach().numpy() for k, v in parts.items()})
    # ---- N1: joint heat-maps (utils/label_conversions.py:90-127) incl. border / out-of-range joints ----------
    rng = np.random.RandomState(31)
    j = rng.uniform(-12, 270, (4, 17, 2)).astype(np.float32)
    j[0, 0] = [0., 0.]; j[0, 1] = [255., 255.]; j[0, 2] = [-7.9, 262.9]; j[0, 3] = [-8., 100.]; j[0, 4] = [263., 5.]
    j[0, 5] = [247.5, 8.2]; j[0, 6] = [254.99, 0.5]
    np.savez_compressed(os.path.join(GOLDEN, 'heatmaps_b4.npz'), joints2d=j, heatmaps=ref.heatmaps(torch.from_numpy(j), 256).numpy())
    # ---- N2: SMPL / camera augmentation + perspective projection (augmentation/*.py, utils/cam_utils.py:40-71) ----------
    np.savez_compressed(os.path.join(GOLDEN, 'synth_b4.npz'), **synth_fixture(ref))
    # ---- N4: metric sums of the tracker + aligned point sets (metrics/..tracker.py:102-213, utils/eval_utils.py) --------
    np.savez_compressed(os.path.join(GOLDEN, 'metrics_b4.npz'), **metrics_fixture(ref))
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == '__main__':
    main()

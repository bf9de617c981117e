"""SURVEY 8f rows N2 / N4 timed on one GPU: the target side of the synthetic loop and the device-side metric tracker,
next to the reference's own way of doing the same work (oracle restatement of its numpy / torch-CPU code on the host cores).

Usage: python tools/bench_widening.py [--batch 64] [--iters 50]      -> JSON lines
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'straps-3dhumanshapepose_b200'))
sys.path.insert(0, os.path.join(REPO, 'oracle'))          # checker / CPU baseline only
os.environ.setdefault('STRAPS_ASSETS_ROOT', os.path.join(REPO, 'tests', '_scratch', 'assets'))
import numpy as np
import torch
from straps_b200 import synthetic_assets, _lib

synthetic_assets.write_synthetic_assets(os.environ['STRAPS_ASSETS_ROOT'], seed=0)
import config
from models.smpl_official import SMPL
from augmentation.smpl_augmentation import augment_smpl
from augmentation.cam_augmentation import augment_cam_t
from utils.cam_utils import perspective_project_torch
from metrics.train_loss_and_metrics_tracker import TrainingLossesAndMetricsTracker
import straps_oracle as O

DEV = 'cuda:0'
METRICS = ['pves', 'pves_sc', 'pves_pa', 'pve-ts', 'pve-ts_sc', 'mpjpes', 'mpjpes_sc', 'mpjpes_pa', 'pose_mses', 'shape_mses',
           'joints2D_l2es']                                  # run_train.py:63-65
TASKS = ['verts', 'joints2D', 'pose_params', 'shape_params', 'joints3D']


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=50)
    args = ap.parse_args()
    B = args.batch
    rng = np.random.RandomState(0)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)
    pose, shape0 = cu(rng.normal(0, 0.4, (B, 72))), cu(rng.normal(0, 1, (B, 10)))
    mean_shape = cu(np.zeros(10))
    params = {'augment_shape': True, 'delta_betas_distribution': 'normal', 'delta_betas_range': [-3., 3.],
              'delta_betas_std_vector': torch.tensor([1.5] * 10, device=DEV)}
    mean_cam_t = torch.tensor([0., 0.2, 42.], device=DEV)[None].expand(B, -1)
    cam_R = torch.eye(3, device=DEV)[None].expand(B, -1, -1).contiguous()
    K = torch.tensor([[5000., 0., 128.], [0., 5000., 128.], [0., 0., 1.]], device=DEV)[None].expand(B, -1, -1).contiguous()
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=B).to(DEV)
    out = {}

    def target_side():
        with torch.no_grad():
            shape, pose_rm, glob_rm = augment_smpl(shape0, pose[:, 3:], pose[:, :3], mean_shape, params)
            cam_t = augment_cam_t(mean_cam_t, xy_std=0.05, delta_z_range=[-5, 5])
            o = smpl(body_pose=pose_rm, global_orient=glob_rm, betas=shape, pose2rot=False)
            j2d = perspective_project_torch(o.joints[:, config.ALL_JOINTS_TO_COCO_MAP, :], cam_R, cam_t, cam_K=K)
            rep = smpl(betas=shape).vertices
        out.update(verts=o.vertices, joints=o.joints, j2d=j2d, reposed=rep, shape=shape, rot=torch.cat([glob_rm, pose_rm], 1))
    n0 = _lib.launch_count()
    ms = timed(target_side, args.iters)
    launches = (_lib.launch_count() - n0) / (args.iters + 3)
    print(json.dumps({'workload': 'N2 target side: augment_smpl + augment_cam_t + SMPL(rotmats) + perspective + SMPL(T-pose) '
                                  '(train/train_synthetic_otf_rendering.py:121-145)', 'batch': B, 'ms': ms,
                      'bodies_per_s': B / (ms * 1e-3), 'library_launches_per_call': launches}))

    # ---- N4: one tracker update (all run_train.py metrics) ----
    noise = lambda t, s: t + s * torch.randn_like(t)
    pred = {'verts': noise(out['verts'], 0.02), 'joints3D': noise(out['joints'][:, 73:87], 0.02), 'joints2D': torch.rand(B, 17, 2, device=DEV) * 2 - 1,
            'shape_params': noise(out['shape'], 0.3), 'pose_params_rot_matrices': noise(out['rot'], 0.1)}
    target = {'verts': out['verts'], 'joints3D': out['joints'][:, 73:87].contiguous(), 'joints2D': out['j2d'], 'shape_params': out['shape'],
              'pose_params_rot_matrices': out['rot']}
    pred_rep = noise(out['reposed'], 0.01)
    loss = torch.tensor(1.0, device=DEV)
    tl = {t: torch.tensor(0.2, device=DEV) for t in TASKS}
    tracker = TrainingLossesAndMetricsTracker(TASKS, METRICS, 256, os.path.join(REPO, 'tests', '_scratch', 'bench_tracker.pkl'))
    tracker.initialise_loss_metric_sums()
    n0 = _lib.launch_count()
    ms = timed(lambda: tracker.update_per_batch('train', loss, tl, pred, target, B, pred_reposed_vertices=pred_rep,
                                                target_reposed_vertices=out['reposed']), args.iters)
    launches = (_lib.launch_count() - n0) / (args.iters + 3)
    sums = tracker.sync()
    # the reference's way: copy everything to the host, align sample by sample in numpy
    t0 = time.perf_counter()
    npd = lambda d: {k: v.cpu().numpy() for k, v in d.items()}
    ref = O.metric_sums(npd(pred), npd(target), pred_reposed=pred_rep.cpu().numpy(), target_reposed=out['reposed'].cpu().numpy())
    cpu_ms = (time.perf_counter() - t0) * 1e3
    n_updates = args.iters + 3
    worst = max(abs(sums['train_' + k] / n_updates - v) / abs(v) for k, v in ref.items() if k in METRICS)
    bytes_ = 2 * B * (2 * 6890 + 14) * 12
    print(json.dumps({'workload': 'N4 tracker.update_per_batch, 11 metrics of run_train.py:63-65 (device sums, no host sync)', 'batch': B,
                      'ms': ms, 'library_launches_per_call': launches, 'point_set_bytes_read_once': bytes_,
                      'achieved_gbs_of_that': bytes_ / (ms * 1e-3) / 1e9,
                      'cpu_numpy_ms (D2H + oracle restatement of the reference loop, vectorised SVD)': cpu_ms,
                      'max_rel_diff_vs_numpy': worst}))


if __name__ == '__main__':
    main()
